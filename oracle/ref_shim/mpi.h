// TEST INFRASTRUCTURE ONLY (oracle/): thread-rank stand-in for <mpi.h>.
//
// The reference (medical-genomics-group/gmrm) needs MPI, which this image does not
// have.  This header declares exactly the 20 MPI entry points the reference calls
// (SURVEY.md section 8c lists them); oracle/ref_shim/mpi_shim.cpp implements them
// with every "rank" being a std::thread of ONE process, so the UNMODIFIED reference
// sources under /root/reference/src can be run with R > 1 ranks.  Nothing here is
// shipped in, linked into, or called by the product (gmrm_b200/).
#pragma once
#include <cstddef>
#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_File;
typedef long long MPI_Offset;
typedef struct { int count; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0

#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 2
#define MPI_C_BOOL 3
#define MPI_INT 4
#define MPI_INTEGER 5
#define MPI_UNSIGNED 6
#define MPI_DOUBLE 7
#define MPI_UNSIGNED_LONG_LONG 8

#define MPI_SUM 1
#define MPI_MAX 2

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_EXCL 8

extern "C++" {
int MPI_Init(int*, char***);
int MPI_Finalize();
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Barrier(MPI_Comm);
double MPI_Wtime();
int MPI_Abort(MPI_Comm, int);
int MPI_Type_size(MPI_Datatype, int*);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File*);
int MPI_File_close(MPI_File*);
int MPI_File_delete(const char*, MPI_Info);
int MPI_File_get_size(MPI_File, MPI_Offset*);
int MPI_File_read_at(MPI_File, MPI_Offset, void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_read_at_all(MPI_File, MPI_Offset, void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_write_at(MPI_File, MPI_Offset, const void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_write_at_all(MPI_File, MPI_Offset, const void*, int, MPI_Datatype, MPI_Status*);
}
