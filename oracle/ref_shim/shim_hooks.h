// TEST INFRASTRUCTURE ONLY (oracle/).  Hooks shared by the MPI shim and the Boost
// stand-in headers: which thread-rank is calling, and the per-rank variate log that
// makes "replay mode" possible.
#pragma once
#include <cstddef>

extern "C" {
int  gmrm_shim_rank();                                  // rank of the calling thread
void gmrm_shim_log(const void* bytes, size_t nbytes);   // append to this rank's variate log
}

// Variate-log record kinds (1 tag byte, then the payload, all little-endian):
//   'B'  double a, b, value                    beta_rng           (distributions.hpp:39-46)
//   'N'  double mean, sd, z                    norm_rng: value = mean + sd*z (distributions.hpp:48-53)
//   'U'  double u                              unif_rng           (distributions.hpp:55-59)
//   'G'  double shape, scale, unit             rgamma: value = unit*scale   (distributions.hpp:32-37)
//   'P'  int32 n, int32 perm[n]                midx after random_shuffle    (phenotype.cpp:314-323)
