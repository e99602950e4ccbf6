// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime.h> when gmrm_b200/csrc/kernels.cuh is compiled for the host
// by the kernel emulation tests (tests/emu/cuda_emu.h).  Only the one type the declarations need.
#pragma once
typedef void* cudaStream_t;
