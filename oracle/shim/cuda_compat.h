/* Test / baseline infrastructure (NOT product code): lets nvcc 12.9 compile the reference's BackendCUDA.cu (CUDA 6.5 era)
 * for sm_100a without editing it: the mask-less warp shuffle it uses was removed with Volta's independent thread scheduling. */
#pragma once
#define __shfl_xor(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
