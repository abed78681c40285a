// Shared helpers of libtefem_b200: error plumbing, launch checks, block reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tefem.h"

#ifdef __CUDACC__
#define TEFEM_HD __host__ __device__ __forceinline__
#else
#define TEFEM_HD inline
#endif

namespace tefem {

void set_error(const char* fmt, ...);
void count_launch();

#define TEFEM_CUDA_CHECK(expr)                                                           \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            tefem::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                             cudaGetErrorString(_e));                                    \
            return TEFEM_ERR_CUDA;                                                       \
        }                                                                                \
    } while (0)

// every kernel launch of the library is followed by exactly one TEFEM_LAUNCH_CHECK(): it also feeds the
// launch counter that bench.py reports as "gpu_launches".
#define TEFEM_LAUNCH_CHECK()                      \
    do {                                          \
        tefem::count_launch();                    \
        TEFEM_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

#define TEFEM_REQUIRE(cond, msg)                                                         \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            tefem::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);   \
            return TEFEM_ERR_INVALID;                                                    \
        }                                                                                \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Quadrature tables of the reference (elements.py:65-69, 76-87, 165-172), built on the host
// with the reference's own operation order; uploaded to constant memory by each translation
// unit that needs them.
struct QuadTables {
    double W4;       // w + w + w + w
    double S[4];     // sum_g w N_a(g)
    double NN[16];   // sum_g w N_a(g) N_b(g)
};
const QuadTables& quad_tables();

#ifdef __CUDACC__
// Sum over the CTA in a FIXED order (warp shuffle tree, then warp 0 over the warp partials):
// deterministic for a given block size.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double* smem /* >= NV*32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = (lane < nwarp) ? smem[k * 32 + lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
            v[k] = t;
        }
    }
    __syncthreads();
}
#endif

}  // namespace tefem
