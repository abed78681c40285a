// Three-level aggregation preconditioner for the block-Jacobi-scaled system (right preconditioner of BiCGSTAB).
//
// The reference solves with a sparse direct solver (spsolve, solvers.py:26,172); the Krylov replacement with the
// per-node block-Jacobi scaling alone needs ~10 iterations per mesh cell along the longest axis (SURVEY.md H5:
// ~2000 at config B).  This adds a coarse-space correction in the additive (BPX-like) form
//
//       z = omega * r  +  P1 * M1( P1^T r )
//
// on the ALREADY block-Jacobi-scaled fine operator A^ = D^-1 A, where
//   * P1 is the prolongation of a geometric aggregation of the nodes whose coarse space holds the RIGID-BODY MODES
//     of every aggregate - 3 translations and 3 rotations about the aggregate's centre for the displacement, a
//     constant for theta: a fine node i of aggregate I at offset d_i from the centre receives
//     u_i = t_I + w_I x d_i, theta_i = theta_I.  Each aggregate is stored as two 4-DOF "pseudo nodes"
//     (t_x, t_y, t_z, theta) and (w_x, w_y, w_z, 0), so that level 1 is a 4x4-block CSR like the fine level;
//     A1 = P1^T A^ P1 (Galerkin, built on the host side with torch) is scaled by the inverse of its 8x8 diagonal
//     blocks, A1^ = D1^-1 A1, and stored in single precision (preconditioner data only; every vector and the
//     fine operator stay fp64).  Measured (scipy prototype, config-2 operator, 32^3 box): 99 iterations with
//     translations only, 66 with the rotations;
//   * M1 is one V-cycle on level 1: damped Jacobi on A1^ (= a scaling by omega_c), a level-2 correction
//     P2 * inv(P2^T A1^ P2) * P2^T with the same rigid-body transfer between aggregate centres (P2, P2^T stored
//     as small block-CSR matrices) and a DENSE inverse (a few hundred aggregates), and a post-smoothing step;
//   * the fine level needs NO extra SpMV: per application one restriction pass and one prolongation pass over
//     the fine vector, the level-1 work is ~1/100 of a fine SpMV.
// Every step is a fixed linear operator (fixed sweeps, no inner Krylov), so BiCGSTAB stays valid.  In the
// multi-GPU case level 1 and 2 are replicated on every rank: the restricted residual (8 * n_agg doubles) is
// summed over the ranks through NVLink peer memory inside the kernel that starts the V-cycle.
#include "tefem_common.cuh"
#include "tefem_comm.h"
#include "tefem_precond.h"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace tefem {

// rc[8I + 0..3] = sum_i r_i ; rc[8I + 4..6] = sum_i d_i x r_u,i ; rc[8I + 7] = 0 over the nodes i of aggregate I
// (d_i = offset from the aggregate's centre, single precision); one warp per aggregate, fixed shuffle tree.
__global__ void __launch_bounds__(256) restrict_kernel(int32_t n_agg, const int32_t* __restrict__ agg_ptr,
                                                       const int32_t* __restrict__ agg_nodes,
                                                       const float4* __restrict__ node_off,
                                                       const double* __restrict__ r, double* __restrict__ rc) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_agg) return;
    const int32_t s = agg_ptr[warp], t = agg_ptr[warp + 1];
    double a[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int32_t k = s + lane; k < t; k += 32) {
        const int32_t i = agg_nodes[k];
        const double* v = r + 4 * (int64_t)i;
        const float4 d = node_off[i];
        const double v0 = v[0], v1 = v[1], v2 = v[2];
        a[0] += v0; a[1] += v1; a[2] += v2; a[3] += v[3];
        a[4] += (double)d.y * v2 - (double)d.z * v1;
        a[5] += (double)d.z * v0 - (double)d.x * v2;
        a[6] += (double)d.x * v1 - (double)d.y * v0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] += __shfl_down_sync(0xffffffffu, a[k], off);
    if (lane == 0) {
        double* o = rc + 8 * (int64_t)warp;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = a[k];
        o[7] = 0.0;
    }
}

// z_i = omega * r_i + (t_I + w_I x d_i, theta_I) with (t, theta, w) = xc[8I ..]; one thread per node, 256-bit accesses
__global__ void __launch_bounds__(256) prolong_combine_kernel(int64_t n_nodes, const int32_t* __restrict__ node_agg,
                                                              const float4* __restrict__ node_off, double omega,
                                                              const double* __restrict__ r, const double* __restrict__ xc,
                                                              double* __restrict__ z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double r0, r1, r2, r3, c0, c1, c2, c3, w0, w1, w2, w3;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r0), "=d"(r1), "=d"(r2), "=d"(r3) : "l"(r + 4 * i));
    const double* c = xc + 8 * (int64_t)node_agg[i];
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(c));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(w0), "=d"(w1), "=d"(w2), "=d"(w3) : "l"(c + 4));
    const float4 d = node_off[i];
    r0 = omega * r0 + c0 + (w1 * (double)d.z - w2 * (double)d.y);
    r1 = omega * r1 + c1 + (w2 * (double)d.x - w0 * (double)d.z);
    r2 = omega * r2 + c2 + (w0 * (double)d.y - w1 * (double)d.x);
    r3 = omega * r3 + c3;
    (void)w3;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(z + 4 * i), "d"(r0), "d"(r1), "d"(r2), "d"(r3) : "memory");
}

// bc = D1^-1 rc (8x8 block per aggregate, row-major) ; xc = wc * bc.   One thread per scalar row.
__global__ void __launch_bounds__(256) coarse_start_kernel(int64_t n, const double* __restrict__ dinv,
                                                           const double* __restrict__ rc, double wc,
                                                           double* __restrict__ bc, double* __restrict__ xc) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* d = dinv + 8 * k;                    // row (k & 7) of block (k >> 3)
    const double* v = rc + (k & ~(int64_t)7);
    double b = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) b += d[c] * v[c];
    bc[k] = b;
    xc[k] = wc * b;
}

// Multi-GPU form of coarse_start_kernel, fused with the all-reduce of the restricted residual over NVLink peer
// memory (tefem_peer.cuh): every rank's restriction kernel wrote its partial P1^T r into its own peer buffer; this
// kernel raises the flags, waits for all ranks, adds the partial vectors in rank order (bitwise identical on every
// rank: level 1 and 2 stay consistent replicas) and applies D1^-1 and the first damped-Jacobi step.
// agg_ranks[i] (optional): bit r set iff rank r owns fine nodes of aggregate i - only those ranks' partial sums are
// read (the others are exact zeros), i.e. 1 - 2 remote reads per aggregate for a slab partition instead of world:
// 6 x fewer bytes over NVLink at 8 GPUs.  One thread per aggregate (8 doubles, four 128-bit volatile loads per rank).
__global__ void __launch_bounds__(128) peer_coarse_start_kernel(const PeerCtx pc, int32_t n_agg,
                                                                const uint8_t* __restrict__ agg_ranks,
                                                                const double* __restrict__ dinv, double wc,
                                                                double* __restrict__ bc, double* __restrict__ xc) {
    if (blockIdx.x == 0) {
        __threadfence_system();
        peer_signal(pc, threadIdx.x);
    }
    peer_wait(pc, threadIdx.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agg) return;
    const unsigned mask = agg_ranks ? (unsigned)agg_ranks[i] : 0xffu;
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < pc.world; ++r) {
        if (!((mask >> r) & 1u)) continue;
        const double* src = peer_big(pc, r) + 8 * i;
        double v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) ld_volatile_v2(src + 2 * q, v[2 * q], v[2 * q + 1]);
#pragma unroll
        for (int q = 0; q < 8; ++q) s[q] += v[q];
    }
    const double* D = dinv + 64 * i;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double b = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) b += D[8 * k + c] * s[c];
        bc[8 * i + k] = b;
        xc[8 * i + k] = wc * b;
    }
}

// out = c0 * xin + c1 * (cb * b - A x) for a 4x4-block CSR matrix A in SINGLE precision (b == nullptr: 0; xin is read
// only when c0 != 0): the level-1 operator, and the transfer operators P2^T / P2 between levels 1 and 2.  It only
// shapes the preconditioner (a fixed linear operator either way), the Krylov vectors and the fine operator stay
// fp64; half the bytes keep the whole level in L2.
// COARSE_LANES (16) lanes share one block row: lane l multiplies the blocks s + l, s + l + 16, ...; the column index of
// the NEXT trip is loaded one trip ahead, so that a trip costs ONE memory latency (block and x_j loads issue together)
// instead of two, at full occupancy (few registers); a fixed shuffle tree adds the 4 row sums and lanes 0 .. 3 store.
// These kernels are latency bound, not bandwidth bound: with one thread per scalar row (round 1) a sweep was a 27-trip
// chain of dependent loads, with 8 lanes per row (first version of round 2) still 7 trips of two latencies each: ~10 us
// per launch whether the operator came from L2 or from HBM (profiles/README.md).  The sum order is fixed (lane-strided
// partial sums, then the tree), so every rank computes bitwise the same coarse vectors.
constexpr int COARSE_LANES = 16;

// One 4x4 block (row-major) of a coarse operator as 16 floats: stored in single precision (transfer operators), or in
// HALF precision (the level-1 operator A1^ = D1^-1 A1: unit diagonal blocks, off-diagonal entries of magnitude <~ 1, so
// the 11-bit mantissa perturbs the preconditioner by ~5e-4 relative - measured: no change of the BiCGSTAB iteration
// counts at 150^3, profiles/README.md - and halves the bytes of the six level-1 sweeps of an application: the 190^3
// level then fits the L2 next to the vectors).
__device__ __forceinline__ void load_block16(const float* __restrict__ sys, int64_t p, float4& r0, float4& r1, float4& r2, float4& r3) {
    const float4* blk = reinterpret_cast<const float4*>(sys + 16 * p);
    r0 = __ldg(blk); r1 = __ldg(blk + 1); r2 = __ldg(blk + 2); r3 = __ldg(blk + 3);
}
__device__ __forceinline__ void load_block16(const __half* __restrict__ sys, int64_t p, float4& r0, float4& r1, float4& r2, float4& r3) {
    const uint4* blk = reinterpret_cast<const uint4*>(sys + 16 * p);
    const uint4 lo = __ldg(blk), hi = __ldg(blk + 1);
    const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&lo.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&lo.y));
    const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&lo.z)), a3 = __half22float2(*reinterpret_cast<const __half2*>(&lo.w));
    const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
    const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&hi.z)), b3 = __half22float2(*reinterpret_cast<const __half2*>(&hi.w));
    r0 = make_float4(a0.x, a0.y, a1.x, a1.y); r1 = make_float4(a2.x, a2.y, a3.x, a3.y);
    r2 = make_float4(b0.x, b0.y, b1.x, b1.y); r3 = make_float4(b2.x, b2.y, b3.x, b3.y);
}

template <typename VT>
__global__ void __launch_bounds__(256) coarse_spmv_kernel(const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ colidx, int32_t n_rows,
                                                          const VT* __restrict__ sys, const double* __restrict__ x,
                                                          const double* __restrict__ b, const double* xin,
                                                          double c0, double c1, double cb, double* out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t i = (int32_t)(gid / COARSE_LANES);
    const int l = (int)(gid % COARSE_LANES);
    const bool live = i < n_rows;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (live) {
        const int32_t s = rowptr[i], t = rowptr[i + 1];
        int32_t p = s + l;
        int32_t j = p < t ? colidx[p] : 0;
        while (p < t) {
            const int32_t pn = p + COARSE_LANES;
            const int32_t jn = pn < t ? colidx[pn] : 0;            // next trip's column: in flight during this trip
            float4 r0, r1, r2, r3;
            load_block16(sys, (int64_t)p, r0, r1, r2, r3);
            const double2 u0 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j);
            const double2 u1 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j + 2);
            a0 += (double)r0.x * u0.x + (double)r0.y * u0.y + (double)r0.z * u1.x + (double)r0.w * u1.y;
            a1 += (double)r1.x * u0.x + (double)r1.y * u0.y + (double)r1.z * u1.x + (double)r1.w * u1.y;
            a2 += (double)r2.x * u0.x + (double)r2.y * u0.y + (double)r2.z * u1.x + (double)r2.w * u1.y;
            a3 += (double)r3.x * u0.x + (double)r3.y * u0.y + (double)r3.z * u1.x + (double)r3.w * u1.y;
            p = pn; j = jn;
        }
    }
    // all 32 lanes of the warp take part in the shuffles (rows never straddle a warp: 16 divides 32)
#pragma unroll
    for (int off = COARSE_LANES / 2; off > 0; off >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, off, COARSE_LANES);
        a1 += __shfl_down_sync(0xffffffffu, a1, off, COARSE_LANES);
        a2 += __shfl_down_sync(0xffffffffu, a2, off, COARSE_LANES);
        a3 += __shfl_down_sync(0xffffffffu, a3, off, COARSE_LANES);
    }
    // lane r < 4 keeps row r: a_r of lane 0 is fetched by lane r
    const double s1 = __shfl_sync(0xffffffffu, a1, 0, COARSE_LANES);
    const double s2 = __shfl_sync(0xffffffffu, a2, 0, COARSE_LANES);
    const double s3 = __shfl_sync(0xffffffffu, a3, 0, COARSE_LANES);
    const double s0 = __shfl_sync(0xffffffffu, a0, 0, COARSE_LANES);
    if (live && l < 4) {
        const double acc = l == 0 ? s0 : (l == 1 ? s1 : (l == 2 ? s2 : s3));
        const int64_t k = 4 * (int64_t)i + l;
        out[k] = (c0 != 0.0 ? c0 * xin[k] : 0.0) + c1 * ((b ? cb * b[k] : 0.0) - acc);
    }
}

// y = M x for a dense row-major n x n single-precision matrix; one CTA of 128 threads per row, fixed reduction tree.
__global__ void __launch_bounds__(128) dense_matvec_kernel(int32_t n, const float* __restrict__ M,
                                                           const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double red[32];
    const float* m = M + (int64_t)blockIdx.x * n;
    double a[1] = {0.0};
    for (int k = threadIdx.x; k < n; k += 128) a[0] += (double)m[k] * x[k];
    block_reduce_sum<1>(a, red);
    if (threadIdx.x == 0) y[blockIdx.x] = a[0];
}

// (A fused form - restriction, all-reduce, V-cycle and prolongation in ONE persistent cooperative kernel with grid.sync()
// between the phases - was built and measured in round 2: 0.32 ms per application against 0.19 ms for the separate
// launches at 150^3 on one GPU, 0.26 against 0.15 ms on two: a grid-wide barrier costs as much as a kernel boundary here
// and the fine-level passes lose their occupancy.  Removed; numbers in profiles/README.md.)

}  // namespace tefem

using namespace tefem;

struct tefem_precond {
    int32_t n_rows = 0, n1 = 0, n2 = 0;          // fine node rows, level-1 aggregates, level-2 aggregates
    const int32_t *agg_ptr = nullptr, *agg_nodes = nullptr, *node_agg = nullptr;
    const float4* node_off = nullptr;
    const int32_t *rowptr1 = nullptr, *colidx1 = nullptr;
    const float* sys1 = nullptr;
    const double* dinv1 = nullptr;
    const int32_t *r2_rowptr = nullptr, *r2_colidx = nullptr, *p2_rowptr = nullptr, *p2_colidx = nullptr;
    const float *r2_vals = nullptr, *p2_vals = nullptr;
    const float* inv2 = nullptr;
    double omega = 1.0, omega_c = 0.8;
    int n_pre = 1, n_post = 1;                   // damped-Jacobi sweeps on level 1 before / after the level-2 correction
    double *rc = nullptr, *bc = nullptr, *xc = nullptr, *tc = nullptr, *r2 = nullptr, *x2 = nullptr;
    tefem_comm* comm = nullptr;
    const uint8_t* agg_ranks = nullptr;          // multi-GPU: which ranks contribute to an aggregate's restricted residual
    const __half* sys1h = nullptr;               // level-1 operator in half precision (tefem_precond_set_level1_half)
    bool overlap = true;                         // last pre-smoothing sweep and level-2 correction from one residual
    // TEFEM_PROFILE_PRECOND=1 (diagnostics, stderr at destroy): device time of the three parts of an application -
    // restriction + all-reduce over the ranks + first sweep (includes waiting for the slowest rank), the level-1 / level-2
    // cycle (replicated work), the prolongation.  Synchronises the stream after every application: not for timed runs.
    bool profile = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double part_ms[3] = {0.0, 0.0, 0.0};
    int64_t n_apply = 0;
};

static int64_t pc_align(int64_t n) { return (n + 15) / 16 * 16; }

extern "C" int64_t tefem_precond_work_doubles(int32_t n_agg, int32_t n_agg2) {
    return 4 * pc_align(8 * (int64_t)n_agg) + 2 * pc_align(8 * (int64_t)n_agg2);
}

extern "C" int tefem_precond_create(tefem_precond** out, int32_t n_rows, int32_t n_agg, const int32_t* agg_ptr,
                                    const int32_t* agg_nodes, const int32_t* node_agg, const float* node_off, double omega,
                                    const int32_t* rowptr1, const int32_t* colidx1, const float* sys1,
                                    const double* dinv1, double omega_c, int n_pre, int n_post, int32_t n_agg2,
                                    const int32_t* r2_rowptr,
                                    const int32_t* r2_colidx, const float* r2_vals, const int32_t* p2_rowptr,
                                    const int32_t* p2_colidx, const float* p2_vals, const float* inv2, double* work,
                                    void* comm) {
    TEFEM_REQUIRE(out && agg_ptr && agg_nodes && node_agg && node_off && rowptr1 && colidx1 && sys1 && dinv1 && work,
                  "null pointer");
    TEFEM_REQUIRE(n_agg > 0 && n_rows >= 0 && n_agg2 > 0, "sizes");
    TEFEM_REQUIRE(n_pre >= 1 && n_pre <= 8 && n_post >= 0 && n_post <= 8, "1 .. 8 pre- and 0 .. 8 post-smoothing sweeps");
    TEFEM_REQUIRE(r2_rowptr && r2_colidx && r2_vals && p2_rowptr && p2_colidx && p2_vals && inv2, "null level-2 arrays");
    TEFEM_REQUIRE((((uintptr_t)work) & 127) == 0 && (((uintptr_t)sys1) & 63) == 0 && (((uintptr_t)node_off) & 15) == 0 &&
                      (((uintptr_t)r2_vals) & 63) == 0 && (((uintptr_t)p2_vals) & 63) == 0,
                  "work must be 128-byte, the block arrays 64-byte, node_off 16-byte aligned");
    tefem_precond* p = new tefem_precond();
    p->n_rows = n_rows; p->n1 = n_agg; p->n2 = n_agg2;
    p->agg_ptr = agg_ptr; p->agg_nodes = agg_nodes; p->node_agg = node_agg;
    p->node_off = reinterpret_cast<const float4*>(node_off);
    p->rowptr1 = rowptr1; p->colidx1 = colidx1; p->sys1 = sys1; p->dinv1 = dinv1;
    p->r2_rowptr = r2_rowptr; p->r2_colidx = r2_colidx; p->r2_vals = r2_vals;
    p->p2_rowptr = p2_rowptr; p->p2_colidx = p2_colidx; p->p2_vals = p2_vals; p->inv2 = inv2;
    p->omega = omega; p->omega_c = omega_c; p->n_pre = n_pre; p->n_post = n_post;
    const int64_t a1 = pc_align(8 * (int64_t)n_agg), a2 = pc_align(8 * (int64_t)n_agg2);
    double* w = work;
    p->rc = w; w += a1; p->bc = w; w += a1; p->xc = w; w += a1; p->tc = w; w += a1;
    p->r2 = w; w += a2; p->x2 = w;
    p->comm = reinterpret_cast<tefem_comm*>(comm);
    const char* pe = getenv("TEFEM_PROFILE_PRECOND");
    if (pe && pe[0] == '1') {
        p->profile = true;
        for (int k = 0; k < 4; ++k) {
            if (cudaEventCreate(&p->ev[k]) != cudaSuccess) { p->profile = false; break; }
        }
    }
    *out = p;
    return TEFEM_OK;
}

extern "C" int tefem_precond_set_level1_half(tefem_precond* p, const void* sys1_half) {
    TEFEM_REQUIRE(p, "null preconditioner");
    TEFEM_REQUIRE((((uintptr_t)sys1_half) & 31) == 0, "the half-precision blocks must be 32-byte aligned");
    p->sys1h = reinterpret_cast<const __half*>(sys1_half);
    return TEFEM_OK;
}

extern "C" int tefem_precond_set_cycle(tefem_precond* p, int overlapped) {
    TEFEM_REQUIRE(p, "null preconditioner");
    p->overlap = overlapped != 0;
    return TEFEM_OK;
}

extern "C" int tefem_precond_set_agg_ranks(tefem_precond* p, const uint8_t* agg_ranks) {
    TEFEM_REQUIRE(p, "null preconditioner");
    p->agg_ranks = agg_ranks;
    return TEFEM_OK;
}

extern "C" int tefem_precond_destroy(tefem_precond* p) {
    if (p && p->profile) {
        if (p->n_apply > 0)
            fprintf(stderr, "[tefem precond profile] %lld applications: restrict + all-reduce + first sweep %.4f ms, level-1/2 cycle %.4f ms, "
                            "prolongation %.4f ms per application\n", (long long)p->n_apply, p->part_ms[0] / p->n_apply,
                    p->part_ms[1] / p->n_apply, p->part_ms[2] / p->n_apply);
        for (int k = 0; k < 4; ++k)
            if (p->ev[k]) cudaEventDestroy(p->ev[k]);
    }
    delete p;
    return TEFEM_OK;
}

int tefem_precond_apply_stream(tefem_precond* p, const double* r, double* z, cudaStream_t st) {
    TEFEM_REQUIRE(p && r && z, "null pointer");
    const int64_t n1 = 8 * (int64_t)p->n1;                    // scalar size of level 1
    const int32_t np1 = 2 * p->n1, np2 = 2 * p->n2;           // pseudo nodes (4-DOF block rows) of levels 1 and 2
    const unsigned g1 = (unsigned)ceil_div(n1, 256);                                           // one thread per scalar of level 1
    const unsigned gs1 = (unsigned)ceil_div((int64_t)COARSE_LANES * np1, 256);                 // coarse_spmv: 16 lanes per block row
    const unsigned gs2 = (unsigned)ceil_div((int64_t)COARSE_LANES * np2, 256);
    const unsigned gr = (unsigned)ceil_div(32 * (int64_t)p->n1, 256);
    if (p->profile) cudaEventRecord(p->ev[0], st);
    // level 0 -> 1: rc = P1^T r, summed over the ranks (levels 1 and 2 are replicated): the restriction writes
    // into this rank's peer buffer and the next kernel does all-reduce + D1^-1 + first smoothing step in one
    if (p->comm) {
        TEFEM_REQUIRE(tefem_comm_has_peer(p->comm) && tefem_comm_peer_cap(p->comm) >= n1,
                      "the peer-memory region must hold 8 * n_agg doubles (tefem_comm_shm_alloc)");
        const PeerCtx pc = tefem_comm_peer_ctx(p->comm, PEER_CH_COARSE);
        double* mine = reinterpret_cast<double*>(pc.base[pc.rank] + PEER_BIG_OFFSET) + (int64_t)(pc.epoch & 1u) * pc.cap;
        restrict_kernel<<<gr, 256, 0, st>>>(p->n1, p->agg_ptr, p->agg_nodes, p->node_off, r, mine);
        TEFEM_LAUNCH_CHECK();
        peer_coarse_start_kernel<<<(unsigned)ceil_div((int64_t)p->n1, 128), 128, 0, st>>>(pc, p->n1, p->agg_ranks, p->dinv1, p->omega_c,
                                                                                          p->bc, p->xc);
        TEFEM_LAUNCH_CHECK();
    } else {
        restrict_kernel<<<gr, 256, 0, st>>>(p->n1, p->agg_ptr, p->agg_nodes, p->node_off, r, p->rc);
        TEFEM_LAUNCH_CHECK();
        coarse_start_kernel<<<g1, 256, 0, st>>>(n1, p->dinv1, p->rc, p->omega_c, p->bc, p->xc);
        TEFEM_LAUNCH_CHECK();
    }
    if (p->profile) cudaEventRecord(p->ev[1], st);
    // the level-1 operator: half precision when attached, else single
#define LEVEL1_SPMV(X, B, XIN, C0, C1, OUT)                                                                                         \
    do {                                                                                                                            \
        if (p->sys1h) coarse_spmv_kernel<__half><<<gs1, 256, 0, st>>>(p->rowptr1, p->colidx1, np1, p->sys1h, X, B, XIN, C0, C1, 1.0, OUT); \
        else coarse_spmv_kernel<float><<<gs1, 256, 0, st>>>(p->rowptr1, p->colidx1, np1, p->sys1, X, B, XIN, C0, C1, 1.0, OUT);           \
    } while (0)
    // level 1 V-cycle on A1^ = D1^-1 A1 (bc = D1^-1 rc, x = wc bc so far = the first damped-Jacobi sweep from 0):
    // further pre-smoothing sweeps, correction from level 2, post-smoothing.  A sweep reads all of x, so it writes
    // into the other buffer: x' = x + wc (bc - A1^ x).
    //   classic cycle:  n_pre - 1 sweeps, residual t = bc - A1^ x, x += P2 A2^-1 P2^T t
    //   overlapped (default, n_pre >= 2): n_pre - 2 sweeps, residual t = bc - A1^ x, then the LAST pre-smoothing sweep and
    //     the level-2 correction are both formed from that one residual: x += wc t + P2 A2^-1 P2^T t - one level-1 product
    //     less per application, still a fixed linear operator; BiCGSTAB iterations on the scipy prototype of config 4's
    //     operator: 92 -> 92 (24^3 box), 111 -> 101 (48^3) (profiles/README.md)
    double *cur = p->xc, *alt = p->rc;
    const bool overlap = p->overlap && p->n_pre >= 2;
    for (int k = 1; k < p->n_pre - (overlap ? 1 : 0); ++k) {
        LEVEL1_SPMV(cur, p->bc, cur, 1.0, p->omega_c, alt);
        TEFEM_LAUNCH_CHECK();
        double* t = cur; cur = alt; alt = t;
    }
    LEVEL1_SPMV(cur, p->bc, nullptr, 0.0, 1.0, p->tc);
    TEFEM_LAUNCH_CHECK();                                                                       // tc = bc - A1^ x
    coarse_spmv_kernel<float><<<gs2, 256, 0, st>>>(p->r2_rowptr, p->r2_colidx, np2, p->r2_vals, p->tc, nullptr, nullptr, 0.0, -1.0, 1.0, p->r2);
    TEFEM_LAUNCH_CHECK();                                                                       // r2 = P2^T tc
    dense_matvec_kernel<<<(unsigned)(4 * np2), 128, 0, st>>>(4 * np2, p->inv2, p->r2, p->x2);
    TEFEM_LAUNCH_CHECK();                                                                       // x2 = A2^-1 r2
    if (overlap)   // x += wc tc + P2 x2  ( = x - ( -wc tc - P2 x2 ) )
        coarse_spmv_kernel<float><<<gs1, 256, 0, st>>>(p->p2_rowptr, p->p2_colidx, np1, p->p2_vals, p->x2, p->tc, cur, 1.0, -1.0, -p->omega_c, cur);
    else           // x += P2 x2
        coarse_spmv_kernel<float><<<gs1, 256, 0, st>>>(p->p2_rowptr, p->p2_colidx, np1, p->p2_vals, p->x2, nullptr, cur, 1.0, -1.0, 1.0, cur);
    TEFEM_LAUNCH_CHECK();
    for (int k = 0; k < p->n_post; ++k) {
        LEVEL1_SPMV(cur, p->bc, cur, 1.0, p->omega_c, alt);
        TEFEM_LAUNCH_CHECK();
        double* t = cur; cur = alt; alt = t;
    }
#undef LEVEL1_SPMV
    if (p->profile) cudaEventRecord(p->ev[2], st);
    // level 1 -> 0
    if (p->n_rows > 0) {
        prolong_combine_kernel<<<(unsigned)ceil_div((int64_t)p->n_rows, 256), 256, 0, st>>>(p->n_rows, p->node_agg, p->node_off,
                                                                                            p->omega, r, cur, z);
        TEFEM_LAUNCH_CHECK();
    }
    if (p->profile) {
        cudaEventRecord(p->ev[3], st);
        if (cudaEventSynchronize(p->ev[3]) == cudaSuccess) {
            for (int k = 0; k < 3; ++k) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]) == cudaSuccess) p->part_ms[k] += ms;
            }
            p->n_apply += 1;
        }
    }
    return TEFEM_OK;
}

extern "C" int tefem_precond_apply(tefem_precond* p, const double* r, double* z, void* stream) {
    return tefem_precond_apply_stream(p, r, z, as_stream(stream));
}
