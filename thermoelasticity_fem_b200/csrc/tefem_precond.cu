// Three-level aggregation preconditioner for the block-Jacobi-scaled system (right preconditioner of BiCGSTAB).
//
// The reference solves with a sparse direct solver (spsolve, solvers.py:26,172); the Krylov replacement with the
// per-node block-Jacobi scaling alone needs ~10 iterations per mesh cell along the longest axis (SURVEY.md H5:
// ~2000 at config B).  This adds a coarse-space correction in the additive (BPX-like) form
//
//       z = omega * r  +  P1 * M1( P1^T r )
//
// on the ALREADY block-Jacobi-scaled fine operator A^ = D^-1 A, where
//   * P1 is the piecewise-constant prolongation of a geometric aggregation of the nodes (one coarse node per
//     aggregate, the 4 DOFs u_x, u_y, u_z, theta kept separate), A1 = P1^T A^ P1 (Galerkin, built on the host
//     side with torch, 4x4-block CSR like the fine level) and A1^ = D1^-1 A1 its block-Jacobi-scaled form, stored
//     in single precision (preconditioner data only; every vector and the fine operator stay fp64);
//   * M1 is one V-cycle on level 1: damped Jacobi on A1^ (= a scaling by omega_c, the diagonal blocks are the
//     identity), a level-2 correction P2 * inv(P2^T A1^ P2) * P2^T with a DENSE inverse (a few hundred nodes), and
//     a post-smoothing step;
//   * the fine level needs NO extra SpMV: per application one restriction pass and one prolongation pass over
//     the fine vector, the level-1 work is ~1/500 of a fine SpMV.
// Every step is a fixed linear operator (fixed sweeps, no inner Krylov), so BiCGSTAB stays valid.  In the
// multi-GPU case level 1 and 2 are replicated on every rank: the restricted residual (4 * n_coarse doubles) is
// summed with one all-reduce per application.
// Measured in the prototype (scipy) on the config-2 operator, Kuhn beam 120 x 30 x 18 cells (285 k DOF):
// 1260 BiCGSTAB iterations with block-Jacobi alone, 306 with aggregation factors (8, 4), 190 with (4, 4).
#include "tefem_common.cuh"
#include "tefem_comm.h"
#include "tefem_precond.h"

namespace tefem {

// rc[4I + c] = sum over the nodes i of aggregate I of r[4i + c]; one warp per aggregate, fixed shuffle tree.
__global__ void __launch_bounds__(256) restrict_kernel(int32_t n_agg, const int32_t* __restrict__ agg_ptr,
                                                       const int32_t* __restrict__ agg_nodes,
                                                       const double* __restrict__ r, double* __restrict__ rc) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_agg) return;
    const int32_t s = agg_ptr[warp], t = agg_ptr[warp + 1];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int32_t k = s + lane; k < t; k += 32) {
        const double* v = r + 4 * (int64_t)agg_nodes[k];
        a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, off);
        a1 += __shfl_down_sync(0xffffffffu, a1, off);
        a2 += __shfl_down_sync(0xffffffffu, a2, off);
        a3 += __shfl_down_sync(0xffffffffu, a3, off);
    }
    if (lane == 0) {
        double* o = rc + 4 * (int64_t)warp;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
    }
}

// z[4i + c] = omega * r[4i + c] + xc[4 * node_agg[i] + c]; one thread per node, 256-bit accesses
__global__ void __launch_bounds__(256) prolong_combine_kernel(int64_t n_nodes, const int32_t* __restrict__ node_agg, double omega,
                                                              const double* __restrict__ r, const double* __restrict__ xc,
                                                              double* __restrict__ z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double r0, r1, r2, r3, c0, c1, c2, c3;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r0), "=d"(r1), "=d"(r2), "=d"(r3) : "l"(r + 4 * i));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3)
                 : "l"(xc + 4 * (int64_t)node_agg[i]));
    r0 = omega * r0 + c0; r1 = omega * r1 + c1; r2 = omega * r2 + c2; r3 = omega * r3 + c3;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(z + 4 * i), "d"(r0), "d"(r1), "d"(r2), "d"(r3) : "memory");
}

// x[4i + c] += x2[4 * node_agg[i] + c]
__global__ void __launch_bounds__(256) prolong_add_kernel(int64_t n, const int32_t* __restrict__ node_agg,
                                                          const double* __restrict__ x2, double* __restrict__ x) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    x[k] += x2[4 * (int64_t)node_agg[k >> 2] + (k & 3)];
}

// bc = D1^-1 rc (4x4 block per coarse node, row-major) ; xc = wc * bc.   One thread per scalar row.
__global__ void __launch_bounds__(256) coarse_start_kernel(int64_t n, const double* __restrict__ dinv,
                                                           const double* __restrict__ rc, double wc,
                                                           double* __restrict__ bc, double* __restrict__ xc) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* d = dinv + 4 * k;                    // row (k & 3) of block (k >> 2)
    const double* v = rc + (k & ~(int64_t)3);
    const double b = d[0] * v[0] + d[1] * v[1] + d[2] * v[2] + d[3] * v[3];
    bc[k] = b;
    xc[k] = wc * b;
}

// Multi-GPU form of coarse_start_kernel, fused with the all-reduce of the restricted residual over NVLink peer
// memory (tefem_peer.cuh): every rank's restriction kernel wrote its partial P1^T r into its own peer buffer; this
// kernel raises the flags, waits for all ranks, adds the partial vectors of all ranks in rank order (bitwise
// identical on every rank: level 1 and 2 stay consistent replicas) and applies D1^-1 and the first damped-Jacobi
// step.  One thread per coarse node (4 doubles, two 128-bit volatile loads per rank).
__global__ void __launch_bounds__(256) peer_coarse_start_kernel(const PeerCtx pc, int32_t n_nodes,
                                                                const double* __restrict__ dinv, double wc,
                                                                double* __restrict__ bc, double* __restrict__ xc) {
    if (blockIdx.x == 0) {
        __threadfence_system();
        peer_signal(pc, threadIdx.x);
    }
    peer_wait(pc, threadIdx.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int r = 0; r < pc.world; ++r) {
        const double* src = peer_big(pc, r) + 4 * i;
        double a, b, c, d;
        ld_volatile_v2(src, a, b);
        ld_volatile_v2(src + 2, c, d);
        s0 += a; s1 += b; s2 += c; s3 += d;
    }
    const double* D = dinv + 16 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double b = D[4 * k] * s0 + D[4 * k + 1] * s1 + D[4 * k + 2] * s2 + D[4 * k + 3] * s3;
        bc[4 * i + k] = b;
        xc[4 * i + k] = wc * b;
    }
}

// out = c0 * xin + c1 * (b - A1^ x) on level 1.  The level-1 operator is stored in SINGLE precision: it only
// shapes the preconditioner (a fixed linear operator either way), the Krylov vectors and the fine operator stay
// fp64; half the bytes keep the whole level in L2.  4 threads per coarse node, one 128-bit load per block row.
__global__ void __launch_bounds__(256) coarse_spmv_kernel(const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ colidx, int32_t n_rows,
                                                          const float* __restrict__ sys, const double* __restrict__ x,
                                                          const double* __restrict__ b, const double* __restrict__ xin,
                                                          double c0, double c1, double* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= 4 * (int64_t)n_rows) return;
    const int32_t i = (int32_t)(tid >> 2);
    const int r = (int)(tid & 3);
    const int32_t s = rowptr[i], t = rowptr[i + 1];
    double acc0 = 0.0, acc1 = 0.0;
    int32_t p = s;
    for (; p + 2 <= t; p += 2) {
        const int32_t j0 = colidx[p], j1 = colidx[p + 1];
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(sys + 16 * (int64_t)p + 4 * r));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(sys + 16 * (int64_t)(p + 1) + 4 * r));
        const double2 u0 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j0);
        const double2 u1 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j0 + 2);
        const double2 w0 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j1);
        const double2 w1 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j1 + 2);
        acc0 += (double)a0.x * u0.x + (double)a0.y * u0.y + (double)a0.z * u1.x + (double)a0.w * u1.y;
        acc1 += (double)a1.x * w0.x + (double)a1.y * w0.y + (double)a1.z * w1.x + (double)a1.w * w1.y;
    }
    if (p < t) {
        const int32_t j0 = colidx[p];
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(sys + 16 * (int64_t)p + 4 * r));
        const double2 u0 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j0);
        const double2 u1 = *reinterpret_cast<const double2*>(x + 4 * (int64_t)j0 + 2);
        acc0 += (double)a0.x * u0.x + (double)a0.y * u0.y + (double)a0.z * u1.x + (double)a0.w * u1.y;
    }
    out[tid] = (c0 != 0.0 ? c0 * xin[tid] : 0.0) + c1 * (b[tid] - (acc0 + acc1));
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

}  // namespace tefem

using namespace tefem;

struct tefem_precond {
    int32_t n_rows = 0, n1 = 0, n2 = 0;
    const int32_t *agg_ptr = nullptr, *agg_nodes = nullptr, *node_agg = nullptr;
    const int32_t *rowptr1 = nullptr, *colidx1 = nullptr;
    const float* sys1 = nullptr;
    const double* dinv1 = nullptr;
    const int32_t *agg2_ptr = nullptr, *agg2_nodes = nullptr, *node_agg2 = nullptr;
    const float* inv2 = nullptr;
    double omega = 0.7, omega_c = 0.6;
    double *rc = nullptr, *bc = nullptr, *xc = nullptr, *tc = nullptr, *r2 = nullptr, *x2 = nullptr;
    tefem_comm* comm = nullptr;
};

static int64_t pc_align(int64_t n) { return (n + 15) / 16 * 16; }

extern "C" int64_t tefem_precond_work_doubles(int32_t n_coarse, int32_t n_coarse2) {
    return 4 * pc_align(4 * (int64_t)n_coarse) + 2 * pc_align(4 * (int64_t)n_coarse2);
}

extern "C" int tefem_precond_create(tefem_precond** out, int32_t n_rows, int32_t n_coarse, const int32_t* agg_ptr,
                                    const int32_t* agg_nodes, const int32_t* node_agg, double omega,
                                    const int32_t* rowptr1, const int32_t* colidx1, const float* sys1,
                                    const double* dinv1, double omega_c, int32_t n_coarse2, const int32_t* agg2_ptr,
                                    const int32_t* agg2_nodes, const int32_t* node_agg2, const float* inv2, double* work,
                                    void* comm) {
    TEFEM_REQUIRE(out && agg_ptr && agg_nodes && node_agg && rowptr1 && colidx1 && sys1 && dinv1 && work, "null pointer");
    TEFEM_REQUIRE(n_coarse > 0 && n_rows >= 0, "sizes");
    TEFEM_REQUIRE(n_coarse2 == 0 || (agg2_ptr && agg2_nodes && node_agg2 && inv2), "null level-2 arrays");
    TEFEM_REQUIRE((((uintptr_t)work) & 127) == 0 && (((uintptr_t)sys1) & 63) == 0, "work must be 128-byte, sys1 64-byte aligned");
    tefem_precond* p = new tefem_precond();
    p->n_rows = n_rows; p->n1 = n_coarse; p->n2 = n_coarse2;
    p->agg_ptr = agg_ptr; p->agg_nodes = agg_nodes; p->node_agg = node_agg;
    p->rowptr1 = rowptr1; p->colidx1 = colidx1; p->sys1 = sys1; p->dinv1 = dinv1;
    p->agg2_ptr = agg2_ptr; p->agg2_nodes = agg2_nodes; p->node_agg2 = node_agg2; p->inv2 = inv2;
    p->omega = omega; p->omega_c = omega_c;
    const int64_t a1 = pc_align(4 * (int64_t)n_coarse), a2 = pc_align(4 * (int64_t)n_coarse2);
    double* w = work;
    p->rc = w; w += a1; p->bc = w; w += a1; p->xc = w; w += a1; p->tc = w; w += a1;
    p->r2 = w; w += a2; p->x2 = w;
    p->comm = reinterpret_cast<tefem_comm*>(comm);
    *out = p;
    return TEFEM_OK;
}

extern "C" int tefem_precond_destroy(tefem_precond* p) {
    delete p;
    return TEFEM_OK;
}

int tefem_precond_apply_stream(tefem_precond* p, const double* r, double* z, cudaStream_t st) {
    TEFEM_REQUIRE(p && r && z, "null pointer");
    const int64_t nf = 4 * (int64_t)p->n_rows, n1 = 4 * (int64_t)p->n1;
    // level 0 -> 1: rc = P1^T r, summed over the ranks (levels 1 and 2 are replicated): the restriction writes
    // into this rank's peer buffer and the next kernel does all-reduce + D1^-1 + first smoothing step in one
    const unsigned g1 = (unsigned)ceil_div(n1, 256);
    if (p->comm) {
        TEFEM_REQUIRE(tefem_comm_has_peer(p->comm) && tefem_comm_peer_cap(p->comm) >= n1,
                      "the peer-memory region must hold 4 * n_coarse doubles (tefem_comm_shm_alloc)");
        const PeerCtx pc = tefem_comm_peer_ctx(p->comm, PEER_CH_COARSE);
        double* mine = reinterpret_cast<double*>(pc.base[pc.rank] + PEER_BIG_OFFSET) + (int64_t)(pc.epoch & 1u) * pc.cap;
        restrict_kernel<<<(unsigned)ceil_div(32 * (int64_t)p->n1, 256), 256, 0, st>>>(p->n1, p->agg_ptr, p->agg_nodes, r, mine);
        TEFEM_LAUNCH_CHECK();
        peer_coarse_start_kernel<<<(unsigned)ceil_div((int64_t)p->n1, 256), 256, 0, st>>>(pc, p->n1, p->dinv1, p->omega_c, p->bc, p->xc);
        TEFEM_LAUNCH_CHECK();
    } else {
        restrict_kernel<<<(unsigned)ceil_div(32 * (int64_t)p->n1, 256), 256, 0, st>>>(p->n1, p->agg_ptr, p->agg_nodes, r, p->rc);
        TEFEM_LAUNCH_CHECK();
        coarse_start_kernel<<<g1, 256, 0, st>>>(n1, p->dinv1, p->rc, p->omega_c, p->bc, p->xc);
        TEFEM_LAUNCH_CHECK();
    }
    // level 1 V-cycle on A1^ = D1^-1 A1:  bc = D1^-1 rc ; xc = wc bc ; correction from level 2 ; xc += wc (bc - A1^ xc)
    const double* xcur = p->xc;
    if (p->n2 > 0) {
        coarse_spmv_kernel<<<g1, 256, 0, st>>>(p->rowptr1, p->colidx1, p->n1, p->sys1, p->xc, p->bc, nullptr, 0.0, 1.0, p->tc);
        TEFEM_LAUNCH_CHECK();                                                                   // tc = bc - A1^ xc
        restrict_kernel<<<(unsigned)ceil_div(32 * (int64_t)p->n2, 256), 256, 0, st>>>(p->n2, p->agg2_ptr, p->agg2_nodes, p->tc, p->r2);
        TEFEM_LAUNCH_CHECK();
        dense_matvec_kernel<<<(unsigned)(4 * p->n2), 128, 0, st>>>(4 * p->n2, p->inv2, p->r2, p->x2);
        TEFEM_LAUNCH_CHECK();
        prolong_add_kernel<<<g1, 256, 0, st>>>(n1, p->node_agg2, p->x2, p->xc);
        TEFEM_LAUNCH_CHECK();
    }
    // post-smoothing into a second buffer (the SpMV reads all of xc): rc <- xc + wc (bc - A1^ xc)
    coarse_spmv_kernel<<<g1, 256, 0, st>>>(p->rowptr1, p->colidx1, p->n1, p->sys1, p->xc, p->bc, p->xc, 1.0, p->omega_c, p->rc);
    TEFEM_LAUNCH_CHECK();
    xcur = p->rc;
    // level 1 -> 0
    if (nf > 0) {
        prolong_combine_kernel<<<(unsigned)ceil_div((int64_t)p->n_rows, 256), 256, 0, st>>>(p->n_rows, p->node_agg, p->omega, r, xcur, z);
        TEFEM_LAUNCH_CHECK();
    }
    return TEFEM_OK;
}

extern "C" int tefem_precond_apply(tefem_precond* p, const double* r, double* z, void* stream) {
    return tefem_precond_apply_stream(p, r, z, as_stream(stream));
}
