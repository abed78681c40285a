// K3 / K4 / K6 / K7: block SpMV, fused BLAS-1, the BiCGSTAB driver, and the Newmark vector kernels.
//
// Reference call sites replaced (rcapillon/thermoelasticity_fem):
//   solvers.py:26, :172   spsolve(K_ff, F_f), spsolve(K_dyn, rhs)   -> tefem_bicgstab
//   solvers.py:167-178    Newmark right-hand side and updates        -> tefem_newmark_predict / _correct
//   model.py:120-286      lumped loads                               -> tefem_load_axpy
//
// BiCGSTAB is run on the block-Jacobi left-scaled system (tefem_block_jacobi_scale), so no
// preconditioner application appears in the iteration.  One iteration is
//     v = A p            + partial (rhat.v)                      [spmv<1>]
//     alpha = rho / (rhat.v)                                      [reduce + scalars]
//     s = r - alpha v                                             [axpy]
//     t = A s            + partials (t.s, t.t, s.s per field, rhat.s, rhat.t)  [spmv<8>]
//     omega = t.s / t.t ; rho' = rhat.s - omega rhat.t ; beta = (rho'/rho)(alpha/omega)
//     |r_f|^2 = s.s_f - 2 omega t.s_f + omega^2 t.t_f  (f = u, theta)  [reduce + scalars]
//     x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v)   [fused update]
// i.e. two reduction points per iteration (each one packed all-reduce in the multi-GPU case), all
// scalars stay on the device, and the host only polls a stop flag every `check_every` iterations.
// Dot products are reduced in a fixed order (per-CTA partials, then one CTA over the partials):
// results are bitwise reproducible for a given grid.
#include "tefem_common.cuh"
#include "tefem_comm.h"
#include "tefem_precond.h"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace cg = cooperative_groups;

namespace tefem {

constexpr int SPMV_THREADS = 256;
// 4 resident CTAs per SM (<= 64 registers) and a grid of exactly one resident wave: ncu showed spmv<1> at 66
// registers -> 3 CTAs per SM, 2.67 waves of the old 1184-CTA grid and a 13 % longer launch than spmv<8> (64
// registers).  5 CTAs per SM (48 registers) spills and is slower: live 2.61 ms vs 2.43 ms at config B
// (profiles/README.md).
#ifndef SPMV_MIN_CTAS
#define SPMV_MIN_CTAS 4
#endif
constexpr int MAX_PARTIAL_CTAS = 148 * 8;   // grid of the reducing kernels: 8 CTAs of 256 threads per SM
constexpr int NSCAL = 24;

// device scalar block
// (SC_RR / SC_BB hold max_f |r_f|^2 / thr_f and 1: the convergence measure is per field, see scalars_omega_kernel)
enum { SC_RHO = 0, SC_ALPHA, SC_OMEGA, SC_BETA, SC_RR, SC_BB, SC_RHV, SC_THR_U = 7, SC_PKT = 8 /* 8 packet slots 8..15 */,
       SC_THR_T = 16, SC_BNORM2 = 17 };
// device int block
enum { IN_STOP = 0, IN_STOP_ITER = 1, IN_ITERS = 2, IN_REASON = 3,   // reason: 1 converged, 2 breakdown
       IN_TICKET = 8 };                                                  // CTA ticket of push_halo_kernel

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// x is rewritten between launches but never during one: the non-coherent path is safe inside a kernel.
__device__ __forceinline__ void ld256_x(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// y = A x over node rows; 4 consecutive threads own the 4 scalar rows of a node.
// DOTS = 0: none; 1: partial of (w . y); 8: partials of (y.s_u, y.s_t, y.y_u, y.y_t, s.s_u, s.s_t, w.s, w.y) with
// s = sd (the un-preconditioned vector whose image y is; sd == x without a preconditioner) and the first three
// split by field (u rows r < 3, theta rows r == 3).
// rows: optional list of node rows (interior / boundary split), else rows [0, n_rows).
// The grid-stride row loop of the SpMV; d[] receives this thread's dot-product contributions.
__device__ __forceinline__ void ld256_cg(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// XCG: x is read through L2 only (rows whose columns another GPU writes WHILE the kernel runs: spmv_fused_halo_kernel)
template <int DOTS, int UNROLL, bool XCG = false>
__device__ __forceinline__ void spmv_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                          const int32_t* __restrict__ rows, int32_t n_rows, const double* __restrict__ sys,
                                          const double* __restrict__ x, double* __restrict__ y,
                                          const double* __restrict__ w, const double* __restrict__ sd, double (&d)[5],
                                          int32_t row0 = 0) {
    const int64_t n = 4 * (int64_t)n_rows;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = rows ? rows[tid >> 2] : row0 + (int32_t)(tid >> 2);
        const int r = (int)(tid & 3);
        const int32_t s = rowptr[i], t = rowptr[i + 1];
        double acc0 = 0.0, acc1 = 0.0;
        int32_t p = s;
        // UNROLL blocks per trip: all their 256-bit loads are issued before the first use
        for (; p + UNROLL <= t; p += UNROLL) {
            double a[UNROLL][4], v[UNROLL][4];
            int32_t j[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) j[u] = colidx[p + u];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) ld256(sys + 16 * (int64_t)(p + u) + 4 * r, a[u][0], a[u][1], a[u][2], a[u][3]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (XCG) ld256_cg(x + 4 * (int64_t)j[u], v[u][0], v[u][1], v[u][2], v[u][3]);
                else ld256_x(x + 4 * (int64_t)j[u], v[u][0], v[u][1], v[u][2], v[u][3]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const double c = a[u][0] * v[u][0] + a[u][1] * v[u][1] + a[u][2] * v[u][2] + a[u][3] * v[u][3];
                if (u & 1) acc1 += c; else acc0 += c;
            }
        }
        for (; p < t; ++p) {
            double a0, a1, a2, a3, x0, x1, x2, x3;
            ld256(sys + 16 * (int64_t)p + 4 * r, a0, a1, a2, a3);
            if (XCG) ld256_cg(x + 4 * (int64_t)colidx[p], x0, x1, x2, x3);
            else ld256_x(x + 4 * (int64_t)colidx[p], x0, x1, x2, x3);
            acc0 += a0 * x0 + a1 * x1 + a2 * x2 + a3 * x3;
        }
        const double yv = acc0 + acc1;
        const int64_t row = 4 * (int64_t)i + r;
        y[row] = yv;
        if (DOTS == 1) d[0] += w[row] * yv;
        if (DOTS == 8) {
            const double sv = sd[row], wv = w[row];
            d[0] += yv * sv; d[1] += yv * yv; d[2] += sv * sv; d[3] += wv * sv; d[4] += wv * yv;
        }
    }
}

// CTA sums of the dot-product contributions -> partials[blockIdx.x * DOTS + k] (fixed reduction tree).
template <int DOTS>
__device__ __forceinline__ void spmv_store_partials(const double (&d)[5], double* red, double* __restrict__ partials) {
    if (DOTS == 1) {
        double v1[1] = {d[0]};
        block_reduce_sum<1>(v1, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = v1[0];
    }
    if (DOTS == 8) {
        // a thread always owns the same scalar row type (the grid stride is a multiple of 4)
        const bool th = (threadIdx.x & 3) == 3;
        double v8[8] = {th ? 0.0 : d[0], th ? d[0] : 0.0, th ? 0.0 : d[1], th ? d[1] : 0.0,
                        th ? 0.0 : d[2], th ? d[2] : 0.0, d[3], d[4]};
        block_reduce_sum<8>(v8, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) partials[(size_t)blockIdx.x * 8 + k] = v8[k];
        }
    }
}

// Flags of one halo exchange (peer-memory halo, tefem_peer.cuh; resolved on the host so that the kernels index nothing
// dynamically): `mine[q]` = the flag neighbour q raises in THIS rank's region when its boundary values have landed in my
// vector; `theirs[q]` = the flag I raise in neighbour q's region.  Slab partitions have <= 2 neighbours.
struct HaloFlags {
    int n = 0;
    uint32_t epoch = 0;
    long long timeout = PEER_TIMEOUT_DEFAULT;
    int* err = nullptr;
    const uint32_t* mine[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t* theirs[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// Threads 0 .. n-1 of the CTA wait for one neighbour flag each (bounded; a time-out marks the communicator failed and
// later waits return at once), then the CTA synchronises: every thread may read the halo tail afterwards.
__device__ __forceinline__ void halo_wait(const HaloFlags& hf) {
#pragma unroll
    for (int q = 0; q < PEER_MAX_RANKS; ++q) {
        if (q < hf.n && (int)threadIdx.x == q) {
            const uint32_t* f = hf.mine[q];
            if ((int32_t)(ld_acquire_sys(f) - hf.epoch) < 0 && !*reinterpret_cast<volatile int*>(hf.err)) {
                const long long t0 = clock64();
                while ((int32_t)(ld_acquire_sys(f) - hf.epoch) < 0) {
                    if (clock64() - t0 > hf.timeout) { *hf.err = 1; break; }
                }
            }
        }
    }
    __syncthreads();
}

// HALO: multi-GPU form.  The halo tail of x (rows n_rows ..) is written by the neighbouring ranks' push_halo_kernel
// straight into this rank's peer-mapped vector (remote stores over NVLink); the kernel only has to wait for their
// "pushed" flags of this exchange (an acquire at system scope, which also drops stale L1 lines) before it reads x -
// no packing, no NCCL call, no second launch for the boundary rows.
template <int DOTS, int UNROLL, bool HALO>
__global__ void __launch_bounds__(SPMV_THREADS, SPMV_MIN_CTAS) spmv_kernel(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx,
                                                            const int32_t* __restrict__ rows, int32_t n_rows,
                                                            const double* __restrict__ sys, const double* __restrict__ x,
                                                            double* __restrict__ y, const double* __restrict__ w,
                                                            const double* __restrict__ sd, double* __restrict__ partials,
                                                            const int* __restrict__ ictl, const HaloFlags hf) {
    __shared__ double red[8 * 32];
    double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const bool active = (ictl == nullptr) || (ictl[IN_STOP] == 0);
    if (HALO && active) halo_wait(hf);
    if (active) spmv_rows<DOTS, UNROLL>(rowptr, colidx, rows, n_rows, sys, x, y, w, sd, d);
    spmv_store_partials<DOTS>(d, red, partials);
}

// Halo push: entry k sends the 4 values of owned node idx[k] to node dpos[k] of rank dst[k]'s copy of the SAME vector
// slot (all ranks lay their slots out identically, tefem_comm_vec_alloc) with one 256-bit remote store.  The CTA that
// takes the last ticket - every other CTA's stores are then visible system-wide - raises this rank's flag of the
// exchange on the neighbours.  The consumer is the neighbour's next SpMV (or pull_halo_kernel).
struct HaloSlots {
    double* base[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
__global__ void __launch_bounds__(256) push_halo_kernel(int64_t n, const int32_t* __restrict__ idx,
                                                        const int32_t* __restrict__ dst, const int32_t* __restrict__ dpos,
                                                        const double* __restrict__ x, const HaloSlots slots,
                                                        const HaloFlags hf, const int* ictl, int* ticket) {
    if (ictl && ictl[IN_STOP]) return;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        double a, b, c, d;
        ld256_x(x + 4 * (int64_t)idx[k], a, b, c, d);
        double* o = slots.base[dst[k]] + 4 * (int64_t)dpos[k];
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        if (t == (int)gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
#pragma unroll
            for (int q = 0; q < PEER_MAX_RANKS; ++q)
                if (q < hf.n) st_release_sys(hf.theirs[q], hf.epoch);
        }
    }
}

// SpMV FUSED WITH ITS HALO EXCHANGE (default multi-GPU form): one kernel is producer and consumer.  The first CTAs store
// this rank's boundary values of x into the neighbours' copies of the vector (as push_halo_kernel) and the one that takes
// the last ticket raises the flags; then every CTA multiplies the INTERIOR rows [lo, hi) - rows without halo columns, a
// contiguous range for row blocks in node order - and only then waits for the neighbours' flags and multiplies the
// boundary rows [0, lo) and [hi, n_rows): the NVLink transfer and the skew between the ranks hide behind the interior rows,
// and the separate push launch is gone.
struct HaloPush {
    int64_t n = 0;
    const int32_t *idx = nullptr, *dst = nullptr, *dpos = nullptr;
    HaloSlots slots;
    int* ticket = nullptr;
    int n_ctas = 0;              // CTAs that push (and take a ticket)
    int32_t lo = 0, hi = 0;      // interior rows
};
template <int DOTS, int UNROLL>
__global__ void __launch_bounds__(SPMV_THREADS, SPMV_MIN_CTAS) spmv_fused_halo_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t n_rows, const double* __restrict__ sys,
    const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ w, const double* __restrict__ sd,
    double* __restrict__ partials, const int* __restrict__ ictl, const HaloFlags hf, const HaloPush hp) {
    __shared__ double red[8 * 32];
    double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const bool active = (ictl == nullptr) || (ictl[IN_STOP] == 0);
    if (active) {
        if ((int)blockIdx.x < hp.n_ctas) {
            for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hp.n; k += (int64_t)hp.n_ctas * blockDim.x) {
                double a, b, c, e;
                ld256_x(x + 4 * (int64_t)hp.idx[k], a, b, c, e);
                double* o = hp.slots.base[hp.dst[k]] + 4 * (int64_t)hp.dpos[k];
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(a), "d"(b), "d"(c), "d"(e) : "memory");
            }
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                const int t = atomicAdd(hp.ticket, 1);
                if (t == hp.n_ctas - 1) {
                    *hp.ticket = 0;
                    __threadfence_system();
#pragma unroll
                    for (int q = 0; q < PEER_MAX_RANKS; ++q)
                        if (q < hf.n) st_release_sys(hf.theirs[q], hf.epoch);
                }
            }
        }
        spmv_rows<DOTS, UNROLL>(rowptr, colidx, nullptr, hp.hi - hp.lo, sys, x, y, w, sd, d, hp.lo);
        halo_wait(hf);
        spmv_rows<DOTS, UNROLL, true>(rowptr, colidx, nullptr, hp.lo, sys, x, y, w, sd, d, 0);
        spmv_rows<DOTS, UNROLL, true>(rowptr, colidx, nullptr, n_rows - hp.hi, sys, x, y, w, sd, d, hp.hi);
    }
    spmv_store_partials<DOTS>(d, red, partials);
}

// Consumer of a push outside the SpMV (tefem_solver_halo_exchange): waits for the neighbours' flags and copies the halo
// tail of the slot into the caller's vector.
__global__ void __launch_bounds__(256) pull_halo_kernel(int64_t n_tail, const double* __restrict__ slot_tail,
                                                        double* __restrict__ x_tail, const HaloFlags hf) {
    halo_wait(hf);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_tail; k += (int64_t)gridDim.x * blockDim.x)
        x_tail[k] = ld_volatile_f64(slot_tail + k);
}

// r = b - y ; rhat = r ; p = r ; partials (r.r_u, r.r_t, b.b_u, b.b_t, x.x_u, x.x_t)
__global__ void __launch_bounds__(256) init_residual_kernel(int64_t n, const double* __restrict__ b,
                                                            const double* __restrict__ y, const double* __restrict__ x,
                                                            double* __restrict__ r, double* __restrict__ rhat,
                                                            double* __restrict__ p, double* __restrict__ partials) {
    __shared__ double red[6 * 32];
    double rr = 0.0, bb = 0.0, xx = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double bv = b[k], rv = bv - y[k], xv = x[k];
        r[k] = rv; rhat[k] = rv; p[k] = rv;
        rr += rv * rv; bb += bv * bv; xx += xv * xv;
    }
    const bool th = (threadIdx.x & 3) == 3;
    double d[6] = {th ? 0.0 : rr, th ? rr : 0.0, th ? 0.0 : bb, th ? bb : 0.0, th ? 0.0 : xx, th ? xx : 0.0};
    block_reduce_sum<6>(d, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) partials[6 * blockIdx.x + k] = d[k];
    }
}

// Convergence is judged PER FIELD on the block-Jacobi-scaled residual: |r_u| <= rtol |b_u| and
// |r_theta| <= rtol |b_theta|.  A field whose right-hand side vanishes (e.g. theta rows with no thermal load)
// is measured against |x_f| of the current iterate instead - the rows have a unit diagonal block after the
// scaling, so |x_f| is the natural magnitude of that row block - and is left out while x_f is still 0; the
// true-residual check at every (re)start then brings it in.
// The u rows (~1e10) and theta rows (~1) of the coupled operator make a single norm blind to one of the
// fields; measured on data/sandwich.msh: one norm at 1e-10 leaves a 2e-7 error in u.
__device__ __forceinline__ double field_measure(double rr_u, double rr_t, double thr_u, double thr_t) {
    const double mu = thr_u > 0.0 ? rr_u / thr_u : 0.0;
    const double mt = thr_t > 0.0 ? rr_t / thr_t : 0.0;
    return fmax(mu, mt);
}

__device__ __forceinline__ void scalars_init(double* sc, int* ictl, double rtol2) {
    const double rr_u = sc[SC_PKT], rr_t = sc[SC_PKT + 1], bb_u = sc[SC_PKT + 2], bb_t = sc[SC_PKT + 3];
    const double xx_u = sc[SC_PKT + 4], xx_t = sc[SC_PKT + 5];
    const double bb = bb_u + bb_t;
    sc[SC_THR_U] = bb_u > 0.0 ? bb_u : xx_u;
    sc[SC_THR_T] = bb_t > 0.0 ? bb_t : xx_t;
    sc[SC_BNORM2] = bb;
    sc[SC_RHO] = rr_u + rr_t;
    sc[SC_RR] = field_measure(rr_u, rr_t, sc[SC_THR_U], sc[SC_THR_T]);
    sc[SC_BB] = 1.0;
    sc[SC_ALPHA] = 1.0; sc[SC_OMEGA] = 1.0; sc[SC_BETA] = 0.0;
    ictl[IN_STOP] = 0; ictl[IN_STOP_ITER] = -1; ictl[IN_REASON] = 0;
    if (sc[SC_RR] <= rtol2) { ictl[IN_STOP] = 1; ictl[IN_REASON] = 1; }
}

__device__ __forceinline__ void scalars_alpha(double* sc, int* ictl, int iter) {
    if (ictl[IN_STOP]) return;
    const double rhv = sc[SC_PKT];
    sc[SC_RHV] = rhv;
    if (rhv == 0.0 || !isfinite(rhv)) { ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = -2; ictl[IN_REASON] = 2; return; }
    sc[SC_ALPHA] = sc[SC_RHO] / rhv;
}

__device__ __forceinline__ void scalars_omega(double* sc, int* ictl, int iter, double rtol2) {
    if (ictl[IN_STOP]) return;
    const double ts_u = sc[SC_PKT], ts_t = sc[SC_PKT + 1], tt_u = sc[SC_PKT + 2], tt_t = sc[SC_PKT + 3];
    const double ss_u = sc[SC_PKT + 4], ss_t = sc[SC_PKT + 5], hs = sc[SC_PKT + 6], ht = sc[SC_PKT + 7];
    const double ts = ts_u + ts_t, tt = tt_u + tt_t;
    const double thr_u = sc[SC_THR_U], thr_t = sc[SC_THR_T];
    ictl[IN_ITERS] += 1;
    if (tt == 0.0 || !isfinite(tt)) {
        // t = A s = 0: s = 0 (converged at the half step) or breakdown
        sc[SC_OMEGA] = 0.0; sc[SC_BETA] = 0.0; sc[SC_RR] = field_measure(ss_u, ss_t, thr_u, thr_t);
        ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = iter; ictl[IN_REASON] = (sc[SC_RR] <= rtol2) ? 1 : 2;
        return;
    }
    const double omega = ts / tt;
    const double rho_new = hs - omega * ht;
    const double rr_u = fmax(ss_u - 2.0 * omega * ts_u + omega * omega * tt_u, 0.0);
    const double rr_t = fmax(ss_t - 2.0 * omega * ts_t + omega * omega * tt_t, 0.0);
    const double rr = field_measure(rr_u, rr_t, thr_u, thr_t);
    sc[SC_OMEGA] = omega;
    sc[SC_RR] = rr;
    const double rho = sc[SC_RHO], alpha = sc[SC_ALPHA];
    sc[SC_RHO] = rho_new;
    if (rr <= rtol2) {
        sc[SC_BETA] = 0.0;
        ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = iter; ictl[IN_REASON] = 1;
        return;
    }
    if (omega == 0.0 || rho_new == 0.0 || !isfinite(omega) || !isfinite(rho_new)) {
        sc[SC_BETA] = 0.0;
        ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = iter; ictl[IN_REASON] = 2;
        return;
    }
    sc[SC_BETA] = (rho_new / rho) * (alpha / omega);
}

// One reduction point of the iteration in ONE kernel (one CTA): fixed-order sum of the per-CTA partials, the
// all-reduce over the ranks through NVLink peer memory (tefem_peer.cuh: every rank adds the packets of all ranks
// in rank order, so the scalars - and the stop decision - are bitwise identical everywhere), and the BiCGSTAB
// scalar update.  MODE 0: (re)start, 1: alpha, 2: omega / beta / convergence.  The exchange is executed even after
// the stop flag is set: the epochs of all ranks must advance together.
template <int NV, int MODE>
__global__ void __launch_bounds__(256) reduce_scalars_kernel(const double* __restrict__ partials, int n_cta, double* sc,
                                                             int* ictl, int iter, double rtol2, const PeerCtx pc) {
    __shared__ double red[NV * 32];
    __shared__ double pk[PEER_MAX_RANKS][NV];
    double d[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) d[k] = 0.0;
    for (int c = threadIdx.x; c < n_cta; c += blockDim.x)
#pragma unroll
        for (int k = 0; k < NV; ++k) d[k] += partials[(size_t)c * NV + k];
    block_reduce_sum<NV>(d, red);
    if (pc.world > 1) {
        if (threadIdx.x == 0) {
            double* mine = peer_small(pc, pc.rank);
#pragma unroll
            for (int k = 0; k < NV; ++k) mine[k] = d[k];
            __threadfence_system();
        }
        __syncthreads();
        peer_signal(pc, threadIdx.x);
        peer_wait(pc, threadIdx.x);
        if ((int)threadIdx.x < pc.world) {
            const double* src = peer_small(pc, threadIdx.x);
#pragma unroll
            for (int k = 0; k < NV; ++k) pk[threadIdx.x][k] = ld_volatile_f64(src + k);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double a = 0.0;
                for (int r = 0; r < pc.world; ++r) a += pk[r][k];
                d[k] = a;
            }
        }
    }
    if (threadIdx.x == 0) {
        if (MODE == 0 || !ictl[IN_STOP]) {
#pragma unroll
            for (int k = 0; k < NV; ++k) sc[SC_PKT + k] = d[k];
        }
        if (MODE == 0) scalars_init(sc, ictl, rtol2);
        if (MODE == 1) scalars_alpha(sc, ictl, iter);
        if (MODE == 2) scalars_omega(sc, ictl, iter, rtol2);
    }
}

// s = r - alpha v
__global__ void __launch_bounds__(256) update_s_kernel(int64_t n, const double* __restrict__ r,
                                                       const double* __restrict__ v, double* __restrict__ s,
                                                       const double* __restrict__ sc, const int* __restrict__ ictl) {
    if (ictl[IN_STOP]) return;
    const double alpha = sc[SC_ALPHA];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        s[k] = r[k] - alpha * v[k];
}

// x += alpha ph + omega sh ; r = s - omega t ; p = r + beta (p - omega v)
// (ph = N p, sh = N s: the right-preconditioned directions; ph == p, sh == s without a preconditioner)
// (p / ph and s / sh name the SAME buffers when there is no preconditioner: no __restrict__ on them)
__global__ void __launch_bounds__(256) update_xrp_kernel(int64_t n, double* __restrict__ x, double* __restrict__ r,
                                                         double* p, const double* s,
                                                         const double* __restrict__ t, const double* __restrict__ v,
                                                         const double* ph, const double* sh,
                                                         const double* __restrict__ sc, const int* __restrict__ ictl,
                                                         int iter) {
    if (ictl[IN_STOP] && ictl[IN_STOP_ITER] != iter) return;
    const double alpha = sc[SC_ALPHA], omega = sc[SC_OMEGA], beta = sc[SC_BETA];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double pv = p[k], sv = s[k];
        const double phv = ph[k], shv = sh[k];
        x[k] += alpha * phv + omega * shv;
        const double rv = sv - omega * t[k];
        r[k] = rv;
        p[k] = rv + beta * (pv - omega * v[k]);
    }
}

// ---- persistent cooperative iteration for SMALL single-GPU systems ---------------------------------------------
// Below ~1e5 rows an iteration of the launch-per-operation driver is bound by launch latency (10 launches of 2-4 us
// around microsecond kernels: the reference's own example_transient_2 runs 458 iterations per step).  Here ONE
// cooperative kernel runs a whole batch of iterations: the grid stays resident, the phases are separated by
// grid.sync() (4 per iteration), every CTA reduces the per-CTA partials itself in the same fixed order, so all
// CTAs hold bitwise identical copies of the scalars and take the same (uniform) stop decision without a broadcast.
// Same arithmetic per row and per reduction as the multi-kernel path.
struct CoopArgs {
    const int32_t* rowptr; const int32_t* colidx; const double* sys;
    int32_t n_rows;
    double *x, *r, *rhat, *p, *v, *s, *t, *partials, *sc;
    int* ictl;
    int it0, it1;
    double rtol2;
};

template <int NV>
__device__ __forceinline__ void coop_reduce(const double* partials, int n_cta, double* red, double* out /* shared */) {
    double d[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) d[k] = 0.0;
    for (int c = threadIdx.x; c < n_cta; c += blockDim.x)
#pragma unroll
        for (int k = 0; k < NV; ++k) d[k] += __ldcg(partials + (size_t)c * NV + k);
    block_reduce_sum<NV>(d, red);
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = d[k];
}

__global__ void __launch_bounds__(SPMV_THREADS, 2) coop_bicgstab_kernel(const CoopArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[8 * 32];
    __shared__ double sc[NSCAL];
    __shared__ int ictl[4];
    if (threadIdx.x < NSCAL) sc[threadIdx.x] = a.sc[threadIdx.x];
    if (threadIdx.x < 4) ictl[threadIdx.x] = a.ictl[threadIdx.x];
    __syncthreads();
    const int64_t n = 4 * (int64_t)a.n_rows;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    double* part1 = a.partials;
    double* part8 = a.partials + MAX_PARTIAL_CTAS;
    for (int it = a.it0; it < a.it1; ++it) {
        if (ictl[IN_STOP]) break;                       // identical in every CTA
        {   // v = A p, (rhat . v)
            double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            spmv_rows<1, 2>(a.rowptr, a.colidx, nullptr, a.n_rows, a.sys, a.p, a.v, a.rhat, nullptr, d);
            spmv_store_partials<1>(d, red, part1);
        }
        grid.sync();
        coop_reduce<1>(part1, gridDim.x, red, sc + SC_PKT);
        if (threadIdx.x == 0) scalars_alpha(sc, ictl, it);
        __syncthreads();
        if (ictl[IN_STOP]) break;
        {   // s = r - alpha v
            const double alpha = sc[SC_ALPHA];
            for (int64_t k = t0; k < n; k += stride) a.s[k] = a.r[k] - alpha * a.v[k];
        }
        grid.sync();
        {   // t = A s and the 8 packed dots
            double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            spmv_rows<8, 2>(a.rowptr, a.colidx, nullptr, a.n_rows, a.sys, a.s, a.t, a.rhat, a.s, d);
            spmv_store_partials<8>(d, red, part8);
        }
        grid.sync();
        coop_reduce<8>(part8, gridDim.x, red, sc + SC_PKT);
        if (threadIdx.x == 0) scalars_omega(sc, ictl, it, a.rtol2);
        __syncthreads();
        {   // x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v)   (also in the stopping iteration)
            const double alpha = sc[SC_ALPHA], omega = sc[SC_OMEGA], beta = sc[SC_BETA];
            for (int64_t k = t0; k < n; k += stride) {
                const double pv = a.p[k], sv = a.s[k];
                a.x[k] += alpha * pv + omega * sv;
                const double rv = sv - omega * a.t[k];
                a.r[k] = rv;
                a.p[k] = rv + beta * (pv - omega * a.v[k]);
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x < NSCAL) a.sc[threadIdx.x] = sc[threadIdx.x];
        if (threadIdx.x < 4) a.ictl[threadIdx.x] = ictl[threadIdx.x];
    }
}

// ---- preconditioned CG for the SPD field systems of the staged steady-state solve ---------------------------------
// K of the reference has no theta -> u block (model.py:94-103: uu <- Kuu, ut <- Kut, tt <- Ktt), so
// spsolve(K_ff, F_f) (solvers.py:26) is a block back-substitution: K_tt theta = F_t, then K_uu u = F_u - K_ut theta,
// two SYMMETRIC POSITIVE DEFINITE solves.  CG needs one SpMV and one preconditioner application per iteration
// (BiCGSTAB: two of each) and cannot break down.  Preconditioner: the inverse 4x4 diagonal node blocks (block
// Jacobi per node, applied symmetrically as z = D^-1 r: the operator itself stays unscaled and symmetric).
// Convergence is judged like in tefem_bicgstab, per field on the block-Jacobi-scaled residual z = D^-1 r:
// |z_f| <= rtol |D^-1 b|_f, and verified with the true residual before returning.
// One persistent cooperative kernel runs a batch of iterations (3 grid.sync() per iteration); every CTA reduces
// the per-CTA partials itself in the same fixed order, so all CTAs hold identical scalars and stop together.
enum { CG_RZ = 0, CG_BB_U, CG_BB_T, CG_RR, CG_ALPHA, CG_BETA, CG_PKT = 8 };

struct CgArgs {
    const int32_t* rowptr; const int32_t* colidx; const double* sys; const double* dinv;
    int32_t n_rows;
    const double* b;
    double *x, *r, *z, *p, *q, *partials, *sc;
    int* ictl;
    int mode;          // 0: (re)start - r = b - A x, z = D^-1 r, p = z, scalars; 1: iterate [it0, it1)
    int it0, it1;
    double rtol2;
};

// z_row = sum_c dinv[i][row][c] r[4 i + c] for the scalar row `k` of this lane; the four lanes of a node exchange
// their r values with shuffles (all 32 lanes of the warp call this together).
__device__ __forceinline__ double quad_block_apply(const double* __restrict__ dinv, int64_t k, bool live, double rv) {
    const int lane = threadIdx.x & 31, q0 = lane & ~3;
    const double r0 = __shfl_sync(0xffffffffu, rv, q0), r1 = __shfl_sync(0xffffffffu, rv, q0 + 1);
    const double r2 = __shfl_sync(0xffffffffu, rv, q0 + 2), r3 = __shfl_sync(0xffffffffu, rv, q0 + 3);
    if (!live) return 0.0;
    double d0, d1, d2, d3;
    ld256(dinv + 4 * k, d0, d1, d2, d3);          // row (k & 3) of block (k >> 2): dinv[16 i + 4 r + c] = dinv[4 k + c]
    return d0 * r0 + d1 * r1 + d2 * r2 + d3 * r3;
}

__global__ void __launch_bounds__(SPMV_THREADS, 2) coop_cg_kernel(const CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[8 * 32];
    __shared__ double sc[NSCAL];
    __shared__ int ictl[4];
    if (threadIdx.x < NSCAL) sc[threadIdx.x] = a.sc[threadIdx.x];
    if (threadIdx.x < 4) ictl[threadIdx.x] = a.ictl[threadIdx.x];
    __syncthreads();
    const int64_t n = 4 * (int64_t)a.n_rows;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t w0 = t0 - (threadIdx.x & 31);                       // first scalar row of this warp's trip
    const bool th = (threadIdx.x & 3) == 3;                           // a thread always owns the same row type
    double* part = a.partials;
    if (a.mode == 0) {
        {   // q = A x
            double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            spmv_rows<0, 2>(a.rowptr, a.colidx, nullptr, a.n_rows, a.sys, a.x, a.q, nullptr, nullptr, d);
        }
        grid.sync();
        double rz = 0.0, zz = 0.0, bb = 0.0;
        for (int64_t w = w0; w < n; w += stride) {
            const int64_t k = w + (threadIdx.x & 31);
            const bool live = k < n;
            const double bv = live ? a.b[k] : 0.0;
            const double rv = live ? bv - a.q[k] : 0.0;
            const double zv = quad_block_apply(a.dinv, k, live, rv);
            const double sb = quad_block_apply(a.dinv, k, live, bv);
            if (live) { a.r[k] = rv; a.z[k] = zv; a.p[k] = zv; }
            rz += rv * zv; zz += zv * zv; bb += sb * sb;
        }
        double v[5] = {rz, th ? 0.0 : zz, th ? zz : 0.0, th ? 0.0 : bb, th ? bb : 0.0};
        block_reduce_sum<5>(v, red);
        if (threadIdx.x == 0)
#pragma unroll
            for (int j = 0; j < 5; ++j) part[(size_t)blockIdx.x * 5 + j] = v[j];
        grid.sync();
        coop_reduce<5>(part, gridDim.x, red, sc + CG_PKT);
        if (threadIdx.x == 0) {
            const double bb_u = sc[CG_PKT + 3], bb_t = sc[CG_PKT + 4];
            sc[CG_RZ] = sc[CG_PKT]; sc[CG_BB_U] = bb_u; sc[CG_BB_T] = bb_t;
            sc[SC_BNORM2] = bb_u + bb_t;
            sc[CG_RR] = field_measure(sc[CG_PKT + 1], sc[CG_PKT + 2], bb_u, bb_t);
            ictl[IN_STOP] = 0; ictl[IN_STOP_ITER] = -1; ictl[IN_REASON] = 0;
            if (sc[CG_RR] <= a.rtol2) { ictl[IN_STOP] = 1; ictl[IN_REASON] = 1; }
        }
        __syncthreads();
    } else {
        for (int it = a.it0; it < a.it1; ++it) {
            if (ictl[IN_STOP]) break;                       // identical in every CTA
            {   // q = A p, (p . q)
                double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                spmv_rows<1, 2>(a.rowptr, a.colidx, nullptr, a.n_rows, a.sys, a.p, a.q, a.p, nullptr, d);
                spmv_store_partials<1>(d, red, part);
            }
            grid.sync();
            coop_reduce<1>(part, gridDim.x, red, sc + CG_PKT);
            if (threadIdx.x == 0) {
                const double pq = sc[CG_PKT];
                if (!(pq > 0.0) || !isfinite(pq)) { ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = -2; ictl[IN_REASON] = 2; }
                else sc[CG_ALPHA] = sc[CG_RZ] / pq;
            }
            __syncthreads();
            if (ictl[IN_STOP]) break;
            {   // x += alpha p ; r -= alpha q ; z = D^-1 r ; (r . z), |z|^2 per field
                const double alpha = sc[CG_ALPHA];
                double rz = 0.0, zz = 0.0;
                for (int64_t w = w0; w < n; w += stride) {
                    const int64_t k = w + (threadIdx.x & 31);
                    const bool live = k < n;
                    double rv = 0.0;
                    if (live) {
                        a.x[k] += alpha * a.p[k];
                        rv = a.r[k] - alpha * a.q[k];
                        a.r[k] = rv;
                    }
                    const double zv = quad_block_apply(a.dinv, k, live, rv);
                    if (live) a.z[k] = zv;
                    rz += rv * zv; zz += zv * zv;
                }
                double v[3] = {rz, th ? 0.0 : zz, th ? zz : 0.0};
                block_reduce_sum<3>(v, red);
                if (threadIdx.x == 0)
#pragma unroll
                    for (int j = 0; j < 3; ++j) part[(size_t)(gridDim.x + blockIdx.x) * 3 + j] = v[j];
            }
            grid.sync();
            coop_reduce<3>(part + (size_t)gridDim.x * 3, gridDim.x, red, sc + CG_PKT);
            if (threadIdx.x == 0) {
                const double rz_new = sc[CG_PKT];
                ictl[IN_ITERS] += 1;
                sc[CG_RR] = field_measure(sc[CG_PKT + 1], sc[CG_PKT + 2], sc[CG_BB_U], sc[CG_BB_T]);
                if (sc[CG_RR] <= a.rtol2) { ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = it; ictl[IN_REASON] = 1; }
                else if (!(rz_new > 0.0) || !isfinite(rz_new)) { ictl[IN_STOP] = 1; ictl[IN_STOP_ITER] = it; ictl[IN_REASON] = 2; }
                else { sc[CG_BETA] = rz_new / sc[CG_RZ]; sc[CG_RZ] = rz_new; }
            }
            __syncthreads();
            if (ictl[IN_STOP]) break;
            {   // p = z + beta p
                const double beta = sc[CG_BETA];
                for (int64_t k = t0; k < n; k += stride) a.p[k] = a.z[k] + beta * a.p[k];
            }
            grid.sync();
        }
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x < NSCAL) a.sc[threadIdx.x] = sc[threadIdx.x];
        if (threadIdx.x < 4) a.ictl[threadIdx.x] = ictl[threadIdx.x];
    }
}

// send[k*4 + c] = x[4*idx[k] + c]
__global__ void __launch_bounds__(256) pack_halo_kernel(int64_t n_nodes, const int32_t* __restrict__ idx,
                                                        const double* __restrict__ x, double* __restrict__ send,
                                                        const int* __restrict__ ictl) {
    if (ictl && ictl[IN_STOP]) return;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 4 * n_nodes) return;
    send[k] = x[4 * (int64_t)idx[k >> 2] + (k & 3)];
}

__global__ void __launch_bounds__(256) newmark_predict_kernel(int64_t n, const double* __restrict__ x,
                                                              const double* __restrict__ v, const double* __restrict__ a,
                                                              double c_v, double c_xv, double c_xa,
                                                              const uint8_t* __restrict__ mask, double* __restrict__ w1,
                                                              double* __restrict__ w2) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (mask && mask[k]) { w1[k] = 0.0; w2[k] = 0.0; return; }
    const double vv = v[k], av = a[k];
    w1[k] = vv + c_v * av;
    w2[k] = x[k] + c_xv * vv + c_xa * av;
}

__global__ void __launch_bounds__(256) newmark_correct_kernel(int64_t n, double* __restrict__ x, double* __restrict__ v,
                                                              double* __restrict__ a, const double* __restrict__ an,
                                                              double dt, double gamma, double beta,
                                                              const uint8_t* __restrict__ mask) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (mask && mask[k]) { v[k] = 0.0; a[k] = 0.0; return; }
    const double vv = v[k], av = a[k], nv = an[k];
    // solvers.py:173-178, same association order
    v[k] = vv + (1.0 - gamma) * dt * av + gamma * dt * nv;
    x[k] = x[k] + dt * vv + 0.5 * (dt * dt) * ((1.0 - 2.0 * beta) * av + 2.0 * beta * nv);
    a[k] = nv;
}

// out = c0 x0 + c1 x1 + c2 x2 (x1 / x2 may be null): extrapolated initial guess for the next time step
__global__ void __launch_bounds__(256) lincomb3_kernel(int64_t n, double c0, const double* __restrict__ x0, double c1,
                                                       const double* __restrict__ x1, double c2,
                                                       const double* __restrict__ x2, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double v = c0 * x0[k];
    if (x1) v += c1 * x1[k];
    if (x2) v += c2 * x2[k];
    out[k] = v;
}

__global__ void __launch_bounds__(256) load_axpy_kernel(int64_t n, const int32_t* __restrict__ node,
                                                        const double* __restrict__ w, double c0, double c1, double c2,
                                                        double c3, double* __restrict__ F) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double wk = w[k];
    double* f = F + 4 * (int64_t)node[k];
    if (c0 != 0.0) f[0] += wk * c0;
    if (c1 != 0.0) f[1] += wk * c1;
    if (c2 != 0.0) f[2] += wk * c2;
    if (c3 != 0.0) f[3] += wk * c3;
}

// same with the four amplitudes read from device memory (end-to-end path: the step's amplitudes arrive by an H2D copy)
__global__ void __launch_bounds__(256) load_axpy_dev_kernel(int64_t n, const int32_t* __restrict__ node,
                                                            const double* __restrict__ w, const double* __restrict__ coef,
                                                            double* __restrict__ F) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double wk = w[k], c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3];
    double* f = F + 4 * (int64_t)node[k];
    if (c0 != 0.0) f[0] += wk * c0;
    if (c1 != 0.0) f[1] += wk * c1;
    if (c2 != 0.0) f[2] += wk * c2;
    if (c3 != 0.0) f[3] += wk * c3;
}

}  // namespace tefem

using namespace tefem;

struct tefem_solver {
    int32_t n_rows = 0, n_halo = 0;
    int64_t n_own = 0, n_loc = 0;   // scalar sizes: owned, owned + halo
    double* work = nullptr;
    double *r = nullptr, *rhat = nullptr, *p = nullptr, *v = nullptr, *s = nullptr, *t = nullptr;
    double *ph = nullptr, *sh = nullptr;   // preconditioned directions (used only with a preconditioner)
    tefem_precond* pc = nullptr;
    double *partials = nullptr, *sc = nullptr, *sendbuf = nullptr;
    int* ictl = nullptr;
    int* h_ictl = nullptr;      // pinned
    double* h_sc = nullptr;     // pinned
    tefem_comm* comm = nullptr;
    // halo
    int n_nbr = 0;
    std::vector<int32_t> nbr, send_ptr, recv_ptr;
    const int32_t* send_idx = nullptr;
    // peer-memory halo (default): destination rank / node position of every send entry; the SpMV input vectors live in
    // the communicator's peer-mapped slots
    const int32_t* send_dst = nullptr;
    const int32_t* send_dpos = nullptr;
    bool peer_halo = false;
    double *xs = nullptr, *xa = nullptr;       // staging slots: restart / true-residual SpMV input, generic exchange
    // NCCL halo (measurement alternative: communicator without vector slots): interior / boundary row split
    const int32_t* boundary_rows = nullptr;
    int32_t n_boundary_rows = 0;
    const int32_t* interior_rows = nullptr;
    int32_t n_interior_rows = 0;
    bool fused_halo = true;                  // peer halo: the SpMV kernel pushes the boundary values itself
    int32_t interior_lo = 0, interior_hi = 0;   // rows [lo, hi) have no halo column (tefem_solver_set_interior_range)
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
    // optional SpMV timing (bench.py roofline): event pairs around every SpMV launch, harvested at the
    // host polls of the stop flag
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cat;      // category of every event pair (PROF_*)
    size_t ev_used = 0;
    double spmv_ms = 0.0;
    long long spmv_count = 0;
    int last_iters = 0;           // iterations of the previous solve: where the host starts polling the stop flag
    bool profile_all = false;     // TEFEM_PROFILE_ALL=1: also time the other parts of an iteration (diagnostics, stderr)
    double cat_ms[4] = {0.0, 0.0, 0.0, 0.0};
    long long cat_count[4] = {0, 0, 0, 0};
};

enum { SLOT_P = 0, SLOT_S = 1, SLOT_PH = 2, SLOT_SH = 3, SLOT_XS = 4, SLOT_XA = 5, SLOT_COUNT = 6 };
extern "C" int tefem_solver_peer_slots() { return SLOT_COUNT; }

enum { PROF_SPMV = 0, PROF_PRECOND = 1, PROF_REDUCE = 2, PROF_UPDATE = 3 };
static int prof_begin(tefem_solver* s, cudaStream_t st, int cat = PROF_SPMV) {
    if (!s->profile || (cat != PROF_SPMV && !s->profile_all)) return TEFEM_OK;
    if (s->ev_cat.size() < s->ev_used / 2 + 1) s->ev_cat.resize(s->ev_used / 2 + 1);
    s->ev_cat[s->ev_used / 2] = cat;
    if (s->ev_used + 2 > s->ev_pool.size()) {
        for (int k = 0; k < 2; ++k) {
            cudaEvent_t e;
            TEFEM_CUDA_CHECK(cudaEventCreate(&e));
            s->ev_pool.push_back(e);
        }
    }
    TEFEM_CUDA_CHECK(cudaEventRecord(s->ev_pool[s->ev_used], st));
    return TEFEM_OK;
}
static int prof_end(tefem_solver* s, cudaStream_t st, int cat = PROF_SPMV) {
    if (!s->profile || (cat != PROF_SPMV && !s->profile_all)) return TEFEM_OK;
    TEFEM_CUDA_CHECK(cudaEventRecord(s->ev_pool[s->ev_used + 1], st));
    s->ev_used += 2;
    return TEFEM_OK;
}
// call only after the stream has been synchronised; only the first n_valid SpMVs of the batch did real work (the
// launches queued after the stop flag was raised return immediately and must not dilute the average)
static int prof_harvest(tefem_solver* s, size_t n_valid) {
    size_t seen[4] = {0, 0, 0, 0};      // every category has two pairs per iteration
    for (size_t k = 0; k + 1 < s->ev_used; k += 2) {
        const int cat = s->ev_cat[k / 2];
        if (seen[cat]++ >= n_valid) continue;
        float ms = 0.f;
        TEFEM_CUDA_CHECK(cudaEventElapsedTime(&ms, s->ev_pool[k], s->ev_pool[k + 1]));
        s->cat_ms[cat] += ms;
        s->cat_count[cat] += 1;
        if (cat == PROF_SPMV) { s->spmv_ms += ms; s->spmv_count += 1; }
    }
    s->ev_used = 0;
    return TEFEM_OK;
}

static int64_t vec_len(int32_t n_rows, int32_t n_halo) { return 4 * ((int64_t)n_rows + n_halo); }

static int64_t align16(int64_t n) { return (n + 15) / 16 * 16; }   // keep every vector 128-byte aligned

extern "C" int64_t tefem_solver_work_doubles(int32_t n_rows, int32_t n_halo) {
    const int64_t nl = align16(vec_len(n_rows, n_halo));
    return 8 * nl + align16((int64_t)MAX_PARTIAL_CTAS * 2 * 8) + align16(NSCAL) + align16(16) + align16(4 * (int64_t)n_rows);
}

extern "C" int tefem_solver_create(tefem_solver** out, int32_t n_rows, int32_t n_halo, double* work, void* comm) {
    TEFEM_REQUIRE(out && work, "null pointer");
    TEFEM_REQUIRE((((uintptr_t)work) & 127) == 0, "work must be 128-byte aligned");
    tefem_solver* s = new tefem_solver();
    s->n_rows = n_rows; s->n_halo = n_halo;
    s->n_own = 4 * (int64_t)n_rows; s->n_loc = vec_len(n_rows, n_halo);
    s->work = work;
    const int64_t nl = align16(s->n_loc);
    double* w = work;
    s->r = w; w += nl; s->rhat = w; w += nl; s->p = w; w += nl; s->v = w; w += nl; s->s = w; w += nl; s->t = w; w += nl;
    s->ph = w; w += nl; s->sh = w; w += nl;
    s->partials = w; w += align16((int64_t)MAX_PARTIAL_CTAS * 2 * 8);
    s->sc = w; w += align16(NSCAL);
    s->ictl = reinterpret_cast<int*>(w); w += align16(16);
    s->sendbuf = w;
    s->comm = reinterpret_cast<tefem_comm*>(comm);
    if (s->comm && tefem_comm_has_vectors(s->comm)) {
        // SpMV inputs (p, s, their preconditioned images, two staging vectors) in the peer-mapped slots: the
        // neighbours write the halo tails directly
        TEFEM_REQUIRE(tefem_comm_n_slots(s->comm) >= SLOT_COUNT && tefem_comm_slot_doubles(s->comm) >= nl,
                      "peer-memory vector slots too few / too small (tefem_comm_vec_alloc)");
        const int me = tefem_comm_rank(s->comm);
        s->p = tefem_comm_slot(s->comm, me, SLOT_P);
        s->s = tefem_comm_slot(s->comm, me, SLOT_S);
        s->ph = tefem_comm_slot(s->comm, me, SLOT_PH);
        s->sh = tefem_comm_slot(s->comm, me, SLOT_SH);
        s->xs = tefem_comm_slot(s->comm, me, SLOT_XS);
        s->xa = tefem_comm_slot(s->comm, me, SLOT_XA);
        s->peer_halo = true;
    }
    TEFEM_CUDA_CHECK(cudaMallocHost(&s->h_ictl, 8 * sizeof(int)));     // [0..3] control block, [4] peer time-out flag
    s->h_ictl[4] = 0;
    TEFEM_CUDA_CHECK(cudaMallocHost(&s->h_sc, NSCAL * sizeof(double)));
    if (s->comm) {
        TEFEM_CUDA_CHECK(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
        TEFEM_CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming));
        TEFEM_CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming));
    }
    *out = s;
    return TEFEM_OK;
}

extern "C" int tefem_solver_profile(tefem_solver* s, int enable, double* h_spmv_ms, int64_t* h_spmv_count) {
    TEFEM_REQUIRE(s, "null solver");
    if (s->profile_all && s->cat_count[0] > 0)
        fprintf(stderr, "tefem profile (ms per call): spmv %.4f x %lld | precond %.4f x %lld | reduce+scalars %.4f x %lld | "
                        "vector updates %.4f x %lld\n",
                s->cat_ms[0] / s->cat_count[0], s->cat_count[0], s->cat_ms[1] / (s->cat_count[1] ? s->cat_count[1] : 1),
                s->cat_count[1], s->cat_ms[2] / (s->cat_count[2] ? s->cat_count[2] : 1), s->cat_count[2],
                s->cat_ms[3] / (s->cat_count[3] ? s->cat_count[3] : 1), s->cat_count[3]);
    if (h_spmv_ms) *h_spmv_ms = s->spmv_ms;
    if (h_spmv_count) *h_spmv_count = s->spmv_count;
    if (enable >= 0) {
        s->profile = enable != 0;
        s->spmv_ms = 0.0;
        s->spmv_count = 0;
        for (int k = 0; k < 4; ++k) { s->cat_ms[k] = 0.0; s->cat_count[k] = 0; }
        const char* pa = getenv("TEFEM_PROFILE_ALL");
        s->profile_all = pa && pa[0] == '1';
    }
    return TEFEM_OK;
}

extern "C" int tefem_solver_destroy(tefem_solver* s) {
    if (!s) return TEFEM_OK;
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    if (s->h_ictl) cudaFreeHost(s->h_ictl);
    if (s->h_sc) cudaFreeHost(s->h_sc);
    if (s->ev_ready) cudaEventDestroy(s->ev_ready);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    delete s;
    return TEFEM_OK;
}

// Halo description of a row-partitioned solve.  h_nbr / h_send_ptr / h_recv_ptr: neighbour ranks (ascending), the
// segments of send_idx (owned nodes whose values neighbour k needs, ascending global id) and of the halo tail.
// Peer-memory halo (the communicator has vector slots): send_dst[k] / send_dpos[k] = destination rank and node position
// in that rank's local vector of send entry k (device arrays; the position is n_own + halo offset THERE);
// boundary_rows / interior_rows may be NULL.  NCCL halo (no vector slots): send_dst / send_dpos may be NULL and the
// interior / boundary row lists are required (the exchange overlaps the interior rows).
extern "C" int tefem_solver_set_halo(tefem_solver* s, int32_t n_halo, int n_nbr, const int32_t* h_nbr,
                                     const int32_t* h_send_ptr, const int32_t* send_idx, const int32_t* send_dst,
                                     const int32_t* send_dpos, const int32_t* h_recv_ptr,
                                     const int32_t* boundary_rows, int32_t n_boundary_rows,
                                     const int32_t* interior_rows, int32_t n_interior_rows) {
    TEFEM_REQUIRE(s, "null solver");
    TEFEM_REQUIRE(n_halo == s->n_halo, "n_halo differs from tefem_solver_create");
    TEFEM_REQUIRE(n_nbr >= 0 && n_nbr <= PEER_MAX_RANKS, "0 .. 8 neighbour ranks");
    TEFEM_REQUIRE(n_nbr == 0 || (h_nbr && h_send_ptr && h_recv_ptr && send_idx), "null halo arrays");
    if (s->peer_halo) TEFEM_REQUIRE(n_nbr == 0 || (send_dst && send_dpos), "the peer-memory halo needs send_dst / send_dpos");
    else TEFEM_REQUIRE(n_boundary_rows + n_interior_rows == s->n_rows, "interior + boundary rows must cover the owned rows");
    s->n_nbr = n_nbr;
    s->nbr.assign(h_nbr, h_nbr + n_nbr);
    s->send_ptr.assign(h_send_ptr, h_send_ptr + n_nbr + 1);
    s->recv_ptr.assign(h_recv_ptr, h_recv_ptr + n_nbr + 1);
    TEFEM_REQUIRE(n_nbr == 0 || s->send_ptr[n_nbr] <= s->n_rows, "send list larger than the owned rows");
    s->send_idx = send_idx; s->send_dst = send_dst; s->send_dpos = send_dpos;
    s->boundary_rows = boundary_rows; s->n_boundary_rows = n_boundary_rows;
    s->interior_rows = interior_rows; s->n_interior_rows = n_interior_rows;
    return TEFEM_OK;
}

// Peer-memory halo: rows [lo, hi) have no halo column (a contiguous interior range; lo == hi: none known) - the fused
// SpMV multiplies them before it waits for the neighbours.  fused = 0 restores the separate push launch (comparison).
extern "C" int tefem_solver_set_interior_range(tefem_solver* s, int32_t lo, int32_t hi, int fused) {
    TEFEM_REQUIRE(s && lo >= 0 && lo <= hi && hi <= s->n_rows, "0 <= lo <= hi <= n_rows");
    s->interior_lo = lo; s->interior_hi = hi;
    s->fused_halo = fused != 0;
    return TEFEM_OK;
}

static int grid_for(int64_t n_threads, int per_cta) {
    int64_t g = ceil_div(n_threads, per_cta);
    if (g > MAX_PARTIAL_CTAS) g = MAX_PARTIAL_CTAS;
    if (g < 1) g = 1;
    return (int)g;
}

// Launch configuration of the SpMV (threads per CTA, CTAs per SM of the grid-stride grid, blocks per trip):
// tuned on the B200 at config B (profiles/README.md); tefem_spmv_set_config(threads, ctas_per_sm, unroll) overrides it
// for experiments (single-GPU launches; the multi-GPU path uses the tuned configuration).
struct SpmvCfg {
    int threads = 256, ctas_per_sm = 8, unroll = 2;
};
static SpmvCfg g_spmv_cfg;
static const SpmvCfg& spmv_cfg() { return g_spmv_cfg; }

template <int DOTS, int UNROLL, bool HALO>
static int spmv_resident_ctas(int threads, int* out) {
    static int cached_threads = 0, cached = 0;
    if (cached_threads != threads) {
        int occ = 0, dev = 0, sms = 0;
        TEFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_kernel<DOTS, UNROLL, HALO>, threads, 0));
        TEFEM_CUDA_CHECK(cudaGetDevice(&dev));
        TEFEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cached = (occ < 1 ? 1 : occ) * sms;
        cached_threads = threads;
    }
    *out = cached;
    return TEFEM_OK;
}

// Optional halo wait of an SpMV launch (peer-memory halo)
struct SpmvHalo {
    bool on = false;
    HaloFlags hf;
};

template <int DOTS, int UNROLL, bool HALO>
static int launch_spmv_u(const SpmvCfg& c, const int32_t* rowptr, const int32_t* colidx, const int32_t* rows, int32_t n_rows,
                         const double* sys, const double* x, double* y, const double* w, const double* sd, double* partials,
                         const int* ictl, int* n_cta_out, const SpmvHalo& h, cudaStream_t st) {
    // one resident wave of grid-striding CTAs (no tail), capped by the partial-sum buffers
    int resident = 0;
    int rc = spmv_resident_ctas<DOTS, UNROLL, HALO>(c.threads, &resident);
    if (rc) return rc;
    if (c.ctas_per_sm > 0 && resident > 148 * c.ctas_per_sm) resident = 148 * c.ctas_per_sm;
    if (resident > MAX_PARTIAL_CTAS) resident = MAX_PARTIAL_CTAS;
    int64_t g64 = ceil_div(4 * (int64_t)n_rows, c.threads);
    if (g64 > resident) g64 = resident;
    const int g = g64 < 1 ? 1 : (int)g64;
    if (n_cta_out) *n_cta_out = g;
    spmv_kernel<DOTS, UNROLL, HALO><<<g, c.threads, 0, st>>>(rowptr, colidx, rows, n_rows, sys, x, y, w, sd ? sd : x, partials, ictl,
                                                            h.hf);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

template <int DOTS>
static int launch_spmv(const int32_t* rowptr, const int32_t* colidx, const int32_t* rows, int32_t n_rows,
                       const double* sys, const double* x, double* y, const double* w, const double* sd, double* partials,
                       const int* ictl, int* n_cta_out, cudaStream_t st, const SpmvHalo& h = SpmvHalo()) {
    const SpmvCfg& c = spmv_cfg();
    if (h.on) {   // multi-GPU: the tuned configuration only
        return launch_spmv_u<DOTS, 2, true>(c, rowptr, colidx, rows, n_rows, sys, x, y, w, sd, partials, ictl, n_cta_out, h, st);
    }
    if (c.unroll == 4) return launch_spmv_u<DOTS, 4, false>(c, rowptr, colidx, rows, n_rows, sys, x, y, w, sd, partials, ictl, n_cta_out, h, st);
    if (c.unroll == 1) return launch_spmv_u<DOTS, 1, false>(c, rowptr, colidx, rows, n_rows, sys, x, y, w, sd, partials, ictl, n_cta_out, h, st);
    return launch_spmv_u<DOTS, 2, false>(c, rowptr, colidx, rows, n_rows, sys, x, y, w, sd, partials, ictl, n_cta_out, h, st);
}

// Flags and slot addresses of the next halo exchange of slot vector x.
static int halo_context(tefem_solver* s, const double* x, SpmvHalo* h, HaloSlots* hs) {
    const int me = tefem_comm_rank(s->comm);
    const int64_t off = x - tefem_comm_slot(s->comm, me, 0);
    TEFEM_REQUIRE(off >= 0 && off % tefem_comm_slot_doubles(s->comm) == 0 &&
                      off / tefem_comm_slot_doubles(s->comm) < tefem_comm_n_slots(s->comm),
                  "the SpMV input of a multi-GPU solve must be a peer-memory vector slot");
    const int slot = (int)(off / tefem_comm_slot_doubles(s->comm));
    h->on = true;
    const PeerCtx pc = tefem_comm_peer_ctx(s->comm, PEER_CH_HALO);      // next epoch of the halo channel
    h->hf.n = s->n_nbr; h->hf.epoch = pc.epoch; h->hf.timeout = pc.timeout; h->hf.err = pc.err;
    for (int k = 0; k < s->n_nbr; ++k) {
        h->hf.mine[k] = peer_flag(pc, pc.rank, s->nbr[k]);
        h->hf.theirs[k] = peer_flag(pc, s->nbr[k], pc.rank);
    }
    for (int r = 0; r < tefem_comm_world(s->comm); ++r) hs->base[r] = tefem_comm_slot(s->comm, r, slot);
    return TEFEM_OK;
}

// Peer-memory halo, fused form: ONE launch pushes this rank's boundary values, multiplies the interior rows, waits for the
// neighbours and multiplies the boundary rows (spmv_fused_halo_kernel).
template <int DOTS>
static int launch_spmv_fused(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys, const double* x,
                             double* y, const double* w, const double* sd, const int* ictl, int* n_cta_out, cudaStream_t st);

// Peer-memory halo, producer side: push the boundary values of slot vector x (one of this rank's slots) into the same
// slot of the neighbours and raise the flags of a new exchange, whose context is returned for the consumer.
static int push_halo(tefem_solver* s, const double* x, const int* ictl, SpmvHalo* h, cudaStream_t st) {
    HaloSlots hs;
    int rc = halo_context(s, x, h, &hs);
    if (rc) return rc;
    const int64_t n_send = s->send_ptr[s->n_nbr];
    const unsigned g = (unsigned)(n_send > 0 ? ceil_div(n_send, 256) : 1);
    push_halo_kernel<<<g, 256, 0, st>>>(n_send, s->send_idx, s->send_dst, s->send_dpos, x, hs, h->hf, ictl,
                                        s->ictl + IN_TICKET);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

template <int DOTS>
static int launch_spmv_fused(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys, const double* x,
                             double* y, const double* w, const double* sd, const int* ictl, int* n_cta_out, cudaStream_t st) {
    SpmvHalo h;
    HaloPush hp;
    int rc = halo_context(s, x, &h, &hp.slots);
    if (rc) return rc;
    int resident = 0;
    rc = spmv_resident_ctas<DOTS, 2, true>(SPMV_THREADS, &resident);     // same register budget as the HALO instantiation
    if (rc) return rc;
    if (resident > MAX_PARTIAL_CTAS) resident = MAX_PARTIAL_CTAS;
    int64_t g64 = ceil_div(4 * (int64_t)s->n_rows, SPMV_THREADS);
    if (g64 > resident) g64 = resident;
    const int g = g64 < 1 ? 1 : (int)g64;
    if (n_cta_out) *n_cta_out = g;
    hp.n = s->send_ptr[s->n_nbr];
    hp.idx = s->send_idx; hp.dst = s->send_dst; hp.dpos = s->send_dpos;
    hp.ticket = s->ictl + IN_TICKET;
    int64_t pc = ceil_div(hp.n > 0 ? hp.n : 1, SPMV_THREADS);
    hp.n_ctas = (int)(pc > g ? g : pc);                                  // >= 1: an empty send list still raises the flags
    hp.lo = s->interior_lo; hp.hi = s->interior_hi;
    spmv_fused_halo_kernel<DOTS, 2><<<g, SPMV_THREADS, 0, st>>>(rowptr, colidx, s->n_rows, sys, x, y, w, sd ? sd : x, s->partials, ictl,
                                                                h.hf, hp);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

// Exchange the halo part of x (owned rows listed in send_idx go out, halo rows come in).
static int halo_exchange(tefem_solver* s, double* x, const int* ictl, cudaStream_t st) {
    if (!s->comm || s->n_nbr == 0) return TEFEM_OK;
    if (s->peer_halo) {   // stage through slot XA: owned part in, push, wait, halo tail out
        TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->xa, x, sizeof(double) * (size_t)s->n_own, cudaMemcpyDeviceToDevice, st));
        SpmvHalo h;
        int rc = push_halo(s, s->xa, nullptr, &h, st);
        if (rc) return rc;
        const int64_t n_tail = s->n_loc - s->n_own;
        pull_halo_kernel<<<grid_for(n_tail, 256), 256, 0, st>>>(n_tail, s->xa + s->n_own, x + s->n_own, h.hf);
        TEFEM_LAUNCH_CHECK();
        return TEFEM_OK;
    }
    const int64_t n_send = s->send_ptr[s->n_nbr];
    if (n_send > 0) {
        pack_halo_kernel<<<(unsigned)ceil_div(4 * n_send, 256), 256, 0, st>>>(n_send, s->send_idx, x, s->sendbuf, ictl);
        TEFEM_LAUNCH_CHECK();
    }
    return tefem_comm_neighbor_exchange(s->comm, s->n_nbr, s->nbr.data(), s->send_ptr.data(), s->sendbuf,
                                        s->recv_ptr.data(), x + s->n_own, st);
}

// y = A x with the halo of x.  Peer-memory halo (default): push kernel + ONE SpMV launch over all rows in natural
// order that waits for the neighbours' flags.  NCCL halo (communicator without vector slots): pack + send/recv on a
// second stream overlapped with the interior rows, boundary rows after the halo event.  DOTS partials packed.
template <int DOTS>
static int dist_spmv(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys, double* x,
                     double* y, const double* w, const double* sd, const int* ictl, int* n_part_ctas, cudaStream_t st) {
    if (!s->comm || s->n_nbr == 0) {
        int rc0 = prof_begin(s, st);
        if (rc0) return rc0;
        rc0 = launch_spmv<DOTS>(rowptr, colidx, nullptr, s->n_rows, sys, x, y, w, sd, s->partials, ictl, n_part_ctas, st);
        if (rc0) return rc0;
        return prof_end(s, st);
    }
    int rc = prof_begin(s, st);
    if (rc) return rc;
    if (s->peer_halo && s->fused_halo) {
        rc = launch_spmv_fused<DOTS>(s, rowptr, colidx, sys, x, y, w, sd, ictl, n_part_ctas, st);
        if (rc) return rc;
        return prof_end(s, st);
    }
    if (s->peer_halo) {
        SpmvHalo h;
        rc = push_halo(s, x, ictl, &h, st);
        if (rc) return rc;
        rc = launch_spmv<DOTS>(rowptr, colidx, nullptr, s->n_rows, sys, x, y, w, sd, s->partials, ictl, n_part_ctas, st, h);
        if (rc) return rc;
        return prof_end(s, st);
    }
    // comm stream: pack + send/recv ; main stream: interior rows ; then boundary rows
    TEFEM_CUDA_CHECK(cudaEventRecord(s->ev_ready, st));
    TEFEM_CUDA_CHECK(cudaStreamWaitEvent(s->comm_stream, s->ev_ready, 0));
    rc = halo_exchange(s, x, ictl, s->comm_stream);
    if (rc) return rc;
    TEFEM_CUDA_CHECK(cudaEventRecord(s->ev_halo, s->comm_stream));
    int g0 = 0, g1 = 0;
    if (s->n_interior_rows > 0) {
        rc = launch_spmv<DOTS>(rowptr, colidx, s->interior_rows, s->n_interior_rows, sys, x, y, w, sd, s->partials, ictl, &g0, st);
        if (rc) return rc;
    }
    TEFEM_CUDA_CHECK(cudaStreamWaitEvent(st, s->ev_halo, 0));
    if (s->n_boundary_rows > 0) {
        rc = launch_spmv<DOTS>(rowptr, colidx, s->boundary_rows, s->n_boundary_rows, sys, x, y, w, sd,
                               s->partials + (size_t)g0 * (DOTS > 0 ? DOTS : 1), ictl, &g1, st);
        if (rc) return rc;
    }
    if (n_part_ctas) *n_part_ctas = g0 + g1;
    return prof_end(s, st);
}

template <int NV, int MODE>
static int reduce_scalars(tefem_solver* s, int n_cta, int iter, double rtol2, cudaStream_t st) {
    PeerCtx pc;                                           // world = 1: no exchange
    if (s->comm) {
        TEFEM_REQUIRE(tefem_comm_has_peer(s->comm), "multi-GPU solves need the peer-memory region (tefem_comm_shm_alloc / _open)");
        pc = tefem_comm_peer_ctx(s->comm, PEER_CH_DOTS);
    }
    reduce_scalars_kernel<NV, MODE><<<1, 256, 0, st>>>(s->partials, n_cta, s->sc, s->ictl, iter, rtol2, pc);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_solver_set_precond(tefem_solver* s, tefem_precond* pc) {
    TEFEM_REQUIRE(s, "null solver");
    s->pc = pc;
    return TEFEM_OK;
}

extern "C" int tefem_solver_halo_exchange(tefem_solver* s, double* x, void* stream) {
    TEFEM_REQUIRE(s && x, "null pointer");
    return halo_exchange(s, x, nullptr, as_stream(stream));
}

extern "C" int tefem_spmv_set_config(int threads, int ctas_per_sm, int unroll) {
    TEFEM_REQUIRE(threads >= 32 && threads <= 256 && (threads & 31) == 0, "threads per CTA: a multiple of 32 in 32 .. 256");
    TEFEM_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 8, "1 .. 8 CTAs per SM (the partial-sum buffers hold 148 * 8)");
    TEFEM_REQUIRE(unroll == 1 || unroll == 2 || unroll == 4, "blocks per trip: 1, 2 or 4");
    g_spmv_cfg.threads = threads; g_spmv_cfg.ctas_per_sm = ctas_per_sm; g_spmv_cfg.unroll = unroll;
    return TEFEM_OK;
}

extern "C" int tefem_spmv(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, const double* sys,
                          const double* x, double* y, void* stream) {
    TEFEM_REQUIRE(rowptr && colidx && sys && x && y, "null pointer");
    TEFEM_REQUIRE((((uintptr_t)sys) & 127) == 0 && (((uintptr_t)x) & 31) == 0, "sys / x alignment");
    if (n_rows == 0) return TEFEM_OK;
    return launch_spmv<0>(rowptr, colidx, nullptr, n_rows, sys, x, y, nullptr, nullptr, nullptr, nullptr, nullptr, as_stream(stream));
}

// Persistent cooperative path: systems small enough that an iteration is launch-latency bound.
constexpr int32_t COOP_MAX_ROWS = 65536;
static int coop_grid(int32_t n_rows) {
    static int resident = -1;
    if (resident < 0) {
        int dev = 0, sms = 0, per_sm = 0, coop_ok = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_bicgstab_kernel, SPMV_THREADS, 0) != cudaSuccess)
            resident = 0;
        else
            resident = coop_ok ? sms * per_sm : 0;
        if (resident > MAX_PARTIAL_CTAS) resident = MAX_PARTIAL_CTAS;
    }
    int64_t g = ceil_div(4 * (int64_t)n_rows, SPMV_THREADS);
    if (g > resident) g = resident;
    return (int)g;
}

extern "C" int tefem_bicgstab(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys,
                              const double* b, double* x, double rtol, int max_iter, int check_every, double floor_factor,
                              double* h_info, void* stream) {
    TEFEM_REQUIRE(s && rowptr && colidx && sys && b && x, "null pointer");
    TEFEM_REQUIRE((((uintptr_t)sys) & 127) == 0 && (((uintptr_t)x) & 31) == 0, "sys / x alignment");
    TEFEM_REQUIRE(rtol > 0 && max_iter > 0 && floor_factor >= 1.0, "rtol / max_iter / floor_factor");
    if (check_every < 1) check_every = 1;
    cudaStream_t st = as_stream(stream);
    const int64_t n = s->n_own;
    const int vgrid = grid_for(n, 256);
    const double rtol2 = rtol * rtol;
    const int max_restarts = 20;
    int total_iters = 0, restarts = 0, stalls = 0;
    bool floor_hit = false;
    double relres = -1.0, bnorm = 0.0, prev_relres = INFINITY;
    int status = TEFEM_ERR_NO_CONVERGE;
    TEFEM_CUDA_CHECK(cudaMemsetAsync(s->ictl, 0, 16 * sizeof(int), st));
    for (;;) {
        // ---- (re)start: r = b - A x, rhat = p = r
        int npc = 0;
        double* xin = x;
        if (s->peer_halo && s->n_nbr > 0) {   // the operator input must be a peer-memory slot: stage the owned part of x
            TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->xs, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
            xin = s->xs;
        }
        int rc = dist_spmv<0>(s, rowptr, colidx, sys, xin, s->t, nullptr, nullptr, nullptr, &npc, st);
        if (rc) return rc;
        init_residual_kernel<<<vgrid, 256, 0, st>>>(n, b, s->t, x, s->r, s->rhat, s->p, s->partials);
        TEFEM_LAUNCH_CHECK();
        rc = reduce_scalars<6, 0>(s, vgrid, 0, rtol2, st);
        if (rc) return rc;
        TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_ictl, s->ictl, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_sc, s->sc, NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
        TEFEM_CUDA_CHECK(cudaStreamSynchronize(st));
        bnorm = sqrt(s->h_sc[SC_BNORM2]);
        relres = sqrt(s->h_sc[SC_RR]);   // max over the fields of |r_f| / |b_f|
        if (bnorm == 0.0) {   // homogeneous system: the solution is 0
            TEFEM_CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
            relres = 0.0; status = TEFEM_OK; break;
        }
        if (!isfinite(relres)) { set_error("non-finite residual in BiCGSTAB"); status = TEFEM_ERR_BREAKDOWN; break; }
        if (s->h_ictl[IN_STOP] && s->h_ictl[IN_REASON] == 1) { status = TEFEM_OK; break; }   // true residual converged
        // the TRUE residual cannot be pushed below ~eps*cond in working precision: two restarts in a row
        // that fail to improve it mean that floor has been reached
        if (restarts > 0 && relres > 0.5 * prev_relres) ++stalls; else stalls = 0;
        prev_relres = relres;
        if (stalls >= 2) {      // accepted only within floor_factor * rtol, and reported (h_info[4])
            if (relres <= floor_factor * rtol) { status = TEFEM_OK; floor_hit = relres > rtol; }
            break;
        }
        if (total_iters >= max_iter || restarts > max_restarts) break;
        // ---- iterate
        bool stopped = false;
        int it = 0;
        const bool coop = !s->comm && !s->pc && !s->profile && s->n_rows <= COOP_MAX_ROWS && coop_grid(s->n_rows) > 0;
        if (s->profile) s->ev_used = 0;          // drop the (re)start SpMV from the timing
        int iters_seen = total_iters;
        // Polling schedule: the launches queued behind a raised stop flag are no-ops, but each still costs ~2 us, and every
        // poll drains the launch pipeline.  Consecutive solves of a time loop need nearly the same number of iterations,
        // so the first poll of a (re)start is placed a few iterations before the previous solve's count and the following
        // ones come at short intervals; without history (or after a restart) the regular check_every applies.
        const int fine_every = check_every < 4 ? check_every : 4;
        int first_poll = check_every;
        if (restarts == 0 && s->last_iters > 2 * check_every) first_poll = s->last_iters - fine_every;
        while (!stopped && total_iters + it < max_iter) {
            const int batch_end = it == 0 ? first_poll : it + (first_poll > check_every ? fine_every : check_every);
            if (coop) {   // one cooperative launch runs the whole batch (small systems)
                int it1 = batch_end;
                if (total_iters + it1 > max_iter) it1 = max_iter - total_iters;
                CoopArgs ca = {rowptr, colidx, sys, s->n_rows, x, s->r, s->rhat, s->p, s->v, s->s, s->t, s->partials, s->sc,
                               s->ictl, it, it1, rtol2};
                void* kargs[] = {&ca};
                TEFEM_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)coop_bicgstab_kernel, dim3(coop_grid(s->n_rows)),
                                                             dim3(SPMV_THREADS), kargs, 0, st));
                count_launch();
                it = it1;
            }
            for (; !coop && it < batch_end && total_iters + it < max_iter; ++it) {
                double* pdir = s->p;      // direction fed to the operator: p, or N p with a preconditioner
                if (s->pc) {
                    prof_begin(s, st, PROF_PRECOND);
                    rc = tefem_precond_apply_stream(s->pc, s->p, s->ph, st);
                    if (rc) return rc;
                    prof_end(s, st, PROF_PRECOND);
                    pdir = s->ph;
                }
                rc = dist_spmv<1>(s, rowptr, colidx, sys, pdir, s->v, s->rhat, nullptr, s->ictl, &npc, st);
                if (rc) return rc;
                prof_begin(s, st, PROF_REDUCE);
                rc = reduce_scalars<1, 1>(s, npc, it, rtol2, st);
                if (rc) return rc;
                prof_end(s, st, PROF_REDUCE);
                prof_begin(s, st, PROF_UPDATE);
                update_s_kernel<<<vgrid, 256, 0, st>>>(n, s->r, s->v, s->s, s->sc, s->ictl);
                TEFEM_LAUNCH_CHECK();
                prof_end(s, st, PROF_UPDATE);
                double* sdir = s->s;
                if (s->pc) {
                    prof_begin(s, st, PROF_PRECOND);
                    rc = tefem_precond_apply_stream(s->pc, s->s, s->sh, st);
                    if (rc) return rc;
                    prof_end(s, st, PROF_PRECOND);
                    sdir = s->sh;
                }
                rc = dist_spmv<8>(s, rowptr, colidx, sys, sdir, s->t, s->rhat, s->s, s->ictl, &npc, st);
                if (rc) return rc;
                prof_begin(s, st, PROF_REDUCE);
                rc = reduce_scalars<8, 2>(s, npc, it, rtol2, st);
                if (rc) return rc;
                prof_end(s, st, PROF_REDUCE);
                prof_begin(s, st, PROF_UPDATE);
                update_xrp_kernel<<<vgrid, 256, 0, st>>>(n, x, s->r, s->p, s->s, s->t, s->v, pdir, sdir, s->sc, s->ictl, it);
                TEFEM_LAUNCH_CHECK();
                prof_end(s, st, PROF_UPDATE);
            }
            TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_ictl, s->ictl, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
            if (s->comm && tefem_comm_err_ptr(s->comm))     // the peer time-out flag rides on the same synchronisation
                TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_ictl + 4, tefem_comm_err_ptr(s->comm), sizeof(int), cudaMemcpyDeviceToHost, st));
            TEFEM_CUDA_CHECK(cudaStreamSynchronize(st));
            if (s->comm && s->h_ictl[4]) return tefem_comm_report_timeout();
            if (s->profile) {
                const int done_now = s->h_ictl[IN_ITERS] - iters_seen;      // iterations that really ran in this batch
                rc = prof_harvest(s, 2 * (size_t)(done_now > 0 ? done_now : 0));
                if (rc) return rc;
            }
            iters_seen = s->h_ictl[IN_ITERS];
            stopped = s->h_ictl[IN_STOP] != 0;
        }
        total_iters = s->h_ictl[IN_ITERS];
        // loop back: the true residual is recomputed; converged -> exit, otherwise restart from x
        if (!stopped && total_iters >= max_iter) {
            // fall through to one last true-residual evaluation
        }
        ++restarts;
    }
    s->last_iters = (status == TEFEM_OK && restarts <= 1) ? total_iters : 0;
    if (h_info) {
        h_info[0] = (double)total_iters;
        h_info[1] = (double)restarts;
        h_info[2] = relres;
        h_info[3] = bnorm;
        h_info[4] = floor_hit ? 1.0 : 0.0;
        h_info[5] = (double)stalls;
    }
    if (status == TEFEM_ERR_NO_CONVERGE)
        set_error("BiCGSTAB did not converge: %d iterations, %d restarts, relative residual %.3e (rtol %.3e)%s",
                  total_iters, restarts, relres, rtol,
                  stalls >= 2 ? "; the true residual stalled (eps * cond floor of working precision)" : "");
    return status;
}

// Preconditioned CG (coop_cg_kernel): same host protocol as tefem_bicgstab - (re)start with the true residual, batches
// of check_every iterations, true-residual verification, stall detection.
extern "C" int tefem_cg(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys, const double* dinv,
                        const double* b, double* x, double rtol, int max_iter, int check_every, double floor_factor,
                        double* h_info, void* stream) {
    TEFEM_REQUIRE(s && rowptr && colidx && sys && dinv && b && x, "null pointer");
    TEFEM_REQUIRE((((uintptr_t)sys) & 127) == 0 && (((uintptr_t)x) & 31) == 0 && (((uintptr_t)dinv) & 31) == 0, "sys / x / dinv alignment");
    TEFEM_REQUIRE(rtol > 0 && max_iter > 0 && floor_factor >= 1.0, "rtol / max_iter / floor_factor");
    TEFEM_REQUIRE(!s->comm, "tefem_cg is a single-GPU solver");
    if (check_every < 1) check_every = 1;
    cudaStream_t st = as_stream(stream);
    const int64_t n = s->n_own;
    static int resident = -1;
    if (resident < 0) {
        int dev = 0, sms = 0, per_sm = 0, coop_ok = 0;
        TEFEM_CUDA_CHECK(cudaGetDevice(&dev));
        TEFEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TEFEM_CUDA_CHECK(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev));
        TEFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_cg_kernel, SPMV_THREADS, 0));
        TEFEM_REQUIRE(coop_ok && per_sm > 0, "cooperative launch unavailable");
        resident = sms * per_sm;
        if (resident > MAX_PARTIAL_CTAS) resident = MAX_PARTIAL_CTAS;
    }
    int64_t g64 = ceil_div(n, SPMV_THREADS);
    if (g64 > resident) g64 = resident;
    const dim3 grid((unsigned)(g64 < 1 ? 1 : g64)), block(SPMV_THREADS);
    const double rtol2 = rtol * rtol;
    const int max_restarts = 20;
    int total_iters = 0, restarts = 0, stalls = 0;
    bool floor_hit = false;
    double relres = -1.0, bnorm = 0.0, prev_relres = INFINITY;
    int status = TEFEM_ERR_NO_CONVERGE;
    TEFEM_CUDA_CHECK(cudaMemsetAsync(s->ictl, 0, 16 * sizeof(int), st));
    TEFEM_CUDA_CHECK(cudaMemsetAsync(s->sc, 0, NSCAL * sizeof(double), st));
    CgArgs ca = {rowptr, colidx, sys, dinv, s->n_rows, b, x, s->r, s->rhat, s->p, s->v, s->partials, s->sc, s->ictl, 0, 0, 0, rtol2};
    for (;;) {
        ca.mode = 0;
        void* kargs[] = {&ca};
        TEFEM_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)coop_cg_kernel, grid, block, kargs, 0, st));
        count_launch();
        TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_ictl, s->ictl, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_sc, s->sc, NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
        TEFEM_CUDA_CHECK(cudaStreamSynchronize(st));
        bnorm = sqrt(s->h_sc[SC_BNORM2]);
        relres = sqrt(s->h_sc[CG_RR]);
        if (bnorm == 0.0) {
            TEFEM_CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
            relres = 0.0; status = TEFEM_OK; break;
        }
        if (!isfinite(relres)) { set_error("non-finite residual in CG"); status = TEFEM_ERR_BREAKDOWN; break; }
        if (s->h_ictl[IN_STOP] && s->h_ictl[IN_REASON] == 1) { status = TEFEM_OK; break; }
        if (restarts > 0 && relres > 0.5 * prev_relres) ++stalls; else stalls = 0;
        prev_relres = relres;
        if (stalls >= 2) {
            if (relres <= floor_factor * rtol) { status = TEFEM_OK; floor_hit = relres > rtol; }
            break;
        }
        if (total_iters >= max_iter || restarts > max_restarts) break;
        bool stopped = false;
        int it = 0;
        while (!stopped && total_iters + it < max_iter) {
            int it1 = it + check_every;
            if (total_iters + it1 > max_iter) it1 = max_iter - total_iters;
            ca.mode = 1; ca.it0 = it; ca.it1 = it1;
            void* ka[] = {&ca};
            TEFEM_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)coop_cg_kernel, grid, block, ka, 0, st));
            count_launch();
            it = it1;
            TEFEM_CUDA_CHECK(cudaMemcpyAsync(s->h_ictl, s->ictl, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
            TEFEM_CUDA_CHECK(cudaStreamSynchronize(st));
            stopped = s->h_ictl[IN_STOP] != 0;
        }
        total_iters = s->h_ictl[IN_ITERS];
        ++restarts;
    }
    if (h_info) {
        h_info[0] = (double)total_iters;
        h_info[1] = (double)restarts;
        h_info[2] = relres;
        h_info[3] = bnorm;
        h_info[4] = floor_hit ? 1.0 : 0.0;
        h_info[5] = (double)stalls;
    }
    if (status == TEFEM_ERR_NO_CONVERGE)
        set_error("CG did not converge: %d iterations, %d restarts, relative residual %.3e (rtol %.3e)", total_iters, restarts,
                  relres, rtol);
    return status;
}

extern "C" int tefem_newmark_predict(int64_t n, const double* x, const double* v, const double* a, double dt,
                                     double gamma, double beta, const uint8_t* dir_mask, double* w1, double* w2,
                                     void* stream) {
    TEFEM_REQUIRE(x && v && a && w1 && w2, "null pointer");
    if (n == 0) return TEFEM_OK;
    // solvers.py:168-170: v + (1-gamma) dt a ; x + dt v + 0.5 dt^2 (1-2 beta) a
    newmark_predict_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        n, x, v, a, (1.0 - gamma) * dt, dt, 0.5 * (dt * dt) * (1.0 - 2.0 * beta), dir_mask, w1, w2);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_newmark_correct(int64_t n, double* x, double* v, double* a, const double* a_new, double dt,
                                     double gamma, double beta, const uint8_t* dir_mask, void* stream) {
    TEFEM_REQUIRE(x && v && a && a_new, "null pointer");
    if (n == 0) return TEFEM_OK;
    newmark_correct_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, x, v, a, a_new, dt, gamma, beta,
                                                                                       dir_mask);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_load_axpy(int64_t n_entries, const int32_t* node, const double* w, const double* h_coef, double* F,
                               void* stream) {
    TEFEM_REQUIRE(node && w && h_coef && F, "null pointer");
    if (n_entries == 0) return TEFEM_OK;
    load_axpy_kernel<<<(unsigned)ceil_div(n_entries, 256), 256, 0, as_stream(stream)>>>(n_entries, node, w, h_coef[0],
                                                                                         h_coef[1], h_coef[2], h_coef[3], F);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_load_axpy_dev(int64_t n_entries, const int32_t* node, const double* w, const double* d_coef, double* F,
                                   void* stream) {
    TEFEM_REQUIRE(node && w && d_coef && F, "null pointer");
    if (n_entries == 0) return TEFEM_OK;
    load_axpy_dev_kernel<<<(unsigned)ceil_div(n_entries, 256), 256, 0, as_stream(stream)>>>(n_entries, node, w, d_coef, F);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_lincomb3(int64_t n, double c0, const double* x0, double c1, const double* x1, double c2,
                              const double* x2, double* out, void* stream) {
    TEFEM_REQUIRE(x0 && out, "null pointer");
    if (n == 0) return TEFEM_OK;
    lincomb3_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, c0, x0, c1, x1, c2, x2, out);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}
