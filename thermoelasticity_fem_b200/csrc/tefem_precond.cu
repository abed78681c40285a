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
// Every step is a fixed linear operator (fixed sweeps, no inner Krylov), so BiCGSTAB stays valid.  On several GPUs the
// coarse cycle runs either on levels replicated on every rank (separate launches; the restricted residual, 8 * n_agg
// doubles, is summed over the ranks through NVLink peer memory inside the kernel that starts the V-cycle) or - the
// default - with the rows of levels 1 and 2 partitioned over the ranks in one persistent dataflow kernel
// (coarse_cycle_kernel below).
#include "tefem_common.cuh"
#include "tefem_comm.h"
#include "tefem_precond.h"

#include <cuda_fp16.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace tefem {

// rc[8I + 0..3] = sum_i r_i ; rc[8I + 4..6] = sum_i d_i x r_u,i ; rc[8I + 7] = 0 over the nodes i of aggregate I
// (d_i = offset from the aggregate's centre, single precision).  One CTA of 128 threads per aggregate (a cube of 8 h has
// ~500 nodes: four trips of dependent loads instead of the sixteen of a single warp, which made this pass latency bound:
// 24 us for 0.9 M nodes on one of eight GPUs); thread-strided partial sums, fixed shuffle tree, the four warp sums added
// in warp order.  An aggregate without local nodes (most of them in a multi-GPU solve) writes zeros and leaves.
constexpr int RESTRICT_THREADS = 128;
// agg_list (optional): the aggregates with local nodes - one CTA each instead of one per aggregate of the whole mesh (on one
// of eight GPUs 7 of 8 CTAs found an empty aggregate: ~7 us of launch waves).  Only together with agg_ranks: nobody reads
// the partial sums of a rank without nodes in the aggregate.
// Persistent coarse cycle on several GPUs (push.world > 1): the partial sums go straight to the rank that OWNS the aggregate
// in the cycle (aggregate i -> rank floor(((i + 1) world - 1) / n_agg)), as tagged values (tefem "LL" layout, below) in slot
// [this rank] of the owner's partial-sum area - a remote store over NVLink.  The owner's cycle kernel waits for the tags: no
// flag exchange and no remote read opens the cycle.
struct RestrictPush {
    int world = 0, rank = 0;
    uint4* base[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // partial-sum areas
    int64_t slot_units = 0;                      // 16-byte units per source-rank slot
    uint32_t tag = 0;
};
__device__ __forceinline__ void ll_store1_early(uint4* p, uint32_t tag, double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = (unsigned long long)(uint32_t)b | ((unsigned long long)tag << 32);
    const unsigned long long hi = (unsigned long long)(uint32_t)(b >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(lo), "l"(hi) : "memory");
}
__global__ void __launch_bounds__(RESTRICT_THREADS) restrict_kernel(int32_t n_agg, int32_t n_agg_total, const int32_t* __restrict__ agg_list,
                                                                    const int32_t* __restrict__ agg_ptr,
                                                                    const int32_t* __restrict__ agg_nodes,
                                                                    const float4* __restrict__ node_off,
                                                                    const double* __restrict__ r, double* __restrict__ rc,
                                                                    const RestrictPush push) {
    __shared__ double part[RESTRICT_THREADS / 32][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x >= n_agg) return;
    const int agg = agg_list ? agg_list[blockIdx.x] : (int)blockIdx.x;
    const int32_t s = agg_ptr[agg], t = agg_ptr[agg + 1];
    uint4* dst = nullptr;
    if (push.world > 1) {
        const int owner = (int)((((int64_t)agg + 1) * push.world - 1) / n_agg_total);
        dst = push.base[owner] + (int64_t)push.rank * push.slot_units + 8 * (int64_t)agg;
    }
    double* o = rc + 8 * (int64_t)agg;
    if (s == t) {
        if (threadIdx.x < 8) {
            if (dst) ll_store1_early(dst + threadIdx.x, push.tag, 0.0);
            else o[threadIdx.x] = 0.0;
        }
        return;
    }
    double a[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int32_t k = s + (int32_t)threadIdx.x; k < t; k += RESTRICT_THREADS) {
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
#pragma unroll
        for (int k = 0; k < 7; ++k) part[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        if (threadIdx.x < 7) {
            v = part[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < RESTRICT_THREADS / 32; ++w) v += part[w][threadIdx.x];
        }
        if (dst) ll_store1_early(dst + threadIdx.x, push.tag, v);
        else o[threadIdx.x] = v;
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

// ---- ROW-PARTITIONED coarse cycle: one persistent DATAFLOW kernel per application, vectors exchanged over peer memory ----
//
// The replicated cycle above is ~10 dependent launches of ~10 - 14 us each on EVERY rank, whatever the number of GPUs: at
// 8 GPUs it was 29 % of a BiCGSTAB iteration and the first limit of strong scaling (profiles/README.md).  Here the rows of
// levels 1 and 2 are PARTITIONED: rank q owns the aggregates [n1 q / world, n1 (q + 1) / world) (both pseudo nodes) and
// the level-2 aggregates [n2 q / world, ..): their rows of P2^T and of the dense level-2 inverse.  Every coarse vector of
// the cycle has its own buffer (one per phase: nothing is overwritten inside an application) with one REPLICA per rank in
// the peer-mapped region, in a SELF-VALIDATING layout: a double is stored as the 16 bytes {low word, tag, high word, tag}
// (two 8-byte units, each written atomically), where tag = 32 * application + vector number.  The rank that owns a row
// computes it and stores it (strong 256-bit stores, remote ones over NVLink) into the replicas of the ranks that read it
// (destination masks, tefem_precond_set_cycle_masks); a consumer loads the 16 bytes and simply retries until both tags
// match.  There is NO barrier, NO fence and NO flag: the restriction kernel pushes its partial sums as tagged values to the
// owner of the aggregate, and a phase starts on a row as soon as the values that row reads have arrived, so the latency of
// an application is its dependency chain (per level-1 product one NVLink store-to-visible latency plus the row's L2
// reads), not ten synchronisations.  The tagged vectors of consecutive applications alternate between two areas: a rank
// runs ahead of another by less than one application (the dense level-2 product needs P2^T t of EVERY rank), so an area
// is rewritten only after its readers have left the application that used it.
// Versions that were measured and dropped (profiles/README.md): flag barriers over all CTAs of all ranks between the
// phases (system fence after remote stores + ticket + flag + poll: ~12 us per phase, slower than the launches); weak
// stores for the tagged values (they waited microseconds in a write-combining stage); a start-flag exchange; a
// __nanosleep back-off in the retry loop; batched loads in the level-1 rows.
// A value has exactly one producer, so the replicas are bitwise identical and so are the stop decisions of the ranks.
// Waits are bounded like every peer wait.  All CTAs must be co-resident (consumers spin on producers of the same grid):
// grid <= the occupancy limit of the device.  world == 1 runs the same kernel on local buffers.
constexpr int CYC_THREADS = 256;
constexpr int CYC_GATHER_CTAS = 4;       // CTAs that copy r2 out of its tagged replica for everybody else (no polling storm)
constexpr int CYC_MAX_X = 24;            // vector numbers: level-1 iterates 0 .. CYC_MAX_X - 1, then t, x2, bc, r2
enum { CYC_V_T = 24, CYC_V_X2 = 25, CYC_V_BC = 26, CYC_V_R2 = 27, CYC_V_RC = 28 };

struct CycleArgs {
    const int32_t *rowptr1, *colidx1;
    const void* sys1;                    // float or __half blocks (template parameter)
    const double* dinv1;
    const uint8_t* agg_ranks;
    const int32_t *r2_rowptr, *r2_colidx, *p2_rowptr, *p2_colidx;
    const float *r2_vals, *p2_vals, *inv2;
    int32_t n1, n2, a0, a1;              // own aggregates [a0, a1)
    double wc;
    int n_pre, n_post, overlap;
    int world, rank;
    const double* rc[PEER_MAX_RANKS];    // world == 1: rc[0] = the restricted residual (plain doubles)
    int64_t part_off;                    // world > 1: offset (units, from X[r]) of the tagged partial sums, slot q = rank q's
    uint4* X[PEER_MAX_RANKS];            // tagged vectors of rank r: iterate k at k * vec_units, t at m, x2 at m + 1, r2 after x2 (m = n_iterates)
    uint4* bc;                           // local tagged vector: D1^-1 rc on the own rows
    const uint8_t *x_mask, *t_mask;      // per aggregate: the ranks that read its level-1 iterates / its residual t (null: all)
    int32_t j0, j1;                      // own level-2 aggregates: rows of P2^T and of the dense inverse
    double* xc_plain;                    // local: the final iterate as plain doubles (aggregates with local fine nodes)
    double* r2_plain;                    // local: r2 gathered from its tagged replica by the first CYC_GATHER_CTAS CTAs
    unsigned long long* gathered;        // local counter of finished gather CTAs (never reset)
    unsigned long long gathered_target;  // its value when this application's r2 is complete
    int64_t a2_units;                    // 16-byte units of a level-2 vector
    int64_t vec_units;                   // 16-byte units between consecutive vectors
    int n_iterates;
    uint32_t tag_base;                   // 32 * application
    int* err;
    long long timeout;
    unsigned long long* stamps;          // diagnostics (TEFEM_PROFILE_PRECOND=1): globaltimer of CTA 0 after each phase, or null
    unsigned long long* all_stamps;      // diagnostics: [n_cta][16] stamps of every CTA, or null
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void cyc_stamp(const CycleArgs& a, int k) {
    if (a.stamps && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (blockIdx.x == 0) a.stamps[k] = t;
        if (a.all_stamps) a.all_stamps[16 * (size_t)blockIdx.x + k] = t;      // diagnostics: every CTA's time line
    }
}

// L2 prefetch of [p, p + bytes) by the whole grid (128-byte lines): the operators of this rank's rows were evicted by the
// fine-level SpMV since the last application; fetched at kernel start they are L2 hits when their phase comes.
__device__ __forceinline__ void cyc_prefetch(const void* p, int64_t bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    for (int64_t o = 128 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); o < bytes; o += 128 * (int64_t)gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(c + o));
}

// ---- tagged ("LL") doubles: {lo, tag, hi, tag} ----
__device__ __forceinline__ unsigned long long ll_pack(uint32_t w, uint32_t tag) { return (unsigned long long)w | ((unsigned long long)tag << 32); }
__device__ __forceinline__ void ll_store1(uint4* p, uint32_t tag, double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(ll_pack((uint32_t)b, tag)), "l"(ll_pack((uint32_t)(b >> 32), tag)) : "memory");
}
__device__ __forceinline__ void ll_store4(uint4* p, uint32_t tag, double v0, double v1, double v2, double v3) {
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long b2 = (unsigned long long)__double_as_longlong(v2), b3 = (unsigned long long)__double_as_longlong(v3);
    // volatile (strong) stores: a weak store to peer memory may wait in a write-combining stage for microseconds
    asm volatile("st.volatile.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(ll_pack((uint32_t)b0, tag)), "l"(ll_pack((uint32_t)(b0 >> 32), tag)),
                 "l"(ll_pack((uint32_t)b1, tag)), "l"(ll_pack((uint32_t)(b1 >> 32), tag)) : "memory");
    asm volatile("st.volatile.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p + 2), "l"(ll_pack((uint32_t)b2, tag)), "l"(ll_pack((uint32_t)(b2 >> 32), tag)),
                 "l"(ll_pack((uint32_t)b3, tag)), "l"(ll_pack((uint32_t)(b3 >> 32), tag)) : "memory");
}
__device__ __forceinline__ uint4 ll_ld(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// two consecutive values with ONE 256-bit load (p 32-byte aligned)
__device__ __forceinline__ void ll_ld2(const uint4* p, uint4& v0, uint4& v1) {
    unsigned long long a, b, c, d;
    asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
    v0 = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    v1 = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)d, (uint32_t)(d >> 32));
}
__device__ __forceinline__ double ll_value(const uint4& v) { return __hiloint2double((int)v.z, (int)v.x); }
template <int N>
__device__ __forceinline__ bool ll_try(const uint4* p, uint32_t tag, uint4 (&v)[N]) {
    bool ok = true;
    if (N == 1) {
        v[0] = ll_ld(p);
    } else {
#pragma unroll
        for (int k = 0; k + 1 < N; k += 2) ll_ld2(p + k, v[k], v[k + 1]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) ok = ok && v[k].y == tag && v[k].w == tag;
    return ok;
}
// Spin until the tags of N (1 or even; p 32-byte aligned for N > 1) consecutive values match.  Bounded: a time-out marks
// the communicator failed and later waits give up at once.  (A __nanosleep(40) between retries cost ~1 us per wait:
// measured slower, profiles/README.md.)
template <int N>
__device__ __forceinline__ void ll_load(const uint4* p, uint32_t tag, double (&out)[N], int* err, long long timeout) {
    static_assert(N == 1 || (N & 1) == 0, "1 or an even number of values");
    uint4 v[N];
    if (!ll_try<N>(p, tag, v)) {
        const long long t0 = clock64();
        for (;;) {
            if (ll_try<N>(p, tag, v)) break;
            if (*reinterpret_cast<volatile int*>(err)) break;
            if (clock64() - t0 > timeout) { *err = 1; break; }
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = ll_value(v[k]);
}

// A tagged vector as the kernel sees it: the local replica to read, and (for outputs) the unit offset inside every rank's region.
struct CycVec {
    const uint4* rd;      // local replica
    int64_t off;          // offset (units) from CycleArgs::X[r]; < 0: local-only vector, written through `wr`
    uint4* wr;
    uint32_t tag;
};

// out = c0 * xin + c1 * (cb * b - A x) on the block rows [r0, r1) of a 4x4-block CSR operator, stored into the n_dst
// replicas of out.  One WARP per block row; lane l multiplies the blocks s + l, s + l + 32, ... - two trips (64 blocks, a
// whole level-1 row) are in flight together, so a row costs one chain row pointer -> column indices -> {blocks, x} whatever
// its length; the lanes that store first fetch xin and b (available before x); fixed shuffle tree; lane q < n_dst writes the
// row's four values into replica q.  (The sum order differs from coarse_spmv_kernel's 16-lane mapping in the last bits: the
// replicas of the ranks are identical by construction - one producer per value - not identical to the single-GPU cycle.)
constexpr int CYC_LANES = 32;

template <typename VT>
__device__ __forceinline__ void cyc_block_fma(const VT* __restrict__ sys, int32_t p, const double (&u)[4], double& a0, double& a1,
                                              double& a2, double& a3) {
    float4 q0, q1, q2, q3;
    load_block16(sys, (int64_t)p, q0, q1, q2, q3);
    a0 += (double)q0.x * u[0] + (double)q0.y * u[1] + (double)q0.z * u[2] + (double)q0.w * u[3];
    a1 += (double)q1.x * u[0] + (double)q1.y * u[1] + (double)q1.z * u[2] + (double)q1.w * u[3];
    a2 += (double)q2.x * u[0] + (double)q2.y * u[1] + (double)q2.z * u[2] + (double)q2.w * u[3];
    a3 += (double)q3.x * u[0] + (double)q3.y * u[1] + (double)q3.z * u[2] + (double)q3.w * u[3];
}

// Raw (unchecked) loads of 4 tagged values and the check / retry that follows: a batch of independent loads is ISSUED before
// the first tag is looked at, so that the batch costs one memory latency (a check is a branch on the loaded data: checking
// between the loads would serialise them).
__device__ __forceinline__ void ll_issue4(const uint4* p, uint4 (&v)[4]) { ll_ld2(p, v[0], v[1]); ll_ld2(p + 2, v[2], v[3]); }
__device__ __forceinline__ bool ll_ok4(const uint4 (&v)[4], uint32_t tag) {
    return v[0].y == tag && v[0].w == tag && v[1].y == tag && v[1].w == tag && v[2].y == tag && v[2].w == tag && v[3].y == tag && v[3].w == tag;
}
__device__ __forceinline__ void ll_finish4(const uint4* p, uint32_t tag, uint4 (&v)[4], double (&out)[4], int* err, long long timeout) {
    if (!ll_ok4(v, tag)) {
        const long long t0 = clock64();
        for (;;) {
            ll_issue4(p, v);
            if (ll_ok4(v, tag)) break;
            if (*reinterpret_cast<volatile int*>(err)) break;
            if (clock64() - t0 > timeout) { *err = 1; break; }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = ll_value(v[k]);
}

// (Measured and dropped: issuing the blocks and both x loads of a trip, and xin with b, as batches before the first tag check
// - fewer dependent latencies on paper, but 200 - 400 bytes of spills at the 80 registers that three CTAs per SM allow, and
// the stragglers of the post-smoothing sweeps got slower: 90 against 75 us per application, profiles/README.md.)
template <typename VT>
__device__ __forceinline__ void cyc_rows(const CycleArgs& a, uint4* const* reps, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ colidx, const VT* __restrict__ sys, int32_t r0, int32_t r1,
                                         const CycVec& x, const CycVec* b, const CycVec* xin, double c0, double c1, double cb,
                                         const CycVec& out, int n_dst, const uint8_t* __restrict__ dst_mask) {
    const int32_t gid = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
    const int32_t group = gid / CYC_LANES, n_groups = (int32_t)(gridDim.x * blockDim.x) / CYC_LANES;
    const int l = gid % CYC_LANES;
    for (int32_t base = r0; base < r1; base += n_groups) {          // uniform trip count: all lanes reach the shuffles
        const int32_t i = base + group;
        const bool live = i < r1;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        double o[4] = {0.0, 0.0, 0.0, 0.0}, bb[4] = {0.0, 0.0, 0.0, 0.0};
        bool writes = false;
        if (live) {
            const int32_t s = rowptr[i], t = rowptr[i + 1];
            for (int32_t p = s + l; p < t; p += 2 * CYC_LANES) {
                const int32_t p2 = p + CYC_LANES;
                const bool two = p2 < t;
                const int32_t j = colidx[p], j2 = two ? colidx[p2] : 0;
                double u[4], u2[4];
                ll_load<4>(x.rd + 4 * (int64_t)j, x.tag, u, a.err, a.timeout);
                if (two) ll_load<4>(x.rd + 4 * (int64_t)j2, x.tag, u2, a.err, a.timeout);
                cyc_block_fma<VT>(sys, p, u, a0, a1, a2, a3);
                if (two) cyc_block_fma<VT>(sys, p2, u2, a0, a1, a2, a3);
            }
            // lane q stores into replica q when rank q reads this row (dst_mask per aggregate = pair of block rows)
            writes = l < n_dst && (!dst_mask || ((dst_mask[i >> 1] >> l) & 1));
            if (writes) {
                const int64_t k = 4 * (int64_t)i;
                if (xin) {
                    ll_load<4>(xin->rd + k, xin->tag, o, a.err, a.timeout);
                    o[0] *= c0; o[1] *= c0; o[2] *= c0; o[3] *= c0;
                }
                if (b) {
                    ll_load<4>(b->rd + k, b->tag, bb, a.err, a.timeout);
                    bb[0] *= cb; bb[1] *= cb; bb[2] *= cb; bb[3] *= cb;
                }
            }
        }
#pragma unroll
        for (int off = CYC_LANES / 2; off > 0; off >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, off);
            a1 += __shfl_down_sync(0xffffffffu, a1, off);
            a2 += __shfl_down_sync(0xffffffffu, a2, off);
            a3 += __shfl_down_sync(0xffffffffu, a3, off);
        }
        const double s0 = __shfl_sync(0xffffffffu, a0, 0);
        const double s1 = __shfl_sync(0xffffffffu, a1, 0);
        const double s2 = __shfl_sync(0xffffffffu, a2, 0);
        const double s3 = __shfl_sync(0xffffffffu, a3, 0);
        if (writes) {
            const int64_t k = 4 * (int64_t)i;
            uint4* dst = out.off >= 0 ? reps[l] + out.off + k : out.wr + k;
            ll_store4(dst, out.tag, o[0] + c1 * (bb[0] - s0), o[1] + c1 * (bb[1] - s1), o[2] + c1 * (bb[2] - s2), o[3] + c1 * (bb[3] - s3));
        }
    }
}

template <typename VT>
__global__ void __launch_bounds__(CYC_THREADS, 3) coarse_cycle_kernel(const CycleArgs a) {
    extern __shared__ double s_r2[];                     // 8 n2 doubles: P2^T t, staged once per CTA for the dense product
    __shared__ uint4* reps[PEER_MAX_RANKS];
    const int tid = (int)threadIdx.x;
    const VT* sys1 = reinterpret_cast<const VT*>(a.sys1);
    const int me = a.rank;
    if (tid < PEER_MAX_RANKS) reps[tid] = a.X[tid];
    int stamp = 0;
    cyc_stamp(a, stamp++);                               // kernel entry
    // No start synchronisation: the partial sums of the restriction arrive as tagged values (phase 0 waits for them), and
    // the tagged vectors of consecutive applications live in two alternating areas (apply_persistent).
    __syncthreads();
    {   // operators of the own rows -> L2 (level 1, dense level-2 rows, transfer operators)
        const int32_t b0 = a.rowptr1[2 * a.a0], b1 = a.rowptr1[2 * a.a1];
        cyc_prefetch(reinterpret_cast<const char*>(a.sys1) + (int64_t)b0 * 16 * sizeof(VT), (int64_t)(b1 - b0) * 16 * sizeof(VT));
        cyc_prefetch(a.colidx1 + b0, (int64_t)(b1 - b0) * 4);
        cyc_prefetch(a.inv2 + (int64_t)8 * a.j0 * 8 * a.n2, (int64_t)8 * (a.j1 - a.j0) * 8 * a.n2 * 4);
        const int32_t c0 = a.r2_rowptr[2 * a.j0], c1 = a.r2_rowptr[2 * a.j1];
        cyc_prefetch(a.r2_vals + (int64_t)c0 * 16, (int64_t)(c1 - c0) * 64);
        const int32_t d0 = a.p2_rowptr[2 * a.a0], d1 = a.p2_rowptr[2 * a.a1];
        cyc_prefetch(a.p2_vals + (int64_t)d0 * 16, (int64_t)(d1 - d0) * 64);
    }
    cyc_stamp(a, stamp++);                               // operator prefetches issued
    const uint32_t T = a.tag_base + 1u;                  // tag of vector v: T + v
    const int64_t U = a.vec_units;
    const uint4* mine = a.X[me];
    CycVec vbc = {a.bc, -1, a.bc, T + CYC_V_BC};
    const int64_t off_r2 = (int64_t)(a.n_iterates + 1) * U + a.a2_units;
    CycVec vr2 = {mine + off_r2, off_r2, nullptr, T + CYC_V_R2};
    CycVec vt = {mine + (int64_t)a.n_iterates * U, (int64_t)a.n_iterates * U, nullptr, T + CYC_V_T};
    CycVec vx2 = {mine + (int64_t)(a.n_iterates + 1) * U, (int64_t)(a.n_iterates + 1) * U, nullptr, T + CYC_V_X2};
    int it = 0;                                          // current level-1 iterate
    auto iterate = [&](int k) { CycVec v = {mine + (int64_t)k * U, (int64_t)k * U, nullptr, T + (uint32_t)k}; return v; };
    // phase 0: bc = D1^-1 sum_ranks rc (rank order; only the ranks that own nodes of the aggregate), x_0 = wc bc.
    // 8 lanes per own aggregate, lane k = row k of the 8x8 block.
    {
        const int32_t gid = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
        const int32_t n_groups = (int32_t)(gridDim.x * blockDim.x) >> 3;
        const int k = gid & 7;
        for (int32_t i = a.a0 + (gid >> 3); i < a.a1; i += n_groups) {
            const unsigned mask = (a.world > 1 && a.agg_ranks) ? (unsigned)a.agg_ranks[i] : 0xffu;
            double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int r = 0; r < a.world; ++r) {
                if (!((mask >> r) & 1u)) continue;
                double v[8];
                if (a.world > 1) {      // tagged partial sums of rank r, pushed into this rank's area by its restriction
                    ll_load<8>(mine + a.part_off + (int64_t)r * U + 8 * (int64_t)i, T + CYC_V_RC, v, a.err, a.timeout);
                } else {
                    const double* src = a.rc[0] + 8 * (int64_t)i;
#pragma unroll
                    for (int q = 0; q < 4; ++q) ld_volatile_v2(src + 2 * q, v[2 * q], v[2 * q + 1]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) s[q] += v[q];
            }
            const double* D = a.dinv1 + 64 * (int64_t)i + 8 * k;
            double b = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) b += D[c] * s[c];
            ll_store1(a.bc + 8 * (int64_t)i + k, vbc.tag, b);
            const double x = a.wc * b;
            const unsigned xm = a.x_mask ? (unsigned)a.x_mask[i] : 0xffu;
            for (int r = 0; r < a.world; ++r)
                if ((xm >> r) & 1u) ll_store1(a.X[r] + 8 * (int64_t)i + k, T, x);
        }
    }
    cyc_stamp(a, stamp++);
    const int32_t r0 = 2 * a.a0, r1 = 2 * a.a1;           // own pseudo nodes (block rows) of level 1
    const bool overlap = a.overlap && a.n_pre >= 2;
    for (int k = 1; k < a.n_pre - (overlap ? 1 : 0); ++k) {                          // x' = x + wc (bc - A1^ x)
        const CycVec x = iterate(it), xn = iterate(it + 1);
        cyc_rows<VT>(a, reps, a.rowptr1, a.colidx1, sys1, r0, r1, x, &vbc, &x, 1.0, a.wc, 1.0, xn, a.world, a.x_mask);
        ++it;
        cyc_stamp(a, stamp++);
    }
    {   // t = bc - A1^ x
        const CycVec x = iterate(it);
        cyc_rows<VT>(a, reps, a.rowptr1, a.colidx1, sys1, r0, r1, x, &vbc, nullptr, 0.0, 1.0, 1.0, vt, a.world, a.t_mask);
    }
    cyc_stamp(a, stamp++);
    // r2 = P2^T t on the own level-2 aggregates (their children's t rows were sent here), stored into every replica
    cyc_rows<float>(a, reps, a.r2_rowptr, a.r2_colidx, a.r2_vals, 2 * a.j0, 2 * a.j1, vt, nullptr, nullptr, 0.0, -1.0, 1.0, vr2, a.world, nullptr);
    cyc_stamp(a, stamp++);
    {   // x2 = A2^-1 r2 on the own rows of the dense inverse: one CTA per row with all the row's loads in flight at once (the
        // first row is loaded BEFORE the wait for r2, which is staged in shared memory); thread-strided sums, fixed tree
        __shared__ double wsum[CYC_THREADS / 32];
        const int n = 8 * a.n2, n4 = n / 4;
        const int32_t q0 = 8 * a.j0, q1 = 8 * a.j1;
        const int lane = tid & 31, warp = tid >> 5;
        if ((int)blockIdx.x < CYC_GATHER_CTAS) {   // gather: tagged replica -> plain local array, then one counter tick per CTA
            // thread -> 4 consecutive values (one block row of r2), all its rows issued before the first check
            for (int c = 4 * ((int)blockIdx.x * CYC_THREADS + tid); c < n; c += 4 * CYC_GATHER_CTAS * CYC_THREADS) {
                uint4 v[4];
                double d[4];
                ll_issue4(vr2.rd + c, v);
                ll_finish4(vr2.rd + c, vr2.tag, v, d, a.err, a.timeout);
                a.r2_plain[c] = d[0]; a.r2_plain[c + 1] = d[1]; a.r2_plain[c + 2] = d[2]; a.r2_plain[c + 3] = d[3];
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(a.gathered, 1ULL);
            }
        }
        int32_t row = q0 + (int32_t)blockIdx.x;
        if (row < q1) {                                                                      // uniform over the CTA
            constexpr int PRE = 4;                                                           // float4 per thread held in registers
            float4 m[PRE];
            const float4* mrow = reinterpret_cast<const float4*>(a.inv2 + (int64_t)row * n);
#pragma unroll
            for (int u = 0; u < PRE; ++u) m[u] = (tid + u * CYC_THREADS < n4) ? __ldg(mrow + tid + u * CYC_THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
            // r2: every CTA polling all 8 n2 tagged values (hundreds of CTAs x 64 KB per retry) starved the few warps that
            // still produce them; the gather CTAs above copied it to a plain array - wait for their counter, read it once
            if (tid == 0) {
                const long long t0 = clock64();
                while (ld_acquire_gpu_u64(a.gathered) < a.gathered_target) {
                    if (*reinterpret_cast<volatile int*>(a.err)) break;
                    if (clock64() - t0 > a.timeout) { *a.err = 1; break; }
                }
            }
            __syncthreads();
            for (int c0 = tid; c0 < n; c0 += 8 * CYC_THREADS) {           // eight loads in flight per thread, then the stores
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (c0 + u * CYC_THREADS < n) ? __ldcg(a.r2_plain + c0 + u * CYC_THREADS) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u * CYC_THREADS < n) s_r2[c0 + u * CYC_THREADS] = v[u];
            }
            __syncthreads();
            for (; row < q1; row += (int32_t)gridDim.x) {
                double acc = 0.0;
#pragma unroll
                for (int u = 0; u < PRE; ++u) {
                    const int c = tid + u * CYC_THREADS;
                    if (c < n4) {
                        const double2 u0 = *reinterpret_cast<const double2*>(s_r2 + 4 * c), u1 = *reinterpret_cast<const double2*>(s_r2 + 4 * c + 2);
                        acc += (double)m[u].x * u0.x + (double)m[u].y * u0.y + (double)m[u].z * u1.x + (double)m[u].w * u1.y;
                    }
                }
                mrow = reinterpret_cast<const float4*>(a.inv2 + (int64_t)row * n);
                for (int c = tid + PRE * CYC_THREADS; c < n4; c += CYC_THREADS) {                // rows longer than PRE * 4 * 256 values
                    const float4 v = __ldg(mrow + c);
                    const double2 u0 = *reinterpret_cast<const double2*>(s_r2 + 4 * c), u1 = *reinterpret_cast<const double2*>(s_r2 + 4 * c + 2);
                    acc += (double)v.x * u0.x + (double)v.y * u0.y + (double)v.z * u1.x + (double)v.w * u1.y;
                }
                const int32_t next = row + (int32_t)gridDim.x;
                if (next < q1) {                                                             // the next row's loads fly during the reduction
                    const float4* mn = reinterpret_cast<const float4*>(a.inv2 + (int64_t)next * n);
#pragma unroll
                    for (int u = 0; u < PRE; ++u) m[u] = (tid + u * CYC_THREADS < n4) ? __ldg(mn + tid + u * CYC_THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
                if (lane == 0) wsum[warp] = acc;
                __syncthreads();
                if (tid < a.world) {
                    double t = wsum[0];
#pragma unroll
                    for (int w = 1; w < CYC_THREADS / 32; ++w) t += wsum[w];
                    ll_store1(a.X[tid] + vx2.off + row, vx2.tag, t);
                }
                __syncthreads();
            }
        }
    }
    cyc_stamp(a, stamp++);
    {   // x' = x + P2 x2 (+ wc t: the last pre-smoothing sweep of the overlapped cycle)
        const CycVec x = iterate(it), xn = iterate(it + 1);
        cyc_rows<float>(a, reps, a.p2_rowptr, a.p2_colidx, a.p2_vals, r0, r1, vx2, overlap ? &vt : nullptr, &x, 1.0, -1.0,
                        overlap ? -a.wc : 1.0, xn, a.world, a.x_mask);
        ++it;
    }
    cyc_stamp(a, stamp++);
    for (int k = 0; k < a.n_post; ++k) {
        const CycVec x = iterate(it), xn = iterate(it + 1);
        cyc_rows<VT>(a, reps, a.rowptr1, a.colidx1, sys1, r0, r1, x, &vbc, &x, 1.0, a.wc, 1.0, xn, a.world, a.x_mask);
        ++it;
        cyc_stamp(a, stamp++);
    }
    {   // the final iterate of the aggregates with local fine nodes as PLAIN doubles: the prolongation reads the 8 values of an
        // aggregate once per node (~500 times) - from L1 when they are plain, 128 bytes of L2 traffic per node when tagged
        const CycVec x = iterate(it);
        const int64_t n = 8 * (int64_t)a.n1;
        for (int64_t c = (int64_t)blockIdx.x * blockDim.x + tid; c < n; c += (int64_t)gridDim.x * blockDim.x) {
            if (a.world > 1 && a.agg_ranks && !((a.agg_ranks[c >> 3] >> me) & 1)) continue;
            double v[1];
            ll_load<1>(x.rd + c, x.tag, v, a.err, a.timeout);
            a.xc_plain[c] = v[0];
        }
    }
}

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
    // row-partitioned persistent coarse cycle (coarse_cycle_kernel; tefem_precond_set_coarse_mode)
    bool persistent = false;
    int cyc_grid = 0;                            // CTAs of the persistent kernel (<= co-resident limit)
    uint4* cyc_local = nullptr;                  // caller's zeroed buffer of tagged vectors (tefem_precond_cycle_doubles)
    const uint8_t *x_mask = nullptr, *t_mask = nullptr;      // tefem_precond_set_cycle_masks
    const int32_t* agg_list = nullptr;           // aggregates with local nodes (tefem_precond_set_agg_ranks), or null: all
    int32_t n_agg_list = 0;
    double* xc_plain = nullptr;                  // persistent cycle: the final iterate as plain doubles for the prolongation
    int* d_err_local = nullptr;                  // time-out flag of the tagged loads without a communicator
    uint32_t local_epoch = 0;
    unsigned long long gather_total = 0;         // ticks of the r2 gather counter so far
    // TEFEM_PROFILE_PRECOND=1 (diagnostics, stderr at destroy): device time of the three parts of an application -
    // restriction + all-reduce over the ranks + first sweep (includes waiting for the slowest rank), the level-1 / level-2
    // cycle (replicated work), the prolongation.  Synchronises the stream after every application: not for timed runs.
    bool profile = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double part_ms[3] = {0.0, 0.0, 0.0};
    int64_t n_apply = 0;
    double stamp_us[32] = {0.0};                 // persistent cycle: time of CTA 0 between its phase stamps
    int n_stamps = 0;
    unsigned long long* d_all_stamps = nullptr;  // diagnostics only (allocated when TEFEM_PROFILE_PRECOND=1)
    double phase_end_us[3][16] = {{0.0}};        // per phase: min / median / max over the CTAs of the time since the kernel's first stamp
};

static int64_t pc_align(int64_t n) { return (n + 15) / 16 * 16; }

extern "C" int64_t tefem_precond_work_doubles(int32_t n_agg, int32_t n_agg2) {
    return 4 * pc_align(8 * (int64_t)n_agg) + 2 * pc_align(8 * (int64_t)n_agg2) + 16;
}

// Level-1 iterates of one application of the persistent cycle (each has its own buffer): x_0, the pre-smoothing sweeps, the
// level-2 update, the post-smoothing sweeps.
static int cycle_iterates(int n_pre, int n_post, bool overlap) {
    return 2 + ((overlap && n_pre >= 2) ? n_pre - 2 : n_pre - 1) + n_post;
}
// 16-byte units of the tagged vectors every rank's region holds: the iterates (for any cycle form), t, x2
static int64_t cycle_shared_units(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post) {
    return (int64_t)(cycle_iterates(n_pre, n_post, false) + 1) * pc_align(8 * (int64_t)n_agg) + 2 * pc_align(8 * (int64_t)n_agg2);
}
// ... and, with a communicator, one slot of tagged partial restricted residuals per source rank
static int64_t cycle_peer_units(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post) {
    return cycle_shared_units(n_agg, n_agg2, n_pre, n_post) + (int64_t)PEER_MAX_RANKS * pc_align(8 * (int64_t)n_agg);
}

// Capacity (doubles per parity) of the peer-memory region the row-partitioned coarse cycle needs (tefem_comm_shm_alloc):
// the partial restricted residual, then the tagged vectors (16 bytes per value).
extern "C" int64_t tefem_precond_peer_doubles(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post) {
    return pc_align(8 * (int64_t)n_agg) + 4 * cycle_peer_units(n_agg, n_agg2, n_pre, n_post);       // two application parities
}
// Doubles of the caller's ZEROED, 128-byte aligned buffer for tefem_precond_set_coarse_mode(mode 1): the local tagged vectors
// (bc, r2) and - used without a communicator - the shared ones.
extern "C" int64_t tefem_precond_cycle_doubles(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post) {
    return 2 * (pc_align(8 * (int64_t)n_agg) + 2 * cycle_shared_units(n_agg, n_agg2, n_pre, n_post)) + pc_align(8 * (int64_t)n_agg2) + 32;
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
    p->r2 = w; w += a2; p->x2 = w; w += a2;
    p->d_err_local = reinterpret_cast<int*>(w);                      // the caller zeroed `work`
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

// mode 0: the coarse cycle as separate launches on replicated levels; 1: row-partitioned persistent dataflow kernel
// (coarse_cycle_kernel).  cycle_work: zeroed buffer of tefem_precond_cycle_doubles doubles (mode 1).  With a communicator
// the peer region must hold tefem_precond_peer_doubles per parity.
template <typename VT>
static int cycle_configure(tefem_precond* p, size_t smem, int* per_sm) {
    if (smem > 48 * 1024)
        TEFEM_CUDA_CHECK(cudaFuncSetAttribute(coarse_cycle_kernel<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TEFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, coarse_cycle_kernel<VT>, CYC_THREADS, smem));
    (void)p;
    return TEFEM_OK;
}

extern "C" int tefem_precond_set_coarse_mode(tefem_precond* p, int mode, double* cycle_work) {
    TEFEM_REQUIRE(p && (mode == 0 || mode == 1), "mode 0 (launches) or 1 (persistent)");
    if (mode == 1) {
        TEFEM_REQUIRE(cycle_work && (((uintptr_t)cycle_work) & 127) == 0, "cycle_work: zeroed, 128-byte aligned buffer");
        TEFEM_REQUIRE(cycle_iterates(p->n_pre, p->n_post, p->overlap) <= CYC_MAX_X, "too many sweeps for the persistent coarse cycle");
        const size_t smem = sizeof(double) * 8 * (size_t)p->n2;
        TEFEM_REQUIRE(smem <= 96 * 1024, "level 2 too large for the persistent coarse cycle (8 n_agg2 doubles of shared memory)");
        if (p->comm)
            TEFEM_REQUIRE(tefem_comm_has_peer(p->comm) &&
                              tefem_comm_peer_cap(p->comm) >= tefem_precond_peer_doubles(p->n1, p->n2, p->n_pre, p->n_post),
                          "the peer-memory region must hold tefem_precond_peer_doubles(...) doubles per parity");
        int dev = 0, sms = 0, per_sm = 0;
        TEFEM_CUDA_CHECK(cudaGetDevice(&dev));
        TEFEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int rc = p->sys1h ? cycle_configure<__half>(p, smem, &per_sm) : cycle_configure<float>(p, smem, &per_sm);
        if (rc) return rc;
        TEFEM_REQUIRE(per_sm >= 1, "the coarse-cycle kernel does not fit on an SM");
        const int world = p->comm ? tefem_comm_world(p->comm) : 1;
        // one pass over the own block rows of level 1 (one warp each) when the co-resident limit allows
        const int64_t own_rows = 2 * ceil_div((int64_t)p->n1, world);
        int64_t g = ceil_div(own_rows * CYC_LANES, CYC_THREADS);
        if (g < 64) g = 64;
        if (g > (int64_t)sms * per_sm) g = (int64_t)sms * per_sm;
        p->cyc_grid = (int)g;
        p->cyc_local = reinterpret_cast<uint4*>(cycle_work);
    }
    p->persistent = mode == 1;
    return TEFEM_OK;
}

// Persistent cycle on several GPUs: x_mask[i] / t_mask[i] (uint8 device arrays over the aggregates, bit q) = rank q reads
// the level-1 iterates / the residual t of aggregate i: the owner stores a row only into those replicas (without the
// masks every row goes to every rank: 7 x the NVLink traffic of a slab partition).  The owner's own bit must be set.
// Ownership: aggregate i belongs to rank q with n_agg q / world <= i < n_agg (q + 1) / world (integer division), level-2
// aggregate J likewise with n_agg2.
extern "C" int tefem_precond_set_cycle_masks(tefem_precond* p, const uint8_t* x_mask, const uint8_t* t_mask) {
    TEFEM_REQUIRE(p, "null preconditioner");
    p->x_mask = x_mask; p->t_mask = t_mask;
    return TEFEM_OK;
}

extern "C" int tefem_precond_set_agg_ranks(tefem_precond* p, const uint8_t* agg_ranks) {
    TEFEM_REQUIRE(p, "null preconditioner");
    p->agg_ranks = agg_ranks;
    return TEFEM_OK;
}

// Multi-GPU: the aggregates that have fine nodes on this rank (device array of n indices): the restriction launches one
// CTA per listed aggregate.  Requires agg_ranks (partial sums of unlisted aggregates are never written nor read).
extern "C" int tefem_precond_set_agg_list(tefem_precond* p, const int32_t* agg_list, int32_t n) {
    TEFEM_REQUIRE(p && n >= 0 && (n == 0 || agg_list), "arguments");
    TEFEM_REQUIRE(p->agg_ranks, "set agg_ranks first");
    p->agg_list = n > 0 ? agg_list : nullptr;
    p->n_agg_list = n;
    return TEFEM_OK;
}

extern "C" int tefem_precond_destroy(tefem_precond* p) {
    if (p && p->profile) {
        if (p->n_apply > 0)
            fprintf(stderr, "[tefem precond profile] %lld applications: restrict + all-reduce + first sweep %.4f ms, level-1/2 cycle %.4f ms, "
                            "prolongation %.4f ms per application\n", (long long)p->n_apply, p->part_ms[0] / p->n_apply,
                    p->part_ms[1] / p->n_apply, p->part_ms[2] / p->n_apply);
        if (p->n_apply > 0 && p->n_stamps > 0) {
            fprintf(stderr, "[tefem precond profile] persistent cycle, CTA 0, us per phase (prefetch, phase 0, sweeps.., t, r2, x2, update, post..):");
            for (int k = 1; k < p->n_stamps; ++k) fprintf(stderr, " %.2f", p->stamp_us[k] / p->n_apply);
            fprintf(stderr, "\n");
            if (p->d_all_stamps) {
                const char* nm[3] = {"first", "median", "last"};
                for (int q = 0; q < 3; ++q) {
                    fprintf(stderr, "[tefem precond profile] end of phase over all CTAs, us since the first CTA started (%s CTA):", nm[q]);
                    for (int k = 0; k < p->n_stamps; ++k) fprintf(stderr, " %.2f", p->phase_end_us[q][k] / p->n_apply);
                    fprintf(stderr, "\n");
                }
            }
        }
        if (p->d_all_stamps) cudaFree(p->d_all_stamps);
        for (int k = 0; k < 4; ++k)
            if (p->ev[k]) cudaEventDestroy(p->ev[k]);
    }
    delete p;
    return TEFEM_OK;
}

// One application with the row-partitioned persistent coarse cycle: restriction, coarse_cycle_kernel, prolongation.
static int apply_persistent(tefem_precond* p, const double* r, double* z, cudaStream_t st) {
    const int64_t a1 = pc_align(8 * (int64_t)p->n1), a2 = pc_align(8 * (int64_t)p->n2);
    const int32_t n_restrict = p->agg_list ? p->n_agg_list : p->n1;                             // restriction: one CTA per (local) aggregate
    const unsigned gr = (unsigned)(n_restrict > 0 ? n_restrict : 1);
    CycleArgs a;
    memset(&a, 0, sizeof(a));
    a.rowptr1 = p->rowptr1; a.colidx1 = p->colidx1;
    a.sys1 = p->sys1h ? (const void*)p->sys1h : (const void*)p->sys1;
    a.dinv1 = p->dinv1; a.agg_ranks = p->agg_ranks;
    a.r2_rowptr = p->r2_rowptr; a.r2_colidx = p->r2_colidx; a.r2_vals = p->r2_vals;
    a.p2_rowptr = p->p2_rowptr; a.p2_colidx = p->p2_colidx; a.p2_vals = p->p2_vals; a.inv2 = p->inv2;
    a.n1 = p->n1; a.n2 = p->n2;
    a.wc = p->omega_c; a.n_pre = p->n_pre; a.n_post = p->n_post; a.overlap = p->overlap ? 1 : 0;
    a.bc = p->cyc_local;
    a.vec_units = a1; a.a2_units = a2;
    a.part_off = cycle_shared_units(p->n1, p->n2, p->n_pre, p->n_post);
    a.x_mask = p->x_mask; a.t_mask = p->t_mask;
    a.n_iterates = cycle_iterates(p->n_pre, p->n_post, p->overlap);
    double* rc_mine = p->rc;
    a.rc[0] = p->rc;                                              // read by phase 0 when world == 1 (with or without a communicator)
    if (p->comm) {
        const PeerCtx pc = tefem_comm_peer_ctx(p->comm, PEER_CH_CYCLE);
        a.world = pc.world; a.rank = pc.rank;
        a.tag_base = pc.epoch * 32u;
        a.err = pc.err; a.timeout = pc.timeout;
        for (int q = 0; q < pc.world; ++q) {
            double* big = reinterpret_cast<double*>(pc.base[q] + PEER_BIG_OFFSET);
            // the tagged vectors of consecutive applications alternate between two areas: a rank can run ahead of another by
            // less than one application (the dense level-2 product needs P2^T t of every rank), so an area is rewritten only
            // after all its readers have left the application that used it
            a.X[q] = reinterpret_cast<uint4*>(big + a1) + (int64_t)(pc.epoch & 1u) * cycle_peer_units(p->n1, p->n2, p->n_pre, p->n_post);
        }
    } else {
        a.world = 1; a.rank = 0;
        a.tag_base = (++p->local_epoch) * 32u;
        a.err = p->d_err_local;                                   // local producers cannot be lost
        a.timeout = PEER_TIMEOUT_DEFAULT;
        a.rc[0] = p->rc;
        a.X[0] = p->cyc_local + a1 + (int64_t)(p->local_epoch & 1u) * cycle_shared_units(p->n1, p->n2, p->n_pre, p->n_post);
    }
    // tail of cycle_work: r2 as a plain vector, then 32 words: the phase stamps of CTA 0 (diagnostics) and the gather counter
    double* tail = reinterpret_cast<double*>(p->cyc_local) + tefem_precond_cycle_doubles(p->n1, p->n2, p->n_pre, p->n_post) - 32;
    unsigned long long* d_stamps = reinterpret_cast<unsigned long long*>(tail);
    a.stamps = p->profile ? d_stamps : nullptr;
    if (p->profile && !p->d_all_stamps) {
        if (cudaMalloc(reinterpret_cast<void**>(&p->d_all_stamps), sizeof(unsigned long long) * 16 * (size_t)p->cyc_grid) != cudaSuccess)
            p->d_all_stamps = nullptr;
    }
    a.all_stamps = p->profile ? p->d_all_stamps : nullptr;
    a.r2_plain = tail - a2;
    a.xc_plain = p->xc;
    a.gathered = reinterpret_cast<unsigned long long*>(tail + 16);
    p->gather_total += CYC_GATHER_CTAS;
    a.gathered_target = p->gather_total;
    a.a0 = (int32_t)((int64_t)p->n1 * a.rank / a.world); a.a1 = (int32_t)((int64_t)p->n1 * (a.rank + 1) / a.world);
    a.j0 = (int32_t)((int64_t)p->n2 * a.rank / a.world); a.j1 = (int32_t)((int64_t)p->n2 * (a.rank + 1) / a.world);
    if (p->profile) cudaEventRecord(p->ev[0], st);
    RestrictPush push;
    if (p->comm) {      // partial sums go to the owner of the aggregate as tagged values
        push.world = a.world; push.rank = a.rank; push.slot_units = a.vec_units; push.tag = a.tag_base + 1u + CYC_V_RC;
        for (int q = 0; q < a.world; ++q) push.base[q] = a.X[q] + a.part_off;
    }
    restrict_kernel<<<gr, RESTRICT_THREADS, 0, st>>>(n_restrict, p->n1, p->agg_list, p->agg_ptr, p->agg_nodes, p->node_off, r, rc_mine, push);
    TEFEM_LAUNCH_CHECK();
    if (p->profile) cudaEventRecord(p->ev[1], st);
    void* kargs[] = {&a};
    const void* fn = p->sys1h ? (const void*)coarse_cycle_kernel<__half> : (const void*)coarse_cycle_kernel<float>;
    // A plain launch (a cooperative one costs more): the grid is within the co-resident limit of the device, and a CTA that
    // waits for a producer of the same grid can only be delayed - not blocked - by independent work of other streams.
    TEFEM_CUDA_CHECK(cudaLaunchKernel(fn, dim3((unsigned)p->cyc_grid), dim3(CYC_THREADS), kargs, sizeof(double) * 8 * (size_t)p->n2, st));
    count_launch();
    TEFEM_CUDA_CHECK(cudaGetLastError());
    if (p->profile) cudaEventRecord(p->ev[2], st);
    if (p->n_rows > 0) {
        prolong_combine_kernel<<<(unsigned)ceil_div((int64_t)p->n_rows, 256), 256, 0, st>>>(p->n_rows, p->node_agg, p->node_off,
                                                                                            p->omega, r, p->xc, z);
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
            unsigned long long h[16];
            const int ns = 6 + a.n_iterates - 1 < 16 ? 6 + a.n_iterates - 1 : 16;    // entry, prefetch, phase 0, sweeps, t, r2, x2, update, post
            if (cudaMemcpy(h, d_stamps, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
                for (int k = 1; k < ns; ++k) p->stamp_us[k] += 1e-3 * (double)(h[k] - h[k - 1]);
                p->n_stamps = ns;
            }
            if (p->d_all_stamps) {
                std::vector<unsigned long long> all(16 * (size_t)p->cyc_grid);
                if (cudaMemcpy(all.data(), p->d_all_stamps, sizeof(unsigned long long) * all.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
                    unsigned long long t0 = ~0ULL;
                    for (int c = 0; c < p->cyc_grid; ++c) t0 = all[16 * (size_t)c] < t0 ? all[16 * (size_t)c] : t0;
                    std::vector<double> v((size_t)p->cyc_grid);
                    for (int k = 0; k < ns; ++k) {
                        for (int c = 0; c < p->cyc_grid; ++c) v[(size_t)c] = 1e-3 * (double)(all[16 * (size_t)c + k] - t0);
                        std::sort(v.begin(), v.end());
                        p->phase_end_us[0][k] += v.front(); p->phase_end_us[1][k] += v[v.size() / 2]; p->phase_end_us[2][k] += v.back();
                    }
                }
            }
        }
    }
    return TEFEM_OK;
}

int tefem_precond_apply_stream(tefem_precond* p, const double* r, double* z, cudaStream_t st) {
    TEFEM_REQUIRE(p && r && z, "null pointer");
    if (p->persistent) return apply_persistent(p, r, z, st);
    const int64_t n1 = 8 * (int64_t)p->n1;                    // scalar size of level 1
    const int32_t np1 = 2 * p->n1, np2 = 2 * p->n2;           // pseudo nodes (4-DOF block rows) of levels 1 and 2
    const unsigned g1 = (unsigned)ceil_div(n1, 256);                                           // one thread per scalar of level 1
    const unsigned gs1 = (unsigned)ceil_div((int64_t)COARSE_LANES * np1, 256);                 // coarse_spmv: 16 lanes per block row
    const unsigned gs2 = (unsigned)ceil_div((int64_t)COARSE_LANES * np2, 256);
    const int32_t n_restrict = p->agg_list ? p->n_agg_list : p->n1;                             // restriction: one CTA per (local) aggregate
    const unsigned gr = (unsigned)(n_restrict > 0 ? n_restrict : 1);
    if (p->profile) cudaEventRecord(p->ev[0], st);
    // level 0 -> 1: rc = P1^T r, summed over the ranks (levels 1 and 2 are replicated): the restriction writes
    // into this rank's peer buffer and the next kernel does all-reduce + D1^-1 + first smoothing step in one
    if (p->comm) {
        TEFEM_REQUIRE(tefem_comm_has_peer(p->comm) && tefem_comm_peer_cap(p->comm) >= n1,
                      "the peer-memory region must hold 8 * n_agg doubles (tefem_comm_shm_alloc)");
        const PeerCtx pc = tefem_comm_peer_ctx(p->comm, PEER_CH_COARSE);
        double* mine = reinterpret_cast<double*>(pc.base[pc.rank] + PEER_BIG_OFFSET) + (int64_t)(pc.epoch & 1u) * pc.cap;
        restrict_kernel<<<gr, RESTRICT_THREADS, 0, st>>>(n_restrict, p->n1, p->agg_list, p->agg_ptr, p->agg_nodes, p->node_off, r, mine, RestrictPush());
        TEFEM_LAUNCH_CHECK();
        peer_coarse_start_kernel<<<(unsigned)ceil_div((int64_t)p->n1, 128), 128, 0, st>>>(pc, p->n1, p->agg_ranks, p->dinv1, p->omega_c,
                                                                                          p->bc, p->xc);
        TEFEM_LAUNCH_CHECK();
    } else {
        restrict_kernel<<<gr, RESTRICT_THREADS, 0, st>>>(n_restrict, p->n1, p->agg_list, p->agg_ptr, p->agg_nodes, p->node_off, r, p->rc, RestrictPush());
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
