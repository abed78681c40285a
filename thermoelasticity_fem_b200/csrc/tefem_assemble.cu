// K1: element quadrature + atomics-free owner-computes scatter into the node-block pattern.
//
// Replaces the reference's assembly loops (model.py:87-118): per element the six
// Element.compute_mat_*_e (elements.py:220-286), the np.ix_ scatter into a dense n_dofs^2 array
// and csc_array(dense).  Design (DESIGN.md "K1"):
//
//   * the symbolic phase groups node rows into TILES and stores everything a tile needs in per-tile
//     segments (16-byte aligned): node coordinates, byte-packed tile-local connectivity, incidence
//     entries (the slot map), work items, split lists;
//   * the kernel is PERSISTENT (grid = CTAs that fit on the 148 SMs, 3 per SM): while a CTA computes
//     tile t it prefetches the segments of its next tile into the other half of a double buffer with
//     TMA bulk copies (cp.async.bulk -> UBLKCP) that complete on an mbarrier, so no global-memory
//     latency sits on the critical path of a tile (v1-v3 were latency bound, profiles/README.md);
//   * per tile: phase 1 builds the GEOMETRY record (inverse Jacobian columns = shape-function
//     gradients, det; no material) of every staged element in shared memory; phase 2: one thread per
//     WORK ITEM - a node-pair block, or a piece of a block whose incidence list is long (the diagonal
//     blocks) or crosses a material interface - walks its incidence entries (element, a, b) in
//     ascending element order and accumulates the MATERIAL-FREE geometric sums (9 + 3 + 1 + 3 values)
//     in registers; all elements of an item share one material, whose constants are applied once per
//     item (7 / 10 shared-memory loads per incidence instead of 11 / 17: the kernel is bound by the
//     shared-memory pipe).  Items are in block order, so the structure-of-arrays plane stores of a
//     warp are contiguous; pieces of split blocks park their values in shared memory and one thread
//     per split block adds them in list order (phase 3).  Every output value is written exactly once:
//     no atomics, bitwise reproducible, summation in the reference's element order (piecewise for the
//     split blocks).
//
// The same per-thread routines are compiled for the host when TEFEM_HOST_EMU is defined; that
// build is used only by tests/host_emulation to check the schedule logic in a GPU-less container
// and is never loaded by the package.
#include "tefem_element.cuh"

#include <limits.h>
#include <stdlib.h>
#include <vector>

namespace tefem {

struct AssembleArgs {
    const double* coords;      // element_matrices path only
    const int32_t* conn;       // element_matrices path only
    const int32_t* mat_id;     // element_matrices path only
    const double* mat_table;
    const int32_t* cta_hdr;    // [n_cta, 12]: n_items, n_elems, n_inc, n_split, n_nodes, o_items, o_elems, o_inc, o_split, o_nodes, block0
    const int32_t* cta_elems;  // staged element numbers (error reporting only)
    const int32_t* cta_conn8;  // connectivity of the staged elements as 4 tile-local node bytes
    const double* cta_xyz;     // coordinates of the tile's nodes, [.., 3]
    const int32_t* item_inc;   // first incidence entry of the work item (relative to the tile's segment) | tile-local block << 16
    const int32_t* item_meta;  // length | (partial slot + 1) << 8 | material << 24   (slot + 1 == 0: the block's only item)
    const int32_t* split_block;  // tile-local block with several items
    const int32_t* split_meta;   // first partial slot | n_items << 16
    const uint16_t* inc;
    int32_t n_cta;
    int64_t n_blocks;
    double alpha_M, alpha_K;
    double* vals_M;
    double* vals_K;
    double* vals_D;
    int32_t* bad_elem;
};

enum { H_NITEMS = 0, H_NELEMS, H_NINC, H_NSPLIT, H_NNODES, H_OITEMS, H_OELEMS, H_OINC, H_OSPLIT, H_ONODES, H_BLOCK0, H_NPIECE, HDR_STRIDE = 12 };

// ---- per-thread routines (host + device) -------------------------------------------------------

// Geometry record of element e read through the global connectivity (element_matrices path).
TEFEM_HD double stage_element(const AssembleArgs& A, int32_t e, double* rec) {
    const int32_t n0 = A.conn[4 * (int64_t)e], n1 = A.conn[4 * (int64_t)e + 1];
    const int32_t n2 = A.conn[4 * (int64_t)e + 2], n3 = A.conn[4 * (int64_t)e + 3];
    return element_record<true>(A.coords + 3 * (int64_t)n0, A.coords + 3 * (int64_t)n1, A.coords + 3 * (int64_t)n2,
                                A.coords + 3 * (int64_t)n3, rec);
}

// Phase 1 from the tile's staged node coordinates xyz[tile-local node][3] and the byte-packed tile-local
// connectivity (no gather from global memory per element).
template <bool WITH_GT>
TEFEM_HD double stage_element_local(const double* xyz, uint32_t conn8, double* rec) {
    return element_record<WITH_GT>(xyz + 3 * (conn8 & 0xff), xyz + 3 * ((conn8 >> 8) & 0xff), xyz + 3 * ((conn8 >> 16) & 0xff),
                                   xyz + 3 * (conn8 >> 24), rec);
}

template <int OPS, int DUU>
struct AsmTraits {
    static constexpr bool WANT_M = (OPS & TEFEM_OP_M) != 0, WANT_K = (OPS & TEFEM_OP_K) != 0, WANT_D = (OPS & TEFEM_OP_D) != 0;
    static constexpr bool NEED_K = WANT_K || (WANT_D && DUU == 2);
    static constexpr bool NEED_M = WANT_M || (WANT_D && DUU >= 1);
    typedef BlockAcc<NEED_K, NEED_M, WANT_D> Acc;
    typedef GeomAcc<NEED_K, NEED_M, WANT_D> Geom;
    static constexpr int REC = WANT_D ? REC_STRIDE_D : REC_STRIDE_K;     // record stride (doubles)
    static constexpr int SLOT = Acc::LIVE | 1;                           // slot stride (doubles), odd
};

// item_meta fields
TEFEM_HD int meta_len(int32_t meta) { return meta & 0xff; }
TEFEM_HD int meta_slot(int32_t meta) { return ((meta >> 8) & 0xffff) - 1; }
TEFEM_HD int meta_mat(int32_t meta) { return (int)((uint32_t)meta >> 24); }

// Write the planes of block p.  DUU: 0 = D has no uu part, 1 = alpha_M only (3 diagonal planes),
// 2 = alpha_K != 0 (9 planes); model.py:112-117.
template <int OPS, int DUU>
TEFEM_HD void store_block(const AssembleArgs& A, int32_t p, const typename AsmTraits<OPS, DUU>::Acc& acc) {
    typedef AsmTraits<OPS, DUU> T;
    const int64_t P = A.n_blocks;
    if (T::WANT_M) {
        A.vals_M[p] = acc.m;
        A.vals_M[P + p] = acc.m;
        A.vals_M[2 * P + p] = acc.m;
    }
    if (T::WANT_K) {
#pragma unroll
        for (int i = 0; i < 9; ++i) A.vals_K[i * P + p] = acc.kuu[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) A.vals_K[(9 + i) * P + p] = acc.kut[i];
        A.vals_K[12 * P + p] = acc.ktt;
    }
    if (T::WANT_D) {
#pragma unroll
        for (int i = 0; i < 3; ++i) A.vals_D[i * P + p] = acc.dtu[i];
        A.vals_D[3 * P + p] = acc.dtt;
        if (DUU == 1) {
            const double m = A.alpha_M * acc.m;
            A.vals_D[4 * P + p] = m;
            A.vals_D[5 * P + p] = m;
            A.vals_D[6 * P + p] = m;
        }
        if (DUU == 2) {
            const double m = A.alpha_M * acc.m;   // alpha_M may be 0 here: adds an exact 0
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    A.vals_D[(4 + 3 * i + j) * P + p] = (i == j ? m : 0.0) + A.alpha_K * acc.kuu[3 * i + j];
        }
    }
}

// Geometric sums of one material-uniform list inc[s .. s + len) (ascending element order) with the material
// constants applied once.
template <int OPS, int DUU>
TEFEM_HD void gather_item(const AssembleArgs& A, uint32_t inc_word, int32_t meta, const double* recs, const uint16_t* inc_s,
                          double W4, const QuadS& S, const QuadNN& NN, typename AsmTraits<OPS, DUU>::Acc& acc) {
    typedef AsmTraits<OPS, DUU> T;
    typename T::Geom g;
    g.clear();
    const uint32_t s = inc_word & 0xffffu, t = s + (uint32_t)meta_len(meta);
#pragma unroll 2
    for (uint32_t k = s; k < t; ++k) {
        const uint32_t w = inc_s[k];
        block_accumulate<T::NEED_K, T::NEED_M, T::WANT_D>(g, recs + (size_t)(w >> 4) * T::REC, (w >> 2) & 3, w & 3, W4, S, NN);
    }
    block_finalize<T::NEED_K, T::NEED_M, T::WANT_D>(g, A.mat_table + MAT_STRIDE * meta_mat(meta), acc);
}

// One work item of the inline layout: either write the block (items are in block order: consecutive threads write
// consecutive addresses of every plane) or park the values of a piece of a split block in its shared-memory slot.
template <int OPS, int DUU>
TEFEM_HD void process_item(const AssembleArgs& A, int32_t block0, uint32_t inc_word, int32_t meta, const double* recs,
                           const uint16_t* inc_s, double* slots, double W4, const QuadS& S, const QuadNN& NN) {
    typedef AsmTraits<OPS, DUU> T;
    typename T::Acc acc;
    gather_item<OPS, DUU>(A, inc_word, meta, recs, inc_s, W4, S, NN, acc);
    const int slot = meta_slot(meta);
    if (slot < 0) store_block<OPS, DUU>(A, block0 + (int32_t)(inc_word >> 16), acc);
    else acc.dump(slots + (size_t)slot * T::SLOT);
}

// Pieces-first layout.  First pass (p < 0): a piece of a split block sums its list and parks the values in its slot.
// Second pass (p = the block, one thread per block in block order): a plain block sums its whole list; a combine item
// (slot field set) adds its meta_len(meta) parked pieces in list order.  Either way the block is stored by the thread
// at its block-order position: the plane stores of a warp are contiguous.
template <int OPS, int DUU>
TEFEM_HD void process_pf_item(const AssembleArgs& A, int32_t p, uint32_t inc_word, int32_t meta, const double* recs,
                              const uint16_t* inc_s, double* slots, double W4, const QuadS& S, const QuadNN& NN) {
    typedef AsmTraits<OPS, DUU> T;
    typename T::Acc acc;
    const int slot = meta_slot(meta);
    if (p < 0 || slot < 0) {
        gather_item<OPS, DUU>(A, inc_word, meta, recs, inc_s, W4, S, NN, acc);
    } else {
        acc.clear();
        const int n = meta_len(meta);
        for (int c = 0; c < n; ++c) acc.add_from(slots + (size_t)(slot + c) * T::SLOT);
    }
    if (p < 0) acc.dump(slots + (size_t)slot * T::SLOT);
    else store_block<OPS, DUU>(A, p, acc);
}

// Phase 3 for one split block: add the values of its items in list order (= ascending element) and write the block.
template <int OPS, int DUU>
TEFEM_HD void combine_split(const AssembleArgs& A, int32_t p, int32_t meta, const double* slots) {
    typedef AsmTraits<OPS, DUU> T;
    typename T::Acc acc;
    acc.clear();
    const int slot0 = meta & 0xffff, n = (int)((uint32_t)meta >> 16);
    for (int c = 0; c < n; ++c) acc.add_from(slots + (size_t)(slot0 + c) * T::SLOT);
    store_block<OPS, DUU>(A, p, acc);
}

// Dense element matrices of one element (parity path; elements.py:220-286).
TEFEM_HD void element_dense(const double* rec, const double* mat, double W4, const QuadS& S, const QuadNN& NN, int kinds,
                            double* Muu, double* Dtu, double* Dtt, double* Kuu, double* Kut, double* Ktt) {
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            GeomAcc<true, true, true> g;
            g.clear();
            block_accumulate<true, true, true>(g, rec, a, b, W4, S, NN);
            BlockAcc<true, true, true> acc;
            block_finalize<true, true, true>(g, mat, acc);
            if (kinds & TEFEM_KIND_MUU)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Muu[(3 * a + i) * 12 + 3 * b + j] = (i == j) ? acc.m : 0.0;
            if (kinds & TEFEM_KIND_KUU)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Kuu[(3 * a + i) * 12 + 3 * b + j] = acc.kuu[3 * i + j];
            if (kinds & TEFEM_KIND_KUT)
                for (int i = 0; i < 3; ++i) Kut[(3 * a + i) * 4 + b] = acc.kut[i];
            if (kinds & TEFEM_KIND_KTT) Ktt[a * 4 + b] = acc.ktt;
            if (kinds & TEFEM_KIND_DTU)
                for (int c = 0; c < 3; ++c) Dtu[a * 12 + 3 * b + c] = acc.dtu[c];
            if (kinds & TEFEM_KIND_DTT) Dtt[a * 4 + b] = acc.dtt;
        }
}

// ---- host side: quadrature tables ----------------------------------------------------------------

static QuadTables make_quad_tables() {
    // elements.py:65-69: truncated 10-digit literals; :76-87: N_a = c0 + c1 x + c2 y + c3 z evaluated
    // left to right with the coefficient table of :56-61 ([1,-1,-1,-1], [0,1,0,0], ...).
    const double w = 0.0416666667, ga = 0.5854101966, gb = 0.1381966011;
    const double pts[4][3] = {{ga, gb, gb}, {gb, ga, gb}, {gb, gb, ga}, {gb, gb, gb}};
    const double coef[4][4] = {{1., 0., 0., 0.}, {-1., 1., 0., 0.}, {-1., 0., 1., 0.}, {-1., 0., 0., 1.}};  // [row][a]
    QuadTables q;
    memset(&q, 0, sizeof(q));
    for (int g = 0; g < 4; ++g) {
        volatile double N[4];  // volatile: keep the reference's rounding sequence (no contraction)
        for (int a = 0; a < 4; ++a) {
            volatile double v = coef[0][a];
            volatile double t1 = coef[1][a] * pts[g][0];
            v = v + t1;
            volatile double t2 = coef[2][a] * pts[g][1];
            v = v + t2;
            volatile double t3 = coef[3][a] * pts[g][2];
            v = v + t3;
            N[a] = v;
        }
        q.W4 += w;
        for (int a = 0; a < 4; ++a) {
            volatile double wn = w * N[a];
            q.S[a] += wn;
            for (int b = 0; b < 4; ++b) {
                volatile double nn = N[a] * N[b];
                volatile double wnn = w * nn;
                q.NN[4 * a + b] += wnn;
            }
        }
    }
    return q;
}

const QuadTables& quad_tables() {
    static const QuadTables q = make_quad_tables();
    return q;
}

// The select form of the table look-ups (tefem_element.cuh) must return the table entries bit for bit.
static bool quad_select_checked() {
    static const bool ok = quad_select_exact(quad_tables().S, quad_tables().NN);
    return ok;
}

struct TileDims {
    int n_cta, max_items, max_elems, max_inc, max_slots, max_split, max_nodes;
    int pieces_first;   // item layout of every tile: 0 = inline (split lists), 1 = pieces first + one item per block
};

}  // namespace tefem

using namespace tefem;

extern "C" int tefem_quadrature_tables(double* h_W4, double* h_S, double* h_NN) {
    const QuadTables& q = quad_tables();
    if (h_W4) *h_W4 = q.W4;
    if (h_S) memcpy(h_S, q.S, sizeof(q.S));
    if (h_NN) memcpy(h_NN, q.NN, sizeof(q.NN));
    return TEFEM_OK;
}

#ifdef __CUDACC__
// ================================================================================================
// Device kernels
// ================================================================================================
namespace tefem {

__constant__ QuadTables c_quad;

static int upload_quad() {
    TEFEM_REQUIRE(quad_select_checked(), "quadrature tables do not have the value structure the select form assumes");
    static bool done = false;   // per process; constant memory is per device context of this module
    static int dev_done = -1;
    int dev = 0;
    TEFEM_CUDA_CHECK(cudaGetDevice(&dev));
    if (!done || dev_done != dev) {
        TEFEM_CUDA_CHECK(cudaMemcpyToSymbol(c_quad, &quad_tables(), sizeof(QuadTables)));
        done = true;
        dev_done = dev;
    }
    return TEFEM_OK;
}

// Threads per CTA: 256 (default; the 512-block tiles of symbolic.py are two passes).  The pieces-first kernel is also
// instantiated for 128 and 64 threads - more, smaller CTAs per SM, i.e. more independent barrier domains (a third of the
// stall samples of the 256-thread kernel wait at one of the three barriers of a tile, profiles/README.md) - selected
// with tefem_assemble_set_threads (measurement knob; pair it with items_per_cta = 2 x threads).
// 512 threads per SM in every case, so that no instantiation is capped below 128 registers.
#ifndef ASM_THREADS_N
#define ASM_THREADS_N 256
#endif
constexpr int ASM_THREADS = ASM_THREADS_N;
#ifndef ASM_THREADS_PER_SM
#define ASM_THREADS_PER_SM 512
#endif

// ---- TMA bulk copy + mbarrier (sm_90+; SASS: UBLKCP / SYNCS) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline uint32_t up16(uint32_t bytes) { return (bytes + 15u) & ~15u; }

// Shared-memory layout of one prefetch stage (byte offsets, all multiples of 16).
struct StageLayout {
    uint32_t hdr, xyz, conn8, it_inc, it_meta, sp_block, sp_meta, inc, bytes;
};
__host__ __device__ inline StageLayout stage_layout(int max_items, int max_elems, int max_inc, int max_split, int max_nodes) {
    StageLayout L;
    uint32_t o = 0;
    L.hdr = o; o += 64;
    L.xyz = o; o += up16(24u * (uint32_t)((max_nodes + 1) & ~1));
    L.conn8 = o; o += up16(4u * (uint32_t)(max_elems + 3));
    L.it_inc = o; o += up16(4u * (uint32_t)(max_items + 3));
    L.it_meta = o; o += up16(4u * (uint32_t)(max_items + 3));
    L.sp_block = o; o += up16(4u * (uint32_t)(max_split + 3));
    L.sp_meta = o; o += up16(4u * (uint32_t)(max_split + 3));
    L.inc = o; o += up16(2u * (uint32_t)(max_inc + 7));
    L.bytes = o;
    return L;
}

// Thread 0: issue the bulk copies of tile `t`, whose header (h0, h1, h2) was loaded one iteration earlier, into
// stage buffer `st` (completing on `bar`).
__device__ __forceinline__ void prefetch_tile(const AssembleArgs& A, int t, const int4& h0, const int4& h1, const int4& h2,
                                              char* st, const StageLayout& L, uint64_t* bar) {
    const int4* hg = reinterpret_cast<const int4*>(A.cta_hdr) + 3 * (size_t)t;
    const int n_items = h0.x, n_elems = h0.y, n_inc = h0.z, n_split = h0.w, n_nodes = h1.x;
    const int o_items = h1.y, o_elems = h1.z, o_inc = h1.w, o_split = h2.x, o_nodes = h2.y;
    const uint32_t b_items = up16(4u * n_items), b_elems = up16(4u * n_elems), b_inc = up16(2u * n_inc);
    const uint32_t b_split = up16(4u * n_split), b_xyz = up16(24u * n_nodes);
    const uint32_t total = 48u + b_xyz + b_elems + 2u * b_items + 2u * b_split + b_inc;
    mbar_expect_tx(bar, total);
    bulk_g2s(st + L.hdr, hg, 48u, bar);
    if (b_xyz) bulk_g2s(st + L.xyz, A.cta_xyz + 3 * (size_t)o_nodes, b_xyz, bar);
    if (b_elems) bulk_g2s(st + L.conn8, A.cta_conn8 + o_elems, b_elems, bar);
    if (b_items) {
        bulk_g2s(st + L.it_inc, A.item_inc + o_items, b_items, bar);
        bulk_g2s(st + L.it_meta, A.item_meta + o_items, b_items, bar);
    }
    if (b_split) {
        bulk_g2s(st + L.sp_block, A.split_block + o_split, b_split, bar);
        bulk_g2s(st + L.sp_meta, A.split_meta + o_split, b_split, bar);
    }
    if (b_inc) bulk_g2s(st + L.inc, A.inc + o_inc, b_inc, bar);
}

// Shared memory: [element records][partial-sum slots][stage 0][stage 1][2 mbarriers][S, NN tables].
// The quadrature tables are indexed by the per-incidence local node numbers: from shared memory (20 distinct
// banks, broadcast) instead of constant memory, where a warp's divergent indices would be serialised.
template <int OPS, int DUU, bool PF, int NT>
__global__ void __launch_bounds__(NT, ASM_THREADS_PER_SM / NT) assemble_kernel(const AssembleArgs A, const TileDims D, int rec_doubles,
                                                                  int part_doubles) {
    extern __shared__ __align__(128) double smem[];
    double* recs = smem;
    double* partials = recs + rec_doubles;
    const StageLayout L = stage_layout(D.max_items, D.max_elems, D.max_inc, D.max_split, D.max_nodes);
    char* stage0 = reinterpret_cast<char*>(partials + part_doubles);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + 2 * (size_t)L.bytes);
#ifdef ASM_TAB_SELECT
    const QuadS tabS = make_quad_s(c_quad.S);        // selected by compares, no shared-memory tables (tefem_element.cuh)
    const QuadNN tabNN = make_quad_nn(c_quad.NN);
#else
    double* tabS_s = reinterpret_cast<double*>(bars + 2);
    double* tabNN_s = tabS_s + 4;
    if (threadIdx.x < 4) tabS_s[threadIdx.x] = c_quad.S[threadIdx.x];
    if (threadIdx.x < 16) tabNN_s[threadIdx.x] = c_quad.NN[threadIdx.x];
    const QuadS tabS = make_quad_s(tabS_s);
    const QuadNN tabNN = make_quad_nn(tabNN_s);
#endif
    typedef AsmTraits<OPS, DUU> T;
    const int G = gridDim.x;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_proxy_async();
    }
    __syncthreads();
    // thread 0 keeps the header of the next tile to prefetch in registers, loaded one iteration ahead
    const int4* hdr_g = reinterpret_cast<const int4*>(A.cta_hdr);
    int4 nh0 = make_int4(0, 0, 0, 0), nh1 = nh0, nh2 = nh0;
    if (threadIdx.x == 0 && (int)blockIdx.x < A.n_cta) {
        const int4* hg = hdr_g + 3 * (size_t)blockIdx.x;
        prefetch_tile(A, blockIdx.x, __ldg(hg), __ldg(hg + 1), __ldg(hg + 2), stage0, L, &bars[0]);
        if ((int)blockIdx.x + G < A.n_cta) {
            const int4* hn = hdr_g + 3 * (size_t)(blockIdx.x + G);
            nh0 = __ldg(hn); nh1 = __ldg(hn + 1); nh2 = __ldg(hn + 2);
        }
    }
    const double W4 = c_quad.W4;
    int k = 0;
    for (int t = blockIdx.x; t < A.n_cta; t += G, ++k) {
        const int s = k & 1;
        char* st = stage0 + (size_t)s * L.bytes;
        // prefetch the next tile of this CTA into the other stage (its last readers passed the barrier that
        // ends the previous iteration; the proxy fence orders their generic reads before the async writes)
        if (threadIdx.x == 0 && t + G < A.n_cta) {
            fence_proxy_async();
            prefetch_tile(A, t + G, nh0, nh1, nh2, stage0 + (size_t)(s ^ 1) * L.bytes, L, &bars[s ^ 1]);
            if (t + 2 * G < A.n_cta) {
                const int4* hn = hdr_g + 3 * (size_t)(t + 2 * G);
                nh0 = __ldg(hn); nh1 = __ldg(hn + 1); nh2 = __ldg(hn + 2);
            }
        }
        // wait for this tile's segments
        const uint32_t parity = (uint32_t)(k >> 1) & 1u;
        while (!mbar_try_wait(&bars[s], parity)) {}
        const int32_t* hdr = reinterpret_cast<const int32_t*>(st + L.hdr);
        const int n_items = hdr[H_NITEMS], n_elems = hdr[H_NELEMS], o_elems = hdr[H_OELEMS];
        const int32_t block0 = hdr[H_BLOCK0];
        const double* xyz = reinterpret_cast<const double*>(st + L.xyz);
        const uint32_t* conn8 = reinterpret_cast<const uint32_t*>(st + L.conn8);
        const int32_t* it_inc = reinterpret_cast<const int32_t*>(st + L.it_inc);
        const int32_t* it_meta = reinterpret_cast<const int32_t*>(st + L.it_meta);
        const uint16_t* inc_s = reinterpret_cast<const uint16_t*>(st + L.inc);
        // phase 1: geometry records of the staged elements
        for (int le = threadIdx.x; le < n_elems; le += NT) {
            const double det = stage_element_local<T::WANT_D>(xyz, conn8[le], recs + (size_t)le * T::REC);
            if (!(det > 0.0)) {  // error path only: elements.py:207-210
                const int32_t e = A.cta_elems[o_elems + le];
                if (det < 0.0) atomicMin(A.bad_elem, e);
                else atomicMin(A.bad_elem + 1, e);
            }
        }
        __syncthreads();
        if (PF) {   // pieces-first layout
            const int n_piece = hdr[H_NPIECE];
            // phase 2a: the pieces of the split blocks (diagonal chunks, material interfaces; lists of similar length)
            // park their values; phase 2b: one thread per block, in block order (one copy of the incidence loop)
#pragma unroll 1
            for (int ph = (n_piece > 0 ? 0 : 1); ph < 2; ++ph) {
                const int lo = ph ? n_piece : 0, hi = ph ? n_items : n_piece;
                for (int it = lo + threadIdx.x; it < hi; it += NT)
                    process_pf_item<OPS, DUU>(A, ph ? block0 + (it - lo) : -1, (uint32_t)it_inc[it], it_meta[it], recs, inc_s,
                                              partials, W4, tabS, tabNN);
                if (ph == 0) __syncthreads();
            }
        } else {
            // phase 2: work items in block order
            for (int it = threadIdx.x; it < n_items; it += NT)
                process_item<OPS, DUU>(A, block0, (uint32_t)it_inc[it], it_meta[it], recs, inc_s, partials, W4, tabS, tabNN);
            // phase 3: blocks with several items (the diagonal blocks, material interfaces)
            const int n_split = hdr[H_NSPLIT];
            if (n_split > 0) {   // uniform over the CTA
                __syncthreads();
                const int32_t* sp_block = reinterpret_cast<const int32_t*>(st + L.sp_block);
                const int32_t* sp_meta = reinterpret_cast<const int32_t*>(st + L.sp_meta);
                for (int q = threadIdx.x; q < n_split; q += NT)
                    combine_split<OPS, DUU>(A, block0 + sp_block[q], sp_meta[q], partials);
            }
        }
        __syncthreads();   // records, partials and this stage are free again
    }
}

__global__ void __launch_bounds__(128) element_matrices_kernel(
    const double* __restrict__ coords, const int32_t* __restrict__ conn, const int32_t* __restrict__ mat_id,
    const double* __restrict__ mat_table, const int64_t* __restrict__ elem_idx, int64_t n, int kinds,
    double* Muu, double* Dtu, double* Dtt, double* Kuu, double* Kut, double* Ktt, int32_t* bad_elem) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t e = elem_idx ? elem_idx[t] : t;
    AssembleArgs A;
    A.coords = coords; A.conn = conn; A.mat_id = mat_id; A.mat_table = mat_table;
    double rec[REC_STRIDE_MAX];
    const double det = stage_element(A, (int32_t)e, rec);
    if (!(det > 0.0)) {
        if (det < 0.0) atomicMin(bad_elem, (int32_t)e);
        else atomicMin(bad_elem + 1, (int32_t)e);
    }
    element_dense(rec, mat_table + MAT_STRIDE * mat_id[e], c_quad.W4, make_quad_s(c_quad.S), make_quad_nn(c_quad.NN), kinds,
                  Muu ? Muu + t * 144 : nullptr, Dtu ? Dtu + t * 48 : nullptr, Dtt ? Dtt + t * 16 : nullptr,
                  Kuu ? Kuu + t * 144 : nullptr, Kut ? Kut + t * 48 : nullptr, Ktt ? Ktt + t * 16 : nullptr);
}

template <int OPS, int DUU, bool PF, int NT>
static int launch_assemble_pf(const AssembleArgs& A, const TileDims& d, cudaStream_t st) {
    const int rec_doubles = (d.max_elems * AsmTraits<OPS, DUU>::REC + 15) & ~15;   // keep what follows 128-byte aligned
    const int part_doubles = (d.max_slots * AsmTraits<OPS, DUU>::SLOT + 15) & ~15;
    const StageLayout L = stage_layout(d.max_items, d.max_elems, d.max_inc, d.max_split, d.max_nodes);
    const size_t smem = ((size_t)rec_doubles + part_doubles) * sizeof(double) + 2 * (size_t)L.bytes + 16 + 20 * sizeof(double);
    TEFEM_REQUIRE(smem <= 220 * 1024, "tile exceeds shared memory");
    static int sm_count = 0;
    static bool attr_set = false;
    if (!attr_set) {
        TEFEM_CUDA_CHECK(cudaFuncSetAttribute(assemble_kernel<OPS, DUU, PF, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        int dev = 0;
        TEFEM_CUDA_CHECK(cudaGetDevice(&dev));
        TEFEM_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    int per_sm = 0;
    TEFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, assemble_kernel<OPS, DUU, PF, NT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count * per_sm;
    if (grid > d.n_cta) grid = d.n_cta;
    assemble_kernel<OPS, DUU, PF, NT><<<grid, NT, smem, st>>>(A, d, rec_doubles, part_doubles);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

// Threads per CTA of the pieces-first kernel: 256 (default) | 128 | 64, set with tefem_assemble_set_threads (measurement
// knob; the environment is not consulted on the launch path).  A fourth form - ONE persistent 512-thread CTA per SM
// whose warps grab 32-entry chunks from shared-memory counters, no CTA barrier, records double-buffered, TMA
// triple-buffered - was built and measured in round 2 (bitwise equal planes; K 1.39 - 1.52 ms, M + D + K 1.66 - 1.75 ms
// against 0.83 / 1.28 ms on the 100^3 box) and removed: all 16 warps then sit in the dependency chain of the SAME tile,
// and what hides the latencies of a tile is other, independent tiles in flight (profiles/README.md).
static int g_asm_threads = ASM_THREADS;
static int asm_threads() { return g_asm_threads; }

template <int OPS, int DUU>
static int launch_assemble(const AssembleArgs& A, const TileDims& d, cudaStream_t st) {
    const int nt = asm_threads();
    if (!d.pieces_first) {
        TEFEM_REQUIRE(nt == ASM_THREADS, "tefem_assemble_set_threads applies to the pieces-first item layout only");
        return launch_assemble_pf<OPS, DUU, false, ASM_THREADS>(A, d, st);
    }
    switch (nt) {
        case ASM_THREADS: return launch_assemble_pf<OPS, DUU, true, ASM_THREADS>(A, d, st);
#if ASM_THREADS_N == 256
        case 128: return launch_assemble_pf<OPS, DUU, true, 128>(A, d, st);
        case 64: return launch_assemble_pf<OPS, DUU, true, 64>(A, d, st);
#endif
        default: break;
    }
    set_error("threads per CTA of the assembly kernel must be 64, 128 or 256 (tefem_assemble_set_threads)");
    return TEFEM_ERR_INVALID;
}

template <int OPS>
static int dispatch_duu(const AssembleArgs& A, int duu, const TileDims& d, cudaStream_t st) {
    if (!(OPS & TEFEM_OP_D)) return launch_assemble<OPS, 0>(A, d, st);
    switch (duu) {
        case 0: return launch_assemble<OPS, 0>(A, d, st);
        case 1: return launch_assemble<OPS, 1>(A, d, st);
        default: return launch_assemble<OPS, 2>(A, d, st);
    }
}

}  // namespace tefem

extern "C" int tefem_element_matrices(const double* coords, const int32_t* conn, const int32_t* mat_id,
                                      const double* mat_table, int n_mat, const int64_t* elem_idx, int64_t n,
                                      int kinds, double* Muu, double* Dtu, double* Dtt, double* Kuu, double* Kut,
                                      double* Ktt, int32_t* bad_elem, void* stream) {
    TEFEM_REQUIRE(coords && conn && mat_id && mat_table && bad_elem, "null input");
    TEFEM_REQUIRE(n_mat > 0 && n >= 0, "sizes");
    if (n == 0) return TEFEM_OK;
    int rc = upload_quad();
    if (rc) return rc;
    element_matrices_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream)>>>(
        coords, conn, mat_id, mat_table, elem_idx, n, kinds, Muu, Dtu, Dtt, Kuu, Kut, Ktt, bad_elem);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_assemble(const double* mat_table, int n_mat, int32_t n_cta, const int32_t* h_max,
                              const int32_t* cta_hdr, const int32_t* cta_elems, const int32_t* cta_conn8,
                              const double* cta_xyz, const int32_t* item_inc, const int32_t* item_meta,
                              const int32_t* split_block, const int32_t* split_meta, const uint16_t* inc, int64_t n_blocks,
                              int ops, double alpha_M,
                              double alpha_K, double* vals_M, double* vals_K, double* vals_D, int32_t* bad_elem,
                              void* stream) {
    TEFEM_REQUIRE(mat_table && bad_elem && h_max, "null input");
    TEFEM_REQUIRE(cta_hdr && cta_elems && cta_conn8 && cta_xyz && item_inc && item_meta && inc, "null schedule");
    const int max_items = h_max[0], max_elems = h_max[1], max_inc = h_max[2], max_slots = h_max[3], max_split = h_max[4],
              max_nodes = h_max[5], pieces_first = h_max[6];
    TEFEM_REQUIRE(pieces_first == 0 || (pieces_first == 1 && max_split == 0), "item layout flag");
    TEFEM_REQUIRE(max_split == 0 || (split_block && split_meta), "null split lists");
    const uintptr_t align_or = (uintptr_t)cta_hdr | (uintptr_t)cta_conn8 | (uintptr_t)cta_xyz | (uintptr_t)item_inc |
                               (uintptr_t)item_meta | (uintptr_t)split_block | (uintptr_t)split_meta | (uintptr_t)inc;
    TEFEM_REQUIRE((align_or & 15) == 0, "schedule arrays must be 16-byte aligned");
    TEFEM_REQUIRE(ops > 0 && ops < 8, "ops mask");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_M) || vals_M, "vals_M");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_K) || vals_K, "vals_K");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_D) || vals_D, "vals_D");
    TEFEM_REQUIRE(max_elems > 0 && max_elems <= 4096, "max elements per tile must fit 12 bits");
    TEFEM_REQUIRE(max_nodes > 0 && max_nodes <= 256, "max nodes per tile must fit a byte");
    TEFEM_REQUIRE(max_items > 0 && max_inc >= 0 && max_inc < 65536 && max_slots >= 0 && max_slots < 65536 && max_split >= 0,
                  "tile limits");
    TEFEM_REQUIRE(n_mat > 0 && n_mat <= 256, "1 .. 256 materials (the work items carry the material index in 8 bits)");
    if (n_cta == 0) return TEFEM_OK;
    int rc = upload_quad();
    if (rc) return rc;
    AssembleArgs A;
    memset(&A, 0, sizeof(A));
    A.mat_table = mat_table;
    A.cta_hdr = cta_hdr; A.cta_elems = cta_elems; A.cta_conn8 = cta_conn8; A.cta_xyz = cta_xyz;
    A.item_inc = item_inc; A.item_meta = item_meta; A.split_block = split_block; A.split_meta = split_meta; A.inc = inc;
    A.n_cta = n_cta; A.n_blocks = n_blocks; A.alpha_M = alpha_M; A.alpha_K = alpha_K;
    A.vals_M = vals_M; A.vals_K = vals_K; A.vals_D = vals_D; A.bad_elem = bad_elem;
    const int duu = (alpha_K != 0.0) ? 2 : (alpha_M != 0.0 ? 1 : 0);
    const TileDims d = {n_cta, max_items, max_elems, max_inc, max_slots, max_split, max_nodes, pieces_first};
    cudaStream_t st = as_stream(stream);
    switch (ops) {
        case 1: return dispatch_duu<1>(A, duu, d, st);
        case 2: return dispatch_duu<2>(A, duu, d, st);
        case 3: return dispatch_duu<3>(A, duu, d, st);
        case 4: return dispatch_duu<4>(A, duu, d, st);
        case 5: return dispatch_duu<5>(A, duu, d, st);
        case 6: return dispatch_duu<6>(A, duu, d, st);
        default: return dispatch_duu<7>(A, duu, d, st);
    }
}

extern "C" int tefem_assemble_set_threads(int threads) {
    TEFEM_REQUIRE(threads == 64 || threads == 128 || threads == 256, "threads per CTA must be 64, 128 or 256");
    g_asm_threads = threads;
    return TEFEM_OK;
}

extern "C" int tefem_check_jacobian(const int32_t* bad_elem, int32_t* h_elem, void* stream) {
    int32_t h[2];
    TEFEM_CUDA_CHECK(cudaMemcpyAsync(h, bad_elem, sizeof(h), cudaMemcpyDeviceToHost, as_stream(stream)));
    TEFEM_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    if (h[0] == INT_MAX && h[1] == INT_MAX) return TEFEM_OK;
    const bool neg_first = h[0] < h[1];
    const int32_t e = neg_first ? h[0] : h[1];
    if (h_elem) *h_elem = e;
    set_error("Element %d has %s jacobian.", e, neg_first ? "negative" : "zero");
    return neg_first ? TEFEM_ERR_NEG_JACOBIAN : TEFEM_ERR_ZERO_JACOBIAN;
}

#else  // ------------------------------------------------------------------------------------------
// TEFEM_HOST_EMU: sequential execution of the SAME per-thread routines (tests/host_emulation only).
// ================================================================================================
template <int OPS, int DUU>
static int emu_run(const AssembleArgs& A, const TileDims& d, int n_mat) {
    typedef typename AsmTraits<OPS, DUU>::Acc Acc;
    const QuadTables& q = quad_tables();
    typedef AsmTraits<OPS, DUU> T;
    std::vector<double> recs((size_t)d.max_elems * T::REC);
    std::vector<double> partials((size_t)d.max_slots * T::SLOT + 1);
    for (int cta = 0; cta < d.n_cta; ++cta) {
        const int32_t* h = A.cta_hdr + (size_t)HDR_STRIDE * cta;
        if (h[H_NELEMS] > d.max_elems || h[H_NITEMS] > d.max_items || h[H_NINC] > d.max_inc || h[H_NSPLIT] > d.max_split ||
            h[H_NNODES] > d.max_nodes)
            return TEFEM_ERR_INVALID;
        // every segment must start on a 16-byte boundary (bulk copies)
        if ((h[H_OITEMS] & 3) || (h[H_OELEMS] & 3) || (h[H_OINC] & 7) || (h[H_OSPLIT] & 3) || (h[H_ONODES] & 1)) return TEFEM_ERR_INVALID;
        const double* xyz = A.cta_xyz + 3 * (size_t)h[H_ONODES];
        const uint16_t* inc_s = A.inc + h[H_OINC];
        for (int le = 0; le < h[H_NELEMS]; ++le) {
            const uint32_t cn = (uint32_t)A.cta_conn8[h[H_OELEMS] + le];
            for (int a = 0; a < 4; ++a)
                if ((int)((cn >> (8 * a)) & 0xff) >= h[H_NNODES]) return TEFEM_ERR_INVALID;
            const double det = stage_element_local<T::WANT_D>(xyz, cn, recs.data() + (size_t)le * T::REC);
            if (!(det > 0.0)) {
                const int32_t e = A.cta_elems[h[H_OELEMS] + le];
                int32_t* slot = A.bad_elem + (det < 0.0 ? 0 : 1);
                if (e < *slot) *slot = e;
            }
        }
        const int n_piece = h[H_NPIECE];
        if (n_piece < 0 || n_piece > h[H_NITEMS] || n_piece > d.max_slots || (n_piece > 0 && h[H_NSPLIT] != 0)) return TEFEM_ERR_INVALID;
        if (!d.pieces_first && n_piece != 0) return TEFEM_ERR_INVALID;
        for (int it = 0; it < h[H_NITEMS]; ++it) {
            const int32_t meta = A.item_meta[h[H_OITEMS] + it];
            const int slot = meta_slot(meta);
            const uint32_t iw = (uint32_t)A.item_inc[h[H_OITEMS] + it];
            const int32_t s = (int32_t)(iw & 0xffffu);
            if (meta_mat(meta) >= n_mat) return TEFEM_ERR_INVALID;
            if (d.pieces_first) {
                if (it < n_piece) {   // piece k parks in slot k
                    if (slot != it || s + meta_len(meta) > h[H_NINC]) return TEFEM_ERR_INVALID;
                    process_pf_item<OPS, DUU>(A, -1, iw, meta, recs.data(), inc_s, partials.data(), q.W4, make_quad_s(q.S), make_quad_nn(q.NN));
                } else {              // block items: all pieces were processed before (the kernel synchronises)
                    const int b = it - n_piece;
                    if ((int)(iw >> 16) != b) return TEFEM_ERR_INVALID;
                    if (slot < 0 ? s + meta_len(meta) > h[H_NINC] : slot + meta_len(meta) > n_piece) return TEFEM_ERR_INVALID;
                    process_pf_item<OPS, DUU>(A, h[H_BLOCK0] + b, iw, meta, recs.data(), inc_s, partials.data(), q.W4, make_quad_s(q.S), make_quad_nn(q.NN));
                }
                continue;
            }
            if (slot >= d.max_slots || s + meta_len(meta) > h[H_NINC]) return TEFEM_ERR_INVALID;
            process_item<OPS, DUU>(A, h[H_BLOCK0], iw, meta, recs.data(), inc_s, partials.data(), q.W4, make_quad_s(q.S), make_quad_nn(q.NN));
        }
        for (int k = 0; k < h[H_NSPLIT]; ++k) {
            const int32_t sm = A.split_meta[h[H_OSPLIT] + k];
            if ((sm & 0xffff) + (int)((uint32_t)sm >> 16) > d.max_slots) return TEFEM_ERR_INVALID;
            combine_split<OPS, DUU>(A, h[H_BLOCK0] + A.split_block[h[H_OSPLIT] + k], sm, partials.data());
        }
    }
    return TEFEM_OK;
}

extern "C" int tefem_emu_assemble(const double* mat_table, int n_mat, int32_t n_cta, const int32_t* h_max, const int32_t* cta_hdr,
                                  const int32_t* cta_elems, const int32_t* cta_conn8,
                                  const double* cta_xyz, const int32_t* item_inc, const int32_t* item_meta,
                                  const int32_t* split_block, const int32_t* split_meta, const uint16_t* inc, int64_t n_blocks,
                                  int ops, double alpha_M, double alpha_K,
                                  double* vals_M, double* vals_K, double* vals_D, int32_t* bad_elem) {
    AssembleArgs A;
    memset(&A, 0, sizeof(A));
    A.mat_table = mat_table;
    A.cta_hdr = cta_hdr; A.cta_elems = cta_elems; A.cta_conn8 = cta_conn8; A.cta_xyz = cta_xyz;
    A.item_inc = item_inc; A.item_meta = item_meta; A.split_block = split_block; A.split_meta = split_meta; A.inc = inc;
    A.n_cta = n_cta; A.n_blocks = n_blocks; A.alpha_M = alpha_M; A.alpha_K = alpha_K;
    A.vals_M = vals_M; A.vals_K = vals_K; A.vals_D = vals_D; A.bad_elem = bad_elem;
    const TileDims d = {n_cta, h_max[0], h_max[1], h_max[2], h_max[3], h_max[4], h_max[5], h_max[6]};
    const int duu = (alpha_K != 0.0) ? 2 : (alpha_M != 0.0 ? 1 : 0);
#define TEFEM_EMU_CASE(O)                                                     \
    case O:                                                                   \
        if (duu == 0 || !((O) & TEFEM_OP_D)) return emu_run<O, 0>(A, d, n_mat);      \
        if (duu == 1) return emu_run<O, 1>(A, d, n_mat);                             \
        return emu_run<O, 2>(A, d, n_mat);
    switch (ops) {
        TEFEM_EMU_CASE(1) TEFEM_EMU_CASE(2) TEFEM_EMU_CASE(3) TEFEM_EMU_CASE(4)
        TEFEM_EMU_CASE(5) TEFEM_EMU_CASE(6) TEFEM_EMU_CASE(7)
        default: return TEFEM_ERR_INVALID;
    }
#undef TEFEM_EMU_CASE
}

extern "C" int tefem_emu_element_matrices(const double* coords, const int32_t* conn, const int32_t* mat_id,
                                          const double* mat_table, int64_t n, int kinds, double* Muu, double* Dtu,
                                          double* Dtt, double* Kuu, double* Kut, double* Ktt, int32_t* bad_elem) {
    AssembleArgs A;
    memset(&A, 0, sizeof(A));
    A.coords = coords; A.conn = conn; A.mat_id = mat_id; A.mat_table = mat_table;
    const QuadTables& q = quad_tables();
    for (int64_t e = 0; e < n; ++e) {
        double rec[REC_STRIDE_MAX];
        const double det = stage_element(A, (int32_t)e, rec);
        if (!(det > 0.0)) {
            int32_t* slot = bad_elem + (det < 0.0 ? 0 : 1);
            if (e < *slot) *slot = (int32_t)e;
        }
        element_dense(rec, mat_table + MAT_STRIDE * mat_id[e], q.W4, make_quad_s(q.S), make_quad_nn(q.NN), kinds, Muu ? Muu + e * 144 : nullptr, Dtu ? Dtu + e * 48 : nullptr,
                      Dtt ? Dtt + e * 16 : nullptr, Kuu ? Kuu + e * 144 : nullptr, Kut ? Kut + e * 48 : nullptr,
                      Ktt ? Ktt + e * 16 : nullptr);
    }
    return TEFEM_OK;
}
#endif
