// K1: element quadrature + atomics-free owner-computes scatter into the node-block pattern.
//
// Replaces the reference's assembly loops (model.py:87-118): per element the six
// Element.compute_mat_*_e (elements.py:220-286), the np.ix_ scatter into a dense n_dofs^2 array
// and csc_array(dense).  Design (DESIGN.md "K1"):
//
//   * the symbolic phase groups node rows into CTA tiles; a CTA first stages, in shared memory, the
//     geometry record (inverse Jacobian columns = shape-function gradients, det, material-scaled
//     weights) of every element that touches one of its rows (phase 1: coalesced int4 loads of the
//     tile's connectivity copy, gathers of node coordinates through L2) and, with 16-byte loads, the
//     tile's contiguous range of incidence entries;
//   * then one thread per WORK ITEM - a node-pair block, or a chunk of <= max_chunk incidences of a
//     block with a long incidence list (the diagonal blocks) - walks its precomputed incidence
//     entries (element, a, b) in ascending element order - the slot map - and accumulates the block's
//     13 (K) / 1 (M) / 4 (D) values in registers (phase 2).  Unsplit blocks are written straight to
//     the structure-of-arrays planes; chunks park their partial sums in shared memory and one thread
//     per split block adds them in chunk order (phase 3).  Every output value is written exactly
//     once: no atomics, bitwise reproducible, summation in the reference's element order (chunked
//     for the split blocks).
//
// The same per-thread routines are compiled for the host when TEFEM_HOST_EMU is defined; that
// build is used only by tests/host_emulation to check the schedule logic in a GPU-less container
// and is never loaded by the package.
#include "tefem_element.cuh"

#include <limits.h>
#include <vector>

namespace tefem {

struct AssembleArgs {
    const double* coords;
    const int32_t* conn;       // element_matrices path only
    const int32_t* mat_id;     // element_matrices path only
    const double* mat_table;
    const int32_t* cta_ptr;    // [n_cta, 12]: item0, item1, elem0, elem1, inc0, inc1, split0, split1, node0, node1, -, -
    const int32_t* cta_elems;  // staged element numbers (error reporting only)
    const int32_t* cta_conn8;  // [n_staged] connectivity of the staged elements as 4 tile-local node bytes
    const int32_t* cta_mat;    // [n_staged] material index of the staged elements
    const int32_t* cta_nodes;  // [n_tile_nodes] nodes whose coordinates the tile stages
    const int32_t* item_block; // [n_items] block p of the work item
    const int32_t* item_blockT;  // [n_items] mirror block (j, i) of an upper block (i, j), or -1
    const uint32_t* item_inc;  // [n_items] first incidence entry of the work item (absolute index into inc)
    const int32_t* item_meta;  // [n_items] length | (partial slot + 1) << 8 | upper << 24  (slot + 1 == 0: unsplit)
    const int32_t* split_block;  // [n_split] block p of a split block
    const int32_t* split_meta;   // [n_split] first partial slot | n_chunks << 16
    const uint16_t* inc;
    int64_t n_blocks;
    double alpha_M, alpha_K;
    double* vals_M;
    double* vals_K;
    double* vals_D;
    int32_t* bad_elem;
};

enum { CTA_ITEM0 = 0, CTA_ITEM1, CTA_ELEM0, CTA_ELEM1, CTA_INC0, CTA_INC1, CTA_SPLIT0, CTA_SPLIT1, CTA_NODE0, CTA_NODE1,
       CTA_STRIDE = 12 };
enum { ITEM_UPPER = 1 << 24 };

// ---- per-thread routines (host + device) -------------------------------------------------------

// Geometry record of element e read through the global connectivity (element_matrices path).
TEFEM_HD double stage_element(const AssembleArgs& A, int32_t e, double W4, double* rec) {
    const int32_t n0 = A.conn[4 * (int64_t)e], n1 = A.conn[4 * (int64_t)e + 1];
    const int32_t n2 = A.conn[4 * (int64_t)e + 2], n3 = A.conn[4 * (int64_t)e + 3];
    const double* mat = A.mat_table + MAT_STRIDE * A.mat_id[e];
    return element_record(A.coords + 3 * (int64_t)n0, A.coords + 3 * (int64_t)n1,
                          A.coords + 3 * (int64_t)n2, A.coords + 3 * (int64_t)n3, mat, W4, rec);
}

// Phase 1 from the tile's staged node coordinates xyz[tile-local node][3] and the byte-packed tile-local
// connectivity (no gather from global memory per element).
TEFEM_HD double stage_element_local(const AssembleArgs& A, const double* xyz, uint32_t conn8, int32_t mat_idx, double W4,
                                    double* rec) {
    return element_record(xyz + 3 * (conn8 & 0xff), xyz + 3 * ((conn8 >> 8) & 0xff), xyz + 3 * ((conn8 >> 16) & 0xff),
                          xyz + 3 * (conn8 >> 24), A.mat_table + MAT_STRIDE * mat_idx, W4, rec);
}

template <int OPS, int DUU>
struct AsmTraits {
    static constexpr bool WANT_M = (OPS & TEFEM_OP_M) != 0, WANT_K = (OPS & TEFEM_OP_K) != 0, WANT_D = (OPS & TEFEM_OP_D) != 0;
    static constexpr bool NEED_K = WANT_K || (WANT_D && DUU == 2);
    static constexpr bool NEED_M = WANT_M || (WANT_D && DUU >= 1);
    typedef BlockAcc<NEED_K, NEED_M, WANT_D> Acc;
};

// Phase 2: acc = sum over the incidence entries inc[s .. t) (ascending element order).
template <int OPS, int DUU>
TEFEM_HD void accumulate_list(typename AsmTraits<OPS, DUU>::Acc& acc, const double* recs, const uint16_t* inc,
                              uint32_t s, uint32_t t, const double* S, const double* NN) {
    typedef AsmTraits<OPS, DUU> T;
    for (uint32_t k = s; k < t; ++k) {
        const uint32_t w = inc[k];
        block_accumulate<T::NEED_K, T::NEED_M, T::WANT_D>(acc, recs + (size_t)(w >> 4) * REC_STRIDE, (w >> 2) & 3, w & 3, S, NN);
    }
}

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

// One work item.  Upper block (i, j): accumulate it together with the non-symmetric parts of its mirror (j, i)
// and write both.  Diagonal chunk: accumulate, then write the block or park the partial sums.
template <int OPS, int DUU>
TEFEM_HD void process_item(const AssembleArgs& A, int32_t p, int32_t pT, uint32_t s, int32_t meta, const double* recs,
                           const uint16_t* inc_s, double* partials, const double* S, const double* NN) {
    typedef AsmTraits<OPS, DUU> T;
    typename T::Acc acc;
    acc.clear();
    const uint32_t t = s + (uint32_t)(meta & 0xff);
    if (meta & ITEM_UPPER) {
        MirrorAcc<T::NEED_K, T::WANT_D> mir;
        mir.clear();
        for (uint32_t k = s; k < t; ++k) {
            const uint32_t w = inc_s[k];
            pair_accumulate<T::NEED_K, T::NEED_M, T::WANT_D>(acc, mir, recs + (size_t)(w >> 4) * REC_STRIDE, (w >> 2) & 3,
                                                             w & 3, S, NN);
        }
        store_block<OPS, DUU>(A, p, acc);
        if (pT >= 0) {
            if (T::NEED_K) {   // Kuu_ji = Kuu_ij^T ; Kut_ji from the mirror accumulators
                double tmp;
                tmp = acc.kuu[1]; acc.kuu[1] = acc.kuu[3]; acc.kuu[3] = tmp;
                tmp = acc.kuu[2]; acc.kuu[2] = acc.kuu[6]; acc.kuu[6] = tmp;
                tmp = acc.kuu[5]; acc.kuu[5] = acc.kuu[7]; acc.kuu[7] = tmp;
                for (int i = 0; i < 3; ++i) acc.kut[i] = mir.kut[i];
            }
            if (T::WANT_D)
                for (int i = 0; i < 3; ++i) acc.dtu[i] = mir.dtu[i];
            store_block<OPS, DUU>(A, pT, acc);
        }
    } else {
        accumulate_list<OPS, DUU>(acc, recs, inc_s, s, t, S, NN);
        const int slot = ((meta >> 8) & 0xffff) - 1;
        if (slot < 0) store_block<OPS, DUU>(A, p, acc);
        else acc.dump(partials + (size_t)slot * T::Acc::LIVE);
    }
}

// Phase 3 for one split block: add the chunk partials in chunk order and write the block.
template <int OPS, int DUU>
TEFEM_HD void combine_split(const AssembleArgs& A, int32_t p, int32_t meta, const double* partials) {
    typedef typename AsmTraits<OPS, DUU>::Acc Acc;
    Acc acc;
    acc.clear();
    const int slot0 = meta & 0xffff, n = meta >> 16;
    for (int c = 0; c < n; ++c) acc.add_from(partials + (size_t)(slot0 + c) * Acc::LIVE);
    store_block<OPS, DUU>(A, p, acc);
}

// Dense element matrices of one element (parity path; elements.py:220-286).
TEFEM_HD void element_dense(const double* rec, const double* S, const double* NN, int kinds,
                            double* Muu, double* Dtu, double* Dtt, double* Kuu, double* Kut, double* Ktt) {
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            BlockAcc<true, true, true> acc;
            acc.clear();
            block_accumulate<true, true, true>(acc, rec, a, b, S, NN);
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

static void fill_args(AssembleArgs& A, const double* coords, const double* mat_table, const int32_t* cta_ptr,
                      const int32_t* cta_elems, const int32_t* cta_conn8, const int32_t* cta_mat, const int32_t* cta_nodes,
                      const int32_t* item_block, const int32_t* item_blockT, const uint32_t* item_inc, const int32_t* item_meta,
                      const int32_t* split_block, const int32_t* split_meta, const uint16_t* inc, int64_t n_blocks,
                      double alpha_M, double alpha_K, double* vals_M, double* vals_K, double* vals_D, int32_t* bad_elem) {
    A.coords = coords; A.conn = nullptr; A.mat_id = nullptr; A.mat_table = mat_table;
    A.cta_ptr = cta_ptr; A.cta_elems = cta_elems; A.cta_conn8 = cta_conn8; A.cta_mat = cta_mat; A.cta_nodes = cta_nodes;
    A.item_block = item_block; A.item_blockT = item_blockT; A.item_inc = item_inc; A.item_meta = item_meta;
    A.split_block = split_block; A.split_meta = split_meta; A.inc = inc;
    A.n_blocks = n_blocks; A.alpha_M = alpha_M; A.alpha_K = alpha_K;
    A.vals_M = vals_M; A.vals_K = vals_K; A.vals_D = vals_D; A.bad_elem = bad_elem;
}

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

constexpr int ASM_THREADS = 256;

// Shared memory: [element records: rec_doubles][partial sums: part_doubles][node coordinates: xyz_doubles]
// [incidence entries (uint16)].
// A tile's incidence entries are one contiguous range of `inc`: they are copied with 16-byte loads while phase 0
// waits on the coordinates, and the phase-2 loop reads them from shared memory instead of chasing a global load
// per iteration (profiles/README.md: 19 % of all stall samples sat on that load in v1).  Node coordinates are
// staged once per tile (coalesced for consecutive node numbers) instead of 12 scattered 8-byte gathers per element.
template <int OPS, int DUU>
__global__ void __launch_bounds__(ASM_THREADS, 2) assemble_kernel(const AssembleArgs A, int rec_doubles, int part_doubles,
                                                                  int xyz_doubles) {
    extern __shared__ double recs[];
    double* partials = recs + rec_doubles;
    double* xyz = partials + part_doubles;
    uint16_t* inc_s = reinterpret_cast<uint16_t*>(xyz + xyz_doubles);
    const int4* hdr = reinterpret_cast<const int4*>(A.cta_ptr) + 3 * blockIdx.x;
    const int4 c0 = __ldg(hdr), c1 = __ldg(hdr + 1), c2 = __ldg(hdr + 2);
    const int i0 = c0.x, i1 = c0.y, e0 = c0.z, ne = c0.w - c0.z;
    const uint32_t inc_lo = (uint32_t)c1.x, inc_hi = (uint32_t)c1.y;
    const int sp0 = c1.z, sp1 = c1.w;
    const int n0 = c2.x, nn = c2.y - c2.x;
    // phase 0: node coordinates -> shared memory
    for (int k = threadIdx.x; k < nn; k += ASM_THREADS) {
        const double* X = A.coords + 3 * (int64_t)__ldg(A.cta_nodes + n0 + k);
        xyz[3 * k] = X[0]; xyz[3 * k + 1] = X[1]; xyz[3 * k + 2] = X[2];
    }
    // first work item of this thread: issue its index loads now, they are consumed after the barriers
    const int it0 = i0 + threadIdx.x;
    int32_t p0 = 0, pT0 = -1, m0 = 0;
    uint32_t s0 = 0;
    if (it0 < i1) {
        p0 = __ldg(A.item_block + it0);
        pT0 = __ldg(A.item_blockT + it0);
        s0 = __ldg(A.item_inc + it0);
        m0 = __ldg(A.item_meta + it0);
    }
    // incidence entries [inc_lo, inc_hi) -> shared memory, from the 16-byte aligned address below inc_lo
    const uint32_t base = inc_lo & ~7u;
    {
        const uint4* src = reinterpret_cast<const uint4*>(A.inc + base);
        uint4* dst = reinterpret_cast<uint4*>(inc_s);
        const uint32_t n16 = (inc_hi - base + 7u) >> 3;
        for (uint32_t k = threadIdx.x; k < n16; k += ASM_THREADS) dst[k] = __ldg(src + k);
    }
    // connectivity / material of this thread's first staged element (consumed after the barrier)
    uint32_t cn0 = 0;
    int32_t mt0 = 0;
    if ((int)threadIdx.x < ne) {
        cn0 = (uint32_t)__ldg(A.cta_conn8 + e0 + threadIdx.x);
        mt0 = __ldg(A.cta_mat + e0 + threadIdx.x);
    }
    __syncthreads();
    // phase 1: geometry records of the staged elements
    const double W4 = c_quad.W4;
    for (int le = threadIdx.x; le < ne; le += ASM_THREADS) {
        const uint32_t cn = (le == (int)threadIdx.x) ? cn0 : (uint32_t)__ldg(A.cta_conn8 + e0 + le);
        const int32_t m = (le == (int)threadIdx.x) ? mt0 : __ldg(A.cta_mat + e0 + le);
        const double det = stage_element_local(A, xyz, cn, m, W4, recs + (size_t)le * REC_STRIDE);
        if (!(det > 0.0)) {  // error path only: elements.py:207-210
            const int32_t e = A.cta_elems[e0 + le];
            if (det < 0.0) atomicMin(A.bad_elem, e);
            else atomicMin(A.bad_elem + 1, e);
        }
    }
    __syncthreads();
    // phase 2: work items
    if (it0 < i1) process_item<OPS, DUU>(A, p0, pT0, s0 - base, m0, recs, inc_s, partials, c_quad.S, c_quad.NN);
    for (int it = it0 + ASM_THREADS; it < i1; it += ASM_THREADS)
        process_item<OPS, DUU>(A, __ldg(A.item_block + it), __ldg(A.item_blockT + it), __ldg(A.item_inc + it) - base,
                               __ldg(A.item_meta + it), recs, inc_s, partials, c_quad.S, c_quad.NN);
    // phase 3: split (diagonal) blocks
    if (sp1 > sp0) {   // uniform over the CTA
        __syncthreads();
        for (int k = sp0 + threadIdx.x; k < sp1; k += ASM_THREADS)
            combine_split<OPS, DUU>(A, __ldg(A.split_block + k), __ldg(A.split_meta + k), partials);
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
    double rec[REC_STRIDE];
    const double det = stage_element(A, (int32_t)e, c_quad.W4, rec);
    if (!(det > 0.0)) {
        if (det < 0.0) atomicMin(bad_elem, (int32_t)e);
        else atomicMin(bad_elem + 1, (int32_t)e);
    }
    element_dense(rec, c_quad.S, c_quad.NN, kinds,
                  Muu ? Muu + t * 144 : nullptr, Dtu ? Dtu + t * 48 : nullptr, Dtt ? Dtt + t * 16 : nullptr,
                  Kuu ? Kuu + t * 144 : nullptr, Kut ? Kut + t * 48 : nullptr, Ktt ? Ktt + t * 16 : nullptr);
}

struct LaunchDims {
    int n_cta, max_cta_elems, max_cta_inc, max_cta_slots, max_cta_nodes;
};

template <int OPS, int DUU>
static int launch_assemble(const AssembleArgs& A, const LaunchDims& d, cudaStream_t st) {
    const int rec_doubles = (d.max_cta_elems * REC_STRIDE + 1) & ~1;    // keep the areas behind it 16-byte aligned
    const int part_doubles = (d.max_cta_slots * AsmTraits<OPS, DUU>::Acc::LIVE + 1) & ~1;
    const int xyz_doubles = (3 * d.max_cta_nodes + 1) & ~1;
    const size_t smem = ((size_t)rec_doubles + part_doubles + xyz_doubles) * sizeof(double) +
                        ((size_t)d.max_cta_inc + 16) * sizeof(uint16_t);
    TEFEM_REQUIRE(smem <= 200 * 1024, "CTA tile exceeds shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        TEFEM_CUDA_CHECK(cudaFuncSetAttribute(assemble_kernel<OPS, DUU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    assemble_kernel<OPS, DUU><<<d.n_cta, ASM_THREADS, smem, st>>>(A, rec_doubles, part_doubles, xyz_doubles);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

template <int OPS>
static int dispatch_duu(const AssembleArgs& A, int duu, const LaunchDims& d, cudaStream_t st) {
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

extern "C" int tefem_assemble(const double* coords, const double* mat_table, int n_mat, int32_t n_cta,
                              int32_t max_cta_elems, int32_t max_cta_inc, int32_t max_cta_slots, int32_t max_cta_nodes,
                              const int32_t* cta_ptr, const int32_t* cta_elems, const int32_t* cta_conn8,
                              const int32_t* cta_mat, const int32_t* cta_nodes, const int32_t* item_block,
                              const int32_t* item_blockT, const uint32_t* item_inc,
                              const int32_t* item_meta, const int32_t* split_block, const int32_t* split_meta,
                              const uint16_t* inc, int64_t n_blocks, int ops, double alpha_M, double alpha_K,
                              double* vals_M, double* vals_K, double* vals_D, int32_t* bad_elem, void* stream) {
    TEFEM_REQUIRE(coords && mat_table && bad_elem, "null mesh input");
    TEFEM_REQUIRE(cta_ptr && cta_elems && cta_conn8 && cta_mat && cta_nodes && item_block && item_blockT && item_inc &&
                      item_meta && inc, "null schedule");
    TEFEM_REQUIRE(max_cta_slots == 0 || (split_block && split_meta), "null split lists");
    TEFEM_REQUIRE((((uintptr_t)inc) & 15) == 0 && (((uintptr_t)cta_ptr) & 15) == 0, "inc / cta_ptr must be 16-byte aligned");
    TEFEM_REQUIRE(max_cta_nodes > 0 && max_cta_nodes <= 256, "max_cta_nodes must fit a byte");
    TEFEM_REQUIRE(ops > 0 && ops < 8, "ops mask");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_M) || vals_M, "vals_M");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_K) || vals_K, "vals_K");
    TEFEM_REQUIRE(!(ops & TEFEM_OP_D) || vals_D, "vals_D");
    TEFEM_REQUIRE(max_cta_elems > 0 && max_cta_elems <= 4096, "max_cta_elems must fit 12 bits");
    TEFEM_REQUIRE(max_cta_inc >= 0 && max_cta_slots >= 0 && max_cta_slots < 65536, "max_cta_inc / max_cta_slots");
    TEFEM_REQUIRE(n_mat > 0, "n_mat");
    if (n_cta == 0) return TEFEM_OK;
    int rc = upload_quad();
    if (rc) return rc;
    AssembleArgs A;
    fill_args(A, coords, mat_table, cta_ptr, cta_elems, cta_conn8, cta_mat, cta_nodes, item_block, item_blockT, item_inc,
              item_meta, split_block, split_meta, inc, n_blocks, alpha_M, alpha_K, vals_M, vals_K, vals_D, bad_elem);
    const int duu = (alpha_K != 0.0) ? 2 : (alpha_M != 0.0 ? 1 : 0);
    const LaunchDims d = {n_cta, max_cta_elems, max_cta_inc, max_cta_slots, max_cta_nodes};
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
static int emu_run(const AssembleArgs& A, int n_cta, int max_cta_elems, int max_cta_inc, int max_cta_slots,
                   int max_cta_nodes) {
    typedef typename AsmTraits<OPS, DUU>::Acc Acc;
    const QuadTables& q = quad_tables();
    std::vector<double> recs((size_t)max_cta_elems * REC_STRIDE);
    std::vector<double> partials((size_t)max_cta_slots * Acc::LIVE + 1);
    std::vector<uint16_t> inc_s((size_t)max_cta_inc + 16);
    std::vector<double> xyz((size_t)3 * max_cta_nodes + 1);
    for (int cta = 0; cta < n_cta; ++cta) {
        const int32_t* c = A.cta_ptr + (size_t)CTA_STRIDE * cta;
        const int e0 = c[CTA_ELEM0], ne = c[CTA_ELEM1] - e0;
        if (ne > max_cta_elems) return TEFEM_ERR_INVALID;
        const int n0 = c[CTA_NODE0], nn = c[CTA_NODE1] - n0;
        if (nn > max_cta_nodes) return TEFEM_ERR_INVALID;
        for (int k = 0; k < nn; ++k)
            for (int d = 0; d < 3; ++d) xyz[3 * k + d] = A.coords[3 * (size_t)A.cta_nodes[n0 + k] + d];
        // same staging arithmetic as the kernel: aligned-down copy of the tile's incidence range
        const uint32_t inc_lo = (uint32_t)c[CTA_INC0], inc_hi = (uint32_t)c[CTA_INC1];
        const uint32_t base = inc_lo & ~7u;
        const uint32_t n16 = (inc_hi - base + 7u) >> 3;
        if ((size_t)n16 * 8 > inc_s.size()) return TEFEM_ERR_INVALID;
        for (uint32_t k = 0; k < n16 * 8; ++k) inc_s[k] = A.inc[base + k];
        for (int le = 0; le < ne; ++le) {
            const double det = stage_element_local(A, xyz.data(), (uint32_t)A.cta_conn8[e0 + le], A.cta_mat[e0 + le], q.W4,
                                                   recs.data() + (size_t)le * REC_STRIDE);
            if (!(det > 0.0)) {
                const int32_t e = A.cta_elems[e0 + le];
                int32_t* slot = A.bad_elem + (det < 0.0 ? 0 : 1);
                if (e < *slot) *slot = e;
            }
        }
        for (int it = c[CTA_ITEM0]; it < c[CTA_ITEM1]; ++it) {
            const int slot = ((A.item_meta[it] >> 8) & 0xffff) - 1;
            if (slot >= max_cta_slots) return TEFEM_ERR_INVALID;
            process_item<OPS, DUU>(A, A.item_block[it], A.item_blockT[it], A.item_inc[it] - base, A.item_meta[it],
                                   recs.data(), inc_s.data(), partials.data(), q.S, q.NN);
        }
        for (int k = c[CTA_SPLIT0]; k < c[CTA_SPLIT1]; ++k)
            combine_split<OPS, DUU>(A, A.split_block[k], A.split_meta[k], partials.data());
    }
    return TEFEM_OK;
}

extern "C" int tefem_emu_assemble(const double* coords, const double* mat_table, int32_t n_cta, int32_t max_cta_elems,
                                  int32_t max_cta_inc, int32_t max_cta_slots, int32_t max_cta_nodes, const int32_t* cta_ptr,
                                  const int32_t* cta_elems, const int32_t* cta_conn8, const int32_t* cta_mat,
                                  const int32_t* cta_nodes, const int32_t* item_block, const int32_t* item_blockT,
                                  const uint32_t* item_inc, const int32_t* item_meta,
                                  const int32_t* split_block, const int32_t* split_meta, const uint16_t* inc,
                                  int64_t n_blocks, int ops, double alpha_M, double alpha_K,
                                  double* vals_M, double* vals_K, double* vals_D, int32_t* bad_elem) {
    AssembleArgs A;
    fill_args(A, coords, mat_table, cta_ptr, cta_elems, cta_conn8, cta_mat, cta_nodes, item_block, item_blockT, item_inc,
              item_meta, split_block, split_meta, inc, n_blocks, alpha_M, alpha_K, vals_M, vals_K, vals_D, bad_elem);
    const int duu = (alpha_K != 0.0) ? 2 : (alpha_M != 0.0 ? 1 : 0);
#define TEFEM_EMU_CASE(O)                                                                                     \
    case O:                                                                                                   \
        if (duu == 0 || !((O) & TEFEM_OP_D)) return emu_run<O, 0>(A, n_cta, max_cta_elems, max_cta_inc, max_cta_slots, max_cta_nodes); \
        if (duu == 1) return emu_run<O, 1>(A, n_cta, max_cta_elems, max_cta_inc, max_cta_slots, max_cta_nodes);              \
        return emu_run<O, 2>(A, n_cta, max_cta_elems, max_cta_inc, max_cta_slots, max_cta_nodes);
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
    A.coords = coords; A.conn = conn; A.mat_id = mat_id; A.mat_table = mat_table;
    const QuadTables& q = quad_tables();
    for (int64_t e = 0; e < n; ++e) {
        double rec[REC_STRIDE];
        const double det = stage_element(A, (int32_t)e, q.W4, rec);
        if (!(det > 0.0)) {
            int32_t* slot = bad_elem + (det < 0.0 ? 0 : 1);
            if (e < *slot) *slot = (int32_t)e;
        }
        element_dense(rec, q.S, q.NN, kinds, Muu ? Muu + e * 144 : nullptr, Dtu ? Dtu + e * 48 : nullptr,
                      Dtt ? Dtt + e * 16 : nullptr, Kuu ? Kuu + e * 144 : nullptr, Kut ? Kut + e * 48 : nullptr,
                      Ktt ? Ktt + e * 16 : nullptr);
    }
    return TEFEM_OK;
}
#endif
