/*
 * tefem.h - C ABI of the B200-native thermoelastic FEM hot path (libtefem_b200.so).
 *
 * The reference (rcapillon/thermoelasticity_fem) is pure Python and has no FFI layer; its
 * "plugin boundary" for this path is the set of numpy/scipy call sites listed in SURVEY.md 8(b).
 * Every entry point below names the reference call site(s) it replaces (file:line relative to the
 * reference repository).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers (taken from torch tensors with .data_ptr()) unless the
 *     parameter name starts with `h_` (host pointer);
 *   - the library never allocates device memory for the caller's data: Python/torch owns every
 *     buffer; the library owns only its constant tables and an opaque solver/communicator handle;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless documented;
 *   - every function returns an int status: 0 = TEFEM_OK, negative = error class; the message of
 *     the last error of the calling thread is returned by tefem_last_error();
 *   - fp64 everywhere; node-level indices are int32, value offsets are computed in 64 bits.
 *
 * Storage formats
 *   pattern   node-block CSR: rowptr[N+1] (int32), colidx[P] (int32, ascending inside a row);
 *             P = number of touched node pairs.  Block p of row i couples node i with node colidx[p].
 *   planes    operator values as structure-of-arrays planes: vals[plane*P + p].  Plane layouts:
 *               K (13 planes): uu(i,j) -> 3*i+j (0..8), ut(i) -> 9+i, tt -> 12      (no tu block)
 *               M ( 3 planes): uu(i,i) -> i                                          (component diagonal)
 *               D (4 / 7 / 13 planes): tu(j) -> j, tt -> 3; then uu: none (alpha_M = alpha_K = 0),
 *                 uu(i,i) -> 4+i (alpha_K = 0), uu(i,j) -> 4+3*i+j (alpha_K != 0)
 *             so that the bytes written equal 8 * (touched non-zeros), SURVEY.md 8(d).
 *   system    solver matrix as array-of-structures 4x4 blocks: sys[p*16 + 4*r + c], 128-byte aligned.
 */
#ifndef TEFEM_H
#define TEFEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEFEM_OK                 0
#define TEFEM_ERR_INVALID       -1   /* bad argument                                             */
#define TEFEM_ERR_CUDA          -2   /* CUDA runtime error (message has the CUDA string)         */
#define TEFEM_ERR_NEG_JACOBIAN  -3   /* elements.py:207-208  'Element {n} has negative jacobian.' */
#define TEFEM_ERR_ZERO_JACOBIAN -4   /* elements.py:209-210  'Element {n} has zero jacobian.'     */
#define TEFEM_ERR_NO_CONVERGE   -5   /* Krylov solve did not reach the tolerance                 */
#define TEFEM_ERR_BREAKDOWN     -6   /* Krylov breakdown that restarts could not cure            */
#define TEFEM_ERR_NCCL          -7   /* NCCL error                                               */

#define TEFEM_OP_M 1
#define TEFEM_OP_K 2
#define TEFEM_OP_D 4

#define TEFEM_KIND_MUU 1
#define TEFEM_KIND_DTU 2
#define TEFEM_KIND_DTT 4
#define TEFEM_KIND_KUU 8
#define TEFEM_KIND_KUT 16
#define TEFEM_KIND_KTT 32

/* Message of the last error raised on the calling thread (NUL-terminated, truncated to n). */
int tefem_last_error(char* h_buf, size_t n);

/* Number of CUDA kernels this library has launched in the calling process (monotonic). */
int64_t tefem_launch_count(void);

/* Library / build information: writes "tefem_b200 <version> sm_100a cuda <ver>" into h_buf. */
int tefem_version(char* h_buf, size_t n);

/* Quadrature tables exactly as the reference builds them at import time
 * (elements.py:65-69 truncated Gauss literals, :76-87 shape values, :165-172):
 *   h_W4 = w+w+w+w, h_S[4] = sum_g w N_a(g), h_NN[16] = sum_g w N_a(g) N_b(g).                */
int tefem_quadrature_tables(double* h_W4, double* h_S, double* h_NN);

/* Per-element matrices for a list of elements (parity / Element.compute_mat_*_e):
 * replaces Element.compute_mat_{Muu,Dtu,Dtt,Kuu,Kut,Ktt}_e, elements.py:220-286, including the
 * Jacobian of elements.py:199-218.  elem_idx may be NULL (= 0..n-1).  Outputs (row-major, may be
 * NULL when the kind is not requested): Muu[n,12,12] Dtu[n,4,12] Dtt[n,4,4] Kuu[n,12,12]
 * Kut[n,12,4] Ktt[n,4,4].  mat_table[n_mat,8] = (rho, lambda, mu, k, c, beta, T0, pad),
 * materials.py:18-29.  bad_elem[2] (device int32, initialised to INT32_MAX by the caller):
 * smallest element index with det < 0 / det == 0.                                              */
int tefem_element_matrices(const double* coords, const int32_t* conn, const int32_t* mat_id,
                           const double* mat_table, int n_mat,
                           const int64_t* elem_idx, int64_t n, int kinds,
                           double* Muu, double* Dtu, double* Dtt, double* Kuu, double* Kut, double* Ktt,
                           int32_t* bad_elem, void* stream);

/* Element quadrature + atomics-free scatter into the node-block pattern (kernel K1).
 * Replaces Model.assemble_M / assemble_K / assemble_D, model.py:87-118 (the per-element
 * compute_mat_*_e calls, the np.ix_ scatter into the dense matrix and csc_array(dense)).
 *
 * Owner-computes TILE schedule built once per mesh by the symbolic phase (host side, symbolic.py): a tile owns a
 * contiguous range of node rows and everything it needs lives in per-tile segments that start on 16-byte
 * boundaries, so that the persistent kernel can prefetch its next tile with TMA bulk copies while it computes:
 *   cta_hdr[c][12] = {n_items, n_elems, n_inc, n_split, n_nodes, o_items, o_elems, o_inc, o_split, o_nodes,
 *                     first block, n_pieces}   (a tile owns a contiguous block range starting at `first block`);
 *   cta_xyz[o_nodes ..][3]     coordinates of the tile's nodes (copy of coords[cta_nodes]);
 *   cta_conn8 / cta_elems [o_elems ..]   staged elements: 4 tile-local node numbers (one byte each), element
 *                              number (error reporting);
 *   inc[o_inc ..]              incidence entries (tile-local element << 4) | (a << 2) | b, a, b = local node numbers
 *                              of (row, column), block after block, ASCENDING ELEMENT inside a block (the order of
 *                              the reference's loop, model.py:89,96,107);
 *   item_inc / item_meta [o_items ..]   work items: one block, or one piece of a block's list (lists are cut where
 *                              the element material changes, and long runs - the diagonal blocks - into chunks):
 *                              ALL ELEMENTS OF AN ITEM SHARE ONE MATERIAL, row (item_meta >> 24) & 0xff of mat_table,
 *                              whose constants multiply the item's geometric sums once; the thread sums
 *                              inc[(item_inc & 0xffff) .. + (item_meta & 0xff)) (tile-relative) and either writes
 *                              block `first block + (item_inc >> 16)` (items are in block order: the plane stores
 *                              of a warp are contiguous) or, when ((item_meta >> 8) & 0xffff) - 1 >= 0, parks the
 *                              values in that shared-memory slot;
 *   split_block / split_meta [o_split ..]   blocks with several items: tile-local block, first slot | n_items << 16
 *                              (added in list order = ascending element);
 *   h_max[7] = largest n_items, n_elems, n_inc, partial slots, n_split, n_nodes of a tile (shared-memory sizes), and
 *              the ITEM LAYOUT of the tiles: 0 = inline (above), 1 = pieces first: cta_hdr[c][11] = n_pieces; the item
 *              list of a tile starts with all pieces of its split blocks (piece k parks its values in slot k) and
 *              continues with exactly ONE item per block in block order - a plain block, or a combine item
 *              (item_meta = n_pieces | (first slot + 1) << 8) that adds the parked pieces in list order; the kernel
 *              runs the pieces, synchronises, then the block items, so that warps hold lists of similar length and
 *              every block is stored at its block-order position (no split lists: n_split = 0).  Same chunks, same
 *              summation order, bitwise the same values as layout 0.
 *   Every output value is written exactly once: no atomics, bitwise reproducible.
 * ops = bit mask of TEFEM_OP_*; vals_* are planes (see above) and may be NULL when not requested.
 * The D layout follows alpha_M / alpha_K (model.py:112-117).
 * bad_elem[2] as above; the call itself does not synchronise - use tefem_check_jacobian().      */
int tefem_assemble(const double* mat_table, int n_mat, int32_t n_cta, const int32_t* h_max,
                   const int32_t* cta_hdr, const int32_t* cta_elems, const int32_t* cta_conn8,
                   const double* cta_xyz, const int32_t* item_inc, const int32_t* item_meta,
                   const int32_t* split_block, const int32_t* split_meta, const uint16_t* inc,
                   int64_t n_blocks, int ops, double alpha_M, double alpha_K,
                   double* vals_M, double* vals_K, double* vals_D,
                   int32_t* bad_elem, void* stream);

/* Threads per CTA of the pieces-first assembly kernel: 64, 128 or 256 (default; measurement knob - more, smaller CTAs per
 * SM need tiles of 2 x threads items).  Host-side setting of the calling process; no reference counterpart.      */
int tefem_assemble_set_threads(int threads);

/* Bank-aware tile-local numbering of the staged elements of every tile (HOST function, all pointers are host pointers;
 * no reference counterpart - part of the symbolic phase).  The record of staged element e sits at rec[stride * e] in shared
 * memory, so e mod 16 decides the banks its gathers hit; for every pair of elements that the lanes of one half-warp read in
 * the same load the function counts at which class differences they collide, then exchanges the classes of two elements
 * inside a block of 16 consecutive staged elements while that lowers the count (the numbering stays a bijection onto
 * 0 .. n_elems - 1 and phase 1 of the kernel keeps its access pattern).
 * strides: bit 0 = kernels without D (record stride 13), bit 1 = with D (stride 17).  h_new_local[o_elems + old number] <- new
 * number; h_n_slots[c] <- n_elems of tile c; h_cost (optional, [2]) <- pairwise collision count before / after.
 * The caller permutes cta_elems / cta_conn8 and rewrites the element field of the incidence entries; lists, order and
 * values are unchanged (bitwise).                                                                                  */
int tefem_schedule_bank_numbering(int32_t n_cta, const int32_t* h_cta_hdr, const int32_t* h_item_inc,
                                  const int32_t* h_item_meta, const uint16_t* h_inc, int strides, int n_sweeps,
                                  int n_threads, int32_t* h_new_local, int32_t* h_n_slots, int64_t* h_cost);

/* Synchronises the stream, reads bad_elem[2] and returns TEFEM_OK, TEFEM_ERR_NEG_JACOBIAN or
 * TEFEM_ERR_ZERO_JACOBIAN for the FIRST offending element in element order (the one the
 * reference's loop would have raised on); *h_elem receives its index.                          */
int tefem_check_jacobian(const int32_t* bad_elem, int32_t* h_elem, void* stream);

/* y[4N] (+)= A x for an operator in plane storage (used for the Dirichlet lifting
 * K[free, dir] @ x_dir of model.py:326-339 and the Newmark right-hand side D@(.) + K@(.) of
 * solvers.py:167-170).  plane_map[16]: plane index of entry (r, c) of a block, or -1.
 * y = beta_y * y + alpha * A x.                                                                */
int tefem_planes_spmv(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int64_t n_blocks,
                      const double* vals, const int8_t* h_plane_map,
                      const double* x, double alpha, double beta_y, double* y, void* stream);

/* Dirichlet elimination + solver-matrix formation (kernel K2).  Replaces
 * apply_dirichlet_matrices (model.py:296-302) and `M_ff + gamma dt D_ff + beta dt^2 K_ff`
 * (solvers.py:159-161): sys = cM*M + cD*D + cK*K on the shared pattern, with every row and column
 * of a constrained DOF (dir_mask[4N] != 0) replaced by the identity, which is the reduced system
 * A[free,:][:,free] embedded in the uniform 4x4 block structure.  Any of the
 * plane pointers may be NULL (coefficient ignored).  d_layout = number of D planes (4, 7, 13).   */
int tefem_form_system(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int64_t n_blocks,
                      const double* vals_M, double cM, const double* vals_D, int d_layout, double cD,
                      const double* vals_K, double cK, const uint8_t* dir_mask,
                      double* sys, void* stream);

/* Block-Jacobi setup (kernel K5): inverts the 4x4 diagonal block of every node row (partial
 * pivoting) into dinv[N,16] and, when scale != 0, left-scales the system in place,
 * sys <- blockdiag(dinv) * sys, so that BiCGSTAB needs no separate preconditioner application
 * (scale = 0 keeps sys: tefem_cg applies dinv symmetrically).
 * diag_idx[N] = position of the diagonal block of each row; singular: one int32 of caller-owned
 * device memory (scratch; the library allocates nothing).  Synchronises the stream.            */
int tefem_block_jacobi_scale(const int32_t* rowptr, const int32_t* diag_idx, int32_t n_rows,
                             double* sys, double* dinv, int32_t* singular, int scale, void* stream);

/* z[4N] = blockdiag(dinv) * r   (applies the same scaling to a right-hand side); entries of the
 * constrained DOFs (dir_mask != 0, may be NULL) are set to 0 (their unknown is eliminated).      */
int tefem_block_jacobi_apply(const double* dinv, int32_t n_rows, const double* r, double* z,
                             const uint8_t* dir_mask, void* stream);

/* y = sys * x on the 4x4 block format (kernel K3), 4 threads per node row.                     */
int tefem_spmv(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows,
               const double* sys, const double* x, double* y, void* stream);
/* Launch configuration of the SpMV (threads per CTA, resident CTAs per SM of the grid-stride grid, blocks per trip);
 * default 256, 8, 2 (tuned at config B).  Measurement knob, host-side setting of the calling process.              */
int tefem_spmv_set_config(int threads, int ctas_per_sm, int unroll);

/* Opaque Krylov solver handle (kernel K4/K6 driver).  Work vectors are provided by the caller:
 * work must hold tefem_solver_work_doubles(n_rows, n_halo) doubles, 128-byte aligned.                                    */
typedef struct tefem_solver tefem_solver;
int64_t tefem_solver_work_doubles(int32_t n_rows, int32_t n_halo);
int tefem_solver_create(tefem_solver** out, int32_t n_rows, int32_t n_halo, double* work, void* comm /* tefem_comm* or NULL */);
int tefem_solver_destroy(tefem_solver* s);
/* SpMV timing for the roofline report: returns the accumulated device time (CUDA events on the launching
 * stream around every SpMV of the Krylov iteration, halo exchange included in the multi-GPU case) and the
 * number of SpMVs since profiling was enabled; enable = 1 / 0 switches it and clears the counters,
 * enable < 0 only reads.                                                                            */
int tefem_solver_profile(tefem_solver* s, int enable, double* h_spmv_ms, int64_t* h_spmv_count);

/* Halo description for the row-partitioned multi-GPU solve (ignored when comm == NULL):
 * x vectors have n_rows owned node rows followed by n_halo halo node rows;
 * send: for each neighbour k, the owned node rows send_idx[send_ptr[k] .. send_ptr[k+1]) go to
 * rank h_nbr[k]; recv: halo rows [recv_ptr[k], recv_ptr[k+1]) (relative to n_rows) arrive from it.
 * PEER-MEMORY HALO (default; the communicator has vector slots, tefem_comm_vec_alloc): the SpMV input vectors of
 * the solver live in the slots; before an SpMV a push kernel stores the boundary values straight into the halo
 * tails of the neighbours' copies (remote stores over NVLink) and raises a flag, and the SpMV - ONE launch over
 * all rows - waits for its neighbours' flags: no packing, no NCCL call, no second launch.  send_dst[k] /
 * send_dpos[k] (device arrays, one per send entry): destination rank and node position in THAT rank's local
 * vector (its n_rows + halo offset).  boundary_rows / interior_rows may be NULL.
 * NCCL HALO (communicator without vector slots; kept as the measured alternative): pack kernel + grouped
 * ncclSend/ncclRecv on a second stream overlapped with the interior rows, boundary rows in a second launch;
 * needs the interior / boundary row lists, send_dst / send_dpos may be NULL.                                   */
int tefem_solver_set_halo(tefem_solver* s, int32_t n_halo, int n_nbr, const int32_t* h_nbr,
                          const int32_t* h_send_ptr, const int32_t* send_idx, const int32_t* send_dst,
                          const int32_t* send_dpos, const int32_t* h_recv_ptr,
                          const int32_t* boundary_rows, int32_t n_boundary_rows,
                          const int32_t* interior_rows, int32_t n_interior_rows);
/* Peer-memory halo: rows [lo, hi) have no halo column (for row blocks in node order the interior is one contiguous
 * range; lo == hi when unknown).  fused != 0 (default): the SpMV is ONE kernel that stores this rank's boundary values
 * into the neighbours' vectors, multiplies the interior rows, waits for the neighbours' flags and multiplies the
 * boundary rows; 0: separate push launch, the SpMV waits before its first row (comparison).                          */
int tefem_solver_set_interior_range(tefem_solver* s, int32_t lo, int32_t hi, int fused);
/* number of peer-memory vector slots a solver needs (tefem_comm_vec_alloc) */
int tefem_solver_peer_slots(void);

/* Aggregation-based coarse correction, used as the right preconditioner of tefem_bicgstab on top of the block-Jacobi
 * scaling: z = omega r + P1 M1(P1^T r) on the scaled system, with a replicated level 1 (one V-cycle with damped
 * Jacobi) and a dense level 2 (csrc/tefem_precond.cu).  The coarse space of an aggregate holds its rigid-body modes
 * (3 translations, 3 rotations about its centre) and a constant theta, stored as two 4-DOF pseudo nodes
 * (t_x, t_y, t_z, theta), (w_x, w_y, w_z, 0).  The reference has no counterpart (spsolve is a direct solve,
 * solvers.py:26,172).  All arrays are built by the host side (multilevel.py) and owned by the caller:
 *   agg_ptr[n_agg+1], agg_nodes[n_rows]: owned fine node rows grouped by aggregate; node_agg[n_rows]: aggregate of
 *   each owned fine node; node_off[n_rows][4] (float): offset of the node from its aggregate's centre (x, y, z, 0);
 *   rowptr1 / colidx1 / sys1: 4x4-block CSR over the 2 n_agg pseudo nodes of A1^ = D1^-1 P1^T A^ P1 (D1 = its 8x8
 *   diagonal blocks); dinv1[n_agg][64]: D1^-1, row-major 8x8; omega_c, n_pre >= 1, n_post >= 0: damping and number
 *   of the Jacobi sweeps on level 1 before / after the level-2 correction;
 *   r2_* / p2_*: the transfer operators P2^T (2 n_agg2 block rows) and P2 (2 n_agg block rows) between the pseudo
 *   nodes of levels 1 and 2 as 4x4-block CSR; inv2: dense row-major inverse [8 n_agg2, 8 n_agg2] of P2^T A1^ P2;
 *   sys1, r2_vals, p2_vals and inv2 are SINGLE precision (preconditioner data only; all vectors and the fine
 *   operator are fp64);
 *   work: tefem_precond_work_doubles(n_agg, n_agg2) doubles, 128-byte aligned; comm: tefem_comm* or NULL (the
 *   restricted residual, 8 n_agg doubles, is summed over the ranks through peer memory; levels 1 and 2 are
 *   replicated on every rank).                                                                                  */
typedef struct tefem_precond tefem_precond;
int64_t tefem_precond_work_doubles(int32_t n_agg, int32_t n_agg2);
int tefem_precond_create(tefem_precond** out, int32_t n_rows, int32_t n_agg, const int32_t* agg_ptr,
                         const int32_t* agg_nodes, const int32_t* node_agg, const float* node_off, double omega,
                         const int32_t* rowptr1, const int32_t* colidx1, const float* sys1, const double* dinv1,
                         double omega_c, int n_pre, int n_post, int32_t n_agg2, const int32_t* r2_rowptr,
                         const int32_t* r2_colidx,
                         const float* r2_vals, const int32_t* p2_rowptr, const int32_t* p2_colidx, const float* p2_vals,
                         const float* inv2, double* work, void* comm);
int tefem_precond_destroy(tefem_precond* pc);
int tefem_precond_apply(tefem_precond* pc, const double* r, double* z, void* stream);
/* Multi-GPU: agg_ranks[n_agg] (uint8 device array, bit r set iff rank r owns fine nodes of the aggregate) restricts the
 * in-kernel all-reduce of the restricted residual to the ranks that contribute (the others hold exact zeros).      */
int tefem_precond_set_agg_ranks(tefem_precond* pc, const uint8_t* agg_ranks);
/* Multi-GPU, after set_agg_ranks: agg_list[n] (int32 device array) = the aggregates with fine nodes on THIS rank; the
 * restriction then launches one CTA per listed aggregate instead of one per aggregate of the whole mesh.           */
int tefem_precond_set_agg_list(tefem_precond* pc, const int32_t* agg_list, int32_t n);
/* Level-1 operator in HALF precision (same block layout as sys1, 16 halves per block, 32-byte aligned; NULL detaches):
 * the six level-1 sweeps of an application then read half the bytes.  Preconditioner data only.                     */
int tefem_precond_set_level1_half(tefem_precond* pc, const void* sys1_half);
/* Level-1 cycle: overlapped != 0 (default): the last pre-smoothing sweep and the level-2 correction are formed from the
 * same residual (one level-1 product less per application); 0: classic V-cycle (all pre-sweeps, then the residual).  */
int tefem_precond_set_cycle(tefem_precond* pc, int overlapped);
/* Coarse cycle of an application.  mode 0: separate launches on levels replicated on every rank (~10 dependent small
 * kernels per application whatever the number of GPUs).  mode 1: ONE persistent kernel (plain launch, grid within the
 * co-resident limit); the rows of levels 1 and 2 are PARTITIONED over the ranks, every coarse vector has one replica per
 * rank in the peer-memory region in a self-validating layout (16 bytes per value: two 8-byte units {32 data bits, 32-bit
 * tag}), the rank that owns a row computes it and stores it into the replicas of the ranks that read it (remote stores
 * over NVLink, tefem_precond_set_cycle_masks), the restriction pushes its partial sums the same way to the owner of the
 * aggregate, and a consumer retries its load until the tags match: no barrier, fence or flag (csrc/tefem_precond.cu,
 * coarse_cycle_kernel).  cycle_work: ZEROED, 128-byte aligned buffer of tefem_precond_cycle_doubles(...) doubles owned
 * by the caller (mode 1; NULL for mode 0).  With a communicator its ZEROED region must have been allocated with
 * big_cap_doubles >= tefem_precond_peer_doubles(...) and every rank must apply the preconditioner the same number of
 * times; without one the same kernel runs on cycle_work alone.  Call after set_cycle / set_level1_half.              */
int tefem_precond_set_coarse_mode(tefem_precond* pc, int mode, double* cycle_work);
int64_t tefem_precond_peer_doubles(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post);
/* Persistent cycle on several GPUs: x_mask[n_agg] / t_mask[n_agg] (uint8 device arrays, bit q) = rank q reads the level-1
 * iterates / the residual t of the aggregate; the owner stores a row only into those replicas (NULL: into all).  Row
 * ownership: aggregate i belongs to rank q with n_agg q / world <= i < n_agg (q + 1) / world (integer division), level-2
 * aggregate J likewise with n_agg2; the owner of J computes rows 8 J .. 8 J + 7 of P2^T t and of the dense product.  */
int tefem_precond_set_cycle_masks(tefem_precond* pc, const uint8_t* x_mask, const uint8_t* t_mask);
int64_t tefem_precond_cycle_doubles(int32_t n_agg, int32_t n_agg2, int n_pre, int n_post);
/* Attach (or detach with NULL) a preconditioner to the Krylov driver. */
int tefem_solver_set_precond(tefem_solver* s, tefem_precond* pc);

/* Fills the halo part of x (rows n_rows .. n_rows + n_halo) with the owners' values: pack kernel +
 * grouped ncclSend/ncclRecv on `stream`.  No-op without a communicator.                          */
int tefem_solver_halo_exchange(tefem_solver* s, double* x, void* stream);

/* Left-block-Jacobi-scaled BiCGSTAB (replaces spsolve at solvers.py:26 and :172).
 * Solves sys x = b where sys and b have ALREADY been scaled by tefem_block_jacobi_scale/apply.
 * Stops when, for each field f in {u rows, theta rows}, ||(b - sys x)_f|| <= rtol * ||b_f|| (a field with
 * b_f = 0 is measured against ||x_f||): the TRUE residual is re-evaluated before returning and
 * recursive-residual drift triggers a restart from the current x.  When two restarts in a row cannot
 * improve the true residual (the eps*cond floor of working precision) the solve stops: it returns OK only
 * if the residual is within floor_factor * rtol (floor_factor >= 1; 1 = never accept above rtol) and then
 * raises h_info[4], otherwise TEFEM_ERR_NO_CONVERGE.  x holds the initial guess on entry.
 * h_info[6] = {iterations, restarts, final relative residual, ||b||, floor flag (1: stalled and accepted
 * above rtol), stall count}.                                                                    */
int tefem_bicgstab(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx,
                   const double* sys, const double* b, double* x,
                   double rtol, int max_iter, int check_every, double floor_factor, double* h_info,
                   void* stream);

/* Block-Jacobi-preconditioned CG for a SYMMETRIC POSITIVE DEFINITE system in the same block layout (sys UNSCALED,
 * dinv from tefem_block_jacobi_scale(.., scale = 0)).  Used by the staged steady-state solve: the reference's K has
 * no theta -> u block (model.py:94-103), so spsolve(K_ff, F_f) (solvers.py:26) is K_tt theta = F_t followed by
 * K_uu u = F_u - K_ut theta, each SPD once the other field's DOFs are eliminated (identity rows).  One SpMV per
 * iteration, a persistent cooperative kernel per batch of check_every iterations.  Same stopping rule (per field,
 * on dinv * residual, true residual verified) and h_info as tefem_bicgstab.  Single GPU.          */
int tefem_cg(tefem_solver* s, const int32_t* rowptr, const int32_t* colidx, const double* sys,
             const double* dinv, const double* b, double* x, double rtol, int max_iter, int check_every,
             double floor_factor, double* h_info, void* stream);

/* Fused Newmark kernels (kernel K7), solvers.py:167-178.
 * predictor: w1 = v + (1-gamma) dt a ; w2 = x + dt v + 0.5 dt^2 (1-2 beta) a   (zero on constrained DOFs)
 * corrector: v <- v + (1-gamma) dt a + gamma dt a_new ; x <- x + dt v_old + 0.5 dt^2 ((1-2beta) a + 2 beta a_new);
 *            a <- a_new, all masked to the free DOFs (constrained DOFs keep x, v = a = 0).       */
int tefem_newmark_predict(int64_t n, const double* x, const double* v, const double* a,
                          double dt, double gamma, double beta, const uint8_t* dir_mask,
                          double* w1, double* w2, void* stream);
int tefem_newmark_correct(int64_t n, double* x, double* v, double* a, const double* a_new,
                          double dt, double gamma, double beta, const uint8_t* dir_mask, void* stream);

/* out = c0 x0 + c1 x1 + c2 x2 (x1, x2 may be NULL).  Used to extrapolate the previous accelerations into the
 * initial guess of the next step's solve (3 a_n - 3 a_{n-1} + a_{n-2}); the reference has no counterpart
 * (spsolve is direct, solvers.py:172).                                                            */
int tefem_lincomb3(int64_t n, double c0, const double* x0, double c1, const double* x1, double c2,
                   const double* x2, double* out, void* stream);

/* F[4*node[i] + c] += w[i] * h_coef[c], c = 0..3: the load vector as geometric nodal weights x a
 * time-dependent amplitude (model.py:120-286: area/3 and volume/4 lumping).  node[] unique.     */
int tefem_load_axpy(int64_t n_entries, const int32_t* node, const double* w, const double* h_coef,
                    double* F, void* stream);
/* Same, with the four amplitudes in DEVICE memory (d_coef[4]): the end-to-end path copies the step's amplitudes from
 * pinned host memory instead of passing them as kernel arguments.                                                  */
int tefem_load_axpy_dev(int64_t n_entries, const int32_t* node, const double* w, const double* d_coef,
                        double* F, void* stream);

/* ---- multi-GPU communicator (NCCL over NVLink), one process per GPU ------------------------- */
typedef struct tefem_comm tefem_comm;
/* h_id: 128-byte ncclUniqueId buffer; created on rank 0, broadcast by the caller (torch.distributed). */
int tefem_comm_unique_id(const char* h_nccl_lib_path, char* h_id);
int tefem_comm_create(tefem_comm** out, const char* h_nccl_lib_path, const char* h_id, int n_ranks, int rank);
int tefem_comm_destroy(tefem_comm* c);
/* in-place sum all-reduce of n doubles with NCCL (set-up paths) */
int tefem_comm_allreduce_sum(tefem_comm* c, double* buf, int n, void* stream);
/* Peer-memory region for the reductions of the solve (csrc/tefem_peer.cuh): the dot-product packets of BiCGSTAB and
 * the restricted residual of the coarse correction are all-reduced INSIDE the kernels that consume them, by direct
 * loads from the other GPUs' memory over NVLink / NVSwitch (no NCCL call, no extra launch).  Every rank allocates
 * its region (big_cap_doubles >= 4 * n_coarse of the preconditioner, 0 without one) and receives a 64-byte CUDA IPC
 * handle in h_handle; the caller gathers the handles of all ranks (rank order, n_ranks * 64 bytes) and passes them
 * to tefem_comm_shm_open on every rank.  One node, at most 8 ranks.  tefem_comm_check reports a time-out of a
 * peer wait (call after a stream synchronisation; tefem_bicgstab does).                                          */
int tefem_comm_shm_alloc(tefem_comm* c, int64_t big_cap_doubles, char* h_handle);
int tefem_comm_shm_open(tefem_comm* c, const char* h_handles);
/* A time-out marks the communicator failed: later waits return at once and tefem_comm_check keeps reporting it until
 * tefem_comm_reset_error (after the ranks have been resynchronised).  tefem_comm_set_timeout: seconds a peer wait may
 * spin (default 30).                                                                                               */
int tefem_comm_check(tefem_comm* c);
int tefem_comm_reset_error(tefem_comm* c);
int tefem_comm_set_timeout(tefem_comm* c, double seconds);
/* Peer-mapped vector slots: n_slots vectors of slot_doubles doubles (identical on all ranks; slot_doubles >= the
 * largest 4 * (n_rows + n_halo) over the ranks), allocated by the library with cudaMalloc because they must be
 * shareable over CUDA IPC - the one exception to "the caller owns every buffer".  Same handle exchange as above.
 * Create them BEFORE tefem_solver_create: a solver of a communicator with slots keeps its SpMV input vectors there. */
int tefem_comm_vec_alloc(tefem_comm* c, int n_slots, int64_t slot_doubles, char* h_handle);
int tefem_comm_vec_open(tefem_comm* c, const char* h_handles);

#ifdef __cplusplus
}
#endif
#endif /* TEFEM_H */
