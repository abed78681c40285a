// K2 / K5: Dirichlet elimination, solver-matrix formation, block-Jacobi scaling, and the
// plane-storage SpMV used for lifting and the Newmark right-hand side.
//
// Reference call sites replaced (rcapillon/thermoelasticity_fem):
//   model.py:296-302   A[free,:][:,free]                        -> tefem_form_system (identity rows/cols)
//   solvers.py:159-161 M_ff + gamma dt D_ff + beta dt^2 K_ff    -> tefem_form_system (cM, cD, cK)
//   model.py:326-339   K[free,:][:,dir] @ x_dir                 -> tefem_planes_spmv
//   solvers.py:168-169 D_ff @ (.) , K_ff @ (.)                  -> tefem_planes_spmv
#include "tefem_common.cuh"

namespace tefem {

struct PlaneMap {
    int8_t m[16];
};

// y = beta_y * y + alpha * A x ; one thread per scalar row (4 consecutive threads share a node row).
__global__ void __launch_bounds__(256) planes_spmv_kernel(const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ colidx, int64_t n_scalar_rows,
                                                          int64_t P, const double* __restrict__ vals, PlaneMap pm,
                                                          const double* __restrict__ x, double alpha, double beta_y,
                                                          double* __restrict__ y) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_scalar_rows) return;
    const int32_t i = (int32_t)(row >> 2);
    const int r = (int)(row & 3);
    const int p0 = pm.m[4 * r + 0], p1 = pm.m[4 * r + 1], p2 = pm.m[4 * r + 2], p3 = pm.m[4 * r + 3];
    double acc = 0.0;
    if (p0 >= 0 || p1 >= 0 || p2 >= 0 || p3 >= 0) {
        const int32_t s = rowptr[i], t = rowptr[i + 1];
        for (int32_t p = s; p < t; ++p) {
            const double* xj = x + 4 * (int64_t)colidx[p];
            if (p0 >= 0) acc += vals[p0 * P + p] * xj[0];
            if (p1 >= 0) acc += vals[p1 * P + p] * xj[1];
            if (p2 >= 0) acc += vals[p2 * P + p] * xj[2];
            if (p3 >= 0) acc += vals[p3 * P + p] * xj[3];
        }
    }
    y[row] = (beta_y == 0.0 ? 0.0 : beta_y * y[row]) + alpha * acc;
}

struct FormArgs {
    const double* vM; const double* vD; const double* vK;
    double cM, cD, cK;
    PlaneMap mM, mD, mK;
};

// sys[p][r][c] = cM M + cD D + cK K, constrained rows/columns replaced by the identity.
__global__ void __launch_bounds__(256) form_system_kernel(const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ colidx, int64_t n_scalar_rows,
                                                          int64_t P, FormArgs F, const uint8_t* __restrict__ dir_mask,
                                                          double* __restrict__ sys) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_scalar_rows) return;
    const int32_t i = (int32_t)(row >> 2);
    const int r = (int)(row & 3);
    const bool row_fixed = dir_mask && dir_mask[row];
    const int32_t s = rowptr[i], t = rowptr[i + 1];
    for (int32_t p = s; p < t; ++p) {
        const int32_t j = colidx[p];
        double out[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double v = 0.0;
            const int pM = F.mM.m[4 * r + c], pD = F.mD.m[4 * r + c], pK = F.mK.m[4 * r + c];
            if (F.vM && pM >= 0) v += F.cM * F.vM[pM * P + p];
            if (F.vD && pD >= 0) v += F.cD * F.vD[pD * P + p];
            if (F.vK && pK >= 0) v += F.cK * F.vK[pK * P + p];
            const bool col_fixed = dir_mask && dir_mask[4 * (int64_t)j + c];
            if (row_fixed || col_fixed) v = (row_fixed && i == j && r == c) ? 1.0 : 0.0;
            out[c] = v;
        }
        double4* dst = reinterpret_cast<double4*>(sys + 16 * (int64_t)p + 4 * r);
        // two 16-byte stores (double4 is 16-byte aligned in CUDA 12.x)
        reinterpret_cast<double2*>(dst)[0] = make_double2(out[0], out[1]);
        reinterpret_cast<double2*>(dst)[1] = make_double2(out[2], out[3]);
    }
}

// 4x4 inverse by Gauss-Jordan with partial pivoting; returns false when singular.
__device__ __forceinline__ bool invert4(const double* __restrict__ a, double* __restrict__ inv) {
    double m[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            m[r][c] = a[4 * r + c];
            m[r][4 + c] = (r == c) ? 1.0 : 0.0;
        }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        double best = fabs(m[k][k]);
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const double v = fabs(m[r][k]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) ok = false;
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            if (r == piv) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const double tmp = m[k][c]; m[k][c] = m[r][c]; m[r][c] = tmp; }
            }
        }
        const double d = 1.0 / m[k][k];
#pragma unroll
        for (int c = 0; c < 8; ++c) m[k][c] *= d;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r != k) {
                const double f = m[r][k];
#pragma unroll
                for (int c = 0; c < 8; ++c) m[r][c] -= f * m[k][c];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) inv[4 * r + c] = m[r][4 + c];
    return ok;
}

__global__ void __launch_bounds__(128) block_invert_kernel(const int32_t* __restrict__ diag_idx, int32_t n_rows,
                                                           const double* __restrict__ sys, double* __restrict__ dinv,
                                                           int32_t* __restrict__ singular) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double a[16], inv[16];
    const double* blk = sys + 16 * (int64_t)diag_idx[i];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = blk[k];
    if (!invert4(a, inv)) atomicMin(singular, i);
#pragma unroll
    for (int k = 0; k < 16; ++k) dinv[16 * (int64_t)i + k] = inv[k];
}

// sys[p] <- dinv[i] * sys[p] for every block of row i; one thread per node row.
__global__ void __launch_bounds__(128) block_scale_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows,
                                                          const double* __restrict__ dinv, double* __restrict__ sys) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = dinv[16 * (int64_t)i + k];
    const int32_t s = rowptr[i], t = rowptr[i + 1];
    for (int32_t p = s; p < t; ++p) {
        double a[16], o[16];
        double* blk = sys + 16 * (int64_t)p;
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = blk[k];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                o[4 * r + c] = d[4 * r] * a[c] + d[4 * r + 1] * a[4 + c] + d[4 * r + 2] * a[8 + c] + d[4 * r + 3] * a[12 + c];
#pragma unroll
        for (int k = 0; k < 16; ++k) blk[k] = o[k];
    }
}

__global__ void __launch_bounds__(256) block_apply_kernel(const double* __restrict__ dinv, int64_t n_scalar_rows,
                                                          const double* __restrict__ rin, double* __restrict__ z,
                                                          const uint8_t* __restrict__ mask) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_scalar_rows) return;
    if (mask && mask[row]) { z[row] = 0.0; return; }
    const int64_t i = row >> 2;
    const int r = (int)(row & 3);
    const double* d = dinv + 16 * i + 4 * r;
    const double* v = rin + 4 * i;
    z[row] = d[0] * v[0] + d[1] * v[1] + d[2] * v[2] + d[3] * v[3];
}

static PlaneMap make_map_K() {
    PlaneMap pm;
    for (int k = 0; k < 16; ++k) pm.m[k] = -1;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) pm.m[4 * r + c] = (int8_t)(3 * r + c);
        pm.m[4 * r + 3] = (int8_t)(9 + r);
    }
    pm.m[15] = 12;
    return pm;
}
static PlaneMap make_map_M() {
    PlaneMap pm;
    for (int k = 0; k < 16; ++k) pm.m[k] = -1;
    for (int r = 0; r < 3; ++r) pm.m[5 * r] = (int8_t)r;
    return pm;
}
static PlaneMap make_map_D(int d_layout) {
    PlaneMap pm;
    for (int k = 0; k < 16; ++k) pm.m[k] = -1;
    for (int c = 0; c < 3; ++c) pm.m[12 + c] = (int8_t)c;
    pm.m[15] = 3;
    if (d_layout == 7)
        for (int r = 0; r < 3; ++r) pm.m[5 * r] = (int8_t)(4 + r);
    if (d_layout == 13)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) pm.m[4 * r + c] = (int8_t)(4 + 3 * r + c);
    return pm;
}

}  // namespace tefem

using namespace tefem;

extern "C" int tefem_planes_spmv(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int64_t n_blocks,
                                 const double* vals, const int8_t* h_plane_map, const double* x, double alpha,
                                 double beta_y, double* y, void* stream) {
    TEFEM_REQUIRE(rowptr && colidx && vals && h_plane_map && x && y, "null pointer");
    if (n_rows == 0) return TEFEM_OK;
    PlaneMap pm;
    memcpy(pm.m, h_plane_map, 16);
    const int64_t n = 4 * (int64_t)n_rows;
    planes_spmv_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(rowptr, colidx, n, n_blocks, vals, pm,
                                                                                   x, alpha, beta_y, y);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_form_system(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int64_t n_blocks,
                                 const double* vals_M, double cM, const double* vals_D, int d_layout, double cD,
                                 const double* vals_K, double cK, const uint8_t* dir_mask, double* sys, void* stream) {
    TEFEM_REQUIRE(rowptr && colidx && sys, "null pointer");
    TEFEM_REQUIRE(!vals_D || d_layout == 4 || d_layout == 7 || d_layout == 13, "d_layout must be 4, 7 or 13");
    TEFEM_REQUIRE((((uintptr_t)sys) & 127) == 0, "sys must be 128-byte aligned");
    if (n_rows == 0) return TEFEM_OK;
    FormArgs F;
    F.vM = vals_M; F.vD = vals_D; F.vK = vals_K;
    F.cM = cM; F.cD = cD; F.cK = cK;
    F.mM = make_map_M(); F.mD = make_map_D(vals_D ? d_layout : 4); F.mK = make_map_K();
    const int64_t n = 4 * (int64_t)n_rows;
    form_system_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(rowptr, colidx, n, n_blocks, F,
                                                                                   dir_mask, sys);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}

extern "C" int tefem_block_jacobi_scale(const int32_t* rowptr, const int32_t* diag_idx, int32_t n_rows, double* sys,
                                        double* dinv, int32_t* singular, int scale, void* stream) {
    TEFEM_REQUIRE(rowptr && diag_idx && sys && dinv && singular, "null pointer");
    if (n_rows == 0) return TEFEM_OK;
    cudaStream_t st = as_stream(stream);
    const int32_t big = INT32_MAX;
    TEFEM_CUDA_CHECK(cudaMemcpyAsync(singular, &big, sizeof(big), cudaMemcpyHostToDevice, st));
    block_invert_kernel<<<(unsigned)ceil_div(n_rows, 128), 128, 0, st>>>(diag_idx, n_rows, sys, dinv, singular);
    TEFEM_LAUNCH_CHECK();
    if (scale) {
        block_scale_kernel<<<(unsigned)ceil_div(n_rows, 128), 128, 0, st>>>(rowptr, n_rows, dinv, sys);
        TEFEM_LAUNCH_CHECK();
    }
    int32_t h_sing = INT32_MAX;
    TEFEM_CUDA_CHECK(cudaMemcpyAsync(&h_sing, singular, sizeof(h_sing), cudaMemcpyDeviceToHost, st));
    TEFEM_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_sing != INT32_MAX) {
        set_error("singular 4x4 diagonal block at node row %d", h_sing);
        return TEFEM_ERR_BREAKDOWN;
    }
    return TEFEM_OK;
}

extern "C" int tefem_block_jacobi_apply(const double* dinv, int32_t n_rows, const double* r, double* z,
                                        const uint8_t* dir_mask, void* stream) {
    TEFEM_REQUIRE(dinv && r && z, "null pointer");
    if (n_rows == 0) return TEFEM_OK;
    const int64_t n = 4 * (int64_t)n_rows;
    block_apply_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(dinv, n, r, z, dir_mask);
    TEFEM_LAUNCH_CHECK();
    return TEFEM_OK;
}
