// Tet4 thermoelastic element arithmetic shared by the assembly kernel (K1) and the
// per-element matrix kernel.  Closed forms of the reference's operator chains
// (rcapillon/thermoelasticity_fem elements.py:199-286), SURVEY.md 8(a) rows A3-A6:
//
//   Jt          rows X1-X0, X2-X0, X3-X0                           elements.py:200-204
//   det, inv    det(Jt), inv(Jt) (cofactors instead of LAPACK LU)   elements.py:205,212
//   G[i][b]     = dN_b/dx_i = (inv @ Dt)[i][b]: G[:,0] = -(row sums of inv), G[:,b] = inv[:,b-1]
//   GT[c][b]    = (inv^T @ Dt)[c][b]: the divergence of elements.py:237 misses mat_P, i.e. it uses
//                 the TRANSPOSED inverse; GT[:,0] = -(column sums of inv), GT[c][b] = inv[b-1][c]
//   Kuu[3a+i,3b+j] = W4 det (lam G[i][a] G[j][b] + mu G[j][a] G[i][b] + mu d_ij g_a.g_b)   :254-263
//   Kut[3a+i,b]    = -beta det S_a G[i][b]                                                  :265-275
//   Ktt[a,b]       = W4 det k g_a.g_b                                                       :277-286
//   Muu[3a+i,3b+j] = rho det NN[a][b] d_ij                                                  :220-229
//   Dtt[a,b]       = rho c det NN[a][b]                                                     :243-252
//   Dtu[a,3b+c]    = -beta T0 det S_a GT[c][b]                                              :231-241
#pragma once
#include "tefem_common.cuh"

namespace tefem {

// Staged per-element record (doubles).  The odd stride keeps 8-byte shared-memory accesses of
// consecutive records on distinct banks.
enum {
    REC_G = 0,     // 12: G column-major, G[i][a] at REC_G + 3*a + i
    REC_GT0 = 12,  //  3: GT[c][0]
    REC_CL = 15,   // W4 det lambda
    REC_CM = 16,   // W4 det mu
    REC_CK = 17,   // W4 det k
    REC_CB = 18,   // -beta det
    REC_RM = 19,   // rho det
    REC_RC = 20,   // rho c det
    REC_CD = 21,   // -beta T0 det
    REC_STRIDE = 23
};

// Material table row (materials.py:5-29): rho, lambda, mu, k, c, beta, T0, pad.
enum { MAT_RHO = 0, MAT_LAM = 1, MAT_MU = 2, MAT_K = 3, MAT_C = 4, MAT_BETA = 5, MAT_T0 = 6, MAT_STRIDE = 8 };

// Geometry + material scalars of one element.  Returns det(Jt).
TEFEM_HD double element_record(const double* __restrict__ X0, const double* __restrict__ X1,
                               const double* __restrict__ X2, const double* __restrict__ X3,
                               const double* __restrict__ mat, double W4, double* __restrict__ rec) {
    // Jt[a][b] = d x_b / d xi_a
    const double j00 = X1[0] - X0[0], j01 = X1[1] - X0[1], j02 = X1[2] - X0[2];
    const double j10 = X2[0] - X0[0], j11 = X2[1] - X0[1], j12 = X2[2] - X0[2];
    const double j20 = X3[0] - X0[0], j21 = X3[1] - X0[1], j22 = X3[2] - X0[2];
    const double c00 = j11 * j22 - j12 * j21;
    const double c01 = j12 * j20 - j10 * j22;
    const double c02 = j10 * j21 - j11 * j20;
    const double det = j00 * c00 + j01 * c01 + j02 * c02;
    const double r = 1.0 / det;
    // inv[i][k] = cofactor[k][i] / det
    const double i00 = c00 * r, i10 = c01 * r, i20 = c02 * r;
    const double i01 = (j02 * j21 - j01 * j22) * r, i11 = (j00 * j22 - j02 * j20) * r, i21 = (j01 * j20 - j00 * j21) * r;
    const double i02 = (j01 * j12 - j02 * j11) * r, i12 = (j02 * j10 - j00 * j12) * r, i22 = (j00 * j11 - j01 * j10) * r;
    // G[:,0] = -(inv[:,0] + inv[:,1] + inv[:,2]) ; G[:,b] = inv[:,b-1]
    rec[REC_G + 0] = -(i00 + i01 + i02);
    rec[REC_G + 1] = -(i10 + i11 + i12);
    rec[REC_G + 2] = -(i20 + i21 + i22);
    rec[REC_G + 3] = i00; rec[REC_G + 4] = i10; rec[REC_G + 5] = i20;
    rec[REC_G + 6] = i01; rec[REC_G + 7] = i11; rec[REC_G + 8] = i21;
    rec[REC_G + 9] = i02; rec[REC_G + 10] = i12; rec[REC_G + 11] = i22;
    // GT[c][0] = -(inv[0][c] + inv[1][c] + inv[2][c])
    rec[REC_GT0 + 0] = -(i00 + i10 + i20);
    rec[REC_GT0 + 1] = -(i01 + i11 + i21);
    rec[REC_GT0 + 2] = -(i02 + i12 + i22);
    const double wd = W4 * det;
    rec[REC_CL] = wd * mat[MAT_LAM];
    rec[REC_CM] = wd * mat[MAT_MU];
    rec[REC_CK] = wd * mat[MAT_K];
    rec[REC_CB] = -(mat[MAT_BETA] * det);
    rec[REC_RM] = mat[MAT_RHO] * det;
    rec[REC_RC] = mat[MAT_RHO] * mat[MAT_C] * det;
    rec[REC_CD] = -(mat[MAT_BETA] * mat[MAT_T0] * det);
    return det;
}

// GT[c][b] from the record: b == 0 -> REC_GT0 + c ; b >= 1 -> inv[b-1][c] = G[b-1][c+1] = rec[3*(c+1) + (b-1)]
TEFEM_HD double rec_gt(const double* rec, int c, int b) {
    return b == 0 ? rec[REC_GT0 + c] : rec[REC_G + 3 * (c + 1) + (b - 1)];
}

// Accumulators of one node-pair block.  WITH_KUU/… select which parts are live so that the
// compiler drops the dead ones per kernel instantiation.
template <bool WITH_K, bool WITH_M, bool WITH_D>
struct BlockAcc {
    double kuu[WITH_K ? 9 : 1];
    double kut[WITH_K ? 3 : 1];
    double ktt;
    double m;        // rho det NN (one value: Muu and its three diagonal copies)
    double dtu[WITH_D ? 3 : 1];
    double dtt;
    TEFEM_HD void clear() {
        for (int i = 0; i < (WITH_K ? 9 : 1); ++i) kuu[i] = 0.0;
        for (int i = 0; i < (WITH_K ? 3 : 1); ++i) kut[i] = 0.0;
        for (int i = 0; i < (WITH_D ? 3 : 1); ++i) dtu[i] = 0.0;
        ktt = 0.0; m = 0.0; dtt = 0.0;
    }
    // number of live accumulators, and (de)serialisation for the partial sums of split incidence lists
    static constexpr int LIVE = (WITH_K ? 13 : 0) + (WITH_M ? 1 : 0) + (WITH_D ? 4 : 0);
    TEFEM_HD void dump(double* dst) const {
        int k = 0;
        if (WITH_K) {
            for (int i = 0; i < 9; ++i) dst[k++] = kuu[i];
            for (int i = 0; i < 3; ++i) dst[k++] = kut[i];
            dst[k++] = ktt;
        }
        if (WITH_M) dst[k++] = m;
        if (WITH_D) {
            for (int i = 0; i < 3; ++i) dst[k++] = dtu[i];
            dst[k++] = dtt;
        }
    }
    TEFEM_HD void add_from(const double* src) {
        int k = 0;
        if (WITH_K) {
            for (int i = 0; i < 9; ++i) kuu[i] += src[k++];
            for (int i = 0; i < 3; ++i) kut[i] += src[k++];
            ktt += src[k++];
        }
        if (WITH_M) m += src[k++];
        if (WITH_D) {
            for (int i = 0; i < 3; ++i) dtu[i] += src[k++];
            dtt += src[k++];
        }
    }
};

// acc += contribution of block (a, b) of the element whose record is `rec`.
template <bool WITH_K, bool WITH_M, bool WITH_D>
TEFEM_HD void block_accumulate(BlockAcc<WITH_K, WITH_M, WITH_D>& acc, const double* __restrict__ rec,
                               int a, int b, const double* __restrict__ S, const double* __restrict__ NN) {
    if (WITH_K) {
        const double ga0 = rec[REC_G + 3 * a], ga1 = rec[REC_G + 3 * a + 1], ga2 = rec[REC_G + 3 * a + 2];
        const double gb0 = rec[REC_G + 3 * b], gb1 = rec[REC_G + 3 * b + 1], gb2 = rec[REC_G + 3 * b + 2];
        const double cl = rec[REC_CL], cm = rec[REC_CM];
        const double dot = ga0 * gb0 + ga1 * gb1 + ga2 * gb2;
        const double la0 = cl * ga0, la1 = cl * ga1, la2 = cl * ga2;
        const double ma0 = cm * ga0, ma1 = cm * ga1, ma2 = cm * ga2;
        const double md = cm * dot;
        // kuu[i][j] += lam ga_i gb_j + mu ga_j gb_i (+ mu dot on the diagonal)
        acc.kuu[0] += la0 * gb0 + ma0 * gb0 + md;
        acc.kuu[1] += la0 * gb1 + ma1 * gb0;
        acc.kuu[2] += la0 * gb2 + ma2 * gb0;
        acc.kuu[3] += la1 * gb0 + ma0 * gb1;
        acc.kuu[4] += la1 * gb1 + ma1 * gb1 + md;
        acc.kuu[5] += la1 * gb2 + ma2 * gb1;
        acc.kuu[6] += la2 * gb0 + ma0 * gb2;
        acc.kuu[7] += la2 * gb1 + ma1 * gb2;
        acc.kuu[8] += la2 * gb2 + ma2 * gb2 + md;
        const double t = rec[REC_CB] * S[a];
        acc.kut[0] += t * gb0;
        acc.kut[1] += t * gb1;
        acc.kut[2] += t * gb2;
        acc.ktt += rec[REC_CK] * dot;
    }
    if (WITH_M) acc.m += rec[REC_RM] * NN[4 * a + b];
    if (WITH_D) {
        const double t = rec[REC_CD] * S[a];
        acc.dtu[0] += t * rec_gt(rec, 0, b);
        acc.dtu[1] += t * rec_gt(rec, 1, b);
        acc.dtu[2] += t * rec_gt(rec, 2, b);
        acc.dtt += rec[REC_RC] * NN[4 * a + b];
    }
}

// The non-symmetric parts of the MIRROR block (j, i) of an upper block (i, j).  Its symmetric parts are those of
// (i, j): Kuu_ji = Kuu_ij^T, Ktt, Muu, Dtt equal (NN and g_a.g_b are symmetric bit for bit); only
//   Kut_ji[i] = -beta det S_b G[i][a]      and      Dtu_ji[c] = -beta T0 det S_b GT[c][a]
// differ, and they need nothing beyond what (i, j) already loads except GT[:, a].
template <bool WITH_K, bool WITH_D>
struct MirrorAcc {
    double kut[WITH_K ? 3 : 1];
    double dtu[WITH_D ? 3 : 1];
    TEFEM_HD void clear() {
        for (int i = 0; i < (WITH_K ? 3 : 1); ++i) kut[i] = 0.0;
        for (int i = 0; i < (WITH_D ? 3 : 1); ++i) dtu[i] = 0.0;
    }
};

// acc += block (a, b) and mir += the non-symmetric parts of block (b, a) of the same element.
template <bool WITH_K, bool WITH_M, bool WITH_D>
TEFEM_HD void pair_accumulate(BlockAcc<WITH_K, WITH_M, WITH_D>& acc, MirrorAcc<WITH_K, WITH_D>& mir,
                              const double* __restrict__ rec, int a, int b, const double* __restrict__ S,
                              const double* __restrict__ NN) {
    if (WITH_K) {
        const double ga0 = rec[REC_G + 3 * a], ga1 = rec[REC_G + 3 * a + 1], ga2 = rec[REC_G + 3 * a + 2];
        const double gb0 = rec[REC_G + 3 * b], gb1 = rec[REC_G + 3 * b + 1], gb2 = rec[REC_G + 3 * b + 2];
        const double cl = rec[REC_CL], cm = rec[REC_CM];
        const double dot = ga0 * gb0 + ga1 * gb1 + ga2 * gb2;
        const double la0 = cl * ga0, la1 = cl * ga1, la2 = cl * ga2;
        const double ma0 = cm * ga0, ma1 = cm * ga1, ma2 = cm * ga2;
        const double md = cm * dot;
        acc.kuu[0] += la0 * gb0 + ma0 * gb0 + md;
        acc.kuu[1] += la0 * gb1 + ma1 * gb0;
        acc.kuu[2] += la0 * gb2 + ma2 * gb0;
        acc.kuu[3] += la1 * gb0 + ma0 * gb1;
        acc.kuu[4] += la1 * gb1 + ma1 * gb1 + md;
        acc.kuu[5] += la1 * gb2 + ma2 * gb1;
        acc.kuu[6] += la2 * gb0 + ma0 * gb2;
        acc.kuu[7] += la2 * gb1 + ma1 * gb2;
        acc.kuu[8] += la2 * gb2 + ma2 * gb2 + md;
        const double cb = rec[REC_CB];
        const double t = cb * S[a], u = cb * S[b];
        acc.kut[0] += t * gb0;
        acc.kut[1] += t * gb1;
        acc.kut[2] += t * gb2;
        mir.kut[0] += u * ga0;
        mir.kut[1] += u * ga1;
        mir.kut[2] += u * ga2;
        acc.ktt += rec[REC_CK] * dot;
    }
    if (WITH_M) acc.m += rec[REC_RM] * NN[4 * a + b];
    if (WITH_D) {
        const double cd = rec[REC_CD];
        const double t = cd * S[a], u = cd * S[b];
        acc.dtu[0] += t * rec_gt(rec, 0, b);
        acc.dtu[1] += t * rec_gt(rec, 1, b);
        acc.dtu[2] += t * rec_gt(rec, 2, b);
        mir.dtu[0] += u * rec_gt(rec, 0, a);
        mir.dtu[1] += u * rec_gt(rec, 1, a);
        mir.dtu[2] += u * rec_gt(rec, 2, a);
        acc.dtt += rec[REC_RC] * NN[4 * a + b];
    }
}

}  // namespace tefem
