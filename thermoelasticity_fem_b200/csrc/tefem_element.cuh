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
// Every form is (material constant) x (geometric factor): the geometric factors are summed per work item and the
// constants applied once (block_finalize).
#pragma once
#include "tefem_common.cuh"

namespace tefem {

// Staged per-element record (doubles): geometry only.  Work items are material-uniform (the symbolic phase cuts
// a block's incidence list where the material changes), so the material constants multiply the accumulated
// geometric sums ONCE per item instead of once per incidence: 7 (K) / 10 (M + D + K) shared-memory loads per
// incidence instead of 11 / 17.  The odd strides keep 8-byte accesses of consecutive records on distinct banks.
enum {
    REC_G = 0,      // 12: G column-major, G[i][a] at REC_G + 3*a + i
    REC_DET = 12,   // det(Jt)
    REC_GT0 = 13,   //  3: GT[c][0] (D only)
    REC_STRIDE_K = 13,
    REC_STRIDE_D = 17,
    REC_STRIDE_MAX = 17
};

// Material table row (materials.py:5-29): rho, lambda, mu, k, c, beta, T0, pad.
enum { MAT_RHO = 0, MAT_LAM = 1, MAT_MU = 2, MAT_K = 3, MAT_C = 4, MAT_BETA = 5, MAT_T0 = 6, MAT_STRIDE = 8 };

// Geometry of one element.  Returns det(Jt).
template <bool WITH_GT>
TEFEM_HD double element_record(const double* __restrict__ X0, const double* __restrict__ X1,
                               const double* __restrict__ X2, const double* __restrict__ X3, double* __restrict__ rec) {
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
    rec[REC_DET] = det;
    if (WITH_GT) {   // GT[c][0] = -(inv[0][c] + inv[1][c] + inv[2][c])
        rec[REC_GT0 + 0] = -(i00 + i10 + i20);
        rec[REC_GT0 + 1] = -(i01 + i11 + i21);
        rec[REC_GT0 + 2] = -(i02 + i12 + i22);
    }
    return det;
}

// GT[c][b] from the record: b == 0 -> REC_GT0 + c ; b >= 1 -> inv[b-1][c] = G[b-1][c+1] = rec[3*(c+1) + (b-1)]
TEFEM_HD double rec_gt(const double* rec, int c, int b) {
    return b == 0 ? rec[REC_GT0 + c] : rec[REC_G + 3 * (c + 1) + (b - 1)];
}

// Quadrature tables S_a = sum_g w N_a(g) and NN_ab = sum_g w N_a(g) N_b(g) as seen by the incidence loop.
// Default (ASM_TAB_SELECT, K1 is bound by the shared-memory pipe, profiles/README.md): the tables take only 2 and 5
// BITWISE distinct values (S: a == 0 or not; NN: (0,0), a == b >= 1, a != b both >= 1, {0, 1|2}, {0, 3} - the last two
// differ in the last bit because the reference sums the Gauss points in a fixed order), so the entry is selected with
// compares on the local node numbers from values held in registers / the constant bank: no shared-memory load.
// -DASM_TAB_LOAD restores the indexed loads (one wavefront per half-warp and table; the round-1 kernel).
// Both forms return the table entries themselves: the assembled values are bitwise identical.
#ifndef ASM_TAB_LOAD
#define ASM_TAB_SELECT 1
#endif
struct QuadS {
#ifdef ASM_TAB_SELECT
    double s0, s1;
    TEFEM_HD double at(int a) const { return a == 0 ? s0 : s1; }
#else
    const double* s;
    TEFEM_HD double at(int a) const { return s[a]; }
#endif
};
struct QuadNN {
#ifdef ASM_TAB_SELECT
    double v00, vdd, vod, v0x, v03;
    TEFEM_HD double at(int a, int b) const {
        const double diag = a == 0 ? v00 : vdd;
        const double with0 = (a | b) == 3 ? v03 : v0x;      // one of a, b is 0 here: a | b is the other one
        const double off = (a == 0 || b == 0) ? with0 : vod;
        return a == b ? diag : off;
    }
#else
    const double* nn;
    TEFEM_HD double at(int a, int b) const { return nn[4 * a + b]; }
#endif
};
// True when the select form reproduces every entry of the tables bit for bit (checked once on the host before the
// first launch; the tables are built in the reference's operation order, tefem_assemble.cu make_quad_tables).
TEFEM_HD QuadS make_quad_s(const double* S) {
    QuadS q;
#ifdef ASM_TAB_SELECT
    q.s0 = S[0]; q.s1 = S[1];
#else
    q.s = S;
#endif
    return q;
}
TEFEM_HD QuadNN make_quad_nn(const double* NN) {
    QuadNN q;
#ifdef ASM_TAB_SELECT
    q.v00 = NN[0]; q.vdd = NN[5]; q.vod = NN[6]; q.v0x = NN[1]; q.v03 = NN[3];
#else
    q.nn = NN;
#endif
    return q;
}
inline bool quad_select_exact(const double* S, const double* NN) {
    const QuadS s = make_quad_s(S);
    const QuadNN n = make_quad_nn(NN);
    for (int a = 0; a < 4; ++a) {
        if (s.at(a) != S[a]) return false;
        for (int b = 0; b < 4; ++b)
            if (n.at(a, b) != NN[4 * a + b]) return false;
    }
    return true;
}

// Values of one node-pair block.  WITH_K/... select which parts are live so that the compiler drops the dead
// ones per kernel instantiation.
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
    // number of live values, and (de)serialisation for the partial sums of split incidence lists
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

// Material-free sums of one work item over its incidence entries (element e, local nodes a, b):
//   P[3i+j] = sum W4 det G[i][a] G[j][b]      q[i] = sum det S_a G[i][b]
//   nn      = sum det NN[a][b]                r[c] = sum det S_a GT[c][b]
template <bool WITH_K, bool WITH_M, bool WITH_D>
struct GeomAcc {
    double P[WITH_K ? 9 : 1];
    double q[WITH_K ? 3 : 1];
    double nn;
    double r[WITH_D ? 3 : 1];
    TEFEM_HD void clear() {
        for (int i = 0; i < (WITH_K ? 9 : 1); ++i) P[i] = 0.0;
        for (int i = 0; i < (WITH_K ? 3 : 1); ++i) q[i] = 0.0;
        for (int i = 0; i < (WITH_D ? 3 : 1); ++i) r[i] = 0.0;
        nn = 0.0;
    }
};

// acc += geometric contribution of block (a, b) of the element whose record is `rec`.
template <bool WITH_K, bool WITH_M, bool WITH_D>
TEFEM_HD void block_accumulate(GeomAcc<WITH_K, WITH_M, WITH_D>& acc, const double* __restrict__ rec, int a, int b,
                               double W4, const QuadS& S, const QuadNN& NN) {
    const double det = rec[REC_DET];
    double ds = 0.0;
    if (WITH_K || WITH_D) ds = det * S.at(a);
    if (WITH_K) {
        const double ga0 = rec[REC_G + 3 * a], ga1 = rec[REC_G + 3 * a + 1], ga2 = rec[REC_G + 3 * a + 2];
        double gb0 = ga0, gb1 = ga1, gb2 = ga2;
        if (a != b) {   // diagonal incidences (all lanes of a warp of diagonal chunks) skip three shared-memory loads
            gb0 = rec[REC_G + 3 * b]; gb1 = rec[REC_G + 3 * b + 1]; gb2 = rec[REC_G + 3 * b + 2];
        }
        const double wd = W4 * det;
        const double w0 = wd * ga0, w1 = wd * ga1, w2 = wd * ga2;
        acc.P[0] += w0 * gb0; acc.P[1] += w0 * gb1; acc.P[2] += w0 * gb2;
        acc.P[3] += w1 * gb0; acc.P[4] += w1 * gb1; acc.P[5] += w1 * gb2;
        acc.P[6] += w2 * gb0; acc.P[7] += w2 * gb1; acc.P[8] += w2 * gb2;
        acc.q[0] += ds * gb0;
        acc.q[1] += ds * gb1;
        acc.q[2] += ds * gb2;
    }
    if (WITH_M || WITH_D) acc.nn += det * NN.at(a, b);
    if (WITH_D) {
        acc.r[0] += ds * rec_gt(rec, 0, b);
        acc.r[1] += ds * rec_gt(rec, 1, b);
        acc.r[2] += ds * rec_gt(rec, 2, b);
    }
}

// The block values of a material-uniform item: out = material constants x geometric sums
// (Kuu/Kut/Ktt elements.py:254-286, Muu :220-229, Dtu :231-241, Dtt :243-252).
template <bool WITH_K, bool WITH_M, bool WITH_D>
TEFEM_HD void block_finalize(const GeomAcc<WITH_K, WITH_M, WITH_D>& g, const double* __restrict__ mat,
                             BlockAcc<WITH_K, WITH_M, WITH_D>& out) {
    out.clear();
    if (WITH_K) {
        const double lam = mat[MAT_LAM], mu = mat[MAT_MU];
        const double md = mu * (g.P[0] + g.P[4] + g.P[8]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                out.kuu[3 * i + j] = lam * g.P[3 * i + j] + mu * g.P[3 * j + i] + (i == j ? md : 0.0);
        const double nb = -mat[MAT_BETA];
        out.kut[0] = nb * g.q[0]; out.kut[1] = nb * g.q[1]; out.kut[2] = nb * g.q[2];
        out.ktt = mat[MAT_K] * (g.P[0] + g.P[4] + g.P[8]);
    }
    if (WITH_M || WITH_D) out.m = mat[MAT_RHO] * g.nn;
    if (WITH_D) {
        const double nbt = -(mat[MAT_BETA] * mat[MAT_T0]);
        out.dtu[0] = nbt * g.r[0]; out.dtu[1] = nbt * g.r[1]; out.dtu[2] = nbt * g.r[2];
        out.dtt = mat[MAT_RHO] * mat[MAT_C] * g.nn;
    }
}

}  // namespace tefem
