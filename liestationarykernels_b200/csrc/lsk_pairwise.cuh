// Parameter block shared by the host API and the pairwise kernels.
#pragma once
#include <cstdint>

#define LSK_MAX_FORMS 5
#define LSK_MAX_VARS 5
#define LSK_MAX_COEF 400
#define LSK_MAX_ROWS 24
#define LSK_MAX_SNAPSHOT 7
#define LSK_STAIR_MAX_ROWS 12    // STAIR2: rows sit at a fixed stride of LSK_STAIR_STRIDE coefficients
#define LSK_STAIR_STRIDE 16

// Epilogue kinds: how the class function is evaluated from the conjugation invariants.
#define LSK_EPI_CHEB1   1   // rank 1:  sum_k coef[k] T_k(var0)                          (Clenshaw; SO(3), SU(2))
#define LSK_EPI_STAIR2  2   // rank 2:  sum_b var1^b sum_a C[b][a] var0^a                (nested Horner; SO(5) fast path)
#define LSK_EPI_ORBIT2  3   // rank 2:  sum_ab W[a][b] T_a(c1) T_b(c2), c1+c2=var0, c1*c2=var1 (Weyl-orbit form; robust path)
#define LSK_EPI_SPARSE  4   // any:     sum_t coef[2t] prod_v var_v^{e_tv}, exponents packed 8 bits each in coef[2t+1]
#define LSK_EPI_LINEAR  5   // out = scale * form_0                                   (Gram matrix of two feature maps)
#define LSK_EPI_FEAT_CHEB1  6   // per-irrep outputs, rank 1:  out[i, l*stride + j] = sum_k C_l[k] T_k(var0), row l has row_len[l] coefficients
#define LSK_EPI_FEAT_SPARSE 7   // per-irrep outputs, any rank: row l is a sparse polynomial with row_len[l] terms (SPARSE encoding)

#define LSK_EPI_HORNER      8   // any:     nested Horner over a tape of (coefficient, op) steps (lsk_pairwise.cu epi_tape)
#define LSK_EPI_FEAT_HORNER 9   // per-irrep outputs, any rank: row l is a tape of row_len[l] steps
#define LSK_EPI_SYRK       10   // out += scale * form_0 on the tiles that reach the diagonal or lie above it, nothing mirrored
                                // (trailing update of the blocked Cholesky, lsk_cholesky.cu; square row-major output)

#define LSK_OUT_ROW_MAJOR   0   // out[i * ld_out + j]
#define LSK_OUT_PANEL_MAJOR 1   // operand layout of this kernel: [i / 64][j / 4][i % 64][j % 4], out_kg k-groups per row

// One pairwise problem: `nforms` bilinear forms u_f(i,j) = sum_{k in groups [g0_f, g1_f)} A[i,k] B[j,k], then
//   var_v = aff[v][0] + sum_f aff[v][1+f] * u_f,   out[i,j] = scale * P(var_0..var_{nvars-1}).
// Passed by value as a kernel parameter, so the coefficient table lives in the constant bank and is read through
// uniform loads (LDCU) - no device allocation, no host->device copy, no global state.
struct LskEpilogue {
    int32_t kind;
    int32_t nforms;
    int32_t nvars;
    int32_t nterms;                        // CHEB1: #coefficients; STAIR2: #rows; ORBIT2: matrix size D; SPARSE: #terms
    int32_t group_begin[LSK_MAX_FORMS];    // k-group (4 doubles) range of each form inside the embedded operands
    int32_t group_end[LSK_MAX_FORMS];
    int32_t row_len[LSK_MAX_ROWS];         // STAIR2: #coefficients in row b;  FEAT_*: #coefficients / #terms of irrep l
    int32_t out_layout;                    // LSK_OUT_*
    int32_t out_kg;                        // PANEL_MAJOR: k-groups per row of the output operand
    int32_t feat_stride;                   // FEAT_*: column offset between consecutive irreps (= number of phases)
    int32_t reserved;
    double scale;
    unsigned long long stair_shape;        // STAIR2: coefficient pairs of row slot k (evaluation order) in bits [4k, 4k+4)
    double aff[LSK_MAX_VARS][1 + LSK_MAX_FORMS];
    double coef[LSK_MAX_COEF];
};
