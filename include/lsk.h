/* liblsk — C ABI of the B200-native hot path of LieStationaryKernels.
 *
 * The reference (imbirik/LieStationaryKernels) is pure Python and has no FFI; the boundary a maintainer binds is
 * this header (ctypes stub: INTEGRATION.md).  Every entry point
 *   - is `extern "C"`, takes plain pointers and sizes, returns int: 0 ok, < 0 argument error (LSK_E*), > 0 cudaError_t;
 *   - takes DEVICE pointers for bulk data (contiguous, row-major float64 unless stated; complex128 = interleaved re,im),
 *     caller-owned, 16-byte aligned; small descriptors (`LskEpilogue`, int arrays) are HOST pointers read during the call;
 *   - launches asynchronously on the `cudaStream_t` passed as `void* stream` (NULL = default stream); the embedding and
 *     pairwise kernels are launched with programmatic stream serialisation and wait (griddepcontrol.wait) for the previous
 *     kernel of the stream before their first global-memory access, so ordinary stream semantics hold for the caller;
 *   - allocates nothing that outlives the call and keeps no mutable global state (safe from several host threads on
 *     distinct streams).
 * "reference" citations are paths below /root/reference/src/lie_stationary_kernels/.
 */
#ifndef LSK_H
#define LSK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSK_EARG         (-1)   /* invalid argument            */
#define LSK_ESIZE        (-2)   /* size out of supported range */
#define LSK_EUNSUPPORTED (-3)   /* configuration not compiled  */

#define LSK_MAX_FORMS 5
#define LSK_MAX_VARS  5
#define LSK_MAX_COEF  400
#define LSK_MAX_ROWS  24

/* epilogue kinds */
#define LSK_EPI_CHEB1       1   /* rank 1: sum_k coef[k] T_k(var0)  (Clenshaw)                                   */
#define LSK_EPI_STAIR2      2   /* rank 2: sum_b var1^b sum_a C[b][a] var0^a  (nested Horner, fast form)         */
#define LSK_EPI_ORBIT2      3   /* rank 2: sum_ab W[a][b] T_a(c1) T_b(c2), c1+c2 = var0, c1 c2 = var1 (robust)   */
#define LSK_EPI_SPARSE      4   /* any  : sum_t coef[2t] prod_v var_v^e_tv, e packed 8 bits each in coef[2t+1]
                                   (monomial list; read by lsk_homog_pairs only)                                  */
#define LSK_EPI_LINEAR      5   /* out = scale * form_0  (Gram of two feature maps)                              */
#define LSK_EPI_FEAT_CHEB1  6   /* per-irrep rows, rank 1 (row l: row_len[l] Chebyshev coefficients)             */
#define LSK_EPI_FEAT_SPARSE 7   /* per-irrep rows, any rank (row l: row_len[l] sparse terms; lsk_homog_pairs)    */
#define LSK_EPI_HORNER      8   /* any  : nested Horner tape, nterms steps (coef[2t], op = low word of coef[2t+1]):
                                   op 0: a0 = a0 v0 + c;  1: a0 = c;  2+3(l-1): al = al vl + a(l-1);  +1: al = a(l-1);
                                   +2: al = al vl;  result a(nvars-1)  (SU(3..5), SO(4), SO(6), SO(7))               */
#define LSK_EPI_FEAT_HORNER 9   /* per-irrep rows, any rank: row l is a tape of row_len[l] steps                  */
#define LSK_EPI_SYRK       10   /* out += scale * form_0 on the tiles that reach the diagonal or lie above it (square
                                   row-major output, nothing mirrored): trailing update of lsk_potrf_upper          */

/* output layouts */
#define LSK_OUT_ROW_MAJOR   0   /* out[i * ld_out + j]                                                            */
#define LSK_OUT_PANEL_MAJOR 1   /* [i / 64][j / 4][i % 64][j % 4]: the operand layout the pairwise kernel reads   */
#define LSK_FEAT_COMPLEX    2   /* complex128 [N, m] interleaved (feature kernels only)                           */

/* One pairwise problem.  nforms bilinear forms u_f(i,j) = sum_{k in 4*[group_begin_f, group_end_f)} A[i,k] B[j,k],
 *   var_v = aff[v][0] + sum_f aff[v][1+f] u_f,     out[i,j] = scale * P(var_0 .. var_{nvars-1}).
 * Passed by value to the kernel: the coefficient table is read from the constant bank. */
typedef struct LskEpilogue {
    int32_t kind;
    int32_t nforms;
    int32_t nvars;
    int32_t nterms;                      /* CHEB1 #coef; STAIR2 #rows; ORBIT2 matrix size; SPARSE #terms; HORNER #steps; FEAT_* #irreps */
    int32_t group_begin[LSK_MAX_FORMS];
    int32_t group_end[LSK_MAX_FORMS];
    int32_t row_len[LSK_MAX_ROWS];
    int32_t out_layout;
    int32_t out_kg;                      /* PANEL_MAJOR: k-groups (4 columns) per output row */
    int32_t feat_stride;                 /* FEAT_*: column offset between irreps (= number of phases) */
    int32_t reserved;                    /* bit 0 (lsk_homog_pairs): pair (y_j, x_i) instead of (x_i, y_j);
                                            bits 4-7 (lsk_homog_pairs, n = 5): t in {0, 2, 3}: the caller vouches that every
                                            subgroup sample is blockdiag(I_t, R) (Stiefel: H = SO(n - m), stiefel.py:25-36);
                                            11: t = 3 and every R is a plane rotation [[c, s], [-s, c]];
                                            bit 1 (lsk_pairwise_poly): symmetric problem (A, B describe the same points, square
                                            row-major output): only the upper triangle of tiles is computed and mirrored;
                                            bits 8+f (lsk_pairwise_poly): form f enters the affine map squared */
    double scale;
    unsigned long long stair_shape;      /* STAIR2: coefficient pairs of row slot k (evaluation order) in bits [4k, 4k+4);
                                            row slot k starts at coef[16 k] */
    double aff[LSK_MAX_VARS][1 + LSK_MAX_FORMS];
    double coef[LSK_MAX_COEF];
} LskEpilogue;

int lsk_version(void);

/* Stream-ordered snapshot of up to LSK_MAX_SNAPSHOT device doubles into host-mapped pinned memory: slot i (16 bytes, one store)
 *   dst[2 i] = *src[i],   dst[2 i + 1] (read as uint64) = bits(*src[i]) ^ (seq * 0x9E3779B97F4A7C15 + 0x7F4A7C159E3779B9).
 * One 32-thread kernel, no copy engine, no event, no system-scope fence: the host polls until every slot's tag matches its value
 * for the sequence number it passed.  The Python layer uses it to verify, without draining the stream, that the kernel
 * hyper-parameters (reference spectral_measure.py:20-21: device-resident nn.Parameters that the reference re-reads on every
 * call) still have the values the cached coefficient block was built from.
 * src: HOST array of n device pointers; dst_host_mapped: pinned host memory (cudaHostAlloc / torch pin_memory) of
 * 2 * LSK_MAX_SNAPSHOT doubles, 16-byte aligned, addressed by its host pointer (unified addressing). */
#define LSK_MAX_SNAPSHOT 7
int lsk_snapshot_doubles(const void* const* src, int n, void* dst_host_mapped, unsigned long long seq, void* stream);

/* lsk_embed_points2 (below) whose kernel also writes the snapshot (same slot format): the hyper-parameter check of a covariance call
 * then costs no launch of its own.  Both operands must be non-empty; n_snap = 0 makes it a plain lsk_embed_points2. */
int lsk_embed_points2_snapshot(const void* x, long long num_x, const int* part_x, void* out_x,
                               const void* y, long long num_y, const int* part_y, void* out_y,
                               int n, int is_complex, int nseg, const int* wedge, const int* group_begin, int kg_total,
                               const void* const* snap_src, int n_snap, void* snap_dst_host_mapped, unsigned long long seq,
                               void* stream);

/* Text for a return code of any entry point (0 ok, < 0 LSK_E*, > 0 cudaError_t).  Stateless: the library keeps no
 * "last error"; every call reports through its return value. */
const char* lsk_error_string(int code);

/* Number of doubles of the panel-major embedding buffer for `num` points and `kg_total` k-groups. */
long long lsk_embed_buffer_doubles(long long num, int kg_total);

/* Per-point compound-matrix embeddings (k = 1..3; k = 4 for real matrices) of SO(n)/SU(n) elements, panel-major output.
 * Replaces operand preparation of CompactLieGroup.pairwise_diff: inv + cartesian_prod + bmm (reference space.py:62-78,
 * utils.py:40-52, spaces/so.py:124-127, spaces/su.py:79-82).
 * x: [num, n, n] float64 (is_complex = 0) or complex128 (is_complex = 1).  wedge/part/group_begin: host int[nseg].
 * part: 0 real; 1 complex (re,im); 2 complex (re,im) [second operand, real form]; 3 complex (-im,re) [imaginary form].
 * parts 6..9 (complex): one real number per entry of Lambda^k x: re, im, re + im, re - im.  With A = (6, 7, 8) and B = (6, 7, 9)
 * the complex form <Lambda^k x, conj Lambda^k y> = (f0 + f1) + i (f2 - f0 + f1) costs three real forms instead of four.
 * part 5 (complex, n = 4, wedge = 2): Lambda^2 x in the real basis of Lambda^2 C^4 (the SO(6) image of SU(4)), 36 reals, both sides.
 * part 4 (real, n = 2 * wedge): Lambda^k x with the Hodge star on the row index, so that <*Lambda^r x, Lambda^r y> is the
 * invariant prod_i 2 sin(theta_i) (up to sign) that separates the two SO(2r) classes with equal spectrum.
 * wedge = 100 (n = 5, real): the 16 rotor coefficients of the Spin(5) lift of x; <R(x), R(y)> = +-cos(t1/2)cos(t2/2)
 * replaces the 100-term Lambda^2 form for SO(5). */
int lsk_embed_points(const void* x, long long num, int n, int is_complex, int nseg, const int* wedge, const int* part,
                     const int* group_begin, int kg_total, void* out, void* stream);

/* Both operands of one pairwise problem with a single launch (x: side A, y: side B; same n, wedge, group_begin, kg_total;
 * part_x / part_y as `part` above).  kernel(x, y) = this + lsk_pairwise_poly: two launches per covariance. */
int lsk_embed_points2(const void* x, long long num_x, const int* part_x, void* out_x,
                      const void* y, long long num_y, const int* part_y, void* out_y,
                      int n, int is_complex, int nseg, const int* wedge, const int* group_begin, int kg_total, void* stream);

/* Fused pairwise kernel: bilinear forms on the fp64 tensor-core path + class-polynomial epilogue + store.
 * Replaces, for all N*M pairs, pairwise_embed + the character loop + the weighted sum:
 *   reference space.py:80-85, spaces/so.py:136-168,288-293, spaces/su.py:91-92,175-179, spectral_kernel.py:45-54
 *   (EigenbasisSumKernel.forward), spectral_kernel.py:96-109 (RandomPhaseKernel.make_embedding, FEAT_* kinds) and the
 *   lazy Gram spectral_kernel.py:125-127,195-197 (LINEAR kind).
 * A: [n_rows] x kg_total, B: [n_cols] x kg_total panel-major operands; out: device, layout per epi->out_layout. */
int lsk_pairwise_poly(const void* A, const void* B, long long n_rows, long long n_cols, int kg_total,
                      const LskEpilogue* epi, void* out, long long ld_out, void* stream);

/* Lift of n x m frames to SO(n): complete Householder QR (LAPACK conventions) + the sign logic of
 * Stiefel.M_to_G (reference spaces/stiefel.py:38-48, spaces/grassmannian.py:74-85,120-130).
 * x: [num, n, m]; out: [num, n, n]; bad_count: device int32, incremented for frames the lift does not reproduce
 * (the reference's assert, stiefel.py:47). */
int lsk_lift_frames(const void* x, long long num, int n, int m, void* out, void* bad_count, void* stream);

/* Homogeneous-space pair kernel: for all (i, j), mean over subgroup samples h_k of the class polynomials at
 * lift(x_i)^T lift(y_j) h_k^T.  Replaces CompactHomogeneousSpace.pairwise_diff/pairwise_embed and
 * AveragedLieGroupCharacter.forward (reference space.py:125-135, 273-275).  n in {3, 4, 5}.
 * lifted_x: [n_rows, n, n], lifted_y: [n_cols, n, n], h_samples: [n_avg, n, n]; epi kind FEAT_CHEB1 (n = 3) or
 * FEAT_SPARSE (n = 4: variables s_1, s_2, z;  n = 5: s_1, s_2) with one row per output (irrep, or a single collapsed polynomial). */
int lsk_homog_pairs(const void* lifted_x, const void* lifted_y, const void* h_samples, long long n_rows, long long n_cols,
                    int n, int n_avg, const LskEpilogue* epi, void* out, long long ld_out, void* stream);

/* Fused prior sample on SO(5)/SO(2) (Stiefel V(5,3), V(5,4)):
 *   out[i] = sum_l sum_j weights[l * P + j] * Phi[i, l * P + j],   Phi = the per-irrep feature map lsk_homog_pairs would write
 * for the same arguments (epi: LSK_EPI_FEAT_SPARSE rows, feat_stride = P = n_phases; bits 4-7 of `reserved` must be 11: the caller
 * vouches that every subgroup sample is blockdiag(I_3, [[c, s], [-s, c]])).  Replaces make_embedding + einsum('nm,m->n') of
 * RandomPhaseApproximation.forward (reference prior_approximation.py:70-88) without materialising Phi.  lifted_x: [n_rows, 5, 5],
 * lifted_phases: [n_phases, 5, 5] (both from lsk_lift_frames), h_samples: [n_avg, 5, 5], weights: [nterms * feat_stride], out:
 * [n_rows]; returns LSK_EUNSUPPORTED (-3) for any other problem shape. */
int lsk_homog_sample(const void* lifted_x, const void* lifted_phases, const void* h_samples, long long n_rows, long long n_phases,
                     int n, int n_avg, const LskEpilogue* epi, const void* weights, void* out, void* stream);

/* Spherical-function features on hyperbolic space (ball model):
 *   F[i,k] = scale * coef[k] * ((1-|x_i|^2)/|x_i - shift_k|^2)^(-i lmd_k + (n-1)/2)
 * (reference spaces/hyperbolic.py:118-129, 153-158).  x: [n_points, n], shift: [m, n], lmd, coef: [m].
 * layout ROW_MAJOR / PANEL_MAJOR write the real operand [Re F | Im F] (2m columns); FEAT_COMPLEX writes complex128. */
int lsk_hyp_features(const void* x, const void* shift, const void* lmd, const void* coef, long long n_points, int m, int n,
                     double scale, void* out, int layout, long long ld_out, int out_kg, void* stream);

/* Spherical-function features on SPD(n), n = 2..6:
 *   F[i,k] = scale * coef[k] * exp(sum_r (i lmd[k,r] + rho_r) log a_r(U_i Hinv_k)),  a = |diag R| (Householder QR)
 * (reference space.py:338-346, 370-380, spaces/spd.py:83-106, 125-129).  chol_upper: [n_points, n, n] = to_group(x),
 * shift_inv: [m, n, n] = inv(shift), lmd: [m, n], coef: [m]. */
int lsk_spd_features(const void* chol_upper, const void* shift_inv, const void* lmd, const void* coef, long long n_points,
                     int m, int n, double scale, void* out, int layout, long long ld_out, int out_kg, void* stream);

/* Prior samples on H^n / SPD(n) with the feature map fused away: out[i] = sum_k (Re F[i,k] w[k] + Im F[i,k] w[m + k]), F as in
 * lsk_hyp_features / lsk_spd_features (reference prior_approximation.py:107-115, RandomFourierApproximation.forward, with
 * w = [weights_real | -weights_imag]).  The [N, m] feature matrix is never written; deterministic (no atomics). */
int lsk_hyp_sample(const void* x, const void* shift, const void* lmd, const void* coef, const void* w, long long n_points,
                   int m, int n, double scale, void* out, void* stream);
int lsk_spd_sample(const void* chol_upper, const void* shift_inv, const void* lmd, const void* coef, const void* w,
                   long long n_points, int m, int n, double scale, void* out, void* stream);

/* Iwasawa logarithms log a_r(U_i Hinv_k), out [n_points, m, n]: what the hyper-parameter gradient of the SPD Fourier-feature
 * kernel needs next to the features (dF/dlambda_kr = i log a_r F; reference spd.py:92-106 under torch autograd). */
int lsk_spd_logdiag(const void* chol_upper, const void* shift_inv, long long n_points, int m, int n, void* out, void* stream);

/* Batched small-matrix prologues for SPD(n): mode 0 = upper Cholesky factor (reference spaces/spd.py:45-46),
 * mode 1 = inverse (spaces/spd.py:66-67).  x, out: [num, n, n]; bad_count: device int32 (not PD / singular). */
int lsk_small_matrix(const void* x, long long num, int n, int mode, void* out, void* bad_count, void* stream);

/* Feature map of the flat torus [0,1)^dim (reference spaces/torus.py:14-97): panel-major [num, 2 * order] operand with
 *   F[i, 2k] = w_k cos(two_pi * freq_k . x_i),  F[i, 2k+1] = w_k sin(two_pi * freq_k . x_i)      (w = 1 when weight is NULL),
 * so that sum_k a_k Re chi_k(x_i - y_j) (TorusCharacter.chi, torus.py:91-95, summed as in spectral_kernel.py:45-54) is one
 * LINEAR lsk_pairwise_poly of F(x; w = a) and F(y).  x: [num, dim]; freq: [order, dim] float64; out_kg >= ceil(2 order / 4). */
int lsk_torus_features(const void* x, long long num, int dim, const void* freq, const void* weight, int order,
                       double two_pi, void* out, int out_kg, void* stream);

/* Haar samples from Gaussian draws: Householder QR (LAPACK conventions), q *= sign(diag r) (SU: conjugate phase),
 * q[:, :, 0] *= sign(det q) (SU: conjugate phase of det) - the deterministic part of SO.rand / SU.rand / SPD.rand_phase
 * (reference spaces/so.py:93-100, spaces/su.py:56-64, spaces/spd.py:48-55).  gauss, out: [num, n, n] float64 or complex128. */
int lsk_haar_from_gaussian(const void* gauss, long long num, int n, int is_complex, void* out, void* stream);

/* Unit-quaternion parametrisation of SO(3) (to_su2 = 0, reference spaces/so.py:74-92) or SU(2) (to_su2 = 1,
 * spaces/su.py:47-55).  gauss: [num, 4] float64;  out: [num, 3, 3] float64 or [num, 2, 2] complex128. */
int lsk_quaternion_to_group(const void* gauss, long long num, int to_su2, void* out, void* stream);

/* Prior sample from a feature map that sits in the panel-major operand layout: out[i] = scale * sum_c A[i, c] w[c]
 * (reference prior_approximation.py:85-88 `einsum('nm,m->n', embedding, weights)`, :114-115 for the Fourier-feature sampler).
 * A: [n_rows] x 4 kg_total panel-major; w: device, 4 * kg_total doubles (zero beyond the last feature); HBM-bound, deterministic. */
int lsk_panel_gemv(const void* A, long long n_rows, int kg_total, const void* w, double scale, void* out, void* stream);

/* ---- GP regression step on the covariance (the caller of the hot path: reference examples/gpr_model.py:26-39 hands
 * K + sigma_n^2 I to gpytorch.models.ExactGP / ExactMarginalLogLikelihood, i.e. to a LAPACK / cuSOLVER potrf + potrs) ---- */

/* Workspace (doubles) of lsk_potrf_upper for an n x n matrix: the inverses of the 128 x 128 diagonal blocks of the factor
 * (kept for lsk_potrs_upper) followed by one panel-major row panel. */
long long lsk_potrf_work_doubles(long long n);

/* Blocked fp64 Cholesky A = U^T U in place: on return the upper triangle of the row-major A [n, lda] holds U (the lower
 * factor L = U^T of the same buffer read column-major); the strict lower triangle is scratch.  Only the upper triangle of
 * A is read.  info_dev: device int32, 0 on success, else the 1-based index of the first non-positive pivot (LAPACK dpotrf).
 * work: device, lsk_potrf_work_doubles(n) doubles, must be kept for lsk_potrs_upper. */
int lsk_potrf_upper(void* A, long long n, long long lda, void* work, void* info_dev, void* stream);

/* Triangular solves with the factor and workspace of lsk_potrf_upper, one right-hand side, in place in b [n]:
 * which = 1: U^T y = b;  2: U x = b;  3: both (A x = b, LAPACK dpotrs).  tmp: device scratch, n + 128 doubles. */
int lsk_potrs_upper(const void* U, long long n, long long lda, const void* work, void* b, void* tmp, int which, void* stream);

/* Forward substitution with many right-hand sides, in place: B [n, nrhs] (row-major, ldb) <- U^{-T} B = L^{-1} B — the
 * whitening solve of the posterior covariance K** - V^T V, V = L^{-1} K*.  work: lsk_trsm_work_doubles(n, nrhs) doubles. */
long long lsk_trsm_work_doubles(long long n, long long nrhs);
int lsk_trsm_lower(const void* U, long long n, long long lda, const void* potrf_work, void* B, long long nrhs, long long ldb,
                   void* work, void* stream);

/* Dense inverse from the factor and workspace of lsk_potrf_upper (LAPACK dpotri, but the full symmetric matrix is written):
 * out [n, ld_out] = (U^T U)^{-1}.  The gradient of the GP marginal likelihood needs it: d/dtheta log det K = tr(K^{-1} dK).
 * work: device, lsk_potri_work_doubles(n) doubles (scratch). */
long long lsk_potri_work_doubles(long long n);
int lsk_potri_upper(const void* U, long long n, long long lda, const void* potrf_work, void* work, void* out, long long ld_out,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSK_H */
