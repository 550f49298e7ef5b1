// Homogeneous spaces M = SO(n)/H (Stiefel, Grassmannian, oriented Grassmannian).
//
// (1) lsk_lift_frames: the lift M -> G of `Stiefel.M_to_G` (spaces/stiefel.py:38-48; identical copies in
//     spaces/grassmannian.py:74-85,120-130): complete Householder QR with LAPACK's conventions
//     (beta = -sign(alpha) * norm, v_0 = 1, Q = H_0 H_1 ... H_{m-1}), columns scaled by diag(R), the row-sign
//     check of the reference, and the determinant fix on the last column.  With a finite number of H samples the
//     kernel value depends on which orthonormal completion is chosen, so this is reproduced
//     exactly rather than replaced by any other completion.
// (2) lsk_homog_pairs: for every pair (x_i, y_j) and every subgroup sample h_k the element
//         g_ijk = lift(x_i)^T lift(y_j) h_k^T                (space.py:125-135)
//     is formed in registers, its conjugation invariants tr g, tr g^2 are taken, the class polynomials are
//     evaluated and averaged over k (AveragedLieGroupCharacter.forward, space.py:273-275).  One launch produces
//     either per-irrep features Phi[i, l*stride + j] or, with a single collapsed polynomial, the covariance
//     entry itself.  Nothing of size N*M*A is ever stored (the reference materialises [N*M*A, n, n]).
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"
#include <cstring>

namespace lsk {

// ------------------------------------------------------------------------------------------------------------
// lift
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lift_kernel(const double* __restrict__ x, long long num, int n, int m, double* __restrict__ out, int* __restrict__ bad) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    double a[8][8], q[8][8], rd[8], v[8], tau[8];
    const double* xb = x + b * n * m;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < m; ++c) a[r][c] = xb[r * m + c];
    // ---- Householder QR (dgeqr2 / dlarfg) ----
    for (int j = 0; j < m; ++j) {
        const double alpha = a[j][j];
        double ss = 0.0;
        for (int r = j + 1; r < n; ++r) ss += a[r][j] * a[r][j];
        if (ss == 0.0) {
            tau[j] = 0.0;
            rd[j] = alpha;
        } else {
            const double nrm = sqrt(alpha * alpha + ss);
            const double beta = alpha >= 0.0 ? -nrm : nrm;
            tau[j] = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            for (int r = j + 1; r < n; ++r) a[r][j] *= sc;        // v (v_j = 1 implicit) stored below the diagonal
            rd[j] = beta;
            for (int c = j + 1; c < m; ++c) {                     // apply H_j to the trailing columns
                double w = a[j][c];
                for (int r = j + 1; r < n; ++r) w += a[r][j] * a[r][c];
                w *= tau[j];
                a[j][c] -= w;
                for (int r = j + 1; r < n; ++r) a[r][c] -= a[r][j] * w;
            }
        }
    }
    // ---- Q = H_0 ... H_{m-1} applied to the identity (dorg2r order: last reflector first) ----
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) q[r][c] = (r == c) ? 1.0 : 0.0;
    for (int j = m - 1; j >= 0; --j) {
        if (tau[j] == 0.0) continue;
        v[j] = 1.0;
        for (int r = j + 1; r < n; ++r) v[r] = a[r][j];
        for (int c = j; c < n; ++c) {
            double w = 0.0;
            for (int r = j; r < n; ++r) w += v[r] * q[r][c];
            w *= tau[j];
            for (int r = j; r < n; ++r) q[r][c] -= v[r] * w;
        }
    }
    // ---- stiefel.py:40-46 ----
    for (int c = 0; c < m; ++c)
        for (int r = 0; r < n; ++r) q[r][c] *= rd[c];             // g = g * r_diag  (columns; complement scaled by 1)
    bool mismatch = false;
    for (int r = 0; r < n; ++r) {
        bool close = true;
        for (int c = 0; c < m; ++c) {
            const double ref = xb[r * m + c];
            close = close && (fabs(q[r][c] - ref) <= 1e-8 + 1e-5 * fabs(ref));   // torch.isclose defaults
        }
        if (!close)
            for (int c = 0; c < n; ++c) q[r][c] = -q[r][c];
    }
    // determinant by Gaussian elimination with partial pivoting (only its sign is used)
    double lu[8][8];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) lu[r][c] = q[r][c];
    double det = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(lu[r][c]) > fabs(lu[piv][c])) piv = r;
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const double t = lu[c][k]; lu[c][k] = lu[piv][k]; lu[piv][k] = t; }
            det = -det;
        }
        det *= lu[c][c];
        if (lu[c][c] == 0.0) break;
        for (int r = c + 1; r < n; ++r) {
            const double f = lu[r][c] / lu[c][c];
            for (int k = c + 1; k < n; ++k) lu[r][k] -= f * lu[c][k];
        }
    }
    const double sgn = det > 0.0 ? 1.0 : (det < 0.0 ? -1.0 : 0.0);
    for (int r = 0; r < n; ++r) q[r][n - 1] *= sgn;
    // the reference asserts that the lift reproduces x (stiefel.py:47, torch.allclose defaults)
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < m; ++c) {
            const double ref = xb[r * m + c];
            if (!(fabs(q[r][c] - ref) <= 1e-8 + 1e-5 * fabs(ref))) mismatch = true;
        }
    if (mismatch) atomicAdd(bad, 1);
    double* ob = out + b * n * n;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) ob[r * n + c] = q[r][c];
}

// ------------------------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ double eval_sparse_row(const double* coef, int nterms, const double (&v)[NV]) {
    double acc = 0.0;
    for (int t = 0; t < nterms; ++t) {
        double m = coef[2 * t];
        const unsigned long long e = (unsigned long long)__double_as_longlong(coef[2 * t + 1]);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int ei = (int)((e >> (8 * i)) & 0xffu);
            for (int k = 0; k < ei; ++k) m *= v[i];
        }
        acc += m;
    }
    return acc;
}

__device__ __forceinline__ double eval_cheb_row(const double* coef, int nterms, double x) {
    double b1 = 0.0, b2 = 0.0;
    const double x2 = 2.0 * x;
    for (int k = nterms - 1; k >= 1; --k) {
        const double t = fma(x2, b1, coef[k] - b2);
        b2 = b1;
        b1 = t;
    }
    return fma(x, b1, coef[0] - b2);
}

// N_: matrix size (3, 4 or 5).  Invariants: tr g (n = 3);  tr g, e_2(g) (n = 5);  tr g, e_2(g), tr(* Lambda^2 g) (n = 4: the
// Hodge-star form separates the two SO(4) classes with equal spectrum).  blockDim = (32, 8): x = column j (fastest), y = row i.
template <int N_>
__global__ void __launch_bounds__(256)
homog_pair_kernel(const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ h,
                  int n_rows, int n_cols, int n_avg, double* __restrict__ out, long long ld_out,
                  const __grid_constant__ LskEpilogue p)
{
    constexpr int RANK = (N_ == 3) ? 1 : (N_ == 4 ? 3 : 2);      // number of invariants / polynomial variables
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int i = blockIdx.y * 8 + threadIdx.y;
    if (i >= n_rows || j >= n_cols) return;
    double M[N_][N_];
    {
        double xi[N_][N_], yj[N_][N_];
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                xi[a][b] = gx[(size_t)i * N_ * N_ + a * N_ + b];
                yj[a][b] = gy[(size_t)j * N_ * N_ + a * N_ + b];
            }
        // M = x_i^T y_j   (pairwise_diff(..., reverse=True), space.py:125-129);  with p.reserved & 1 the roles are
        // swapped, M = y_j^T x_i: feature maps pair (phase_j, point_i) as pairwise_embed(phases, x) does
        const bool swap = (p.reserved & 1) != 0;
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < N_; ++c) s = swap ? fma(yj[c][a], xi[c][b], s) : fma(xi[c][a], yj[c][b], s);
                M[a][b] = s;
            }
    }
    const int rows = p.nterms;
    double sum[LSK_MAX_ROWS];
#pragma unroll
    for (int l = 0; l < LSK_MAX_ROWS; ++l) sum[l] = 0.0;

    for (int k = 0; k < n_avg; ++k) {
        const double* hk = h + (size_t)k * N_ * N_;
        double form[RANK];
        if constexpr (RANK == 1) {
            double t1 = 0.0;                                     // tr(M h^T) = <M, h>_F
#pragma unroll
            for (int a = 0; a < N_; ++a)
#pragma unroll
                for (int b = 0; b < N_; ++b) t1 = fma(M[a][b], __ldg(hk + a * N_ + b), t1);
            form[0] = t1;
        } else {
            double G[N_][N_];                                    // G = M h_k^T
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                double hb[N_];
#pragma unroll
                for (int c = 0; c < N_; ++c) hb[c] = __ldg(hk + b * N_ + c);
#pragma unroll
                for (int a = 0; a < N_; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < N_; ++c) s = fma(M[a][c], hb[c], s);
                    G[a][b] = s;
                }
            }
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int a = 0; a < N_; ++a) {
                t1 += G[a][a];
#pragma unroll
                for (int b = 0; b < N_; ++b) t2 = fma(G[a][b], G[b][a], t2);
            }
            form[0] = t1;
            form[1] = 0.5 * (t1 * t1 - t2);                      // e_2(g)
            if constexpr (N_ == 4) {
                // tr(* Lambda^2 g) = sum_I sign(I, I^c) minor(rows I^c, cols I)
                auto mn = [&](int r0, int r1, int c0, int c1) { return G[r0][c0] * G[r1][c1] - G[r0][c1] * G[r1][c0]; };
                form[2] = mn(2, 3, 0, 1) + mn(0, 1, 2, 3) - mn(1, 3, 0, 2) - mn(0, 2, 1, 3) + mn(1, 2, 0, 3) + mn(0, 3, 1, 2);
            }
        }
        double var[RANK];
#pragma unroll
        for (int v = 0; v < RANK; ++v) {
            double xv = p.aff[v][0];
#pragma unroll
            for (int f = 0; f < RANK; ++f) xv = fma(p.aff[v][1 + f], form[f], xv);
            var[v] = xv;
        }
        int pos = 0;
#pragma unroll
        for (int l = 0; l < LSK_MAX_ROWS; ++l) {
            if (l < rows) {
                const int len = p.row_len[l];
                if constexpr (RANK == 1) sum[l] += eval_cheb_row(p.coef + pos, len, var[0]);
                else sum[l] += eval_sparse_row<RANK>(p.coef + 2 * pos, len, var);
                pos += len;
            }
        }
    }
    const double inv_avg = p.scale / (double)n_avg;
#pragma unroll
    for (int l = 0; l < LSK_MAX_ROWS; ++l) {
        if (l < rows) {
            const long long c = (long long)l * p.feat_stride + j;
            double* dst = (p.out_layout == LSK_OUT_PANEL_MAJOR)
                ? out + (((size_t)(i >> 6) * p.out_kg + (size_t)(c >> 2)) * PANEL + (i & 63)) * 4 + (c & 3)
                : out + (size_t)i * ld_out + c;
            *dst = sum[l] * inv_avg;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// pair kernel, fast variant: the class polynomials are linear in the monomials of the invariants, so the average over
// the subgroup samples is taken on the MONOMIALS (a fixed, fully unrolled table in registers) and the per-irrep
// coefficients are applied once per pair instead of once per (pair, sample).  Subgroup samples are staged in shared
// memory and read as broadcasts.  rank 1: S[a] = mean_k T_a(c_k), a <= AMAX;  rank 2: S[a][b] = mean_k s1^a s2^b.
// ------------------------------------------------------------------------------------------------------------
template <int N_, int AMAX, int BMAX>
__global__ void __launch_bounds__(256)
homog_pair_kernel_v2(const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ h,
                     int n_rows, int n_cols, int n_avg, double* __restrict__ out, long long ld_out,
                     const __grid_constant__ LskEpilogue p)
{
    constexpr int RANK = N_ / 2;
    constexpr int HCHUNK = 64;                                   // subgroup samples staged per pass
    __shared__ __align__(16) double sh[HCHUNK * N_ * N_ + 2];
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int i = blockIdx.y * 8 + threadIdx.y;
    const bool active = (i < n_rows) && (j < n_cols);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    double M[N_][N_];
    if (active) {
        double xi[N_][N_], yj[N_][N_];
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                xi[a][b] = gx[(size_t)i * N_ * N_ + a * N_ + b];
                yj[a][b] = gy[(size_t)j * N_ * N_ + a * N_ + b];
            }
        const bool swap = (p.reserved & 1) != 0;
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < N_; ++c) s = swap ? fma(yj[c][a], xi[c][b], s) : fma(xi[c][a], yj[c][b], s);
                M[a][b] = s;
            }
    } else {
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) M[a][b] = 0.0;
    }
    double S[(AMAX + 1) * (RANK == 1 ? 1 : BMAX + 1)];
#pragma unroll
    for (int q = 0; q < (AMAX + 1) * (RANK == 1 ? 1 : BMAX + 1); ++q) S[q] = 0.0;

    for (int k0 = 0; k0 < n_avg; k0 += HCHUNK) {
        const int kc = min(HCHUNK, n_avg - k0);
        __syncthreads();
        for (int q = tid; q < kc * N_ * N_; q += 256) sh[q] = h[(size_t)k0 * N_ * N_ + q];
        __syncthreads();
        for (int k = 0; k < kc; ++k) {
            const double* hk = sh + k * N_ * N_;
            double var[RANK];
            if constexpr (RANK == 1) {
                double t1 = 0.0;
#pragma unroll
                for (int a = 0; a < N_; ++a)
#pragma unroll
                    for (int b = 0; b < N_; ++b) t1 = fma(M[a][b], hk[a * N_ + b], t1);
                var[0] = fma(p.aff[0][1], t1, p.aff[0][0]);
            } else {
                double G[N_][N_];
#pragma unroll
                for (int b = 0; b < N_; ++b) {
                    double hb[N_];
#pragma unroll
                    for (int c = 0; c < N_; ++c) hb[c] = hk[b * N_ + c];
#pragma unroll
                    for (int a = 0; a < N_; ++a) {
                        double s = 0.0;
#pragma unroll
                        for (int c = 0; c < N_; ++c) s = fma(M[a][c], hb[c], s);
                        G[a][b] = s;
                    }
                }
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int a = 0; a < N_; ++a) {
                    t1 += G[a][a];
#pragma unroll
                    for (int b = 0; b < N_; ++b) t2 = fma(G[a][b], G[b][a], t2);
                }
                const double e2 = 0.5 * (t1 * t1 - t2);
                var[0] = fma(p.aff[0][2], e2, fma(p.aff[0][1], t1, p.aff[0][0]));
                var[1] = fma(p.aff[1][2], e2, fma(p.aff[1][1], t1, p.aff[1][0]));
            }
            if constexpr (RANK == 1) {
                double t0 = 1.0, t1c = var[0];
                const double x2 = 2.0 * var[0];
                S[0] += 1.0;
#pragma unroll
                for (int a = 1; a <= AMAX; ++a) {
                    S[a] += t1c;
                    const double nx = fma(x2, t1c, -t0);
                    t0 = t1c; t1c = nx;
                }
            } else {
                double pa[AMAX + 1], pb[BMAX + 1];
                pa[0] = 1.0; pb[0] = 1.0;
#pragma unroll
                for (int a = 1; a <= AMAX; ++a) pa[a] = pa[a - 1] * var[0];
#pragma unroll
                for (int b = 1; b <= BMAX; ++b) pb[b] = pb[b - 1] * var[1];
#pragma unroll
                for (int b = 0; b <= BMAX; ++b)
#pragma unroll
                    for (int a = 0; a <= AMAX; ++a) S[b * (AMAX + 1) + a] = fma(pa[a], pb[b], S[b * (AMAX + 1) + a]);
            }
        }
    }
    if (!active) return;
    // per-irrep coefficients applied once to the averaged monomials (dynamic exponents -> local array)
    double Sl[(AMAX + 1) * (RANK == 1 ? 1 : BMAX + 1)];
    const double inv_avg = p.scale / (double)n_avg;
#pragma unroll
    for (int q = 0; q < (AMAX + 1) * (RANK == 1 ? 1 : BMAX + 1); ++q) Sl[q] = S[q] * inv_avg;
    int pos = 0;
    for (int l = 0; l < p.nterms; ++l) {
        const int len = p.row_len[l];
        double v = 0.0;
        if constexpr (RANK == 1) {
            for (int a = 0; a < len; ++a) v = fma(p.coef[pos + a], Sl[a], v);
        } else {
            for (int t = 0; t < len; ++t) {
                const unsigned long long e = (unsigned long long)__double_as_longlong(p.coef[2 * (pos + t) + 1]);
                const int ea = (int)(e & 0xffu), eb = (int)((e >> 8) & 0xffu);
                v = fma(p.coef[2 * (pos + t)], Sl[eb * (AMAX + 1) + ea], v);
            }
        }
        pos += len;
        const long long c = (long long)l * p.feat_stride + j;
        double* dst = (p.out_layout == LSK_OUT_PANEL_MAJOR)
            ? out + (((size_t)(i >> 6) * p.out_kg + (size_t)(c >> 2)) * PANEL + (i & 63)) * 4 + (c & 3)
            : out + (size_t)i * ld_out + c;
        *dst = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// pair kernel for SO(5)/H with the shape of the problem fixed at compile time:
//   SHAPE: 4 bits per power b of s2 = number of powers of s1 that occur with it (the staircase of monomials of the first L
//          irreps: L = 10 -> rows {5, 3, 2}, L = 20 -> {7, 5, 4, 3, 1}); only those monomials are averaged;
//   MTOP:  the subgroup samples are blockdiag(I_MTOP, R_k) (Stiefel: H = SO(n - m) acts on the last n - m columns,
//          stiefel.py:25-36), so g = M h^T differs from M only in its last n - MTOP columns: 5 (5 - MTOP)^2 multiply-adds per
//          sample instead of 125, and tr g, tr g^2 split into a per-pair constant and a per-sample part.  MTOP = 0: general h.
// ------------------------------------------------------------------------------------------------------------
template <unsigned long long SHAPE> struct StairShape {
    static constexpr int rows() { int r = 0; while (r < 16 && ((SHAPE >> (4 * r)) & 15ull)) ++r; return r; }
    static constexpr int len(int b) { return (int)((SHAPE >> (4 * b)) & 15ull); }
    static constexpr int offset(int b) { int o = 0; for (int q = 0; q < b; ++q) o += len(q); return o; }
    static constexpr int total() { return offset(rows()); }
    static constexpr int amax() { int m = 0; for (int q = 0; q < rows(); ++q) m = len(q) - 1 > m ? len(q) - 1 : m; return m; }
};

template <unsigned long long SHAPE, int MTOP>
__global__ void __launch_bounds__(256)
homog5_stair_kernel(const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ h,
                    int n_rows, int n_cols, int n_avg, double* __restrict__ out, long long ld_out,
                    const __grid_constant__ LskEpilogue p)
{
    constexpr int N_ = 5, W = N_ - MTOP;                         // W: columns of g that depend on the sample
    using SH = StairShape<SHAPE>;
    constexpr int ROWS = SH::rows(), TOTAL = SH::total(), AMAX = SH::amax();
    constexpr int HCHUNK = 64;
    __shared__ __align__(16) double sh[HCHUNK * N_ * W];
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int i = blockIdx.y * 8 + threadIdx.y;
    const bool active = (i < n_rows) && (j < n_cols);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    double M[N_][N_];
#pragma unroll
    for (int a = 0; a < N_; ++a)
#pragma unroll
        for (int b = 0; b < N_; ++b) M[a][b] = 0.0;
    if (active) {
        double xi[N_][N_], yj[N_][N_];
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                xi[a][b] = gx[(size_t)i * N_ * N_ + a * N_ + b];
                yj[a][b] = gy[(size_t)j * N_ * N_ + a * N_ + b];
            }
        const bool swap = (p.reserved & 1) != 0;
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < N_; ++c) s = swap ? fma(yj[c][a], xi[c][b], s) : fma(xi[c][a], yj[c][b], s);
                M[a][b] = s;
            }
    }
    // per-pair constants: the part of tr g and tr g^2 that does not see the sample
    double t1_const = 0.0, t2_const = 0.0;
#pragma unroll
    for (int a = 0; a < MTOP; ++a) {
        t1_const += M[a][a];
#pragma unroll
        for (int b = 0; b < MTOP; ++b) t2_const = fma(M[a][b], M[b][a], t2_const);
    }
    double S[TOTAL];
#pragma unroll
    for (int q = 0; q < TOTAL; ++q) S[q] = 0.0;

    for (int k0 = 0; k0 < n_avg; k0 += HCHUNK) {
        const int kc = min(HCHUNK, n_avg - k0);
        __syncthreads();
        // stage rows MTOP.. of every sample, columns MTOP.. only (the rest is the identity / zero by contract)
        for (int q = tid; q < kc * W * W; q += 256) {
            const int k = q / (W * W), rc = q - k * (W * W), r = rc / W, c = rc - r * W;
            sh[q] = h[(size_t)(k0 + k) * N_ * N_ + (MTOP + r) * N_ + (MTOP + c)];
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < kc; ++k) {
            const double* hk = sh + k * W * W;
            // G[a][MTOP + b] = sum_c M[a][MTOP + c] h[MTOP + b][MTOP + c]
            double G[N_][W];
#pragma unroll
            for (int b = 0; b < W; ++b) {
                double hb[W];
#pragma unroll
                for (int c = 0; c < W; ++c) hb[c] = hk[b * W + c];
#pragma unroll
                for (int a = 0; a < N_; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < W; ++c) s = fma(M[a][MTOP + c], hb[c], s);
                    G[a][b] = s;
                }
            }
            double t1 = t1_const, t2 = t2_const;
#pragma unroll
            for (int b = 0; b < W; ++b) t1 += G[MTOP + b][b];
#pragma unroll
            for (int a = 0; a < MTOP; ++a)                     // cross terms: g[a][b] (changed column) * g[b][a] = M[b][a]
#pragma unroll
                for (int b = 0; b < W; ++b) t2 = fma(2.0 * G[a][b], M[MTOP + b][a], t2);
#pragma unroll
            for (int a = 0; a < W; ++a)
#pragma unroll
                for (int b = 0; b < W; ++b) t2 = fma(G[MTOP + a][b], G[MTOP + b][a], t2);
            const double e2 = 0.5 * (t1 * t1 - t2);
            const double v0 = fma(p.aff[0][2], e2, fma(p.aff[0][1], t1, p.aff[0][0]));
            const double v1 = fma(p.aff[1][2], e2, fma(p.aff[1][1], t1, p.aff[1][0]));
            double pa[AMAX + 1];
            pa[0] = 1.0;
#pragma unroll
            for (int a = 1; a <= AMAX; ++a) pa[a] = pa[a - 1] * v0;
            double pb = 1.0;
#pragma unroll
            for (int b = 0; b < ROWS; ++b) {
#pragma unroll
                for (int a = 0; a < SH::len(b); ++a) S[SH::offset(b) + a] = fma(pa[a], pb, S[SH::offset(b) + a]);
                pb *= v1;
            }
        }
    }
    if (!active) return;
    const double inv_avg = p.scale / (double)n_avg;
    // dense copy of the averaged monomials indexed [b][a] (dynamic exponents in the coefficient loop -> local array)
    double Sl[ROWS * (AMAX + 1)];
#pragma unroll
    for (int b = 0; b < ROWS; ++b)
#pragma unroll
        for (int a = 0; a <= AMAX; ++a) Sl[b * (AMAX + 1) + a] = a < SH::len(b) ? S[SH::offset(b) + a] * inv_avg : 0.0;
    int pos = 0;
    for (int l = 0; l < p.nterms; ++l) {
        const int len = p.row_len[l];
        double v = 0.0;
        for (int t = 0; t < len; ++t) {
            const unsigned long long e = (unsigned long long)__double_as_longlong(p.coef[2 * (pos + t) + 1]);
            const int ea = (int)(e & 0xffu), eb = (int)((e >> 8) & 0xffu);
            v = fma(p.coef[2 * (pos + t)], Sl[eb * (AMAX + 1) + ea], v);
        }
        pos += len;
        const long long c = (long long)l * p.feat_stride + j;
        double* dst = (p.out_layout == LSK_OUT_PANEL_MAJOR)
            ? out + (((size_t)(i >> 6) * p.out_kg + (size_t)(c >> 2)) * PANEL + (i & 63)) * 4 + (c & 3)
            : out + (size_t)i * ld_out + c;
        *dst = v;
    }
}


// ------------------------------------------------------------------------------------------------------------
// SO(5)/SO(2) with rotation samples (BASELINE config C4: V(5,3), H = SO(2); also V(5,4), whose samples are the identity):
//   h_k = blockdiag(I_3, R_k),  R_k = [[c, s], [-s, c]]   (stiefel.py:25-36 with so.py:62-70)
// For such samples the invariants of g = M h_k^T are TRIGONOMETRIC POLYNOMIALS of the sample angle with per-pair constants
// (M = [[T, U], [V, W]] with T 3x3, W = [[p, q], [r, t]] 2x2, Q = V U):
//   tr g   = tr T + (p + t) c + (q - r) s
//   tr g^2 = tr T^2 + (u + v)/2 + 2 (Q00 + Q11) c + 2 (Q01 - Q10) s + (u - v)/2 cos 2phi + (w/2) sin 2phi,
//            u = p^2 + t^2 + 2 q r,  v = q^2 + r^2 - 2 p t,  w = 2 (p q - r t + q t - p r)
// so a sample costs 6 multiply-adds for the invariants (the general kernel above: 42) + the staircase monomials; the 5 x 5
// matrix M is dead once the eight constants exist, which takes the kernel from 126 registers to ~70.  The per-irrep coefficients are
// applied to the averaged monomials through a DENSE [irrep][monomial] table (built by the host from the sparse list): every
// index is a compile-time constant, nothing lives in local memory.
//   MODE 0: per-irrep features, CTA = 64 rows x 4 columns -> every store instruction of a warp writes 8 rows x 32 B of one
//           k-group of the panel-major map (256 contiguous bytes);
//   MODE 1: fused prior sample f[i] = sum_l sum_j w[l P + j] Phi[i, l P + j] (prior_approximation.py:85-88): CTA = 8 rows x 32
//           phase lanes looping over all phases, fixed-order shuffle reduction - the 5.4 GB feature map of C4 is never written.
// ------------------------------------------------------------------------------------------------------------
template <unsigned long long SHAPE>
struct RotPair {
    using SH = StairShape<SHAPE>;
    static constexpr int ROWS = SH::rows(), TOTAL = SH::total(), AMAX = SH::amax();

    // eight per-pair constants of the two trigonometric polynomials
    struct Consts { double t1_0, t1_c, t1_s, t2_0, t2_c, t2_s, t2_c2, t2_s2; };

    __device__ static __forceinline__ Consts pair_constants(const double* __restrict__ gi, const double* __restrict__ gj, bool swap) {
        // M = lift_i^T lift_j (swap: lift_j^T lift_i), accumulated row pair by row pair: 25 accumulators + 10 operands live
        double M[5][5];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int b = 0; b < 5; ++b) M[a][b] = 0.0;
        const double* __restrict__ ga = swap ? gj : gi;       // swap: M = lift_j^T lift_i - the same code on exchanged pointers
        const double* __restrict__ gb = swap ? gi : gj;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            double ar[5], br[5];
#pragma unroll
            for (int a = 0; a < 5; ++a) { ar[a] = ga[c * 5 + a]; br[a] = gb[c * 5 + a]; }
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int b = 0; b < 5; ++b) M[a][b] = fma(ar[a], br[b], M[a][b]);
        }
        Consts k;
        double trT = 0.0, trT2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            trT += M[a][a];
#pragma unroll
            for (int b = 0; b < 3; ++b) trT2 = fma(M[a][b], M[b][a], trT2);
        }
        const double p = M[3][3], q = M[3][4], r = M[4][3], t = M[4][4];
        double Q[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) s = fma(M[3 + a][c], M[c][3 + b], s);
                Q[a][b] = s;
            }
        const double u = fma(p, p, fma(t, t, 2.0 * q * r)), v = fma(q, q, fma(r, r, -2.0 * p * t));
        const double w = 2.0 * (fma(p, q, -r * t) + fma(q, t, -p * r));
        k.t1_0 = trT; k.t1_c = p + t; k.t1_s = q - r;
        k.t2_0 = trT2 + 0.5 * (u + v);
        k.t2_c = 2.0 * (Q[0][0] + Q[1][1]); k.t2_s = 2.0 * (Q[0][1] - Q[1][0]);
        k.t2_c2 = 0.5 * (u - v); k.t2_s2 = 0.5 * w;
        return k;
    }

    // S[m] += monomials of (v0, v1) for one sample (c, s, cos 2phi, sin 2phi)
    __device__ static __forceinline__ void add_sample(const Consts& k, const double4 cs, const double (&aff)[2][3], double (&S)[TOTAL]) {
        const double t1 = fma(k.t1_s, cs.y, fma(k.t1_c, cs.x, k.t1_0));
        const double t2 = fma(k.t2_s2, cs.w, fma(k.t2_c2, cs.z, fma(k.t2_s, cs.y, fma(k.t2_c, cs.x, k.t2_0))));
        const double e2 = 0.5 * fma(t1, t1, -t2);
        const double v0 = fma(aff[0][2], e2, fma(aff[0][1], t1, aff[0][0]));
        const double v1 = fma(aff[1][2], e2, fma(aff[1][1], t1, aff[1][0]));
        double pa[AMAX + 1];
        pa[0] = 1.0;
#pragma unroll
        for (int a = 1; a <= AMAX; ++a) pa[a] = pa[a - 1] * v0;
        double pb = 1.0;
#pragma unroll
        for (int b = 0; b < ROWS; ++b) {
#pragma unroll
            for (int a = 0; a < SH::len(b); ++a) S[SH::offset(b) + a] = fma(pa[a], pb, S[SH::offset(b) + a]);
            if (b + 1 < ROWS) pb *= v1;
        }
    }
};

constexpr int ROT_CHUNK = 128;          // subgroup samples staged per pass (4 doubles each)
constexpr int ROT_COLS_PER_CTA = 32;    // MODE 0: columns per CTA (8 groups of 4)

template <unsigned long long SHAPE, int MODE>
__global__ void __launch_bounds__(256, 3)
homog5_rot_kernel(const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ h,
                  int n_rows, int n_cols, int n_avg, double* __restrict__ out, long long ld_out,
                  const double* __restrict__ weights, const __grid_constant__ LskEpilogue p)
{
    using RP = RotPair<SHAPE>;
    constexpr int TOTAL = RP::TOTAL;
    __shared__ __align__(32) double4 sh[ROT_CHUNK];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    // MODE 0: blockDim = (4, 64): x = column lane, y = row inside the 64-row panel.  MODE 1: blockDim = (32, 8).
    const int i = MODE == 0 ? blockIdx.y * 64 + threadIdx.y : blockIdx.x * 8 + threadIdx.y;
    const bool swap = (p.reserved & 1) != 0;
    double aff[2][3];
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) aff[v][c] = p.aff[v][c];
    const double inv_avg = p.scale / (double)n_avg;
    // MODE 0: ROT_COLS_PER_CTA / 4 column groups per CTA (the subgroup samples are staged once per CTA, and a CTA lives long enough
    // for its launch and drain to disappear); MODE 1: all columns, 32 at a time
    const int j_begin = MODE == 0 ? blockIdx.x * ROT_COLS_PER_CTA + threadIdx.x : threadIdx.x;
    const int j_step = MODE == 0 ? 4 : 32;
    const int j_end = MODE == 0 ? min(blockIdx.x * ROT_COLS_PER_CTA + ROT_COLS_PER_CTA, ((n_cols + 3) / 4) * 4) : ((n_cols + 31) / 32) * 32;
    double fsum = 0.0;
    for (int j0 = j_begin; j0 < j_end; j0 += j_step) {
        const int j = j0;
        const bool active = (i < n_rows) && (j < n_cols);
        typename RP::Consts k = {};
        if (active) k = RP::pair_constants(gx + (size_t)i * 25, gy + (size_t)j * 25, swap);
        double S[TOTAL];
#pragma unroll
        for (int m = 0; m < TOTAL; ++m) S[m] = 0.0;
        for (int k0 = 0; k0 < n_avg; k0 += ROT_CHUNK) {
            const int kc = min(ROT_CHUNK, n_avg - k0);
            if (n_avg > ROT_CHUNK || j0 == j_begin) {                   // few samples: staged once per CTA
                __syncthreads();
                for (int q = tid; q < kc; q += 256) {
                    const double c = h[(size_t)(k0 + q) * 25 + 18], s = h[(size_t)(k0 + q) * 25 + 19];   // R[0][0], R[0][1]
                    sh[q] = make_double4(c, s, fma(c, c, -s * s), 2.0 * c * s);
                }
                __syncthreads();
            }
#pragma unroll 2
            for (int q = 0; q < kc; ++q) RP::add_sample(k, sh[q], aff, S);
        }
        // dense coefficient table: irrep l, monomial m -> coef[l * TOTAL + m]
        if (MODE == 0) {
            if (active)
            for (int l = 0; l < p.nterms; ++l) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < TOTAL; ++m) v = fma(p.coef[l * TOTAL + m], S[m], v);
                v *= inv_avg;
                const long long c = (long long)l * p.feat_stride + j;
                double* dst = (p.out_layout == LSK_OUT_PANEL_MAJOR)
                    ? out + (((size_t)(i >> 6) * p.out_kg + (size_t)(c >> 2)) * PANEL + (i & 63)) * 4 + (c & 3)
                    : out + (size_t)i * ld_out + c;
                *dst = v;
            }
        } else if (active) {
            for (int l = 0; l < p.nterms; ++l) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < TOTAL; ++m) v = fma(p.coef[l * TOTAL + m], S[m], v);
                fsum = fma(v * inv_avg, weights[(size_t)l * p.feat_stride + j], fsum);
            }
        }
    }
    if (MODE == 1) {
        // fixed-order reduction over the 32 phase lanes of the row
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
        if (threadIdx.x == 0 && i < n_rows) out[i] = fsum;
    }
}

// sparse per-irrep monomial lists -> dense [irrep][monomial of the staircase SHAPE]; false if a term falls outside the staircase
template <unsigned long long SHAPE>
static bool densify_rows(const LskEpilogue& p, LskEpilogue& q) {
    using SH = StairShape<SHAPE>;
    constexpr int TOTAL = SH::total();
    if (p.nterms * TOTAL > LSK_MAX_COEF) return false;
    q = p;
    for (int t = 0; t < LSK_MAX_COEF; ++t) q.coef[t] = 0.0;
    int pos = 0;
    for (int l = 0; l < p.nterms; ++l) {
        for (int t = 0; t < p.row_len[l]; ++t) {
            unsigned long long e;
            memcpy(&e, &p.coef[2 * (pos + t) + 1], 8);
            const int ea = (int)(e & 0xffu), eb = (int)((e >> 8) & 0xffu);
            if (eb >= SH::rows() || ea >= SH::len(eb)) return false;
            q.coef[l * TOTAL + SH::offset(eb) + ea] += p.coef[2 * (pos + t)];
        }
        pos += p.row_len[l];
    }
    return true;
}

}  // namespace lsk

extern "C" int lsk_lift_frames(const void* x, long long num, int n, int m, void* out, void* bad_count, void* stream) {
    using namespace lsk;
    if (!x || !out || !bad_count) return LSK_EARG;
    if (num < 0 || n < 2 || n > 8 || m < 1 || m >= n) return LSK_ESIZE;
    if (num == 0) return 0;
    const long long blocks = (num + 127) / 128;
    lift_kernel<<<(unsigned)blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(x), num, n, m, static_cast<double*>(out), static_cast<int*>(bad_count));
    return (int)cudaGetLastError();
}

extern "C" int lsk_homog_pairs(const void* lifted_x, const void* lifted_y, const void* h_samples, long long n_rows,
                               long long n_cols, int n, int n_avg, const LskEpilogue* epi, void* out, long long ld_out,
                               void* stream) {
    using namespace lsk;
    if (!lifted_x || !lifted_y || !h_samples || !epi || !out) return LSK_EARG;
    if (n_rows < 0 || n_cols < 0 || n_avg < 1) return LSK_ESIZE;
    if (n_rows == 0 || n_cols == 0) return 0;
    // the pair kernels index rows / columns with int and put 8 rows per CTA on grid.y (<= 65535)
    if (n_rows > 8ll * 65535 || n_cols > 0x7fffffffll - 32) return LSK_ESIZE;
    const LskEpilogue& p = *epi;
    if (p.nterms < 1 || p.nterms > LSK_MAX_ROWS) return LSK_EARG;
    if (p.out_layout == LSK_OUT_ROW_MAJOR && ld_out < n_cols) return LSK_ESIZE;
    if (p.out_layout == LSK_OUT_PANEL_MAJOR && p.out_kg < 1) return LSK_ESIZE;
    int tot = 0;
    for (int l = 0; l < p.nterms; ++l) { if (p.row_len[l] < 1) return LSK_EARG; tot += p.row_len[l]; }
    dim3 block(32, 8), grid((unsigned)((n_cols + 31) / 32), (unsigned)((n_rows + 7) / 8));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* gx = static_cast<const double*>(lifted_x);
    const double* gy = static_cast<const double*>(lifted_y);
    const double* hh = static_cast<const double*>(h_samples);
    double* o = static_cast<double*>(out);
    if (n == 3) {
        if (p.kind != LSK_EPI_FEAT_CHEB1 || tot > LSK_MAX_COEF) return LSK_EARG;
        int maxlen = 0;
        for (int l = 0; l < p.nterms; ++l) maxlen = p.row_len[l] > maxlen ? p.row_len[l] : maxlen;
        if (maxlen <= 16)          // Chebyshev degree <= 15: averaged-monomial kernel
            homog_pair_kernel_v2<3, 15, 0><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p);
        else
            homog_pair_kernel<3><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p);
    } else if (n == 4) {
        if (p.kind != LSK_EPI_FEAT_SPARSE || 2 * tot > LSK_MAX_COEF) return LSK_EARG;
        homog_pair_kernel<4><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p);
    } else if (n == 5) {
        if (p.kind != LSK_EPI_FEAT_SPARSE || 2 * tot > LSK_MAX_COEF) return LSK_EARG;
        int amax = 0, bmax = 0;
        for (int t = 0; t < tot; ++t) {
            unsigned long long e;
            memcpy(&e, &p.coef[2 * t + 1], 8);
            const int ea = (int)(e & 0xffu), eb = (int)((e >> 8) & 0xffu);
            amax = ea > amax ? ea : amax;
            bmax = eb > bmax ? eb : bmax;
        }
        // staircase of the monomials that occur: rows[b] = 1 + largest power of s1 next to s2^b
        int rows[16] = {0};
        bool fits16 = true;
        for (int t = 0; t < tot; ++t) {
            unsigned long long e;
            memcpy(&e, &p.coef[2 * t + 1], 8);
            const int ea = (int)(e & 0xffu), eb = (int)((e >> 8) & 0xffu);
            if (eb >= 16 || ea >= 15) { fits16 = false; break; }
            rows[eb] = ea + 1 > rows[eb] ? ea + 1 : rows[eb];
        }
        auto covered = [&](unsigned long long shape) {
            if (!fits16) return false;
            for (int b = 0; b < 16; ++b) if (rows[b] > (int)((shape >> (4 * b)) & 15ull)) return false;
            return true;
        };
        int mtop = (p.reserved >> 4) & 15;               // the caller vouches: samples are blockdiag(I_mtop, R)
        constexpr unsigned long long SH10 = 0x235ull, SH20 = 0x13457ull;
        if (mtop == 11) {
            // t = 3 and every R is a plane rotation [[c, s], [-s, c]]: trigonometric-polynomial kernel
            LskEpilogue q;
            dim3 rblock(4, 64), rgrid((unsigned)((n_cols + ROT_COLS_PER_CTA - 1) / ROT_COLS_PER_CTA), (unsigned)((n_rows + 63) / 64));
            if (rgrid.y <= 65535u) {
                if (covered(SH10) && densify_rows<SH10>(p, q)) {
                    homog5_rot_kernel<SH10, 0><<<rgrid, rblock, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, nullptr, q);
                    return (int)cudaGetLastError();
                }
                if (covered(SH20) && densify_rows<SH20>(p, q)) {
                    homog5_rot_kernel<SH20, 0><<<rgrid, rblock, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, nullptr, q);
                    return (int)cudaGetLastError();
                }
            }
            mtop = 3;                                    // other shapes: the block-structured kernel below
        }
#define LSK_H5(SHAPE, MT) homog5_stair_kernel<SHAPE, MT><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p)
        if (mtop > 4 || mtop == 1) return LSK_EARG;
        if (covered(SH10)) {
            if (mtop >= 3) LSK_H5(SH10, 3); else if (mtop == 2) LSK_H5(SH10, 2); else LSK_H5(SH10, 0);
        } else if (covered(SH20)) {
            if (mtop >= 3) LSK_H5(SH20, 3); else if (mtop == 2) LSK_H5(SH20, 2); else LSK_H5(SH20, 0);
        } else if (amax <= 6 && bmax <= 3)
            homog_pair_kernel_v2<5, 6, 3><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p);
#undef LSK_H5
        else
            homog_pair_kernel<5><<<grid, block, 0, st>>>(gx, gy, hh, (int)n_rows, (int)n_cols, n_avg, o, ld_out, p);
    } else {
        return LSK_EUNSUPPORTED;
    }
    return (int)cudaGetLastError();
}


// Fused prior sample on SO(5)/SO(2) (V(5,3) / V(5,4) with rotation samples):  out[i] = sum_l sum_j w[l * P + j] * Phi[i, l * P + j]
// with Phi the per-irrep feature map lsk_homog_pairs would write (same epilogue struct, FEAT_SPARSE rows, feat_stride = P).
// Replaces make_embedding + einsum('nm,m->n') of prior_approximation.py:70-88; the feature map is never materialised.
// LSK_EUNSUPPORTED when the problem is not of that form (the caller then writes the map and reduces it with lsk_panel_gemv).
extern "C" int lsk_homog_sample(const void* lifted_x, const void* lifted_phases, const void* h_samples, long long n_rows,
                                long long n_phases, int n, int n_avg, const LskEpilogue* epi, const void* weights, void* out,
                                void* stream) {
    using namespace lsk;
    if (!lifted_x || !lifted_phases || !h_samples || !epi || !weights || !out) return LSK_EARG;
    if (n_rows < 0 || n_phases < 0 || n_avg < 1) return LSK_ESIZE;
    if (n_rows > 0x7fffffffll - 64 || n_phases > 0x7fffffffll - 64) return LSK_ESIZE;
    const LskEpilogue& p = *epi;
    if (n != 5 || p.kind != LSK_EPI_FEAT_SPARSE || ((p.reserved >> 4) & 15) != 11) return LSK_EUNSUPPORTED;
    if (p.nterms < 1 || p.nterms > LSK_MAX_ROWS || p.feat_stride < n_phases) return LSK_EARG;
    int tot = 0;
    for (int l = 0; l < p.nterms; ++l) { if (p.row_len[l] < 1) return LSK_EARG; tot += p.row_len[l]; }
    if (2 * tot > LSK_MAX_COEF) return LSK_EARG;
    if (n_rows == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_phases == 0) return (int)cudaMemsetAsync(out, 0, (size_t)n_rows * sizeof(double), st);
    constexpr unsigned long long SH10 = 0x235ull, SH20 = 0x13457ull;
    LskEpilogue q;
    dim3 block(32, 8), grid((unsigned)((n_rows + 7) / 8));
    const double* gx = static_cast<const double*>(lifted_x);
    const double* gp = static_cast<const double*>(lifted_phases);
    const double* hh = static_cast<const double*>(h_samples);
    const double* w = static_cast<const double*>(weights);
    double* o = static_cast<double*>(out);
    if (densify_rows<SH10>(p, q))
        homog5_rot_kernel<SH10, 1><<<grid, block, 0, st>>>(gx, gp, hh, (int)n_rows, (int)n_phases, n_avg, o, 0, w, q);
    else if (densify_rows<SH20>(p, q))
        homog5_rot_kernel<SH20, 1><<<grid, block, 0, st>>>(gx, gp, hh, (int)n_rows, (int)n_phases, n_avg, o, 0, w, q);
    else
        return LSK_EUNSUPPORTED;
    return (int)cudaGetLastError();
}
