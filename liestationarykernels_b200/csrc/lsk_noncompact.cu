// Spherical-function features on non-compact symmetric spaces (random Fourier features).
//
//   H^n (ball model):  F[i,k] = c_k * ((1 - |x_i|^2) / |x_i - b_k|^2)^(-i lambda_k + rho),  rho = (n-1)/2
//                      (spaces/hyperbolic.py:118-129, 153-158)
//   SPD(n) = GL(n)/O(n):  F[i,k] = c_k * exp( sum_r (i lambda_kr + rho_r) log a_r(U_i h_k^{-1}) ),
//                      U_i = chol(x_i) upper, a = |diag R| of the QR factorisation, rho_r = r + 1 - (n+1)/2
//                      (space.py:338-346, 370-380; spaces/spd.py:83-106, 125-129)
//
// One thread per (point, feature); everything between the two small inputs and the feature value stays in
// registers (the reference materialises [N*m, n, n] products and runs a batched LAPACK QR over them).  The feature
// matrix is written once, either as complex128 [N, m] (the reference's `lb_eigenspaces(x)` layout) or directly as
// the real operand Psi = [Re F | Im F] of the Gram kernel, row-major or in the tensor-core panel layout.
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"

#define LSK_FEAT_COMPLEX   2   // out[i, k] complex128 interleaved (re, im), row-major with ld_out complex elements per row

namespace lsk {

__device__ __forceinline__ void store_feature(double* out, int layout, long long ld_out, int out_kg, int i, int k, int m,
                                              double re, double im) {
    if (layout == LSK_FEAT_COMPLEX) {
        double* dst = out + ((size_t)i * ld_out + k) * 2;
        dst[0] = re; dst[1] = im;
    } else if (layout == LSK_OUT_PANEL_MAJOR) {
        const size_t base = ((size_t)(i >> 6) * out_kg) * (PANEL * 4) + (size_t)(i & 63) * 4;
        const int c0 = k, c1 = m + k;
        out[base + (size_t)(c0 >> 2) * (PANEL * 4) + (c0 & 3)] = re;
        out[base + (size_t)(c1 >> 2) * (PANEL * 4) + (c1 & 3)] = im;
    } else {
        out[(size_t)i * ld_out + k] = re;
        out[(size_t)i * ld_out + m + k] = im;
    }
}

// sum over the 64 threads that share threadIdx.y (two warps), fixed order: deterministic
__device__ __forceinline__ double reduce_row_of_64(double v, double (*part)[2]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.y][threadIdx.x >> 5] = v;
    __syncthreads();
    return part[threadIdx.y][0] + part[threadIdx.y][1];
}

// blockDim = (64, 4): x -> feature k (fastest, coalesced), y -> point i.
// SAMPLE: the prior sample f[i] = sum_k (Re F[i,k] w[k] + Im F[i,k] w[m + k]) is reduced in the kernel (grid.x = 1, every
// thread strides over the features) and the feature matrix is never written (prior_approximation.py:107-115).
template <bool SAMPLE>
__global__ void __launch_bounds__(256)
hyp_features_kernel(const double* __restrict__ x, const double* __restrict__ shift, const double* __restrict__ lmd,
                    const double* __restrict__ coef, const double* __restrict__ w, int n_points, int m, int n, double scale,
                    double* __restrict__ out, int layout, long long ld_out, int out_kg)
{
    __shared__ double part[4][2];
    const int i = blockIdx.y * 4 + threadIdx.y;
    const bool row_ok = i < n_points;
    if (!SAMPLE && !row_ok) return;
    const double rho = 0.5 * (n - 1);
    double acc = 0.0;
    for (int k = blockIdx.x * 64 + threadIdx.x; k < m; k += SAMPLE ? 64 : m) {
        double x2 = 0.0, d2 = 0.0;
        for (int a = 0; a < n; ++a) {
            const double xa = row_ok ? x[(size_t)i * n + a] : 0.0;
            const double d = xa - shift[(size_t)k * n + a];
            x2 += xa * xa;            // sum(square(x))            (hyperbolic.py:126)
            d2 += d * d;              // sum(square(x - b))        (hyperbolic.py:124)
        }
        const double log_xb = log(1.0 - x2) - log(d2);
        const double mag = exp(rho * log_xb) * coef[k] * scale;
        double s, c;
        sincos(-lmd[k] * log_xb, &s, &c);
        if constexpr (SAMPLE) acc = fma(mag * c, w[k], fma(mag * s, w[m + k], acc));
        else store_feature(out, layout, ld_out, out_kg, i, k, m, mag * c, mag * s);
    }
    if constexpr (SAMPLE) {
        const double tot = reduce_row_of_64(acc, part);
        if (threadIdx.x == 0 && row_ok) out[i] = tot;
    }
}

// Householder QR of an N_ x N_ matrix held in registers; returns log|R_rr|
template <int N_>
__device__ __forceinline__ void log_abs_diag_r(double (&A)[N_][N_], double (&la)[N_]) {
#pragma unroll
    for (int j = 0; j < N_; ++j) {
        const double alpha = A[j][j];
        double ss = 0.0;
#pragma unroll
        for (int r = j + 1; r < N_; ++r) ss = fma(A[r][j], A[r][j], ss);
        if (j == N_ - 1 || ss == 0.0) {
            la[j] = log(fabs(alpha));
        } else {
            const double nrm = sqrt(fma(alpha, alpha, ss));
            const double beta = alpha >= 0.0 ? -nrm : nrm;
            la[j] = log(nrm);
            const double tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
#pragma unroll
            for (int r = j + 1; r < N_; ++r) A[r][j] *= sc;
#pragma unroll
            for (int c = j + 1; c < N_; ++c) {
                double w = A[j][c];
#pragma unroll
                for (int r = j + 1; r < N_; ++r) w = fma(A[r][j], A[r][c], w);
                w *= tau;
                A[j][c] -= w;
#pragma unroll
                for (int r = j + 1; r < N_; ++r) A[r][c] = fma(-A[r][j], w, A[r][c]);
            }
        }
    }
}

template <int N_, bool SAMPLE>
__global__ void __launch_bounds__(256)
spd_features_kernel(const double* __restrict__ U, const double* __restrict__ Hinv, const double* __restrict__ lmd,
                    const double* __restrict__ coef, const double* __restrict__ w, int n_points, int m, double scale,
                    double* __restrict__ out, int layout, long long ld_out, int out_kg)
{
    __shared__ double part[4][2];
    const int i = blockIdx.y * 4 + threadIdx.y;
    const bool row_ok = i < n_points;
    if (!SAMPLE && !row_ok) return;
    double u[N_][N_];
#pragma unroll
    for (int a = 0; a < N_; ++a)
#pragma unroll
        for (int b = 0; b < N_; ++b) u[a][b] = row_ok ? U[(size_t)i * N_ * N_ + a * N_ + b] : (a == b ? 1.0 : 0.0);
    double acc = 0.0;
    for (int k = blockIdx.x * 64 + threadIdx.x; k < m; k += SAMPLE ? 64 : m) {
        double A[N_][N_];
        {
            double h[N_][N_];
#pragma unroll
            for (int a = 0; a < N_; ++a)
#pragma unroll
                for (int b = 0; b < N_; ++b) h[a][b] = Hinv[(size_t)k * N_ * N_ + a * N_ + b];
#pragma unroll
            for (int a = 0; a < N_; ++a)
#pragma unroll
                for (int b = 0; b < N_; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < N_; ++c) s = fma(u[a][c], h[c][b], s);
                    A[a][b] = s;                                       // U_i h_k^{-1}   (space.py:338-346)
                }
        }
        double la[N_];
        log_abs_diag_r<N_>(A, la);
        double re = 0.0, ph = 0.0;
#pragma unroll
        for (int r = 0; r < N_; ++r) {
            const double rho = (r + 1) - 0.5 * (N_ + 1);               // spd.py:87-90
            re = fma(rho, la[r], re);
            ph = fma(lmd[(size_t)k * N_ + r], la[r], ph);
        }
        const double mag = exp(re) * coef[k] * scale;
        double s, c;
        sincos(ph, &s, &c);
        if constexpr (SAMPLE) acc = fma(mag * c, w[k], fma(mag * s, w[m + k], acc));
        else store_feature(out, layout, ld_out, out_kg, i, k, m, mag * c, mag * s);
    }
    if constexpr (SAMPLE) {
        const double tot = reduce_row_of_64(acc, part);
        if (threadIdx.x == 0 && row_ok) out[i] = tot;
    }
}

// log a_r(U_i h_k^{-1}) for every (point, feature): the Iwasawa logarithms the gradient of the SPD Fourier-feature kernel
// needs (dF/dlambda_kr = i log a_r F);  out[i, k, r]
template <int N_>
__global__ void __launch_bounds__(256)
spd_logdiag_kernel(const double* __restrict__ U, const double* __restrict__ Hinv, int n_points, int m, double* __restrict__ out)
{
    const int k = blockIdx.x * 64 + threadIdx.x;
    const int i = blockIdx.y * 4 + threadIdx.y;
    if (i >= n_points || k >= m) return;
    double A[N_][N_];
    {
        double u[N_][N_], h[N_][N_];
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                u[a][b] = U[(size_t)i * N_ * N_ + a * N_ + b];
                h[a][b] = Hinv[(size_t)k * N_ * N_ + a * N_ + b];
            }
#pragma unroll
        for (int a = 0; a < N_; ++a)
#pragma unroll
            for (int b = 0; b < N_; ++b) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < N_; ++c) s = fma(u[a][c], h[c][b], s);
                A[a][b] = s;
            }
    }
    double la[N_];
    log_abs_diag_r<N_>(A, la);
#pragma unroll
    for (int r = 0; r < N_; ++r) out[((size_t)i * m + k) * N_ + r] = la[r];
}

// mode 0: upper Cholesky factor (torch.linalg.cholesky(x, upper=True), spd.py:45-46);  mode 1: inverse (spd.py:66-67)
__global__ void __launch_bounds__(128)
small_matrix_kernel(const double* __restrict__ x, long long num, int n, int mode, double* __restrict__ out, int* __restrict__ bad) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    double a[8][8];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) a[r][c] = x[b * n * n + r * n + c];
    double* o = out + b * n * n;
    if (mode == 0) {
        // x = U^T U: column-oriented (LAPACK dpotrf 'U' order)
        double u[8][8];
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) u[r][c] = 0.0;
        bool ok = true;
        for (int j = 0; j < n; ++j) {
            double d = a[j][j];
            for (int k = 0; k < j; ++k) d -= u[k][j] * u[k][j];
            if (!(d > 0.0)) { ok = false; d = nan(""); }
            const double ujj = sqrt(d);
            u[j][j] = ujj;
            for (int c = j + 1; c < n; ++c) {
                double s = a[j][c];
                for (int k = 0; k < j; ++k) s -= u[k][j] * u[k][c];
                u[j][c] = s / ujj;
            }
        }
        if (!ok) atomicAdd(bad, 1);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) o[r * n + c] = u[r][c];
    } else {
        // Gauss-Jordan with partial pivoting
        double inv[8][8];
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) inv[r][c] = (r == c) ? 1.0 : 0.0;
        for (int c = 0; c < n; ++c) {
            int piv = c;
            for (int r = c + 1; r < n; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
            if (a[piv][c] == 0.0) { atomicAdd(bad, 1); break; }
            if (piv != c)
                for (int k = 0; k < n; ++k) {
                    double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t;
                    t = inv[c][k]; inv[c][k] = inv[piv][k]; inv[piv][k] = t;
                }
            const double d = 1.0 / a[c][c];
            for (int k = 0; k < n; ++k) { a[c][k] *= d; inv[c][k] *= d; }
            for (int r = 0; r < n; ++r) {
                if (r == c) continue;
                const double f = a[r][c];
                if (f != 0.0)
                    for (int k = 0; k < n; ++k) { a[r][k] -= f * a[c][k]; inv[r][k] -= f * inv[c][k]; }
            }
        }
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) o[r * n + c] = inv[r][c];
    }
}

}  // namespace lsk

static int check_layout(int layout, long long ld_out, int m, int out_kg) {
    if (layout == LSK_FEAT_COMPLEX) return ld_out >= m ? 0 : LSK_ESIZE;
    if (layout == LSK_OUT_ROW_MAJOR) return ld_out >= 2LL * m ? 0 : LSK_ESIZE;
    if (layout == LSK_OUT_PANEL_MAJOR) return out_kg * 4LL >= 2LL * m ? 0 : LSK_ESIZE;
    return LSK_EARG;
}

extern "C" int lsk_hyp_features(const void* x, const void* shift, const void* lmd, const void* coef, long long n_points,
                                int m, int n, double scale, void* out, int layout, long long ld_out, int out_kg, void* stream) {
    using namespace lsk;
    if (!x || !shift || !lmd || !coef || !out) return LSK_EARG;
    if (n_points < 0 || m < 0 || n < 1 || n > 64) return LSK_ESIZE;
    if (n_points == 0 || m == 0) return 0;
    if (int rc = check_layout(layout, ld_out, m, out_kg)) return rc;
    dim3 block(64, 4), grid((unsigned)((m + 63) / 64), (unsigned)((n_points + 3) / 4));
    hyp_features_kernel<false><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(x), static_cast<const double*>(shift), static_cast<const double*>(lmd),
        static_cast<const double*>(coef), nullptr, (int)n_points, m, n, scale, static_cast<double*>(out), layout, ld_out, out_kg);
    return (int)cudaGetLastError();
}

extern "C" int lsk_hyp_sample(const void* x, const void* shift, const void* lmd, const void* coef, const void* w,
                              long long n_points, int m, int n, double scale, void* out, void* stream) {
    using namespace lsk;
    if (!x || !shift || !lmd || !coef || !w || !out) return LSK_EARG;
    if (n_points < 0 || m < 0 || n < 1 || n > 64) return LSK_ESIZE;
    if (n_points == 0) return 0;
    dim3 block(64, 4), grid(1, (unsigned)((n_points + 3) / 4));
    hyp_features_kernel<true><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(x), static_cast<const double*>(shift), static_cast<const double*>(lmd),
        static_cast<const double*>(coef), static_cast<const double*>(w), (int)n_points, m, n, scale,
        static_cast<double*>(out), 0, 0, 0);
    return (int)cudaGetLastError();
}

extern "C" int lsk_spd_features(const void* chol_upper, const void* shift_inv, const void* lmd, const void* coef,
                                long long n_points, int m, int n, double scale, void* out, int layout, long long ld_out,
                                int out_kg, void* stream) {
    using namespace lsk;
    if (!chol_upper || !shift_inv || !lmd || !coef || !out) return LSK_EARG;
    if (n_points < 0 || m < 0) return LSK_ESIZE;
    if (n_points == 0 || m == 0) return 0;
    if (int rc = check_layout(layout, ld_out, m, out_kg)) return rc;
    dim3 block(64, 4), grid((unsigned)((m + 63) / 64), (unsigned)((n_points + 3) / 4));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* u = static_cast<const double*>(chol_upper);
    const double* h = static_cast<const double*>(shift_inv);
    const double* l = static_cast<const double*>(lmd);
    const double* c = static_cast<const double*>(coef);
    double* o = static_cast<double*>(out);
#define LSK_SPD(N_) spd_features_kernel<N_, false><<<grid, block, 0, st>>>(u, h, l, c, nullptr, (int)n_points, m, scale, o, layout, ld_out, out_kg)
    switch (n) {
    case 2: LSK_SPD(2); break;
    case 3: LSK_SPD(3); break;
    case 4: LSK_SPD(4); break;
    case 5: LSK_SPD(5); break;
    case 6: LSK_SPD(6); break;
    default: return LSK_EUNSUPPORTED;
    }
#undef LSK_SPD
    return (int)cudaGetLastError();
}

extern "C" int lsk_spd_sample(const void* chol_upper, const void* shift_inv, const void* lmd, const void* coef, const void* w,
                              long long n_points, int m, int n, double scale, void* out, void* stream) {
    using namespace lsk;
    if (!chol_upper || !shift_inv || !lmd || !coef || !w || !out) return LSK_EARG;
    if (n_points < 0 || m < 0) return LSK_ESIZE;
    if (n_points == 0) return 0;
    dim3 block(64, 4), grid(1, (unsigned)((n_points + 3) / 4));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* u = static_cast<const double*>(chol_upper);
    const double* h = static_cast<const double*>(shift_inv);
    const double* l = static_cast<const double*>(lmd);
    const double* c = static_cast<const double*>(coef);
    const double* ww = static_cast<const double*>(w);
    double* o = static_cast<double*>(out);
#define LSK_SPD(N_) spd_features_kernel<N_, true><<<grid, block, 0, st>>>(u, h, l, c, ww, (int)n_points, m, scale, o, 0, 0, 0)
    switch (n) {
    case 2: LSK_SPD(2); break;
    case 3: LSK_SPD(3); break;
    case 4: LSK_SPD(4); break;
    case 5: LSK_SPD(5); break;
    case 6: LSK_SPD(6); break;
    default: return LSK_EUNSUPPORTED;
    }
#undef LSK_SPD
    return (int)cudaGetLastError();
}

extern "C" int lsk_spd_logdiag(const void* chol_upper, const void* shift_inv, long long n_points, int m, int n, void* out,
                               void* stream) {
    using namespace lsk;
    if (!chol_upper || !shift_inv || !out) return LSK_EARG;
    if (n_points < 0 || m < 0) return LSK_ESIZE;
    if (n_points == 0 || m == 0) return 0;
    dim3 block(64, 4), grid((unsigned)((m + 63) / 64), (unsigned)((n_points + 3) / 4));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* u = static_cast<const double*>(chol_upper);
    const double* h = static_cast<const double*>(shift_inv);
    double* o = static_cast<double*>(out);
#define LSK_SPD(N_) spd_logdiag_kernel<N_><<<grid, block, 0, st>>>(u, h, (int)n_points, m, o)
    switch (n) {
    case 2: LSK_SPD(2); break;
    case 3: LSK_SPD(3); break;
    case 4: LSK_SPD(4); break;
    case 5: LSK_SPD(5); break;
    case 6: LSK_SPD(6); break;
    default: return LSK_EUNSUPPORTED;
    }
#undef LSK_SPD
    return (int)cudaGetLastError();
}

extern "C" int lsk_small_matrix(const void* x, long long num, int n, int mode, void* out, void* bad_count, void* stream) {
    using namespace lsk;
    if (!x || !out || !bad_count) return LSK_EARG;
    if (num < 0 || n < 1 || n > 8 || (mode != 0 && mode != 1)) return LSK_ESIZE;
    if (num == 0) return 0;
    small_matrix_kernel<<<(unsigned)((num + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(x), num, n, mode, static_cast<double*>(out), static_cast<int*>(bad_count));
    return (int)cudaGetLastError();
}
