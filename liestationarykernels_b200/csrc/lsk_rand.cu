// Haar sampling on SO(n) / SU(n) from Gaussian draws — the deterministic part of `space.rand`
// (spaces/so.py:62-100, spaces/su.py:45-64; also SPD.rand_phase, spaces/spd.py:48-55).  The Gaussian source stays
// `torch.randn` so that seeded scripts consume the generator exactly like the reference; this file replaces the batched
// cuSOLVER QR + sign / determinant fixes:
//     h = randn(num, n, n);  q, r = qr(h);  q *= sign(diag r)  [SU: conj phase];  q[:, :, 0] *= sign(det q)  [SU: conj phase]
// with one thread per matrix doing a Householder QR with LAPACK's conventions (geqr2 / larfg, org2r).
// n = 3 (SO) and n = 2 (SU) use the reference's unit-quaternion parametrisation.
#include "lsk_common.cuh"

namespace lsk {

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.re, -a.im}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// real and complex share the code path (imaginary parts are zero for SO(n))
template <bool CPLX>
__global__ void __launch_bounds__(64)
haar_qr_kernel(const double* __restrict__ gauss, long long num, int n, double* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    constexpr int S = CPLX ? 2 : 1;
    cplx a[8][8], q[8][8], tau[8];
    double rdiag[8];
    const double* g = gauss + b * n * n * S;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) a[r][c] = {g[(r * n + c) * S], CPLX ? g[(r * n + c) * S + 1] : 0.0};
    // ---- Householder QR: H_j = I - tau_j v_j v_j^H, v_j(j) = 1, beta_j = -sign(Re alpha) * ||(alpha, x)|| ----
    for (int j = 0; j < n; ++j) {
        const cplx alpha = a[j][j];
        double ss = 0.0;
        for (int r = j + 1; r < n; ++r) ss += a[r][j].re * a[r][j].re + a[r][j].im * a[r][j].im;
        if (ss == 0.0 && alpha.im == 0.0) {
            tau[j] = {0.0, 0.0};
            rdiag[j] = alpha.re;
            continue;
        }
        const double nrm = sqrt(alpha.re * alpha.re + alpha.im * alpha.im + ss);
        const double beta = alpha.re >= 0.0 ? -nrm : nrm;
        tau[j] = {(beta - alpha.re) / beta, -alpha.im / beta};
        const cplx sc = cdiv({1.0, 0.0}, {alpha.re - beta, alpha.im});
        for (int r = j + 1; r < n; ++r) a[r][j] = cmul(a[r][j], sc);
        rdiag[j] = beta;
        // apply H_j^H to the trailing columns (zgeqr2 applies conj(tau) from the left)
        const cplx tc = cconj(tau[j]);
        for (int c = j + 1; c < n; ++c) {
            cplx w = a[j][c];
            for (int r = j + 1; r < n; ++r) w = cadd(w, cmul(cconj(a[r][j]), a[r][c]));
            w = cmul(tc, w);
            a[j][c] = csub(a[j][c], w);
            for (int r = j + 1; r < n; ++r) a[r][c] = csub(a[r][c], cmul(a[r][j], w));
        }
    }
    // ---- Q = H_0 H_1 ... H_{n-1} applied to the identity, last reflector first (org2r / ung2r) ----
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) q[r][c] = {r == c ? 1.0 : 0.0, 0.0};
    for (int j = n - 1; j >= 0; --j) {
        if (tau[j].re == 0.0 && tau[j].im == 0.0) continue;
        for (int c = j; c < n; ++c) {
            cplx w = q[j][c];
            for (int r = j + 1; r < n; ++r) w = cadd(w, cmul(cconj(a[r][j]), q[r][c]));
            w = cmul(tau[j], w);
            q[j][c] = csub(q[j][c], w);
            for (int r = j + 1; r < n; ++r) q[r][c] = csub(q[r][c], cmul(a[r][j], w));
        }
    }
    // ---- q *= sign(diag r) (SO) / conj(r_jj / |r_jj|) (SU): the diagonal of R is real in both cases ----
    for (int c = 0; c < n; ++c) {
        const double sg = rdiag[c] > 0.0 ? 1.0 : (rdiag[c] < 0.0 ? -1.0 : 0.0);
        for (int r = 0; r < n; ++r) { q[r][c].re *= sg; q[r][c].im *= sg; }
    }
    // ---- determinant by LU with partial pivoting; column 0 is multiplied by the conjugate phase of det ----
    cplx lu[8][8];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) lu[r][c] = q[r][c];
    cplx det = {1.0, 0.0};
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = lu[c][c].re * lu[c][c].re + lu[c][c].im * lu[c][c].im;
        for (int r = c + 1; r < n; ++r) {
            const double m = lu[r][c].re * lu[r][c].re + lu[r][c].im * lu[r][c].im;
            if (m > best) { best = m; piv = r; }
        }
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const cplx t = lu[c][k]; lu[c][k] = lu[piv][k]; lu[piv][k] = t; }
            det = {-det.re, -det.im};
        }
        det = cmul(det, lu[c][c]);
        if (best == 0.0) break;
        for (int r = c + 1; r < n; ++r) {
            const cplx f = cdiv(lu[r][c], lu[c][c]);
            for (int k = c + 1; k < n; ++k) lu[r][k] = csub(lu[r][k], cmul(f, lu[c][k]));
        }
    }
    cplx fix;
    if (CPLX) {
        const double m = sqrt(det.re * det.re + det.im * det.im);
        fix = {det.re / m, -det.im / m};
    } else {
        fix = {det.re > 0.0 ? 1.0 : (det.re < 0.0 ? -1.0 : 0.0), 0.0};
    }
    for (int r = 0; r < n; ++r) q[r][0] = cmul(q[r][0], fix);
    double* o = out + b * n * n * S;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            o[(r * n + c) * S] = q[r][c].re;
            if (CPLX) o[(r * n + c) * S + 1] = q[r][c].im;
        }
}

// unit quaternion -> SO(3) (so.py:74-92) or SU(2) (su.py:47-55); gauss: [num, 4]
__global__ void __launch_bounds__(256)
quaternion_kernel(const double* __restrict__ gauss, long long num, int to_su2, double* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    double x = gauss[4 * b], y = gauss[4 * b + 1], z = gauss[4 * b + 2], w = gauss[4 * b + 3];
    const double nrm = sqrt(x * x + y * y + z * z + w * w);        // torch.linalg.vector_norm
    x /= nrm; y /= nrm; z /= nrm; w /= nrm;
    if (to_su2) {
        // a = x + i y, b = z + i w;  q = [[a, b], [-conj b, conj a]]  (su.py:50-55: r1 = (a, -b*) and r2 = (b, a*) are COLUMNS)
        double* o = out + 8 * b;
        o[0] = x;  o[1] = y;   o[2] = z;  o[3] = w;
        o[4] = -z; o[5] = w;   o[6] = x;  o[7] = -y;
        return;
    }
    const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, xw = x * w, yz = y * z, yw = y * w, zw = z * w;
    double* o = out + 9 * b;
    // r1, r2, r3 of the reference are COLUMNS (cat along the last axis)
    o[0] = 1 - 2 * (yy + zz); o[1] = 2 * (xy + zw);     o[2] = 2 * (xz - yw);
    o[3] = 2 * (xy - zw);     o[4] = 1 - 2 * (xx + zz); o[5] = 2 * (yz + xw);
    o[6] = 2 * (xz + yw);     o[7] = 2 * (yz - xw);     o[8] = 1 - 2 * (xx + yy);
}

}  // namespace lsk

extern "C" int lsk_haar_from_gaussian(const void* gauss, long long num, int n, int is_complex, void* out, void* stream) {
    using namespace lsk;
    if (!gauss || !out) return LSK_EARG;
    if (num < 0 || n < 1 || n > 8) return LSK_ESIZE;
    if (num == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((num + 63) / 64);
    if (is_complex) haar_qr_kernel<true><<<blocks, 64, 0, st>>>(static_cast<const double*>(gauss), num, n, static_cast<double*>(out));
    else haar_qr_kernel<false><<<blocks, 64, 0, st>>>(static_cast<const double*>(gauss), num, n, static_cast<double*>(out));
    return (int)cudaGetLastError();
}

extern "C" int lsk_quaternion_to_group(const void* gauss, long long num, int to_su2, void* out, void* stream) {
    using namespace lsk;
    if (!gauss || !out) return LSK_EARG;
    if (num < 0) return LSK_ESIZE;
    if (num == 0) return 0;
    quaternion_kernel<<<(unsigned)((num + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(gauss), num, to_su2, static_cast<double*>(out));
    return (int)cudaGetLastError();
}
