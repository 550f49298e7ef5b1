// Prior samples f = Phi(x) w without a Gram launch.
//
//   lsk_panel_gemv:  f[i] = scale * sum_c Phi[i, c] w[c] for a feature map that already sits in the panel-major operand layout
//                    (RandomPhaseApproximation.forward, prior_approximation.py:85-88: einsum('nm,m->n')).  HBM-bound: every
//                    entry of Phi is read once (the 128 x 64 tile kernel with a one-row right operand computes 64 columns to
//                    keep one).  Deterministic: fixed summation order, no atomics.
#include "lsk_common.cuh"

namespace lsk {

// one CTA per 64-row panel; thread = (row r, slice q of the k-groups: g = q, q + 4, ...)
__global__ void __launch_bounds__(256)
panel_gemv_kernel(const double* __restrict__ A, long long n_rows, int kg_total, const double* __restrict__ w, double scale,
                  double* __restrict__ out)
{
    __shared__ double part[4][64];
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    const double* base = A + ((size_t)blockIdx.x * kg_total) * (PANEL * 4) + r * 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int g = q;
#pragma unroll 1
    for (; g + 12 < kg_total; g += 16) {                       // four 32-byte loads in flight per thread
        double2 v[4][2], ww[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double2* p = reinterpret_cast<const double2*>(base + (size_t)(g + 4 * u) * (PANEL * 4));
            v[u][0] = __ldcs(p); v[u][1] = __ldcs(p + 1);
            const double2* pw = reinterpret_cast<const double2*>(w + 4 * (g + 4 * u));
            ww[u][0] = __ldg(pw); ww[u][1] = __ldg(pw + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 = fma(v[u][0].x, ww[u][0].x, a0); a1 = fma(v[u][0].y, ww[u][0].y, a1);
            a2 = fma(v[u][1].x, ww[u][1].x, a2); a3 = fma(v[u][1].y, ww[u][1].y, a3);
        }
    }
    for (; g < kg_total; g += 4) {
        const double2* p = reinterpret_cast<const double2*>(base + (size_t)g * (PANEL * 4));
        const double2 v0 = __ldcs(p), v1 = __ldcs(p + 1);
        const double2* pw = reinterpret_cast<const double2*>(w + 4 * g);
        const double2 w0 = __ldg(pw), w1 = __ldg(pw + 1);
        a0 = fma(v0.x, w0.x, a0); a1 = fma(v0.y, w0.y, a1);
        a2 = fma(v1.x, w1.x, a2); a3 = fma(v1.y, w1.y, a3);
    }
    part[q][r] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    const long long row = 64ll * blockIdx.x + r;
    if (q == 0 && row < n_rows) out[row] = scale * ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r]));
}

}  // namespace lsk

// A: panel-major [n_rows (padded to 128), 4 * kg_total]; w: device, 4 * kg_total doubles (zero-padded); out: n_rows doubles
extern "C" int lsk_panel_gemv(const void* A, long long n_rows, int kg_total, const void* w, double scale, void* out, void* stream) {
    using namespace lsk;
    if (!A || !w || !out) return LSK_EARG;
    if (n_rows < 0 || kg_total < 1) return LSK_ESIZE;
    if (n_rows == 0) return 0;
    const long long panels = (n_rows + 63) / 64;
    if (panels > 0x7fffffffLL) return LSK_ESIZE;
    panel_gemv_kernel<<<(unsigned)panels, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const double*>(A), n_rows, kg_total, static_cast<const double*>(w), scale, static_cast<double*>(out));
    return (int)cudaGetLastError();
}
