// Feature map of the flat torus T^n = [0,1)^n  (reference spaces/torus.py:14-97).
//
// The characters are chi_k(x) = exp(2 pi i k.x), so the covariance  sum_k a_k Re chi_k(x_i - y_j)  is the Gram of
//     F[i, 2k] = w_k cos(2 pi k.x_i),   F[i, 2k+1] = w_k sin(2 pi k.x_i)
// (weights on one side only), and the N x M work goes to the fp64 tensor-core Gram kernel (lsk_pairwise_poly, LINEAR).
// This kernel writes F in the panel-major operand layout that kernel reads: [row / 64][col / 4][row % 64][col % 4],
// rows zero-padded to a multiple of 128.  `two_pi` is passed by the host: the reference evaluates its characters with
// pi = 2 * acos(0) rounded to float32 (torus.py:11), and parity at 1e-10 needs the same constant.
#include "lsk_common.cuh"

namespace lsk {

__global__ void __launch_bounds__(256)
torus_features_kernel(const double* __restrict__ x, long long num, int dim, const double* __restrict__ freq,
                      const double* __restrict__ weight, int order, double two_pi, double* __restrict__ out, int out_kg) {
    pdl_wait();
    const long long rows_padded = ((num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
    const long long total = rows_padded * out_kg;                 // one thread per (row, k-group) = two frequencies
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx & (PANEL - 1));
        const long long pg = idx >> 6;                             // panel * out_kg + group
        const int g = (int)(pg % out_kg);
        const long long row = (pg / out_kg) * PANEL + r;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (row < num) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 2 * g + h;
                if (k < order) {
                    double t = 0.0;
                    for (int d = 0; d < dim; ++d) t += x[row * dim + d] * freq[(size_t)k * dim + d];
                    double s, c;
                    sincos(two_pi * t, &s, &c);
                    const double w = weight ? weight[k] : 1.0;
                    v[2 * h] = w * c;
                    v[2 * h + 1] = w * s;
                }
            }
        }
        double2* dst = reinterpret_cast<double2*>(out + (size_t)pg * (PANEL * 4) + (size_t)r * 4);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
    }
}

}  // namespace lsk

extern "C" int lsk_torus_features(const void* x, long long num, int dim, const void* freq, const void* weight, int order,
                                  double two_pi, void* out, int out_kg, void* stream) {
    using namespace lsk;
    if (!x || !freq || !out) return LSK_EARG;
    if (num < 0 || dim < 1 || order < 1 || out_kg < (2 * order + 3) / 4) return LSK_ESIZE;
    if (num == 0) return 0;
    const long long rows_padded = ((num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
    long long blocks = (rows_padded * out_kg + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    return (int)launch_pdl(torus_features_kernel, dim3((unsigned)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                           static_cast<const double*>(x), num, dim, static_cast<const double*>(freq),
                           static_cast<const double*>(weight), order, two_pi, static_cast<double*>(out), out_kg);
}
