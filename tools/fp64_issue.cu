// Issue-slot model of the B200 fp64 pipe: how much do non-DFMA instructions cost next to DFMA?
//
// Variants (all 16 warps/SMSP resident, independent accumulator chains):
//   base        : DFMA only
//   ffma_K      : K independent FFMA per DFMA        (fma pipe, fp32)
//   iadd_K      : K independent IADD3/LOP per DFMA   (alu pipe)
//   lds128_per_N: one broadcast LDS.128 per N DFMA, both doubles used as DFMA operands
//   cparam      : DFMA operand read straight from the kernel-parameter constant bank, immediate offsets
//                 (fully unrolled Horner over R entries, coefficient table passed by value)
// Output: JSON with the DFMA TFLOP/s each variant sustains (2 flop per DFMA, other work not counted).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 2048;
constexpr int NACC = 8;

template <int KF, int KI>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b, float fa, int ia) {
    double acc[NACC];
    float f[NACC * (KF > 0 ? KF : 1)];
    int n[NACC * (KI > 0 ? KI : 1)];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = a + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NACC * (KF > 0 ? KF : 1); ++i) f[i] = fa + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NACC * (KI > 0 ? KI : 1); ++i) n[i] = ia + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            acc[i] = fma(acc[i], b, a);
#pragma unroll
            for (int k = 0; k < KF; ++k) f[i * KF + k] = fmaf(f[i * KF + k], fa, 1.0f);
#pragma unroll
            for (int k = 0; k < KI; ++k) n[i * KI + k] = (n[i * KI + k] ^ ia) + it;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    for (int i = 0; i < NACC * (KF > 0 ? KF : 1); ++i) s += f[i];
    for (int i = 0; i < NACC * (KI > 0 ? KI : 1); ++i) s += n[i];
    if (s == 123.456) out[0] = s;
}

// one LDS.128 (broadcast) per 2*R DFMA: R entries share each pair of coefficients
template <int R>
__global__ void __launch_bounds__(256) k_lds128(double* out, double a) {
    __shared__ double2 tab[128];
    if (threadIdx.x < 128) tab[threadIdx.x] = make_double2(a * threadIdx.x, a + threadIdx.x);
    __syncthreads();
    double acc[R], x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r] = a + r + threadIdx.x; x[r] = a * (r + 1); }
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            double2 c = tab[j + (it & 1) * 64];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(acc[r], x[r], c.x);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(acc[r], x[r], c.y);
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    if (s == 123.456) out[0] = s;
}

struct Coefs { double c[128]; };

template <int R>
__global__ void __launch_bounds__(256) k_cparam(double* out, double a, const __grid_constant__ Coefs p) {
    double acc[R], x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r] = a + r + threadIdx.x; x[r] = a * (r + 1); }
    for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
        for (int j = 0; j < 128; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(acc[r], x[r], p.c[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 8; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    const double nth = double(blocks) * threads;
    double* out; CK(cudaMalloc(&out, 8));
    const double a = 1.0000001, b = 0.9999999;
    Coefs co; for (int i = 0; i < 128; ++i) co.c[i] = 1.0 / (i + 1);
    auto tf = [&](double dfma_per_thread, double ms) { return nth * dfma_per_thread * 2.0 / (ms * 1e-3) / 1e12; };
    const double n_mix = double(ITERS) * NACC;
    printf("{\"base\": %.2f", tf(n_mix, time_ms([&] { k_mix<0, 0><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"ffma_1\": %.2f", tf(n_mix, time_ms([&] { k_mix<1, 0><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"ffma_2\": %.2f", tf(n_mix, time_ms([&] { k_mix<2, 0><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"ffma_4\": %.2f", tf(n_mix, time_ms([&] { k_mix<4, 0><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"iadd_1\": %.2f", tf(n_mix, time_ms([&] { k_mix<0, 1><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"iadd_2\": %.2f", tf(n_mix, time_ms([&] { k_mix<0, 2><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    printf(", \"ffma1_iadd1\": %.2f", tf(n_mix, time_ms([&] { k_mix<1, 1><<<blocks, threads>>>(out, a, b, 1.0f, 3); })));
    const double n_l = double(ITERS / 8) * 64 * 2;
    printf(",\n \"lds128_per_2dfma\": %.2f", tf(n_l * 1, time_ms([&] { k_lds128<1><<<blocks, threads>>>(out, a); })));
    printf(", \"lds128_per_4dfma\": %.2f", tf(n_l * 2, time_ms([&] { k_lds128<2><<<blocks, threads>>>(out, a); })));
    printf(", \"lds128_per_8dfma\": %.2f", tf(n_l * 4, time_ms([&] { k_lds128<4><<<blocks, threads>>>(out, a); })));
    printf(", \"lds128_per_16dfma\": %.2f", tf(n_l * 8, time_ms([&] { k_lds128<8><<<blocks, threads>>>(out, a); })));
    const double n_c = double(ITERS / 16) * 128;
    printf(",\n \"cparam_R1\": %.2f", tf(n_c * 1, time_ms([&] { k_cparam<1><<<blocks, threads>>>(out, a, co); })));
    printf(", \"cparam_R2\": %.2f", tf(n_c * 2, time_ms([&] { k_cparam<2><<<blocks, threads>>>(out, a, co); })));
    printf(", \"cparam_R4\": %.2f", tf(n_c * 4, time_ms([&] { k_cparam<4><<<blocks, threads>>>(out, a, co); })));
    printf(", \"cparam_R8\": %.2f}\n", tf(n_c * 8, time_ms([&] { k_cparam<8><<<blocks, threads>>>(out, a, co); })));
    CK(cudaGetLastError());
    return 0;
}
