// fp64 pipe microbenchmark for B200 (sm_100a).
//
// Measures the three numbers every roofline fraction in this repo is quoted against:
//   * DFMA   : dependent-chain-free vector fp64 FMA throughput (TFLOP/s, 2 flop per FMA)
//   * DMMA   : mma.sync.aligned.m8n8k4.f64 throughput (TFLOP/s, 2*8*8*4 flop per warp instruction)
//   * MIXED  : both issued from the same warps, to see whether the two share one datapath
// plus a DFMA run whose second operand comes from the constant bank and one whose operand
// comes from a shared-memory broadcast, since those are the operand sources the fused
// character kernels use.
//
// Output: one JSON object on stdout (bench.py / profiles/ copy it into MEASURED fp64 peaks).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NACC  = 16;      // independent accumulators per thread

__constant__ double c_coef[1024];

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dfma_const(double* out, double a) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, c_coef[(it & 15) * 64 + j * NACC + i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// every DFMA takes one operand from a shared-memory broadcast load that feeds R entries
template <int R>
__global__ void __launch_bounds__(256) k_dfma_smem(double* out, double a) {
    __shared__ double tab[256];
    tab[threadIdx.x] = a * threadIdx.x;
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int i = 0; i < NACC; i += R) {
                double c = tab[(it + j * 16 + i) & 255];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[i + r] = fma(acc[i + r], a, c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int NMMA = 8;
__global__ void __launch_bounds__(256) k_dmma(double* out, double a, double b) {
    double d[NMMA][2];
#pragma unroll
    for (int i = 0; i < NMMA; ++i) { d[i][0] = a + i; d[i][1] = b + threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) dmma884(d[i][0], d[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NMMA; ++i) s += d[i][0] + d[i][1];
    if (s == 123.456) out[0] = s;
}

// NF DFMAs per DMMA in the same instruction stream
template <int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
    double d[NMMA][2];
    double acc[NMMA * NF];
#pragma unroll
    for (int i = 0; i < NMMA; ++i) { d[i][0] = a + i; d[i][1] = b + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < NMMA * NF; ++i) acc[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
            dmma884(d[i][0], d[i][1], a, b);
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[i * NF + f] = fma(acc[i * NF + f], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NMMA; ++i) s += d[i][0] + d[i][1];
#pragma unroll
    for (int i = 0; i < NMMA * NF; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    double h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0 / (i + 1);
    CK(cudaMemcpyToSymbol(c_coef, h, sizeof(h)));
    const int blocks = sms * 8, threads = 256;
    const double nthreads = double(blocks) * threads;
    const double a = 1.0000001, b = 0.9999999;

    double ms_fma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, a, b); }, 10);
    double ms_fmc = time_ms([&] { k_dfma_const<<<blocks, threads>>>(out, a); }, 10);
    double ms_s1  = time_ms([&] { k_dfma_smem<1><<<blocks, threads>>>(out, a); }, 10);
    double ms_s2  = time_ms([&] { k_dfma_smem<2><<<blocks, threads>>>(out, a); }, 10);
    double ms_s4  = time_ms([&] { k_dfma_smem<4><<<blocks, threads>>>(out, a); }, 10);
    double ms_mma = time_ms([&] { k_dmma<<<blocks, threads>>>(out, a, b); }, 10);
    double ms_m1  = time_ms([&] { k_mixed<1><<<blocks, threads>>>(out, a, b); }, 10);
    double ms_m2  = time_ms([&] { k_mixed<2><<<blocks, threads>>>(out, a, b); }, 10);
    double ms_m4  = time_ms([&] { k_mixed<4><<<blocks, threads>>>(out, a, b); }, 10);
    double ms_m8  = time_ms([&] { k_mixed<8><<<blocks, threads>>>(out, a, b); }, 10);
    CK(cudaGetLastError());

    const double fl_fma = nthreads * ITERS * NACC * 2.0;
    const double fl_mma = (nthreads / 32) * ITERS * NMMA * (2.0 * 8 * 8 * 4);
    auto tf = [](double fl, double ms) { return fl / (ms * 1e-3) / 1e12; };
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, sms, clk);
    printf(" \"dfma_tflops\": %.3f, \"dfma_const_operand_tflops\": %.3f,\n", tf(fl_fma, ms_fma), tf(fl_fma, ms_fmc));
    printf(" \"dfma_smem_operand_tflops\": {\"1\": %.3f, \"2\": %.3f, \"4\": %.3f},\n",
           tf(fl_fma, ms_s1), tf(fl_fma, ms_s2), tf(fl_fma, ms_s4));
    printf(" \"dmma_m8n8k4_tflops\": %.3f,\n", tf(fl_mma, ms_mma));
    printf(" \"mixed\": {");
    const int nf[4] = {1, 2, 4, 8}; const double msm[4] = {ms_m1, ms_m2, ms_m4, ms_m8};
    for (int i = 0; i < 4; ++i) {
        double fl_f = nthreads * ITERS * NMMA * nf[i] * 2.0;
        printf("\"dfma_per_dmma_%d\": {\"total_tflops\": %.3f, \"dfma_part\": %.3f, \"dmma_part\": %.3f}%s",
               nf[i], tf(fl_f + fl_mma, msm[i]), tf(fl_f, msm[i]), tf(fl_mma, msm[i]), i < 3 ? ", " : "");
    }
    printf("},\n \"nominal_dfma_tflops_at_max_clock\": %.3f}\n", sms * 64 * 2.0 * clk * 1e3 / 1e12);
    return 0;
}
