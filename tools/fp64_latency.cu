// DFMA / DMMA dependent-issue latency and the in-flight count needed to saturate the B200 fp64 pipe.
// One CTA per SM; W warps per SM sub-partition (blockDim = 128 * W); C independent accumulator chains per thread.
// Reports cycles per DFMA (per SMSP) for every (W, C): throughput saturates at 2 cycles; below saturation
// cycles/DFMA = latency / (W * C).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 8192;

template <int C>
__global__ void k_chain(double* out, double a, double b, long long* cycles) {
    double acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = a + i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) acc[i] = fma(acc[i], b, a);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int C>
__global__ void k_mma_chain(double* out, double a, double b, long long* cycles) {
    double d[C][2];
#pragma unroll
    for (int i = 0; i < C; ++i) { d[i][0] = a + i; d[i][1] = b + threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += d[i][0] + d[i][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int C>
static void run(int W, double* out, long long* dcyc, bool mma) {
    long long h = 0;
    if (mma) k_mma_chain<C><<<148, 128 * W>>>(out, 1.0000001, 0.9999999, dcyc);
    else k_chain<C><<<148, 128 * W>>>(out, 1.0000001, 0.9999999, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("  {\"op\": \"%s\", \"warps_per_smsp\": %d, \"chains\": %d, \"cycles_per_op_per_smsp\": %.2f},\n",
           mma ? "dmma" : "dfma", W, C, double(h) / (double(ITERS) * C * W));
}

int main() {
    double* out; long long* dcyc;
    CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&dcyc, 8));
    printf("[\n");
    for (int W : {1, 2, 4, 8}) {
        run<1>(W, out, dcyc, false); run<2>(W, out, dcyc, false); run<4>(W, out, dcyc, false);
        run<8>(W, out, dcyc, false); run<16>(W, out, dcyc, false);
    }
    for (int W : {1, 2, 4}) {
        run<1>(W, out, dcyc, true); run<2>(W, out, dcyc, true); run<4>(W, out, dcyc, true); run<8>(W, out, dcyc, true);
    }
    printf("  {}\n]\n");
    return 0;
}
