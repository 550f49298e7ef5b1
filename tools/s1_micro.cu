// Development: where do the cycles of the 16 x 16 register-resident elimination (diag_tile_kernel, step S1) go?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/s1_micro tools/s1_micro.cu && tools/s1_micro
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    const double e2 = fma(-x * y, y, 1.0);
    return fma(0.5 * y, e2, y);
}
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st_pred(double* p, double v, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f64 [%0], %1; }" :: "l"(p), "d"(v), "r"((int)pred) : "memory");
}
constexpr int TS = 18;

// EPI 0: branchy epilogue (as in the kernel), 1: predicated stores;   OTHER 0: other warps idle at the barrier, 1: busy with DMMA
template <int EPI, int OTHER>
__global__ void __launch_bounds__(256, 1) k(const double* __restrict__ D, double* __restrict__ gout, long long* cycles, int nb)
{
    extern __shared__ __align__(128) double sm[];
    double* Dt = sm;                 // 16 x TS
    double* Bt = sm + 16 * TS;
    double* rsv = Bt + 16 * TS;      // 16
    double* piv = rsv + 16;          // 64
    double* junk = piv + 64;         // 8 x 16 x TS for the other warps
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t < 256) Dt[(t >> 4) * TS + (t & 15)] = D[t];
    for (int i = t; i < 8 * 16 * TS; i += 256) junk[i] = 1e-3 * i;
    __syncthreads();
    if (warp == 0) {
        const int c = lane & 15, o = 0, k0 = 0;
        const bool iside = lane >= 16;
        double* mine = (iside ? Bt : Dt) + c;
        double w[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) { const double x = Dt[r * TS + c]; w[r] = iside ? ((r == c) ? 1.0 : 0.0) : x; }
        double dc = 1.0; int bad = -1;
        const long long t0 = clk();
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {
            double* pb = piv + (j & 1) * 32;
            pb[(iside ? 16 : 0) + ((c - j) & 15)] = w[0];
            mine[j * TS] = w[0];
            __syncwarp();
            double row[16];
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const double2 x = *reinterpret_cast<const double2*>(pb + i); row[i] = x.x; row[i + 1] = x.y; }
            const double d = row[0];
            const double id = fast_rcp(d);
            dc = (c == j) ? d : dc;
            bad = (bad < 0 && !(d > 0.0) && o + j < nb) ? j : bad;
            const double pc = w[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) w[i - 1] = fma(-(row[i] * pc), id, w[i]);
            w[15] = 0.0;
        }
        const long long t1 = clk();
        if (bad >= 0 && lane == 0) gout[4000] = bad;
        const double rs = fast_rsqrt(dc);
        if (EPI == 0) { if (!iside) rsv[o + c] = rs; } else rsv[o + c] = rs;
        __syncwarp();
        double* gdst = iside ? gout + 2048 + c : gout + (size_t)(k0 + o) * 128 + k0 + o + c;
        const long long gstep = iside ? 4 : 128;
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const double u = mine[r * TS] * rsv[o + r];
            mine[r * TS] = (iside || r <= c) ? u : 0.0;
            const bool pred = iside ? (r >= c) : (r <= c && o + c < nb);
            if (EPI == 0) { if (pred) gdst[r * gstep] = u; } else st_pred(gdst + r * gstep, u, pred);
        }
        const long long t2 = clk();
        if (lane == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; }
    } else if (OTHER == 1) {
        const int g = lane >> 2, tig = lane & 3;
        double* P = junk + (warp - 1) * 16 * TS;
        double acc[2][2][2] = {};
        for (int rep = 0; rep < 12; ++rep)
#pragma unroll
            for (int kk = 0; kk < 16; kk += 4) {
                const double a0 = P[(kk + tig) * TS + g], a1 = P[(kk + tig) * TS + 8 + g];
                dmma884(acc[0][0], a0, a0); dmma884(acc[0][1], a0, a1); dmma884(acc[1][0], a1, a0); dmma884(acc[1][1], a1, a1);
            }
        if (acc[0][0][0] == 12345.0) gout[5000] = acc[1][1][1];
    }
    __syncthreads();
}

template <int EPI, int OTHER>
void run(const double* D, double* out, long long* cyc, const char* name) {
    long long c[2];
    const int smem = (2 * 16 * TS + 16 + 64 + 8 * 16 * TS) * 8;
    for (int rep = 0; rep < 3; ++rep) { k<EPI, OTHER><<<1, 256, smem>>>(D, out, cyc, 128); cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost); }
    printf("%-44s 16 steps %5lld cycles, epilogue %5lld   (%s)\n", name, c[0], c[1], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double h[256];
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) h[r * 16 + c] = (r == c ? 20.0 : 0.0) + 1.0 / (1 + r + c);
    double *D, *out; long long* cyc;
    cudaMalloc(&D, sizeof(h)); cudaMalloc(&out, 8192 * 8); cudaMalloc(&cyc, 16);
    cudaMemcpy(D, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0, 0>(D, out, cyc, "branchy epilogue, other warps idle");
    run<1, 0>(D, out, cyc, "predicated epilogue, other warps idle");
    run<0, 1>(D, out, cyc, "branchy epilogue, other warps on DMMA");
    run<1, 1>(D, out, cyc, "predicated epilogue, other warps on DMMA");
    return 0;
}
