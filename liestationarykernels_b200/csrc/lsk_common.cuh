// Shared device helpers for liblsk (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define LSK_VERSION 100

// error codes of the C ABI: 0 ok, <0 argument error, >0 cudaError_t
#define LSK_EARG  (-1)
#define LSK_ESIZE (-2)
#define LSK_EUNSUPPORTED (-3)

namespace lsk {

constexpr int PANEL = 64;       // rows per panel of the embedded operand layout
constexpr int BM = 128;         // rows per CTA tile  (two A panels)
constexpr int BN = 64;          // columns per CTA tile (one B panel)
constexpr int ROW_PAD = 128;    // embedded operands are zero-padded to a multiple of this many rows
constexpr int KG_PER_STAGE = 8; // k-groups (of 4 doubles) per pipeline stage -> 16 KB per panel per stage
constexpr int STAGES = 4;       // 4 x (2 A panels + 1 B panel) x 16 KB = 192 KB: a whole K=128 tile is prefetched
constexpr int CONSUMER_WARPS = 16;
constexpr int THREADS = CONSUMER_WARPS * 32;         // the TMA producer is lane 0 of warp 0

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LSK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LSK_DONE_%=;\n"
        "bra LSK_WAIT_%=;\n"
        "LSK_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// single hardware-suspending wait attempt (returns after the phase completes or after the HW time limit)
__device__ __forceinline__ bool mbar_try_wait_once(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine; SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col); fp64 tensor-core path (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// shared-memory loads through an explicit 32-bit shared-window address (computed once per kernel): keeps the
// generic->shared conversion (an S2UR of the CTA id on sm_100a) out of inner loops
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void st_cs_f64x2(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_cs_f64(double* p, double a) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}

// programmatic dependent launch (PDL): a kernel launched with `launch_pdl` may start while the previous kernel of the
// stream is still draining; it must execute pdl_wait() before its first global-memory access (the wait returns once the
// previous grid has completed and its writes are visible).  pdl_launch_dependents() lets the NEXT kernel start early.
// Both are no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace lsk
