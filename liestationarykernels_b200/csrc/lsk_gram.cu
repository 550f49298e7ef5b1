// Dedicated Gram kernel:  out = scale * A B^T  for panel-major feature maps (EPI_LINEAR of lsk_pairwise_poly with a row-major
// output) - the dense contraction behind RandomPhaseKernel / RandomFourierFeatureKernel `.evaluate()` and the samplers' `_cov`
// (reference: gpytorch MatmulLazyTensor of spectral_kernel.py:125-127, 195-197, i.e. a DGEMM).  At C4 / C5 this one kernel is
// 98 % of the step (K = 20 480 / 16 384, 4.4e13 / 1.4e14 flop).
//
// Same operand layout and DMMA fragment mapping as the generic tile kernel (lsk_pairwise.cu); what differs is everything
// around the tensor-core instructions, because at large K the only thing that matters is that the fp64 pipe never drains:
//   * a DEDICATED producer warp (17th warp, one lane) owns the TMA ring and the tile order: the consumer warps carry no
//     producer duty, no shared chunk counter, no compare-and-swap - a stage boundary is one try_wait and one arrive;
//   * the consumer loop is SOFTWARE-PIPELINED ACROSS STAGE BOUNDARIES: the fragments of k-group g+1 are loaded (double-buffered
//     registers) before the DMMAs of k-group g are issued, and at the last k-group of a stage the wait on the NEXT stage's full
//     barrier and the loads of its first fragments come before the last DMMAs of the current stage.  The four warps of an SM
//     sub-partition run in lockstep (they share the pipe fairly), so any instruction sequence that has to complete between
//     two DMMAs of one warp is a bubble in all four; with the pipelining the DMMA stream of a warp has no such gap;
//   * tiles are handed out in column bands of ~1536 columns, row-major inside a band, so the CTAs of one wave share
//     ~12 x 12 tiles worth of operand panels in L2 instead of 1 x 148 (the generic order re-reads B from HBM for every row tile:
//     ncu at 8192^2 x 16384: 63.6 GB of DRAM reads for 2.1 GB of operands, profiles/gram_ncu_summary_r02.json);
//   * NT = 4 instantiation: 128 x 128 tiles (16 flop per byte staged instead of 10.7, 0.5 fragment loads per DMMA instead of 0.75);
//     its consumers need ~110 registers, which 17 warps do not leave (96): it runs with a producer WARPGROUP and `setmaxnreg`;
//   * symmetric mode (A and B are the same points): only tiles that reach the diagonal or lie above it, mirrored on store.
#include <cstdlib>
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"

namespace lsk {
namespace gram {

constexpr int KGS = 8;                              // k-groups (of 4 doubles) per pipeline stage
constexpr int PANEL_STAGE = KGS * PANEL * 4;        // one 64-row panel, one stage: 2048 doubles = 16 KB
constexpr int CONS_WARPS = 16;
constexpr int THREADS_G = (CONS_WARPS + 1) * 32;    // + the producer warp
// 128 x 128 tiles need ~110 registers per consumer thread; 17 warps leave 96 (five warps on one sub-partition).  That variant runs
// with a whole producer WARPGROUP (20 warps, 96 registers each at launch) and moves registers from it to the consumers:
// 4 x 32 x 24 + 16 x 32 x 112 = 60 416 <= 640 x 96.
constexpr int THREADS_G4 = (CONS_WARPS + 4) * 32;
constexpr int CONS_REGS_G4 = 112, PROD_REGS_G4 = 24;
constexpr int TILE_SLOTS = 8;                       // ring of tile descriptors (producer -> consumers)

template <int NT> struct Cfg {
    static constexpr int BNT = 32 * NT;             // columns per CTA tile: 64 (NT = 2) or 128 (NT = 4)
    static constexpr int B_PANELS = NT / 2;
    static constexpr int STAGE_DBL = (2 + B_PANELS) * PANEL_STAGE;
    static constexpr int NST = NT == 2 ? 4 : 3;     // 4 x 48 KB or 3 x 64 KB = 192 KB
    static constexpr int RATIO = 128 / BNT;         // column tiles per row tile
    static constexpr int BAND = 1536 / BNT;         // column tiles per band of the tile order
    static constexpr int SMEM_BYTES = NST * STAGE_DBL * 8 + 2 * NST * 8 + TILE_SLOTS * 8 + 64;
};

// number of tiles of the (possibly triangular) schedule; also used by the host for the grid size
template <int NT, bool SYM>
__host__ __device__ inline long long count_tiles(int tiles_r, int tiles_c) {
    if (!SYM) return (long long)tiles_r * tiles_c;
    long long n = 0;
    for (int r = 0; r < tiles_r; ++r) { const int first = Cfg<NT>::RATIO * r; if (first < tiles_c) n += tiles_c - first; }
    return n;
}

// MODE 0: out = scale * A B^T (SYM: upper triangle of tiles, mirrored);  MODE 1: out += scale * A B^T in place (SYM: the tiles
// that reach the diagonal or lie above it, nothing mirrored) - the rank-k update of the blocked Cholesky (EPI_SYRK).  The old
// entries of a tile are requested when the tile starts and consumed after its DMMA phase.
template <int NT, bool SYM, int MODE>
__global__ void __launch_bounds__((NT == 4 || MODE == 1) ? THREADS_G4 : THREADS_G, 1)
gram_kernel(const double* __restrict__ A, const double* __restrict__ B, int n_rows, int n_cols, int kg_stride, int kg_len,
            double* __restrict__ out, long long ld_out, double scale, long long n_tiles)
{
    using C = Cfg<NT>;
    constexpr int NST = C::NST, STAGE_DBL = C::STAGE_DBL, BNT = C::BNT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sbuf = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + NST * STAGE_DBL);
    uint64_t* empty_bar = full_bar + NST;
    int2* tile_desc = reinterpret_cast<int2*>(empty_bar + NST);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_r = (n_rows + BM - 1) / BM, tiles_c = (n_cols + BNT - 1) / BNT;
    const int n_chunks = (kg_len + KGS - 1) / KGS;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONS_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();

    if constexpr (NT == 4 || MODE == 1) {
        if (warp >= CONS_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS_G4));
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS_G4));
    }
    if (warp >= CONS_WARPS) {
        // ===================== producer warp: one lane owns the ring and the tile order =====================
        if (warp != CONS_WARPS || lane != 0) return;
        // band state of the tile order: tiles [band_first, band_first + band_count) lie in column tiles [c0, c0 + w)
        int band = 0, c0 = 0, w = min(C::BAND, tiles_c);
        long long band_first = 0;
        auto band_tiles = [&](int bc0, int bw) -> long long {
            if constexpr (!SYM) return (long long)tiles_r * bw;
            long long n = 0;
            for (int r = 0; r < tiles_r && C::RATIO * r < bc0 + bw; ++r) n += bc0 + bw - max(bc0, C::RATIO * r);
            return n;
        };
        long long band_count = band_tiles(c0, w);
        int stage = 0;
        uint32_t parity = 1;                        // first pass over the ring: the empty barriers pass immediately
        for (int k = 0; k < my_tiles; ++k) {
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            while (t >= band_first + band_count) {
                band_first += band_count;
                ++band; c0 = band * C::BAND; w = min(C::BAND, tiles_c - c0);
                band_count = band_tiles(c0, w);
            }
            long long rem = t - band_first;
            int pr, pc;
            if constexpr (!SYM) { pr = (int)(rem / w); pc = c0 + (int)(rem - (long long)pr * w); }
            else {
                const int full_rows = min(tiles_r, c0 / C::RATIO + 1);          // rows whose whole band segment is kept
                if (rem < (long long)full_rows * w) { pr = (int)(rem / w); pc = c0 + (int)(rem - (long long)pr * w); }
                else {
                    rem -= (long long)full_rows * w;
                    pr = full_rows;
                    int cnt = c0 + w - C::RATIO * pr;
                    while (rem >= cnt) { rem -= cnt; ++pr; cnt = c0 + w - C::RATIO * pr; }
                    pc = C::RATIO * pr + (int)rem;
                }
            }
            const double* srcA = A + (size_t)(2 * pr) * kg_stride * (PANEL * 4);
            const double* srcB = B + (size_t)(C::B_PANELS * pc) * kg_stride * (PANEL * 4);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty_bar[stage], parity);
                if (c == 0) tile_desc[k % TILE_SLOTS] = make_int2(pr, pc);     // published by the full barrier of this chunk
                const int ng = min(KGS, kg_len - c * KGS);
                const uint32_t bytes = (uint32_t)ng * PANEL * 4 * 8;
                double* dst = sbuf + stage * STAGE_DBL;
                mbar_expect_tx(&full_bar[stage], (2 + C::B_PANELS) * bytes);
                const size_t koff = (size_t)c * KGS * (PANEL * 4);
                bulk_g2s(dst, srcA + koff, bytes, &full_bar[stage]);
                bulk_g2s(dst + PANEL_STAGE, srcA + (size_t)kg_stride * (PANEL * 4) + koff, bytes, &full_bar[stage]);
                bulk_g2s(dst + 2 * PANEL_STAGE, srcB + koff, bytes, &full_bar[stage]);
                if constexpr (NT == 4)
                    bulk_g2s(dst + 3 * PANEL_STAGE, srcB + (size_t)kg_stride * (PANEL * 4) + koff, bytes, &full_bar[stage]);
                if (++stage == NST) { stage = 0; parity ^= 1; }
            }
        }
        return;
    }

    // ===================== consumer warps: 4 (rows) x 4 (cols), warp tile 32 x 8 NT =====================
    const int wr = warp >> 2, wc = warp & 3;
    const int frag = (lane >> 2) * 4 + (lane & 3);                  // offset of this lane inside an 8 x 4 fragment
    const int a_off = (wr >> 1) * PANEL_STAGE + ((wr & 1) * 32) * 4 + frag;
    // B: warp columns [wc * 8 NT, +8 NT) of the tile; NT = 4: 32 columns, two warps per 64-row panel
    const int b_off = 2 * PANEL_STAGE + (NT == 4 ? (wc >> 1) * PANEL_STAGE + ((wc & 1) * 32) * 4 : (wc * 16) * 4) + frag;
    const bool vec_ok = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

    // Shared-window byte addresses: every fragment load is `ld.shared.f64 [stage base + immediate]`, issued through volatile
    // asm so that the order below - the order that keeps the fp64 pipe fed - is the order in the instruction stream.
    const uint32_t s_base = smem_u32(sbuf);
    const uint32_t a_byte = (uint32_t)a_off * 8u, b_byte = (uint32_t)b_off * 8u;
    constexpr uint32_t STAGE_BYTES = (uint32_t)STAGE_DBL * 8u, KG_BYTES = PANEL * 4 * 8, FRAG_BYTES = 32 * 8;
    auto load_frags = [&](uint32_t st_addr, int gi, double (&ra)[4], double (&rb)[NT]) {
#pragma unroll
        for (int m = 0; m < 4; ++m) ra[m] = lds_f64(st_addr + a_byte + (uint32_t)gi * KG_BYTES + (uint32_t)m * FRAG_BYTES);
#pragma unroll
        for (int n = 0; n < NT; ++n) rb[n] = lds_f64(st_addr + b_byte + (uint32_t)gi * KG_BYTES + (uint32_t)n * FRAG_BYTES);
    };

    int stage = 0;
    uint32_t phase = 0;
    for (int k = 0; k < my_tiles; ++k) {
        double acc[4][NT][2];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) { acc[m][n][0] = 0.0; acc[m][n][1] = 0.0; }
        // fragments of the next k-group, double-buffered in registers
        double fa[2][4], fb[2][NT];
        mbar_wait(&full_bar[stage], phase);
        const int2 desc = tile_desc[k % TILE_SLOTS];
        load_frags(s_base + (uint32_t)stage * STAGE_BYTES, 0, fa[0], fb[0]);
        double prev[MODE == 1 ? 4 : 1][MODE == 1 ? NT : 1][2];
        if constexpr (MODE == 1) {
            const int r_base = desc.x * BM + wr * 32 + (lane >> 2), c_base = desc.y * BNT + wc * (8 * NT) + 2 * (lane & 3);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int row = r_base + m * 8, col = c_base + n * 8;
                    prev[m][n][0] = 0.0; prev[m][n][1] = 0.0;
                    if (row < n_rows && col < n_cols) {
                        const double* src = out + (size_t)row * ld_out + col;
                        if (vec_ok && col + 1 < n_cols) {
                            const double2 v = *reinterpret_cast<const double2*>(src);
                            prev[m][n][0] = v.x; prev[m][n][1] = v.y;
                        } else {
                            prev[m][n][0] = src[0];
                            if (col + 1 < n_cols) prev[m][n][1] = src[1];
                        }
                    }
                }
        }
        int prev_stage = -1;                        // stage whose release is still owed (one k-group late, see below)
        for (int c = 0; c < n_chunks; ++c) {
            const int ng = min(KGS, kg_len - c * KGS);
            const uint32_t st_addr = s_base + (uint32_t)stage * STAGE_BYTES;
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == NST) { nstage = 0; nphase ^= 1; }
            const bool more = c + 1 < n_chunks;                         // another chunk of this tile follows
            if (ng == KGS) {
                // A warp owns the pipe one DMMA in four (16 of 64 cycles), so up to ~48 cycles of other instructions fit between
                // two of its DMMAs for free.  Everything that is not a DMMA is therefore placed right AFTER the first DMMA of a
                // k-group and spread over the k-groups: fragment loads of the next k-group (gi < 7), the wait on the next
                // stage's full barrier + its first fragments (gi = 7), the release of the PREVIOUS stage (gi = 0).
#pragma unroll
                for (int gi = 0; gi < KGS; ++gi) {
                    const int cur = gi & 1, nxt = cur ^ 1;
                    dmma884(acc[0][0], fa[cur][0], fb[cur][0]);
                    if (gi == 0 && prev_stage >= 0 && lane == 0) mbar_arrive(&empty_bar[prev_stage]);
                    if (gi + 1 < KGS) load_frags(st_addr, gi + 1, fa[nxt], fb[nxt]);
                    else if (more) {
                        mbar_wait(&full_bar[nstage], nphase);
                        load_frags(s_base + (uint32_t)nstage * STAGE_BYTES, 0, fa[nxt], fb[nxt]);
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int n = 0; n < NT; ++n)
                            if (m + n > 0) dmma884(acc[m][n], fa[cur][m], fb[cur][n]);
                }
            } else {
                // K tail (last chunk of a tile only; KGS is even, so the fragments of its first k-group sit in buffer 0)
                if (prev_stage >= 0 && lane == 0) mbar_arrive(&empty_bar[prev_stage]);
#pragma unroll 1
                for (int gi = 0; gi < ng; ++gi) {
                    if (gi > 0) load_frags(st_addr + (uint32_t)gi * KG_BYTES, 0, fa[0], fb[0]);
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int n = 0; n < NT; ++n) dmma884(acc[m][n], fa[0][m], fb[0][n]);
                }
            }
            // the loads of this stage were consumed by DMMAs that have issued (mma.sync is warp-collective: every lane's operands
            // were ready); the release is posted after the first DMMA of the next stage, or below at the end of the tile
            prev_stage = stage;
            stage = nstage; phase = nphase;
        }
        if (lane == 0) mbar_arrive(&empty_bar[prev_stage]);

        // ---- store: out[i, j] = scale * acc ----
        const int pr = desc.x, pc = desc.y;
        const int row0 = pr * BM + wr * 32 + (lane >> 2);
        const int col0 = pc * BNT + wc * (8 * NT) + 2 * (lane & 3);
        const bool interior = vec_ok && pr * BM + BM <= n_rows && pc * BNT + BNT <= n_cols;
        auto value = [&](int m, int n, int i) -> double {
            if constexpr (MODE == 1) return fma(scale, acc[m][n][i], prev[m][n][i]);
            else return scale * acc[m][n][i];
        };
        if (interior) {
            double* base = out + (size_t)row0 * ld_out + col0;
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    double* dst = base + (size_t)(m * 8) * ld_out + n * 8;
                    if constexpr (MODE == 1) *reinterpret_cast<double2*>(dst) = make_double2(value(m, n, 0), value(m, n, 1));   // read again soon: keep in L2
                    else st_cs_f64x2(dst, value(m, n, 0), value(m, n, 1));
                }
        } else {
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int row = row0 + m * 8, col = col0 + n * 8;
                    if (row < n_rows && col < n_cols) {
                        double* dst = out + (size_t)row * ld_out + col;
                        if (vec_ok && col + 1 < n_cols) st_cs_f64x2(dst, value(m, n, 0), value(m, n, 1));
                        else {
                            st_cs_f64(dst, value(m, n, 0));
                            if (col + 1 < n_cols) st_cs_f64(dst + 1, value(m, n, 1));
                        }
                    }
                }
        }
        if constexpr (SYM && MODE == 0) {
            // mirror: every entry strictly above the diagonal also goes to its transposed position (8 lanes of a quad column
            // write 64 contiguous bytes)
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int row = row0 + m * 8, col = col0 + n * 8;
                    if (row < n_rows) {
                        if (col > row && col < n_cols) st_cs_f64(out + (size_t)col * ld_out + row, scale * acc[m][n][0]);
                        if (col + 1 > row && col + 1 < n_cols) st_cs_f64(out + (size_t)(col + 1) * ld_out + row, scale * acc[m][n][1]);
                    }
                }
        }
    }
}

template <int NT, bool SYM, int MODE>
static int launch(const double* A, const double* B, int n_rows, int n_cols, int kg_stride, int kg_len, double* out, long long ld_out,
                  double scale, int sms, cudaStream_t stream) {
    using C = Cfg<NT>;
    auto kern = gram_kernel<NT, SYM, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int tiles_r = (n_rows + BM - 1) / BM, tiles_c = (n_cols + C::BNT - 1) / C::BNT;
    const long long tiles = count_tiles<NT, SYM>(tiles_r, tiles_c);
    const int grid = (int)(tiles < sms ? tiles : sms);
    if ((tiles + grid - 1) / grid > 0x7fffffffLL) return LSK_ESIZE;
    return (int)launch_pdl(kern, dim3(grid), dim3((NT == 4 || MODE == 1) ? THREADS_G4 : THREADS_G), (size_t)C::SMEM_BYTES, stream, A, B,
                           n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, tiles);
}

}  // namespace gram

// lsk_pairwise.cu dispatches EPI_LINEAR with a row-major output here.  `variant`: 0 = automatic, 2 / 4 = force the 128 x 64 /
// 128 x 128 tile (development: LSK_GRAM_TILE in the environment of the calling process, read once).
int gram_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_stride, int kg_len, double* out, long long ld_out,
                double scale, bool symmetric, int sms, cudaStream_t stream) {
    static const int forced = [] { const char* s = getenv("LSK_GRAM_TILE"); return s ? atoi(s) : 0; }();
    // 128 x 128 tiles (two thirds of the fragment loads and 25 % less L2 -> SM traffic per flop: 36.0-36.4 against 35.8-36.2 TFLOP/s)
    // once the problem fills at least four waves of them; smaller problems keep the finer 128 x 64 grain
    const long long t128 = (long long)((n_rows + 127) / 128) * ((n_cols + 127) / 128);
    int nt = (t128 >= 4ll * sms) ? 4 : 2;
    if (forced == 2 || forced == 4) nt = forced;
    if (symmetric) {
        if (nt == 4) return gram::launch<4, true, 0>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
        return gram::launch<2, true, 0>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
    }
    if (nt == 4) return gram::launch<4, false, 0>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
    return gram::launch<2, false, 0>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
}

// out += scale * A B^T in place (EPI_SYRK of lsk_pairwise_poly: the strip and trailing updates of lsk_potrf_upper / lsk_potri_upper);
// `symmetric`: only the tiles that reach the diagonal or lie above it.  `sms` may be smaller than the device's SM count: the
// look-ahead of the factorisation keeps a few SMs free for the panel chain that runs beside the trailing update.
int gram_update_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_stride, int kg_len, double* out, long long ld_out,
                       double scale, bool symmetric, int sms, cudaStream_t stream) {
    if (symmetric) return gram::launch<2, true, 1>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
    return gram::launch<2, false, 1>(A, B, n_rows, n_cols, kg_stride, kg_len, out, ld_out, scale, sms, stream);
}

}  // namespace lsk
