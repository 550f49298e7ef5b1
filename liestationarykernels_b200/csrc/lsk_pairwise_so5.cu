// Dedicated covariance kernel for SO(5) with the monomial-staircase class polynomial (the BASELINE headline:
// EigenbasisSumKernel on SO(5), Matern / heat, first 10 / 20 / 100 irreps).
//
// Same mathematics and operand layout as the generic kernel of lsk_pairwise.cu (which it replaces for these shapes;
// reference chain space.py:62-85, so.py:136-168,288-293, spectral_kernel.py:45-54), with everything the generic kernel
// decides at run time fixed at compile time:
//   * operands are the standard SO(5) embedding [matrix (25 + 3 zeros) | Spin(5) rotor (16)] = 11 k-groups, so ONE
//     pipeline stage holds the operands of a whole 128 x 64 tile (two bulk copies, 45 + 22.5 KB, one mbarrier round trip
//     per tile) and the DMMA phase is straight-line code;
//   * the variable map is s1 = (tr g - 1) / 2,  s2 = 4 <R_x, R_y>^2 - (tr g + 1) / 2  with immediate constants;
//   * the staircase shape is a template constant: nested Horner is straight-line code, one DFMA per coefficient, the
//     coefficients are fetched two at a time by uniform loads at immediate offsets of the kernel-parameter bank.
// What this buys is measured in profiles/README.md: with four warps per SM sub-partition running the same phases in
// lockstep, step time = 4 x (fp64-pipe work per warp) + (everything that is not fp64-pipe work); this kernel attacks
// the second term (tile index arithmetic, variable map, store addressing, barrier traffic).
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"

#ifdef LSK_TRACE
// development build only (tools/trace_build.py): per-warp clock64 stamps of CTA 0 for the first tiles
#define LSK_TRACE_TILES 12
#define LSK_TRACE_EVENTS 8
__device__ long long g_lsk_trace[16][LSK_TRACE_TILES][LSK_TRACE_EVENTS];
#define LSK_STAMP(ev) do { if (blockIdx.x == 0 && lane == 0 && tile_no < LSK_TRACE_TILES) g_lsk_trace[warp][tile_no][ev] = clock64(); } while (0)
extern "C" int lsk_debug_trace(void* host) { return (int)cudaMemcpyFromSymbol(host, g_lsk_trace, sizeof(g_lsk_trace)); }
#else
#define LSK_STAMP(ev) do { } while (0)
#endif

namespace lsk {
namespace so5 {

constexpr int KG = 11;                          // k-groups per embedded point: [0, 7) matrix, [7, 11) rotor
constexpr int KG_MATRIX = 7;
constexpr int PANEL_DBL = KG * PANEL * 4;       // one 64-row panel, all k-groups: 2816 doubles = 22 528 B, contiguous in HBM
constexpr int STAGE_DBL = 3 * PANEL_DBL;        // two A panels + one B panel: the operands of one tile
constexpr int NSTAGE = 3;                       // 3 x 67.6 KB = 203 KB: two tiles of look-ahead
constexpr int SMEM_BYTES = NSTAGE * STAGE_DBL * 8 + 2 * NSTAGE * 8 + 64;
// 16 consumer warps + a producer WARPGROUP (one active lane).  Twenty warps start with 96 registers each; the producer group gives
// most of its share back and the consumers grow to 112 (`setmaxnreg`): 4 x 32 x 24 + 16 x 32 x 112 <= 640 x 96.
constexpr int THREADS_SO5 = (CONSUMER_WARPS + 4) * 32;
constexpr int CONS_REGS = 112, PROD_REGS = 24;

// Nested Horner over a compile-time staircase: bits [4k, 4k+4) of SHAPE = coefficient pairs of row slot k (evaluation
// order, highest power of s2 first; slot k starts at coef[16 k], highest power of s1 first), bit 48 + k = the slot's first
// coefficient is padding (odd row length).  Each row is the body of a loop that runs exactly once (`once` is a run-time
// 1): the back edge keeps ptxas from interleaving the independent rows, which costs registers (spills at 128) and buys
// nothing - 8 entries x 2 issue cycles already cover the 8-cycle DFMA latency.
template <int R, unsigned long long SHAPE>
__device__ __forceinline__ void stair_fixed(const LskEpilogue& p, int once, const double (&s1)[R], const double (&s2)[R], double (&out)[R]) {
    double outer[R];
#pragma unroll
    for (int k = 0; k < LSK_STAIR_MAX_ROWS; ++k) {
        const int pairs = (int)((SHAPE >> (4 * k)) & 15ull);
        const int odd = (int)((SHAPE >> (48 + k)) & 1ull);
        const int len = 2 * pairs - odd, base = LSK_STAIR_STRIDE * k + odd;
        if (pairs > 0) {
            int it = 0;
#pragma unroll 1
            do {
                double inner[R];
                if (len == 1) {
#pragma unroll
                    for (int r = 0; r < R; ++r) inner[r] = p.coef[base];
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) inner[r] = fma(s1[r], p.coef[base], p.coef[base + 1]);
#pragma unroll
                    for (int q = 2; q < len; ++q) {
#pragma unroll
                        for (int r = 0; r < R; ++r) inner[r] = fma(inner[r], s1[r], p.coef[base + q]);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) outer[r] = (k == 0) ? inner[r] : fma(outer[r], s2[r], inner[r]);
            } while (++it < once);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = outer[r];
}

template <unsigned long long SHAPE, bool SYM>
__global__ void __launch_bounds__(THREADS_SO5, 1)
stair_kernel(const double* __restrict__ A, const double* __restrict__ B, int n_rows, int n_cols,
             double* __restrict__ out, long long ld_out, const __grid_constant__ LskEpilogue p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sbuf = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + NSTAGE * STAGE_DBL);
    uint64_t* empty_bar = full_bar + NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_tiles = (n_rows + BM - 1) / BM, col_panels = (n_cols + BN - 1) / BN;
    // symmetric mode (A and B describe the same points): only tiles that reach the diagonal or lie above it
    const long long n_tiles = SYM ? (long long)row_tiles * col_panels - (long long)row_tiles * (row_tiles - 1)
                                  : (long long)row_tiles * col_panels;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();         // launched with programmatic serialisation: everything above overlapped the tail of the embedding kernel

    // Tile cursor: tile index t = blockIdx.x + k gridDim.x  ->  (row tile, column panel).  General problems advance it
    // incrementally (no division in the tile loop); the triangular schedule of the symmetric mode inverts a quadratic.
    struct Cursor { long long t; int r, c; };
    const int step_r = SYM ? 0 : (int)(gridDim.x / (unsigned)col_panels), step_c = SYM ? 0 : (int)(gridDim.x % (unsigned)col_panels);
    auto locate = [&](Cursor& cur) {
        if constexpr (!SYM) { cur.r = (int)(cur.t / col_panels); cur.c = (int)(cur.t % col_panels); return; }
        // row tile r owns col_panels - 2 r tiles; S(r) = r * col_panels - r (r - 1) tiles precede it
        const double c1 = (double)col_panels + 1.0;
        int r = (int)(0.5 * (c1 - sqrt(fmax(c1 * c1 - 4.0 * (double)cur.t, 0.0))));
        r = max(0, min(r, row_tiles - 1));
        if ((long long)r * col_panels - (long long)r * (r - 1) > cur.t) --r;
        if (r + 1 < row_tiles && (long long)(r + 1) * col_panels - (long long)(r + 1) * r <= cur.t) ++r;
        cur.r = r;
        cur.c = 2 * r + (int)(cur.t - ((long long)r * col_panels - (long long)r * (r - 1)));
    };
    auto advance = [&](Cursor& cur) {
        cur.t += gridDim.x;
        if constexpr (SYM) { if (cur.t < n_tiles) locate(cur); }
        else {
            cur.r += step_r; cur.c += step_c;
            if (cur.c >= col_panels) { cur.c -= col_panels; ++cur.r; }
        }
    };
    Cursor cons{(long long)blockIdx.x, 0, 0};
    if (cons.t < n_tiles) locate(cons);

    // ===================== producer: a warpgroup of its own, one active lane =====================
    // (Round 1 folded the producer duty into thread 0 of consumer warp 0: ~50 single-thread instructions per tile during which
    // that warp issues nothing.  Its three sub-partition mates take the pipe meanwhile, but the warp itself falls behind, the
    // ring is released at ITS pace, and every other warp spins on the next full barrier: 55 try_wait iterations per warp and tile
    // in the source-level profile, profiles/so5_ncu_summary_r02.json.)
    if (warp >= CONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS));
        if (warp != CONSUMER_WARPS || lane != 0) return;
        Cursor prod = cons;
        int p_stage = 0;
        uint32_t p_parity = 1;                       // first pass over the ring: the empty barriers pass immediately
        for (; prod.t < n_tiles; advance(prod)) {
            mbar_wait(&empty_bar[p_stage], p_parity);
            const double* srcA = A + (size_t)(2 * prod.r) * PANEL_DBL;       // operands are padded to ROW_PAD rows: both panels exist
            const double* srcB = B + (size_t)prod.c * PANEL_DBL;
            double* dst = sbuf + p_stage * STAGE_DBL;
            mbar_expect_tx(&full_bar[p_stage], STAGE_DBL * 8);
            bulk_g2s(dst, srcA, 2 * PANEL_DBL * 8, &full_bar[p_stage]);      // the two A panels are adjacent in HBM
            bulk_g2s(dst + 2 * PANEL_DBL, srcB, PANEL_DBL * 8, &full_bar[p_stage]);
#ifdef LSK_TRACE
            if (blockIdx.x == 0) { const long long pt = (prod.t - blockIdx.x) / gridDim.x; if (pt < LSK_TRACE_TILES) g_lsk_trace[0][pt][7] = clock64(); }
#endif
            if (++p_stage == NSTAGE) { p_stage = 0; p_parity ^= 1; }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS));

    // ===================== consumer warps: 4 (rows) x 4 (cols), warp tile 32 x 16 =====================
    const int wr = warp >> 2, wc = warp & 3;
    const int frag = (lane >> 2) * 4 + (lane & 3);           // offset of this lane inside an 8x4 fragment
    const int a_off = (wr >> 1) * PANEL_DBL + ((wr & 1) * 32) * 4 + frag;
    const int b_off = 2 * PANEL_DBL + (wc * 16) * 4 + frag;
    const bool vec_ok = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int once = p.nvars - 1;                            // = 1 (see stair_fixed)
    const long long ld8 = 8 * ld_out;

    int stage = 0;
    uint32_t phase = 0;
#ifdef LSK_TRACE
    int tile_no = -1;
#endif
    for (; cons.t < n_tiles; advance(cons)) {
#ifdef LSK_TRACE
        ++tile_no;
#endif
        LSK_STAMP(0);
        double acc[2][4][2][2];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) { acc[f][m][n][0] = 0.0; acc[f][m][n][1] = 0.0; }

        mbar_wait(&full_bar[stage], phase);
        LSK_STAMP(1);
        // ---- both bilinear forms on the tensor-core path, straight-line: tr g over k-groups [0, 7), <R_x, R_y> over [7, 11) ----
        {
            const double* pa = sbuf + stage * STAGE_DBL + a_off;
            const double* pb = sbuf + stage * STAGE_DBL + b_off;
#pragma unroll
            for (int gi = 0; gi < KG; ++gi) {
                const int f = gi < KG_MATRIX ? 0 : 1;
                double a[4], b[2];
#pragma unroll
                for (int m = 0; m < 4; ++m) a[m] = pa[gi * (PANEL * 4) + m * 32];
#pragma unroll
                for (int n = 0; n < 2; ++n) b[n] = pb[gi * (PANEL * 4) + n * 32];
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 2; ++n) dmma884(acc[f][m][n], a[m], b[n]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);      // hand the stage back to the producer
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        LSK_STAMP(2);

        // ---- class polynomial on the accumulator fragments: ONE copy of the straight-line code, run for both halves of
        //      the warp tile (the second half of the accumulators is shifted into the registers of the first) ----
        const int row0 = cons.r * BM + wr * 32 + (lane >> 2);
        const int col0 = cons.c * BN + wc * 16 + 2 * (lane & 3);
        const bool interior = vec_ok && !SYM && (cons.r * BM + BM <= n_rows) && (cons.c * BN + BN <= n_cols);
        double* tile_out = out + (size_t)row0 * ld_out + col0;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            constexpr int R = 8;
            double s1[R], s2[R], res[R];
#pragma unroll
            for (int e = 0; e < R; ++e) {
                const int ml = e >> 2, n = (e >> 1) & 1, i = e & 1;
                const double u0 = acc[0][ml][n][i], u1 = acc[1][ml][n][i];
                s1[e] = fma(0.5, u0, -0.5);                          // s1 = e_1(cos theta) = (tr g - 1) / 2
                s2[e] = fma(4.0 * u1, u1, fma(-0.5, u0, -0.5));      // s2 = (4 <R_x,R_y>)^2 / 4 - 1 - s1
            }
            LSK_STAMP(3 + 2 * h);
            stair_fixed<R, SHAPE>(p, once, s1, s2, res);
            LSK_STAMP(4 + 2 * h);
            if (interior) {
#pragma unroll
                for (int e = 0; e < R; e += 2) {
                    const int ml = e >> 2, n = (e >> 1) & 1;
                    st_cs_f64x2(tile_out + (2 * h + ml) * ld8 + n * 8, res[e], res[e + 1]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < R; e += 2) {
                    const int row = row0 + (2 * h + (e >> 2)) * 8, col = col0 + ((e >> 1) & 1) * 8;
                    const double r0 = res[e], r1 = res[e + 1];
                    if (row < n_rows && col < n_cols) {
                        double* dst = out + (size_t)row * ld_out + col;
                        if constexpr (SYM) {      // mirror: entries strictly above the diagonal
                            if (col > row) st_cs_f64(out + (size_t)col * ld_out + row, r0);
                            if (col + 1 > row && col + 1 < n_cols) st_cs_f64(out + (size_t)(col + 1) * ld_out + row, r1);
                        }
                        if (vec_ok && col + 1 < n_cols) st_cs_f64x2(dst, r0, r1);
                        else {
                            st_cs_f64(dst, r0);
                            if (col + 1 < n_cols) st_cs_f64(dst + 1, r1);
                        }
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 2; ++n) { acc[f][m][n][0] = acc[f][m + 2][n][0]; acc[f][m][n][1] = acc[f][m + 2][n][1]; }
        }
    }
}

template <unsigned long long SHAPE>
static int launch_shape(const double* A, const double* B, int n_rows, int n_cols, double* out, long long ld_out,
                        const LskEpilogue& p, int sms, bool sym, cudaStream_t stream) {
    const long long rt = (n_rows + BM - 1) / BM, cp = (n_cols + BN - 1) / BN;
    const long long tiles = sym ? rt * cp - rt * (rt - 1) : rt * cp;
    const int grid = (int)(tiles < sms ? tiles : sms);
    cudaError_t e;
    if (sym) {
        auto kern = stair_kernel<SHAPE, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        return (int)launch_pdl(kern, dim3(grid), dim3(THREADS_SO5), (size_t)SMEM_BYTES, stream, A, B, n_rows, n_cols, out, ld_out, p);
    }
    auto kern = stair_kernel<SHAPE, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    return (int)launch_pdl(kern, dim3(grid), dim3(THREADS_SO5), (size_t)SMEM_BYTES, stream, A, B, n_rows, n_cols, out, ld_out, p);
}

}  // namespace so5

int so5_stair_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_total, double* out, long long ld_out,
                     const LskEpilogue& p, int sms, cudaStream_t stream) {
    using namespace so5;
    // the dedicated kernel hard-codes the standard SO(5) embedding, its variable map and row-major output
    if (kg_total != KG || p.nforms != 2 || p.nvars != 2 || p.out_layout != LSK_OUT_ROW_MAJOR) return LSK_EUNSUPPORTED;
    if (p.group_begin[0] != 0 || p.group_end[0] != KG_MATRIX || p.group_begin[1] != KG_MATRIX || p.group_end[1] != KG) return LSK_EUNSUPPORTED;
    if (((p.reserved >> 8) & 0xff) != 0x2) return LSK_EUNSUPPORTED;                 // only the rotor form enters squared
    if (p.scale != 1.0) return LSK_EUNSUPPORTED;                                     // the host folds the scale into the table
    const double want[2][3] = {{-0.5, 0.5, 0.0}, {-0.5, -0.5, 4.0}};
    for (int v = 0; v < 2; ++v)
        for (int i = 0; i < 3; ++i)
            if (p.aff[v][i] != want[v][i]) return LSK_EUNSUPPORTED;
    const bool sym = ((p.reserved >> 1) & 1) != 0;
#define LSK_SO5_SHAPE(S) if (p.stair_shape == (S)) return launch_shape<(S)>(A, B, n_rows, n_cols, out, ld_out, p, sms, sym, stream);
    LSK_SO5_SHAPE(0x54e087766544321ull)      // first 100 irreps
    LSK_SO5_SHAPE(0x1b000000043221ull)       // first 20 irreps (default order)
    LSK_SO5_SHAPE(0x6000000000321ull)        // first 10 irreps
#undef LSK_SO5_SHAPE
    return LSK_EUNSUPPORTED;
}

}  // namespace lsk
