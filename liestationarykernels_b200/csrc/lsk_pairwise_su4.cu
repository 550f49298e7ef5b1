// Dedicated covariance kernel for SU(4) (BASELINE config C3: EigenbasisSumKernel on SU(4), heat / Matern, first 10 / 20 irreps).
//
// Same mathematics and operand layout as the generic kernel of lsk_pairwise.cu, which it replaces for these shapes (reference chain
// space.py:62-85, su.py:91-92,175-179, spectral_kernel.py:45-54): the standard SU(4) embedding has 21 k-groups in four bilinear forms
//   f0 = sum a c, f1 = sum b d, f2 = sum (a+b)(c-d)      (three-form complex trace, x = a + i b, y = c + i d: 4 k-groups each)
//   f3 = <R(x), R(y)>                                     (e_2 through the real SO(6) image of Lambda^2 C^4: 9 k-groups)
// and the class polynomial is a generated straight-line nested-Horner tape in (Re e_1, Im e_1, e_2) = (f0 + f1, f2 - f0 + f1, f3).
//
// The generic kernel spends 1155 warp-instructions per warp and 64 x 64 tile on this problem, of which 84 are DMMAs and ~170 DFMAs
// (ncu, profiles/su4_ncu_summary_r01_s3.json: fp64 datapath 71 % busy, issue slots 46 %): five pipeline chunks per tile, each with
// a claim-the-producer-duty attempt, a vote and an mbarrier round trip, a run-time variable map and predicated, layout-generic
// stores.  Here everything that does not depend on the data is fixed at compile time, as in so5::stair_kernel:
//   * ONE pipeline stage = the operands of a whole tile (two bulk copies of 42 KB, one mbarrier round trip per tile, 2 stages),
//     fed by a dedicated producer warpgroup;
//   * the DMMA phase is straight-line code over the 21 k-groups with the form boundaries as immediates;
//   * the variable map is two additions and a subtraction; the tape is the generated code selected by a template constant;
//   * incremental tile cursor, unpredicated 16-byte streaming stores on interior tiles.
#include <cstdlib>
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"
#include "lsk_generated_epilogues.cuh"

namespace lsk {
namespace su4 {

constexpr int KG = 21;                          // k-groups per embedded point
constexpr int PANEL_DBL = KG * PANEL * 4;       // one 64-row panel, all k-groups: 5376 doubles = 43 008 B, contiguous in HBM
constexpr int STAGE_DBL = 2 * PANEL_DBL;        // one A panel + one B panel: the operands of a 64 x 64 tile
constexpr int NSTAGE = 2;                       // 2 x 86 KB = 172 KB: one tile of look-ahead
constexpr int SMEM_BYTES = NSTAGE * STAGE_DBL * 8 + 2 * NSTAGE * 8 + 64;
constexpr int TM = 64, TN = 64;                 // CTA tile
// 16 consumer warps + a producer warpgroup with one active lane; registers move from the producer group to the consumers
// (`setmaxnreg`: 4 x 32 x 24 + 16 x 32 x 112 <= 640 x 96), as in so5::stair_kernel
constexpr int THREADS_SU4 = (CONSUMER_WARPS + 4) * 32;
constexpr int CONS_REGS = 112, PROD_REGS = 24;

__host__ __device__ constexpr int form_of(int gi) { return gi < 4 ? 0 : gi < 8 ? 1 : gi < 12 ? 2 : 3; }

template <int GEN>
__global__ void __launch_bounds__(THREADS_SU4, 1)
tile_kernel(const double* __restrict__ A, const double* __restrict__ B, int n_rows, int n_cols,
            double* __restrict__ out, long long ld_out, const __grid_constant__ LskEpilogue p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sbuf = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + NSTAGE * STAGE_DBL);
    uint64_t* empty_bar = full_bar + NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_tiles = (n_rows + TM - 1) / TM, col_tiles = (n_cols + TN - 1) / TN;
    const long long n_tiles = (long long)row_tiles * col_tiles;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();         // launched with programmatic serialisation: everything above overlapped the tail of the embedding kernel

    // tile cursor: t = blockIdx.x + k gridDim.x -> (row tile, column tile), advanced without a division
    struct Cursor { long long t; int r, c; };
    const int step_r = (int)(gridDim.x / (unsigned)col_tiles), step_c = (int)(gridDim.x % (unsigned)col_tiles);
    auto advance = [&](Cursor& cur) {
        cur.t += gridDim.x;
        cur.r += step_r; cur.c += step_c;
        if (cur.c >= col_tiles) { cur.c -= col_tiles; ++cur.r; }
    };
    Cursor cons{(long long)blockIdx.x, (int)(blockIdx.x / (unsigned)col_tiles), (int)(blockIdx.x % (unsigned)col_tiles)};

    // ===================== producer: a warpgroup of its own, one active lane =====================
    if (warp >= CONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS));
        if (warp != CONSUMER_WARPS || lane != 0) return;
        Cursor prod = cons;
        int p_stage = 0;
        uint32_t p_parity = 1;                       // first pass over the ring: the empty barriers pass immediately
        for (; prod.t < n_tiles; advance(prod)) {
            mbar_wait(&empty_bar[p_stage], p_parity);
            double* dst = sbuf + p_stage * STAGE_DBL;
            mbar_expect_tx(&full_bar[p_stage], STAGE_DBL * 8);
            bulk_g2s(dst, A + (size_t)prod.r * PANEL_DBL, PANEL_DBL * 8, &full_bar[p_stage]);
            bulk_g2s(dst + PANEL_DBL, B + (size_t)prod.c * PANEL_DBL, PANEL_DBL * 8, &full_bar[p_stage]);
            if (++p_stage == NSTAGE) { p_stage = 0; p_parity ^= 1; }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS));

    // ===================== consumer warps: 4 (rows) x 4 (cols), warp tile 16 x 16 =====================
    const int wr = warp >> 2, wc = warp & 3;
    const int frag = (lane >> 2) * 4 + (lane & 3);           // offset of this lane inside an 8x4 fragment
    const int a_off = (wr * 16) * 4 + frag;
    const int b_off = PANEL_DBL + (wc * 16) * 4 + frag;
    const bool vec_ok = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const long long ld8 = 8 * ld_out;

    int stage = 0;
    uint32_t phase = 0;
    for (; cons.t < n_tiles; advance(cons)) {
        double acc[4][2][2][2];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) { acc[f][m][n][0] = 0.0; acc[f][m][n][1] = 0.0; }

        mbar_wait(&full_bar[stage], phase);
        // ---- the four bilinear forms on the tensor-core path, straight-line over the 21 k-groups ----
        {
            const double* pa = sbuf + stage * STAGE_DBL + a_off;
            const double* pb = sbuf + stage * STAGE_DBL + b_off;
#pragma unroll
            for (int gi = 0; gi < KG; ++gi) {
                const int f = form_of(gi);
                double a[2], b[2];
#pragma unroll
                for (int m = 0; m < 2; ++m) a[m] = pa[gi * (PANEL * 4) + m * 32];
#pragma unroll
                for (int n = 0; n < 2; ++n) b[n] = pb[gi * (PANEL * 4) + n * 32];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 2; ++n) dmma884(acc[f][m][n], a[m], b[n]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);      // hand the stage back to the producer
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }

        // ---- class polynomial on the 8 accumulator entries of this lane ----
        constexpr int R = 8;
        double var[3][R], res[R];
#pragma unroll
        for (int e = 0; e < R; ++e) {
            const int m = e >> 2, n = (e >> 1) & 1, i = e & 1;
            const double f0 = acc[0][m][n][i], f1 = acc[1][m][n][i], f2 = acc[2][m][n][i];
            var[0][e] = f0 + f1;                             // Re tr(x y^H)
            var[1][e] = (f1 - f0) + f2;                      // Im tr(x y^H)  (same association as the generic variable map)
            var[2][e] = acc[3][m][n][i];                     // e_2(x y^H) (real)
        }
        gen_epilogue<GEN, R, 3>(p, var, res);
        const int row0 = cons.r * TM + wr * 16 + (lane >> 2);
        const int col0 = cons.c * TN + wc * 16 + 2 * (lane & 3);
        const bool interior = vec_ok && (cons.r * TM + TM <= n_rows) && (cons.c * TN + TN <= n_cols);
        if (interior) {
            double* tile_out = out + (size_t)row0 * ld_out + col0;
#pragma unroll
            for (int e = 0; e < R; e += 2) {
                const int m = e >> 2, n = (e >> 1) & 1;
                st_cs_f64x2(tile_out + m * ld8 + n * 8, p.scale * res[e], p.scale * res[e + 1]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < R; e += 2) {
                const int row = row0 + (e >> 2) * 8, col = col0 + ((e >> 1) & 1) * 8;
                const double r0 = p.scale * res[e], r1 = p.scale * res[e + 1];
                if (row < n_rows && col < n_cols) {
                    double* dst = out + (size_t)row * ld_out + col;
                    if (vec_ok && col + 1 < n_cols) st_cs_f64x2(dst, r0, r1);
                    else {
                        st_cs_f64(dst, r0);
                        if (col + 1 < n_cols) st_cs_f64(dst + 1, r1);
                    }
                }
            }
        }
    }
}

template <int GEN>
static int launch_gen(const double* A, const double* B, int n_rows, int n_cols, double* out, long long ld_out,
                      const LskEpilogue& p, int sms, cudaStream_t stream) {
    const long long tiles = (long long)((n_rows + TM - 1) / TM) * ((n_cols + TN - 1) / TN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    auto kern = tile_kernel<GEN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    return (int)launch_pdl(kern, dim3(grid), dim3(THREADS_SU4), (size_t)SMEM_BYTES, stream, A, B, n_rows, n_cols, out, ld_out, p);
}

}  // namespace su4

// Called by lsk_pairwise_poly when a HORNER problem matched generated tape `gen_id`; LSK_EUNSUPPORTED unless it is the standard
// SU(4) problem (embedding, variable map, row-major output) of one of the compiled truncations.
int su4_tile_launch(int gen_id, const double* A, const double* B, int n_rows, int n_cols, int kg_total, double* out, long long ld_out,
                    const LskEpilogue& p, int sms, cudaStream_t stream) {
    using namespace su4;
    if (gen_id != 3 && gen_id != 4) return LSK_EUNSUPPORTED;
    if (kg_total != KG || p.nforms != 4 || p.nvars != 3 || p.out_layout != LSK_OUT_ROW_MAJOR) return LSK_EUNSUPPORTED;
    const int want_begin[4] = {0, 4, 8, 12}, want_end[4] = {4, 8, 12, 21};
    for (int f = 0; f < 4; ++f)
        if (p.group_begin[f] != want_begin[f] || p.group_end[f] != want_end[f]) return LSK_EUNSUPPORTED;
    if (((p.reserved >> 8) & 0xff) != 0 || ((p.reserved >> 1) & 1)) return LSK_EUNSUPPORTED;     // no squared forms, no symmetric mode
    const double want[3][5] = {{0, 1, 1, 0, 0}, {0, -1, 1, 1, 0}, {0, 0, 0, 0, 1}};
    for (int v = 0; v < 3; ++v)
        for (int i = 0; i < 5; ++i)
            if (p.aff[v][i] != want[v][i]) return LSK_EUNSUPPORTED;
    if (gen_id == 3) return launch_gen<3>(A, B, n_rows, n_cols, out, ld_out, p, sms, stream);
    return launch_gen<4>(A, B, n_rows, n_cols, out, ld_out, p, sms, stream);
}

}  // namespace lsk
