// Fused pairwise covariance kernel:  out[i,j] = scale * P( invariants of x_i y_j^{-1} ).
//
// Replaces, for all N*M pairs at once, the reference chain
//   cartesian_prod (utils.py:40-52) -> bmm (space.py:62-78) -> torus_representative (so.py:136-168, su.py:91-92)
//   -> per-monomial character loop (so.py:288-293, su.py:175-179) -> weighted sum (spectral_kernel.py:45-54).
//
// Formulation: the conjugation invariants of g = x_i y_j^{-1} that the characters depend on
// (tr g, e_2(g), ...) are *bilinear* in per-point embeddings (vec x, Lambda^2 x, ...: Cauchy-Binet), so the
// pair stage is a small-K fp64 GEMM  u_f = A_f B_f^T  run on the fp64 tensor-core path (DMMA m8n8k4) with operands
// staged through shared memory by 1-D bulk async copies (TMA engine) in a 4-stage mbarrier ring fed by a
// dedicated producer warp.  The class polynomial is then evaluated in registers on the accumulator fragments
// (coefficients come from the kernel-parameter constant bank through uniform loads) and the covariance tile is
// written with 16-byte streaming stores.  Nothing of size N*M exists except the output.
//
// Operand layout ("panel-major", produced by lsk_embed.cu): rows are grouped in panels of 64 (zero-padded to a multiple of 128 rows); inside a panel
// the data is [k-group g][row r][4 doubles], so (a) one pipeline stage of a panel is one contiguous 16 KB
// block (one bulk copy) and (b) a DMMA A/B fragment (8 rows x 4 k) is one contiguous, conflict-free 256 B read.
#include <cstdlib>
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"
#include "lsk_generated_epilogues.cuh"

#ifdef LSK_TRACE
// development build only (tools/trace_build.py): clock64 stamps of CTA 0 of the generic kernel
#define LSK_GT_TILES 10
#define LSK_GT_EVENTS 16
__device__ long long g_lsk_gtrace[16][LSK_GT_TILES][LSK_GT_EVENTS];
__device__ long long g_lsk_gissue[128];
#define LSK_GSTAMP(ev) do { if (blockIdx.x == 0 && lane == 0 && tile_no < LSK_GT_TILES && (ev) < LSK_GT_EVENTS) g_lsk_gtrace[warp][tile_no][ev] = clock64(); } while (0)
extern "C" int lsk_debug_gtrace(void* host, void* issue) {
    cudaError_t e = cudaMemcpyFromSymbol(host, g_lsk_gtrace, sizeof(g_lsk_gtrace));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyFromSymbol(issue, g_lsk_gissue, sizeof(g_lsk_gissue));
}
#else
#define LSK_GSTAMP(ev) do { } while (0)
#endif

namespace lsk {

// lsk_pairwise_so5.cu: LSK_EUNSUPPORTED when the problem is not one of the compiled SO(5) shapes
int so5_stair_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_total, double* out, long long ld_out,
                     const LskEpilogue& p, int sms, cudaStream_t stream);

// lsk_pairwise_su4.cu: LSK_EUNSUPPORTED unless the problem is the standard SU(4) one of generated tape `gen_id`
int su4_tile_launch(int gen_id, const double* A, const double* B, int n_rows, int n_cols, int kg_total, double* out, long long ld_out,
                    const LskEpilogue& p, int sms, cudaStream_t stream);

// lsk_gram.cu: out = scale * A B^T on the dedicated Gram kernel (row-major output)
int gram_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_stride, int kg_len, double* out, long long ld_out,
                double scale, bool symmetric, int sms, cudaStream_t stream);

int gram_update_launch(const double* A, const double* B, int n_rows, int n_cols, int kg_stride, int kg_len, double* out, long long ld_out,
                       double scale, bool symmetric, int sms, cudaStream_t stream);

constexpr int PANEL_STAGE = KG_PER_STAGE * PANEL * 4;            // one panel, one stage: 2048 doubles = 16 KB
constexpr int B_STAGE = PANEL_STAGE;
// per tile variant (MT = 8-row fragments per warp): A panels per stage, ring depth, dynamic shared memory.  The 64-row
// variant serves problems with many forms (five or more chunks per tile): its smaller stages buy a ring deep enough to
// prefetch a whole tile during the previous tile's epilogue.
template <int MT> struct TileCfg {
    static constexpr int A_STAGE = (MT / 2) * PANEL_STAGE;
    static constexpr int NST = MT == 4 ? STAGES : 7;
    static constexpr int SMEM_BYTES = NST * (A_STAGE + B_STAGE) * 8 + 2 * NST * 8 + 128;   // + producer state (16 ints)
};

// ------------------------------------------------------------------------------------------------------------
// epilogues: R entries per call so that every uniform coefficient load feeds R DFMAs
// ------------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void epi_cheb1(const double* coef, int nterms, const double (&v0)[R], double (&out)[R]) {
    // Clenshaw for sum_k c_k T_k(x):  b_k = c_k + 2x b_{k+1} - b_{k+2};  result = c_0 + x b_1 - b_2
    double b1[R], b2[R], x2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { b1[r] = 0.0; b2[r] = 0.0; x2[r] = 2.0 * v0[r]; }
    for (int k = nterms - 1; k >= 1; --k) {
        const double c = coef[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double t = fma(x2[r], b1[r], c - b2[r]);
            b2[r] = b1[r];
            b1[r] = t;
        }
    }
    const double c0 = coef[0];
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = fma(v0[r], b1[r], c0 - b2[r]);
}

template <int R>
__device__ __forceinline__ void epi_stair2(const LskEpilogue& p, const double (&s1)[R], const double (&s2)[R], double (&out)[R]) {
    // sum_b s2^b ( sum_a C[b][a] s1^a ) by nested Horner.  Row slot k (evaluation order) sits at the fixed offset 16 k
    // of coef[], holds 2 * pairs_k coefficients (highest power first, zero-padded to even length) and pairs_k comes
    // from 4 bits of the uniform shape word.  The row loop is unrolled, so the code is a sequence of simple
    // runtime-trip-count loops whose bounds and coefficient addresses are warp-uniform: ptxas keeps all of it on the
    // uniform datapath (LDCU + DFMA with a uniform-register operand, UISETP/BRA.U loop control).  Nested
    // data-dependent loops, a per-step branch or shared-memory coefficients each cost 10-25 % of the DFMA rate
    // (tools/horner_loop.cu, profiles/horner_loop_r01.json).
    double outer[R];
#pragma unroll
    for (int r = 0; r < R; ++r) outer[r] = 0.0;
    const unsigned long long shape = p.stair_shape;
#pragma unroll
    for (int k = 0; k < LSK_STAIR_MAX_ROWS; ++k) {
        const int pairs = (int)((shape >> (4 * k)) & 15ull);
        if (pairs > 0) {
            double inner[R];
            {
                const double c0 = p.coef[LSK_STAIR_STRIDE * k], c1 = p.coef[LSK_STAIR_STRIDE * k + 1];
#pragma unroll
                for (int r = 0; r < R; ++r) inner[r] = fma(s1[r], c0, c1);
            }
#pragma unroll 1          // code size: 16 row copies x 2 tile halves must stay inside the instruction cache
            for (int q = 1; q < pairs; ++q) {
                const double c0 = p.coef[LSK_STAIR_STRIDE * k + 2 * q], c1 = p.coef[LSK_STAIR_STRIDE * k + 2 * q + 1];
#pragma unroll
                for (int r = 0; r < R; ++r) inner[r] = fma(inner[r], s1[r], c0);
#pragma unroll
                for (int r = 0; r < R; ++r) inner[r] = fma(inner[r], s1[r], c1);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) outer[r] = fma(outer[r], s2[r], inner[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = outer[r];
}

template <int R, int D>
__device__ __forceinline__ void epi_orbit2(const LskEpilogue& p, const double (&s1)[R], const double (&s2)[R], double (&out)[R]) {
    // c1, c2 = roots of z^2 - s1 z + s2;  sum_{a,b<D} W[a][b] T_a(c1) T_b(c2), W dense row-major D x D
    double tb[R][D], ta0[R], ta1[R], c1x2[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double disc = fmax(fma(s1[r], s1[r], -4.0 * s2[r]), 0.0);
        const double rt = sqrt(disc);
        const double c1 = 0.5 * (s1[r] + rt), c2 = 0.5 * (s1[r] - rt);
        tb[r][0] = 1.0;
        tb[r][1] = c2;
#pragma unroll
        for (int b = 2; b < D; ++b) tb[r][b] = fma(2.0 * c2, tb[r][b - 1], -tb[r][b - 2]);
        ta0[r] = 1.0; ta1[r] = c1; c1x2[r] = 2.0 * c1; acc[r] = 0.0;
    }
    for (int a = 0; a < p.nterms; ++a) {
        double row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) row[r] = 0.0;
#pragma unroll
        for (int b = 0; b < D; ++b) {
            const double w = p.coef[a * D + b];
#pragma unroll
            for (int r = 0; r < R; ++r) row[r] = fma(w, tb[r][b], row[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r] = fma(ta0[r], row[r], acc[r]);
            const double nxt = fma(c1x2[r], ta1[r], -ta0[r]);   // T_{a+2}
            ta0[r] = ta1[r];
            ta1[r] = nxt;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = acc[r];
}

// Nested Horner over a run-time "tape" (any number of variables, any sparsity pattern): step t is (c, op) =
// (tape[2t], low word of tape[2t+1]);  a_l is the accumulator of nesting level l (variable v_l, level 0 innermost):
//   op 0: a_0 = a_0 v_0 + c        op 1: a_0 = c                      (coefficient chain of the innermost variable)
//   op 2 + 3(l-1): a_l = a_l v_l + a_{l-1}     +1: a_l = a_{l-1}     +2: a_l = a_l v_l   (close a level-(l-1) chain / gap)
// The result is a_{NV-1}.  One step = R fp64 operations for one uniform (c, op) fetch; all branches are warp-uniform.
// The host (classfn.horner_tape) emits one step per node of the exponent trie, so a polynomial with T terms costs about
// T + (number of distinct exponent prefixes) fused multiply-adds instead of T * degree multiplications.
template <int R, int NV>
__device__ __forceinline__ void epi_tape(const double* tape, int nsteps, const double (&v)[NV][R], double (&out)[R]) {
    double a[NV][R];
#pragma unroll
    for (int l = 0; l < NV; ++l)
#pragma unroll
        for (int r = 0; r < R; ++r) a[l][r] = 0.0;
    for (int t = 0; t < nsteps; ++t) {
        const double c = tape[2 * t];
        const int op = (int)__double_as_longlong(tape[2 * t + 1]);
        if (op == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) a[0][r] = fma(a[0][r], v[0][r], c);
        } else if (op == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) a[0][r] = c;
        } else {
#pragma unroll
            for (int l = 1; l < NV; ++l) {
                const int base = 2 + 3 * (l - 1);
                if (op == base) {
#pragma unroll
                    for (int r = 0; r < R; ++r) a[l][r] = fma(a[l][r], v[l][r], a[l - 1][r]);
                } else if (op == base + 1) {
#pragma unroll
                    for (int r = 0; r < R; ++r) a[l][r] = a[l - 1][r];
                } else if (op == base + 2) {
#pragma unroll
                    for (int r = 0; r < R; ++r) a[l][r] *= v[l][r];
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = a[NV - 1][r];
}

// k-groups of one form are cut into ceil(len / KG_PER_STAGE) chunks of nearly equal size (one pipeline stage each)
__device__ __forceinline__ int chunk_len(int remaining) {
    const int chunks = (remaining + KG_PER_STAGE - 1) / KG_PER_STAGE;
    return (remaining + chunks - 1) / chunks;
}

// ------------------------------------------------------------------------------------------------------------
// main kernel
//   EPI: epilogue kind;  NF: number of bilinear forms;  NV: number of variables (HORNER kinds);  D: ORBIT2 size
// ------------------------------------------------------------------------------------------------------------
//   MT:  8-row accumulator fragments per warp (4: 128 x 64 tile, two A panels;  2: 64 x 64 tile, one A panel - for
//        problems with three or four bilinear forms, whose accumulators would not fit in registers otherwise)
//   GEN: > 0 selects a generated straight-line epilogue (lsk_generated_epilogues.cuh) instead of the tape interpreter
template <int EPI, int NF, int NV, int D, bool SYM, int MT, int GEN>
__global__ void __launch_bounds__(THREADS, 1)
pairwise_kernel(const double* __restrict__ A, const double* __restrict__ B,
                int n_rows, int n_cols, int kg_total,               // kg_total: k-groups per row of A and B
                double* __restrict__ out, long long ld_out,
                const __grid_constant__ LskEpilogue p)
{
    constexpr int A_STAGE = TileCfg<MT>::A_STAGE, NST = TileCfg<MT>::NST;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + NST * A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + NST * B_STAGE);
    uint64_t* empty_bar = full_bar + NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BMT = 32 * MT;                    // rows per CTA tile
    static_assert(MT == 4 || (MT == 2 && !SYM), "the triangular schedule assumes 128 x 64 tiles");
    const int row_tiles = (n_rows + BMT - 1) / BMT, col_panels = (n_cols + BN - 1) / BN;
    // symmetric mode (bit 1 of `reserved`; A and B describe the same points): only tiles that reach the diagonal or lie
    // above it are computed, every entry above the diagonal is also stored at its mirror position
    constexpr bool sym = SYM;        // compile-time: the general kernel carries none of the triangular-schedule code
    const long long n_tiles = sym ? (long long)row_tiles * col_panels - (long long)row_tiles * (row_tiles - 1)
                                  : (long long)row_tiles * col_panels;
    // tile index -> (row tile, column panel); loop-free (it is also evaluated under the producer's divergent branch, once
    // per pipeline chunk) and integer-only.  Triangular schedule: row tile r owns the col_panels - 2 r tiles that reach the
    // diagonal or lie right of it; rows r and row_tiles - 1 - r together always own W = 2 col_panels - 2 (row_tiles - 1)
    // tiles, so a tile index splits into (pair, offset) with one division.
    const int pair_w = 2 * col_panels - 2 * (row_tiles - 1);
    const bool small_index = n_tiles <= 0x7fffffffLL;
    auto tile_coords = [&](long long t, int& pr, int& pc) {
        if constexpr (!SYM) {
            if (small_index) { const unsigned q = (unsigned)t / (unsigned)col_panels; pr = (int)q; pc = (int)((unsigned)t - q * (unsigned)col_panels); }
            else { pr = (int)(t / col_panels); pc = (int)(t % col_panels); }
            return;
        }
        int pair, off;
        if (small_index) { const unsigned q = (unsigned)t / (unsigned)pair_w; pair = (int)q; off = (int)((unsigned)t - q * (unsigned)pair_w); }
        else { pair = (int)(t / pair_w); off = (int)(t % pair_w); }
        const int first = col_panels - 2 * pair;
        if (off < first) { pr = pair; pc = 2 * pair + off; }
        else { pr = row_tiles - 1 - pair; pc = 2 * pr + (off - first); }
    };

    // producer state in shared memory: [0] next chunk (CTA-local sequence number) to be issued, [1] chunks per tile,
    // [2] chunks of this CTA, [4 + 3 f ..] per form: first chunk, floor chunk size, remainder (same cut as chunk_len():
    // the first `remainder` chunks of a form are one k-group longer)
    int* pstate = reinterpret_cast<int*>(empty_bar + NST);
    int* next_q = pstate;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        int cpt = 0;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int len = p.group_end[f] - p.group_begin[f], cnt = (len + KG_PER_STAGE - 1) / KG_PER_STAGE;
            pstate[4 + 3 * f] = cpt;
            pstate[5 + 3 * f] = len / cnt;
            pstate[6 + 3 * f] = len - (len / cnt) * cnt;
            cpt += cnt;
        }
        pstate[0] = 0;
        pstate[1] = cpt;
        pstate[2] = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * cpt;      // < 2^31: checked by the host
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();         // launched with programmatic serialisation: everything above overlapped the tail of the embedding kernel

    // ===================== producer duty: shared by all warps, whoever gets there first =====================
    // 16 consumer warps = 4 per SM sub-partition keeps the register budget at 128/thread (a 17th, dedicated producer warp
    // costs spills).  The chunks of this CTA form one sequence q = 0, 1, ... (tile-major, then form, then chunk); chunk q
    // lives in ring stage q % NST.  Lane 0 of ANY warp may issue the next chunk: it reads the shared counter, tests the
    // stage's empty barrier and claims q with a compare-and-swap.  With a fixed producer warp that warp falls behind the
    // others by its producer work, every other warp runs into the ring limit behind it and the whole CTA moves at the
    // pace of its slowest warp (per-warp clock traces, profiles/README.md); with claiming, the warps that are ahead do
    // the producer work and the laggard catches up.
    // One non-blocking attempt to issue the next chunk.  LOOP-FREE on purpose: a loop (even an inline-asm spin) under a
    // thread-divergent branch makes ptxas give up warp-uniformity for the whole enclosing tile loop, which moves the
    // epilogue's coefficient fetch from the uniform datapath (LDCU + uniform-register operands) to per-thread LDC and
    // costs ~15 % of its DFMA rate.  `wait_hw` = use the hardware-suspending try_wait instead of a plain test.
    auto try_issue = [&](bool wait_hw) {
        const int q = *reinterpret_cast<volatile int*>(next_q);
        if (q >= pstate[2]) return;
        const int st = q % NST;
        const uint32_t parity = (uint32_t)(((q / NST) & 1) ^ 1);  // previous use of the stage; first ring pass: immediately true
        const bool free_slot = wait_hw ? mbar_try_wait_once(&empty_bar[st], parity) : mbar_test(&empty_bar[st], parity);
        if (!free_slot) return;
        if (atomicCAS(next_q, q, q + 1) != q) return;             // another warp took it
        const int cpt = pstate[1], k = q / cpt, c = q - k * cpt;
        int gb = p.group_begin[0], base = pstate[5], rem = pstate[6], cum = 0;
#pragma unroll
        for (int f = 1; f < NF; ++f)
            if (c >= pstate[4 + 3 * f]) { gb = p.group_begin[f]; base = pstate[5 + 3 * f]; rem = pstate[6 + 3 * f]; cum = pstate[4 + 3 * f]; }
        const int i = c - cum;
        const int g0 = gb + i * base + min(i, rem), ng = base + (i < rem ? 1 : 0);
        int pr, pc;
        tile_coords(blockIdx.x + (long long)k * gridDim.x, pr, pc);
        const double* srcA0 = A + ((size_t)((MT / 2) * pr) * kg_total + g0) * (PANEL * 4);   // operands are padded to
        const double* srcA1 = srcA0 + (size_t)kg_total * (PANEL * 4);                         // ROW_PAD rows: both exist
        const double* srcB = B + ((size_t)pc * kg_total + g0) * (PANEL * 4);
        const uint32_t bytes = (uint32_t)ng * PANEL * 4 * 8;
#ifdef LSK_TRACE
        if (blockIdx.x == 0 && q < 128) g_lsk_gissue[q] = clock64();
#endif
        mbar_expect_tx(&full_bar[st], (MT / 2 + 1) * bytes);
        bulk_g2s(sA + st * A_STAGE, srcA0, bytes, &full_bar[st]);
        if constexpr (MT == 4) bulk_g2s(sA + st * A_STAGE + PANEL_STAGE, srcA1, bytes, &full_bar[st]);
        bulk_g2s(sB + st * B_STAGE, srcB, bytes, &full_bar[st]);
    };
    // keep the ring full without ever blocking (lane 0 of every warp; called wherever it is cheap)
    auto produce_ahead = [&]() {
        if (lane == 0) try_issue(false);
    };
    // the chunk a warp is about to wait for must be in flight: warp-uniform poll loop (the vote makes the trip count
    // uniform; normally the chunk was issued long ago and the loop body runs once)
    auto ensure_issued = [&](int qc) {
        unsigned done = 0;
        while (!done) {
            unsigned ok = 1;
            if (lane == 0 && *reinterpret_cast<volatile int*>(next_q) <= qc) {
                try_issue(true);
                ok = *reinterpret_cast<volatile int*>(next_q) > qc ? 1u : 0u;
            }
            done = __all_sync(0xffffffffu, ok);
        }
    };
#ifdef LSK_TRACE
    int tile_no = -1;
#endif

    // ===================== consumer warps: 4 (rows) x 4 (cols), warp tile 32 x 16 =====================
    const int wr = warp >> 2, wc = warp & 3;
    const int frag = (lane >> 2) * 4 + (lane & 3);           // offset of this lane inside an 8x4 fragment
    const int a_off = MT == 4 ? (wr >> 1) * PANEL_STAGE + ((wr & 1) * 32) * 4 + frag : (wr * 16) * 4 + frag;
    const int b_off = (wc * 16) * 4 + frag;
    const bool vec_ok = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

    int stage = 0, qc = 0;          // ring stage and sequence number of the chunk this warp consumes next
    uint32_t phase = 0;
    // the cut of every form into chunks does not depend on the tile: count, floor size and remainder come from per-kernel
    // state (the two integer divisions of chunk_len() sat on the critical path of every chunk of every tile)
    // (the producer state in shared memory already holds them: pstate[5 + 3 f] floor size, pstate[6 + 3 f] remainder)
    // tile cursor of the rectangular schedule: advanced by (gridDim / col_panels, gridDim % col_panels) with a carry
    int cur_r = 0, cur_c = 0, step_r = 0, step_c = 0;
    if constexpr (!SYM) {
        cur_r = (int)((long long)blockIdx.x / col_panels); cur_c = (int)((long long)blockIdx.x % col_panels);
        step_r = (int)((long long)gridDim.x / col_panels); step_c = (int)((long long)gridDim.x % col_panels);
    }
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#ifdef LSK_TRACE
        ++tile_no;
        int chunk_no = 0;
#endif
        LSK_GSTAMP(0);
        int pr, pc;
        if constexpr (SYM) tile_coords(t, pr, pc);
        else {
            pr = cur_r; pc = cur_c;
            cur_r += step_r; cur_c += step_c;
            if (cur_c >= col_panels) { cur_c -= col_panels; ++cur_r; }
        }
        double acc[NF][MT][2][2];
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) { acc[f][m][n][0] = 0.0; acc[f][m][n][1] = 0.0; }
        // SYRK: the entries this thread will update are requested now and consumed after the DMMA phase, so the
        // read half of the read-modify-write hides behind the tile's tensor-core work
        double prev[EPI == LSK_EPI_SYRK ? MT : 1][2][2];
        if constexpr (EPI == LSK_EPI_SYRK) {
            const int r_base = pr * BMT + wr * (8 * MT) + (lane >> 2), c_base = pc * BN + wc * 16 + 2 * (lane & 3);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) {
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

        // ---- bilinear forms on the tensor-core path: every form is cut into chunks of <= KG_PER_STAGE k-groups,
        //      one pipeline stage each; inside a stage all offsets are immediates ----
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int cut_cnt = (p.group_end[f] - p.group_begin[f] + KG_PER_STAGE - 1) / KG_PER_STAGE;     // constant divisor: a shift
            const int cut_base = pstate[5 + 3 * f], cut_rem = pstate[6 + 3 * f];
            for (int ci = 0; ci < cut_cnt; ++ci) {
                const int ng = cut_base + (ci < cut_rem ? 1 : 0);
                produce_ahead();
                ensure_issued(qc);
                LSK_GSTAMP(1 + 2 * chunk_no);
                mbar_wait(&full_bar[stage], phase);
                LSK_GSTAMP(2 + 2 * chunk_no);
#ifdef LSK_TRACE
                ++chunk_no;
#endif
                const double* pa = sA + stage * A_STAGE + a_off;
                const double* pb = sB + stage * B_STAGE + b_off;
                if (ng == KG_PER_STAGE) {
#pragma unroll
                    for (int gi = 0; gi < KG_PER_STAGE; ++gi) {
                        double a[MT], b[2];
#pragma unroll
                        for (int m = 0; m < MT; ++m) a[m] = pa[gi * (PANEL * 4) + m * 32];
#pragma unroll
                        for (int n = 0; n < 2; ++n) b[n] = pb[gi * (PANEL * 4) + n * 32];
#pragma unroll
                        for (int m = 0; m < MT; ++m)
#pragma unroll
                            for (int n = 0; n < 2; ++n) dmma884(acc[f][m][n], a[m], b[n]);
                    }
                } else {
#pragma unroll
                    for (int gi = 0; gi < KG_PER_STAGE - 1; ++gi) {
                        if (gi < ng) {
                            double a[MT], b[2];
#pragma unroll
                            for (int m = 0; m < MT; ++m) a[m] = pa[gi * (PANEL * 4) + m * 32];
#pragma unroll
                            for (int n = 0; n < 2; ++n) b[n] = pb[gi * (PANEL * 4) + n * 32];
#pragma unroll
                            for (int m = 0; m < MT; ++m)
#pragma unroll
                                for (int n = 0; n < 2; ++n) dmma884(acc[f][m][n], a[m], b[n]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);      // hand the stage back to the producer
                if (++stage == NST) { stage = 0; phase ^= 1; }
                ++qc;
            }
        }

        LSK_GSTAMP(13);
        // ---- class polynomial on the accumulator fragments, 8 entries at a time ----
        const int row0 = pr * BMT + wr * (8 * MT) + (lane >> 2);
        const int col0 = pc * BN + wc * 16 + 2 * (lane & 3);
        // Gram / rank-k-update tiles that lie fully inside a row-major output: straight 16-byte stores at constant offsets
        // from one base pointer (the general path below pays ~25 instructions of predicates and 64-bit addressing per pair,
        // a measurable share of a K = 128 tile)
        const bool interior_tile = vec_ok && p.out_layout == LSK_OUT_ROW_MAJOR && pr * BMT + BMT <= n_rows && pc * BN + BN <= n_cols;
        if constexpr ((EPI == LSK_EPI_LINEAR && !SYM) || EPI == LSK_EPI_SYRK) {
            if (interior_tile) {
                produce_ahead();
                double* base = out + (size_t)row0 * ld_out + col0;
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        double r0, r1;
                        if constexpr (EPI == LSK_EPI_SYRK) {
                            r0 = fma(p.scale, acc[0][m][n][0], prev[m][n][0]);
                            r1 = fma(p.scale, acc[0][m][n][1], prev[m][n][1]);
                        } else {
                            r0 = p.scale * acc[0][m][n][0];
                            r1 = p.scale * acc[0][m][n][1];
                        }
                        st_cs_f64x2(base + (size_t)(m * 8) * ld_out + n * 8, r0, r1);
                    }
                continue;
            }
        }
        // writes the pair (r0, r1) at (row, col + col_shift), (row, col + col_shift + 1); `col` is even
        auto store_pair = [&](int row, int col, long long col_shift, double r0, double r1) {
            if (row >= n_rows || col >= n_cols) return;
            const long long c = col_shift + col;
            double* dst;
            bool pair_ok;
            if (p.out_layout == LSK_OUT_PANEL_MAJOR) {
                dst = out + (((size_t)(row >> 6) * p.out_kg + (size_t)(c >> 2)) * PANEL + (row & 63)) * 4 + (c & 3);
                pair_ok = (c & 1) == 0;
            } else {
                dst = out + (size_t)row * ld_out + c;
                pair_ok = vec_ok && ((c & 1) == 0);
            }
            if constexpr (SYM && EPI != LSK_EPI_SYRK) {          // mirror (row-major only): entries strictly above the diagonal
                if (col > row) st_cs_f64(out + (size_t)col * ld_out + row, r0);
                if (col + 1 > row && col + 1 < n_cols) st_cs_f64(out + (size_t)(col + 1) * ld_out + row, r1);
            }
            if (pair_ok && col + 1 < n_cols) st_cs_f64x2(dst, r0, r1);
            else {
                st_cs_f64(dst, r0);
                if (col + 1 < n_cols) {
                    const long long c1 = c + 1;
                    double* d1 = (p.out_layout == LSK_OUT_PANEL_MAJOR)
                        ? out + (((size_t)(row >> 6) * p.out_kg + (size_t)(c1 >> 2)) * PANEL + (row & 63)) * 4 + (c1 & 3)
                        : dst + 1;
                    st_cs_f64(d1, r1);
                }
            }
        };
#pragma unroll
        for (int h = 0; h < MT / 2; ++h) {
            constexpr int R = 8;
            produce_ahead();
            double var[NV][R], res[R];
#pragma unroll
            for (int e = 0; e < R; ++e) {
                const int ml = e >> 2, n = (e >> 1) & 1, i = e & 1;
                double u[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) u[f] = acc[f][2 * h + ml][n][i];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if constexpr (EPI == LSK_EPI_LINEAR || EPI == LSK_EPI_SYRK) var[v][e] = u[0];
                    else {
                        // forms flagged in bits 8.. of `reserved` enter squared (sign-ambiguous spinor form)
                        double x = p.aff[v][0];
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            // generated epilogues know which forms enter which variable: zero terms cost nothing
                            if constexpr (GEN > 0) { if (!((LSK_GEN_AFFMASK[GEN] >> (8 * v + f)) & 1ull)) continue; }
                            const double w = ((p.reserved >> (8 + f)) & 1) ? u[f] * u[f] : u[f];
                            x = fma(p.aff[v][1 + f], w, x);
                        }
                        var[v][e] = x;
                    }
                }
            }
            if constexpr (EPI == LSK_EPI_FEAT_CHEB1 || EPI == LSK_EPI_FEAT_HORNER) {
                // one output column block per irrep: Phi[i, l * feat_stride + j]
                int pos = 0;
                for (int l = 0; l < p.nterms; ++l) {
                    const int len = p.row_len[l];
                    if constexpr (EPI == LSK_EPI_FEAT_CHEB1) epi_cheb1<R>(p.coef + pos, len, var[0], res);
                    else epi_tape<R, NV>(p.coef + 2 * pos, len, var, res);
                    pos += len;
#pragma unroll
                    for (int e = 0; e < R; e += 2) {
                        const int m = 2 * h + (e >> 2), n = (e >> 1) & 1;
                        store_pair(row0 + m * 8, col0 + n * 8, (long long)l * p.feat_stride, p.scale * res[e], p.scale * res[e + 1]);
                    }
                }
            } else {
                if constexpr (EPI == LSK_EPI_CHEB1) epi_cheb1<R>(p.coef, p.nterms, var[0], res);
                else if constexpr (EPI == LSK_EPI_STAIR2) epi_stair2<R>(p, var[0], var[1], res);
                else if constexpr (EPI == LSK_EPI_HORNER) {
                    if constexpr (GEN > 0) gen_epilogue<GEN, R, NV>(p, var, res);
                    else epi_tape<R, NV>(p.coef, p.nterms, var, res);
                }
                else if constexpr (EPI == LSK_EPI_LINEAR || EPI == LSK_EPI_SYRK) {
#pragma unroll
                    for (int e = 0; e < R; ++e) res[e] = var[0][e];
                } else {
                    // ORBIT2 keeps a Chebyshev table per entry: run it 2 entries at a time to bound registers
#pragma unroll
                    for (int e2 = 0; e2 < R; e2 += 2) {
                        double a1[2] = {var[0][e2], var[0][e2 + 1]}, a2[2] = {var[1][e2], var[1][e2 + 1]}, o[2];
                        epi_orbit2<2, D>(p, a1, a2, o);
                        res[e2] = o[0]; res[e2 + 1] = o[1];
                    }
                }
                if (!SYM && interior_tile) {          // tile fully inside a row-major output: no predicates, constant offsets
                    double* base = out + (size_t)row0 * ld_out + col0;
#pragma unroll
                    for (int e = 0; e < R; e += 2) {
                        const int m = 2 * h + (e >> 2), n = (e >> 1) & 1;
                        st_cs_f64x2(base + (size_t)(m * 8) * ld_out + n * 8, p.scale * res[e], p.scale * res[e + 1]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < R; e += 2) {
                        const int m = 2 * h + (e >> 2), n = (e >> 1) & 1;
                        if constexpr (EPI == LSK_EPI_SYRK)
                            store_pair(row0 + m * 8, col0 + n * 8, 0, fma(p.scale, res[e], prev[m][n][0]), fma(p.scale, res[e + 1], prev[m][n][1]));
                        else
                            store_pair(row0 + m * 8, col0 + n * 8, 0, p.scale * res[e], p.scale * res[e + 1]);
                    }
                }
            }
        }
    }
}

template <int EPI, int NF, int NV, int D, bool SYM = false, int MT = 4, int GEN = 0>
static int launch(const double* A, const double* B, int n_rows, int n_cols, int kg_total, double* out, long long ld_out,
                  const LskEpilogue& p, int max_ctas, cudaStream_t stream) {
    auto kern = pairwise_kernel<EPI, NF, NV, D, SYM, MT, GEN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<MT>::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const long long rt = (n_rows + 32 * MT - 1) / (32 * MT), cp = (n_cols + BN - 1) / BN;
    const long long tiles = SYM ? rt * cp - rt * (rt - 1) : rt * cp;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, TileCfg<MT>::SMEM_BYTES);
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)max_ctas * per_sm;
    int grid = (int)(tiles < cap ? tiles : cap);
    long long cpt = 0;                                  // chunks per tile; the CTA-local chunk sequence number is an int
    for (int f = 0; f < NF; ++f) cpt += (p.group_end[f] - p.group_begin[f] + KG_PER_STAGE - 1) / KG_PER_STAGE;
    if (((tiles + grid - 1) / grid) * cpt > 0x7fffffffLL) return LSK_ESIZE;
    return (int)launch_pdl(kern, dim3(grid), dim3(THREADS), (size_t)TileCfg<MT>::SMEM_BYTES, stream, A, B, n_rows, n_cols, kg_total, out, ld_out, p);
}

}  // namespace lsk

extern "C" int lsk_pairwise_poly(const void* A, const void* B, long long n_rows, long long n_cols, int kg_total,
                                 const LskEpilogue* epi, void* out, long long ld_out, void* stream) {
    using namespace lsk;
    if (!A || !B || !out || !epi) return LSK_EARG;
    if (n_rows < 0 || n_cols < 0 || kg_total <= 0) return LSK_ESIZE;
    if (epi->out_layout == LSK_OUT_ROW_MAJOR && ld_out < n_cols) return LSK_ESIZE;
    if (epi->out_layout == LSK_OUT_PANEL_MAJOR && epi->out_kg < 1) return LSK_ESIZE;
    if (epi->out_layout != LSK_OUT_ROW_MAJOR && epi->out_layout != LSK_OUT_PANEL_MAJOR) return LSK_EARG;
    if (((epi->reserved >> 1) & 1) && (n_rows != n_cols || epi->out_layout != LSK_OUT_ROW_MAJOR)) return LSK_EARG;
    if (n_rows == 0 || n_cols == 0) return 0;
    if (n_rows > 0x7fffffffLL || n_cols > 0x7fffffffLL) return LSK_ESIZE;
    const LskEpilogue& p = *epi;
    if (p.nforms < 1 || p.nforms > LSK_MAX_FORMS || p.nvars < 1 || p.nvars > LSK_MAX_VARS) return LSK_EARG;
    for (int f = 0; f < p.nforms; ++f)
        if (p.group_begin[f] < 0 || p.group_end[f] > kg_total || p.group_begin[f] >= p.group_end[f]) return LSK_EARG;
    for (int f = 1; f < p.nforms; ++f)
        if (p.group_begin[f] != p.group_end[f - 1]) return LSK_EARG;   // forms tile the k range in order
    // kg_total is the operand row stride; the forms may stop short of it (SYRK on the first half of a two-panel operand,
    // LINEAR on a column block of a feature map)
    if (p.group_begin[0] != 0 || p.group_end[p.nforms - 1] > kg_total) return LSK_EARG;
    if (p.kind != LSK_EPI_SYRK && p.kind != LSK_EPI_LINEAR && p.group_end[p.nforms - 1] != kg_total) return LSK_EARG;

    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* a = static_cast<const double*>(A);
    const double* b = static_cast<const double*>(B);
    double* o = static_cast<double*>(out);
    const int nr = (int)n_rows, nc = (int)n_cols;

#define LSK_GO(EPI, NF, NV, D) return launch<EPI, NF, NV, D, false, ((NF) >= 3 ? 2 : 4)>(a, b, nr, nc, kg_total, o, ld_out, p, sms, st)
#define LSK_GO_SYM(EPI, NF, NV, D) do { if (want_sym) return launch<EPI, NF, NV, D, true>(a, b, nr, nc, kg_total, o, ld_out, p, sms, st); \
                                        return launch<EPI, NF, NV, D, false>(a, b, nr, nc, kg_total, o, ld_out, p, sms, st); } while (0)
    const bool want_sym = ((p.reserved >> 1) & 1) != 0;
    switch (p.kind) {
    case LSK_EPI_CHEB1:
        if (p.nterms < 1 || p.nterms > LSK_MAX_COEF || p.nvars != 1) return LSK_EARG;
        if (p.nforms == 1) LSK_GO_SYM(LSK_EPI_CHEB1, 1, 1, 1);
        return LSK_EUNSUPPORTED;
    case LSK_EPI_STAIR2: {
        if (p.nterms < 1 || p.nterms > LSK_STAIR_MAX_ROWS || p.nvars != 2) return LSK_EARG;
        for (int k = 0; k < LSK_STAIR_MAX_ROWS; ++k) {
            const int pairs = (int)((p.stair_shape >> (4 * k)) & 15ull);
            if (k < p.nterms ? (pairs < 1 || 2 * pairs > LSK_STAIR_STRIDE || p.row_len[k] != 2 * pairs) : pairs != 0) return LSK_EARG;
        }
        if (p.nforms == 2) {
            // the SO(5) tables of the common truncations have a dedicated kernel (lsk_pairwise_so5.cu)
            const int rc = so5_stair_launch(a, b, nr, nc, kg_total, o, ld_out, p, sms, st);
            if (rc != LSK_EUNSUPPORTED) return rc;
            LSK_GO_SYM(LSK_EPI_STAIR2, 2, 2, 1);
        }
        return LSK_EUNSUPPORTED;
    }
    case LSK_EPI_ORBIT2:
        if (p.nvars != 2 || p.nforms != 2 || p.nterms < 1) return LSK_EARG;
        if (p.nterms <= 8) LSK_GO(LSK_EPI_ORBIT2, 2, 2, 8);
        if (p.nterms <= 12) LSK_GO(LSK_EPI_ORBIT2, 2, 2, 12);
        if (p.nterms <= 16) LSK_GO(LSK_EPI_ORBIT2, 2, 2, 16);
        if (p.nterms <= 20) LSK_GO(LSK_EPI_ORBIT2, 2, 2, 20);
        return LSK_ESIZE;
    case LSK_EPI_HORNER: {
        if (p.nterms < 1 || 2 * p.nterms > LSK_MAX_COEF) return LSK_ESIZE;
        // a tape whose structure was known at build time runs as generated straight-line code
        const unsigned long long h = gen_hash(p);
        for (const GenEntry& g : LSK_GEN_TABLE) {
            if (g.hash != h || g.nsteps != p.nterms || g.nforms != p.nforms || g.nvars != p.nvars) continue;
            bool map_ok = true;                 // the affine map must vanish wherever the generated code skips a term
            for (int v = 0; v < p.nvars; ++v)
                for (int f = 0; f < p.nforms; ++f)
                    if (!((g.affmask >> (8 * v + f)) & 1ull) && p.aff[v][1 + f] != 0.0) map_ok = false;
            if (!map_ok) continue;
            {
                // SU(4) has a dedicated kernel (lsk_pairwise_su4.cu); LSK_SU4_GENERIC=1 keeps the generic one (development)
                static const bool generic_only = [] { const char* s = getenv("LSK_SU4_GENERIC"); return s && atoi(s) > 0; }();
                if (!generic_only) {
                    const int rc = su4_tile_launch(g.id, a, b, nr, nc, kg_total, o, ld_out, p, sms, st);
                    if (rc != LSK_EUNSUPPORTED) return rc;
                }
            }
            switch (g.id) {
#define LSK_GEN_CASE(ID, NFV, NVV) case ID: return launch<LSK_EPI_HORNER, NFV, NVV, 1, false, ((NFV) >= 3 ? 2 : 4), ID>(a, b, nr, nc, kg_total, o, ld_out, p, sms, st);
                LSK_GEN_LIST(LSK_GEN_CASE)
#undef LSK_GEN_CASE
            default: break;
            }
        }
        if (p.nforms == 1 && p.nvars == 1) LSK_GO(LSK_EPI_HORNER, 1, 1, 1);
        if (p.nforms == 2 && p.nvars == 2) LSK_GO(LSK_EPI_HORNER, 2, 2, 1);
        if (p.nforms == 3 && p.nvars == 3) LSK_GO(LSK_EPI_HORNER, 3, 3, 1);
        if (p.nforms == 4 && p.nvars == 4) LSK_GO(LSK_EPI_HORNER, 4, 4, 1);
        if (p.nforms == 4 && p.nvars == 3) LSK_GO(LSK_EPI_HORNER, 4, 3, 1);        // SU(4): three-form complex trace + real e_2
        if (p.nforms == 5 && p.nvars == 5) LSK_GO(LSK_EPI_HORNER, 5, 5, 1);
        return LSK_EUNSUPPORTED;
    }
    case LSK_EPI_LINEAR:
        if (p.nforms != 1 || p.nvars != 1) return LSK_EARG;
        {
            // row-major Grams run on the dedicated kernel (lsk_gram.cu); LSK_GRAM_TILE=-1 keeps the generic tile kernel (development)
            static const bool generic_only = [] { const char* s = getenv("LSK_GRAM_TILE"); return s && atoi(s) < 0; }();
            if (p.out_layout == LSK_OUT_ROW_MAJOR && !generic_only)
                return gram_launch(a, b, nr, nc, kg_total, p.group_end[0], o, ld_out, p.scale, want_sym, sms, st);
        }
        LSK_GO_SYM(LSK_EPI_LINEAR, 1, 1, 1);
    case LSK_EPI_SYRK:
        if (p.nforms != 1 || p.nvars != 1 || p.out_layout != LSK_OUT_ROW_MAJOR) return LSK_EARG;
        {
            // rank-k updates run on the Gram kernel's pipeline (lsk_gram.cu, read-modify-write mode); LSK_GRAM_TILE=-1: generic kernel
            static const bool generic_only = [] { const char* s = getenv("LSK_GRAM_TILE"); return s && atoi(s) < 0; }();
            if (!generic_only)
                return gram_update_launch(a, b, nr, nc, kg_total, p.group_end[0], o, ld_out, p.scale, want_sym, sms, st);
        }
        LSK_GO_SYM(LSK_EPI_SYRK, 1, 1, 1);
    case LSK_EPI_SPARSE:
    case LSK_EPI_FEAT_SPARSE:
        return LSK_EUNSUPPORTED;        // monomial-list encoding: only lsk_homog_pairs reads it
    case LSK_EPI_FEAT_CHEB1:
    case LSK_EPI_FEAT_HORNER: {
        if (p.nterms < 1 || p.nterms > LSK_MAX_ROWS || p.feat_stride < n_cols) return LSK_EARG;
        int tot = 0;
        for (int l = 0; l < p.nterms; ++l) { if (p.row_len[l] < 1) return LSK_EARG; tot += p.row_len[l]; }
        if (p.kind == LSK_EPI_FEAT_CHEB1) {
            if (tot > LSK_MAX_COEF || p.nvars != 1 || p.nforms != 1) return LSK_ESIZE;
            LSK_GO(LSK_EPI_FEAT_CHEB1, 1, 1, 1);
        }
        if (2 * tot > LSK_MAX_COEF) return LSK_ESIZE;
        if (p.nforms == 2 && p.nvars == 2) LSK_GO(LSK_EPI_FEAT_HORNER, 2, 2, 1);
        if (p.nforms == 3 && p.nvars == 3) LSK_GO(LSK_EPI_FEAT_HORNER, 3, 3, 1);
        if (p.nforms == 4 && p.nvars == 4) LSK_GO(LSK_EPI_FEAT_HORNER, 4, 4, 1);
        if (p.nforms == 4 && p.nvars == 3) LSK_GO(LSK_EPI_FEAT_HORNER, 4, 3, 1);
        if (p.nforms == 5 && p.nvars == 5) LSK_GO(LSK_EPI_FEAT_HORNER, 5, 5, 1);
        return LSK_EUNSUPPORTED;
    }
    default:
        return LSK_EARG;
    }
#undef LSK_GO
#undef LSK_GO_SYM
}
