// Blocked fp64 Cholesky factorisation, triangular solves and inverse for the GP regression step that consumes the covariance
// (the reference hands K + sigma_n^2 I to gpytorch's ExactGP, examples/gpr_model.py:26-39, whose
// marginal likelihood / prediction is a LAPACK/cuSOLVER potrf + potrs, and whose gradient needs tr(K^{-1} dK)).
//
//   A = U^T U,  U upper triangular, factorised IN PLACE in the upper triangle of a row-major matrix (equivalently: the
//   lower factor L = U^T of the same matrix read column-major).  Right-looking, 128-column panels, two panels per
//   trailing update (lsk_potrf_upper):
//     diag_tile_kernel   one CTA: U_kk = chol(A_kk) and W_kk = U_kk^{-T} on 16 x 16 tiles (register-resident elimination of the
//                   diagonal tile by one warp, DMMA tile products by the others, look-ahead; see the kernel)
//     trsm_kernel   row panel U_k,: = W_kk^T A_k,: on the fp64 tensor-core path (DMMA), written twice: row-major into A
//                   (the factor) and panel-major into the workspace (the operand layout of the pairwise kernel)
//     pairwise_kernel<SYRK>  strip / trailing update A_ij -= U_ki^T U_kj (K = 128 / 256) on the tiles that reach the
//                   diagonal or lie above it (lsk_pairwise.cu: TMA ring + DMMA, read-modify-write epilogue)
//   The inverses W_kk stay in the workspace:
//     lsk_potrs_upper  U^T y = b, U x = y: one launch per sweep (sweep_kernel): every block owned by one CTA, solved with its
//                      inverse block, the blocks pipelined across the SMs through flags - no divisions, no grid-wide barriers
//     lsk_trsm_lower   the forward solve for a block of right-hand sides (trsm_kernel + repack_kernel + tile kernel)
//     lsk_potri_upper  Y = U^{-T} by the same sweep on the identity and A^{-1} = sum_k Y_k^T Y_k on the triangular tile
//                      schedule: 2 n^3 / 3 flop, the zero structure of U^{-T} used exactly at block level
#include <cstdlib>

#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"

extern "C" int lsk_pairwise_poly(const void* A, const void* B, long long n_rows, long long n_cols, int kg_total,
                                 const LskEpilogue* epi, void* out, long long ld_out, void* stream);

namespace lsk {
namespace chol {

constexpr int NB = 128;                 // block size = K of the trailing update (32 k-groups)
constexpr int WBLK = NB * NB;           // doubles per stored inverse block

// W block layout ("k-group major", the A-operand layout of trsm_kernel):  B[kk][q] = (U^{-T})[kk][q] = W[q][kk]
// sits at  (q / 4) * 512 + kk * 4 + q % 4.
__device__ __forceinline__ int wg_index(int kk, int q) { return (q >> 2) * (NB * 4) + kk * 4 + (q & 3); }

// 1 / d to within an ulp or two in 5 instructions (hardware seed + two Newton steps); the IEEE division sequence is ~4x
// longer and sits on the critical path of every elimination step, in every thread
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}

// ------------------------------------------------------------------------------------------------------------
// diagonal block:  U_kk = chol(A_kk) and B = U_kk^{-T} of one 128 x 128 block, one CTA, the dependent chain cut to what it has to be
//
// (Round 1 swept the block with 2 x 128 rank-1 steps, each paying a CTA-wide barrier, a shared-memory round trip and a
// reciprocal on its dependent path: ~190 ns per step, 52 us per block; this kernel: 23.7 us, profiles/diag_trace_r02.txt.)
// The block is cut into 16 x 16 tiles (the 36 on or above the diagonal, row stride 18 in shared memory), and per tile row jb:
//   S1  one warp eliminates the 16 x 16 diagonal tile in registers: lane c < 16 owns column c of the (symmetric) tile, lane
//       16 + c owns column c of an identity that receives the same row operations.  The pivot row is spread over the lanes and
//       all-gathered through shared memory; the update is  v_r -= (D_jr D_jc) / d_j  with the product formed before 1 / d_j is
//       known, so a step's dependent path is store -> load -> reciprocal -> one fused multiply-add, and there is no barrier inside
//       the 16 steps.  Out come U_jj = diag(d^-1/2) R and, from the identity side, its inverse transpose
//       B_jj = diag(d^-1/2) L_unit^{-1} - no separate inversion of the diagonal tiles.
//   S2  the tile row right of it, U_j,: = B_jj A_j,: - one 16 x 16 x 16 product per warp on the fp64 tensor-core path,
//   S3a the next tile row receives its rank-16 update (so that S1 of the next tile can start),
//   S3b WHILE warp 0 runs S1 of the next tile, six helper warps update the rest of the trailing tiles and compute tile row
//       jb of B by block forward substitution, B(i, j) = -B(i, i) sum_k U(k, i)^T B(k, j) (one warp per tile).  The seventh warp
//       shares warp 0's sub-partition and idles: its DMMAs would delay every fp64 instruction of the dependent chain.
//   The factor and the non-zero half of B go to global memory tile by tile as they are finished (the zeros of the W block come
//   from a memset of the workspace): a lone SM drains 17-24 bytes per clock, 205 KB at the end of the kernel cost 12 000 cycles.
// Every tile product is DMMA.8x8x4: with plain DFMA a 2 x 4 register block per lane needs three 16-byte shared-memory loads
// per eight multiply-adds and the ONE shared-memory pipe of the SM (not the fp64 pipe) sets the pace - measured 1 480 cycles
// per tile product; the DMMA fragments need four 8-byte loads per 1 024 multiply-adds.
// ------------------------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 256;
constexpr int TS = 18;                           // row stride of a 16 x 16 tile in shared memory (doubles)
constexpr int TILE = 16 * TS;
constexpr int NTL = NB / 16;                     // tiles per side
constexpr int TRI = NTL * (NTL + 1) / 2;         // tiles in a triangle
constexpr int DIAGT_SMEM = ((2 * TRI + 16) * TILE + NB + 64) * 8;

__device__ __forceinline__ int ut(int bi, int bj) { return bj * (bj + 1) / 2 + bi; }     // upper tile (bi <= bj)
__device__ __forceinline__ int lt(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }     // lower tile (bi >= bj)

// 1 / sqrt(x) for normal positive x: hardware seed + a cubic and a linear Newton step
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    const double e2 = fma(-x * y, y, 1.0);
    return fma(0.5 * y, e2, y);
}

// Store under a lane-dependent condition as ONE predicated instruction.  Written as `if (p) *ptr = v` the compiler emitted a
// divergent branch per store, and sixteen of those in a row cost the lone warp of step S1 5 900 cycles (tools/s1_micro.cu).
__device__ __forceinline__ void st_global_if(double* ptr, double v, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f64 [%0], %1; }" :: "l"(ptr), "d"(v), "r"((int)p) : "memory");
}

// Accumulator of a 16 x 16 tile held by one warp as 2 x 2 DMMA fragments: acc[a][b][e] = C[8 a + g][8 b + 2 tig + e],
// g = lane / 4, tig = lane % 4.
struct TileAcc {
    double c[2][2][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) c[a][b][0] = c[a][b][1] = 0.0;
    }
    __device__ __forceinline__ void load(const double* __restrict__ C, int g, int tig) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double2 v = *reinterpret_cast<const double2*>(C + (8 * a + g) * TS + 8 * b + 2 * tig);
                c[a][b][0] = v.x; c[a][b][1] = v.y;
            }
    }
    __device__ __forceinline__ void store(double* __restrict__ C, int g, int tig) const {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                *reinterpret_cast<double2*>(C + (8 * a + g) * TS + 8 * b + 2 * tig) = make_double2(c[a][b][0], c[a][b][1]);
    }
    // C (+/-)= P^T Q:  C[i][j] += sum_k P[k][i] Q[k][j]
    template <bool NEG>
    __device__ __forceinline__ void mma_tn(const double* __restrict__ P, const double* __restrict__ Q, int g, int tig) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            double a0 = P[(k + tig) * TS + g], a1 = P[(k + tig) * TS + 8 + g];
            const double b0 = Q[(k + tig) * TS + g], b1 = Q[(k + tig) * TS + 8 + g];
            if (NEG) { a0 = -a0; a1 = -a1; }
            dmma884(c[0][0], a0, b0); dmma884(c[0][1], a0, b1);
            dmma884(c[1][0], a1, b0); dmma884(c[1][1], a1, b1);
        }
    }
    // C (+/-)= P Q:  C[i][j] += sum_k P[i][k] Q[k][j]
    template <bool NEG>
    __device__ __forceinline__ void mma_nn(const double* __restrict__ P, const double* __restrict__ Q, int g, int tig) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            double a0 = P[g * TS + k + tig], a1 = P[(8 + g) * TS + k + tig];
            const double b0 = Q[(k + tig) * TS + g], b1 = Q[(k + tig) * TS + 8 + g];
            if (NEG) { a0 = -a0; a1 = -a1; }
            dmma884(c[0][0], a0, b0); dmma884(c[0][1], a0, b1);
            dmma884(c[1][0], a1, b0); dmma884(c[1][1], a1, b1);
        }
    }
};

#ifdef LSK_TRACE
__device__ long long g_diag_trace[8][80];
#define DIAG_STAMP(ev) do { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); if (lane == 0) g_diag_trace[warp][ev] = c_; } while (0)
#else
#define DIAG_STAMP(ev) do { } while (0)
#endif

__global__ void __launch_bounds__(DIAG_THREADS, 1)
diag_tile_kernel(double* __restrict__ A, long long lda, int k0, int nb, double* __restrict__ Wg, int* __restrict__ info)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Up = reinterpret_cast<double*>(smem_raw);         // 36 upper tiles: A_kk, then U_kk
    double* Bp = Up + TRI * TILE;                             // 36 lower tiles: B = U_kk^{-T}
    double* Tt = Bp + TRI * TILE;                             // 16 scratch tiles of the doubling steps
    double* rsv = Tt + 16 * TILE;                             // 1 / U_jj
    double* piv = rsv + NB;                                   // 2 x 32: the pivot row of S1, double-buffered
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int g = lane >> 2, tig = lane & 3;                  // DMMA fragment coordinates
    DIAG_STAMP(0);

    // ---- load the upper tiles (diagonal tiles mirrored: S1 works on the full symmetric tile); identity padding when ragged
    {
        const int r = t >> 4, c = t & 15;
        double tmp[TRI];                                      // all 36 loads in flight before the first store
        if (nb == NB) {                                       // full block: one multiply-add of address arithmetic per load
            const size_t row16 = 16 * (size_t)lda;
            const double* p = A + (size_t)(k0 + r) * lda + k0 + c;
            const double* pm = (c < r) ? A + (size_t)(k0 + c) * lda + k0 + r : p;      // mirrored position inside a diagonal tile
#pragma unroll
            for (int bj = 0; bj < NTL; ++bj)
#pragma unroll
                for (int bi = 0; bi <= bj; ++bi) tmp[bj * (bj + 1) / 2 + bi] = ((bi == bj) ? pm : p)[bi * row16 + 16 * bj];
        } else {
#pragma unroll
            for (int bj = 0; bj < NTL; ++bj)
#pragma unroll
                for (int bi = 0; bi <= bj; ++bi) {
                    int R = 16 * bi + r, C = 16 * bj + c;
                    if (bi == bj && c < r) { const int x = R; R = C; C = x; }
                    const bool in = R < nb && C < nb;             // ragged block: read a valid entry, then select
                    const double x = A[(size_t)(k0 + (in ? R : 0)) * lda + k0 + (in ? C : 0)];
                    tmp[bj * (bj + 1) / 2 + bi] = in ? x : ((R == C) ? 1.0 : 0.0);
                }
        }
#pragma unroll
        for (int i = 0; i < TRI; ++i) Up[i * TILE + r * TS + c] = tmp[i];
    }
    __syncthreads();
    DIAG_STAMP(1);

    auto inverse_row = [&](int bi, int first, int step) {
        double* scratch = Tt + warp * TILE;
        for (int bj = first; bj < bi; bj += step) {
            TileAcc acc;
            acc.zero();
            for (int k = bj; k < bi; ++k) acc.mma_tn<false>(Up + ut(k, bi) * TILE, Bp + lt(k, bj) * TILE, g, tig);
            acc.store(scratch, g, tig);
            __syncwarp();
            TileAcc x;
            x.zero();
            x.mma_nn<true>(Bp + lt(bi, bi) * TILE, scratch, g, tig);
            __syncwarp();
            x.store(Bp + lt(bi, bj) * TILE, g, tig);
            const int kk0 = 16 * bi + g, q0 = 16 * bj + 2 * tig;                 // and into the W block (fragment pairs are adjacent there)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
                    *reinterpret_cast<double2*>(Wg + wg_index(kk0 + 8 * a, q0 + 8 * b)) = make_double2(x.c[a][b][0], x.c[a][b][1]);
        }
    };

#pragma unroll 1
    for (int jb = 0; jb < NTL; ++jb) {
        const int o = 16 * jb;
        if (warp == 0) {
            // ---- S1: diagonal tile (lanes 0-15) and the identity that receives the same row operations (lanes 16-31).
            // Nothing in here branches on the lane: 1 450 cycles for the 16 steps (tools/s1_micro.cu).
            double* Dt = Up + ut(jb, jb) * TILE;
            const int c = lane & 15;
            const bool iside = lane >= 16;
            double v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const double x = Dt[r * TS + c];
                v[r] = iside ? ((r == c) ? 1.0 : 0.0) : x;
            }
            double dc = 1.0;
            int bad = -1;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                // the pivot row D[j][:] is spread over the lanes (lane c holds D[j][c] in v[j]): all-gather through shared memory
                double* pb = piv + (j & 1) * 32;
                pb[lane] = v[j];                                  // identity side: scratch slots 16 .. 31
                __syncwarp();
                double row[16];
#pragma unroll
                for (int r = (j & ~1); r < 16; r += 2) {
                    const double2 x = *reinterpret_cast<const double2*>(pb + r);
                    row[r] = x.x; row[r + 1] = x.y;
                }
                const double d = row[j];
                const double id = fast_rcp(d);
                dc = (c == j) ? d : dc;
                bad = (bad < 0 && !(d > 0.0) && o + j < nb) ? j : bad;
#pragma unroll
                for (int r = j + 1; r < 16; ++r) v[r] = fma(-(row[r] * v[j]), id, v[r]);
            }
            if (bad >= 0) {                                       // warp-uniform: not positive definite (LAPACK info convention)
                if (lane == 0) atomicCAS(info, 0, k0 + o + bad + 1);
            }
            rsv[o + c] = fast_rsqrt(dc);                          // both halves hold the same pivots
            __syncwarp();
            // rows scaled by d_r^-1/2: U[r][c] (kept for c >= r) / B[r][c] (zero for c > r by itself).  Global copies: U into A,
            // B into the W block (whose zeros come from the memset of the workspace).
            double* dst = Up + (iside ? (TRI + lt(jb, jb)) * TILE : ut(jb, jb) * TILE) + c;
            double* gdst = iside ? Wg + wg_index(o, o + c) : A + (size_t)(k0 + o) * lda + k0 + o + c;
            const long long gstep = iside ? 4 : lda;
            const bool col_ok = iside || o + c < nb;
#pragma unroll
            for (int r = 0; r < 16; r += 2) {                    // all loads before the first store (the stores are ordered asm)
                const double2 x = *reinterpret_cast<const double2*>(rsv + o + r);
                v[r] *= x.x; v[r + 1] *= x.y;
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                dst[r * TS] = (iside || r <= c) ? v[r] : 0.0;
                st_global_if(gdst + r * gstep, v[r], col_ok && (iside ? (r >= c) : (r <= c)));
            }
        } else if (jb > 0 && warp != 4) {
            // ---- S3b of the previous tile row: trailing tiles below the row that S3a already served.  Six helpers: warp 4 sits
            // on the sub-partition of warp 0 and its DMMAs would delay every fp64 instruction of the dependent chain of S1.
            const int hw = warp - (warp > 4 ? 2 : 1);             // 0 .. 5
            const int rows = NTL - 1 - jb, tiles = rows * (rows + 1) / 2;
            for (int idx = hw; idx < tiles; idx += 6) {
                int rem = idx, bi = jb + 1, len = rows;
                while (rem >= len) { rem -= len; ++bi; --len; }
                const int bj = bi + rem;
                TileAcc acc;
                double* Ct = Up + ut(bi, bj) * TILE;
                acc.load(Ct, g, tig);
                acc.mma_tn<true>(Up + ut(jb - 1, bi) * TILE, Up + ut(jb - 1, bj) * TILE, g, tig);
                acc.store(Ct, g, tig);
            }
            // ---- and tile row jb - 1 of B = U^{-T} (block forward substitution; everything it needs exists since S1 of that row):
            //      B(bi, bj) = -B(bi, bi) sum_{k = bj}^{bi - 1} L(bi, k) B(k, bj),   L(bi, k) = U(k, bi)^T.   One warp per tile.
            inverse_row(jb - 1, 5 - hw, 6);
        }
        DIAG_STAMP(2 + 6 * jb);
        __syncthreads();
        DIAG_STAMP(3 + 6 * jb);
        // ---- S2: tile row jb right of the diagonal, U_j,bj = B_jj A_j,bj (in place)
        if (warp < NTL - 1 - jb) {
            double* Xt = Up + ut(jb, jb + 1 + warp) * TILE;
            TileAcc acc;
            acc.zero();
            acc.mma_nn<false>(Bp + lt(jb, jb) * TILE, Xt, g, tig);
            __syncwarp();
            acc.store(Xt, g, tig);
            const int C0 = 16 * (jb + 1 + warp) + 2 * tig;   // the factor goes back to A as it is produced (columns < nb: rows are then < nb too)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    double* gp = A + (size_t)(k0 + o + 8 * a + g) * lda + k0 + C0 + 8 * b;
                    st_global_if(gp, acc.c[a][b][0], C0 + 8 * b < nb);
                    st_global_if(gp + 1, acc.c[a][b][1], C0 + 8 * b + 1 < nb);
                }
        }
        DIAG_STAMP(4 + 6 * jb);
        __syncthreads();
        DIAG_STAMP(5 + 6 * jb);
        // ---- S3a: rank-16 update of the next tile row
        if (warp < NTL - 1 - jb) {
            const int bi = jb + 1, bj = jb + 1 + warp;
            TileAcc acc;
            double* Ct = Up + ut(bi, bj) * TILE;
            acc.load(Ct, g, tig);
            acc.mma_tn<true>(Up + ut(jb, bi) * TILE, Up + ut(jb, bj) * TILE, g, tig);
            acc.store(Ct, g, tig);
        }
        DIAG_STAMP(6 + 6 * jb);
        __syncthreads();
        DIAG_STAMP(7 + 6 * jb);
    }

    DIAG_STAMP(50);

    // ---- last tile row of B (rows 0 .. 6 were produced beside S1 of the following rows)
    inverse_row(NTL - 1, warp, 8);
    DIAG_STAMP(55);
}
#ifdef LSK_TRACE
}  // namespace chol
}  // namespace lsk
extern "C" int lsk_debug_diag_trace(void* host) { return (int)cudaMemcpyFromSymbol(host, lsk::chol::g_diag_trace, sizeof(lsk::chol::g_diag_trace)); }
namespace lsk {
namespace chol {
#endif

// ------------------------------------------------------------------------------------------------------------
// row panel:  U[k0 + kk, j] = sum_q B[kk][q] A[k0 + q, j]   for the 64 columns j0 .. j0 + 63 of this CTA
// ------------------------------------------------------------------------------------------------------------
constexpr int TRSM_THREADS = 256;
constexpr int TRSM_SMEM = (WBLK + 32 * 64 * 4) * 8 + 16;

__global__ void __launch_bounds__(TRSM_THREADS, 1)
trsm_kernel(double* __restrict__ A, long long lda, long long n_rows, long long col0, long long col_end, int k0,
            const double* __restrict__ Wg, double* __restrict__ P,
            int kg_stride, int g_off)         // P: panel-major with kg_stride k-groups per row; this panel fills groups g_off .. g_off + 31
{                                             // columns [col0, col_end) of the 128 rows from k0 (rows >= n_rows read as zero)
    const long long n = col_end;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sW = reinterpret_cast<double*>(smem_raw);         // B[kk][q], k-group major (128 KB)
    double* sB = sW + WBLK;                                   // A tile, [q / 4][j][q % 4] (64 KB)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 64 * 4);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long j0 = col0 + 64ll * blockIdx.x;

    if (t == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, WBLK * 8);
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(sW + c * (WBLK / 4), Wg + c * (WBLK / 4), (WBLK / 4) * 8, bar);
    }
    {
        // all 32 loads of a thread are issued before the first store: the panel comes from HBM and this kernel is on the
        // serial path of the factorisation (with the loads consumed one k-group at a time it was 80 % load latency)
        const int j = t & 63;
        const bool in = j0 + j < n;
        double v[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = (t >> 6) + 4 * u;
#pragma unroll
            for (int e = 0; e < 4; ++e) v[u][e] = (in && k0 + 4 * g + e < n_rows) ? A[(size_t)(k0 + 4 * g + e) * lda + j0 + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = (t >> 6) + 4 * u;
            double2* dst = reinterpret_cast<double2*>(sB + g * 256 + j * 4);
            dst[0] = make_double2(v[u][0], v[u][1]);
            dst[1] = make_double2(v[u][2], v[u][3]);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // warp w: rows kk = 16 w .. 16 w + 15 (2 fragments) x 64 columns (8 fragments); B is lower triangular: q <= kk
    double acc[2][8][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    const int frag = (lane >> 2) * 4 + (lane & 3);
    const int g_end = 4 * warp + 4;
    for (int g = 0; g < g_end; ++g) {
        double af[2], bf[8];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) af[mi] = sW[g * (NB * 4) + (16 * warp + 8 * mi) * 4 + frag];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) bf[ni] = sB[g * 256 + (8 * ni) * 4 + frag];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
    }
    const bool vec_ok = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((j0 & 1) == 0);
    double* Pp = P + ((size_t)blockIdx.x * kg_stride + g_off) * 256;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            const int kk = 16 * warp + 8 * mi + (lane >> 2), j = 8 * ni + 2 * (lane & 3);
            const double v0 = acc[mi][ni][0], v1 = acc[mi][ni][1];
            Pp[(kk >> 2) * 256 + j * 4 + (kk & 3)] = v0;          // zero beyond column n (the A tile was zero-filled)
            Pp[(kk >> 2) * 256 + (j + 1) * 4 + (kk & 3)] = v1;
            double* dst = A + (size_t)(k0 + kk) * lda + j0 + j;
            if (k0 + kk >= n_rows) continue;
            if (j0 + j + 1 < n && vec_ok) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            else {
                if (j0 + j < n) dst[0] = v0;
                if (j0 + j + 1 < n) dst[1] = v1;
            }
        }
}

// ------------------------------------------------------------------------------------------------------------
// triangular solves with one right-hand side: one launch per sweep, the blocks pipelined across the SMs
// ------------------------------------------------------------------------------------------------------------
// U^T y = b (forward) or U x = y (backward) over the 128-row blocks.  Round 1 ran a block step as two stream-ordered launches
// (the 128 x 128 product with the stored inverse block, then the update of the rest of the vector): 4 n / 128 dependent launches,
// 5.2 ms at n = 16384 for 2 GB of reads.  Here block j belongs to ONE CTA, which keeps the running right-hand side of its block
// in shared memory, applies the contributions of the earlier blocks in the order they are finished (the 128 x 128 tile of U is
// requested BEFORE the CTA waits for that block's solution, so on the dependent path the tile is already in registers), solves
// with the inverse block (staged in shared memory at the start), writes its 128 entries of the solution in place and raises a
// flag.  No grid-wide barrier: a step of the dependent chain is flag -> 128 values from L2 -> two 128 x 128 products.
// CTAs take their positions in the sweep from a ticket, so a CTA only ever waits for CTAs that started before it.
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_SMEM = (WBLK + 2 * NB + 8 * NB) * 8 + 16;

template <bool BACKWARD>
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
sweep_kernel(const double* __restrict__ U, long long lda, long long n, const double* __restrict__ Wall, double* b,
             int* flags, int* ticket, int nblk)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Ws = reinterpret_cast<double*>(smem_raw);         // inverse block of the block being solved, global layout (wg_index)
    double* zs = Ws + WBLK;                                   // solution of the source block
    double* bj = zs + NB;                                     // running right-hand side of this CTA's block
    double* part = bj + NB;                                   // up to 8 x NB partial sums
    int* vid_s = reinterpret_cast<int*>(part + 8 * NB);
    const int t = threadIdx.x, lo = t & (NB - 1), q4 = t >> 7;
    const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(U) & 15) == 0);      // rows start on 16-byte boundaries

    for (;;) {
        // Sweep positions are handed out by a ticket, one per round: whoever holds position s only waits for positions < s, and
        // those were all taken by CTAs that are running or done - no CTA ever waits for one that has not started, whatever
        // share of the device this launch gets.
        if (t == 0) *vid_s = atomicAdd(ticket, 1);
        __syncthreads();
        const int s = *vid_s;
        if (s >= nblk) break;
        const int j = BACKWARD ? nblk - 1 - s : s;
        const long long j0 = (long long)j * NB;
        const int nbj = (int)((n - j0) < NB ? (n - j0) : NB);
        __syncthreads();                                      // previous round: Ws / bj no longer read
        {
            const double2* src = reinterpret_cast<const double2*>(Wall + (size_t)j * WBLK);
            double2* dst = reinterpret_cast<double2*>(Ws);
#pragma unroll 4
            for (int i = t; i < WBLK / 2; i += SWEEP_THREADS) dst[i] = src[i];
        }
        if (t < NB) bj[t] = (t < nbj) ? __ldcg(b + j0 + t) : 0.0;

        for (int i = 0; i < s; ++i) {
            const int k = BACKWARD ? nblk - 1 - i : i;        // source block, finished (or being finished) by an earlier CTA
            const long long k0 = (long long)k * NB;
            const int nbk = (int)((n - k0) < NB ? (n - k0) : NB);
            // The tile of U that couples block k to block j, requested before the wait.  16-byte loads when the matrix allows
            // (ncu on the 8-byte version: 57 stall cycles per issue on the load-store queue, lg_throttle).
            double u[32];
            if (!BACKWARD) {                                  // rows of block k (a full block: k < j), columns of block j
                if (vec) {                                    // thread: columns 2 c2, 2 c2 + 1; rows 16 r8 .. 16 r8 + 15
                    const int c2 = t & 63, r8 = t >> 6;
                    const double* src = U + (size_t)(k0 + 16 * r8) * lda + j0 + 2 * c2;
                    const bool ok0 = 2 * c2 < nbj, ok1 = 2 * c2 + 1 < nbj;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        double2 x = make_double2(0.0, 0.0);
                        if (ok1) x = *reinterpret_cast<const double2*>(src + (size_t)r * lda);
                        else if (ok0) x.x = src[(size_t)r * lda];
                        u[2 * r] = x.x; u[2 * r + 1] = x.y;
                    }
                } else {                                      // thread: column lo, rows 32 q4 .. 32 q4 + 31
                    const double* src = U + (size_t)(k0 + 32 * q4) * lda + j0 + lo;
                    const bool ok = lo < nbj;
#pragma unroll
                    for (int r = 0; r < 32; ++r) u[r] = ok ? src[(size_t)r * lda] : 0.0;
                }
            } else {                                          // row j0 + lo (a full block: j < k), columns k0 + 32 q4 + c
                const double* src = U + (size_t)(j0 + lo) * lda + k0 + 32 * q4;
                const int lim = nbk - 32 * q4;
                if (vec) {
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        double2 x = make_double2(0.0, 0.0);
                        if (c + 1 < lim) x = *reinterpret_cast<const double2*>(src + c);
                        else if (c < lim) x.x = src[c];
                        u[c] = x.x; u[c + 1] = x.y;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) u[c] = (c < lim) ? src[c] : 0.0;
                }
            }
            if (t == 0) {
                int v = 0;
                long long spins = 0;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + k) : "memory");
                } while (v == 0 && ++spins < (1ll << 25));    // bounded (tens of seconds): a lost flag must not hang the device
            }
            __syncthreads();                                  // also: zs / part of the previous step are no longer read
            if (t < NB) zs[t] = (t < nbk) ? __ldcg(b + k0 + t) : 0.0;
            __syncthreads();
            if (!BACKWARD && vec) {                           // two columns x 16 rows per thread: eight partial sums per column
                const int c2 = t & 63, r8 = t >> 6;
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double z = zs[16 * r8 + r];
                    acc0 = fma(u[2 * r], z, acc0);
                    acc1 = fma(u[2 * r + 1], z, acc1);
                }
                part[r8 * NB + 2 * c2] = acc0;
                part[r8 * NB + 2 * c2 + 1] = acc1;
                __syncthreads();
                if (t < NB) {
                    double sum = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) sum += part[q * NB + t];
                    bj[t] -= sum;
                }
            } else {
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int r = 0; r < 32; r += 2) {
                    acc0 = fma(u[r], zs[32 * q4 + r], acc0);
                    acc1 = fma(u[r + 1], zs[32 * q4 + r + 1], acc1);
                }
                part[q4 * NB + lo] = acc0 + acc1;
                __syncthreads();
                if (t < NB) bj[t] -= (part[t] + part[NB + t]) + (part[2 * NB + t] + part[3 * NB + t]);
            }
        }
        __syncthreads();
        // the block itself: forward z[kk] = sum_{q <= kk} B[kk][q] r[q], backward z[q] = sum_{kk >= q} B[kk][q] r[kk]   (B = U_jj^{-T})
        {
            const int row = t >> 2, sub = t & 3;              // four threads per entry
            double acc = 0.0;
            if (!BACKWARD) {
                for (int g = sub; g <= (row >> 2); g += 4) {
                    const double2* w = reinterpret_cast<const double2*>(Ws + g * (NB * 4) + row * 4);
                    const double2 w0 = w[0], w1 = w[1];
                    acc = fma(w0.x, bj[4 * g], acc); acc = fma(w0.y, bj[4 * g + 1], acc);
                    acc = fma(w1.x, bj[4 * g + 2], acc); acc = fma(w1.y, bj[4 * g + 3], acc);
                }
            } else {
                const double* w = Ws + (row >> 2) * (NB * 4) + (row & 3);
                for (int kk = row + sub; kk < NB; kk += 4) acc = fma(w[kk * 4], bj[kk], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (sub == 0 && row < nbj) b[j0 + row] = acc;
        }
        __syncthreads();
        if (t == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(flags + j), "r"(1) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// inverse from the factor (lsk_potri_upper): helpers
// ------------------------------------------------------------------------------------------------------------
// row panel of the factor, columns [col0, n), into the panel-major operand layout: P[c][kk] = U[k0 + kk][col0 + c]
__global__ void __launch_bounds__(256)
repack_kernel(const double* __restrict__ U, long long lda, long long n, int k0, long long col0, double* __restrict__ P,
              int kg_stride, int g_off)
{
    const int t = threadIdx.x, j = t & 63;
    const long long c = col0 + 64ll * blockIdx.x + j;
    double* Pp = P + ((size_t)blockIdx.x * kg_stride + g_off) * 256;
    for (int g = t >> 6; g < 32; g += 4) {
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (c < n && k0 + 4 * g + e < n) ? U[(size_t)(k0 + 4 * g + e) * lda + c] : 0.0;
        double2* dst = reinterpret_cast<double2*>(Pp + g * 256 + j * 4);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
    }
}

__global__ void __launch_bounds__(256)
identity_kernel(double* __restrict__ Y, long long n)
{
    const long long i = blockIdx.x, j = 256ll * blockIdx.y + threadIdx.x;
    if (j < n) Y[i * n + j] = (i == j) ? 1.0 : 0.0;
}

// out[j][i] = out[i][j] for i < j (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
mirror_kernel(double* __restrict__ out, long long ld, long long n)
{
    __shared__ double tile[32][33];
    const long long bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long i = 32 * bi + r, j = 32 * bj + tx;
        tile[r][tx] = (i < n && j < n) ? out[i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long j = 32 * bj + r, i = 32 * bi + tx;        // writes out[j][i], coalesced along i
        if (i < n && j < n && i < j) out[j * ld + i] = tile[tx][r];
    }
}

static inline long long n_blocks(long long n) { return (n + NB - 1) / NB; }
static inline long long panel_doubles(long long n) { return (((n + 127) / 128) * 128 + 128) * (2 * NB); }

}  // namespace chol
}  // namespace lsk

extern "C" long long lsk_potrf_work_doubles(long long n) {
    using namespace lsk::chol;
    if (n < 0) return LSK_ESIZE;
    return n_blocks(n) * WBLK + panel_doubles(n);
}

// Two 128-column panel steps per trailing update, so that the update runs with K = 256 (half the read-modify-write traffic
// and half the per-tile overhead of K = 128):
//   diag(k0); trsm(k0) -> P[:, 0:128];  strip: rows [k0+128, k0+256) x cols [k0+128, n) -= P1^T P1 (K = 128);
//   diag(k0+128); trsm(k0+128) -> P[:, 128:256];  trailing rows/cols [k0+256, n) -= [P1 P2]^T [P1 P2] (K = 256).
extern "C" int lsk_potrf_upper(void* A_, long long n, long long lda, void* work_, void* info_dev, void* stream) {
    using namespace lsk;
    using namespace lsk::chol;
    if (!A_ || !work_ || !info_dev) return LSK_EARG;
    if (n < 0 || lda < n || n > 0x7fffffffLL) return LSK_ESIZE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* A = static_cast<double*>(A_);
    double* W = static_cast<double*>(work_);
    double* P = W + n_blocks(n) * WBLK;
    int* info = static_cast<int*>(info_dev);
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    e = cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(diag_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAGT_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(W, 0, (size_t)n_blocks(n) * WBLK * sizeof(double), st);      // diag_tile_kernel writes the non-zero half of each block only
    if (e != cudaSuccess) return (int)e;
    auto diag = [&](int at, int nb, double* w) { diag_tile_kernel<<<1, DIAG_THREADS, DIAGT_SMEM, st>>>(A, lda, at, nb, w, info); };

    constexpr int KG2 = 2 * NB / 4;                      // k-groups per row of the two-panel operand
    LskEpilogue ep = {};
    ep.kind = LSK_EPI_SYRK; ep.nforms = 1; ep.nvars = 1; ep.nterms = 1;
    ep.out_layout = LSK_OUT_ROW_MAJOR;
    ep.scale = -1.0;
    ep.group_begin[0] = 0;

    for (long long k0 = 0, blk = 0; k0 < n; k0 += 2 * NB, blk += 2) {
        // ---- first panel
        const int nb1 = (int)((n - k0) < NB ? (n - k0) : NB);
        diag((int)k0, nb1, W + blk * WBLK);
        const long long s1 = k0 + NB, m1 = n - s1;
        if (m1 <= 0) break;
        trsm_kernel<<<(unsigned)(((m1 + 127) / 128) * 2), TRSM_THREADS, TRSM_SMEM, st>>>(A, lda, n, s1, n, (int)k0, W + blk * WBLK, P, KG2, 0);
        // ---- strip of the next 128 rows (general accumulate, K = 128)
        const long long strip = m1 < NB ? m1 : NB;
        ep.group_end[0] = NB / 4;
        ep.reserved = 0;
        int rc = lsk_pairwise_poly(P, P, strip, m1, KG2, &ep, A + s1 * lda + s1, lda, stream);
        if (rc != 0) return rc;
        // ---- second panel
        diag((int)s1, (int)strip, W + (blk + 1) * WBLK);
        const long long s2 = s1 + NB, m2 = n - s2;
        if (m2 <= 0) break;
        double* P2 = P + 2ll * KG2 * 256;                // operand rows are relative to s2: two 64-row panels further
        trsm_kernel<<<(unsigned)(((m2 + 127) / 128) * 2), TRSM_THREADS, TRSM_SMEM, st>>>(A, lda, n, s2, n, (int)s1, W + (blk + 1) * WBLK, P2, KG2, NB / 4);
        // ---- trailing update (upper triangle of tiles, K = 256)
        ep.group_end[0] = KG2;
        ep.reserved = 2;
        rc = lsk_pairwise_poly(P2, P2, m2, m2, KG2, &ep, A + s2 * lda + s2, lda, stream);
        if (rc != 0) return rc;
    }
    return (int)cudaGetLastError();
}

// which: 1 = forward only (U^T y = b), 2 = backward only (U x = b), 3 = both (U^T U x = b).  `b` is overwritten with the
// solution; `tmp` is a scratch vector of n + 128 doubles.
extern "C" int lsk_potrs_upper(const void* U_, long long n, long long lda, const void* work_, void* b_, void* tmp_, int which,
                               void* stream) {
    using namespace lsk;
    using namespace lsk::chol;
    if (!U_ || !work_ || !b_ || !tmp_) return LSK_EARG;
    if (n < 0 || lda < n || n > 0x7fffffffLL) return LSK_ESIZE;
    if (which < 1 || which > 3) return LSK_EARG;
    if (n == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* U = static_cast<const double*>(U_);
    const double* W = static_cast<const double*>(work_);
    double* b = static_cast<double*>(b_);
    const long long nblk = n_blocks(n);
    // flags (one per block and sweep) and the two tickets live in the caller's scratch vector
    int* flags = static_cast<int*>(tmp_);
    const size_t words = (size_t)(2 * nblk + 2);
    if (words * sizeof(int) > (size_t)(n + 128) * sizeof(double)) return LSK_ESIZE;
    cudaError_t e = cudaMemsetAsync(flags, 0, words * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    static const int cap = [] { const char* v = getenv("LSK_SWEEP_CTAS"); return v ? atoi(v) : 0; }();   // tests: fewer CTAs than blocks
    if (cap > 0 && cap < sms) sms = cap;
    const unsigned grid = (unsigned)(nblk < sms ? nblk : sms);
    if (which & 1) {
        if ((e = cudaFuncSetAttribute(sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM)) != cudaSuccess) return (int)e;
        sweep_kernel<false><<<grid, SWEEP_THREADS, SWEEP_SMEM, st>>>(U, lda, n, W, b, flags, flags + 2 * nblk, (int)nblk);
    }
    if (which & 2) {
        if ((e = cudaFuncSetAttribute(sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM)) != cudaSuccess) return (int)e;
        sweep_kernel<true><<<grid, SWEEP_THREADS, SWEEP_SMEM, st>>>(U, lda, n, W, b, flags + nblk, flags + 2 * nblk + 1, (int)nblk);
    }
    return (int)cudaGetLastError();
}

// Inverse from the factor: out = (U^T U)^{-1}, full symmetric n x n.  Right-looking over the 128-row blocks of the factor:
//   Y_k,: = W_kk^T Yb_k,:                  (trsm_kernel; Yb starts as the identity and ends as U^{-T}, lower triangular)
//   Yb_i,: -= U_k,i^T Y_k,:   for i > k    (tile kernel, read-modify-write, K = 128)
//   out[0:n_k, 0:n_k] += Y_k,:^T Y_k,:     (tile kernel, triangular schedule: A^{-1} = U^{-1} U^{-T} = sum_k Y_k^T Y_k)
// 2 n^3 / 3 flop, all of it on the tensor-core tile kernel; the zero structure of U^{-T} is used exactly (block level).
extern "C" long long lsk_potri_work_doubles(long long n) {
    using namespace lsk::chol;
    if (n < 0) return LSK_ESIZE;
    return ((n * n + 15) / 16) * 16 + 2 * ((((n + 127) / 128) * 128) + 128) * (2 * NB);
}

// Two block rows of Y per pair of updates (K = 256, like the factorisation):
//   Y1 = W_k^T Yb_k;   strip: Yb_{k+1} -= U_k[:, block k+1]^T Y1 (K = 128);   Y2 = W_{k+1}^T Yb_{k+1};
//   Yb_i -= [U_k; U_{k+1}][:, i]^T [Y1; Y2]  for the rows behind;   out[0:n2, 0:n2] += [Y1; Y2]^T [Y1; Y2]  (triangular schedule)
extern "C" int lsk_potri_upper(const void* U_, long long n, long long lda, const void* potrf_work, void* work_, void* out_,
                               long long ld_out, void* stream) {
    using namespace lsk;
    using namespace lsk::chol;
    if (!U_ || !potrf_work || !work_ || !out_) return LSK_EARG;
    if (n < 0 || lda < n || ld_out < n || n > 0x7fffffffLL) return LSK_ESIZE;
    if (n == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* U = static_cast<const double*>(U_);
    const double* W = static_cast<const double*>(potrf_work);
    double* Yb = static_cast<double*>(work_);
    constexpr int KG2 = 2 * NB / 4;
    const long long panel_rows = (((n + 127) / 128) * 128) + 128;
    double* PU = Yb + ((n * n + 15) / 16) * 16;          // 128-byte aligned operands (bulk copies)
    double* PY = PU + panel_rows * (2 * NB);
    double* out = static_cast<double*>(out_);
    cudaError_t e = cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset2DAsync(out, (size_t)ld_out * 8, 0, (size_t)n * 8, (size_t)n, st);
    if (e != cudaSuccess) return (int)e;
    identity_kernel<<<dim3((unsigned)n, (unsigned)((n + 255) / 256)), 256, 0, st>>>(Yb, n);

    LskEpilogue ep = {};
    ep.kind = LSK_EPI_SYRK; ep.nforms = 1; ep.nvars = 1; ep.nterms = 1;
    ep.out_layout = LSK_OUT_ROW_MAJOR;
    ep.group_begin[0] = 0;
    const long long nblk = n_blocks(n);
    auto panels = [](long long cols) { return (unsigned)(((cols + 127) / 128) * 2); };
    for (long long blk = 0; blk < nblk; blk += 2) {
        const long long k0 = blk * NB;
        const long long n1 = (k0 + NB) < n ? (k0 + NB) : n;       // columns the first block row reaches
        const bool two = n1 < n;
        const long long n2 = two ? ((n1 + NB) < n ? (n1 + NB) : n) : n1;
        int rc;
        // first block row; its panel is zero-filled up to column n2 so that it can sit next to the second one
        trsm_kernel<<<panels(n2), TRSM_THREADS, TRSM_SMEM, st>>>(Yb, n, n, 0, n1, (int)k0, W + blk * WBLK, PY, KG2, 0);
        if (two) {
            const long long m1 = n - n1;
            repack_kernel<<<panels(m1), 256, 0, st>>>(U, lda, n, (int)k0, n1, PU, KG2, 0);
            ep.group_end[0] = NB / 4; ep.scale = -1.0; ep.reserved = 0;
            rc = lsk_pairwise_poly(PU, PY, n2 - n1, n1, KG2, &ep, Yb + n1 * n, n, stream);             // strip, K = 128
            if (rc != 0) return rc;
            trsm_kernel<<<panels(n2), TRSM_THREADS, TRSM_SMEM, st>>>(Yb, n, n, 0, n2, (int)n1, W + (blk + 1) * WBLK, PY, KG2, NB / 4);
            const long long m2 = n - n2;
            if (m2 > 0) {
                double* PU2 = PU + 2ll * KG2 * 256;               // operand rows relative to n2 = n1 + 128: two panels further
                repack_kernel<<<panels(m2), 256, 0, st>>>(U, lda, n, (int)n1, n2, PU2, KG2, NB / 4);
                ep.group_end[0] = KG2; ep.scale = -1.0; ep.reserved = 0;
                rc = lsk_pairwise_poly(PU2, PY, m2, n2, KG2, &ep, Yb + n2 * n, n, stream);             // rows behind, K = 256
                if (rc != 0) return rc;
            }
        }
        ep.group_end[0] = two ? KG2 : NB / 4; ep.scale = 1.0; ep.reserved = 2;
        rc = lsk_pairwise_poly(PY, PY, n2, n2, KG2, &ep, out, ld_out, stream);                          // A^{-1} += Y^T Y
        if (rc != 0) return rc;
    }
    mirror_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32)), 256, 0, st>>>(out, ld_out, n);
    return (int)cudaGetLastError();
}

// Forward substitution with many right-hand sides: B [n, nrhs] (row-major, ldb) <- U^{-T} B = L^{-1} B, the whitening solve of
// the GP posterior covariance (K** - V^T V with V = L^{-1} K*).  Same block sweep as the first phase of lsk_potri_upper:
//   B_k,: = W_kk^T B_k,:   (trsm_kernel)      B_i,: -= U_k,i^T B_k,:  for i > k   (tile kernel, read-modify-write)
extern "C" long long lsk_trsm_work_doubles(long long n, long long nrhs) {
    using namespace lsk::chol;
    if (n < 0 || nrhs < 0) return LSK_ESIZE;
    return ((((n + 127) / 128) * 128) + (((nrhs + 127) / 128) * 128)) * NB;
}

extern "C" int lsk_trsm_lower(const void* U_, long long n, long long lda, const void* potrf_work, void* B_, long long nrhs,
                              long long ldb, void* work_, void* stream) {
    using namespace lsk;
    using namespace lsk::chol;
    if (!U_ || !potrf_work || !B_ || !work_) return LSK_EARG;
    if (n < 0 || nrhs < 0 || lda < n || ldb < nrhs || n > 0x7fffffffLL || nrhs > 0x7fffffffLL) return LSK_ESIZE;
    if (n == 0 || nrhs == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double* U = static_cast<const double*>(U_);
    const double* W = static_cast<const double*>(potrf_work);
    double* B = static_cast<double*>(B_);
    double* PU = static_cast<double*>(work_);
    double* PY = PU + (((n + 127) / 128) * 128) * NB;
    cudaError_t e = cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM);
    if (e != cudaSuccess) return (int)e;
    LskEpilogue ep = {};
    ep.kind = LSK_EPI_SYRK; ep.nforms = 1; ep.nvars = 1; ep.nterms = 1;
    ep.out_layout = LSK_OUT_ROW_MAJOR;
    ep.group_begin[0] = 0; ep.group_end[0] = NB / 4;
    ep.scale = -1.0; ep.reserved = 0;
    const long long nblk = n_blocks(n);
    for (long long blk = 0; blk < nblk; ++blk) {
        const long long k0 = blk * NB;
        const long long nk = (k0 + NB) < n ? (k0 + NB) : n;
        const long long m = n - nk;
        trsm_kernel<<<(unsigned)(((nrhs + 127) / 128) * 2), TRSM_THREADS, TRSM_SMEM, st>>>(B, ldb, n, 0, nrhs, (int)k0, W + blk * WBLK, PY, NB / 4, 0);
        if (m > 0) {
            repack_kernel<<<(unsigned)(((m + 127) / 128) * 2), 256, 0, st>>>(U, lda, n, (int)k0, nk, PU, NB / 4, 0);
            const int rc = lsk_pairwise_poly(PU, PY, m, nrhs, NB / 4, &ep, B + nk * ldb, ldb, stream);
            if (rc != 0) return rc;
        }
    }
    return (int)cudaGetLastError();
}
