// Per-point embeddings: the O(N) prologue of the pairwise kernels.
//
// For g = x y^{-1} in SO(n) / SU(n) the coefficients of the characteristic polynomial are bilinear in the
// compound matrices of x and y (Cauchy-Binet):  e_k(x y^T) = < Lambda^k x , Lambda^k y >_F  (with a conjugate on y
// for SU(n)).  This kernel writes, for every point, the concatenation of the requested compound matrices
// (k = 1: the matrix itself) into the panel-major operand layout of lsk_pairwise.cu:
//     out[panel][k-group][row in panel (64)][4]      (doubles; rows beyond `num` and pad columns are zero)
// It replaces the reference's `inv` + `cartesian_prod` + `bmm` operand preparation (space.py:62-78).
#include "lsk_common.cuh"
#include "lsk_pairwise.cuh"

#define LSK_SEG_SPIN5 100     // segment code: rotor coefficients of the Spin(5) lift (n = 5, real)

namespace lsk {

// Snapshot of a few device scalars into host-mapped pinned memory (lsk_snapshot_doubles, lsk_embed_points2_snapshot): the Python layer
// verifies its cached coefficient block against the live hyper-parameters without draining the stream.  No fence: a system-scope
// fence has to reach every peer the process has mapped.  Instead every value travels with a tag, tag = bits(value) ^ key(seq), in one
// 16-byte store; the host accepts a slot only when the pair is consistent, so a torn or reordered update is simply polled again.
struct Snapshot { const double* src[LSK_MAX_SNAPSHOT]; int n; double* dst; unsigned long long seq; };

__host__ __device__ inline unsigned long long snapshot_key(unsigned long long seq) { return seq * 0x9E3779B97F4A7C15ull + 0x7F4A7C159E3779B9ull; }

__device__ __forceinline__ void snapshot_write(const Snapshot& snap, int t) {
    if (t < snap.n) {
        const double v = *snap.src[t];
        const unsigned long long tag = (unsigned long long)__double_as_longlong(v) ^ snapshot_key(snap.seq);
        asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(snap.dst + 2 * t), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
    }
}

struct EmbedSpec {
    int nseg;
    int wedge[LSK_MAX_FORMS];        // 1, 2 or 3
    int part[LSK_MAX_FORMS];         // 0: real matrix;  1: complex (re, im);  2: complex (re, im) [B side, real form];
                         // 3: complex (-im, re) [B side, imaginary form];  4: real, Hodge star on the row index (n = 2k)
    int group_begin[LSK_MAX_FORMS];  // first k-group of the segment
    int length[LSK_MAX_FORMS];       // number of doubles in the segment (before padding to a multiple of 4)
};

__device__ __forceinline__ int binom(int n, int k) {
    int r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

// lexicographic rank -> increasing index tuple of size k (k <= 4) out of n
__device__ __forceinline__ void unrank(int idx, int n, int k, int (&t)[4]) {
    int start = 0;
    for (int pos = 0; pos < k; ++pos) {
        for (int v = start; v < n; ++v) {
            const int cnt = binom(n - v - 1, k - pos - 1);
            if (idx < cnt) { t[pos] = v; start = v + 1; break; }
            idx -= cnt;
        }
    }
}

template <bool CPLX>
__device__ __forceinline__ void load(const double* x, int n, int r, int c, double& re, double& im) {
    if (CPLX) { re = x[2 * (r * n + c)]; im = x[2 * (r * n + c) + 1]; }
    else { re = x[r * n + c]; im = 0.0; }
}

__device__ __forceinline__ void cmul(double ar, double ai, double br, double bi, double& cr, double& ci) {
    cr = ar * br - ai * bi;
    ci = ar * bi + ai * br;
}

template <bool CPLX>
__device__ void minor(const double* x, int n, int k, const int (&I)[4], const int (&J)[4], double& re, double& im) {
    if (k == 1) { load<CPLX>(x, n, I[0], J[0], re, im); return; }
    if (k == 2) {
        double ar, ai, br, bi, cr, ci, dr, di, p1r, p1i, p2r, p2i;
        load<CPLX>(x, n, I[0], J[0], ar, ai); load<CPLX>(x, n, I[1], J[1], dr, di);
        load<CPLX>(x, n, I[0], J[1], br, bi); load<CPLX>(x, n, I[1], J[0], cr, ci);
        cmul(ar, ai, dr, di, p1r, p1i); cmul(br, bi, cr, ci, p2r, p2i);
        re = p1r - p2r; im = p1i - p2i;
        return;
    }
    if (k == 4) {
        // 4 x 4 minor (SO(8): Lambda^4), real matrices only: Laplace expansion by complementary 2 x 2 minors of the two
        // row pairs (rows 0,1 against rows 2,3)
        double a[4][4];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) { double t; load<CPLX>(x, n, I[r], J[c], a[r][c], t); }
        double det = 0.0;
        for (int c0 = 0; c0 < 4; ++c0)
            for (int c1 = c0 + 1; c1 < 4; ++c1) {
                int d0 = -1, d1 = -1;
                for (int c = 0; c < 4; ++c)
                    if (c != c0 && c != c1) { if (d0 < 0) d0 = c; else d1 = c; }
                const double top = a[0][c0] * a[1][c1] - a[0][c1] * a[1][c0];
                const double bot = a[2][d0] * a[3][d1] - a[2][d1] * a[3][d0];
                const int sgn = ((c0 + c1 + 0 + 1) & 1) ? -1 : 1;       // (-1)^(row indices 0 + 1 + column indices c0 + c1)
                det += sgn * top * bot;
            }
        re = det; im = 0.0;
        return;
    }
    // k == 3: cofactor expansion along the first row
    double m[3][3][2];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) load<CPLX>(x, n, I[a], J[b], m[a][b][0], m[a][b][1]);
    re = 0.0; im = 0.0;
    for (int c = 0; c < 3; ++c) {
        const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;     // cyclic: sign is always +
        double p1r, p1i, p2r, p2i, tr, ti;
        cmul(m[1][c1][0], m[1][c1][1], m[2][c2][0], m[2][c2][1], p1r, p1i);
        cmul(m[1][c2][0], m[1][c2][1], m[2][c1][0], m[2][c1][1], p2r, p2i);
        cmul(m[0][c][0], m[0][c][1], p1r - p2r, p1i - p2i, tr, ti);
        re += tr; im += ti;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Spin(5) lift: the rotor R(x) in the even Clifford algebra Cl+(5) (16 real coefficients, index = blade mask >> 1)
// with R e_a R~ = sum_b x_ba e_b.  x is reduced to the identity by ten Givens rotations, x = G_1^T ... G_10^T, and the
// rotor is the product of the plane rotors cos(phi/2) - sin(phi/2) e_j e_i.  For g = x y^T:
//     < R(x), R(y) > = +- cos(theta_1 / 2) cos(theta_2 / 2)
// (the character of the 4-dimensional spin representation / 4), which together with tr g determines the class of g.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr double blade_sign(int a, int b) {
    int s = 0;
    a >>= 1;
    while (a) {
        int t = a & b, c = 0;
        while (t) { c += t & 1; t >>= 1; }
        s += c;
        a >>= 1;
    }
    return (s & 1) ? -1.0 : 1.0;
}
__host__ __device__ constexpr int even_mask(int k) {          // even-parity 5-bit mask A with A >> 1 == k
    int c = 0, t = k;
    while (t) { c += t & 1; t >>= 1; }
    return (k << 1) | (c & 1);
}

template <int J, int I>
struct RotorTable {
    int dst[16];
    bool neg[16];
    constexpr RotorTable() : dst{}, neg{} {
        const int B = (1 << J) | (1 << I);
        for (int k = 0; k < 16; ++k) {
            const int A = even_mask(k);
            dst[k] = (A ^ B) >> 1;
            neg[k] = blade_sign(A, B) < 0.0;
        }
    }
};

// 1 / sqrt(x) for normal positive x: hardware seed (2^-22) + one third-order Newton step, no special-case paths
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    const double e2 = fma(-x * y, y, 1.0);           // second (linear) step: the seed error cubed is ~2^-60, this
    return fma(0.5 * y, e2, y);                      // step removes what the rounding of the first one left
}

// R <- R * (ch - sh e_J e_I): all blade indices and signs are compile-time constants, everything stays in registers
template <int J, int I>
__device__ __forceinline__ void rotor_step(double (&m)[5][5], double (&R)[16]) {
    // Branch-free: this kernel is one long dependent instruction chain per point (latency-bound, a few warps per SM), so
    // every avoided branch, special-case path and division shortens the launch that precedes each pair kernel.
    const double a = m[J][J], b = m[I][J];
    const double r2 = fma(a, a, b * b);              // entries of an orthogonal matrix: no overflow to guard
    const bool live = r2 > 1e-280;                   // column already aligned (a = b = 0): identity step
    const double ir = rsqrt_pos(live ? r2 : 1.0);
    const double c = live ? a * ir : 1.0, s = live ? b * ir : 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double rj = m[J][k], ri = m[I][k];
        m[J][k] = fma(c, rj, s * ri);
        m[I][k] = fma(c, ri, -s * rj);
    }
    // half-angle without cancellation: h = (1 + |c|) / 2 in [1/2, 1];  c >= 0: (ch, sh) = (sqrt h, s / (2 sqrt h)),
    // c < 0: (|s| / (2 sqrt h), sign(s) sqrt h)
    const double h = 0.5 * (1.0 + fabs(c)), t = rsqrt_pos(h);
    const double big = h * t, small = 0.5 * t;
    const double ch = (c >= 0.0) ? big : fabs(s) * small;
    const double sh = (c >= 0.0) ? s * small : copysign(big, s);
    // target blade (A ^ B) >> 1 receives -sh * sign(A, B) * R[k]; destinations and signs are compile-time tables
    constexpr RotorTable<J, I> tab{};
    double nw[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) nw[k] = ch * R[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) nw[tab.dst[k]] = fma(tab.neg[k] ? sh : -sh, R[k], nw[tab.dst[k]]);
#pragma unroll
    for (int k = 0; k < 16; ++k) R[k] = nw[k];
}

__device__ __forceinline__ void spin5_rotor(const double* __restrict__ x, double (&R)[16]) {
    double m[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) m[r][c] = x[r * 5 + c];
#pragma unroll
    for (int k = 0; k < 16; ++k) R[k] = 0.0;
    R[0] = 1.0;
    rotor_step<0, 1>(m, R); rotor_step<0, 2>(m, R); rotor_step<0, 3>(m, R); rotor_step<0, 4>(m, R);
    rotor_step<1, 2>(m, R); rotor_step<1, 3>(m, R); rotor_step<1, 4>(m, R);
    rotor_step<2, 3>(m, R); rotor_step<2, 4>(m, R);
    rotor_step<3, 4>(m, R);
}

// One operand of an embedding launch.  A launch carries one or two of them (blockIdx.y): kernel(x, y) embeds both
// sides with a single launch, which matters once the row-sharded pair kernel is down to a few hundred microseconds.
struct EmbedJob {
    const double* x;
    long long num;
    double* out;
    EmbedSpec spec;
};

template <bool CPLX>
__global__ void __launch_bounds__(256)
embed_kernel(const __grid_constant__ EmbedJob job0, const __grid_constant__ EmbedJob job1, int n, int kg_total,
             const __grid_constant__ Snapshot snap) {
    pdl_launch_dependents();
    const EmbedJob& job = blockIdx.y ? job1 : job0;
    const double* __restrict__ x = job.x;
    double* __restrict__ out = job.out;
    const long long num = job.num;
    const EmbedSpec& spec = job.spec;
    pdl_wait();                                   // x may come from the kernel launched just before (e.g. space.rand)
    if (blockIdx.x == 0 && blockIdx.y == 0) snapshot_write(snap, threadIdx.x);
    const long long total = ((num + ROW_PAD - 1) / ROW_PAD) * (ROW_PAD / PANEL) * (long long)kg_total * (PANEL * 4);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx & 3);
        const int r = (int)((idx >> 2) & (PANEL - 1));
        const long long gp = idx / (PANEL * 4);
        const int g = (int)(gp % kg_total);
        const long long row = (gp / kg_total) * PANEL + r;
        double val = 0.0;
        if (row < num) {
            int seg = 0;
            for (int s = 1; s < spec.nseg; ++s) if (g >= spec.group_begin[s]) seg = s;
            const int e = (g - spec.group_begin[seg]) * 4 + j;
            if (!CPLX && spec.wedge[seg] == LSK_SEG_SPIN5) {
                continue;                 // written by spin5_embed_kernel (one thread per point)
            } else if (CPLX && spec.part[seg] == 5) {
                // SU(4): Lambda^2 C^4 = C^6 carries a real structure (Hodge star composed with conjugation) that commutes
                // with SU(4); in the basis  u_{2p} = (e_P + s_p e_P')/sqrt2,  u_{2p+1} = i (e_P - s_p e_P')/sqrt2  (P' the
                // complementary index pair, s = +,-,+ for P = 01, 02, 03) Lambda^2 g is a REAL orthogonal 6 x 6 matrix (the
                // double cover SU(4) -> SO(6)), and e_2(x y^{-1}) = <R(x), R(y)> costs 36 real multiply-adds instead of the
                // 72 of the complex form.  Entry (a, b) = Re sum conj(B[I][a]) Lambda^2x[I][J] B[J][b] (2 x 2 minors).
                if (e < 36) {
                    const int a = e / 6, b = e - 6 * a, pa = a >> 1, pb = b >> 1;
                    const double sa = (pa == 1) ? -1.0 : 1.0, sb = (pb == 1) ? -1.0 : 1.0;
                    double acc = 0.0;
                    for (int u = 0; u < 2; ++u)
                        for (int v = 0; v < 2; ++v) {
                            int I[4] = {0, 0, 0, 0}, J[4] = {0, 0, 0, 0};
                            unrank(u ? 5 - pa : pa, 4, 2, I);
                            unrank(v ? 5 - pb : pb, 4, 2, J);
                            double re, im;
                            minor<CPLX>(x + row * 2ll * 16, 4, 2, I, J, re, im);
                            // conj(beta_a) beta_b with beta = (1 | s) for even columns, (i | -i s) for odd ones
                            double car = (a & 1) ? 0.0 : (u ? sa : 1.0), cai = (a & 1) ? (u ? sa : -1.0) : 0.0;   // conj(beta_a)
                            double cbr = (b & 1) ? 0.0 : (v ? sb : 1.0), cbi = (b & 1) ? (v ? -sb : 1.0) : 0.0;   // beta_b
                            const double wr = car * cbr - cai * cbi, wi = car * cbi + cai * cbr;
                            acc += wr * re - wi * im;
                        }
                    val = 0.5 * acc;
                }
            } else if (e < spec.length[seg]) {
                const int k = spec.wedge[seg], part = spec.part[seg];
                const int C = binom(n, k);
                const bool one_real = CPLX && part >= 6;          // parts 6..9: one real number per complex entry
                const int el = (CPLX && !one_real) ? (e >> 1) : e;
                int I[4] = {0, 0, 0, 0}, J[4] = {0, 0, 0, 0};
                unrank(el / C, n, k, I);
                unrank(el % C, n, k, J);
                double hodge = 1.0;
                if (!CPLX && part == 4) {
                    // Hodge star on the row index (n = 2k): I -> complement I^c, sign of the permutation (I, I^c)
                    int comp[4] = {0, 0, 0, 0}, nc = 0, inv = 0;
                    for (int v = 0; v < n; ++v) {
                        bool in = false;
                        for (int q = 0; q < k; ++q) in = in || (I[q] == v);
                        if (!in) {
                            for (int q = 0; q < k; ++q) inv += (I[q] > v);      // inversions between I and I^c
                            if (nc < 4) comp[nc] = v;
                            ++nc;
                        }
                    }
                    hodge = (inv & 1) ? -1.0 : 1.0;
                    for (int q = 0; q < 4; ++q) I[q] = comp[q];
                }
                double re, im;
                minor<CPLX>(x + row * (long long)(CPLX ? 2 : 1) * n * n, n, k, I, J, re, im);
                if (!CPLX) val = hodge * re;
                else if (one_real) val = part == 6 ? re : part == 7 ? im : part == 8 ? re + im : re - im;
                else if (part == 3) val = (e & 1) ? re : -im;
                else val = (e & 1) ? im : re;
            }
        }
        out[idx] = val;
    }
}

// one thread per point: rotor of the Spin(5) lift into k-groups [g0, g0 + 4) of the panel-major buffer; with
// WITH_MATRIX the same thread also writes the matrix itself (25 values + 3 zeros) into k-groups [gm, gm + 7), which
// makes this kernel the whole SO(5) embedding (one launch per operand)
template <bool WITH_MATRIX>
__global__ void __launch_bounds__(64)
spin5_embed_kernel(const double* __restrict__ x0, long long num0, double* __restrict__ out0,
                   const double* __restrict__ x1, long long num1, double* __restrict__ out1, int kg_total, int g0, int gm,
                   const __grid_constant__ Snapshot snap) {
    pdl_launch_dependents();
    const double* __restrict__ x = blockIdx.y ? x1 : x0;
    double* __restrict__ out = blockIdx.y ? out1 : out0;
    const long long num = blockIdx.y ? num1 : num0;
    pdl_wait();
    if (blockIdx.x == 0 && blockIdx.y == 0) snapshot_write(snap, threadIdx.x);
    const long long rows_padded = ((num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_padded) return;
    double R[16];
    if (row < num) spin5_rotor(x + row * 25, R);
    else for (int k = 0; k < 16; ++k) R[k] = 0.0;
    if (WITH_MATRIX) {
        double v[28];
#pragma unroll
        for (int k = 0; k < 28; ++k) v[k] = (k < 25 && row < num) ? x[row * 25 + k] : 0.0;
        double* mb = out + ((row >> 6) * kg_total + gm) * (PANEL * 4) + (row & 63) * 4;
#pragma unroll
        for (int g = 0; g < 7; ++g) {
            double2* dst = reinterpret_cast<double2*>(mb + (size_t)g * (PANEL * 4));
            dst[0] = make_double2(v[4 * g], v[4 * g + 1]);
            dst[1] = make_double2(v[4 * g + 2], v[4 * g + 3]);
        }
    }
    double* base = out + ((row >> 6) * kg_total + g0) * (PANEL * 4) + (row & 63) * 4;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        double2* dst = reinterpret_cast<double2*>(base + (size_t)g * (PANEL * 4));
        dst[0] = make_double2(R[4 * g], R[4 * g + 1]);
        dst[1] = make_double2(R[4 * g + 2], R[4 * g + 3]);
    }
}


// one thread per point: the whole standard SU(4) embedding [Re x | Im x | Re x +- Im x | R(x)] (21 k-groups; the generic kernel
// above spends one thread per OUTPUT ELEMENT, each re-deriving index pairs and four 2 x 2 minors: 58 us for 2 x 16384 points, 3 %
// of a C3 step).  R(x) is the real 6 x 6 image of Lambda^2 x described there; with m[P][Q] the 36 complex 2 x 2 minors,
// P' = 5 - P the complementary pair and s = (+, -, +) for P = 01, 02, 03:
//   R[2p  ][2q  ] = ( Re m[p][q] + sq Re m[p][q'] + sp Re m[p'][q] + sp sq Re m[p'][q']) / 2
//   R[2p  ][2q+1] = (-Im m[p][q] + sq Im m[p][q'] - sp Im m[p'][q] + sp sq Im m[p'][q']) / 2
//   R[2p+1][2q  ] = ( Im m[p][q] + sq Im m[p][q'] - sp Im m[p'][q] - sp sq Im m[p'][q']) / 2
//   R[2p+1][2q+1] = ( Re m[p][q] - sq Re m[p][q'] - sp Re m[p'][q] + sp sq Re m[p'][q']) / 2
// (same association of the four terms as the generic kernel: results are bit-identical).
__global__ void __launch_bounds__(64)
su4_embed_kernel(const double* __restrict__ x0, long long num0, double* __restrict__ out0, int third0,
                 const double* __restrict__ x1, long long num1, double* __restrict__ out1, int third1,
                 const __grid_constant__ Snapshot snap) {
    pdl_launch_dependents();
    const double* __restrict__ x = blockIdx.y ? x1 : x0;
    double* __restrict__ out = blockIdx.y ? out1 : out0;
    const long long num = blockIdx.y ? num1 : num0;
    const int third = blockIdx.y ? third1 : third0;           // 8: Re + Im (side A), 9: Re - Im (side B)
    pdl_wait();
    if (blockIdx.x == 0 && blockIdx.y == 0) snapshot_write(snap, threadIdx.x);
    constexpr int KGT = 21;
    const long long rows_padded = ((num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_padded) return;
    double* base = out + ((row >> 6) * KGT) * (PANEL * 4) + (row & 63) * 4;
    auto put4 = [&](int g, double a, double b, double c, double d) {
        double2* dst = reinterpret_cast<double2*>(base + (size_t)g * (PANEL * 4));
        dst[0] = make_double2(a, b);
        dst[1] = make_double2(c, d);
    };
    if (row >= num) {
#pragma unroll
        for (int g = 0; g < KGT; ++g) put4(g, 0.0, 0.0, 0.0, 0.0);
        return;
    }
    double re[4][4], im[4][4];
    const double2* xp = reinterpret_cast<const double2*>(x + row * 32);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { const double2 v = xp[r * 4 + c]; re[r][c] = v.x; im[r][c] = v.y; }
    const double sg = third == 8 ? 1.0 : -1.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        put4(r, re[r][0], re[r][1], re[r][2], re[r][3]);
        put4(4 + r, im[r][0], im[r][1], im[r][2], im[r][3]);
        put4(8 + r, re[r][0] + sg * im[r][0], re[r][1] + sg * im[r][1], re[r][2] + sg * im[r][2], re[r][3] + sg * im[r][3]);
    }
    // 2 x 2 minors, index pairs in lexicographic order
    constexpr int P0[6] = {0, 0, 0, 1, 1, 2}, P1[6] = {1, 2, 3, 2, 3, 3};
    double mr[6][6], mi[6][6];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int i0 = P0[p], i1 = P1[p], j0 = P0[q], j1 = P1[q];
            // (a d - b c) with a = x[i0][j0], d = x[i1][j1], b = x[i0][j1], c = x[i1][j0]; products as in `cmul`
            const double p1r = re[i0][j0] * re[i1][j1] - im[i0][j0] * im[i1][j1], p1i = re[i0][j0] * im[i1][j1] + im[i0][j0] * re[i1][j1];
            const double p2r = re[i0][j1] * re[i1][j0] - im[i0][j1] * im[i1][j0], p2i = re[i0][j1] * im[i1][j0] + im[i0][j1] * re[i1][j0];
            mr[p][q] = p1r - p2r;
            mi[p][q] = p1i - p2i;
        }
    double R[36];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double sp = p == 1 ? -1.0 : 1.0, sq = q == 1 ? -1.0 : 1.0;
            const int pc = 5 - p, qc = 5 - q;
            // the generic kernel accumulates acc += w_r * re - w_i * im over (u, v) = (0,0), (0,1), (1,0), (1,1) starting from 0
            R[(2 * p) * 6 + 2 * q] = 0.5 * (((0.0 + mr[p][q]) + sq * mr[p][qc]) + sp * mr[pc][q] + (sp * sq) * mr[pc][qc]);
            R[(2 * p) * 6 + 2 * q + 1] = 0.5 * (((0.0 - mi[p][q]) + sq * mi[p][qc]) - sp * mi[pc][q] + (sp * sq) * mi[pc][qc]);
            R[(2 * p + 1) * 6 + 2 * q] = 0.5 * (((0.0 + mi[p][q]) + sq * mi[p][qc]) - sp * mi[pc][q] - (sp * sq) * mi[pc][qc]);
            R[(2 * p + 1) * 6 + 2 * q + 1] = 0.5 * (((0.0 + mr[p][q]) - sq * mr[p][qc]) - sp * mr[pc][q] + (sp * sq) * mr[pc][qc]);
        }
#pragma unroll
    for (int g = 0; g < 9; ++g) put4(12 + g, R[4 * g], R[4 * g + 1], R[4 * g + 2], R[4 * g + 3]);
}

}  // namespace lsk

namespace lsk {

static int make_spec(long long num, int n, int is_complex, int nseg, const int* wedge, const int* part, const int* group_begin,
                     int kg_total, EmbedSpec& spec) {
    if (!wedge || !part || !group_begin) return LSK_EARG;
    if (num < 0 || n < 1 || n > 8 || nseg < 1 || nseg > LSK_MAX_FORMS || kg_total < 1) return LSK_ESIZE;
    spec.nseg = nseg;
    for (int s = 0; s < LSK_MAX_FORMS; ++s) { spec.wedge[s] = 1; spec.part[s] = 0; spec.group_begin[s] = 0; spec.length[s] = 0; }
    for (int s = 0; s < nseg; ++s) {
        if (wedge[s] == LSK_SEG_SPIN5) {
            if (is_complex || n != 5 || part[s] != 0) return LSK_EARG;
            spec.wedge[s] = wedge[s]; spec.part[s] = 0; spec.group_begin[s] = group_begin[s]; spec.length[s] = 16;
            if (group_begin[s] < 0 || group_begin[s] + 4 > kg_total) return LSK_ESIZE;
            if (s > 0 && group_begin[s] < spec.group_begin[s - 1] + (spec.length[s - 1] + 3) / 4) return LSK_EARG;
            continue;
        }
        if (wedge[s] < 1 || wedge[s] > 4 || wedge[s] > n || (wedge[s] == 4 && is_complex)) return LSK_EARG;
        const bool su4_real = is_complex && part[s] == 5 && n == 4 && wedge[s] == 2;
        const bool one_real = is_complex && part[s] >= 6 && part[s] <= 9;
        if (!su4_real && !one_real && (is_complex ? (part[s] < 1 || part[s] > 3) : (part[s] != 0 && !(part[s] == 4 && n == 2 * wedge[s])))) return LSK_EARG;
        int c = 1;
        for (int i = 1; i <= wedge[s]; ++i) c = c * (n - wedge[s] + i) / i;
        spec.wedge[s] = wedge[s]; spec.part[s] = part[s]; spec.group_begin[s] = group_begin[s];
        spec.length[s] = su4_real ? 36 : one_real ? c * c : c * c * (is_complex ? 2 : 1);
        const int end = group_begin[s] + (spec.length[s] + 3) / 4;
        if (group_begin[s] < 0 || end > kg_total) return LSK_ESIZE;
        if (s > 0 && group_begin[s] < spec.group_begin[s - 1] + (spec.length[s - 1] + 3) / 4) return LSK_EARG;
    }
    return 0;
}

static bool is_so5_standard(const EmbedSpec& spec, int n, int is_complex, int kg_total) {
    return !is_complex && n == 5 && spec.nseg == 2 && spec.wedge[0] == 1 && spec.wedge[1] == LSK_SEG_SPIN5 &&
           spec.group_begin[0] == 0 && spec.group_begin[1] == 7 && kg_total == 11;
}

static bool is_su4_standard(const EmbedSpec& spec, int n, int is_complex, int kg_total) {
    return is_complex && n == 4 && spec.nseg == 4 && kg_total == 21 &&
           spec.wedge[0] == 1 && spec.wedge[1] == 1 && spec.wedge[2] == 1 && spec.wedge[3] == 2 &&
           spec.part[0] == 6 && spec.part[1] == 7 && (spec.part[2] == 8 || spec.part[2] == 9) && spec.part[3] == 5 &&
           spec.group_begin[0] == 0 && spec.group_begin[1] == 4 && spec.group_begin[2] == 8 && spec.group_begin[3] == 12;
}

// embeds one or two operands (njobs) that share n, is_complex and kg_total with ONE launch where the layout allows it
static int embed_jobs(EmbedJob* jobs, int njobs, int n, int is_complex, int kg_total, cudaStream_t st, const Snapshot& snap = Snapshot{}) {
    long long rows_max = 0, total_max = 0;
    for (int j = 0; j < njobs; ++j) {
        const long long rp = ((jobs[j].num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
        rows_max = rp > rows_max ? rp : rows_max;
        const long long total = rp / PANEL * (long long)kg_total * (PANEL * 4);
        total_max = total > total_max ? total : total_max;
    }
    if (rows_max == 0) return 0;
    const EmbedJob& j0 = jobs[0];
    const EmbedJob& j1 = jobs[njobs - 1];
    bool all_so5 = true;
    for (int j = 0; j < njobs; ++j) all_so5 = all_so5 && is_so5_standard(jobs[j].spec, n, is_complex, kg_total);
    if (all_so5) {
        // the standard SO(5) layout [matrix | rotor]: one thread per point writes both segments
        cudaError_t e = launch_pdl(spin5_embed_kernel<true>, dim3((unsigned)((rows_max + 63) / 64), njobs), dim3(64), 0, st,
                                   j0.x, j0.num, j0.out, j1.x, j1.num, j1.out, kg_total, 7, 0, snap);
        return (int)e;
    }
    bool all_su4 = true;
    for (int j = 0; j < njobs; ++j) all_su4 = all_su4 && is_su4_standard(jobs[j].spec, n, is_complex, kg_total);
    if (all_su4) {
        // the standard SU(4) layout [Re | Im | Re +- Im | SO(6) image of Lambda^2]: one thread per point writes all 21 k-groups
        cudaError_t e = launch_pdl(su4_embed_kernel, dim3((unsigned)((rows_max + 63) / 64), njobs), dim3(64), 0, st,
                                   j0.x, j0.num, j0.out, j0.spec.part[2], j1.x, j1.num, j1.out, j1.spec.part[2], snap);
        return (int)e;
    }
    long long blocks = (total_max + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    cudaError_t e;
    if (is_complex) e = launch_pdl(embed_kernel<true>, dim3((unsigned)blocks, njobs), dim3(256), 0, st, j0, j1, n, kg_total, snap);
    else e = launch_pdl(embed_kernel<false>, dim3((unsigned)blocks, njobs), dim3(256), 0, st, j0, j1, n, kg_total, snap);
    if (e != cudaSuccess) return (int)e;
    for (int j = 0; j < njobs; ++j)
        for (int s = 0; s < jobs[j].spec.nseg; ++s)
            if (jobs[j].spec.wedge[s] == LSK_SEG_SPIN5 && jobs[j].num > 0) {
                const long long rp = ((jobs[j].num + ROW_PAD - 1) / ROW_PAD) * ROW_PAD;
                e = launch_pdl(spin5_embed_kernel<false>, dim3((unsigned)((rp + 63) / 64), 1), dim3(64), 0, st,
                               jobs[j].x, jobs[j].num, jobs[j].out, jobs[j].x, jobs[j].num, jobs[j].out,
                               kg_total, jobs[j].spec.group_begin[s], 0, Snapshot{});
                if (e != cudaSuccess) return (int)e;
            }
    return (int)cudaGetLastError();
}

}  // namespace lsk

extern "C" int lsk_embed_points(const void* x, long long num, int n, int is_complex, int nseg,
                                const int* wedge, const int* part, const int* group_begin, int kg_total,
                                void* out, void* stream) {
    using namespace lsk;
    if (!x || !out) return LSK_EARG;
    EmbedJob job;
    const int rc = make_spec(num, n, is_complex, nseg, wedge, part, group_begin, kg_total, job.spec);
    if (rc) return rc;
    if (num == 0) return 0;
    job.x = static_cast<const double*>(x); job.num = num; job.out = static_cast<double*>(out);
    return embed_jobs(&job, 1, n, is_complex, kg_total, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int lsk_embed_points2(const void* x, long long num_x, const int* part_x, void* out_x,
                                 const void* y, long long num_y, const int* part_y, void* out_y,
                                 int n, int is_complex, int nseg, const int* wedge, const int* group_begin, int kg_total,
                                 void* stream) {
    using namespace lsk;
    if (!x || !y || !out_x || !out_y) return LSK_EARG;
    EmbedJob jobs[2];
    int rc = make_spec(num_x, n, is_complex, nseg, wedge, part_x, group_begin, kg_total, jobs[0].spec);
    if (rc) return rc;
    rc = make_spec(num_y, n, is_complex, nseg, wedge, part_y, group_begin, kg_total, jobs[1].spec);
    if (rc) return rc;
    jobs[0].x = static_cast<const double*>(x); jobs[0].num = num_x; jobs[0].out = static_cast<double*>(out_x);
    jobs[1].x = static_cast<const double*>(y); jobs[1].num = num_y; jobs[1].out = static_cast<double*>(out_y);
    return embed_jobs(jobs, 2, n, is_complex, kg_total, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int lsk_version(void) { return LSK_VERSION; }

// Stateless: the text for a return code of any entry point (0 ok, < 0 LSK_E*, > 0 cudaError_t).
// ---- snapshot of a few device scalars into host-mapped pinned memory: stand-alone launch ----------------------------------
namespace lsk {
__global__ void __launch_bounds__(32)
snapshot_kernel(const __grid_constant__ Snapshot snap) {
    pdl_launch_dependents();
    pdl_wait();                                   // the scalars may have been written by the kernel launched just before
    snapshot_write(snap, threadIdx.x);
}
}  // namespace lsk

static int make_snapshot(const void* const* src, int n, void* dst_host_mapped, unsigned long long seq, lsk::Snapshot& snap) {
    snap = lsk::Snapshot{};
    if (n == 0) return 0;
    if (!src || !dst_host_mapped) return LSK_EARG;
    if (n < 0 || n > LSK_MAX_SNAPSHOT) return LSK_ESIZE;
    for (int i = 0; i < n; ++i) { if (!src[i]) return LSK_EARG; snap.src[i] = static_cast<const double*>(src[i]); }
    snap.n = n; snap.dst = static_cast<double*>(dst_host_mapped); snap.seq = seq;
    return 0;
}

extern "C" int lsk_snapshot_doubles(const void* const* src, int n, void* dst_host_mapped, unsigned long long seq, void* stream) {
    using namespace lsk;
    if (!src || !dst_host_mapped) return LSK_EARG;
    Snapshot snap;
    const int rc = make_snapshot(src, n, dst_host_mapped, seq, snap);
    if (rc) return rc;
    return (int)launch_pdl(snapshot_kernel, dim3(1), dim3(32), 0, reinterpret_cast<cudaStream_t>(stream), snap);
}

// lsk_embed_points2 with the snapshot written by the embedding kernel itself (no extra launch in front of the pair kernel)
extern "C" int lsk_embed_points2_snapshot(const void* x, long long num_x, const int* part_x, void* out_x,
                                          const void* y, long long num_y, const int* part_y, void* out_y,
                                          int n, int is_complex, int nseg, const int* wedge, const int* group_begin, int kg_total,
                                          const void* const* snap_src, int n_snap, void* snap_dst_host_mapped, unsigned long long seq,
                                          void* stream) {
    using namespace lsk;
    if (!x || !y || !out_x || !out_y) return LSK_EARG;
    if (num_x <= 0 || num_y <= 0) return LSK_ESIZE;              // the snapshot rides on a launch: both operands must be non-empty
    EmbedJob jobs[2];
    int rc = make_spec(num_x, n, is_complex, nseg, wedge, part_x, group_begin, kg_total, jobs[0].spec);
    if (rc) return rc;
    rc = make_spec(num_y, n, is_complex, nseg, wedge, part_y, group_begin, kg_total, jobs[1].spec);
    if (rc) return rc;
    Snapshot snap;
    rc = make_snapshot(snap_src, n_snap, snap_dst_host_mapped, seq, snap);
    if (rc) return rc;
    jobs[0].x = static_cast<const double*>(x); jobs[0].num = num_x; jobs[0].out = static_cast<double*>(out_x);
    jobs[1].x = static_cast<const double*>(y); jobs[1].num = num_y; jobs[1].out = static_cast<double*>(out_y);
    return embed_jobs(jobs, 2, n, is_complex, kg_total, reinterpret_cast<cudaStream_t>(stream), snap);
}

extern "C" const char* lsk_error_string(int code) {
    switch (code) {
    case 0: return "ok";
    case LSK_EARG: return "invalid argument";
    case LSK_ESIZE: return "size out of supported range";
    case LSK_EUNSUPPORTED: return "configuration not compiled";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error code";
    }
}

extern "C" long long lsk_embed_buffer_doubles(long long num, int kg_total) {
    return ((num + lsk::ROW_PAD - 1) / lsk::ROW_PAD) * (lsk::ROW_PAD / lsk::PANEL) * (long long)kg_total * (lsk::PANEL * 4);
}
