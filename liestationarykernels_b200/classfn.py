"""Host-side plans for class functions on compact Lie groups: which per-point embeddings the CUDA kernels need,
how the bilinear forms map to polynomial variables, and the integer matrix that turns spectral weights into
the coefficient table of the fused epilogue (`struct LskEpilogue`, include/lsk.h).

Everything here is exact integer/rational arithmetic done once per (group, n, order); per call only
`coef = M^T (a_lambda * d_lambda)` (a few hundred flops) runs on the host.

Reference behaviour covered: `torus_representative` + `chi` + the weighted sum of `EigenbasisSumKernel.forward`
(`spaces/so.py:136-168,288-293`, `spaces/su.py:91-92,175-179`, `spectral_kernel.py:45-54`).
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache

import numpy as np

from . import characters as ch
from . import invariants as inv
from ._lib import (EPI_CHEB1, EPI_HORNER, EPI_ORBIT2, EPI_SPARSE, EPI_STAIR2, MAX_COEF, MAX_FORMS, MAX_ROWS, MAX_VARS,
                   LskEpilogue)


HODGE = 4       # embedding part code: Lambda^k x with the Hodge star applied to the row index (csrc/lsk_embed.cu)
Z_SIGN = {4: -1.0, 8: 1.0}   # z = Z_SIGN * <*Lambda^r x, Lambda^r y>, matching the reference's class labelling
SU4_REAL = 5    # embedding part code: Lambda^2 of an SU(4) element in the real basis of Lambda^2 C^4 (its SO(6) image)
SPIN5 = 100     # embedding segment code: the 16 rotor coefficients of the Spin(5) lift (csrc/lsk_embed.cu)


def _binom(n, k):
    return math.comb(n, k)


class GroupPlan:
    """Embedding layout + polynomial tables for one (family, n, order)."""

    def __init__(self, family: str, n: int, order: int, use_spin: bool = True):
        self.family, self.n, self.order = family, n, order
        self.use_spin = use_spin
        self.square_mask = 0
        kept, dropped = ch.truncated_irreps(family, n, order) if order else ([], [])
        self.irreps = kept
        self.dropped = dropped
        self.is_complex = family == "SU"
        self._build_forms()
        if order:
            self._build_polynomials()

    # ------------------------------------------------------------------ embeddings / bilinear forms
    def _build_forms(self):
        n = self.n
        segs_a, segs_b = [], []        # (wedge, part)
        if self.family == "SO":
            rank = n // 2
            if rank > 4:
                raise NotImplementedError("SO(%d): compound matrices beyond Lambda^4 are not implemented "
                                          "(supported: SO(3) ... SO(8))" % n)
            if n % 2 == 0:
                # e_1..e_r from the compound matrices; for even r additionally z = prod 2 sin(theta_i), the invariant that
                # separates the two SO(2r) classes with equal spectrum, as the Hodge-star form <*Lambda^r x, Lambda^r y>
                for k in range(1, rank + 1):
                    segs_a.append((k, 0))
                    segs_b.append((k, 0))
                if rank % 2 == 0:
                    segs_a.append((rank, HODGE))
                    segs_b.append((rank, 0))
            elif n == 5 and self.use_spin:
                # Spin(5) lift: tr g (25-term dot) + the rotor inner product <R_x, R_y> = cos(t1/2) cos(t2/2)
                # (16-term dot) replace the 100-term Lambda^2 form: K = 44 instead of 128
                segs_a += [(1, 0), (SPIN5, 0)]
                segs_b += [(1, 0), (SPIN5, 0)]
            else:
                for k in range(1, rank + 1):
                    segs_a.append((k, 0))
                    segs_b.append((k, 0))
            self.real_vars = rank
        else:
            if n > 6:
                raise NotImplementedError("SU(%d): needs more than %d real invariants (supported: SU(2) ... SU(6))" % (n, MAX_VARS))
            # e_k for k <= n/2; e_{n-k} = conj(e_k); e_{n/2} is real for even n
            half = n // 2
            for k in range(1, half + 1):
                real_only = (n % 2 == 0 and k == half)
                if n == 4 and k == 1:
                    # tr(x y^dagger) = sum (a + ib)(c - id) with three real forms instead of four (Karatsuba):
                    #   f0 = sum a c,  f1 = sum b d,  f2 = sum (a + b)(c - d);   Re = f0 + f1,  Im = f2 - f0 + f1
                    segs_a += [(1, 6), (1, 7), (1, 8)]
                    segs_b += [(1, 6), (1, 7), (1, 9)]
                    continue
                if n == 4 and k == 2:
                    # Lambda^2 C^4 is a real representation (SU(4) -> SO(6)): both sides carry the real 6 x 6 image and the
                    # form costs 36 multiply-adds instead of 72 (csrc/lsk_embed.cu, part 5)
                    segs_a.append((2, SU4_REAL))
                    segs_b.append((2, SU4_REAL))
                    continue
                segs_a.append((k, 1))
                segs_b.append((k, 2))
                if not real_only:
                    segs_a.append((k, 1))
                    segs_b.append((k, 3))
            self.real_vars = len(segs_a)
        assert len(segs_a) <= MAX_FORMS
        self.segments_a, self.segments_b = segs_a, segs_b
        begin, g = [], 0
        for (k, part) in segs_a:
            begin.append(g)
            length = 16 if k == SPIN5 else 36 if part == SU4_REAL else _binom(n, k) ** 2 * (2 if self.is_complex and part < 6 else 1)
            g += (length + 3) // 4
        self.group_begin = begin
        self.kg_total = g
        self.group_end = begin[1:] + [g]
        self.nforms = len(segs_a)

    # ------------------------------------------------------------------ class polynomials
    def _build_polynomials(self):
        sigs = [r[0] for r in self.irreps]
        n = self.n
        if self.family == "SO":
            rank = n // 2
            if rank == 1:
                self.kind = "cheb1"
                # chi_l = sum_mu m(mu) orb_mu,  orb_mu = 2 T_mu(c) (mu > 0), 1 (mu = 0);  c = (tr g - 1) / 2
                deg = max(s[0] for s in sigs)
                M = np.zeros((len(sigs), deg + 1))
                for l, s in enumerate(sigs):
                    for mu, m in ch.dominant_weight_multiplicities("B", tuple(s)).items():
                        M[l, mu[0]] += m * (2 if mu[0] else 1)
                self.M = M
                self.aff = [[-0.5, 0.5]]
                return
            if n % 2 == 0:
                self.kind = "sparse"
                self._build_so_even(sigs)
                return
            monos, mat = inv.character_matrix("SO", n, sigs)
            self.monos = monos
            self.M = np.array(mat, dtype=np.float64)
            assert np.all(np.abs(self.M) < 2.0 ** 53)
            self.aff = _so_odd_affine(n)
            if n == 5 and self.use_spin:
                # forms: u0 = tr g, u1 = <R_x, R_y> (sign-ambiguous, enters squared):
                #   s1 = (u0 - 1) / 2,   s2 = (4 u1)^2 / 4 - 1 - s1 = -1/2 - u0 / 2 + 4 u1^2
                self.aff = [[-0.5, 0.5, 0.0], [-0.5, -0.5, 4.0]]
                self.square_mask = 0b10
            if rank == 2:
                self.kind = "rank2"
                self._build_stair(monos)
                self._build_orbit(sigs)
            else:
                self.kind = "sparse"
                self.sparse_expo = monos
            return
        # SU(n)
        if n == 2:
            self.kind = "cheb1"
            deg = max(s[0] for s in sigs)
            M = np.zeros((len(sigs), deg + 1))
            for l, s in enumerate(sigs):
                for w, m in ch.character_weights("SU", 2, s).items():
                    M[l, abs(w[0] - w[1])] += m          # e^{i(w0-w1)theta}; the +-pair sums to 2 T_k(cos theta)
            self.M = M
            self.aff = [[0.0, 0.5]]                        # cos(theta) = Re tr(g) / 2
            return
        self.kind = "sparse"
        monos, mat = _su_real_character_matrix(n, sigs)
        self.sparse_expo = monos
        self.M = np.array(mat, dtype=np.float64)
        assert np.all(np.abs(self.M) < 2.0 ** 53)
        self.aff = [[0.0] + [1.0 if f == v else 0.0 for f in range(self.nforms)] for v in range(self.nforms)]
        if n == 4:          # forms (f0, f1, f2, e2) -> variables (Re e1, Im e1, e2), see _build_forms
            self.aff = [[0.0, 1.0, 1.0, 0.0, 0.0], [0.0, -1.0, 1.0, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0, 1.0]]

    def _build_so_even(self, sigs):
        """Re chi = ( P(s) + (-1)^{r/2} z Q(s) ) / 2 for even r, P(s) / 2 for odd r (invariants.so_even_character_polys);
        variables (s_1..s_r[, z]).  The sign convention of z (which of the two classes with equal spectrum is which) is
        the reference's `torus_representative` (so.py:151-168); Z_SIGN was fixed against its golden output."""
        n, r = self.n, self.n // 2
        with_z = r % 2 == 0
        polys = []
        for sgn in sigs:
            P, Q = inv.so_even_character_polys(n, tuple(sgn))
            poly = {e + ((0,) if with_z else ()): Fraction(c, 2) for e, c in P.items()}
            if with_z:
                for e, c in Q.items():
                    poly[e + (1,)] = poly.get(e + (1,), 0) + Fraction((-1) ** (r // 2) * c, 2)
            polys.append(poly)
        monos = sorted(set().union(*[set(q) for q in polys]))
        index = {m: i for i, m in enumerate(monos)}
        M = np.zeros((len(sigs), len(monos)))
        for l, q in enumerate(polys):
            for m, c in q.items():
                M[l, index[m]] = float(c)
        self.sparse_expo, self.M = monos, M
        aff = _so_even_affine(n)                       # rows over (1, e_1..e_r)
        nf = self.nforms
        self.aff = [row + [0.0] * (1 + nf - len(row)) for row in aff]
        if with_z:
            self.aff.append([0.0] * nf + [Z_SIGN[n]])

    def _build_stair(self, monos):
        """Layout of the fast (monomial) form for csrc/lsk_pairwise.cu `epi_stair2`: row slot k (evaluation order,
        highest power of s2 first) holds sum_a C[b][a] s1^a, highest a first, padded with zeros at the high end to an
        even number of coefficients, at the fixed offset 16 k; the pair counts go into a 64-bit shape word (4 bits per
        slot; bit 48 + slot flags a row whose first coefficient is that padding)."""
        rows = {}
        for (a, b) in monos:
            rows[b] = max(rows.get(b, -1), a)
        nrows = max(rows) + 1
        self.stair_rows = [max(2, (rows.get(b, 0) + 2) // 2 * 2) for b in range(nrows)]
        self.stair_ok = nrows <= 12 and max(self.stair_rows) <= 16      # LSK_STAIR_MAX_ROWS / LSK_STAIR_STRIDE
        index = {m: i for i, m in enumerate(monos)}
        self.stair_slots = []            # (offset in coef[], monomial index or -1)
        self.stair_shape = 0
        for slot, b in enumerate(range(nrows - 1, -1, -1)):
            ln = self.stair_rows[b]
            if self.stair_ok:
                self.stair_shape |= (ln // 2) << (4 * slot)
                if (rows.get(b, 0) + 1) % 2:            # odd row: the slot's first coefficient is padding (bit 48 + slot);
                    self.stair_shape |= 1 << (48 + slot)   # only the shape-specialised kernel instantiations read this
            for pos, a in enumerate(range(ln - 1, -1, -1)):
                self.stair_slots.append((16 * slot + pos, index.get((a, b), -1)))

    def _build_orbit(self, sigs):
        # W[mu] = sum_l w_l m_l(mu);  full symmetric Chebyshev-product matrix (see lsk_pairwise.cu epi_orbit2)
        dom = sorted(set().union(*[set(ch.dominant_weight_multiplicities("B", tuple(s))) for s in sigs]))
        self.orbit_weights = dom
        D = max(mu[0] for mu in dom) + 1
        self.orbit_D = D
        Mo = np.zeros((len(sigs), len(dom)))
        for l, s in enumerate(sigs):
            for mu, m in ch.dominant_weight_multiplicities("B", tuple(s)).items():
                Mo[l, dom.index(mu)] = m
        self.M_orbit = Mo
        fac = []
        for (a, b) in dom:
            fac.append(1.0 if a == 0 else (2.0 if b == 0 else 4.0))
        self.orbit_fac = np.array(fac)

    # ------------------------------------------------------------------ epilogue for given irrep weights
    def _base_epilogue(self) -> LskEpilogue:
        ep = LskEpilogue()
        ep.nforms = self.nforms
        for f in range(self.nforms):
            ep.group_begin[f] = self.group_begin[f]
            ep.group_end[f] = self.group_end[f]
        ep.nvars = len(self.aff)
        ep.reserved = self.square_mask << 8
        for v, row in enumerate(self.aff):
            for i, val in enumerate(row):
                ep.aff[v][i] = float(val)
        ep.scale = 1.0
        return ep

    def epilogue(self, weights: np.ndarray, scale: float, form: str = "auto") -> LskEpilogue:
        """weights[l] multiplies chi_l (typically a_l * d_l).  form in {'auto', 'fast', 'robust'} (rank 2 only)."""
        w = np.asarray(weights, dtype=np.float64)
        ep = self._base_epilogue()
        ep.scale = float(scale)
        if self.kind == "cheb1":
            coef = _dot_long(w, self.M)
            if len(coef) > MAX_COEF:
                raise ValueError("order too large for the rank-1 epilogue (%d > %d)" % (len(coef), MAX_COEF))
            ep.kind, ep.nterms = EPI_CHEB1, len(coef)
            for i, c in enumerate(coef):
                ep.coef[i] = c
            return ep
        if self.kind == "rank2":
            use_fast = form == "fast" or (form == "auto" and self.stair_ok and self.fast_form_is_safe(w))
            if use_fast:
                if not self.stair_ok:
                    raise ValueError("staircase table too large for the fast form")
                coef = _dot_long(w, self.M) * float(scale)     # the scale is folded into the coefficients: one multiply
                ep.scale = 1.0                                  # per entry less (the dedicated SO(5) kernel requires it)
                ep.kind, ep.nterms = EPI_STAIR2, len(self.stair_rows)
                ep.stair_shape = self.stair_shape
                for slot, ln in enumerate(reversed(self.stair_rows)):
                    ep.row_len[slot] = ln
                for off, j in self.stair_slots:
                    ep.coef[off] = coef[j] if j >= 0 else 0.0
                return ep
            D = self.orbit_D
            if D > 20:
                raise ValueError("order too large for the Weyl-orbit epilogue (max exponent %d > 19)" % (D - 1))
            Dp = 8 if D <= 8 else 12 if D <= 12 else 16 if D <= 16 else 20
            W = _dot_long(w, self.M_orbit) * self.orbit_fac
            ep.kind, ep.nterms = EPI_ORBIT2, D
            full = np.zeros((Dp, Dp))
            for (a, b), val in zip(self.orbit_weights, W):
                full[a, b] = val
                full[b, a] = val
            for a in range(D):
                for b in range(Dp):
                    ep.coef[a * Dp + b] = full[a, b]
            return ep
        # sparse polynomial in nvars variables: nested-Horner tape (csrc/lsk_pairwise.cu `epi_tape`)
        coef = _dot_long(w, self.M)
        # every monomial of the plan is kept (zero coefficients included): the tape's structure then depends on group and
        # order only, which is what lets lsk_pairwise_poly pick a generated straight-line epilogue by structure hash
        terms = {tuple(e): float(c) for c, e in zip(coef, self.sparse_expo)}
        tape = horner_tape(terms)
        if 2 * len(tape) > MAX_COEF:
            raise ValueError("polynomial too large for the Horner-tape epilogue (%d steps > %d)" % (len(tape), MAX_COEF // 2))
        ep.kind, ep.nterms = EPI_HORNER, len(tape)
        _write_tape(ep, 0, tape)
        return ep

    # ------------------------------------------------------------------ per-irrep feature epilogues
    def feature_epilogues(self, row_scales, stride: int, monomial_list: bool = False):
        """Yields (epilogue, first output column) for batches of irreps: Phi[:, l*stride + p] = row_scales[l] chi_l.
        Rows of rank >= 2 are nested-Horner tapes (FEAT_HORNER, the pairwise kernel) or, with `monomial_list`, plain
        (coefficient, exponents) lists (FEAT_SPARSE, what the homogeneous-space kernel averages over subgroup samples)."""
        from ._lib import EPI_FEAT_CHEB1, EPI_FEAT_HORNER, EPI_FEAT_SPARSE
        scales = np.asarray(row_scales, dtype=np.float64)
        L = len(self.irreps)
        rows = []
        for l in range(L):
            if self.kind == "cheb1":
                nz = np.nonzero(self.M[l])[0]
                deg = int(nz.max()) if len(nz) else 0
                rows.append([float(scales[l] * self.M[l, k]) for k in range(deg + 1)])
            else:
                expo = self.monos if self.kind == "rank2" else self.sparse_expo
                terms = {tuple(expo[t]): float(scales[l] * self.M[l, t]) for t in np.nonzero(self.M[l])[0]}
                terms = terms if terms else {tuple(expo[0]): 0.0}
                rows.append([(c, _pack_exponents(e)) for e, c in terms.items()] if monomial_list else horner_tape(terms))
        cap = MAX_COEF if self.kind == "cheb1" else MAX_COEF // 2
        start = 0
        while start < L:
            stop, used = start, 0
            while stop < L and stop - start < MAX_ROWS and used + len(rows[stop]) <= cap:
                used += len(rows[stop])
                stop += 1
            if stop == start:
                raise ValueError("irrep %d has too many polynomial terms for the feature epilogue" % start)
            ep = self._base_epilogue()
            ep.kind = EPI_FEAT_CHEB1 if self.kind == "cheb1" else EPI_FEAT_SPARSE if monomial_list else EPI_FEAT_HORNER
            ep.nterms = stop - start
            ep.feat_stride = stride
            pos = 0
            for i, l in enumerate(range(start, stop)):
                ep.row_len[i] = len(rows[l])
                if self.kind == "cheb1":
                    for item in rows[l]:
                        ep.coef[pos] = item
                        pos += 1
                else:
                    _write_tape(ep, pos, rows[l])
                    pos += len(rows[l])
            yield ep, start * stride
            start = stop

    def collapsed_row_epilogue(self, weights, scale: float):
        """One output row = sum_l weights[l] chi_l, in the per-row encoding of the FEAT_* epilogues (used by the
        homogeneous-space pair kernel, where the polynomial is evaluated once per subgroup sample)."""
        from ._lib import EPI_FEAT_CHEB1, EPI_FEAT_SPARSE
        coef = _dot_long(np.asarray(weights, dtype=np.float64), self.M)
        ep = self._base_epilogue()
        ep.scale = float(scale)
        ep.nterms = 1
        ep.feat_stride = 0
        if self.kind == "cheb1":
            if len(coef) > MAX_COEF:
                raise ValueError("order too large")
            ep.kind = EPI_FEAT_CHEB1
            ep.row_len[0] = len(coef)
            for i, c in enumerate(coef):
                ep.coef[i] = c
            return ep
        expo = self.monos if self.kind == "rank2" else self.sparse_expo
        keep = [(c, e) for c, e in zip(coef, expo) if c != 0.0]
        if 2 * len(keep) > MAX_COEF:
            raise ValueError("too many polynomial terms (%d) for the homogeneous-space kernel" % len(keep))
        ep.kind = EPI_FEAT_SPARSE
        ep.row_len[0] = len(keep)
        for t, (c, e) in enumerate(keep):
            packed = 0
            for j, ej in enumerate(e):
                packed |= int(ej) << (8 * j)
            ep.coef[2 * t] = c
            ep.coef[2 * t + 1] = np.array([packed], dtype=np.uint64).view(np.float64)[0]
        return ep

    # ------------------------------------------------------------------ fast-form safety check (rank 2)
    def _probe(self):
        if getattr(self, "_probe_cache", None) is not None:
            return self._probe_cache
        rng = np.random.default_rng(12345)
        th = rng.uniform(0, math.pi, size=(2, 4096))
        th[:, :64] *= 0.05                                   # near the identity class, where the kernel peaks
        c1, c2 = np.cos(th[0]), np.cos(th[1])
        s1, s2 = c1 + c2, c1 * c2
        D = self.orbit_D
        T1 = np.cos(np.arange(D)[:, None] * th[0][None, :])
        T2 = np.cos(np.arange(D)[:, None] * th[1][None, :])
        orb = np.stack([(T1[a] * T2[b] + T1[b] * T2[a]) * (0.5 if a == b else 1.0) for (a, b) in self.orbit_weights])
        self._probe_cache = (s1, s2, orb)
        return self._probe_cache

    def fast_form_is_safe(self, w: np.ndarray, tol: float = 2e-13) -> bool:
        """Evaluates the monomial (fast) and Weyl-orbit (robust) forms in float64 on 4096 probe classes and accepts
        the fast one only when they agree to `tol` relative to the local kernel value — the monomial basis loses
        digits for slowly decaying spectra (Matern-1/2, l=0.4 gives 5e-11)."""
        s1, s2, orb = self._probe()
        robust = (_dot_long(w, self.M_orbit) * self.orbit_fac) @ orb
        coef = _dot_long(w, self.M)
        fast = np.zeros_like(s1)
        table = {off: (coef[j] if j >= 0 else 0.0) for off, j in self.stair_slots}
        for slot, ln in enumerate(reversed(self.stair_rows)):
            inner = np.full_like(s1, table[16 * slot])
            for a in range(1, ln):
                inner = inner * s1 + table[16 * slot + a]
            fast = fast * s2 + inner
        denom = np.maximum(np.abs(robust), 1e-300)
        return bool(np.max(np.abs(fast - robust) / denom) < tol)


def horner_tape(terms: dict):
    """Nested-Horner evaluation order of a sparse polynomial {exponent tuple: coefficient} as a list of (c, op) steps
    for `epi_tape` (csrc/lsk_pairwise.cu): variable 0 is the innermost level; one step per node of the exponent trie,
    missing exponents inside a chain cost one multiplication (inner level: a zero coefficient)."""
    nv = len(next(iter(terms)))

    def rec(sub, level):
        if level == 0:
            top = max(e[0] for e in sub)
            return [(sub.get((e,), 0.0), 1 if e == top else 0) for e in range(top, -1, -1)]
        groups = {}
        for ex, c in sub.items():
            groups.setdefault(ex[-1], {})[ex[:-1]] = c
        top, base, out = max(groups), 2 + 3 * (level - 1), []
        for e in range(top, -1, -1):
            if e in groups:
                out += rec(groups[e], level - 1)
                out.append((0.0, base + (1 if e == top else 0)))
            else:
                out.append((0.0, base + 2))
        return out

    return rec(terms, nv - 1)


def _pack_exponents(e) -> int:
    packed = 0
    for j, ej in enumerate(e):
        assert 0 <= ej < 256
        packed |= int(ej) << (8 * j)
    return packed


def _write_tape(ep: LskEpilogue, pos: int, tape):
    for t, (c, op) in enumerate(tape):
        ep.coef[2 * (pos + t)] = c
        ep.coef[2 * (pos + t) + 1] = np.array([op], dtype=np.uint64).view(np.float64)[0]


def _dot_long(w: np.ndarray, M: np.ndarray) -> np.ndarray:
    """w @ M accumulated in extended precision: the integer entries of M reach 1e7 and alternate in sign."""
    return np.asarray((w.astype(np.longdouble)[None, :] @ M.astype(np.longdouble))[0], dtype=np.float64)


def _so_odd_affine(n: int):
    """Rows [const, coef of e_1, ..., coef of e_r]: s_k(cos theta) as an affine function of e_1..e_r of g in SO(n).
    det(zI - g) = (z - 1) prod_i (z^2 - 2 c_i z + 1) = sum_k (-1)^k e_k z^{n-k}."""
    r = n // 2

    def pmul(a, b):
        out = [Fraction(0)] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                out[i + j] += x * y
        return out

    # polys in z, low -> high;  basis_j(z) = (z - 1) (-2z)^j (z^2 + 1)^(r - j), multiplied by s_j (s_0 = 1)
    basis = []
    for j in range(r + 1):
        p = [Fraction(1)]
        for _ in range(j):
            p = pmul(p, [Fraction(0), Fraction(-2)])
        for _ in range(r - j):
            p = pmul(p, [Fraction(1), Fraction(0), Fraction(1)])
        basis.append(pmul(p, [Fraction(-1), Fraction(1)]))
    # e_k = (-1)^k * coefficient of z^(n-k):   e_k = T[k][0] + sum_j T[k][j] s_j
    T = [[(-1) ** k * (basis[j][n - k] if n - k < len(basis[j]) else Fraction(0)) for j in range(r + 1)] for k in range(r + 1)]
    # solve for s_1..s_r from e_1..e_r (triangular: e_k involves s_j for j <= k)
    rows = []
    sol = {0: ([Fraction(1)] + [Fraction(0)] * r)}            # s_0 = 1 as affine function of (1, e_1..e_r)
    for k in range(1, r + 1):
        # e_k = sum_{j<=k} T[k][j] s_j  ->  s_k = (e_k - sum_{j<k} T[k][j] s_j) / T[k][k]
        acc = [Fraction(0)] * (r + 1)
        acc[k] = Fraction(1)
        for j in range(k):
            for i in range(r + 1):
                acc[i] -= T[k][j] * sol[j][i]
        assert T[k][k] != 0 and all(T[k][j] == 0 for j in range(k + 1, r + 1))
        sol[k] = [a / T[k][k] for a in acc]
        rows.append([float(v) for v in sol[k]])
    return rows


def _so_even_affine(n: int):
    """Rows [const, coef of e_1, ..., coef of e_r]: s_k(cos theta) as an affine function of e_1..e_r of g in SO(2r).
    det(zI - g) = prod_i (z^2 - 2 c_i z + 1) = sum_j s_j (-2z)^j (z^2 + 1)^(r - j)."""
    r = n // 2

    def pmul(a, b):
        out = [Fraction(0)] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                out[i + j] += x * y
        return out

    basis = []
    for j in range(r + 1):
        p = [Fraction(1)]
        for _ in range(j):
            p = pmul(p, [Fraction(0), Fraction(-2)])
        for _ in range(r - j):
            p = pmul(p, [Fraction(1), Fraction(0), Fraction(1)])
        basis.append(p)
    T = [[(-1) ** k * (basis[j][n - k] if n - k < len(basis[j]) else Fraction(0)) for j in range(r + 1)] for k in range(r + 1)]
    sol = {0: [Fraction(1)] + [Fraction(0)] * r}
    rows = []
    for k in range(1, r + 1):
        acc = [Fraction(0)] * (r + 1)
        acc[k] = Fraction(1)
        for j in range(k):
            for i in range(r + 1):
                acc[i] -= T[k][j] * sol[j][i]
        assert T[k][k] != 0
        sol[k] = [a / T[k][k] for a in acc]
        rows.append([float(v) for v in sol[k]])
    return rows


def _su_real_character_matrix(n: int, sigs):
    """Characters of SU(n) as real polynomials in the real invariants
       (Re e_1, Im e_1, Re e_2, Im e_2, ...) [and e_{n/2} real for even n] - only the real part, which is all the
       kernels use (`spectral_kernel.py:50` takes `.real`)."""
    half = n // 2
    names = []           # variable index of (k, 're'/'im')
    for k in range(1, half + 1):
        names.append((k, "re"))
        if not (n % 2 == 0 and k == half):
            names.append((k, "im"))
    nv = len(names)
    vid = {nm: i for i, nm in enumerate(names)}

    def unit(i):
        return tuple(1 if j == i else 0 for j in range(nv))

    zero = tuple([0] * nv)

    def e_complex(k):
        """e_k as (re_poly, im_poly) in the real variables."""
        kk = k if k <= half else n - k
        re = {unit(vid[(kk, "re")]): 1}
        im = {unit(vid[(kk, "im")]): (1 if k <= half else -1)} if (kk, "im") in vid else {}
        return re, im

    def cmul(a, b):
        re = inv.p_add(inv.p_mul(a[0], b[0]), inv.p_mul(a[1], b[1]), -1)
        im = inv.p_add(inv.p_mul(a[0], b[1]), inv.p_mul(a[1], b[0]))
        return re, im

    pow_cache = {}

    def cpow(k, p):
        if (k, p) not in pow_cache:
            if p == 0:
                pow_cache[(k, p)] = ({zero: 1}, {})
            else:
                pow_cache[(k, p)] = cmul(cpow(k, p - 1), e_complex(k))
        return pow_cache[(k, p)]

    polys = []
    for s in sigs:
        cp = inv.su_character_poly(n, tuple(s))
        real = {}
        for expo, c in cp.items():
            term = ({zero: 1}, {})
            for k, p in enumerate(expo, start=1):
                if p:
                    term = cmul(term, cpow(k, p))
            real = inv.p_add(real, term[0], c)
        polys.append(real)
    monos = sorted(set().union(*[set(p) for p in polys]))
    index = {m: i for i, m in enumerate(monos)}
    mat = [[0] * len(monos) for _ in polys]
    for l, p in enumerate(polys):
        for m, c in p.items():
            mat[l][index[m]] = c
    return monos, mat


@lru_cache(maxsize=None)
def group_plan(family: str, n: int, order: int, use_spin: bool = True) -> GroupPlan:
    return GroupPlan(family, n, order, use_spin)
