"""Characters as polynomials in conjugation invariants ("collapse" step).

A character is a class function, so instead of eigen-decomposing every pairwise difference
(`spaces/so.py:136-168`, `spaces/su.py:91-92`) and looping over monomials (`so.py:288-293`, `su.py:175-179`)
the CUDA kernels evaluate

    SO(2r+1):  chi_lambda(g) = P_lambda(s_1..s_r),   s_k = e_k(cos theta_1 .. cos theta_r)
    SO(2r)  :  chi_lambda(g) = P_lambda(s_1..s_r) + pf * Q_lambda(s_1..s_r),  pf = prod_i sin(theta_i)  (sign-carrying)
    SU(n)   :  chi_lambda(g) = P_lambda(e_1..e_{n-1}),  e_k = elementary symmetric functions of the eigenvalues

with integer coefficient matrices computed here in exact arithmetic from the weight multiplicities of
`characters.py`.  Only the (hyper-parameter dependent) weights a_lambda * d_lambda are floating point.

All polynomials are dicts {exponent tuple: int}.
"""
from __future__ import annotations

import itertools
from fractions import Fraction
from functools import lru_cache

from . import characters as ch


# --------------------------------------------------------------------------------------------------------------
# tiny exact sparse polynomial algebra
# --------------------------------------------------------------------------------------------------------------
def p_add(a: dict, b: dict, scale=1) -> dict:
    out = dict(a)
    for k, v in b.items():
        nv = out.get(k, 0) + scale * v
        if nv == 0:
            out.pop(k, None)
        else:
            out[k] = nv
    return out


def p_mul(a: dict, b: dict) -> dict:
    out = {}
    for ka, va in a.items():
        for kb, vb in b.items():
            k = tuple(x + y for x, y in zip(ka, kb))
            nv = out.get(k, 0) + va * vb
            if nv == 0:
                out.pop(k, None)
            else:
                out[k] = nv
    return out


@lru_cache(maxsize=None)
def cheb_T(k: int) -> tuple:
    """Coefficients (low -> high) of the Chebyshev polynomial T_k, cos(k t) = T_k(cos t)."""
    if k == 0:
        return (1,)
    if k == 1:
        return (0, 1)
    a, b = cheb_T(k - 1), cheb_T(k - 2)
    out = [0] * (k + 1)
    for i, v in enumerate(a):
        out[i + 1] += 2 * v
    for i, v in enumerate(b):
        out[i] -= v
    return tuple(out)


@lru_cache(maxsize=None)
def _elem_sym(r: int, k: int) -> dict:
    """e_k(c_1..c_r) as a polynomial in c."""
    out = {}
    for idx in itertools.combinations(range(r), k):
        out[tuple(1 if i in idx else 0 for i in range(r))] = 1
    return out


_POW_CACHE = {}


def _elem_pow(r: int, k: int, p: int) -> dict:
    key = (r, k, p)
    if key not in _POW_CACHE:
        if p == 0:
            _POW_CACHE[key] = {tuple([0] * r): 1}
        else:
            _POW_CACHE[key] = p_mul(_elem_pow(r, k, p - 1), _elem_sym(r, k))
    return _POW_CACHE[key]


def symmetric_to_elementary(poly: dict, r: int) -> dict:
    """Rewrite a symmetric polynomial in c_1..c_r as a polynomial in e_1..e_r (exponent tuple over e_k).
    Classic leading-term elimination; exact for integer/Fraction coefficients."""
    work = dict(poly)
    out = {}
    while work:
        lead = max(work)                       # lexicographically largest exponent; sorted descending by symmetry
        coef = work[lead]
        assert all(lead[i] >= lead[i + 1] for i in range(r - 1)), "polynomial is not symmetric"
        expo = tuple(lead[i] - (lead[i + 1] if i + 1 < r else 0) for i in range(r))
        out[expo] = out.get(expo, 0) + coef
        term = {tuple([0] * r): 1}
        for k in range(r):
            if expo[k]:
                term = p_mul(term, _elem_pow(r, k + 1, expo[k]))
        work = p_add(work, term, -coef)
    return out


# --------------------------------------------------------------------------------------------------------------
# SO(2r+1): characters as polynomials in s_k = e_k(cos theta)
# --------------------------------------------------------------------------------------------------------------
def _cos_orbit_poly(mu, r: int) -> dict:
    """sum over the hyperoctahedral orbit of mu of e^{i<w,theta>} as a polynomial in c_i = cos theta_i:
    sum over distinct permutations p of mu of prod_{p_i != 0} 2 T_{p_i}(c_i)."""
    out = {}
    for p in set(itertools.permutations(mu)):
        term = {tuple([0] * r): 1}
        for i, m in enumerate(p):
            if m == 0:
                continue
            t = cheb_T(abs(m))
            f = {tuple(d if j == i else 0 for j in range(r)): 2 * c for d, c in enumerate(t) if c}
            term = p_mul(term, f)
        out = p_add(out, term)
    return out


@lru_cache(maxsize=None)
def _so_odd_orbit_in_elementary(mu: tuple) -> tuple:
    r = len(mu)
    return tuple(sorted(symmetric_to_elementary(_cos_orbit_poly(mu, r), r).items()))


@lru_cache(maxsize=None)
def so_odd_character_poly(n: int, signature: tuple) -> dict:
    """chi_signature of SO(n), n odd, as {exponents over (s_1..s_r): int}."""
    assert n % 2 == 1
    dom = ch.dominant_weight_multiplicities("B", tuple(signature))
    out = {}
    for mu, m in dom.items():
        out = p_add(out, dict(_so_odd_orbit_in_elementary(mu)), m)
    return out


# --------------------------------------------------------------------------------------------------------------
# SU(n): Schur polynomial s_lambda(gamma_1..gamma_n) as a polynomial in e_1..e_{n-1} (e_n = det = 1)
# --------------------------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def su_character_poly(n: int, signature: tuple) -> dict:
    """chi_signature of SU(n) as {exponents over (e_1..e_{n-1}): int}; e_n = 1 has been substituted."""
    weights = ch.character_weights("SU", n, signature)
    full = symmetric_to_elementary(dict(weights), n)
    out = {}
    for expo, c in full.items():
        k = expo[:-1]                      # e_n^p = 1
        out[k] = out.get(k, 0) + c
    return {k: v for k, v in out.items() if v}


@lru_cache(maxsize=None)
def cheb_U(k: int) -> tuple:
    """Coefficients (low -> high) of the Chebyshev polynomial of the second kind, sin((k+1) t) = sin t U_k(cos t)."""
    if k == 0:
        return (1,)
    if k == 1:
        return (0, 2)
    a, b = cheb_U(k - 1), cheb_U(k - 2)
    out = [0] * (k + 1)
    for i, v in enumerate(a):
        out[i + 1] += 2 * v
    for i, v in enumerate(b):
        out[i] -= v
    return tuple(out)


def _sin_orbit_poly(mu, r: int) -> dict:
    """sum over distinct permutations p of mu (all entries > 0) of prod_i U_{p_i - 1}(c_i), a symmetric polynomial in c.
    With prod_i (2 i sin theta_i) in front it is the sign-weighted sum over the hyperoctahedral orbit of mu."""
    out = {}
    for p in set(itertools.permutations(mu)):
        term = {tuple([0] * r): 1}
        for i, m in enumerate(p):
            u = cheb_U(m - 1)
            f = {tuple(d if j == i else 0 for j in range(r)): c for d, c in enumerate(u) if c}
            term = p_mul(term, f)
        out = p_add(out, term)
    return out


@lru_cache(maxsize=None)
def so_even_character_polys(n: int, signature: tuple):
    """chi_signature of SO(n), n = 2r even, as  ( P(s) + i^r z Q(s) ) / 2  with s_k = e_k(cos theta) and
    z = prod_i 2 sin theta_i (the invariant that separates the two SO(2r) classes with the same spectrum).
    Returns (2P, 2Q) as {exponents over (s_1..s_r): int}.  For odd r the z-term is purely imaginary and drops out of
    Re chi, which is all the kernels use."""
    assert n % 2 == 0
    r = n // 2
    dom = ch.dominant_weight_multiplicities("D", tuple(signature))
    P, Q = {}, {}
    for mu, m in dom.items():
        amu = tuple(sorted((abs(v) for v in mu), reverse=True))
        cos_part = dict(symmetric_to_elementary(_cos_orbit_poly(amu, r), r))
        if all(v != 0 for v in mu):
            # W(D_r) mu is half of the hyperoctahedral orbit: (Omega_B + sgn * Xi) / 2
            P = p_add(P, cos_part, m)
            sgn = 1 if mu[-1] > 0 else -1
            sin_part = dict(symmetric_to_elementary(_sin_orbit_poly(amu, r), r))
            Q = p_add(Q, sin_part, m * sgn)
        else:
            P = p_add(P, cos_part, 2 * m)
    return P, Q


def character_poly(group: str, n: int, signature) -> dict:
    if group == "SO":
        if n % 2 == 1:
            return so_odd_character_poly(n, tuple(signature))
        raise NotImplementedError("SO(2r) needs the Pfaffian-sign invariant; see DESIGN.md")
    if group == "SU":
        return su_character_poly(n, tuple(signature))
    raise ValueError(group)


def character_matrix(group: str, n: int, signatures):
    """-> (monomials, M): sorted list of exponent tuples, and M[l][t] integer coefficient of monomial t in chi_l."""
    polys = [character_poly(group, n, s) for s in signatures]
    monos = sorted(set().union(*[set(p) for p in polys]))
    index = {m: i for i, m in enumerate(monos)}
    mat = [[0] * len(monos) for _ in polys]
    for l, p in enumerate(polys):
        for m, c in p.items():
            mat[l][index[m]] = c
    return monos, mat
