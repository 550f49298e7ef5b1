"""Irreducible-representation bookkeeping for SO(n) and SU(n): signatures, Weyl dimensions, Casimir
eigenvalues and *weight multiplicities* (the content of the reference's `precomputed_characters.json`).

The reference ships its characters as sympy-generated JSON tables (`spaces/precomputed_characters.json`,
`plots/precomputed_characters.json`; loaded at `spaces/so.py:210-223`, `spaces/su.py:139-153`) holding,
per irrep, every weight with its multiplicity, 20 irreps per group (100 for SO(3..5) in `plots/`).  Here
the same numbers are produced on demand by Freudenthal's recursion over dominant weights, so any
`order` works for any n, and nothing has to be parsed at construction time.  `tests/test_characters.py`
checks the output against the reference tables entry by entry (when `/root/reference` is mounted) and
against a committed digest of them (always).

Conventions follow the reference exactly, because they fix the truncation set and the feature column
order:

* SO(n): signature = partition with `rank = n//2` parts (last part signed when n is even), enumerated as
  in `spaces/so.py:102-122`; Casimir `|rho+sig|^2 - |rho|^2` with numpy norms (`so.py:196-201`);
  dimension by the Weyl formula (`so.py:180-194`).
* SU(n): signature = n-tuple, last entry 0 (`spaces/su.py:66-77`); Casimir with the mean-centred
  signature (`su.py:124-130`); dimension (`su.py:116-122`).
* irreps are *stably* sorted by the float Casimir and the first `order` are kept (`space.py:36-42`).
"""
from __future__ import annotations

import itertools
import math
import operator
from fractions import Fraction
from functools import lru_cache, reduce

import numpy as np


# --------------------------------------------------------------------------------------------------------------
# signatures, dimensions, Casimir eigenvalues (restated from the reference; order matters)
# --------------------------------------------------------------------------------------------------------------
def _partitions_exact_parts(total: int, parts: int):
    """Partitions of `total` into exactly `parts` positive parts, in the colex order produced by the
    Hindenburg algorithm the reference uses (`utils.py:55-96`): the enumeration order breaks Casimir ties."""
    if parts == 0:
        if total == 0:
            yield []
        return
    if parts == 1:
        if total > 0:
            yield [total]
        return
    if total < parts:
        return
    cur = [total - parts + 1] + [1] * (parts - 1)
    while True:
        yield list(cur)
        if cur[0] - 1 > cur[1]:
            cur[0] -= 1
            cur[1] += 1
            continue
        j = 2
        s = cur[0] + cur[1] - 1
        while j < parts and cur[j] >= cur[0] - 1:
            s += cur[j]
            j += 1
        if j >= parts:
            return
        cur[j] = x = cur[j] + 1
        j -= 1
        while j > 0:
            cur[j] = x
            s -= x
            j -= 1
        cur[0] = s


def so_signatures(n: int) -> list:
    """All candidate SO(n) signatures in the reference's generation order (`so.py:102-122`)."""
    rank = n // 2
    limit = 200 if n == 3 else 30
    out = []
    for total in range(limit):
        for parts in range(rank + 1):
            for p in _partitions_exact_parts(total, parts):
                sig = p + [0] * (rank - parts)
                out.append(tuple(sig))
                if n % 2 == 0 and sig[-1] != 0:
                    sig = list(sig)
                    sig[-1] = -sig[-1]
                    out.append(tuple(sig))
    return out


def su_signatures(n: int) -> list:
    """All candidate SU(n) signatures in the reference's order (`su.py:66-77`): sorted ascending."""
    rank = n - 1
    lim = 100 if n in (1, 2) else 30 if n == 3 else 10
    sigs = [s + (0,) for s in itertools.combinations_with_replacement(range(lim, -1, -1), r=rank)]
    sigs.sort()
    return sigs


def so_rho(n: int) -> np.ndarray:
    rank = n // 2
    rho = np.arange(rank - 1, -1, -1)
    return rho if n % 2 == 0 else rho + 0.5


def su_rho(n: int) -> np.ndarray:
    return np.arange(n - 1, -n, -2) * 0.5


def so_dimension(n: int, sig) -> int:
    rank = n // 2
    if n % 2 == 1:
        qs = [pk + rank - k - 1 / 2 for k, pk in enumerate(sig)]
        d = reduce(operator.mul, (2 * qs[k] / math.factorial(2 * k + 1) for k in range(rank))) * \
            reduce(operator.mul, ((qs[i] - qs[j]) * (qs[i] + qs[j]) for i, j in itertools.combinations(range(rank), 2)), 1)
        return int(round(d))
    qs = [pk + rank - k - 1 if k != rank - 1 else abs(pk) for k, pk in enumerate(sig)]
    d = int(reduce(operator.mul, (2 / math.factorial(2 * k) for k in range(1, rank)), 1)
            * reduce(operator.mul, ((qs[i] - qs[j]) * (qs[i] + qs[j]) for i, j in itertools.combinations(range(rank), 2)), 1))
    return int(round(d))


def su_dimension(n: int, sig) -> int:
    if n == 1:
        return 1
    d = reduce(operator.mul,
               (reduce(operator.mul, (sig[i - 1] - sig[j - 1] + j - i for j in range(i + 1, n + 1))) / math.factorial(n - i)
                for i in range(1, n)))
    return int(round(d))


def so_casimir(n: int, sig) -> float:
    rho = so_rho(n)
    return (np.linalg.norm(rho + np.array(sig)) ** 2 - np.linalg.norm(rho) ** 2).item()


def su_casimir(n: int, sig) -> float:
    s = np.array(sig, dtype=float)
    s -= np.mean(s)
    rho = su_rho(n)
    return (np.linalg.norm(rho + s) ** 2 - np.linalg.norm(rho) ** 2).item()


def truncated_irreps(group: str, n: int, order: int):
    """-> (kept, dropped): lists of (signature, dimension, casimir), stably sorted by casimir (`space.py:36-42`)."""
    if group == "SO":
        sigs = so_signatures(n)
        recs = [(s, so_dimension(n, s), so_casimir(n, s)) for s in sigs]
    elif group == "SU":
        sigs = su_signatures(n)
        recs = [(s, su_dimension(n, s), su_casimir(n, s)) for s in sigs]
    else:
        raise ValueError(group)
    recs.sort(key=operator.itemgetter(2))   # list.sort is stable, as in the reference
    return recs[:order], recs[order:]


# --------------------------------------------------------------------------------------------------------------
# weight multiplicities by Freudenthal's formula
#   root systems in the epsilon basis with the standard inner product:
#     B_r (SO(2r+1)), D_r (SO(2r)), A_{n-1} realised in Z^n (SU(n), weights = exponent vectors of the gammas)
#   coordinates are doubled for B_r so that rho is integral.
# --------------------------------------------------------------------------------------------------------------
def _positive_roots(kind: str, r: int):
    roots = []
    e = lambda i: tuple(1 if k == i else 0 for k in range(r))
    add = lambda a, b: tuple(x + y for x, y in zip(a, b))
    sub = lambda a, b: tuple(x - y for x, y in zip(a, b))
    for i in range(r):
        for j in range(i + 1, r):
            roots.append(sub(e(i), e(j)))
            if kind in ("B", "D"):
                roots.append(add(e(i), e(j)))
    if kind == "B":
        roots.extend(e(i) for i in range(r))
    return roots


def _dominant(kind: str, w):
    """Representative of the Weyl orbit of `w` in the closed dominant chamber."""
    if kind == "A":
        return tuple(sorted(w, reverse=True))
    if kind == "B":
        return tuple(sorted((abs(x) for x in w), reverse=True))
    # D_r: signed permutations with an even number of sign changes
    neg = sum(1 for x in w if x < 0)
    a = sorted((abs(x) for x in w), reverse=True)
    if neg % 2 == 1 and a[-1] != 0:
        a[-1] = -a[-1]
    return tuple(a)


def _is_dominant(kind: str, w) -> bool:
    return _dominant(kind, w) == tuple(w)


def _rho2(kind: str, r: int):
    """2*rho as an integer vector."""
    if kind == "A":
        return tuple(2 * (r - 1 - i) for i in range(r))          # any shift by (c,..,c) gives the same formula
    if kind == "B":
        return tuple(2 * (r - i) - 1 for i in range(r))
    return tuple(2 * (r - 1 - i) for i in range(r))


@lru_cache(maxsize=None)
def dominant_weight_multiplicities(kind: str, highest: tuple) -> dict:
    """{dominant weight mu: multiplicity of mu in V(highest)} by Freudenthal's recursion."""
    r = len(highest)
    roots = _positive_roots(kind, r)
    rho2 = _rho2(kind, r)
    # dominant weights below `highest`: closure under subtracting positive roots inside the chamber
    dom = {tuple(highest)}
    frontier = [tuple(highest)]
    while frontier:
        nxt = []
        for mu in frontier:
            for a in roots:
                nu = tuple(x - y for x, y in zip(mu, a))
                if _is_dominant(kind, nu) and nu not in dom:
                    dom.add(nu)
                    nxt.append(nu)
        frontier = nxt

    def norm2_shift(w):   # |2w + 2rho|^2  (a common factor 4 cancels in the recursion)
        return sum((2 * x + p) ** 2 for x, p in zip(w, rho2))

    top = norm2_shift(highest)
    mult = {tuple(highest): 1}
    # process in order of decreasing "level" <highest - mu, rho-ish>; sorting by norm2_shift descending is a valid
    # topological order because mu + k*alpha (dominant-ised) always has a strictly larger |.+rho|
    todo = sorted(dom - {tuple(highest)}, key=norm2_shift, reverse=True)
    for mu in todo:
        acc = 0
        for a in roots:
            k = 1
            while True:
                nu = tuple(x + k * y for x, y in zip(mu, a))
                m = mult.get(_dominant(kind, nu))
                if m is None:
                    # outside the weight diagram (all further k are outside as well)
                    if _dominant(kind, nu) not in dom:
                        break
                    raise AssertionError("topological order violated")
                acc += m * sum(x * y for x, y in zip(nu, a))
                k += 1
        denom = top - norm2_shift(mu)
        # Freudenthal: (|l+rho|^2 - |mu+rho|^2) m(mu) = 2 * sum <mu+k a, a> m(mu+k a);  denom carries a factor 4
        val = Fraction(8 * acc, denom)
        assert val.denominator == 1 and val > 0, (kind, highest, mu, val)
        mult[mu] = int(val)
    return mult


def _orbit(kind: str, mu):
    """All distinct images of the dominant weight `mu` under the Weyl group."""
    r = len(mu)
    out = set()
    for perm in set(itertools.permutations(mu)):
        if kind == "A":
            out.add(perm)
            continue
        nz = [i for i, x in enumerate(perm) if x != 0]
        for signs in itertools.product((1, -1), repeat=len(nz)):
            w = list(perm)
            for i, s in zip(nz, signs):
                w[i] = s * w[i]
            w = tuple(w)
            if kind == "D" and _dominant("D", w) != tuple(mu):
                continue
            out.add(w)
    return out


def group_kind(group: str, n: int):
    """-> (root-system letter, rank of the weight vectors we carry)."""
    if group == "SU":
        return "A", n
    if group == "SO":
        return ("B" if n % 2 else "D"), n // 2
    raise ValueError(group)


def character_weights(group: str, n: int, signature) -> dict:
    """{weight (tuple of ints): multiplicity} for the irrep `signature` — every weight, not only dominant ones.
    For SO(n) a weight (w_1..w_r) is the Laurent monomial prod gamma_i^{w_i}; for SU(n) the monomial
    prod gamma_i^{w_i} with all w_i >= 0 (a Schur polynomial term)."""
    kind, _ = group_kind(group, n)
    if group == "SO" and n == 3:
        kind = "B"
    dom = dominant_weight_multiplicities(kind, tuple(signature))
    out = {}
    for mu, m in dom.items():
        for w in _orbit(kind, mu):
            out[w] = m
    return out


def reference_table_entry(group: str, n: int, signature):
    """The irrep's character in the reference JSON layout `[coeffs, monomial-exponents]`
    (`so.py:214-223`): SO exponents over (gamma_1..gamma_r, conj gamma_1..conj gamma_r), SU over gamma_1..gamma_n."""
    weights = character_weights(group, n, signature)
    coeffs, monoms = [], []
    for w in sorted(weights, reverse=True):
        if group == "SO":
            monoms.append([max(x, 0) for x in w] + [max(-x, 0) for x in w])
        else:
            monoms.append(list(w))
        coeffs.append(weights[w])
    return coeffs, monoms
