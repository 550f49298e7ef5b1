"""numpy emulation of the CUDA data path (embedding -> bilinear forms -> epilogue) driven by the very same
`LskEpilogue` parameter block the kernels receive.  Test infrastructure: lets the CPU suite validate the host
tables (classfn.py) end to end against the reference goldens without a GPU."""
import itertools

import numpy as np

from liestationarykernels_b200 import _lib


def compound(x, k):
    """Lambda^k of each matrix in the batch: [num, C(n,k), C(n,k)] (same lexicographic order as lsk_embed.cu)."""
    n = x.shape[-1]
    idx = list(itertools.combinations(range(n), k))
    out = np.zeros((x.shape[0], len(idx), len(idx)), dtype=x.dtype)
    for a, I in enumerate(idx):
        for b, J in enumerate(idx):
            out[:, a, b] = np.linalg.det(x[:, I, :][:, :, J]) if k > 1 else x[:, I[0], J[0]]
    return out


_EVEN = [m for m in range(32) if bin(m).count("1") % 2 == 0]


def _blade_sign(a, b):
    s = 0
    a >>= 1
    while a:
        s += bin(a & b).count("1")
        a >>= 1
    return -1.0 if s & 1 else 1.0


def spin5_rotor(x):
    """16 rotor coefficients (index = even blade mask >> 1) of the Spin(5) lift, by Givens elimination."""
    m = np.array(x, dtype=np.float64)
    R = np.zeros(16)
    R[0] = 1.0
    for j in range(4):
        for i in range(j + 1, 5):
            a, b = m[j, j], m[i, j]
            r = np.hypot(a, b)
            if r == 0:
                continue
            c, s = a / r, b / r
            rj, ri = m[j].copy(), m[i].copy()
            m[j], m[i] = c * rj + s * ri, -s * rj + c * ri
            if c >= 0:
                ch = np.sqrt((1 + c) / 2)
                sh = s / (2 * ch)
            else:
                sh = (1.0 if s >= 0 else -1.0) * np.sqrt((1 - c) / 2)
                ch = s / (2 * sh)
            B = (1 << j) | (1 << i)
            new = ch * R
            for idx in range(16):
                A = _EVEN[idx]
                new[(A ^ B) >> 1] += -sh * _blade_sign(A, B) * R[idx]
            R = new
    return R


def embed(plan, x, side):
    segs = plan.segments_a if side == "a" else plan.segments_b
    cols = []
    for (k, part) in segs:
        if k == 100:
            cols.append(np.stack([spin5_rotor(xi) for xi in x]))
            continue
        c = compound(x, k)
        if part == 5:                       # SU(4): Lambda^2 in the real basis of Lambda^2 C^4 (csrc/lsk_embed.cu, part 5)
            s2 = 1.0 / np.sqrt(2.0)
            B = np.zeros((6, 6), dtype=complex)
            for p_, sg in ((0, 1.0), (1, -1.0), (2, 1.0)):
                B[p_, 2 * p_], B[5 - p_, 2 * p_] = s2, sg * s2
                B[p_, 2 * p_ + 1], B[5 - p_, 2 * p_ + 1] = 1j * s2, -1j * sg * s2
            r = np.einsum("ia,nij,jb->nab", B.conj(), c, B)
            assert np.abs(r.imag).max() < 1e-12
            cols.append(r.real.reshape(x.shape[0], -1))
            continue
        if part == 4:                       # Hodge star on the row index: (I, J) -> sign(I, I^c) * minor(I^c, J)
            n = x.shape[-1]
            idx = list(itertools.combinations(range(n), k))
            star = np.zeros_like(c)
            for a, I in enumerate(idx):
                comp = tuple(v for v in range(n) if v not in I)
                perm = list(I) + list(comp)
                sign = 1
                for i in range(n):
                    for j in range(i + 1, n):
                        if perm[i] > perm[j]:
                            sign = -sign
                star[:, a, :] = sign * c[:, idx.index(comp), :]
            c = star
        c = c.reshape(x.shape[0], -1)
        if part == 0 or part == 4:
            v = c.real
        elif part >= 6:                    # one real number per complex entry: re, im, re + im, re - im
            v = {6: c.real, 7: c.imag, 8: c.real + c.imag, 9: c.real - c.imag}[part]
        else:
            if part == 3:
                v = np.stack((-c.imag, c.real), axis=-1).reshape(x.shape[0], -1)
            else:
                v = np.stack((c.real, c.imag), axis=-1).reshape(x.shape[0], -1)
        pad = (-v.shape[1]) % 4
        cols.append(np.pad(v, ((0, 0), (0, pad))))
    return np.concatenate(cols, axis=1)


def run_epilogue(ep, forms):
    """forms: [nforms, N, M] -> [N, M]"""
    sq = (ep.reserved >> 8) & 0xF
    w = [forms[f] * forms[f] if (sq >> f) & 1 else forms[f] for f in range(ep.nforms)]
    var = [ep.aff[v][0] + sum(ep.aff[v][1 + f] * w[f] for f in range(ep.nforms)) for v in range(ep.nvars)]
    coef = np.array(ep.coef[:])
    if ep.kind == _lib.EPI_CHEB1:
        x = var[0]
        b1 = np.zeros_like(x)
        b2 = np.zeros_like(x)
        for k in range(ep.nterms - 1, 0, -1):
            b1, b2 = 2 * x * b1 + (coef[k] - b2), b1
        res = x * b1 + (coef[0] - b2)
    elif ep.kind == _lib.EPI_STAIR2:
        s1, s2 = var
        res = np.zeros_like(s1)
        for slot in range(ep.nterms):
            ln = 2 * ((ep.stair_shape >> (4 * slot)) & 15)
            assert ln == ep.row_len[slot]
            inner = np.full_like(s1, coef[16 * slot])
            for a in range(1, ln):
                inner = inner * s1 + coef[16 * slot + a]
            res = res * s2 + inner
    elif ep.kind == _lib.EPI_ORBIT2:
        s1, s2 = var
        D = ep.nterms
        Dp = 8 if D <= 8 else 12 if D <= 12 else 16 if D <= 16 else 20
        rt = np.sqrt(np.maximum(s1 * s1 - 4 * s2, 0))
        c1, c2 = 0.5 * (s1 + rt), 0.5 * (s1 - rt)
        T1 = [np.ones_like(c1), c1]
        T2 = [np.ones_like(c2), c2]
        for k in range(2, Dp):
            T1.append(2 * c1 * T1[-1] - T1[-2])
            T2.append(2 * c2 * T2[-1] - T2[-2])
        res = np.zeros_like(s1)
        for a in range(D):
            row = sum(coef[a * Dp + b] * T2[b] for b in range(Dp))
            res = res + T1[a] * row
    elif ep.kind == _lib.EPI_SPARSE:
        res = np.zeros_like(var[0])
        for t in range(ep.nterms):
            packed = int(np.array([coef[2 * t + 1]]).view(np.uint64)[0])
            m = np.full_like(var[0], coef[2 * t])
            for i in range(ep.nvars):
                e = (packed >> (8 * i)) & 0xFF
                m = m * var[i] ** e
            res = res + m
    elif ep.kind == _lib.EPI_HORNER:
        a = [np.zeros_like(var[0]) for _ in range(ep.nvars)]
        for t in range(ep.nterms):
            c, op = coef[2 * t], int(np.array([coef[2 * t + 1]]).view(np.uint64)[0])
            if op == 0:
                a[0] = a[0] * var[0] + c
            elif op == 1:
                a[0] = np.full_like(var[0], c)
            else:
                lvl, kind = 1 + (op - 2) // 3, (op - 2) % 3
                a[lvl] = a[lvl] * var[lvl] + a[lvl - 1] if kind == 0 else a[lvl - 1].copy() if kind == 1 else a[lvl] * var[lvl]
        res = a[ep.nvars - 1]
    else:
        raise ValueError(ep.kind)
    return ep.scale * res


def pairwise(plan, x, y, ep):
    A, B = embed(plan, x, "a"), embed(plan, y, "b")
    forms = []
    for f in range(ep.nforms):
        lo, hi = 4 * ep.group_begin[f], 4 * ep.group_end[f]
        forms.append(A[:, lo:hi] @ B[:, lo:hi].T)
    return run_epilogue(ep, np.array(forms))
