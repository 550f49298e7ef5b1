"""Kernel front-ends — drop-in for the reference's `spectral_kernel.py` (same classes, constructor arguments,
`kernel(x, y=None, normalize=True)`, `.normalizer`, training-mode side effects and RNG consumption order).

The forward passes do not materialise anything of size N*M except the result: see csrc/lsk_pairwise.cu."""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np
import torch

from . import _lib, ops
from .space import AbstractManifold, CompactLieGroup
from .spectral_measure import AbstractSpectralMeasure

dtype = torch.float64
# snapshot polling (`EigenbasisSumKernel._finish_hyper_read`): spin this long, then sleep in slices.  Measured on 8 GPUs (one rank per
# GPU, 0.31 ms steps): pure spinning keeps every rank within 0.5 % of the others; sleeping after 40 us costs the slowest rank 3 %.
SPIN_SECONDS = float(os.environ.get("LSK_SPIN_US", "5000")) * 1e-6
SLEEP_SECONDS = float(os.environ.get("LSK_SLEEP_US", "20")) * 1e-6


class AbstractSpectralKernel(torch.nn.Module):
    def __init__(self, measure: AbstractSpectralMeasure, manifold: AbstractManifold):
        super().__init__()
        self.measure = measure
        self.manifold = manifold

    def forward(self, *args):
        raise NotImplementedError


def spectral_weights(measure, manifold):
    """Host copy of a(lambda_l) for the manifold's irreps (spectral_measure.py:24-26, 35-36): one tiny device op and
    one L-element device->host read per call; all N*M work is in the CUDA kernel."""
    with torch.no_grad():
        a = measure(manifold._lambdas)
    return a.detach().cpu().numpy()


def spectral_weights_tensor(measure, manifold):
    with torch.no_grad():
        return measure(manifold._lambdas).detach().reshape(-1)


class _WeightedCharacterSum(torch.autograd.Function):
    """K = sum_l c_l(theta) d_l Re chi_l(x_i y_j^{-1}) with hyper-parameter gradients.

    d/dtheta of the covariance is the same class-function sum with the weights dc_l/dtheta, so the backward pass is
    one more launch of the fused kernel per hyper-parameter followed by <G, K'>: nothing per-irrep of size N*M is ever
    stored (the reference keeps L such tensors alive in its autograd graph).  dc/dtheta comes from forward-mode AD of
    the tiny L-vector function theta -> c(theta).  Inputs x, y get no gradient."""

    @staticmethod
    def forward(ctx, kernel, x, y, normalize, out, *params):
        ctx.kernel, ctx.normalize = kernel, normalize
        ctx.x, ctx.y = x, y
        ctx.n_params = len(params)
        with torch.no_grad():
            c = kernel._coefficient_vector(normalize, params)
        if out is not None:
            ctx.mark_dirty(out)                           # `out` is an input written in place and returned
        return kernel._evaluate(x, y, c.detach().cpu().numpy(), out)

    @staticmethod
    def backward(ctx, grad):
        kernel = ctx.kernel
        params = kernel._hyper_parameters()
        grads = []
        for i in range(ctx.n_params):
            if not ctx.needs_input_grad[5 + i]:
                grads.append(None)
                continue

            def f(p, i=i):
                ps = list(params)
                ps[i] = p
                return kernel._coefficient_vector(ctx.normalize, ps)

            with torch.enable_grad():
                _, tangent = torch.func.jvp(f, (params[i].detach(),), (torch.ones_like(params[i]),))
            with torch.no_grad():
                dk = kernel._evaluate(ctx.x, ctx.y, tangent.detach().cpu().numpy(), None)
                grads.append(torch.dot(grad.reshape(-1), dk.reshape(-1)).reshape(params[i].shape))    # one pass, no N x M temporary
        return (None, None, None, None, None, *grads)


class EigenbasisSumKernel(AbstractSpectralKernel):
    """k(x, y) = |sigma^2| / norm * sum_l a(lambda_l) d_l Re chi_l(x y^{-1})   (spectral_kernel.py:26-54)."""

    def __init__(self, measure, manifold, form="auto"):
        super().__init__(measure, manifold)
        self.form = form
        self.assume_static = False
        self._ep_cache = {}
        self._pinned = self._pinned_np = None
        point = manifold.rand()
        self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def compute_normalizer(self):
        point = self.manifold.rand()
        self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def __getstate__(self):          # copies / pickles start without the launch-time caches (coefficient block, pinned buffer)
        state = self.__dict__.copy()
        state["_ep_cache"], state["_pinned"], state["_pinned_np"] = {}, None, None
        return state

    def forward(self, x, y=None, normalize=True, out=None):
        if y is None:
            y = x
        if self.training:
            if normalize:
                self.compute_normalizer()
        manifold = self.manifold
        params = self._hyper_parameters()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params) and \
                (isinstance(manifold, CompactLieGroup) or hasattr(manifold, "_eigensum_coefs")):
            return _WeightedCharacterSum.apply(self, x, y, normalize, out, *params)
        if not isinstance(manifold, CompactLieGroup):
            return manifold._eigensum(self, x, y, normalize, out)
        plan = manifold.plan
        same = y is x and plan.segments_a == plan.segments_b      # one embedding serves both sides
        cached = self._ep_cache.get("ep") if self._ep_cache.get("flags") == (normalize, self.form) else None
        pending = None
        capturing = cached is not None and torch.cuda.is_current_stream_capturing()
        if cached is not None and not self.assume_static and not capturing:
            pending = self._enqueue_hyper_read(normalize)
        if capturing:
            ep = cached               # inside a CUDA-graph capture the host cannot wait for the device: the captured launch keeps the block
        elif pending is not None:
            # Optimistic launch: start the stream-ordered read of the hyper-parameters, launch with the cached coefficient
            # block, and only then wait for the read.  The read completes when the work queued BEFORE this call has finished, so
            # the host stays one call ahead of the GPU instead of draining it on every call (a blocking read cost 90 us of idle
            # GPU per 2.4 ms step at C2).  If a parameter did change, the block is rebuilt and the launch repeated into the same
            # output - stream order makes the second result the final one.
            ep = cached
        else:
            ep = self._epilogue(normalize)
        if same:
            if pending is not None:
                ops.snapshot_launch(pending[4])
            ea = eb = ops.embed(plan, x, "a")
        else:
            # one launch for both operands; it also writes the hyper-parameter snapshot
            ea, eb = ops.embed_pair(plan, x, y, snapshot=None if pending is None else pending[4])
        res = ops.pairwise_poly(ea, eb, x.shape[0], y.shape[0], plan.kg_total, ep, out=out, symmetric=same)
        if pending is not None and self._finish_hyper_read(pending) != self._ep_cache["key"]:
            ep = self._epilogue(normalize)
            res = ops.pairwise_poly(ea, eb, x.shape[0], y.shape[0], plan.kg_total, ep, out=res, symmetric=same)
        return res

    # ---- differentiable path ---------------------------------------------------------------------------------
    def _hyper_parameters(self):
        return [self.measure.lengthscale, self.measure.variance]

    def _coefficient_vector(self, normalize, params):
        """c_l = |sigma^2| a_l(lengthscale) / normaliser as a torch function of the hyper-parameters.  In training mode
        the normaliser is the kernel at the identity class, sum_l a_l d_l^2, recomputed from the same parameters
        (spectral_kernel.py:33-35, 41-43); in eval mode it is the stored constant."""
        ls, var = params
        m = self.measure
        lam = self.manifold._lambdas
        if hasattr(m, "nu"):
            a = torch.pow(m.nu / (ls * ls / 2.0) + lam, -m.nu - m.dim / 2)
        else:
            a = torch.exp(-ls * ls / 2.0 * lam)
        if not normalize:
            return a
        if not self.training:
            norm = torch.as_tensor(self.normalizer, dtype=dtype, device=a.device).detach()
        elif isinstance(self.manifold, CompactLieGroup):
            dims = torch.as_tensor(self.manifold._dims, dtype=dtype, device=a.device)
            norm = (a * dims * dims).sum()
        else:       # homogeneous space: phi_l(id, id) = d_l * (subgroup average of chi_l), fixed by the h samples
            norm = (a * self.manifold._identity_values().to(a.device)).sum()
        return torch.abs(var[0]) * a / norm

    def _evaluate(self, x, y, coefs, out):
        manifold = self.manifold
        if not isinstance(manifold, CompactLieGroup):
            return manifold._eigensum_coefs(x, y, np.asarray(coefs, dtype=np.float64), 1.0, out)
        plan = manifold.plan
        ep = plan.epilogue(np.asarray(coefs, dtype=np.float64) * manifold._dims, 1.0, form=self.form)
        same = y is x and plan.segments_a == plan.segments_b      # one embedding serves both sides
        if same:
            ea = eb = ops.embed(plan, x, "a")
        else:
            ea, eb = ops.embed_pair(plan, x, y)                    # one launch for both operands
        return ops.pairwise_poly(ea, eb, x.shape[0], y.shape[0], plan.kg_total, ep, out=out, symmetric=same)


    def _hyper_tensors(self, normalize):
        m = self.measure
        dev, host = [], []
        for t in (m.lengthscale, m.variance, getattr(m, "nu", None), self.normalizer if normalize else None):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                dev.append(t.detach().reshape(-1)[:1].to(dtype))
            elif t is not None:
                host.append(float(t))
        return dev, host

    def _hyper_values(self, normalize):
        """Current (lengthscale, variance, nu, normaliser) as host floats: one concatenation and ONE device->host read."""
        dev, host = self._hyper_tensors(normalize)
        vals = torch.cat(dev).tolist() if dev else []
        return (normalize, self.form) + tuple(vals + host)

    def _enqueue_hyper_read(self, normalize):
        """Asynchronous version of `_hyper_values`: the device scalars are copied into host-mapped pinned memory by the embedding
        launch of this call (`lsk_embed_points2_snapshot`; a 32-thread launch of its own, `lsk_snapshot_doubles`, where there is
        no two-operand embedding), each with a tag the host polls; no copy engine, no event.  (torch.cat + a pinned copy + an event
        put two extra launches and a DMA in front of the embedding kernel: +24 us per step at 8 GPUs, where a step is 0.31 ms.)
        Returns the pending-read record; its last field is what the launch needs."""
        m = self.measure
        ptrs, host = [], []
        for t in (m.lengthscale, m.variance, getattr(m, "nu", None), self.normalizer if normalize else None):
            if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.numel() >= 1:
                ptrs.append(t.data_ptr())
            elif isinstance(t, torch.Tensor):
                return None                                   # unusual parameter (dtype / device): blocking read instead
            elif t is not None:
                host.append(float(t))
        if self._pinned is None:
            self._pinned = torch.zeros(16, dtype=dtype).pin_memory()
            self._pinned_np = self._pinned.numpy().view("uint64")
            self._seq = 0
        self._seq += 1
        arr = (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
        launch = (arr, len(ptrs), ctypes.c_void_p(self._pinned.data_ptr()), self._seq)     # handed to the embedding launch
        return (len(ptrs), host, normalize, self._seq, launch)

    def _finish_hyper_read(self, pending):
        """Poll the pinned slots until every [value, tag] pair is consistent for this call's sequence number (see lsk.h)."""
        n, host, normalize, seq, _ = pending
        words = self._pinned_np
        key = (seq * 0x9E3779B97F4A7C15 + 0x7F4A7C159E3779B9) & 0xFFFFFFFFFFFFFFFF
        # Spin (the wait is at most one call of GPU work), then sleep in short slices if the queue ahead of this call is long.
        t_start = time.perf_counter()
        t_spin = t_start + SPIN_SECONDS
        while True:
            vals = [int(words[2 * t]) for t in range(n)]
            if all((vals[t] ^ key) == int(words[2 * t + 1]) for t in range(n)):
                break
            now = time.perf_counter()
            if now > t_spin:
                if now - t_start > 2.0:                        # never observed; keeps a lost update from hanging the host
                    torch.cuda.current_stream().synchronize()
                    vals = [int(words[2 * t]) for t in range(n)]
                    if all((vals[t] ^ key) == int(words[2 * t + 1]) for t in range(n)):
                        break
                    raise _lib.LskError("hyper-parameter snapshot did not arrive")
                time.sleep(SLEEP_SECONDS)
        floats = np.array(vals, dtype=np.uint64).view(np.float64).tolist() if n else []
        return (normalize, self.form) + tuple(floats + host)

    def _epilogue(self, normalize):
        """Coefficient block for the current hyper-parameters.  The cache is keyed on the parameter VALUES, read back on
        every call (a write through `.data` changes neither `_version` nor the storage pointer, and the reference re-evaluates
        the measure on every call, so nothing weaker is a drop-in).  A caller that guarantees frozen hyper-parameters can set
        `kernel.assume_static = True` to skip the read (tensor identity / version keys only) and `invalidate()` by hand."""
        m = self.measure
        if self.assume_static:
            key = [normalize, self.form]
            for t in (m.lengthscale, m.variance, getattr(m, "nu", None), self.normalizer if normalize else None):
                key.append(None if t is None else (id(t), t._version, t.data_ptr()) if isinstance(t, torch.Tensor) else t)
            key = tuple(key)
        else:
            key = self._hyper_values(normalize)
        hit = self._ep_cache.get("key") == key
        if not hit:
            manifold = self.manifold
            w = spectral_weights(m, manifold) * manifold._dims
            scale = float(torch.abs(m.variance[0]).item() / float(self.normalizer)) if normalize else 1.0
            self._ep_cache = {"key": key, "flags": (normalize, self.form), "ep": manifold.plan.epilogue(w, scale, form=self.form)}
        return self._ep_cache["ep"]

    def invalidate(self):
        """Drop the cached coefficient block (only needed with `assume_static = True`)."""
        self._ep_cache = {}

    # ---- host-buffer entry: covariance written into host memory ------------------------------------------------
    HOST_ROW_BLOCK = 2048        # rows per device block of `evaluate_to_host` (a multiple of 128)

    def evaluate_to_host(self, x, y=None, out=None, normalize=True, row_block=None):
        """kernel(x, y) with HOST buffers on both sides: `x`, `y` may be host tensors (pinned for asynchronous copies), the
        result is written into `out`, a float64 [N, M] host tensor (allocated pinned when None) and returned once complete.
        The N x M covariance never exists on the device: the operands are embedded once, row blocks of `row_block` rows are
        computed into two alternating device buffers, and each block's device->host copy runs on a second stream while the
        next block is computed, so the call costs the PCIe time of the result plus one block of compute (at C2, 16384^2:
        the 2.1 GB copy dominates 2.4 ms of compute).  Values are those of `forward` bit for bit (same tiles, same
        coefficient block)."""
        manifold = self.manifold
        dev = manifold._lambdas.device
        xd = x.to(dev, non_blocking=True)
        yd = xd if y is None or y is x else y.to(dev, non_blocking=True)
        N, M = xd.shape[0], yd.shape[0]
        if out is None:
            out = torch.empty((N, M), dtype=torch.float64).pin_memory()
        elif out.is_cuda or out.dtype != torch.float64 or tuple(out.shape) != (N, M) or (M and out.stride(1) != 1):
            raise ValueError("out must be a float64 [%d, %d] host tensor with unit column stride" % (N, M))
        rb = int(row_block or self.HOST_ROW_BLOCK)
        if rb <= 0 or rb % 128:
            raise ValueError("row_block must be a positive multiple of 128")
        with torch.no_grad():
            if not isinstance(manifold, CompactLieGroup) or N <= rb or M == 0:
                if N and M:
                    out.copy_(self.forward(xd, yd, normalize=normalize))
                return out
            if self.training and normalize:
                self.compute_normalizer()                 # as `forward` does (one draw from the generator per call)
            plan = manifold.plan
            ep = self._epilogue(normalize)
            ea, eb = ops.embed_pair(plan, xd, yd)
            main = torch.cuda.current_stream(dev)
            side = getattr(self, "_copy_stream", None)
            if side is None or side.device != main.device:
                side = self._copy_stream = torch.cuda.Stream(dev)
            bufs = [torch.empty((rb, M), dtype=torch.float64, device=dev) for _ in range(2)]
            for b in bufs:
                b.record_stream(side)
            drained = [None, None]
            for i, r0 in enumerate(range(0, N, rb)):
                rows = min(rb, N - r0)
                if drained[i % 2] is not None:
                    main.wait_event(drained[i % 2])                  # the block's buffer is free once its last copy has left
                block = bufs[i % 2][:rows]
                ops.pairwise_poly(ea[(r0 // 64) * plan.kg_total * 256:], eb, rows, M, plan.kg_total, ep, out=block)
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    out[r0:r0 + rows].copy_(block, non_blocking=True)
                    drained[i % 2] = torch.cuda.Event()
                    drained[i % 2].record(side)
            main.wait_stream(side)                                   # stream order: work queued after this call sees the copies done
            if not out.is_pinned():
                return out                                           # pageable destination: the copies above were synchronous
            side.synchronize()
        return out


class EigenbasisKernel(AbstractSpectralKernel):
    """k(x, y) = sum_l a(lambda_l) sum_k f_{l,k}(x) conj(f_{l,k}(y)) over an explicit orthonormal basis of every eigenspace
    (spectral_kernel.py:57-75; `eigenspace.basis` = TranslatedCharactersBasis on Lie groups).  Equal to EigenbasisSumKernel in
    exact arithmetic (the basis functions of an irrep sum to d chi(x y^{-1})); it exists for API parity and as a cross-check,
    and is usable for small irreps only (dim^2 basis functions, a dim^2 x dim^2 Cholesky per irrep).

    The reference's forward has a shape slip: with `f_x = f(x).T` ([dim^2, N]) it accumulates `f_x @ f_y.T`, a
    [dim^2, dim^2] matrix, into the [N, M] result, which only broadcasts when N = M = dim^2.  The evident intent -
    `f(x) @ f(y)^H`, real part - is what is computed here."""

    def __init__(self, measure, manifold):
        super().__init__(measure, manifold)
        point = self.manifold.id
        self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def forward(self, x, y=None, normalize=True):
        if y is None:
            y = x
        cov = torch.zeros(len(x), len(y), dtype=dtype, device=x.device)
        for eigenspace in self.manifold.lb_eigenspaces:
            f = eigenspace.basis
            f_x = f(x)                                   # [N, dim^2]
            f_y = f_x if y is x else f(y)
            cov = cov + self.measure(eigenspace.lb_eigenvalue) * (f_x @ f_y.conj().T).real
        if normalize:
            return cov / self.normalizer
        return cov


class RandomPhaseKernel(AbstractSpectralKernel):
    """Feature-map approximation with random phases (spectral_kernel.py:78-127):
       Phi[n, l*P + p] = sqrt(a_l / P) * Re phase_function_l(phase_p, x_n),   k = |sigma^2| / norm * Phi_x Phi_y^T."""

    def __init__(self, measure, manifold, phase_order=100):
        super().__init__(measure, manifold)
        self.approx_order = manifold.order
        self.phase_order = phase_order
        self.phases = self.sample_phases()
        point = self.manifold.rand()
        with torch.no_grad():
            self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def compute_normalizer(self):
        point = self.manifold.id
        with torch.no_grad():
            self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def sample_phases(self):
        return self.manifold.rand(self.phase_order)

    def _row_scales(self, extra=1.0):
        a = spectral_weights(self.measure, self.manifold)
        return np.sqrt(a * extra) / np.sqrt(self.phase_order)

    def _features(self, x, scales, panel):
        return self.manifold._phase_features(x, self.phases, scales, panel)

    def make_embedding(self, x):
        """[N, L*P] float64, irrep-major / phase-minor columns (spectral_kernel.py:96-109)."""
        return self._features(x, self._row_scales(), panel=False)

    def _coefficients(self, normalize, q_identity):
        """c_l = a_l / P  or  |sigma^2| a_l / (P norm) as a torch function of (lengthscale, variance).  In training mode the
        normaliser is the un-normalised kernel at the identity, sum_l a_l q_l with q_l = mean_p phi_l(phase_p, id)^2,
        recomputed from the same parameters (spectral_kernel.py:89-91, 115-117); in eval mode the stored constant."""
        m = self.measure
        lam = self.manifold._lambdas
        ls, var = m.lengthscale, m.variance
        if hasattr(m, "nu"):
            a = torch.pow(m.nu / (ls * ls / 2.0) + lam, -m.nu - m.dim / 2)
        else:
            a = torch.exp(-ls * ls / 2.0 * lam)
        if not normalize:
            return a / self.phase_order
        norm = (a * q_identity).sum() if q_identity is not None else torch.as_tensor(self.normalizer, dtype=dtype, device=a.device).detach()
        return torch.abs(var[0]) * a / (self.phase_order * norm)

    def _differentiable(self, x, y, normalize):
        """forward with hyper-parameter gradients (what `examples/gpr_experiments.py` trains on homogeneous spaces)"""
        from .lazy import DifferentiableLazyGram
        L, P = len(self.manifold.lb_eigenspaces), self.phase_order
        q = None
        if self.training and normalize:
            with torch.no_grad():
                unit = self.manifold._phase_features(self.manifold.id, self.phases, np.ones(L), False)      # [1, L * P]
                q = (unit.reshape(L, P) ** 2).mean(dim=1)
        c = self._coefficients(normalize, q)
        if q is not None:               # the state side effect of compute_normalizer (spectral_kernel.py:89-91)
            with torch.no_grad():
                self.normalizer = (spectral_weights_tensor(self.measure, self.manifold) * q).sum()
        scales = np.sqrt(c.detach().cpu().numpy())
        fx = self._features(x, scales, panel=True)
        fy = fx if y is x else self._features(y, scales, panel=True)
        return DifferentiableLazyGram(fx, fy, c, L, P)

    def forward(self, x, y=None, normalize=True):
        if y is None:
            y = x
        if torch.is_grad_enabled() and (self.measure.lengthscale.requires_grad or self.measure.variance.requires_grad):
            return self._differentiable(x, y, normalize)
        if self.training:
            if normalize:
                self.compute_normalizer()
        scales = self._row_scales()
        fx = self._features(x, scales, panel=True)
        fy = fx if y is x else self._features(y, scales, panel=True)
        scale = float(torch.abs(self.measure.variance[0]).item() / float(self.normalizer)) if normalize else 1.0
        from .lazy import LazyGram, as_reference_lazy
        return as_reference_lazy(LazyGram(fx, fy, scale))


def _rff_scale(kernel, normalize):
    if not normalize:
        return 1.0
    return (abs(float(kernel.measure.variance[0].detach())) / float(kernel.normalizer)) ** 0.5


class RandomFourierFeatureKernel(AbstractSpectralKernel):
    """Random Fourier features on a non-compact symmetric space (spectral_kernel.py:164-197):
       Psi = sqrt(|sigma^2| / norm) [Re F, Im F],  k = Psi_x Psi_y^T  (returned lazily)."""

    def __init__(self, measure, manifold):
        super().__init__(measure, manifold)
        manifold.generate_lb_eigenspaces(measure)
        point = self.manifold.id
        with torch.no_grad():
            self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def compute_normalizer(self):
        point = self.manifold.id
        with torch.no_grad():
            self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def _differentiable(self, x, y, normalize):
        """Forward with lengthscale / variance gradients.  The spectral samples move with the lengthscale (lmd = samples *
        scale(l), hyperbolic.py:43-52, spd.py:27-43) and the c-function with them; both stay torch functions of the parameters
        on the host, the N * M * m work of forward and backward runs on the feature and tile kernels (`lazy._FourierGramFn`).
        Same state side effects as the plain path (training mode: features regenerated, normaliser recomputed)."""
        from .lazy import DifferentiableFourierGram
        from .spaces.noncompact import _SphericalFunctions
        m, meas = self.manifold, self.measure
        lmd, shift, coeff = m._differentiable_spectrum(meas, resample=self.training)
        funcs = _SphericalFunctions(lmd.detach(), shift, m, coeff.detach())
        if self.training:
            m.lb_eigenspaces = funcs
        if normalize:
            if self.training:
                norm = (coeff * coeff).sum()
                self.normalizer = norm.detach()
            else:
                norm = torch.as_tensor(self.normalizer, dtype=dtype, device=lmd.device).detach()
            amp = torch.abs(meas.variance[0]) / norm
        else:
            amp = torch.ones((), dtype=dtype, device=lmd.device)
        return DifferentiableFourierGram(m, funcs, x, y, lmd, coeff, amp)

    def forward(self, x, y=None, normalize=True):
        if y is None:
            y = x
        if torch.is_grad_enabled() and hasattr(self.manifold, "_differentiable_spectrum") and \
                (self.measure.lengthscale.requires_grad or self.measure.variance.requires_grad):
            return self._differentiable(x, y, normalize)
        if self.training:
            self.manifold.generate_lb_eigenspaces(self.measure)
            if normalize:
                self.compute_normalizer()
        manifold = self.manifold
        s = _rff_scale(self, normalize)
        px = manifold._features(manifold.to_group(x), manifold.lb_eigenspaces, s, "panel")
        py = px if y is x else manifold._features(manifold.to_group(y), manifold.lb_eigenspaces, s, "panel")
        from .lazy import LazyGram, as_reference_lazy
        return as_reference_lazy(LazyGram(px, py, 1.0))


class RandomSpectralKernel(AbstractSpectralKernel):
    """Pairwise form (spectral_kernel.py:130-161): k(x, y) = Re sum_k F_k(x y^{-1}) conj(F_k(id)).
    Like the reference this needs O(N*M*m) feature evaluations; it exists for API parity (the reference's tests use it
    to cross-check the feature-map kernel on small inputs)."""

    def __init__(self, measure, manifold):
        super().__init__(measure, manifold)
        manifold.generate_lb_eigenspaces(measure)
        point = self.manifold.rand()
        self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def compute_normalizer(self):
        point = self.manifold.rand()
        self.normalizer = self.forward(point, point, normalize=False)[0, 0]

    def _differentiable(self, x, y, normalize):
        """Forward with lengthscale / variance gradients (the reference's forward is plain torch and therefore differentiable:
        spectral_kernel.py:140-161).  Same machinery as the feature-map kernel: the pairwise differences are the "points", the
        identity is the only other point, K = Gram(F(diff), F(id)) through `lazy._FourierGramFn`.  Training mode mirrors the
        reference's call order: features are generated, `compute_normalizer` draws a point and its nested forward generates them
        AGAIN (a fresh draw on SPD(n), spd.py:28) - that last set serves both the normaliser and the covariance."""
        from .lazy import DifferentiableFourierGram
        from .spaces.noncompact import _SphericalFunctions
        m, meas = self.manifold, self.measure
        lmd, shift, coeff = m._differentiable_spectrum(meas, resample=self.training)
        if self.training and normalize:
            m.rand()                                                             # the point of compute_normalizer (RNG order)
            lmd, shift, coeff = m._differentiable_spectrum(meas, resample=True)
        funcs = _SphericalFunctions(lmd.detach(), shift, m, coeff.detach())
        if self.training:
            m.lb_eigenspaces = funcs
        x_, y_ = m.to_group(x), m.to_group(y)
        diff = m.pairwise_diff(x_, y_)
        ident = m.to_group(m.id) if m.id.dim() == 3 else m.id
        one = torch.ones((), dtype=dtype, device=lmd.device)
        cov = DifferentiableFourierGram(m, funcs, diff, ident, lmd, coeff, one, grouped=True).evaluate()[:, 0].view(x.size()[0], y.size()[0])
        if not normalize:
            return cov
        if self.training:
            norm = (coeff * coeff).sum()                                         # the kernel at coinciding points: sum_k |F_k(id)|^2
            self.normalizer = norm.detach()
        else:
            norm = torch.as_tensor(self.normalizer, dtype=dtype, device=lmd.device).detach()
        return meas.variance[0] * cov / norm

    def forward(self, x, y=None, normalize=True):
        if y is None:
            y = x
        if torch.is_grad_enabled() and hasattr(self.manifold, "_differentiable_spectrum") and \
                (self.measure.lengthscale.requires_grad or self.measure.variance.requires_grad):
            return self._differentiable(x, y, normalize)
        if self.training:
            self.manifold.generate_lb_eigenspaces(self.measure)
            if normalize:
                self.compute_normalizer()
        manifold = self.manifold
        x_, y_ = manifold.to_group(x), manifold.to_group(y)
        diff = manifold.pairwise_diff(x_, y_)
        f = manifold.lb_eigenspaces
        p_diff = manifold._features(diff, f, 1.0, "panel")
        p_id = manifold._features(manifold.to_group(manifold.id) if manifold.id.dim() == 3 else manifold.id, f, 1.0, "panel")
        cov = ops.gram(p_diff, p_id)[:, 0].view(x.size()[0], y.size()[0])
        if normalize:
            return self.measure.variance[0].detach() * cov / self.normalizer
        return cov


# north_star spellings
EigenSumKernel = EigenbasisSumKernel
RandomFourierKernel = RandomFourierFeatureKernel
