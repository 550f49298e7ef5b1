import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_err(got, ref):
    """max |got-ref| / |ref| per entry (the north-star tolerance is <= 1e-10 on this)."""
    import torch
    return (torch.abs(got - ref) / torch.abs(ref)).max().item()


def scaled_err(got, ref):
    """max |got-ref| / max|ref| (used where entries cross zero)."""
    import torch
    return (torch.abs(got - ref).max() / torch.abs(ref).max()).item()


def dot_err(got, ref, abs_gram):
    """max over entries of |got-ref| / (|Phi_x| |Phi_y|^T)_ij: the componentwise bound of an inner product.  Used for Gram
    matrices of random-feature maps, whose entries are Monte-Carlo sums that cross zero (a bound relative to |ref_ij| is not
    well-posed there: two BLAS summation orders already differ by 1e-16 sqrt(K) |Phi_x||Phi_y|^T)."""
    import torch
    return (torch.abs(got - ref) / abs_gram).max().item()


def rel_err_conditioned(got, ref, abs_gram, kappa=1e3):
    """max |got-ref| / |ref| over the entries whose inner product loses fewer than log10(kappa) digits to cancellation
    (|ref_ij| >= abs_gram_ij / kappa); returns (error, fraction of entries covered)."""
    import torch
    mask = torch.abs(ref) * kappa >= abs_gram
    if not mask.any():
        return 0.0, 0.0
    return (torch.abs(got - ref)[mask] / torch.abs(ref)[mask]).max().item(), mask.double().mean().item()


def record(test, **values):
    """Append observed parity errors to gpurun_out/parity_errors.jsonl (diagnostics that travel back from the GPU box)."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as fh:
            fh.write(json.dumps(dict(test=test, **values)) + "\n")
    except OSError:
        pass


def cov_err(got, ref, floor=1e-3):
    """Per-entry error of a covariance matrix: |got-ref| / max(|ref_ij|, floor * max|ref|).  Entries within three orders of
    magnitude of the largest one are held to the north-star bound RELATIVE TO THEMSELVES; smaller ones (truncated spectral sums
    and Monte-Carlo Grams oscillate through zero far from the diagonal) to floor * that bound in units of the largest entry,
    i.e. 1e-13 of scale at TOL = 1e-10."""
    import torch
    denom = torch.clamp(torch.abs(ref), min=floor * torch.abs(ref).max().item())
    return (torch.abs(got - ref) / denom).max().item()
