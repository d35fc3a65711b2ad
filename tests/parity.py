"""Shared parity metrics (tolerances of BASELINE.json north_star, stated once here)."""
import torch

Y_RTOL = 1e-4        # nonbasic / corrected actions: per-instance max-norm relative error
GRAD_RTOL = 1e-3     # reduced gradients
EQ_TOL = 1e-5        # reference Newton tolerance (evopf.py:786)
OUTLIER_FRAC = 1e-4  # instances within f32 round-off of an inequality bound may flip their mask


def rel_err(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per-instance max-norm relative error of a against the reference b (both (B, n), CPU)."""
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return (a - b).abs().max(dim=1).values / b.abs().max(dim=1).values.clamp_min(1e-30)


MARGIN_TOL = 1e-4    # an instance is "mask-stable" if no inequality residual comes closer than this to zero


def assert_close_or_marginal(err: torch.Tensor, margin: torch.Tensor, tol: float, max_outlier_frac: float = 5e-3):
    """Every instance whose inequality margins stay above MARGIN_TOL along the reference trajectory must match to
    `tol`; the others may flip a mask but must stay a small fraction."""
    bad = err > tol
    assert not bool((bad & (margin.cpu().double() >= MARGIN_TOL)).any()), \
        "mask-stable instances exceed tolerance: max err %g" % float(err[bad & (margin.cpu().double() >= MARGIN_TOL)].max())
    assert float(bad.double().mean()) <= max_outlier_frac, "outlier fraction %g" % float(bad.double().mean())
