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
