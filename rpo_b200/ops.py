"""Tensor-level wrappers over the C ABI (device pointers of torch CUDA tensors, torch's current stream).

torch is plumbing here: it owns the device memory and the stream.  Every function launches exactly one
kernel of librpo_b200.so (plus one memset for the loop-test flags) and returns fresh tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _is_device_tensor(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda


def _dev(t: torch.Tensor, name: str, cols: Optional[int] = None) -> torch.Tensor:
    if not _is_device_tensor(t):
        raise _lib.RpoLibraryError("%s must be a CUDA tensor (rpo_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if t.dim() == 1:
        t = t.view(1, -1) if cols is None or t.numel() == cols else t.view(-1, cols)
    if cols is not None and t.shape[-1] != cols:
        raise ValueError("%s must have %d columns, got %s" % (name, cols, tuple(t.shape)))
    return t.contiguous()


def _rows(t: torch.Tensor, name: str, min_cols: int) -> Tuple[torch.Tensor, int]:
    """A (B, S>=min_cols) float32 CUDA tensor whose rows are contiguous; returns (tensor, leading dimension)."""
    if not _is_device_tensor(t):
        raise _lib.RpoLibraryError("%s must be a CUDA tensor (rpo_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if t.dim() == 1:
        t = t.view(1, -1)
    if t.shape[1] < min_cols:
        raise ValueError("%s needs at least %d columns, got %d" % (name, min_cols, t.shape[1]))
    if t.shape[0] > 1 and (t.stride(1) != 1 or t.stride(0) < t.shape[1]):
        t = t.contiguous()
    if t.shape[0] <= 1:
        t = t.contiguous()
        return t, t.shape[1]
    # The kernels bulk-copy whole row blocks of `ld` floats (TMA): the LAST row must own ld floats inside the allocation
    # too.  A column-offset view (e.g. s[:, 24:57] of a (B,57) tensor) ends `offset` floats before that: copy it.
    ld = t.stride(0)
    if ld != t.shape[1]:
        end = (t.storage_offset() + t.shape[0] * ld) * t.element_size()
        if end > t.untyped_storage().nbytes():
            t = t.contiguous()
            return t, t.shape[1]
    return t, ld


def _same_batch(B: int, **tensors) -> None:
    """Every row-wise argument must carry the batch size the kernel is launched with: the kernels index raw pointers."""
    for name, t in tensors.items():
        if t is not None and t.shape[0] != B:
            raise ValueError("%s has %d rows, expected %d" % (name, t.shape[0], B))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------- EVOPF
def evopf_complete_fwd(state, z, tol: float = 1e-5, max_iters: int = 50, return_status: bool = False, refine: bool = True):
    """refine=True: rpo_evopf_complete_fwd (float32 Newton + float64-residual chord step); refine=False: the plain float32
    Newton sequence of the fused projection kernel (rpo_evopf_complete_fwd_plain), bit-identical to evopf_project's y."""
    lib = _lib.load()
    state, ld = _rows(state, "state", 28)
    z = _dev(z, "z", 14)
    B = z.shape[0]
    _same_batch(B, state=state)
    y = torch.empty(B, 43, dtype=torch.float32, device=z.device)
    status = torch.empty(B, dtype=torch.int32, device=z.device) if return_status else None
    with torch.cuda.device(z.device):
        fn = lib.rpo_evopf_complete_fwd if refine else lib.rpo_evopf_complete_fwd_plain
        _lib.check(fn(state.data_ptr(), ld, z.data_ptr(), y.data_ptr(), B, tol, max_iters, _ptr(status), _stream(z)),
                   "rpo_evopf_complete_fwd")
    return (y, status) if return_status else y


def evopf_complete_bwd(y, dl_dy, need_dx: bool = True):
    lib = _lib.load()
    y = _dev(y, "y", 43)
    dl_dy = _dev(dl_dy, "dl_dy", 43)
    B = y.shape[0]
    _same_batch(B, dl_dy=dl_dy)
    dz = torch.empty(B, 14, dtype=torch.float32, device=y.device)
    dx = torch.empty(B, 28, dtype=torch.float32, device=y.device) if need_dx else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_complete_bwd(y.data_ptr(), dl_dy.data_ptr(), dz.data_ptr(), _ptr(dx), B, _stream(y)),
                   "rpo_evopf_complete_bwd")
    return dz, dx


def evopf_direction(state, y):
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    B = y.shape[0]
    _same_batch(B, state=state)
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_direction(state.data_ptr(), ld, y.data_ptr(), out.data_ptr(), B, _stream(y)),
                   "rpo_evopf_direction")
    return out


def evopf_correct(state, y, lr: float, momentum: float, corr_eps: float, n_steps: int, want_flags: bool = True):
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    B = y.shape[0]
    _same_batch(B, state=state)
    out = torch.empty_like(y)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=y.device) if want_flags else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_correct(state.data_ptr(), ld, y.data_ptr(), out.data_ptr(), B, lr, momentum, corr_eps,
                                         n_steps, _ptr(flags), _stream(y)), "rpo_evopf_correct")
    return out, flags


def evopf_project(state, z, lr: float, momentum: float, corr_eps: float, n_steps: int, tol: float = 1e-5,
                  max_iters: int = 50, want_y: bool = False, want_flags: bool = True, want_status: bool = False):
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    z = _dev(z, "z", 14)
    B = z.shape[0]
    _same_batch(B, state=state)
    y = torch.empty(B, 43, dtype=torch.float32, device=z.device) if want_y else None
    yc = torch.empty(B, 43, dtype=torch.float32, device=z.device)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=z.device) if want_flags else None
    status = torch.empty(B, dtype=torch.int32, device=z.device) if want_status else None
    with torch.cuda.device(z.device):
        _lib.check(lib.rpo_evopf_project(state.data_ptr(), ld, z.data_ptr(), _ptr(y), yc.data_ptr(), B, tol, max_iters,
                                         lr, momentum, corr_eps, n_steps, _ptr(flags), _ptr(status), _stream(z)),
                   "rpo_evopf_project")
    return yc, y, flags, status


def evopf_resid(state, y, want_eq: bool = True, want_ineq: bool = True):
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    B = y.shape[0]
    _same_batch(B, state=state)
    eq = torch.empty(B, 28, dtype=torch.float32, device=y.device) if want_eq else None
    ineq = torch.empty(B, 58, dtype=torch.float32, device=y.device) if want_ineq else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_resid(state.data_ptr(), ld, y.data_ptr(), _ptr(eq), _ptr(ineq), B, _stream(y)),
                   "rpo_evopf_resid")
    return eq, ineq


def evopf_resid_bwd(state, y, g_eq=None, g_ineq=None, ineq_mode: int = 0):
    """dL/dy of eq_resid / ineq_resid (ineq_mode 0) / ineq_dist (ineq_mode 1) given the upstream gradients (one launch)."""
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    B = y.shape[0]
    g_eq = None if g_eq is None else _dev(g_eq, "g_eq", 28)
    g_ineq = None if g_ineq is None else _dev(g_ineq, "g_ineq", 58)
    _same_batch(B, state=state, g_eq=g_eq, g_ineq=g_ineq)
    dy = torch.empty(B, 43, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_resid_bwd(state.data_ptr(), ld, y.data_ptr(), _ptr(g_eq), _ptr(g_ineq), ineq_mode,
                                           dy.data_ptr(), B, _stream(y)), "rpo_evopf_resid_bwd")
    return dy


def evopf_complete_bwd_penalty(state, y, dl_dy, w, row_scale=None, need_dx: bool = True):
    """GRG backward of L(y) + row_scale * w . relu(ineq_resid(state, y)) in one launch (the actor loss of rpo_ddpg.py:307-324)."""
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    dl_dy = _dev(dl_dy, "dl_dy", 43)
    w = _dev(w, "w", 58)
    B = y.shape[0]
    if row_scale is not None:
        row_scale = _dev(row_scale.reshape(-1, 1), "row_scale", 1)
    _same_batch(B, state=state, dl_dy=dl_dy, row_scale=row_scale)
    dz = torch.empty(B, 14, dtype=torch.float32, device=y.device)
    dx = torch.empty(B, 28, dtype=torch.float32, device=y.device) if need_dx else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_complete_bwd_penalty(state.data_ptr(), ld, y.data_ptr(), dl_dy.data_ptr(), w.data_ptr(),
                                                      _ptr(row_scale), dz.data_ptr(), _ptr(dx), B, _stream(y)),
                   "rpo_evopf_complete_bwd_penalty")
    return dz, dx


def _box14(lo, hi):
    lo, hi = _dev(lo, "box_lo").view(-1), _dev(hi, "box_hi").view(-1)
    if lo.numel() != 14 or hi.numel() != 14:
        raise ValueError("the EVOPF partial-action box has 14 bounds")
    return lo, hi


def evopf_project_squashed(state, u, lo, hi, box_mode: int, lr: float, momentum: float, corr_eps: float, n_steps: int,
                           tol: float = 1e-5, max_iters: int = 50, want_y: bool = False, want_flags: bool = True,
                           want_status: bool = False):
    """evopf_project(state, box(u), ...) in ONE launch: the action box (rpo_box_fwd, n_vol = 5) is applied inside the
    projection kernel; bit-identical to box_fwd -> evopf_project."""
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    u = _dev(u, "u", 14)
    lo, hi = _box14(lo, hi)
    B = u.shape[0]
    _same_batch(B, state=state)
    y = torch.empty(B, 43, dtype=torch.float32, device=u.device) if want_y else None
    yc = torch.empty(B, 43, dtype=torch.float32, device=u.device)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=u.device) if want_flags else None
    status = torch.empty(B, dtype=torch.int32, device=u.device) if want_status else None
    with torch.cuda.device(u.device):
        _lib.check(lib.rpo_evopf_project_squashed(state.data_ptr(), ld, u.data_ptr(), lo.data_ptr(), hi.data_ptr(), box_mode,
                                                  _ptr(y), yc.data_ptr(), B, tol, max_iters, lr, momentum, corr_eps, n_steps,
                                                  _ptr(flags), _ptr(status), _stream(u)), "rpo_evopf_project_squashed")
    return yc, y, flags, status


def evopf_complete_fwd_squashed(state, u, lo, hi, box_mode: int, tol: float = 1e-5, max_iters: int = 50):
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    u = _dev(u, "u", 14)
    lo, hi = _box14(lo, hi)
    B = u.shape[0]
    _same_batch(B, state=state)
    y = torch.empty(B, 43, dtype=torch.float32, device=u.device)
    with torch.cuda.device(u.device):
        _lib.check(lib.rpo_evopf_complete_fwd_squashed(state.data_ptr(), ld, u.data_ptr(), lo.data_ptr(), hi.data_ptr(), box_mode,
                                                       y.data_ptr(), B, tol, max_iters, None, _stream(u)),
                   "rpo_evopf_complete_fwd_squashed")
    return y


def evopf_complete_bwd_squashed(state, y, dl_dy, u, lo, hi, box_mode: int, w=None, row_scale=None, need_dx: bool = False):
    """dL/du (and dL/dx) of the squashed completion, optionally with the nju . ineq_dist penalty folded in (one launch)."""
    lib = _lib.load()
    state, ld = _rows(state, "state", 33)
    y = _dev(y, "y", 43)
    dl_dy = _dev(dl_dy, "dl_dy", 43)
    u = _dev(u, "u", 14)
    lo, hi = _box14(lo, hi)
    w = None if w is None else _dev(w, "w", 58)
    B = y.shape[0]
    if row_scale is not None:
        row_scale = _dev(row_scale.reshape(-1, 1), "row_scale", 1)
    _same_batch(B, state=state, dl_dy=dl_dy, u=u, row_scale=row_scale)
    du = torch.empty(B, 14, dtype=torch.float32, device=y.device)
    dx = torch.empty(B, 28, dtype=torch.float32, device=y.device) if need_dx else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_complete_bwd_squashed(state.data_ptr(), ld, y.data_ptr(), dl_dy.data_ptr(), u.data_ptr(),
                                                       lo.data_ptr(), hi.data_ptr(), box_mode, _ptr(w), _ptr(row_scale),
                                                       du.data_ptr(), _ptr(dx), B, _stream(y)), "rpo_evopf_complete_bwd_squashed")
    return du, dx


def evopf_step(state, y, next_exo, want_viol: bool = True):
    lib = _lib.load()
    state = _dev(state, "state", 57)
    y = _dev(y, "y", 43)
    next_exo = _dev(next_exo, "next_exo", 52)
    B = y.shape[0]
    _same_batch(B, state=state, next_exo=next_exo)
    ns = torch.empty(B, 57, dtype=torch.float32, device=y.device)
    rew = torch.empty(B, dtype=torch.float32, device=y.device)
    eqv = torch.empty(B, 28, dtype=torch.float32, device=y.device) if want_viol else None
    iqv = torch.empty(B, 58, dtype=torch.float32, device=y.device) if want_viol else None
    with torch.cuda.device(y.device):
        _lib.check(lib.rpo_evopf_step(state.data_ptr(), y.data_ptr(), next_exo.data_ptr(), ns.data_ptr(), rew.data_ptr(),
                                      _ptr(eqv), _ptr(iqv), B, _stream(y)), "rpo_evopf_step")
    return ns, rew, eqv, iqv


# ------------------------------------------------------------------------------------------- cart
def cart_complete_fwd(z, partial_idx: int):
    lib = _lib.load()
    z = _dev(z, "z", 1)
    B = z.shape[0]
    a = torch.empty(B, 2, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(lib.rpo_cart_complete_fwd(z.data_ptr(), a.data_ptr(), B, partial_idx, _stream(z)), "rpo_cart_complete_fwd")
    return a


def cart_complete_bwd(dl_da, partial_idx: int):
    lib = _lib.load()
    dl_da = _dev(dl_da, "dl_da", 2)
    B = dl_da.shape[0]
    dz = torch.empty(B, 1, dtype=torch.float32, device=dl_da.device)
    with torch.cuda.device(dl_da.device):
        _lib.check(lib.rpo_cart_complete_bwd(dl_da.data_ptr(), dz.data_ptr(), B, partial_idx, _stream(dl_da)),
                   "rpo_cart_complete_bwd")
    return dz


def cart_direction(a, partial_idx: int):
    lib = _lib.load()
    a = _dev(a, "a", 2)
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_cart_direction(a.data_ptr(), out.data_ptr(), a.shape[0], partial_idx, _stream(a)), "rpo_cart_direction")
    return out


def cart_correct(a, partial_idx: int, lr, momentum, corr_eps, n_steps, want_flags=True):
    lib = _lib.load()
    a = _dev(a, "a", 2)
    out = torch.empty_like(a)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=a.device) if want_flags else None
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_cart_correct(a.data_ptr(), out.data_ptr(), a.shape[0], partial_idx, lr, momentum, corr_eps,
                                        n_steps, _ptr(flags), _stream(a)), "rpo_cart_correct")
    return out, flags


def cart_project(z, partial_idx: int, lr, momentum, corr_eps, n_steps, want_a=False, want_flags=True):
    lib = _lib.load()
    z = _dev(z, "z", 1)
    B = z.shape[0]
    a = torch.empty(B, 2, dtype=torch.float32, device=z.device) if want_a else None
    ac = torch.empty(B, 2, dtype=torch.float32, device=z.device)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=z.device) if want_flags else None
    with torch.cuda.device(z.device):
        _lib.check(lib.rpo_cart_project(z.data_ptr(), _ptr(a), ac.data_ptr(), B, partial_idx, lr, momentum, corr_eps,
                                        n_steps, _ptr(flags), _stream(z)), "rpo_cart_project")
    return ac, a, flags


def cart_resid(a, want_eq=True, want_ineq=True):
    lib = _lib.load()
    a = _dev(a, "a", 2)
    B = a.shape[0]
    eq = torch.empty(B, 1, dtype=torch.float32, device=a.device) if want_eq else None
    ineq = torch.empty(B, 6, dtype=torch.float32, device=a.device) if want_ineq else None
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_cart_resid(a.data_ptr(), _ptr(eq), _ptr(ineq), B, _stream(a)), "rpo_cart_resid")
    return eq, ineq


def cart_resid_bwd(a, g_eq=None, g_ineq=None, ineq_mode: int = 0):
    lib = _lib.load()
    a = _dev(a, "a", 2)
    B = a.shape[0]
    g_eq = None if g_eq is None else _dev(g_eq, "g_eq", 1)
    g_ineq = None if g_ineq is None else _dev(g_ineq, "g_ineq", 6)
    _same_batch(B, g_eq=g_eq, g_ineq=g_ineq)
    da = torch.empty(B, 2, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_cart_resid_bwd(a.data_ptr(), _ptr(g_eq), _ptr(g_ineq), ineq_mode, da.data_ptr(), B, _stream(a)),
                   "rpo_cart_resid_bwd")
    return da


def cart_step(state, a, want_viol=True):
    lib = _lib.load()
    state = _dev(state, "state", 6)
    a = _dev(a, "a", 2)
    B = a.shape[0]
    _same_batch(B, state=state)
    ns = torch.empty(B, 6, dtype=torch.float32, device=a.device)
    done = torch.empty(B, dtype=torch.int32, device=a.device)
    eqv = torch.empty(B, 1, dtype=torch.float32, device=a.device) if want_viol else None
    iqv = torch.empty(B, 6, dtype=torch.float32, device=a.device) if want_viol else None
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_cart_step(state.data_ptr(), a.data_ptr(), ns.data_ptr(), done.data_ptr(), _ptr(eqv), _ptr(iqv), B,
                                     _stream(a)), "rpo_cart_step")
    return ns, done, eqv, iqv


# ------------------------------------------------------------------------------------------- pendulum
def pen_complete_fwd(obs, z):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    z = _dev(z, "z", 1)
    B = z.shape[0]
    _same_batch(B, obs=obs)
    a = torch.empty(B, 2, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(lib.rpo_pen_complete_fwd(obs.data_ptr(), z.data_ptr(), a.data_ptr(), B, _stream(z)), "rpo_pen_complete_fwd")
    return a


def pen_complete_bwd(obs, dl_da):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    dl_da = _dev(dl_da, "dl_da", 2)
    B = dl_da.shape[0]
    _same_batch(B, obs=obs)
    dz = torch.empty(B, 1, dtype=torch.float32, device=dl_da.device)
    with torch.cuda.device(dl_da.device):
        _lib.check(lib.rpo_pen_complete_bwd(obs.data_ptr(), dl_da.data_ptr(), dz.data_ptr(), B, _stream(dl_da)),
                   "rpo_pen_complete_bwd")
    return dz


def pen_direction(obs, a):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    a = _dev(a, "a", 2)
    _same_batch(a.shape[0], obs=obs)
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_pen_direction(obs.data_ptr(), a.data_ptr(), out.data_ptr(), a.shape[0], _stream(a)), "rpo_pen_direction")
    return out


def pen_correct(obs, a, lr, momentum, corr_eps, n_steps, want_flags=True):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    a = _dev(a, "a", 2)
    _same_batch(a.shape[0], obs=obs)
    out = torch.empty_like(a)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=a.device) if want_flags else None
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_pen_correct(obs.data_ptr(), a.data_ptr(), out.data_ptr(), a.shape[0], lr, momentum, corr_eps,
                                       n_steps, _ptr(flags), _stream(a)), "rpo_pen_correct")
    return out, flags


def pen_project(obs, z, lr, momentum, corr_eps, n_steps, want_a=False, want_flags=True):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    z = _dev(z, "z", 1)
    B = z.shape[0]
    _same_batch(B, obs=obs)
    a = torch.empty(B, 2, dtype=torch.float32, device=z.device) if want_a else None
    ac = torch.empty(B, 2, dtype=torch.float32, device=z.device)
    flags = torch.empty(max(n_steps, 1), dtype=torch.int32, device=z.device) if want_flags else None
    with torch.cuda.device(z.device):
        _lib.check(lib.rpo_pen_project(obs.data_ptr(), z.data_ptr(), _ptr(a), ac.data_ptr(), B, lr, momentum, corr_eps,
                                       n_steps, _ptr(flags), _stream(z)), "rpo_pen_project")
    return ac, a, flags


def pen_resid(obs, a, want_eq=True, want_ineq=True):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    a = _dev(a, "a", 2)
    B = a.shape[0]
    _same_batch(B, obs=obs)
    eq = torch.empty(B, 1, dtype=torch.float32, device=a.device) if want_eq else None
    ineq = torch.empty(B, 1, dtype=torch.float32, device=a.device) if want_ineq else None
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_pen_resid(obs.data_ptr(), a.data_ptr(), _ptr(eq), _ptr(ineq), B, _stream(a)), "rpo_pen_resid")
    return eq, ineq


def pen_resid_bwd(obs, a, g_eq=None, g_ineq=None, ineq_mode: int = 0):
    lib = _lib.load()
    obs = _dev(obs, "obs", 5)
    a = _dev(a, "a", 2)
    B = a.shape[0]
    g_eq = None if g_eq is None else _dev(g_eq, "g_eq", 1)
    g_ineq = None if g_ineq is None else _dev(g_ineq, "g_ineq", 1)
    _same_batch(B, obs=obs, g_eq=g_eq, g_ineq=g_ineq)
    da = torch.empty(B, 2, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.rpo_pen_resid_bwd(obs.data_ptr(), a.data_ptr(), _ptr(g_eq), _ptr(g_ineq), ineq_mode, da.data_ptr(), B,
                                         _stream(a)), "rpo_pen_resid_bwd")
    return da


def pen_step(st, a, want_viol=True):
    lib = _lib.load()
    st = _dev(st, "st", 4)
    a = _dev(a, "a", 2)
    B = a.shape[0]
    _same_batch(B, st=st)
    dev = a.device
    nst = torch.empty(B, 4, dtype=torch.float32, device=dev)
    nobs = torch.empty(B, 5, dtype=torch.float32, device=dev)
    rew = torch.empty(B, dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.int32, device=dev)
    eqv = torch.empty(B, 1, dtype=torch.float32, device=dev) if want_viol else None
    iqv = torch.empty(B, 1, dtype=torch.float32, device=dev) if want_viol else None
    with torch.cuda.device(dev):
        _lib.check(lib.rpo_pen_step(st.data_ptr(), a.data_ptr(), nst.data_ptr(), nobs.data_ptr(), rew.data_ptr(),
                                    done.data_ptr(), _ptr(eqv), _ptr(iqv), B, _stream(a)), "rpo_pen_step")
    return nst, nobs, rew, done, eqv, iqv


# ------------------------------------------------------------------------------------------- action box (f-1)
BOX_AFFINE, BOX_TANH, BOX_CLIP, BOX_AFFINE01, BOX_SCALE, BOX_LOW, BOX_HIGH = range(7)


def _box_args(x, lo, hi, state, n_vol, mode, B):
    lo = _dev(lo, "lo").view(-1)
    hi = _dev(hi, "hi").view(-1)
    W = lo.numel()
    if hi.numel() != W:
        raise ValueError("lo and hi disagree on the box width")
    if mode <= BOX_AFFINE01:
        x = _dev(x, "x", W)
        B = x.shape[0]
    elif B is None:
        if state is None:
            raise ValueError("modes scale / low / high need `state` or `B` for the batch size")
        B = state.shape[0] if state.dim() > 1 else 1
    soc_ptr, ld = None, 0
    if n_vol:
        if state is None:
            raise ValueError("a volatile box needs the state (EVOPFEnv.update)")
        state, ld = _rows(state, "state", 33)
        if state.shape[0] != B:
            raise ValueError("state and x disagree on the batch size")
        soc_ptr = state.data_ptr() + 28 * 4
    return x, lo, hi, W, state, soc_ptr, ld, B


def box_fwd(x, lo, hi, mode: int = BOX_AFFINE, state=None, n_vol: int = 0, out=None, B=None):
    """One launch of rpo_box_fwd: x (B,W) -> (B,W) inside the box [lo, hi] whose last n_vol columns follow state[:, 28:33]."""
    lib = _lib.load()
    x, lo, hi, W, state, soc_ptr, ld, B = _box_args(x, lo, hi, state, n_vol, mode, B)
    if out is None:
        out = torch.empty(B, W, dtype=torch.float32, device=lo.device)
    with torch.cuda.device(lo.device):
        _lib.check(lib.rpo_box_fwd(_ptr(x) if mode <= BOX_AFFINE01 else None, lo.data_ptr(), hi.data_ptr(), W, soc_ptr, ld, n_vol,
                                   out.data_ptr(), mode, B, _stream(lo)), "rpo_box_fwd")
    return out


def box_bwd(x, grad_out, lo, hi, mode: int = BOX_AFFINE, state=None, n_vol: int = 0):
    lib = _lib.load()
    x, lo, hi, W, state, soc_ptr, ld, B = _box_args(x, lo, hi, state, n_vol, mode, None)
    g = _dev(grad_out, "grad_out", W)
    _same_batch(B, grad_out=g)
    gx = torch.empty(B, W, dtype=torch.float32, device=lo.device)
    with torch.cuda.device(lo.device):
        _lib.check(lib.rpo_box_bwd(x.data_ptr(), g.data_ptr(), lo.data_ptr(), hi.data_ptr(), W, soc_ptr, ld, n_vol, gx.data_ptr(),
                                   mode, B, _stream(lo)), "rpo_box_bwd")
    return gx


def fp32_peak_probe(device, blocks: int, threads: int, iters: int):
    """Launch the FMA micro-kernel once; returns flops executed (time it with CUDA events around the call)."""
    lib = _lib.load()
    out = torch.empty(blocks * threads, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.rpo_fp32_peak_probe(out.data_ptr(), blocks, threads, iters, torch.cuda.current_stream(device).cuda_stream),
                   "rpo_fp32_peak_probe")
    return 2.0 * 32 * iters * blocks * threads, out
