"""TEST INFRASTRUCTURE ONLY: CPU stand-ins for `rpo_b200.ops` backed by the host simulator (tests/hostsim), i.e. by
the very same per-instance code the CUDA kernels run.  `install(monkeypatch)` swaps them into `rpo_b200.ops` so that
the drop-in environments and drivers can be exercised end to end — together with the UNMODIFIED reference classes —
in a container without a GPU.  The product never imports this module; without it the ops raise on CPU tensors."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

HS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", HS])
        _lib = ctypes.CDLL(os.path.join(HS, "libhostsim_f32.so"))
    return _lib


def _P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _np(t, cols=None):
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(1, -1) if cols is None or a.size == cols else a.reshape(-1, cols)
    return a


def _t(a):
    return torch.from_numpy(a)


f = ctypes.c_float


def evopf_complete_fwd(state, z, tol=1e-5, max_iters=50, return_status=False, refine=True):
    s, zz = _np(state), _np(z, 14)
    B = zz.shape[0]
    y = np.zeros((B, 43), np.float32)
    st = np.zeros(B, np.int32)
    lib().hs_evopf_complete(_P(s), s.shape[1], _P(zz), _P(y), B, f(tol), max_iters, _P(st), int(bool(refine)))
    return (_t(y), _t(st)) if return_status else _t(y)


def evopf_complete_bwd(y, dl_dy, need_dx=True):
    yy, g = _np(y, 43), _np(dl_dy, 43)
    B = yy.shape[0]
    dz = np.zeros((B, 14), np.float32)
    dx = np.zeros((B, 28), np.float32) if need_dx else None
    lib().hs_evopf_backward(_P(yy), _P(g), _P(dz), _P(dx), B)
    return _t(dz), (_t(dx) if need_dx else None)


def evopf_direction(state, y):
    s, yy = _np(state), _np(y, 43)
    B = yy.shape[0]
    out, em, im = np.zeros((B, 43), np.float32), np.zeros(B, np.float32), np.zeros(B, np.float32)
    lib().hs_evopf_direction(_P(s), s.shape[1], _P(yy), _P(out), _P(em), _P(im), B)
    return _t(out)


def evopf_correct(state, y, lr, momentum, corr_eps, n_steps, want_flags=True):
    s, yy = _np(state), _np(y, 43).copy()
    flags = np.zeros(max(n_steps, 1), np.int32)
    lib().hs_evopf_correct(_P(s), s.shape[1], _P(yy), yy.shape[0], f(lr), f(momentum), f(corr_eps), n_steps, _P(flags))
    return _t(yy), (_t(flags) if want_flags else None)


def evopf_project(state, z, lr, momentum, corr_eps, n_steps, tol=1e-5, max_iters=50, want_y=False, want_flags=True,
                  want_status=False):
    s, zz = _np(state), _np(z, 14)
    B = zz.shape[0]
    if n_steps == 0:
        y = evopf_complete_fwd(state, z, tol, max_iters, refine=False)
        return y.clone(), (y if want_y else None), (_t(np.zeros(1, np.int32)) if want_flags else None), None
    y, yc = np.zeros((B, 43), np.float32), np.zeros((B, 43), np.float32)
    flags, st = np.zeros(max(n_steps, 1), np.int32), np.zeros(B, np.int32)
    lib().hs_evopf_project(_P(s), s.shape[1], _P(zz), _P(y), _P(yc), B, f(tol), max_iters, f(lr), f(momentum), f(corr_eps),
                           n_steps, _P(flags), _P(st))
    return _t(yc), (_t(y) if want_y else None), (_t(flags) if want_flags else None), (_t(st) if want_status else None)


def evopf_resid(state, y, want_eq=True, want_ineq=True):
    s, yy = _np(state), _np(y, 43)
    B = yy.shape[0]
    eq, iq = np.zeros((B, 28), np.float32), np.zeros((B, 58), np.float32)
    lib().hs_evopf_resid(_P(s), s.shape[1], _P(yy), _P(eq), _P(iq), B)
    return (_t(eq) if want_eq else None), (_t(iq) if want_ineq else None)


def evopf_step(state, y, next_exo, want_viol=True):
    s, yy, ex = _np(state, 57), _np(y, 43), _np(next_exo, 52)
    B = yy.shape[0]
    ns, rew = np.zeros((B, 57), np.float32), np.zeros(B, np.float32)
    eqv = np.zeros((B, 28), np.float32) if want_viol else None
    iqv = np.zeros((B, 58), np.float32) if want_viol else None
    lib().hs_evopf_step(_P(s), _P(yy), _P(ex), _P(ns), _P(rew), _P(eqv), _P(iqv), B)
    return _t(ns), _t(rew), (_t(eqv) if want_viol else None), (_t(iqv) if want_viol else None)


def cart_complete_fwd(z, partial_idx):
    zz = _np(z, 1)
    a = np.zeros((zz.shape[0], 2), np.float32)
    lib().hs_cart_complete(_P(zz), _P(a), zz.shape[0], partial_idx)
    return _t(a)


def cart_complete_bwd(dl_da, partial_idx):
    d = _np(dl_da, 2)
    dz = np.zeros((d.shape[0], 1), np.float32)
    lib().hs_cart_complete_bwd(_P(d), _P(dz), d.shape[0], partial_idx)
    return _t(dz)


def cart_direction(a, partial_idx):
    aa = _np(a, 2)
    out = np.zeros_like(aa)
    lib().hs_cart_direction(_P(aa), _P(out), aa.shape[0], partial_idx)
    return _t(out)


def cart_correct(a, partial_idx, lr, momentum, corr_eps, n_steps, want_flags=True):
    aa = _np(a, 2).copy()
    flags = np.zeros(max(n_steps, 1), np.int32)
    lib().hs_cart_correct(_P(aa), aa.shape[0], partial_idx, f(lr), f(momentum), f(corr_eps), n_steps, _P(flags))
    return _t(aa), (_t(flags) if want_flags else None)


def cart_project(z, partial_idx, lr, momentum, corr_eps, n_steps, want_a=False, want_flags=True):
    a = cart_complete_fwd(z, partial_idx)
    ac, flags = cart_correct(a, partial_idx, lr, momentum, corr_eps, n_steps, want_flags)
    return ac, (a if want_a else None), flags


def cart_resid(a, want_eq=True, want_ineq=True):
    aa = _np(a, 2)
    eq, iq = np.zeros((aa.shape[0], 1), np.float32), np.zeros((aa.shape[0], 6), np.float32)
    lib().hs_cart_resid(_P(aa), _P(eq), _P(iq), aa.shape[0])
    return (_t(eq) if want_eq else None), (_t(iq) if want_ineq else None)


def cart_step(state, a, want_viol=True):
    s, aa = _np(state, 6), _np(a, 2)
    B = aa.shape[0]
    ns, done = np.zeros((B, 6), np.float32), np.zeros(B, np.int32)
    lib().hs_cart_step(_P(s), _P(aa), _P(ns), _P(done), B)
    eq, iq = cart_resid(_t(aa))
    return _t(ns), _t(done), (eq if want_viol else None), (torch.clamp(iq, 0) if want_viol else None)


def pen_complete_fwd(obs, z):
    o, zz = _np(obs, 5), _np(z, 1)
    a = np.zeros((zz.shape[0], 2), np.float32)
    lib().hs_pen_complete(_P(o), _P(zz), _P(a), zz.shape[0])
    return _t(a)


def pen_complete_bwd(obs, dl_da):
    o, d = _np(obs, 5), _np(dl_da, 2)
    dz = np.zeros((d.shape[0], 1), np.float32)
    lib().hs_pen_complete_bwd(_P(o), _P(d), _P(dz), d.shape[0])
    return _t(dz)


def pen_direction(obs, a):
    o, aa = _np(obs, 5), _np(a, 2)
    out = np.zeros_like(aa)
    lib().hs_pen_direction(_P(o), _P(aa), _P(out), aa.shape[0])
    return _t(out)


def pen_correct(obs, a, lr, momentum, corr_eps, n_steps, want_flags=True):
    o, aa = _np(obs, 5), _np(a, 2).copy()
    flags = np.zeros(max(n_steps, 1), np.int32)
    lib().hs_pen_correct(_P(o), _P(aa), aa.shape[0], f(lr), f(momentum), f(corr_eps), n_steps, _P(flags))
    return _t(aa), (_t(flags) if want_flags else None)


def pen_project(obs, z, lr, momentum, corr_eps, n_steps, want_a=False, want_flags=True):
    a = pen_complete_fwd(obs, z)
    ac, flags = pen_correct(obs, a, lr, momentum, corr_eps, n_steps, want_flags)
    return ac, (a if want_a else None), flags


def pen_resid(obs, a, want_eq=True, want_ineq=True):
    o, aa = _np(obs, 5), _np(a, 2)
    eq, iq = np.zeros((aa.shape[0], 1), np.float32), np.zeros((aa.shape[0], 1), np.float32)
    lib().hs_pen_resid(_P(o), _P(aa), _P(eq), _P(iq), aa.shape[0])
    return (_t(eq) if want_eq else None), (_t(iq) if want_ineq else None)


def pen_step(st, a, want_viol=True):
    s, aa = _np(st, 4), _np(a, 2)
    B = aa.shape[0]
    nst, nobs, rew, done = np.zeros((B, 4), np.float32), np.zeros((B, 5), np.float32), np.zeros(B, np.float32), np.zeros(B, np.int32)
    lib().hs_pen_step(_P(s), _P(aa), _P(nst), _P(nobs), _P(rew), _P(done), B)
    obs = np.stack([np.cos(s[:, 0].astype(np.float64)), np.sin(s[:, 0].astype(np.float64)), s[:, 1], s[:, 2], s[:, 3]], 1).astype(np.float32)
    eq, iq = pen_resid(_t(obs), _t(aa))
    return _t(nst), _t(nobs), _t(rew), _t(done), (eq if want_viol else None), (torch.clamp(iq, 0) if want_viol else None)


def evopf_resid_bwd(state, y, g_eq=None, g_ineq=None, ineq_mode=0):
    s, yy = _np(state), _np(y, 43)
    B = yy.shape[0]
    ge = None if g_eq is None else _np(g_eq, 28)
    gi = None if g_ineq is None else _np(g_ineq, 58)
    dy = np.zeros((B, 43), np.float32)
    lib().hs_evopf_resid_bwd(_P(s), s.shape[1], _P(yy), _P(ge), _P(gi), ineq_mode, _P(dy), B)
    return _t(dy)


def evopf_complete_bwd_penalty(state, y, dl_dy, w, row_scale=None, need_dx=True):
    s, yy, g = _np(state), _np(y, 43), _np(dl_dy, 43)
    B = yy.shape[0]
    ww = _np(w, 58).reshape(-1)
    rs = None if row_scale is None else np.ascontiguousarray(_np(row_scale).reshape(-1))
    dz = np.zeros((B, 14), np.float32)
    dx = np.zeros((B, 28), np.float32) if need_dx else None
    lib().hs_evopf_backward_penalty(_P(s), s.shape[1], _P(yy), _P(g), _P(ww), _P(rs), _P(dz), _P(dx), B)
    return _t(dz), (_t(dx) if need_dx else None)


def cart_resid_bwd(a, g_eq=None, g_ineq=None, ineq_mode=0):
    aa = _np(a, 2)
    ge = None if g_eq is None else _np(g_eq, 1)
    gi = None if g_ineq is None else _np(g_ineq, 6)
    da = np.zeros((aa.shape[0], 2), np.float32)
    lib().hs_cart_resid_bwd(_P(aa), _P(ge), _P(gi), ineq_mode, _P(da), aa.shape[0])
    return _t(da)


def pen_resid_bwd(obs, a, g_eq=None, g_ineq=None, ineq_mode=0):
    o, aa = _np(obs, 5), _np(a, 2)
    ge = None if g_eq is None else _np(g_eq, 1)
    gi = None if g_ineq is None else _np(g_ineq, 1)
    da = np.zeros((aa.shape[0], 2), np.float32)
    lib().hs_pen_resid_bwd(_P(o), _P(aa), _P(ge), _P(gi), ineq_mode, _P(da), aa.shape[0])
    return _t(da)


def _box(bwd, x, g, lo, hi, mode, state, n_vol, B):
    lo_, hi_ = _np(lo).reshape(-1), _np(hi).reshape(-1)
    W = lo_.size
    xx = None if x is None else _np(x, W)
    if xx is not None:
        B = xx.shape[0]
    elif B is None:
        B = _np(state).shape[0]
    soc, ld = None, 0
    if n_vol:
        s = _np(state)
        if s.shape[0] != B:
            s = np.ascontiguousarray(np.broadcast_to(s, (B, s.shape[1])))
        soc, ld = np.ascontiguousarray(s[:, 28:33]), 5
    out = np.zeros((B, W), np.float32)
    if bwd:
        gg = _np(g, W)
        lib().hs_box_bwd(_P(xx), _P(gg), _P(lo_), _P(hi_), W, _P(soc), ld, n_vol, _P(out), mode, B)
    else:
        lib().hs_box_fwd(_P(xx), _P(lo_), _P(hi_), W, _P(soc), ld, n_vol, _P(out), mode, B)
    return _t(out)


def box_fwd(x, lo, hi, mode=0, state=None, n_vol=0, out=None, B=None):
    return _box(False, x, None, lo, hi, mode, state, n_vol, B)


def box_bwd(x, grad_out, lo, hi, mode=0, state=None, n_vol=0):
    return _box(True, x, grad_out, lo, hi, mode, state, n_vol, None)


def evopf_project_squashed(state, u, lo, hi, box_mode, lr, momentum, corr_eps, n_steps, tol=1e-5, max_iters=50, want_y=False,
                           want_flags=True, want_status=False):
    z = box_fwd(u, lo, hi, box_mode, state, 5)
    return evopf_project(state, z, lr, momentum, corr_eps, n_steps, tol, max_iters, want_y, want_flags, want_status)


def evopf_complete_fwd_squashed(state, u, lo, hi, box_mode, tol=1e-5, max_iters=50):
    return evopf_complete_fwd(state, box_fwd(u, lo, hi, box_mode, state, 5), tol, max_iters)


def evopf_complete_bwd_squashed(state, y, dl_dy, u, lo, hi, box_mode, w=None, row_scale=None, need_dx=False):
    if w is None:
        dz, dx = evopf_complete_bwd(y, dl_dy, need_dx)
    else:
        dz, dx = evopf_complete_bwd_penalty(state, y, dl_dy, w, row_scale, need_dx)
    return box_bwd(u, dz, lo, hi, box_mode, state, 5), dx


NAMES = [n for n in dir() if n.startswith(("evopf_", "cart_", "pen_", "box_"))]


def install(monkeypatch):
    """Swap every op of rpo_b200.ops for its host-simulator twin (test scope only, undone by monkeypatch)."""
    import rpo_b200.ops as ops
    g = globals()
    for n in NAMES:
        assert hasattr(ops, n), n
        monkeypatch.setattr(ops, n, g[n])
