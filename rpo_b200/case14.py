"""IEEE 14-bus test case (public MATPOWER/PYPOWER ``case14``) and the bus-admittance matrix.

The reference obtains these constants from the third-party package ``pypower`` (un-pinned, absent
from /root/reference): ``case14()`` at rpo/env/electrical_grid/evopf.py:211, ``makeYbus`` at
evopf.py:263 and the ``idx_bus`` / ``idx_gen`` column constants at evopf.py:222-258,299-302.
pypower contributes *constants only* to the hot path; they are restated here from the published
case file and pinned by two known-answer tests (tests/test_case14_kat.py):

* KAT-1: Y-bus checksums (sum, Frobenius norms, nnz, a few entries),
* KAT-2: a Newton power flow from the case set-points reproduces the solved case that is printed
  in the case file itself (|dVm| < 2e-3 pu, slack Pg = 232.39 MW).
"""
from __future__ import annotations

import numpy as np

# --- column indices (pypower.idx_bus / pypower.idx_gen) -------------------------------------
BUS_I, BUS_TYPE, PD, QD, GS, BS, BUS_AREA, VM, VA, BASE_KV, ZONE, VMAX, VMIN = range(13)
GEN_BUS, PG, QG, QMAX, QMIN, VG, MBASE, GEN_STATUS, PMAX, PMIN = range(10)
F_BUS, T_BUS, BR_R, BR_X, BR_B, RATE_A, RATE_B, RATE_C, TAP, SHIFT, BR_STATUS, ANGMIN, ANGMAX = range(13)

_BUS = [
    # i type  Pd    Qd   Gs  Bs  area Vm     Va      baseKV zone Vmax  Vmin
    [1, 3, 0.0, 0.0, 0, 0, 1, 1.060, 0.00, 0, 1, 1.06, 0.94],
    [2, 2, 21.7, 12.7, 0, 0, 1, 1.045, -4.98, 0, 1, 1.06, 0.94],
    [3, 2, 94.2, 19.0, 0, 0, 1, 1.010, -12.72, 0, 1, 1.06, 0.94],
    [4, 1, 47.8, -3.9, 0, 0, 1, 1.019, -10.33, 0, 1, 1.06, 0.94],
    [5, 1, 7.6, 1.6, 0, 0, 1, 1.020, -8.78, 0, 1, 1.06, 0.94],
    [6, 2, 11.2, 7.5, 0, 0, 1, 1.070, -14.22, 0, 1, 1.06, 0.94],
    [7, 1, 0.0, 0.0, 0, 0, 1, 1.062, -13.37, 0, 1, 1.06, 0.94],
    [8, 2, 0.0, 0.0, 0, 0, 1, 1.090, -13.36, 0, 1, 1.06, 0.94],
    [9, 1, 29.5, 16.6, 0, 19, 1, 1.056, -14.94, 0, 1, 1.06, 0.94],
    [10, 1, 9.0, 5.8, 0, 0, 1, 1.051, -15.10, 0, 1, 1.06, 0.94],
    [11, 1, 3.5, 1.8, 0, 0, 1, 1.057, -14.79, 0, 1, 1.06, 0.94],
    [12, 1, 6.1, 1.6, 0, 0, 1, 1.055, -15.07, 0, 1, 1.06, 0.94],
    [13, 1, 13.5, 5.8, 0, 0, 1, 1.050, -15.16, 0, 1, 1.06, 0.94],
    [14, 1, 14.9, 5.0, 0, 0, 1, 1.036, -16.04, 0, 1, 1.06, 0.94],
]

_GEN = [
    # bus Pg     Qg    Qmax Qmin Vg     mBase status Pmax  Pmin  (+11 zero columns)
    [1, 232.4, -16.9, 10, 0, 1.060, 100, 1, 332.4, 0],
    [2, 40.0, 42.4, 50, -40, 1.045, 100, 1, 140.0, 0],
    [3, 0.0, 23.4, 40, 0, 1.010, 100, 1, 100.0, 0],
    [6, 0.0, 12.2, 24, -6, 1.070, 100, 1, 100.0, 0],
    [8, 0.0, 17.4, 24, -6, 1.090, 100, 1, 100.0, 0],
]

_BRANCH = [
    # f  t   r        x        b       tap
    [1, 2, 0.01938, 0.05917, 0.0528, 0],
    [1, 5, 0.05403, 0.22304, 0.0492, 0],
    [2, 3, 0.04699, 0.19797, 0.0438, 0],
    [2, 4, 0.05811, 0.17632, 0.0340, 0],
    [2, 5, 0.05695, 0.17388, 0.0346, 0],
    [3, 4, 0.06701, 0.17103, 0.0128, 0],
    [4, 5, 0.01335, 0.04211, 0.0, 0],
    [4, 7, 0.0, 0.20912, 0.0, 0.978],
    [4, 9, 0.0, 0.55618, 0.0, 0.969],
    [5, 6, 0.0, 0.25202, 0.0, 0.932],
    [6, 11, 0.09498, 0.19890, 0.0, 0],
    [6, 12, 0.12291, 0.25581, 0.0, 0],
    [6, 13, 0.06615, 0.13027, 0.0, 0],
    [7, 8, 0.0, 0.17615, 0.0, 0],
    [7, 9, 0.0, 0.11001, 0.0, 0],
    [9, 10, 0.03181, 0.08450, 0.0, 0],
    [9, 14, 0.12711, 0.27038, 0.0, 0],
    [10, 11, 0.08205, 0.19207, 0.0, 0],
    [12, 13, 0.22092, 0.19988, 0.0, 0],
    [13, 14, 0.17093, 0.34802, 0.0, 0],
]

_GENCOST = [
    # model startup shutdown n c2 c1 c0
    [2, 0, 0, 3, 0.0430293, 20, 0],
    [2, 0, 0, 3, 0.25, 20, 0],
    [2, 0, 0, 3, 0.01, 40, 0],
    [2, 0, 0, 3, 0.01, 40, 0],
    [2, 0, 0, 3, 0.01, 40, 0],
]


def case14() -> dict:
    """Return a fresh PYPOWER-format case dict (float64 arrays, 1-based bus ids)."""
    bus = np.array(_BUS, dtype=np.float64)
    gen = np.zeros((len(_GEN), 21), dtype=np.float64)
    gen[:, :10] = np.array(_GEN, dtype=np.float64)
    branch = np.zeros((len(_BRANCH), 13), dtype=np.float64)
    br = np.array(_BRANCH, dtype=np.float64)
    branch[:, F_BUS] = br[:, 0]
    branch[:, T_BUS] = br[:, 1]
    branch[:, BR_R] = br[:, 2]
    branch[:, BR_X] = br[:, 3]
    branch[:, BR_B] = br[:, 4]
    branch[:, RATE_A] = 9900.0
    branch[:, TAP] = br[:, 5]
    branch[:, BR_STATUS] = 1.0
    branch[:, ANGMIN] = -360.0
    branch[:, ANGMAX] = 360.0
    return {
        "version": "2",
        "baseMVA": 100.0,
        "bus": bus,
        "gen": gen,
        "branch": branch,
        "gencost": np.array(_GENCOST, dtype=np.float64),
    }


def make_ybus(base_mva: float, bus: np.ndarray, branch: np.ndarray) -> np.ndarray:
    """Dense complex bus-admittance matrix (the published MATPOWER ``makeYbus`` formula).

    ``bus[:, BUS_I]`` and the branch end columns must already be 0-based consecutive (the reference
    subtracts 1 before the call, evopf.py:260-263).  Ys = 1/(r+jx); Ytt = Ys + j b/2;
    Yff = Ytt/|tap|^2; Yft = -Ys/conj(tap); Ytf = -Ys/tap; Ysh = (Gs + j Bs)/baseMVA.
    """
    nb = bus.shape[0]
    stat = branch[:, BR_STATUS]
    ys = stat / (branch[:, BR_R] + 1j * branch[:, BR_X])
    bc = stat * branch[:, BR_B]
    tap = np.ones(branch.shape[0], dtype=np.complex128)
    nz = branch[:, TAP] != 0
    tap[nz] = branch[nz, TAP]
    tap = tap * np.exp(1j * np.pi / 180.0 * branch[:, SHIFT])
    ytt = ys + 1j * bc / 2.0
    yff = ytt / (tap * np.conj(tap))
    yft = -ys / np.conj(tap)
    ytf = -ys / tap
    ysh = (bus[:, GS] + 1j * bus[:, BS]) / base_mva
    f = branch[:, F_BUS].astype(int)
    t = branch[:, T_BUS].astype(int)
    ybus = np.zeros((nb, nb), dtype=np.complex128)
    for k in range(branch.shape[0]):
        ybus[f[k], f[k]] += yff[k]
        ybus[f[k], t[k]] += yft[k]
        ybus[t[k], f[k]] += ytf[k]
        ybus[t[k], t[k]] += ytt[k]
    ybus[np.arange(nb), np.arange(nb)] += ysh
    return ybus


def ybus14() -> np.ndarray:
    ppc = case14()
    bus = ppc["bus"].copy()
    branch = ppc["branch"].copy()
    bus[:, BUS_I] -= 1
    branch[:, [F_BUS, T_BUS]] -= 1
    return make_ybus(ppc["baseMVA"], bus, branch)
