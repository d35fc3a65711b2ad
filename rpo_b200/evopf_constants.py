"""Constants and index maps of the EVOPF-v0 problem (IEEE-14 + 5 batteries).

Derived from the case data exactly the way EVOPFEnv.__init__ derives them
(rpo/env/electrical_grid/evopf.py:220-337); consumed by the host mirror (rpo_b200/evopf.py), by the
code generator of the CUDA kernels (tools/gen_evopf.py) and by the parity tests.
"""
from __future__ import annotations

import numpy as np

from .case14 import (BUS_TYPE, MBASE, PMAX, PMIN, QMAX, QMIN, VA, VM, VMAX, VMIN, case14, ybus14)

_ppc = case14()
_bus, _gen = _ppc["bus"], _ppc["gen"]

NBUS = _bus.shape[0]            # 14
NG = _gen.shape[0]              # 5
NE = NG                         # 5 batteries, one per generator bus (evopf.py:242)
BASE_MVA = float(_ppc["baseMVA"])
GENBASE = _gen[:, MBASE].copy()

SLACK = np.where(_bus[:, BUS_TYPE] == 3)[0]                      # [0]
PV = np.where(_bus[:, BUS_TYPE] == 2)[0]                         # [1, 2, 5, 7]
SPV = np.sort(np.concatenate([SLACK, PV]))                       # [0, 1, 2, 5, 7]
PQ = np.setdiff1d(np.arange(NBUS), SPV)                          # 9 load buses
SLACK_ = np.array([int(np.where(x == SPV)[0][0]) for x in SLACK])
PV_ = np.array([int(np.where(x == SPV)[0][0]) for x in PV])
SPV_ = np.arange(NG)
NPV, NSLACK, NPQ = len(PV), len(SLACK), len(PQ)

XDIM = 2 * NBUS                  # 28
YDIM = 2 * NG + 2 * NBUS + NE    # 43
ZDIM = NPV + NG + NE             # 14
STATE_DIM = 57
NAHEAD = 24
EQ_NUM, INEQ_NUM = 28, 58

PG0, QG0, VM0, VA0, PE0 = 0, NG, 2 * NG, 2 * NG + NBUS, 2 * NG + 2 * NBUS

PARTIAL_VARS = np.concatenate([PG0 + PV_, VM0 + SPV, VA0 + SLACK, PE0 + SPV_])
OTHER_VARS = np.setdiff1d(np.arange(YDIM), PARTIAL_VARS)
PARTIAL_ACTIONS = np.concatenate([PG0 + PV_, VM0 + SPV, PE0 + SPV_])
OTHER_ACTIONS = OTHER_VARS.copy()

PG_PV_ZIDX = np.arange(NPV)
VM_SPV_ZIDX = np.arange(NPV, 2 * NPV + NSLACK)
PE_SPV_ZIDX = np.arange(2 * NPV + NSLACK, 2 * NPV + NSLACK + NE)
PE_PV_ZIDX = PE_SPV_ZIDX[1:]

KEEP_CONSTR = np.concatenate([PV, PQ, NBUS + PQ])                         # Newton rows (evopf.py:809-812)
NEWTON_GUESS_INDS = np.concatenate([VM0 + PQ, VA0 + PV, VA0 + PQ])        # Newton cols (evopf.py:813-816)

PMAX_PU = _gen[:, PMAX] / GENBASE
PMIN_PU = _gen[:, PMIN] / GENBASE
QMAX_PU = _gen[:, QMAX] / GENBASE
QMIN_PU = _gen[:, QMIN] / GENBASE
VMAX_PU = _bus[:, VMAX].copy()
VMIN_PU = _bus[:, VMIN].copy()
VM_INIT = _bus[:, VM].copy()
VA_INIT = np.deg2rad(_bus[:, VA])
SLACK_VA = VA_INIT[SLACK].copy()

QUAD_COSTS = _ppc["gencost"][:, 4].copy()
LIN_COSTS = _ppc["gencost"][:, 5].copy()
CONST_COST = float(_ppc["gencost"][:, 6].sum())
WG, WE = 1.0, 5.0                                                           # evopf.py:244-245

YBUS = ybus14()
G = YBUS.real.copy()
B = YBUS.imag.copy()

# Battery (evopf.py:25-26, 243)
B_LOW, B_HIGH, B_PMIN, B_PMAX, ETA_IN, ETA_OUT = 0.1, 0.8, -0.2, 0.2, 0.9, 0.9

ACTION_HIGH = np.concatenate([PMAX_PU, QMAX_PU, VMAX_PU, np.full(NBUS, np.pi), np.full(NE, B_PMAX)]).astype(np.float32)
ACTION_LOW = np.concatenate([PMIN_PU, QMIN_PU, VMIN_PU, np.full(NBUS, -np.pi), np.full(NE, B_PMIN)]).astype(np.float32)

NEWTON_TOL, NEWTON_MAX_ITERS, NEWTON_BSZ = 1e-5, 50, 256                    # evopf.py:786
