"""Exogenous demand / price series for EVOPF-v0 with the reference loaders' semantics
(rpo/env/electrical_grid/data/demand.py:7-75, price.py:7-55).

The reference averages two pickled time series into 24-step daily curves.  Those data files are assets of the
reference and are not redistributed: pass ``curve=`` explicitly or ``pickle_path=`` pointing at the reference's
``crawlers/*.pickle``; EVOPFEnv locates them by itself when the reference package is importable
(`find_reference_data_dir`); only in stand-alone use (no reference anywhere) a smooth synthetic daily profile of the
same mean level is used, with a warning (documented in DESIGN.md).
"""
from __future__ import annotations

import pickle

import numpy as np

from ..case14 import PD, QD


def find_reference_data_dir():
    """The reference keeps its recorded demand / price series beside its loaders (rpo/env/electrical_grid/data/crawlers,
    evopf.py:212-216 -> data/demand.py:14, data/price.py:13).  When the reference package `rpo` is importable — which it
    is whenever rpo_b200 is used as a drop-in under the reference's own algorithm classes — use its files, so that the
    drop-in EVOPFEnv sees the same exogenous curves.  Returns None when `rpo` cannot be located (stand-alone use)."""
    import importlib.util
    import os
    try:
        spec = importlib.util.find_spec("rpo")
    except (ImportError, ValueError):
        spec = None
    for root in (list(spec.submodule_search_locations) if spec is not None and spec.submodule_search_locations else []):
        d = os.path.join(root, "env", "electrical_grid", "data", "crawlers")
        if os.path.isfile(os.path.join(d, "demand.pickle")) and os.path.isfile(os.path.join(d, "price.pickle")):
            return d
    return None


def _daily_curve(pickle_path, key, T):
    with open(pickle_path, "rb") as f:
        data = np.array(pickle.load(f)[key])
    total = len(data) // T
    curve = np.zeros(T)
    for i in range(total):
        curve += data[i * T: i * T + T]
    return curve / total


def synthetic_curve(T=24, level=1.0, swing=0.16):
    h = np.arange(T)
    return level * (1.0 + swing * np.sin((h - 9.0) * 2.0 * np.pi / T) + 0.5 * swing * np.sin((h - 3.0) * 4.0 * np.pi / T))


class DemandLoader(object):
    def __init__(self, ppc, rho=0.5, q_rand=False, T=24, min_pf=0.9, max_pf=1.0, regularized=True, curve=None,
                 pickle_path=None):
        if curve is None:
            curve = _daily_curve(pickle_path, "value", T) if pickle_path else synthetic_curve(T, level=1000.0)
        self.curve = np.asarray(curve, dtype=np.float64)
        self.T, self.q_rand, self.rho = T, q_rand, rho
        self.baseMVA, self.bus, self.gen = ppc["baseMVA"], ppc["bus"], ppc["gen"]
        self.min_pf, self.max_pf = min_pf, max_pf
        self.regularized = regularized
        self.cache = None
        self.counter = None

    def reset(self):
        self.counter = 0
        pd, qd = self.process(self.curve)
        self.cache = np.concatenate((pd, qd), axis=1) / self.baseMVA
        return self.fetch()

    def process(self, ps):                                   # demand.py:47-67
        nbus = self.bus.shape[0]
        idx = np.nonzero(self.bus[:, PD])[0]
        p_total = self.bus[:, PD].sum() * self.T * 1.0
        ps = (ps / ps.sum() * p_total)[:, None]
        rand_mat = np.random.dirichlet(np.ones(len(idx)), size=self.T)
        p_ratio = self.bus[:, PD] / self.bus[:, PD].sum()
        p_ratio = p_ratio[None].repeat(self.T, axis=0) * (1 - self.rho)
        p_ratio[:, idx] += self.rho * rand_mat
        pd = p_ratio * ps
        if self.q_rand:
            pf = np.random.rand(self.T, nbus) * (self.max_pf - self.min_pf) + self.min_pf
            qd = pd * np.tan(np.arccos(pf)) * np.sign(self.bus[:, QD])
        else:
            qd = np.broadcast_to(self.bus[:, QD], pd.shape).copy()
        return pd, qd

    def fetch(self):                                          # demand.py:69-75
        done = self.counter >= self.T
        tmp = np.zeros_like(self.cache[0]) if done else self.cache[self.counter].copy()
        self.counter += 1
        return tmp, done


class PriceLoader(object):
    def __init__(self, T=24, ahead=24, regularized=True, reg_bias=0.1, reg_sigma=0.05, curve=None, pickle_path=None):
        if curve is None:
            curve = _daily_curve(pickle_path, "SYS", T) if pickle_path else synthetic_curve(T, level=49.0, swing=0.25)
        self.curve = np.asarray(curve, dtype=np.float64)
        self.T, self.ahead = T, ahead
        self.regularized, self.reg_bias, self.reg_sigma = regularized, reg_bias, reg_sigma
        self.cache = None
        self.counter = None

    def reset(self):
        self.counter = 0
        tmp = np.concatenate([self.curve, np.zeros(self.T)])
        self.cache = self.process(tmp)
        return self.fetch()

    def process(self, pc):                                    # price.py:40-47
        mag_r = self.reg_bias * np.random.randn() + 1
        return mag_r * pc * (1 + np.random.randn(self.T * 2))

    def fetch(self):                                          # price.py:49-55
        done = self.counter >= self.T
        tmp = np.zeros(self.ahead) if done else self.cache[self.counter:self.counter + self.ahead].copy()
        self.counter += 1
        return tmp, done
