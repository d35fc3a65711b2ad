# dry-run of a -m gpu test module on CPU with the host simulator swapped in (checks the test logic itself)
import sys, importlib, inspect, itertools
sys.path.insert(0, "/root/repo")
import pytest, torch
from _pytest.monkeypatch import MonkeyPatch
from tests import hostsim_ops
mp = MonkeyPatch(); hostsim_ops.install(mp)
torch.Tensor.cuda = lambda self, *a, **k: self          # the tests move tensors with .cuda(): stay on the host
torch.cuda.synchronize = lambda *a, **k: None
modname = sys.argv[1]
m = importlib.import_module(modname)
m.DEV = "cpu"
from oracle.rpo_oracle import EvopfOracle
fix = {"evopf_oracle": EvopfOracle(), "evopf_oracle64": EvopfOracle(dtype=torch.float64)}
only = sys.argv[2:] 
for name, fn in inspect.getmembers(m, inspect.isfunction):
    if not name.startswith("test_") or (only and name not in only): continue
    marks = [mk for mk in getattr(fn, "pytestmark", []) if mk.name == "parametrize"]
    names = [mk.args[0] for mk in marks]; vals = [mk.args[1] for mk in marks]
    for combo in itertools.product(*vals) if marks else [()]:
        kw = dict(zip(names, combo))
        for a in inspect.signature(fn).parameters:
            if a in fix: kw[a] = fix[a]
        if kw.get("B", 0) and kw["B"] > 300: continue
        try:
            fn(**kw); print("ok  ", name, {k: v for k, v in kw.items() if k not in fix})
        except Exception as e:
            print("FAIL", name, {k: v for k, v in kw.items() if k not in fix}, type(e).__name__, str(e)[:300])
