"""Host-side logic that needs no GPU: loop-exit reconstruction, env bookkeeping, loaders, sharding."""
import copy

import numpy as np
import torch


def test_reference_step_count():
    from rpo_b200.algo import reference_step_count as f
    t = lambda l: torch.tensor(l, dtype=torch.int32)
    assert f(t([1] * 10), 10) == 10
    assert f(t([0] * 10), 10) == 1            # `step == 0 or ...`: at least one step (rpo_ddpg.py:271)
    assert f(t([1, 1, 1, 0, 1, 0, 0, 0, 0, 0]), 10) == 3
    assert f(t([1, 0]), 2) == 1
    assert f(t([1]), 1) == 1
    assert f(t([1]), 0) == 0


def test_shard_rows_partition():
    from rpo_b200.dist import shard_rows
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 5):
        for w in (1, 2, 3, 8):
            parts = [shard_rows(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_evopf_env_host_side():
    from oracle.rpo_oracle import EvopfOracle
    from rpo_b200.env import EVOPFEnv
    np.random.seed(3)
    env = EVOPFEnv()
    orc = EvopfOracle()
    assert env.volatile is True and env.nahead == 24
    for name in ("partial_actions", "other_actions", "partial_vars", "other_vars"):
        assert list(getattr(env, name)) == list(getattr(orc, name if name in ("partial_actions", "partial_vars", "other_vars") else "other_vars"))
    lo, hi = env.box_constraint_partial
    assert lo.shape == (14,) and np.allclose(hi[:4], [1.4, 1, 1, 1]) and np.allclose(lo[4:9], 0.94)
    s = env.reset()
    assert s.shape == (57,) and np.allclose(s[28:33], 0.2) and s[0] == 0.0       # bus 1 carries no load
    assert abs(s[:14].sum() - 2.59) < 0.6
    lo, hi = env.update(torch.tensor(s, dtype=torch.float32).view(1, -1))
    lo_o, hi_o = orc.partial_box(torch.tensor(s, dtype=torch.float32).view(1, -1))
    assert np.array_equal(lo, lo_o) and np.array_equal(hi, hi_o)
    e2 = copy.deepcopy(env)
    assert e2 is not env and np.array_equal(e2.state, env.state)
    # 24 fetches then done (demand.py:69-75); reset() consumed one
    d = [env.data.fetch()[1] for _ in range(24)]
    assert d[:23] == [False] * 23 and d[23] is True


def test_loaders_follow_reference_semantics():
    from rpo_b200.case14 import case14
    from rpo_b200.env.loaders import DemandLoader, PriceLoader
    np.random.seed(0)
    d = DemandLoader(case14(), rho=0.5, q_rand=True)
    row, done = d.reset()
    assert row.shape == (28,) and not done
    pd = d.cache[:, :14]
    assert np.all(pd[:, [0, 6, 7]] == 0) and np.all(pd >= 0)
    assert abs(pd.sum() - 2.59 * 24) < 1e-6                       # total energy preserved (demand.py:51-52)
    ratio = np.abs(d.cache[:, 14:]) / np.maximum(pd, 1e-12)
    assert ratio[pd > 0].max() <= np.tan(np.arccos(0.9)) + 1e-9
    p = PriceLoader()
    w, done = p.reset()
    assert w.shape == (24,) and p.cache.shape == (48,) and np.all(p.cache[24:] == 0)


def test_classic_env_host_side():
    from rpo_b200.env import CartSafeEnv, SpringPendulumEnv
    np.random.seed(123)
    c = CartSafeEnv()
    assert list(c.partial_actions) == [1]                         # seed 123 -> [1]
    assert (c.state_dim, c.action_dim, c.eq_num, c.ineq_num) == (6, 2, 1, 6)
    assert c.reset().shape == (6,) and c.volatile is False and c.update is None
    p = SpringPendulumEnv()
    assert list(p.partial_actions) == [0] and p.reset().shape == (5,)
    copy.deepcopy(c), copy.deepcopy(p)


def test_rows_helper_accepts_strided_views():
    from rpo_b200 import ops
    if not torch.cuda.is_available():
        import pytest
        with pytest.raises(Exception):
            ops._rows(torch.zeros(4, 57), "state", 28)
        return


def test_generated_header_is_up_to_date(tmp_path):
    """tools/gen_evopf.py is deterministic and the committed rpo_b200/csrc/evopf_gen.cuh is its current output."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_evopf", os.path.join(root, "tools", "gen_evopf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = mod.Gen()
    txt = g.generate()
    assert txt == open(os.path.join(root, "rpo_b200", "csrc", "evopf_gen.cuh")).read()
    assert g.order == [7, 2, 10, 9, 11, 12, 13, 1, 4, 5, 3, 6, 8] and g.n == 22 and g.nL == 70 and g.nU == 92
    st = mod.op_stats(txt)
    assert 1500 < st["newton_iter_flop"] < 3000 and 2000 < st["corr_step_flop"] < 4000
