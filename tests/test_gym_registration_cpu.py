"""The drop-in environments under gym's own `make` protocol.  gym is not installed in this image (the reference needs
0.19, README.md:16), so this emulates what gym 0.19-0.21's registration does to an environment instance —
gym/envs/registration.py: EnvSpec.make sets `env.unwrapped.spec = spec`, EnvRegistry keeps its specs in `.env_specs`
(no __contains__/__delitem__ on the registry object), and the TimeLimit wrapper reads `spec`, `reward_range`,
`metadata` and the spaces in gym.Wrapper.__init__ — with the same attribute accesses, to pin the surface the classes
must offer (they were plain `object`s without `unwrapped` / `spec` / `reward_range`)."""
import copy
import importlib
import sys
import types

import pytest


class _EnvSpec(object):
    def __init__(self, id, entry_point, max_episode_steps=None, **kw):
        self.id, self.entry_point, self.max_episode_steps, self._kwargs = id, entry_point, max_episode_steps, kw

    def make(self, **kwargs):
        mod, name = self.entry_point.split(":")
        env = getattr(importlib.import_module(mod), name)(**{**self._kwargs, **kwargs})
        spec = copy.deepcopy(self)
        spec._kwargs = kwargs
        env.unwrapped.spec = spec                      # gym 0.19 registration.py: EnvSpec.make
        return env


class _EnvRegistry(object):                            # gym <= 0.21: not a dict
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kw):
        if id in self.env_specs:
            raise RuntimeError("Cannot re-register id: %s" % id)
        self.env_specs[id] = _EnvSpec(id, **kw)

    def make(self, id, **kwargs):
        env = self.env_specs[id].make(**kwargs)
        if env.spec.max_episode_steps is not None:
            env = _TimeLimit(env, env.spec.max_episode_steps)
        return env


class _TimeLimit(object):                              # gym.Wrapper.__init__ + TimeLimit
    def __init__(self, env, max_episode_steps):
        self.env = env
        self.action_space, self.observation_space = env.action_space, env.observation_space
        self.reward_range, self.metadata = env.reward_range, env.metadata
        self._max_episode_steps = max_episode_steps

    @property
    def unwrapped(self):
        return self.env.unwrapped


@pytest.fixture()
def fake_gym(monkeypatch):
    reg = _EnvRegistry()
    regmod = types.ModuleType("gym.envs.registration")
    regmod.registry, regmod.register = reg, (lambda id, **kw: reg.register(id, **kw))
    envs = types.ModuleType("gym.envs")
    envs.registration = regmod
    gym = types.ModuleType("gym")
    gym.envs, gym.make = envs, reg.make
    for name, m in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", regmod)):
        monkeypatch.setitem(sys.modules, name, m)
    return gym, reg


def test_dropin_envs_register_and_make_under_gym019_semantics(fake_gym):
    gym, reg = fake_gym
    from rpo_b200.env import register_all
    assert register_all(force=False)
    assert set(reg.env_specs) == {"CartSafe-v0", "SpringPendulum-v0", "EVOPF-v0"}
    assert register_all(force=True)                    # re-registration over an EnvRegistry (no __delitem__ on it)
    import numpy as np
    for id_, steps in (("CartSafe-v0", 200), ("SpringPendulum-v0", 200), ("EVOPF-v0", None)):
        kw = dict(device="cpu")
        if id_ == "EVOPF-v0":
            kw.update(demand_curve=np.ones(24), price_curve=np.ones(24))
        env = gym.make(id_, **kw)
        base = env.unwrapped
        assert base.spec.id == id_ and base.spec.max_episode_steps == steps
        assert (steps is None) == (env is base)        # EVOPF is registered without a TimeLimit (electrical_grid/__init__.py:4-7)
        assert base.reward_range == (-float("inf"), float("inf")) and isinstance(base.metadata, dict)
        assert base.action_space.shape == (base.action_dim,) and base.observation_space.shape == (base.state_dim,)
        with base as e:                                # gym.Env is a context manager
            assert e is base
