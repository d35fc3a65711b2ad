"""Minimal stand-in for gym 0.19 — test infrastructure only (oracle/ is never imported by the product).

Supplies exactly the surface the reference touches: ``Env``, ``spaces.Box``,
``utils.seeding.np_random``, ``logger.warn``, ``envs.registration.register``, ``make`` with a
TimeLimit wrapper.  No arithmetic of the hot path lives here.
"""
import copy
import importlib

from . import spaces, logger  # noqa: F401
from .envs.registration import register, make, registry  # noqa: F401


class Env(object):
    metadata = {}
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]
