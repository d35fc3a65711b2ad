"""Tiny Box space (gym.spaces.Box when gym is installed) and the environment base class."""
import numpy as np


class _PlainEnv(object):
    """What gym 0.19-0.21's `gym.make` touches on an environment instance (gym/envs/registration.py: `env.unwrapped.spec
    = spec`, then TimeLimit / OrderEnforcing wrappers read `spec`, `reward_range`, `metadata`, the spaces): used as the
    base class when gym itself is not importable, so the same classes register and wrap under a real gym as well."""
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False

    def close(self):
        pass


try:  # pragma: no cover - depends on the host
    import gym as _gym
    EnvBase = _gym.Env if all(hasattr(_gym.Env, a) for a in ("unwrapped", "reward_range", "spec")) else _PlainEnv
except Exception:
    EnvBase = _PlainEnv

try:  # pragma: no cover - depends on the host
    from gym.spaces import Box  # type: ignore
except Exception:
    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.asarray(low).shape
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)
