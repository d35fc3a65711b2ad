import importlib

registry = {}


class TimeLimit(object):
    """Sets done=True after max_episode_steps; forwards every other attribute to the wrapped env."""

    def __init__(self, env, max_episode_steps):
        self.__dict__["env"] = env
        self.__dict__["_max_episode_steps"] = max_episode_steps
        self.__dict__["_elapsed_steps"] = 0

    def __getattr__(self, name):
        if name in ("env", "__setstate__", "__getstate__", "__deepcopy__", "__reduce_ex__", "__reduce__"):
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)

    def __setattr__(self, name, value):
        if name in self.__dict__:
            self.__dict__[name] = value
        else:
            setattr(self.__dict__["env"], name, value)

    def reset(self):
        self.__dict__["_elapsed_steps"] = 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.__dict__["_elapsed_steps"] += 1
        if self._max_episode_steps is not None and self._elapsed_steps >= self._max_episode_steps:
            done = True
        return obs, reward, done, info

    def __deepcopy__(self, memo):
        import copy
        new = TimeLimit.__new__(TimeLimit)
        new.__dict__["env"] = copy.deepcopy(self.__dict__["env"], memo)
        new.__dict__["_max_episode_steps"] = self._max_episode_steps
        new.__dict__["_elapsed_steps"] = self._elapsed_steps
        return new


def register(id, entry_point, max_episode_steps=None, **kwargs):
    registry[id] = (entry_point, max_episode_steps, kwargs)


def make(id, **kwargs):
    entry_point, max_episode_steps, kw = registry[id]
    mod_name, cls_name = entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    env = cls(**{**kw, **kwargs})
    if max_episode_steps is not None:
        env = TimeLimit(env, max_episode_steps)
    return env
