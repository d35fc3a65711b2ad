"""Gym registration under the reference's ids (rpo/env/classic_control/__init__.py:6-16,
rpo/env/electrical_grid/__init__.py:4-7).  gym is optional: without it the classes are used directly."""


def register_all(force: bool = False) -> bool:
    try:
        from gym.envs.registration import register, registry
    except Exception:
        return False
    specs = [
        ("CartSafe-v0", "rpo_b200.env:CartSafeEnv", 200),
        ("SpringPendulum-v0", "rpo_b200.env:SpringPendulumEnv", 200),
        ("EVOPF-v0", "rpo_b200.env:EVOPFEnv", None),
    ]
    for id_, entry, steps in specs:
        # gym <= 0.21: `registry` is an EnvRegistry whose specs live in `.env_specs` (it has no __contains__ / __delitem__);
        # gym >= 0.22 and the test stub: a plain dict
        specs_map = getattr(registry, "env_specs", None)
        if specs_map is None:
            specs_map = registry
        try:
            present = id_ in specs_map
        except Exception:
            present = False
        if present and force:
            try:
                del specs_map[id_]
            except Exception:
                pass
        if not present or force:
            kw = dict(id=id_, entry_point=entry)
            if steps is not None:
                kw["max_episode_steps"] = steps
            register(**kw)
    return True
