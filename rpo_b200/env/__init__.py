"""Drop-in environments: same constructor-free API, attributes and hard-constraint protocol as rpo.env
(rpo/env/__init__.py:1-2), with every protocol method and the env step running on the sm_100a kernels."""
from .evopf import EVOPFEnv  # noqa: F401
from .cartpole import CartSafeEnv  # noqa: F401
from .pendulum import SpringPendulumEnv  # noqa: F401
from .registration import register_all  # noqa: F401
