"""Test infrastructure (NOT product code): makes the read-only reference importable in THIS container.

Only `oracle/gen_golden.py`, `tools/gen_tables.py` and the `-m "not gpu"` pin tests import this. The GPU box has
no /root/reference, so nothing on the product path, in `-m gpu` tests, smoke() or bench.py may use it.

The reference's gym adapter imports `gymnasium`, which is absent here; a ~25-line stand-in providing
`Env`, `spaces.Discrete/Box` and `envs.registration.register` is enough for
`splendor.splendor.gym.envs.{actions,utils,splendor_env}` to import unmodified.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "splendor"))


def _install_gymnasium_stub() -> None:
    if "gymnasium" in sys.modules:
        return
    g = types.ModuleType("gymnasium")

    class Env:  # noqa: D401 - minimal stand-in
        def __init__(self):
            pass

        def reset(self, *, seed=None, options=None):
            return None

    sp = types.ModuleType("gymnasium.spaces")

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Box:
        def __init__(self, low, high, shape):
            self.low, self.high, self.shape = low, high, shape

    sp.Discrete, sp.Box = Discrete, Box
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **kwargs: None
    envs.registration = reg
    g.Env, g.spaces, g.envs = Env, sp, envs
    sys.modules.update(
        {"gymnasium": g, "gymnasium.spaces": sp, "gymnasium.envs": envs, "gymnasium.envs.registration": reg}
    )


def install() -> None:
    """Put the reference on sys.path (read-only use) with the gymnasium stand-in."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_SRC)
    _install_gymnasium_stub()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
