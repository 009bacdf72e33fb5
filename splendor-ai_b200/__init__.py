"""splendor-ai_b200: a B200-native batched engine for the game-rule path of roeey777/Splendor-AI.

Public surface (mirrors the reference's rule path):
  BatchedSplendor, playout_games      batched engine over the C ABI (include/splendor_b200.h)
  state                               packed HBM layout <-> numpy fields
  build()                             compile csrc/ for sm_100a
"""
from . import _native, state  # noqa: F401
from ._native import SplendorNativeError, build  # noqa: F401


def __getattr__(name):  # torch is imported lazily so that `build()` works in a bare container
    if name in ("BatchedSplendor", "playout_games", "counters_dict", "engine"):
        import importlib

        engine = importlib.import_module(__name__ + ".engine")
        return engine if name == "engine" else getattr(engine, name)
    if name in ("SplendorEnv", "SplendorVecEnv"):
        import importlib

        return getattr(importlib.import_module(__name__ + ".env"), name)
    if name in ("actions", "game_rule", "env", "template", "dist", "minimax", "game", "rollout", "sampling"):
        import importlib

        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
