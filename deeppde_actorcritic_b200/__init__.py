"""B200-native rollout + TD training step of MoZhou1995/DeepPDE_ActorCritic (drop-in API).

    from deeppde_actorcritic_b200 import equation, ActorCriticSolver, load_config
"""
from . import _cabi  # noqa: F401  (ctypes table only; the .so is loaded on first use)
from .config import Config, load_config, munchify  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import deeppde_actorcritic_b200` stays cheap
    if name in ("equation", "solver", "engine"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in ("ActorCriticSolver", "CriticModel", "ActorModel", "DeepNN"):
        from . import solver
        return getattr(solver, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
