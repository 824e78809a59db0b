"""gn-ode-sir_b200: B200-native GN-ODE rollout (hot path of sissykosm/GN-ODE-SIR).

Import name: ``gn_ode_sir_b200`` (the repo-root shim ``gn_ode_sir_b200.py`` maps the
hyphenated directory onto it).
"""
from . import _lib                                             # noqa: F401
from ._build import LIB_PATH, build_library                    # noqa: F401
from .graph import BatchCache, DeviceBatch, DeviceGraph        # noqa: F401
from . import rollout                                          # noqa: F401  (rollout.rollout / .aggregate / .odefunc_eval)
from . import ode_sim, ode_ngraphs, parallel, synth, harness    # noqa: F401
