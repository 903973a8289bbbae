"""TEST INFRASTRUCTURE (build container only): load the UNMODIFIED reference env layer.

Registers a synthetic package ``WindGym`` whose ``__path__`` points at ``/root/reference/WindGym`` but whose
``__init__`` is empty (the real one drags in xarray / torch examples and a dangling gymnasium registration,
SURVEY.md Q12), and puts ``oracle/shims`` on ``sys.path`` so the un-installed third-party *names* resolve.
``/root/reference`` does not exist on the GPU box: anything that calls this must skip there.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WINDGYM_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "WindGym", "Wind_Farm_Env.py"))


def load_reference():
    """Returns a namespace with the reference's own classes (unmodified source, executed in place)."""
    if not reference_available():
        raise FileNotFoundError(f"reference checkout not found under {REFERENCE_ROOT}")
    for p in (_SHIMS, _REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "WindGym" not in sys.modules:
        pkg = types.ModuleType("WindGym")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "WindGym")]
        sys.modules["WindGym"] = pkg
        # Agents/__init__ imports PyWakeAgent (needs py_wake's steady-state models: out of scope)
        agents = types.ModuleType("WindGym.Agents")
        agents.__path__ = [os.path.join(REFERENCE_ROOT, "WindGym", "Agents")]
        sys.modules["WindGym.Agents"] = agents
    ns = types.SimpleNamespace()
    ns.MesClass = importlib.import_module("WindGym.MesClass")
    ns.WindEnv = importlib.import_module("WindGym.WindEnv")
    ns.BasicControllers = importlib.import_module("WindGym.BasicControllers")
    ns.Wind_Farm_Env = importlib.import_module("WindGym.Wind_Farm_Env")
    ns.FarmEval = importlib.import_module("WindGym.FarmEval")
    ns.WindEnvMulti = importlib.import_module("WindGym.WindEnvMulti")
    ns.ConstantAgent = importlib.import_module("WindGym.Agents.ConstantAgent").ConstantAgent
    ns.BaseAgent = importlib.import_module("WindGym.Agents.BaseAgent").BaseAgent
    ns.WindFarmEnv = ns.Wind_Farm_Env.WindFarmEnv
    ns.FarmEvalCls = ns.FarmEval.FarmEval
    ns.examples = os.path.join(REFERENCE_ROOT, "examples", "EnvConfigs")
    return ns
