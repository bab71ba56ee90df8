"""Run the reference's own ``train.py`` (or any of its caller scripts) UNCHANGED on top of the CUDA environment.

    python -m multi_agent_rl_for_fjsp_b200.run_reference_caller /path/to/reference train.py --timesteps 512 ...

``sys.path`` is arranged so that ``from FJSPParallelEnvWrapper import FJSPParallelEnv`` (train.py:20) resolves to
``dropin/FJSPParallelEnvWrapper.py`` while ``a2c``, ``networks``, ``transition_memory``, ``visualization``,
``constants``, ``utils.Logger`` and ``enums`` still come from the reference tree; ``gymnasium`` / ``matplotlib`` /
``pettingzoo`` stand-ins are used only when the real packages are missing.  Nothing of the reference's simulation
(FJSPSimulation, agents/, models/, simpy) is imported.
"""
from __future__ import annotations

import importlib.util
import os
import runpy
import sys
import types


def _load_stand_in(compat_dir, name):
    """Register compat/<name> (a package directory) as top-level module `name`, submodules resolved lazily from it."""
    init = os.path.join(compat_dir, name, "__init__.py")
    spec = importlib.util.spec_from_file_location(name, init, submodule_search_locations=[os.path.join(compat_dir, name)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(__doc__)
    ref_root, script = os.path.abspath(argv[0]), argv[1]
    pkg = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, ref_root)
    # stand-ins ONLY for the packages that are really missing: each one is loaded from compat/<name> under its own name,
    # so an installed gymnasium is never shadowed because, say, matplotlib is absent
    for mod in ("gymnasium", "pettingzoo", "matplotlib"):
        if importlib.util.find_spec(mod) is None:
            _load_stand_in(os.path.join(pkg, "compat"), mod)
    sys.path.insert(0, os.path.join(pkg, "dropin"))
    sys.path.insert(0, os.path.dirname(pkg))
    # the reference's `utils` / `enums` are namespace directories that site-packages may shadow
    for name in ("utils", "enums"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(ref_root, name)]
        sys.modules[name] = m
    sys.argv = [os.path.join(ref_root, script)] + argv[2:]
    runpy.run_path(os.path.join(ref_root, script), run_name="__main__")


if __name__ == "__main__":
    main()
