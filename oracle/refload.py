"""Load and drive the UNMODIFIED reference sources from ``/root/reference`` (TEST INFRASTRUCTURE ONLY).

Used to (a) generate the golden vectors committed under ``tests/golden/`` (``oracle/gen_golden.py``)
and (b) differential-test the C restatement ``oracle/fjsp_oracle.c`` against the live reference in
this container.  ``/root/reference`` does not exist on the GPU box, so everything that imports this
module is skipped there (``reference_available()``).

Three obstacles, handled here:
  1. ``simpy`` / ``gymnasium`` / ``pettingzoo`` / ``matplotlib`` are not installed -> the SimPy
     restatement in ``oracle/shims`` and the shape-only stand-ins in the package's ``compat/`` are
     put on ``sys.path`` (only if the real package is missing).
  2. ``site-packages/agents`` is an unrelated regular package that shadows the reference's
     ``agents/`` namespace directory -> namespace modules for ``agents``, ``models``, ``utils``,
     ``enums`` are pre-registered with ``__path__`` pointing into the reference tree.
  3. The reference prints on every order completion (``FJSPSimulation.py:258``) -> callers may
     silence stdout with ``quiet()``.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("FJSP_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(REPO, "oracle", "shims")
_COMPAT = os.path.join(REPO, "multi_agent_rl_for_fjsp_b200", "compat")

AGENT_IDS = [
    "pickup_station", "agv", "small_machine", "big_machine",
    "packaging_blue_1", "packaging_blue_2", "packaging_red", "packaging_green",
]
N_ACTIONS = (3, 8, 3, 3, 3, 3, 3, 3)

_loaded = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "FJSPSimulation.py"))


def _missing(mod: str) -> bool:
    try:
        return importlib.util.find_spec(mod) is None
    except (ImportError, ValueError):
        return True


def load_reference():
    """Import the reference modules; returns a namespace with FJSPParallelEnv, FJSPSimulation, enums."""
    if _loaded:
        for p in _loaded["paths"]:  # a test fixture may have restored sys.path since the first load
            if p not in sys.path:
                sys.path.insert(0, p)
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    paths = [REFERENCE_ROOT]
    if _missing("simpy"):
        sys.path.insert(0, _SHIMS)
        paths.append(_SHIMS)
    if _missing("gymnasium") or _missing("pettingzoo") or _missing("matplotlib"):
        sys.path.insert(0, _COMPAT)
        paths.append(_COMPAT)
    for pkg in ("agents", "models", "utils", "enums"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
        m.__package__ = pkg
        sys.modules[pkg] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for stale in ("FJSPSimulation", "FJSPParallelEnvWrapper", "constants"):
        sys.modules.pop(stale, None)
    ns = types.SimpleNamespace()
    ns.constants = importlib.import_module("constants")
    ns.FJSPSimulation = importlib.import_module("FJSPSimulation").FJSPSimulation
    ns.FJSPParallelEnv = importlib.import_module("FJSPParallelEnvWrapper").FJSPParallelEnv
    ns.ProductType = importlib.import_module("enums.ProductType").ProductType
    ns.PackagingColor = importlib.import_module("enums.PackagingColor").PackagingColor
    ns.LocationType = importlib.import_module("enums.LocationType").LocationType
    ns.Product = importlib.import_module("models.Product").Product
    ns.Order = importlib.import_module("models.Order").Order
    _loaded["ns"] = ns
    _loaded["paths"] = paths
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def reset_with_orders(env, orders):
    """Reset the reference env and install an explicit order table.

    ``orders`` is a sequence of ``(n_products, product_type, colour)`` with the 1-based enum values
    (ProductType SMALL=1/MEDIUM=2/BIG=3, PackagingColor RED=1/BLUE=2/GREEN=3).  The reference draws
    orders from the global NumPy RNG inside ``reset`` (``FJSPSimulation.py:107-112``); to feed both
    sides identical streams we reset with ``num_orders=0`` and then replay ``generate_order``'s body
    (``FJSPSimulation.py:109-129``) with the given values instead of RNG draws.
    """
    ns = load_reference()
    obs, infos = env.reset(options={"num_orders": 0})
    sim = env.unwrapped.simulation
    for (n, t, c) in orders:
        order_id = len(sim.orders)
        ptype = ns.ProductType(int(t))
        colour = ns.PackagingColor(int(c))
        products = [
            ns.Product(id=order_id * 100 + i, product_type=ptype, packaging_color=colour, order_id=order_id)
            for i in range(int(n))
        ]
        order = ns.Order(id=order_id, products=products, arrival_time=sim.env.now)
        sim.orders.append(order)
        sim.pickup_station.add_order(order)
    # registry of every Tray object of this episode, so delivered/lost trays stay enumerable
    sim._all_trays = list(sim.pickup_station.trays_at_station)
    return sim.get_observations(), infos
