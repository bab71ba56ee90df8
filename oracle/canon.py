"""Canonical integer state record S (TEST INFRASTRUCTURE ONLY).

numpy mirror of ``FjspCanonState`` in ``include/fjsp_b200.h`` plus the exporter that walks the
LIVE reference object graph (``/root/reference``: FJSPSimulation.py, agents/*.py, models/*.py)
into that record, so the reference, the C restatement and the CUDA path can be compared bit for
bit.
"""
from __future__ import annotations

import numpy as np

MAX_ORDERS = 32
MAXQ = 64
PS_READY = 256
MAXPQ = 256

MACHINE_DT = np.dtype([
    ("is_busy", "<i4"), ("current_tray", "<i4"), ("progress_done", "<i4"),
    ("queue_n", "<i4"), ("queue", "<i4", (MAXQ,)),
    ("ready_n", "<i4"), ("ready", "<i4", (MAXQ,)),
])
PACK_DT = np.dtype([
    ("is_busy", "<i4"), ("current_product", "<i4"), ("progress_L", "<i4"),
    ("products_completed", "<i4"), ("users", "<i4"),
    ("queue_n", "<i4"), ("queue", "<i4", (MAXPQ,)),
])
CANON_DT = np.dtype([
    ("current_step", "<i4"), ("num_orders", "<i4"), ("fault", "<i4"),
    ("agv_row", "<i4"), ("agv_col", "<i4"), ("agv_carry", "<i4"), ("agv_is_moving", "<i4"),
    ("ps_order_queue_len", "<i4"), ("ps_current_order", "<i4"), ("ps_product_idx", "<i4"),
    ("ps_current_tray", "<i4"), ("ps_trays_at_station", "<i4"),
    ("ps_ready_n", "<i4"), ("ps_ready", "<i4", (PS_READY,)),
    ("machine", MACHINE_DT, (2,)),
    ("storage_n", "<i4"), ("storage", "<i4", (MAXQ,)),
    ("pack", PACK_DT, (4,)),
    ("processed_mask", "<i4", (MAX_ORDERS,)),
    ("packaged_mask", "<i4", (MAX_ORDERS,)),
    ("order_complete", "<i4", (MAX_ORDERS,)),
    ("order_completion_step", "<i4", (MAX_ORDERS,)),
    ("total_products_packaged", "<i4"), ("completed_orders", "<i4"),
])

PACK_IDS = ["packaging_blue_1", "packaging_blue_2", "packaging_red", "packaging_green"]


def empty_canon() -> np.ndarray:
    s = np.zeros((), dtype=CANON_DT)
    s["agv_carry"] = -1
    s["ps_current_order"] = -1
    s["ps_current_tray"] = -1
    s["ps_ready"][...] = -1
    for m in range(2):
        s["machine"][m]["current_tray"] = -1
        s["machine"][m]["queue"][...] = -1
        s["machine"][m]["ready"][...] = -1
    s["storage"][...] = -1
    for p in range(4):
        s["pack"][p]["current_product"] = -1
        s["pack"][p]["queue"][...] = -1
    s["order_completion_step"][...] = -1
    return s


def tray_entry(tray) -> int:
    """tray_id | order<<16 | first<<22 | count<<26 for a reference ``Tray`` (models/Tray.py:7-58)."""
    if tray is None:
        return -1
    prods = tray.products
    if not prods:
        return int(tray.id)
    return int(tray.id) | (int(prods[0].order_id) << 16) | ((int(prods[0].id) % 100) << 22) | (len(prods) << 26)


def _fill(dst_n_field, dst_arr, items):
    n = len(items)
    assert n <= dst_arr.shape[0], "canonical record overflow (%d > %d)" % (n, dst_arr.shape[0])
    dst_arr[:n] = items
    return n


def export_reference(sim) -> np.ndarray:
    """Walk a live reference ``FJSPSimulation`` into the canonical record."""
    s = empty_canon()
    step_size = sim.config["step_size"]
    s["current_step"] = sim.current_step
    s["num_orders"] = len(sim.orders)
    agv = sim.agv
    s["agv_row"], s["agv_col"] = int(agv.position[0]), int(agv.position[1])
    s["agv_carry"] = tray_entry(agv.carrying_tray)
    s["agv_is_moving"] = int(bool(agv.is_moving))
    ps = sim.pickup_station
    s["ps_order_queue_len"] = len(ps.order_queue)
    s["ps_current_order"] = ps.current_order.id if ps.current_order else -1
    s["ps_product_idx"] = ps.current_order_product_idx
    s["ps_current_tray"] = tray_entry(ps.current_tray)
    s["ps_trays_at_station"] = len(ps.trays_at_station)
    s["ps_ready_n"] = _fill(None, s["ps_ready"], [tray_entry(t) for t in ps.ready_trays])
    for mi, m in enumerate((sim.small_machine, sim.big_machine)):
        rec = s["machine"][mi]
        rec["is_busy"] = int(bool(m.is_busy))
        rec["current_tray"] = tray_entry(m.current_tray)
        assert m.processing_progress in (0.0, 1.0)
        rec["progress_done"] = int(m.processing_progress == 1.0)
        rec["queue_n"] = _fill(None, rec["queue"], [tray_entry(t) for t in m.tray_queue])
        rec["ready_n"] = _fill(None, rec["ready"], [tray_entry(t) for t in m.ready_trays])
    s["storage_n"] = _fill(None, s["storage"], [tray_entry(t) for t in sim.storage.trays])
    for pi, pid in enumerate(PACK_IDS):
        st = sim.packaging_stations[pid]
        rec = s["pack"][pi]
        rec["is_busy"] = int(bool(st.is_busy))
        rec["current_product"] = st.current_product.id if st.current_product is not None else -1
        rec["progress_L"] = int(round(100.0 / st.processing_progress)) if st.processing_progress else 0
        rec["products_completed"] = st.products_completed
        rec["users"] = st.resource.count
        rec["queue_n"] = _fill(None, rec["queue"], [p.id for p in st.product_queue])
    for o in sim.orders:
        pm = 0
        km = 0
        for i, p in enumerate(o.products):
            pm |= int(bool(p.is_processed)) << i
            km |= int(bool(p.is_packaged)) << i
        s["processed_mask"][o.id] = pm
        s["packaged_mask"][o.id] = km
        s["order_complete"][o.id] = int(bool(o.is_complete))
        if o.completion_time is not None:
            k = o.completion_time / step_size - 1
            assert k == int(k)
            s["order_completion_step"][o.id] = int(k)
    s["total_products_packaged"] = sim.total_products_packaged
    s["completed_orders"] = len(sim.completed_orders)
    return s


def tray_entry_long(tray) -> int:
    """FJSP_TRAY_ENTRY_LONG: tray_id | order<<12 | first<<24 | count<<28 (long order streams, include/fjsp_b200.h)."""
    if tray is None:
        return -1
    prods = tray.products
    if not prods:
        return int(tray.id)
    return int(tray.id) | (int(prods[0].order_id) << 12) | ((int(prods[0].id) % 100) << 24) | (len(prods) << 28)


def export_reference_long(sim) -> np.ndarray:
    """The canonical record of a live reference simulation under the conventions of the port's LONG layout
    (include/fjsp_b200.h fjsp_export_orders): tray entries are FJSP_TRAY_ENTRY_LONG; the per-order arrays describe the 32
    most recently popped orders; an order reads "complete, all bits, step -1" from the second step after its completion
    (the port has given its slot back by then); list fields are cut at the record's capacity with the true length kept."""
    s = empty_canon()
    step_size = sim.config["step_size"]
    te = tray_entry_long
    s["current_step"] = sim.current_step
    s["num_orders"] = len(sim.orders)
    agv = sim.agv
    s["agv_row"], s["agv_col"] = int(agv.position[0]), int(agv.position[1])
    s["agv_carry"] = te(agv.carrying_tray)
    s["agv_is_moving"] = int(bool(agv.is_moving))
    ps = sim.pickup_station
    s["ps_order_queue_len"] = len(ps.order_queue)
    s["ps_current_order"] = ps.current_order.id if ps.current_order else -1
    s["ps_product_idx"] = ps.current_order_product_idx
    s["ps_current_tray"] = te(ps.current_tray)
    s["ps_trays_at_station"] = len(ps.trays_at_station)
    ready = [te(t) for t in ps.ready_trays]
    s["ps_ready"][:min(len(ready), PS_READY)] = ready[:PS_READY]
    s["ps_ready_n"] = len(ready)
    for mi, m in enumerate((sim.small_machine, sim.big_machine)):
        rec = s["machine"][mi]
        rec["is_busy"] = int(bool(m.is_busy))
        rec["current_tray"] = te(m.current_tray)
        rec["progress_done"] = int(m.processing_progress == 1.0)
        rec["queue_n"] = _fill(None, rec["queue"], [te(t) for t in m.tray_queue])
        rec["ready_n"] = _fill(None, rec["ready"], [te(t) for t in m.ready_trays])
    s["storage_n"] = _fill(None, s["storage"], [te(t) for t in sim.storage.trays])
    for pi, pid in enumerate(PACK_IDS):
        st = sim.packaging_stations[pid]
        rec = s["pack"][pi]
        rec["is_busy"] = int(bool(st.is_busy))
        rec["current_product"] = st.current_product.id if st.current_product is not None else -1
        rec["progress_L"] = int(round(100.0 / st.processing_progress)) if st.processing_progress else 0
        rec["products_completed"] = st.products_completed
        rec["users"] = st.resource.count
        rec["queue_n"] = _fill(None, rec["queue"], [p.id for p in st.product_queue])
    popped = len(sim.orders) - len(ps.order_queue)
    base = max(0, popped - MAX_ORDERS)
    for i in range(MAX_ORDERS):
        oid = base + i
        if oid >= popped:
            continue
        o = sim.orders[oid]
        k = None if o.completion_time is None else int(o.completion_time / step_size - 1)
        if o.is_complete and k < sim.current_step - 1:
            s["processed_mask"][i] = s["packaged_mask"][i] = 0x1FF
            s["order_complete"][i] = 1
            continue
        pm = km = 0
        for j, p in enumerate(o.products):
            pm |= int(bool(p.is_processed)) << j
            km |= int(bool(p.is_packaged)) << j
        s["processed_mask"][i], s["packaged_mask"][i] = pm, km
        s["order_complete"][i] = int(bool(o.is_complete))
        if k is not None:
            s["order_completion_step"][i] = k
    s["total_products_packaged"] = sim.total_products_packaged
    s["completed_orders"] = len(sim.completed_orders)
    return s


def diff(a: np.ndarray, b: np.ndarray, prefix: str = "") -> list:
    """Field-wise differences between two canonical records (empty list = bit-exact)."""
    out = []
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if a.dtype[name].names:
            if x.shape:
                for i in range(x.shape[0]):
                    out += diff(x[i], y[i], "%s%s[%d]." % (prefix, name, i))
            else:
                out += diff(x, y, prefix + name + ".")
        elif not np.array_equal(x, y):
            out.append("%s%s: %s != %s" % (prefix, name, np.asarray(x).tolist(), np.asarray(y).tolist()))
    return out


# observation layout O: a2c._flatten_obs order (a2c.py:137-151) over the 8 agents
AGENT_IDS = [
    "pickup_station", "agv", "small_machine", "big_machine",
    "packaging_blue_1", "packaging_blue_2", "packaging_red", "packaging_green",
]
MASK_OFFSETS = [0, 3, 11, 14, 17, 20, 23, 26, 29]


def flatten_reference_obs(obs: dict):
    """Reference observation dicts -> (float32[38] in layout O, int8[32] masks)."""
    vals = []
    masks = np.zeros(32, dtype=np.int8)
    for ai, aid in enumerate(AGENT_IDS):
        o = obs[aid]
        for key in sorted(o.keys()):
            if key == "action_mask":
                continue
            vals.extend(np.asarray(o[key]).flatten().tolist())
        m = np.asarray(o["action_mask"])
        masks[MASK_OFFSETS[ai]:MASK_OFFSETS[ai] + m.shape[0]] = m
    v = np.array(vals, dtype=np.float32)
    assert v.shape == (38,)
    return v, masks
