"""Dump / export formats of the reference, fed from the GPU-resident FORK env.

The fork's product is a data set: per step four JSON files (``save_layout_and_data_rates``,
reference mobile_env/core/base.py:298-349) and per epoch three CSV files (``save_epoch_data``,
base.py:351-404) plus the station positions (``save_base_station_positions``,
scenarios/custom.py:79-85).  The notebooks (GNN.ipynb, analysisData.ipynb,
chooseBaseStation.ipynb) consume exactly these files.

Here one *env* of the batch plays the role of one *epoch* of the reference loop
(collectData2.ipynb cell 4: ``for epoch: reset(); for step: env.step(epoch, step)``).
``ReferenceDumpWriter`` copies the few tensors a dump needs to pinned host memory on a side
stream after every step (the step path itself is untouched) and a writer thread formats the
files.  Formatting is byte-compatible with the reference, including the quirks of its types
(``np.float64(..)`` / ``np.int64(..)`` reprs inside the CSV lists, python ``0.0`` for unconnected
UEs, python ints on the step a UE snaps onto its waypoint).  One documented difference: inside
``data_rates_*.json`` the reference lists the UEs of a BS in Python-set order (address
dependent); here they are listed by (bs_id, ue_id).

The per-UE QoE is recomputed on the host in FP64 from the exact FP64 rate with the reference's own
formula (utilities.py:44-55), so the two-decimal values match the reference's exactly.
"""
from __future__ import annotations

import json
import os
import queue
import threading
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

STEP_DIRS = {
    "stations": ("collectData", "BaseStationPosition", "stations_info_{e}_{s}.json"),
    "users": ("collectData", "UserEquipmentPosition", "user_positions_{e}_{s}.json"),
    "rates": ("collectData", "DataRate", "data_rates_{e}_{s}.json"),
    "qoe": ("collectData", "UserQoE", "user_qoe_{e}_{s}.json"),
}
EPOCH_DIRS = {
    "stations": ("collectData2", "BaseStationPosition", "stations_{e}.json"),
    "rates": ("collectData2", "DataRate", "datarates_{e}.csv"),
    "traj": ("collectData2", "UserEquipmentPosition", "user_positions_{e}.csv"),
    "qoe": ("collectData2", "UserQoE", "user_qoe_{e}.csv"),
}


def scaled_utility_fp64(rate: float, lower, upper, coeffs):
    """calculateUtility + scaleUtility (utilities.py:44-55) with the reference's types:
    np.float64 for a positive rate, python numbers for the lower bound."""
    w1, w2, w3 = coeffs
    if rate <= 0.0:
        u = lower
    else:
        u = np.clip(w1 * np.log(w2 + rate) / np.log(w3), lower, upper)
    return 2 * (u - lower) / (upper - lower) - 1


def format_step_files(e: int, s: int, bs_xy, pos, assoc, rate, util_params) -> Dict[str, str]:
    """The four per-step JSON files of base.py:298-349 for one env. Returns {relative path: text}."""
    lower, upper, coeffs = util_params
    out = {}
    stations = [{"bs_id": b, "x": round(float(int(x)), 2), "y": round(float(int(y)), 2)} for b, (x, y) in enumerate(bs_xy)]
    users = [{"ue_id": u, "x": round(float(x), 2), "y": round(float(y), 2)} for u, (x, y) in enumerate(pos)]
    pairs = sorted((int(b), u) for u, b in enumerate(assoc) if b >= 0)
    rates = [{"ue_id": u, "bs_id": b, "data_rate": round(float(rate[u]), 2)} for b, u in pairs]
    qoe = []
    for u in range(len(pos)):
        r = np.float64(rate[u]) if assoc[u] >= 0 else 0.0
        qoe.append({"ue_id": u, "qoe": round(scaled_utility_fp64(r, lower, upper, coeffs), 2)})
    for key, obj in (("stations", stations), ("users", users), ("rates", rates), ("qoe", qoe)):
        out[os.path.join(*STEP_DIRS[key]).format(e=e, s=s)] = json.dumps(obj, indent=4)
    return out


def format_epoch_files(e: int, bs_xy, pos_steps, arrived_steps, assoc_steps, rate_steps, util_params) -> Dict[str, str]:
    """The per-epoch files (base.py:351-404, custom.py:79-85) for one env from its whole episode:
    pos_steps [T,U,2], arrived_steps [T,U] bool, assoc_steps [T,U], rate_steps [T,U]."""
    import pandas as pd

    lower, upper, coeffs = util_params
    T, U = len(pos_steps), len(pos_steps[0])
    out = {}
    positions = {b: (int(x), int(y)) for b, (x, y) in enumerate(bs_xy)}
    out[os.path.join(*EPOCH_DIRS["stations"]).format(e=e)] = json.dumps(positions)
    rates, traj, qoes = [], [], []
    for u in range(U):
        r_list, t_list, q_list = [], [], []
        for t in range(T):
            connected = assoc_steps[t][u] >= 0
            # allUserDataRates.get(ue, 0.0): np.float64 for a connected UE, python 0.0 otherwise (base.py:265)
            r = np.float64(rate_steps[t][u]) if connected else 0.0
            r_list.append(round(r, 2))
            x, y = pos_steps[t][u]
            # movement.py:54-56 returns the popped waypoint (python ints) on arrival, else np.int64s (60-62)
            t_list.append((int(x), int(y)) if arrived_steps[t][u] else (np.int64(x), np.int64(y)))
            q_list.append(round(scaled_utility_fp64(r, lower, upper, coeffs), 2))
        rates.append({"User ID": u, "Data Rates": r_list})
        traj.append({"User ID": u, "Trajectory": t_list})
        qoes.append({"User ID": u, "QoE": q_list})
    for key, rows in (("rates", rates), ("traj", traj), ("qoe", qoes)):
        out[os.path.join(*EPOCH_DIRS[key]).format(e=e)] = pd.DataFrame(rows).to_csv(index=False)
    return out


class ReferenceDumpWriter:
    """Asynchronous exporter for a FORK-mode batched env.

    >>> w = ReferenceDumpWriter(env, "/data/run1", envs=range(1000))
    >>> env.reset(); w.begin_episode()
    >>> for s in range(20):
    ...     env.step(0, s); w.after_step(s)
    >>> w.end_episode(); w.close()
    """

    def __init__(self, env, root: str, envs: Optional[Iterable[int]] = None, per_step: bool = True,
                 epoch_offset: int = 0, depth: int = 4):
        if env.assoc is None:
            raise ValueError("ReferenceDumpWriter needs a FORK-mode env (the dumps are the fork's format)")
        self.env, self.root, self.per_step = env, root, per_step
        self.sel = torch.as_tensor(list(range(env.num_envs)) if envs is None else list(envs), device=env.device)
        self.ids = [int(i) + epoch_offset for i in self.sel.tolist()]
        up = env.config["utility_params"]
        self.util_params = (up["lower"], up["upper"], tuple(up["coeffs"]))
        self.copy_stream = torch.cuda.Stream(device=env.device)
        n, U = len(self.ids), env.plan.num_ues
        self.slots = [
            {
                "pos": torch.empty(n, U, 2, dtype=torch.int16).pin_memory(),
                "wp": torch.empty(n, U, 2, dtype=torch.int16).pin_memory(),
                "assoc": torch.empty(n, U, dtype=torch.int32).pin_memory(),
                "rate": torch.empty(n, U, dtype=torch.float64).pin_memory(),
                "event": torch.cuda.Event(),
                "free": threading.Event(),
            }
            for _ in range(depth)
        ]
        for sl in self.slots:
            sl["free"].set()
        self.q: "queue.Queue" = queue.Queue()
        self.err: List[BaseException] = []
        self.thread = threading.Thread(target=self._writer, daemon=True)
        self.thread.start()
        self.step_count = 0
        self.bs = None
        self.history = None

    # -- producer side (caller's thread) --------------------------------------------------
    def begin_episode(self):
        torch.cuda.current_stream(self.env.device).synchronize()
        env = self.env
        if env.nbs is None:
            bs = env.bs_xy.cpu().numpy()
            self.bs = [bs for _ in self.ids]
        else:
            bs_all = env.bs_xy[self.sel].cpu().numpy()
            nbs = env.nbs[self.sel].cpu().numpy()
            self.bs = [bs_all[i, : nbs[i]] for i in range(len(self.ids))]
        self.history = {"pos": [], "arrived": [], "assoc": [], "rate": []}
        self.step_count = 0

    def after_step(self, step: int):
        env = self.env
        sl = self.slots[self.step_count % len(self.slots)]
        sl["free"].wait()
        sl["free"].clear()
        self.copy_stream.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(self.copy_stream):
            sl["pos"].copy_(env.pos[self.sel], non_blocking=True)
            sl["wp"].copy_(env.wp[self.sel], non_blocking=True)
            sl["assoc"].copy_(env.assoc[self.sel], non_blocking=True)
            sl["rate"].copy_(env.rate[self.sel], non_blocking=True)
            sl["event"].record(self.copy_stream)
        # the next step must not overwrite the tensors before the gather kernels above have read them
        torch.cuda.current_stream(env.device).wait_stream(self.copy_stream)
        self.q.put(("step", step, sl))
        self.step_count += 1

    def write_rollout(self, series):
        """A whole episode at once: ``series = env.rollout(T, record=("pos", "wp", "assoc", "rate"))``
        (one launch for the fork's own scenario) instead of ``after_step`` per step.  Call
        ``begin_episode()`` before the rollout (it snapshots the BS layouts) and ``end_episode()`` after."""
        missing = [n for n in ("pos", "wp", "assoc", "rate") if n not in series]
        if missing:
            raise ValueError(f"write_rollout needs the series {missing}: record=('pos', 'wp', 'assoc', 'rate')")
        host = {n: series[n][:, self.sel].cpu().numpy() for n in ("pos", "wp", "assoc", "rate")}
        self.q.put(("rollout", self.step_count, host))
        self.step_count += host["pos"].shape[0]

    def end_episode(self):
        self.q.put(("epoch", None, None))
        self.q.join()
        self._raise()

    def close(self):
        self.q.put(("stop", None, None))
        self.thread.join()
        self._raise()

    def _raise(self):
        if self.err:
            raise self.err[0]

    # -- consumer side (writer thread) -----------------------------------------------------
    def _write(self, files: Dict[str, str]):
        for rel, text in files.items():
            path = os.path.join(self.root, rel)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(text)

    def _writer(self):
        while True:
            kind, step, sl = self.q.get()
            try:
                if kind == "stop":
                    return
                if kind == "step":
                    sl["event"].synchronize()
                    pos, wp = sl["pos"].numpy().copy(), sl["wp"].numpy().copy()
                    assoc, rate = sl["assoc"].numpy().copy(), sl["rate"].numpy().copy()
                    sl["free"].set()
                    self.history["pos"].append(pos)
                    self.history["arrived"].append(wp[:, :, 0] < 0)
                    self.history["assoc"].append(assoc)
                    self.history["rate"].append(rate)
                    if self.per_step:
                        for i, e in enumerate(self.ids):
                            self._write(format_step_files(e, step, self.bs[i], pos[i], assoc[i], rate[i], self.util_params))
                elif kind == "rollout":
                    host, sl = sl, None
                    for t in range(host["pos"].shape[0]):
                        pos, assoc, rate = host["pos"][t], host["assoc"][t], host["rate"][t]
                        self.history["pos"].append(pos)
                        self.history["arrived"].append(host["wp"][t][:, :, 0] < 0)
                        self.history["assoc"].append(assoc)
                        self.history["rate"].append(rate)
                        if self.per_step:
                            for i, e in enumerate(self.ids):
                                self._write(format_step_files(e, step + t, self.bs[i], pos[i], assoc[i], rate[i],
                                                              self.util_params))
                elif kind == "epoch":
                    h = self.history
                    for i, e in enumerate(self.ids):
                        self._write(format_epoch_files(
                            e, self.bs[i], [p[i] for p in h["pos"]], [a[i] for a in h["arrived"]],
                            [a[i] for a in h["assoc"]], [r[i] for r in h["rate"]], self.util_params))
            except BaseException as exc:  # surfaced by end_episode()/close()
                self.err.append(exc)
                if sl is not None:
                    sl["free"].set()
            finally:
                self.q.task_done()
