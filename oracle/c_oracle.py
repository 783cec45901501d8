"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/mbe_oracle_c.c (the compiled CPU restatement).

Only ``tests/`` and the ``cpu_baseline`` leg of ``bench.py`` may import this module.  ``build()`` compiles
the C file with gcc into ``oracle/_build/libmbe_oracle_c.so`` (git-ignored, travels to the GPU box with
the snapshot like the product's own ``.so``)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mbe_oracle_c.c")
LIB = os.path.join(HERE, "_build", "libmbe_oracle_c.so")


class CParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("width", "height", "velocity", "snr_tr", "noise", "ue_height",
                                          "util_lower", "util_upper", "w1", "w2", "w3")] + [
        ("ep_time", C.c_int32), ("handler", C.c_int32), ("scheduler", C.c_int32), ("pad_", C.c_int32)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mbo_fork_step.restype = None
        _lib.mbo_gym_step.restype = None
        _lib.mbo_set_tables.restype = None
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class CEnvBatch:
    """E independent envs stepped by the C library.  ``p``: oracle.mbe_oracle.Params; ``bs_over``: per-BS
    overrides like the fixtures hold (keys bw / freq / tx / bs_height)."""

    def __init__(self, p, bs_xy, num_envs, num_ues, handler="central", bs_over=None, nbs=None, pinned_tables=False):
        sched = {"resource_fair": 0, "proportional_fair": 1, "rate_fair": 2}[p.scheduler]
        self.lib = load()
        bs_xy = np.asarray(bs_xy, dtype=np.int32)
        self.per_env = bs_xy.ndim == 3
        self.E, self.U, self.B = int(num_envs), int(num_ues), int(bs_xy.shape[-2])
        assert self.B <= self.lib.mbo_max_b() and self.U <= self.lib.mbo_max_u()
        self.bs_xy = np.ascontiguousarray(bs_xy)
        self.nbs = None if nbs is None else np.ascontiguousarray(nbs, dtype=np.int32)
        w1, w2, w3 = p.util_coeffs
        self.ma = handler != "central"
        self.cp = CParams(p.width, p.height, p.velocity, p.snr_tr, p.noise, p.ue_height, p.util_lower, p.util_upper,
                          w1, w2, w3, int(p.ep_time), int(self.ma), sched, 0)
        par = np.tile(np.array([p.bw, p.freq, p.tx, p.bs_height], dtype=np.float64), (self.B, 1))
        for b, over in enumerate(bs_over or []):
            for k, v in (over or {}).items():
                par[b, {"bw": 0, "freq": 1, "tx": 2, "bs_height": 3}[k]] = v
        self.bs_par = np.ascontiguousarray(par)
        E, U, B = self.E, self.U, self.B
        self.F = (4 if self.ma else 2) * B + 1
        self.pos = np.zeros((E, U, 2), dtype=np.int32)
        self.wp = np.full((E, U, 2), -1, dtype=np.int32)
        self.t = np.zeros(E, dtype=np.int32)
        self.conn = np.zeros((E, U, B), dtype=np.uint8)
        self.assoc = np.full((E, U), -1, dtype=np.int32)
        self.rate = np.zeros((E, U), dtype=np.float64)
        self.util = np.zeros((E, U), dtype=np.float64)
        self.reward = np.zeros((E, U) if self.ma else (E,), dtype=np.float64)
        self.done = np.zeros(E, dtype=np.uint8)
        self.drew = np.zeros((E, U), dtype=np.int32)
        self.metrics = np.zeros((E, 4), dtype=np.float64)
        self.bs_util = np.zeros((E, B), dtype=np.float64)
        self.obs = np.zeros((E, U, self.F), dtype=np.float32)
        self.tables = None
        if pinned_tables:
            self.tables = self._numpy_tables(p, bs_over)

    def _numpy_tables(self, p, bs_over):
        """snr / Shannon rate per integer squared distance through the numpy scalar chain of
        oracle/mbe_oracle.py (the reference's operation order: power_loss -> calculateSNR -> datarate,
        channels.py:132-146, 24-27, 78-83), one table per BS parameter set.  With them the C steps return the
        numpy rates bit for bit (see mbo_set_tables in the C file)."""
        import dataclasses
        import math

        from oracle import mbe_oracle as orc

        tab_len = int(p.width) ** 2 + int(p.height) ** 2 + 1
        sets = [dict(o or {}) for o in (bs_over or [])] if any(bs_over or []) else [{}]
        n_tabs = len(sets) if len(sets) > 1 else 1
        if n_tabs > 1:
            assert n_tabs == self.B
        snr = np.zeros((n_tabs, tab_len), dtype=np.float64)
        rate = np.zeros((n_tabs, tab_len), dtype=np.float64)
        with np.errstate(divide="ignore", over="ignore"):
            for i in range(n_tabs):
                pp = dataclasses.replace(p, **sets[i]) if sets[i] else p
                for d2 in range(tab_len):
                    s_ = orc.snr_of(pp, math.sqrt(d2))
                    snr[i, d2] = s_
                    rate[i, d2] = orc.datarate_of(pp, s_)
        return np.ascontiguousarray(snr), np.ascontiguousarray(rate), n_tabs, tab_len

    def _tables_on(self):
        if self.tables is None:
            self.lib.mbo_set_tables(None, None, 0, 0)
        else:
            snr, rate, n_tabs, tab_len = self.tables
            self.lib.mbo_set_tables(_p(snr, C.c_double), _p(rate, C.c_double), n_tabs, tab_len)

    def reset(self, init_pos):
        self.pos[:] = np.asarray(init_pos, dtype=np.int32)
        self.wp[:] = -1
        self.t[:] = 0
        self.conn[:] = 0

    def step_fork(self, new_wp):
        new_wp = np.ascontiguousarray(np.broadcast_to(np.asarray(new_wp, dtype=np.int32), self.pos.shape))
        self._tables_on()
        self.lib.mbo_fork_step(C.byref(self.cp), self.E, self.U, self.B, _p(self.bs_par, C.c_double),
                               _p(self.bs_xy, C.c_int32), int(self.per_env), _p(self.nbs, C.c_int32),
                               _p(self.pos, C.c_int32), _p(self.wp, C.c_int32), _p(new_wp, C.c_int32),
                               _p(self.drew, C.c_int32), _p(self.t, C.c_int32), _p(self.assoc, C.c_int32),
                               _p(self.rate, C.c_double), _p(self.util, C.c_double), _p(self.done, C.c_uint8),
                               _p(self.metrics, C.c_double))

    def step_gym(self, actions, new_wp):
        if self.per_env:
            raise NotImplementedError("GYM step of the C restatement: shared layouts only")
        new_wp = np.ascontiguousarray(np.broadcast_to(np.asarray(new_wp, dtype=np.int32), self.pos.shape))
        acts = np.ascontiguousarray(np.broadcast_to(np.asarray(actions, dtype=np.int32), (self.E, self.U)))
        self._tables_on()
        self.lib.mbo_gym_step(C.byref(self.cp), self.E, self.U, self.B, _p(self.bs_par, C.c_double),
                              _p(self.bs_xy, C.c_int32), _p(self.pos, C.c_int32), _p(self.wp, C.c_int32),
                              _p(new_wp, C.c_int32), _p(self.drew, C.c_int32), _p(self.t, C.c_int32),
                              _p(self.conn, C.c_uint8), _p(acts, C.c_int32), _p(self.rate, C.c_double),
                              _p(self.util, C.c_double), _p(self.reward, C.c_double), _p(self.done, C.c_uint8),
                              _p(self.obs, C.c_float), _p(self.bs_util, C.c_double), _p(self.metrics, C.c_double))


if __name__ == "__main__":
    print(build(force=True))
