"""Test helper: drives oracle/mbe_oracle.py exactly like the CUDA env is driven (Philox draws,
resets, autoreset, random BS layouts) so that outputs can be compared tensor by tensor."""
from __future__ import annotations

import numpy as np

from oracle import mbe_oracle as orc


def params_from_env(env) -> orc.Params:
    cfg, p = env.config, env.plan
    return orc.Params(
        width=float(cfg["width"]), height=float(cfg["height"]), ep_time=p.ep_time,
        bw=cfg["bs"]["bw"], freq=cfg["bs"]["freq"], tx=cfg["bs"]["tx"], bs_height=cfg["bs"]["height"],
        velocity=cfg["ue"]["velocity"], snr_tr=cfg["ue"]["snr_tr"], noise=cfg["ue"]["noise"],
        ue_height=cfg["ue"]["height"], util_lower=cfg["utility_params"]["lower"],
        util_upper=cfg["utility_params"]["upper"], util_coeffs=tuple(cfg["utility_params"]["coeffs"]),
        scheduler={0: "resource_fair", 1: "proportional_fair", 2: "rate_fair"}[p.scheduler],
    )


class Mirror:
    def __init__(self, env):
        self.p = params_from_env(env)
        pl = env.plan
        self.E, self.U, self.B = pl.num_envs, pl.num_ues, pl.num_bs
        self.seed, self.off = pl.seed, pl.env_offset
        self.gym = env.config["mode"] == "gym"
        self.handler = "ma" if pl.handler == 1 else "central"
        self.autoreset = pl.autoreset
        self.rre = pl.reset_rng_episode
        self.bs_random = pl.bs_random if pl.bs_random[1] > 0 else None
        self.gid = self.off + np.arange(self.E)
        # UE trajectories shared by all envs (MBE_FLAG_SHARED_TRAJECTORY): env id 0 in their counters
        self.traj_gid = np.zeros_like(self.gid) if getattr(pl, "shared_trajectory", False) else self.gid
        self.ue = np.arange(self.U)
        self.episode = np.full(self.E, -1, dtype=np.int64)
        self.t = np.zeros(self.E, dtype=np.int64)
        self.pos = np.zeros((self.E, self.U, 2), dtype=np.int64)
        self.wp = np.full((self.E, self.U, 2), -1, dtype=np.int64)
        self.conn = np.zeros((self.E, self.U, self.B), dtype=bool)
        if self.bs_random:
            self.bs = np.zeros((self.E, self.B, 2), dtype=np.int64)
            self.nbs = np.full(self.E, self.B, dtype=np.int64)
        else:
            self.bs = pl.bs_xy.astype(np.int64)
            self.nbs = None

    def _salt(self):
        return np.zeros_like(self.episode) if self.rre else self.episode

    def _reinit(self, sel):
        sel = np.asarray(sel, dtype=bool)
        self.episode = np.where(sel, self.episode + 1, self.episode)
        self.t = np.where(sel, 0, self.t)
        x, y = orc.philox_point(self.seed, self.traj_gid[:, None], self.ue[None, :], 0, orc.PURPOSE_INITPOS,
                                self._salt()[:, None], self.p.width, self.p.height)
        init = np.stack([x, y], axis=-1)
        self.pos = np.where(sel[:, None, None], init, self.pos)
        self.wp = np.where(sel[:, None, None], -1, self.wp)
        self.conn = np.where(sel[:, None, None], False, self.conn)
        if self.bs_random:
            n = orc.philox_bs_count(self.seed, self.gid, self.episode, *self.bs_random)
            bx, by = orc.philox_point(self.seed, self.gid[:, None], np.arange(self.B)[None, :], 0,
                                      orc.PURPOSE_BSLAYOUT, self.episode[:, None], self.p.width, self.p.height)
            live = np.arange(self.B)[None, :] < n[:, None]
            bs = np.stack([np.where(live, bx, 0), np.where(live, by, 0)], axis=-1)
            self.bs = np.where(sel[:, None, None], bs, self.bs)
            self.nbs = np.where(sel, n, self.nbs)

    def reset(self, mask=None):
        sel = np.ones(self.E, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        self._reinit(sel)
        if self.gym:
            return self.observe_fresh()
        return None

    def _bs_for_oracle(self):
        if self.nbs is None:
            return self.bs
        # absent BS slots: push them far away so they are never eligible and never the max SNR
        live = np.arange(self.B)[None, :] < self.nbs[:, None]
        return np.where(live[:, :, None], self.bs, 10**7)

    def observe_fresh(self):
        obs = orc.batch_observe(self.p, self.pos, self._bs_for_oracle(), self.conn, None, self.handler)
        return self._mask_absent(obs)

    def _mask_absent(self, obs):
        if self.nbs is None:
            return obs
        B = self.B
        live = (np.arange(B)[None, :] < self.nbs[:, None])[:, None, :]
        obs = obs.copy()
        obs[:, :, B:2 * B] = np.where(live, obs[:, :, B:2 * B], 0)
        if self.handler != "central":
            pass  # bcast of an absent BS is idle (-1) and its count 0 on both sides
        return obs

    def new_wp(self):
        x, y = orc.philox_point(self.seed, self.traj_gid[:, None], self.ue[None, :], self.t[:, None],
                                orc.PURPOSE_WAYPOINT, self._salt()[:, None], self.p.width, self.p.height)
        return np.stack([x, y], axis=-1)

    def step_fork(self, new_wp=None):
        if new_wp is None:
            new_wp = self.new_wp()
        out = orc.batch_step_fork(self.p, self.pos, self.wp, new_wp, self._bs_for_oracle(), self.t, self.nbs)
        self.pos, self.wp, self.t = out["pos"], out["wp"], out["t"]
        if self.autoreset:
            self._reinit(out["done"])
        out["pos_after"] = self.pos.copy()
        return out

    def step_gym(self, actions, new_wp=None):
        if new_wp is None:
            new_wp = self.new_wp()
        a = np.asarray(actions)
        if self.nbs is not None:
            a = np.where(a > self.nbs[:, None], 0, a)
        out = orc.batch_step_gym(self.p, self.pos, self.wp, new_wp, self._bs_for_oracle(), self.conn, a,
                                 self.t, self.handler)
        self.pos, self.wp, self.t, self.conn = out["pos"], out["wp"], out["t"], out["conn"]
        out["obs"] = self._mask_absent(out["obs"])
        if self.autoreset and out["done"].any():
            self._reinit(out["done"])
            fresh = self.observe_fresh()
            out["obs"] = np.where(out["done"][:, None, None], fresh, out["obs"])
        out["pos_after"] = self.pos.copy()
        out["conn_after"] = self.conn.copy()
        return out


def conn_bool_from_words(words, B):
    """GPU connection words (int32 [E,U] or [E,U,MW]) -> bool [E,U,B]."""
    w = np.asarray(words).astype(np.int64) & 0xFFFFFFFF
    if w.ndim == 2:
        w = w[:, :, None]
    b = np.arange(B)
    return ((w[:, :, b // 32] >> (b % 32)[None, None, :]) & 1).astype(bool)


def conn_bits(conn_bool):
    """[E,U,B] bool -> uint32-as-int64 bitmask [E,U]."""
    B = conn_bool.shape[2]
    return (conn_bool.astype(np.int64) << np.arange(B)[None, None, :]).sum(axis=2)
