"""Reference-named host view of one environment (slow path, for inspection and export)."""
from __future__ import annotations

from collections import Counter, defaultdict

from .entities import BaseStation, UserEquipment


class EnvView:
    """Snapshot with the attribute names of reference base.py:69-79."""

    def __init__(self, env, e: int):
        p = env.plan
        cfg = env.config
        self.channelModel, self.schedulerModel, self.utilityModel = env.channelModel, env.schedulerModel, env.utilityModel
        pos = env.pos[e].cpu().tolist()
        if env.nbs is None:
            bs_xy, nbs = env.bs_xy.cpu().tolist(), p.num_bs
        else:
            bs_xy, nbs = env.bs_xy[e].cpu().tolist(), int(env.nbs[e])
        if env.stationDict:
            stations = [env.stationDict[k] for k in sorted(env.stationDict)]
        else:
            stations = [BaseStation(b, tuple(bs_xy[b]), **cfg["bs"]) for b in range(nbs)]
        self.stationDict = {bs.bs_id: bs for bs in stations}
        users = []
        for u, proto in sorted(env.userDict.items()):
            ue = UserEquipment(proto.ue_id, proto.velocity, proto.snr_threshold, proto.noise, proto.height)
            ue.x, ue.y = pos[u]
            ue.startTime, ue.exitTime = 0, p.ep_time
            users.append(ue)
        self.userDict = {ue.ue_id: ue for ue in users}
        self.time = float(env.t[e])
        self.activeUsers = list(users) if self.time < p.ep_time else []
        rate = env.rate[e].cpu().tolist()
        util = env.utility_scaled[e].cpu().tolist()
        self.bs2ue_connections = defaultdict(set)
        self.bs2ue_dataRates = {}
        if env.assoc is not None:
            for u, b in enumerate(env.assoc[e].cpu().tolist()):
                if b >= 0:
                    self.bs2ue_connections[stations[b]].add(users[u])
                    self.bs2ue_dataRates[(stations[b], users[u])] = rate[u]
        else:
            for u, mask in enumerate(env.conn[e].cpu().tolist()):
                if isinstance(mask, list):  # B > 32: [E,U,MW] words, word i holds BS 32i..32i+31
                    mask = sum((w & 0xFFFFFFFF) << (32 * i) for i, w in enumerate(mask))
                for b in range(nbs):
                    if (mask >> b) & 1:
                        self.bs2ue_connections[stations[b]].add(users[u])
        self.allUserDataRates = {users[u]: rate[u] for u in range(len(users)) if rate[u] != 0.0}
        self.ue_utilities = {users[u]: util[u] for u in range(len(users))}
        # upstream spellings
        self.stations, self.users, self.active = self.stationDict, self.userDict, self.activeUsers
        self.connections, self.datarates = self.bs2ue_connections, self.bs2ue_dataRates
        self.macro, self.utilities = self.allUserDataRates, self.ue_utilities

    # ---- the reference's per-entity queries on this snapshot (scalar host code, like base.py) ----
    def check_connectivity(self, bs, ue) -> bool:
        """``snr > snr_threshold`` for one link (reference base.py:212-214)."""
        return self.channelModel.calculateSNR(bs, ue) > ue.snr_threshold

    def available_connections(self, ue) -> set:
        """The stations ``ue`` could connect to (reference base.py:216-218)."""
        return {bs for bs in self.stationDict.values() if self.check_connectivity(bs, ue)}

    def update_connections(self) -> None:
        """Keeps only the links that still pass ``check_connectivity`` (reference base.py:221-227)."""
        kept = {bs: {ue for ue in ues if self.check_connectivity(bs, ue)} for bs, ues in self.bs2ue_connections.items()}
        self.bs2ue_connections.clear()
        self.bs2ue_connections.update(kept)

    def allocateDataRate2User(self, bs) -> dict:
        """``{(bs, ue): rate}`` of one station: scheduler share of the Shannon rates, two decimals
        (reference base.py:421-435)."""
        ues = list(self.bs2ue_connections.get(bs, ()))
        snrs = [self.channelModel.calculateSNR(bs, ue) for ue in ues]
        caps = [self.channelModel.datarate(bs, ue, snr) for ue, snr in zip(ues, snrs)]
        shares = self.schedulerModel.share(bs, caps) if ues else []
        return {(bs, ue): round(rate, 2) for ue, rate in zip(ues, shares)}

    def user_total_datarates(self, bs2ue_dataRates) -> Counter:
        """Per-UE sum over its links (reference base.py:413-418)."""
        totals = Counter()
        for (_, ue), rate in bs2ue_dataRates.items():
            totals.update({ue: rate})
        return totals

    def allStationUtilities(self) -> dict:
        """Mean utility of every station's UEs, the scaled lower bound for an idle one (reference
        base.py:438-447)."""
        idle = self.utilityModel.scaleUtility(self.utilityModel.lower)
        out = {}
        for bs in self.stationDict.values():
            ues = self.bs2ue_connections.get(bs)
            out[bs] = sum(self.ue_utilities[ue] for ue in ues) / len(ues) if ues else idle
        return out

    def save_layout_and_data_rates(self, epoch_number: int, curr_step: int, root: str = ".."):
        """Writes this snapshot's four per-step JSON files exactly like the reference does from inside
        ``step`` (base.py:298-349: ``../collectData/{BaseStationPosition,UserEquipmentPosition,DataRate,
        UserQoE}/..._{epoch}_{step}.json``; ``root`` replaces the leading ``..``).  For whole batches
        use ``export.ReferenceDumpWriter`` -- this is the one-env convenience with the reference's name."""
        import os

        from ..export import format_step_files

        stations = [self.stationDict[k] for k in sorted(self.stationDict)]
        users = [self.userDict[k] for k in sorted(self.userDict)]
        assoc = [-1] * len(users)
        for bs, ues in self.bs2ue_connections.items():
            for ue in ues:
                assoc[ue.ue_id] = bs.bs_id
        rate = [self.allUserDataRates.get(ue, 0.0) for ue in users]
        util = self.utilityModel
        files = format_step_files(int(epoch_number), int(curr_step), [(bs.x, bs.y) for bs in stations],
                                  [(ue.x, ue.y) for ue in users], assoc, rate,
                                  (util.lower, util.upper, tuple(util.coeffs)))
        for rel, text in files.items():
            path = os.path.join(root, rel)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(text)
        return sorted(files)
