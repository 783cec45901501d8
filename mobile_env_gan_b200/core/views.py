"""Reference-named host view of one environment (slow path, for inspection and export)."""
from __future__ import annotations

from collections import defaultdict

from .entities import BaseStation, UserEquipment


class EnvView:
    """Snapshot with the attribute names of reference base.py:69-79."""

    def __init__(self, env, e: int):
        p = env.plan
        cfg = env.config
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
                for b in range(nbs):
                    if (mask >> b) & 1:
                        self.bs2ue_connections[stations[b]].add(users[u])
        self.allUserDataRates = {users[u]: rate[u] for u in range(len(users)) if rate[u] != 0.0}
        self.ue_utilities = {users[u]: util[u] for u in range(len(users))}
        # upstream spellings
        self.stations, self.users, self.active = self.stationDict, self.userDict, self.activeUsers
        self.connections, self.datarates = self.bs2ue_connections, self.bs2ue_dataRates
        self.macro, self.utilities = self.allUserDataRates, self.ue_utilities
