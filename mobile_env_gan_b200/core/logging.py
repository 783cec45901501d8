"""Monitor with the reference's interface (mobile_env/core/logging.py:6-87), batched: every
metric callable receives the env and returns a tensor with a leading env axis; results are
kept on the device and only moved by ``load_results``.  Off the step path: ``MComCore.step``
never calls ``update``; callers that want a trace call ``env.monitor.update(env)``."""
from __future__ import annotations

from typing import Dict

import torch


class Monitor:
    def __init__(self, scalar_metrics: Dict, ue_metrics: Dict, bs_metrics: Dict, **kwargs):
        self.scalar_metrics = scalar_metrics
        self.ue_metrics = ue_metrics
        self.bs_metrics = bs_metrics
        self.scalar_results = None
        self.ue_results = None
        self.bs_results = None

    def reset(self):
        self.scalar_results = {name: [] for name in self.scalar_metrics}
        self.ue_results = {name: [] for name in self.ue_metrics}
        self.bs_results = {name: [] for name in self.bs_metrics}

    def update(self, simulation):
        for results, metrics in (
            (self.scalar_results, self.scalar_metrics),
            (self.ue_results, self.ue_metrics),
            (self.bs_results, self.bs_metrics),
        ):
            for name, metric in metrics.items():
                results[name].append(torch.as_tensor(metric(simulation)).detach().clone())  # O(1) per step

    def load_results(self, env: int = 0):
        """Pandas frames for one env, shaped like the reference's (logging.py:44-75)."""
        import pandas as pd

        scalar = pd.DataFrame({k: [float(v[env]) for v in vs] for k, vs in self.scalar_results.items()})
        scalar.index.names = ["Time Step"]

        def frame(results, id_name):
            rows = {}
            for metric, entries in results.items():
                for step, vals in enumerate(entries):
                    for i, v in enumerate(vals[env].tolist()):
                        rows.setdefault((step, i), {})[metric] = v
            df = pd.DataFrame.from_dict(rows, orient="index")
            if len(df):
                df.index = pd.MultiIndex.from_tuples(df.index, names=["Time Step", id_name])
                df = df.sort_index()[sorted(df.columns)]
            df.columns.name = "Metric"
            return df

        return scalar, frame(self.ue_results, "UE ID"), frame(self.bs_results, "BS ID")

    def info(self, env: int = 0):
        if not self.scalar_results or any(len(r) == 0 for r in self.scalar_results.values()):
            return {}
        out = {k: float(v[-1][env]) for k, v in self.scalar_results.items()}
        out.update({k: v[-1][env].tolist() for k, v in self.ue_results.items()})
        out.update({k: v[-1][env].tolist() for k, v in self.bs_results.items()})
        return out
