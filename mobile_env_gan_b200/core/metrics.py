"""The four built-in scalar metrics (reference mobile_env/core/metrics.py:5-28), read from the
per-env accumulator the step kernel writes (``metrics`` f32 [E,4]); each returns a tensor [E]."""


def number_connections(sim):
    return sim.metrics[:, 0]


def number_connected(sim):
    return sim.metrics[:, 1]


def mean_utility(sim):
    return sim.metrics[:, 2]


def mean_datarate(sim):
    return sim.metrics[:, 3]
