"""Alias of :mod:`.channels` under the module name the reference README uses in its customisation
example (``from mobile_env.core.channel import Channel``, README.md:108-121; the fork's file is
``core/channels.py``)."""
from .channels import EPSILON, Channel, LogDistance, OkumuraHata  # noqa: F401
