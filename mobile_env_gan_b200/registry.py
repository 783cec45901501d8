"""Minimal ``make`` with Gymnasium ids.  gymnasium itself is not a dependency; when it is
installed the same ids are also registered there (entry points return the batched env)."""
from __future__ import annotations

from .scenarios import MComCustom, MComLarge, MComMedium, MComSmall, MComSynthetic

_REGISTRY = {}


def register(env_id: str, entry, **kwargs):
    _REGISTRY[env_id] = (entry, kwargs)


def make(env_id: str, **kwargs):
    if env_id not in _REGISTRY:
        raise KeyError(f"unknown environment id {env_id!r}; known: {sorted(_REGISTRY)}")
    entry, defaults = _REGISTRY[env_id]
    config = {}
    config.update(kwargs.pop("config", {}) or {})
    # the id fixes the step semantics and the handler (a full MComCore.default_config() passed as
    # `config`, as in the reference README's customisation example, must not undo them)
    config.update(defaults.get("config", {}))
    for key in ("num_envs", "device", "autoreset", "env_offset"):
        if key in kwargs:
            config[key] = kwargs.pop(key)
    return entry(config=config, **kwargs)


for _size, _cls in (("small", MComSmall), ("medium", MComMedium), ("large", MComLarge),
                    ("synthetic", MComSynthetic)):
    for _h in ("central", "ma"):
        register(f"mobile-{_size}-{_h}-v0", _cls, config={"handler": _h, "mode": "gym"})
register("mobile-custom-v0", MComCustom)

try:  # optional: mirror into gymnasium's registry
    import gymnasium as _gym

    for _id, (_entry, _kw) in _REGISTRY.items():
        if _id not in _gym.registry:
            _gym.register(id=_id, entry_point=_entry, kwargs=_kw, disable_env_checker=True)
except Exception:  # gymnasium absent
    pass
