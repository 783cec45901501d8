"""Config helpers (mirrors what reference mobile_env/core/util.py:31-40 offers to callers;
the rendering glyph of util.py:6-28 is out of scope)."""
from __future__ import annotations

from typing import Dict


def deep_dict_merge(dest: Dict, source: Dict) -> Dict:
    """Recursively overlays ``source`` onto ``dest`` in place and returns ``dest``
    (same contract as reference core/util.py:31-40)."""
    for key, val in source.items():
        if isinstance(val, dict):
            deep_dict_merge(dest.setdefault(key, {}), val)
        else:
            dest[key] = val
    return dest
