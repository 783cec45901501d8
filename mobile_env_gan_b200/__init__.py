"""mobile-env on B200: batched GPU replacement for ``MComCore.step``."""
from .core.base import MComCore  # noqa: F401
from .registry import make, register  # noqa: F401
from .scenarios import MComCustom, MComLarge, MComMedium, MComSmall, MComSynthetic  # noqa: F401

__all__ = ["MComCore", "MComCustom", "MComSmall", "MComMedium", "MComLarge", "MComSynthetic", "make", "register"]
