from .custom import MComCustom  # noqa: F401
from .gym_scenarios import MComLarge, MComMedium, MComSmall, MComSynthetic  # noqa: F401
