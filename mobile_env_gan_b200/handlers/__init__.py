from .central import MComCentralHandler  # noqa: F401
from .multi_agent import MComMAHandler  # noqa: F401

HANDLERS = {"central": MComCentralHandler, "ma": MComMAHandler}
