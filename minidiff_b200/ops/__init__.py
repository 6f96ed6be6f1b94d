from .definitions import *  # noqa: F401,F403
from .wrapping import *  # noqa: F401,F403
