'''`twx.interp` on the B200: same public names as the reference package (twx/interp/__init__.py:1-4).'''
from .station_select import *      # noqa: F401,F403
from .interp_tair import *         # noqa: F401,F403
from .optimize import *            # noqa: F401,F403
from .tiling import *              # noqa: F401,F403
from .feed import *             # noqa: F401,F403
