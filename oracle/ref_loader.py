"""Execute the reference's own pure-numpy hot-path code verbatim (container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` exists in the build
container but not on the GPU box, so this module is used solely by ``tests/golden/make_golden.py``
to generate committed golden vectors and by CPU tests that are skipped when the tree is absent.

What is loaded, unmodified, with two Python-2 shims (``unicode = str``; ``np.alltrue``/``np.bool``):
* ``twx/utils/util_geo.py``            -> ``grt_circle_dist``
* ``twx/interp/station_select.py``     -> ``StationSelect``
* ``twx/interp/interp_tair.py:1099-1146`` -> ``_gwr_series`` (the file as a whole is Python-2 only)
* ``twx/interp/interp_tair.py:143-197``   -> ``tmin_tmax_fixer``
"""
import builtins
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("TWX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "twx", "interp", "station_select.py"))


def _shims():
    if not hasattr(builtins, "unicode"):
        builtins.unicode = str
    if not hasattr(np, "alltrue"):
        np.alltrue = np.all
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float


def _read(rel, first=None, last=None):
    with open(os.path.join(REF_ROOT, rel)) as f:
        lines = f.readlines()
    if first is not None:
        lines = lines[first - 1:last]
    return "".join(lines)


def load():
    """Returns a namespace with the reference's grt_circle_dist, StationSelect, _gwr_series,
    tmin_tmax_fixer, executed from the reference sources where they lie."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _shims()
    ns = types.SimpleNamespace()

    g_geo = {"__name__": "ref_util_geo"}
    exec(compile(_read("twx/utils/util_geo.py"), "util_geo.py", "exec"), g_geo)
    ns.grt_circle_dist = g_geo["grt_circle_dist"]

    # stub packages so that `from twx.db import LON, LAT, STN_ID` and `from twx.utils import
    # grt_circle_dist` in station_select.py resolve (field names: twx/db/station_data.py:49-68)
    saved = {k: sys.modules.get(k) for k in ("twx", "twx.db", "twx.utils")}
    twx = types.ModuleType("twx")
    twx_db = types.ModuleType("twx.db")
    twx_db.LON, twx_db.LAT, twx_db.STN_ID = "longitude", "latitude", "station_id"
    twx_utils = types.ModuleType("twx.utils")
    twx_utils.grt_circle_dist = ns.grt_circle_dist
    twx.db, twx.utils = twx_db, twx_utils
    sys.modules.update({"twx": twx, "twx.db": twx_db, "twx.utils": twx_utils})
    try:
        g_ss = {"__name__": "ref_station_select"}
        exec(compile(_read("twx/interp/station_select.py"), "station_select.py", "exec"), g_ss)
        ns.StationSelect = g_ss["StationSelect"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    g_it = {"np": np, "__name__": "ref_interp_tair"}
    exec(compile(_read("twx/interp/interp_tair.py", 1099, 1146), "interp_tair.py[1099:1146]", "exec"), g_it)
    ns._gwr_series = g_it["_gwr_series"]
    exec(compile(_read("twx/interp/interp_tair.py", 143, 197), "interp_tair.py[143:197]", "exec"), g_it)
    ns.tmin_tmax_fixer = g_it["tmin_tmax_fixer"]
    return ns
