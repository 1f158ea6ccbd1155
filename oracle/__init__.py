"""CPU oracle for the TopoWx interpolation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and there only as the checker (or as the CPU baseline being timed), never on the product path.
The product (``topowx_b200``) never imports this package and fails loudly without its CUDA library.

Parity status
-------------
* Neighbour search, bisquare weights, neighbour-count averaging, variogram smoothing and the GWR
  series (``twx/utils/util_geo.py``, ``twx/interp/station_select.py``, ``twx/interp/interp_tair.py``)
  are restated in numpy in :mod:`oracle.twx_oracle` and PINNED against the reference's own code,
  executed verbatim from ``/root/reference`` by :mod:`oracle.ref_loader` (see
  ``tests/golden/make_golden.py`` and ``tests/test_oracle_vs_golden.py``).
* The kriging arithmetic of ``krig_meantair`` (``twx/interp/rpy/interp.R:198-270``) runs inside the
  third-party R packages ``gstat`` 1.0-25 and ``sp`` 1.1-1 (``INSTALL.rst:56,59``), which are neither
  vendored in the reference nor installable here (no R, no network).  :func:`oracle.twx_oracle.ked_gstat`
  restates their published algorithm (kriging with external drift on WGS-84 great-circle distances);
  it is pinned only by analytic known-answer tests, so against real gstat it is **parity unpinned**.
"""
