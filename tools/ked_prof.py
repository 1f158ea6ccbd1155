"""Cycle breakdown of the KED kernel from the instrumented build (python -m topowx_b200.build --prof;
run with TWXI_LIB=topowx_b200/libtwxi_prof.so).  Averages per problem, in SM cycles."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from topowx_b200 import db, _lib
from topowx_b200.context import TwxiContext, interp_chunk
da, wrk = bench.build_inputs(0)
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
wrk_d = torch.from_numpy(wrk).cuda()
lib = _lib.lib
buf = (C.c_ulonglong * 16)()
out = interp_chunk(ctx[0], ctx[1], wrk_d)
lib.twxi_ked_prof(buf, 1)
out = interp_chunk(ctx[0], ctx[1], wrk_d, out=out)
lib.twxi_ked_prof(buf, 1)
v = np.array(list(buf), dtype=np.float64)
n = max(v[0], 1)
print("problems", int(v[0]))
for i, nm in ((1, "top of problem -> inputs landed"), (4, "diag: stage loop total"),
              (5, "  barrier wait"), (6, "  chol8_inverse"), (7, "  post-barrier DMMA part")):
    print("   %-28s %9.0f cycles" % (nm, v[i] / n))
n = max(v[8], 1)
for i, nm in ((9, "worker0: stage loop total"), (10, "  barrier wait"), (11, "  stage_rows"), (12, "  wait for L(K+1,K)")):
    print("   %-28s %9.0f cycles" % (nm, v[i] / n))
