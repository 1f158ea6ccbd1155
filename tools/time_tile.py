"""Stage timing of the benchmark tile through twxi_interp_chunk (device-resident buffers).
usage: python tools/time_tile.py [reps]   -> prints stage_ms (knn, nngh_params, krig, gwr_daily, fixer_quantise)"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from topowx_b200 import db, _lib
from topowx_b200.context import TwxiContext, interp_chunk
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
da, wrk = bench.build_inputs(0)
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
wrk_d = torch.from_numpy(wrk).cuda()
lib = _lib.lib
out = None
tot = np.zeros(5); ked = 0.0
for i in range(reps + 1):
    lib.twxi_set_stage_timing(1)
    out = interp_chunk(ctx[0], ctx[1], wrk_d, out=out)
    torch.cuda.synchronize()
    s5 = (C.c_float * 5)(); lib.twxi_get_stage_ms(s5)
    kk = C.c_float(); lib.twxi_get_ked_kernel_ms(C.byref(kk))
    if i: tot += np.array(list(s5)); ked += kk.value
print(os.environ.get("TWXI_KED_CFG", "default"), "stage_ms", np.round(tot / reps, 3).tolist(), "sum", round(float(tot.sum() / reps), 3),
      "ked_kernel only", round(ked / reps, 3), "ok cells", int((out["status"].cpu().numpy() == 0).sum()))
