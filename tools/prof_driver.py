import sys; sys.path.insert(0,'.')
import numpy as np
import mrs_uav_trajectory_generation_b200 as tg
from mrs_uav_trajectory_generation_b200 import workloads as W
ctx = tg.Context(tg.Library(), 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wp_off, wp = W.random_flier_paths_fast(B, first_index=0)
P = ctx.L.default_params(max_deviation_iters=int(sys.argv[2]) if len(sys.argv) > 2 else 6)
for i in range(2):
    res, tot = ctx.optimize_batch(wp_off, wp, None, None, P)
print("ok", ctx.last_device_ms(), res["rounds"].mean())
