"""Experiment: extrema kernels on replicated (divergence-free) vs real segments."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import mrs_uav_trajectory_generation_b200 as tg
from mrs_uav_trajectory_generation_b200 import workloads as W
import os
libpath = (sys.argv[1] or None) if len(sys.argv) > 1 else None
ctx = tg.Context(tg.Library(libpath), 0)
print('library', libpath or 'default')
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wp_off, wp = W.random_flier_paths_fast(B, first_index=0)
P = ctx.L.default_params(check_deviation=0)
res, tot = ctx.optimize_batch(wp_off, wp, None, None, P)
out = ctx.fetch_outputs()
coef, times = out["coef"], out["times"]
S = len(times)
def run(c, t, label):
    ctx.set_profiling(True)
    ctx.extrema(c, t)
    pr = ctx.profile()
    ctx.set_profiling(False)
    tot = sum(v[0] for v in pr.values())
    print(label, "segments", len(t), "total ms %.3f" % tot, {k.split("<")[1][:6]: round(v[0], 3) for k, v in sorted(pr.items())})
run(coef, times, "real      ")
run(coef, times, "real again")
rep = np.repeat(coef[5:6], S, axis=0); rt = np.repeat(times[5:6], S)
run(rep, rt, "replicated")
# sorted by warp: groups of 32 identical segments
idx = np.repeat(np.arange(S // 32), 32)
run(coef[idx], times[idx], "warp-uniform")
