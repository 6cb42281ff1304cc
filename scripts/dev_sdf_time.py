"""Time the SDF pipeline stages at BASELINE configs[2] size (400^3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from or_cdchomp_b200 import models
from or_cdchomp_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
prims, apos, aext = models.clutter_scene()
ce = 0.005 * 400 / N
sizes, lengths, gpose = models.field_geometry(apos, aext, ce, 0.2)
print(sizes, lengths)
gp = models.prims_to_grid_frame(prims, gpose)
eng = Engine(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
n = int(np.prod(sizes))
d_occ = torch.empty(n, dtype=torch.float64, device="cuda")
d_obs = torch.empty(n, dtype=torch.float64, device="cuda")
d_sdf = torch.empty(n, dtype=torch.float64, device="cuda")
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t_occ = timed(lambda: eng.occupancy_device(gp, sizes, lengths, ce, d_occ.data_ptr()))
def flood():
    d_obs.copy_(d_occ); eng.flood_relabel_device(d_obs.data_ptr(), sizes, 0)
t_flood = timed(flood)
t_sdf = timed(lambda: eng.sdf_build_device(d_obs.data_ptr(), sizes, lengths, d_sdf.data_ptr()))
print("occupancy ms", t_occ, "flood+relabel ms", t_flood, "sdf ms", t_sdf)
print("SDF build Mvox/s", n / t_sdf / 1e3, " obstacle fraction", float(torch.isinf(d_obs).double().mean()))
