"""Time the general fp64 SDF path (anisotropic / non-binary grids) next to the integer fast path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from or_cdchomp_b200 import models
from or_cdchomp_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
prims, apos, aext = models.clutter_scene()
ce = 0.005 * 400 / N
sizes, lengths, gpose = models.field_geometry(apos, aext, ce, 0.2)
gp = models.prims_to_grid_frame(prims, gpose)
eng = Engine(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
n = int(np.prod(sizes))
d_obs = torch.empty(n, dtype=torch.float64, device="cuda")
d_sdf = torch.empty(n, dtype=torch.float64, device="cuda")
eng.occupancy_device(gp, sizes, lengths, ce, d_obs.data_ptr())
eng.flood_relabel_device(d_obs.data_ptr(), sizes, 0)
for general in (False, True):
    eng.force_general_sdf(general)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); eng.sdf_build_device(d_obs.data_ptr(), sizes, lengths, d_sdf.data_ptr()); e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("general" if general else "fast   ", sizes, "%.2f ms -> %.1f Mvox/s" % (best, n / best / 1e3))
