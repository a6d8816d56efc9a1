"""When do the three halo classes really run?  SOGPU_DEBUG_TIMELINE=1 python tools/timeline_probe.py"""
import os, sys
os.environ["SOGPU_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_b200 import api, synth
s = synth.config(1, 1.0)
g = api.SoGpu()
g.set_particles(s.pos, s.mass)
dev = torch.device("cuda")
d_c = torch.from_numpy(s.centers).to(dev); d_r = torch.from_numpy(s.rgtp).to(dev)
d_n = torch.empty(s.h, dtype=torch.int32, device=dev); d_m = torch.empty(s.h, dtype=torch.float32, device=dev)
for rep in range(4):
    g.build_grid()
    g.so_device(d_c.data_ptr(), d_r.data_ptr(), s.h, np.float32(200.0), 8, d_n.data_ptr(), d_m.data_ptr())
    t = g.debug_timeline()
    t0 = min(t[0], t[2], t[4])
    names = ["query<1024>", "query<256>", "query<32>", "query<256> deferred"]
    print("rep", rep, " | ".join("%s start %+7.1f us end %+7.1f us" % (names[k], (t[2*k] - t0) / 1e3, (t[2*k+1] - t0) / 1e3) for k in range(4) if t[2*k+1]))
