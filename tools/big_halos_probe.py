import sys, numpy as np, time
sys.path.insert(0, "/root/repo")
from so_b200 import api, synth
s = synth.config(1, 1.0)
g = api.SoGpu()
g.profile_enable(True)
g.set_particles(s.pos, s.mass)
g.build_grid()
order = np.argsort(-s.rgtp)
for name, sel in (("top1", order[:1]), ("top10", order[:10]), ("top40", order[:40]), ("rank 100-140", order[100:140]), ("rank 1000-1040", order[1000:1040])):
    for rep in range(3):
        g.profile_read(reset=True)
        r = g.so(s.centers[sel], s.rgtp[sel], 200.0)
        st = g.stats()
        prof = {k: round(v[0], 4) for k, v in g.profile_read(reset=True).items() if v[1] and "query" in k or "emit" in k and v[1]}
    print(name, "ndelta", r["ndelta"][:3], "evals", st["last_evals"], "hist evals", st["last_evals_first"], "members", st["last_members"], prof)
