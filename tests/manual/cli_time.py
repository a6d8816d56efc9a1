"""Time the drop-in `so` program (so_b200/host/so) next to the reference binary on one BASELINE
config, with the phase breakdown (SO_TIMING=1), and diff their output files.  Run under gpurun:
    python tests/manual/cli_time.py [config] [scale] [--no-ref]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from so_b200 import synth, tipsy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "so_b200", "host", "so")
REF = os.path.join(ROOT, "oracle", "_ref", "so_ref")


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    no_ref = "--no-ref" in sys.argv
    extra = [a for a in sys.argv[3:] if a != "--no-ref"]
    s = synth.config(idx, scale)
    tmp = "/dev/shm/so_cli_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    snap, gtp = os.path.join(tmp, "s.tipsy"), os.path.join(tmp, "h.gtp")
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    flags = ["-delta", "200", "-grp", "-gtp", "-O", repr(float(s.omega0))] + extra
    print("[%s] N=%d H=%d  flags %s" % (s.name, s.n, s.h, " ".join(flags)), flush=True)
    env = dict(os.environ, SO_TIMING="1")
    res = {}
    for name, exe in (("ours", SO), ("ref", REF)):
        if name == "ref" and (no_ref or not os.path.exists(REF)):
            continue
        out = os.path.join(tmp, name)
        for rep in range(2 if name == "ours" else 1):         # second run: page cache + driver warm
            t0 = time.time()
            with open(snap, "rb") as fin:
                r = subprocess.run([exe, "-i", gtp, "-o", out] + flags, stdin=fin, capture_output=True, text=True, env=env)
            dt = time.time() - t0
        assert r.returncode == 0, r.stderr[-3000:]
        res[name] = dt
        print("== %s: %.3f s wall" % (name, dt))
        print("\n".join(l for l in r.stderr.splitlines() if "timing" in l or "[sogpu]" in l or "[members]" in l or "CPU Time" in l or l.strip().startswith("SO")))
    if "ref" in res:
        a, b = tipsy.read_sogrp(os.path.join(tmp, "ours.sogrp")), tipsy.read_sogrp(os.path.join(tmp, "ref.sogrp"))
        print(".sogrp identical:", bool(np.array_equal(a, b)))
        ha, ra = tipsy.parse_sovcirc(os.path.join(tmp, "ours.sovcirc"))
        hb, rb = tipsy.parse_sovcirc(os.path.join(tmp, "ref.sovcirc"))
        print(".sovcirc rows identical:", bool(np.array_equal(np.array(ra), np.array(rb))))
        print("speed-up (wall, whole program): %.1fx" % (res["ref"] / res["ours"]))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
