# A/B of tuning knobs on the GPU box: usage  bash tools/ab.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do env $var=$v python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err; done
python - "$@" <<EOF2
import json, sys
for v in sys.argv[1:]:
    try:
        d=json.load(open("gpurun_out/ab_%s.json"%v))
        print(v, round(d["ms_per_step"],4), {k:(round(x["ms_per_step"],4), round(x.get("gbs",0))) for k,x in d["kernels"].items() if x["ms_per_step"]>0.02})
    except Exception as e: print(v, "failed", e)
EOF2
