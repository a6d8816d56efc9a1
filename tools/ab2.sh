# A/B of several env settings: each argument is "NAME=V,NAME2=V2"
i=0
for cfg in "$@"; do i=$((i+1)); env $(echo $cfg | tr ',' ' ') python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab2_$i.json 2>gpurun_out/ab2_$i.err; done
python - "$@" <<EOF2
import json, sys
for i,v in enumerate(sys.argv[1:]):
    try:
        d=json.load(open("gpurun_out/ab2_%d.json"%(i+1)))
        print(v, round(d["ms_per_step"],4), {k.replace("k_so_",""):round(x["ms_per_step"],3) for k,x in d["kernels"].items() if x["ms_per_step"]>0.02 and "so_" in k})
    except Exception as e: print(v, "failed", e)
EOF2
