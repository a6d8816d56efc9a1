python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for fb in 2 1; do SOGPU_FIRST_BALL=$fb python bench.py --no-cpu-baseline > gpurun_out/b_fb$fb.json 2>gpurun_out/b_fb$fb.err; done
python - <<EOF2
import json
for f in ("b_fb2","b_fb1"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["ms_per_step"], d["evals_per_step"], {k:(round(v["ms_per_step"],4), round(v.get("gbs",0))) for k,v in d["kernels"].items()})
EOF2
