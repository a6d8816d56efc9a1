set -x
timeout 600 python tools/virtual_ranks_probe.py --ranks 8 --scale 0.5 --steps 1 > gpurun_out/vr8_half.json 2> gpurun_out/vr8_half.err || exit 1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_route_stage --launch-skip 26 --launch-count 2 -o gpurun_out/r2_route_multi -f python tools/virtual_ranks_probe.py --ranks 8 --scale 0.5 --steps 1 > gpurun_out/ncu_vr.log 2>&1
tail -3 gpurun_out/ncu_vr.log
