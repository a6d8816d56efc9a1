# ncu --set full on the kernels one rank of an 8-rank domain step runs (virtual ranks on one GPU); usage: bash tools/ncu_vr.sh
# EXPENSIVE: the process holds ~40 GB of buffers that ncu saves and restores around every replay — 31 launches at
# full scale took 18 GPU-minutes (profiles/r2_experiments.md).  Prefer --scale 0.5 and one or two launches.
set -x
timeout 600 python tools/virtual_ranks_probe.py --ranks 8 --steps 1 > gpurun_out/vr8_s1.json 2> gpurun_out/vr8_s1.err || exit 1
timeout 1000 ncu --set full --clock-control none --import-source on -k regex:'k_bucket_sort_sparse|k_lvl_partition|k_so_query|k_mark_table|k_route_split|k_lvl_hist' --launch-skip 264 --launch-count 32 -o gpurun_out/r2_rank8 -f python tools/virtual_ranks_probe.py --ranks 8 --steps 1 > gpurun_out/ncu_vr.log 2>&1
tail -3 gpurun_out/ncu_vr.log
