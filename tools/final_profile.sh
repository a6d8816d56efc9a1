# the two ncu passes behind profiles/r2_*: launch list of the bench command, then --set full of one device-resident step
# usage (GPU box): bash tools/final_profile.sh   -> gpurun_out/r2f_*
set -x
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cfg1 --no-cpu-baseline"
timeout 500 $CMD > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu1.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'k_route_stage|k_lvl_partition_rt|k_bucket_sort_sparse|k_so_query_fused|k_lvl_hist' --launch-skip 27 --launch-count 9 -o gpurun_out/r2f_full -f $CMD > gpurun_out/r2f_ncu2.log 2>&1
tail -2 gpurun_out/r2f_ncu2.log
