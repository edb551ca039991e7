# round 2, final build (path cache on): same-box A/B of the bench with the cache off / on, then ncu IN THE BENCH WINDOW as scripts/gpu_profile_r2.sh:
# launch list of move 5 and one --set full capture of two consecutive simulation steps.  bench.py numbers printed under ncu are not bench values.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
for pc in off on; do
  $K 200 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-secondary --path-cache $pc > gpurun_out/bench_r2_pc_$pc.json 2> gpurun_out/bench_r2_pc_$pc.err
  echo "path cache $pc:"; python scripts/show_bench.py gpurun_out/bench_r2_pc_$pc.json | head -4 | cut -c1-400
done
$K 300 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv|k_tc_fc|k_select_expand|k_backup" --launch-skip 10100 --launch-count 8 \
    -f -o gpurun_out/prof_r2b python bench.py --steps 1 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full_r2b.log 2>&1
$K 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15200 --launch-count 1200 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 1 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench_r2b.log 2>&1
tail -2 gpurun_out/launches_r2b.csv | cut -c1-200; ls -la gpurun_out | tail -6
