# round-end evidence refresh (tests are run separately: scripts/gpu_all_tests.sh): headline bench, cfg3 bench, structure bench,
# reverse-step microbench, ncu launch lists (sequence step graph replays; structure step graph replays)
set +e
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_$R.json
timeout 1200 python bench.py --workload cfg3 --steps 2 --warmup 3 > gpurun_out/bench_cfg3_$R.json 2> gpurun_out/bench_cfg3_$R.err; echo "bench cfg3 exit $?"; cut -c1-200 gpurun_out/bench_cfg3_$R.json
timeout 900 python scripts/struct_bench.py > gpurun_out/struct_bench_$R.json 2> gpurun_out/struct_bench_$R.err; echo "struct bench exit $?"; cut -c1-200 gpurun_out/struct_bench_$R.json
timeout 300 python scripts/revstep_bench.py > gpurun_out/revstep_$R.log 2>&1; cat gpurun_out/revstep_$R.log
export SEQDIFF_PROFILER_RANGE=1
PCMD="python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --launch-skip 10 --launch-count 240 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
unset SEQDIFF_PROFILER_RANGE
SCMD="python scripts/struct_bench.py --no-cpu --steps 1 --warmup 1 --timesteps 4"
timeout 600 $SCMD > gpurun_out/plain_struct.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 2600 --launch-count 300 --csv \
    --log-file gpurun_out/launches_struct_$R.csv $SCMD > gpurun_out/ncu_launches_struct.log 2>&1
echo "ncu struct launches exit $?"
ls -la gpurun_out/launches_$R.csv gpurun_out/launches_struct_$R.csv
