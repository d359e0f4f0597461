# round-end evidence run: full GPU tests, smoke, headline bench, cfg3 bench, ncu launch list (caches left warm between kernels) + full captures
set +e
mkdir -p gpurun_out
R=${ROUND:-r01}
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -1 gpurun_out/$name.log; }
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py -s
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"; cat gpurun_out/bench_$R.json
timeout 1200 python bench.py --workload cfg3 --steps 2 --warmup 3 > gpurun_out/bench_cfg3_$R.json 2> gpurun_out/bench_cfg3_$R.err; echo "bench cfg3 exit $?"
timeout 600 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep_$R.log 2>&1
timeout 300 python scripts/revstep_bench.py > gpurun_out/revstep_$R.log 2>&1; cat gpurun_out/revstep_$R.log
export SEQDIFF_PROFILER_RANGE=1  # bench.py brackets its timed region with cudaProfilerStart/Stop: ncu sees graph replays only, not the GEMM tuner
PCMD="python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --launch-skip 10 --launch-count 240 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 $PCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 --launch-skip 10 --launch-count 10 \
    -o gpurun_out/prof_gemm_$R -f $PCMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 600 $PCMD > gpurun_out/plain3.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_ --launch-skip 2 --launch-count 3 \
    -o gpurun_out/prof_attn_$R -f $PCMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"
ITERS=2 timeout 300 python scripts/revstep_bench.py > /dev/null 2>&1 && \
ITERS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:reverse_step -c 1 -o gpurun_out/prof_rev_$R -f python scripts/revstep_bench.py > gpurun_out/ncu_rev.log 2>&1
echo "ncu rev exit $?"
ls gpurun_out | head -50
