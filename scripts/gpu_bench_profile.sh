# bench + profiles on one B200.  Numbers printed under ncu are never bench values.
set +e
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/fwd.log 2>&1; echo "fwd exit $?"
tail -3 gpurun_out/fwd.log
timeout 900 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
cat gpurun_out/bench_$R.json; tail -3 gpurun_out/bench_$R.err
PCMD="python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 650 --launch-count 230 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
timeout 600 $PCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 120 --launch-count 4 \
    -o gpurun_out/prof_gemm_$R -f $PCMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"; tail -2 gpurun_out/ncu_gemm.log
timeout 600 $PCMD > gpurun_out/plain3.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_16 --launch-skip 16 --launch-count 2 \
    -o gpurun_out/prof_attn_$R -f $PCMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"; tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out
