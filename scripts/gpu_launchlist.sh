# warm-cache launch list of graph-replayed steps: per-kernel device time with caches left alone between kernels
set +e
mkdir -p gpurun_out
R=${ROUND:-r01b}
PCMD="python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 700 --launch-count 240 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
