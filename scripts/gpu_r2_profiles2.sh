# round-2 ncu evidence of the TIMED region only (cudaProfilerStart / Stop around it: the GEMM tuner's candidate launches of the warm-up stay out)
set +e
mkdir -p gpurun_out
R=r02
export SEQDIFF_PROFILER_RANGE=1
PCMD="python bench.py --steps 1 --warmup 2 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-count 400 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 55 --launch-count 12 \
    -o gpurun_out/prof_gemm_$R -f $PCMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"; tail -2 gpurun_out/ncu_gemm.log
