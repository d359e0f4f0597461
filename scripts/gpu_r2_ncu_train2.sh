# ncu evidence of the late round-2 training step (one B200): launch list of ONE step (profiler range) and a --set full capture of the fused
# LayerNorm backward (operand + bias gradient).  Reports stay small (no source import, two captures): gpurun_out/ is capped at 64 MiB.
set +e
mkdir -p gpurun_out
export SEQDIFF_PROFILER_RANGE=1
TCMD="python scripts/train_profile.py --batch 128"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train_r02.csv $TCMD > gpurun_out/ncu_launches_train2.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"layernorm_bwd" \
    --launch-count 2 -o gpurun_out/prof_trainrow_r02 -f $TCMD > gpurun_out/ncu_trainrow.log 2>&1
echo "ncu trainrow exit $?"; tail -2 gpurun_out/ncu_trainrow.log
ls -la gpurun_out
