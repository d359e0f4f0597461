# ncu evidence of the late round-2 training step (one B200): launch list of ONE step (profiler range) and --set full captures of the fused
# rowwise kernels (LayerNorm backward with operand + bias gradient, dropout + residual + LayerNorm forward, dense embedding backward)
set +e
mkdir -p gpurun_out
export SEQDIFF_PROFILER_RANGE=1
TCMD="python scripts/train_profile.py --batch 128"
timeout 600 $TCMD > gpurun_out/plain_train2.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train_r02.csv $TCMD > gpurun_out/ncu_launches_train2.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"layernorm_bwd|dropout_add_layernorm|embed_bwd" \
    --launch-count 40 -o gpurun_out/prof_trainrow_r02 -f $TCMD > gpurun_out/ncu_trainrow.log 2>&1
echo "ncu trainrow exit $?"; tail -2 gpurun_out/ncu_trainrow.log
