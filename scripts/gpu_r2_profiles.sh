# round-2 ncu evidence on one B200.  Numbers printed under ncu are never bench values.
set +e
mkdir -p gpurun_out
R=r02
PCMD="python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras"
timeout 600 $PCMD > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 700 --launch-count 280 --csv \
    --log-file gpurun_out/launches_$R.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 120 --launch-count 6 \
    -o gpurun_out/prof_gemm_$R -f $PCMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"; tail -2 gpurun_out/ncu_gemm.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_pipe --launch-skip 30 --launch-count 2 \
    -o gpurun_out/prof_attn_$R -f $PCMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"; tail -2 gpurun_out/ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reverse_step --launch-skip 3 --launch-count 1 \
    -o gpurun_out/prof_rev_$R -f python scripts/revstep_bench.py > gpurun_out/ncu_rev.log 2>&1
echo "ncu rev exit $?"; tail -2 gpurun_out/ncu_rev.log
TCMD="python scripts/train_profile.py --batch 128"
timeout 600 $TCMD > gpurun_out/train_profile_b128_$R.log 2>&1; echo "train profile exit $?"; head -8 gpurun_out/train_profile_b128_$R.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_train_tc --launch-skip 30 --launch-count 3 \
    -o gpurun_out/prof_trainattn_$R -f $TCMD > gpurun_out/ncu_trainattn.log 2>&1
echo "ncu trainattn exit $?"; tail -2 gpurun_out/ncu_trainattn.log
ls -la gpurun_out | tail -12
