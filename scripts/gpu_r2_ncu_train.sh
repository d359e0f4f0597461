# ncu evidence of the round-2 training kernels (one B200)
set +e
mkdir -p gpurun_out
TCMD="python scripts/train_profile.py --batch 128"
timeout 600 $TCMD > gpurun_out/plain_train.log 2>&1; echo "plain exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_pipe --launch-skip 28 --launch-count 4 \
    -o gpurun_out/prof_attnbwdpipe_r02 -f $TCMD > gpurun_out/ncu_attnbwdpipe.log 2>&1
echo "ncu attnbwdpipe exit $?"; tail -2 gpurun_out/ncu_attnbwdpipe.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 700 --launch-count 12 \
    -o gpurun_out/prof_traingemm_r02 -f $TCMD > gpurun_out/ncu_traingemm.log 2>&1
echo "ncu traingemm exit $?"; tail -2 gpurun_out/ncu_traingemm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 1100 --launch-count 380 --csv \
    --log-file gpurun_out/launches_train_r02.csv $TCMD > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu launches exit $?"
