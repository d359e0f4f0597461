set +e
mkdir -p gpurun_out
TCMD="python scripts/train_profile.py --batch 128"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_train_tc --launch-skip 28 --launch-count 8 \
    -o gpurun_out/prof_trainbwd_r02 -f $TCMD > gpurun_out/ncu_trainbwd.log 2>&1
echo "ncu trainbwd exit $?"; tail -2 gpurun_out/ncu_trainbwd.log
