set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -k "gemm_tn" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/train.log 2>&1; echo "train exit $?"; grep -E "per-tensor|passed|failed|Error" gpurun_out/train.log | tail -8
timeout 600 python scripts/train_profile.py --batch 128 > gpurun_out/train_profile_b128_r02c.log 2>&1; head -10 gpurun_out/train_profile_b128_r02c.log
timeout 600 python scripts/train_profile.py --batch 16 > gpurun_out/train_profile_b16_r02c.log 2>&1; head -8 gpurun_out/train_profile_b16_r02c.log
