set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train.py -q -s -m gpu -k adamw -p no:cacheprovider 2>&1 | tail -3
for b in 16 128; do for p in 0.1 0.0; do timeout 600 python scripts/train_profile.py --batch $b --p $p; done; done 2>&1 | tee gpurun_out/train_profile_r02.log
