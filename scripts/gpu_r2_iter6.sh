set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/train.log 2>&1; echo "train exit $?"; tail -5 gpurun_out/train.log
timeout 600 python scripts/train_profile.py --batch 128 > gpurun_out/train_profile_b128_r02b.log 2>&1; head -9 gpurun_out/train_profile_b128_r02b.log
timeout 600 python scripts/train_profile.py --batch 16 > gpurun_out/train_profile_b16_r02b.log 2>&1; head -9 gpurun_out/train_profile_b16_r02b.log
timeout 600 python bench.py --workload cfg4 --steps 8 --warmup 3 > gpurun_out/bench_cfg4_1gpu.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_1gpu.json')); print(d['value'], d['ms_per_step'], d['cfg4_train']['train_flops_frac_of_peak'])"
