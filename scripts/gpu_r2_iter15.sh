# round 2, iteration 15: predictor-tail backward with thread-owned dW2 columns in registers (F = 20)
set +e
mkdir -p gpurun_out
L=gpurun_out/iter15.log
: > $L
echo "== test_train" >> $L
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -s 2>&1 | grep -E "cosine|passed|failed|Error|error|assert|L2 rel" | head -40 >> $L
echo "== train_profile batch 128" >> $L
timeout 300 python scripts/train_profile.py --batch 128 2>&1 | grep -v Warning | head -36 >> $L
echo "== train_profile batch 16" >> $L
timeout 300 python scripts/train_profile.py --batch 16 2>&1 | grep -v Warning | head -3 >> $L
echo "== bench cfg4" >> $L
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 2>/dev/null > gpurun_out/bench_cfg4_iter15.json
python -c "import sys,json; d=json.load(open('gpurun_out/bench_cfg4_iter15.json')); c=d['cfg4_train']; print({k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','launches_per_step')})" >> $L
cat $L
