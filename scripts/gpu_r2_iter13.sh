# round 2, iteration 13: dropout + residual folded into the forward LayerNorm pass, hidden-dropout masks handed to the fused LayerNorm
# backward as bits; new 16-bit parity tests (small tensors; dropout on vs the fp32 mode under the same masks)
set +e
mkdir -p gpurun_out
L=gpurun_out/iter13.log
: > $L
echo "== test_train (defaults)" >> $L
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -s 2>&1 | grep -E "cosine|passed|failed|Error|error|assert|L2 rel" | head -40 >> $L
echo "== test_train 16-bit parity with the unfused forms" >> $L
SEQDIFF_DROP_LN_FUSE=0 SEQDIFF_LN_BWD_FUSE=0 timeout 600 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -s -k "16bit" 2>&1 | grep -E "cosine|passed|failed|L2 rel" >> $L
for env in "" "SEQDIFF_DROP_LN_FUSE=0"; do
  echo "== train_profile batch 128 [$env]" >> $L
  env $env timeout 300 python scripts/train_profile.py --batch 128 2>&1 | grep -v Warning | head -24 >> $L
done
echo "== train_profile batch 16" >> $L
timeout 300 python scripts/train_profile.py --batch 16 2>&1 | grep -v Warning | head -3 >> $L
for env in "" "SEQDIFF_DROP_LN_FUSE=0"; do
  echo "== bench cfg4 [$env]" >> $L
  env $env timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['cfg4_train']; print({k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','launches_per_step')})" >> $L
done
cat $L
