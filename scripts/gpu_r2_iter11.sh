# round 2, iteration 11: fused LayerNorm backward (+ 16-bit dY operand + bias gradient), float4 AdamW, deeper sumsq loads, opt-in PDL in the
# training step -- gradient / optimizer parity first, then A/B timings of every switch, then the full bench line
set +e
mkdir -p gpurun_out
L=gpurun_out/iter11.log
: > $L
echo "== test_train (default: fused LN backward)" >> $L
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -6 >> $L
echo "== test_train -k 'grad or parity or ddp or overlap' under SEQDIFF_TRAIN_PDL=3" >> $L
SEQDIFF_TRAIN_PDL=3 timeout 600 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -k "grad or parity or matches or dropout" 2>&1 | tail -4 >> $L
for env in "" "SEQDIFF_LN_BWD_FUSE=0" "SEQDIFF_LNBWD_OCC=2" "SEQDIFF_TRAIN_PDL=3" "SEQDIFF_TRAIN_PDL=1"; do
  echo "== train_profile batch 128 [$env]" >> $L
  env $env timeout 300 python scripts/train_profile.py --batch 128 2>&1 | grep -v Warning | head -14 >> $L
done
for env in "" "SEQDIFF_TRAIN_PDL=3" "SEQDIFF_TRAIN_PDL=1"; do
  echo "== train_profile batch 16 [$env]" >> $L
  env $env timeout 300 python scripts/train_profile.py --batch 16 2>&1 | grep -v Warning | head -3 >> $L
done
echo "== bench (full line)" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_iter11.json 2> gpurun_out/bench_iter11.err; echo "bench exit $?" >> $L
python - >> $L <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter11.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'packed', d['packed']['value'], 'clocks', d['clocks'])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], 'share', r['share_of_forward']); print(r['kernel_ms_per_forward'])
c=d['cfg4_train']; print('cfg4', {k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','launches_per_step','error')})
c=d['cfg3_strong']; print('cfg3', {k:c.get(k) for k in ('value','ms_per_sampling','error')}, (c.get('packed') or {}).get('value'))
PY
cat $L
