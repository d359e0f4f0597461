# round 2, iteration 12: cp.async-pipelined fused LayerNorm backward, dense embedding backward (register accumulators), PDL on by default in
# the training step; A/B of each switch and of the early-trigger library variant (-DSEQDIFF_TRAIN_TRIGGER)
set +e
mkdir -p gpurun_out
L=gpurun_out/iter12.log
: > $L
TRIG=$PWD/e3-invaraint-diffusion-model_b200/libseqdiff_b200_trig.so
echo "== test_train (defaults)" >> $L
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -6 >> $L
echo "== test_train parity subset, early-trigger variant" >> $L
SEQDIFF_LIB=$TRIG timeout 600 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -k "grad or parity or matches or dropout or resume" 2>&1 | tail -4 >> $L
for env in "" "SEQDIFF_LNBWD_PIPE=0" "SEQDIFF_EMBED_BWD_DENSE=0" "SEQDIFF_TRAIN_PDL=0" "SEQDIFF_LIB=$TRIG" "SEQDIFF_LIB=$TRIG SEQDIFF_TRAIN_PDL=1"; do
  echo "== train_profile batch 128 [$env]" >> $L
  env $env timeout 300 python scripts/train_profile.py --batch 128 2>&1 | grep -v Warning | head -22 >> $L
done
for env in "" "SEQDIFF_LIB=$TRIG" "SEQDIFF_LIB=$TRIG SEQDIFF_TRAIN_PDL=1"; do
  echo "== train_profile batch 16 [$env]" >> $L
  env $env timeout 300 python scripts/train_profile.py --batch 16 2>&1 | grep -v Warning | head -3 >> $L
done
for env in "" "SEQDIFF_LIB=$TRIG"; do
  echo "== bench cfg4 [$env]" >> $L
  env $env timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['cfg4_train']; print({k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','launches_per_step')})" >> $L
done
cat $L
