set +e
mkdir -p gpurun_out
for f in 0 1 0 1; do
  SEQDIFF_LN_FUSE=$f python bench.py --steps 4 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 LN_FUSE=$f', round(d['value'],1), 'padded;', round(d['packed']['value'],1), 'packed', d['packed']['identical_to_padded_at_valid_positions'])"
done | tee gpurun_out/ln_fuse_ab_r02.log
for f in 0 1; do
  SEQDIFF_LN_FUSE=$f python bench.py --batch 8 --timesteps 100 --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=8 LN_FUSE=$f', round(d['ms_per_step']*10,2), 'us/step padded;', round(d['packed']['ms_per_step']*10,2), 'packed')"
done | tee -a gpurun_out/ln_fuse_ab_r02.log
