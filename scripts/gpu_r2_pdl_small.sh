set +e
mkdir -p gpurun_out
for b in 1 4 16; do for pdl in 0 1; do
  SEQDIFF_PDL=$pdl timeout 300 python bench.py --batch $b --timesteps 100 --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=$b PDL=$pdl', round(d['value'],1), 'graph-steps/s', round(d['ms_per_step']*10,2), 'us/step')"
done; done | tee gpurun_out/pdl_small_r02.log
