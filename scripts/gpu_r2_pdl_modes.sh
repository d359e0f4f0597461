set +e
mkdir -p gpurun_out
for pdl in 0 1 2 3 0 2; do
  SEQDIFF_PDL=$pdl timeout 300 python bench.py --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 B=64 PDL=$pdl', round(d['value'],1), 'graph-steps/s', round(d['ms_per_step']*2,2), 'us/step; packed', round(d['packed']['value'],1))"
done | tee gpurun_out/pdl_modes_r02.log
