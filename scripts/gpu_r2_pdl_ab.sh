set +e
mkdir -p gpurun_out
for rep in 1 2 3; do for pdl in 0 3; do
  SEQDIFF_PDL=$pdl timeout 300 python bench.py --steps 4 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 B=64 PDL=$pdl', round(d['value'],1), 'padded;', round(d['packed']['value'],1), 'packed')"
done; done | tee gpurun_out/pdl_mode3_ab_r02.log
for pdl in 0 3; do
  SEQDIFF_PDL=$pdl timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 2 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 PDL=$pdl', round(d['value'],1), 'padded;', round(d['packed']['value'],1), 'packed')"
done | tee -a gpurun_out/pdl_mode3_ab_r02.log
