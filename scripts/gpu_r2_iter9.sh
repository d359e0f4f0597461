set +e
for pdl in unset 1 0 unset 1 0; do
  if [ $pdl = unset ]; then unset SEQDIFF_PDL; else export SEQDIFF_PDL=$pdl; fi
  python bench.py --batch 1 --timesteps 100 --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=1 PDL=$pdl', round(d['ms_per_step']*10,2), 'us/step')"
done
unset SEQDIFF_PDL
python scripts/struct_bench.py --steps 2 --warmup 2 --timesteps 300 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('struct default', round(d['value'],1))"
timeout 900 python -m pytest tests/test_structure_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -1
