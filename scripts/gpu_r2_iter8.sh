set +e
timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_structure_model.py -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 4 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', round(d['value'],1), 'padded;', round(d['packed']['value'],1), 'packed', d['packed']['identical_to_padded_at_valid_positions'])"; done
python bench.py --workload cfg3 --steps 2 --warmup 2 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3', round(d['value'],1), 'padded;', round(d['packed']['value'],1), 'packed')"
python bench.py --batch 1 --timesteps 100 --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=1', round(d['ms_per_step']*10,2), 'us/step')"
for pdl in 1 3 1 3; do SEQDIFF_PDL=$pdl python scripts/struct_bench.py --steps 2 --warmup 2 --timesteps 300 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('struct PDL=$pdl', round(d['value'],1))"; done
