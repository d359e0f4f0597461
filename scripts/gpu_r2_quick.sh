set +e
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 2 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'packed', d['packed']['value'])"
