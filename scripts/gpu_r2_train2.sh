set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -s -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
run train tests/test_train.py
timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -h "rel err\|losses\|derivative\|Error\|error" gpurun_out/train.log | head -40
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02b.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","e2e","cfg3_strong","cfg4_train")}, indent=1)[:3500])
PY
tail -5 gpurun_out/bench_r02b.err
