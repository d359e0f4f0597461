set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train.py -q -s -m gpu -p no:cacheprovider > gpurun_out/train.log 2>&1; echo "train exit $?"
grep -h "rel err\|losses\|derivative\|Error\|error\|FAILED\|passed\|failed\|^E " gpurun_out/train.log | head -40
for b in 16 128; do timeout 600 python scripts/train_profile.py --batch $b --p 0.1; done 2>&1 | tee gpurun_out/train_profile_tc_r02.log | grep -E "ms per step|kernel gaps|attention|gemm|transpose " 
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02c.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","packed","e2e")}, indent=1)[:2500])
print("cfg3", {k:d["cfg3_strong"].get(k) for k in ("value","ms_per_sampling","packed")})
print("cfg4", {k:d["cfg4_train"].get(k) for k in ("value","ms_per_step","launches_per_step","error")})
PY
