set +e
mkdir -p gpurun_out
for pdl in 1 0; do
  echo "== SEQDIFF_PDL=$pdl"
  SEQDIFF_PDL=$pdl timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_pdl$pdl.json'))
    print("value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_pdl$pdl.err').read()[-1500:])
PY
done
SEQDIFF_PDL=1 timeout 1500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ops.py -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -3
