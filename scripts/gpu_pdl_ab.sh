set +e
mkdir -p gpurun_out
for pdl in 0 1 0 1; do
  SEQDIFF_PDL=$pdl timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl$pdl.json'))
print("PDL=$pdl value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), d["clocks"])
PY
done
SEQDIFF_PDL=1 timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -2
