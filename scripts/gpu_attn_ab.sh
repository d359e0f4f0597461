set +e
mkdir -p gpurun_out
for bq in 128 64; do
  echo "== SEQDIFF_ATTN_BQ=$bq"
  SEQDIFF_ATTN_BQ=$bq timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k attention -p no:cacheprovider 2>&1 | tail -1
  SEQDIFF_ATTN_BQ=$bq timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bq$bq.json 2> gpurun_out/bench_bq$bq.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_bq$bq.json'))
print("value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), {k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_forward"].items() if 'attention' in k})
PY
done
