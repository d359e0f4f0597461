set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "gemm_tcgen05" > gpurun_out/ops.log 2>&1; echo "ops exit $?"; tail -2 gpurun_out/ops.log
for h in 1 0; do echo "L2HINT=$h"; SEQDIFF_GEMM_L2HINT=$h TM=16384 TN=4608 TCFG=256 TLIM=600 python scripts/gemm_trace.py 2>&1 | grep committed | head -5; SEQDIFF_GEMM_L2HINT=$h python scripts/gemm_sweep.py 2>&1 | grep "ada2\|mlp0\|kv_all\|sum"; done
for h in 1 0 1 0; do
  SEQDIFF_GEMM_L2HINT=$h timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_h$h.json 2> gpurun_out/bench_h$h.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_h$h.json'))
print("L2HINT=$h value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), d["clocks"])
PY
done
