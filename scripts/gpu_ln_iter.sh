set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "fused_layernorm or gemm_tcgen05" > gpurun_out/ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/ops.log
TLN=1 python scripts/gemm_trace.py > gpurun_out/gtrace_ln.log 2>&1; grep -v "tma slot\|stage full" gpurun_out/gtrace_ln.log | tail -32
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/fwd.log
for f in 1 0 1 0; do
  SEQDIFF_LN_FUSE=$f timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ln$f.json 2> gpurun_out/bench_ln$f.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ln$f.json'))
print("LN_FUSE=$f value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), d["clocks"])
PY
done
