set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "gemm" > gpurun_out/ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/ops.log
TCFG=192 python scripts/gemm_trace.py > gpurun_out/gtrace_192.log 2>&1; grep -h "M=\|epi\|committed" gpurun_out/gtrace_192.log
TM=16384 TN=4608 TCFG=256 TLIM=600 python scripts/gemm_trace.py > gpurun_out/gtrace_big.log 2>&1; grep "committed" gpurun_out/gtrace_big.log | head -5
timeout 600 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; cat gpurun_out/gemm_sweep.log
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/fwd.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_iter.json'))
    print("value", round(d["value"]), d["unit"], "ms/sampling", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "gemm TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "iso", d.get("roofline_isolated"), "clocks", d["clocks"])
    print({k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_forward"].items()})
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_iter.err').read()[-2000:])
PY
