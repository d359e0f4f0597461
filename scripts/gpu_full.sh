# full GPU regression + both bench workloads
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/$name.log; }
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py -s
grep -E "cfg3 L=512" gpurun_out/fwd.log | sed 's/^[.F]*//'
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 exit $?"
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench cfg3 exit $?"
python - <<'PY'
import json
for n in ("cfg2","cfg3"):
    try:
        d=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, "value", round(d["value"]), "ms/sampling", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "gemm TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "step-frac", round(d["roofline"]["whole_step_model_flops_frac"],3), d.get("cpu_baseline",{}).get("value"))
        print("   ", {k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_forward"].items()})
        print("   rev", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["reverse_step_roofline"].items() if k!='note'})
    except Exception as e:
        print(n, "parse failed", e); print(open(f'gpurun_out/bench_{n}.err').read()[-1500:])
PY
