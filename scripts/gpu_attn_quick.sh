set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "attention" > gpurun_out/ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/ops.log
TREL=0 TLIM=140 python scripts/attn_trace.py > gpurun_out/trace_norel.log 2>&1; TREL=1 TLIM=140 python scripts/attn_trace.py > gpurun_out/trace_rel.log 2>&1
timeout 300 python scripts/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1; cat gpurun_out/attn_sweep.log
