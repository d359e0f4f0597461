set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "attention" > gpurun_out/ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/ops.log
timeout 300 python scripts/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1; cat gpurun_out/attn_sweep.log
ATTN_RAGGED=1 timeout 300 python scripts/attn_sweep.py > gpurun_out/attn_sweep_ragged.log 2>&1; echo ragged; cat gpurun_out/attn_sweep_ragged.log
