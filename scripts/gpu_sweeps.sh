set +e
mkdir -p gpurun_out
timeout 300 python scripts/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1; cat gpurun_out/attn_sweep.log
timeout 600 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; cat gpurun_out/gemm_sweep.log
