set +e
mkdir -p gpurun_out
python scripts/kernel_floor.py 2>&1 | tee gpurun_out/kernel_floor_r02.log
TM=128 TN=768 TK=768 TCFG=128 TLIM=80 python scripts/gemm_trace.py 2>&1 | tee gpurun_out/gemm_trace_m128_r02.log | head -70
