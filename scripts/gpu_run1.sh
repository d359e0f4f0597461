set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
run gemm_tc tests/test_gpu_ops.py -k "gemm_tcgen05" -x
run gemm_f32 tests/test_gpu_ops.py -k "gemm_fp32"
run attn tests/test_gpu_ops.py -k "attention"
run rev tests/test_gpu_ops.py -k "reverse or philox or aa_noise"
run fwd tests/test_gpu_forward.py -s
cat gpurun_out/summary.txt
