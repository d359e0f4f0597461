set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/fwd.log
grep -E "rel err|L2 rel" gpurun_out/fwd.log | sed 's/^[.F]*//' | grep -E "checkpoint|autocast" 
timeout 600 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; cat gpurun_out/gemm_sweep.log
