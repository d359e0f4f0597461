set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 600 -p no:cacheprovider -k "reverse or philox or aa_noise" > gpurun_out/ops_rev.log 2>&1; echo "ops_rev exit $?"; tail -3 gpurun_out/ops_rev.log
timeout 300 python scripts/revstep_bench.py > gpurun_out/revstep_r02.log 2>&1; cat gpurun_out/revstep_r02.log
bash scripts/gpu_r2_bench.sh
