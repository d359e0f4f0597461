set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -q -s -m gpu -k "packed" -p no:cacheprovider > gpurun_out/pack.log 2>&1; echo "pack exit $?"
grep -h "valid of\|Error\|error\|FAILED\|passed\|failed" gpurun_out/pack.log | head -40
