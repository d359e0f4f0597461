# full GPU test pass, one pytest process per group (a faulting kernel poisons only its own CUDA context)
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1200 python -m pytest "$@" -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py -s
cat gpurun_out/summary.txt
