# round 2: training-step tests (gradient parity vs oracle autograd, AdamW, dropout, fit) + the two files that changed semantics
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -s -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
run train tests/test_train.py
run struct tests/test_structure_model.py
run fwd tests/test_gpu_forward.py -k "denoise_loop or philox or consecutive"
cat gpurun_out/summary.txt
grep -h "rel err\|losses\|derivative\|Error\|error" gpurun_out/train.log | head -40
