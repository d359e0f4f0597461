set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python scripts/train_steptimes.py --batch 128 --steps 10 2>&1 | tail -10 | tee gpurun_out/steptimes128.log
python scripts/train_steptimes.py --batch 16 --steps 8 2>&1 | tail -5 | tee gpurun_out/steptimes16.log
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
run decloss tests/test_decode_loss.py
run struct tests/test_structure_model.py
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py
run train tests/test_train.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
cat gpurun_out/summary.txt
