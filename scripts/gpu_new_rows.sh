# GPU pass for the rows added late in round 1: structure model, Gaussian step, decode, loss terms (+ the whole suite)
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -25 gpurun_out/$name.log; }
run struct tests/test_structure_model.py -x
run decloss tests/test_decode_loss.py
cat gpurun_out/summary.txt
