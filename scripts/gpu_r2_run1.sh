# round 2, run 1: full GPU suite (one pytest process per file), smoke, default bench
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 1500 python -m pytest "$@" -q -s -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
run decloss tests/test_decode_loss.py
run struct tests/test_structure_model.py
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -h "rel err" gpurun_out/fwd.log | head -80
head -c 1500 gpurun_out/bench_r02a.json
