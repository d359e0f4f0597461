# end-of-round checks on one B200 after the last training-step changes: full GPU suite (product build), smoke, headline bench, reference arm
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $? : $(tail -1 gpurun_out/$name.log)" >> gpurun_out/summary.txt; }
run decloss tests/test_decode_loss.py
run struct tests/test_structure_model.py
run ops tests/test_gpu_ops.py
run fwd tests/test_gpu_forward.py
run train tests/test_train.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $? : $(tail -1 gpurun_out/smoke.log)" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_final2.json 2>/dev/null; echo "reference arm exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_final2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'packed', d['packed']['value'], 'clocks', d['clocks'])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], 'share', r['share_of_forward']); ri=d['roofline_isolated']; print('iso', ri.get('achieved'), ri.get('frac'), ri.get('frac_of_two_floor_bound'))
print('rev', d['reverse_step_roofline']['avg_launch_us'], d['reverse_step_roofline']['frac'])
print('cfg1', d.get('cfg1_latency'))
c=d['cfg4_train']; print('cfg4', {k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','launches_per_step','error')})
c=d['cfg3_strong']; print('cfg3', {k:c.get(k) for k in ('value','ms_per_sampling','error')}, (c.get('packed') or {}).get('value'))
print('cpu', d.get('cpu_baseline'))
PY
cut -c1-400 gpurun_out/bench_ref_r02_final2.json
