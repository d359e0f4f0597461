set +e
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02d.json'))
for k in ('value','e2e','cfg1_latency','clocks'):
    print(k, d.get(k))
print('packed', d['packed']['value'])
r=d['roofline']; print('roofline', r['achieved'], r['frac'])
ri=d['roofline_isolated']; print('iso', ri.get('achieved'), ri.get('frac'), ri.get('frac_of_two_floor_bound'), ri.get('error'))
for s in ri.get('per_shape', []): print(s)
print('rev', d['reverse_step_roofline'])
c=d['cfg4_train']; print('cfg4', {k:c.get(k) for k in ('value','ms_per_step','train_flops_frac_of_peak','error')})
c=d['cfg3_strong']; print('cfg3', {k:c.get(k) for k in ('value','ms_per_sampling','packed','error')})
PY
tail -3 gpurun_out/bench_r02d.err
