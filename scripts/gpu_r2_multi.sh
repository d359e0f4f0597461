# multi-GPU records of round 2: bench.py (cfg2 weak + cfg3 strong + cfg4 strong / weak) and the sharded structure sampler, N = $1 GPUs
set +e
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_r02.json 2> gpurun_out/bench_${N}gpu_r02.err; echo "bench $N exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu_r02.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'packed', d['packed']['value'])
c=d['cfg3_strong']; print('cfg3', {k:c.get(k) for k in ('value','ms_per_sampling','graphs_per_gpu','error')}, (c.get('packed') or {}).get('value'))
c=d['cfg4_train']; print('cfg4', {k:c.get(k) for k in ('value','ms_per_step','ms_per_step_without_allreduce','ms_per_step_allreduce_after_backward','ms_allreduce_alone','exposed_allreduce_frac','allreduce_bus_GBps','error')})
w=c.get('weak_128_per_gpu') or {}; print('cfg4 weak', {k:w.get(k) for k in ('value','ms_per_step','exposed_allreduce_frac','global_batch','error')})
PY
tail -3 gpurun_out/bench_${N}gpu_r02.err
timeout 900 $TR --master-port 29522 scripts/struct_sharded_bench.py > gpurun_out/struct_sharded_${N}gpu_r02.json 2> gpurun_out/struct_sharded_${N}gpu_r02.err; echo "struct sharded $N exit $?"
cat gpurun_out/struct_sharded_${N}gpu_r02.json; tail -3 gpurun_out/struct_sharded_${N}gpu_r02.err
