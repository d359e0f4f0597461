set +e
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_r02.json 2> gpurun_out/bench_2gpu_r02.err
echo "bench2 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_r02.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ("value","n_gpus","ms_per_step","e2e","cfg3_strong","cfg4_train")}, indent=1)[:4000])
PY
tail -5 gpurun_out/bench_2gpu_r02.err
