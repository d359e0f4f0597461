set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/train.log 2>&1; echo "train exit $?"; tail -3 gpurun_out/train.log
timeout 600 python scripts/train_profile.py > gpurun_out/train_profile_r02b.log 2>&1; head -12 gpurun_out/train_profile_r02b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg4 --steps 8 --warmup 3 > gpurun_out/bench_cfg4_2gpu.json 2> gpurun_out/bench_cfg4_2gpu.err; echo "cfg4 2gpu exit $?"; tail -c 1800 gpurun_out/bench_cfg4_2gpu.json; tail -5 gpurun_out/bench_cfg4_2gpu.err
