set +e
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/train_ddp_check.py > gpurun_out/train_ddp_check_2gpu_r02.json 2> gpurun_out/train_ddp_check.err; echo "ddp check exit $?"
cat gpurun_out/train_ddp_check_2gpu_r02.json; tail -3 gpurun_out/train_ddp_check.err
