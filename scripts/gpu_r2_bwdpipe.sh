set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train.py -q -m gpu --timeout 300 -p no:cacheprovider -k "attention_train_kernels" -x 2>&1 | tail -25
