set +e
timeout 900 python -m pytest tests/test_train.py -q -m gpu --timeout 600 -p no:cacheprovider -k "resume or fit" 2>&1 | tail -15
