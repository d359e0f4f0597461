set +e
mkdir -p gpurun_out
SEQDIFF_DEBUG_BOUNDS=1 timeout 600 python -m pytest tests/test_decode_loss.py -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/dbg_decode.log
SEQDIFF_DEBUG_BOUNDS=1 timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/dbg_fwd.log
cat gpurun_out/dbg_decode.log; cat gpurun_out/dbg_fwd.log
