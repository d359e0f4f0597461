set +e
mkdir -p gpurun_out
: > gpurun_out/ln_sweep.log
for v in 2561 1281 5121 641 2562 1282; do
  echo "SEQDIFF_LN_VAR=$v" >> gpurun_out/ln_sweep.log
  SEQDIFF_LN_VAR=$v timeout 120 python scripts/ln_bench.py >> gpurun_out/ln_sweep.log 2>&1
done
cat gpurun_out/ln_sweep.log
