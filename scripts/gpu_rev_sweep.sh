set +e
mkdir -p gpurun_out
: > gpurun_out/rev_sweep.log
for c in ${REV_CFGS:-25622 25621 25632 25631 12841 12842 12851 12852 12861 12862 6431 6432 6434}; do
  echo "cfg $c" >> gpurun_out/rev_sweep.log
  SEQDIFF_REV_CFG=$c timeout 120 python scripts/revstep_bench.py 2>&1 | head -1 >> gpurun_out/rev_sweep.log
done
cat gpurun_out/rev_sweep.log
