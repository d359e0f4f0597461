set +e
mkdir -p gpurun_out
export SEQDIFF_GEMM_TUNE_LOG=1
( python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras > /dev/null
  python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras --batch 1 > /dev/null
  python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras --batch 8 > /dev/null
  python bench.py --steps 1 --warmup 1 --timesteps 2 --no-extras --workload cfg3 > /dev/null
  python bench.py --steps 1 --warmup 1 --timesteps 2 --no-extras --workload cfg3 --batch 32 > /dev/null
  python scripts/struct_bench.py --steps 1 --warmup 1 --timesteps 4 --no-cpu > /dev/null
  python bench.py --steps 1 --warmup 1 --timesteps 4 --no-extras --precision fp16 > /dev/null
  python scripts/train_profile.py --batch 128 > /dev/null
  python scripts/train_profile.py --batch 16 > /dev/null ) 2>&1 | grep "gemm tune" | sort | uniq > gpurun_out/gemm_tune_selections_r02.log
wc -l gpurun_out/gemm_tune_selections_r02.log
awk '{print $(NF-3), $(NF-2)}' gpurun_out/gemm_tune_selections_r02.log | sort | uniq -c
