set +e
mkdir -p gpurun_out
export ATTN_SHAPES=1,2
python scripts/attn_sweep.py > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_pipe -s 3 -c 1 -f -o gpurun_out/prof_attn_pipe_rel python scripts/attn_sweep.py > gpurun_out/ncu_attn_rel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_pipe -s 36 -c 1 -f -o gpurun_out/prof_attn_pipe_norel python scripts/attn_sweep.py > gpurun_out/ncu_attn_norel.log 2>&1
cat gpurun_out/plain_attn.log; tail -3 gpurun_out/ncu_attn_rel.log gpurun_out/ncu_attn_norel.log
