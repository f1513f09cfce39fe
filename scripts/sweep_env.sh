#!/bin/bash
# Env-stage tuning sweep (run under gpurun): one bench line per option set, PPO skipped.
out=gpurun_out/sweep_env.jsonl
: > $out
for envs in 65536 4096; do
  for ctas in 0 2 3 4 6; do
    for unroll in 4 8; do
      echo "{\"envs\": $envs, \"stack_ctas_per_sm\": $ctas, \"stack_unroll\": $unroll}" >> $out
      python bench.py --envs $envs --steps 30 --skip-ppo --no-cpu-baseline --opt stack_ctas_per_sm=$ctas --opt stack_unroll=$unroll >> $out 2>> gpurun_out/sweep_env.err
    done
  done
done
