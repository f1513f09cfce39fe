#!/bin/bash
# A/B of a library option over the env stage: scripts/ab_bench_opt.sh <option> "<values>" "<env counts>"
opt=$1; vals=$2; sizes=${3:-4096}
for n in $sizes; do for v in $vals; do
  python bench.py --envs $n --steps 30 --warmup 5 --skip-ppo --no-cpu-baseline --opt $opt=$v 2>/dev/null > /tmp/ab_line.json
  python - "$opt" "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ab_line.json"))
k = d["kernels"]
print(f"envs {d['config']['envs_per_gpu']} {sys.argv[1]}={sys.argv[2]}: {d['ms_per_step']*1e3:.2f} us/step  post {k['post_physics_ms']*1e3:.2f}  stack {k['stack_pair_ms']*1e3:.2f}  pd {k['pd_ms']*1e3:.2f}", flush=True)
PY
done; done
