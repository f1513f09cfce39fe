#!/bin/bash
# gpurun --gpus N: data-parallel parity test and bench lines (fused peer-memory optimizer; NCCL path for comparison).
# usage: scripts/collect_evidence_multi.sh <N> [tag]
N=$1; tag=${2:-r02f}; out=gpurun_out
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -8 > $out/${tag}_dp${N}_test.log
cat $out/${tag}_dp${N}_test.log
$run --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --skip-3xtf32 > $out/${tag}_bench_n${N}_4096_peer_optimizer.json 2> $out/${tag}_bench_n${N}_4096_peer_optimizer.err
$run --master-port 29522 bench.py --gpus $N --envs 8192 --steps 30 --warmup 5 --no-cpu-baseline --skip-3xtf32 > $out/${tag}_bench_n${N}_8192_peer_optimizer.json 2> $out/${tag}_bench_n${N}_8192_peer_optimizer.err
HB_DP_FUSED=0 $run --master-port 29523 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --skip-3xtf32 > $out/${tag}_bench_n${N}_4096_nccl.json 2> $out/${tag}_bench_n${N}_4096_nccl.err
for f in 4096_peer_optimizer 8192_peer_optimizer 4096_nccl; do
  python - $out/${tag}_bench_n${N}_$f.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); p = d["ppo"]
print(sys.argv[1], "value %.1f M" % (d["value"] / 1e6), "ppo %.1f M samples/s, %.2f ms/update (%s)" % (p["value"] / 1e6, p["ms_per_update"], p["gradient_exchange"]),
      "e2e %.2f M" % (d["e2e"]["value"] / 1e6))
PY
done
