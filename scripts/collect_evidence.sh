#!/bin/bash
# One gpurun call (1 GPU) that produces the round's committed evidence under gpurun_out/ (tag = $1, default r02f):
#   GPU test log, bench lines at 4096 (default run, with the CPU arm inside) / 8192 / 16384 / 32768 / 65536 envs, the reference
#   arm, the ncu launch list of a short bench run, and one `ncu --set full` capture of every kernel of the path at 4096 envs.
# Every ncu pass runs after the same command has exited 0 without ncu.  scripts/summarize_ncu.py turns the CSVs into profiles/*.txt.
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/${tag}_gputest.log
cat $out/${tag}_gputest.log
python bench.py > $out/${tag}_bench_n1_4096.json 2> $out/${tag}_bench_n1_4096.err || exit 1
for n in 8192 16384 32768 65536; do
  python bench.py --envs $n --no-cpu-baseline --skip-3xtf32 > $out/${tag}_bench_n1_$n.json 2> $out/${tag}_bench_n1_$n.err || exit 1
done
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference_cpu.json 2> $out/${tag}_bench_reference_cpu.err
short="python bench.py --steps 5 --warmup 3 --ppo-updates 1 --no-cpu-baseline --skip-3xtf32"
$short > /dev/null 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_bench_4096.csv $short > $out/${tag}_ncu_list.log 2>&1
python scripts/profile_r02.py 4096 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/${tag}_full python scripts/profile_r02.py 4096 > $out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_4096.csv 2>> $out/${tag}_ncu_full.log
ls -la $out | tail -20
