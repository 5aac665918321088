#!/bin/bash
# ncu captures of the round's kernels on the GPU box (each command first exits 0 WITHOUT ncu, then is captured once):
#   gpurun --timeout 2400 -- 'bash tools/profile_round.sh'   ->   gpurun_out/prof_*.ncu-rep, prof_launches.csv
# then here:  python tools/ncu_summary.py "title=gpurun_out/prof_x.ncu-rep:kernel-regex" ... > profiles/rNN_ncu_kernels.md
mkdir -p gpurun_out
cap() {  # name, kernel regex, launch-skip, command...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 600 "$@" > gpurun_out/prof_$name.plain.log 2>&1 || { echo "$name: plain run failed"; return; }
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/prof_$name.ncu.log 2>&1
  echo "$name rc=$?"
}
cap sweep40 elastic_sweep_kernel 2 python tools/bench_sweep_quick.py 3000 2
cap sweep64 elastic_sweep_kernel 2 python tools/bench_sweep_quick.py 4000 3
ONLY4=1 cap packed packed_solve_kernel 1 python tools/bench_general.py 20000
ONLY5=1 cap block block_solve_kernel 1 python tools/bench_general.py 4000
cap obs elastic_observables 1 python bench.py --samples 2000 --steps 1 --warmup 1 --no-parity
timeout 600 python bench.py --samples 2000 --steps 2 --warmup 1 --no-parity > gpurun_out/prof_bench_samples2000.json 2> gpurun_out/prof_bench_samples2000.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv python bench.py --samples 2000 --steps 2 --warmup 1 --no-parity > gpurun_out/prof_launches.log 2>&1
ls -la gpurun_out/prof_*
