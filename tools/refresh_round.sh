#!/bin/bash
# Round-end refresh on the GPU box: full GPU suite, smoke, one bench line per BASELINE config (N = 1) and the reference arm.
#   gpurun --timeout 3000 -- 'bash tools/refresh_round.sh'   ->   gpurun_out/final_*
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_gputests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_gputests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
for c in 1 3 4 5; do timeout 600 python bench.py --config $c > gpurun_out/final_bench_c$c.json 2> gpurun_out/final_bench_c$c.err; done
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -3 gpurun_out/final_gputests.log; tail -2 gpurun_out/final_smoke.log
