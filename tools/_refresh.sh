set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "configs_3_4_5" > gpurun_out/r2f_gputests_b.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_gputests_b.log
timeout 600 python bench.py --config 3 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
tail -5 gpurun_out/r2f_gputests_b.log; cat gpurun_out/r2f_bench_c3.json | cut -c1-600
