set -x
mkdir -p gpurun_out
timeout 300 python bench.py --cpu-steps 1 > gpurun_out/bench_r1_v9.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v9.json
timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_v9_C4.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v9_C4.json
timeout 300 python bench.py --workload C3 --steps 30 --cpu-steps 1 > gpurun_out/bench_r1_v9_C3.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v9_C3.json
