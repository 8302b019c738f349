set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --cpu-steps 1 > gpurun_out/bench_r1_v12.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v12.json
timeout 300 python bench.py --contexts 1 --cpu-steps 1 > gpurun_out/bench_r1_v12_c1.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v12_c1.json
timeout 300 python bench.py --workload C5 --contexts 1 --steps 16 --warmup 3 --cpu-steps 1 > gpurun_out/bench_r1_v12_C5.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v12_C5.json
