set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --cpu-steps 1 > gpurun_out/bench_r1_v11.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v11.json
timeout 300 python bench.py --contexts 1 --cpu-steps 1 > gpurun_out/bench_r1_v11_c1.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v11_c1.json
timeout 300 python bench.py --contexts 3 --cpu-steps 1 > gpurun_out/bench_r1_v11_c3.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v11_c3.json
timeout 300 python bench.py --workload C5 --contexts 1 --steps 16 --warmup 3 --cpu-steps 1 > gpurun_out/bench_r1_v11_C5.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v11_C5.json
timeout 300 python bench.py --workload C4 --contexts 1 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_v11_C4.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v11_C4.json
