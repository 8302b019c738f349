set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -40
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 300 python bench.py --path tensor > gpurun_out/bench_r1_tc.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_tc.json
timeout 300 python bench.py --path tensor --workload C3 --steps 30 --warmup 3 --cpu-steps 2 > gpurun_out/bench_r1_tc_C3.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_tc_C3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tc.csv python bench.py --path tensor --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncc_tc -s 3 -c 2 -o gpurun_out/prof_tc python bench.py --path tensor --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
