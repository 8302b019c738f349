set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_n2_err.log; tail -5 gpurun_out/bench_n2_err.log; cat gpurun_out/bench_r1_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 5 --cpu-steps 1 > gpurun_out/bench_r1_n1.json 2>/dev/null; cat gpurun_out/bench_r1_n1.json
