set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
MTM_B200_TS=1 timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -8
MTM_B200_TS=1 MTM_B200_PROF=1 timeout 120 python - <<'PY' 2>&1 | tail -3
import numpy as np, MTM
from oracle import synth
from mtm_b200 import _native
image, temps, params = synth.config("C2")
ctx = _native.Context(0)
for i in range(2):
    hits = MTM.matchTemplates(temps, image, context=ctx, **params)
print(len(hits))
PY
timeout 300 python bench.py --cpu-steps 3 > gpurun_out/bench_r1_v4.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v4.json
MTM_B200_TS=1 timeout 300 python bench.py --cpu-steps 1 > gpurun_out/bench_r1_v4_ts.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v4_ts.json
timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_v4_C4.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v4_C4.json
MTM_B200_TS=1 timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_v4_C4ts.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v4_C4ts.json
timeout 300 python bench.py --workload C5 --steps 10 --warmup 3 --cpu-steps 1 > gpurun_out/bench_r1_v4_C5.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v4_C5.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_bench.log 2>&1
