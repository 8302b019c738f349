set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
MTM_B200_PROF=1 timeout 120 python - <<'PY' 2>&1 | tail -6
import numpy as np, MTM
from oracle import synth
from mtm_b200 import _native
image, temps, params = synth.config("C2")
ctx = _native.Context(0)
for i in range(2):
    hits = MTM.matchTemplates(temps, image, context=ctx, **params)
print(len(hits))
PY
timeout 300 python bench.py --cpu-steps 3 > gpurun_out/bench_r1_ts2.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_ts2.json
MTM_B200_NO_TS=1 timeout 300 python bench.py --cpu-steps 3 > gpurun_out/bench_r1_ss2.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_ss2.json
timeout 300 python bench.py --workload C3 --steps 30 --cpu-steps 1 > gpurun_out/bench_r1_ts2_C3.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_ts2_C3.json
timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_ts2_C4.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_ts2_C4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_ts2.csv python bench.py --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_bench.log 2>&1
