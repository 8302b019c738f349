set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
cat > c2run_tmp.py <<'PY'
import numpy as np, MTM
from oracle import synth
from mtm_b200 import _native
image, temps, params = synth.config("C2")
ctx = _native.Context(0)
for i in range(3):
    hits = MTM.matchTemplates(temps, image, context=ctx, **params)
print(len(hits))
PY
MTM_B200_STAGES=1 timeout 120 python c2run_tmp.py 2>&1 | tail -3
MTM_B200_TS=1 MTM_B200_PROF=1 timeout 120 python c2run_tmp.py 2>&1 | grep "mtm prof" | tail -1
MTM_B200_TS=1 timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -3
rm -f c2run_tmp.py
timeout 300 python bench.py --cpu-steps 3 > gpurun_out/bench_r1_v6.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v6.json
MTM_B200_TS=1 timeout 300 python bench.py --cpu-steps 1 > gpurun_out/bench_r1_v6_ts.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v6_ts.json
MTM_B200_TS=1 timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 1 > gpurun_out/bench_r1_v6_C4ts.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_r1_v6_C4ts.json
