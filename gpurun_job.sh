mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 200 python bench.py --cpu-steps 1 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'ms/step %.4f sync %.4f e2e %.4f frac %.3f kern_ms %.4f d2h %d'%(d['ms_per_step'], d['sync_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['e2e']['d2h_bytes_per_step']))"; }
run a ""
run a "--contexts 4"
run a "--contexts 1"
run a "--workload C4"
MTM_B200_STAGES=1 timeout 100 python - <<'PY' 2>&1 | tail -3
import sys, os
sys.path.insert(0, os.getcwd())
import MTM
from oracle import synth
image, temps, params = synth.config("C2")
for _ in range(4):
    MTM.matchTemplates(temps, image, **params)
PY
