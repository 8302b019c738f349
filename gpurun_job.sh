mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_r1_p3.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r1_p3_reference.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --workload C4 --steps 50 --cpu-steps 2 > gpurun_out/bench_r1_p3_C4.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --workload C5 --contexts 1 --steps 16 --warmup 3 --cpu-steps 1 > gpurun_out/bench_r1_p3_C5.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --workload C3 --steps 30 --cpu-steps 2 > gpurun_out/bench_r1_p3_C3.json 2>> gpurun_out/bench_err.log
MTM_B200_STAGES=1 timeout 100 python - <<'PY' 2>&1 | tail -2
import sys, os
sys.path.insert(0, os.getcwd())
import MTM
from oracle import synth
image, temps, params = synth.config("C2")
for _ in range(4):
    MTM.matchTemplates(temps, image, **params)
PY
python - <<'PY'
import json
for n in ['','_C3','_C4','_C5']:
    d=json.loads(open('gpurun_out/bench_r1_p3%s.json'%n).read().strip().splitlines()[-1])
    print(n, 'ms/step %.4f sync %.4f e2e %.4f frac %.3f kern_ms %.4f share %.2f'%(d['ms_per_step'], d['sync_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step']))
d=json.loads(open('gpurun_out/bench_r1_p3_reference.json').read().strip().splitlines()[-1]); print('ref', d['ms_per_step'], d['value'])
PY
