mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 200 python bench.py --cpu-steps 1 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'ms/step %.4f sync %.4f e2e %.4f frac %.3f kern_ms %.4f'%(d['ms_per_step'], d['sync_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step']))"; }
run a ""
cat > m_tmp.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import MTM
from mtm_b200 import _native
from oracle import synth, mtm_port
image, temps, params = synth.config("C2")
ct = _native.Context(0); cd = _native.Context(0); cd.set_path(_native.PATH_DIRECT)
for method in (1, 3, 4, 5):
    kw = dict(method=method, N_object=1)
    res = {}
    for name, c in (("tensor", ct), ("direct", cd)):
        for _ in range(3): h = MTM.matchTemplates(temps, image, context=c, **kw)
        c.synchronize(); t0 = time.perf_counter()
        for _ in range(30): h = MTM.matchTemplates(temps, image, context=c, **kw)
        res[name] = ((time.perf_counter() - t0) / 30 * 1e3, h)
    t0 = time.perf_counter(); want = mtm_port.match_templates(temps, image, **kw); cpu = (time.perf_counter() - t0) * 1e3
    ok = [(a[0], tuple(a[1])) for a in res["tensor"][1]] == [(a[0], tuple(a[1])) for a in want]
    print("method %d N_object=1: tensor %.3f ms  direct %.3f ms  cpu port %.1f ms  tensor==direct %s  tensor==port %s" % (method, res["tensor"][0], res["direct"][0], cpu, res["tensor"][1] == res["direct"][1], ok))
PY
timeout 200 python m_tmp.py 2>&1 | tail -5
rm -f m_tmp.py
