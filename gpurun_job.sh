mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_uint16.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
cat > m_tmp.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import MTM
from mtm_b200 import _native
from oracle import synth
image, temps, params = synth.config("C2")
img16 = (image.astype(np.uint16) * 200 + 1000).astype(np.uint16)
t16 = [(n, (t.astype(np.uint16) * 200 + 1000).astype(np.uint16)) for n, t in temps]
ct = _native.Context(0); cd = _native.Context(0); cd.set_path(_native.PATH_DIRECT)
res = {}
for name, c in (("tensor", ct), ("fp32", cd)):
    for _ in range(3): h = MTM.matchTemplates(t16, img16, context=c, **params)
    c.synchronize(); t0 = time.perf_counter()
    for _ in range(20): h = MTM.matchTemplates(t16, img16, context=c, **params)
    res[name] = ((time.perf_counter() - t0) / 20 * 1e3, h)
same = [(a[0], a[1]) for a in res["tensor"][1]] == [(a[0], a[1]) for a in res["fp32"][1]]
print("uint16 C2-shaped: tensor %.3f ms  fp32 kernel %.3f ms  hits %d  same %s" % (res["tensor"][0], res["fp32"][0], len(res["tensor"][1]), same))
PY
timeout 200 python m_tmp.py 2>&1 | tail -3
rm -f m_tmp.py
timeout 200 python bench.py --cpu-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 u8', 'ms/step %.4f sync %.4f e2e %.4f frac %.3f kern_ms %.4f'%(d['ms_per_step'], d['sync_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step']))"
