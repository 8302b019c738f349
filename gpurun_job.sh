set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
cat > rgbrun_tmp.py <<'PY'
import numpy as np, MTM, time
from oracle import synth, mtm_port
from mtm_b200 import _native
image, temps, params = synth.config("C2")
rgb = np.ascontiguousarray(np.stack([image, np.roll(image, 3, 1), np.roll(image, 5, 0)], axis=-1))
trgb = [(n, np.ascontiguousarray(np.stack([t, np.roll(t, 3, 1), np.roll(t, 5, 0)], axis=-1))) for n, t in temps]
for path in (_native.PATH_AUTO, _native.PATH_DIRECT):
    ctx = _native.Context(0); ctx.set_path(path)
    for i in range(3): hits = MTM.matchTemplates(trgb, rgb, context=ctx, **params)
    ctx.timer_begin()
    for i in range(10): hits = MTM.matchTemplates(trgb, rgb, context=ctx, **params)
    print("path", path, "ms/step", ctx.timer_end() / 10, len(hits))
t0 = time.perf_counter(); want = mtm_port.match_templates(trgb, rgb, **params); print("cpu ms", (time.perf_counter() - t0) * 1e3, len(want))
PY
timeout 300 python rgbrun_tmp.py 2>&1 | tail -4; rm -f rgbrun_tmp.py
