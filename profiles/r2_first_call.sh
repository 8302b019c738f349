#!/bin/bash
# First GPU call of the next session (one B200, ~12 min of box time): everything that was written or changed without a GPU gets run,
# the two unmeasured knobs get their A/B, and the captures that the epilogue / small-kernel work needs are taken.
#   /usr/local/graft/bin/gpurun --timeout 1100 -- 'bash profiles/r2_first_call.sh'
# Outputs land in gpurun_out/ (merged back); nothing printed under ncu is a bench value.
set +e
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/r2_gpu.txt 2>&1

# 1. the whole -m gpu suite (incl. the tests that have only ever run on the CPU emulation) and the knob parity scripts
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?" > $O/r2_rc.txt
MTM_B200_TEST_KNOBS=1 timeout 200 python -m pytest tests/test_gpu_knobs.py -m gpu -q -p no:cacheprovider > $O/r2_knob_tests.log 2>&1; echo "knob tests rc=$?" >> $O/r2_rc.txt

# 2. A/B of the unmeasured moment kernels (same process settings otherwise): box sums on every config, row walking where sweeps are long
for w in C2 C3 C4 C5; do
  steps=300; [ $w = C3 ] && steps=60; [ $w = C4 ] && steps=100; [ $w = C5 ] && steps=20
  for v in 0 1; do
    MTM_B200_MOM_BOX=$v timeout 120 python bench.py --workload $w --steps $steps --cpu-steps 1 > $O/r2_ab_${w}_box$v.json 2> $O/r2_ab_${w}_box$v.err
  done
done
for v in 0 1; do
  MTM_B200_MOM_ROWS=$v timeout 120 python bench.py --workload C5 --steps 20 --cpu-steps 1 > $O/r2_ab_C5_rows$v.json 2> $O/r2_ab_C5_rows$v.err
done
python - <<'PY' > gpurun_out/r2_ab_summary.txt 2>&1
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ab_*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms/step %.4f  sync %.4f  kernel %.4f  e2e %.1f" % (d["ms_per_step"], d["sync_ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "failed:", e)
PY

# 3. launch lists (cold, serialised: shares of a step, not absolute times) with and without the box-sum route
for v in 0 1; do
  MTM_B200_MOM_BOX=$v timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_C2_box$v.csv \
      python bench.py --steps 2 --warmup 1 --cpu-steps 1 --contexts 1 > /dev/null 2>&1
done

# 4. full captures: the numerator kernel's epilogue where it is the bound (C4), the one-CTA sort/NMS kernel, the new moment kernel
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ncc_tc_persist -s 4 -c 1 -o $O/r2_tc_C4 \
    python bench.py --workload C4 --steps 2 --warmup 2 --cpu-steps 1 --contexts 1 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:finalize_small -s 4 -c 1 -o $O/r2_fin_C2 \
    python bench.py --steps 2 --warmup 2 --cpu-steps 1 --contexts 1 > /dev/null 2>&1
MTM_B200_MOM_BOX=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:box_moments -s 4 -c 1 -o $O/r2_box_C2 \
    python bench.py --steps 2 --warmup 2 --cpu-steps 1 --contexts 1 > /dev/null 2>&1
MTM_B200_MOM_BOX=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:box_moments -s 2 -c 1 -o $O/r2_box_C5 \
    python bench.py --workload C5 --steps 2 --warmup 1 --cpu-steps 1 --contexts 1 > /dev/null 2>&1

# 5. the contract lines of this tree (default settings)
timeout 200 python bench.py > $O/r2_bench_C2.json 2> $O/r2_bench_C2.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 2 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
cat $O/r2_rc.txt; tail -2 $O/r2_gpu_tests.log; tail -2 $O/r2_knob_tests.log; cat $O/r2_ab_summary.txt
