mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -2
MTM_B200_EW=16 timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -2
MTM_B200_EW=12 timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q 2>&1 | tail -2
run() { timeout 200 python bench.py --cpu-steps 1 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'ms/step %.4f sync %.4f e2e %.4f frac %.3f kern_ms %.4f'%(d['ms_per_step'], d['sync_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step']))"; }
for ew in 8 12 16; do
  export MTM_B200_EW=$ew
  run "EW=$ew" "--contexts 1 --steps 100"
  run "EW=$ew" "--workload C4 --contexts 1 --steps 40"
  run "EW=$ew" "--workload C5 --contexts 1 --steps 12"
done
unset MTM_B200_EW
run auto ""
run auto "--workload C4"
run auto "--workload C5 --contexts 1 --steps 12"
MTM_B200_EW=8 run "EW=8 2ctx" ""
MTM_B200_EW=12 run "EW=12 2ctx" ""
