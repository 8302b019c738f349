set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_C2.csv python bench.py --steps 3 --warmup 3 --cpu-steps 1 --contexts 1 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ncc_tc_kernel -s 6 -c 2 -o gpurun_out/prof_tc_final python bench.py --steps 3 --warmup 3 --cpu-steps 1 --contexts 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"peaks2d|finalize_small|sat_cols|window_moments" -s 8 -c 4 -o gpurun_out/prof_small_final python bench.py --steps 3 --warmup 3 --cpu-steps 1 --contexts 1 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -8
