set -x
python tools/prof_generate.py generate 5 > gpurun_out/r2_gen_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_remap_warp -s 2 -c 1 -o gpurun_out/r2_warp0 -f python tools/prof_generate.py generate 3 > gpurun_out/r2_ncu_warp0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 1 -c 1 -o gpurun_out/r2_apply0 -f python tools/prof_generate.py apply 2 > gpurun_out/r2_ncu_apply0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rbf_exact -s 2 -c 1 -o gpurun_out/r2_exact0 -f python tools/prof_generate.py generate 3 > gpurun_out/r2_ncu_exact0.log 2>&1
nvidia-smi > gpurun_out/r2_smi.log 2>&1
