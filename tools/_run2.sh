set -x
python tools/prof_generate.py generate 8 > gpurun_out/r2_gen_time6.log 2>&1
LUTGEN_CUDA_STATS=1 python tools/prof_generate.py generate 1 > gpurun_out/r2_gen_stats6.log 2>&1
python tools/prof_generate.py nn 4 >> gpurun_out/r2_gen_time6.log 2>&1
python tools/prof_generate.py shepard256 4 >> gpurun_out/r2_gen_time6.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1
tail -5 gpurun_out/r2_pytest6.log
ncu --set full --clock-control none --import-source on -k regex:k_remap_warp -s 2 -c 1 -o gpurun_out/r2_warp6 -f python tools/prof_generate.py generate 3 > gpurun_out/r2_ncu_warp6.log 2>&1
# variant: 4 CTAs per SM (64 registers)
cp lutgen_rs_b200/liblutgen_cuda.so /tmp/lib_a.so
cp tools/_variant_minb4.so lutgen_rs_b200/liblutgen_cuda.so
touch lutgen_rs_b200/liblutgen_cuda.so
python tools/prof_generate.py generate 8 > gpurun_out/r2_gen_time6_minb4.log 2>&1
cp /tmp/lib_a.so lutgen_rs_b200/liblutgen_cuda.so
