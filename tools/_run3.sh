set -x
python tools/prof_generate.py generate 8 > gpurun_out/r2_gen_time7.log 2>&1
LUTGEN_CUDA_WARP_DEEP=1 python tools/prof_generate.py generate 8 >> gpurun_out/r2_gen_time7.log 2>&1
cp lutgen_rs_b200/liblutgen_cuda.so /tmp/lib_a.so
cp tools/_variant_minb2.so lutgen_rs_b200/liblutgen_cuda.so
touch lutgen_rs_b200/liblutgen_cuda.so
python tools/prof_generate.py generate 8 >> gpurun_out/r2_gen_time7.log 2>&1
cp /tmp/lib_a.so lutgen_rs_b200/liblutgen_cuda.so
touch lutgen_rs_b200/liblutgen_cuda.so
ncu --set full --clock-control none --import-source on -k regex:k_remap_warp -s 2 -c 1 -o gpurun_out/r2_warp7deep -f env LUTGEN_CUDA_WARP_DEEP=1 python tools/prof_generate.py generate 3 > gpurun_out/r2_ncu_warp7.log 2>&1
