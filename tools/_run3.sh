set -x
L=gpurun_out/r2_gen_time11.log
echo base > $L; python tools/prof_generate.py generate 8 >> $L 2>&1
cp lutgen_rs_b200/liblutgen_cuda.so /tmp/lib_a.so
for v in a b c; do
cp tools/_variant_$v.so lutgen_rs_b200/liblutgen_cuda.so; touch lutgen_rs_b200/liblutgen_cuda.so
echo variant_$v >> $L; python tools/prof_generate.py generate 8 >> $L 2>&1
done
cp tools/_variant_a.so lutgen_rs_b200/liblutgen_cuda.so; touch lutgen_rs_b200/liblutgen_cuda.so
ncu --set full --clock-control none --import-source on -k regex:k_remap_warp -s 2 -c 1 -o gpurun_out/r2_warp11a -f python tools/prof_generate.py generate 3 > gpurun_out/r2_ncu_warp11.log 2>&1
cp /tmp/lib_a.so lutgen_rs_b200/liblutgen_cuda.so; touch lutgen_rs_b200/liblutgen_cuda.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1
tail -5 gpurun_out/r2_pytest11.log
