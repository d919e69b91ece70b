set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1
tail -5 gpurun_out/r2_pytest12.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python tools/prof_generate.py apply 4 > gpurun_out/r2_apply_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 1 -c 1 -o gpurun_out/r2_apply1 -f python tools/prof_generate.py apply 2 > gpurun_out/r2_ncu_apply1.log 2>&1
