#!/bin/bash
# round-2 profile captures (each after the same command ran clean without ncu in this round's earlier calls)
set -x
cd "$GRAFT_REPO_ROOT"
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:k_remap_warp -s 2 -c 1 -f -o gpurun_out/r2b_remap python tools/prof_generate.py generate 4 > gpurun_out/ncu_remap.log 2>&1
$NCU --set full --import-source on -k regex:k_png_deflate -s 1 -c 1 -f -o gpurun_out/r2b_png python tools/prof_png.py > gpurun_out/ncu_png.log 2>&1
$NCU --set full --import-source on -k regex:k_blur_axis -s 4 -c 1 -f -o gpurun_out/r2b_blur python tools/prof_generate.py blur 2 > gpurun_out/ncu_blur.log 2>&1
$NCU --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --csv --log-file gpurun_out/r2b_png_launches.csv python tools/prof_png.py > gpurun_out/ncu_png2.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2b_shepard256_launches.csv python tools/prof_generate.py shepard256 2 > gpurun_out/ncu_shep.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2b_blur_launches.csv python tools/prof_generate.py blur 2 > gpurun_out/ncu_blurl.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
