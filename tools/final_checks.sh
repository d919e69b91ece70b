#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:k_blur_axis -s 4 -c 1 -f -o gpurun_out/r2f_blur python tools/prof_generate.py blur 2 > gpurun_out/ncu_blur_f.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2f_blur_launches.csv python tools/prof_generate.py blur 2 > /dev/null 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
