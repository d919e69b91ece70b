set -x
L=gpurun_out/r2_gen_time13.log
python tools/prof_generate.py generate 6 > $L 2>&1
python tools/prof_generate.py shepard256 5 >> $L 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1
tail -5 gpurun_out/r2_pytest14.log
