set -x
timeout 600 python -m pytest tests/test_gpu_png_decode.py -m gpu -x -q > gpurun_out/r2_pytest_pngd6.log 2>&1
tail -3 gpurun_out/r2_pytest_pngd6.log
python tools/time_png_decode.py both 16 > gpurun_out/r2_pngd_time6.log 2>&1
cat gpurun_out/r2_pngd_time6.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_pngd_launches_own5.csv python tools/time_png_decode.py own 16 > /dev/null 2>&1
