set -x
nvidia-smi -L > gpurun_out/r2_multi_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -m gpu -x -q > gpurun_out/r2_pytest_multi.log 2>&1
tail -15 gpurun_out/r2_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
