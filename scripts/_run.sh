mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multictx.py tests/test_gpu_dropin.py -q -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests.log; tail -3 gpurun_out/r2q_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2q_multi_gpu_check.log 2>&1; tail -2 gpurun_out/r2q_multi_gpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2q_bench2.err
