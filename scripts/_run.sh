mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2u_topo8.txt 2>&1
free -g | head -2 > gpurun_out/r2u_mem.txt; nproc >> gpurun_out/r2u_mem.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2u_bench8.json 2> gpurun_out/r2u_bench8.err; echo "bench8 rc=$?"
grep -v "^chunkset.c\|^mem.c" gpurun_out/r2u_bench8.err | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2u_bench4.json 2> gpurun_out/r2u_bench4.err; echo "bench4 rc=$?"
