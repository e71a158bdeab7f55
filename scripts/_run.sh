mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2p_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_ref.json 2>/dev/null; echo "ref rc=$?"
