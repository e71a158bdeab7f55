NCU=1 NCU_K="k_splat|k_mesh" NCU_COUNT=4 NCU_FLAGS=3 bash scripts/gpu_run.sh r2h
bash scripts/variants_run.sh pf1 > gpurun_out/r2h_variants.log 2>&1
cat gpurun_out/r2h_variants.log
FLAGS=3 timeout 200 python scripts/perf_probe.py > gpurun_out/r2h_perf_meshall.log 2>&1; tail -3 gpurun_out/r2h_perf_meshall.log
