#!/bin/bash
# gpu_run.sh TAG : one GPU-box session -- parity tests, perf probe, optional ncu capture of the splat / mesh kernels.
# Every step runs under its own timeout so a hanging kernel cannot eat the GPU budget.
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
FLAGS=1 timeout 120 python scripts/perf_probe.py > gpurun_out/${TAG}_perf_splat.log 2>&1; tail -3 gpurun_out/${TAG}_perf_splat.log
if [ -n "$NCU" ]; then
  FLAGS=${NCU_FLAGS:-1} STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_splat}" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -f -o gpurun_out/${TAG}_full python scripts/perf_probe.py > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log
fi
