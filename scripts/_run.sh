#!/bin/bash
# Scratch script for `gpurun -- 'bash scripts/_run.sh'` calls.  RULE (learned the hard way in round 2):
# every multi-GPU command gets its OWN tight timeout -- the box is charged N x wall time, and a deadlocked NCCL job
# otherwise sits in the watchdog for ten minutes per launch.
N=${1:-2}
mkdir -p gpurun_out
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err; echo "bench$N rc=$?"
