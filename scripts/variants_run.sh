#!/bin/bash
# variants_run.sh v1 v2 ... : splat parity test + perf probe for each build_variants/<v>.so ("main" = the in-tree library).
# Every step runs under its own timeout so a hanging variant cannot eat the GPU budget.
for v in "$@"; do
  if [ "$v" = main ]; then unset VOXPLAT_B200_LIB; else export VOXPLAT_B200_LIB=$PWD/build_variants/$v.so; fi
  echo "=== $v"
  timeout 90 python -m pytest tests/test_gpu_splat.py tests/test_gpu_golden.py tests/test_gpu_edges.py -x -q 2>&1 | tail -1
  timeout 60 python scripts/perf_probe.py 2>&1 | tail -2
done
