mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log; tail -3 gpurun_out/r2s_tests.log
for b in 8 16 32; do
  VP_BENCH_E2E_BLOCKS=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2s_bench_$b.json 2>/dev/null; echo "blocks $b rc=$?"
done
python - <<'PY'
import json
for p in (8,16,32):
    l=json.loads(open('gpurun_out/r2s_bench_%s.json'%p).read().strip().splitlines()[-1])
    print(p, "step %.4f"%l["ms_per_step"], "e2e %.3f floor %.3f"%(l["e2e"]["ms_per_step"], l["e2e"]["pcie_floor_ms"]), l["parity"]["ok"], l["e2e"]["gpu_launches_per_step"])
PY
