set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "medium or specialised or gymref or full_size" 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_driver_graph$i.json 2>/dev/null
MBE_BENCH_GATE_CYCLES=1000000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/r02_b_driver_stream$i.json 2>/dev/null
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_b_default.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --workload mobile-medium-ma-v0 --envs 131072 > gpurun_out/r02_b_ma.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.3f us"%(d["ms_per_step"]*1e3), "frac %.4f"%d["roofline"]["frac"], d["timing"]["block_ms_min"], d["timing"]["block_ms_max"], d["config"]["launch"])
    except Exception as e: print(f, "failed", e)
PY
