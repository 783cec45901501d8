# round 2, call t: step tickets (a step does not wait for the whole previous grid) -- parity, then speed
set -x
timeout 600 python -m pytest tests/test_gpu_more.py -m gpu -q -x -k "step_tickets" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "medium or specialised or gymref or window or state or host" 2>&1 | tail -4
for c in 1 0; do
MBE_CHAIN=$c timeout 300 python bench.py --no-cpu-baseline --steps 1024 > gpurun_out/r02_t_bench_chain$c.json 2>/dev/null
MBE_CHAIN=$c timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_t_bench_driver_chain$c.json 2>/dev/null
MBE_CHAIN=$c timeout 300 python bench.py --no-cpu-baseline --steps 1024 --workload mobile-medium-ma-v0 --envs 131072 > gpurun_out/r02_t_bench_ma_chain$c.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_t_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.3f us"%(d["ms_per_step"]*1e3), "frac %.4f layout %.4f"%(d["roofline"]["frac"], d["roofline"]["frac_layout"]), "two %.2f us"%(d["two_env_groups_in_flight"]["ms_per_step"]*1e3), "e2e %.4g"%d["e2e"]["value"])
    except Exception as e: print(f, "failed", e)
PY
