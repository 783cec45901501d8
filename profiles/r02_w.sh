# round 2, call w: whole GPU suite + smoke() + examples + the driver-form lines on the tree as committed
set -x
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -6 | tee gpurun_out/r02_w_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python examples/policy_in_the_loop.py 2>&1 | tail -2 | tee gpurun_out/r02_w_policy_loop.txt
timeout 300 python examples/policy_in_the_loop.py --workload mobile-medium-ma-v0 2>&1 | tail -1 | tee -a gpurun_out/r02_w_policy_loop.txt
timeout 300 python examples/layout_search.py 2>&1 | tail -3
timeout 300 python examples/collect_data.py --epochs 256 --out /tmp/mbe_collect 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_w_bench_reference_arm.json 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_w_bench_driver.json 2>gpurun_out/r02_w_err.txt || tail -5 gpurun_out/r02_w_err.txt
python - <<'PY'
import json
for n in ("driver","reference_arm"):
    d=json.loads(open(f"gpurun_out/r02_w_bench_{n}.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(n, d["value"], "%.3f us"%(d["ms_per_step"]*1e3), "frac", r.get("frac"), "layout", r.get("frac_layout"), "traffic", r.get("traffic"), "e2e %.4g"%d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("kind"), d.get("clocks"))
PY
