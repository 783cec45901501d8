# round 2, call zk: last validation of the tree as committed (after the redux change): whole GPU suite, smoke, driver-form line
set -x
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -6 | tee gpurun_out/r02_zk_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tests/sanitize_smoke.py 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_zk_bench_driver.json 2>gpurun_out/r02_zk_err.txt || tail -5 gpurun_out/r02_zk_err.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_zk_bench_driver.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("driver", "%.4g"%d["value"], "%.3f us"%(d["ms_per_step"]*1e3), "frac %.3f layout %.3f"%(r["frac"], r["frac_layout"]), "e2e %.4g"%d["e2e"]["value"], d["cpu_baseline"]["kind"], "%.4g"%d["cpu_baseline"]["value"], d["clocks"])
PY
