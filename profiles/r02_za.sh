# round 2, call za: block-per-env kernel with the observation rows leaving as aligned bulk tiles
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "wide or synthetic or big or pf or errors or compact_wire" 2>&1 | tail -3
for w in "mobile-synthetic-central-v0 16384" "mobile-synthetic-ma-v0 8192"; do set -- $w
timeout 300 python bench.py --no-cpu-baseline --workload $1 --envs $2 --steps 512 > gpurun_out/r02_za_bench_$1_$2.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_za_bench*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "%.1f us"%(d["ms_per_step"]*1e3), "frac %.3f layout %.3f"%(d["roofline"]["frac"], d["roofline"]["frac_layout"]), "%.4g env-steps/s"%d["value"])
PY
