# round 2, call i: full GPU suite on the current tree; bench lines (driver form and default) for every workload
set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -12 | tee gpurun_out/r02_i_pytest.txt
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_i_bench_reference_arm.json 2>gpurun_out/r02_i_err.txt || tail -5 gpurun_out/r02_i_err.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_i_bench_driver.json 2>gpurun_out/r02_i_err.txt || tail -5 gpurun_out/r02_i_err.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_i_bench_default.json 2>/dev/null
for w in "mobile-medium-ma-v0 131072" "mobile-large-central-v0 262144" "mobile-large-ma-v0 131072" "mobile-small-central-v0 65536" "mobile-synthetic-central-v0 16384" "mobile-custom-v0 262144"; do set -- $w
timeout 400 python bench.py --workload $1 --envs $2 --steps 512 --cpu-seconds 2 > gpurun_out/r02_i_bench_$1_$2.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_i_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        cb=d.get("cpu_baseline") or {}
        print(f, "%.3f us"%(d["ms_per_step"]*1e3), "frac %s"%(d.get("roofline") or {}).get("frac"), "e2e %.4g"%d["e2e"]["value"], "cpu", cb.get("kind"), cb.get("value"), d.get("clocks"))
    except Exception as e: print(f, "failed", e)
PY
