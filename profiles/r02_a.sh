set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_a_bench_driver1.json 2> gpurun_out/r02_a_driver1.err; tail -c 400 gpurun_out/r02_a_driver1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_a_bench_driver2.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_a_bench_default.json 2>/dev/null
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072" > gpurun_out/r02_a_variants.txt 2>&1
cat gpurun_out/r02_a_variants.txt | tail -12
python - <<'PY'
import json
for n in ("driver1","driver2","default"):
    try:
        d=json.loads(open(f"gpurun_out/r02_a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["timing"], d["e2e"]["value"], d["clocks"])
    except Exception as e: print(n, "failed", e)
PY
