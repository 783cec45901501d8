# round 2, call ze: final records of the tree -- whole GPU suite, bench lines of every workload, steady-state traffic and a
# full capture of the tiled block-per-env kernel
set -x
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -6 | tee gpurun_out/r02_ze_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_ze_bench_driver.json 2>gpurun_out/r02_ze_err.txt || tail -5 gpurun_out/r02_ze_err.txt
for w in "mobile-medium-ma-v0 131072" "mobile-large-central-v0 262144" "mobile-large-ma-v0 131072" "mobile-synthetic-central-v0 16384" "mobile-synthetic-ma-v0 8192" "mobile-custom-v0 262144"; do set -- $w
timeout 400 python bench.py --workload $1 --envs $2 --steps 512 --no-cpu-baseline > gpurun_out/r02_ze_bench_$1_$2.json 2>/dev/null; done
B="--steps 24 --warmup 8 --no-graph --no-cpu-baseline --preheat-seconds 0 --repeats 1"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum"
timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:step_big --launch-skip 12 --launch-count 8 --metrics $M --csv --log-file gpurun_out/r02_ze_traffic_mobile-synthetic-central-v0.csv python bench.py --workload mobile-synthetic-central-v0 --envs 16384 $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_big --launch-skip 6 --launch-count 1 -o gpurun_out/r02_ze_big python bench.py --workload mobile-synthetic-central-v0 --envs 4096 $B > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_ze_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "%.2f us"%(d["ms_per_step"]*1e3), "frac %.3f layout %.3f"%(r["frac"], r["frac_layout"]), "%.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(f, "failed", e)
PY
