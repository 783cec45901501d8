# round 2, call f: full GPU suite, driver-style bench lines, steady-state ncu traffic, full captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -12 | tee gpurun_out/r02_f_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_f_bench_driver.json 2>gpurun_out/r02_f_err.txt || tail -5 gpurun_out/r02_f_err.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_f_bench_default.json 2>/dev/null
for w in "mobile-medium-ma-v0 131072" "mobile-large-central-v0 262144" "mobile-small-central-v0 65536" "mobile-synthetic-central-v0 16384" "mobile-custom-v0 262144"; do set -- $w
timeout 300 python bench.py --no-cpu-baseline --workload $1 --envs $2 --steps 512 > gpurun_out/r02_f_bench_$1_$2.json 2>/dev/null; done
B="--steps 24 --warmup 8 --no-graph --no-cpu-baseline --preheat-seconds 0 --repeats 1"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum"
for w in "mobile-medium-central-v0 65536 step_upt" "mobile-medium-ma-v0 131072 step_upt" "mobile-large-central-v0 262144 step_spec" "mobile-synthetic-central-v0 16384 step_big"; do set -- $w
timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:$3 --launch-skip 12 --launch-count 8 --metrics $M --csv --log-file gpurun_out/r02_f_traffic_$1.csv python bench.py --workload $1 --envs $2 $B > /dev/null 2>&1
tail -9 gpurun_out/r02_f_traffic_$1.csv | cut -c1-300
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_big --launch-skip 6 --launch-count 1 -o gpurun_out/r02_f_big python bench.py --workload mobile-synthetic-central-v0 --envs 4096 $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_upt --launch-skip 12 --launch-count 1 -o gpurun_out/r02_f_upt python bench.py $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_spec --launch-skip 12 --launch-count 1 -o gpurun_out/r02_f_small python bench.py --workload mobile-small-central-v0 $B > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_f_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph --preheat-seconds 0 --repeats 2 > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_f_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.3f us"%(d["ms_per_step"]*1e3), "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], "lite", (d["e2e"].get("obs_stays_on_device") or {}).get("value"), d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
ls -la gpurun_out | tail -24
