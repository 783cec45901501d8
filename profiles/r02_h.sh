# round 2, call h: gathers-then-move variant of the UEs-per-thread kernel; e2e with the spin crew
set -x
timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "wire or host or medium or specialised" 2>&1 | tail -3
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072" 1024 > gpurun_out/r02_h_variants.txt 2>&1
cat gpurun_out/r02_h_variants.txt
for wire in compact raw; do
MBE_HOST_WIRE=$wire timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r02_h_e2e_central_$wire.json 2>/dev/null
MBE_HOST_WIRE=$wire timeout 300 python bench.py --no-cpu-baseline --steps 200 --workload mobile-medium-ma-v0 --envs 131072 > gpurun_out/r02_h_e2e_ma_$wire.json 2>/dev/null
done
for n in 4 16; do MBE_HOST_THREADS=$n timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r02_h_e2e_central_compact_t$n.json 2>/dev/null; done
for n in 4 16; do MBE_HOST_WINDOWS=$n timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r02_h_e2e_central_compact_w$n.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_h_e2e*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "e2e %.4g"%d["e2e"]["value"], "lite %.4g"%d["e2e"]["obs_stays_on_device"]["value"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
