# round 2, call g: end-to-end leg, raw vs compact wire with the rewritten host expansion; host microbench
set -x
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA" | head -8
nvcc -O3 -std=c++17 -o /tmp/hxb profiles/host_expand_bench.cu -lpthread 2>/dev/null
for t in 1 4 8 16; do /tmp/hxb $t 0 | tail -2; done
/tmp/hxb 16 1 | tail -2
timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "wire or host" 2>&1 | tail -3
for wire in compact raw; do
MBE_HOST_WIRE=$wire timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r02_g_e2e_central_$wire.json 2>/dev/null
MBE_HOST_WIRE=$wire timeout 300 python bench.py --no-cpu-baseline --steps 200 --workload mobile-medium-ma-v0 --envs 131072 > gpurun_out/r02_g_e2e_ma_$wire.json 2>/dev/null
done
for n in 4 16; do MBE_HOST_WINDOWS=$n timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r02_g_e2e_central_compact_w$n.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_g_e2e*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "e2e %.4g"%d["e2e"]["value"], "lite %.4g"%d["e2e"]["obs_stays_on_device"]["value"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
