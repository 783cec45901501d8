# round 2, call v: step tickets vs the previous library on the same box
set -x
MBE_CHAIN=0 timeout 300 python bench.py --no-cpu-baseline --steps 1024 > gpurun_out/r02_v_chain0.json 2>/dev/null
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536" 1024 > gpurun_out/r02_v_variants.txt 2>&1
cat gpurun_out/r02_v_variants.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v_chain0.json").read().strip().splitlines()[-1])
print("MBE_CHAIN=0 libmbe.so", "%.3f us"%(d["ms_per_step"]*1e3), "two %.2f"%(d["two_env_groups_in_flight"]["ms_per_step"]*1e3))
PY
