# round 2: N = 1 and N = 8 driver-form lines back to back on ONE 8-GPU box (what the driver's SCALE run does), final tree
set -x
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale8_n1.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale8_n8.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --workload mobile-synthetic-central-v0 --envs 16384 --steps 256 > gpurun_out/r02_scale8_synthetic_n8.json 2>/dev/null
python - <<'PY'
import json
for n in ("n1","n8","synthetic_n8"):
    d=json.loads(open(f"gpurun_out/r02_scale8_{n}.json").read().strip().splitlines()[-1])
    print(n, "%.4g"%d["value"], "%.3f us"%(d["ms_per_step"]*1e3), "frac %.3f"%d["roofline"]["frac"], d["timing"].get("per_rank_block_ms_median"), "e2e %.4g"%d["e2e"]["value"], d["clocks"])
PY
