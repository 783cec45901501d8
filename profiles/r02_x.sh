# round 2, call x: N=1 and N=2 driver-form lines back to back on ONE box (what the driver's SCALE run does)
set -x
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_x_n1.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_x_n1_gpu1.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_x_n2.json 2>/dev/null
MBE_BENCH_AFFINITY=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_x_n2_noaff.json 2>/dev/null
python - <<'PY'
import json
for n in ("n1","n1_gpu1","n2","n2_noaff"):
    d=json.loads(open(f"gpurun_out/r02_x_{n}.json").read().strip().splitlines()[-1])
    print(n, "%.4g"%d["value"], "%.3f us"%(d["ms_per_step"]*1e3), d["timing"].get("per_rank_block_ms_median"), d["timing"]["block_ms_min"], d["timing"]["block_ms_max"])
PY
