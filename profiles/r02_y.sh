# round 2, call y: is the +1.5 % per rank at N = 2 the peer mapping NCCL sets up?  (same box)
set -x
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_y_n1.json 2>/dev/null
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_y_n2.json 2>/dev/null
NCCL_P2P_DISABLE=1 timeout 600 $T --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_y_n2_nop2p.json 2>/dev/null
MBE_BENCH_GATE_CYCLES=2000000 timeout 600 $T --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_y_n2_gate1ms.json 2>/dev/null
timeout 600 $T --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02_y_n2_k200.json 2>/dev/null
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_y_n1_k200.json 2>/dev/null
python - <<'PY'
import json
for n in ("n1","n2","n2_nop2p","n2_gate1ms","n1_k200","n2_k200"):
    d=json.loads(open(f"gpurun_out/r02_y_{n}.json").read().strip().splitlines()[-1])
    print(n, "%.4g"%d["value"], "%.3f us"%(d["ms_per_step"]*1e3), d["timing"].get("per_rank_block_ms_median"))
PY
