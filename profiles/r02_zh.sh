# round 2, call zh: end-to-end leg on 8 GPUs of one host -- DMA rows against the compact wire format (4 and 3 host threads per rank)
set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nproc
MBE_HOST_WIRE=compact MBE_HOST_THREADS=4 timeout 600 $T --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_zh_n8_compact_t4.json 2>/dev/null
MBE_HOST_WIRE=compact MBE_HOST_THREADS=3 timeout 600 $T --master-port 29572 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_zh_n8_compact_t3.json 2>/dev/null
timeout 600 $T --master-port 29573 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_zh_n8_raw.json 2>/dev/null
python - <<'PY'
import json
for n in ("compact_t4","compact_t3","raw"):
    d=json.loads(open(f"gpurun_out/r02_zh_n8_{n}.json").read().strip().splitlines()[-1])
    print(n, "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "lite %.4g"%d["e2e"]["obs_stays_on_device"]["value"])
PY
