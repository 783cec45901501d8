# round 2: multi-GPU lines.  usage: bash profiles/r02_scale.sh N   (under gpurun --gpus N)
# configs[1] weak (65,536 envs per GPU), configs[2] medium-ma (131,072 envs per GPU: 1M over 8 GPUs),
# configs[3] large-central strong scaling (262,144 envs over N GPUs)
N=$1
set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ "$2" = "nodefault" ] || timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_scale_default_driver_n$N.json 2>gpurun_out/r02_scale_err.txt || tail -5 gpurun_out/r02_scale_err.txt
timeout 600 $T bench.py --gpus $N --workload mobile-medium-ma-v0 --envs 131072 --steps 512 > gpurun_out/r02_scale_medium-ma_n$N.json 2>gpurun_out/r02_scale_err.txt || tail -5 gpurun_out/r02_scale_err.txt
timeout 600 $T bench.py --gpus $N --workload mobile-large-central-v0 --total-envs 262144 --steps 512 > gpurun_out/r02_scale_large-central_strong_n$N.json 2>gpurun_out/r02_scale_err.txt || tail -5 gpurun_out/r02_scale_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "value %.4g"%d["value"], "%.3f us"%(d["ms_per_step"]*1e3), "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["scaling"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
