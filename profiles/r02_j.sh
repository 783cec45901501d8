# round 2, call j: fused episodes for the scenario shapes (parity + speed against stepping)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py tests/test_export.py -m gpu -q --maxfail=10 -k "rollout or fused or collect" 2>&1 | tail -5
timeout 600 python profiles/fork_rollout_bench.py 65536 2>&1 | tee gpurun_out/r02_j_fork_rollout.txt
