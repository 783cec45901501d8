# round 2, call s: full-size parity with bit-exact FP64 rates (compiled restatement fed the pinned numpy tables)
set -x
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "full_size or many_envs or compiled" 2>&1 | tail -8 | tee gpurun_out/r02_s_pytest.txt
