# round 2, call o: block-per-env kernel -- sparse connectivity test for the central handler; 3 vs 4 CTAs per SM
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "wide or synthetic or big or pf or errors" 2>&1 | tail -3
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384,mobile-synthetic-ma-v0:8192" 256 > gpurun_out/r02_o_variants.txt 2>&1
cat gpurun_out/r02_o_variants.txt
