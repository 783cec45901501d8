# round 2, call l: change flags instead of kept input words in the medium kernel (fewer live registers): 72- vs 64-register budgets
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "medium or specialised or gymref or full_size or window or state" 2>&1 | tail -3
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072" 1024 > gpurun_out/r02_l_variants.txt 2>&1
cat gpurun_out/r02_l_variants.txt
