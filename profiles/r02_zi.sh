# round 2, call zi: medium kernel, register budget 6 vs 7 and K = 3 (five UEs per thread) re-measured with the round's final code
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536" 1024 > gpurun_out/r02_zi_variants.txt 2>&1
cat gpurun_out/r02_zi_variants.txt
