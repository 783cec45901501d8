# round 2, call zf: medium kernel, staggered start of the first wave
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536" 1024 > gpurun_out/r02_zf_variants.txt 2>&1
cat gpurun_out/r02_zf_variants.txt
