# round 2, call r: block-per-env kernel -- fewer resident CTAs (2 per SM) and interleaved observation rows
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384" 512 >> gpurun_out/r02_r_variants.txt 2>&1
cat gpurun_out/r02_r_variants.txt
