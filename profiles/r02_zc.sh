# round 2, call zc: tiled block-per-env kernel (2 UEs per thread): 4 vs 5 CTAs per SM, 8- vs 4-row tiles
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384" 512 > gpurun_out/r02_zc_variants.txt 2>&1
cat gpurun_out/r02_zc_variants.txt
