# round 2, call p: block-per-env kernel, sparse connectivity test A/B in one run (twice, interleaved)
set -x
for i in 1 2; do
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384" 512 >> gpurun_out/r02_p_variants.txt 2>&1
done
cat gpurun_out/r02_p_variants.txt
