# round 2, call zj: large shapes -- per-BS counts by one warp-wide integer add (redux) instead of ballot + popc + and
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "large or specialised or gymref or full_size_large" 2>&1 | tail -3
for i in 1 2; do
timeout 900 python profiles/variant_sweep.py run "mobile-large-central-v0:262144,mobile-large-ma-v0:131072" 512 >> gpurun_out/r02_zj_variants.txt 2>&1
done
cat gpurun_out/r02_zj_variants.txt
