# round 2, call zd: medium kernel -- per-UE outputs (rate, utility, pos, reward) staged in shared memory and bulk-stored
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "medium or specialised or gymref or window or state or host or full_size_gym" 2>&1 | tail -3
for i in 1 2; do
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072" 1024 >> gpurun_out/r02_zd_variants.txt 2>&1
done
cat gpurun_out/r02_zd_variants.txt
