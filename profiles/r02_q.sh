# round 2, call q: block-per-env kernel with 2 UEs per thread (64 registers, 4 CTAs per SM) against 4 UEs per thread (80 / 3)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "wide or synthetic or big or pf or errors" 2>&1 | tail -3
for i in 1 2; do
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384,mobile-synthetic-ma-v0:8192" 512 >> gpurun_out/r02_q_variants.txt 2>&1
done
cat gpurun_out/r02_q_variants.txt
