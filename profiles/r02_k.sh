# round 2, call k: CTA-shape / store variants of the medium kernel; hybrid short-lived-tail experiment
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536" 1024 > gpurun_out/r02_k_variants.txt 2>&1
cat gpurun_out/r02_k_variants.txt
timeout 600 python profiles/hybrid_tail_experiment.py 2>&1 | tee gpurun_out/r02_k_hybrid_tail.txt | tail -12
