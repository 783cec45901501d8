# round 2, call n: every kernel family once (tests/sanitize_smoke.py, plain run: compute-sanitizer is closed on this pool);
# full ncu capture of the block-per-env kernel after the row specialisation
set -x
timeout 900 python tests/sanitize_smoke.py 2>&1 | tail -3
B="--steps 24 --warmup 8 --no-graph --no-cpu-baseline --preheat-seconds 0 --repeats 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_big --launch-skip 6 --launch-count 1 -o gpurun_out/r02_n_big python bench.py --workload mobile-synthetic-central-v0 --envs 4096 $B > /dev/null 2>&1
ls -la gpurun_out/r02_n_big.ncu-rep
