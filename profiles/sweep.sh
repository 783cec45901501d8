set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01_u_bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r01_u_bench_reference_arm.json 2> gpurun_out/bench_ref.err
for w in "mobile-medium-ma-v0 131072" "mobile-large-central-v0 262144" "mobile-large-ma-v0 131072" "mobile-small-central-v0 65536" "mobile-synthetic-central-v0 16384" "mobile-custom-v0 262144" "mobile-medium-central-v0 262144"; do set -- $w; timeout 300 python bench.py --workload $1 --envs $2 --steps 1024 --no-cpu-baseline > gpurun_out/r01_u_bench_$1_$2.json 2>/dev/null; done
timeout 300 python profiles/rollout_profile.py
ls -la gpurun_out
