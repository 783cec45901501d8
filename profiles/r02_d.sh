set -x
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -30
for w in "mobile-synthetic-central-v0 16384" "mobile-synthetic-ma-v0 8192" "mobile-large-ma-v0 131072" "mobile-large-central-v0 262144"; do set -- $w; timeout 300 python bench.py --workload $1 --envs $2 --steps 256 --no-cpu-baseline > gpurun_out/r02_d_bench_$1_$2.json 2>gpurun_out/r02_d_err.txt || tail -3 gpurun_out/r02_d_err.txt; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_d_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1', d['ms_per_step']*1e3, d['roofline']['frac'])"; done
