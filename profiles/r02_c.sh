set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072" > gpurun_out/r02_c_variants.txt 2>&1
cat gpurun_out/r02_c_variants.txt | tail -16
for w in "mobile-large-central-v0 262144" "mobile-small-central-v0 65536" "mobile-large-ma-v0 131072"; do set -- $w; timeout 300 python bench.py --workload $1 --envs $2 --steps 1024 --no-cpu-baseline > gpurun_out/r02_c_bench_$1_$2.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_c_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1', d['ms_per_step']*1e3, d['roofline']['frac'])"; done
