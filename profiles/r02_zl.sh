# round 2, call zl: medium kernel, shared-memory carve-out of the unified L1
set -x
for c in default 25 40 50 60 75 100; do
if [ $c = default ]; then unset MBE_UPT_CARVEOUT; else export MBE_UPT_CARVEOUT=$c; fi
timeout 300 python bench.py --no-cpu-baseline --steps 1024 > gpurun_out/r02_zl_$c.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_zl_$c.json').read().strip().splitlines()[-1]); print('carveout $c', '%.2f us'%(d['ms_per_step']*1e3), 'two %.2f'%(d['two_env_groups_in_flight']['ms_per_step']*1e3))" | tee -a gpurun_out/r02_zl_carveout.txt
done
