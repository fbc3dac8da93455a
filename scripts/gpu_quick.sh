#!/bin/bash
# quick 1-GPU check: parity subset + cfg2 bench at the 1-GPU and the 8-way-shard geometry
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 --timeout-method=thread -x > gpurun_out/pytestq.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytestq.log
for extra in "" "--nq 2000" ; do
timeout -k 10 600 python bench.py --steps 20 --no-cpu-baseline $extra > gpurun_out/benchq.json 2> gpurun_out/benchq.err
grep '^{' gpurun_out/benchq.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f launches %d phases %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],d['gpu_launches'],d['phases_ms_per_step_rank0']))"
tail -n 3 gpurun_out/benchq.err
done
