#!/bin/bash
mkdir -p gpurun_out
for opt in "--emulate-shards 8" "--emulate-shards 8 --no-overlap" "--emulate-shards 4" "--emulate-shards 8 --nq 40000"; do
timeout -k 10 600 python bench.py --steps 20 --no-cpu-baseline $opt > gpurun_out/benchq.json 2> gpurun_out/benchq.err
grep '^{' gpurun_out/benchq.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$opt: ms/step %.3f K2 %.3f ms achieved %.0f GB/s phases %s'%(d['ms_per_step'],r['kernel_ms'],r['achieved'],d['phases_ms_per_step_rank0']))"
tail -n 2 gpurun_out/benchq.err
done
