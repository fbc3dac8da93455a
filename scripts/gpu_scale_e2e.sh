#!/bin/bash
mkdir -p gpurun_out
for tag in a b; do
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/sf2_cfg2_x8_$tag.json 2> gpurun_out/sf2_$tag.err
grep '^{' gpurun_out/sf2_cfg2_x8_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value %.4g ms/step %.3f e2e %.4g K2 %.3f ms'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms']))"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/sf2_$tag.err | tail -n 3
done
