#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_construct.py -m gpu -q --timeout 600 --timeout-method=thread -x > gpurun_out/pytestq.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytestq.log
N=${1:-2}
for wl in cfg2; do
 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --workload $wl > gpurun_out/benchq_$wl.json 2> gpurun_out/benchq.err
grep '^{' gpurun_out/benchq_$wl.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f launches %d d2h %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],d['gpu_launches'],d['e2e']['d2h_bytes_per_step']))"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/benchq.err | tail -n 5
done
