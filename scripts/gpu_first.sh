#!/bin/bash
# first GPU contact: smoke, parity tests, tiny + full bench.  Run under gpurun from the repo root.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" | tee gpurun_out/first.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/first.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/first.log
echo "== pytest" | tee -a gpurun_out/first.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/first.log
tail -25 gpurun_out/pytest.log
echo "== bench tiny" | tee -a gpurun_out/first.log
timeout -k 10 300 python bench.py --workload tiny --steps 5 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
echo "bench tiny rc=$?" | tee -a gpurun_out/first.log
cat gpurun_out/bench_tiny.json; tail -5 gpurun_out/bench_tiny.err
echo "== bench cfg2" | tee -a gpurun_out/first.log
timeout -k 10 600 python bench.py --steps 10 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
echo "bench cfg2 rc=$?" | tee -a gpurun_out/first.log
cat gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
