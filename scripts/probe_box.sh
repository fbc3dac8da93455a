#!/bin/bash
# what the GPU box offers (host cores, RAM, scratch space, GPU topology) -> gpurun_out/probe.txt
{
echo "== nproc"; nproc
echo "== free -g"; free -g
echo "== df"; df -h /dev/shm /tmp / 2>/dev/null
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
echo "== topo"; nvidia-smi topo -m 2>/dev/null | head -20
echo "== cpu"; grep -m1 "model name" /proc/cpuinfo; grep -c processor /proc/cpuinfo
echo "== ulimit"; ulimit -a | grep -E "locked|open files"
echo "== cgroup mem"; cat /sys/fs/cgroup/memory.max 2>/dev/null
} > gpurun_out/probe.txt 2>&1
