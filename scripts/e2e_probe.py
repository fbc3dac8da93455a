# probe: where does the multi-GPU end-to-end time go?  (torchrun, N ranks)
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobs_b200
from cobs_b200.dist import ShardedSearch
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
cfg = bench.WORKLOADS["cfg2"]; nq = cfg["nq"]
ix = cobs_b200.GpuIndex.procedural(cfg["kind"], cfg["n_docs"], cfg["sig"], cfg["h"], fill_seed=1, device=lr, shard_index=rank, shard_count=world)
ix.set_option("max_batch", nq)
steps = 20
batches = [bench.make_batch(10 + i, nq) for i in range(steps + 3)]
pinned = [torch.from_numpy(b).pin_memory() for b, _ in batches]; off = batches[0][1]
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def run(name, overlap, prefetch, mode):
    s = ShardedSearch(ix, rank, world, 64, overlap=overlap)
    ix.set_option("prefetch", int(prefetch)); ix.set_option("inputs_ready", 0)
    for i in range(3): s.search_host(pinned[i], off, 0.8, 0)
    barrier(); t0 = time.perf_counter(); tsub = tcol = 0.0
    if mode == "sync":
        for i in range(steps): s.search_host(pinned[3 + i], off, 0.8, 0)
    else:
        pend = None
        for i in range(steps):
            a = time.perf_counter(); t = s.submit_host(pinned[3 + i], off, 0.8, 0); b = time.perf_counter()
            if pend is not None: s.collect(pend)
            c = time.perf_counter(); tsub += b - a; tcol += c - b; pend = t
        s.collect(pend)
    barrier(); dt = time.perf_counter() - t0
    if rank == 0: print("%-34s %.3f ms/step  (submit %.3f collect %.3f)" % (name, 1e3 * dt / steps, 1e3 * tsub / steps, 1e3 * tcol / steps), flush=True)
run("sync no-overlap no-prefetch", False, False, "sync")
run("sync overlap prefetch", True, True, "sync")
run("pipelined no-overlap no-prefetch", False, False, "pipe")
run("pipelined overlap no-prefetch", True, False, "pipe")
run("pipelined overlap prefetch", True, True, "pipe")
run("pipelined no-overlap prefetch", False, True, "pipe")
if world > 1: dist.destroy_process_group()
