#!/usr/bin/env python3
"""loader experiments: threads x chunk size x (read only / dma only / both) on a file in /dev/shm"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobs_b200

gb = float(os.environ.get("GB", "12"))
rows = int(gb * 1e9 / 12_500)
path = "/dev/shm/diag_loader.cobs_classic"
g = cobs_b200.GpuIndex.procedural(0, 100_000, [rows], 3, fill_seed=5)
g.save(path)
g.close()
size = os.path.getsize(path)
for mode in ("both", "read", "dma"):
    for threads in (8, 16, 24, 32):
        for chunk in (4, 8, 32):
            os.environ["COBSGPU_LOADER_MODE"] = mode
            os.environ["COBSGPU_LOADER_THREADS"] = str(threads)
            os.environ["COBSGPU_LOADER_CHUNK_MB"] = str(chunk)
            t0 = time.perf_counter()
            g = cobs_b200.GpuIndex.open_file(path)
            dt = time.perf_counter() - t0
            print("mode %-4s threads %2d chunk %2d MB: open %.3f s  stream %.3f s = %.1f GB/s" % (
                mode, threads, chunk, dt, g.info.load_seconds, size / g.info.load_seconds / 1e9), flush=True)
            g.close()
os.unlink(path)
