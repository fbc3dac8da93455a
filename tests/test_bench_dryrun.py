"""CPU dry run of bench.py's own arm: the GPU objects are replaced by stand-ins so that the whole
control flow of run_ours() (warm-up, timed legs, roofline arithmetic, the JSON contract) executes
without a device.  Numbers are meaningless here -- the point is that the default command the driver
runs cannot trip over a typo."""
import argparse
import collections
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cobs_b200  # noqa: E402


class FakeEvent:
    def __init__(self, *a, **kw):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 20.0

    def synchronize(self):
        pass


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **kw):
        pass

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


Info = collections.namedtuple("Info", "bytes_per_kmer hbm_bytes")


class FakeIndex:
    def __init__(self, n_docs, h):
        self.info = Info(h * ((n_docs + 7) // 8), 123)
        self.calls = 0
        self.tickets = 0

    def set_option(self, name, value):
        pass

    def timers(self, reset=False):
        return {"hashes_ms": 1.0, "score_ms": 10.0, "select_ms": 0.1, "h2d_ms": 0.1, "d2h_ms": 0.0,
                "kernel_launches": 60, "score_launches": 20, "kmers": 1, "queries": 1}

    def search_device(self, dq, off, thr, k, rpq, counts, keys, stream):
        self.calls += 1

    def _empty(self, off):
        nq = len(off) - 1
        return np.zeros(nq + 1, dtype=np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32)

    def search_packed(self, blob, off, thr, k, raw=False):
        return self._empty(off)

    def submit(self, blob, off, thr, k):
        self.tickets += 1
        return (self.tickets, len(off) - 1, blob, off)

    def collect(self, ticket, raw=False):
        return self._empty(ticket[3])

    def close(self):
        pass


def test_run_ours_control_flow(monkeypatch, capfd):
    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a, **kw: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **kw: FakeStream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    made = []

    def fake_procedural(kind, n_docs, sig, h, **kw):
        assert kw["shard_index"] == 0 and kw["shard_count"] == 1
        made.append(FakeIndex(n_docs, h))
        return made[-1]

    monkeypatch.setattr(cobs_b200.GpuIndex, "procedural", staticmethod(fake_procedural))
    monkeypatch.setattr(bench, "_REAL_STDOUT", None)
    for env in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(env, raising=False)
    args = argparse.Namespace(gpus=1, steps=4, warmup=3, impl="ours", workload="cfg2", nq=50,
                              rows=1000, results_per_query=64, cpu_seconds=1.0, ref_queries=10,
                              no_cpu_baseline=True, no_overlap=False, emulate_shards=0,
                              parallelism="docs", doc_shards=2, no_secondary=True, no_patterns=False,
                              no_parity=True, no_load=True, parity_queries=8, load_gb=1.0)
    bench.run_ours(args)
    out = capfd.readouterr().out
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "roofline", "e2e", "gpu_launches", "clocks", "legs", "parity_check", "patterns"):
        assert key in d, key
    assert d["metric"] == "query_kmers_per_s" and d["n_gpus"] == 1 and d["steps"] == 4
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    assert set(d["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 50 * 100 + 51 * 8
    assert "workload" in d["config"] and d["vs_baseline"] is None
    # value = k-mers of the timed steps / the (fake) 20 ms of CUDA-event time
    assert d["value"] == pytest.approx(50 * 70 * 4 / 0.020)
    assert set(d["legs"]) == {"hits", "default_threshold"}
    assert set(d["patterns"]) >= {"threshold0_limit10", "kmers1000_threshold0_limit10",
                                  "benchmark_fpr_default", "single_query_latency"}
    # the e2e leg goes through the asynchronous submit/collect pair of the public API
    assert made[0].tickets == 8 + 4          # warm-up (two turns of the slot ring) + timed


def test_auto_shard_policy():
    cfg2 = bench.index_bytes_of(bench.WORKLOADS["cfg2"], bench.WORKLOADS["cfg2"]["sig"])
    cfg5 = bench.index_bytes_of(bench.WORKLOADS["cfg5"], bench.WORKLOADS["cfg5"]["sig"])
    assert 100e9 < cfg2 < 110e9 and 550e9 < cfg5 < 650e9
    for world in (1, 2, 4, 8):
        assert bench.auto_doc_shards(cfg2, world) == 1        # fits one GPU: replicas
    assert bench.auto_doc_shards(cfg5, 8) == 8 and bench.auto_doc_shards(cfg5, 4) == 4   # 150 GB > budget
    assert bench.auto_doc_shards(200e9, 8) == 2
