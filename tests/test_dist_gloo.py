"""CPU test (world_size 2, gloo) of the multi-GPU host logic in cobs_b200/dist.py: every rank
holds a document-axis shard, all-gathers its fixed-size result block and merges.  The CUDA
hooks (_local_search, _merge) are replaced by CPU stand-ins built on the oracle -- this checks
the protocol: shard bounds, global document ids, rank order, overflow propagation, truncation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

N_DOCS, SIG, H, SEED = 3000, [53], 3, 9
RPQ = 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _key(score, doc):
    return ((~np.uint64(score) & np.uint64(0xFFFFFFFF)) << np.uint64(32)) | np.uint64(doc)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cobs_b200.dist import ShardedSearch, shard_bounds_classic, OVERFLOW
    from oracle import oracle

    o = oracle.Index.procedural(oracle.KIND_CLASSIC, N_DOCS, SIG, H, fill_seed=SEED,
                                materialize=True)
    row = (N_DOCS + 7) // 8
    b0, b1 = shard_bounds_classic(row, world)[rank]
    queries = [oracle.random_query(i, 100) for i in range(12)]

    class CpuShard(ShardedSearch):
        def _local_search(self, d_queries, off, threshold, num_results, counts, keys):
            blob = bytes(d_queries.numpy())
            for i in range(len(off) - 1):
                q = blob[int(off[i]):int(off[i + 1])]
                sc = o.scores(q, 8 * b0, 8 * b1)
                T = len(q) - 31 + 1
                thr = int(np.ceil(threshold * T))
                docs = [d for d in range(8 * b0, min(8 * b1, N_DOCS)) if sc[d - 8 * b0] >= thr]
                docs.sort(key=lambda d: (-int(sc[d - 8 * b0]), d))
                if len(docs) > self.rpq and num_results == 0:
                    counts[i] = np.int32(-1)        # 0xFFFFFFFF: candidates overflowed
                    continue
                if num_results:
                    docs = docs[:num_results]
                docs = docs[:self.rpq]
                counts[i] = len(docs)
                for j, d in enumerate(docs):
                    keys[i, j] = int(np.int64(_key(int(sc[d - 8 * b0]), d).view(np.int64)))

        def _merge(self, all_counts, all_keys, num_results, out_counts, out_keys):
            c = all_counts.numpy().view(np.uint32)
            k = all_keys.numpy().view(np.uint64)
            for i in range(c.shape[1]):
                if (c[:, i] == OVERFLOW).any():
                    out_counts[i] = np.int32(-1)
                    continue
                m = np.sort(np.concatenate([k[r, i, :c[r, i]] for r in range(c.shape[0])]))
                if num_results:
                    m = m[:num_results]
                m = m[:out_keys.shape[1]]
                out_counts[i] = len(m)
                out_keys[i, :len(m)] = torch.from_numpy(m.view(np.int64))

    s = CpuShard(None, rank, world, RPQ)
    blob = torch.frombuffer(bytearray(b"".join(queries)), dtype=torch.uint8)
    off = np.arange(len(queries) + 1, dtype=np.uint64) * 100
    ok = True
    for thr, k in ((0.12, 5), (0.12, 0), (0.05, 0), (0.3, 0)):
        counts, keys = s.search_device(blob, off, thr, k)
        c = counts.numpy().view(np.uint32)
        kk = keys.numpy().view(np.uint64)
        for i, q in enumerate(queries):
            want = oracle.search(o, q, thr, k)
            if c[i] == OVERFLOW:
                # only legal when some shard really had more than RPQ candidates
                ok &= (k == 0 and len(want) > RPQ)
                continue
            got = [(0, int(x & np.uint64(0xFFFFFFFF)),
                    int(~(x >> np.uint64(32)) & np.uint64(0xFFFFFFFF))) for x in kk[i, :c[i]]]
            ok &= got == want[:keys.shape[1]]
    # every rank must hold the same merged result
    c = counts.numpy().view(np.uint32)
    valid = [keys.numpy()[i, :c[i]].tobytes() if c[i] != OVERFLOW else b"ovf"
             for i in range(len(queries))]
    gathered = [None] * world
    dist.all_gather_object(gathered, (counts.numpy().tobytes(), valid))
    ok &= all(g == gathered[0] for g in gathered)
    # ---- the pipelined control flow (overlap=True) with stand-ins for the CUDA stream API:
    # ring of buffer sets, two-batch throttle, side-stream exchange.  Results of several
    # consecutive batches must all be right and must not clobber each other. ----
    import contextlib

    class _FakeEvent:
        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

    class _FakeStream:
        def __init__(self, *a, **kw):
            pass

        def wait_event(self, ev):
            pass

        def wait_stream(self, st):
            pass

    torch.cuda.Event = _FakeEvent
    torch.cuda.Stream = _FakeStream
    torch.cuda.current_stream = lambda *a, **kw: _FakeStream()
    torch.cuda.stream = lambda st: contextlib.nullcontext()
    p = CpuShard(None, rank, world, RPQ, overlap=True, depth=3)
    outs = []
    for step in range(5):
        qs = queries[step:step + 4]
        bl = torch.frombuffer(bytearray(b"".join(qs)), dtype=torch.uint8)
        of = np.arange(len(qs) + 1, dtype=np.uint64) * 100
        c_, k_ = p.search_device(bl, of, 0.12, 5)
        outs.append((qs, c_.numpy().view(np.uint32).copy(), k_.numpy().view(np.uint64).copy()))
        ok &= len(p._recent_done) <= 3      # pruned to the last two at every call
    p.join()
    for qs, c_, k_ in outs:
        for i, q in enumerate(qs):
            got = [(0, int(x & np.uint64(0xFFFFFFFF)),
                    int(~(x >> np.uint64(32)) & np.uint64(0xFFFFFFFF))) for x in k_[i, :c_[i]]]
            ok &= got == oracle.search(o, q, 0.12, 5)

    # ---- replicas: every rank holds all columns and searches its slice of the batch ----
    from cobs_b200.dist import QuerySplitSearch

    class CpuReplica(QuerySplitSearch):
        def _local_search(self, d_queries, off, threshold, num_results, counts, keys):
            blob = bytes(d_queries.numpy())
            for i in range(len(off) - 1):
                q = blob[int(off[i]):int(off[i + 1])]
                res = oracle.search(o, q, threshold, num_results)[:self.rpq]
                counts[i] = len(res)
                for j, (_, d, sc_) in enumerate(res):
                    keys[i, j] = int(np.int64(_key(sc_, d).view(np.int64)))

    r = CpuReplica(None, rank, world, 64)
    odd = queries[:11]                      # 11 queries over 2 ranks: ragged last slice
    blob = torch.frombuffer(bytearray(b"".join(odd)), dtype=torch.uint8)
    off = np.arange(len(odd) + 1, dtype=np.uint64) * 100
    counts, keys = r.search_device(blob, off, 0.12, 7)
    c = counts.numpy().view(np.uint32).reshape(-1)[:len(odd)]
    kk = keys.numpy().view(np.uint64).reshape(-1, keys.shape[-1])[:len(odd)]
    for i, q in enumerate(odd):
        got = [(0, int(x & np.uint64(0xFFFFFFFF)),
                int(~(x >> np.uint64(32)) & np.uint64(0xFFFFFFFF))) for x in kk[i, :c[i]]]
        ok &= got == oracle.search(o, q, 0.12, 7)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_bounds_cover_all_columns():
    sys.path.insert(0, ROOT)
    from cobs_b200.dist import shard_bounds_classic
    for row in (1, 15, 16, 17, 12_500, 125_000):
        for world in (1, 2, 3, 8):
            b = shard_bounds_classic(row, world)
            assert b[0][0] == 0 and b[-1][1] == row
            for (a0, a1), (c0, c1) in zip(b, b[1:]):
                assert a1 == c0 or (a1 == row and c0 >= row)
                assert a0 % 16 == 0 and c0 % 16 == 0


def test_world2_gloo_shard_merge():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _grid_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cobs_b200.dist import GridSearch, ShardedSearch, shard_bounds_classic, OVERFLOW
    from oracle import oracle

    o = oracle.Index.procedural(oracle.KIND_CLASSIC, N_DOCS, SIG, H, fill_seed=SEED,
                                materialize=True)
    row = (N_DOCS + 7) // 8

    class CpuShard(ShardedSearch):
        def _local_search(self, d_queries, off, threshold, num_results, counts, keys):
            b0, b1 = shard_bounds_classic(row, self.world)[self.rank]
            blob = bytes(d_queries.numpy())
            for i in range(len(off) - 1):
                q = blob[int(off[i]):int(off[i + 1])]
                sc = o.scores(q, 8 * b0, 8 * b1)
                thr = int(np.ceil(threshold * (len(q) - 30)))
                docs = [d for d in range(8 * b0, min(8 * b1, N_DOCS)) if sc[d - 8 * b0] >= thr]
                docs.sort(key=lambda d: (-int(sc[d - 8 * b0]), d))
                docs = (docs[:num_results] if num_results else docs)[:self.rpq]
                counts[i] = len(docs)
                for j, d in enumerate(docs):
                    keys[i, j] = int(np.int64(_key(int(sc[d - 8 * b0]), d).view(np.int64)))

        def _merge(self, all_counts, all_keys, num_results, out_counts, out_keys):
            c = all_counts.numpy().view(np.uint32)
            k = all_keys.numpy().view(np.uint64)
            for i in range(c.shape[1]):
                m = np.sort(np.concatenate([k[r, i, :c[r, i]] for r in range(c.shape[0])]))
                m = (m[:num_results] if num_results else m)[:out_keys.shape[1]]
                out_counts[i] = len(m)
                out_keys[i, :len(m)] = torch.from_numpy(m.view(np.int64))

    class CpuGrid(GridSearch):
        def _make_inner(self, rpq, overlap, depth):
            return CpuShard(None, self.shard_index, self.doc_shards, rpq, group=self.group)

    g = CpuGrid(lambda si, sc: None, rank, world, doc_shards=2, results_per_query=64)
    ok = (g.query_groups == 2 and g.group_index == rank // 2 and g.shard_index == rank % 2)
    queries = [oracle.random_query(i, 100) for i in range(9)]       # 9 queries: slices 5 + 4
    blob = torch.frombuffer(bytearray(b"".join(queries)), dtype=torch.uint8)
    off = np.arange(len(queries) + 1, dtype=np.uint64) * 100
    for thr, k in ((0.12, 6), (0.3, 0)):
        lo, hi, counts, keys = g.search_device(blob, off, thr, k)
        ok &= (lo, hi) == ((0, 5) if rank < 2 else (5, 9))
        c = counts.numpy().view(np.uint32)
        kk = keys.numpy().view(np.uint64)
        for i, q in enumerate(queries[lo:hi]):
            got = [(0, int(x & np.uint64(0xFFFFFFFF)),
                    int(~(x >> np.uint64(32)) & np.uint64(0xFFFFFFFF))) for x in kk[i, :c[i]]]
            ok &= got == oracle.search(o, q, thr, k)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_world4_gloo_grid_docs_x_queries():
    """2 document shards x 2 query groups on 4 gloo ranks: subgroup exchange only"""
    world = 4
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grid_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True, 2: True, 3: True}
