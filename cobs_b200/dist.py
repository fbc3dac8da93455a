"""Document-axis sharding over several GPUs: one process per GPU (torch.distributed, NCCL over
NVLink), every rank holds a column shard of the signature matrix, sees every query, and the
only exchange is one all-gather of the fixed-size per-rank result blocks
[nq][results_per_query] followed by a per-query merge.

The reference has no distributed mode; its closest analogue is the multi-index merge of
counts_to_result (cobs/query/classic_search.cpp:158-201), whose ordering this reproduces.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api

OVERFLOW = 0xFFFFFFFF


class ShardedSearch:
    """Batch search over a document-sharded index.

    index: the GpuIndex holding this rank's shard (shard_index == rank, shard_count == world).
    All ranks must call search_device() with the same queries; every rank ends up with the
    same merged result."""

    def __init__(self, index, rank=0, world=1, results_per_query=64, group=None, overlap=False,
                 depth=4):
        """overlap=True (world > 1): the all-gather + merge of a call run on a side stream and
        rotate through `depth` buffer sets, so they overlap the next call's score kernel.  The
        returned tensors are then only valid after join()."""
        self.index = index
        self.rank = rank
        self.world = world
        self.rpq = results_per_query
        self.group = group
        self.overlap = bool(overlap) and world > 1
        self.depth = max(2, int(depth))      # result-buffer sets in rotation: depth - 1 batches
        self._bufs = {}                      # may be in flight (submit_host) at any time
        self._parity = 0
        self._comm_stream = None
        self._comm_done = [None] * self.depth
        # The main stream never runs more than two batches ahead of the exchange stream, whatever
        # the ring depth: measured on 8 GPUs, letting the score kernels run three ahead of their
        # all-gathers cost 17 % (0.71 vs 0.60 ms per step) -- the exchange only gets SMs in the
        # gaps between the persistent score kernels.
        self._recent_done = []
        self._copy_stream = None
        self._upload_stream = None
        self._pinned = {}
        self._pin_next = 0

    def _set_input_stream(self, handle):
        self.index.set_option("input_stream", handle)

    # -- hooks (overridden by the CPU/gloo protocol test) -------------------------------
    def _local_search(self, d_queries, off, threshold, num_results, counts, keys):
        stream = torch.cuda.current_stream().cuda_stream
        self.index.search_device(d_queries.data_ptr(), off, threshold, num_results, self.rpq,
                                 counts.data_ptr(), keys.data_ptr(), stream)

    def _merge(self, all_counts, all_keys, num_results, out_counts, out_keys):
        stream = torch.cuda.current_stream().cuda_stream
        nq = all_counts.shape[1]
        api.merge_device(all_counts.device.index or 0, self.world, nq, self.rpq,
                         all_counts.data_ptr(), all_keys.data_ptr(), num_results,
                         out_keys.shape[1], out_counts.data_ptr(), out_keys.data_ptr(), stream,
                         counts_list_stride=all_counts.stride(0),
                         keys_list_stride=all_keys.stride(0))

    # ------------------------------------------------------------------------------------
    def out_per_query(self, num_results):
        cap = self.rpq * self.world
        return cap if num_results == 0 else min(num_results, cap)

    def _buffers(self, nq, k, dev, parity=0):
        """Per (nq, k) work buffers, allocated once.  A rank's result block is ONE int64
        tensor [counts (nq x int32, padded) | keys (nq x rpq)], so a single all-gather moves
        everything."""
        key = (nq, k, str(dev), parity)
        b = self._bufs.get(key)
        if b is None:
            # every buffer set of the ring at once: an allocation stalls the device, and a set
            # that is first needed in the middle of a stream of batches would stall it there
            for par in range(self.depth):
                self._bufs[(nq, k, str(dev), par)] = self._make_buffers(nq, k, dev)
            b = self._bufs[key]
        return b

    def _make_buffers(self, nq, k, dev):
        h = (nq + 1) // 2                      # int64 words holding the int32 counts
        L = h + nq * self.rpq
        block = torch.zeros(L, dtype=torch.int64, device=dev)
        gathered = torch.zeros(self.world * L, dtype=torch.int64, device=dev)
        g2 = gathered.view(self.world, L)
        return {
            "block": block,
            "counts": block[:h].view(torch.int32)[:nq],
            "keys": block[h:].view(nq, self.rpq),
            "gathered": gathered,
            "all_counts": g2[:, :h].view(torch.int32)[:, :nq],
            "all_keys": g2[:, h:].view(self.world, nq, self.rpq),
            "out_counts": torch.zeros(nq, dtype=torch.int32, device=dev),
            "out_keys": torch.zeros((nq, k), dtype=torch.int64, device=dev),
        }

    def search_device(self, d_queries, off, threshold, num_results):
        """d_queries: uint8 tensor on this rank's device holding the packed batch; off: host
        uint64[nq+1].  Returns (counts int32[nq], keys int64[nq, out_per_query]) on the device
        (buffers owned by this object, overwritten by a later call with the same shape);
        key = (~score << 32) | global_doc, ascending == (score desc, doc asc); a count of
        0xFFFFFFFF (as uint32) flags a query whose candidates overflowed on some rank."""
        nq = len(off) - 1
        k = self.out_per_query(num_results)
        if not self.overlap:
            self._parity = (self._parity + 1) % self.depth
            b = self._buffers(nq, k, d_queries.device, self._parity)
            self._local_search(d_queries, off, threshold, num_results, b["counts"], b["keys"])
            if self.world == 1:
                return b["counts"], b["keys"][:, :k]
            # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
            dist.all_gather_into_tensor(b["gathered"], b["block"], group=self.group)
            self._merge(b["all_counts"], b["all_keys"], num_results, b["out_counts"],
                        b["out_keys"])
            return b["out_counts"], b["out_keys"]
        # pipelined: local search on the current stream, exchange + merge on a side stream
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=d_queries.device, priority=-1)
        self._parity = (self._parity + 1) % self.depth
        par = self._parity
        b = self._buffers(nq, k, d_queries.device, par)
        main = torch.cuda.current_stream()
        self._throttle(main)
        self._local_search(d_queries, off, threshold, num_results, b["counts"], b["keys"])
        ready, done = self._events(b)
        ready.record(main)
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ready)
            dist.all_gather_into_tensor(b["gathered"], b["block"], group=self.group)
            self._merge(b["all_counts"], b["all_keys"], num_results, b["out_counts"],
                        b["out_keys"])
            done.record(self._comm_stream)
        self._comm_done[par] = done
        self._recent_done.append(done)
        return b["out_counts"], b["out_keys"]

    def _throttle(self, main):
        """main stream waits for the exchange of the batch before the previous one (which also
        implies that the buffer set about to be reused, `depth` >= 2 batches old, is free)"""
        if len(self._recent_done) >= 2:
            main.wait_event(self._recent_done[-2])
            del self._recent_done[:-2]

    @staticmethod
    def _events(b):
        """the (ready, done) event pair of a buffer set, created once and re-recorded"""
        if "ev" not in b:
            b["ev"] = (torch.cuda.Event(), torch.cuda.Event())
        return b["ev"]

    def join(self):
        """make the current stream wait for the pending exchange/merge work"""
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)

    def search_host(self, h_queries, off, threshold, num_results):
        """end-to-end variant: h_queries is a (pinned) host uint8 tensor; returns numpy
        (counts uint32[nq], keys uint64[nq, kmax]) on the host, kmax = longest result list."""
        return self.collect(self.submit_host(h_queries, off, threshold, num_results))

    # -- streaming: keep a few batches in flight -------------------------------------------
    def submit_host(self, h_queries, off, threshold, num_results):
        """Enqueue one batch end to end: H2D of the (pinned) queries, search, exchange/merge,
        D2H of the per-query counts.  Returns a ticket for collect(); up to depth - 1 tickets
        may be outstanding (the result buffers rotate through `depth` sets)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        # the upload runs on its own stream: on the compute stream it (and with it the hash kernel
        # of this batch) would queue behind the score kernel of the previous batch
        if self._upload_stream is None:
            self._upload_stream = torch.cuda.Stream(device=dev, priority=-1)
        with torch.cuda.stream(self._upload_stream):
            d_q = h_queries.to(dev, non_blocking=True)
        # the compute stream orders itself behind the upload too (it is what runs K1 when the
        # handle's "prefetch" option is off); with prefetch on, K1 waits for the upload stream only
        torch.cuda.current_stream().wait_stream(self._upload_stream)
        self._set_input_stream(self._upload_stream.cuda_stream)
        try:
            counts, keys = self.search_device(d_q, off, threshold, num_results)
        finally:
            self._set_input_stream(-1)      # direct search_device() calls: inputs on the compute stream
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev, priority=-1)
        ready = torch.cuda.Event()
        ready.record(self._comm_stream if (self.overlap and self._comm_stream is not None)
                     else torch.cuda.current_stream())
        # Pinned landing buffers are recycled round-robin (allocation is slow); 4 > tickets in
        # flight.  The counts and the first KEYS_AHEAD keys of every query travel together,
        # asynchronously: lists are short, so collect() rarely has to go back for more -- a
        # blocking, pageable copy of the keys at collect() time cost 0.4 ms per batch.
        k0 = min(int(keys.shape[-1]), self.KEYS_AHEAD)
        shape = (tuple(counts.shape), tuple(keys.shape[:-1]) + (k0,))
        pool = self._pinned.get(shape)
        if pool is None:     # all four at once, on first use (pinned allocation is slow)
            pool = [(torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True),
                     torch.empty(shape[1], dtype=keys.dtype, pin_memory=True)) for _ in range(4)]
            self._pinned[shape] = pool
        self._pin_next = (self._pin_next + 1) % 4
        h_counts, h_keys = pool[self._pin_next]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            h_counts.copy_(counts, non_blocking=True)
            h_keys.copy_(keys[..., :k0], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return {"counts": counts, "keys": keys, "h_counts": h_counts, "h_keys": h_keys, "k0": k0,
                "done": done, "d_q": d_q, "nq": len(off) - 1}

    KEYS_AHEAD = 64

    def collect(self, ticket):
        """wait for a submitted batch; returns numpy (counts uint32[nq], keys uint64[nq, kmax])"""
        ticket["done"].synchronize()
        nq = ticket["nq"]
        # the pinned buffers are recycled: copy; [world, per] layouts flatten to query order
        c = ticket["h_counts"].numpy().view(np.uint32).reshape(-1)[:nq].copy()
        valid = c[c < 0xFFFFFFFE]
        kmax = int(valid.max()) if valid.size else 0
        kmax = min(kmax, ticket["keys"].shape[-1])
        if kmax == 0:
            return c, np.zeros((nq, 0), dtype=np.uint64)
        if kmax <= ticket["k0"]:
            k0 = ticket["k0"]
            k = ticket["h_keys"].numpy().view(np.uint64).reshape(-1, k0)[:nq, :kmax].copy()
            return c, k
        with torch.cuda.stream(self._copy_stream):      # a longer list than travelled ahead
            k = ticket["keys"][..., :kmax].contiguous().cpu()
        return c, k.numpy().view(np.uint64).reshape(-1, kmax)[:nq]


class QuerySplitSearch(ShardedSearch):
    """The other way to use several GPUs, for an index that fits into one GPU's HBM: every rank
    holds the WHOLE index and searches its contiguous slice of each query batch; the per-rank
    result blocks are all-gathered, which already is the batch in query order -- no merge.
    (BASELINE.json's north_star shards the document axis "where the matrix overflows one GPU";
    below that size replicas keep the long row slices that the score kernel likes best.)"""

    def search_device(self, d_queries, off, threshold, num_results):
        nq = len(off) - 1
        k = self.out_per_query(num_results)
        if self.world == 1:
            return super().search_device(d_queries, off, threshold, num_results)
        per = (nq + self.world - 1) // self.world
        lo = min(self.rank * per, nq)
        hi = min(lo + per, nq)
        self._parity = (self._parity + 1) % self.depth
        par = self._parity
        b = self._buffers(per, k, d_queries.device, par)
        main = torch.cuda.current_stream() if self.overlap else None
        if self.overlap:
            self._throttle(main)
        if hi < lo + per:
            b["counts"].zero_()                       # ragged last rank: unused slots stay empty
        if hi > lo:
            n = hi - lo
            self._local_search(d_queries, off[lo:hi + 1], threshold, num_results,
                               b["counts"][:n], b["keys"][:n])
        if not self.overlap:
            dist.all_gather_into_tensor(b["gathered"], b["block"], group=self.group)
        else:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=d_queries.device, priority=-1)
            ready, done = self._events(b)
            ready.record(main)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ready)
                dist.all_gather_into_tensor(b["gathered"], b["block"], group=self.group)
                done.record(self._comm_stream)
            self._comm_done[par] = done
            self._recent_done.append(done)
        # rank-major == query order: query q lives at [q // per][q % per]; collect() flattens
        return b["all_counts"], b["all_keys"]

    def out_per_query(self, num_results):
        return self.rpq if num_results == 0 else min(num_results, self.rpq)


class GridSearch:
    """Documents AND queries sharded: world = doc_shards x query_groups.  Rank r belongs to query
    group r // doc_shards and holds document shard r % doc_shards; every query group searches its
    contiguous slice of each batch over its own complete copy of the index (spread over
    doc_shards GPUs) and exchanges only inside the group.  With doc_shards == world this is
    ShardedSearch, with doc_shards == 1 QuerySplitSearch without the final gather.  It keeps the
    row slices per GPU longer than a pure document split when the index is small enough to be
    held query_groups times."""

    def __init__(self, index_factory, rank, world, doc_shards, results_per_query=64, overlap=False,
                 depth=4):
        """index_factory(shard_index, shard_count) -> GpuIndex for this rank's document shard"""
        if world % doc_shards != 0:
            raise ValueError("world size must be a multiple of doc_shards")
        self.rank, self.world, self.doc_shards = rank, world, doc_shards
        self.query_groups = world // doc_shards
        self.group_index = rank // doc_shards
        self.shard_index = rank % doc_shards
        # every rank has to take part in the creation of every subgroup
        self.group = None
        for g in range(self.query_groups):
            ranks = list(range(g * doc_shards, (g + 1) * doc_shards))
            pg = dist.new_group(ranks) if (world > 1 and doc_shards > 1) else None
            if g == self.group_index:
                self.group = pg
        self.index = index_factory(self.shard_index, doc_shards)
        self.inner = self._make_inner(results_per_query, overlap, depth)

    def _make_inner(self, rpq, overlap, depth):
        return ShardedSearch(self.index, self.shard_index, self.doc_shards, rpq, group=self.group,
                             overlap=overlap, depth=depth)

    def slice_of(self, nq):
        """[lo, hi) of a batch of nq queries handled by this rank's query group"""
        per = (nq + self.query_groups - 1) // self.query_groups
        lo = min(self.group_index * per, nq)
        return lo, min(lo + per, nq)

    def search_device(self, d_queries, off, threshold, num_results):
        """searches this group's slice of the batch; returns (lo, hi, counts, keys) with the
        merged results of queries [lo, hi) as ShardedSearch.search_device returns them"""
        lo, hi = self.slice_of(len(off) - 1)
        if hi == lo:      # more query groups than queries: nothing to do for this whole group
            return lo, hi, None, None
        counts, keys = self.inner.search_device(d_queries, off[lo:hi + 1], threshold, num_results)
        return lo, hi, counts, keys

    def join(self):
        self.inner.join()

    # streaming end-to-end use, see ShardedSearch.submit_host / collect
    @property
    def depth(self):
        return self.inner.depth

    def submit_host(self, h_queries, off, threshold, num_results):
        lo, hi = self.slice_of(len(off) - 1)
        return self.inner.submit_host(h_queries, off[lo:hi + 1], threshold, num_results)

    def collect(self, ticket):
        return self.inner.collect(ticket)

    def search_host(self, h_queries, off, threshold, num_results):
        return self.collect(self.submit_host(h_queries, off, threshold, num_results))


def shard_bounds_classic(row_size, shard_count):
    """column-byte ranges of a classic index cut at multiples of 128 documents -- must match
    build_layout() in cobs_b200/csrc/cobsgpu.cu"""
    gran = (row_size + 15) // 16
    out = []
    for g in range(shard_count):
        lo, hi = gran * g // shard_count, gran * (g + 1) // shard_count
        out.append((lo * 16, min(hi * 16, row_size)))
    return out
