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

    def __init__(self, index, rank=0, world=1, results_per_query=64, group=None):
        self.index = index
        self.rank = rank
        self.world = world
        self.rpq = results_per_query
        self.group = group

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
                         out_keys.shape[1], out_counts.data_ptr(), out_keys.data_ptr(), stream)

    # ------------------------------------------------------------------------------------
    def out_per_query(self, num_results):
        cap = self.rpq * self.world
        return cap if num_results == 0 else min(num_results, cap)

    def search_device(self, d_queries, off, threshold, num_results):
        """d_queries: uint8 tensor on this rank's device holding the packed batch; off: host
        uint64[nq+1].  Returns (counts int32[nq], keys int64[nq, out_per_query]) on the device;
        key = (~score << 32) | global_doc, ascending == (score desc, doc asc); a count of
        0xFFFFFFFF (as uint32) flags a query whose candidates overflowed on some rank."""
        nq = len(off) - 1
        dev = d_queries.device
        counts = torch.empty(nq, dtype=torch.int32, device=dev)
        keys = torch.empty((nq, self.rpq), dtype=torch.int64, device=dev)
        self._local_search(d_queries, off, threshold, num_results, counts, keys)
        if self.world == 1:
            k = self.out_per_query(num_results)
            return counts, keys[:, :k]
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
        flat_counts = torch.empty(self.world * nq, dtype=torch.int32, device=dev)
        flat_keys = torch.empty((self.world * nq, self.rpq), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat_counts, counts, group=self.group)
        dist.all_gather_into_tensor(flat_keys, keys, group=self.group)
        all_counts = flat_counts.view(self.world, nq)
        all_keys = flat_keys.view(self.world, nq, self.rpq)
        k = self.out_per_query(num_results)
        out_counts = torch.empty(nq, dtype=torch.int32, device=dev)
        out_keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        self._merge(all_counts, all_keys, num_results, out_counts, out_keys)
        return out_counts, out_keys

    def search_host(self, h_queries, off, threshold, num_results):
        """end-to-end variant: h_queries is a (pinned) host uint8 tensor; returns numpy
        (counts uint32[nq], keys uint64[nq, k]) on the host."""
        dev = torch.device("cuda", torch.cuda.current_device())
        d_q = h_queries.to(dev, non_blocking=True)
        counts, keys = self.search_device(d_q, off, threshold, num_results)
        c = counts.cpu().numpy().view(np.uint32)
        k = keys.cpu().numpy().view(np.uint64)
        return c, k


def shard_bounds_classic(row_size, shard_count):
    """column-byte ranges of a classic index cut at multiples of 128 documents -- must match
    build_layout() in cobs_b200/csrc/cobsgpu.cu"""
    gran = (row_size + 15) // 16
    out = []
    for g in range(shard_count):
        lo, hi = gran * g // shard_count, gran * (g + 1) // shard_count
        out.append((lo * 16, min(hi * 16, row_size)))
    return out
