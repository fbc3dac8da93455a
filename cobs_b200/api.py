"""Python face of the B200 COBS query path.

`GpuIndex` is a thin object wrapper over the C ABI (include/cobsgpu.h).  `Search` mirrors the
reference's Python class (python/module.cpp:351-386: `cobs_index.Search(path).search(query,
threshold=0.0, num_results=0)` returning objects with `.doc_name` and `.score`), including the
multi-index merge and the ordering rules of cobs::ClassicSearch
(cobs/query/classic_search.cpp:109-202, 403-505).
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _lib
from ._lib import CobsGpuError, KIND_CLASSIC, KIND_COMPACT  # noqa: F401

SearchResult = namedtuple("SearchResult", ["doc_name", "score"])


def _pack_queries(queries):
    """list of bytes/str -> (blob bytes, uint64 offsets[nq+1])"""
    qs = [q if isinstance(q, (bytes, bytearray)) else q.encode("ascii") for q in queries]
    off = np.zeros(len(qs) + 1, dtype=np.uint64)
    if qs:
        off[1:] = np.cumsum([len(q) for q in qs], dtype=np.uint64)
    return b"".join(qs), off


class GpuIndex:
    """One index (or one document-axis shard of it) resident in HBM."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        info = _lib.IndexInfo()
        _lib.check(_lib.lib().cobsgpu_index_get_info(self._h, C.byref(info)))
        self.info = info

    # -- construction --------------------------------------------------------------------
    @classmethod
    def open_file(cls, path, device=0, shard_index=0, shard_count=1):
        h = C.c_void_p()
        p = path if isinstance(path, bytes) else str(path).encode()
        _lib.check(_lib.lib().cobsgpu_index_open_file(p, device, shard_index, shard_count,
                                                      C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, kind, n_docs, pages, num_hashes, term_size=31, canonicalize=1,
                    device=0, shard_index=0, shard_count=1):
        """pages: list of uint8 arrays [signature_size_p, page_size] in the reference's
        row-major, LSB-first layout (classic: one array with page_size = ceil(n_docs/8))."""
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        sig = np.array([p.shape[0] for p in pages], dtype=np.uint64)
        ptrs = (C.c_void_p * len(pages))(*[p.ctypes.data for p in pages])
        d = _lib.IndexDesc()
        d.struct_size = C.sizeof(_lib.IndexDesc)
        d.kind = kind
        d.term_size = term_size
        d.canonicalize = canonicalize
        d.num_hashes = num_hashes
        d.n_docs = n_docs
        d.n_pages = len(pages)
        d.page_size = pages[0].shape[1]
        d.signature_sizes = sig.ctypes.data_as(C.POINTER(C.c_uint64))
        d.page_data = ptrs
        d.device = device
        d.shard_index = shard_index
        d.shard_count = shard_count
        h = C.c_void_p()
        _lib.check(_lib.lib().cobsgpu_index_open(C.byref(d), C.byref(h)))
        return cls(h.value)

    @classmethod
    def procedural(cls, kind, n_docs, signature_sizes, num_hashes, page_size=0, term_size=31,
                   canonicalize=1, fill_seed=1, device=0, shard_index=0, shard_count=1):
        """synthetic index whose bits are generated on the device"""
        sig = np.ascontiguousarray(signature_sizes, dtype=np.uint64)
        d = _lib.IndexDesc()
        d.struct_size = C.sizeof(_lib.IndexDesc)
        d.kind = kind
        d.term_size = term_size
        d.canonicalize = canonicalize
        d.num_hashes = num_hashes
        d.n_docs = n_docs
        d.n_pages = len(sig)
        d.page_size = page_size if kind == KIND_COMPACT else (n_docs + 7) // 8
        d.signature_sizes = sig.ctypes.data_as(C.POINTER(C.c_uint64))
        d.page_data = None
        d.fill_seed = fill_seed
        d.device = device
        d.shard_index = shard_index
        d.shard_count = shard_count
        h = C.c_void_p()
        _lib.check(_lib.lib().cobsgpu_index_open(C.byref(d), C.byref(h)))
        return cls(h.value)

    @classmethod
    def construct_classic(cls, documents, num_hashes=1, false_positive_rate=0.3, term_size=31,
                          canonicalize=1, signature_size=0, device=0):
        """documents: list of (name, [sequence, ...]); builds the classic index on the device
        (the reference's classic_construct, cobs/construction/classic_index.cpp:565-600)"""
        names, seqs, seq_doc = [], [], []
        for i, (name, parts) in enumerate(documents):
            names.append(name.encode() if isinstance(name, str) else name)
            for p in parts:
                seqs.append(p if isinstance(p, (bytes, bytearray)) else p.encode("ascii"))
                seq_doc.append(i)
        blob, off = _pack_queries(seqs)
        sd = np.asarray(seq_doc, dtype=np.uint32)
        d = _lib.ConstructDesc()
        d.struct_size = C.sizeof(_lib.ConstructDesc)
        d.term_size = term_size
        d.canonicalize = canonicalize
        d.num_hashes = num_hashes
        d.signature_size = signature_size
        d.false_positive_rate = false_positive_rate
        d.n_docs = len(names)
        d.n_seqs = len(seqs)
        d.doc_names = (C.c_char_p * len(names))(*names)
        d.sequences = blob
        d.seq_offsets = off.ctypes.data_as(C.POINTER(C.c_uint64))
        d.seq_doc = sd.ctypes.data_as(C.POINTER(C.c_uint32))
        d.device = device
        h = C.c_void_p()
        _lib.check(_lib.lib().cobsgpu_construct_classic(C.byref(d), C.byref(h)))
        return cls(h.value)

    def save(self, path):
        """write the index in the reference's file format"""
        p = path if isinstance(path, bytes) else str(path).encode()
        _lib.check(_lib.lib().cobsgpu_index_save(self._h, p))

    def signature_size(self, page=0):
        return _lib.lib().cobsgpu_index_signature_size(self._h, page)

    def close(self):
        if self._h:
            _lib.lib().cobsgpu_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- geometry (IndexSearchFile getters, cobs/query/index_file.hpp:27-34) ---------------
    @property
    def term_size(self): return self.info.term_size
    @property
    def canonicalize(self): return self.info.canonicalize
    @property
    def num_hashes(self): return self.info.num_hashes
    @property
    def n_docs(self): return self.info.n_docs
    @property
    def counts_size(self): return self.info.counts_size
    @property
    def row_size(self): return self.info.row_size
    @property
    def page_size(self): return self.info.page_size

    def doc_name(self, doc):
        s = _lib.lib().cobsgpu_index_doc_name(self._h, doc)
        return s.decode() if s is not None else "doc_%06d" % doc

    def set_option(self, name, value):
        _lib.check(_lib.lib().cobsgpu_set_option(self._h, name.encode(), int(value)))

    # -- hot path --------------------------------------------------------------------------
    def hash(self, queries):
        """K1: raw XXH64 values, one uint64 array per query ([T*h], k-mer major)"""
        blob, off = _pack_queries(queries)
        k, h = self.term_size, self.num_hashes
        n = [max(int(off[i + 1] - off[i]) - k + 1, 0) * h for i in range(len(queries))]
        out = np.zeros(max(sum(n), 1), dtype=np.uint64)
        _lib.check(_lib.lib().cobsgpu_hash(self._h, blob, off.ctypes.data, len(queries),
                                           out.ctypes.data))
        res, p = [], 0
        for c in n:
            res.append(out[p:p + c])
            p += c
        return res

    def scores(self, queries, out=None):
        """K1+K2: uint32 [nq, counts_size] hit counts (the reference's score_list)"""
        blob, off = _pack_queries(queries)
        if out is None:
            out = np.zeros((len(queries), self.counts_size), dtype=np.uint32)
        _lib.check(_lib.lib().cobsgpu_scores(self._h, blob, off.ctypes.data, len(queries),
                                             out.ctypes.data))
        return out

    def search_batch(self, queries, threshold=0.0, num_results=0):
        """K1+K2+K3 for a batch: list (per query) of (doc uint32[], score uint32[])"""
        blob, off = _pack_queries(queries)
        return self.search_packed(blob, off, threshold, num_results)

    def search_packed(self, blob, off, threshold=0.0, num_results=0, raw=False):
        """same on an already packed batch (blob: bytes/ndarray, off: uint64[nq+1])"""
        nq = len(off) - 1
        r = _lib.Result()
        bp = blob.ctypes.data if isinstance(blob, np.ndarray) else blob
        _lib.check(_lib.lib().cobsgpu_search_batch(self._h, bp, off.ctypes.data, nq,
                                                   threshold, num_results, C.byref(r)))
        return self._unpack(r, nq, raw)

    def submit(self, blob, off, threshold=0.0, num_results=0):
        """asynchronous half of search_packed: enqueue one batch (at most `max_batch` queries),
        returns a ticket for collect().  Up to 4 tickets may be outstanding; `blob` and `off`
        must stay alive until the ticket is collected."""
        t = C.c_uint64()
        bp = blob.ctypes.data if isinstance(blob, np.ndarray) else blob
        _lib.check(_lib.lib().cobsgpu_submit(self._h, bp, off.ctypes.data, len(off) - 1,
                                             threshold, num_results, C.byref(t)))
        return (t.value, len(off) - 1, blob, off)

    def collect(self, ticket, raw=False):
        r = _lib.Result()
        _lib.check(_lib.lib().cobsgpu_collect(self._h, ticket[0], C.byref(r)))
        return self._unpack(r, ticket[1], raw)

    @staticmethod
    def _unpack(r, nq, raw):
        # raw == "view": no copies -- the arrays alias the library's result buffers and are only
        # valid until the handle's next calls (see include/cobsgpu.h); for result volumes where
        # a copy costs more than the search (every document of a large index per query)
        view = raw == "view"
        roff = np.ctypeslib.as_array(r.offsets, shape=(nq + 1,))
        if not view:
            roff = roff.copy()
        total = int(roff[nq])
        if total:
            doc = np.ctypeslib.as_array(r.doc, shape=(total,))
            score = np.ctypeslib.as_array(r.score, shape=(total,))
            if not view:
                doc, score = doc.copy(), score.copy()
        else:
            doc = np.zeros(0, dtype=np.uint32)
            score = np.zeros(0, dtype=np.uint32)
        if raw:
            return roff, doc, score
        return [(doc[int(roff[i]):int(roff[i + 1])], score[int(roff[i]):int(roff[i + 1])])
                for i in range(nq)]

    def search_device(self, d_queries_ptr, off, threshold, num_results, results_per_query,
                      d_counts_ptr, d_keys_ptr, stream=0):
        """device-resident variant: raw device pointers (e.g. torch .data_ptr())"""
        _lib.check(_lib.lib().cobsgpu_search_batch_device(
            self._h, d_queries_ptr, off.ctypes.data, len(off) - 1, threshold, num_results,
            results_per_query, d_counts_ptr, d_keys_ptr, stream))

    def timers(self, reset=False):
        t = _lib.Timers()
        _lib.check(_lib.lib().cobsgpu_get_timers(self._h, C.byref(t)))
        if reset:
            _lib.lib().cobsgpu_reset_timers(self._h)
        return {f: getattr(t, f) for f, _ in t._fields_}

    def read_row(self, page, row, begin, nbytes):
        out = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(_lib.lib().cobsgpu_debug_read_row(self._h, page, row, begin, nbytes,
                                                     out.ctypes.data))
        return out


class GpuGroup:
    """One index sharded along the document axis over several GPUs of THIS process (shard g on
    devices[g]); the leader GPU merges the shards' result blocks over NVLink.  Same search
    interface as GpuIndex."""

    def __init__(self, handle):
        self._g = C.c_void_p(handle)
        n = _lib.lib().cobsgpu_group_size(self._g)
        self.shards = []
        for i in range(n):
            ix = GpuIndex(_lib.lib().cobsgpu_group_shard(self._g, i))
            ix.close = lambda: None          # borrowed: the group owns its shards
            self.shards.append(ix)
        self.info = self.shards[0].info

    @classmethod
    def open_file(cls, path, devices):
        g = C.c_void_p()
        p = path if isinstance(path, bytes) else str(path).encode()
        dev = (C.c_int32 * len(devices))(*devices)
        _lib.check(_lib.lib().cobsgpu_group_open_file(p, dev, len(devices), C.byref(g)))
        return cls(g.value)

    @classmethod
    def procedural(cls, kind, n_docs, signature_sizes, num_hashes, devices, page_size=0,
                   term_size=31, canonicalize=1, fill_seed=1):
        sig = np.ascontiguousarray(signature_sizes, dtype=np.uint64)
        d = _lib.IndexDesc()
        d.struct_size = C.sizeof(_lib.IndexDesc)
        d.kind = kind
        d.term_size = term_size
        d.canonicalize = canonicalize
        d.num_hashes = num_hashes
        d.n_docs = n_docs
        d.n_pages = len(sig)
        d.page_size = page_size if kind == KIND_COMPACT else (n_docs + 7) // 8
        d.signature_sizes = sig.ctypes.data_as(C.POINTER(C.c_uint64))
        d.page_data = None
        d.fill_seed = fill_seed
        g = C.c_void_p()
        dev = (C.c_int32 * len(devices))(*devices)
        _lib.check(_lib.lib().cobsgpu_group_open(C.byref(d), dev, len(devices), C.byref(g)))
        return cls(g.value)

    def set_option(self, name, value):
        for s in self.shards:
            s.set_option(name, value)

    def search_batch(self, queries, threshold=0.0, num_results=0):
        blob, off = _pack_queries(queries)
        return self.search_packed(blob, off, threshold, num_results)

    def search_packed(self, blob, off, threshold=0.0, num_results=0, raw=False):
        nq = len(off) - 1
        r = _lib.Result()
        bp = blob.ctypes.data if isinstance(blob, np.ndarray) else blob
        _lib.check(_lib.lib().cobsgpu_group_search_batch(self._g, bp, off.ctypes.data, nq,
                                                         threshold, num_results, C.byref(r)))
        return GpuIndex._unpack(r, nq, raw)

    def close(self):
        if self._g:
            for s in self.shards:
                s._h = None
            _lib.lib().cobsgpu_group_close(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_device(device, n_lists, nq, results_per_query, d_counts_ptr, d_keys_ptr, num_results,
                 out_per_query, d_out_counts_ptr, d_out_keys_ptr, stream=0,
                 counts_list_stride=0, keys_list_stride=0):
    """strides in elements (uint32 / uint64); 0 = densely packed lists"""
    _lib.check(_lib.lib().cobsgpu_merge_device(device, n_lists, nq, results_per_query,
                                               d_counts_ptr, counts_list_stride, d_keys_ptr,
                                               keys_list_stride, num_results, out_per_query,
                                               d_out_counts_ptr, d_out_keys_ptr, stream))


def decode_keys(keys):
    """uint64 sort keys -> (doc, score)"""
    keys = np.asarray(keys, dtype=np.uint64)
    doc = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    score = (~(keys >> np.uint64(32))).astype(np.uint32)
    return doc, score


class Search:
    """Drop-in for the reference's `cobs_index.Search` / cobs::ClassicSearch over one or more
    index files (python/module.cpp:367-386; multi-index rules classic_search.cpp:158-201)."""

    def __init__(self, paths, device=0):
        if isinstance(paths, (str, bytes)):
            paths = [paths]
        self.indices = [GpuIndex.open_file(p, device) for p in paths]

    def close(self):
        for ix in self.indices:
            ix.close()
        self.indices = []

    def search(self, query, threshold=0.0, num_results=0):
        return self.search_batch([query], threshold, num_results)[0]

    def search_batch(self, queries, threshold=0.0, num_results=0):
        """extension of the reference API: many queries per call (one GPU batch)"""
        nq = len(queries)
        if not self.indices:
            return [[] for _ in range(nq)]
        qlens = [len(q) for q in queries]
        total_docs = sum(ix.counts_size for ix in self.indices)

        def hashes_of(qi):
            return sum(ix.num_hashes * (qlens[qi] - ix.term_size + 1) for ix in self.indices)

        # The reference skips the sort when the query produced a single hash in total
        # (classic_search.cpp:130, max_counts = total_hashes): it then returns the first
        # `limit` kept documents in column order.  Those queries need the untruncated list.
        quirk = [qi for qi in range(nq) if qlens[qi] >= max(ix.term_size for ix in self.indices)
                 and hashes_of(qi) <= 1]
        quirk_set = set(quirk)
        normal = [qi for qi in range(nq) if qi not in quirk_set]
        per_index = [dict() for _ in self.indices]
        for f, ix in enumerate(self.indices):
            # per-index lists, each already ordered (score desc, doc asc) and cut at
            # num_results: the global top-k is contained in the union of per-index top-ks
            for ids, k in ((normal, num_results), (quirk, 0)):
                if ids:
                    res = ix.search_batch([queries[i] for i in ids], threshold, k)
                    per_index[f].update(zip(ids, res))
        out = []
        for qi in range(nq):
            limit = total_docs if num_results == 0 else min(num_results, total_docs)
            ents = []
            for f in range(len(self.indices)):
                doc, score = per_index[f][qi]
                ents.extend((int(s), f, int(d)) for d, s in zip(doc, score))
            if hashes_of(qi) <= 1:
                ents.sort(key=lambda e: (e[1], e[2]))
            elif len(self.indices) > 1:
                # score desc, then (file, doc) asc (classic_search.cpp:173-177)
                ents.sort(key=lambda e: (-e[0], e[1], e[2]))
            ents = ents[:limit]
            out.append([SearchResult(self.indices[f].doc_name(d), s) for s, f, d in ents])
        return out
