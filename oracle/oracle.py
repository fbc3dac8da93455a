"""ctypes view of oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/cobs_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK = 0
ERR_INVALID_BASE = -1
ERR_TOO_SHORT = -2
ERR_BAD_FILE = -3
ERR_BAD_PARAM = -4
ERR_IO = -5

KIND_CLASSIC = 0
KIND_COMPACT = 1


class OracleError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("oracle: %s failed with code %d" % (what, code))
        self.code = code


class _Index(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("term_size", C.c_uint32),
        ("canonicalize", C.c_uint8),
        ("num_hashes", C.c_uint64),
        ("n_docs", C.c_uint32),
        ("n_pages", C.c_uint32),
        ("page_size", C.c_uint64),
        ("signature_sizes", C.POINTER(C.c_uint64)),
        ("page_data", C.POINTER(C.POINTER(C.c_uint8))),
        ("doc_names", C.POINTER(C.c_char_p)),
        ("procedural", C.c_int),
        ("fill_seed", C.c_uint64),
        ("file_blob", C.POINTER(C.c_uint8)),
        ("file_size", C.c_size_t),
        ("owns_pages", C.c_int),
    ]


def build():
    """(Re)build liboracle.so with gcc; cheap, used by __graft_entry__.build()."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_xxh64.restype = C.c_uint64
        L.oracle_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.oracle_canonicalize_kmer.restype = C.c_int
        L.oracle_canonicalize_kmer.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.oracle_create_hashes.restype = C.c_int
        L.oracle_create_hashes.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32,
                                           C.c_uint64, C.c_uint8, C.c_void_p]
        L.oracle_counts_size.restype = C.c_uint64
        L.oracle_counts_size.argtypes = [C.POINTER(_Index)]
        L.oracle_scores.restype = C.c_int
        L.oracle_scores.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_size_t,
                                    C.c_void_p]
        L.oracle_scores_range.restype = C.c_int
        L.oracle_scores_range.argtypes = [C.POINTER(_Index), C.c_char_p,
                                          C.c_size_t, C.c_uint64, C.c_uint64,
                                          C.c_void_p]
        L.oracle_search.restype = C.c_int
        L.oracle_search.argtypes = [C.POINTER(C.POINTER(_Index)), C.c_size_t,
                                    C.c_char_p, C.c_size_t, C.c_double,
                                    C.c_size_t, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]
        L.oracle_index_load.restype = C.c_int
        L.oracle_index_load.argtypes = [C.c_char_p, C.POINTER(_Index)]
        L.oracle_index_free.restype = None
        L.oracle_index_free.argtypes = [C.POINTER(_Index)]
        L.oracle_write_classic.restype = C.c_int
        L.oracle_write_classic.argtypes = [C.c_char_p, C.c_uint32, C.c_uint8,
                                           C.c_uint32, C.c_uint64, C.c_uint64,
                                           C.c_void_p, C.c_void_p]
        L.oracle_write_compact.restype = C.c_int
        L.oracle_write_compact.argtypes = [C.c_char_p, C.c_uint32, C.c_uint8,
                                           C.c_uint32, C.c_uint64, C.c_uint32,
                                           C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.c_void_p]
        L.oracle_fill_word.restype = C.c_uint64
        L.oracle_fill_word.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64,
                                       C.c_uint64]
        L.oracle_index_procedural.restype = C.c_int
        L.oracle_index_procedural.argtypes = [C.POINTER(_Index), C.c_int,
                                              C.c_uint32, C.c_uint8, C.c_uint64,
                                              C.c_uint32, C.c_uint64, C.c_uint32,
                                              C.c_void_p, C.c_uint64]
        L.oracle_index_materialize.restype = C.c_int
        L.oracle_index_materialize.argtypes = [C.POINTER(_Index)]
        L.oracle_write_classic_procedural.restype = C.c_int
        L.oracle_write_classic_procedural.argtypes = [C.c_char_p, C.POINTER(_Index)]
        L.oracle_write_compact_procedural.restype = C.c_int
        L.oracle_write_compact_procedural.argtypes = [C.c_char_p, C.POINTER(_Index)]
        L.oracle_random_query.restype = None
        L.oracle_random_query.argtypes = [C.c_uint64, C.c_size_t, C.c_char_p]
        _lib = L
    return _lib


def _b(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode("ascii")


def xxh64(data, seed=0):
    data = _b(data)
    return lib().oracle_xxh64(data, len(data), seed)


def canonicalize_kmer(kmer):
    """returns (canonical bytes, good)"""
    kmer = _b(kmer)
    out = C.create_string_buffer(len(kmer) + 1)
    good = lib().oracle_canonicalize_kmer(kmer, out, len(kmer))
    return out.raw[:len(kmer)], bool(good)


def create_hashes(query, term_size, num_hashes, canonicalize):
    query = _b(query)
    if len(query) < term_size:
        raise OracleError(ERR_TOO_SHORT, "create_hashes")
    n = (len(query) - term_size + 1) * num_hashes
    out = np.zeros(max(n, 1), dtype=np.uint64)
    rc = lib().oracle_create_hashes(query, len(query), term_size, num_hashes,
                                    canonicalize, out.ctypes.data)
    if rc != OK:
        raise OracleError(rc, "create_hashes")
    return out[:n]


def random_query(seed, length):
    out = C.create_string_buffer(length + 1)
    lib().oracle_random_query(seed, length, out)
    return out.raw[:length]


def fill_word(seed, page, row, word):
    return lib().oracle_fill_word(seed, page, row, word)


class Index:
    """One classic/compact index held by the oracle (file-backed, in-memory or
    procedural)."""

    def __init__(self):
        self._c = _Index()
        self._keep = []
        self._open = False

    # -- constructors ------------------------------------------------------
    @classmethod
    def load(cls, path):
        self = cls()
        rc = lib().oracle_index_load(_b(path), C.byref(self._c))
        if rc != OK:
            raise OracleError(rc, "index_load(%s)" % path)
        self._open = True
        return self

    @classmethod
    def procedural(cls, kind, n_docs, signature_sizes, num_hashes, page_size=0,
                   term_size=31, canonicalize=1, fill_seed=1, materialize=False):
        self = cls()
        sig = np.ascontiguousarray(signature_sizes, dtype=np.uint64)
        rc = lib().oracle_index_procedural(
            C.byref(self._c), kind, term_size, canonicalize, num_hashes, n_docs,
            page_size, len(sig), sig.ctypes.data, fill_seed)
        if rc != OK:
            raise OracleError(rc, "index_procedural")
        self._open = True
        if materialize:
            rc = lib().oracle_index_materialize(C.byref(self._c))
            if rc != OK:
                raise OracleError(rc, "index_materialize")
        return self

    @classmethod
    def from_arrays(cls, kind, n_docs, pages, num_hashes, term_size=31,
                    canonicalize=1, page_size=None):
        """pages: list of uint8 arrays of shape [sig_p, page_size]"""
        self = cls()
        c = self._c
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        if kind == KIND_CLASSIC:
            assert len(pages) == 1
            page_size = (n_docs + 7) // 8
        elif page_size is None:
            page_size = pages[0].shape[1]
        for p in pages:
            assert p.ndim == 2 and p.shape[1] == page_size
        c.kind = kind
        c.term_size = term_size
        c.canonicalize = canonicalize
        c.num_hashes = num_hashes
        c.n_docs = n_docs
        c.n_pages = len(pages)
        c.page_size = page_size
        sig = np.array([p.shape[0] for p in pages], dtype=np.uint64)
        ptrs = (C.POINTER(C.c_uint8) * len(pages))(
            *[p.ctypes.data_as(C.POINTER(C.c_uint8)) for p in pages])
        c.signature_sizes = sig.ctypes.data_as(C.POINTER(C.c_uint64))
        c.page_data = ptrs
        self._keep = [pages, sig, ptrs]
        self._borrowed = True
        return self

    def close(self):
        if self._open:
            lib().oracle_index_free(C.byref(self._c))
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry ----------------------------------------------------------
    @property
    def kind(self): return self._c.kind
    @property
    def term_size(self): return self._c.term_size
    @property
    def canonicalize(self): return self._c.canonicalize
    @property
    def num_hashes(self): return self._c.num_hashes
    @property
    def n_docs(self): return self._c.n_docs
    @property
    def n_pages(self): return self._c.n_pages
    @property
    def page_size(self): return self._c.page_size
    @property
    def signature_sizes(self):
        return [self._c.signature_sizes[i] for i in range(self._c.n_pages)]
    @property
    def counts_size(self): return lib().oracle_counts_size(C.byref(self._c))
    @property
    def doc_names(self):
        if not self._c.doc_names:
            return ["doc_%06d" % i for i in range(self.n_docs)]
        return [self._c.doc_names[i].decode() for i in range(self.n_docs)]

    def page_array(self, p):
        """numpy view [sig_p, page_size] of page p (materialised indices only)"""
        n = self.signature_sizes[p] * self.page_size
        buf = (C.c_uint8 * n).from_address(
            C.addressof(self._c.page_data[p].contents))
        a = np.frombuffer(buf, dtype=np.uint8).reshape(-1, self.page_size)
        return a

    # -- queries -----------------------------------------------------------
    def scores(self, query, doc_begin=None, doc_end=None):
        query = _b(query)
        if doc_begin is None:
            doc_begin, doc_end = 0, self.counts_size
        out = np.zeros(max(doc_end - doc_begin, 1), dtype=np.uint32)
        rc = lib().oracle_scores_range(C.byref(self._c), query, len(query),
                                       doc_begin, doc_end, out.ctypes.data)
        if rc != OK:
            raise OracleError(rc, "scores")
        return out[:doc_end - doc_begin]

    def write(self, path):
        fn = (lib().oracle_write_classic_procedural if self.kind == KIND_CLASSIC
              else lib().oracle_write_compact_procedural)
        rc = fn(_b(path), C.byref(self._c))
        if rc != OK:
            raise OracleError(rc, "write(%s)" % path)


def search(indices, query, threshold=0.0, num_results=0):
    """ClassicSearch::search over a list of Index objects.
    Returns a list of (file, doc, score)."""
    if isinstance(indices, Index):
        indices = [indices]
    query = _b(query)
    arr = (C.POINTER(_Index) * len(indices))(
        *[C.pointer(i._c) for i in indices])
    cap = sum(i.n_docs for i in indices)
    f = np.zeros(max(cap, 1), dtype=np.uint32)
    d = np.zeros(max(cap, 1), dtype=np.uint32)
    s = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_size_t(0)
    rc = lib().oracle_search(arr, len(indices), query, len(query), threshold,
                             num_results, f.ctypes.data, d.ctypes.data,
                             s.ctypes.data, cap, C.byref(n))
    if rc != OK:
        raise OracleError(rc, "search")
    k = n.value
    return list(zip(f[:k].tolist(), d[:k].tolist(), s[:k].tolist()))


def write_classic(path, n_docs, matrix, num_hashes, term_size=31, canonicalize=1,
                  names=None):
    matrix = np.ascontiguousarray(matrix, dtype=np.uint8)
    assert matrix.ndim == 2 and matrix.shape[1] == (n_docs + 7) // 8
    nm = None
    if names is not None:
        nm = (C.c_char_p * n_docs)(*[_b(x) for x in names])
    rc = lib().oracle_write_classic(_b(path), term_size, canonicalize, n_docs,
                                    matrix.shape[0], num_hashes, nm,
                                    matrix.ctypes.data)
    if rc != OK:
        raise OracleError(rc, "write_classic")


def write_compact(path, n_docs, pages, num_hashes, term_size=31, canonicalize=1,
                  names=None):
    pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
    page_size = pages[0].shape[1]
    sig = np.array([p.shape[0] for p in pages], dtype=np.uint64)
    data = np.concatenate([p.reshape(-1) for p in pages])
    nm = None
    if names is not None:
        nm = (C.c_char_p * n_docs)(*[_b(x) for x in names])
    rc = lib().oracle_write_compact(_b(path), term_size, canonicalize, n_docs,
                                    page_size, len(pages), sig.ctypes.data,
                                    num_hashes, nm, data.ctypes.data)
    if rc != OK:
        raise OracleError(rc, "write_compact")
