"""ctypes view of oracle/_ref/libcobs_ref_*.so -- the UNMODIFIED reference, built by
`make -C oracle ref` from /root/reference.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used to (a) pin the C restatement, (b) generate tests/golden/, (c) time the
reference's CPU path for bench.py (cpu_baseline.kind == "reference").
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return all(x in fl for x in ("avx2", "bmi2", "fma"))
    except OSError:
        pass
    return False


def lib_path():
    v = "v3" if _has_avx2() else "v1"
    return os.path.join(_HERE, "_ref", "libcobs_ref_%s.so" % v)


def available():
    return os.path.exists(lib_path())


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(lib_path())
        L.ref_last_error.restype = C.c_char_p
        L.ref_xxh64.restype = C.c_uint64
        L.ref_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.ref_canonicalize_kmer.restype = C.c_int
        L.ref_canonicalize_kmer.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.ref_random_sequence.restype = None
        L.ref_random_sequence.argtypes = [C.c_size_t, C.c_size_t, C.c_char_p]
        L.ref_random_queries_mt19937.restype = None
        L.ref_random_queries_mt19937.argtypes = [C.c_size_t, C.c_size_t,
                                                 C.c_size_t, C.c_char_p]
        L.ref_set_threads.argtypes = [C.c_size_t]
        L.ref_set_load_complete.argtypes = [C.c_int]
        L.ref_set_disable.argtypes = [C.c_int] * 4
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.POINTER(C.c_char_p), C.c_size_t]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_num_docs.restype = C.c_uint32
        L.ref_num_docs.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_doc_name.restype = C.c_char_p
        L.ref_doc_name.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_counts_size.restype = C.c_uint64
        L.ref_counts_size.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_search.restype = C.c_int
        L.ref_search.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_double,
                                 C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_bench.restype = C.c_int
        L.ref_bench.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t,
                                C.c_double, C.c_size_t, C.POINTER(C.c_double),
                                C.POINTER(C.c_uint64)]
        L.ref_classic_construct.restype = C.c_int
        L.ref_classic_construct.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p,
                                            C.c_uint, C.c_uint, C.c_double,
                                            C.c_int, C.c_uint64]
        L.ref_compact_construct.restype = C.c_int
        L.ref_compact_construct.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p,
                                            C.c_uint, C.c_uint, C.c_double,
                                            C.c_int, C.c_uint64]
        L.ref_classic_construct_random.restype = C.c_int
        L.ref_classic_construct_random.argtypes = [C.c_char_p, C.c_uint64,
                                                   C.c_uint64, C.c_size_t,
                                                   C.c_uint64, C.c_size_t]
        L.ref_write_kmer_doc.restype = C.c_int
        L.ref_write_kmer_doc.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p,
                                         C.c_size_t, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


class RefError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise RefError("%s: %s" % (what, lib().ref_last_error().decode()))


def _b(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode("ascii")


def xxh64(data, seed=0):
    data = _b(data)
    return lib().ref_xxh64(data, len(data), seed)


def canonicalize_kmer(kmer):
    kmer = _b(kmer)
    out = C.create_string_buffer(len(kmer) + 1)
    good = lib().ref_canonicalize_kmer(kmer, out, len(kmer))
    return out.raw[:len(kmer)], bool(good)


def random_sequence(size, seed):
    """cobs::random_sequence (util/misc.cpp:32-35, std::default_random_engine)"""
    out = C.create_string_buffer(size + 1)
    lib().ref_random_sequence(size, seed, out)
    return out.raw[:size]


def random_queries_mt19937(seed, n, length):
    """the query generator of `cobs benchmark-fpr` (src/cobs.cpp:709-720)"""
    out = C.create_string_buffer(n * length + 1)
    lib().ref_random_queries_mt19937(seed, n, length, out)
    raw = out.raw
    return [raw[i * length:(i + 1) * length] for i in range(n)]


def set_threads(n):
    lib().ref_set_threads(n)


def set_load_complete(flag):
    lib().ref_set_load_complete(1 if flag else 0)


def set_disable(d8=False, d16=False, d32=False, dsse2=False):
    lib().ref_set_disable(int(d8), int(d16), int(d32), int(dsse2))


class Search:
    """cobs::ClassicSearch over one or more index files (the real reference)."""

    def __init__(self, paths):
        if isinstance(paths, (str, bytes)):
            paths = [paths]
        arr = (C.c_char_p * len(paths))(*[_b(p) for p in paths])
        self._h = lib().ref_open(arr, len(paths))
        if not self._h:
            raise RefError("open: " + lib().ref_last_error().decode())
        self.n_files = len(paths)
        self.n_docs = [lib().ref_num_docs(self._h, i) for i in range(len(paths))]

    def close(self):
        if self._h:
            lib().ref_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def doc_name(self, file, doc):
        return lib().ref_doc_name(self._h, file, doc).decode()

    def counts_size(self, file=0):
        return lib().ref_counts_size(self._h, file)

    def search(self, query, threshold=0.0, num_results=0):
        """list of (file, doc, score) in the reference's output order"""
        query = _b(query)
        cap = max(sum(self.n_docs), 1)
        f = np.zeros(cap, dtype=np.uint32)
        d = np.zeros(cap, dtype=np.uint32)
        s = np.zeros(cap, dtype=np.uint32)
        n = C.c_size_t(0)
        rc = lib().ref_search(self._h, query, len(query), threshold, num_results,
                              f.ctypes.data, d.ctypes.data, s.ctypes.data, cap,
                              C.byref(n))
        _check(rc, "search")
        k = n.value
        return list(zip(f[:k].tolist(), d[:k].tolist(), s[:k].tolist()))

    def bench(self, queries, threshold=0.8, num_results=0):
        """wall seconds for a search() loop over `queries` (list of bytes)"""
        blob = b"".join(queries)
        off = np.zeros(len(queries) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(q) for q in queries])
        sec = C.c_double(0)
        tot = C.c_uint64(0)
        rc = lib().ref_bench(self._h, blob, off.ctypes.data, len(queries),
                             threshold, num_results, C.byref(sec), C.byref(tot))
        _check(rc, "bench")
        return sec.value, tot.value


def classic_construct(in_dir, out_file, tmp_dir, term_size=31, num_hashes=1,
                      fpr=0.3, canonicalize=1, signature_size=0):
    _check(lib().ref_classic_construct(_b(in_dir), _b(out_file), _b(tmp_dir),
                                       term_size, num_hashes, fpr, canonicalize,
                                       signature_size), "classic_construct")


def compact_construct(in_dir, out_file, tmp_dir, term_size=31, num_hashes=1,
                      fpr=0.3, canonicalize=1, page_size=4096):
    _check(lib().ref_compact_construct(_b(in_dir), _b(out_file), _b(tmp_dir),
                                       term_size, num_hashes, fpr, canonicalize,
                                       page_size), "compact_construct")


def classic_construct_random(out_file, signature_size, num_documents,
                             document_size, num_hashes, seed):
    _check(lib().ref_classic_construct_random(_b(out_file), signature_size,
                                              num_documents, document_size,
                                              num_hashes, seed),
           "classic_construct_random")


def write_kmer_doc(path, name, seq, positions):
    seq = _b(seq)
    pos = np.ascontiguousarray(positions, dtype=np.uint64)
    _check(lib().ref_write_kmer_doc(_b(path), _b(name), seq, len(seq),
                                    pos.ctypes.data, len(pos)), "write_kmer_doc")
