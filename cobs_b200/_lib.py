"""ctypes loader for cobs_b200/lib/libcobsgpu.so (the C ABI of include/cobsgpu.h).

There is no CPU fallback anywhere in this package: if the CUDA library has not been built
(`make` / `__graft_entry__.build()`), importing it raises; if no GPU is present the compute
entry points return COBSGPU_ERR_CUDA, surfaced as CobsGpuError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcobsgpu.so")

OK = 0
ERR_INVALID_ARG = 1
ERR_CUDA = 2
ERR_OOM = 3
ERR_QUERY_TOO_SHORT = 4
ERR_INVALID_BASE = 5
ERR_BAD_FILE = 6
ERR_IO = 7

KIND_CLASSIC = 0
KIND_COMPACT = 1


class CobsGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cobsgpu error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class IndexDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("kind", C.c_int32),
        ("term_size", C.c_uint32),
        ("canonicalize", C.c_uint32),
        ("num_hashes", C.c_uint32),
        ("n_docs", C.c_uint32),
        ("n_pages", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("page_size", C.c_uint64),
        ("signature_sizes", C.POINTER(C.c_uint64)),
        ("page_data", C.POINTER(C.c_void_p)),
        ("fill_seed", C.c_uint64),
        ("device", C.c_int32),
        ("shard_index", C.c_uint32),
        ("shard_count", C.c_uint32),
        ("reserved1", C.c_uint32),
    ]


class ConstructDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("term_size", C.c_uint32),
        ("canonicalize", C.c_uint32),
        ("num_hashes", C.c_uint32),
        ("signature_size", C.c_uint64),
        ("false_positive_rate", C.c_double),
        ("n_docs", C.c_uint32),
        ("n_seqs", C.c_uint32),
        ("doc_names", C.POINTER(C.c_char_p)),
        ("sequences", C.c_char_p),
        ("seq_offsets", C.POINTER(C.c_uint64)),
        ("seq_doc", C.POINTER(C.c_uint32)),
        ("device", C.c_int32),
        ("reserved", C.c_uint32),
    ]


class IndexInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("term_size", C.c_uint32),
        ("canonicalize", C.c_uint32),
        ("num_hashes", C.c_uint32),
        ("n_docs", C.c_uint32),
        ("n_pages", C.c_uint32),
        ("page_size", C.c_uint64),
        ("row_size", C.c_uint64),
        ("counts_size", C.c_uint64),
        ("shard_index", C.c_uint32),
        ("shard_count", C.c_uint32),
        ("shard_doc_begin", C.c_uint32),
        ("shard_doc_end", C.c_uint32),
        ("hbm_bytes", C.c_uint64),
        ("bytes_per_kmer", C.c_uint64),
        ("load_seconds", C.c_double),
        ("load_bytes", C.c_uint64),
        ("load_threads", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("offsets", C.POINTER(C.c_uint64)),
        ("doc", C.POINTER(C.c_uint32)),
        ("score", C.POINTER(C.c_uint32)),
    ]


class Timers(C.Structure):
    _fields_ = [
        ("hashes_ms", C.c_double),
        ("score_ms", C.c_double),
        ("select_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("kernel_launches", C.c_uint64),
        ("score_launches", C.c_uint64),
        ("kmers", C.c_uint64),
        ("queries", C.c_uint64),
    ]


# every symbol include/cobsgpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cobsgpu_last_error": (C.c_char_p, []),
    "cobsgpu_version": (C.c_int, []),
    "cobsgpu_device_count": (C.c_int, []),
    "cobsgpu_index_open": (C.c_int, [C.POINTER(IndexDesc), C.POINTER(C.c_void_p)]),
    "cobsgpu_index_open_file": (C.c_int, [C.c_char_p, C.c_int, C.c_uint32, C.c_uint32,
                                          C.POINTER(C.c_void_p)]),
    "cobsgpu_construct_classic": (C.c_int, [C.POINTER(ConstructDesc), C.POINTER(C.c_void_p)]),
    "cobsgpu_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cobsgpu_index_close": (None, [C.c_void_p]),
    "cobsgpu_index_get_info": (C.c_int, [C.c_void_p, C.POINTER(IndexInfo)]),
    "cobsgpu_index_signature_size": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "cobsgpu_index_doc_name": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "cobsgpu_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "cobsgpu_hash": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cobsgpu_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cobsgpu_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_double, C.c_uint64, C.POINTER(Result)]),
    "cobsgpu_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                 C.c_uint64, C.POINTER(C.c_uint64)]),
    "cobsgpu_collect": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(Result)]),
    "cobsgpu_group_open_file": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.c_uint32,
                                          C.POINTER(C.c_void_p)]),
    "cobsgpu_group_open": (C.c_int, [C.POINTER(IndexDesc), C.POINTER(C.c_int32), C.c_uint32,
                                     C.POINTER(C.c_void_p)]),
    "cobsgpu_group_close": (None, [C.c_void_p]),
    "cobsgpu_group_size": (C.c_uint32, [C.c_void_p]),
    "cobsgpu_group_shard": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "cobsgpu_group_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                             C.c_double, C.c_uint64, C.POINTER(Result)]),
    "cobsgpu_search_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                              C.c_double, C.c_uint64, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "cobsgpu_merge_device": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                       C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                       C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cobsgpu_get_timers": (C.c_int, [C.c_void_p, C.POINTER(Timers)]),
    "cobsgpu_reset_timers": (C.c_int, [C.c_void_p]),
    "cobsgpu_debug_read_row": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_void_p]),
}

_lib = None


def lib():
    """Load libcobsgpu.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "cobs_b200: %s is missing -- build it with `make` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise CobsGpuError(rc, lib().cobsgpu_last_error().decode(errors="replace"))
