"""cobs_index -- the import name of the reference's Python module (python/module.cpp), served by
the B200 query path: `import cobs_index as cobs; cobs.Search(path).search(query)` works
unchanged (python/tests/test_cobs_index.py:36-40, 57-61).

In scope: Search / SearchResult (module.cpp:351-386) and, because a device-side classic
construction exists (scope row f4), classic_construct over FASTA / plain-text documents with
ClassicIndexParameters (module.cpp:127-236).  Compact construction and the other document
formats are outside the query hot path and raise NotImplementedError.
"""
import os

import cobs_b200

__version__ = "b200-dev"


class SearchResult:
    """Return objects for Search (module.cpp:351-363): mutable, `doc_name` and `score`"""

    def __init__(self, doc_name="", score=0):
        self.doc_name = doc_name
        self.score = score

    def __repr__(self):
        return "SearchResult(doc_name=%r, score=%d)" % (self.doc_name, self.score)

    def __iter__(self):
        return iter((self.doc_name, self.score))


class Search:
    """Search object to run queries on COBS indices (module.cpp:367-386); loads the given
    classic or compact index file into HBM."""

    def __init__(self, index_path):
        self._s = cobs_b200.Search(index_path)

    def search(self, query, threshold=0.0, num_results=0):
        return [SearchResult(r.doc_name, r.score)
                for r in self._s.search(query, threshold, num_results)]

    def search_batch(self, queries, threshold=0.0, num_results=0):
        """extension: many queries per call (one GPU batch)"""
        return [[SearchResult(r.doc_name, r.score) for r in res]
                for res in self._s.search_batch(queries, threshold, num_results)]


def disable_cache(disable=True):
    """module.cpp:61-67: the FastA/FastQ index caches belong to construction; nothing to do"""


# ------------------------------------------------------------------------------------------
# classic construction on the device (FASTA / text documents only)

_FASTA = (".fa", ".fasta", ".fna", ".ffn", ".faa", ".frn")
_TEXT = (".txt",)


class DocumentEntry:
    def __init__(self, path):
        self.path = path
        base = os.path.basename(path)
        self.name = base[:base.rfind(".")] if "." in base else base
        self.size = os.path.getsize(path)


class DocumentList:
    """module.cpp:73-121, restricted to FASTA (one document per file) and plain text files"""

    def __init__(self, root=None):
        self.list = []
        if root is not None:
            self.add_recursive(root)

    def add(self, path):
        if path.lower().endswith(_FASTA + _TEXT):
            self.list.append(DocumentEntry(path))

    def add_recursive(self, root):
        if os.path.isfile(root):
            self.add(root)
            return
        for dirpath, _, files in sorted(os.walk(root)):
            for f in sorted(files):
                self.add(os.path.join(dirpath, f))

    def size(self):
        return len(self.list)

    def __len__(self):
        return len(self.list)

    def sort_by_path(self):
        self.list.sort(key=lambda e: e.path)

    def sort_by_size(self):
        self.list.sort(key=lambda e: e.size)


class ClassicIndexParameters:
    """module.cpp:127-168, defaults of cobs/construction/classic_index.hpp:27-58"""

    def __init__(self):
        self.term_size = 31
        self.canonicalize = 1
        self.num_hashes = 1
        self.false_positive_rate = 0.3
        self.signature_size = 0
        self.mem_bytes = 0
        self.num_threads = 0
        self.log_prefix = ""
        self.clobber = False
        self.continue_ = False
        self.keep_temporary = False


def _read_document(path):
    """sequences of one document: the records of a FASTA file, or the whole text file"""
    with open(path, "rb") as f:
        data = f.read()
    if not path.lower().endswith(_FASTA):
        return [data]
    # records the way the reference walks a FASTA file (cobs/fasta_file.hpp:156-182): '>' / ';'
    # lines and empty lines end a record, the other lines of a record are concatenated
    seqs, cur = [], []
    for line in data.split(b"\n"):
        if not line or line[:1] in (b">", b";"):
            if cur:
                seqs.append(b"".join(cur))
            cur = []
        else:
            cur.append(line)
    if cur:
        seqs.append(b"".join(cur))
    return seqs


def classic_construct(input=None, out_file=None, index_params=None, file_type="any",
                      tmp_path="", list=None):
    """module.cpp:172-236: builds a classic index (on the GPU) and writes it in the
    reference's file format"""
    p = index_params or ClassicIndexParameters()
    docs = list if list is not None else DocumentList(input)
    if isinstance(docs, str):
        docs = DocumentList(docs)
    if os.path.exists(out_file) and not p.clobber:
        raise RuntimeError("Output file exists, will not overwrite without --clobber.")
    docs.sort_by_path()
    documents = [(e.name, _read_document(e.path)) for e in docs.list]
    g = cobs_b200.GpuIndex.construct_classic(
        documents, num_hashes=p.num_hashes, false_positive_rate=p.false_positive_rate,
        term_size=p.term_size, canonicalize=p.canonicalize, signature_size=p.signature_size)
    try:
        g.save(out_file)
    finally:
        g.close()


class CompactIndexParameters:
    def __init__(self):
        raise NotImplementedError("compact index construction is outside the B200 query path")


def compact_construct(*a, **kw):
    raise NotImplementedError("compact index construction is outside the B200 query path")
