"""CPU robustness test of the index-file parser behind cobsgpu_index_open_file: truncated and
corrupted copies of the golden files must be rejected with a status code (BAD_FILE / IO /
INVALID_ARG; on a box without a GPU a still-valid header ends in ERR_CUDA) -- never a crash, a
hang or a giant allocation.  Runs in a child process so that a crash cannot take pytest down."""
import os
import subprocess
import sys

from conftest import ROOT

CHILD = r'''
import os, random, sys
sys.path.insert(0, sys.argv[1])
import ctypes as C
from cobs_b200 import _lib
L = _lib.lib()
golden = os.path.join(sys.argv[1], "tests", "golden")
tmp = sys.argv[2]
rng = random.Random(7)
allowed = {_lib.ERR_BAD_FILE, _lib.ERR_IO, _lib.ERR_INVALID_ARG, _lib.ERR_CUDA, _lib.ERR_OOM}
n = 0
for name in ("all160.cobs_classic", "all160.cobs_compact", "python_test.cobs_compact"):
    data = open(os.path.join(golden, name), "rb").read()
    header = min(len(data), 2600)
    variants = [data[:cut] for cut in list(range(0, 80)) + [rng.randrange(80, len(data)) for _ in range(40)]]
    for _ in range(250):
        b = bytearray(data)
        for _ in range(rng.randrange(1, 4)):
            pos = rng.randrange(0, header)
            b[pos] = rng.randrange(256)
        variants.append(bytes(b))
    for _ in range(40):                      # blow up the 32/64-bit count fields
        b = bytearray(data)
        pos = rng.randrange(18, 60)
        b[pos:pos + 4] = b"\xff\xff\xff\x7f"
        variants.append(bytes(b))
    for v in variants:
        p = os.path.join(tmp, "fuzz.idx")
        with open(p, "wb") as f:
            f.write(v)
        h = C.c_void_p()
        rc = L.cobsgpu_index_open_file(p.encode(), 0, 0, 1, C.byref(h))
        if rc == 0:
            L.cobsgpu_index_close(h)         # only possible on a GPU box: a harmless mutation
        else:
            assert rc in allowed, (name, rc, L.cobsgpu_last_error())
            assert L.cobsgpu_last_error()
        n += 1
print("fuzzed", n)
'''


def test_corrupt_index_files_are_rejected_cleanly(tmp_path):
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, str(tmp_path)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert "fuzzed 1230" in r.stdout


def test_valid_files_pass_the_parser():
    """every golden file written by the reference gets through header parsing and the size
    checks; without a GPU the call then stops at the device (ERR_CUDA), never at the parser"""
    import ctypes as C
    import glob

    import pytest
    import cobs_b200
    from cobs_b200 import _lib
    L = cobs_b200.lib()
    if L.cobsgpu_device_count() > 0:
        pytest.skip("GPU present: the loader itself is covered by the gpu tests")
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.cobs_*")))
    assert len(files) >= 14
    for p in files:
        h = C.c_void_p()
        assert L.cobsgpu_index_open_file(p.encode(), 0, 0, 1, C.byref(h)) == _lib.ERR_CUDA, p
