"""CPU tests of the host-side C++ layer: binaries exist, usage text, and -- without a GPU --
the query path fails loudly instead of falling back to the CPU."""
import os
import subprocess

import pytest

import cobs_b200
from conftest import ROOT, golden_path

COBS = os.path.join(ROOT, "build", "cobs")


def test_binaries_built():
    for f in ("cobs", "host_tests", "host_unit_tests", "kernel_unit_tests", "libcobs_b200.so"):
        assert os.path.exists(os.path.join(ROOT, "build", f)), f


def test_host_unit_tests_binary():
    """Timer, header sniffing, error conventions of the C++ mirror (no GPU needed)"""
    exe = os.path.join(ROOT, "build", "host_unit_tests")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_unit_tests: ok" in r.stdout
    # assert_exit: message on stderr + exit(EXIT_FAILURE) (cobs/util/error_handling.cpp:19-28)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden"), "exit_error"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1
    assert r.stderr.strip() == "query too short, needs to be at least 31 characters long"


def test_kernel_arithmetic_unit_tests():
    """the __host__ __device__ arithmetic of the kernels (XXH64 forms, canonical k-mers, bit-sliced
    counter helpers, sort keys, procedural fill), executed on the host against the oracle"""
    r = subprocess.run([os.path.join(ROOT, "build", "kernel_unit_tests")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernel_unit_tests: ok" in r.stdout


def test_usage_and_version():
    r = subprocess.run([COBS], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "query" in r.stdout
    r = subprocess.run([COBS, "version"], stdout=subprocess.PIPE, text=True)
    assert "C ABI version 2" in r.stdout
    r = subprocess.run([COBS, "query", "--help"], stdout=subprocess.PIPE, text=True)
    for flag in ("--index", "--file", "--threshold", "--limit", "--load-complete", "--threads"):
        assert flag in r.stdout       # the reference's flags (src/cobs.cpp:474-505)
    assert "default: 0.8" in r.stdout


@pytest.mark.skipif(cobs_b200.lib().cobsgpu_device_count() > 0, reason="GPU present")
def test_cli_without_gpu_fails_loudly():
    r = subprocess.run([COBS, "query", "-i", golden_path("all160.cobs_classic"), "ACGT" * 10],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert r.stdout == ""
