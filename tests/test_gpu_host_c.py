"""The C host mirror (host/dna_host.c: same names and error texts as the reference, work done by
libdnagpu) and the stand-alone C harness, on the GPU, against the oracle."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


class Kmer(C.Structure):
    _fields_ = [("length", C.c_int32), ("bit_sequence", C.c_uint64)]


class Qkmer(C.Structure):
    _fields_ = [("sequence", C.c_char * 33)]


class KmerCount(C.Structure):
    _fields_ = [("kmer", Kmer), ("count", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("total", C.c_uint64), ("distinct", C.c_uint64), ("unique", C.c_uint64)]


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(PKG, "libdnahost.so"))
    lib.dna_make.restype = C.c_void_p
    lib.dna_make.argtypes = [C.c_char_p]
    lib.dna_free.argtypes = [C.c_void_p]
    lib.dnah_last_error.restype = C.c_char_p
    lib.dnah_open.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.dnah_close.argtypes = [C.c_void_p]
    lib.kmer_make.argtypes = [C.c_char_p, C.POINTER(Kmer)]
    lib.qkmer_make.argtypes = [C.c_char_p, C.POINTER(Qkmer)]
    lib.generate_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.POINTER(Kmer)), C.POINTER(C.c_uint64)]
    lib.generate_kmers_where.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Kmer), C.POINTER(Qkmer),
                                         C.POINTER(C.POINTER(Kmer)), C.POINTER(C.c_uint64)]
    lib.count_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Kmer), C.POINTER(Qkmer),
                                C.POINTER(C.POINTER(KmerCount)), C.POINTER(C.c_uint64), C.POINTER(Stats)]
    s = C.c_void_p()
    assert lib.dnah_open(C.byref(s), 0) == 0, lib.dnah_last_error()
    yield lib, s
    lib.dnah_close(s)


def test_reference_statements_through_the_c_mirror(host, kats):
    lib, s = host
    for v in kats["generate_kmers"]:
        d = lib.dna_make(v["dna"].encode())
        rows, n = C.POINTER(Kmer)(), C.c_uint64()
        assert lib.generate_kmers(s, d, v["k"], C.byref(rows), C.byref(n)) == 0
        assert [R.decode_kmer(rows[i].bit_sequence, rows[i].length) for i in range(n.value)] == v["rows"]
    for v in kats["starts_with"]:
        d = lib.dna_make(v["dna"].encode())
        p = Kmer()
        lib.kmer_make(v["prefix"].encode(), C.byref(p))
        rows, n = C.POINTER(Kmer)(), C.c_uint64()
        assert lib.generate_kmers_where(s, d, v["k"], C.byref(p), None, C.byref(rows), C.byref(n)) == 0
        assert [R.decode_kmer(rows[i].bit_sequence, v["k"]) for i in range(n.value)] == v["rows"]
    for v in kats["contains"]:
        d = lib.dna_make(v["dna"].encode())
        q = Qkmer()
        assert lib.qkmer_make(v["pattern"].encode(), C.byref(q)) == 0
        rows, n = C.POINTER(Kmer)(), C.c_uint64()
        assert lib.generate_kmers_where(s, d, v["k"], None, C.byref(q), C.byref(rows), C.byref(n)) == 0
        assert [R.decode_kmer(rows[i].bit_sequence, v["k"]) for i in range(n.value)] == v["rows"]
    for v in kats["group_by"]:
        d = lib.dna_make(v["dna"].encode())
        rows, n, st = C.POINTER(KmerCount)(), C.c_uint64(), Stats()
        assert lib.count_kmers(s, d, v["k"], None, None, C.byref(rows), C.byref(n), C.byref(st)) == 0
        got = {R.decode_kmer(rows[i].kmer.bit_sequence, v["k"]): rows[i].count for i in range(n.value)}
        assert got == v["counts"]
    for v in kats["stats"]:
        d = lib.dna_make(v["dna"].encode())
        st = Stats()
        assert lib.count_kmers(s, d, v["k"], None, None, None, None, C.byref(st)) == 0
        assert (st.total, st.distinct, st.unique) == (v["total"], v["distinct"], v["unique"])


def test_c_mirror_error_behaviour(host):
    lib, s = host
    d = lib.dna_make(b"ACGTACGTAC")
    rows, n = C.POINTER(Kmer)(), C.c_uint64()
    assert lib.generate_kmers(s, d, 0, C.byref(rows), C.byref(n)) != 0
    assert lib.dnah_last_error() == b"Invalid k value: must be between 1 and 32"
    p = Kmer()
    lib.kmer_make(b"ACGT", C.byref(p))
    assert lib.generate_kmers_where(s, d, 3, C.byref(p), None, C.byref(rows), C.byref(n)) != 0
    assert lib.dnah_last_error() == b"Prefix length cannot exceed kmer length"
    q = Qkmer()
    lib.qkmer_make(b"NN", C.byref(q))
    assert lib.generate_kmers_where(s, d, 3, None, C.byref(q), C.byref(rows), C.byref(n)) != 0
    assert lib.dnah_last_error() == b"Qkmer pattern and kmer lengths do not match"


def test_c_harness_runs_and_agrees_with_oracle():
    exe = os.path.join(PKG, "dnagpu_bench")
    for extra in ([], ["--host"], ["--prefix", "AC"]):
        out = subprocess.run([exe, "--bases", "3000000", "--k", "21", "--seed", "7", "--steps", "2"] + extra,
                             capture_output=True, text=True, check=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        words = R.synth_seq(7, 3_000_000)
        want = R.count_query(words, 1, 3_000_000, words.size, 21, prefix=R.kmer_make("AC") if extra[:1] == ["--prefix"] else None,
                             faithful=False, threads=4, want_rows=False)
        assert (d["total"], d["distinct"], d["unique"]) == want.stats, extra
