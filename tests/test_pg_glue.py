"""The PostgreSQL glue (pg/dna_gpu.c) and the host C mirror compile cleanly; the glue is checked
against the PostgreSQL API shim the oracle owns (there is no server in the image).  No compute."""
import os
import subprocess

from conftest import PKG, ROOT


def test_pg_glue_compiles_against_the_fmgr_api():
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                    "-I", os.path.join(ROOT, "oracle", "pgshim"), os.path.join(PKG, "pg", "dna_gpu.c")],
                   check=True)


def test_reference_module_links_with_the_glue_in_place_of_generate_kmers():
    """dna.c (generate_kmers renamed away, as pg/Makefile does) + pg/dna_gpu.c + libdnagpu link into one module
    that exports the SQL-visible symbols; the scalar functions of the reference still answer through it.
    (Calling the GPU entry points is tests/test_gpu_pg_glue.py.)"""
    import pytest
    from oracle import ref_real
    if not ref_real.glue_available():
        pytest.skip("oracle/_ref/libdnaglue.so not built and /root/reference absent")
    glue = ref_real.glue()
    out = subprocess.run(["nm", "-D", "--defined-only", ref_real.GLUE_SO], capture_output=True, text=True,
                         check=True).stdout
    for sym in ("generate_kmers", "generate_kmers_cpu", "generate_kmers_where", "kmer_stats", "count_kmers",
                "kmer_stats_agg_trans", "kmer_stats_agg_final", "starts_with", "contains",
                "kmer_hash", "kmer_eq", "dna_in", "kmer_in", "qkmer_in"):
        assert f" T {sym}\n" in out, sym
    assert glue.kmer_in("ACGT") == (0x78, 4)
    assert glue.kmer_out(0xE4, 4) == "ATCG"


def test_host_mirror_and_c_harness_build():
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(PKG, "libdnahost.so"))
    assert os.path.exists(os.path.join(PKG, "dnagpu_bench"))
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libdnahost.so")], capture_output=True,
                         text=True, check=True).stdout
    for sym in ("dna_make", "kmer_make", "qkmer_make", "generate_kmers", "generate_kmers_where", "count_kmers",
                "dnah_open", "dnah_close"):
        assert f" T {sym}\n" in out, sym


def test_host_mirror_scalar_glue_matches_oracle(ref):
    """dna_make / kmer_make / qkmer_make of the host mirror against the oracle (CPU-only functions)."""
    import ctypes as C
    lib = C.CDLL(os.path.join(PKG, "libdnahost.so"))
    lib.dna_make.restype = C.c_void_p
    lib.dna_make.argtypes = [C.c_char_p]
    lib.dnah_last_error.restype = C.c_char_p
    lib.dna_free.argtypes = [C.c_void_p]

    class Kmer(C.Structure):
        _fields_ = [("length", C.c_int32), ("bit_sequence", C.c_uint64)]
    lib.kmer_make.argtypes = [C.c_char_p, C.POINTER(Kmer)]
    for s in ("ACGT", "ATCGATCGATCGATCGACG", "G" * 70):
        p = lib.dna_make(s.encode())
        assert p
        length = C.cast(p, C.POINTER(C.c_uint64))[0]
        words, n = ref.encode_dna(s)
        got = [C.cast(p, C.POINTER(C.c_uint64))[1 + i] for i in range(len(words))]
        assert length == n and got == [int(w) for w in words]
        lib.dna_free(p)
    assert not lib.dna_make(b"ACGN")
    assert b"Invalid character in DNA sequence: N" in lib.dnah_last_error()
    assert not lib.dna_make(b"")
    assert b"cannot be empty" in lib.dnah_last_error()
    for s in ("A", "ATCG", "ACGTX", "G" * 32):
        k = Kmer()
        assert lib.kmer_make(s.encode(), C.byref(k)) == 0
        assert (k.bit_sequence, k.length) == ref.kmer_make(s)
    k = Kmer()
    assert lib.kmer_make(b"A" * 33, C.byref(k)) != 0
    assert b"cannot exceed 32" in lib.dnah_last_error()
