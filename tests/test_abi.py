"""The drop-in boundary on a box without a GPU: libdnagpu.so loads, exports every symbol
include/dnagpu.h declares, reports the reference's error texts, and refuses to run
without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "dnagpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dnagpu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_is_plain_c():
    src = os.path.join(ROOT, "tests", "_abi_probe.c")
    with open(src, "w") as f:
        f.write('#include "dnagpu.h"\n#include "dnagpu_synth.h"\nint main(void){dnagpu_where w={0,0,0,0};'
                'dnagpu_stats s={0,0,0};(void)w;(void)s;return (int)dnagpu_splitmix64(1)&0;}\n')
    try:
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), src], check=True)
    finally:
        os.unlink(src)


def test_library_exports_every_declared_symbol():
    from dnagpu import _lib
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/dnagpu.h but not exported"
    # and the Python binding describes exactly the declared surface
    assert sorted(_lib.SIGNATURES) == declared


def test_error_texts_are_the_references():
    from dnagpu import _lib
    lib = _lib.load()
    assert lib.dnagpu_version() == 100
    msgs = {1: "Invalid k value: must be between 1 and 32",            # dna.c:773
            2: "Prefix length cannot exceed kmer length",              # dna.c:855
            3: "Qkmer pattern and kmer lengths do not match",          # dna.c:1107
            5: "qkmer pattern cannot be empty",                        # dna.c:878
            6: "Qkmer pattern length cannot exceed 32 characters"}     # dna.c:884
    for code, text in msgs.items():
        assert lib.dnagpu_strerror(code).decode() == text


def test_owner_of_is_a_pure_host_function():
    from dnagpu import owner_of
    for parts in (1, 2, 3, 8, 64):
        owners = [owner_of(x * 0x9E3779B97F4A7C15 & (2**64 - 1), parts) for x in range(2000)]
        assert min(owners) >= 0 and max(owners) < parts
        if parts > 1:
            assert len(set(owners)) == parts


def test_owner_of_forms():
    """2 / 4 / 8 owners: GF(2)-linear in the k-mer's low word (kernels.cuh owner_of: parities under three tap masks), so
    owner(x ^ y) = owner(x) ^ owner(y); every other count: multiplicative hash of the low word.  Both ignore the
    high word and spread random k-mers evenly."""
    import random
    from dnagpu import owner_of
    rnd = random.Random(7)
    xs = [rnd.getrandbits(64) for _ in range(4000)]
    for parts in (2, 4, 8):
        assert owner_of(0, parts) == 0
        for x, y in zip(xs[:500], xs[500:1000]):
            assert owner_of(x ^ y, parts) == owner_of(x, parts) ^ owner_of(y, parts)
        # the ranks of 4 and 2 owners are the low bits of the rank among 8
        assert all(owner_of(x, parts) == owner_of(x, 8) & (parts - 1) for x in xs[:500])
    for parts in (2, 3, 4, 5, 8, 16):
        assert all(owner_of(x, parts) == owner_of(x & 0xFFFFFFFF, parts) for x in xs[:500])
        share = [0] * parts
        for x in xs:
            share[owner_of(x, parts)] += 1
        assert max(share) - min(share) < 0.35 * len(xs) / parts, (parts, share)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dnagpu
    with pytest.raises(dnagpu.DnaError) as e:
        dnagpu.Context(0)
    assert e.value.code == 24  # DNAGPU_ENODEVICE
    assert "no CPU path" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dna-sequences-pg-extension_b200")
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                text = open(os.path.join(d, fn), errors="replace").read()
                assert "ref_cpu" not in text and "oracle" not in text.lower().replace("test infrastructure", ""), \
                    f"{fn} refers to the oracle"
