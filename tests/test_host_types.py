"""Host-side dna / kmer / qkmer values (scalar glue that stays on the CPU in the
reference too) against the oracle's restatement of dna_make / kmer_make / qkmer_make."""
import numpy as np
import pytest

import dnagpu
from dnagpu import Dna, DnaError, Kmer, Qkmer


def test_dna_roundtrip_and_layout(ref):
    rng = np.random.default_rng(0)
    for n in (1, 4, 31, 32, 33, 64, 65, 1000):
        s = "".join(rng.choice(list("ATCG"), size=n))
        d = Dna(s)
        words, length = ref.encode_dna(s)
        assert d.length == length == n
        assert np.array_equal(d.words, words)
        assert str(d) == s == ref.decode_dna(words, n)
        assert d == Dna(s) and len(d) == n


def test_dna_errors():
    with pytest.raises(DnaError, match="cannot be empty"):
        Dna("")
    with pytest.raises(DnaError, match="Invalid character in DNA sequence: N"):
        Dna("ACGTN")
    with pytest.raises(DnaError):
        Dna("acgt")  # lowercase is rejected (dna.c:165)


def test_kmer_matches_oracle(ref):
    for s in ("A", "ATCG", "ACGT", "G" * 32, "ACGTX", "TTTTTTTTTTTTTTTTTTTTT"):
        k = Kmer(s)
        bits, length = ref.kmer_make(s)
        assert (k.bits, k.length) == (bits, length)
        assert str(k) == s.replace("X", "A")
    with pytest.raises(DnaError, match="cannot be empty"):
        Kmer("")
    with pytest.raises(DnaError, match="cannot exceed 32"):
        Kmer("A" * 33)
    with pytest.raises(DnaError, match="Invalid character"):
        Kmer("ACGU")


def test_kmer_scalar_starts_with(ref):
    assert Kmer("ACGT").starts_with(Kmer("AC"))
    assert not Kmer("ACGT").starts_with(Kmer("CA"))
    with pytest.raises(DnaError, match="Prefix length"):
        Kmer("AC").starts_with(Kmer("ACG"))
    g = Kmer("G" * 32)
    assert g.starts_with(g) and not g.starts_with(Kmer("A" * 32))


def test_qkmer_validation(kats):
    assert str(Qkmer(kats["qkmer_alphabet"]["pattern"])) == "ATCGUWSMKRYBDHVN"
    with pytest.raises(DnaError, match="cannot be empty"):
        Qkmer("")
    with pytest.raises(DnaError, match="cannot exceed 32"):
        Qkmer("N" * 33)
    with pytest.raises(DnaError, match="Invalid character in qkmer pattern: X"):
        Qkmer("ANX")


def test_kmer_strings():
    bits = np.array([Kmer("ATCG").bits, Kmer("GGGA").bits], dtype=np.uint64)
    assert dnagpu.kmer_strings(bits, 4) == ["ATCG", "GGGA"]
