"""Properties of the oracle itself: the faithful per-k-mer restatement of
generate_kmers (dna.c:803-825) equals the two-word window form for every k, the
predicates obey their definitions, the synthetic generator is position-addressable."""
import numpy as np
import pytest

from oracle import ref_cpu as R


@pytest.mark.parametrize("k", [1, 2, 3, 5, 16, 21, 31, 32])
def test_window_equals_faithful(k):
    rng = np.random.default_rng(k)
    for n in (k, k + 1, 31, 32, 33, 63, 64, 65, 1000, 4097):
        if n < k:
            continue
        words = rng.integers(0, 2**64, size=(n + 31) // 32, dtype=np.uint64)
        if n % 32:
            words[-1] &= np.uint64((1 << (2 * (n % 32))) - 1)
        a = R.generate_kmers(words, n, k)
        b = R.generate_kmers(words, n, k, window=True)
        assert a.size == n - k + 1
        assert np.array_equal(a, b)


def test_rows_when_shorter_than_k():
    words = np.zeros(1, dtype=np.uint64)
    assert R.generate_kmers(words, 4, 5).size == 0       # Q2: no unsigned wrap
    assert R.generate_kmers(words, 5, 5).size == 1
    with pytest.raises(R.RefError) as e:
        R.generate_kmers(words, 10, 0)
    assert e.value.code == 6
    with pytest.raises(R.RefError):
        R.generate_kmers(words, 10, 33)


def test_kmer_hash_is_lookup3():
    # hash_any over 8 bytes with both halves zero except a: equals hash_uint32-style mixing
    # of PG's hash_bytes for len=8; fixed regression values of this restatement
    assert R.kmer_hash(0) == R.kmer_hash(0)
    vals = {R.kmer_hash(x) for x in range(1000)}
    assert len(vals) == 1000  # no collisions on a tiny dense range
    assert all(0 <= v < 2**32 for v in vals)


def test_starts_with_definition():
    rng = np.random.default_rng(1)
    for _ in range(200):
        k = int(rng.integers(1, 33))
        x = int(rng.integers(0, 2**63)) & ((1 << (2 * k)) - 1)
        pl = int(rng.integers(1, k + 1))
        pref = x & ((1 << (2 * pl)) - 1)
        assert R.starts_with(x, k, pref, pl)
        other = pref ^ 1
        assert not R.starts_with(x, k, other, pl)
    with pytest.raises(R.RefError) as e:
        R.starts_with(0, 3, 0, 4)
    assert e.value.code == 7
    full = (1 << 64) - 1
    assert R.starts_with(full, 32, full, 32)
    assert not R.starts_with(full, 32, 0, 32)
    assert R.starts_with(full, 32, 0, 32, x86=True)  # Q1: what the reference's UB does on x86


IUPAC = {"A": "A", "T": "T", "C": "C", "G": "G", "U": "", "W": "AT", "S": "CG", "M": "AC", "K": "GT",
         "R": "AG", "Y": "CT", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}


def test_nucleotide_matches_table():
    for code, allowed in IUPAC.items():
        for nt in "ATCG":
            assert bool(R.lib().ref_nucleotide_matches(nt.encode(), code.encode())) == (nt in allowed)


def test_contains_definition():
    rng = np.random.default_rng(2)
    codes = list(IUPAC)
    for _ in range(300):
        k = int(rng.integers(1, 33))
        pat = "".join(rng.choice(codes, size=k))
        x = int(rng.integers(0, 2**63)) & ((1 << (2 * k)) - 1)
        s = R.decode_kmer(x, k)
        want = all(ch in IUPAC[p] for ch, p in zip(s, pat))
        assert R.contains(pat, x, k) == want
    with pytest.raises(R.RefError) as e:
        R.contains("NN", 0, 3)
    assert e.value.code == 8


def test_mt_equals_single_thread():
    words = R.synth_seq(11, 200_000)
    a = R.count_query(words, 1, 200_000, words.size, 11, faithful=True)
    b = R.count_query(words, 1, 200_000, words.size, 11, faithful=False, threads=4)
    assert a.stats == b.stats and np.array_equal(a.kmers, b.kmers) and np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.digest, b.digest)
    assert np.array_equal(R.pairs_digest(a.kmers, a.counts), a.digest)
    reads = R.synth_reads(3, 5000, 150, 5)
    a = R.count_query(reads, 5000, 150, 5, 31, faithful=True)
    b = R.count_query(reads, 5000, 150, 5, 31, faithful=False, threads=3)
    assert a.stats == b.stats and np.array_equal(a.counts, b.counts)
    assert a.total == 5000 * 120


def test_synth_is_position_addressable_and_plants_repeats():
    n = 100_007
    full = R.synth_seq(5, n)
    part = R.synth_seq(5, n, first_word=1000, n_words=500)
    assert np.array_equal(full[1000:1500], part)
    assert full[-1] >> np.uint64(2 * (n % 32)) == 0          # tail bits zero (dna.c:186)
    assert full[8] == np.uint64(2**64 - 1) and full[16] == 0    # planted G x 64 / A x 64 windows
    r = R.count_query(full, 1, n, full.size, 31, faithful=False, want_rows=False)
    assert r.distinct < r.total and r.unique < r.distinct      # counts > 1 exist at k = 31
    raw = R.synth_seq(5, n, repeat_every=0)
    r0 = R.count_query(raw, 1, n, raw.size, 31, faithful=False, want_rows=False)
    assert r0.distinct > r.distinct
