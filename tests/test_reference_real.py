"""The oracle's restatement against the REAL reference: /root/reference/dna.c compiled unmodified
against the PostgreSQL API shim (oracle/pgshim) and driven like the executor (SRF loop, per-row
quals, hash aggregate through kmer_hash/kmer_eq).  This is what pins the oracle.  CPU only; skipped
where neither the built library nor /root/reference exists."""
import numpy as np
import pytest

from oracle import ref_cpu as R
from oracle import ref_real as P

pytestmark = pytest.mark.skipif(not P.available(), reason="oracle/_ref not built and /root/reference absent")

IUPAC = "ATCGUWSMKRYBDHVN"


def rand_words(rng, n):
    w = rng.integers(0, 2**64, size=(n + 31) // 32, dtype=np.uint64)
    if n % 32:
        w[-1] &= np.uint64((1 << (2 * (n % 32))) - 1)
    return w


def test_reference_kats_run_through_the_real_code(kats):
    for v in kats["generate_kmers"]:
        w, n = P.dna_in(v["dna"])
        assert [P.kmer_out(b, v["k"]) for b in P.generate_kmers(w, n, v["k"])] == v["rows"]
    for v in kats["starts_with"]:
        w, n = P.dna_in(v["dna"])
        rows = P.generate_kmers(w, n, v["k"], prefix=P.kmer_in(v["prefix"]))
        assert [P.kmer_out(b, v["k"]) for b in rows] == v["rows"]
    for v in kats["contains"]:
        w, n = P.dna_in(v["dna"])
        rows = P.generate_kmers(w, n, v["k"], pattern=v["pattern"])
        assert [P.kmer_out(b, v["k"]) for b in rows] == v["rows"]
    for v in kats["group_by"]:
        w, n = P.dna_in(v["dna"])
        r = P.count(w, 1, n, len(w), v["k"])
        assert {P.kmer_out(b, v["k"]): int(c) for b, c in zip(r.kmers, r.counts)} == v["counts"]
    for v in kats["stats"]:
        w, n = P.dna_in(v["dna"])
        assert P.count(w, 1, n, len(w), v["k"]).stats == (v["total"], v["distinct"], v["unique"])
    for v in kats["encoding"]:
        assert P.kmer_in(v["kmer"]) == (int(v["bits"], 16), len(v["kmer"]))


def test_codecs_agree():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 64, 200, 1001):
        s = "".join(rng.choice(list("ATCG"), size=n))
        wp, np_ = P.dna_in(s)
        wr, nr = R.encode_dna(s)
        assert np_ == nr and np.array_equal(wp, wr)
        assert P.dna_out(wp, n) == R.decode_dna(wr, n) == s
    for s in ("A", "ACGT", "ACGTX", "G" * 32, "TTTTCCCCAAAAGGGG"):
        assert P.kmer_in(s) == R.kmer_make(s)
    for bad, msg in (("", "cannot be empty"), ("ACGN", "Invalid character"), ("acgt", "Invalid character")):
        with pytest.raises(P.PgError, match=msg):
            P.dna_in(bad)
        with pytest.raises(R.RefError):
            R.encode_dna(bad)
    with pytest.raises(P.PgError, match="cannot exceed 32"):
        P.kmer_in("A" * 33)


@pytest.mark.parametrize("k", [1, 2, 5, 16, 21, 31, 32])
def test_generate_kmers_restatement_equals_reference(k):
    rng = np.random.default_rng(k)
    for n in (k, k + 1, 40, 64, 65, 333, 5000):
        if n < k:
            continue
        w = rand_words(rng, n)
        real = P.generate_kmers(w, n, k)
        assert np.array_equal(real, R.generate_kmers(w, n, k))
        assert np.array_equal(real, R.generate_kmers(w, n, k, window=True))
    with pytest.raises(P.PgError, match="Invalid k value"):
        P.generate_kmers(rand_words(rng, 100), 100, 0)
    with pytest.raises(P.PgError, match="Invalid k value"):
        P.generate_kmers(rand_words(rng, 100), 100, 33)


def test_kmer_hash_value_is_pinned_by_the_reference_call_site():
    """dna.c:732 calls hash_any on the 8 key bytes; the shim's hash_any is lookup3 as PostgreSQL
    defines it; the oracle's 8-byte restatement must agree for every key."""
    rng = np.random.default_rng(3)
    for x in [0, 1, 2**32, 2**64 - 1] + [int(v) for v in rng.integers(0, 2**63, size=2000)]:
        assert P.kmer_hash(x) == R.kmer_hash(x)
    assert P.kmer_eq(5, 3, 5, 3) and not P.kmer_eq(5, 3, 5, 4) and not P.kmer_eq(5, 3, 6, 3)


def test_predicates_restatement_equals_reference():
    rng = np.random.default_rng(4)
    for _ in range(400):
        k = int(rng.integers(1, 33))
        x = int(rng.integers(0, 2**63)) & ((1 << (2 * k)) - 1)
        pl = int(rng.integers(1, min(k, 31) + 1))
        pb = int(rng.integers(0, 2**62)) & ((1 << (2 * pl)) - 1) if rng.random() < 0.5 else x & ((1 << (2 * pl)) - 1)
        assert P.starts_with(x, k, pb, pl) == R.starts_with(x, k, pb, pl)
        pat = "".join(rng.choice(list(IUPAC), size=k))
        assert P.contains(pat, x, k) == R.contains(pat, x, k)
    with pytest.raises(P.PgError, match="Prefix length cannot exceed kmer length"):
        P.starts_with(0, 3, 0, 4)
    with pytest.raises(P.PgError, match="lengths do not match"):
        P.contains("NN", 0, 3)
    with pytest.raises(P.PgError, match="Invalid character in qkmer pattern"):
        P.contains("NXN", 0, 3)
    # Q1: a 32-base prefix makes the reference shift by 64 (UB); on x86 the mask becomes 0
    full = 2**64 - 1
    assert P.starts_with(full, 32, 0, 32) == R.starts_with(full, 32, 0, 32, x86=True)
    # Q4: U is accepted as a pattern character but matches nothing
    assert not any(P.contains("U", b, 1) for b in range(4))


@pytest.mark.parametrize("k,prefix,pattern", [(5, None, None), (21, None, None), (31, "AC", None),
                                              (6, None, "DNMSRN"), (12, "T", "NNNNWSNNNNRY"), (32, None, None)])
def test_group_by_restatement_equals_reference(k, prefix, pattern):
    n = 60_000
    w = R.synth_seq(17 + k, n)
    pk = R.kmer_make(prefix) if prefix else None
    real = P.count(w, 1, n, len(w), k, prefix=pk, pattern=pattern)
    port = R.count_query(w, 1, n, len(w), k, prefix=pk, pattern=pattern, faithful=True)
    assert real.stats == port.stats
    assert np.array_equal(real.kmers, port.kmers) and np.array_equal(real.counts, port.counts)
    assert np.array_equal(real.digest, port.digest)
    reads = R.synth_reads(5, 400, 150, 5)
    real = P.count(reads, 400, 150, 5, min(k, 31), prefix=pk if k <= 31 else None, threads=3)
    port = R.count_query(reads, 400, 150, 5, min(k, 31), prefix=pk if k <= 31 else None, faithful=False)
    assert real.stats == port.stats and np.array_equal(real.counts, port.counts)
