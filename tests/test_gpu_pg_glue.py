"""The drop-in, run: the reference's dna.c with its generate_kmers swapped for the GPU glue (pg/dna_gpu.c ->
libdnagpu), driven through the fmgr / SRF protocol by the same executor-like driver that drives the unmodified
reference (oracle/pgshim/driver.c).  Row for row and error for error the two modules must agree; the quals
(starts_with, contains) and the HashAggregate functions (kmer_hash, kmer_eq) are the reference's own in both.

Both libraries are built in the container (where /root/reference exists) into oracle/_ref/ and travel with the
snapshot; nothing here reads /root/reference at run time."""
import numpy as np
import pytest

from oracle import ref_cpu as R
from oracle import ref_real

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref_real.available() and ref_real.glue_available()),
                                 reason="oracle/_ref/libdnaref.so / libdnaglue.so not built")]


@pytest.fixture(scope="module")
def mods(gpu):
    return ref_real.reference(), ref_real.glue()


def _random_dna(rng, n):
    return "".join(rng.choice(list("ATCG"), size=n))


def test_reference_kats_through_the_glue(mods):
    _, glue = mods
    def rows(seq, k, prefix=None, pattern=None):
        words, n = glue.dna_in(seq)
        got = glue.generate_kmers(words, n, k, prefix=glue.kmer_in(prefix) if prefix else None, pattern=pattern)
        return [glue.kmer_out(b, k) for b in got]
    assert rows("ATCGTAGCGT", 3) == ["ATC", "TCG", "CGT", "GTA", "TAG", "AGC", "GCG", "CGT"]      # test.sql:46-58
    assert rows("ACTGACGTACC", 3, prefix="AC") == ["ACT", "ACG", "ACC"]                            # test.sql:67-73
    assert rows("ACGTACGCACGT", 6, pattern="DNMSRN") == ["GTACGC", "GCACGT"]                       # test.sql:86-92
    words, n = glue.dna_in("ATCGATCGATCGATCGACG")                                                    # test.sql:95-104
    kk, cc = glue.count_kmers(words, n, 5)
    assert {glue.kmer_out(k, 5): int(c) for k, c in zip(kk, cc)} == {"ATCGA": 4, "CGATC": 3, "GATCG": 3, "TCGAT": 3,
                                                                      "TCGAC": 1, "CGACG": 1}
    words, n = glue.dna_in("ACGTACGTACGTAG")
    assert glue.kmer_stats(words, n, 8) == (7, 5, 3)                                                 # test.sql:107-119
    assert glue.kmer_stats(words, n, 5) == (10, 5, 1)                                                # README.md:122-134


def test_generate_kmers_rows_equal_the_reference_module(mods):
    ref, glue = mods
    rng = np.random.default_rng(2024)
    for n in (1, 2, 31, 32, 33, 64, 65, 100, 1000, 4097, 50_001):
        words, nb = ref.dna_in(_random_dna(rng, n))
        for k in sorted({1, 2, 5, 16, 31, 32, min(n, 32), min(n + 1, 32)}):
            if nb + 1 < k:           # the reference wraps around below k - 1 bases (dna.c:781): not called there
                assert glue.generate_kmers(words, nb, k).size == 0
                continue
            assert np.array_equal(glue.generate_kmers(words, nb, k), ref.generate_kmers(words, nb, k)), (n, k)


def test_quals_and_hash_aggregate_over_the_glue_rows(mods):
    """WHERE ^@ / @> evaluated by the reference's operators on the rows the GPU produced, and GROUP BY through
    kmer_hash / kmer_eq: the plans of test.sql:67-92 and 95-119 with only generate_kmers replaced."""
    ref, glue = mods
    rng = np.random.default_rng(7)
    words, nb = ref.dna_in(_random_dna(rng, 20_000))
    for k, prefix, pattern in ((5, "AC", None), (9, None, "NNWNNSNNN"), (12, "G", "N" * 11 + "Y"), (31, None, None)):
        pk = ref.kmer_in(prefix) if prefix else None
        assert np.array_equal(glue.generate_kmers(words, nb, k, prefix=pk, pattern=pattern),
                              ref.generate_kmers(words, nb, k, prefix=pk, pattern=pattern))
        a = glue.count(words, 1, nb, words.size, k, prefix=pk, pattern=pattern)
        b = ref.count(words, 1, nb, words.size, k, prefix=pk, pattern=pattern)
        assert a.stats == b.stats and np.array_equal(a.kmers, b.kmers) and np.array_equal(a.counts, b.counts)


def test_pushdown_functions_equal_the_reference_group_by(mods):
    ref, glue = mods
    rng = np.random.default_rng(99)
    for n, k in ((300, 3), (5000, 6), (70_000, 8), (70_000, 21), (200_000, 32), (40, 32), (5, 8)):
        words, nb = ref.dna_in(_random_dna(rng, n) if n != 40 else "G" * 40)
        want = R.count_query(words, 1, nb, words.size, k, faithful=False)
        if nb + 1 >= k:              # the reference's own plan, where it is defined
            b = ref.count(words, 1, nb, words.size, k)
            assert b.stats == (want.total, want.distinct, want.unique)
        assert glue.kmer_stats(words, nb, k) == (want.total, want.distinct, want.unique), (n, k)
        kk, cc = glue.count_kmers(words, nb, k)
        assert np.array_equal(kk, want.kmers) and np.array_equal(cc.astype(np.uint64), want.counts), (n, k)


def test_errors_carry_the_reference_text(mods):
    ref, glue = mods
    words, nb = ref.dna_in("ACGTACGT")
    for k in (0, -3, 33):
        with pytest.raises(ref_real.PgError) as e_ref:
            ref.generate_kmers(words, nb, k)
        for fn in (glue.generate_kmers, glue.kmer_stats, glue.count_kmers):
            with pytest.raises(ref_real.PgError) as e_glue:
                fn(words, nb, k)
            assert str(e_glue.value) == str(e_ref.value) == "Invalid k value: must be between 1 and 32"
    # the module keeps working after an ereport
    assert glue.generate_kmers(words, nb, 6).size == 3
