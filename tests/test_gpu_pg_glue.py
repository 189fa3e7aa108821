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


def test_where_pushdown_rows_equal_the_reference_quals(mods):
    """generate_kmers_where / kmer_stats / count_kmers with the WHERE clause on the GPU against the reference's own
    plan: generate_kmers rows filtered by ITS starts_with / contains (test.sql:67-73, 86-92)."""
    ref, glue = mods
    rng = np.random.default_rng(31)
    words, nb = ref.dna_in(_random_dna(rng, 30_000))
    for k, prefix, pattern in ((3, "AC", None), (6, None, "DNMSRN"), (9, "G", "NNWNNSNNN"), (31, "ACG", None),
                               (32, None, "N" * 31 + "R"), (5, None, None)):
        pk = ref.kmer_in(prefix) if prefix else None
        want_rows = ref.generate_kmers(words, nb, k, prefix=pk, pattern=pattern)
        assert np.array_equal(glue.generate_kmers_where(words, nb, k, prefix=pk, pattern=pattern), want_rows)
        b = ref.count(words, 1, nb, words.size, k, prefix=pk, pattern=pattern)
        assert glue.kmer_stats(words, nb, k, prefix=pk, pattern=pattern, where_form=True) == b.stats
        kk, cc = glue.count_kmers(words, nb, k, prefix=pk, pattern=pattern)
        assert np.array_equal(kk, b.kmers) and np.array_equal(cc.astype(np.uint64), b.counts)
    # the reference's per-row ERRORs (dna.c:854-856, 1106-1108) with the reference's texts
    for fn in (glue.generate_kmers_where, glue.kmer_stats, glue.count_kmers):
        with pytest.raises(ref_real.PgError, match="Prefix length cannot exceed kmer length"):
            fn(words, nb, 3, prefix=ref.kmer_in("ACGT"))
        with pytest.raises(ref_real.PgError, match="Qkmer pattern and kmer lengths do not match"):
            fn(words, nb, 5, pattern="NNN")
    assert glue.live_tables() == 0 and glue.live_contexts() == 0


def test_generate_kmers_windows_cover_long_values(mods):
    """The drop-in extracts in windows (GLUE_WINDOW_ROWS = 4 Mi rows): a value longer than two windows, row for row
    against the oracle, with and without a pushed-down clause."""
    ref, glue = mods
    n, k = 9_000_011, 21
    words = R.synth_seq(5, n)
    assert np.array_equal(glue.generate_kmers(words, n, k), R.generate_kmers(words, n, k, window=True))
    pk = R.kmer_make("ACGTT")
    assert np.array_equal(glue.generate_kmers_where(words, n, k, prefix=pk),
                          R.filter_kmers(words, n, k, prefix=pk))
    kk, cc = glue.count_kmers(words, n, k)             # > 2 windows of grouped rows, fetched from the GPU table
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8)
    assert np.array_equal(kk, want.kmers) and np.array_equal(cc.astype(np.uint64), want.counts)
    assert glue.live_tables() == 0


def test_abandoned_scan_releases_the_gpu_table(mods):
    """LIMIT / cancel: the executor stops calling count_kmers half way and resets the SRF's memory context; the GPU
    table must go with it (MemoryContextRegisterResetCallback), not leak until the backend exits."""
    ref, glue = mods
    n, k = 6_000_000, 31
    words = R.synth_seq(9, n)
    kk, cc = glue.count_kmers(words, n, k, stop_after=10)
    assert kk.size == 10 and glue.live_tables() == 0 and glue.live_contexts() == 0
    # an ERROR raised while the table is alive (bad k on the NEXT query is not it: fail inside the scan instead)
    kk, cc = glue.count_kmers(words, n, k, stop_after=5_000_000)   # crosses a window boundary, then abandons
    assert kk.size == 5_000_000 and glue.live_tables() == 0


def test_table_form_aggregate_equals_the_reference_plan(mods):
    """SELECT (kmer_stats_agg(sequence, k)).* FROM dna_sequences  ==  the reference's table-form query
    (test.sql:140-150): k-mers never span rows, counts merge across rows, NULL rows contribute nothing."""
    ref, glue = mods
    rng = np.random.default_rng(5)
    seqs = []
    for n in (1, 9, 10, 11, 64, 1000, 4097, 33, 5000, 10):
        seqs.append(ref.dna_in(_random_dna(rng, n)))
    seqs.insert(3, None)
    seqs.append(seqs[5])          # a duplicated row: its k-mers count twice
    for k in (1, 5, 10, 32):
        want = R.count_ragged([s for s in seqs if s is not None], k, faithful=False)
        assert glue.kmer_stats_agg(seqs, k) == want.stats, k
    assert glue.kmer_stats_agg([None, None], 5) == (0, 0, 0)
    assert glue.kmer_stats_agg([], 5) == (0, 0, 0)
    with pytest.raises(ref_real.PgError, match="Invalid k value"):
        glue.kmer_stats_agg(seqs, 33)
    # the reference's own numbers shape: 1 M random nt in rows of 100 k, k = 10 (test.sql:151-154: ~ 644 k distinct)
    rows = [ref.dna_in(_random_dna(rng, 100_000)) for _ in range(10)]
    total, distinct, unique = glue.kmer_stats_agg(rows, 10)
    assert total == 10 * (100_000 - 9) and 640_000 < distinct < 650_000 and 380_000 < unique < 390_000


_ALL_GPUS_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import ref_cpu as R
from oracle import ref_real
glue = ref_real.glue()
assert glue.device_count() == 0                      # the backend has not touched a dna value yet
rng = np.random.default_rng(5)
for n, k in ((3_000_000, 31), (3_000_000, 21), (500_000, 13), (70_000, 8), (40, 32)):
    text = "".join(rng.choice(list("ATCG"), size=n)) if n != 40 else "G" * 40
    words, nb = glue.dna_in(text)
    want = R.count_query(words, 1, nb, words.size, k, faithful=False, threads=8)
    assert glue.kmer_stats(words, nb, k) == (want.total, want.distinct, want.unique), (n, k)
    kk, cc = glue.count_kmers(words, nb, k)
    assert np.array_equal(kk, want.kmers) and np.array_equal(cc.astype(np.uint64), want.counts), (n, k)
    # a WHERE clause and the SRF run on the first GPU of the same context
    assert np.array_equal(glue.generate_kmers(words, nb, k), R.generate_kmers(words, nb, k))
    got = glue.generate_kmers(words, nb, k, prefix=(2, 1), pattern="N" * k)       # WHERE kmer ^@ 'C' AND 'NN..N' @> kmer
    assert np.array_equal(got, R.filter_kmers(words, nb, k, prefix=(2, 1), pattern="N" * k))
assert glue.device_count() == int(sys.argv[2]), glue.device_count()
assert glue.live_tables() == 0
print("glue on", glue.device_count(), "GPUs ok")
"""


def test_glue_backend_context_over_all_gpus(gpu):
    """DNAGPU_DEVICES (the stand-in for the GUC dna_gpu.devices) = every GPU of the box: the backend's context is a
    multi-GPU one (dnagpu_create_multi) and kmer_stats / count_kmers shard the sequence over the GPUs -- from C, through
    the fmgr surface, no Python orchestration.  Own process: a backend creates its context once."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DNAGPU_DEVICES=",".join(str(i) for i in range(n)))
    p = subprocess.run([sys.executable, "-c", _ALL_GPUS_SCRIPT, root, str(n)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"glue on {n} GPUs ok" in p.stdout
