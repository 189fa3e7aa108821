"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Bit-exact: every
value on this path is an integer.  Mirrors the statements of the reference's test.sql."""
import numpy as np
import pytest

import dnagpu
from dnagpu import Dna, DnaError, Kmer
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu

IUPAC = "ATCGUWSMKRYBDHVN"


def rand_dna_words(rng, n):
    words = rng.integers(0, 2**64, size=(n + 31) // 32, dtype=np.uint64)
    if n % 32:
        words[-1] &= np.uint64((1 << (2 * (n % 32))) - 1)
    return words


# ---- the reference's own statements (test.sql:46-119, README.md:66-135) -----------------
def test_kat_generate_kmers(gpu, kats):
    for v in kats["generate_kmers"]:
        assert gpu.generate_kmers(v["dna"], v["k"]).strings() == v["rows"], v["source"]


def test_kat_starts_with(gpu, kats):
    for v in kats["starts_with"]:
        assert gpu.filter_kmers(v["dna"], v["k"], prefix=v["prefix"]).strings() == v["rows"], v["source"]


def test_kat_contains(gpu, kats):
    for v in kats["contains"]:
        assert gpu.filter_kmers(v["dna"], v["k"], pattern=v["pattern"]).strings() == v["rows"], v["source"]


def test_kat_equality_filter(gpu, kats):
    for v in kats["equality_filter"]:  # kmer = 'ACGTAC'  <=>  prefix of full length
        assert gpu.filter_kmers(v["dna"], v["k"], prefix=v["equals"]).strings() == v["rows"], v["source"]


def test_kat_group_by(gpu, kats):
    for v in kats["group_by"]:
        assert dnagpu.count_kmers(v["dna"], v["k"], ctx=gpu) == v["counts"], v["source"]


def test_kat_stats(gpu, kats):
    for v in kats["stats"]:
        assert dnagpu.kmer_stats(v["dna"], v["k"], ctx=gpu) == (v["total"], v["distinct"], v["unique"]), v["source"]


# ---- generate_kmers ----------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 13, 16, 17, 21, 31, 32])
def test_extract_matches_oracle(gpu, k):
    rng = np.random.default_rng(100 + k)
    for n in (k, k + 1, 32, 33, 64, 65, 1000, 2049, 70_001):
        if n < k:
            continue
        words = rand_dna_words(rng, n)
        want = R.generate_kmers(words, n, k)            # faithful decode/re-encode form
        got = gpu.generate_kmers(Dna.from_words(words, n), k).bits
        assert np.array_equal(got, want), (k, n)


def test_extract_edge_cases(gpu):
    d = Dna("ACGT")
    assert len(gpu.generate_kmers(d, 5)) == 0                       # Q2: length < k -> 0 rows
    assert gpu.generate_kmers(d, 4).strings() == ["ACGT"]
    for k in (0, -1, 33):
        with pytest.raises(DnaError, match="Invalid k value: must be between 1 and 32") as e:
            gpu.generate_kmers(d, k)
        assert e.value.code == 1
    g = Dna("G" * 40)
    assert set(gpu.generate_kmers(g, 32).bits.tolist()) == {2**64 - 1}
    a = Dna("A" * 40)
    assert set(gpu.generate_kmers(a, 32).bits.tolist()) == {0}


def test_extract_device_path_and_synth(gpu):
    import torch
    n = 1_000_003
    seq = gpu.synth(n, seed=9)
    words = R.synth_seq(9, n)
    assert np.array_equal(seq.download(), words)                 # device generator == host generator
    for k in (21, 32):
        out = gpu.extract(seq, k)
        gpu.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        assert np.array_equal(got, R.generate_kmers(words, n, k, window=True))
    seq.free()


def test_extract_reads_and_ragged(gpu):
    rng = np.random.default_rng(5)
    reads = R.synth_reads(3, 1000, 150, 5)
    for k in (31, 30, 7):
        seq = gpu.upload_reads(reads, 1000, 150, 5)
        got = gpu.extract(seq, k).cpu().numpy().view(np.uint64)
        want = np.concatenate([R.generate_kmers(reads[r * 5:(r + 1) * 5], 150, k) for r in range(1000)])
        assert np.array_equal(got, want), k
        seq.free()
    lens = [1, 5, 31, 32, 33, 100, 7, 64, 2000, 3]
    dnas = [Dna.from_words(rand_dna_words(rng, n), n) for n in lens]
    seq = gpu.upload_ragged(dnas)
    for k in (1, 5, 32):
        got = gpu.extract(seq, k).cpu().numpy().view(np.uint64)
        parts = [R.generate_kmers(d.words, d.length, k) for d in dnas]
        assert np.array_equal(got, np.concatenate(parts)), k
        assert seq.kmer_count(k) == sum(p.size for p in parts)
    seq.free()


# ---- WHERE ^@ / @> -------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 3, 6, 16, 21, 31, 32])
def test_filter_matches_oracle(gpu, k):
    rng = np.random.default_rng(200 + k)
    n = 20_011
    words = rand_dna_words(rng, n)
    d = Dna.from_words(words, n)
    rows = R.generate_kmers(words, n, k)
    for trial in range(6):
        plen = int(rng.integers(0, min(k, 4) + 1))
        prefix = None
        if plen:
            src = int(rows[int(rng.integers(0, rows.size))])
            prefix = Kmer(bits=src & ((1 << (2 * plen)) - 1), length=plen)
        pattern = None
        if trial % 2 == 0:
            # mostly N with a few degenerate / exact positions so that some rows survive
            pat = ["N"] * k
            for pos in rng.choice(k, size=min(k, 3), replace=False):
                pat[pos] = str(rng.choice(list(IUPAC)))
            pattern = "".join(pat)
        want = R.filter_kmers(words, n, k, prefix=(prefix.bits, prefix.length) if prefix else None,
                              pattern=pattern)
        got = gpu.filter_kmers(d, k, prefix=prefix, pattern=pattern).bits
        assert np.array_equal(got, want), (k, prefix, pattern)


def test_filter_every_iupac_code_and_u(gpu):
    n = 5000
    words = rand_dna_words(np.random.default_rng(1), n)
    d = Dna.from_words(words, n)
    for code in IUPAC:
        for pos in (0, 2, 4):
            pat = "".join(code if i == pos else "N" for i in range(5))
            want = R.filter_kmers(words, n, 5, pattern=pat)
            got = gpu.filter_kmers(d, 5, pattern=pat).bits
            assert np.array_equal(got, want), pat
            if code == "U":
                assert got.size == 0        # Q4: U matches nothing (dna.c:1070)


def test_filter_errors_are_the_references(gpu):
    d = Dna("ACGTACGTAC")
    with pytest.raises(DnaError, match="Prefix length cannot exceed kmer length") as e:
        gpu.filter_kmers(d, 3, prefix="ACGT")
    assert e.value.code == 2
    with pytest.raises(DnaError, match="Qkmer pattern and kmer lengths do not match") as e:
        gpu.filter_kmers(d, 3, pattern="NN")
    assert e.value.code == 3
    with pytest.raises(DnaError, match="Invalid character in qkmer pattern: X") as e:
        gpu.filter_kmers(d, 3, pattern="NXN")
    assert e.value.code == 4
    with pytest.raises(DnaError, match="qkmer pattern cannot be empty"):
        gpu.filter_kmers(d, 3, pattern="")
    with pytest.raises(DnaError, match="cannot exceed 32 characters"):
        gpu.filter_kmers(d, 3, pattern="N" * 33)
    # no rows evaluated -> the per-row ERRORs never fire (dna.c:854, 1106 run per row)
    assert len(gpu.filter_kmers(Dna("AC"), 3, prefix="ACGT")) == 0


def test_prefix_of_32_bases_uses_the_full_mask(gpu):
    s = "ACGT" * 8 + "GGGG" + "ACGT" * 8
    d = Dna(s)
    got = gpu.filter_kmers(d, 32, prefix="ACGT" * 8).strings()       # Q1
    assert got == ["ACGT" * 8] * 2


def test_filter_keys_over_a_kmer_column(gpu):
    import torch
    rng = np.random.default_rng(3)
    k = 5
    col = rng.integers(0, 4**k, size=100_003, dtype=np.uint64)    # the 1 M-row kmer table of test.sql:168-179
    dev = torch.from_numpy(col.view(np.int64)).cuda()
    for prefix, pattern in (("ACTG", None), (None, "MRKYN"), ("AT", "NNSNN")):   # test.sql:223, 253
        got = gpu.filter_keys(dev, k, prefix=prefix, pattern=pattern).cpu().numpy().view(np.uint64)
        keep = np.ones(col.size, dtype=bool)
        if prefix:
            pb, pl = R.kmer_make(prefix)
            keep &= (col & np.uint64((1 << (2 * pl)) - 1)) == np.uint64(pb)
        want = np.array([x for x in col[keep] if pattern is None or R.contains(pattern, int(x), k)], dtype=np.uint64)
        assert np.array_equal(got, want), (prefix, pattern)


# ---- GROUP BY kmer --------------------------------------------------------------------------------
def _check_count(gpu, seq, oracle, k, **kw):
    st, table = gpu.count(seq, k, table=True, **kw)
    assert (st.total, st.distinct, st.unique) == oracle.stats, (k, kw)
    kmers, counts = table.sorted()
    assert table.rows == oracle.distinct
    assert np.array_equal(kmers, oracle.kmers) and np.array_equal(counts, oracle.counts), (k, kw)
    table.free()


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 16, 17, 21, 31, 32])
def test_count_matches_oracle_all_methods(gpu, k):
    n = 150_001
    words = R.synth_seq(40 + k, n)                  # planted repeats + G x 64 / A x 64 windows
    seq = gpu.upload(Dna.from_words(words, n))
    oracle = R.count_query(words, 1, n, words.size, k, faithful=(k in (3, 21, 32)))
    _check_count(gpu, seq, oracle, k)                                  # AUTO
    _check_count(gpu, seq, oracle, k, method=dnagpu.COUNT_HASH)
    _check_count(gpu, seq, oracle, k, method=dnagpu.COUNT_PARTITION)
    if k <= 12:
        _check_count(gpu, seq, oracle, k, method=dnagpu.COUNT_DENSE)
    seq.free()


@pytest.mark.parametrize("k,n", [(31, 9_000_011), (32, 5_000_000), (21, 9_000_011), (14, 9_000_011)])
def test_count_partition_two_levels(gpu, k, n):
    """Inputs large enough that the radix partition needs both levels (> 4 M keys), rows included."""
    words = R.synth_seq(90 + k, n)
    seq = gpu.upload(Dna.from_words(words, n))
    oracle = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8)
    _check_count(gpu, seq, oracle, k)                                   # AUTO -> partition
    _check_count(gpu, seq, oracle, k, method=dnagpu.COUNT_PARTITION)
    filtered = R.count_query(words, 1, n, words.size, k, prefix=R.kmer_make("AC"), faithful=False, threads=8)
    _check_count(gpu, seq, filtered, k, method=dnagpu.COUNT_PARTITION, prefix="AC")
    seq.free()


def test_count_partition_skewed_input_spills_correctly(gpu):
    """Low-complexity data: a few k-mers with huge counts plus a bucket with more distinct keys
    than the shared-memory table holds must still give exact rows."""
    import torch
    rng = np.random.default_rng(12)
    k = 31
    hot = rng.integers(0, 2**62, size=5, dtype=np.uint64)
    keys = np.concatenate([np.repeat(hot, 400_000), rng.integers(0, 2**62, size=3_000_000, dtype=np.uint64)])
    rng.shuffle(keys)
    dev = torch.from_numpy(keys.view(np.int64)).cuda()
    uk, uc = np.unique(keys, return_counts=True)
    for method in (dnagpu.COUNT_PARTITION, dnagpu.COUNT_HASH):
        st, table = gpu.count_keys(dev, k, table=True, method=method)
        kk, cc = table.sorted()
        assert (st.total, st.distinct, st.unique) == (keys.size, uk.size, int((uc == 1).sum()))
        assert np.array_equal(kk, uk) and np.array_equal(cc, uc.astype(np.uint64))
    # poly-A: one k-mer, millions of copies
    seq = gpu.upload(Dna("A" * 3_000_000))
    st, table = gpu.count(seq, k, table=True, method=dnagpu.COUNT_PARTITION)
    assert (st.total, st.distinct, st.unique) == (3_000_000 - 30, 1, 0)
    assert table.fetch()[1].tolist() == [3_000_000 - 30]
    seq.free()


def test_count_partition_every_multiplicity_around_the_bin_limits(gpu):
    """The bin / place / compare bucket count holds up to 16 copies of a key in a bin and passes buckets with more
    (or with more keys than its staging area) on to the table kernel: multiplicities on both sides of the limits,
    mixed into ordinary keys, must give the exact aggregates and rows."""
    import torch
    rng = np.random.default_rng(21)
    k = 31
    parts = [rng.integers(0, 2**62, size=2_600_000, dtype=np.uint64)]
    for copies, n_keys in ((2, 200_000), (3, 50_000), (7, 20_000), (15, 5_000), (16, 5_000), (17, 5_000), (18, 2_000),
                           (40, 2_000), (300, 300), (2000, 40), (3072, 3), (3073, 3), (5000, 5)):
        parts.append(np.repeat(rng.integers(0, 2**62, size=n_keys, dtype=np.uint64), copies))
    keys = np.concatenate(parts)
    rng.shuffle(keys)
    dev = torch.from_numpy(keys.view(np.int64)).cuda()
    uk, uc = np.unique(keys, return_counts=True)
    st, table = gpu.count_keys(dev, k, table=True, method=dnagpu.COUNT_PARTITION)
    assert (st.total, st.distinct, st.unique) == (keys.size, uk.size, int((uc == 1).sum()))
    kk, cc = table.sorted()
    assert np.array_equal(kk, uk) and np.array_equal(cc, uc.astype(np.uint64))
    st, _ = gpu.count_keys(dev, k, method=dnagpu.COUNT_PARTITION)      # aggregates only
    assert (st.total, st.distinct, st.unique) == (keys.size, uk.size, int((uc == 1).sum()))


def test_host_buffer_call_pipelined_upload_and_overflow_fallback(gpu):
    """dnagpu_count_kmers on host words: chunked upload overlapped with the optimistic level 1 (no histogram
    pass, fixed-capacity regions).  Heavily repeated input overflows a region and must fall back to the exact
    two-pass level 1 with identical results."""
    n, k = 6_000_003, 31
    words = R.synth_seq(55, n)                                   # ordinary data: optimistic path holds
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8)
    st, table = gpu.count_kmers(Dna.from_words(words, n), k)
    kk, cc = table.sorted()
    assert (st.total, st.distinct, st.unique) == want.stats
    assert np.array_equal(kk, want.kmers) and np.array_equal(cc, want.counts)
    mixed = words.copy()                                          # a third of it becomes poly-A / poly-G
    mixed[10_000:70_000] = 0
    mixed[100_000:130_000] = np.uint64(2**64 - 1)
    want = R.count_query(mixed, 1, n, mixed.size, k, faithful=False, threads=8)
    for kq in (k, 32):
        want = R.count_query(mixed, 1, n, mixed.size, kq, faithful=False, threads=8)
        st, table = gpu.count_kmers(Dna.from_words(mixed, n), kq)
        kk, cc = table.sorted()
        assert (st.total, st.distinct, st.unique) == want.stats, kq
        assert np.array_equal(kk, want.kmers) and np.array_equal(cc, want.counts), kq
    st, table = gpu.count_kmers(Dna("A" * 5_000_000), k)
    assert (st.total, st.distinct, st.unique) == (5_000_000 - 30, 1, 0)
    assert table.fetch()[1].tolist() == [5_000_000 - 30]


def test_count_k32_all_g_sentinel(gpu):
    """'G' x 32 has the bit pattern of the hash table's EMPTY marker: it must still be counted."""
    for s, want_g in (("G" * 32, 1), ("G" * 40 + "ACGT" * 20 + "G" * 33, 11), ("ACGT" * 20, 0)):
        d = Dna(s)
        words, n = R.encode_dna(s)
        oracle = R.count_query(words, 1, n, words.size, 32)
        st, table = gpu.count_kmers(d, 32)
        assert (st.total, st.distinct, st.unique) == oracle.stats
        kmers, counts = table.sorted()
        assert np.array_equal(kmers, oracle.kmers) and np.array_equal(counts, oracle.counts)
        got_g = int(counts[kmers == np.uint64(2**64 - 1)].sum())
        assert got_g == want_g


def test_count_with_where_clause_fused(gpu):
    n = 300_000
    words = R.synth_seq(77, n)
    seq = gpu.upload(Dna.from_words(words, n))
    cases = [(5, "AC", None), (5, None, "MRKYN"), (21, "AC", None), (21, "A", "NNNNNNNNWSNNNNNNNNNRY"),
             (31, "AC", "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY"), (12, None, "SNNNNNNNNNNW"), (31, "T", "U" + "N" * 30)]
    for k, prefix, pattern in cases:
        pk = R.kmer_make(prefix) if prefix else None
        oracle = R.count_query(words, 1, n, words.size, k, prefix=pk, pattern=pattern, faithful=False)
        _check_count(gpu, seq, oracle, k, prefix=prefix, pattern=pattern)
    seq.free()


def test_count_reads_never_span_rows(gpu):
    n_reads, bpr, stride = 20_000, 150, 5
    reads = R.synth_reads(3, n_reads, bpr, stride)
    seq = gpu.upload_reads(reads, n_reads, bpr, stride)
    for k, prefix, pattern in ((31, None, None), (31, "AC", None), (12, None, None), (4, None, None),
                               (31, "AC", "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY")):
        pk = R.kmer_make(prefix) if prefix else None
        oracle = R.count_query(reads, n_reads, bpr, stride, k, prefix=pk, pattern=pattern, faithful=False)
        if prefix is None and pattern is None:
            assert oracle.total == n_reads * (bpr - k + 1)
        _check_count(gpu, seq, oracle, k, prefix=prefix, pattern=pattern)
    seq.free()
    st, table = gpu.count_reads(reads, n_reads, bpr, stride, 31)     # the host-buffer C-ABI call
    oracle = R.count_query(reads, n_reads, bpr, stride, 31, faithful=False)
    assert (st.total, st.distinct, st.unique) == oracle.stats


def test_count_reads_host_call_with_where_clause_pipelined(gpu):
    """dnagpu_count_reads on host words with a WHERE clause: chunked upload overlapped with the predicate scan
    (needs >= 2^24 rows).  A selective clause is counted from the collected list; one that keeps more than 1/8
    of the rows overflows the list and must fall back to the ordinary path with identical results."""
    n_reads, bpr, stride, k = 150_000, 150, 5, 31
    reads = R.synth_reads(8, n_reads, bpr, stride)
    for prefix, pattern in (("AC", "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY"), ("AC", None), ("A", None),
                            (None, "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNU")):
        pk = R.kmer_make(prefix) if prefix else None
        oracle = R.count_query(reads, n_reads, bpr, stride, k, prefix=pk, pattern=pattern, faithful=False, threads=8)
        st, table = gpu.count_reads(reads, n_reads, bpr, stride, k, prefix=prefix, pattern=pattern, table=True)
        assert (st.total, st.distinct, st.unique) == oracle.stats, (prefix, pattern)
        kk, cc = table.sorted()
        assert np.array_equal(kk, oracle.kmers) and np.array_equal(cc, oracle.counts), (prefix, pattern)
    with pytest.raises(DnaError, match="Prefix length cannot exceed kmer length"):
        gpu.count_reads(reads, n_reads, bpr, stride, 3, prefix="ACGT")


def test_count_ragged_table_of_sequences(gpu):
    """SELECT ... FROM dna_sequences d, generate_kmers(d.sequence, k) GROUP BY kmer (test.sql:140-150)."""
    rng = np.random.default_rng(8)
    lens = [int(x) for x in rng.integers(1, 400, size=300)] + [3, 9, 10, 11, 5000]
    dnas = [Dna.from_words(rand_dna_words(rng, n), n) for n in lens]
    seq = gpu.upload_ragged(dnas)
    for k in (10, 3, 32):
        oracle = R.count_ragged([(d.words, d.length) for d in dnas], k, faithful=False)
        _check_count(gpu, seq, oracle, k)
        _check_count(gpu, seq, oracle, k, method=dnagpu.COUNT_HASH)
    seq.free()


def test_count_empty_and_tiny_inputs(gpu):
    st, table = gpu.count_kmers(Dna("ACG"), 5)
    assert (st.total, st.distinct, st.unique) == (0, 0, 0) and table.rows == 0
    st, table = gpu.count_kmers(Dna("A"), 1)
    assert (st.total, st.distinct, st.unique) == (1, 1, 1)
    assert table.fetch()[0].tolist() == [0] and table.fetch()[1].tolist() == [1]
    st, _ = gpu.count_kmers(Dna("ACGTACGT"), 4, prefix="GG", table=False)     # WHERE keeps nothing
    assert (st.total, st.distinct, st.unique) == (0, 0, 0)


def test_count_keys_and_partition_compose_to_the_single_gpu_answer(gpu):
    """Owner routing (the multi-GPU GROUP BY) on one device: bucket, count each bucket,
    add the aggregates -- must equal the direct count."""
    import torch
    n, k = 400_000, 31
    words = R.synth_seq(4, n)
    seq = gpu.upload(Dna.from_words(words, n))
    oracle = R.count_query(words, 1, n, words.size, k, faithful=False)
    for parts in (1, 2, 3, 8):
        buf, counts = gpu.partition(seq, k, parts)
        assert int(counts.sum()) == oracle.total
        assert np.array_equal(counts, gpu.partition_counts(seq, k, parts))
        host = buf.cpu().numpy().view(np.uint64)
        assert np.array_equal(np.sort(host), np.sort(R.generate_kmers(words, n, k, window=True)))
        tot = [0, 0, 0]
        off = 0
        all_k, all_c = [], []
        for p in range(parts):
            c = int(counts[p])
            part = host[off:off + c]
            assert all(dnagpu.owner_of(int(x), parts) == p for x in part[:200])
            st, table = gpu.count_keys(buf[off:off + c], k, table=True,
                                       method=dnagpu.COUNT_PARTITION if parts == 2 else dnagpu.COUNT_AUTO)
            kk, cc = table.fetch()
            all_k.append(kk); all_c.append(cc)
            tot[0] += st.total; tot[1] += st.distinct; tot[2] += st.unique
            off += c
        assert tuple(tot) == oracle.stats, parts
        kk = np.concatenate(all_k); cc = np.concatenate(all_c)
        order = np.argsort(kk)
        assert np.array_equal(kk[order], oracle.kmers) and np.array_equal(cc[order], oracle.counts)
    seq.free()


@pytest.mark.parametrize("k,prefix", [(31, None), (32, None), (21, "AC")])
def test_fused_owner_routing_composes_to_the_single_gpu_answer(gpu, k, prefix):
    """dnagpu_shuffle_send / _count: G ranks emulated one after the other on one device.  Every shard is laid
    out by partition digit, each owner receives its digit slices peer-major, merges and counts."""
    import torch
    from dnagpu.distributed import owner_digits, shard_of
    n = 3_000_000
    words = R.synth_seq(31, n)
    pk = R.kmer_make(prefix) if prefix else None
    oracle = R.count_query(words, 1, n, words.size, k, prefix=pk, faithful=False, threads=4)
    for G in (1, 2, 3, 8):
        plan = gpu.shuffle_plan(n - k + 1, G)
        sends = []
        kept_all = side_all = 0
        for r in range(G):
            first, starts = shard_of(n, k, G, r)
            seq = gpu.synth_range(n, 31, 8, first, starts, k)
            buf = torch.empty(seq.kmer_count(k) + 2, dtype=torch.int64, device="cuda")
            counts, kept, side = gpu.shuffle_send(seq, k, plan, buf, prefix=prefix)
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            sends.append((buf, counts, offs))
            kept_all += kept
            side_all += side
            seq.free()
        assert kept_all == oracle.total
        tot = [0, 0]
        kk_all, cc_all = [], []
        for o in range(G):
            lo, hi = owner_digits(plan, o)
            assert all(gpu.lib.dnagpu_shuffle_owner(plan, d) == o for d in (lo, hi - 1)) if hi > lo else True
            pieces, chunks = [], []
            for buf, counts, offs in sends:                    # peer-major, digits in order
                pieces.extend(int(c) for c in counts[lo:hi])
                chunks.append(buf[int(offs[lo]):int(offs[hi])])
            recv = torch.cat(chunks) if chunks else torch.empty(0, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()               # torch's stream is not the library's
            st, table = gpu.shuffle_count(recv, np.array(pieces, dtype=np.uint64), hi - lo, plan, k, table=True)
            assert st.total == sum(pieces)
            tot[0] += st.distinct
            tot[1] += st.unique
            a, b = table.fetch()
            kk_all.append(a)
            cc_all.append(b)
        want_d, want_u = oracle.distinct, oracle.unique
        assert tot[0] + (side_all > 0) == want_d and tot[1] + (side_all == 1) == want_u, (G, k)
        kk = np.concatenate(kk_all)
        cc = np.concatenate(cc_all)
        if side_all:
            kk = np.append(kk, np.uint64(2**64 - 1))
            cc = np.append(cc, np.uint64(side_all))
        order = np.argsort(kk)
        assert np.array_equal(kk[order], oracle.kmers) and np.array_equal(cc[order], oracle.counts), (G, k)


@pytest.mark.parametrize("k,prefix,pattern", [(31, None, None), (32, None, None), (21, "AC", None),
                                              (12, None, "NNNNNNNNNNNN"), (31, "G", "S" + "N" * 30)])
def test_scatter_to_destinations_composes_to_the_single_gpu_answer(gpu, k, prefix, pattern):
    """dnagpu_shuffle_hist / _scatter_to (and, with a WHERE clause, dnagpu_collect + the _keys forms): the
    exchange that stores straight into the owners' buffers, G ranks emulated on one device with local buffers
    standing in for the peer-mapped ones (the layout arithmetic of distributed.count_sharded_peer)."""
    import torch
    from dnagpu.distributed import owner_digits, shard_of
    n = 2_500_000
    words = R.synth_seq(32, n)
    oracle = R.count_query(words, 1, n, words.size, k, prefix=R.kmer_make(prefix) if prefix else None,
                           pattern=pattern, faithful=False, threads=4)
    where = dict(prefix=prefix, pattern=pattern)
    for G in (1, 2, 5):
        plan = gpu.shuffle_plan(n - k + 1, G)
        seqs = [gpu.synth_range(n, 32, 8, *shard_of(n, k, G, r), k) for r in range(G)]
        listed = [gpu.collect(s, k, **where) if (prefix or pattern) else None for s in seqs]
        counts = np.stack([gpu.shuffle_hist_keys(l, plan) if l is not None else gpu.shuffle_hist(s, k, plan)
                           for s, l in zip(seqs, listed)])
        if prefix or pattern:                                   # the two histogram forms agree
            assert np.array_equal(counts, np.stack([gpu.shuffle_hist(s, k, plan, **where) for s in seqs]))
        ranges = [owner_digits(plan, r) for r in range(G)]
        recv = [torch.full((int(counts[:, a:b].sum()) + 2,), -7, dtype=torch.int64, device="cuda") for a, b in ranges]
        torch.cuda.synchronize()                   # torch's stream is not the library's
        kept_all = side_all = 0
        for r in range(G):
            dest = np.zeros(plan.n_digits, dtype=np.uint64)
            for o, (a, b) in enumerate(ranges):
                block = counts[:, a:b]
                within = np.concatenate([[0], np.cumsum(block[r])[:-1]]).astype(np.uint64) if b > a else np.zeros(0, np.uint64)
                dest[a:b] = np.uint64(recv[o].data_ptr()) + np.uint64(8) * (np.uint64(int(block[:r].sum())) + within)
            if listed[r] is not None:
                kept, side = int(listed[r].numel()), gpu.shuffle_scatter_keys_to(listed[r], plan, dest)
            else:
                kept, side = gpu.shuffle_scatter_to(seqs[r], k, plan, dest)
            kept_all += kept
            side_all += side
        assert kept_all == oracle.total
        distinct = unique = 0
        for o, (a, b) in enumerate(ranges):
            assert int((recv[o][:-2] == -7).sum()) == 0 or k == 32   # every slot of the layout was written
            st = gpu.shuffle_count_addr(recv[o].data_ptr(), counts[:, a:b].reshape(-1), b - a, plan, k)
            distinct += st.distinct
            unique += st.unique
        assert (distinct + (side_all > 0), unique + (side_all == 1)) == (oracle.distinct, oracle.unique), (G, k)
        for s in seqs:
            s.free()


def test_collect_keeps_the_rows_of_filter_in_any_order(gpu):
    """dnagpu_collect == dnagpu_filter as a multiset; a short buffer reports the need."""
    import torch
    n, k = 1_000_003, 9
    seq = gpu.synth(n, 41)
    for where in (dict(prefix="ACG"), dict(pattern="NNWNNSNNN"), dict(prefix="T", pattern="NNNNNNNNN"), dict()):
        a = gpu.filter(seq, k, **where).cpu().numpy()
        b = gpu.collect(seq, k, **where).cpu().numpy()
        assert np.array_equal(np.sort(a), np.sort(b))
    out = torch.empty(16, dtype=torch.int64, device="cuda")
    import ctypes
    need = ctypes.c_uint64()
    w, _keep = dnagpu._where("A", None)
    rc = gpu.lib.dnagpu_collect(gpu.handle, seq.handle, k, ctypes.byref(w), out.data_ptr(), 16, ctypes.byref(need))
    assert rc == 21 and need.value == gpu.filter_count(seq, k, prefix="A")
    seq.free()


def test_shards_with_overlap_cover_the_sequence_once(gpu):
    """Base-range shards with a (k-1)-base overlap (synth_range + start limit) reproduce the
    k-mers of the whole sequence exactly once."""
    n, k, G = 1_000_000, 31, 4
    words = R.synth_seq(6, n)
    whole = R.generate_kmers(words, n, k, window=True)
    per = ((n + G - 1) // G + 31) // 32 * 32
    pieces = []
    for g in range(G):
        first = g * per
        starts = min(per, max(0, (n - k + 1) - first))
        seq = gpu.synth_range(n, 6, 8, first, starts, k)
        assert seq.kmer_count(k) == starts
        pieces.append(gpu.extract(seq, k).cpu().numpy().view(np.uint64))
        seq.free()
    assert np.array_equal(np.concatenate(pieces), whole)


def test_two_gpu_peer_exchange_end_to_end():
    """bench.py's N = 2 path (exchange fused into the scatter kernel over peer memory) on a small workload;
    needs two GPUs, skipped otherwise."""
    import json
    import os
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, k = 20_000_000, 21
    words = R.synth_seq(2, 100_000_000, first_word=0, n_words=(n + 31) // 32)  # c2's stream, cut to n bases
    for exchange in ("gather", "peer", "fused", "routed"):
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                              "--gpus", "2", "--workload", "c2", "--steps", "1", "--warmup", "3", "--e2e-steps", "1",
                              "--cpu-sample", "1000000", "--no-extract", "--exchange", exchange],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads(out.stdout.strip().splitlines()[-1])
        assert d["n_gpus"] == 2 and exchange in d["config"]["parallelism"]
        assert {q: d["result"][q] for q in ("total", "distinct", "unique")} == \
            {"total": 99999980, "distinct": 87735270, "unique": 77017094}
        assert d["result"]["oracle"].startswith("equal")
    del words


# ---- ingest codec (dna_in / dna_out on the device) -------------------------------------------------
def test_encode_decode_dna_match_the_oracle(gpu):
    rng = np.random.default_rng(77)
    for n in (1, 15, 31, 32, 33, 64, 100, 4097, 1_000_003):
        s = "".join(rng.choice(list("ATCG"), size=n))
        d = gpu.encode_dna(s)
        words, length = R.encode_dna(s)
        assert d.length == length and np.array_equal(d.words, words), n
        assert gpu.decode_dna(d) == s == R.decode_dna(words, n)
    seq = gpu.seq_from_text("ATCGATCGATCGATCGACG")
    st, table = gpu.count(seq, 5, table=True)
    kk, cc = table.sorted()
    assert dict(zip(dnagpu.kmer_strings(kk, 5), cc.tolist())) == {"ATCGA": 4, "CGATC": 3, "GATCG": 3, "TCGAT": 3,
                                                                  "TCGAC": 1, "CGACG": 1}   # test.sql:95-104


def test_encode_dna_errors_are_the_references(gpu):
    with pytest.raises(DnaError, match="DNA sequence cannot be empty") as e:
        gpu.encode_dna("")
    assert e.value.code == 9
    good = "ACGT" * 5000
    for pos, ch in ((0, "N"), (31, "x"), (32, "U"), (19_999, "a"), (7777, " ")):
        bad = good[:pos] + ch + good[pos + 1:]
        with pytest.raises(DnaError, match=f"Invalid character in DNA sequence: {ch}") as e:
            gpu.encode_dna(bad)
        assert e.value.code == 8
    two = good[:100] + "Z" + good[101:5000] + "Q" + good[5001:]          # the FIRST offender is reported
    with pytest.raises(DnaError, match="Invalid character in DNA sequence: Z"):
        gpu.encode_dna(two)
    with pytest.raises(R.RefError):
        R.encode_dna(two)


# ---- the unordered predicate scan in both forms: Shift-And automaton (default) and plane test ----------------
def _rand_pattern(rng, k, density):
    return "".join(IUPAC[int(rng.integers(0, 16))] if rng.random() < density else "N" for _ in range(k))


@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 17, 21, 31, 32])
def test_collect_shift_and_equals_plane_test_equals_oracle(gpu, k):
    rng = np.random.default_rng(900 + k)
    for n in (k, k + 1, 127, 128, 129, 160, 1000, 4097, 200_003):
        words = rand_dna_words(rng, n)
        seq = gpu.upload(Dna.from_words(words, n))
        for trial in range(3):
            pattern = _rand_pattern(rng, k, 0.3) if trial else "N" * k
            plen = int(rng.integers(0, min(k, 4) + 1))
            prefix = R.decode_kmer(int(rng.integers(0, 4 ** plen)), plen) if plen else None
            want = R.filter_kmers(words, n, k, prefix=R.kmer_make(prefix) if prefix else None, pattern=pattern)
            a = gpu.collect(seq, k, prefix=prefix, pattern=pattern).cpu().numpy().view(np.uint64)
            b = gpu.collect(seq, k, prefix=prefix, pattern=pattern, planes=True).cpu().numpy().view(np.uint64)
            assert np.array_equal(np.sort(a), np.sort(want)), (n, k, prefix, pattern)
            assert np.array_equal(np.sort(b), np.sort(want)), (n, k, prefix, pattern)
            # the ordered scan (generate_kmers ... WHERE, rows in sequence order) in both forms
            c = gpu.filter(seq, k, prefix=prefix, pattern=pattern).cpu().numpy().view(np.uint64)
            d = gpu.filter(seq, k, prefix=prefix, pattern=pattern, planes=True).cpu().numpy().view(np.uint64)
            assert np.array_equal(c, want) and np.array_equal(d, want), (n, k, prefix, pattern)
        seq.free()


@pytest.mark.parametrize("bases,stride,k", [(150, 5, 31), (150, 6, 21), (33, 2, 32), (64, 2, 5), (129, 5, 13), (300, 10, 31),
                                            (31, 1, 31)])
def test_collect_shift_and_on_reads_never_spans_rows(gpu, bases, stride, k):
    """Reads: every read is its own dna value; the automaton restarts at every row (runs of <= 128 starts)."""
    rng = np.random.default_rng(bases * 7 + k)
    n_reads = 3000
    words = np.zeros(n_reads * stride, dtype=np.uint64)
    for r in range(n_reads):
        w = rand_dna_words(rng, bases)
        words[r * stride: r * stride + w.size] = w
    seq = gpu.upload_reads(words, n_reads, bases, stride)
    for pattern, prefix in (("N" * (k - 1) + "W", None), (_rand_pattern(rng, k, 0.2), "A"), (None, "AC"[:min(2, k)])):
        pk = R.kmer_make(prefix) if prefix else None
        want = np.concatenate([R.filter_kmers(words[r * stride:(r + 1) * stride], bases, k, prefix=pk, pattern=pattern)
                               for r in range(n_reads)])
        a = gpu.collect(seq, k, prefix=prefix, pattern=pattern).cpu().numpy().view(np.uint64)
        b = gpu.collect(seq, k, prefix=prefix, pattern=pattern, planes=True).cpu().numpy().view(np.uint64)
        assert np.array_equal(np.sort(a), np.sort(want)), (pattern, prefix)
        assert np.array_equal(np.sort(b), np.sort(want)), (pattern, prefix)
        assert np.array_equal(gpu.filter(seq, k, prefix=prefix, pattern=pattern).cpu().numpy().view(np.uint64), want)
        assert gpu.filter_count(seq, k, prefix=prefix, pattern=pattern) == want.size
        st, _ = gpu.count(seq, k, prefix=prefix, pattern=pattern, method=dnagpu.COUNT_HASH)
        u, c = np.unique(want, return_counts=True)
        assert (st.total, st.distinct, st.unique) == (int(c.sum()), int(u.size), int((c == 1).sum()))
    seq.free()


def test_collect_shift_and_dense_matches_take_several_rounds(gpu):
    """A clause that keeps (almost) every row: a tile's matches exceed one staging round."""
    n, k = 3_000_000, 11
    words = R.synth_seq(77, n)
    seq = gpu.upload(Dna.from_words(words, n))
    for pattern in ("N" * k, "N" * (k - 1) + "B"):
        want = R.filter_kmers(words, n, k, pattern=pattern)
        a = gpu.collect(seq, k, pattern=pattern).cpu().numpy().view(np.uint64)
        assert np.array_equal(np.sort(a), np.sort(want))
        assert np.array_equal(gpu.filter(seq, k, pattern=pattern).cpu().numpy().view(np.uint64), want)
    seq.free()
