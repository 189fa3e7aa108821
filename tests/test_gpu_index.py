"""The sorted k-mer index (dnagpu_index_*), the replacement for the reference's SP-GiST operator class
(dna.c:1137-1737): `kmer = x`, `kmer ^@ prefix`, `qkmer @> kmer` over a stored kmer column (test.sql:186-262).
An index scan must return exactly the rows of the sequential scan -- the oracle here is the reference's operator
semantics (oracle/ref_cpu.py: starts_with, contains, kmer_eq by (length, bits)) applied row by row."""
import numpy as np
import pytest

import dnagpu
from dnagpu import DnaError
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


def _column(gpu, col):
    import torch
    return torch.from_numpy(col.view(np.int64)).cuda()


def _scan(col, k, prefix=None, pattern=None):
    """Row numbers a sequential scan keeps."""
    keep = np.ones(col.size, dtype=bool)
    if prefix:
        pb, pl = R.kmer_make(prefix)
        keep &= (col & np.uint64((1 << (2 * pl)) - 1)) == np.uint64(pb)
    rows = np.nonzero(keep)[0]
    if pattern is not None:
        rows = np.array([i for i in rows if R.contains(pattern, int(col[i]), k)], dtype=np.int64)
    return rows.astype(np.int64)


def _rev_key(x, k):
    s = 0
    for j in range(k):
        s = (s << 2) | ((int(x) >> (2 * j)) & 3)
    return s


@pytest.mark.parametrize("k,n", [(1, 1000), (3, 5000), (5, 100_003), (8, 70_001), (16, 50_000), (31, 40_000),
                                 (32, 40_000), (5, 1), (5, 4096), (5, 4097)])
def test_sorted_column_is_a_stable_sort_by_base_string(gpu, k, n):
    rng = np.random.default_rng(100 + k + n)
    hi = 4**k if k < 32 else 2**64
    col = rng.integers(0, hi, size=n, dtype=np.uint64)
    if k >= 16:                                  # few distinct values would never collide: plant repeats
        col[rng.integers(0, n, size=n // 3)] = col[rng.integers(0, n, size=n // 3)]
    ix = gpu.index_build(_column(gpu, col), k)
    assert (ix.rows, ix.k) == (n, k)
    skeys, rows = (t.cpu().numpy() for t in ix.sorted_column())
    want_keys = np.array([_rev_key(x, k) for x in col], dtype=np.uint64)
    order = np.argsort(want_keys, kind="stable")
    assert np.array_equal(skeys.view(np.uint64), want_keys[order])
    assert np.array_equal(rows, order)
    ix.free()


def test_reference_index_queries_return_the_rows_of_the_sequential_scan(gpu):
    """test.sql:186-262: a 1 M-row column of 5-mers; = 'ATCGC', ^@ 'ACTG', 'MRKYN' @>.  (The reference's own trie
    loses rows: 1021 of 1025 and 4036 of 4044, test.sql:186-240.)"""
    rng = np.random.default_rng(5)
    k, n = 5, 999_996
    col = rng.integers(0, 4**k, size=n, dtype=np.uint64)
    ix = gpu.index_build(_column(gpu, col), k)
    bits, length = R.kmer_make("ATCGC")
    got = ix.equal("ATCGC").cpu().numpy()
    assert np.array_equal(got, np.nonzero(col == np.uint64(bits))[0])
    assert np.array_equal(ix.search(prefix="ACTG").cpu().numpy(), _scan(col, k, prefix="ACTG"))
    sub = slice(0, 60_000)                       # the per-row Python oracle for @> is slow: a smaller column
    ix2 = gpu.index_build(_column(gpu, col[sub].copy()), k)
    assert np.array_equal(ix2.search(pattern="MRKYN").cpu().numpy(), _scan(col[sub], k, pattern="MRKYN"))
    assert np.array_equal(ix2.search(prefix="AT", pattern="NNSNN").cpu().numpy(),
                          _scan(col[sub], k, prefix="AT", pattern="NNSNN"))
    # and the index agrees with the column scan of the library on the full column
    for where in (dict(pattern="MRKYN"), dict(prefix="G", pattern="NWNNB"), dict(pattern="ACNNN"), dict(prefix="ATCGC")):
        kept = gpu.filter_keys(_column(gpu, col), k, **where).cpu().numpy().view(np.uint64)
        rows = ix.search(**where).cpu().numpy()
        assert np.array_equal(col[rows], kept) and np.all(np.diff(rows) > 0), where
    ix.free()
    ix2.free()


@pytest.mark.parametrize("k", [1, 2, 7, 13, 21, 31, 32])
def test_equal_and_prefix_ranges_for_every_prefix_length(gpu, k):
    rng = np.random.default_rng(k)
    n = 30_000
    words = R.synth_seq(60 + k, n + k)           # planted repeats and G x 64 / A x 64 windows: equal keys exist
    col = R.generate_kmers(words, n + k - 1, k, window=True)
    ix = gpu.index_build(_column(gpu, col), k)
    for probe in list(col[rng.integers(0, col.size, size=6)]) + [np.uint64(0), np.uint64((1 << (2 * k)) - 1 if k < 32 else 2**64 - 1)]:
        text = R.decode_kmer(int(probe), k)
        assert np.array_equal(ix.equal(text).cpu().numpy(), np.nonzero(col == probe)[0]), text
        for plen in sorted(p for p in {1, 2, k // 2, k - 1, k} if 1 <= p <= k):
            assert np.array_equal(ix.search(prefix=text[:plen]).cpu().numpy(), _scan(col, k, prefix=text[:plen])), (text, plen)
    # a k-mer of another length equals no row (kmer_eq compares the lengths, dna.c:655-668)
    if k > 1:
        assert ix.equal(R.decode_kmer(int(col[0]), k)[:k - 1]).numel() == 0
    ix.free()


def test_patterns_narrow_by_their_leading_bases_and_filter_the_rest(gpu):
    rng = np.random.default_rng(11)
    k, n = 9, 20_000
    col = rng.integers(0, 4**k, size=n, dtype=np.uint64)
    ix = gpu.index_build(_column(gpu, col), k)
    for pattern in ("ACGNNNNNN", "ANNNNNNNT", "NNNNNNNNN", "RYNNNNNNN", "ACGTACGTA", "ACSWNNBNN", "UNNNNNNNN", "AUNNNNNNN"):
        assert np.array_equal(ix.search(pattern=pattern).cpu().numpy(), _scan(col, k, pattern=pattern)), pattern
    assert np.array_equal(ix.search(prefix="ACG", pattern="NNNTNNNNN").cpu().numpy(),
                          _scan(col, k, prefix="ACG", pattern="NNNTNNNNN"))
    assert ix.search(prefix="A", pattern="TNNNNNNNN").numel() == 0      # contradictory leading base
    assert ix.search().numel() == n                                     # no clause: every row
    ix.free()


def test_errors_and_empty_columns(gpu):
    import torch
    rng = np.random.default_rng(2)
    col = rng.integers(0, 4**4, size=500, dtype=np.uint64)
    ix = gpu.index_build(_column(gpu, col), 4)
    with pytest.raises(DnaError, match="Prefix length cannot exceed kmer length"):   # dna.c:854-856
        ix.search(prefix="ACGTA")
    with pytest.raises(DnaError, match="[Ll]ength"):                                # dna.c:1106-1108
        ix.search(pattern="NNN")
    with pytest.raises(DnaError):                                                    # dna.c:876-900
        ix.search(pattern="NNNZ")
    ix.free()
    empty = gpu.index_build(torch.empty(0, dtype=torch.int64, device="cuda"), 4)
    assert empty.rows == 0 and empty.equal("ACGT").numel() == 0 and empty.search(prefix="AC").numel() == 0
    assert empty.search(prefix="ACGTA").numel() == 0     # no row is evaluated, so nothing raises (as in the reference)
    empty.free()
    with pytest.raises(DnaError, match="Invalid k value"):
        gpu.index_build(_column(gpu, col), 33)
    # a short result buffer reports the need
    import ctypes as C
    ix = gpu.index_build(_column(gpu, col), 4)
    need = C.c_uint64()
    out = torch.empty(1, dtype=torch.int64, device="cuda")
    w, _keep = dnagpu._where("A", None)
    rc = gpu.lib.dnagpu_index_search(gpu.handle, ix.handle, C.byref(w), out.data_ptr(), 1, C.byref(need))
    assert rc == 21 and need.value == int(((col & np.uint64(3)) == 0).sum())
    ix.free()


def test_large_column_build_and_queries(gpu):
    """1e8 rows of 31-mers straight from generate_kmers: sortedness, permutation, and range queries."""
    import torch
    n, k = 100_000_000, 31
    seq = gpu.synth(n + k - 1, 77)
    col = gpu.extract(seq, k)
    ix = gpu.index_build(col, k)
    skeys, rows = ix.sorted_column()
    assert bool((skeys[1:].view(torch.int64) != skeys[:-1].view(torch.int64)).any())
    # sorted as unsigned: compare the top bit flipped as signed
    flipped = skeys ^ torch.tensor(-2**63, dtype=torch.int64, device="cuda")
    assert bool((flipped[1:] >= flipped[:-1]).all())
    assert int(rows.sum()) == n * (n - 1) // 2 and int(rows.min()) == 0 and int(rows.max()) == n - 1
    # stability: equal keys keep ascending rows
    same = skeys[1:] == skeys[:-1]
    assert bool((rows[1:][same] > rows[:-1][same]).all())
    probe = int(col[12345].item()) & (2**64 - 1)
    text = R.decode_kmer(probe, k)
    got = ix.equal(text)
    assert bool((col[got] == col[12345]).all()) and int((col == col[12345]).sum()) == got.numel()
    got = ix.search(prefix=text[:6])
    mask6 = (1 << 12) - 1
    assert int(((col & mask6) == (probe & mask6)).sum()) == got.numel() and bool((got[1:] > got[:-1]).all())
    assert bool(((col[got] & mask6) == (probe & mask6)).all())
    ix.free()
    seq.free()
