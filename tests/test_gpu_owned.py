"""The multi-GPU data plane, checked on ONE device: every GPU walks the whole sequence (resident as base-range
pieces) and keeps the k-mers dnagpu_owner_of assigns to it (dnagpu_count_opts.owner_parts / owner_part,
k_part_scatter_owned).  Running the G owner shares one after the other on one GPU must give disjoint key sets
whose union is exactly the oracle's grouped result."""
import numpy as np
import pytest
import torch

import dnagpu
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


def _owned_union(gpu, seq, k, G, **kw):
    total = distinct = unique = 0
    kmers, counts = [], []
    for part in range(G):
        st, table = gpu.count(seq, k, owner=(G, part), table=True, **kw)
        a, b = table.fetch()
        table.free()
        assert a.size == st.distinct and int(b.sum()) == st.total
        if a.size:
            sample = a[:: max(1, a.size // 300)]
            assert all(dnagpu.owner_of(int(x), G) == part for x in sample)
        total, distinct, unique = total + st.total, distinct + st.distinct, unique + st.unique
        kmers.append(a)
        counts.append(b)
    kmers, counts = np.concatenate(kmers), np.concatenate(counts)
    order = np.argsort(kmers, kind="stable")
    return (total, distinct, unique), kmers[order], counts[order]


@pytest.mark.parametrize("G", [2, 3, 4, 5, 8, 16])
@pytest.mark.parametrize("k", [5, 14, 21, 31, 32])
def test_owner_shares_are_disjoint_and_add_up_to_the_oracle(gpu, G, k):
    n, seed = 3_000_000, 5            # seed 5 / words 8-9: contains 'G' x 32 (the k = 32 sentinel key)
    words = R.synth_seq(seed, n)
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8)
    seq = gpu.synth(n, seed)
    stats, kmers, counts = _owned_union(gpu, seq, k, G)
    assert stats == want.stats
    assert np.array_equal(kmers, want.kmers) and np.array_equal(counts, want.counts)
    seq.free()


def test_owner_share_exact_flag_and_tiny_inputs(gpu):
    for n, k in ((31, 31), (40, 5), (1000, 8), (70_001, 21)):
        words = R.synth_seq(3, n)
        want = R.count_query(words, 1, n, words.size, k, faithful=False)
        seq = gpu.upload(dnagpu.Dna.from_words(words, n))
        for G in (2, 5, 8):
            for exact in (False, True):
                stats, kmers, counts = _owned_union(gpu, seq, k, G, exact=exact)
                assert stats == want.stats, (n, k, G, exact)
                assert np.array_equal(kmers, want.kmers) and np.array_equal(counts, want.counts)
        seq.free()


@pytest.mark.parametrize("G", [3, 4])
def test_owner_share_of_heavily_repeated_input_falls_back_exactly(gpu, G):
    """poly-A + a short tandem repeat: one owner keeps millions of copies of a few k-mers -- the optimistic regions
    and the per-CTA lists overflow and the exact key-list form takes over (both ownership forms: 4 owners = linear,
    3 = multiplicative)."""
    n, k = 4_000_000, 21
    rng = np.random.default_rng(7)
    words = rng.integers(0, 2**64, size=n // 32, dtype=np.uint64)
    words[1000:60_000] = 0                                        # ~1.9 M x 'A'
    words[70_000:100_000] = np.uint64(0x1B1B1B1B1B1B1B1B)        # (GCTA)n
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8)
    seq = gpu.upload(dnagpu.Dna.from_words(words, n))
    stats, kmers, counts = _owned_union(gpu, seq, k, G)
    assert stats == want.stats
    assert np.array_equal(kmers, want.kmers) and np.array_equal(counts, want.counts)
    seq.free()


def _pieces(words, n, cuts):
    """Cut a packed sequence at base positions `cuts` (multiples of 32) into device tensors that each carry the
    31-base overlap and a zero pad word, the way one shard per GPU would hold them."""
    bounds = [0] + list(cuts) + [n]
    tensors, first, starts = [], [], []
    n_words = (n + 31) // 32
    for a, b in zip(bounds[:-1], bounds[1:]):
        w0, w1 = a // 32, min(n_words, (b + 31 + 31) // 32)
        piece = np.concatenate([words[w0:w1], np.zeros(3, np.uint64)])
        if len(piece) % 2:
            piece = np.concatenate([piece, np.zeros(1, np.uint64)])
        tensors.append(torch.from_numpy(piece.view(np.int64)).cuda())
        first.append(a)
        starts.append(b - a)
    return tensors, first, starts


@pytest.mark.parametrize("k", [13, 31, 32])
def test_sequence_in_pieces_walked_in_ring_order(gpu, k):
    n, seed, G = 5_000_000, 5, 4
    words = R.synth_seq(seed, n)
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=8, want_rows=False)
    tensors, first, starts = _pieces(words, n, [32 * 40_000, 32 * 41_000, 32 * 120_000])
    torch.cuda.synchronize()
    total = distinct = unique = 0
    for part in range(G):
        ring = [(part + i) % G for i in range(G)]          # own piece first, the others in ring order
        seq = gpu.wrap_pieces([tensors[i].data_ptr() for i in ring], [first[i] for i in ring],
                              [starts[i] for i in ring], n, keep=tensors)
        assert seq.kmer_count(k) == n - k + 1
        st, _ = gpu.count(seq, k, owner=(G, part))
        total, distinct, unique = total + st.total, distinct + st.distinct, unique + st.unique
        seq.free()
    assert (total, distinct, unique) == want.stats


def test_pieces_argument_errors(gpu):
    n = 32 * 1000
    words = R.synth_seq(1, n)
    tensors, first, starts = _pieces(words, n, [32 * 500])
    ptrs = [t.data_ptr() for t in tensors]
    with pytest.raises(dnagpu.DnaError):                    # gap: the ranges do not tile the sequence
        gpu.wrap_pieces(ptrs, [0, 32 * 600], [32 * 500, 32 * 400], n)
    with pytest.raises(dnagpu.DnaError):                    # not a multiple of 32
        gpu.wrap_pieces(ptrs, [0, 16010], [16010, n - 16010], n)
    seq = gpu.wrap_pieces(ptrs, first, starts, n, keep=tensors)
    with pytest.raises(dnagpu.DnaError):                    # pieces need an owner restriction
        gpu.count(seq, 21)
    with pytest.raises(dnagpu.DnaError):                    # no WHERE clause with an owner restriction
        gpu.count(seq, 21, prefix="AC", owner=(2, 0))
    with pytest.raises(dnagpu.DnaError):
        gpu.count(seq, 21, owner=(2, 2))
    seq.free()


def test_count_keys_of_a_large_list_takes_the_optimistic_level_1(gpu):
    """dnagpu_count_keys on >= 2^24 keys: fixed regions + one scatter from the list (what every GPU of a multi-GPU
    count runs on the k-mers it kept), against the oracle; a skewed list falls back to the exact form."""
    n, k, seed = 40_000_000, 31, 5
    words = R.synth_seq(seed, n)
    want = R.count_query_big(words, 1, n, words.size, k, threads=8)
    seq = gpu.synth(n, seed)
    keys = gpu.extract(seq, k)
    st, table = gpu.count_keys(keys, k, table=True)
    assert (st.total, st.distinct, st.unique) == want.stats
    a, b = table.fetch()
    assert np.array_equal(R.pairs_digest(a, b), want.digest)
    table.free()
    skew = keys.clone()
    skew[: n // 2] = keys[12345]                      # 20 M copies of one k-mer
    st, _ = gpu.count_keys(skew, k)
    u, c = np.unique(skew.cpu().numpy().view(np.uint64), return_counts=True)
    assert (st.total, st.distinct, st.unique) == (int(c.sum()), int(u.size), int((c == 1).sum()))
    seq.free()
