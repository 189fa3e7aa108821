"""BASELINE.json sizes.  configs[1] (100 Mbp, k = 21) is compared with the oracle through an
order-independent digest of the grouped rows and the three aggregates; sizes the oracle cannot finish in
seconds are checked through size-independent properties (two unrelated GPU methods must agree, totals
must equal the row count, shards must add up)."""
import numpy as np
import pytest

import dnagpu
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


def test_config2_100mbp_k21_rows_digest_matches_oracle(gpu):
    n, k, seed = 100_000_000, 21, 2
    words = R.synth_seq(seed, n)
    want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=16, want_rows=False, expected_keys=n)
    seq = gpu.synth(n, seed)
    assert np.array_equal(seq.download()[:1000], words[:1000])
    st, table = gpu.count(seq, k, table=True)
    assert (st.total, st.distinct, st.unique) == want.stats == (n - k + 1, want.distinct, want.unique)
    kmers, counts = table.fetch()
    assert kmers.size == want.distinct and int(counts.sum()) == n - k + 1
    assert np.array_equal(R.pairs_digest(kmers, counts), want.digest)
    table.free()
    seq.free()


def test_1gbp_k31_two_methods_and_shards_agree(gpu):
    n, k, seed = 1_000_000_000, 31, 5
    seq = gpu.synth(n, seed)
    a, _ = gpu.count(seq, k, method=dnagpu.COUNT_PARTITION)
    b, _ = gpu.count(seq, k, method=dnagpu.COUNT_HASH)
    assert (a.total, a.distinct, a.unique) == (b.total, b.distinct, b.unique)
    assert a.total == n - k + 1 and a.unique <= a.distinct <= a.total
    seq.free()
    # the same sequence as 3 base-range shards routed to 3 owners and counted per owner
    from dnagpu.distributed import owner_digits, shard_of
    import torch
    G = 3
    plan = gpu.shuffle_plan(n - k + 1, G)
    sends, kept_all = [], 0
    for r in range(G):
        first, starts = shard_of(n, k, G, r)
        s = gpu.synth_range(n, seed, 8, first, starts, k)
        buf = torch.empty(s.kmer_count(k) + 2, dtype=torch.int64, device="cuda")
        counts, kept, side = gpu.shuffle_send(s, k, plan, buf)
        sends.append((buf, counts, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)))
        kept_all += kept
        s.free()
    assert kept_all == a.total
    distinct = unique = 0
    for o in range(G):
        lo, hi = owner_digits(plan, o)
        pieces = np.concatenate([c[lo:hi] for _, c, _ in sends])
        recv = torch.cat([b_[int(off[lo]):int(off[hi])] for b_, _, off in sends])
        torch.cuda.synchronize()                   # torch's stream is not the library's
        st, _ = gpu.shuffle_count(recv, pieces, hi - lo, plan, k)
        distinct += st.distinct
        unique += st.unique
        del recv
    assert (distinct, unique) == (a.distinct, a.unique)


def test_k_sweep_edges_on_100mbp(gpu):
    """configs[4] in small: every k from 1 to 32 on one sequence, AUTO method, against the oracle's aggregates."""
    n, seed = 20_000_000, 5
    words = R.synth_seq(seed, n)
    seq = gpu.synth(n, seed)
    for k in list(range(1, 33)):
        want = R.count_query(words, 1, n, words.size, k, faithful=False, threads=16, want_rows=False,
                             expected_keys=min(n, 4 ** min(k, 13)))
        st, _ = gpu.count(seq, k)
        assert (st.total, st.distinct, st.unique) == want.stats, k
    seq.free()
