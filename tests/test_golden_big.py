"""tests/golden/big_expected.json cannot drift from the oracle: the entries small enough for the CPU suite are
re-derived here, both through the partitioned form that produced the file and through the plain hash aggregate."""
import json
import os

import numpy as np

from conftest import ROOT
from oracle import ref_cpu as R

with open(os.path.join(ROOT, "tests", "golden", "big_expected.json")) as _f:
    GOLD = json.load(_f)


def test_every_configuration_has_an_entry():
    need = ["c1", "c2", "c2_faithful", "c3", "c3_10m", "c3_10m_prefix", "c3_10m_pattern", "c4"] + \
           [f"c5_k{k}" for k in range(3, 33)]
    for name in need:
        e = GOLD[name]
        assert e["unique"] <= e["distinct"] <= e["total"] and len(e["digest"]) == 4, name
    assert GOLD["c4"]["total"] == 3_100_000_000 - 31 + 1
    for k in range(3, 33):
        assert GOLD[f"c5_k{k}"]["total"] == 1_000_000_000 - k + 1
        if k <= 12:  # every k-mer of a 1 Gbp random sequence occurs for small k
            assert GOLD[f"c5_k{k}"]["distinct"] == 4 ** k


def test_c1_and_c2_re_derived():
    e = GOLD["c1"]
    w = R.synth_seq(e["seed"], e["n_bases"])
    a = R.count_query(w, 1, e["n_bases"], w.size, e["k"], faithful=True)
    assert a.stats == (e["total"], e["distinct"], e["unique"]) and [int(x) for x in a.digest] == e["digest"]
    e = GOLD["c2"]
    w = R.synth_seq(e["seed"], e["n_bases"])
    b = R.count_query_big(w, 1, e["n_bases"], w.size, e["k"], passes=2, threads=8)
    assert b.stats == (e["total"], e["distinct"], e["unique"]) and [int(x) for x in b.digest] == e["digest"]
    assert GOLD["c2_faithful"]["digest"] == e["digest"]


def test_partitioned_form_equals_the_plain_aggregate():
    for n, k, seed in ((3_000_000, 31, 4), (2_000_000, 32, 5), (1_000_000, 5, 5), (1_500_000, 13, 5)):
        w = R.synth_seq(seed, n)
        a = R.count_query(w, 1, n, w.size, k, faithful=False, threads=4, want_rows=False, expected_keys=n)
        for passes in (1, 3):
            b = R.count_query_big(w, 1, n, w.size, k, passes=passes, threads=5)
            assert a.stats == b.stats and np.array_equal(a.digest, b.digest), (n, k, passes)
    # reads with both predicates, faithful per-row starts_with / contains
    nr = 100_000
    w = R.synth_reads(3, nr, 150, 5)
    pk, pat = R.kmer_make("AC"), "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY"
    a = R.count_query(w, nr, 150, 5, 31, prefix=pk, pattern=pat, faithful=True, threads=4, want_rows=False)
    b = R.count_query_big(w, nr, 150, 5, 31, prefix=pk, pattern=pat, passes=2, threads=3)
    c = R.count_query_big(w, nr, 150, 5, 31, prefix=pk, pattern=pat, passes=1, threads=2, faithful=True)
    assert a.stats == b.stats == c.stats and np.array_equal(a.digest, b.digest) and np.array_equal(a.digest, c.digest)


def test_c3_10m_slice_re_derived():
    e = GOLD["c3_10m"]
    nr = 1_000_000  # the first tenth of the reads: a sanity bound, not equality
    w = R.synth_reads(e["seed"], nr, e["bases"], e["stride"])
    b = R.count_query_big(w, nr, e["bases"], e["stride"], e["k"], prefix=R.kmer_make(e["prefix"]), pattern=e["pattern"],
                          threads=8)
    assert 0.08 * e["total"] < b.total < 0.12 * e["total"]
