"""Golden vectors produced by the reference's own dna.c (tests/golden/make_golden_from_reference.py,
run in the build container) -- checked against the oracle on CPU and against the CUDA path on GPU."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import ref_cpu as R


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    return z, json.loads(bytes(z["cases_json"]).decode())


def test_oracle_reproduces_reference_vectors(golden):
    z, cases = golden
    assert len(cases) >= 40
    for c in cases:
        w, n, k = z[c["id"] + "_words"], c["n_bases"], c["k"]
        prefix = tuple(c["prefix"]) if c["prefix"] else None
        assert np.array_equal(R.generate_kmers(w, n, k), z[c["id"] + "_rows"]), c
        assert np.array_equal(R.filter_kmers(w, n, k, prefix=prefix, pattern=c["pattern"]), z[c["id"] + "_rows_where"]), c
        r = R.count_query(w, 1, n, len(w), k)
        assert list(r.stats) == c["stats"]
        assert np.array_equal(r.kmers, z[c["id"] + "_kmers"]) and np.array_equal(r.counts, z[c["id"] + "_counts"])
        r = R.count_query(w, 1, n, len(w), k, prefix=prefix, pattern=c["pattern"])
        assert list(r.stats) == c["stats_where"]
        assert np.array_equal(r.kmers, z[c["id"] + "_kmers_where"])
        assert np.array_equal(r.counts, z[c["id"] + "_counts_where"])
    assert [R.kmer_hash(int(x)) for x in z["hash_keys"]] == [int(v) for v in z["hash_values"]]


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_vectors(golden, gpu):
    import dnagpu
    from dnagpu import Dna, Kmer
    z, cases = golden
    for c in cases:
        w, n, k = z[c["id"] + "_words"], c["n_bases"], c["k"]
        d = Dna.from_words(w, n)
        prefix = Kmer(bits=c["prefix"][0], length=c["prefix"][1]) if c["prefix"] else None
        assert np.array_equal(gpu.generate_kmers(d, k).bits, z[c["id"] + "_rows"]), c
        assert np.array_equal(gpu.filter_kmers(d, k, prefix=prefix, pattern=c["pattern"]).bits,
                              z[c["id"] + "_rows_where"]), c
        for method in (dnagpu.COUNT_AUTO, dnagpu.COUNT_HASH, dnagpu.COUNT_PARTITION):
            seq = gpu.upload(d)
            st, table = gpu.count(seq, k, table=True, method=method)
            kk, cc = table.sorted()
            assert [st.total, st.distinct, st.unique] == c["stats"], (c, method)
            assert np.array_equal(kk, z[c["id"] + "_kmers"]) and np.array_equal(cc, z[c["id"] + "_counts"])
            st, table = gpu.count(seq, k, prefix=prefix, pattern=c["pattern"], table=True, method=method)
            kk, cc = table.sorted()
            assert [st.total, st.distinct, st.unique] == c["stats_where"], (c, method)
            assert np.array_equal(kk, z[c["id"] + "_kmers_where"])
            assert np.array_equal(cc, z[c["id"] + "_counts_where"])
            seq.free()
