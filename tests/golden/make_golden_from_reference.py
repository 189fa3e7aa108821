#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.npz by running the reference's OWN code
(/root/reference/dna.c compiled unmodified against oracle/pgshim, driven by oracle/pgshim/driver.c)
on seeded random inputs.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_from_reference.py

The .npz is committed: the GPU box has no /root/reference, and tests compare both the oracle and
the CUDA path against these vectors."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_real as P  # noqa: E402

IUPAC = "ATCGUWSMKRYBDHVN"


def rand_words(rng, n):
    w = rng.integers(0, 2**64, size=(n + 31) // 32, dtype=np.uint64)
    if n % 32:
        w[-1] &= np.uint64((1 << (2 * (n % 32))) - 1)
    return w


def main():
    rng = np.random.default_rng(20261018)
    arrays, cases = {}, []
    cid = 0
    for k in (1, 2, 3, 5, 8, 13, 16, 21, 31, 32):
        for n in (k, 33, 64, 257, 1500):
            if n < k:
                continue
            # low-entropy stretches so that counts > 1 exist at every k
            w = rand_words(rng, n)
            if n >= 257:
                w[2:4] = w[0:2]
            plen = int(rng.integers(0, min(k, 3) + 1))
            rows_all = P.generate_kmers(w, n, k)
            prefix = None
            if plen:
                prefix = (int(rows_all[int(rng.integers(0, rows_all.size))]) & ((1 << (2 * plen)) - 1), plen)
            pat = ["N"] * k
            for pos in rng.choice(k, size=min(k, 2), replace=False):
                pat[pos] = str(rng.choice(list(IUPAC)))
            pattern = "".join(pat)
            rows_f = P.generate_kmers(w, n, k, prefix=prefix, pattern=pattern)
            cnt = P.count(w, 1, n, len(w), k)
            cnt_f = P.count(w, 1, n, len(w), k, prefix=prefix, pattern=pattern)
            tag = f"c{cid}"
            arrays[f"{tag}_words"] = w
            arrays[f"{tag}_rows"] = rows_all
            arrays[f"{tag}_rows_where"] = rows_f
            arrays[f"{tag}_kmers"] = cnt.kmers
            arrays[f"{tag}_counts"] = cnt.counts
            arrays[f"{tag}_kmers_where"] = cnt_f.kmers
            arrays[f"{tag}_counts_where"] = cnt_f.counts
            cases.append({"id": tag, "n_bases": n, "k": k, "prefix": prefix, "pattern": pattern,
                          "stats": cnt.stats, "stats_where": cnt_f.stats})
            cid += 1
    # kmer_hash values straight from dna.c:722-735 over the shim's hash_any
    keys = rng.integers(0, 2**63, size=64, dtype=np.uint64)
    arrays["hash_keys"] = keys
    arrays["hash_values"] = np.array([P.kmer_hash(int(x)) for x in keys], dtype=np.uint32)
    arrays["cases_json"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(out, **arrays)
    print(f"{out}: {len(cases)} cases, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
