"""ctypes loader for oracle/libref_cpu.so (ref_cpu.c, the plain-C restatement of dna.c).

TEST INFRASTRUCTURE ONLY: the checker, never the thing measured or shipped.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libref_cpu.so")
u64 = C.c_uint64
vp = C.c_void_p
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "libref_cpu.so"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    sig = {
        "ref_errmsg": (C.c_char_p, [C.c_int]),
        "ref_dna_words": (u64, [u64]),
        "ref_encode_dna": (C.c_int, [C.c_char_p, u64, vp]),
        "ref_decode_dna": (None, [vp, u64, C.c_char_p]),
        "ref_kmer_make": (C.c_int, [C.c_char_p, C.POINTER(u64), C.POINTER(C.c_int)]),
        "ref_decode_kmer": (C.c_int, [u64, C.c_int, C.c_char_p]),
        "ref_kmer_rows": (u64, [u64, C.c_int]),
        "ref_generate_kmers": (C.c_int, [vp, u64, C.c_int, vp, C.POINTER(u64)]),
        "ref_generate_kmers_window": (C.c_int, [vp, u64, C.c_int, vp, C.POINTER(u64)]),
        "ref_kmer_eq": (C.c_int, [u64, C.c_int, u64, C.c_int]),
        "ref_kmer_hash": (C.c_uint32, [u64]),
        "ref_starts_with": (C.c_int, [u64, C.c_int, u64, C.c_int, C.POINTER(C.c_int)]),
        "ref_starts_with_x86": (C.c_int, [u64, C.c_int, u64, C.c_int, C.POINTER(C.c_int)]),
        "ref_validate_qkmer": (C.c_int, [C.c_char_p]),
        "ref_nucleotide_matches": (C.c_int, [C.c_char, C.c_char]),
        "ref_contains": (C.c_int, [C.c_char_p, u64, C.c_int, C.POINTER(C.c_int)]),
        "ref_filter_kmers": (C.c_int, [vp, u64, C.c_int, u64, C.c_int, C.c_char_p, vp, C.POINTER(u64)]),
        "ref_agg_new": (vp, [u64]),
        "ref_agg_free": (None, [vp]),
        "ref_agg_add": (C.c_int, [vp, u64, u64]),
        "ref_agg_groups": (u64, [vp]),
        "ref_agg_stats": (None, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "ref_agg_sorted": (None, [vp, vp, vp]),
        "ref_agg_digest": (None, [vp, vp]),
        "ref_pairs_digest": (None, [vp, vp, u64, vp]),
        "ref_count_query": (C.c_int, [vp, u64, u64, u64, C.c_int, u64, C.c_int, C.c_char_p, C.c_int, vp]),
        "ref_count_query_mt": (C.c_int, [vp, u64, u64, u64, C.c_int, u64, C.c_int, C.c_char_p, C.c_int,
                                         C.c_int, vp]),
        "ref_count_query_big": (C.c_int, [vp, u64, u64, u64, C.c_int, u64, C.c_int, C.c_char_p, C.c_int,
                                          C.c_int, C.c_int, vp, vp]),
        "ref_synth_seq": (None, [u64, C.c_uint32, u64, u64, u64, vp]),
        "ref_synth_reads": (None, [u64, C.c_uint32, u64, u64, C.c_uint32, C.c_uint32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class RefError(ValueError):
    def __init__(self, code):
        super().__init__(lib().ref_errmsg(code).decode())
        self.code = code


def _ok(rc):
    if rc != 0:
        raise RefError(rc)


def encode_dna(text: str):
    """dna_in: text -> (words, n_bases)."""
    raw = text.encode("ascii", "replace")
    words = np.zeros(max(1, int(lib().ref_dna_words(len(raw)))) + 1, dtype=np.uint64)  # +1: pad like PG never has
    _ok(lib().ref_encode_dna(raw, len(raw), words.ctypes.data))
    return words[:int(lib().ref_dna_words(len(raw)))].copy(), len(raw)


def decode_dna(words, n):
    buf = C.create_string_buffer(n + 1)
    lib().ref_decode_dna(np.ascontiguousarray(words, dtype=np.uint64).ctypes.data, n, buf)
    return buf.value.decode()


def kmer_make(text: str):
    bits, length = u64(), C.c_int()
    _ok(lib().ref_kmer_make(text.encode("ascii", "replace"), C.byref(bits), C.byref(length)))
    return bits.value, length.value


def decode_kmer(bits, k):
    buf = C.create_string_buffer(k + 1)
    _ok(lib().ref_decode_kmer(int(bits), k, buf))
    return buf.value.decode()


def _padded(words):
    """The oracle reads only ceil(n/32) words; keep the array alive and contiguous."""
    return np.ascontiguousarray(words, dtype=np.uint64)


def generate_kmers(words, n_bases, k, window=False):
    words = _padded(words)
    rows = int(lib().ref_kmer_rows(n_bases, k))
    out = np.empty(rows, dtype=np.uint64)
    n = u64()
    fn = lib().ref_generate_kmers_window if window else lib().ref_generate_kmers
    _ok(fn(words.ctypes.data, n_bases, k, out.ctypes.data, C.byref(n)))
    return out[:n.value]


def filter_kmers(words, n_bases, k, prefix=None, pattern=None):
    words = _padded(words)
    pb, pl = (0, 0) if prefix is None else prefix
    rows = int(lib().ref_kmer_rows(n_bases, k))
    out = np.empty(rows, dtype=np.uint64)
    n = u64()
    _ok(lib().ref_filter_kmers(words.ctypes.data, n_bases, k, pb, pl,
                               None if pattern is None else pattern.encode("ascii", "replace"),
                               out.ctypes.data, C.byref(n)))
    return out[:n.value]


def starts_with(kmer_bits, k, prefix_bits, prefix_len, x86=False):
    err = C.c_int()
    fn = lib().ref_starts_with_x86 if x86 else lib().ref_starts_with
    r = fn(int(kmer_bits), k, int(prefix_bits), prefix_len, C.byref(err))
    _ok(err.value)
    return bool(r)


def contains(pattern, kmer_bits, k):
    err = C.c_int()
    r = lib().ref_contains(pattern.encode("ascii", "replace"), int(kmer_bits), k, C.byref(err))
    _ok(err.value)
    return bool(r)


def kmer_hash(bits):
    return int(lib().ref_kmer_hash(int(bits)))


class CountResult:
    def __init__(self, total, distinct, unique, kmers, counts, digest):
        self.total, self.distinct, self.unique = total, distinct, unique
        self.kmers, self.counts, self.digest = kmers, counts, digest

    @property
    def stats(self):
        return (self.total, self.distinct, self.unique)


def count_query(words, n_seqs, bases_per_seq, stride_words, k, prefix=None, pattern=None,
                faithful=True, threads=1, want_rows=True, expected_keys=0):
    """generate_kmers [+ WHERE] + GROUP BY kmer over n_seqs fixed-stride sequences."""
    words = _padded(words)
    pb, pl = (0, 0) if prefix is None else prefix
    L = lib()
    agg = L.ref_agg_new(expected_keys or 1024)
    try:
        pat = None if pattern is None else pattern.encode("ascii", "replace")
        if threads > 1:
            rc = L.ref_count_query_mt(words.ctypes.data, n_seqs, bases_per_seq, stride_words, k, pb, pl, pat,
                                      1 if faithful else 0, threads, agg)
        else:
            rc = L.ref_count_query(words.ctypes.data, n_seqs, bases_per_seq, stride_words, k, pb, pl, pat,
                                   1 if faithful else 0, agg)
        _ok(rc)
        t, d, u = u64(), u64(), u64()
        L.ref_agg_stats(agg, C.byref(t), C.byref(d), C.byref(u))
        digest = np.zeros(4, dtype=np.uint64)
        L.ref_agg_digest(agg, digest.ctypes.data)
        kmers = counts = None
        if want_rows:
            g = int(L.ref_agg_groups(agg))
            kmers = np.empty(g, dtype=np.uint64)
            counts = np.empty(g, dtype=np.uint64)
            L.ref_agg_sorted(agg, kmers.ctypes.data, counts.ctypes.data)
        return CountResult(t.value, d.value, u.value, kmers, counts, digest)
    finally:
        L.ref_agg_free(agg)


def count_query_big(words, n_seqs, bases_per_seq, stride_words, k, prefix=None, pattern=None, faithful=False,
                    passes=1, threads=8):
    """count_query for results that do not fit in memory: rows grouped in hash partitions, each aggregated by the
    oracle's own ref_agg -> CountResult without rows (stats + order-independent digest)."""
    words = _padded(words)
    pb, pl = (0, 0) if prefix is None else prefix
    stats = np.zeros(3, dtype=np.uint64)
    digest = np.zeros(4, dtype=np.uint64)
    _ok(lib().ref_count_query_big(words.ctypes.data, n_seqs, bases_per_seq, stride_words, k, pb, pl,
                                  None if pattern is None else pattern.encode("ascii", "replace"),
                                  1 if faithful else 0, passes, threads, stats.ctypes.data, digest.ctypes.data))
    return CountResult(int(stats[0]), int(stats[1]), int(stats[2]), None, None, digest)


def count_ragged(seqs, k, prefix=None, pattern=None, faithful=True):
    """The table form (test.sql:140-150): seqs = [(words, n_bases), ...], one aggregate."""
    L = lib()
    pb, pl = (0, 0) if prefix is None else prefix
    pat = None if pattern is None else pattern.encode("ascii", "replace")
    agg = L.ref_agg_new(1024)
    try:
        for words, n in seqs:
            words = _padded(words)
            _ok(L.ref_count_query(words.ctypes.data, 1, n, len(words), k, pb, pl, pat, 1 if faithful else 0, agg))
        t, d, u = u64(), u64(), u64()
        L.ref_agg_stats(agg, C.byref(t), C.byref(d), C.byref(u))
        g = int(L.ref_agg_groups(agg))
        kmers = np.empty(g, dtype=np.uint64)
        counts = np.empty(g, dtype=np.uint64)
        L.ref_agg_sorted(agg, kmers.ctypes.data, counts.ctypes.data)
        digest = np.zeros(4, dtype=np.uint64)
        L.ref_agg_digest(agg, digest.ctypes.data)
        return CountResult(t.value, d.value, u.value, kmers, counts, digest)
    finally:
        L.ref_agg_free(agg)


def pairs_digest(kmers, counts):
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    digest = np.zeros(4, dtype=np.uint64)
    lib().ref_pairs_digest(kmers.ctypes.data, counts.ctypes.data, kmers.size, digest.ctypes.data)
    return digest


def synth_seq(seed, n_bases, repeat_every=8, first_word=0, n_words=None):
    if n_words is None:
        n_words = (n_bases + 31) // 32
    out = np.empty(n_words, dtype=np.uint64)
    lib().ref_synth_seq(seed, repeat_every, n_bases, first_word, n_words, out.ctypes.data)
    return out


def synth_reads(seed, n_reads, bases_per_read, stride_words, repeat_every=8, first_read=0):
    out = np.empty(n_reads * stride_words, dtype=np.uint64)
    lib().ref_synth_reads(seed, repeat_every, first_read, n_reads, bases_per_read, stride_words, out.ctypes.data)
    return out
