"""ctypes loader for oracle/_ref/libdnaref.so: the reference's UNMODIFIED dna.c compiled against
the PostgreSQL API shim in oracle/pgshim/ and driven like the executor drives it (driver.c).

TEST INFRASTRUCTURE ONLY.  The library can be built only where /root/reference exists
(`make -C oracle ref`); the built .so travels to the GPU box with the repo snapshot.
`make -C oracle glue` builds the same module with generate_kmers swapped for the GPU glue
(pg/dna_gpu.c -> libdnagpu): the drop-in itself, driven through the same fmgr / SRF protocol.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libdnaref.so")
# the same module with generate_kmers swapped for the GPU glue (pg/dna_gpu.c) + the glue's pushdown functions
GLUE_SO = os.path.join(_HERE, "_ref", "libdnaglue.so")
REFERENCE_SRC = "/root/reference/dna.c"
u64, vp = C.c_uint64, C.c_void_p
_default = None


def available():
    return os.path.exists(SO) or os.path.exists(REFERENCE_SRC)


def glue_available():
    return os.path.exists(GLUE_SO) or os.path.exists(REFERENCE_SRC)


class PgError(ValueError):
    """An ereport(ERROR) raised inside the reference's (or the glue's) code."""


class CountResult:
    def __init__(self, stats, digest, kmers, counts):
        self.total, self.distinct, self.unique = stats
        self.stats = tuple(stats)
        self.digest, self.kmers, self.counts = digest, kmers, counts


def _run(fn, *args):
    buf = C.create_string_buffer(512)
    rc = fn(*args, buf, 512)
    if rc != 0:
        raise PgError(buf.value.decode())


class Module:
    """One built module (the reference, or the reference + GPU glue) behind the executor-like driver."""

    def __init__(self, so, target):
        if not os.path.exists(so):
            if not os.path.exists(REFERENCE_SRC):
                raise FileNotFoundError(f"{so} is not built and /root/reference is absent")
            subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        err = [C.c_char_p, C.c_size_t]
        sig = {
            "dnaref_dna_in": (C.c_int, [C.c_char_p, vp, u64, C.POINTER(u64)] + err),
            "dnaref_dna_out": (C.c_int, [vp, u64, C.c_char_p, C.c_size_t] + err),
            "dnaref_kmer_in": (C.c_int, [C.c_char_p, C.POINTER(u64), C.POINTER(C.c_int32)] + err),
            "dnaref_kmer_out": (C.c_int, [u64, C.c_int32, C.c_char_p, C.c_size_t] + err),
            "dnaref_qkmer_in": (C.c_int, [C.c_char_p] + err),
            "dnaref_starts_with": (C.c_int, [u64, C.c_int32, u64, C.c_int32, C.POINTER(C.c_int)] + err),
            "dnaref_contains": (C.c_int, [C.c_char_p, u64, C.c_int32, C.POINTER(C.c_int)] + err),
            "dnaref_kmer_hash": (C.c_uint32, [u64]),
            "dnaref_kmer_eq": (C.c_int, [u64, C.c_int32, u64, C.c_int32]),
            "dnaref_generate_kmers": (C.c_int, [vp, u64, C.c_int, u64, C.c_int32, C.c_char_p, vp, u64,
                                                C.POINTER(u64)] + err),
            "dnaref_count": (C.c_int, [vp, u64, u64, u64, C.c_int, u64, C.c_int32, C.c_char_p, C.c_int, vp, vp, vp, vp,
                                       u64, C.POINTER(u64)] + err),
        }
        if hasattr(L, "dnaref_kmer_stats"):   # only the glue build has the pushdown functions
            sig["dnaref_kmer_stats"] = (C.c_int, [vp, u64, C.c_int, C.c_int, u64, C.c_int32, C.c_char_p, vp] + err)
            sig["dnaref_count_kmers"] = (C.c_int, [vp, u64, C.c_int, C.c_int, u64, C.c_int32, C.c_char_p, u64, vp, vp,
                                                   u64, C.POINTER(u64)] + err)
            sig["dnaref_generate_kmers_where"] = (C.c_int, [vp, u64, C.c_int, u64, C.c_int32, C.c_char_p, vp, u64,
                                                            C.POINTER(u64)] + err)
            sig["dnaref_kmer_stats_agg"] = (C.c_int, [vp, vp, vp, u64, C.c_int, vp] + err)
            sig["dnaref_live_tables"] = (C.c_int, [])
            sig["dnaref_device_count"] = (C.c_int, [])
            sig["dnaref_live_contexts"] = (C.c_int, [])
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.L = L

    def dna_in(self, text):
        words = np.zeros(len(text) // 32 + 2, dtype=np.uint64)
        n = u64()
        _run(self.L.dnaref_dna_in, text.encode("ascii", "replace"), words.ctypes.data, words.size, C.byref(n))
        return words[:(n.value + 31) // 32].copy(), n.value

    def dna_out(self, words, n):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        out = C.create_string_buffer(n + 1)
        _run(self.L.dnaref_dna_out, words.ctypes.data, n, out, n + 1)
        return out.value.decode()

    def kmer_in(self, text):
        bits, length = u64(), C.c_int32()
        _run(self.L.dnaref_kmer_in, text.encode("ascii", "replace"), C.byref(bits), C.byref(length))
        return bits.value, length.value

    def kmer_out(self, bits, length):
        out = C.create_string_buffer(40)
        _run(self.L.dnaref_kmer_out, int(bits), length, out, 40)
        return out.value.decode()

    def qkmer_in(self, text):
        _run(self.L.dnaref_qkmer_in, text.encode("ascii", "replace"))
        return text

    def starts_with(self, kbits, klen, pbits, plen):
        r = C.c_int()
        _run(self.L.dnaref_starts_with, int(kbits), klen, int(pbits), plen, C.byref(r))
        return bool(r.value)

    def contains(self, pattern, kbits, klen):
        r = C.c_int()
        _run(self.L.dnaref_contains, pattern.encode("ascii", "replace"), int(kbits), klen, C.byref(r))
        return bool(r.value)

    def kmer_hash(self, bits):
        return int(self.L.dnaref_kmer_hash(int(bits)))

    def kmer_eq(self, a, alen, b, blen):
        return bool(self.L.dnaref_kmer_eq(int(a), alen, int(b), blen))

    def generate_kmers(self, words, n_bases, k, prefix=None, pattern=None):
        """SELECT * FROM generate_kmers(dna, k) AS g(kmer) [WHERE kmer ^@ prefix AND pattern @> kmer]."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        cap = max(0, n_bases - k + 1) if 1 <= k <= 32 else 0
        out = np.empty(cap + 1, dtype=np.uint64)
        n = u64()
        pb, pl = (0, 0) if prefix is None else prefix
        _run(self.L.dnaref_generate_kmers, words.ctypes.data, n_bases, k, pb, pl,
             None if pattern is None else pattern.encode("ascii", "replace"), out.ctypes.data, out.size, C.byref(n))
        return out[:n.value].copy()

    def count(self, words, n_seqs, bases_per_seq, stride, k, prefix=None, pattern=None, threads=1, want_rows=True):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        stats = np.zeros(3, dtype=np.uint64)
        digest = np.zeros(4, dtype=np.uint64)
        pb, pl = (0, 0) if prefix is None else prefix
        cap = n_seqs * max(0, bases_per_seq - k + 1) if want_rows else 0
        kk = np.empty(cap + 1, dtype=np.uint64)
        cc = np.empty(cap + 1, dtype=np.uint64)
        n = u64()
        _run(self.L.dnaref_count, words.ctypes.data, n_seqs, bases_per_seq, stride, k, pb, pl,
             None if pattern is None else pattern.encode("ascii", "replace"), threads, stats.ctypes.data,
             digest.ctypes.data, kk.ctypes.data if want_rows else None, cc.ctypes.data if want_rows else None, cap,
             C.byref(n))
        st = [int(x) for x in stats]
        return CountResult(st, digest, kk[:n.value].copy() if want_rows else None,
                           cc[:n.value].copy() if want_rows else None)

    # ---- the glue's pushdown functions (libdnaglue.so only) ----
    @staticmethod
    def _where(prefix, pattern):
        pb, pl = (0, 0) if prefix is None else prefix
        return pb, pl, None if pattern is None else pattern.encode("ascii", "replace")

    def kmer_stats(self, words, n_bases, k, prefix=None, pattern=None, where_form=None):
        """SELECT * FROM kmer_stats(dna, k [, prefix, pattern]) -> (total, distinct, uniq).  where_form=True calls the
        four-argument form even with both predicates NULL."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        st = np.zeros(3, dtype=np.int64)
        pb, pl, pat = self._where(prefix, pattern)
        four = (prefix is not None or pattern is not None) if where_form is None else where_form
        _run(self.L.dnaref_kmer_stats, words.ctypes.data, n_bases, k, 1 if four else 0, pb, pl, pat, st.ctypes.data)
        return tuple(int(x) for x in st)

    def count_kmers(self, words, n_bases, k, prefix=None, pattern=None, stop_after=None):
        """SELECT * FROM count_kmers(dna, k [, prefix, pattern]) [LIMIT stop_after] -> (kmers, counts) sorted by kmer
        (unsorted when the scan is abandoned after stop_after rows), plus the Kmer.length check."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        cap = max(0, n_bases - k + 1) if 1 <= k <= 32 else 0
        kk = np.empty(cap + 1, dtype=np.uint64)
        cc = np.empty(cap + 1, dtype=np.int64)
        n = u64()
        pb, pl, pat = self._where(prefix, pattern)
        _run(self.L.dnaref_count_kmers, words.ctypes.data, n_bases, k, 1 if (prefix is not None or pattern is not None) else 0,
             pb, pl, pat, 2**64 - 1 if stop_after is None else stop_after, kk.ctypes.data, cc.ctypes.data, kk.size,
             C.byref(n))
        if stop_after is not None:
            return kk[:n.value].copy(), cc[:n.value].copy()
        order = np.argsort(kk[:n.value], kind="stable")
        return kk[:n.value][order], cc[:n.value][order]

    def generate_kmers_where(self, words, n_bases, k, prefix=None, pattern=None):
        """SELECT * FROM generate_kmers_where(dna, k, prefix, pattern): rows in sequence order."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        cap = max(0, n_bases - k + 1) if 1 <= k <= 32 else 0
        out = np.empty(cap + 1, dtype=np.uint64)
        n = u64()
        pb, pl, pat = self._where(prefix, pattern)
        _run(self.L.dnaref_generate_kmers_where, words.ctypes.data, n_bases, k, pb, pl, pat, out.ctypes.data, out.size,
             C.byref(n))
        assert n.value <= cap, "a row came back with the wrong Kmer.length"
        return out[:n.value].copy()

    def kmer_stats_agg(self, seqs, k):
        """SELECT (kmer_stats_agg(sequence, k)).* FROM t; seqs = [(words, n_bases) | None (a SQL NULL), ...]."""
        offs, lens, parts, pos = [], [], [], 0
        for s in seqs:
            if s is None:
                offs.append(pos)
                lens.append(2**64 - 1)
                continue
            w, n = s
            w = np.ascontiguousarray(w, dtype=np.uint64)[: (n + 31) // 32]
            offs.append(pos)
            lens.append(n)
            parts.append(w)
            pos += w.size
        words = np.concatenate(parts + [np.zeros(1, dtype=np.uint64)]) if parts else np.zeros(1, dtype=np.uint64)
        offs, lens = np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint64)
        st = np.zeros(3, dtype=np.int64)
        _run(self.L.dnaref_kmer_stats_agg, words.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(seqs), k, st.ctypes.data)
        return tuple(int(x) for x in st)

    def live_tables(self):
        return int(self.L.dnaref_live_tables())

    def device_count(self):
        """GPUs of the glue's backend context (0 before its first use; DNAGPU_DEVICES selects them)."""
        return int(self.L.dnaref_device_count())

    def live_contexts(self):
        return int(self.L.dnaref_live_contexts())


def reference():
    """The reference's own module (libdnaref.so)."""
    global _default
    if _default is None:
        _default = Module(SO, "ref")
    return _default


def glue():
    """The reference's module with generate_kmers swapped for the GPU glue (libdnaglue.so; needs a GPU to call)."""
    return Module(GLUE_SO, "glue")


def lib():
    return reference().L


def _forward(name):
    def f(*a, **kw):
        return getattr(reference(), name)(*a, **kw)
    f.__name__ = name
    f.__doc__ = getattr(Module, name).__doc__
    return f


for _n in ("dna_in", "dna_out", "kmer_in", "kmer_out", "qkmer_in", "starts_with", "contains", "kmer_hash", "kmer_eq",
           "generate_kmers", "count"):
    globals()[_n] = _forward(_n)
