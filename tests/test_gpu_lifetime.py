"""Handle lifetime and stream ordering at the C-ABI boundary.

dnagpu_destroy with live child handles (the round-1 smoke() crash: a Table freed into a destroyed context) must
orphan them, *_free after the destroy must be a no-op on the device, and a torch tensor returned by the binding
must be complete without a manual synchronize even though the library runs on its own non-blocking stream."""
import ctypes as C
import gc

import numpy as np
import pytest

import dnagpu
from dnagpu import _lib
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


def test_destroy_with_live_children_then_free_them_c_abi():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dnagpu_create(C.byref(h), 0) == 0
    n, k = 100_000, 21
    words = R.synth_seq(7, n)
    seq, table, index = C.c_void_p(), C.c_void_p(), C.c_void_p()
    st = _lib.Stats()
    assert lib.dnagpu_seq_upload(h, words.ctypes.data, n, C.byref(seq)) == 0
    assert lib.dnagpu_count(h, seq, k, None, None, C.byref(st), C.byref(table)) == 0
    kp, cp = C.c_void_p(), C.c_void_p()
    assert lib.dnagpu_table_device(table, C.byref(kp), C.byref(cp)) == 0
    assert lib.dnagpu_index_build(h, kp, lib.dnagpu_table_rows(table), k, C.byref(index)) == 0
    rows = lib.dnagpu_table_rows(table)
    lib.dnagpu_destroy(h)                       # children still alive
    # host-side facts survive, device use is refused, frees do not crash
    assert lib.dnagpu_table_rows(table) == rows and lib.dnagpu_table_k(table) == k
    h2 = C.c_void_p()
    assert lib.dnagpu_create(C.byref(h2), 0) == 0
    kk = np.zeros(4, dtype=np.uint64)
    assert lib.dnagpu_table_fetch(h2, table, 0, 4, kk.ctypes.data, kk.ctypes.data) == 20  # DNAGPU_EARG
    n_out = C.c_uint64()
    assert lib.dnagpu_extract(h2, seq, k, None, 0, C.byref(n_out)) == 20
    assert lib.dnagpu_index_equal(h2, index, 0, k, None, 0, C.byref(n_out)) == 20
    assert b"destroyed" in lib.dnagpu_last_error(h2)
    lib.dnagpu_table_free(table)
    lib.dnagpu_seq_free(seq)
    lib.dnagpu_index_free(index)
    # the new context is unaffected
    seq2 = C.c_void_p()
    assert lib.dnagpu_seq_upload(h2, words.ctypes.data, n, C.byref(seq2)) == 0
    st2 = _lib.Stats()
    assert lib.dnagpu_count(h2, seq2, k, None, None, C.byref(st2), None) == 0
    assert (st2.total, st2.distinct, st2.unique) == (st.total, st.distinct, st.unique)
    lib.dnagpu_destroy(h2)                      # frees seq2 itself
    lib.dnagpu_seq_free(seq2)


def test_context_close_with_live_objects_python():
    ctx = dnagpu.Context(0)
    n, k = 50_000, 5
    words = R.synth_seq(9, n)
    dna = dnagpu.Dna.from_words(words, n)
    seq = ctx.upload(dna)
    st, table = ctx.count_kmers(dna, k)
    keys = ctx.extract(seq, k)
    index = ctx.index_build(keys, k)
    ctx.close()                                 # what smoke() did in round 1 with `table` alive
    assert table.handle is None and seq.handle is None and index.handle is None
    del table, seq, index
    gc.collect()
    # the opposite order: context collected before its children
    ctx = dnagpu.Context(0)
    seq = ctx.upload(dna)
    st2, table = ctx.count(seq, k, table=True)
    assert (st2.total, st2.distinct, st2.unique) == (st.total, st.distinct, st.unique)
    del ctx
    gc.collect()
    table.free()
    seq.free()


def test_large_extract_read_back_without_manual_sync(gpu):
    """Context.extract() returned while its kernel was still queued on the library's stream."""
    n, k = 64_000_000, 31
    seq = gpu.synth(n, 11)
    for _ in range(3):
        rows = gpu.extract(seq, k)
        got = rows[-1_000_000:].cpu().numpy().view(np.uint64)   # torch's stream; no ctx.synchronize()
        words = R.synth_seq(11, n)
        want = R.generate_kmers(words, n, k, window=True)[-1_000_000:]
        assert np.array_equal(got, want)
        del rows
    kept = gpu.filter(seq, k, prefix="ACGT")
    want = R.filter_kmers(words, n, k, prefix=R.kmer_make("ACGT"))
    assert np.array_equal(kept.cpu().numpy().view(np.uint64), want)
    seq.free()


def test_torch_default_stream_is_lent_not_replaced():
    import torch
    ctx = dnagpu.Context(0, torch_stream=True)   # torch's current stream is the legacy default stream (handle 0)
    n, k = 8_000_000, 21
    words = R.synth_seq(13, n)
    t = torch.from_numpy(np.concatenate([words, np.zeros(2, np.uint64)]).view(np.int64)).cuda()  # queued on torch's stream
    seq = ctx.wrap(t, n)
    rows = ctx.extract(seq, k)
    got = rows.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, R.generate_kmers(words, n, k, window=True))
    seq.free()
    ctx.close()


def test_upload_returns_with_the_callers_buffer_free(gpu):
    """dnagpu_seq_upload "copies; the caller keeps ownership": overwrite the pinned buffer right after the call."""
    import torch
    n, k = 32_000_000, 21
    words = R.synth_seq(17, n)
    pinned = torch.empty(words.size + 2, dtype=torch.int64).pin_memory()
    pinned.zero_()
    pinned[:words.size] = torch.from_numpy(words.view(np.int64))
    seq = gpu.upload_words(C.c_void_p(pinned.data_ptr()), n)
    pinned.fill_(-1)                            # the caller reuses its buffer
    got = seq.download()
    assert np.array_equal(got, words)
    seq.free()


def test_exact_flag_gives_the_same_answer(gpu):
    """DNAGPU_COUNT_FLAG_EXACT: histogram + scan at both partition levels instead of the optimistic regions."""
    n, seed = 30_000_000, 5
    seq = gpu.synth(n, seed)
    for k in (15, 31, 32):
        a, _ = gpu.count(seq, k, method=dnagpu.COUNT_PARTITION)
        b, tb = gpu.count(seq, k, method=dnagpu.COUNT_PARTITION, exact=True, table=True)
        assert (a.total, a.distinct, a.unique) == (b.total, b.distinct, b.unique)
        assert tb.rows == b.distinct
        tb.free()
    seq.free()
