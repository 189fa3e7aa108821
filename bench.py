#!/usr/bin/env python
"""bench.py -- the k-mer hot path on B200, one JSON line per run.

Workload (BASELINE.json configs[3], the one the metric is quoted on): one synthetic
3.1 Gbp sequence (seeded generator of include/dnagpu_synth.h, every 8th 1024-base block a
planted repeat), k = 31, `GROUP BY kmer` with total / distinct / unique.  At N > 1 the
sequence is sharded by base range with a (k-1)-base overlap, k-mers are routed to their
owner rank by hash (NCCL all-to-all) and each rank counts its partition: total work is
fixed, so scaling is "strong".

  value   Gkmer/s with the packed words already resident in HBM (table init + extract +
          count + aggregates; at N > 1 also partition + exchange), CUDA-event timed.
  e2e     the same query through the host-buffer C-ABI call dnagpu_count_kmers():
          pinned host words -> H2D -> count -> D2H of the three aggregates, every step.
  roofline  the kernel with the largest share of the step (CUDA events around every launch, taken
          inside the timed region): its algorithmic bytes / its average duration, plus the same
          figure for every kernel of the pipeline and for the whole step.
  cpu_baseline  the reference's own dna.c (oracle/_ref: compiled unmodified against a PostgreSQL
          API shim and driven like the executor) on the host cores, bounded sample; falls back
          to the oracle's faithful restatement when that library is not built.

`--impl reference` times the reference's CPU implementation of the same query
(oracle/_ref = the reference's own dna.c when it could be compiled, else the oracle port).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dna-sequences-pg-extension_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (n_bases, k, seed, description)      -- BASELINE.json configs[3], [1], [4]
    "c4": (3_100_000_000, 31, 4, "3.1 Gbp synthetic sequence, k=31 GROUP BY kmer count + total/distinct/unique"),
    "c2": (100_000_000, 21, 2, "100 Mbp synthetic sequence, k=21 full count + total/distinct/unique"),
    "c5": (1_000_000_000, 31, 5, "1 Gbp synthetic sequence, k=31 count"),
    # BASELINE.json configs[2]: reads (each its own dna value), WHERE ^@ AND @> fused into the count
    "c3": (100_000_000 * 150, 31, 3, "100M synthetic 150 bp reads, k=31, WHERE kmer ^@ 'AC' AND "
                                     "'NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY' @> kmer, fused filter-count"),
}
READS = {"c3": {"n_reads": 100_000_000, "bases": 150, "stride": 5, "prefix": "AC",
                "pattern": "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY"}}
REPEAT_EVERY = 8
METRIC = "Gkmer/s counted (k=31) at 1/2/4/8 B200; extraction HBM GB/s vs peak"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ============================== the reference arm ==============================
CHUNK_BASES = 1 << 16  # the CPU arms cut the sample into overlapping chunks, one dna value each


def cpu_reference_rate(n_bases, k, seed, threads, sample_bases, reads=None):
    """Time the reference's CPU implementation of the query on a prefix of the workload.

    oracle/_ref (the reference's own dna.c, compiled unmodified against the PostgreSQL API shim and
    driven like the executor: SRF loop + HashAggregate through kmer_hash/kmer_eq) when it is built,
    else the oracle's faithful restatement.  The prefix is cut into chunks of 65536 start positions
    (each chunk a dna value overlapping the next by k-1 bases, so every k-mer is produced once) so
    that all host threads have work; PostgreSQL itself would run the plan on ONE core because
    generate_kmers is not PARALLEL SAFE (dna--1.0.sql:188-191).
    Returns (Gkmer/s, seconds, (total, distinct, unique), kind)."""
    from oracle import ref_cpu as R
    from oracle import ref_real as P
    kind = "reference" if os.path.exists(P.SO) else "port"
    if reads is not None:  # a prefix of the batch of reads, each read one dna value
        n = max(threads, sample_bases // reads["bases"])
        words = R.synth_reads(seed, n, reads["bases"], reads["stride"], REPEAT_EVERY)
        pk = R.kmer_make(reads["prefix"])
        t0 = time.perf_counter()
        if kind == "reference":
            r = P.count(words, n, reads["bases"], reads["stride"], k, prefix=pk, pattern=reads["pattern"],
                        threads=threads, want_rows=False)
        else:
            r = R.count_query(words, n, reads["bases"], reads["stride"], k, prefix=pk, pattern=reads["pattern"],
                              faithful=True, threads=threads, want_rows=False)
        dt = time.perf_counter() - t0
        rows = n * (reads["bases"] - k + 1)  # k-mers generated and tested, the unit of the metric
        return rows / dt / 1e9, dt, (rows,) + tuple(r.stats[1:]), kind, n * reads["bases"]
    n_chunks = max(1, (sample_bases - (k - 1)) // CHUNK_BASES)
    sample = n_chunks * CHUNK_BASES + k - 1
    words = R.synth_seq(seed, n_bases, REPEAT_EVERY, first_word=0, n_words=(sample + 31) // 32 + 1)
    t0 = time.perf_counter()
    if kind == "reference":
        r = P.count(words, n_chunks, CHUNK_BASES + k - 1, CHUNK_BASES // 32, k, threads=threads, want_rows=False)
    else:
        r = R.count_query(words, n_chunks, CHUNK_BASES + k - 1, CHUNK_BASES // 32, k, faithful=True,
                          threads=threads, want_rows=False, expected_keys=sample)
    dt = time.perf_counter() - t0
    assert r.total == n_chunks * CHUNK_BASES
    return r.total / dt / 1e9, dt, r.stats, kind, sample


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    n_bases, k, seed, desc = WORKLOADS[args.workload]
    threads = max(1, min(os.cpu_count() or 1, 64))
    # calibrate the per-step sample so that the whole run ends within a few minutes
    reads = READS.get(args.workload)
    rate, dt, _, kind, _ = cpu_reference_rate(n_bases, k, seed, threads, 4_000_000, reads)
    # the rate on a small (cache-friendly) sample is optimistic: keep 40 % of the time budget
    budget_s = 0.4 * max(0.5, min(4.0, 100.0 / max(1, args.steps + args.warmup)))
    sample = int(max(1_000_000, min(16_000_000, rate * 1e9 * budget_s)))
    for _ in range(args.warmup):
        cpu_reference_rate(n_bases, k, seed, threads, sample, reads)
    wall, total = 0.0, 0
    for _ in range(args.steps):
        _, dt, stats, kind, sample_used = cpu_reference_rate(n_bases, k, seed, threads, sample, reads)
        wall += dt
        total += stats[0]
    value = total / wall / 1e9
    what = ("the reference's own dna.c (unmodified, PostgreSQL API shim): generate_kmers SRF loop + hash aggregate "
            "through kmer_hash/kmer_eq" if kind == "reference" else
            "oracle port: faithful per-k-mer decode/validate/encode (dna.c:803-825) + kmer_hash/kmer_eq aggregate")
    sample_desc = (f"first {sample_used} bases of the workload per step, {what}, {threads} threads "
                   "(Postgres itself would run this serially: generate_kmers is PARALLEL UNSAFE)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gkmer/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": desc, "n_bases": n_bases, "k": k, "seed": seed, "repeat_every": REPEAT_EVERY,
                   "sample_bases_per_step": sample_used},
        "cpu_baseline": {"value": value, "unit": "Gkmer/s", "cores": threads, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": "Gkmer/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ================================= our arm =====================================
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dnagpu
    from dnagpu.distributed import (GpuEngine, PeerExchange, count_sharded, count_sharded_fused, count_sharded_peer,
                                    reads_shard_of, shard_of)

    rank, world, local = dist_env()
    n_bases, k, seed, desc = WORKLOADS[args.workload]
    reads = READS.get(args.workload)
    if args.n_bases:
        n_bases = args.n_bases
        if reads:
            reads = dict(reads, n_reads=max(world, n_bases // reads["bases"]))
            n_bases = reads["n_reads"] * reads["bases"]
    where = {"prefix": reads["prefix"], "pattern": reads["pattern"]} if reads else {}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    peak, peak_src = measured_peaks()

    with torch.cuda.stream(stream):
        ctx = dnagpu.Context(local, torch_stream=True)
        if reads:
            first, starts = reads_shard_of(reads["n_reads"], world, rank)  # first read, reads of this rank
            seq = ctx.synth_reads(first, starts, reads["bases"], reads["stride"], seed, REPEAT_EVERY)
            n_rows_total = reads["n_reads"] * (reads["bases"] - k + 1)
        else:
            first, starts = shard_of(n_bases, k, world, rank)
            seq = ctx.synth_range(n_bases, seed, REPEAT_EVERY, first, starts, k)
            n_rows_total = n_bases - k + 1
        n_rows_local = seq.kmer_count(k)
        n_words_local = seq.n_words
        # host copy of this rank's packed words, pinned (the dna value a backend would hold)
        host = torch.empty(n_words_local + 2, dtype=torch.int64, pin_memory=True)
        host.zero_()
        ctx.synchronize()
        rc = ctx.lib.dnagpu_seq_download(ctx.handle, seq.handle, C.c_void_p(host.data_ptr()), n_words_local)
        assert rc == 0
        local_bases = starts * reads["bases"] if reads else min(n_bases - first, starts + k - 1)

        def barrier():
            if world > 1:
                dist.barrier(device_ids=[local])
            torch.cuda.synchronize(dev)

        engine = GpuEngine(ctx) if world > 1 else None
        xbuf = {}
        px = None
        if world > 1 and args.exchange == "peer":
            try:  # collective: raises on every rank or on none
                px = PeerExchange(ctx, world, rank, int(n_rows_total / world * 1.15) + (1 << 20))
            except RuntimeError as e:
                if rank == 0:
                    print(f"bench: {e}; falling back to --exchange fused", file=sys.stderr)
                args.exchange = "fused"

        def count_resident(s):
            """One pass of the hot path with the packed words resident in HBM -> (total, distinct, unique)."""
            if world == 1:
                st, _ = ctx.count(s, k, table=False, load_factor=args.load_factor, **where)
                return st.total, st.distinct, st.unique
            if args.exchange == "peer":
                return count_sharded_peer(ctx, s, k, n_rows_total, world, rank, px, **where)
            if args.exchange == "fused":
                return count_sharded_fused(ctx, s, k, n_rows_total, world, rank, xbuf, chunks=args.chunks, **where)
            return count_sharded(engine, s, k, world, load_factor=args.load_factor, **where)

        def count_e2e():
            """The reference-facing call: host words in, aggregates out (H2D and D2H inside)."""
            hp = C.c_void_p(host.data_ptr())
            if world == 1 and reads:
                st = ctx.count_reads_ptr(hp, starts, reads["bases"], reads["stride"], k, **where)
                return st.total, st.distinct, st.unique
            if world == 1:
                st = ctx.count_kmers_ptr(hp, local_bases, k)
                return st.total, st.distinct, st.unique
            if reads:
                s = ctx.upload_reads_ptr(hp, starts, reads["bases"], reads["stride"])
            else:
                s = ctx.upload_words(hp, local_bases)
                s.set_start_limit(starts)
            r = count_resident(s)
            s.free()
            return r

        # ---- value leg: resident inputs, CUDA events ----
        for _ in range(args.warmup):
            stats = count_resident(seq)
        ctx.profile(True)
        ctx.profile_reset()
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            stats = count_resident(seq)
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        kernels = ctx.profile_dump()
        ctx.profile(False)
        ctx.profile_reset()

        # ---- e2e leg: host buffers through the C ABI, wall clock around synchronous calls ----
        for _ in range(min(args.warmup, 2)):
            count_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            stats_e2e = count_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0

        # ---- extraction GB/s (the second half of the metric), timed on its own ----
        extract = None
        if not args.no_extract and not reads:
            xs_bases = min(local_bases, 1_000_000_000)
            xs = ctx.synth_range(n_bases, seed, REPEAT_EVERY, first, max(0, xs_bases - k + 1), k)
            xr = xs.kmer_count(k)
            xout = torch.empty(xr + 2, dtype=torch.int64, device=dev)
            for _ in range(3):
                ctx.extract(xs, k, out=xout)
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            x0.record(stream)
            xn = 10
            for _ in range(xn):
                ctx.extract(xs, k, out=xout)
            x1.record(stream)
            torch.cuda.synchronize(dev)
            xms = x0.elapsed_time(x1) / xn
            xbytes = 0.25 * xs_bases + 8.0 * xr
            extract = {"rows": xr, "ms": xms, "gkmer_s": xr / xms / 1e6, "gbs": xbytes / xms / 1e6,
                       "frac": xbytes / xms / 1e6 / peak, "bytes_per_row": 8.25,
                       "note": "output (8 B/row) larger than L2; 10 back-to-back launches"}
            # ... and the WHERE clause over those materialised rows (kmer ^@ 'AC': 1/16 of them pass), rows kept in order
            from dnagpu import _where
            fw, _keep = _where("AC", None)
            fres = torch.empty(xr // 8 + 2, dtype=torch.int64, device=dev)
            fn = C.c_uint64()

            def scan():
                rc_ = ctx.lib.dnagpu_filter_keys(ctx.handle, xout.data_ptr(), xr, k, C.byref(fw), fres.data_ptr(),
                                                 fres.numel(), C.byref(fn))
                assert rc_ == 0, rc_
            for _ in range(2):
                scan()
            torch.cuda.synchronize(dev)
            x0.record(stream)
            for _ in range(xn):
                scan()
            x1.record(stream)
            torch.cuda.synchronize(dev)
            fms = x0.elapsed_time(x1) / xn
            fbytes = 2 * 8.0 * xr + 8.0 * fn.value  # the column is read twice (count per tile, ordered write)
            extract["filter_keys"] = {"predicate": "kmer ^@ 'AC'", "rows": xr, "matches": int(fn.value), "ms": fms,
                                      "gkmer_s": xr / fms / 1e6, "gbs": fbytes / fms / 1e6,
                                      "frac": fbytes / fms / 1e6 / peak,
                                      "note": "bytes = 2 passes over the 8 B rows + rows written"}
            del xout, fres
            xs.free()

    # ---- reduce over ranks ----
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    assert reads or stats[0] == n_rows_total, (stats, n_rows_total)
    assert tuple(stats) == tuple(stats_e2e), (stats, stats_e2e)
    # parity gate: the CPU oracle's answer for this exact workload (tests/golden/big_expected.json, produced by
    # tests/golden/make_golden_big.py).  A run whose result differs prints NO line.
    gold = None
    if not args.n_bases:
        try:
            with open(os.path.join(ROOT, "tests", "golden", "big_expected.json")) as f:
                gold = json.load(f).get({"c4": "c4", "c2": "c2", "c5": "c5_k31", "c3": "c3"}[args.workload])
        except Exception:
            gold = None
        if gold is not None and tuple(stats) != (gold["total"], gold["distinct"], gold["unique"]):
            if rank == 0:
                print(f"bench: result {tuple(stats)} differs from the oracle's "
                      f"{(gold['total'], gold['distinct'], gold['unique'])}: no line printed", file=sys.stderr)
            return 1

    if rank == 0:
        value = n_rows_total * args.steps / (ms * 1e-3) / 1e9
        e2e_value = n_rows_total * args.e2e_steps / e2e_s / 1e9
        launches = sum(v["launches"] for v in kernels.values())
        # algorithmic bytes per launch of every kernel of the count pipeline:
        # rows = k-mers one launch handles on this rank; base reads are 0.25 B/base
        rows_r = stats[0] / world          # keys that reach the partition / count stages (after WHERE)
        base_b = 8.0 * n_words_local       # the packed words one launch reads (0.25 B/base)
        listed = "filter_collect" in kernels and kernels["filter_collect"]["launches"] > 0
        ALG = {
            "count_hash": base_b + 16.0 * rows_r, "count_hash_keys": 8.0 * rows_r + 16.0 * rows_r,
            "count_dense": base_b + 4.0 * rows_r, "count_dense_smem": base_b,
            "part_hist": base_b if world == 1 and not listed else 8.0 * rows_r,
            "part_scatter": (base_b if world == 1 and not listed else 8.0 * rows_r) + 8.0 * rows_r,
            # a WHERE clause is evaluated once into a key list: packed words in, matching rows out
            "filter_collect": base_b + 8.0 * rows_r,
            "part_hist2": 8.0 * rows_r, "part_scatter2": 16.0 * rows_r, "count_buckets": 8.0 * rows_r,
            "partition_count": base_b, "partition_write": base_b + 8.0 * rows_r,
        }
        timed = {n: v for n, v in kernels.items() if n in ALG and v["launches"]}
        roofline = None
        if timed:
            dom = max(timed, key=lambda n: timed[n]["ms"])
            d = timed[dom]
            per_launch_ms = d["ms"] / d["launches"]
            achieved = ALG[dom] / per_launch_ms / 1e6
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get(f"{args.workload}:n{world}:{dom}")
            except Exception:
                pass
            pipe_bytes = sum(ALG[n] * v["launches"] / args.steps for n, v in timed.items())
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": ALG[dom], "ms_per_launch": per_launch_ms,
                        "share_of_step": d["ms"] / ms,
                        "pipeline": {"algorithmic_bytes_per_step": pipe_bytes,
                                     "bytes_per_kmer": pipe_bytes / rows_r,
                                     "achieved": pipe_bytes / (ms / args.steps) / 1e6,
                                     "frac": pipe_bytes / (ms / args.steps) / 1e6 / peak,
                                     "note": "all kernels of one step on this rank / whole step time"},
                        "per_kernel": {n: {"ms_per_launch": v["ms"] / v["launches"],
                                           "achieved": ALG[n] / (v["ms"] / v["launches"]) / 1e6,
                                           "frac": ALG[n] / (v["ms"] / v["launches"]) / 1e6 / peak}
                                       for n, v in timed.items()}}
        threads = max(1, min(os.cpu_count() or 1, 64))
        cv, cdt, cstats, ckind, csample = cpu_reference_rate(n_bases, k, seed, threads, args.cpu_sample, reads)
        cpu = {"value": cv, "unit": "Gkmer/s", "cores": threads, "kind": ckind,
               "sample": f"first {csample} bases of the workload ({cdt:.1f} s), " +
                         ("the reference's own dna.c (unmodified, PostgreSQL API shim) driven like the executor: "
                          "generate_kmers SRF loop + hash aggregate through kmer_hash/kmer_eq"
                          if ckind == "reference" else
                          "oracle port: faithful per-k-mer decode/validate/encode of dna.c:803-825 + "
                          "kmer_hash/kmer_eq hash aggregate")}
        line = {
            "metric": METRIC, "value": value, "unit": "Gkmer/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "n_bases": n_bases, "k": k, "seed": seed,
                       "unit_of_value": "k-mers generated and tested per second" if reads else "k-mers counted per second",
                       "repeat_every": REPEAT_EVERY, "block_bases": 1024,
                       "l2": "inputs (packed words + hash table) larger than L2; no flush needed",
                       "parallelism": "single GPU" if world == 1 else
                       f"{world} base-range shards, k-mers routed to owner GPUs by hash; exchange={args.exchange} (" +
                       {"peer": "level-1 scatter kernel stores into the owners' memory over NVLink, no all-to-all",
                        "fused": "level-1 layout + one NCCL all-to-all",
                        "routed": "dnagpu_partition + NCCL all-to-all + dnagpu_count_keys"}[args.exchange] + ")",
                       "load_factor": args.load_factor or 0.5},
            "result": {"total": stats[0], "distinct": stats[1], "unique": stats[2],
                       "oracle": "equal to tests/golden/big_expected.json (CPU oracle at the full size)"
                       if gold is not None else "not compared (size overridden or no golden entry)"},
            "e2e": {"value": e2e_value, "unit": "Gkmer/s", "steps": args.e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / args.e2e_steps,
                    "h2d_bytes_per_step": int(8 * reads["n_reads"] * reads["stride"]) if reads
                    else int(8 * ((n_bases + 31) // 32)), "d2h_bytes_per_step": 24 * world,
                    "api": ("dnagpu_count_reads(ctx, host_words, n_reads, 150, 5, k, &where, &stats, NULL)" if reads else
                            "dnagpu_count_kmers(ctx, host_words, n_bases, k, NULL, &stats, NULL)")},
            "gpu_launches": launches, "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu,
            "extract": extract, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if px is not None:
        px.close()
    seq.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--n-bases", type=int, default=0, help="override the workload size (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--load-factor", type=float, default=0.0)
    ap.add_argument("--cpu-sample", type=int, default=16_000_000)
    ap.add_argument("--no-extract", action="store_true")
    ap.add_argument("--chunks", type=int, default=4, help="N > 1: digit sub-ranges the exchange is pipelined in")
    ap.add_argument("--exchange", default="peer", choices=["peer", "fused", "routed"],
                    help="N > 1: peer = scatter kernel stores into the owners' memory over NVLink (no all-to-all); "
                         "fused = level-1 layout + NCCL all-to-all; routed = separate dnagpu_partition pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
