/*
 * bench.c -- stand-alone C harness over the libdnagpu C ABI (no Python, no torch).
 *
 *   dnagpu_bench [--bases N] [--k K] [--seed S] [--steps T] [--reads R --read-bases B]
 *                [--prefix ACGT] [--pattern IUPAC] [--host] [--gpus G]
 *
 * Generates the synthetic workload on the device (include/dnagpu_synth.h), runs the
 * GROUP BY kmer query T times and prints one JSON line with the per-kernel CUDA-event
 * times reported by dnagpu_profile_*.  --host times the host-buffer entry point
 * (dnagpu_count_kmers: H2D + count + D2H of the aggregates) instead.  --gpus G (with --host) runs that
 * entry point on a multi-GPU context (dnagpu_create_multi over devices 0 .. G-1): shards uploaded on all
 * PCIe links at once, every GPU counts the k-mers it owns out of the whole sequence.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/dnagpu.h"
#include "../../include/dnagpu_synth.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != DNAGPU_OK) {                                                       \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, dnagpu_last_error(ctx)); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    uint64_t n_bases = 100000000ull, seed = 2, n_reads = 0;
    uint32_t read_bases = 150;
    int k = 21, steps = 5, host = 0, gpus = 1, i;
    int devices[16];
    const char *prefix = NULL, *pattern = NULL;
    dnagpu_ctx *ctx = NULL;
    dnagpu_seq *seq = NULL;
    dnagpu_where where;
    dnagpu_stats st = {0, 0, 0};
    char prof[8192], name[128];
    int sms = 0;
    double t0, dt;

    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--bases") && i + 1 < argc) n_bases = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "--k") && i + 1 < argc) k = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--reads") && i + 1 < argc) n_reads = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "--read-bases") && i + 1 < argc) read_bases = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--prefix") && i + 1 < argc) prefix = argv[++i];
        else if (!strcmp(argv[i], "--pattern") && i + 1 < argc) pattern = argv[++i];
        else if (!strcmp(argv[i], "--host")) host = 1;
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    if (gpus < 1 || gpus > 16 || (gpus > 1 && !host)) {
        fprintf(stderr, "--gpus takes 1..16 and needs --host (the multi-GPU entry point is dnagpu_count_kmers)\n");
        return 2;
    }
    for (i = 0; i < gpus; i++) devices[i] = i;
    if ((gpus > 1 ? dnagpu_create_multi(&ctx, devices, gpus) : dnagpu_create(&ctx, 0)) != DNAGPU_OK) {
        fprintf(stderr, "dnagpu_create: %s\n", dnagpu_last_error(NULL));
        return 1;
    }
    CHECK(dnagpu_device_info(ctx, name, sizeof name, &sms, NULL, NULL));
    memset(&where, 0, sizeof where);
    if (prefix) { /* kmer_make of the literal (dna.c:397-420) */
        size_t j, len = strlen(prefix);
        for (j = 0; j < len && j < 32; j++) {
            uint64_t c = prefix[j] == 'T' ? 1 : prefix[j] == 'C' ? 2 : prefix[j] == 'G' ? 3 : 0;
            where.prefix_bits |= c << (2 * j);
        }
        where.prefix_len = (int32_t)len;
    }
    where.qkmer = pattern;
    if (n_reads)
        CHECK(dnagpu_seq_synth_reads(ctx, 0, n_reads, read_bases, (read_bases + 31) / 32, seed,
                                     DNAGPU_SYNTH_REPEAT_EVERY, &seq));
    else
        CHECK(dnagpu_seq_synth(ctx, n_bases, seed, DNAGPU_SYNTH_REPEAT_EVERY, &seq));

    if (host) {
        uint64_t n_words = dnagpu_seq_words(seq);
        void *pinned = NULL;
        CHECK(dnagpu_host_alloc(ctx, &pinned, (n_words + 2) * 8));
        CHECK(dnagpu_seq_download(ctx, seq, (uint64_t *)pinned, n_words));
        CHECK(dnagpu_count_kmers(ctx, (const uint64_t *)pinned, n_bases, k, &where, &st, NULL)); /* warm-up */
        t0 = now_s();
        for (i = 0; i < steps; i++)
            CHECK(dnagpu_count_kmers(ctx, (const uint64_t *)pinned, n_bases, k, &where, &st, NULL));
        dt = now_s() - t0;
        dnagpu_host_free(ctx, pinned);
        prof[0] = '{', prof[1] = '}', prof[2] = 0;
    } else {
        CHECK(dnagpu_count(ctx, seq, k, &where, NULL, &st, NULL)); /* warm-up */
        CHECK(dnagpu_profile_enable(ctx, 1));
        CHECK(dnagpu_profile_reset(ctx));
        t0 = now_s();
        for (i = 0; i < steps; i++) CHECK(dnagpu_count(ctx, seq, k, &where, NULL, &st, NULL));
        CHECK(dnagpu_synchronize(ctx));
        dt = now_s() - t0;
        CHECK(dnagpu_profile_dump(ctx, prof, sizeof prof));
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"gpus\": %d, \"k\": %d, \"rows\": %llu, \"steps\": %d, \"entry\": \"%s\", "
           "\"ms_per_step\": %.4f, \"gkmer_s\": %.3f, \"total\": %llu, \"distinct\": %llu, \"unique\": %llu, "
           "\"kernels_ms_total\": %s}\n",
           name, sms, dnagpu_device_count(ctx), k, (unsigned long long)dnagpu_seq_kmer_count(seq, k), steps,
           host ? "dnagpu_count_kmers (host words)" : "dnagpu_count (device resident)", 1e3 * dt / steps,
           (double)dnagpu_seq_kmer_count(seq, k) * steps / dt / 1e9, (unsigned long long)st.total,
           (unsigned long long)st.distinct, (unsigned long long)st.unique, prof);
    dnagpu_seq_free(seq);
    dnagpu_destroy(ctx);
    return 0;
}
