/*
 * cpu_bench.c — ORACLE (test infrastructure): times the CPU restatement of seqhasher.go:241-259 as its own
 * process, record-sharded over OpenMP threads, on the same synthetic reads the GPU benchmark uses
 * (splitmix64 generator of csrc/kernels.cu synth_fill_kernel).  bench.py launches it for the `cpu_baseline`
 * object and the `--impl reference` arm; a standalone process is used because OpenMP teams created inside the
 * Python interpreter of this image stay on one CPU.
 *
 * usage: cpu_bench <n_records> <read_len | lo-hi> <seed> <lower_permille> <flags> <threads> <accel> <steps> <algo>[,<algo>..] [n_permille]
 * read_len "lo-hi": record i has length lo + ((splitmix64((seed ^ LEN_SALT) + i * GOLDEN) >> 32) * (hi - lo + 1) >> 32)
 * (bench.py builds the same offsets with numpy).  n_permille of the bases become 'N' (config C4).
 * prints one JSON line: seqs/s (best step), per-step ms, and a 64-bit XOR-fold checksum of every digest.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle.h"

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + t.tv_nsec * 1e-9;
}

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: %s n len seed lower_permille flags threads accel steps algos\n", argv[0]);
        return 2;
    }
    const size_t n = strtoull(argv[1], NULL, 10);
    char *dash = NULL;
    const size_t L = strtoull(argv[2], &dash, 10);
    const size_t Lhi = (dash && *dash == '-') ? strtoull(dash + 1, NULL, 10) : L;
    const unsigned n_permille = argc > 10 ? (unsigned)atoi(argv[10]) : 0;
    const uint64_t seed = strtoull(argv[3], NULL, 0);
    const unsigned lower = (unsigned)atoi(argv[4]);
    const int flags = atoi(argv[5]), threads = atoi(argv[6]), accel = atoi(argv[7]), steps = atoi(argv[8]);
    int algos[SHO_N_ALGOS], n_algos = 0;
    char *list = strdup(argv[9]);
    for (char *tok = strtok(list, ","); tok && n_algos < SHO_N_ALGOS; tok = strtok(NULL, ",")) {
        int id = sho_algo_id(tok);
        if (id < 0) { fprintf(stderr, "unknown algo %s\n", tok); return 2; }
        algos[n_algos++] = id;
    }
    int64_t *offs = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    if (!offs) return 3;
    offs[0] = 0;
    for (size_t i = 0; i < n; i++) {
        size_t len = L;
        if (Lhi > L) {
            const uint64_t z = splitmix64((seed ^ 0x6c656e67746873ULL) + (uint64_t)i * 0x9E3779B97F4A7C15ULL);
            len = L + (size_t)(((z >> 32) * (uint64_t)(Lhi - L + 1)) >> 32);
        }
        offs[i + 1] = offs[i] + (int64_t)len;
    }
    const size_t total = (size_t)offs[n];
    uint8_t *res = (uint8_t *)malloc(total + 64);
    if (!res) return 3;
#pragma omp parallel for schedule(static)
    for (long long j = 0; j < (long long)total; j++) {
        const uint64_t z = splitmix64(seed + (uint64_t)j * 0x9E3779B97F4A7C15ULL);
        uint8_t c = (uint8_t)"ACGT"[z & 3];
        if ((((z >> 32) & 0xFFFFF) * 1000 >> 20) < n_permille) c = 'N';
        if ((((z >> 8) & 0xFFFFF) * 1000 >> 20) < lower) c |= 0x20;
        res[j] = c;
    }
    uint8_t *dig[SHO_N_ALGOS];
    for (int k = 0; k < n_algos; k++) dig[k] = (uint8_t *)malloc(n * sho_digest_size(algos[k]) + 8);
    double best = 1e30, sum = 0;
    for (int s = 0; s < steps + 1; s++) {   /* step 0 warms up (page faults, thread team) */
        const double t0 = now();
        if (sho_batch(res, offs, n, algos, n_algos, flags, dig, NULL, threads, accel) != 0) return 4;
        const double dt = now() - t0;
        if (s) { sum += dt; if (dt < best) best = dt; }
    }
    uint64_t chk = 0;
    for (int k = 0; k < n_algos; k++) {
        const size_t bytes = n * sho_digest_size(algos[k]);
        for (size_t i = 0; i + 8 <= bytes; i += 8) { uint64_t v; memcpy(&v, dig[k] + i, 8); chk ^= v; }
        for (size_t i = bytes & ~(size_t)7; i < bytes; i++) chk ^= (uint64_t)dig[k][i] << (8 * (i & 7));
    }
    printf("{\"records\": %zu, \"read_len\": %zu, \"read_len_hi\": %zu, \"residue_bytes\": %zu, \"threads\": %d, \"accel\": %d, \"accel_available\": %d, \"steps\": %d, "
           "\"ms_per_step\": %.3f, \"best_ms\": %.3f, \"seqs_per_s\": %.1f, \"checksum\": \"%016llx\"}\n",
           n, L, Lhi, total, threads, accel ? sho_accel_available() : 0, sho_accel_available(), steps, sum / steps * 1e3, best * 1e3,
           n / (sum / steps), (unsigned long long)chk);
    return 0;
}
