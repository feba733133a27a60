/* A C caller of include/spb200_host.h, as a player written in C / C++ would bind it (no Python anywhere in this process).
 * Reads one problem from argv[1] (written by tests/test_inference_host.py), calls spb_host_onset_times, writes the results to argv[2].
 * File layout, little endian: int32 header[12] = {n, n_groups, n_tempos, n_pairs, from_tokens, re_estimate, min_onsets, quantize,
 * n_table, 0, 0, 0}; float64 scalars[5] = {scale, initial_tempo, min_onset_dist, tempo_window, 0}; then the arrays in the order of
 * the function's arguments. */
#include <stdio.h>
#include <stdlib.h>
#include "spb200_host.h"

static void* take(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t* h = take(f, 12 * sizeof(int32_t));
    double* s = take(f, 5 * sizeof(double));
    const int n = h[0], groups = h[1], nt = h[2], np = h[3];
    double* ticks = take(f, n * sizeof(double));
    double* durations = take(f, n * sizeof(double));
    double* note_bpm = take(f, n * sizeof(double));
    double* rel_dev = take(f, n * sizeof(double));
    double* rel_held = take(f, n * sizeof(double));
    uint8_t* performed = take(f, n);
    int64_t* order = take(f, n * sizeof(int64_t));
    int64_t* bounds = take(f, (groups + 1) * sizeof(int64_t));
    double* tempos_in = take(f, 3 * nt * sizeof(double));
    double* pairs_in = take(f, 3 * np * sizeof(double));
    double* table = take(f, h[8] * sizeof(double));
    fclose(f);

    double* tempos = calloc(3 * (size_t)(nt + groups), sizeof(double));
    double* pairs = calloc(3 * (size_t)(np + groups), sizeof(double));
    double* on = calloc(n ? n : 1, sizeof(double));
    double* off = calloc(n ? n : 1, sizeof(double));
    for (int i = 0; i < 3 * nt; ++i) tempos[i] = tempos_in[i];
    for (int i = 0; i < 3 * np; ++i) pairs[i] = pairs_in[i];
    int out_t = 0, out_p = 0, resumed = 0;
    const int rc = spb_host_onset_times(n, ticks, durations, note_bpm, rel_dev, rel_held, performed, order, groups, bounds, tempos, nt, pairs,
                                        np, s[0], s[1], h[4], h[5], s[2], s[3], h[6], h[7], table, h[8], on, off, &out_t, &out_p, &resumed);
    if (rc != 0 || spb_host_abi_version() < 1) { fprintf(stderr, "spb_host_onset_times: %d\n", rc); return 3; }

    FILE* o = fopen(argv[2], "wb");
    int32_t head[4] = {out_t, out_p, resumed, 0};
    fwrite(head, sizeof(int32_t), 4, o);
    fwrite(on, sizeof(double), n, o);
    fwrite(off, sizeof(double), n, o);
    fwrite(tempos, sizeof(double), 3 * (size_t)out_t, o);
    fwrite(pairs, sizeof(double), 3 * (size_t)out_p, o);
    fclose(o);
    return 0;
}
