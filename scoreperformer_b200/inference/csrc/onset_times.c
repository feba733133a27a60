/* Onset recurrence of the SPMuple2 messenger (see include/spb200_host.h).  Plain C, float64 throughout; every expression keeps the
 * operand order of the numpy statement in scoreperformer_b200/inference/messengers.py (which keeps the reference's), and sums use
 * numpy's pairwise scheme, so results are bit-identical to the Python path.  Compile with -ffp-contract=off. */
#include "spb200_host.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>

int spb_host_abi_version(void) { return 1; }

/* numpy's float64 add-reduction over a contiguous array (umath loops: pairwise sum, blocks of 128, eight accumulators) */
static double np_sum(const double* a, ptrdiff_t n) {
    if (n < 8) {
        double r = 0.;
        for (ptrdiff_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        ptrdiff_t i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    ptrdiff_t half = n / 2;
    half -= half % 8;
    return np_sum(a, half) + np_sum(a + half, n - half);
}

/* index of the table entry nearest to v; ties go to the upper neighbour */
static int closest(const double* table, int n, double v) {
    int lo = 0, hi = n;                                  /* lower bound: first entry >= v */
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (table[mid] < v) lo = mid + 1; else hi = mid;
    }
    const int up = lo < n ? lo : n - 1, down = lo > 0 ? lo - 1 : 0;
    return (lo == n || fabs(v - table[down]) < fabs(v - table[up])) ? down : up;
}

/* tempo re-estimated from the onsets performed so far: the rows of `pairs` before the current one that lie at least `min_dist`
 * seconds back and inside `window` seconds (widened to the last `min_onsets` within four windows) vote with a weight that falls
 * with their distance in time */
static double local_tempo(const double* pairs, int n_pairs, double min_dist, double window, int min_onsets, double scale,
                          const double* table, int n_table, int quantize, int* idx, double* buf) {
    const double* here = pairs + 3 * (size_t)(n_pairs - 1);
    const double now = here[1];
    const int n_past = n_pairs - 1;
    int n_far = 0;
    for (int i = 0; i < n_past; ++i)
        if (pairs[3 * (size_t)i + 1] <= now - min_dist) idx[n_far++] = i;
    if (n_far == 0)
        for (int i = 0; i < n_past; ++i) idx[n_far++] = i;
    int* win = idx + n_past;                             /* second half of the scratch */
    int n_win = 0;
    for (int i = 0; i < n_far; ++i)
        if (pairs[3 * (size_t)idx[i] + 1] >= now - window) win[n_win++] = idx[i];
    if (n_win < min_onsets) {
        n_win = 0;
        const int from = n_far - min_onsets > 0 ? n_far - min_onsets : 0;
        for (int i = from; i < n_far; ++i)
            if (pairs[3 * (size_t)idx[i] + 1] >= now - 4 * window) win[n_win++] = idx[i];
    }
    if (n_win == 0) {
        for (int i = 0; i < n_far; ++i) win[i] = idx[i];
        n_win = n_far;
    }
    double* dt = buf;
    double* local = buf + n_win;
    double* w = buf + 2 * (size_t)n_win;
    double dt_max = -INFINITY;
    for (int i = 0; i < n_win; ++i) {
        const double* p = pairs + 3 * (size_t)win[i];
        dt[i] = here[1] - p[1];
        local[i] = (here[0] - p[0]) / dt[i] * scale;
        if (dt[i] > dt_max || i == 0) dt_max = dt[i];
    }
    for (int i = 0; i < n_win; ++i) w[i] = 1 - dt[i] / (dt_max + 0.01);
    const double total = np_sum(w, n_win);
    for (int i = 0; i < n_win; ++i) w[i] = w[i] / total * local[i];
    double tempo = np_sum(w, n_win);
    if (!(tempo > table[0])) tempo = table[0];
    if (quantize) tempo = table[closest(table, n_table, tempo)];
    return tempo;
}

int spb_host_onset_times(int n, const double* ticks, const double* durations, const double* note_bpm, const double* rel_dev,
                         const double* rel_held, const uint8_t* performed, const int64_t* order, int n_groups,
                         const int64_t* group_start, double* tempos, int n_tempos, double* pairs, int n_pairs, double scale,
                         double initial_tempo, int from_tokens, int re_estimate, double min_onset_dist, double tempo_window,
                         int min_onsets, int quantize, const double* tempo_table, int n_table, double* on, double* off,
                         int* out_n_tempos, int* out_n_pairs, int* resumed_first) {
    if (n < 0 || n_groups < 0 || n_tempos < 1 || n_pairs < 1 || !ticks || !order || !group_start || !tempos || !pairs || !on || !off ||
        !out_n_tempos || !out_n_pairs || !resumed_first || (re_estimate && (!tempo_table || n_table < 1)))
        return -1;
    const size_t cap = (size_t)n_pairs + (size_t)n_groups + 1;
    int* idx = (int*)malloc(2 * cap * sizeof(int));
    double* buf = (double*)malloc((3 * cap + (size_t)n + 1) * sizeof(double));
    if (!idx || !buf) { free(idx); free(buf); return -2; }
    double* tmp = buf + 3 * cap;                         /* gathered values of one onset */

    double bpm = tempos[3 * (size_t)(n_tempos - 1)];
    double last_tick = pairs[3 * (size_t)(n_pairs - 1)], last_time = pairs[3 * (size_t)(n_pairs - 1) + 1],
           last_n = pairs[3 * (size_t)(n_pairs - 1) + 2];
    int first = 1;
    *resumed_first = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int64_t* members = order + group_start[g];
        const int count = (int)(group_start[g + 1] - group_start[g]);
        int n_live = 0;
        for (int i = 0; i < count; ++i) n_live += performed[members[i]] != 0;
        if (n_live == 0) continue;
        const double tick = ticks[members[0]];
        const int resumed = tick == tempos[3 * (size_t)(n_tempos - 1) + 1] && tick > 0;
        if (resumed) {
            if (n_pairs < 2 || n_tempos < 2) { free(idx); free(buf); return -3; }
            const double* before = pairs + 3 * (size_t)(n_pairs - 2);
            last_tick = before[0]; last_time = before[1]; last_n = before[2];
            bpm = tempos[3 * (size_t)(n_tempos - 2)];
            if (first) *resumed_first = 1;
        }
        first = 0;
        if (from_tokens) {
            for (int i = 0; i < count; ++i) tmp[i] = note_bpm[members[i]];
            const double s = np_sum(tmp, count);
            bpm = resumed ? (bpm * last_n + s) / (last_n + count) : s / count;
        }
        const double step = (tick - last_tick) / bpm * scale;
        const double base = last_time + step;
        int k = 0;
        for (int i = 0; i < count; ++i) {
            const int64_t m = members[i];
            const double played = base + rel_dev[m] * step;
            on[m] = played;
            off[m] = played + rel_held[m] * (durations[m] / bpm * scale);
            if (performed[m]) tmp[k++] = played;
        }
        double onset_time;
        double* row;
        if (resumed) {
            row = pairs + 3 * (size_t)(n_pairs - 1);
            onset_time = row[1] * last_n + np_sum(tmp, n_live);
            onset_time /= (last_n + count);
            row[0] = tick; row[1] = onset_time; row[2] = last_n + count;
        } else {
            onset_time = np_sum(tmp, n_live) / n_live;
            row = pairs + 3 * (size_t)n_pairs++;
            row[0] = tick; row[1] = onset_time; row[2] = count;
        }
        if (re_estimate) {
            if (onset_time < 2 * min_onset_dist) bpm = initial_tempo;
            else bpm = local_tempo(pairs, n_pairs, min_onset_dist, tempo_window, min_onsets, scale, tempo_table, n_table, quantize, idx, buf);
        }
        double* trow = resumed ? tempos + 3 * (size_t)(n_tempos - 1) : tempos + 3 * (size_t)n_tempos++;
        trow[0] = bpm; trow[1] = tick; trow[2] = onset_time;
        last_tick = row[0]; last_time = row[1]; last_n = row[2];
    }
    *out_n_tempos = n_tempos;
    *out_n_pairs = n_pairs;
    free(idx);
    free(buf);
    return 0;
}
