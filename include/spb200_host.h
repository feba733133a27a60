/* scoreperformer_b200 -- host-side C ABI of the rendering path (no CUDA, no Python types).
 *
 * One entry point: the onset recurrence of the SPMuple2 messenger, i.e. the tempo tracking that turns rendered note tuples into
 * performed times.  It replaces the per-onset Python / numpy loop of the reference's
 *   scoreperformer/inference/messengers.py:258-328      (SPMuple2Messenger.tokens_to_messages, loop over score onsets)
 *   scoreperformer/data/tokenizers/spmuple/spmuple2.py:548-593   (filter_onsets_in_window, compute_local_tempo)
 *   scoreperformer/utils/functions.py find_closest
 * and reproduces their float64 results bit for bit (numpy's pairwise summation included).  Built by
 * scoreperformer_b200/inference/native.py with gcc into scoreperformer_b200/inference/csrc/libspb200_host.so.
 */
#ifndef SPB200_HOST_H
#define SPB200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int spb_host_abi_version(void);

/* Notes are given as float64 arrays of length n (score ticks, durations in ticks, Tempo-token values, relative onset deviations,
 * relative performed durations) and a byte mask of performed notes.  `order` lists the notes grouped by score tick in increasing
 * tick order (stable), `group_start[g] .. group_start[g + 1]` delimits group g (n_groups + 1 entries).
 * `tempos` / `pairs` are row-major [cap, 3] buffers whose first n_tempos / n_pairs rows hold the state so far -- rows
 * (tempo, tick, time) and (tick, time, notes averaged) -- and need room for n_groups more rows each.
 * Writes on / off times of every note of a performed onset (others stay untouched), appends to tempos / pairs, stores the new row
 * counts, and sets *resumed_first when the first performed onset continued the last onset of the incoming state (the caller then
 * mirrors the overwritten last rows into the arrays it was handed, as the reference does).
 * Returns 0, or a negative value on invalid arguments / allocation failure. */
int spb_host_onset_times(int n, const double* ticks, const double* durations, const double* note_bpm, const double* rel_dev,
                         const double* rel_held, const uint8_t* performed, const int64_t* order, int n_groups,
                         const int64_t* group_start, double* tempos, int n_tempos, double* pairs, int n_pairs, double scale,
                         double initial_tempo, int from_tokens, int re_estimate, double min_onset_dist, double tempo_window,
                         int min_onsets, int quantize, const double* tempo_table, int n_table, double* on, double* off,
                         int* out_n_tempos, int* out_n_pairs, int* resumed_first);

#ifdef __cplusplus
}
#endif
#endif
