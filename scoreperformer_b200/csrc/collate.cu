// Device side of the input pipeline: the collated batch crosses PCIe / NVLink-C2C in its natural
// width -- uint16 tokens, int32 segment ids, uint8 direction labels, one length per sequence: 65 bytes per note-tuple instead of
// the 466 bytes of the reference's int64 tensors -- and ONE kernel expands it into everything the model consumes, including the
// MixedLM masking, which is a pure function of the performance tokens:
//   data/collators/performance.py:239-255 (MixedLMPerformanceCollator.mask_sequence)
//       no_mask  = token in mask_ignore_token_ids            dim_mask = field in mask_ignore_token_dims
//       masked   = MASK where !no_mask && !dim_mask, else token
//       labels   = token where !no_mask (&& !dim_mask when label_pad_ignored_dims), else label_pad_token_id
//   data/collators/score_performance.py:186-234 (MixedLMScorePerformanceCollator.__call__): masks from the sequence lengths.
// Integer work, HBM-bound: 65 B in, 466 B out per note-tuple.
#include "common.cuh"

namespace {

struct UnpackParams {
    const uint16_t* perf;        // [n, Fp]
    const uint16_t* score;       // [n, Fs] or null
    const int32_t* segs;         // [3, n] bars | beats | onsets, or null
    const uint8_t* dirs;         // [n, Fd] or null
    const int32_t* perf_len;     // [B]
    const int32_t* score_len;    // [B] or null (= perf_len)
    int64_t* o_perf;             // [n, Fp]
    int64_t* o_masked;           // [n, Fp] or null
    int64_t* o_labels;           // [n, Fp] or null
    int64_t* o_score;            // [n, Fs]
    int64_t* o_segs;             // [3, n]
    int64_t* o_dirs;             // [n, Fd]
    uint8_t* o_perf_mask;        // [n]
    uint8_t* o_score_mask;       // [n]
    int B, T, Fp, Fs, Fd;
    uint32_t ignore_dims;        // bit f set: field f is never masked / never labelled
    uint32_t ignore_ids;         // bit v set (v < 32): token id v is never masked / labelled (PAD, MASK, SOS, EOS)
    int mask_token;
    long long label_pad;
    int label_pad_ignored_dims;
};

__global__ void __launch_bounds__(256)
unpack_batch_kernel(UnpackParams p) {
    const long long n = (long long)p.B * p.T;
    const int W = p.Fp;                                   // slots per tuple: slot f handles field f of every array that has one
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n * W; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / W;
        const int f = (int)(idx - row * W);
        const int tok = p.perf[idx];
        p.o_perf[idx] = tok;
        if (p.o_masked != nullptr) {
            const bool no_mask = tok < 32 && ((p.ignore_ids >> tok) & 1u);
            const bool dim_ignored = (p.ignore_dims >> f) & 1u;
            p.o_masked[idx] = (!no_mask && !dim_ignored) ? p.mask_token : tok;
            const bool labelled = !no_mask && !(p.label_pad_ignored_dims && dim_ignored);
            p.o_labels[idx] = labelled ? (long long)tok : p.label_pad;
        }
        if (f < p.Fs && p.score != nullptr) p.o_score[row * p.Fs + f] = p.score[row * p.Fs + f];
        if (f < 3 && p.segs != nullptr) p.o_segs[(long long)f * n + row] = p.segs[(long long)f * n + row];
        if (f < p.Fd && p.dirs != nullptr) p.o_dirs[row * p.Fd + f] = p.dirs[row * p.Fd + f];
        if (f == 0) {
            const int b = (int)(row / p.T), t = (int)(row - (long long)b * p.T);
            p.o_perf_mask[row] = t < p.perf_len[b];
            if (p.o_score_mask != nullptr) p.o_score_mask[row] = t < (p.score_len != nullptr ? p.score_len[b] : p.perf_len[b]);
        }
    }
}

}  // namespace

// Expand a packed batch on the device (see the header of this file).  Every output is written in full; `o_masked` / `o_labels`
// may be NULL together (no MixedLM masking), as may the score / segment / direction groups (input and output together).
extern "C" int spb_unpack_batch(const uint16_t* perf, const uint16_t* score, const int32_t* segs, const uint8_t* dirs,
                                const int32_t* perf_len, const int32_t* score_len, int64_t* o_perf, int64_t* o_masked, int64_t* o_labels,
                                int64_t* o_score, int64_t* o_segs, int64_t* o_dirs, uint8_t* o_perf_mask, uint8_t* o_score_mask, int B, int T,
                                int Fp, int Fs, int Fd, uint32_t ignore_dims, uint32_t ignore_ids, int mask_token, long long label_pad,
                                int label_pad_ignored_dims, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return SPB_OK;
    SPB_CHECK_ARG(perf && perf_len && o_perf && o_perf_mask, "spb_unpack_batch: null pointer");
    SPB_CHECK_ARG((o_masked == nullptr) == (o_labels == nullptr), "spb_unpack_batch: masked tokens and labels come together");
    SPB_CHECK_ARG((score == nullptr) == (o_score == nullptr) && (segs == nullptr) == (o_segs == nullptr) && (dirs == nullptr) == (o_dirs == nullptr),
                  "spb_unpack_batch: an input group and its output go together");
    SPB_CHECK_ARG(Fp >= 3 && Fp <= 32 && Fs <= Fp && Fd <= Fp, "spb_unpack_batch: 3 <= Fp <= 32 and Fs, Fd <= Fp (got %d, %d, %d)", Fp, Fs, Fd);
    UnpackParams p;
    p.perf = perf; p.score = score; p.segs = segs; p.dirs = dirs; p.perf_len = perf_len; p.score_len = score_len;
    p.o_perf = o_perf; p.o_masked = o_masked; p.o_labels = o_labels; p.o_score = o_score; p.o_segs = o_segs; p.o_dirs = o_dirs;
    p.o_perf_mask = o_perf_mask; p.o_score_mask = o_score_mask;
    p.B = B; p.T = T; p.Fp = Fp; p.Fs = Fs; p.Fd = Fd;
    p.ignore_dims = ignore_dims; p.ignore_ids = ignore_ids; p.mask_token = mask_token; p.label_pad = label_pad;
    p.label_pad_ignored_dims = label_pad_ignored_dims;
    const long long total = (long long)B * T * Fp;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)spb_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    unpack_batch_kernel<<<(int)blocks, 256, 0, stream>>>(p);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
