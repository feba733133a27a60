// Output-side kernels: masked cross-entropy over per-field logits (loss + unscaled dlogits in one pass), the nine
// direction-classifier heads with class-weighted CE (forward + backward fused), and on-device greedy/top-k sampling
// support for rendering.
//
// Reference semantics:
//   per-field CE, ignore_index=-100, mean over labelled positions      models/scoreperformer/wrappers.py:49-59
//   classifier heads: Dropout -> Linear(64, C_g) -> weighted CE         models/classifiers/model.py:74-82, 202-216
#include "common.cuh"

namespace {

// One warp per row.  logits fp32 [n, ld]; writes nll/lse stats and (optionally) dlogits = softmax - onehot (bf16, unscaled;
// rows with an ignored label are zero) and the row argmax.
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ labels, int ld_lab, int V, long long ignore_index,
               float* __restrict__ loss_sum, float* __restrict__ count, __nv_bfloat16* __restrict__ dlogits, int ld_d,
               int* __restrict__ argmax_out, int n_rows) {
    __shared__ float s_loss[8], s_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float my_loss = 0.f, my_cnt = 0.f;
    for (int row = blockIdx.x * 8 + warp; row < n_rows; row += gridDim.x * 8) {
        const float* lr = logits + (size_t)row * ld;
        const long long lab = labels[(size_t)row * ld_lab];
        const bool use = lab != ignore_index && lab >= 0 && lab < V;
        float mx = -INFINITY;
        int amax = 0;
        for (int c = lane; c < V; c += 32) {
            const float v = lr[c];
            if (v > mx) { mx = v; amax = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
            if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
        }
        if (argmax_out != nullptr && lane == 0) argmax_out[row] = amax;
        if (!use && dlogits == nullptr) continue;
        float se = 0.f;
        for (int c = lane; c < V; c += 32) se += __expf(lr[c] - mx);
        se = warp_sum(se);
        const float lse = mx + __logf(se);
        if (use) {
            if (lane == 0) { my_loss += lse - lr[lab]; my_cnt += 1.f; }
        }
        if (dlogits != nullptr) {
            __nv_bfloat16* dr = dlogits + (size_t)row * ld_d;
            const float inv = 1.f / se;
            for (int c = lane; c < ld_d; c += 32) {
                float g = 0.f;
                if (use && c < V) g = __expf(lr[c] - mx) * inv - (c == lab ? 1.f : 0.f);
                dr[c] = __float2bfloat16_rn(g);
            }
        }
    }
    if (lane == 0) { s_loss[warp] = my_loss; s_cnt[warp] = my_cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += s_loss[i]; b += s_cnt[i]; }
        if (b > 0.f) { atomicAdd(loss_sum, a); atomicAdd(count, b); }
    }
}

// ------------------------------------------------------------------------------------------- classifier heads
constexpr int CLF_MAX_HEADS = 16;
constexpr int CLF_MAX_CLASSES = 64;   // total classes over all heads
struct ClfHeads {
    int n_heads;
    int n_classes[CLF_MAX_HEADS];
    int class_off[CLF_MAX_HEADS];
    int total;
};

// One WARP per note: lane k owns input dims k and k+32 (in_dim <= 64); each class logit is a warp-shuffle reduction, so the
// weight reads are conflict-free and the input row is read coalesced.  x fp32 [n, in_dim] (style embeddings, detached).
// Forward (dl == nullptr): per head num[g] += w[y] * nll, den[g] += w[y].
// Backward (dl != nullptr): dl[row, off_g + c] = scale_g * w[y] * (p_c - onehot_c)   (zero for unclassified rows).
__global__ void __launch_bounds__(256)
clf_heads_kernel(const float* __restrict__ x, int ldx, const uint8_t* __restrict__ rowmask, const int64_t* __restrict__ labels,
                 int ld_lab, const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ class_w, ClfHeads hd,
                 float* __restrict__ num, float* __restrict__ den, const float* __restrict__ dlogit_scale, float* __restrict__ dl,
                 int n_rows, int in_dim, uint64_t seed, const uint64_t* __restrict__ rng_offset, uint32_t drop_thresh24,
                 float keep_scale) {
    if (rng_offset != nullptr) seed += *rng_offset * 0x9E3779B97F4A7C15ull;
    extern __shared__ float sm[];
    float* sW = sm;                               // [total][64] (zero padded beyond in_dim)
    float* sB = sW + hd.total * 64;               // [total]
    float* sCW = sB + hd.total;                   // [total]
    __shared__ float s_num[CLF_MAX_HEADS], s_den[CLF_MAX_HEADS];
    for (int i = threadIdx.x; i < hd.total * 64; i += blockDim.x) {
        const int c = i >> 6, k = i & 63;
        sW[i] = k < in_dim ? W[(size_t)c * in_dim + k] : 0.f;
    }
    for (int i = threadIdx.x; i < hd.total; i += blockDim.x) { sB[i] = bias[i]; sCW[i] = class_w[i]; }
    if (threadIdx.x < CLF_MAX_HEADS) { s_num[threadIdx.x] = 0.f; s_den[threadIdx.x] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * warps_per_block) {
        const bool row_on = rowmask[row] != 0;
        if (!row_on) {
            if (dl != nullptr) for (int c = lane; c < hd.total; c += 32) dl[(size_t)row * hd.total + c] = 0.f;
            continue;
        }
        const float x0 = lane < in_dim ? x[(size_t)row * ldx + lane] : 0.f;
        const float x1 = lane + 32 < in_dim ? x[(size_t)row * ldx + lane + 32] : 0.f;
        for (int g = 0; g < hd.n_heads; ++g) {
            const int C = hd.n_classes[g], off = hd.class_off[g];
            const long long y = labels[(size_t)row * ld_lab + g];
            if (y < 0 || y >= C) {
                if (dl != nullptr && lane < C) dl[(size_t)row * hd.total + off + lane] = 0.f;
                continue;
            }
            float a0 = x0, a1 = x1;
            if (drop_thresh24 != 0) {
                // one hash per (row, head, lane): low half decides input dim `lane`, high half dim `lane + 32`
                const uint32_t hsh = spb_pair_hash(spb_seed32(seed), ((uint32_t)row * CLF_MAX_HEADS + (uint32_t)g) * 32u + (uint32_t)lane);
                a0 = spb_keep16(hsh, 0, drop_thresh24 >> 8) ? x0 * keep_scale : 0.f;
                a1 = spb_keep16(hsh, 1, drop_thresh24 >> 8) ? x1 * keep_scale : 0.f;
            }
            // All (<= 16) class dot products of the head at once: every lane forms its partial for each class, then a
            // transposing butterfly (16 shuffles instead of 5 per class) leaves the total of class c in lanes 2c and 2c+1.
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c)
                v[c] = c < C ? a0 * sW[(off + c) * 64 + lane] + a1 * sW[(off + c) * 64 + lane + 32] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool hi = lane & 16;
                const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool hi = lane & 8;
                const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool hi = lane & 4;
                const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            {
                const bool hi = lane & 2;
                const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
                v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            const int cls = lane >> 1;                                  // class held by this lane pair
            const bool owner = (lane & 1) == 0 && cls < C;
            const float my_logit = owner ? v[0] + sB[off + cls] : -INFINITY;
            const float mx = warp_max(my_logit);
            const float e = owner ? __expf(my_logit - mx) : 0.f;
            const float se = warp_sum(e);
            const float lse = mx + __logf(se);
            const float ly = __shfl_sync(0xffffffffu, my_logit, 2 * (int)y);
            const float wy = sCW[off + y];
            if (dl == nullptr) {
                if (lane == 0) {
                    atomicAdd(&s_num[g], wy * (lse - ly));
                    atomicAdd(&s_den[g], wy);
                }
            } else if (owner) {
                dl[(size_t)row * hd.total + off + cls] = dlogit_scale[g] * wy * (e / se - (cls == y ? 1.f : 0.f));
            }
        }
    }
    if (dl == nullptr) {
        __syncthreads();
        if (threadIdx.x < hd.n_heads && s_den[threadIdx.x] != 0.f) {
            atomicAdd(num + threadIdx.x, s_num[threadIdx.x]);
            atomicAdd(den + threadIdx.x, s_den[threadIdx.x]);
        }
    }
}

// dW[off_g + c, k] += sum_rows dl[row, off_g + c] * dropout_g(x[row, k]);  db[c] += sum_rows dl[row, c].
// Each block owns 64 rows; dl and the (per-head re-dropped) inputs are staged in shared memory.
__global__ void __launch_bounds__(256)
clf_wgrad_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dl, ClfHeads hd, float* __restrict__ dW,
                 float* __restrict__ db, int n_rows, int in_dim, uint64_t seed, const uint64_t* __restrict__ rng_offset,
                 uint32_t drop_thresh24, float keep_scale) {
    if (rng_offset != nullptr) seed += *rng_offset * 0x9E3779B97F4A7C15ull;
    __shared__ float sx[64][65];
    __shared__ float sd[64][17];
    const int r0 = blockIdx.x * 64;
    const int nr = min(64, n_rows - r0);
    for (int g = 0; g < hd.n_heads; ++g) {
        const int C = hd.n_classes[g], off = hd.class_off[g];
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 64; i += 256) {
            const int r = i >> 6, k = i & 63;
            float v = 0.f;
            if (r < nr && k < in_dim) {
                v = x[(size_t)(r0 + r) * ldx + k];
                if (drop_thresh24 != 0) {
                    const uint32_t hsh = spb_pair_hash(spb_seed32(seed), ((uint32_t)(r0 + r) * CLF_MAX_HEADS + (uint32_t)g) * 32u + (uint32_t)(k & 31));
                    v = spb_keep16(hsh, k >> 5, drop_thresh24 >> 8) ? v * keep_scale : 0.f;
                }
            }
            sx[r][k] = v;
        }
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            sd[r][c] = (r < nr && c < C) ? dl[(size_t)(r0 + r) * hd.total + off + c] : 0.f;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < C * in_dim; o += 256) {
            const int c = o / in_dim, k = o % in_dim;
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < 64; ++r) s += sd[r][c] * sx[r][k];
            if (s != 0.f) atomicAdd(dW + (size_t)(off + c) * in_dim + k, s);
        }
        if (threadIdx.x < C) {
            float s = 0.f;
            for (int r = 0; r < 64; ++r) s += sd[r][threadIdx.x];
            if (s != 0.f) atomicAdd(db + off + threadIdx.x, s);
        }
    }
}

// logits fp32 [n, total] for API consumers (evaluation / inspection); not used by the training step.
__global__ void clf_logits_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ out, int n_rows, int in_dim, int total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rows * total) return;
    const int c = (int)(i % total);
    const int64_t row = i / total;
    float s = bias[c];
    for (int k = 0; k < in_dim; ++k) s += x[row * ldx + k] * W[(size_t)c * in_dim + k];
    out[i] = s;
}

int fill_heads(ClfHeads& hd, const int* n_classes, int n_heads) {
    SPB_CHECK_ARG(n_heads > 0 && n_heads <= CLF_MAX_HEADS, "classifier: 1..%d heads supported", CLF_MAX_HEADS);
    hd.n_heads = n_heads;
    int off = 0;
    for (int g = 0; g < n_heads; ++g) {
        SPB_CHECK_ARG(n_classes[g] > 0 && n_classes[g] <= 16, "classifier: at most 16 classes per head, got %d", n_classes[g]);
        hd.n_classes[g] = n_classes[g];
        hd.class_off[g] = off;
        off += n_classes[g];
    }
    SPB_CHECK_ARG(off <= CLF_MAX_CLASSES, "classifier: at most %d classes in total", CLF_MAX_CLASSES);
    hd.total = off;
    return SPB_OK;
}

}  // namespace

// loss_sum / count (fp32 scalars) are ACCUMULATED into.  dlogits (bf16 [n, ld_d], ld_d >= V, zero-padded) and
// argmax_out (int32 [n]) are optional.
extern "C" int spb_ce_rows(const float* logits, int ld, const int64_t* labels, int ld_lab, int V, long long ignore_index, float* loss_sum,
                           float* count, void* dlogits, int ld_d, int* argmax_out, int n_rows, cudaStream_t stream) {
    if (n_rows <= 0) return SPB_OK;
    SPB_CHECK_ARG(logits && labels && loss_sum && count && V > 0 && ld >= V, "spb_ce_rows: bad arguments");
    SPB_CHECK_ARG(dlogits == nullptr || ld_d >= V, "spb_ce_rows: ld_d < V");
    int grid = ceil_div(n_rows, 8);
    if (grid > spb_num_sms() * 8) grid = spb_num_sms() * 8;
    ce_rows_kernel<<<grid, 256, 0, stream>>>(logits, ld, labels, ld_lab, V, ignore_index, loss_sum, count,
                                             reinterpret_cast<__nv_bfloat16*>(dlogits), ld_d, argmax_out, n_rows);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// W fp32 [total_classes, in_dim] (heads concatenated), bias / class_w fp32 [total_classes].
// forward:  num/den fp32 [n_heads] accumulated into (loss_g = num/den).
// backward: dlogit_scale fp32 [n_heads] on device (= upstream * loss_weight / n_heads / den_g); dW/db accumulated into;
//           dl_scratch fp32 [n, total_classes].
extern "C" int spb_clf_heads(const float* x, int ldx, const uint8_t* rowmask, const int64_t* labels, int ld_lab, const float* W,
                             const float* bias, const float* class_w, const int* n_classes, int n_heads, float* num, float* den,
                             const float* dlogit_scale, float* dW, float* db, float* dl_scratch, int n_rows, int in_dim, float dropout_p,
                             uint64_t seed, const uint64_t* rng_offset, int backward, cudaStream_t stream) {
    if (n_rows <= 0) return SPB_OK;
    SPB_CHECK_ARG(x && rowmask && labels && W && bias && class_w, "spb_clf_heads: null pointer");
    SPB_CHECK_ARG(in_dim > 0 && in_dim <= 64, "spb_clf_heads: in_dim must be <= 64, got %d", in_dim);
    ClfHeads hd;
    int rc = fill_heads(hd, n_classes, n_heads);
    if (rc != SPB_OK) return rc;
    double t = (double)dropout_p * 16777216.0;
    const uint32_t th = dropout_p > 0.f ? (uint32_t)(t < 1 ? 1 : t) : 0;
    const float ks = 1.f / (1.f - dropout_p);
    int grid = ceil_div(n_rows, 8);
    if (grid > 4 * spb_num_sms()) grid = 4 * spb_num_sms();
    const size_t smem = (size_t)(hd.total * 64 + 2 * hd.total) * sizeof(float);
    if (!backward) {
        SPB_CHECK_ARG(num && den, "spb_clf_heads: forward needs num/den");
        clf_heads_kernel<<<grid, 256, smem, stream>>>(x, ldx, rowmask, labels, ld_lab, W, bias, class_w, hd, num, den, nullptr, nullptr,
                                                      n_rows, in_dim, seed, rng_offset, th, ks);
    } else {
        SPB_CHECK_ARG(dlogit_scale && dW && db && dl_scratch, "spb_clf_heads: backward needs dlogit_scale/dW/db/dl_scratch");
        clf_heads_kernel<<<grid, 256, smem, stream>>>(x, ldx, rowmask, labels, ld_lab, W, bias, class_w, hd, nullptr, nullptr, dlogit_scale,
                                                      dl_scratch, n_rows, in_dim, seed, rng_offset, th, ks);
        SPB_CHECK_LAUNCH();
        clf_wgrad_kernel<<<ceil_div(n_rows, 64), 256, 0, stream>>>(x, ldx, dl_scratch, hd, dW, db, n_rows, in_dim, seed, rng_offset, th, ks);
    }
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_clf_logits(const float* x, int ldx, const float* W, const float* bias, float* out, int n_rows, int in_dim, int total,
                              cudaStream_t stream) {
    if (n_rows <= 0) return SPB_OK;
    SPB_CHECK_ARG(x && W && bias && out, "spb_clf_logits: null pointer");
    const int64_t n = (int64_t)n_rows * total;
    clf_logits_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(x, ldx, W, bias, out, n_rows, in_dim, total);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
