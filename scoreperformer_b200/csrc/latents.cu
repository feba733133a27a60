// Hierarchical MMD-VAE latent kernels: segmented mean-pooling over bar/beat/onset runs, the per-segment
// latent projection, broadcast-back to notes, their backward, and the fused pairwise-RBF MMD loss.
//
// Reference semantics (models/scoreperformer/mmd_transformer.py):
//   :325-342  pooling  (dense one-hot alignment + bmm there; a segmented reduction here)
//   :342      latents_mask = all(pooled != 0)                      (reproduced exactly)
//   :346-347  latents = Linear(pooled) * latents_mask
//   :362-366  embeddings[b,t] = latents[b, segments[b,t]] * mask[b,t]
//   :505-534  MMDLoss: k(x,y) = exp(-||x-y||^2 / d^2), mean over all pairs incl. the diagonal
// The pooled input of level l is cat(hidden*mask, style[:, :w_style]) (hierarchical_with_context, :259-261);
// it is read from its two sources and never materialised.
#include "common.cuh"

namespace {

constexpr int MAX_D = 320;             // 256 + 64 style columns
constexpr int DREGS = MAX_D / 32;      // features per lane

__device__ __forceinline__ float load_feat(const float* hidden, const float* style, int ld_style, size_t tok, int c, int d_hidden,
                                           int d_total, bool valid) {
    if (c >= d_total) return 0.f;
    if (c < d_hidden) return valid ? hidden[tok * d_hidden + c] : 0.f;   // hidden * mask
    return style[tok * ld_style + (c - d_hidden)];                         // style is stored already masked
}

// One warp per (sample, 32 consecutive notes, group of 64 features).  Lane <-> 2 adjacent features; run boundaries come from
// the segment ids held one-per-lane and broadcast with warp shuffles; equal-id runs are accumulated in registers and flushed
// with one fp32 atomic per feature per run (runs are contiguous in real data, so almost every segment is flushed exactly
// once; ids need not be sorted for correctness).  Loads are issued 8 notes ahead of the reduction to hide HBM latency.
__global__ void __launch_bounds__(128)
segpool_sum_kernel(const float* __restrict__ hidden, const float* __restrict__ style, int ld_style, const uint8_t* __restrict__ mask,
                   const int64_t* __restrict__ segments, float* __restrict__ pooled, int* __restrict__ counts, int B, int T, int S,
                   int d_hidden, int d_total) {
    const int lane = threadIdx.x & 31;
    const int chunks_per_sample = ceil_div(T, 32);
    const int n_fg = ceil_div(d_total, 64);
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= B * chunks_per_sample * n_fg) return;
    const int fg = gw % n_fg;
    const int bc = gw / n_fg;
    const int b = bc / chunks_per_sample;
    const int t0 = (bc % chunks_per_sample) * 32;
    const int t_me = t0 + lane;
    long long my_id = -1;
    bool my_valid = false;
    if (t_me < T) {
        my_id = segments != nullptr ? segments[(size_t)b * T + t_me] : (mask[(size_t)b * T + t_me] ? 1 : 0);
        my_valid = mask[(size_t)b * T + t_me] != 0;
        if (my_id < 0 || my_id >= S) my_id = -1;   // out-of-range ids are dropped (reference would raise)
    }
    const int c0 = fg * 64 + lane * 2;
    float acc0 = 0.f, acc1 = 0.f;
    int run_id = -1, run_count = 0;
    const int n_tok = min(32, T - t0);
    auto flush = [&]() {
        if (run_id >= 0 && run_count > 0) {
            float* dst = pooled + ((size_t)b * S + run_id) * MAX_D;
            if (c0 < d_total && acc0 != 0.f) atomicAdd(dst + c0, acc0);
            if (c0 + 1 < d_total && acc1 != 0.f) atomicAdd(dst + c0 + 1, acc1);
            if (lane == 0 && fg == 0) atomicAdd(counts + (size_t)b * S + run_id, run_count);
        }
        acc0 = acc1 = 0.f;
        run_count = 0;
    };
    for (int j0 = 0; j0 < n_tok; j0 += 8) {
        float v0[8], v1[8];
        int ids[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            ids[u] = j < n_tok ? (int)__shfl_sync(0xffffffffu, (int)my_id, j & 31) : -2;
            const bool valid = j < n_tok ? __shfl_sync(0xffffffffu, (int)my_valid, j & 31) != 0 : false;
            const size_t tok = (size_t)b * T + t0 + j;
            v0[u] = (j < n_tok && ids[u] >= 0) ? load_feat(hidden, style, ld_style, tok, c0, d_hidden, d_total, valid) : 0.f;
            v1[u] = (j < n_tok && ids[u] >= 0) ? load_feat(hidden, style, ld_style, tok, c0 + 1, d_hidden, d_total, valid) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u >= n_tok) break;
            if (ids[u] != run_id) {
                flush();
                run_id = ids[u];
            }
            if (ids[u] >= 0) {
                acc0 += v0[u];
                acc1 += v1[u];
                ++run_count;
            }
        }
    }
    flush();
}

// One warp per segment slot: mean, validity, latent = W pooled + bias.
__global__ void __launch_bounds__(128)
seg_latent_kernel(float* __restrict__ pooled, const int* __restrict__ counts, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ latents, uint8_t* __restrict__ lmask, int n_slots, int d_total, int z, int force_valid) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const int cnt = counts[slot];
    if (cnt == 0 && !force_valid) {
        // empty slot (most of the [B, T+4] table in the sync-free layout): its pooled row is all zero, so it is invalid
        // and its latent is zero -- nothing to read or multiply
        if (lane == 0) lmask[slot] = 0;
        for (int o = lane; o < z; o += 32) latents[(size_t)slot * z + o] = 0.f;
        return;
    }
    const float inv = 1.f / (float)max(cnt, 1);
    float v[DREGS];
    bool all_nz = true;
#pragma unroll
    for (int k = 0; k < DREGS; ++k) {
        const int c = k * 32 + lane;
        v[k] = c < d_total ? pooled[(size_t)slot * MAX_D + c] * inv : 0.f;
        if (c < d_total) {
            pooled[(size_t)slot * MAX_D + c] = v[k];          // keep the mean for the backward
            all_nz = all_nz && (v[k] != 0.f);
        }
    }
    const bool valid = force_valid ? true : (__all_sync(0xffffffffu, all_nz) != 0);
    if (lane == 0) lmask[slot] = valid;
    for (int o = 0; o < z; ++o) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < DREGS; ++k) {
            const int c = k * 32 + lane;
            if (c < d_total) s += v[k] * W[(size_t)o * d_total + c];
        }
        s = warp_sum(s);
        if (lane == 0) latents[(size_t)slot * z + o] = valid ? s + bias[o] : 0.f;
    }
}

// style[b,t,col0:col0+z] = latents[b, seg[b,t]] * mask[b,t]
__global__ void seg_broadcast_kernel(const float* __restrict__ latents, const int64_t* __restrict__ segments,
                                     const uint8_t* __restrict__ mask, float* __restrict__ style, int ld_style, int col0, int B, int T,
                                     int S, int z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T * z) return;
    const int o = (int)(i % z);
    const int64_t tok = i / z;
    const int b = (int)(tok / T);
    long long id = segments != nullptr ? segments[tok] : (mask[tok] ? 1 : 0);
    float v = 0.f;
    if (mask[tok] && id >= 0 && id < S) v = latents[((size_t)b * S + id) * z + o];
    style[tok * ld_style + col0 + o] = v;
}

// dlat_sum[b, seg, :] += mask * d_style[b, t, col0:col0+z]  (one thread per (token, o); z <= 32 so traffic is tiny)
__global__ void seg_gather_grad_kernel(const float* __restrict__ d_style, int ld_style, int col0, const int64_t* __restrict__ segments,
                                       const uint8_t* __restrict__ mask, float* __restrict__ dlat, int B, int T, int S, int z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T * z) return;
    const int o = (int)(i % z);
    const int64_t tok = i / z;
    if (!mask[tok]) return;
    const int b = (int)(tok / T);
    long long id = segments != nullptr ? segments[tok] : 1;
    if (id < 0 || id >= S) return;
    const float g = d_style[tok * ld_style + col0 + o];
    if (g != 0.f) atomicAdd(dlat + ((size_t)b * S + id) * z + o, g);
}

// One warp per slot: dlat = lmask * (dlat_sum + dlat_direct) (written back), dpooled = dlat W / max(count,1)
__global__ void __launch_bounds__(128)
seg_latent_bwd_kernel(float* __restrict__ dlat, const float* __restrict__ dlat_direct, const uint8_t* __restrict__ lmask,
                      const int* __restrict__ counts, const float* __restrict__ W, float* __restrict__ dpooled, int n_slots, int d_total,
                      int z) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const bool valid = lmask[slot] != 0;
    if (!valid) {
        // invalid slot: its latent gradient is dropped; dpooled is only ever read for slots that own tokens
        for (int o = lane; o < z; o += 32) dlat[(size_t)slot * z + o] = 0.f;
        if (counts[slot] > 0)
            for (int c = lane; c < d_total; c += 32) dpooled[(size_t)slot * MAX_D + c] = 0.f;
        return;
    }
    const float inv = 1.f / (float)max(counts[slot], 1);
    float acc[DREGS];
#pragma unroll
    for (int k = 0; k < DREGS; ++k) acc[k] = 0.f;
    for (int o = 0; o < z; ++o) {
        float g = 0.f;
        if (lane == 0) {
            if (valid) {
                g = dlat[(size_t)slot * z + o];
                if (dlat_direct != nullptr) g += dlat_direct[(size_t)slot * z + o];
            }
            dlat[(size_t)slot * z + o] = g;
        }
        g = __shfl_sync(0xffffffffu, g, 0);
#pragma unroll
        for (int k = 0; k < DREGS; ++k) {
            const int c = k * 32 + lane;
            if (c < d_total) acc[k] += g * W[(size_t)o * d_total + c];
        }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < DREGS; ++k) {
        const int c = k * 32 + lane;
        if (c < d_total) dpooled[(size_t)slot * MAX_D + c] = acc[k] * inv;
    }
}

// d_hidden[b,t,:] += mask * dpooled[b, seg, :256] ; d_style[b,t,:w_style] += dpooled[b, seg, 256:256+w_style]
// (style inputs were stored already masked, and their own mask is applied again when they were produced).
// One warp per note-tuple, 16-byte read-modify-writes (d_hidden_dim, w_style and ld_style are multiples of 4): the kernel is a
// stream over d_hidden, nothing else.
__global__ void __launch_bounds__(256)
seg_scatter_grad_kernel(const float* __restrict__ dpooled, const int64_t* __restrict__ segments, const uint8_t* __restrict__ mask,
                        float* __restrict__ d_hidden, float* __restrict__ d_style, int ld_style, int B, int T, int S, int d_hidden_dim,
                        int d_total) {
    const int lane = threadIdx.x & 31;
    const int n_tok = B * T;
    const int q_hidden = d_hidden_dim >> 2, q_total = d_total >> 2;
    for (int tok = blockIdx.x * 8 + (threadIdx.x >> 5); tok < n_tok; tok += gridDim.x * 8) {
        const bool live = mask[tok] != 0;
        const long long id = segments != nullptr ? segments[tok] : (live ? 1 : 0);
        if (id < 0 || id >= S) continue;
        const float4* src = reinterpret_cast<const float4*>(dpooled + ((size_t)(tok / T) * S + id) * MAX_D);
        float4* dh = reinterpret_cast<float4*>(d_hidden + (size_t)tok * d_hidden_dim);
        float4* ds = reinterpret_cast<float4*>(d_style + (size_t)tok * ld_style);
        for (int c = lane; c < q_total; c += 32) {
            float4* dst = c < q_hidden ? (live ? dh + c : nullptr) : ds + (c - q_hidden);
            if (dst == nullptr) continue;
            const float4 g = __ldg(src + c);
            float4 v = *dst;
            v.x += g.x; v.y += g.y; v.z += g.z; v.w += g.w;
            *dst = v;
        }
    }
}

// C[m, n] += A[K, m]^T B[K, n] in fp32 (tiny m <= MR, n <= 512, long K): blocks split K; thread t owns columns t and t + 256 for
// ALL rows, so a k step costs two conflict-free B loads and MR / 4 broadcast float4 loads of A for 2 * MR FMAs.
template <int MR>
__global__ void __launch_bounds__(256)
small_gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, float* __restrict__ C, int ldc, int K,
                     int m, int n, int k_per_block) {
    extern __shared__ __align__(16) float sm[];
    float* sA = sm;                       // [32][MR], rows >= m zero
    float* sB = sm + 32 * MR;             // [32][512]
    const int k0 = blockIdx.x * k_per_block, k1 = min(K, k0 + k_per_block);
    const int c0 = threadIdx.x, c1 = threadIdx.x + 256;
    float acc0[MR], acc1[MR];
#pragma unroll
    for (int r = 0; r < MR; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
    for (int kk = k0; kk < k1; kk += 32) {
        const int kt = min(32, k1 - kk);
        __syncthreads();
        // A = dlatents: rows of empty / padded segments are exactly zero (most rows in the sync-free [B, T+4] layout), and a
        // chunk of 32 all-zero rows contributes nothing
        int live = 0;
        for (int i = threadIdx.x; i < 32 * MR; i += 256) {
            const int k = i / MR, r = i % MR;
            const float a = (k < kt && r < m) ? A[(size_t)(kk + k) * lda + r] : 0.f;
            sA[i] = a;
            live |= (a != 0.f);
        }
        if (!__syncthreads_or(live)) continue;
        {
            // all 64 loads of the chunk leave before the first store (a loop of load -> store pairs pays the memory latency per row)
            float b0v[32], b1v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                b0v[k] = (k < kt && c0 < n) ? __ldg(Bm + (size_t)(kk + k) * ldb + c0) : 0.f;
                b1v[k] = (k < kt && c1 < n) ? __ldg(Bm + (size_t)(kk + k) * ldb + c1) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                sB[k * 512 + c0] = b0v[k];
                sB[k * 512 + c1] = b1v[k];
            }
        }
        __syncthreads();
        for (int k = 0; k < kt; ++k) {
            const float b0 = sB[k * 512 + c0], b1 = sB[k * 512 + c1];
            const float4* a4 = reinterpret_cast<const float4*>(sA + k * MR);
#pragma unroll
            for (int r = 0; r < MR; r += 4) {
                const float4 a = a4[r >> 2];
                acc0[r] = fmaf(a.x, b0, acc0[r]);         acc1[r] = fmaf(a.x, b1, acc1[r]);
                acc0[r + 1] = fmaf(a.y, b0, acc0[r + 1]); acc1[r + 1] = fmaf(a.y, b1, acc1[r + 1]);
                acc0[r + 2] = fmaf(a.z, b0, acc0[r + 2]); acc1[r + 2] = fmaf(a.z, b1, acc1[r + 2]);
                acc0[r + 3] = fmaf(a.w, b0, acc0[r + 3]); acc1[r + 3] = fmaf(a.w, b1, acc1[r + 3]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        if (r < m) {
            if (c0 < n && acc0[r] != 0.f) atomicAdd(C + (size_t)r * ldc + c0, acc0[r]);
            if (c1 < n && acc1[r] != 0.f) atomicAdd(C + (size_t)r * ldc + c1, acc1[r]);
        }
    }
}

// column sums: out[n] += sum_k A[k, n]
__global__ void colsum_kernel(const float* __restrict__ A, int lda, float* __restrict__ out, int K, int n, int k_per_block) {
    const int c = threadIdx.x;
    if (c >= n) return;
    const int k0 = blockIdx.x * k_per_block, k1 = min(K, k0 + k_per_block);
    float s = 0.f;
    for (int k = k0; k < k1; ++k) s += A[(size_t)k * lda + c];
    if (s != 0.f) atomicAdd(out + c, s);
}

// ------------------------------------------------------------------------------------------- MMD
// Points x = [z_prior (n_z rows) ; y (n_y rows)], coefficients a = [+1/n_z ; -w/n] with n = sum(w):
//   loss = sum_ij a_i a_j k(x_i, x_j),   d loss / d y_i = -(4 a_i / d^2) sum_j a_j k_ij (y_i - x_j)
constexpr int MMD_MAX_D = 32;

__global__ void mmd_coef_kernel(const uint8_t* __restrict__ w, int n_z, int n_y, float* __restrict__ coef) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n_y; i += blockDim.x) s += w[i] ? 1.f : 0.f;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const float n = red[0];
    for (int i = threadIdx.x; i < n_z + n_y; i += blockDim.x)
        coef[i] = i < n_z ? 1.f / (float)n_z : (w[i - n_z] && n > 0.f ? -1.f / n : 0.f);
}

template <int D>
__global__ void __launch_bounds__(128)
mmd_pair_kernel(const float* __restrict__ zp, const float* __restrict__ y, const float* __restrict__ coef, int n_z, int n_y,
                int j_per_block, float* __restrict__ loss, float* __restrict__ grad_y) {
    __shared__ float sx[64][D + 1];
    __shared__ float sa[64];
    const int m = n_z + n_y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool active = i < m;
    float xi[D], gi[D];
    float ai = 0.f;
    if (active) {
        const float* src = i < n_z ? zp + (size_t)i * D : y + (size_t)(i - n_z) * D;
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = src[k];
        ai = coef[i];
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) gi[k] = 0.f;
    float ri = 0.f;
    const float inv_d2 = 1.f / (float)(D * D);
    const int j0 = blockIdx.y * j_per_block, j1 = min(m, j0 + j_per_block);
    for (int jt = j0; jt < j1; jt += 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * D; e += 128) {
            const int r = e / D, k = e % D, j = jt + r;
            float v = 0.f;
            if (j < j1) v = j < n_z ? zp[(size_t)j * D + k] : y[(size_t)(j - n_z) * D + k];
            sx[r][k] = v;
        }
        if (threadIdx.x < 64) sa[threadIdx.x] = (jt + threadIdx.x) < j1 ? coef[jt + threadIdx.x] : 0.f;
        __syncthreads();
        if (active && ai != 0.f) {
            for (int r = 0; r < 64; ++r) {
                const float aj = sa[r];
                if (aj == 0.f) continue;
                float d2 = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) { const float df = xi[k] - sx[r][k]; d2 += df * df; }
                const float kv = aj * __expf(-d2 * inv_d2);
                ri += kv;
#pragma unroll
                for (int k = 0; k < D; ++k) gi[k] += kv * (xi[k] - sx[r][k]);
            }
        }
    }
    // loss partial: block reduce of a_i * r_i
    float part = active ? ai * ri : 0.f;
    part = warp_sum(part);
    __shared__ float red[4];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(loss, red[0] + red[1] + red[2] + red[3]);
    if (active && i >= n_z && ai != 0.f && grad_y != nullptr) {
        const float f = -4.f * ai * inv_d2;
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(grad_y + (size_t)(i - n_z) * D + k, f * gi[k]);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------- C-ABI
// Forward of one latent level.  Buffers: pooled fp32 [B*S, 320] and counts int32 [B*S] must be ZEROED by the caller;
// outputs latents fp32 [B*S, z], lmask uint8 [B*S]; style fp32 [B*T, ld_style] receives columns [col0, col0+z).
// segments == NULL selects aggregate mode 'mean' (S must be 2: slot 1 = the sample, slot 0 = padding).
extern "C" int spb_latent_level_fwd(const float* hidden, const float* style_in, int ld_style, const uint8_t* mask,
                                    const int64_t* segments, const float* W, const float* bias, float* pooled, int* counts,
                                    float* latents, uint8_t* lmask, float* style_out, int col0, int B, int T, int S, int d_hidden,
                                    int w_style, int z, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return SPB_OK;
    SPB_CHECK_ARG(hidden && mask && W && bias && pooled && counts && latents && lmask && style_out, "spb_latent_level_fwd: null pointer");
    const int d_total = d_hidden + w_style;
    SPB_CHECK_ARG(d_total <= MAX_D && z <= 64 && S >= 1, "spb_latent_level_fwd: d_total=%d (max %d), z=%d", d_total, MAX_D, z);
    SPB_CHECK_ARG(segments != nullptr || S == 2, "spb_latent_level_fwd: mode 'mean' expects S == 2");
    const int warps = B * ceil_div(T, 32) * ceil_div(d_total, 64);
    segpool_sum_kernel<<<ceil_div(warps, 4), 128, 0, stream>>>(hidden, style_in, ld_style, mask, segments, pooled, counts, B, T, S, d_hidden, d_total);
    SPB_CHECK_LAUNCH();
    seg_latent_kernel<<<ceil_div(B * S, 4), 128, 0, stream>>>(pooled, counts, W, bias, latents, lmask, B * S, d_total, z, segments == nullptr ? 1 : 0);
    SPB_CHECK_LAUNCH();
    const int64_t n = (int64_t)B * T * z;
    seg_broadcast_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(latents, segments, mask, style_out, ld_style, col0, B, T, S, z);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// Backward of one latent level.  dlat fp32 [B*S, z] must be ZEROED by the caller (it returns the effective latent
// gradient); dlat_direct (MMD + deadpan terms) may be NULL; dpooled fp32 [B*S, 320] is scratch.
// d_hidden [B*T, d_hidden] and d_style [B*T, ld_style] are accumulated into; dW [z, d_total] / dbias [z] are accumulated into.
extern "C" int spb_latent_level_bwd(float* d_style, int ld_style, int col0, const float* dlat_direct, const uint8_t* mask,
                                    const int64_t* segments, const float* W, const float* pooled, const int* counts,
                                    const uint8_t* lmask, float* dlat, float* dpooled, float* d_hidden, float* dW, float* dbias, int B,
                                    int T, int S, int d_hidden_dim, int w_style, int z, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return SPB_OK;
    SPB_CHECK_ARG(d_style && mask && W && pooled && counts && lmask && dlat && dpooled && d_hidden && dW && dbias, "spb_latent_level_bwd: null pointer");
    const int d_total = d_hidden_dim + w_style;
    SPB_CHECK_ARG(d_total <= MAX_D && z <= 32, "spb_latent_level_bwd: d_total=%d z=%d out of range", d_total, z);
    const int64_t n = (int64_t)B * T * z;
    seg_gather_grad_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(d_style, ld_style, col0, segments, mask, dlat, B, T, S, z);
    SPB_CHECK_LAUNCH();
    seg_latent_bwd_kernel<<<ceil_div(B * S, 4), 128, 0, stream>>>(dlat, dlat_direct, lmask, counts, W, dpooled, B * S, d_total, z);
    SPB_CHECK_LAUNCH();
    // dW[z, d_total] += dlat^T pooled ; dbias += colsum(dlat)
    const int K = B * S;
    int k_per_block = ceil_div(K, 2 * spb_num_sms());
    if (k_per_block < 32) k_per_block = 32;
#define SMALL_GEMM_CASE(MR)                                                                                                          \
    {                                                                                                                                \
        constexpr int smem = 32 * (MR + 512) * (int)sizeof(float);                                                                   \
        SPB_CHECK_CUDA(cudaFuncSetAttribute(small_gemm_tn_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));           \
        small_gemm_tn_kernel<MR><<<ceil_div(K, k_per_block), 256, smem, stream>>>(dlat, z, pooled, MAX_D, dW, d_total, K, z, d_total, \
                                                                                  k_per_block);                                      \
    }
    if (z <= 4) SMALL_GEMM_CASE(4)
    else if (z <= 8) SMALL_GEMM_CASE(8)
    else SMALL_GEMM_CASE(32)
#undef SMALL_GEMM_CASE
    SPB_CHECK_LAUNCH();
    colsum_kernel<<<ceil_div(K, k_per_block), 64, 0, stream>>>(dlat, z, dbias, K, z, k_per_block);
    SPB_CHECK_LAUNCH();
    SPB_CHECK_ARG(d_hidden_dim % 4 == 0 && w_style % 4 == 0 && ld_style % 4 == 0 && (reinterpret_cast<uintptr_t>(d_style) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(d_hidden) & 15) == 0,
                  "spb_latent_level_bwd: widths must be multiples of 4 and gradients 16-byte aligned");
    {
        const int want = ceil_div(B * T, 8), cap = 8 * spb_num_sms();
        seg_scatter_grad_kernel<<<want < cap ? want : cap, 256, 0, stream>>>(dpooled, segments, mask, d_hidden, d_style, ld_style, B, T, S, d_hidden_dim, d_total);
    }
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// loss (fp32 scalar, ZEROED by caller) = MMD(z_prior, y[w]);  grad_y fp32 [n_y, d] (ZEROED by caller) = d loss / d y.
// coef is fp32 [n_z + n_y] scratch.  d in {4, 8, 20, 32} (latent dims of the recipes); other d <= 32 use the 32 path padded by caller.
extern "C" int spb_mmd_fwd_bwd(const float* z_prior, const float* y, const uint8_t* w, int n_z, int n_y, int d, float* coef, float* loss,
                               float* grad_y, cudaStream_t stream) {
    SPB_CHECK_ARG(z_prior && y && w && coef && loss, "spb_mmd_fwd_bwd: null pointer");
    SPB_CHECK_ARG(n_z > 0 && n_y > 0, "spb_mmd_fwd_bwd: empty inputs");
    mmd_coef_kernel<<<1, 256, 0, stream>>>(w, n_z, n_y, coef);
    SPB_CHECK_LAUNCH();
    const int m = n_z + n_y;
    int splits = ceil_div(2 * spb_num_sms(), ceil_div(m, 128));
    if (splits < 1) splits = 1;
    int j_per_block = ceil_div(ceil_div(m, splits), 64) * 64;
    splits = ceil_div(m, j_per_block);
    dim3 grid(ceil_div(m, 128), splits);
#define MMD_CASE(D) case D: mmd_pair_kernel<D><<<grid, 128, 0, stream>>>(z_prior, y, coef, n_z, n_y, j_per_block, loss, grad_y); break;
    switch (d) {
        MMD_CASE(4) MMD_CASE(8) MMD_CASE(16) MMD_CASE(20) MMD_CASE(32)
        default:
            spb_set_error("spb_mmd_fwd_bwd: latent dim %d not compiled (4, 8, 16, 20, 32)", d);
            return SPB_ERR_ARG;
    }
#undef MMD_CASE
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
