// Optimiser step of the training path on the FLAT parameter / gradient buffers:
// gradient clipping by global norm + AdamW + refresh of the bf16 weight shadow, one pass over memory.
//
// Reference: experiments/optimizers.py:151-169 (`clip_grad_norm_(max_norm)` then `torch.optim.AdamW.step()`), i.e.
//   coef = min(1, max_norm / (||g|| + 1e-6));  g <- coef * g
//   p <- p * (1 - lr*wd);  m <- b1*m + (1-b1)*g;  v <- b2*v + (1-b2)*g*g
//   p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The 236 parameter tensors are views of one buffer (train_step.py), so the whole model is ONE launch that reads p, g, m, v
// and writes p, m, v (+ 2 bytes of bf16 shadow): 30 bytes per parameter, HBM-bound (11.6 M parameters ~ 350 MB ~ 55 us).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  __nv_bfloat16* __restrict__ shadow, int64_t n4, const float* __restrict__ grad_norm, float grad_scale, float max_norm,
                  float lr, const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float weight_decay,
                  const int64_t* __restrict__ step) {
    if (lr_dev != nullptr) lr = *lr_dev;       // schedulers write the rate on the device: a captured graph never has to be rebuilt
    const float t = (float)(*step);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    float gs = grad_scale;
    if (grad_norm != nullptr && max_norm > 0.f) {
        const float total = *grad_norm * grad_scale;
        gs *= fminf(max_norm / (total + 1e-6f), 1.f);
    }
    const float decay = 1.f - lr * weight_decay;
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = reinterpret_cast<float*>(&pp);
        const float* ga = reinterpret_cast<const float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = ga[j] * gs;
            ma[j] = beta1 * ma[j] + (1.f - beta1) * gj;
            va[j] = beta2 * va[j] + (1.f - beta2) * gj * gj;
            const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
            pa[j] = pa[j] * decay - step_size * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow != nullptr) {
            uint2 o;
            o.x = pack_bf16x2(pa[0], pa[1]);
            o.y = pack_bf16x2(pa[2], pa[3]);
            reinterpret_cast<uint2*>(shadow)[i] = o;
        }
    }
}

// dst_i[0:n_i] += src_i[0:n_i] for up to MULTI_ADD_MAX (dst, src) pairs in one launch (blockIdx.y = pair)
constexpr int MULTI_ADD_MAX = 48;
struct MultiAdd {
    float* dst[MULTI_ADD_MAX];
    const float* src[MULTI_ADD_MAX];
    int n[MULTI_ADD_MAX];
};
__global__ void __launch_bounds__(256) multi_add_kernel(MultiAdd a) {
    const int k = blockIdx.y;
    float* __restrict__ dst = a.dst[k];
    const float* __restrict__ src = a.src[k];
    const int n = a.n[k];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) dst[i] += src[i];
}

// dst_i[0:n_i) = src_i[0:n_i) (bytes) for up to MULTI_ADD_MAX pairs in one launch; 16-byte units when both ends are aligned
struct MultiCopy {
    uint8_t* dst[MULTI_ADD_MAX];
    const uint8_t* src[MULTI_ADD_MAX];
    long long n[MULTI_ADD_MAX];
};
__global__ void __launch_bounds__(256) multi_copy_kernel(MultiCopy a) {
    const int k = blockIdx.y;
    uint8_t* __restrict__ dst = a.dst[k];
    const uint8_t* __restrict__ src = a.src[k];
    const long long n = a.n[k];
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
    const long long n16 = vec ? n / 16 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (long long i = n16 * 16 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) dst[i] = src[i];
}

}  // namespace

// Many small device-to-device copies in one launch: dst_i = src_i (n_i bytes).  dsts / srcs / nbytes are HOST arrays.  (The tensors
// of a collated batch -- tokens, masks, segment ids, labels -- go into the static inputs of the captured step this way.)
extern "C" int spb_multi_copy(void* const* dsts, const void* const* srcs, const long long* nbytes, int n_pairs, cudaStream_t stream) {
    SPB_CHECK_ARG(n_pairs == 0 || (dsts && srcs && nbytes), "spb_multi_copy: null pointer");
    for (int base = 0; base < n_pairs; base += MULTI_ADD_MAX) {
        MultiCopy a;
        const int cnt = n_pairs - base < MULTI_ADD_MAX ? n_pairs - base : MULTI_ADD_MAX;
        long long n_max = 0;
        for (int i = 0; i < MULTI_ADD_MAX; ++i) {
            const int j = i < cnt ? base + i : base;
            a.dst[i] = reinterpret_cast<uint8_t*>(dsts[j]); a.src[i] = reinterpret_cast<const uint8_t*>(srcs[j]);
            a.n[i] = i < cnt ? nbytes[j] : 0;
            if (a.n[i] > n_max) n_max = a.n[i];
        }
        if (n_max == 0) continue;
        long long bx = (n_max / 16 + 255) / 256;
        if (bx < 1) bx = 1;
        if (bx > 128) bx = 128;
        multi_copy_kernel<<<dim3((unsigned)bx, cnt), 256, 0, stream>>>(a);
        SPB_CHECK_LAUNCH();
    }
    return SPB_OK;
}

// Gradient accumulation of many small tensors in one launch: dst_i += src_i (fp32, n_i elements), i < n_pairs.  dsts / srcs / sizes
// are HOST arrays.  (The AdaLN projections of a stack get their gradients from one batched GEMM; this adds each slice into its
// parameter's gradient -- modules/layers.py:31-47 has one nn.Linear per norm.)
extern "C" int spb_multi_add_f32(float* const* dsts, const float* const* srcs, const int* sizes, int n_pairs, cudaStream_t stream) {
    SPB_CHECK_ARG(n_pairs == 0 || (dsts && srcs && sizes), "spb_multi_add_f32: null pointer");
    for (int base = 0; base < n_pairs; base += MULTI_ADD_MAX) {
        MultiAdd a;
        const int cnt = n_pairs - base < MULTI_ADD_MAX ? n_pairs - base : MULTI_ADD_MAX;
        int n_max = 0;
        for (int i = 0; i < MULTI_ADD_MAX; ++i) {
            const int j = i < cnt ? base + i : base;
            a.dst[i] = dsts[j]; a.src[i] = srcs[j]; a.n[i] = i < cnt ? sizes[j] : 0;
            if (a.n[i] > n_max) n_max = a.n[i];
        }
        if (n_max == 0) continue;
        int bx = (n_max + 255) / 256;
        if (bx > 64) bx = 64;
        multi_add_kernel<<<dim3(bx, cnt), 256, 0, stream>>>(a);
        SPB_CHECK_LAUNCH();
    }
    return SPB_OK;
}

// p / g / m / v fp32 [n] (n % 4 == 0, 16-byte aligned); shadow bf16 [n] or null.  grad_norm: device scalar holding the L2 norm of
// g BEFORE grad_scale (null or max_norm <= 0 disables clipping).  step: device int64 holding the 1-based step number t.
// lr_dev: optional device scalar that overrides `lr` (learning-rate schedules under CUDA-graph replay).
extern "C" int spb_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, int64_t n, const float* grad_norm,
                              float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                              const int64_t* step, const float* lr_dev, cudaStream_t stream) {
    if (n <= 0) return SPB_OK;
    SPB_CHECK_ARG(p && g && m && v && step, "spb_adamw_step: null pointer");
    SPB_CHECK_ARG(n % 4 == 0, "spb_adamw_step: n must be a multiple of 4 (flat buffers are padded to 8), got %lld", (long long)n);
    SPB_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(shadow) & 7) == 0,
                  "spb_adamw_step: unaligned buffers");
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)spb_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adamw_flat_kernel<<<(int)blocks, 256, 0, stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow), n4, grad_norm, grad_scale,
                                                       max_norm, lr, lr_dev, beta1, beta2, eps, weight_decay, step);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
