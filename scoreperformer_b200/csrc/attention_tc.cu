// tcgen05 / TMEM / TMA forward of the fused multi-query attention (same math as attention.cu, see there for the reference
// citations: modules/transformer/attend.py:58-126, attention.py:139-197).
//
// MQA makes the 4 query heads of a position share K and V, so one CTA stacks 32 positions x 4 heads into a 128-row tile:
//   S[128 x 128 keys] = Q[128 x 64] K^T      one tcgen05.mma chain (M=128, N=128, K=64), accumulator in TMEM
//   P = softmax-numerators(S)                 4 warps, ONE ROW PER THREAD (TMEM lane == row): no shuffles for row max / sum
//   O_part[128 x 64] = P[128 x 128] V         P goes registers -> bf16 -> 128B-swizzled smem (A operand), V is the MN-major B
// Warp 4 owns all asynchronous work (TMA loads of Q / K / V tiles through 3-D tensor maps, MMA issue, commits); warps 0-3
// own the arithmetic.  Two CTAs fit per SM (96 KB smem, 256 TMEM columns each) and overlap each other's MMA / softmax phases.
#include "attention_tc.cuh"

namespace {
using namespace attn_tc;


constexpr int SQ_OFF = 0;
constexpr int SK_OFF = 16384;           // 2 stages x 16 KB
constexpr int SV_OFF = 49152;
constexpr int SP_OFF = 65536;           // 2 K-atoms x 16 KB
constexpr int BAR_OFF = 98304;
constexpr int TC_SMEM_BYTES = BAR_OFF + 256 + 1024;

struct TcParams {
    const uint32_t* mask_bits;   // [B, words_per_row] key validity bits (bit j%32 of word j/32), or null
    int words_per_row;
    const float* logslopes;
    __nv_bfloat16* out;
    int ld_out;
    float* lse;                  // [B, H, T] base-2 units
    float* edist;                // [B, H, T] E_i[|i-j|] under the undropped attention weights (backward's slope gradient), or null
    int B, T;
    float scale;
    int causal;
    uint64_t seed;
    const uint64_t* rng_offset;
    uint32_t thr32;
    float keep_scale;
    int kcol, vcol;
};

__global__ void __launch_bounds__(160, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* k_empty = bars + 3;     // [2]
    uint64_t* v_full = bars + 5;
    uint64_t* v_empty = bars + 6;
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_full = bars + 9;
    uint64_t* o_full = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * QP, b = blockIdx.y;
    const int T = p.T;
    const int n_tiles_all = (T + TKEY - 1) / TKEY;
    const int n_kt = p.causal ? min(n_tiles_all, (min(q0 + QP, T) - 1) / TKEY + 1) : n_tiles_all;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
        mbar_init(&k_empty[0], 1); mbar_init(&k_empty[1], 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(s_free, 4);
        mbar_init(p_full, 4); mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;            // columns [0, 128)
    const uint32_t tmem_o = tmem_base + 128;      // columns [128, 192)

    if (warp == 4) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, TKEY, false, false);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DH, false, true);
            const uint32_t sq = smem_u32(smem + SQ_OFF), sk = smem_u32(smem + SK_OFF), sv = smem_u32(smem + SV_OFF),
                           sp = smem_u32(smem + SP_OFF);
            mbar_arrive_expect_tx(q_full, NH * QP * 128);
#pragma unroll
            for (int h = 0; h < NH; ++h) tma_load_3d(smem + SQ_OFF + h * (QP * 128), &tmQ, q_full, h * DH, q0, b);
            mbar_arrive_expect_tx(&k_full[0], TKEY * 128);
            tma_load_3d(smem + SK_OFF, &tmKV, &k_full[0], p.kcol, 0, b);
            mbar_arrive_expect_tx(v_full, TKEY * 128);
            tma_load_3d(smem + SV_OFF, &tmKV, v_full, p.vcol, 0, b);
            for (int kt = 0; kt < n_kt; ++kt) {
                const int st = kt & 1;
                if (kt + 1 < n_kt) {          // prefetch the next K tile into the other stage
                    const int s2 = (kt + 1) & 1;
                    mbar_wait(&k_empty[s2], ((((kt + 1) >> 1) & 1) ^ 1));
                    mbar_arrive_expect_tx(&k_full[s2], TKEY * 128);
                    tma_load_3d(smem + SK_OFF + s2 * (TKEY * 128), &tmKV, &k_full[s2], p.kcol, (kt + 1) * TKEY, b);
                }
                if (kt == 0) mbar_wait(q_full, 0);
                mbar_wait(&k_full[st], (kt >> 1) & 1);
                mbar_wait(s_free, (kt & 1) ^ 1);          // softmax warps have read S of the previous tile out of TMEM
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16(tmem_s, umma_smem_desc_sw128(sq + k * 32, 0, 1024),
                              umma_smem_desc_sw128(sk + st * (TKEY * 128) + k * 32, 0, 1024), idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(&k_empty[st]);
                mbar_wait(p_full, kt & 1);                // P tile written (and O_part of the previous tile consumed)
                mbar_wait(v_full, kt & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TKEY / 16; ++ks)
                    umma_bf16(tmem_o, umma_smem_desc_sw128(sp + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                              umma_smem_desc_sw128(sv + ks * 2048, 8192, 1024), idesc_pv, ks > 0 ? 1u : 0u);
                umma_commit(o_full);
                umma_commit(v_empty);
                if (kt + 1 < n_kt) {
                    mbar_wait(v_empty, kt & 1);
                    mbar_arrive_expect_tx(v_full, TKEY * 128);
                    tma_load_3d(smem + SV_OFF, &tmKV, v_full, p.vcol, (kt + 1) * TKEY, b);
                }
            }
        }
    } else {
        // ---------------- softmax warps: warp == head, lane == position, thread == one row of the 128-row tile
        const int h = warp;
        const int i = q0 + lane;
        const int r = warp * QP + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const float slope = __expf(p.logslopes[h]) * LOG2E;
        const float scale2 = p.scale * LOG2E;
        DropParams drop;
        drop.seedmix = drop_seedmix(p.seed, p.rng_offset);
        drop.thr32 = p.thr32;
        drop.quarter_t = (uint32_t)((T + 3) >> 2);
        drop.keep_scale = p.keep_scale;
        const uint32_t drop_row = drop_row_base(drop, (uint32_t)((b * NH + h) * T + i));
        const bool drop_on = p.thr32 != 0;
        uint8_t* sP = smem + SP_OFF;

        float o_reg[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) o_reg[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, d_run = 0.f;      // running max, sum of numerators, sum of numerators * |i-j|

        for (int kt = 0; kt < n_kt; ++kt) {
            // validity bits of this tile's four 32-key chunks (key padding, sequence tail, causal limit for THIS row)
            uint32_t vbits[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j0 = kt * TKEY + c * 32;
                uint32_t bits = p.mask_bits != nullptr ? p.mask_bits[(size_t)b * p.words_per_row + (j0 >> 5)] : 0xffffffffu;
                if (j0 + 32 > T) bits &= (T > j0) ? ((1u << (T - j0)) - 1u) : 0u;
                if (p.causal) {
                    const int lim = i - j0;     // keys j0 .. j0+lim allowed
                    bits &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
                }
                vbits[c] = bits;
            }
            mbar_wait(s_full, kt & 1);
            tc_fence_after();
            // ---- pass 1: row maximum of this tile (scores in base-2 units: s*scale2 - slope*|i-j|)
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, v);
                tmem_ld_wait();
                const float dbase = (float)(i - (kt * TKEY + c * 32));
                const uint32_t bits = vbits[c];
                if (bits == 0xffffffffu) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        tmax = fmaxf(tmax, fmaf(-slope, fabsf(dbase - (float)jj), __uint_as_float(v[jj]) * scale2));
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const float x = fmaf(-slope, fabsf(dbase - (float)jj), __uint_as_float(v[jj]) * scale2);
                        tmax = fmaxf(tmax, ((bits >> jj) & 1u) ? x : -INFINITY);
                    }
                }
            }
            const float m_new = fmaxf(m_run, tmax);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            const float corr = exp2f(m_run - m_use);
            m_run = m_new;
            // ---- pass 2: numerators, dropout, bf16 pack into the swizzled A-operand tile
            float rsum = 0.f, dsum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, v);
                tmem_ld_wait();
                if (c == 3) {                      // S is fully in registers: let the next QK^T overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
                const int j0 = kt * TKEY + c * 32;
                const float dbase = (float)(i - j0);
                const uint32_t bits = vbits[c];
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const float dist = fabsf(dbase - (float)jj);
                    const float x = fmaf(-slope, dist, __uint_as_float(v[jj]) * scale2);
                    float e = exp2f(x - m_use);
                    if (bits != 0xffffffffu) e = ((bits >> jj) & 1u) ? e : 0.f;
                    rsum += e;
                    dsum = fmaf(e, dist, dsum);
                    v[jj] = __float_as_uint(e);
                }
                if (drop_on) {
                    const uint32_t base = drop_row + (uint32_t)(j0 >> 2) * DROP_K;
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 4) {
                        const uint32_t qh = drop_quad(base + (uint32_t)(jj >> 2) * DROP_K);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            v[jj + e] = drop_keep(qh, e, drop.thr32) ? __float_as_uint(__uint_as_float(v[jj + e]) * drop.keep_scale) : 0u;
                    }
                }
                uint8_t* dst_row = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(v[qd * 8 + 0]), __uint_as_float(v[qd * 8 + 1]));
                    u.y = pack_bf16x2(__uint_as_float(v[qd * 8 + 2]), __uint_as_float(v[qd * 8 + 3]));
                    u.z = pack_bf16x2(__uint_as_float(v[qd * 8 + 4]), __uint_as_float(v[qd * 8 + 5]));
                    u.w = pack_bf16x2(__uint_as_float(v[qd * 8 + 6]), __uint_as_float(v[qd * 8 + 7]));
                    const int c16 = (c & 1) * 4 + qd;
                    *reinterpret_cast<uint4*>(dst_row + ((c16 ^ (r & 7)) << 4)) = u;
                }
            }
            l_run = l_run * corr + rsum;
            d_run = d_run * corr + dsum;
            fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
#pragma unroll
            for (int d = 0; d < DH; ++d) o_reg[d] *= corr;
            mbar_wait(o_full, kt & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_o + lane_addr + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 32; ++d) o_reg[half * 32 + d] += __uint_as_float(v[d]);
            }
            tc_fence_before();
        }
        if (i < T) {
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            __nv_bfloat16* dst = p.out + ((size_t)b * T + i) * p.ld_out + h * DH;
#pragma unroll
            for (int d = 0; d < DH; d += 8) {
                uint4 u;
                u.x = pack_bf16x2(o_reg[d] * inv, o_reg[d + 1] * inv);
                u.y = pack_bf16x2(o_reg[d + 2] * inv, o_reg[d + 3] * inv);
                u.z = pack_bf16x2(o_reg[d + 4] * inv, o_reg[d + 5] * inv);
                u.w = pack_bf16x2(o_reg[d + 6] * inv, o_reg[d + 7] * inv);
                *reinterpret_cast<uint4*>(dst + d) = u;
            }
            if (p.lse != nullptr) p.lse[((size_t)b * NH + h) * T + i] = l_run > 0.f ? (m_run + log2f(l_run)) : INFINITY;
            if (p.edist != nullptr) p.edist[((size_t)b * NH + h) * T + i] = d_run * inv;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<256>(tmem_base);
}

// key_mask bytes [B, T] -> bit words [B, ceil(T/32)]
__global__ void mask_bits_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits, int B, int T, int words) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= B * words) return;
    const int b = w / words, j = (w % words) * 32 + lane;
    const bool ok = j < T && mask[(size_t)b * T + j] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) bits[w] = word;
}

}  // namespace

// tcgen05 forward.  Same contract as spb_attention_fwd plus `mask_bits_scratch` (uint32 [B, ceil(T/32)], only touched when
// key_mask != NULL; the backward reads it again) and `edist` (fp32 [B, H, T] or NULL): E_i[|i-j|] under the attention weights,
// which spb_attention_bwd_tc needs for the slope gradient.  Requires H == 4, dim_head == 64, ld and ld_out multiples of 8.
extern "C" int spb_attention_fwd_tc(const void* qkv, int ld, const uint8_t* key_mask, uint32_t* mask_bits_scratch, const float* logslopes,
                                    void* out, int ld_out, float* lse, float* edist, int B, int T, int H, int dim_head, int causal,
                                    float dropout_p, uint64_t seed, const uint64_t* rng_offset, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return SPB_OK;
    SPB_CHECK_ARG(qkv && logslopes && out, "spb_attention_fwd_tc: null pointer");
    SPB_CHECK_ARG(H == NH && dim_head == DH, "spb_attention_fwd_tc: needs 4 heads of dim 64 (got %d x %d)", H, dim_head);
    SPB_CHECK_ARG(ld % 8 == 0 && ld >= H * DH + 2 * DH && ld_out % 8 == 0, "spb_attention_fwd_tc: bad leading dims");
    SPB_CHECK_ARG(key_mask == nullptr || mask_bits_scratch != nullptr, "spb_attention_fwd_tc: mask_bits_scratch required with key_mask");
    SPB_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "spb_attention_fwd_tc: dropout_p must be in [0,1)");
    const int words = ceil_div(T, 32);
    if (key_mask != nullptr) {
        mask_bits_kernel<<<ceil_div(B * words, 8), 256, 0, stream>>>(key_mask, mask_bits_scratch, B, T, words);
        SPB_CHECK_LAUNCH();
    }
    CUtensorMap tmQ, tmKV;
    int rc = spb_make_tmap_bf16_3d(&tmQ, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)T * ld * 2, DH, QP);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_3d(&tmKV, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)T * ld * 2, DH, TKEY);
    if (rc != SPB_OK) return rc;
    TcParams p;
    p.mask_bits = key_mask != nullptr ? mask_bits_scratch : nullptr;
    p.words_per_row = words;
    p.logslopes = logslopes;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ld_out = ld_out;
    p.lse = lse;
    p.edist = edist;
    p.B = B; p.T = T;
    p.scale = 1.f / sqrtf((float)dim_head);
    p.causal = causal;
    p.seed = seed;
    p.rng_offset = rng_offset;
    p.thr32 = host_drop_thr32(dropout_p);     // spb_attention_bwd_tc evaluates the same mask function (attention_tc.cuh)
    p.keep_scale = 1.f / (1.f - dropout_p);
    p.kcol = H * DH;
    p.vcol = H * DH + DH;
    SPB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attn_fwd_tc_kernel<<<dim3(ceil_div(T, QP), B), 160, TC_SMEM_BYTES, stream>>>(tmQ, tmKV, p);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
