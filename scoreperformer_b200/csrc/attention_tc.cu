// tcgen05 / TMEM / TMA forward of the fused multi-query attention (same math as attention.cu, see there for the reference
// citations: modules/transformer/attend.py:58-126, attention.py:139-197).
//
// MQA makes the 4 query heads of a position share K and V, so 32 positions x 4 heads stack into one 128-row UMMA tile
// (row = head*32 + position).  One CTA per SM owns FOUR such query tiles (128 consecutive positions of one sequence: every K / V
// tile fetched from L2 serves 512 rows) and walks the 64-key tiles they can see:
//   S_q[128 x 64 keys] = Q_q[128 x 64] K^T       accumulator in TMEM (4 x 64 columns)
//   P_q = exp2(S_q - m)  (+ dropout)              one row per thread (TMEM lane == row), bf16 into 128B-swizzled smem
//   O_q[128 x 64]     += P_q V                    accumulates IN TMEM over all key tiles (4 x 64 columns), read once at the end
// 16 element-wise warps = 4 quartets, quartet q owns query tile q (warp == head == TMEM lane quarter), so the four quartets run
// four independent S -> P -> PV pipelines that the one MMA thread (warp 16) serves round-robin, with S of the next key tile
// issued together with PV of this one; warp 17 is the TMA producer (Q once, K / V through 3-stage rings).
// Softmax: there is no exact running maximum.  Per (row, key tile) the shift is an UPPER BOUND of the scores,
//   bound = scale * max_j raw_ij - slope * (distance of row i to the tile),
// which one 3-input max per two scores yields; it exceeds the true maximum by at most the ALiBi spread inside a tile
// (slope * 63 <= 23 in log2 units), harmless in fp32 / bf16 floating point.  The row's shift only moves (and O / l are only
// rescaled, in TMEM, by the row's own thread) when the bound grows by more than 2^8 -- after the first tiles practically never.
#include "attention_tc.cuh"

namespace {
using namespace attn_tc;

constexpr int QG = 4;                   // query tiles per CTA
constexpr int FK = 64;                  // keys per tile of the forward
constexpr int KST = 3;                  // K / V ring depth
constexpr int SQ_OFF = 0;                         // Q tiles   [QG][128 rows][64]     64 KB
constexpr int SK_OFF = SQ_OFF + QG * 16384;       // K ring    [KST][64 keys][64]     24 KB
constexpr int SV_OFF = SK_OFF + KST * 8192;       // V ring                           24 KB
constexpr int SP_OFF = SV_OFF + KST * 8192;       // P tiles   [QG][128 rows][64 keys] 64 KB
constexpr int BAR_OFF = SP_OFF + QG * 16384;
constexpr int TC_SMEM_BYTES = BAR_OFF + 512 + 1024;
static_assert(SP_OFF % 1024 == 0 && TC_SMEM_BYTES <= 232448, "shared-memory plan");
constexpr int FWD_EW_WARPS = 4 * QG;
constexpr int FWD_THREADS = 32 * (FWD_EW_WARPS + 2);
constexpr float RESCALE_STEP = 8.f;     // log2 units

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

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* q_full = bars + 0;          // [QG]
    uint64_t* s_full = bars + 4;          // [QG]  S_q of the current key tile is in TMEM
    uint64_t* p_full = bars + 8;          // [QG]  P_q is in shared memory (4 warps) -- S_q has been read out, O_q rescaled if needed
    uint64_t* pv_done = bars + 12;        // [QG]  O_q += P_q V has completed: P_q may be overwritten, O_q may be read
    uint64_t* k_full = bars + 16;         // [KST]
    uint64_t* k_empty = bars + 19;
    uint64_t* v_full = bars + 22;
    uint64_t* v_empty = bars + 25;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (QG * QP), b = blockIdx.y;
    const int T = p.T;
    const int n_tiles_all = (T + FK - 1) / FK;
    // key tiles visible to query tile q (0 when the tile lies beyond the sequence end); non-decreasing in q up to the last active one
    auto tiles_of = [&](int q) -> int {
        const int first = q0 + q * QP;
        if (first >= T) return 0;
        return p.causal ? min(n_tiles_all, (min(first + QP, T) - 1) / FK + 1) : n_tiles_all;
    };

    if (warp == FWD_EW_WARPS && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        for (int q = 0; q < QG; ++q) {
            mbar_init(&q_full[q], 1); mbar_init(&s_full[q], 1); mbar_init(&p_full[q], 4); mbar_init(&pv_done[q], 1);
        }
        for (int s_ = 0; s_ < KST; ++s_) {
            mbar_init(&k_full[s_], 1); mbar_init(&k_empty[s_], 1); mbar_init(&v_full[s_], 1); mbar_init(&v_empty[s_], 1);
        }
        fence_mbar_init();
    }
    if (warp == FWD_EW_WARPS) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == FWD_EW_WARPS) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, FK, false, false);        // S:  [128 x 64 keys], K-major A and B
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DH, false, true);      // PV: [128 x 64 dh], V read MN-major
            const uint32_t sq = smem_u32(smem + SQ_OFF), sk = smem_u32(smem + SK_OFF), sv = smem_u32(smem + SV_OFF),
                           sp = smem_u32(smem + SP_OFF);
            int nt[QG];
#pragma unroll
            for (int q = 0; q < QG; ++q) nt[q] = tiles_of(q);
            const int nt_max = max(max(nt[0], nt[1]), max(nt[2], nt[3]));
            auto issue_s = [&](int q, int kt) {
                const uint32_t kb = sk + (kt % KST) * 8192;
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16(tmem_base + q * FK, umma_smem_desc_sw128(sq + q * 16384 + k * 32, 0, 1024),
                              umma_smem_desc_sw128(kb + k * 32, 0, 1024), idesc, k > 0 ? 1u : 0u);
                umma_commit(&s_full[q]);
            };
            SPB_MBAR_WAIT(&k_full[0], 0);
#pragma unroll
            for (int q = 0; q < QG; ++q)
                if (nt[q] > 0) {
                    SPB_MBAR_WAIT(&q_full[q], 0);
                    tc_fence_after();
                    issue_s(q, 0);
                }
            umma_commit(&k_empty[0]);
            for (int kt = 0; kt < nt_max; ++kt) {
                const int st = kt % KST, nxt = kt + 1;
                SPB_MBAR_WAIT(&v_full[st], (kt / KST) & 1);
                if (nxt < nt_max) SPB_MBAR_WAIT(&k_full[nxt % KST], (nxt / KST) & 1);
#pragma unroll
                for (int q = 0; q < QG; ++q) {
                    if (kt >= nt[q]) continue;
                    SPB_MBAR_WAIT(&p_full[q], kt & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < FK / 16; ++ks)
                        umma_bf16(tmem_base + QG * FK + q * DH, umma_smem_desc_sw128(sp + q * 16384 + ks * 32, 0, 1024),
                                  umma_smem_desc_sw128(sv + st * 8192 + ks * 2048, 8192, 1024), idesc_pv, (kt > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(&pv_done[q]);
                    if (nxt < nt[q]) issue_s(q, nxt);
                }
                umma_commit(&v_empty[st]);
                if (nxt < nt_max) umma_commit(&k_empty[nxt % KST]);
            }
        }
    } else if (warp == FWD_EW_WARPS + 1) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int nt_max = 0;
            for (int q = 0; q < QG; ++q) {
                const int n = tiles_of(q);
                nt_max = max(nt_max, n);
                if (n == 0) continue;
                mbar_arrive_expect_tx(&q_full[q], NH * QP * 128);
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    tma_load_3d(smem + SQ_OFF + q * 16384 + h * (QP * 128), &tmQ, &q_full[q], h * DH, q0 + q * QP, b);
                if (q == 0) {             // the first K tile right behind the first Q tile
                    mbar_arrive_expect_tx(&k_full[0], FK * 128);
                    tma_load_3d(smem + SK_OFF, &tmKV, &k_full[0], p.kcol, 0, b);
                }
            }
            for (int kt = 0; kt < nt_max; ++kt) {
                const int st = kt % KST;
                const uint32_t ph = ((kt / KST) & 1) ^ 1;
                if (kt > 0) {
                    SPB_MBAR_WAIT(&k_empty[st], ph);
                    mbar_arrive_expect_tx(&k_full[st], FK * 128);
                    tma_load_3d(smem + SK_OFF + st * 8192, &tmKV, &k_full[st], p.kcol, kt * FK, b);
                }
                SPB_MBAR_WAIT(&v_empty[st], ph);
                mbar_arrive_expect_tx(&v_full[st], FK * 128);
                tma_load_3d(smem + SV_OFF + st * 8192, &tmKV, &v_full[st], p.vcol, kt * FK, b);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps: quartet q = warp / 4, head = warp % 4
        const int q = warp >> 2, h = warp & 3;
        const int n_kt = tiles_of(q);
        if (n_kt > 0) {
            const int i_first = q0 + q * QP;               // first position of the quartet's tile
            const int i = i_first + lane;
            const int r = h * QP + lane;
            const uint32_t lane_addr = (uint32_t)(h * 32) << 16;
            const uint32_t tm_s = tmem_base + lane_addr + q * FK;
            const uint32_t tm_o = tmem_base + lane_addr + QG * FK + q * DH;
            const float slope = __expf(p.logslopes[h]) * LOG2E;
            const float scale2 = p.scale * LOG2E;
            DropParams drop;
            drop.seedmix = drop_seedmix(p.seed, p.rng_offset);
            drop.thr32 = p.thr32;
            drop.quarter_t = (uint32_t)((T + 3) >> 2);
            drop.keep_scale = p.keep_scale;
            const uint32_t drop_row = drop_row_base(drop, (uint32_t)((b * NH + h) * T + i));
            const bool drop_on = p.thr32 != 0;
            const uint32_t sP_row = smem_u32(smem + SP_OFF + q * 16384 + r * 128);
            const uint32_t swz = (uint32_t)(r & 7);

            float m_run = -INFINITY, l_run = 0.f, d_run = 0.f;   // shift, sum of numerators, sum of numerators * |i-j|

            for (int kt = 0; kt < n_kt; ++kt) {
                const int j0t = kt * FK;
                // validity bits of the tile's two 32-key chunks (key padding, sequence tail, causal limit of THIS row)
                uint32_t vbits[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int j0 = j0t + c * 32;
                    uint32_t bits = 0u;
                    if (j0 < T) {
                        bits = p.mask_bits != nullptr ? __ldg(p.mask_bits + (size_t)b * p.words_per_row + (j0 >> 5)) : 0xffffffffu;
                        if (j0 + 32 > T) bits &= (1u << (T - j0)) - 1u;
                        if (p.causal) {
                            const int lim = i - j0;     // keys j0 .. j0+lim allowed
                            bits &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
                        }
                    }
                    vbits[c] = bits;
                }
                const bool plain = __all_sync(0xffffffffu, (vbits[0] & vbits[1]) == 0xffffffffu);
                SPB_MBAR_WAIT(&s_full[q], kt & 1);
                tc_fence_after();
                // ---- pass 1: upper bound of the row's scores in this tile
                float rawmax = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tm_s + c * 32, v);
                    tmem_ld_wait();
                    if (plain) {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 2) rawmax = fmax3(rawmax, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
                    } else {
                        const uint32_t bits = vbits[c];
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) rawmax = fmaxf(rawmax, ((bits >> jj) & 1u) ? __uint_as_float(v[jj]) : -INFINITY);
                    }
                }
                const int gap = i < j0t ? j0t - i : (i > j0t + FK - 1 ? i - (j0t + FK - 1) : 0);
                const float bound = fmaf(rawmax, scale2, -slope * (float)gap);       // -inf when the row sees no key of the tile
                if (kt == 0) {
                    m_run = bound;
                } else {
                    const bool grow = bound > m_run + RESCALE_STEP;
                    if (__any_sync(0xffffffffu, grow)) {
                        // rare: move the shift of the rows that need it and rescale their O / l / d (O lives in TMEM)
                        const float corr = grow ? exp2f(m_run - bound) : 1.f;        // m_run == -inf -> 0
                        if (grow) m_run = bound;
                        l_run *= corr;
                        d_run *= corr;
                        SPB_MBAR_WAIT(&pv_done[q], (kt - 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t v[32];
                            tmem_ld_32x32b_x32(tm_o + half * 32, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int d = 0; d < 32; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) * corr);
                            tmem_st_32x32b_x32(tm_o + half * 32, v);
                        }
                        tmem_st_wait();
                    }
                }
                const float m_use = m_run == -INFINITY ? 0.f : m_run;
                // ---- pass 2: numerators, dropout, bf16 pack into the swizzled A-operand tile
                float rsum = 0.f, dsum = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tm_s + c * 32, v);
                    tmem_ld_wait();
                    const int j0 = j0t + c * 32;
                    const float dbase = (float)(i - j0);
                    // the 32 keys lie on one side of all 32 rows of the warp: |i-j| is linear in j
                    const bool left = j0 + 31 <= i_first, right = j0 >= i_first + QP - 1;
                    if (plain && (left || right)) {
                        const float ss = left ? slope : -slope;                      // x = s*scale2 + ss*jj + c0
                        const float c0 = fmaf(-ss, dbase, -m_use);
                        float cs = 0.f, ws = 0.f;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const float e = exp2f(fmaf(__uint_as_float(v[jj]), scale2, fmaf(ss, (float)jj, c0)));
                            cs += e;
                            ws = fmaf(e, (float)jj, ws);
                            v[jj] = __float_as_uint(e);
                        }
                        rsum += cs;
                        const float dd = fmaf(dbase, cs, -ws);                       // sum e * (dbase - jj)
                        dsum += left ? dd : -dd;
                    } else {
                        const uint32_t bits = vbits[c];
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const float dist = fabsf(dbase - (float)jj);
                            float e = exp2f(fmaf(-slope, dist, fmaf(__uint_as_float(v[jj]), scale2, -m_use)));
                            e = ((bits >> jj) & 1u) ? e : 0.f;
                            rsum += e;
                            dsum = fmaf(e, dist, dsum);
                            v[jj] = __float_as_uint(e);
                        }
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
                    if (c == 0 && kt > 0) SPB_MBAR_WAIT(&pv_done[q], (kt - 1) & 1);      // PV of the previous tile has read P_q
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(v[qd * 8 + 0]), __uint_as_float(v[qd * 8 + 1]));
                        u.y = pack_bf16x2(__uint_as_float(v[qd * 8 + 2]), __uint_as_float(v[qd * 8 + 3]));
                        u.z = pack_bf16x2(__uint_as_float(v[qd * 8 + 4]), __uint_as_float(v[qd * 8 + 5]));
                        u.w = pack_bf16x2(__uint_as_float(v[qd * 8 + 6]), __uint_as_float(v[qd * 8 + 7]));
                        sts_u4(sP_row + ((((uint32_t)(c * 4 + qd)) ^ swz) << 4), u);
                    }
                }
                l_run += rsum;
                d_run += dsum;
                tc_fence_before();                     // S_q is in registers / O_q rescaled: the MMA thread may go on
                fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[q]);
            }
            // ---- epilogue: O_q / l
            SPB_MBAR_WAIT(&pv_done[q], (n_kt - 1) & 1);
            tc_fence_after();
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tm_o + half * 32, v);
                tmem_ld_wait();
                if (i < T) {
                    __nv_bfloat16* dst = p.out + ((size_t)b * T + i) * p.ld_out + h * DH + half * 32;
#pragma unroll
                    for (int d = 0; d < 32; d += 8) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(v[d]) * inv, __uint_as_float(v[d + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(v[d + 2]) * inv, __uint_as_float(v[d + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(v[d + 4]) * inv, __uint_as_float(v[d + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(v[d + 6]) * inv, __uint_as_float(v[d + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + d) = u;
                    }
                }
            }
            if (i < T) {
                if (p.lse != nullptr) p.lse[((size_t)b * NH + h) * T + i] = l_run > 0.f ? (m_run + log2f(l_run)) : INFINITY;
                if (p.edist != nullptr) p.edist[((size_t)b * NH + h) * T + i] = d_run * inv;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FWD_EW_WARPS) tmem_dealloc<512>(tmem_base);
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
    rc = spb_make_tmap_bf16_3d(&tmKV, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)T * ld * 2, DH, FK);
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
    attn_fwd_tc_kernel<<<dim3(ceil_div(T, QG * QP), B), FWD_THREADS, TC_SMEM_BYTES, stream>>>(tmQ, tmKV, p);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
