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
constexpr int FK = 64;                  // keys per K / V ring stage
constexpr int HK = 32;                  // keys per pipeline step (one S product, one P half-tile, one PV product)
constexpr int KST = 3;                  // K / V ring depth
constexpr int SQ_OFF = 0;                         // Q tiles   [QG][128 rows][64]     64 KB
constexpr int SK_OFF = SQ_OFF + QG * 16384;       // K ring    [KST][64 keys][64]     24 KB
constexpr int SV_OFF = SK_OFF + KST * 8192;       // V ring                           24 KB
constexpr int SP_OFF = SV_OFF + KST * 8192;       // P tiles   [QG][128 rows][64 keys] 64 KB
constexpr int BAR_OFF = SP_OFF + QG * 16384;
constexpr int TC_SMEM_BYTES = BAR_OFF + 512 + 1024;
static_assert(SP_OFF % 1024 == 0 && TC_SMEM_BYTES <= 232448, "shared-memory plan");
constexpr int FWD_EW_WARPS = 4 * QG;
constexpr int FWD_THREADS = 32 * (FWD_EW_WARPS + 3);     // + S issuer, PV issuer, TMA producer
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
// packed fp32 pairs (sm_100 FFMA2 / FADD2): two lanes of arithmetic per issue slot
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// non-blocking probe of an mbarrier phase (try_wait may suspend the thread for a while: not what a polling loop wants)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// eight numerators of a row -> keep / drop (the 1/(1-p) factor is folded into the final normalisation) -> bf16 -> one 16-byte
// piece of the swizzled P tile
__device__ __forceinline__ void emit8(float (&e)[8], bool drop_on, uint32_t drop_pre, uint32_t thr32, uint32_t dst) {
    if (drop_on) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint32_t qh = drop_quad(drop_pre + (uint32_t)t * DROP_K);
#pragma unroll
            for (int k = 0; k < 4; ++k) e[4 * t + k] = drop_keep(qh, k, thr32) ? e[4 * t + k] : 0.f;
        }
    }
    sts_u4(dst, make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7])));
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* q_full = bars + 0;          // [QG]
    uint64_t* s_full = bars + 4;          // [QG]  S_q of the current key tile is in TMEM
    uint64_t* s_free = bars + 8;          // [QG]  ... and has been read into registers for good (4 warps): S_q of the next tile may land
    uint64_t* p_full = bars + 12;         // [QG]  P_q is in shared memory (4 warps), O_q rescaled if needed
    uint64_t* pv_done = bars + 16;        // [QG]  O_q += P_q V has completed: P_q may be overwritten, O_q may be read
    uint64_t* k_full = bars + 20;         // [KST]
    uint64_t* k_empty = bars + 20 + KST;
    uint64_t* v_full = bars + 20 + 2 * KST;
    uint64_t* v_empty = bars + 20 + 3 * KST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20 + 4 * KST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (QG * QP), b = blockIdx.y;
    const int T = p.T;
    // The CTA walks the key tiles from its own diagonal outwards: tile kd (the one holding its last query row, or the last
    // visible one) down to 0, then kd+1 upwards.  ALiBi makes the near tiles carry the large scores, so the softmax shift
    // settles on the first tile a quartet sees and the O rescale (which has to wait for the tensor pipe) stays rare.
    // Under the causal mask quartet q only sees tiles <= its own diagonal: it joins the walk n0(q) steps late.
    const int n_tiles_all = (T + FK - 1) / FK;
    const int kd = min(n_tiles_all, (min(q0 + QG * QP, T) - 1) / FK + 1) - 1;
    const int n_walk = p.causal ? kd + 1 : n_tiles_all;
    auto tile_at = [&](int n) -> int { return n <= kd ? kd - n : n; };
    auto first_step = [&](int q) -> int {           // n_walk when the quartet's rows lie beyond the sequence end
        const int first = q0 + q * QP;
        if (first >= T) return n_walk;
        return p.causal ? kd - (min(n_tiles_all, (min(first + QP, T) - 1) / FK + 1) - 1) : 0;
    };

    if (warp == FWD_EW_WARPS && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        for (int q = 0; q < QG; ++q) {
            mbar_init(&q_full[q], 1); mbar_init(&s_full[q], 1); mbar_init(&s_free[q], 4); mbar_init(&p_full[q], 4); mbar_init(&pv_done[q], 1);
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
        // ------------------------------------------------------------------ S issuer: S_q(kt+1) = Q_q K(kt+1)^T as soon as quartet q
        // has read S_q(kt) out of TMEM -- the tail of its tile (last exp2s, P stores) covers the product's latency.  A second
        // thread (next branch) issues the PV products, so a quartet waiting for one kind of product never holds up the other.
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, FK, false, false);      // S:  [128 x 64 keys], K-major A and B
            const uint32_t sq = smem_u32(smem + SQ_OFF), sk = smem_u32(smem + SK_OFF);
            int n0[QG];
#pragma unroll
            for (int q = 0; q < QG; ++q) n0[q] = first_step(q);
            for (int n = 0; n < n_walk; ++n) {
                SPB_MBAR_WAIT(&k_full[n % KST], (n / KST) & 1);
                const uint32_t kb = sk + (n % KST) * 8192;
#pragma unroll
                for (int q = 0; q < QG; ++q) {
                    if (n < n0[q]) continue;
                    if (n == n0[q]) SPB_MBAR_WAIT(&q_full[q], 0);
                    else SPB_MBAR_WAIT(&s_free[q], (n - n0[q] - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k)
                        umma_bf16(tmem_base + q * FK, umma_smem_desc_sw128(sq + q * 16384 + k * 32, 0, 1024),
                                  umma_smem_desc_sw128(kb + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
                    umma_commit(&s_full[q]);
                }
                umma_commit(&k_empty[n % KST]);
            }
        }
    } else if (warp == FWD_EW_WARPS + 2) {
        // ------------------------------------------------------------------ PV issuer: O_q += P_q V(kt) as each P_q arrives
        if (lane == 0) {
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DH, false, true);      // PV: [128 x 64 dh], V read MN-major
            const uint32_t sv = smem_u32(smem + SV_OFF), sp = smem_u32(smem + SP_OFF);
            int n0[QG];
#pragma unroll
            for (int q = 0; q < QG; ++q) n0[q] = first_step(q);
            for (int n = 0; n < n_walk; ++n) {
                const int st = n % KST;
                SPB_MBAR_WAIT(&v_full[st], (n / KST) & 1);
#pragma unroll
                for (int q = 0; q < QG; ++q) {
                    if (n < n0[q]) continue;
                    SPB_MBAR_WAIT(&p_full[q], (n - n0[q]) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < FK / 16; ++ks)
                        umma_bf16(tmem_base + QG * FK + q * DH, umma_smem_desc_sw128(sp + q * 16384 + ks * 32, 0, 1024),
                                  umma_smem_desc_sw128(sv + st * 8192 + ks * 2048, 8192, 1024), idesc_pv, (n > n0[q] || ks > 0) ? 1u : 0u);
                    umma_commit(&pv_done[q]);
                }
                umma_commit(&v_empty[st]);
            }
        }
    } else if (warp == FWD_EW_WARPS + 1) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&k_full[0], FK * 128);      // the first K tile ahead of everything
            tma_load_3d(smem + SK_OFF, &tmKV, &k_full[0], p.kcol, tile_at(0) * FK, b);
            for (int q = QG - 1; q >= 0; --q) {               // the quartets nearest the first tile start first
                if (first_step(q) >= n_walk) continue;
                mbar_arrive_expect_tx(&q_full[q], NH * QP * 128);
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    tma_load_3d(smem + SQ_OFF + q * 16384 + h * (QP * 128), &tmQ, &q_full[q], h * DH, q0 + q * QP, b);
            }
            for (int n = 0; n < n_walk; ++n) {
                const int st = n % KST, kt = tile_at(n);
                const uint32_t ph = ((n / KST) & 1) ^ 1;
                if (n > 0) {
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
        const int n_first = first_step(q);
        if (n_first < n_walk) {
            const int i_first = q0 + q * QP;               // first position of the quartet's tile
            const int i = i_first + lane;
            const int r = h * QP + lane;
            const uint32_t lane_addr = (uint32_t)(h * 32) << 16;
            const uint32_t tm_s = tmem_base + lane_addr + q * FK;
            const uint32_t tm_o = tmem_base + lane_addr + QG * FK + q * DH;
            const float slope = __expf(p.logslopes[h]) * LOG2E;
            const float scale2 = p.scale * LOG2E;
            const uint64_t scale2_2 = pk2(scale2, scale2);
            DropParams drop;
            drop.seedmix = drop_seedmix(p.seed, p.rng_offset);
            drop.thr32 = p.thr32;
            drop.quarter_t = (uint32_t)((T + 3) >> 2);
            drop.keep_scale = p.keep_scale;
            const uint32_t drop_row = drop_row_base(drop, (uint32_t)((b * NH + h) * T + i));
            const bool drop_on = p.thr32 != 0;
            const uint32_t sP_row = smem_u32(smem + SP_OFF + q * 16384 + r * 128);
            const uint32_t swz = (uint32_t)(r & 7);
            const uint32_t* mask_row = p.mask_bits != nullptr ? p.mask_bits + (size_t)b * p.words_per_row : nullptr;

            float m_run = -INFINITY, l_run = 0.f, d_run = 0.f;   // shift, sum of numerators, sum of numerators * |i-j|

            for (int n = n_first; n < n_walk; ++n) {
                const int kt = n - n_first;         // the quartet's own step count (barrier phases)
                const int j0t = tile_at(n) * FK;
                // validity bits of the tile's two 32-key halves (key padding, sequence tail, causal limit of THIS row)
                uint32_t vbits[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int j0 = j0t + c * 32;
                    uint32_t bits = 0u;
                    if (j0 < T) {
                        bits = mask_row != nullptr ? __ldg(mask_row + (j0 >> 5)) : 0xffffffffu;
                        if (j0 + 32 > T) bits &= (1u << (T - j0)) - 1u;
                        if (p.causal) {
                            const int lim = i - j0;     // keys j0 .. j0+lim allowed
                            bits &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
                        }
                    }
                    vbits[c] = bits;
                }
                SPB_MBAR_WAIT(&s_full[q], kt & 1);
                tc_fence_after();
                // ---- pass 1: upper bound of the row's scores in this tile (sixteen keys at a time; rolled loops keep the code small)
                float rawmax = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tm_s + c * 32, v);
                    tmem_ld_wait();
                    const uint32_t bits = vbits[c];
                    if (__all_sync(0xffffffffu, bits == 0xffffffffu)) {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 2) rawmax = fmax3(rawmax, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
                    } else {
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
                            uint32_t o[32];
                            tmem_ld_32x32b_x32(tm_o + half * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int d = 0; d < 32; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * corr);
                            tmem_st_32x32b_x32(tm_o + half * 32, o);
                        }
                        tmem_st_wait();
                    }
                }
                const float m_use = m_run == -INFINITY ? 0.f : m_run;
                // ---- pass 2: numerators, dropout, bf16 pack into the swizzled A-operand tile, eight keys (one 16-byte piece) at a time
                float rsum = 0.f, dsum = 0.f;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tm_s + c * 32, v);
                    tmem_ld_wait();
                    if (c == 1) {            // S_q is in registers for good: the next tile's product may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_free[q]);
                    }
                    if (c == 0 && kt > 0) SPB_MBAR_WAIT(&pv_done[q], (kt - 1) & 1);      // PV of the previous tile has read P_q
                    const int j0 = j0t + c * 32;
                    const float dbase = (float)(i - j0);
                    const uint32_t drop_base = drop_row + (uint32_t)(j0 >> 2) * DROP_K;
                    const uint32_t piece = (uint32_t)(c * 4);
                    // the 32 keys lie on one side of all 32 rows of the warp: |i-j| is linear in j
                    const bool left = j0 + 31 <= i_first, right = j0 >= i_first + QP - 1;
                    const uint32_t bits = vbits[c];
                    if ((left || right) && __all_sync(0xffffffffu, bits == 0xffffffffu)) {
                        // x_jj = s_jj * scale2 + (ss * jj + c0): the bias term runs as a packed pair stepped by 2*ss
                        const float ss = left ? slope : -slope;
                        const float c0 = fmaf(-ss, dbase, -m_use);
                        uint64_t cb = pk2(c0, c0 + ss);
                        const uint64_t step2 = pk2(2.f * ss, 2.f * ss);
                        uint64_t es = pk2(0.f, 0.f), wsum = pk2(0.f, 0.f);       // sum e, sum e * (ss*jj + c0)
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float e[8];
#pragma unroll
                            for (int u = 0; u < 8; u += 2) {
                                const int jj = g * 8 + u;
                                const uint64_t x2 = ffma2(pk2(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1])), scale2_2, cb);
                                float x0, x1;
                                upk2(x2, x0, x1);
                                e[u] = ex2(x0);
                                e[u + 1] = ex2(x1);
                                const uint64_t e2 = pk2(e[u], e[u + 1]);
                                es = fadd2(es, e2);
                                wsum = ffma2(e2, cb, wsum);
                                cb = fadd2(cb, step2);
                            }
                            emit8(e, drop_on, drop_base + (uint32_t)(2 * g) * DROP_K, drop.thr32, sP_row + (((piece + g) ^ swz) << 4));
                        }
                        float a0, a1, w0, w1;
                        upk2(es, a0, a1);
                        upk2(wsum, w0, w1);
                        const float cs = a0 + a1;
                        rsum += cs;
                        // sum e*jj = (W - c0*cs) / ss;  sum e*|i-j| = +-(dbase*cs - sum e*jj)
                        const float ejj = __fdividef(fmaf(-c0, cs, w0 + w1), ss);
                        const float dd = fmaf(dbase, cs, -ejj);
                        dsum += left ? dd : -dd;
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float e[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int jj = g * 8 + u;
                                const float dist = fabsf(dbase - (float)jj);
                                float ev = ex2(fmaf(-slope, dist, fmaf(__uint_as_float(v[jj]), scale2, -m_use)));
                                ev = ((bits >> jj) & 1u) ? ev : 0.f;
                                rsum += ev;
                                dsum = fmaf(ev, dist, dsum);
                                e[u] = ev;
                            }
                            emit8(e, drop_on, drop_base + (uint32_t)(2 * g) * DROP_K, drop.thr32, sP_row + (((piece + g) ^ swz) << 4));
                        }
                    }
                }
                l_run += rsum;
                d_run += dsum;
                tc_fence_before();                     // O_q rescaled (if it was): ordered before the MMA thread's next accumulate
                fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[q]);
            }
            // ---- epilogue: O_q / l (dropout's 1/(1-p) rides along)
            SPB_MBAR_WAIT(&pv_done[q], (n_walk - n_first - 1) & 1);
            tc_fence_after();
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            const float inv_o = drop_on ? inv * p.keep_scale : inv;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tm_o + half * 32, o);
                tmem_ld_wait();
                if (i < T) {
                    __nv_bfloat16* dst = p.out + ((size_t)b * T + i) * p.ld_out + h * DH + half * 32;
#pragma unroll
                    for (int d = 0; d < 32; d += 8) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[d]) * inv_o, __uint_as_float(o[d + 1]) * inv_o);
                        u.y = pack_bf16x2(__uint_as_float(o[d + 2]) * inv_o, __uint_as_float(o[d + 3]) * inv_o);
                        u.z = pack_bf16x2(__uint_as_float(o[d + 4]) * inv_o, __uint_as_float(o[d + 5]) * inv_o);
                        u.w = pack_bf16x2(__uint_as_float(o[d + 6]) * inv_o, __uint_as_float(o[d + 7]) * inv_o);
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
