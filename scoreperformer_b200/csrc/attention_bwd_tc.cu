// tcgen05 / TMEM / TMA backward of the fused multi-query attention: dQ, dK, dV and d(logslope) in ONE kernel.
//
// Reference semantics: the autograd of modules/transformer/attend.py:58-126 + attention.py:139-197 (see attention.cu for the
// forward statement).  With P = softmax(S), S_ij = scale q_i.k_j - slope_h |i-j| and Pd = dropout(P):
//   dV = Pd^T dO          dP = dO V^T          dS = P o (keep o dP - delta),  delta_i = sum_d dO_id O_id
//   dQ = scale dS K       dK = scale dS^T Q    d logslope_h = -slope_h sum_ij dS_ij |i-j|
//
// One CTA owns a tile of 128 KEYS of one sequence and walks the query tiles that can see it.  MQA lets the four query heads
// share K and V, so a query tile stacks 32 positions x 4 heads into the 128 rows of one UMMA tile (row = head*32 + position).
// Per query tile five tcgen05.mma chains run, all operands in 128B-swizzled shared memory, all accumulators in TMEM:
//   S  [128 q x 128 k] = Q K^T          dP [128 q x 128 k] = dO V^T                 (K-major A and B)
//   dV [128 k x 64]   += Pd^T dO        dK [128 k x 64]   += dS^T Q                 (MN-major A and B: no transposes staged)
//   dQ [128 q x 64]    = dS K                                                        (K-major A, MN-major B)
// dK / dV stay in TMEM for the whole CTA (summed over heads and queries for free); dQ leaves every iteration through a TMA
// tensor reduce-add into an fp32 accumulator (the only cross-CTA reduction of the design).
// Warp roles (576 threads, one CTA per SM): warp 0 = MMA issuer, warp 17 = TMA producer; warps 1-16 = the element-wise
// part: one S/dP row per thread (TMEM lane == row), four warps per lane quarter each taking a 32-key chunk, so there are no
// shuffles and no row reductions anywhere (lse, delta and E[|i-j|] come from the forward).  The kernel is bound by the
// element-wise instruction stream (exp2, dropout, dS), not by the tensor pipe: 16 warps keep four per scheduler in flight.
#include "attention_tc.cuh"
#include <stdlib.h>

namespace {
using namespace attn_tc;

constexpr int QSTAGES = 3;              // Q / dO ring: tiles are requested two query tiles before their first MMA
constexpr int SK_OFF = 0;               // K tile  [128 keys][64]            16 KB
constexpr int SV_OFF = 16384;           // V tile                             16 KB
constexpr int SQ_OFF = 32768;           // Q tiles [QSTAGES][128 rows][64]    48 KB
constexpr int SDO_OFF = SQ_OFF + QSTAGES * 16384;      // dO tiles            48 KB
constexpr int SP_OFF = SDO_OFF + QSTAGES * 16384;      // dropped P, [2 key halves][128 rows][64 keys]   32 KB
constexpr int SDS_OFF = SP_OFF + 32768;  // dS, same layout                   32 KB
constexpr int SDQ_OFF = SDS_OFF + 32768; // dQ staging: 8 warps x one 4 KB box
constexpr int BAR_OFF = SDQ_OFF + 32768;
constexpr int BWD_SMEM_BYTES = BAR_OFF + 512 + 1024;
static_assert(BWD_SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may have");
constexpr int EW_WARPS = 16;            // element-wise warps: 4 TMEM lane quarters x 4 key chunks of 32
constexpr int BWD_THREADS = 32 * (2 + EW_WARPS);       // + MMA issuer (warp 0) + TMA producer (last warp)

constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

struct BwdParams {
    const uint32_t* mask_bits;   // [B, words_per_row] key validity bits, or null
    int words_per_row;
    const float* logslopes;      // [H]
    const float* lse;            // [B, H, T] base-2 log-sum-exp of the forward
    const float* delta;          // [B, H, T] rowsum(dO * O)
    const float* edist;          // [B, H, T] E_i[|i-j|] under the (undropped) attention distribution of the forward
    __nv_bfloat16* dqkv;         // [B*T, ld_dqkv]: the k and v columns are written here
    int ld_dqkv;
    float* dlogslopes;           // [H], accumulated into (may be null)
    int B, T;
    float scale;
    int causal;
    uint64_t seed;
    const uint64_t* rng_offset;
    uint32_t thr32;
    float keep_scale;
    int kcol, vcol;
};

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ, BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = bars + 0;
    uint64_t* q_full = bars + 1;       // [QSTAGES]
    uint64_t* q_empty = bars + 5;      // [QSTAGES]
    uint64_t* sdp_full = bars + 9;     // S and dP of this query tile are in TMEM
    uint64_t* sdp_free = bars + 10;    // ... and have been read out (16 warps)
    uint64_t* pds_full = bars + 11;    // Pd and dS tiles are in shared memory (16 warps)
    uint64_t* pds_free = bars + 12;    // ... and the three MMAs that read them have completed
    uint64_t* dq_full = bars + 13;
    uint64_t* dq_free = bars + 14;     // dQ accumulator read out (8 warps)
    uint64_t* dkv_full = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int kt = blockIdx.x / p.B, b = blockIdx.x - kt * p.B;     // all CTAs of key tile 0 first: the longest (causal) go first
    const int k0 = kt * TKEY;
    const int n_qt = (T + QP - 1) / QP;
    const int qt_begin = p.causal ? k0 / QP : 0;
    const int iters = n_qt - qt_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmDO);
        tma_prefetch_desc(&tmDQ);
        mbar_init(kv_full, 1);
        for (int s_ = 0; s_ < QSTAGES; ++s_) { mbar_init(&q_full[s_], 1); mbar_init(&q_empty[s_], 1); }
        mbar_init(sdp_full, 1); mbar_init(sdp_free, EW_WARPS);
        mbar_init(pds_full, EW_WARPS); mbar_init(pds_free, 1);
        mbar_init(dq_full, 1); mbar_init(dq_free, EW_WARPS / 2);
        mbar_init(dkv_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ control thread: TMA loads and MMA issue
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, TKEY, false, false);     // S, dP
            constexpr uint32_t idesc_kv = umma_idesc_bf16(128, DH, true, true);        // dV, dK
            constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, false, true);        // dQ
            const uint32_t sk = smem_u32(smem + SK_OFF), sv = smem_u32(smem + SV_OFF), sp = smem_u32(smem + SP_OFF),
                           sds = smem_u32(smem + SDS_OFF);
            SPB_MBAR_WAIT(kv_full, 0);
            // S / dP of tile `it` are issued ONE TILE AHEAD: the element-wise warps copy a tile's S / dP out of TMEM into registers
            // first thing (sdp_free), so the next tile's two products run on the tensor pipe underneath this tile's exp2 / dS
            // arithmetic, and so do this tile's dV / dK / dQ products underneath the next tile's.
            auto issue_s_dp = [&](int it) {
                const int st = it % QSTAGES;
                const uint32_t sq = smem_u32(smem + SQ_OFF + st * 16384), sdo = smem_u32(smem + SDO_OFF + st * 16384);
                SPB_MBAR_WAIT(&q_full[st], (it / QSTAGES) & 1);
                SPB_MBAR_WAIT(sdp_free, (it & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16(tmem_base + TM_S, umma_smem_desc_sw128(sq + k * 32, 0, 1024), umma_smem_desc_sw128(sk + k * 32, 0, 1024),
                              idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16(tmem_base + TM_DP, umma_smem_desc_sw128(sdo + k * 32, 0, 1024), umma_smem_desc_sw128(sv + k * 32, 0, 1024),
                              idesc_s, k > 0 ? 1u : 0u);
                umma_commit(sdp_full);
            };
            issue_s_dp(0);
            for (int it = 0; it < iters; ++it) {
                const int st = it % QSTAGES;
                const uint32_t sq = smem_u32(smem + SQ_OFF + st * 16384), sdo = smem_u32(smem + SDO_OFF + st * 16384);
                if (it + 1 < iters) issue_s_dp(it + 1);
                SPB_MBAR_WAIT(pds_full, it & 1);
                SPB_MBAR_WAIT(dq_free, (it & 1) ^ 1);
                tc_fence_after();
                // dV += Pd^T dO, dK += dS^T Q: the 128 query rows are the contraction; A = [q][keys] read MN-major (two 64-key
                // atoms 16 KB apart), B = [q][d] read MN-major
#pragma unroll
                for (int ks = 0; ks < 128 / 16; ++ks)
                    umma_bf16(tmem_base + TM_DV, umma_smem_desc_sw128(sp + ks * 2048, 16384, 1024),
                              umma_smem_desc_sw128(sdo + ks * 2048, 16384, 1024), idesc_kv, (it > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 128 / 16; ++ks)
                    umma_bf16(tmem_base + TM_DK, umma_smem_desc_sw128(sds + ks * 2048, 16384, 1024),
                              umma_smem_desc_sw128(sq + ks * 2048, 16384, 1024), idesc_kv, (it > 0 || ks > 0) ? 1u : 0u);
                // dQ = dS K: the 128 keys are the contraction; A = dS K-major (two 64-key blocks), B = K tile read MN-major
#pragma unroll
                for (int ks = 0; ks < TKEY / 16; ++ks)
                    umma_bf16(tmem_base + TM_DQ, umma_smem_desc_sw128(sds + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                              umma_smem_desc_sw128(sk + ks * 2048, 16384, 1024), idesc_q, ks > 0 ? 1u : 0u);
                umma_commit(pds_free);
                umma_commit(&q_empty[st]);
                umma_commit(dq_full);
            }
            umma_commit(dkv_full);
        }
    } else if (warp == EW_WARPS + 1) {
        // ------------------------------------------------------------------ TMA producer: K / V once, then the Q / dO ring
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TKEY * 128);
            tma_load_3d(smem + SK_OFF, &tmKV, kv_full, p.kcol, k0, b);
            tma_load_3d(smem + SV_OFF, &tmKV, kv_full, p.vcol, k0, b);
            for (int it = 0; it < iters; ++it) {
                const int st = it % QSTAGES;
                const int q0 = (qt_begin + it) * QP;
                SPB_MBAR_WAIT(&q_empty[st], ((it / QSTAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&q_full[st], 2 * NH * QP * 128);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    tma_load_3d(smem + SQ_OFF + st * 16384 + h * (QP * 128), &tmQ, &q_full[st], h * DH, q0, b);
                    tma_load_3d(smem + SDO_OFF + st * 16384 + h * (QP * 128), &tmDO, &q_full[st], h * DH, q0, b);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ element-wise warps: thread == (head, position) row,
        // warp == (TMEM lane quarter = head, 32-key chunk).  Chunks 0 and 1 also drain the dQ accumulator (32 columns each).
        const int ew = warp - 1;                      // 0..15
        const int h = warp & 3;                       // TMEM lane quarter == head
        const int chunk = ew >> 2;                    // keys [32*chunk, 32*chunk + 32) of the tile
        const int r = h * QP + lane;                  // row of the stacked tile
        const uint32_t lane_addr = (uint32_t)(h * 32) << 16;
        const float slope_nat = __expf(p.logslopes[h]);
        const float slope = slope_nat * LOG2E;
        const float scale2 = p.scale * LOG2E;
        DropParams drop;
        drop.seedmix = drop_seedmix(p.seed, p.rng_offset);
        drop.thr32 = p.thr32;
        drop.quarter_t = (uint32_t)((T + 3) >> 2);
        drop.keep_scale = p.keep_scale;
        const bool drop_on = p.thr32 != 0;
        const int j0 = k0 + chunk * 32;
        // validity of this warp's 32 keys (padding mask and sequence tail): loop invariant
        uint32_t kbits = 0u;
        if (j0 < T) {
            kbits = p.mask_bits != nullptr ? p.mask_bits[(size_t)b * p.words_per_row + (j0 >> 5)] : 0xffffffffu;
            if (j0 + 32 > T) kbits &= (1u << (T - j0)) - 1u;
        }
        uint8_t* sP = smem + SP_OFF + (chunk >> 1) * 16384 + r * 128;
        uint8_t* sDS = smem + SDS_OFF + (chunk >> 1) * 16384 + r * 128;
        const uint32_t swz = (uint32_t)(r & 7);
        const bool drains = chunk < 2;
        uint8_t* stage = smem + SDQ_OFF + (drains ? (chunk * 4 + h) * 4096 : 0);
        float slope_acc = 0.f;                        // sum_ij dS_ij (|i-j| - E_i) over this thread's rows / columns

        // dQ of query tile `it_done`: TMEM -> fp32 box (here) -> TMA reduce-add (issue_dq, after the tile's one proxy fence)
        auto stage_dq = [&](int it_done) {
            SPB_MBAR_WAIT(dq_full, it_done & 1);      // also implies pds_free of that tile: both are committed after the same MMAs
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + TM_DQ + lane_addr + chunk * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(dq_free);
                bulk_wait_group_read<0>();            // the previous reduce has read the box
            }
            __syncwarp();
            const uint32_t my = smem_u32(stage) + (uint32_t)(lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 x;
                x.x = __uint_as_float(v[4 * j]) * p.scale;
                x.y = __uint_as_float(v[4 * j + 1]) * p.scale;
                x.z = __uint_as_float(v[4 * j + 2]) * p.scale;
                x.w = __uint_as_float(v[4 * j + 3]) * p.scale;
                sts_f4(my + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4), x);
            }
        };
        auto issue_dq = [&](int it_done) {            // lane 0, after fence.proxy.async + __syncwarp
            tma_reduce_add_3d(&tmDQ, stage, h * DH + chunk * 32, (qt_begin + it_done) * QP, b);
            bulk_commit_group();
        };

        // per-row statistics of the forward, fetched one tile ahead so their latency hides under the arithmetic
        float nlse_n, delta_n, e_n;
        auto fetch_row = [&](int it) {
            const int i = (qt_begin + it) * QP + lane;
            const bool row_ok = it < iters && i < T;
            const size_t rl = ((size_t)b * NH + h) * T + (row_ok ? i : 0);
            nlse_n = row_ok ? -__ldg(p.lse + rl) : -INFINITY;
            delta_n = row_ok ? __ldg(p.delta + rl) : 0.f;
            e_n = row_ok ? __ldg(p.edist + rl) : 0.f;
        };
        fetch_row(0);
        for (int it = 0; it < iters; ++it) {
            const int q0 = (qt_begin + it) * QP;
            const int i = q0 + lane;
            const float nlse = nlse_n, delta_i = delta_n, e_i = e_n;
            const uint32_t drop_base = drop_row_base(drop, (uint32_t)(((size_t)b * NH + h) * T + i)) + (uint32_t)(j0 >> 2) * DROP_K;
            SPB_MBAR_WAIT(sdp_full, it & 1);
            tc_fence_after();
            uint32_t vs[32], vp[32];
            tmem_ld_32x32b_x32(tmem_base + TM_S + lane_addr + chunk * 32, vs);
            tmem_ld_32x32b_x32(tmem_base + TM_DP + lane_addr + chunk * 32, vp);
            fetch_row(it + 1);
            tmem_ld_wait();
            tc_fence_before();                        // S and dP are in registers: the next query tile's MMAs may overwrite them
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
            uint32_t bits = kbits;
            if (p.causal) {
                const int lim = i - j0;               // keys j0 .. j0 + lim are visible to query i
                bits &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
            }
            const bool plain = __all_sync(0xffffffffu, bits == 0xffffffffu);
            const float dbase = (float)(i - j0);
            uint32_t out_p[16], out_ds[16];           // the tile's Pd / dS pieces stay in registers until the buffers are free
#pragma unroll
            for (int g = 0; g < 8; ++g) {             // 4 keys share one dropout hash
                uint32_t qh = 0;
                if (drop_on) qh = drop_quad(drop_base + (uint32_t)g * DROP_K);
                float pv[4], dsv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int jj = g * 4 + e;
                    const float dist = fabsf(dbase - (float)jj);
                    const float x = fmaf(-slope, dist, fmaf(__uint_as_float(vs[jj]), scale2, nlse));
                    float pe = exp2f(x);
                    if (!plain) pe = ((bits >> jj) & 1u) ? pe : 0.f;
                    float keepf = 1.f;
                    if (drop_on) keepf = drop_keep(qh, e, drop.thr32) ? drop.keep_scale : 0.f;
                    const float ds = pe * fmaf(__uint_as_float(vp[jj]), keepf, -delta_i);
                    slope_acc = fmaf(ds, dist - e_i, slope_acc);
                    pv[e] = pe * keepf;
                    dsv[e] = ds;
                }
                out_p[g * 2] = pack_bf16x2(pv[0], pv[1]);
                out_p[g * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
                out_ds[g * 2] = pack_bf16x2(dsv[0], dsv[1]);
                out_ds[g * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
            }
            // by now the previous tile's dV / dK / dQ products have long completed: drain its dQ, then reuse the Pd / dS buffers
            if (it > 0) {
                if (drains) stage_dq(it - 1);
                else SPB_MBAR_WAIT(pds_free, (it - 1) & 1);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t off = (((uint32_t)((chunk & 1) * 4 + g)) ^ swz) << 4;
                *reinterpret_cast<uint4*>(sP + off) = make_uint4(out_p[4 * g], out_p[4 * g + 1], out_p[4 * g + 2], out_p[4 * g + 3]);
                *reinterpret_cast<uint4*>(sDS + off) = make_uint4(out_ds[4 * g], out_ds[4 * g + 1], out_ds[4 * g + 2], out_ds[4 * g + 3]);
            }
            fence_proxy_async();                      // ONE fence per tile: Pd / dS -> tensor core, dQ box -> TMA (both async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(pds_full);
                if (drains && it > 0) issue_dq(it - 1);
            }
        }
        if (drains) {
            stage_dq(iters - 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                issue_dq(iters - 1);
                bulk_wait_group<0>();
            }
        }
        // d logslope_h = -slope_h * sum_ij dS_ij (|i-j| - E_i)   (the E_i term removes the row residual of the bf16-rounded delta)
        slope_acc = warp_sum(slope_acc);
        if (lane == 0 && p.dlogslopes != nullptr) atomicAdd(p.dlogslopes + h, -slope_acc * slope_nat);

        // ---- dK / dV of this CTA's keys: chunks 0,1 take the two halves of dV, chunks 2,3 those of dK; thread == key row
        SPB_MBAR_WAIT(dkv_full, 0);
        tc_fence_after();
        {
            const int j = k0 + h * 32 + lane;
            const bool is_v = chunk < 2;
            const int c0 = (chunk & 1) * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (is_v ? TM_DV : TM_DK) + lane_addr + c0, v);
            tmem_ld_wait();
            if (j < T) {
                const float mul = is_v ? 1.f : p.scale;
                __nv_bfloat16* dst = p.dqkv + ((size_t)b * T + j) * p.ld_dqkv + (is_v ? p.vcol : p.kcol) + c0;
#pragma unroll
                for (int d = 0; d < 32; d += 8) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(v[d]) * mul, __uint_as_float(v[d + 1]) * mul);
                    u.y = pack_bf16x2(__uint_as_float(v[d + 2]) * mul, __uint_as_float(v[d + 3]) * mul);
                    u.z = pack_bf16x2(__uint_as_float(v[d + 4]) * mul, __uint_as_float(v[d + 5]) * mul);
                    u.w = pack_bf16x2(__uint_as_float(v[d + 6]) * mul, __uint_as_float(v[d + 7]) * mul);
                    *reinterpret_cast<uint4*>(dst + d) = u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// dq accumulator fp32 [n, 256] -> bf16 columns [0, 256) of dqkv; the accumulator is zeroed again for the next layer
__global__ void attn_dq_convert_kernel(float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int ld, long long n_rows, int cols) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one 8-column piece per thread
    const int per_row = cols / 8;
    if (idx >= n_rows * per_row) return;
    const long long row = idx / per_row;
    const int c = (int)(idx - row * per_row) * 8;
    float4* src = reinterpret_cast<float4*>(acc + row * cols + c);
    const float4 a = src[0], b2 = src[1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b2.x, b2.y); u.w = pack_bf16x2(b2.z, b2.w);
    *reinterpret_cast<uint4*>(dqkv + row * ld + c) = u;
    src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

// tcgen05 backward.  qkv / dout / lse / delta as in spb_attention_bwd; `edist` and `mask_bits` are the side outputs of
// spb_attention_fwd_tc; `dq_acc` is an fp32 [B*T, H*64] scratch accumulator that must be ZERO on entry and is zero again on
// return (so one buffer serves every layer of a stack).  dqkv receives dq | dk | dv in the qkv column layout; dlogslopes [H] is
// accumulated into.  Requires H == 4, dim_head == 64.
extern "C" int spb_attention_bwd_tc(const void* qkv, int ld, const uint32_t* mask_bits, const float* logslopes, const void* dout,
                                    int ld_do, const float* lse, const float* delta, const float* edist, float* dq_acc, void* dqkv,
                                    int ld_dqkv, float* dlogslopes, int B, int T, int H, int dim_head, int causal, float dropout_p,
                                    uint64_t seed, const uint64_t* rng_offset, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return SPB_OK;
    SPB_CHECK_ARG(qkv && logslopes && dout && lse && delta && edist && dq_acc && dqkv, "spb_attention_bwd_tc: null pointer");
    SPB_CHECK_ARG(H == NH && dim_head == DH, "spb_attention_bwd_tc: needs 4 heads of dim 64 (got %d x %d)", H, dim_head);
    SPB_CHECK_ARG(ld % 8 == 0 && ld >= H * DH + 2 * DH && ld_do % 8 == 0 && ld_dqkv % 8 == 0, "spb_attention_bwd_tc: bad leading dims");
    SPB_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "spb_attention_bwd_tc: dropout_p must be in [0,1)");
    CUtensorMap tmQ, tmKV, tmDO, tmDQ;
    int rc = spb_make_tmap_bf16_3d(&tmQ, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)T * ld * 2, DH, QP);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_3d(&tmKV, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)T * ld * 2, DH, TKEY);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_3d(&tmDO, dout, (uint64_t)ld_do, (uint64_t)T, (uint64_t)B, (uint64_t)ld_do * 2, (uint64_t)T * ld_do * 2, DH, QP);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_f32_3d(&tmDQ, dq_acc, (uint64_t)(H * DH), (uint64_t)T, (uint64_t)B, (uint64_t)(H * DH) * 4,
                              (uint64_t)T * (H * DH) * 4, 32, QP);
    if (rc != SPB_OK) return rc;
    BwdParams p;
    p.mask_bits = mask_bits;
    p.words_per_row = ceil_div(T, 32);
    p.logslopes = logslopes;
    p.lse = lse; p.delta = delta; p.edist = edist;
    p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
    p.ld_dqkv = ld_dqkv;
    p.dlogslopes = dlogslopes;
    p.B = B; p.T = T;
    p.scale = 1.f / sqrtf((float)dim_head);
    p.causal = causal;
    p.seed = seed;
    p.rng_offset = rng_offset;
    p.thr32 = host_drop_thr32(dropout_p);
    p.keep_scale = 1.f / (1.f - dropout_p);
    p.kcol = H * DH;
    p.vcol = H * DH + DH;
    SPB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
    attn_bwd_tc_kernel<<<ceil_div(T, TKEY) * B, BWD_THREADS, BWD_SMEM_BYTES, stream>>>(tmQ, tmKV, tmDO, tmDQ, p);
    SPB_CHECK_LAUNCH();
    const long long pieces = (long long)B * T * (H * DH / 8);
    attn_dq_convert_kernel<<<(unsigned)((pieces + 255) / 256), 256, 0, stream>>>(dq_acc, p.dqkv, ld_dqkv, (long long)B * T, H * DH);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
