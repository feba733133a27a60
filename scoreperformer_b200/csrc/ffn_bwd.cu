// Fused feed-forward sub-layer backward, data path, for sm_100a (tcgen05 / TMEM / TMA, CTA pairs):
//     dh   = dy . W2                     dy   bf16 [n, 256]   gradient of the sub-layer output
//     du   = GLU'(u) * dropout-mask . dh u    bf16 [n, 2048]  saved pre-activation (value | gate, bias included)
//     dxn  = du . W1                     dxn  bf16 [n, 256]   gradient of the normalised input
//     db1 += column sums of du
// Reference: the autograd of modules/transformer/feedforward.py:13-22 (GLU), :35-64 (proj -> GLU -> dropout -> out).
// The weight gradients (dW1 = du^T xn, dW2 = dy^T h) contract over all rows and stay ordinary GEMMs over du / h in HBM.
//
// Same shape of kernel as ffn_fwd_pair_kernel (ffn.cu): the two CTAs of a cluster own 256 rows, each stages ITS 128 rows and
// HALF of every weight tile, one thread of the leader issues cta_group::2 MMAs for both.  Per chunk c of 64 hidden units:
//   GEMM1(c)  dh_c[256 x 64]   = dy[256 x 256] . W2T_c^T        W2T = W2 transposed once per step ([1024, 256], K-major rows),
//                                                               each CTA stages 32 of the 64 rows; accumulator in TMEM (x2)
//   GLU'(c)   16 warps, thread == row: u_c (TMA-loaded [128 x (64 value | 64 gate)] tile) and dh_c -> du_c, written IN PLACE
//             into the same 128B-swizzled tile = A operand of GEMM2(c) and source of the TMA store of du
//   GEMM2(c)  dxn[256 x 256]  += du_c[256 x 128] . W1_c         W1_c = value rows [64c, +64) and gate rows [1024 + 64c, +64)
//                                                               of W1 [2048, 256] as an MN-major B operand (no transpose)
// dh never exists in HBM, u is read once and du written once (in place if the caller passes du == u).
// HBM traffic per row: 512 B (dy) + 4 KB (u) + 4 KB (du) + 512 B (dxn); the weights (1.5 MB) stream from L2 once per 256 rows.
#include "attention_tc.cuh"
#include <string.h>

namespace {
using attn_tc::named_bar_sync;
using attn_tc::tmem_ld_32x32b_x16;

constexpr int D = 256;
constexpr int HID = 1024;
constexpr int BMF = 128;          // rows per CTA
constexpr int CH = 64;            // hidden units per chunk
constexpr int NCH = HID / CH;
constexpr int KB = D / 64;        // k-blocks of GEMM1
constexpr int W2_STAGES = 6;      // [32 rows x 64 k] 4 KB
constexpr int W1_STAGES = 2;      // (value | gate) x 2 column atoms x [64 k-rows x 128 B] = 32 KB
constexpr int SG_OFF = 0;                                  // dy tile (own 128 rows): KB x [128 x 64]   64 KB
constexpr int SW2_OFF = SG_OFF + KB * 16384;               //                                           24 KB
constexpr int SW1_OFF = SW2_OFF + W2_STAGES * 4096;        //                                           64 KB
constexpr int SU_OFF = SW1_OFF + W1_STAGES * 32768;        // u -> du tiles: 2 x (value | gate)         64 KB
constexpr int SDB_OFF = SU_OFF + 2 * 32768;                // bias-gradient partial sums, fp32 [2048]    8 KB
constexpr int BAR_OFF = SDB_OFF + 2 * HID * 4;
constexpr int BWD_SMEM_BYTES = BAR_OFF + 512;
static_assert(SW1_OFF % 1024 == 0 && SU_OFF % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(BWD_SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may have");
constexpr int EPI_WARPS = 16;
constexpr int BWD_THREADS = 128 + EPI_WARPS * 32;          // producer, GEMM1 issuer, GEMM2 issuer, forwarding / store thread, epilogue
constexpr uint32_t TM_DX = 0, TM_DH = 256;

struct FfnBwdParams {
    float* db1;                  // fp32 [2 * HID], accumulated into; may be null
    __nv_bfloat16* dxn;          // bf16 [n, ld_dxn]
    int ld_dxn;
    int n_rows;
    uint64_t seed;
    const uint64_t* rng_offset;
    uint32_t thr32;
    float keep_scale;
};

__global__ void __launch_bounds__(BWD_THREADS, 1)
ffn_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmW2T, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmDU, FfnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* a_full = bars + 0;                       // leader
    uint64_t* a_empty = bars + 1;                      // both CTAs (multicast commit)
    uint64_t* w2_full = bars + 2;                      // [W2_STAGES] leader
    uint64_t* w2_empty = w2_full + W2_STAGES;          // both
    uint64_t* w1_full = w2_empty + W2_STAGES;          // [W1_STAGES] leader
    uint64_t* w1_empty = w1_full + W1_STAGES;          // both
    uint64_t* ul_full = w1_empty + W1_STAGES;          // [2] local: the u tile of this CTA's rows has landed
    uint64_t* dh_full = ul_full + 2;                   // [2] both
    uint64_t* dh_empty = dh_full + 2;                  // [2] leader, one arrival per CTA (its forwarding thread)
    uint64_t* du_full = dh_empty + 2;                  // [2] leader, one arrival per CTA
    uint64_t* du_empty = du_full + 2;                  // [2] both: GEMM2 has consumed the du tile
    uint64_t* dx_full = du_empty + 2;                  // both
    uint64_t* dx_empty = dx_full + 1;                  // leader, 2 x 16 warps
    uint64_t* dh_rd = dx_empty + 1;                    // [2] local: dh chunk read out by this CTA's 16 warps
    uint64_t* du_wr = dh_rd + 2;                       // [2] local: du chunk written by this CTA's 16 warps
    uint64_t* st_empty = du_wr + 2;                    // [2] local: the TMA store of the du tile has read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int n_ptiles = (p.n_rows + 2 * BMF - 1) / (2 * BMF);
    const int pt0 = blockIdx.x >> 1, pt_step = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("spb200: ffn_bwd_pair_kernel needs 1024-byte aligned dynamic shared memory\n");
            __trap();
        }
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmW2T);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmU);
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < W2_STAGES; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
        for (int s = 0; s < W1_STAGES; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&ul_full[b], 1);
            mbar_init(&dh_full[b], 1); mbar_init(&dh_empty[b], 2);
            mbar_init(&du_full[b], 2); mbar_init(&du_empty[b], 1);
            mbar_init(&dh_rd[b], EPI_WARPS); mbar_init(&du_wr[b], EPI_WARPS);
            mbar_init(&st_empty[b], 1);
        }
        mbar_init(dx_full, 1); mbar_init(dx_empty, 2 * EPI_WARPS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs: own rows, own weight halves;
        // the weight / dy byte counts land on the LEADER's full barriers, which gate the MMAs; the u tile is a local matter)
        if (lane == 0) {
            int s1 = 0, s2 = 0;
            uint32_t ph1 = 0, ph2 = 0;
            int it = 0;
            auto leader_bar = [](uint64_t* bar) { return smem_u32(bar) & 0xFEFFFFFFu; };     // same offset in the even CTA of the pair
            auto load_w2 = [&](int c) {
                for (int kb = 0; kb < KB; ++kb) {
                    SPB_MBAR_WAIT(&w2_empty[s2], ph2 ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&w2_full[s2], 2 * 4096);
                    tma_load_2d_pair(smem + SW2_OFF + s2 * 4096, &tmW2T, leader_bar(&w2_full[s2]), kb * 64, c * CH + rank * 32);
                    if (++s2 == W2_STAGES) { s2 = 0; ph2 ^= 1u; }
                }
            };
            if (pt0 < n_ptiles) load_w2(0);
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++it) {
                const int m0 = pt * 2 * BMF + rank * BMF;
                SPB_MBAR_WAIT(a_empty, (uint32_t)(it & 1) ^ 1u);
                if (rank == 0) mbar_arrive_expect_tx(a_full, 2 * KB * 16384);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_load_2d_pair(smem + SG_OFF + kb * 16384, &tmG, leader_bar(a_full), kb * 64, m0);
                for (int c = 0; c < NCH; ++c) {
                    const int g = it * NCH + c, b = g & 1;
                    const uint32_t ph = (uint32_t)((g >> 1) & 1);
                    // W2T one chunk ahead (GEMM1 runs up to two chunks ahead of the epilogue)
                    if (c + 1 < NCH) load_w2(c + 1);
                    else if (pt + pt_step < n_ptiles) load_w2(0);
                    // u tile: its buffer is free once GEMM2 and the TMA store of chunk g - 2 have read it
                    SPB_MBAR_WAIT(&du_empty[b], ph ^ 1u);
                    SPB_MBAR_WAIT(&st_empty[b], ph ^ 1u);
                    mbar_arrive_expect_tx(&ul_full[b], 32768);
                    tma_load_2d(smem + SU_OFF + b * 32768, &tmU, &ul_full[b], c * CH, m0);
                    tma_load_2d(smem + SU_OFF + b * 32768 + 16384, &tmU, &ul_full[b], HID + c * CH, m0);
                    // W1 rows of the chunk (k of GEMM2), this CTA's 128 output columns as two 64-column atoms
                    SPB_MBAR_WAIT(&w1_empty[s1], ph1 ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&w1_full[s1], 2 * 32768);
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                        for (int atom = 0; atom < 2; ++atom)
                            tma_load_2d_pair(smem + SW1_OFF + s1 * 32768 + blk * 16384 + atom * 8192, &tmW1, leader_bar(&w1_full[s1]),
                                             rank * 128 + atom * 64, blk * HID + c * CH);
                    if (++s1 == W1_STAGES) { s1 = 0; ph1 ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ GEMM1 issuer (leader CTA): dh chunks for both CTAs
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc1 = umma_idesc_bf16(2 * BMF, CH, false, false);
            const uint32_t sg = smem_u32(smem + SG_OFF), sw2 = smem_u32(smem + SW2_OFF);
            int s2 = 0;
            uint32_t ph2 = 0;
            int it = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++it) {
                SPB_MBAR_WAIT(a_full, (uint32_t)(it & 1));
                for (int c = 0; c < NCH; ++c) {
                    const int g = it * NCH + c, b = g & 1;
                    SPB_MBAR_WAIT(&dh_empty[b], (uint32_t)((g >> 1) & 1) ^ 1u);
                    tc_fence_after();
                    for (int kb = 0; kb < KB; ++kb) {
                        SPB_MBAR_WAIT(&w2_full[s2], ph2);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_pair(tmem_base + TM_DH + b * CH, umma_smem_desc_sw128(sg + kb * 16384 + k * 32, 0, 1024),
                                           umma_smem_desc_sw128(sw2 + s2 * 4096 + k * 32, 0, 1024), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit_pair(&w2_empty[s2]);
                        if (++s2 == W2_STAGES) { s2 = 0; ph2 ^= 1u; }
                    }
                    umma_commit_pair(&dh_full[b]);
                }
                umma_commit_pair(a_empty);                 // every GEMM1 of this tile has been issued: dy may be replaced
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ GEMM2 issuer (leader CTA)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc2 = umma_idesc_bf16(2 * BMF, D, false, true);
            const uint32_t sw1 = smem_u32(smem + SW1_OFF), su = smem_u32(smem + SU_OFF);
            int s1 = 0;
            uint32_t ph1 = 0;
            int it = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++it) {
                for (int c = 0; c < NCH; ++c) {
                    const int g = it * NCH + c, b = g & 1;
                    SPB_MBAR_WAIT(&du_full[b], (uint32_t)((g >> 1) & 1));
                    SPB_MBAR_WAIT(&w1_full[s1], ph1);
                    if (c == 0) SPB_MBAR_WAIT(dx_empty, (uint32_t)(it & 1) ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_pair(tmem_base + TM_DX, umma_smem_desc_sw128(su + b * 32768 + blk * 16384 + k * 32, 0, 1024),
                                           umma_smem_desc_sw128(sw1 + s1 * 32768 + blk * 16384 + k * 2048, 8192, 1024), idesc2,
                                           (c > 0 || blk > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&w1_empty[s1]);
                    umma_commit_pair(&du_empty[b]);
                    if (++s1 == W1_STAGES) { s1 = 0; ph1 ^= 1u; }
                }
                umma_commit_pair(dx_full);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ forwarding / store thread (both CTAs): turns the local
        // barriers of the 16 epilogue warps into ONE cluster-scope arrival each on the leader, and sends the du tiles off
        if (lane == 0) {
            tma_prefetch_desc(&tmDU);
            const uint32_t dh_empty_l[2] = {mapa_u32(&dh_empty[0], 0), mapa_u32(&dh_empty[1], 0)};
            const uint32_t du_full_l[2] = {mapa_u32(&du_full[0], 0), mapa_u32(&du_full[1], 0)};
            int it = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++it) {
                const int m0 = pt * 2 * BMF + rank * BMF;
                for (int c = 0; c < NCH; ++c) {
                    const int g = it * NCH + c, b = g & 1;
                    const uint32_t ph = (uint32_t)((g >> 1) & 1);
                    SPB_MBAR_WAIT(&dh_rd[b], ph);
                    mbar_arrive_cluster_relaxed(dh_empty_l[b]);
                    SPB_MBAR_WAIT(&du_wr[b], ph);
                    mbar_arrive_cluster(du_full_l[b]);
                    if (m0 < p.n_rows) {
                        tma_store_2d(&tmDU, smem + SU_OFF + b * 32768, c * CH, m0);
                        tma_store_2d(&tmDU, smem + SU_OFF + b * 32768 + 16384, HID + c * CH, m0);
                        bulk_commit_group();
                        bulk_wait_group_read<0>();             // the tile is wanted back for the u of chunk g + 2 right away
                    }
                    mbar_arrive(&st_empty[b]);
                }
            }
            bulk_wait_group<0>();
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows): thread == row
        constexpr int HW = CH / 4;                         // 16 hidden units per warp and chunk
        const int ew = warp - 4;
        const int q = warp & 3;                            // TMEM lane quarter this warp may touch
        const int part = ew >> 2;                          // which 16 of the chunk's 64 hidden units
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int r = q * 32 + lane;
        float* sdb = reinterpret_cast<float*>(smem + SDB_OFF);
        for (int i = threadIdx.x - 128; i < 2 * HID; i += EPI_WARPS * 32) sdb[i] = 0.f;
        named_bar_sync(1, EPI_WARPS * 32);
        uint64_t seed = p.seed;
        if (p.rng_offset != nullptr) seed += *p.rng_offset * 0x9E3779B97F4A7C15ull;
        const uint32_t seed32 = spb_seed32(seed);
        const bool drop_on = p.thr32 != 0;
        const uint32_t swz = (uint32_t)(r & 7);
        const uint32_t su_row = smem_u32(smem + SU_OFF) + (uint32_t)(r * 128);
        const uint32_t off0 = (((uint32_t)(part * 2)) ^ swz) << 4, off1 = (((uint32_t)(part * 2 + 1)) ^ swz) << 4;
        const uint32_t dx_empty_l = mapa_u32(dx_empty, 0);
        int it = 0;
        for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++it) {
            const int row = pt * 2 * BMF + rank * BMF + r;
            const bool row_ok = row < p.n_rows;
            for (int c = 0; c < NCH; ++c) {
                const int g = it * NCH + c, b = g & 1;
                const uint32_t ph = (uint32_t)((g >> 1) & 1);
                const int hid0 = c * CH + part * HW;
                const uint32_t base = su_row + (uint32_t)(b * 32768);
                SPB_MBAR_WAIT(&ul_full[b], ph);
                const uint4 v0 = lds_u4(base + off0), v1 = lds_u4(base + off1);
                const uint4 g0 = lds_u4(base + 16384 + off0), g1 = lds_u4(base + 16384 + off1);
                SPB_MBAR_WAIT(&dh_full[b], ph);
                tc_fence_after();
                uint32_t dh[HW];
                tmem_ld_32x32b_x16(tmem_base + TM_DH + b * CH + lane_addr + part * HW, dh);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dh_rd[b]);
                const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                float x[2 * HW];                           // [0, 16) d value, [16, 32) d gate
                const uint32_t quad0 = (uint32_t)row * (uint32_t)(HID >> 2) + (uint32_t)(hid0 >> 2);
#pragma unroll
                for (int j = 0; j < HW; j += 4) {
                    float d[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[e] = __uint_as_float(dh[j + e]);
                    if (drop_on) {
                        const uint32_t qh = spb_quad_hash(seed32, quad0 + (uint32_t)(j >> 2));
#pragma unroll
                        for (int e = 0; e < 4; ++e) d[e] = spb_quad_keep(qh, e, p.thr32) ? d[e] * p.keep_scale : 0.f;
                    }
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const float2 vv = unpack_bf16x2(vw[(j >> 1) + e2]), gg = unpack_bf16x2(gw[(j >> 1) + e2]);
                        const float vs[2] = {vv.x, vv.y}, gs[2] = {gg.x, gg.y};
#pragma unroll
                        for (int e1 = 0; e1 < 2; ++e1) {
                            const int e = 2 * e2 + e1;
                            const float hg = 0.5f * gs[e1];        // sigmoid(g) = 1/2 + 1/2 tanh(g/2)
                            float th;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hg));
                            const float sig = fmaf(0.5f, th, 0.5f);
                            const float silu = gs[e1] * sig;
                            x[j + e] = d[e] * silu;
                            x[HW + j + e] = d[e] * vs[e1] * fmaf(silu, 1.f - sig, sig);     // silu' = sig + silu (1 - sig)
                        }
                    }
                }
                // du over u, in place (this thread read exactly these 64 bytes)
                sts_u4(base + off0, make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7])));
                sts_u4(base + off1, make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15])));
                sts_u4(base + 16384 + off0, make_uint4(pack_bf16x2(x[16], x[17]), pack_bf16x2(x[18], x[19]), pack_bf16x2(x[20], x[21]), pack_bf16x2(x[22], x[23])));
                sts_u4(base + 16384 + off1, make_uint4(pack_bf16x2(x[24], x[25]), pack_bf16x2(x[26], x[27]), pack_bf16x2(x[28], x[29]), pack_bf16x2(x[30], x[31])));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&du_wr[b]);
                // bias gradient: column sums over the warp's 32 rows by a transposing butterfly (31 shuffles for 32 columns);
                // lane L ends up with the total of column L, added to the CTA's partial sums
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const bool upper = (lane & s) != 0;
#pragma unroll
                    for (int i = 0; i < s; ++i) {
                        const float send = upper ? x[i] : x[i + s];
                        const float keep = upper ? x[i + s] : x[i];
                        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                    }
                }
                atomicAdd(&sdb[(lane < HW ? 0 : HID - HW) + hid0 + lane], x[0]);
            }
            // ---- dxn: this warp's 64 columns of its 32 rows, bf16
            __nv_bfloat16* dst = p.dxn + (size_t)(row_ok ? row : 0) * p.ld_dxn + part * 64;
            SPB_MBAR_WAIT(dx_full, (uint32_t)(it & 1));
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + TM_DX + lane_addr + part * 64 + half * 32, v);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(dx_empty_l);
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + half * 32 + j * 8) =
                            make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                }
            }
        }
        named_bar_sync(1, EPI_WARPS * 32);
        if (p.db1 != nullptr && pt0 < n_ptiles)
            for (int i = threadIdx.x - 128; i < 2 * HID; i += EPI_WARPS * 32) atomicAdd(p.db1 + i, sdb[i]);
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();              // the peer's smem / barriers must outlive the leader's last MMA and commit
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// dst_z[c, r] = src_z[r, c] for up to 16 bf16 matrices of one shape (the once-per-step W2 -> W2T of the kernel above, all layers
// of a stack in one launch); dst_z = dst + z * rows * cols
struct TransposeBatch {
    const __nv_bfloat16* src[16];
};
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(TransposeBatch b, __nv_bfloat16* __restrict__ dst_all, int rows, int cols) {
    __shared__ __nv_bfloat16 tile[64][66];
    const __nv_bfloat16* __restrict__ src = b.src[blockIdx.z];
    __nv_bfloat16* __restrict__ dst = dst_all + (size_t)blockIdx.z * rows * cols;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int r = i >> 6, c = i & 63;
        tile[r][c] = (r0 + r < rows && c0 + c < cols) ? src[(size_t)(r0 + r) * cols + c0 + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int c = i >> 6, r = i & 63;
        if (r0 + r < rows && c0 + c < cols) dst[(size_t)(c0 + c) * rows + r0 + r] = tile[r][c];
    }
}

}  // namespace

extern "C" int spb_transpose_bf16(const void* const* srcs, int n_mats, void* dst, int rows, int cols, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0 || n_mats <= 0) return SPB_OK;
    SPB_CHECK_ARG(srcs && dst && n_mats <= 16, "spb_transpose_bf16: 1..16 matrices per call (got %d)", n_mats);
    TransposeBatch b;
    for (int i = 0; i < 16; ++i) b.src[i] = reinterpret_cast<const __nv_bfloat16*>(srcs[i < n_mats ? i : 0]);
    dim3 grid(ceil_div(cols, 64), ceil_div(rows, 64), n_mats);
    transpose_bf16_kernel<<<grid, 256, 0, stream>>>(b, reinterpret_cast<__nv_bfloat16*>(dst), rows, cols);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// Fused feed-forward backward, data path (see the header of this file).  dy bf16 [n, 256] (ld_dy elements per row), w2t bf16
// [1024, 256] = nn.Linear(inner, dim).weight TRANSPOSED, w1 bf16 [2048, 256] (value rows, then gate rows), u bf16 [n, 2048] as
// saved by spb_ffn_fwd, du bf16 [n, 2048] (may be u itself), db1 fp32 [2048] accumulated into (or NULL), dxn bf16 [n, ld_dxn].
// Dropout arguments as given to spb_ffn_fwd.
extern "C" int spb_ffn_bwd(const void* dy, int ld_dy, const void* w2t, const void* w1, const void* u, void* du, float* db1, void* dxn,
                           int ld_dxn, int n_rows, int dim, int hidden, float dropout_p, uint64_t seed, const uint64_t* rng_offset,
                           cudaStream_t stream) {
    if (n_rows <= 0) return SPB_OK;
    SPB_CHECK_ARG(dy && w2t && w1 && u && du && dxn, "spb_ffn_bwd: null pointer");
    SPB_CHECK_ARG(dim == D && hidden == HID, "spb_ffn_bwd: built for dim 256 / hidden 1024 (got %d / %d)", dim, hidden);
    SPB_CHECK_ARG(ld_dy % 8 == 0 && ld_dxn % 8 == 0, "spb_ffn_bwd: leading dimensions must be multiples of 8");
    SPB_CHECK_ARG((reinterpret_cast<uintptr_t>(dxn) & 15) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(du) & 15) == 0,
                  "spb_ffn_bwd: u / du / dxn must be 16-byte aligned");
    SPB_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "spb_ffn_bwd: dropout_p must be in [0,1)");
    FfnBwdParams p;
    p.db1 = db1;
    p.dxn = reinterpret_cast<__nv_bfloat16*>(dxn);
    p.ld_dxn = ld_dxn;
    p.n_rows = n_rows;
    p.seed = seed; p.rng_offset = rng_offset;
    p.thr32 = spb_drop_thr32(dropout_p);
    p.keep_scale = 1.f / (1.f - dropout_p);
    CUtensorMap tmG, tmW2T, tmW1, tmU, tmDU;
    int rc = spb_make_tmap_bf16_2d(&tmG, dy, (uint64_t)D, (uint64_t)n_rows, (uint64_t)ld_dy * 2, 64, BMF);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_2d(&tmW2T, w2t, (uint64_t)D, (uint64_t)HID, (uint64_t)D * 2, 64, 32);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_2d(&tmW1, w1, (uint64_t)D, (uint64_t)(2 * HID), (uint64_t)D * 2, 64, 64);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_2d(&tmU, u, (uint64_t)(2 * HID), (uint64_t)n_rows, (uint64_t)(2 * HID) * 2, 64, BMF);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_2d(&tmDU, du, (uint64_t)(2 * HID), (uint64_t)n_rows, (uint64_t)(2 * HID) * 2, 64, BMF);
    if (rc != SPB_OK) return rc;
    SPB_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
    const int n_ptiles = ceil_div(n_rows, 2 * BMF);
    const int max_pairs = spb_num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (n_ptiles < max_pairs ? n_ptiles : max_pairs));
    cfg.blockDim = dim3(BWD_THREADS);
    cfg.dynamicSmemBytes = BWD_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_bwd_pair_kernel, tmG, tmW2T, tmW1, tmU, tmDU, p));
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
