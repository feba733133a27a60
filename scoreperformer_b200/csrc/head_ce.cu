// Tied output head of ONE tuple field fused with its masked cross-entropy:
//
//   logits[n, V] = e_f[n, 128] * table_f[V, 128]^T          models/scoreperformer/embeddings.py:345-353 (weights tied to the
//                                                            input embedding table of the field)
//   loss_sum += sum_n CE(logits[n, :], label[n]),  count += #labelled rows      wrappers.py:49-59 (ignore_index = -100)
//   dlogits[n, :] = softmax(logits[n, :]) - onehot(label[n])   (bf16, zero for ignored rows; consumed by the two backward GEMMs)
//
// The logits never leave the SM: tcgen05.mma accumulates a 128 x V tile in TMEM (V <= 256 is ONE n-tile), and the epilogue
// thread that owns an accumulator row makes three passes over its TMEM row (max, sum of exp2, gradient).  Per CTA:
//   warp 0        TMA producer: the field's table slice once (resident for the whole launch), then e tiles, 2-stage ring
//   warp 1        MMA issuer: 8 tcgen05.mma (K = 128) per tile, accumulator double-buffered in TMEM (2 x 256 columns)
//   warps 2..5    epilogue group 0: even tiles of this CTA  (thread = row; warp w reads TMEM lanes 32*(w%4)..)
//   warps 6..9    epilogue group 1: odd tiles
// Algorithmic HBM bytes per note-tuple and field: 256 B of e in, 8 B label in, 2*V B of dlogits out (none in inference).
#include "common.cuh"

namespace {

constexpr int HC_BM = 128;
constexpr int HC_K = 128;                 // embedding width of a field
constexpr int HC_BK = 64;
constexpr int HC_THREADS = 320;
constexpr int HC_A_STAGE = HC_BM * HC_K * 2;      // 32 KB: both k-blocks of one e tile
constexpr int HC_MAXN = 256;
constexpr float HC_LOG2E = 1.4426950408889634f;

struct HeadCeParams {
    int n_rows, V, n_mma;       // n_mma = V rounded up to 16 (UMMA N)
    uint32_t idesc;
    const int64_t* labels;
    int ld_lab;
    long long ignore_index;
    float* loss_sum;
    float* count;
    __nv_bfloat16* dlogits;     // [n_rows, ld_d] or null
    int ld_d;
    int* argmax;                // [n_rows] or null
    const float* token_values;  // fp32 [V]: the value a token of this field stands for (evaluator distances), or null
    float* stats;               // [3] accumulated over labelled rows: #(argmax == label), sum |tv[argmax] - tv[label]|,
                                //     sum_v softmax_v |tv[label] - tv[v]|  (evaluator.py:38-46,72-104); or null
};

__global__ void __launch_bounds__(HC_THREADS, 1)
head_ce_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HeadCeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                                   // 2 k-blocks x [HC_MAXN rows x 128 B]
    uint8_t* sA = smem + 2 * HC_MAXN * 128;               // 2 stages x 2 k-blocks x [128 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * HC_A_STAGE);
    uint64_t* b_full = bars;            // [1]
    uint64_t* a_full = bars + 1;        // [2]
    uint64_t* a_empty = bars + 3;       // [2]
    uint64_t* acc_full = bars + 5;      // [2]
    uint64_t* acc_empty = bars + 7;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* s_tv = reinterpret_cast<float*>(bars + 16);          // token values of the field (HC_MAXN floats)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = (p.n_rows + HC_BM - 1) / HC_BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);          // the four warps of one epilogue group
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (p.stats != nullptr)
        for (int i = threadIdx.x; i < HC_MAXN; i += HC_THREADS) s_tv[i] = (p.token_values != nullptr && i < p.V) ? p.token_values[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // table slice: rows beyond V are zero-filled by TMA, so the padded logits are exactly 0 and get masked below
            mbar_arrive_expect_tx(b_full, 2 * p.n_mma * 128);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sB + kb * HC_MAXN * 128, &tmB, b_full, kb * HC_BK, 0);
            int s = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&a_empty[s], phase ^ 1);
                mbar_arrive_expect_tx(&a_full[s], HC_A_STAGE);
                for (int kb = 0; kb < 2; ++kb) tma_load_2d(sA + s * HC_A_STAGE + kb * (HC_BM * 128), &tmA, &a_full[s], kb * HC_BK, t * HC_BM);
                if (++s == 2) { s = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(b_full, 0);
            int s = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                mbar_wait(&a_full[s], phase);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * HC_MAXN);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t a_addr = smem_u32(sA + s * HC_A_STAGE + kb * (HC_BM * 128));
                    const uint32_t b_addr = smem_u32(sB + kb * HC_MAXN * 128);
#pragma unroll
                    for (int k = 0; k < HC_BK / 16; ++k) {
                        const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32, 0, 1024);
                        const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                        umma_bf16(tmem_d, da, db, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&a_empty[s]);
                umma_commit(&acc_full[acc]);
                if (++s == 2) { s = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int grp = (warp - 2) >> 2;                  // epilogue group = TMEM accumulator buffer
        const int q = warp & 3;                           // TMEM lane quarter this warp may read
        const uint32_t tmem_acc = tmem_base + (uint32_t)(grp * HC_MAXN) + ((uint32_t)(q * 32) << 16);
        const int V = p.V;
        const int nchunk = (V + 31) / 32;
        float loss_acc = 0.f, cnt_acc = 0.f;
        float hit_acc = 0.f, dist_acc = 0.f, wdist_acc = 0.f;      // evaluator statistics of the labelled rows
        const bool stats_on = p.stats != nullptr;
        const bool wdist_on = stats_on && p.token_values != nullptr;
        uint32_t ph = 0;
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            const int row = t * HC_BM + q * 32 + lane;
            long long lab = p.ignore_index;
            if (row < p.n_rows) lab = p.labels[(size_t)row * p.ld_lab];
            const bool use = row < p.n_rows && lab != p.ignore_index && lab >= 0 && lab < V;
            mbar_wait(&acc_full[grp], ph);
            ph ^= 1;
            tc_fence_after();
            // pass 1: row maximum (and arg max) over the V real columns
            float mx = -INFINITY;
            int amax = 0;
            for (int c = 0; c < nchunk; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = __uint_as_float(v[j]);
                    if (c * 32 + j < V && x > mx) { mx = x; amax = c * 32 + j; }
                }
            }
            if (p.argmax != nullptr && row < p.n_rows) p.argmax[row] = amax;
            const float tv_lab = (wdist_on && use) ? s_tv[(int)lab] : 0.f;
            if (stats_on && use) {
                hit_acc += amax == (int)lab ? 1.f : 0.f;
                dist_acc += fabsf(s_tv[amax] - tv_lab);
            }
            // pass 2: sum of exponentials and the label's logit (and, for the evaluator, the expected |value - target|)
            float se = 0.f, xl = 0.f, wd = 0.f;
            const float mxs = mx * HC_LOG2E;
            for (int c = 0; c < nchunk; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = __uint_as_float(v[j]);
                    if (c * 32 + j < V) {
                        const float ex = exp2f(fmaf(x, HC_LOG2E, -mxs));
                        se += ex;
                        if (wdist_on) wd = fmaf(ex, fabsf(tv_lab - s_tv[c * 32 + j]), wd);
                    }
                    if (c * 32 + j == (int)lab) xl = x;
                }
            }
            if (use) {
                loss_acc += mx + __logf(se) - xl;
                cnt_acc += 1.f;
                if (wdist_on) wdist_acc += wd / se;
            }
            // pass 3: gradient rows (bf16), 16 bytes at a time; columns [V, ld_d) are written as zeros
            if (p.dlogits != nullptr) {
                const float inv = use ? 1.f / se : 0.f;
                __nv_bfloat16* dr = p.dlogits + (size_t)row * p.ld_d;
                for (int c = 0; c < nchunk; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    if (row < p.n_rows) {
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const int c0 = c * 32 + j8 * 8;
                            if (c0 >= p.ld_d) break;
                            float gq[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int col = c0 + e;
                                const float x = __uint_as_float(v[j8 * 8 + e]);
                                float gv = (col < V) ? exp2f(fmaf(x, HC_LOG2E, -mxs)) * inv : 0.f;
                                if (use && col == (int)lab) gv -= 1.f;
                                gq[e] = gv;
                            }
                            const uint4 o = make_uint4(pack_bf16x2(gq[0], gq[1]), pack_bf16x2(gq[2], gq[3]), pack_bf16x2(gq[4], gq[5]),
                                                       pack_bf16x2(gq[6], gq[7]));
                            *reinterpret_cast<uint4*>(dr + c0) = o;       // ld_d % 8 == 0 (host-checked)
                        }
                    }
                }
            }
            // all TMEM reads of this warp are done: give the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[grp]);
        }
        loss_acc = warp_sum(loss_acc);
        cnt_acc = warp_sum(cnt_acc);
        if (stats_on) {
            hit_acc = warp_sum(hit_acc);
            dist_acc = warp_sum(dist_acc);
            wdist_acc = warp_sum(wdist_acc);
        }
        if (lane == 0 && cnt_acc > 0.f) {
            atomicAdd(p.loss_sum, loss_acc);
            atomicAdd(p.count, cnt_acc);
            if (stats_on) {
                atomicAdd(p.stats, hit_acc);
                atomicAdd(p.stats + 1, dist_acc);
                atomicAdd(p.stats + 2, wdist_acc);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

constexpr int HC_SMEM = 2 * HC_MAXN * 128 + 2 * HC_A_STAGE + 1024 + 128 + HC_MAXN * 4;

}  // namespace

// e bf16 [n_rows, lde] (the 128 columns of this field), table bf16 [V, 128] (ldt = row stride), labels int64 with stride
// ld_lab.  loss_sum / count are ACCUMULATED into.  dlogits bf16 [n_rows, ld_d] (ld_d >= V, multiple of 8) and argmax int32
// [n_rows] are optional.  `stats` (fp32 [3], accumulated; may be NULL) receives the evaluator's sums over the labelled rows:
// #(argmax == label), sum |tv[argmax] - tv[label]| and sum_v softmax_v |tv[label] - tv[v]| with tv = `token_values` (fp32 [V];
// NULL: only the hit count) -- models/scoreperformer/evaluator.py:38-46,72-104 without materialising the logits.
extern "C" int spb_head_ce(const void* e, int lde, const void* table, int ldt, int V, const int64_t* labels, int ld_lab,
                           long long ignore_index, float* loss_sum, float* count, void* dlogits, int ld_d, int* argmax,
                           const float* token_values, float* stats, int n_rows, cudaStream_t stream) {
    if (n_rows <= 0) return SPB_OK;
    SPB_CHECK_ARG(e && table && labels && loss_sum && count, "spb_head_ce: null pointer");
    SPB_CHECK_ARG(V > 0 && V <= HC_MAXN, "spb_head_ce: field vocabulary must be in 1..%d, got %d", HC_MAXN, V);
    SPB_CHECK_ARG(lde % 8 == 0 && ldt % 8 == 0, "spb_head_ce: lde / ldt must be multiples of 8 (TMA 16 B rule)");
    SPB_CHECK_ARG(dlogits == nullptr || (ld_d >= V && ld_d % 8 == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0),
                  "spb_head_ce: dlogits needs ld_d >= V, ld_d %% 8 == 0 and a 16-byte aligned base");
    const int n_mma = ceil_div(V, 16) * 16;
    CUtensorMap tmA, tmB;
    int rc = spb_make_tmap_bf16_2d(&tmA, e, (uint64_t)HC_K, (uint64_t)n_rows, (uint64_t)lde * 2, HC_BK, HC_BM);
    if (rc != SPB_OK) return rc;
    rc = spb_make_tmap_bf16_2d(&tmB, table, (uint64_t)HC_K, (uint64_t)V, (uint64_t)ldt * 2, HC_BK, (uint32_t)n_mma);
    if (rc != SPB_OK) return rc;
    HeadCeParams p;
    p.n_rows = n_rows; p.V = V; p.n_mma = n_mma;
    p.idesc = umma_idesc_bf16(HC_BM, n_mma, false, false);
    p.labels = labels; p.ld_lab = ld_lab; p.ignore_index = ignore_index;
    p.loss_sum = loss_sum; p.count = count;
    p.dlogits = reinterpret_cast<__nv_bfloat16*>(dlogits); p.ld_d = ld_d; p.argmax = argmax;
    p.token_values = token_values; p.stats = stats;
    static bool configured = false;
    if (!configured) {
        SPB_CHECK_CUDA(cudaFuncSetAttribute(head_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
        configured = true;
    }
    const int tiles = ceil_div(n_rows, HC_BM);
    const int grid = tiles < spb_num_sms() ? tiles : spb_num_sms();
    head_ce_kernel<<<grid, HC_THREADS, HC_SMEM, stream>>>(tmA, tmB, p);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
