// Table-gradient scatter of the fused tuple-embedding + LayerNorm backward on tensor cores.
//
// Reference semantics: modules/transformer/embeddings.py:80-137 (per-field nn.Embedding with padding_idx=0, concatenated,
// LayerNorm'ed); autograd scatters dx[n, f*128:(f+1)*128] into row tokens[n, f] of field f's table.
//
// The scatter is a contraction with a one-hot matrix,  dtable_f^T [128, V_f] = dx_f^T [128, n] * onehot_f [n, V_f],  so it runs
// as bf16 mma.sync.m16n8k16 with fp32 accumulators held in registers for the whole token range of a CTA:
//   * grid (chunks, F): CTA (c, f) owns field f and a contiguous range of note-tuples;
//   * per 64-tuple tile all 512 threads rebuild dx = rstd * (dy*w - c1 - xhat*c2) from dy, the table gather and the per-row
//     statistics, and park it as bf16 in shared memory (row = tuple, padded to 272 B so ldmatrix.trans is conflict-free);
//   * warp (dim half, vocabulary group) then walks the tile 16 tuples at a time: A = dx^T via ldmatrix.trans, B = the one-hot
//     block built in registers from the token ids (1.0 = 0x3F80), skipped when no tuple of the step hits the block;
//   * one flush of the register accumulators with fp32 atomics at the end (zero entries skipped).
// Measured on B200 at 32k x 12-field tuples: 125 us against 201 us for the shared-memory float-atomic version it replaces.  Vocabulary rows >= 256 (only the Bar field can have them) take a direct global-atomic path.
#include "common.cuh"

namespace {

constexpr int ES_THREADS = 512;
constexpr int ES_WARPS = ES_THREADS / 32;
constexpr int ES_TILE = 64;            // tuples per shared-memory tile
constexpr int ES_ROW_BYTES = 272;      // 128 bf16 + 16 B pad
constexpr int ES_MAXNT = 2;            // 8-row vocabulary blocks per warp (16 warps -> 256 rows)
constexpr int ES_VMAX = 8 * ES_WARPS * ES_MAXNT;

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Work list: field f owns blocks [first[f], first[f+1]) and each of them rows_per_chunk[f] tuples.
struct ScatterPlan {
    int first[MAX_FIELDS + 1];
    int rows_per_chunk[MAX_FIELDS];
};

struct RawTile {          // what one thread holds of the next tile: one tuple, 16 dims
    long long tok;        // raw id; validated when the tile is produced
    float4 x[4];
    uint4 dy[2];
    float mean, rstd, c1, c2;
};

__global__ void __launch_bounds__(ES_THREADS, 1)
embed_scatter_mma_kernel(const __nv_bfloat16* __restrict__ dy, int ld_dy, const int64_t* __restrict__ tokens, int ld_tok,
                         const float* __restrict__ table, FieldTable ft, const float* __restrict__ w,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ c1_in,
                         const float* __restrict__ c2_in, float* __restrict__ dtable, int n_rows, ScatterPlan plan) {
    // two tiles: tile i+1 is produced while stragglers still read tile i, so one barrier per tile is enough
    __shared__ __align__(16) uint8_t tile[2][ES_TILE * ES_ROW_BYTES];
    __shared__ __align__(8) int tok_s[2][ES_TILE];
    int f = 0;
    while (f + 1 < ft.n_fields && (int)blockIdx.x >= plan.first[f + 1]) ++f;
    const int V = ft.size[f], off = ft.offset[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = ((int)blockIdx.x - plan.first[f]) * plan.rows_per_chunk[f];
    const int r1 = min(n_rows, r0 + plan.rows_per_chunk[f]);
    // producer mapping: 8 threads per tuple, 16 dims each
    const int p_tok = threadIdx.x >> 3;
    const int p_dim = (threadIdx.x & 7) * 16;
    float4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wv[i] = *reinterpret_cast<const float4*>(w + f * 128 + p_dim + 4 * i);

    // warp `warp` accumulates vocabulary blocks warp and warp+16 (8 rows each) for all 128 dims (8 m-tiles)
    float acc[8][ES_MAXNT][4];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < ES_MAXNT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.f;

    // token ids run two tiles ahead of the tensor cores, the gathered operands one tile ahead: every global round trip of
    // the dependent chain (id -> table row) overlaps a full tile of work
    auto fetch_tok = [&](int base) -> long long {
        const int row = base + p_tok;
        return row < r1 ? tokens[(size_t)row * ld_tok + f] : 0ll;
    };
    auto fetch = [&](int base, long long tk, RawTile& rt) {
        const int row = base + p_tok;
        rt.tok = tk;
        if (tk > 0 && tk < V) {                               // PAD (0) and out-of-range ids get no gradient
            const float* xr = table + (size_t)(off + tk) * 128 + p_dim;
#pragma unroll
            for (int i = 0; i < 4; ++i) rt.x[i] = *reinterpret_cast<const float4*>(xr + 4 * i);
            const __nv_bfloat16* dr = dy + (size_t)row * ld_dy + f * 128 + p_dim;
            rt.dy[0] = *reinterpret_cast<const uint4*>(dr);
            rt.dy[1] = *reinterpret_cast<const uint4*>(dr + 8);
            rt.mean = mean_in[row]; rt.rstd = rstd_in[row]; rt.c1 = c1_in[row]; rt.c2 = c2_in[row];
        }
    };

    RawTile cur;
    fetch(r0, fetch_tok(r0), cur);
    long long tok_next = fetch_tok(r0 + ES_TILE);
    int buf = 0;
    for (int base = r0; base < r1; base += ES_TILE, buf ^= 1) {
        const uint32_t tile_a = smem_u32(tile[buf]);
        // ---- produce: dx of this tile -> shared memory (bf16; dy itself is bf16, so this rounding is at the level of the input)
        {
            uint4 o[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
            int tk = (cur.tok > 0 && cur.tok < V) ? (int)cur.tok : -1;
            if (tk >= 0) {
                const uint32_t dyw[8] = {cur.dy[0].x, cur.dy[0].y, cur.dy[0].z, cur.dy[0].w, cur.dy[1].x, cur.dy[1].y, cur.dy[1].z, cur.dy[1].w};
                float dx[16];
                const float a = cur.rstd * cur.c2, b = cur.c1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 d0 = unpack_bf16x2(dyw[2 * i]), d1 = unpack_bf16x2(dyw[2 * i + 1]);
                    dx[4 * i + 0] = cur.rstd * (d0.x * wv[i].x - b - (cur.x[i].x - cur.mean) * a);
                    dx[4 * i + 1] = cur.rstd * (d0.y * wv[i].y - b - (cur.x[i].y - cur.mean) * a);
                    dx[4 * i + 2] = cur.rstd * (d1.x * wv[i].z - b - (cur.x[i].z - cur.mean) * a);
                    dx[4 * i + 3] = cur.rstd * (d1.y * wv[i].w - b - (cur.x[i].w - cur.mean) * a);
                }
                if (tk >= ES_VMAX) {
                    // rows beyond the register-resident vocabulary range: rare, straight to global memory
                    float* dst = dtable + (size_t)(off + tk) * 128 + p_dim;
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(dst + i, dx[i]);
                    tk = -1;
                } else {
                    o[0] = make_uint4(pack_bf16x2(dx[0], dx[1]), pack_bf16x2(dx[2], dx[3]), pack_bf16x2(dx[4], dx[5]), pack_bf16x2(dx[6], dx[7]));
                    o[1] = make_uint4(pack_bf16x2(dx[8], dx[9]), pack_bf16x2(dx[10], dx[11]), pack_bf16x2(dx[12], dx[13]), pack_bf16x2(dx[14], dx[15]));
                }
            }
            const uint32_t dst = tile_a + (uint32_t)(p_tok * ES_ROW_BYTES + p_dim * 2);
            sts_u4(dst, o[0]);
            sts_u4(dst + 16, o[1]);
            if ((threadIdx.x & 7) == 0) tok_s[buf][p_tok] = tk;
        }
        __syncthreads();
        // ---- prefetch the next tile's operands while the tensor cores chew on this one
        if (base + ES_TILE < r1) {
            fetch(base + ES_TILE, tok_next, cur);
            tok_next = fetch_tok(base + 2 * ES_TILE);
        }
        // ---- consume: 4 steps of 16 tuples
#pragma unroll 1
        for (int ks = 0; ks < ES_TILE / 16; ++ks) {
            const int2 ta = *reinterpret_cast<const int2*>(&tok_s[buf][ks * 16 + 2 * t]);
            const int2 tb = *reinterpret_cast<const int2*>(&tok_s[buf][ks * 16 + 8 + 2 * t]);
            uint32_t b0[ES_MAXNT], b1[ES_MAXNT];
            bool hit[ES_MAXNT];
            bool any = false;
#pragma unroll
            for (int n = 0; n < ES_MAXNT; ++n) {
                const int v = (warp + ES_WARPS * n) * 8 + g;
                b0[n] = (ta.x == v ? 0x3F80u : 0u) | (ta.y == v ? 0x3F800000u : 0u);
                b1[n] = (tb.x == v ? 0x3F80u : 0u) | (tb.y == v ? 0x3F800000u : 0u);
                hit[n] = __any_sync(0xffffffffu, (b0[n] | b1[n]) != 0u);
                any |= hit[n];
            }
            if (!any) continue;                               // no tuple of this step lands in this warp's blocks
            const int tok_row = ks * 16 + (lane & 7) + ((lane >> 4) << 3);
            const uint32_t src = tile_a + (uint32_t)(tok_row * ES_ROW_BYTES + (((lane >> 3) & 1) << 4));
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                uint32_t a[4];
                ldsm_x4_t(src + (uint32_t)(m * 32), a[0], a[1], a[2], a[3]);
#pragma unroll
                for (int n = 0; n < ES_MAXNT; ++n)
                    if (hit[n]) mma16816(acc[m][n], a, b0[n], b1[n]);
            }
        }
    }
    // ---- flush: acc[m][n] holds rows dim = m*16 + g (+8), columns v = (warp + 16n)*8 + 2t (+1)
#pragma unroll
    for (int n = 0; n < ES_MAXNT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int v = (warp + ES_WARPS * n) * 8 + 2 * t + e;
            if (v >= V || v == 0) continue;
            float* dst = dtable + (size_t)(off + v) * 128 + g;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (acc[m][n][e] != 0.f) atomicAdd(dst + m * 16, acc[m][n][e]);
                if (acc[m][n][2 + e] != 0.f) atomicAdd(dst + m * 16 + 8, acc[m][n][2 + e]);
            }
        }
    }
}

}  // namespace

// Called by spb_embed_ln_bwd (rowops.cu) after the statistics pass.  dtable is accumulated into.
int spb_embed_scatter_mma(const __nv_bfloat16* dy, int ld_dy, const int64_t* tokens, int ld_tok, const float* table,
                          const FieldTable& ft, const float* w, const float* mean, const float* rstd, const float* c1,
                          const float* c2, float* dtable, int n_rows, cudaStream_t stream) {
    SPB_CHECK_ARG(ld_dy % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0, "embed scatter: dy must be 16-byte aligned with ld %% 8 == 0");
    // Equal shares measured best on B200: small vocabularies concentrate their MMAs in one or two warps, large ones spread
    // them over all sixteen, and the two effects roughly cancel (a V-proportional split was 40% slower).
    float cost[MAX_FIELDS], total = 0.f;
    for (int f = 0; f < ft.n_fields; ++f) {
        cost[f] = 1.f;
        total += cost[f];
    }
    ScatterPlan plan;
    int blocks = 0;
    const int budget = spb_num_sms();
    for (int f = 0; f < ft.n_fields; ++f) {
        int chunks = (int)(budget * cost[f] / total);
        if (chunks < 1) chunks = 1;
        int rpc = ceil_div(ceil_div(n_rows, chunks), ES_TILE) * ES_TILE;
        chunks = ceil_div(n_rows, rpc);
        plan.first[f] = blocks;
        plan.rows_per_chunk[f] = rpc;
        blocks += chunks;
    }
    plan.first[ft.n_fields] = blocks;
    embed_scatter_mma_kernel<<<blocks, ES_THREADS, 0, stream>>>(dy, ld_dy, tokens, ld_tok, table, ft, w, mean, rstd, c1, c2, dtable, n_rows,
                                                               plan);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
