// Shared pieces of the tcgen05 attention kernels (attention_tc.cu forward, attention_bwd_tc.cu backward).
#pragma once
#include "common.cuh"

namespace attn_tc {

constexpr int DH = 64;         // head dimension
constexpr int QP = 32;         // query positions per 128-row tile
constexpr int NH = 4;          // query heads stacked into the M dimension (MQA: they share K and V)
constexpr int TKEY = 128;      // keys per tile
constexpr float LOG2E = 1.4426950408889634f;

// ---- dropout of the attention weights (attend.py:118-124, `dropout_p` of SDPA).  Counter based, nothing stored: forward and
// backward evaluate the same function of (seed, batch, head, query, key).  ONE 32-bit mix per FOUR adjacent keys of a row; the
// four decisions are the top bits of four odd multiples of it (the same lattice structure as consecutive outputs of a
// multiplicative generator), each compared with a 32-bit threshold: 3.75 integer instructions per element instead of 8.5 for a
// full hash per pair -- these kernels are bound by the element-wise instruction count, not by the tensor pipe.
struct DropParams {
    uint32_t seedmix;      // 64-bit seed (+ device-side step counter) folded by a full splitmix round, once per thread
    uint32_t thr32;        // drop when u < thr32;  0 = dropout off
    uint32_t quarter_t;    // ceil(T / 4): quads per row
    float keep_scale;      // 1 / (1 - p)
};
__device__ __forceinline__ uint32_t drop_seedmix(uint64_t seed, const uint64_t* rng_offset) {
    if (rng_offset != nullptr) seed += *rng_offset * 0x9E3779B97F4A7C15ull;
    return spb_hash32(seed, 0x5bd1e995ull);
}
// per-row base of the quad counter, with the first multiply and the seed folded in: quad q of the row hashes base + q * DROP_K
constexpr uint32_t DROP_K = 0x9E3779B9u;
__device__ __forceinline__ uint32_t drop_row_base(const DropParams& d, uint32_t row_lin) {
    return row_lin * d.quarter_t * DROP_K + d.seedmix;
}
__device__ __forceinline__ uint32_t drop_quad(uint32_t pre) {       // pre = row base + quad index * DROP_K
    uint32_t h = pre ^ (pre >> 15);
    return h * 0x846ca68bu;
}
__device__ __forceinline__ bool drop_keep(uint32_t quad_hash, int k, uint32_t thr32) {      // k = key & 3 (compile-time in the loops)
    const uint32_t mul = k == 0 ? 1u : (k == 1 ? 0x7feb352du : (k == 2 ? 0xc2b2ae35u : 0x27d4eb2fu));
    return quad_hash * mul >= thr32;
}

__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// named barrier among `n_threads` threads (a multiple of 32) of the CTA
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

static inline uint32_t host_drop_thr32(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 4294967296.0;
    return t < 1.0 ? 1u : (t > 4294967295.0 ? 4294967295u : (uint32_t)t);
}

}  // namespace attn_tc
