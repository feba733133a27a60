// Microbenchmark: cycles for chains of small tcgen05.mma (M=128, K=16, bf16) issued by one thread, operands in 128B-swizzled smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scoreperformer_b200/csrc -o build/umma_latency \
//        tests/cuda/umma_latency.cu scoreperformer_b200/csrc/api.cu -lcuda
// Per N and per number of independent accumulators ACCS: a fully unrolled run of 32 instructions issued round-robin over the
// accumulators (each accumulator's instructions form a dependent chain); "issue" = cycles until the last one is issued, "done" =
// until the commit has arrived on the mbarrier.
#include "common.cuh"
#include <cstdio>

template <int N, int ACCS, bool PRE>
__device__ __forceinline__ void chain(uint32_t tm, uint32_t sa, uint32_t sb, uint64_t* bar, uint32_t& phase, long long* out) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { da[k] = umma_smem_desc_sw128(sa + k * 32, 0, 1024); db[k] = umma_smem_desc_sw128(sb + k * 32, 0, 1024); }
    for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int it = 0; it < 32; ++it) {
            const int a = it % ACCS;
            if (PRE) umma_bf16(tm + a * N, da[it & 3], db[it & 3], idesc, it >= ACCS ? 1u : 0u);
            else umma_bf16(tm + a * N, umma_smem_desc_sw128(sa + (it & 3) * 32 + a * 16384, 0, 1024),
                           umma_smem_desc_sw128(sb + (it & 3) * 32, 0, 1024), idesc, it >= ACCS ? 1u : 0u);
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1;
        const long long t2 = clock64();
        if (rep == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
}

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (warp == 0 && lane == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 65536);
        uint32_t phase = 0;
        chain<N, 1, false>(tm, sa, sb, bar, phase, out + 0);
        if (2 * N <= 512) chain<N, 2, false>(tm, sa, sb, bar, phase, out + 2);
        if (4 * N <= 512) chain<N, 4, false>(tm, sa, sb, bar, phase, out + 4);
        chain<N, 1, true>(tm, sa, sb, bar, phase, out + 6);
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
        for (int rep = 0; rep < 2; ++rep) {       // one MMA + commit + wait: the round-trip latency
            const long long t0 = clock64();
            umma_bf16(tm, umma_smem_desc_sw128(sa, 0, 1024), umma_smem_desc_sw128(sb, 0, 1024), idesc, 0u);
            umma_commit(bar);
            mbar_wait(bar, phase);
            phase ^= 1;
            if (rep == 1) out[8] = clock64() - t0;
        }
        for (int rep = 0; rep < 2; ++rep) {       // 4 dependent MMAs + commit + wait
            const long long t0 = clock64();
#pragma unroll
            for (int it = 0; it < 4; ++it)
                umma_bf16(tm, umma_smem_desc_sw128(sa + it * 32, 0, 1024), umma_smem_desc_sw128(sb + it * 32, 0, 1024), idesc, it ? 1u : 0u);
            umma_commit(bar);
            mbar_wait(bar, phase);
            phase ^= 1;
            if (rep == 1) out[9] = clock64() - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

// four warps, each with its own accumulator and mbarrier, issue 32 MMAs at the same time: is the ~61-cycle issue interval per
// thread or per SM?
template <int N>
__global__ void __launch_bounds__(128, 1) k4(long long* out, int n_issuers) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) mbar_init(bar + w, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    for (int rep = 0; rep < 2; ++rep) {
        __syncthreads();
        if (lane == 0 && warp < n_issuers) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
            const uint32_t sa = smem_u32(smem) + warp * 16384, sb = smem_u32(smem + 65536);
            const long long t0 = clock64();
#pragma unroll
            for (int it = 0; it < 32; ++it)
                umma_bf16(tm + warp * N, umma_smem_desc_sw128(sa + (it & 3) * 32, 0, 1024), umma_smem_desc_sw128(sb + (it & 3) * 32, 0, 1024),
                          idesc, it ? 1u : 0u);
            const long long t1 = clock64();
            umma_commit(bar + warp);
            mbar_wait(bar + warp, rep);
            const long long t2 = clock64();
            if (rep == 1) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N>
void run4(long long* d, int n) {
    cudaMemset(d, 0, 128);
    cudaFuncSetAttribute(k4<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 1024);
    k4<N><<<1, 128, 100 * 1024 + 1024>>>(d, n);
    long long h[8];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    printf("N=%3d %d issuing warps x 32 MMAs: issue / done per warp:", N, n);
    for (int w = 0; w < n; ++w) printf("  %lld / %lld", h[2 * w], h[2 * w + 1]);
    printf("\n");
}

template <int N>
void run(long long* d) {
    cudaMemset(d, 0, 128);
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 1024);
    k<N><<<1, 128, 100 * 1024 + 1024>>>(d);
    long long h[10];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    printf("N=%3d 32 MMAs: 1 acc issue %5lld done %5lld (%.0f cyc/MMA) | 2 acc %5lld / %5lld (%.0f) | 4 acc %5lld / %5lld (%.0f) | "
           "descs precomputed %5lld / %5lld (%.0f) | 1 MMA round trip %lld | 4 dependent + commit %lld\n",
           N, h[0], h[1], h[1] / 32.0, h[2], h[3], h[3] / 32.0, h[4], h[5], h[5] / 32.0, h[6], h[7], h[7] / 32.0, h[8], h[9]);
}

int main() {
    long long* d;
    cudaMalloc(&d, 128);
    run<32>(d); run<64>(d); run<128>(d); run<256>(d);
    run4<32>(d, 1); run4<32>(d, 2); run4<32>(d, 4); run4<64>(d, 1); run4<64>(d, 2); run4<64>(d, 4); run4<128>(d, 2); run4<128>(d, 4);
    return 0;
}
