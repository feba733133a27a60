// Standalone check of spb_gemm_bf16 (no torch): all four operand-major combinations, ragged sizes,
// epilogues and split-K against a CPU double reference, then a few C2-shaped timings.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

extern "C" int spb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int trans_a, int trans_b, int lda,
                             int ldb, int ldc, const float* bias, const float* residual, int ldr, const uint8_t* rowmask,
                             int c_fp32, int split_k, int accumulate, const float* alpha, cudaStream_t stream);
extern "C" const char* spb_last_error();

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static int run_case(int M, int N, int K, int ta, int tb, int epi, int c_fp32, int split) {
    // logical A[M,K], B[N,K]
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), res((size_t)M * N);
    std::vector<uint8_t> mask(M);
    for (auto& v : A) v = bf(frand());
    for (auto& v : B) v = bf(frand());
    for (auto& v : bias) v = frand();
    for (auto& v : res) v = frand();
    for (auto& v : mask) v = rand() % 4 != 0;
    int lda = ta ? M : K, ldb = tb ? N : K;
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) hA[ta ? (size_t)k * M + m : (size_t)m * K + k] = __float2bfloat16_rn(A[(size_t)m * K + k]);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) hB[tb ? (size_t)k * N + n : (size_t)n * K + k] = __float2bfloat16_rn(B[(size_t)n * K + k]);
    void *dA, *dB, *dC;
    float *dbias, *dres;
    uint8_t* dmask;
    size_t csz = (size_t)M * N * (c_fp32 ? 4 : 2);
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, csz);
    cudaMalloc(&dbias, N * 4); cudaMalloc(&dres, (size_t)M * N * 4); cudaMalloc(&dmask, M);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dres, res.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dmask, mask.data(), M, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, csz);
    int rc = spb_gemm_bf16(dA, dB, dC, M, N, K, ta, tb, lda, ldb, N, (epi & 1) ? dbias : nullptr, (epi & 2) ? dres : nullptr, N,
                           (epi & 4) ? dmask : nullptr, c_fp32, split, 0, nullptr, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) {
        printf("FAIL launch M=%d N=%d K=%d ta=%d tb=%d: rc=%d %s / %s\n", M, N, K, ta, tb, rc, spb_last_error(), cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> out((size_t)M * N);
    if (c_fp32) cudaMemcpy(out.data(), dC, csz, cudaMemcpyDeviceToHost);
    else {
        std::vector<__nv_bfloat16> ob((size_t)M * N);
        cudaMemcpy(ob.data(), dC, csz, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < ob.size(); ++i) out[i] = __bfloat162float(ob[i]);
    }
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
            if (epi & 1) acc += bias[n];
            if ((epi & 4) && !mask[m]) acc = 0;
            if (epi & 2) acc += res[(size_t)m * N + n];
            double err = fabs(acc - out[(size_t)m * N + n]);
            if (err > max_err) max_err = err;
            if (fabs(acc) > max_ref) max_ref = fabs(acc);
        }
    double tol = (c_fp32 ? 2e-3 : 1.2e-2) * (max_ref > 1 ? max_ref : 1);
    int bad = !(max_err <= tol);
    printf("%s M=%d N=%d K=%d ta=%d tb=%d epi=%d fp32=%d split=%d max_err=%.3e max_ref=%.2f\n", bad ? "FAIL" : "ok  ", M, N, K,
           ta, tb, epi, c_fp32, split, max_err, max_ref);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias); cudaFree(dres); cudaFree(dmask);
    return bad;
}

static void bench(int M, int N, int K, int ta, int tb, int c_fp32, int split, const char* what) {
    void *dA, *dB, *dC;
    cudaMalloc(&dA, (size_t)M * K * 2); cudaMalloc(&dB, (size_t)N * K * 2); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMemset(dA, 0, (size_t)M * K * 2); cudaMemset(dB, 0, (size_t)N * K * 2);
    int lda = ta ? M : K, ldb = tb ? N : K;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) spb_gemm_bf16(dA, dB, dC, M, N, K, ta, tb, lda, ldb, N, 0, 0, 0, 0, c_fp32, split, 0, nullptr, 0);
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; ++i) spb_gemm_bf16(dA, dB, dC, M, N, K, ta, tb, lda, ldb, N, 0, 0, 0, 0, c_fp32, split, 0, nullptr, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    double flops = 2.0 * M * N * K;
    double bytes = ((double)M * K + (double)N * K) * 2 + (double)M * N * (c_fp32 ? 4 : 2);
    printf("bench %-22s M=%d N=%d K=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s (err=%s)\n", what, M, N, K, ms * 1e3, flops / ms * 1e-9,
           bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

int main() {
    int fails = 0;
    srand(1);
    for (int ta = 0; ta < 2; ++ta)
        for (int tb = 0; tb < 2; ++tb) {
            fails += run_case(128, 128, 64, ta, tb, 0, 1, 1);
            fails += run_case(256, 64, 128, ta, tb, 0, 1, 1);
            fails += run_case(304, 200, 256, ta, tb, 0, 1, 1);
            fails += run_case(304, 200, 200, ta, tb, 7, 0, 1);
        }
    fails += run_case(1000, 384, 1024, 0, 0, 3, 0, 1);
    fails += run_case(511, 256, 256, 0, 0, 6, 1, 1);
    fails += run_case(256, 256, 8192, 1, 1, 0, 1, 0);
    fails += run_case(384, 256, 4088, 1, 1, 0, 1, 5);
    fails += run_case(64, 40, 2048, 1, 1, 0, 1, 0);
    printf("%s (%d failures)\n", fails ? "GEMM TEST FAILED" : "GEMM TEST PASSED", fails);
    if (!fails) {
        bench(32768, 384, 256, 0, 0, 0, 1, "qkv fwd");
        bench(32768, 256, 256, 0, 0, 1, 1, "out fwd fp32");
        bench(32768, 2048, 256, 0, 0, 0, 1, "ffn1 fwd");
        bench(32768, 256, 1024, 0, 0, 1, 1, "ffn2 fwd fp32");
        bench(32768, 256, 1536, 0, 0, 0, 1, "emb proj");
        bench(32768, 256, 2048, 0, 1, 0, 1, "ffn1 dgrad");
        bench(2048, 256, 32768, 1, 1, 1, 0, "ffn1 wgrad splitK");
        bench(256, 256, 32768, 1, 1, 1, 0, "out wgrad splitK");
        bench(8192, 8192, 8192, 0, 0, 0, 1, "square 8k");
    }
    return fails != 0;
}
