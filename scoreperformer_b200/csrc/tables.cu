// Computed per-field embedding tables: W_f = index rows {discrete ids} + MLP(token_values) with those rows zeroed.
//
// Reference: modules/transformer/embeddings.py:124-143 (token_weight / value_weight) and :199-211 (value MLP =
// Linear(1,E) -> Mish -> Linear(E,E)); the reference rebuilds every table with ~10 tiny ATen ops per field per use.
// Here ONE launch builds all fields' tables and ONE launch back-propagates a table gradient into every field's
// (index_weight, W0, b0, W1, b1) gradients.  E = 128.
#include "common.cuh"

namespace {

constexpr int E = 128;
constexpr int MAXF = 16;

struct TableFields {
    int n_fields;
    int size[MAXF], offset[MAXF];
    const float* index_w[MAXF];     // [V, E]
    const float* values[MAXF];      // [V] token values
    const float* disc[MAXF];        // [V] 1.0 where the row is a discrete id (index row kept, MLP row zeroed)
    const float* w0[MAXF];          // [E] (Linear(1,E).weight[:,0])
    const float* b0[MAXF];          // [E]
    const float* w1[MAXF];          // [E, E]
    const float* b1[MAXF];          // [E]
    float* d_index_w[MAXF];
    float* d_w0[MAXF];
    float* d_b0[MAXF];
    float* d_w1[MAXF];
    float* d_b1[MAXF];
};

__device__ __forceinline__ float mish(float x) {
    const float sp = x > 20.f ? x : log1pf(__expf(x));
    return x * tanhf(sp);
}
__device__ __forceinline__ float mish_grad(float x) {
    const float sp = x > 20.f ? x : log1pf(__expf(x));
    const float t = tanhf(sp);
    const float sig = 1.f / (1.f + __expf(-x));
    return t + x * (1.f - t * t) * sig;
}

// grid (row blocks, F), 128 threads = the E output columns; 8 rows per CTA iteration.  W1 is staged once per CTA in shared
// memory with a padded row stride (thread c walks row c: a stride of E floats would be a 32-way bank conflict, and reading it
// straight from global memory touches 32 different lines per warp load -- that was 65 us per launch).
constexpr int FWD_W1_STRIDE = E + 1;
__global__ void __launch_bounds__(E)
table_fwd_kernel(TableFields tf, float* __restrict__ table) {
    extern __shared__ float s_w1p[];                // [E][E + 1]
    __shared__ float sh[8][E];
    const int f = blockIdx.y, c = threadIdx.x;
    const int V = tf.size[f];
    if ((int)blockIdx.x * 8 >= V) return;
    const float w0 = tf.w0[f][c], b0 = tf.b0[f][c], b1 = tf.b1[f][c];
    for (int i = c; i < E * E; i += E) s_w1p[(i / E) * FWD_W1_STRIDE + (i % E)] = tf.w1[f][i];      // coalesced read
    const float* __restrict__ w1 = s_w1p + c * FWD_W1_STRIDE;
    for (int r0 = blockIdx.x * 8; r0 < V; r0 += gridDim.x * 8) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r0 + k;
            sh[k][c] = r < V ? mish(fmaf(w0, tf.values[f][r], b0)) : 0.f;
        }
        __syncthreads();
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = b1;
        for (int j = 0; j < E; ++j) {
            const float w = w1[j];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, sh[k][j], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r0 + k;
            if (r < V) {
                const bool d = tf.disc[f][r] != 0.f;
                table[(size_t)(tf.offset[f] + r) * E + c] = d ? tf.index_w[f][(size_t)r * E + c] : acc[k];
            }
        }
    }
}

// grid (row splits, F), 256 threads.  W1 sits in shared memory; dW1 [E,E] is accumulated in registers (64 entries per thread)
// over the CTA's rows of the field and flushed with one atomic per entry.
constexpr int BWD_SPLITS = 8;
__global__ void __launch_bounds__(256)
table_bwd_kernel(TableFields tf, const float* __restrict__ dtable) {
    extern __shared__ float s_w1[];                 // [E][E]
    __shared__ float s_dv[E], s_h[E], s_pre[E];
    const int f = blockIdx.y, t = threadIdx.x;
    const int V = tf.size[f];
    for (int i = t; i < E * E; i += 256) s_w1[i] = tf.w1[f][i];
    float dw1[64];                    // thread t owns outputs (i, j) with i = t/2, j in [(t%2)*64, +64)
#pragma unroll
    for (int k = 0; k < 64; ++k) dw1[k] = 0.f;
    float db1 = 0.f, dw0 = 0.f, db0 = 0.f;   // for threads < E (column t)
    const int oi = t >> 1, oj0 = (t & 1) * 64;
    const float w0 = t < E ? tf.w0[f][t] : 0.f, b0 = t < E ? tf.b0[f][t] : 0.f;
    for (int r = blockIdx.x; r < V; r += BWD_SPLITS) {
        const bool d = tf.disc[f][r] != 0.f;
        const float* g = dtable + (size_t)(tf.offset[f] + r) * E;
        if (d) {                                             // index row: gradient flows to index_weight only
            if (t < E) atomicAdd(tf.d_index_w[f] + (size_t)r * E + t, g[t]);
            continue;
        }
        __syncthreads();
        if (t < E) {
            const float pre = fmaf(w0, tf.values[f][r], b0);
            s_pre[t] = pre;
            s_h[t] = mish(pre);
            s_dv[t] = g[t];
            db1 += g[t];
        }
        __syncthreads();
        {   // dW1[i, j] += dv[i] * h[j]
            const float dv = s_dv[oi];
#pragma unroll
            for (int k = 0; k < 64; ++k) dw1[k] = fmaf(dv, s_h[oj0 + k], dw1[k]);
        }
        if (t < E) {                                         // dh[t] = sum_i W1[i, t] * dv[i]
            float acc = 0.f;
#pragma unroll 16
            for (int i = 0; i < E; ++i) acc = fmaf(s_w1[i * E + t], s_dv[i], acc);
            const float dpre = acc * mish_grad(s_pre[t]);
            dw0 = fmaf(dpre, tf.values[f][r], dw0);
            db0 += dpre;
        }
    }
#pragma unroll
    for (int k = 0; k < 64; ++k)
        if (dw1[k] != 0.f) atomicAdd(tf.d_w1[f] + (size_t)oi * E + oj0 + k, dw1[k]);
    if (t < E) {
        atomicAdd(tf.d_b1[f] + t, db1);
        atomicAdd(tf.d_w0[f] + t, dw0);
        atomicAdd(tf.d_b0[f] + t, db0);
    }
}

int fill(TableFields& tf, const int* sizes, int n_fields, const void* const* ptrs, int per_field, bool backward) {
    SPB_CHECK_ARG(n_fields > 0 && n_fields <= MAXF, "tables: 1..%d fields supported", MAXF);
    SPB_CHECK_ARG(ptrs != nullptr, "tables: null pointer table");
    tf.n_fields = n_fields;
    int off = 0;
    for (int f = 0; f < n_fields; ++f) {
        tf.size[f] = sizes[f];
        tf.offset[f] = off;
        off += sizes[f];
        const void* const* p = ptrs + (size_t)f * per_field;
        tf.index_w[f] = (const float*)p[0]; tf.values[f] = (const float*)p[1]; tf.disc[f] = (const float*)p[2];
        tf.w0[f] = (const float*)p[3]; tf.b0[f] = (const float*)p[4]; tf.w1[f] = (const float*)p[5]; tf.b1[f] = (const float*)p[6];
        if (backward) {
            tf.d_index_w[f] = (float*)p[7]; tf.d_w0[f] = (float*)p[8]; tf.d_b0[f] = (float*)p[9];
            tf.d_w1[f] = (float*)p[10]; tf.d_b1[f] = (float*)p[11];
        }
        for (int k = 0; k < per_field; ++k) SPB_CHECK_ARG(p[k] != nullptr, "tables: null pointer for field %d slot %d", f, k);
    }
    return SPB_OK;
}

}  // namespace

// ptrs: HOST array of n_fields * 7 device pointers per field: index_weight [V,128], token_values [V], discrete mask [V] (fp32 0/1),
// W0 [128], b0 [128], W1 [128,128], b1 [128].  table fp32 [sum V, 128] is written.
extern "C" int spb_table_build_fwd(const int* field_sizes, int n_fields, const void* const* ptrs, float* table, cudaStream_t stream) {
    TableFields tf;
    int rc = fill(tf, field_sizes, n_fields, ptrs, 7, false);
    if (rc != SPB_OK) return rc;
    SPB_CHECK_ARG(table != nullptr, "spb_table_build_fwd: null output");
    constexpr int FWD_SMEM = E * FWD_W1_STRIDE * 4;
    static bool fwd_configured = false;
    if (!fwd_configured) {
        SPB_CHECK_CUDA(cudaFuncSetAttribute(table_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
        fwd_configured = true;
    }
    table_fwd_kernel<<<dim3(8, n_fields), E, FWD_SMEM, stream>>>(tf, table);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ptrs: n_fields * 12 device pointers per field: the 7 of the forward + d_index_weight, dW0, db0, dW1, db1 (all ACCUMULATED into).
extern "C" int spb_table_build_bwd(const int* field_sizes, int n_fields, const void* const* ptrs, const float* dtable, cudaStream_t stream) {
    TableFields tf;
    int rc = fill(tf, field_sizes, n_fields, ptrs, 12, true);
    if (rc != SPB_OK) return rc;
    SPB_CHECK_ARG(dtable != nullptr, "spb_table_build_bwd: null gradient");
    static bool configured = false;
    if (!configured) {
        SPB_CHECK_CUDA(cudaFuncSetAttribute(table_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, E * E * 4));
        configured = true;
    }
    table_bwd_kernel<<<dim3(BWD_SPLITS, n_fields), 256, E * E * 4, stream>>>(tf, dtable);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
