// Error plumbing + small utility kernels shared by the C-ABI (include/spb200.h).
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

static thread_local char g_last_error[512] = "";

void spb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* spb_last_error() { return g_last_error; }

extern "C" int spb_abi_version() { return 1; }

namespace {

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n,
                                        const uint8_t* __restrict__ rowmask, int row_len) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i + 3 < n; i += stride) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        if (rowmask != nullptr && !rowmask[i / row_len]) v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i) = o;
    }
    if (i < n && i + 3 >= n) {
        for (int64_t j = i; j < n; ++j) {
            float v = src[j];
            if (rowmask != nullptr && !rowmask[j / row_len]) v = 0.f;
            dst[j] = __float2bfloat16_rn(v);
        }
    }
}

}  // namespace

// dst[i] = bf16(src[i]) (optionally zeroing whole rows of length row_len where rowmask is false;
// row_len must be a multiple of 4 in that case).
extern "C" int spb_cast_f32_bf16(const float* src, void* dst, int64_t n, const uint8_t* rowmask, int row_len,
                                 cudaStream_t stream) {
    if (n <= 0) return SPB_OK;
    SPB_CHECK_ARG(src && dst, "spb_cast_f32_bf16: null pointer");
    SPB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                  "spb_cast_f32_bf16: unaligned pointers");
    SPB_CHECK_ARG(rowmask == nullptr || (row_len > 0 && row_len % 4 == 0), "spb_cast_f32_bf16: row_len must be a multiple of 4");
    const int threads = 256;
    int64_t blocks = (n / 4 + threads - 1) / threads;
    int max_blocks = spb_num_sms() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    cast_f32_to_bf16_kernel<<<(int)blocks, threads, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n, rowmask,
                                                                 row_len > 0 ? row_len : 1);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

namespace {

// out[c] += sum_r x[r, c];  thread owns 2 adjacent columns, 8 row-lanes per block, rows split over blockIdx.y
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int ld, float* __restrict__ out, int n_rows, int n_cols, int rows_per_block) {
    __shared__ float red[8][64];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 2;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
    float s0 = 0.f, s1 = 0.f;
    if (c < n_cols) {
        const bool pair = (c + 1 < n_cols);
        for (int r = r0 + ty; r < r1; r += 8) {
            s0 += (float)x[(size_t)r * ld + c];
            if (pair) s1 += (float)x[(size_t)r * ld + c + 1];
        }
    }
    red[ty][tx * 2] = s0;
    red[ty][tx * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        const int cc = blockIdx.x * 64 + threadIdx.x;
        if (cc < n_cols && s != 0.f) atomicAdd(out + cc, s);
    }
}

// Vector form: a lane owns 16 bytes of every row (8 bf16 / 4 fp32 columns), a warp one 512-byte row segment, the 8 warps of a
// block interleave rows (4 rows in flight per thread), partial sums meet in shared memory and leave as one atomic per column.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, int ld, float* __restrict__ out, int n_rows, int n_cols, int rows_per_block) {
    constexpr int VEC = 16 / (int)sizeof(T);
    __shared__ float red[8][32 * VEC + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * VEC;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (c < n_cols) {      // n_cols % VEC == 0 is guaranteed by the host
        auto add = [&](const uint4& u) {
            if (sizeof(T) == 2) {
                const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
                acc[0] += a.x; acc[1] += a.y; acc[2 % VEC] += b.x; acc[3 % VEC] += b.y;
                acc[4 % VEC] += cc.x; acc[5 % VEC] += cc.y; acc[6 % VEC] += d.x; acc[7 % VEC] += d.y;
            } else {
                acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y); acc[2 % VEC] += __uint_as_float(u.z);
                acc[3 % VEC] += __uint_as_float(u.w);
            }
        };
        int r = r0 + w;
        for (; r + 24 < r1; r += 32) {
            const uint4 u0 = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + c);
            const uint4 u1 = *reinterpret_cast<const uint4*>(x + (size_t)(r + 8) * ld + c);
            const uint4 u2 = *reinterpret_cast<const uint4*>(x + (size_t)(r + 16) * ld + c);
            const uint4 u3 = *reinterpret_cast<const uint4*>(x + (size_t)(r + 24) * ld + c);
            add(u0); add(u1); add(u2); add(u3);
        }
        for (; r < r1; r += 8) add(*reinterpret_cast<const uint4*>(x + (size_t)r * ld + c));
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) red[w][lane * VEC + j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * VEC; i += 256) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][i];
        const int cc = blockIdx.x * 32 * VEC + i;
        if (cc < n_cols && sum != 0.f) atomicAdd(out + cc, sum);
    }
}

}  // namespace

// out fp32 [n_cols] += column sums of x [n_rows, ld] (bias gradients).  x_fp32 selects the input dtype.
extern "C" int spb_colsum(const void* x, int x_fp32, int ld, float* out, int n_rows, int n_cols, cudaStream_t stream) {
    if (n_rows <= 0 || n_cols <= 0) return SPB_OK;
    SPB_CHECK_ARG(x && out, "spb_colsum: null pointer");
    const int esize = x_fp32 ? 4 : 2, vec = 16 / esize;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((size_t)ld * esize) % 16 == 0 && n_cols % vec == 0) {
        const int col_blocks = ceil_div(n_cols, 32 * vec);
        int chunks = ceil_div(4 * spb_num_sms(), col_blocks);
        int rows_per_block = ceil_div(n_rows, chunks);
        if (rows_per_block < 64) rows_per_block = 64;
        chunks = ceil_div(n_rows, rows_per_block);
        dim3 grid(col_blocks, chunks);
        if (x_fp32) colsum_vec_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), ld, out, n_rows, n_cols, rows_per_block);
        else colsum_vec_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, out, n_rows, n_cols, rows_per_block);
        SPB_CHECK_LAUNCH();
        return SPB_OK;
    }
    int chunks = ceil_div(2 * spb_num_sms(), ceil_div(n_cols, 64));
    int rows_per_block = ceil_div(n_rows, chunks);
    if (rows_per_block < 64) rows_per_block = 64;
    chunks = ceil_div(n_rows, rows_per_block);
    dim3 grid(ceil_div(n_cols, 64), chunks);
    if (x_fp32) colsum_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), ld, out, n_rows, n_cols, rows_per_block);
    else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, out, n_rows, n_cols, rows_per_block);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
