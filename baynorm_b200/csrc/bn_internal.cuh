// bn_internal.cuh -- shared plumbing of libbaynorm_b200 (context, staging buffers,
// launch accounting, block/warp reductions).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string>
#include <vector>

#include "../../include/baynorm_b200.h"

// ---------------------------------------------------------------------------------
// opaque handles
// ---------------------------------------------------------------------------------
struct bn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    int64_t launches = 0;
    bn_allreduce_fn ar = nullptr;
    void *ar_user = nullptr;
    int rank = 0, world = 1;
    cudaEvent_t ev[8] = {};
    double *pinned = nullptr; // 64 doubles of page-locked host memory for truly asynchronous scalar read-backs
    // grow-only device scratch kept across calls (slots 0-1, 8-9: the 3-D sampler's lists and side buffers; 2-4: counters of the
    // gene-major build; 5: DownSampling's list; 6-7: result tiles): re-allocating GB-sized buffers per call through the stream-ordered
    // pool was measured to stall the host for 5-400 ms every few calls
    void *scratch[28] = {};        // slots 10-14: the second set of the 3-D sampler's per-chunk buffers
    size_t scratch_bytes[28] = {};
    // side streams of the 3-D sampler (bn_sample.cu): [0] prepares chunk c + 1 (classification, side-buffer draws) while the
    // main kernel of chunk c runs on `stream`; [1..4] carry the table classes that run next to each other
    cudaStream_t aux[5] = {};
    cudaEvent_t ev_aux[12] = {};
    bool sample_attr_set = false;  // function attributes are per device: set once per context
    bool fit2_attr_set = false;
    // result streaming (bn_stream.cu): a second stream for device -> host copies, two result tiles in flight
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_comp[2] = {}, ev_copy[2] = {};
    double *host_tile[2] = {};      // pinned, grow-only
    size_t host_tile_bytes = 0;
    int tune[BN_TUNE_COUNT] = {};   // bn_debug_tune: A/B switches of profiling runs (0 = default)
    int64_t last_fit[4] = {};       // bn_debug_last_fit
    unsigned int *fit_parked = nullptr; // pinned: genes parked by the last sweep kernel
};
// returns ctx->scratch[slot] with at least `bytes` bytes (contents undefined)
int bn_ctx_scratch(bn_ctx *ctx, int slot, size_t bytes, void **out);

struct bn_mat {
    bn_ctx *ctx = nullptr;
    int64_t G = 0, C = 0, nnz = 0;
    int64_t gene0 = 0, cell0 = 0;
    // CSC, device-owned
    int64_t *colptr = nullptr; // [C+1]
    int32_t *rowidx = nullptr; // [nnz]
    double *val = nullptr;     // [nnz]
    bool is_count = false;     // all values are non-negative integers
    // gene-major twin (built on demand by bn_mat_ensure_csr)
    bool has_csr = false;
    int64_t *rowptr = nullptr; // [G+1]
    // count matrices (is_count): one packed word per non-zero, (cell << 32) | count, cells
    // increasing within a gene -- 8 B per non-zero for every per-gene stage
    uint64_t *gm = nullptr;    // [nnz]
    // general real-valued matrices (EstPrior on already normalised data): separate arrays
    int32_t *cell = nullptr;   // [nnz]
    double *rval = nullptr;    // [nnz]
    // blocked column pointers (built on demand by bn_mat_ensure_blockptr): blockptr[t * C + j] = offset, from the
    // start colptr[j] of column j, of its first entry whose row is >= t * BN_ROWBLOCK, t = 0 .. nblock (row nblock =
    // the column length).  The posterior kernels give every warp one row block: its slice of a column is
    // colptr[j] + [blockptr[t][j], blockptr[t + 1][j]).  Offsets, not positions: a shard may hold >= 2^32 non-zeros.
    uint32_t *blockptr = nullptr; // [(nblock + 1) * C]
    bool has_blockptr = false;
    int64_t nblock = 0;
};
#define BN_ROWBLOCK 128

// ---------------------------------------------------------------------------------
// error helpers
// ---------------------------------------------------------------------------------
int bn_fail(bn_ctx *ctx, int code, const char *fmt, ...);

#define BN_CUDA(ctx, expr)                                                                      \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return bn_fail((ctx), e__ == cudaErrorMemoryAllocation ? BN_ERR_OOM : BN_ERR_CUDA,  \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define BN_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != BN_OK) return rc__; \
    } while (0)

// Launch + account + check.  Every kernel of the library goes through this so that
// bn_ctx_launch_count() is the number bench.py reports as gpu_launches.
#define BN_LAUNCH(ctx, kernel, grid, block, smem, ...)                                          \
    do {                                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return bn_fail((ctx), BN_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,      \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                        \
    } while (0)

// the same on another stream of the context (the side streams of the 3-D sampler)
#define BN_LAUNCH_ON(ctx, st, kernel, grid, block, smem, ...)                                   \
    do {                                                                                        \
        kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                 \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return bn_fail((ctx), BN_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,      \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                        \
    } while (0)

// ---------------------------------------------------------------------------------
// pointer classification + staging
// ---------------------------------------------------------------------------------
inline bool bn_is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Stream-ordered device buffer (cudaMallocAsync pool of the context's device).
template <typename T> struct DevBuf {
    bn_ctx *ctx = nullptr;
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(bn_ctx *c, size_t count)
    {
        release();
        ctx = c;
        n = count;
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), c->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            return bn_fail(c, BN_ERR_OOM, "device allocation of %zu bytes failed: %s", count * sizeof(T),
                           cudaGetErrorString(e));
        }
        return BN_OK;
    }
    void release()
    {
        if (p) cudaFreeAsync(p, ctx->stream);
        p = nullptr;
    }
    T *take()
    {
        T *q = p;
        p = nullptr;
        return q;
    }
};

// Read-only argument that may live on host or device.
template <typename T> struct DevIn {
    DevBuf<T> tmp;
    const T *p = nullptr;
    int bind(bn_ctx *ctx, const T *src, size_t count)
    {
        if (!src || bn_is_device_ptr(src)) {
            p = src;
            return BN_OK;
        }
        BN_TRY(tmp.alloc(ctx, count));
        BN_CUDA(ctx, cudaMemcpyAsync(tmp.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        p = tmp.p;
        return BN_OK;
    }
};

// Output argument that may live on host or device (NULL = not wanted unless force).
template <typename T> struct DevOut {
    DevBuf<T> tmp;
    T *p = nullptr;    // device pointer kernels write to
    T *host = nullptr; // where to copy afterwards (NULL: p is the caller's device memory or scratch)
    size_t n = 0;
    int bind(bn_ctx *ctx, T *dst, size_t count, bool force = false)
    {
        n = count;
        if (dst && bn_is_device_ptr(dst)) {
            p = dst;
            return BN_OK;
        }
        if (!dst && !force) {
            p = nullptr;
            return BN_OK;
        }
        BN_TRY(tmp.alloc(ctx, count));
        p = tmp.p;
        host = dst;
        return BN_OK;
    }
    int commit(bn_ctx *ctx)
    {
        if (host && n) BN_CUDA(ctx, cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        return BN_OK;
    }
};

int bn_sync(bn_ctx *ctx);
// posterior kernels on device pointers, queued on the context's stream without a final synchronisation
// (kind: 0 mean, 1 mode, 2 S draws); out is G x ncol (x S, planes G*ncol apart)
int bn_post_device(bn_ctx *ctx, int kind, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                   int64_t col0, int64_t ncol, int S, uint64_t seed, double *d_out);
// the same for a HOST destination: column tiles, tile t+1 computed while tile t travels (bn_stream.cu)
int bn_post_to_host(bn_ctx *ctx, int kind, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                    int64_t col0, int64_t ncol, int S, uint64_t seed, double *h_out);
int bn_mat_ensure_csr(bn_ctx *ctx, bn_mat *m);
int bn_mat_ensure_blockptr(bn_ctx *ctx, const bn_mat *m);
// largest status code over all gene shards (bn_ctx.cu): called before the first collective of an entry point
int bn_agree(bn_ctx *ctx, int local_rc);
// in-place all-reduce over shards through the user hook (no-op when world == 1)
int bn_allreduce(bn_ctx *ctx, double *dev, int64_t count, int op);

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double bn_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v; // butterfly: every lane ends with the same bits (fp add is commutative)
}
__device__ __forceinline__ double bn_warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double bn_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over a group of TPG threads (TPG == 32: one warp; otherwise the whole CTA, which
// must have exactly TPG threads).  scratch: >= TPG/32 doubles of shared memory.  Every
// thread of the group receives the same bits; order of additions is fixed.
template <int TPG> __device__ __forceinline__ double bn_group_sum(double v, double *scratch)
{
    v = bn_warp_sum(v);
    if (TPG == 32) return v;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads(); // scratch may still be read by the previous call
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < TPG / 32; k++) s += scratch[k];
    return s;
}
#endif
