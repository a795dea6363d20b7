// bn_ctx.cu -- context lifetime, error reporting, matrix upload and the gene-major
// (CSR) twin the per-gene kernels read.  Replaces the per-call `as<arma::sp_mat>(Data)`
// copies of the reference's wrappers (src/RcppExports.cpp:76-90): the matrix is placed
// in HBM once and every later stage reads it there.
#include <cub/cub.cuh>
#include <stdarg.h>

#include "bn_internal.cuh"

static thread_local std::string g_create_err;

int bn_fail(bn_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_err = buf;
    return code;
}

extern "C" {

int bn_version(void) { return BN_VERSION; }

const char *bn_last_error(const bn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bn_ctx_create(int device, bn_ctx **out)
{
    if (!out) return bn_fail(nullptr, BN_ERR_INVALID, "bn_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return bn_fail(nullptr, BN_ERR_NO_DEVICE,
                       "no CUDA device visible (%s); libbaynorm_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return bn_fail(nullptr, BN_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    bn_ctx *c = new bn_ctx();
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess) {
        delete c;
        return bn_fail(nullptr, BN_ERR_CUDA, "cudaSetDevice/GetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (c->prop.major != 10) {
        int maj = c->prop.major, mnr = c->prop.minor;
        delete c;
        return bn_fail(nullptr, BN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       maj, mnr);
    }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return bn_fail(nullptr, BN_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    cudaHostAlloc((void **)&c->pinned, 64 * sizeof(double), cudaHostAllocDefault);
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return BN_OK;
}

void bn_ctx_destroy(bn_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (void *p : ctx->scratch)
        if (p) cudaFree(p);
    delete ctx;
}

int bn_ctx_synchronize(bn_ctx *ctx) { return bn_sync(ctx); }
void *bn_ctx_stream(bn_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t bn_ctx_launch_count(const bn_ctx *ctx) { return ctx ? ctx->launches : 0; }

int bn_ctx_device_info(bn_ctx *ctx, int64_t out[4])
{
    if (!ctx || !out) return BN_ERR_INVALID;
    out[0] = ctx->prop.multiProcessorCount;
    out[1] = ctx->prop.clockRate;
    out[2] = ctx->prop.major * 10 + ctx->prop.minor;
    out[3] = (int64_t)ctx->prop.totalGlobalMem;
    return BN_OK;
}

int bn_ctx_set_allreduce(bn_ctx *ctx, bn_allreduce_fn fn, void *user, int rank, int world)
{
    if (!ctx) return BN_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return bn_fail(ctx, BN_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    if (world > 1 && !fn) return bn_fail(ctx, BN_ERR_INVALID, "world > 1 needs an all-reduce hook");
    ctx->ar = fn;
    ctx->ar_user = user;
    ctx->rank = rank;
    ctx->world = world;
    return BN_OK;
}

} // extern "C"

int bn_ctx_scratch(bn_ctx *ctx, int slot, size_t bytes, void **out)
{
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            BN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            BN_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        const size_t want = bytes + bytes / 8; // headroom: column blocks of slightly different width re-use it
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) return bn_fail(ctx, BN_ERR_OOM, "scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return BN_OK;
}

int bn_sync(bn_ctx *ctx)
{
    BN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BN_OK;
}

int bn_allreduce(bn_ctx *ctx, double *dev, int64_t count, int op)
{
    if (ctx->world <= 1 || count <= 0) return BN_OK;
    int rc = ctx->ar(ctx->ar_user, dev, count, 0, op, (void *)ctx->stream);
    if (rc != 0) return bn_fail(ctx, BN_ERR_COLLECTIVE, "all-reduce hook returned %d", rc);
    return BN_OK;
}

// ---------------------------------------------------------------------------------
// upload + validation
// ---------------------------------------------------------------------------------
__global__ void k_widen_colptr(const int32_t *__restrict__ p32, int64_t *__restrict__ p64, int64_t n)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) p64[k] = p32[k];
}

// flags[0] |= 1: colptr not monotone / out of range; |= 2: row index out of range or not
// strictly increasing within a column; |= 4: some value is not a non-negative integer.
__global__ void k_validate_csc(int64_t G, int64_t C, int64_t nnz, const int64_t *__restrict__ colptr,
                               const int32_t *__restrict__ rowidx, const double *__restrict__ val, int *flags)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    int f = 0;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        if (a > b || a < 0 || b > nnz) {
            f |= 1;
            continue;
        }
        for (int64_t k = a + lane; k < b; k += 32) {
            const int32_t r = rowidx[k];
            if (r < 0 || r >= G) f |= 2;
            if (k > a && rowidx[k - 1] >= r) f |= 2;
            const double v = val[k];
            if (!(v >= 0.0) || v != floor(v) || v > 4294967295.0) f |= 4;
        }
    }
    if (warp == 0 && lane == 0 && (colptr[0] != 0 || colptr[C] != nnz)) f |= 1;
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(flags, f);
}

static int bn_mat_finish(bn_ctx *ctx, bn_mat *m)
{
    DevBuf<int> flags;
    BN_TRY(flags.alloc(ctx, 1));
    BN_CUDA(ctx, cudaMemsetAsync(flags.p, 0, sizeof(int), ctx->stream));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8 > 0 ? (m->C + 7) / 8 : 1, 148 * 16);
    BN_LAUNCH(ctx, k_validate_csc, blocks, 256, 0, m->G, m->C, m->nnz, m->colptr, m->rowidx, m->val, flags.p);
    int h = 0;
    BN_CUDA(ctx, cudaMemcpyAsync(&h, flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(bn_sync(ctx));
    if (h & 1) return bn_fail(ctx, BN_ERR_INVALID, "CSC column pointers are not a monotone [0..nnz] sequence");
    if (h & 2) return bn_fail(ctx, BN_ERR_INVALID, "CSC row indices out of range or not strictly increasing within a column");
    m->is_count = !(h & 4);
    return BN_OK;
}

extern "C" int bn_mat_from_csc(bn_ctx *ctx, int64_t G, int64_t C, const void *p, int p_is_int64, const int32_t *i,
                               const double *x, bn_mat **out)
{
    if (!ctx || !out) return BN_ERR_INVALID;
    *out = nullptr;
    if (G <= 0 || C <= 0 || !p) return bn_fail(ctx, BN_ERR_INVALID, "bn_mat_from_csc: need G > 0, C > 0 and p");
    if (G > 2147483647LL) return bn_fail(ctx, BN_ERR_INVALID, "G exceeds int32 row indices");
    cudaSetDevice(ctx->device);
    // nnz = p[C]
    int64_t nnz = 0;
    {
        const char *last = (const char *)p + (size_t)C * (p_is_int64 ? 8 : 4);
        if (bn_is_device_ptr(p)) {
            if (p_is_int64) BN_CUDA(ctx, cudaMemcpy(&nnz, last, 8, cudaMemcpyDeviceToHost));
            else {
                int32_t t;
                BN_CUDA(ctx, cudaMemcpy(&t, last, 4, cudaMemcpyDeviceToHost));
                nnz = t;
            }
        } else {
            nnz = p_is_int64 ? *(const int64_t *)last : (int64_t) * (const int32_t *)last;
        }
    }
    if (nnz < 0) return bn_fail(ctx, BN_ERR_INVALID, "negative nnz");
    if (nnz > 0 && (!i || !x)) return bn_fail(ctx, BN_ERR_INVALID, "i/x missing");
    if (nnz >= 4294967295LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "nnz >= 2^32 per shard: split the matrix by gene blocks");
    bn_mat *m = new bn_mat();
    m->ctx = ctx;
    m->G = G;
    m->C = C;
    m->nnz = nnz;
    auto fail = [&](int rc) {
        bn_mat_free(m);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaMallocAsync((void **)&m->colptr, (size_t)(C + 1) * 8, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&m->rowidx, (size_t)(nnz ? nnz : 1) * 4, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&m->val, (size_t)(nnz ? nnz : 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "matrix allocation failed: %s", cudaGetErrorString(e)));
    if (p_is_int64) {
        if ((e = cudaMemcpyAsync(m->colptr, p, (size_t)(C + 1) * 8, cudaMemcpyDefault, ctx->stream)) != cudaSuccess)
            return fail(bn_fail(ctx, BN_ERR_CUDA, "copy p: %s", cudaGetErrorString(e)));
    } else {
        DevIn<int32_t> p32;
        int rc = p32.bind(ctx, (const int32_t *)p, (size_t)C + 1);
        if (rc) return fail(rc);
        k_widen_colptr<<<(unsigned)((C + 1 + 255) / 256), 256, 0, ctx->stream>>>(p32.p, m->colptr, C + 1);
        ctx->launches++;
    }
    if (nnz) {
        if ((e = cudaMemcpyAsync(m->rowidx, i, (size_t)nnz * 4, cudaMemcpyDefault, ctx->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(m->val, x, (size_t)nnz * 8, cudaMemcpyDefault, ctx->stream)) != cudaSuccess)
            return fail(bn_fail(ctx, BN_ERR_CUDA, "copy i/x: %s", cudaGetErrorString(e)));
    }
    int rc = bn_mat_finish(ctx, m);
    if (rc) return fail(rc);
    *out = m;
    return BN_OK;
}

// dense -> CSC -------------------------------------------------------------------
__global__ void k_dense_count(int64_t G, int64_t C, const double *__restrict__ A, int64_t *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const double *col = A + G * j;
        int c = 0;
        for (int64_t i = lane; i < G; i += 32) c += (col[i] != 0.0);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) cnt[j] = c;
    }
}

__global__ void k_dense_fill(int64_t G, int64_t C, const double *__restrict__ A, const int64_t *__restrict__ colptr,
                             int32_t *__restrict__ rowidx, double *__restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const double *col = A + G * j;
        int64_t base = colptr[j];
        for (int64_t i0 = 0; i0 < G; i0 += 32) {
            const int64_t i = i0 + lane;
            const double v = (i < G) ? col[i] : 0.0;
            const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
            if (v != 0.0) {
                const int64_t q = base + __popc(mask & ((1u << lane) - 1u));
                rowidx[q] = (int32_t)i;
                val[q] = v;
            }
            base += __popc(mask);
        }
    }
}

extern "C" int bn_mat_from_dense(bn_ctx *ctx, int64_t G, int64_t C, const double *dense, bn_mat **out)
{
    if (!ctx || !out) return BN_ERR_INVALID;
    *out = nullptr;
    if (G <= 0 || C <= 0 || !dense) return bn_fail(ctx, BN_ERR_INVALID, "bn_mat_from_dense: bad arguments");
    if (G > 2147483647LL) return bn_fail(ctx, BN_ERR_INVALID, "G exceeds int32 row indices");
    cudaSetDevice(ctx->device);
    DevIn<double> A;
    BN_TRY(A.bind(ctx, dense, (size_t)G * C));
    bn_mat *m = new bn_mat();
    m->ctx = ctx;
    m->G = G;
    m->C = C;
    auto fail = [&](int rc) {
        bn_mat_free(m);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaMallocAsync((void **)&m->colptr, (size_t)(C + 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "colptr: %s", cudaGetErrorString(e)));
    DevBuf<int64_t> cnt;
    int rc = cnt.alloc(ctx, (size_t)C + 1);
    if (rc) return fail(rc);
    cudaMemsetAsync(cnt.p, 0, (size_t)(C + 1) * 8, ctx->stream);
    const int blocks = (int)std::min<int64_t>((C + 7) / 8, 148 * 16);
    k_dense_count<<<blocks, 256, 0, ctx->stream>>>(G, C, A.p, cnt.p);
    ctx->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, m->colptr, (int)(C + 1), ctx->stream);
    DevBuf<char> tmp;
    if ((rc = tmp.alloc(ctx, tb))) return fail(rc);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, m->colptr, (int)(C + 1), ctx->stream);
    ctx->launches++;
    int64_t nnz = 0;
    if ((e = cudaMemcpyAsync(&nnz, m->colptr + C, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_CUDA, "dense scan: %s", cudaGetErrorString(e)));
    if (nnz >= 4294967295LL) return fail(bn_fail(ctx, BN_ERR_UNSUPPORTED, "nnz >= 2^32 per shard"));
    m->nnz = nnz;
    if ((e = cudaMallocAsync((void **)&m->rowidx, (size_t)(nnz ? nnz : 1) * 4, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&m->val, (size_t)(nnz ? nnz : 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "matrix allocation failed: %s", cudaGetErrorString(e)));
    k_dense_fill<<<blocks, 256, 0, ctx->stream>>>(G, C, A.p, m->colptr, m->rowidx, m->val);
    ctx->launches++;
    if ((rc = bn_mat_finish(ctx, m))) return fail(rc);
    *out = m;
    return BN_OK;
}

extern "C" void bn_mat_free(bn_mat *m)
{
    if (!m) return;
    // stream-ordered frees: the memory returns to the context's pool once the work already queued
    // on the stream has finished, and the next matrix re-uses it without going to the driver
    if (m->ctx) {
        cudaSetDevice(m->ctx->device);
        cudaStream_t st = m->ctx->stream;
        if (m->colptr) cudaFreeAsync(m->colptr, st);
        if (m->rowidx) cudaFreeAsync(m->rowidx, st);
        if (m->val) cudaFreeAsync(m->val, st);
        if (m->rowptr) cudaFreeAsync(m->rowptr, st);
        if (m->gm) cudaFreeAsync(m->gm, st);
        if (m->cell) cudaFreeAsync(m->cell, st);
        if (m->rval) cudaFreeAsync(m->rval, st);
        if (m->blockptr) cudaFreeAsync(m->blockptr, st);
    }
    delete m;
}

extern "C" int bn_mat_info(const bn_mat *m, int64_t out[4])
{
    if (!m || !out) return BN_ERR_INVALID;
    out[0] = m->G;
    out[1] = m->C;
    out[2] = m->nnz;
    out[3] = m->is_count ? 1 : 0;
    return BN_OK;
}

// Drop the gene-major twin so the next per-gene stage rebuilds it (benchmarks time the
// build as part of every step; a fresh matrix would pay it too).
extern "C" int bn_mat_drop_gene_major(bn_mat *m)
{
    if (!m) return BN_ERR_INVALID;
    m->has_csr = false; // the buffers stay with the matrix: the rebuild overwrites them in place
    return BN_OK;
}

// Measured FP64 FMA throughput of this device (flop/s), the denominator for the likelihood
// kernels' roofline: 8 independent DFMA chains per thread, all SMs, about 10 ms.
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int bn_debug_fp64_peak(bn_ctx *ctx, double *flops_per_s)
{
    if (!ctx || !flops_per_s) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const int blocks = ctx->prop.multiProcessorCount * 8, iters = 20000;
    DevBuf<double> o;
    BN_TRY(o.alloc(ctx, (size_t)blocks * 256));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        BN_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
        BN_LAUNCH(ctx, k_fp64_peak, blocks, 256, 0, o.p, iters);
        BN_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
        BN_TRY(bn_sync(ctx));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
        if (rep > 0 && ms < best) best = ms;
    }
    *flops_per_s = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (best * 1e-3);
    return BN_OK;
}

// Measured HBM WRITE bandwidth (GB/s): the posterior kernels are write-only streams, so this -- not the
// read+write copy figure -- is what they can reach.  pattern 0: one linear stream, every warp store
// 256 contiguous bytes; pattern 1: the 3-D layout, a warp stores 256 B to each of `planes` planes
// (bytes/planes apart), `run` stores of 256 B per visit, before moving on, like the S draws of one gene block.
__global__ void __launch_bounds__(256) k_write_peak(double *out, int64_t n, int planes, int run)
{
    const int64_t per_plane = n / planes;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t step = 32 * (int64_t)run; // elements a warp writes to one plane before switching
    for (int64_t c = warp * step; c + step <= per_plane; c += nwarps * step)
        for (int s = 0; s < planes; s++)
            for (int q = 0; q < run; q++) __stcs(out + s * per_plane + c + q * 32 + lane, (double)(c + s));
}

extern "C" int bn_debug_write_peak(bn_ctx *ctx, int64_t bytes, int planes, int run, double *gb_per_s)
{
    if (!ctx || !gb_per_s || bytes < (1 << 20) || planes < 1 || run < 1) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const int64_t n = bytes / 8;
    DevBuf<double> o;
    BN_TRY(o.alloc(ctx, (size_t)n));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        BN_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
        BN_LAUNCH(ctx, k_write_peak, ctx->prop.multiProcessorCount * 16, 256, 0, o.p, n, planes, run);
        BN_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
        BN_TRY(bn_sync(ctx));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
        if (rep > 0 && ms < best) best = ms;
    }
    *gb_per_s = (double)(n / planes / (32 * run) * (32 * run) * planes) * 8.0 / (best * 1e-3) / 1e9;
    return BN_OK;
}

extern "C" int bn_mat_set_origin(bn_mat *m, int64_t gene0, int64_t cell0)
{
    if (!m || gene0 < 0 || cell0 < 0) return BN_ERR_INVALID;
    m->gene0 = gene0;
    m->cell0 = cell0;
    return BN_OK;
}

extern "C" int bn_mat_export_csc(bn_ctx *ctx, const bn_mat *m, int64_t *p, int32_t *i, double *x)
{
    if (!ctx || !m) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (p) BN_CUDA(ctx, cudaMemcpyAsync(p, m->colptr, (size_t)(m->C + 1) * 8, cudaMemcpyDefault, ctx->stream));
    if (i && m->nnz) BN_CUDA(ctx, cudaMemcpyAsync(i, m->rowidx, (size_t)m->nnz * 4, cudaMemcpyDefault, ctx->stream));
    if (x && m->nnz) BN_CUDA(ctx, cudaMemcpyAsync(x, m->val, (size_t)m->nnz * 8, cudaMemcpyDefault, ctx->stream));
    return bn_sync(ctx);
}

// column subset -------------------------------------------------------------------
__global__ void k_sel_len(const int64_t *__restrict__ colptr, const int64_t *__restrict__ idx, int64_t n, int64_t C,
                          int64_t *__restrict__ len, int *bad)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) {
        const int64_t j = idx[k];
        if (j < 0 || j >= C) {
            *bad = 1;
            len[k] = 0;
        } else
            len[k] = colptr[j + 1] - colptr[j];
    }
    if (k == n) len[k] = 0;
}

__global__ void k_sel_copy(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                           const double *__restrict__ val, const int64_t *__restrict__ idx, int64_t n,
                           const int64_t *__restrict__ ncolptr, int32_t *__restrict__ nrow, double *__restrict__ nval)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += nwarp) {
        const int64_t src = colptr[idx[k]], dst = ncolptr[k], len = ncolptr[k + 1] - dst;
        for (int64_t t = lane; t < len; t += 32) {
            nrow[dst + t] = rowidx[src + t];
            nval[dst + t] = val[src + t];
        }
    }
}

extern "C" int bn_mat_select_cols(bn_ctx *ctx, const bn_mat *m, const int64_t *idx, int64_t n, bn_mat **out)
{
    if (!ctx || !m || !out || !idx || n <= 0) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_mat_select_cols: bad arguments") : BN_ERR_INVALID;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    DevIn<int64_t> di;
    BN_TRY(di.bind(ctx, idx, (size_t)n));
    DevBuf<int64_t> len;
    BN_TRY(len.alloc(ctx, (size_t)n + 1));
    DevBuf<int> bad;
    BN_TRY(bad.alloc(ctx, 1));
    BN_CUDA(ctx, cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
    BN_LAUNCH(ctx, k_sel_len, (unsigned)((n + 1 + 255) / 256), 256, 0, m->colptr, di.p, n, m->C, len.p, bad.p);
    bn_mat *s = new bn_mat();
    s->ctx = ctx;
    s->G = m->G;
    s->C = n;
    s->gene0 = m->gene0;
    s->is_count = m->is_count;
    auto fail = [&](int rc) {
        bn_mat_free(s);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaMallocAsync((void **)&s->colptr, (size_t)(n + 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "colptr: %s", cudaGetErrorString(e)));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, s->colptr, (int)(n + 1), ctx->stream);
    DevBuf<char> tmp;
    int rc = tmp.alloc(ctx, tb);
    if (rc) return fail(rc);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.p, s->colptr, (int)(n + 1), ctx->stream);
    ctx->launches++;
    int64_t nnz = 0;
    int hbad = 0;
    cudaMemcpyAsync(&nnz, s->colptr + n, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&hbad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_CUDA, "select_cols: %s", cudaGetErrorString(e)));
    if (hbad) return fail(bn_fail(ctx, BN_ERR_INVALID, "column index out of range"));
    s->nnz = nnz;
    if ((e = cudaMallocAsync((void **)&s->rowidx, (size_t)(nnz ? nnz : 1) * 4, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&s->val, (size_t)(nnz ? nnz : 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "matrix allocation failed: %s", cudaGetErrorString(e)));
    const int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 16);
    k_sel_copy<<<blocks, 256, 0, ctx->stream>>>(m->colptr, m->rowidx, m->val, di.p, n, s->colptr, s->rowidx, s->val);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(bn_fail(ctx, BN_ERR_CUDA, "k_sel_copy: %s", cudaGetErrorString(e)));
    *out = s;
    return bn_sync(ctx);
}

// ---------------------------------------------------------------------------------
// gene-major twin: stable radix sort of (row, position) pairs, then a gather.
// Stable => cells stay in increasing order inside each gene => every per-gene
// reduction has one fixed summation order, independent of launch geometry.
// ---------------------------------------------------------------------------------
__global__ void k_iota_u32(uint32_t *__restrict__ v, int64_t n)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) v[k] = (uint32_t)k;
}

// Row pointer from the SORTED row keys: position k opens every row in (key[k-1], key[k]]; rows after the last
// key are empty.  One coalesced 4 B read per non-zero (a histogram with atomics took 4x as long).
__global__ void __launch_bounds__(256) k_rowptr_from_sorted(const int32_t *__restrict__ key, int64_t nnz, int64_t G,
                                                            int64_t *__restrict__ rowptr)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += gridDim.x * (int64_t)blockDim.x) {
        const int32_t cur = __ldg(key + k), prev = k > 0 ? __ldg(key + k - 1) : -1;
        for (int32_t r = prev + 1; r <= cur; r++) rowptr[r] = k;
        if (k == nnz - 1)
            for (int64_t r = (int64_t)cur + 1; r <= G; r++) rowptr[r] = nnz;
    }
}

__global__ void k_csr_gather(int64_t nnz, int64_t C, const uint32_t *__restrict__ pos, const int64_t *__restrict__ colptr,
                             const double *__restrict__ val, int32_t *__restrict__ cell, double *__restrict__ rval)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * (int64_t)blockDim.x) {
        const int64_t k = pos[q];
        // column of position k: last j with colptr[j] <= k
        int64_t lo = 0, hi = C; // invariant: colptr[lo] <= k < colptr[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (colptr[mid] <= k) lo = mid;
            else hi = mid;
        }
        cell[q] = (int32_t)lo;
        rval[q] = val[k];
    }
}

// (cell << 32) | count for every non-zero, in CSC order: one warp per column
__global__ void __launch_bounds__(256) k_pack_nz(int64_t C, const int64_t *__restrict__ colptr, const double *__restrict__ val,
                                                 uint64_t *__restrict__ packed)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        const uint64_t hi = (uint64_t)j << 32;
        int64_t k = a + lane;
        for (; k + 96 < b; k += 128) { // four loads in flight per lane
            const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + 32), v2 = __ldcs(val + k + 64), v3 = __ldcs(val + k + 96);
            packed[k] = hi | (uint64_t)(uint32_t)v0;
            packed[k + 32] = hi | (uint64_t)(uint32_t)v1;
            packed[k + 64] = hi | (uint64_t)(uint32_t)v2;
            packed[k + 96] = hi | (uint64_t)(uint32_t)v3;
        }
        for (; k < b; k += 32) packed[k] = hi | (uint64_t)(uint32_t)__ldcs(val + k);
    }
}

// Blocked column pointers: one warp per column walks the (sorted) row indices once and records where every
// BN_ROWBLOCK-row block starts.  4 B per non-zero read, (G / 128 + 1) * C * 4 B written, once per matrix: it
// replaces a 12-step binary search per (warp, column) in every posterior launch.
__global__ void __launch_bounds__(256) k_blockptr(int64_t C, int64_t nblock, const int64_t *__restrict__ colptr,
                                                  const int32_t *__restrict__ rowidx, uint32_t *__restrict__ blockptr)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        for (int64_t k = a + lane; k < b; k += 32) {
            const int64_t t = rowidx[k] / BN_ROWBLOCK, tp = (k > a) ? rowidx[k - 1] / BN_ROWBLOCK : -1;
            for (int64_t q = tp + 1; q <= t; q++) blockptr[q * C + j] = (uint32_t)k; // blocks opened by entry k
        }
        // blocks after the last entry (and the closing row) are empty: they start at the end of the column
        const int64_t tl = (b > a) ? rowidx[b - 1] / BN_ROWBLOCK : -1;
        for (int64_t q = tl + 1 + lane; q <= nblock; q += 32) blockptr[q * C + j] = (uint32_t)b;
    }
}

int bn_mat_ensure_blockptr(bn_ctx *ctx, const bn_mat *cm)
{
    bn_mat *m = const_cast<bn_mat *>(cm); // a cache, like the gene-major twin
    if (m->blockptr) return BN_OK;
    // positions are 32-bit in the posterior kernels, which also look up to 6 x 32 entries past a slice start
    if (m->nnz >= 4294967295LL - 4096) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "nnz >= 2^32 - 4096 per shard");
    cudaSetDevice(ctx->device);
    m->nblock = (m->G + BN_ROWBLOCK - 1) / BN_ROWBLOCK;
    BN_CUDA(ctx, cudaMallocAsync((void **)&m->blockptr, (size_t)(m->nblock + 1) * (size_t)m->C * 4, ctx->stream));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
    BN_LAUNCH(ctx, k_blockptr, blocks, 256, 0, m->C, m->nblock, m->colptr, m->rowidx, m->blockptr);
    return BN_OK;
}

int bn_mat_ensure_csr(bn_ctx *ctx, bn_mat *m)
{
    if (m->has_csr) return BN_OK;
    cudaSetDevice(ctx->device);
    const int64_t nnz = m->nnz, G = m->G;
    if (nnz > 2147483647LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "nnz > 2^31-1 per shard in the gene-major build");
    if (!m->rowptr) BN_CUDA(ctx, cudaMallocAsync((void **)&m->rowptr, (size_t)(G + 1) * 8, ctx->stream));
    int bits = 1;
    while ((1LL << bits) < G) bits++;
    if (m->is_count) {
        // stable sort of (row -> packed (cell, count)) pairs: the sorted values ARE the gene-major matrix
        if (!m->gm) BN_CUDA(ctx, cudaMallocAsync((void **)&m->gm, (size_t)(nnz ? nnz : 1) * 8, ctx->stream));
        if (nnz) {
            struct { uint64_t *p; } packed;
            struct { int32_t *p; } kout;
            BN_TRY(bn_ctx_scratch(ctx, 2, (size_t)nnz * 8, (void **)&packed.p));
            BN_TRY(bn_ctx_scratch(ctx, 3, (size_t)nnz * 4, (void **)&kout.p));
            const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
            BN_LAUNCH(ctx, k_pack_nz, blocks, 256, 0, m->C, m->colptr, m->val, packed.p);
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t *)m->rowidx, kout.p, (const uint64_t *)packed.p, m->gm,
                                            (int)nnz, 0, bits, ctx->stream);
            struct { char *p; } tmp;
            BN_TRY(bn_ctx_scratch(ctx, 4, tb, (void **)&tmp.p));
            BN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, (const int32_t *)m->rowidx, kout.p,
                                                         (const uint64_t *)packed.p, m->gm, (int)nnz, 0, bits, ctx->stream));
            ctx->launches += (bits + 7) / 8 + 1;
            BN_LAUNCH(ctx, k_rowptr_from_sorted, (unsigned)std::min<int64_t>((nnz + 255) / 256, 148 * 32), 256, 0, kout.p, nnz, G, m->rowptr);
        }
    } else {
        if (!m->cell) BN_CUDA(ctx, cudaMallocAsync((void **)&m->cell, (size_t)(nnz ? nnz : 1) * 4, ctx->stream));
        if (!m->rval) BN_CUDA(ctx, cudaMallocAsync((void **)&m->rval, (size_t)(nnz ? nnz : 1) * 8, ctx->stream));
        if (nnz) {
            DevBuf<int32_t> kout;
            DevBuf<uint32_t> v0, v1;
            BN_TRY(kout.alloc(ctx, (size_t)nnz));
            BN_TRY(v0.alloc(ctx, (size_t)nnz));
            BN_TRY(v1.alloc(ctx, (size_t)nnz));
            BN_LAUNCH(ctx, k_iota_u32, (unsigned)((nnz + 255) / 256), 256, 0, v0.p, nnz);
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t *)m->rowidx, kout.p, (const uint32_t *)v0.p, v1.p,
                                            (int)nnz, 0, bits, ctx->stream);
            DevBuf<char> tmp;
            BN_TRY(tmp.alloc(ctx, tb));
            BN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, (const int32_t *)m->rowidx, kout.p, (const uint32_t *)v0.p,
                                                         v1.p, (int)nnz, 0, bits, ctx->stream));
            ctx->launches += (bits + 7) / 8 + 1;
            BN_LAUNCH(ctx, k_rowptr_from_sorted, (unsigned)std::min<int64_t>((nnz + 255) / 256, 148 * 32), 256, 0, kout.p, nnz, G, m->rowptr);
            BN_LAUNCH(ctx, k_csr_gather, 148 * 8, 256, 0, nnz, m->C, v1.p, m->colptr, m->val, m->cell, m->rval);
        }
    }
    if (!nnz) BN_CUDA(ctx, cudaMemsetAsync(m->rowptr, 0, (size_t)(G + 1) * 8, ctx->stream));
    m->has_csr = true;
    return BN_OK;
}
