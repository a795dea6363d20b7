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
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2; k++) {
        cudaEventCreateWithFlags(&c->ev_comp[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming);
    }
    {
        // side streams get the highest priority: their short, low-occupancy kernels slot in between the CTAs of the long
        // kernel on `stream` instead of queueing behind its whole grid
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        for (auto &st : c->aux) cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi);
    }
    for (auto &ev : c->ev_aux) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaHostAlloc((void **)&c->pinned, 64 * sizeof(double), cudaHostAllocDefault);
    c->fit_parked = reinterpret_cast<unsigned int *>(c->pinned + 40);
    *c->fit_parked = 0;
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
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (int k = 0; k < 2; k++) {
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
        if (ctx->ev_copy[k]) cudaEventDestroy(ctx->ev_copy[k]);
        if (ctx->host_tile[k]) cudaFreeHost(ctx->host_tile[k]);
    }
    for (auto &st : ctx->aux)
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (auto &ev : ctx->ev_aux)
        if (ev) cudaEventDestroy(ev);
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

// Gene-sharded runs: every shard must take the same branch before a collective, or the others wait forever in the hook.
// Returns the largest status code over all shards (one max all-reduce of a double); with world == 1 the local code.
int bn_agree(bn_ctx *ctx, int local_rc)
{
    if (ctx->world <= 1) return local_rc;
    DevBuf<double> d;
    if (d.alloc(ctx, 1) != BN_OK) return BN_ERR_OOM;
    const double v = (double)local_rc;
    double *h = ctx->pinned + 48;
    *h = v;
    if (cudaMemcpyAsync(d.p, h, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return BN_ERR_CUDA;
    const std::string keep = ctx->err;
    const int rc = bn_allreduce(ctx, d.p, 1, 2);
    if (rc != BN_OK) return rc;
    if (cudaMemcpyAsync(h, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return BN_ERR_CUDA;
    const int g = (int)*h;
    if (g != BN_OK && local_rc == BN_OK)
        bn_fail(ctx, g, "another gene shard failed with status %d (this shard's inputs were fine)", g);
    else
        ctx->err = keep;
    return g;
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
    m->has_blockptr = false; // the blocked column pointers are part of the build (and shared with the posterior kernels)
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
// Blocked column pointers: one warp per column walks the (sorted) row indices once and records where every
// BN_ROWBLOCK-row block starts, as an OFFSET from the start of the column (uint32: a column has at most G < 2^31
// entries, while absolute positions need 64 bits once a shard holds >= 2^32 non-zeros -- BASELINE config 5 on one
// GPU).  4 B per non-zero read, (G / 128 + 1) * C * 4 B written, once per matrix: it replaces a 12-step binary
// search per (warp, column) in every posterior launch and gives the gene-major build its row tiles.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_blockptr(int64_t C, int64_t nblock, const int64_t *__restrict__ colptr,
                                                  const int32_t *__restrict__ rowidx, uint32_t *__restrict__ blockptr)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        for (int64_t k0 = a; k0 < b; k0 += 128) { // four loads in flight per lane; the previous entry comes by shuffle
            int32_t r[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r[u] = (k0 + u * 32 + lane < b) ? __ldg(rowidx + k0 + u * 32 + lane) : 0x7fffffff;
            int32_t last = (k0 > a) ? __ldg(rowidx + k0 - 1) : -1; // entry before this batch (warp-uniform address)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int32_t prev = __shfl_up_sync(0xffffffffu, r[u], 1);
                if (lane == 0) prev = last;
                last = __shfl_sync(0xffffffffu, r[u], 31);
                const int64_t k = k0 + u * 32 + lane;
                if (k < b) {
                    const int t = r[u] / BN_ROWBLOCK, tp = prev < 0 ? -1 : prev / BN_ROWBLOCK;
                    for (int q = tp + 1; q <= t; q++) blockptr[(int64_t)q * C + j] = (uint32_t)(k - a); // blocks opened by entry k
                }
            }
        }
        // blocks after the last entry (and the closing row) are empty: they start at the end of the column
        const int64_t tl = (b > a) ? rowidx[b - 1] / BN_ROWBLOCK : -1;
        for (int64_t q = tl + 1 + lane; q <= nblock; q += 32) blockptr[q * C + j] = (uint32_t)(b - a);
    }
}

int bn_mat_ensure_blockptr(bn_ctx *ctx, const bn_mat *cm)
{
    bn_mat *m = const_cast<bn_mat *>(cm); // a cache, like the gene-major twin
    if (m->has_blockptr) return BN_OK;
    cudaSetDevice(ctx->device);
    m->nblock = (m->G + BN_ROWBLOCK - 1) / BN_ROWBLOCK;
    if (!m->blockptr) BN_CUDA(ctx, cudaMallocAsync((void **)&m->blockptr, (size_t)(m->nblock + 1) * (size_t)m->C * 4, ctx->stream));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
    BN_LAUNCH(ctx, k_blockptr, blocks, 256, 0, m->C, m->nblock, m->colptr, m->rowidx, m->blockptr);
    m->has_blockptr = true;
    return BN_OK;
}

// ---------------------------------------------------------------------------------
// Gene-major twin (what `Data[g, ]` at R/PRIOR_FUNCTIONS.R:442 extracts, for all genes at once): a hand-written
// stable counting transpose of the CSC arrays.
//
// Work unit = (tile of GM_TILE_ROWS rows) x (range of columns), one WARP each.  Row indices are sorted inside a column, so
// the tile's part of a column is one contiguous slice, and the blocked column pointers give its bounds with two loads.
// Rows are distinct inside a slice, so the warp advances a per-row cursor in shared memory with plain loads/stores (no
// atomics) and walks its columns in increasing order: cells come out ascending inside every gene -- the same order a
// stable sort by row gives, whatever the launch geometry, so every per-gene reduction downstream keeps ONE summation
// order (bit-reproducible across GPU counts).
//   pass 1 (k_gm_count):   cnt[p][r] = entries of row r in column range p             reads 4 B per non-zero
//   k_gm_offsets:          cnt[p][r] -> entries of row r in ranges < p; rowtot[r]     P x G words
//   exclusive scan of rowtot -> rowptr
//   pass 2 (k_gm_scatter): position = rowptr[r] + cnt[p][r] + (rank inside the range) reads 12 B, writes 8 B per non-zero
// 24 B of DRAM traffic per non-zero (+ 4 B for the blocked column pointers, shared with the posterior kernels) against
// ~64 B for pack + two cub onesweep passes + row-pointer extraction, and no nnz-sized scratch: that is what lets the
// 3.4e9 non-zeros of BASELINE config 5 (41 GB CSC + 27 GB gene-major) sit on one 180 GB GPU.
// ---------------------------------------------------------------------------------
#define GM_TILE_BLOCKS 8                              // row blocks per tile
#define GM_TILE_ROWS (GM_TILE_BLOCKS * BN_ROWBLOCK)   // 1024 rows: ~84 entries per column slice at 8 % non-zeros
#define GM_WARPS 8
#define GM_PF 3                                       // slice entries per lane prefetched one column ahead

// Pass 1: warp per (row tile, fine column range); rows are distinct inside a column slice, so the per-row counters in
// shared memory take plain increments, one column at a time.
__global__ void __launch_bounds__(GM_WARPS * 32, 3)
    k_gm_count(int64_t G, int64_t C, int64_t nblock, int64_t ntile, int P, const int64_t *__restrict__ colptr,
               const int32_t *__restrict__ rowidx, const uint32_t *__restrict__ blockptr, uint32_t *__restrict__ cnt)
{
    extern __shared__ uint32_t s_cnt_all[]; // [GM_WARPS][GM_TILE_ROWS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *cur = s_cnt_all + warp * GM_TILE_ROWS;
    const int64_t ntask = ntile * P;
    // neighbouring warps take neighbouring row tiles of the SAME column range: their slices are adjacent in memory
    for (int64_t task = (int64_t)blockIdx.x * GM_WARPS + warp; task < ntask; task += (int64_t)gridDim.x * GM_WARPS) {
        const int64_t p = task / ntile, t = task - p * ntile;
        const int64_t row0 = t * GM_TILE_ROWS;
        const int nrow = (int)(G - row0 < GM_TILE_ROWS ? G - row0 : GM_TILE_ROWS);
        const int64_t c0 = C * p / P, c1 = C * (p + 1) / P;
        const int64_t blo = t * GM_TILE_BLOCKS, bhi = blo + GM_TILE_BLOCKS < nblock ? blo + GM_TILE_BLOCKS : nblock;
        const uint32_t *bp_lo = blockptr + blo * C, *bp_hi = blockptr + bhi * C;
        for (int r = lane; r < nrow; r += 32) cur[r] = 0;
        __syncwarp();
        // bounds of 32 columns at a time (coalesced), handed out by shuffle; the first GM_PF x 32 row indices of the
        // next column's slice are loaded while the current column is counted
        int32_t pr[GM_PF];
        int64_t ka = 0, kb = 0; // slice of the column held in pr
        for (int64_t jg = c0; jg < c1; jg += 32) {
            const int ng = (int)(c1 - jg < 32 ? c1 - jg : 32);
            int64_t my_a = 0, my_b = 0;
            if (lane < ng) {
                const int64_t cs = __ldg(colptr + jg + lane);
                my_a = cs + __ldg(bp_lo + jg + lane);
                my_b = cs + __ldg(bp_hi + jg + lane);
            }
            for (int jj = 0; jj < ng; jj++) {
                if (jj == 0) { // first column of the group: nothing was prefetched for it
                    ka = __shfl_sync(0xffffffffu, my_a, 0);
                    kb = __shfl_sync(0xffffffffu, my_b, 0);
#pragma unroll
                    for (int q = 0; q < GM_PF; q++) pr[q] = (ka + q * 32 + lane < kb) ? __ldg(rowidx + ka + q * 32 + lane) : -1;
                }
                int32_t r_[GM_PF];
#pragma unroll
                for (int q = 0; q < GM_PF; q++) r_[q] = pr[q];
                const int64_t a = ka, b = kb;
                if (jj + 1 < ng) {
                    ka = __shfl_sync(0xffffffffu, my_a, jj + 1);
                    kb = __shfl_sync(0xffffffffu, my_b, jj + 1);
#pragma unroll
                    for (int q = 0; q < GM_PF; q++) pr[q] = (ka + q * 32 + lane < kb) ? __ldg(rowidx + ka + q * 32 + lane) : -1;
                }
#pragma unroll
                for (int q = 0; q < GM_PF; q++)
                    if (r_[q] >= 0) cur[r_[q] - (int)row0]++;
                // a longer slice (dense matrices): the rest straight from memory, four loads in flight per lane
                for (int64_t k = a + GM_PF * 32 + lane; k < b; k += 128) {
                    int32_t rr[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) rr[u] = (k + u * 32 < b) ? __ldg(rowidx + k + u * 32) : -1;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (rr[u] >= 0) cur[rr[u] - (int)row0]++;
                }
                __syncwarp(); // column j's increments are visible before column j + 1 touches the same rows
            }
        }
        for (int r = lane; r < nrow; r += 32) cnt[p * G + row0 + r] = cur[r];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Pass 2, CTA per (row tile, coarse column range).  Scattering one entry at a time (8 B per row per visit) ran at
// 3-6e10 stores/s whether a warp walked its columns alone (4.0 ms per 1e8 non-zeros: millions of partly written
// sectors open in L2) or a CTA placed 64 columns together (1.7 ms): the store TRANSACTION rate is the limit.  So the
// CTA takes as many consecutive columns as fit GS_CAP staged entries (<= 64), transposes that block in shared memory
// and writes every row's run of entries contiguously (1.1 ms):
//   a  mask[r] |= 1 << jj for every entry (r, column jj of the block)            (order-free shared-memory atomics)
//   b  row counts = popc(mask), exclusive scan -> start of every row's run in the stage, advance the global cursors
//   c  entry (r, jj) -> stage[start[r] + popc(mask[r] & below(jj))]             (rank = columns before jj: ascending cells)
//   d  stage -> gm, thread per staged entry: a row's run is contiguous on both sides
// ---------------------------------------------------------------------------------
#define GS_THREADS 512
#define GS_CAP 8192
#define GS_WMAX 64
#define GS_MINB 2

template <bool PACKED>
__global__ void __launch_bounds__(GS_THREADS, GS_MINB)
    k_gm_scatter(int64_t G, int64_t C, int64_t nblock, int64_t ntile, int P, int F, int Q, const int64_t *__restrict__ colptr,
                 const int32_t *__restrict__ rowidx, const double *__restrict__ val, const uint32_t *__restrict__ blockptr,
                 const uint32_t *__restrict__ cnt, const int64_t *__restrict__ rowptr, uint64_t *__restrict__ gm,
                 int32_t *__restrict__ cell, double *__restrict__ rval)
{
    extern __shared__ __align__(16) unsigned char gs_smem[];
    uint64_t *stage = reinterpret_cast<uint64_t *>(gs_smem);                        // [GS_CAP] packed word or value bits
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(stage + GS_CAP); // [GM_TILE_ROWS]
    int64_t *gcur = reinterpret_cast<int64_t *>(mask + GM_TILE_ROWS);               // [GM_TILE_ROWS] next free position of the row
    int64_t *gbase = gcur + GM_TILE_ROWS;                                           // [GM_TILE_ROWS] position of stage[e] = gbase[row] + e
    uint32_t *rstart = reinterpret_cast<uint32_t *>(gbase + GM_TILE_ROWS);          // [GM_TILE_ROWS]
    uint16_t *srow = reinterpret_cast<uint16_t *>(rstart + GM_TILE_ROWS);           // [GS_CAP]
    int32_t *scell = reinterpret_cast<int32_t *>(srow + GS_CAP);                    // [GS_CAP] (real-valued matrices only)
    __shared__ int64_t s_ka[GS_WMAX];
    __shared__ uint32_t s_len[GS_WMAX];
    __shared__ uint32_t s_wsum[GS_THREADS / 32];
    __shared__ int s_W;
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = GS_THREADS / 32, RPT = GM_TILE_ROWS / GS_THREADS; // rows per thread in the scan
    for (int64_t task = blockIdx.x; task < ntile * Q; task += gridDim.x) {
        const int64_t q = task / ntile, t = task - q * ntile; // neighbouring CTAs: neighbouring row tiles of the same columns
        const int64_t row0 = t * GM_TILE_ROWS;
        const int nrow = (int)(G - row0 < GM_TILE_ROWS ? G - row0 : GM_TILE_ROWS);
        const int64_t p0 = q * F, p1 = (q + 1) * F < P ? (q + 1) * F : P;
        const int64_t c0 = C * p0 / P, c1 = C * p1 / P;
        const int64_t blo = t * GM_TILE_BLOCKS, bhi = blo + GM_TILE_BLOCKS < nblock ? blo + GM_TILE_BLOCKS : nblock;
        const uint32_t *bp_lo = blockptr + blo * C, *bp_hi = blockptr + bhi * C;
        __syncthreads();
        for (int r = tid; r < GM_TILE_ROWS; r += GS_THREADS) {
            mask[r] = 0ull;
            gcur[r] = r < nrow ? rowptr[row0 + r] + (int64_t)cnt[p0 * G + row0 + r] : 0;
        }
        for (int64_t j0 = c0; j0 < c1;) {
            // ---- the block: as many columns as fit the stage ---------------------------------------
            if (warp < 2) {
                const int jj = tid; // 0..63
                int64_t ka = 0;
                uint32_t len = 0;
                if (j0 + jj < c1) {
                    const int64_t cs = __ldg(colptr + j0 + jj);
                    const uint32_t lo = __ldg(bp_lo + j0 + jj), hi = __ldg(bp_hi + j0 + jj);
                    ka = cs + lo;
                    len = hi - lo;
                }
                uint32_t inc = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                s_ka[jj] = ka;
                s_len[jj] = len;
                asm volatile("bar.sync 1, 64;" ::: "memory"); // the two warps only
                if (warp == 1) inc += s_wsum[0];
                // columns whose cumulative length fits: a prefix (lengths are non-negative)
                const unsigned fits = __ballot_sync(0xffffffffu, inc <= GS_CAP && j0 + jj < c1);
                if (lane == 0) s_wsum[2 + warp] = (uint32_t)__popc(fits);
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (tid == 0) {
                    int W = (int)s_wsum[2];
                    if (W == 32) W += (int)s_wsum[3];
                    s_W = W > 0 ? W : 1; // one column's slice always fits (<= GM_TILE_ROWS <= GS_CAP entries)
                }
            }
            __syncthreads();
            const int W = s_W;
            // ---- a: which columns of the block hold each row -----------------------------------------
            for (int jj = warp; jj < W; jj += NW) {
                const int64_t ka = s_ka[jj];
                const uint32_t len = s_len[jj];
                const unsigned long long bit = 1ull << jj;
                for (uint32_t i = lane; i < len; i += 256) {
                    int32_t rr[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) rr[u] = (i + u * 32 < len) ? __ldg(rowidx + ka + i + u * 32) : -1;
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (rr[u] >= 0) atomicOr(mask + (rr[u] - (int)row0), bit);
                }
            }
            __syncthreads();
            // ---- b: run starts (exclusive scan of the row counts), global cursors ---------------------
            {
                uint32_t c[RPT], sum = 0;
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    c[i] = (uint32_t)__popcll(mask[tid * RPT + i]);
                    sum += c[i];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                uint32_t before = 0;
                for (int w = 0; w < warp; w++) before += s_wsum[w];
                uint32_t run = before + inc - sum;
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    const int r = tid * RPT + i;
                    rstart[r] = run;
                    const int64_t g0 = gcur[r];
                    gbase[r] = g0 - (int64_t)run;
                    gcur[r] = g0 + c[i];
                    run += c[i];
                }
                if (tid == GS_THREADS - 1) s_total = run;
            }
            __syncthreads();
            // ---- c: place the entries, cells ascending inside every row -------------------------------
            for (int jj = warp; jj < W; jj += NW) {
                const int64_t ka = s_ka[jj];
                const uint32_t len = s_len[jj];
                const unsigned long long below = (1ull << jj) - 1ull;
                const int64_t j = j0 + jj;
                for (uint32_t i = lane; i < len; i += 128) {
                    int32_t rr[4];
                    double vv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        rr[u] = (i + u * 32 < len) ? __ldg(rowidx + ka + i + u * 32) : -1;
                        vv[u] = (i + u * 32 < len) ? __ldcs(val + ka + i + u * 32) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (rr[u] >= 0) {
                            const int rl = rr[u] - (int)row0;
                            const uint32_t e = rstart[rl] + (uint32_t)__popcll(mask[rl] & below);
                            stage[e] = PACKED ? (((uint64_t)j << 32) | (uint64_t)(uint32_t)vv[u]) : (uint64_t)__double_as_longlong(vv[u]);
                            srow[e] = (uint16_t)rl;
                            if (!PACKED) scell[e] = (int32_t)j;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- d: stream the stage out; a row's run is contiguous in the stage and in the output ----
            {
                const uint32_t total = s_total;
                for (uint32_t e = tid; e < total; e += GS_THREADS) {
                    const int64_t pos = gbase[srow[e]] + (int64_t)e;
                    if (PACKED) gm[pos] = stage[e];
                    else {
                        rval[pos] = __longlong_as_double((long long)stage[e]);
                        cell[pos] = scell[e];
                    }
                }
                for (int r = tid; r < GM_TILE_ROWS; r += GS_THREADS) mask[r] = 0ull;
            }
            __syncthreads();
            j0 += W;
        }
    }
}

// cnt[p][r] (entries of row r in column range p) -> exclusive prefix over p; rowtot[r] = row length
__global__ void __launch_bounds__(256) k_gm_offsets(int64_t G, int P, uint32_t *__restrict__ cnt, int64_t *__restrict__ rowtot)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > G) return;
    if (r == G) {
        rowtot[G] = 0;
        return;
    }
    int64_t run = 0; // a row has at most C < 2^32 entries: the prefixes fit 32 bits
    int p = 0;
    for (; p + 8 <= P; p += 8) { // eight independent loads in flight, then the dependent prefix
        uint32_t c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) c[u] = cnt[(int64_t)(p + u) * G + r];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            cnt[(int64_t)(p + u) * G + r] = (uint32_t)run;
            run += c[u];
        }
    }
    for (; p < P; p++) {
        const uint32_t c = cnt[(int64_t)p * G + r];
        cnt[(int64_t)p * G + r] = (uint32_t)run;
        run += c;
    }
    rowtot[r] = run;
}

int bn_mat_ensure_csr(bn_ctx *ctx, bn_mat *m)
{
    if (m->has_csr) return BN_OK;
    cudaSetDevice(ctx->device);
    const int64_t nnz = m->nnz, G = m->G, C = m->C;
    if (C > 4294967295LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "more than 2^32 - 1 cells");
    if (!m->rowptr) BN_CUDA(ctx, cudaMallocAsync((void **)&m->rowptr, (size_t)(G + 1) * 8, ctx->stream));
    if (m->is_count) {
        if (!m->gm) BN_CUDA(ctx, cudaMallocAsync((void **)&m->gm, (size_t)(nnz ? nnz : 1) * 8, ctx->stream));
    } else {
        if (!m->cell) BN_CUDA(ctx, cudaMallocAsync((void **)&m->cell, (size_t)(nnz ? nnz : 1) * 4, ctx->stream));
        if (!m->rval) BN_CUDA(ctx, cudaMallocAsync((void **)&m->rval, (size_t)(nnz ? nnz : 1) * 8, ctx->stream));
    }
    if (!nnz) {
        BN_CUDA(ctx, cudaMemsetAsync(m->rowptr, 0, (size_t)(G + 1) * 8, ctx->stream));
        m->has_csr = true;
        return BN_OK;
    }
    BN_TRY(bn_mat_ensure_blockptr(ctx, m));
    const int sms = ctx->prop.multiProcessorCount;
    const int64_t ntile = (G + GM_TILE_ROWS - 1) / GM_TILE_ROWS;
    // column ranges: enough (tile, range) warps to fill the device 3 CTAs deep, at least 64 columns per range
    int64_t P = ((int64_t)sms * 3 * GM_WARPS * 2 + ntile - 1) / ntile;
    P = std::max<int64_t>(1, std::min<int64_t>(P, (C + 63) / 64));
    struct { uint32_t *p; } cnt;
    struct { int64_t *p; } rowtot;
    struct { char *p; } tmp;
    BN_TRY(bn_ctx_scratch(ctx, 2, (size_t)P * (size_t)G * 4, (void **)&cnt.p));
    BN_TRY(bn_ctx_scratch(ctx, 3, (size_t)(G + 1) * 8, (void **)&rowtot.p));
    const size_t smem = (size_t)GM_WARPS * GM_TILE_ROWS * sizeof(uint32_t);
    const int grid = (int)std::min<int64_t>((ntile * P + GM_WARPS - 1) / GM_WARPS, (int64_t)sms * 3);
    BN_LAUNCH(ctx, k_gm_count, grid, GM_WARPS * 32, smem, G, C, m->nblock, ntile, (int)P, m->colptr, m->rowidx, m->blockptr, cnt.p);
    BN_LAUNCH(ctx, k_gm_offsets, (unsigned)((G + 1 + 255) / 256), 256, 0, G, (int)P, cnt.p, rowtot.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, rowtot.p, m->rowptr, (int)(G + 1), ctx->stream);
    BN_TRY(bn_ctx_scratch(ctx, 4, tb, (void **)&tmp.p));
    BN_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, rowtot.p, m->rowptr, (int)(G + 1), ctx->stream));
    ctx->launches++;
    {
        // pass 2: CTAs over (row tile, coarse range of F fine column ranges), about two waves of GS_MINB CTAs per SM
        const int64_t want = (int64_t)sms * GS_MINB * 2;
        int64_t Q = std::max<int64_t>(1, std::min<int64_t>(P, (want + ntile - 1) / ntile));
        const int64_t F = (P + Q - 1) / Q;
        Q = (P + F - 1) / F;
        const size_t smem2 = (size_t)GS_CAP * 8 + (size_t)GM_TILE_ROWS * (8 + 8 + 8 + 4) + (size_t)GS_CAP * 2 + (m->is_count ? 0 : (size_t)GS_CAP * 4);
        const int grid2 = (int)std::min<int64_t>(ntile * Q, (int64_t)sms * GS_MINB);
#define GS_ARGS G, C, m->nblock, ntile, (int)P, (int)F, (int)Q, m->colptr, m->rowidx, m->val, m->blockptr, cnt.p, m->rowptr, m->gm, m->cell, m->rval
        if (m->is_count) {
            BN_CUDA(ctx, cudaFuncSetAttribute(k_gm_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            BN_LAUNCH(ctx, (k_gm_scatter<true>), grid2, GS_THREADS, smem2, GS_ARGS);
        } else {
            BN_CUDA(ctx, cudaFuncSetAttribute(k_gm_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            BN_LAUNCH(ctx, (k_gm_scatter<false>), grid2, GS_THREADS, smem2, GS_ARGS);
        }
#undef GS_ARGS
    }
    m->has_csr = true;
    return BN_OK;
}
