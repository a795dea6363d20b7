// bn_stream.cu -- delivery of the gene x cell posterior: results are 8 (x S) bytes per gene x cell, 26 GB - 6.9 TB at the
// BASELINE shapes, so they leave the device as column tiles:
//   * bn_post_to_host: a host destination of bn_post_mean / bn_post_mode / bn_post_sample is filled tile by tile; tile
//     t + 1 is computed on the context's stream while tile t travels on a second stream (PCIe carries ~55 GB/s, the
//     kernels write 3-5 TB/s: the bus is the bound, the compute hides behind it);
//   * bn_post_stream: the same pipeline into pinned tiles owned by the library, handed to a caller-supplied sink -- the
//     tile writer for results larger than host memory;
//   * bn_post_mean_csc: the posterior mean as a device-resident CSC matrix, what Main_mean_NB_spBay returns
//     (src/bayNorm_main.cpp:1598: the dense mean converted to sp_mat; out.sparse = TRUE at R/bayNorm.r:140-150).
#include <cub/cub.cuh>

#include "bn_internal.cuh"

struct bn_csc_result {
    bn_ctx *ctx = nullptr;
    int64_t G = 0, ncol = 0, nnz = 0;
    int64_t *colptr = nullptr;
    int32_t *rowidx = nullptr;
    double *val = nullptr;
};

static int64_t default_tile_cols(const bn_mat *m, int planes, int64_t ncol, int64_t tile_cols)
{
    if (tile_cols <= 0) tile_cols = (int64_t)(((size_t)1 << 30) / ((size_t)m->G * (size_t)planes * 8)); // ~1 GB of result per tile
    return std::max<int64_t>(1, std::min<int64_t>(tile_cols, ncol));
}

// The tile pipeline.  h_out != NULL: tiles are copied straight into the caller's G x ncol (x S) array (planes G*ncol
// apart); else into the library's pinned tiles and passed to `sink`.
static int post_tiles(bn_ctx *ctx, int kind, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                      int64_t col0, int64_t ncol, int S, uint64_t seed, int64_t tile_cols, double *h_out, bn_tile_sink sink,
                      void *user)
{
    const int planes = kind == 2 ? S : 1;
    const int64_t G = m->G;
    const int64_t tc = default_tile_cols(m, planes, ncol, tile_cols);
    const size_t tile_bytes = (size_t)G * (size_t)tc * (size_t)planes * 8;
    double *dt[2];
    BN_TRY(bn_ctx_scratch(ctx, 6, tile_bytes, (void **)&dt[0]));
    BN_TRY(bn_ctx_scratch(ctx, 7, tile_bytes, (void **)&dt[1]));
    if (!h_out && ctx->host_tile_bytes < tile_bytes) {
        for (int k = 0; k < 2; k++) {
            if (ctx->host_tile[k]) cudaFreeHost(ctx->host_tile[k]);
            ctx->host_tile[k] = nullptr;
        }
        ctx->host_tile_bytes = 0;
        for (int k = 0; k < 2; k++) {
            cudaError_t e = cudaHostAlloc((void **)&ctx->host_tile[k], tile_bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) return bn_fail(ctx, BN_ERR_OOM, "pinned tile of %zu bytes: %s", tile_bytes, cudaGetErrorString(e));
        }
        ctx->host_tile_bytes = tile_bytes;
    }
    const int64_t nt = (ncol + tc - 1) / tc;
    int rc = BN_OK;
    int64_t prev_c0 = 0, prev_nc = 0;
    for (int64_t t = 0; t < nt && rc == BN_OK; t++) {
        const int slot = (int)(t & 1);
        const int64_t c0 = t * tc, nc = std::min<int64_t>(tc, ncol - c0);
        // the device tile is free once the copy of tile t - 2 has left it
        if (t >= 2) BN_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[slot], 0));
        rc = bn_post_device(ctx, kind, m, d_beta, d_size, d_mu, col0 + c0, nc, S, seed, dt[slot]);
        if (rc != BN_OK) break;
        BN_CUDA(ctx, cudaEventRecord(ctx->ev_comp[slot], ctx->stream));
        BN_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_comp[slot], 0));
        if (h_out) {
            // plane s of the tile (G*nc values) lands at column c0 of plane s of the result (planes G*ncol apart)
            BN_CUDA(ctx, cudaMemcpy2DAsync(h_out + G * c0, (size_t)G * (size_t)ncol * 8, dt[slot], (size_t)G * (size_t)nc * 8,
                                           (size_t)G * (size_t)nc * 8, (size_t)planes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        } else {
            BN_CUDA(ctx, cudaMemcpyAsync(ctx->host_tile[slot], dt[slot], (size_t)G * (size_t)nc * (size_t)planes * 8,
                                         cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        BN_CUDA(ctx, cudaEventRecord(ctx->ev_copy[slot], ctx->copy_stream));
        if (sink && t >= 1) { // tile t - 1 is (about to be) complete: hand it over while tile t is computed and copied
            BN_CUDA(ctx, cudaEventSynchronize(ctx->ev_copy[slot ^ 1]));
            if (sink(user, col0 + prev_c0, prev_nc, ctx->host_tile[slot ^ 1]) != 0) rc = bn_fail(ctx, BN_ERR_INVALID, "tile sink stopped the iteration");
        }
        prev_c0 = c0;
        prev_nc = nc;
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
    if (rc != BN_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return bn_fail(ctx, BN_ERR_CUDA, "result tiles: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (sink && nt >= 1 && sink(user, col0 + prev_c0, prev_nc, ctx->host_tile[(nt - 1) & 1]) != 0)
        return bn_fail(ctx, BN_ERR_INVALID, "tile sink stopped the iteration");
    return BN_OK;
}

int bn_post_to_host(bn_ctx *ctx, int kind, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                    int64_t col0, int64_t ncol, int S, uint64_t seed, double *h_out)
{
    return post_tiles(ctx, kind, m, d_beta, d_size, d_mu, col0, ncol, S, seed, 0, h_out, nullptr, nullptr);
}

extern "C" int bn_post_stream(bn_ctx *ctx, const bn_mat *m, int kind, const double *BETA_vec, const double *size,
                              const double *mu, int S, uint64_t seed, int64_t tile_cols, bn_tile_sink sink, void *user)
{
    if (!ctx || !m || !BETA_vec || !size || !mu || !sink || kind < 0 || kind > 2)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_post_stream: bad arguments") : BN_ERR_INVALID;
    if (kind == 2 && S <= 0) return bn_fail(ctx, BN_ERR_INVALID, "posterior: S must be positive");
    cudaSetDevice(ctx->device);
    DevIn<double> b, sz, mm;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(sz.bind(ctx, size, (size_t)m->G));
    BN_TRY(mm.bind(ctx, mu, (size_t)m->G));
    return post_tiles(ctx, kind, m, b.p, sz.p, mm.p, 0, m->C, S, seed, tile_cols, nullptr, sink, user);
}

// ---------------------------------------------------------------------------------
// sparse result: count the non-zeros of every column of a dense tile, scan, fill
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tile_count(int64_t G, int64_t nc, const double *__restrict__ A, int64_t *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < nc; j += nwarp) {
        const double *col = A + G * j;
        int c = 0;
        for (int64_t i = lane; i < G; i += 32) c += (col[i] != 0.0);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) cnt[j] = c;
    }
}

__global__ void __launch_bounds__(256) k_tile_fill(int64_t G, int64_t nc, const double *__restrict__ A, const int64_t *__restrict__ colptr,
                                                   int32_t *__restrict__ rowidx, double *__restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < nc; j += nwarp) {
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

__global__ void k_narrow_colptr(const int64_t *__restrict__ p64, int32_t *__restrict__ p32, int64_t n)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) p32[k] = (int32_t)p64[k];
}

extern "C" void bn_csc_result_free(bn_csc_result *r)
{
    if (!r) return;
    if (r->ctx) {
        cudaSetDevice(r->ctx->device);
        if (r->colptr) cudaFreeAsync(r->colptr, r->ctx->stream);
        if (r->rowidx) cudaFreeAsync(r->rowidx, r->ctx->stream);
        if (r->val) cudaFreeAsync(r->val, r->ctx->stream);
    }
    delete r;
}

extern "C" int bn_post_mean_csc(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                                int64_t col0, int64_t ncol, bn_csc_result **out)
{
    if (!ctx || !m || !BETA_vec || !size || !mu || !out)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_post_mean_csc: NULL argument") : BN_ERR_INVALID;
    *out = nullptr;
    if (col0 < 0 || ncol <= 0 || col0 + ncol > m->C) return bn_fail(ctx, BN_ERR_INVALID, "posterior: column block outside the matrix");
    cudaSetDevice(ctx->device);
    DevIn<double> b, sz, mm;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(sz.bind(ctx, size, (size_t)m->G));
    BN_TRY(mm.bind(ctx, mu, (size_t)m->G));
    const int64_t G = m->G;
    const int64_t tc = default_tile_cols(m, 1, ncol, 0);
    double *dt = nullptr;
    BN_TRY(bn_ctx_scratch(ctx, 6, (size_t)G * (size_t)tc * 8, (void **)&dt));
    bn_csc_result *r = new bn_csc_result();
    r->ctx = ctx;
    r->G = G;
    r->ncol = ncol;
    auto fail = [&](int rc) {
        bn_csc_result_free(r);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaMallocAsync((void **)&r->colptr, (size_t)(ncol + 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "colptr: %s", cudaGetErrorString(e)));
    DevBuf<int64_t> cnt;
    int rc = cnt.alloc(ctx, (size_t)ncol + 1);
    if (rc) return fail(rc);
    cudaMemsetAsync(cnt.p, 0, (size_t)(ncol + 1) * 8, ctx->stream);
    // pass 1: the mean of every tile, counted; pass 2: the same tiles again, written out (the mean costs 9 B of HBM traffic
    // per element: cheaper to recompute than to keep a dense copy of a result that may not fit)
    for (int pass = 0; pass < 2; pass++) {
        for (int64_t c0 = 0; c0 < ncol; c0 += tc) {
            const int64_t nc = std::min<int64_t>(tc, ncol - c0);
            if ((rc = bn_post_device(ctx, 0, m, b.p, sz.p, mm.p, col0 + c0, nc, 1, 0, dt))) return fail(rc);
            const int blocks = (int)std::min<int64_t>((nc + 7) / 8, 148 * 16);
            if (pass == 0) k_tile_count<<<blocks, 256, 0, ctx->stream>>>(G, nc, dt, cnt.p + c0);
            else k_tile_fill<<<blocks, 256, 0, ctx->stream>>>(G, nc, dt, r->colptr + c0, r->rowidx, r->val);
            ctx->launches++;
        }
        if (pass == 0) {
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, r->colptr, (int)(ncol + 1), ctx->stream);
            DevBuf<char> tmp;
            if ((rc = tmp.alloc(ctx, tb))) return fail(rc);
            cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, r->colptr, (int)(ncol + 1), ctx->stream);
            ctx->launches++;
            if ((e = cudaMemcpyAsync(&r->nnz, r->colptr + ncol, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
                return fail(bn_fail(ctx, BN_ERR_CUDA, "sparse result scan: %s", cudaGetErrorString(e)));
            const size_t n = (size_t)(r->nnz ? r->nnz : 1);
            if ((e = cudaMallocAsync((void **)&r->rowidx, n * 4, ctx->stream)) != cudaSuccess ||
                (e = cudaMallocAsync((void **)&r->val, n * 8, ctx->stream)) != cudaSuccess)
                return fail(bn_fail(ctx, BN_ERR_OOM, "sparse result of %lld non-zeros: %s", (long long)r->nnz, cudaGetErrorString(e)));
        }
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(bn_fail(ctx, BN_ERR_CUDA, "sparse result: %s", cudaGetErrorString(e)));
    if ((rc = bn_sync(ctx))) return fail(rc);
    *out = r;
    return BN_OK;
}

extern "C" int bn_csc_result_info(const bn_csc_result *r, int64_t out[3])
{
    if (!r || !out) return BN_ERR_INVALID;
    out[0] = r->G;
    out[1] = r->ncol;
    out[2] = r->nnz;
    return BN_OK;
}

extern "C" int bn_csc_result_fetch(bn_ctx *ctx, const bn_csc_result *r, int64_t *p64, int32_t *p32, int32_t *i, double *x)
{
    if (!ctx || !r) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (p64) BN_CUDA(ctx, cudaMemcpyAsync(p64, r->colptr, (size_t)(r->ncol + 1) * 8, cudaMemcpyDefault, ctx->stream));
    DevBuf<int32_t> narrow;
    if (p32) {
        if (r->nnz > 2147483647LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "more than 2^31 - 1 non-zeros: 32-bit column pointers cannot hold them");
        BN_TRY(narrow.alloc(ctx, (size_t)r->ncol + 1));
        BN_LAUNCH(ctx, k_narrow_colptr, (unsigned)((r->ncol + 1 + 255) / 256), 256, 0, r->colptr, narrow.p, r->ncol + 1);
        BN_CUDA(ctx, cudaMemcpyAsync(p32, narrow.p, (size_t)(r->ncol + 1) * 4, cudaMemcpyDefault, ctx->stream));
    }
    if (i && r->nnz) BN_CUDA(ctx, cudaMemcpyAsync(i, r->rowidx, (size_t)r->nnz * 4, cudaMemcpyDefault, ctx->stream));
    if (x && r->nnz) BN_CUDA(ctx, cudaMemcpyAsync(x, r->val, (size_t)r->nnz * 8, cudaMemcpyDefault, ctx->stream));
    return bn_sync(ctx);
}

extern "C" int bn_csc_result_to_mat(bn_ctx *ctx, bn_csc_result *r, bn_mat **out)
{
    if (!ctx || !r || !out) return BN_ERR_INVALID;
    // the arrays change hands: the result handle is left empty (still to be freed by the caller)
    bn_mat *m = new bn_mat();
    m->ctx = ctx;
    m->G = r->G;
    m->C = r->ncol;
    m->nnz = r->nnz;
    m->colptr = r->colptr;
    m->rowidx = r->rowidx;
    m->val = r->val;
    m->is_count = false; // a posterior mean is real-valued
    r->colptr = nullptr;
    r->rowidx = nullptr;
    r->val = nullptr;
    r->nnz = 0;
    *out = m;
    return BN_OK;
}
