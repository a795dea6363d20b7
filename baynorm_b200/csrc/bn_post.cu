// bn_post.cu -- the gene x cell posterior kernels (Main_mean_NB_Bay, Main_mode_NB_Bay,
// Main_NB_Bay; src/bayNorm_main.cpp:1063-1299), binomial DownSampling (:892-915) and the
// synthetic NB count generator the benchmarks use.
//
// Layout: every WARP owns 128 consecutive genes (one row block of the blocked column pointers) and walks a chunk of
// <= 64 consecutive cell columns without any CTA barrier; size/mu of its genes live in registers for the whole walk.
// For every column the warp scatters its slice of the column's non-zeros into a 1 KB shared-memory strip of counts,
// evaluates the closed form for its genes and streams the result out: each output element is written exactly once,
// 256 contiguous bytes per store instruction.  HBM traffic per element: 8 B written + 12 B * nnz-fraction read;
// beta/size/mu are O(G + C) and stay in L2.
//
// Bit-exactness (mode): the reference's expression order is reproduced with explicit
// round-to-nearest intrinsics, which the compiler never contracts into FMAs:
//   tempmu = ((x + size) * (mu - mu*beta)) / (size + mu*beta)          (:1118, :1196, :1278)
//   mode   = floor((tempmu / (size + x)) * ((size + x) - 1)) + x       (:1281), 0 if size+x <= 1
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include "bn_internal.cuh"
#include "bn_rng.cuh"

#define POST_NCMAX 64 // columns walked by one CTA

enum { POST_MEAN = 0, POST_MODE = 1, POST_SAMPLE = 2 };

__device__ __forceinline__ double post_tempmu(double x, double size, double mu, double beta)
{
    const double mb = __dmul_rn(mu, beta);
    return __ddiv_rn(__dmul_rn(__dadd_rn(x, size), __dsub_rn(mu, mb)), __dadd_rn(size, mb));
}

// Posterior mode, bit-exact with the reference's operation order.
//   reference:  floor( ((x+size)*(mu-mu*beta)/(size+mu*beta)) / (size+x) * ((size+x)-1) ) + x
// The two IEEE divisions dominate the FP64 cost of the kernel, and floor() only needs to know on
// which side of an integer the value lies.  Fast path: the same real-number quantity with ONE
// approximate reciprocal (rcp.approx seed + two Newton steps, relative error < 2^-50 overall);
// if an integer lies within 2^-42 * v of it, redo the reference sequence exactly.  The fast value
// and the reference value both differ from the real number by far less than that margin, so
// whenever the fast path is taken floor() of both is the same integer.
__device__ __forceinline__ double post_mode(double x, double size, double mu, double beta)
{
    const double r = __dadd_rn(size, x);
    if (!(r > 1.0)) return 0.0;
    const double mb = __dmul_rn(mu, beta);
    const double a = __dsub_rn(mu, mb), den = __dadd_rn(size, mb), rm1 = __dsub_rn(r, 1.0);
    // in real arithmetic the (size+x) factors cancel: v = (mu - mu*beta) * (size+x-1) / (size + mu*beta)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    y = fma(y, fma(-den, y, 1.0), y); // rcp.approx is good to ~2^-20: two Newton steps reach full precision
    y = fma(y, fma(-den, y, 1.0), y);
    const double v = (a * rm1) * y;
    const double fl = floor(v);
    const double frac = v - fl, tol = fabs(v) * 2.2737367544323206e-13; // 2^-42
    if (frac > tol && frac < 1.0 - tol) return fl + x;                    // (NaN/inf/huge fall through)
    const double tm = __ddiv_rn(__dmul_rn(r, a), den);
    return __dadd_rn(floor(__dmul_rn(__ddiv_rn(tm, r), rm1)), x);
}

// ---------------------------------------------------------------------------------
// Warp-autonomous posterior mean / mode: every warp owns 32 * EPL consecutive genes (EPL per lane; a store
// instruction writes 256 contiguous bytes, a column EPL x 256 B) and walks the CTA's column chunk without any
// __syncthreads().  Its slice of every column comes from the blocked column pointers of the matrix
// (bn_mat_ensure_blockptr: two coalesced loads per column instead of a binary search), so slice lengths are
// known and up to PF x 32 non-zeros are prefetched into registers one column ahead; size/mu of the warp's genes
// stay in registers.  EPL = 4 (64 registers, 4 CTAs per SM) measured best: mode on the 50 % dense config
// 0.93 -> 0.66 ms against the first, CTA-synchronous kernel (removed).
// ---------------------------------------------------------------------------------
#define PW_WARPS 8

template <int MODE, int PF, int MINB, int EPL>
__global__ void __launch_bounds__(PW_WARPS * 32, MINB)
    k_posterior_w(int64_t G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                  const double *__restrict__ val, const double *__restrict__ beta, const double *__restrict__ size,
                  const double *__restrict__ mu, int64_t col0, int64_t ncol, int nc_per_cta, double *__restrict__ out,
                  const uint32_t *__restrict__ blockptr, int64_t Ctot)
{
    constexpr int PW_GENES = 32 * EPL, PW_TG = PW_GENES * PW_WARPS;
    __shared__ double s_x[PW_WARPS][PW_GENES];               // counts of the warp's genes in the current column
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * PW_TG;
    const int64_t w0 = t0 + (int64_t)warp * PW_GENES;
    const int64_t jb = (int64_t)blockIdx.y * nc_per_cta;
    const int nc = (int)((jb + nc_per_cta < ncol ? jb + nc_per_cta : ncol) - jb);
    if (w0 >= G) return;
    double *xt = s_x[warp];
    // slice bounds of columns lane and lane + 32 from the blocked column pointers (the warp's genes are EPL / 4
    // row blocks), and their BETA
    const int64_t nblock = (G + BN_ROWBLOCK - 1) / BN_ROWBLOCK, blk_lo = w0 / BN_ROWBLOCK;
    const int64_t blk_hi = blk_lo + PW_GENES / BN_ROWBLOCK < nblock ? blk_lo + PW_GENES / BN_ROWBLOCK : nblock; // row nblock = column end
    const uint32_t *bp_lo = blockptr + blk_lo * Ctot + col0 + jb;
    const uint32_t *bp_hi = blockptr + blk_hi * Ctot + col0 + jb;
    // positions inside the chunk are 32-bit offsets from the chunk's first column (<= 64 columns of < 2^31 rows);
    // only that base is 64-bit, so a shard may hold more than 2^32 non-zeros
    const int64_t base = __ldg(colptr + col0 + jb);
    rowidx += base;
    val += base;
    uint32_t lo_a = 0, hi_a = 0, lo_b = 0, hi_b = 0;
    double beta_a = 0.0, beta_b = 0.0;
    if (lane < nc) {
        const uint32_t cs = (uint32_t)(__ldg(colptr + col0 + jb + lane) - base);
        lo_a = cs + __ldg(bp_lo + lane);
        hi_a = cs + __ldg(bp_hi + lane);
        beta_a = __ldg(beta + col0 + jb + lane);
    }
    if (lane + 32 < nc) {
        const uint32_t cs = (uint32_t)(__ldg(colptr + col0 + jb + lane + 32) - base);
        lo_b = cs + __ldg(bp_lo + lane + 32);
        hi_b = cs + __ldg(bp_hi + lane + 32);
        beta_b = __ldg(beta + col0 + jb + lane + 32);
    }
    double sz[EPL], m[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        const int64_t i = w0 + e * 32 + lane;
        sz[e] = (i < G) ? __ldg(size + i) : 1.0;
        m[e] = (i < G) ? __ldg(mu + i) : 0.0;
        xt[e * 32 + lane] = 0.0;
    }
    __syncwarp();
    const bool full = w0 + PW_GENES <= G; // warp-uniform: no per-element bounds test
    int32_t pr[PF];
    double pv[PF];
    uint32_t kcur = 0, kend = 0;
#define PW_PREFETCH(COL)                                                    \
    do {                                                                    \
        kcur = __shfl_sync(0xffffffffu, ((COL) < 32) ? lo_a : lo_b, (COL) & 31); \
        kend = __shfl_sync(0xffffffffu, ((COL) < 32) ? hi_a : hi_b, (COL) & 31); \
        _Pragma("unroll") for (int q = 0; q < PF; q++)                      \
        {                                                                   \
            const uint32_t k = kcur + q * 32 + lane;                        \
            pr[q] = (k < kend) ? __ldg(rowidx + k) : -1;                    \
            pv[q] = (k < kend) ? __ldcs(val + k) : 0.0;                     \
        }                                                                   \
    } while (0)
    if (nc > 0) PW_PREFETCH(0);
    double *po = out + G * jb + w0 + lane;
    for (int jj = 0; jj < nc; jj++) {
        // scatter the slice of column jj: the prefetched part from registers, a longer slice from memory
#pragma unroll
        for (int q = 0; q < PF; q++)
            if (pr[q] >= 0) xt[pr[q] - w0] = pv[q];
        for (uint32_t k = kcur + PF * 32 + lane; k < kend; k += 32) xt[__ldg(rowidx + k) - w0] = __ldcs(val + k);
        __syncwarp();
        const double b = __shfl_sync(0xffffffffu, (jj < 32) ? beta_a : beta_b, jj & 31);
        if (jj + 1 < nc) PW_PREFETCH(jj + 1);
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            const double xv = xt[e * 32 + lane];
            xt[e * 32 + lane] = 0.0; // own slot: ready for the next column's scatter
            const double v = (MODE == POST_MEAN) ? __dadd_rn(post_tempmu(xv, sz[e], m[e], b), xv) : post_mode(xv, sz[e], m[e], b);
            if (full || w0 + e * 32 + lane < G) __stcs(po + e * 32, v);
        }
        po += G;
        __syncwarp();
    }
#undef PW_PREFETCH
}

int bn_post_sample_device(bn_ctx *ctx, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                          int64_t col0, int64_t ncol, int S, uint64_t seed, double *d_out); // bn_sample.cu

// device pointers in and out, nothing synchronised: the core every posterior entry point shares
int bn_post_device(bn_ctx *ctx, int mode, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                   int64_t col0, int64_t ncol, int S, uint64_t seed, double *d_out)
{
    if (mode == POST_SAMPLE) return bn_post_sample_device(ctx, m, d_beta, d_size, d_mu, col0, ncol, S, seed, d_out);
    BN_TRY(bn_mat_ensure_blockptr(ctx, m));
    const int sms = ctx->prop.multiProcessorCount;
    const bool densew = (double)m->nnz > 0.2 * (double)m->G * (double)m->C;
    const int64_t PW_TG = 32 * 4 * PW_WARPS; // 4 genes per lane (128 per warp, 4 CTAs per SM); 8 per lane at 3 CTAs per SM measured slower
    const int64_t tiles_w = (m->G + PW_TG - 1) / PW_TG;
    int ncw = POST_NCMAX;
    while (ncw > 16 && tiles_w * ((ncol + ncw - 1) / ncw) < (int64_t)sms * 24) ncw >>= 1;
    if (ctx->tune[BN_TUNE_POST_NCS] > 0) ncw = std::max(1, std::min(POST_NCMAX, ctx->tune[BN_TUNE_POST_NCS]));
    if ((ncol + ncw - 1) / ncw > 65535) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: at most %d columns per call", 65535 * 16);
    dim3 gridw((unsigned)tiles_w, (unsigned)((ncol + ncw - 1) / ncw));
#define POSTW_ARGS m->G, m->colptr, m->rowidx, m->val, d_beta, d_size, d_mu, col0, ncol, ncw, d_out, m->blockptr, m->C
    // slices prefetched per column: 3 x 32 non-zeros for >= 20 % dense matrices (5 x 32 spills at 64 registers), 32 otherwise
    if (mode == POST_MEAN) {
        if (densew) BN_LAUNCH(ctx, (k_posterior_w<POST_MEAN, 3, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
        else BN_LAUNCH(ctx, (k_posterior_w<POST_MEAN, 1, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
    } else {
        if (densew) BN_LAUNCH(ctx, (k_posterior_w<POST_MODE, 3, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
        else BN_LAUNCH(ctx, (k_posterior_w<POST_MODE, 1, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
    }
#undef POSTW_ARGS
    return BN_OK;
}

static int post_launch(bn_ctx *ctx, int mode, const bn_mat *m, const double *BETA_vec, const double *size,
                       const double *mu, int64_t col0, int64_t ncol, int S, uint64_t seed, double *out)
{
    if (!ctx || !m || !BETA_vec || !size || !mu || !out)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "posterior: NULL argument (mu is required: the NB model has no default)") : BN_ERR_INVALID;
    if (col0 < 0 || ncol <= 0 || col0 + ncol > m->C) return bn_fail(ctx, BN_ERR_INVALID, "posterior: column block [%lld,+%lld) outside 0..%lld", (long long)col0, (long long)ncol, (long long)m->C);
    if (mode == POST_SAMPLE && S <= 0) return bn_fail(ctx, BN_ERR_INVALID, "posterior: S must be positive");
    cudaSetDevice(ctx->device);
    DevIn<double> b, sz, mm;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(sz.bind(ctx, size, (size_t)m->G));
    BN_TRY(mm.bind(ctx, mu, (size_t)m->G));
    if (bn_is_device_ptr(out)) {
        BN_TRY(bn_post_device(ctx, mode, m, b.p, sz.p, mm.p, col0, ncol, S, seed, out));
        return bn_sync(ctx);
    }
    // host destination: column tiles through two device buffers, tile t+1 computed while tile t is on the bus
    return bn_post_to_host(ctx, mode, m, b.p, sz.p, mm.p, col0, ncol, S, seed, out);
}

extern "C" int bn_post_mean(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                            int64_t col0, int64_t ncol, double *out)
{
    return post_launch(ctx, POST_MEAN, m, BETA_vec, size, mu, col0, ncol, 1, 0, out);
}
extern "C" int bn_post_mode(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                            int64_t col0, int64_t ncol, double *out)
{
    return post_launch(ctx, POST_MODE, m, BETA_vec, size, mu, col0, ncol, 1, 0, out);
}
extern "C" int bn_post_sample(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                              int64_t col0, int64_t ncol, int S, uint64_t seed, double *out)
{
    return post_launch(ctx, POST_SAMPLE, m, BETA_vec, size, mu, col0, ncol, S, seed, out);
}

// ---------------------------------------------------------------------------------
// DownSampling (src/bayNorm_main.cpp:892-915): dense column-major in/out, out[i,j] ~ Binomial(Data[i,j], beta_j),
// 16 B/element of HBM traffic.  Four genes of a cell share ONE Philox block -- counter (gene / 4, cell), word =
// gene % 4 -- so a draw is a pure function of (seed, gene, cell) and costs a quarter of a block; zeros cost nothing.
// Counts with n min(p, 1-p) < 30 are inverted from 0 in float; larger ones (and counts beyond float's integer grid)
// are listed and drawn by k_downsample_big with the FP64 BTRS path of bn_rbinom on a sub-stream of their own.
// ---------------------------------------------------------------------------------
// the float path: sequential inversion from 0 with p <= 1/2:  f(0) = (1-p)^n,  f(k) = f(k-1) s (n-k+1)/k,  s = p/(1-p)
__device__ __forceinline__ double ds_draw_small(double n, float pf, float L, float sratio, bool flip, uint32_t w)
{
    const float nf = (float)n;
    float f = __expf(nf * L), u = bn_u01f(w), k = 0.0f;
    while (u > f && k < nf) {
        u -= f;
        k += 1.0f;
        f *= sratio * __fdividef(nf - k + 1.0f, k);
    }
    return flip ? n - (double)k : (double)k;
}
__device__ __forceinline__ bool ds_is_big(double n, float pf) { return (float)n * pf >= 30.0f || n >= 16777216.0; }

// A warp owns 32-gene x 32-cell tiles.  Loads and stores run with lane = gene (every instruction 256 contiguous
// bytes); the draws run with lane = CELL through a padded shared-memory tile, one gene at a time, so the 32 lanes of a
// draw instruction hold counts of the same gene (same scale -> similar inversion lengths) and each lane keeps its
// column's p, log1p(-p) and p/q for the whole tile.  Measured against lane = gene for the draws: 15 -> 2x active
// lanes per instruction.
#define DS_WARPS 8
__global__ void __launch_bounds__(DS_WARPS * 32, 3)
    k_downsample(int64_t G, int64_t C, const double *__restrict__ in, const double *__restrict__ beta,
                 const __grid_constant__ PhiloxKeys keys, double *__restrict__ out, int64_t *__restrict__ biglist,
                 unsigned long long *__restrict__ bigcount)
{
    extern __shared__ double s_t[]; // [warp][cell][gene], rows padded to 33
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *T = s_t + wib * (32 * 33);
    const int64_t tg = (G + 31) / 32, tc = (C + 31) / 32, total = tg * tc;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t t = warp; t < total; t += nwarp) {
        const int64_t ct = t / tg, gt = t - ct * tg;
        const int64_t g0 = gt * 32, c0 = ct * 32;
        const int ncell = (int)(C - c0 < 32 ? C - c0 : 32);
        const bool gin = g0 + lane < G;
        // load: lane = gene
#pragma unroll 8
        for (int cc = 0; cc < 32; cc++) T[cc * 33 + lane] = (gin && cc < ncell) ? __ldcs(in + (c0 + cc) * G + g0 + lane) : 0.0;
        __syncwarp();
        // draw: lane = cell
        const int64_t j = c0 + lane;
        const double pp = (lane < ncell) ? __ldg(beta + j) : 0.0;
        const bool flip = pp > 0.5;
        const float pf = (float)(flip ? 1.0 - pp : pp);
        const float L = log1pf(-pf), sratio = pf / (1.0f - pf);
        for (int gi = 0; gi < 32; gi += 4) {
            double n[4];
#pragma unroll
            for (int q = 0; q < 4; q++) n[q] = T[lane * 33 + gi + q];
            double r[4] = {0.0, 0.0, 0.0, 0.0};
            int nbig = 0;
            if ((n[0] > 0.0 || n[1] > 0.0 || n[2] > 0.0 || n[3] > 0.0) && pp > 0.0) {
                if (pp >= 1.0) {
#pragma unroll
                    for (int q = 0; q < 4; q++) r[q] = n[q];
                } else {
                    uint32_t wd[4]; // one Philox block per (4 genes, cell): counter (gene / 4, cell), word = gene % 4
                    const uint64_t gq = (uint64_t)(g0 + gi) >> 2;
                    bn_philox_block(keys, (uint32_t)gq, (uint32_t)(gq >> 32), (uint32_t)j, (uint32_t)((uint64_t)j >> 32), wd[0], wd[1], wd[2],
                                    wd[3]);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (!(n[q] > 0.0)) continue;
                        if (ds_is_big(n[q], pf)) nbig++; // drawn by k_downsample_big: BTRS never shares a warp with the inversion
                        else r[q] = ds_draw_small(n[q], pf, L, sratio, flip, wd[q]);
                    }
                }
            }
            // warp-aggregated append of the big elements (element index = i + G j)
            const unsigned any = __ballot_sync(0xffffffffu, nbig > 0);
            if (any) {
                int pre = nbig;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, pre, o);
                    if (lane >= o) pre += v;
                }
                unsigned long long base = 0;
                if (lane == 31) base = atomicAdd(bigcount, (unsigned long long)pre);
                base = __shfl_sync(0xffffffffu, base, 31) + (unsigned long long)(pre - nbig);
                if (nbig > 0) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (n[q] > 0.0 && ds_is_big(n[q], pf)) biglist[base++] = j * G + g0 + gi + q;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) T[lane * 33 + gi + q] = r[q];
        }
        __syncwarp();
        // store: lane = gene
        if (gin) {
#pragma unroll 8
            for (int cc = 0; cc < 32; cc++)
                if (cc < ncell) __stcs(out + (c0 + cc) * G + g0 + lane, T[cc * 33 + lane]);
        }
        __syncwarp();
    }
}

// counts with n min(p, 1-p) >= 30 (or beyond float's integer grid): Hoermann's BTRS in FP64, one element per thread
__global__ void __launch_bounds__(256) k_downsample_big(int64_t G, const double *__restrict__ in, const double *__restrict__ beta,
                                                        uint64_t seed, const int64_t *__restrict__ biglist,
                                                        const unsigned long long *__restrict__ bigcount, double *__restrict__ out)
{
    const int64_t n = (int64_t)*bigcount;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += gridDim.x * (int64_t)blockDim.x) {
        const int64_t e = biglist[t], j = e / G, i = e - j * G;
        Philox g;
        g.init(seed, (uint64_t)i + (uint64_t)j * 0x100000000ULL, 0x44533032u /* "DS02" */);
        out[e] = bn_rbinom(g, in[e], __ldg(beta + j));
    }
}

extern "C" int bn_downsample(bn_ctx *ctx, int64_t G, int64_t C, const double *Data, const double *BETA_vec,
                             uint64_t seed, double *out)
{
    if (!ctx || !Data || !BETA_vec || !out || G <= 0 || C <= 0)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_downsample: bad arguments") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> a, b;
    BN_TRY(a.bind(ctx, Data, (size_t)G * C));
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)C));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, out, (size_t)G * C));
    const int64_t want = (((G + 31) / 32) * ((C + 31) / 32) + DS_WARPS - 1) / DS_WARPS; // one warp per 32 x 32 tile
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->prop.multiProcessorCount * 24);
    // elements for the BTRS kernel: at most every element (8 B each, context scratch)
    int64_t *biglist = nullptr;
    BN_TRY(bn_ctx_scratch(ctx, 5, (size_t)G * (size_t)C * 8, (void **)&biglist));
    DevBuf<unsigned long long> bigcount;
    BN_TRY(bigcount.alloc(ctx, 1));
    BN_CUDA(ctx, cudaMemsetAsync(bigcount.p, 0, 8, ctx->stream));
    const size_t ds_smem = (size_t)DS_WARPS * 32 * 33 * sizeof(double);
    BN_CUDA(ctx, cudaFuncSetAttribute(k_downsample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ds_smem));
    BN_LAUNCH(ctx, k_downsample, grid, DS_WARPS * 32, ds_smem, G, C, a.p, b.p, bn_philox_keys(seed), o.p, biglist, bigcount.p);
    BN_LAUNCH(ctx, k_downsample_big, ctx->prop.multiProcessorCount * 8, 256, 0, G, a.p, b.p, seed, biglist, bigcount.p, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

// ---------------------------------------------------------------------------------
// synthetic counts: x_ij ~ NB(size phi_i, mean mu_i beta_j), written straight to CSC.
// Two passes with the same counter-based draws: count per column, scan, fill in row order.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double synth_draw(int64_t gi, int64_t j, double phi, double mu, double beta, uint64_t seed)
{
    const double M = mu * beta;
    const NBParams P = bn_nb_setup(phi, M);
    Philox g;
    g.init(seed, (uint64_t)gi + (uint64_t)j * 0x100000000ULL, 0x53594e30u /* "SYN0" */);
    return bn_nb_draw(P, g);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_synth(int64_t G, int64_t C, const double *__restrict__ mu,
                                               const double *__restrict__ size, const double *__restrict__ beta,
                                               uint64_t seed, int64_t gene0, int64_t *__restrict__ cnt,
                                               const int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                                               double *__restrict__ val)
{
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    for (int64_t j = blockIdx.x; j < C; j += gridDim.x) {
        const double b = beta[j];
        int64_t base = FILL ? colptr[j] : 0;
        int total = 0;
        for (int64_t i0 = 0; i0 < G; i0 += 256) {
            const int64_t i = i0 + threadIdx.x;
            double x = 0.0;
            if (i < G) x = synth_draw(gene0 + i, j, size[i], mu[i], b, seed);
            const int f = (x > 0.0);
            if (FILL) {
                int pos, agg;
                Scan(tmp).ExclusiveSum(f, pos, agg);
                if (f) {
                    rowidx[base + pos] = (int32_t)i;
                    val[base + pos] = x;
                }
                base += agg;
                __syncthreads();
            } else
                total += f;
        }
        if (!FILL) {
            int incl;
            Scan(tmp).InclusiveSum(total, incl);
            if (threadIdx.x == 255) cnt[j] = incl;
            __syncthreads();
        }
    }
}

extern "C" int bn_mat_synth_nb(bn_ctx *ctx, int64_t G, int64_t C, const double *mu, const double *size,
                               const double *beta, uint64_t seed, int64_t gene0, bn_mat **out)
{
    if (!ctx || !out || !mu || !size || !beta || G <= 0 || C <= 0)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_mat_synth_nb: bad arguments") : BN_ERR_INVALID;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    DevIn<double> dmu, dsz, db;
    BN_TRY(dmu.bind(ctx, mu, (size_t)G));
    BN_TRY(dsz.bind(ctx, size, (size_t)G));
    BN_TRY(db.bind(ctx, beta, (size_t)C));
    bn_mat *m = new bn_mat();
    m->ctx = ctx;
    m->G = G;
    m->C = C;
    m->gene0 = gene0;
    m->is_count = true;
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
    const int grid = (int)std::min<int64_t>(C, 148 * 8);
    k_synth<false><<<grid, 256, 0, ctx->stream>>>(G, C, dmu.p, dsz.p, db.p, seed, gene0, cnt.p, nullptr, nullptr, nullptr);
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
        return fail(bn_fail(ctx, BN_ERR_CUDA, "synth scan: %s", cudaGetErrorString(e)));
    m->nnz = nnz;
    if ((e = cudaMallocAsync((void **)&m->rowidx, (size_t)(nnz ? nnz : 1) * 4, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&m->val, (size_t)(nnz ? nnz : 1) * 8, ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_OOM, "matrix allocation failed: %s", cudaGetErrorString(e)));
    k_synth<true><<<grid, 256, 0, ctx->stream>>>(G, C, dmu.p, dsz.p, db.p, seed, gene0, nullptr, m->colptr, m->rowidx, m->val);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(bn_fail(ctx, BN_ERR_CUDA, "k_synth: %s", cudaGetErrorString(e)));
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        return fail(bn_fail(ctx, BN_ERR_CUDA, "k_synth: %s", cudaGetErrorString(e)));
    *out = m;
    return BN_OK;
}

// ---------------------------------------------------------------------------------
// debug hook: raw Philox4x32-10 block, so tests can pin the generator against the
// published known-answer vectors (Random123 kat_vectors).
// ---------------------------------------------------------------------------------
__global__ void k_philox_block(const uint32_t *__restrict__ ctr, const uint32_t *__restrict__ key, uint32_t *__restrict__ out)
{
    Philox g;
    g.k0 = key[0];
    g.k1 = key[1];
    g.c0 = ctr[0];
    g.c1 = ctr[1];
    g.c2 = ctr[2];
    g.c3 = ctr[3];
    g.have = 0;
    g.refill();
    out[0] = g.out[0];
    out[1] = g.out[1];
    out[2] = g.out[2];
    out[3] = g.out[3];
}

extern "C" int bn_debug_philox(bn_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (!ctx || !ctr || !key || !out) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<uint32_t> c, k;
    BN_TRY(c.bind(ctx, ctr, 4));
    BN_TRY(k.bind(ctx, key, 2));
    DevOut<uint32_t> o;
    BN_TRY(o.bind(ctx, out, 4));
    BN_LAUNCH(ctx, k_philox_block, 1, 1, 0, c.p, k.p, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}
