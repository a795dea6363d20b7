// bn_reduce.cu -- the reductions around the fit: library-size column sums and the
// default BETA (R/bayNorm.r:164-170), method-of-moments prior (rowMeansFast / rowVarsFast
// / EstPrior_sprcpp, src/bayNorm_main.cpp:1303-1394 with the x/beta normalisation of
// R/PRIOR_FUNCTIONS.R:253 fused in), and AdjustSIZE_fun (R/sup_funs.R:27-33).
//
// All of these are HBM streams over the non-zeros (12 B per non-zero per pass) followed
// by O(G) or O(C) epilogues; every reduction has a fixed order so results do not depend
// on launch geometry, and cross-shard pieces go through bn_allreduce().
#include <math.h>

#include "bn_internal.cuh"
#include "bn_rng.cuh"

// ---------------------------------------------------------------------------------
// column sums: one warp per cell column, 8 columns per 256-thread CTA
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_colsums(int64_t C, const int64_t *__restrict__ colptr,
                                                 const double *__restrict__ val, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        double s0 = 0.0, s1 = 0.0;
        int64_t k = a + lane;
        for (; k + 32 < b; k += 64) {
            s0 += __ldcs(val + k);
            s1 += __ldcs(val + k + 32);
        }
        if (k < b) s0 += __ldcs(val + k);
        const double s = bn_warp_sum(s0 + s1);
        if (lane == 0) out[j] = s;
    }
}

extern "C" int bn_colsums(bn_ctx *ctx, const bn_mat *m, double *colsums)
{
    if (!ctx || !m || !colsums) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_colsums: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevOut<double> o;
    BN_TRY(o.bind(ctx, colsums, (size_t)m->C));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
    BN_LAUNCH(ctx, k_colsums, blocks, 256, 0, m->C, m->colptr, m->val, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

// ---------------------------------------------------------------------------------
// one-CTA block reductions (fixed order: thread-strided partials, then a tree)
// ---------------------------------------------------------------------------------
enum { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };

template <int OP> __device__ __forceinline__ double red_op(double a, double b)
{
    return OP == RED_SUM ? a + b : (OP == RED_MIN ? fmin(a, b) : fmax(a, b));
}
template <int OP> __device__ double block_reduce_1024(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = sh[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) r = red_op<OP>(r, sh[k]);
    return r;
}

// beta = xx/mean(xx)*mean_beta; beta<=0 -> min positive; beta>=1 -> max below 1
// (R/bayNorm.r:166-169).  status[0] = 1 when the range check of :185-187 fails.
__global__ void __launch_bounds__(1024) k_beta_from_colsums(int64_t C, const double *__restrict__ xx, double mean_beta,
                                                            double *__restrict__ beta, int *__restrict__ status)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) s += xx[j];
    const double mean = block_reduce_1024<RED_SUM>(s, sh) / (double)C;
    double mn = INFINITY;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        const double b = __dmul_rn(__ddiv_rn(xx[j], mean), mean_beta);
        beta[j] = b;
        if (b > 0.0) mn = fmin(mn, b);
    }
    const double minpos = block_reduce_1024<RED_MIN>(mn, sh);
    double mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        double b = beta[j];
        if (b <= 0.0) {
            b = minpos;
            beta[j] = b;
        }
        if (b < 1.0) mx = fmax(mx, b);
    }
    const double maxlt1 = block_reduce_1024<RED_MAX>(mx, sh);
    double lo = INFINITY, hi = -INFINITY;
    int nan = 0;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        double b = beta[j];
        if (b >= 1.0) {
            b = maxlt1;
            beta[j] = b;
        }
        if (b != b) nan = 1;
        lo = fmin(lo, b);
        hi = fmax(hi, b);
    }
    lo = block_reduce_1024<RED_MIN>(lo, sh);
    hi = block_reduce_1024<RED_MAX>(hi, sh);
    nan = __syncthreads_or(nan);
    if (threadIdx.x == 0) status[0] = (nan || !(lo > 0.0) || !(hi < 1.0)) ? 1 : 0;
}

extern "C" int bn_beta_default(bn_ctx *ctx, const bn_mat *m, double mean_beta, double *beta)
{
    if (!ctx || !m || !beta) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_beta_default: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevOut<double> o;
    BN_TRY(o.bind(ctx, beta, (size_t)m->C));
    DevBuf<double> xx;
    BN_TRY(xx.alloc(ctx, (size_t)m->C));
    DevBuf<int> st;
    BN_TRY(st.alloc(ctx, 1));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
    BN_LAUNCH(ctx, k_colsums, blocks, 256, 0, m->C, m->colptr, m->val, xx.p);
    BN_TRY(bn_allreduce(ctx, xx.p, m->C, 0)); // per-cell totals over gene shards
    BN_LAUNCH(ctx, k_beta_from_colsums, 1, 1024, 0, m->C, xx.p, mean_beta, o.p, st.p);
    int h = 0;
    BN_CUDA(ctx, cudaMemcpyAsync(&h, st.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(o.commit(ctx));
    BN_TRY(bn_sync(ctx));
    if (h) return bn_fail(ctx, BN_ERR_BETA_RANGE, "The range of BETA must be within (0,1).");
    return BN_OK;
}

template <bool PACKED>
__device__ __forceinline__ void nz_load(const uint64_t *__restrict__ gm, const int32_t *__restrict__ cell,
                                        const double *__restrict__ rval, int64_t k, int &c, double &x)
{
    if (PACKED) {
        const uint64_t e = __ldcs(gm + k);
        c = (int)(e >> 32);
        x = (double)(uint32_t)e;
    } else {
        c = cell[k];
        x = rval[k];
    }
}

// ---------------------------------------------------------------------------------
// per-gene moments on the gene-major twin: one warp per gene.
// mode 0: full EstPrior (MU, SIZE with NaN rule, V)     -- means computed here
// mode 1: rowMeansFast only (MU)
// mode 2: rowVarsFast with caller-supplied means (VARS = raw variance, /(C-1))
// mode 3: as mode 0 but SIZE is the raw quotient m^2/(v-m) without the R-level NaN rule
//         (what EstPrior_sprcpp / EstPrior_rcpp themselves return)
// ---------------------------------------------------------------------------------
// TPG = 32: warp per gene, genes below `split` stored entries; TPG = 256: CTA per gene, the long ones (a 1.3 M-entry gene on a
// single warp took 30 of the 41 ms this kernel needed on config 5).  Either way ONE summation order per gene.
template <bool PACKED, int TPG>
__global__ void __launch_bounds__(256) k_row_moments(int mode, int64_t split, int64_t G, int64_t C, const int64_t *__restrict__ rowptr,
                                                     const uint64_t *__restrict__ gm, const int32_t *__restrict__ cell,
                                                     const double *__restrict__ rval,
                                                     const double *__restrict__ beta, const double *__restrict__ means_in,
                                                     double *__restrict__ MU, double *__restrict__ SIZE,
                                                     double *__restrict__ V)
{
    __shared__ double scratch[8];
    const int lane = TPG == 32 ? (threadIdx.x & 31) : threadIdx.x; // position inside the gene's group
    const int64_t warp = TPG == 32 ? (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5 : blockIdx.x;
    const int64_t nwarp = TPG == 32 ? (gridDim.x * (int64_t)blockDim.x) >> 5 : gridDim.x;
    const double n = (double)C;
    for (int64_t g = warp; g < G; g += nwarp) {
        const int64_t a = rowptr[g], b = rowptr[g + 1];
        if ((b - a >= split) != (TPG != 32)) continue; // the other kernel's gene (uniform over the group)
        double mean;
        if (mode == 2) {
            mean = means_in[g];
        } else {
            // four independent loads / divisions / partial sums per lane: the loop is latency-bound otherwise
            double s = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t k = a + lane;
            for (; k + 3 * TPG < b; k += 4 * TPG) {
                double x0, x1, x2, x3;
                int c0, c1, c2, c3;
                nz_load<PACKED>(gm, cell, rval, k, c0, x0);
                nz_load<PACKED>(gm, cell, rval, k + TPG, c1, x1);
                nz_load<PACKED>(gm, cell, rval, k + 2 * TPG, c2, x2);
                nz_load<PACKED>(gm, cell, rval, k + 3 * TPG, c3, x3);
                if (beta) {
                    const double b0 = __ldg(beta + c0), b1 = __ldg(beta + c1), b2 = __ldg(beta + c2), b3 = __ldg(beta + c3);
                    x0 = __ddiv_rn(x0, b0), x1 = __ddiv_rn(x1, b1), x2 = __ddiv_rn(x2, b2), x3 = __ddiv_rn(x3, b3);
                }
                s += x0, s1 += x1, s2 += x2, s3 += x3;
            }
            for (; k < b; k += TPG) {
                double x;
                int c;
                nz_load<PACKED>(gm, cell, rval, k, c, x);
                s += beta ? __ddiv_rn(x, __ldg(beta + c)) : x;
            }
            s = (s + s1) + (s2 + s3);
            mean = __ddiv_rn(bn_group_sum<TPG>(s, scratch), n); // :1316-1318
            if (lane == 0 && MU) MU[g] = mean;
            if (mode == 1) continue;
        }
        double ss = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
        int64_t k = a + lane;
        for (; k + 3 * TPG < b; k += 4 * TPG) {
            double x0, x1, x2, x3;
            int c0, c1, c2, c3;
            nz_load<PACKED>(gm, cell, rval, k, c0, x0);
            nz_load<PACKED>(gm, cell, rval, k + TPG, c1, x1);
            nz_load<PACKED>(gm, cell, rval, k + 2 * TPG, c2, x2);
            nz_load<PACKED>(gm, cell, rval, k + 3 * TPG, c3, x3);
            if (beta) {
                const double b0 = __ldg(beta + c0), b1 = __ldg(beta + c1), b2 = __ldg(beta + c2), b3 = __ldg(beta + c3);
                x0 = __ddiv_rn(x0, b0), x1 = __ddiv_rn(x1, b1), x2 = __ddiv_rn(x2, b2), x3 = __ddiv_rn(x3, b3);
            }
            const double d0 = __dsub_rn(x0, mean), d1 = __dsub_rn(x1, mean), d2 = __dsub_rn(x2, mean), d3 = __dsub_rn(x3, mean);
            ss = __dadd_rn(ss, __dmul_rn(d0, d0)), ss1 = __dadd_rn(ss1, __dmul_rn(d1, d1));
            ss2 = __dadd_rn(ss2, __dmul_rn(d2, d2)), ss3 = __dadd_rn(ss3, __dmul_rn(d3, d3));
        }
        for (; k < b; k += TPG) {
            double x;
            int c;
            nz_load<PACKED>(gm, cell, rval, k, c, x);
            const double v = beta ? __ddiv_rn(x, __ldg(beta + c)) : x;
            const double d = __dsub_rn(v, mean);
            ss = __dadd_rn(ss, __dmul_rn(d, d)); // :1331
        }
        ss = bn_group_sum<TPG>((ss + ss1) + (ss2 + ss3), scratch);
        if (lane == 0) {
            const double nz = (double)(b - a);
            double var = __dadd_rn(ss, __dmul_rn(__dsub_rn(n, nz), __dmul_rn(mean, mean))); // :1338
            var = __ddiv_rn(var, (double)(C - 1));                                            // :1339
            if (mode == 2) {
                V[g] = var;
            } else {
                const double v = __dmul_rn(__ddiv_rn(__dsub_rn(n, 1.0), n), var);         // :1385
                double size = __ddiv_rn(__dmul_rn(mean, mean), __dsub_rn(v, mean));       // :1386
                if (mode == 0 && v <= mean) size = nan("");                               // R/PRIOR_FUNCTIONS.R:163
                if (SIZE) SIZE[g] = size;
                if (V) V[g] = v;
            }
        }
    }
}

static int launch_row_moments(bn_ctx *ctx, bn_mat *m, int mode, const double *beta, const double *means_in, double *MU,
                              double *SIZE, double *V)
{
    BN_TRY(bn_mat_ensure_csr(ctx, m));
    const int blocks = (int)std::min<int64_t>((m->G + 7) / 8, 148 * 16);
    const int64_t split = 16384; // stored entries from which a gene gets a CTA
    const bool any_long = m->nnz >= split; // no gene can reach the split otherwise
    const int blocks_cta = (int)std::min<int64_t>(m->G, 148 * 8);
#define RM_ARGS mode, split, m->G, m->C, m->rowptr, m->gm, m->cell, m->rval, beta, means_in, MU, SIZE, V
    if (m->is_count) {
        if (any_long) BN_LAUNCH(ctx, (k_row_moments<true, 256>), blocks_cta, 256, 0, RM_ARGS);
        BN_LAUNCH(ctx, (k_row_moments<true, 32>), blocks, 256, 0, RM_ARGS);
    } else {
        if (any_long) BN_LAUNCH(ctx, (k_row_moments<false, 256>), blocks_cta, 256, 0, RM_ARGS);
        BN_LAUNCH(ctx, (k_row_moments<false, 32>), blocks, 256, 0, RM_ARGS);
    }
#undef RM_ARGS
    return BN_OK;
}

// device-pointer version used by bn_prior
int bn_mme_device(bn_ctx *ctx, bn_mat *m, const double *d_beta, double *d_MU, double *d_SIZE, double *d_V)
{
    return launch_row_moments(ctx, m, 0, d_beta, nullptr, d_MU, d_SIZE, d_V);
}

static int bn_mme_impl(bn_ctx *ctx, bn_mat *m, int mode, const double *beta, double *MU, double *SIZE, double *V)
{
    if (!ctx || !m) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_mme: NULL argument") : BN_ERR_INVALID;
    if (m->C < 2) return bn_fail(ctx, BN_ERR_INVALID, "bn_mme needs at least 2 cells");
    cudaSetDevice(ctx->device);
    DevIn<double> b;
    BN_TRY(b.bind(ctx, beta, (size_t)m->C));
    DevOut<double> oMU, oS, oV;
    BN_TRY(oMU.bind(ctx, MU, (size_t)m->G));
    BN_TRY(oS.bind(ctx, SIZE, (size_t)m->G));
    BN_TRY(oV.bind(ctx, V, (size_t)m->G));
    BN_TRY(launch_row_moments(ctx, m, mode, b.p, nullptr, oMU.p, oS.p, oV.p));
    BN_TRY(oMU.commit(ctx));
    BN_TRY(oS.commit(ctx));
    BN_TRY(oV.commit(ctx));
    return bn_sync(ctx);
}

extern "C" int bn_mme(bn_ctx *ctx, bn_mat *m, const double *beta, double *MU, double *SIZE, double *V)
{
    return bn_mme_impl(ctx, m, 0, beta, MU, SIZE, V);
}

extern "C" int bn_mme_raw(bn_ctx *ctx, bn_mat *m, const double *beta, double *MU, double *SIZE, double *V)
{
    return bn_mme_impl(ctx, m, 3, beta, MU, SIZE, V);
}

extern "C" int bn_row_means(bn_ctx *ctx, bn_mat *m, double *means)
{
    if (!ctx || !m || !means) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_row_means: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevOut<double> o;
    BN_TRY(o.bind(ctx, means, (size_t)m->G));
    BN_TRY(launch_row_moments(ctx, m, 1, nullptr, nullptr, o.p, nullptr, nullptr));
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

extern "C" int bn_row_vars(bn_ctx *ctx, bn_mat *m, const double *means, double *vars)
{
    if (!ctx || !m || !means || !vars) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_row_vars: NULL argument") : BN_ERR_INVALID;
    if (m->C < 2) return bn_fail(ctx, BN_ERR_INVALID, "bn_row_vars needs at least 2 cells");
    cudaSetDevice(ctx->device);
    DevIn<double> mi;
    BN_TRY(mi.bind(ctx, means, (size_t)m->G));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, vars, (size_t)m->G));
    BN_TRY(launch_row_moments(ctx, m, 2, nullptr, mi.p, nullptr, nullptr, o.p));
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

extern "C" int bn_mme_dense(bn_ctx *ctx, int64_t G, int64_t C, const double *dense, double *MU, double *SIZE, double *V)
{
    if (!ctx) return BN_ERR_INVALID;
    bn_mat *m = nullptr;
    BN_TRY(bn_mat_from_dense(ctx, G, C, dense, &m));
    // EstPrior_rcpp returns pow(m,2)/(v-m) as is; the NaN rule lives in R's EstPrior only.
    int rc = bn_mme_impl(ctx, m, 3, nullptr, MU, SIZE, V);
    bn_mat_free(m);
    return rc;
}

// ---------------------------------------------------------------------------------
// SyntheticControl (R/synthetic_controls.r:57-70): lambda_i = rowMeans(t(t(Data)/colSums(Data)) * Mean_depth)
// with Mean_depth = mean(colSums(Data)/BETA), then N_c[i,j] = DownSampling(rpois(lambda_i), BETA_j)
// = rbinom(1, rpois(1, lambda_i), BETA_j), i.e. Poisson(lambda_i BETA_j) by the thinning property (the reference
// keeps only the thinned matrix).  The column sums, the per-gene means and the draws are three streaming passes;
// the G x C Poisson matrix of the reference is never materialised.
// ---------------------------------------------------------------------------------
// depth[0] = mean_j xx_j / beta_j ; status[0] |= 1 when a cell has no counts (the reference then propagates NaN)
__global__ void __launch_bounds__(1024) k_mean_depth(int64_t C, const double *__restrict__ xx, const double *__restrict__ beta,
                                                     double *__restrict__ depth, int *__restrict__ status)
{
    __shared__ double sh[32];
    double s = 0.0;
    int bad = 0;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        s += __ddiv_rn(xx[j], beta[j]);
        bad |= !(xx[j] > 0.0);
    }
    const double tot = block_reduce_1024<RED_SUM>(s, sh);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        depth[0] = __ddiv_rn(tot, (double)C);
        status[0] = bad;
    }
}

__global__ void __launch_bounds__(256) k_scale_vec(int64_t n, double *__restrict__ v, const double *__restrict__ factor)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) v[k] = __dmul_rn(v[k], factor[0]);
}

// A warp owns 32-gene x 32-cell tiles: the draws run with lane = CELL, one gene at a time (the 32 lanes of a draw share
// lambda_i, so they take the same sampler branch and similar loop lengths: 8.7 -> ~28 active lanes per instruction),
// the stores with lane = gene through a padded shared-memory tile (256 contiguous bytes per instruction).
#define SC_WARPS 8
__global__ void __launch_bounds__(SC_WARPS * 32, 3)
    k_synthetic_control(int64_t G, int64_t C, const double *__restrict__ lambda, const double *__restrict__ beta, uint64_t seed,
                        int64_t gene0, int64_t cell0, double *__restrict__ out)
{
    extern __shared__ double s_t[]; // [warp][cell][gene], rows padded to 33
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *T = s_t + wib * (32 * 33);
    const int64_t tg = (G + 31) / 32, tc = (C + 31) / 32, total = tg * tc;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t t = warp; t < total; t += nwarp) {
        const int64_t ct = t / tg, gt = t - ct * tg;
        const int64_t g0 = gt * 32, c0 = ct * 32;
        const int ncell = (int)(C - c0 < 32 ? C - c0 : 32), ngene = (int)(G - g0 < 32 ? G - g0 : 32);
        const int64_t j = c0 + lane;
        const double b = (lane < ncell) ? __ldg(beta + j) : 0.0;
        const double lam_mine = (lane < ngene) ? __ldg(lambda + g0 + lane) : 0.0;
        for (int gi = 0; gi < ngene; gi++) {
            const double lam = __shfl_sync(0xffffffffu, lam_mine, gi);
            double r = 0.0;
            if (lane < ncell) {
                Philox g;
                g.init(seed, (uint64_t)(gene0 + g0 + gi) + (uint64_t)(cell0 + j) * 0x100000000ULL, 0x53433031u /* "SC01" */);
                // rbinom(1, rpois(1, lambda_i), beta_j) is Poisson(lambda_i beta_j) (binomial thinning of a Poisson
                // count): one draw in float instead of a Poisson and a binomial draw in double; FP64 only for means
                // beyond float's integer grid
                const double lb = __dmul_rn(lam, b);
                r = lb < 4.0e6 ? (double)bn_rpoisf(g, (float)lb) : bn_rpois(g, lb);
            }
            T[lane * 33 + gi] = r;
        }
        __syncwarp();
        if (lane < ngene) {
#pragma unroll 8
            for (int cc = 0; cc < 32; cc++)
                if (cc < ncell) __stcs(out + (c0 + cc) * G + g0 + lane, T[cc * 33 + lane]);
        }
        __syncwarp();
    }
}

extern "C" int bn_synthetic_control(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, uint64_t seed, double *lambda,
                                    double *N_c)
{
    if (!ctx || !m || !BETA_vec) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_synthetic_control: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> b;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    DevOut<double> olam, onc;
    BN_TRY(olam.bind(ctx, lambda, (size_t)m->G, true));
    BN_TRY(onc.bind(ctx, N_c, (size_t)m->G * (size_t)m->C));
    DevBuf<double> xx, depth;
    DevBuf<int> st;
    BN_TRY(xx.alloc(ctx, (size_t)m->C));
    BN_TRY(depth.alloc(ctx, 1));
    BN_TRY(st.alloc(ctx, 1));
    const int blocks = (int)std::min<int64_t>((m->C + 7) / 8, 148 * 16);
    BN_LAUNCH(ctx, k_colsums, blocks, 256, 0, m->C, m->colptr, m->val, xx.p);
    BN_TRY(bn_allreduce(ctx, xx.p, m->C, 0)); // per-cell totals over gene shards
    BN_LAUNCH(ctx, k_mean_depth, 1, 1024, 0, m->C, xx.p, b.p, depth.p, st.p);
    // rowMeans(Data / colSums): the MME mean pass with the column sums in the place of BETA, then * Mean_depth
    BN_TRY(launch_row_moments(ctx, m, 1, xx.p, nullptr, olam.p, nullptr, nullptr));
    BN_LAUNCH(ctx, k_scale_vec, (unsigned)((m->G + 255) / 256), 256, 0, m->G, olam.p, depth.p);
    int h = 0;
    BN_CUDA(ctx, cudaMemcpyAsync(&h, st.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(bn_sync(ctx));
    if (h) return bn_fail(ctx, BN_ERR_INVALID, "SyntheticControl: a cell has no counts (colSums == 0 makes every lambda NaN in the reference)");
    if (onc.p) {
        const int64_t tiles = ((m->G + 31) / 32) * ((m->C + 31) / 32);
        const int grid = (int)std::min<int64_t>((tiles + SC_WARPS - 1) / SC_WARPS, (int64_t)ctx->prop.multiProcessorCount * 24);
        const size_t smem = (size_t)SC_WARPS * 32 * 33 * sizeof(double);
        BN_CUDA(ctx, cudaFuncSetAttribute(k_synthetic_control, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BN_LAUNCH(ctx, k_synthetic_control, grid, SC_WARPS * 32, smem, m->G, m->C, olam.p, b.p, seed, m->gene0, m->cell0, onc.p);
    }
    BN_TRY(olam.commit(ctx));
    BN_TRY(onc.commit(ctx));
    return bn_sync(ctx);
}

// ---------------------------------------------------------------------------------
// cross-gene statistics (single CTA; G is tens of thousands)
// ---------------------------------------------------------------------------------
// out[0] = min over non-NaN v, out[1] = -max over non-NaN v   (so one min-all-reduce serves both)
__global__ void __launch_bounds__(1024) k_minmax_nonan(int64_t n, const double *__restrict__ v, double *__restrict__ out)
{
    __shared__ double sh[32];
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        const double x = v[k];
        if (x == x) {
            mn = fmin(mn, x);
            mx = fmax(mx, x);
        }
    }
    mn = block_reduce_1024<RED_MIN>(mn, sh);
    mx = block_reduce_1024<RED_MAX>(mx, sh);
    if (threadIdx.x == 0) {
        out[0] = mn;
        out[1] = -mx;
    }
}

// size_est[is.na(size_est)] <- min(size_est[!is.na(size_est)])   R/PRIOR_FUNCTIONS.R:259
__global__ void k_fill_nan(int64_t n, double *__restrict__ v, const double *__restrict__ mm)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n && v[k] != v[k]) v[k] = mm[0];
}

// sums over genes with min < BB < max of {1, x, y, xx, xy}, x = log MME_SIZE, y = log BB
__global__ void __launch_bounds__(1024) k_ols_sums(int64_t n, const double *__restrict__ bb, const double *__restrict__ ms,
                                                   const double *__restrict__ mm, double *__restrict__ out)
{
    __shared__ double sh[32];
    const double mn = mm[0], mx = -mm[1];
    double s[5] = {0, 0, 0, 0, 0};
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        const double b = bb[k];
        if (b < mx && b > mn) {
            const double x = log(ms[k]), y = log(b);
            s[0] += 1.0;
            s[1] += x;
            s[2] += y;
            s[3] += x * x;
            s[4] += x * y;
        }
    }
    for (int q = 0; q < 5; q++) {
        const double r = block_reduce_1024<RED_SUM>(s[q], sh);
        if (threadIdx.x == 0) out[q] = r;
    }
}

// coefficients from the five sums, then exp(a + b log MME_SIZE) for every gene
__global__ void k_adjust_apply(int64_t n, const double *__restrict__ ms, const double *__restrict__ sums,
                               double *__restrict__ out, double *__restrict__ coef)
{
    const double N = sums[0], sx = sums[1], sy = sums[2], sxx = sums[3], sxy = sums[4];
    const double mxv = sx / N, myv = sy / N;
    const double b = (sxy - N * mxv * myv) / (sxx - N * mxv * mxv);
    const double a = myv - b * mxv;
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = exp(a + b * log(ms[k]));
    if (k == 0 && coef) {
        coef[0] = a;
        coef[1] = b;
        coef[2] = N;
    }
}

int bn_size_bounds_device(bn_ctx *ctx, int64_t G, double *d_size, double *d_mm /*[2]*/)
{
    BN_LAUNCH(ctx, k_minmax_nonan, 1, 1024, 0, G, d_size, d_mm);
    BN_TRY(bn_allreduce(ctx, d_mm, 2, 1));
    BN_LAUNCH(ctx, k_fill_nan, (unsigned)((G + 255) / 256), 256, 0, G, d_size, d_mm);
    return BN_OK;
}

// min and max of the non-NaN entries of a (gene-sharded) vector over ALL shards
extern "C" int bn_vec_range(bn_ctx *ctx, const double *v, int64_t n, double out[2])
{
    if (!ctx || !v || !out || n <= 0) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_vec_range: bad arguments") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> in;
    BN_TRY(in.bind(ctx, v, (size_t)n));
    DevBuf<double> mm;
    BN_TRY(mm.alloc(ctx, 2));
    BN_LAUNCH(ctx, k_minmax_nonan, 1, 1024, 0, n, in.p, mm.p);
    BN_TRY(bn_allreduce(ctx, mm.p, 2, 1));
    double h[2];
    BN_CUDA(ctx, cudaMemcpyAsync(h, mm.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(bn_sync(ctx));
    out[0] = h[0];
    out[1] = -h[1];
    return BN_OK;
}

int bn_adjust_size_device(bn_ctx *ctx, int64_t G, const double *d_bb, const double *d_ms, double *d_out, double *d_coef)
{
    DevBuf<double> st;
    BN_TRY(st.alloc(ctx, 8));
    BN_LAUNCH(ctx, k_minmax_nonan, 1, 1024, 0, G, d_bb, st.p);
    BN_TRY(bn_allreduce(ctx, st.p, 2, 1));
    BN_LAUNCH(ctx, k_ols_sums, 1, 1024, 0, G, d_bb, d_ms, st.p, st.p + 2);
    BN_TRY(bn_allreduce(ctx, st.p + 2, 5, 0));
    BN_LAUNCH(ctx, k_adjust_apply, (unsigned)((G + 255) / 256), 256, 0, G, d_ms, st.p + 2, d_out, d_coef);
    return BN_OK;
}

extern "C" int bn_adjust_size(bn_ctx *ctx, int64_t G, const double *BB_SIZE, const double *MME_SIZE, double *out,
                              double coef[3])
{
    if (!ctx || !BB_SIZE || !MME_SIZE || !out || G <= 0)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_adjust_size: bad arguments") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> bb, ms;
    BN_TRY(bb.bind(ctx, BB_SIZE, (size_t)G));
    BN_TRY(ms.bind(ctx, MME_SIZE, (size_t)G));
    DevOut<double> o, c;
    BN_TRY(o.bind(ctx, out, (size_t)G));
    BN_TRY(c.bind(ctx, coef, 3));
    BN_TRY(bn_adjust_size_device(ctx, G, bb.p, ms.p, o.p, c.p));
    BN_TRY(o.commit(ctx));
    BN_TRY(c.commit(ctx));
    return bn_sync(ctx);
}
