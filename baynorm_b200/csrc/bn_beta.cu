// bn_beta.cu -- BetaFun (R/PRIOR_FUNCTIONS.R:30-71): cell capture efficiencies from the library sizes of
// a robustly chosen gene subset.
//
//   xx        = colSums(Data);  Normcount = t(t(Data)/xx) * mean(xx)                        (:36-39)
//   per gene over ALL cells of log(Normcount + 1):  med = median, mad = 1.4826 median|. - med|, max   (:44-51)
//   select    = (max < med + 3 mad)  and  (fraction of zero cells < 0.35)                    (:52-57)
//   BETA      = colSums(Data[select, ]) / mean(.) * MeanBETA, >= 1 -> max below 1, <= 0 -> min above 0   (:60-67)
//
// The reference runs three R-level `apply` loops over a dense G x C matrix.  Here only genes that can
// pass the zero-fraction test ("candidates", expressed in > 65 % of the cells) are looked at: their
// stored entries are transformed into one contiguous buffer, sorted per gene by a segmented radix sort,
// and the order statistics of the full length-C vector follow by index arithmetic, because the
// C - nnz absent cells all contribute log(0 + 1) = 0 (and |0 - med| = med to the deviations).
// Gene-sharded runs exchange the two length-C column sums through bn_allreduce(); everything per gene
// is shard-local.
#include <cub/cub.cuh>
#include <math.h>

#include "bn_internal.cuh"

// lane-strided read of a gene-major row (count matrices: packed (cell << 32) | count; otherwise cell[] + rval[])
__device__ __forceinline__ void bf_load(const uint64_t *__restrict__ gm, const int32_t *__restrict__ cell,
                                        const double *__restrict__ rval, int64_t k, int &c, double &x)
{
    if (gm) {
        const uint64_t e = __ldg(gm + k);
        c = (int)(e >> 32);
        x = (double)(uint32_t)e;
    } else {
        c = cell[k];
        x = rval[k];
    }
}

__global__ void __launch_bounds__(1024) k_bf_mean(int64_t n, const double *__restrict__ v, double *__restrict__ out)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += v[j];
    s = bn_warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; k++) t += sh[k];
        out[0] = t / (double)n;
    }
}

// warp per gene: zero fraction; candidates get their stored-entry count as segment length
__global__ void __launch_bounds__(256) k_bf_count(int64_t G, int64_t C, const int64_t *__restrict__ rowptr,
                                                  const uint64_t *__restrict__ gm, const int32_t *__restrict__ cell,
                                                  const double *__restrict__ rval, double *__restrict__ dropout,
                                                  int64_t *__restrict__ seglen)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t g = warp; g < G; g += nwarp) {
        const int64_t a = rowptr[g], b = rowptr[g + 1];
        int64_t npos = 0;
        for (int64_t k = a + lane; k < b; k += 32) {
            int c;
            double x;
            bf_load(gm, cell, rval, k, c, x);
            npos += (x != 0.0);
        }
        for (int o = 16; o > 0; o >>= 1) npos += __shfl_xor_sync(0xffffffffu, npos, o);
        if (lane == 0) {
            const double d = (double)(C - npos) / (double)C; // length(which(x == 0)) / length(x)
            dropout[g] = d;
            seglen[g] = (d < 0.35) ? (b - a) : 0;
        }
    }
    if (warp == 0 && lane == 0) seglen[G] = 0;
}

// warp per candidate gene: log(x / xx_c * mean(xx) + 1) of every stored entry into the gene's segment
__global__ void __launch_bounds__(256) k_bf_fill(int64_t G, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
                                                 const int32_t *__restrict__ cell, const double *__restrict__ rval,
                                                 const double *__restrict__ xx, const double *__restrict__ mean_xx,
                                                 const int64_t *__restrict__ segoff, double *__restrict__ buf)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const double mx = mean_xx[0];
    for (int64_t g = warp; g < G; g += nwarp) {
        const int64_t o = segoff[g], n = segoff[g + 1] - o;
        if (n == 0) continue;
        const int64_t a = rowptr[g];
        for (int64_t k = lane; k < n; k += 32) {
            int c;
            double x;
            bf_load(gm, cell, rval, a + k, c, x);
            buf[o + k] = log(__dadd_rn(__dmul_rn(__ddiv_rn(x, xx[c]), mx), 1.0));
        }
    }
}

// k-th smallest (0-based) of {z zeros} U sorted[0..n)
__device__ __forceinline__ double bf_kth_with_zeros(const double *__restrict__ sorted, int64_t z, int64_t k)
{
    return k < z ? 0.0 : sorted[k - z];
}
// k-th smallest of {z copies of m} U sorted[0..n), c_lt = #{sorted < m}
__device__ __forceinline__ double bf_kth_with_const(const double *__restrict__ sorted, int64_t z, int64_t c_lt, double m, int64_t k)
{
    return k < c_lt ? sorted[k] : (k < c_lt + z ? m : sorted[k - z]);
}

// thread per gene: median and maximum of the full length-C vector
__global__ void __launch_bounds__(256) k_bf_median(int64_t G, int64_t C, const int64_t *__restrict__ segoff,
                                                   const double *__restrict__ sorted, double *__restrict__ med,
                                                   double *__restrict__ mx)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t o = segoff[g], n = segoff[g + 1] - o;
    if (n == 0) {
        med[g] = nan("");
        mx[g] = nan("");
        return;
    }
    const int64_t z = C - n;
    const double *s = sorted + o;
    // stats::median: sort(x)[half] for odd n, mean(sort(x)[half + 0:1]) otherwise
    med[g] = (C & 1) ? bf_kth_with_zeros(s, z, C / 2)
                     : (bf_kth_with_zeros(s, z, C / 2 - 1) + bf_kth_with_zeros(s, z, C / 2)) / 2.0;
    mx[g] = s[n - 1]; // >= 0 = the value of the absent cells
}

// |v - med| in place (warp per candidate gene)
__global__ void __launch_bounds__(256) k_bf_dev(int64_t G, const int64_t *__restrict__ segoff, const double *__restrict__ med,
                                                double *__restrict__ buf)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t g = warp; g < G; g += nwarp) {
        const int64_t o = segoff[g], n = segoff[g + 1] - o;
        const double m = med[g];
        for (int64_t k = lane; k < n; k += 32) buf[o + k] = fabs(__dsub_rn(buf[o + k], m));
    }
}

// thread per gene: mad = 1.4826 * median(|x - med|), selection flag
__global__ void __launch_bounds__(256) k_bf_select(int64_t G, int64_t C, const int64_t *__restrict__ segoff,
                                                   const double *__restrict__ sdev, const double *__restrict__ med,
                                                   const double *__restrict__ mx, double *__restrict__ mad,
                                                   int32_t *__restrict__ sel)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t o = segoff[g], n = segoff[g + 1] - o;
    if (n == 0) { // zero fraction >= 0.35: excluded whatever the spread
        mad[g] = nan("");
        sel[g] = 0;
        return;
    }
    const int64_t z = C - n;
    const double *s = sdev + o;
    const double m = fabs(med[g]); // the absent cells deviate by |0 - med|
    int64_t lo = 0, hi = n;        // c_lt = #{s < m}
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (s[mid] < m) lo = mid + 1;
        else hi = mid;
    }
    const double md = (C & 1) ? bf_kth_with_const(s, z, lo, m, C / 2)
                              : (bf_kth_with_const(s, z, lo, m, C / 2 - 1) + bf_kth_with_const(s, z, lo, m, C / 2)) / 2.0;
    const double d = __dmul_rn(1.4826, md); // stats::mad, constant = 1.4826
    mad[g] = d;
    sel[g] = (mx[g] < __dadd_rn(med[g], __dmul_rn(3.0, d))) ? 1 : 0;
}

// colSums(Data): warp per column
__global__ void __launch_bounds__(256) k_bf_colsums(int64_t C, const int64_t *__restrict__ colptr, const double *__restrict__ val,
                                                    double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        double s = 0.0;
        for (int64_t k = a + lane; k < b; k += 32) s += __ldcs(val + k);
        s = bn_warp_sum(s);
        if (lane == 0) out[j] = s;
    }
}

// colSums(Data[select, ]): warp per column over the CSC
__global__ void __launch_bounds__(256) k_bf_colsums_sel(int64_t C, const int64_t *__restrict__ colptr,
                                                        const int32_t *__restrict__ rowidx, const double *__restrict__ val,
                                                        const int32_t *__restrict__ sel, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t j = warp; j < C; j += nwarp) {
        const int64_t a = colptr[j], b = colptr[j + 1];
        double s = 0.0;
        for (int64_t k = a + lane; k < b; k += 32)
            if (__ldg(sel + __ldg(rowidx + k))) s += __ldcs(val + k);
        s = bn_warp_sum(s);
        if (lane == 0) out[j] = s;
    }
}

// BETA = t/mean(t) * MeanBETA; >= 1 -> max(BETA < 1) first, then <= 0 -> min(BETA > 0)   (:61-67, in that order)
__global__ void __launch_bounds__(1024) k_bf_beta(int64_t C, const double *__restrict__ t, double mean_beta,
                                                  double *__restrict__ beta)
{
    __shared__ double sh[32];
    __shared__ double bc;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto reduce = [&](double v, int op) { // 0 sum, 1 min, 2 max; fixed order
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = op == 0 ? v + u : (op == 1 ? fmin(v, u) : fmax(v, u));
        }
        __syncthreads();
        if (lane == 0) sh[w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = sh[0];
            for (int k = 1; k < 32; k++) r = op == 0 ? r + sh[k] : (op == 1 ? fmin(r, sh[k]) : fmax(r, sh[k]));
            bc = r;
        }
        __syncthreads();
        return bc;
    };
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) s += t[j];
    const double mean = reduce(s, 0) / (double)C;
    double mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        const double b = __dmul_rn(__ddiv_rn(t[j], mean), mean_beta);
        beta[j] = b;
        if (b < 1.0) mx = fmax(mx, b);
    }
    const double maxlt1 = reduce(mx, 2);
    double mn = INFINITY;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
        double b = beta[j];
        if (b >= 1.0) {
            b = maxlt1;
            beta[j] = b;
        }
        if (b > 0.0) mn = fmin(mn, b);
    }
    const double minpos = reduce(mn, 1);
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x)
        if (beta[j] <= 0.0) beta[j] = minpos;
}

extern "C" int bn_beta_fun(bn_ctx *ctx, bn_mat *m, double MeanBETA, double *BETA, int32_t *selected, double *stats)
{
    if (!ctx || !m || !BETA) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_beta_fun: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    BN_TRY(bn_mat_ensure_csr(ctx, m));
    const int64_t G = m->G, C = m->C;
    DevOut<double> obeta, ostats;
    DevOut<int32_t> osel;
    BN_TRY(obeta.bind(ctx, BETA, (size_t)C));
    BN_TRY(osel.bind(ctx, selected, (size_t)G, true));
    BN_TRY(ostats.bind(ctx, stats, (size_t)G * 4, true)); // [med | mad | max | zero fraction], G each
    double *med = ostats.p, *mad = ostats.p + G, *mx = ostats.p + 2 * G, *drop = ostats.p + 3 * G;
    DevBuf<double> xx, mean_xx, tsel;
    DevBuf<int64_t> seglen, segoff;
    BN_TRY(xx.alloc(ctx, (size_t)C));
    BN_TRY(mean_xx.alloc(ctx, 1));
    BN_TRY(tsel.alloc(ctx, (size_t)C));
    BN_TRY(seglen.alloc(ctx, (size_t)G + 1));
    BN_TRY(segoff.alloc(ctx, (size_t)G + 1));
    const int cblocks = (int)std::min<int64_t>((C + 7) / 8, 148 * 16), gblocks = (int)std::min<int64_t>((G + 7) / 8, 148 * 16);
    // column sums of the whole matrix (all gene shards) and their mean
    BN_LAUNCH(ctx, k_bf_colsums, cblocks, 256, 0, C, m->colptr, m->val, xx.p);
    BN_TRY(bn_allreduce(ctx, xx.p, C, 0));
    BN_LAUNCH(ctx, k_bf_mean, 1, 1024, 0, C, xx.p, mean_xx.p);
    BN_LAUNCH(ctx, k_bf_count, gblocks, 256, 0, G, C, m->rowptr, m->gm, m->cell, m->rval, drop, seglen.p);
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, seglen.p, segoff.p, (int)(G + 1), ctx->stream);
        DevBuf<char> tmp;
        BN_TRY(tmp.alloc(ctx, tb));
        BN_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, seglen.p, segoff.p, (int)(G + 1), ctx->stream));
        ctx->launches++;
    }
    int64_t total = 0;
    BN_CUDA(ctx, cudaMemcpyAsync(&total, segoff.p + G, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(bn_sync(ctx));
    DevBuf<double> bufa, bufb;
    BN_TRY(bufa.alloc(ctx, (size_t)std::max<int64_t>(total, 1)));
    BN_TRY(bufb.alloc(ctx, (size_t)std::max<int64_t>(total, 1)));
    DevBuf<char> sort_tmp;
    size_t sort_bytes = 0;
    if (total > 0) {
        if (total > 2147483647LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "BetaFun: more than 2^31-1 entries in candidate genes per shard");
        cub::DeviceSegmentedSort::SortKeys(nullptr, sort_bytes, (const double *)bufa.p, bufb.p, (int)total, (int)G, segoff.p, segoff.p + 1,
                                           ctx->stream);
        BN_TRY(sort_tmp.alloc(ctx, sort_bytes));
        BN_LAUNCH(ctx, k_bf_fill, gblocks, 256, 0, G, m->rowptr, m->gm, m->cell, m->rval, xx.p, mean_xx.p, segoff.p, bufa.p);
        BN_CUDA(ctx, cub::DeviceSegmentedSort::SortKeys(sort_tmp.p, sort_bytes, (const double *)bufa.p, bufb.p, (int)total, (int)G,
                                                        segoff.p, segoff.p + 1, ctx->stream));
        ctx->launches++;
    }
    BN_LAUNCH(ctx, k_bf_median, (unsigned)((G + 255) / 256), 256, 0, G, C, segoff.p, bufb.p, med, mx);
    if (total > 0) {
        BN_LAUNCH(ctx, k_bf_dev, gblocks, 256, 0, G, segoff.p, med, bufb.p);
        BN_CUDA(ctx, cub::DeviceSegmentedSort::SortKeys(sort_tmp.p, sort_bytes, (const double *)bufb.p, bufa.p, (int)total, (int)G,
                                                        segoff.p, segoff.p + 1, ctx->stream));
        ctx->launches++;
    }
    BN_LAUNCH(ctx, k_bf_select, (unsigned)((G + 255) / 256), 256, 0, G, C, segoff.p, bufa.p, med, mx, mad, osel.p);
    BN_LAUNCH(ctx, k_bf_colsums_sel, cblocks, 256, 0, C, m->colptr, m->rowidx, m->val, osel.p, tsel.p);
    BN_TRY(bn_allreduce(ctx, tsel.p, C, 0));
    BN_LAUNCH(ctx, k_bf_beta, 1, 1024, 0, C, tsel.p, MeanBETA, obeta.p);
    BN_TRY(obeta.commit(ctx));
    BN_TRY(osel.commit(ctx));
    BN_TRY(ostats.commit(ctx));
    return bn_sync(ctx);
}
