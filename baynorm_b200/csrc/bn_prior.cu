// bn_prior.cu -- Prior_fun (R/PRIOR_FUNCTIONS.R:241-311, FIX_MU=TRUE) as one native call:
// x/beta normalisation fused into the moments (:253-256), NaN -> global minimum size (:259),
// bounds from the global min / ceiling(max) (:277-279), the per-gene fit (BB_Fun :271-283) and
// AdjustSIZE_fun (:294-296).  The three cross-gene steps are also the three cross-SHARD steps
// of a gene-sharded run; they go through bn_allreduce (<= 7 doubles each).
#include <math.h>

#include "bn_internal.cuh"

int bn_mme_device(bn_ctx *ctx, bn_mat *m, const double *d_beta, double *d_MU, double *d_SIZE, double *d_V);
int bn_size_bounds_device(bn_ctx *ctx, int64_t G, double *d_size, double *d_mm);
int bn_adjust_size_device(bn_ctx *ctx, int64_t G, const double *d_bb, const double *d_ms, double *d_out, double *d_coef);
int bn_fit_size_device(bn_ctx *ctx, bn_mat *m, const double *d_beta, const double *d_mu0, const double *d_size0,
                       double lower, double upper, const double *d_mm, const bn_fit_opts *opts, double *d_size_out,
                       int32_t *d_conv, int32_t *d_nfe);

extern "C" int bn_prior(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, int BB_SIZE, const bn_fit_opts *opts,
                        bn_prior_out *out)
{
    if (!ctx || !m || !BETA_vec || !out) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_prior: NULL argument") : BN_ERR_INVALID;
    if (m->C < 2) return bn_fail(ctx, BN_ERR_INVALID, "bn_prior needs at least 2 cells");
    cudaSetDevice(ctx->device);
    // a shard whose gene block holds a non-count value would leave bn_fit_size_device between the collectives of this call
    // and strand the other shards in the hook: agree on the verdict first
    {
        int rc = (BB_SIZE && !m->is_count) ? BN_ERR_NONCOUNT : BN_OK;
        if (rc) bn_fail(ctx, rc, "the likelihood fit needs non-negative integer counts (dnbinom_mu is 0 elsewhere)");
        rc = bn_agree(ctx, rc);
        if (rc != BN_OK) return rc;
    }
    const size_t G = (size_t)m->G;
    DevIn<double> beta;
    BN_TRY(beta.bind(ctx, BETA_vec, (size_t)m->C));
    DevOut<double> oMU, oSIZE, oBB, oADJ;
    DevOut<int32_t> oConv, oNfe;
    BN_TRY(oMU.bind(ctx, out->MME_MU, G, true));
    BN_TRY(oSIZE.bind(ctx, out->MME_SIZE, G, true));
    BN_TRY(oBB.bind(ctx, out->BB_SIZE, G, BB_SIZE != 0));
    BN_TRY(oADJ.bind(ctx, out->MME_SIZE_adjust, G, BB_SIZE != 0));
    BN_TRY(oConv.bind(ctx, out->conv, G));
    BN_TRY(oNfe.bind(ctx, out->nfeval, G));
    DevBuf<double> stat; // [0..1] min / -max of MME_SIZE, [2..4] coef
    BN_TRY(stat.alloc(ctx, 8));
    cudaEvent_t *ev = ctx->ev;
    BN_CUDA(ctx, cudaEventRecord(ev[0], ctx->stream));
    BN_TRY(bn_mme_device(ctx, m, beta.p, oMU.p, oSIZE.p, nullptr));
    BN_TRY(bn_size_bounds_device(ctx, m->G, oSIZE.p, stat.p));
    // nothing below waits for the host: the box of the fit is read from device memory, so the whole
    // prior (MME -> bounds -> fit -> adjust) is queued in one go and host scheduling noise (8 ranks,
    // NCCL proxy threads) cannot open bubbles between the stages
    double *mm = ctx->pinned, *coef_h = ctx->pinned + 2; // page-locked: the copies really are asynchronous
    mm[0] = mm[1] = 0.0;
    coef_h[0] = coef_h[1] = coef_h[2] = 0.0;
    BN_CUDA(ctx, cudaMemcpyAsync(mm, stat.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BN_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    if (BB_SIZE) {
        BN_TRY(bn_fit_size_device(ctx, m, beta.p, oMU.p, oSIZE.p, 0.0, 0.0, stat.p, opts, oBB.p, oConv.p, oNfe.p));
        BN_CUDA(ctx, cudaEventRecord(ev[2], ctx->stream));
        BN_TRY(bn_adjust_size_device(ctx, m->G, oBB.p, oSIZE.p, oADJ.p, stat.p + 2));
        BN_CUDA(ctx, cudaMemcpyAsync(coef_h, stat.p + 2, 24, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        BN_CUDA(ctx, cudaEventRecord(ev[2], ctx->stream));
    }
    BN_CUDA(ctx, cudaEventRecord(ev[3], ctx->stream));
    BN_TRY(oMU.commit(ctx));
    BN_TRY(oSIZE.commit(ctx));
    BN_TRY(oBB.commit(ctx));
    BN_TRY(oADJ.commit(ctx));
    BN_TRY(oConv.commit(ctx));
    BN_TRY(oNfe.commit(ctx));
    BN_TRY(bn_sync(ctx));
    out->size_lower = mm[0];
    out->size_upper = ceil(-mm[1]);
    out->coef[0] = coef_h[0];
    out->coef[1] = coef_h[1];
    out->coef[2] = coef_h[2];
    if (BB_SIZE && (!(mm[0] > 0.0) || !isfinite(mm[0]) || !isfinite(mm[1])))
        return bn_fail(ctx, BN_ERR_INVALID, "no gene has a finite positive MME size (min %g, max %g)", mm[0], -mm[1]);
    float t01 = 0, t12 = 0, t23 = 0, t03 = 0;
    cudaEventElapsedTime(&t01, ev[0], ev[1]);
    cudaEventElapsedTime(&t12, ev[1], ev[2]);
    cudaEventElapsedTime(&t23, ev[2], ev[3]);
    cudaEventElapsedTime(&t03, ev[0], ev[3]);
    out->seconds[0] = t01 * 1e-3;
    out->seconds[1] = t12 * 1e-3;
    out->seconds[2] = t23 * 1e-3;
    out->seconds[3] = t03 * 1e-3;
    return BN_OK;
}
