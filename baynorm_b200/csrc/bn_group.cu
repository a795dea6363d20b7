// bn_group.cu -- one process driving several GPUs (SURVEY.md 8(b) "Threading", 8(e)): a context per device, the gene-sharded
// path's all-reduces over NCCL, and group versions of the entry points an R session calls (default BETA, Prior_fun, the
// posterior).  Stands where the reference starts a SOCK cluster over genes (R/PRIOR_FUNCTIONS.R:426-427, parallel = TRUE):
// there every worker receives the whole matrix; here device k holds rows [bounds[k], bounds[k+1]) and the devices exchange
// C doubles once plus <= 7 doubles per prior set.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 on the first bn_group_create) and never linked: a process that
// already carries its own NCCL (torch.distributed drives the one-process-per-GPU path through the all-reduce hook) keeps
// exactly one copy, and a single-GPU installation needs none.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "bn_internal.cuh"

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
    bool load()
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) {
            err = std::string("cannot load libnccl.so.2: ") + dlerror();
            return false;
        }
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GetErrorString) {
            err = "libnccl.so.2 lacks ncclCommInitAll / ncclCommDestroy / ncclAllReduce";
            dlclose(lib);
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl; // written once, by the first bn_group_create (callers create groups from one thread)

struct Member {
    ncclComm_t comm = nullptr;
};

} // namespace

struct bn_group {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<bn_ctx *> ctx;
    std::vector<Member> mem;
    std::string err;
};

struct bn_gmat {
    bn_group *grp = nullptr;
    int64_t G = 0, C = 0;
    std::vector<int64_t> bounds; // ndev + 1
    std::vector<bn_mat *> shard;
};

static thread_local std::string g_group_err;

// the all-reduce hook of a member: in place on the context's stream
static int group_allreduce(void *user, void *buf, int64_t count, int dtype, int op, void *stream)
{
    Member *m = static_cast<Member *>(user);
    if (dtype != 0 || op < 0 || op > 2) return 2;
    const ncclRedOp_t rop = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
    const ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, rop, m->comm, (cudaStream_t)stream);
    return r == ncclSuccess ? 0 : 1;
}

// fn(k) on one host thread per device; the largest status code
template <typename F> static int group_run(bn_group *g, F fn)
{
    std::vector<int> rc((size_t)g->ndev, BN_OK);
    if (g->ndev == 1) {
        cudaSetDevice(g->devices[0]);
        return fn(0);
    }
    std::vector<std::thread> th;
    for (int k = 0; k < g->ndev; k++)
        th.emplace_back([&, k] {
            cudaSetDevice(g->devices[(size_t)k]);
            rc[(size_t)k] = fn(k);
        });
    for (auto &t : th) t.join();
    int worst = BN_OK;
    for (int k = 0; k < g->ndev; k++)
        if (rc[(size_t)k] != BN_OK) {
            if (worst == BN_OK) g->err = std::string("device ") + std::to_string(g->devices[(size_t)k]) + ": " + bn_last_error(g->ctx[(size_t)k]);
            worst = std::max(worst, rc[(size_t)k]);
        }
    return worst;
}

extern "C" {

const char *bn_group_last_error(const bn_group *g) { return g ? g->err.c_str() : g_group_err.c_str(); }

int bn_group_create(int ndev, const int *devices, bn_group **out)
{
    if (!out || ndev < 1) return BN_ERR_INVALID;
    *out = nullptr;
    bn_group *g = new bn_group;
    g->ndev = ndev;
    for (int k = 0; k < ndev; k++) g->devices.push_back(devices ? devices[k] : k);
    g->ctx.assign((size_t)ndev, nullptr);
    g->mem.assign((size_t)ndev, Member());
    for (int k = 0; k < ndev; k++) {
        const int rc = bn_ctx_create(g->devices[(size_t)k], &g->ctx[(size_t)k]);
        if (rc != BN_OK) {
            g_group_err = bn_last_error(nullptr);
            bn_group_destroy(g);
            return rc;
        }
    }
    if (ndev > 1) {
        if (!g_nccl.load()) {
            g_group_err = g_nccl.err;
            bn_group_destroy(g);
            return BN_ERR_COLLECTIVE;
        }
        std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, g->devices.data());
        if (r != ncclSuccess) {
            g_group_err = std::string("ncclCommInitAll failed: ") + g_nccl.GetErrorString(r);
            bn_group_destroy(g);
            return BN_ERR_COLLECTIVE;
        }
        for (int k = 0; k < ndev; k++) {
            g->mem[(size_t)k].comm = comms[(size_t)k];
            bn_ctx_set_allreduce(g->ctx[(size_t)k], group_allreduce, &g->mem[(size_t)k], k, ndev);
        }
    }
    *out = g;
    return BN_OK;
}

void bn_group_destroy(bn_group *g)
{
    if (!g) return;
    for (int k = 0; k < g->ndev; k++) {
        if (g->ctx[(size_t)k]) {
            cudaSetDevice(g->devices[(size_t)k]);
            bn_ctx_synchronize(g->ctx[(size_t)k]);
        }
        if (g->mem[(size_t)k].comm) g_nccl.CommDestroy(g->mem[(size_t)k].comm);
        if (g->ctx[(size_t)k]) bn_ctx_destroy(g->ctx[(size_t)k]);
    }
    delete g;
}

int bn_group_size(const bn_group *g) { return g ? g->ndev : 0; }
bn_ctx *bn_group_ctx(bn_group *g, int k) { return (g && k >= 0 && k < g->ndev) ? g->ctx[(size_t)k] : nullptr; }

int bn_group_run(bn_group *g, bn_group_fn fn, void *user)
{
    if (!g || !fn) return BN_ERR_INVALID;
    return group_run(g, [&](int k) { return fn(user, k, g->ctx[(size_t)k]); });
}

// ---- the gene-sharded matrix ---------------------------------------------------------------------------------------
int bn_group_mat_from_csc(bn_group *g, int64_t G, int64_t C, const void *p, int p_is_int64, const int32_t *i, const double *x,
                          bn_gmat **out)
{
    if (!g || !out || G < 1 || C < 1 || !p) return BN_ERR_INVALID;
    *out = nullptr;
    if (G < g->ndev) {
        g->err = "fewer genes than devices";
        return BN_ERR_INVALID;
    }
    auto P = [&](int64_t j) -> int64_t { return p_is_int64 ? static_cast<const int64_t *>(p)[j] : (int64_t) static_cast<const int32_t *>(p)[j]; };
    const int64_t nnz = P(C);
    if (nnz < 0 || (nnz > 0 && (!i || !x))) return BN_ERR_INVALID;
    // contiguous gene blocks of near-equal cost C + 4 nnz_gene (the fit's dense and per-entry parts, the posterior's per-cell part)
    std::vector<int64_t> rownnz((size_t)G, 0);
    for (int64_t k = 0; k < nnz; k++) {
        if (i[k] < 0 || i[k] >= G) {
            g->err = "row index outside [0, G)";
            return BN_ERR_INVALID;
        }
        rownnz[(size_t)i[k]]++;
    }
    bn_gmat *gm = new bn_gmat;
    gm->grp = g;
    gm->G = G;
    gm->C = C;
    gm->bounds.assign((size_t)g->ndev + 1, 0);
    gm->shard.assign((size_t)g->ndev, nullptr);
    {
        const double total = (double)C * (double)G + 4.0 * (double)nnz;
        double acc = 0.0;
        int k = 1;
        for (int64_t r = 0; r < G && k < g->ndev; r++) {
            acc += (double)C + 4.0 * (double)rownnz[(size_t)r];
            if (acc >= total * k / g->ndev || G - (r + 1) <= g->ndev - k) { // every device keeps at least one gene
                gm->bounds[(size_t)k] = std::max<int64_t>(r + 1, gm->bounds[(size_t)k - 1] + 1);
                k++;
            }
        }
        for (; k < g->ndev; k++) gm->bounds[(size_t)k] = gm->bounds[(size_t)k - 1] + 1;
        gm->bounds[(size_t)g->ndev] = G;
    }
    const int rc = group_run(g, [&](int k) -> int {
        const int64_t g0 = gm->bounds[(size_t)k], g1 = gm->bounds[(size_t)k + 1];
        // rows are ascending inside a column: the block's entries of a column are one run
        std::vector<int64_t> sp((size_t)C + 1, 0);
        std::vector<int64_t> lo((size_t)C, 0);
        for (int64_t j = 0; j < C; j++) {
            const int32_t *b = i + P(j), *e = i + P(j + 1);
            const int32_t *a = std::lower_bound(b, e, (int32_t)g0), *z = std::lower_bound(a, e, (int32_t)g1);
            lo[(size_t)j] = a - i;
            sp[(size_t)j + 1] = sp[(size_t)j] + (z - a);
        }
        const int64_t n = sp[(size_t)C];
        std::vector<int32_t> si((size_t)std::max<int64_t>(n, 1));
        std::vector<double> sx((size_t)std::max<int64_t>(n, 1));
        for (int64_t j = 0; j < C; j++) {
            const int64_t a = lo[(size_t)j], cnt = sp[(size_t)j + 1] - sp[(size_t)j], d = sp[(size_t)j];
            for (int64_t t = 0; t < cnt; t++) {
                si[(size_t)(d + t)] = i[a + t] - (int32_t)g0;
                sx[(size_t)(d + t)] = x[a + t];
            }
        }
        int r = bn_mat_from_csc(g->ctx[(size_t)k], g1 - g0, C, sp.data(), 1, si.data(), sx.data(), &gm->shard[(size_t)k]);
        if (r == BN_OK) r = bn_mat_set_origin(gm->shard[(size_t)k], g0, 0);
        return r;
    });
    if (rc != BN_OK) {
        bn_gmat_free(gm);
        return rc;
    }
    *out = gm;
    return BN_OK;
}

void bn_gmat_free(bn_gmat *gm)
{
    if (!gm) return;
    for (size_t k = 0; k < gm->shard.size(); k++)
        if (gm->shard[k]) {
            cudaSetDevice(gm->grp->devices[k]);
            bn_mat_free(gm->shard[k]);
        }
    delete gm;
}

int bn_gmat_bounds(const bn_gmat *gm, int64_t *bounds)
{
    if (!gm || !bounds) return BN_ERR_INVALID;
    for (size_t k = 0; k < gm->bounds.size(); k++) bounds[k] = gm->bounds[k];
    return BN_OK;
}

bn_mat *bn_gmat_shard(bn_gmat *gm, int k) { return (gm && k >= 0 && k < (int)gm->shard.size()) ? gm->shard[(size_t)k] : nullptr; }

// ---- group versions of the entry points ------------------------------------------------------------------------------
int bn_group_beta_default(bn_group *g, bn_gmat *gm, double mean_beta, double *beta)
{
    if (!g || !gm || !beta || gm->grp != g) return BN_ERR_INVALID;
    // every device ends with the same vector (the column sums are all-reduced); device 0's is returned
    std::vector<std::vector<double>> tmp((size_t)g->ndev);
    return group_run(g, [&](int k) -> int {
        double *dst = beta;
        if (k > 0) {
            tmp[(size_t)k].resize((size_t)gm->C);
            dst = tmp[(size_t)k].data();
        }
        return bn_beta_default(g->ctx[(size_t)k], gm->shard[(size_t)k], mean_beta, dst);
    });
}

int bn_group_prior(bn_group *g, bn_gmat *gm, const double *BETA_vec, int BB_SIZE, const bn_fit_opts *opts, bn_prior_out *out)
{
    if (!g || !gm || !BETA_vec || !out || gm->grp != g || !out->MME_MU || !out->MME_SIZE) return BN_ERR_INVALID;
    std::vector<bn_prior_out> po((size_t)g->ndev);
    const int rc = group_run(g, [&](int k) -> int {
        const int64_t g0 = gm->bounds[(size_t)k];
        bn_prior_out &o = po[(size_t)k];
        o = *out;
        o.MME_MU = out->MME_MU + g0;
        o.MME_SIZE = out->MME_SIZE + g0;
        o.BB_SIZE = out->BB_SIZE ? out->BB_SIZE + g0 : nullptr;
        o.MME_SIZE_adjust = out->MME_SIZE_adjust ? out->MME_SIZE_adjust + g0 : nullptr;
        o.conv = out->conv ? out->conv + g0 : nullptr;
        o.nfeval = out->nfeval ? out->nfeval + g0 : nullptr;
        return bn_prior(g->ctx[(size_t)k], gm->shard[(size_t)k], BETA_vec, BB_SIZE, opts, &o);
    });
    if (rc != BN_OK) return rc;
    out->size_lower = po[0].size_lower;
    out->size_upper = po[0].size_upper;
    for (int q = 0; q < 3; q++) out->coef[q] = po[0].coef[q];
    for (int q = 0; q < 4; q++) {
        out->seconds[q] = 0.0;
        for (int k = 0; k < g->ndev; k++) out->seconds[q] = std::max(out->seconds[q], po[(size_t)k].seconds[q]);
    }
    return BN_OK;
}

int bn_group_post(bn_group *g, bn_gmat *gm, int kind, const double *BETA_vec, const double *size, const double *mu, int64_t col0,
                  int64_t ncol, int S, uint64_t seed, double *out)
{
    if (!g || !gm || gm->grp != g || !BETA_vec || !size || !mu || !out || kind < 0 || kind > 2 || col0 < 0 || ncol < 1 ||
        col0 + ncol > gm->C || (kind == 2 && S < 1))
        return BN_ERR_INVALID;
    const int planes = kind == 2 ? S : 1;
    const int64_t G = gm->G;
    return group_run(g, [&](int k) -> int {
        bn_ctx *ctx = g->ctx[(size_t)k];
        const bn_mat *m = gm->shard[(size_t)k];
        const int64_t g0 = gm->bounds[(size_t)k], Gk = gm->bounds[(size_t)k + 1] - g0;
        DevIn<double> b, sz, mm;
        BN_TRY(b.bind(ctx, BETA_vec, (size_t)gm->C));
        BN_TRY(sz.bind(ctx, size + g0, (size_t)Gk));
        BN_TRY(mm.bind(ctx, mu + g0, (size_t)Gk));
        // column tiles of <= 256 MB of result; each device writes its row block of the caller's G x ncol (x S) array
        int64_t tc = std::max<int64_t>(1, (int64_t)(256.0e6 / (8.0 * (double)Gk * (double)planes)));
        tc = std::min(tc, ncol);
        double *tile = nullptr;
        BN_TRY(bn_ctx_scratch(ctx, 6, (size_t)Gk * (size_t)tc * (size_t)planes * 8, (void **)&tile));
        for (int64_t c = 0; c < ncol; c += tc) {
            const int64_t n = std::min(tc, ncol - c);
            BN_TRY(bn_post_device(ctx, kind, m, b.p, sz.p, mm.p, col0 + c, n, S, seed, tile));
            for (int s = 0; s < planes; s++)
                BN_CUDA(ctx, cudaMemcpy2DAsync(out + g0 + G * (c + ncol * (int64_t)s), (size_t)G * 8, tile + Gk * n * (int64_t)s, (size_t)Gk * 8,
                                               (size_t)Gk * 8, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
            BN_TRY(bn_sync(ctx));
        }
        return BN_OK;
    });
}

} // extern "C"
