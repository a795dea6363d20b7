// bn_post.cu -- the gene x cell posterior kernels (Main_mean_NB_Bay, Main_mode_NB_Bay,
// Main_NB_Bay; src/bayNorm_main.cpp:1063-1299), binomial DownSampling (:892-915) and the
// synthetic NB count generator the benchmarks use.
//
// Layout: a CTA owns a tile of TG = 2048 consecutive genes and walks NC consecutive cell
// columns.  size/mu of the tile live in registers for the whole walk.  For every column the
// CTA scatters that column's non-zeros (row indices are sorted, so the tile's slice is one
// contiguous run found by binary search) into a shared-memory tile of counts, then every
// thread evaluates the closed form for its genes and streams the result out: each output
// element is written exactly once, 256 threads x 8 B contiguous per store instruction.
// HBM traffic per element: 8 B written + 12 B * nnz-fraction read; beta/size/mu are
// O(G + C) and stay in L2.
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

#define POST_THREADS 256
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

// first index in [lo, hi) with rowidx[k] >= key
__device__ __forceinline__ int64_t lower_bound_row(const int32_t *__restrict__ rowidx, int64_t lo, int64_t hi, int64_t key)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowidx + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Dynamic shared memory: three count tiles of POST_TG doubles.  In iteration jj the CTA
//   (1) issues the loads of column jj's non-zeros (tile slice) into registers,
//   (2) computes and stores column jj-1 from tile (jj-1)%3 and zeroes tile (jj+1)%3,
//   (3) scatters the prefetched non-zeros into tile jj%3 (zeroed one iteration earlier),
// then one __syncthreads().  Global-load latency of (1) is covered by the arithmetic of (2).
// POST_PF: non-zeros per thread prefetched two columns ahead (2 covers 512 per tile column: sparse
// single-cell matrices; 4 covers 1024: dense-ish matrices such as config 2)
template <int MODE, int POST_EPT, int POST_PF>
__global__ void __launch_bounds__(POST_THREADS, POST_EPT == 8 ? 3 : 6)
    k_posterior(int64_t G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                const double *__restrict__ val, const double *__restrict__ beta, const double *__restrict__ size,
                const double *__restrict__ mu, int64_t col0, int64_t ncol, int nc_per_cta, double *__restrict__ out,
                int S, uint64_t seed, int64_t gene0, int64_t cell0, int64_t sstride)
{
    constexpr int POST_TG = POST_THREADS * POST_EPT;
    extern __shared__ double xs[]; // [3][POST_TG]
    __shared__ int64_t s_lo[POST_NCMAX + 2], s_end[POST_NCMAX + 2];
    __shared__ double s_beta[POST_NCMAX];
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * POST_TG;
    const int64_t tend = t0 + POST_TG;
    const int64_t jb = (int64_t)blockIdx.y * nc_per_cta;
    const int nc = (int)((jb + nc_per_cta < ncol ? jb + nc_per_cta : ncol) - jb);
    // every column's slice start, all columns in parallel (one thread each)
    if (tid < nc) {
        const int64_t j = col0 + jb + tid;
        const int64_t cs = __ldg(colptr + j), ce = __ldg(colptr + j + 1);
        s_lo[tid] = (t0 == 0) ? cs : lower_bound_row(rowidx, cs, ce, t0);
        s_end[tid] = ce;
        s_beta[tid] = __ldg(beta + j);
    } else if (tid < nc + 2) {
        s_lo[tid] = 0; // empty slices past the chunk keep the pipeline code branch-free
        s_end[tid] = 0;
    }
    double sz[POST_EPT], m[POST_EPT];
#pragma unroll
    for (int e = 0; e < POST_EPT; e++) {
        const int64_t i = t0 + e * POST_THREADS + tid;
        sz[e] = (i < G) ? __ldg(size + i) : 1.0;
        m[e] = (i < G) ? __ldg(mu + i) : 0.0;
    }
#pragma unroll
    for (int e = 0; e < POST_EPT; e++) {
        xs[e * POST_THREADS + tid] = 0.0;
        xs[POST_TG + e * POST_THREADS + tid] = 0.0;
    }
    __syncthreads();
    // register sets for the slices of two columns in flight (loaded one iteration before use)
    int32_t prA[POST_PF], prB[POST_PF];
    double pvA[POST_PF], pvB[POST_PF];
#define POST_LOAD(PR, PV, COL)                                                           \
    do {                                                                                 \
        const int64_t kb_ = s_lo[COL], ke_ = s_end[COL];                                 \
        _Pragma("unroll") for (int q = 0; q < POST_PF; q++)                              \
        {                                                                                \
            const int64_t k = kb_ + q * POST_THREADS + tid;                              \
            PR[q] = (k < ke_) ? __ldg(rowidx + k) : 0x7fffffff;                          \
            PV[q] = (k < ke_) ? __ldcs(val + k) : 0.0;                                   \
        }                                                                                \
    } while (0)
    // one pipeline stage: load slice of column jj+1, compute column jj-1, zero tile jj+1, scatter column jj
#define POST_STAGE(JJ, PR_USE, PV_USE, PR_LOAD, PV_LOAD)                                                 \
    do {                                                                                                 \
        const int jj = (JJ);                                                                             \
        double *xcur = xs + (jj % 3) * POST_TG, *xprev = xs + ((jj + 2) % 3) * POST_TG,                  \
               *xnext = xs + ((jj + 1) % 3) * POST_TG;                                                   \
        POST_LOAD(PR_LOAD, PV_LOAD, jj + 1);                                                             \
        if (jj > 0) {                                                                                    \
            const int64_t jl = jb + jj - 1;                                                              \
            const double b = s_beta[jj - 1];                                                             \
            double *o = out + G * jl;                                                                    \
            _Pragma("unroll") for (int e = 0; e < POST_EPT; e++)                                         \
            {                                                                                            \
                const int64_t i = t0 + e * POST_THREADS + tid;                                           \
                const double xv = xprev[e * POST_THREADS + tid];                                         \
                if (i < G) {                                                                             \
                    if (MODE == POST_MEAN) __stcs(o + i, __dadd_rn(post_tempmu(xv, sz[e], m[e], b), xv)); \
                    else __stcs(o + i, post_mode(xv, sz[e], m[e], b));                                   \
                }                                                                                        \
            }                                                                                            \
        }                                                                                                \
        _Pragma("unroll") for (int e = 0; e < POST_EPT; e++) xnext[e * POST_THREADS + tid] = 0.0;        \
        if (jj < nc) {                                                                                   \
            bool more = true;                                                                            \
            _Pragma("unroll") for (int q = 0; q < POST_PF; q++)                                          \
            {                                                                                            \
                if ((int64_t)PR_USE[q] < tend) xcur[PR_USE[q] - t0] = PV_USE[q];                         \
                else more = false;                                                                       \
            }                                                                                            \
            if (more) { /* dense tile column: the rest of the slice */                                   \
                const int64_t ke_ = s_end[jj];                                                           \
                for (int64_t k = s_lo[jj] + POST_PF * POST_THREADS + tid; k < ke_; k += POST_THREADS) {  \
                    const int64_t r = __ldg(rowidx + k);                                                 \
                    if (r >= tend) break;                                                                \
                    xcur[r - t0] = __ldcs(val + k);                                                      \
                }                                                                                        \
            }                                                                                            \
        }                                                                                                \
        __syncthreads();                                                                                 \
    } while (0)
    POST_LOAD(prA, pvA, 0);
    for (int j2 = 0; j2 <= nc; j2 += 2) {
        POST_STAGE(j2, prA, pvA, prB, pvB);
        if (j2 + 1 <= nc) POST_STAGE(j2 + 1, prB, pvB, prA, pvA);
    }
#undef POST_STAGE
#undef POST_LOAD
}

// ---------------------------------------------------------------------------------
// 3-D draws (Main_NB_Bay, src/bayNorm_main.cpp:1063-1139): S draws x + NB(size+x, tempmu) per
// gene x cell, out[i + G*(j + ncol*s)].
//
// Same tile walk as k_posterior (1024-gene tiles).  Per tile column:
//   phase 1  every thread handles its 4 genes: NB parameters in FP64 (the reference's expression),
//            then, if the mean is small (< SAMPLE_LIGHT_MAX: all x = 0 entries of ordinary genes,
//            i.e. ~92 % of the matrix), S draws by single-uniform CDF inversion in float, stored
//            straight to the S output planes (each store instruction writes 256 contiguous
//            doubles); otherwise the gene is appended to the tile column's heavy list;
//   phase 2  the heavy list x S is flattened over the CTA: each thread takes (entry, s) items and
//            draws gamma-Poisson.  Light and heavy work never share a warp instruction, so a
//            non-zero count does not stall 31 cheap lanes for hundreds of instructions.
// Draw (gene, cell, s) is a pure function of (seed, global gene, global cell, s): light draws use
// Philox block s/4 word s%4 of the element's stream, heavy draws their own sub-stream 2^31 + s.
// ---------------------------------------------------------------------------------
#define SAMPLE_EPT 4
#define SAMPLE_TG (POST_THREADS * SAMPLE_EPT)
#define SAMPLE_LIGHT_MAX 20.0f  // means above this never put enough mass on SAMPLE_K points
#define SAMPLE_LIGHT_MASS 0.95f // mass the 16-point CDF must cover for the table path (the rest: sequential tail)

struct HeavyEntry {
    float r, M;
    int gene; // local to the tile
};
#define SAMPLE_K 32 // support points in the per-element CDF table
#define SAMPLE_SMEM_BYTES (3 * SAMPLE_TG * 8 + SAMPLE_TG * 12 + SAMPLE_K * POST_THREADS * 4)

__global__ void __launch_bounds__(POST_THREADS, 4)
    k_post_sample(int64_t G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                  const double *__restrict__ val, const double *__restrict__ beta, const double *__restrict__ size,
                  const double *__restrict__ mu, int64_t col0, int64_t ncol, int nc_per_cta, double *__restrict__ out,
                  int S, uint64_t seed, int64_t gene0, int64_t cell0, int64_t sstride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*xs)[SAMPLE_TG] = reinterpret_cast<double(*)[SAMPLE_TG]>(smem_raw);                      // [3][TG] count tiles
    HeavyEntry *heavy = reinterpret_cast<HeavyEntry *>(smem_raw + 3 * SAMPLE_TG * 8);                // [TG]
    float *cdf = reinterpret_cast<float *>(smem_raw + 3 * SAMPLE_TG * 8 + SAMPLE_TG * 12);           // [SAMPLE_K][threads]
    __shared__ int64_t s_lo[POST_NCMAX], s_end[POST_NCMAX];
    __shared__ float inv_tab[64];
    __shared__ int n_heavy[2]; // alternating per column so the reset never races with the next column's appends
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * SAMPLE_TG;
    const int64_t tend = t0 + SAMPLE_TG;
    const int64_t jb = (int64_t)blockIdx.y * nc_per_cta;
    const int nc = (int)((jb + nc_per_cta < ncol ? jb + nc_per_cta : ncol) - jb);
    if (tid < nc) {
        const int64_t j = col0 + jb + tid;
        const int64_t cs = __ldg(colptr + j), ce = __ldg(colptr + j + 1);
        s_lo[tid] = (t0 == 0) ? cs : lower_bound_row(rowidx, cs, ce, t0);
        s_end[tid] = ce;
    }
    if (tid < 64) inv_tab[tid] = 1.0f / (float)(tid + 1);
    if (tid == 0) n_heavy[0] = n_heavy[1] = 0;
    double sz[SAMPLE_EPT], m[SAMPLE_EPT];
#pragma unroll
    for (int e = 0; e < SAMPLE_EPT; e++) {
        const int64_t i = t0 + e * POST_THREADS + tid;
        sz[e] = (i < G) ? __ldg(size + i) : 1.0;
        m[e] = (i < G) ? __ldg(mu + i) : 0.0;
        xs[0][e * POST_THREADS + tid] = 0.0;
        xs[1][e * POST_THREADS + tid] = 0.0;
    }
    __syncthreads();
    for (int jj = 0; jj <= nc; jj++) {
        double *xcur = xs[jj % 3], *xprev = xs[(jj + 2) % 3], *xnext = xs[(jj + 1) % 3];
        int32_t pr[2];
        double pv[2];
        int64_t kbase = 0, kend = 0;
        if (jj < nc) {
            kbase = s_lo[jj];
            kend = s_end[jj];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int64_t k = kbase + q * POST_THREADS + tid;
                pr[q] = (k < kend) ? __ldg(rowidx + k) : 0x7fffffff;
                pv[q] = (k < kend) ? __ldcs(val + k) : 0.0;
            }
        }
        const int64_t jl = jb + jj - 1, j = col0 + jl;
        double *o = out + G * jl;
        if (jj > 0) { // phase 1 of column jj-1
            const double b = __ldg(beta + j);
#pragma unroll 1
            for (int e = 0; e < SAMPLE_EPT; e++) {
                const int64_t i = t0 + e * POST_THREADS + tid;
                if (i >= G) continue;
                const double xv = xprev[e * POST_THREADS + tid];
                const double tm = post_tempmu(xv, sz[e], m[e], b);
                const double r = __dadd_rn(sz[e], xv);
                if (!(tm > 0.0)) { // rgamma(shape, scale = 0) = 0 -> rpois(0) = 0: the draw is x itself
                    for (int s = 0; s < S; s++) __stcs(o + i + sstride * s, xv);
                    continue;
                }
                // CDF of the first SAMPLE_K support points, built once per element and shared by its S draws
                const float rf = (float)r, Mf = (float)tm;
                float ctop = 0.0f, p = 0.0f;
                if (Mf < SAMPLE_LIGHT_MAX) {
                    const float q = Mf / (rf + Mf);
                    p = __expf(-rf * log1pf(Mf / rf));
                    ctop = p;
                    cdf[0 * POST_THREADS + tid] = ctop;
#pragma unroll
                    for (int k = 1; k < SAMPLE_K; k++) {
                        p *= q * (rf + (float)(k - 1)) * inv_tab[k - 1];
                        ctop += p;
                        cdf[k * POST_THREADS + tid] = ctop;
                    }
                }
                if (ctop >= SAMPLE_LIGHT_MASS) {
                    // light element: each draw is a branch-free 5-step search of the CDF table; a uniform
                    // beyond it (probability < 5 %) continues the recurrence sequentially
                    const float q = Mf / (rf + Mf);
                    Philox g;
                    g.init(seed, (uint64_t)(gene0 + i) + (uint64_t)(cell0 + j) * 0x100000000ULL, 0u);
                    for (int s0 = 0; s0 < S; s0 += 4) {
                        g.c2 = (uint32_t)(s0 >> 2);
                        g.c3 = 0;
                        g.refill();
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            if (s0 + w >= S) break;
                            const float u = bn_u01f(g.out[w]);
                            int k = (u > cdf[15 * POST_THREADS + tid]) ? 16 : 0;
                            k += (u > cdf[(k + 7) * POST_THREADS + tid]) ? 8 : 0;
                            k += (u > cdf[(k + 3) * POST_THREADS + tid]) ? 4 : 0;
                            k += (u > cdf[(k + 1) * POST_THREADS + tid]) ? 2 : 0;
                            k += (u > cdf[k * POST_THREADS + tid]) ? 1 : 0;
                            if (k == SAMPLE_K - 1 && u > ctop) { // tail
                                float pk = p, ck = ctop, kk = (float)(SAMPLE_K - 1);
                                // stop once p(k) no longer moves the float CDF: u then lies in the last ~1e-7 of
                                // mass and k is already at that quantile
                                while (u > ck && pk > ck * 3e-8f && kk < 4096.0f) {
                                    pk *= q * (rf + kk) / (kk + 1.0f);
                                    ck += pk;
                                    kk += 1.0f;
                                }
                                k = (int)kk;
                            }
                            __stcs(o + i + sstride * (s0 + w), (double)k + xv);
                        }
                    }
                } else {
                    const int slot = atomicAdd(&n_heavy[jj & 1], 1);
                    HeavyEntry h;
                    h.r = rf;
                    h.M = Mf;
                    h.gene = e * POST_THREADS + tid;
                    heavy[slot] = h;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < SAMPLE_EPT; e++) xnext[e * POST_THREADS + tid] = 0.0;
        if (jj < nc) {
            bool more = true;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if ((int64_t)pr[q] < tend) xcur[pr[q] - t0] = pv[q];
                else more = false;
            }
            if (more) {
                for (int64_t k = kbase + 2 * POST_THREADS + tid; k < kend; k += POST_THREADS) {
                    const int64_t r = __ldg(rowidx + k);
                    if (r >= tend) break;
                    xcur[r - t0] = __ldcs(val + k);
                }
            }
        }
        __syncthreads();
        if (jj > 0) { // phase 2 of column jj-1: heavy entries x S, flattened over the CTA
            const int nh = n_heavy[jj & 1];
            const int items = nh * S;
            for (int it = tid; it < items; it += POST_THREADS) {
                const int s = it / nh, ent = it - s * nh;
                const HeavyEntry h = heavy[ent];
                const int64_t i = t0 + h.gene;
                Philox g;
                g.init(seed, (uint64_t)(gene0 + i) + (uint64_t)(cell0 + j) * 0x100000000ULL, 0x80000000u + (uint32_t)s);
                // x may not be exactly representable in float for non-integer inputs: re-read it
                __stcs(o + i + sstride * s, bn_nb_heavy_draw(h.r, h.M, g) + xprev[h.gene]);
            }
            __syncthreads();
            if (tid == 0) n_heavy[jj & 1] = 0;
        }
    }
}

static int post_launch(bn_ctx *ctx, int mode, const bn_mat *m, const double *BETA_vec, const double *size,
                       const double *mu, int64_t col0, int64_t ncol, int S, uint64_t seed, double *out)
{
    if (!ctx || !m || !BETA_vec || !size || !mu || !out)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "posterior: NULL argument (mu is required: the NB model has no default)") : BN_ERR_INVALID;
    if (col0 < 0 || ncol <= 0 || col0 + ncol > m->C) return bn_fail(ctx, BN_ERR_INVALID, "posterior: column block [%lld,+%lld) outside 0..%lld", (long long)col0, (long long)ncol, (long long)m->C);
    if (mode == POST_SAMPLE && S <= 0) return bn_fail(ctx, BN_ERR_INVALID, "posterior: S must be positive");
    cudaSetDevice(ctx->device);
    const size_t nout = (size_t)m->G * (size_t)ncol * (size_t)(mode == POST_SAMPLE ? S : 1);
    DevIn<double> b, sz, mm;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(sz.bind(ctx, size, (size_t)m->G));
    BN_TRY(mm.bind(ctx, mu, (size_t)m->G));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, out, nout));
    if (mode == POST_SAMPLE) {
        const int64_t tiles_s = (m->G + SAMPLE_TG - 1) / SAMPLE_TG;
        int ncs = POST_NCMAX;
        while (ncs > 1 && tiles_s * ((ncol + ncs - 1) / ncs) < 148 * 8) ncs >>= 1;
        if ((ncol + ncs - 1) / ncs > 65535) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: at most %d columns per call", 65535 * POST_NCMAX);
        dim3 grid_s((unsigned)tiles_s, (unsigned)((ncol + ncs - 1) / ncs));
        BN_CUDA(ctx, cudaFuncSetAttribute(k_post_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, SAMPLE_SMEM_BYTES));
        BN_LAUNCH(ctx, k_post_sample, grid_s, POST_THREADS, SAMPLE_SMEM_BYTES, m->G, m->colptr, m->rowidx, m->val, b.p, sz.p, mm.p, col0, ncol,
                  ncs, o.p, S, seed, m->gene0, m->cell0, m->G * ncol);
        BN_TRY(o.commit(ctx));
        return bn_sync(ctx);
    }
    static int ept = 0;
    if (!ept) {
        const char *e = getenv("BN_POST_EPT");
        ept = (e && atoi(e) == 4) ? 4 : 8;
    }
    const int64_t TG = (int64_t)POST_THREADS * ept;
    const int64_t tiles = (m->G + TG - 1) / TG;
    // long column walks amortise the per-CTA slice search and the size/mu loads; keep >= ~8 CTAs per SM
    int nc = POST_NCMAX;
    while (nc > 1 && tiles * ((ncol + nc - 1) / nc) < 148 * 8) nc >>= 1;
    if ((ncol + nc - 1) / nc > 65535) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: at most %d columns per call", 65535 * POST_NCMAX);
    dim3 grid((unsigned)tiles, (unsigned)((ncol + nc - 1) / nc));
    const int64_t sstride = m->G * ncol;
    const size_t smem = 3 * (size_t)TG * sizeof(double);
    const bool dense = (double)m->nnz > 0.2 * (double)m->G * (double)m->C;
#define POST_GO2(MODE_, EPT_, PF_)                                                                                   \
    do {                                                                                                             \
        BN_CUDA(ctx, cudaFuncSetAttribute(k_posterior<MODE_, EPT_, PF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        BN_LAUNCH(ctx, (k_posterior<MODE_, EPT_, PF_>), grid, POST_THREADS, smem, m->G, m->colptr, m->rowidx, m->val, b.p, sz.p, \
                  mm.p, col0, ncol, nc, o.p, S, seed, m->gene0, m->cell0, sstride);                                  \
    } while (0)
#define POST_GO(MODE_, EPT_)                  \
    do {                                      \
        if (dense) POST_GO2(MODE_, EPT_, 4);  \
        else POST_GO2(MODE_, EPT_, 2);        \
    } while (0)
    if (ept == 8) {
        if (mode == POST_MEAN) POST_GO(POST_MEAN, 8);
        else POST_GO(POST_MODE, 8);
    } else {
        if (mode == POST_MEAN) POST_GO(POST_MEAN, 4);
        else POST_GO(POST_MODE, 4);
    }
#undef POST_GO
#undef POST_GO2
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
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
// DownSampling: dense column-major in/out, one binomial draw per element, 16 B/element.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_downsample(int64_t G, int64_t C, const double *__restrict__ in,
                                                    const double *__restrict__ beta, uint64_t seed,
                                                    double *__restrict__ out)
{
    const int64_t total = G * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += gridDim.x * (int64_t)blockDim.x) {
        const double n = __ldcs(in + e);
        double r = 0.0;
        if (n > 0.0) {
            const int64_t j = e / G, i = e - j * G;
            Philox g;
            g.init(seed, (uint64_t)i + (uint64_t)j * 0x100000000ULL, 0x44533031u /* "DS01" */);
            r = bn_rbinom(g, n, __ldg(beta + j));
        }
        __stcs(out + e, r);
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
    const int64_t want = (G * C + 255) / 256;
    const int grid = (int)std::min<int64_t>(want, 148 * 32);
    BN_LAUNCH(ctx, k_downsample, grid, 256, 0, G, C, a.p, b.p, seed, o.p);
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
    if (nnz >= 4294967295LL) return fail(bn_fail(ctx, BN_ERR_UNSUPPORTED, "nnz >= 2^32 per shard"));
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
