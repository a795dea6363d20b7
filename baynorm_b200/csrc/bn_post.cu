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
// 0.93 -> 0.66 ms against the CTA-synchronous k_posterior above (kept behind BN_POST_V1 for A/B runs).
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

// ---------------------------------------------------------------------------------
// 3-D draws (Main_NB_Bay, src/bayNorm_main.cpp:1063-1139): S draws x + NB(size+x, tempmu) per
// gene x cell, out[i + G*(j + ncol*s)].
//
// Two kernels.  k_post_sample walks the matrix exactly like k_posterior (1024-gene tiles).  Per
// element it evaluates the NB parameters in FP64 (the reference's expression) and tabulates the first
// SAMPLE_K points of the CDF as 32-bit fixed-point thresholds in shared memory (one column of the
// table per thread: conflict-free).  If the table covers >= 95 % of the mass (all x = 0 entries of
// ordinary genes, ~97 % of a single-cell matrix) each of the S draws is one Philox word searched
// branch-free through the table with integer compares -- 3 instructions per level -- and stored
// straight to its output plane (every store instruction writes 256 contiguous doubles); a word
// beyond the table continues the recurrence sequentially.  Everything else (non-zero counts of
// expressed genes) is appended to a list in global memory and drawn by k_post_sample_heavy with the
// gamma-Poisson mixture the reference uses, one (entry, s) item per thread, so rejection loops never
// share a warp with the table path and the two code bodies do not evict each other from the
// instruction cache.
// Draw (gene, cell, s) is a pure function of (seed, global gene, global cell, s): table draws use
// Philox block s/4 word s%4 of the element's stream, heavy draws their own sub-stream 2^31 + s.
// ---------------------------------------------------------------------------------
#define SAMPLE_LIGHT_MAX 30.0f  // means above this never put enough mass on SAMPLE_K points
#define SAMPLE_LIGHT_MASS 0.85f // mass the table must cover (the rest: sequential tail); 30 / 0.85 measured best of a sweep (0.95: +16 %)
#define SAMPLE_K 32             // support points in the per-element CDF table

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// A Philox word at or beyond the last tabulated threshold: continue the pmf recurrence from k = SAMPLE_K - 1
// until the (2^32-scaled) CDF passes it.  The walk ends as well once p(k) no longer moves the float CDF: the
// word then lies in the last ~1e-7 of mass and k is at that quantile.  If it cannot even start -- p(K-1) is
// already negligible, i.e. the table is complete and the word fell into the rounding deficit of the float sum
// (probability ~1e-7) -- return -1: the caller takes the last tabulated point that carries mass.
__device__ __noinline__ float sample_tail(uint32_t u, float pk, float ck, float c1, float q)
{
    const float uf = (float)u;
    float kk = (float)(32 - 1);
    if (!(pk > ck * 3e-8f)) return -1.0f;
    while (uf >= ck && pk > ck * 3e-8f && kk < 4096.0f) {
        kk += 1.0f;
        pk *= fmaf(c1, __fdividef(1.0f, kk), q);
        ck += pk;
    }
    return kk;
}

struct HeavyItem {
    int32_t gene, col; // local to the launch: row of the shard, column relative to col0
    float r, M;
    double x;
};

// Every WARP is autonomous: it owns SW_GENES consecutive genes (4 per lane: stores stay 256 B
// contiguous per instruction) and walks the CTA's column chunk at its own pace -- own slice starts
// (binary-searched up front, two columns per lane, kept in registers and shuffled out), own count
// strip, own table columns, own block of the heavy list.  There is no __syncthreads() in the loop:
// a warp delayed by a tail walk never holds up the other seven.
#define SW_GENES 128
#define SW_HBLOCK 64 // heavy-list slots a warp reserves per global atomic
#define SW_PLANES 32 // planes staged per pass (one byte per draw)
#define SW_SMEM_PER_WARP (SAMPLE_K * 32 * 4 + SW_GENES * 8 + SW_PLANES * SW_GENES)

template <int SW_WARPS, int MINB>
__global__ void __launch_bounds__(SW_WARPS * 32, MINB)
    k_post_sample(int64_t G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                  const double *__restrict__ val, const double *__restrict__ beta, const double *__restrict__ size,
                  const double *__restrict__ mu, int64_t col0, int64_t ncol, int nc_per_cta, double *__restrict__ out,
                  int S, const __grid_constant__ PhiloxKeys keys, int64_t gene0, int64_t cell0, int64_t sstride,
                  HeavyItem *__restrict__ hlist, unsigned int *__restrict__ hcount, unsigned int hcap,
                  const uint32_t *__restrict__ blockptr, int64_t Ctot, float light_max, float light_mass)
{
    constexpr int SW_TG = SW_GENES * SW_WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xt = reinterpret_cast<double *>(smem_raw) + warp * SW_GENES;                                  // the warp's counts of the current column (0 = absent)
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem_raw + SW_WARPS * SW_GENES * 8) + warp * SAMPLE_K * 32; // [k][lane]: one conflict-free column per thread
    unsigned char *s_stage = smem_raw + SW_WARPS * (SW_GENES * 8 + SAMPLE_K * 32 * 4) + warp * (SW_PLANES * SW_GENES); // [plane][gene]: draws of the current column
    const int64_t w0 = (int64_t)blockIdx.x * SW_TG + (int64_t)warp * SW_GENES; // first gene of this warp
    const int64_t wend = w0 + SW_GENES;
    const int64_t jb = (int64_t)blockIdx.y * nc_per_cta;
    const int nc = (int)((jb + nc_per_cta < ncol ? jb + nc_per_cta : ncol) - jb);
    if (w0 >= G) return;
    // slice bounds of columns lane and lane + 32 of the chunk from the blocked column pointers (the warp's 128
    // genes are one row block), and their BETA
    const uint32_t *bp_lo = blockptr + (w0 / BN_ROWBLOCK) * Ctot + col0 + jb;
    const int64_t base = __ldg(colptr + col0 + jb); // 32-bit offsets from the chunk's first column below
    rowidx += base;
    val += base;
    uint32_t lo_a = 0, end_a = 0, lo_b = 0, end_b = 0;
    float beta_a = 0.0f, beta_b = 0.0f;
    if (lane < nc) {
        const uint32_t cs = (uint32_t)(__ldg(colptr + col0 + jb + lane) - base);
        lo_a = cs + __ldg(bp_lo + lane);
        end_a = cs + __ldg(bp_lo + Ctot + lane);
        beta_a = (float)__ldg(beta + col0 + jb + lane);
    }
    if (lane + 32 < nc) {
        const uint32_t cs = (uint32_t)(__ldg(colptr + col0 + jb + lane + 32) - base);
        lo_b = cs + __ldg(bp_lo + lane + 32);
        end_b = cs + __ldg(bp_lo + Ctot + lane + 32);
        beta_b = (float)__ldg(beta + col0 + jb + lane + 32);
    }
    // The draws are made from a float pmf, so the NB parameters size + x and
    // tempmu = (x + size)(mu - mu beta)/(size + mu beta) (:1118) are evaluated in float too: a relative 1e-7 in a
    // parameter is far below what any test of the draws can resolve, and it frees 8 registers and an FP64 division.
    float sz[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int64_t i = w0 + e * 32 + lane;
        sz[e] = (i < G) ? (float)__ldg(size + i) : 1.0f;
        m[e] = (i < G) ? (float)__ldg(mu + i) : 0.0f;
        xt[e * 32 + lane] = 0.0;
    }
    // this thread's column of the threshold table, as a 32-bit shared-memory address: entry k at tsa + SOFF(k)
    const uint32_t tsa = (uint32_t)__cvta_generic_to_shared(s_tab + lane);
    const uint32_t ssa = (uint32_t)__cvta_generic_to_shared(s_stage + lane); // draw (plane s, element e) at ssa + s * SW_GENES + e * 32
#define SOFF(k) ((k) * 32 * 4)
    unsigned hblock = 0, hused = SW_HBLOCK; // heavy-list block of this warp (warp-uniform)
    // slice of column 0 into registers
    int32_t pr = 0x7fffffff;
    double pv = 0.0;
    uint32_t kcur = __shfl_sync(0xffffffffu, lo_a, 0), kend = __shfl_sync(0xffffffffu, end_a, 0);
    if (nc > 0 && kcur + lane < kend) {
        pr = __ldg(rowidx + kcur + lane);
        pv = __ldcs(val + kcur + lane);
    }
    __syncwarp();
    for (int jj = 0; jj < nc; jj++) {
        // scatter column jj (prefetched; dense columns continue synchronously)
        {
            bool more = (int64_t)pr < wend;
            if (more) xt[pr - w0] = pv;
            uint32_t k = kcur + 32;
            while (__all_sync(0xffffffffu, more) && k < kend) {
                int32_t r2 = 0x7fffffff;
                double v2 = 0.0;
                if (k + lane < kend) {
                    r2 = __ldg(rowidx + k + lane);
                    v2 = __ldcs(val + k + lane);
                }
                more = (int64_t)r2 < wend;
                if (more) xt[r2 - w0] = v2;
                k += 32;
            }
        }
        __syncwarp();
        // prefetch the slice of column jj + 1
        pr = 0x7fffffff;
        pv = 0.0;
        if (jj + 1 < nc) {
            const int src = (jj + 1) & 31;
            kcur = __shfl_sync(0xffffffffu, (jj + 1 < 32) ? lo_a : lo_b, src);
            kend = __shfl_sync(0xffffffffu, (jj + 1 < 32) ? end_a : end_b, src);
            if (kcur + lane < kend) {
                pr = __ldg(rowidx + kcur + lane);
                pv = __ldcs(val + kcur + lane);
            }
        }
        const int64_t jl = jb + jj, j = col0 + jl;
        const float bf = __shfl_sync(0xffffffffu, (jj < 32) ? beta_a : beta_b, jj & 31);
        // The S draws of the column go through a byte-per-draw staging strip and leave as 1 KB runs per plane
        // (4 x 256 B, the warp's 128 genes): the dim = c(G, C, S) layout puts the planes G*ncol apart, and HBM
        // sustains 5.5 TB/s with 1 KB per plane visit against 4.0 TB/s with 256 B (bn_debug_write_peak).
        // 255 in the strip = not written here (big-mean element: other kernels; draw >= 255: stored directly).
        for (int sb = 0; sb < S; sb += SW_PLANES) { // one pass unless S > SW_PLANES
        const int nS = min(S - sb, SW_PLANES);
#pragma unroll 1
        for (int e = 0; e < 4; e++) {
            const int64_t i = w0 + e * 32 + lane;
            const bool in = i < G;
            const double xv = xt[e * 32 + lane];
            // size/mu of the current gene sit in sz[0], m[0]; the arrays are rotated at the end of the step so
            // that they stay in registers (a dynamic index would put them in local memory)
            const float mbf = m[0] * bf;
            const float rf = sz[0] + (float)xv, Mf = __fdividef(rf * (m[0] - mbf), sz[0] + mbf);
            const bool pos = Mf > 0.0f;
            {
                const float s_ = sz[0], m_ = m[0];
                sz[0] = sz[1], sz[1] = sz[2], sz[2] = sz[3], sz[3] = s_;
                m[0] = m[1], m[1] = m[2], m[2] = m[3], m[3] = m_;
            }
            double *o = out + G * jl + i + sstride * sb;
            // first SAMPLE_K points of the CDF, scaled by 2^32:  p(0) = (r/(r+M))^r,
            // p(k) = p(k-1) * q (r + k - 1)/k = p(k-1) * (q + q (r - 1)/k),  q = M/(r+M)
            const float q = __fdividef(Mf, rf + Mf), c1 = q * (rf - 1.0f);
            float p = 0.0f, ctop = 0.0f;
            if (in && Mf < light_max && pos) {
                p = __expf(-rf * log1pf(__fdividef(Mf, rf))) * 4294967296.0f;
                ctop = p;
                sts_u32(tsa, __float2uint_rz(ctop)); // the conversion saturates at 2^32 - 1
#pragma unroll
                for (int k = 1; k < SAMPLE_K; k++) {
                    p *= fmaf(c1, 1.0f / (float)k, q);
                    ctop += p;
                    sts_u32(tsa + SOFF(k), __float2uint_rz(ctop));
                }
            }
            const bool light = ctop >= light_mass * 4294967296.0f;
            const bool heavy = in && pos && !light;
            // append the heavy elements of this step to the warp's block of the global list; a new block
            // (one global atomic per SW_HBLOCK slots) is reserved when the current one cannot take them,
            // the slots left over are marked empty
            const unsigned hm = (sb == 0) ? __ballot_sync(0xffffffffu, heavy) : 0u;
            if (hm) {
                const unsigned n = (unsigned)__popc(hm);
                if (hused + n > SW_HBLOCK) {
                    if (hblock + SW_HBLOCK <= hcap) {
                        if (lane < (int)(SW_HBLOCK - hused)) hlist[hblock + hused + lane].gene = -1;
                        if (lane + 32 < (int)(SW_HBLOCK - hused)) hlist[hblock + hused + lane + 32].gene = -1;
                    }
                    if (lane == 0) hblock = atomicAdd(hcount, (unsigned)SW_HBLOCK);
                    hblock = __shfl_sync(0xffffffffu, hblock, 0);
                    hused = 0;
                }
                // a block past the capacity is dropped: hcount > hcap tells the host, which repeats the call with
                // worst-case lists (draws are pure functions of their coordinates, so the repeat is idempotent)
                if (heavy && hblock + SW_HBLOCK <= hcap) {
                    HeavyItem h;
                    h.gene = (int32_t)i;
                    h.col = (int32_t)jl;
                    h.r = rf;
                    h.M = Mf;
                    h.x = xv;
                    hlist[hblock + hused + __popc(hm & ((1u << lane) - 1u))] = h;
                }
                hused += n;
            }
            uint32_t sp = ssa + e * 32; // staging slot of (plane sb, this element)
            if (!light) {
                // mean 0 (rgamma(shape, scale = 0) = 0 -> rpois(0) = 0): the draw is x itself, staged as k = 0;
                // big-mean and out-of-range elements: 255 = not written by this kernel
                const uint32_t fill = (in && !pos) ? 0u : 255u;
                for (int s = 0; s < nS; s++) sts_u8(sp + s * SW_GENES, fill);
            } else {
                const uint32_t c0 = (uint32_t)(gene0 + i), c1p = (uint32_t)(cell0 + j);
                const uint32_t t31 = lds_u32(tsa + SOFF(SAMPLE_K - 1));
                double *po = o;
                // one search level: thresholds are non-decreasing, k = #{thresholds <= u}
#define SLEVEL(A, U, PROBE, STEP)                                                                                   \
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tld.shared.u32 t, [%0+%2];\n\tsetp.ge.u32 p, %1, t;\n\t"            \
                 "@p add.u32 %0, %0, %3;\n\t}"                                                                         \
                 : "+r"(A)                                                                                             \
                 : "r"(U), "n"(SOFF(PROBE)), "n"(SOFF(STEP)))
#define SKVAL(A) ((A - tsa) / SOFF(1))
                // slow path of one draw: beyond the table (~0.1 % of draws), or a draw that does not fit a byte
#define SDRAW(A, U, PLANE)                                                                                       \
    do {                                                                                                         \
        uint32_t k_ = SKVAL(A);                                                                                  \
        if (U >= t31) {                                                                                          \
            const float kk_ = sample_tail(U, p, ctop, c1, q);                                                    \
            if (kk_ >= 0.0f) k_ = (uint32_t)kk_;                                                                 \
            else { /* rounding deficit of a complete table: the last point that still carries mass */            \
                uint32_t a_ = tsa;                                                                               \
                const uint32_t u_ = t31 - 1u;                                                                    \
                SLEVEL(a_, u_, 15, 16);                                                                          \
                SLEVEL(a_, u_, 7, 8);                                                                            \
                SLEVEL(a_, u_, 3, 4);                                                                            \
                SLEVEL(a_, u_, 1, 2);                                                                            \
                SLEVEL(a_, u_, 0, 1);                                                                            \
                k_ = SKVAL(a_);                                                                                  \
            }                                                                                                    \
        }                                                                                                        \
        if (k_ >= 255u) {                                                                                        \
            __stcs(po + sstride * (PLANE), (double)k_ + xv);                                                     \
            k_ = 255u;                                                                                           \
        }                                                                                                        \
        sts_u8(sp + (PLANE) * SW_GENES, k_);                                                                     \
    } while (0)
                int s0 = 0;
                for (; s0 + 4 <= nS; s0 += 4) {
                    uint32_t u0, u1, u2, u3;
                    bn_philox_block(keys, c0, c1p, (uint32_t)((sb + s0) >> 2), 0u, u0, u1, u2, u3);
                    uint32_t a0 = tsa, a1 = tsa, a2 = tsa, a3 = tsa;
                    SLEVEL(a0, u0, 15, 16); SLEVEL(a1, u1, 15, 16); SLEVEL(a2, u2, 15, 16); SLEVEL(a3, u3, 15, 16);
                    SLEVEL(a0, u0, 7, 8);   SLEVEL(a1, u1, 7, 8);   SLEVEL(a2, u2, 7, 8);   SLEVEL(a3, u3, 7, 8);
                    SLEVEL(a0, u0, 3, 4);   SLEVEL(a1, u1, 3, 4);   SLEVEL(a2, u2, 3, 4);   SLEVEL(a3, u3, 3, 4);
                    SLEVEL(a0, u0, 1, 2);   SLEVEL(a1, u1, 1, 2);   SLEVEL(a2, u2, 1, 2);   SLEVEL(a3, u3, 1, 2);
                    SLEVEL(a0, u0, 0, 1);   SLEVEL(a1, u1, 0, 1);   SLEVEL(a2, u2, 0, 1);   SLEVEL(a3, u3, 0, 1);
                    const uint32_t umax = max(max(u0, u1), max(u2, u3));
                    if (umax < t31) { // all four inside the table (the common case; k < 32 fits the byte)
                        sts_u8(sp, SKVAL(a0));
                        sts_u8(sp + SW_GENES, SKVAL(a1));
                        sts_u8(sp + 2 * SW_GENES, SKVAL(a2));
                        sts_u8(sp + 3 * SW_GENES, SKVAL(a3));
                    } else {
                        SDRAW(a0, u0, 0);
                        SDRAW(a1, u1, 1);
                        SDRAW(a2, u2, 2);
                        SDRAW(a3, u3, 3);
                    }
                    sp += 4 * SW_GENES;
                    po += 4 * sstride;
                }
                if (s0 < nS) { // S is not a multiple of 4: the last, partial Philox block
                    uint32_t u[4];
                    bn_philox_block(keys, c0, c1p, (uint32_t)((sb + s0) >> 2), 0u, u[0], u[1], u[2], u[3]);
#pragma unroll
                    for (int w = 0; w < 3; w++) {
                        if (s0 + w < nS) {
                            uint32_t a0 = tsa;
                            SLEVEL(a0, u[w], 15, 16);
                            SLEVEL(a0, u[w], 7, 8);
                            SLEVEL(a0, u[w], 3, 4);
                            SLEVEL(a0, u[w], 1, 2);
                            SLEVEL(a0, u[w], 0, 1);
                            SDRAW(a0, u[w], w);
                        }
                    }
                }
#undef SDRAW
#undef SKVAL
#undef SLEVEL
            }
            __syncwarp(); // reconverge before the next element's table is written
        }
        // flush: plane by plane, the warp's 128 genes as four back-to-back 256 B stores
        {
            double xe[4];
            bool ine[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                xe[e] = xt[e * 32 + lane];
                ine[e] = w0 + e * 32 + lane < G;
            }
            double *po = out + G * jl + w0 + lane + sstride * sb;
            uint32_t sp = ssa;
            for (int s = 0; s < nS; s++) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const uint32_t kb = lds_u8(sp + e * 32);
                    if (kb != 255u && ine[e]) __stcs(po + e * 32, xe[e] + (double)kb);
                }
                sp += SW_GENES;
                po += sstride;
            }
        }
        __syncwarp(); // the strip is rewritten by the next pass / column
        }
#pragma unroll
        for (int e = 0; e < 4; e++) xt[e * 32 + lane] = 0.0; // own slots: ready for the next column's scatter
        __syncwarp();
    }
    // close the warp's last heavy-list block
    if (hused < SW_HBLOCK && hblock + SW_HBLOCK <= hcap) {
        if (lane < (int)(SW_HBLOCK - hused)) hlist[hblock + hused + lane].gene = -1;
        if (lane + 32 < (int)(SW_HBLOCK - hused)) hlist[hblock + hused + lane + 32].gene = -1;
    }
#undef SOFF
}

// ---------------------------------------------------------------------------------
// Elements the 32-point table cannot cover (non-zero counts of expressed genes, ~3 % of a
// single-cell matrix): one WARP per element tabulates up to 1024 points of the same CDF
// cooperatively -- each lane multiplies out 8 consecutive pmf ratios, a multiplicative warp scan
// chains the lanes, a second scan accumulates the CDF -- in FP64 (p(0) can be 1e-200), stores it
// as 32-bit thresholds in the warp's shared-memory strip and draws the S samples by binary search,
// lane s drawing sample s with the SAME Philox word the 32-point path would use.  Cost is fixed
// (no rejection loops, no divergence).  Elements whose first 1024 points hold < 99.9 % of the mass
// (means in the hundreds) go to a second list for the gamma-Poisson kernel below.
// ---------------------------------------------------------------------------------
#define HT_KMAX 1024
#define HT_BLOCK 256 // points per cooperative pass (8 per lane)
#define HT_WARPS 8
#define HT_GROUP 8 // list slots a warp takes per fetch

__global__ void __launch_bounds__(HT_WARPS * 32) k_post_sample_table(int64_t G, const HeavyItem *__restrict__ hlist,
                                                                     const unsigned int *__restrict__ hcount, int S,
                                                                     const __grid_constant__ PhiloxKeys keys, int64_t gene0,
                                                                     int64_t cell0, int64_t col0, int64_t sstride,
                                                                     double *__restrict__ out, HeavyItem *__restrict__ flist,
                                                                     unsigned int *__restrict__ fcount, unsigned int hcap,
                                                                     unsigned int *__restrict__ cursor)
{
    __shared__ uint32_t tab[HT_WARPS][HT_KMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *T = tab[warp];
    const unsigned n = min(*hcount, hcap);
    const double two32 = 4294967296.0;
    // groups of HT_GROUP slots are handed out through a global cursor (an element costs ~600 warp instructions,
    // so static striding would leave the last warps a long tail); the next group is fetched before this one is worked
    unsigned grp = 0;
    if (lane == 0) grp = atomicAdd(cursor, (unsigned)HT_GROUP);
    grp = __shfl_sync(0xffffffffu, grp, 0);
    HeavyItem mine;
    mine.gene = -1;
    if (lane < HT_GROUP && grp + lane < n) mine = hlist[grp + lane];
    while (grp < n) { // grp: the group held in `mine`
        const HeavyItem cur = mine;
        if (lane == 0) grp = atomicAdd(cursor, (unsigned)HT_GROUP);
        grp = __shfl_sync(0xffffffffu, grp, 0); // next group (>= n: nothing left, the loop ends after this pass)
        mine.gene = -1;
        if (lane < HT_GROUP && grp + lane < n) mine = hlist[grp + lane];
        unsigned todo = __ballot_sync(0xffffffffu, cur.gene >= 0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        HeavyItem h;
        h.gene = __shfl_sync(0xffffffffu, cur.gene, src);
        h.col = __shfl_sync(0xffffffffu, cur.col, src);
        h.r = __shfl_sync(0xffffffffu, cur.r, src);
        h.M = __shfl_sync(0xffffffffu, cur.M, src);
        h.x = __shfl_sync(0xffffffffu, cur.x, src);
        const double r = (double)h.r, M = (double)h.M;
        const double q = M / (r + M), c1 = q * (r - 1.0);
        // p(0) * 2^32 = 2^(32 - r log2(1 + M/r)) (everything below is scaled by 2^32).  The parameters are floats, so
        // float logarithm/exponential cores lose nothing (p(0) already moves by r * 6e-8 relative per ulp of M);
        // the exponent is split off in double so that p(0) down to 1e-300 stays representable.
        double p0;
        {
            const double t = 32.0 - r * 1.4426950408889634 * (double)log1pf(h.M / h.r);
            const double nf = floor(t);
            const int ne = (int)nf;
            p0 = (ne < -1000) ? 0.0 : (double)exp2f((float)(t - nf)) * __hiloint2double((1023 + ne) << 20, 0);
        }
        double carry_p = 1.0, carry_c = 0.0;
        int K = 0;
        for (int blk = 0; blk < HT_KMAX / HT_BLOCK; blk++) {
            const int kb = blk * HT_BLOCK + lane * 8;
            double pl[8], prod = 1.0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = kb + j;
                const float kf = (float)k;
                double y = (double)__fdividef(1.0f, kf); // 1/k: float reciprocal + one Newton step in double
                y = fma(y, fma(-(double)kf, y, 1.0), y);
                const double t = (k == 0) ? p0 : fma(c1, y, q);
                prod *= t;
                pl[j] = prod;
            }
            // exclusive multiplicative scan of the lane products, then chain to the previous block
            double incl = prod;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl *= v;
            }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0;
            excl *= carry_p;
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                sum += pl[j] * excl; // p(kb + j)
                pl[j] = sum;
            }
            double incs = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, incs, o);
                if (lane >= o) incs += v;
            }
            const double base = incs - sum + carry_c;
#pragma unroll
            for (int j = 0; j < 8; j++) T[kb + j] = __double2uint_rz(base + pl[j]); // saturates at 2^32 - 1
            carry_p = __shfl_sync(0xffffffffu, incl, 31) * carry_p;
            carry_c += __shfl_sync(0xffffffffu, incs, 31);
            K += HT_BLOCK;
            if (carry_c >= 0.9998 * two32) break;
        }
        __syncwarp();
        if (!(carry_c >= 0.999 * two32)) { // mass beyond 1024 points: gamma-Poisson kernel
            if (lane == 0) flist[atomicAdd(fcount, 1u)] = h;
            continue;
        }
        const uint32_t c0 = (uint32_t)(gene0 + h.gene), c1p = (uint32_t)(cell0 + col0 + h.col);
        double *o = out + (int64_t)h.gene + G * (int64_t)h.col;
        // A table whose last point carries no mass any more is complete: its top is 2^32 only up to the float
        // error of p(0) (~1e-6), so the Philox word is rescaled to the tabulated total -- an exact renormalisation.
        const bool complete = !(carry_p > carry_c * 1e-12);
        const double uscale = complete ? carry_c * (1.0 / two32) : 1.0;
        for (int s = lane; s < S; s += 32) {
            uint32_t w[4];
            bn_philox_block(keys, c0, c1p, (uint32_t)(s >> 2), 0u, w[0], w[1], w[2], w[3]);
            uint32_t u = (s & 2) ? ((s & 1) ? w[3] : w[2]) : ((s & 1) ? w[1] : w[0]);
            if (complete) u = __double2uint_rd((double)u * uscale);
            int lo = 0; // k = #{thresholds <= u}: branch-free lower bound over the K tabulated points
            for (int step = 1 << (31 - __clz(K)); step > 0; step >>= 1) {
                const int idx = lo + step;
                if (idx <= K && u >= T[idx - 1]) lo = idx;
            }
            if (lo == K && complete) { // u == top exactly (rounding): the last point that carries mass
                const uint32_t u2 = T[K - 1] - 1u;
                lo = 0;
                for (int step = 1 << (31 - __clz(K)); step > 0; step >>= 1) {
                    const int idx = lo + step;
                    if (idx <= K && u2 >= T[idx - 1]) lo = idx;
                }
            }
            double kk = (double)lo;
            if (lo == K) { // beyond an incomplete table (< 0.1 % of draws): continue the recurrence
                double pk = carry_p, ck = carry_c;
                const double ud = (double)u;
                kk -= 1.0;
                while (ud >= ck && pk > ck * 1e-17 && kk < 1.0e6) {
                    kk += 1.0;
                    pk *= fma(c1, 1.0 / kk, q);
                    ck += pk;
                }
            }
            __stcs(o + sstride * s, kk + h.x);
        }
        __syncwarp(); // the strip is rewritten by the next element
      }
    }
}

// heavy entries x S, one item per thread (items of one entry are neighbours: same gamma shape)
__global__ void __launch_bounds__(256) k_post_sample_heavy(int64_t G, const HeavyItem *__restrict__ hlist,
                                                           const unsigned int *__restrict__ hcount, int S, uint64_t seed,
                                                           int64_t gene0, int64_t cell0, int64_t col0, int64_t sstride,
                                                           double *__restrict__ out, unsigned int *__restrict__ counters)
{
    // last kernel of a chunk: fold the chunk's slot count into the running maximum the host checks
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(counters + 2, counters[0]);
    const int64_t items = (int64_t)(*hcount) * S;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += gridDim.x * (int64_t)blockDim.x) {
        const int64_t ent = it / S;
        const int s = (int)(it - ent * S);
        const HeavyItem h = hlist[ent];
        Philox g;
        g.init(seed, (uint64_t)(gene0 + h.gene) + (uint64_t)(cell0 + col0 + h.col) * 0x100000000ULL, 0x80000000u + (uint32_t)s);
        __stcs(out + (int64_t)h.gene + G * (int64_t)h.col + sstride * s, bn_nb_heavy_draw(h.r, h.M, g) + h.x);
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
        constexpr int SW_WARPS = 8; // 8 warps x 9 KB of shared memory: 3 CTAs per SM
        const int64_t SW_TG = (int64_t)SW_GENES * SW_WARPS;
        const int64_t tiles_s = (m->G + SW_TG - 1) / SW_TG;
        const int sms = ctx->prop.multiProcessorCount;
        // Capacity of the list of elements the 32-point table cannot cover.  Worst case: every element, plus the
        // slots warps leave empty when they close a block (fewer than the elements they appended).  First attempt:
        // 4 entries per expected non-zero of the chunk (single-cell data: ~0.4 per non-zero); if a chunk overflows
        // (the kernels then drop what does not fit and leave the count above the capacity) the whole call is
        // repeated with worst-case lists.
        const double rho = (double)m->nnz / ((double)m->G * (double)m->C);
        BN_TRY(bn_mat_ensure_blockptr(ctx, m));
        const float light_max = SAMPLE_LIGHT_MAX, light_mass = SAMPLE_LIGHT_MASS;
        for (int attempt = 0; attempt < 2; attempt++) {
            const double per_elem = attempt == 0 ? std::min(2.0, std::max(0.125, 6.0 * rho)) : 2.0;
            const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncol, (int64_t)(((int64_t)1 << 30) / (sizeof(HeavyItem) * per_elem * (double)m->G))));
            // CTAs of any chunk: tiles * ceil(chunk / 64) when 64 columns per CTA already give 64 CTAs per SM; otherwise
            // the halving stops below 2 * 64 per SM (+ one row of tiles) or at 4 columns per CTA
            const int64_t ctas64 = tiles_s * ((chunk + POST_NCMAX - 1) / POST_NCMAX), ctas4 = tiles_s * ((chunk + 3) / 4);
            const size_t max_ctas = (size_t)std::max<int64_t>(ctas64, std::min<int64_t>(ctas4, (int64_t)sms * 128 + tiles_s));
            const size_t cap = (((size_t)(per_elem * (double)m->G * (double)chunk) + SW_HBLOCK - 1) / SW_HBLOCK + (size_t)SW_WARPS * max_ctas) * SW_HBLOCK; // whole blocks
            if (cap > 0xfffffff0u) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: element list too long");
            struct { HeavyItem *p; } hl, fl; // context scratch: kept across calls (see bn_ctx_scratch)
            DevBuf<unsigned int> hc; // per chunk: {list slots, fallback entries, -, table-kernel cursor}; hc[2] = max slots over the chunks
            BN_TRY(bn_ctx_scratch(ctx, 0, cap * sizeof(HeavyItem), (void **)&hl.p));
            BN_TRY(bn_ctx_scratch(ctx, 1, cap * sizeof(HeavyItem), (void **)&fl.p));
            BN_TRY(hc.alloc(ctx, 4));
            BN_CUDA(ctx, cudaMemsetAsync(hc.p, 0, 4 * sizeof(unsigned int), ctx->stream));
            for (int64_t c0 = 0; c0 < ncol; c0 += chunk) {
                const int64_t nc_ = std::min<int64_t>(chunk, ncol - c0);
                int ncs = POST_NCMAX; // columns per CTA: short walks keep the columns in flight (and their 20 planes) close together
                while (ncs > 4 && tiles_s * ((nc_ + ncs - 1) / ncs) < (int64_t)sms * 64) ncs >>= 1;
                if (ctx->tune[BN_TUNE_SAMPLE_NCS] > 0) ncs = std::max(1, std::min(POST_NCMAX, ctx->tune[BN_TUNE_SAMPLE_NCS]));
                if ((nc_ + ncs - 1) / ncs > 65535) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: at most %d columns per call", 65535 * POST_NCMAX);
                dim3 grid_s((unsigned)tiles_s, (unsigned)((nc_ + ncs - 1) / ncs));
                BN_CUDA(ctx, cudaMemsetAsync(hc.p, 0, 2 * sizeof(unsigned int), ctx->stream));
                BN_CUDA(ctx, cudaMemsetAsync(hc.p + 3, 0, sizeof(unsigned int), ctx->stream));
                // out is addressed relative to the chunk's first column; planes stay G*ncol apart
#define SAMPLE_ARGS                                                                                                    \
    m->G, m->colptr, m->rowidx, m->val, b.p, sz.p, mm.p, col0 + c0, nc_, ncs, o.p + m->G * c0, S, bn_philox_keys(seed), m->gene0, \
        m->cell0, m->G * ncol, hl.p, hc.p, (unsigned)cap, m->blockptr, m->C, light_max, light_mass
                const size_t smem_s = (size_t)SW_WARPS * SW_SMEM_PER_WARP;
                BN_CUDA(ctx, cudaFuncSetAttribute(k_post_sample<SW_WARPS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
                BN_LAUNCH(ctx, (k_post_sample<SW_WARPS, 3>), grid_s, SW_WARPS * 32, smem_s, SAMPLE_ARGS);
#undef SAMPLE_ARGS
                BN_LAUNCH(ctx, k_post_sample_table, sms * 4, HT_WARPS * 32, 0, m->G, hl.p, hc.p, S, bn_philox_keys(seed), m->gene0, m->cell0,
                          col0 + c0, m->G * ncol, o.p + m->G * c0, fl.p, hc.p + 1, (unsigned)cap, hc.p + 3);
                BN_LAUNCH(ctx, k_post_sample_heavy, sms * 8, 256, 0, m->G, fl.p, hc.p + 1, S, seed, m->gene0, m->cell0, col0 + c0, m->G * ncol,
                          o.p + m->G * c0, hc.p);
            }
            unsigned int *h_max = reinterpret_cast<unsigned int *>(ctx->pinned + 32);
            BN_CUDA(ctx, cudaMemcpyAsync(h_max, hc.p + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            BN_TRY(bn_sync(ctx));
            if (*h_max <= cap) break;
            if (attempt == 1) return bn_fail(ctx, BN_ERR_CUDA, "posterior: element list overflow with worst-case capacity");
        }
        BN_TRY(o.commit(ctx));
        return bn_sync(ctx);
    }
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
#define POSTW_ARGS m->G, m->colptr, m->rowidx, m->val, b.p, sz.p, mm.p, col0, ncol, ncw, o.p, m->blockptr, m->C
    // slices prefetched per column: 3 x 32 non-zeros for >= 20 % dense matrices (5 x 32 spills at 64 registers), 32 otherwise
    if (mode == POST_MEAN) {
        if (densew) BN_LAUNCH(ctx, (k_posterior_w<POST_MEAN, 3, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
        else BN_LAUNCH(ctx, (k_posterior_w<POST_MEAN, 1, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
    } else {
        if (densew) BN_LAUNCH(ctx, (k_posterior_w<POST_MODE, 3, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
        else BN_LAUNCH(ctx, (k_posterior_w<POST_MODE, 1, 4, 4>), gridw, PW_WARPS * 32, 0, POSTW_ARGS);
    }
#undef POSTW_ARGS
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
