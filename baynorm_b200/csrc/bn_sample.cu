// bn_sample.cu -- 3-D posterior draws (Main_NB_Bay, src/bayNorm_main.cpp:1063-1139): S draws
// x + NB(size = phi_i + x, mean = tempmu) per gene x cell, out[i + G*(j + ncol*s)], tempmu = (x + phi)(mu - mu beta)/(phi + mu beta)
// (:1118-1121).
//
// The draw (gene, cell, s) is the inversion of ONE Philox4x32-10 word -- block s/4, word s%4 of the stream with counter
// (global gene, global cell) and key = seed -- through the element's CDF, so it is a pure function of (seed, gene, cell, s):
// independent of S, tiling, sharding and GPU count.  At HBM speed a B200 has ~45 issue slots per draw, so everything is
// organised around instructions per draw:
//
//  * x = 0 elements (92 % of a single-cell matrix) are drawn by k_post_sample_main.  A warp owns 128 consecutive genes and
//    walks a chunk of columns on its own (no CTA barrier).  Its genes are dealt to the four "slots" (lane = gene) in the
//    order of their posterior spread (k_sample_perm), so the 32 lanes of a slot need CDF tables of similar length: the
//    table is tabulated 8, 16, 32 or SM_KCAP points deep until every lane holds 1 - 2^-12 of its mass (warp vote), and
//    the S draws are branch-free binary searches of exactly that depth with integer compares.
//  * Genes whose x = 0 elements need more than SM_KCAP points (small size with a moderate mean: ~12 % of the genes) would
//    make their whole slot walk window after window; their x = 0 elements are drawn FIRST, by k_hz_sample, thread per
//    (gene, column) with a table of the gene's class (64-512 points), into a byte side buffer the main kernel copies
//    into its strip.
//  * x > 0 elements (8 %) have size phi + x and a mean (phi + x)/phi times larger: tables of 64-512 points.  Drawn inside
//    the warps above they would dictate the table length of their 31 neighbours, so they are drawn FIRST, by
//    k_nz_sample: every stored entry of the column chunk is classified by an upper bound of its 1 - 2^-12 quantile, and
//    each class runs thread-per-element with a table of its own size -- full, homogeneous warps.  The draws land in a
//    side buffer (uint16 per draw, in CSC order) that the main kernel picks up with the column's non-zeros.
//  * All S draws of a column leave through a byte-per-draw strip and are flushed plane by plane as 1 KB runs (the
//    dim = c(G, C, S) layout puts the planes G*ncol apart; HBM sustains 5.5 TB/s at 1 KB per plane visit against 4.0 at
//    256 B).  Values >= 255 are stored directly; elements beyond 512 table points go to a list for the gamma-Poisson
//    kernel (the mixture the reference itself uses), which writes them after the main kernel.
#include <math.h>
#include <stdlib.h>

#include "bn_internal.cuh"
#include "bn_rng.cuh"

#define SM_GENES 128  // genes per warp: one row block of the blocked column pointers
#define SM_PLANES 28  // planes staged per pass (one byte per draw)
#define SM_STRIDE 7   // 32-bit words per gene in the strip: odd => conflict-free for lane = gene
#define SM_NCMAX 64   // columns walked by one CTA
#define SM_KMAX 512   // table points beyond which an element goes to the gamma-Poisson list
#define SM_HBLOCK 64  // deferred-list slots a warp reserves per global atomic
#define SM_TOPF 4293918720.0f // (1 - 2^-12) * 2^32: a table holding this much mass is "complete"
#define SM_LIGHT_MASS 0.85f   // mass 32 points must hold (at the gene's largest mean) for its x = 0 elements to stay in the main kernel
#define SM_LIGHT_MAX 30.0f    // ... and that mean must stay below this
#define SM_P0_MAXEXP 80.0f    // -log p(0) beyond which the element is drawn by the gamma-Poisson kernel

// 1/k correctly rounded (the value of the compile-time constant 1.0f / k): every kernel tabulates the SAME thresholds
__constant__ float c_rcp[1025] = {
#include "bn_rcp_table.inc"
};

struct HeavyItem {
    int32_t gene, col; // local to the launch: row of the shard, column relative to the chunk
    float r, M;
    double x;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v)); }

// Number of CDF points that hold 1 - 2^-12 of the mass of NB(size r, mean M): the normal-theory value
// M + 3.5 sd + 2 times a cubic correction in (log M, log r) fitted to the exact quantile over r in [0.03, 300],
// M in [0.02, 600] (residual -11 % .. +16 % for >= 8 points).  Only scheduling depends on it (which kernel and which table
// size draws an element): words beyond a table that turns out short of its mass walk the recurrence on from its end.
__device__ __forceinline__ float nb_points_est(float r, float M)
{
    const float x = __logf(M), y = __logf(r);
    const float base = M + 3.5f * sqrtf(M + __fdividef(M * M, r)) + 2.0f;
    const float poly = 4.92271701e-01f + x * (7.04511583e-02f + x * (-1.26723009e-02f + x * 5.61575493e-04f)) +
                       y * (-2.06393332e-01f + y * (2.61045443e-02f - y * 4.97898745e-04f)) +
                       x * y * (-2.03019819e-02f + x * 2.56334976e-03f + y * 3.66568844e-04f);
    return base * __expf(poly);
}

// ---------------------------------------------------------------------------------
// Per 128-gene block: genes in the order of the table length their x = 0 elements need at an average BETA.
// perm[block * 128 + rank] = offset of the gene inside its block.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample_perm(int64_t G, const double *__restrict__ size, const double *__restrict__ mu,
                                                     const float *__restrict__ beta_min, uint8_t *__restrict__ perm)
{
    __shared__ float keys[8][SM_GENES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t blk = (int64_t)blockIdx.x * 8 + warp;
    const int64_t w0 = blk * SM_GENES;
    if (w0 >= G) return;
    const float bf = beta_min[0]; // the smallest BETA of the whole vector: the order only has to be roughly right
    float key[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int64_t i = w0 + e * 32 + lane;
        float k = 3.0e38f; // past the end of the matrix: sorted last
        if (i < G) {
            const float s = (float)__ldg(size + i), m = (float)__ldg(mu + i), mb = m * bf;
            const float M = __fdividef(s * (m - mb), s + mb);
            k = (M > 0.0f) ? nb_points_est(s, M) : 0.0f;
            if (!(k == k)) k = 0.0f;
        }
        key[e] = k;
        keys[warp][e * 32 + lane] = k;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int a = e * 32 + lane;
        int rank = 0;
        for (int b = 0; b < SM_GENES; b++) {
            const float kb = keys[warp][b];
            rank += (kb < key[e] || (kb == key[e] && b < a)) ? 1 : 0;
        }
        perm[w0 + rank] = (uint8_t)a;
    }
}

// Thresholds are T(k) = rz(2^32 CDF(k)), the CDF accumulated in float by the recurrence
//     p(0) = (r/(r+M))^r,   p(k) = p(k-1) (q + q (r-1) / k)   with 1/k correctly rounded,   q = M/(r+M),
// and the draw for a Philox word u is the first k with u < T(k).  Every kernel tabulates exactly these thresholds, so
// which kernel (or how deep a table) drew an element never shows in the result: a gene may be "heavy" in one column
// chunk and "light" in another.  A word no table resolves -- beyond the tabulated points (2^-12 of the draws), or in the
// rounding deficit 2^32 - T(inf) of the float sum (~1e-7) -- takes the same recurrence from k = 0: the walk ends at the
// first k whose pmf no longer moves the sum.
__device__ __noinline__ uint32_t sample_from0(uint32_t u, float p0, float c1, float q)
{
    float pk = p0, ck = p0, kk = 0.0f;
    if (u < __float2uint_rz(ck)) return 0u;
    while (pk > ck * 3e-8f && kk < 60000.0f) {
        kk += 1.0f;
        pk *= fmaf(c1, __frcp_rn(kk), q);
        ck += pk;
        if (u < __float2uint_rz(ck)) break;
    }
    return (uint32_t)kk;
}

// The same recurrence continued from the last tabulated point kk (pmf pk, scaled CDF ck, T(kk) <= u): what a table that
// holds less than all of the mass does with the words beyond it.  Bit-identical to sample_from0 passing through kk, so
// the caller must only use it while pk still moves the sum at kk (otherwise: sample_from0).
__device__ __noinline__ uint32_t sample_walk(uint32_t u, float pk, float ck, float c1, float q, float kk)
{
    while (pk > ck * 3e-8f && kk < 60000.0f) {
        kk += 1.0f;
        pk *= fmaf(c1, __frcp_rn(kk), q);
        ck += pk;
        if (u < __float2uint_rz(ck)) break;
    }
    return (uint32_t)kk;
}

// ---------------------------------------------------------------------------------
// Thread-per-element sampling with a table of KT points per thread ([k][thread]: conflict-free), shared by the kernels
// for the stored entries (k_nz_sample) and for the x = 0 elements of heavy genes (k_hz_sample).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void nb_params(double xv, float sz, float m, float bf, float &rf, float &Mf)
{
    const float mbf = m * bf;
    rf = sz + (float)xv;
    Mf = __fdividef(rf * (m - mbf), sz + mbf);
}
// table class of an element: 0..4 -> 32 << class points, 5 -> gamma-Poisson list
#define NB_CLASSES 6
__device__ __forceinline__ int nb_class(float rf, float Mf)
{
    if (!(Mf > 0.0f)) return 0; // mean 0: every draw is x itself, any table will do
    // p(0) = exp(-r log1p(M/r)) must stay a normal float after the 2^32 scaling (the table recurrence starts from it)
    if (rf * log1pf(__fdividef(Mf, rf)) > SM_P0_MAXEXP) return NB_CLASSES - 1;
    const float k = nb_points_est(rf, Mf);
    return k <= 32.0f ? 0 : k <= 64.0f ? 1 : k <= 128.0f ? 2 : k <= 256.0f ? 3 : k <= (float)SM_KMAX ? 4 : 5;
}

// number of table entries <= u among 2^LEVELS entries at stride `stride` bytes from address `base`
template <int LEVELS> __device__ __forceinline__ uint32_t table_count(uint32_t base, uint32_t stride, uint32_t u)
{
    uint32_t lo = 0;
#pragma unroll
    for (int l = LEVELS - 1; l >= 0; l--) {
        const uint32_t half = 1u << l;
        if (u >= lds_u32(base + (lo + half - 1u) * stride)) lo += half;
    }
    return lo;
}

// KT points of the element's CDF in the thread's table column, then draws of planes [s_begin, s_end) (s_begin a multiple of
// 4): emit(s, k) receives them.  Two Philox blocks -- eight searches -- are in flight together: these kernels hold few
// warps per SM (the tables fill shared memory), so the latency of the dependent shared-memory probes has to be covered
// by independent searches of the same thread.
template <int KT, int STRIDE> struct NbTable {
    static constexpr int LEVELS = (KT == 32) ? 5 : (KT == 64) ? 6 : (KT == 128) ? 7 : (KT == 256) ? 8 : 9;
    uint32_t tsa, tlast;
    float p0, c1, q, pl, cl; // pl, cl: pmf and scaled CDF at the last tabulated point
    __device__ __forceinline__ void build(uint32_t tsa_, float rf, float Mf)
    {
        tsa = tsa_;
        q = __fdividef(Mf, rf + Mf);
        c1 = q * (rf - 1.0f);
        p0 = __expf(-rf * log1pf(__fdividef(Mf, rf))) * 4294967296.0f;
        float p = p0, ctop = p0;
        sts_u32(tsa, __float2uint_rz(ctop)); // the conversion saturates at 2^32 - 1
#pragma unroll 8
        for (int kk = 1; kk < KT; kk++) {
            p *= fmaf(c1, c_rcp[kk], q);
            ctop += p;
            sts_u32(tsa + kk * STRIDE, __float2uint_rz(ctop));
        }
        tlast = lds_u32(tsa + (KT - 1) * STRIDE);
        pl = p;
        cl = ctop;
    }
    template <typename Emit>
    __device__ __forceinline__ void draws(const PhiloxKeys &keys, uint32_t c0, uint32_t c1p, int s_begin, int s_end, Emit emit) const
    {
        for (int s0 = s_begin; s0 < s_end; s0 += 8) {
            uint32_t u[8], kd[8];
            bn_philox_block(keys, c0, c1p, (uint32_t)(s0 >> 2), 0u, u[0], u[1], u[2], u[3]);
            bn_philox_block(keys, c0, c1p, (uint32_t)(s0 >> 2) + 1u, 0u, u[4], u[5], u[6], u[7]);
#pragma unroll
            for (int w = 0; w < 8; w++) kd[w] = 0;
#pragma unroll
            for (int l = LEVELS - 1; l >= 0; l--) { // level by level across the eight searches
                const uint32_t half = 1u << l;
#pragma unroll
                for (int w = 0; w < 8; w++)
                    if (u[w] >= lds_u32(tsa + (kd[w] + half - 1u) * STRIDE)) kd[w] += half;
            }
#pragma unroll
            for (int w = 0; w < 8; w++)
                if (s0 + w < s_end)
                    emit(s0 + w, u[w] < tlast ? kd[w] : (pl > cl * 3e-8f ? sample_walk(u[w], pl, cl, c1, q, (float)(KT - 1)) : sample_from0(u[w], p0, c1, q)));
        }
    }
};

// ---------------------------------------------------------------------------------
// Stored entries of the column chunk: classification (two passes: per-column class counts, scan, fill -- no atomics, the
// list order is deterministic) and sampling.
// ---------------------------------------------------------------------------------
struct NzItem {
    uint32_t krel; // stored entry: position in the CSC arrays relative to the chunk's first entry;
                   // x = 0 element of a heavy gene: NZ_PAIR | row of the gene in the byte side buffer
    int32_t col;   // column relative to the chunk
};
#define NZ_PAIR 0x80000000u

// Thread per element, NZC_PER_CTA elements per CTA: the nnz stored entries of the chunk, then the nheavy x ncol (heavy gene,
// column) pairs -- the x = 0 elements of genes the main kernel does not draw; a pair that turns out to be a stored entry
// (looked up in the column's slice of the gene's row block) is dropped.  Every element is classified on its own: most columns
// of a heavy gene need far fewer points than its worst column.
// FILL = false: cnt[c * ncta + cta] = elements of class c this CTA saw.  FILL = true: cnt holds the list position of the
// CTA's first element of class c (k_nz_offsets); the elements take the next slots in whatever order the shared-memory
// counters hand out -- the list order only decides which thread draws an element.
#define NZC_PER_CTA 2048
template <bool FILL>
__global__ void __launch_bounds__(256) k_nz_classify(int64_t ncol, int64_t col0, int64_t ncta, int64_t nheavy, const int32_t *__restrict__ hgenes,
                                                     const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                     const double *__restrict__ val, const double *__restrict__ beta,
                                                     const double *__restrict__ size, const double *__restrict__ mu,
                                                     const uint32_t *__restrict__ blockptr, int64_t Ctot, unsigned int *__restrict__ cnt,
                                                     NzItem *__restrict__ list)
{
    __shared__ unsigned int s_cnt[NB_CLASSES];
    __shared__ int64_t s_span[2]; // columns (relative to the chunk) of the CTA's first and last stored entry
    const int64_t kfirst = __ldg(colptr + col0), nnz = __ldg(colptr + col0 + ncol) - kfirst, nall = nnz + nheavy * ncol;
    if (threadIdx.x < NB_CLASSES) s_cnt[threadIdx.x] = FILL ? cnt[threadIdx.x * ncta + blockIdx.x] : 0u;
    const int64_t vb = (int64_t)blockIdx.x * NZC_PER_CTA;
    if (threadIdx.x < 2 && vb < nnz) {
        // the CTA's entries are consecutive: their columns lie between those of the first and the last one (two searches per
        // CTA instead of one 11-step search per entry; a CTA spans one or two columns of a single-cell matrix)
        const int64_t k = kfirst + (threadIdx.x == 0 ? vb : (vb + NZC_PER_CTA - 1 < nnz ? vb + NZC_PER_CTA - 1 : nnz - 1));
        int64_t lo = 0, hi = ncol;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(colptr + col0 + mid) <= k) lo = mid;
            else hi = mid;
        }
        s_span[threadIdx.x] = lo;
    }
    __syncthreads();
#pragma unroll 2
    for (int t = threadIdx.x; t < NZC_PER_CTA; t += 256) {
        const int64_t v = vb + t;
        if (v >= nall) break;
        NzItem it;
        int32_t i;
        double xv = 0.0;
        if (v < nnz) {
            const int64_t k = kfirst + v;
            // column of entry k: last jl with colptr[col0 + jl] <= k
            int64_t lo = s_span[0], hi = s_span[1] + 1;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(colptr + col0 + mid) <= k) lo = mid;
                else hi = mid;
            }
            i = __ldg(rowidx + k);
            xv = __ldg(val + k);
            it.krel = (uint32_t)v;
            it.col = (int32_t)lo;
        } else {
            const int64_t pr = v - nnz, h = pr / ncol;
            i = __ldg(hgenes + h);
            it.krel = NZ_PAIR | (uint32_t)h;
            it.col = (int32_t)(pr - h * ncol);
            // stored? the column's entries of the gene's 128-row block are a run of ~10 sorted row indices
            const int64_t j = col0 + it.col, cs = __ldg(colptr + j);
            const uint32_t *bp = blockptr + (int64_t)(i / BN_ROWBLOCK) * Ctot + j;
            bool stored = false;
            for (int64_t k = cs + __ldg(bp), ke = cs + __ldg(bp + Ctot); k < ke; k++) {
                const int32_t r = __ldg(rowidx + k);
                if (r >= i) {
                    stored = r == i;
                    break;
                }
            }
            if (stored) continue;
        }
        float rf, Mf;
        nb_params(xv, (float)__ldg(size + i), (float)__ldg(mu + i), (float)__ldg(beta + col0 + it.col), rf, Mf);
        int cls = nb_class(rf, Mf);
        // a pair is always drawn here (the gamma-Poisson list is for stored entries and for genes the main kernel owns): the
        // estimate is a fitted curve, not a monotone bound, so an element may come out above its gene's class -- the 512-point
        // table then simply sends more words down the recurrence
        if (v >= nnz && cls == NB_CLASSES - 1) cls = NB_CLASSES - 2;
        const unsigned slot = atomicAdd(&s_cnt[cls], 1u);
        if (FILL) list[slot] = it;
    }
    if (!FILL) {
        __syncthreads();
        if (threadIdx.x < NB_CLASSES) cnt[threadIdx.x * ncta + blockIdx.x] = s_cnt[threadIdx.x];
    }
}

// exclusive scan of cnt[c * ncta + cta] in (class, CTA) order, in place; offsets[c] = start of class c, offsets[6] = total
__global__ void __launch_bounds__(1024) k_nz_offsets(int64_t n /* = NB_CLASSES * ncta */, int64_t ncta, unsigned int *__restrict__ cnt,
                                                     unsigned long long *__restrict__ offsets)
{
    __shared__ unsigned int wsum[32];
    __shared__ unsigned int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t k = base + threadIdx.x;
        const unsigned int v = k < n ? cnt[k] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned int before = carry;
        for (int w = 0; w < warp; w++) before += wsum[w];
        const unsigned int excl = before + inc - v;
        if (k < n) {
            cnt[k] = excl;
            if (k % ncta == 0) offsets[k / ncta] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[NB_CLASSES] = carry;
}

template <int KT, int THREADS>
__global__ void __launch_bounds__(THREADS) k_nz_sample(int cls, const NzItem *__restrict__ list,
                                                       const unsigned long long *__restrict__ offsets, int64_t G, int64_t ncol, int64_t col0,
                                                       const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                       const double *__restrict__ val, const int32_t *__restrict__ hgenes,
                                                       const double *__restrict__ beta, const double *__restrict__ size,
                                                       const double *__restrict__ mu, int S, const __grid_constant__ PhiloxKeys keys,
                                                       int64_t gene0, int64_t cell0, uint8_t *__restrict__ side,
                                                       uint8_t *__restrict__ sideb, double *__restrict__ out, int64_t sstride)
{
    extern __shared__ uint32_t nz_tab[]; // [KT][THREADS]
    const uint32_t tsa = (uint32_t)__cvta_generic_to_shared(nz_tab + threadIdx.x);
    const int64_t kfirst = __ldg(colptr + col0);
    const unsigned long long lo = offsets[cls], hi = offsets[cls + 1];
    const bool vec = (S & 3) == 0; // four draws (one Philox block) leave as one store: single uint16 / byte stores were bound by
                                   // the store transaction rate, not by the sampling
    for (unsigned long long t = lo + blockIdx.x * (unsigned long long)THREADS + threadIdx.x; t < hi;
         t += (unsigned long long)gridDim.x * THREADS) {
        const NzItem it = list[t];
        const bool pair = (it.krel & NZ_PAIR) != 0u;
        const int64_t j = col0 + it.col;
        int32_t i;
        double xv = 0.0;
        if (pair) i = __ldg(hgenes + (it.krel & ~NZ_PAIR));
        else {
            i = __ldg(rowidx + kfirst + it.krel);
            xv = __ldg(val + kfirst + it.krel);
        }
        float rf, Mf;
        nb_params(xv, (float)__ldg(size + i), (float)__ldg(mu + i), (float)__ldg(beta + j), rf, Mf);
        NbTable<KT, THREADS * 4> T;
        const bool table = Mf > 0.0f; // else rgamma(shape, scale = 0) = 0 -> rpois(0) = 0: the draw is x itself
        if (table) T.build(tsa, rf, Mf);
        {
            // bytes for the main kernel's strip -- in CSC order for a stored entry, at (row of the heavy gene, column) for a
            // pair -- and values >= 255 straight to the output
            uint8_t *o = pair ? sideb + ((size_t)(it.krel & ~NZ_PAIR) * (size_t)ncol + (size_t)it.col) * (size_t)S
                              : side + (size_t)it.krel * (size_t)S;
            double *po = out + i + G * (int64_t)it.col;
            uint32_t pack = 0;
            auto emit = [&](int s, uint32_t kv) {
                if (kv >= 255u) {
                    __stcs(po + sstride * s, xv + (double)kv);
                    kv = 255u;
                }
                if (vec) {
                    pack |= kv << (8 * (s & 3));
                    if ((s & 3) == 3) {
                        *reinterpret_cast<uint32_t *>(o + (s - 3)) = pack;
                        pack = 0;
                    }
                } else
                    o[s] = (uint8_t)kv;
            };
            if (table) T.draws(keys, (uint32_t)(gene0 + i), (uint32_t)(cell0 + j), 0, S, emit);
            else
                for (int s = 0; s < S; s++) emit(s, 0u);
        }
    }
}

// class 5: onto the gamma-Poisson list; the side buffer says "not here"
__global__ void __launch_bounds__(256) k_nz_defer(const NzItem *__restrict__ list, const unsigned long long *__restrict__ offsets,
                                                  int64_t col0, const int64_t *__restrict__ colptr,
                                                  const int32_t *__restrict__ rowidx, const double *__restrict__ val,
                                                  const double *__restrict__ beta, const double *__restrict__ size,
                                                  const double *__restrict__ mu, int S, uint8_t *__restrict__ side,
                                                  HeavyItem *__restrict__ hlist, unsigned int *__restrict__ hcount, unsigned int hcap)
{
    const int64_t kfirst = __ldg(colptr + col0);
    const unsigned long long lo = offsets[NB_CLASSES - 1], hi = offsets[NB_CLASSES];
    for (unsigned long long t = lo + blockIdx.x * 256ull + threadIdx.x; t < hi; t += (unsigned long long)gridDim.x * 256ull) {
        const NzItem it = list[t];
        const int64_t k = kfirst + it.krel, j = col0 + it.col;
        const int32_t i = __ldg(rowidx + k);
        HeavyItem h;
        h.gene = i;
        h.col = it.col;
        h.x = __ldg(val + k);
        nb_params(h.x, (float)__ldg(size + i), (float)__ldg(mu + i), (float)__ldg(beta + j), h.r, h.M);
        uint8_t *o = side + (size_t)it.krel * (size_t)S;
        for (int s = 0; s < S; s++) o[s] = 255;
        const unsigned slot = atomicAdd(hcount, 1u);
        if (slot < hcap) hlist[slot] = h; // past the capacity: hcount > hcap tells the host, which repeats the call
    }
}

// ---------------------------------------------------------------------------------
// Heavy genes: x = 0 elements that need more than SM_KCAP table points at the chunk's smallest BETA (largest mean).
// gclass[g] = 0 light (main kernel), 1..4 table of 32 << class points (k_hz_sample), 5 beyond SM_KMAX at that BETA (the main
// kernel tests every x = 0 element of such a gene and sends those past the limit to the gamma-Poisson list; the test is per
// ELEMENT, so which sampler draws an element never depends on the chunk).  hidx[g] = row of the gene in the byte side
// buffer (-1 light, -2 class 5), hgenes[c] lists the genes of class c (order left to the atomics: it only decides where a
// gene's bytes sit).
// ---------------------------------------------------------------------------------
// out[0] = min over v (as float), out pre-set to a large positive float by the caller (bytes 0x7f).  Positive floats order like
// their bit patterns, so the CTAs combine with an integer atomicMin; a vector with non-positive entries (rejected by the BETA
// range check of every R-level entry) still gives ONE value per vector, which is all the callers need.  One CTA per 8192
// elements: the single-CTA version took 190 us on 1.3 M cells, 8 % of a posterior call of config 5.
__global__ void __launch_bounds__(256) k_min_vec(int64_t n, const double *__restrict__ v, float *__restrict__ out)
{
    __shared__ float sh[8];
    float mn = 3.0e38f;
    const int64_t base = (int64_t)blockIdx.x * 8192;
    const int64_t end = base + 8192 < n ? base + 8192 : n;
    for (int64_t k = base + threadIdx.x; k < end; k += 256) mn = fminf(mn, (float)__ldg(v + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) mn = fminf(mn, sh[k]);
        if (mn > 0.0f) atomicMin(reinterpret_cast<int *>(out), __float_as_int(mn));
        else atomicMin(reinterpret_cast<int *>(out), 0); // a non-positive entry: the callers' formulas only need a fixed value
    }
}

__global__ void __launch_bounds__(256) k_gene_class(int64_t G, const double *__restrict__ size, const double *__restrict__ mu,
                                                    const float *__restrict__ beta_min, uint8_t *__restrict__ gclass,
                                                    int32_t *__restrict__ hidx, int32_t *__restrict__ hgenes /*[G]*/,
                                                    unsigned int *__restrict__ hcnt /*number of heavy genes*/)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    float rf, Mf;
    nb_params(0.0, (float)size[g], (float)mu[g], beta_min[0], rf, Mf);
    int c = nb_class(rf, Mf);
    if (!(rf > 0.0f)) c = 0;
    if (c >= 1 && c <= 4) {
        // the main kernel keeps a gene whose first SM_KCAP points hold SM_LIGHT_MASS of the mass even at its largest mean: the
        // few words beyond walk the recurrence there (cheaper than a 64-point table per column for every such gene)
        const float q = __fdividef(Mf, rf + Mf), c1 = q * (rf - 1.0f);
        float p = __expf(-rf * log1pf(__fdividef(Mf, rf))), ctop = p;
        for (int k = 1; k < 32; k++) {
            p *= fmaf(c1, c_rcp[k], q);
            ctop += p;
        }
        if (ctop >= SM_LIGHT_MASS && Mf < SM_LIGHT_MAX) c = 0;
    }
    gclass[g] = (uint8_t)c;
    int32_t h = c == NB_CLASSES - 1 ? -2 : -1; // -2: some element of this gene may need the gamma-Poisson list
    if (c >= 1 && c <= 4) {
        h = (int32_t)atomicAdd(hcnt, 1u);
        hgenes[h] = (int32_t)g;
    }
    hidx[g] = h;
}

// ---------------------------------------------------------------------------------
// The main kernel: x = 0 elements of light genes, the strip and the flush.
// ---------------------------------------------------------------------------------
#define SM_SMEM_PER_WARP(KCAP) (SM_GENES * 8 + (KCAP) * 32 * 4 + SM_GENES * SM_STRIDE * 4)

template <int WARPS, int MINB, int KCAP, bool COUNTS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    k_post_sample_main(int64_t G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                       const double *__restrict__ val, const double *__restrict__ beta, const double *__restrict__ size,
                       const double *__restrict__ mu, int64_t col0, int64_t ncol, int nc_per_cta, double *__restrict__ out, int S,
                       const __grid_constant__ PhiloxKeys keys, int64_t gene0, int64_t cell0, int64_t sstride,
                       const uint8_t *__restrict__ perm, const uint8_t *__restrict__ side, const uint8_t *__restrict__ gclass,
                       const int32_t *__restrict__ hidx, const uint8_t *__restrict__ sideb, HeavyItem *__restrict__ hlist,
                       unsigned int *__restrict__ hcount, unsigned int hcap, const uint32_t *__restrict__ blockptr, int64_t Ctot)
{
    constexpr int LCAP = (KCAP == 32) ? 5 : 6;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + warp * SM_SMEM_PER_WARP(KCAP);
    double *xt = reinterpret_cast<double *>(wbase);                                   // counts of the warp's genes in the current column (0 = absent)
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(wbase + SM_GENES * 8);             // [k][lane]: one conflict-free column per thread
    uint32_t *s_strip = s_tab + KCAP * 32;                                            // [gene][SM_STRIDE]: byte s of a gene's run = draw of plane s
    const int64_t w0 = ((int64_t)blockIdx.x * WARPS + warp) * SM_GENES; // first gene of this warp
    const int64_t jb = (int64_t)blockIdx.y * nc_per_cta;
    const int nc = (int)((jb + nc_per_cta < ncol ? jb + nc_per_cta : ncol) - jb);
    if (w0 >= G) return;
    const int64_t kfirst = __ldg(colptr + col0); // the side buffer is indexed from the chunk's first stored entry
    // slice bounds of columns lane and lane + 32 of the chunk from the blocked column pointers (the warp's 128 genes are one
    // row block) as 32-bit offsets from the first column of the chunk, and their BETA
    const uint32_t *bp_lo = blockptr + (w0 / BN_ROWBLOCK) * Ctot + col0 + jb;
    const int64_t base = __ldg(colptr + col0 + jb);
    rowidx += base;
    val += base;
    side += (size_t)(base - kfirst) * (size_t)S;
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
    // The draws are made from a float pmf, so the NB parameters are evaluated in float too: a relative 1e-7 in a
    // parameter is far below what any test of the draws can resolve.  Slot e of lane l works on gene w0 + go[e];
    // hz[e] >= 0: a heavy gene, its x = 0 draws wait in the byte side buffer at row hz[e].
    float sz[4], m[4];
    int go[4], hz[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int64_t q = w0 + e * 32 + lane; // a position in the block's sorted order (the table is padded to whole blocks)
        go[e] = (int)__ldg(perm + q);
        const int64_t i = w0 + go[e];
        sz[e] = (i < G) ? (float)__ldg(size + i) : 1.0f;
        m[e] = (i < G) ? (float)__ldg(mu + i) : 0.0f;
        hz[e] = (i < G) ? __ldg(hidx + i) : -1;
        xt[e * 32 + lane] = 0.0;
    }
    // strip bytes of rows past the end of the matrix stay 255 ("not written here") for good: the flush needs no bounds test
    for (int q = lane; q < SM_GENES * SM_STRIDE; q += 32) s_strip[q] = 0xffffffffu;
    uint32_t tsa = (uint32_t)__cvta_generic_to_shared(s_tab + lane);   // this thread's table column: entry k at tsa + k * 128
    uint32_t ssa = (uint32_t)__cvta_generic_to_shared(s_strip);        // strip word w of gene g at ssa + (g * SM_STRIDE + w) * 4
    // opaque to the compiler from here on: it would otherwise re-derive both from %tid / %ctaid inside the draw loop (15
    // instructions per four draws) instead of holding two registers
    asm volatile("" : "+r"(tsa), "+r"(ssa));
    unsigned hblock = 0, hused = SM_HBLOCK; // deferred-list block of this warp (warp-uniform)
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
        const int64_t jl = jb + jj, j = col0 + jl;
        const float bf = __shfl_sync(0xffffffffu, (jj < 32) ? beta_a : beta_b, jj & 31);
        for (int sb = 0; sb < S; sb += SM_PLANES) { // one pass unless S > SM_PLANES
            const int nS = min(S - sb, SM_PLANES);
            const int nW = (nS + 3) >> 2;
            // ---- stored entries of the column: counts into xt, their draws (k_nz_sample) into the strip ----------------
            {
                uint32_t k = kcur;
                int32_t r2 = pr;
                double v2 = pv;
                for (;;) {
                    const bool mine = k + lane < kend;
                    if (mine) {
                        const int g = r2 - (int)w0;
                        // an explicitly stored zero is a stored entry like any other (its draws wait in the side buffer): -0.0
                        // keeps it apart from "absent" below and adds like 0 in the flush
                        if (sb == 0) xt[g] = (v2 == 0.0) ? -0.0 : v2;
                        // the entry's draws (k_nz_sample): one byte per plane, 255 = stored elsewhere (value >= 255: written by
                        // k_nz_sample itself; gamma-Poisson list: written after this kernel)
                        const uint8_t *sp = side + (size_t)(k + lane) * (size_t)S + sb;
                        const uint32_t sw = ssa + g * (SM_STRIDE * 4);
                        if ((S & 3) == 0) {
                            const uint32_t *spw = reinterpret_cast<const uint32_t *>(sp);
                            for (int w = 0; w < nW; w++) sts_u32(sw + w * 4, __ldg(spw + w));
                        } else {
                            for (int w = 0; w < nW; w++) {
                                uint32_t word = 0;
                                for (int b = 0; b < 4; b++) word |= (uint32_t)((w * 4 + b < nS) ? __ldg(sp + w * 4 + b) : 255u) << (8 * b);
                                sts_u32(sw + w * 4, word);
                            }
                        }
                    }
                    k += 32;
                    if (k >= kend) break;
                    r2 = 0x7fffffff;
                    v2 = 0.0;
                    if (k + lane < kend) { // a longer slice (dense matrices): the rest straight from memory
                        r2 = __ldg(rowidx + k + lane);
                        v2 = __ldcs(val + k + lane);
                    }
                }
            }
            __syncwarp();
            // prefetch the slice of column jj + 1 (after the last pass of this column)
            if (sb + SM_PLANES >= S) {
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
            }
            // ---- x = 0 elements, slot by slot ------------------------------------------------------------------------
#pragma unroll 1
            for (int e = 0; e < 4; e++) {
                // size/mu/offset of the slot's gene sit in element 0 of the arrays, which are rotated at the end of the step so
                // that they stay in registers (a dynamic index would put them in local memory)
                const int g = go[0], hz_ = hz[0];
                const float s_ = sz[0], m_ = m[0];
                sz[0] = sz[1], sz[1] = sz[2], sz[2] = sz[3], sz[3] = s_;
                m[0] = m[1], m[1] = m[2], m[2] = m[3], m[3] = m_;
                go[0] = go[1], go[1] = go[2], go[2] = go[3], go[3] = g;
                hz[0] = hz[1], hz[1] = hz[2], hz[2] = hz[3], hz[3] = hz_;
                const int64_t i = w0 + g;
                const bool zero = (i < G) && __double_as_longlong(xt[g]) == 0ll; // absent from the column (+0.0 exactly)
                const float mbf = m_ * bf;
                const float rf = s_, Mf = __fdividef(rf * (m_ - mbf), s_ + mbf);
                const bool pos = Mf > 0.0f;
                const uint32_t sw = ssa + g * (SM_STRIDE * 4);
                // -log p(0); elements beyond SM_KMAX table points, or whose p(0) leaves the float range, go to the gamma-Poisson
                // list: nb_class() == 5 for THIS element (only genes that were class 5 at the chunk's smallest BETA can have any)
                const float lp0 = rf * log1pf(__fdividef(Mf, rf));
                const bool defer = zero && pos && hz_ == -2 && nb_class(rf, Mf) == NB_CLASSES - 1;
                const bool heavy = zero && pos && !defer && hz_ >= 0;
                const bool work = zero && pos && !defer && !heavy;
                if (zero && !work && !heavy) { // mean 0: the draw is x itself (k = 0); deferred: 255 = not written here
                    const uint32_t fill = defer ? 0xffffffffu : 0u;
                    for (int w = 0; w < nW; w++) sts_u32(sw + w * 4, fill);
                }
                if (heavy) { // the draws k_hz_sample made for this (gene, column)
                    const uint8_t *hb = sideb + ((size_t)hz_ * (size_t)ncol + (size_t)jl) * (size_t)S + sb;
                    if ((S & 3) == 0) {
                        const uint32_t *hw = reinterpret_cast<const uint32_t *>(hb);
                        for (int w = 0; w < nW; w++) sts_u32(sw + w * 4, __ldg(hw + w));
                    } else {
                        for (int w = 0; w < nW; w++) {
                            uint32_t word = 0;
                            for (int b = 0; b < 4; b++) word |= (uint32_t)((w * 4 + b < nS) ? __ldg(hb + w * 4 + b) : 255u) << (8 * b);
                            sts_u32(sw + w * 4, word);
                        }
                    }
                }
                const unsigned hm = (sb == 0) ? __ballot_sync(0xffffffffu, defer) : 0u;
                if (hm) {
                    // append to the warp's block of the global list; a new block (one global atomic per SM_HBLOCK slots) is
                    // reserved when the current one cannot take them, the slots left over are marked empty
                    const unsigned n = (unsigned)__popc(hm);
                    if (hused + n > SM_HBLOCK) {
                        if (hblock + SM_HBLOCK <= hcap) {
                            if (lane < (int)(SM_HBLOCK - hused)) hlist[hblock + hused + lane].gene = -1;
                            if (lane + 32 < (int)(SM_HBLOCK - hused)) hlist[hblock + hused + lane + 32].gene = -1;
                        }
                        if (lane == 0) hblock = atomicAdd(hcount, (unsigned)SM_HBLOCK);
                        hblock = __shfl_sync(0xffffffffu, hblock, 0);
                        hused = 0;
                    }
                    // a block past the capacity is dropped: hcount > hcap tells the host, which repeats the call with
                    // worst-case lists (draws are pure functions of their coordinates, so the repeat is idempotent)
                    if (defer && hblock + SM_HBLOCK <= hcap) {
                        HeavyItem h;
                        h.gene = (int32_t)i;
                        h.col = (int32_t)jl;
                        h.r = rf;
                        h.M = Mf;
                        h.x = 0.0;
                        hlist[hblock + hused + __popc(hm & ((1u << lane) - 1u))] = h;
                    }
                    hused += n;
                }
                if (!__any_sync(0xffffffffu, work)) continue;
                // ---- the CDF table, 8 / 16 / 32 / KCAP points deep: until every lane holds SM_TOPF of its mass ----------
                const float q = __fdividef(Mf, rf + Mf), c1 = q * (rf - 1.0f);
                const float p0 = __expf(-lp0) * 4294967296.0f;
                float p = work ? p0 : 0.0f;
                float ctop = work ? p : 4294967296.0f;
                sts_u32(tsa, __float2uint_rz(ctop)); // the conversion saturates at 2^32 - 1
#pragma unroll
                for (int k = 1; k < 8; k++) {
                    p *= fmaf(c1, 1.0f / (float)k, q);
                    ctop += p;
                    sts_u32(tsa + k * 128, __float2uint_rz(ctop));
                }
                int levels = 3;
                // complete: the mass is there, or the pmf no longer moves the float sum (a complete table whose float total
                // fell short of 2^32 by rounding)
#define SM_COMPLETE() (!work || ctop >= SM_TOPF || !(p > ctop * 3e-8f))
                if (!__all_sync(0xffffffffu, SM_COMPLETE())) {
#pragma unroll
                    for (int k = 8; k < 16; k++) {
                        p *= fmaf(c1, 1.0f / (float)k, q);
                        ctop += p;
                        sts_u32(tsa + k * 128, __float2uint_rz(ctop));
                    }
                    levels = 4;
                    if (!__all_sync(0xffffffffu, SM_COMPLETE())) {
#pragma unroll
                        for (int k = 16; k < 32; k++) {
                            p *= fmaf(c1, 1.0f / (float)k, q);
                            ctop += p;
                            sts_u32(tsa + k * 128, __float2uint_rz(ctop));
                        }
                        levels = 5;
                        if (KCAP > 32 && !__all_sync(0xffffffffu, SM_COMPLETE())) {
#pragma unroll
                            for (int k = 32; k < KCAP; k++) {
                                p *= fmaf(c1, 1.0f / (float)k, q);
                                ctop += p;
                                sts_u32(tsa + k * 128, __float2uint_rz(ctop));
                            }
                            levels = LCAP;
                        }
                    }
                }
                const int K = 1 << levels;
                const bool walkable = p > ctop * 3e-8f; // the pmf still moves the sum at the last tabulated point
                const uint32_t tlast = lds_u32(tsa + (K - 1) * 128);
                const uint32_t c0 = (uint32_t)(gene0 + i), c1p = (uint32_t)(cell0 + j);
                double *po = out + G * jl + i + sstride * sb;
                // one search level: thresholds are non-decreasing, k = #{thresholds <= u}
#define SLEVEL(A, U, PROBE, STEP)                                                                                   \
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tld.shared.u32 t, [%0+%2];\n\tsetp.ge.u32 p, %1, t;\n\t"            \
                 "@p add.u32 %0, %0, %3;\n\t}"                                                                         \
                 : "+r"(A)                                                                                             \
                 : "r"(U), "n"((PROBE) * 128), "n"((STEP) * 128))
#define SKVAL(A) ((A - tsa) >> 7)
                // a draw beyond the table: the recurrence goes on from the last point (tables of genes near the light limit hold
                // 85-99.97 % of the mass), or restarts from 0 when the pmf no longer moves the sum there (rounding deficit of
                // a complete table; restarting makes the result independent of the depth this slot happened to build);
                // values >= 255 are stored directly
#define SM_SLOW(KV, U, PLANE)                                                                                    \
    do {                                                                                                         \
        if ((U) >= tlast) KV = walkable ? sample_walk((U), p, ctop, c1, q, (float)(K - 1)) : sample_from0((U), p0, c1, q); \
        if (KV >= 255u) {                                                                                        \
            __stcs(po + sstride * (PLANE), (double)KV);                                                          \
            KV = 255u;                                                                                           \
        }                                                                                                        \
    } while (0)
                // ONE copy of the draw loop; the depth of the search is a warp-uniform run-time value (three uniform branches per
                // Philox block).  Four specialised copies were 1.3x the instruction-cache footprint (6.8 % of the warp stalls were
                // "no instruction") and each re-loaded the round keys into its own uniform registers.
                // (a partial last Philox block -- S not a multiple of 4 -- fills bytes of planes >= nS, which the flush ignores)
                if (work) {
#pragma unroll 1
                    for (int s0 = 0; s0 < nS; s0 += 4) {
                        uint32_t u0, u1, u2, u3;
                        bn_philox_block(keys, c0, c1p, (uint32_t)((sb + s0) >> 2), 0u, u0, u1, u2, u3);
                        uint32_t a0 = tsa, a1 = tsa, a2 = tsa, a3 = tsa;
                        if (LCAP >= 6 && levels >= 6) { SLEVEL(a0, u0, 31, 32); SLEVEL(a1, u1, 31, 32); SLEVEL(a2, u2, 31, 32); SLEVEL(a3, u3, 31, 32); }
                        if (levels >= 5) { SLEVEL(a0, u0, 15, 16); SLEVEL(a1, u1, 15, 16); SLEVEL(a2, u2, 15, 16); SLEVEL(a3, u3, 15, 16); }
                        if (levels >= 4) { SLEVEL(a0, u0, 7, 8); SLEVEL(a1, u1, 7, 8); SLEVEL(a2, u2, 7, 8); SLEVEL(a3, u3, 7, 8); }
                        SLEVEL(a0, u0, 3, 4); SLEVEL(a1, u1, 3, 4); SLEVEL(a2, u2, 3, 4); SLEVEL(a3, u3, 3, 4);
                        SLEVEL(a0, u0, 1, 2); SLEVEL(a1, u1, 1, 2); SLEVEL(a2, u2, 1, 2); SLEVEL(a3, u3, 1, 2);
                        SLEVEL(a0, u0, 0, 1); SLEVEL(a1, u1, 0, 1); SLEVEL(a2, u2, 0, 1); SLEVEL(a3, u3, 0, 1);
                        uint32_t k0 = SKVAL(a0), k1 = SKVAL(a1), k2 = SKVAL(a2), k3 = SKVAL(a3);
                        const uint32_t umax = max(max(u0, u1), max(u2, u3));
                        if (umax >= tlast) { /* some draw beyond the table (rare): K <= 64 < 255, so inside draws need no check */
                            SM_SLOW(k0, u0, s0);
                            SM_SLOW(k1, u1, s0 + 1);
                            SM_SLOW(k2, u2, s0 + 2);
                            SM_SLOW(k3, u3, s0 + 3);
                        }
                        sts_u32(sw + s0, k0 | (k1 << 8) | (k2 << 16) | (k3 << 24));
                    }
                }
                __syncwarp();
#undef SM_SLOW
#undef SKVAL
#undef SLEVEL
#undef SM_COMPLETE
                __syncwarp(); // reconverge before the next slot's table is written
            }
            __syncwarp();
            // ---- flush: plane by plane, the warp's 128 genes (natural order) as four back-to-back 256 B stores -------------
            {
                // x + k without an integer -> double conversion: the byte becomes the low word of the double 2^52 + k, then
                // (2^52 + k) + (x - 2^52), exact for counts; real-valued x (posterior on normalised data) takes
                // ((2^52 + k) - 2^52) + x.  The store is predicated on the byte (255 = written elsewhere), no branch.
                double xa[4];
#pragma unroll
                for (int e = 0; e < 4; e++) xa[e] = COUNTS ? xt[e * 32 + lane] - 4503599627370496.0 : xt[e * 32 + lane];
                double *po = out + G * jl + w0 + lane + sstride * sb;
                const uint32_t sl = ssa + lane * (SM_STRIDE * 4);
                for (int w = 0; w < nW; w++) {
                    uint32_t wd[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) wd[e] = lds_u32(sl + e * (32 * SM_STRIDE * 4) + w * 4);
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        if (w * 4 + b < nS) {
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t kb = __byte_perm(wd[e], 0u, 0x4440u + b);
                                const double kd = __hiloint2double(0x43300000, (int)kb);
                                const double v = COUNTS ? kd + xa[e] : (kd - 4503599627370496.0) + xa[e];
                                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 255;\n\t@p st.global.cs.f64 [%0], %1;\n\t}" ::"l"(po + e * 32),
                                             "d"(v), "r"(kb)
                                             : "memory");
                            }
                            po += sstride;
                        }
                    }
                }
            }
            __syncwarp(); // the strip is rewritten by the next pass / column
        }
#pragma unroll
        for (int e = 0; e < 4; e++) xt[e * 32 + lane] = 0.0; // own slots: ready for the next column's scatter
        __syncwarp();
    }
    // close the warp's last deferred-list block
    if (hused < SM_HBLOCK && hblock + SM_HBLOCK <= hcap) {
        if (lane < (int)(SM_HBLOCK - hused)) hlist[hblock + hused + lane].gene = -1;
        if (lane + 32 < (int)(SM_HBLOCK - hused)) hlist[hblock + hused + lane + 32].gene = -1;
    }
}

// deferred entries x S, one item per thread (items of one entry are neighbours: same gamma shape); the gamma-Poisson
// mixture the reference uses (Marsaglia-Tsang + PTRS) in float, sub-stream 2^31 + s of the element
__global__ void __launch_bounds__(256) k_post_sample_heavy(int64_t G, const HeavyItem *__restrict__ hlist,
                                                           const unsigned int *__restrict__ hcount, unsigned int hcap, int S,
                                                           uint64_t seed, int64_t gene0, int64_t cell0, int64_t col0,
                                                           int64_t sstride, double *__restrict__ out,
                                                           unsigned int *__restrict__ hmax)
{
    // last kernel of a chunk: fold the chunk's slot count into the running maximum the host checks
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(hmax, *hcount);
    const int64_t items = (int64_t)min(*hcount, hcap) * S;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += gridDim.x * (int64_t)blockDim.x) {
        const int64_t ent = it / S;
        const int s = (int)(it - ent * S);
        const HeavyItem h = hlist[ent];
        if (h.gene < 0) continue; // a slot its warp left empty
        Philox g;
        g.init(seed, (uint64_t)(gene0 + h.gene) + (uint64_t)(cell0 + col0 + h.col) * 0x100000000ULL, 0x80000000u + (uint32_t)s);
        __stcs(out + (int64_t)h.gene + G * (int64_t)h.col + sstride * s, bn_nb_heavy_draw(h.r, h.M, g) + h.x);
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
struct SampleArgs {
    bn_ctx *ctx;
    const bn_mat *m;
    const double *b, *sz, *mm;
    int S;
    uint64_t seed;
};

template <int KT, int THREADS>
static int launch_nz_class(const SampleArgs &a, cudaStream_t st, int cls, int grid, const NzItem *list, const unsigned long long *offsets,
                           int64_t ncol, int64_t col0, const int32_t *hgenes, uint8_t *side, uint8_t *sideb, double *out, int64_t sstride)
{
    bn_ctx *ctx = a.ctx;
    const size_t smem = (size_t)KT * THREADS * 4;
    BN_CUDA(ctx, cudaFuncSetAttribute(k_nz_sample<KT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BN_CUDA(ctx, cudaFuncSetAttribute(k_nz_sample<KT, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    BN_LAUNCH_ON(ctx, st, (k_nz_sample<KT, THREADS>), grid, THREADS, smem, cls, list, offsets, a.m->G, ncol, col0, a.m->colptr, a.m->rowidx,
                 a.m->val, hgenes, a.b, a.sz, a.mm, a.S, bn_philox_keys(a.seed), a.m->gene0, a.m->cell0, side, sideb, out, sstride);
    return BN_OK;
}

// device pointers in, columns [col0, col0 + ncol) of the matrix; out is G x ncol x S
int bn_post_sample_device(bn_ctx *ctx, const bn_mat *m, const double *d_beta, const double *d_size, const double *d_mu,
                          int64_t col0, int64_t ncol, int S, uint64_t seed, double *d_out)
{
    constexpr int SW_WARPS = 8, KCAP = 32, MINB = 3;
    const int sms = ctx->prop.multiProcessorCount;
    const int64_t G = m->G;
    const int64_t SW_TG = (int64_t)SM_GENES * SW_WARPS;
    const int64_t tiles_s = (G + SW_TG - 1) / SW_TG;
    const SampleArgs A{ctx, m, d_beta, d_size, d_mu, S, seed};
    {
        // one shared-memory carveout for every kernel of the pipeline: CTAs of kernels that ask for different carveouts cannot
        // share an SM, which would serialise the side streams behind the main kernel
        if (!ctx->sample_attr_set) {
            cudaFuncSetAttribute(k_nz_classify<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_nz_classify<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_nz_offsets, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_nz_defer, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_post_sample_heavy, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_post_sample_main<8, 3, 32, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_post_sample_main<8, 3, 32, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            ctx->sample_attr_set = true;
        }
    }
    BN_TRY(bn_mat_ensure_blockptr(ctx, m));
    // per call: the smallest BETA of the WHOLE vector (largest posterior mean of every gene: which kernel draws a gene's x = 0
    // elements must not depend on the column block of the call), the slot order of every 128-gene block (padded to whole
    // blocks) and the heavy genes
    const int64_t nblk = (G + SM_GENES - 1) / SM_GENES;
    DevBuf<uint8_t> perm, gclass;
    DevBuf<int32_t> hidx, hgenes;
    DevBuf<float> bmin;
    DevBuf<unsigned int> hc; // {gamma-Poisson list slots of the chunk, max over the chunks, heavy genes}
    BN_TRY(perm.alloc(ctx, (size_t)nblk * SM_GENES));
    BN_TRY(gclass.alloc(ctx, (size_t)G));
    BN_TRY(hidx.alloc(ctx, (size_t)G));
    BN_TRY(hgenes.alloc(ctx, (size_t)G));
    BN_TRY(bmin.alloc(ctx, 1));
    BN_TRY(hc.alloc(ctx, 4));
    BN_CUDA(ctx, cudaMemsetAsync(hc.p, 0, 4 * sizeof(unsigned int), ctx->stream));
    BN_CUDA(ctx, cudaMemsetAsync(bmin.p, 0x7f, sizeof(float), ctx->stream));
    BN_LAUNCH(ctx, k_min_vec, (unsigned)((m->C + 8191) / 8192), 256, 0, m->C, d_beta, bmin.p);
    BN_LAUNCH(ctx, k_sample_perm, (unsigned)((nblk + 7) / 8), 256, 0, G, d_size, d_mu, bmin.p, perm.p);
    BN_LAUNCH(ctx, k_gene_class, (unsigned)((G + 255) / 256), 256, 0, G, d_size, d_mu, bmin.p, gclass.p, hidx.p, hgenes.p, hc.p + 2);
    unsigned int *h_nh = reinterpret_cast<unsigned int *>(ctx->pinned + 34);
    BN_CUDA(ctx, cudaMemcpyAsync(h_nh, hc.p + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    BN_TRY(bn_sync(ctx));
    const int64_t nheavy = (int64_t)*h_nh;
    // column chunks: the side buffers (2 S bytes per stored entry, S bytes per (heavy gene, column), 8 bytes of list per element of
    // either kind) stay below ~1 GB
    const double rho = (double)m->nnz / ((double)G * (double)m->C);
    const double per_col = (std::max(1.0, rho * (double)G) + (double)nheavy) * (S + 8.0);
    for (int attempt = 0; attempt < 2; attempt++) {
        int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncol, (int64_t)(((double)(1 << 30)) / per_col)));
        // BN_TUNE_SAMPLE_CHUNKS: split the call into at least that many chunks (A/B switch).  Measured on config 5 (calls of 6060
        // columns): 1 chunk 2748 ms, 2: 2734, 4: 2777, 8: 2823 -- preparing chunk c + 1 next to the main kernel of chunk c buys
        // nothing once the main kernel fills the SMs; the phase is the sum of its kernels.
        if (ctx->tune[BN_TUNE_SAMPLE_CHUNKS] > 0) {
            const int want = ctx->tune[BN_TUNE_SAMPLE_CHUNKS];
            const int64_t per = ((ncol + want - 1) / want + SM_NCMAX - 1) / SM_NCMAX * SM_NCMAX;
            chunk = std::min(chunk, std::max<int64_t>(per, SM_NCMAX));
        }
        // capacity of the gamma-Poisson list (whole blocks of SM_HBLOCK): first attempt 1/8 entry per expected non-zero
        // plus the blocks warps leave partly empty; a chunk that overflows makes the whole call repeat with worst-case lists
        const double per_elem = attempt == 0 ? std::min(1.0, std::max(1.0 / 64, 0.125 * rho)) : 1.0;
        const int64_t ctas64 = tiles_s * ((chunk + SM_NCMAX - 1) / SM_NCMAX), ctas4 = tiles_s * ((chunk + 3) / 4);
        const size_t max_ctas = (size_t)std::max<int64_t>(ctas64, std::min<int64_t>(ctas4, (int64_t)sms * 128 + tiles_s));
        const size_t cap = (((size_t)(per_elem * (double)G * (double)chunk) + SM_HBLOCK - 1) / SM_HBLOCK + (size_t)SW_WARPS * max_ctas + 1) * SM_HBLOCK;
        if (cap > 0xfffffff0u) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: element list too long");
        // Two sets of per-chunk buffers: chunk c + 1 is prepared (classification, side-buffer draws of the stored entries and of the
        // heavy genes' x = 0 elements) on the side streams while the main kernel of chunk c runs on the context's stream.  The five
        // table classes are independent launches; the deep ones (128-512 points per thread) hold few warps per SM, so they run
        // next to each other on their own streams instead of one after the other.
        struct Set {
            HeavyItem *hl;
            uint8_t *side, *sideb;
            NzItem *list;
            unsigned int *ccnt;
            unsigned long long *offs;
            unsigned int *hcnt; // gamma-Poisson list slots of the chunk
        } set[2];
        DevBuf<unsigned long long> offs; // start of every class in the element list [7], two sets
        DevBuf<unsigned int> ccnt;      // per (class, CTA) counts -> positions, two sets
        DevBuf<unsigned int> hcs;       // heavy-list counters of the two sets
        BN_TRY(offs.alloc(ctx, 16));
        BN_TRY(hcs.alloc(ctx, 2));
        BN_CUDA(ctx, cudaMemsetAsync(hc.p, 0, 2 * sizeof(unsigned int), ctx->stream));
        // stored entries per chunk: read the column pointers of the chunk boundaries
        const int64_t nchunk = (ncol + chunk - 1) / chunk;
        std::vector<int64_t> bounds((size_t)nchunk + 1);
        for (int64_t c = 0; c <= nchunk; c++) {
            const int64_t jc = col0 + std::min<int64_t>(c * chunk, ncol);
            BN_CUDA(ctx, cudaMemcpyAsync(&bounds[(size_t)c], m->colptr + jc, 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        BN_TRY(bn_sync(ctx));
        int64_t max_nnz = 1;
        for (int64_t c = 0; c < nchunk; c++) max_nnz = std::max(max_nnz, bounds[(size_t)c + 1] - bounds[(size_t)c]);
        const int64_t max_all = max_nnz + nheavy * chunk;
        if (max_all >= 2147483647LL) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: more than 2^31 list elements in a column chunk");
        const size_t ccnt_n = (size_t)NB_CLASSES * (size_t)((max_all + NZC_PER_CTA - 1) / NZC_PER_CTA + 1);
        BN_TRY(ccnt.alloc(ctx, 2 * ccnt_n));
        const int nset = nchunk > 1 ? 2 : 1;
        for (int k = 0; k < nset; k++) {
            static const int slots[2][4] = {{0, 1, 8, 9}, {10, 11, 12, 13}};
            BN_TRY(bn_ctx_scratch(ctx, slots[k][0], cap * sizeof(HeavyItem), (void **)&set[k].hl));
            BN_TRY(bn_ctx_scratch(ctx, slots[k][1], (size_t)max_nnz * (size_t)S, (void **)&set[k].side));
            BN_TRY(bn_ctx_scratch(ctx, slots[k][2], (size_t)std::max<int64_t>(1, nheavy * chunk) * (size_t)S, (void **)&set[k].sideb));
            BN_TRY(bn_ctx_scratch(ctx, slots[k][3], (size_t)max_all * sizeof(NzItem), (void **)&set[k].list));
            set[k].ccnt = ccnt.p + (size_t)k * ccnt_n;
            set[k].offs = offs.p + 8 * k;
            set[k].hcnt = hcs.p + k;
        }
        if (nset == 1) set[1] = set[0];
        const int64_t sstride = G * ncol; // planes stay G*ncol apart; a chunk is addressed relative to its first column
        const bool serial = ctx->tune[BN_TUNE_SAMPLE_SERIAL] == 1;
        cudaStream_t sp = serial ? ctx->stream : ctx->aux[0];
        cudaStream_t scl[5] = {sp, serial ? sp : ctx->aux[1], serial ? sp : ctx->aux[2], serial ? sp : ctx->aux[3], serial ? sp : ctx->aux[4]};
        cudaEvent_t ev_ready = ctx->ev_aux[0];                         // everything queued on the main stream so far (inputs, buffers)
        cudaEvent_t *ev_prep = ctx->ev_aux + 1, *ev_free = ctx->ev_aux + 3; // [set]: side buffers filled / released by the main kernel
        cudaEvent_t ev_fork = ctx->ev_aux[5], *ev_cls = ctx->ev_aux + 6;   // class streams
        BN_CUDA(ctx, cudaEventRecord(ev_ready, ctx->stream));
        BN_CUDA(ctx, cudaStreamWaitEvent(sp, ev_ready, 0));
        // prepare chunk c on the side streams
        auto prepare = [&](int64_t c) -> int {
            const Set &B = set[c & 1];
            const int64_t c0 = c * chunk, nc_ = std::min<int64_t>(chunk, ncol - c0), jc = col0 + c0;
            double *o = d_out + G * c0;
            if (c >= 2) BN_CUDA(ctx, cudaStreamWaitEvent(sp, ev_free[c & 1], 0)); // the main kernel of chunk c - 2 has read this set
            BN_CUDA(ctx, cudaMemsetAsync(B.hcnt, 0, sizeof(unsigned int), sp));
            const int64_t nnz_c = bounds[(size_t)c + 1] - bounds[(size_t)c];
            const int64_t nall = nnz_c + nheavy * nc_;
            const int64_t ncta = std::max<int64_t>(1, (nall + NZC_PER_CTA - 1) / NZC_PER_CTA);
#define NZC_ARGS nc_, jc, ncta, nheavy, hgenes.p, m->colptr, m->rowidx, m->val, d_beta, d_size, d_mu, m->blockptr, m->C, B.ccnt, B.list
            BN_LAUNCH_ON(ctx, sp, k_nz_classify<false>, (unsigned)ncta, 256, 0, NZC_ARGS);
            BN_LAUNCH_ON(ctx, sp, k_nz_offsets, 1, 1024, 0, (int64_t)NB_CLASSES * ncta, ncta, B.ccnt, B.offs);
            BN_LAUNCH_ON(ctx, sp, k_nz_classify<true>, (unsigned)ncta, 256, 0, NZC_ARGS);
#undef NZC_ARGS
            BN_CUDA(ctx, cudaEventRecord(ev_fork, sp));
            for (int k = 1; k <= 4; k++) BN_CUDA(ctx, cudaStreamWaitEvent(scl[k], ev_fork, 0));
            BN_TRY((launch_nz_class<32, 256>(A, sp, 0, sms * 4, B.list, B.offs, nc_, jc, hgenes.p, B.side, B.sideb, o, sstride)));
            BN_TRY((launch_nz_class<64, 256>(A, scl[1], 1, sms * 3, B.list, B.offs, nc_, jc, hgenes.p, B.side, B.sideb, o, sstride)));
            BN_TRY((launch_nz_class<128, 128>(A, scl[2], 2, sms * 3, B.list, B.offs, nc_, jc, hgenes.p, B.side, B.sideb, o, sstride)));
            BN_TRY((launch_nz_class<256, 128>(A, scl[3], 3, sms, B.list, B.offs, nc_, jc, hgenes.p, B.side, B.sideb, o, sstride)));
            BN_TRY((launch_nz_class<512, 64>(A, scl[4], 4, sms, B.list, B.offs, nc_, jc, hgenes.p, B.side, B.sideb, o, sstride)));
            BN_LAUNCH_ON(ctx, sp, k_nz_defer, sms, 256, 0, B.list, B.offs, jc, m->colptr, m->rowidx, m->val, d_beta, d_size, d_mu, S, B.side, B.hl,
                         B.hcnt, (unsigned)cap);
            for (int k = 1; k <= 4; k++) {
                BN_CUDA(ctx, cudaEventRecord(ev_cls[k], scl[k]));
                BN_CUDA(ctx, cudaStreamWaitEvent(sp, ev_cls[k], 0));
            }
            BN_CUDA(ctx, cudaEventRecord(ev_prep[c & 1], sp));
            return BN_OK;
        };
        BN_TRY(prepare(0));
        for (int64_t c = 0; c < nchunk; c++) {
            const Set &B = set[c & 1];
            const int64_t c0 = c * chunk, nc_ = std::min<int64_t>(chunk, ncol - c0), jc = col0 + c0;
            double *o = d_out + G * c0;
            if (c + 1 < nchunk) BN_TRY(prepare(c + 1));
            BN_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_prep[c & 1], 0));
            // ---- x = 0 elements of light genes, strip, flush --------------------------------------------------------------
            int ncs = SM_NCMAX; // columns per CTA: short walks keep the columns in flight (and their S planes) close together
            while (ncs > 4 && tiles_s * ((nc_ + ncs - 1) / ncs) < (int64_t)sms * 64) ncs >>= 1;
            if (ctx->tune[BN_TUNE_SAMPLE_NCS] > 0) ncs = std::max(1, std::min(SM_NCMAX, ctx->tune[BN_TUNE_SAMPLE_NCS]));
            if ((nc_ + ncs - 1) / ncs > 65535) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "posterior: at most %d columns per call", 65535 * 4);
            dim3 grid_s((unsigned)tiles_s, (unsigned)((nc_ + ncs - 1) / ncs));
            const size_t smem_s = (size_t)SW_WARPS * SM_SMEM_PER_WARP(KCAP);
#define SM_MAIN_ARGS                                                                                                               \
    G, m->colptr, m->rowidx, m->val, d_beta, d_size, d_mu, jc, nc_, ncs, o, S, bn_philox_keys(seed), m->gene0, m->cell0, sstride, perm.p, \
        B.side, gclass.p, hidx.p, B.sideb, B.hl, B.hcnt, (unsigned)cap, m->blockptr, m->C
            if (m->is_count) {
                BN_CUDA(ctx, cudaFuncSetAttribute(k_post_sample_main<SW_WARPS, MINB, KCAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
                BN_LAUNCH(ctx, (k_post_sample_main<SW_WARPS, MINB, KCAP, true>), grid_s, SW_WARPS * 32, smem_s, SM_MAIN_ARGS);
            } else {
                BN_CUDA(ctx, cudaFuncSetAttribute(k_post_sample_main<SW_WARPS, MINB, KCAP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
                BN_LAUNCH(ctx, (k_post_sample_main<SW_WARPS, MINB, KCAP, false>), grid_s, SW_WARPS * 32, smem_s, SM_MAIN_ARGS);
            }
#undef SM_MAIN_ARGS
            // ---- the gamma-Poisson list ---------------------------------------------------------------------------
            BN_LAUNCH(ctx, k_post_sample_heavy, sms * 8, 256, 0, G, B.hl, B.hcnt, (unsigned)cap, S, seed, m->gene0, m->cell0, jc, sstride, o,
                      hc.p + 1);
            BN_CUDA(ctx, cudaEventRecord(ev_free[c & 1], ctx->stream));
        }
        unsigned int *h_max = reinterpret_cast<unsigned int *>(ctx->pinned + 32);
        BN_CUDA(ctx, cudaMemcpyAsync(h_max, hc.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        BN_TRY(bn_sync(ctx));
        if (*h_max <= cap) break;
        if (attempt == 1) return bn_fail(ctx, BN_ERR_CUDA, "posterior: element list overflow with worst-case capacity");
    }
    return BN_OK;
}
