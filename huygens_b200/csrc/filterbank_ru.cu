// filterbank_ru.cu — K1-ru: exact Filterbank kernel for the throughput regime (more lanes than the
// warp-specialised kernel can keep resident: BASELINE configs[2] on one GPU, configs[3]).
//
// Same bit-exact arithmetic as filterbank_ws.inc (reference: src/filterbank.h:170-187), different
// division of labour: with many warps per scheduler the 12-cycle dependency chain of one lane hides
// behind the other warps, so the goal is the lowest instruction count per filter-sample, not the
// shortest critical path.
//     RU warps  lane = filter: feed-forward taps, pre-gain and the recurrence (x by LDS.128 broadcast) into a 2-deep
//               ring of 32-sample tiles; own y1, y2 and p
//     S warps   lane = sample: add the CTA's filters, in index order and with the reference's rounding
//               (src/filterbank.h:130,137), onto the running bank-sum row that comes from the CTA of the filters before
//               and goes to the next CTA (chain.cuh); the chain's tail writes out[t]
//     G warp    feeds the input ring by TMA bulk copies; gateway of the links that a single warp reads (a device's
//               inbox, short chains, every link of the y-tile kernels): keeps the inbound packets of the next tiles in
//               flight and drops the running row into shared memory
// Tile formats.
//   * Term tiles (strict arithmetic, no per-band rows, no ring followers): the RU warps also own the mix gains and leave
//     the mixdown TERMS v = fl(y*g) in the tile, sample-major [32 samples][32 filters].  S warp j adds ALL the CTA's
//     filters for the tiles k = j (mod NS), walking the RU warps' tiles of that index (terms fetched a batch ahead of the
//     additions, slots given back as soon as their terms are in registers): the CTA is one stage of the chain, and on
//     long chains the S warps poll the links themselves, one credit word per reading warp.  Hand-over barriers of an RU
//     warp are per S warp (an mbarrier's phase bit only tells one phase ahead).
//   * y tiles (fused multiply-add, per-band rows, ring followers reading y back): the tile holds y, filter-major;
//     ceil(W/2) S warps own the gains (replayed per sample while they still move), stream the per-band rows and hand
//     the running row to each other through shared memory; S warp 0 gets it from the gateway.
// One CTA = W RU warps + NS S warps + G, of one stream; no __syncthreads in the steady state, where an RU warp runs whole
// input chunks with nothing but the two hand-over barriers per tile around the unrolled recurrence.  9.75 (taps
// {g,0,-g}) to 12.25 (general taps) RU instructions per filter-sample with term tiles, 13.4 issued in total on
// BASELINE configs[2] (ncu) against 11 algorithmic issue slots.
// ENV = 1 / 2 fuses one RMSbank envelope follower per band behind the filters: ring of y^2 in HBM, or a
// lagged replica of every filter that regenerates y[t-width] bit for bit (see FbRuParams).
#include "hg_common.cuh"
#include "filterbank_handle.cuh"
#include "chain.cuh"

namespace hg {

namespace ru {

constexpr int TILE = 32;      // samples per hand-over tile (S lane = sample)
constexpr int NB = 2;         // tile ring depth per RU warp (y tiles)
#ifndef HG_RU_NB_VF
#define HG_RU_NB_VF 2
#endif
constexpr int NB_VF = HG_RU_NB_VF;   // ... of a term-tile kernel (no gain tiles: the shared memory goes to a deeper ring)
constexpr int XT = 1024;      // input chunk
constexpr int XB = 2;         // input ring depth
constexpr int W_MAX = 14;     // RU warps per CTA
constexpr int W_MAX_ENV = 8;  // ... with fused envelope followers: fewer warps, so the 32 prefetched ring values stay in registers
constexpr int GPS = 2;        // RU warps served by one S warp: an S warp must not need more than its fair share
                              // of its scheduler's issue slots, or the RU warps starve behind it
constexpr int NS_MAX = (W_MAX + GPS - 1) / GPS;
constexpr int YB_MAX = 14;    // hand-over barrier pairs per RU warp (>= NB and >= the S warps of a term-tile kernel)
#ifndef HG_RU_PD
#define HG_RU_PD 4
#endif
#ifndef HG_RU_GD
#define HG_RU_GD 2
#endif
constexpr int GD = HG_RU_GD;  // inbound packets (tiles) the gateway keeps in flight
constexpr int PD = HG_RU_PD;         // depth of the running-row hand-over between consecutive S warps
// S warps of a CTA.  y tiles: an S warp serves GPS RU warps and hands the running row to the next S warp.  Term tiles
// (VF): S warp j adds ALL the CTA's filters for the tiles k = j (mod NS); its additions are one dependent chain whose
// every link waits ~10-20 cycles for an issue slot among the RU warps, so the S warps' number is what bounds the
// mixdown's throughput: as many as the register file allows (fb_ru_config), at most one per RU warp.
__host__ __device__ constexpr int s_warps(int W) { return (W + GPS - 1) / GPS; }
#ifndef HG_RU_NS_CAP
#define HG_RU_NS_CAP 8
#endif
constexpr int CTA_WARPS_MAX = 24;   // 24 warps x 32 lanes x 80 registers fit the 65 536 of an SM (ptxas grants 25 warps only 72)
__host__ __device__ constexpr int s_warps_vf(int W, int per_sm) {
    const int room = CTA_WARPS_MAX / per_sm - W - 1;
    return room < HG_RU_NS_CAP ? room : HG_RU_NS_CAP;
}
// which kernels keep the mixdown TERMS fl(y*g) in their tiles (the ring followers read y back; fma needs y and g)
template <bool FMA, bool BANDS, int ENV> struct TermTiles { static constexpr bool value = !FMA && !BANDS && ENV != 1; };

template <typename T> struct RowStride;      // RS*sizeof(T)/16 odd: conflict-free 128-bit rows
template <> struct RowStride<float> { static constexpr int RS = TILE + 4; };
template <> struct RowStride<double> { static constexpr int RS = TILE + 2; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_poll(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// makes the instruction stream wait, at this point, until the load that produces x has returned
__device__ __forceinline__ void reg_fence(float x) { asm volatile("" ::"f"(x) : "memory"); }
__device__ __forceinline__ void reg_fence(double x) { asm volatile("" ::"d"(x) : "memory"); }
template <typename T> __device__ __forceinline__ void ld4(const T* p, T (&x)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&x)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double* p, double (&x)[4]) {
    double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const T (&x)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&x)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}
template <> __device__ __forceinline__ void st4<double>(double* p, const double (&x)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(x[2], x[3]);
}

// dynamic shared memory layout (W known at launch):
//   xs[XB][XT] | y[W][NB][32*RS] | gt[NS][32*RS] | gs[NS][GPS][32] | psum[NS][PD][32] | barriers
template <typename T> struct Layout {
    static constexpr int RS = RowStride<T>::RS;
    int W, NS, LAG, NBk, GT;       // LAG: the replica followers add a second input ring; NBk: tile slots per RU warp; GT: gain tiles
    __host__ __device__ Layout(int w, int ns, bool lag, bool vf) : W(w), NS(ns), LAG(lag ? 1 : 0), NBk(vf ? NB_VF : NB), GT(vf ? 0 : 1) {}
    __host__ __device__ size_t xs_off() const { return 0; }
    __host__ __device__ size_t y_off() const { return sizeof(T) * XB * XT; }
    __host__ __device__ size_t gt_off() const { return y_off() + sizeof(T) * (size_t)W * NBk * 32 * RS; }
    __host__ __device__ size_t gs_off() const { return gt_off() + sizeof(T) * GT * NS * 32 * RS; }
    __host__ __device__ size_t ps_off() const { return gs_off() + sizeof(T) * NS * GPS * 32; }
    __host__ __device__ size_t xl_off() const { return (ps_off() + sizeof(T) * PD * NS * 32 + 15) & ~(size_t)15; }
    __host__ __device__ size_t yl_off() const { return xl_off() + sizeof(T) * (size_t)LAG * XB * XT; }
    __host__ __device__ size_t bar_off() const { return (yl_off() + 15) & ~(size_t)15; }
    __host__ __device__ size_t bytes() const { return bar_off() + 8 * (size_t)(2 * XB + 2 * W * YB_MAX + 2 * PD * NS); }
};

}  // namespace ru

// one sample of one lane, bit-exact (src/filterbank.h:178-179).  TAPS: 0 general; 1: b1 == 0; 2: b1 == 0 and
// b2 == -b0 (the skipped operations are exact identities, see filterbank_ws.inc).
template <typename T, bool FMA, int TAPS>
__device__ __forceinline__ T ru_sample(T x0, T b0, T b1, T b2, T a1, T a2, T p, T& y1, T& y2, T& x1, T& x2, T& m1, T& m2, T& t2) {
    typedef Ar<T, FMA> A;
    T fw;
    if (TAPS == 2 && !FMA) {
        const T m0 = A::mul(b0, x0);
        fw = A::sub(m0, m2);
        m2 = m1; m1 = m0;
    } else {
        fw = A::mul(b0, x0);
        if (TAPS == 0) fw = A::muladd(b1, x1, fw);
        fw = A::muladd(b2, x2, fw);
    }
    T y;
    if (FMA) {
        const T fb = A::muladd(a2, y2, A::mul(a1, y1));
        y = A::mulsub(fw, p, fb);
    } else {
        const T u = A::mul(fw, p);
        const T fb = A::add(A::mul(a1, y1), t2);
        y = A::sub(u, fb);
        t2 = A::mul(a2, y1);
    }
    y2 = y1; y1 = y;
    x2 = x1; x1 = x0;
    return y;
}

// a full tile, unrolled; the next group's inputs are requested before this group is computed
template <typename T, bool FMA, int TAPS>
__device__ __forceinline__ void ru_tile(const T* __restrict__ xc, T* __restrict__ yrow, T b0, T b1, T b2, T a1, T a2, T p,
                                        T& y1, T& y2, T& x1, T& x2, T& m1, T& m2, T& t2) {
    using namespace ru;
    T xq[2][4];
    ld4<T>(xc, xq[0]);
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        if (q + 1 < TILE / 4) ld4<T>(xc + 4 * (q + 1), xq[(q + 1) & 1]);
        T yy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) yy[j] = ru_sample<T, FMA, TAPS>(xq[q & 1][j], b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
        st4<T>(yrow + 4 * q, yy);
    }
}

// VF tiles: the same recurrence, the mixdown term v = fl(y*g) stored sample-major (vcol = tile + lane, row stride VS)
template <typename T, bool FMA, int TAPS>
__device__ __forceinline__ void ru_tile_v(const T* __restrict__ xc, T* __restrict__ vcol, T b0, T b1, T b2, T a1, T a2, T p, T g,
                                          T& y1, T& y2, T& x1, T& x2, T& m1, T& m2, T& t2) {
    using namespace ru;
    typedef Ar<T, FMA> A;
    constexpr int VS = RowStride<T>::RS;
    T xq[2][4];
    ld4<T>(xc, xq[0]);
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        if (q + 1 < TILE / 4) ld4<T>(xc + 4 * (q + 1), xq[(q + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const T y = ru_sample<T, FMA, TAPS>(xq[q & 1][j], b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
            vcol[(4 * q + j) * VS] = A::mul(y, g);               // filterbank.h:130  y_n * g_n
        }
    }
}

// ENV == 2: a full tile of the filter AND of its `width`-lagged replica, interleaved (two independent
// dependency chains per lane), with the envelope follower folded in: q = y^2, q_old = (replica y)^2,
// S <- (q - q_old) + S  (rmsbank.h:58-61; the "+ 0*0" of a real sample's |z|^2 is exact for q >= 0 and dropped).
// CAP: S after tick `je` of the tile (0 <= je < TILE) is captured in s_at_je for the envelope row; without CAP
// the caller takes S at the end of the tile (rows that fall on tile ends, e.g. hop = 64 on aligned blocks).
template <typename T, bool FMA, int TAPS, bool CAP, bool VF>
__device__ __forceinline__ void ru_tile_env(const T* __restrict__ xc, const T* __restrict__ xlc, T* __restrict__ yrow, T g,
                                            T b0, T b1, T b2, T a1, T a2, T p, T b0l, T b1l, T b2l, T a1l, T a2l, T pl,
                                            T& y1, T& y2, T& x1, T& x2, T& m1, T& m2, T& t2,
                                            T& y1l, T& y2l, T& x1l, T& x2l, T& m1l, T& m2l, T& t2l,
                                            T& rsum, int je, T& s_at_je) {
    using namespace ru;
    typedef Ar<T, FMA> A;
    T xq[2][4], xlq[2][4];
    ld4<T>(xc, xq[0]);
    ld4<T>(xlc, xlq[0]);
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        if (q + 1 < TILE / 4) { ld4<T>(xc + 4 * (q + 1), xq[(q + 1) & 1]); ld4<T>(xlc + 4 * (q + 1), xlq[(q + 1) & 1]); }
        T yy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            yy[j] = ru_sample<T, FMA, TAPS>(xq[q & 1][j], b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
            const T yl = ru_sample<T, FMA, TAPS>(xlq[q & 1][j], b0l, b1l, b2l, a1l, a2l, pl, y1l, y2l, x1l, x2l, m1l, m2l, t2l);
            rsum = A::add(A::sub(A::mul(yy[j], yy[j]), A::mul(yl, yl)), rsum);
            if (CAP && 4 * q + j == je) s_at_je = rsum;
            if (VF) yrow[(4 * q + j) * RowStride<T>::RS] = A::mul(yy[j], g);          // sample-major mixdown term (yrow = tile + lane)
        }
        if (!VF) st4<T>(yrow + 4 * q, yy);
    }
}

template <typename T> struct FbRuParams {
    FbK1Params<T> k;
    int W;            // RU warps per CTA
    int NS;           // S warps per CTA
    int cps;          // CTAs per stream = ceil(ceil(N/32) / W)
    // ENV variant: per-band RMSbank envelope followers fused behind the filters (src/rmsbank.h:54-74 on
    // the real band outputs: q = y^2, S <- (q - q[t-width]) + S, env = sqrt(S / width))
    T* env;               // [S][env_rows][N], one row every `hop` samples (row j = sample (j+1)*hop - 1 of the block)
    T* rms_sum;           // [S][N] running sums
    T* rms_ring;          // [S][ceil(N/32)][width][32] last `width` values of y^2 (HBM / L2 resident)
    int width, hop;
    long env_rows;
    long t_base;          // first sample of this launch within the block
    long long tick0;      // ticks rendered before this block (ring position)
    // ENV == 2, replica followers: y[t-width] is regenerated by a second copy of every filter that runs
    // `width` ticks behind (same arithmetic, same input => same bits), instead of being stored
    const T* in_lag;      // per stream: in_lag[t] = x[t - width] for launch-relative t (t = -1, -2 valid)
    long lag_stride;
    T* lag_state;         // [3][S][N]: y1, y2, p of the replicas
    const T* lag_par;     // [7][N] the parameters the replicas run with: what the filters ran with `width` ticks ago
    int lag_taps;         // tap structure of lag_par
};

// ENV: 0 no followers; 1 followers with the ring of y^2 in HBM; 2 followers with lagged replicas
template <typename T, bool FMA, bool BANDS, int ENV>
__global__ void __launch_bounds__(ENV == 1 ? (ru::W_MAX_ENV + ru::s_warps(ru::W_MAX_ENV) + 1) * 32 : (ru::TermTiles<FMA, BANDS, ENV>::value ? ru::CTA_WARPS_MAX : ru::W_MAX + ru::NS_MAX + 1) * 32, 1)
fb_exact_ru_kernel(FbRuParams<T> Q) {
    using namespace ru;
    typedef Ar<T, FMA> A;
    constexpr int RS = RowStride<T>::RS;
    constexpr bool VF = TermTiles<FMA, BANDS, ENV>::value;   // tiles hold the mixdown terms, sample-major
    constexpr int NB = VF ? ru::NB_VF : ru::NB;              // tile slots per RU warp
    const FbK1Params<T>& P = Q.k;
    const int W = Q.W;
    const int NS = Q.NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout<T> lay(W, NS, ENV == 2, VF);
    T* xs = reinterpret_cast<T*>(smem_raw + lay.xs_off());
    T* xls = reinterpret_cast<T*>(smem_raw + lay.xl_off());          // ENV == 2: lagged input ring
    T* ytiles = reinterpret_cast<T*>(smem_raw + lay.y_off());
    T* gtiles = reinterpret_cast<T*>(smem_raw + lay.gt_off());
    T* gsmem = reinterpret_cast<T*>(smem_raw + lay.gs_off());
    T* psum = reinterpret_cast<T*>(smem_raw + lay.ps_off());
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + lay.bar_off());
    unsigned long long* x_full = bars;                      // [XB]
    unsigned long long* x_empty = bars + XB;                // [XB]
    // Tile hand-over barriers of RU warp w: tile k uses pair k % YB.  y tiles: YB = NB, one pair per tile slot.  Term
    // tiles: YB = NS, one pair per S warp (S warp j takes the tiles k = j mod NS, so every barrier is waited on by one
    // warp, phase after phase — a waiter that skipped a phase would misread the parity); the slot is still k % NB.
    const int YB = VF ? NS : NB;
    unsigned long long* y_full = bars + 2 * XB;                  // [W][YB]
    unsigned long long* y_empty = bars + 2 * XB + W * YB_MAX;    // [W][YB]
    unsigned long long* ps_full = bars + 2 * XB + 2 * W * YB_MAX;    // [NS][PD] running row of S warp i ready for S warp i+1
    unsigned long long* ps_empty = ps_full + PD * NS;                // [NS][PD] ... consumed

#ifdef HG_EXP_S_FIRST
    const int lane = threadIdx.x & 31, hw_warp = threadIdx.x >> 5;
    const int warp = hw_warp <= NS ? W + hw_warp : hw_warp - (NS + 1);       // experiment: S and G warps get the lowest warp ids
#else
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#endif
    // chain position = arrival order (chain.cuh): CTA (s, cta) owns the filter groups [cta W, cta W + W) of stream s
    const unsigned ticket = chain::take_ticket(P.chain);
    const int cta = (int)(ticket % gridDim.x), s = (int)(ticket / gridDim.x);
    const int N = P.N;
    const long L = P.L;
    const int groups = (N + 31) / 32;                       // RU warps of the whole stream
    const int w_first = cta * W;                            // first filter group of this CTA
    const int w_here = min(W, groups - w_first);            // RU warps with work
    const size_t SN = (size_t)P.S * N;
    const int ntiles = (int)((L + TILE - 1) / TILE);
    const int nchunks = (int)((L + XT - 1) / XT);
    const int last_tn = (int)(L - (long)(ntiles - 1) * TILE);
    constexpr int TPC = XT / TILE;

    if (threadIdx.x == 0) {
        for (int i = 0; i < XB; i++) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], w_here); }
        for (int i = 0; i < W * YB; i++) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < PD * NS; i++) { mbar_init(&ps_full[i], 1); mbar_init(&ps_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // the chain (chain.cuh): where this CTA's running bank-sum row comes from
    const chain::Params& C = P.chain;
    const int G = gridDim.x;
    const bool summing = P.out != nullptr || C.outbox != nullptr;
    const bool ordered = C.ordered != 0;
    const bool head = !ordered || cta == 0, tail = !ordered || cta == G - 1;
    const bool has_in = ordered && summing && (!head || C.inbox != nullptr);
    constexpr int WPS = chain::Words<T>::WPS;
    // Term tiles, long chains: the links between the CTAs of this launch are polled by the S warps themselves (see the S
    // role); short chains of many streams (BASELINE configs[3]: 4 CTAs per stream) keep the gateway warp, whose
    // prefetch costs the few S warps of those CTAs nothing.
    const bool dlinks = VF && G > 8;

    if (warp == W + NS) {
        // ------------------------------ G: input loader + chain gateway ------------------------------
        // One polling loop, never blocked on either duty: an input chunk is issued (TMA bulk copy) as soon as every
        // RU warp has released its ring slot — the RU warps of the later S warps run a few tiles behind those of
        // the first, so whoever waits for that must not be a stage of the chain; the inbound packets of the next
        // two tiles are kept in flight and land in shared memory as soon as their tags are complete.
        const T* inp = P.in + (size_t)s * P.in_stride;
        const T* inlp = ENV == 2 ? Q.in_lag + (size_t)s * Q.lag_stride : nullptr;
        chain::Port<T, 1> pin;
        if (has_in) {
            if (!head) pin.open(C.ring + (((size_t)s * G + cta) << C.ring_log2) * WPS, C.credit + (size_t)s * G + cta, C);
            else pin.open(C.inbox + ((size_t)s << C.ring_log2) * WPS, C.inbox_credit + s, C);
        }
        typename chain::Port<T, 1>::Raw rq[GD];            // polls in flight: tiles gk .. gk + GD - 1
        // term tiles: the S warps poll the links between CTAs themselves (below); the gateway only serves a device's inbox
        const int gtiles = (has_in && (!dlinks || head)) ? ntiles : 0;
#pragma unroll
        for (int i = 0; i < GD; i++)
            if (i < gtiles && (long)i * TILE + lane < L) pin.issue(C, (long)i * TILE + lane, rq[i]);
        unsigned long long* gfull = ps_full + (NS - 1) * PD;        // the hand-over slots of the (unused) link after the last S warp
        unsigned long long* gempty = ps_empty + (NS - 1) * PD;
        T* gslot = psum + (size_t)(NS - 1) * PD * 32;
        int c = 0, gk = 0;
        unsigned spins = 0;
#ifdef HG_PROF_G
        long long g_slot = 0, g_data = 0, g_t = clock64();
#endif
        while (c < nchunks || gk < gtiles) {
            bool progress = false;
#ifdef HG_PROF_G
            { const long long now = clock64();
              if (gk < gtiles) { if (!mbar_poll(&gempty[gk % PD], (unsigned)(((gk / PD) & 1) ^ 1))) g_slot += now - g_t; else g_data += now - g_t; }
              g_t = now; }
#endif
            const int b = c % XB;
            if (c < nchunks && mbar_poll(&x_empty[b], (unsigned)(((c / XB) & 1) ^ 1))) {
                const long base = (long)c * XT;
                const int cnt = (int)min((long)XT, L - base);
                const T* src = inp + base;
                const T* srcl = ENV == 2 ? inlp + base : src;
                const unsigned bytes = (unsigned)(cnt * sizeof(T));
                if ((((uintptr_t)src) & 15) == 0 && (((uintptr_t)srcl) & 15) == 0 && (bytes & 15) == 0) {
                    if (lane == 0) {
                        mbar_arrive_tx(&x_full[b], ENV == 2 ? 2 * bytes : bytes);
                        bulk_g2s(xs + b * XT, src, bytes, &x_full[b]);
                        if (ENV == 2) bulk_g2s(xls + b * XT, srcl, bytes, &x_full[b]);
                    }
                } else {
                    for (int i = lane; i < cnt; i += 32) {
                        xs[b * XT + i] = src[i];
                        if (ENV == 2) xls[b * XT + i] = srcl[i];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&x_full[b]);
                }
                c++;
                progress = true;
            }
            const int pb = gk % PD;
            if (gk < gtiles && mbar_poll(&gempty[pb], (unsigned)(((gk / PD) & 1) ^ 1))) {
                const long t = (long)gk * TILE + lane;
                const bool got = !(t < L) || pin.ready(C, t, rq[0]);
                if (__all_sync(0xffffffffu, got)) {
                    T v[1];
                    memcpy(v, rq[0].w, sizeof(v));
                    gslot[pb * 32 + lane] = v[0];
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&gfull[pb]);
                        pin.release(C, (long)(gk + 1) * TILE);       // the packet is in shared memory: its ring slot is free
                    }
                    gk++;
#pragma unroll
                    for (int i = 0; i + 1 < GD; i++) rq[i] = rq[i + 1];
                    if (gk + GD - 1 < gtiles && (long)(gk + GD - 1) * TILE + lane < L) pin.issue(C, (long)(gk + GD - 1) * TILE + lane, rq[GD - 1]);
                    progress = true;
                } else if (!got) {
                    // poll again (this lane's packet only).  More polls in flight do not help: strong loads of one warp do
                    // not overlap, so a gateway takes at most one tile per L2 round trip (measured: deeper windows are slower)
                    pin.issue(C, t, rq[0]);
                }
            }
            if (progress) spins = 0;
            else {
#ifndef HG_RU_GSLEEP
#define HG_RU_GSLEEP 32
#endif
                __nanosleep(HG_RU_GSLEEP);          // nothing to do right now: leave the issue and shared-memory pipes to the other roles
                if (++spins > chain::SPIN_LIMIT / 8) { *C.error = 1; break; }
            }
        }
#ifdef HG_PROF_G
        if (HG_PROF_ON(P) && lane == 0) { long long* q = P.prof + ((size_t)s * gridDim.x + cta) * 8; q[6] = g_slot; q[7] = g_data; }
#endif
        return;
    }
    if (warp < W) {
        // ------------------------------ RU: taps, pre-gain, recurrence ------------------------------
        if (warp >= w_here) return;
        const int n = (w_first + warp) * 32 + lane;
        const bool ok = n < N;
        const size_t si = (size_t)s * N + n;
        const T b0 = ok ? P.par[0 * (size_t)N + n] : (T)0, b1 = ok ? P.par[1 * (size_t)N + n] : (T)0;
        const T b2 = ok ? P.par[2 * (size_t)N + n] : (T)0;
        const T a1 = ok ? P.par[3 * (size_t)N + n] : (T)0, a2 = ok ? P.par[4 * (size_t)N + n] : (T)0;
        const T pin = ok ? P.par[5 * (size_t)N + n] : (T)0;
        T y1 = ok ? P.state[0 * SN + si] : (T)0, y2 = ok ? P.state[1 * SN + si] : (T)0;
        T p = ok ? P.state[2 * SN + si] : (T)0;
        const T cp = A::mul(P.c_p, pin), s_p = P.s_p;
        T x1 = P.xh[2 * s], x2 = P.xh[2 * s + 1];
        T m1 = A::mul(b0, x1), m2 = A::mul(b0, x2);         // taps == 2: fl(b0 x[t-1]), fl(b0 x[t-2])
        const int taps = P.taps;
        bool conv = __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p);
        // VF: this warp owns the mix gain smoother too (filterbank.h:173) and writes v = fl(y*g)
        T g = (VF && ok) ? P.state[3 * SN + si] : (T)0;
        const T cg = VF ? A::mul(P.c_g, ok ? P.par[6 * (size_t)N + n] : (T)0) : (T)0, s_g = P.s_g;
        if (VF) conv = conv && __all_sync(0xffffffffu, A::muladd(s_g, g, cg) == g);     // `conv`: both smoothers at their fixed point
        T* ybase = ytiles + (size_t)warp * NB * 32 * RS + (VF ? lane : lane * RS);
        unsigned long long* yf = y_full + warp * YB;
        unsigned long long* ye = y_empty + warp * YB;
        int xb = 0, yb = 0, kc = 0;
        unsigned xpar = 0;
        // tile k: slot yb = k % NB; filled -> yf[fi], fi = k % YB; before writing, the slot's previous tile k - NB must have
        // been taken: ye[ei], ei = (k - NB) % YB, phase (k - NB) / YB (a fresh barrier passes the wait for "phase -1")
        int fi = 0, ei = (YB - NB % YB) % YB;
        unsigned epar = 1;
        T t2 = A::mul(a2, y2);
        // ENV: envelope follower state of this lane
        T rsum = (ENV && ok) ? Q.rms_sum[si] : (T)0;
        const T wT = (T)Q.width;
        // ring layout [S][filter group of 32][width][32]: a warp's consecutive positions are consecutive 128-byte
        // lines, so a tile of 32 samples is one contiguous 4 KB burst (DRAM-row friendly)
        T* ring = ENV == 1 ? Q.rms_ring + (((size_t)s * groups + (w_first + warp)) * Q.width) * 32 + lane : nullptr;
        int rpos = ENV == 1 ? (int)((Q.tick0 + Q.t_base) % Q.width) : 0;
        // ENV == 2: the replica of this lane, idle until the stream is `width` ticks old
        T y1l = (T)0, y2l = (T)0, pl = (T)0, x1l = (T)0, x2l = (T)0, m1l = (T)0, m2l = (T)0, t2l = (T)0;
        bool convl = false, was_idle = false;
        long lag_left = 0;
        const T* inl = nullptr;
        // the replicas' own parameter set: a change reaches them `width` ticks after it reached the filters
        T b0l = b0, b1l = b1, b2l = b2, a1l = a1, a2l = a2, cpl = cp;
        int tapsl = taps;
        if (ENV == 2) {
            if (ok) {
                y1l = Q.lag_state[0 * SN + si]; y2l = Q.lag_state[1 * SN + si]; pl = Q.lag_state[2 * SN + si];
                b0l = Q.lag_par[0 * (size_t)N + n]; b1l = Q.lag_par[1 * (size_t)N + n]; b2l = Q.lag_par[2 * (size_t)N + n];
                a1l = Q.lag_par[3 * (size_t)N + n]; a2l = Q.lag_par[4 * (size_t)N + n];
                cpl = A::mul(P.c_p, Q.lag_par[5 * (size_t)N + n]);
            }
            tapsl = Q.lag_taps;
            inl = Q.in_lag + (size_t)s * Q.lag_stride;
            const long long started = Q.tick0 + Q.t_base - Q.width;      // replica ticks already done (< 0: idle)
            lag_left = started < 0 ? (long)(-started) : 0;
            x1l = inl[-1]; x2l = inl[-2];
            m1l = A::mul(b0l, x1l); m2l = A::mul(b0l, x2l); t2l = A::mul(a2l, y2l);
            convl = __all_sync(0xffffffffu, A::muladd(s_p, pl, cpl) == pl);
        }
        long hop_left = ENV ? (long)Q.hop - (Q.t_base % Q.hop) : 0;        // samples until the next envelope row
        long env_row = ENV ? Q.t_base / Q.hop : 0;
        // ENV: the delayed squares q[t-width] live in one register buffer that is refilled in place: as soon as
        // slot j of tile k has been consumed, slot j of tile k+1 is requested, so every ring load has a whole
        // tile of arithmetic to land (width >= 2 * TILE keeps the two tiles' positions apart)
        T qo[ENV == 1 ? TILE : 1];
        if (ENV == 1 && ok) {
            const int cnt0 = (ntiles == 1) ? last_tn : TILE;
#pragma unroll
            for (int j = 0; j < (ENV == 1 ? TILE : 1); j++) {
                int pj = rpos + j;
                if (pj >= Q.width) pj -= Q.width;
                qo[j] = (j < cnt0) ? ring[(size_t)pj * 32] : (T)0;
            }
        }
        long long pf_gate = 0, pf_x = 0;
        const long long pf_t0 = HG_PROF_ON(P) ? clock64() : 0;
#ifdef HG_EXP_HEAD_NS
        const bool throttle = ordered && summing && head && G > 1;
#define HG_HEAD_THROTTLE if (throttle) __nanosleep(HG_EXP_HEAD_NS);
#else
#define HG_HEAD_THROTTLE
#endif
        for (int k = 0; k < ntiles; k++) {
            if (ENV == 0 && conv && kc == 0 && (long)(k + TPC) * TILE <= L) {
                // steady state: a whole input chunk of full tiles with converged smoothers.  Nothing but the two
                // hand-over barriers per tile around the unrolled recurrence (the general path below spends ~75
                // instructions per tile on cases that cannot occur here).
                { const long long pc0 = HG_PROF_ON(P) ? clock64() : 0; mbar_wait(&x_full[xb], xpar); if (HG_PROF_ON(P)) pf_x += clock64() - pc0; }
                const T* xcb = xs + xb * XT;
#define HG_RU_CHUNK(TP)                                                                                          \
                _Pragma("unroll 1") for (int i = 0; i < TPC; i++) {                                             \
                    const long long pc0 = HG_PROF_ON(P) ? clock64() : 0;                                               \
                    mbar_wait(&ye[ei], epar);                                                                   \
                    if (HG_PROF_ON(P)) pf_gate += clock64() - pc0;                                                     \
                    HG_HEAD_THROTTLE                                                                            \
                    if (VF) ru_tile_v<T, FMA, TP>(xcb + i * TILE, ybase + yb * 32 * RS, b0, b1, b2, a1, a2, p, g, y1, y2, x1, x2, m1, m2, t2); \
                    else ru_tile<T, FMA, TP>(xcb + i * TILE, ybase + yb * 32 * RS, b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2); \
                    __syncwarp();                                                                               \
                    if (lane == 0) mbar_arrive(&yf[fi]);                                                        \
                    if (++yb == NB) yb = 0;                                                                     \
                    if (++fi == YB) fi = 0;                                                                     \
                    if (++ei == YB) { ei = 0; epar ^= 1u; }                                                     \
                }
                if (taps == 2) { HG_RU_CHUNK(2) } else if (taps == 1) { HG_RU_CHUNK(1) } else { HG_RU_CHUNK(0) }
#undef HG_RU_CHUNK
                if (lane == 0) mbar_arrive(&x_empty[xb]);
                if (++xb == XB) { xb = 0; xpar ^= 1u; }
                k += TPC - 1;
                continue;
            }
            if (ENV == 2 && conv && convl && taps == tapsl && kc == 0 && lag_left == 0 && !was_idle && (long)(k + TPC) * TILE <= L &&
                Q.hop % TILE == 0 && hop_left % TILE == 0) {
                // the same steady state with replica followers: filter + replica + follower per tile, envelope rows
                // on tile ends
                mbar_wait(&x_full[xb], xpar);
                const T* xcb = xs + xb * XT;
                const T* xlcb = xls + xb * XT;
#define HG_RU_CHUNK_ENV(TP)                                                                                      \
                _Pragma("unroll 1") for (int i = 0; i < TPC; i++) {                                             \
                    mbar_wait(&ye[ei], epar);                                                                   \
                    T unused = (T)0;                                                                            \
                    ru_tile_env<T, FMA, TP, false, VF>(xcb + i * TILE, xlcb + i * TILE, ybase + yb * 32 * RS, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, \
                                                   y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, -1, unused); \
                    hop_left -= TILE;                                                                           \
                    if (hop_left == 0) {                                                                        \
                        if (ok && Q.env) Q.env[((size_t)s * Q.env_rows + env_row) * N + n] = A::sqrt(A::div(rsum, wT)); \
                        env_row++;                                                                              \
                        hop_left = Q.hop;                                                                       \
                    }                                                                                           \
                    __syncwarp();                                                                               \
                    if (lane == 0) mbar_arrive(&yf[fi]);                                                        \
                    if (++yb == NB) yb = 0;                                                                     \
                    if (++fi == YB) fi = 0;                                                                     \
                    if (++ei == YB) { ei = 0; epar ^= 1u; }                                                     \
                }
                if (taps == 2) { HG_RU_CHUNK_ENV(2) } else if (taps == 1) { HG_RU_CHUNK_ENV(1) } else { HG_RU_CHUNK_ENV(0) }
#undef HG_RU_CHUNK_ENV
                if (lane == 0) mbar_arrive(&x_empty[xb]);
                if (++xb == XB) { xb = 0; xpar ^= 1u; }
                k += TPC - 1;
                continue;
            }
            const int tn = (k == ntiles - 1) ? last_tn : TILE;
            T ykeep[(VF && ENV == 2) ? TILE : 1];      // slow path of the replica followers only (local memory)
            if (kc == 0) mbar_wait(&x_full[xb], xpar);
            mbar_wait(&ye[ei], epar);
            const T* xc = xs + xb * XT + kc * TILE;
            T* yrow = ybase + yb * 32 * RS;
            int tt = 0;
            bool fused = false;
            if (ENV == 2 && conv && convl && taps == tapsl && lag_left == 0 && !was_idle && tn == TILE && Q.hop >= TILE) {
                // filter + replica + follower in one unrolled tile
                const T* xlc = xls + xb * XT + kc * TILE;
                const int je = hop_left <= TILE ? (int)hop_left - 1 : -1;
                T s_je = (T)0;
                if (je < 0 || je == TILE - 1) {
                    // no row, or the row falls on the tile's last tick: S at the end of the tile is what is needed
                    if (taps == 2) ru_tile_env<T, FMA, 2, false, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                    else if (taps == 1) ru_tile_env<T, FMA, 1, false, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                    else ru_tile_env<T, FMA, 0, false, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                    s_je = rsum;
                } else {
                    if (taps == 2) ru_tile_env<T, FMA, 2, true, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                    else if (taps == 1) ru_tile_env<T, FMA, 1, true, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                    else ru_tile_env<T, FMA, 0, true, VF>(xc, xlc, yrow, g, b0, b1, b2, a1, a2, p, b0l, b1l, b2l, a1l, a2l, pl, y1, y2, x1, x2, m1, m2, t2, y1l, y2l, x1l, x2l, m1l, m2l, t2l, rsum, je, s_je);
                }
                if (je >= 0) {
                    if (ok && Q.env) Q.env[((size_t)s * Q.env_rows + env_row) * N + n] = A::sqrt(A::div(s_je, wT));   // rmsbank.h:64
                    env_row++;
                    hop_left += Q.hop;
                }
                hop_left -= TILE;
                tt = TILE;
                fused = true;
            } else if (conv && tn == TILE && !(VF && ENV == 2)) {
                // whole tile unrolled; `taps` is uniform, so pick the specialised body once per tile
                if (VF) {
                    if (taps == 2) ru_tile_v<T, FMA, 2>(xc, yrow, b0, b1, b2, a1, a2, p, g, y1, y2, x1, x2, m1, m2, t2);
                    else if (taps == 1) ru_tile_v<T, FMA, 1>(xc, yrow, b0, b1, b2, a1, a2, p, g, y1, y2, x1, x2, m1, m2, t2);
                    else ru_tile_v<T, FMA, 0>(xc, yrow, b0, b1, b2, a1, a2, p, g, y1, y2, x1, x2, m1, m2, t2);
                }
                else if (taps == 2) ru_tile<T, FMA, 2>(xc, yrow, b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
                else if (taps == 1) ru_tile<T, FMA, 1>(xc, yrow, b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
                else ru_tile<T, FMA, 0>(xc, yrow, b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
                tt = TILE;
            } else if (conv && !VF) {
                for (; tt + 4 <= tn; tt += 4) {
                    T xv[4], yy[4];
                    ld4<T>(xc + tt, xv);
#pragma unroll
                    for (int j = 0; j < 4; j++) yy[j] = ru_sample<T, FMA, 0>(xv[j], b0, b1, b2, a1, a2, p, y1, y2, x1, x2, m1, m2, t2);
                    st4<T>(yrow + tt, yy);
                }
            }
            for (; tt < tn; tt++) {
                if (!conv) p = A::muladd(s_p, p, cp);                              // filterbank.h:172
                if (VF && !conv) g = A::muladd(s_g, g, cg);                        // filterbank.h:173
                const T x0 = xc[tt];
                T fw = A::mul(b0, x0);
                fw = A::muladd(b1, x1, fw);
                fw = A::muladd(b2, x2, fw);
                T fb = A::mul(a1, y1);
                fb = A::muladd(a2, y2, fb);
                const T y = A::mulsub(fw, p, fb);
                y2 = y1; y1 = y;
                x2 = x1; x1 = x0;
                if (VF) yrow[tt * RS] = A::mul(y, g);                              // filterbank.h:130  y_n * g_n
                else yrow[tt] = y;
                if (VF && ENV == 2) ykeep[tt] = y;                                 // the follower below needs y itself
            }
            if (!conv || tn != TILE || (VF && ENV == 2 && !fused)) { m1 = A::mul(b0, x1); m2 = A::mul(b0, x2); t2 = A::mul(a2, y2); }   // the per-sample loop ran
            if (!conv) conv = __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p && (!VF || A::muladd(s_g, g, cg) == g));
            if (ENV == 2 && !fused) {
                // replica + follower, sample by sample (stream start, smoothers still moving, ragged tile, hop < TILE)
                const T* xlc = xls + xb * XT + kc * TILE;
                int j0 = 0;                                                       // ticks of this tile before the replica starts
                if (lag_left > 0) { j0 = (int)min((long)tn, lag_left); lag_left -= j0; was_idle = true; }
                if (was_idle && j0 < tn) {
                    // the replica starts in this tile: its input history is what precedes its first sample
                    const long t1 = (long)k * TILE + j0;                          // launch-relative tick of that sample
                    x1l = inl[t1 - 1]; x2l = inl[t1 - 2];
                    convl = __all_sync(0xffffffffu, A::muladd(s_p, pl, cpl) == pl);
                    was_idle = false;
                }
                for (int j = 0; j < tn; j++) {
                    T yl = (T)0;                                                  // idle: y[t - width] = 0 (rmsbank.h:36-44 setZero)
                    if (j >= j0) {
                        if (!convl) pl = A::muladd(s_p, pl, cpl);
                        const T x0 = xlc[j];
                        T fw = A::mul(b0l, x0);
                        fw = A::muladd(b1l, x1l, fw);
                        fw = A::muladd(b2l, x2l, fw);
                        T fbk = A::mul(a1l, y1l);
                        fbk = A::muladd(a2l, y2l, fbk);
                        yl = A::mulsub(fw, pl, fbk);
                        y2l = y1l; y1l = yl;
                        x2l = x1l; x1l = x0;
                    }
                    const T y = VF ? ykeep[j] : yrow[j];
                    rsum = A::add(A::sub(A::mul(y, y), A::mul(yl, yl)), rsum);     // rmsbank.h:58-61
                    if (--hop_left == 0) {
                        if (ok && Q.env) Q.env[((size_t)s * Q.env_rows + env_row) * N + n] = A::sqrt(A::div(rsum, wT));   // :64
                        env_row++;
                        hop_left = Q.hop;
                    }
                }
                m1l = A::mul(b0l, x1l); m2l = A::mul(b0l, x2l); t2l = A::mul(a2l, y2l);
                if (!convl && lag_left == 0 && !was_idle) convl = __all_sync(0xffffffffu, A::muladd(s_p, pl, cpl) == pl);
            }
            if (ENV == 1 && ok) {
                // envelope follower over the tile just rendered (own row of the y tile), rmsbank.h:58-64
                const int tn_next = (k + 1 >= ntiles) ? 0 : ((k + 1 == ntiles - 1) ? last_tn : TILE);
                if (tn == TILE && tn_next == TILE && rpos + 2 * TILE <= Q.width && Q.hop >= TILE) {
                    // fast path: full tile, full next tile, no ring wrap, at most one envelope row due: every ring
                    // access is base + immediate offset
                    const int je = hop_left <= TILE ? (int)hop_left - 1 : -1;
                    T* base = ring + (size_t)rpos * 32;
#pragma unroll
                    for (int g4 = 0; g4 < TILE / 4; g4++) {
                        T yv[4];
                        ld4<T>(yrow + 4 * g4, yv);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int j = 4 * g4 + i;
                            const T q = A::muladd((T)0, (T)0, A::mul(yv[i], yv[i]));
                            rsum = A::add(A::sub(q, qo[j]), rsum);
                            base[j * 32] = q;
                            qo[j] = base[(TILE + j) * 32];
                            if (j == je && Q.env) Q.env[((size_t)s * Q.env_rows + env_row) * N + n] = A::sqrt(A::div(rsum, wT));   // :64
                        }
                    }
                    if (je >= 0) { env_row++; hop_left += Q.hop; }
                    hop_left -= TILE;
                } else {
#pragma unroll
                    for (int j = 0; j < TILE; j++) {
                        int pj = rpos + j;
                        if (pj >= Q.width) pj -= Q.width;
                        if (j < tn) {
                            const T y = yrow[j];
                            const T q = A::muladd((T)0, (T)0, A::mul(y, y));       // |z|^2 of a real sample: y*y + 0*0
                            rsum = A::add(A::sub(q, qo[j]), rsum);                 // (new - old) + last, :60
                            ring[(size_t)pj * 32] = q;
                            if (--hop_left == 0) {
                                if (Q.env) Q.env[((size_t)s * Q.env_rows + env_row) * N + n] = A::sqrt(A::div(rsum, wT));   // :64
                                env_row++;
                                hop_left = Q.hop;
                            }
                        }
                        if (j < tn_next) {                                         // refill slot j for the next tile
                            int pn = pj + TILE;
                            if (pn >= Q.width) pn -= Q.width;
                            qo[j] = ring[(size_t)pn * 32];
                        }
                    }
                }
                rpos += tn;
                if (rpos >= Q.width) rpos -= Q.width;
            }
            __syncwarp();
            const bool chunk_done = (kc == TPC - 1) || (k == ntiles - 1);
            if (lane == 0) {
                mbar_arrive(&yf[fi]);
                if (chunk_done) mbar_arrive(&x_empty[xb]);
            }
            if (++yb == NB) yb = 0;
            if (++fi == YB) fi = 0;
            if (++ei == YB) { ei = 0; epar ^= 1u; }
            if (++kc == TPC) { kc = 0; if (++xb == XB) { xb = 0; xpar ^= 1u; } }
        }
        if (ok) {
            P.state[0 * SN + si] = y1; P.state[1 * SN + si] = y2; P.state[2 * SN + si] = p;
            if (VF) P.state[3 * SN + si] = g;
            if (ENV) Q.rms_sum[si] = rsum;
            if (ENV == 2) { Q.lag_state[0 * SN + si] = y1l; Q.lag_state[1 * SN + si] = y2l; Q.lag_state[2 * SN + si] = pl; }
        }
        if (cta == 0 && warp == 0 && lane == 0 && L > 0) { P.xh_out[2 * s] = x1; P.xh_out[2 * s + 1] = x2; }
#ifdef HG_PROF_G
        if (HG_PROF_ON(P) && lane == 0 && warp == 0) {
#else
        if (HG_PROF_ON(P) && lane == 0 && (warp == 0 || warp == w_here - 1)) {
#endif
            long long* q = P.prof + ((size_t)s * gridDim.x + cta) * 8 + (warp == 0 ? 4 : 6);
            q[0] = pf_gate; q[1] = warp == 0 ? pf_x : clock64() - pf_t0;
#ifdef HG_PROF_G
            { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); q[1] = (long long)smid * L; }
#endif
        }
    } else {
        // ------------------------------ S: mixdown ------------------------------
        const int sidx = warp - W;                          // 0 .. NS-1
        if (sidx >= NS) return;
        if constexpr (VF) {
            // Term tiles.  S warp j adds ALL the CTA's filters, in index order, onto the running row of the tiles
            // k = j (mod NS): it walks the RU warps' tiles of that tile index (lane = sample, this lane's 32 terms of an
            // RU warp are one row), the terms fetched one batch ahead of the additions and every tile slot given back
            // as soon as its terms are in registers.  The CTA is ONE stage of the chain — no hand-over between S
            // warps; the row comes from the gateway's slot k % PD and leaves for the next CTA / device / `out`.
            constexpr int CH = 64 / (int)sizeof(T), NBAT = 32 / CH;     // terms per batch, batches per RU warp
            // Links between CTAs of this launch are read by the S warp that needs the tile (NS polls in flight per link: a
            // poll is an L2 round trip of 1-2 k cycles and strong loads of one warp do not overlap, so a single polling
            // warp would cap the chain at a tile per round trip) — each reading warp publishes its own credit word, and
            // the warp that writes tile k checks the word of the warp that read tile k - ring.  A device's inbox is
            // still served by the gateway warp (one credit word: the writer may be another kernel).
            const bool direct = dlinks && ordered && summing && !head;
            const bool via_g = has_in && !direct;
            chain::Port<T, 1> pin, pout;
            typename chain::Port<T, 1>::Raw raw;
            if (direct) pin.open(C.ring + (((size_t)s * G + cta) << C.ring_log2) * WPS, C.credit_w + ((size_t)s * G + cta) * chain::CW + sidx, C);
            if (ordered && summing) {
                if (!tail) {
                    const int ring_tiles = (1 << C.ring_log2) / TILE;
                    const int reader = ((sidx - ring_tiles % NS) % NS + NS) % NS;       // the next CTA's warp that read tile k - ring_tiles
                    pout.open(C.ring + (((size_t)s * G + cta + 1) << C.ring_log2) * WPS,
                              dlinks ? C.credit_w + ((size_t)s * G + cta + 1) * chain::CW + reader : C.credit + (size_t)s * G + cta + 1, C);
                } else if (C.outbox) pout.open(C.outbox + ((size_t)s << C.ring_log2) * WPS, C.outbox_credit + s, C);
            }
            T* out_row = nullptr;
            if (P.out && tail && !pout.on())
                out_row = ordered ? (P.out + (size_t)s * P.out_stride) : (P.out + ((size_t)cta * P.S + s) * (size_t)L);   // tree mode: one partial row per CTA
            const T* carry = (ordered && head && !has_in && P.carry_in) ? P.carry_in + (size_t)s * P.carry_stride : nullptr;
            const bool dist = P.dist != 0;
            const int gl = NS - 1;                          // the gateway's hand-over slots
            const int nb = w_here * NBAT;                   // batches per tile (even)
            long long pf_in = 0, pf_y = 0, pf_mix = 0, pf_out = 0;
            if (direct && sidx < ntiles && (long)sidx * TILE + lane < L) pin.issue(C, (long)sidx * TILE + lane, raw);
#pragma unroll 1
            for (int k = sidx; k < ntiles; k += NS) {
                const int tn = (k == ntiles - 1) ? last_tn : TILE;
                const long tg = (long)k * TILE;
                const int yb = k % NB;                          // tile slot; barrier pair of this S warp: sidx = k % NS
                const unsigned ypar = (unsigned)((k / NS) & 1);
                const int pb = k % PD;
                const unsigned ppar = (unsigned)((k / PD) & 1);
                const T* row0 = ytiles + (size_t)yb * 32 * RS + lane * RS;      // this lane's row in RU warp 0's slot
                long long pc0 = HG_PROF_ON(P) ? clock64() : 0, pc1;
                auto fetch = [&](int b, T(&v)[CH]) {
                    const int w = b / NBAT, part = b % NBAT;
                    if (part == 0) {
                        if (HG_PROF_ON(P)) { const long long q0 = clock64(); mbar_wait(&y_full[w * YB + sidx], ypar); pf_y += clock64() - q0; }
                        else mbar_wait(&y_full[w * YB + sidx], ypar);
                    }
                    const T* src = row0 + (size_t)w * NB * 32 * RS + part * CH;
#pragma unroll
                    for (int q = 0; q < CH / 4; q++) ld4<T>(src + 4 * q, *reinterpret_cast<T(*)[4]>(&v[4 * q]));
                };
                auto landed = [&](int b, const T(&v)[CH]) {       // batch b is in registers: its tile slot is free after the last part
                    if (b % NBAT == NBAT - 1) {
                        reg_fence(v[CH - 1]);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&y_empty[(b / NBAT) * YB + sidx]);
                    }
                };
                T va[CH], vb[CH];
                if (!summing) {                                   // nothing to add: just recycle the slots
                    for (int w = 0; w < w_here; w++) {
                        mbar_wait(&y_full[w * YB + sidx], ypar);
                        if (lane == 0) mbar_arrive(&y_empty[w * YB + sidx]);
                    }
                    continue;
                }
                fetch(0, va);
                T acc = (T)0;
                if (HG_PROF_ON(P)) pc0 = clock64();
                if (direct) {
                    const long t = tg + lane;
                    bool got = !(t < L) || chain::Port<T, 1>::ready(C, t, raw);
                    unsigned spins = 0;
                    while (!__all_sync(0xffffffffu, got)) {
                        if (!got) { pin.issue(C, t, raw); got = chain::Port<T, 1>::ready(C, t, raw); }
                        if (++spins > chain::SPIN_LIMIT) { *C.error = 1; break; }
                    }
                    if (t < L) { T v[1]; memcpy(v, raw.w, sizeof(v)); acc = v[0]; }
                    __syncwarp();
                    if (lane == 0) pin.release(C, tg + TILE);
                    // ask for this warp's next row now: the poll is in flight during the whole walk (stale if the CTA before
                    // is not that far ahead; it is repeated above when the row is needed)
                    if (k + NS < ntiles && t + (long)NS * TILE < L) pin.issue(C, t + (long)NS * TILE, raw);
                } else if (via_g) {
                    mbar_wait(&ps_full[gl * PD + pb], ppar);
                    acc = psum[((size_t)gl * PD + pb) * 32 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ps_empty[gl * PD + pb]);
                } else if (carry) {
                    if (lane < tn) acc = carry[tg + lane];
                }
                if (HG_PROF_ON(P)) { pc1 = clock64(); pf_in += pc1 - pc0; pc0 = pc1; }
                // acc = fl(v_l[t] + acc), l ascending over the CTA's filters: src/filterbank.h:130 (:137 with the per-band
                // distortion).  Lanes past the bank's last filter wrote +-0 terms, which leave any sum unchanged.
#pragma unroll 1
                for (int b = 0; b < nb; b += 2) {
                    fetch(b + 1, vb);
                    if (dist) {
#pragma unroll
                        for (int l = 0; l < CH; l++) acc = A::add(acc, softclip_dev<T>(va[l]));
                    } else {
#pragma unroll
                        for (int l = 0; l < CH; l++) acc = A::add(va[l], acc);
                    }
                    landed(b + 1, vb);
                    if (b + 2 < nb) fetch(b + 2, va);
                    if (dist) {
#pragma unroll
                        for (int l = 0; l < CH; l++) acc = A::add(acc, softclip_dev<T>(vb[l]));
                    } else {
#pragma unroll
                        for (int l = 0; l < CH; l++) acc = A::add(vb[l], acc);
                    }
                }
                if (HG_PROF_ON(P)) { pc1 = clock64(); pf_mix += pc1 - pc0; pc0 = pc1; }
                if (pout.on()) {
                    pout.reserve(C, tg + TILE);
                    T a1v[1] = {acc};
                    pout.send(C, tg + lane, a1v);
                } else if (out_row) {
                    if (lane < tn) {
                        if (P.epi.kind && ordered) acc = epilogue_dev<T>(P.epi, acc, P.dry_in[(size_t)s * P.in_stride + tg + lane]);
                        out_row[tg + lane] = acc;
                    }
                }
                if (HG_PROF_ON(P)) pf_out += clock64() - pc0;
#ifdef HG_PROF_G
                if (HG_PROF_ON(P) && lane == 0 && sidx == 1 && ((k / NS) & 31) == 0 && (k / NS) / 32 < 128) {     // progress stamps: tile k done at time t
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    P.prof[(size_t)(1024 + cta * 16) * 8 + (k / NS) / 32] = (long long)now;
                }
#endif
            }
            if (HG_PROF_ON(P) && lane == 0 && sidx == (NS > 1 ? 1 : 0)) {
                long long* q = P.prof + ((size_t)s * gridDim.x + cta) * 8;
                q[0] = pf_in; q[1] = pf_y; q[2] = pf_mix; q[3] = pf_out;
            }
            return;
        }
        // ---- y tiles (fused arithmetic, per-band rows, ring followers): the S warps own the mix gains and form a chain ----
        // RU warps served by this S warp: an even split of the CTA's active RU warps
        const int per = (w_here + NS - 1) / NS;
        const int wa = min(w_here, sidx * per), wb = min(w_here, wa + per);
        T* gt = gtiles + (size_t)sidx * 32 * RS;
        T* gs = gsmem + (size_t)sidx * GPS * 32;            // converged gains, [served RU warp][filter]
        // lane l holds g, cg of filter l of every served RU warp
        T g[GPS], cg[GPS];
        bool gconv = true;
#pragma unroll
        for (int i = 0; i < GPS; i++) {
            const int w = wa + i;
            const int n = (w_first + w) * 32 + lane;
            const bool ok = w < wb && n < N;
            g[i] = ok ? P.state[3 * SN + (size_t)s * N + n] : (T)0;
            cg[i] = A::mul(P.c_g, ok ? P.par[6 * (size_t)N + n] : (T)0);
            gconv = gconv && (A::muladd(P.s_g, g[i], cg[i]) == g[i]);
            gs[i * 32 + lane] = g[i];
        }
        gconv = __all_sync(0xffffffffu, gconv);
        __syncwarp();
        // the running bank-sum row: in from the S warp before (S warp 0: from the gateway), out to the next S warp /
        // CTA / device
        const bool first = sidx == 0, last = sidx == NS - 1;
        const bool smem_in = !first || has_in;
        const int in_link = first ? NS - 1 : sidx - 1;
        chain::Port<T, 1> pout;
        if (ordered && summing && last) {
            if (!tail) pout.open(C.ring + (((size_t)s * G + cta + 1) << C.ring_log2) * WPS, C.credit + (size_t)s * G + cta + 1, C);
            else if (C.outbox) pout.open(C.outbox + ((size_t)s << C.ring_log2) * WPS, C.outbox_credit + s, C);
        }
        T* out_row = nullptr;
        if (P.out && last && tail && !pout.on())
            out_row = ordered ? (P.out + (size_t)s * P.out_stride) : (P.out + ((size_t)cta * P.S + s) * (size_t)L);   // tree mode: one partial row per CTA
        const T* carry = (ordered && first && head && !has_in && P.carry_in) ? P.carry_in + (size_t)s * P.carry_stride : nullptr;
        const int dist = P.dist;
        int yb = 0;
        unsigned ypar = 0;
        long long pf_in = 0, pf_y = 0, pf_mix = 0, pf_out = 0;
        for (int k = 0; k < ntiles; k++) {
            const int tn = (k == ntiles - 1) ? last_tn : TILE;
            const long tg = (long)k * TILE;
            T acc = (T)0;
            bool still = false;
            const int pb = k % PD;
            const unsigned ppar = (unsigned)((k / PD) & 1);
            long long pc0 = HG_PROF_ON(P) ? clock64() : 0, pc1;
            if (summing) {
                if (smem_in) {
                    // running row after the filters before this S warp's (left by S warp sidx-1, or by the gateway)
                    mbar_wait(&ps_full[in_link * PD + pb], ppar);
                    acc = psum[((size_t)in_link * PD + pb) * 32 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ps_empty[in_link * PD + pb]);
                } else if (carry) {
                    if (lane < tn) acc = carry[tg + lane];
                }
            }
            if (HG_PROF_ON(P)) { pc1 = clock64(); pf_in += pc1 - pc0; pc0 = pc1; }
#pragma unroll 1
            for (int w = wa; w < wb; w++) {
                const int i = w - wa;
                mbar_wait(&y_full[w * NB + yb], ypar);
                if (HG_PROF_ON(P)) { pc1 = clock64(); pf_y += pc1 - pc0; pc0 = pc1; }
                const T* yt = ytiles + ((size_t)w * NB + yb) * 32 * RS + lane;
                const int rows = min(32, N - (w_first + w) * 32);
                // acc = fl(fl(y_l[t] * g_l[t]) + acc) (fused in fma mode), l ascending: src/filterbank.h:130 — the same
                // function of (y, g) whether the gains still move or not
                if (!gconv) {
                    T gi = (T)0, cgi = (T)0;
#pragma unroll
                    for (int j = 0; j < GPS; j++) if (j == i) { gi = g[j]; cgi = cg[j]; }
                    // replay the gain smoother of filter `lane` over the tile (filterbank.h:173) into gt[l][t]
                    for (int t = 0; t < tn; t++) { gi = A::muladd(P.s_g, gi, cgi); gt[lane * RS + t] = gi; }
                    still = still || (A::muladd(P.s_g, gi, cgi) != gi);
#pragma unroll
                    for (int j = 0; j < GPS; j++) if (j == i) g[j] = gi;
                    gs[i * 32 + lane] = gi;
                    __syncwarp();
                    if (dist) {
                        for (int l = 0; l < rows; l++) acc = A::add(acc, softclip_dev<T>(A::mul(yt[l * RS], gt[l * RS + lane])));   // :137
                    } else {
#pragma unroll 8
                        for (int l = 0; l < rows; l++) acc = A::muladd(yt[l * RS], gt[l * RS + lane], acc);
                    }
                    __syncwarp();
                } else {
                    const T* gw = gs + i * 32;
                    if (dist) {
                        for (int l = 0; l < rows; l++) acc = A::add(acc, softclip_dev<T>(A::mul(yt[l * RS], gw[l])));               // :137
                    } else if (rows == 32) {
#pragma unroll
                        for (int l = 0; l < 32; l += 4) {
                            T g4[4];
                            ld4<T>(gw + l, g4);                                  // broadcast, one LDS per 4 filters
                            acc = A::muladd(yt[(l + 0) * RS], g4[0], acc);
                            acc = A::muladd(yt[(l + 1) * RS], g4[1], acc);
                            acc = A::muladd(yt[(l + 2) * RS], g4[2], acc);
                            acc = A::muladd(yt[(l + 3) * RS], g4[3], acc);
                        }
                    } else {
                        for (int l = 0; l < rows; l++) acc = A::muladd(yt[l * RS], gw[l], acc);
                    }
                }
                if (BANDS) {
                    for (int l = 0; l < rows; l++)
                        if (lane < tn) P.bands[((size_t)s * N + (w_first + w) * 32 + l) * P.bands_stride + tg + lane] = yt[l * RS];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&y_empty[w * NB + yb]);
                if (HG_PROF_ON(P)) { pc1 = clock64(); pf_mix += pc1 - pc0; pc0 = pc1; }
            }
            if (!gconv) gconv = !__any_sync(0xffffffffu, still);
            if (summing) {
                if (!last) {
                    mbar_wait(&ps_empty[sidx * PD + pb], ppar ^ 1u);
                    psum[((size_t)sidx * PD + pb) * 32 + lane] = acc;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ps_full[sidx * PD + pb]);
                } else if (pout.on()) {
                    pout.reserve(C, tg + TILE);
                    T a1v[1] = {acc};
                    pout.send(C, tg + lane, a1v);
                } else if (out_row) {
                    if (lane < tn) {
                        if (P.epi.kind && ordered) acc = epilogue_dev<T>(P.epi, acc, P.dry_in[(size_t)s * P.in_stride + tg + lane]);
                        out_row[tg + lane] = acc;
                    }
                }
            }
            if (HG_PROF_ON(P)) pf_out += clock64() - pc0;
            if (++yb == NB) { yb = 0; ypar ^= 1u; }
        }
#pragma unroll
        for (int i = 0; i < GPS; i++) {
            const int w = wa + i;
            const int n = (w_first + w) * 32 + lane;
            if (w < wb && n < N) P.state[3 * SN + (size_t)s * N + n] = g[i];
        }
        if (HG_PROF_ON(P) && lane == 0 && sidx == (NS > 1 ? 1 : 0)) {       // a middle S warp (inbound from shared memory)
            long long* q = P.prof + ((size_t)s * gridDim.x + cta) * 8;
            q[0] = pf_in; q[1] = pf_y; q[2] = pf_mix; q[3] = pf_out;
        }
    }
}

// pick RU warps per CTA (and 1 or 2 CTAs per SM, as shared memory allows) so that the CTAs fill whole
// waves: each CTA runs the whole block, so a nearly empty last wave costs a full pass
template <typename T>
void fb_ru_config(const hg_fb* fb, bool bands, int* W_out, int* NS_out, int* cps_out) {
    const int groups = (fb->N + 31) / 32;
    const bool vf = fb->arith != HG_ARITH_FMA && !bands && !(fb->cur_env && fb->env_mode != 2);     // TermTiles of the kernel that will run
    const int w_max = (fb->cur_env && fb->env_mode != 2) ? ru::W_MAX_ENV : ru::W_MAX;
    const size_t smem_sm = (size_t)227 * 1024;
    int bestW = 4, bestNS = vf ? ru::s_warps_vf(4, 1) : ru::s_warps(4);
    double best = -1;
    for (int per_sm = 1; per_sm <= 2; per_sm++) {
        for (int W = 4; W <= w_max; W++) {
            const int NS = vf ? ru::s_warps_vf(W, per_sm) : ru::s_warps(W);
            if (NS < (vf ? ru::NB_VF : 1)) continue;         // term tiles: one hand-over barrier pair per S warp, at least one per tile slot
            if ((ru::Layout<T>(W, NS, fb->env_mode == 2 && fb->cur_env, vf).bytes() + 1024) * per_sm > smem_sm) continue;
            // registers: 80 per thread (launch bounds), W RU warps + their S warps + the gateway
            if ((long)(W + NS + 1) * 32 * 80 * per_sm > 65536) continue;
            const long slots = (long)fb->sm_count * per_sm;
            const long ctas = (long)((groups + W - 1) / W) * fb->S;
            const long waves = (ctas + slots - 1) / slots;
            // useful RU warp-passes over the warp slots of the waves that run, weighted by how many
            // RU warps an SM keeps busy (latency hiding needs >= ~3 per scheduler)
            const double fill = (double)groups * fb->S / ((double)waves * slots * W);
            const double busy = (double)(W * per_sm) / 16.0;
            const double eff = fill * (busy > 1.0 ? 1.0 : busy);
            if (eff > best + 1e-9) { best = eff; bestW = W; bestNS = NS; }
        }
    }
    *W_out = bestW;
    *NS_out = bestNS;
    *cps_out = (groups + bestW - 1) / bestW;
}
template void fb_ru_config<float>(const hg_fb*, bool, int*, int*, int*);
template void fb_ru_config<double>(const hg_fb*, bool, int*, int*, int*);

template <typename T, bool FMA, bool BANDS, int ENV>
static hg_status launch_ru(const FbRuParams<T>& Q, cudaStream_t st) {
    constexpr bool VF = ru::TermTiles<FMA, BANDS, ENV>::value;
    const size_t smem = ru::Layout<T>(Q.W, Q.NS, ENV == 2, VF).bytes();
    const int threads = (Q.W + Q.NS + 1) * 32;
    auto kern = fb_exact_ru_kernel<T, FMA, BANDS, ENV>;
    HG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (Q.k.chain.ordered && (Q.k.out || Q.k.chain.outbox)) {
        // every CTA of a stream's chain must be resident at once: one that has not started yet would stall its
        // predecessor for good once the link's ring is full
        int per_sm = 0, dev = 0, sms = 0;
        HG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        HG_CUDA(cudaGetDevice(&dev));
        HG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if ((long)Q.cps > (long)per_sm * sms)
            return fail(HG_ERR_UNSUPPORTED, "ordered mixdown: %d filters of one stream need %d CTAs, this GPU holds %d at once — split the bank over GPUs (hg_fb_process_sharded) or select HG_SUM_TREE",
                        Q.k.N, Q.cps, per_sm * sms);
    }
    kern<<<dim3(Q.cps, Q.k.S), threads, smem, st>>>(Q);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

template <typename T>
hg_status fb_launch_exact_ru(bool fma, bool bands, const FbK1Params<T>& P, int W, int NS, int cps, cudaStream_t st, const FbEnvArgs<T>* env) {
    FbRuParams<T> Q;
    Q.k = P; Q.W = W; Q.NS = NS; Q.cps = cps;
    Q.env = nullptr; Q.rms_sum = nullptr; Q.rms_ring = nullptr; Q.width = 1; Q.hop = 1; Q.env_rows = 0; Q.t_base = 0; Q.tick0 = 0;
    Q.in_lag = nullptr; Q.lag_stride = 0; Q.lag_state = nullptr; Q.lag_par = nullptr; Q.lag_taps = 0;
    if (env) {
        Q.env = env->env; Q.rms_sum = env->rms_sum; Q.rms_ring = env->rms_ring; Q.width = env->width; Q.hop = env->hop;
        Q.env_rows = env->env_rows; Q.t_base = env->t_base; Q.tick0 = env->tick0;
        Q.in_lag = env->in_lag; Q.lag_stride = env->lag_stride; Q.lag_state = env->lag_state;
        Q.lag_par = env->lag_par; Q.lag_taps = env->lag_taps;
        if (bands) return fail(HG_ERR_UNSUPPORTED, "envelope followers and per-band outputs cannot be combined");
        if (env->in_lag) return fma ? launch_ru<T, true, false, 2>(Q, st) : launch_ru<T, false, false, 2>(Q, st);
        return fma ? launch_ru<T, true, false, 1>(Q, st) : launch_ru<T, false, false, 1>(Q, st);
    }
    if (bands) return fma ? launch_ru<T, true, true, 0>(Q, st) : launch_ru<T, false, true, 0>(Q, st);
    return fma ? launch_ru<T, true, false, 0>(Q, st) : launch_ru<T, false, false, 0>(Q, st);
}
template hg_status fb_launch_exact_ru<float>(bool, bool, const FbK1Params<float>&, int, int, int, cudaStream_t, const FbEnvArgs<float>*);
template hg_status fb_launch_exact_ru<double>(bool, bool, const FbK1Params<double>&, int, int, int, cudaStream_t, const FbEnvArgs<double>*);

}  // namespace hg

extern "C" hg_status hg_fb_debug_ru_config(int dtype, int arith, int n_filters, int n_streams, int sm_count, int bands, int followers,
                                           long long out[5]) {
    if (!out || n_filters < 1 || n_streams < 1 || sm_count < 1 || (dtype != HG_F32 && dtype != HG_F64))
        return hg::fail(HG_ERR_INVALID, "hg_fb_debug_ru_config: bad argument");
    hg_fb fb;
    fb.dtype = dtype; fb.arith = arith; fb.order = 2; fb.N = n_filters; fb.S = n_streams; fb.flags = 0; fb.device = 0;
    fb.sm_count = sm_count;
    int dummy = 0;
    fb.cur_env = followers ? &dummy : nullptr;
    fb.env_mode = followers == 1 ? 1 : 2;
    const bool vf = arith != HG_ARITH_FMA && !bands && followers != 1;
    int W = 0, NS = 0, cps = 0;
    size_t smem = 0;
    if (dtype == HG_F32) { hg::fb_ru_config<float>(&fb, bands != 0, &W, &NS, &cps); smem = hg::ru::Layout<float>(W, NS, followers == 2, vf).bytes(); }
    else { hg::fb_ru_config<double>(&fb, bands != 0, &W, &NS, &cps); smem = hg::ru::Layout<double>(W, NS, followers == 2, vf).bytes(); }
    out[0] = W; out[1] = NS; out[2] = cps; out[3] = (long long)(W + NS + 1) * 32; out[4] = (long long)smem;
    return HG_OK;
}
