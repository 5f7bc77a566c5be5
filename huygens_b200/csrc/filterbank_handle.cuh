// filterbank_handle.cuh — hg_fb handle layout and kernel parameter blocks.
#pragma once
#include <cstdint>
#include <deque>
#include <vector>
#include "hg_common.cuh"
#include "chain.cuh"

// Role cycle counters of the exact kernels (hg_fb_debug_profile): compiled in only with -DHG_FB_PROFILE, so the
// production kernels carry no timing code.
#ifdef HG_FB_PROFILE
#define HG_PROF_ON(P) ((P).prof != nullptr)
#else
#define HG_PROF_ON(P) false
#endif

constexpr int K1_NT = 128;   // threads per CTA of the exact kernel (one warp per SM sub-partition)

// Device layout (all of the handle's dtype T):
//   d_par   [7][N]      b0 b1 b2 a1 a2 pin gin          (shared by all streams)
//   d_state [4][S][N]   y1 y2 p g                        (per stream)
//   d_xh    [2][S][2]   x[t-1], x[t-2] per stream, ping-pong between launches
struct hg_fb {
    int dtype, arith, order, N, S, flags, device;
    int sm_count;
    double c_p, s_p, c_g, s_g;   // already rounded to T
    void* h_par = nullptr;       // pinned host shadow of d_par
    void* d_par = nullptr;
    bool dirty = true;
    int taps = 0;                // structure of the forward taps, detected at upload
    void* d_state = nullptr;
    void* d_xh = nullptr;
    int xh_cur = 0;
    int cur_dist = 0;            // distortion of the block being processed
    const void* cur_tv = nullptr;   // device rows of the block's parameter table (hg_fb_process_tv), else null
    int cur_tv_hop = 0;
    hg::DevBuf in_stage, out_stage, bands_stage, partial, scan_a, scan_b;
    int force_F = 0;      // HG_FB_FORCE_F: 1,2,4,8 = filters per thread of the lane kernel; -1 = warp-specialised kernel
    int force_chunks = 0;
    // order-faithful mixdown (chain.cuh): link rings and credits, the virtual sample position, the CTA ticket counter
    bool smoothers_custom = false;         // hg_fb_set_state gave the streams their own p / g (the scan's prepass reads stream 0)
    int sum_mode = 1;                      // HG_SUM_ORDERED (1, default) or HG_SUM_TREE (0)
    hg::Epilogue epi = {0, 0.0, 0.0};      // hg_fb_set_epilogue
    const void* cur_dry_in = nullptr;      // device pointer of the block's dry signal (hg_fb_process_voice), null: the input
    hg::DevBuf dry_stage;
    void* d_chain_ring = nullptr;          // [links][ring samples] LL words
    size_t chain_ring_words = 0;
    int chain_ring_log2 = 0;
    unsigned long long* d_chain_credit = nullptr;
    size_t chain_links = 0;
    unsigned long long chain_vbase = 0;    // multiple of the ring size, grows with every launch
    unsigned* d_ticket = nullptr;          // [1] CTA arrival counter, never reset: launches subtract `ticket_base`
    unsigned ticket_base = 0;
    int* d_chain_error = nullptr;          // [1] watchdog flag
    long long* d_prof = nullptr;           // debug (HG_FB_PROF=1 at create): [ctas][8] role cycle counters of the last launch
    // the chain's ends for the block being processed (set by the sharded entry points, else null)
    const void* cur_carry_in = nullptr;    // [S][L] running row to start from (null: zero)
    hg::chain::Params cur_peer = {};       // inbox / outbox of a multi-GPU chain (ring, credit fields unused)
    const void* chain_owner = nullptr;     // the hg_comm whose position sequence the rings last carried (null: the handle's own)
    // last launch report
    int last_mode = 0, last_launches = 0, last_F = 0, last_ctas = 0;
    int ws_level = 0;
    int ws_nf = 1;               // filters per recurrence lane of the last ws launch (a CTA owns 32 * ws_nf filters)
    int force_ws_nf = 0;         // HG_FB_WS_NF (read at create): 1 or 2 forces it, 0 = by lane count
    // fused per-band envelope followers (hg_fb_process_env)
    void* d_rms_sum = nullptr;   // [S][N]
    void* d_rms_ring = nullptr;  // [S][ceil(N/32)][width][32]
    int rms_width = 0;
    long long rms_ticks = 0;
    void* cur_env = nullptr;     // non-null while hg_fb_process_env runs (followers on)
    void* cur_env_out = nullptr; // device pointer of the block's envelope output, may be null
    int cur_hop = 0;
    size_t cur_env_rows = 0;
    int env_mode = 2;            // HG_ENV_REPLICA (2, default: FMA-bound, 2.3x faster on B200) or HG_ENV_RING (1); fixed once the followers have started
    void* d_lag_state = nullptr; // replica followers: [3][S][N] y1, y2, p of the lagged copies
    void* d_xhist = nullptr;     // replica followers: [S][width + 4] input samples before the current block
    void* d_lag_par = nullptr;   // replica followers: [7][N] parameters the replicas currently run with
    int lag_taps = 0;
    bool lag_live = false;       // d_lag_par holds a parameter set (from the first block with followers on)
    struct LagPending { long long tick; int taps; std::vector<unsigned char> par; };
    std::deque<LagPending> lag_pending;   // parameter sets on their way to the replicas: active from global tick `tick`
    hg::DevBuf xcat;             // replica followers: [S][width + 4 + L] history + block
    hg::DevBuf tv_stage;         // hg_fb_process_tv: device copy of a host parameter table
    int* d_sm_slot = nullptr;        // per-SM CTA arrival counters (role rotation of co-resident ws CTAs)
    hg::DevBuf env_stage;            // residency level chosen for the warp-specialised kernel (1 << level CTAs per SM)
};

namespace hg {

template <typename T> struct FbK1Params {
    const T* par;
    T* state;
    const T* xh;
    T* xh_out;
    const T* in;
    T* out;
    T* bands;
    long in_stride, out_stride, bands_stride;
    int N, S;
    long L;
    T c_p, s_p, c_g, s_g;
    int direct;
    int bands_vec;
    int dist;      // HG_DIST_*: per-band nonlinearity before the sum (lane kernel only)
    int taps;      // 0 general; 1: b1 == 0 everywhere; 2: b1 == 0 and b2 == -b0 everywhere
    chain::Params chain;      // how the bank sum travels from CTA to CTA (chain.cuh)
    Epilogue epi;             // post-processing of out[t] at the chain's tail (kind 0: none)
    const T* dry_in;          // [S][L] (stride in_stride): the dry signal of the epilogue (the block's input unless the caller gave another)
    long long* prof;          // debug (HG_FB_PROF): per-CTA cycle counters of the S / R roles, or null
    const T* carry_in;        // [S] rows (stride carry_stride): the running sum the chain's head starts from; null = 0
    long carry_stride;
    int* sm_slot;  // [>= SM count] zeroed before a launch with several CTAs per SM: arrival order of the CTAs on each SM
    // time-varying parameters (lane kernel only): row r = [b0 b1 b2 a1 a2 pin gin][N] replaces the bank's
    // parameters before block tick r * tv_hop (setter calls between two ticks of the caller loop)
    const T* tv_rows;
    int tv_hop;
    long tv_t0;    // block tick of this launch's first sample
};

template <typename T>
hg_status fb_process_exact(hg_fb* fb, const T* d_in, T* d_out, T* d_bands, size_t L, cudaStream_t st);
template <typename T>
hg_status fb_launch_exact_ws(bool fma, bool bands, int level, int nf, const FbK1Params<T>& P, int groups, cudaStream_t st);
template <typename T>
bool fb_ws_describe(int level, int nf, int* tile, int* u_slots, int* y_slots, long long* smem_bytes);
template <typename T> struct FbEnvArgs {
    T* env; T* rms_sum; T* rms_ring;
    int width, hop;
    long env_rows, t_base;
    long long tick0;
    const T* in_lag = nullptr;    // replica followers: lagged input of this launch (null => ring followers)
    long lag_stride = 0;
    T* lag_state = nullptr;
    const T* lag_par = nullptr;   // the replicas' parameter set [7][N] and its tap structure
    int lag_taps = 0;
};
template <typename T>
hg_status fb_launch_exact_ru(bool fma, bool bands, const FbK1Params<T>& P, int W, int NS, int ctas_per_stream, cudaStream_t st,
                             const FbEnvArgs<T>* env = nullptr);
template <typename T>
void fb_ru_config(const hg_fb* fb, bool bands, int* W_out, int* NS_out, int* ctas_per_stream_out);
template <typename T>
hg_status fb_process_scan(hg_fb* fb, const T* d_in, T* d_out, size_t L, cudaStream_t st);

}  // namespace hg
