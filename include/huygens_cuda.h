/* huygens_cuda.h — C ABI of libhuygens_cuda.so (hand-written sm_100a CUDA).
 *
 * The reference (amcerbu/huygens) has no FFI: its boundary is the public surface of the
 * header-only class templates in namespace soundmath.  Each entry point below names the
 * reference member it replaces (paths relative to the reference root).  The C++ drop-in
 * templates in include/huygens/ forward to these; see INTEGRATION.md for the binding.
 *
 * Conventions
 *   - dtype: HG_F32 / HG_F64.  Every `const void*` / `void*` array is of the handle's dtype.
 *   - Array pointers given to *_process may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes).  Device pointers: fully asynchronous on `cuda_stream`
 *     (a cudaStream_t; NULL = the legacy default stream).  Host
 *     pointers: staged through internal device buffers and synchronised before returning.
 *   - Setters write a host shadow; it is uploaded at the next *_process call, so parameter
 *     changes land on block boundaries (the reference's setters race with the audio thread,
 *     tests/filterbank.cpp:217-252; per-sample smoothing continues on the device).
 *   - State persists in the handle between calls: rendering L samples in one call is
 *     bit-identical to any split into consecutive calls (exact mode).
 *   - Every function returns hg_status; hg_last_error() gives a thread-local message.
 *     Nothing throws across the ABI.  There is no CPU fallback: without a CUDA device
 *     hg_*_create fails with HG_ERR_CUDA.
 */
#ifndef HUYGENS_CUDA_H
#define HUYGENS_CUDA_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int hg_status;
enum { HG_OK = 0, HG_ERR_INVALID = 1, HG_ERR_UNSUPPORTED = 2, HG_ERR_CUDA = 3, HG_ERR_NOMEM = 4, HG_ERR_NCCL = 5 };
enum { HG_F32 = 0, HG_F64 = 1 };
/* arithmetic mode: STRICT = every mul/add rounded separately (x86-64 g++ -O3 build of the
 * reference, tests/build.sh:8); FMA = the fixed contraction pattern of DESIGN.md §3. */
enum { HG_ARITH_STRICT = 0, HG_ARITH_FMA = 1 };
/* time strategy: EXACT = serial-in-time lanes, bit-identical per filter to the reference loop;
 * SCAN = chunked time-parallel scan of the 2x2 state recurrence (not bit-identical; see
 * DESIGN.md §5 for the measured error); AUTO = EXACT unless few lanes and a long block. */
enum { HG_MODE_AUTO = 0, HG_MODE_EXACT = 1, HG_MODE_SCAN = 2 };
/* hg_fb_create flags */
enum { HG_FB_FIXED = 1 /* FFilterbank<T,N,order> smoothing semantics, src/filterbank.h:203,291 */ };
/* per-band nonlinearity applied to y_n*g_n before the sum (src/filterbank.h:133-139) */
enum { HG_DIST_NONE = 0, HG_DIST_SOFTCLIP = 1 /* tests/filterbank.cpp:158-171, width 0.125 */ };

typedef struct hg_fb hg_fb;
typedef struct hg_osc hg_osc;
typedef struct hg_casc hg_casc;
typedef struct hg_rms hg_rms;
typedef struct hg_comm hg_comm;

const char* hg_last_error(void);
/* number of visible CUDA devices (0 if none / no driver); never fails */
int hg_device_count(void);
/* Measured FP32 (dtype HG_F32) or FP64 FMA issue rate of `device`, in lane-FMA slots per second,
 * from a register-resident FFMA/DFMA microbenchmark; also returns SM count and the SM clock (MHz)
 * observed during it.  This is the denominator of the FMA roofline. */
hg_status hg_fma_peak(int device, int dtype, double* slots_per_s, int* sm_count, double* sm_mhz);
/* Measured cycles per sample of the recurrence's dependency chain y <- fl(u - fl(fl(a1*y) + t2)) for one
 * warp alone on its scheduler: the floor of any bit-exact serial evaluation of one filter lane. */
hg_status hg_chain_latency(int device, int dtype, double* cycles_per_sample);

/* ---- Filterbank<T> / FFilterbank<T,N,order>  (src/filterbank.h) -------------------------- */
/* Filterbank(int order, int N, double k_p, double k_g)  src/filterbank.h:36-70; n_streams
 * independent banks sharing coefficients and gain targets (tests/filterbanks.cpp:185,195-208),
 * each with its own input and state.  order <= 2 (every bank in the reference is order 2). */
hg_status hg_fb_create(hg_fb** out, int dtype, int arith, int order, int n_filters, int n_streams,
                       double k_p, double k_g, int flags, int device);
hg_status hg_fb_destroy(hg_fb* fb);
/* coefficients(n, forward, back)  src/filterbank.h:73-82 for filters n0..n0+count-1.
 * fwd: [count][n_fwd], back: [count][n_back] (back = a1,a2 with no leading a0).  Only the first
 * min(order+1, n_fwd) / min(order, n_back) entries are used, the rest keep their value. */
hg_status hg_fb_set_coefficients(hg_fb* fb, int n0, int count, const void* fwd, int n_fwd,
                                 const void* back, int n_back);
/* boost(n, v) / boost(vector)  src/filterbank.h:85-96: targets of the smoothed pre-gain */
hg_status hg_fb_set_boost(hg_fb* fb, int n0, int count, const void* values);
/* mix(n, v) / mix(vector)  src/filterbank.h:99-110: targets of the smoothed mixdown gain */
hg_status hg_fb_set_mix(hg_fb* fb, int n0, int count, const void* values);
/* open()  src/filterbank.h:112-116 */
hg_status hg_fb_open(hg_fb* fb);
/* Block form of the caller loop `out[t] = bank(in[t]); bank.tick();` (operator() :125-131,
 * compute :170-187, tick :142-148; tests/resynthesis.cpp:33-42).
 *   in        [n_streams][L]
 *   out_sum   [n_streams][L]              or NULL
 *   out_bands [n_streams][n_filters][L]   or NULL  (y_n[t], the private `outputs` row)
 * distortion != NONE selects operator()(sample, distortion) :133-139. */
hg_status hg_fb_process(hg_fb* fb, const void* in, void* out_sum, void* out_bands, size_t L,
                        int mode, int distortion, void* cuda_stream);
/* new: the same block render with one RMSbank<T,.>(width) envelope follower per (stream, filter) fused
 * behind the filters (src/rmsbank.h:54-74 applied to the real band output y_n: q = y^2,
 * S <- (q - q[t-width]) + S, env = sqrt(S/width)); BASELINE configs[3].  out_env (nullable)
 * [n_streams][L/hop][n_filters] receives env at block samples hop-1, 2*hop-1, ...; the follower state
 * (running sums and the ring of the last `width` squares) persists in the handle; width >= 64 and
 * fixed at the first call.  Bit-identical to Filterbank bands -> RMSbank. */
/* new: a block with parameter changes at exact ticks — the reference's setter calls between two ticks of the
 * caller loop (audio-rate modulation, src/subtractive.h:215-230: coefficients() + mix() every tick).  Event e
 * takes effect before tick events[e].tick of this block (ticks ascending, <= L); the block is rendered
 * in segments between events, so results equal per-tick setter calls bit for bit.  n_streams must be 1. */
enum { HG_EV_COEFFICIENTS = 0, HG_EV_BOOST = 1, HG_EV_MIX = 2, HG_EV_OPEN = 3 };
typedef struct {
    size_t tick;
    int kind;            /* HG_EV_* */
    int n0, count;       /* filters [n0, n0 + count) */
    const void* a;       /* COEFFICIENTS: forward [count][na];  BOOST / MIX: values [count] */
    int na;
    const void* b;       /* COEFFICIENTS: back [count][nb] */
    int nb;
} hg_fb_event;
hg_status hg_fb_process_events(hg_fb* fb, const void* in, void* out_sum, size_t L, const hg_fb_event* events,
                               size_t n_events, int mode, void* cuda_stream);
/* new: a block whose parameters change every `hop` ticks (hop = 1: audio-rate modulation).  Row r of `rows`
 * [n_rows][7][n_filters] = b0 b1 b2 a1 a2 boost mix replaces the bank's parameters before block tick r*hop —
 * coefficients(n, {b0,b1,b2}, {a1,a2}), boost(values), mix(values) for every filter between two ticks of the
 * caller loop (filterbank.h:73-110; src/subtractive.h:215-230).  One kernel, lanes reload their row at the hop
 * boundaries; boost and mix keep their per-sample smoothing.  n_rows * hop >= L; the bank keeps the last row. */
hg_status hg_fb_process_tv(hg_fb* fb, const void* in, void* out_sum, size_t L, const void* rows, size_t n_rows, int hop,
                           void* cuda_stream);
/* How the followers obtain y[t-width].  HG_ENV_RING: a ring of the last `width` squares per filter
 * in device memory (8 bytes of HBM traffic per filter-sample, width*sizeof(T) bytes per filter).
 * HG_ENV_REPLICA (default): a second copy of every filter runs `width` ticks behind on the same input and regenerates
 * y[t-width] bit for bit (no ring, twice the arithmetic, ~2x faster on B200).  Parameter changes reach the
 * replicas `width` ticks after the filters (the library keeps the pending sets and splits blocks at the
 * hand-over ticks); the one restriction is that every tick must then be rendered through hg_fb_process_env
 * (plain hg_fb_process is refused while replica followers run).  Call before the first hg_fb_process_env. */
enum { HG_ENV_RING = 1, HG_ENV_REPLICA = 2 };
hg_status hg_fb_set_follower_mode(hg_fb* fb, int mode);
/* Drop the followers (running sums, ring or replicas): the next hg_fb_process_env starts them afresh, like a newly
 * constructed RMSbank.  While followers run, plain hg_fb_process / hg_fb_set_state are refused in BOTH modes — the
 * followers must see every tick, and their state is not part of hg_fb_get_state. */
hg_status hg_fb_reset_followers(hg_fb* fb);
hg_status hg_fb_process_env(hg_fb* fb, const void* in, void* out_sum, void* out_env, size_t L, int width,
                            int hop, void* cuda_stream);
/* state access (streaming / checkpoint tests): y1,y2,p,g are [n_streams][n_filters]; xhist is
 * [n_streams][2] = x[t-1], x[t-2].  Host pointers. NULL entries are skipped. */
hg_status hg_fb_get_state(hg_fb* fb, void* y1, void* y2, void* p, void* g, void* xhist);
hg_status hg_fb_set_state(hg_fb* fb, const void* y1, const void* y2, const void* p, const void* g,
                          const void* xhist);
/* what the last hg_fb_process did: mode used, kernel launches, filters per thread, CTAs */
hg_status hg_fb_last_launch(hg_fb* fb, int* mode_used, int* launches, int* filters_per_thread,
                            int* grid_ctas);
/* How operator()'s N-way mixdown is evaluated (src/filterbank.h:130, :137).
 * HG_SUM_ORDERED (default): sum_n fl(y_n*g_n) accumulated in filter-index order in T, exactly the reference loop
 *   (fused multiply-add in HG_ARITH_FMA): the running row travels from CTA to CTA through L2-resident rings.
 *   out_sum is bit-identical to the reference/oracle.
 * HG_SUM_TREE: every CTA sums its own filters and the partial rows are added in a fixed tree order — same
 *   per-filter values, different summation order (float: ~3e-5 of peak at 4096 filters, 3e-4 at 65 536). */
enum { HG_SUM_TREE = 0, HG_SUM_ORDERED = 1 };
hg_status hg_fb_set_sum_mode(hg_fb* fb, int mode);
/* new: post-processing of the finished bank sum, fused into the store of out_sum[t] (kernel epilogue) — the last
 * step of the reference's caller loops:
 *   HG_EPI_SOFTCLIP  out = softclip(dry * in[t] + sum[t], width)   tests/filterbank.cpp:205-207 (outer softclip(., 0.9))
 *   HG_EPI_LIMITER   out = 2/PI * atan(dry * in[t] + sum[t])       src/wave.h:150 `limiter`, tests/harmbank.cpp:76
 * evaluated in double and rounded to T (device atan: within tolerance of libm, not bit-identical).  HG_EPI_NONE
 * restores the plain sum.  In a sharded render the rank that holds the chain's tail applies it. */
enum { HG_EPI_NONE = 0, HG_EPI_SOFTCLIP = 1, HG_EPI_LIMITER = 2 };
hg_status hg_fb_set_epilogue(hg_fb* fb, int kind, double dry, double width);
/* Block form of the tests/filterbank.cpp:200-215 process body: out[t] = epilogue(dry * dry_in[t] + F(in[t], distortion))
 * with the epilogue set by hg_fb_set_epilogue — `in` is the (pre-clipped, driven) signal the bank filters, `dry_in`
 * the raw input of the dry path (NULL: the same block).  Exact mode. */
hg_status hg_fb_process_voice(hg_fb* fb, const void* in, const void* dry_in, void* out_sum, size_t L, int distortion,
                              void* cuda_stream);
/* Synchronous health check (copies one flag from the device): HG_ERR_CUDA if the ordered mixdown's watchdog fired
 * during an earlier asynchronous block.  Host-pointer blocks check this before returning. */
hg_status hg_fb_check(hg_fb* fb);
/* Debug aid: with HG_FB_PROF=1 in the environment at hg_fb_create, the warp-specialised kernel records per-CTA cycle
 * counters of its S and R roles (waits for the y tile, for the running row, mixdown, hand-over; R's barrier waits and
 * total).  Copies [n_ctas][8] counters of the last launch (chain order).  Not part of the reference's surface. */
hg_status hg_fb_debug_profile(hg_fb* fb, long long* counters, int n_ctas);
/* Debug aid, host only (no CUDA call): the launch geometry the throughput-regime kernel (K1-ru) would use for a bank of
 * n_filters x n_streams on a GPU with sm_count SMs.  followers: 0 none, 1 HG_ENV_RING, 2 HG_ENV_REPLICA.
 * out[5] = RU warps per CTA, S warps per CTA, CTAs per stream, threads per CTA, dynamic shared memory per CTA (bytes).
 * The ordered mixdown needs out[2] <= sm_count x CTAs per SM: tests use this to check the one-wave rule off the GPU. */
hg_status hg_fb_debug_ru_config(int dtype, int arith, int n_filters, int n_streams, int sm_count, int bands, int followers,
                                long long out[5]);
/* Debug aid, host only (no CUDA call): the launch geometry of the warp-specialised kernel (K1-ws) for a bank of
 * n_filters x n_streams on a GPU with sm_count SMs.  filters_per_lane: 0 or 1 = the default CTAs of 32 filters, 2 = the
 * opt-in CTAs of 64 filters (float; what HG_FB_WS_NF=2 selects).  out[8] = fits (0: the dispatcher picks K1-ru, the rest
 * is 0), residency level (CTAs per SM = 1 << level), filters per lane, CTAs per stream, CTAs in total, tile length
 * (samples), y-tile slots, dynamic shared memory per CTA (bytes).  The ordered mixdown needs all CTAs resident at once:
 * out[4] <= sm_count << level and (out[7] + 1024) << level <= 228 KB. */
hg_status hg_fb_debug_ws_config(int dtype, int n_filters, int n_streams, int sm_count, int filters_per_lane, long long out[8]);

/* ---- filter-partitioned rendering over the GPUs of one box (BASELINE configs[2]) ------------------------------ */
/* new (nothing in the reference to match: it is single-threaded host code).  A bank of N filters is split over G
 * ranks, rank r holding the filters [r N/G, (r+1) N/G) in an hg_fb of its own device; every rank renders the same
 * input.  One rank per process: rank 0 calls hg_comm_get_unique_id (128 bytes, an ncclUniqueId), hands it to the
 * others by any means, then every rank calls hg_comm_init_rank.  One process driving several GPUs (a C++ audio
 * host): hg_comm_init_all fills comms[0..n_devices).  NCCL (libnccl.so.2, loaded at run time) bootstraps the group;
 * neighbouring ranks map a small mailbox of each other's device memory (CUDA IPC / peer access over NVLink). */
hg_status hg_comm_get_unique_id(void* id128);
hg_status hg_comm_init_rank(hg_comm** out, int n_ranks, int rank, const void* id128, int device);
hg_status hg_comm_init_all(hg_comm** comms, int n_devices, const int* devices /* NULL: 0..n-1 */);
hg_status hg_comm_destroy(hg_comm* comm);
/* ncclGroupStart / ncclGroupEnd: when ONE host thread issues the collective call of several ranks
 * (hg_comm_init_all), bracket the ranks' hg_fb_process_sharded calls with these (required for HG_SHARD_REDUCE and
 * for a `root` other than the chain's tail; harmless otherwise).  Use device buffers there: a host-buffer call
 * blocks until its own kernel has finished, which needs the other ranks' kernels to be running. */
hg_status hg_comm_group_start(void);
hg_status hg_comm_group_end(void);
int hg_comm_rank(const hg_comm* comm);
int hg_comm_size(const hg_comm* comm);
/* HG_SHARD_CHAIN: operator()'s index-order sum (src/filterbank.h:130) is kept across the devices — the running
 *   bank-sum row leaves the last CTA of rank r for rank r+1 through peer memory, tile by tile under the compute, one
 *   kernel per GPU; the result is bit-identical to one hg_fb of all N filters and lands on the last rank.
 * HG_SHARD_REDUCE: every rank renders its own partial sum and one ncclReduce adds the rows (order-free: within
 *   ~N eps of the reference sum). */
enum { HG_SHARD_CHAIN = 0, HG_SHARD_REDUCE = 1 };
/* Collective: every rank of `comm` calls it with its own handle, the same input block and L.  out_sum [L] is written
 * on rank `root` only (root < 0: the last rank for CHAIN, rank 0 for REDUCE; a root other than the chain's tail costs
 * one ncclSend/ncclRecv of the row); other ranks may pass NULL.  One stream per handle (n_streams == 1). */
hg_status hg_fb_process_sharded(hg_fb* fb, hg_comm* comm, const void* in, void* out_sum, size_t L, int shard_mode,
                                int root, void* cuda_stream);

/* ---- Oscbank<T,N>  (src/oscbank.h, src/multichannel.h) ------------------------------------ */
hg_status hg_osc_create(hg_osc** out, int dtype, int arith, int n_partials, int device);   /* oscbank.h:37-47 */
hg_status hg_osc_destroy(hg_osc* osc);
/* freqmod(index, target_hz)  oscbank.h:49-56; out-of-range index is silently ignored (:51) */
hg_status hg_osc_freqmod(hg_osc* osc, int index, double target_hz);
hg_status hg_osc_set_frequencies(hg_osc* osc, int n0, int count, const double* target_hz);
/* Multichannel activate/deactivate/open/close  multichannel.h:87-118 */
hg_status hg_osc_activate(hg_osc* osc, const int* indices, int count, int on);
hg_status hg_osc_open(hg_osc* osc);
hg_status hg_osc_close(hg_osc* osc);
int hg_osc_active_count(hg_osc* osc);
/* L ticks.  Per tick t, BEFORE advancing: out_mix[t] = mixdown() (complex interleaved, sum over
 * the active set, oscbank.h:81-90); out_amix[t] = Mixer(Modbank(amps, osc())) = sum over ALL n of
 * amps[n]*Re(phase_n) (modbank.h:41-45, mixer.h:32-35); then tick() (oscbank.h:59-63).
 * Any of out_mix / out_amix may be NULL; amps is [n_partials] and required iff out_amix. */
hg_status hg_osc_process(hg_osc* osc, void* out_mix, const void* amps, void* out_amix, size_t L,
                         void* cuda_stream);
/* The same, with time-varying amplitudes: row r of amp_rows [n_rows][n_partials] drives ticks
 * [r*hop, (r+1)*hop) of this block — the resynthesis half of tests/resynthesis.cpp-style chains, where the
 * amplitudes are the envelope rows of hg_fb_process_env / hg_rms_process (modbank.h:41-45 with a new
 * amplitude vector every hop ticks).  n_rows * hop >= L; out_amix is required, out_mix optional. */
hg_status hg_osc_process_env(hg_osc* osc, void* out_mix, const void* amp_rows, size_t n_rows, int hop,
                             void* out_amix, size_t L, void* cuda_stream);
hg_status hg_osc_get_phases(hg_osc* osc, void* phases /* complex interleaved [n_partials], host */);
hg_status hg_osc_set_phases(hg_osc* osc, const void* phases);

/* ---- RMSbank<T,N>  (src/rmsbank.h) --------------------------------------------------------- */
hg_status hg_rms_create(hg_rms** out, int dtype, int arith, int n_channels, int width, int device); /* rmsbank.h:32-45 */
hg_status hg_rms_destroy(hg_rms* rms);
/* L ticks of `out = rms(z); rms.tick();` (operator() :54-66 evaluated once per tick, tick :68-74).
 * zre, zim, out: [n_channels][L] channel-major; zim may be NULL (real input). */
hg_status hg_rms_process(hg_rms* rms, const void* zre, const void* zim, void* out, size_t L,
                         void* cuda_stream);

/* ---- Latchbank<T,N>  (src/latchbank.h) ------------------------------------------------------- */
/* Latchbank(T thresh, T ratio) src/latchbank.h:34-53, attached to this RMSbank handle (flags reset). */
hg_status hg_rms_set_latch(hg_rms* rms, double thresh, double ratio);
/* L ticks of `gated = latch(&rms, z); rms.tick();` (latchbank.h:81-100, tests/harmbank.cpp:84-99):
 * out_re/out_im [n_channels][L] = z * engaged; out_level (nullable) = the RMS level that drove the gate.
 * zim / out_im may both be NULL (real signal). */
hg_status hg_rms_process_latch(hg_rms* rms, const void* zre, const void* zim, void* out_re, void* out_im,
                               void* out_level, size_t L, void* cuda_stream);

/* ---- Slidebank<T,N> / Stickbank<T,N>  (src/slidebank.h, src/stickbank.h) ---------------------- */
/* Slidebank(order, radii) slidebank.h:53-97: per channel a cascade of max(1, order) complex one-pole
 * low-passes with radius radii[n] (complex interleaved [n_channels]), one tick of delay per stage.
 * Stickbank(order, rad) stickbank.h:46-78: every channel through the all-pole smoother with max(1, order)
 * poles from coefficients() (:198-218); rmod (:69-79) re-derives them.  order <= 16. */
hg_status hg_slide_create(hg_casc** out, int dtype, int arith, int n_channels, int order, const double* radii, int device);
hg_status hg_stick_create(hg_casc** out, int dtype, int arith, int n_channels, int order, double rad, int device);
hg_status hg_stick_rmod(hg_casc* bank, double rad);
hg_status hg_casc_destroy(hg_casc* bank);
/* L ticks of `out = bank(&z); bank.tick();` (slidebank.h:100-124,166-175; stickbank.h:82-107,181-195).
 * zre, zim, out_re, out_im: [n_channels][L] channel-major; zim may be NULL (real input, :111-117 / :91-98). */
hg_status hg_casc_process(hg_casc* bank, const void* zre, const void* zim, void* out_re, void* out_im, size_t L,
                          void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* HUYGENS_CUDA_H */
