// chain.cuh — the order-faithful mixdown: a running bank-sum row handed from CTA to CTA (and from GPU to GPU).
//
// The reference returns  out[t] = (...((0 + v_0) + v_1) + ...) + v_{N-1},  v_n = fl(y_n[t] * g_n[t])  summed in
// filter-index order in T (src/filterbank.h:130, FF :270; with a distortion :137; Oscbank::mixdown
// src/oscbank.h:81-90).  Floating-point addition is not associative, and at BASELINE sizes the difference between
// summation orders exceeds the parity tolerance, so the kernels replay that order: the CTA that owns filters
// [a, b) receives the running row after filter a-1, adds its filters in index order and passes the row on.
// Time is the parallel axis: CTA c works on tile k while CTA c+1 works on tile k-1 (a systolic pipeline).
//
// Link = a ring of RING samples in global memory (L2-resident) using the "LL" protocol: every 8-byte word
// carries 32 bits of payload and a 32-bit tag = lap number of its ring position + 1.  An aligned 8-byte store
// is single-copy atomic, so a reader that sees the expected tag has the payload of that lap — no flag, no
// fence, one L2 (or NVLink) hop per hand-over.  Positions are virtual: vbase (a multiple of RING that grows
// monotonically over the handle's life) + the launch-relative sample index, so tags never repeat and a link may
// connect kernels with different tile sizes, or two GPUs (the ring then lives in the reader's memory and the
// writer stores through a peer mapping; the credit word lives in the writer's memory).  Back-pressure: the
// reader publishes the virtual position it has consumed up to; the writer stays at most RING ahead of it.
#pragma once
#include <cuda_runtime.h>
#include <cstring>

namespace hg {
namespace chain {

struct Params {
    unsigned long long* ring;           // local links: link (s, c), c = 1..G-1, at ring + ((s*G + c) << ring_log2) * WPS
    unsigned long long* credit;         // [S*G] consumed virtual position of link (s, c), written by its reader
    unsigned long long* credit_w;       // [S*G][CW] the same per reading warp, for links read by several warps (tile k by warp k mod n)
    unsigned long long* inbox;          // [S] rings the previous device writes (this device's memory); null: chain starts here
    unsigned long long* inbox_credit;   // [S] words in the previous device's memory where the head publishes its position
    unsigned long long* outbox;         // [S] rings in the next device's memory; null: the tail writes `out`
    unsigned long long* outbox_credit;  // [S] local words the next device's head publishes to
    unsigned long long vbase;           // virtual position of this launch's first sample (multiple of the ring size)
    unsigned* ticket;                   // CTA arrival counter: chain position = arrival order, so a CTA's predecessor
    unsigned ticket_base;               //   is always resident (no deadlock whatever order the CTAs are scheduled in)
    int* error;                         // set to 1 if a poll gives up (watchdog); the host reports it
    int ring_log2;
    int ordered;                        // 1: chained, index-order sum; 0: every CTA writes its own partial row (tree mode)
};

constexpr int CW = 16;                          // credit words per link in credit_w
constexpr unsigned SPIN_LIMIT = 1u << 27;      // polls before the watchdog gives up (tens of seconds)

// CTA's linear chain ticket (call from every thread; one atomic per CTA)
__device__ __forceinline__ unsigned take_ticket(const Params& C) {
    __shared__ unsigned ticket_s;
    if (threadIdx.x == 0) ticket_s = atomicAdd(C.ticket, 1u) - C.ticket_base;
    __syncthreads();
    return ticket_s;
}

template <typename T> struct Words { static constexpr int WPS = sizeof(T) / 4; };

__device__ __forceinline__ void st_pair2(unsigned long long* p, unsigned a, unsigned b, unsigned tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(tag), "r"(b), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_pair1(unsigned long long* p, unsigned a, unsigned tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(tag) : "memory");
}
__device__ __forceinline__ unsigned long long ld_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One end of a link, held by one warp.  K = consecutive samples per lane (the lane's packet).  The launch-wide
// constants (vbase, ring size, watchdog flag) are read from the kernel parameter block `C` at every use, so a
// port costs its holder only the ring and credit pointers.
template <typename T, int K> struct Port {
    static constexpr int WPS = Words<T>::WPS, NW = K * WPS;
    static_assert(NW == 1 || NW % 2 == 0, "packet of 1 word or pairs of words");
    unsigned long long* base = nullptr;      // ring of this link
    unsigned long long* cred = nullptr;      // its credit word
    unsigned long long seen = 0;             // writer: reader position last read from `cred`

    __device__ __forceinline__ void open(unsigned long long* ring, unsigned long long* credit, const Params& C) {
        base = ring; cred = credit; seen = C.vbase;
    }
    __device__ __forceinline__ bool on() const { return base != nullptr; }
    __device__ __forceinline__ unsigned long long* slot(const Params& C, long t) const {
        return base + (size_t)((unsigned)t & ((1u << C.ring_log2) - 1u)) * WPS;
    }
    static __device__ __forceinline__ unsigned tag(const Params& C, long t) {
        return (unsigned)((C.vbase + (unsigned long long)t) >> C.ring_log2) + 1u;
    }

    // writer: the K samples starting at launch-relative sample t (a multiple of K)
    __device__ __forceinline__ void send(const Params& C, long t, const T (&x)[K]) {
        unsigned w[NW];
        memcpy(w, x, sizeof(x));
        unsigned long long* p = slot(C, t);
        const unsigned tg = tag(C, t);
        if constexpr (NW == 1) st_pair1(p, w[0], tg);
        else {
#pragma unroll
            for (int i = 0; i < NW / 2; i++) st_pair2(p + 2 * i, w[2 * i], w[2 * i + 1], tg);
        }
    }
    // writer: wait until the reader has consumed everything that would be overwritten by samples < t_end
    __device__ __forceinline__ void reserve(const Params& C, long t_end) {
        const unsigned long long need = C.vbase + (unsigned long long)t_end;
        unsigned spins = 0;
        while (need - seen > (1ull << C.ring_log2)) {
            unsigned long long c = ld_u64(cred);
            if ((long long)(c - C.vbase) < 0) c = C.vbase;      // a position published by an earlier launch
            seen = c;
            if (++spins > SPIN_LIMIT) { *C.error = 1; break; }
        }
    }
    // reader.  A poll is split in two so that the load can be issued a tile ahead and tested when it is needed:
    // issue() only loads (payload and tags stay in registers), ready() compares the tags.
    struct Raw { unsigned w[NW]; unsigned tg[NW]; };
    __device__ __forceinline__ void issue(const Params& C, long t, Raw& r) const {
        const unsigned long long* p = slot(C, t);
        if constexpr (NW == 1)
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.w[0]), "=r"(r.tg[0]) : "l"(p) : "memory");
        else {
#pragma unroll
            for (int i = 0; i < NW / 2; i++)
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r.w[2 * i]), "=r"(r.tg[2 * i]), "=r"(r.w[2 * i + 1]), "=r"(r.tg[2 * i + 1]) : "l"(p + 2 * i) : "memory");
        }
    }
    static __device__ __forceinline__ bool ready(const Params& C, long t, const Raw& r) {
        const unsigned tg = tag(C, t);
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NW; i++) ok = ok && (r.tg[i] == tg);
        return ok;
    }
    // reader: the lane's packet at sample t; `r` may hold an earlier issue(t, r).  Polls until the tags match.
    __device__ __forceinline__ void recv(const Params& C, long t, T (&x)[K], Raw& r, bool issued) {
        if (!issued) issue(C, t, r);
        unsigned spins = 0;
        while (!ready(C, t, r)) {
            issue(C, t, r);
            if (++spins > SPIN_LIMIT) { *C.error = 1; break; }
        }
        memcpy(x, r.w, sizeof(x));
    }
    // reader: everything before launch-relative sample t_end has been consumed (call from one lane, after a
    // __syncwarp that follows every lane's recv)
    __device__ __forceinline__ void release(const Params& C, long t_end) { st_u64(cred, C.vbase + (unsigned long long)t_end); }
};

}  // namespace chain
}  // namespace hg
