// srx_fused.cu — the whole overlap step as ONE persistent kernel (one CTA per SM), and its frame-sharded form in which
// the cross-GPU exchange of the key accumulator runs inside the same kernel over NVLink peer memory.
//
// Replaces OverlapCorresponder.step_finished (source/common_utils/stable_render_utils/corresponder.py:298-376) for the
// common geometry (8x8 id pixels per latent cell, 4 channels, float accumulators); everything else takes the split
// kernels of srx_overlap.cu.
//
// Phases:
//   A  stream   : a producer warp keeps a ring of FZ_STAGES shared-memory stages full with bulk async copies
//                 (cp.async.bulk -> UBLKCP, completion on an mbarrier): per stage 8 id rows x 32 cells (32 KB of ids,
//                 L2 evict-first: read once) plus the 32 cells' latents.  Sixteen consumer warps key the pixels out of
//                 shared memory (a warp per cell, two adjacent pixels per lane), reduce equal keys inside the warp
//                 (REDUX for single-key cells, in-lane and row-pair merges otherwise) and issue one 128-bit vector
//                 reduction + one count reduction per distinct key into the L2-resident accumulator; the cell's last
//                 valid pixel is its winner.  While the first copies are in flight the consumers clear the accumulator
//                 of the NEXT step (double buffered by step parity).  In the cached-plan regime this phase reads the
//                 (key, cell, multiplicity) pairs stored by the bucketing passes (MARK / EMIT modes) instead of the ids.
//   X  exchange : (frame-sharded runs) after a box-wide barrier rank r PULLS the partial sums of its 1/world slice of
//                 the accumulator from every peer over NVLink, adds them in rank order and keeps the totals as 32-byte
//                 flagged records in its own memory; one signal round later everybody fetches the records it needs
//                 from their owners.  Only arrival counters are ever stored to a peer; no NCCL call.
//   B  gather   : each latent frame is handled by a group of CTAs: winner's mean, blend, sums of x, x^2, b, b^2 (double)
//   C  AdaIN    : the CTAs of a frame exchange their sums and meet at the frame's own counter, then re-standardise x
//                 in place (the latents stay in shared memory between B and C).
// Barriers are monotonic arrival counters in the workspace (targets derive from a device-side step counter, so the
// kernel replays unchanged from a CUDA graph); the launch is cooperative so that all CTAs are co-resident.
#include "srx_plan.cuh"
#include <stdlib.h>
#include <vector>

#define FZ_CONS 16                        // consumer warps
#define FZ_THREADS ((FZ_CONS + 1) * 32)   // + one producer warp
#define FZ_CELLS 32                       // cells per stage
#define FZ_CPW (FZ_CELLS / FZ_CONS)       // cells per consumer warp per stage
#define FZ_STAGES 6
#define FZ_SLOT_BITS 25
#define FZ_BU 4                           // cells per thread per round of the gather phase

enum { FZ_ST_KEY_RANGE = 0, FZ_ST_TIMEOUT = 2 };   // status words (srx_plan_check); 1 = ST_CELL_RANGE of the split kernels
#define FZ_SPIN_FAST 8192u                // polls before a waiter starts backing off
#define FZ_SPIN_LIMIT (1u << 21)          // backed-off polls (~1 us each) before a wait gives up: ~2 s
// What the kernel does with the ids:
//   STEP    the streaming overlap step (ids new for this call)
//   MARK    bucketing pass 1 — phase A only: winners, the byte map of winner keys, pairs per CTA.  No reductions.
//   EMIT    bucketing pass 2 — phase A only: every CTA appends the (key, cell, multiplicity) pairs whose key is in the
//           map to its own region of the pool (same static deal of items as MARK, so the MARK counts bound the regions)
//   CACHED  the overlap step from the pool: no id traffic; reductions only for keys that matter
enum { FZ_MODE_STEP = 0, FZ_MODE_MARK = 1, FZ_MODE_EMIT = 2, FZ_MODE_CACHED = 3 };

template <typename IdT> struct FzId;
// row pitch = row bytes + a skew that makes the consumers' 128-bit shared loads bank-conflict free
template <> struct FzId<int4> { enum { PX = 16, ROW = FZ_CELLS * 8 * 16, PITCH = ROW + 16 }; };
template <> struct FzId<short4> { enum { PX = 8, ROW = FZ_CELLS * 8 * 8, PITCH = ROW + 64 }; };

template <typename IdT> struct FzLayout {
    enum {
        LAT_OFF = 8 * FzId<IdT>::PITCH,
        DESC_OFF = LAT_OFF + 4 * FZ_CELLS * 4,
        STAGE = (DESC_OFF + 16 + 127) / 128 * 128,
        BAR_OFF = FZ_STAGES * STAGE,                 // full[FZ_STAGES], empty[FZ_STAGES]
        MISC_OFF = BAR_OFF + 2 * FZ_STAGES * 8,      // epoch word
        RED_OFF = (MISC_OFF + 16 + 15) / 16 * 16,    // [FZ_THREADS/32][16] double block-reduce scratch
        COEF_OFF = RED_OFF + (FZ_THREADS / 32) * 16 * 8,
        TOTAL = COEF_OFF + 16 * 4
    };
};

struct FzParams {
    const void *ids;
    void *x;
    const int *fmap;
    char *ws;                 // this rank's workspace
    char *peers[SRX_MAX_PEERS];
    char *mc;                 // multicast (NVLS) view of the workspace: one address reaches every rank's copy; NULL = pull exchange
    long long ll_off;         // total records of 32 B: this rank's slice [ceil(kcap / world)] (pull) or all kcap slots (NVLS)
    int ll_slice;             // slots per owner slice = ceil(kcap / world)
    long long accum_stride, pads_off, ctrl_off, stats_off;
    int *winner;
    int *status;
    int H, W, h, w, batch;
    unsigned kcap;
    int nrows, chunks;
    float ratio, one_minus;
    int adain;
    int world, rank;
    int accum_vec;            // float4 vectors in one accumulator (sums then counts)
    double inv_n, inv_nm1;    // 1 / (h*w), 1 / (h*w - 1)
    int mode;                 // FZ_MODE_*
    unsigned char *need;      // [kcap] 1 = the key wins some cell on some rank (the only keys whose mean is ever read)
    int *cta_tab;             // [3][gridDim]: pairs seen by each CTA (MARK), entries kept (EMIT), first entry of its region
    int2 *pool;               // cached plan: (slot, cell << 6 | multiplicity - 1) entries, one region per CTA
    float *cnt_plan;          // cached plan: [kcap] entries per key on this rank (filled by EMIT; counts depend on the ids only)
    int dbg;                  // SRX_FZ_DEBUG experiment bits (results are wrong when set): 1 = no reductions,
                              // 2 = consumers only drain the ring, 4 = stop after phase A
};

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------
// release/acquire fence at GPU scope (MEMBAR.ALL.GPU).  __threadfence() is the sequentially-consistent fence.sc.gpu,
// which is measurably slower and stronger than anything the counters below need.
__device__ __forceinline__ void fz_fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ int4 lds128(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds64(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
template <typename XT> __device__ __forceinline__ float lds_x(uint32_t base, int idx);
template <> __device__ __forceinline__ float lds_x<float>(uint32_t base, int idx) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + idx * 4));
    return v;
}
template <> __device__ __forceinline__ float lds_x<__half>(uint32_t base, int idx) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + idx * 2));
    return __half2float(__ushort_as_half(v));
}
template <> __device__ __forceinline__ float lds_x<__nv_bfloat16>(uint32_t base, int idx) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + idx * 2));
    return __uint_as_float((unsigned)v << 16);
}

// arrival counters: release-add / acquire-load, gpu scope on one GPU, system scope across NVLink peers
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v, bool sys) {
    if (sys) asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p, bool sys) {
    unsigned v;
    if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p, bool sys) {
    unsigned v;
    if (sys) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// phase timestamps (SM clock) of the first and the last CTA, kept in the control block: [ctrl + 64 + 64*which][8]
#define FZ_TRACE(idx)                                                                                              \
    do {                                                                                                           \
        if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {                                        \
            const long long now_ = clock64();                                                                      \
            reinterpret_cast<long long *>(P.ws + P.ctrl_off + 64 + (blockIdx.x == 0 ? 0 : 64))[idx] = now_;        \
            if (blockIdx.x == 0) reinterpret_cast<long long *>(P.ws + P.ctrl_off + 1024)[(fz_epoch_ & 31u) * 8 + idx] = now_; \
        }                                                                                                          \
    } while (0)

// 32-byte exchange record {sum.xyzw, count, flag, -, -}: written with ONE 256-bit store and read with ONE 256-bit load,
// so data and flag travel in the same sector and a reader that sees this step's flag also sees this step's data — no
// fence, no acknowledgement round trip, no separate barrier (same idea as NCCL's LL protocols).  Flags are the step
// number: records never need clearing.
struct FzRec { float4 s; float c; unsigned flag; };
__device__ __forceinline__ void fz_rec_store(char *p, float4 s, float c, unsigned flag) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)), "r"(__float_as_uint(s.z)), "r"(__float_as_uint(s.w)),
                   "r"(__float_as_uint(c)), "r"(flag), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ FzRec fz_rec_load(const char *p) {
    unsigned a, b, c, d, e, f, g, h;
    asm volatile("ld.relaxed.sys.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p) : "memory");
    FzRec r;
    r.s = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
    r.c = __uint_as_float(e);
    r.flag = f;
    return r;
}
__device__ __forceinline__ FzRec fz_rec_wait(const char *p, unsigned flag, int *status) {
    FzRec r = fz_rec_load(p);
    unsigned spins = 0;
    while (r.flag != flag) {
        __nanosleep(200);   // thousands of threads wait at once: back off instead of saturating L2 with polls
        if (++spins > FZ_SPIN_LIMIT || *reinterpret_cast<volatile int *>(status + FZ_ST_TIMEOUT)) {
            atomicOr(status + FZ_ST_TIMEOUT, 1);
            break;
        }
        r = fz_rec_load(p);
    }
    return r;
}

// NVLS exchange (multimem.*: one instruction addresses the same offset on every rank through the NVSwitch).
//   multimem.ld_reduce: the switch reads the addressed words on all ranks and returns their sum — the owner receives its
//                       slice already reduced (1/world of the bytes a pull moves over its links);
//   multimem.st:        one store lands in every rank's copy — the all-gather of the totals as a broadcast;
//   multimem.red:       one arrival increments a counter on every rank.
__device__ __forceinline__ float4 mm_ld_reduce_f4(const char *p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_f4(char *p, float a, float b, float c, float d) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mm_red_add_u32(char *p, unsigned v) {
    asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Pushed total record (NVLS mode): two 16-byte halves {sum.x, sum.y, sum.z, step} {sum.w, count, step, 0}, each written by ONE
// 16-byte multicast store, so each half carries its own flag; a reader takes the record once both flags show this step.
__device__ __forceinline__ FzRec fz_rec2_wait(const char *p, unsigned flag, int *status) {
    unsigned a, b, c, d, e, f, g, h;
    unsigned spins = 0;
    while (true) {
        asm volatile("ld.relaxed.sys.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p) : "memory");
        if (d == flag && g == flag) break;
        __nanosleep(100);
        if (++spins > FZ_SPIN_LIMIT || *reinterpret_cast<volatile int *>(status + FZ_ST_TIMEOUT)) {
            atomicOr(status + FZ_ST_TIMEOUT, 1);
            break;
        }
    }
    FzRec r;
    r.s = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(e));
    r.c = __uint_as_float(f);
    r.flag = d;
    return r;
}

// 32-byte statistics record {3 doubles, step}: the 16 AdaIN partial sums of one CTA travel as six of them, each written
// with ONE 256-bit store and polled with ONE 256-bit load, so a reader that sees this step's number also sees the data.
// No atomics, no fence, no arrival counter; records never need clearing (step numbers only grow).
__device__ __forceinline__ void fz_stat_store(char *p, double a, double b, double c, unsigned flag) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(__double2loint(a)), "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b)),
                   "r"(__double2loint(c)), "r"(__double2hiint(c)), "r"(flag), "r"(0u) : "memory");
}
__device__ __forceinline__ unsigned fz_stat_load(const char *p, double &a, double &b, double &c) {
    unsigned r0, r1, r2, r3, r4, r5, f, z;
    asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(f), "=r"(z) : "l"(p) : "memory");
    a = __hiloint2double((int)r1, (int)r0);
    b = __hiloint2double((int)r3, (int)r2);
    c = __hiloint2double((int)r5, (int)r4);
    return f;
}

// Bounded spin: wait until *p has reached `target` (monotonic counter).  A peer that died, skipped a step or launched a
// different grid would otherwise hang every GPU of the box inside a cooperative kernel: after ~2 s the waiter sets the
// TIMEOUT status word (srx_plan_check reports it) and carries on, and every later wait of this launch gives up at once.
__device__ __forceinline__ bool fz_wait_ge(const unsigned *p, unsigned target, bool sys, int *status) {
    unsigned spins = 0;
    while ((int)(ld_relaxed(p, sys) - target) < 0) {
        if (++spins > FZ_SPIN_FAST) {
            __nanosleep(500);
            if (spins > FZ_SPIN_FAST + FZ_SPIN_LIMIT || *reinterpret_cast<volatile int *>(status + FZ_ST_TIMEOUT)) {
                atomicOr(status + FZ_ST_TIMEOUT, 1);
                return false;
            }
        }
    }
    return true;
}

// Grid-wide (and, with peers, box-wide) barrier `b`: every CTA adds 1 to counter [b][my rank] on each participant and
// waits until each of its own counters [b][p] has reached target = step * gridDim.x.  All ranks launch the same grid.
__device__ __forceinline__ void fz_barrier(const FzParams &P, int b, unsigned target, bool cross) {
    fz_fence_gpu();
    __syncthreads();
    const int nsrc = cross ? P.world : 1;
    const bool sys = cross && P.world > 1;
    if (threadIdx.x < nsrc) {
        const int dst = cross ? (int)threadIdx.x : P.rank;
        unsigned *pad = reinterpret_cast<unsigned *>((cross ? P.peers[dst] : P.ws) + P.pads_off) + b * SRX_MAX_PEERS + P.rank;
        // the __threadfence() above made this CTA's writes visible in this GPU's L2, which is also where peers read
        // them; the arrival itself is a relaxed add (a system-scope release would stall for microseconds)
        if (sys) red_relaxed_add(pad, 1u);
        else red_release_add(pad, 1u, false);
        const int src = cross ? (int)threadIdx.x : P.rank;
        const unsigned *mine = reinterpret_cast<const unsigned *>(P.ws + P.pads_off) + b * SRX_MAX_PEERS + src;
        fz_wait_ge(mine, target, sys, P.status);
        // no acquire fence: everything read after this barrier that another SM (or GPU) wrote is loaded past L1
        // (ld.cg / ld.relaxed.sys); the fence would only add an L1 invalidation and ~0.7 us
    }
    __syncthreads();
}

// the dense slot of float32(vertexID) (corresponder.py:331-334, corrmap.py:256-261)
// float32(v) in integer arithmetic: exact below 2^24; in [2^24, 2^25) floats are 2 apart and ties go to the even
// mantissa, i.e. the multiple of 4; from 2^25 on the value is past any slot table (capacity <= 2^25) whatever it rounds to.
__device__ __forceinline__ int fz_slot_of(int v) {
    return ((unsigned)v >> 24) == 1u ? ((v + ((v >> 1) & 1)) & ~1) : v;
}
// key of one pixel: the dense slot of float32(vertexID), or -1 when the pixel is no entry (map_index == 2048 or an all-zero
// id, corrmap.py:266-275)
__device__ __forceinline__ int fz_key(int s, int m, int i, int v, unsigned kcap, int *status) {
    const bool valid = (i != SRX_NO_ID_MAP_INDEX) & ((s | m | i | v) != 0);
    const int slot = fz_slot_of(v);
    const bool inr = (unsigned)slot < kcap;
    if (valid & !inr) atomicOr(status + FZ_ST_KEY_RANGE, 1);
    return (valid & inr) ? slot : -1;
}

template <typename IdT> __device__ __forceinline__ void fz_keys(uint32_t a, unsigned kcap, int *status, int &ka, int &kb);
template <> __device__ __forceinline__ void fz_keys<int4>(uint32_t a, unsigned kcap, int *status, int &ka, int &kb) {
    const int4 A = lds128(a), B = lds128(a + 16);
    ka = fz_key(A.x, A.y, A.z, A.w, kcap, status);
    kb = fz_key(B.x, B.y, B.z, B.w, kcap, status);
}
template <> __device__ __forceinline__ void fz_keys<short4>(uint32_t a, unsigned kcap, int *status, int &ka, int &kb) {
    const int4 v = lds128(a);
    ka = fz_key((int)(short)(v.x & 0xffff), v.x >> 16, (int)(short)(v.y & 0xffff), v.y >> 16, kcap, status);
    kb = fz_key((int)(short)(v.z & 0xffff), v.z >> 16, (int)(short)(v.w & 0xffff), v.w >> 16, kcap, status);
}

__device__ __forceinline__ double fz_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums 16 per-lane values over the warp with 16 shuffles instead of 80: every butterfly stage halves the number of
// values a lane still carries (the lane keeps the half selected by its own bit and hands the other half to its partner).
// On return lane L (L < 16) holds in s[0] the warp-wide sum of value index L's ... see mapping below: index = bit-reversal
// free layout: after stages xor 16, 8, 4, 2 a lane holds value j = (lane >> 1) & 15 ... we only need: lane 2*j holds sum j.
__device__ __forceinline__ double fz_warp_sum16(double (&s)[16], int lane) {
    const unsigned FULL = 0xffffffffu;
    // stage xor 16: 16 -> 8 values; lanes with bit 4 set keep the upper half
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double give = up ? s[j] : s[j + 8];
            const double keep = up ? s[j + 8] : s[j];
            s[j] = keep + __shfl_xor_sync(FULL, give, 16);
        }
    }
    {   // xor 8: 8 -> 4
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double give = up ? s[j] : s[j + 4];
            const double keep = up ? s[j + 4] : s[j];
            s[j] = keep + __shfl_xor_sync(FULL, give, 8);
        }
    }
    {   // xor 4: 4 -> 2
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double give = up ? s[j] : s[j + 2];
            const double keep = up ? s[j + 2] : s[j];
            s[j] = keep + __shfl_xor_sync(FULL, give, 4);
        }
    }
    {   // xor 2: 2 -> 1
        const bool up = lane & 2;
        const double give = up ? s[0] : s[1];
        const double keep = up ? s[1] : s[0];
        s[0] = keep + __shfl_xor_sync(FULL, give, 2);
    }
    // xor 1: both lanes of a pair end with the full sum of value index ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1)
    return s[0] + __shfl_xor_sync(FULL, s[0], 1);
}

// Exchange, owner side: pull the partial sums of this rank's slots from every peer, add in rank order, keep the totals
// as flagged records.  U slots per thread and round, all their loads in flight together (one NVLink round trip per round);
// U * MAXP is bounded by the register budget.
template <int U, int MAXP>
__device__ __forceinline__ void fz_pull_slice(const FzParams &P, const float *acc, long long aoff, int slice, unsigned epoch,
                                              int gtid, int gthreads) {
    (void)acc;
    for (int i0 = gtid; i0 < slice; i0 += U * gthreads) {
        float4 ps[U][MAXP];
        float pc[U][MAXP];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * gthreads;
            const long long k = (long long)P.rank * slice + i;
            live[u] = i < slice && k < (long long)P.kcap;
            // ring order starting at this rank's neighbour: at any moment the ranks read from different peers
            // (array index j = ring position; the sum below still runs in rank order)
#pragma unroll
            for (int j = 0; j < MAXP; ++j) {
                if (j < P.world && live[u]) {
                    int p = P.rank + 1 + j;
                    if (p >= P.world) p -= P.world;
                    const float *pa = reinterpret_cast<const float *>(P.peers[p] + aoff);
                    ps[u][j] = ld_volatile_f4(reinterpret_cast<const float4 *>(pa) + k);
                    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(pc[u][j]) : "l"(pa + (long long)P.kcap * 4 + k) : "memory");
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            // rank p sits at ring position j = p - rank - 1 (mod world); add in rank order so that the total does not
            // depend on which rank owns the slot
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            float c = 0.f;
#pragma unroll
            for (int p = 0; p < MAXP; ++p) {
                if (p < P.world) {
                    int j = p - P.rank - 1;
                    if (j < 0) j += P.world;
                    float4 v = ps[u][0];
                    float vc = pc[u][0];
#pragma unroll
                    for (int q = 1; q < MAXP; ++q)
                        if (q == j) { v = ps[u][q]; vc = pc[u][q]; }
                    if (p == 0) { s = v; c = vc; }
                    else { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; c += vc; }
                }
            }
            fz_rec_store(P.ws + P.ll_off + (long long)(i0 + u * gthreads) * 32, s, c, epoch);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename IdT, typename XT>
__global__ void __launch_bounds__(FZ_THREADS, 1) k_overlap_fused(const __grid_constant__ FzParams P) {
    typedef FzLayout<IdT> L;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    volatile unsigned *s_epoch = reinterpret_cast<volatile unsigned *>(smem + L::MISC_OFF);

    if (tid == 0) {
        for (int s = 0; s < FZ_STAGES; ++s) {
            mbar_init(sbase + L::BAR_OFF + s * 8, 1);                      // full: the producer's arrive + tx bytes
            mbar_init(sbase + L::BAR_OFF + (FZ_STAGES + s) * 8, FZ_CONS);  // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *s_epoch = *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off) + 1u;
        *reinterpret_cast<volatile unsigned *>(smem + L::MISC_OFF + 4) = 0u;
    }
    __syncthreads();
    const unsigned epoch = *s_epoch;                 // 1-based index of this step
    const unsigned fz_epoch_ = epoch;
    FZ_TRACE(0);
    if (tid == 0) {   // ring of per-step kernel start / end stamps on the GPU's global timer
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(reinterpret_cast<unsigned long long *>(P.ws + P.ctrl_off + 256) + (epoch & 31u) * 2, now);
    }
    const unsigned target = epoch * gridDim.x;       // arrival count that completes this step's barriers
    const int par = (int)((epoch - 1u) & 1u);
    float *acc = reinterpret_cast<float *>(P.ws + (long long)par * P.accum_stride);
    float *cnt = acc + (long long)P.kcap * 4;
    const int n = P.h * P.w;
    const unsigned nitems = (unsigned)(P.nrows * P.chunks);

    // ------------------------------------------------------------------------------------------------- phase A
    const bool with_x = P.mode == FZ_MODE_STEP;
    unsigned *s_fill = reinterpret_cast<unsigned *>(smem + L::MISC_OFF + 4);   // pairs / entries of this CTA
    if (P.mode == FZ_MODE_CACHED) {
        // cached plan: this CTA's region of the pool, all warps
        if (warp != FZ_CONS) {
            float4 *other = reinterpret_cast<float4 *>(P.ws + (long long)(par ^ 1) * P.accum_stride);
            if (P.world > 1) {
                const unsigned done = (epoch - 1u) * gridDim.x;
                for (int q = 0; q < P.world; ++q)
                    fz_wait_ge(reinterpret_cast<const unsigned *>(P.ws + P.pads_off) + 1 * SRX_MAX_PEERS + q, done, true, P.status);
            }
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int stride = gridDim.x * FZ_CONS * 32;
            for (int v = blockIdx.x * FZ_CONS * 32 + tid; v < P.accum_vec; v += stride) other[v] = z;
        }
        // the counts depend on the ids only: copy them from the plan instead of reducing them again (the count region
        // of this step's accumulator was cleared during the previous step and nothing else writes it in this mode)
        {
            const float4 *src = reinterpret_cast<const float4 *>(P.cnt_plan);
            float4 *dst = reinterpret_cast<float4 *>(cnt);
            for (int v = blockIdx.x * FZ_THREADS + tid; v < (int)(P.kcap / 4); v += gridDim.x * FZ_THREADS) dst[v] = __ldg(src + v);
        }
        const int G = gridDim.x;
        const int e0 = P.cta_tab[2 * G + blockIdx.x], e1 = e0 + P.cta_tab[G + blockIdx.x];
        const XT *xs = reinterpret_cast<const XT *>(P.x);
        for (int e = e0 + tid; e < e1; e += FZ_THREADS) {
            const int2 en = __ldg(P.pool + e);
            const unsigned cc = (unsigned)en.y;
            const int cell = (int)(cc >> 6);
            const float fm = (float)((cc & 63u) + 1u);
            const int fr = cell / n, ci = cell - fr * n;
            const XT *xp = xs + (long long)fr * 4 * n + ci;
            const float x0 = XIo<XT>::ld(xp), x1 = XIo<XT>::ld(xp + n), x2 = XIo<XT>::ld(xp + 2 * n), x3 = XIo<XT>::ld(xp + 3 * n);
            red_add_f32x4(acc + (long long)en.x * 4, fm * x0, fm * x1, fm * x2, fm * x3);
        }
    } else if (warp == FZ_CONS) {
        // producer.  Items (8 id rows x 32 cells) are numbered in CHUNK-MAJOR order (all rows of column chunk 0, then
        // chunk 1, ...): an item's cost follows the number of entries in it, which varies mostly with the screen
        // position, and the SM count is a multiple of the usual chunks-per-row, so a row-major round-robin would hand the
        // same (busy or empty) screen column to one CTA every time.  The first 70 % of the items are dealt round-robin;
        // the rest is drawn in small batches from a ticket counter so that the CTAs finish together (a counter for ALL
        // items serialises: ~7 ns per same-address atomic is as long as an item takes to stream).  Every CTA overdraws
        // exactly one ticket, so the counter advances by a fixed amount per step and never needs a reset.  The bucketing
        // passes (MARK / EMIT) must see identical deals, so they stay fully static.
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        int stage = 0;
        unsigned ph = 0;
        const char *ids = reinterpret_cast<const char *>(P.ids);
        const XT *x = reinterpret_cast<const XT *>(P.x);
        const unsigned G = gridDim.x;
        const bool dynamic_tail = P.mode == FZ_MODE_STEP;
        const unsigned per_cta = dynamic_tail ? (nitems / G) * 70u / 100u : (nitems + G - 1) / G;
        const unsigned static_end = dynamic_tail ? per_cta * G : nitems;
        const unsigned rest = nitems - static_end;
        // batches of 3..8 items: small enough that the last ones end within a few us of each other, large enough that the
        // ticket counter (one same-address atomic per ~7.5 ns chip-wide) stays far from saturation
        const unsigned batch = min(8u, max(3u, rest / (G * 8u)));
        const unsigned nbatches = (rest + batch - 1) / batch;
        unsigned *tickets = reinterpret_cast<unsigned *>(P.ws + P.ctrl_off + 8);
        // tickets drawn by all earlier streaming steps (cached steps draw none, so this is not a function of `epoch`)
        const unsigned tbase = *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off + 12) * (nbatches + G);
        unsigned ticket = 0, next_ticket = 0, in_batch = 0;
        unsigned k = 0;
        bool drawing = false;
        while (true) {
            unsigned item;
            if (!drawing) {
                item = blockIdx.x + k * G;
                ++k;
                if (k > per_cta || item >= static_end) {          // static share done
                    if (!dynamic_tail) item = nitems;
                    else {
                        drawing = true;
                        if (lane == 0) ticket = atomicAdd(tickets, 1u) - tbase;
                        ticket = __shfl_sync(FULL, ticket, 0);
                        in_batch = 0;
                    }
                }
            }
            if (drawing) {
                if (in_batch == batch) { ticket = __shfl_sync(FULL, next_ticket, 0); in_batch = 0; }
                item = ticket < nbatches ? static_end + ticket * batch + in_batch : nitems;
                if (item < nitems && in_batch == 0 && lane == 0) next_ticket = atomicAdd(tickets, 1u) - tbase;   // hides behind this batch
                if (item >= nitems && ticket < nbatches) {       // ragged last batch: fetch the (overdrawn) next ticket
                    ticket = __shfl_sync(FULL, next_ticket, 0);
                    in_batch = 0;
                    continue;
                }
                ++in_batch;
            }
            mbar_wait(sbase + L::BAR_OFF + (FZ_STAGES + stage) * 8, ph ^ 1u);
            const uint32_t sb = sbase + stage * L::STAGE;
            const uint32_t full = sbase + L::BAR_OFF + stage * 8;
            if (item >= nitems) {                    // end marker for the consumers
                if (lane == 0) {
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sb + L::DESC_OFF), "r"(0), "r"(-1) : "memory");
                    mbar_arrive(full);
                }
                break;
            }
            const int chunk = (int)item / P.nrows;
            const int row = (int)item - chunk * P.nrows;
            const int g = row / P.h;
            const int sy = row - g * P.h;
            const int fl = __ldg(P.fmap + g);
            const int sx0 = chunk * FZ_CELLS;
            const int ncell = min(FZ_CELLS, P.w - sx0);
            if (lane == 0) {
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sb + L::DESC_OFF), "r"((fl * P.h + sy) * P.w + sx0), "r"(ncell) : "memory");
                mbar_expect_tx(full, (uint32_t)(ncell * (64 * FzId<IdT>::PX + (with_x ? 4 * (int)sizeof(XT) : 0))));
            }
            __syncwarp();
            if (lane < 8) {
                const long long px = ((long long)g * P.H + sy * 8 + lane) * P.W + sx0 * 8;
                bulk_g2s_hint(sb + lane * FzId<IdT>::PITCH, ids + px * FzId<IdT>::PX, (uint32_t)(ncell * 8 * FzId<IdT>::PX), full, pol);
            } else if (lane < 12 && with_x) {
                const int ch = lane - 8;
                bulk_g2s(sb + L::LAT_OFF + ch * FZ_CELLS * (int)sizeof(XT),
                         x + ((long long)(fl * 4 + ch) * P.h + sy) * P.w + sx0, (uint32_t)(ncell * (int)sizeof(XT)), full);
            }
            if (++stage == FZ_STAGES) { stage = 0; ph ^= 1u; }
        }
    } else {
        // consumers: first clear the next step's accumulator and statistics while the first copies are in flight
        if (P.mode == FZ_MODE_STEP) {
            if (P.world > 1) {   // peers pulled from that accumulator during the previous step: wait until all are done
                const unsigned done = (epoch - 1u) * gridDim.x;
                for (int q = 0; q < P.world; ++q)
                    fz_wait_ge(reinterpret_cast<const unsigned *>(P.ws + P.pads_off) + 1 * SRX_MAX_PEERS + q, done, true, P.status);
            }
            float4 *other = reinterpret_cast<float4 *>(P.ws + (long long)(par ^ 1) * P.accum_stride);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int stride = gridDim.x * FZ_CONS * 32;
            for (int v = blockIdx.x * FZ_CONS * 32 + tid; v < P.accum_vec; v += stride) other[v] = z;
        }
        const int r = lane >> 2, pr = lane & 3;
        const uint32_t lane_off = r * FzId<IdT>::PITCH + pr * 2 * FzId<IdT>::PX;
        int stage = 0;
        unsigned ph = 0;
        while (true) {
            mbar_wait(sbase + L::BAR_OFF + stage * 8, ph);
            const uint32_t sb = sbase + stage * L::STAGE;
            const int2 desc = lds64(sb + L::DESC_OFF);
            if (desc.y < 0) break;                    // end marker
            int ka[FZ_CPW], kb[FZ_CPW];
            float xv[FZ_CPW][4];
#pragma unroll
            for (int u = 0; u < FZ_CPW; ++u) {
                const int cell = warp * FZ_CPW + u;
                ka[u] = kb[u] = -1;
                if (cell < desc.y)   // warp-uniform (ragged last chunk of a row: the stage holds stale bytes there)
                    fz_keys<IdT>(sb + lane_off + cell * 8 * FzId<IdT>::PX, P.kcap, P.status, ka[u], kb[u]);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) xv[u][ch] = lds_x<XT>(sb + L::LAT_OFF + ch * FZ_CELLS * (int)sizeof(XT), cell);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sbase + L::BAR_OFF + (FZ_STAGES + stage) * 8);   // stage may be refilled
            if (++stage == FZ_STAGES) { stage = 0; ph ^= 1u; }
            if (P.dbg & 2) continue;

#pragma unroll
            for (int u = 0; u < FZ_CPW; ++u) {
                const int cell = warp * FZ_CPW + u;
                if (cell >= desc.y) break;            // warp-uniform
                const int a = ka[u], b = kb[u];
                const int hi = __reduce_max_sync(FULL, max(a, b));
                if (hi < 0) {                         // no entry in this cell
                    if (lane == 0) P.winner[desc.x + cell] = -1;
                    continue;
                }
                // winner = last valid pixel in row-major order: positions 2*lane (a) and 2*lane+1 (b), stored +1
                unsigned wp = 0;
                if (b >= 0) wp = ((unsigned)(2 * lane + 2) << FZ_SLOT_BITS) | (unsigned)b;
                else if (a >= 0) wp = ((unsigned)(2 * lane + 1) << FZ_SLOT_BITS) | (unsigned)a;
                wp = __reduce_max_sync(FULL, wp);
                if (lane == 0) P.winner[desc.x + cell] = (int)(wp & ((1u << FZ_SLOT_BITS) - 1u));
                const unsigned lo = __reduce_min_sync(FULL, min((unsigned)a, (unsigned)b));
                if (P.dbg & 1) continue;
                // reduce-by-key inside the warp -> up to two (key, multiplicity) pairs per lane:
                //   * one key in the whole cell (REDUX min == max): a single pair for the cell;
                //   * otherwise the lane's two horizontally adjacent pixels merge when equal, and vertically adjacent rows
                //     (lane L = row 2i, lane L+4 = row 2i+1, same columns) merge pairwise through two shuffles.  Magnified
                //     textures (several pixels per texel) are where this pays: every merge saves two L2 atomics.
                int k1 = a, k2 = (a == b) ? -1 : b, m1 = (a == b) ? 2 : 1, m2 = 1;
                if ((int)lo == hi) {
                    const int total = __reduce_add_sync(FULL, (a >= 0) + (b >= 0));
                    k1 = lane == 0 ? hi : -1;
                    k2 = -1;
                    m1 = total;
                } else {
                    const bool upper = (lane & 4) == 0;                       // even row of the cell
                    const int o1 = __shfl_xor_sync(FULL, k1, 4), o2 = __shfl_xor_sync(FULL, k2, 4);
                    const int om1 = __shfl_xor_sync(FULL, m1, 4);
                    // first slots: same column pair, rows 2i / 2i+1.  Equal keys: the upper lane takes both.
                    if (k1 >= 0 && o1 == k1) {
                        if (upper) m1 += om1; else k1 = -1;
                    }
                    if (k2 >= 0 && o2 == k2) {                                // second slots (multiplicity 1 on both sides)
                        if (upper) m2 += 1; else k2 = -1;
                    }
                }
                if (P.mode != FZ_MODE_STEP) {
                    // bucketing passes: the pairs are counted / stored instead of reduced
                    if (P.mode == FZ_MODE_MARK) {
                        const int w = (int)(wp & ((1u << FZ_SLOT_BITS) - 1u));
                        const int np = __popc(__ballot_sync(FULL, k1 >= 0)) + __popc(__ballot_sync(FULL, k2 >= 0));
                        if (lane == 0) {
                            P.need[w] = 1;
                            atomicAdd(s_fill, (unsigned)np);
                        }
                    } else {   // EMIT
                        const bool fa = k1 >= 0 && __ldg(P.need + k1) != 0;
                        const bool fb = k2 >= 0 && __ldg(P.need + k2) != 0;
                        const unsigned ba = __ballot_sync(FULL, fa), bb = __ballot_sync(FULL, fb);
                        const int tot = __popc(ba) + __popc(bb);
                        unsigned base = 0;
                        if (lane == 0 && tot) base = atomicAdd(s_fill, (unsigned)tot);
                        base = __shfl_sync(FULL, base, 0);
                        const unsigned lt = (1u << lane) - 1u;
                        const unsigned cellg = (unsigned)(desc.x + cell) << 6;
                        int2 *dst = P.pool + P.cta_tab[2 * gridDim.x + blockIdx.x] + base;
                        if (fa) {
                            dst[__popc(ba & lt)] = make_int2(k1, (int)(cellg | (unsigned)(m1 - 1)));
                            red_add_f32(P.cnt_plan + k1, (float)m1);
                        }
                        if (fb) {
                            dst[__popc(ba) + __popc(bb & lt)] = make_int2(k2, (int)(cellg | (unsigned)(m2 - 1)));
                            red_add_f32(P.cnt_plan + k2, (float)m2);
                        }
                    }
                    continue;
                }
                if (k1 >= 0) {
                    const float fm = (float)m1;
                    red_add_f32x4(acc + (long long)k1 * 4, fm * xv[u][0], fm * xv[u][1], fm * xv[u][2], fm * xv[u][3]);
                    red_add_f32(cnt + k1, fm);
                }
                if (k2 >= 0) {
                    const float fm = (float)m2;
                    red_add_f32x4(acc + (long long)k2 * 4, fm * xv[u][0], fm * xv[u][1], fm * xv[u][2], fm * xv[u][3]);
                    red_add_f32(cnt + k2, fm);
                }
            }
        }
    }

    if (P.mode == FZ_MODE_MARK || P.mode == FZ_MODE_EMIT) {   // bucketing passes end here; the step counter does not move
        __syncthreads();
        if (tid == 0) P.cta_tab[(P.mode == FZ_MODE_MARK ? 0 : gridDim.x) + blockIdx.x] = (int)*s_fill;
        return;
    }

    // ------------------------------------------------------------------------------------------------- barrier 0
    FZ_TRACE(1);
    if (P.dbg & 4) {
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {   // (racy against slow producers: experiments only)
            *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off) = epoch;
            *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off + 12) += 1u;
        }
        return;
    }
    fz_barrier(P, 0, target, true);    // every rank's reductions have landed in its own L2
    FZ_TRACE(2);
    if (blockIdx.x == 0 && tid == 0) {
        *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off) = epoch;
        if (P.mode == FZ_MODE_STEP) *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off + 12) += 1u;   // streaming steps so far
    }

    // ------------------------------------------------------------------------------------------------- phase X
    // Frame-sharded runs.  Rank r owns the slots [r*slice, (r+1)*slice): it PULLS the partial sums of its slice from
    // every peer's accumulator (coalesced loads over NVLink), adds them in rank order and keeps the totals in its own
    // memory as 32-byte records {sum.xyzw, count, step} written with one 256-bit store.  Phase B then fetches the
    // winner's record from the owner (one 256-bit load, remote for foreign slots) and checks the step number inside it,
    // so no second barrier is needed.  Nothing is ever stored to a peer except arrival counters: a system-scope fence
    // behind remote stores costs ~6 us on this fabric, a pull costs one round trip (~2 us) — measured with
    // tools/nvl_probe.py.
    const bool xchg = P.world > 1;
    const int slice = P.ll_slice;
    if (xchg && P.mc) {
        // NVLS form.  Owner side: the switch reduces this rank's slice over all ranks (multimem.ld_reduce), the totals go
        // out to every rank as flagged records (multimem.st) — no pull of world-1 partial slices, no signal round, and
        // phase B below polls LOCAL records instead of fetching them from their owners.
        const int gthreads = gridDim.x * FZ_THREADS;
        const char *mca = P.mc + (long long)par * P.accum_stride;
        char *mct = P.mc + P.ll_off;
        for (int base = blockIdx.x * FZ_THREADS + (tid - lane); base < slice; base += gthreads) {
            const int i = base + lane;
            const long long k = (long long)P.rank * slice + i;
            const bool live = i < slice && k < (long long)P.kcap;     // uniform over every aligned group of 4 lanes
            float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), cv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) sv = mm_ld_reduce_f4(mca + k * 16);
            if (live && (lane & 3) == 0) cv = mm_ld_reduce_f4(mca + (long long)P.kcap * 16 + k * 4);   // counts of slots k..k+3
            const int src = lane & ~3;
            const float c0 = __shfl_sync(FULL, cv.x, src), c1 = __shfl_sync(FULL, cv.y, src), c2 = __shfl_sync(FULL, cv.z, src),
                        c3 = __shfl_sync(FULL, cv.w, src);
            const float c = (lane & 2) ? ((lane & 1) ? c3 : c2) : ((lane & 1) ? c1 : c0);
            if (live && c > 0.f) {                    // slots nobody contributed to are never looked up
                mm_st_f4(mct + k * 32, sv.x, sv.y, sv.z, __uint_as_float(epoch));
                mm_st_f4(mct + k * 32 + 16, sv.w, c, __uint_as_float(epoch), 0.f);
            }
        }
        FZ_TRACE(3);
        // this CTA no longer reads any rank's accumulator of this step (they are cleared two steps on): one multicast
        // arrival tells every rank
        __syncthreads();
        if (tid == 0) mm_red_add_u32(P.mc + P.pads_off + (1 * SRX_MAX_PEERS + P.rank) * 4, 1u);
        FZ_TRACE(4);
    } else if (xchg) {
        const int gtid = blockIdx.x * FZ_THREADS + tid, gthreads = gridDim.x * FZ_THREADS;
        const long long aoff = (long long)par * P.accum_stride;
        if (P.world <= 4) fz_pull_slice<2, 4>(P, acc, aoff, slice, epoch, gtid, gthreads);
        else fz_pull_slice<1, SRX_MAX_PEERS>(P, acc, aoff, slice, epoch, gtid, gthreads);
        FZ_TRACE(3);
        // Tell every peer that this CTA (a) has stored its share of this rank's total records — the fence puts them in
        // this GPU's L2 first — and (b) no longer reads the peer's accumulator of this step (peers clear it two steps on).
        fz_fence_gpu();
        __syncthreads();
        if (tid < P.world)
            red_relaxed_add(reinterpret_cast<unsigned *>(P.peers[tid] + P.pads_off) + 1 * SRX_MAX_PEERS + P.rank, 1u);
        // ... and wait until every rank's records are complete: polling a record over NVLink before it is written costs
        // a round trip per retry, polling these local counters costs nothing
        if (tid < P.world) {
            fz_wait_ge(reinterpret_cast<const unsigned *>(P.ws + P.pads_off) + 1 * SRX_MAX_PEERS + tid, target, true, P.status);
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        __syncthreads();
        FZ_TRACE(4);
    }

    // ------------------------------------------------------------------------------------------------- phases B, C
    // All cells of all latent frames form one linear range that is cut into gridDim equal pieces: CTA c owns the cells
    // [start(c), start(c+1)), i.e. one or more SEGMENTS (frame f, cells [a, b) of it).  Equal pieces keep every SM busy
    // whatever the frame count (96 frames on 148 SMs used to leave 52 SMs idle and the others with a whole 16 K-cell
    // frame each: 39 us of gather + AdaIN on cfg3).
    //   B  per segment: winner's mean, blend, the 16 partial sums (x, x^2, b, b^2 per channel), block-reduced and
    //      published as six flagged 32-byte records in slot (c + f) — unique, since c and f only grow along the range.
    //   C  per segment: poll the records of every CTA that owns a piece of frame f (one L2 round trip once they have
    //      landed; nobody waits inside phase B, so the polls cannot deadlock), fold them in CTA order — every CTA of a
    //      frame computes bit-identical statistics — and re-standardise the own cells.
    // The latents read in B stay in shared memory (the ring is idle by now) for C.
    XT *x = reinterpret_cast<XT *>(P.x);
    double *red = reinterpret_cast<double *>(smem + L::RED_OFF);
    float *coef = reinterpret_cast<float *>(smem + L::COEF_OFF);
    const int G = gridDim.x;
    float4 *sx4 = reinterpret_cast<float4 *>(smem);                       // staged latents of this CTA's cells
    const long long smem_cells = (long long)(FZ_STAGES * L::STAGE / 16);
    const long long NC = (long long)P.batch * n;
    // Two ways of cutting the range.  ALIGNED (at least two CTAs per frame): every frame is cut into grp = G / batch equal
    // pieces, one per CTA — one segment, one block reduction and one record poll per CTA, which is what matters when the
    // whole phase takes a few microseconds.  LINEAR (otherwise): gridDim equal pieces regardless of frame boundaries.
    const int grp = 2 * P.batch <= G ? G / P.batch : 0;
    auto cta_start = [&](int c) -> long long {
        if (grp) {
            if (c >= P.batch * grp) return NC;
            const int f = c / grp, part = c - f * grp;
            return (long long)f * n + (((long long)n * part / grp) & ~7ll);
        }
        return c >= G ? NC : ((NC * c) / G) & ~7ll;
    };
    const long long beg = cta_start((int)blockIdx.x), end = cta_start((int)blockIdx.x + 1);
    char *slots = P.ws + P.stats_off;
    for (long long p0 = beg; p0 < end;) {
        const int f = (int)(p0 / n);
        const int s0 = (int)(p0 - (long long)f * n);
        const int s1 = (int)min((long long)n, end - (long long)f * n);
        const long long soff = p0 - beg - s0;                             // staged index of cell ci of this segment = soff + ci
        p0 += s1 - s0;
        XT *xf = x + (long long)f * 4 * n;
        const int *wf = P.winner + (long long)f * n;
        double sums[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sums[j] = 0.0;
        for (int c0 = s0 + tid; c0 < s1; c0 += FZ_BU * FZ_THREADS) {
            // FZ_BU cells per thread per round, loads grouped by dependency level (winner -> accumulator) so that a
            // round costs two L2 round trips instead of 2 * FZ_BU  (requesting the next round's winners a round ahead was
            // measured slower: 21 instead of 17.5 us on cfg3)
            int slot[FZ_BU];
            float xv[FZ_BU][4];
            float4 a[FZ_BU];
            float cn[FZ_BU];
#pragma unroll
            for (int u = 0; u < FZ_BU; ++u) {
                const int ci = c0 + u * FZ_THREADS;
                slot[u] = ci < s1 ? __ldcg(wf + ci) : -2;
            }
#pragma unroll
            for (int u = 0; u < FZ_BU; ++u) {
                const int ci = min(c0 + u * FZ_THREADS, s1 - 1);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) xv[u][ch] = XIo<XT>::ld(xf + (long long)ch * n + ci);
            }
            if (P.adain) {
#pragma unroll
                for (int u = 0; u < FZ_BU; ++u) {
                    const int ci = c0 + u * FZ_THREADS;
                    if (ci < s1 && soff + ci < smem_cells) sx4[soff + ci] = make_float4(xv[u][0], xv[u][1], xv[u][2], xv[u][3]);
                }
            }
#pragma unroll
            for (int u = 0; u < FZ_BU; ++u) {
                a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                cn[u] = 1.f;
                if (slot[u] >= 0) {
                    if (xchg && P.mc) {               // NVLS: the owner pushed this step's total into every rank's table
                        const FzRec r = fz_rec2_wait(P.ws + P.ll_off + (long long)slot[u] * 32, epoch, P.status);
                        a[u] = r.s;
                        cn[u] = r.c;
                    } else if (xchg) {                // the owner's record of this step (remote for foreign slots)
                        const int o = slot[u] / slice;
                        const FzRec r = fz_rec_wait(P.peers[o] + P.ll_off + (long long)(slot[u] - o * slice) * 32, epoch, P.status);
                        a[u] = r.s;
                        cn[u] = r.c;
                    } else {
                        a[u] = __ldcg(reinterpret_cast<const float4 *>(acc) + slot[u]);
                        cn[u] = __ldcg(cnt + slot[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < FZ_BU; ++u) {
                if (slot[u] == -2) continue;
                const int ci = c0 + u * FZ_THREADS;
                float bv[4];
                if (slot[u] >= 0) {
                    // mean, then (1-r)*x + r*m : mul, mul, add, each rounded (corresponder.py:351-352)
                    bv[0] = __fadd_rn(__fmul_rn(P.one_minus, xv[u][0]), __fmul_rn(P.ratio, __fdiv_rn(a[u].x, cn[u])));
                    bv[1] = __fadd_rn(__fmul_rn(P.one_minus, xv[u][1]), __fmul_rn(P.ratio, __fdiv_rn(a[u].y, cn[u])));
                    bv[2] = __fadd_rn(__fmul_rn(P.one_minus, xv[u][2]), __fmul_rn(P.ratio, __fdiv_rn(a[u].z, cn[u])));
                    bv[3] = __fadd_rn(__fmul_rn(P.one_minus, xv[u][3]), __fmul_rn(P.ratio, __fdiv_rn(a[u].w, cn[u])));
                    if (!P.adain) {
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) XIo<XT>::st(xf + (long long)ch * n + ci, bv[ch]);
                    }
                } else {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) bv[ch] = xv[u][ch];
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const double dx = xv[u][ch], db = bv[ch];
                    sums[ch * 4 + 0] += dx; sums[ch * 4 + 1] += dx * dx; sums[ch * 4 + 2] += db; sums[ch * 4 + 3] += db * db;
                }
            }
        }
        if (!P.adain) continue;
        {
            const double v = fz_warp_sum16(sums, lane);
            // lane pair (2q, 2q+1) holds value index j with bits: j3 = lane bit 4, j2 = bit 3, j1 = bit 2, j0 = bit 1
            const int j = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0) red[warp * 16 + j] = v;
        }
        __syncthreads();
        if (tid < 16) {
            double v = 0.0;
            for (int k = 0; k < FZ_THREADS / 32; ++k) v += red[k * 16 + tid];
            red[tid] = v;                              // thread t only ever touches column t of the scratch rows
        }
        if (warp == 0) {
            __syncwarp();
            if (lane < 6) {
                const double va = red[3 * lane], vb = lane < 5 ? red[3 * lane + 1] : 0.0, vc = lane < 5 ? red[3 * lane + 2] : 0.0;
                fz_stat_store(slots + ((long long)(blockIdx.x + f) * 6 + lane) * 32, va, vb, vc, epoch);
            }
        }
        __syncthreads();                               // the scratch is reused by the next segment
    }
    FZ_TRACE(5);
    if (P.adain) {
        for (long long p0 = beg; p0 < end;) {
            const int f = (int)(p0 / n);
            const int s0 = (int)(p0 - (long long)f * n);
            const int s1 = (int)min((long long)n, end - (long long)f * n);
            const long long soff = p0 - beg - s0;
            p0 += s1 - s0;
            XT *xf = x + (long long)f * 4 * n;
            if (warp == 0) {
                // the CTAs that own a piece of frame f: c_lo holds its first cell, c_hi its last
                const long long fa = (long long)f * n, fb = fa + n - 1;
                int c_lo, c_hi;
                if (grp) { c_lo = f * grp; c_hi = c_lo + grp - 1; }
                else {
                    c_lo = (int)((fa * G) / NC), c_hi = (int)((fb * G) / NC);
                    while (c_lo + 1 < G && cta_start(c_lo + 1) <= fa) ++c_lo;
                    while (c_lo > 0 && cta_start(c_lo) > fa) --c_lo;
                    while (c_hi + 1 < G && cta_start(c_hi + 1) <= fb) ++c_hi;
                    while (c_hi > 0 && cta_start(c_hi) > fb) --c_hi;
                }
                // lanes 0..29: five CTAs at a time, record r6 = lane % 6 of CTA c_lo + lane / 6 (+ 5 per round)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                if (lane < 30) {
                    const int r6 = lane % 6;
                    for (int q = c_lo + lane / 6; q <= c_hi; q += 5) {
                        if (cta_start(q) >= cta_start(q + 1)) continue;       // a CTA without cells publishes nothing
                        const char *src = slots + ((long long)(q + f) * 6 + r6) * 32;
                        double va, vb, vc;
                        unsigned spins = 0;
                        while (fz_stat_load(src, va, vb, vc) != epoch) {
                            if (++spins > FZ_SPIN_FAST) {
                                __nanosleep(500);
                                if (spins > FZ_SPIN_FAST + FZ_SPIN_LIMIT || *reinterpret_cast<volatile int *>(P.status + FZ_ST_TIMEOUT)) {
                                    atomicOr(P.status + FZ_ST_TIMEOUT, 1);
                                    break;
                                }
                            }
                        }
                        a0 += va; a1 += vb; a2 += vc;
                    }
                }
                // fold the five lane groups: lanes r6, r6 + 6, ..., r6 + 24 hold partial sums of the same three values
                double *fold = red + 32;                   // [5][18] doubles of the scratch
                if (lane < 30) { fold[lane * 3] = a0; fold[lane * 3 + 1] = a1; fold[lane * 3 + 2] = a2; }
                __syncwarp();
                if (lane < 4) {
                    double q4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double v = 0.0;
                        for (int g5 = 0; g5 < 5; ++g5) v += fold[g5 * 18 + lane * 4 + j];
                        q4[j] = v;
                    }
                    const double sx = q4[0], sxx = q4[1], sb = q4[2], sbb = q4[3];
                    // unbiased variance + 1e-5, sqrt (math_utils.py:39-47); 1/n and 1/(n-1) come from the host in double
                    coef[lane * 4 + 0] = (float)(sx * P.inv_n);
                    coef[lane * 4 + 1] = __fsqrt_rn(__fadd_rn((float)((sxx - sx * sx * P.inv_n) * P.inv_nm1), 1e-5f));
                    coef[lane * 4 + 2] = __fsqrt_rn(__fadd_rn((float)((sbb - sb * sb * P.inv_n) * P.inv_nm1), 1e-5f));
                    coef[lane * 4 + 3] = (float)(sb * P.inv_n);
                }
            }
            __syncthreads();
            FZ_TRACE(6);
#pragma unroll 4
            for (int ci = s0 + tid; ci < s1; ci += FZ_THREADS) {
                float v[4];
                if (soff + ci < smem_cells) {
                    const float4 t = sx4[soff + ci];
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) v[ch] = XIo<XT>::ld(xf + (long long)ch * n + ci);
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    // ((x - mu_c) / sigma_c) * sigma_s + mu_s : one rounding per op (math_utils.py:78-80)
                    XIo<XT>::st(xf + (long long)ch * n + ci,
                                __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v[ch], coef[ch * 4 + 0]), coef[ch * 4 + 1]), coef[ch * 4 + 2]),
                                          coef[ch * 4 + 3]));
                }
            }
            __syncthreads();
        }
    }
    FZ_TRACE(7);
    if (tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        unsigned long long *ring = reinterpret_cast<unsigned long long *>(P.ws + P.ctrl_off + 256);
        atomicMax(ring + (epoch & 31u) * 2 + 1, now);
        if (blockIdx.x == 0) {   // prepare the slot of the step after next
            ring[((epoch + 2u) & 31u) * 2] = ~0ull;
            ring[((epoch + 2u) & 31u) * 2 + 1] = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool fz_disabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SRX_NO_FUSED");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

// geometry for which the persistent step applies (decided once, at plan creation)
bool srx_fused_applicable(const srx_plan *p) {
    const srx_plan_desc &d = p->d;
    if (fz_disabled()) return false;
    if (!p->fast_r8 || d.accum_mode != SRX_ACCUM_FAST) return false;
    if (d.lat_w % 8 != 0) return false;   // 16-byte granularity of the bulk copies for 2-byte latents
    if (p->kcap * 5 / 4 > INT_MAX) return false;
    return true;
}

static int fz_grid(const srx_plan *p) { return p->fused_grid > 0 ? p->fused_grid : srx_sm_count_cached(); }

template <typename IdT, typename XT>
static int launch_fused_t(srx_plan *p, const srx_step_args *a, cudaStream_t st, int mode) {
    typedef FzLayout<IdT> L;
    const srx_plan_desc &d = p->d;
    auto kern = k_overlap_fused<IdT, XT>;
    static bool configured = false;   // per template instance
    static int dev_configured = -1;
    int dev = 0;
    SRX_CUDA_CHECK(cudaGetDevice(&dev));
    if (!configured || dev_configured != dev) {
        SRX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
        configured = true;
        dev_configured = dev;
    }
    FzParams P;
    memset(&P, 0, sizeof(P));
    P.ids = a->ids_dev;
    P.x = a->x_dev;
    P.fmap = p->fmap;
    P.ws = p->ws;
    for (int i = 0; i < SRX_MAX_PEERS; ++i) P.peers[i] = p->world > 1 ? p->peers[i] : nullptr;
    P.peers[p->world > 1 ? p->rank : 0] = p->ws;
    P.ll_off = p->ll_off;
    P.mc = (p->world > 1 && p->mc && p->kcap % 4 == 0 && p->ll_bytes >= p->kcap * 32) ? p->mc : nullptr;
    P.ll_slice = (int)(((p->kcap + p->world - 1) / p->world + 3) / 4 * 4);   // multiple of 4: count vectors never straddle owners
    P.accum_stride = p->accum_stride;
    P.pads_off = p->pads_off;
    P.ctrl_off = p->ctrl_off;
    P.stats_off = p->stats_off;
    P.winner = reinterpret_cast<int *>(p->ws + p->winner_off);
    P.status = reinterpret_cast<int *>(p->ws + p->status_off);
    P.H = d.height; P.W = d.width; P.h = d.lat_h; P.w = d.lat_w; P.batch = d.batch;
    P.kcap = (unsigned)p->kcap;
    P.nrows = d.frames * d.lat_h;
    P.chunks = (d.lat_w + FZ_CELLS - 1) / FZ_CELLS;
    P.ratio = a->ratio;
    P.one_minus = (float)(1.0 - (double)a->ratio);   // python evaluates (1 - ratio) in double
    P.adain = a->adain;
    P.world = p->world;
    P.rank = p->world > 1 ? p->rank : 0;
    P.accum_vec = (int)(p->accum_bytes / 16);
    P.inv_n = 1.0 / ((double)d.lat_h * d.lat_w);
    P.inv_nm1 = 1.0 / ((double)d.lat_h * d.lat_w - 1.0);
    P.mode = mode;
    P.need = reinterpret_cast<unsigned char *>(p->ws + p->need_off);
    P.cta_tab = reinterpret_cast<int *>(p->ws + p->ctatab_off);
    P.pool = reinterpret_cast<int2 *>(p->pool);
    P.cnt_plan = reinterpret_cast<float *>(p->ws + p->cntp_off);
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("SRX_FZ_DEBUG"); dbg = e ? atoi(e) : 0; }
        P.dbg = dbg;
    }

    cudaLaunchConfig_t cfg = {};
    // one CTA per SM on every rank (barrier targets rely on equal grids)
    cfg.gridDim = dim3((unsigned)fz_grid(p));
    cfg.blockDim = dim3(FZ_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    static int coop = -1;
    if (coop < 0) { const char *e = getenv("SRX_FZ_COOP"); coop = (e && e[0] == '0') ? 0 : 1; }
    cfg.numAttrs = coop ? 1 : 0;
    SRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, P));
    return SRX_OK;
}

template <typename IdT>
static int launch_fused_x(srx_plan *p, const srx_step_args *a, cudaStream_t st, int mode) {
    switch (a->x_dtype) {
        case SRX_F32: return launch_fused_t<IdT, float>(p, a, st, mode);
        case SRX_F16: return launch_fused_t<IdT, __half>(p, a, st, mode);
        default: return launch_fused_t<IdT, __nv_bfloat16>(p, a, st, mode);
    }
}

int srx_launch_fused(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->ids_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->x_dev) & 15) == 0,
                SRX_ERR_INVALID, "id buffers and latents must be 16-byte aligned");
    int mode = FZ_MODE_STEP;
    if (!a->ids_dev) {
        SRX_REQUIRE(p->cache_ready && p->cache_grid == fz_grid(p), SRX_ERR_INVALID,
                    "no ids given and no cached plan: call srx_plan_build_cache first (or pass ids_dev)");
        mode = FZ_MODE_CACHED;
    }
    return p->d.id_dtype == SRX_I32 ? launch_fused_x<int4>(p, a, st, mode) : launch_fused_x<short4>(p, a, st, mode);
}

// Bucketing pass (SURVEY.md §8a K4-K5 for the cached-plan regime).  Two calls:
//   1. pool_dev == NULL: streams the ids once (MARK): per-cell winners, bitmap of winner keys, pairs per CTA; syncs and
//      lays the per-CTA regions out; *pool_bytes receives the size the caller must allocate.
//   2. pool_dev != NULL: streams the ids again (EMIT) and fills the pool with the (key, cell, multiplicity) pairs whose
//      key wins at least one cell.  Afterwards srx_overlap_step with ids_dev == NULL runs from the pool.
extern "C" int srx_plan_build_cache(srx_plan *p, const void *ids_dev, void *pool_dev, int64_t *pool_bytes, void *stream) {
    SRX_REQUIRE(p && p->ws && ids_dev && pool_bytes, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(p->fused, SRX_ERR_UNSUPPORTED, "the cached plan needs the persistent step kernel (8x8 pixels per cell, 4 channels, float accumulators)");
    SRX_REQUIRE((long long)p->d.batch * p->d.lat_h * p->d.lat_w < (1ll << 26), SRX_ERR_UNSUPPORTED, "more than 2^26 latent cells");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(ids_dev) & 15) == 0, SRX_ERR_INVALID, "id buffers must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int G = fz_grid(p);
    int *tab = reinterpret_cast<int *>(p->ws + p->ctatab_off);
    srx_step_args a;
    memset(&a, 0, sizeof(a));
    a.ids_dev = ids_dev;
    a.x_dev = nullptr;
    a.x_dtype = SRX_F32;
    if (!pool_dev) {
        p->cache_ready = false;
        SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->need_off, 0, (size_t)p->kcap, st));
        int rc = p->d.id_dtype == SRX_I32 ? launch_fused_x<int4>(p, &a, st, FZ_MODE_MARK) : launch_fused_x<short4>(p, &a, st, FZ_MODE_MARK);
        if (rc) return rc;
        std::vector<int> host(3 * (size_t)G, 0);
        SRX_CUDA_CHECK(cudaMemcpyAsync(host.data(), tab, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
        SRX_CUDA_CHECK(cudaStreamSynchronize(st));
        long long total = 0;
        for (int b = 0; b < G; ++b) {
            host[2 * G + b] = (int)total;
            total += host[b];
            SRX_REQUIRE(total < (1ll << 31), SRX_ERR_UNSUPPORTED, "more than 2^31 cached pairs");
        }
        SRX_CUDA_CHECK(cudaMemcpyAsync(tab + 2 * G, host.data() + 2 * G, sizeof(int) * G, cudaMemcpyHostToDevice, st));
        SRX_CUDA_CHECK(cudaStreamSynchronize(st));   // host vector goes out of scope
        p->cache_entries_cap = total;
        p->cache_grid = G;
        *pool_bytes = (total > 0 ? total : 1) * 8;
        return SRX_OK;
    }
    SRX_REQUIRE(p->cache_grid == G && p->cache_entries_cap >= 0, SRX_ERR_INVALID, "call with pool_dev == NULL first");
    SRX_REQUIRE(*pool_bytes >= (p->cache_entries_cap > 0 ? p->cache_entries_cap : 1) * 8, SRX_ERR_INVALID, "pool too small");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(pool_dev) & 15) == 0, SRX_ERR_INVALID, "pool must be 16-byte aligned");
    p->pool = pool_dev;
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->cntp_off, 0, (size_t)p->kcap * 4, st));
    int rc = p->d.id_dtype == SRX_I32 ? launch_fused_x<int4>(p, &a, st, FZ_MODE_EMIT) : launch_fused_x<short4>(p, &a, st, FZ_MODE_EMIT);
    if (rc) return rc;
    p->cache_ready = true;
    return SRX_OK;
}

// entries actually kept by the last srx_plan_build_cache (syncs); profiling / tests
extern "C" int srx_plan_cache_entries(srx_plan *p, int64_t *kept, int64_t *capacity, void *stream) {
    SRX_REQUIRE(p && p->ws && kept && capacity, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(p->cache_ready, SRX_ERR_INVALID, "no cached plan");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int G = p->cache_grid;
    std::vector<int> host((size_t)G, 0);
    SRX_CUDA_CHECK(cudaMemcpyAsync(host.data(), reinterpret_cast<int *>(p->ws + p->ctatab_off) + G, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    long long t = 0;
    for (int b = 0; b < G; ++b) t += host[b];
    *kept = t;
    *capacity = p->cache_entries_cap;
    return SRX_OK;
}

// Frame-sharded peer mode: `peer_ws[i]` is rank i's workspace as mapped into this process (CUDA IPC or symmetric
// memory); every rank binds a workspace of the same layout (same key capacity and channel count).
extern "C" int srx_plan_bind_peers(srx_plan *p, int rank, int world, void *const *peer_ws) {
    SRX_REQUIRE(p && p->ws, SRX_ERR_INVALID, "bind the local workspace first");
    SRX_REQUIRE(world >= 1 && world <= SRX_MAX_PEERS && rank >= 0 && rank < world, SRX_ERR_INVALID, "bad rank/world (%d/%d)", rank, world);
    SRX_REQUIRE(p->fused, SRX_ERR_UNSUPPORTED, "peer mode needs the persistent step kernel (8x8 pixels per cell, 4 channels, float accumulators)");
    SRX_REQUIRE(world == 1 || peer_ws, SRX_ERR_INVALID, "null peer table");
    for (int i = 0; i < world; ++i) {
        char *q = world == 1 ? p->ws : reinterpret_cast<char *>(peer_ws[i]);
        if (i == rank) q = p->ws;
        SRX_REQUIRE(q, SRX_ERR_INVALID, "null peer workspace %d", i);
        p->peers[i] = q;
    }
    p->world = world;
    p->rank = rank;
    return SRX_OK;
}

// NVLS: `mc_ws` is the multicast address of the (symmetric) workspace — the same offset on every rank behind one pointer
// (torch symmetric memory: multicast_ptr).  Call after srx_plan_bind_peers, on every rank or on none; NULL switches back to
// the pull exchange.
extern "C" int srx_plan_bind_multicast(srx_plan *p, void *mc_ws) {
    SRX_REQUIRE(p && p->ws, SRX_ERR_INVALID, "bind the local workspace first");
    SRX_REQUIRE(!mc_ws || p->world > 1, SRX_ERR_INVALID, "bind the peers first (srx_plan_bind_peers)");
    SRX_REQUIRE(!mc_ws || (p->kcap % 4 == 0 && p->ll_bytes >= p->kcap * 32), SRX_ERR_UNSUPPORTED,
                "the NVLS exchange needs a key capacity that is a multiple of 4 and at most 4 Mi slots");
    p->mc = reinterpret_cast<char *>(mc_ws);
    return SRX_OK;
}

extern "C" int srx_plan_set_grid(srx_plan *p, int ctas) {
    SRX_REQUIRE(p, SRX_ERR_INVALID, "null plan");
    SRX_REQUIRE(ctas >= 0 && ctas <= srx_sm_count_cached(), SRX_ERR_INVALID, "grid must be between 0 and the SM count");
    p->fused_grid = ctas;
    p->cache_ready = false;   // the cached plan is laid out per CTA
    if (p->ws) {
        // Barrier targets, the accumulator parity, the ticket counter and the statistics records all derive from the
        // device-side step counter times the grid: restart them together (device idle; in a frame-sharded run every rank
        // must do the same before any of them steps again).
        SRX_CUDA_CHECK(cudaDeviceSynchronize());
        SRX_CUDA_CHECK(cudaMemset(p->ws, 0, (size_t)(p->ll_off + p->ll_bytes)));
        SRX_CUDA_CHECK(cudaMemset(p->ws + p->stats_off, 0, (size_t)p->stats_bytes));
    }
    return SRX_OK;
}

// Phase timestamps of the last step (SM clock ticks relative to kernel entry): out[0..7] first CTA, out[8..15] last CTA.
// Indices: 0 entry, 1 phase A done, 2 barrier 0 passed, 3 exchange done, 4 signal round done, 5 gather loop done,
// 6 statistics reduced, 7 exit.  Syncs the stream.  Profiling aid.
extern "C" int srx_plan_read_trace(srx_plan *p, int64_t *out16, void *stream) {
    SRX_REQUIRE(p && p->ws && out16, SRX_ERR_INVALID, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out16 == reinterpret_cast<int64_t *>(-1)) return SRX_ERR_INVALID;
    SRX_CUDA_CHECK(cudaMemcpyAsync(out16, p->ws + p->ctrl_off + 64, 128, cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    return SRX_OK;
}

// Profiling aid: ring of the last 32 steps' (earliest CTA start, latest CTA end) on the GPU global timer [ns], slot = step & 31,
// followed by the first CTA's 8 phase stamps (SM clock) of each of those steps: out64[64 + 8 * slot + idx].
extern "C" int srx_plan_read_step_ring(srx_plan *p, uint64_t *out64, void *stream) {
    SRX_REQUIRE(p && p->ws && out64, SRX_ERR_INVALID, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SRX_CUDA_CHECK(cudaMemcpyAsync(out64, p->ws + p->ctrl_off + 256, 512, cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaMemcpyAsync(out64 + 64, p->ws + p->ctrl_off + 1024, 2048, cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    return SRX_OK;
}
