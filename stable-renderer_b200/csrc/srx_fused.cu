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
//                 shared memory (a warp per cell, two adjacent pixels per lane), reduce equal keys inside the warp with
//                 REDUX, and issue one 128-bit vector reduction + one count reduction per distinct key into the
//                 L2-resident accumulator; the cell's last valid pixel is its winner.  While the first copies are in
//                 flight the consumers clear the accumulator of the NEXT step (double buffered by step parity).
//   X  exchange : (frame-sharded runs) every CTA signals every peer; rank r then owns a 1/world slice of the
//                 accumulator: it loads that slice from all peers, adds in rank order and stores the total back into
//                 every peer — reduce-scatter and all-gather in one pass over NVLink, 2 signal rounds, no NCCL call.
//   B  gather   : cells are split evenly over the CTAs: winner's mean, blend, per-frame sums of x, x^2, b, b^2 (double)
//   C  AdaIN    : after one more grid-wide barrier: re-standardise x in place.
// Barriers are monotonic arrival counters in the workspace (targets derive from a device-side step counter, so the
// kernel replays unchanged from a CUDA graph); the launch is cooperative so that all CTAs are co-resident.
#include "srx_plan.cuh"
#include <stdlib.h>

#define FZ_CONS 16                        // consumer warps
#define FZ_THREADS ((FZ_CONS + 1) * 32)   // + one producer warp
#define FZ_CELLS 32                       // cells per stage
#define FZ_CPW (FZ_CELLS / FZ_CONS)       // cells per consumer warp per stage
#define FZ_STAGES 6
#define FZ_SLOT_BITS 25

enum { FZ_ST_KEY_RANGE = 0 };

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
    int dbg;                  // SRX_FZ_DEBUG experiment bits (results are wrong when set): 1 = no reductions,
                              // 2 = consumers only drain the ring, 4 = stop after phase A
};

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p, bool sys) {
    unsigned v;
    if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Grid-wide (and, with peers, box-wide) barrier `b`: every CTA adds 1 to counter [b][my rank] on each participant and
// waits until each of its own counters [b][p] has reached target = step * gridDim.x.  All ranks launch the same grid.
__device__ __forceinline__ void fz_barrier(const FzParams &P, int b, unsigned target, bool cross) {
    __threadfence();
    __syncthreads();
    const int nsrc = cross ? P.world : 1;
    const bool sys = cross && P.world > 1;
    if (threadIdx.x < nsrc) {
        if (sys) __threadfence_system();
        const int dst = cross ? (int)threadIdx.x : P.rank;
        unsigned *pad = reinterpret_cast<unsigned *>((cross ? P.peers[dst] : P.ws) + P.pads_off) + b * SRX_MAX_PEERS + P.rank;
        red_release_add(pad, 1u, sys);
        const int src = cross ? (int)threadIdx.x : P.rank;
        const unsigned *mine = reinterpret_cast<const unsigned *>(P.ws + P.pads_off) + b * SRX_MAX_PEERS + src;
        while ((int)(ld_acquire(mine, sys) - target) < 0) __nanosleep(32);
    }
    __syncthreads();
}

// key of one pixel: the dense slot of float32(vertexID) (corresponder.py:331-334, corrmap.py:256-261) or -1 when the
// pixel is no entry (map_index == 2048 or an all-zero id, corrmap.py:266-275)
__device__ __forceinline__ int fz_key(int s, int m, int i, int v, unsigned kcap, int *status) {
    const bool valid = (i != SRX_NO_ID_MAP_INDEX) & ((s | m | i | v) != 0);
    const int slot = __float2int_rz(__int2float_rn(v));
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
    }
    __syncthreads();
    const unsigned epoch = *s_epoch;                 // 1-based index of this step
    const unsigned target = epoch * gridDim.x;       // arrival count that completes this step's barriers
    const int par = (int)((epoch - 1u) & 1u);
    float *acc = reinterpret_cast<float *>(P.ws + (long long)par * P.accum_stride);
    float *cnt = acc + (long long)P.kcap * 4;
    const int n = P.h * P.w;
    const int nitems = P.nrows * P.chunks;

    // ------------------------------------------------------------------------------------------------- phase A
    if (warp == FZ_CONS) {
        // producer
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        int stage = 0;
        unsigned ph = 0;
        const char *ids = reinterpret_cast<const char *>(P.ids);
        const XT *x = reinterpret_cast<const XT *>(P.x);
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            mbar_wait(sbase + L::BAR_OFF + (FZ_STAGES + stage) * 8, ph ^ 1u);
            const int row = item / P.chunks;
            const int chunk = item - row * P.chunks;
            const int g = row / P.h;
            const int sy = row - g * P.h;
            const int fl = __ldg(P.fmap + g);
            const int sx0 = chunk * FZ_CELLS;
            const int ncell = min(FZ_CELLS, P.w - sx0);
            const uint32_t sb = sbase + stage * L::STAGE;
            const uint32_t full = sbase + L::BAR_OFF + stage * 8;
            if (lane == 0) {
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sb + L::DESC_OFF), "r"((fl * P.h + sy) * P.w + sx0), "r"(ncell) : "memory");
                mbar_expect_tx(full, (uint32_t)(ncell * (64 * FzId<IdT>::PX + 4 * (int)sizeof(XT))));
            }
            __syncwarp();
            if (lane < 8) {
                const long long px = ((long long)g * P.H + sy * 8 + lane) * P.W + sx0 * 8;
                bulk_g2s_hint(sb + lane * FzId<IdT>::PITCH, ids + px * FzId<IdT>::PX, (uint32_t)(ncell * 8 * FzId<IdT>::PX), full, pol);
            } else if (lane < 12) {
                const int ch = lane - 8;
                bulk_g2s(sb + L::LAT_OFF + ch * FZ_CELLS * (int)sizeof(XT),
                         x + ((long long)(fl * 4 + ch) * P.h + sy) * P.w + sx0, (uint32_t)(ncell * (int)sizeof(XT)), full);
            }
            if (++stage == FZ_STAGES) { stage = 0; ph ^= 1u; }
        }
    } else {
        // consumers: first clear the next step's accumulator and statistics while the first copies are in flight
        {
            float4 *other = reinterpret_cast<float4 *>(P.ws + (long long)(par ^ 1) * P.accum_stride);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int stride = gridDim.x * FZ_CONS * 32;
            for (int v = blockIdx.x * FZ_CONS * 32 + tid; v < P.accum_vec; v += stride) other[v] = z;
            float4 *st_other = reinterpret_cast<float4 *>(P.ws + P.stats_off + (long long)(par ^ 1) * P.batch * 128);
            for (int v = blockIdx.x * FZ_CONS * 32 + tid; v < P.batch * 8; v += stride) st_other[v] = z;
        }
        const int r = lane >> 2, pr = lane & 3;
        const uint32_t lane_off = r * FzId<IdT>::PITCH + pr * 2 * FzId<IdT>::PX;
        int stage = 0;
        unsigned ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            mbar_wait(sbase + L::BAR_OFF + stage * 8, ph);
            const uint32_t sb = sbase + stage * L::STAGE;
            const int2 desc = lds64(sb + L::DESC_OFF);
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
                if (cell >= desc.y) break;            // warp-uniform (ragged last chunk of a row)
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
                if ((int)lo == hi) {                  // one key in the whole cell: a single reduction
                    const int total = __reduce_add_sync(FULL, (a >= 0) + (b >= 0));
                    if (lane == 0) {
                        const float fm = (float)total;
                        red_add_f32x4(acc + (long long)hi * 4, fm * xv[u][0], fm * xv[u][1], fm * xv[u][2], fm * xv[u][3]);
                        red_add_f32(cnt + hi, fm);
                    }
                } else if (a == b) {                  // both valid (a >= 0 here, otherwise hi < 0 or a != b)
                    if (a >= 0) {
                        red_add_f32x4(acc + (long long)a * 4, 2.f * xv[u][0], 2.f * xv[u][1], 2.f * xv[u][2], 2.f * xv[u][3]);
                        red_add_f32(cnt + a, 2.f);
                    }
                } else {
                    if (a >= 0) {
                        red_add_f32x4(acc + (long long)a * 4, xv[u][0], xv[u][1], xv[u][2], xv[u][3]);
                        red_add_f32(cnt + a, 1.f);
                    }
                    if (b >= 0) {
                        red_add_f32x4(acc + (long long)b * 4, xv[u][0], xv[u][1], xv[u][2], xv[u][3]);
                        red_add_f32(cnt + b, 1.f);
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------------------------------------- barrier 0
    if (P.dbg & 4) {
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off) = epoch;
        return;
    }
    fz_barrier(P, 0, target, true);
    if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile unsigned *>(P.ws + P.ctrl_off) = epoch;

    // ------------------------------------------------------------------------------------------------- phase X
    if (P.world > 1) {
        const int per = (P.accum_vec + P.world - 1) / P.world;
        const int v0 = P.rank * per, v1 = min(P.accum_vec, v0 + per);
        const long long aoff = (long long)par * P.accum_stride;
        for (int v = v0 + blockIdx.x * FZ_THREADS + tid; v < v1; v += gridDim.x * FZ_THREADS) {
            float4 part[SRX_MAX_PEERS];
#pragma unroll
            for (int p = 0; p < SRX_MAX_PEERS; ++p)
                if (p < P.world) part[p] = ld_volatile_f4(reinterpret_cast<const float4 *>(P.peers[p] + aoff) + v);
            float4 s = part[0];
#pragma unroll
            for (int p = 1; p < SRX_MAX_PEERS; ++p)
                if (p < P.world) { s.x += part[p].x; s.y += part[p].y; s.z += part[p].z; s.w += part[p].w; }
#pragma unroll
            for (int p = 0; p < SRX_MAX_PEERS; ++p)
                if (p < P.world) reinterpret_cast<float4 *>(P.peers[p] + aoff)[v] = s;
        }
        fz_barrier(P, 1, target, true);
    }

    // ------------------------------------------------------------------------------------------------- phase B
    XT *x = reinterpret_cast<XT *>(P.x);
    const long long total = (long long)P.batch * n;
    const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
    const long long c0 = min(total, (long long)blockIdx.x * per_cta), c1 = min(total, c0 + per_cta);
    double *red = reinterpret_cast<double *>(smem + L::RED_OFF);
    float *coef = reinterpret_cast<float *>(smem + L::COEF_OFF);
    double *stats = reinterpret_cast<double *>(P.ws + P.stats_off) + (long long)par * P.batch * 16;
    const int f_lo = (int)(c0 / n), f_hi = c1 > c0 ? (int)((c1 - 1) / n) : f_lo - 1;

    for (int f = f_lo; f <= f_hi; ++f) {
        const int s0 = (int)(max(c0, (long long)f * n) - (long long)f * n);
        const int s1 = (int)(min(c1, (long long)(f + 1) * n) - (long long)f * n);
        XT *xf = x + (long long)f * 4 * n;
        const int *wf = P.winner + (long long)f * n;
        double sums[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sums[j] = 0.0;
        for (int ci = s0 + tid; ci < s1; ci += FZ_THREADS) {
            const int slot = __ldcg(wf + ci);
            float xv[4], bv[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) xv[ch] = XIo<XT>::ld(xf + (long long)ch * n + ci);
            if (slot >= 0) {
                const float4 a = __ldcg(reinterpret_cast<const float4 *>(acc) + slot);
                const float cn = __ldcg(cnt + slot);
                // mean, then (1-r)*x + r*m : mul, mul, add, each rounded (corresponder.py:351-352)
                bv[0] = __fadd_rn(__fmul_rn(P.one_minus, xv[0]), __fmul_rn(P.ratio, __fdiv_rn(a.x, cn)));
                bv[1] = __fadd_rn(__fmul_rn(P.one_minus, xv[1]), __fmul_rn(P.ratio, __fdiv_rn(a.y, cn)));
                bv[2] = __fadd_rn(__fmul_rn(P.one_minus, xv[2]), __fmul_rn(P.ratio, __fdiv_rn(a.z, cn)));
                bv[3] = __fadd_rn(__fmul_rn(P.one_minus, xv[3]), __fmul_rn(P.ratio, __fdiv_rn(a.w, cn)));
                if (!P.adain) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) XIo<XT>::st(xf + (long long)ch * n + ci, bv[ch]);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) bv[ch] = xv[ch];
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const double dx = xv[ch], db = bv[ch];
                sums[ch * 4 + 0] += dx; sums[ch * 4 + 1] += dx * dx; sums[ch * 4 + 2] += db; sums[ch * 4 + 3] += db * db;
            }
        }
        if (P.adain) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double v = fz_warp_sum(sums[j]);
                if (lane == 0) red[warp * 16 + j] = v;
            }
            __syncthreads();
            if (tid < 16) {
                double v = 0.0;
                for (int k = 0; k < FZ_THREADS / 32; ++k) v += red[k * 16 + tid];
                atomicAdd(stats + (long long)f * 16 + tid, v);
            }
            __syncthreads();
        }
    }
    if (!P.adain) return;

    // ------------------------------------------------------------------------------------------------- phase C
    fz_barrier(P, 2, target, false);
    for (int f = f_lo; f <= f_hi; ++f) {
        const int s0 = (int)(max(c0, (long long)f * n) - (long long)f * n);
        const int s1 = (int)(min(c1, (long long)(f + 1) * n) - (long long)f * n);
        XT *xf = x + (long long)f * 4 * n;
        if (tid < 4) {
            const double *s = stats + (long long)f * 16 + tid * 4;
            const double sx = __ldcg(s), sxx = __ldcg(s + 1), sb = __ldcg(s + 2), sbb = __ldcg(s + 3);
            const double dn = (double)n;
            // unbiased variance + 1e-5, sqrt (math_utils.py:39-47)
            coef[tid * 4 + 0] = (float)(sx / dn);
            coef[tid * 4 + 1] = __fsqrt_rn(__fadd_rn((float)((sxx - sx * sx / dn) / (dn - 1.0)), 1e-5f));
            coef[tid * 4 + 2] = __fsqrt_rn(__fadd_rn((float)((sbb - sb * sb / dn) / (dn - 1.0)), 1e-5f));
            coef[tid * 4 + 3] = (float)(sb / dn);
        }
        __syncthreads();
        for (int ci = s0 + tid; ci < s1; ci += FZ_THREADS) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                XT *p = xf + (long long)ch * n + ci;
                const float v = XIo<XT>::ld(p);
                // ((x - mu_c) / sigma_c) * sigma_s + mu_s : one rounding per op (math_utils.py:78-80)
                XIo<XT>::st(p, __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, coef[ch * 4 + 0]), coef[ch * 4 + 1]), coef[ch * 4 + 2]),
                                         coef[ch * 4 + 3]));
            }
        }
        __syncthreads();
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

template <typename IdT, typename XT>
static int launch_fused_t(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
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
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("SRX_FZ_DEBUG"); dbg = e ? atoi(e) : 0; }
        P.dbg = dbg;
    }

    cudaLaunchConfig_t cfg = {};
    // one CTA per SM on every rank (barrier targets rely on equal grids)
    cfg.gridDim = dim3((unsigned)(p->fused_grid > 0 ? p->fused_grid : srx_sm_count_cached()));
    cfg.blockDim = dim3(FZ_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, P));
    return SRX_OK;
}

template <typename IdT>
static int launch_fused_x(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    switch (a->x_dtype) {
        case SRX_F32: return launch_fused_t<IdT, float>(p, a, st);
        case SRX_F16: return launch_fused_t<IdT, __half>(p, a, st);
        default: return launch_fused_t<IdT, __nv_bfloat16>(p, a, st);
    }
}

int srx_launch_fused(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->ids_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->x_dev) & 15) == 0,
                SRX_ERR_INVALID, "id buffers and latents must be 16-byte aligned");
    return p->d.id_dtype == SRX_I32 ? launch_fused_x<int4>(p, a, st) : launch_fused_x<short4>(p, a, st);
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

extern "C" int srx_plan_set_grid(srx_plan *p, int ctas) {
    SRX_REQUIRE(p, SRX_ERR_INVALID, "null plan");
    SRX_REQUIRE(ctas >= 0 && ctas <= srx_sm_count_cached(), SRX_ERR_INVALID, "grid must be between 0 and the SM count");
    p->fused_grid = ctas;
    return SRX_OK;
}
