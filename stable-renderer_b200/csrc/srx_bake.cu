// srx_bake.cu — UV-texture bake: projects decoded frames into the CorrespondMap atlas.
//
// Replaces CorrespondMap.update / _update (source/engine/static/corrmap.py:578-736), called from
// DefaultCorresponder.finished (source/common_utils/stable_render_utils/corresponder.py:130-155).
//
// Reference semantics (frames strictly sequential; within a frame `index_put_` with duplicate texels):
//   replace*: a texel ends with the colour of the LAST kept pixel (row-major) of the LAST frame that touches it;
//   first*  : texels written before the call are skipped (corrmap.py:720-725); otherwise the last kept pixel of the
//             EARLIEST frame that touches the texel.
// Both are "max over an order key", so all frames are processed in parallel:
//   B1 claim : owner[texel] = atomicMax(order(frame, pixel) + 1)          ids (+mask) streamed once
//   B2 write : the pixel whose order equals owner[texel] converts its colour to fp16 and stores it — pushed by the pixels
//              (small updates) or pulled by the texels through the owner word (k_bake_write_texels)
// The weighted multi-view bake (not in the reference) replaces B1/B2 with a vector-atomic
// weighted sum per texel and a per-texel finalize.
#include "srx_common.cuh"

#include <type_traits>

enum { BK_ST_INDEX = 0 };

struct BakeGeom {
    int k2, texels, C, Cin;
    int H, W;
    int sprite, material, ignore_filter;
    int inverse_masks;
    int first_mode;
    int frames_total;  // frames in this chunk (view-sharded bake: frames of ALL ranks)
    unsigned int pix_offset;   // view-sharded bake: linear index of this rank's first pixel among all ranks' views, else 0
};

// mask and id tests of one pixel that is already in registers: true + its texel when the pixel takes part in the bake
__device__ __forceinline__ bool bake_texel_eval(const IdPx &p, float m, bool has_mask, const BakeGeom &g, bool require_id,
                                                long long *tex, int *status) {
    if (has_mask) {
        if (g.inverse_masks) m = __fsub_rn(1.f, m);  // corrmap.py:651-654
        if (!(m > 0.f)) return false;                // corrmap.py:707
    }
    if (require_id && !id_valid(p)) return false;
    if (!g.ignore_filter) {                         // corrmap.py:710-715
        if (g.sprite >= 0 && p.s != g.sprite) return false;
        if (g.material >= 0 && p.m != g.material) return false;
    }
    int mi = p.i, vid = p.v;
    if (mi < 0) mi += g.k2;          // torch negative-index wrap
    if (vid < 0) vid += g.texels;
    if (mi < 0 || mi >= g.k2 || vid < 0 || vid >= g.texels) {
        atomicOr(status + BK_ST_INDEX, 1);  // the reference raises IndexError at corrmap.py:723/735
        return false;
    }
    *tex = (long long)mi * g.texels + vid;
    return true;
}

template <typename IdT>
__device__ __forceinline__ bool bake_texel(const IdT *__restrict__ ids, const float *__restrict__ masks, long long i,
                                           const BakeGeom &g, bool require_id, long long *tex, int *status) {
    float m = 0.f;
    if (masks) {
        m = masks[i];
        if (!((g.inverse_masks ? __fsub_rn(1.f, m) : m) > 0.f)) return false;   // masked out: the id is not even loaded
    }
    return bake_texel_eval(load_id(ids + i), m, masks != nullptr, g, require_id, tex, status);
}

// order key of pixel i of the chunk (32-bit arithmetic: a chunk holds fewer than 2^32 pixels): `replace` ranks pixels in
// (frame, pixel) order, `first` ranks earlier frames higher
__device__ __forceinline__ unsigned int bake_order1(unsigned int i, unsigned int hw, const BakeGeom &g) {
    if (!g.first_mode) return i + 1u;
    const unsigned int f = i / hw;
    return ((unsigned int)(g.frames_total - 1) - f) * hw + (i - f * hw) + 1u;
}

// Texel 0 of map 0 is what every pixel WITHOUT an id addresses when the caller passes no masks (all-zero id -> map_index 0,
// vertexID 0; the reference writes their colours there too).  With `replace` keys growing along the sweep nearly each of those
// pixels would have to update the same word (measured: the claim pass took 1.4 ms instead of 0.3 ms on config 4).  Claims on
// texel 0 are therefore reduced per thread and warp and leave the warp as one atomic when its loop ends (a block-level
// reduction cost 17 % of the warps' time at the barrier: warps over background finish early).
__device__ __forceinline__ void bake_claim_texel0(unsigned int zmax, const uint8_t *__restrict__ writtens,
                                                  unsigned int *__restrict__ owner, const BakeGeom &g) {
    zmax = __reduce_max_sync(0xffffffffu, zmax);
    if ((threadIdx.x & 31) == 0 && zmax && !(g.first_mode && writtens[0]) && __ldcg(owner) < zmax) atomicMax(owner, zmax);
}

template <typename IdT>
__global__ void __launch_bounds__(256) k_bake_claim(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                     const uint8_t *__restrict__ writtens, unsigned int *__restrict__ owner,
                                                     int *__restrict__ status, BakeGeom g, long long npx) {
    const long long hw = (long long)g.H * g.W;
    unsigned int zmax = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        long long tex;
        if (!bake_texel(ids, masks, i, g, false, &tex, status)) continue;
        const unsigned int order1 = bake_order1((unsigned int)i + g.pix_offset, (unsigned int)hw, g);
        if (tex == 0) { zmax = order1 > zmax ? order1 : zmax; continue; }
        if (g.first_mode && writtens[tex]) continue;
        // the owner word only grows: skip the atomic when a later pixel already claimed the texel (removes almost all
        // traffic to hot texels such as (0,0), which every background pixel addresses when no mask is given)
        if (__ldcg(owner + tex) >= order1) continue;
        atomicMax(owner + tex, order1);
    }
    bake_claim_texel0(zmax, writtens, owner, g);
}

// NP pairs of horizontally adjacent pixels per thread (2 NP consecutive pixels): one 256-bit id load per pair, all owner
// probes in flight together.  The pass is a chain of dependent loads (ids -> owner word -> atomic); ncu: scoreboard stalls on
// top, DRAM at 44 %, 45 G atomic sectors/s.
template <typename IdT, int NP>
__device__ __forceinline__ void bake_load_group(const IdT *__restrict__ ids, const float *__restrict__ masks, long long i0,
                                                IdPx (&px)[2 * NP], float (&m)[2 * NP]) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        load_id_pair(ids + i0 + 2 * q, px[2 * q], px[2 * q + 1]);
        m[2 * q] = 0.f; m[2 * q + 1] = 0.f;
        if (masks) {
            const float2 mm = *reinterpret_cast<const float2 *>(masks + i0 + 2 * q);
            m[2 * q] = mm.x; m[2 * q + 1] = mm.y;
        }
    }
}

template <typename IdT, int NP>
__global__ void __launch_bounds__(256) k_bake_claim_pair(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                          const uint8_t *__restrict__ writtens, unsigned int *__restrict__ owner,
                                                          int *__restrict__ status, BakeGeom g, long long ngroups) {
    const unsigned int hw = (unsigned int)((long long)g.H * g.W);
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int zmax = 0u;
    long long gr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    IdPx px[2 * NP], nx[2 * NP];
    float m[2 * NP], nm[2 * NP];
    if (gr < ngroups) bake_load_group<IdT, NP>(ids, masks, gr * (2 * NP), px, m);
    while (gr < ngroups) {
        // the ids of the next round are requested before this round's owner probes: the two HBM / L2 round trips of a round
        // overlap (ncu source view: 36 % of the stall samples sat on the first use of the ids, 30 % on the owner word)
        if (gr + stride < ngroups) bake_load_group<IdT, NP>(ids, masks, (gr + stride) * (2 * NP), nx, nm);
        const long long i0 = gr * (2 * NP);
        long long t[2 * NP];
        unsigned int cur[2 * NP], ord[2 * NP];
        bool ok[2 * NP];
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) {
            t[k] = 0;
            ok[k] = bake_texel_eval(px[k], m[k], masks != nullptr, g, false, &t[k], status);
            ord[k] = bake_order1((unsigned int)i0 + (unsigned int)k + g.pix_offset, hw, g);
            if (ok[k] && t[k] == 0) { zmax = ord[k] > zmax ? ord[k] : zmax; ok[k] = false; }
            cur[k] = 0xffffffffu;
            if (ok[k]) cur[k] = __ldcg(owner + t[k]);
        }
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) {
            if (!ok[k]) continue;
            if (g.first_mode && writtens[t[k]]) continue;
            // a neighbour on the same texel with a larger key makes this claim redundant (magnified textures)
            if (k + 1 < 2 * NP && ok[k + 1] && t[k + 1] == t[k] && ord[k + 1] > ord[k]) continue;
            if (cur[k] < ord[k]) atomicMax(owner + t[k], ord[k]);
        }
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) { px[k] = nx[k]; m[k] = nm[k]; }
        gr += stride;
    }
    bake_claim_texel0(zmax, writtens, owner, g);
}

// The same pass as a three-stage software pipeline: in one loop round a thread requests the ids of round n+1, issues the owner
// probes of round n and consumes the probes of round n-1, so neither of the two dependent round trips (ids -> texel, texel -> owner
// word) is waited for in the round that issued it.  (ncu source view of k_bake_claim_pair on config 4: 58 % of all warp samples
// sit on the compare that consumes the owner word, profiles/r2_bake_claim_pair_stalls.txt.)  MODE 1 skips the probe and issues
// the atomicMax for every kept pixel (no dependent load at all, four times the atomics).
template <typename IdT, int MODE>
__global__ void __launch_bounds__(256) k_bake_claim_pipe(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                          const uint8_t *__restrict__ writtens, unsigned int *__restrict__ owner,
                                                          int *__restrict__ status, BakeGeom g, long long ngroups) {
    const unsigned int hw = (unsigned int)((long long)g.H * g.W);
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int zmax = 0u;
    long long gr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    IdPx px[2], nx[2];
    float m[2], nm[2];
    long long pt[2] = {0, 0};
    unsigned int pcur[2] = {0u, 0u}, pord[2] = {0u, 0u};
    bool pok[2] = {false, false};
    if (gr < ngroups) bake_load_group<IdT, 1>(ids, masks, gr * 2, px, m);
    while (gr < ngroups) {
        if (gr + stride < ngroups) bake_load_group<IdT, 1>(ids, masks, (gr + stride) * 2, nx, nm);
        const long long i0 = gr * 2;
        long long t[2];
        unsigned int cur[2], ord[2];
        bool ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            t[k] = 0;
            ok[k] = bake_texel_eval(px[k], m[k], masks != nullptr, g, false, &t[k], status);
            ord[k] = bake_order1((unsigned int)i0 + (unsigned int)k + g.pix_offset, hw, g);
            if (ok[k] && t[k] == 0) { zmax = ord[k] > zmax ? ord[k] : zmax; ok[k] = false; }
        }
        if (ok[0] && ok[1] && t[0] == t[1] && ord[1] > ord[0]) ok[0] = false;     // the neighbour's larger key makes this claim redundant
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            cur[k] = 0u;
            if (ok[k] && g.first_mode && writtens[t[k]]) ok[k] = false;
            if (MODE == 1) { if (ok[k]) atomicMax(owner + t[k], ord[k]); }
            else if (ok[k]) cur[k] = __ldcg(owner + t[k]);
        }
        if (MODE != 1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (pok[k] && pcur[k] < pord[k]) atomicMax(owner + pt[k], pord[k]);
                pt[k] = t[k]; pcur[k] = cur[k]; pord[k] = ord[k]; pok[k] = ok[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) { px[k] = nx[k]; m[k] = nm[k]; }
        gr += stride;
    }
    if (MODE != 1) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (pok[k] && pcur[k] < pord[k]) atomicMax(owner + pt[k], pord[k]);
    }
    bake_claim_texel0(zmax, writtens, owner, g);
}

// Which form of the claim pass runs.  Config 4 (64 M pixels onto 16.7 M texels), ms per `replace` bake: probes consumed in the round
// that issued them 0.407, consumed a round later 0.376, no probes 0.355 (profiles/r2_bake_claim_pair_stalls.txt) — an atomic is
// fire-and-forget, a probe is a dependent round trip.  Probes pay when most claims are redundant (many more pixels than texels:
// the atomics on one word serialise), so the probe-free form runs up to 16 pixels per texel.  SRX_BAKE_CLAIM = 0 | 1 | 2 forces
// probe / no probe / pipelined probe.
template <typename IdT>
static void bake_launch_claim_pair(int grid, cudaStream_t st, const IdT *ids, const float *masks, const uint8_t *writtens,
                                   unsigned int *owner, int *status, const BakeGeom &g, long long ngroups, long long ntex) {
    const char *e = getenv("SRX_BAKE_CLAIM");
    const int mode = e ? atoi(e) : (2 * ngroups <= 16 * ntex ? 1 : 0);
    if (mode == 1) k_bake_claim_pipe<IdT, 1><<<grid, 256, 0, st>>>(ids, masks, writtens, owner, status, g, ngroups);
    else if (mode == 2) k_bake_claim_pipe<IdT, 2><<<grid, 256, 0, st>>>(ids, masks, writtens, owner, status, g, ngroups);
    else k_bake_claim_pair<IdT, 1><<<grid, 256, 0, st>>>(ids, masks, writtens, owner, status, g, ngroups);
}

template <typename CT> __device__ __forceinline__ float color_ld(const CT *p);
template <> __device__ __forceinline__ float color_ld<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float color_ld<__half>(const __half *p) { return __half2float(*p); }
template <> __device__ __forceinline__ float color_ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename IdT, typename CT>
__global__ void __launch_bounds__(256) k_bake_write(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                     const CT *__restrict__ colors, uint8_t *__restrict__ writtens,
                                                     const unsigned int *__restrict__ owner, __half *__restrict__ values,
                                                     int *__restrict__ status, BakeGeom g, long long npx) {
    const long long hw = (long long)g.H * g.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        long long tex;
        if (!bake_texel(ids, masks, i, g, false, &tex, status)) continue;
        const long long f = i / hw, pix = i - f * hw;
        const long long order = (g.first_mode ? (g.frames_total - 1 - f) : f) * hw + pix;
        if (owner[tex] != (unsigned int)(order + 1)) continue;  // in first mode owner stays 0 for already-written texels
        const CT *c = colors + i * g.Cin;
        __half *v = values + tex * g.C;
        for (int ch = 0; ch < g.C; ++ch)  // channel fix-up of corrmap.py:681-684
            v[ch] = __float2half_rn(ch < g.Cin ? color_ld<CT>(c + ch) : 1.f);
        writtens[tex] = 1;
    }
}

// B2, texel-major: the owner word names the winning pixel, so the atlas can pull instead of the pixels pushing — one
// pass over the owner words, a colour gather for claimed texels only; the ids are not read a second time and the colours of
// losing pixels are never read.  Used when the views hold at least half as many pixels as the atlas has texels.
template <typename CT>
__global__ void __launch_bounds__(256) k_bake_write_texels(const CT *__restrict__ colors, uint8_t *__restrict__ writtens,
                                                            const unsigned int *__restrict__ owner, __half *__restrict__ values,
                                                            BakeGeom g, long long ntex, int frame_offset, int frames_local,
                                                            int set_written) {
    const long long hw = (long long)g.H * g.W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntex; t += (long long)gridDim.x * blockDim.x) {
        const unsigned int o = __ldcs(owner + t);
        if (o == 0u) continue;
        const long long order = (long long)o - 1;
        const long long fo = order / hw, pix = order - fo * hw;
        const long long fl = (g.first_mode ? (g.frames_total - 1 - fo) : fo) - frame_offset;   // frame among this rank's views
        if (fl < 0 || fl >= frames_local) continue;                                          // another rank holds the winner
        const long long i = fl * hw + pix;
        const CT *c = colors + i * g.Cin;
        __half *v = values + t * g.C;
        if (g.C == 4) {
            __half h4[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) h4[ch] = __float2half_rn(ch < g.Cin ? color_ld<CT>(c + ch) : 1.f);
            *reinterpret_cast<uint2 *>(v) = *reinterpret_cast<const uint2 *>(h4);
        } else {
            for (int ch = 0; ch < g.C; ++ch) v[ch] = __float2half_rn(ch < g.Cin ? color_ld<CT>(c + ch) : 1.f);
        }
        if (set_written) writtens[t] = 1;
    }
}

// View-sharded bake, last step: texels claimed in this call take the winner's value (delta = the ranks' partial atlases summed
// as int32 words: every texel is non-zero on exactly one rank, so the sum is that rank's bit pattern).
__global__ void __launch_bounds__(256) k_bake_merge(const unsigned int *__restrict__ owner, const __half *__restrict__ delta,
                                                     __half *__restrict__ values, uint8_t *__restrict__ writtens, long long ntex, int C) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntex; t += (long long)gridDim.x * blockDim.x) {
        if (__ldcs(owner + t) == 0u) continue;
        for (int ch = 0; ch < C; ++ch) values[t * C + ch] = delta[t * C + ch];
        writtens[t] = 1;
    }
}

// ---- weighted multi-view bake -----------------------------------------------------------------------------------
__device__ __forceinline__ float bake_weight(const __half *__restrict__ nd, long long i, int mode) {
    if (mode == SRX_WEIGHT_UNIFORM || nd == nullptr) return 1.f;
    const float nz = __half2float(nd[i * 4 + 2]);
    const float vn = fabsf(__fsub_rn(__fmul_rn(2.f, nz), 1.f));         // |n . (0,0,1)| of the *0.5+0.5 encoded normal
    float w = __fdiv_rn(1.f, __fadd_rn(fabsf(__fsub_rn(1.f, vn)), 1.f));  // algorithms.py:111-113
    if (mode == SRX_WEIGHT_VIEW_NORMAL_DEPTH) w = __fmul_rn(w, __half2float(nd[i * 4 + 3]));
    return w;
}

template <typename IdT, typename CT>
__global__ void __launch_bounds__(256) k_bake_accum(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                     const CT *__restrict__ colors, const __half *__restrict__ nd,
                                                     float *__restrict__ acc, float *__restrict__ wsum,
                                                     int *__restrict__ status, BakeGeom g, int weight_mode, long long npx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        long long tex;
        if (!bake_texel(ids, masks, i, g, true, &tex, status)) continue;
        const float w = bake_weight(nd, i, weight_mode);
        const CT *c = colors + i * g.Cin;
        float v[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) v[ch] = ch < g.C ? __fmul_rn(w, ch < g.Cin ? color_ld<CT>(c + ch) : 1.f) : 0.f;
        red_add_f32x4(acc + tex * 4, v[0], v[1], v[2], v[3]);
        // RGB colours into an RGBA atlas: the appended alpha is 1 (corrmap.py:681-684), so channel 3 accumulates w * 1 — the
        // weight sum itself; the separate scalar reduction (half of all L2 atomics, which bound this kernel) is not needed
        if (!(g.C == 4 && g.Cin == 3)) red_add_f32(wsum + tex, w);
    }
}

__device__ __forceinline__ float bake_weight_bits(unsigned int nz_depth, int mode) {   // low half n_z, high half depth
    if (mode == SRX_WEIGHT_UNIFORM) return 1.f;
    const float nz = __half2float(__ushort_as_half((unsigned short)(nz_depth & 0xffffu)));
    const float vn = fabsf(__fsub_rn(__fmul_rn(2.f, nz), 1.f));
    float w = __fdiv_rn(1.f, __fadd_rn(fabsf(__fsub_rn(1.f, vn)), 1.f));
    if (mode == SRX_WEIGHT_VIEW_NORMAL_DEPTH) w = __fmul_rn(w, __half2float(__ushort_as_half((unsigned short)(nz_depth >> 16))));
    return w;
}

// RGB float32 colours, two adjacent pixels per thread: 256-bit id load, 128-bit normal+depth load, three 64-bit colour loads.
template <typename IdT>
__global__ void __launch_bounds__(256) k_bake_accum_pair(const IdT *__restrict__ ids, const float *__restrict__ masks,
                                                          const float *__restrict__ colors, const __half *__restrict__ nd,
                                                          float *__restrict__ acc, float *__restrict__ wsum,
                                                          int *__restrict__ status, BakeGeom g, int weight_mode, long long npairs) {
    const bool w_in_alpha = g.C == 4;      // Cin == 3 here: alpha accumulates w * 1
    // (requesting the next round's ids before this round's colours, as the claim kernel does, measured 9 % slower here)
    for (long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x; pr < npairs; pr += (long long)gridDim.x * blockDim.x) {
        const long long i0 = pr * 2;
        IdPx px[2];
        float m[2];
        bake_load_group<IdT, 1>(ids, masks, i0, px, m);
        long long t0 = 0, t1 = 0;
        const bool ok0 = bake_texel_eval(px[0], m[0], masks != nullptr, g, true, &t0, status);
        const bool ok1 = bake_texel_eval(px[1], m[1], masks != nullptr, g, true, &t1, status);
        if (!(ok0 || ok1)) continue;
        float w0 = 1.f, w1 = 1.f;
        if (nd != nullptr && weight_mode != SRX_WEIGHT_UNIFORM) {
            const uint4 q = *reinterpret_cast<const uint4 *>(nd + i0 * 4);
            w0 = bake_weight_bits(q.y, weight_mode);
            w1 = bake_weight_bits(q.w, weight_mode);
        }
        const float2 *c = reinterpret_cast<const float2 *>(colors + i0 * 3);
        const float2 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);       // r0 g0 | b0 r1 | g1 b1
        if (ok0) {
            red_add_f32x4(acc + t0 * 4, __fmul_rn(w0, c0.x), g.C > 1 ? __fmul_rn(w0, c0.y) : 0.f, g.C > 2 ? __fmul_rn(w0, c1.x) : 0.f,
                          w_in_alpha ? w0 : 0.f);
            if (!w_in_alpha) red_add_f32(wsum + t0, w0);
        }
        if (ok1) {
            red_add_f32x4(acc + t1 * 4, __fmul_rn(w1, c1.y), g.C > 1 ? __fmul_rn(w1, c2.x) : 0.f, g.C > 2 ? __fmul_rn(w1, c2.y) : 0.f,
                          w_in_alpha ? w1 : 0.f);
            if (!w_in_alpha) red_add_f32(wsum + t1, w1);
        }
    }
}

__global__ void __launch_bounds__(256) k_bake_finalize(const float *__restrict__ acc, const float *__restrict__ wsum,
                                                        __half *__restrict__ values, uint8_t *__restrict__ writtens,
                                                        long long ntex, int C, int first_mode, int w_in_alpha) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntex; t += (long long)gridDim.x * blockDim.x) {
        const float w = w_in_alpha ? acc[t * 4 + 3] : wsum[t];
        if (!(w > 0.f)) continue;
        if (first_mode && writtens[t]) continue;
        for (int ch = 0; ch < C; ++ch) values[t * C + ch] = __float2half_rn(__fdiv_rn(acc[t * 4 + ch], w));
        writtens[t] = 1;
    }
}

// ---- host ------------------------------------------------------------------------------------------------------------
static inline int64_t bk_align(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" int64_t srx_bake_sharded_workspace_bytes(int k2, int texels, int channels) {
    const int64_t ntex = (int64_t)k2 * texels;
    return bk_align(ntex * 4) + 256 + bk_align(ntex * channels * 2);
}

extern "C" int64_t srx_bake_workspace_bytes(int k2, int texels, int channels, int weight_mode) {
    (void)channels;
    const int64_t ntex = (int64_t)k2 * texels;
    if (weight_mode == SRX_WEIGHT_NONE) return bk_align(ntex * 4) + 256;
    return bk_align(ntex * 16) + bk_align(ntex * 4) + 256;
}

template <typename IdT, typename CT>
static int bake_impl(const srx_bake_args *a, cudaStream_t st) {
    const int64_t ntex = (int64_t)a->k2 * a->texels;
    const long long hw = (long long)a->height * a->width;
    char *ws = reinterpret_cast<char *>(a->workspace_dev);
    BakeGeom g;
    g.k2 = a->k2; g.texels = a->texels; g.C = a->channels; g.Cin = a->color_channels;
    g.H = a->height; g.W = a->width;
    g.sprite = a->sprite_id; g.material = a->material_id; g.ignore_filter = a->ignore_obj_mat_id;
    g.inverse_masks = a->inverse_masks;
    g.first_mode = (a->mode == SRX_BAKE_FIRST || a->mode == SRX_BAKE_FIRST_AVG) ? 1 : 0;
    g.pix_offset = 0u;
    const int sms = srx_sm_count_cached();
    const IdT *ids = reinterpret_cast<const IdT *>(a->ids_dev);
    const CT *colors = reinterpret_cast<const CT *>(a->colors_dev);
    __half *values = reinterpret_cast<__half *>(a->values_dev);
    int *status;
    if (a->weight_mode == SRX_WEIGHT_NONE) {
        unsigned int *owner = reinterpret_cast<unsigned int *>(ws);
        status = reinterpret_cast<int *>(ws + bk_align(ntex * 4));
        if (a->phase <= 1 && !a->defer_status) SRX_CUDA_CHECK(cudaMemsetAsync(status, 0, 256, st));
        if (a->phase != 0) {
            // view-sharded bake of the reference modes: order keys number the views of ALL ranks, the owner words are
            // MAX-reduced between phase 1 and 2, the ranks' partial atlases SUM-reduced (as int32 words) between phase 2 and 3
            SRX_REQUIRE(a->frame_offset >= 0 && a->frames_global >= a->frame_offset + a->frames, SRX_ERR_INVALID,
                        "frames_global must cover frame_offset + frames");
            SRX_REQUIRE((long long)a->frames_global * hw < (1ll << 31), SRX_ERR_UNSUPPORTED,
                        "view-sharded bake: all ranks' views together must stay below 2^31 pixels (order keys travel as int32)");
            SRX_REQUIRE(a->workspace_bytes >= srx_bake_sharded_workspace_bytes(a->k2, a->texels, a->channels), SRX_ERR_INVALID, "workspace too small");
            __half *delta = reinterpret_cast<__half *>(ws + bk_align(ntex * 4) + 256);
            g.frames_total = a->frames_global;
            g.pix_offset = (unsigned int)((long long)a->frame_offset * hw);
            const long long npx = (long long)a->frames * hw;
            long long nbt = (ntex + 255) / 256;
            const int gridt = (int)(nbt < (long long)sms * 8 ? nbt : (long long)sms * 8);
            if (a->phase == 1) {
                SRX_CUDA_CHECK(cudaMemsetAsync(owner, 0, (size_t)ntex * 4, st));
                long long nb = (npx + 255) / 256;
                const int grid = (int)(nb < (long long)sms * 8 ? nb : (long long)sms * 8);
                const bool pair_ok = (npx & 1) == 0 && (reinterpret_cast<uintptr_t>(ids) & 31) == 0 && (reinterpret_cast<uintptr_t>(a->masks_dev) & 7) == 0;
                if (pair_ok) {
                    long long nbp = (npx / 2 + 255) / 256;
                    const int gridp = (int)(nbp < (long long)sms * 8 ? nbp : (long long)sms * 8);
                    bake_launch_claim_pair<IdT>(gridp, st, ids, a->masks_dev, a->writtens_dev, owner, status, g, npx / 2, ntex);
                } else {
                    k_bake_claim<IdT><<<grid, 256, 0, st>>>(ids, a->masks_dev, a->writtens_dev, owner, status, g, npx);
                }
            } else if (a->phase == 2) {
                SRX_CUDA_CHECK(cudaMemsetAsync(delta, 0, (size_t)bk_align(ntex * a->channels * 2), st));
                k_bake_write_texels<CT><<<gridt, 256, 0, st>>>(colors, a->writtens_dev, owner, delta, g, ntex, a->frame_offset, a->frames, 0);
            } else {
                k_bake_merge<<<gridt, 256, 0, st>>>(owner, delta, values, a->writtens_dev, ntex, g.C);
            }
            SRX_CUDA_CHECK(cudaGetLastError());
            if (a->phase != 1 || a->defer_status) return SRX_OK;      // the status word is written by the claim pass only
            int st_claim = 0;
            SRX_CUDA_CHECK(cudaMemcpyAsync(&st_claim, status, sizeof(int), cudaMemcpyDeviceToHost, st));
            SRX_CUDA_CHECK(cudaStreamSynchronize(st));
            if (st_claim)
                return srx_set_error(SRX_ERR_INDEX, "index out of range: a kept pixel addresses (map_index, vertexID) outside the "
                                     "%d x %d atlas (corrmap.py:735)", a->k2, a->texels);
            return SRX_OK;
        }
        // the 32-bit order key holds frames_per_chunk*H*W + 1; longer sequences run as ordered chunks
        const long long max_frames = ((1ll << 32) - 2) / hw;
        SRX_REQUIRE(max_frames >= 1, SRX_ERR_UNSUPPORTED, "frame larger than 2^32 pixels");
        const int step = (int)(max_frames < a->frames ? max_frames : a->frames);
        for (int f0 = 0; f0 < a->frames; f0 += step) {
            const int nf = a->frames - f0 < step ? a->frames - f0 : step;
            g.frames_total = nf;
            const long long npx = (long long)nf * hw;
            long long nb = (npx + 255) / 256;
            const int grid = (int)(nb < (long long)sms * 8 ? nb : (long long)sms * 8);
            const float *masks = a->masks_dev ? a->masks_dev + (long long)f0 * hw : nullptr;
            SRX_CUDA_CHECK(cudaMemsetAsync(owner, 0, (size_t)ntex * 4, st));
            const bool pair_ok = (npx & 1) == 0 && (reinterpret_cast<uintptr_t>(ids + (long long)f0 * hw) & 31) == 0 &&
                                 (reinterpret_cast<uintptr_t>(masks) & 7) == 0;
            if (pair_ok) {   // NP = 2 (four pixels per thread) measured 5 % slower on config 4: 0.436 against 0.415 ms per bake
                long long nbp = (npx / 2 + 255) / 256;
                const int gridp = (int)(nbp < (long long)sms * 8 ? nbp : (long long)sms * 8);
                bake_launch_claim_pair<IdT>(gridp, st, ids + (long long)f0 * hw, masks, a->writtens_dev, owner, status, g, npx / 2, ntex);
            } else {
                k_bake_claim<IdT><<<grid, 256, 0, st>>>(ids + (long long)f0 * hw, masks, a->writtens_dev, owner, status, g, npx);
            }
            if (npx * 2 >= ntex) {
                long long nbt = (ntex + 255) / 256;
                const int gridt = (int)(nbt < (long long)sms * 8 ? nbt : (long long)sms * 8);
                k_bake_write_texels<CT><<<gridt, 256, 0, st>>>(colors + (long long)f0 * hw * g.Cin, a->writtens_dev, owner, values, g, ntex,
                                                               0, nf, 1);
            } else {
                k_bake_write<IdT, CT><<<grid, 256, 0, st>>>(ids + (long long)f0 * hw, masks, colors + (long long)f0 * hw * g.Cin,
                                                            a->writtens_dev, owner, values, status, g, npx);
            }
        }
    } else {
        float *acc = reinterpret_cast<float *>(ws);
        float *wsum = reinterpret_cast<float *>(ws + bk_align(ntex * 16));
        status = reinterpret_cast<int *>(ws + bk_align(ntex * 16) + bk_align(ntex * 4));
        if (a->phase != 2) {
            // (deferred status: the status block stays as it is — sticky until srx_bake_check reads and clears it)
            SRX_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)(bk_align(ntex * 16) + bk_align(ntex * 4) + (a->defer_status ? 0 : 256)), st));
            g.frames_total = a->frames;
            const long long npx = (long long)a->frames * hw;
            long long nb = (npx + 255) / 256;
            const int grid = (int)(nb < (long long)sms * 8 ? nb : (long long)sms * 8);
            const bool pair_ok = std::is_same<CT, float>::value && g.Cin == 3 && (npx & 1) == 0 &&
                                 (reinterpret_cast<uintptr_t>(ids) & 31) == 0 && (reinterpret_cast<uintptr_t>(a->masks_dev) & 7) == 0 &&
                                 (reinterpret_cast<uintptr_t>(a->colors_dev) & 7) == 0 && (reinterpret_cast<uintptr_t>(a->normal_depth_dev) & 15) == 0;
            if (pair_ok) {
                long long nbp = (npx / 2 + 255) / 256;
                const int gridp = (int)(nbp < (long long)sms * 8 ? nbp : (long long)sms * 8);
                k_bake_accum_pair<IdT><<<gridp, 256, 0, st>>>(ids, a->masks_dev, reinterpret_cast<const float *>(a->colors_dev),
                                                              reinterpret_cast<const __half *>(a->normal_depth_dev), acc, wsum, status, g,
                                                              a->weight_mode, npx / 2);
            } else {
                k_bake_accum<IdT, CT><<<grid, 256, 0, st>>>(ids, a->masks_dev, colors, reinterpret_cast<const __half *>(a->normal_depth_dev),
                                                            acc, wsum, status, g, a->weight_mode, npx);
            }
        }
        if (a->phase != 1) {
            long long nb2 = (ntex + 255) / 256;
            const int grid2 = (int)(nb2 < (long long)sms * 8 ? nb2 : (long long)sms * 8);
            k_bake_finalize<<<grid2, 256, 0, st>>>(acc, wsum, values, a->writtens_dev, ntex, g.C, g.first_mode,
                                                   (g.C == 4 && g.Cin == 3) ? 1 : 0);
        }
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    if (a->defer_status) return SRX_OK;           // no host sync: srx_bake_check reports (and clears) the status later
    int st_host = 0;
    SRX_CUDA_CHECK(cudaMemcpyAsync(&st_host, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (st_host)
        return srx_set_error(SRX_ERR_INDEX, "index out of range: a kept pixel addresses (map_index, vertexID) outside the "
                             "%d x %d atlas (corrmap.py:735)", a->k2, a->texels);
    return SRX_OK;
}

template <typename IdT>
static int bake_color_dispatch(const srx_bake_args *a, cudaStream_t st) {
    switch (a->color_dtype) {
        case SRX_F32: return bake_impl<IdT, float>(a, st);
        case SRX_F16: return bake_impl<IdT, __half>(a, st);
        case SRX_BF16: return bake_impl<IdT, __nv_bfloat16>(a, st);
        default: return srx_set_error(SRX_ERR_INVALID, "colour dtype must be f32/f16/bf16");
    }
}

extern "C" int srx_bake_update(const srx_bake_args *a, void *stream) {
    SRX_REQUIRE(a, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(a->values_dev && a->writtens_dev && a->colors_dev && a->ids_dev && a->workspace_dev, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE(a->k2 > 0 && a->texels > 0 && a->channels > 0 && a->channels <= 4, SRX_ERR_INVALID, "atlas must have 1..4 channels");
    SRX_REQUIRE(a->frames > 0 && a->height > 0 && a->width > 0 && a->color_channels > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(a->mode >= SRX_BAKE_REPLACE && a->mode <= SRX_BAKE_FIRST_AVG, SRX_ERR_INVALID, "unknown update mode");
    SRX_REQUIRE(a->weight_mode >= SRX_WEIGHT_NONE && a->weight_mode <= SRX_WEIGHT_VIEW_NORMAL_DEPTH, SRX_ERR_INVALID, "unknown weight mode");
    SRX_REQUIRE(a->phase >= 0 && a->phase <= (a->weight_mode == SRX_WEIGHT_NONE ? 3 : 2), SRX_ERR_INVALID,
                "phase must be 0..2 (weighted bake) or 0..3 (reference modes)");
    // channel fix-up (corrmap.py:681-684): truncate, or append alpha = 1 when C == 4 and the colour has 3 channels
    SRX_REQUIRE(a->color_channels >= a->channels || (a->channels == 4 && a->color_channels == 3), SRX_ERR_INVALID,
                "shape mismatch: colour has %d channels, atlas has %d", a->color_channels, a->channels);
    SRX_REQUIRE(a->weight_mode < SRX_WEIGHT_VIEW_NORMAL || a->normal_depth_dev, SRX_ERR_INVALID, "normal/depth buffer required for this weight mode");
    SRX_REQUIRE(a->workspace_bytes >= srx_bake_workspace_bytes(a->k2, a->texels, a->channels, a->weight_mode), SRX_ERR_INVALID, "workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->id_dtype == SRX_I32) return bake_color_dispatch<int4>(a, st);
    if (a->id_dtype == SRX_I16) return bake_color_dispatch<short4>(a, st);
    return srx_set_error(SRX_ERR_INVALID, "id dtype must be int32 or int16");
}

// Deferred status (args->defer_status = 1): srx_bake_update enqueues its kernels and returns without touching the host; the
// status word in the workspace stays sticky across such calls (the workspace must start zeroed) until this call reads it,
// clears it and reports SRX_ERR_INDEX if any of them addressed a texel outside the atlas.  Syncs the stream.
extern "C" int srx_bake_check(const srx_bake_args *a, void *stream) {
    SRX_REQUIRE(a && a->workspace_dev, SRX_ERR_INVALID, "null argument");
    const long long ntex = (long long)a->k2 * a->texels;
    char *ws = reinterpret_cast<char *>(a->workspace_dev);
    int *status = reinterpret_cast<int *>(a->weight_mode == SRX_WEIGHT_NONE ? ws + bk_align(ntex * 4) : ws + bk_align(ntex * 16) + bk_align(ntex * 4));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int st_host = 0;
    SRX_CUDA_CHECK(cudaMemcpyAsync(&st_host, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaMemsetAsync(status, 0, 256, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (st_host)
        return srx_set_error(SRX_ERR_INDEX, "index out of range: a kept pixel addresses (map_index, vertexID) outside the "
                             "%d x %d atlas (corrmap.py:735)", a->k2, a->texels);
    return SRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// On-disk atlas format (CorrespondMap.dump / Load, source/engine/static/corrmap.py:738-872): k*k PNGs of
// uint8 = clip(255 * value, 0, 255) plus written-flag images.  The quantisation runs here so that only bytes cross PCIe.
// numpy evaluates `255. * float16_array` in float16 (the product of two halves is exact in float32, so one rounding
// to half reproduces it), clips, and truncates to uint8.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_atlas_to_u8(const __half *__restrict__ values, const unsigned char *__restrict__ writtens,
                                                      unsigned char *__restrict__ out_values, unsigned char *__restrict__ out_flags,
                                                      long long n_values, long long n_flags) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_values; i += (long long)gridDim.x * blockDim.x) {
        const float p = __half2float(__float2half_rn(255.f * __half2float(values[i])));   // half product, as numpy
        const float c = fminf(fmaxf(p, 0.f), 255.f);                                      // np.clip(img, 0, 255)
        out_values[i] = (unsigned char)(int)c;                                            // .astype(np.uint8): truncation
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_flags; i += (long long)gridDim.x * blockDim.x)
        out_flags[i] = writtens[i] ? 255 : 0;
}

__global__ void __launch_bounds__(256) k_u8_to_atlas(const unsigned char *__restrict__ in_values, const unsigned char *__restrict__ in_flags,
                                                      __half *__restrict__ values, unsigned char *__restrict__ writtens,
                                                      long long n_values, long long n_flags) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_values; i += (long long)gridDim.x * blockDim.x)
        values[i] = __float2half_rn(__fdiv_rn((float)in_values[i], 255.f));               // float32 / 255., stored as half
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_flags; i += (long long)gridDim.x * blockDim.x)
        writtens[i] = in_flags[i] ? 1 : 0;                                                // (img / 255.).bool()
}

extern "C" int srx_atlas_quantize(const void *values_f16_dev, const uint8_t *writtens_dev, uint8_t *out_values_dev,
                                  uint8_t *out_flags_dev, int64_t n_values, int64_t n_flags, void *stream) {
    SRX_REQUIRE(values_f16_dev && writtens_dev && out_values_dev && out_flags_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(n_values >= 0 && n_flags >= 0, SRX_ERR_INVALID, "bad sizes");
    const long long nb = (n_values + 255) / 256;
    const int grid = (int)(nb < 1 ? 1 : (nb < (long long)srx_sm_count_cached() * 8 ? nb : (long long)srx_sm_count_cached() * 8));
    k_atlas_to_u8<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __half *>(values_f16_dev), writtens_dev,
                                                                            out_values_dev, out_flags_dev, n_values, n_flags);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

extern "C" int srx_atlas_dequantize(const uint8_t *in_values_dev, const uint8_t *in_flags_dev, void *values_f16_dev,
                                    uint8_t *writtens_dev, int64_t n_values, int64_t n_flags, void *stream) {
    SRX_REQUIRE(values_f16_dev && writtens_dev && in_values_dev && in_flags_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(n_values >= 0 && n_flags >= 0, SRX_ERR_INVALID, "bad sizes");
    const long long nb = (n_values + 255) / 256;
    const int grid = (int)(nb < 1 ? 1 : (nb < (long long)srx_sm_count_cached() * 8 ? nb : (long long)srx_sm_count_cached() * 8));
    k_u8_to_atlas<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in_values_dev, in_flags_dev, reinterpret_cast<__half *>(values_f16_dev),
                                                                            writtens_dev, n_values, n_flags);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
