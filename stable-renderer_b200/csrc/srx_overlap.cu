// srx_overlap.cu — the correspondence-map latent overlap step on B200.
//
// Replaces (reference paths relative to /root/reference):
//   IDMap.create_vertex_screen_info            source/engine/static/corrmap.py:220-280
//   OverlapCorresponder.step_finished          source/common_utils/stable_render_utils/corresponder.py:298-376
//   tensor_group_by_then_average               source/common_utils/math_utils.py:86-161
//   adaptive_instance_normalization            source/common_utils/math_utils.py:27-80
//
// Data flow of one step:
//   K1  accumulate : one streaming pass over the id buffers (16 B/pixel, the only HBM-sized traffic).  A warp owns
//                    one latent cell = 8x8 id pixels: it loads the 64 ids with two coalesced 128-bit loads per lane,
//                    reduces equal keys inside the warp (match.any), gathers the cell's latent once and issues one
//                    128-bit vector reduction (REDG.F32x4) + one count reduction per distinct key into the
//                    key-indexed accumulator, which is L2 resident.  The last valid pixel of the cell is the cell's
//                    "winner" (duplicate-index write-back of corresponder.py:354-359).
//   [NCCL all-reduce of the accumulator when frames are sharded over GPUs]
//   K2  finalize   : one thread-block cluster per latent frame: gather the winner key's mean, blend, per-(frame,
//                    channel) statistics reduced over distributed shared memory, AdaIN, in-place write-back.
#include "srx_plan.cuh"
#include <cooperative_groups.h>
#include <new>
#include <stdlib.h>
#include <vector>

namespace cg = cooperative_groups;

// =================================================================================================================
// plan
// =================================================================================================================
// cell coordinate of a pixel coordinate, in the reference's float32 arithmetic:
// trunc(fl32(fl32(p) / fl32(divisor)) * fl32(cells))   (corrmap.py:239,249 ; corresponder.py:312-313)
static int host_cell(int p, int divisor, int cells) {
    volatile float ratio = (float)p / (float)divisor;
    volatile float scaled = ratio * (float)cells;
    return (int)scaled;
}

// -----------------------------------------------------------------------------------------------------------------
// scan: key range, entry count and range check of the cells (plan creation only)
// -----------------------------------------------------------------------------------------------------------------
struct ScanOut {
    unsigned long long n_valid;
    long long key_min, key_max;
    int bad_cell;
    int pad;
};

template <typename IdT>
__global__ void __launch_bounds__(256) k_scan_ids(const IdT *__restrict__ ids, const int *__restrict__ colcell,
                                                   const int *__restrict__ rowcell, int H, int W, int h, int w,
                                                   long long npx, int key_mode, ScanOut *out) {
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;
    unsigned long long cnt = 0;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        IdPx p = load_id(ids + i);
        bool valid = key_mode == SRX_KEY_VERTEX ? id_valid(p) : ((p.s | p.m | p.i | p.v) != 0);
        if (valid) {
            long long k = vertex_slot(p.v);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
            ++cnt;
            int x = (int)(i % W), y = (int)((i / W) % H);
            if (colcell[x] >= w || rowcell[y] >= h) bad = 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) {
            atomicMin(&out->key_min, kmin);
            atomicMax(&out->key_max, kmax);
            atomicAdd(&out->n_valid, cnt);
        }
        if (bad) atomicOr(&out->bad_cell, 1);
    }
}

// =================================================================================================================
// K1 fast path: 8x8 id pixels per latent cell, C == 4, one warp per cell
// =================================================================================================================
enum { ST_KEY_RANGE = 0, ST_CELL_RANGE = 1 };

// Work item = 8 id rows x (8*U cells).  Each of the CTA's 8 warps owns U consecutive cells.  A lane holds two
// horizontally adjacent pixels of a cell (one 256-bit load: lane = row*4 + pair, so lane order is the row-major pixel
// order of the cell) and issues the loads of all U cells plus the cells' latents before consuming any: 4 KB of ids in
// flight per warp, and the latent loads (which do not depend on the ids) overlap with them.
// Reduce-by-key inside the warp uses no shuffles or match instructions: REDUX min/max detect a cell whose valid pixels
// all share one key (one reduction for the whole cell); otherwise equal pixel pairs are merged in-lane.  The winner
// (last valid pixel in row-major order) is a REDUX max over (position, slot) packed in 32 bits.
#define K1_U 4
#define SRX_SLOT_BITS 25   // dense slot limit 2^25: leaves 7 bits for (pixel position + 1) in the packed winner word

template <typename IdT, typename XT, bool DET, bool FROM_SLOTS, bool WRITE_SLOTS>
__global__ void __launch_bounds__(256, 3)
k_accum_r8(const IdT *__restrict__ ids, int *__restrict__ slotmap, const XT *__restrict__ x,
           const int *__restrict__ fmap, void *__restrict__ accum, int *__restrict__ winner, int *__restrict__ status,
           int H, int W, int h, int w, long long kcap, int nrows, int chunks_per_row, int dbg) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, pr = lane & 3;
    float *acc_f = reinterpret_cast<float *>(accum);
    float *cnt_f = acc_f + kcap * 4;
    long long *acc_q = reinterpret_cast<long long *>(accum);
    long long *cnt_q = acc_q + kcap * 4;
    const int plane = h * w;
    const int nitems = nrows * chunks_per_row;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int row = item / chunks_per_row;
        const int chunk = item - row * chunks_per_row;
        const int g = row / h;
        const int sy = row - g * h;
        const int fl = fmap[g];
        const int sx0 = chunk * (8 * K1_U) + warp * K1_U;
        const long long px_base = ((long long)g * H + sy * 8 + r) * W + pr * 2;   // pixel pair (row r, cols 2pr, 2pr+1) of cell 0
        const XT *xrow = x + ((long long)fl * 4 * h + sy) * w;                   // channel 0 of this cell row
        int *wrow = winner + ((long long)fl * h + sy) * w;

        int ka[K1_U], kb[K1_U];   // slot of the lane's two pixels, -1 = no entry
        float xv[K1_U][4];
        if (FROM_SLOTS) {
#pragma unroll
            for (int u = 0; u < K1_U; ++u) {
                const int sx = sx0 + u;
                ka[u] = kb[u] = -1;
                if (sx < w) {
                    const int2 v = __ldg(reinterpret_cast<const int2 *>(slotmap + px_base + sx * 8));
                    ka[u] = v.x;
                    kb[u] = v.y;
                }
            }
        } else {
            IdPx pa[K1_U], pb[K1_U];
#pragma unroll
            for (int u = 0; u < K1_U; ++u) {
                const int sx = sx0 + u;
                pa[u] = IdPx{0, 0, 0, 0};
                pb[u] = IdPx{0, 0, 0, 0};
                if (sx < w) load_id_pair(ids + px_base + sx * 8, pa[u], pb[u]);
            }
#pragma unroll
            for (int u = 0; u < K1_U; ++u) {
                const bool va = id_valid(pa[u]), vb = id_valid(pb[u]);
                long long sa = va ? vertex_slot(pa[u].v) : -1;
                long long sb = vb ? vertex_slot(pb[u].v) : -1;
                if ((sa >= kcap) | (sb >= kcap) | (va & (sa < 0)) | (vb & (sb < 0))) {
                    atomicOr(status + ST_KEY_RANGE, 1);
                    if (sa >= kcap) sa = -1;
                    if (sb >= kcap) sb = -1;
                }
                ka[u] = sa < 0 ? -1 : (int)sa;
                kb[u] = sb < 0 ? -1 : (int)sb;
                if (WRITE_SLOTS && sx0 + u < w)
                    *reinterpret_cast<int2 *>(slotmap + px_base + (sx0 + u) * 8) = make_int2(ka[u], kb[u]);
            }
        }
        // the cells' latents: every lane reads the same 4 addresses (one L1/L2 transaction each, no shuffles)
#pragma unroll
        for (int u = 0; u < K1_U; ++u) {
            const int sx = min(sx0 + u, w - 1);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) xv[u][ch] = XIo<XT>::ld(xrow + ch * plane + sx);
        }

#pragma unroll
        for (int u = 0; u < K1_U; ++u) {
            const int sx = sx0 + u;
            if (sx >= w) break;   // warp-uniform
            const int a = ka[u], b = kb[u];
            const int nvalid = (a >= 0) + (b >= 0);
            // winner: highest (pixel position, slot); positions 2*lane (a) and 2*lane+1 (b), stored +1 so that 0 = none
            unsigned wpack = 0;
            if (b >= 0) wpack = ((unsigned)(2 * lane + 2) << SRX_SLOT_BITS) | (unsigned)b;
            else if (a >= 0) wpack = ((unsigned)(2 * lane + 1) << SRX_SLOT_BITS) | (unsigned)a;
            wpack = __reduce_max_sync(FULL, wpack);
            if (lane == 0) wrow[sx] = wpack ? (int)(wpack & ((1u << SRX_SLOT_BITS) - 1)) : -1;
            if (wpack == 0) continue;   // no valid pixel in this cell (warp-uniform)
            const int lo = __reduce_min_sync(FULL, min(a >= 0 ? a : INT_MAX, b >= 0 ? b : INT_MAX));
            const int hi = __reduce_max_sync(FULL, max(a, b));
            int k_red[2], m_red[2];   // up to two reductions from this lane
            k_red[0] = k_red[1] = -1;
            m_red[0] = m_red[1] = 0;
            if (lo == hi) {           // every valid pixel of the cell carries the same key: one reduction for the cell
                const int total = __reduce_add_sync(FULL, nvalid);
                if (lane == 0) { k_red[0] = lo; m_red[0] = total; }
            } else if (a >= 0 && a == b) {
                k_red[0] = a; m_red[0] = 2;
            } else {
                k_red[0] = a; m_red[0] = 1;
                k_red[1] = b; m_red[1] = 1;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = k_red[j];
                if (k < 0 || (dbg & 1)) continue;
                if (DET) {
                    const long long q = (long long)m_red[j];
                    red_add_s64(acc_q + (long long)k * 4 + 0, q * to_fix(xv[u][0]));
                    red_add_s64(acc_q + (long long)k * 4 + 1, q * to_fix(xv[u][1]));
                    red_add_s64(acc_q + (long long)k * 4 + 2, q * to_fix(xv[u][2]));
                    red_add_s64(acc_q + (long long)k * 4 + 3, q * to_fix(xv[u][3]));
                    red_add_s64(cnt_q + k, q);
                } else {
                    const float fm = (float)m_red[j];
                    red_add_f32x4(acc_f + (long long)k * 4, fm * xv[u][0], fm * xv[u][1], fm * xv[u][2], fm * xv[u][3]);
                    red_add_f32(cnt_f + k, fm);
                }
            }
        }
    }
}

// =================================================================================================================
// K1 generic path: any H/h, W/w ratio, any channel count, several id frames per latent frame.  Thread per pixel.
// =================================================================================================================
template <typename IdT, typename XT, bool DET>
__global__ void __launch_bounds__(256)
k_accum_generic(const IdT *__restrict__ ids, const XT *__restrict__ x, const int *__restrict__ fmap,
                const int *__restrict__ colcell, const int *__restrict__ rowcell, void *__restrict__ accum,
                unsigned long long *__restrict__ winner64, int *__restrict__ status, int H, int W, int h, int w, int C,
                long long kcap, long long npx) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float *acc_f = reinterpret_cast<float *>(accum);
    float *cnt_f = acc_f + kcap * C;
    long long *acc_q = reinterpret_cast<long long *>(accum);
    long long *cnt_q = acc_q + kcap * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // all lanes of a warp run the same number of iterations (warp collectives inside)
    const long long iters = (npx + stride - 1) / stride;
    for (long long it = 0; it < iters; ++it) {
        const long long i = start + it * stride;
        long long cell = -1;
        int slot = -1;
        if (i < npx) {
            IdPx p = load_id(ids + i);
            if (id_valid(p)) {
                const int px = (int)(i % W);
                const long long t = i / W;
                const int py = (int)(t % H);
                const int g = (int)(t / H);
                const int sx = colcell[px], sy = rowcell[py];
                long long s = vertex_slot(p.v);
                if (s < 0 || s >= kcap) {
                    atomicOr(status + ST_KEY_RANGE, 1);
                } else if (sx >= w || sy >= h) {
                    atomicOr(status + ST_CELL_RANGE, 1);
                } else {
                    slot = (int)s;
                    const int fl = fmap[g];
                    cell = ((long long)fl * h + sy) * w + sx;
                    const XT *xp = x + ((long long)fl * C * h + sy) * w + sx;
                    const long long plane = (long long)h * w;
                    if (C == 4 && !DET) {
                        red_add_f32x4(acc_f + (long long)slot * 4, XIo<XT>::ld(xp), XIo<XT>::ld(xp + plane),
                                      XIo<XT>::ld(xp + 2 * plane), XIo<XT>::ld(xp + 3 * plane));
                        red_add_f32(cnt_f + slot, 1.f);
                    } else {
                        for (int ch = 0; ch < C; ++ch) {
                            const float v = XIo<XT>::ld(xp + ch * plane);
                            if (DET) red_add_s64(acc_q + (long long)slot * C + ch, to_fix(v));
                            else red_add_f32(acc_f + (long long)slot * C + ch, v);
                        }
                        if (DET) red_add_s64(cnt_q + slot, 1LL);
                        else red_add_f32(cnt_f + slot, 1.f);
                    }
                }
            }
        }
        // winner: highest entry order per cell.  Lanes of a warp hold consecutive pixels, so the last lane of each
        // run of equal cells carries the warp's candidate; one 64-bit atomicMax per (warp, cell).
        const unsigned same = __match_any_sync(FULL, cell);
        if (cell >= 0 && lane == 31 - __clz(same)) {
            const unsigned long long packed = ((unsigned long long)(i + 1) << 27) | (unsigned long long)slot;
            atomicMax(winner64 + cell, packed);
        }
    }
}

__global__ void k_winner_unpack(const unsigned long long *__restrict__ winner64, int *__restrict__ winner, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long v = winner64[i];
        winner[i] = v ? (int)(v & ((1ull << 27) - 1)) : -1;
    }
}

// =================================================================================================================
// K2 finalize: one thread-block cluster per latent frame
// =================================================================================================================
#define K2_THREADS 512
#define K2_MAXC 4

template <typename XT, bool DET>
__global__ void __launch_bounds__(K2_THREADS)
k_finalize_c4(XT *__restrict__ x, const void *__restrict__ accum, const int *__restrict__ winner, int h, int w,
              long long kcap, float ratio, float one_minus, int adain) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = cluster.num_blocks();
    const int rank = cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int n = h * w;
    const int lo = (int)((long long)n * rank / CL), hi = (int)((long long)n * (rank + 1) / CL);
    const float *acc_f = reinterpret_cast<const float *>(accum);
    const float *cnt_f = acc_f + kcap * 4;
    const long long *acc_q = reinterpret_cast<const long long *>(accum);
    const long long *cnt_q = acc_q + kcap * 4;
    XT *xb = x + (long long)b * 4 * n;
    const int *wb = winner + (long long)b * n;

    // sums[ch][0..3] = sum x, sum x^2, sum b, sum b^2  (double: exact enough that the variance matches torch's
    // double-accumulated CPU var to a float32 ulp)
    double sums[4][4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int j = 0; j < 4; ++j) sums[ch][j] = 0.0;

    for (int cell = lo + threadIdx.x; cell < hi; cell += K2_THREADS) {
        const int slot = wb[cell];
        float xv[4], bv[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) xv[ch] = XIo<XT>::ld(xb + (long long)ch * n + cell);
        if (slot >= 0) {
            float m[4];
            if (DET) {
                const double cn = (double)cnt_q[slot];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) m[ch] = (float)(from_fix(acc_q[(long long)slot * 4 + ch]) / cn);
            } else {
                const float4 a = *reinterpret_cast<const float4 *>(acc_f + (long long)slot * 4);
                const float cn = cnt_f[slot];
                m[0] = __fdiv_rn(a.x, cn); m[1] = __fdiv_rn(a.y, cn); m[2] = __fdiv_rn(a.z, cn); m[3] = __fdiv_rn(a.w, cn);
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)  // (1-r)*x + r*m : mul, mul, add, each rounded (corresponder.py:351-352)
                bv[ch] = __fadd_rn(__fmul_rn(one_minus, xv[ch]), __fmul_rn(ratio, m[ch]));
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) bv[ch] = xv[ch];
        }
        if (!adain) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) XIo<XT>::st(xb + (long long)ch * n + cell, bv[ch]);
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const double dx = xv[ch], db = bv[ch];
                sums[ch][0] += dx; sums[ch][1] += dx * dx; sums[ch][2] += db; sums[ch][3] += db * db;
            }
        }
    }
    if (!adain) return;

    __shared__ double warp_part[K2_THREADS / 32][16];
    __shared__ double cta_part[16];
    __shared__ float coef[4][4];  // mu_c, sigma_c, sigma_s, mu_s per channel
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double v = warp_sum(sums[ch][j]);
            if (lane == 0) warp_part[wid][ch * 4 + j] = v;
        }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
        for (int k = 0; k < K2_THREADS / 32; ++k) v += warp_part[k][threadIdx.x];
        cta_part[threadIdx.x] = v;
    }
    cluster.sync();  // every CTA's partials are visible cluster-wide (distributed shared memory)
    if (threadIdx.x < 4) {
        const int ch = threadIdx.x;
        double s[4] = {0, 0, 0, 0};
        for (int rk = 0; rk < CL; ++rk) {  // fixed rank order: deterministic
            const double *remote = cluster.map_shared_rank(cta_part, rk);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] += remote[ch * 4 + j];
        }
        const double dn = (double)n;
        const double mean_c = s[0] / dn, mean_s = s[2] / dn;
        const double var_c = (s[1] - s[0] * s[0] / dn) / (dn - 1.0);  // unbiased (torch .var default), math_utils.py:39
        const double var_s = (s[3] - s[2] * s[2] / dn) / (dn - 1.0);
        coef[ch][0] = (float)mean_c;
        coef[ch][1] = __fsqrt_rn(__fadd_rn((float)var_c, 1e-5f));
        coef[ch][2] = __fsqrt_rn(__fadd_rn((float)var_s, 1e-5f));
        coef[ch][3] = (float)mean_s;
    }
    cluster.sync();  // remote reads done before any CTA may exit; also orders coef[] for this CTA
    for (int cell = lo + threadIdx.x; cell < hi; cell += K2_THREADS) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            XT *p = xb + (long long)ch * n + cell;
            const float v = XIo<XT>::ld(p);
            // ((x - mu_c) / sigma_c) * sigma_s + mu_s : sub, div, mul, add each rounded (math_utils.py:78-80)
            const float y = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, coef[ch][0]), coef[ch][1]), coef[ch][2]), coef[ch][3]);
            XIo<XT>::st(p, y);
        }
    }
}

// generic channel count: one CTA per (frame, channel) plane
template <typename XT, bool DET>
__global__ void __launch_bounds__(K2_THREADS)
k_finalize_plane(XT *__restrict__ x, const void *__restrict__ accum, const int *__restrict__ winner, int C, int h, int w,
                 long long kcap, float ratio, float one_minus, int adain) {
    const int b = blockIdx.x / C, ch = blockIdx.x % C;
    const int n = h * w;
    const float *acc_f = reinterpret_cast<const float *>(accum);
    const float *cnt_f = acc_f + kcap * C;
    const long long *acc_q = reinterpret_cast<const long long *>(accum);
    const long long *cnt_q = acc_q + kcap * C;
    XT *xp = x + ((long long)b * C + ch) * n;
    const int *wb = winner + (long long)b * n;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int cell = threadIdx.x; cell < n; cell += K2_THREADS) {
        const int slot = wb[cell];
        const float xv = XIo<XT>::ld(xp + cell);
        float bv = xv;
        if (slot >= 0) {
            float m;
            if (DET) m = (float)(from_fix(acc_q[(long long)slot * C + ch]) / (double)cnt_q[slot]);
            else m = __fdiv_rn(acc_f[(long long)slot * C + ch], cnt_f[slot]);
            bv = __fadd_rn(__fmul_rn(one_minus, xv), __fmul_rn(ratio, m));
        }
        if (!adain) XIo<XT>::st(xp + cell, bv);
        else { const double dx = xv, db = bv; s0 += dx; s1 += dx * dx; s2 += db; s3 += db * db; }
    }
    if (!adain) return;
    __shared__ double part[K2_THREADS / 32][4];
    __shared__ float coef[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) { part[wid][0] = s0; part[wid][1] = s1; part[wid][2] = s2; part[wid][3] = s3; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[4] = {0, 0, 0, 0};
        for (int k = 0; k < K2_THREADS / 32; ++k)
            for (int j = 0; j < 4; ++j) s[j] += part[k][j];
        const double dn = (double)n;
        coef[0] = (float)(s[0] / dn);
        coef[1] = __fsqrt_rn(__fadd_rn((float)((s[1] - s[0] * s[0] / dn) / (dn - 1.0)), 1e-5f));
        coef[2] = __fsqrt_rn(__fadd_rn((float)((s[3] - s[2] * s[2] / dn) / (dn - 1.0)), 1e-5f));
        coef[3] = (float)(s[2] / dn);
    }
    __syncthreads();
    for (int cell = threadIdx.x; cell < n; cell += K2_THREADS) {
        const float v = XIo<XT>::ld(xp + cell);
        XIo<XT>::st(xp + cell, __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, coef[0]), coef[1]), coef[2]), coef[3]));
    }
}

// =================================================================================================================
// host side
// =================================================================================================================
static int grid_for(const void *kernel, int threads, long long work_items_per_block_hint, long long total_blocks_needed) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    long long resident = (long long)per_sm * srx_sm_count_cached();
    (void)work_items_per_block_hint;
    long long g = total_blocks_needed < resident ? total_blocks_needed : resident;
    return (int)(g < 1 ? 1 : g);
}

template <typename IdT>
static int run_scan(srx_plan *p, const void *ids, cudaStream_t st) {
    const srx_plan_desc &d = p->d;
    ScanOut *dout = nullptr;
    SRX_CUDA_CHECK(cudaMalloc(&dout, sizeof(ScanOut)));
    ScanOut init{0ull, LLONG_MAX, LLONG_MIN, 0, 0};
    SRX_CUDA_CHECK(cudaMemcpyAsync(dout, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    const long long npx = (long long)d.frames * d.height * d.width;
    const int grid = grid_for((const void *)k_scan_ids<IdT>, 256, 0, (npx + 255) / 256);
    k_scan_ids<IdT><<<grid, 256, 0, st>>>(reinterpret_cast<const IdT *>(ids), p->colcell, p->rowcell, d.height, d.width,
                                          d.lat_h, d.lat_w, npx, d.key_mode, dout);
    ScanOut res;
    cudaError_t e = cudaMemcpyAsync(&res, dout, sizeof(res), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dout);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "id scan failed: %s", cudaGetErrorString(e));
    p->n_valid = (int64_t)res.n_valid;
    p->key_min = res.n_valid ? res.key_min : 0;
    p->key_max = res.n_valid ? res.key_max : -1;
    if (res.bad_cell)
        return srx_set_error(SRX_ERR_INDEX, "index out of range: a valid id pixel maps outside the %dx%d latent "
                             "(x is divided by the id height and y by the id width, corrmap.py:239,249)", d.lat_h, d.lat_w);
    return SRX_OK;
}

extern "C" int srx_plan_create(srx_plan **out, const srx_plan_desc *desc, const void *ids_dev, void *stream) {
    SRX_REQUIRE(out && desc, SRX_ERR_INVALID, "null argument");
    *out = nullptr;
    const srx_plan_desc &d = *desc;
    SRX_REQUIRE(d.id_dtype == SRX_I32 || d.id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    SRX_REQUIRE(d.frames > 0 && d.height > 0 && d.width > 0 && d.batch > 0 && d.channels > 0 && d.lat_h > 0 && d.lat_w > 0,
                SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(d.channels <= 64, SRX_ERR_UNSUPPORTED, "more than 64 latent channels");
    SRX_REQUIRE(d.key_mode == SRX_KEY_VERTEX, SRX_ERR_UNSUPPORTED, "srx_plan_create handles SRX_KEY_VERTEX; tuple keys go through srx_legacy_*");
    SRX_REQUIRE(d.accum_mode == SRX_ACCUM_FAST || d.accum_mode == SRX_ACCUM_DETERMINISTIC || d.accum_mode == SRX_ACCUM_FAST_SPLIT,
                SRX_ERR_INVALID, "bad accum mode");
    SRX_REQUIRE(d.key_capacity > 0 || ids_dev, SRX_ERR_INVALID, "ids are required when key_capacity is 0 (scan)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    srx_plan *p = new (std::nothrow) srx_plan();
    SRX_REQUIRE(p, SRX_ERR_INVALID, "out of host memory");
    p->d = d;
    p->d.frame_map = nullptr;
    const bool allow_fused = d.accum_mode == SRX_ACCUM_FAST;
    if (d.accum_mode == SRX_ACCUM_FAST_SPLIT) p->d.accum_mode = SRX_ACCUM_FAST;

    // host tables in the reference's float32 arithmetic
    std::vector<int> tab((size_t)d.width + d.height + d.frames);
    int *colcell = tab.data(), *rowcell = colcell + d.width, *fmap = rowcell + d.height;
    bool r8 = (d.height == 8 * d.lat_h) && (d.width == 8 * d.lat_w) && d.channels == 4;
    for (int x = 0; x < d.width; ++x) {
        colcell[x] = host_cell(x, d.height, d.lat_w);  // x / H * w   (sic)
        if (colcell[x] != (x >> 3)) r8 = false;
    }
    for (int y = 0; y < d.height; ++y) {
        rowcell[y] = host_cell(y, d.width, d.lat_h);   // y / W * h   (sic)
        if (rowcell[y] != (y >> 3)) r8 = false;
    }
    std::vector<char> seen((size_t)d.batch, 0);
    for (int g = 0; g < d.frames; ++g) {
        int v = desc->frame_map ? desc->frame_map[g] : g;
        if (v < 0) v += d.batch;  // torch negative-index wrap
        if (v < 0 || v >= d.batch) {
            delete p;
            return srx_set_error(SRX_ERR_INDEX, "index %d is out of bounds for dimension 0 with size %d (frame index used as batch index, corresponder.py:314)",
                                 desc->frame_map ? desc->frame_map[g] : g, d.batch);
        }
        if (seen[v]) r8 = false;  // several id frames write one latent frame: ordered 64-bit winner path
        seen[v] = 1;
        fmap[g] = v;
    }
    p->fast_r8 = r8;
    cudaError_t e = cudaMalloc(&p->tables, tab.size() * sizeof(int));
    if (e != cudaSuccess) { delete p; return srx_set_error(SRX_ERR_CUDA, "cudaMalloc(tables): %s", cudaGetErrorString(e)); }
    p->colcell = p->tables; p->rowcell = p->colcell + d.width; p->fmap = p->rowcell + d.height;
    e = cudaMemcpyAsync(p->tables, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // tab is a stack-lifetime buffer
    if (e != cudaSuccess) { cudaFree(p->tables); delete p; return srx_set_error(SRX_ERR_CUDA, "table upload: %s", cudaGetErrorString(e)); }

    if (d.key_capacity > 0) {
        p->kcap = d.key_capacity;
    } else {
        int rc = d.id_dtype == SRX_I32 ? run_scan<int4>(p, ids_dev, st) : run_scan<short4>(p, ids_dev, st);
        if (rc != SRX_OK) { cudaFree(p->tables); delete p; return rc; }
        if (p->key_min < 0) {
            cudaFree(p->tables); delete p;
            return srx_set_error(SRX_ERR_KEY_RANGE, "negative vertex ids are not supported by the dense slot table");
        }
        p->kcap = p->key_max + 1 > 1 ? p->key_max + 1 : 1;
    }
    if (p->kcap > (1ll << SRX_SLOT_BITS)) {
        int64_t k = p->kcap;
        cudaFree(p->tables); delete p;
        return srx_set_error(SRX_ERR_KEY_RANGE, "key capacity %lld exceeds the dense slot table limit 2^25", (long long)k);
    }
    p->kcap = align_up(p->kcap, 64);
    p->fused = allow_fused && srx_fused_applicable(p);
    plan_layout(p);
    const int n = d.lat_h * d.lat_w;
    p->cluster = n >= 16384 ? 8 : (n >= 4096 ? 4 : (n >= 1024 ? 2 : 1));
    *out = p;
    return SRX_OK;
}

extern "C" int srx_plan_get_info(const srx_plan *p, srx_plan_info *info) {
    SRX_REQUIRE(p && info, SRX_ERR_INVALID, "null argument");
    info->n_valid = p->n_valid;
    info->key_min = p->key_min;
    info->key_max = p->key_max;
    info->key_capacity = p->kcap;
    info->workspace_bytes = p->total_bytes;
    info->accum_offset = p->accum_off;
    info->accum_bytes = p->accum_bytes;
    info->accum_dtype = p->d.accum_mode == SRX_ACCUM_DETERMINISTIC ? -64 : SRX_F32;
    info->fast_path = p->fast_r8 ? 1 : 0;
    info->fused = p->fused ? 1 : 0;
    info->need_offset = p->need_off;
    return SRX_OK;
}

extern "C" int srx_plan_bind_workspace(srx_plan *p, void *ws, int64_t bytes, void *stream) {
    SRX_REQUIRE(p && ws, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(bytes >= p->total_bytes, SRX_ERR_INVALID, "workspace too small: %lld < %lld", (long long)bytes, (long long)p->total_bytes);
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    p->ws = reinterpret_cast<char *>(ws);
    p->ws_bytes = bytes;
    // accumulator starts clean; latent frames that no id frame maps to keep winner = -1 forever
    // A0, A1, A2, pads, ctrl, exchange records (their flags must start below step 1)
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws, 0, (size_t)(p->ll_off + p->ll_bytes), st));
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->stats_off, 0, (size_t)p->stats_bytes, st));
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->winner_off, 0xFF, p->winner_bytes, st));
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->status_off, 0, 256, st));
    return SRX_OK;
}

extern "C" int srx_plan_destroy(srx_plan *p) {
    if (!p) return SRX_OK;
    if (p->tables) cudaFree(p->tables);
    delete p;
    return SRX_OK;
}

extern "C" int srx_plan_check(srx_plan *p, void *stream) {
    SRX_REQUIRE(p && p->ws, SRX_ERR_INVALID, "plan has no workspace");
    int st_host[4] = {0, 0, 0, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SRX_CUDA_CHECK(cudaMemcpyAsync(st_host, p->ws + p->status_off, sizeof(st_host), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (st_host[2])   // FZ_ST_TIMEOUT of the persistent step kernel (srx_fused.cu)
        return srx_set_error(SRX_ERR_PEER_LOST, "a wait inside the step kernel gave up after ~2 s: a peer rank (or a CTA group) never "
                             "arrived — ranks out of step, a dead peer, or unequal grids; the latents of that step are invalid");
    if (st_host[ST_CELL_RANGE])
        return srx_set_error(SRX_ERR_INDEX, "index out of range: a valid id pixel maps outside the latent");
    if (st_host[ST_KEY_RANGE])
        return srx_set_error(SRX_ERR_KEY_RANGE, "a vertex id fell outside the dense slot table (capacity %lld)", (long long)p->kcap);
    return SRX_OK;
}

// -----------------------------------------------------------------------------------------------------------------
// launches
// -----------------------------------------------------------------------------------------------------------------
// SRX_K1_DEBUG bit 0: skip the reductions (isolates the streaming + keying cost in experiments; results are wrong)
static int k1_debug_flags() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SRX_K1_DEBUG");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <typename IdT, typename XT, bool DET>
static int launch_accum(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    const srx_plan_desc &d = p->d;
    void *accum = p->ws + p->accum_off;
    int *winner = reinterpret_cast<int *>(p->ws + p->winner_off);
    int *status = reinterpret_cast<int *>(p->ws + p->status_off);
    const XT *x = reinterpret_cast<const XT *>(a->x_dev);
    const IdT *ids = reinterpret_cast<const IdT *>(a->ids_dev);
    if (p->fast_r8) {
        const int nrows = d.frames * d.lat_h;
        const int chunks = (d.lat_w + 8 * K1_U - 1) / (8 * K1_U);
        auto kern = k_accum_r8<IdT, XT, DET, false, false>;
        const int grid = grid_for((const void *)kern, 256, 0, (long long)nrows * chunks);
        kern<<<grid, 256, 0, st>>>(ids, nullptr, x, p->fmap, accum, winner, status, d.height, d.width, d.lat_h, d.lat_w,
                                   p->kcap, nrows, chunks, k1_debug_flags());
    } else {
        const long long npx = (long long)d.frames * d.height * d.width;
        const long long ncell_lat = (long long)d.batch * d.lat_h * d.lat_w;
        unsigned long long *w64 = reinterpret_cast<unsigned long long *>(p->ws + p->winner64_off);
        SRX_CUDA_CHECK(cudaMemsetAsync(w64, 0, p->winner64_bytes, st));
        auto kern = k_accum_generic<IdT, XT, DET>;
        const int grid = grid_for((const void *)kern, 256, 0, (npx + 255) / 256);
        kern<<<grid, 256, 0, st>>>(ids, x, p->fmap, p->colcell, p->rowcell, accum, w64, status, d.height, d.width,
                                   d.lat_h, d.lat_w, d.channels, p->kcap, npx);
        const int g2 = (int)((ncell_lat + 255) / 256 < 1184 ? (ncell_lat + 255) / 256 : 1184);
        k_winner_unpack<<<g2, 256, 0, st>>>(w64, winner, ncell_lat);
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

template <typename XT, bool DET>
static int launch_finalize(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    const srx_plan_desc &d = p->d;
    const void *accum = p->ws + p->accum_off;
    const int *winner = reinterpret_cast<const int *>(p->ws + p->winner_off);
    XT *x = reinterpret_cast<XT *>(a->x_dev);
    const float ratio = a->ratio;
    const float one_minus = (float)(1.0 - (double)a->ratio);  // python: (1 - ratio) in double, then cast to the tensor dtype
    if (d.channels == 4) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(d.batch * p->cluster));
        cfg.blockDim = dim3(K2_THREADS);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p->cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_finalize_c4<XT, DET>, x, accum, winner, d.lat_h, d.lat_w,
                                          (long long)p->kcap, ratio, one_minus, a->adain));
    } else {
        k_finalize_plane<XT, DET><<<d.batch * d.channels, K2_THREADS, 0, st>>>(x, accum, winner, d.channels, d.lat_h,
                                                                              d.lat_w, p->kcap, ratio, one_minus, a->adain);
        SRX_CUDA_CHECK(cudaGetLastError());
    }
    // leave the accumulator clean for the next step
    SRX_CUDA_CHECK(cudaMemsetAsync(p->ws + p->accum_off, 0, p->accum_bytes, st));
    return SRX_OK;
}

template <typename XT, bool DET>
static int dispatch_accum_ids(srx_plan *p, const srx_step_args *a, cudaStream_t st) {
    return p->d.id_dtype == SRX_I32 ? launch_accum<int4, XT, DET>(p, a, st) : launch_accum<short4, XT, DET>(p, a, st);
}

static int check_step(srx_plan *p, const srx_step_args *a) {
    SRX_REQUIRE(p && a, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(p->ws, SRX_ERR_INVALID, "srx_plan_bind_workspace was not called");
    SRX_REQUIRE(a->x_dev, SRX_ERR_INVALID, "null latents");
    SRX_REQUIRE(a->x_dtype == SRX_F32 || a->x_dtype == SRX_F16 || a->x_dtype == SRX_BF16, SRX_ERR_INVALID, "latent dtype must be f32/f16/bf16");
    return SRX_OK;
}

extern "C" int srx_accum_reduce(srx_plan *p, const srx_step_args *a, void *stream) {
    int rc = check_step(p, a);
    if (rc) return rc;
    SRX_REQUIRE(a->ids_dev, SRX_ERR_UNSUPPORTED, "cached-slot regime is not built yet: pass ids_dev");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool det = p->d.accum_mode == SRX_ACCUM_DETERMINISTIC;
    switch (a->x_dtype) {
        case SRX_F32: return det ? dispatch_accum_ids<float, true>(p, a, st) : dispatch_accum_ids<float, false>(p, a, st);
        case SRX_F16: return det ? dispatch_accum_ids<__half, true>(p, a, st) : dispatch_accum_ids<__half, false>(p, a, st);
        default: return det ? dispatch_accum_ids<__nv_bfloat16, true>(p, a, st) : dispatch_accum_ids<__nv_bfloat16, false>(p, a, st);
    }
}

extern "C" int srx_accum_finalize_gather(srx_plan *p, const srx_step_args *a, void *stream) {
    int rc = check_step(p, a);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool det = p->d.accum_mode == SRX_ACCUM_DETERMINISTIC;
    switch (a->x_dtype) {
        case SRX_F32: return det ? launch_finalize<float, true>(p, a, st) : launch_finalize<float, false>(p, a, st);
        case SRX_F16: return det ? launch_finalize<__half, true>(p, a, st) : launch_finalize<__half, false>(p, a, st);
        default: return det ? launch_finalize<__nv_bfloat16, true>(p, a, st) : launch_finalize<__nv_bfloat16, false>(p, a, st);
    }
}

extern "C" int srx_overlap_step(srx_plan *p, const srx_step_args *a, void *stream) {
    if (p && a && p->fused) {
        int rc0 = check_step(p, a);
        if (rc0) return rc0;
        return srx_launch_fused(p, a, reinterpret_cast<cudaStream_t>(stream));
    }
    SRX_REQUIRE(p && p->world == 1, SRX_ERR_UNSUPPORTED, "peer mode is only available with the persistent step kernel");
    int rc = srx_accum_reduce(p, a, stream);
    if (rc) return rc;
    return srx_accum_finalize_gather(p, a, stream);
}
