// srx_legacy_ordered.cu — legacy overlap with kernel_radius > 0, in the reference's evaluation order.
//
// Overlap.__call__ (legacy_codes/stable_rendering_algo/overlap/overlap.py:83-152) walks the traces in dict order and
// writes every blended trace into the storage it keeps reading (`frame_seq_stack_copy = frame_seq_stack.detach()`, :97,:145).
// With kernel_radius == 0 a trace reads only its own pixels, so the order is irrelevant (srx_legacy.cu).  With a radius
// the pooled neighbours `(y+d, x+d), d in [-r, r]` (:61-80,:137-138) may already have been rewritten by earlier traces: the
// result is a Gauss-Seidel sweep in dict order (it differs from a read-original evaluation by 5.7e-2), so the
// order is part of the contract.  Here:
//   1. rank of every pixel's key among the traces (>= 2 entries) in insertion order  — srx_corrmap_trace_ranks
//   2. stable radix sort of the pixels by that rank (CUB, setup)                     — CSR of traces, entries in (frame,row,col) order
//   3. LEVEL SCHEDULE.  Trace t depends on an earlier trace s exactly when a pixel of one lies in the pooling neighbourhood
//      of the other (t must see what s wrote, and s must not see what t writes).  The longest-path level of every trace in
//      that conflict graph is found by parallel relaxation (k_ord_levels: level[t] = max over earlier neighbours + 1, to the
//      fixpoint); traces of one level are pairwise independent, so a level runs in parallel — a warp per trace, entries over
//      the lanes (gather + diagonal pooling, strategy weights, blend, write back) — and the levels run in order behind grid
//      barriers.  Identical to walking the traces one after another (the reference's order), far faster than the one-CTA
//      sweep it replaces (k_ord_sweep, kept behind SRX_ORD_SEQUENTIAL=1 as the cross-check).
// ResizeOverlap's nearest up-sample / down-sample / where() (:205-221) wrap the sweep, as in the reference.
#include "srx_common.cuh"

#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>

namespace cg = cooperative_groups;

struct OrdGeom {
    int T, H, W, h, w, C;
    int strategy, radius, resize;
    float up_sy, up_sx, down_sy, down_sx;
    float alpha, one_minus, inv_span;     // 1 / (2r + 1)
};

__device__ __forceinline__ int ord_nearest(int dst, float scale, int in_size) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}
__device__ __forceinline__ float ord_vn_weight(float vn) {  // algorithms.py:111-113
    return __fdiv_rn(1.f, __fadd_rn(fabsf(__fsub_rn(1.f, vn)), 1.f));
}

__global__ void __launch_bounds__(256) k_ord_iota(unsigned int *v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = (unsigned int)i;
}

// offsets[t] = first sorted position of trace t; offsets[n_traces] = number of entries (keys 0xFFFFFFFF = no trace sort last)
__global__ void __launch_bounds__(256) k_ord_offsets(const unsigned int *__restrict__ keys, long long n, int *__restrict__ offsets,
                                                      long long n_traces) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int k = i < n ? keys[i] : 0xFFFFFFFFu;
        const unsigned int prev = i > 0 ? keys[i - 1] : 0xFFFFFFFEu;   // differs from every valid key and from the sentinel
        if (i == 0 || k != prev) {
            if (k != 0xFFFFFFFFu) offsets[k] = (int)i;
            else if (i == 0 || prev != 0xFFFFFFFFu) offsets[n_traces] = (int)i;
        }
    }
}

template <typename XT>
__global__ void __launch_bounds__(256) k_ord_upsample(const XT *__restrict__ x, float *__restrict__ work, OrdGeom g) {
    const long long total = (long long)g.T * g.C * g.H * g.W;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(o % g.W);
        const long long t = o / g.W;
        const int py = (int)(t % g.H);
        const long long fc = t / g.H;
        const int uy = g.resize ? ord_nearest(py, g.up_sy, g.h) : py, ux = g.resize ? ord_nearest(px, g.up_sx, g.w) : px;
        work[o] = XIo<XT>::ld(x + (fc * g.h + uy) * g.w + ux);
    }
}

template <typename XT>
__global__ void __launch_bounds__(256) k_ord_downsample(XT *__restrict__ x, const float *__restrict__ work, OrdGeom g) {
    const long long total = (long long)g.T * g.C * g.h * g.w;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int sx = (int)(o % g.w);
        const long long t = o / g.w;
        const int sy = (int)(t % g.h);
        const long long fc = t / g.h;
        const int py = g.resize ? ord_nearest(sy, g.down_sy, g.H) : sy, px = g.resize ? ord_nearest(sx, g.down_sx, g.W) : sx;
        const float v = work[(fc * g.H + py) * g.W + px];
        if (g.resize == 1) {
            if (v != 0.f) XIo<XT>::st(x + o, v);        // torch.where(ovlp != 0, ovlp, original), overlap.py:221
        } else {
            XIo<XT>::st(x + o, v);
        }
    }
}

// The ordered sweep: ONE CTA, traces one after another.  scratch: [n_entries][C] latent, [n_entries][C] pooled, [n_entries] weight.
__global__ void __launch_bounds__(256) k_ord_sweep(float *__restrict__ work, const unsigned int *__restrict__ entries,
                                                    const int *__restrict__ offsets, long long n_traces,
                                                    const float *__restrict__ vnmap, float *__restrict__ s_lat,
                                                    float *__restrict__ s_pool, float *__restrict__ s_w, OrdGeom g) {
    __shared__ float s_tot[64];          // per-channel totals (average / view-normal), C <= 64 on this path
    const int C = g.C;
    const long long hw = (long long)g.H * g.W;
    for (long long t = 0; t < n_traces; ++t) {
        const int beg = offsets[t], end = offsets[t + 1], L = end - beg;
        // phase 1: gather + diagonal pooling (reads may see what earlier traces wrote)
        for (int idx = threadIdx.x; idx < L * C; idx += blockDim.x) {
            const int i = idx / C, c = idx - i * C;
            const unsigned int p = entries[beg + i];
            const int f = (int)(p / hw);
            const int r0 = (int)(p - (long long)f * hw);
            const int y = r0 / g.W, x = r0 - y * g.W;
            const float *plane = work + ((long long)f * C + c) * hw;
            const float lat = plane[(long long)y * g.W + x];
            float pooled = lat;
            if (g.radius > 0) {
                float acc = 0.f;
                for (int d = -g.radius; d <= g.radius; ++d) {
                    const int yy = min(max(y + d, 0), g.H - 1), xx = min(max(x + d, 0), g.W - 1);    // overlap.py:72-75
                    acc = __fadd_rn(acc, plane[(long long)yy * g.W + xx]);
                }
                pooled = __fmul_rn(acc, g.inv_span);
            }
            s_lat[(long long)(beg + i) * C + c] = lat;
            s_pool[(long long)(beg + i) * C + c] = pooled;
            if (c == 0 && g.strategy == SRX_STRATEGY_VIEW_NORMAL) s_w[beg + i] = ord_vn_weight(vnmap[p]);
        }
        __syncthreads();
        // phase 2a: totals that do not depend on the receiving entry
        if (g.strategy == SRX_STRATEGY_AVERAGE || g.strategy == SRX_STRATEGY_VIEW_NORMAL) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
            for (int c = warp; c < C; c += nwarp) {
                float acc = 0.f;
                for (int j = lane; j < L; j += 32) {
                    const float v = s_pool[(long long)(beg + j) * C + c];
                    acc += g.strategy == SRX_STRATEGY_VIEW_NORMAL ? __fmul_rn(s_w[beg + j], v) : v;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s_tot[c] = acc;
            }
            __syncthreads();
        }
        // phase 2b + 3: W @ X / W.sum(0) (algorithms.py:34-118), blend, write into the storage being read (overlap.py:145)
        for (int idx = threadIdx.x; idx < L * C; idx += blockDim.x) {
            const int i = idx / C, c = idx - i * C;
            const unsigned int p = entries[beg + i];
            const int f = (int)(p / hw);
            const int r0 = (int)(p - (long long)f * hw);
            const int y = r0 / g.W, x = r0 - y * g.W;
            float ov;
            if (g.strategy == SRX_STRATEGY_AVERAGE) {
                ov = __fdiv_rn(s_tot[c], (float)L);
            } else if (g.strategy == SRX_STRATEGY_VIEW_NORMAL) {
                ov = __fdiv_rn(s_tot[c], __fmul_rn((float)L, s_w[beg + i]));        // column sum of identical rows = L * w_i
            } else {
                float num = 0.f, den = 0.f;
                for (int j = 0; j < L; ++j) {
                    const unsigned int q = entries[beg + j];
                    const int fj = (int)(q / hw);
                    float wgt;
                    if (g.strategy == SRX_STRATEGY_FRAME_DISTANCE) {
                        wgt = __fdiv_rn(1.f, (float)(abs(f - fj) + 1));                               // algorithms.py:66-70
                    } else {
                        const int rj = (int)(q - (long long)fj * hw);
                        const int yj = rj / g.W, xj = rj - yj * g.W;
                        wgt = __fdiv_rn(1.f, (float)(abs(x - xj) + abs(y - yj) + 1));               // algorithms.py:87-93
                    }
                    num = __fadd_rn(num, __fmul_rn(wgt, s_pool[(long long)(beg + j) * C + c]));
                    den = __fadd_rn(den, wgt);
                }
                ov = __fdiv_rn(num, den);
            }
            const float lat = s_lat[(long long)(beg + i) * C + c];
            work[((long long)f * C + c) * hw + (long long)y * g.W + x] = __fadd_rn(__fmul_rn(g.alpha, ov), __fmul_rn(g.one_minus, lat));
        }
        __syncthreads();
    }
}

// ---- level schedule --------------------------------------------------------------------------------------------------------
// level[t] = 1 + max level of the earlier traces that own a pixel in t's pooling neighbourhood (0 without such a trace): the
// longest path in the conflict graph, by relaxation to the fixpoint.  Cooperative launch; flag[3] rotates (a round sets its
// own flag, clears the next one, and everybody reads it behind the grid barrier).
__global__ void __launch_bounds__(256) k_ord_levels(const int *__restrict__ rank, const unsigned int *__restrict__ entries,
                                                     const int *__restrict__ offsets, long long n_traces, OrdGeom g,
                                                     int *level, int *flag, int *max_level) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long hw = (long long)g.H * g.W;
    for (int round = 0;; ++round) {
        if (blockIdx.x == 0 && threadIdx.x == 0) flag[(round + 1) % 3] = 0;
        for (long long t = warp; t < n_traces; t += nwarps) {
            const int beg = offsets[t], end = offsets[t + 1];
            const int cur = *reinterpret_cast<volatile int *>(level + t);
            int cand = cur;
            for (int i = beg + lane; i < end; i += 32) {
                const unsigned int p = entries[i];
                const int f = (int)(p / hw);
                const int r0 = (int)(p - (long long)f * hw);
                const int y = r0 / g.W, x = r0 - y * g.W;
                for (int d = -g.radius; d <= g.radius; ++d) {
                    if (d == 0) continue;
                    const int yy = min(max(y + d, 0), g.H - 1), xx = min(max(x + d, 0), g.W - 1);
                    const int s = rank[(long long)f * hw + (long long)yy * g.W + xx];
                    if (s >= 0 && s < t) cand = max(cand, *reinterpret_cast<volatile int *>(level + s) + 1);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
            if (lane == 0 && cand > cur) {
                *reinterpret_cast<volatile int *>(level + t) = cand;
                *reinterpret_cast<volatile int *>(flag + round % 3) = 1;
            }
        }
        __threadfence();
        grid.sync();
        if (*reinterpret_cast<volatile int *>(flag + round % 3) == 0) break;
    }
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_traces; t += (long long)gridDim.x * blockDim.x)
        atomicMax(max_level, level[t]);
}

__global__ void __launch_bounds__(256) k_ord_level_hist(const int *__restrict__ level, long long n_traces, int *__restrict__ hist) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_traces; t += (long long)gridDim.x * blockDim.x)
        atomicAdd(hist + level[t], 1);
}
__global__ void __launch_bounds__(256) k_ord_level_scatter(const int *__restrict__ level, long long n_traces, int *__restrict__ cursor,
                                                            int *__restrict__ order) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_traces; t += (long long)gridDim.x * blockDim.x)
        order[atomicAdd(cursor + level[t], 1)] = (int)t;
}

// One trace on one warp: the body of k_ord_sweep with the entries over the lanes.
__device__ __forceinline__ void ord_trace_warp(float *__restrict__ work, const unsigned int *__restrict__ entries, int beg, int L,
                                               const float *__restrict__ vnmap, float *__restrict__ s_lat, float *__restrict__ s_pool,
                                               float *__restrict__ s_w, float *tot, const OrdGeom &g, int lane) {
    const int C = g.C;
    const long long hw = (long long)g.H * g.W;
    for (int idx = lane; idx < L * C; idx += 32) {
        const int i = idx / C, c = idx - i * C;
        const unsigned int p = entries[beg + i];
        const int f = (int)(p / hw);
        const int r0 = (int)(p - (long long)f * hw);
        const int y = r0 / g.W, x = r0 - y * g.W;
        const float *plane = work + ((long long)f * C + c) * hw;
        const float lat = __ldcg(plane + (long long)y * g.W + x);
        float pooled = lat;
        if (g.radius > 0) {
            float acc = 0.f;
            for (int d = -g.radius; d <= g.radius; ++d) {
                const int yy = min(max(y + d, 0), g.H - 1), xx = min(max(x + d, 0), g.W - 1);    // overlap.py:72-75
                acc = __fadd_rn(acc, __ldcg(plane + (long long)yy * g.W + xx));
            }
            pooled = __fmul_rn(acc, g.inv_span);
        }
        s_lat[(long long)(beg + i) * C + c] = lat;
        s_pool[(long long)(beg + i) * C + c] = pooled;
        if (c == 0 && g.strategy == SRX_STRATEGY_VIEW_NORMAL) s_w[beg + i] = ord_vn_weight(vnmap[p]);
    }
    __syncwarp();
    if (g.strategy == SRX_STRATEGY_AVERAGE || g.strategy == SRX_STRATEGY_VIEW_NORMAL) {
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            for (int j = lane; j < L; j += 32) {
                const float v = s_pool[(long long)(beg + j) * C + c];
                acc += g.strategy == SRX_STRATEGY_VIEW_NORMAL ? __fmul_rn(s_w[beg + j], v) : v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) tot[c] = acc;
        }
        __syncwarp();
    }
    for (int idx = lane; idx < L * C; idx += 32) {
        const int i = idx / C, c = idx - i * C;
        const unsigned int p = entries[beg + i];
        const int f = (int)(p / hw);
        const int r0 = (int)(p - (long long)f * hw);
        const int y = r0 / g.W, x = r0 - y * g.W;
        float ov;
        if (g.strategy == SRX_STRATEGY_AVERAGE) {
            ov = __fdiv_rn(tot[c], (float)L);
        } else if (g.strategy == SRX_STRATEGY_VIEW_NORMAL) {
            ov = __fdiv_rn(tot[c], __fmul_rn((float)L, s_w[beg + i]));
        } else {
            float num = 0.f, den = 0.f;
            for (int j = 0; j < L; ++j) {
                const unsigned int q = entries[beg + j];
                const int fj = (int)(q / hw);
                float wgt;
                if (g.strategy == SRX_STRATEGY_FRAME_DISTANCE) {
                    wgt = __fdiv_rn(1.f, (float)(abs(f - fj) + 1));
                } else {
                    const int rj = (int)(q - (long long)fj * hw);
                    const int yj = rj / g.W, xj = rj - yj * g.W;
                    wgt = __fdiv_rn(1.f, (float)(abs(x - xj) + abs(y - yj) + 1));
                }
                num = __fadd_rn(num, __fmul_rn(wgt, s_pool[(long long)(beg + j) * C + c]));
                den = __fadd_rn(den, wgt);
            }
            ov = __fdiv_rn(num, den);
        }
        const float lat = s_lat[(long long)(beg + i) * C + c];
        __stcg(work + ((long long)f * C + c) * hw + (long long)y * g.W + x, __fadd_rn(__fmul_rn(g.alpha, ov), __fmul_rn(g.one_minus, lat)));
    }
}

// All levels in one cooperative launch: the traces of a level over the warps of the grid, a grid barrier between levels.
__global__ void __launch_bounds__(256) k_ord_sweep_levels(float *__restrict__ work, const unsigned int *__restrict__ entries,
                                                           const int *__restrict__ offsets, const int *__restrict__ order,
                                                           const int *__restrict__ level_start, const int *__restrict__ max_level,
                                                           const float *__restrict__ vnmap, float *__restrict__ s_lat,
                                                           float *__restrict__ s_pool, float *__restrict__ s_w, OrdGeom g) {
    __shared__ float s_tot[8][64];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int n_levels = *max_level + 1;
    for (int lv = 0; lv < n_levels; ++lv) {
        const int lo = level_start[lv], hi = level_start[lv + 1];
        for (long long k = lo + warp; k < hi; k += nwarps) {
            const int t = order[k];
            const int beg = offsets[t];
            ord_trace_warp(work, entries, beg, offsets[t + 1] - beg, vnmap, s_lat, s_pool, s_w, s_tot[wib], g, lane);
        }
        __threadfence();
        grid.sync();
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
struct OrdLayout {
    int64_t rank, keys_out, vals_in, vals_out, offsets, work, s_lat, s_pool, s_w, cub, cub_bytes, tr, tr_bytes, level, hist, order, misc, total;
};
static inline int64_t ord_align(int64_t v) { return (v + 255) / 256 * 256; }

static int ord_layout(const srx_legacy_desc *d, OrdLayout *L) {
    const int64_t npx = (int64_t)d->frames * d->height * d->width;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned int *)nullptr, (unsigned int *)nullptr,
                                                    (const unsigned int *)nullptr, (unsigned int *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(e));
    size_t scan_bytes = 0;
    e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int *)nullptr, (int *)nullptr, (int)(npx / 2 + 4));
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceScan size query failed: %s", cudaGetErrorString(e));
    if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
    int64_t off = 0;
    L->rank = off; off = ord_align(off + npx * 4);
    L->keys_out = off; off = ord_align(off + npx * 4);
    L->vals_in = off; off = ord_align(off + npx * 4);
    L->vals_out = off; off = ord_align(off + npx * 4);
    L->offsets = off; off = ord_align(off + (npx / 2 + 2) * 4);
    L->work = off; off = ord_align(off + npx * d->channels * 4);
    L->s_lat = off; off = ord_align(off + npx * d->channels * 4);
    L->s_pool = off; off = ord_align(off + npx * d->channels * 4);
    L->s_w = off; off = ord_align(off + npx * 4);
    L->cub = off; L->cub_bytes = (int64_t)cub_bytes; off = ord_align(off + (int64_t)cub_bytes);
    L->tr_bytes = srx_corrmap_trace_ranks_workspace_bytes(npx);
    if (L->tr_bytes < 0) return srx_set_error(SRX_ERR_UNSUPPORTED, "too many pixels");
    L->tr = off; off = ord_align(off + L->tr_bytes);
    L->level = off; off = ord_align(off + (npx / 2 + 2) * 4);      // per trace
    L->hist = off; off = ord_align(off + (npx / 2 + 4) * 4);       // per level: counts, then (scanned) first trace of the level
    L->order = off; off = ord_align(off + (npx / 2 + 2) * 4);      // traces grouped by level
    L->misc = off; off += 256;                                       // relaxation flags [3], deepest level
    L->total = off;
    return SRX_OK;
}

static int ord_validate(const srx_legacy_desc *d) {
    SRX_REQUIRE(d, SRX_ERR_INVALID, "null descriptor");
    SRX_REQUIRE(d->id_dtype == SRX_I32 || d->id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    SRX_REQUIRE(d->frames > 0 && d->height > 0 && d->width > 0 && d->channels > 0 && d->lat_h > 0 && d->lat_w > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(d->channels <= 64, SRX_ERR_UNSUPPORTED, "the ordered sweep supports up to 64 channels (B*C)");
    SRX_REQUIRE(d->strategy >= SRX_STRATEGY_AVERAGE && d->strategy <= SRX_STRATEGY_VIEW_NORMAL, SRX_ERR_INVALID, "Unknown algorithm %d", d->strategy);
    SRX_REQUIRE((int64_t)d->frames * d->height * d->width < (1ll << 31), SRX_ERR_UNSUPPORTED, "too many pixels for 32-bit entries");
    return SRX_OK;
}

extern "C" int64_t srx_legacy_ordered_workspace_bytes(const srx_legacy_desc *d) {
    if (ord_validate(d) != SRX_OK) return -1;
    OrdLayout L;
    if (ord_layout(d, &L) != SRX_OK) return -1;
    return L.total;
}

template <typename XT>
static int ord_impl(const srx_legacy_desc *d, const srx_legacy_args *a, int radius, cudaStream_t st) {
    OrdLayout L;
    int rc = ord_layout(d, &L);
    if (rc) return rc;
    SRX_REQUIRE(a->workspace_bytes >= L.total, SRX_ERR_INVALID, "workspace too small");
    char *ws = reinterpret_cast<char *>(a->workspace_dev);
    const long long npx = (long long)d->frames * d->height * d->width;
    int *rank = reinterpret_cast<int *>(ws + L.rank);
    unsigned int *keys_out = reinterpret_cast<unsigned int *>(ws + L.keys_out);
    unsigned int *vals_in = reinterpret_cast<unsigned int *>(ws + L.vals_in);
    unsigned int *vals_out = reinterpret_cast<unsigned int *>(ws + L.vals_out);
    int *offsets = reinterpret_cast<int *>(ws + L.offsets);
    float *work = reinterpret_cast<float *>(ws + L.work);

    int64_t n_traces = 0;
    rc = srx_corrmap_trace_ranks(a->ids_dev, d->id_dtype, d->frames, d->height, d->width, d->merge_len, rank, &n_traces, ws + L.tr,
                                 L.tr_bytes, st);
    if (rc) return rc;

    OrdGeom g;
    g.T = d->frames; g.H = d->height; g.W = d->width; g.h = d->lat_h; g.w = d->lat_w; g.C = d->channels;
    g.strategy = d->strategy; g.radius = radius;
    g.resize = (d->lat_h != d->height || d->lat_w != d->width) ? 1 : 0;
    volatile float usy = (float)d->lat_h / (float)d->height, usx = (float)d->lat_w / (float)d->width;
    volatile float dsy = (float)d->height / (float)d->lat_h, dsx = (float)d->width / (float)d->lat_w;
    g.up_sy = usy; g.up_sx = usx; g.down_sy = dsy; g.down_sx = dsx;
    g.alpha = a->alpha; g.one_minus = (float)(1.0 - (double)a->alpha);
    volatile float inv = 1.0f / (float)(2 * radius + 1);
    g.inv_span = inv;
    const int sms = srx_sm_count_cached();
    auto blocks = [&](long long n) { long long nb = (n + 255) / 256; return (int)(nb < (long long)sms * 8 ? (nb < 1 ? 1 : nb) : (long long)sms * 8); };
    XT *x = reinterpret_cast<XT *>(a->x_dev);
    k_ord_upsample<XT><<<blocks(npx * d->channels), 256, 0, st>>>(x, work, g);
    if (n_traces > 0) {
        k_ord_iota<<<blocks(npx), 256, 0, st>>>(vals_in, npx);
        size_t cub_bytes = (size_t)L.cub_bytes;
        // keys: ranks 0..n_traces-1 and the 0xFFFFFFFF sentinel of pixels outside every trace — all 32 bits are sorted
        SRX_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, reinterpret_cast<const unsigned int *>(rank), keys_out,
                                                       vals_in, vals_out, (int)npx, 0, 32, st));
        k_ord_offsets<<<blocks(npx + 1), 256, 0, st>>>(keys_out, npx, offsets, n_traces);
        const char *seq_env = getenv("SRX_ORD_SEQUENTIAL");       // read per call: tests switch it to cross-check the two sweeps
        const int sequential = (seq_env && seq_env[0] && seq_env[0] != '0') ? 1 : 0;
        float *s_lat = reinterpret_cast<float *>(ws + L.s_lat), *s_pool = reinterpret_cast<float *>(ws + L.s_pool), *s_w = reinterpret_cast<float *>(ws + L.s_w);
        if (sequential || radius == 0) {
            // (radius 0: no trace reads another trace's pixels — one level; the plain kernels of srx_legacy.cu serve that case)
            k_ord_sweep<<<1, 256, 0, st>>>(work, vals_out, offsets, n_traces, a->view_normal_dev, s_lat, s_pool, s_w, g);
        } else {
            int *level = reinterpret_cast<int *>(ws + L.level), *hist = reinterpret_cast<int *>(ws + L.hist);
            int *order = reinterpret_cast<int *>(ws + L.order), *misc = reinterpret_cast<int *>(ws + L.misc);
            const int *rank_c = rank;
            const unsigned int *entries_c = vals_out;
            const int *offsets_c = offsets;
            SRX_CUDA_CHECK(cudaMemsetAsync(level, 0, (size_t)(n_traces + 1) * 4, st));
            SRX_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t)(n_traces + 4) * 4, st));
            SRX_CUDA_CHECK(cudaMemsetAsync(misc, 0, 256, st));
            int per_sm = 0, per_sm2 = 0;
            SRX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ord_levels, 256, 0));
            SRX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_ord_sweep_levels, 256, 0));
            const int grid1 = sms * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm)), grid2 = sms * (per_sm2 < 1 ? 1 : (per_sm2 > 4 ? 4 : per_sm2));
            long long nt = n_traces;
            int *flag = misc, *max_level = misc + 4;
            void *args1[] = {(void *)&rank_c, (void *)&entries_c, (void *)&offsets_c, (void *)&nt, (void *)&g, (void *)&level, (void *)&flag, (void *)&max_level};
            SRX_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_ord_levels, dim3(grid1), dim3(256), args1, 0, st));
            // traces grouped by level: histogram, exclusive scan (level l starts at hist[l]), scatter through a cursor copy
            k_ord_level_hist<<<blocks(n_traces), 256, 0, st>>>(level, n_traces, hist);
            size_t scan_bytes = (size_t)L.cub_bytes;
            SRX_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws + L.cub, scan_bytes, hist, hist, (int)(n_traces + 2), st));
            int *cursor = reinterpret_cast<int *>(ws + L.keys_out);          // the sort's key buffer is free again
            SRX_CUDA_CHECK(cudaMemcpyAsync(cursor, hist, (size_t)(n_traces + 2) * 4, cudaMemcpyDeviceToDevice, st));
            k_ord_level_scatter<<<blocks(n_traces), 256, 0, st>>>(level, n_traces, cursor, order);
            const int *order_c = order, *hist_c = hist, *max_c = max_level;
            const float *vn = a->view_normal_dev;
            void *args2[] = {(void *)&work, (void *)&entries_c, (void *)&offsets_c, (void *)&order_c, (void *)&hist_c, (void *)&max_c, (void *)&vn,
                             (void *)&s_lat, (void *)&s_pool, (void *)&s_w, (void *)&g};
            SRX_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_ord_sweep_levels, dim3(grid2), dim3(256), args2, 0, st));
        }
    }
    k_ord_downsample<XT><<<blocks((long long)d->frames * d->channels * d->lat_h * d->lat_w), 256, 0, st>>>(x, work, g);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// johnny_overlap (legacy_codes/legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-141): the experimental
// frame-distance variant of the diffusers pipeline.  Per trace, entry after entry in (frame,row,col) order:
//     value = sum_j x_j / (|t_i - t_j| + 1),  count = sum_j 1 / (|t_i - t_j| + 1)          (:95-107)
//     x_i  <- alpha * value / count + (1 - alpha) * x_i                                      (:109-111)
//     x_i  <- beta * base + (1 - beta) * x_i, base = the noised original latent at the trace's FIRST entry   (:112-116)
// written into the storage the next entry reads (:118) — a Gauss-Seidel recurrence INSIDE the trace; traces touch disjoint
// pixels, so they run in parallel (a warp each), lanes over j.  Nearest up-sample before, nearest down-sample after (no
// `where`: :132).  gamma (extra noise, :139-143) is hard-wired to 0 in the reference (:39) and not implemented.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_johnny_sweep(float *__restrict__ work, const unsigned int *__restrict__ entries,
                                                       const int *__restrict__ offsets, long long n_traces,
                                                       const float *__restrict__ base, OrdGeom g, float beta, float one_minus_beta) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long hw = (long long)g.H * g.W;
    const int C = g.C;
    for (long long t = warp; t < n_traces; t += nwarps) {
        const int beg = offsets[t], end = offsets[t + 1], L = end - beg;
        // the trace's first entry names the base colour (the dict's first appearance, :113-115)
        const unsigned int p0 = entries[beg];
        const int f0 = (int)(p0 / hw);
        const int r00 = (int)(p0 - (long long)f0 * hw);
        const int by = g.resize ? ord_nearest(r00 / g.W, g.up_sy, g.h) : r00 / g.W, bx = g.resize ? ord_nearest(r00 % g.W, g.up_sx, g.w) : r00 % g.W;
        for (int i = 0; i < L; ++i) {
            const unsigned int pi = entries[beg + i];
            const int fi = (int)(pi / hw);
            const long long ri = pi - (long long)fi * hw;
            float cnt = 0.f;
            for (int j = lane; j < L; j += 32) {
                const int fj = (int)(entries[beg + j] / hw);
                cnt += __fdiv_rn(1.f, (float)(abs(fi - fj) + 1));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            for (int c = 0; c < C; ++c) {
                float acc = 0.f;
                for (int j = lane; j < L; j += 32) {
                    const unsigned int pj = entries[beg + j];
                    const int fj = (int)(pj / hw);
                    const float v = *reinterpret_cast<volatile float *>(work + ((long long)fj * C + c) * hw + (pj - (long long)fj * hw));
                    acc += __fmul_rn(v, __fdiv_rn(1.f, (float)(abs(fi - fj) + 1)));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) {
                    float *dst = work + ((long long)fi * C + c) * hw + ri;
                    float ov = __fadd_rn(__fmul_rn(g.alpha, __fdiv_rn(acc, cnt)), __fmul_rn(g.one_minus, *dst));
                    if (base) ov = __fadd_rn(__fmul_rn(beta, base[(((long long)f0 * C + c) * g.h + by) * g.w + bx]), __fmul_rn(one_minus_beta, ov));
                    *reinterpret_cast<volatile float *>(dst) = ov;
                }
            }
            __syncwarp();      // entry i is in place before entry i + 1 reads the trace
        }
    }
}

template <typename XT>
static int johnny_impl(const srx_legacy_desc *d, const srx_legacy_args *a, float beta, const float *base, cudaStream_t st) {
    OrdLayout L;
    int rc = ord_layout(d, &L);
    if (rc) return rc;
    SRX_REQUIRE(a->workspace_bytes >= L.total, SRX_ERR_INVALID, "workspace too small");
    char *ws = reinterpret_cast<char *>(a->workspace_dev);
    const long long npx = (long long)d->frames * d->height * d->width;
    int *rank = reinterpret_cast<int *>(ws + L.rank);
    unsigned int *keys_out = reinterpret_cast<unsigned int *>(ws + L.keys_out);
    unsigned int *vals_in = reinterpret_cast<unsigned int *>(ws + L.vals_in);
    unsigned int *vals_out = reinterpret_cast<unsigned int *>(ws + L.vals_out);
    int *offsets = reinterpret_cast<int *>(ws + L.offsets);
    float *work = reinterpret_cast<float *>(ws + L.work);
    int64_t n_traces = 0;
    rc = srx_corrmap_trace_ranks(a->ids_dev, d->id_dtype, d->frames, d->height, d->width, d->merge_len, rank, &n_traces, ws + L.tr,
                                 L.tr_bytes, st);
    if (rc) return rc;
    OrdGeom g;
    g.T = d->frames; g.H = d->height; g.W = d->width; g.h = d->lat_h; g.w = d->lat_w; g.C = d->channels;
    g.strategy = SRX_STRATEGY_FRAME_DISTANCE; g.radius = 0;
    g.resize = (d->lat_h != d->height || d->lat_w != d->width) ? 1 : 0;
    volatile float usy = (float)d->lat_h / (float)d->height, usx = (float)d->lat_w / (float)d->width;
    volatile float dsy = (float)d->height / (float)d->lat_h, dsx = (float)d->width / (float)d->lat_w;
    g.up_sy = usy; g.up_sx = usx; g.down_sy = dsy; g.down_sx = dsx;
    g.alpha = a->alpha; g.one_minus = (float)(1.0 - (double)a->alpha);
    g.inv_span = 1.f;
    const int sms = srx_sm_count_cached();
    auto blocks = [&](long long n) { long long nb = (n + 255) / 256; return (int)(nb < (long long)sms * 8 ? (nb < 1 ? 1 : nb) : (long long)sms * 8); };
    XT *x = reinterpret_cast<XT *>(a->x_dev);
    k_ord_upsample<XT><<<blocks(npx * d->channels), 256, 0, st>>>(x, work, g);
    if (n_traces > 0) {
        k_ord_iota<<<blocks(npx), 256, 0, st>>>(vals_in, npx);
        size_t cub_bytes = (size_t)L.cub_bytes;
        SRX_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, reinterpret_cast<const unsigned int *>(rank), keys_out,
                                                       vals_in, vals_out, (int)npx, 0, 32, st));
        k_ord_offsets<<<blocks(npx + 1), 256, 0, st>>>(keys_out, npx, offsets, n_traces);
        k_johnny_sweep<<<blocks(n_traces * 32), 256, 0, st>>>(work, vals_out, offsets, n_traces, base, g, beta, (float)(1.0 - (double)beta));
    }
    g.resize = g.resize ? 2 : 0;      // plain nearest down-sample, no where(): johnny_overlap.py:132
    k_ord_downsample<XT><<<blocks((long long)d->frames * d->channels * d->lat_h * d->lat_w), 256, 0, st>>>(x, work, g);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

// base_dev: [T, B*C, h, w] float32 base-colour latents (the noised originals, johnny_overlap.py:63-65) or NULL / beta == 0
extern "C" int srx_johnny_overlap(const srx_legacy_desc *d, const srx_legacy_args *a, float beta, const float *base_dev, void *stream) {
    int rc = ord_validate(d);
    if (rc) return rc;
    SRX_REQUIRE(a && a->x_dev && a->ids_dev && a->workspace_dev, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace_dev) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    SRX_REQUIRE(beta >= 0.f && (beta == 0.f || base_dev), SRX_ERR_INVALID, "beta > 0 needs the base-colour latents");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float *base = beta > 0.f ? base_dev : nullptr;
    switch (a->x_dtype) {
        case SRX_F32: return johnny_impl<float>(d, a, beta, base, st);
        case SRX_F16: return johnny_impl<__half>(d, a, beta, base, st);
        case SRX_BF16: return johnny_impl<__nv_bfloat16>(d, a, beta, base, st);
        default: return srx_set_error(SRX_ERR_INVALID, "latent dtype must be f32/f16/bf16");
    }
}

extern "C" int srx_legacy_overlap_ordered(const srx_legacy_desc *d, const srx_legacy_args *a, int kernel_radius, void *stream) {
    int rc = ord_validate(d);
    if (rc) return rc;
    SRX_REQUIRE(a && a->x_dev && a->ids_dev && a->workspace_dev, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE(kernel_radius >= 0, SRX_ERR_INVALID, "negative kernel radius");
    SRX_REQUIRE(d->strategy != SRX_STRATEGY_VIEW_NORMAL || a->view_normal_dev, SRX_ERR_INVALID,
                "perpendicular_view_normal needs view_normal_map (algorithms.py:106)");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace_dev) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a->x_dtype) {
        case SRX_F32: return ord_impl<float>(d, a, kernel_radius, st);
        case SRX_F16: return ord_impl<__half>(d, a, kernel_radius, st);
        case SRX_BF16: return ord_impl<__nv_bfloat16>(d, a, kernel_radius, st);
        default: return srx_set_error(SRX_ERR_INVALID, "latent dtype must be f32/f16/bf16");
    }
}
