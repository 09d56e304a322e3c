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
//   3. one CTA walks the traces in order; the entries of a trace are processed in parallel (gather + diagonal pooling,
//      strategy weights, blend, write back), traces strictly one after another.
// ResizeOverlap's nearest up-sample / down-sample / where() (:205-221) wrap the sweep, as in the reference.
#include "srx_common.cuh"

#include <cub/device/device_radix_sort.cuh>

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
        if (g.resize) {
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

// ---- host ---------------------------------------------------------------------------------------------------------------
struct OrdLayout {
    int64_t rank, keys_out, vals_in, vals_out, offsets, work, s_lat, s_pool, s_w, cub, cub_bytes, tr, tr_bytes, total;
};
static inline int64_t ord_align(int64_t v) { return (v + 255) / 256 * 256; }

static int ord_layout(const srx_legacy_desc *d, OrdLayout *L) {
    const int64_t npx = (int64_t)d->frames * d->height * d->width;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned int *)nullptr, (unsigned int *)nullptr,
                                                    (const unsigned int *)nullptr, (unsigned int *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(e));
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
        k_ord_sweep<<<1, 256, 0, st>>>(work, vals_out, offsets, n_traces, a->view_normal_dev, reinterpret_cast<float *>(ws + L.s_lat),
                                       reinterpret_cast<float *>(ws + L.s_pool), reinterpret_cast<float *>(ws + L.s_w), g);
    }
    k_ord_downsample<XT><<<blocks((long long)d->frames * d->channels * d->lat_h * d->lat_w), 256, 0, st>>>(x, work, g);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
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
