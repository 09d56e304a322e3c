// srx_legacy.cu — legacy-generation overlap: full-tuple keys, four weighting strategies.
//
// Replaces (reference paths relative to /root/reference/legacy_codes/stable_rendering_algo):
//   CorrespondenceMap.FromExisting / merge_nearby   data_classes/correspondence_map.py:145-170, 276-286
//   Overlap.__call__ (kernel_radius 0)              overlap/overlap.py:83-152
//   ResizeOverlap.__call__                          overlap/overlap.py:180-222
//   AverageDistance / FrameDistance / PixelDistance / PerpendicularViewNormal   overlap/algorithms.py:34-118
//
// The reference up-samples the latents to the id resolution, walks a Python dict of per-key traces doing an [L,L]
// weight matmul for each, and down-samples again.  Only ONE id pixel per latent cell survives the nearest
// down-sample, so the work is restated gather-side:
//   L0 seed     : every latent cell looks at the id pixel it will be sampled from, packs its id tuple into an exact
//                 64-bit key and inserts it into an L2-resident open-addressing table (atomicCAS); cells sharing a
//                 key are chained.  Keys that no sampled pixel carries are never stored.
//   L1 accum    : one streaming pass over all id pixels (coalesced 128-bit loads).  A pixel whose key is in the
//                 table adds weight(i,j) * latent_j into the record of every chained cell i (vector atomics) —
//                 the row of the reference's weight matrix that belongs to cell i, without materialising it.
//   L2 finalize : ov = sum / weight-sum (or sum / (L * w_i) for the view-normal strategy, algorithms.py:114-117),
//                 alpha blend, `where(out != 0, out, original)` (overlap.py:221), in-place write.
#include "srx_common.cuh"
#include <vector>

#define LG_EMPTY 0xFFFFFFFFFFFFFFFFull
enum { LG_ST_KEY = 0 };

struct LegacyGeom {
    int T, H, W, h, w, C;
    int merge;      // >= 1
    int strategy;
    int resize;     // 1: ResizeOverlap (apply where()), 0: Overlap at id resolution
    unsigned int mask;  // table capacity - 1
    float up_sy, up_sx;      // fl32(h / H), fl32(w / W): id pixel -> latent cell (F.interpolate up-sample, overlap.py:205-206)
    float down_sy, down_sx;  // fl32(H / h), fl32(W / w): latent cell -> sampled id pixel (down-sample, overlap.py:217-218)
};

// F.interpolate(mode='nearest') source index: min(floor(dst * fl32(in/out)), in-1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

__device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

// exact 64-bit packing of the id tuple (after merge_nearby's floor division of components 2 and 3)
template <typename IdT> __device__ __forceinline__ bool pack_key(const IdPx &p, int merge, unsigned long long *key);
template <> __device__ __forceinline__ bool pack_key<short4>(const IdPx &p, int merge, unsigned long long *key) {
    const int a = floordiv(p.i, merge), b = floordiv(p.v, merge);
    *key = ((unsigned long long)(unsigned short)p.s << 48) | ((unsigned long long)(unsigned short)p.m << 32) |
           ((unsigned long long)(unsigned short)a << 16) | (unsigned long long)(unsigned short)b;
    return true;  // 4 x 16 bits: always exact
}
template <> __device__ __forceinline__ bool pack_key<int4>(const IdPx &p, int merge, unsigned long long *key) {
    const int a = floordiv(p.i, merge), b = floordiv(p.v, merge);
    if (p.s < 0 || p.s >= 1024 || p.m < 0 || p.m >= 1024 || a < 0 || a >= 4096 || b < 0) return false;
    *key = ((unsigned long long)p.s << 54) | ((unsigned long long)p.m << 44) | ((unsigned long long)a << 32) |
           (unsigned long long)(unsigned int)b;
    return true;
}

__device__ __forceinline__ unsigned int hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned int)k;
}

__device__ __forceinline__ float vn_weight(float vn) {  // algorithms.py:111-113
    return __fdiv_rn(1.f, __fadd_rn(fabsf(__fsub_rn(1.f, vn)), 1.f));
}

struct LegacyBufs {
    unsigned long long *keys;  // [cap]
    int *head;                 // [cap] first chained cell, -1 = none
    int *next;                 // [ncell]
    int *cellslot;             // [ncell] table slot of the cell's key, -1 = the sampled pixel has no id
    float *selfx;              // [ncell][C] latent seen through the resize round trip
    float *sum;                // [ncell][C]
    float *wsum;               // [ncell]
    float *cnt;                // [ncell]  L
    float *wself;              // [ncell]  w_i (view-normal)
    int *status;
};

template <typename IdT, typename XT>
__global__ void __launch_bounds__(256) k_legacy_seed(const IdT *__restrict__ ids, const XT *__restrict__ x,
                                                      const float *__restrict__ vnmap, LegacyBufs b, LegacyGeom g, long long ncell) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int sx = (int)(cell % g.w);
        const long long t = cell / g.w;
        const int sy = (int)(t % g.h);
        const int f = (int)(t / g.h);
        const int py = nearest_src(sy, g.down_sy, g.H), px = nearest_src(sx, g.down_sx, g.W);
        const long long pix = ((long long)f * g.H + py) * g.W + px;
        const int uy = nearest_src(py, g.up_sy, g.h), ux = nearest_src(px, g.up_sx, g.w);
        for (int c = 0; c < g.C; ++c)
            b.selfx[cell * g.C + c] = XIo<XT>::ld(x + (((long long)f * g.C + c) * g.h + uy) * g.w + ux);
        const IdPx p = load_id(ids + pix);
        int slot = -1;
        if ((p.s | p.m | p.i | p.v) != 0) {  // correspondence_map.py:153 — only all-zero ids are skipped
            unsigned long long key;
            if (!pack_key<IdT>(p, g.merge, &key)) {
                atomicOr(b.status + LG_ST_KEY, 1);
            } else {
                unsigned int s = hash64(key) & g.mask;
                while (true) {
                    const unsigned long long prev = atomicCAS(b.keys + s, LG_EMPTY, key);
                    if (prev == LG_EMPTY || prev == key) break;
                    s = (s + 1) & g.mask;
                }
                slot = (int)s;
                b.next[cell] = atomicExch(b.head + s, (int)cell);
                if (g.strategy == SRX_STRATEGY_VIEW_NORMAL) b.wself[cell] = vn_weight(vnmap[pix]);
            }
        }
        b.cellslot[cell] = slot;
    }
}

template <typename IdT, typename XT>
__global__ void __launch_bounds__(256) k_legacy_accum(const IdT *__restrict__ ids, const XT *__restrict__ x,
                                                       const float *__restrict__ vnmap, LegacyBufs b, LegacyGeom g, long long npx) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < npx; j += (long long)gridDim.x * blockDim.x) {
        const IdPx p = load_id(ids + j);
        if ((p.s | p.m | p.i | p.v) == 0) continue;
        unsigned long long key;
        if (!pack_key<IdT>(p, g.merge, &key)) { atomicOr(b.status + LG_ST_KEY, 1); continue; }
        unsigned int s = hash64(key) & g.mask;
        bool found = false;
        while (true) {
            const unsigned long long k = b.keys[s];
            if (k == key) { found = true; break; }
            if (k == LG_EMPTY) break;
            s = (s + 1) & g.mask;
        }
        if (!found) continue;  // no latent cell samples this surface point: nothing downstream reads its sum
        const int xj = (int)(j % g.W);
        const long long t = j / g.W;
        const int yj = (int)(t % g.H);
        const int fj = (int)(t / g.H);
        const XT *xp = x + (((long long)fj * g.C) * g.h + nearest_src(yj, g.up_sy, g.h)) * g.w + nearest_src(xj, g.up_sx, g.w);
        const long long plane = (long long)g.h * g.w;
        float wj = 1.f;
        if (g.strategy == SRX_STRATEGY_VIEW_NORMAL) wj = vn_weight(vnmap[j]);
        for (int i = b.head[s]; i >= 0; i = b.next[i]) {
            float wgt = wj;
            if (g.strategy == SRX_STRATEGY_FRAME_DISTANCE) {
                const int fi = i / (g.h * g.w);
                wgt = __fdiv_rn(1.f, (float)(abs(fi - fj) + 1));                   // algorithms.py:66-70
            } else if (g.strategy == SRX_STRATEGY_PIXEL_DISTANCE) {
                const int r = i % (g.h * g.w);
                const int yi = nearest_src(r / g.w, g.down_sy, g.H), xi = nearest_src(r % g.w, g.down_sx, g.W);
                wgt = __fdiv_rn(1.f, (float)(abs(xi - xj) + abs(yi - yj) + 1));    // algorithms.py:87-93
            }
            float *dst = b.sum + (long long)i * g.C;
            if ((g.C & 3) == 0) {
                for (int c = 0; c < g.C; c += 4)
                    red_add_f32x4(dst + c, wgt * XIo<XT>::ld(xp + c * plane), wgt * XIo<XT>::ld(xp + (c + 1) * plane),
                                  wgt * XIo<XT>::ld(xp + (c + 2) * plane), wgt * XIo<XT>::ld(xp + (c + 3) * plane));
            } else {
                for (int c = 0; c < g.C; ++c) red_add_f32(dst + c, wgt * XIo<XT>::ld(xp + c * plane));
            }
            red_add_f32(b.wsum + i, wgt);
            red_add_f32(b.cnt + i, 1.f);
        }
    }
}

template <typename XT>
__global__ void __launch_bounds__(256) k_legacy_finalize(XT *__restrict__ x, LegacyBufs b, LegacyGeom g, float alpha,
                                                          float one_minus, long long ncell) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(cell % ((long long)g.h * g.w));
        const int f = (int)(cell / ((long long)g.h * g.w));
        const bool mixed = b.cellslot[cell] >= 0 && b.cnt[cell] >= 2.f;  // traces of length 1 are skipped (overlap.py:123-127)
        float denom = 1.f;
        if (mixed) denom = g.strategy == SRX_STRATEGY_VIEW_NORMAL ? __fmul_rn(b.cnt[cell], b.wself[cell]) : b.wsum[cell];
        for (int c = 0; c < g.C; ++c) {
            const float xi = b.selfx[cell * g.C + c];
            float out = xi;
            if (mixed) {
                const float ov = __fdiv_rn(b.sum[cell * g.C + c], denom);
                out = __fadd_rn(__fmul_rn(alpha, ov), __fmul_rn(one_minus, xi));    // overlap.py:145
            }
            XT *p = x + ((long long)f * g.C + c) * g.h * g.w + r;
            if (g.resize) {
                if (out != 0.f) XIo<XT>::st(p, out);   // torch.where(ovlp != 0, ovlp, original), overlap.py:221
            } else {
                XIo<XT>::st(p, out);
            }
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
struct LegacyLayout {
    int64_t cap, ncell;
    int64_t keys, head, next, cellslot, selfx, sum, wsum, cnt, wself, status, total;
    int64_t zero_begin, zero_end;  // [sum .. cnt, status] cleared per call; [keys .. head] set to 0xFF per call
};

static inline int64_t lg_align(int64_t v) { return (v + 255) / 256 * 256; }

static LegacyLayout legacy_layout(const srx_legacy_desc *d) {
    LegacyLayout L;
    L.ncell = (int64_t)d->frames * d->lat_h * d->lat_w;
    int64_t cap = 1024;
    while (cap < 2 * L.ncell) cap <<= 1;
    L.cap = cap;
    int64_t off = 0;
    L.keys = off; off = lg_align(off + cap * 8);
    L.head = off; off = lg_align(off + cap * 4);
    L.next = off; off = lg_align(off + L.ncell * 4);
    L.cellslot = off; off = lg_align(off + L.ncell * 4);
    L.selfx = off; off = lg_align(off + L.ncell * d->channels * 4);
    L.zero_begin = off;
    L.sum = off; off = lg_align(off + L.ncell * d->channels * 4);
    L.wsum = off; off = lg_align(off + L.ncell * 4);
    L.cnt = off; off = lg_align(off + L.ncell * 4);
    L.status = off; off += 256;
    L.zero_end = off;
    L.wself = off; off = lg_align(off + L.ncell * 4);
    L.total = off;
    return L;
}

static int legacy_validate(const srx_legacy_desc *d) {
    SRX_REQUIRE(d, SRX_ERR_INVALID, "null descriptor");
    SRX_REQUIRE(d->id_dtype == SRX_I32 || d->id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    SRX_REQUIRE(d->frames > 0 && d->height > 0 && d->width > 0 && d->channels > 0 && d->lat_h > 0 && d->lat_w > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(d->strategy >= SRX_STRATEGY_AVERAGE && d->strategy <= SRX_STRATEGY_VIEW_NORMAL, SRX_ERR_INVALID, "Unknown algorithm %d", d->strategy);
    SRX_REQUIRE((int64_t)d->frames * d->lat_h * d->lat_w < (1ll << 30), SRX_ERR_UNSUPPORTED, "too many latent cells for 32-bit chaining");
    return SRX_OK;
}

extern "C" int64_t srx_legacy_workspace_bytes(const srx_legacy_desc *d) {
    if (legacy_validate(d) != SRX_OK) return -1;
    return legacy_layout(d).total;
}

static float nearest_scale(int in_size, int out_size) {
    volatile float s = (float)in_size / (float)out_size;   // one float32 division, as torch computes the scale
    return s;
}

template <typename IdT, typename XT>
static int legacy_impl(const srx_legacy_desc *d, const srx_legacy_args *a, cudaStream_t st) {
    const LegacyLayout L = legacy_layout(d);
    char *ws = reinterpret_cast<char *>(a->workspace_dev);
    LegacyBufs b;
    b.keys = reinterpret_cast<unsigned long long *>(ws + L.keys);
    b.head = reinterpret_cast<int *>(ws + L.head);
    b.next = reinterpret_cast<int *>(ws + L.next);
    b.cellslot = reinterpret_cast<int *>(ws + L.cellslot);
    b.selfx = reinterpret_cast<float *>(ws + L.selfx);
    b.sum = reinterpret_cast<float *>(ws + L.sum);
    b.wsum = reinterpret_cast<float *>(ws + L.wsum);
    b.cnt = reinterpret_cast<float *>(ws + L.cnt);
    b.wself = reinterpret_cast<float *>(ws + L.wself);
    b.status = reinterpret_cast<int *>(ws + L.status);

    SRX_CUDA_CHECK(cudaMemsetAsync(b.keys, 0xFF, (size_t)(L.next - L.keys), st));                       // keys = EMPTY, head = -1
    // sums, counts, status (the status block stays sticky when its check is deferred)
    SRX_CUDA_CHECK(cudaMemsetAsync(ws + L.zero_begin, 0, (size_t)(L.zero_end - L.zero_begin - (a->defer_status ? 256 : 0)), st));

    LegacyGeom g;
    g.T = d->frames; g.H = d->height; g.W = d->width; g.h = d->lat_h; g.w = d->lat_w; g.C = d->channels;
    g.merge = d->merge_len > 1 ? d->merge_len : 1;
    g.strategy = d->strategy;
    g.resize = (d->lat_h != d->height || d->lat_w != d->width) ? 1 : 0;
    g.mask = (unsigned int)(L.cap - 1);
    g.up_sy = nearest_scale(d->lat_h, d->height); g.up_sx = nearest_scale(d->lat_w, d->width);
    g.down_sy = nearest_scale(d->height, d->lat_h); g.down_sx = nearest_scale(d->width, d->lat_w);
    const int sms = srx_sm_count_cached();
    const long long npx = (long long)d->frames * d->height * d->width;
    const IdT *ids = reinterpret_cast<const IdT *>(a->ids_dev);
    XT *x = reinterpret_cast<XT *>(a->x_dev);
    auto blocks = [&](long long n) { long long nb = (n + 255) / 256; return (int)(nb < (long long)sms * 8 ? (nb < 1 ? 1 : nb) : (long long)sms * 8); };
    k_legacy_seed<IdT, XT><<<blocks(L.ncell), 256, 0, st>>>(ids, x, a->view_normal_dev, b, g, L.ncell);
    k_legacy_accum<IdT, XT><<<blocks(npx), 256, 0, st>>>(ids, x, a->view_normal_dev, b, g, npx);
    k_legacy_finalize<XT><<<blocks(L.ncell), 256, 0, st>>>(x, b, g, a->alpha, (float)(1.0 - (double)a->alpha), L.ncell);
    SRX_CUDA_CHECK(cudaGetLastError());
    if (a->defer_status) return SRX_OK;           // no host sync: srx_legacy_check reports (and clears) the status later
    int st_host = 0;
    SRX_CUDA_CHECK(cudaMemcpyAsync(&st_host, b.status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (st_host)
        return srx_set_error(SRX_ERR_KEY_RANGE, "an id component does not fit the packed 64-bit key "
                             "(int32 ids: sprite, material < 1024, third component < 4096 after merge, fourth >= 0)");
    return SRX_OK;
}

template <typename IdT>
static int legacy_x_dispatch(const srx_legacy_desc *d, const srx_legacy_args *a, cudaStream_t st) {
    switch (a->x_dtype) {
        case SRX_F32: return legacy_impl<IdT, float>(d, a, st);
        case SRX_F16: return legacy_impl<IdT, __half>(d, a, st);
        case SRX_BF16: return legacy_impl<IdT, __nv_bfloat16>(d, a, st);
        default: return srx_set_error(SRX_ERR_INVALID, "latent dtype must be f32/f16/bf16");
    }
}

// Deferred status (args->defer_status = 1): reads and clears the sticky status word of earlier srx_legacy_overlap calls on this
// workspace (which must start zeroed); SRX_ERR_KEY_RANGE if any of them met an id that does not fit the packed key.  Syncs.
extern "C" int srx_legacy_check(const srx_legacy_desc *d, const srx_legacy_args *a, void *stream) {
    int rc = legacy_validate(d);
    if (rc) return rc;
    SRX_REQUIRE(a && a->workspace_dev, SRX_ERR_INVALID, "null buffer");
    const LegacyLayout L = legacy_layout(d);
    int *status = reinterpret_cast<int *>(reinterpret_cast<char *>(a->workspace_dev) + L.status);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int st_host = 0;
    SRX_CUDA_CHECK(cudaMemcpyAsync(&st_host, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaMemsetAsync(status, 0, 256, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (st_host)
        return srx_set_error(SRX_ERR_KEY_RANGE, "an id component does not fit the packed 64-bit key "
                             "(int32 ids: sprite, material < 1024, third component < 4096 after merge, fourth >= 0)");
    return SRX_OK;
}

extern "C" int srx_legacy_overlap(const srx_legacy_desc *d, const srx_legacy_args *a, void *stream) {
    int rc = legacy_validate(d);
    if (rc) return rc;
    SRX_REQUIRE(a && a->x_dev && a->ids_dev && a->workspace_dev, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE(d->strategy != SRX_STRATEGY_VIEW_NORMAL || a->view_normal_dev, SRX_ERR_INVALID,
                "perpendicular_view_normal needs view_normal_map (algorithms.py:106)");
    SRX_REQUIRE(a->workspace_bytes >= legacy_layout(d).total, SRX_ERR_INVALID, "workspace too small");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace_dev) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return d->id_dtype == SRX_I32 ? legacy_x_dispatch<int4>(d, a, st) : legacy_x_dispatch<short4>(d, a, st);
}

// =================================================================================================================
// CorrespondenceMap maintenance on the id buffers: which pixel introduces a key (the reference dict's insertion
// order), and dropping whole keys (dropout_index / dropout_in_rectangle, correspondence_map.py:207-274).
// The "map" here is the id buffers themselves; deleting a key = clearing the ids of every pixel that carries it.
// =================================================================================================================
struct KeyTable {
    unsigned long long *keys;   // [cap] packed id tuple, LG_EMPTY = free
    unsigned long long *first;  // [cap] smallest linear pixel index that carries the key
    unsigned int mask;
};

template <typename IdT>
__device__ __forceinline__ bool km_key(const IdT *ids, long long i, int merge, unsigned long long *key, int *status) {
    const IdPx p = load_id(ids + i);
    if ((p.s | p.m | p.i | p.v) == 0) return false;              // correspondence_map.py:153
    if (!pack_key<IdT>(p, merge, key)) { atomicOr(status, 1); return false; }
    return true;
}

__device__ __forceinline__ unsigned int km_insert(const KeyTable &t, unsigned long long key) {
    unsigned int s = hash64(key) & t.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(t.keys + s, LG_EMPTY, key);
        if (prev == LG_EMPTY || prev == key) return s;
        s = (s + 1) & t.mask;
    }
}
__device__ __forceinline__ int km_find(const KeyTable &t, unsigned long long key) {
    unsigned int s = hash64(key) & t.mask;
    while (true) {
        const unsigned long long cur = t.keys[s];
        if (cur == key) return (int)s;
        if (cur == LG_EMPTY) return -1;
        s = (s + 1) & t.mask;
    }
}

template <typename IdT>
__global__ void __launch_bounds__(256) k_km_first(const IdT *__restrict__ ids, long long npx, int merge, KeyTable t, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!km_key(ids, i, merge, &key, status)) continue;
        atomicMin(t.first + km_insert(t, key), (unsigned long long)i);
    }
}
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_first_mask(const IdT *__restrict__ ids, long long npx, int merge, KeyTable t,
                                                        unsigned char *__restrict__ mask, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        unsigned char m = 0;
        if (km_key(ids, i, merge, &key, status)) {
            const int s = km_find(t, key);
            m = (s >= 0 && t.first[s] == (unsigned long long)i) ? 1 : 0;
        }
        mask[i] = m;
    }
}
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_seed(const IdT *__restrict__ ids, const long long *__restrict__ seeds, long long n_seeds,
                                                  long long npx, int merge, KeyTable t, int *status) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_seeds; j += (long long)gridDim.x * blockDim.x) {
        const long long i = seeds[j];
        if (i < 0 || i >= npx) { atomicOr(status, 2); continue; }
        unsigned long long key;
        if (km_key(ids, i, merge, &key, status)) km_insert(t, key);
    }
}
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_drop(IdT *__restrict__ ids, long long npx, int merge, KeyTable t, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!km_key(ids, i, merge, &key, status)) continue;
        if (km_find(t, key) >= 0) {
            IdT z;
            z.x = 0; z.y = 0; z.z = 0; z.w = 0;
            ids[i] = z;
        }
    }
}

static long long km_capacity(long long n) {
    long long cap = 64;
    while (cap < 2 * n + 2) cap <<= 1;
    return cap;
}
extern "C" int64_t srx_corrmap_keys_workspace_bytes(int64_t n_keys_upper_bound) {
    if (n_keys_upper_bound < 0 || n_keys_upper_bound > (1ll << 30)) return -1;
    return km_capacity(n_keys_upper_bound) * 16 + 256;
}

static int km_setup(KeyTable *t, int **status, void *ws, int64_t ws_bytes, long long n, cudaStream_t st) {
    const long long cap = km_capacity(n);
    SRX_REQUIRE(ws && ws_bytes >= cap * 16 + 256, SRX_ERR_INVALID, "workspace too small");
    t->keys = reinterpret_cast<unsigned long long *>(ws);
    t->first = t->keys + cap;
    t->mask = (unsigned int)(cap - 1);
    *status = reinterpret_cast<int *>(t->first + cap);
    SRX_CUDA_CHECK(cudaMemsetAsync(ws, 0xFF, (size_t)cap * 16, st));   // keys = EMPTY, first = max
    SRX_CUDA_CHECK(cudaMemsetAsync(*status, 0, 256, st));
    return SRX_OK;
}
static int km_finish(int *status, cudaStream_t st) {
    int h = 0;
    SRX_CUDA_CHECK(cudaGetLastError());
    SRX_CUDA_CHECK(cudaMemcpyAsync(&h, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h & 2) return srx_set_error(SRX_ERR_INDEX, "seed pixel index out of range");
    if (h & 1) return srx_set_error(SRX_ERR_KEY_RANGE, "an id tuple does not fit the exact 64-bit key (sprite, material < 1024, third component < 4096)");
    return SRX_OK;
}
static int km_grid(long long n) {
    const long long nb = (n + 255) / 256, cap = (long long)srx_sm_count_cached() * 8;
    return (int)(nb < 1 ? 1 : (nb < cap ? nb : cap));
}

// mask_out[i] = 1 where pixel i (frame, row, col order) is the first one carrying its key: the keys of the reference's dict
// in insertion order are the keys of the marked pixels in index order (correspondence_map.py:148-168).  Syncs.
extern "C" int srx_corrmap_first_appearance(const void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                                            uint8_t *mask_out_dev, void *workspace_dev, int64_t workspace_bytes, void *stream) {
    SRX_REQUIRE(ids_dev && mask_out_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(id_dtype == SRX_I32 || id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long npx = (long long)frames * height * width;
    const int merge = merge_len > 1 ? merge_len : 1;
    KeyTable t;
    int *status;
    int rc = km_setup(&t, &status, workspace_dev, workspace_bytes, npx, st);
    if (rc) return rc;
    if (id_dtype == SRX_I32) {
        k_km_first<int4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), npx, merge, t, status);
        k_km_first_mask<int4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), npx, merge, t, mask_out_dev, status);
    } else {
        k_km_first<short4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<const short4 *>(ids_dev), npx, merge, t, status);
        k_km_first_mask<short4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<const short4 *>(ids_dev), npx, merge, t, mask_out_dev, status);
    }
    return km_finish(status, st);
}

// Clears the ids of every pixel whose key equals the key of one of the seed pixels (linear indices into [F,H,W]) — deleting
// keys from the reference's dict (correspondence_map.py:219-223, 268-274).  In place.  Syncs.
extern "C" int srx_corrmap_drop_keys(void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                                     const int64_t *seed_pixels_dev, int64_t n_seeds, void *workspace_dev, int64_t workspace_bytes,
                                     void *stream) {
    SRX_REQUIRE(ids_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0 && n_seeds >= 0, SRX_ERR_INVALID, "bad sizes");
    SRX_REQUIRE(id_dtype == SRX_I32 || id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    if (n_seeds == 0) return SRX_OK;
    SRX_REQUIRE(seed_pixels_dev, SRX_ERR_INVALID, "null seed list");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long npx = (long long)frames * height * width;
    const int merge = merge_len > 1 ? merge_len : 1;
    KeyTable t;
    int *status;
    int rc = km_setup(&t, &status, workspace_dev, workspace_bytes, n_seeds, st);
    if (rc) return rc;
    const long long *seeds = reinterpret_cast<const long long *>(seed_pixels_dev);
    if (id_dtype == SRX_I32) {
        k_km_seed<int4><<<km_grid(n_seeds), 256, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), seeds, n_seeds, npx, merge, t, status);
        k_km_drop<int4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<int4 *>(ids_dev), npx, merge, t, status);
    } else {
        k_km_seed<short4><<<km_grid(n_seeds), 256, 0, st>>>(reinterpret_cast<const short4 *>(ids_dev), seeds, n_seeds, npx, merge, t, status);
        k_km_drop<short4><<<km_grid(npx), 256, 0, st>>>(reinterpret_cast<short4 *>(ids_dev), npx, merge, t, status);
    }
    return km_finish(status, st);
}

// =================================================================================================================
// CorrMapLatentNoiseInitializer (legacy_codes/nodes/latent.py:26-40): every key with at least two entries ("trace") gets
// one random 4-vector for the latent and one for the noise, drawn in the dict's insertion order; all its pixels receive
// them; then a nearest down-sample.  The random rows are drawn by the caller (the reference's CPU generator stream);
// here: which row belongs to which pixel (rank of the pixel's key among the traces, in insertion order) and the fill of
// exactly the pixels the down-sample keeps.
// =================================================================================================================
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_first_count(const IdT *__restrict__ ids, long long npx, int merge, KeyTable t,
                                                         unsigned int *__restrict__ count, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!km_key(ids, i, merge, &key, status)) continue;
        const unsigned int s = km_insert(t, key);
        atomicMin(t.first + s, (unsigned long long)i);
        atomicAdd(count + s, 1u);
    }
}
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_trace_flag(const IdT *__restrict__ ids, long long npx, int merge, KeyTable t,
                                                        const unsigned int *__restrict__ count, int *__restrict__ flag, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        int f = 0;
        if (km_key(ids, i, merge, &key, status)) {
            const int s = km_find(t, key);
            f = (s >= 0 && t.first[s] == (unsigned long long)i && count[s] >= 2u) ? 1 : 0;   // latent.py:29-30 skips singletons
        }
        flag[i] = f;
    }
}
// rank[i] <- rank of the pixel that introduced i's key.  In place: an introducing pixel rewrites its own value, the
// others read only introducing pixels.
template <typename IdT>
__global__ void __launch_bounds__(256) k_km_trace_rank(const IdT *__restrict__ ids, long long npx, int merge, KeyTable t,
                                                        int *rank, int *status) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!km_key(ids, i, merge, &key, status)) continue;
        const int s = km_find(t, key);
        if (s < 0) continue;
        const long long f = (long long)t.first[s];
        if (f != i) rank[i] = rank[f];
    }
}

extern "C" int64_t srx_corrmap_trace_ranks_workspace_bytes(int64_t n_pixels) {
    if (n_pixels < 0 || n_pixels > (1ll << 30)) return -1;
    return km_capacity(n_pixels) * 20 + 256 + lg_align(srx_flags_scan_scratch_ints(n_pixels) * 4);
}

// rank_out[i] = index of pixel i's key among the keys with >= 2 entries, in dict insertion order; -1 for pixels without an
// id and for single-entry keys.  *n_traces_out (host) = number of such keys.  Syncs.
extern "C" int srx_corrmap_trace_ranks(const void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                                       int32_t *rank_out_dev, int64_t *n_traces_out, void *workspace_dev, int64_t workspace_bytes,
                                       void *stream) {
    SRX_REQUIRE(ids_dev && rank_out_dev && n_traces_out, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(id_dtype == SRX_I32 || id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long npx = (long long)frames * height * width;
    SRX_REQUIRE(npx <= (1ll << 30), SRX_ERR_INVALID, "too many pixels");
    SRX_REQUIRE(workspace_dev && workspace_bytes >= srx_corrmap_trace_ranks_workspace_bytes(npx), SRX_ERR_INVALID, "workspace too small");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    const int merge = merge_len > 1 ? merge_len : 1;
    const long long cap = km_capacity(npx);
    KeyTable t;
    int *status;
    int rc = km_setup(&t, &status, workspace_dev, cap * 16 + 256, npx, st);   // keys, first, status
    if (rc) return rc;
    unsigned char *after = reinterpret_cast<unsigned char *>(status) + 256;
    unsigned int *count = reinterpret_cast<unsigned int *>(after);
    int *scratch = reinterpret_cast<int *>(after + cap * 4);
    SRX_CUDA_CHECK(cudaMemsetAsync(count, 0, (size_t)cap * 4, st));
    const int grid = km_grid(npx);
    if (id_dtype == SRX_I32) {
        const int4 *ids = reinterpret_cast<const int4 *>(ids_dev);
        k_km_first_count<int4><<<grid, 256, 0, st>>>(ids, npx, merge, t, count, status);
        k_km_trace_flag<int4><<<grid, 256, 0, st>>>(ids, npx, merge, t, count, rank_out_dev, status);
    } else {
        const short4 *ids = reinterpret_cast<const short4 *>(ids_dev);
        k_km_first_count<short4><<<grid, 256, 0, st>>>(ids, npx, merge, t, count, status);
        k_km_trace_flag<short4><<<grid, 256, 0, st>>>(ids, npx, merge, t, count, rank_out_dev, status);
    }
    rc = srx_flags_to_ranks(rank_out_dev, npx, scratch, n_traces_out, st);
    if (rc) return rc;
    if (id_dtype == SRX_I32)
        k_km_trace_rank<int4><<<grid, 256, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), npx, merge, t, rank_out_dev, status);
    else
        k_km_trace_rank<short4><<<grid, 256, 0, st>>>(reinterpret_cast<const short4 *>(ids_dev), npx, merge, t, rank_out_dev, status);
    return km_finish(status, st);
}

// out[b,c,y,x] = rows[rank, which, c] when frame b's pixel sampled by F.interpolate(mode="nearest") (latent.py:37-38)
// belongs to a trace, else base[which, c, sy, sx] (the same base frame repeats over the batch, latent.py:22-26).
__global__ void __launch_bounds__(256) k_corrmap_noise_fill(const int *__restrict__ rank, const float *__restrict__ base,
                                                             const float *__restrict__ rows, float *__restrict__ latent_out,
                                                             float *__restrict__ noise_out, int frames, int H, int W, int batch,
                                                             int h, int w, float scale_y, float scale_x) {
    const long long total = (long long)batch * h * w;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(o % w), y = (int)((o / w) % h), b = (int)(o / ((long long)w * h));
        int sy = (int)floorf(__fmul_rn((float)y, scale_y)), sx = (int)floorf(__fmul_rn((float)x, scale_x));
        sy = sy < H - 1 ? sy : H - 1;
        sx = sx < W - 1 ? sx : W - 1;
        const long long src = (long long)sy * W + sx;
        const int r = b < frames ? rank[(long long)b * H * W + src] : -1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long dst = (((long long)b * 4 + c) * h + y) * w + x;
            latent_out[dst] = r >= 0 ? rows[(long long)r * 8 + c] : base[(long long)c * H * W + src];
            noise_out[dst] = r >= 0 ? rows[(long long)r * 8 + 4 + c] : base[((long long)4 + c) * H * W + src];
        }
    }
}

extern "C" int srx_corrmap_noise_fill(const int32_t *rank_dev, int frames, int height, int width, int batch, int lat_h, int lat_w,
                                      const float *base_dev, const float *rows_dev, float *latent_out_dev, float *noise_out_dev,
                                      void *stream) {
    SRX_REQUIRE(rank_dev && base_dev && latent_out_dev && noise_out_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0 && batch > 0 && lat_h > 0 && lat_w > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(batch >= frames, SRX_ERR_INDEX, "batch_size %d is smaller than the map's %d frames (IndexError in latent.py:34)", batch, frames);
    const long long total = (long long)batch * lat_h * lat_w;
    const volatile float sy = (float)height / (float)lat_h, sx = (float)width / (float)lat_w;
    k_corrmap_noise_fill<<<km_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        rank_dev, base_dev, rows_dev, latent_out_dev, noise_out_dev, frames, height, width, batch, lat_h, lat_w, sy, sx);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
