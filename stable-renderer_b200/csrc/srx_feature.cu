// srx_feature.cu — wide-channel feature overlap: the body of OverlapCorresponder.post_atten_inject
// (source/common_utils/stable_render_utils/corresponder.py:236-295, unreachable in the reference behind an early return)
// on [B, h*w, c] post-attention features, c = 320 ... 1280.
//
// The reference up-samples the features to the id-map size (nearest), gathers one c-vector per id pixel ([N, c]: gigabytes),
// sorts the keys (`unique`), scatter-adds, blends, writes back with duplicate indices, down-samples and runs AdaIN.  Every entry
// of one (key, feature cell) pair contributes the same c-vector, and only the up-sampled cells the down-sampling reads matter, so:
//   1. k_fo_scan      one pass over the ids: the last entry of every up-sampled cell (its "winner", 64-bit atomicMax) and one
//                     64-bit code (key << 32 | feature cell) per pixel;
//   2. bucketing      CUB radix sort of the codes + run-length encoding = the distinct (key, feature cell) pairs with their
//                     multiplicities, grouped by key — the reference's `unique(return_inverse)` as one sort per id batch;
//   3. k_fo_style     per output cell: winner key -> its pair segment (binary search) -> mean = sum(mult * row) / sum(mult),
//                     the rows (c * 4 bytes, contiguous in the "b (h w) c" layout) fetched by a producer warp with bulk async
//                     copies (cp.async.bulk -> UBLKCP) into a shared-memory ring and reduced there; blend; per-(frame,
//                     channel) sums of style and content for AdaIN;
//   4. k_fo_adain     (x - mu_c) / sigma_c * sigma_s + mu_s, one rounding per op (math_utils.py:78-80).
// Roofline: HBM.  Algorithmic bytes: 16 B per id pixel + (P + 2 B h w) rows of c * s bytes (P pair rows read, every feature
// row read once and written once).
#include "srx_common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#define FO_THREADS 160            // 4 consumer warps + 1 producer warp
#define FO_CONS 128
#define FO_STAGES 4
#define FO_MAXC 1280
#define FO_CPT (FO_MAXC / FO_CONS)   // channels per consumer thread
#define FO_CELLS 8                // output cells per CTA

enum { FO_ST_RANGE = 0 };

__device__ __forceinline__ uint32_t fo_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fo_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fo_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void fo_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fo_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fo_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// F.interpolate(mode='nearest') source index: min(floor(dst * fl32(in / out)), in - 1)
__device__ __forceinline__ int fo_nearest(int dst, int in_size, int out_size) {
    const float scale = __fdiv_rn((float)in_size, (float)out_size);
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

struct FoGeom {
    int F, H, W, B, h, w, c, mh, mw;
    unsigned kcap;
};

template <typename IdT>
__global__ void __launch_bounds__(256) k_fo_scan(const IdT *__restrict__ ids, const int *__restrict__ fmap, FoGeom g,
                                                  unsigned long long *__restrict__ winner, unsigned long long *__restrict__ codes,
                                                  int *status) {
    const long long npx = (long long)g.F * g.H * g.W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (long long)gridDim.x * blockDim.x) {
        const IdPx id = load_id(ids + p);
        unsigned long long code = ~0ull;
        if (id_valid(id)) {
            const long long f = p / ((long long)g.H * g.W);
            const int rem = (int)(p - f * g.H * g.W), y = rem / g.W, x = rem - y * g.W;
            // corrmap.py:239,249: x / height, y / width in float32; corresponder.py:252-253: * id_map.width / .height, truncated
            const int sx = (int)__fmul_rn(__fdiv_rn((float)x, (float)g.H), (float)g.mw);
            const int sy = (int)__fmul_rn(__fdiv_rn((float)y, (float)g.W), (float)g.mh);
            const int b = fmap[f];
            const long long slot = vertex_slot(id.v);
            if (sx >= g.mw || sy >= g.mh || b < 0 || b >= g.B) atomicOr(status + FO_ST_RANGE, 1);
            else if (slot < 0 || slot >= (long long)g.kcap) atomicOr(status + FO_ST_RANGE, 2);
            else {
                const long long U = ((long long)b * g.mh + sy) * g.mw + sx;
                const int src = b * g.h * g.w + fo_nearest(sy, g.h, g.mh) * g.w + fo_nearest(sx, g.w, g.mw);
                atomicMax(winner + U, ((unsigned long long)(rem + 1) << 32) | (unsigned long long)slot);   // entry order = (y, x)
                code = ((unsigned long long)slot << 32) | (unsigned)src;
            }
        }
        codes[p] = code;
    }
}

template <typename XT> __device__ __forceinline__ float fo_ld(const XT *p);
template <> __device__ __forceinline__ float fo_ld<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float fo_ld<__half>(const __half *p) { return __half2float(*p); }
template <> __device__ __forceinline__ float fo_ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

__device__ __forceinline__ int fo_lower_bound(const unsigned long long *a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One CTA = FO_CELLS consecutive output cells of one frame.  Warp 4 (producer) walks the cells' pair segments and keeps a ring of
// FO_STAGES feature rows in flight (one bulk copy per row); warps 0-3 reduce the rows out of shared memory, each thread owning
// channels t, t + 128, ...
template <typename XT>
__global__ void __launch_bounds__(FO_THREADS) k_fo_style(const XT *__restrict__ feat, FoGeom g, const unsigned long long *__restrict__ winner,
                                                          const unsigned long long *__restrict__ pairs, const int *__restrict__ mult,
                                                          const int *__restrict__ npairs_p, float ratio, float one_minus,
                                                          float *__restrict__ style, double *__restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int row_bytes = g.c * (int)sizeof(XT);
    const int stage_bytes = (row_bytes + 127) & ~127;
    const uint32_t sbase = fo_smem_u32(smem);
    const uint32_t bar0 = sbase + FO_STAGES * stage_bytes;              // full[FO_STAGES], empty[FO_STAGES]
    int *s_seg = reinterpret_cast<int *>(smem + FO_STAGES * stage_bytes + 2 * FO_STAGES * 8);   // [FO_CELLS][4]: lo, hi, src0, has
    float *s_mult = reinterpret_cast<float *>(s_seg + FO_CELLS * 4);                                // [FO_STAGES] multiplicity of the staged row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = g.h * g.w;
    const int cells_per_frame = (hw + FO_CELLS - 1) / FO_CELLS;
    const int b = blockIdx.x / cells_per_frame, cell0 = (blockIdx.x - b * cells_per_frame) * FO_CELLS;
    const int ncell = min(FO_CELLS, hw - cell0);
    const int npairs = *npairs_p;
    if (tid == 0) {
        for (int s = 0; s < FO_STAGES; ++s) { fo_mbar_init(bar0 + s * 8, 1); fo_mbar_init(bar0 + (FO_STAGES + s) * 8, FO_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < ncell) {
        // the up-sampled cell this output cell samples (nearest down-sampling), its winner, the winner key's pair segment
        const int cell = cell0 + tid, i = cell / g.w, j = cell - i * g.w;
        const int Y = fo_nearest(i, g.mh, g.h), X = fo_nearest(j, g.mw, g.w);
        const unsigned long long wv = winner[((long long)b * g.mh + Y) * g.mw + X];
        int lo = 0, hi = 0;
        if (wv) {
            const unsigned long long key = wv & 0xffffffffull;
            lo = fo_lower_bound(pairs, npairs, key << 32);
            hi = fo_lower_bound(pairs, npairs, (key + 1) << 32);
        }
        s_seg[tid * 4 + 0] = lo;
        s_seg[tid * 4 + 1] = hi;
        s_seg[tid * 4 + 2] = b * hw + fo_nearest(Y, g.h, g.mh) * g.w + fo_nearest(X, g.w, g.mw);   // the up-sampled value there
        s_seg[tid * 4 + 3] = wv ? 1 : 0;
    }
    __syncthreads();
    const char *fbytes = reinterpret_cast<const char *>(feat);
    if (warp == FO_CONS / 32) {
        // producer: for every cell its base row, then the rows of its pair segment
        int stage = 0;
        unsigned ph = 0;
        for (int q = 0; q < ncell; ++q) {
            const int lo = s_seg[q * 4], hi = s_seg[q * 4 + 1];
            for (int e = lo - 1; e < hi; ++e) {
                fo_mbar_wait(bar0 + (FO_STAGES + stage) * 8, ph ^ 1u);
                if (lane == 0) {
                    const long long row = e < lo ? (long long)s_seg[q * 4 + 2] : (long long)(unsigned)(__ldg(pairs + e) & 0xffffffffull);
                    s_mult[stage] = e < lo ? 0.f : (float)__ldg(mult + e);
                    fo_mbar_expect_tx(bar0 + stage * 8, (uint32_t)row_bytes);
                    fo_bulk_g2s(sbase + stage * stage_bytes, fbytes + row * row_bytes, (uint32_t)row_bytes, bar0 + stage * 8);
                }
                if (++stage == FO_STAGES) { stage = 0; ph ^= 1u; }
            }
        }
        return;
    }
    // consumers
    int stage = 0;
    unsigned ph = 0;
    double st[FO_CPT][2];                         // per-channel sums of style and style^2 over this CTA's cells
#pragma unroll
    for (int k = 0; k < FO_CPT; ++k) st[k][0] = st[k][1] = 0.0;
    for (int q = 0; q < ncell; ++q) {
        const int lo = s_seg[q * 4], hi = s_seg[q * 4 + 1], has = s_seg[q * 4 + 3];
        float base[FO_CPT], acc[FO_CPT];
        float cnt = 0.f;
#pragma unroll
        for (int k = 0; k < FO_CPT; ++k) base[k] = acc[k] = 0.f;
        for (int e = lo - 1; e < hi; ++e) {
            fo_mbar_wait(bar0 + stage * 8, ph);
            const XT *row = reinterpret_cast<const XT *>(smem + stage * stage_bytes);
            const float m = s_mult[stage];
            if (e < lo) {
#pragma unroll
                for (int k = 0; k < FO_CPT; ++k) { const int ch = tid + k * FO_CONS; if (ch < g.c) base[k] = fo_ld<XT>(row + ch); }
            } else {
                cnt += m;
#pragma unroll
                for (int k = 0; k < FO_CPT; ++k) { const int ch = tid + k * FO_CONS; if (ch < g.c) acc[k] = __fadd_rn(acc[k], __fmul_rn(m, fo_ld<XT>(row + ch))); }
            }
            __syncwarp();
            if (lane == 0) fo_mbar_arrive(bar0 + (FO_STAGES + stage) * 8);
            if (++stage == FO_STAGES) { stage = 0; ph ^= 1u; }
        }
        float *out = style + ((long long)b * hw + cell0 + q) * g.c;
#pragma unroll
        for (int k = 0; k < FO_CPT; ++k) {
            const int ch = tid + k * FO_CONS;
            if (ch >= g.c) break;
            // (1 - r) * x + r * mean: mul, mul, add, each rounded (corresponder.py:272-273)
            const float v = has ? __fadd_rn(__fmul_rn(one_minus, base[k]), __fmul_rn(ratio, __fdiv_rn(acc[k], cnt))) : base[k];
            out[ch] = v;
            st[k][0] += (double)v;
            st[k][1] += (double)v * (double)v;
        }
    }
#pragma unroll
    for (int k = 0; k < FO_CPT; ++k) {
        const int ch = tid + k * FO_CONS;
        if (ch >= g.c) break;
        atomicAdd(stats + ((long long)b * 4 + 2) * g.c + ch, st[k][0]);
        atomicAdd(stats + ((long long)b * 4 + 3) * g.c + ch, st[k][1]);
    }
}

// content statistics: per (frame, channel) sums of x and x^2 over the h*w cells.  One CTA = 64 cells x all channels.
template <typename XT>
__global__ void __launch_bounds__(256) k_fo_content_stats(const XT *__restrict__ feat, FoGeom g, double *__restrict__ stats) {
    const int hw = g.h * g.w;
    const int chunks = (hw + 63) / 64;
    const int b = blockIdx.x / chunks, c0 = (blockIdx.x - b * chunks) * 64;
    const int n = min(64, hw - c0);
    for (int ch = threadIdx.x; ch < g.c; ch += blockDim.x) {
        double s = 0.0, s2 = 0.0;
        const XT *p = feat + ((long long)b * hw + c0) * g.c + ch;
        for (int q = 0; q < n; ++q) { const double v = (double)fo_ld<XT>(p + (long long)q * g.c); s += v; s2 += v * v; }
        atomicAdd(stats + ((long long)b * 4 + 0) * g.c + ch, s);
        atomicAdd(stats + ((long long)b * 4 + 1) * g.c + ch, s2);
    }
}

template <typename XT> __device__ __forceinline__ void fo_st(XT *p, float v);
template <> __device__ __forceinline__ void fo_st<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void fo_st<__half>(__half *p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void fo_st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

template <typename XT>
__global__ void __launch_bounds__(256) k_fo_adain(const XT *__restrict__ feat, XT *__restrict__ out, FoGeom g, const double *__restrict__ stats) {
    const int hw = g.h * g.w;
    const long long total = (long long)g.B * hw * g.c;
    const double n = (double)hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % g.c);
        const int b = (int)(i / ((long long)hw * g.c));
        const double *s = stats + (long long)b * 4 * g.c + ch;
        const double sx = s[0], sxx = s[g.c], sb = s[2 * g.c], sbb = s[3 * g.c];
        // unbiased variance + 1e-5, sqrt (math_utils.py:39-47)
        const float mc = (float)(sx / n), ms = (float)(sb / n);
        const float sc = __fsqrt_rn(__fadd_rn((float)((sxx - sx * sx / n) / (n - 1.0)), 1e-5f));
        const float ss = __fsqrt_rn(__fadd_rn((float)((sbb - sb * sb / n) / (n - 1.0)), 1e-5f));
        const float x = fo_ld<XT>(feat + i);
        fo_st<XT>(out + i, __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mc), sc), ss), ms));
    }
}

struct FoLayout {
    int64_t status, npairs, stats, winner, codes_a, codes_b, uniq, mult, style, cub, cub_bytes, total;
};

static inline int64_t fo_align(int64_t v) { return (v + 255) / 256 * 256; }

static int fo_layout(const srx_feature_args *a, FoLayout *L) {
    const int64_t npx = (int64_t)a->frames * a->height * a->width;
    SRX_REQUIRE(npx > 0 && npx < (1ll << 31), SRX_ERR_UNSUPPORTED, "id batch of %lld pixels (limit 2^31)", (long long)npx);
    size_t sort_bytes = 0, rle_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(e));
    e = cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr,
                                           (int *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRunLengthEncode size query failed: %s", cudaGetErrorString(e));
    int64_t off = 0;
    L->status = off; off += 256;
    L->npairs = off; off += 256;
    L->stats = off; off = fo_align(off + (int64_t)a->batch * 4 * a->channels * 8);
    L->winner = off; off = fo_align(off + (int64_t)a->batch * a->map_height * a->map_width * 8);
    L->codes_a = off; off = fo_align(off + npx * 8);
    L->codes_b = off; off = fo_align(off + npx * 8);
    L->uniq = off; off = fo_align(off + npx * 8);
    L->mult = off; off = fo_align(off + npx * 4);
    L->style = off; off = fo_align(off + (int64_t)a->batch * a->lat_h * a->lat_w * a->channels * 4);
    L->cub_bytes = (int64_t)(sort_bytes > rle_bytes ? sort_bytes : rle_bytes);
    L->cub = off; off = fo_align(off + L->cub_bytes);
    L->total = off;
    return SRX_OK;
}

static int fo_validate(const srx_feature_args *a) {
    SRX_REQUIRE(a, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(a->frames > 0 && a->height > 0 && a->width > 0 && a->batch > 0 && a->lat_h > 0 && a->lat_w > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(a->channels > 0 && a->channels <= FO_MAXC, SRX_ERR_UNSUPPORTED, "1..%d channels are supported, got %d", FO_MAXC, a->channels);
    SRX_REQUIRE(a->map_height > 0 && a->map_width > 0, SRX_ERR_INVALID, "non-positive up-sampling size");
    SRX_REQUIRE(a->key_capacity > 0 && a->key_capacity <= (1ll << 31), SRX_ERR_INVALID, "key capacity must be in 1..2^31");
    SRX_REQUIRE(a->id_dtype == SRX_I32 || a->id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    SRX_REQUIRE(a->x_dtype == SRX_F32 || a->x_dtype == SRX_F16 || a->x_dtype == SRX_BF16, SRX_ERR_INVALID, "feature dtype must be f32, f16 or bf16");
    const int64_t rb = (int64_t)a->channels * (a->x_dtype == SRX_F32 ? 4 : 2);
    SRX_REQUIRE(rb % 16 == 0, SRX_ERR_UNSUPPORTED, "feature rows must be a multiple of 16 bytes (bulk copies): %lld", (long long)rb);
    return SRX_OK;
}

extern "C" int64_t srx_feature_overlap_workspace_bytes(const srx_feature_args *a) {
    if (fo_validate(a)) return -1;
    FoLayout L;
    if (fo_layout(a, &L)) return -1;
    return L.total;
}

template <typename XT>
static int fo_run(const srx_feature_args *a, const FoLayout &L, const int *fmap_dev, cudaStream_t st) {
    char *ws = reinterpret_cast<char *>(a->workspace);
    FoGeom g{a->frames, a->height, a->width, a->batch, a->lat_h, a->lat_w, a->channels, a->map_height, a->map_width, (unsigned)a->key_capacity};
    const int64_t npx = (int64_t)a->frames * a->height * a->width;
    int *status = reinterpret_cast<int *>(ws + L.status);
    int *npairs = reinterpret_cast<int *>(ws + L.npairs);
    double *stats = reinterpret_cast<double *>(ws + L.stats);
    unsigned long long *winner = reinterpret_cast<unsigned long long *>(ws + L.winner);
    unsigned long long *ca = reinterpret_cast<unsigned long long *>(ws + L.codes_a), *cb = reinterpret_cast<unsigned long long *>(ws + L.codes_b);
    unsigned long long *uniq = reinterpret_cast<unsigned long long *>(ws + L.uniq);
    int *mult = reinterpret_cast<int *>(ws + L.mult);
    float *style = reinterpret_cast<float *>(ws + L.style);
    SRX_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)L.codes_a, st));          // status, pair count, statistics, winners
    const int sms = srx_sm_count_cached();
    const long long nb = (npx + 255) / 256;
    const int grid = (int)(nb < (long long)sms * 16 ? nb : (long long)sms * 16);
    if (a->id_dtype == SRX_I32) k_fo_scan<int4><<<grid, 256, 0, st>>>(reinterpret_cast<const int4 *>(a->ids_dev), fmap_dev, g, winner, ca, status);
    else k_fo_scan<short4><<<grid, 256, 0, st>>>(reinterpret_cast<const short4 *>(a->ids_dev), fmap_dev, g, winner, ca, status);
    SRX_CUDA_CHECK(cudaGetLastError());
    size_t cub_bytes = (size_t)L.cub_bytes;
    SRX_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ws + L.cub, cub_bytes, ca, cb, (int)npx, 0, 64, st));
    cub_bytes = (size_t)L.cub_bytes;
    SRX_CUDA_CHECK(cub::DeviceRunLengthEncode::Encode(ws + L.cub, cub_bytes, cb, uniq, mult, npairs, (int)npx, st));
    // (the run of ~0 codes — pixels without an entry — sorts last and is never inside a key's segment)
    const XT *feat = reinterpret_cast<const XT *>(a->feat_dev);
    const int hw = a->lat_h * a->lat_w;
    k_fo_content_stats<XT><<<a->batch * ((hw + 63) / 64), 256, 0, st>>>(feat, g, stats);
    const int row_bytes = a->channels * (int)sizeof(XT);
    const int smem = FO_STAGES * ((row_bytes + 127) & ~127) + 2 * FO_STAGES * 8 + FO_CELLS * 16 + FO_STAGES * 4 + 64;
    static bool configured[3] = {false, false, false};
    const int ti = sizeof(XT) == 4 ? 0 : (a->x_dtype == SRX_F16 ? 1 : 2);
    if (!configured[ti]) {
        SRX_CUDA_CHECK(cudaFuncSetAttribute(k_fo_style<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FO_STAGES * FO_MAXC * 4 + 1024));
        configured[ti] = true;
    }
    k_fo_style<XT><<<a->batch * ((hw + FO_CELLS - 1) / FO_CELLS), FO_THREADS, smem, st>>>(
        feat, g, winner, uniq, mult, npairs, a->ratio, (float)(1.0 - (double)a->ratio), style, stats);
    SRX_CUDA_CHECK(cudaGetLastError());
    const long long total = (long long)a->batch * hw * a->channels;
    const long long nb2 = (total + 255) / 256;
    k_fo_adain<XT><<<(int)(nb2 < (long long)sms * 16 ? nb2 : (long long)sms * 16), 256, 0, st>>>(feat, reinterpret_cast<XT *>(a->out_dev), g, stats);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

extern "C" int srx_feature_overlap(const srx_feature_args *a, void *stream) {
    int rc = fo_validate(a);
    if (rc) return rc;
    SRX_REQUIRE(a->ids_dev && a->feat_dev && a->out_dev && a->workspace && a->frame_map_dev, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE((reinterpret_cast<uintptr_t>(a->feat_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, SRX_ERR_INVALID,
                "features must be 16-byte aligned, the workspace 256-byte aligned");
    FoLayout L;
    if ((rc = fo_layout(a, &L))) return rc;
    SRX_REQUIRE(a->workspace_bytes >= L.total, SRX_ERR_INVALID, "workspace too small: %lld < %lld", (long long)a->workspace_bytes, (long long)L.total);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->x_dtype == SRX_F32) return fo_run<float>(a, L, a->frame_map_dev, st);
    if (a->x_dtype == SRX_F16) return fo_run<__half>(a, L, a->frame_map_dev, st);
    return fo_run<__nv_bfloat16>(a, L, a->frame_map_dev, st);
}

// device-side failures of the last srx_feature_overlap on this workspace (syncs the stream)
extern "C" int srx_feature_overlap_check(const srx_feature_args *a, void *stream) {
    SRX_REQUIRE(a && a->workspace, SRX_ERR_INVALID, "null argument");
    int host[2] = {0, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SRX_CUDA_CHECK(cudaMemcpyAsync(host, a->workspace, sizeof(host), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (host[FO_ST_RANGE] & 1) return srx_set_error(SRX_ERR_INDEX, "index out of range: an id pixel maps outside the up-sampled features or names a frame outside the batch");
    if (host[FO_ST_RANGE] & 2) return srx_set_error(SRX_ERR_KEY_RANGE, "a vertex id fell outside the key capacity (%lld)", (long long)a->key_capacity);
    return SRX_OK;
}

// =====================================================================================================================
// Cell-similarity overlap — taichi_cells_overlap (source/common_utils/stable_render_utils/corr_utils.py:110-134).
// new[A] = (new[A] + val[A] + sum_{B != A} sim(A,B) val[B]) * (1 / (1 + sum_{B != A} sim(A,B))),
// sim(A,B) = sum_{x in A, i in B} contrib[x] contrib[i] [id[x] == id[i]]  (all four components; no validity filter, :28-44);
// a cell = pixels // cells CONSECUTIVE pixels of the flattened frame (:130).  The reference loops over all cell pairs and all
// pixel pairs (O(cells^2 px^2)); with c_A[key] = the summed contribution of A's pixels that carry `key`,
//     sim(A,B) = sum_key c_A[key] c_B[key]   =>   sum_{B != A} sim(A,B) val[B] = sum_key c_A[key] (S[key] - c_A[key] val[A]),
//     S[key] = sum_B c_B[key] val[B],  T[key] = sum_B c_B[key]   —   linear in pixels + pairs.
//   k_cs_insert   id tuple -> open-addressing table (exact 64-bit packing), dense rank per distinct key
//   k_cs_pairs    a warp per cell: the distinct keys of the cell with their summed contributions (leader pixel per key)
//   k_cs_scatter  S[key] += c * val[cell] (channels over the threads), T[key] += c
//   k_cs_finish   the formula above per cell
// =====================================================================================================================
#define CS_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ bool cs_pack(const int4 &p, unsigned long long *key) {
    if (p.x < 0 || p.x >= 1024 || p.y < 0 || p.y >= 1024 || p.z < 0 || p.z >= 4096 || p.w < 0) return false;
    *key = ((unsigned long long)p.x << 54) | ((unsigned long long)p.y << 44) | ((unsigned long long)p.z << 32) | (unsigned long long)(unsigned)p.w;
    return true;
}
__device__ __forceinline__ unsigned cs_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__global__ void __launch_bounds__(256) k_cs_insert(const int4 *__restrict__ ids, long long npx_total, unsigned long long *keys,
                                                    int *__restrict__ rank_of, int *__restrict__ px_slot, unsigned mask, int *counters) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npx_total; p += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        int slot = -1;
        if (!cs_pack(ids[p], &key)) atomicOr(counters + 1, 1);
        else {
            unsigned s = cs_hash(key) & mask;
            while (true) {
                const unsigned long long prev = atomicCAS(keys + s, CS_EMPTY, key);
                if (prev == CS_EMPTY) { rank_of[s] = atomicAdd(counters, 1); break; }
                if (prev == key) break;
                s = (s + 1) & mask;
            }
            slot = (int)s;
        }
        px_slot[p] = slot;
    }
}

// pair_rank / pair_w [cells_total][cpp]: entry j of a cell is (rank, summed contribution) when pixel j is the first of the
// cell with its key, else rank = -1
__global__ void __launch_bounds__(256) k_cs_pairs(const int *__restrict__ px_slot, const int *__restrict__ rank_of, const float *__restrict__ contrib,
                                                   long long cells_total, int cells, int cpp, int npx, int *__restrict__ pair_rank,
                                                   float *__restrict__ pair_w) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long cell = warp; cell < cells_total; cell += nwarps) {
        const long long b = cell / cells, ci = cell - b * cells;
        const long long base = b * npx + ci * cpp;
        for (int j = lane; j < cpp; j += 32) {
            const int sj = px_slot[base + j];
            bool leader = sj >= 0;
            for (int i = 0; i < j && leader; ++i) leader = px_slot[base + i] != sj;
            float w = 0.f;
            if (leader)
                for (int i = j; i < cpp; ++i) if (px_slot[base + i] == sj) w += contrib[base + i];
            pair_rank[cell * cpp + j] = leader ? rank_of[sj] : -1;
            pair_w[cell * cpp + j] = w;
        }
    }
}

__global__ void __launch_bounds__(128) k_cs_scatter(const float *__restrict__ val, const int *__restrict__ pair_rank, const float *__restrict__ pair_w,
                                                     int cpp, int c, float *__restrict__ S, float *__restrict__ T) {
    const long long cell = blockIdx.x;
    for (int j = 0; j < cpp; ++j) {
        const int r = pair_rank[cell * cpp + j];
        if (r < 0) continue;
        const float w = pair_w[cell * cpp + j];
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) red_add_f32(S + (long long)r * c + ch, __fmul_rn(w, val[cell * c + ch]));
        if (threadIdx.x == 0) red_add_f32(T + r, w);
    }
}

__global__ void __launch_bounds__(128) k_cs_finish(const float *__restrict__ val, const int *__restrict__ pair_rank, const float *__restrict__ pair_w,
                                                    int cpp, int c, const float *__restrict__ S, const float *__restrict__ T, float *__restrict__ out) {
    const long long cell = blockIdx.x;
    float den = 1.f;
    for (int j = 0; j < cpp; ++j) {
        const int r = pair_rank[cell * cpp + j];
        if (r >= 0) { const float w = pair_w[cell * cpp + j]; den += w * (T[r] - w); }
    }
    const float inv = __fdiv_rn(1.f, den);                  // corr_utils.py:108: values *= 1.0 / total_sim
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float v = val[cell * c + ch];
        float num = out[cell * c + ch] + v;                 // :85-90: the target cell's own value is added to the placeholder first
        for (int j = 0; j < cpp; ++j) {
            const int r = pair_rank[cell * cpp + j];
            if (r >= 0) { const float w = pair_w[cell * cpp + j]; num += w * (S[(long long)r * c + ch] - w * v); }
        }
        out[cell * c + ch] = __fmul_rn(num, inv);
    }
}

static inline int64_t cs_cap(int64_t npx_total) { int64_t cap = 1024; while (cap < 2 * npx_total) cap <<= 1; return cap; }

// workspace of the key pass: [counters 256][keys cap*8][rank_of cap*4][px_slot npx*4][pair_rank npx*4][pair_w npx*4]
extern "C" int64_t srx_cells_overlap_workspace_bytes(int batch, int pixels) {
    if (batch <= 0 || pixels <= 0) return -1;
    const int64_t n = (int64_t)batch * pixels, cap = cs_cap(n);
    return 256 + fo_align(cap * 8) + fo_align(cap * 4) + 3 * fo_align(n * 4);
}

// Pass 1: distinct id tuples -> *n_keys_out (host; syncs) — the caller then allocates the [n_keys, c] + [n_keys] float key sums.
extern "C" int srx_cells_overlap_keys(const int32_t *ids_dev, const float *contrib_dev, int batch, int pixels, int cells, void *workspace,
                                      int64_t workspace_bytes, int64_t *n_keys_out, void *stream) {
    SRX_REQUIRE(ids_dev && contrib_dev && workspace && n_keys_out, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(batch > 0 && pixels > 0 && cells > 0 && pixels >= cells, SRX_ERR_INVALID, "bad sizes");
    SRX_REQUIRE(workspace_bytes >= srx_cells_overlap_workspace_bytes(batch, pixels), SRX_ERR_INVALID, "workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)batch * pixels, cap = cs_cap(n);
    char *ws = reinterpret_cast<char *>(workspace);
    int *counters = reinterpret_cast<int *>(ws);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + 256);
    int *rank_of = reinterpret_cast<int *>(ws + 256 + fo_align(cap * 8));
    int *px_slot = rank_of + fo_align(cap * 4) / 4;
    int *pair_rank = px_slot + fo_align(n * 4) / 4;
    float *pair_w = reinterpret_cast<float *>(pair_rank + fo_align(n * 4) / 4);
    SRX_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256, st));
    SRX_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, (size_t)cap * 8, st));
    const int sms = srx_sm_count_cached();
    const long long nb = (n + 255) / 256;
    k_cs_insert<<<(int)(nb < (long long)sms * 16 ? nb : (long long)sms * 16), 256, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), n, keys, rank_of,
                                                                                             px_slot, (unsigned)(cap - 1), counters);
    const int cpp = pixels / cells;
    const long long cells_total = (long long)batch * cells;
    // pixels past cells * cpp belong to no cell (corr_utils.py:130) — the pair pass never reads them
    const long long nbw = (cells_total * 32 + 255) / 256;
    k_cs_pairs<<<(int)(nbw < (long long)sms * 16 ? nbw : (long long)sms * 16), 256, 0, st>>>(px_slot, rank_of, contrib_dev, cells_total, cells, cpp, pixels,
                                                                                              pair_rank, pair_w);
    SRX_CUDA_CHECK(cudaGetLastError());
    int host[2] = {0, 0};
    SRX_CUDA_CHECK(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (host[1]) return srx_set_error(SRX_ERR_KEY_RANGE, "an id component does not fit the packed 64-bit key (sprite, material < 1024, "
                                      "map index < 4096, vertex id >= 0)");
    *n_keys_out = host[0];
    return SRX_OK;
}

// Pass 2: values [batch, cells, c] f32, new_values [batch, cells, c] f32 (the reference's placeholder: its content is added to,
// corr_utils.py:85-90), key_sums = zeroed [n_keys * (c + 1)] floats.
extern "C" int srx_cells_overlap(const float *values_dev, float *new_values_dev, int batch, int pixels, int cells, int channels, void *workspace,
                                 float *key_sums_dev, int64_t n_keys, void *stream) {
    SRX_REQUIRE(values_dev && new_values_dev && workspace && key_sums_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(batch > 0 && pixels > 0 && cells > 0 && channels > 0 && n_keys >= 0, SRX_ERR_INVALID, "bad sizes");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)batch * pixels, cap = cs_cap(n);
    char *ws = reinterpret_cast<char *>(workspace);
    int *rank_of = reinterpret_cast<int *>(ws + 256 + fo_align(cap * 8));
    int *px_slot = rank_of + fo_align(cap * 4) / 4;
    int *pair_rank = px_slot + fo_align(n * 4) / 4;
    float *pair_w = reinterpret_cast<float *>(pair_rank + fo_align(n * 4) / 4);
    float *S = key_sums_dev, *T = key_sums_dev + n_keys * channels;
    const int cpp = pixels / cells;
    const unsigned cells_total = (unsigned)((long long)batch * cells);
    k_cs_scatter<<<cells_total, 128, 0, st>>>(values_dev, pair_rank, pair_w, cpp, channels, S, T);
    k_cs_finish<<<cells_total, 128, 0, st>>>(values_dev, pair_rank, pair_w, cpp, channels, S, T, new_values_dev);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
