// srx_feature.cu — wide-channel feature overlap: the body of OverlapCorresponder.post_atten_inject
// (source/common_utils/stable_render_utils/corresponder.py:236-295, unreachable in the reference behind an early return)
// on [B, h*w, c] post-attention features, c = 320 ... 1280.
//
// The reference up-samples the features to the id-map size (nearest), gathers one c-vector per id pixel ([N, c]: gigabytes),
// sorts the keys (`unique`), scatter-adds, blends, writes back with duplicate indices, down-samples and runs AdaIN.  Every entry
// of one (key, feature cell) pair contributes the same c-vector, and only the up-sampled cells the down-sampling reads matter, so:
//   bucketing pass — depends on the ids and the sizes only, NOT on the features: done once per id batch and feature size, then
//   reused by every attention layer and every denoise step that sees the same ids (`reuse_buckets`)
//     1. k_fo_scan      one pass over the ids: the last entry of every up-sampled cell (its "winner", 64-bit atomicMax) and one
//                       code (key << sb | feature cell) per pixel, sb = bits of the cell index;
//     2. CUB radix sort of the codes over their sb + kb + 1 significant bits + run-length encoding = the distinct (key, feature
//        cell) pairs with their multiplicities, grouped by key — the reference's `unique(return_inverse)`;
//   per call
//     3. k_fo_content_stats  per-(frame, channel) sums of x and x^2 (a warp per cell, 16-byte loads);
//     4. k_fo_style     per output cell: winner key -> its pair segment (binary search) -> mean = sum(mult * row) / sum(mult),
//                       the rows (c * s bytes, contiguous in the "b (h w) c" layout) fetched by a producer warp with bulk async
//                       copies (cp.async.bulk -> UBLKCP) into a shared-memory ring and reduced there; blend; per-(frame,
//                       channel) sums of the blended features (the blended tensor itself is never stored: AdaIN needs only its
//                       statistics);
//     5. k_fo_coef + k_fo_adain   (x - mu_c) / sigma_c * sigma_s + mu_s, one rounding per op (math_utils.py:78-80), 16 bytes per thread.
// Roofline: HBM for the streaming parts, L2 for the row gathers of step 4.  Algorithmic bytes per call: (R + 3 B h w) rows of
// c * s bytes (R = the rows step 4 gathers = sum over output cells of the winner key's pair count; every feature row read twice —
// statistics, AdaIN — and written once) + 16 B per id pixel when the buckets are built.
#include "srx_common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#define FO_THREADS 160            // 4 consumer warps + 1 producer warp
#define FO_CONS 128
#define FO_MIN_STAGES 4
#define FO_MAX_STAGES 16
#define FO_RING_BYTES (24 * 1024) // ring budget per CTA: stages = clamp(FO_RING_BYTES / stage bytes, 4, 16)
#define FO_MAXC 1280
#define FO_CELLS 8                // output cells per CTA

enum { FO_ST_RANGE = 0, FO_ST_ROWS = 2 };      // status words: [0] range failures, [2..3] rows gathered by the last call (u64)

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
    int sb, kb;                    // bits of a feature-cell index / of a key inside a code; bit sb + kb marks "no entry"
};

template <typename IdT>
__global__ void __launch_bounds__(256) k_fo_scan(const IdT *__restrict__ ids, const int *__restrict__ fmap, FoGeom g,
                                                  unsigned long long *__restrict__ winner, unsigned long long *__restrict__ codes,
                                                  int *status) {
    const long long npx = (long long)g.F * g.H * g.W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (long long)gridDim.x * blockDim.x) {
        const IdPx id = load_id(ids + p);
        unsigned long long code = 1ull << (g.sb + g.kb);           // sorts behind every (key, cell) code
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
                code = ((unsigned long long)slot << g.sb) | (unsigned)src;
            }
        }
        codes[p] = code;
    }
}

template <typename XT> __device__ __forceinline__ float fo_ld(const XT *p);
template <> __device__ __forceinline__ float fo_ld<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float fo_ld<__half>(const __half *p) { return __half2float(*p); }
template <> __device__ __forceinline__ float fo_ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

// 16 bytes of features <-> floats
template <typename XT> struct FoVec;
template <> struct FoVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4 &v, float (&o)[4]) {
        o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y); o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
    }
    static __device__ __forceinline__ uint4 pack(const float (&o)[4]) {
        return make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
    }
};
template <> struct FoVec<__half> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4 &v, float (&o)[8]) {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[k]));
            o[2 * k] = f.x; o[2 * k + 1] = f.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&o)[8]) {
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 hh = __halves2half2(__float2half_rn(o[2 * k]), __float2half_rn(o[2 * k + 1]));
            w[k] = *reinterpret_cast<const unsigned *>(&hh);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct FoVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4 &v, float (&o)[8]) {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[2 * k] = __uint_as_float(w[k] << 16); o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
    }
    static __device__ __forceinline__ uint4 pack(const float (&o)[8]) {
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 hh = __halves2bfloat162(__float2bfloat16_rn(o[2 * k]), __float2bfloat16_rn(o[2 * k + 1]));
            w[k] = *reinterpret_cast<const unsigned *>(&hh);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ int fo_lower_bound(const unsigned long long *a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// once per bucketing pass: for every output cell the up-sampled cell it samples (nearest down-sampling), that cell's winner and the
// winner key's pair segment — (lo, hi, row of the up-sampled value, has a winner)
__global__ void __launch_bounds__(256) k_fo_segments(FoGeom g, const unsigned long long *__restrict__ winner,
                                                      const unsigned long long *__restrict__ pairs, const int *__restrict__ npairs_p,
                                                      int4 *__restrict__ cellseg) {
    const int hw = g.h * g.w;
    const int n = g.B * hw;
    const int npairs = *npairs_p;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, cell = idx - b * hw, i = cell / g.w, j = cell - i * g.w;
        const int Y = fo_nearest(i, g.mh, g.h), X = fo_nearest(j, g.mw, g.w);
        const unsigned long long wv = winner[((long long)b * g.mh + Y) * g.mw + X];
        int lo = 0, hi = 0;
        if (wv) {
            const unsigned long long key = wv & 0xffffffffull;
            lo = fo_lower_bound(pairs, npairs, key << g.sb);
            hi = fo_lower_bound(pairs, npairs, (key + 1) << g.sb);
        }
        cellseg[idx] = make_int4(lo, hi, b * hw + fo_nearest(Y, g.h, g.mh) * g.w + fo_nearest(X, g.w, g.mw), wv ? 1 : 0);
    }
}

// One CTA = FO_CELLS consecutive output cells of one frame.  Warp 4 (producer) walks the cells' pair segments — its 32 lanes resolve 32
// rows at a time, the next 32 are requested before the current ones are issued — and keeps a ring of `nst` feature rows in flight (one
// bulk copy per row); the consumer warps reduce the rows out of shared memory, a thread owning the 16-byte vectors t, t + 128, ... of a
// row (KV of them).  Row order per cell: the pair rows, then the cell's own ("base") row, at which the cell is finalised straight
// from shared memory.  Narrow rows do not need all four consumer warps: only the first `nact` take part in the ring.
template <typename XT, int KV>
__global__ void __launch_bounds__(FO_THREADS) k_fo_style(const XT *__restrict__ feat, FoGeom g, const int4 *__restrict__ cellseg,
                                                          const unsigned long long *__restrict__ pairs, const int *__restrict__ mult,
                                                          float ratio, float one_minus, int nst,
                                                          double *__restrict__ stats, unsigned long long *__restrict__ rows_out) {
    constexpr int VEC = FoVec<XT>::N;
    extern __shared__ __align__(128) unsigned char smem[];
    const int row_bytes = g.c * (int)sizeof(XT);
    const int stage_bytes = (row_bytes + 127) & ~127;
    const int nvec = row_bytes >> 4;
    const int nact = min(FO_CONS / 32, (nvec + 31) >> 5);               // consumer warps that own a vector
    const uint32_t sbase = fo_smem_u32(smem);
    const uint32_t bar0 = sbase + nst * stage_bytes;                    // full[nst], empty[nst]
    int *s_seg = reinterpret_cast<int *>(smem + nst * stage_bytes + 2 * nst * 8);   // [FO_CELLS][4]: lo, hi, src0, has
    float *s_mult = reinterpret_cast<float *>(s_seg + FO_CELLS * 4);                  // [nst] multiplicity of the staged row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = g.h * g.w;
    const int cells_per_frame = (hw + FO_CELLS - 1) / FO_CELLS;
    const int b = blockIdx.x / cells_per_frame, cell0 = (blockIdx.x - b * cells_per_frame) * FO_CELLS;
    const int ncell = min(FO_CELLS, hw - cell0);
    const unsigned long long src_mask = (1ull << g.sb) - 1ull;
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { fo_mbar_init(bar0 + s * 8, 1); fo_mbar_init(bar0 + (nst + s) * 8, nact); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < ncell) reinterpret_cast<int4 *>(s_seg)[tid] = __ldg(cellseg + (long long)b * hw + cell0 + tid);
    __syncthreads();
    const char *fbytes = reinterpret_cast<const char *>(feat);
    if (warp == FO_CONS / 32) {
        // producer.  The CTA's rows are numbered 0..total-1 across the cells (per cell: its pair rows, then its base row).
        int stage = 0;
        unsigned ph = 0;
        int total = 0;
        for (int q = 0; q < ncell; ++q) total += s_seg[q * 4 + 1] - s_seg[q * 4] + 1;
        auto resolve = [&](int r, unsigned &row, float &m) {
            row = 0u; m = 0.f;
            if (r >= total) return;
            int q = 0, first = 0, len;
            for (;;) {                                   // ncell <= FO_CELLS segments
                len = s_seg[q * 4 + 1] - s_seg[q * 4] + 1;
                if (r < first + len) break;
                first += len; ++q;
            }
            const int k = r - first;
            if (k == len - 1) { row = (unsigned)s_seg[q * 4 + 2]; return; }
            const int e = s_seg[q * 4] + k;
            row = (unsigned)(__ldg(pairs + e) & src_mask);
            m = (float)__ldg(mult + e);
        };
        unsigned nrow; float nm;
        resolve(lane, nrow, nm);
        for (int r0 = 0; r0 < total; r0 += 32) {
            const unsigned my_row = nrow;
            const float my_m = nm;
            resolve(r0 + 32 + lane, nrow, nm);
            const int cnt = min(32, total - r0);
            for (int j = 0; j < cnt; ++j) {
                const unsigned row = __shfl_sync(0xffffffffu, my_row, j);
                const float m = __shfl_sync(0xffffffffu, my_m, j);
                fo_mbar_wait(bar0 + (nst + stage) * 8, ph ^ 1u);
                if (lane == 0) {
                    s_mult[stage] = m;
                    fo_mbar_expect_tx(bar0 + stage * 8, (uint32_t)row_bytes);
                    fo_bulk_g2s(sbase + stage * stage_bytes, fbytes + (long long)row * row_bytes, (uint32_t)row_bytes, bar0 + stage * 8);
                }
                if (++stage == nst) { stage = 0; ph ^= 1u; }
            }
        }
        if (lane == 0) atomicAdd(rows_out, (unsigned long long)total);
        return;
    }
    if (warp >= nact) return;
    // consumers
    int stage = 0;
    unsigned ph = 0;
    double st[KV * VEC][2];                       // per-channel sums of style and style^2 over this CTA's cells
    float acc[KV * VEC];
#pragma unroll
    for (int k = 0; k < KV * VEC; ++k) { st[k][0] = st[k][1] = 0.0; acc[k] = 0.f; }
    for (int q = 0; q < ncell; ++q) {
        const int lo = s_seg[q * 4], hi = s_seg[q * 4 + 1], has = s_seg[q * 4 + 3];
        float cnt = 0.f;
        for (int e = lo; e <= hi; ++e) {
            fo_mbar_wait(bar0 + stage * 8, ph);
            const uint4 *row = reinterpret_cast<const uint4 *>(smem + stage * stage_bytes);
            if (e < hi) {
                const float m = s_mult[stage];
                cnt += m;
#pragma unroll
                for (int k = 0; k < KV; ++k) {
                    const int vi = tid + k * FO_CONS;
                    if (vi < nvec) {
                        float x[VEC];
                        FoVec<XT>::unpack(row[vi], x);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc[k * VEC + j] = __fadd_rn(acc[k * VEC + j], __fmul_rn(m, x[j]));
                    }
                }
            } else {
                // the cell's own row: (1 - r) * x + r * mean — mul, mul, add, each rounded (corresponder.py:272-273)
#pragma unroll
                for (int k = 0; k < KV; ++k) {
                    const int vi = tid + k * FO_CONS;
                    if (vi < nvec) {
                        float x[VEC];
                        FoVec<XT>::unpack(row[vi], x);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float v = has ? __fadd_rn(__fmul_rn(one_minus, x[j]), __fmul_rn(ratio, __fdiv_rn(acc[k * VEC + j], cnt))) : x[j];
                            st[k * VEC + j][0] += (double)v;
                            st[k * VEC + j][1] += (double)v * (double)v;
                            acc[k * VEC + j] = 0.f;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) fo_mbar_arrive(bar0 + (nst + stage) * 8);
            if (++stage == nst) { stage = 0; ph ^= 1u; }
        }
    }
#pragma unroll
    for (int k = 0; k < KV; ++k) {
        const int vi = tid + k * FO_CONS;
        if (vi >= nvec) break;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            atomicAdd(stats + ((long long)b * 4 + 2) * g.c + vi * VEC + j, st[k * VEC + j][0]);
            atomicAdd(stats + ((long long)b * 4 + 3) * g.c + vi * VEC + j, st[k * VEC + j][1]);
        }
    }
}

// Narrow rows (up to 2 KB: c = 320, 640 in half precision, c = 320 in fp32): a warp covers a whole row in at most four 16-byte passes,
// so every warp runs its OWN cells — FOW_CELLS consecutive ones — and fetches its rows with plain 128-bit loads, UN rows in flight.
// (Bulk copies do not pay for rows this small: measured on config "16 x 64^2 x 320", 503 k rows of 640 B, one CTA-wide ring 146 us,
// a ring per warp 185 us — about one row per 50 ns and SM whatever the structure; see profiles/r2_feature_overlap.txt.)
// Statistics are summed over the warp's cells in float and folded over the CTA in double.
#define FOW_WARPS 4
#define FOW_CELLS 8                // cells per warp (default; SRX_FO_CELLS = 1..FOW_MAX_CELLS)
#define FOW_MAX_CELLS 16
template <typename XT, int P>
__global__ void __launch_bounds__(FOW_WARPS * 32) k_fo_style_warp(const XT *__restrict__ feat, FoGeom g, const int4 *__restrict__ cellseg,
                                                                   const unsigned long long *__restrict__ pairs, const int *__restrict__ mult,
                                                                   float ratio, float one_minus, int cpw,
                                                                   double *__restrict__ stats, unsigned long long *__restrict__ rows_out) {
    constexpr int VEC = FoVec<XT>::N;
    constexpr int UN = P <= 2 ? 4 : 2;
    extern __shared__ __align__(16) unsigned char smem[];
    float *s_stat = reinterpret_cast<float *>(smem);                  // [FOW_WARPS][2][c]
    const int row_vecs = (g.c * (int)sizeof(XT)) >> 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = g.h * g.w;
    const int per_cta = FOW_WARPS * cpw;
    const int ctas_per_frame = (hw + per_cta - 1) / per_cta;
    const int b = blockIdx.x / ctas_per_frame;
    const int wc0 = (blockIdx.x - b * ctas_per_frame) * per_cta + warp * cpw;
    const int ncell = max(0, min(cpw, hw - wc0));
    const unsigned long long src_mask = (1ull << g.sb) - 1ull;
    // lane q < ncell holds cell q's segment; rows are numbered 0..total-1 across the cells (pair rows, then the cell's own row)
    int4 seg = make_int4(0, 0, 0, 0);
    if (lane < ncell) seg = __ldg(cellseg + (long long)b * hw + wc0 + lane);
    const int len = lane < ncell ? seg.y - seg.x + 1 : 0;
    int first = len;
#pragma unroll
    for (int d = 1; d < FOW_MAX_CELLS; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, first, d); if (lane >= d) first += t; }
    const int total = __shfl_sync(0xffffffffu, first, FOW_MAX_CELLS - 1);
    first -= len;
    auto resolve = [&](int r, unsigned &row, float &m, int &meta) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < FOW_MAX_CELLS; ++k) { const int fk = __shfl_sync(0xffffffffu, first, k); if (k < ncell && r >= fk) q = k; }
        const int lo_q = __shfl_sync(0xffffffffu, seg.x, q), len_q = __shfl_sync(0xffffffffu, len, q);
        const int src_q = __shfl_sync(0xffffffffu, seg.z, q), has_q = __shfl_sync(0xffffffffu, seg.w, q);
        const int first_q = __shfl_sync(0xffffffffu, first, q);
        row = 0u; m = 0.f; meta = 0;
        if (r >= total) return;
        const int k = r - first_q;
        if (k == len_q - 1) { row = (unsigned)src_q; meta = 1 | (has_q << 1); return; }
        row = (unsigned)(__ldg(pairs + lo_q + k) & src_mask);
        m = (float)__ldg(mult + lo_q + k);
    };
    float acc[P * VEC], ssum[P * VEC], ssq[P * VEC];
#pragma unroll
    for (int k = 0; k < P * VEC; ++k) acc[k] = ssum[k] = ssq[k] = 0.f;
    float cnt = 0.f;
    const uint4 *fv = reinterpret_cast<const uint4 *>(feat);
    unsigned c_row, n_row; float c_m, n_m; int c_meta, n_meta;
    resolve(lane, n_row, n_m, n_meta);
    for (int r0 = 0; r0 < total; r0 += 32) {
        c_row = n_row; c_m = n_m; c_meta = n_meta;
        resolve(r0 + 32 + lane, n_row, n_m, n_meta);                  // the next 32 rows are requested before these are worked on
        const int nr = min(32, total - r0);
        for (int j0 = 0; j0 < nr; j0 += UN) {
            uint4 v[UN][P];
            float mm[UN];
            int mt[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int j = min(j0 + u, 31);
                const unsigned row = __shfl_sync(0xffffffffu, c_row, j);
                mm[u] = __shfl_sync(0xffffffffu, c_m, j);
                mt[u] = __shfl_sync(0xffffffffu, c_meta, j);
                if (j0 + u < nr) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int vi = lane + p * 32;
                        if (vi < row_vecs) v[u][p] = __ldg(fv + (long long)row * row_vecs + vi);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (j0 + u >= nr) break;
                if (!(mt[u] & 1)) {
                    cnt += mm[u];
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (lane + p * 32 < row_vecs) {
                            float x[VEC];
                            FoVec<XT>::unpack(v[u][p], x);
#pragma unroll
                            for (int k = 0; k < VEC; ++k) acc[p * VEC + k] = __fadd_rn(acc[p * VEC + k], __fmul_rn(mm[u], x[k]));
                        }
                    }
                } else {
                    // the cell's own row: (1 - r) * x + r * mean — mul, mul, add, each rounded (corresponder.py:272-273)
                    const bool has = (mt[u] & 2) != 0;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (lane + p * 32 < row_vecs) {
                            float x[VEC];
                            FoVec<XT>::unpack(v[u][p], x);
#pragma unroll
                            for (int k = 0; k < VEC; ++k) {
                                const float val = has ? __fadd_rn(__fmul_rn(one_minus, x[k]), __fmul_rn(ratio, __fdiv_rn(acc[p * VEC + k], cnt))) : x[k];
                                ssum[p * VEC + k] += val;
                                ssq[p * VEC + k] = fmaf(val, val, ssq[p * VEC + k]);
                                acc[p * VEC + k] = 0.f;
                            }
                        }
                    }
                    cnt = 0.f;
                }
            }
        }
    }
    if (lane == 0 && total) atomicAdd(rows_out, (unsigned long long)total);
    float *mine = s_stat + (size_t)warp * 2 * g.c;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int vi = lane + p * 32;
        if (vi < row_vecs) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) { mine[vi * VEC + k] = ssum[p * VEC + k]; mine[g.c + vi * VEC + k] = ssq[p * VEC + k]; }
        }
    }
    __syncthreads();
    for (int ch = tid; ch < g.c; ch += FOW_WARPS * 32) {
        double a = 0.0, a2 = 0.0;
#pragma unroll
        for (int w = 0; w < FOW_WARPS; ++w) { a += (double)s_stat[(size_t)w * 2 * g.c + ch]; a2 += (double)s_stat[(size_t)w * 2 * g.c + g.c + ch]; }
        atomicAdd(stats + ((long long)b * 4 + 2) * g.c + ch, a);
        atomicAdd(stats + ((long long)b * 4 + 3) * g.c + ch, a2);
    }
}

// content statistics: per (frame, channel) sums of x and x^2 over the h*w cells.  One CTA = `chunk` cells x one slab of 32 * VEC
// channels; a warp per cell, 16 bytes per lane, the eight warps' sums folded through shared memory.
template <typename XT>
__global__ void __launch_bounds__(256) k_fo_content_stats(const XT *__restrict__ feat, FoGeom g, int chunk, double *__restrict__ stats) {
    constexpr int VEC = FoVec<XT>::N;
    __shared__ double red[8][32 * VEC][2];
    const int hw = g.h * g.w, per = 32 * VEC;
    const int slabs = (g.c + per - 1) / per, chunks = (hw + chunk - 1) / chunk;
    int bid = blockIdx.x;
    const int ck = bid % chunks; bid /= chunks;
    const int sl = bid % slabs, b = bid / slabs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch0 = sl * per + lane * VEC;
    double s[VEC], s2[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = s2[k] = 0.0;
    const int c_hi = min(hw, (ck + 1) * chunk);
    if (ch0 < g.c) {
        for (int cell = ck * chunk + warp; cell < c_hi; cell += 32) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cell + 8 * u < c_hi) v[u] = __ldg(reinterpret_cast<const uint4 *>(feat + ((long long)b * hw + cell + 8 * u) * g.c + ch0));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (cell + 8 * u >= c_hi) break;
                float x[VEC];
                FoVec<XT>::unpack(v[u], x);
#pragma unroll
                for (int k = 0; k < VEC; ++k) { const double d = (double)x[k]; s[k] += d; s2[k] += d * d; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) { red[warp][lane * VEC + k][0] = s[k]; red[warp][lane * VEC + k][1] = s2[k]; }
    __syncthreads();
    for (int i = threadIdx.x; i < per; i += 256) {
        const int ch = sl * per + i;
        if (ch >= g.c) break;
        double a = 0.0, a2 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[w][i][0]; a2 += red[w][i][1]; }
        atomicAdd(stats + ((long long)b * 4 + 0) * g.c + ch, a);
        atomicAdd(stats + ((long long)b * 4 + 1) * g.c + ch, a2);
    }
}

// per (frame, channel): content mean / std, style std / mean — unbiased variance + 1e-5, sqrt (math_utils.py:39-47)
__global__ void __launch_bounds__(256) k_fo_coef(FoGeom g, const double *__restrict__ stats, float4 *__restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.B * g.c) return;
    const int b = i / g.c, ch = i - b * g.c;
    const double n = (double)(g.h * g.w);
    const double *s = stats + (long long)b * 4 * g.c + ch;
    const double sx = s[0], sxx = s[g.c], sb = s[2 * g.c], sbb = s[3 * g.c];
    const float mc = (float)(sx / n), ms = (float)(sb / n);
    const float sc = __fsqrt_rn(__fadd_rn((float)((sxx - sx * sx / n) / (n - 1.0)), 1e-5f));
    const float ss = __fsqrt_rn(__fadd_rn((float)((sbb - sb * sb / n) / (n - 1.0)), 1e-5f));
    coef[i] = make_float4(mc, sc, ss, ms);
}

// A thread keeps ONE channel vector (its stride over the vectors is a multiple of c / VEC), so the four coefficients of its VEC
// channels live in registers and are reloaded only when it crosses into the next frame; four vectors in flight, few threads (the
// 128 bytes of coefficients per thread have to be amortised over many 16-byte vectors).
template <typename XT>
__global__ void __launch_bounds__(256) k_fo_adain(const XT *__restrict__ feat, XT *__restrict__ out, FoGeom g, const float4 *__restrict__ coef) {
    constexpr int VEC = FoVec<XT>::N;
    const int hw = g.h * g.w, cv = g.c / VEC;
    const long long nvec = (long long)g.B * hw * cv, per_frame = (long long)hw * cv;
    const long long threads = (long long)gridDim.x * blockDim.x;
    const long long stride = threads / cv * cv;                      // host: threads >= cv
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    const int chv = (int)(t % cv);
    float4 q[VEC];
    int cur_b = -1;
    constexpr int UN = VEC == 8 ? 2 : 4;
    for (long long i0 = t; i0 < nvec; i0 += UN * stride) {
        uint4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (i0 + u * stride < nvec) v[u] = __ldg(reinterpret_cast<const uint4 *>(feat) + i0 + u * stride);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = i0 + u * stride;
            if (i >= nvec) break;
            const int b = (int)(i / per_frame);
            if (b != cur_b) {
                cur_b = b;
#pragma unroll
                for (int k = 0; k < VEC; ++k) q[k] = __ldg(coef + (long long)b * g.c + chv * VEC + k);
            }
            float x[VEC];
            FoVec<XT>::unpack(v[u], x);
#pragma unroll
            for (int k = 0; k < VEC; ++k) x[k] = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x[k], q[k].x), q[k].y), q[k].z), q[k].w);
            reinterpret_cast<uint4 *>(out)[i] = FoVec<XT>::pack(x);
        }
    }
}

// Workspace: the BUCKETS (what a reuse call needs: status, pair count, statistics, coefficients, per-cell segments, pairs,
// multiplicities — independent of the channel count and the feature dtype) first, the scratch of the bucketing pass (winners, codes,
// CUB) behind them.
struct FoLayout {
    int64_t status, npairs, stats, coef, cellseg, uniq, mult, bucket_bytes, winner, codes_a, codes_b, cub, cub_bytes, total;
    int sb, kb;
};

static inline int64_t fo_align(int64_t v) { return (v + 255) / 256 * 256; }
static inline int fo_bits(uint64_t v) { int b = 1; while (b < 63 && (v >> b)) ++b; return b; }     // bits needed for values 0..v

static int fo_layout(const srx_feature_args *a, FoLayout *L) {
    const int64_t npx = (int64_t)a->frames * a->height * a->width;
    SRX_REQUIRE(npx > 0 && npx < (1ll << 31), SRX_ERR_UNSUPPORTED, "id batch of %lld pixels (limit 2^31)", (long long)npx);
    const int64_t cells = (int64_t)a->batch * a->lat_h * a->lat_w;
    SRX_REQUIRE(cells < (1ll << 31), SRX_ERR_UNSUPPORTED, "%lld feature cells (limit 2^31)", (long long)cells);
    L->sb = fo_bits((uint64_t)(cells - 1));
    L->kb = fo_bits((uint64_t)(a->key_capacity - 1));
    size_t sort_bytes = 0, rle_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(e));
    e = cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr,
                                           (int *)nullptr, (int)npx);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cub::DeviceRunLengthEncode size query failed: %s", cudaGetErrorString(e));
    int64_t off = 0;
    L->status = off; off += 256;
    L->npairs = off; off += 256;
    L->stats = off; off = fo_align(off + (int64_t)a->batch * 4 * FO_MAXC * 8);
    L->coef = off; off = fo_align(off + (int64_t)a->batch * FO_MAXC * 16);
    L->cellseg = off; off = fo_align(off + cells * 16);
    L->uniq = off; off = fo_align(off + npx * 8);
    L->mult = off; off = fo_align(off + npx * 4);
    L->bucket_bytes = off;
    L->winner = off; off = fo_align(off + (int64_t)a->batch * a->map_height * a->map_width * 8);
    L->codes_a = off; off = fo_align(off + npx * 8);
    L->codes_b = off; off = fo_align(off + npx * 8);
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

extern "C" int64_t srx_feature_overlap_bucket_bytes(const srx_feature_args *a) {
    if (fo_validate(a)) return -1;
    FoLayout L;
    if (fo_layout(a, &L)) return -1;
    return L.bucket_bytes;
}

template <typename XT>
static int fo_run(const srx_feature_args *a, const FoLayout &L, const int *fmap_dev, cudaStream_t st) {
    char *ws = reinterpret_cast<char *>(a->workspace);
    FoGeom g{a->frames, a->height, a->width, a->batch, a->lat_h, a->lat_w, a->channels, a->map_height, a->map_width, (unsigned)a->key_capacity,
             L.sb, L.kb};
    const int64_t npx = (int64_t)a->frames * a->height * a->width;
    int *status = reinterpret_cast<int *>(ws + L.status);
    int *npairs = reinterpret_cast<int *>(ws + L.npairs);
    double *stats = reinterpret_cast<double *>(ws + L.stats);
    float4 *coef = reinterpret_cast<float4 *>(ws + L.coef);
    int4 *cellseg = reinterpret_cast<int4 *>(ws + L.cellseg);
    unsigned long long *uniq = reinterpret_cast<unsigned long long *>(ws + L.uniq);
    int *mult = reinterpret_cast<int *>(ws + L.mult);
    const int sms = srx_sm_count_cached();
    if (!a->reuse_buckets) {
        unsigned long long *winner = reinterpret_cast<unsigned long long *>(ws + L.winner);
        unsigned long long *ca = reinterpret_cast<unsigned long long *>(ws + L.codes_a), *cb = reinterpret_cast<unsigned long long *>(ws + L.codes_b);
        SRX_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)L.cellseg, st));        // status, pair count, statistics, coefficients
        SRX_CUDA_CHECK(cudaMemsetAsync(winner, 0, (size_t)(L.codes_a - L.winner), st));
        const long long nb = (npx + 255) / 256;
        const int grid = (int)(nb < (long long)sms * 16 ? nb : (long long)sms * 16);
        if (a->id_dtype == SRX_I32) k_fo_scan<int4><<<grid, 256, 0, st>>>(reinterpret_cast<const int4 *>(a->ids_dev), fmap_dev, g, winner, ca, status);
        else k_fo_scan<short4><<<grid, 256, 0, st>>>(reinterpret_cast<const short4 *>(a->ids_dev), fmap_dev, g, winner, ca, status);
        SRX_CUDA_CHECK(cudaGetLastError());
        size_t cub_bytes = (size_t)L.cub_bytes;
        SRX_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ws + L.cub, cub_bytes, ca, cb, (int)npx, 0, L.sb + L.kb + 1, st));
        cub_bytes = (size_t)L.cub_bytes;
        SRX_CUDA_CHECK(cub::DeviceRunLengthEncode::Encode(ws + L.cub, cub_bytes, cb, uniq, mult, npairs, (int)npx, st));
        // (the run of "no entry" codes sorts last and is never inside a key's segment)
        const int ncells = a->batch * a->lat_h * a->lat_w;
        k_fo_segments<<<(ncells + 255) / 256, 256, 0, st>>>(g, winner, uniq, npairs, cellseg);
        SRX_CUDA_CHECK(cudaGetLastError());
    } else {
        SRX_CUDA_CHECK(cudaMemsetAsync(ws + L.stats, 0, (size_t)(L.coef - L.stats), st));
        SRX_CUDA_CHECK(cudaMemsetAsync(status + FO_ST_ROWS, 0, 8, st));
    }
    const XT *feat = reinterpret_cast<const XT *>(a->feat_dev);
    const int hw = a->lat_h * a->lat_w;
    {
        const int per = 32 * FoVec<XT>::N, slabs = (a->channels + per - 1) / per;
        const int chunk = hw >= 1024 ? 128 : 32;
        k_fo_content_stats<XT><<<a->batch * slabs * ((hw + chunk - 1) / chunk), 256, 0, st>>>(feat, g, chunk, stats);
    }
    const int row_bytes = a->channels * (int)sizeof(XT);
    const int stage_bytes = (row_bytes + 127) & ~127;
    int nst = FO_RING_BYTES / stage_bytes;
    nst = nst < FO_MIN_STAGES ? FO_MIN_STAGES : (nst > FO_MAX_STAGES ? FO_MAX_STAGES : nst);
    const int smem = nst * stage_bytes + 2 * nst * 8 + FO_CELLS * 16 + nst * 4 + 64;
    unsigned long long *rows = reinterpret_cast<unsigned long long *>(status + FO_ST_ROWS);
    const float one_minus = (float)(1.0 - (double)a->ratio);
    if (row_bytes <= 2048) {
        const int smem_w = FOW_WARPS * 2 * a->channels * 4;
        int cpw = FOW_CELLS;
        if (const char *e = getenv("SRX_FO_CELLS")) { cpw = atoi(e); cpw = cpw < 1 ? 1 : (cpw > FOW_MAX_CELLS ? FOW_MAX_CELLS : cpw); }
        const int per_cta = FOW_WARPS * cpw;
        const int grid_w = a->batch * ((hw + per_cta - 1) / per_cta);
#define FO_LAUNCH_WARP(P_) k_fo_style_warp<XT, P_><<<grid_w, FOW_WARPS * 32, smem_w, st>>>(feat, g, cellseg, uniq, mult, a->ratio, one_minus, cpw, stats, rows)
        const int nvec = row_bytes / 16;
        if (nvec <= 32) FO_LAUNCH_WARP(1);
        else if (nvec <= 64) FO_LAUNCH_WARP(2);
        else if (nvec <= 96) FO_LAUNCH_WARP(3);
        else FO_LAUNCH_WARP(4);
#undef FO_LAUNCH_WARP
    } else {
        const int grid_style = a->batch * ((hw + FO_CELLS - 1) / FO_CELLS);
#define FO_LAUNCH_STYLE(KV_)                                                                                                        \
    do {                                                                                                                            \
        static bool configured = false;                                                                                             \
        if (!configured) {                                                                                                          \
            SRX_CUDA_CHECK(cudaFuncSetAttribute(k_fo_style<XT, KV_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                                FO_RING_BYTES + FO_MIN_STAGES * FO_MAXC * 4 + 1024));                               \
            configured = true;                                                                                                      \
        }                                                                                                                           \
        k_fo_style<XT, KV_><<<grid_style, FO_THREADS, smem, st>>>(feat, g, cellseg, uniq, mult, a->ratio, one_minus, nst, stats, rows); \
    } while (0)
        if (row_bytes <= 2 * FO_CONS * 16) FO_LAUNCH_STYLE(2);
        else FO_LAUNCH_STYLE(3);
#undef FO_LAUNCH_STYLE
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    k_fo_coef<<<(a->batch * a->channels + 255) / 256, 256, 0, st>>>(g, stats, coef);
    const long long nvec = (long long)a->batch * hw * a->channels / FoVec<XT>::N;
    long long nb2 = (nvec + 4 * 256 - 1) / (4 * 256);                   // >= 4 vectors per thread, one resident wave
    static int occ_adain = 0;
    if (!occ_adain) {
        SRX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_adain, k_fo_adain<XT>, 256, 0));
        if (occ_adain < 1) occ_adain = 1;
    }
    nb2 = nb2 < (long long)sms * occ_adain ? nb2 : (long long)sms * occ_adain;
    const long long cvb = (a->channels / FoVec<XT>::N + 255) / 256;       // at least c / VEC threads
    k_fo_adain<XT><<<(int)(nb2 > cvb ? nb2 : cvb), 256, 0, st>>>(feat, reinterpret_cast<XT *>(a->out_dev), g, coef);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

extern "C" int srx_feature_overlap(const srx_feature_args *a, void *stream) {
    int rc = fo_validate(a);
    if (rc) return rc;
    SRX_REQUIRE(a->feat_dev && a->out_dev && a->workspace, SRX_ERR_INVALID, "null buffer");
    SRX_REQUIRE(a->reuse_buckets || (a->ids_dev && a->frame_map_dev), SRX_ERR_INVALID, "null id buffer");
    SRX_REQUIRE(((reinterpret_cast<uintptr_t>(a->feat_dev) | reinterpret_cast<uintptr_t>(a->out_dev)) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, SRX_ERR_INVALID,
                "features and output must be 16-byte aligned, the workspace 256-byte aligned");
    FoLayout L;
    if ((rc = fo_layout(a, &L))) return rc;
    const int64_t need = a->reuse_buckets ? L.bucket_bytes : L.total;
    SRX_REQUIRE(a->workspace_bytes >= need, SRX_ERR_INVALID, "workspace too small: %lld < %lld", (long long)a->workspace_bytes, (long long)need);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->x_dtype == SRX_F32) return fo_run<float>(a, L, a->frame_map_dev, st);
    if (a->x_dtype == SRX_F16) return fo_run<__half>(a, L, a->frame_map_dev, st);
    return fo_run<__nv_bfloat16>(a, L, a->frame_map_dev, st);
}

// device-side failures of the bucketing pass behind this workspace (syncs the stream)
extern "C" int srx_feature_overlap_check(const srx_feature_args *a, void *stream) {
    SRX_REQUIRE(a && a->workspace, SRX_ERR_INVALID, "null argument");
    int host[4] = {0, 0, 0, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SRX_CUDA_CHECK(cudaMemcpyAsync(host, a->workspace, sizeof(host), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (host[FO_ST_RANGE] & 1) return srx_set_error(SRX_ERR_INDEX, "index out of range: an id pixel maps outside the up-sampled features or names a frame outside the batch");
    if (host[FO_ST_RANGE] & 2) return srx_set_error(SRX_ERR_KEY_RANGE, "a vertex id fell outside the key capacity (%lld)", (long long)a->key_capacity);
    return SRX_OK;
}

extern "C" int64_t srx_feature_overlap_rows(const srx_feature_args *a, void *stream) {
    if (!a || !a->workspace) return -1;
    unsigned long long rows = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(&rows, reinterpret_cast<char *>(a->workspace) + FO_ST_ROWS * 4, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    return (int64_t)rows;
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
