// srx_group.cu — same-key broadcast initialisers: every pixel that shows the same surface
// point receives the same random vector.
//
// Replaces (reference paths relative to /root/reference):
//   tensor_group_by_then_randn_init        source/common_utils/math_utils.py:164-229
//   CreateNoiseSequenceFromIdMap.__call__  source/comfyUI/stable_rendering/_nodes/loaders.py:193-271
//
// The reference sorts the N keys (`unique(return_inverse=True)`), draws one random row per unique key with torch.randn
// and expands through the inverse.  Keys here are texel ids in a bounded range, so the sort becomes a dense presence
// table + exclusive scan: table[key] = rank of the key among the sorted unique keys — exactly `inverse`, bit for bit.
// The random rows themselves are still drawn by torch (same call, same generator, same shape as the reference, so the
// values are identical on the same device); everything else — ranking, expansion, the scatter into the noise frames
// and the 8x down-sampling — runs here, without the [N,7] entry list and without the full-resolution noise tensors.
#include "srx_common.cuh"

#define GR_THREADS 256
#define GR_TILE 2048   // table entries per scan tile

// ---------------------------------------------------------------------------------------------------------------
// presence
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GR_THREADS) k_gr_mark_keys(const float *__restrict__ keys, long long n, long long kcap,
                                                              int *__restrict__ table, int *__restrict__ bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float kf = keys[i];
        const long long s = (long long)kf;
        if (!(kf >= 0.f) || (float)s != kf || s >= kcap) { atomicOr(bad, 1); continue; }
        table[s] = 1;
    }
}

template <typename IdT>
__global__ void __launch_bounds__(GR_THREADS) k_gr_mark_ids(const IdT *__restrict__ ids, long long npx, long long kcap,
                                                             int *__restrict__ table, int *__restrict__ bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        const IdPx p = load_id(ids + i);
        if (!id_valid(p)) continue;
        const long long s = vertex_slot(p.v);
        if (s < 0 || s >= kcap) { atomicOr(bad, 1); continue; }
        table[s] = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan of the 0/1 table, in place: tile sums -> scan of the tile sums (one CTA) -> per-tile scan + offset.
// Absent keys get rank -1.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GR_THREADS) k_gr_tile_sums(const int *__restrict__ table, long long kcap, int *__restrict__ tile_sum) {
    const long long base = (long long)blockIdx.x * GR_TILE;
    int c = 0;
    for (int j = threadIdx.x; j < GR_TILE; j += GR_THREADS) {
        const long long i = base + j;
        if (i < kcap) c += table[i];
    }
    __shared__ int sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = sh;
}

__global__ void __launch_bounds__(1024) k_gr_scan_tiles(int *__restrict__ tile_sum, int ntiles, int *__restrict__ total) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tile_sum[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (wid ? warp_tot[wid - 1] : 0) + incl - v;
        if (i < ntiles) tile_sum[i] = before;   // exclusive offset of the tile
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(GR_THREADS) k_gr_rank_tiles(int *__restrict__ table, long long kcap, const int *__restrict__ tile_off) {
    // GR_TILE / GR_THREADS = 8 consecutive entries per thread
    const long long base = (long long)blockIdx.x * GR_TILE + (long long)threadIdx.x * (GR_TILE / GR_THREADS);
    int v[GR_TILE / GR_THREADS];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < GR_TILE / GR_THREADS; ++j) {
        v[j] = (base + j < kcap) ? table[base + j] : 0;
        mine += v[j];
    }
    __shared__ int warp_tot[GR_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int before = tile_off[blockIdx.x] + incl - mine;
    for (int k = 0; k < wid; ++k) before += warp_tot[k];
#pragma unroll
    for (int j = 0; j < GR_TILE / GR_THREADS; ++j) {
        if (base + j < kcap) table[base + j] = v[j] ? before : -1;
        before += v[j];
    }
}

static int gr_rank_table(int *table, long long kcap, int *scratch /* ntiles + 2 ints */, int64_t *n_unique, int *bad_dev, cudaStream_t st) {
    const int ntiles = (int)((kcap + GR_TILE - 1) / GR_TILE);
    k_gr_tile_sums<<<ntiles, GR_THREADS, 0, st>>>(table, kcap, scratch);
    k_gr_scan_tiles<<<1, 1024, 0, st>>>(scratch, ntiles, scratch + ntiles);
    k_gr_rank_tiles<<<ntiles, GR_THREADS, 0, st>>>(table, kcap, scratch);
    SRX_CUDA_CHECK(cudaGetLastError());
    int host[2] = {0, 0};
    SRX_CUDA_CHECK(cudaMemcpyAsync(&host[0], scratch + ntiles, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaMemcpyAsync(&host[1], bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRX_CUDA_CHECK(cudaStreamSynchronize(st));
    if (host[1]) return srx_set_error(SRX_ERR_KEY_RANGE, "group keys must be non-negative integers below key_capacity=%lld", kcap);
    *n_unique = host[0];
    return SRX_OK;
}

// In-place exclusive scan of a 0/1 int array for other translation units (srx_legacy.cu): entry -> its rank among the set
// entries, -1 where clear.  scratch: srx_flags_scan_scratch_ints(n) ints.  Syncs.
long long srx_flags_scan_scratch_ints(long long n) { return (n + GR_TILE - 1) / GR_TILE + 4; }
int srx_flags_to_ranks(int *flags, long long n, int *scratch, int64_t *total, cudaStream_t st) {
    const int ntiles = (int)((n + GR_TILE - 1) / GR_TILE);
    SRX_CUDA_CHECK(cudaMemsetAsync(scratch + ntiles, 0, 4 * sizeof(int), st));
    return gr_rank_table(flags, n, scratch, total, scratch + ntiles + 1, st);
}

extern "C" int64_t srx_group_rank_workspace_ints(int64_t key_capacity) {
    return key_capacity + (key_capacity + GR_TILE - 1) / GR_TILE + 4;
}

__global__ void __launch_bounds__(GR_THREADS) k_gr_lookup(const float *__restrict__ keys, long long n, const int *__restrict__ table,
                                                           int *__restrict__ rank_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        rank_out[i] = table[(long long)keys[i]];
}

static int gr_grid(long long work) {
    const long long nb = (work + GR_THREADS - 1) / GR_THREADS;
    const long long cap = (long long)srx_sm_count_cached() * 8;
    return (int)(nb < 1 ? 1 : (nb < cap ? nb : cap));
}

// table_ws: srx_group_rank_workspace_ints(key_capacity) ints; on return table_ws[k] = rank of key k among the sorted
// unique keys (or -1), rank_out[i] = rank of keys[i] (= torch.unique's inverse), *n_unique = number of unique keys.  Syncs.
extern "C" int srx_group_rank(const float *keys_dev, int64_t n, int64_t key_capacity, int32_t *table_ws, int32_t *rank_out,
                              int64_t *n_unique, void *stream) {
    SRX_REQUIRE(keys_dev && table_ws && n_unique, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(n >= 0 && key_capacity > 0 && key_capacity < (1ll << 31), SRX_ERR_INVALID, "bad sizes");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ntiles = (int)((key_capacity + GR_TILE - 1) / GR_TILE);
    int *scratch = table_ws + key_capacity;
    int *bad = scratch + ntiles + 1;
    SRX_CUDA_CHECK(cudaMemsetAsync(table_ws, 0, (size_t)srx_group_rank_workspace_ints(key_capacity) * sizeof(int), st));
    if (n > 0) k_gr_mark_keys<<<gr_grid(n), GR_THREADS, 0, st>>>(keys_dev, n, key_capacity, table_ws, bad);
    int rc = gr_rank_table(table_ws, key_capacity, scratch, n_unique, bad, st);
    if (rc) return rc;
    if (rank_out && n > 0) {
        k_gr_lookup<<<gr_grid(n), GR_THREADS, 0, st>>>(keys_dev, n, table_ws, rank_out);
        SRX_CUDA_CHECK(cudaGetLastError());
    }
    return SRX_OK;
}

// the same table built straight from id buffers (key = float32(vertexID) of every valid pixel).  Syncs.
extern "C" int srx_ids_rank_table(const void *ids_dev, int id_dtype, int frames, int height, int width, int64_t key_capacity,
                                  int32_t *table_ws, int64_t *n_unique, void *stream) {
    SRX_REQUIRE(ids_dev && table_ws && n_unique, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0 && key_capacity > 0 && key_capacity < (1ll << 31), SRX_ERR_INVALID, "bad sizes");
    SRX_REQUIRE(id_dtype == SRX_I32 || id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ntiles = (int)((key_capacity + GR_TILE - 1) / GR_TILE);
    int *scratch = table_ws + key_capacity;
    int *bad = scratch + ntiles + 1;
    const long long npx = (long long)frames * height * width;
    SRX_CUDA_CHECK(cudaMemsetAsync(table_ws, 0, (size_t)srx_group_rank_workspace_ints(key_capacity) * sizeof(int), st));
    if (id_dtype == SRX_I32)
        k_gr_mark_ids<int4><<<gr_grid(npx), GR_THREADS, 0, st>>>(reinterpret_cast<const int4 *>(ids_dev), npx, key_capacity, table_ws, bad);
    else
        k_gr_mark_ids<short4><<<gr_grid(npx), GR_THREADS, 0, st>>>(reinterpret_cast<const short4 *>(ids_dev), npx, key_capacity, table_ws, bad);
    return gr_rank_table(table_ws, key_capacity, scratch, n_unique, bad, st);
}

// out[i, :] = table[rank[i], :]   — `random_values[inverse_indices]` (math_utils.py:224)
__global__ void __launch_bounds__(GR_THREADS) k_gr_broadcast(const float *__restrict__ table, const int *__restrict__ rank, long long n,
                                                              int C, float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / C;
        out[i] = table[(long long)rank[row] * C + (int)(i - row * C)];
    }
}

extern "C" int srx_group_broadcast(const float *table_dev, const int32_t *rank_dev, int64_t n, int channels, float *out_dev, void *stream) {
    SRX_REQUIRE(table_dev && rank_dev && out_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(n >= 0 && channels > 0, SRX_ERR_INVALID, "bad sizes");
    if (n == 0) return SRX_OK;
    k_gr_broadcast<<<gr_grid(n * channels), GR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table_dev, rank_dev, n, channels, out_dev);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// CreateNoiseSequenceFromIdMap, fused: value(f, c, y, x) = key_noise[rank(id[f,y,x])][c] where the pixel has an id,
// else base[c, y, x]; then the node's down-sampling.  Two tensors (latent, noise) share one pass over the ids.
// ---------------------------------------------------------------------------------------------------------------
template <typename IdT>
__device__ __forceinline__ int gr_pixel_rank(const IdT *ids, long long px, const int *__restrict__ table) {
    const IdPx p = load_id(ids + px);
    if (!id_valid(p)) return -1;
    return table[vertex_slot(p.v)];
}

// Which entry ends up in full-resolution pixel (Y, X) of latent frame f?  The node writes `latent[frame, :, sy, sx] = row`
// for every entry with sx = trunc(fl32(x / H) * S), sy = trunc(fl32(y / W) * S) (loaders.py:218-219 with corrmap.py:239,249);
// duplicates resolve to the LAST entry in (id frame, y, x) order.  With an id map of the working size the candidates are the
// pixel (Y, X) itself; a larger map sends several pixels to one target, a smaller one leaves holes (the base draw stays);
// several id frames may write one latent frame (the chain inv_frame[f] -> prev_frame[g] -> ..., latest first).
struct NoiseGeom {
    int F, H, W, S;
    const int *inv_frame, *prev_frame;
};
__device__ __forceinline__ int gr_axis_map(int p, int div, int S) { return (int)__fmul_rn(__fdiv_rn((float)p, (float)div), (float)S); }
// pixels p in [0, n) with gr_axis_map(p, div, S) == T form one run [lo, hi] (the map is monotone); hi < lo = none
__device__ __forceinline__ void gr_axis_range(int T, int n, int div, int S, int &lo, int &hi) {
    int p = (int)(((long long)T * div) / S) - 2;
    if (p < 0) p = 0;
    while (p < n && gr_axis_map(p, div, S) < T) ++p;
    lo = p;
    while (p < n && gr_axis_map(p, div, S) == T) ++p;
    hi = p - 1;
}
template <typename IdT>
__device__ __forceinline__ int gr_target_rank(const IdT *ids, const NoiseGeom &g, const int *__restrict__ table, int f, int Y, int X) {
    int ylo = Y, yhi = Y, xlo = X, xhi = X;
    if (g.H != g.S || g.W != g.S) {
        gr_axis_range(Y, g.H, g.W, g.S, ylo, yhi);      // y / width
        gr_axis_range(X, g.W, g.H, g.S, xlo, xhi);      // x / height
        if (yhi < ylo || xhi < xlo) return -1;
    }
    for (int gi = g.inv_frame[f]; gi >= 0; gi = g.prev_frame ? g.prev_frame[gi] : -1) {
        for (int y = yhi; y >= ylo; --y)
            for (int x = xhi; x >= xlo; --x) {
                const int r = gr_pixel_rank(ids, ((long long)gi * g.H + y) * g.W + x, table);
                if (r >= 0) return r;
            }
    }
    return -1;
}

// 'nearest': F.interpolate(size = (S/8, S/8)) keeps source pixel (8i, 8j)   (loaders.py:252-255)
template <typename IdT>
__global__ void __launch_bounds__(GR_THREADS) k_noise_nearest(const IdT *__restrict__ ids, NoiseGeom g,
                                                               const int *__restrict__ table, const float *__restrict__ key_a,
                                                               const float *__restrict__ key_b, const float *__restrict__ base_a,
                                                               const float *__restrict__ base_b, float *__restrict__ out_a,
                                                               float *__restrict__ out_b) {
    const int S = g.S, h = S / 8, w = S / 8;
    const long long total = (long long)g.F * h * w;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % w);
        const int i = (int)((t / w) % h);
        const int f = (int)(t / ((long long)w * h));
        const long long src = (long long)(8 * i) * S + 8 * j;
        const int r = gr_target_rank(ids, g, table, f, 8 * i, 8 * j);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long o = (((long long)f * 4 + c) * h + i) * w + j;
            out_a[o] = r >= 0 ? key_a[(long long)r * 4 + c] : base_a[(long long)c * S * S + src];
            out_b[o] = r >= 0 ? key_b[(long long)r * 4 + c] : base_b[(long long)c * S * S + src];
        }
    }
}

// 'mean' / 'max' / 'min': the node views the full-resolution tensor as [-1, 4, 8, 8] — 256 CONSECUTIVE floats of a row,
// not an 8x8 block — and reduces dims (1, 2): out[chunk, k] = op over a < 4, b < 8 of row[chunk*256 + a*64 + b*8 + k]
// (loaders.py:257-268).  The result holds F*4*S*S/32 floats, which the node then views as [-1, 4, S/8, S/8] (2F frames).
template <typename IdT>
__global__ void __launch_bounds__(GR_THREADS) k_noise_pool(const IdT *__restrict__ ids, NoiseGeom g,
                                                            const int *__restrict__ table, const float *__restrict__ key_b,
                                                            const float *__restrict__ base_b, float *__restrict__ out_b, int op) {
    const int S = g.S;
    const long long total = (long long)g.F * 4 * S * S / 32;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t & 7);
        const long long first = (t >> 3) * 256;            // flat index of the chunk's first float in [F,4,S,S]
        const int x0 = (int)(first % S);
        const int y = (int)((first / S) % S);
        const int c = (int)((first / ((long long)S * S)) % 4);
        const int f = (int)(first / ((long long)S * S * 4));
        float acc = op == 1 ? 0.f : (op == 2 ? -INFINITY : INFINITY);
        for (int ab = 0; ab < 32; ++ab) {
            const int x = x0 + ab * 8 + k;
            const int r = gr_target_rank(ids, g, table, f, y, x);
            const float v = r >= 0 ? key_b[(long long)r * 4 + c] : base_b[((long long)c * S + y) * S + x];
            acc = op == 1 ? acc + v : (op == 2 ? fmaxf(acc, v) : fminf(acc, v));
        }
        out_b[t] = op == 1 ? acc / 32.f : acc;
    }
}

extern "C" int srx_noise_from_ids(const srx_noise_args *a, void *stream) {
    SRX_REQUIRE(a && a->ids_dev && a->inv_frame_dev && a->rank_table_dev && a->key_noise_dev && a->base_noise_dev && a->noise_out_dev,
                SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(a->id_dtype == SRX_I32 || a->id_dtype == SRX_I16, SRX_ERR_INVALID, "id dtype must be int32 or int16");
    const int S = a->work_size > 0 ? a->work_size : a->height;
    SRX_REQUIRE(a->frames > 0 && a->height > 0 && a->width > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(S % 256 == 0, SRX_ERR_INVALID, "the working size must be a multiple of 256 (the node uses 512 / 1024)");
    SRX_REQUIRE(a->work_size > 0 || a->height == a->width, SRX_ERR_INVALID, "work_size == 0 means a square id map of the working size");
    SRX_REQUIRE(a->mode >= 0 && a->mode <= 3, SRX_ERR_INVALID, "mode must be 0 nearest, 1 mean, 2 max, 3 min");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    NoiseGeom g{a->frames, a->height, a->width, S, a->inv_frame_dev, a->prev_frame_dev};
    if (a->mode == 0) {
        SRX_REQUIRE(a->key_latent_dev && a->base_latent_dev && a->latent_out_dev, SRX_ERR_INVALID, "nearest mode needs the latent tensors too");
        const long long total = (long long)g.F * (S / 8) * (S / 8);
        if (a->id_dtype == SRX_I32)
            k_noise_nearest<int4><<<gr_grid(total), GR_THREADS, 0, st>>>(reinterpret_cast<const int4 *>(a->ids_dev), g, a->rank_table_dev,
                a->key_latent_dev, a->key_noise_dev, a->base_latent_dev, a->base_noise_dev, a->latent_out_dev, a->noise_out_dev);
        else
            k_noise_nearest<short4><<<gr_grid(total), GR_THREADS, 0, st>>>(reinterpret_cast<const short4 *>(a->ids_dev), g, a->rank_table_dev,
                a->key_latent_dev, a->key_noise_dev, a->base_latent_dev, a->base_noise_dev, a->latent_out_dev, a->noise_out_dev);
    } else {
        const long long total = (long long)g.F * 4 * S * S / 32;
        if (a->id_dtype == SRX_I32)
            k_noise_pool<int4><<<gr_grid(total), GR_THREADS, 0, st>>>(reinterpret_cast<const int4 *>(a->ids_dev), g, a->rank_table_dev,
                a->key_noise_dev, a->base_noise_dev, a->noise_out_dev, a->mode);
        else
            k_noise_pool<short4><<<gr_grid(total), GR_THREADS, 0, st>>>(reinterpret_cast<const short4 *>(a->ids_dev), g, a->rank_table_dev,
                a->key_noise_dev, a->base_noise_dev, a->noise_out_dev, a->mode);
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
