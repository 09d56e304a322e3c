// srx_keying.cu — API-parity kernels around the keying stage:
//   srx_vertex_screen_info      IDMap.create_vertex_screen_info         source/engine/static/corrmap.py:220-280
//   srx_group_by_then_average   tensor_group_by_then_average            source/common_utils/math_utils.py:86-161
// The overlap step itself never materialises these arrays (srx_overlap.cu keys on the fly); they exist so that the
// reference's intermediate tensors can be produced — and compared bit for bit — on the GPU.
#include "srx_common.cuh"
#include <vector>

#define VSI_THREADS 256
#define VSI_ITERS 16
#define VSI_TILE (VSI_THREADS * VSI_ITERS)

template <typename IdT>
__global__ void __launch_bounds__(VSI_THREADS) k_vsi_count(const IdT *__restrict__ ids, long long npx,
                                                            unsigned int *__restrict__ block_counts) {
    const long long base = (long long)blockIdx.x * VSI_TILE;
    unsigned int cnt = 0;
#pragma unroll 4
    for (int j = 0; j < VSI_ITERS; ++j) {
        const long long i = base + (long long)j * VSI_THREADS + threadIdx.x;
        if (i < npx) cnt += id_valid(load_id(ids + i)) ? 1u : 0u;
    }
    __shared__ unsigned int sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh, cnt);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = sh;
}

// exclusive scan of the per-tile counts (single CTA; tiles are few: F*H*W/4096)
__global__ void __launch_bounds__(1024) k_vsi_scan(const unsigned int *__restrict__ block_counts,
                                                    unsigned long long *__restrict__ block_offsets, int nblocks,
                                                    unsigned long long *__restrict__ total) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry, chunk_total;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < nblocks ? block_counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = warp_tot[lane];
            unsigned long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;  // exclusive offset of each warp inside the chunk
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        if (i < nblocks) block_offsets[i] = carry + warp_tot[wid] + (incl - v);
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

template <typename IdT>
__global__ void __launch_bounds__(VSI_THREADS) k_vsi_write(const IdT *__restrict__ ids, long long npx, int H, int W,
                                                            const int *__restrict__ frame_values,
                                                            const unsigned long long *__restrict__ block_offsets,
                                                            float *__restrict__ out) {
    __shared__ unsigned int warp_cnt[VSI_THREADS / 32];
    __shared__ unsigned long long run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * VSI_TILE;
    if (threadIdx.x == 0) run = block_offsets[blockIdx.x];
    __syncthreads();
    const float fH = __int2float_rn(H), fW = __int2float_rn(W);
    for (int j = 0; j < VSI_ITERS; ++j) {
        const long long i = base + (long long)j * VSI_THREADS + threadIdx.x;
        IdPx p{0, 0, 0, 0};
        bool valid = false;
        if (i < npx) {
            p = load_id(ids + i);
            valid = id_valid(p);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        unsigned int before = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < VSI_THREADS / 32; ++k) {
            const unsigned int c = warp_cnt[k];
            if (k < wid) before += c;
            tot += c;
        }
        if (valid) {
            const unsigned long long row = run + before + __popc(bal & ((1u << lane) - 1u));
            const int x = (int)(i % W);
            const long long t = i / W;
            const int y = (int)(t % H);
            const int g = (int)(t / H);
            float *o = out + row * 7;
            o[0] = __int2float_rn(p.s);
            o[1] = __int2float_rn(p.m);
            o[2] = __int2float_rn(p.i);
            o[3] = __int2float_rn(p.v);
            o[4] = __fdiv_rn(__int2float_rn(x), fH);  // x / height (sic, corrmap.py:239)
            o[5] = __fdiv_rn(__int2float_rn(y), fW);  // y / width  (sic, corrmap.py:249)
            o[6] = __int2float_rn(frame_values[g]);
        }
        __syncthreads();
        if (threadIdx.x == 0) run += tot;
        __syncthreads();
    }
}

template <typename IdT>
static int vsi_impl(const void *ids_dev, int F, int H, int W, const int32_t *frame_values_host, float *out_dev,
                    int64_t *n_rows, cudaStream_t st) {
    const long long npx = (long long)F * H * W;
    const int nblocks = (int)((npx + VSI_TILE - 1) / VSI_TILE);
    char *scratch = nullptr;
    const size_t bytes = (size_t)nblocks * (sizeof(unsigned int) + sizeof(unsigned long long)) + 16 + (size_t)F * sizeof(int) + 64;
    SRX_CUDA_CHECK(cudaMalloc(&scratch, bytes));
    unsigned long long *offsets = reinterpret_cast<unsigned long long *>(scratch);
    unsigned long long *total = offsets + nblocks;
    unsigned int *counts = reinterpret_cast<unsigned int *>(total + 1);
    int *fv = reinterpret_cast<int *>(counts + nblocks + (nblocks & 1));
    std::vector<int> fvh((size_t)F);
    for (int g = 0; g < F; ++g) fvh[g] = frame_values_host ? frame_values_host[g] : g;
    cudaError_t e = cudaMemcpyAsync(fv, fvh.data(), (size_t)F * sizeof(int), cudaMemcpyHostToDevice, st);
    const IdT *ids = reinterpret_cast<const IdT *>(ids_dev);
    if (e == cudaSuccess) {
        k_vsi_count<IdT><<<nblocks, VSI_THREADS, 0, st>>>(ids, npx, counts);
        k_vsi_scan<<<1, 1024, 0, st>>>(counts, offsets, nblocks, total);
        k_vsi_write<IdT><<<nblocks, VSI_THREADS, 0, st>>>(ids, npx, H, W, fv, offsets, out_dev);
        e = cudaGetLastError();
    }
    unsigned long long tot = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&tot, total, sizeof(tot), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "vertex_screen_info failed: %s", cudaGetErrorString(e));
    *n_rows = (int64_t)tot;
    return SRX_OK;
}

extern "C" int srx_vertex_screen_info(const void *ids_dev, int id_dtype, int frames, int height, int width,
                                      const int32_t *frame_values_host, float *out_dev, int64_t *n_rows, void *stream) {
    SRX_REQUIRE(ids_dev && out_dev && n_rows, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(frames > 0 && height > 0 && width > 0, SRX_ERR_INVALID, "non-positive dimension");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (id_dtype == SRX_I32) return vsi_impl<int4>(ids_dev, frames, height, width, frame_values_host, out_dev, n_rows, st);
    if (id_dtype == SRX_I16) return vsi_impl<short4>(ids_dev, frames, height, width, frame_values_host, out_dev, n_rows, st);
    return srx_set_error(SRX_ERR_INVALID, "id dtype must be int32 or int16");
}

// -----------------------------------------------------------------------------------------------------------------
// tensor_group_by_then_average with a dense slot table
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gba_accum(const float *__restrict__ values, const float *__restrict__ keys,
                                                    long long n, int C, float *__restrict__ acc, float *__restrict__ cnt,
                                                    long long kcap, int *__restrict__ bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float kf = keys[i];
        const long long s = (long long)kf;
        if (!(kf >= 0.f) || (float)s != kf || s >= kcap) { atomicOr(bad, 1); continue; }
        const float *v = values + i * C;
        if (C == 4) red_add_f32x4(acc + s * 4, v[0], v[1], v[2], v[3]);
        else for (int ch = 0; ch < C; ++ch) red_add_f32(acc + s * C + ch, v[ch]);
        red_add_f32(cnt + s, 1.f);
    }
}

__global__ void __launch_bounds__(256) k_gba_expand(const float *__restrict__ keys, long long n, int C,
                                                     const float *__restrict__ acc, const float *__restrict__ cnt,
                                                     long long kcap, float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / C;
        const int ch = (int)(i % C);
        const float kf = keys[row];
        const long long s = (long long)kf;
        if (!(kf >= 0.f) || s >= kcap) { out[i] = __int_as_float(0x7fc00000); continue; }
        out[i] = __fdiv_rn(acc[s * C + ch], cnt[s]);  // average_values = sum_values / counts (math_utils.py:153)
    }
}

extern "C" int srx_group_by_then_average(const float *values_dev, const float *keys_dev, int64_t n, int channels,
                                         float *out_dev, float *workspace_dev, int64_t key_capacity, void *stream) {
    SRX_REQUIRE(values_dev && keys_dev && out_dev && workspace_dev, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(n >= 0 && channels > 0 && key_capacity > 0, SRX_ERR_INVALID, "bad sizes");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n == 0) return SRX_OK;
    float *acc = workspace_dev;
    float *cnt = acc + key_capacity * channels;
    int *bad = nullptr;
    SRX_CUDA_CHECK(cudaMalloc(&bad, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(workspace_dev, 0, (size_t)key_capacity * (channels + 1) * sizeof(float), st);
    int bad_h = 0;
    if (e == cudaSuccess) {
        const int sms = srx_sm_count_cached();
        long long nb = (n + 255) / 256;
        const int grid = (int)(nb < (long long)sms * 8 ? nb : (long long)sms * 8);
        k_gba_accum<<<grid, 256, 0, st>>>(values_dev, keys_dev, n, channels, acc, cnt, key_capacity, bad);
        long long nb2 = (n * channels + 255) / 256;
        const int grid2 = (int)(nb2 < (long long)sms * 8 ? nb2 : (long long)sms * 8);
        k_gba_expand<<<grid2, 256, 0, st>>>(keys_dev, n, channels, acc, cnt, key_capacity, out_dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(bad);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "group_by_then_average failed: %s", cudaGetErrorString(e));
    if (bad_h) return srx_set_error(SRX_ERR_KEY_RANGE, "group keys must be non-negative integers below key_capacity=%lld", (long long)key_capacity);
    return SRX_OK;
}
