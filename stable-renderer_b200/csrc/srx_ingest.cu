// srx_ingest.cu — G-buffer frame ingest: one frame of attachments -> the batch tensors the overlap step and the bake read.
//
// Replaces RenderManager._save_frame_data (source/engine/managers/renderManager.py:877-948) and the "closer pixel wins"
// merge of identical-G-buffer tasks (renderManager.py:121-133).  The reference reads each of the six attachments back
// with Texture.tensor(update=True, flip=True) (map, Memcpy2D, device-wide sync, flip: texture.py:221-254), clones it,
// slices it and torch.cat()s it onto a growing batch; the noise attachment additionally goes through mask mixing, a
// 64-pixel mean, and AdaIN against the raw attachment.  Here: one pass over the frame writes every output straight into
// slot `frame_slot` of preallocated batch tensors with the row flip fused, and reduces the statistics AdaIN needs; a
// second, tiny kernel applies AdaIN to the (H/8)*(W/8) pooled values.
//
// Numerics follow the reference's dtypes operation by operation (renderManager.py:882-935, math_utils.py:27-80):
//   mask  = half(1 - alpha)                                   fp16   (:883)
//   mixed = float(half(noise * half(1 - mask))) + bg * mask   fp32   (:929; the fp16 product rounds before the sum)
//   pooled = mean of 64 CONSECUTIVE pixels of a row: `noise.view(-1, 8, 8, 4).mean(dim=(1, 2))` on an NHWC tensor groups
//            256 consecutive floats, not an 8x8 block (:933) — reproduced as is
//   AdaIN(content = pooled fp32, style = raw noise attachment fp16, 'NHWC'): style variance / mean round to fp16,
//            `var + eps` in fp16, sqrt in fp32, std back to fp16 (math_utils.py:39-51); content statistics in fp32.
#include "srx_common.cuh"

#include <cuda_fp16.h>

struct IngestStats {          // workspace header: 16 doubles
    double style_sum[4], style_sq[4], content_sum[4], content_sq[4];
};

// Where the attachments come from.  Linear device buffers in the layout of Texture.tensor() (texture.py:166-254), or the
// mapped cudaArrays of the GL textures themselves, read through surface objects — the zero-copy path: no Memcpy2D into
// a staging tensor, no device-wide sync, no flip pass; the ingest kernel is the only reader of the texture memory.
// CUDA arrays hold 1, 2 or 4 channels: three-channel attachments (position, canny: RGB32F, renderManager.py:268,352) are
// four-channel arrays whose last channel is ignored.
struct SrcLinear {
    const __half *color, *normal_depth, *noise;
    const int4 *ids;
    const float *pos;
    const void *canny;
    int canny_f16, W;
    __device__ __forceinline__ bool has_color() const { return color != nullptr; }
    __device__ __forceinline__ bool has_ids() const { return ids != nullptr; }
    __device__ __forceinline__ bool has_pos() const { return pos != nullptr; }
    __device__ __forceinline__ bool has_nd() const { return normal_depth != nullptr; }
    __device__ __forceinline__ bool has_noise() const { return noise != nullptr; }
    __device__ __forceinline__ bool has_canny() const { return canny != nullptr; }
    __device__ __forceinline__ uint2 h4(const __half *b, int x, int y) const { return *reinterpret_cast<const uint2 *>(b + ((long long)y * W + x) * 4); }
    __device__ __forceinline__ uint2 ld_color(int x, int y) const { return h4(color, x, y); }
    __device__ __forceinline__ uint2 ld_nd(int x, int y) const { return h4(normal_depth, x, y); }
    __device__ __forceinline__ uint2 ld_noise(int x, int y) const { return h4(noise, x, y); }
    __device__ __forceinline__ int4 ld_ids(int x, int y) const { return ids[(long long)y * W + x]; }
    __device__ __forceinline__ void ld_pos(int x, int y, float *o) const { const float *q = pos + ((long long)y * W + x) * 3; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
    __device__ __forceinline__ void ld_canny_f(int x, int y, float *o) const { const float *q = reinterpret_cast<const float *>(canny) + ((long long)y * W + x) * 3; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
    __device__ __forceinline__ void ld_canny_h(int x, int y, __half *o) const { const __half *q = reinterpret_cast<const __half *>(canny) + ((long long)y * W + x) * 3; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
};
struct SrcArrays {
    cudaSurfaceObject_t color, normal_depth, noise, ids, pos, canny;   // 0 = absent
    int canny_f16;
    __device__ __forceinline__ bool has_color() const { return color != 0; }
    __device__ __forceinline__ bool has_ids() const { return ids != 0; }
    __device__ __forceinline__ bool has_pos() const { return pos != 0; }
    __device__ __forceinline__ bool has_nd() const { return normal_depth != 0; }
    __device__ __forceinline__ bool has_noise() const { return noise != 0; }
    __device__ __forceinline__ bool has_canny() const { return canny != 0; }
    __device__ __forceinline__ uint2 ld_color(int x, int y) const { uint2 v; surf2Dread(&v, color, x * 8, y); return v; }
    __device__ __forceinline__ uint2 ld_nd(int x, int y) const { uint2 v; surf2Dread(&v, normal_depth, x * 8, y); return v; }
    __device__ __forceinline__ uint2 ld_noise(int x, int y) const { uint2 v; surf2Dread(&v, noise, x * 8, y); return v; }
    __device__ __forceinline__ int4 ld_ids(int x, int y) const { int4 v; surf2Dread(&v, ids, x * 16, y); return v; }
    __device__ __forceinline__ void ld_pos(int x, int y, float *o) const { float4 v; surf2Dread(&v, pos, x * 16, y); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
    __device__ __forceinline__ void ld_canny_f(int x, int y, float *o) const { float4 v; surf2Dread(&v, canny, x * 16, y); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
    __device__ __forceinline__ void ld_canny_h(int x, int y, __half *o) const { uint2 v; surf2Dread(&v, canny, x * 8, y); const __half *q = reinterpret_cast<const __half *>(&v); o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
};

struct IngestPtrs {
    const float *bg;
    int canny_f16;            // the canny attachment (and canny_maps) is fp16 instead of f32
    __half *color_maps, *masks, *normal_maps, *depth_maps;
    int4 *id_maps;
    float *pos_maps;
    void *canny_maps;
    float *pooled;            // [G][4] fp32 scratch
    IngestStats *stats;
    int with_noise;
};

template <typename T> __device__ __forceinline__ void copy3(T *dst, const T *src) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }

// One warp per group of 64 consecutive output pixels, two pixels per lane.
template <typename Src>
__global__ void __launch_bounds__(256) k_ingest_frame(const Src src, IngestPtrs p, int H, int W, int flip, long long groups) {
    __shared__ double sh[8][16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * 8 + wid;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    if (g < groups) {
        float mix[4] = {0.f, 0.f, 0.f, 0.f};
        double ssum[4] = {0, 0, 0, 0}, ssq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const long long o = g * 64 + lane * 2 + k;          // output pixel (top-left origin)
            const int y = (int)(o / W), x = (int)(o % W);
            const int ys = flip ? H - 1 - y : y;                // source row (GL origin when flip)
            __half c[4];
            *reinterpret_cast<uint2 *>(c) = src.ld_color(x, ys);
            const __half mask = __float2half_rn(1.0f - __half2float(c[3]));
            if (p.color_maps) { p.color_maps[o * 3] = c[0]; p.color_maps[o * 3 + 1] = c[1]; p.color_maps[o * 3 + 2] = c[2]; }
            if (p.masks) p.masks[o] = mask;
            if (p.id_maps) p.id_maps[o] = src.ld_ids(x, ys);
            if (p.pos_maps) { float q[3]; src.ld_pos(x, ys, q); copy3(p.pos_maps + o * 3, q); }
            if (p.canny_maps) {
                if (p.canny_f16) { __half q[3]; src.ld_canny_h(x, ys, q); copy3(reinterpret_cast<__half *>(p.canny_maps) + o * 3, q); }
                else { float q[3]; src.ld_canny_f(x, ys, q); copy3(reinterpret_cast<float *>(p.canny_maps) + o * 3, q); }
            }
            if (src.has_nd() && (p.normal_maps || p.depth_maps)) {
                __half nd[4];
                *reinterpret_cast<uint2 *>(nd) = src.ld_nd(x, ys);
                if (p.normal_maps) { p.normal_maps[o * 3] = nd[0]; p.normal_maps[o * 3 + 1] = nd[1]; p.normal_maps[o * 3 + 2] = nd[2]; }
                if (p.depth_maps) { p.depth_maps[o * 3] = nd[3]; p.depth_maps[o * 3 + 1] = nd[3]; p.depth_maps[o * 3 + 2] = nd[3]; }
            }
            if (p.with_noise) {
                __half n[4];
                *reinterpret_cast<uint2 *>(n) = src.ld_noise(x, ys);
                const float4 bg = *reinterpret_cast<const float4 *>(p.bg + o * 4);   // GlobalBGNoise is already top-left
                const float bgv[4] = {bg.x, bg.y, bg.z, bg.w};
                const float mf = __half2float(mask);
                const float om = __half2float(__float2half_rn(1.0f - mf));
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const float nf = __half2float(n[ch]);
                    const float a = __half2float(__float2half_rn(__fmul_rn(nf, om)));
                    mix[ch] += __fadd_rn(a, __fmul_rn(bgv[ch], mf));
                    ssum[ch] += (double)nf;
                    ssq[ch] += (double)nf * (double)nf;
                }
            }
        }
        if (p.with_noise) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mix[ch] += __shfl_xor_sync(0xffffffffu, mix[ch], o);
                    ssum[ch] += __shfl_xor_sync(0xffffffffu, ssum[ch], o);
                    ssq[ch] += __shfl_xor_sync(0xffffffffu, ssq[ch], o);
                }
                const float pooled = mix[ch] * (1.0f / 64.0f);
                if (lane == 0) p.pooled[g * 4 + ch] = pooled;
                acc[ch] = ssum[ch];
                acc[4 + ch] = ssq[ch];
                acc[8 + ch] = (double)pooled;
                acc[12 + ch] = (double)pooled * (double)pooled;
            }
        }
    }
    if (!p.with_noise) return;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sh[wid][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        atomicAdd(reinterpret_cast<double *>(p.stats) + threadIdx.x, t);
    }
}

// noise_maps[slot, c, p] = (pooled[p,c] - mean_c) / std_c * style_std_c + style_mean_c (math_utils.py:78-80, one rounding per op)
__global__ void __launch_bounds__(256) k_ingest_adain(const float *__restrict__ pooled, const IngestStats *__restrict__ st,
                                                       float *__restrict__ out, long long groups, double n_style) {
    __shared__ float mc[4], sc[4], ms[4], ss[4];
    if (threadIdx.x < 4) {
        const int c = threadIdx.x;
        const double n = (double)groups;
        const double cm = st->content_sum[c] / n;
        const double cv = (st->content_sq[c] - st->content_sum[c] * cm) / (n - 1.0);          // unbiased (torch.var default)
        mc[c] = (float)cm;
        sc[c] = sqrtf(__fadd_rn((float)cv, 1e-5f));
        const double sm = st->style_sum[c] / n_style;
        const double sv = (st->style_sq[c] - st->style_sum[c] * sm) / (n_style - 1.0);
        const __half var_h = __float2half_rn((float)sv);                                        // fp16 var (math_utils.py:39)
        const __half var_eps = __float2half_rn(__fadd_rn(__half2float(var_h), 1e-5f));          // + eps in fp16
        ss[c] = __half2float(__float2half_rn(sqrtf(__half2float(var_eps))));                    // sqrt in fp32, std back to fp16 (:42-51)
        ms[c] = __half2float(__float2half_rn((float)sm));
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups * 4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 3);
        const long long g = i >> 2;
        const float v = __fdiv_rn(__fsub_rn(pooled[i], mc[c]), sc[c]);
        out[(long long)c * groups + g] = __fadd_rn(__fmul_rn(v, ss[c]), ms[c]);
    }
}

extern "C" int64_t srx_ingest_workspace_bytes(int height, int width) {
    if (height <= 0 || width <= 0) return -1;
    return 256 + (int64_t)height * width / 64 * 4 * (int64_t)sizeof(float);
}

static int make_surf(void *cuda_array, int want_bytes, const char *what, cudaSurfaceObject_t *out) {
    *out = 0;
    if (!cuda_array) return SRX_OK;
    cudaChannelFormatDesc desc;
    cudaExtent ext;
    unsigned int flags = 0;
    SRX_CUDA_CHECK(cudaArrayGetInfo(&desc, &ext, &flags, reinterpret_cast<cudaArray_t>(cuda_array)));
    const int bytes = (desc.x + desc.y + desc.z + desc.w) / 8;
    SRX_REQUIRE(bytes == want_bytes, SRX_ERR_INVALID, "%s array has %d bytes per texel, expected %d (three-channel GL formats map to "
                "four-channel CUDA arrays)", what, bytes, want_bytes);
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = reinterpret_cast<cudaArray_t>(cuda_array);
    SRX_CUDA_CHECK(cudaCreateSurfaceObject(out, &rd));
    return SRX_OK;
}

struct SurfSet {   // surface objects live until the launches that use them have been enqueued (they hold their own reference)
    cudaSurfaceObject_t s[6] = {0, 0, 0, 0, 0, 0};
    ~SurfSet() { for (int i = 0; i < 6; ++i) if (s[i]) cudaDestroySurfaceObject(s[i]); }
};

static int ingest_common(const srx_ingest_args *a, const srx_gbuffer_arrays *arr, void *stream) {
    SRX_REQUIRE(a, SRX_ERR_INVALID, "null argument");
    const int H = a->height, W = a->width;
    SRX_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, SRX_ERR_INVALID, "frame size must be a positive multiple of 8 (renderManager.py:933)");
    const bool from_arrays = arr != nullptr;
    const void *s_color = from_arrays ? arr->color : a->src.color, *s_ids = from_arrays ? arr->ids : a->src.ids,
               *s_pos = from_arrays ? arr->pos : (const void *)a->src.pos, *s_nd = from_arrays ? arr->normal_depth : a->src.normal_depth,
               *s_noise = from_arrays ? arr->noise : a->src.noise, *s_canny = from_arrays ? arr->canny : a->src.canny;
    const int canny_dtype = from_arrays ? arr->canny_dtype : a->src.canny_dtype;
    SRX_REQUIRE(s_color, SRX_ERR_INVALID, "the colour attachment is required (its alpha is the mask, renderManager.py:883)");
    SRX_REQUIRE(a->frame_slot >= 0, SRX_ERR_INVALID, "negative frame slot");
    SRX_REQUIRE(!s_canny || canny_dtype == SRX_F32 || canny_dtype == SRX_F16, SRX_ERR_INVALID, "canny dtype must be f32 or f16");
    SRX_REQUIRE(!a->id_maps || s_ids, SRX_ERR_INVALID, "id_maps requested without an id attachment");
    SRX_REQUIRE(!a->pos_maps || s_pos, SRX_ERR_INVALID, "pos_maps requested without a position attachment");
    SRX_REQUIRE(!a->canny_maps || s_canny, SRX_ERR_INVALID, "canny_maps requested without a canny attachment");
    SRX_REQUIRE((!a->normal_maps && !a->depth_maps) || s_nd, SRX_ERR_INVALID, "normal/depth maps requested without the attachment");
    const bool with_noise = a->noise_maps != nullptr;
    SRX_REQUIRE(!with_noise || (s_noise && a->bg_noise), SRX_ERR_INVALID, "noise_maps needs the noise attachment and the background noise");
    SRX_REQUIRE(!with_noise || (a->workspace && a->workspace_bytes >= srx_ingest_workspace_bytes(H, W)), SRX_ERR_INVALID, "workspace too small");
    SRX_REQUIRE(!with_noise || (reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long npx = (long long)H * W, groups = npx / 64, slot = a->frame_slot;
    IngestPtrs p;
    p.canny_f16 = canny_dtype == SRX_F16;
    p.bg = a->bg_noise;
    p.with_noise = with_noise ? 1 : 0;
    p.color_maps = a->color_maps ? reinterpret_cast<__half *>(a->color_maps) + slot * npx * 3 : nullptr;
    p.masks = a->masks ? reinterpret_cast<__half *>(a->masks) + slot * npx : nullptr;
    p.normal_maps = a->normal_maps ? reinterpret_cast<__half *>(a->normal_maps) + slot * npx * 3 : nullptr;
    p.depth_maps = a->depth_maps ? reinterpret_cast<__half *>(a->depth_maps) + slot * npx * 3 : nullptr;
    p.id_maps = a->id_maps ? reinterpret_cast<int4 *>(a->id_maps) + slot * npx : nullptr;
    p.pos_maps = a->pos_maps ? a->pos_maps + slot * npx * 3 : nullptr;
    p.canny_maps = a->canny_maps ? reinterpret_cast<unsigned char *>(a->canny_maps) + slot * npx * 3 * (p.canny_f16 ? 2 : 4) : nullptr;
    p.stats = reinterpret_cast<IngestStats *>(a->workspace);
    p.pooled = with_noise ? reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(a->workspace) + 256) : nullptr;
    if (with_noise) SRX_CUDA_CHECK(cudaMemsetAsync(a->workspace, 0, 256, st));
    const unsigned int grid = (unsigned int)((groups + 7) / 8);
    if (from_arrays) {
        SurfSet ss;
        int rc;
        if ((rc = make_surf(const_cast<void *>(s_color), 8, "colour", &ss.s[0]))) return rc;
        if ((rc = make_surf(const_cast<void *>(s_nd), 8, "normal+depth", &ss.s[1]))) return rc;
        if ((rc = make_surf(with_noise ? const_cast<void *>(s_noise) : nullptr, 8, "noise", &ss.s[2]))) return rc;
        if ((rc = make_surf(const_cast<void *>(s_ids), 16, "id", &ss.s[3]))) return rc;
        if ((rc = make_surf(const_cast<void *>(s_pos), 16, "position", &ss.s[4]))) return rc;
        if ((rc = make_surf(const_cast<void *>(s_canny), p.canny_f16 ? 8 : 16, "canny", &ss.s[5]))) return rc;
        SrcArrays src{ss.s[0], ss.s[1], ss.s[2], ss.s[3], ss.s[4], ss.s[5], p.canny_f16};
        k_ingest_frame<SrcArrays><<<grid, 256, 0, st>>>(src, p, H, W, a->flip_rows ? 1 : 0, groups);
        SRX_CUDA_CHECK(cudaGetLastError());
    } else {
        SrcLinear src{reinterpret_cast<const __half *>(s_color), reinterpret_cast<const __half *>(s_nd),
                      with_noise ? reinterpret_cast<const __half *>(s_noise) : nullptr, reinterpret_cast<const int4 *>(s_ids),
                      reinterpret_cast<const float *>(s_pos), s_canny, p.canny_f16, W};
        k_ingest_frame<SrcLinear><<<grid, 256, 0, st>>>(src, p, H, W, a->flip_rows ? 1 : 0, groups);
    }
    if (with_noise) {
        const long long work = groups * 4;
        const int g2 = (int)((work + 255) / 256 < 296 ? (work + 255) / 256 : 296);
        k_ingest_adain<<<g2, 256, 0, st>>>(p.pooled, p.stats, a->noise_maps + slot * groups * 4, groups, (double)npx);
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

extern "C" int srx_frame_ingest(const srx_ingest_args *a, void *stream) { return ingest_common(a, nullptr, stream); }

// Zero-copy form: the attachments are the mapped cudaArrays of the GL textures (srx_gl_map); `args->src` is ignored.
extern "C" int srx_frame_ingest_arrays(const srx_ingest_args *a, const srx_gbuffer_arrays *arrays, void *stream) {
    SRX_REQUIRE(arrays, SRX_ERR_INVALID, "null argument");
    return ingest_common(a, arrays, stream);
}

// -----------------------------------------------------------------------------------------------------------------
// "closer pixel wins" (renderManager.py:121-133): where the current draw's reversed depth (normal_depth alpha; larger =
// closer, default_Gbuffer.frag.glsl:123) exceeds the stored one, every attachment of the pixel replaces the stored one.
// canny is stored as fp16 (renderManager.py:357); the attachment is read as fp16 by the reference (data_type HALF, :353) or f32.
// -----------------------------------------------------------------------------------------------------------------
template <typename Src>
__global__ void __launch_bounds__(256) k_gbuffer_merge_closer(const Src cur, srx_gbuffer_temp tmp, int canny_f16, int H, int W, int flip) {
    const long long npx = (long long)H * W;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < npx; o += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(o / W), x = (int)(o % W);
        const int ys = flip ? H - 1 - y : y;
        __half nd[4];
        *reinterpret_cast<uint2 *>(nd) = cur.ld_nd(x, ys);
        __half *depth = reinterpret_cast<__half *>(tmp.depth);
        if (!(__half2float(nd[3]) > __half2float(depth[o]))) continue;
        depth[o] = nd[3];
        if (tmp.normal) {
            __half *n = reinterpret_cast<__half *>(tmp.normal) + o * 3;
            n[0] = nd[0]; n[1] = nd[1]; n[2] = nd[2];
        }
        if (tmp.color && cur.has_color()) *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(tmp.color) + o * 4) = cur.ld_color(x, ys);
        if (tmp.ids && cur.has_ids()) reinterpret_cast<int4 *>(tmp.ids)[o] = cur.ld_ids(x, ys);
        if (tmp.pos && cur.has_pos()) { float q[3]; cur.ld_pos(x, ys, q); copy3(tmp.pos + o * 3, q); }
        if (tmp.noise && cur.has_noise()) *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(tmp.noise) + o * 4) = cur.ld_noise(x, ys);
        if (tmp.canny && cur.has_canny()) {
            __half *c = reinterpret_cast<__half *>(tmp.canny) + o * 3;
            if (canny_f16) {
                __half q[3]; cur.ld_canny_h(x, ys, q); copy3(c, q);
            } else {
                float f[3]; cur.ld_canny_f(x, ys, f);
                c[0] = __float2half_rn(f[0]); c[1] = __float2half_rn(f[1]); c[2] = __float2half_rn(f[2]);
            }
        }
    }
}

extern "C" int srx_gbuffer_merge_closer(const srx_gbuffer *cur, int height, int width, int flip_rows, const srx_gbuffer_temp *temp,
                                        void *stream) {
    SRX_REQUIRE(cur && temp, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(height > 0 && width > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(cur->normal_depth && temp->depth, SRX_ERR_INVALID, "the normal+depth attachment and the depth buffer are required");
    SRX_REQUIRE(!cur->canny || cur->canny_dtype == SRX_F32 || cur->canny_dtype == SRX_F16, SRX_ERR_INVALID, "canny dtype must be f32 or f16");
    const long long npx = (long long)height * width;
    const long long nb = (npx + 255) / 256, cap = (long long)srx_sm_count_cached() * 8;
    SrcLinear src{reinterpret_cast<const __half *>(cur->color), reinterpret_cast<const __half *>(cur->normal_depth),
                  reinterpret_cast<const __half *>(cur->noise), reinterpret_cast<const int4 *>(cur->ids), cur->pos, cur->canny,
                  cur->canny_dtype == SRX_F16, width};
    k_gbuffer_merge_closer<SrcLinear><<<(int)(nb < cap ? nb : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, *temp, cur->canny_dtype == SRX_F16, height, width, flip_rows ? 1 : 0);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

extern "C" int srx_gbuffer_merge_closer_arrays(const srx_gbuffer_arrays *cur, int height, int width, int flip_rows,
                                               const srx_gbuffer_temp *temp, void *stream) {
    SRX_REQUIRE(cur && temp, SRX_ERR_INVALID, "null argument");
    SRX_REQUIRE(height > 0 && width > 0, SRX_ERR_INVALID, "non-positive dimension");
    SRX_REQUIRE(cur->normal_depth && temp->depth, SRX_ERR_INVALID, "the normal+depth attachment and the depth buffer are required");
    SRX_REQUIRE(!cur->canny || cur->canny_dtype == SRX_F32 || cur->canny_dtype == SRX_F16, SRX_ERR_INVALID, "canny dtype must be f32 or f16");
    const bool cf16 = cur->canny_dtype == SRX_F16;
    SurfSet ss;
    int rc;
    if ((rc = make_surf(cur->color, 8, "colour", &ss.s[0]))) return rc;
    if ((rc = make_surf(cur->normal_depth, 8, "normal+depth", &ss.s[1]))) return rc;
    if ((rc = make_surf(cur->noise, 8, "noise", &ss.s[2]))) return rc;
    if ((rc = make_surf(cur->ids, 16, "id", &ss.s[3]))) return rc;
    if ((rc = make_surf(cur->pos, 16, "position", &ss.s[4]))) return rc;
    if ((rc = make_surf(cur->canny, cf16 ? 8 : 16, "canny", &ss.s[5]))) return rc;
    SrcArrays src{ss.s[0], ss.s[1], ss.s[2], ss.s[3], ss.s[4], ss.s[5], cf16};
    const long long npx = (long long)height * width;
    const long long nb = (npx + 255) / 256, cap = (long long)srx_sm_count_cached() * 8;
    k_gbuffer_merge_closer<SrcArrays><<<(int)(nb < cap ? nb : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, *temp, cf16, height, width, flip_rows ? 1 : 0);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
