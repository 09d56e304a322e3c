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

struct IngestPtrs {
    const __half *color, *normal_depth, *noise;
    const int4 *ids;
    const float *pos, *bg;
    const void *canny;
    int canny_f16;            // the canny attachment (and canny_maps) is fp16 instead of f32
    __half *color_maps, *masks, *normal_maps, *depth_maps;
    int4 *id_maps;
    float *pos_maps;
    void *canny_maps;
    float *pooled;            // [G][4] fp32 scratch
    IngestStats *stats;
};

template <typename T> __device__ __forceinline__ void copy3(T *dst, const T *src) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }

// One warp per group of 64 consecutive output pixels, two pixels per lane.
__global__ void __launch_bounds__(256) k_ingest_frame(IngestPtrs p, int H, int W, int flip, long long groups) {
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
            const long long s = (long long)(flip ? H - 1 - y : y) * W + x;   // source pixel (GL origin when flip)
            __half c[4];
            *reinterpret_cast<uint2 *>(c) = *reinterpret_cast<const uint2 *>(p.color + s * 4);
            const __half mask = __float2half_rn(1.0f - __half2float(c[3]));
            if (p.color_maps) { p.color_maps[o * 3] = c[0]; p.color_maps[o * 3 + 1] = c[1]; p.color_maps[o * 3 + 2] = c[2]; }
            if (p.masks) p.masks[o] = mask;
            if (p.id_maps) p.id_maps[o] = p.ids[s];
            if (p.pos_maps) copy3(p.pos_maps + o * 3, p.pos + s * 3);
            if (p.canny_maps) {
                if (p.canny_f16) copy3(reinterpret_cast<__half *>(p.canny_maps) + o * 3, reinterpret_cast<const __half *>(p.canny) + s * 3);
                else copy3(reinterpret_cast<float *>(p.canny_maps) + o * 3, reinterpret_cast<const float *>(p.canny) + s * 3);
            }
            if (p.normal_depth) {
                __half nd[4];
                *reinterpret_cast<uint2 *>(nd) = *reinterpret_cast<const uint2 *>(p.normal_depth + s * 4);
                if (p.normal_maps) { p.normal_maps[o * 3] = nd[0]; p.normal_maps[o * 3 + 1] = nd[1]; p.normal_maps[o * 3 + 2] = nd[2]; }
                if (p.depth_maps) { p.depth_maps[o * 3] = nd[3]; p.depth_maps[o * 3 + 1] = nd[3]; p.depth_maps[o * 3 + 2] = nd[3]; }
            }
            if (p.noise) {
                __half n[4];
                *reinterpret_cast<uint2 *>(n) = *reinterpret_cast<const uint2 *>(p.noise + s * 4);
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
        if (p.noise) {
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
    if (!p.noise) return;
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

extern "C" int srx_frame_ingest(const srx_ingest_args *a, void *stream) {
    SRX_REQUIRE(a, SRX_ERR_INVALID, "null argument");
    const int H = a->height, W = a->width;
    SRX_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, SRX_ERR_INVALID, "frame size must be a positive multiple of 8 (renderManager.py:933)");
    SRX_REQUIRE(a->src.color, SRX_ERR_INVALID, "the colour attachment is required (its alpha is the mask, renderManager.py:883)");
    SRX_REQUIRE(a->frame_slot >= 0, SRX_ERR_INVALID, "negative frame slot");
    SRX_REQUIRE(!a->src.canny || a->src.canny_dtype == SRX_F32 || a->src.canny_dtype == SRX_F16, SRX_ERR_INVALID, "canny dtype must be f32 or f16");
    SRX_REQUIRE(!a->id_maps || a->src.ids, SRX_ERR_INVALID, "id_maps requested without an id attachment");
    SRX_REQUIRE(!a->pos_maps || a->src.pos, SRX_ERR_INVALID, "pos_maps requested without a position attachment");
    SRX_REQUIRE(!a->canny_maps || a->src.canny, SRX_ERR_INVALID, "canny_maps requested without a canny attachment");
    SRX_REQUIRE((!a->normal_maps && !a->depth_maps) || a->src.normal_depth, SRX_ERR_INVALID, "normal/depth maps requested without the attachment");
    const bool with_noise = a->noise_maps != nullptr;
    SRX_REQUIRE(!with_noise || (a->src.noise && a->bg_noise), SRX_ERR_INVALID, "noise_maps needs the noise attachment and the background noise");
    SRX_REQUIRE(!with_noise || (a->workspace && a->workspace_bytes >= srx_ingest_workspace_bytes(H, W)), SRX_ERR_INVALID, "workspace too small");
    SRX_REQUIRE(!with_noise || (reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, SRX_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long npx = (long long)H * W, groups = npx / 64, slot = a->frame_slot;
    IngestPtrs p;
    p.color = reinterpret_cast<const __half *>(a->src.color);
    p.normal_depth = reinterpret_cast<const __half *>(a->src.normal_depth);
    p.noise = with_noise ? reinterpret_cast<const __half *>(a->src.noise) : nullptr;
    p.ids = reinterpret_cast<const int4 *>(a->src.ids);
    p.pos = a->src.pos;
    p.canny = a->src.canny;
    p.canny_f16 = a->src.canny_dtype == SRX_F16;
    p.bg = a->bg_noise;
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
    k_ingest_frame<<<(unsigned int)((groups + 7) / 8), 256, 0, st>>>(p, H, W, a->flip_rows ? 1 : 0, groups);
    if (with_noise) {
        const long long work = groups * 4;
        const int grid = (int)((work + 255) / 256 < 296 ? (work + 255) / 256 : 296);
        k_ingest_adain<<<grid, 256, 0, st>>>(p.pooled, p.stats, a->noise_maps + slot * groups * 4, groups, (double)npx);
    }
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}

// -----------------------------------------------------------------------------------------------------------------
// "closer pixel wins" (renderManager.py:121-133): where the current draw's reversed depth (normal_depth alpha; larger =
// closer, default_Gbuffer.frag.glsl:123) exceeds the stored one, every attachment of the pixel replaces the stored one.
// canny is stored as fp16 (renderManager.py:357); the attachment is read as fp16 by the reference (data_type HALF, :353) or f32.
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gbuffer_merge_closer(srx_gbuffer cur, srx_gbuffer_temp tmp, int H, int W, int flip) {
    const long long npx = (long long)H * W;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < npx; o += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(o / W), x = (int)(o % W);
        const long long s = (long long)(flip ? H - 1 - y : y) * W + x;
        __half nd[4];
        *reinterpret_cast<uint2 *>(nd) = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(cur.normal_depth) + s * 4);
        __half *depth = reinterpret_cast<__half *>(tmp.depth);
        if (!(__half2float(nd[3]) > __half2float(depth[o]))) continue;
        depth[o] = nd[3];
        if (tmp.normal) {
            __half *n = reinterpret_cast<__half *>(tmp.normal) + o * 3;
            n[0] = nd[0]; n[1] = nd[1]; n[2] = nd[2];
        }
        if (tmp.color && cur.color)
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(tmp.color) + o * 4) = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(cur.color) + s * 4);
        if (tmp.ids && cur.ids) reinterpret_cast<int4 *>(tmp.ids)[o] = reinterpret_cast<const int4 *>(cur.ids)[s];
        if (tmp.pos && cur.pos) copy3(tmp.pos + o * 3, cur.pos + s * 3);
        if (tmp.noise && cur.noise)
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(tmp.noise) + o * 4) = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(cur.noise) + s * 4);
        if (tmp.canny && cur.canny) {
            __half *c = reinterpret_cast<__half *>(tmp.canny) + o * 3;
            if (cur.canny_dtype == SRX_F16) {
                copy3(c, reinterpret_cast<const __half *>(cur.canny) + s * 3);
            } else {
                const float *f = reinterpret_cast<const float *>(cur.canny) + s * 3;
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
    k_gbuffer_merge_closer<<<(int)(nb < cap ? nb : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*cur, *temp, height, width,
                                                                                                          flip_rows ? 1 : 0);
    SRX_CUDA_CHECK(cudaGetLastError());
    return SRX_OK;
}
