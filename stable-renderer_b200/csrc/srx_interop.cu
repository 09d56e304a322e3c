// srx_interop.cu — texture <-> tensor interop without staging copies or device-wide syncs.
//
// Replaces Texture._init_tensor / tensor / set_data (source/engine/static/texture/texture.py:166-254, 326-408), which
// map the GL texture with pycuda, Memcpy2D it into a linear buffer, call torch.cuda.synchronize() and then flip the
// rows with a second pass, and the cuda-python wrappers of source/common_utils/cuda_utils.py:101-190.
// Here the mapped cudaArray is bound to a surface object and one kernel on the caller's stream moves texels straight
// between the array and the tensor with the bottom-left -> top-left flip (texture.py:236,253) fused in.
#include "srx_common.cuh"

// cuda_gl_interop.h needs GL headers that are not part of this image; the two entry points are declared by hand
// (GLuint / GLenum are unsigned int).
extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource **resource, unsigned int image,
                                                   unsigned int target, unsigned int flags);

struct srx_gl_resource {
    cudaGraphicsResource *res = nullptr;
    bool mapped = false;
};

extern "C" int srx_gl_register_image(srx_gl_resource **out, unsigned int gl_texture, unsigned int gl_target, unsigned int flags) {
    SRX_REQUIRE(out, SRX_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaGraphicsResource *res = nullptr;
    cudaError_t e = cudaGraphicsGLRegisterImage(&res, gl_texture, gl_target, flags);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the (non-sticky) error so that later launch checks do not see it
        return srx_set_error(SRX_ERR_CUDA, "cudaGraphicsGLRegisterImage(tex=%u, target=0x%x) failed: %s (is a GL context current on this thread?)",
                             gl_texture, gl_target, cudaGetErrorString(e));
    }
    srx_gl_resource *r = new srx_gl_resource();
    r->res = res;
    *out = r;
    return SRX_OK;
}

extern "C" int srx_gl_map(srx_gl_resource *r, void **cuda_array_out, void *stream) {
    SRX_REQUIRE(r && cuda_array_out, SRX_ERR_INVALID, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!r->mapped) {
        SRX_CUDA_CHECK(cudaGraphicsMapResources(1, &r->res, st));
        r->mapped = true;
    }
    cudaArray_t arr = nullptr;
    SRX_CUDA_CHECK(cudaGraphicsSubResourceGetMappedArray(&arr, r->res, 0, 0));
    *cuda_array_out = arr;
    return SRX_OK;
}

extern "C" int srx_gl_unmap(srx_gl_resource *r, void *stream) {
    SRX_REQUIRE(r, SRX_ERR_INVALID, "null argument");
    if (r->mapped) {
        SRX_CUDA_CHECK(cudaGraphicsUnmapResources(1, &r->res, reinterpret_cast<cudaStream_t>(stream)));
        r->mapped = false;
    }
    return SRX_OK;
}

extern "C" int srx_gl_unregister(srx_gl_resource *r) {
    if (!r) return SRX_OK;
    if (r->mapped) cudaGraphicsUnmapResources(1, &r->res, 0);
    cudaError_t e = cudaGraphicsUnregisterResource(r->res);
    delete r;
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "cudaGraphicsUnregisterResource failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

// one thread per texel; texel type = the widest POD matching the texel size so that tensor accesses are coalesced
template <typename T>
__global__ void __launch_bounds__(256) k_array_to_tensor(cudaSurfaceObject_t surf, T *__restrict__ dst, int width, int height, int flip) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    T v;
    surf2Dread(&v, surf, x * (int)sizeof(T), y);
    const int oy = flip ? height - 1 - y : y;
    dst[(long long)oy * width + x] = v;
}

template <typename T>
__global__ void __launch_bounds__(256) k_tensor_to_array(cudaSurfaceObject_t surf, const T *__restrict__ src, int width, int height,
                                                          int flip, int x_off, int y_off) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const int sy = flip ? height - 1 - y : y;
    const T v = src[(long long)sy * width + x];
    surf2Dwrite(v, surf, (x + x_off) * (int)sizeof(T), y + y_off);
}

static int make_surface(void *cuda_array, int texel_bytes, cudaSurfaceObject_t *surf) {
    SRX_REQUIRE(cuda_array, SRX_ERR_INVALID, "null cudaArray");
    cudaChannelFormatDesc desc;
    cudaExtent ext;
    unsigned int flags = 0;
    SRX_CUDA_CHECK(cudaArrayGetInfo(&desc, &ext, &flags, reinterpret_cast<cudaArray_t>(cuda_array)));
    const int bytes = (desc.x + desc.y + desc.z + desc.w) / 8;
    SRX_REQUIRE(bytes == texel_bytes, SRX_ERR_INVALID, "texel size mismatch: array has %d bytes per texel, tensor %d", bytes, texel_bytes);
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = reinterpret_cast<cudaArray_t>(cuda_array);
    SRX_CUDA_CHECK(cudaCreateSurfaceObject(surf, &rd));
    return SRX_OK;
}

extern "C" int srx_array_to_tensor(void *cuda_array, void *dst_dev, int width, int height, int texel_bytes, int flip, void *stream) {
    SRX_REQUIRE(dst_dev && width > 0 && height > 0, SRX_ERR_INVALID, "bad argument");
    cudaSurfaceObject_t surf = 0;
    int rc = make_surface(cuda_array, texel_bytes, &surf);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((width + 31) / 32, (height + 7) / 8), block(256);
    switch (texel_bytes) {
        case 1: k_array_to_tensor<unsigned char><<<grid, block, 0, st>>>(surf, (unsigned char *)dst_dev, width, height, flip); break;
        case 2: k_array_to_tensor<unsigned short><<<grid, block, 0, st>>>(surf, (unsigned short *)dst_dev, width, height, flip); break;
        case 4: k_array_to_tensor<unsigned int><<<grid, block, 0, st>>>(surf, (unsigned int *)dst_dev, width, height, flip); break;
        case 8: k_array_to_tensor<uint2><<<grid, block, 0, st>>>(surf, (uint2 *)dst_dev, width, height, flip); break;
        case 16: k_array_to_tensor<uint4><<<grid, block, 0, st>>>(surf, (uint4 *)dst_dev, width, height, flip); break;
        default: cudaDestroySurfaceObject(surf); return srx_set_error(SRX_ERR_UNSUPPORTED, "texel size %d (CUDA arrays hold 1, 2 or 4 channels)", texel_bytes);
    }
    cudaError_t e = cudaGetLastError();
    cudaDestroySurfaceObject(surf);  // the launch holds its own reference to the array
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "array_to_tensor launch failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

extern "C" int srx_tensor_to_array(void *cuda_array, const void *src_dev, int width, int height, int texel_bytes, int flip,
                                   int x_offset, int y_offset, void *stream) {
    SRX_REQUIRE(src_dev && width > 0 && height > 0 && x_offset >= 0 && y_offset >= 0, SRX_ERR_INVALID, "bad argument");
    cudaSurfaceObject_t surf = 0;
    int rc = make_surface(cuda_array, texel_bytes, &surf);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((width + 31) / 32, (height + 7) / 8), block(256);
    switch (texel_bytes) {
        case 1: k_tensor_to_array<unsigned char><<<grid, block, 0, st>>>(surf, (const unsigned char *)src_dev, width, height, flip, x_offset, y_offset); break;
        case 2: k_tensor_to_array<unsigned short><<<grid, block, 0, st>>>(surf, (const unsigned short *)src_dev, width, height, flip, x_offset, y_offset); break;
        case 4: k_tensor_to_array<unsigned int><<<grid, block, 0, st>>>(surf, (const unsigned int *)src_dev, width, height, flip, x_offset, y_offset); break;
        case 8: k_tensor_to_array<uint2><<<grid, block, 0, st>>>(surf, (const uint2 *)src_dev, width, height, flip, x_offset, y_offset); break;
        case 16: k_tensor_to_array<uint4><<<grid, block, 0, st>>>(surf, (const uint4 *)src_dev, width, height, flip, x_offset, y_offset); break;
        default: cudaDestroySurfaceObject(surf); return srx_set_error(SRX_ERR_UNSUPPORTED, "texel size %d", texel_bytes);
    }
    cudaError_t e = cudaGetLastError();
    cudaDestroySurfaceObject(surf);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "tensor_to_array launch failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

extern "C" int srx_array_alloc(void **cuda_array_out, int width, int height, int channels, int bits_per_channel, int kind) {
    SRX_REQUIRE(cuda_array_out && width > 0 && height > 0, SRX_ERR_INVALID, "bad argument");
    SRX_REQUIRE(channels == 1 || channels == 2 || channels == 4, SRX_ERR_INVALID, "CUDA arrays hold 1, 2 or 4 channels");
    cudaChannelFormatKind k = kind == 0 ? cudaChannelFormatKindSigned : (kind == 1 ? cudaChannelFormatKindUnsigned : cudaChannelFormatKindFloat);
    cudaChannelFormatDesc desc = cudaCreateChannelDesc(bits_per_channel, channels >= 2 ? bits_per_channel : 0,
                                                       channels >= 4 ? bits_per_channel : 0, channels >= 4 ? bits_per_channel : 0, k);
    cudaArray_t arr = nullptr;
    SRX_CUDA_CHECK(cudaMallocArray(&arr, &desc, (size_t)width, (size_t)height, cudaArraySurfaceLoadStore));
    *cuda_array_out = arr;
    return SRX_OK;
}

extern "C" int srx_array_free(void *cuda_array) {
    if (cuda_array) SRX_CUDA_CHECK(cudaFreeArray(reinterpret_cast<cudaArray_t>(cuda_array)));
    return SRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Channel-aware copies for the Texture wrapper (texture.py:221-254, 326-408).  A GL texture with three channels is a
// four-channel CUDA array; Texture.tensor() still hands out [H,W,3], and Texture.set_data() pads RGB data with alpha = 1
// (texture.py:379-380), broadcasts single-channel data (:377-378) and truncates wider data (:383-384).  E = element type.
// ---------------------------------------------------------------------------------------------------------------
template <typename E, int AC>
__global__ void __launch_bounds__(256) k_array_to_tensor_ch(cudaSurfaceObject_t surf, E *__restrict__ dst, int width, int height,
                                                             int dst_ch, int flip) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    E v[AC];
    if (AC * sizeof(E) == 16) { uint4 t; surf2Dread(&t, surf, x * 16, y); memcpy(v, &t, 16); }
    else if (AC * sizeof(E) == 8) { uint2 t; surf2Dread(&t, surf, x * 8, y); memcpy(v, &t, 8); }
    else if (AC * sizeof(E) == 4) { unsigned t; surf2Dread(&t, surf, x * 4, y); memcpy(v, &t, 4); }
    else if (AC * sizeof(E) == 2) { unsigned short t; surf2Dread(&t, surf, x * 2, y); memcpy(v, &t, 2); }
    else { unsigned char t; surf2Dread(&t, surf, x, y); memcpy(v, &t, 1); }
    const int oy = flip ? height - 1 - y : y;
    E *o = dst + ((long long)oy * width + x) * dst_ch;
    for (int c = 0; c < dst_ch && c < AC; ++c) o[c] = v[c];
}

template <typename E, int AC>
__global__ void __launch_bounds__(256) k_tensor_to_array_ch(cudaSurfaceObject_t surf, const E *__restrict__ src, int width, int height,
                                                             int src_ch, int flip, int x_off, int y_off, E one) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const int sy = flip ? height - 1 - y : y;
    const E *in = src + ((long long)sy * width + x) * src_ch;
    E v[AC];
#pragma unroll
    for (int c = 0; c < AC; ++c) {
        if (c < src_ch) v[c] = in[c];
        else if (src_ch == 1) v[c] = in[0];          // single channel: repeated (texture.py:377-378)
        else if (c == 3) v[c] = one;                  // RGB -> RGBA: alpha = 1 (texture.py:379-380)
        else memset(&v[c], 0, sizeof(E));
    }
    if (AC * sizeof(E) == 16) { uint4 t; memcpy(&t, v, 16); surf2Dwrite(t, surf, (x + x_off) * 16, y + y_off); }
    else if (AC * sizeof(E) == 8) { uint2 t; memcpy(&t, v, 8); surf2Dwrite(t, surf, (x + x_off) * 8, y + y_off); }
    else if (AC * sizeof(E) == 4) { unsigned t; memcpy(&t, v, 4); surf2Dwrite(t, surf, (x + x_off) * 4, y + y_off); }
    else if (AC * sizeof(E) == 2) { unsigned short t; memcpy(&t, v, 2); surf2Dwrite(t, surf, (x + x_off) * 2, y + y_off); }
    else { unsigned char t; memcpy(&t, v, 1); surf2Dwrite(t, surf, x + x_off, y + y_off); }
}

static int array_channels(void *cuda_array, int elem_bytes, int *channels, int *aw, int *ah, cudaSurfaceObject_t *surf) {
    SRX_REQUIRE(cuda_array, SRX_ERR_INVALID, "null cudaArray");
    cudaChannelFormatDesc desc;
    cudaExtent ext;
    unsigned int flags = 0;
    SRX_CUDA_CHECK(cudaArrayGetInfo(&desc, &ext, &flags, reinterpret_cast<cudaArray_t>(cuda_array)));
    SRX_REQUIRE(desc.x == elem_bytes * 8, SRX_ERR_INVALID, "element size mismatch: array channels have %d bits, tensor elements %d", desc.x,
                elem_bytes * 8);
    *channels = (desc.x ? 1 : 0) + (desc.y ? 1 : 0) + (desc.z ? 1 : 0) + (desc.w ? 1 : 0);
    *aw = (int)ext.width;
    *ah = (int)ext.height;
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = reinterpret_cast<cudaArray_t>(cuda_array);
    SRX_CUDA_CHECK(cudaCreateSurfaceObject(surf, &rd));
    return SRX_OK;
}

template <typename E>
static int a2t_ch(cudaSurfaceObject_t surf, int ac, void *dst, int w, int h, int dst_ch, int flip, cudaStream_t st) {
    dim3 grid((w + 31) / 32, (h + 7) / 8), block(256);
    if (ac == 1) k_array_to_tensor_ch<E, 1><<<grid, block, 0, st>>>(surf, (E *)dst, w, h, dst_ch, flip);
    else if (ac == 2) k_array_to_tensor_ch<E, 2><<<grid, block, 0, st>>>(surf, (E *)dst, w, h, dst_ch, flip);
    else k_array_to_tensor_ch<E, 4><<<grid, block, 0, st>>>(surf, (E *)dst, w, h, dst_ch, flip);
    return SRX_OK;
}
template <typename E>
static int t2a_ch(cudaSurfaceObject_t surf, int ac, const void *src, int w, int h, int src_ch, int flip, int xo, int yo, E one, cudaStream_t st) {
    dim3 grid((w + 31) / 32, (h + 7) / 8), block(256);
    if (ac == 1) k_tensor_to_array_ch<E, 1><<<grid, block, 0, st>>>(surf, (const E *)src, w, h, src_ch, flip, xo, yo, one);
    else if (ac == 2) k_tensor_to_array_ch<E, 2><<<grid, block, 0, st>>>(surf, (const E *)src, w, h, src_ch, flip, xo, yo, one);
    else k_tensor_to_array_ch<E, 4><<<grid, block, 0, st>>>(surf, (const E *)src, w, h, src_ch, flip, xo, yo, one);
    return SRX_OK;
}

// tensor [height,width,dst_channels] <- array; dst_channels <= the array's channel count (extra array channels are dropped)
extern "C" int srx_array_to_tensor_ch(void *cuda_array, void *dst_dev, int width, int height, int elem_bytes, int dst_channels,
                                      int flip, void *stream) {
    SRX_REQUIRE(dst_dev && width > 0 && height > 0 && dst_channels >= 1 && dst_channels <= 4, SRX_ERR_INVALID, "bad argument");
    int ac = 0, aw = 0, ah = 0;
    cudaSurfaceObject_t surf = 0;
    int rc = array_channels(cuda_array, elem_bytes, &ac, &aw, &ah, &surf);
    if (rc) return rc;
    if (dst_channels > ac || width > aw || height > ah) {
        cudaDestroySurfaceObject(surf);
        return srx_set_error(SRX_ERR_INVALID, "tensor [%d,%d,%d] does not fit the %dx%d array of %d channels", height, width, dst_channels, aw, ah, ac);
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (elem_bytes == 1) a2t_ch<unsigned char>(surf, ac, dst_dev, width, height, dst_channels, flip, st);
    else if (elem_bytes == 2) a2t_ch<unsigned short>(surf, ac, dst_dev, width, height, dst_channels, flip, st);
    else if (elem_bytes == 4) a2t_ch<unsigned int>(surf, ac, dst_dev, width, height, dst_channels, flip, st);
    else { cudaDestroySurfaceObject(surf); return srx_set_error(SRX_ERR_UNSUPPORTED, "element size %d", elem_bytes); }
    cudaError_t e = cudaGetLastError();
    cudaDestroySurfaceObject(surf);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "array_to_tensor_ch launch failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

// array region [y_offset .. +height, x_offset .. +width] <- tensor [height,width,src_channels]; `one_bits` = the value 1 in the
// element type (alpha of padded RGB data)
extern "C" int srx_tensor_to_array_ch(void *cuda_array, const void *src_dev, int width, int height, int elem_bytes, int src_channels,
                                      int flip, int x_offset, int y_offset, unsigned int one_bits, void *stream) {
    SRX_REQUIRE(src_dev && width > 0 && height > 0 && x_offset >= 0 && y_offset >= 0 && src_channels >= 1, SRX_ERR_INVALID, "bad argument");
    int ac = 0, aw = 0, ah = 0;
    cudaSurfaceObject_t surf = 0;
    int rc = array_channels(cuda_array, elem_bytes, &ac, &aw, &ah, &surf);
    if (rc) return rc;
    if (x_offset + width > aw || y_offset + height > ah) {
        cudaDestroySurfaceObject(surf);
        return srx_set_error(SRX_ERR_INVALID, "region %dx%d at (%d,%d) exceeds the %dx%d array", width, height, x_offset, y_offset, aw, ah);
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (elem_bytes == 1) t2a_ch<unsigned char>(surf, ac, src_dev, width, height, src_channels, flip, x_offset, y_offset, (unsigned char)one_bits, st);
    else if (elem_bytes == 2) t2a_ch<unsigned short>(surf, ac, src_dev, width, height, src_channels, flip, x_offset, y_offset, (unsigned short)one_bits, st);
    else if (elem_bytes == 4) t2a_ch<unsigned int>(surf, ac, src_dev, width, height, src_channels, flip, x_offset, y_offset, one_bits, st);
    else { cudaDestroySurfaceObject(surf); return srx_set_error(SRX_ERR_UNSUPPORTED, "element size %d", elem_bytes); }
    cudaError_t e = cudaGetLastError();
    cudaDestroySurfaceObject(surf);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "tensor_to_array_ch launch failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Atlas layer -> GL_TEXTURE_2D_ARRAY layer (CorrespondMap.load, corrmap.py:443-489): the reference uploads
// get_map(i, order='whc') through the host, i.e. (square atlas) texel (x, y) of layer i = _values[i, x * width + y], the
// transpose of the [height,width] map; the shader samples it with (uv.y, uv.x), default_Gbuffer.frag.glsl:186-200.  Here one
// kernel writes the layer's width x height cudaArray from the fp16 atlas, the transpose done through a shared-memory tile.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_atlas_to_array(cudaSurfaceObject_t surf, const T *__restrict__ values, int Ht, int Wt, int mode) {
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (mode == 0) {                                   // plain 'hwc' layer
        const int x = blockIdx.x * 32 + tx;
        for (int r = ty; r < 32; r += 8) {
            const int y = blockIdx.y * 32 + r;
            if (x < Wt && y < Ht) surf2Dwrite(values[(long long)y * Wt + x], surf, x * (int)sizeof(T), y);
        }
        return;
    }
    if (mode == 1) {                                   // square atlas: texel (x, y) <- values[x * Wt + y], through a tile
        for (int r = ty; r < 32; r += 8) {
            const int row = blockIdx.x * 32 + r, col = blockIdx.y * 32 + tx;     // atlas (row, col)
            if (row < Ht && col < Wt) tile[r][tx] = values[(long long)row * Wt + col];
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 32 + r;         // texel x = atlas row, y = atlas col
            if (x < Ht && y < Wt) surf2Dwrite(tile[tx][r], surf, x * (int)sizeof(T), y);
        }
        return;
    }
    // non-square atlas: the reference hands the bytes of the [W,H,C] array to a width x height upload, i.e. texel (x, y) is
    // element y * Wt + x of the flattened [W][H] array = atlas (h = flat % Ht, w = flat / Ht)
    const int x = blockIdx.x * 32 + tx;
    for (int r = ty; r < 32; r += 8) {
        const int y = blockIdx.y * 32 + r;
        if (x < Wt && y < Ht) {
            const long long flat = (long long)y * Wt + x;
            const long long w = flat / Ht, h = flat - w * Ht;
            surf2Dwrite(values[h * Wt + w], surf, x * (int)sizeof(T), y);
        }
    }
}

extern "C" int srx_atlas_to_array(void *cuda_array, const void *values_dev, int height, int width, int channels, int transpose, void *stream) {
    SRX_REQUIRE(values_dev && height > 0 && width > 0, SRX_ERR_INVALID, "bad argument");
    SRX_REQUIRE(channels == 1 || channels == 2 || channels == 4, SRX_ERR_UNSUPPORTED,
                "CUDA arrays hold 1, 2 or 4 channels (RGB16F cannot be registered, renderManager.py:268)");
    cudaSurfaceObject_t surf = 0;
    int rc = make_surface(cuda_array, channels * 2, &surf);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int mode = !transpose ? 0 : (height == width ? 1 : 2);
    dim3 grid((width + 31) / 32, (height + 31) / 32), block(256);
    if (channels == 1) k_atlas_to_array<unsigned short><<<grid, block, 0, st>>>(surf, (const unsigned short *)values_dev, height, width, mode);
    else if (channels == 2) k_atlas_to_array<unsigned int><<<grid, block, 0, st>>>(surf, (const unsigned int *)values_dev, height, width, mode);
    else k_atlas_to_array<uint2><<<grid, block, 0, st>>>(surf, (const uint2 *)values_dev, height, width, mode);
    cudaError_t e = cudaGetLastError();
    cudaDestroySurfaceObject(surf);
    if (e != cudaSuccess) return srx_set_error(SRX_ERR_CUDA, "atlas_to_array launch failed: %s", cudaGetErrorString(e));
    return SRX_OK;
}

// mapped array of one layer of a registered GL_TEXTURE_2D_ARRAY (the resource must be mapped: srx_gl_map)
extern "C" int srx_gl_mapped_layer(srx_gl_resource *r, int layer, void **cuda_array_out) {
    SRX_REQUIRE(r && cuda_array_out && layer >= 0, SRX_ERR_INVALID, "bad argument");
    SRX_REQUIRE(r->mapped, SRX_ERR_INVALID, "map the resource first (srx_gl_map)");
    cudaArray_t arr = nullptr;
    SRX_CUDA_CHECK(cudaGraphicsSubResourceGetMappedArray(&arr, r->res, (unsigned int)layer, 0));
    *cuda_array_out = arr;
    return SRX_OK;
}
