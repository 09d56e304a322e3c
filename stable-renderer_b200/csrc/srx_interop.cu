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
