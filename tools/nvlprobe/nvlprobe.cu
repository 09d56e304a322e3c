// NVLink peer-access micro-probe (profiling aid, not part of libsrx): times, on the local SM clock / globaltimer, how long
// one GPU needs to (a) load contiguously, (b) gather random 32-byte sectors, (c) store contiguously + fence.sys from / to
// a peer's memory.  Built by tools/nvlprobe/build.sh, driven by tools/nvl_probe.py.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// mode 0: contiguous 16 B loads; 1: random 32 B gathers (idx = hash); 2: contiguous 16 B stores + fence.sys;
// 3: contiguous 32 B stores + fence.sys; 4: contiguous 32 B loads
__global__ void k_probe(const char *peer, char *peer_w, long long bytes, int mode, float *sink, unsigned long long *tout) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    unsigned long long t0 = gtime();
    float acc = 0.f;
    if (mode == 0) {
        for (long long i = tid; i < bytes / 16; i += nth) {
            float4 v;
            asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(peer + i * 16));
            acc += v.x + v.w;
        }
    } else if (mode == 1 || mode == 4) {
        const long long n = bytes / 32;
        for (long long i = tid; i < n; i += nth) {
            long long j = mode == 1 ? (long long)((unsigned long long)(i * 2654435761ull + 12345) % (unsigned long long)n) : i;
            unsigned a, b, c, d, e, f, g, h;
            asm volatile("ld.relaxed.sys.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(peer + j * 32));
            acc += __uint_as_float(a) + __uint_as_float(h);
        }
    } else if (mode == 2) {
        for (long long i = tid; i < bytes / 16; i += nth)
            asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(peer_w + i * 16), "f"(1.0f) : "memory");
        __threadfence_system();
    } else if (mode == 3) {
        for (long long i = tid; i < bytes / 32; i += nth)
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(peer_w + i * 32), "r"(7u) : "memory");
        __threadfence_system();
    }
    unsigned long long t1 = gtime();
    if (acc == 123.456f) *sink = acc;
    // per-CTA (start, end); host takes min start / max end
    if (threadIdx.x == 0) { tout[2 * blockIdx.x] = t0; tout[2 * blockIdx.x + 1] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        // end = when the slowest thread of the CTA finished: approximate with a second read after the barrier
        tout[2 * blockIdx.x + 1] = gtime();
    }
}

extern "C" int nvl_probe(const void *peer, void *peer_w, long long bytes, int mode, int ctas, int threads, void *tout_dev, void *stream) {
    static float *sink = nullptr;
    if (!sink) cudaMalloc(&sink, 4);
    k_probe<<<ctas, threads, 0, (cudaStream_t)stream>>>((const char *)peer, (char *)peer_w, bytes, mode, sink, (unsigned long long *)tout_dev);
    return (int)cudaGetLastError();
}
