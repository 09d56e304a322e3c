// streamprobe — how fast can one B200 stream the id buffers of the overlap step into the SMs?
// Variants: plain 256-bit loads, and the shared-memory ring fed by cp.async.bulk with different stage shapes, depths,
// item orders and L2 policies.  Prints GB/s per variant (CUDA events, best of N after warm-up).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o streamprobe streamprobe.cu && ./streamprobe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// ---- plain loads: every thread streams 32-byte vectors, U in flight -------------------------------------------------
template <int U, int HINT>
__global__ void __launch_bounds__(512) k_ldg(const char *base, long long bytes, unsigned *sink) {
    const long long nvec = bytes / 32;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride * U) {
        unsigned r[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = v + u * stride;
            if (i < nvec) {
                if (HINT)
                    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7]) : "l"(base + i * 32));
                else
                    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7]) : "l"(base + i * 32));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[u][j] = 0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc ^= r[u][j];
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// ---- ring fed by bulk copies ----------------------------------------------------------------------------------------
struct RingParams {
    const char *base;
    int rows;            // image rows in total (frames * H)
    long long pitch;     // bytes per image row
    int rpi;             // rows per item
    int rb;              // bytes per row copy (item = rpi x rb bytes)
    int chunks;          // row chunks per image row = pitch / rb
    int order;           // 0 chunk-major (the step kernel's deal), 1 row-major, 2 contiguous per CTA (row-major blocks)
    int stages, hint, cons;
    int skew;            // bytes added to the shared-memory row pitch (the step kernel skews rows by 16 B against bank conflicts)
    int extra_n, extra_bytes;   // extra tiny copies per stage (the step kernel fetches 4 latent rows per stage)
    const char *extra;
    unsigned *sink;
};

__global__ void __launch_bounds__(544, 1) k_ring(const RingParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stage_bytes = P.rpi * (P.rb + P.skew) + ((P.extra_n * P.extra_bytes + 127) & ~127);
    const int data_bytes = P.rpi * P.rb + P.extra_n * P.extra_bytes;
    const uint32_t bar0 = sbase + P.stages * stage_bytes;
    if (tid == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(bar0 + s * 8, 1); mbar_init(bar0 + (P.stages + s) * 8, P.cons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long items = (long long)(P.rows / P.rpi) * P.chunks;
    const int G = gridDim.x;
    const int rowblocks = P.rows / P.rpi;
    if (warp == P.cons) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        int stage = 0; unsigned ph = 0;
        const long long per = (items + G - 1) / G;
        for (long long k = 0;; ++k) {
            long long item = P.order == 2 ? (long long)blockIdx.x * per + k : (long long)blockIdx.x + k * G;
            if (P.order == 2 && k >= per) item = items;
            mbar_wait(bar0 + (P.stages + stage) * 8, ph ^ 1u);
            const uint32_t full = bar0 + stage * 8;
            const uint32_t sb = sbase + stage * stage_bytes;
            if (item >= items) {
                if (lane == 0) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb), "r"(-1) : "memory"); mbar_arrive(full); }
                break;
            }
            long long rblk, chunk;
            if (P.order == 0) { chunk = (int)item / rowblocks; rblk = (int)item - chunk * rowblocks; }
            else { rblk = (int)item / P.chunks; chunk = (int)item - rblk * P.chunks; if (P.order == 3) chunk = (chunk + rblk) % P.chunks; }
            if (lane == 0) mbar_expect_tx(full, (uint32_t)data_bytes);
            __syncwarp();
            for (int r = lane; r < P.rpi; r += 32) {
                const char *src = P.base + (rblk * P.rpi + r) * P.pitch + chunk * P.rb;
                if (P.hint) bulk_g2s_hint(sb + r * (P.rb + P.skew), src, (uint32_t)P.rb, full, pol);
                else bulk_g2s(sb + r * (P.rb + P.skew), src, (uint32_t)P.rb, full);
            }
            if (lane >= 8 && lane < 8 + P.extra_n)
                bulk_g2s(sb + P.rpi * (P.rb + P.skew) + (lane - 8) * P.extra_bytes, P.extra + ((item * 4 + lane - 8) % 4096) * 512, (uint32_t)P.extra_bytes, full);
            if (++stage == P.stages) { stage = 0; ph ^= 1u; }
        }
    } else if (warp < P.cons) {
        int stage = 0; unsigned ph = 0;
        unsigned acc = 0;
        const int per_warp = P.rpi * P.rb / P.cons;
        while (true) {
            mbar_wait(bar0 + stage * 8, ph);
            const uint32_t sb = sbase + stage * stage_bytes;
            int first;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(first) : "r"(sb));
            // (an end marker of -1 in the first word; real data never starts with -1 here: the buffer is filled with small ints)
            if (first == -1) break;
            for (int o = lane * 16; o < per_warp; o += 512) {
                int4 v;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sb + warp * per_warp + o));
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + (P.stages + stage) * 8);
            if (++stage == P.stages) { stage = 0; ph ^= 1u; }
        }
        if (acc == 0x12345u) P.sink[0] = acc;
    }
}

static float time_best(void (*launch)(void *), void *arg, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int i = 0; i < reps + 2; ++i) {
        CK(cudaEventRecord(a));
        launch(arg);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 2) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

struct LdgArg { const char *p; long long bytes; unsigned *sink; int grid, u, hint; };
static void launch_ldg(void *a_) {
    LdgArg *a = (LdgArg *)a_;
    if (a->u == 2) { if (a->hint) k_ldg<2, 1><<<a->grid, 512>>>(a->p, a->bytes, a->sink); else k_ldg<2, 0><<<a->grid, 512>>>(a->p, a->bytes, a->sink); }
    else if (a->u == 4) { if (a->hint) k_ldg<4, 1><<<a->grid, 512>>>(a->p, a->bytes, a->sink); else k_ldg<4, 0><<<a->grid, 512>>>(a->p, a->bytes, a->sink); }
    else { if (a->hint) k_ldg<8, 1><<<a->grid, 512>>>(a->p, a->bytes, a->sink); else k_ldg<8, 0><<<a->grid, 512>>>(a->p, a->bytes, a->sink); }
}
struct RingArg { RingParams P; int grid; };
static void launch_ring(void *a_) {
    RingArg *a = (RingArg *)a_;
    const int smem = a->P.stages * (a->P.rpi * (a->P.rb + a->P.skew) + ((a->P.extra_n * a->P.extra_bytes + 127) & ~127)) + 2 * a->P.stages * 8 + 64;
    CK(cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_ring<<<a->grid, (a->P.cons + 1) * 32, smem>>>(a->P);
}

int main(int argc, char **argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 96, H = argc > 2 ? atoi(argv[2]) : 1024;
    const long long bytes = (long long)frames * H * H * 16;
    char *buf; unsigned *sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(buf, 1, bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("streamprobe: %d frames of %dx%d x 16 B = %.1f MB, %d SMs\n", frames, H, H, bytes / 1e6, sms);
    for (int hint = 1; hint < 2; ++hint)
        for (int u : {8})
            for (int mult : {2, 4}) {
                LdgArg a{buf, bytes, sink, sms * mult, u, hint};
                const float ms = time_best(launch_ldg, &a, 5);
                printf("ldg256  U=%d ctas/sm=%d evict_first=%d : %8.1f us  %7.1f GB/s\n", u, mult, hint, ms * 1e3, bytes / ms / 1e6);
            }
    struct V { int rpi, rb, stages, order, hint, skew, en, eb; };
    const long long pitch = (long long)H * 16;
    char *extra; CK(cudaMalloc(&extra, 4096 * 512)); CK(cudaMemset(extra, 1, 4096 * 512));
    std::vector<V> vs = {
        {8, 4096, 6, 0, 1, 0, 0, 0},      // chunk-major (the step kernel's deal)
        {8, 4096, 6, 1, 1, 0, 0, 0},      // row-major
        {8, 4096, 6, 3, 1, 0, 0, 0},      // row-major, chunk rotated by the row block
        {8, 4096, 6, 2, 1, 0, 0, 0},      // contiguous block of items per CTA
        {8, 4096, 6, 0, 1, 16, 0, 0},     // + 16 B row skew
        {8, 4096, 6, 1, 1, 16, 0, 0},
        {8, 4096, 6, 0, 1, 0, 4, 64},     // + 4 tiny copies per stage (bf16 latents: 32 cells x 2 B)
        {8, 4096, 6, 1, 1, 0, 4, 64},
        {8, 4096, 6, 1, 1, 0, 4, 128},
        {8, 4096, 6, 1, 1, 0, 1, 512},
        {8, 4096, 6, 0, 1, 16, 4, 64},    // both = the step kernel's ring
        {8, 4096, 6, 1, 1, 16, 4, 64},
        {8, 4096, 6, 3, 1, 16, 4, 64},
        {8, 4096, 6, 3, 1, 16, 4, 128},
        {8, 4096, 5, 3, 1, 16, 4, 128},
        {8, 4096, 4, 3, 1, 16, 4, 128},
        {8, 4096, 3, 3, 1, 16, 4, 128},
        {16, 4096, 3, 3, 1, 16, 0, 0},
        {8, 8192, 3, 3, 1, 16, 0, 0},
    };
    for (const V &v : vs) {
        if ((long long)v.rb > pitch || pitch % v.rb) continue;
        RingArg a;
        a.P = RingParams{buf, frames * H, pitch, v.rpi, v.rb, (int)(pitch / v.rb), v.order, v.stages, v.hint, 16, v.skew, v.en, v.eb, extra, sink};
        a.grid = sms;
        const float ms = time_best(launch_ring, &a, 5);
        printf("ring  rows/item=%2d row_bytes=%5d stages=%2d order=%d evict_first=%d skew=%2d extra=%dx%3dB : %8.1f us  %7.1f GB/s\n", v.rpi, v.rb, v.stages,
               v.order, v.hint, v.skew, v.en, v.eb, ms * 1e3, bytes / ms / 1e6);
    }
    return 0;
}
