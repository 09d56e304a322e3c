// L2 gather micro-probe (profiling aid, not part of libsrx): how fast can B200 gather small records at random from an L2-resident
// table — the access pattern of a key-major (atomic-free) cached-plan reduction: per (key, cell) pair one 4-byte packed entry
// (streamed) and one 8- or 16-byte latent record gathered from a cell-major copy of the latents.
//   ./gatherprobe [pairs = 26.5M] [cells = 1.57M]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <typename REC, int UN>
__global__ void __launch_bounds__(256) k_gather(const unsigned *__restrict__ pairs, const REC *__restrict__ table, long long n, float *out) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (long long i = tid; i < n; i += nth * UN) {
        unsigned e[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) e[u] = i + u * nth < n ? __ldg(pairs + i + u * nth) : 0u;
        REC r[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) r[u] = __ldg(table + (e[u] >> 6));
#pragma unroll
        for (int u = 0; u < UN; ++u) acc += (float)((e[u] & 63u) + 1u) * __uint_as_float(((const unsigned *)&r[u])[0] & 0x3fffffffu);
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void k_fill(unsigned *pairs, long long n, unsigned cells) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        pairs[i] = ((hash((unsigned)i) % cells) << 6) | (unsigned)(i & 3);
}

template <typename REC, int UN>
static void run(const char *name, const unsigned *pairs, const void *table, long long n, float *out, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        k_gather<REC, UN><<<148 * ctas_per_sm, 256>>>(pairs, (const REC *)table, n, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-34s UN=%d ctas/sm=%d : %8.1f us  %6.1f G gathers/s  (%s)\n", name, UN, ctas_per_sm, best * 1e3, n / (best * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 26500000ll;
    const unsigned cells = argc > 2 ? (unsigned)atoll(argv[2]) : 1572864u;
    unsigned *pairs;
    void *table;
    float *out;
    cudaMalloc(&pairs, n * 4);
    cudaMalloc(&table, (size_t)cells * 16);
    cudaMalloc(&out, 4);
    cudaMemset(table, 0, (size_t)cells * 16);
    k_fill<<<148 * 8, 256>>>(pairs, n, cells);
    cudaDeviceSynchronize();
    printf("gatherprobe: %lld pairs (4 B each, streamed) gathering from %u cells\n", n, cells);
    for (int c = 4; c <= 8; c += 4) {
        run<uint2, 4>("8-byte records (4 x bf16)", pairs, table, n, out, c);
        run<uint2, 8>("8-byte records (4 x bf16)", pairs, table, n, out, c);
        run<uint4, 4>("16-byte records (4 x f32)", pairs, table, n, out, c);
        run<uint4, 8>("16-byte records (4 x f32)", pairs, table, n, out, c);
    }
    return 0;
}
