// L2 reduction micro-probe (profiling aid, not part of libsrx): how fast can B200 retire scattered {sum.xyzw, count} updates
//   mode 0: red.global.add.v4.f32 + red.global.add.f32 into [K][4] + [K]            (what phase A does today)
//   mode 1: red.global.add.v4.f32 + red.global.add.f32 into one 32-byte record [K][8] (same sector for both)
//   mode 2: cp.reduce.async.bulk.global.shared::cta.add.f32 of a 32-byte record per key (TMA engine)
//   mode 3: red.global.add.v4.f32 only
//   mode 4: scalar red only (the count)      mode 5: mode 0 with every other lane idle (same updates, twice the instructions)
//   mode 6: v4 sum + v4 count vector (three zeros)   mode 7: v4 sum, the count only for every other update (half merged)
//   mode 8: TMA bulk reduce of RUNS of 5 adjacent keys: one 80-byte sum chunk + one 32-byte count chunk per 5 updates
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(512) k_probe(float *acc, long long K, long long n_updates, int mode) {
    extern __shared__ __align__(128) float sm[];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    float *rec = sm + threadIdx.x * 8;
    for (long long i = tid; i < n_updates; i += nth) {
        // texel-like locality: consecutive threads hit nearby (not identical) keys, different per iteration
        const long long k = (long long)((hash((unsigned)(i >> 5)) + (unsigned)(i & 31) * 3u) % (unsigned long long)K);
        const float v = 1.0f;
        if (mode == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k * 4), "f"(v) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + K * 4 + k), "f"(v) : "memory");
        } else if (mode == 1) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k * 8), "f"(v) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + k * 8 + 4), "f"(v) : "memory");
        } else if (mode == 3) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k * 4), "f"(v) : "memory");
        } else if (mode == 4) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + K * 4 + k), "f"(v) : "memory");
        } else if (mode == 5) {
            if ((threadIdx.x & 1) == 0) {
                for (int r = 0; r < 2; ++r) {
                    const long long k2 = (k + r * 977) % K;
                    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k2 * 4), "f"(v) : "memory");
                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + K * 4 + k2), "f"(v) : "memory");
                }
            }
        } else if (mode == 6) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k * 4), "f"(v) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%2,%2};" ::"l"(acc + K * 4 + (k & ~3ll)), "f"(v), "f"(0.f) : "memory");
        } else if (mode == 7) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(acc + k * 4), "f"(v) : "memory");
            if (i & 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + K * 4 + k), "f"(2.f * v) : "memory");
        } else if (mode == 8) {
            if (i % 5 == 0) {      // this thread stands for a run of 5 adjacent keys
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                float *run = sm + threadIdx.x * 28;
#pragma unroll
                for (int j = 0; j < 28; ++j) run[j] = (j < 20 || (j >= 20 && j < 25)) ? v : 0.f;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const unsigned s = (unsigned)__cvta_generic_to_shared(run);
                const long long kb = (k & ~3ll) % (K - 8);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 80;" ::"l"(acc + kb * 4), "r"(s) : "memory");
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 32;" ::"l"(acc + K * 4 + kb), "r"(s + 80) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // wait until this thread's previous bulk reduction has read its record
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) rec[j] = j < 5 ? v : 0.f;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned s = (unsigned)__cvta_generic_to_shared(rec);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 32;" ::"l"(acc + k * 8), "r"(s) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (mode == 2 || mode == 8) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const long long K = 262144, N = 6400000;
    float *acc;
    cudaMalloc(&acc, K * 8 * sizeof(float));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 112);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[9] = {"v4 + scalar RED, split arrays", "v4 + scalar RED, one 32 B record", "TMA bulk reduce of a 32 B record", "v4 RED only",
                            "scalar RED only", "v4 + scalar, half the lanes idle", "v4 sum + v4 count vector", "v4 sum + count for every 2nd",
                            "TMA bulk reduce, runs of 5 keys"};
    for (int mode = 0; mode < 9; ++mode) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaMemset(acc, 0, K * 8 * sizeof(float));
            cudaEventRecord(e0);
            k_probe<<<148 * 2, 512, 512 * 112>>>(acc, K, N, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        float h[8];
        cudaMemcpy(h, acc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-36s %8.1f us for %lld updates = %6.1f G updates/s  (%s, acc[0..4] = %.0f %.0f %.0f %.0f %.0f)\n", names[mode],
               best * 1e3, N, N / (best * 1e-3) / 1e9, cudaGetErrorString(err), h[0], h[1], h[2], h[3], h[4]);
    }
    return 0;
}
