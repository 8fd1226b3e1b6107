// lat_probe.cu -- dependent-chain latency (cycles per operation, one warp, clock64) of the warp primitives the FPS
// exchange is built from: REDUX (__reduce_max_sync), SHFL, VOTE (ballot), LDS, and a 5-step shuffle arg-max.
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void probe(long long* out, unsigned seed) {
    __shared__ unsigned sm[64];
    const int lane = threadIdx.x;
    sm[lane] = lane * 7 + seed;
    sm[lane + 32] = lane;
    __syncwarp();
    unsigned v = seed + lane;
    long long t0, t1;
    // REDUX chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) v = __reduce_max_sync(0xffffffffu, v + lane) ^ (unsigned)i;
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / 64;
    // SHFL chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + (unsigned)i;
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / 64;
    // ballot chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) v = __ballot_sync(0xffffffffu, (v + lane) & 1) + (unsigned)i;
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / 64;
    // LDS chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) v = sm[(v + i) & 63];
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / 64;
    // shuffle arg-max (value, index), 5 steps
    unsigned val = v * 2654435761u + lane, idx = lane;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned ov = __shfl_xor_sync(0xffffffffu, val, o), oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
        }
        val = val * 1664525u + lane + idx;
    }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / 16;
    // two dependent REDUX (max then min of the index among the maxima), as in fps.cu
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const unsigned m = __reduce_max_sync(0xffffffffu, val);
        const unsigned k = __reduce_min_sync(0xffffffffu, val == m ? idx : 0xffffffffu);
        val = val * 1664525u + lane + k;
    }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / 16;
    if (lane == 0) out[6] = v + val + idx;
}

int main() {
    long long* d;
    cudaMalloc(&d, 8 * sizeof(long long));
    probe<<<1, 32>>>(d, 3);
    probe<<<1, 32>>>(d, 5);
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("cycles per dependent op: REDUX %lld, SHFL %lld, VOTE %lld, LDS %lld; shuffle arg-max (5 steps) %lld; REDUX max + REDUX min %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5]);
    return 0;
}
