// tc_probe.cu -- validates the tcgen05 conventions mlp_tc.cu relies on, one MMA group at a time:
//   D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, packed bf16, written with tcgen05.st) * B[N x K]^T (smem, K-major,
//   no swizzle, core matrices of 8 rows x 16 bytes).  Integer-valued inputs make the expected result exact.
// Variants: lbo_is_k (which descriptor field is the K-direction core-matrix stride), a_even_low (which half of a
// 32-bit TMEM cell holds the even k).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int N, int K, int lbo_is_k, int a_even_low) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- B into smem: element (n,k) at (n/8)*SBO + (k/8)*LBO + (n%8)*16 + (k%8)*2
    const int LBO = 128, SBO = (K / 8) * 128;
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        const uint32_t off = (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16_rn(B[n * K + k]);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_ptr;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    // ---- A row `tid` into TMEM columns [128, 128 + K/2): packed pairs
    {
        uint32_t r[16];
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 2 * (c0 + j);
                float e = 0.f, o = 0.f;
                if (k < K) { e = A[tid * K + k]; o = A[tid * K + k + 1]; }
                const uint32_t eb = __bfloat16_as_ushort(__float2bfloat16_rn(e)), ob = __bfloat16_as_ushort(__float2bfloat16_rn(o));
                r[j] = a_even_low ? (eb | (ob << 16)) : (ob | (eb << 16));
            }
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                    lane_addr + 128 + c0),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // instruction descriptor: c=F32(1)<<4, a=BF16(1)<<7, b=BF16(1)<<10, K-major both, N>>3 <<17, M>>4 <<24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t b_addr = smem_u32(smem) + ks * 2 * LBO;
            const uint32_t lbo = lbo_is_k ? LBO : SBO, sbo = lbo_is_k ? SBO : LBO;
            uint64_t desc = 0;
            desc |= (uint64_t)((b_addr >> 4) & 0x3FFF);
            desc |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
            desc |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
            desc |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
            const uint32_t a_addr = tbase + 128 + ks * 8;
            const uint32_t acc = ks > 0;
            asm volatile(
                "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tbase),
                "r"(a_addr), "l"(desc), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(lane_addr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

int main() {
    const int Ns[] = {32, 128, 256}, Ks[] = {32, 64, 128};
    for (int N : Ns)
        for (int K : Ks) {
            std::vector<float> A(M * K), B(N * K), Dref(M * N), D(M * N);
            srand(N * 131 + K);
            for (auto& v : A) v = (float)(rand() % 7 - 3);
            for (auto& v : B) v = (float)(rand() % 5 - 2);
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    float s = 0;
                    for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
                    Dref[m * N + n] = s;
                }
            float *dA, *dB, *dD;
            cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
            cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
            for (int lbo_is_k = 1; lbo_is_k >= 0; --lbo_is_k)
                for (int a_even_low = 1; a_even_low >= 0; --a_even_low) {
                    cudaMemset(dD, 0xFF, D.size() * 4);
                    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, N * K * 2 + 1024);
                    probe<<<1, 128, N * K * 2 + 1024>>>(dA, dB, dD, N, K, lbo_is_k, a_even_low);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                    int bad = 0; float maxerr = 0;
                    for (size_t i = 0; i < D.size(); ++i) { float d = fabsf(D[i] - Dref[i]); if (!(d <= 1e-3f)) ++bad; if (d > maxerr) maxerr = d; }
                    printf("N=%3d K=%3d lbo_is_k=%d a_even_low=%d: %s bad=%d/%zu maxerr=%g  D[0..3]=%g %g %g %g ref=%g %g %g %g\n", N, K, lbo_is_k,
                           a_even_low, e == cudaSuccess ? "ok " : cudaGetErrorString(e), bad, D.size(), maxerr, D[0], D[1], D[2], D[3],
                           Dref[0], Dref[1], Dref[2], Dref[3]);
                    if (e != cudaSuccess) return 1;
                }
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    return 0;
}
