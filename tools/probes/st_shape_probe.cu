// st_shape_probe.cu -- pins the register <-> (TMEM lane, column) mapping of tcgen05.st.16x256b.x1 as used by the
// coalesced row producer of mlp_tc.cu: every thread stores tags with the 16x256b shape (two instructions per warp:
// lane offset 0 and 16), then the block reads TMEM back row-per-thread (32x32b) and dumps [lane][column].
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_ptr;
    for (int blk = 0; blk < 2; ++blk) {
        const uint32_t addr = tbase + ((uint32_t)(warp * 32 + blk * 16) << 16);
        uint32_t r[4];
        for (int j = 0; j < 4; ++j) r[j] = (blk << 24) | (warp << 16) | (lane << 8) | j;
        asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    const uint32_t laddr = tbase + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(laddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * 8 + j] = v[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(32));
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 128 * 8 * 4);
    cudaMemset(d, 0xff, 128 * 8 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[128 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 128; ++row)
        for (int c = 0; c < 8; ++c) {
            const uint32_t t = h[row * 8 + c];
            const int blk = t >> 24, warp = (t >> 16) & 0xff, lane = (t >> 8) & 0xff, reg = t & 0xff;
            // expected: row = warp*32 + blk*16 + lane/4 + 8*(reg/2), column = 2*(lane%4) + reg%2
            const int erow = warp * 32 + blk * 16 + lane / 4 + 8 * (reg / 2), ecol = 2 * (lane % 4) + reg % 2;
            if (erow != row || ecol != c) {
                if (bad < 16) printf("row %3d col %d: tag blk=%d warp=%d lane=%2d reg=%d (expected row %d col %d)\n", row, c, blk, warp, lane, reg, erow, ecol);
                ++bad;
            }
        }
    printf("16x256b.x1 mapping: %s (%d mismatches)\n", bad ? "DIFFERENT" : "as expected: reg{0,1} = row lane/4, cols 2*(lane%%4)+{0,1}; reg{2,3} = row lane/4+8", bad);
    return 0;
}
