// linear.cu -- one shared-MLP layer in exact fp32: y = act(x * w^T + bias), BatchNorm pre-folded.
// Reference: the Conv2d/Conv1d(kernel 1) + BatchNorm + ReLU triples of model/pointnet_util.py:195-197,
// :253-255, :310-312, pointnet2.py:172-173, pointnet.py; nn.Linear heads; torch.bmm transforms.
//
// CUDA-core SGEMM (fp32 FMA accumulation, K ascending).  Both operands are K-contiguous in
// global memory (x rows are points, w rows are output channels), so tiles are fetched with 16-byte
// loads along K and transposed into shared memory; each thread then owns an 8x8 (or 8x4) register
// tile.  This is the fp32-exact path; the tensor-core path for the heavy levels lives in mlp_tc.cu.
#include "common.cuh"

namespace pn {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_kernel(const float* __restrict__ x, int64_t ldx, int64_t x_bstride, const float* __restrict__ w, int64_t w_bstride,
              const float* __restrict__ bias_base, int64_t bias_bstride, int relu, int64_t rows, int cin, int cout,
              float* __restrict__ y, int64_t ldy, int64_t y_bstride, int vec_ok) {
    constexpr int THREADS = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];

    const int bz = blockIdx.z;
    const float* __restrict__ xb = x + (int64_t)bz * x_bstride;
    const float* __restrict__ wb = w + (int64_t)bz * w_bstride;
    float* __restrict__ yb = y + (int64_t)bz * y_bstride;
    const float* __restrict__ bias = bias_base ? bias_base + (int64_t)bz * bias_bstride : nullptr;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);  // column group
    const int ty = tid / (BN / TN);  // row group

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    constexpr int KV = BK / 4;  // float4 per tile row
    for (int k0 = 0; k0 < cin; k0 += BK) {
        // ---- A tile: BM rows x BK
        for (int e = tid; e < BM * KV; e += THREADS) {
            const int r = e / KV, kq = (e % KV) * 4;
            const int64_t gr = row0 + r;
            const int gk = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < rows) {
                const float* src = xb + gr * ldx + gk;
                if ((vec_ok & 1) && gk + 3 < cin) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (gk < cin) v.x = src[0];
                    if (gk + 1 < cin) v.y = src[1];
                    if (gk + 2 < cin) v.z = src[2];
                    if (gk + 3 < cin) v.w = src[3];
                }
            }
            As[kq][r] = v.x;
            As[kq + 1][r] = v.y;
            As[kq + 2][r] = v.z;
            As[kq + 3][r] = v.w;
        }
        // ---- B tile: BN output channels x BK
        for (int e = tid; e < BN * KV; e += THREADS) {
            const int r = e / KV, kq = (e % KV) * 4;
            const int gc = col0 + r;
            const int gk = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gc < cout) {
                const float* src = wb + (int64_t)gc * cin + gk;
                if ((vec_ok & 2) && gk + 3 < cin) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (gk < cin) v.x = src[0];
                    if (gk + 1 < cin) v.y = src[1];
                    if (gk + 2 < cin) v.z = src[2];
                    if (gk + 3 < cin) v.w = src[3];
                }
            }
            Bs[kq][r] = v.x;
            Bs[kq + 1][r] = v.y;
            Bs[kq + 2][r] = v.z;
            Bs[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
                b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // ---- epilogue: bias + ReLU
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gr = row0 + ty * TM + i;
        if (gr >= rows) continue;
        float* __restrict__ dst = yb + gr * ldy;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gc = col0 + tx * TN + j;
            if (gc >= cout) continue;
            float v = acc[i][j] + (bias ? bias[gc] : 0.0f);
            if (relu) v = fmaxf(v, 0.0f);
            dst[gc] = v;
        }
    }
}

}  // namespace pn

PN_EXPORT int pn_linear_f32(const float* x, int64_t ldx, int64_t x_bstride, const float* w, int64_t w_bstride,
                            const float* bias, int64_t bias_bstride, int relu, int B, int64_t rows, int cin, int cout,
                            float* y, int64_t ldy, int64_t y_bstride, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(x && w && y, PN_ERR_BAD_ARG, "pn_linear_f32: null pointer");
    PN_REQUIRE(B > 0 && rows > 0 && cin > 0 && cout > 0, PN_ERR_BAD_ARG, "pn_linear_f32: sizes must be positive");
    PN_REQUIRE(ldx >= cin && ldy >= cout, PN_ERR_BAD_ARG, "pn_linear_f32: leading dimension smaller than the row (ldx=%lld cin=%d ldy=%lld cout=%d)",
               (long long)ldx, cin, (long long)ldy, cout);
    PN_REQUIRE(B <= 65535, PN_ERR_UNSUPPORTED, "pn_linear_f32: B=%d exceeds 65535", B);
    // bit 0: x rows can be read with 16-byte loads; bit 1: same for w rows
    int vec_ok = 0;
    if (((uintptr_t)x % 16 == 0) && (ldx % 4 == 0) && (x_bstride % 4 == 0)) vec_ok |= 1;
    if (((uintptr_t)w % 16 == 0) && (cin % 4 == 0) && (w_bstride % 4 == 0)) vec_ok |= 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (cout > 32) {
        constexpr int BM = 128, BN = 64;
        dim3 grid((unsigned)ceil_div(rows, BM), (unsigned)ceil_div(cout, BN), (unsigned)B);
        linear_kernel<BM, BN, 16, 8, 4><<<grid, 256, 0, st>>>(x, ldx, x_bstride, w, w_bstride, bias, bias_bstride, relu, rows, cin, cout,
                                                              y, ldy, y_bstride, vec_ok);
    } else {
        constexpr int BM = 128, BN = 32;
        dim3 grid((unsigned)ceil_div(rows, BM), (unsigned)ceil_div(cout, BN), (unsigned)B);
        linear_kernel<BM, BN, 16, 4, 4><<<grid, 256, 0, st>>>(x, ldx, x_bstride, w, w_bstride, bias, bias_bstride, relu, rows, cin, cout,
                                                              y, ldy, y_bstride, vec_ok);
    }
    return finish_launch("pn_linear_f32");
}
