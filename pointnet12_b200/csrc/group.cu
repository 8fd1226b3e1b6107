// group.cu -- gather / grouping / pooling / interpolation kernels.
//   index_points            model/pointnet_util.py:43-60
//   group (gather+recentre) model/pointnet_util.py:127-131 (SSG order), :243-247 (MSG order)
//   group_max               model/pointnet_util.py:199, :256; pointnet.py:35,74,122
//   three_nn / interpolate  model/pointnet_util.py:295-301 (+ concat of :303-307)
//   log_softmax             model/pointnet2.py:174; pointnet.py:251
// All of them are HBM/L2-bound data movement: coalesced along the channel dimension, 16-byte vector
// accesses where alignment allows, grids sized from the element count.
#include <math_constants.h>

#include "common.cuh"

namespace pn {

// ------------------------------------------------------------------------------------------------
__global__ void index_points_kernel(const float* __restrict__ points, int64_t pB, int64_t pN, int64_t pC, int N, int C,
                                    const int64_t* __restrict__ idx, int64_t M, int64_t total, float* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t row = e / C;
        const int64_t b = row / M;
        int64_t j = idx[row];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);  // memory safety only; valid inputs are never clamped
        out[e] = points[b * pB + j * pN + c * pC];
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void group_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC,
                             const float* __restrict__ feat, int64_t fB, int64_t fN, int64_t fC, int D,
                             const float* __restrict__ qxyz, int64_t qB, int64_t qN, int64_t qC,
                             const int64_t* __restrict__ idx, int N, int S, int K, int msg_order, int64_t total,
                             float* __restrict__ out, int64_t ldo) {
    const int C = 3 + D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t row = e / C;  // (b, s, k)
        const int64_t bs = row / K;
        const int64_t b = bs / S;
        const int s = (int)(bs % S);
        int64_t j = idx[row];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        const int cx = msg_order ? c - D : c;  // channel inside the xyz block, if 0 <= cx < 3
        float v;
        if (cx >= 0 && cx < 3) {
            v = __fsub_rn(xyz[b * xB + j * xN + cx * xC], qxyz[b * qB + (int64_t)s * qN + cx * qC]);
        } else {
            const int cf = msg_order ? c : c - 3;
            v = feat[b * fB + j * fN + cf * fC];
        }
        out[row * ldo + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void group_max_kernel(const float* __restrict__ x, int64_t ldx, int64_t groups, int K, int C,
                                 float* __restrict__ y, int64_t ldy) {
    const int64_t total = groups * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t g = e / C;
        const float* __restrict__ r = x + g * K * ldx + c;
        float m = r[0];
        for (int k = 1; k < K; ++k) m = fmaxf(m, r[(int64_t)k * ldx]);
        y[g * ldy + c] = m;
    }
}

// Large-K variant (global max over the points of a cloud): one CTA per (group, 32-channel slab).
__global__ void __launch_bounds__(256)
group_max_tall_kernel(const float* __restrict__ x, int64_t ldx, int K, int C, float* __restrict__ y, int64_t ldy) {
    __shared__ float red[8][33];
    const int64_t g = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    float m = -CUDART_INF_F;
    if (c < C) {
        const float* __restrict__ r = x + g * K * ldx + c;
        for (int k = ry; k < K; k += 8) m = fmaxf(m, r[(int64_t)k * ldx]);
    }
    red[ry][threadIdx.x & 31] = m;
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][threadIdx.x]);
        y[g * ldy + c] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// 3-NN: one thread per query point, sources staged in shared memory as (x, y, z, |p|^2); every lane of
// a warp reads the same source (broadcast).  Top-3 kept in registers, ordered by (distance, index).
constexpr int kNnThreads = 128;
constexpr int kNnTile = 1024;

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float* __restrict__ xyz1, int64_t aB, int64_t aN, int64_t aC, const float* __restrict__ xyz2,
                int64_t bB, int64_t bN, int64_t bC, int N, int S, int64_t* __restrict__ idx,
                float* __restrict__ weight) {
    __shared__ float4 tile[kNnTile];
    const int b = blockIdx.y;
    const int n = blockIdx.x * kNnThreads + threadIdx.x;
    const bool ok = n < N;
    const float* a = xyz1 + (int64_t)b * aB + (int64_t)(ok ? n : 0) * aN;
    const float ax = a[0], ay = a[aC], az = a[2 * aC];
    const float sa = sqnorm3(ax, ay, az);
    float d0 = CUDART_INF_F, d1 = CUDART_INF_F, d2 = CUDART_INF_F;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int t0 = 0; t0 < S; t0 += kNnTile) {
        const int tn = min(kNnTile, S - t0);
        if (t0) __syncthreads();
        for (int i = threadIdx.x; i < tn; i += kNnThreads) {
            const float* r = xyz2 + (int64_t)b * bB + (int64_t)(t0 + i) * bN;
            const float x = r[0], y = r[bC], z = r[2 * bC];
            tile[i] = make_float4(x, y, z, sqnorm3(x, y, z));
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < tn; ++j) {
            const float4 v = tile[j];
            const float d = sqdist_expand(ax, ay, az, sa, v.x, v.y, v.z, v.w);
            if (d < d2) {
                const int jj = t0 + j;
                if (d < d1) {
                    d2 = d1;
                    i2 = i1;
                    if (d < d0) {
                        d1 = d0;
                        i1 = i0;
                        d0 = d;
                        i0 = jj;
                    } else {
                        d1 = d;
                        i1 = jj;
                    }
                } else {
                    d2 = d;
                    i2 = jj;
                }
            }
        }
    }
    if (!ok) return;
    // dists[dists < 1e-10] = 1e-10 ; weight = 1/d ; weight /= sum(weight)   (pointnet_util.py:298-300)
    const float c0 = d0 < 1e-10f ? 1e-10f : d0, c1 = d1 < 1e-10f ? 1e-10f : d1, c2 = d2 < 1e-10f ? 1e-10f : d2;
    const float w0 = __fdiv_rn(1.0f, c0), w1 = __fdiv_rn(1.0f, c1), w2 = __fdiv_rn(1.0f, c2);
    const float norm = __fadd_rn(__fadd_rn(w0, w1), w2);
    const int64_t o = ((int64_t)b * N + n) * 3;
    idx[o] = i0;
    idx[o + 1] = i1;
    idx[o + 2] = i2;
    weight[o] = __fdiv_rn(w0, norm);
    weight[o + 1] = __fdiv_rn(w1, norm);
    weight[o + 2] = __fdiv_rn(w2, norm);
}

// Weighted gather + concat: one warp per output row, lanes across channels.
__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float* __restrict__ p1, int64_t p1B, int64_t p1N, int64_t p1C, int D1,
                         const float* __restrict__ p2, int64_t p2B, int64_t p2N, int64_t p2C, int D2, int S,
                         const int64_t* __restrict__ idx, const float* __restrict__ weight, int N, int64_t rows,
                         float* __restrict__ out, int64_t ldo, int64_t o_bstride) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const int64_t b = row / N;
        const int64_t n = row % N;
        float* __restrict__ o = out + b * o_bstride + n * ldo;
        if (p1) {
            const float* __restrict__ r = p1 + b * p1B + n * p1N;
            for (int c = lane; c < D1; c += 32) o[c] = r[c * p1C];
        }
        int64_t j0 = idx[row * 3], j1 = idx[row * 3 + 1], j2 = idx[row * 3 + 2];
        j0 = min(max(j0, (int64_t)0), (int64_t)S - 1);
        j1 = min(max(j1, (int64_t)0), (int64_t)S - 1);
        j2 = min(max(j2, (int64_t)0), (int64_t)S - 1);
        const float w0 = weight[row * 3], w1 = weight[row * 3 + 1], w2 = weight[row * 3 + 2];
        const float* __restrict__ r0 = p2 + b * p2B + j0 * p2N;
        const float* __restrict__ r1 = p2 + b * p2B + j1 * p2N;
        const float* __restrict__ r2 = p2 + b * p2B + j2 * p2N;
        for (int c = lane; c < D2; c += 32) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(r0[c * p2C], w0), __fmul_rn(r1[c * p2C], w1)),
                                      __fmul_rn(r2[c * p2C], w2));
            o[D1 + c] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
log_softmax_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, float* __restrict__ y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const float* __restrict__ r = x + row * ldx;
        float m = -CUDART_INF_F;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, r[c]);
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s += expf(r[c] - m);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float ls = logf(s);
        for (int c = lane; c < C; c += 32) y[row * ldy + c] = (r[c] - m) - ls;
    }
}

// pred.argmax(-1) of the evaluation loop (pcdseg.py:75): the label of every point as one byte; the first maximum wins, as
// torch.argmax does.  One thread per row (rows of <= 64 classes are a handful of 16-byte loads).
__global__ void __launch_bounds__(256)
argmax_labels_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, unsigned char* __restrict__ y) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
        const float* __restrict__ r = x + row * ldx;
        float m = r[0];
        int best = 0;
        for (int c = 1; c < C; ++c) {
            const float v = r[c];
            if (v > m) {
                m = v;
                best = c;
            }
        }
        y[row] = (unsigned char)best;
    }
}

static inline unsigned grid_for(int64_t total, int threads, int per_sm = 8) {
    const int64_t want = ceil_div(total, threads);
    const int64_t cap = 148LL * per_sm;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace pn

PN_EXPORT int pn_index_points_f32(const float* points, int64_t pB, int64_t pN, int64_t pC, int B, int N, int C,
                                  const int64_t* idx, int64_t M, float* out, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(points && idx && out, PN_ERR_BAD_ARG, "pn_index_points_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && C > 0 && M > 0, PN_ERR_BAD_ARG, "pn_index_points_f32: sizes must be positive");
    const int64_t total = (int64_t)B * M * C;
    cudaError_t e = launch_kernel(index_points_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, points, pB, pN,
                               pC, N, C, idx, M, total, out);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("pn_index_points_f32: launch failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return PN_OK;
}

PN_EXPORT int pn_group_f32(const float* xyz, int64_t xB, int64_t xN, int64_t xC, const float* feat, int64_t fB,
                           int64_t fN, int64_t fC, int D, const float* new_xyz, int64_t qB, int64_t qN, int64_t qC,
                           const int64_t* idx, int B, int N, int S, int K, int msg_order, float* out, int64_t ldo,
                           pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && new_xyz && idx && out, PN_ERR_BAD_ARG, "pn_group_f32: null pointer");
    PN_REQUIRE((feat != nullptr) == (D > 0) && D >= 0, PN_ERR_BAD_ARG, "pn_group_f32: feat pointer and D=%d disagree", D);
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0, PN_ERR_BAD_ARG, "pn_group_f32: sizes must be positive");
    PN_REQUIRE(ldo >= 3 + D, PN_ERR_BAD_ARG, "pn_group_f32: ldo=%lld smaller than 3+D=%d", (long long)ldo, 3 + D);
    const int64_t total = (int64_t)B * S * K * (3 + D);
    group_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(xyz, xB, xN, xC, feat, fB, fN, fC, D, new_xyz, qB, qN,
                                                                         qC, idx, N, S, K, msg_order, total, out, ldo);
    return finish_launch("pn_group_f32");
}

PN_EXPORT int pn_group_max_f32(const float* x, int64_t ldx, int64_t groups, int K, int C, float* y, int64_t ldy,
                               pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(x && y, PN_ERR_BAD_ARG, "pn_group_max_f32: null pointer");
    PN_REQUIRE(groups > 0 && K > 0 && C > 0 && ldx >= C && ldy >= C, PN_ERR_BAD_ARG, "pn_group_max_f32: bad sizes");
    if (K >= 256 && groups <= 65535) {
        dim3 grid((unsigned)ceil_div(C, 32), (unsigned)groups);
        group_max_tall_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, K, C, y, ldy);
    } else {
        group_max_kernel<<<grid_for(groups * C, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, groups, K, C, y, ldy);
    }
    return finish_launch("pn_group_max_f32");
}

PN_EXPORT int pn_three_nn_f32(const float* xyz1, int64_t aB, int64_t aN, int64_t aC, const float* xyz2, int64_t bB,
                              int64_t bN, int64_t bC, int B, int N, int S, int64_t* idx, float* weight,
                              pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz1 && xyz2 && idx && weight, PN_ERR_BAD_ARG, "pn_three_nn_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S >= 3, PN_ERR_BAD_ARG, "pn_three_nn_f32: need B, N > 0 and S >= 3 (got %d, %d, %d)", B, N, S);
    PN_REQUIRE(B <= 65535, PN_ERR_UNSUPPORTED, "pn_three_nn_f32: B=%d exceeds 65535", B);
    dim3 grid((unsigned)ceil_div(N, kNnThreads), (unsigned)B);
    three_nn_kernel<<<grid, kNnThreads, 0, (cudaStream_t)stream>>>(xyz1, aB, aN, aC, xyz2, bB, bN, bC, N, S, idx, weight);
    return finish_launch("pn_three_nn_f32");
}

PN_EXPORT int pn_three_interpolate_f32(const float* points1, int64_t p1B, int64_t p1N, int64_t p1C, int D1,
                                       const float* points2, int64_t p2B, int64_t p2N, int64_t p2C, int D2, int S,
                                       const int64_t* idx, const float* weight, int B, int N, float* out, int64_t ldo,
                                       int64_t o_bstride, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(points2 && idx && weight && out, PN_ERR_BAD_ARG, "pn_three_interpolate_f32: null pointer");
    PN_REQUIRE((points1 != nullptr) == (D1 > 0) && D1 >= 0, PN_ERR_BAD_ARG,
               "pn_three_interpolate_f32: points1 pointer and D1=%d disagree", D1);
    PN_REQUIRE(B > 0 && N > 0 && D2 > 0 && S > 0 && ldo >= D1 + D2, PN_ERR_BAD_ARG, "pn_three_interpolate_f32: bad sizes");
    const int64_t rows = (int64_t)B * N;
    three_interpolate_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        points1, p1B, p1N, p1C, D1, points2, p2B, p2N, p2C, D2, S, idx, weight, N, rows, out, ldo, o_bstride);
    return finish_launch("pn_three_interpolate_f32");
}

PN_EXPORT int pn_log_softmax_f32(const float* x, int64_t ldx, int64_t rows, int C, float* y, int64_t ldy,
                                 pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(x && y, PN_ERR_BAD_ARG, "pn_log_softmax_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldx >= C && ldy >= C, PN_ERR_BAD_ARG, "pn_log_softmax_f32: bad sizes");
    log_softmax_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, y, ldy);
    return finish_launch("pn_log_softmax_f32");
}

PN_EXPORT int pn_argmax_labels_u8(const float* x, int64_t ldx, int64_t rows, int C, uint8_t* labels, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(x && labels, PN_ERR_BAD_ARG, "pn_argmax_labels_u8: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && C <= 256 && ldx >= C, PN_ERR_BAD_ARG, "pn_argmax_labels_u8: bad sizes (1 <= C <= 256)");
    argmax_labels_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, labels);
    return finish_launch("pn_argmax_labels_u8");
}
