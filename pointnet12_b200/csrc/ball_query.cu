// ball_query.cu -- query_ball_point (reference: model/pointnet_util.py:87-107) and square_distance (:19-40).
//
// The reference builds a [B,S,N] distance cube and an int64 index cube and full-sorts the latter;
// what it returns is simply "the first nsample in-ball points in ascending original index, padded
// with the first hit".  Here one warp owns a query and scans the cloud in index order, 32 points per
// step, with a ballot + prefix popcount to keep hits ordered, and stops as soon as nsample hits are
// found.  The points of a tile are staged once per CTA in shared memory as float4 (x, y, z, |p|^2) so
// that a lane fetches its point with a single conflict-free 16-byte load.  Nothing of size S*N ever
// exists.  Membership uses the reference's expansion formula bit for bit (common.cuh).
#include "common.cuh"

namespace pn {

constexpr int kBqThreads = 256;
constexpr int kBqWarps = kBqThreads / 32;

// QPW queries per warp share every shared-memory read; two 32-point chunks are tested per loop trip so that each
// warp has 2*QPW independent load -> distance -> ballot chains in flight (the loop is latency-, not issue-bound).
template <int QPW, int TILE>
__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC,
                  const float* __restrict__ qxyz, int64_t qB, int64_t qN, int64_t qC, int N, int S,
                  float radius2, int K, int64_t* __restrict__ out) {
    extern __shared__ __align__(16) float4 tile[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * kBqWarps + warp) * QPW;
    const float* __restrict__ p = xyz + (int64_t)b * xB;

    float ax[QPW], ay[QPW], az[QPW], sa[QPW];
    int cnt[QPW];
    int64_t first[QPW];
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        const int s = q0 + q;
        const bool ok = s < S;
        const float* a = qxyz + (int64_t)b * qB + (int64_t)(ok ? s : 0) * qN;
        ax[q] = a[0];
        ay[q] = a[qC];
        az[q] = a[2 * qC];
        sa[q] = sqnorm3(ax[q], ay[q], az[q]);
        cnt[q] = ok ? 0 : K;  // out-of-range queries are born finished
        first[q] = N;
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int t0 = 0; t0 < N; t0 += TILE) {
        const int tn = min(TILE, N - t0);
        for (int i = threadIdx.x; i < tn; i += kBqThreads) {
            const float* r = p + (int64_t)(t0 + i) * xN;
            const float x = r[0], y = r[xC], z = r[2 * xC];
            tile[i] = make_float4(x, y, z, sqnorm3(x, y, z));
        }
        __syncthreads();
        bool open = false;
#pragma unroll
        for (int q = 0; q < QPW; ++q) open |= cnt[q] < K;
        if (open) {  // warp-uniform
            for (int c = 0; c < tn; c += 64) {
                const int i0 = c + lane, i1 = c + 32 + lane;
                // out-of-tile lanes read a far-away dummy point (never a hit: d = +inf > radius2)
                const float4 v0 = i0 < tn ? tile[i0] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
                const float4 v1 = i1 < tn ? tile[i1] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
                bool any_open = false;
#pragma unroll
                for (int q = 0; q < QPW; ++q) {
                    const float d0 = sqdist_expand(ax[q], ay[q], az[q], sa[q], v0.x, v0.y, v0.z, v0.w);
                    const float d1 = sqdist_expand(ax[q], ay[q], az[q], sa[q], v1.x, v1.y, v1.z, v1.w);
                    const bool h0 = !(d0 > radius2), h1 = !(d1 > radius2);
                    const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                    if ((m0 | m1) && cnt[q] < K) {
                        int64_t* __restrict__ o = out + ((int64_t)b * S + (q0 + q)) * K;
                        if (cnt[q] == 0) first[q] = t0 + c + (m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1);
                        const int p0 = cnt[q] + __popc(m0 & lt_mask);
                        if (h0 && p0 < K) o[p0] = t0 + i0;
                        const int n0 = cnt[q] + __popc(m0);
                        const int p1 = n0 + __popc(m1 & lt_mask);
                        if (h1 && p1 < K) o[p1] = t0 + i1;
                        cnt[q] = n0 + __popc(m1);
                    }
                    any_open |= cnt[q] < K;
                }
                if (!any_open) break;
            }
        }
        open = false;
#pragma unroll
        for (int q = 0; q < QPW; ++q) open |= cnt[q] < K;
        // leave early when every query of the CTA is complete
        if (!__syncthreads_or(open)) break;
    }
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        if (q0 + q >= S) continue;
        int64_t* __restrict__ o = out + ((int64_t)b * S + (q0 + q)) * K;
        for (int k = min(cnt[q], K) + lane; k < K; k += 32) o[k] = first[q];  // pad with the first hit (or N)
    }
}

__global__ void square_distance_kernel(const float* __restrict__ src, int64_t aB, int64_t aN, int64_t aC,
                                       const float* __restrict__ dst, int64_t bB, int64_t bN, int64_t bC, int N,
                                       int M, float* __restrict__ out) {
    const int b = blockIdx.z;
    const int i = blockIdx.y;
    const float* a = src + (int64_t)b * aB + (int64_t)i * aN;
    const float ax = a[0], ay = a[aC], az = a[2 * aC];
    const float sa = sqnorm3(ax, ay, az);
    float* __restrict__ o = out + ((int64_t)b * N + i) * M;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const float* q = dst + (int64_t)b * bB + (int64_t)j * bN;
        const float bx = q[0], by = q[bC], bz = q[2 * bC];
        o[j] = sqdist_expand(ax, ay, az, sa, bx, by, bz, sqnorm3(bx, by, bz));
    }
}

}  // namespace pn

PN_EXPORT int pn_ball_query_f32(const float* xyz, int64_t xB, int64_t xN, int64_t xC, const float* new_xyz,
                                int64_t qB, int64_t qN, int64_t qC, int B, int N, int S, float radius2, int nsample,
                                int64_t* out_idx, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && new_xyz && out_idx, PN_ERR_BAD_ARG, "pn_ball_query_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, PN_ERR_BAD_ARG,
               "pn_ball_query_f32: B, N, S, nsample must be positive (got %d, %d, %d, %d)", B, N, S, nsample);
    PN_REQUIRE(B <= 65535, PN_ERR_UNSUPPORTED, "pn_ball_query_f32: B=%d exceeds 65535", B);
    cudaStream_t st = (cudaStream_t)stream;
    // Few queries: one per warp so that the grid still covers the SMs; many: two per warp share every tile read.
    const int64_t total_q = (int64_t)B * S;
    if (total_q >= 4096) {
        constexpr int TILE = 4096;
        auto kern = ball_query_kernel<2, TILE>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16);
        dim3 grid((unsigned)ceil_div(S, kBqWarps * 2), (unsigned)B);
        kern<<<grid, kBqThreads, TILE * 16, st>>>(xyz, xB, xN, xC, new_xyz, qB, qN, qC, N, S, radius2, nsample, out_idx);
    } else {
        constexpr int TILE = 2048;
        auto kern = ball_query_kernel<1, TILE>;
        dim3 grid((unsigned)ceil_div(S, kBqWarps), (unsigned)B);
        kern<<<grid, kBqThreads, TILE * 16, st>>>(xyz, xB, xN, xC, new_xyz, qB, qN, qC, N, S, radius2, nsample, out_idx);
    }
    return finish_launch("pn_ball_query_f32");
}

PN_EXPORT int pn_square_distance_f32(const float* src, int64_t aB, int64_t aN, int64_t aC, const float* dst, int64_t bB,
                                     int64_t bN, int64_t bC, int B, int N, int M, float* out, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(src && dst && out, PN_ERR_BAD_ARG, "pn_square_distance_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && M > 0, PN_ERR_BAD_ARG, "pn_square_distance_f32: B, N, M must be positive");
    PN_REQUIRE(B <= 65535 && N <= 65535, PN_ERR_UNSUPPORTED, "pn_square_distance_f32: B and N are limited to 65535");
    const int64_t gx = ceil_div(M, 256);
    dim3 grid((unsigned)(gx < 64 ? gx : 64), (unsigned)N, (unsigned)B);
    square_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, aB, aN, aC, dst, bB, bN, bC, N, M, out);
    return finish_launch("pn_square_distance_f32");
}
