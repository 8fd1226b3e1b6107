// ball_grid.cu -- query_ball_point (reference: model/pointnet_util.py:87-107) through a uniform-grid bucket pass.
//
// What the reference returns is "the first nsample in-ball points in ascending ORIGINAL index, padded with the
// first hit" (it sorts an index cube).  The ordered scan of ball_query.cu reproduces that with an early exit, which is
// cheap where the cloud is dense (32 hits arrive after a short prefix) and expensive where it is sparse (the whole
// cloud is scanned and fewer than 32 hits are found) -- and farthest-point centroids favour sparse regions.  Here:
//
//   build  (one CTA per cloud, depends on xyz and the radius only -> runs beside farthest-point sampling):
//          bounding box -> cell size >= radius (with slack for fp32 rounding, see grid_plan) -> counting sort of
//          the points by cell (x fastest) into float4 (x, y, z, original index).
//   query  (one warp per centroid): the 27 neighbouring cells are 9 contiguous runs of the sorted array.
//          few candidates  -> test them all (coalesced 16-byte loads), mark hits in a per-warp shared-memory BITMAP
//                             over original indices, then read the first nsample set bits: ascending order for free;
//          many candidates -> the ball is dense, so the ordered scan over the raw cloud ends after a short prefix.
//
// Membership uses the reference's expansion formula bit for bit (common.cuh: sqdist_expand), so both paths return
// exactly what the scan returns; the grid only decides WHICH points are tested, conservatively.
#include "common.cuh"

namespace pn {

constexpr int kGridMaxAxis = 32;
constexpr int kGridMaxCells = 32768;
constexpr int kGridHdrWords = 16;   // lo[3], inv[3], G[3], ncell, pad
constexpr int kGridBuildThreads = 1024;

struct GridHdr {
    float lo[3];
    float inv[3];
    int g[3];
    int ncell;
    int pad[6];
};
static_assert(sizeof(GridHdr) == kGridHdrWords * 4, "GridHdr layout");

__host__ __device__ inline size_t grid_cloud_bytes(int N) {
    // header | cell_start[kGridMaxCells + 1] (padded to 16 bytes) | sorted float4[N]
    return (size_t)kGridHdrWords * 4 + (((size_t)kGridMaxCells + 1 + 3) / 4) * 16 + (size_t)N * 16;
}
__device__ __forceinline__ const GridHdr* grid_hdr(const unsigned char* ws) { return reinterpret_cast<const GridHdr*>(ws); }
__device__ __forceinline__ const int* grid_cells(const unsigned char* ws) {
    return reinterpret_cast<const int*>(ws + kGridHdrWords * 4);
}
__device__ __forceinline__ const float4* grid_sorted(const unsigned char* ws) {
    return reinterpret_cast<const float4*>(ws + kGridHdrWords * 4 + (((size_t)kGridMaxCells + 1 + 3) / 4) * 16);
}

// cell coordinate along one axis; monotone in v, NOT clamped above/below the grid except to keep the int sane
__device__ __forceinline__ int grid_coord(float v, float lo, float inv) {
    const float t = __fmul_rn(__fsub_rn(v, lo), inv);
    return (int)fminf(fmaxf(t, -2.0f), (float)(kGridMaxAxis + 2));   // NaN -> -2 (fmaxf drops the NaN)
}

__global__ void __launch_bounds__(kGridBuildThreads, 1)
ball_grid_build_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC, int N, float radius2,
                       unsigned char* __restrict__ ws_all, size_t ws_stride) {
    extern __shared__ int hist[];                  // kGridMaxCells counters / cursors
    __shared__ float red[6][kGridBuildThreads / 32];
    __shared__ GridHdr hdr;
    __shared__ int warp_tot[kGridBuildThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ p = xyz + (int64_t)b * xB;
    unsigned char* ws = ws_all + (size_t)b * ws_stride;

    // ---- 1. bounding box
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = tid; i < N; i += kGridBuildThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = p[(int64_t)i * xN + a * xC];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            red[a][warp] = mn[a];
            red[3 + a][warp] = mx[a];
        }
    }
    __syncthreads();
    // ---- 2. plan: cell size h >= sqrt(r^2 + E) * 1.001 where E bounds the rounding error of the reference's
    // expansion formula (a few ulps of |a|^2 + |b|^2), so that "in ball by the formula" implies "within one cell".
    if (tid == 0) {
        float m2 = 0.0f;
        float lo[3], ext[3];
        for (int a = 0; a < 3; ++a) {
            float l = red[a][0], h = red[3 + a][0];
            for (int w = 1; w < kGridBuildThreads / 32; ++w) {
                l = fminf(l, red[a][w]);
                h = fmaxf(h, red[3 + a][w]);
            }
            if (!(l <= h)) { l = 0.0f; h = 0.0f; }   // empty / all-NaN cloud
            lo[a] = l;
            ext[a] = h - l;
            const float m = fmaxf(fabsf(l), fabsf(h));
            m2 += m * m;
        }
        const float slack = m2 * (1.0f / 262144.0f);             // 2^-18 * max|p|^2  >>  rounding of the formula
        const float h0 = sqrtf(fmaxf(radius2, 0.0f) + slack) * 1.001f + 1e-30f;
        int ncell = 1;
        for (int a = 0; a < 3; ++a) {
            const float cells = ext[a] / h0;
            int g = cells < (float)(kGridMaxAxis - 1) ? (int)cells + 1 : kGridMaxAxis;
            float h = h0;
            if (g == kGridMaxAxis) h = fmaxf(h0, ext[a] / (float)kGridMaxAxis * 1.001f);
            hdr.lo[a] = lo[a];
            hdr.inv[a] = 1.0f / h;
            hdr.g[a] = g;
            ncell *= g;
        }
        hdr.ncell = ncell;
        for (int i = 0; i < 6; ++i) hdr.pad[i] = 0;
        *reinterpret_cast<GridHdr*>(ws) = hdr;
    }
    __syncthreads();
    const int gx = hdr.g[0], gy = hdr.g[1], gz = hdr.g[2], ncell = hdr.ncell;
    const float lx = hdr.lo[0], ly = hdr.lo[1], lz = hdr.lo[2], ix = hdr.inv[0], iy = hdr.inv[1], iz = hdr.inv[2];
    auto cell_of = [&](float x, float y, float z) {
        const int cx = min(max(grid_coord(x, lx, ix), 0), gx - 1);
        const int cy = min(max(grid_coord(y, ly, iy), 0), gy - 1);
        const int cz = min(max(grid_coord(z, lz, iz), 0), gz - 1);
        return (cz * gy + cy) * gx + cx;
    };
    // ---- 3. histogram
    for (int c = tid; c < ncell; c += kGridBuildThreads) hist[c] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += kGridBuildThreads) {
        const float x = p[(int64_t)i * xN], y = p[(int64_t)i * xN + xC], z = p[(int64_t)i * xN + 2 * xC];
        atomicAdd(&hist[cell_of(x, y, z)], 1);
    }
    __syncthreads();
    // ---- 4. exclusive scan over the cells (each thread owns a contiguous run)
    int* __restrict__ cell_start = reinterpret_cast<int*>(ws + kGridHdrWords * 4);
    const int per = (ncell + kGridBuildThreads - 1) / kGridBuildThreads;
    const int c0 = tid * per, c1 = min(c0 + per, ncell);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += hist[c];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane], inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        warp_tot[lane] = inc - t;
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
    for (int c = c0; c < c1; ++c) {
        const int n = hist[c];
        hist[c] = run;
        cell_start[c] = run;
        run += n;
    }
    if (tid == 0) cell_start[ncell] = N;
    __syncthreads();
    // ---- 5. scatter (order inside a cell is arbitrary: the query orders hits by original index itself)
    float4* __restrict__ sorted = reinterpret_cast<float4*>(ws + kGridHdrWords * 4 + (((size_t)kGridMaxCells + 1 + 3) / 4) * 16);
    for (int i = tid; i < N; i += kGridBuildThreads) {
        const float x = p[(int64_t)i * xN], y = p[(int64_t)i * xN + xC], z = p[(int64_t)i * xN + 2 * xC];
        const int pos = atomicAdd(&hist[cell_of(x, y, z)], 1);
        sorted[pos] = make_float4(x, y, z, __int_as_float(i));
    }
}

// ------------------------------------------------------------------------------------------------ query
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
ball_query_grid_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC,
                       const float* __restrict__ qxyz, int64_t qB, int64_t qN, int64_t qC, int N, int S, int64_t total_q,
                       float radius2, int K, const unsigned char* __restrict__ ws_all, size_t ws_stride, int threshold,
                       int bm_words, int64_t* __restrict__ out) {
    extern __shared__ unsigned bitmaps[];   // WARPS x bm_words, zero between queries
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* __restrict__ bm = bitmaps + (size_t)warp * bm_words;
    for (int w = lane; w < bm_words; w += 32) bm[w] = 0u;
    __syncwarp();
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int64_t q = (int64_t)blockIdx.x * WARPS + warp; q < total_q; q += (int64_t)gridDim.x * WARPS) {
        const int b = (int)(q / S);
        const unsigned char* ws = ws_all + (size_t)b * ws_stride;
        const GridHdr* h = grid_hdr(ws);
        const int* __restrict__ cell_start = grid_cells(ws);
        const float4* __restrict__ sorted = grid_sorted(ws);
        const float* a = qxyz + (int64_t)b * qB + (q % S) * qN;
        const float ax = a[0], ay = a[qC], az = a[2 * qC];
        const float sa = sqnorm3(ax, ay, az);
        const int gx = h->g[0], gy = h->g[1], gz = h->g[2];
        const int cx = grid_coord(ax, h->lo[0], h->inv[0]);
        const int cy = grid_coord(ay, h->lo[1], h->inv[1]);
        const int cz = grid_coord(az, h->lo[2], h->inv[2]);
        // the 27 neighbouring cells = 9 runs along x; lane r < 9 owns run (dy, dz) = (r % 3 - 1, r / 3 - 1)
        int r_beg = 0, r_end = 0;
        if (lane < 9) {
            const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, gx - 1);
            if (y >= 0 && y < gy && z >= 0 && z < gz && x0 <= x1) {
                const int row = (z * gy + y) * gx;
                r_beg = cell_start[row + x0];
                r_end = cell_start[row + x1 + 1];
            }
        }
        int cand = r_end - r_beg;
#pragma unroll
        for (int o = 16; o; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);

        int64_t* __restrict__ o = out + q * K;
        int cnt = 0;
        int64_t first = N;
        if (cand <= threshold) {
            // ---- sparse neighbourhood: test every candidate, mark hits by original index
            int hits = 0;
            for (int r = 0; r < 9; ++r) {
                const int beg = __shfl_sync(0xffffffffu, r_beg, r), end = __shfl_sync(0xffffffffu, r_end, r);
                for (int i0 = beg; i0 < end; i0 += 32) {
                    const int i = i0 + lane;
                    bool hit = false;
                    int j = 0;
                    if (i < end) {
                        const float4 v = sorted[i];
                        const float d = sqdist_expand(ax, ay, az, sa, v.x, v.y, v.z, sqnorm3(v.x, v.y, v.z));
                        hit = !(d > radius2);
                        j = __float_as_int(v.w);
                    }
                    if (hit) atomicOr(&bm[j >> 5], 1u << (j & 31));
                    hits += __popc(__ballot_sync(0xffffffffu, hit));
                }
            }
            __syncwarp();
            if (hits > 0) {
                for (int w0 = 0; w0 < bm_words; w0 += 32) {
                    unsigned word = bm[w0 + lane];     // bm_words is a multiple of 32
                    bm[w0 + lane] = 0u;
                    const int pc = __popc(word);
                    int incl = pc;
#pragma unroll
                    for (int s = 1; s < 32; s <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, s);
                        if (lane >= s) incl += t;
                    }
                    const unsigned has = __ballot_sync(0xffffffffu, pc > 0);
                    if (cnt == 0 && has) {
                        const int src = __ffs(has) - 1;
                        const int lowest = (w0 + lane) * 32 + __ffs(word) - 1;
                        first = __shfl_sync(0xffffffffu, lowest, src);
                    }
                    int pos = cnt + incl - pc;
                    while (word && pos < K) {
                        o[pos++] = (int64_t)(w0 + lane) * 32 + (__ffs(word) - 1);
                        word &= word - 1;
                    }
                    cnt += __shfl_sync(0xffffffffu, incl, 31);
                    hits -= __shfl_sync(0xffffffffu, incl, 31);
                    if (cnt >= K || hits <= 0) {
                        // words not yet visited may still hold bits (cnt >= K): clear them
                        if (hits > 0)
                            for (int w = w0 + 32 + lane; w < bm_words; w += 32) bm[w] = 0u;
                        break;
                    }
                }
                __syncwarp();
            }
        } else {
            // ---- dense neighbourhood: ordered scan over the raw cloud, ends after a short prefix
            const float* __restrict__ p = xyz + (int64_t)b * xB;
            for (int c = 0; c < N && cnt < K; c += 64) {
                const int i0 = c + lane, i1 = c + 32 + lane;
                bool h0 = false, h1 = false;
                if (i0 < N) {
                    const float x = p[(int64_t)i0 * xN], y = p[(int64_t)i0 * xN + xC], z = p[(int64_t)i0 * xN + 2 * xC];
                    h0 = !(sqdist_expand(ax, ay, az, sa, x, y, z, sqnorm3(x, y, z)) > radius2);
                }
                if (i1 < N) {
                    const float x = p[(int64_t)i1 * xN], y = p[(int64_t)i1 * xN + xC], z = p[(int64_t)i1 * xN + 2 * xC];
                    h1 = !(sqdist_expand(ax, ay, az, sa, x, y, z, sqnorm3(x, y, z)) > radius2);
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                if (m0 | m1) {
                    if (cnt == 0) first = c + (m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1);
                    const int p0 = cnt + __popc(m0 & lt_mask);
                    if (h0 && p0 < K) o[p0] = i0;
                    const int n0 = cnt + __popc(m0);
                    const int p1 = n0 + __popc(m1 & lt_mask);
                    if (h1 && p1 < K) o[p1] = i1;
                    cnt = n0 + __popc(m1);
                }
            }
        }
        for (int k = min(cnt, K) + lane; k < K; k += 32) o[k] = first;   // pad with the first hit (or N)
    }
}

}  // namespace pn

PN_EXPORT size_t pn_ball_grid_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * pn::grid_cloud_bytes(N);
}

PN_EXPORT int pn_ball_grid_build_f32(const float* xyz, int64_t xB, int64_t xN, int64_t xC, int B, int N, float radius2,
                                     void* grid, size_t grid_bytes, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && grid, PN_ERR_BAD_ARG, "pn_ball_grid_build_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && radius2 >= 0.0f, PN_ERR_BAD_ARG, "pn_ball_grid_build_f32: B, N must be positive, radius2 >= 0");
    PN_REQUIRE(((uintptr_t)grid & 15) == 0, PN_ERR_ALIGNMENT, "pn_ball_grid_build_f32: grid must be 16-byte aligned");
    PN_REQUIRE(grid_bytes >= pn_ball_grid_bytes(B, N), PN_ERR_BAD_ARG,
               "pn_ball_grid_build_f32: grid buffer holds %zu bytes, %zu needed", grid_bytes, pn_ball_grid_bytes(B, N));
    auto kern = ball_grid_build_kernel;
    const size_t smem = (size_t)kGridMaxCells * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("pn_ball_grid_build_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    kern<<<B, kGridBuildThreads, smem, (cudaStream_t)stream>>>(xyz, xB, xN, xC, N, radius2, static_cast<unsigned char*>(grid),
                                                              grid_cloud_bytes(N));
    return finish_launch("pn_ball_grid_build_f32");
}

PN_EXPORT int pn_ball_query_grid_f32(const float* xyz, int64_t xB, int64_t xN, int64_t xC, const float* new_xyz, int64_t qB,
                                     int64_t qN, int64_t qC, int B, int N, int S, float radius2, int nsample,
                                     const void* grid, size_t grid_bytes, int threshold, int64_t* out_idx,
                                     pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && new_xyz && out_idx && grid, PN_ERR_BAD_ARG, "pn_ball_query_grid_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, PN_ERR_BAD_ARG,
               "pn_ball_query_grid_f32: B, N, S, nsample must be positive (got %d, %d, %d, %d)", B, N, S, nsample);
    PN_REQUIRE(grid_bytes >= pn_ball_grid_bytes(B, N), PN_ERR_BAD_ARG,
               "pn_ball_query_grid_f32: grid buffer holds %zu bytes, %zu needed", grid_bytes, pn_ball_grid_bytes(B, N));
    PN_REQUIRE(N <= 1048576, PN_ERR_UNSUPPORTED, "pn_ball_query_grid_f32: N=%d exceeds 1048576", N);
    if (threshold == 0) {
        // break-even between testing C candidates and scanning ~ 33 N / (hits + 1) points, hits ~ 0.155 C
        int t = 64;
        while ((int64_t)t * t < (int64_t)213 * N) t += 64;
        threshold = t;
    }
    const int bm_words = (int)ceil_div(ceil_div(N, 32), 32) * 32;
    const int64_t total_q = (int64_t)B * S;
    cudaStream_t st = (cudaStream_t)stream;
    auto launch = [&](auto kern, int warps) -> int {
        const size_t smem = (size_t)warps * bm_words * 4;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("pn_ball_query_grid_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        const int64_t ctas = ceil_div(total_q, warps);
        const int64_t cap = 148LL * 16;
        kern<<<(unsigned)(ctas < cap ? ctas : cap), warps * 32, smem, st>>>(
            xyz, xB, xN, xC, new_xyz, qB, qN, qC, N, S, total_q, radius2, nsample, static_cast<const unsigned char*>(grid),
            grid_cloud_bytes(N), threshold, bm_words, out_idx);
        return finish_launch("pn_ball_query_grid_f32");
    };
    if (N <= 32768) return launch(ball_query_grid_kernel<8>, 8);
    if (N <= 262144) return launch(ball_query_grid_kernel<4>, 4);
    return launch(ball_query_grid_kernel<1>, 1);
}
