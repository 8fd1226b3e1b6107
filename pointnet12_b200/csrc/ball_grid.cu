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

size_t ball_grid_cloud_bytes(int N) { return grid_cloud_bytes(N); }
size_t ball_grid_sorted_offset() { return (size_t)kGridHdrWords * 4 + (((size_t)kGridMaxCells + 1 + 3) / 4) * 16; }

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
constexpr int kGqTile = 2048;   // points per staged tile in the dense phase (32 KB as float4)

// One CTA = WARPS centroids of one cloud.
//   phase 1 (per warp): count the candidates of the 27 neighbouring cells; at most `threshold` -> test them all,
//                       128 per trip with the four 16-byte loads of a lane in flight together, hits -> bitmap ->
//                       first K set bits.  More -> the centroid is left pending.
//   phase 2 (per warp): a pending (dense) centroid scans the raw cloud in index order, 128 points per trip, until its
//                       K-th hit.
// The work of one CTA on WARPS centroids of cloud b (warp w: centroid (ax, ay, az), result row o); the warps are independent.
template <int WARPS>
__device__ __forceinline__ void gq_block(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC, int N, float radius2,
                                         int K, const unsigned char* __restrict__ ws, int threshold, int bm_words,
                                         unsigned char* gq_smem, int b, bool valid, float ax, float ay, float az,
                                         int64_t* __restrict__ o) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float sa = sqnorm3(ax, ay, az);
    int cnt = valid ? 0 : K;          // out-of-range warps are born finished
    int64_t first = N;
    bool pending = false;

    if (valid) {
        const GridHdr* h = grid_hdr(ws);
        const int* __restrict__ cell_start = grid_cells(ws);
        const int gx = h->g[0], gy = h->g[1], gz = h->g[2];
        const int cx = grid_coord(ax, h->lo[0], h->inv[0]);
        const int cy = grid_coord(ay, h->lo[1], h->inv[1]);
        const int cz = grid_coord(az, h->lo[2], h->inv[2]);
        // the 27 neighbouring cells = 9 runs along x; lane r < 9 owns run (dy, dz) = (r % 3 - 1, r / 3 - 1)
        int r_beg = 0, r_len = 0;
        if (lane < 9) {
            const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, gx - 1);
            if (y >= 0 && y < gy && z >= 0 && z < gz && x0 <= x1) {
                const int row = (z * gy + y) * gx;
                r_beg = cell_start[row + x0];
                r_len = cell_start[row + x1 + 1] - r_beg;
            }
        }
        int beg[9], len[9], cand = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            beg[r] = __shfl_sync(0xffffffffu, r_beg, r);
            len[r] = __shfl_sync(0xffffffffu, r_len, r);
            cand += len[r];
        }
        if (cand <= threshold) {
            // ---- sparse neighbourhood: every candidate is tested; hits are ordered through the bitmap
            unsigned* __restrict__ bm = reinterpret_cast<unsigned*>(gq_smem) + (size_t)warp * bm_words;
            for (int w = lane; w < bm_words; w += 32) bm[w] = 0u;
            __syncwarp();
            const float4* __restrict__ sorted = grid_sorted(ws);
            int hits = 0;
            for (int t0 = 0; t0 < cand; t0 += 128) {
                float4 v[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int off = t0 + u * 32 + lane, i = -1;
#pragma unroll
                    for (int r = 0; r < 9; ++r) {
                        if (i < 0 && off < len[r]) i = beg[r] + off;
                        off -= len[r];
                    }
                    ok[u] = i >= 0;
                    v[u] = ok[u] ? sorted[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float d = sqdist_expand(ax, ay, az, sa, v[u].x, v[u].y, v[u].z, sqnorm3(v[u].x, v[u].y, v[u].z));
                    const bool hit = ok[u] && !(d > radius2);
                    const int j = __float_as_int(v[u].w);
                    if (hit) atomicOr(&bm[j >> 5], 1u << (j & 31));
                    hits += __popc(__ballot_sync(0xffffffffu, hit));
                }
            }
            __syncwarp();
            for (int w0 = 0; w0 < bm_words && hits > 0 && cnt < K; w0 += 32) {
                unsigned word = bm[w0 + lane];     // bm_words is a multiple of 32
                const int pc = __popc(word);
                int incl = pc;
#pragma unroll
                for (int sh = 1; sh < 32; sh <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, sh);
                    if (lane >= sh) incl += t;
                }
                const unsigned has = __ballot_sync(0xffffffffu, pc > 0);
                if (cnt == 0 && has) {
                    const int lowest = (w0 + lane) * 32 + __ffs(word) - 1;
                    first = __shfl_sync(0xffffffffu, lowest, __ffs(has) - 1);
                }
                int pos = cnt + incl - pc;
                while (word && pos < K) {
                    o[pos++] = (int64_t)(w0 + lane) * 32 + (__ffs(word) - 1);
                    word &= word - 1;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                cnt += total;
                hits -= total;
            }
        } else {
            pending = true;
        }
    }

    // ---- dense neighbourhoods: the warp scans the raw cloud in index order by itself, 128 points per trip (the loads of a
    // trip in flight together), and stops at the K-th hit -- a dense ball has it after a short prefix.  (Round 1 staged tiles
    // of the cloud through shared memory for the whole CTA: ncu showed 51 % of all warp samples waiting at those CTA barriers,
    // seven warps idling while one scanned.)
    if (pending) {   // warp-uniform
        const float* __restrict__ p = xyz + (int64_t)b * xB;
        for (int c = 0; c < N && cnt < K; c += 128) {
            float px[4], py[4], pz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(c + u * 32 + lane, N - 1);
                const float* r = p + (int64_t)i * xN;
                px[u] = r[0];
                py[u] = r[xC];
                pz[u] = r[2 * xC];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = c + u * 32 + lane;
                const bool h = i < N && !(sqdist_expand(ax, ay, az, sa, px[u], py[u], pz[u], sqnorm3(px[u], py[u], pz[u])) > radius2);
                const unsigned m = __ballot_sync(0xffffffffu, h);
                if (m) {
                    if (cnt == 0) first = c + u * 32 + __ffs(m) - 1;
                    const int pos = cnt + __popc(m & lt_mask);
                    if (h && pos < K) o[pos] = i;
                    cnt += __popc(m);
                }
            }
        }
    }
    if (valid)
        for (int k = min(cnt, K) + lane; k < K; k += 32) o[k] = first;   // pad with the first hit (or N)
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
ball_query_grid_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC,
                       const float* __restrict__ qxyz, int64_t qB, int64_t qN, int64_t qC, int N, int S,
                       float radius2, int K, const unsigned char* __restrict__ ws_all, size_t ws_stride, int threshold,
                       int bm_words, const int* __restrict__ done, int64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char gq_smem[];
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    // (the grid may be capped: a CTA then walks several groups of WARPS centroids)
    for (int sg = blockIdx.x; sg * WARPS < S; sg += gridDim.x) {
        const int s = sg * WARPS + warp;
        // `done` (may be NULL): rows already produced by ball_query_stream_kernel; a group whose rows are all done is skipped
        bool valid = s < S;
        if (done) {
            if (valid && done[(int64_t)b * S + s]) valid = false;
            if (!__syncthreads_or(valid)) continue;
        }
        const float* a = qxyz + (int64_t)b * qB + (int64_t)(s < S ? s : 0) * qN;
        gq_block<WARPS>(xyz, xB, xN, xC, N, radius2, K, ws_all + (size_t)b * ws_stride, threshold, bm_words, gq_smem, b, valid, a[0],
                        a[qC], a[2 * qC], out + ((int64_t)b * S + (s < S ? s : 0)) * K);
    }
}

// The same search fed by a RUNNING farthest-point-sampling kernel: a persistent grid on the SMs that sampling leaves idle
// polls the progress feed of pn_fps_progress_f32 (one 8-byte word per centroid: index << 32 | 1) and answers the ball
// query of every centroid as soon as it exists, so that the level's grouping is (almost) complete when sampling ends.
// CTA c serves cloud c % B: rounds of WARPS consecutive centroids.  A warp that waits too long gives up and the CTA
// leaves; rows it did not finish keep done == 0 and are computed by ball_query_grid_kernel afterwards (correctness never
// depends on the two kernels actually running side by side).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
ball_query_stream_kernel(const float* __restrict__ xyz, int64_t xB, int64_t xN, int64_t xC,
                         const unsigned long long* __restrict__ seq, int B, int N, int S, int s_end, float radius2, int K,
                         const unsigned char* __restrict__ ws_all, size_t ws_stride, int threshold, int bm_words,
                         long long timeout_ns, int* __restrict__ done, int64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char gq_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x % B;
    const int per_cloud = gridDim.x / B;           // CTAs serving this cloud (the host launches a multiple of B)
    if ((int)blockIdx.x >= per_cloud * B) return;
    const float* __restrict__ p = xyz + (int64_t)b * xB;
    // centroids [s_end, S) are left to the follow-up kernel: the last few appear when sampling ends, and the whole GPU
    // answers them faster than the few CTAs of this grid
    for (int s0 = (blockIdx.x / B) * WARPS; s0 < s_end; s0 += per_cloud * WARPS) {
        const int s = s0 + warp;
        const bool valid = s < s_end;
        float ax = 0.f, ay = 0.f, az = 0.f;
        bool gave_up = false;
        if (valid) {
            unsigned long long w = 0ull;
            if (lane == 0) {
                const volatile unsigned long long* slot = seq + (int64_t)b * S + s;
                long long t0 = 0;
                while (((w = *slot) & 1ull) == 0ull) {
                    __nanosleep(200);
                    long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > timeout_ns) break;
                }
            }
            w = __shfl_sync(0xffffffffu, w, 0);
            gave_up = (w & 1ull) == 0ull;
            if (!gave_up) {
                int j = (int)(w >> 32);
                j = j < 0 ? 0 : (j >= N ? N - 1 : j);
                ax = p[(int64_t)j * xN];
                ay = p[(int64_t)j * xN + xC];
                az = p[(int64_t)j * xN + 2 * xC];
            }
        }
        if (__syncthreads_or(gave_up)) return;     // sampling is not running beside us: leave everything to the fallback
        gq_block<WARPS>(xyz, xB, xN, xC, N, radius2, K, ws_all + (size_t)b * ws_stride, threshold, bm_words, gq_smem, b, valid, ax,
                        ay, az, out + ((int64_t)b * S + (valid ? s : 0)) * K);
        __syncwarp();
        if (valid && lane == 0) {
            __threadfence();                       // the row before its flag
            done[(int64_t)b * S + s] = 1;
        }
        __syncthreads();                           // shared memory is reused by the next round
    }
}


}  // namespace pn

// Break-even between testing C candidates of the 27 cells (hits ~ 0.155 C) and scanning ~ 33 N / (hits + 1) points of the raw
// cloud from global memory until the K-th hit.  Measured at N = 24000, B = 8, radius 0.1 for thresholds 1280 / 2560 / 5120 / 8192 /
// always cells: 151 / 80 / 60 / 70 / 215 us alone (a warp that scans most of the cloud by itself is a long tail), and 0.343 / 0.3415 /
// 0.346 / 0.348 ms per batch with ten batches in flight (within 2 %): C^2 ~ 1024 N.
static int gq_auto_threshold(int N) {
    int t = 64;
    while ((int64_t)t * t < (int64_t)1024 * N) t += 64;
    return t;
}

PN_EXPORT size_t pn_ball_grid_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * pn::grid_cloud_bytes(N);
}

PN_EXPORT int pn_ball_grid_order(const void* grid, int N, const int32_t** order, int64_t* estride, int64_t* bstride) {
    PN_REQUIRE(grid && order && estride && bstride && N > 0, PN_ERR_BAD_ARG, "pn_ball_grid_order: bad arguments");
    const size_t sorted_off = (size_t)pn::kGridHdrWords * 4 + (((size_t)pn::kGridMaxCells + 1 + 3) / 4) * 16;
    *order = reinterpret_cast<const int32_t*>(static_cast<const unsigned char*>(grid) + sorted_off) + 3;   // float4.w
    *estride = 4;
    *bstride = (int64_t)(pn::grid_cloud_bytes(N) / 4);
    return PN_OK;
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
                                     const void* grid_ws, size_t grid_bytes, int threshold, const int32_t* done,
                                     int64_t* out_idx, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && new_xyz && out_idx && grid_ws, PN_ERR_BAD_ARG, "pn_ball_query_grid_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, PN_ERR_BAD_ARG,
               "pn_ball_query_grid_f32: B, N, S, nsample must be positive (got %d, %d, %d, %d)", B, N, S, nsample);
    PN_REQUIRE(grid_bytes >= pn_ball_grid_bytes(B, N), PN_ERR_BAD_ARG,
               "pn_ball_query_grid_f32: grid buffer holds %zu bytes, %zu needed", grid_bytes, pn_ball_grid_bytes(B, N));
    PN_REQUIRE(N <= 1048576, PN_ERR_UNSUPPORTED, "pn_ball_query_grid_f32: N=%d exceeds 1048576", N);
    PN_REQUIRE(B <= 65535, PN_ERR_UNSUPPORTED, "pn_ball_query_grid_f32: B=%d exceeds 65535", B);
    if (threshold == 0) threshold = gq_auto_threshold(N);
    const int bm_words = (int)ceil_div(ceil_div(N, 32), 32) * 32;
    cudaStream_t st = (cudaStream_t)stream;
    auto launch = [&](auto kern, int warps) -> int {
        size_t smem = (size_t)warps * bm_words * 4;
        if (smem < (size_t)kGqTile * 16) smem = (size_t)kGqTile * 16;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("pn_ball_query_grid_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        int64_t gx = ceil_div(S, warps);
        dim3 grid((unsigned)gx, (unsigned)B);
        e = launch_kernel(kern, grid, dim3(warps * 32), smem, st, xyz, xB, xN, xC, new_xyz, qB, qN, qC, N, S, radius2, nsample,
                       static_cast<const unsigned char*>(grid_ws), grid_cloud_bytes(N), threshold, bm_words, done, out_idx);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("pn_ball_query_grid_f32: launch failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        return PN_OK;
    };
    if (N <= 32768) return launch(ball_query_grid_kernel<8>, 8);      // bitmaps: <= 4 KB per warp
    if (N <= 262144) return launch(ball_query_grid_kernel<4>, 4);     // <= 32 KB per warp
    return launch(ball_query_grid_kernel<1>, 1);                      // <= 128 KB
}

PN_EXPORT int pn_ball_query_stream_f32(const float* xyz, int64_t xB, int64_t xN, int64_t xC, const uint64_t* progress, int B, int N,
                                       int S, int s_end, float radius2, int nsample, const void* grid_ws, size_t grid_bytes,
                                       int ctas, size_t min_smem_bytes, int32_t* done, int64_t* out_idx, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && progress && grid_ws && done && out_idx, PN_ERR_BAD_ARG, "pn_ball_query_stream_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, PN_ERR_BAD_ARG, "pn_ball_query_stream_f32: sizes must be positive");
    PN_REQUIRE(grid_bytes >= pn_ball_grid_bytes(B, N), PN_ERR_BAD_ARG, "pn_ball_query_stream_f32: grid buffer too small");
    PN_REQUIRE(N <= 32768, PN_ERR_UNSUPPORTED, "pn_ball_query_stream_f32: N=%d exceeds 32768", N);
    PN_REQUIRE(ctas >= B && ctas % B == 0, PN_ERR_BAD_ARG, "pn_ball_query_stream_f32: ctas=%d must be a positive multiple of B=%d",
               ctas, B);
    const int bm_words = (int)ceil_div(ceil_div(N, 32), 32) * 32;
    size_t smem = (size_t)8 * bm_words * 4;
    if (smem < (size_t)kGqTile * 16) smem = (size_t)kGqTile * 16;
    if (smem < min_smem_bytes) smem = min_smem_bytes;   // large enough that a CTA cannot share an SM with the sampling kernel
    PN_REQUIRE(smem <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_ball_query_stream_f32: %zu bytes of shared memory requested", smem);
    auto kern = ball_query_stream_kernel<8>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("pn_ball_query_stream_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    PN_REQUIRE(s_end >= 0 && s_end <= S, PN_ERR_BAD_ARG, "pn_ball_query_stream_f32: s_end=%d outside [0, S=%d]", s_end, S);
    kern<<<ctas, 256, smem, (cudaStream_t)stream>>>(xyz, xB, xN, xC, reinterpret_cast<const unsigned long long*>(progress), B, N, S,
                                                    s_end, radius2, nsample, static_cast<const unsigned char*>(grid_ws), grid_cloud_bytes(N),
                                                    gq_auto_threshold(N), bm_words, 3000000LL, done, out_idx);
    return finish_launch("pn_ball_query_stream_f32");
}
