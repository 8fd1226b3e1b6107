// fps.cu -- farthest-point sampling (reference: model/pointnet_util.py:63-84).
//
// One thread-block CLUSTER per cloud.  The cloud is split into contiguous chunks, one per CTA; every
// thread keeps PTS points (x, y, z, running min-distance) in registers for the whole kernel, so an
// iteration touches no global memory at all.  Per iteration:
//   1. each thread updates its PTS running distances against the current centroid and tracks its
//      best (distance, index);
//   2. warp arg-max with two REDUX instructions (max over the distance bits -- distances are >= 0, so
//      unsigned order equals float order -- then min over the indices that attain it);
//   3. one __syncthreads, then every warp redundantly reduces the per-warp winners (no second barrier);
//   4. cluster exchange: the CTA's winner (packed key + its coordinates) is stored into every peer's
//      shared memory through DSMEM, one barrier.cluster, then all threads pick the cluster winner.
// Slots are double-buffered by iteration parity so that one barrier per level and iteration is enough.
//
// Bit-exactness: distance = ((dx*dx + dy*dy) + dz*dz) with explicit _rn intrinsics (no FMA
// contraction), running min, ties resolved to the lowest point index -- as torch.max does on CPU.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pn {

constexpr int kMaxCluster = 16;
constexpr unsigned kNoIndex = 0xFFFFFFFFu;

struct __align__(32) FpsMsg {
    unsigned long long key;  // (distance bits << 32) | ~index : max key = max distance, then min index
    float x, y, z;
    float pad[3];
};

constexpr int kFpsSmemHeader = 2 * kMaxCluster * (int)sizeof(FpsMsg) + 2 * 32 * (int)sizeof(uint2);

template <int THREADS, int PTS, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int64_t sB, int64_t sN, int64_t sC, int N, int npoint,
           const int64_t* __restrict__ start, int64_t* __restrict__ out, int chunk) {
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    FpsMsg* cl_slots = reinterpret_cast<FpsMsg*>(smem_raw);                                   // [2][16]
    uint2* warp_slots = reinterpret_cast<uint2*>(smem_raw + 2 * kMaxCluster * sizeof(FpsMsg));  // [2][32]
    float* sx = reinterpret_cast<float*>(smem_raw + kFpsSmemHeader);
    float* sy = sx + chunk;
    float* sz = sy + chunk;

    unsigned CL = 1, rank = 0;
    if constexpr (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        CL = cluster.num_blocks();
        rank = cluster.block_rank();
    }
    const int b = blockIdx.x / CL;
    const float* __restrict__ p = xyz + (int64_t)b * sB;
    const int base = rank * chunk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float x[PTS], y[PTS], z[PTS], d[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) {
        const int li = tid + k * THREADS;
        const int j = base + li;
        const bool ok = li < chunk && j < N;
        x[k] = ok ? p[(int64_t)j * sN] : 0.0f;
        y[k] = ok ? p[(int64_t)j * sN + sC] : 0.0f;
        z[k] = ok ? p[(int64_t)j * sN + 2 * sC] : 0.0f;
        d[k] = ok ? 1e10f : -2.0f;  // -2 marks an empty slot: never selected, never updated
        if (li < chunk) {
            sx[li] = x[k];
            sy[li] = y[k];
            sz[li] = z[k];
        }
    }
    int far = (int)start[b];
    float cx = p[(int64_t)far * sN], cy = p[(int64_t)far * sN + sC], cz = p[(int64_t)far * sN + 2 * sC];
    if constexpr (CLUSTER) cg::this_cluster().sync();  // peers must be resident before DSMEM stores
    else __syncthreads();

    int64_t* __restrict__ o = out + (int64_t)b * npoint;
    for (int i = 0; i < npoint; ++i) {
        if (rank == 0 && tid == 0) o[i] = far;
        float bv = -1.0f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const float dd = sqdist_diff(x[k], y[k], z[k], cx, cy, cz);
            d[k] = fminf(d[k], dd);
            if (d[k] > bv) {  // strict: the lowest index wins inside a thread (k ascending = index ascending)
                bv = d[k];
                bk = k;
            }
        }
        const unsigned vb = __float_as_uint(fmaxf(bv, 0.0f));
        const unsigned gi = bv < 0.0f ? kNoIndex : (unsigned)(base + tid + bk * THREADS);
        const unsigned wm = __reduce_max_sync(0xffffffffu, vb);
        const unsigned wi = __reduce_min_sync(0xffffffffu, vb == wm ? gi : kNoIndex);
        const int par = i & 1;
        unsigned bm, bi;
        if constexpr (NW > 1) {
            if (lane == 0) warp_slots[par * 32 + warp] = make_uint2(wm, wi);
            __syncthreads();
            const uint2 s = lane < NW ? warp_slots[par * 32 + lane] : make_uint2(0u, kNoIndex);
            bm = __reduce_max_sync(0xffffffffu, s.x);
            bi = __reduce_min_sync(0xffffffffu, s.x == bm ? s.y : kNoIndex);
        } else {
            bm = wm;
            bi = wi;
        }
        if constexpr (!CLUSTER) {
            far = (int)bi;
            cx = sx[far];
            cy = sy[far];
            cz = sz[far];
        } else {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0 && lane < (int)CL) {
                FpsMsg m;
                m.key = ((unsigned long long)bm << 32) | (unsigned long long)(~bi);
                const int li = bi == kNoIndex ? 0 : (int)bi - base;
                m.x = sx[li];
                m.y = sy[li];
                m.z = sz[li];
                m.pad[0] = m.pad[1] = m.pad[2] = 0.0f;
                FpsMsg* dst = cluster.map_shared_rank(&cl_slots[par * kMaxCluster + rank], lane);
                *dst = m;
            }
            cluster.sync();
            unsigned long long best = 0ull;
            int bc = 0;
            for (int c = 0; c < (int)CL; ++c) {
                const unsigned long long k = cl_slots[par * kMaxCluster + c].key;
                if (k > best) {
                    best = k;
                    bc = c;
                }
            }
            far = (int)(~(unsigned)(best & 0xFFFFFFFFull));
            const FpsMsg& w = cl_slots[par * kMaxCluster + bc];
            cx = w.x;
            cy = w.y;
            cz = w.z;
        }
    }
}

static int g_force_cluster = 0;
static int g_force_threads = 0;

template <int THREADS, int PTS>
static int launch_fps(bool use_cluster, int CL, const float* xyz, int64_t sB, int64_t sN, int64_t sC, int B, int N,
                      int npoint, const int64_t* start, int64_t* out, int chunk, cudaStream_t stream) {
    const size_t smem = (size_t)kFpsSmemHeader + (size_t)chunk * 3 * sizeof(float);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    cudaError_t e;
    if (use_cluster) {
        auto kern = fps_kernel<THREADS, PTS, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && CL > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) {
            set_error("pn_fps_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, xyz, sB, sN, sC, N, npoint, start, out, chunk);
    } else {
        auto kern = fps_kernel<THREADS, PTS, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("pn_fps_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        e = cudaLaunchKernelEx(&cfg, kern, xyz, sB, sN, sC, N, npoint, start, out, chunk);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("pn_fps_f32: launch failed (cluster=%d threads=%d pts=%d smem=%zu): %s", CL, THREADS, PTS, smem,
                  cudaGetErrorString(e));
        return (int)e;
    }
    return PN_OK;
}

template <int THREADS>
static int dispatch_pts(int pts, bool use_cluster, int CL, const float* xyz, int64_t sB, int64_t sN, int64_t sC, int B,
                        int N, int npoint, const int64_t* start, int64_t* out, int chunk, cudaStream_t stream) {
#define PN_FPS_CASE(P)                                                                                           \
    if (pts <= P)                                                                                                \
    return launch_fps<THREADS, P>(use_cluster, CL, xyz, sB, sN, sC, B, N, npoint, start, out, chunk, stream)
    PN_FPS_CASE(1);
    PN_FPS_CASE(2);
    PN_FPS_CASE(4);
    PN_FPS_CASE(8);
    if constexpr (THREADS <= 512) {
        PN_FPS_CASE(16);
    }
#undef PN_FPS_CASE
    set_error("pn_fps_f32: %d points per thread at %d threads exceeds the register-resident limit", pts, THREADS);
    return PN_ERR_UNSUPPORTED;
}

}  // namespace pn

PN_EXPORT int pn_fps_set_config(int cluster_size, int threads) {
    const bool cl_ok = cluster_size == 0 || cluster_size == 1 || cluster_size == 2 || cluster_size == 4 ||
                       cluster_size == 8 || cluster_size == 16;
    const bool th_ok = threads == 0 || threads == 64 || threads == 128 || threads == 256 || threads == 512 ||
                       threads == 1024;
    PN_REQUIRE(cl_ok && th_ok, PN_ERR_BAD_ARG, "pn_fps_set_config: cluster_size in {0,1,2,4,8,16}, threads in {0,64..1024}");
    pn::g_force_cluster = cluster_size;
    pn::g_force_threads = threads;
    return PN_OK;
}

PN_EXPORT int pn_fps_f32(const float* xyz, int64_t sB, int64_t sN, int64_t sC, int B, int N, int npoint,
                         const int64_t* start_idx, int64_t* out_idx, pn_stream_t stream) {
    using namespace pn;
    PN_REQUIRE(xyz && start_idx && out_idx, PN_ERR_BAD_ARG, "pn_fps_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && npoint > 0, PN_ERR_BAD_ARG, "pn_fps_f32: B, N, npoint must be positive (got %d, %d, %d)", B,
               N, npoint);
    int CL = g_force_cluster;
    if (CL == 0) {
        if (N <= 3072) CL = 1;
        else if (N <= 6144) CL = 2;
        else if (N <= 12288) CL = 4;
        else if (N <= 24576) CL = 8;
        else CL = 16;
    }
    int threads = g_force_threads;
    if (threads == 0) threads = N <= 128 ? 64 : N <= 512 ? 128 : N <= 2048 ? 256 : 512;
    const int chunk = (int)ceil_div(N, CL);
    const int pts = (int)ceil_div(chunk, threads);
    PN_REQUIRE((size_t)kFpsSmemHeader + (size_t)chunk * 12 <= 227 * 1024, PN_ERR_UNSUPPORTED,
               "pn_fps_f32: N=%d needs %d points per CTA at cluster size %d; shared memory holds 19200", N, chunk, CL);
    cudaStream_t st = (cudaStream_t)stream;
    const bool uc = CL > 1;
    switch (threads) {
        case 64: return dispatch_pts<64>(pts, uc, CL, xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, chunk, st);
        case 128: return dispatch_pts<128>(pts, uc, CL, xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, chunk, st);
        case 256: return dispatch_pts<256>(pts, uc, CL, xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, chunk, st);
        case 512: return dispatch_pts<512>(pts, uc, CL, xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, chunk, st);
        default: return dispatch_pts<1024>(pts, uc, CL, xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, chunk, st);
    }
}
