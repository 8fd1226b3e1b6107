// preprocess.cu -- raw SemanticKITTI scans -> network input, on the device.
// Reference: Semantic_KITTI_Utils.get (data_utils/kitti_utils.py:183-227: label = raw & 0xFFFF -> learning_map, drop
// class 0, label - 1, in-view filter), points_basic_filter / hv_in_range / box_in_range (:238-280), and
// SemKITTI_Loader.__getitem__ (data_utils/SemKITTI_Loader.py:91-115: pcd_normalize :23-30, pcd_jitter :17-21,
// np.random.choice(length, npoints, replace=True)), followed by points.transpose(2, 1) in pcdseg.py:167.
//
// The wire format is the dataset's own: float32 x 4 per point (.bin) and uint32 per point (.label), B scans
// concatenated with an offsets table.  Two steps, all HBM-bound streaming:
//   pn_scan_filter_f32   keep flag per point (label map + field-of-view), ORDER-PRESERVING compaction into a list of
//                        kept indices per scan (tile counts -> scan -> scatter; the reference's boolean-mask indexing
//                        keeps the file order, and np.random.choice indexes into that order)
//   pn_scan_sample_f32   resample with replacement, normalise / clip, jitter, write [B, 4, npoints] channel-major and
//                        the int64 labels [B, npoints] -- the tensors PointNet2SemSeg.forward and the loss take.
// Randomness: the reference draws from numpy's global generator, which cannot be reproduced on a GPU; the caller either
// passes the draws (choice indices, jitter noise -- the parity tests do) or a Philox {seed, offset} and the kernel draws.
#include <math_constants.h>

#include "common.cuh"

namespace pn {

constexpr int PP_THREADS = 256, PP_PER_THREAD = 8, PP_TILE = PP_THREADS * PP_PER_THREAD;

struct FovArgs {
    float h_lo, h_hi, v_lo, v_hi;
};

// kitti_utils.py:213-225 + :262-280 for one point.  atan2f / sqrtf in fp32 like numpy's float32 ufuncs.
__device__ __forceinline__ bool scan_keep(const float4 p, unsigned raw, const uint8_t* __restrict__ lut, int lut_size, FovArgs f,
                                          int inview) {
    const unsigned sem = raw & 0xFFFFu;
    const int mapped = sem < (unsigned)lut_size ? lut[sem] : 0;
    if (mapped == 0) return false;
    if (!inview) return true;
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z)));
    const float h = atan2f(p.y, p.x), v = atan2f(p.z, d);
    const bool fov = h > f.h_lo && h < f.h_hi && v < f.v_hi && v > f.v_lo;
    const bool box = p.x > -10000.f && p.x < 10000.f && p.y > -10000.f && p.y < 10000.f && p.z > -10000.f && p.z < 10000.f &&
                     d > -10000.f && d < 10000.f;
    return fov && box;
}

// pass 0: per-tile counts; pass 1: scatter the kept indices at tile_offset + rank inside the tile
template <int PASS>
__global__ void __launch_bounds__(PP_THREADS)
scan_filter_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ raw_label, const int64_t* __restrict__ offsets,
                   const uint8_t* __restrict__ lut, int lut_size, FovArgs fov, int inview, int tiles, int32_t* __restrict__ tile_counts,
                   int32_t* __restrict__ kept) {
    __shared__ int warp_sums[PP_THREADS / 32];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int64_t base = offsets[b];
    const int64_t n = offsets[b + 1] - base;
    const int64_t first = (int64_t)tile * PP_TILE + (int64_t)threadIdx.x * PP_PER_THREAD;
    unsigned flags = 0;
    int cnt = 0;
    if ((int64_t)tile * PP_TILE < n) {
#pragma unroll
        for (int i = 0; i < PP_PER_THREAD; ++i) {
            const int64_t q = first + i;
            if (q < n && scan_keep(points[base + q], raw_label[base + q], lut, lut_size, fov, inview)) {
                flags |= 1u << i;
                ++cnt;
            }
        }
    }
    // block-wide exclusive scan of cnt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int warp_base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PP_THREADS / 32; ++w) {
        if (w < warp) warp_base += warp_sums[w];
        total += warp_sums[w];
    }
    if (PASS == 0) {
        if (threadIdx.x == 0) tile_counts[b * tiles + tile] = total;
    } else {
        int pos = tile_counts[b * tiles + tile] + warp_base + incl - cnt;      // tile_counts now holds exclusive offsets
#pragma unroll
        for (int i = 0; i < PP_PER_THREAD; ++i)
            if (flags & (1u << i)) kept[base + pos++] = (int32_t)(first + i);
    }
}

// exclusive scan of one scan's tile counts (in place), total -> kept_count[b]; one block per scan
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(int32_t* __restrict__ tile_counts, int tiles, int32_t* __restrict__ kept_count) {
    __shared__ int part[1024];
    int32_t* t = tile_counts + (int64_t)blockIdx.x * tiles;
    const int per = (tiles + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(tiles, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += t[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        const int c = t[i];
        t[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) kept_count[blockIdx.x] = part[1023];
}

__device__ __forceinline__ uint4 pp_philox(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

__device__ __forceinline__ float2 pp_box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);       // (0, 1]
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);                // [0, 1)
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// one thread per output point (b, j)
__global__ void scan_sample_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ raw_label,
                                   const int64_t* __restrict__ offsets, int B, const uint8_t* __restrict__ lut, int lut_size,
                                   const int32_t* __restrict__ kept, const int32_t* __restrict__ kept_count, int npoints,
                                   const int64_t* __restrict__ choice, const float* __restrict__ noise, float sigma, float clip,
                                   const uint64_t* __restrict__ seed_offset, float* __restrict__ out, int64_t* __restrict__ labels) {
    const int64_t total = (int64_t)B * npoints;
    uint2 key = make_uint2(0, 0);
    uint32_t off_lo = 0;
    if (seed_offset) {
        key = make_uint2((uint32_t)seed_offset[0], (uint32_t)(seed_offset[0] >> 32));
        off_lo = (uint32_t)seed_offset[1];
    }
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / npoints);
        const int j = (int)(e - (int64_t)b * npoints);
        const int64_t base = offsets[b];
        const int count = kept_count[b];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t lab = 0;
        if (count > 0) {
            int64_t i;
            if (choice) {
                i = choice[e];
                i = i < 0 ? 0 : (i >= count ? count - 1 : i);          // memory safety only
            } else {
                const uint4 r = pp_philox(make_uint4((uint32_t)j, (uint32_t)b, off_lo, 0u), key);
                i = (int64_t)__umulhi(r.x, (uint32_t)count);             // uniform in [0, count)
            }
            const int64_t src = base + kept[base + i];
            const float4 p = points[src];
            // pcd_normalize (SemKITTI_Loader.py:23-30): x/70, y/70, z/3, (r - 0.5)*2, clip to [-1, 1]
            o.x = fminf(fmaxf(__fdiv_rn(p.x, 70.0f), -1.0f), 1.0f);
            o.y = fminf(fmaxf(__fdiv_rn(p.y, 70.0f), -1.0f), 1.0f);
            o.z = fminf(fmaxf(__fdiv_rn(p.z, 3.0f), -1.0f), 1.0f);
            o.w = fminf(fmaxf(__fmul_rn(__fsub_rn(p.w, 0.5f), 2.0f), -1.0f), 1.0f);
            // pcd_jitter (:17-21): clip(sigma * randn, -clip, clip) per KEPT point and channel, drawn before resampling
            if (noise) {
                const float4 nz = *reinterpret_cast<const float4*>(noise + (base + i) * 4);
                o.x = __fadd_rn(nz.x, o.x); o.y = __fadd_rn(nz.y, o.y); o.z = __fadd_rn(nz.z, o.z); o.w = __fadd_rn(nz.w, o.w);
            } else if (seed_offset && sigma > 0.0f) {
                const uint4 r = pp_philox(make_uint4((uint32_t)i, (uint32_t)b, off_lo, 1u), key);
                const float2 g0 = pp_box_muller(r.x, r.y), g1 = pp_box_muller(r.z, r.w);
                o.x += fminf(fmaxf(sigma * g0.x, -clip), clip);
                o.y += fminf(fmaxf(sigma * g0.y, -clip), clip);
                o.z += fminf(fmaxf(sigma * g1.x, -clip), clip);
                o.w += fminf(fmaxf(sigma * g1.y, -clip), clip);
            }
            const unsigned sem = raw_label[src] & 0xFFFFu;
            lab = (int64_t)(sem < (unsigned)lut_size ? lut[sem] : 0) - 1;
        }
        float* ob = out + (int64_t)b * 4 * npoints + j;
        ob[0] = o.x;
        ob[npoints] = o.y;
        ob[2 * (int64_t)npoints] = o.z;
        ob[3 * (int64_t)npoints] = o.w;
        labels[e] = lab;
    }
}

}  // namespace pn

using namespace pn;

PN_EXPORT size_t pn_scan_workspace_bytes(int B, int64_t max_points) {
    if (B <= 0 || max_points <= 0) return 0;
    const int64_t tiles = ceil_div(max_points, PP_TILE);
    return (size_t)(B * tiles) * sizeof(int32_t);
}

PN_EXPORT int pn_scan_filter_f32(const float* points, const uint32_t* raw_label, const int64_t* offsets, int B,
                                 int64_t max_points, const uint8_t* lut, int lut_size, int inview, float h_lo, float h_hi,
                                 float v_lo, float v_hi, int32_t* kept, int32_t* kept_count, void* workspace,
                                 size_t workspace_bytes, pn_stream_t stream) {
    PN_REQUIRE(points && raw_label && offsets && lut && kept && kept_count && workspace, PN_ERR_BAD_ARG, "pn_scan_filter_f32: null pointer");
    PN_REQUIRE(B > 0 && B <= 65535 && max_points > 0 && lut_size > 0, PN_ERR_BAD_ARG, "pn_scan_filter_f32: bad sizes");
    PN_REQUIRE(((uintptr_t)points & 15) == 0, PN_ERR_ALIGNMENT, "pn_scan_filter_f32: points must be 16-byte aligned");
    const int64_t tiles = ceil_div(max_points, PP_TILE);
    PN_REQUIRE(tiles <= (1 << 20), PN_ERR_UNSUPPORTED, "pn_scan_filter_f32: scans longer than 2^31 points are not supported");
    PN_REQUIRE(workspace_bytes >= pn_scan_workspace_bytes(B, max_points), PN_ERR_BAD_ARG, "pn_scan_filter_f32: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const FovArgs fov = {h_lo, h_hi, v_lo, v_hi};
    int32_t* tc = static_cast<int32_t*>(workspace);
    dim3 grid((unsigned)tiles, (unsigned)B);
    const float4* p4 = reinterpret_cast<const float4*>(points);
    scan_filter_kernel<0><<<grid, PP_THREADS, 0, st>>>(p4, raw_label, offsets, lut, lut_size, fov, inview, (int)tiles, tc, kept);
    scan_tile_offsets_kernel<<<B, 1024, 0, st>>>(tc, (int)tiles, kept_count);
    scan_filter_kernel<1><<<grid, PP_THREADS, 0, st>>>(p4, raw_label, offsets, lut, lut_size, fov, inview, (int)tiles, tc, kept);
    return finish_launch("pn_scan_filter_f32");
}

PN_EXPORT int pn_scan_sample_f32(const float* points, const uint32_t* raw_label, const int64_t* offsets, int B,
                                 const uint8_t* lut, int lut_size, const int32_t* kept, const int32_t* kept_count,
                                 int npoints, const int64_t* choice, const float* noise, float sigma, float clip,
                                 const uint64_t* seed_offset, float* out, int64_t* labels, pn_stream_t stream) {
    PN_REQUIRE(points && raw_label && offsets && lut && kept && kept_count && out && labels, PN_ERR_BAD_ARG,
               "pn_scan_sample_f32: null pointer");
    PN_REQUIRE(B > 0 && npoints > 0 && lut_size > 0, PN_ERR_BAD_ARG, "pn_scan_sample_f32: bad sizes");
    PN_REQUIRE(choice || seed_offset, PN_ERR_BAD_ARG, "pn_scan_sample_f32: pass the choice indices or a Philox seed");
    PN_REQUIRE(((uintptr_t)points & 15) == 0 && (!noise || ((uintptr_t)noise & 15) == 0), PN_ERR_ALIGNMENT,
               "pn_scan_sample_f32: points / noise must be 16-byte aligned");
    const int64_t total = (int64_t)B * npoints;
    const unsigned blocks = (unsigned)(ceil_div(total, 256) > 148 * 16 ? 148 * 16 : ceil_div(total, 256));
    scan_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(points), raw_label, offsets, B, lut,
                                                                lut_size, kept, kept_count, npoints, choice, noise, sigma, clip,
                                                                seed_offset, out, labels);
    return finish_launch("pn_scan_sample_f32");
}
