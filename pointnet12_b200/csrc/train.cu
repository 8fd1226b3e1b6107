// train.cu -- kernels of the TRAINING step of PointNet2SemSeg and of the
// evaluation metrics (row f-2).  Reference: pcdseg.py:157-186 (forward in train mode, CrossEntropyLoss on the
// log-probabilities, backward, Adam) and pcdseg.py:58-97 (argmax, per-class I/U, accuracy).
//
// The training forward differs from inference in one way that matters for the kernels: BatchNorm normalises with the
// statistics of the CURRENT batch (over every row of the layer: B*S*K rows for a set-abstraction level, B*N for a
// feature-propagation level), so a layer is  GEMM -> column statistics -> normalise+ReLU  with a grid-wide reduction
// in the middle, and the activations of every layer are kept for the backward pass.  Everything here works on the
// point-major row matrices [rows, C] of the inference path.
//
//   forward   pn_linear_f32 (linear.cu)           y = x W^T + b
//             pn_bn_stats_f32                     column sums / sums of squares in fp64
//             pn_bn_finalize_f32                  mean, 1/sqrt(var+eps), scale/shift, running statistics (momentum)
//             pn_bn_act_f32 / pn_bn_act_max_f32   z = relu(y*scale+shift) / the same + max over nsample + arg-max
//             pn_dropout_f32                      Philox mask (or a given mask), scale 1/(1-p)
//             pn_log_softmax_f32 (group.cu), pn_cross_entropy_f32
//   backward  pn_log_softmax_bwd_f32, pn_dropout_f32 (mask reuse)
//             pn_bn_bwd_stats_f32                 sum(g), sum(g*xhat) with g = dz * [z > 0] (pooled: routed by arg-max)
//             pn_bn_bwd_apply_f32                 dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); dgamma, dbeta
//             pn_grad_weight_f32                  dW += dy^T x, db += sum(dy)     (split over the rows, fp32 atomics)
//             pn_linear_f32 on W^T (pn_transpose_f32)   dx = dy W
//             pn_group_bwd_f32, pn_three_interpolate_bwd_f32    scatter-add through the gathers
//   step      pn_adam_f32                         torch.optim.Adam (L2 weight decay) on one flat parameter buffer
//   metrics   pn_seg_metrics_f32 + pn_seg_metrics_accumulate     pcdseg.py:72-83 without host round trips
//
// All of these are HBM-bound streaming / reduction kernels except pn_grad_weight_f32 (CUDA-core SGEMM with a long K).
#include <math_constants.h>

#include "common.cuh"

namespace pn {

static inline int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ------------------------------------------------------------------------------------------------
// Column statistics.  Block = 32 channels x 16 row lanes; a warp reads 128 contiguous bytes of a row.
// MODE 0: s1 = sum y, s2 = sum y^2.
// MODE 1: g = dz * [y*scale+shift > 0] (relu) ; xhat = (y-mean)*invstd ; s1 = sum g, s2 = sum g*xhat.
// MODE 2: as 1 with dz routed from the pooled gradient: g = dpool[row/K] if argmax[row/K] == row%K.
constexpr int ST_LANES = 16;

template <int MODE>
__global__ void __launch_bounds__(32 * ST_LANES)
col_stats_kernel(const float* __restrict__ y, int64_t ldy, int64_t rows, int C, int64_t rows_per_block,
                 const float* __restrict__ dz, int64_t lddz, const int32_t* __restrict__ argmax, int K,
                 const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                 const float* __restrict__ invstd, int relu, double* __restrict__ s1, double* __restrict__ s2) {
    __shared__ double r1[ST_LANES][33], r2[ST_LANES][33];
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1e = min(rows, r0 + rows_per_block);
    double a1 = 0.0, a2 = 0.0;
    if (c < C) {
        float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
        if (MODE != 0) { sc = scale[c]; sh = shift[c]; mu = mean[c]; is = invstd[c]; }
        for (int64_t r = r0 + ry; r < r1e; r += ST_LANES) {
            const float v = y[r * ldy + c];
            if (MODE == 0) {
                const double d = (double)v;
                a1 += d;
                a2 = fma(d, d, a2);
            } else {
                float g;
                if (MODE == 1) {
                    g = dz[r * lddz + c];
                } else {
                    const int64_t grp = r / K;
                    g = (argmax[grp * C + c] == (int)(r - grp * K)) ? dz[grp * lddz + c] : 0.0f;
                }
                if (relu && !(fmaf(v, sc, sh) > 0.0f)) g = 0.0f;
                const float xh = (v - mu) * is;
                a1 += (double)g;
                a2 = fma((double)g, (double)xh, a2);
            }
        }
    }
    r1[ry][lane] = a1;
    r2[ry][lane] = a2;
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < ST_LANES; ++i) { a1 += r1[i][lane]; a2 += r2[i][lane]; }
        atomicAdd(&s1[c], a1);
        atomicAdd(&s2[c], a2);
    }
}

static inline void stats_grid(int64_t rows, int C, dim3* grid, int64_t* rows_per_block) {
    const int cy = (int)ceil_div(C, 32);
    int64_t bx = ceil_div((int64_t)sm_count() * 8, cy);
    int64_t rpb = ceil_div(rows, bx);
    rpb = ceil_div(rpb < 64 ? 64 : rpb, ST_LANES) * ST_LANES;
    *rows_per_block = rpb;
    *grid = dim3((unsigned)ceil_div(rows, rpb), (unsigned)cy, 1);
}

__global__ void bn_finalize_kernel(const double* __restrict__ s1, const double* __restrict__ s2, int64_t n, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   int64_t* __restrict__ num_batches_tracked, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    const double mu = s1[c] / (double)n;
    double var = s2[c] / (double)n - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    mean[c] = (float)mu;
    invstd[c] = is;
    scale[c] = g * is;
    shift[c] = b - (float)mu * (g * is);
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mu;
    if (running_var) {
        const double unbiased = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// z = act(y*scale + shift), elementwise over [rows, C].
__global__ void bn_act_kernel(const float* __restrict__ y, int64_t ldy, int64_t rows, int C,
                              const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                              float* __restrict__ z, int64_t ldz) {
    const int64_t total = rows * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t r = e / C;
        float v = fmaf(y[r * ldy + c], scale[c], shift[c]);
        if (relu) v = fmaxf(v, 0.0f);
        z[r * ldz + c] = v;
    }
}

// pooled[g,c] = max_k act(y[g*K+k,c]*scale+shift), argmax = the first k attaining it (torch.max(dim) on CPU).
__global__ void bn_act_max_kernel(const float* __restrict__ y, int64_t ldy, int64_t groups, int K, int C,
                                  const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                  float* __restrict__ out, int64_t ldo, int32_t* __restrict__ argmax) {
    const int64_t total = groups * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t g = e / C;
        const float sc = scale[c], sh = shift[c];
        const float* __restrict__ r = y + g * K * ldy + c;
        float m = -CUDART_INF_F;
        int am = 0;
        for (int k = 0; k < K; ++k) {
            float v = fmaf(r[(int64_t)k * ldy], sc, sh);
            if (relu) v = fmaxf(v, 0.0f);
            if (v > m) { m = v; am = k; }
        }
        out[g * ldo + c] = m;
        argmax[g * C + c] = am;
    }
}

// Tall variant (K >= 256: the global max over the N points of a cloud in the PointNet nets and group-all levels): one CTA per
// (group, 32-channel slab), 8 row lanes per channel, then a shared-memory reduction that keeps the FIRST maximum.
__global__ void __launch_bounds__(256)
bn_act_max_tall_kernel(const float* __restrict__ y, int64_t ldy, int K, int C, const float* __restrict__ scale,
                       const float* __restrict__ shift, int relu, float* __restrict__ out, int64_t ldo,
                       int32_t* __restrict__ argmax) {
    __shared__ float vmax[8][33];
    __shared__ int vidx[8][33];
    const int64_t g = blockIdx.y;
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float m = -CUDART_INF_F;
    int am = 0;
    if (c < C) {
        const float sc = scale[c], sh = shift[c];
        const float* __restrict__ r = y + g * K * ldy + c;
        for (int k = ry; k < K; k += 8) {
            float v = fmaf(r[(int64_t)k * ldy], sc, sh);
            if (relu) v = fmaxf(v, 0.0f);
            if (v > m) { m = v; am = k; }           // a lane sees its rows in ascending order: first maximum of the lane
        }
    }
    vmax[ry][lane] = m;
    vidx[ry][lane] = am;
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            const float v = vmax[i][lane];
            const int k = vidx[i][lane];
            if (v > m || (v == m && k < am)) { m = v; am = k; }      // ties across lanes: the lowest row index
        }
        out[g * ldo + c] = m;
        argmax[g * C + c] = am;
    }
}

// dy = gamma*invstd * (g - s1/n - xhat*s2/n); block 0 also adds s2 to dgamma and s1 to dbeta.
template <int MODE>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ y, int64_t ldy, int64_t rows, int C,
                                    const float* __restrict__ dz, int64_t lddz, const int32_t* __restrict__ argmax, int K,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                    const double* __restrict__ s1, const double* __restrict__ s2,
                                    float* __restrict__ dy, int64_t lddy, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
    const int64_t total = rows * C;
    const double inv_n = 1.0 / (double)rows;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dgamma) dgamma[c] += (float)s2[c];      // accumulated, like dW / db (a single writer: no atomics)
            if (dbeta) dbeta[c] += (float)s1[c];
        }
    }
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t r = e / C;
        const float v = y[r * ldy + c];
        const float sc = scale[c], sh = shift[c];
        float g;
        if (MODE == 1) {
            g = dz[r * lddz + c];
        } else {
            const int64_t grp = r / K;
            g = (argmax[grp * C + c] == (int)(r - grp * K)) ? dz[grp * lddz + c] : 0.0f;
        }
        if (relu && !(fmaf(v, sc, sh) > 0.0f)) g = 0.0f;
        const float xh = (v - mean[c]) * invstd[c];
        const float m1 = (float)(s1[c] * inv_n), m2 = (float)(s2[c] * inv_n);
        dy[r * lddy + c] = sc * (g - m1 - xh * m2);      // scale = gamma * invstd
    }
}

// ------------------------------------------------------------------------------------------------
// dW[co,ci] += sum_r dy[r,co] * x[r,ci]; db[co] += sum_r dy[r,co].  Both operands are read along their rows
// (channels contiguous), so a K step is BK rows of each: straight 16-byte loads into shared memory, no transpose.
// grid = (co tiles, ci tiles, row splits); fp32 atomics into the pre-zeroed (or accumulating) dW.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
grad_weight_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx, int64_t rows,
                   int64_t rows_per_split, int cout, int cin, float* __restrict__ dw, int64_t lddw,
                   float* __restrict__ db, int vec_ok) {
    constexpr int THREADS = (BM / TM) * (BN / TN);
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int co0 = blockIdx.x * BM, ci0 = blockIdx.y * BN;
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r1 = min(rows, r0 + rows_per_split);
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const bool do_bias = db != nullptr && blockIdx.y == 0 && tx == 0;
    float acc[TM][TN];
    float bsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        bsum[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    }
    constexpr int AV = BM / 4, BV = BN / 4;
    for (int64_t k0 = r0; k0 < r1; k0 += BK) {
        for (int e = tid; e < BK * AV; e += THREADS) {
            const int k = e / AV, cq = (e % AV) * 4;
            const int64_t r = k0 + k;
            const int gc = co0 + cq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
                const float* src = dy + r * lddy + gc;
                if ((vec_ok & 1) && gc + 3 < cout) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (gc < cout) v.x = src[0];
                    if (gc + 1 < cout) v.y = src[1];
                    if (gc + 2 < cout) v.z = src[2];
                    if (gc + 3 < cout) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(&As[k][cq]) = v;
        }
        for (int e = tid; e < BK * BV; e += THREADS) {
            const int k = e / BV, cq = (e % BV) * 4;
            const int64_t r = k0 + k;
            const int gc = ci0 + cq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
                const float* src = x + r * ldx + gc;
                if ((vec_ok & 2) && gc + 3 < cin) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (gc < cin) v.x = src[0];
                    if (gc + 1 < cin) v.y = src[1];
                    if (gc + 2 < cin) v.z = src[2];
                    if (gc + 3 < cin) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(&Bs[k][cq]) = v;
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
            if (do_bias) {
#pragma unroll
                for (int i = 0; i < TM; ++i) bsum[i] += a[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int co = co0 + ty * TM + i;
        if (co >= cout) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ci = ci0 + tx * TN + j;
            if (ci < cin) atomicAdd(&dw[(int64_t)co * lddw + ci], acc[i][j]);
        }
        if (do_bias) atomicAdd(&db[co], bsum[i]);
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int x = blockIdx.x * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int yy = blockIdx.y * 32 + j;
        if (x < cols && yy < rows) tile[j][threadIdx.x] = in[(int64_t)yy * cols + x];
    }
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x;   // column of out = row of in
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int oy = blockIdx.x * 32 + j;
        if (ox < rows && oy < cols) out[(int64_t)oy * rows + ox] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------------
// Scatter-add through the gathers.
// grouped rows (b,s,k) had channels [xyz_rel(3), feat(D)] (or [feat, xyz_rel] in MSG order); only feat carries grad.
__global__ void group_bwd_kernel(const float* __restrict__ dg, int64_t ldg, int col0, int D,
                                 const int64_t* __restrict__ idx, int N, int64_t SK, int64_t total,
                                 float* __restrict__ dfeat) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % D);
        const int64_t row = e / D;
        const int64_t b = row / SK;
        int64_t j = idx[row];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        atomicAdd(&dfeat[(b * N + j) * D + c], dg[row * ldg + col0 + c]);
    }
}

// rows (b,n) of dx = [d points1 (D1) | d interpolated (D2)]
__global__ void three_interpolate_bwd_kernel(const float* __restrict__ dx, int64_t ldx, int D1, int D2,
                                             const int64_t* __restrict__ idx, const float* __restrict__ weight, int N,
                                             int S, int64_t total, float* __restrict__ dp1, float* __restrict__ dp2) {
    const int C = D1 + D2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t row = e / C;
        const float g = dx[row * ldx + c];
        if (c < D1) {
            dp1[row * D1 + c] = g;
        } else {
            const int64_t b = row / N;
            const int cc = c - D1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int64_t j = idx[row * 3 + k];
                j = j < 0 ? 0 : (j >= S ? S - 1 : j);
                atomicAdd(&dp2[(b * S + j) * D2 + cc], g * weight[row * 3 + k]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dropout.  Philox4x32-10 keyed by (seed), counter = (element / 4, offset): four uniform words per call.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
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

// y = x * keep / (1-p).  mask_in != NULL: keep = mask_in (bytes); else keep is drawn (P(keep) = 1-p) from the Philox
// stream named by *seed_offset = {seed, offset} (device memory: a CUDA graph replays with fresh values) and written
// to mask_out.  Dense [rows*C] element indexing for the random stream, leading dimensions for the data.
__global__ void dropout_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, float p,
                               const uint64_t* __restrict__ seed_offset, const uint8_t* __restrict__ mask_in,
                               uint8_t* __restrict__ mask_out, float* __restrict__ y, int64_t ldy) {
    const int64_t total = rows * C;
    const float inv_keep = 1.0f / (1.0f - p);
    const int64_t quads = (total + 3) / 4;
    uint2 key = make_uint2(0, 0);
    uint32_t off_lo = 0, off_hi = 0;
    if (!mask_in) {
        const uint64_t seed = seed_offset[0], off = seed_offset[1];
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        off_lo = (uint32_t)off;
        off_hi = (uint32_t)(off >> 32);
    }
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t rnd[4] = {0, 0, 0, 0};
        if (!mask_in) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)((uint64_t)q >> 32), off_lo, off_hi), key);
            rnd[0] = r.x; rnd[1] = r.y; rnd[2] = r.z; rnd[3] = r.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t e = q * 4 + i;
            if (e >= total) break;
            const int c = (int)(e % C);
            const int64_t r = e / C;
            uint8_t keep;
            if (mask_in) {
                keep = mask_in[e];
            } else {
                // uniform in [0,1) from the top 24 bits; keep when u >= p
                keep = ((float)(rnd[i] >> 8) * (1.0f / 16777216.0f)) >= p ? 1 : 0;
                if (mask_out) mask_out[e] = keep;
            }
            y[r * ldy + c] = keep ? x[r * ldx + c] * inv_keep : 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// nn.CrossEntropyLoss()(x.transpose(2,1), target) of pcdseg.py:177-178 on rows x [rows, C] (the reference feeds the
// network's log-probabilities, so CrossEntropyLoss normalises them once more): loss = mean_i (lse(x_i) - x_i[t_i]);
// dx[i,c] = (softmax(x_i)[c] - [c == t_i]) * grad_scale / rows.  One row per thread; loss summed in fp64.
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ target, int64_t rows, int C,
                     double* __restrict__ loss_sum, float* __restrict__ dx, int64_t lddx, float grad_scale) {
    __shared__ double red[8];
    double local = 0.0;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const float* __restrict__ xr = x + r * ldx;
        float m = -CUDART_INF_F;
        for (int c = 0; c < C; ++c) m = fmaxf(m, xr[c]);
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += expf(xr[c] - m);
        const float lse = m + logf(s);
        const int64_t t = target[r];
        if (t < 0 || t >= C) {
            // nn.CrossEntropyLoss raises on a label outside [0, C); an asynchronous kernel cannot, so the loss and this
            // row's gradient are poisoned with NaN instead of being computed against a clamped label
            local += (double)CUDART_NAN_F;
            if (dx)
                for (int c = 0; c < C; ++c) dx[r * lddx + c] = CUDART_NAN_F;
            continue;
        }
        local += (double)(lse - xr[t]);
        if (dx) {
            const float k = grad_scale / (float)rows;
            for (int c = 0; c < C; ++c) dx[r * lddx + c] = (expf(xr[c] - lse) - (c == (int)t ? 1.0f : 0.0f)) * k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        atomicAdd(loss_sum, t);
    }
}

__global__ void loss_finalize_kernel(const double* __restrict__ loss_sum, int64_t rows, float* __restrict__ loss) {
    *loss = (float)(*loss_sum / (double)rows);
}

// log_softmax backward: dx = dy - exp(y) * sum_c dy   (y = the log-probabilities)
__global__ void log_softmax_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y,
                                       int64_t ldy, int64_t rows, int C, float* __restrict__ dx, int64_t lddx) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += dy[r * lddy + c];
        for (int c = 0; c < C; ++c) dx[r * lddx + c] = dy[r * lddy + c] - expf(y[r * ldy + c]) * s;
    }
}

// ------------------------------------------------------------------------------------------------
// torch.optim.Adam (amsgrad off, L2 weight decay added to the gradient) on a flat buffer.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_over_bc1, float beta1, float beta2, float eps,
                            float weight_decay, float inv_sqrt_bc2, float grad_scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);                // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.0f - beta2, gi * gi, v[i] * beta2);          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - lr_over_bc1 * (mi / denom);
    }
}

// The same update with the learning rate and the step count in DEVICE memory, so that a captured CUDA graph replays
// with the current values: *step is incremented by a one-thread kernel first, then every thread derives the bias
// corrections from it (double precision, like the host variant).
__global__ void adam_step_increment_kernel(int64_t* step) { *step += 1; }

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ lr_ptr,
                                const int64_t* __restrict__ step_ptr, float beta1, float beta2, float eps,
                                float weight_decay, float grad_scale) {
    const double step = (double)*step_ptr;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float lr_over_bc1 = (float)((double)*lr_ptr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);
        const float vi = fmaf(1.0f - beta2, gi * gi, v[i] * beta2);
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - lr_over_bc1 * (mi / denom);
    }
}

// ------------------------------------------------------------------------------------------------
// Evaluation metrics of pcdseg.py:72-83: arg-max label per point, per-class intersection / prediction / target
// counts and the number of correct points -- one pass over the log-probabilities, block-local histograms.
// counts layout (int64): [0,C) intersection, [C,2C) predicted, [2C,3C) target, [3C] correct.
constexpr int METRIC_MAX_CLASSES = 64;
__global__ void __launch_bounds__(256)
seg_metrics_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ target, int64_t rows, int C,
                   int64_t* __restrict__ pred_out, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int h[3 * METRIC_MAX_CLASSES + 1];
    for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const float* __restrict__ xr = x + r * ldx;
        float m = xr[0];
        int am = 0;
        for (int c = 1; c < C; ++c) {
            const float v = xr[c];
            if (v > m) { m = v; am = c; }        // first maximum, like torch.argmax on CPU
        }
        if (pred_out) pred_out[r] = am;
        const int64_t t = target[r];
        atomicAdd(&h[C + am], 1u);
        if (t >= 0 && t < C) atomicAdd(&h[2 * C + (int)t], 1u);
        if (t == am) {
            atomicAdd(&h[am], 1u);
            atomicAdd(&h[3 * C], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x)
        if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

// One batch folded into the running totals exactly as the reference's Python loop does it (pcdseg.py:75-83):
// iou = 1 if U == 0 else I/U (double division, rounded to fp32 when added to the fp32 array), count += 1,
// accuracy list entry = correct / points (double).  state: ious fp32[C], count u32[C], acc_sum f64, batches i64.
__global__ void seg_metrics_accumulate_kernel(const unsigned long long* __restrict__ counts, int C, int64_t points,
                                              float* __restrict__ ious, unsigned int* __restrict__ count,
                                              double* __restrict__ acc_sum, int64_t* __restrict__ batches) {
    const int c = threadIdx.x;
    if (c < C) {
        const unsigned long long I = counts[c], U = counts[C + c] + counts[2 * C + c] - I;
        const double iou = U == 0 ? 1.0 : (double)I / (double)U;
        ious[c] = ious[c] + (float)iou;
        count[c] += 1;
    }
    if (c == 0) {
        *acc_sum += (double)counts[3 * C] / (double)points;
        *batches += 1;
    }
}

static inline unsigned ew_blocks(int64_t total, int threads) {
    const int64_t want = ceil_div(total, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace pn

using namespace pn;

PN_EXPORT int pn_bn_stats_f32(const float* y, int64_t ldy, int64_t rows, int C, double* sum, double* sumsq,
                              pn_stream_t stream) {
    PN_REQUIRE(y && sum && sumsq, PN_ERR_BAD_ARG, "pn_bn_stats_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldy >= C, PN_ERR_BAD_ARG, "pn_bn_stats_f32: bad shape");
    dim3 grid;
    int64_t rpb;
    stats_grid(rows, C, &grid, &rpb);
    col_stats_kernel<0><<<grid, 32 * ST_LANES, 0, (cudaStream_t)stream>>>(y, ldy, rows, C, rpb, nullptr, 0, nullptr, 1, nullptr,
                                                                          nullptr, nullptr, nullptr, 0, sum, sumsq);
    return finish_launch("pn_bn_stats_f32");
}

PN_EXPORT int pn_bn_finalize_f32(const double* sum, const double* sumsq, int64_t n, int C, const float* gamma,
                                 const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                 int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd,
                                 pn_stream_t stream) {
    PN_REQUIRE(sum && sumsq && scale && shift && mean && invstd, PN_ERR_BAD_ARG, "pn_bn_finalize_f32: null pointer");
    PN_REQUIRE(n > 0 && C > 0, PN_ERR_BAD_ARG, "pn_bn_finalize_f32: bad shape");
    bn_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
        sum, sumsq, n, C, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, scale, shift, mean, invstd);
    return finish_launch("pn_bn_finalize_f32");
}

PN_EXPORT int pn_bn_act_f32(const float* y, int64_t ldy, int64_t rows, int C, const float* scale, const float* shift,
                            int relu, float* z, int64_t ldz, pn_stream_t stream) {
    PN_REQUIRE(y && scale && shift && z, PN_ERR_BAD_ARG, "pn_bn_act_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldy >= C && ldz >= C, PN_ERR_BAD_ARG, "pn_bn_act_f32: bad shape");
    bn_act_kernel<<<ew_blocks(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, rows, C, scale, shift, relu, z, ldz);
    return finish_launch("pn_bn_act_f32");
}

PN_EXPORT int pn_bn_act_max_f32(const float* y, int64_t ldy, int64_t groups, int K, int C, const float* scale,
                                const float* shift, int relu, float* out, int64_t ldo, int32_t* argmax,
                                pn_stream_t stream) {
    PN_REQUIRE(y && scale && shift && out && argmax, PN_ERR_BAD_ARG, "pn_bn_act_max_f32: null pointer");
    PN_REQUIRE(groups > 0 && K > 0 && C > 0 && ldy >= C && ldo >= C, PN_ERR_BAD_ARG, "pn_bn_act_max_f32: bad shape");
    if (K >= 256 && groups <= 65535) {
        bn_act_max_tall_kernel<<<dim3((unsigned)ceil_div(C, 32), (unsigned)groups), 256, 0, (cudaStream_t)stream>>>(y, ldy, K, C, scale, shift,
                                                                                                              relu, out, ldo, argmax);
    } else {
        bn_act_max_kernel<<<ew_blocks(groups * C, 128), 128, 0, (cudaStream_t)stream>>>(y, ldy, groups, K, C, scale, shift, relu, out,
                                                                                       ldo, argmax);
    }
    return finish_launch("pn_bn_act_max_f32");
}

PN_EXPORT int pn_bn_bwd_stats_f32(const float* y, int64_t ldy, int64_t rows, int C, const float* dz, int64_t lddz,
                                  const int32_t* argmax, int K, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, int relu, double* s1, double* s2,
                                  pn_stream_t stream) {
    PN_REQUIRE(y && dz && scale && shift && mean && invstd && s1 && s2, PN_ERR_BAD_ARG, "pn_bn_bwd_stats_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldy >= C && lddz >= C, PN_ERR_BAD_ARG, "pn_bn_bwd_stats_f32: bad shape");
    PN_REQUIRE(!argmax || (K > 0 && rows % K == 0), PN_ERR_BAD_ARG, "pn_bn_bwd_stats_f32: rows must be a multiple of K");
    dim3 grid;
    int64_t rpb;
    stats_grid(rows, C, &grid, &rpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (argmax)
        col_stats_kernel<2><<<grid, 32 * ST_LANES, 0, st>>>(y, ldy, rows, C, rpb, dz, lddz, argmax, K, scale, shift, mean, invstd,
                                                            relu, s1, s2);
    else
        col_stats_kernel<1><<<grid, 32 * ST_LANES, 0, st>>>(y, ldy, rows, C, rpb, dz, lddz, nullptr, 1, scale, shift, mean, invstd,
                                                            relu, s1, s2);
    return finish_launch("pn_bn_bwd_stats_f32");
}

PN_EXPORT int pn_bn_bwd_apply_f32(const float* y, int64_t ldy, int64_t rows, int C, const float* dz, int64_t lddz,
                                  const int32_t* argmax, int K, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, int relu, const double* s1, const double* s2,
                                  float* dy, int64_t lddy, float* dgamma, float* dbeta, pn_stream_t stream) {
    PN_REQUIRE(y && dz && scale && shift && mean && invstd && s1 && s2 && dy, PN_ERR_BAD_ARG, "pn_bn_bwd_apply_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldy >= C && lddz >= C && lddy >= C, PN_ERR_BAD_ARG, "pn_bn_bwd_apply_f32: bad shape");
    PN_REQUIRE(!argmax || (K > 0 && rows % K == 0), PN_ERR_BAD_ARG, "pn_bn_bwd_apply_f32: rows must be a multiple of K");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = ew_blocks(rows * C, 256);
    if (argmax)
        bn_bwd_apply_kernel<2><<<blocks, 256, 0, st>>>(y, ldy, rows, C, dz, lddz, argmax, K, scale, shift, mean, invstd, relu, s1, s2,
                                                       dy, lddy, dgamma, dbeta);
    else
        bn_bwd_apply_kernel<1><<<blocks, 256, 0, st>>>(y, ldy, rows, C, dz, lddz, nullptr, 1, scale, shift, mean, invstd, relu, s1,
                                                       s2, dy, lddy, dgamma, dbeta);
    return finish_launch("pn_bn_bwd_apply_f32");
}

PN_EXPORT int pn_grad_weight_f32(const float* dy, int64_t lddy, const float* x, int64_t ldx, int64_t rows, int cout,
                                 int cin, float* dw, int64_t lddw, float* db, pn_stream_t stream) {
    PN_REQUIRE(dy && x && dw, PN_ERR_BAD_ARG, "pn_grad_weight_f32: null pointer");
    PN_REQUIRE(rows > 0 && cout > 0 && cin > 0 && lddy >= cout && ldx >= cin && lddw >= cin, PN_ERR_BAD_ARG,
               "pn_grad_weight_f32: bad shape");
    int vec_ok = 0;
    if (((uintptr_t)dy % 16 == 0) && (lddy % 4 == 0)) vec_ok |= 1;
    if (((uintptr_t)x % 16 == 0) && (ldx % 4 == 0)) vec_ok |= 2;
    // (A 128 x 128 tile with 8 x 8 register tiles measured SLOWER on B200 -- 73 vs 49 us for 128 x 128 over 64 k rows: two
    // resident CTAs cannot hide the global-load latency of a BK = 8 step -- so the 64 x 64 tile is used throughout; the
    // fast path for wide layers is the tensor-core kernel pn_grad_weight_bf16x3.)
    const bool wide = false;
    const int BM = wide ? 128 : 64, BN = BM, BK = wide ? 8 : 16;
    const int64_t tiles = ceil_div(cout, BM) * ceil_div(cin, BN);
    int64_t splits = ceil_div((int64_t)sm_count() * (wide ? 2 : 4), tiles);
    int64_t rps = ceil_div(ceil_div(rows, splits), BK) * BK;
    if (rps < 64) rps = 64;
    splits = ceil_div(rows, rps);
    PN_REQUIRE(splits <= 65535, PN_ERR_UNSUPPORTED, "pn_grad_weight_f32: too many row splits");
    dim3 grid((unsigned)ceil_div(cout, BM), (unsigned)ceil_div(cin, BN), (unsigned)splits);
    if (wide)
        grad_weight_kernel<128, 128, 8, 8, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, lddy, x, ldx, rows, rps, cout, cin, dw, lddw,
                                                                                     db, vec_ok);
    else
        grad_weight_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, lddy, x, ldx, rows, rps, cout, cin, dw, lddw,
                                                                                    db, vec_ok);
    return finish_launch("pn_grad_weight_f32");
}

PN_EXPORT int pn_transpose_f32(const float* in, int rows, int cols, float* out, pn_stream_t stream) {
    PN_REQUIRE(in && out && rows > 0 && cols > 0, PN_ERR_BAD_ARG, "pn_transpose_f32: bad arguments");
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, rows, cols, out);
    return finish_launch("pn_transpose_f32");
}

PN_EXPORT int pn_group_bwd_f32(const float* dgrouped, int64_t ldg, int col0, int D, const int64_t* idx, int B, int N,
                               int S, int K, float* dfeat, pn_stream_t stream) {
    PN_REQUIRE(dgrouped && idx && dfeat, PN_ERR_BAD_ARG, "pn_group_bwd_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D > 0 && col0 >= 0 && ldg >= col0 + D, PN_ERR_BAD_ARG,
               "pn_group_bwd_f32: bad shape");
    const int64_t total = (int64_t)B * S * K * D;
    group_bwd_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(dgrouped, ldg, col0, D, idx, N, (int64_t)S * K, total,
                                                                             dfeat);
    return finish_launch("pn_group_bwd_f32");
}

PN_EXPORT int pn_three_interpolate_bwd_f32(const float* dx, int64_t ldx, int D1, int D2, const int64_t* idx,
                                           const float* weight, int B, int N, int S, float* dpoints1, float* dpoints2,
                                           pn_stream_t stream) {
    PN_REQUIRE(dx && idx && weight && dpoints2 && (D1 == 0 || dpoints1), PN_ERR_BAD_ARG, "pn_three_interpolate_bwd_f32: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && S > 0 && D1 >= 0 && D2 > 0 && ldx >= D1 + D2, PN_ERR_BAD_ARG,
               "pn_three_interpolate_bwd_f32: bad shape");
    const int64_t total = (int64_t)B * N * (D1 + D2);
    three_interpolate_bwd_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(dx, ldx, D1, D2, idx, weight, N, S, total,
                                                                                         dpoints1, dpoints2);
    return finish_launch("pn_three_interpolate_bwd_f32");
}

PN_EXPORT int pn_dropout_f32(const float* x, int64_t ldx, int64_t rows, int C, float p, const uint64_t* seed_offset,
                             const uint8_t* mask_in, uint8_t* mask_out, float* y, int64_t ldy, pn_stream_t stream) {
    PN_REQUIRE(x && y && (mask_in || seed_offset), PN_ERR_BAD_ARG, "pn_dropout_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldx >= C && ldy >= C && p >= 0.0f && p < 1.0f, PN_ERR_BAD_ARG, "pn_dropout_f32: bad arguments");
    dropout_kernel<<<ew_blocks((rows * C + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, p, seed_offset, mask_in,
                                                                                        mask_out, y, ldy);
    return finish_launch("pn_dropout_f32");
}

PN_EXPORT int pn_cross_entropy_f32(const float* x, int64_t ldx, const int64_t* target, int64_t rows, int C,
                                   double* loss_sum, float* loss, float* dx, int64_t lddx, float grad_scale,
                                   pn_stream_t stream) {
    PN_REQUIRE(x && target && loss_sum && loss, PN_ERR_BAD_ARG, "pn_cross_entropy_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && ldx >= C && (!dx || lddx >= C), PN_ERR_BAD_ARG, "pn_cross_entropy_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(double), st);
    if (e != cudaSuccess) {
        set_error("pn_cross_entropy_f32: %s", cudaGetErrorString(e));
        return (int)e;
    }
    cross_entropy_kernel<<<ew_blocks(rows, 256), 256, 0, st>>>(x, ldx, target, rows, C, loss_sum, dx, lddx, grad_scale);
    loss_finalize_kernel<<<1, 1, 0, st>>>(loss_sum, rows, loss);
    return finish_launch("pn_cross_entropy_f32");
}

PN_EXPORT int pn_log_softmax_bwd_f32(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t rows, int C,
                                     float* dx, int64_t lddx, pn_stream_t stream) {
    PN_REQUIRE(dy && y && dx, PN_ERR_BAD_ARG, "pn_log_softmax_bwd_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && lddy >= C && ldy >= C && lddx >= C, PN_ERR_BAD_ARG, "pn_log_softmax_bwd_f32: bad shape");
    log_softmax_bwd_kernel<<<ew_blocks(rows, 256), 256, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, rows, C, dx, lddx);
    return finish_launch("pn_log_softmax_bwd_f32");
}

PN_EXPORT int pn_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                          pn_stream_t stream) {
    PN_REQUIRE(param && grad && exp_avg && exp_avg_sq, PN_ERR_BAD_ARG, "pn_adam_f32: null pointer");
    PN_REQUIRE(n > 0 && step > 0, PN_ERR_BAD_ARG, "pn_adam_f32: n and step must be positive");
    // bias corrections in double like torch (python floats), then the per-element arithmetic in fp32
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)((double)lr / bc1),
                                                                    beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2)),
                                                                    grad_scale);
    return finish_launch("pn_adam_f32");
}

PN_EXPORT int pn_adam_dev_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const float* lr, int64_t* step, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, pn_stream_t stream) {
    PN_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr && step, PN_ERR_BAD_ARG, "pn_adam_dev_f32: null pointer");
    PN_REQUIRE(n > 0, PN_ERR_BAD_ARG, "pn_adam_dev_f32: n must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    adam_step_increment_kernel<<<1, 1, 0, st>>>(step);
    adam_dev_kernel<<<ew_blocks(n, 256), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, step, beta1, beta2, eps, weight_decay,
                                                      grad_scale);
    return finish_launch("pn_adam_dev_f32");
}

PN_EXPORT int pn_seg_metrics_f32(const float* logp, int64_t ldx, const int64_t* target, int64_t rows, int C,
                                 int64_t* pred, int64_t* counts, pn_stream_t stream) {
    PN_REQUIRE(logp && target && counts, PN_ERR_BAD_ARG, "pn_seg_metrics_f32: null pointer");
    PN_REQUIRE(rows > 0 && C > 0 && C <= METRIC_MAX_CLASSES && ldx >= C, PN_ERR_BAD_ARG, "pn_seg_metrics_f32: bad shape (C <= 64)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int64_t) * (3 * C + 1), st);
    if (e != cudaSuccess) {
        set_error("pn_seg_metrics_f32: %s", cudaGetErrorString(e));
        return (int)e;
    }
    seg_metrics_kernel<<<ew_blocks(rows, 256), 256, 0, st>>>(logp, ldx, target, rows, C, pred, (unsigned long long*)counts);
    return finish_launch("pn_seg_metrics_f32");
}

PN_EXPORT int pn_seg_metrics_accumulate(const int64_t* counts, int C, int64_t points, float* ious, uint32_t* count,
                                        double* acc_sum, int64_t* batches, pn_stream_t stream) {
    PN_REQUIRE(counts && ious && count && acc_sum && batches, PN_ERR_BAD_ARG, "pn_seg_metrics_accumulate: null pointer");
    PN_REQUIRE(C > 0 && C <= METRIC_MAX_CLASSES && points > 0, PN_ERR_BAD_ARG, "pn_seg_metrics_accumulate: bad shape");
    seg_metrics_accumulate_kernel<<<1, METRIC_MAX_CLASSES, 0, (cudaStream_t)stream>>>((const unsigned long long*)counts, C, points, ious,
                                                                                     count, acc_sum, batches);
    return finish_launch("pn_seg_metrics_accumulate");
}

// ------------------------------------------------------------------------------------------------
// Row f-4 helpers.
// chamfer_batch / chamfer_non_batch (model/chamfer.py:7-53): sum over b, n of min_m ||p1[b,n] - p2[b,m]||_2, divided by B.
// One thread per p1 point, p2 staged through shared memory in tiles; the minimum is taken over squared distances
// (sqrt is monotonic) and the root drawn once per point; block sums in fp64.
namespace pn {
constexpr int CH_TILE = 256, CH_MAXD = 8;
__global__ void __launch_bounds__(CH_TILE)
chamfer_kernel(const float* __restrict__ p1, int64_t aB, int64_t aN, int64_t aC, const float* __restrict__ p2, int64_t bB,
               int64_t bN, int64_t bC, int N, int M, int D, float* __restrict__ per_point, double* __restrict__ total) {
    __shared__ float tile[CH_TILE][CH_MAXD];
    __shared__ double red[CH_TILE / 32];
    const int b = blockIdx.y;
    const int n = blockIdx.x * CH_TILE + threadIdx.x;
    float q[CH_MAXD];
#pragma unroll
    for (int c = 0; c < CH_MAXD; ++c) q[c] = (n < N && c < D) ? p1[b * aB + (int64_t)n * aN + c * aC] : 0.0f;
    float best = CUDART_INF_F;
    for (int m0 = 0; m0 < M; m0 += CH_TILE) {
        const int m = m0 + threadIdx.x;
#pragma unroll
        for (int c = 0; c < CH_MAXD; ++c) tile[threadIdx.x][c] = (m < M && c < D) ? p2[b * bB + (int64_t)m * bN + c * bC] : 0.0f;
        __syncthreads();
        const int lim = min(CH_TILE, M - m0);
        for (int j = 0; j < lim; ++j) {
            float d = 0.0f;
#pragma unroll
            for (int c = 0; c < CH_MAXD; ++c) {
                const float t = q[c] - tile[j][c];
                d = fmaf(t, t, d);
            }
            best = fminf(best, d);
        }
        __syncthreads();
    }
    const float dist = n < N ? sqrtf(best) : 0.0f;
    if (n < N && per_point) per_point[(int64_t)b * N + n] = dist;
    double s = (double)dist;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < CH_TILE / 32; ++i) t += red[i];
        atomicAdd(total, t);
    }
}

// SemKITTI_2_Common.__call__ (data_utils/kitti_utils.py:97-110): common[..., j] = max(logits[..., src0[j]], logits[..., src1[j]])
__global__ void class_merge_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int n_out, const int* __restrict__ src0,
                                   const int* __restrict__ src1, float* __restrict__ y) {
    const int64_t total = rows * n_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % n_out);
        const int64_t r = e / n_out;
        y[e] = fmaxf(x[r * ldx + src0[j]], x[r * ldx + src1[j]]);
    }
}
}  // namespace pn

PN_EXPORT int pn_chamfer_f32(const float* p1, int64_t aB, int64_t aN, int64_t aC, const float* p2, int64_t bB, int64_t bN,
                             int64_t bC, int B, int N, int M, int D, float* per_point, double* total, pn_stream_t stream) {
    PN_REQUIRE(p1 && p2 && total, PN_ERR_BAD_ARG, "pn_chamfer_f32: null pointer");
    PN_REQUIRE(B > 0 && B <= 65535 && N > 0 && M > 0 && D > 0 && D <= CH_MAXD, PN_ERR_BAD_ARG, "pn_chamfer_f32: bad sizes (D <= 8)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(total, 0, sizeof(double), st);
    if (e != cudaSuccess) {
        set_error("pn_chamfer_f32: %s", cudaGetErrorString(e));
        return (int)e;
    }
    dim3 grid((unsigned)ceil_div(N, CH_TILE), (unsigned)B);
    chamfer_kernel<<<grid, CH_TILE, 0, st>>>(p1, aB, aN, aC, p2, bB, bN, bC, N, M, D, per_point, total);
    return finish_launch("pn_chamfer_f32");
}

PN_EXPORT int pn_class_merge_f32(const float* x, int64_t ldx, int64_t rows, int n_in, int n_out, const int* src0,
                                 const int* src1, float* y, pn_stream_t stream) {
    PN_REQUIRE(x && src0 && src1 && y, PN_ERR_BAD_ARG, "pn_class_merge_f32: null pointer");
    PN_REQUIRE(rows > 0 && n_in > 0 && n_out > 0 && ldx >= n_in, PN_ERR_BAD_ARG, "pn_class_merge_f32: bad shape");
    class_merge_kernel<<<ew_blocks(rows * n_out, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, n_out, src0, src1, y);
    return finish_launch("pn_class_merge_f32");
}
