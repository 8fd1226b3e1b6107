/*
 * pn_oracle.c -- CPU restatement of the PointNet / PointNet++ forward primitives of
 * Jiang-Muyun/PointNet12 (model/pointnet_util.py).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product path
 * (pointnet12_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * fixtures under tests/golden/ that oracle/gen_golden.py produced by importing the reference's own
 * Python modules from /root/reference (torch 2.11 CPU) on seeded synthetic clouds.
 *
 * The file is compiled with -ffp-contract=off: every fused multiply-add below is an explicit
 * fmaf(), because index parity with the reference depends on which roundings happen.
 *
 * Layout convention: point-major rows, xyz[b][n][c] addressed through element strides so that the
 * permuted [B,3,N] views the reference passes around (pointnet_util.py:184) need no copy.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * farthest_point_sample  (pointnet_util.py:63-84)
 *   distance = 1e10 (:74); farthest = caller-supplied start (the reference draws it with
 *   torch.randint on the CPU generator, :75); per iteration: record, dist = sum((xyz-c)**2)
 *   (:80) which torch evaluates as ((dx*dx + dy*dy) + dz*dz) with each op rounded to fp32,
 *   running min (:81-82), argmax with the lowest index winning ties (:83).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_fps(const float* xyz, int64_t sB, int64_t sN, int64_t sC, int B, int N, int npoint,
                     const int64_t* start, int64_t* out /* [B,npoint] */) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float* p = xyz + (int64_t)b * sB;
        float* px = (float*)malloc(sizeof(float) * (size_t)N * 4);
        float* py = px + N;
        float* pz = py + N;
        float* dist = pz + N;
        for (int j = 0; j < N; ++j) {
            px[j] = p[j * sN];
            py[j] = p[j * sN + sC];
            pz[j] = p[j * sN + 2 * sC];
            dist[j] = 1e10f;
        }
        int64_t far = start[b];
        for (int i = 0; i < npoint; ++i) {
            out[(int64_t)b * npoint + i] = far;
            const float cx = px[far], cy = py[far], cz = pz[far];
            for (int j = 0; j < N; ++j) {
                const float dx = px[j] - cx, dy = py[j] - cy, dz = pz[j] - cz;
                const float d = (dx * dx + dy * dy) + dz * dz;
                if (d < dist[j]) dist[j] = d;
            }
            float best = -1.0f;
            for (int j = 0; j < N; ++j) best = dist[j] > best ? dist[j] : best;
            int64_t arg = 0;
            for (int j = 0; j < N; ++j) {
                if (dist[j] == best) { arg = j; break; }
            }
            far = arg;
        }
        free(px);
    }
}

/* ------------------------------------------------------------------------------------------
 * square_distance  (pointnet_util.py:19-40)
 *   dist = -2*matmul(src, dst^T) (:37); dist += sum(src**2) (:38); dist += sum(dst**2) (:39).
 *   The K=3 sgemm accumulates x, then y, then z with fused multiply-adds; the squared norms are
 *   plain fp32 ((x*x + y*y) + z*z).  The result may be slightly negative for coincident points.
 * ------------------------------------------------------------------------------------------ */
static inline float sqnorm3(float x, float y, float z) { return (x * x + y * y) + z * z; }

static inline float sqdist_expand(float ax, float ay, float az, float sa, float bx, float by, float bz,
                                  float sb) {
    const float dot = fmaf(az, bz, fmaf(ay, by, ax * bx));
    return ((-2.0f * dot) + sa) + sb;
}

ORC_API void orc_square_distance(const float* src, int64_t aB, int64_t aN, int64_t aC, const float* dst,
                                 int64_t bB, int64_t bN, int64_t bC, int B, int N, int M,
                                 float* out /* [B,N,M] */) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < N; ++i) {
            const float* a = src + b * aB + i * aN;
            const float ax = a[0], ay = a[aC], az = a[2 * aC];
            const float sa = sqnorm3(ax, ay, az);
            float* o = out + ((int64_t)b * N + i) * M;
            for (int j = 0; j < M; ++j) {
                const float* q = dst + b * bB + j * bN;
                const float bx = q[0], by = q[bC], bz = q[2 * bC];
                o[j] = sqdist_expand(ax, ay, az, sa, bx, by, bz, sqnorm3(bx, by, bz));
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * query_ball_point  (pointnet_util.py:87-107)
 *   index cube = arange(N); entries with sqrdists > radius**2 become N (:102, fp32 compare);
 *   sort ascending and keep nsample (:103); entries still equal to N are replaced by the row's
 *   first entry (:104-106).  I.e. the first nsample in-ball points in ascending original index,
 *   padded with the first hit -- not the nearest ones.  src = new_xyz, dst = xyz in the
 *   square_distance call (:101).  A row with no hit at all keeps N (the reference would too).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_ball_query(float radius2, int nsample, const float* xyz, int64_t xB, int64_t xN, int64_t xC,
                            const float* new_xyz, int64_t qB, int64_t qN, int64_t qC, int B, int N, int S,
                            int64_t* out /* [B,S,nsample] */) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* px = (float*)malloc(sizeof(float) * (size_t)N * 4);
        float* py = px + N;
        float* pz = py + N;
        float* pn = pz + N;
        for (int j = 0; j < N; ++j) {
            const float* q = xyz + b * xB + j * xN;
            px[j] = q[0];
            py[j] = q[xC];
            pz[j] = q[2 * xC];
            pn[j] = sqnorm3(px[j], py[j], pz[j]);
        }
        for (int s = 0; s < S; ++s) {
            const float* a = new_xyz + b * qB + s * qN;
            const float ax = a[0], ay = a[qC], az = a[2 * qC];
            const float sa = sqnorm3(ax, ay, az);
            int64_t* o = out + ((int64_t)b * S + s) * nsample;
            int cnt = 0;
            for (int j = 0; j < N && cnt < nsample; ++j) {
                const float d = sqdist_expand(ax, ay, az, sa, px[j], py[j], pz[j], pn[j]);
                if (!(d > radius2)) o[cnt++] = j;
            }
            const int64_t first = cnt > 0 ? o[0] : (int64_t)N;
            for (; cnt < nsample; ++cnt) o[cnt] = first;
        }
        free(px);
    }
}

/* ------------------------------------------------------------------------------------------
 * index_points  (pointnet_util.py:43-60): out[b, m, :] = points[b, idx[b, m], :]
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_index_points(const float* points, int64_t pB, int64_t pN, int64_t pC, int B, int C,
                              const int64_t* idx, int M, float* out /* [B,M,C] */) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int m = 0; m < M; ++m) {
            const float* row = points + b * pB + idx[(int64_t)b * M + m] * pN;
            float* o = out + ((int64_t)b * M + m) * C;
            for (int c = 0; c < C; ++c) o[c] = row[c * pC];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * grouping of sample_and_group (pointnet_util.py:127-131, SSG order cat([xyz_rel, feats])) and of
 * PointNetSetAbstractionMsg.forward (:243-247, MSG order cat([feats, xyz_rel])).
 *   out[b, s, k, :] with 3+D channels; xyz_rel = xyz[idx] - new_xyz[s] (fp32 subtract).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_group(const float* xyz, int64_t xB, int64_t xN, int64_t xC, const float* feat, int64_t fB,
                       int64_t fN, int64_t fC, int D, const float* new_xyz, int64_t qB, int64_t qN,
                       int64_t qC, const int64_t* idx, int B, int S, int K, int msg_order,
                       float* out /* [B,S,K,3+D] */) {
    const int C = 3 + D;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int s = 0; s < S; ++s) {
            const float* c = new_xyz + b * qB + s * qN;
            for (int k = 0; k < K; ++k) {
                const int64_t j = idx[((int64_t)b * S + s) * K + k];
                const float* p = xyz + b * xB + j * xN;
                float* o = out + (((int64_t)b * S + s) * K + k) * C;
                float* oxyz = msg_order ? o + D : o;
                float* ofeat = msg_order ? o : o + 3;
                oxyz[0] = p[0] - c[0];
                oxyz[1] = p[xC] - c[qC];
                oxyz[2] = p[2 * xC] - c[2 * qC];
                if (D > 0) {
                    const float* f = feat + b * fB + j * fN;
                    for (int d = 0; d < D; ++d) ofeat[d] = f[d * fC];
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * 1x1 convolution / linear layer over rows, optional eval-mode BatchNorm, optional ReLU.
 *   Conv2d(c_in,c_out,1) + BatchNorm2d + relu (pointnet_util.py:195-197, :253-255),
 *   Conv1d + BatchNorm1d + relu (:310-312, pointnet2.py:172; pointnet.py passim), nn.Linear.
 *   y[r, co] = sum_ci x[r, ci] * w[co, ci] + bias[co];
 *   eval BN: (y - running_mean) / sqrt(running_var + eps) * gamma + beta  (NOT folded: the
 *   oracle keeps conv and BN as the two steps the reference executes).
 *   w_bstride lets each batch item use its own weight matrix (torch.bmm in pointnet.py:105-107).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_linear(const float* x, int64_t ldx, int64_t x_bstride, const float* w, int64_t w_bstride,
                        const float* bias, const float* bn_gamma, const float* bn_beta,
                        const float* bn_mean, const float* bn_var, float bn_eps, int relu, int B,
                        int64_t rows, int cin, int cout, float* y, int64_t ldy, int64_t y_bstride) {
    for (int b = 0; b < B; ++b) {
        const float* wb = w + b * w_bstride;
        float* wt = (float*)malloc(sizeof(float) * (size_t)cin * cout); /* [cin][cout] */
        for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < cin; ++ci) wt[(size_t)ci * cout + co] = wb[(size_t)co * cin + ci];
        const float* xb = x + b * x_bstride;
        float* yb = y + b * y_bstride;
#pragma omp parallel
        {
            float* acc = (float*)malloc(sizeof(float) * (size_t)cout);
#pragma omp for schedule(static)
            for (int64_t r = 0; r < rows; ++r) {
                const float* xr = xb + r * ldx;
                for (int co = 0; co < cout; ++co) acc[co] = 0.0f;
                for (int ci = 0; ci < cin; ++ci) {
                    const float xv = xr[ci];
                    const float* wrow = wt + (size_t)ci * cout;
#pragma omp simd
                    for (int co = 0; co < cout; ++co) acc[co] = fmaf(xv, wrow[co], acc[co]);
                }
                float* yr = yb + r * ldy;
                for (int co = 0; co < cout; ++co) {
                    float v = acc[co] + (bias ? bias[co] : 0.0f);
                    if (bn_gamma) {
                        const float invstd = 1.0f / sqrtf(bn_var[co] + bn_eps);
                        v = (v - bn_mean[co]) * invstd * bn_gamma[co] + bn_beta[co];
                    }
                    if (relu && !(v > 0.0f)) v = 0.0f;
                    yr[co] = v;
                }
            }
            free(acc);
        }
        free(wt);
    }
}

/* max over the K consecutive rows of each group: torch.max(new_points, 2)[0]
 * (pointnet_util.py:199, :256) and the global max over points (pointnet.py:35, 74, 122). */
ORC_API void orc_group_max(const float* x, int64_t ldx, int64_t groups, int K, int C, float* y, int64_t ldy) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < groups; ++g) {
        float* o = y + g * ldy;
        const float* r0 = x + g * K * ldx;
        for (int c = 0; c < C; ++c) o[c] = r0[c];
        for (int k = 1; k < K; ++k) {
            const float* r = r0 + (int64_t)k * ldx;
            for (int c = 0; c < C; ++c) o[c] = r[c] > o[c] ? r[c] : o[c];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * 3-NN inverse-distance interpolation  (PointNetFeaturePropagation.forward, pointnet_util.py:292-301)
 *   dists = square_distance(xyz1, xyz2) (:295, src = xyz1); sort, keep 3 (:296-297);
 *   dists < 1e-10 -> 1e-10 (:298); weight = 1/d (:299), normalised by the sum (:300);
 *   interpolated = sum_k points2[idx_k] * weight_k (:301).
 *   Selection here: the 3 smallest by (distance, index).  The reference's sort is not stable, so
 *   rows whose 3rd and 4th smallest distances are equal are flagged in tie[] (tests skip them).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_three_nn(const float* xyz1, int64_t aB, int64_t aN, int64_t aC, const float* xyz2, int64_t bB,
                          int64_t bN, int64_t bC, int B, int N, int S, int64_t* idx /* [B,N,3] */,
                          float* weight /* [B,N,3] */, float* dist3 /* [B,N,3] raw, may be NULL */,
                          uint8_t* tie /* [B,N], may be NULL */) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* sx = (float*)malloc(sizeof(float) * (size_t)S * 4);
        float* sy = sx + S;
        float* sz = sy + S;
        float* sn = sz + S;
        for (int j = 0; j < S; ++j) {
            const float* q = xyz2 + b * bB + j * bN;
            sx[j] = q[0];
            sy[j] = q[bC];
            sz[j] = q[2 * bC];
            sn[j] = sqnorm3(sx[j], sy[j], sz[j]);
        }
        for (int i = 0; i < N; ++i) {
            const float* a = xyz1 + b * aB + i * aN;
            const float ax = a[0], ay = a[aC], az = a[2 * aC];
            const float sa = sqnorm3(ax, ay, az);
            float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY, d3 = INFINITY;
            int64_t i0 = 0, i1 = 0, i2 = 0;
            for (int j = 0; j < S; ++j) {
                const float d = sqdist_expand(ax, ay, az, sa, sx[j], sy[j], sz[j], sn[j]);
                if (d < d0) { d3 = d2; d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = j; }
                else if (d < d1) { d3 = d2; d2 = d1; i2 = i1; d1 = d; i1 = j; }
                else if (d < d2) { d3 = d2; d2 = d; i2 = j; }
                else if (d < d3) { d3 = d; }
            }
            const int64_t o = ((int64_t)b * N + i) * 3;
            idx[o] = i0; idx[o + 1] = i1; idx[o + 2] = i2;
            if (dist3) { dist3[o] = d0; dist3[o + 1] = d1; dist3[o + 2] = d2; }
            if (tie) tie[(int64_t)b * N + i] = (uint8_t)(S > 3 && d2 == d3);
            const float c0 = d0 < 1e-10f ? 1e-10f : d0;
            const float c1 = d1 < 1e-10f ? 1e-10f : d1;
            const float c2 = d2 < 1e-10f ? 1e-10f : d2;
            const float w0 = 1.0f / c0, w1 = 1.0f / c1, w2 = 1.0f / c2;
            const float norm = (w0 + w1) + w2;
            weight[o] = w0 / norm; weight[o + 1] = w1 / norm; weight[o + 2] = w2 / norm;
        }
        free(sx);
    }
}

/* interpolated[b,n,:] = sum_k points2[b, idx[b,n,k], :] * weight[b,n,k]   (pointnet_util.py:301) */
ORC_API void orc_three_interpolate(const float* points2, int64_t pB, int64_t pN, int64_t pC, int D,
                                   const int64_t* idx, const float* weight, int B, int N, float* out,
                                   int64_t ldo, int64_t o_bstride) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < N; ++i) {
            const int64_t o = ((int64_t)b * N + i) * 3;
            const float* r0 = points2 + b * pB + idx[o] * pN;
            const float* r1 = points2 + b * pB + idx[o + 1] * pN;
            const float* r2 = points2 + b * pB + idx[o + 2] * pN;
            const float w0 = weight[o], w1 = weight[o + 1], w2 = weight[o + 2];
            float* y = out + b * o_bstride + i * ldo;
            for (int d = 0; d < D; ++d)
                y[d] = (r0[d * pC] * w0 + r1[d * pC] * w1) + r2[d * pC] * w2;
        }
    }
}

/* F.log_softmax over the channel dimension of each row (pointnet2.py:174, pointnet.py:251). */
ORC_API void orc_log_softmax(const float* x, int64_t ldx, int64_t rows, int C, float* y, int64_t ldy) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        const float* xr = x + r * ldx;
        float* yr = y + r * ldy;
        float m = xr[0];
        for (int c = 1; c < C; ++c) m = xr[c] > m ? xr[c] : m;
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += expf(xr[c] - m);
        const float ls = logf(s);
        for (int c = 0; c < C; ++c) yr[c] = (xr[c] - m) - ls;
    }
}
