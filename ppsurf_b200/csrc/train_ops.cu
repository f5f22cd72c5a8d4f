// Training step (BASELINE config 5, `pps.py fit`): forward-in-train-mode and backward kernels of everything that is not a
// dense contraction (those go through pps_gemm).  All tensors are row-major fp32 [rows, channels]; "groups" are runs of
// consecutive rows (a sample for InstanceNorm, a neighbourhood / a patch / the 64 interpolation neighbours of a query for the
// segment reductions).  Reference call sites (source/base/nn.py unless noted):
//   norm_*        BatchNorm1d in train mode (batch statistics + running update; 162-190, 305-373, 376-417, 438-450, 508-548)
//                 and InstanceNorm2d(affine) of FKAConvLayer (586-587, 630, 638), with the following ReLU / SiLU fused
//   seg_max_*     max over the patch points (STN, 170-172), over the 16 neighbours weighted by the distance weights (631-633, 639-641)
//                 and over all points of a sample (535)
//   gather_max_*  max_pool(x, ids) (677-680);  gather / scatter_add: batch_gather (655-674) and its gradient
//   attn_pool_*   softmax over the neighbours, mean over the heads, weighted sum (source/poco_model.py:413-416; AttentionPoco 84-96)
//   fka_*         FKAConvLayer.forward 592-652: neighbourhood geometry + norm_radius update (598-616), distance weights (618-624),
//                 the per-point [C_in x 16].[16 x 16] feature product (647-649)
//   ce_*          cross entropy of compute_loss (source/poco_model.py:75-88);  dropout: MLP (376-417, p = 0.3)
#include <algorithm>

#include "common.cuh"

namespace pps {
namespace train {

constexpr int kT = 256;

__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == 1) return fmaxf(z, 0.f);
    if (act == 2) return z / (1.f + __expf(-z));
    return z;
}
// derivative of the activation at pre-activation z
__device__ __forceinline__ float act_grad(float z, int act) {
    if (act == 1) return z > 0.f ? 1.f : 0.f;
    if (act == 2) {
        const float s = 1.f / (1.f + __expf(-z));
        return s * (1.f + z * (1.f - s));
    }
    return 1.f;
}

// ---- column statistics -----------------------------------------------------------------------------------------------------
// sums[g, c, 0..1] += sum / sum of squares over the block's row slab; fp64 atomics (one per block, channel and statistic)
__global__ void __launch_bounds__(kT) norm_stats_kernel(const float* __restrict__ x, long long rows, int c, long long slab, double* sums) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const float* xg = x + (long long)g * rows * c;
    // thread layout: cw channel lanes x (256 / cw) row lanes
    const int cw = c >= 32 ? 32 : (c >= 16 ? 16 : (c >= 8 ? 8 : (c >= 4 ? 4 : (c >= 2 ? 2 : 1))));
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, nry = kT / cw;
    __shared__ double sh[2][kT];
    for (int cc = tx; cc < c; cc += cw) {
        double s = 0.0, s2 = 0.0;
        for (long long r = r0 + ty; r < r1; r += nry) {
            const float v = xg[r * c + cc];
            s += v;
            s2 += (double)v * v;
        }
        sh[0][threadIdx.x] = s;
        sh[1][threadIdx.x] = s2;
        __syncthreads();
        if (ty == 0) {
            for (int j = 1; j < nry; ++j) {
                s += sh[0][j * cw + tx];
                s2 += sh[1][j * cw + tx];
            }
            atomicAdd(&sums[((long long)g * c + cc) * 2 + 0], s);
            atomicAdd(&sums[((long long)g * c + cc) * 2 + 1], s2);
        }
        __syncthreads();
    }
}
__global__ void norm_finalize_kernel(const double* __restrict__ sums, long long count, long long gc, float* mean, float* var) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gc) return;
    const double m = sums[2 * i] / count;
    mean[i] = (float)m;
    var[i] = (float)fmax(sums[2 * i + 1] / count - m * m, 0.0);  // biased
}
// y = act((x - mean) * rsqrt(var + eps) * gamma + beta)
__global__ void __launch_bounds__(kT) norm_apply_kernel(const float* __restrict__ x, long long rows, int c, const float* __restrict__ mean,
                                                        const float* __restrict__ var, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int act, float* __restrict__ y,
                                                        long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cc = (int)(i % c);
    const long long g = i / ((long long)rows * c);
    const float m = mean[g * c + cc], rs = rsqrtf(var[g * c + cc] + eps);
    const float z = (x[i] - m) * rs * gamma[cc] + beta[cc];
    y[i] = act_fwd(z, act);
}
// red[g, c, 0] += sum dz, red[g, c, 1] += sum dz * xhat   (dz = dy * act'(z))
__global__ void __launch_bounds__(kT) norm_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int c,
                                                             long long slab, const float* __restrict__ mean, const float* __restrict__ var,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                             int act, double* red) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const float* xg = x + (long long)g * rows * c;
    const float* dg = dy + (long long)g * rows * c;
    const int cw = c >= 32 ? 32 : (c >= 16 ? 16 : (c >= 8 ? 8 : (c >= 4 ? 4 : (c >= 2 ? 2 : 1))));
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, nry = kT / cw;
    __shared__ double sh[2][kT];
    for (int cc = tx; cc < c; cc += cw) {
        const float m = mean[(long long)g * c + cc], rs = rsqrtf(var[(long long)g * c + cc] + eps), ga = gamma[cc], be = beta[cc];
        double s = 0.0, s2 = 0.0;
        for (long long r = r0 + ty; r < r1; r += nry) {
            const float xh = (xg[r * c + cc] - m) * rs;
            const float dz = dg[r * c + cc] * act_grad(xh * ga + be, act);
            s += dz;
            s2 += (double)dz * xh;
        }
        sh[0][threadIdx.x] = s;
        sh[1][threadIdx.x] = s2;
        __syncthreads();
        if (ty == 0) {
            for (int j = 1; j < nry; ++j) {
                s += sh[0][j * cw + tx];
                s2 += sh[1][j * cw + tx];
            }
            atomicAdd(&red[((long long)g * c + cc) * 2 + 0], s);
            atomicAdd(&red[((long long)g * c + cc) * 2 + 1], s2);
        }
        __syncthreads();
    }
}
// dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)); dgamma[c] += sum_g red[g,c,1], dbeta[c] += sum_g red[g,c,0] (block 0)
__global__ void __launch_bounds__(kT) norm_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int c,
                                                            int groups, const float* __restrict__ mean, const float* __restrict__ var,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                            int act, const double* __restrict__ red, float* __restrict__ dx, float* dgamma,
                                                            float* dbeta, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int cc = threadIdx.x; cc < c; cc += blockDim.x) {
            double a = 0.0, b = 0.0;
            for (int g = 0; g < groups; ++g) {
                b += red[((long long)g * c + cc) * 2 + 0];
                a += red[((long long)g * c + cc) * 2 + 1];
            }
            dgamma[cc] = (float)a;
            dbeta[cc] = (float)b;
        }
    }
    if (i >= total) return;
    const int cc = (int)(i % c);
    const long long g = i / ((long long)rows * c);
    const float m = mean[g * c + cc], rs = rsqrtf(var[g * c + cc] + eps), ga = gamma[cc];
    const float xh = (x[i] - m) * rs;
    const float dz = dy[i] * act_grad(xh * ga + beta[cc], act);
    const double inv = 1.0 / (double)rows;
    const float mdz = (float)(red[(g * c + cc) * 2 + 0] * inv), mdzx = (float)(red[(g * c + cc) * 2 + 1] * inv);
    dx[i] = ga * rs * (dz - mdz - xh * mdzx);
}
// running = (1 - momentum) * running + momentum * batch statistic (unbiased variance), source: torch BatchNorm semantics
__global__ void bn_running_kernel(const float* __restrict__ mean, const float* __restrict__ var, long long count, float momentum, int c,
                                  float* running_mean, float* running_var) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float unbiased = count > 1 ? var[i] * ((float)count / (float)(count - 1)) : var[i];
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * mean[i];
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * unbiased;
}

// ---- elementwise ---------------------------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ x, long long n, int act, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = act_fwd(x[i], act);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n, int act, float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = dy[i] * act_grad(x[i], act);
}
__device__ __forceinline__ uint32_t hash32(uint32_t a, uint32_t b) {
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__global__ void dropout_fwd_kernel(const float* __restrict__ x, long long n, float p, uint32_t seed, const uint32_t* __restrict__ draw,
                                   float* __restrict__ y, uint8_t* mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw) seed += draw[0] * 0x9E3779B9u;  // device-side draw counter: a replayed CUDA graph gets a fresh mask every time
    const uint32_t h = hash32((uint32_t)i ^ seed, (uint32_t)(i >> 32) + seed * 31u);
    const bool keep = (h >> 8) * (1.f / 16777216.f) >= p;
    mask[i] = keep;
    y[i] = keep ? x[i] / (1.f - p) : 0.f;
}
__global__ void dropout_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask, long long n, float p, float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = mask[i] ? dy[i] / (1.f - p) : 0.f;
}
// y[r, c] = x[r, c] * w[r]
__global__ void rowscale_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, long long rows, int c, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * c) y[i] = x[i] * w[i / c];
}
// dx = dy * w; dw[r] = sum_c dy * x   (one thread per row: c is 16 in the kernel-weight MLP)
__global__ void rowscale_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy, long long rows,
                                    int c, float* __restrict__ dx, float* __restrict__ dw) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float wr = w[r];
    float s = 0.f;
    for (int j = 0; j < c; ++j) {
        const float d = dy[r * c + j];
        dx[r * c + j] = d * wr;
        s = fmaf(d, x[r * c + j], s);
    }
    dw[r] = s;
}
// out[g, s, 0:c] = x[g, s, :], out[g, s, c:2c] = v[g, :]
__global__ void concat_bcast_fwd_kernel(const float* __restrict__ x, const float* __restrict__ v, long long groups, int s, int c,
                                        float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * s * 2 * c) return;
    const int cc = (int)(i % (2 * c));
    const long long row = i / (2 * c);
    out[i] = cc < c ? x[row * c + cc] : v[(row / s) * c + (cc - c)];
}
// dx[g, s, :] = dout[g, s, 0:c]; dv[g, :] = sum_s dout[g, s, c:2c]
__global__ void concat_bcast_bwd_kernel(const float* __restrict__ dout, long long groups, int s, int c, float* __restrict__ dx,
                                        float* __restrict__ dv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * c) return;
    const int cc = (int)(i % c);
    const long long g = i / c;
    float acc = 0.f;
    for (int j = 0; j < s; ++j) {
        const long long row = g * s + j;
        dx[row * c + cc] = dout[row * 2 * c + cc];
        acc += dout[row * 2 * c + c + cc];
    }
    dv[i] = acc;
}

// ---- gathers -------------------------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, long long m, int c, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * c) return;
    const long long r = i / c;
    y[i] = x[(long long)idx[r] * c + (i % c)];
}
__global__ void scatter_add_rows_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx, long long m, int c, float* dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * c) return;
    const long long r = i / c;
    atomicAdd(dx + (long long)idx[r] * c + (i % c), dy[i]);
}
// y[g, c] = max_s x[g, s, c] * w[g, s] (w nullable), arg = first maximising s
__global__ void seg_max_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, long long groups, int s, int c,
                                   float* __restrict__ y, int32_t* __restrict__ arg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * c) return;
    const int cc = (int)(i % c);
    const long long g = i / c;
    float best = -INFINITY;
    int bi = 0;
    for (int j = 0; j < s; ++j) {
        float v = x[(g * s + j) * c + cc];
        if (w) v *= w[g * s + j];
        if (v > best) {
            best = v;
            bi = j;
        }
    }
    y[i] = best;
    arg[i] = bi;
}
// dx[g, arg, c] = dy[g, c] * w[g, arg]; dw[g, arg] += dy[g, c] * x[g, arg, c]   (dx and dw zero-filled by the launcher)
__global__ void seg_max_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ arg, const float* __restrict__ x,
                                   const float* __restrict__ w, long long groups, int s, int c, float* dx, float* dw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * c) return;
    const int cc = (int)(i % c);
    const long long g = i / c;
    const long long row = g * s + arg[i];
    const float d = dy[i];
    dx[row * c + cc] = w ? d * w[row] : d;
    if (dw) atomicAdd(dw + row, d * x[row * c + cc]);
}
// max_pool: y[b, n, c] = max_j x[b, ids[b, n, j], c]; arg = the winning SOURCE ROW (b * n_in + id)
__global__ void gather_max_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ ids, long long b, long long n_in, long long n_s,
                                      int c, int kn, float* __restrict__ y, int32_t* __restrict__ arg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * n_s * c) return;
    const int cc = (int)(i % c);
    const long long p = i / c, bi = p / n_s;
    float best = -INFINITY;
    long long br = 0;
    for (int j = 0; j < kn; ++j) {
        const long long row = bi * n_in + ids[p * kn + j];
        const float v = x[row * c + cc];
        if (v > best) {
            best = v;
            br = row;
        }
    }
    y[i] = best;
    arg[i] = (int32_t)br;
}
__global__ void gather_max_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ arg, long long total, int c, float* dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    atomicAdd(dx + (long long)arg[i] * c + (i % c), dy[i]);
}

// ---- attention pooling ---------------------------------------------------------------------------------------------------------
// one block per group: prob[s, h] = softmax_s(scores[s, h]); a[s] = mean_h prob; out[c] = sum_s a[s] * v[s, c].
// The group's score tile sits in shared memory; threads are laid out as (head lanes) x (parts of s) for the column-wise softmax,
// one warp per row for the head mean, one thread per channel for the weighted sum.
__global__ void __launch_bounds__(kT) attn_pool_fwd_kernel(const float* __restrict__ scores, const float* __restrict__ v, int s, int h, int c,
                                                           float* __restrict__ prob, float* __restrict__ a, float* __restrict__ out) {
    extern __shared__ float sh[];  // tile[s * h], a[s], red[kT]
    float* tile = sh;
    float* sa = sh + s * h;
    float* red = sa + s;
    const long long g = blockIdx.x;
    const float* sc = scores + g * s * h;
    for (int i = threadIdx.x; i < s * h; i += kT) tile[i] = sc[i];
    __syncthreads();
    const int lanes_h = h < kT ? h : kT, parts = kT / lanes_h;
    const int hl = threadIdx.x % lanes_h, part = threadIdx.x / lanes_h;
    for (int hh = hl; hh < h; hh += lanes_h) {
        float mx = -INFINITY;
        for (int j = part; j < s; j += parts) mx = fmaxf(mx, tile[j * h + hh]);
        red[threadIdx.x] = mx;
        __syncthreads();
        for (int q = 0; q < parts; ++q) mx = fmaxf(mx, red[q * lanes_h + hl]);
        __syncthreads();
        float sum = 0.f;
        for (int j = part; j < s; j += parts) {
            const float e = __expf(tile[j * h + hh] - mx);
            tile[j * h + hh] = e;
            sum += e;
        }
        red[threadIdx.x] = sum;
        __syncthreads();
        sum = 0.f;
        for (int q = 0; q < parts; ++q) sum += red[q * lanes_h + hl];
        const float inv = 1.f / sum;
        for (int j = part; j < s; j += parts) tile[j * h + hh] *= inv;
        __syncthreads();
    }
    float* pr = prob + g * s * h;
    for (int i = threadIdx.x; i < s * h; i += kT) pr[i] = tile[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < s; j += kT / 32) {
        float acc = 0.f;
        for (int hh = lane; hh < h; hh += 32) acc += tile[j * h + hh];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            sa[j] = acc / h;
            a[g * s + j] = acc / h;
        }
    }
    __syncthreads();
    const float* vg = v + g * s * c;
    for (int cc = threadIdx.x; cc < c; cc += kT) {
        float acc = 0.f;
        for (int j = 0; j < s; ++j) acc = fmaf(sa[j], vg[j * c + cc], acc);
        out[g * c + cc] = acc;
    }
}
// dv[s, c] = a[s] * dout[c]; da[s] = dout . v[s, :]; dscores[s, h] = prob[s, h] * (da[s] - sum_s' prob[s', h] * da[s']) / h
__global__ void __launch_bounds__(kT) attn_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ prob,
                                                           const float* __restrict__ a, const float* __restrict__ v, int s, int h, int c,
                                                           float* __restrict__ dscores, float* __restrict__ dv) {
    extern __shared__ float sh[];  // da[s]
    const long long g = blockIdx.x;
    const float* vg = v + g * s * c;
    const float* dg = dout + g * c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < s; j += nw) {
        float acc = 0.f;
        const float aj = a[g * s + j];
        for (int cc = lane; cc < c; cc += 32) {
            const float d = dg[cc];
            acc = fmaf(d, vg[j * c + cc], acc);
            dv[(g * s + j) * c + cc] = aj * d;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sh[j] = acc;
    }
    __syncthreads();
    const float* pr = prob + g * s * h;
    float* ds = dscores + g * s * h;
    for (int hh = threadIdx.x; hh < h; hh += blockDim.x) {
        float dot = 0.f;
        for (int j = 0; j < s; ++j) dot = fmaf(pr[j * h + hh], sh[j], dot);
        for (int j = 0; j < s; ++j) ds[j * h + hh] = pr[j * h + hh] * (sh[j] - dot) * (1.f / h);
    }
}

// ---- FKAConv ---------------------------------------------------------------------------------------------------------------------
// offs[r, 0..2] = pts[b, ids[r]] - support[p]; dist[r]; radius_sum += sum over the block's points of max_j dist (fp64 atomic)
__global__ void __launch_bounds__(kT) fka_geometry_kernel(const float* __restrict__ pts, const float* __restrict__ support,
                                                          const int32_t* __restrict__ ids, long long b, long long n_in, long long n_s, int kn,
                                                          float* __restrict__ offs, float* __restrict__ dist, double* radius_sum) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // flattened support point
    float mx = 0.f;
    if (p < b * n_s) {
        const long long bi = p / n_s;
        const float sx = support[p * 3], sy = support[p * 3 + 1], sz = support[p * 3 + 2];
        for (int j = 0; j < kn; ++j) {
            const long long src = (bi * n_in + ids[p * kn + j]) * 3;
            const float dx = pts[src] - sx, dy = pts[src + 1] - sy, dz = pts[src + 2] - sz;
            const long long r = p * kn + j;
            offs[r * 3] = dx;
            offs[r * 3 + 1] = dy;
            offs[r * 3 + 2] = dz;
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            dist[r] = d;
            mx = fmaxf(mx, d);
        }
    }
    __shared__ float sh[kT];
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int o = kT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(radius_sum, (double)sh[0]);
}
// norm_radius = (1 - momentum) * norm_radius + momentum * radius_sum / points   (train mode, nn.py:608-613)
__global__ void fka_radius_update_kernel(const double* radius_sum, long long points, float momentum, float* norm_radius) {
    norm_radius[0] = norm_radius[0] * (1.f - momentum) + (float)(radius_sum[0] / (double)points) * momentum;
}
// offs /= norm_radius;  dw[r] = sigmoid(-alpha * d + beta) / (sum + (sum == 0) + 1e-6) * kn  (one thread per point)
__global__ void fka_weights_fwd_kernel(float* offs, const float* __restrict__ dist, long long points, int kn, const float* __restrict__ alpha,
                                       const float* __restrict__ beta, const float* __restrict__ norm_radius, float* __restrict__ sig,
                                       float* __restrict__ dw) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= points) return;
    const float al = alpha[0], be = beta[0], nr = norm_radius[0];
    float sum = 0.f;
    for (int j = 0; j < kn; ++j) {
        const long long r = p * kn + j;
        const float s = 1.f / (1.f + __expf(al * dist[r] - be));
        sig[r] = s;
        sum += s;
        offs[r * 3] /= nr;
        offs[r * 3 + 1] /= nr;
        offs[r * 3 + 2] /= nr;
    }
    const float den = sum + (sum == 0.f ? 1.f : 0.f) + 1e-6f;
    for (int j = 0; j < kn; ++j) dw[p * kn + j] = sig[p * kn + j] / den * kn;
}
// ddw [R] -> dalpha, dbeta (fp64 atomics, one per block)
__global__ void __launch_bounds__(kT) fka_weights_bwd_kernel(const float* __restrict__ ddw, const float* __restrict__ sig,
                                                             const float* __restrict__ dist, long long points, int kn, double* dalpha_dbeta) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float da = 0.f, db = 0.f;
    if (p < points) {
        float sum = 0.f, gs = 0.f;
        for (int j = 0; j < kn; ++j) {
            sum += sig[p * kn + j];
            gs = fmaf(ddw[p * kn + j], sig[p * kn + j], gs);
        }
        const float den = sum + (sum == 0.f ? 1.f : 0.f) + 1e-6f;
        for (int j = 0; j < kn; ++j) {
            const long long r = p * kn + j;
            const float s = sig[r];
            const float ds = kn * (ddw[r] / den - gs / (den * den));  // d loss / d sigmoid_j
            const float dlogit = ds * s * (1.f - s);
            da = fmaf(dlogit, -dist[r], da);
            db += dlogit;
        }
    }
    __shared__ float sh[2][kT];
    sh[0][threadIdx.x] = da;
    sh[1][threadIdx.x] = db;
    __syncthreads();
    for (int o = kT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(&dalpha_dbeta[0], (double)sh[0][0]);
        atomicAdd(&dalpha_dbeta[1], (double)sh[1][0]);
    }
}
// feat[p, c * 16 + m] = sum_j x[b, ids[p, j], c] * mat[p, j, m]     (one block per point, thread = (c, m) pairs)
__global__ void __launch_bounds__(kT) fka_feat_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ ids,
                                                          const float* __restrict__ mat, long long n_in, long long n_s, int kn, int cin,
                                                          float* __restrict__ feat) {
    __shared__ float smat[16 * 16];
    __shared__ int32_t sid[16];
    const long long p = blockIdx.x, bi = p / n_s;
    if (threadIdx.x < kn * 16) smat[threadIdx.x] = mat[p * kn * 16 + threadIdx.x];
    if (threadIdx.x < kn) sid[threadIdx.x] = ids[p * kn + threadIdx.x];
    __syncthreads();
    for (int e = threadIdx.x; e < cin * 16; e += blockDim.x) {
        const int cc = e >> 4, m = e & 15;
        float acc = 0.f;
        for (int j = 0; j < kn; ++j) acc = fmaf(x[(bi * n_in + sid[j]) * cin + cc], smat[j * 16 + m], acc);
        feat[p * cin * 16 + e] = acc;
    }
}
// dmat[p, j, m] = sum_c x[ids[p, j], c] * dfeat[p, c, m];  dx[ids[p, j], c] += sum_m mat[p, j, m] * dfeat[p, c, m]
__global__ void __launch_bounds__(kT) fka_feat_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ x,
                                                          const int32_t* __restrict__ ids, const float* __restrict__ mat, long long n_in,
                                                          long long n_s, int kn, int cin, float* dx, float* __restrict__ dmat) {
    __shared__ float smat[16 * 16];
    __shared__ int32_t sid[16];
    const long long p = blockIdx.x, bi = p / n_s;
    if (threadIdx.x < kn * 16) smat[threadIdx.x] = mat[p * kn * 16 + threadIdx.x];
    if (threadIdx.x < kn) sid[threadIdx.x] = ids[p * kn + threadIdx.x];
    __syncthreads();
    const float* df = dfeat + p * cin * 16;
    // dx: thread = (j, c)
    for (int e = threadIdx.x; e < kn * cin; e += blockDim.x) {
        const int j = e / cin, cc = e % cin;
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < 16; ++m) acc = fmaf(smat[j * 16 + m], df[cc * 16 + m], acc);
        atomicAdd(dx + (bi * n_in + sid[j]) * cin + cc, acc);
    }
    // dmat: thread = (j, m), loop over c
    if (threadIdx.x < kn * 16) {
        const int j = threadIdx.x >> 4, m = threadIdx.x & 15;
        const float* xr = x + (bi * n_in + sid[j]) * cin;
        float acc = 0.f;
        for (int cc = 0; cc < cin; ++cc) acc = fmaf(xr[cc], df[cc * 16 + m], acc);
        dmat[p * kn * 16 + threadIdx.x] = acc;
    }
}

// ---- loss ------------------------------------------------------------------------------------------------------------------------
// loss_rows[i] = logsumexp(logits[i, :]) - logits[i, target[i]]; loss_sum += sum (fp64 atomic)
__global__ void __launch_bounds__(kT) ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, long long m, int c,
                                                    float* __restrict__ loss_rows, double* loss_sum) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < m) {
        float mx = -INFINITY;
        for (int j = 0; j < c; ++j) mx = fmaxf(mx, logits[i * c + j]);
        float s = 0.f;
        for (int j = 0; j < c; ++j) s += expf(logits[i * c + j] - mx);
        const long long t = min(max((long long)target[i], 0ll), (long long)c - 1);  // labels outside [0, c) are clamped, never read out of bounds
        l = logf(s) + mx - logits[i * c + t];
        loss_rows[i] = l;
    }
    __shared__ float sh[kT];
    sh[threadIdx.x] = l;
    __syncthreads();
    for (int o = kT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(loss_sum, (double)sh[0]);
}
// dlogits[i, j] = scale[i] * (softmax(logits[i])[j] - [j == target[i]])
__global__ void ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, const float* __restrict__ scale,
                              long long m, int c, float* __restrict__ dlogits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float mx = -INFINITY;
    for (int j = 0; j < c; ++j) mx = fmaxf(mx, logits[i * c + j]);
    float s = 0.f;
    for (int j = 0; j < c; ++j) s += expf(logits[i * c + j] - mx);
    const float g = scale[i];
    const long long t = min(max((long long)target[i], 0ll), (long long)c - 1);
    for (int j = 0; j < c; ++j) dlogits[i * c + j] = g * (expf(logits[i * c + j] - mx) / s - (j == t ? 1.f : 0.f));
}
// out[c] (+)= sum_r x[r, c]   (bias gradients)
__global__ void __launch_bounds__(kT) colsum_kernel(const float* __restrict__ x, long long rows, int c, long long ld, long long slab, float* out) {
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const int cw = c >= 32 ? 32 : (c >= 16 ? 16 : (c >= 8 ? 8 : (c >= 4 ? 4 : (c >= 2 ? 2 : 1))));
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, nry = kT / cw;
    __shared__ float sh[kT];
    for (int cc = tx; cc < c; cc += cw) {
        float s = 0.f;
        for (long long r = r0 + ty; r < r1; r += nry) s += x[r * ld + cc];
        sh[threadIdx.x] = s;
        __syncthreads();
        if (ty == 0) {
            for (int j = 1; j < nry; ++j) s += sh[j * cw + tx];
            atomicAdd(out + cc, s);
        }
        __syncthreads();
    }
}


// ---- float4 variants (c % 4 == 0, 16-byte aligned rows): a block streams a slab of rows with its threads laid out as
// (c / 4 channel quads) x (row lanes); the per-channel constants are loaded once, the row loop has no index division ------------------
struct Quad {
    int lanes_c, lanes_r, cq, rl;  // channel-quad lanes, row lanes, this thread's quad lane and row lane
    __device__ explicit Quad(int c) {
        const int q = c >> 2;
        lanes_c = q < kT ? q : kT;
        lanes_r = kT / lanes_c;
        cq = threadIdx.x % lanes_c;
        rl = threadIdx.x / lanes_c;
    }
};
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

__global__ void __launch_bounds__(kT) norm_stats_v4_kernel(const float* __restrict__ x, long long rows, int c, long long slab, double* sums) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const float* xg = x + (long long)g * rows * c;
    const Quad t(c);
    __shared__ float sh[2][kT][4];
    for (int cq = t.cq; cq < (c >> 2); cq += t.lanes_c) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        // fp32 partial sums over at most slab / lanes_r rows (tens to hundreds), combined in fp64 below
        for (long long r = r0 + t.rl; r < r1; r += t.lanes_r) {
            const float4 v = ld4(xg + r * c + 4 * cq);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]); s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sh[0][threadIdx.x][j] = s[j];
            sh[1][threadIdx.x][j] = s2[j];
        }
        __syncthreads();
        if (t.rl == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double a = 0.0, b = 0.0;
                for (int l = 0; l < t.lanes_r; ++l) {
                    a += sh[0][l * t.lanes_c + t.cq][j];
                    b += sh[1][l * t.lanes_c + t.cq][j];
                }
                atomicAdd(&sums[((long long)g * c + 4 * cq + j) * 2 + 0], a);
                atomicAdd(&sums[((long long)g * c + 4 * cq + j) * 2 + 1], b);
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kT) norm_apply_v4_kernel(const float* __restrict__ x, long long rows, int c, long long slab,
                                                           const float* __restrict__ mean, const float* __restrict__ var,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                                                           float* __restrict__ y) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const long long base = (long long)g * rows * c;
    const Quad t(c);
    for (int cq = t.cq; cq < (c >> 2); cq += t.lanes_c) {
        float sc[4], sf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = 4 * cq + j;
            const float rs = rsqrtf(var[(long long)g * c + cc] + eps);
            sc[j] = rs * gamma[cc];
            sf[j] = beta[cc] - mean[(long long)g * c + cc] * sc[j];
        }
        for (long long r = r0 + t.rl; r < r1; r += t.lanes_r) {
            const float4 v = ld4(x + base + r * c + 4 * cq);
            float4 o;
            o.x = act_fwd(fmaf(v.x, sc[0], sf[0]), act);
            o.y = act_fwd(fmaf(v.y, sc[1], sf[1]), act);
            o.z = act_fwd(fmaf(v.z, sc[2], sf[2]), act);
            o.w = act_fwd(fmaf(v.w, sc[3], sf[3]), act);
            st4(y + base + r * c + 4 * cq, o);
        }
    }
}
__global__ void __launch_bounds__(kT) norm_bwd_reduce_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows,
                                                                int c, long long slab, const float* __restrict__ mean,
                                                                const float* __restrict__ var, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, int act, double* red) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const long long base = (long long)g * rows * c;
    const Quad t(c);
    __shared__ float sh[2][kT][4];
    for (int cq = t.cq; cq < (c >> 2); cq += t.lanes_c) {
        float m[4], rs[4], ga[4], be[4], s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = 4 * cq + j;
            m[j] = mean[(long long)g * c + cc];
            rs[j] = rsqrtf(var[(long long)g * c + cc] + eps);
            ga[j] = gamma[cc];
            be[j] = beta[cc];
        }
        for (long long r = r0 + t.rl; r < r1; r += t.lanes_r) {
            const float4 v = ld4(x + base + r * c + 4 * cq), d = ld4(dy + base + r * c + 4 * cq);
            const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (xv[j] - m[j]) * rs[j];
                const float dz = dv[j] * act_grad(fmaf(xh, ga[j], be[j]), act);
                s[j] += dz;
                s2[j] = fmaf(dz, xh, s2[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sh[0][threadIdx.x][j] = s[j];
            sh[1][threadIdx.x][j] = s2[j];
        }
        __syncthreads();
        if (t.rl == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double a = 0.0, b = 0.0;
                for (int l = 0; l < t.lanes_r; ++l) {
                    a += sh[0][l * t.lanes_c + t.cq][j];
                    b += sh[1][l * t.lanes_c + t.cq][j];
                }
                atomicAdd(&red[((long long)g * c + 4 * cq + j) * 2 + 0], a);
                atomicAdd(&red[((long long)g * c + 4 * cq + j) * 2 + 1], b);
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kT) norm_bwd_apply_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int c,
                                                               long long slab, const float* __restrict__ mean, const float* __restrict__ var,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                               int act, const double* __restrict__ red, float* __restrict__ dx) {
    const int g = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const long long base = (long long)g * rows * c;
    const Quad t(c);
    const double inv = 1.0 / (double)rows;
    for (int cq = t.cq; cq < (c >> 2); cq += t.lanes_c) {
        float m[4], rs[4], ga[4], be[4], mdz[4], mdzx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gc = (long long)g * c + 4 * cq + j;
            m[j] = mean[gc];
            rs[j] = rsqrtf(var[gc] + eps);
            ga[j] = gamma[4 * cq + j];
            be[j] = beta[4 * cq + j];
            mdz[j] = (float)(red[gc * 2 + 0] * inv);
            mdzx[j] = (float)(red[gc * 2 + 1] * inv);
        }
        for (long long r = r0 + t.rl; r < r1; r += t.lanes_r) {
            const float4 v = ld4(x + base + r * c + 4 * cq), d = ld4(dy + base + r * c + 4 * cq);
            const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (xv[j] - m[j]) * rs[j];
                const float dz = dv[j] * act_grad(fmaf(xh, ga[j], be[j]), act);
                o[j] = ga[j] * rs[j] * (dz - mdz[j] - xh * mdzx[j]);
            }
            st4(dx + base + r * c + 4 * cq, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}
// dgamma[c] = sum_g red[g,c,1], dbeta[c] = sum_g red[g,c,0]
__global__ void norm_param_grads_kernel(const double* __restrict__ red, int groups, int c, float* dgamma, float* dbeta) {
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= c) return;
    double a = 0.0, b = 0.0;
    for (int g = 0; g < groups; ++g) {
        b += red[((long long)g * c + cc) * 2 + 0];
        a += red[((long long)g * c + cc) * 2 + 1];
    }
    dgamma[cc] = (float)a;
    dbeta[cc] = (float)b;
}
__global__ void act_fwd_v4_kernel(const float* __restrict__ x, long long n4, int act, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = ld4(x + 4 * i);
    st4(y + 4 * i, make_float4(act_fwd(v.x, act), act_fwd(v.y, act), act_fwd(v.z, act), act_fwd(v.w, act)));
}
__global__ void act_bwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n4, int act, float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = ld4(x + 4 * i), d = ld4(dy + 4 * i);
    st4(dx + 4 * i, make_float4(d.x * act_grad(v.x, act), d.y * act_grad(v.y, act), d.z * act_grad(v.z, act), d.w * act_grad(v.w, act)));
}
__global__ void __launch_bounds__(kT) colsum_v4_kernel(const float* __restrict__ x, long long rows, int c, long long ld, long long slab,
                                                       float* out) {
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    const Quad t(c);
    __shared__ float sh[kT][4];
    for (int cq = t.cq; cq < (c >> 2); cq += t.lanes_c) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (long long r = r0 + t.rl; r < r1; r += t.lanes_r) {
            const float4 v = ld4(x + r * ld + 4 * cq);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[threadIdx.x][j] = s[j];
        __syncthreads();
        if (t.rl == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = 0.f;
                for (int l = 0; l < t.lanes_r; ++l) a += sh[l * t.lanes_c + t.cq][j];
                atomicAdd(out + 4 * cq + j, a);
            }
        }
        __syncthreads();
    }
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline unsigned blocks_for(long long n) { return (unsigned)std::max<long long>(1, ceil_div(n, kT)); }
inline long long slab_for(long long rows, long long groups) {
    // about 4 blocks per SM over all groups, at least 64 rows per block
    const long long want = std::max<long long>(1, 4 * kNumSMs / std::max<long long>(groups, 1));
    return std::max<long long>(64, ceil_div(rows, want));
}

}  // namespace train
}  // namespace pps

using namespace pps;
using namespace pps::train;
#define ST static_cast<cudaStream_t>(stream)

extern "C" {

size_t pps_norm_workspace_bytes(int64_t groups, int c) { return (size_t)groups * c * 2 * sizeof(double); }

int pps_norm_fwd(const float* x, int64_t groups, int64_t rows, int c, const float* gamma, const float* beta, float eps, int act, float* y,
                 float* mean, float* var, void* workspace, size_t workspace_bytes, void* stream) {
    PPS_CHECK_ARG(x && gamma && beta && y && mean && var && workspace, "pps_norm_fwd: null pointer");
    PPS_CHECK_ARG(groups > 0 && groups <= 65535 && rows > 0 && c > 0 && act >= 0 && act <= 2, "pps_norm_fwd: bad shape");
    if (workspace_bytes < pps_norm_workspace_bytes(groups, c)) {
        set_error("pps_norm_fwd: workspace too small");
        return PPS_ERR_WORKSPACE;
    }
    double* sums = static_cast<double*>(workspace);
    PPS_CUDA(cudaMemsetAsync(sums, 0, pps_norm_workspace_bytes(groups, c), ST));
    const long long slab = slab_for(rows, groups);
    const bool v4 = (c & 3) == 0 && aligned16(x) && aligned16(y);
    const dim3 grid((unsigned)ceil_div(rows, slab), (unsigned)groups);
    if (v4)
        norm_stats_v4_kernel<<<grid, kT, 0, ST>>>(x, rows, c, slab, sums);
    else
        norm_stats_kernel<<<grid, kT, 0, ST>>>(x, rows, c, slab, sums);
    PPS_LAUNCH_CHECK();
    norm_finalize_kernel<<<blocks_for(groups * c), kT, 0, ST>>>(sums, rows, groups * c, mean, var);
    PPS_LAUNCH_CHECK();
    const long long total = groups * rows * c;
    if (v4)
        norm_apply_v4_kernel<<<grid, kT, 0, ST>>>(x, rows, c, slab, mean, var, gamma, beta, eps, act, y);
    else
        norm_apply_kernel<<<blocks_for(total), kT, 0, ST>>>(x, rows, c, mean, var, gamma, beta, eps, act, y, total);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_norm_bwd(const float* x, const float* dy, int64_t groups, int64_t rows, int c, const float* gamma, const float* beta,
                 const float* mean, const float* var, float eps, int act, float* dx, float* dgamma, float* dbeta, void* workspace,
                 size_t workspace_bytes, void* stream) {
    PPS_CHECK_ARG(x && dy && gamma && beta && mean && var && dx && dgamma && dbeta && workspace, "pps_norm_bwd: null pointer");
    PPS_CHECK_ARG(groups > 0 && groups <= 65535 && rows > 0 && c > 0, "pps_norm_bwd: bad shape");
    if (workspace_bytes < pps_norm_workspace_bytes(groups, c)) {
        set_error("pps_norm_bwd: workspace too small");
        return PPS_ERR_WORKSPACE;
    }
    double* red = static_cast<double*>(workspace);
    PPS_CUDA(cudaMemsetAsync(red, 0, pps_norm_workspace_bytes(groups, c), ST));
    const long long slab = slab_for(rows, groups);
    const bool v4 = (c & 3) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx);
    const dim3 grid((unsigned)ceil_div(rows, slab), (unsigned)groups);
    const long long total = groups * rows * c;
    if (v4) {
        norm_bwd_reduce_v4_kernel<<<grid, kT, 0, ST>>>(x, dy, rows, c, slab, mean, var, gamma, beta, eps, act, red);
        PPS_LAUNCH_CHECK();
        norm_param_grads_kernel<<<blocks_for(c), kT, 0, ST>>>(red, (int)groups, c, dgamma, dbeta);
        PPS_LAUNCH_CHECK();
        norm_bwd_apply_v4_kernel<<<grid, kT, 0, ST>>>(x, dy, rows, c, slab, mean, var, gamma, beta, eps, act, red, dx);
    } else {
        norm_bwd_reduce_kernel<<<grid, kT, 0, ST>>>(x, dy, rows, c, slab, mean, var, gamma, beta, eps, act, red);
        PPS_LAUNCH_CHECK();
        norm_bwd_apply_kernel<<<blocks_for(total), kT, 0, ST>>>(x, dy, rows, c, (int)groups, mean, var, gamma, beta, eps, act, red, dx, dgamma,
                                                                dbeta, total);
    }
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_bn_running_update(const float* mean, const float* var, int64_t count, float momentum, int c, float* running_mean, float* running_var,
                          void* stream) {
    PPS_CHECK_ARG(mean && var && running_mean && running_var && c > 0, "pps_bn_running_update: bad argument");
    bn_running_kernel<<<blocks_for(c), kT, 0, ST>>>(mean, var, count, momentum, c, running_mean, running_var);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_act_fwd(const float* x, int64_t n, int act, float* y, void* stream) {
    PPS_CHECK_ARG(x && y && n >= 0, "pps_act_fwd: bad argument");
    if (n == 0) return PPS_OK;
    if ((n & 3) == 0 && aligned16(x) && aligned16(y))
        act_fwd_v4_kernel<<<blocks_for(n / 4), kT, 0, ST>>>(x, n / 4, act, y);
    else
        act_fwd_kernel<<<blocks_for(n), kT, 0, ST>>>(x, n, act, y);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_act_bwd(const float* x, const float* dy, int64_t n, int act, float* dx, void* stream) {
    PPS_CHECK_ARG(x && dy && dx && n >= 0, "pps_act_bwd: bad argument");
    if (n == 0) return PPS_OK;
    if ((n & 3) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx))
        act_bwd_v4_kernel<<<blocks_for(n / 4), kT, 0, ST>>>(x, dy, n / 4, act, dx);
    else
        act_bwd_kernel<<<blocks_for(n), kT, 0, ST>>>(x, dy, n, act, dx);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_dropout_fwd(const float* x, int64_t n, float p, uint32_t seed, const uint32_t* draw_counter, float* y, uint8_t* mask, void* stream) {
    PPS_CHECK_ARG(x && y && mask && n >= 0 && p >= 0.f && p < 1.f, "pps_dropout_fwd: bad argument");
    if (n == 0) return PPS_OK;
    dropout_fwd_kernel<<<blocks_for(n), kT, 0, ST>>>(x, n, p, seed, draw_counter, y, mask);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, void* stream) {
    PPS_CHECK_ARG(dy && dx && mask && n >= 0, "pps_dropout_bwd: bad argument");
    if (n == 0) return PPS_OK;
    dropout_bwd_kernel<<<blocks_for(n), kT, 0, ST>>>(dy, mask, n, p, dx);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_rowscale_fwd(const float* x, const float* w, int64_t rows, int c, float* y, void* stream) {
    PPS_CHECK_ARG(x && w && y && rows >= 0 && c > 0, "pps_rowscale_fwd: bad argument");
    if (rows == 0) return PPS_OK;
    rowscale_fwd_kernel<<<blocks_for(rows * c), kT, 0, ST>>>(x, w, rows, c, y);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_rowscale_bwd(const float* x, const float* w, const float* dy, int64_t rows, int c, float* dx, float* dw, void* stream) {
    PPS_CHECK_ARG(x && w && dy && dx && dw && rows >= 0 && c > 0, "pps_rowscale_bwd: bad argument");
    if (rows == 0) return PPS_OK;
    rowscale_bwd_kernel<<<blocks_for(rows), kT, 0, ST>>>(x, w, dy, rows, c, dx, dw);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_concat_bcast_fwd(const float* x, const float* v, int64_t groups, int s, int c, float* out, void* stream) {
    PPS_CHECK_ARG(x && v && out && groups >= 0 && s > 0 && c > 0, "pps_concat_bcast_fwd: bad argument");
    if (groups == 0) return PPS_OK;
    concat_bcast_fwd_kernel<<<blocks_for(groups * s * 2 * c), kT, 0, ST>>>(x, v, groups, s, c, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_concat_bcast_bwd(const float* dout, int64_t groups, int s, int c, float* dx, float* dv, void* stream) {
    PPS_CHECK_ARG(dout && dx && dv && groups >= 0 && s > 0 && c > 0, "pps_concat_bcast_bwd: bad argument");
    if (groups == 0) return PPS_OK;
    concat_bcast_bwd_kernel<<<blocks_for(groups * c), kT, 0, ST>>>(dout, groups, s, c, dx, dv);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_gather_rows(const float* x, const int32_t* idx, int64_t m, int c, float* y, void* stream) {
    PPS_CHECK_ARG(x && idx && y && m >= 0 && c > 0, "pps_gather_rows: bad argument");
    if (m == 0) return PPS_OK;
    gather_rows_kernel<<<blocks_for(m * c), kT, 0, ST>>>(x, idx, m, c, y);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_scatter_add_rows(const float* dy, const int32_t* idx, int64_t m, int c, float* dx, void* stream) {
    PPS_CHECK_ARG(dy && idx && dx && m >= 0 && c > 0, "pps_scatter_add_rows: bad argument");
    if (m == 0) return PPS_OK;
    scatter_add_rows_kernel<<<blocks_for(m * c), kT, 0, ST>>>(dy, idx, m, c, dx);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_seg_max_fwd(const float* x, const float* w, int64_t groups, int s, int c, float* y, int32_t* arg, void* stream) {
    PPS_CHECK_ARG(x && y && arg && groups >= 0 && s > 0 && c > 0, "pps_seg_max_fwd: bad argument");
    if (groups == 0) return PPS_OK;
    seg_max_fwd_kernel<<<blocks_for(groups * c), kT, 0, ST>>>(x, w, groups, s, c, y, arg);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_seg_max_bwd(const float* dy, const int32_t* arg, const float* x, const float* w, int64_t groups, int s, int c, float* dx, float* dw,
                    void* stream) {
    PPS_CHECK_ARG(dy && arg && x && dx && groups >= 0 && s > 0 && c > 0 && (!dw || w), "pps_seg_max_bwd: bad argument");
    if (groups == 0) return PPS_OK;
    PPS_CUDA(cudaMemsetAsync(dx, 0, (size_t)groups * s * c * sizeof(float), ST));
    if (dw) PPS_CUDA(cudaMemsetAsync(dw, 0, (size_t)groups * s * sizeof(float), ST));
    seg_max_bwd_kernel<<<blocks_for(groups * c), kT, 0, ST>>>(dy, arg, x, w, groups, s, c, dx, dw);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_gather_max_fwd(const float* x, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int c, int kn, float* y, int32_t* arg,
                       void* stream) {
    PPS_CHECK_ARG(x && ids && y && arg && b >= 0 && n_in > 0 && n_s >= 0 && c > 0 && kn > 0, "pps_gather_max_fwd: bad argument");
    PPS_CHECK_ARG(b * n_in < (1ll << 31), "pps_gather_max_fwd: more than 2^31 source rows");
    if (b * n_s == 0) return PPS_OK;
    gather_max_fwd_kernel<<<blocks_for(b * n_s * c), kT, 0, ST>>>(x, ids, b, n_in, n_s, c, kn, y, arg);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_gather_max_bwd(const float* dy, const int32_t* arg, int64_t rows, int c, float* dx, void* stream) {
    PPS_CHECK_ARG(dy && arg && dx && rows >= 0 && c > 0, "pps_gather_max_bwd: bad argument");
    if (rows == 0) return PPS_OK;
    gather_max_bwd_kernel<<<blocks_for(rows * c), kT, 0, ST>>>(dy, arg, rows * c, c, dx);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_attn_pool_fwd(const float* scores, const float* v, int64_t groups, int s, int h, int c, float* prob, float* a, float* out,
                      void* stream) {
    PPS_CHECK_ARG(scores && v && prob && a && out && groups >= 0 && s > 0 && h > 0 && c > 0, "pps_attn_pool_fwd: bad argument");
    PPS_CHECK_ARG(((size_t)s * h + s + kT) * sizeof(float) <= 48 * 1024, "pps_attn_pool_fwd: a group of %d rows x %d heads exceeds the shared-memory tile", s, h);
    if (groups == 0) return PPS_OK;
    attn_pool_fwd_kernel<<<(unsigned)groups, kT, ((size_t)s * h + s + kT) * sizeof(float), ST>>>(scores, v, s, h, c, prob, a, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_attn_pool_bwd(const float* dout, const float* prob, const float* a, const float* v, int64_t groups, int s, int h, int c,
                      float* dscores, float* dv, void* stream) {
    PPS_CHECK_ARG(dout && prob && a && v && dscores && dv && groups >= 0 && s > 0 && s <= 4096 && h > 0 && c > 0,
                  "pps_attn_pool_bwd: bad argument");
    if (groups == 0) return PPS_OK;
    attn_pool_bwd_kernel<<<(unsigned)groups, kT, s * sizeof(float), ST>>>(dout, prob, a, v, s, h, c, dscores, dv);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_fka_geometry_fwd(const float* pts, const float* support, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int kn,
                         const float* alpha, const float* beta, float* norm_radius, float momentum, int update_radius, float* offs,
                         float* dist, float* sig, float* dw, double* scratch, void* stream) {
    PPS_CHECK_ARG(pts && support && ids && alpha && beta && norm_radius && offs && dist && sig && dw && scratch,
                  "pps_fka_geometry_fwd: null pointer");
    PPS_CHECK_ARG(b > 0 && n_in > 0 && n_s > 0 && kn > 0 && kn <= 16, "pps_fka_geometry_fwd: bad shape");
    const long long points = b * n_s;
    PPS_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), ST));
    fka_geometry_kernel<<<blocks_for(points), kT, 0, ST>>>(pts, support, ids, b, n_in, n_s, kn, offs, dist, scratch);
    PPS_LAUNCH_CHECK();
    if (update_radius) {
        fka_radius_update_kernel<<<1, 1, 0, ST>>>(scratch, points, momentum, norm_radius);
        PPS_LAUNCH_CHECK();
    }
    fka_weights_fwd_kernel<<<blocks_for(points), kT, 0, ST>>>(offs, dist, points, kn, alpha, beta, norm_radius, sig, dw);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_fka_weights_bwd(const float* ddw, const float* sig, const float* dist, int64_t points, int kn, double* dalpha_dbeta, void* stream) {
    PPS_CHECK_ARG(ddw && sig && dist && dalpha_dbeta && points > 0 && kn > 0, "pps_fka_weights_bwd: bad argument");
    PPS_CUDA(cudaMemsetAsync(dalpha_dbeta, 0, 2 * sizeof(double), ST));
    fka_weights_bwd_kernel<<<blocks_for(points), kT, 0, ST>>>(ddw, sig, dist, points, kn, dalpha_dbeta);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_fka_feat_fwd(const float* x, const int32_t* ids, const float* mat, int64_t b, int64_t n_in, int64_t n_s, int kn, int cin, float* feat,
                     void* stream) {
    PPS_CHECK_ARG(x && ids && mat && feat && b > 0 && n_in > 0 && n_s > 0 && kn > 0 && kn <= 16 && cin > 0, "pps_fka_feat_fwd: bad argument");
    fka_feat_fwd_kernel<<<(unsigned)(b * n_s), kT, 0, ST>>>(x, ids, mat, n_in, n_s, kn, cin, feat);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_fka_feat_bwd(const float* dfeat, const float* x, const int32_t* ids, const float* mat, int64_t b, int64_t n_in, int64_t n_s, int kn,
                     int cin, float* dx, float* dmat, void* stream) {
    PPS_CHECK_ARG(dfeat && x && ids && mat && dx && dmat && b > 0 && n_in > 0 && n_s > 0 && kn > 0 && kn <= 16 && cin > 0,
                  "pps_fka_feat_bwd: bad argument");
    PPS_CUDA(cudaMemsetAsync(dx, 0, (size_t)b * n_in * cin * sizeof(float), ST));
    fka_feat_bwd_kernel<<<(unsigned)(b * n_s), kT, 0, ST>>>(dfeat, x, ids, mat, n_in, n_s, kn, cin, dx, dmat);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_ce_fwd(const float* logits, const int64_t* target, int64_t m, int c, float* loss_rows, double* loss_sum, void* stream) {
    PPS_CHECK_ARG(logits && target && loss_rows && loss_sum && m > 0 && c > 0, "pps_ce_fwd: bad argument");
    PPS_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
    ce_fwd_kernel<<<blocks_for(m), kT, 0, ST>>>(logits, target, m, c, loss_rows, loss_sum);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_ce_bwd(const float* logits, const int64_t* target, const float* scale, int64_t m, int c, float* dlogits, void* stream) {
    PPS_CHECK_ARG(logits && target && scale && dlogits && m > 0 && c > 0, "pps_ce_bwd: bad argument");
    ce_bwd_kernel<<<blocks_for(m), kT, 0, ST>>>(logits, target, scale, m, c, dlogits);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
int pps_colsum(const float* x, int64_t rows, int c, int64_t ld, float* out, int accumulate, void* stream) {
    PPS_CHECK_ARG(x && out && rows >= 0 && c > 0 && ld >= c, "pps_colsum: bad argument");
    if (!accumulate) PPS_CUDA(cudaMemsetAsync(out, 0, (size_t)c * sizeof(float), ST));
    if (rows == 0) return PPS_OK;
    const long long slab = slab_for(rows, 1);
    if ((c & 3) == 0 && (ld & 3) == 0 && aligned16(x))
        colsum_v4_kernel<<<(unsigned)ceil_div(rows, slab), kT, 0, ST>>>(x, rows, c, ld, slab, out);
    else
        colsum_kernel<<<(unsigned)ceil_div(rows, slab), kT, 0, ST>>>(x, rows, c, ld, slab, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // extern "C"
