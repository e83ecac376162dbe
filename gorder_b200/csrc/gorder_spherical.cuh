// Spherical-clustering leaflets on the device (spherical_clustering.rs:36-275).
// Parity: tests/test_gpu_spherical.py (device vs the oracle, which is pinned by spherical_clustering.rs:299-361).
//
// Per assignment frame, one CTA:
//   distances of the ClusterHeads group (GorderSetup.membrane) from its PBC-aware centre (run_group_center, 3 axes),
//   initial means = the 25th / 75th percentile (exact order statistics by a 4-pass radix select on the float bits),
//   initial variances = the sample variance, then <= 50 EM iterations of a two-component 1-D Gaussian mixture,
//   r < 0.5 -> cluster 1, the cluster farther from the centre is the upper (outer) leaflet.
//
// The reference folds every sum SEQUENTIALLY in f32, and over 1e5 heads such a fold carries a relative error of ~1e-3
// that decides at which iteration |d avg_ll| < 1e-4 stops the fit: a tree or f64 reduction lands on a different
// iteration and, for leaflets whose distance distributions overlap (gap < ~5 sigma), on different assignments (measured
// with a numpy model of both variants against the oracle: up to 98 of 85 027 heads; with sequential folds: none, max
// |dr| 4e-5 with +-1 ulp of noise on every log / exp).  So the element-wise work of a phase runs on all threads and leaves
// its terms in scratch arrays, and each fold is done by ONE lane in the reference's order -- two or three folds of a
// phase on lanes of different warps at once.  A fold is bound by the latency of the dependent FADDs (~0.2 ms for 1e5
// heads), a CTA therefore keeps its SM almost idle: throughput comes from the frames of a batch running side by side
// (256 threads per CTA, several CTAs per SM).
// Arithmetic: the oracle's operation order with explicit _rn intrinsics (no FMA contraction); logf / expf are CUDA's
// (<= 2 ulp from libm's).
#pragma once
#include "gorder_kernels.cuh"

namespace gorder {

constexpr int kGmmMaxIterations = 50;      // spherical_clustering.rs:23
constexpr float kGmmTolerance = 1e-4f;     // spherical_clustering.rs:26
constexpr int kSphThreads = 256;
constexpr int kSphArrays = 5;              // dist, resp, and three term arrays per frame

// sum of a[0..n) in index order, a 16-byte aligned; the array was written by this CTA (plain loads)
__device__ __forceinline__ float seq_sum(const float *a, int n) {
    float s = 0.0f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        const float4 q = *reinterpret_cast<const float4 *>(a + i);
        s = __fadd_rn(s, q.x); s = __fadd_rn(s, q.y); s = __fadd_rn(s, q.z); s = __fadd_rn(s, q.w);
    }
    for (; i < n; i++) s = __fadd_rn(s, a[i]);
    return s;
}
// folds of up to three arrays at once (lane 0 of warps 0..2); every thread returns with s_out[] valid
__device__ __forceinline__ void seq_sums(const float *a0, const float *a1, const float *a2, int n, float *s_out) {
    __syncthreads();   // the arrays were written by all threads; s_out may still be read from the previous call
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && warp < 3) {
        const float *a = warp == 0 ? a0 : (warp == 1 ? a1 : a2);
        if (a) s_out[warp] = seq_sum(a, n);
    }
    __syncthreads();
}

// k-th smallest (0-based) of n non-negative floats: radix select over the bit patterns, most significant byte first
__device__ __forceinline__ float block_select(const float *d, int n, int k, unsigned *s_hist, int *s_pick) {
    unsigned prefix = 0u, mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned b = __float_as_uint(d[i]);
            if ((b & mask) == prefix) atomicAdd(&s_hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, dg = 0;
            for (; dg < 255; dg++) {
                if (acc + (int)s_hist[dg] > k) break;
                acc += (int)s_hist[dg];
            }
            s_pick[0] = dg; s_pick[1] = k - acc;
        }
        __syncthreads();
        prefix |= (unsigned)s_pick[0] << shift; mask |= 255u << shift; k = s_pick[1];
        __syncthreads();
    }
    return __uint_as_float(prefix);
}

// log_gaussian (spherical_clustering.rs:103-107) with ln(2 pi) and ln(variance) hoisted
__device__ __forceinline__ float gmm_log_gaussian(float x, float mean, float variance, float ln_2pi, float ln_var) {
    const float diff = __fsub_rn(x, mean);
    return __fmul_rn(-0.5f, __fadd_rn(__fadd_rn(ln_2pi, ln_var), __fdiv_rn(__fmul_rn(diff, diff), variance)));
}

// grid (n_assign), block kSphThreads.  scratch: [n_assign][kSphArrays][n_pad] floats (n_pad = n rounded up to 4);
// cl_upper: [n_assign][n] result (1 = upper).
__global__ void __launch_bounds__(kSphThreads) spherical_cluster_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                                       const int *__restrict__ frame_list, const float *__restrict__ center,
                                                                       float *scratch, int n_pad, unsigned char *__restrict__ cl_upper) {
    __shared__ float s_sum[3];
    __shared__ unsigned s_hist[256];
    __shared__ int s_pick[2];
    const int ai = blockIdx.x, f = frame_list[ai], n = v.membrane.n;
    const FrameAux &a = aux[f];
    const bool pbc = v.handle_pbc != 0;
    const float cx = center[3 * ai], cy = center[3 * ai + 1], cz = center[3 * ai + 2];
    if ((cx != cx || cy != cy || cz != cz) && threadIdx.x == 0) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, a.frame_index);
    const float *fr = planes + (size_t)f * v.frame_floats;
    float *dist = scratch + (size_t)ai * kSphArrays * n_pad, *resp = dist + n_pad, *t0 = resp + n_pad, *t1 = t0 + n_pad, *t2 = t1 + n_pad;
    const float n_f = (float)n;

    // distances from the centre (PBCHandler::distance, pbc.rs:354; Dimension::XYZ): atom - centre, minimum image
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int off = v.membrane.off[i];
        const size_t cs = (size_t)v.membrane.cs[i];
        float dx = __fsub_rn(fr[off], cx), dy = __fsub_rn(fr[off + cs], cy), dz = __fsub_rn(fr[off + 2 * cs], cz);
        if (pbc) { dx = min_image(dx, a.L[0], a.half[0]); dy = min_image(dy, a.L[1], a.half[1]); dz = min_image(dz, a.L[2], a.half[2]); }
        dist[i] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    // initialize_params (spherical_clustering.rs:116-136): sample mean and variance, quartiles
    seq_sums(dist, nullptr, nullptr, n, s_sum);
    const float gmean = __fdiv_rn(s_sum[0], n_f);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float t = __fsub_rn(dist[i], gmean); t0[i] = __fmul_rn(t, t); }
    seq_sums(t0, nullptr, nullptr, n, s_sum);
    float gvar = __fdiv_rn(s_sum[0], __fsub_rn(n_f, 1.0f));
    if (!isfinite(gvar) || gvar <= 0.0f) gvar = 1.0f;
    const float var_floor = 1e-6f, weight_floor = 1e-4f;
    float mean_a = block_select(dist, n, n / 4, s_hist, s_pick);
    float mean_b = block_select(dist, n, (3 * n) / 4, s_hist, s_pick);
    float weight_a = 0.5f, var_a = fmaxf(gvar, var_floor), var_b = var_a;
    float prev_avg_ll = -CUDART_INF_F;
    for (int i = threadIdx.x; i < n; i += blockDim.x) resp[i] = 0.5f;
    const float ln_2pi = logf(__fmul_rn(2.0f, CUDART_PI_F));

    // fit_gmm_1d_two_components (spherical_clustering.rs:138-237)
    for (int it = 0; it < kGmmMaxIterations; it++) {
        const float lwa = logf(weight_a), lwb = logf(__fsub_rn(1.0f, weight_a)), lva = logf(var_a), lvb = logf(var_b);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i];
            const float ja = __fadd_rn(lwa, gmm_log_gaussian(x, mean_a, var_a, ln_2pi, lva));
            const float jb = __fadd_rn(lwb, gmm_log_gaussian(x, mean_b, var_b, ln_2pi, lvb));
            const float m = fmaxf(ja, jb);
            const float log_px = __fadd_rn(m, logf(__fadd_rn(expf(__fsub_rn(ja, m)), expf(__fsub_rn(jb, m)))));
            t0[i] = log_px;
            resp[i] = expf(__fsub_rn(ja, log_px));
        }
        seq_sums(t0, nullptr, nullptr, n, s_sum);                 // loglik_sum
        const float avg_ll = __fdiv_rn(s_sum[0], n_f);
        if (fabsf(__fsub_rn(avg_ll, prev_avg_ll)) < kGmmTolerance) break;   // (same value in every thread)
        prev_avg_ll = avg_ll;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i], r = resp[i];
            t1[i] = __fmul_rn(r, x); t2[i] = __fmul_rn(__fsub_rn(1.0f, r), x);
        }
        seq_sums(resp, t1, t2, n, s_sum);                         // sum of the responsibilities, numerators of the means
        float sum_a = s_sum[0], sum_b = __fsub_rn(n_f, sum_a);
        sum_a = fmaxf(sum_a, 1e-6f); sum_b = fmaxf(sum_b, 1e-6f);
        weight_a = fminf(fmaxf(__fdiv_rn(sum_a, n_f), weight_floor), __fsub_rn(1.0f, weight_floor));
        mean_a = __fdiv_rn(s_sum[1], sum_a); mean_b = __fdiv_rn(s_sum[2], sum_b);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i], r = resp[i], da = __fsub_rn(x, mean_a), db = __fsub_rn(x, mean_b);
            t1[i] = __fmul_rn(__fmul_rn(r, da), da); t2[i] = __fmul_rn(__fmul_rn(__fsub_rn(1.0f, r), db), db);
        }
        seq_sums(t1, t2, nullptr, n, s_sum);
        var_a = fmaxf(__fdiv_rn(s_sum[0], sum_a), var_floor); var_b = fmaxf(__fdiv_rn(s_sum[1], sum_b), var_floor);
    }

    // Clusters::from_responsibilities (spherical_clustering.rs:239-272): adding 0 leaves a sequential f32 sum unchanged
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool first = resp[i] < 0.5f;
        t0[i] = first ? 1.0f : 0.0f; t1[i] = first ? dist[i] : 0.0f; t2[i] = first ? 0.0f : dist[i];
    }
    seq_sums(t0, t1, t2, n, s_sum);   // (the count is exact in f32 below 2^24 heads)
    const float n1 = s_sum[0], n2 = __fsub_rn(n_f, n1);
    const bool first_is_upper = __fdiv_rn(s_sum[1], n1) > __fdiv_rn(s_sum[2], n2);   // NaN (empty cluster) compares false
    for (int i = threadIdx.x; i < n; i += blockDim.x) cl_upper[(size_t)ai * n + i] = ((resp[i] < 0.5f) == first_is_upper) ? 1 : 0;
}

// leaflet rows from the cluster table (SphericalClusterClassification::set_assignment, leaflets.rs:1351-1367, + maybe_flip):
// cl_index[molpad] = position of the molecule's head in the ClusterHeads group.  grid ((n_molpad + 255) / 256, n_assign)
__global__ void __launch_bounds__(256) spherical_assign_kernel(DeviceView v, const int *__restrict__ molpad_type, const int *__restrict__ cl_index,
                                                               const unsigned char *__restrict__ cl_upper, unsigned char *__restrict__ rows) {
    const int ai = blockIdx.y, mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= v.n_molpad) return;
    const int t = molpad_type[mp];
    if (t < 0) return;
    const TypeDesc &td = v.types[t];
    const int m = mp - __ldg(&td.molpad0);
    unsigned char out = GORDER_UPPER;
    if (m < __ldg(&td.n_mol)) {
        bool upper = cl_upper[(size_t)ai * v.membrane.n + cl_index[mp]] != 0;
        if (v.leaflet_flip) upper = !upper;
        out = upper ? GORDER_UPPER : GORDER_LOWER;
    }
    rows[(size_t)(1 + ai) * v.n_molpad + mp] = out;
}

}  // namespace gorder
