// Spherical-clustering leaflets on the device (spherical_clustering.rs:36-275).
//
// EXPERIMENTAL: written after the last GPU session of round 1 and not yet run on a device.  gorder_gpu_create refuses
// GORDER_LEAFLET_SPHERICAL unless GORDER_EXPERIMENTAL_SPHERICAL is set; tests/test_gpu_spherical.py is skipped without it.
//
// Per assignment frame, one CTA of 1024 threads:
//   distances of the ClusterHeads group (GorderSetup.membrane) from its PBC-aware centre (run_group_center, 3 axes),
//   initial means = the 25th / 75th percentile (exact order statistics by a 4-pass radix select on the float bits),
//   initial variances = the sample variance, then <= 50 EM iterations of a two-component 1-D Gaussian mixture
//   (E step + log-likelihood, means, variances: three block reductions per iteration), r < 0.5 -> cluster 1,
//   the cluster farther from the centre is the upper (outer) leaflet.
// The reference folds its sums sequentially in f32; the block reductions here run in f64.  The responsibilities
// therefore agree to ~1e-6, not in bits: an assignment can only differ for a head with r within that of 0.5, i.e.
// one that sits between the two leaflets of the vesicle.
#pragma once
#include "gorder_kernels.cuh"

namespace gorder {

constexpr int kGmmMaxIterations = 50;      // spherical_clustering.rs:23
constexpr float kGmmTolerance = 1e-4f;     // spherical_clustering.rs:26
constexpr int kSphThreads = 1024;

// sums of up to four doubles over the CTA, result in every thread
__device__ __forceinline__ void block_sum4(double (&x)[4], double (*s_red)[32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
    __syncthreads();   // s_red may still be read from the previous call
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 4; k++) s_red[k][warp] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double t = 0.0;
        for (int w = 0; w < kSphThreads / 32; w++) t += s_red[k][w];   // fixed order: deterministic
        x[k] = t;
    }
}

// k-th smallest (0-based) of n non-negative floats: radix select over the bit patterns, most significant byte first
__device__ __forceinline__ float block_select(const float *__restrict__ d, int n, int k, unsigned *s_hist, int *s_pick) {
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

__device__ __forceinline__ float gmm_log_gaussian(float x, float mean, float variance) {   // spherical_clustering.rs:103-107
    const float diff = x - mean;
    return -0.5f * (logf(2.0f * CUDART_PI_F) + logf(variance) + __fdiv_rn(__fmul_rn(diff, diff), variance));
}

// grid (n_assign), block kSphThreads.  dist / resp: [n_assign][n] scratch; cl_upper: [n_assign][n] result (1 = upper).
__global__ void __launch_bounds__(kSphThreads) spherical_cluster_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                                       const int *__restrict__ frame_list, const float *__restrict__ center,
                                                                       float *__restrict__ dist_all, float *__restrict__ resp_all,
                                                                       unsigned char *__restrict__ cl_upper) {
    __shared__ double s_red[4][32];
    __shared__ unsigned s_hist[256];
    __shared__ int s_pick[2];
    const int ai = blockIdx.x, f = frame_list[ai], n = v.membrane.n;
    const FrameAux &a = aux[f];
    const bool pbc = v.handle_pbc != 0;
    const float cx = center[3 * ai], cy = center[3 * ai + 1], cz = center[3 * ai + 2];
    if ((cx != cx || cy != cy || cz != cz) && threadIdx.x == 0) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, a.frame_index);
    const float *fr = planes + (size_t)f * v.frame_floats;
    float *dist = dist_all + (size_t)ai * n, *resp = resp_all + (size_t)ai * n;
    const float n_f = (float)n;

    // distances from the centre (PBCHandler::distance, pbc.rs:354; Dimension::XYZ) and their mean
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int off = v.membrane.off[i];
        const size_t cs = (size_t)v.membrane.cs[i];
        float dx = __fsub_rn(fr[off], cx), dy = __fsub_rn(fr[off + cs], cy), dz = __fsub_rn(fr[off + 2 * cs], cz);
        if (pbc) { dx = min_image(dx, a.L[0], a.half[0]); dy = min_image(dy, a.L[1], a.half[1]); dz = min_image(dz, a.L[2], a.half[2]); }
        const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        dist[i] = d;
        acc[0] += (double)d;
    }
    block_sum4(acc, s_red);   // (also orders the writes of dist before the reads below)
    const float gmean = (float)acc[0] / n_f;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float t = dist[i] - gmean; acc[0] += (double)(t * t); }
    block_sum4(acc, s_red);
    float gvar = (float)acc[0] / (n_f - 1.0f);
    if (!isfinite(gvar) || gvar <= 0.0f) gvar = 1.0f;

    // initialize_params (spherical_clustering.rs:116-136)
    const float var_floor = 1e-6f, weight_floor = 1e-4f;
    float mean_a = block_select(dist, n, n / 4, s_hist, s_pick);
    float mean_b = block_select(dist, n, (3 * n) / 4, s_hist, s_pick);
    float weight_a = 0.5f, var_a = fmaxf(gvar, var_floor), var_b = var_a;
    float prev_avg_ll = -CUDART_INF_F;
    for (int i = threadIdx.x; i < n; i += blockDim.x) resp[i] = 0.5f;

    // fit_gmm_1d_two_components (spherical_clustering.rs:138-237)
    for (int it = 0; it < kGmmMaxIterations; it++) {
        const float lwa = logf(weight_a), lwb = logf(1.0f - weight_a);
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i];
            const float ja = lwa + gmm_log_gaussian(x, mean_a, var_a), jb = lwb + gmm_log_gaussian(x, mean_b, var_b);
            const float m = fmaxf(ja, jb);
            const float log_px = m + logf(expf(ja - m) + expf(jb - m));
            const float r = expf(ja - log_px);
            resp[i] = r;
            acc[0] += (double)log_px; acc[1] += (double)r;
        }
        block_sum4(acc, s_red);
        const float avg_ll = (float)acc[0] / n_f;
        if (fabsf(avg_ll - prev_avg_ll) < kGmmTolerance) break;
        prev_avg_ll = avg_ll;
        float sum_a = (float)acc[1], sum_b = n_f - sum_a;
        sum_a = fmaxf(sum_a, 1e-6f); sum_b = fmaxf(sum_b, 1e-6f);
        weight_a = fminf(fmaxf(sum_a / n_f, weight_floor), 1.0f - weight_floor);
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i], r = resp[i];
            acc[0] += (double)(r * x); acc[1] += (double)((1.0f - r) * x);
        }
        block_sum4(acc, s_red);
        mean_a = (float)acc[0] / sum_a; mean_b = (float)acc[1] / sum_b;
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = dist[i], r = resp[i], da = x - mean_a, db = x - mean_b;
            acc[0] += (double)(r * da * da); acc[1] += (double)((1.0f - r) * db * db);
        }
        block_sum4(acc, s_red);
        var_a = fmaxf((float)acc[0] / sum_a, var_floor); var_b = fmaxf((float)acc[1] / sum_b, var_floor);
    }

    // Clusters::from_responsibilities (spherical_clustering.rs:239-272)
    acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (resp[i] < 0.5f) { acc[0] += 1.0; acc[1] += (double)dist[i]; } else { acc[2] += 1.0; acc[3] += (double)dist[i]; }
    }
    block_sum4(acc, s_red);
    const bool first_is_upper = ((float)acc[1] / (float)acc[0]) > ((float)acc[3] / (float)acc[2]);   // NaN (empty cluster) compares false
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
