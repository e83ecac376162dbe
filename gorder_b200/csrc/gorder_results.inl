// Results conversion: the step AFTER the hot path, host only -- no GPU, no handle.
//
// Turns the integer accumulators gorder_gpu_finish() returns into the numbers the reference prints:
//   * AnalysisOrder::calc_order            order.rs:97-107 (integer division truncating toward zero, order.rs:34-41)
//   * OrderSummer                          converter.rs:513-559 (element-wise sums over the bonds of an atom /
//                                          the atoms of a molecule / the molecules of the system)
//   * TimeWiseData::estimate_error         timewise.rs:191-231 (block means, sample standard deviation in f32 with the
//                                          sequential sums of `statistical 1.0.0`: mean = fold(+) / n, var = fold(+ d^2) / (n - 1))
//   * TimeWiseData::prefix_average         timewise.rs:259-274 (convergence)
//   * order-map division                   converter.rs:159-308 / ordermap.rs (value / samples, NaN below min_samples)
// Included by gorder_capi.cu; declared in include/gorder_b200.h.

namespace gres {

constexpr double kPrecision = 1e6;   // order.rs:13

inline float calc_order(long long total, unsigned long long n, long long min_samples) {
    if (n < (unsigned long long)std::max<long long>(1, min_samples)) return std::numeric_limits<float>::quiet_NaN();
    // |total| / n with the sign restored: the quotient truncates toward zero for any magnitude of n
    const unsigned long long mag = total < 0 ? 0ull - (unsigned long long)total : (unsigned long long)total;
    const unsigned long long q = mag / n;
    const double v = (double)q / kPrecision;
    return (float)(total < 0 ? -v : v);
}

inline bool valid(const GorderRaw *r) {
    return r && r->n_slots >= 0 && r->n_frames >= 0 && (r->n_slots == 0 || (r->sum && r->count)) &&
           (r->n_frames == 0 || r->n_slots == 0 || (r->tw_sum && r->tw_count) || (!r->tw_sum && !r->tw_count));
}

inline bool slots_ok(const GorderRaw *r, const int32_t *slots, int32_t n_sel) {
    if (n_sel < 0 || (n_sel > 0 && !slots)) return false;
    for (int32_t i = 0; i < n_sel; i++) if (slots[i] < 0 || slots[i] >= r->n_slots) return false;
    return true;
}

}  // namespace gres

extern "C" {

int gorder_results_order(const GorderRaw *raw, const int32_t *slots, int32_t n_sel, int32_t n_blocks, int32_t min_samples, float sign,
                         float *value, float *error) {
    using namespace gres;
    if (!valid(raw) || !slots_ok(raw, slots, n_sel) || !value) return GORDER_ERR_INVALID_ARGUMENT;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    long long tot[3] = {0, 0, 0};
    unsigned long long cnt[3] = {0, 0, 0};
    for (int32_t i = 0; i < n_sel; i++)
        for (int k = 0; k < 3; k++) { tot[k] += raw->sum[3 * (size_t)slots[i] + k]; cnt[k] += raw->count[3 * (size_t)slots[i] + k]; }
    for (int k = 0; k < 3; k++) {
        const float v = calc_order(tot[k], cnt[k], min_samples);
        value[k] = v == v ? sign * v : v;
    }
    if (!error) return GORDER_OK;
    const bool tw = raw->tw_sum && raw->n_frames > 0 && n_blocks > 0;
    if (!tw) { error[0] = error[1] = error[2] = nan; return GORDER_OK; }
    // the reference refuses fewer than two blocks (timewise.rs:196-199) and divides by block_size = n_frames / n_blocks
    // (:201-207: a block size of zero is a division by zero there)
    if (n_blocks < 2 || (long long)n_blocks > raw->n_frames) return GORDER_ERR_INVALID_ARGUMENT;
    const long long block = raw->n_frames / n_blocks;
    const size_t row = 3 * (size_t)raw->n_slots;
    std::vector<float> means(3 * (size_t)n_blocks);
    bool empty[3] = {false, false, false};
    for (int32_t b = 0; b < n_blocks; b++) {
        long long s[3] = {0, 0, 0};
        unsigned long long c[3] = {0, 0, 0};
        for (long long f = b * block; f < (b + 1) * block; f++)
            for (int32_t i = 0; i < n_sel; i++) {
                const size_t at = (size_t)f * row + 3 * (size_t)slots[i];
                for (int k = 0; k < 3; k++) { s[k] += raw->tw_sum[at + k]; c[k] += raw->tw_count[at + k]; }
            }
        for (int k = 0; k < 3; k++) {
            if (c[k] == 0) empty[k] = true;
            means[3 * (size_t)b + k] = calc_order(s[k], c[k], 1);
        }
    }
    for (int k = 0; k < 3; k++) {
        if (empty[k] || cnt[k] < (unsigned long long)std::max(0, min_samples)) { error[k] = nan; continue; }
        float sum = 0.0f;
        for (int32_t b = 0; b < n_blocks; b++) sum += means[3 * (size_t)b + k];
        const float mean = sum / (float)n_blocks;
        float dev2 = 0.0f;
        for (int32_t b = 0; b < n_blocks; b++) { const float d = means[3 * (size_t)b + k] - mean; dev2 += d * d; }
        error[k] = sqrtf(dev2 / (float)(n_blocks - 1));
    }
    return GORDER_OK;
}

int gorder_results_convergence(const GorderRaw *raw, const int32_t *slots, int32_t n_sel, float sign, float *out) {
    using namespace gres;
    if (!valid(raw) || !slots_ok(raw, slots, n_sel) || (raw->n_frames > 0 && (!out || !raw->tw_sum))) return GORDER_ERR_INVALID_ARGUMENT;
    long long s[3] = {0, 0, 0};
    unsigned long long c[3] = {0, 0, 0};
    const size_t row = 3 * (size_t)raw->n_slots;
    for (long long f = 0; f < raw->n_frames; f++) {
        for (int32_t i = 0; i < n_sel; i++) {
            const size_t at = (size_t)f * row + 3 * (size_t)slots[i];
            for (int k = 0; k < 3; k++) { s[k] += raw->tw_sum[at + k]; c[k] += raw->tw_count[at + k]; }
        }
        for (int k = 0; k < 3; k++) {
            const float v = calc_order(s[k], c[k], 1);
            out[3 * (size_t)f + k] = v == v ? sign * v : v;
        }
    }
    return GORDER_OK;
}

int gorder_results_map(const int64_t *map_sum, const uint64_t *map_count, int64_t n, int32_t min_samples, float sign, float *out) {
    if (n < 0 || (n > 0 && (!map_sum || !map_count || !out))) return GORDER_ERR_INVALID_ARGUMENT;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    const unsigned long long need = (unsigned long long)std::max(0, min_samples);
    for (int64_t i = 0; i < n; i++) {
        const unsigned long long c = map_count[i];
        const float v = (float)((double)map_sum[i] / gres::kPrecision) / (float)c;   // 0 / 0 = NaN for an empty bin
        out[i] = c < need ? nan : sign * v;
    }
    return GORDER_OK;
}

}  // extern "C"
