// Several Metropolis-Hastings sweeps per call, for every likelihood model (north_star: "several MH sweeps fused per
// launch"; reference loop SMC_example/Micmem_SMC_main.py:209-249).
//
// One sweep is  proposal (factor read from device memory) -> early-rejection thresholds -> likelihood of the in-box
// proposals -> accept -> merged moments of the updated particles (which also gathers the sweep's counters over all
// shards and rebuilds the proposal factor on the device).  Because the factor never leaves the device, nothing in
// that chain needs the host: smcb_mh_sweeps enqueues n_sweeps of them back to back.  Unlike smcb_mh_fused (kinetic
// model, frozen factor) the covariance IS refreshed every sweep, as the reference does (Micmem_SMC_main.py:212).
// What a batch cannot do is take the reference's two per-sweep host decisions - stop when enough particles have
// moved (:243), halve the step when too few have (:247) - inside the batch: the caller applies them between
// batches (with n_sweeps = 1 that is the reference's rule exactly).
#include "common.cuh"

extern "C" int smcb_mh_sweeps(smcb_handle* h, int model, double* theta_dev, int64_t ld, double* lk_dev, int64_t n, int d,
                              int64_t n_total, const double* w_cov_host, double ratio, const double* low_host,
                              const double* high_host, double gamma, int n_sweeps, int early_reject, uint64_t seed,
                              uint64_t id_offset, uint32_t stage, uint32_t sweep0, double* prop_dev, int64_t ld_prop,
                              double* lk2_dev, double* lkmin_dev, uint8_t* inbox_dev, uint8_t* moved_dev,
                              int64_t* counts_dev, double* blk_dev, void* stream) {
    REQUIRE(h, h && theta_dev && lk_dev && low_host && high_host && prop_dev && lk2_dev && inbox_dev && moved_dev &&
                   counts_dev && blk_dev,
            SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && ld_prop >= n && n_total >= n && n_sweeps >= 1,
            SMCB_ERR_INVALID, "bad size");
    REQUIRE(h, !early_reject || lkmin_dev != nullptr, SMCB_ERR_INVALID, "early rejection needs a threshold buffer");
    const double* F_dev = blk_dev + 4 + d + (size_t)d * d;      // the factor smcb_moments_merged left in the block
    int rc;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint32_t sweep = sweep0 + (uint32_t)s;
        if ((rc = smcb_mh_propose_dev(h, theta_dev, ld, n, d, F_dev, ratio, low_host, high_host, nullptr, seed, id_offset,
                                      stage, sweep, prop_dev, ld_prop, inbox_dev, stream)))
            return rc;
        const double* lkmin = nullptr;
        if (early_reject) {
            if ((rc = smcb_mh_threshold(h, lk_dev, inbox_dev, n, gamma, nullptr, nullptr, seed, id_offset, stage, sweep,
                                        lkmin_dev, stream)))
                return rc;
            lkmin = lkmin_dev;
        }
        if ((rc = smcb_loglik_bounded(h, model, prop_dev, ld_prop, n, d, inbox_dev, lkmin, lk2_dev, stream))) return rc;
        if ((rc = smcb_mh_accept(h, theta_dev, ld, lk_dev, prop_dev, ld_prop, lk2_dev, inbox_dev, n, d, gamma, nullptr,
                                 nullptr, seed, id_offset, stage, sweep, moved_dev, counts_dev, stream)))
            return rc;
        if ((rc = smcb_moments_merged(h, theta_dev, ld, n, d, n_total, counts_dev, w_cov_host, blk_dev, stream))) return rc;
    }
    return SMCB_OK;
}
