// se_api_loss.cu -- multi-resolution STFT loss entry points.
#include "se_host.h"

#include <cstdlib>

using namespace se;

// Default: only |B| is saved by the forward pass and the backward pass re-transforms the estimate.
// SE_MRSTFT_SAVE_SPECTRUM=1 (read once per process) makes the forward pass also save the estimate's spectrum
// (8 bytes per bin per resolution, 3x the workspace) so that the backward pass runs one transform per resolution
// instead of two -- a measured wash on B200 (see k_loss_bwd_saved), kept as an alternative for compute-poorer parts.
static bool loss_recompute() {
    static const bool v = [] { const char* e = std::getenv("SE_MRSTFT_SAVE_SPECTRUM"); return !(e && e[0] == '1'); }();
    return v;
}

static const int kRes[3][3] = {{512, 128, 512}, {1024, 256, 1024}, {2048, 512, 2048}};

#define SE_DISPATCH_LOSS_GEO(n_fft, CALL)                                              \
    do {                                                                               \
        if (frames8() && n_fft == 512) { using G = Geo<512, 128, 128, 8>; CALL; }      \
        else if (frames8() && n_fft == 1024) { using G = Geo<1024, 256, 128, 8>; CALL; } \
        else if (n_fft == 512) { using G = Geo<512, 128, 256>; CALL; }                 \
        else if (n_fft == 1024) { using G = Geo<1024, 256, 256>; CALL; }               \
        else { using G = Geo<2048, 512, 512>; CALL; }                                  \
    } while (0)
// forward at n = 2048: 8-frame groups halve the working set (64 KB) so that two 256-thread CTAs fit per SM
#define SE_DISPATCH_LOSS_FWD_GEO(n_fft, CALL)                                          \
    do {                                                                               \
        if (frames8() && n_fft == 512) { using G = Geo<512, 128, 128, 8>; CALL; }      \
        else if (frames8() && n_fft == 1024) { using G = Geo<1024, 256, 128, 8>; CALL; } \
        else if (n_fft == 512) { using G = Geo<512, 128, 256>; CALL; }                 \
        else if (n_fft == 1024) { using G = Geo<1024, 256, 256>; CALL; }               \
        else { using G = Geo<2048, 512, 256, 8>; CALL; }                               \
    } while (0)
static int loss_fwd_frames(int n_fft) { return (engine_version() == 2 || n_fft >= 2048 || frames8()) ? 8 : 16; }

// signal-pair engine (se_kernels2.cuh): forward pairs (reference, estimate) of a row, backward pairs two rows.
// The backward kernels also hold the right-edge zone in shared memory: tables go undoubled there so that
// two CTAs (n <= 1024) / one CTA (n = 2048) still fit.
#define SE_DISPATCH_LOSS_GEO2(n_fft, DUP, CALL)                                        \
    do {                                                                               \
        if (n_fft == 512) { using G = Geo2<512, 128, 128, DUP>; CALL; }                \
        else if (n_fft == 1024) { using G = Geo2<1024, 256, 256, DUP>; CALL; }         \
        else { using G = Geo2<2048, 512, 512, DUP>; CALL; }                            \
    } while (0)
template <class G>
static cudaError_t run_loss_fwd2(const LossArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_loss_fwd2<G>, (unsigned)(rows * a.nchunks), G::NT, Smem2<G>::ANALYSIS, st, a);
}
template <class G>
static cudaError_t run_loss_bwd2(LossArgs a, int64_t rows, cudaStream_t st) {
    const int64_t prows = (rows + 1) / 2;
    a.nchunks = plan_synthesis(prows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR, 2, 16);
    return launch(k_loss_bwd2<G>, (unsigned)(prows * a.nchunks), G::NT, Smem2<G>::FUSED_ADJ, st, a, (int)rows);
}

template <class G>
static cudaError_t run_loss_fwd(const LossArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_loss_fwd<G>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, st, a);
}
template <class G>
static cudaError_t run_loss_bwd(LossArgs a, int64_t rows, cudaStream_t st) {
    // n = 2048: single-group chunks (16 - (OLA-1) blocks each, no carry); 8-frame groups need >= 2 groups per chunk
    // so that the reflect fold's sources and destinations share a chunk
    a.nchunks = G::N >= 2048 ? (a.b_hi + 12) / 13 : plan_synthesis(rows, a.b_hi, G::OLA, G::MINB, G::FR, G::FR == 8 ? 2 : 1, G::FR == 8 ? 16 : 8);
    return launch(k_loss_bwd<G>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::FUSED_ADJ, st, a);
}
template <class G>
static cudaError_t run_loss_bwd_saved(LossArgs a, int64_t rows, cudaStream_t st) {
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR, G::FR == 8 ? 2 : 1, G::FR == 8 ? 16 : 8);
    return launch(k_loss_bwd_saved<G>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::SYNTH_ADJ, st, a);
}

extern "C" {

// ---------------------------------------------------------------- MR-STFT loss
static int loss_fwd_plan(int64_t rows, int64_t nsample, int r, int& gpc, int& nchunks) {
    const int64_t T = 1 + nsample / kRes[r][1];
    plan_analysis(rows, T, gpc, nchunks, loss_fwd_frames(kRes[r][0]));
    return (int)(rows * nchunks);
}

// workspace layout: [per-CTA partial sums (double) for the 3 resolutions | |B| per resolution (float) |
//                    A per resolution (float2, absent in recompute mode)]
static int64_t loss_partials_bytes(int64_t rows, int64_t nsample) {
    int64_t total = 0;
    for (int r = 0; r < 3; ++r) {
        int gpc, nchunks;
        total += (int64_t)loss_fwd_plan(rows, nsample, r, gpc, nchunks) * 3 * sizeof(double);
    }
    return (total + 255) / 256 * 256;
}
static int64_t loss_refmag_floats(int64_t rows, int64_t nsample, int r) {
    return rows * (kRes[r][0] / 2 + 1) * (1 + nsample / kRes[r][1]);
}

// bins of resolutions 0..r-1 (r = 3: of all three)
static int64_t loss_bins_before(int64_t rows, int64_t nsample, int r) {
    int64_t total = 0;
    for (int q = 0; q < r; ++q) total += loss_refmag_floats(rows, nsample, q);
    return total;
}
// the |B| region, padded so that the float2 region behind it stays 16-byte aligned
static int64_t loss_refmag_region_floats(int64_t rows, int64_t nsample) {
    return (loss_bins_before(rows, nsample, 3) + 3) / 4 * 4;
}

int64_t se_mrstft_workspace_bytes(int64_t rows, int64_t nsample) {
    int64_t total = loss_partials_bytes(rows, nsample);
    total += loss_refmag_region_floats(rows, nsample) * (int64_t)sizeof(float);
    if (!loss_recompute()) total += loss_bins_before(rows, nsample, 3) * (int64_t)sizeof(float2);
    return total;
}

static int loss_fwd_impl(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums, float* loss,
                         void* workspace, void* stream);

int se_mrstft_loss_fwd(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums,
                       void* workspace, void* stream) {
    return loss_fwd_impl(est, ref, rows, nsample, sums, nullptr, workspace, stream);
}
// single-process variant: the reduction launch also writes the loss value (no separate se_mrstft_loss_value launch)
int se_mrstft_loss_fwd_value(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums, float* loss,
                             void* workspace, void* stream) {
    if (!loss) return fail(SE_ERR_BAD_ARG, "null pointer");
    return loss_fwd_impl(est, ref, rows, nsample, sums, loss, workspace, stream);
}

static int loss_fwd_impl(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums, float* loss,
                         void* workspace, void* stream) {
    if (!est || !ref || !sums || !workspace) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nsample < 2048) return fail(SE_ERR_BAD_ARG, "need rows > 0 and nsample >= 2048");
    double* part0 = reinterpret_cast<double*>(workspace);
    float* refmag0 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + loss_partials_bytes(rows, nsample));
    int nres[3] = {0, 0, 0}, gpcs[3], nchs[3];
    for (int r = 0; r < 3; ++r) {
        if (int rc = check_common(rows, nsample, kRes[r][0], kRes[r][1], kRes[r][2])) return rc;
        nres[r] = loss_fwd_plan(rows, nsample, r, gpcs[r], nchs[r]);
    }
    // largest transform first; each follower starts in the previous kernel's tail (see k_loss_fwd)
    for (int r = 2; r >= 0; --r) {
        const int n = kRes[r][0], hop = kRes[r][1], win = kRes[r][2];
        LossArgs a{};
        if (int rc = get_tables(n, hop, win, false, 0.5f, a.tb)) return rc;
        double* part = part0;
        float* refmag = refmag0;
        for (int q = 0; q < r; ++q) part += (size_t)nres[q] * 3;
        refmag += loss_bins_before(rows, nsample, r);
        a.est = est; a.ref = ref; a.partials = part; a.refmag = refmag;
        a.estspec = loss_recompute() ? nullptr
                                     : reinterpret_cast<float2*>(refmag0 + loss_refmag_region_floats(rows, nsample)) + loss_bins_before(rows, nsample, r);
        a.nsample = (int)nsample; a.nframe = (int)(1 + nsample / hop);
        a.gpc = gpcs[r]; a.nchunks = nchs[r];
        a.chained = r != 2;
        cudaError_t e;
        if (engine_version() == 2) SE_DISPATCH_LOSS_GEO2(n, true, (e = run_loss_fwd2<G>(a, rows, (cudaStream_t)stream)));
        else SE_DISPATCH_LOSS_FWD_GEO(n, (e = run_loss_fwd<G>(a, rows, (cudaStream_t)stream)));
        if (e != cudaSuccess) return cuda_fail(e, "se_mrstft_loss_fwd launch");
    }
    // plain stream serialisation: the reduction needs ALL three kernels, not just its immediate predecessor
    double cnt[3];
    for (int r = 0; r < 3; ++r) cnt[r] = (double)rows * (kRes[r][0] / 2 + 1) * (double)(1 + nsample / kRes[r][1]);
    cudaError_t e = launch_ex(false, k_reduce_partials, loss ? 1u : 3u, 256u, 0, (cudaStream_t)stream,
                              (const double*)part0, nres[0], nres[1], nres[2], sums, loss, cnt[0], cnt[1], cnt[2]);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_loss_fwd reduce launch");
}

int se_mrstft_loss_value(const double* sums, int64_t global_rows, int64_t nsample, float* loss, void* stream) {
    if (!sums || !loss || global_rows <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty batch");
    double cnt[3];
    for (int r = 0; r < 3; ++r)
        cnt[r] = (double)global_rows * (kRes[r][0] / 2 + 1) * (double)(1 + nsample / kRes[r][1]);
    cudaError_t e = launch(k_loss_value, 1u, 32u, 0, (cudaStream_t)stream, sums, cnt[0], cnt[1], cnt[2], loss, (const double*)nullptr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_loss_value launch");
}

int se_mrstft_loss_value_dev(const double* sums10, int64_t nsample, float* loss, void* stream) {
    if (!sums10 || !loss) return fail(SE_ERR_BAD_ARG, "null pointer");
    double per_row[3];
    for (int r = 0; r < 3; ++r) per_row[r] = (double)(kRes[r][0] / 2 + 1) * (double)(1 + nsample / kRes[r][1]);
    cudaError_t e = launch(k_loss_value, 1u, 32u, 0, (cudaStream_t)stream, sums10, per_row[0], per_row[1], per_row[2], loss, sums10 + 9);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_loss_value_dev launch");
}

int se_mrstft_loss_bwd(const float* est, const void* workspace, const double* sums, const float* gout, int64_t global_rows,
                       int64_t rows, int64_t nsample, float* g_est, void* stream) {
    if (!est || !workspace || !sums || !gout || !g_est) return fail(SE_ERR_BAD_ARG, "null pointer");
    // global_rows == 0: `sums` holds 10 doubles and sums[9] is the global row count (device side, uneven shards)
    if (rows <= 0 || (global_rows != 0 && global_rows < rows) || nsample < 2048)
        return fail(SE_ERR_BAD_ARG, "need 0 < rows <= global_rows (or global_rows == 0: count in sums[9]), nsample >= 2048");
    const float* refmag0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(workspace) + loss_partials_bytes(rows, nsample));
    for (int r = 0; r < 3; ++r)
        if (int rc = check_common(rows, nsample, kRes[r][0], kRes[r][1], kRes[r][2])) return rc;
    // largest transform first; followers start in the previous kernel's tail and wait only before they accumulate
    // into g_est (see k_loss_bwd).  Saved-spectrum mode goes smallest first: the forward pass wrote that scratch
    // last, so part of it is still in L2 (measured 211 -> 205 us).
    const bool asc = !loss_recompute();
    for (int idx = 0; idx < 3; ++idx) {
        const int r = asc ? idx : 2 - idx;
        const int n = kRes[r][0], hop = kRes[r][1], win = kRes[r][2];
        LossArgs a{};
        if (int rc = get_tables(n, hop, win, false, 0.5f, a.tb)) return rc;
        const float* refmag = refmag0 + loss_bins_before(rows, nsample, r);
        a.est = est; a.refmag = const_cast<float*>(refmag); a.g_est = g_est; a.sums = sums + 3 * r; a.gout = gout;
        a.nsample = (int)nsample; a.nframe = (int)(1 + nsample / hop);
        a.b_lo = 0; a.b_hi = (int)((nsample + n + hop - 1) / hop);
        a.accumulate = idx != 0;
        a.chained = idx != 0;
        a.inv_count = global_rows ? (float)(1.0 / ((double)global_rows * (n / 2 + 1) * (double)a.nframe)) : 0.f;
        a.rows_dev = global_rows ? nullptr : sums + 9;
        a.bins_per_row = (double)(n / 2 + 1) * (double)a.nframe;
        a.inv_res = 1.0f / 3.0f;
        cudaError_t e;
        if (loss_recompute() && engine_version() == 2) {
            SE_DISPATCH_LOSS_GEO2(n, false, (e = run_loss_bwd2<G>(a, rows, (cudaStream_t)stream)));
        } else if (loss_recompute()) {
            SE_DISPATCH_LOSS_GEO(n, (e = run_loss_bwd<G>(a, rows, (cudaStream_t)stream)));
        } else {
            a.estspec = const_cast<float2*>(reinterpret_cast<const float2*>(refmag0 + loss_refmag_region_floats(rows, nsample))) +
                        loss_bins_before(rows, nsample, r);
            SE_DISPATCH_LOSS_GEO(n, (e = run_loss_bwd_saved<G>(a, rows, (cudaStream_t)stream)));
        }
        if (e != cudaSuccess) return cuda_fail(e, "se_mrstft_loss_bwd launch");
    }
    return 0;
}

}  // extern "C"
