// se_api_specloss.cu -- fused STFT-domain mse / l1 losses against a waveform target.
#include "se_host.h"
#include "se_specloss.cuh"

using namespace se;

static int specloss_args(SpecLossArgs& a, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length, float scale,
                         int kind) {
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    if (kind != 0 && kind != 1) return fail(SE_ERR_UNSUPPORTED, "spectral loss kind must be 0 (mse) or 1 (l1)");
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.nsample = (int)nsample; a.nframe = (int)(1 + nsample / hop); a.kind = kind;
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks, 16);     // every SE_DISPATCH_GEO geometry uses 16-frame groups
    return 0;
}

extern "C" int64_t se_spectral_loss_workspace_bytes(int64_t rows, int64_t nsample, int hop) {
    int gpc, nchunks;
    plan_analysis(rows, 1 + nsample / (hop > 0 ? hop : 1), gpc, nchunks, 16);
    return rows * nchunks * (int64_t)sizeof(double);
}

extern "C" int se_spectral_loss_fwd(const float* enh, const float* target, int64_t rows, int64_t nsample, int n_fft, int hop,
                                    int win_length, float scale, int kind, double* sum_out, void* workspace, void* stream) {
    if (!enh || !target || !sum_out || !workspace) return fail(SE_ERR_BAD_ARG, "null pointer");
    SpecLossArgs a{};
    if (int rc = specloss_args(a, rows, nsample, n_fft, hop, win_length, scale, kind)) return rc;
    a.enh = enh; a.target = target; a.partials = reinterpret_cast<double*>(workspace);
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = launch(k_spec_loss<G, false>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS,
                                            (cudaStream_t)stream, a)));
    if (e != cudaSuccess) return cuda_fail(e, "se_spectral_loss_fwd launch");
    e = launch(k_sum_partials, 1u, 256u, 0, (cudaStream_t)stream, (const double*)a.partials, (int)(rows * a.nchunks), sum_out);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_spectral_loss_fwd reduce launch");
}

extern "C" int se_spectral_loss_bwd(const float* enh, const float* target, const float* gout, int64_t global_rows, int64_t rows,
                                    int64_t nsample, int n_fft, int hop, int win_length, float scale, int kind, float* genh,
                                    void* stream) {
    if (!enh || !target || !gout || !genh) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (global_rows < rows) return fail(SE_ERR_BAD_ARG, "global_rows < rows");
    SpecLossArgs a{};
    if (int rc = specloss_args(a, rows, nsample, n_fft, hop, win_length, scale, kind)) return rc;
    a.enh = enh; a.target = target; a.gout = gout; a.genh = genh;
    a.inv_count = (float)(1.0 / ((double)global_rows * (n_fft / 2 + 1) * (double)a.nframe * 2.0));
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = launch(k_spec_loss<G, true>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS,
                                            (cudaStream_t)stream, a)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_spectral_loss_bwd launch");
}
