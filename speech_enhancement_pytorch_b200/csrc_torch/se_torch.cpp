// se_torch.cpp -- the thin PyTorch C++ extension north_star names: TORCH_LIBRARY(se_b200, ...) operators with C++
// autograd nodes over the C-ABI of libse_b200.so (include/se_b200.h).  No kernels here: every op allocates its
// outputs, takes the current CUDA stream and calls the same extern "C" entry points the ctypes binding calls.
//
// Reference call sites these ops stay cheap for: src/solver.py:457-458 (stft_custom / istft_custom around the model),
// :466 (model forward tail), :480 (loss_function(enhanced, sources)); src/evaluate.py:39,72.
// Built by _native.build() with g++ against the torch headers, linked to libse_b200.so ($ORIGIN rpath), in-tree.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/autograd/custom_function.h>
#include <torch/library.h>

#include "../../include/se_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check(int rc) {
    if (rc == 0) return;
    const char* msg = se_last_error();
    if (rc == SE_ERR_BAD_ARG) TORCH_CHECK_VALUE(false, "se_b200[", rc, "]: ", msg);
    if (rc == SE_ERR_UNSUPPORTED) TORCH_CHECK_NOT_IMPLEMENTED(false, "se_b200[", rc, "]: ", msg);
    TORCH_CHECK(false, "se_b200[", rc, "]: ", msg);
}

// CUDA-only, fp32 at the C-ABI: half / bf16 inputs are up-cast ("bf16 model, fp32 spectra", BASELINE cfg 4); float64
// inputs (the reference passes dtype=tensor.dtype to its window and works in double, src/evaluate.py:113,147) are
// computed in fp32 and the result is returned as float64 -- the casts are ordinary differentiable ops, so gradients
// come back in the input's dtype.
Tensor prep(const Tensor& t, const char* what) {
    TORCH_CHECK(t.is_cuda(), "speech_enhancement_pytorch_b200 runs on CUDA tensors only (hand-written sm_100a kernels; "
                             "there is no CPU fallback): ", what, " is on ", t.device());
    Tensor r = t;
    if (r.scalar_type() == at::kHalf || r.scalar_type() == at::kBFloat16 || r.scalar_type() == at::kDouble) r = r.to(at::kFloat);
    TORCH_CHECK_TYPE(r.scalar_type() == at::kFloat, "expected float32 tensors, got ", r.scalar_type(), " for ", what);
    return r.contiguous();
}
void* stream_of(const Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
const float* cp(const Tensor& t) { return t.const_data_ptr<float>(); }
float* mp(Tensor& t) { return t.mutable_data_ptr<float>(); }

// fused = nullptr: any geometry the general path takes; otherwise the name of a fused op built for the tuned geometries only
void check_cfg(int64_t n_fft, int64_t hop, int64_t win, const char* fused = nullptr) {
    TORCH_CHECK_NOT_IMPLEMENTED(n_fft >= 8 && n_fft <= 8192 && !(n_fft & 1), "n_fft=", n_fft,
                                ": even sizes in 8..8192 are built (no CPU / cuFFT fallback path)");
    TORCH_CHECK_NOT_IMPLEMENTED(hop >= 1 && hop <= n_fft, "hop_length=", hop, " must be in [1, n_fft]");
    TORCH_CHECK_NOT_IMPLEMENTED(win >= 2 && win <= n_fft, "win_length=", win, " must be in [2, n_fft]");
    TORCH_CHECK_NOT_IMPLEMENTED(!fused || se_geometry_tuned((int)n_fft, (int)hop), fused,
                                ": built for n_fft 512/1024/2048 at hop n_fft/4 or n_fft/2, got ", n_fft, "/", hop);
}

// ------------------------------------------------------------------ raw calls
Tensor stft_raw(const Tensor& x, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    c10::cuda::CUDAGuard guard(x.device());
    const int64_t rows = x.size(0), n = x.size(1);
    Tensor out = at::empty({rows, n_fft / 2 + 1, 1 + n / hop, 2}, x.options());
    check(se_stft_fwd(cp(x), mp(out), rows, n, (int)n_fft, (int)hop, (int)win, (float)scale, stream_of(x)));
    return out;
}
Tensor stft_adj_raw(const Tensor& g, int64_t nsample, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    c10::cuda::CUDAGuard guard(g.device());
    Tensor out = at::empty({g.size(0), nsample}, g.options());
    check(se_stft_bwd(cp(g), mp(out), g.size(0), nsample, (int)n_fft, (int)hop, (int)win, (float)scale, 0, stream_of(g)));
    return out;
}
Tensor istft_raw(const Tensor& spec, int64_t length, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    c10::cuda::CUDAGuard guard(spec.device());
    Tensor out = at::empty({spec.size(0), length}, spec.options());
    check(se_istft_fwd(cp(spec), mp(out), spec.size(0), spec.size(2), length, (int)n_fft, (int)hop, (int)win, (float)scale,
                       stream_of(spec)));
    return out;
}
Tensor istft_adj_raw(const Tensor& gy, int64_t nframe, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    c10::cuda::CUDAGuard guard(gy.device());
    Tensor out = at::empty({gy.size(0), n_fft / 2 + 1, nframe, 2}, gy.options());
    check(se_istft_bwd(cp(gy), mp(out), gy.size(0), nframe, gy.size(1), (int)n_fft, (int)hop, (int)win, (float)scale,
                       stream_of(gy)));
    return out;
}

// ------------------------------------------------------------------ autograd nodes
struct StftFn : torch::autograd::Function<StftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& x, int64_t n_fft, int64_t hop, int64_t win, double scale) {
        ctx->saved_data["cfg"] = std::vector<int64_t>{x.size(1), n_fft, hop, win};
        ctx->saved_data["scale"] = scale;
        return stft_raw(x, n_fft, hop, win, scale);
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto c = ctx->saved_data["cfg"].toIntVector();
        return {stft_adj_raw(g[0].contiguous(), c[0], c[1], c[2], c[3], ctx->saved_data["scale"].toDouble()), Tensor(), Tensor(),
                Tensor(), Tensor()};
    }
};
// config.center = False: no padding, T = 1 + (N - n_fft) / hop
struct StftNoCenterFn : torch::autograd::Function<StftNoCenterFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& x, int64_t n_fft, int64_t hop, int64_t win, double scale) {
        c10::cuda::CUDAGuard guard(x.device());
        const int64_t rows = x.size(0), n = x.size(1);
        TORCH_CHECK_VALUE(n >= n_fft, "stft (center=False): the input (", n, " samples) is shorter than n_fft = ", n_fft);
        Tensor out = at::empty({rows, n_fft / 2 + 1, 1 + (n - n_fft) / hop, 2}, x.options());
        check(se_stft_nocenter_fwd(cp(x), mp(out), rows, n, (int)n_fft, (int)hop, (int)win, (float)scale, stream_of(x)));
        ctx->saved_data["cfg"] = std::vector<int64_t>{n, n_fft, hop, win};
        ctx->saved_data["scale"] = scale;
        return out;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto c = ctx->saved_data["cfg"].toIntVector();
        const Tensor gs = g[0].contiguous();
        c10::cuda::CUDAGuard guard(gs.device());
        Tensor gx = at::empty({gs.size(0), c[0]}, gs.options());
        check(se_stft_nocenter_bwd(cp(gs), mp(gx), gs.size(0), c[0], (int)c[1], (int)c[2], (int)c[3],
                                   (float)ctx->saved_data["scale"].toDouble(), 0, stream_of(gs)));
        return {gx, Tensor(), Tensor(), Tensor(), Tensor()};
    }
};
struct IstftFn : torch::autograd::Function<IstftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& spec, int64_t length, int64_t n_fft, int64_t hop, int64_t win,
                          double scale) {
        ctx->saved_data["cfg"] = std::vector<int64_t>{spec.size(2), n_fft, hop, win};
        ctx->saved_data["scale"] = scale;
        return istft_raw(spec, length, n_fft, hop, win, scale);
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto c = ctx->saved_data["cfg"].toIntVector();
        return {istft_adj_raw(g[0].contiguous(), c[0], c[1], c[2], c[3], ctx->saved_data["scale"].toDouble()), Tensor(), Tensor(),
                Tensor(), Tensor(), Tensor()};
    }
};

struct MaskFn : torch::autograd::Function<MaskFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& spec, const Tensor& mask, int64_t mode, bool pre_tanh) {
        c10::cuda::CUDAGuard guard(spec.device());
        Tensor out = at::empty_like(spec);
        check(se_mask_fwd(cp(spec), cp(mask), mp(out), spec.numel() / 2, (int)mode, pre_tanh ? 1 : 0, stream_of(spec)));
        ctx->save_for_backward({spec, mask});
        ctx->saved_data["mode"] = mode;
        ctx->saved_data["tanh"] = pre_tanh;
        return out;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &spec = saved[0], &mask = saved[1];
        c10::cuda::CUDAGuard guard(spec.device());
        const Tensor go = g[0].contiguous();
        Tensor gmask = at::empty_like(mask);
        Tensor gspec;
        if (ctx->needs_input_grad(0)) gspec = at::empty_like(spec);
        check(se_mask_bwd(cp(spec), cp(mask), cp(go), mp(gmask), gspec.defined() ? mp(gspec) : nullptr, spec.numel() / 2,
                          (int)ctx->saved_data["mode"].toInt(), ctx->saved_data["tanh"].toBool() ? 1 : 0, stream_of(spec)));
        return {gspec, gmask, Tensor(), Tensor()};
    }
};

// istft_custom(apply_mask(spec, mask)) in one launch each way; gradient flows to the raw mask only
struct MaskIstftFn : torch::autograd::Function<MaskIstftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& spec, const Tensor& mask, int64_t length, int64_t n_fft, int64_t hop,
                          int64_t win, double scale, int64_t mode, bool pre_tanh) {
        c10::cuda::CUDAGuard guard(spec.device());
        Tensor y = at::empty({spec.size(0), length}, spec.options());
        check(se_mask_istft_fwd(cp(spec), cp(mask), mp(y), spec.size(0), spec.size(2), length, (int)n_fft, (int)hop, (int)win,
                                (float)scale, (int)mode, pre_tanh ? 1 : 0, stream_of(spec)));
        ctx->save_for_backward({spec, mask});
        ctx->saved_data["cfg"] = std::vector<int64_t>{length, n_fft, hop, win, mode, pre_tanh ? 1 : 0};
        ctx->saved_data["scale"] = scale;
        return y;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &spec = saved[0], &mask = saved[1];
        const auto c = ctx->saved_data["cfg"].toIntVector();
        c10::cuda::CUDAGuard guard(spec.device());
        const Tensor gy = g[0].contiguous();
        Tensor gmask = at::empty_like(mask);
        check(se_mask_istft_bwd(cp(gy), cp(spec), cp(mask), mp(gmask), spec.size(0), spec.size(2), c[0], (int)c[1], (int)c[2],
                                (int)c[3], (float)ctx->saved_data["scale"].toDouble(), (int)c[4], (int)c[5], stream_of(spec)));
        return {Tensor(), gmask, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

// wave -> STFT -> mask -> iSTFT -> wave in one launch; backward to the raw mask
struct EnhanceFn : torch::autograd::Function<EnhanceFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& x, const Tensor& mask, int64_t n_fft, int64_t hop, int64_t win,
                          int64_t mode, bool pre_tanh) {
        c10::cuda::CUDAGuard guard(x.device());
        Tensor y = at::empty_like(x);
        check(se_enhance_fwd(cp(x), cp(mask), mp(y), x.size(0), x.size(1), (int)n_fft, (int)hop, (int)win, (int)mode,
                             pre_tanh ? 1 : 0, stream_of(x)));
        ctx->save_for_backward({x, mask});
        ctx->saved_data["cfg"] = std::vector<int64_t>{n_fft, hop, win, mode, pre_tanh ? 1 : 0};
        return y;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &x = saved[0], &mask = saved[1];
        const auto c = ctx->saved_data["cfg"].toIntVector();
        TORCH_CHECK_NOT_IMPLEMENTED(!ctx->needs_input_grad(0), "enhance: gradient wrt the input waveform is not built");
        c10::cuda::CUDAGuard guard(x.device());
        const Tensor gy = g[0].contiguous();
        Tensor gmask = at::empty_like(mask);
        if (c[0] > 1024) {
            // two 2048-point working sets do not fit one SM's shared memory: compose the three kernels
            const Tensor spec = stft_raw(x, c[0], c[1], c[2], 1.0 / (double)c[2]);
            const Tensor gspec = istft_adj_raw(gy, spec.size(2), c[0], c[1], c[2], (double)c[2]);
            check(se_mask_bwd(cp(spec), cp(mask), cp(gspec), mp(gmask), nullptr, spec.numel() / 2, (int)c[3], (int)c[4], stream_of(x)));
        } else {
            check(se_enhance_bwd(cp(gy), cp(x), cp(mask), mp(gmask), x.size(0), x.size(1), (int)c[0], (int)c[1], (int)c[2], (int)c[3],
                                 (int)c[4], stream_of(x)));
        }
        return {Tensor(), gmask, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

// MR-STFT loss of one process (no exchange step; the sharded variant stays in ops.py next to torch.distributed)
struct MrstftFn : torch::autograd::Function<MrstftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& est, const Tensor& ref) {
        c10::cuda::CUDAGuard guard(est.device());
        const int64_t rows = est.size(0), n = est.size(1);
        int64_t bytes = se_mrstft_workspace_bytes(rows, n);
        Tensor ws = at::empty({bytes < 8 ? 8 : bytes}, est.options().dtype(at::kByte));
        Tensor sums = at::empty({9}, est.options().dtype(at::kDouble));
        Tensor loss = at::empty({}, est.options());
        void* st = stream_of(est);
        check(se_mrstft_loss_fwd_value(cp(est), cp(ref), rows, n, sums.mutable_data_ptr<double>(), mp(loss), ws.mutable_data_ptr(), st));
        ctx->save_for_backward({est, ws, sums});
        return loss;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &est = saved[0], &ws = saved[1], &sums = saved[2];
        c10::cuda::CUDAGuard guard(est.device());
        const Tensor gout = g[0].to(at::kFloat).contiguous();
        Tensor gx = at::empty_like(est);
        check(se_mrstft_loss_bwd(cp(est), ws.const_data_ptr(), sums.const_data_ptr<double>(), cp(gout), est.size(0), est.size(0),
                                 est.size(1), mp(gx), stream_of(est)));
        return {gx, Tensor()};
    }
};

// ------------------------------------------------------------------ DCCRN in-model transforms (src/model/dccrn.py:669-747)
Tensor conv_stft_raw(const Tensor& x, int64_t win_len, int64_t win_inc, int64_t fft_len, int64_t window_id) {
    c10::cuda::CUDAGuard guard(x.device());
    const int64_t rows = x.size(0), n = x.size(1);
    const int64_t nt = (n + 2 * (win_len - win_inc) - win_len) / win_inc + 1;
    Tensor out = at::empty({rows, 2 * (fft_len / 2 + 1), nt}, x.options());
    check(se_conv_stft_fwd_w(cp(x), mp(out), rows, n, (int)win_len, (int)win_inc, (int)fft_len, (int)window_id, stream_of(x)));
    return out;
}
struct ConvIstftFn : torch::autograd::Function<ConvIstftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& spec, int64_t out_len, int64_t win_len, int64_t win_inc, int64_t fft_len,
                          int64_t window_id) {
        c10::cuda::CUDAGuard guard(spec.device());
        Tensor y = at::empty({spec.size(0), out_len}, spec.options());
        check(se_conv_istft_fwd_w(cp(spec), mp(y), spec.size(0), spec.size(2), out_len, (int)win_len, (int)win_inc, (int)fft_len,
                                  (int)window_id, stream_of(spec)));
        ctx->saved_data["cfg"] = std::vector<int64_t>{spec.size(2), win_len, win_inc, fft_len, window_id};
        return y;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto c = ctx->saved_data["cfg"].toIntVector();
        const Tensor gy = g[0].contiguous();
        c10::cuda::CUDAGuard guard(gy.device());
        Tensor gs = at::empty({gy.size(0), 2 * (c[3] / 2 + 1), c[0]}, gy.options());
        check(se_conv_istft_bwd_w(cp(gy), mp(gs), gy.size(0), c[0], gy.size(1), (int)c[1], (int)c[2], (int)c[3], (int)c[4], stream_of(gy)));
        return {gs, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};
// DCCRN's mask tail (dccrn.py:203-223) + ConviSTFT (:224) in one launch each way; gradient to the two mask planes
struct ConvMaskIstftFn : torch::autograd::Function<ConvMaskIstftFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& spec, const Tensor& mre, const Tensor& mim, int64_t out_len,
                          int64_t win_len, int64_t win_inc, int64_t fft_len, int64_t mode, int64_t window_id) {
        c10::cuda::CUDAGuard guard(spec.device());
        Tensor y = at::empty({spec.size(0), out_len}, spec.options());
        check(se_conv_mask_istft_fwd_w(cp(spec), cp(mre), cp(mim), mp(y), spec.size(0), spec.size(2), out_len, (int)win_len, (int)win_inc,
                                       (int)fft_len, (int)mode, (int)window_id, stream_of(spec)));
        ctx->save_for_backward({spec, mre, mim});
        ctx->saved_data["cfg"] = std::vector<int64_t>{out_len, win_len, win_inc, fft_len, mode, window_id};
        return y;
    }
    static variable_list backward(AutogradContext* ctx, variable_list g) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &spec = saved[0], &mre = saved[1], &mim = saved[2];
        const auto c = ctx->saved_data["cfg"].toIntVector();
        c10::cuda::CUDAGuard guard(spec.device());
        const Tensor gy = g[0].contiguous();
        Tensor gre = at::empty_like(mre), gim = at::empty_like(mim);
        check(se_conv_mask_istft_bwd_w(cp(gy), cp(spec), cp(mre), cp(mim), mp(gre), mp(gim), spec.size(0), spec.size(2), c[0], (int)c[1],
                                       (int)c[2], (int)c[3], (int)c[4], (int)c[5], stream_of(spec)));
        return {Tensor(), gre, gim, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

bool needs_grad(std::initializer_list<const Tensor*> ts) {
    if (!at::GradMode::is_enabled()) return false;
    for (const Tensor* t : ts)
        if (t->requires_grad()) return true;
    return false;
}

// ------------------------------------------------------------------ operators (rows already flattened by the Python shims)
Tensor op_stft(const Tensor& x_in, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    check_cfg(n_fft, hop, win);
    TORCH_CHECK_VALUE(x_in.dim() == 2, "se_b200::stft expects [rows, N]");
    const Tensor x = prep(x_in, "input");
    const Tensor out = needs_grad({&x}) ? StftFn::apply(x, n_fft, hop, win, scale)
                                        : stft_raw(x, n_fft, hop, win, scale);        // inference: no autograd node
    return x_in.scalar_type() == at::kDouble ? out.to(at::kDouble) : out;
}
Tensor op_stft_nocenter(const Tensor& x_in, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    check_cfg(n_fft, hop, win);
    TORCH_CHECK_VALUE(x_in.dim() == 2, "se_b200::stft_nocenter expects [rows, N]");
    const Tensor out = StftNoCenterFn::apply(prep(x_in, "input"), n_fft, hop, win, scale);
    return x_in.scalar_type() == at::kDouble ? out.to(at::kDouble) : out;
}
Tensor op_istft(const Tensor& spec_in, int64_t length, int64_t n_fft, int64_t hop, int64_t win, double scale) {
    check_cfg(n_fft, hop, win);
    TORCH_CHECK_VALUE(spec_in.dim() == 4 && spec_in.size(3) == 2, "se_b200::istft expects [rows, F, T, 2]");
    const Tensor spec = prep(spec_in, "spectrum");
    const Tensor out = needs_grad({&spec}) ? IstftFn::apply(spec, length, n_fft, hop, win, scale)
                                           : istft_raw(spec, length, n_fft, hop, win, scale);
    return spec_in.scalar_type() == at::kDouble ? out.to(at::kDouble) : out;
}
Tensor op_mask(const Tensor& spec, const Tensor& mask, int64_t mode, bool pre_tanh) {
    return MaskFn::apply(prep(spec, "spectrum"), prep(mask, "mask"), mode, pre_tanh);
}
Tensor op_mask_istft(const Tensor& spec, const Tensor& mask, int64_t length, int64_t n_fft, int64_t hop, int64_t win, double scale,
                     int64_t mode, bool pre_tanh) {
    check_cfg(n_fft, hop, win, "mask_istft");
    return MaskIstftFn::apply(prep(spec, "spectrum"), prep(mask, "mask"), length, n_fft, hop, win, scale, mode, pre_tanh);
}
Tensor op_enhance(const Tensor& x, const Tensor& mask, int64_t n_fft, int64_t hop, int64_t win, int64_t mode, bool pre_tanh) {
    check_cfg(n_fft, hop, win, "enhance");
    return EnhanceFn::apply(prep(x, "input"), prep(mask, "mask"), n_fft, hop, win, mode, pre_tanh);
}
Tensor op_mrstft(const Tensor& est, const Tensor& ref) {
    TORCH_CHECK_NOT_IMPLEMENTED(!ref.requires_grad(), "loss_mrstft: gradient flows to `enhanced` only (targets must not require grad)");
    return MrstftFn::apply(prep(est, "enhanced"), prep(ref, "sources"));
}

Tensor op_conv_stft(const Tensor& x, int64_t win_len, int64_t win_inc, int64_t fft_len, int64_t window_id) {
    TORCH_CHECK_VALUE(x.dim() == 2, "se_b200::conv_stft expects [rows, N]");
    TORCH_CHECK_NOT_IMPLEMENTED(!(at::GradMode::is_enabled() && x.requires_grad()),
                                "ConvSTFT: gradient wrt the input waveform is not built (the mixture is data)");
    return conv_stft_raw(prep(x, "input"), win_len, win_inc, fft_len, window_id);
}
Tensor op_conv_istft(const Tensor& spec, int64_t out_len, int64_t win_len, int64_t win_inc, int64_t fft_len, int64_t window_id) {
    TORCH_CHECK_VALUE(spec.dim() == 3, "se_b200::conv_istft expects [rows, 2F, T]");
    return ConvIstftFn::apply(prep(spec, "spectrum"), out_len, win_len, win_inc, fft_len, window_id);
}
Tensor op_conv_mask_istft(const Tensor& spec, const Tensor& mre, const Tensor& mim, int64_t out_len, int64_t win_len, int64_t win_inc,
                          int64_t fft_len, int64_t mode, int64_t window_id) {
    TORCH_CHECK_VALUE(spec.dim() == 3 && mre.dim() == 3 && mim.dim() == 3, "se_b200::conv_mask_istft expects [rows, 2F, T] and two [rows, F, T] masks");
    return ConvMaskIstftFn::apply(prep(spec, "spectrum"), prep(mre, "mask"), prep(mim, "mask"), out_len, win_len, win_inc, fft_len, mode,
                                  window_id);
}

}  // namespace

TORCH_LIBRARY(se_b200, m) {
    m.def("stft_nocenter(Tensor x, int n_fft, int hop, int win_length, float scale) -> Tensor");
    m.def("conv_stft(Tensor x, int win_len, int win_inc, int fft_len, int window_id) -> Tensor");
    m.def("conv_istft(Tensor spec, int out_len, int win_len, int win_inc, int fft_len, int window_id) -> Tensor");
    m.def("conv_mask_istft(Tensor spec, Tensor mask_re, Tensor mask_im, int out_len, int win_len, int win_inc, int fft_len, int mode, int window_id) -> Tensor");
    m.def("stft(Tensor x, int n_fft, int hop, int win_length, float scale) -> Tensor");
    m.def("istft(Tensor spec, int length, int n_fft, int hop, int win_length, float scale) -> Tensor");
    m.def("mask(Tensor spec, Tensor mask, int mode, bool pre_tanh) -> Tensor");
    m.def("mask_istft(Tensor spec, Tensor mask, int length, int n_fft, int hop, int win_length, float scale, int mode, bool pre_tanh) -> Tensor");
    m.def("enhance(Tensor x, Tensor mask, int n_fft, int hop, int win_length, int mode, bool pre_tanh) -> Tensor");
    m.def("mrstft_loss(Tensor est, Tensor ref) -> Tensor");
}
// Registered for every dispatch key: the ops do their own device check (CPU tensors raise) and build the autograd graph
// themselves with C++ autograd nodes.
TORCH_LIBRARY_IMPL(se_b200, CompositeImplicitAutograd, m) {
    m.impl("stft", op_stft);
    m.impl("istft", op_istft);
    m.impl("mask", op_mask);
    m.impl("mask_istft", op_mask_istft);
    m.impl("enhance", op_enhance);
    m.impl("mrstft_loss", op_mrstft);
    m.impl("stft_nocenter", op_stft_nocenter);
    m.impl("conv_stft", op_conv_stft);
    m.impl("conv_istft", op_conv_istft);
    m.impl("conv_mask_istft", op_conv_mask_istft);
}
