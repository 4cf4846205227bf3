"""speech_enhancement_pytorch_b200 -- B200-native (sm_100a) spectral front/back-end behind the
call signatures of ooshyun/Speech-Enhancement-Pytorch's hot path.

(The task names the package `speech-enhancement-pytorch_b200`; a hyphen is not importable, so
the directory uses underscores.)

Public surface (all CUDA-only, no CPU fallback):
    stft_custom, istft_custom          src/evaluate.py:101-162
    evaluate, segment_stft             src/evaluate.py:10-98,164-183 (segment -> STFT without copies)
    apply_mask, apply_mask_dccrn       model forward tails
    magnitude_feature, stft_custom_with_feature   NN input features
    loss_mrstft, MRSTFTLoss            loss_function(enhanced, sources) convention; group=... shards the batch over
                                       GPUs (distributed.peer_exchange: the exchange step over NVLink peer memory)
    ConvSTFT, ConviSTFT                src/model/dccrn.py:669-747; ConviSTFT.forward_masked = DCCRN tail + iSTFT
                                       (dccrn.py:203-224) in one launch each way
    enhance                            fused stft_custom -> mask -> istft_custom
    apply_mask_istft                   fused model tail -> istft_custom (masked spectrum never written)
    si_snr, loss_sisdr, SI_SDR         src/loss.py:14-29, src/metric.py:92-123
    overlap_and_add                    src/model/conv_tasnet.py:11-31
"""
from .evaluate import stft_custom, istft_custom, stft_custom_with_feature, evaluate, segment_stft, stitch_segments
from .masking import apply_mask, apply_mask_dccrn, magnitude_feature
from .loss import (loss_mrstft, MRSTFTLoss, loss_spectral, si_snr, loss_sisdr,
                   loss_phase_sensitive_spectral_approximation, SI_SDR)
from .tasnet import overlap_and_add
from .dccrn import ConvSTFT, ConviSTFT
from .fused import enhance, apply_mask_istft
from . import _native

__all__ = ["stft_custom", "istft_custom", "stft_custom_with_feature", "magnitude_feature", "evaluate", "segment_stft", "stitch_segments", "apply_mask", "apply_mask_dccrn", "loss_mrstft", "MRSTFTLoss", "loss_spectral", "si_snr", "loss_sisdr", "loss_phase_sensitive_spectral_approximation",
           "ConvSTFT", "ConviSTFT", "enhance", "apply_mask_istft", "SI_SDR", "overlap_and_add"]
__version__ = "0.1.0"
