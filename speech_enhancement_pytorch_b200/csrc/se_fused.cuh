// se_fused.cuh -- fused wave -> STFT -> mask -> iSTFT -> wave kernels (placeholder).
#pragma once
#include "se_kernels.cuh"
