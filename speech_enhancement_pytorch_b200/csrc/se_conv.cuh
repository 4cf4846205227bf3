// se_conv.cuh -- DCCRN-convention transforms (src/model/dccrn.py:649-747): placeholder until the
// synthesis side lands; the analysis side is k_analysis<Geo<512,win_inc,256>, LOAD_ZEROPAD, true>.
#pragma once
#include "se_kernels.cuh"
