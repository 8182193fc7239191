// relaxed.cu -- the relaxed-precision builds of the two streaming kernels (k_light_shade, k_taa): this translation unit is
// compiled WITH FMA contraction and instantiates the kernels with FAST == true (MUFU reciprocal / rsqrt / sqrt instead
// of the IEEE sequences, pow(x, 5) by multiplication, constant divisions by the rounded reciprocal).  They are what a
// host gets; LUZRT_DEBUG_EXACT_MATH selects the bit-faithful builds in light_pass.cu / taa.cu (compiled with -fmad=false),
// which are what the oracle is compared with value by value.  Tolerance of the relaxed build: the contract's (radiance
// max-abs 1e-3 in linear HDR or PSNR >= 50 dB); tests/test_gpu_parity.py runs every scene through both.
#include <cstdlib>

#include "passes.h"
#include "shade_kernel.cuh"
#include "taa_kernel.cuh"

namespace luz {

cudaError_t launch_light_shade_relaxed(cudaStream_t stream, const LightArgs& args) {
    const dim3 sgrid((args.fc.width + 31) / 32, (args.rows.rows + 3) / 4, args.rows.n_bands);
    if (args.fc.shadow_type == LUZW_SHADOW_MAP)
        k_light_shade<true, true><<<sgrid, 128, 0, stream>>>(args);
    else
        k_light_shade<false, true><<<sgrid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_taa_relaxed(cudaStream_t stream, const TaaArgs& args) {
    static const int variant = [] { // rows per thread of the sliding window (LUZRT_TAA_VARIANT: tuning runs)
        const char* e = getenv("LUZRT_TAA_VARIANT");
        return e ? atoi(e) : 0;
    }();
    auto grid = [&](int rows_per_thread) {
        return dim3((args.fc.width + 31) / 32, (args.rows.rows + 8 * rows_per_thread - 1) / (8 * rows_per_thread),
                    args.rows.n_bands);
    };
    if (variant == 1) k_taa<true, 8, 2><<<grid(8), 256, 0, stream>>>(args);
    else if (variant == 2) k_taa<true, 4, 3><<<grid(4), 256, 0, stream>>>(args);
    else if (variant == 3) k_taa<true, 2, 4><<<grid(2), 256, 0, stream>>>(args);
    else if (variant == 4) k_taa<true, 4, 2><<<grid(4), 256, 0, stream>>>(args);
    else if (variant == 5) k_taa<true, 4, 4><<<grid(4), 256, 0, stream>>>(args);
    else k_taa<true, 1, 4><<<grid(1), 256, 0, stream>>>(args);
    return cudaGetLastError();
}

} // namespace luz
