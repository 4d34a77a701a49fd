// ce_tma_f32.cu — fp32 instantiations of the TMA-staged K1 (see ce_tma_impl.cuh).
#include "ce_tma_impl.cuh"

namespace cvcs {

// pixels per thread chosen so that one stage (C planes x 256*VECP pixels) stays <= 32 KB
int ce_tma_launch_f32(const CeParams& p, int layout, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (layout == CVCS_NCHW) {
        if (p.C <= 8) {
            if (get_option(CVCS_OPT_TMA_VECP) == 2) return tma::dispatch<float, 2, false, 2, 8>(p, stream, handled);
            return tma::dispatch<float, 4, false, 2, 8>(p, stream, handled);
        }
        if (p.C <= 16) return tma::dispatch<float, 2, false, 9, 16>(p, stream, handled);
        return tma::dispatch<float, 1, false, 17, kMaxRegC>(p, stream, handled);
    }
    // NHWC: the thread's span (VECP pixels x C floats) must be whole 16-byte vectors
    if (p.C <= 12) return tma::dispatch<float, 4, true, 2, 12>(p, stream, handled);
    if (p.C % 4 == 0) return tma::dispatch<float, 1, true, 13, kMaxRegC>(p, stream, handled);   // 16, 20
    if (p.C % 2 == 0) return tma::dispatch<float, 2, true, 13, kMaxRegC>(p, stream, handled);   // 14, 18
    return CVCS_OK;  // odd C > 12: generic variant
}

}  // namespace cvcs
