// ce_tma_bf16.cu — bf16 instantiations of the TMA-staged K1 (see ce_tma_impl.cuh).
// Arithmetic is fp32; gradients are rounded to bf16 once (round-to-nearest-even).
#include "ce_tma_impl.cuh"

namespace cvcs {

int ce_tma_launch_bf16(const CeParams& p, int layout, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (layout == CVCS_NCHW) {
        if (p.C <= 8) {
            // 4 pixels (two bf16x2 pairs) per thread: 96 registers -> two CTAs per SM; 8 per thread needs 150+
            // default: two sub-chunks of 4 pixels per thread per stage — 4 KB bulk copies, half the per-stage bookkeeping
            // per pixel, 8-bit private counters so that two CTAs x three stages fit (cfg3: 0.860 of the copy peak against
            // 0.836 for one sub-chunk x four stages).  OPT_TMA_VECP == 4 selects the single sub-chunk form.
            if (get_option(CVCS_OPT_TMA_VECP) != 4) return tma::dispatch<__nv_bfloat16, 4, false, 2, 8, 2>(p, stream, handled);
            return tma::dispatch<__nv_bfloat16, 4, false, 2, 8>(p, stream, handled);
        }
        if (p.C <= 16) return tma::dispatch<__nv_bfloat16, 4, false, 9, 16>(p, stream, handled);
        return tma::dispatch<__nv_bfloat16, 2, false, 17, kMaxRegC>(p, stream, handled);
    }
    if (p.C <= 12) return tma::dispatch<__nv_bfloat16, 8, true, 2, 12>(p, stream, handled);
    if (p.C % 4 == 0) return tma::dispatch<__nv_bfloat16, 2, true, 13, kMaxRegC>(p, stream, handled);   // 16, 20
    if (p.C % 2 == 0) return tma::dispatch<__nv_bfloat16, 4, true, 13, kMaxRegC>(p, stream, handled);   // 14, 18
    return CVCS_OK;  // odd C > 12: generic variant
}

}  // namespace cvcs
