// ce_direct.cu — K1, direct-load variants.
//
// NCHW "direct": a thread owns VEC consecutive pixels of one image plane and walks the C class
// planes with 128-bit streaming loads (a warp reads 512 contiguous bytes per plane).  All
// C·VEC values stay in registers between the softmax and the gradient, so the logits are read
// exactly once and the gradients written exactly once.  Used when the TMA-staged variant's
// alignment rules do not hold, and as the A/B baseline for it.
//
// "generic": any C <= 1024 and either layout, one pixel per thread, logits re-read from L1/L2
// for the second and third sweep.  Correctness path for shapes without a register-resident
// instantiation.
#include "ce_common.cuh"

namespace cvcs {
namespace {

template <typename T, int C, int VEC, bool PRIV, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) ce_nchw_kernel(const CeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float wsm[C];

    const bool do_grad = p.dlogits != nullptr;
    const bool do_arg = p.argmax != nullptr;
    const bool do_conf = p.confmat != nullptr;

    if (threadIdx.x < C) wsm[threadIdx.x] = p.weight ? p.weight[threadIdx.x] : 1.0f;
    BinAcc<PRIV> conf;
    if (do_conf) conf.init(smem, C * C);  // contains the barrier
    else __syncthreads();

    const float inv_tw = do_grad ? static_cast<float>(p.inv_tw_dev ? *p.inv_tw_dev : p.inv_tw) : 0.f;
    const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
    T* __restrict__ dlogits = reinterpret_cast<T*>(p.dlogits);

    double lsum = 0.0, wsum = 0.0;
    unsigned int bad = 0, since_flush = 0;

    // (b, g) = (image, VEC-pixel group inside the image) of this thread's item, advanced incrementally:
    // one integer division per kernel instead of one per item
    const unsigned int ipi = p.items_per_image;
    const unsigned int stride = gridDim.x * kThreads;
    const unsigned int step_b = stride / ipi, step_g = stride - step_b * ipi;
    unsigned int b, g;
    {
        const unsigned int first = blockIdx.x * kThreads + threadIdx.x;
        b = first / ipi;
        g = first - b * ipi;
    }
    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < p.n_items;
         base += stride, b += step_b, g += step_g) {
        if (g >= ipi) {
            g -= ipi;
            ++b;
        }
        const long long item = base + threadIdx.x;
        if (item < p.n_items) {
            const long long pix = static_cast<long long>(b) * p.hw + static_cast<long long>(g) * VEC;
            const long long off0 = static_cast<long long>(b) * C * p.hw + static_cast<long long>(g) * VEC;

            float x[VEC][C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float v[VEC];
                VecIO<T, VEC>::load(logits + off0 + c * p.hw, v);
#pragma unroll
                for (int k = 0; k < VEC; ++k) x[k][c] = v[k];
            }
            int t[VEC];
            load_targets<VEC>(p, pix, t);

            int amax[VEC];
            float step_l = 0.f, step_w = 0.f;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                amax[k] = pixel_ce<C>(x[k], t[k], wsm, inv_tw, do_grad, step_l, step_w, bad, [&](int c) {
                    float v[1];  // rows with NaN / inf only: one scalar re-read per class
                    VecIO<T, 1>::load(logits + off0 + c * p.hw + k, v);
                    return v[0];
                });
                if (do_conf && static_cast<unsigned int>(t[k]) < static_cast<unsigned int>(C)) conf.add(t[k] * C + amax[k]);
            }
            lsum += static_cast<double>(step_l);
            wsum += static_cast<double>(step_w);
            if (do_grad) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float v[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) v[k] = x[k][c];
                    VecIO<T, VEC>::store(dlogits + off0 + c * p.hw, v);
                }
            }
            if (do_arg) store_argmax<VEC>(p, pix, amax);
        }
        if (PRIV && do_conf) {
            since_flush += VEC;
            if (since_flush > 65535u - VEC) {  // CTA-uniform: the u16 counters cannot overflow
                conf.flush(p.confmat);
                since_flush = 0;
            }
        }
    }
    if (do_conf) conf.flush(p.confmat);
    finish_loss<kWarps>(p, lsum, wsum, bad);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) ce_generic_kernel(const CeParams p, long long class_stride,
                                                              long long pixel_stride, long long image_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int C = p.C;
    const bool do_grad = p.dlogits != nullptr;
    const bool do_conf = p.confmat != nullptr;
    float* wsm = reinterpret_cast<float*>(smem);
    for (int c = threadIdx.x; c < C; c += kThreads) wsm[c] = p.weight ? p.weight[c] : 1.0f;
    BinAcc<false> conf;
    if (do_conf) conf.init(smem + ((C * 4 + 15) / 16) * 16, C * C, p.conf_reps, p.confmat);
    else __syncthreads();
    const float inv_tw = do_grad ? static_cast<float>(p.inv_tw_dev ? *p.inv_tw_dev : p.inv_tw) : 0.f;
    const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
    T* __restrict__ dlogits = reinterpret_cast<T*>(p.dlogits);

    double lsum = 0.0, wsum = 0.0;
    unsigned int bad = 0;
    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < p.n_pixels;
         base += static_cast<long long>(gridDim.x) * kThreads) {
        const long long pix = base + threadIdx.x;
        if (pix >= p.n_pixels) continue;
        const long long b = pix / p.hw;
        const long long off0 = b * image_stride + (pix - b * p.hw) * pixel_stride;
        int t[1];
        load_targets<1>(p, pix, t);
        const int tv = t[0];
        float best, m, xt = 0.f;
        {
            float x0[1];
            VecIO<T, 1>::load(logits + off0, x0);
            best = m = x0[0];
        }
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            float xc[1];
            VecIO<T, 1>::load(logits + off0 + c * class_stride, xc);
            if (better(xc[0], best)) {
                best = xc[0];
                arg = c;
            }
            m = fmaxf(m, xc[0]);
        }
        const bool valid = static_cast<unsigned int>(tv) < static_cast<unsigned int>(C);
        bad += (!valid && tv != -1) ? 1u : 0u;
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            float xc[1];
            VecIO<T, 1>::load(logits + off0 + c * class_stride, xc);
            xt = (c == tv) ? xc[0] : xt;
            s += ex2_ftz((xc[0] - m) * kLog2e);
        }
        const float w = valid ? wsm[tv] : 0.f;
        const float nll = fmaf(lg2_ftz(s), kLn2, m - xt);
        lsum += valid ? static_cast<double>(w * nll) : 0.0;
        wsum += static_cast<double>(w);
        if (do_grad) {
            const float gsc = valid ? w * inv_tw : 0.f;
            const float r = gsc * rcp_ftz(s);
            for (int c = 0; c < C; ++c) {
                float xc[1];
                VecIO<T, 1>::load(logits + off0 + c * class_stride, xc);
                xc[0] = fmaf(ex2_ftz((xc[0] - m) * kLog2e), r, (c == tv) ? -gsc : 0.f);
                VecIO<T, 1>::store(dlogits + off0 + c * class_stride, xc);
            }
        }
        if (p.argmax) {
            int a[1] = {arg};
            store_argmax<1>(p, pix, a);
        }
        if (do_conf && valid) conf.add(tv * C + arg);
    }
    if (do_conf) conf.flush(p.confmat);
    finish_loss<kWarps>(p, lsum, wsum, bad);
}

template <typename T, int C, int VEC>
int launch_nchw(const CeParams& p, cudaStream_t stream) {
    constexpr bool PRIV = C * C <= kPrivBinsMax;
    // register budget: C*VEC fp32 values live per thread (ptxas -v: 80 regs hold C*VEC = 28 with a
    // 12-byte spill; 128 regs hold 56)
    constexpr int MINB = (C * VEC <= 32) ? 3 : (C * VEC <= 64 ? 2 : 1);
    auto kernel = ce_nchw_kernel<T, C, VEC, PRIV, MINB>;
    const int smem = p.confmat ? BinAcc<PRIV>::smem_bytes(C * C) : 0;
    int grid = 0;
    {
        int rc = persistent_grid(kernel, kThreads, smem, &grid);  // cached per (kernel, smem, device)
        if (rc) return rc;
    }
    const long long blocks_needed = (p.n_items + kThreads - 1) / kThreads;
    int g = static_cast<int>(blocks_needed < grid ? blocks_needed : grid);
    if (g < 1) g = 1;
    kernel<<<g, kThreads, smem, stream>>>(p);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

// Register-resident instantiations exist for CLO <= C <= CHI.
template <typename T, int VEC, int CLO, int CHI, int CC = CLO>
int dispatch_c(const CeParams& p, cudaStream_t stream, bool* handled) {
    if constexpr (CC > CHI) {
        *handled = false;
        return CVCS_OK;
    } else {
        if (p.C == CC) {
            *handled = true;
            return launch_nchw<T, CC, VEC>(p, stream);
        }
        return dispatch_c<T, VEC, CLO, CHI, CC + 1>(p, stream, handled);
    }
}

}  // namespace

// vec: pixels per thread the caller validated (f32: 4; bf16: 8 for C <= 12, else 4)
int ce_direct_launch(const CeParams& p, int logits_dtype, int vec, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (p.C < 2 || p.C > kMaxRegC) return CVCS_OK;
    if (logits_dtype == CVCS_F32 && vec == 4) return dispatch_c<float, 4, 2, kMaxRegC>(p, stream, handled);
    if (logits_dtype == CVCS_BF16 && vec == 8) return dispatch_c<__nv_bfloat16, 8, 2, 12>(p, stream, handled);
    if (logits_dtype == CVCS_BF16 && vec == 4) return dispatch_c<__nv_bfloat16, 4, 13, kMaxRegC>(p, stream, handled);
    return CVCS_OK;
}

int ce_generic_launch(const CeParams& p0, int logits_dtype, int layout, cudaStream_t stream) {
    const int C = p0.C;
    CeParams p = p0;
    const long long class_stride = layout == CVCS_NCHW ? p.hw : 1;
    const long long pixel_stride = layout == CVCS_NCHW ? 1 : C;
    const long long image_stride = static_cast<long long>(C) * p.hw;
    p.conf_reps = shared_bin_replicas(C * C);
    const int smem = ((C * 4 + 15) / 16) * 16 + (p.confmat ? BinAcc<false>::smem_bytes(C * C, p.conf_reps) : 0);
    const long long blocks_needed = (p.n_pixels + kThreads - 1) / kThreads;
    int grid = 0;
    if (logits_dtype == CVCS_F32) {
        int rc = persistent_grid(ce_generic_kernel<float>, kThreads, smem, &grid);
        if (rc) return rc;
        if (blocks_needed < grid) grid = static_cast<int>(blocks_needed < 1 ? 1 : blocks_needed);
        ce_generic_kernel<float><<<grid, kThreads, smem, stream>>>(p, class_stride, pixel_stride, image_stride);
    } else {
        int rc = persistent_grid(ce_generic_kernel<__nv_bfloat16>, kThreads, smem, &grid);
        if (rc) return rc;
        if (blocks_needed < grid) grid = static_cast<int>(blocks_needed < 1 ? 1 : blocks_needed);
        ce_generic_kernel<__nv_bfloat16><<<grid, kThreads, smem, stream>>>(p, class_stride, pixel_stride, image_stride);
    }
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

}  // namespace cvcs
