// ce_tma_impl.cuh — K1, TMA-staged variant (the primary path).
//
// A persistent, warp-specialised CTA per resident slot:
//   * one producer lane drives the bulk-copy engine: `cp.async.bulk` (TMA, 1-D) pulls a chunk of
//     P pixels — C class planes of P·esize bytes each for NCHW, or one contiguous P·C·esize span
//     for NHWC — plus the chunk's labels into a shared-memory stage, completing on an mbarrier;
//     finished stages are pushed back to HBM with `cp.async.bulk.global.shared::cta`.
//   * 256 consumer threads wait on the stage's mbarrier, read their VECP pixels from shared
//     memory with conflict-free vector loads, run the per-pixel softmax-CE / gradient / argmax
//     arithmetic in registers, overwrite the logits with the gradients IN PLACE, fence to the
//     async proxy and arrive on the stage's "done" mbarrier.
// Bytes in flight are set by the number of stages, not by registers or occupancy, so the
// memory pipeline stays full while each logit is read once and each gradient written once.
// A CTA handles chunks blockIdx.x, blockIdx.x + gridDim.x, ... (static round-robin).
//
// Requirements checked by the launcher: 16-byte aligned base pointers, H·W % 16 == 0 (NCHW)
// or B·H·W % 16 == 0 (NHWC) so that every bulk copy is a multiple of 16 bytes.
#pragma once
#include "ce_common.cuh"

namespace cvcs {
namespace tma {

constexpr int kProducerWarps = 1;
constexpr int kBlock = kThreads + 32 * kProducerWarps;
constexpr int kMaxStages = 8;
constexpr int kConsumerBar = 1;  // named barrier of the 256 consumer threads

struct Geom {
    int stages;
    int stage_bytes;   // logits + labels, multiple of 128
    int label_off;     // offset of the labels inside a stage
    int hist_off;      // offset of the bin accumulators in dynamic smem
    int stage_off;     // offset of stage 0
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global bulk copy, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- chunk addressing -----------------------------------------------------------------------
struct Chunk {
    long long pix0;    // global pixel index of the chunk's first pixel (b*hw + k*P for NCHW)
    long long elem0;   // NCHW: element offset of plane 0 (b*C*hw + k*P); NHWC: pix0 * C
    int n;             // valid pixels in the chunk
};

template <int C, int P, bool NHWC>
__device__ __forceinline__ Chunk chunk_of(const CeParams& p, long long q) {
    Chunk ck;
    if constexpr (NHWC) {
        ck.pix0 = q * P;
        ck.elem0 = ck.pix0 * C;
        const long long rem = p.n_pixels - ck.pix0;
        ck.n = rem < P ? static_cast<int>(rem) : P;
    } else {
        const unsigned int q32 = static_cast<unsigned int>(q);
        const unsigned int b = q32 / p.items_per_image;
        const unsigned int k = q32 - b * p.items_per_image;
        const long long in_img = static_cast<long long>(k) * P;
        ck.pix0 = static_cast<long long>(b) * p.hw + in_img;
        ck.elem0 = static_cast<long long>(b) * C * p.hw + in_img;
        const long long rem = p.hw - in_img;
        ck.n = rem < P ? static_cast<int>(rem) : P;
    }
    return ck;
}

template <typename T, int C, int VECP, bool NHWC, bool PRIV>
__global__ void __launch_bounds__(kBlock) ce_tma_kernel(const CeParams p, const Geom g) {
    constexpr int P = kThreads * VECP;
    constexpr int ES = sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float wsm[C];
    __shared__ __align__(8) unsigned long long bars[2 * kMaxStages];  // full[S], done[S]

    const int tid = threadIdx.x;
    const bool do_grad = p.dlogits != nullptr;
    const bool do_arg = p.argmax != nullptr;
    const bool do_conf = p.confmat != nullptr;
    const int S = g.stages;
    const int tsize = p.target_i64 ? 8 : 1;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem + g.stage_off);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar0 + 8 * s, 1);                       // full: producer's expect_tx arrive
            mbar_init(bar0 + 8 * (kMaxStages + s), kThreads);  // done: every consumer thread
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (tid < C) wsm[tid] = p.weight ? p.weight[tid] : 1.0f;
    __syncthreads();

    const long long n_chunks = p.n_items;
    const long long mine = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    double lsum = 0.0, wsum = 0.0;
    unsigned int bad = 0;

    if (tid >= kThreads) {
        // ================= producer =================
        if (tid == kThreads) {
            const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
            T* __restrict__ dlogits = reinterpret_cast<T*>(p.dlogits);
            const unsigned char* __restrict__ target = reinterpret_cast<const unsigned char*>(p.target);
            auto issue_load = [&](long long i) {
                const int s = static_cast<int>(i % S);
                const Chunk ck = chunk_of<C, P, NHWC>(p, blockIdx.x + i * gridDim.x);
                const uint32_t dst = stage0 + s * g.stage_bytes;
                const uint32_t bar = bar0 + 8 * s;
                const uint32_t lbytes = static_cast<uint32_t>(ck.n) * tsize;
                mbar_expect_tx(bar, static_cast<uint32_t>(ck.n) * C * ES + lbytes);
                if constexpr (NHWC) {
                    bulk_g2s(dst, logits + ck.elem0, static_cast<uint32_t>(ck.n) * C * ES, bar);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        bulk_g2s(dst + c * P * ES, logits + ck.elem0 + c * p.hw, static_cast<uint32_t>(ck.n) * ES, bar);
                }
                bulk_g2s(dst + g.label_off, target + ck.pix0 * tsize, lbytes, bar);
            };
            auto issue_store = [&](long long i) {
                const int s = static_cast<int>(i % S);
                const Chunk ck = chunk_of<C, P, NHWC>(p, blockIdx.x + i * gridDim.x);
                const uint32_t src = stage0 + s * g.stage_bytes;
                if constexpr (NHWC) {
                    bulk_s2g(dlogits + ck.elem0, src, static_cast<uint32_t>(ck.n) * C * ES);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        bulk_s2g(dlogits + ck.elem0 + c * p.hw, src + c * P * ES, static_cast<uint32_t>(ck.n) * ES);
                }
                bulk_commit();
            };
            const long long pre = mine < S ? mine : S;
            for (long long i = 0; i < pre; ++i) issue_load(i);
            for (long long i = 0; i < mine; ++i) {
                const int s = static_cast<int>(i % S);
                mbar_wait(bar0 + 8 * (kMaxStages + s), static_cast<uint32_t>((i / S) & 1));
                if (do_grad) {
                    issue_store(i);
                    // the stage consumed one step earlier is free once its store has left smem
                    if (i >= 1 && i - 1 + S < mine) {
                        bulk_wait_read<1>();
                        issue_load(i - 1 + S);
                    }
                } else if (i + S < mine) {
                    issue_load(i + S);
                }
            }
            if (do_grad) bulk_wait_all();
        }
    } else {
        // ================= consumers =================
        BinAcc<PRIV, kConsumerBar> conf;
        if (do_conf) conf.init(smem + g.hist_off, C * C);
        const float inv_tw = do_grad ? static_cast<float>(p.inv_tw_dev ? *p.inv_tw_dev : p.inv_tw) : 0.f;
        unsigned int since_flush = 0;

        for (long long i = 0; i < mine; ++i) {
            const int s = static_cast<int>(i % S);
            const Chunk ck = chunk_of<C, P, NHWC>(p, blockIdx.x + i * gridDim.x);
            unsigned char* stage = smem + g.stage_off + s * g.stage_bytes;
            mbar_wait(bar0 + 8 * s, static_cast<uint32_t>((i / S) & 1));
            if (tid * VECP < ck.n) {
                float x[VECP][C];
                // ---- shared -> registers
                if constexpr (NHWC) {
                    constexpr int EPV = 16 / ES;  // elements per 16-byte vector
                    static_assert((VECP * C) % EPV == 0, "NHWC run must be whole 16-byte vectors");
                    const uint4* src = reinterpret_cast<const uint4*>(stage + static_cast<size_t>(tid) * VECP * C * ES);
#pragma unroll
                    for (int j = 0; j < VECP * C / EPV; ++j) {
                        const uint4 v = src[j];
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < EPV; ++e) {
                            const int idx = j * EPV + e;
                            float f;
                            if constexpr (ES == 4) f = __uint_as_float(w[e]);
                            else f = (e & 1) ? bf16_hi(w[e / 2]) : bf16_lo(w[e / 2]);
                            x[idx / C][idx % C] = f;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const unsigned char* src = stage + (static_cast<size_t>(c) * P + tid * VECP) * ES;
                        constexpr int BYTES = VECP * ES;
                        uint32_t w[BYTES >= 4 ? BYTES / 4 : 1];
                        if constexpr (BYTES == 16) {
                            const uint4 v = *reinterpret_cast<const uint4*>(src);
                            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                        } else if constexpr (BYTES == 8) {
                            const uint2 v = *reinterpret_cast<const uint2*>(src);
                            w[0] = v.x; w[1] = v.y;
                        } else if constexpr (BYTES == 4) {
                            w[0] = *reinterpret_cast<const uint32_t*>(src);
                        } else {
                            w[0] = *reinterpret_cast<const unsigned short*>(src);
                        }
#pragma unroll
                        for (int k = 0; k < VECP; ++k) {
                            if constexpr (ES == 4) x[k][c] = __uint_as_float(w[k]);
                            else x[k][c] = (k & 1) ? bf16_hi(w[k / 2]) : bf16_lo(w[k / 2]);
                        }
                    }
                }
                // ---- labels
                int t[VECP];
                const unsigned char* lab = stage + g.label_off;
                if (p.target_i64) {
#pragma unroll
                    for (int k = 0; k < VECP; ++k) {
                        const uint2 v = *reinterpret_cast<const uint2*>(lab + (static_cast<size_t>(tid) * VECP + k) * 8);
                        t[k] = decode_label_i64(v.x, v.y, p.ignore_index);
                    }
                } else {
                    uint32_t w[VECP >= 4 ? VECP / 4 : 1];
                    if constexpr (VECP == 8) {
                        const uint2 v = *reinterpret_cast<const uint2*>(lab + tid * 8);
                        w[0] = v.x; w[1] = v.y;
                    } else if constexpr (VECP == 4) {
                        w[0] = *reinterpret_cast<const uint32_t*>(lab + tid * 4);
                    } else if constexpr (VECP == 2) {
                        w[0] = *reinterpret_cast<const unsigned short*>(lab + tid * 2);
                    } else {
                        w[0] = lab[tid];
                    }
                    decode_labels_u8<VECP>(w, p.ignore_index, t);
                }
                // ---- math
                int amax[VECP];
                float step_l = 0.f, step_w = 0.f;
#pragma unroll
                for (int k = 0; k < VECP; ++k) {
                    amax[k] = pixel_ce<C>(x[k], t[k], wsm, inv_tw, do_grad, step_l, step_w, bad);
                    if (do_conf && static_cast<unsigned int>(t[k]) < static_cast<unsigned int>(C)) conf.add(t[k] * C + amax[k]);
                }
                lsum += static_cast<double>(step_l);
                wsum += static_cast<double>(step_w);
                // ---- registers -> shared (in place)
                if (do_grad) {
                    if constexpr (NHWC) {
                        constexpr int EPV = 16 / ES;
                        uint4* dst = reinterpret_cast<uint4*>(stage + static_cast<size_t>(tid) * VECP * C * ES);
#pragma unroll
                        for (int j = 0; j < VECP * C / EPV; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if constexpr (ES == 4) {
                                    const int idx = j * 4 + e;
                                    w[e] = __float_as_uint(x[idx / C][idx % C]);
                                } else {
                                    const int i0 = j * 8 + 2 * e, i1 = i0 + 1;
                                    w[e] = pack_bf16(x[i0 / C][i0 % C], x[i1 / C][i1 % C]);
                                }
                            }
                            dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            unsigned char* dst = stage + (static_cast<size_t>(c) * P + tid * VECP) * ES;
                            constexpr int BYTES = VECP * ES;
                            uint32_t w[BYTES >= 4 ? BYTES / 4 : 1];
                            if constexpr (ES == 4) {
#pragma unroll
                                for (int k = 0; k < VECP; ++k) w[k] = __float_as_uint(x[k][c]);
                            } else if constexpr (VECP >= 2) {
#pragma unroll
                                for (int k = 0; k < VECP / 2; ++k) w[k] = pack_bf16(x[2 * k][c], x[2 * k + 1][c]);
                            } else {
                                w[0] = pack_bf16(x[0][c], 0.f);
                            }
                            if constexpr (BYTES == 16) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                            else if constexpr (BYTES == 8) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
                            else if constexpr (BYTES == 4) *reinterpret_cast<uint32_t*>(dst) = w[0];
                            else *reinterpret_cast<unsigned short*>(dst) = static_cast<unsigned short>(w[0]);
                        }
                    }
                }
                if (do_arg) store_argmax<VECP>(p, ck.pix0 + tid * VECP, amax);
            }
            if (do_grad) fence_async_smem();  // make the in-place gradients visible to the bulk store
            mbar_arrive(bar0 + 8 * (kMaxStages + s));
            if (PRIV && do_conf) {
                since_flush += VECP;
                if (since_flush > 65535u - VECP) {
                    conf.flush(p.confmat);
                    since_flush = 0;
                }
            }
        }
        if (do_conf) conf.flush(p.confmat);
    }
    finish_loss<kBlock / 32>(p, lsum, wsum, bad);
}

// ---- launch ------------------------------------------------------------------------------------
template <typename T, int C, int VECP, bool NHWC>
int launch(const CeParams& p0, cudaStream_t stream, bool* handled) {
    constexpr bool PRIV = C * C <= kPrivBinsMax;
    constexpr int P = kThreads * VECP;
    constexpr int ES = sizeof(T);
    CeParams p = p0;
    const int tsize = p.target_i64 ? 8 : 1;
    Geom g{};
    g.label_off = C * P * ES;
    g.stage_bytes = ((g.label_off + P * tsize + 127) / 128) * 128;
    g.hist_off = 0;
    const int hist_bytes = p.confmat ? BinAcc<PRIV>::smem_bytes(C * C) : 0;
    g.stage_off = ((hist_bytes + 127) / 128) * 128;
    // two CTAs per SM: each may use up to ~112 KB of the 227 KB; wide stages (i64 labels, large C)
    // that would leave fewer than 3 stages get the whole SM instead
    int stages = (113 * 1024 - g.stage_off) / g.stage_bytes;
    if (stages < 3) stages = (226 * 1024 - g.stage_off) / g.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    const int want_stages = get_option(CVCS_OPT_TMA_STAGES);
    if (want_stages >= 2 && want_stages <= kMaxStages && want_stages <= stages) stages = want_stages;
    if (stages < 2) {  // not even double-buffered: leave the shape to the direct / generic variants
        *handled = false;
        return CVCS_OK;
    }
    g.stages = stages;
    const int smem = g.stage_off + stages * g.stage_bytes;
    auto kernel = ce_tma_kernel<T, C, VECP, NHWC, PRIV>;
    int grid = 0;
    int rc = persistent_grid(kernel, kBlock, smem, &grid);
    if (rc) return rc;
    if (NHWC) {
        p.n_items = (p.n_pixels + P - 1) / P;
    } else {
        p.items_per_image = static_cast<unsigned int>((p.hw + P - 1) / P);
        p.n_items = static_cast<long long>(p.items_per_image) * (p.n_pixels / p.hw);
    }
    if (p.n_items < grid) grid = static_cast<int>(p.n_items < 1 ? 1 : p.n_items);
    kernel<<<grid, kBlock, smem, stream>>>(p, g);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

template <typename T, int VECP, bool NHWC, int CLO, int CHI, int CC = CLO>
int dispatch(const CeParams& p, cudaStream_t stream, bool* handled) {
    if constexpr (CC > CHI) {
        *handled = false;
        return CVCS_OK;
    } else {
        if (p.C == CC) {
            *handled = true;
            return launch<T, CC, VECP, NHWC>(p, stream, handled);
        }
        return dispatch<T, VECP, NHWC, CLO, CHI, CC + 1>(p, stream, handled);
    }
}

}  // namespace tma
}  // namespace cvcs
