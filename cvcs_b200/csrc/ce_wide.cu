// ce_wide.cu — K1 for class counts beyond the register-resident range (C > 21), NCHW.
//
// The register kernels (ce_tma_impl.cuh) keep a pixel's C logits in one thread's registers, which stops at C = 21.  Here
// the same warp-specialised bulk-copy pipeline (loader lane: `cp.async.bulk` of C class-plane rows + the labels into a
// shared-memory stage; storer lane: the gradients back to HBM; 256 consumer threads in between) stages a chunk of P
// pixels, and a consumer thread walks the CLASS dimension of its pixel in shared memory: lane i of a warp reads pixel
// i of every plane row, so the reads are conflict-free whatever C is, and no cross-lane reduction is needed.  Three
// walks per pixel (max / first-max index, Σ exp, gradients written in place) cost 16·C bytes of shared-memory traffic per
// pixel against 8·C bytes of HBM traffic: the kernel stays HBM-bound.  Each logit is read from HBM once and each
// gradient written once; the generic kernel it replaces read every logit three times through the caches (C = 32: 0.44
// of the copy peak, C = 64: 0.26).
//
// Arithmetic and special-value rules are the generic kernel's, term for term (same summation order over the classes,
// torch's NaN-is-maximal argmax), so the two agree bit for bit.  Reference: nn.CrossEntropyLoss + torch.max +
// MulticlassConfusionMatrix.update, utils.py:223-242, 88-94; train.py:122-125.
#include "ce_tma_impl.cuh"

namespace cvcs {
namespace {
using namespace tma;

struct WideGeom {
    int stages;
    int P;             // pixels per chunk (a multiple of kThreads)
    int stage_bytes;   // C plane rows + labels, multiple of 128
    int label_off;
    int wsm_off;       // C class weights (float)
    int stage_off;
    int conf_reps;     // shared-memory replicas of the C x C bins (0: global atomics)
};

struct WideChunk {
    long long pix0;    // global pixel index of the chunk's first pixel
    long long elem0;   // element offset of plane 0
    int n;             // valid pixels; < 0: no more chunks
    int pad;
};

__device__ __forceinline__ WideChunk wide_chunk_of(const CeParams& p, int P, long long q) {
    WideChunk ck;
    ck.pad = 0;
    const unsigned int q32 = static_cast<unsigned int>(q);
    const unsigned int b = q32 / p.items_per_image;
    const unsigned int k = q32 - b * p.items_per_image;
    const long long in_img = static_cast<long long>(k) * P;
    ck.pix0 = static_cast<long long>(b) * p.hw + in_img;
    ck.elem0 = static_cast<long long>(b) * p.C * p.hw + in_img;
    const long long rem = p.hw - in_img;
    ck.n = rem < P ? static_cast<int>(rem) : P;
    return ck;
}

template <typename T, bool GRAD>
__global__ void __launch_bounds__(kBlock, 2) ce_wide_kernel(const CeParams p, const WideGeom g) {
    constexpr int ES = sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[3 * kMaxStages];  // full[S], done[S], free[S]
    __shared__ __align__(16) WideChunk desc[kMaxStages];

    const int tid = threadIdx.x;
    const int C = p.C, P = g.P, S = g.stages;
    const int tsize = p.target_i64 ? 8 : 1;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem + g.stage_off);
    const bool do_conf = p.confmat != nullptr;
    const bool do_loss = !p.no_loss;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (2 * kMaxStages + s), 1);
            mbar_init(bar0 + 8 * (kMaxStages + s), kThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();

    const long long n_chunks = p.n_items;
    double lsum = 0.0, wsum = 0.0;
    unsigned int bad = 0;

    if (tid >= kThreads) {
        // ================= producers: warp 8 loads, warp 9 stores — ALL lanes issue =================
        // A stage here is C bulk copies of a class-plane row (1-2 KB each), not the 8 copies of 4 KB of the register
        // kernels: one issuing lane cannot keep up (measured: ~80 cycles per copy, C = 32 forward-only stuck at 0.57
        // of the copy peak).  Lane l issues rows l, l + 32, ...; lane 0 does the waiting, the chunk claim and the
        // barrier arrivals, and the warp moves in step (__syncwarp).  Bulk async-groups are per thread, so every
        // storing lane commits and waits for its own group before lane 0 hands the stage back.
        const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
        T* __restrict__ dlogits = reinterpret_cast<T*>(p.dlogits);
        const unsigned char* __restrict__ target = reinterpret_cast<const unsigned char*>(p.target);
        const int lane = tid & 31;
        if (tid < kThreads + 32) {
            // ---- loader warp
            Ring ring{0, 0u};
            unsigned int q = blockIdx.x;
            for (long long i = 0;; ++i) {
                if (i >= S && lane == 0) {
                    const uint32_t prev = ring.phase ^ 1u;
                    if constexpr (GRAD) mbar_wait<true>(bar0 + 8 * (2 * kMaxStages + ring.s), prev);
                    else mbar_wait<true>(bar0 + 8 * (kMaxStages + ring.s), prev);
                }
                __syncwarp();
                const uint32_t bar = bar0 + 8 * ring.s;
                if (static_cast<long long>(q) >= n_chunks) {
                    if (lane == 0) {
                        desc[ring.s].n = -1;
                        mbar_arrive(bar);
                    }
                    break;
                }
                const WideChunk ck = wide_chunk_of(p, P, q);
                const uint32_t dst = stage0 + ring.s * g.stage_bytes;
                const uint32_t row = static_cast<uint32_t>(ck.n) * ES;
                const uint32_t lbytes = static_cast<uint32_t>(ck.n) * tsize;
                if (lane == 0) {
                    desc[ring.s] = ck;
                    mbar_expect_tx(bar, row * C + lbytes);          // release: publishes desc
                }
                __syncwarp();
                for (int c = lane; c < C; c += 32) bulk_g2s(dst + c * P * ES, logits + ck.elem0 + c * p.hw, row, bar);
                unsigned int qn = 0u;
                if (lane == 0) {
                    bulk_g2s(dst + g.label_off, target + ck.pix0 * tsize, lbytes, bar);
                    qn = gridDim.x + atomicAdd(&p.ws->next_chunk, 1u);
                }
                q = __shfl_sync(0xffffffffu, qn, 0);
                ring.next(S);
            }
        } else if (GRAD) {
            // ---- storer warp
            Ring ring{0, 0u};
            int pending = -1;
            for (;;) {
                const uint32_t done = bar0 + 8 * (kMaxStages + ring.s);
                int early = 0;
                if (lane == 0) early = (pending >= 0 && !mbar_test(done, ring.phase)) ? 1 : 0;
                early = __shfl_sync(0xffffffffu, early, 0);
                if (early) {
                    bulk_wait_read<0>();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (2 * kMaxStages + pending));
                    pending = -1;
                }
                if (lane == 0) mbar_wait<true>(done, ring.phase);
                __syncwarp();
                const WideChunk ck = desc[ring.s];
                if (ck.n < 0) break;
                const uint32_t src = stage0 + ring.s * g.stage_bytes;
                const uint32_t row = static_cast<uint32_t>(ck.n) * ES;
                for (int c = lane; c < C; c += 32) bulk_s2g(dlogits + ck.elem0 + c * p.hw, src + c * P * ES, row);
                bulk_commit();
                if (pending >= 0) {
                    bulk_wait_read<1>();                               // this lane's older group has left shared memory
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (2 * kMaxStages + pending));
                }
                pending = ring.s;
                ring.next(S);
            }
            bulk_wait_read<0>();
        }
        return;
    }

    // ================= consumers =================
    float* wsm = reinterpret_cast<float*>(smem + g.wsm_off);
    for (int c = tid; c < C; c += kThreads) wsm[c] = p.weight ? p.weight[c] : 1.0f;
    const float inv_tw = GRAD ? static_cast<float>(p.inv_tw_dev ? __ldcg(p.inv_tw_dev) : p.inv_tw) : 0.f;
    using Conf = BinAcc<false, kConsumerBar>;
    Conf conf;
    if (do_conf) conf.init(smem, C * C, g.conf_reps, p.confmat);
    else Conf::sync();
    const int ign8 = ignore_as_int_u8(p.ignore_index);
    const int plane = P * ES;     // bytes between the class rows of a stage

    Ring ring{0, 0u};
    for (;;) {
        mbar_wait<true>(bar0 + 8 * ring.s, ring.phase);
        const WideChunk ck = desc[ring.s];
        if (ck.n < 0) {
            mbar_arrive(bar0 + 8 * (kMaxStages + ring.s));   // pass the sentinel on to the storer
            break;
        }
        unsigned char* stage = smem + g.stage_off + static_cast<size_t>(ring.s) * g.stage_bytes;
        const unsigned char* lab = stage + g.label_off;
        for (int px = tid; px < ck.n; px += kThreads) {
            unsigned char* col = stage + static_cast<size_t>(px) * ES;     // this pixel's entry of class row 0
            // ---- label: [0, C) valid class, -1 ignored, anything else out of bounds
            int tv;
            if (p.target_i64) {
                const uint2 v = *reinterpret_cast<const uint2*>(lab + static_cast<size_t>(px) * 8);
                tv = decode_label_i64(v.x, v.y, p.ignore_index);
            } else {
                const int v = lab[px];
                tv = (v == ign8) ? -1 : v;
            }
            const bool valid = static_cast<unsigned int>(tv) < static_cast<unsigned int>(C);
            bad += (!valid && tv != -1) ? 1u : 0u;
            // ---- walk 1: max (for the shift) and first-max index.  x * 0 summed over the row is NaN exactly when the row
            // holds a NaN or an infinity: those rows (rare) redo the walk with torch's NaN-is-maximal rule and take the
            // recomputing form of walks 2 / 3 below, which is the generic kernel's arithmetic on the original logits
            float best = lds_elem<T>(col, 0);
            int arg = 0;
            float chk = best * 0.f;
#pragma unroll 4
            for (int c = 1; c < C; ++c) {
                const float x = lds_elem<T>(col + static_cast<size_t>(c) * plane, 0);
                chk = fmaf(x, 0.f, chk);
                if (x > best) {
                    best = x;
                    arg = c;
                }
            }
            float m = best;
            const bool special = chk != chk;
            if (special) {
                best = lds_elem<T>(col, 0);
                m = best;
                arg = 0;
                for (int c = 1; c < C; ++c) {
                    const float x = lds_elem<T>(col + static_cast<size_t>(c) * plane, 0);
                    if (better(x, best)) {
                        best = x;
                        arg = c;
                    }
                    m = fmaxf(m, x);
                }
            }
            if (do_loss) {
                const float xt = valid ? lds_elem<T>(col + static_cast<size_t>(tv) * plane, 0) : 0.f;
                const float w = valid ? wsm[tv] : 0.f;
                // fp32 with gradients: walk 2 leaves exp(x - m) in place of x, walk 3 only scales it (the value the
                // recomputing form would produce, bit for bit, with one MUFU and four instructions fewer per element);
                // bf16 keeps its logits until walk 3, because a bf16 copy of exp(x - m) would be rounded twice
                const bool in_place = GRAD && ES == 4 && !special;
                // ---- walk 2: Σ exp(x - m), in class order
                float s = 0.f;
                if (in_place) {
#pragma unroll 4
                    for (int c = 0; c < C; ++c) {
                        unsigned char* e = col + static_cast<size_t>(c) * plane;
                        const float v = ex2_ftz((lds_elem<T>(e, 0) - m) * kLog2e);
                        s += v;
                        sts_elem<T>(e, 0, v);
                    }
                } else {
#pragma unroll 4
                    for (int c = 0; c < C; ++c) s += ex2_ftz((lds_elem<T>(col + static_cast<size_t>(c) * plane, 0) - m) * kLog2e);
                }
                const float nll = fmaf(lg2_ftz(s), kLn2, m - xt);
                lsum += valid ? static_cast<double>(w * nll) : 0.0;
                wsum += static_cast<double>(w);
                if constexpr (GRAD) {
                    // ---- walk 3: gradients in place
                    const float gsc = valid ? w * inv_tw : 0.f;      // exact zeros at ignored pixels
                    const float r = gsc * rcp_ftz(s);
                    const float gt = fmaf(ex2_ftz((xt - m) * kLog2e), r, -gsc);
                    if (in_place) {
#pragma unroll 4
                        for (int c = 0; c < C; ++c) {
                            unsigned char* e = col + static_cast<size_t>(c) * plane;
                            sts_elem<T>(e, 0, fmaf(lds_elem<T>(e, 0), r, 0.f));
                        }
                    } else {
#pragma unroll 4
                        for (int c = 0; c < C; ++c) {
                            unsigned char* e = col + static_cast<size_t>(c) * plane;
                            sts_elem<T>(e, 0, fmaf(ex2_ftz((lds_elem<T>(e, 0) - m) * kLog2e), r, 0.f));
                        }
                    }
                    if (valid) sts_elem<T>(col + static_cast<size_t>(tv) * plane, 0, gt);
                }
            }
            if (p.argmax) {
                int a[1] = {arg};
                store_argmax<1>(p, ck.pix0 + px, a);
            }
            if (do_conf && valid) conf.add(tv * C + arg);
        }
        if constexpr (GRAD) fence_async_smem();   // in-place gradients visible to the bulk store
        mbar_arrive(bar0 + 8 * (kMaxStages + ring.s));
        ring.next(S);
    }
    if (do_conf) conf.flush(p.confmat);
    finish_loss<kWarps, kConsumerBar>(p, lsum, wsum, bad);
}

template <typename T>
int launch_wide(const CeParams& p0, cudaStream_t stream, bool* handled) {
    constexpr int ES = sizeof(T);
    CeParams p = p0;
    const int C = p.C;
    const int tsize = p.target_i64 ? 8 : 1;
    const bool grad = p.dlogits != nullptr;
    auto kernel = grad ? ce_wide_kernel<T, true> : ce_wide_kernel<T, false>;
    static thread_local int static_smem[2] = {-1, -1};
    if (static_smem[grad] < 0) {
        cudaFuncAttributes fa;
        CVCS_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
        static_smem[grad] = static_cast<int>(fa.sharedSizeBytes);
    }
    // Geometry: the consumers are issue / latency limited (three walks of dependent shared-memory accesses), so two CTAs
    // per SM — 16 consumer warps — come first whenever three stages of 256 pixels fit in half an SM's shared memory;
    // otherwise one CTA with 512- or 256-pixel stages.  The C x C bins give way first (fewer replicas).
    const int budget_full = 227 * 1024 - static_smem[grad] - 256;
    const int budget_half = 233472 / 2 - 1024 - static_smem[grad] - 256;
    WideGeom g{};
    const int wsm_bytes = ((C * 4 + 127) / 128) * 128;
    bool found = false;
    struct Try { int budget, P, min_stages; };
    const Try tries[] = {{budget_half, 256, 3}, {budget_full, 512, 3}, {budget_full, 256, 3}, {budget_full, 512, 2}, {budget_full, 256, 2}};
    for (const Try& t : tries) {
        int reps = p.confmat ? shared_bin_replicas(C * C, 32 * 1024, 64 * 1024) : 0;
        for (;;) {
            const int hist = p.confmat ? ((C * C * reps * 4 + 127) / 128) * 128 : 0;
            const int stage_bytes = ((C * t.P * ES + t.P * tsize + 127) / 128) * 128;
            const int stages = (t.budget - hist - wsm_bytes) / stage_bytes;
            if (stages >= t.min_stages) {
                g.P = t.P;
                g.stage_bytes = stage_bytes;
                g.label_off = C * t.P * ES;
                g.wsm_off = hist;
                g.stage_off = hist + wsm_bytes;
                g.conf_reps = reps;
                g.stages = stages > 4 ? 4 : stages;
                found = true;
                break;
            }
            if (reps <= 1) break;
            reps >>= 1;
        }
        if (found) break;
    }
    if (!found) {
        *handled = false;          // a single class-plane stage does not fit twice: the generic kernel takes it
        return CVCS_OK;
    }
    *handled = true;
    const int smem = g.stage_off + g.stages * g.stage_bytes;
    int grid = 0;
    int rc = persistent_grid(kernel, kBlock, smem, &grid);
    if (rc) return rc;
    p.items_per_image = static_cast<unsigned int>((p.hw + g.P - 1) / g.P);
    p.n_items = static_cast<long long>(p.items_per_image) * (p.n_pixels / p.hw);
    if (p.n_items < grid) grid = static_cast<int>(p.n_items < 1 ? 1 : p.n_items);
    kernel<<<grid, kBlock, smem, stream>>>(p, g);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

}  // namespace

// NCHW, C > kMaxRegC, every class-plane row of a chunk a multiple of 16 bytes at a 16-byte aligned address (checked by
// the caller: aligned base pointers, H*W % 16 == 0).  *handled = false when not even two stages fit (C beyond ~200 for
// fp32): the generic kernel remains the path for those.
int ce_wide_launch(const CeParams& p, int logits_dtype, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (p.C * p.C > 16384) return CVCS_OK;      // C x C bins beyond one 64 KB replica
    return logits_dtype == CVCS_F32 ? launch_wide<float>(p, stream, handled) : launch_wide<__nv_bfloat16>(p, stream, handled);
}

}  // namespace cvcs
