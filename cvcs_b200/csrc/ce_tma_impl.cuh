// ce_tma_impl.cuh — K1, TMA-staged variant (the primary path).
//
// A persistent, warp-specialised CTA (256 consumer threads + a loader warp + a storer warp):
//   * the loader lane drives the bulk-copy engine: `cp.async.bulk` (TMA, 1-D) pulls a chunk of
//     P pixels — C class planes of P·esize bytes each for NCHW, or one contiguous P·C·esize span
//     for NHWC — plus the chunk's labels into a shared-memory stage, completing on the stage's
//     `full` mbarrier;
//   * 256 consumer threads wait on `full`, read their VECP pixels from shared memory with
//     conflict-free vector loads, run the per-pixel softmax-CE / gradient / argmax arithmetic in
//     registers, overwrite the logits with the gradients IN PLACE, fence to the async proxy and
//     arrive on the stage's `done` mbarrier;
//   * the storer lane waits on `done`, pushes the stage back to HBM with
//     `cp.async.bulk.global.shared::cta` and arrives on the stage's `free` mbarrier as soon as the
//     store has finished reading shared memory, which is what the loader waits for before refilling.
// Bytes in flight are set by the number of stages, not by registers or occupancy; each logit is
// read once and each gradient written once.  The geometry (stages per CTA, CTAs per SM) is chosen
// from measurements in launch().  A CTA handles chunks blockIdx.x, blockIdx.x + gridDim.x, ...
//
// Requirements checked by the launcher: 16-byte aligned base pointers, H·W % 16 == 0 (NCHW)
// or B·H·W % 16 == 0 (NHWC) so that every bulk copy is a multiple of 16 bytes.
#pragma once
#include "ce_common.cuh"

namespace cvcs {
namespace tma {

constexpr int kProducerWarps = 2;  // warp 8: bulk loads, warp 9: bulk stores
constexpr int kBlock = kThreads + 32 * kProducerWarps;
constexpr int kMaxStages = 8;
constexpr int kConsumerBar = 1;  // named barrier of the 256 consumer threads

struct Geom {
    int stages;
    int wait_hint;     // 1: mbarrier waits pass a long suspend-time hint
    int stage_bytes;   // logits + labels, multiple of 128
    int label_off;     // offset of the labels inside a stage
    int next_off;      // offset of the NEXT batch's label chunk inside a stage (0: not staged)
    int l2_hint;       // bit 0: staged next-batch labels evict_last (the next launch re-reads them: 1 B/px of HBM saved),
                       //        this batch's labels evict_first; bit 1: logits evict_first; bit 2: gradient stores evict_first
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <bool HINT>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread in hardware until the phase completes or a time limit expires;
    // HINT passes an explicit (long) limit so that waiting warps stay out of the issue slots
    if constexpr (HINT) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity), "r"(0x989680)
            : "memory");
    } else {
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
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// shared -> global bulk copy, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- chunk addressing -----------------------------------------------------------------------
// Chunks are claimed dynamically: a CTA starts with chunk blockIdx.x and its loader then draws
// gridDim.x + atomicAdd(next_chunk) — SMs that run faster (nearer L2 slices, fewer DRAM conflicts)
// simply take more chunks, so all CTAs finish within one chunk of each other.  The loader publishes
// the chunk's coordinates in a per-stage descriptor; consumers and storer read them after the
// stage's barrier, so only one thread per chunk does the index arithmetic.
struct Chunk {
    long long pix0;    // global pixel index of the chunk's first pixel (b*hw + k*P for NCHW)
    long long elem0;   // NCHW: element offset of plane 0 (b*C*hw + k*P); NHWC: pix0 * C
    int n;             // valid pixels in the chunk; < 0: no more chunks (sentinel)
    int pad;
};

template <int C, int P, bool NHWC>
__device__ __forceinline__ Chunk chunk_of(const CeParams& p, long long q) {
    Chunk ck;
    ck.pad = 0;
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

// ring position: stage index + phase parity, advanced without a modulo
struct Ring {
    int s;
    uint32_t phase;
    __device__ __forceinline__ void next(int S) {
        if (++s == S) {
            s = 0;
            phase ^= 1u;
        }
    }
};

template <typename T>
__device__ __forceinline__ float lds_elem(const unsigned char* base, int idx) {
    if constexpr (sizeof(T) == 4) return *reinterpret_cast<const float*>(base + 4 * idx);
    else return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(base + 2 * idx)) << 16);
}
template <typename T>
__device__ __forceinline__ void sts_elem(unsigned char* base, int idx, float v) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float*>(base + 4 * idx) = v;
    else *reinterpret_cast<unsigned short*>(base + 2 * idx) = static_cast<unsigned short>(pack_bf16(v, 0.f) & 0xffffu);
}

// ---- Σ v·w[y] inside K1 (tw_mode 1) --------------------------------------------------------------------------------
// nn.CrossEntropyLoss's 'mean' divides every gradient by the total weight of the WHOLE batch (utils.py:230,238), so it
// must be known before the first gradient is written.  Instead of a separate K4 launch the 256 consumer threads of
// every CTA first sum a 256-entry weight table over their slice of the byte labels (16.8 MB for cfg3: ~3 us, and the
// labels are then L2-resident for the main loop), the grid meets at one barrier (all CTAs are resident: persistent
// grid, cooperative launch), every CTA folds the per-CTA partials in the same fixed order, and with several GPUs one
// thread publishes the sum to every peer over NVLink and every CTA adds the ranks' values in rank order.  The loader
// warp is not involved: its first bulk loads are in flight meanwhile.  Called by threads 0..kThreads-1.
// Σ over this CTA's slice of n u8 labels of the 256-entry weight table; four 128-bit loads in flight per thread (the
// labels come from HBM: latency-bound).  The block total is returned to thread 0.  Threads 0..kThreads-1.
__device__ __forceinline__ double scan_label_weights(const float* wlut, double* red, const uint8_t* __restrict__ tgt, long long n, int tid) {
    const auto csync = [] { asm volatile("bar.sync %0, %1;" ::"n"(kConsumerBar), "n"(kThreads) : "memory"); };
    const long long n16 = n / 16;
    const long long per = (n16 + gridDim.x - 1) / gridDim.x;
    const long long lo = static_cast<long long>(blockIdx.x) * per;
    const long long hi = lo + per < n16 ? lo + per : n16;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int U = 4;
    for (long long base = lo + tid; base < hi; base += U * kThreads) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)     // default caching: the labels stay in L2 for the main loop
            v[u] = base + u * kThreads < hi ? *reinterpret_cast<const uint4*>(tgt + 16 * (base + u * kThreads)) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base + u * kThreads >= hi) continue;
            const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k & 3] += wlut[(w4[k / 4] >> (8 * (k % 4))) & 0xff];
        }
    }
    if (blockIdx.x == 0) {
        const long long i = n16 * 16 + tid;
        if (i < n) a[0] += wlut[tgt[i]];
    }
    double s = warp_sum(static_cast<double>((a[0] + a[1]) + (a[2] + a[3])));
    if ((tid & 31) == 0) red[tid >> 5] = s;
    csync();
    double b = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) b += red[w];
    }
    csync();
    return b;
}

template <int C>
__device__ __forceinline__ double prepass_total_weight(const CeParams& p, float* wlut, double* red, int tid) {
    __shared__ double total;
    const auto csync = [] { asm volatile("bar.sync %0, %1;" ::"n"(kConsumerBar), "n"(kThreads) : "memory"); };
    const double mine = scan_label_weights(wlut, red, reinterpret_cast<const uint8_t*>(p.target), p.n_pixels, tid);
    if (tid == 0) {
        p.ws->pre[0][blockIdx.x] = mine;
        __threadfence();
        atomicAdd(&p.ws->gbar, 1u);
        while (ld_acquire_gpu_u32(&p.ws->gbar) < gridDim.x) __nanosleep(32);
    }
    csync();
    if (tid < 32) {
        double v = 0.0;
        for (unsigned int i = tid; i < gridDim.x; i += 32) v += __ldcg(&p.ws->pre[0][i]);   // same order in every CTA
        v = warp_sum(v);
        if (p.xworld > 1) v = xchg_total_weight(p, v, blockIdx.x == 0);      // the whole warp
        if (tid == 0) {
            total = v;
            if (blockIdx.x == 0 && p.tw_out) {
                p.tw_out[0] = v;
                p.tw_out[1] = 1.0 / v;
            }
        }
    }
    csync();
    return total;
}

// tw_mode 2: this rank's Σ v·w[y] comes from a K4 launch (e.g. one step ahead on a side stream); only the cross-GPU
// exchange happens here, in the shadow of the pipeline fill.  Called by threads 0..kThreads-1.
__device__ __forceinline__ double exchanged_total_weight(const CeParams& p, int tid) {
    __shared__ double total;
    if (tid < 32) {
        double v = __ldcg(p.tw_local_dev);                                     // uniform over the warp
        if (p.xworld > 1) v = xchg_total_weight(p, v, blockIdx.x == 0);      // the whole warp
        if (tid == 0) {
            total = v;
            if (blockIdx.x == 0 && p.tw_out) {
                p.tw_out[0] = v;
                p.tw_out[1] = 1.0 / v;
            }
        }
    }
    asm volatile("bar.sync %0, %1;" ::"n"(kConsumerBar), "n"(kThreads) : "memory");
    return total;
}

// 256-entry weight table indexed by a label byte (0 for ignore_index and for anything >= C): static shared memory of
// the kernels that call it (gradient kernels with a label pre-pass / next-batch scan); a constant address, no register
__device__ __forceinline__ float* weight_lut() {
    __shared__ float lut[256];
    return lut;
}
__device__ __forceinline__ double* scan_scratch() {
    __shared__ double red[kWarps];
    return red;
}

template <int SUB>
struct HistCounter {
    using type = unsigned short;
};
template <>
struct HistCounter<2> {
    using type = unsigned char;
};

// SUB: sub-chunks of kThreads*VECP pixels per stage.  A consumer thread handles VECP consecutive pixels of every
// sub-chunk, so SUB > 1 doubles the bytes per bulk copy and halves the per-stage bookkeeping per pixel without
// widening the thread's register working set.
// Register budget: two gradient CTAs per SM (96 registers) or three forward-only ones (64).  The one shape whose thread
// holds more live values than that — NHWC bf16 with an 8-pixel span (C = 7: 56 logits, their labels, arg-maxima and
// target gradients) — spilled 328 bytes per thread at 96 registers and ran at 0.46 of the copy peak; those kernels are
// compiled for one CTA per SM (two forward-only) instead, and launch() reads the kernel's register count and sizes the
// stages for what can be resident.
template <typename T, int VECP, int C, bool NHWC, bool GRAD>
constexpr int min_ctas_per_sm() {
    constexpr bool wide = NHWC && sizeof(T) == 2 && VECP * C >= 48;
    return GRAD ? (wide ? 1 : 2) : (wide ? 2 : 3);
}

template <typename T, int C, int VECP, bool NHWC, bool PRIV, bool GRAD, bool LOSS = true, int SUB = 1>
__global__ void __launch_bounds__(kBlock, min_ctas_per_sm<T, VECP, C, NHWC, GRAD>()) ce_tma_kernel(const CeParams p, const Geom g) {
    constexpr int P = kThreads * VECP * SUB;
    constexpr int ES = sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float wsm[C];
    __shared__ __align__(8) unsigned long long bars[3 * kMaxStages];  // full[S], done[S], free[S]
    __shared__ __align__(16) Chunk desc[kMaxStages];                   // what each stage currently holds

    const int tid = threadIdx.x;
#ifdef CVCS_X_TIMING
    // experiment build: globaltimer stamps of CTAs 0..255 in ws->hist (start / first stage ready / last chunk done /
    // before the loss epilogue), the very end of the grid in hist[1024]; the caller re-zeroes the workspace
    unsigned long long t_start = 0, t_first = 0, t_loop = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
    constexpr bool do_grad = GRAD;
    const bool do_arg = p.argmax != nullptr;
    const bool do_conf = p.confmat != nullptr;
    const int S = g.stages;
    const int tsize = p.target_i64 ? 8 : 1;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem + g.stage_off);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar0 + 8 * s, 1);                       // full: loader's expect_tx arrive
            mbar_init(bar0 + 8 * (2 * kMaxStages + s), 1);    // free: the store warp, once the stage has left smem
            mbar_init(bar0 + 8 * (kMaxStages + s), kThreads);  // done: every consumer thread
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    // Programmatic dependent launch: the next kernel in the stream (e.g. the following K1) may become resident
    // as this grid's CTAs exit and run ITS prologue; everything above touches only shared memory, everything
    // below may read what the predecessor wrote (logits, 1/Σw from K4, the workspace counters) and waits for it.
    // Both are no-ops when the launch carries no programmatic-serialization attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();                                  // barriers initialised
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the loader goes straight to its first bulk loads; the consumers fetch the class weights and 1/Σw meanwhile

    const long long n_chunks = p.n_items;

    double lsum = 0.0, wsum = 0.0;
    unsigned int bad = 0;

    if (tid >= kThreads) {
        // ================= producers: one lane of warp 8 loads, one lane of warp 9 stores =================
        // A stage cycles  load -> full -> (consumers) -> done -> store -> free -> load ...  The two
        // roles never wait on each other's work except through the barriers, so a stage is refilled
        // as soon as its gradients have left shared memory, however long the consumers take.
        const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
        T* __restrict__ dlogits = reinterpret_cast<T*>(p.dlogits);
        const unsigned char* __restrict__ target = reinterpret_cast<const unsigned char*>(p.target);
        auto wait = [&](uint32_t bar, uint32_t phase) {
            if (g.wait_hint) mbar_wait<true>(bar, phase);
            else mbar_wait<false>(bar, phase);
        };
        if (tid == kThreads) {
            // ---- loader
            Ring ring{0, 0u};
            long long q = blockIdx.x;
            const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
            const bool hint_labels = (g.l2_hint & 1) != 0, hint_logits = (g.l2_hint & 2) != 0;
            for (long long i = 0;; ++i) {
                if (i >= S) {
                    // the stage's previous occupant: stored away (grad) / consumed (forward only)
                    const uint32_t prev = ring.phase ^ 1u;
                    if constexpr (do_grad) wait(bar0 + 8 * (2 * kMaxStages + ring.s), prev);
                    else wait(bar0 + 8 * (kMaxStages + ring.s), prev);
                }
                const uint32_t bar = bar0 + 8 * ring.s;
                if (q >= n_chunks) {          // nothing left: tell consumers and storer, then leave
                    desc[ring.s].n = -1;
                    mbar_arrive(bar);
                    break;
                }
                const Chunk ck = chunk_of<C, P, NHWC>(p, q);
                desc[ring.s] = ck;
                const uint32_t dst = stage0 + ring.s * g.stage_bytes;
                const uint32_t lbytes = static_cast<uint32_t>(ck.n) * tsize;
                const uint32_t nbytes = g.next_off ? static_cast<uint32_t>(ck.n) : 0u;   // the next batch's labels of the same pixels
                mbar_expect_tx(bar, static_cast<uint32_t>(ck.n) * C * ES + lbytes + nbytes);   // release: publishes desc
                if (hint_logits) {
                    if constexpr (NHWC) {
                        bulk_g2s_hint(dst, logits + ck.elem0, static_cast<uint32_t>(ck.n) * C * ES, bar, pol_first);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            bulk_g2s_hint(dst + c * P * ES, logits + ck.elem0 + c * p.hw, static_cast<uint32_t>(ck.n) * ES, bar, pol_first);
                    }
                } else {
                    if constexpr (NHWC) {
                        bulk_g2s(dst, logits + ck.elem0, static_cast<uint32_t>(ck.n) * C * ES, bar);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            bulk_g2s(dst + c * P * ES, logits + ck.elem0 + c * p.hw, static_cast<uint32_t>(ck.n) * ES, bar);
                    }
                }
                if (hint_labels) {
                    // the labels launch i stages for launch i+1's divisor are launch i+1's own labels: kept in L2 until
                    // then (evict_last), released by their last reader (evict_first)
                    bulk_g2s_hint(dst + g.label_off, target + ck.pix0 * tsize, lbytes, bar, pol_first);
                    if (nbytes) bulk_g2s_hint(dst + g.next_off, reinterpret_cast<const unsigned char*>(p.next_target) + ck.pix0, nbytes, bar, pol_last);
                } else {
                    bulk_g2s(dst + g.label_off, target + ck.pix0 * tsize, lbytes, bar);
                    if (nbytes) bulk_g2s(dst + g.next_off, reinterpret_cast<const unsigned char*>(p.next_target) + ck.pix0, nbytes, bar);
                }
                // claim the next chunk now: the atomic's round trip overlaps the wait for the next stage
                q = static_cast<long long>(gridDim.x) + atomicAdd(&p.ws->next_chunk, 1u);
                ring.next(S);
            }
        } else if (do_grad && tid == kThreads + 32) {
            // ---- storer: up to two bulk stores in flight; a stage is handed back to the loader as soon
            // as its store has finished reading shared memory — before blocking on the next `done`
            Ring ring{0, 0u};
            int pending = -1;  // stage whose store has been issued but not yet waited for
            const uint64_t pol_first = l2_policy_evict_first();
            const bool hint_grads = (g.l2_hint & 4) != 0;
            for (;;) {
                const uint32_t done = bar0 + 8 * (kMaxStages + ring.s);
                if (pending >= 0 && !mbar_test(done, ring.phase)) {
                    bulk_wait_read<0>();
                    mbar_arrive(bar0 + 8 * (2 * kMaxStages + pending));
                    pending = -1;
                }
                wait(done, ring.phase);
                const Chunk ck = desc[ring.s];
                if (ck.n < 0) break;
                const uint32_t src = stage0 + ring.s * g.stage_bytes;
                if (hint_grads) {
                    if constexpr (NHWC) {
                        bulk_s2g_hint(dlogits + ck.elem0, src, static_cast<uint32_t>(ck.n) * C * ES, pol_first);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            bulk_s2g_hint(dlogits + ck.elem0 + c * p.hw, src + c * P * ES, static_cast<uint32_t>(ck.n) * ES, pol_first);
                    }
                } else if constexpr (NHWC) {
                    bulk_s2g(dlogits + ck.elem0, src, static_cast<uint32_t>(ck.n) * C * ES);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        bulk_s2g(dlogits + ck.elem0 + c * p.hw, src + c * P * ES, static_cast<uint32_t>(ck.n) * ES);
                }
                bulk_commit();
                if (pending >= 0) {
                    bulk_wait_read<1>();                                  // the older store has left shared memory
                    mbar_arrive(bar0 + 8 * (2 * kMaxStages + pending));    // -> the loader may refill that stage
                }
                pending = ring.s;
                ring.next(S);
            }
            bulk_wait_read<0>();   // the last stores have left shared memory: the CTA may exit (the grid's end makes them visible)
#ifdef CVCS_X_TIMING
            unsigned long long t_st;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_st));
            atomicMax(&p.ws->hist[1025], t_st);
#endif
        }
    } else {
        // ================= consumers =================
        if (tid < C) wsm[tid] = p.weight ? p.weight[tid] : 1.0f;
        [[maybe_unused]] float next_acc = 0.f;   // Σ w over the next batch's labels this thread has seen
        float inv_tw = do_grad ? static_cast<float>(p.inv_tw_dev ? __ldcg(p.inv_tw_dev) : p.inv_tw) : 0.f;
        if constexpr (do_grad) {
            if (p.tw_mode == 1 || p.next_target) {
                float* wlut = weight_lut();
                double* scan_red = scan_scratch();
                wlut[tid] = (tid < C && static_cast<long long>(tid) != p.ignore_index) ? (p.weight ? p.weight[tid] : 1.0f) : 0.f;
                asm volatile("bar.sync %0, %1;" ::"n"(kConsumerBar), "n"(kThreads) : "memory");
                if (p.next_target && !g.next_off) {
                    // the NEXT batch's labels (a batch of another size: not staged with this one's chunks): summed now,
                    // while the loader fills the pipeline; published by the last CTA
                    const double nxt = scan_label_weights(wlut, scan_red, reinterpret_cast<const uint8_t*>(p.next_target), p.next_n, tid);
                    if (tid == 0) p.ws->pre[1][blockIdx.x] = nxt;      // made visible by the loss epilogue's release
                }

                if (p.tw_mode == 1) inv_tw = static_cast<float>(1.0 / prepass_total_weight<C>(p, wlut, scan_red, tid));   // grid-wide; see above
            }
            if (p.tw_mode == 2) inv_tw = static_cast<float>(1.0 / exchanged_total_weight(p, tid));
        }
        // private counters: 16-bit, or 8-bit when a stage holds several sub-chunks (half the shared memory, which buys a
        // third stage for two CTAs per SM; a CTA of a 16-tile bf16 batch never reaches 255 pixels per thread anyway)
        using Conf = BinAcc<PRIV, kConsumerBar, typename HistCounter<SUB>::type>;
        Conf conf;
        if (do_conf) conf.init(smem + g.hist_off, C * C);
        else Conf::sync();      // wsm visible to all consumers
        const int ign8 = ignore_as_int_u8(p.ignore_index);
        // byte-parallel label classification constants (u8 labels, C <= 128): see classify_labels_u8x4
        const uint32_t lab_ge_add = static_cast<uint32_t>(128 - C) * 0x01010101u;
        const uint32_t lab_ign4 = static_cast<uint32_t>(ign8 >= 0 ? ign8 : 0) * 0x01010101u;
        const uint32_t lab_ne_or = ign8 >= 0 ? 0u : 0x80808080u;   // no u8 label can equal an ignore_index outside [0, 255]
        unsigned int since_flush = 0;
        // ring position held as addresses (advanced by addition, no per-stage multiplies)
        int rs = 0;
        uint32_t rphase = 0u;
        uint32_t full_bar = bar0;
        unsigned char* stage = smem + g.stage_off;
        const Chunk* dsc = desc;

        for (;;) {
            mbar_wait<true>(full_bar, rphase);   // suspends in hardware (the option only affects the producers)
#ifdef CVCS_X_TIMING
            if (tid == 0 && t_first == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));
#endif
            const Chunk ck = *dsc;
            if (ck.n < 0) {                   // sentinel: pass it on to the storer and leave
                mbar_arrive(full_bar + 8 * kMaxStages);
                break;
            }
#pragma unroll 1
            for (int sub = 0; sub < SUB; ++sub) {
            const int pix_t0 = sub * (kThreads * VECP) + tid * VECP;  // this thread's first pixel of the sub-chunk
            if (pix_t0 < ck.n) {
                // element index of (class c, pixel j of the chunk) inside the stage
                auto eidx = [&](int c, int j) { return NHWC ? j * C + c : c * P + j; };
                // ---- shared -> registers: the thread's VECP pixels x C classes, kept RAW (logit
                // dtype) so that bf16 holds two values per register; converted pixel by pixel
                constexpr int WPP = NHWC ? 1 : VECP * ES / 4;                 // words per plane row (NCHW)
                constexpr int NW = NHWC ? VECP * C * ES / 4 : C * (WPP > 0 ? WPP : 1);
                static_assert(NHWC ? (VECP * C * ES) % 16 == 0 : (VECP * ES) % 4 == 0, "thread span must be whole words");
                uint32_t raw[NW];
                if constexpr (NHWC) {
                    const uint4* src = reinterpret_cast<const uint4*>(stage + static_cast<size_t>(pix_t0) * C * ES);
#pragma unroll
                    for (int j = 0; j < NW / 4; ++j) {
                        const uint4 v = src[j];
                        raw[4 * j] = v.x; raw[4 * j + 1] = v.y; raw[4 * j + 2] = v.z; raw[4 * j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const unsigned char* src = stage + (static_cast<size_t>(c) * P + pix_t0) * ES;
                        if constexpr (WPP == 4) {
                            const uint4 v = *reinterpret_cast<const uint4*>(src);
                            raw[4 * c] = v.x; raw[4 * c + 1] = v.y; raw[4 * c + 2] = v.z; raw[4 * c + 3] = v.w;
                        } else if constexpr (WPP == 2) {
                            const uint2 v = *reinterpret_cast<const uint2*>(src);
                            raw[2 * c] = v.x; raw[2 * c + 1] = v.y;
                        } else {
                            raw[c] = *reinterpret_cast<const uint32_t*>(src);
                        }
                    }
                }
                // flat element number of (class c, thread-local pixel k) inside raw[] (in elements of T)
                auto ridx = [&](int c, int k) { return NHWC ? k * C + c : c * VECP + k; };
                auto raw_get = [&](int e) -> float {
                    if constexpr (ES == 4) return __uint_as_float(raw[e]);
                    else return (e & 1) ? bf16_hi(raw[e >> 1]) : bf16_lo(raw[e >> 1]);
                };
                // ---- labels -> per pixel: the class clamped into [0, C) (what the target-logit gather uses), a
                // valid flag (in range and not ignore_index) and the count of out-of-bounds labels
                int tcl[VECP];
                bool valid[VECP];
                const unsigned char* lab = stage + g.label_off;
                if (p.target_i64) {
#pragma unroll
                    for (int k = 0; k < VECP; ++k) {
                        const uint2 v = *reinterpret_cast<const uint2*>(lab + (static_cast<size_t>(pix_t0) + k) * 8);
                        const int tv = decode_label_i64(v.x, v.y, p.ignore_index);
                        valid[k] = static_cast<unsigned int>(tv) < static_cast<unsigned int>(C);
                        tcl[k] = valid[k] ? tv : 0;
                        bad += (!valid[k] && tv != -1) ? 1u : 0u;
                    }
                } else if constexpr (VECP % 4 == 0 && C <= 128) {
                    // four labels per 32-bit word, classified byte-parallel: two flag words (bit 7 of each byte) give
                    // "not a valid class" and "out of bounds"; one predicate per pixel and one test per word remain
#pragma unroll
                    for (int j = 0; j < VECP / 4; ++j) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(lab + pix_t0 + 4 * j);
                        uint32_t inval7, bad7;
                        classify_labels_u8x4(w, lab_ge_add, lab_ign4, lab_ne_or, inval7, bad7);
                        if (bad7) bad += static_cast<unsigned int>(__popc(bad7));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            valid[4 * j + k] = (inval7 & (0x80u << (8 * k))) == 0u;
                            tcl[4 * j + k] = min(static_cast<int>((w >> (8 * k)) & 0xffu), C - 1);
                        }
                    }
                } else {
                    uint32_t w = 0;
                    if constexpr (VECP == 2) w = *reinterpret_cast<const unsigned short*>(lab + pix_t0);
                    else w = lab[pix_t0];
                    static_assert(VECP <= 2 || (VECP % 4 == 0 && C <= 128), "label decode");
#pragma unroll
                    for (int k = 0; k < VECP; ++k) {
                        const int v = (w >> (8 * k)) & 0xff;
                        const bool in_range = v < C;
                        valid[k] = in_range && v != ign8;
                        tcl[k] = min(v, C - 1);
                        bad += (!in_range && v != ign8) ? 1u : 0u;
                    }
                }
                if constexpr (do_grad) {
                    if (g.next_off) {
                        // the next batch's labels of these pixels arrived with the stage: four table look-ups per word
                        const unsigned char* nl = stage + g.next_off + pix_t0;
                        if constexpr (VECP % 4 == 0) {
#pragma unroll
                            for (int j = 0; j < VECP / 4; ++j) {
                                const uint32_t w = *reinterpret_cast<const uint32_t*>(nl + 4 * j);
                                const float* lut = weight_lut();
                                next_acc += (lut[w & 0xffu] + lut[(w >> 16) & 0xffu]) + (lut[(w >> 8) & 0xffu] + lut[w >> 24]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < VECP; ++k) next_acc += weight_lut()[nl[k]];
                        }
                    }
                }
                // ---- math.  fp32: one pixel at a time; bf16 NCHW: two pixels at a time, because a 32-bit
                // word of a class plane holds the pixel pair (2j, 2j+1) and both gradients are packed with one
                // cvt.rn.bf16x2 (the compiler interleaves the unrolled groups either way)
                constexpr int PP = (ES == 2 && VECP % 2 == 0) ? 2 : 1;
                constexpr bool kPackedPair = ES == 2 && PP == 2 && !NHWC;
                // byte address of (class c, thread-local pixel k) inside the stage: one multiply-add on the
                // thread's base (the target logit is gathered, and its gradient patched, by dynamic class index)
                constexpr int kClassStride = NHWC ? ES : P * ES;
                constexpr int kPixStride = NHWC ? C * ES : ES;
                unsigned char* const px0 = stage + static_cast<size_t>(pix_t0) * kPixStride;
                auto tptr = [&](int cls, int k) { return px0 + cls * kClassStride + k * kPixStride; };
                int amax[VECP];
                float gfix[VECP];     // gradient of the target class, patched into the stage afterwards
                float step_l = 0.f, step_w = 0.f;
                bool anomalous = false;  // some pixel's Σexp is NaN (NaN / +inf / all -inf logits)
#pragma unroll
                for (int k0 = 0; k0 < VECP; k0 += PP) {
                    if constexpr (kPackedPair) {
                        // ---- bf16 NCHW, a pixel pair per 32-bit word: everything that does not need fp32 stays on
                        // the packed words.  max = HMNMX2.BF16, first-max index = HSET2.EQ mask + one LOP3 per class
                        // (both pixels at once), and x - max goes straight from the packed half into fp32 with the
                        // mixed-precision add (sub.f32.bf16 -> FHADD.BF16, exact) — no unpack instructions at all.
                        uint32_t wq[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) wq[c] = raw[ridx(c, k0) >> 1];
                        uint32_t m2 = wq[0];
#pragma unroll
                        for (int c = 1; c < C; ++c) m2 = bf16x2_max(m2, wq[c]);
                        uint32_t arg2 = static_cast<uint32_t>(C - 1) * 0x00010001u;
#pragma unroll
                        for (int c = C - 2; c >= 0; --c)
                            arg2 = lop3_select(static_cast<uint32_t>(c) * 0x00010001u, arg2, bf16x2_eq_mask(wq[c], m2));
                        amax[k0] = static_cast<int>(arg2 & 0xffffu);
                        amax[k0 + 1] = static_cast<int>(arg2 >> 16);
                        if constexpr (!LOSS) {
                            // metrics mode: a NaN anywhere in the row (or +inf with -inf) makes the packed sum NaN
                            uint32_t s2 = wq[0];
#pragma unroll
                            for (int c = 1; c < C; ++c) s2 = bf16x2_add(s2, wq[c]);
                            anomalous |= bf16x2_nan_mask(s2) != 0u;
                        } else {
                            float e[2][C], r[2] = {0.f, 0.f};
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int k = k0 + q;
                                const unsigned short xtb = *reinterpret_cast<const unsigned short*>(tptr(tcl[k], k));
                                const float wc = wsm[tcl[k]];
                                const float w = valid[k] ? wc : 0.f;
                                const float m = q ? bf16_hi(m2) : bf16_lo(m2);
                                float s = 0.f;
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    e[q][c] = ex2_ftz(sub_f32_bf16(bf16_half(wq[c], q), m) * kLog2e);
                                    s = c ? s + e[q][c] : e[q][c];
                                }
                                anomalous |= (s != s);
                                const float dt = sub_f32_bf16(xtb, m);            // x_t - max <= 0, exact
                                const float nll = fmaf(lg2_ftz(s), kLn2, -dt);
                                step_l += valid[k] ? w * nll : 0.f;
                                step_w += w;
                                if constexpr (do_grad) {
                                    const float gsc = valid[k] ? w * inv_tw : 0.f;  // exact zeros at ignored pixels
                                    r[q] = gsc * rcp_ftz(s);
                                    gfix[k] = fmaf(ex2_ftz(dt * kLog2e), r[q], -gsc);
                                }
                            }
                            if constexpr (do_grad) {
#pragma unroll
                                for (int c = 0; c < C; ++c) raw[ridx(c, k0) >> 1] = pack_bf16(e[0][c] * r[0], e[1][c] * r[1]);
                            }
                        }
                        continue;
                    }
                    float x[PP][C];
                    float r[PP];
#pragma unroll
                    for (int q = 0; q < PP; ++q) {
                        const int k = k0 + q;
                        // the target logit and its class weight by dynamic index from shared memory
                        float xt = 0.f, w = 0.f;
                        if constexpr (LOSS) {
                            if constexpr (ES == 4) xt = *reinterpret_cast<const float*>(tptr(tcl[k], k));
                            else xt = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(tptr(tcl[k], k))) << 16);
                            w = wsm[tcl[k]];
                            w = valid[k] ? w : 0.f;
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) x[q][c] = raw_get(ridx(c, k));
                        float m, s;
                        int arg;
                        if constexpr (LOSS) {
                            softmax_core<C, ES == 2>(x[q], m, s, arg);
                        } else {
                            // metrics mode: only the argmax is wanted.  Σ x is NaN exactly when the row holds a NaN
                            // (or +inf and -inf together): those rows take the NaN-aware path below
                            m = x[q][0];
                            s = x[q][0];
#pragma unroll
                            for (int c = 1; c < C; ++c) {
                                m = fmaxf(m, x[q][c]);
                                s += x[q][c];
                            }
                            arg = C - 1;
#pragma unroll
                            for (int c = C - 2; c >= 0; --c) arg = (x[q][c] == m) ? c : arg;
                        }
                        anomalous |= (s != s);
                        amax[k] = arg;
                        if constexpr (LOSS) {
                            const float nll = fmaf(lg2_ftz(s), kLn2, m - xt);
                            step_l += valid[k] ? w * nll : 0.f;
                            step_w += w;
                        }
                        r[q] = 0.f;
                        if constexpr (do_grad) {
                            const float gsc = valid[k] ? w * inv_tw : 0.f;  // exact zeros at ignored pixels
                            r[q] = gsc * rcp_ftz(s);
                            gfix[k] = fmaf(exp_shifted<ES == 2>(xt, m), r[q], -gsc);
                        }
                    }
                    if constexpr (do_grad) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if constexpr (ES == 4) {
                                raw[ridx(c, k0)] = __float_as_uint(x[0][c] * r[0]);
                            } else if constexpr (PP == 2 && NHWC) {
                                // the pixel pair's 2C values are C consecutive words: word c holds flat elements 2c, 2c+1
                                auto flat = [&](int e) { return e < C ? x[0][e < C ? e : 0] * r[0] : x[PP - 1][e < C ? 0 : e - C] * r[PP - 1]; };
                                raw[(ridx(0, k0) >> 1) + c] = pack_bf16(flat(2 * c), flat(2 * c + 1));
                            } else {  // bf16, one pixel per thread: RNE, merged into the half of the word this element owns
                                const int e = ridx(c, k0);
                                const uint32_t h = pack_bf16(x[0][c] * r[0], 0.f) & 0xffffu;
                                raw[e >> 1] = (e & 1) ? ((raw[e >> 1] & 0x0000ffffu) | (h << 16)) : ((raw[e >> 1] & 0xffff0000u) | h);
                            }
                        }
                    }
                }
                lsum += static_cast<double>(step_l);
                wsum += static_cast<double>(step_w);
                // rows with NaN / inf (rare): redo the argmax with torch's NaN rule from the original
                // logits, which are still in the stage
                if (anomalous) {
#pragma unroll
                    for (int k = 0; k < VECP; ++k)
                        amax[k] = argmax_nan_aware<C>([&](int c) { return lds_elem<T>(stage, eidx(c, pix_t0 + k)); });
                }
                if (do_conf) {
                    if constexpr (PRIV) {
#pragma unroll
                        for (int k = 0; k < VECP; ++k)
                            if (valid[k]) conf.add(tcl[k] * C + amax[k]);
                    } else {
                        // shared bins (C*C > 64): one shared-memory atomic per pixel on this warp's replica
#pragma unroll
                        for (int k = 0; k < VECP; ++k)
                            if (valid[k]) conf.add(tcl[k] * C + amax[k]);
                    }
                }
                // ---- registers -> shared (in place), then the target-class entries
                if constexpr (do_grad) {
                    if constexpr (NHWC) {
                        uint4* dst = reinterpret_cast<uint4*>(stage + static_cast<size_t>(pix_t0) * C * ES);
#pragma unroll
                        for (int j = 0; j < NW / 4; ++j) dst[j] = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            unsigned char* dst = stage + (static_cast<size_t>(c) * P + pix_t0) * ES;
                            if constexpr (WPP == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(raw[4 * c], raw[4 * c + 1], raw[4 * c + 2], raw[4 * c + 3]);
                            else if constexpr (WPP == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(raw[2 * c], raw[2 * c + 1]);
                            else *reinterpret_cast<uint32_t*>(dst) = raw[c];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < VECP; ++k) {
                        if (valid[k]) {
                            if constexpr (ES == 4) *reinterpret_cast<float*>(tptr(tcl[k], k)) = gfix[k];
                            else *reinterpret_cast<unsigned short*>(tptr(tcl[k], k)) = static_cast<unsigned short>(pack_bf16(gfix[k], 0.f) & 0xffffu);
                        }
                    }
                }
                if (do_arg) store_argmax<VECP>(p, ck.pix0 + pix_t0, amax);
            }
            if (PRIV && do_conf) since_flush += VECP;
            }   // sub-chunks
            if constexpr (do_grad) fence_async_smem();  // make the in-place gradients visible to the bulk store
            mbar_arrive(full_bar + 8 * kMaxStages);
            if (PRIV && do_conf) {
                if (since_flush > Conf::kMaxCount - VECP * SUB) {
                    conf.flush(p.confmat);
                    since_flush = 0;
                }
            }
            ++dsc;
            full_bar += 8;
            stage += g.stage_bytes;
            if (++rs == S) {
                rs = 0;
                rphase ^= 1u;
                dsc = desc;
                full_bar = bar0;
                stage = smem + g.stage_off;
            }
        }
#ifdef CVCS_X_TIMING
        if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_loop));
#endif
        if constexpr (do_grad) {
            if (g.next_off) {
                __shared__ double next_red[kWarps];
                const double v = warp_sum(static_cast<double>(next_acc));
                if ((tid & 31) == 0) next_red[tid >> 5] = v;
                Conf::sync();
                if (tid == 0) {
                    double b = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) b += next_red[w];
                    p.ws->pre[1][blockIdx.x] = b;          // published by the loss epilogue's release
                }
            }
        }
        if (do_conf) conf.flush(p.confmat);
        // the loss epilogue runs on the consumer warps alone: it overlaps the store warp's last bulk stores
        finish_loss<kWarps, kConsumerBar>(p, lsum, wsum, bad);
#ifdef CVCS_X_TIMING
        if (tid == 0) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            if (blockIdx.x < 256) {
                p.ws->hist[blockIdx.x] = t_start;
                p.ws->hist[256 + blockIdx.x] = t_first;
                p.ws->hist[512 + blockIdx.x] = t_loop;
                p.ws->hist[768 + blockIdx.x] = t_end;
            }
            atomicMax(&p.ws->hist[1024], t_end);
        }
#endif
    }
}

// ---- launch ------------------------------------------------------------------------------------
template <typename T, int C, int VECP, bool NHWC, int SUB = 1>
int launch(const CeParams& p0, cudaStream_t stream, bool* handled) {
    constexpr bool PRIV = C * C <= kPrivBinsMax;
    constexpr int P = kThreads * VECP * SUB;
    constexpr int ES = sizeof(T);
    CeParams p = p0;
    const int tsize = p.target_i64 ? 8 : 1;
    Geom g{};
    g.label_off = C * P * ES;
    g.stage_bytes = ((g.label_off + P * tsize + 127) / 128) * 128;
    if (p.dlogits && p.next_target && p.next_n == p.n_pixels) {
        // the next batch has this batch's geometry: its label chunk rides in the stage (one more bulk copy per stage)
        g.next_off = g.stage_bytes;
        g.stage_bytes += ((P + 127) / 128) * 128;
    }
    g.hist_off = 0;
    const int hist_bytes = p.confmat ? BinAcc<PRIV, 0, typename HistCounter<SUB>::type>::smem_bytes(C * C) : 0;
    g.stage_off = ((hist_bytes + 127) / 128) * 128;
    // Pipeline geometry (measured on B200, profiles/README.md "K1 geometry, round 2").  With gradients the kernel wants
    // 4 stages per CTA (fp32 C=7, 1 CTA/SM: 3 stages 0.945, 4-6 stages 0.99 of the measured copy peak, 7 stages 0.945;
    // bf16 C=7, 2 CTAs/SM: 3 stages 0.81, 4-5 stages 0.835) and two CTAs per SM whenever each can still hold 3 stages
    // (bf16 is issue-limited with 8 consumer warps per SM: 1 CTA 0.59-0.72).  Forward-only kernels (64 registers) run
    // 3 CTAs per SM when each still holds 3 stages, else 2 (fp32 C=7: 2 CTAs x 3..6 stages all 0.93, 3 CTAs x 2: 0.69).
    const bool grad = p.dlogits != nullptr;
    auto kernel = p.dlogits ? ce_tma_kernel<T, C, VECP, NHWC, PRIV, true, true, SUB>
                            : (p.no_loss ? ce_tma_kernel<T, C, VECP, NHWC, PRIV, false, false, SUB> : ce_tma_kernel<T, C, VECP, NHWC, PRIV, false, true, SUB>);
    // static shared memory of the chosen kernel (barriers, descriptors, weight tables, reduction scratch): it counts
    // against the SM's 228 KB like the dynamic part does
    static thread_local int static_smem[3] = {-1, -1, -1};
    static thread_local int reg_ctas[3] = {1, 1, 1};       // CTAs per SM the kernel's register count allows
    const int kslot = p.dlogits ? 0 : (p.no_loss ? 1 : 2);
    if (static_smem[kslot] < 0) {
        cudaFuncAttributes fa;
        CVCS_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
        static_smem[kslot] = static_cast<int>(fa.sharedSizeBytes);
        const int regs = ((fa.numRegs + 7) / 8) * 8;
        reg_ctas[kslot] = regs > 0 ? 65536 / (regs * kBlock) : 1;
        if (reg_ctas[kslot] < 1) reg_ctas[kslot] = 1;
    }
    const int reserve = 1024 + ((static_smem[kslot] + 255) / 256) * 256;   // 1 KB per CTA taken by the system + static
    int target_ctas = get_option(CVCS_OPT_TMA_CTAS);
    const bool ctas_forced = target_ctas >= 1 && target_ctas <= 4;
    if (!ctas_forced) {
        if (grad) target_ctas = ((233472 / 2 - reserve - g.stage_off) / g.stage_bytes >= 3) ? 2 : 1;
        else target_ctas = ((233472 / 3 - reserve - g.stage_off) / g.stage_bytes >= 3) ? 3 : 2;  // forward-only kernels fit 3 CTAs (64 regs)
        if (target_ctas > reg_ctas[kslot]) target_ctas = reg_ctas[kslot];
    }
    const int per_cta = 233472 / target_ctas - reserve;  // 228 KB per SM
    int stages = (per_cta - g.stage_off) / g.stage_bytes;
    if ((stages < 3 && target_ctas > 1 && !ctas_forced) || stages < 2) {  // wide stages (i64 labels, large C): the whole SM
        target_ctas = 1;
        stages = (227 * 1024 - reserve + 1024 - g.stage_off) / g.stage_bytes;
    }
    if (grad && stages > 4) stages = 4;
    if (stages > kMaxStages) stages = kMaxStages;
    const int want_stages = get_option(CVCS_OPT_TMA_STAGES);
    if (want_stages >= 2 && want_stages <= kMaxStages && want_stages <= (per_cta - g.stage_off) / g.stage_bytes) stages = want_stages;
    if (stages < 2) {  // not even double-buffered: leave the shape to the direct / generic variants
        *handled = false;
        return CVCS_OK;
    }
    g.stages = stages;
    g.wait_hint = get_option(CVCS_OPT_TMA_WAIT_HINT) == 1 ? 0 : 1;
    {
        const int h = get_option(CVCS_OPT_L2_HINT);             // 0: default, v in 1..8: bits v - 1
        g.l2_hint = (h >= 1 && h <= 8) ? h - 1 : (g.next_off ? 1 : 0);
    }
    int smem = g.stage_off + stages * g.stage_bytes;
    // keep exactly `target_ctas` CTAs resident: the dynamic request is padded past what target + 1 could share
    // (more bytes in flight per SM than ~120 KB measurably lowers the sustained HBM rate)
    const int min_smem = 233472 / (target_ctas + 1) - 1024 - static_smem[kslot] + 16;
    if (smem < min_smem) smem = min_smem;
    int grid = 0;
    int rc = persistent_grid(kernel, kBlock, smem, &grid);
    if (rc) return rc;
    if (NHWC) {
        p.n_items = (p.n_pixels + P - 1) / P;
    } else {
        p.items_per_image = static_cast<unsigned int>((p.hw + P - 1) / P);
        p.n_items = static_cast<long long>(p.items_per_image) * (p.n_pixels / p.hw);
    }
    {   // leave whole SMs free when asked to (a concurrent collective needs somewhere to run); chunks are
        // claimed dynamically, so any grid size does the same work
        const int reserve = get_option(CVCS_OPT_RESERVE_SMS);
        const int sms = num_sms();
        if (reserve > 0 && reserve <= 32 && reserve < sms) grid -= (grid / sms) * reserve;
    }
    if (p.n_items < grid) grid = static_cast<int>(p.n_items < 1 ? 1 : p.n_items);
    if ((p.tw_mode == 1 || p.next_target) && grid > kMaxPreGrid) grid = kMaxPreGrid;
    if (p.tw_mode == 1) {
        // the pre-pass meets at a grid-wide barrier: every CTA must be resident -> cooperative launch (the grid is the
        // persistent one, sized from the occupancy query, so the launch is refused rather than deadlocked if it is not)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kBlock);
        cfg.dynamicSmemBytes = static_cast<size_t>(smem);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CVCS_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p, g));
    } else if (get_option(CVCS_OPT_PDL) != 2) {
        // programmatic stream serialization: this launch may start while the previous kernel of the stream drains
        // (the kernel itself waits, griddepcontrol.wait, before it reads anything the predecessor wrote)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kBlock);
        cfg.dynamicSmemBytes = static_cast<size_t>(smem);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CVCS_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p, g));
    } else {
        kernel<<<grid, kBlock, smem, stream>>>(p, g);
        CVCS_CUDA_OK(cudaGetLastError());
    }
    return CVCS_OK;
}

// NHWC: a thread's span of VECP pixels x C classes must be whole 16-byte vectors
template <typename T, int C, int VECP, bool NHWC>
constexpr bool span_ok() {
    return !NHWC || (VECP * C * static_cast<int>(sizeof(T))) % 16 == 0;
}

template <typename T, int VECP, bool NHWC, int CLO, int CHI, int SUB = 1, int CC = CLO>
int dispatch(const CeParams& p, cudaStream_t stream, bool* handled) {
    if constexpr (CC > CHI) {
        *handled = false;
        return CVCS_OK;
    } else {
        if constexpr (span_ok<T, CC, VECP, NHWC>()) {
            if (p.C == CC) {
                *handled = true;
                return launch<T, CC, VECP, NHWC, SUB>(p, stream, handled);
            }
        }
        return dispatch<T, VECP, NHWC, CLO, CHI, SUB, CC + 1>(p, stream, handled);
    }
}

}  // namespace tma
}  // namespace cvcs
