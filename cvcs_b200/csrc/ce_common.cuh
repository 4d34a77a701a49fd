// ce_common.cuh — device code shared by the K1 variants (direct-load and TMA-staged):
// the per-pixel softmax-CE / gradient / argmax arithmetic and the deterministic loss epilogue.
//   loss   = Σ v w[y] (lse(x) - x[y]) / Σ v w[y]        nn.CrossEntropyLoss(weight, ignore_index),
//                                                        reference utils.py:223-242, train.py:122
//   dlogit = v w[y] (softmax(x) - onehot(y)) / Σ v w[y]  loss.backward(), train.py:125
//   argmax = first maximal class, NaN is maximal         torch.max(y_pred, dim=0), utils.py:90
#pragma once
#include "common.cuh"

namespace cvcs {

struct CeParams {
    const void* logits;
    const void* target;
    const float* weight;
    void* dlogits;
    void* argmax;
    unsigned long long* confmat;
    const double* inv_tw_dev;
    double inv_tw;
    double* loss_sums;
    float* loss_out;
    Workspace* ws;
    long long ignore_index;
    long long hw;                  // pixels per image
    long long n_pixels;            // B * hw
    long long n_items;             // direct: B*hw/VEC work items; tma: number of chunks
    unsigned int items_per_image;  // direct: hw/VEC; tma (NCHW): chunks per image
    int C;
    int target_i64;  // 0: u8, 1: i64
    int argmax_i64;  // 0: u8, 1: i64
    int conf_reps;   // generic kernel: shared-memory replicas of the C*C bins (0 = global atomics)
    unsigned long long* status;  // nullable: += #out-of-bounds labels (metrics mode, cvcs_eval_fused)
    int no_loss;     // 1: metrics mode — argmax + confusion matrix only, no softmax / loss (loss_sums may be NULL)
    // total weight computed by the kernel itself (TMA variant): a label pre-pass by every CTA, a grid-wide barrier
    // and — across GPUs — a one-shot exchange over peer-mapped memory, all before the first gradient is written
    int tw_mode;                 // 0: inv_tw / inv_tw_dev above; 1: in-kernel label pre-pass (+ exchange);
                                 // 2: this rank's Σ given in tw_local_dev (a K4 launch), exchange in the kernel
    const double* tw_local_dev;  // tw_mode 2: f64[1]
    double* tw_out;              // nullable f64[2]: {Σ v·w[y] (global), 1/Σ}
    // software pipelining across launches: this launch also sums the weights over the NEXT batch's (u8) labels — in its
    // prologue, while the first bulk loads are in flight — and its last CTA publishes {Σ, 1/Σ} for the next launch
    const void* next_target;     // nullable
    long long next_n;            // pixels in the next batch
    double* next_tw_out;         // f64[2]
    int xworld, xrank;           // ranks taking part in the exchange (1 = this GPU only)
    XchgBlock* xpeer[kXMaxRanks];  // every rank's exchange block (xpeer[xrank] is the local one)
};

// what cvcs_ce_fused_tw asks of the launcher: compute the total weight in the kernel, optionally exchanged across ranks
struct TwRequest {
    const void* next_target;       // nullable: u8 labels of the NEXT batch, scanned by this launch
    long long next_n;
    double* next_tw_out;           // f64[2] {Σ v·w[y] of the next batch (this rank), 1/Σ}
    const double* tw_local;        // nullable: this rank's Σ v·w[y] already on the device (skips the label pre-pass)
    double* tw_out;                // nullable f64[2]
    int world, rank;
    XchgBlock* peer[kXMaxRanks];   // peer[rank] = local block; unused when world == 1
};

// launchers (one translation unit each)
int ce_direct_launch(const CeParams& p, int logits_dtype, int vec, cudaStream_t stream, bool* handled);
int ce_tma_launch(const CeParams& p, int logits_dtype, int layout, cudaStream_t stream, bool* handled);
int ce_generic_launch(const CeParams& p, int logits_dtype, int layout, cudaStream_t stream);

constexpr int kPrivBinsMax = 64;  // C*C <= 64 -> per-thread private u16 counters
constexpr int kMaxRegC = 21;      // largest C with register-resident instantiations

#ifdef __CUDACC__

__device__ __forceinline__ int ignore_as_int_u8(long long ignore_index) {
    return (ignore_index >= 0 && ignore_index <= 255) ? static_cast<int>(ignore_index) : -1000;
}

// Label encoding used by all K1 kernels: [0,C) valid class, -1 ignored, anything else
// (>= C or -2) out of bounds.
template <int VEC>
__device__ __forceinline__ void decode_labels_u8(const uint32_t* words, long long ignore_index, int (&t)[VEC]) {
    const int ign = ignore_as_int_u8(ignore_index);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int v = (words[i / 4] >> (8 * (i % 4))) & 0xff;
        t[i] = (v == ign) ? -1 : v;
    }
}
__device__ __forceinline__ int decode_label_i64(uint32_t lo, uint32_t hi, long long ignore_index) {
    const long long a = (static_cast<long long>(hi) << 32) | lo;
    return a == ignore_index ? -1 : ((a < 0 || a > 0x7fffffff) ? -2 : static_cast<int>(a));
}

template <int VEC>
__device__ __forceinline__ void load_targets(const CeParams& p, long long pix, int (&t)[VEC]) {
    if (p.target_i64) {
        load_labels_i64<VEC>(reinterpret_cast<const long long*>(p.target) + pix, p.ignore_index, -1, t);
    } else {
        load_labels_u8<VEC>(reinterpret_cast<const uint8_t*>(p.target) + pix, t);
        const int ign = ignore_as_int_u8(p.ignore_index);
#pragma unroll
        for (int i = 0; i < VEC; ++i) t[i] = (t[i] == ign) ? -1 : t[i];
    }
}

template <int VEC>
__device__ __forceinline__ void store_argmax(const CeParams& p, long long pix, const int (&a)[VEC]) {
    if (p.argmax_i64) {
        long long* out = reinterpret_cast<long long*>(p.argmax) + pix;
        if constexpr (VEC % 2 == 0) {
#pragma unroll
            for (int i = 0; i < VEC / 2; ++i) {
                Raw<16> r;
                r.v = make_uint4(static_cast<uint32_t>(a[2 * i]), 0u, static_cast<uint32_t>(a[2 * i + 1]), 0u);
                r.store(out + 2 * i);
            }
        } else {
            Raw<8> r;
            r.v = make_uint2(static_cast<uint32_t>(a[0]), 0u);
            r.store(out);
        }
    } else {
        uint8_t* out = reinterpret_cast<uint8_t*>(p.argmax) + pix;
        if constexpr (VEC == 1) {
            Raw<1> r;
            r.v = static_cast<uint8_t>(a[0]);
            r.store(out);
        } else if constexpr (VEC == 2) {
            Raw<2> r;
            r.v = static_cast<uint16_t>(a[0] | (a[1] << 8));
            r.store(out);
        } else {
            Raw<VEC> r;
#pragma unroll
            for (int i = 0; i < VEC / 4; ++i)
                r.word(i) = static_cast<uint32_t>(a[4 * i]) | (static_cast<uint32_t>(a[4 * i + 1]) << 8) |
                            (static_cast<uint32_t>(a[4 * i + 2]) << 16) | (static_cast<uint32_t>(a[4 * i + 3]) << 24);
            r.store(out);
        }
    }
}

// Four u8 labels of one 32-bit word classified at once (C <= 128).  Results carry one flag per byte in bit 7:
//   inval7: the label is not a valid class (>= C, or equal to ignore_index)
//   bad7  : the label is out of bounds (>= C and not ignore_index)
// ge_add = (128 - C) * 0x01010101; ign4 = the ignore byte replicated (0 if ignore_index is not a byte value) and
// ne_or = 0x80808080 in that case (no label can match), else 0.
__device__ __forceinline__ void classify_labels_u8x4(uint32_t w, uint32_t ge_add, uint32_t ign4, uint32_t ne_or,
                                                     uint32_t& inval7, uint32_t& bad7) {
    const uint32_t ge = ((w & 0x7f7f7f7fu) + ge_add) | w;                       // bit 7: byte >= C
    const uint32_t z = w ^ ign4;
    const uint32_t ne = (((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) | ne_or;        // bit 7: byte != ignore
    bad7 = ge & ne & 0x80808080u;
    inval7 = (ge | ~ne) & 0x80808080u;
}

// ---- per-pixel arithmetic ------------------------------------------------------------------------
// MUFU approximations without the range-fixing prologue/epilogue the libdevice wrappers add: the
// arguments are range limited by construction (ex2: x - max <= 0, underflow to 0 is the right
// answer; lg2 / rcp: Σ exp in [1, C]).  Each is one SASS instruction.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed bf16x2 helpers (bf16 logits, NCHW: a 32-bit word of a class plane is a pixel pair) ----------
// HMNMX2.BF16 / HSET2.BF16 / HFMA2.BF16 work on both halves at once; sub.f32.bf16 (sm_100: FHADD.BF16 with an
// .H0/.H1 operand selector) subtracts an fp32 from a bf16 half EXACTLY into fp32, so the packed words never have to
// be unpacked.  max ignores NaN operands and returns +0 for {-0, +0}; set.eq is an IEEE compare (-0 == +0, NaN != NaN):
// the same tie rules as the fp32 FMNMX / FSETP path.
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_eq_mask(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("set.eq.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   // 0xffff per half where equal
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_nan_mask(uint32_t a) {
    uint32_t d;
    asm("set.nan.u32.bf16x2 %0, %1, %1;" : "=r"(d) : "r"(a));         // 0xffff per half that is NaN
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// (mask & a) | (~mask & b), one LOP3
__device__ __forceinline__ uint32_t lop3_select(uint32_t a, uint32_t b, uint32_t mask) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(mask), "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned short bf16_half(uint32_t w, int hi) {
    return static_cast<unsigned short>(hi ? (w >> 16) : (w & 0xffffu));
}
// bf16 - fp32 -> fp32, exact (mixed-precision add of PTX ISA 8.6, sm_100+)
__device__ __forceinline__ float sub_f32_bf16(unsigned short a, float b) {
    float d;
    asm("sub.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(a), "f"(b));
    return d;
}

// ---- one-shot exchange of the per-rank total weight over peer-mapped memory (NVLink) ---------------------------------
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// flag store of the fence + relaxed-store release pattern: ONE system-scope fence after the payload stores orders them
// before every flag that follows (a st.release per peer would repeat the fence — microseconds each — once per rank)
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// this rank's value of exchange number `seq` -> slot [seq % depth][rank] of every rank's block, then the flags
__device__ __forceinline__ void xchg_publish(const CeParams& p, double mine, unsigned long long seq) {
    const int slot = static_cast<int>(seq % kXDepth);
    for (int q = 0; q < p.xworld; ++q) *reinterpret_cast<volatile double*>(&p.xpeer[q]->slots[slot][p.xrank]) = mine;
    __threadfence_system();
    for (int q = 0; q < p.xworld; ++q) st_relaxed_sys_u32(&p.xpeer[q]->flags[slot][p.xrank], static_cast<unsigned int>(seq));
}
// One WARP per CTA (all 32 lanes call it, `mine` uniform).  The writer CTA publishes this rank's value to every rank's
// block — lane q serves peer q, and only if the previous launch has not already done so (published >= seq) — then lane q
// of every caller waits for rank q's value of this exchange in its OWN block, and the values are added in rank order
// (shuffles), so that every rank and every CTA gets the bit-identical total after two memory round trips, whatever the
// number of ranks.  A peer that never arrives (or has overrun the ring) ends the wait after ~4 s with NaN and a count in
// errors — never a hang.
__device__ __forceinline__ double xchg_total_weight(const CeParams& p, double mine, bool writer) {
    const int lane = threadIdx.x & 31;
    XchgBlock* local = p.xpeer[p.xrank];
    unsigned long long seq = 0ull, pub = 0ull;
    if (lane == 0) {
        seq = *reinterpret_cast<volatile unsigned long long*>(&local->seq) + 1ull;
        pub = *reinterpret_cast<volatile unsigned long long*>(&local->published);
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    pub = __shfl_sync(0xffffffffu, pub, 0);
    const int slot = static_cast<int>(seq % kXDepth);
    const unsigned int tag = static_cast<unsigned int>(seq);
    if (writer && pub < seq && lane < p.xworld) {
        *reinterpret_cast<volatile double*>(&p.xpeer[lane]->slots[slot][p.xrank]) = mine;
        __threadfence_system();
        st_relaxed_sys_u32(&p.xpeer[lane]->flags[slot][p.xrank], tag);
    }
    double v = 0.0;
    bool failed = false;
    if (lane < p.xworld) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys_u32(&local->flags[slot][lane]) != tag) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ull) {
                failed = true;
                break;
            }
            __nanosleep(64);
        }
        if (!failed) v = *reinterpret_cast<volatile double*>(&local->slots[slot][lane]);
    }
    __syncwarp();
    const bool any_failed = __any_sync(0xffffffffu, failed);
    double tot = 0.0;
    for (int q = 0; q < p.xworld; ++q) tot += __shfl_sync(0xffffffffu, v, q);     // rank order
    if (any_failed) {
        if (lane == 0) atomicAdd(&local->errors, 1ull);
        tot = __longlong_as_double(0x7ff8000000000000ll);
    }
    return tot;
}

// torch.max over one pixel's classes with its NaN rule, from a fetch functor (slow path only).
template <int C, typename F>
__device__ __forceinline__ int argmax_nan_aware(F&& fetch) {
    float best = fetch(0);
    int arg = 0;
#pragma unroll
    for (int c = 1; c < C; ++c) {
        const float v = fetch(c);
        if (better(v, best)) {
            best = v;
            arg = c;
        }
    }
    return arg;
}

// exp(x - m) for x <= m.  FUSED (bf16 logits, 1e-2 tolerance): one FFMA + MUFU, x*log2e - m*log2e,
// whose rounding error in the exponent is <= ulp(m*log2e)/2 (about 1e-6 relative for |m| ~ 20);
// otherwise the subtraction is done first and is exact for nearby values: FADD + FMUL + MUFU.
template <bool FUSED>
__device__ __forceinline__ float exp_shifted(float x, float m) {
    if constexpr (FUSED) return ex2_ftz(fmaf(x, kLog2e, -m * kLog2e));
    else return ex2_ftz((x - m) * kLog2e);
}

// Softmax statistics of one pixel.  On exit x[c] = exp(x[c] - max), m = max (NaNs skipped),
// s = Σ x[c], and arg = first index with x[c] == max, which is torch's argmax whenever the row
// holds no NaN / +inf / all -inf; those rows give s = NaN and the caller redoes the argmax with
// argmax_nan_aware on the original values.
template <int C, bool FUSED = false>
__device__ __forceinline__ void softmax_core(float (&x)[C], float& m, float& s, int& arg) {
    m = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
    arg = C - 1;
#pragma unroll
    for (int c = C - 2; c >= 0; --c) arg = (x[c] == m) ? c : arg;
    x[0] = exp_shifted<FUSED>(x[0], m);
    s = x[0];                     // same bits as 0 + e0 (e0 >= +0), one FADD fewer per pixel
#pragma unroll
    for (int c = 1; c < C; ++c) {
        x[c] = exp_shifted<FUSED>(x[c], m);
        s += x[c];
    }
}

// One pixel, register-only variant (direct kernels): x[0..C) holds the logits on entry and (if
// do_grad) the gradients on exit.  Returns the argmax; accumulates the loss terms.  tv uses the
// label encoding above.  `refetch(c)` re-reads logit c (only used for rows with NaN / inf).
template <int C, typename F>
__device__ __forceinline__ int pixel_ce(float (&x)[C], int tv, const float* __restrict__ wsm, float inv_tw,
                                        bool do_grad, float& step_l, float& step_w, unsigned int& bad, F&& refetch) {
    const bool valid = static_cast<unsigned int>(tv) < static_cast<unsigned int>(C);
    bad += (!valid && tv != -1) ? 1u : 0u;
    float xt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) xt = (c == tv) ? x[c] : xt;
    float m, s;
    int arg;
    softmax_core<C>(x, m, s, arg);
    if (s != s) arg = argmax_nan_aware<C>(refetch);
    const float w = valid ? wsm[tv] : 0.f;
    const float nll = fmaf(lg2_ftz(s), kLn2, m - xt);
    step_l += valid ? w * nll : 0.f;
    step_w += w;
    if (do_grad) {
        const float gsc = valid ? w * inv_tw : 0.f;  // exact zeros at ignored pixels even when 1/Σw = inf
        const float r = gsc * rcp_ftz(s);
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = fmaf(x[c], r, (c == tv) ? -gsc : 0.f);
    }
    return arg;
}

// Block-reduce {Σ w·nll, Σ w}; the last CTA folds the per-block fp64 partials in a fixed
// order -> bit-stable run to run, no float atomics.  Must be called by threads 0..NWARPS*32-1: the whole CTA
// (BAR = 0, __syncthreads) or the first NWARPS warps of a wider one, synchronised on named barrier BAR — the TMA-staged
// kernel runs it on its consumer warps while the store warp is still draining the last gradients.
template <int NWARPS, int BAR = 0>
__device__ __forceinline__ void finish_loss(const CeParams& p, double lsum, double wsum, unsigned int bad) {
    __shared__ double red[2 * NWARPS];
    __shared__ unsigned int is_last;
    const auto sync = [] {
        if constexpr (BAR == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NWARPS * 32) : "memory");
    };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    lsum = warp_sum(lsum);
    wsum = warp_sum(wsum);
    bad = warp_sum(bad);
    if (lane == 0) {
        red[warp] = lsum;
        red[NWARPS + warp] = wsum;
        if (bad) atomicAdd(&p.ws->bad, static_cast<unsigned long long>(bad));
    }
    sync();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            a += red[w];
            b += red[NWARPS + w];
        }
        p.ws->partial[2 * blockIdx.x] = a;
        p.ws->partial[2 * blockIdx.x + 1] = b;
        // one acq_rel ticket instead of fence + atomic + fence: its release publishes this CTA's partials (and, through
        // the barrier above, its warps' out-of-bounds counts), its acquire — in the CTA that draws the last ticket —
        // makes every earlier CTA's visible to the fold below (the barrier after it extends that to the other threads)
        unsigned int t;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(&p.ws->ticket) : "memory");
        is_last = (t == gridDim.x - 1);
    }
    sync();
    if (!is_last) return;
    const unsigned long long nbad = __ldcg(&p.ws->bad);   // every CTA's count landed before its ticket
    double a = 0.0, b = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += NWARPS * 32) {
        a += __ldcg(&p.ws->partial[2 * i]);
        b += __ldcg(&p.ws->partial[2 * i + 1]);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    double next_sum = 0.0;
    if (p.next_tw_out && warp == 0) {
        // the next batch's total weight: per-CTA partials folded in a fixed order by one warp
        for (unsigned int i = lane; i < gridDim.x; i += 32) next_sum += __ldcg(&p.ws->pre[1][i]);
        next_sum = warp_sum(next_sum);
        if (lane == 0) {
            p.next_tw_out[0] = next_sum;
            p.next_tw_out[1] = 1.0 / next_sum;
        }
    }
    sync();
    if (lane == 0) {
        red[warp] = a;
        red[NWARPS + warp] = b;
    }
    sync();
    if (threadIdx.x == 0) {
        a = 0.0;
        b = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            a += red[w];
            b += red[NWARPS + w];
        }
        if (p.status && nbad) atomicAdd(p.status, nbad);
        if (p.loss_sums) {
            p.loss_sums[0] = a;
            p.loss_sums[1] = b;
            p.loss_sums[2] = static_cast<double>(nbad);
        }
        if (p.loss_out) {
            float l = static_cast<float>(a / b);  // 0/0 -> NaN like torch (everything ignored)
            if (nbad) l = __int_as_float(0x7fc00000);
            *p.loss_out = l;
        }
        p.ws->bad = 0ull;
        p.ws->ticket = 0u;
        p.ws->next_chunk = 0u;
        p.ws->gbar = 0u;
        if (p.tw_mode != 0 && p.xworld > 1) {
            XchgBlock* local = p.xpeer[p.xrank];
            const unsigned long long seq = local->seq + 1ull;        // this launch's exchange; every CTA read it long ago
            if (p.next_tw_out) {
                // the NEXT exchange's value is known already: send it to the peers now, a whole step before they need it
                xchg_publish(p, next_sum, seq + 1ull);
                local->published = seq + 1ull;
            }
            local->seq = seq;
        }
        __threadfence();
    }
}

#endif  // __CUDACC__

// persistent grid = resident CTAs per SM x SMs.  The attribute call and the occupancy query cost a few
// microseconds of host time each, so the answer is remembered per (kernel, threads, smem, device).
template <typename K>
int persistent_grid(K kernel, int threads, int smem_bytes, int* grid_out) {
    struct Entry {
        const void* fn;
        int threads, smem, dev, grid;
    };
    static thread_local Entry cache[16];
    static thread_local int n_cached = 0, next_slot = 0;
    int dev = 0;
    CVCS_CUDA_OK(cudaGetDevice(&dev));
    const void* fn = reinterpret_cast<const void*>(kernel);
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].fn == fn && cache[i].threads == threads && cache[i].smem == smem_bytes && cache[i].dev == dev) {
            *grid_out = cache[i].grid;
            return CVCS_OK;
        }
    int per_sm = 0;
    // opt in to the architectural maximum once (227 KB on sm_100a) rather than to this call's size: the
    // attribute is per kernel, and a later, smaller request must not lower it under a cached larger one
    if (smem_bytes > 48 * 1024) {
        cudaFuncAttributes fa;
        CVCS_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
        const int max_dyn = 227 * 1024 - static_cast<int>(fa.sharedSizeBytes);   // static + dynamic <= 227 KB
        if (smem_bytes > max_dyn)
            return set_error(CVCS_ERR_UNSUPPORTED, "kernel needs %d B of dynamic shared memory, %d available", smem_bytes, max_dyn);
        CVCS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    }
    CVCS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes));
    if (per_sm < 1) return set_error(CVCS_ERR_UNSUPPORTED, "kernel does not fit on an SM (smem %d B)", smem_bytes);
    int g = per_sm * num_sms();
    if (g > kMaxGrid) g = kMaxGrid;
    *grid_out = g;
    cache[next_slot] = Entry{fn, threads, smem_bytes, dev, g};
    next_slot = (next_slot + 1) % 16;
    if (n_cached < 16) ++n_cached;
    return CVCS_OK;
}

}  // namespace cvcs
