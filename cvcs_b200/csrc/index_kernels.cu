// index_kernels.cu — the label / index-map kernels around K1:
//   K4 label histogram (+ Σ w[y])   <- Loader._get_class_count, dataset.py:346-358; the 'mean'
//                                       divisor of nn.CrossEntropyLoss (utils.py:230,238)
//   K2 argmax over the class dim     <- torch.max / torch.argmax, utils.py:90,158,504; esa.py:56
//   K3 confusion matrix from indices <- MulticlassConfusionMatrix.update, utils.py:93-94
//   grad scaling                     <- autograd's grad_output * dlogits when grad_output != 1
// All are byte/integer streaming kernels: HBM-bound, 128-bit coalesced loads, persistent grids,
// shared-memory privatised bins flushed with one 64-bit global atomic per bin per CTA.
#include "common.cuh"

namespace cvcs {
namespace {

template <typename K>
int grid_for(K kernel, int smem_bytes, long long blocks_needed, int* grid_out) {
    int per_sm = 0;
    if (smem_bytes > 48 * 1024)
        CVCS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CVCS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem_bytes));
    if (per_sm < 1) return set_error(CVCS_ERR_UNSUPPORTED, "kernel does not fit on an SM (smem %d B)", smem_bytes);
    long long g = static_cast<long long>(per_sm) * num_sms();
    if (g > kMaxGrid) g = kMaxGrid;
    if (g > blocks_needed) g = blocks_needed;
    if (g < 1) g = 1;
    *grid_out = static_cast<int>(g);
    return CVCS_OK;
}

// Classify a label: [0,C) -> itself, ignore_index outside [0,C) -> C, anything else -> C+1.
__device__ __forceinline__ int hist_bin_u8(int t, int C, int ign_outside) {
    return t < C ? t : (t == ign_outside ? C : C + 1);
}
__device__ __forceinline__ int hist_bin_i64(long long t, int C, long long ignore_index) {
    return (t >= 0 && t < C) ? static_cast<int>(t) : (t == ignore_index ? C : C + 1);
}

// ---- K4 ------------------------------------------------------------------------------------
struct HistParams {
    const void* target;
    long long n;            // pixels
    unsigned long long* hist;  // nullable, accumulated into
    const float* weight;    // nullable
    double* tw_out;         // nullable {Σw, 1/Σw}
    Workspace* ws;
    long long ignore_index;
    int C;
    int target_i64;
    unsigned char* narrow;  // nullable (i64 labels only): u8 copy of the labels, 255 = ignored, 254 = out of range
};

template <bool PRIV>
__global__ void __launch_bounds__(kThreads) label_hist_kernel(const HistParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int is_last;
    const int C = p.C;
    const int nb = C + 2;
    BinAcc<PRIV> acc;
    acc.init(smem, nb);
    unsigned int since_flush = 0;

    if (!p.target_i64) {
        const uint8_t* __restrict__ tgt = reinterpret_cast<const uint8_t*>(p.target);
        const int ign_outside = (p.ignore_index >= C && p.ignore_index <= 255) ? static_cast<int>(p.ignore_index) : -1;
        const long long n16 = p.n / 16;  // pointer is 16-byte aligned (checked on the host)
        // label maps have long runs: a 32-bit word whose four bytes agree is one +4 update
        auto add_word = [&](uint32_t w) {
            const int t0 = static_cast<int>(w & 0xff);
            if (w == static_cast<uint32_t>(t0) * 0x01010101u) {
                acc.add(hist_bin_u8(t0, C, ign_outside), 4);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc.add(hist_bin_u8(static_cast<int>((w >> (8 * k)) & 0xff), C, ign_outside), 1);
            }
        };
        constexpr int U = 2;  // independent 128-bit loads in flight per thread
        const long long stride = static_cast<long long>(gridDim.x) * kThreads;
        for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < n16; base += U * stride) {
            Raw<16> r[U];
            bool have[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = base + u * stride + threadIdx.x;
                have[u] = i < n16;
                if (have[u]) r[u].load(tgt + 16 * i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!have[u]) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) add_word(r[u].word(k));
            }
            if (PRIV) {
                since_flush += 16 * U;
                if (since_flush > 65535u - 16u * U) {
                    acc.flush(p.ws->hist);
                    since_flush = 0;
                }
            }
        }
        // tail (< 16 pixels)
        if (blockIdx.x == 0) {
            const long long i = n16 * 16 + threadIdx.x;
            if (i < p.n) acc.add(hist_bin_u8(tgt[i], C, ign_outside));
        }
    } else {
        const long long* __restrict__ tgt = reinterpret_cast<const long long*>(p.target);
        const long long n2 = p.n / 2;
        for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < n2;
             base += static_cast<long long>(gridDim.x) * kThreads) {
            const long long i = base + threadIdx.x;
            if (i < n2) {
                Raw<16> r;
                r.load(tgt + 2 * i);
                const long long a = (static_cast<long long>(r.v.y) << 32) | r.v.x;
                const long long b = (static_cast<long long>(r.v.w) << 32) | r.v.z;
                acc.add(hist_bin_i64(a, C, p.ignore_index));
                acc.add(hist_bin_i64(b, C, p.ignore_index));
            }
            if (PRIV) {
                since_flush += 2;
                if (since_flush > 65535u - 2u) {
                    acc.flush(p.ws->hist);
                    since_flush = 0;
                }
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (p.n & 1)) acc.add(hist_bin_i64(tgt[p.n - 1], C, p.ignore_index));
    }
    acc.flush(p.ws->hist);

    // last CTA: publish this call's histogram, Σw, and zero the scratch bins
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&p.ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double tw_red[kWarps];
    double tw = 0.0;
    for (int b = threadIdx.x; b < nb; b += kThreads) {
        const unsigned long long cnt = __ldcg(&p.ws->hist[b]);
        p.ws->hist[b] = 0ull;
        if (p.hist && cnt) atomicAdd(p.hist + b, cnt);
        if (b < C && static_cast<long long>(b) != p.ignore_index)
            tw += static_cast<double>(cnt) * static_cast<double>(p.weight ? p.weight[b] : 1.0f);
    }
    tw = warp_sum(tw);
    if ((threadIdx.x & 31) == 0) tw_red[threadIdx.x >> 5] = tw;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += tw_red[w];
        if (p.tw_out) {
            p.tw_out[0] = s;
            p.tw_out[1] = 1.0 / s;
        }
        p.ws->ticket = 0u;
        __threadfence();
    }
}

// ---- K4, lean mode: Σ_i v_i w[y_i] without the histogram ------------------------------------------
// When only the 'mean' divisor is wanted (the pre-pass in front of K1), counting is unnecessary: the
// total weight is a plain sum of per-pixel table look-ups, with no read-modify-write chains.  u8 labels
// index a 256-entry shared-memory table directly (0 for ignore_index and for anything >= C).
// Per-thread fp32 partials of a few dozen terms -> fp64 block partials -> fixed-order fold by the last
// CTA (run-to-run bit-stable).
__global__ void __launch_bounds__(kThreads) weight_sum_kernel(const HistParams p) {
    __shared__ float lut[256];
    __shared__ double red[kWarps];
    __shared__ unsigned int is_last;
    const int C = p.C;
    {
        const int t = threadIdx.x;
        lut[t] = (t < C && static_cast<long long>(t) != p.ignore_index) ? (p.weight ? p.weight[t] : 1.0f) : 0.f;
    }
    __syncthreads();
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    if (!p.target_i64) {
        const uint8_t* __restrict__ tgt = reinterpret_cast<const uint8_t*>(p.target);
        const long long n16 = p.n / 16;
        constexpr int U = 2;
        for (long long base = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; base < n16; base += U * stride) {
            Raw<16> r[U];
            bool have[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                have[u] = base + u * stride < n16;
                if (have[u]) r[u].load(tgt + 16 * (base + u * stride));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!have[u]) continue;
#pragma unroll
                for (int k = 0; k < 16; ++k) a[k & 3] += lut[(r[u].word(k / 4) >> (8 * (k % 4))) & 0xff];
            }
        }
        if (blockIdx.x == 0) {
            const long long i = n16 * 16 + threadIdx.x;
            if (i < p.n) a[0] += lut[tgt[i]];
        }
    } else {
        const long long* __restrict__ tgt = reinterpret_cast<const long long*>(p.target);
        const long long n2 = p.n / 2;
        auto w_of = [&](long long v) { return (v >= 0 && v < C) ? lut[static_cast<int>(v)] : 0.f; };  // lut[ignore] == 0
        // u8 code for K1: ignore_index -> 255, other values outside [0, C) -> 254 (counted as out of bounds there)
        auto code_of = [&](long long v) -> unsigned int {
            return v == p.ignore_index ? 255u : ((v >= 0 && v < C) ? static_cast<unsigned int>(v) : 254u);
        };
        constexpr int U = 4;
        // the byte labels are read by the K1 launch that follows: ask L2 to keep them (16.8 MB for a 16-tile batch)
        unsigned long long keep;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
        for (long long base = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; base < n2; base += U * stride) {
            Raw<16> r[U];
            bool have[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                have[u] = base + u * stride < n2;
                if (have[u]) r[u].load(tgt + 2 * (base + u * stride));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!have[u]) continue;
                const long long v0 = (static_cast<long long>(r[u].v.y) << 32) | r[u].v.x;
                const long long v1 = (static_cast<long long>(r[u].v.w) << 32) | r[u].v.z;
                a[u & 3] += w_of(v0);
                a[(u + 2) & 3] += w_of(v1);
                if (p.narrow) {
                    const unsigned int c0 = code_of(v0), c1 = code_of(v1);
                    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(reinterpret_cast<unsigned short*>(p.narrow) + (base + u * stride)),
                                 "h"(static_cast<unsigned short>(c0 | (c1 << 8))), "l"(keep)
                                 : "memory");
                }
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (p.n & 1)) {
            a[0] += w_of(tgt[p.n - 1]);
            if (p.narrow) p.narrow[p.n - 1] = static_cast<unsigned char>(code_of(tgt[p.n - 1]));
        }
    }
    double s = static_cast<double>((a[0] + a[1]) + (a[2] + a[3]));
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) b += red[w];
        p.ws->partial[blockIdx.x] = b;
        __threadfence();
        const unsigned int t = atomicAdd(&p.ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double tot = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) tot += __ldcg(&p.ws->partial[i]);
    tot = warp_sum(tot);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) b += red[w];
        p.tw_out[0] = b;
        p.tw_out[1] = 1.0 / b;
        p.ws->ticket = 0u;
        __threadfence();
    }
}

__global__ void total_weight_kernel(const unsigned long long* __restrict__ hist, const float* __restrict__ weight,
                                    int C, long long ignore_index, double* __restrict__ out) {
    // one warp; fixed order -> deterministic
    double s = 0.0;
    for (int c = threadIdx.x; c < C; c += 32)
        if (static_cast<long long>(c) != ignore_index)
            s += static_cast<double>(hist[c]) * static_cast<double>(weight ? weight[c] : 1.0f);
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        out[0] = s;
        out[1] = 1.0 / s;
    }
}

// ---- grad scaling ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) scale_kernel(T* __restrict__ x, long long n, const float* __restrict__ scale) {
    constexpr int VEC = 16 / sizeof(T);
    const float s = *scale;
    if (s == 1.0f) return;   // loss.backward() feeds exactly 1 (train.py:125): nothing to do, and no host had to look
    const long long nv = n / VEC;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < nv;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        float v[VEC];
        VecIO<T, VEC>::load(x + i * VEC, v);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] *= s;
        VecIO<T, VEC>::store(x + i * VEC, v);
    }
    if (blockIdx.x == 0) {
        const long long i = nv * VEC + threadIdx.x;
        if (i < n) {
            float v[1];
            VecIO<T, 1>::load(x + i, v);
            v[0] *= s;
            VecIO<T, 1>::store(x + i, v);
        }
    }
}

// ---- K2 ------------------------------------------------------------------------------------
struct ArgmaxParams {
    const void* logits;
    void* out;
    long long hw;
    long long n_items;
    unsigned int items_per_image;
    int C;
    int out_i64;
};

template <int VEC>
__device__ __forceinline__ void store_index(void* outp, int out_i64, long long pix, const int (&a)[VEC]) {
    if (out_i64) {
        long long* out = reinterpret_cast<long long*>(outp) + pix;
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
        uint8_t* out = reinterpret_cast<uint8_t*>(outp) + pix;
        if constexpr (VEC == 1) {
            Raw<1> r;
            r.v = static_cast<uint8_t>(a[0]);
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

// NCHW: a thread owns VEC consecutive pixels and walks the class planes (4 planes in flight).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) argmax_nchw_kernel(const ArgmaxParams p) {
    const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
    const int C = p.C;
    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < p.n_items;
         base += static_cast<long long>(gridDim.x) * kThreads) {
        const long long item = base + threadIdx.x;
        if (item >= p.n_items) continue;
        const unsigned int item32 = static_cast<unsigned int>(item);
        const unsigned int b = item32 / p.items_per_image;
        const unsigned int g = item32 - b * p.items_per_image;
        const long long pix = static_cast<long long>(b) * p.hw + static_cast<long long>(g) * VEC;
        const T* src = logits + static_cast<long long>(b) * C * p.hw + static_cast<long long>(g) * VEC;
        float best[VEC];
        int arg[VEC];
        VecIO<T, VEC>::load(src, best);
#pragma unroll
        for (int v = 0; v < VEC; ++v) arg[v] = 0;
        int c = 1;
        for (; c + 4 <= C; c += 4) {
            float x[4][VEC];
#pragma unroll
            for (int j = 0; j < 4; ++j) VecIO<T, VEC>::load(src + (c + j) * p.hw, x[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (better(x[j][v], best[v])) {
                        best[v] = x[j][v];
                        arg[v] = c + j;
                    }
        }
        for (; c < C; ++c) {
            float x[VEC];
            VecIO<T, VEC>::load(src + c * p.hw, x);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (better(x[v], best[v])) {
                    best[v] = x[v];
                    arg[v] = c;
                }
        }
        store_index<VEC>(p.out, p.out_i64, pix, arg);
    }
}

// Any layout via strides, one pixel per thread.
template <typename T>
__global__ void __launch_bounds__(kThreads) argmax_generic_kernel(const ArgmaxParams p, long long class_stride,
                                                                  long long pixel_stride, long long image_stride) {
    const T* __restrict__ logits = reinterpret_cast<const T*>(p.logits);
    const int C = p.C;
    for (long long pix = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; pix < p.n_items;
         pix += static_cast<long long>(gridDim.x) * kThreads) {
        const long long b = pix / p.hw;
        const T* src = logits + b * image_stride + (pix - b * p.hw) * pixel_stride;
        float best[1];
        VecIO<T, 1>::load(src, best);
        int arg[1] = {0};
        for (int c = 1; c < C; ++c) {
            float x[1];
            VecIO<T, 1>::load(src + c * class_stride, x);
            if (better(x[0], best[0])) {
                best[0] = x[0];
                arg[0] = c;
            }
        }
        store_index<1>(p.out, p.out_i64, pix, arg);
    }
}

// ---- K3 ------------------------------------------------------------------------------------
struct ConfParams {
    const void* pred;
    const void* target;
    long long n;
    unsigned long long* confmat;
    unsigned long long* status;
    long long ignore_index;
    int C;
    int pred_i64;
    int target_i64;
    int reps;  // shared-memory replicas of the bins (shared mode); 0 = global atomics
};

template <int VEC>
__device__ __forceinline__ void load_index(const void* base, int is_i64, long long i, long long (&v)[VEC]) {
    if (is_i64) {
        const long long* q = reinterpret_cast<const long long*>(base) + i;
        if constexpr (VEC % 2 == 0) {
#pragma unroll
            for (int k = 0; k < VEC / 2; ++k) {
                Raw<16> r;
                r.load(q + 2 * k);
                v[2 * k] = (static_cast<long long>(r.v.y) << 32) | r.v.x;
                v[2 * k + 1] = (static_cast<long long>(r.v.w) << 32) | r.v.z;
            }
        } else {
            v[0] = q[0];
        }
    } else {
        const uint8_t* q = reinterpret_cast<const uint8_t*>(base) + i;
        if constexpr (VEC == 1) {
            v[0] = q[0];
        } else {
            Raw<VEC> r;
            r.load(q);
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[k] = (r.word(k / 4) >> (8 * (k % 4))) & 0xff;
        }
    }
}

template <bool PRIV, int VEC>
__global__ void __launch_bounds__(kThreads) confmat_kernel(const ConfParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int C = p.C;
    BinAcc<PRIV> acc;
    acc.init(smem, C * C, p.reps, p.confmat);
    unsigned int since_flush = 0, bad = 0;
    const long long nv = p.n / VEC;
    auto one = [&](long long t, long long q) {
        if (t == p.ignore_index) return;
        if (t < 0 || t >= C || q < 0 || q >= C) {
            ++bad;
            return;
        }
        acc.add(static_cast<int>(t) * C + static_cast<int>(q));
    };
    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < nv;
         base += static_cast<long long>(gridDim.x) * kThreads) {
        const long long i = base + threadIdx.x;
        if (i < nv) {
            long long t[VEC], q[VEC];
            load_index<VEC>(p.target, p.target_i64, i * VEC, t);
            load_index<VEC>(p.pred, p.pred_i64, i * VEC, q);
#pragma unroll
            for (int k = 0; k < VEC; ++k) one(t[k], q[k]);
        }
        if (PRIV) {
            since_flush += VEC;
            if (since_flush > 65535u - VEC) {
                acc.flush(p.confmat);
                since_flush = 0;
            }
        }
    }
    if (blockIdx.x == 0) {
        const long long i = nv * VEC + threadIdx.x;
        if (i < p.n) {
            long long t[1], q[1];
            load_index<1>(p.target, p.target_i64, i, t);
            load_index<1>(p.pred, p.pred_i64, i, q);
            one(t[0], q[0]);
        }
    }
    acc.flush(p.confmat);
    bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0 && bad && p.status) atomicAdd(p.status, static_cast<unsigned long long>(bad));
}

// Both maps u8 and 16-byte aligned (the stored-mask / K2-output case): 16 pixels per 128-bit load, two loads
// of each map in flight, and one +4 update for a word whose four (target, prediction) pairs agree — index
// maps have long runs.
template <bool PRIV>
__global__ void __launch_bounds__(kThreads) confmat_u8_kernel(const ConfParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int C = p.C;
    BinAcc<PRIV> acc;
    acc.init(smem, C * C, p.reps, p.confmat);
    unsigned int since_flush = 0, bad = 0;
    const int ign = (p.ignore_index >= 0 && p.ignore_index <= 255) ? static_cast<int>(p.ignore_index) : -1;
    auto one = [&](int t, int q, unsigned int n) {
        if (t == ign) return;
        if (t >= C || q >= C) {
            bad += n;
            return;
        }
        acc.add(t * C + q, n);
    };
    // C <= 16: the key t*C + q of a pixel fits a byte, so a word of four in-range, not ignored pairs — the common case,
    // found with three byte-parallel tests — becomes four keys with ONE multiply-add; a word whose keys agree (index
    // maps have long runs) is a single +4 update.  Anything else takes the pixel-by-pixel path.
    const bool swar = C <= 16;
    const uint32_t ge_add = static_cast<uint32_t>(128 - (swar ? C : 0)) * 0x01010101u;
    const uint32_t ign4 = static_cast<uint32_t>(ign >= 0 ? ign : 0) * 0x01010101u;
    const uint32_t ne_or = ign >= 0 ? 0u : 0x80808080u;
    const uint32_t Cu = static_cast<uint32_t>(C);
    auto word = [&](uint32_t tw, uint32_t qw) {
        if (swar) {
            const uint32_t tge = ((tw & 0x7f7f7f7fu) + ge_add) | tw;                     // bit 7 of a lane: target >= C
            const uint32_t qge = ((qw & 0x7f7f7f7fu) + ge_add) | qw;                     //                  prediction >= C
            const uint32_t z = tw ^ ign4;
            const uint32_t tne = (((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) | ne_or;        //                  target != ignore_index
            if ((((tge | qge) | ~tne) & 0x80808080u) == 0u) {
                const uint32_t k4 = tw * Cu + qw;
                const uint32_t k0 = k4 & 0xffu;
                if (k4 == k0 * 0x01010101u) {
                    acc.add(static_cast<int>(k0), 4u);
                } else {
                    acc.add(static_cast<int>(k0));
                    acc.add(static_cast<int>((k4 >> 8) & 0xffu));
                    acc.add(static_cast<int>((k4 >> 16) & 0xffu));
                    acc.add(static_cast<int>(k4 >> 24));
                }
                return;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) one(static_cast<int>((tw >> (8 * k)) & 0xff), static_cast<int>((qw >> (8 * k)) & 0xff), 1u);
    };
    const uint8_t* __restrict__ tgt = reinterpret_cast<const uint8_t*>(p.target);
    const uint8_t* __restrict__ prd = reinterpret_cast<const uint8_t*>(p.pred);
    const long long n16 = p.n / 16;
    constexpr int U = 2;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < n16; base += U * stride) {
        Raw<16> rt[U], rq[U];
        bool have[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride + threadIdx.x;
            have[u] = i < n16;
            if (have[u]) {
                rt[u].load(tgt + 16 * i);
                rq[u].load(prd + 16 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!have[u]) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) word(rt[u].word(k), rq[u].word(k));
        }
        if (PRIV) {
            since_flush += 16 * U;
            if (since_flush > 65535u - 16u * U) {
                acc.flush(p.confmat);
                since_flush = 0;
            }
        }
    }
    if (blockIdx.x == 0) {
        const long long i = n16 * 16 + threadIdx.x;
        if (i < p.n) one(tgt[i], prd[i], 1u);
    }
    acc.flush(p.confmat);
    bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0 && bad && p.status) atomicAdd(p.status, static_cast<unsigned long long>(bad));
}

bool aligned_to(const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; }

}  // namespace

int label_hist_launch(const void* target, int target_dtype, long long n, int C, long long ignore_index,
                      unsigned long long* hist, const float* weight, double* tw_out, void* workspace,
                      cudaStream_t stream) {
    CVCS_REQUIRE((target || n == 0) && workspace, "cvcs_label_hist: NULL target/workspace");
    CVCS_REQUIRE(target_dtype == CVCS_U8 || target_dtype == CVCS_I64, "cvcs_label_hist: target dtype tag %d (want u8/i64)", target_dtype);
    CVCS_REQUIRE(n >= 0 && C >= 1, "cvcs_label_hist: bad n=%lld C=%d", n, C);
    if (C + 2 > kMaxHistBins) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_label_hist: C=%d > %d", C, kMaxHistBins - 2);
    CVCS_REQUIRE(aligned_to(target, 16), "cvcs_label_hist: target must be 16-byte aligned");
    HistParams p{};
    p.target = target;
    p.n = n;
    p.hist = hist;
    p.weight = weight;
    p.tw_out = tw_out;
    p.ws = reinterpret_cast<Workspace*>(workspace);
    p.ignore_index = ignore_index;
    p.C = C;
    p.target_i64 = target_dtype == CVCS_I64;
    if (!hist && tw_out && C <= 256) {
        // lean mode: only the total weight is wanted -> no counting (see weight_sum_kernel)
        const long long per_thread = p.target_i64 ? 8 : 32;
        long long blocks = (n / per_thread + kThreads - 1) / kThreads;
        const long long cap = 8ll * num_sms();
        if (blocks > cap) blocks = cap;
        if (blocks > kMaxGrid) blocks = kMaxGrid;
        if (blocks < 1) blocks = 1;
        weight_sum_kernel<<<static_cast<int>(blocks), kThreads, 0, stream>>>(p);
        CVCS_CUDA_OK(cudaGetLastError());
        return CVCS_OK;
    }
    const int per = p.target_i64 ? 2 : 32;  // pixels per thread per step (u8: two 128-bit loads)
    long long blocks = (n / per + kThreads - 1) / kThreads;
    // every CTA ends with one global atomic per bin on the same few addresses: cap the grid at four CTAs per SM (two
    // left the kernel latency-bound: 24 % of the issue slots, 13 % of the DRAM rate in ncu, profiles/r2c) and let each
    // thread stream more labels instead
    const long long cap = 4ll * num_sms();
    if (blocks > cap) blocks = cap;
    int grid = 0;
    if (C + 2 <= 64) {
        const int smem = BinAcc<true>::smem_bytes(C + 2);
        int rc = grid_for(label_hist_kernel<true>, smem, blocks, &grid);
        if (rc) return rc;
        label_hist_kernel<true><<<grid, kThreads, smem, stream>>>(p);
    } else {
        const int smem = BinAcc<false>::smem_bytes(C + 2);
        int rc = grid_for(label_hist_kernel<false>, smem, blocks, &grid);
        if (rc) return rc;
        label_hist_kernel<false><<<grid, kThreads, smem, stream>>>(p);
    }
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int labels_prepare_launch(const long long* target, long long n, int C, long long ignore_index, const float* weight,
                          double* tw_out, unsigned char* labels_u8, void* workspace, cudaStream_t stream) {
    CVCS_REQUIRE((target || n == 0) && tw_out && (labels_u8 || n == 0) && workspace, "cvcs_labels_prepare: NULL argument");
    CVCS_REQUIRE(n >= 0 && C >= 1 && C <= 254, "cvcs_labels_prepare: needs 1 <= C <= 254 (got %d)", C);
    CVCS_REQUIRE(aligned_to(target, 16) && aligned_to(labels_u8, 2), "cvcs_labels_prepare: misaligned pointers");
    HistParams p{};
    p.target = target;
    p.n = n;
    p.weight = weight;
    p.tw_out = tw_out;
    p.ws = reinterpret_cast<Workspace*>(workspace);
    p.ignore_index = ignore_index;
    p.C = C;
    p.target_i64 = 1;
    p.narrow = labels_u8;
    long long blocks = (n / 8 + kThreads - 1) / kThreads;
    const long long cap = 8ll * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks > kMaxGrid) blocks = kMaxGrid;
    if (blocks < 1) blocks = 1;
    weight_sum_kernel<<<static_cast<int>(blocks), kThreads, 0, stream>>>(p);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

// ---- one-shot all-reduce of a short f64 vector over peer-mapped memory (pass-end sums) ------------------------------
struct WidePeers {
    XchgRegion* r[kXMaxRanks];
};
static __global__ void __launch_bounds__(kThreads) xchg_allreduce_kernel(const WidePeers peers, int world, int rank, double* __restrict__ buf, int n) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_fail;
    XchgRegion* local = peers.r[rank];
    if (threadIdx.x == 0) {
        s_seq = *reinterpret_cast<volatile unsigned long long*>(&local->wide.seq) + 1ull;
        s_fail = 0;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = static_cast<int>(seq & 1ull);
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const double v = buf[i];
        for (int q = 0; q < world; ++q) *reinterpret_cast<volatile double*>(&peers.r[q]->wide.data[par][rank][i]) = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int q = 0; q < world; ++q)
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&peers.r[q]->wide.flags[par][rank]), "r"(static_cast<unsigned int>(seq)) : "memory");   // released by the fence above
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int q = 0; q < world && !s_fail; ++q) {
            for (;;) {
                unsigned int f;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&local->wide.flags[par][q]) : "memory");
                if (f == static_cast<unsigned int>(seq)) break;
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) {       // a peer never arrived: NaN results and an error count, never a hang
                    s_fail = 1;
                    atomicAdd(&local->block.errors, 1ull);
                    break;
                }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) {
        double s = 0.0;
        for (int q = 0; q < world; ++q) s += *reinterpret_cast<volatile double*>(&local->wide.data[par][q][i]);   // rank order
        buf[i] = s_fail ? __longlong_as_double(0x7ff8000000000000ll) : s;
    }
    if (threadIdx.x == 0) local->wide.seq = seq;
}
int xchg_allreduce_launch(XchgRegion* const* peers, int world, int rank, double* buf, int n, cudaStream_t stream) {
    WidePeers wp{};
    for (int q = 0; q < world; ++q) wp.r[q] = peers[q];
    xchg_allreduce_kernel<<<1, kThreads, 0, stream>>>(wp, world, rank, buf, n);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

static __global__ void reciprocal_kernel(const double* __restrict__ in, double* __restrict__ out) {
    const double v = in[0];
    out[0] = v;
    out[1] = 1.0 / v;
}
int reciprocal_launch(const double* in, double* out2, cudaStream_t stream) {
    reciprocal_kernel<<<1, 1, 0, stream>>>(in, out2);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int total_weight_launch(const unsigned long long* hist, const float* weight, int C, long long ignore_index,
                        double* out, cudaStream_t stream) {
    CVCS_REQUIRE(hist && out && C >= 1, "cvcs_total_weight: NULL hist/out or C < 1");
    total_weight_kernel<<<1, 32, 0, stream>>>(hist, weight, C, ignore_index, out);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int scale_launch(void* x, int dtype, long long n, const float* scale, cudaStream_t stream) {
    CVCS_REQUIRE(x && scale && n >= 0, "cvcs_scale_inplace: NULL x/scale");
    CVCS_REQUIRE(dtype == CVCS_F32 || dtype == CVCS_BF16, "cvcs_scale_inplace: dtype tag %d", dtype);
    CVCS_REQUIRE(aligned_to(x, 16), "cvcs_scale_inplace: x must be 16-byte aligned");
    const long long per = dtype == CVCS_F32 ? 4 : 8;
    long long blocks = (n / per + kThreads - 1) / kThreads;
    long long g = static_cast<long long>(num_sms()) * 8;
    if (g > blocks) g = blocks;
    if (g < 1) g = 1;
    if (dtype == CVCS_F32) scale_kernel<float><<<static_cast<int>(g), kThreads, 0, stream>>>(reinterpret_cast<float*>(x), n, scale);
    else scale_kernel<__nv_bfloat16><<<static_cast<int>(g), kThreads, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(x), n, scale);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int argmax_launch(const void* logits, int logits_dtype, int layout, int B, int C, int H, int W, void* out,
                  int out_dtype, cudaStream_t stream) {
    CVCS_REQUIRE(logits && out, "cvcs_argmax: NULL logits/out");
    CVCS_REQUIRE(logits_dtype == CVCS_F32 || logits_dtype == CVCS_BF16, "cvcs_argmax: logits dtype tag %d", logits_dtype);
    CVCS_REQUIRE(layout == CVCS_NCHW || layout == CVCS_NHWC, "cvcs_argmax: layout %d", layout);
    CVCS_REQUIRE(out_dtype == CVCS_U8 || out_dtype == CVCS_I64, "cvcs_argmax: out dtype tag %d", out_dtype);
    CVCS_REQUIRE(B > 0 && C >= 1 && H > 0 && W > 0, "cvcs_argmax: bad shape");
    CVCS_REQUIRE(!(out_dtype == CVCS_U8 && C > 256), "cvcs_argmax: u8 output needs C <= 256");
    const long long hw = static_cast<long long>(H) * W, n = hw * B;
    CVCS_REQUIRE(n < (1ll << 33), "cvcs_argmax: too many pixels");
    ArgmaxParams p{};
    p.logits = logits;
    p.out = out;
    p.hw = hw;
    p.C = C;
    p.out_i64 = out_dtype == CVCS_I64;
    const int esize = logits_dtype == CVCS_F32 ? 4 : 2;
    const int vec = logits_dtype == CVCS_F32 ? 4 : 8;
    const bool fast = layout == CVCS_NCHW && hw % vec == 0 && aligned_to(logits, static_cast<size_t>(vec) * esize) &&
                      aligned_to(out, p.out_i64 ? 16 : vec);
    int grid = 0;
    if (fast) {
        p.n_items = n / vec;
        p.items_per_image = static_cast<unsigned int>(hw / vec);
        const long long blocks = (p.n_items + kThreads - 1) / kThreads;
        if (logits_dtype == CVCS_F32) {
            int rc = grid_for(argmax_nchw_kernel<float, 4>, 0, blocks, &grid);
            if (rc) return rc;
            argmax_nchw_kernel<float, 4><<<grid, kThreads, 0, stream>>>(p);
        } else {
            int rc = grid_for(argmax_nchw_kernel<__nv_bfloat16, 8>, 0, blocks, &grid);
            if (rc) return rc;
            argmax_nchw_kernel<__nv_bfloat16, 8><<<grid, kThreads, 0, stream>>>(p);
        }
    } else {
        p.n_items = n;
        const long long cs = layout == CVCS_NCHW ? hw : 1, ps = layout == CVCS_NCHW ? 1 : C;
        const long long is = static_cast<long long>(C) * hw;
        const long long blocks = (n + kThreads - 1) / kThreads;
        if (logits_dtype == CVCS_F32) {
            int rc = grid_for(argmax_generic_kernel<float>, 0, blocks, &grid);
            if (rc) return rc;
            argmax_generic_kernel<float><<<grid, kThreads, 0, stream>>>(p, cs, ps, is);
        } else {
            int rc = grid_for(argmax_generic_kernel<__nv_bfloat16>, 0, blocks, &grid);
            if (rc) return rc;
            argmax_generic_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(p, cs, ps, is);
        }
    }
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int confmat_launch(const void* pred, int pred_dtype, const void* target, int target_dtype, long long n, int C,
                   long long ignore_index, unsigned long long* confmat, unsigned long long* status, void* workspace,
                   cudaStream_t stream) {
    (void)workspace;
    CVCS_REQUIRE(((pred && target) || n == 0) && confmat, "cvcs_confmat: NULL pred/target/confmat");
    CVCS_REQUIRE(pred_dtype == CVCS_U8 || pred_dtype == CVCS_I64, "cvcs_confmat: pred dtype tag %d (want u8/i64)", pred_dtype);
    CVCS_REQUIRE(target_dtype == CVCS_U8 || target_dtype == CVCS_I64, "cvcs_confmat: target dtype tag %d (want u8/i64)", target_dtype);
    CVCS_REQUIRE(n >= 0 && C >= 1, "cvcs_confmat: bad n/C");
    if (C > 4096) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_confmat: C=%d > 4096", C);
    ConfParams p{};
    p.pred = pred;
    p.target = target;
    p.n = n;
    p.confmat = confmat;
    p.status = status;
    p.ignore_index = ignore_index;
    p.C = C;
    p.pred_i64 = pred_dtype == CVCS_I64;
    p.target_i64 = target_dtype == CVCS_I64;
    p.reps = shared_bin_replicas(C * C);
    const bool vec_ok = aligned_to(pred, 16) && aligned_to(target, 16);
    const bool priv = C * C <= 64;
    int grid = 0;
    int rc;
#define CVCS_LAUNCH_CONF(PRIV, VEC)                                                     \
    do {                                                                                \
        const int smem = BinAcc<PRIV>::smem_bytes(C * C, p.reps);                       \
        const long long blocks = (n / VEC + kThreads - 1) / kThreads;                   \
        rc = grid_for(confmat_kernel<PRIV, VEC>, smem, blocks, &grid);                  \
        if (rc) return rc;                                                              \
        confmat_kernel<PRIV, VEC><<<grid, kThreads, smem, stream>>>(p);                 \
    } while (0)
    if (vec_ok && !p.pred_i64 && !p.target_i64) {
        // u8 / u8: the fast kernel; at most 4 CTAs per SM (every CTA ends with one global atomic per bin)
        const int smem = priv ? BinAcc<true>::smem_bytes(C * C) : BinAcc<false>::smem_bytes(C * C, p.reps);
        long long blocks = (n / 32 + kThreads - 1) / kThreads;
        const long long cap = 4ll * num_sms();
        if (blocks > cap) blocks = cap;
        if (priv) {
            rc = grid_for(confmat_u8_kernel<true>, smem, blocks, &grid);
            if (rc) return rc;
            confmat_u8_kernel<true><<<grid, kThreads, smem, stream>>>(p);
        } else {
            rc = grid_for(confmat_u8_kernel<false>, smem, blocks, &grid);
            if (rc) return rc;
            confmat_u8_kernel<false><<<grid, kThreads, smem, stream>>>(p);
        }
        CVCS_CUDA_OK(cudaGetLastError());
        return CVCS_OK;
    }
    // 4 pixels per thread when both maps are u8 (32-bit loads) or any i64 (2 x 128-bit loads)
    if (vec_ok) {
        if (priv) CVCS_LAUNCH_CONF(true, 4);
        else CVCS_LAUNCH_CONF(false, 4);
    } else {
        if (priv) CVCS_LAUNCH_CONF(true, 1);
        else CVCS_LAUNCH_CONF(false, 1);
    }
#undef CVCS_LAUNCH_CONF
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

}  // namespace cvcs
