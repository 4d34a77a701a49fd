// common.cuh — shared device/host helpers for the cvcs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cvcs_b200.h"

namespace cvcs {

// ---- host-side error plumbing (abi.cu owns the thread-local buffer) ----------------------
int set_error(int code, const char* fmt, ...);
int num_sms();
int get_option(int option);  // cvcs_set_option values (0 = default / auto)

#define CVCS_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (void)cudaGetLastError(); /* do not leak a non-sticky error into the next call */ \
            return ::cvcs::set_error(CVCS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,          \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);            \
        }                                                                                    \
    } while (0)

#define CVCS_REQUIRE(cond, ...)                                              \
    do {                                                                     \
        if (!(cond)) return ::cvcs::set_error(CVCS_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

// ---- workspace layout -----------------------------------------------------------------------
// Zero-filled once by the caller; every kernel that uses a field restores it to zero.
constexpr int kMaxHistBins = 1032;  // C <= 1024 plus {ignored, out-of-bounds}
constexpr int kMaxGrid = 4096;  // upper bound on persistent grid size (148 SMs x <=16 CTAs, rounded)
constexpr int kMaxPreGrid = 1024;  // the TMA-staged K1 runs at most 4 CTAs per SM
struct Workspace {
    unsigned int ticket;          // last-block election counter
    unsigned int next_chunk;      // K1 (TMA variant): dynamic chunk claim counter
    unsigned long long bad;       // out-of-bounds label counter
    unsigned int gbar;            // K1 with its own label pre-pass: grid-wide arrival counter (zero on exit)
    unsigned int pad0;
    unsigned long long pad1[5];
    double partial[2 * kMaxGrid]; // per-block {Σ w·nll, Σ w}
    unsigned long long hist[kMaxHistBins];  // per-call label histogram (K4 / K5), zero on exit
    double pre[2][kMaxPreGrid];   // K1 (TMA variant) per-block Σ v·w[y]: [0] this batch's pre-pass, [1] the NEXT batch's scan
};
constexpr size_t kWorkspaceBytes = 128 * 1024;
static_assert(sizeof(Workspace) <= kWorkspaceBytes, "workspace too small");

constexpr int kThreads = 256;  // CTA width of every persistent kernel
constexpr int kWarps = kThreads / 32;

// replicas of a shared-memory u32 bin array: 8 (one per warp) while that stays small, halved until
// it fits `budget`; 0 = not even one copy fits (every hit becomes a global atomic)
inline int shared_bin_replicas(int nbins, int budget_bytes = 64 * 1024, int hard_limit_bytes = 160 * 1024) {
    int r = kWarps;
    while (r > 1 && static_cast<long long>(nbins) * r * 4 > budget_bytes) r >>= 1;
    if (static_cast<long long>(nbins) * r * 4 > hard_limit_bytes) return 0;
    return r;
}

// ---- Σw exchange block (one per rank, in device memory mapped into every peer: CUDA IPC over NVLink) ---------------
// Rank r publishes its Σ v·w[y] of exchange number `seq` in slot [seq % kXDepth][r] of EVERY rank's block and then
// releases flags[seq % kXDepth][r] = seq; a consumer spins on the N flags of its own block and adds the N values in
// rank order, so every rank divides by the bit-identical total.  A launch that also computed the NEXT batch's sum
// (next_target) publishes it for exchange seq + 1 as it ends — a whole step ahead of its readers — so that in a
// pipelined sequence no kernel ever waits for a peer.  Depth 8: a rank may run up to 7 exchanges ahead of
// the slowest reader before it would overwrite an unread slot (a reader that finds a newer sequence number reports it).
constexpr int kXDepth = 8;
constexpr int kXMaxRanks = 16;
struct XchgBlock {
    double slots[kXDepth][kXMaxRanks];
    unsigned int flags[kXDepth][kXMaxRanks];
    unsigned long long seq;        // exchanges completed by THIS rank (advanced by the last CTA of each launch)
    unsigned long long errors;     // time-outs / overruns seen by this rank's readers
    unsigned long long published;  // highest exchange number whose value THIS rank has already sent to its peers
};

// Pass-end sums (the C x C confusion matrix, the per-step loss table): a one-shot all-reduce of up to kXWideN doubles
// over the same peer-mapped allocation — every rank stores its vector into slot [parity][rank] of every rank's region,
// releases a flag, waits for the N flags in its own region and adds the N vectors in rank order (bit-identical result
// on every rank).  One 256-thread CTA, no NCCL kernel: ~10 us instead of a 50-100 us collective launch.
constexpr int kXWideN = 2048;
struct XchgWide {
    double data[2][kXMaxRanks][kXWideN];
    unsigned int flags[2][kXMaxRanks];
    unsigned long long seq;
};
struct XchgRegion {
    XchgBlock block;    // first: a peer pointer to the region is a pointer to its block
    XchgWide wide;
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    // round-to-nearest-even, NaN preserving (cvt.rn.bf16x2.f32 packs {hi, lo})
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// Streaming 128/64/32-bit global accesses (.cs = evict-first: every byte is touched once).
template <int BYTES>
struct Raw;
template <>
struct Raw<16> {
    uint4 v;
    __device__ __forceinline__ void load(const void* p) { v = __ldcs(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void store(void* p) const { __stcs(reinterpret_cast<uint4*>(p), v); }
    __device__ __forceinline__ uint32_t& word(int i) { return (&v.x)[i]; }
};
template <>
struct Raw<8> {
    uint2 v;
    __device__ __forceinline__ void load(const void* p) { v = __ldcs(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ void store(void* p) const { __stcs(reinterpret_cast<uint2*>(p), v); }
    __device__ __forceinline__ uint32_t& word(int i) { return (&v.x)[i]; }
};
template <>
struct Raw<4> {
    uint32_t v;
    __device__ __forceinline__ void load(const void* p) { v = __ldcs(reinterpret_cast<const uint32_t*>(p)); }
    __device__ __forceinline__ void store(void* p) const { __stcs(reinterpret_cast<uint32_t*>(p), v); }
    __device__ __forceinline__ uint32_t& word(int) { return v; }
};
template <>
struct Raw<2> {
    uint16_t v;
    __device__ __forceinline__ void load(const void* p) { v = __ldcs(reinterpret_cast<const unsigned short*>(p)); }
    __device__ __forceinline__ void store(void* p) const { __stcs(reinterpret_cast<unsigned short*>(p), v); }
};
template <>
struct Raw<1> {
    uint8_t v;
    __device__ __forceinline__ void load(const void* p) { v = __ldcs(reinterpret_cast<const unsigned char*>(p)); }
    __device__ __forceinline__ void store(void* p) const { __stcs(reinterpret_cast<unsigned char*>(p), v); }
};

// Load / store VEC consecutive logits of type T as fp32.
template <typename T, int VEC>
struct VecIO;

template <int VEC>
struct VecIO<float, VEC> {
    static_assert(VEC == 1 || VEC == 2 || VEC == 4, "f32 VEC");
    __device__ static __forceinline__ void load(const float* p, float (&x)[VEC]) {
        Raw<4 * VEC> r;
        r.load(p);
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = __uint_as_float(r.word(i));
    }
    __device__ static __forceinline__ void store(float* p, const float (&x)[VEC]) {
        Raw<4 * VEC> r;
#pragma unroll
        for (int i = 0; i < VEC; ++i) r.word(i) = __float_as_uint(x[i]);
        r.store(p);
    }
};

template <>
struct VecIO<__nv_bfloat16, 1> {
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&x)[1]) {
        Raw<2> r;
        r.load(p);
        x[0] = __uint_as_float(static_cast<uint32_t>(r.v) << 16);
    }
    __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&x)[1]) {
        Raw<2> r;
        r.v = static_cast<uint16_t>(pack_bf16(x[0], 0.f) & 0xffffu);
        r.store(p);
    }
};
template <int VEC>
struct VecIO<__nv_bfloat16, VEC> {
    static_assert(VEC == 2 || VEC == 4 || VEC == 8, "bf16 VEC");
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&x)[VEC]) {
        Raw<2 * VEC> r;
        r.load(p);
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            x[2 * i] = bf16_lo(r.word(i));
            x[2 * i + 1] = bf16_hi(r.word(i));
        }
    }
    __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&x)[VEC]) {
        Raw<2 * VEC> r;
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) r.word(i) = pack_bf16(x[2 * i], x[2 * i + 1]);
        r.store(p);
    }
};

// Labels: VEC consecutive u8 or i64 -> int (values outside int range are clamped to -2 so
// that they can never alias a valid class; ignore_index is matched on the 64-bit value).
template <int VEC>
__device__ __forceinline__ void load_labels_u8(const uint8_t* p, int (&t)[VEC]) {
    Raw<VEC> r;
    r.load(p);
    if constexpr (VEC == 1) {
        t[0] = r.v;
    } else if constexpr (VEC == 2) {
        t[0] = r.v & 0xff;
        t[1] = r.v >> 8;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) t[i] = (r.word(i / 4) >> (8 * (i % 4))) & 0xff;
    }
}

template <int VEC>
__device__ __forceinline__ void load_labels_i64(const long long* p, long long ignore_index,
                                                int ignore_tag, int (&t)[VEC]) {
    // Maps each label to an int: in-range values stay, ignore_index -> ignore_tag,
    // anything else outside int range -> -2 (out of bounds).
    if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            Raw<16> r;
            r.load(p + 2 * i);
            long long a = (static_cast<long long>(r.v.y) << 32) | r.v.x;
            long long b = (static_cast<long long>(r.v.w) << 32) | r.v.z;
            t[2 * i] = a == ignore_index ? ignore_tag : ((a < 0 || a > 0x7fffffff) ? -2 : static_cast<int>(a));
            t[2 * i + 1] = b == ignore_index ? ignore_tag : ((b < 0 || b > 0x7fffffff) ? -2 : static_cast<int>(b));
        }
    } else {
        Raw<8> r;
        r.load(p);
        long long a = (static_cast<long long>(r.v.y) << 32) | r.v.x;
        t[0] = a == ignore_index ? ignore_tag : ((a < 0 || a > 0x7fffffff) ? -2 : static_cast<int>(a));
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned int warp_sum(unsigned int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// torch.max / argmax tie+NaN rule: first maximal index, NaN is maximal, first NaN wins.
__device__ __forceinline__ bool better(float v, float best) {
    return v > best || (v != v && best == best);
}

// ---- shared-memory bin accumulators (confusion matrix, label histogram) ------------------------
// Private mode: u16 cnt[bin][thread] (no atomics, no same-address serialisation however blocky the
// label map is); shared mode: u32 bins[rep][nb] with shared-memory atomics, warp w using replica
// w % reps (reps = 8 for small matrices, fewer when C*C*4 B would not fit); reps == 0: the bins do
// not fit shared memory at all and every hit is a 64-bit global atomic.  CTAs are kThreads wide.
// BAR = 0: the accumulator is used by the whole CTA (__syncthreads); BAR > 0: by threads
// 0..kThreads-1 of a wider CTA, synchronised on named barrier BAR.
template <bool PRIV, int BAR = 0, typename CT = unsigned short>
struct BinAcc {
    static constexpr unsigned int kMaxCount = sizeof(CT) == 1 ? 255u : 65535u;   // private counter range
    static __device__ __forceinline__ void sync() {
        if constexpr (BAR == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(kThreads) : "memory");
    }
    CT* cnt16;
    CT* mine;                  // PRIV: cnt16 + threadIdx.x (this thread's counter of bin 0)
    unsigned int* bins32;
    unsigned long long* direct;
    int nb;
    int reps;

    __device__ __forceinline__ void init(unsigned char* smem, int nbins, int replicas = kWarps,
                                         unsigned long long* global_bins = nullptr) {
        nb = nbins;
        reps = replicas;
        direct = global_bins;
        if constexpr (PRIV) {
            cnt16 = reinterpret_cast<CT*>(smem);
            mine = cnt16 + threadIdx.x;
            uint32_t* z = reinterpret_cast<uint32_t*>(smem);
            for (int i = threadIdx.x; i < nb * kThreads * static_cast<int>(sizeof(CT)) / 4; i += kThreads) z[i] = 0u;
        } else {
            bins32 = reinterpret_cast<unsigned int*>(smem);
            for (int i = threadIdx.x; i < nb * reps; i += kThreads) bins32[i] = 0u;
        }
        sync();
    }
    __device__ __forceinline__ void add(int key, unsigned int n = 1u) {
        if constexpr (PRIV) {
            CT* c = mine + key * kThreads;
            *c = static_cast<CT>(*c + n);
        } else {
            if (reps) atomicAdd(bins32 + ((threadIdx.x >> 5) & (reps - 1)) * nb + key, n);
            else atomicAdd(direct + key, static_cast<unsigned long long>(n));
        }
    }
    // (Round 2 measured warp-aggregated updates for the shared mode — match.any + popc/redux, leaders only, with and
    // without atomics — on B200: the match instruction costs far more than the same-address atomics it removes (C=16
    // metrics mode 0.94 -> 0.46 of the copy peak, C=20 0.81 -> 0.74; profiles/README.md), so the plain per-lane
    // shared-memory atomic on a per-warp replica stays.)
    static __host__ __device__ int smem_bytes(int nbins, int replicas = kWarps) {
        return PRIV ? nbins * kThreads * static_cast<int>(sizeof(CT)) : nbins * replicas * 4;
    }
    // CTA-wide flush to global u64 bins; leaves the counters zeroed.
    __device__ __forceinline__ void flush(unsigned long long* confmat) {
        sync();
        if constexpr (PRIV) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int b = warp; b < nb; b += kWarps) {
                unsigned int s = 0;
#pragma unroll
                for (int j = 0; j < kThreads / 32; ++j) {
                    CT* c = cnt16 + b * kThreads + j * 32 + lane;
                    s += *c;
                    *c = 0;
                }
                s = warp_sum(s);
                if (lane == 0 && s) atomicAdd(confmat + b, static_cast<unsigned long long>(s));
            }
        } else {
            for (int b = threadIdx.x; b < nb; b += kThreads) {
                unsigned int s = 0;
                for (int w = 0; w < reps; ++w) {
                    s += bins32[w * nb + b];
                    bins32[w * nb + b] = 0u;
                }
                if (s) atomicAdd(confmat + b, static_cast<unsigned long long>(s));
            }
        }
        sync();
    }
};

#endif  // __CUDACC__
}  // namespace cvcs
