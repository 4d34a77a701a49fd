// tile_kernels.cu — the data-format kernels either side of K1:
//   K5 tile gather + cast + per-band normalise  <- _get_cropped_data (torchvision crop, zero
//        padding) dataset.py:28-32,136-150; image.type(float32) train.py:121;
//        SegformerMod.preprocessor = ToDtype(float32) + (x - mean) / std  nets.py:339-342
//   N2 stitch   <- tile re-assembly, inference.py:40-57 (+ CenterCrop of utils.py:146,154)
//   N3 vote     <- Ensemble majority vote via torch.mode, utils.py:499-507
//   N4 colorize <- GID15Converter.iconvert, converters.py:23-36
// Pure byte/gather work: HBM-bound, coalesced 32-bit u8 reads (a warp covers 128 contiguous
// bytes of a scene row) and 128-bit fp32 / 64-bit bf16 writes (512 / 256 contiguous bytes).
#include "common.cuh"

namespace cvcs {
namespace {

struct TileParams {
    const unsigned char* scene;
    const int* tile_yx;
    const int* tile_slot;  // nullable: output slot of tile i (default i)
    const float* mean;
    const float* stdv;
    void* out;
    const unsigned char* label;
    void* label_out;
    unsigned long long* hist;  // nullable
    Workspace* ws;
    long long hist_ignore;
    long long n_rows;     // n_tiles * planes * tile_h  (one "row item" = one tile row of one plane)
    int Cb, H, W, tile_h, tile_w, n_tiles;
    int planes;           // Cb (+1 when the label plane is gathered in the same launch)
    int out_dtype;        // CVCS_U8 / F32 / BF16
    int label_out_i64;
    int hist_C;
    int use_lut;          // per-band 256-entry LUT of (v - mean) / std in shared memory
};

constexpr int kSpan = 4;                 // pixels per thread per access (one 32-bit u8 load)
constexpr int kUnroll = 4;               // independent accesses per thread per step
constexpr int kRowBlock = 32 * kSpan * kUnroll;  // pixels of a tile row one warp covers per step (512)

enum : int { kOutU8 = 0, kOutF32 = 1, kOutBF16 = 2 };
enum : int { kCast = 0, kLut = 1, kDiv = 2 };  // float conversion: plain cast / shared-memory LUT / IEEE sub+div

// 4 consecutive u8 of a scene row starting at column x0 (zero fill outside the scene, as
// torchvision.transforms.functional.crop pads); `row` is NULL for rows outside the scene.
__device__ __forceinline__ uint32_t load_u8x4(const unsigned char* __restrict__ row, int W, int x0) {
    if (row == nullptr) return 0u;
    if (x0 >= 0 && x0 + 3 < W && ((reinterpret_cast<uintptr_t>(row + x0) & 3u) == 0))
        return __ldcs(reinterpret_cast<const unsigned int*>(row + x0));
    uint32_t w = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + k;
        if (x >= 0 && x < W) w |= static_cast<uint32_t>(row[x]) << (8 * k);
    }
    return w;
}

// 4 packed u8 -> the output dtype, written with one (u8: 32-bit, bf16: 64-bit, f32: 128-bit) store
template <int OUT, int MODE>
__device__ __forceinline__ void emit4(void* __restrict__ out, long long o, uint32_t w, const float* __restrict__ lut,
                                      float mean, float stdv) {
    if constexpr (OUT == kOutU8) {
        __stcs(reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(out) + o), w);
    } else {
        float f[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = (w >> (8 * k)) & 0xff;
            if constexpr (MODE == kLut) f[k] = lut[b];
            else if constexpr (MODE == kDiv) f[k] = __fdiv_rn(__fsub_rn(static_cast<float>(b), mean), stdv);  // IEEE, as sub_().div_()
            else f[k] = static_cast<float>(b);
        }
        if constexpr (OUT == kOutF32) VecIO<float, 4>::store(reinterpret_cast<float*>(out) + o, f);
        else VecIO<__nv_bfloat16, 4>::store(reinterpret_cast<__nv_bfloat16*>(out) + o, f);
    }
}

// One warp gathers one tile row of one plane per step: lane l handles the 4-pixel groups
// l, l + 32, l + 64, l + 96 of each 512-pixel block, so every load instruction reads 128
// contiguous bytes of the scene row and every store instruction writes 512 (fp32) / 256 (bf16) /
// 128 (u8) contiguous bytes of the tile row, with kUnroll independent accesses in flight per thread.
// Rows that lie inside the scene with a 4-byte aligned start (every tile of the regular grid when
// W % 4 == 0) take a branch-free path; shifted / overhanging tiles take the general one.
template <int OUT, int MODE, bool PRIV>
__global__ void __launch_bounds__(kThreads) tile_kernel(const TileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int is_last;
    const bool do_hist = p.hist != nullptr;
    float* lut_all = reinterpret_cast<float*>(smem);
    const int lut_bytes = MODE == kLut ? p.Cb * 256 * 4 : 0;
    if constexpr (MODE == kLut) {
        // IEEE sub / div exactly as torch's sub_(mean).div_(std): 256 possible inputs per band
        for (int i = threadIdx.x; i < p.Cb * 256; i += kThreads) {
            const int cb = i >> 8;
            lut_all[i] = __fdiv_rn(__fsub_rn(static_cast<float>(i & 255), __ldg(p.mean + cb)), __ldg(p.stdv + cb));
        }
    }
    BinAcc<PRIV> acc;
    if (do_hist) acc.init(smem + lut_bytes, p.hist_C + 2);  // contains the barrier
    else __syncthreads();
    const long long plane = static_cast<long long>(p.H) * p.W;
    const long long tile_plane = static_cast<long long>(p.tile_h) * p.tile_w;
    const int ign_outside = (p.hist_ignore >= p.hist_C && p.hist_ignore <= 255) ? static_cast<int>(p.hist_ignore) : -1;
    const int lane = threadIdx.x & 31;
    const long long warp0 = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    const long long n_warps = static_cast<long long>(gridDim.x) * kWarps;

    for (long long item = warp0; item < p.n_rows; item += n_warps) {
        // item -> (tile, plane, y): plane-major inside a tile so that a warp's consecutive items
        // walk down the rows of one plane
        const unsigned int it = static_cast<unsigned int>(item);
        const unsigned int tp = it / p.tile_h;
        const int y = static_cast<int>(it - tp * p.tile_h);
        const unsigned int tile = tp / p.planes;
        const int pl = static_cast<int>(tp - tile * p.planes);
        const bool is_label = pl == p.Cb;
        const int sy = __ldg(p.tile_yx + 2 * tile) + y;
        const int sx0 = __ldg(p.tile_yx + 2 * tile + 1);
        const long long slot = p.tile_slot ? __ldg(p.tile_slot + tile) : static_cast<long long>(tile);
        const unsigned char* src_plane = is_label ? p.label : p.scene + pl * plane;
        const unsigned char* row = (sy >= 0 && sy < p.H) ? src_plane + static_cast<long long>(sy) * p.W : nullptr;
        const long long obase = (is_label ? slot : slot * p.Cb + pl) * tile_plane + static_cast<long long>(y) * p.tile_w;
        const bool fast = row != nullptr && sx0 >= 0 && sx0 + p.tile_w <= p.W && ((reinterpret_cast<uintptr_t>(row + sx0) & 3u) == 0);
        const unsigned int* row32 = reinterpret_cast<const unsigned int*>(row + sx0);  // only dereferenced when `fast`
        const float* lut = lut_all + (pl << 8);
        float mean = 0.f, stdv = 1.f;
        if constexpr (MODE == kDiv) {
            if (!is_label) {
                mean = __ldg(p.mean + pl);
                stdv = __ldg(p.stdv + pl);
            }
        }

        for (int xb = 0; xb < p.tile_w; xb += kRowBlock) {
            uint32_t w[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int x = xb + (u * 32 + lane) * kSpan;
                if (fast) w[u] = (x < p.tile_w) ? __ldcs(row32 + (x >> 2)) : 0u;
                else w[u] = (x < p.tile_w) ? load_u8x4(row, p.W, sx0 + x) : 0u;
            }
            if (!is_label) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int x = xb + (u * 32 + lane) * kSpan;
                    if (x < p.tile_w) emit4<OUT, MODE>(p.out, obase + x, w[u], lut, mean, stdv);
                }
            } else {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int x = xb + (u * 32 + lane) * kSpan;
                    if (x >= p.tile_w) continue;
                    const long long o = obase + x;
                    if (p.label_out) {
                        if (p.label_out_i64) {
                            long long* out = reinterpret_cast<long long*>(p.label_out) + o;
                            Raw<16> r;
                            r.v = make_uint4(w[u] & 0xff, 0u, (w[u] >> 8) & 0xff, 0u);
                            r.store(out);
                            r.v = make_uint4((w[u] >> 16) & 0xff, 0u, w[u] >> 24, 0u);
                            r.store(out + 2);
                        } else {
                            __stcs(reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(p.label_out) + o), w[u]);
                        }
                    }
                    if (do_hist) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int t = static_cast<int>((w[u] >> (8 * k)) & 0xff);
                            acc.add(t < p.hist_C ? t : (t == ign_outside ? p.hist_C : p.hist_C + 1));
                        }
                    }
                }
            }
        }
    }
    if (!do_hist) return;
    // PRIV counters are u16: the launcher bounds the label pixels per lane so they cannot overflow
    acc.flush(p.ws->hist);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&p.ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int b = threadIdx.x; b < p.hist_C + 2; b += kThreads) {
        const unsigned long long cnt = __ldcg(&p.ws->hist[b]);
        p.ws->hist[b] = 0ull;
        if (cnt) atomicAdd(p.hist + b, cnt);
    }
    if (threadIdx.x == 0) {
        p.ws->ticket = 0u;
        __threadfence();
    }
}

template <int OUT, int MODE>
int launch_tile(const TileParams& p, bool priv, int grid, int smem, cudaStream_t stream) {
    if (priv) {
        if (smem > 48 * 1024) CVCS_CUDA_OK(cudaFuncSetAttribute(tile_kernel<OUT, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tile_kernel<OUT, MODE, true><<<grid, kThreads, smem, stream>>>(p);
    } else {
        if (smem > 48 * 1024) CVCS_CUDA_OK(cudaFuncSetAttribute(tile_kernel<OUT, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tile_kernel<OUT, MODE, false><<<grid, kThreads, smem, stream>>>(p);
    }
    return CVCS_OK;
}

// Scalar variant for tile widths that are not a multiple of 4 (or unaligned output pointers): one
// pixel per thread, same item order.
__global__ void __launch_bounds__(kThreads) tile_kernel_scalar(const TileParams p) {
    const long long plane = static_cast<long long>(p.H) * p.W;
    const long long tile_plane = static_cast<long long>(p.tile_h) * p.tile_w;
    const long long total = p.n_rows * p.tile_w;
    const int ign_outside = (p.hist_ignore >= p.hist_C && p.hist_ignore <= 255) ? static_cast<int>(p.hist_ignore) : -1;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const long long item = i / p.tile_w;
        const int x = static_cast<int>(i - item * p.tile_w);
        const long long tp = item / p.tile_h;
        const int y = static_cast<int>(item - tp * p.tile_h);
        const long long tile = tp / p.planes;
        const int pl = static_cast<int>(tp - tile * p.planes);
        const bool is_label = pl == p.Cb;
        const int sy = p.tile_yx[2 * tile] + y, sx = p.tile_yx[2 * tile + 1] + x;
        const long long slot = p.tile_slot ? p.tile_slot[tile] : tile;
        const unsigned char* src_plane = is_label ? p.label : p.scene + pl * plane;
        const unsigned int v = (sy >= 0 && sy < p.H && sx >= 0 && sx < p.W) ? src_plane[static_cast<long long>(sy) * p.W + sx] : 0u;
        const long long o = (is_label ? slot : slot * p.Cb + pl) * tile_plane + static_cast<long long>(y) * p.tile_w + x;
        if (is_label) {
            if (p.label_out) {
                if (p.label_out_i64) reinterpret_cast<long long*>(p.label_out)[o] = v;
                else reinterpret_cast<unsigned char*>(p.label_out)[o] = static_cast<unsigned char>(v);
            }
            if (p.hist) {
                const int t = static_cast<int>(v);
                atomicAdd(p.hist + (t < p.hist_C ? t : (t == ign_outside ? p.hist_C : p.hist_C + 1)), 1ull);
            }
        } else if (p.out_dtype == CVCS_U8) {
            reinterpret_cast<unsigned char*>(p.out)[o] = static_cast<unsigned char>(v);
        } else {
            float f = static_cast<float>(v);
            if (p.mean) f = __fdiv_rn(__fsub_rn(f, p.mean[pl]), p.stdv[pl]);
            if (p.out_dtype == CVCS_F32) reinterpret_cast<float*>(p.out)[o] = f;
            else reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(f);
        }
    }
}

int simple_grid(long long n) {
    long long blocks = (n + kThreads - 1) / kThreads;
    long long g = static_cast<long long>(num_sms()) * 8;
    if (g > blocks) g = blocks;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// ---- N2 stitch -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) stitch_kernel(const unsigned char* __restrict__ tiles, int n_tiles, int th,
                                                          int tw, const int* __restrict__ yx, int ch, int cw,
                                                          unsigned char* __restrict__ scene, int H, int W, int oy, int ox) {
    const long long per_tile = static_cast<long long>(ch) * cw;
    const long long total = per_tile * n_tiles;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const int t = static_cast<int>(i / per_tile);
        const int r = static_cast<int>(i - t * per_tile);
        const int y = r / cw, x = r - y * cw;
        const int sy = yx[2 * t] + y, sx = yx[2 * t + 1] + x;
        if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
        scene[static_cast<long long>(sy) * W + sx] = tiles[(static_cast<long long>(t) * th + oy + y) * tw + ox + x];
    }
}

// 4 pixels per thread: crop width, crop offset, tile width and scene width multiples of 4 and aligned
// bases; a group whose destination is not 4-aligned or leaves the scene falls back to byte stores.
__global__ void __launch_bounds__(kThreads) stitch_x4_kernel(const unsigned char* __restrict__ tiles, int n_tiles, int th,
                                                             int tw, const int* __restrict__ yx, int ch, int cw,
                                                             unsigned char* __restrict__ scene, int H, int W, int oy, int ox) {
    const int gpr = cw / 4;  // groups per cropped row
    const long long per_tile = static_cast<long long>(ch) * gpr;
    const long long total = per_tile * n_tiles;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const int t = static_cast<int>(i / per_tile);
        const int r = static_cast<int>(i - t * per_tile);
        const int y = r / gpr, x = (r - y * gpr) * 4;
        const int sy = __ldg(yx + 2 * t) + y, sx = __ldg(yx + 2 * t + 1) + x;
        if (sy < 0 || sy >= H) continue;
        const unsigned int w = __ldcs(reinterpret_cast<const unsigned int*>(tiles + (static_cast<long long>(t) * th + oy + y) * tw + ox + x));
        unsigned char* dst = scene + static_cast<long long>(sy) * W + sx;
        if (sx >= 0 && sx + 3 < W && (sx & 3) == 0) {
            *reinterpret_cast<unsigned int*>(dst) = w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (sx + k >= 0 && sx + k < W) dst[k] = static_cast<unsigned char>(w >> (8 * k));
        }
    }
}

// 16 pixels per thread (one 128-bit load, one 128-bit store): everything a multiple of 16 — crop width and offset,
// tile and scene width, the tile origins' x — which is the usual case (tiles of 224 / 256 / 1024 on a tile-sized grid).
// A thread walks down `rows` consecutive rows of its 16-pixel column, so the index arithmetic is paid once per thread.
__global__ void __launch_bounds__(kThreads) stitch_x16_kernel(const unsigned char* __restrict__ tiles, int n_tiles, int th,
                                                              int tw, const int* __restrict__ yx, int ch, int cw,
                                                              unsigned char* __restrict__ scene, int H, int W, int oy, int ox, int rows) {
    const int gpr = cw / 16;                         // 16-pixel groups per cropped row
    const int bands = (ch + rows - 1) / rows;        // row bands per tile
    const long long per_tile = static_cast<long long>(bands) * gpr;
    const long long total = per_tile * n_tiles;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const int t = static_cast<int>(i / per_tile);
        const int r = static_cast<int>(i - t * per_tile);
        const int band = r / gpr, x = (r - band * gpr) * 16;
        const int y0 = band * rows;
        const int sy0 = __ldg(yx + 2 * t) + y0, sx = __ldg(yx + 2 * t + 1) + x;
        const unsigned char* src = tiles + (static_cast<long long>(t) * th + oy + y0) * tw + ox + x;
        unsigned char* dst = scene + static_cast<long long>(sy0) * W + sx;
        const int ny = min(rows, ch - y0);
        const bool whole = sx >= 0 && sx + 15 < W && (sx & 15) == 0;
#pragma unroll 4
        for (int y = 0; y < ny; ++y) {
            const int sy = sy0 + y;
            if (sy < 0 || sy >= H) continue;
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src + static_cast<long long>(y) * tw));
            unsigned char* d = dst + static_cast<long long>(y) * W;
            if (whole) {
                *reinterpret_cast<uint4*>(d) = v;
            } else {
                const unsigned int w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (sx + k >= 0 && sx + k < W) d[k] = static_cast<unsigned char>(w4[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

// ---- N3 vote -------------------------------------------------------------------------------------
template <typename IT, typename OT>
__global__ void __launch_bounds__(kThreads) vote_kernel(const IT* __restrict__ maps, int n_maps, long long n, OT* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        long long best_v = 0;
        int best_c = 0;
        for (int a = 0; a < n_maps; ++a) {
            const long long va = static_cast<long long>(maps[a * n + i]);
            int cnt = 0;
            for (int b = 0; b < n_maps; ++b) cnt += (static_cast<long long>(maps[b * n + i]) == va) ? 1 : 0;
            // most frequent value; ties -> smallest value (torch.mode)
            if (cnt > best_c || (cnt == best_c && va < best_v)) {
                best_c = cnt;
                best_v = va;
            }
        }
        out[i] = static_cast<OT>(best_v);
    }
}

// u8 maps, n_maps <= 8: byte-parallel voting, four pixels per 32-bit word.  For every pair of maps one word
// operation chain marks the byte lanes where they agree (exact zero-byte test of the XOR) and bumps both maps' per-lane
// counters; the winner per pixel is the maximum of the 16-bit keys (count << 8 | 255 - value) — most frequent value,
// ties to the smallest (torch.mode, utils.py:506) — taken two pixels at a time with the packed u16 max.  Every map is
// read exactly once: a thread takes 16 consecutive pixels (one 128-bit load per map) or, in the tail kernel, 4.
template <int NM>
__device__ __forceinline__ uint32_t vote_word(const uint32_t (&w)[NM]) {
    uint32_t cnt[NM];
#pragma unroll
    for (int a = 0; a < NM; ++a) cnt[a] = 0x01010101u;
#pragma unroll
    for (int a = 0; a < NM; ++a)
#pragma unroll
        for (int b = a + 1; b < NM; ++b) {
            const uint32_t x = w[a] ^ w[b];
            const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;          // bit 7 of a lane: its low 7 bits are non-zero
            const uint32_t e = (~(t | x) & 0x80808080u) >> 7;            // 1 in the lanes where the two maps agree
            cnt[a] += e;
            cnt[b] += e;
        }
    uint32_t best_lo = 0u, best_hi = 0u;                                   // pixels (0, 1) and (2, 3) as u16 lanes
#pragma unroll
    for (int a = 0; a < NM; ++a) {
        const uint32_t klo = __byte_perm(cnt[a], 0u, 0x1404) | (__byte_perm(w[a], 0u, 0x4140) ^ 0x00ff00ffu);
        const uint32_t khi = __byte_perm(cnt[a], 0u, 0x3424) | (__byte_perm(w[a], 0u, 0x4342) ^ 0x00ff00ffu);
        best_lo = __vmaxu2(best_lo, klo);
        best_hi = __vmaxu2(best_hi, khi);
    }
    return ~__byte_perm(best_lo, best_hi, 0x6420);                         // the low byte of every lane is 255 - value
}

template <int NM>
__global__ void __launch_bounds__(kThreads) vote_u8x16_kernel(const uint8_t* __restrict__ maps, long long n,
                                                              uint8_t* __restrict__ out) {
    const long long n16 = n / 16;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n16;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        uint4 v[NM];
#pragma unroll
        for (int a = 0; a < NM; ++a) v[a] = __ldcs(reinterpret_cast<const uint4*>(maps + a * n) + i);
        uint4 res;
        {
            uint32_t w[NM];
#pragma unroll
            for (int a = 0; a < NM; ++a) w[a] = v[a].x;
            res.x = vote_word<NM>(w);
#pragma unroll
            for (int a = 0; a < NM; ++a) w[a] = v[a].y;
            res.y = vote_word<NM>(w);
#pragma unroll
            for (int a = 0; a < NM; ++a) w[a] = v[a].z;
            res.z = vote_word<NM>(w);
#pragma unroll
            for (int a = 0; a < NM; ++a) w[a] = v[a].w;
            res.w = vote_word<NM>(w);
        }
        __stcs(reinterpret_cast<uint4*>(out) + i, res);
    }
}

template <int NM>
__global__ void __launch_bounds__(kThreads) vote_u8x4_kernel(const uint8_t* __restrict__ maps, long long n,
                                                             uint8_t* __restrict__ out) {
    const long long n4 = n / 4;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        uint32_t w[NM];
#pragma unroll
        for (int a = 0; a < NM; ++a) w[a] = __ldcs(reinterpret_cast<const unsigned int*>(maps + a * n) + i);
        __stcs(reinterpret_cast<unsigned int*>(out) + i, vote_word<NM>(w));
    }
}

template <int NM>
bool try_vote_u8x4(const void* maps, int n_maps, long long n, void* out, cudaStream_t stream, bool wide) {
    if constexpr (NM > 8) {
        return false;
    } else {
        if (n_maps == NM) {
            if (wide) vote_u8x16_kernel<NM><<<simple_grid(n / 16), kThreads, 0, stream>>>(reinterpret_cast<const uint8_t*>(maps), n, reinterpret_cast<uint8_t*>(out));
            else vote_u8x4_kernel<NM><<<simple_grid(n / 4), kThreads, 0, stream>>>(reinterpret_cast<const uint8_t*>(maps), n, reinterpret_cast<uint8_t*>(out));
            return true;
        }
        return try_vote_u8x4<NM + 1>(maps, n_maps, n, out, stream, wide);
    }
}

// ---- N4 colorize ---------------------------------------------------------------------------------
template <typename IT>
__global__ void __launch_bounds__(kThreads) colorize_kernel(const IT* __restrict__ idx, long long n,
                                                            const float* __restrict__ lut, int C, float* __restrict__ out) {
    extern __shared__ float slut[];
    for (int i = threadIdx.x; i < 3 * C; i += kThreads) slut[i] = lut[i];
    __syncthreads();
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const long long v = static_cast<long long>(idx[i]);
        float r = 1.f, g = 1.f, b = 1.f;  // iconvert starts from torch.ones
        if (v >= 0 && v < C) {
            r = slut[3 * v];
            g = slut[3 * v + 1];
            b = slut[3 * v + 2];
        }
        out[3 * i] = r;
        out[3 * i + 1] = g;
        out[3 * i + 2] = b;
    }
}


}  // namespace

int tile_launch(const unsigned char* scene, int Cb, int H, int W, const int* tile_yx, const int* tile_slot, int n_tiles, int tile_h,
                int tile_w, const float* mean, const float* stdv, void* out, int out_dtype,
                const unsigned char* label, void* label_out, int label_out_dtype, unsigned long long* hist,
                int hist_C, long long hist_ignore, void* workspace, cudaStream_t stream) {
    CVCS_REQUIRE(scene && tile_yx && out, "cvcs_tile_normalize: NULL scene/tile_yx/out");
    CVCS_REQUIRE(Cb >= 1 && H > 0 && W > 0 && n_tiles >= 0 && tile_h > 0 && tile_w > 0, "cvcs_tile_normalize: bad shape");
    CVCS_REQUIRE(out_dtype == CVCS_U8 || out_dtype == CVCS_F32 || out_dtype == CVCS_BF16, "cvcs_tile_normalize: out dtype tag %d", out_dtype);
    CVCS_REQUIRE((mean == nullptr) == (stdv == nullptr), "cvcs_tile_normalize: mean and std must both be given or both NULL");
    CVCS_REQUIRE(!(out_dtype == CVCS_U8 && mean), "cvcs_tile_normalize: u8 output cannot be normalised");
    CVCS_REQUIRE(!label_out || label, "cvcs_tile_normalize: label_out without label scene");
    CVCS_REQUIRE(!label_out || label_out_dtype == CVCS_U8 || label_out_dtype == CVCS_I64, "cvcs_tile_normalize: label out dtype tag %d", label_out_dtype);
    CVCS_REQUIRE(!hist || (label && workspace && hist_C >= 1), "cvcs_tile_normalize: hist needs label scene, workspace and hist_C");
    if (hist && hist_C + 2 > kMaxHistBins) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_tile_normalize: hist_C too large");
    if (n_tiles == 0) return CVCS_OK;
    const long long rows = static_cast<long long>(n_tiles) * tile_h;
    CVCS_REQUIRE(rows * tile_w < (1ll << 33), "cvcs_tile_normalize: too many output pixels");

    const int osz = out_dtype == CVCS_F32 ? 4 : (out_dtype == CVCS_BF16 ? 2 : 1);
    auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const bool vec4 = tile_w % 4 == 0 && al(out, 4 * static_cast<size_t>(osz)) &&
                      al(label_out, label_out_dtype == CVCS_I64 ? 16 : 4);
    TileParams p{};
    p.scene = scene;
    p.tile_yx = tile_yx;
    p.tile_slot = tile_slot;
    p.mean = mean;
    p.stdv = stdv;
    p.out = out;
    p.label = label;
    p.label_out = label_out;
    p.hist = hist;
    p.ws = reinterpret_cast<Workspace*>(workspace);
    p.hist_ignore = hist_ignore;
    p.Cb = Cb;
    p.H = H;
    p.W = W;
    p.tile_h = tile_h;
    p.tile_w = tile_w;
    p.n_tiles = n_tiles;
    p.planes = Cb + ((label && (label_out || hist)) ? 1 : 0);
    p.n_rows = static_cast<long long>(n_tiles) * p.planes * tile_h;
    p.out_dtype = out_dtype;
    p.label_out_i64 = label_out_dtype == CVCS_I64;
    p.hist_C = hist_C;
    p.use_lut = (mean != nullptr && Cb <= 32) ? 1 : 0;
    CVCS_REQUIRE(p.n_rows < (1ll << 32), "cvcs_tile_normalize: too many work items");
    if (!vec4) {
        const long long total = p.n_rows * tile_w;
        long long g = (total + kThreads - 1) / kThreads;
        const long long cap = static_cast<long long>(num_sms()) * 8;
        if (g > cap) g = cap;
        tile_kernel_scalar<<<static_cast<int>(g < 1 ? 1 : g), kThreads, 0, stream>>>(p);
        CVCS_CUDA_OK(cudaGetLastError());
        return CVCS_OK;
    }
    const bool want_priv = !hist || hist_C + 2 <= 64;
    const int lut_bytes = p.use_lut ? Cb * 256 * 4 : 0;
    // grid: as many CTAs as stay resident; with private u16 histogram counters a warp must not
    // see more than 65535 label pixels per lane, which bounds the row items per warp
    int ctas_per_sm = get_option(CVCS_OPT_TILE_CTAS);
    // 4 CTAs (32 warps) per SM: measured best for this write-heavy stream (2: 0.71, 4: 0.94, 8: 0.81 of the copy peak)
    if (ctas_per_sm < 1 || ctas_per_sm > 8) ctas_per_sm = 4;
    long long g = static_cast<long long>(num_sms()) * ctas_per_sm;
    const long long warps_needed = p.n_rows;
    if (g * kWarps > warps_needed) g = (warps_needed + kWarps - 1) / kWarps;
    if (g > kMaxGrid) g = kMaxGrid;
    if (g < 1) g = 1;
    bool priv = want_priv;
    if (hist && priv) {
        const long long items_per_warp = (p.n_rows + g * kWarps - 1) / (g * kWarps);
        const long long px_per_lane = items_per_warp * ((tile_w + kRowBlock - 1) / kRowBlock) * kUnroll * 4;
        if (px_per_lane > 60000) priv = false;  // fall back to shared-memory atomics (no overflow possible)
    }
    const int smem = lut_bytes + (hist ? (priv ? BinAcc<true>::smem_bytes(hist_C + 2) : BinAcc<false>::smem_bytes(hist_C + 2)) : 0);
    const int grid = static_cast<int>(g);
    const int mode = p.use_lut ? kLut : (mean ? kDiv : kCast);
    int rc = CVCS_OK;
    if (out_dtype == CVCS_U8) rc = launch_tile<kOutU8, kCast>(p, priv, grid, smem, stream);
    else if (out_dtype == CVCS_F32)
        rc = mode == kLut ? launch_tile<kOutF32, kLut>(p, priv, grid, smem, stream)
                          : (mode == kDiv ? launch_tile<kOutF32, kDiv>(p, priv, grid, smem, stream) : launch_tile<kOutF32, kCast>(p, priv, grid, smem, stream));
    else
        rc = mode == kLut ? launch_tile<kOutBF16, kLut>(p, priv, grid, smem, stream)
                          : (mode == kDiv ? launch_tile<kOutBF16, kDiv>(p, priv, grid, smem, stream) : launch_tile<kOutBF16, kCast>(p, priv, grid, smem, stream));
    if (rc) return rc;
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int stitch_launch(const unsigned char* tiles, int n_tiles, int th, int tw, const int* yx, int ch, int cw,
                  unsigned char* scene, int H, int W, cudaStream_t stream) {
    CVCS_REQUIRE(tiles && yx && scene, "cvcs_stitch: NULL argument");
    CVCS_REQUIRE(n_tiles >= 0 && th > 0 && tw > 0 && ch > 0 && cw > 0 && ch <= th && cw <= tw && H > 0 && W > 0, "cvcs_stitch: bad shape");
    if (n_tiles == 0) return CVCS_OK;
    const long long total = static_cast<long long>(n_tiles) * ch * cw;
    // torchvision's CenterCrop offset (utils.py:146,154): int(round(d / 2.0)) with Python's round-half-to-even
    auto center = [](int d) { const int k = d / 2; return (d % 2 == 0 || k % 2 == 0) ? k : k + 1; };
    const int oy = center(th - ch), ox = center(tw - cw);
    const bool x4 = cw % 4 == 0 && tw % 4 == 0 && ox % 4 == 0 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(tiles) & 3u) == 0 &&
                    (reinterpret_cast<uintptr_t>(scene) & 3u) == 0;
    const bool x16 = cw % 16 == 0 && tw % 16 == 0 && ox % 16 == 0 && W % 16 == 0 && (reinterpret_cast<uintptr_t>(tiles) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(scene) & 15u) == 0;
    if (x16) {
        const int rows = 4;      // rows per thread: four independent 128-bit loads in flight
        const long long items = static_cast<long long>(n_tiles) * ((ch + rows - 1) / rows) * (cw / 16);
        stitch_x16_kernel<<<simple_grid(items), kThreads, 0, stream>>>(tiles, n_tiles, th, tw, yx, ch, cw, scene, H, W, oy, ox, rows);
    } else if (x4) stitch_x4_kernel<<<simple_grid(total / 4), kThreads, 0, stream>>>(tiles, n_tiles, th, tw, yx, ch, cw, scene, H, W, oy, ox);
    else stitch_kernel<<<simple_grid(total), kThreads, 0, stream>>>(tiles, n_tiles, th, tw, yx, ch, cw, scene, H, W, oy, ox);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int vote_launch(const void* maps, int dtype, int n_maps, long long n, int C, void* out, int out_dtype,
                cudaStream_t stream) {
    (void)C;
    CVCS_REQUIRE(maps && out && n_maps >= 1 && n >= 0, "cvcs_vote: bad argument");
    CVCS_REQUIRE(dtype == CVCS_U8 || dtype == CVCS_I64, "cvcs_vote: dtype tag %d", dtype);
    CVCS_REQUIRE(out_dtype == CVCS_U8 || out_dtype == CVCS_I64, "cvcs_vote: out dtype tag %d", out_dtype);
    const int g = simple_grid(n);
    auto al4 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 3u) == 0; };
    const bool wide = n % 16 == 0 && (reinterpret_cast<uintptr_t>(maps) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (dtype == CVCS_U8 && out_dtype == CVCS_U8 && n_maps <= 8 && n % 4 == 0 && al4(maps) && al4(out) &&
        try_vote_u8x4<1>(maps, n_maps, n, out, stream, wide)) {
        CVCS_CUDA_OK(cudaGetLastError());
        return CVCS_OK;
    }
    if (dtype == CVCS_U8 && out_dtype == CVCS_U8)
        vote_kernel<<<g, kThreads, 0, stream>>>(reinterpret_cast<const uint8_t*>(maps), n_maps, n, reinterpret_cast<uint8_t*>(out));
    else if (dtype == CVCS_U8)
        vote_kernel<<<g, kThreads, 0, stream>>>(reinterpret_cast<const uint8_t*>(maps), n_maps, n, reinterpret_cast<long long*>(out));
    else if (out_dtype == CVCS_U8)
        vote_kernel<<<g, kThreads, 0, stream>>>(reinterpret_cast<const long long*>(maps), n_maps, n, reinterpret_cast<uint8_t*>(out));
    else
        vote_kernel<<<g, kThreads, 0, stream>>>(reinterpret_cast<const long long*>(maps), n_maps, n, reinterpret_cast<long long*>(out));
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

// (Round 2 tried four pixels per thread with three 128-bit stores: 48-byte-strided 16-byte stores fill half a sector
// each and ran 1.8x SLOWER than the three scalar stores below, which the L1 merges into whole lines — 79 vs 45 us on
// a 16.8 Mpixel map, profiles/r2c.)
int colorize_launch(const void* idx, int dtype, long long n, const float* lut, int C, float* out, cudaStream_t stream) {
    CVCS_REQUIRE(idx && lut && out && n >= 0 && C >= 1 && C <= 4096, "cvcs_colorize: bad argument");
    CVCS_REQUIRE(dtype == CVCS_U8 || dtype == CVCS_I64, "cvcs_colorize: dtype tag %d", dtype);
    const int g = simple_grid(n);
    const int smem = 3 * C * 4;
    if (dtype == CVCS_U8) colorize_kernel<<<g, kThreads, smem, stream>>>(reinterpret_cast<const uint8_t*>(idx), n, lut, C, out);
    else colorize_kernel<<<g, kThreads, smem, stream>>>(reinterpret_cast<const long long*>(idx), n, lut, C, out);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

}  // namespace cvcs
