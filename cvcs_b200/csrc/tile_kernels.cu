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
    long long n_items;  // n_tiles * tile_h * groups_per_row
    int Cb, H, W, tile_h, tile_w, groups_per_row, n_tiles;
    int out_dtype;        // CVCS_U8 / F32 / BF16
    int label_out_i64;
    int hist_C;
};

template <int VEC>
__device__ __forceinline__ void load_u8_row(const unsigned char* __restrict__ plane, int H, int W, int y, int x0,
                                            unsigned int (&px)[VEC]) {
    // zero fill outside the scene (torchvision.transforms.functional.crop semantics)
    if (y < 0 || y >= H) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) px[k] = 0u;
        return;
    }
    const unsigned char* row = plane + static_cast<long long>(y) * W;
    if constexpr (VEC == 4) {
        if (x0 >= 0 && x0 + 3 < W && ((reinterpret_cast<uintptr_t>(row + x0) & 3u) == 0)) {
            const unsigned int w = __ldcs(reinterpret_cast<const unsigned int*>(row + x0));
            px[0] = w & 0xff;
            px[1] = (w >> 8) & 0xff;
            px[2] = (w >> 16) & 0xff;
            px[3] = w >> 24;
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int x = x0 + k;
        px[k] = (x >= 0 && x < W) ? row[x] : 0u;
    }
}

template <int VEC, bool PRIV>
__global__ void __launch_bounds__(kThreads) tile_kernel(const TileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int is_last;
    const bool do_hist = p.hist != nullptr;
    BinAcc<PRIV> acc;
    if (do_hist) acc.init(smem, p.hist_C + 2);
    unsigned int since_flush = 0;
    const long long plane = static_cast<long long>(p.H) * p.W;
    const long long tile_plane = static_cast<long long>(p.tile_h) * p.tile_w;
    const int ign_outside = (p.hist_ignore >= p.hist_C && p.hist_ignore <= 255) ? static_cast<int>(p.hist_ignore) : -1;

    for (long long base = static_cast<long long>(blockIdx.x) * kThreads; base < p.n_items;
         base += static_cast<long long>(gridDim.x) * kThreads) {
        const long long item = base + threadIdx.x;
        if (item < p.n_items) {
            const unsigned int it = static_cast<unsigned int>(item);
            const unsigned int rowi = it / p.groups_per_row;       // tile * tile_h + y
            const int xg = static_cast<int>(it - rowi * p.groups_per_row);
            const unsigned int tile = rowi / p.tile_h;
            const int y = static_cast<int>(rowi - tile * p.tile_h);
            const int sy = __ldg(p.tile_yx + 2 * tile) + y;
            const int sx = __ldg(p.tile_yx + 2 * tile + 1) + xg * VEC;
            const long long opix = static_cast<long long>(y) * p.tile_w + xg * VEC;  // inside a tile plane
            const long long slot = p.tile_slot ? __ldg(p.tile_slot + tile) : static_cast<long long>(tile);

            for (int cb = 0; cb < p.Cb; ++cb) {
                unsigned int px[VEC];
                load_u8_row<VEC>(p.scene + cb * plane, p.H, p.W, sy, sx, px);
                const long long o = (slot * p.Cb + cb) * tile_plane + opix;
                if (p.out_dtype == CVCS_U8) {
                    unsigned char* out = reinterpret_cast<unsigned char*>(p.out) + o;
                    if constexpr (VEC == 4) {
                        __stcs(reinterpret_cast<unsigned int*>(out), px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24));
                    } else {
                        out[0] = static_cast<unsigned char>(px[0]);
                    }
                } else {
                    float f[VEC];
                    if (p.mean) {
                        const float m = __ldg(p.mean + cb), s = __ldg(p.stdv + cb);
#pragma unroll
                        for (int k = 0; k < VEC; ++k)
                            f[k] = __fdiv_rn(__fsub_rn(static_cast<float>(px[k]), m), s);  // IEEE, as sub_().div_()
                    } else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) f[k] = static_cast<float>(px[k]);
                    }
                    if (p.out_dtype == CVCS_F32) VecIO<float, VEC>::store(reinterpret_cast<float*>(p.out) + o, f);
                    else VecIO<__nv_bfloat16, VEC>::store(reinterpret_cast<__nv_bfloat16*>(p.out) + o, f);
                }
            }
            if (p.label) {
                unsigned int lb[VEC];
                load_u8_row<VEC>(p.label, p.H, p.W, sy, sx, lb);
                const long long o = slot * tile_plane + opix;
                if (p.label_out) {
                    if (p.label_out_i64) {
                        long long* out = reinterpret_cast<long long*>(p.label_out) + o;
                        if constexpr (VEC == 4) {
                            Raw<16> r;
                            r.v = make_uint4(lb[0], 0u, lb[1], 0u);
                            r.store(out);
                            r.v = make_uint4(lb[2], 0u, lb[3], 0u);
                            r.store(out + 2);
                        } else {
                            out[0] = lb[0];
                        }
                    } else {
                        unsigned char* out = reinterpret_cast<unsigned char*>(p.label_out) + o;
                        if constexpr (VEC == 4) {
                            __stcs(reinterpret_cast<unsigned int*>(out), lb[0] | (lb[1] << 8) | (lb[2] << 16) | (lb[3] << 24));
                        } else {
                            out[0] = static_cast<unsigned char>(lb[0]);
                        }
                    }
                }
                if (do_hist) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const int t = static_cast<int>(lb[k]);
                        acc.add(t < p.hist_C ? t : (t == ign_outside ? p.hist_C : p.hist_C + 1));
                    }
                }
            }
        }
        if (PRIV && do_hist) {
            since_flush += VEC;
            if (since_flush > 65535u - VEC) {
                acc.flush(p.ws->hist);
                since_flush = 0;
            }
        }
    }
    if (!do_hist) return;
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

// ---- N2 stitch -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) stitch_kernel(const unsigned char* __restrict__ tiles, int n_tiles, int th,
                                                          int tw, const int* __restrict__ yx, int ch, int cw,
                                                          unsigned char* __restrict__ scene, int H, int W) {
    const long long per_tile = static_cast<long long>(ch) * cw;
    const long long total = per_tile * n_tiles;
    const int oy = (th - ch) / 2, ox = (tw - cw) / 2;  // CenterCrop offsets (torchvision rounds (th-ch)/2.0 to even... see host)
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

int simple_grid(long long n) {
    long long blocks = (n + kThreads - 1) / kThreads;
    long long g = static_cast<long long>(num_sms()) * 8;
    if (g > blocks) g = blocks;
    if (g < 1) g = 1;
    return static_cast<int>(g);
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
    p.out_dtype = out_dtype;
    p.label_out_i64 = label_out_dtype == CVCS_I64;
    p.hist_C = hist_C;
    const int vec = vec4 ? 4 : 1;
    p.groups_per_row = tile_w / vec;
    p.n_items = rows * p.groups_per_row;
    CVCS_REQUIRE(p.n_items < (1ll << 32), "cvcs_tile_normalize: too many work items");
    const bool priv = !hist || hist_C + 2 <= 64;
    const int smem = hist ? (priv ? BinAcc<true>::smem_bytes(hist_C + 2) : BinAcc<false>::smem_bytes(hist_C + 2)) : 0;
    const long long blocks = (p.n_items + kThreads - 1) / kThreads;
    long long g = static_cast<long long>(num_sms()) * (smem > 24 * 1024 ? 4 : 8);
    if (g > blocks) g = blocks;
    if (g > kMaxGrid) g = kMaxGrid;
    const int grid = static_cast<int>(g);
    if (vec4) {
        if (priv) tile_kernel<4, true><<<grid, kThreads, smem, stream>>>(p);
        else tile_kernel<4, false><<<grid, kThreads, smem, stream>>>(p);
    } else {
        if (priv) tile_kernel<1, true><<<grid, kThreads, smem, stream>>>(p);
        else tile_kernel<1, false><<<grid, kThreads, smem, stream>>>(p);
    }
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

int stitch_launch(const unsigned char* tiles, int n_tiles, int th, int tw, const int* yx, int ch, int cw,
                  unsigned char* scene, int H, int W, cudaStream_t stream) {
    CVCS_REQUIRE(tiles && yx && scene, "cvcs_stitch: NULL argument");
    CVCS_REQUIRE(n_tiles >= 0 && th > 0 && tw > 0 && ch > 0 && cw > 0 && ch <= th && cw <= tw && H > 0 && W > 0, "cvcs_stitch: bad shape");
    if (n_tiles == 0) return CVCS_OK;
    const long long total = static_cast<long long>(n_tiles) * ch * cw;
    stitch_kernel<<<simple_grid(total), kThreads, 0, stream>>>(tiles, n_tiles, th, tw, yx, ch, cw, scene, H, W);
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
