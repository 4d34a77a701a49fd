// context_kernel.cu — N4: the context view of a patch (reference dataset.py:11-16 `_get_context`):
//     context = Resize(p)( crop(image, tly - p, tlx - p, 3p, 3p) )
// i.e. the 3p x 3p neighbourhood of the patch (zeros outside the scene, torchvision crop semantics) reduced to
// p x p.  The reference's resizer is `torchvision.transforms.Resize` (dataset.py imports torchvision.transforms as
// v2): a uint8 tensor is cast to float32, resized by torch's antialiased bilinear kernel and rounded back
// (round-half-to-even).  The antialiased kernel is a triangle filter whose support is the scale factor (3 input pixels
// each side), applied separably — horizontal pass first — with per-output-pixel normalised float32 weights computed
// exactly as aten does (UpSampleKernel.cpp `_compute_indices_min_size_weights_aa`): for the 3:1 ratio every interior
// output has the five taps [1 2 3 2 1]/9 starting one pixel left of its 3-pixel cell; the first and last outputs lose
// the tap outside the crop and are renormalised ([2 3 2 1]/8, [1 2 3 2]/8).  Each pass is the plain left-to-right
// float32 sum x0*w0 + x1*w1 + ... (separate multiply and add, no contraction).  Results equal the reference's bytes
// except at exact .5 ties of the filtered value, where torch's own output depends on whether its build fuses the
// multiply-add for that pixel (AVX-512 main loop vs scalar tail) — about 1 pixel in 10^6 (tests/test_gpu_context.py).
//
// One CTA produces a 32 x 32 block of one band of one context tile: the 98 x 98 input window is gathered into shared
// memory row by row (one warp per row, zero filled outside the scene), reduced horizontally to 98 x 32 floats, then
// vertically to 32 x 32 bytes.  HBM traffic: 9 bytes read + 1 written per output pixel and band (the windows of
// neighbouring blocks overlap by 2 rows / columns).
#include <math.h>

#include "common.cuh"

namespace cvcs {
namespace {

constexpr int kTapsMax = 6;
struct Taps {
    int off[3];            // first tap relative to 3*i, for i = 0 / interior / last
    int n[3];              // taps in the row
    float w[3][kTapsMax];  // normalised float32 weights
};

constexpr int kBX = 32, kBY = 32;                // output block
constexpr int kWX = 3 * kBX + 2, kWY = 3 * kBY + 2;   // input window (one extra tap each side)
constexpr int kWinWords = (kWX + 3 + 3) / 4;          // 32-bit words per window row, whatever its byte phase (0..3)
constexpr int kWinStride = 4 * kWinWords + 4;         // + the spare columns the zero-weight taps may touch

__global__ void __launch_bounds__(kThreads) context_kernel(const unsigned char* __restrict__ scene, int Cb, int H, int W,
                                                           const int* __restrict__ tile_yx, const int* __restrict__ tile_slot,
                                                           int p, unsigned char* __restrict__ out, const Taps taps_arg,
                                                           const int words_ok) {
    // Every tap loop below runs the full five taps: rows with fewer carry zero weights, and adding x * 0 (x a byte or
    // a finite partial sum, never negative) leaves the float32 sum bit-identical.  The window therefore keeps two
    // spare columns and the horizontal result one spare (zero) row for the taps that fall past the end.
    __shared__ __align__(16) unsigned char win[kWY][kWinStride];
    __shared__ float hbuf[kWY + 1][kBX];
    __shared__ float wts[3][5];                 // indexed by row type at run time
    __shared__ int offs[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 15) {
        const int row = tid / 5, k = tid % 5;
        wts[row][k] = k < taps_arg.n[row] ? taps_arg.w[row][k] : 0.f;
    }
    if (tid < 3) offs[tid] = taps_arg.off[tid];
    if (tid < kBX) hbuf[kWY][tid] = 0.f;
    const int bx = blockIdx.x * kBX, by = blockIdx.y * kBY;     // output block origin inside the tile
    const int tile = blockIdx.z / Cb, band = blockIdx.z % Cb;
    const int slot = tile_slot ? tile_slot[tile] : tile;
    // crop origin in scene coordinates: (tly - p, tlx - p); window origin inside the crop: 3*b - 1
    const int cy0 = tile_yx[2 * tile] - p, cx0 = tile_yx[2 * tile + 1] - p;
    const int wy0 = 3 * by - 1, wx0 = 3 * bx - 1;
    const unsigned char* __restrict__ plane = scene + static_cast<size_t>(band) * H * W;
    // Window rows start at any byte phase of the scene row.  When every scene row is word aligned (W and H*W multiples
    // of 4, aligned base — the usual case) a row is fetched as the 26 aligned 32-bit words that cover it, the bytes
    // outside the scene or the crop masked to zero, and lands in shared memory at the same phase (`shift`); otherwise
    // byte by byte.  Which bytes are valid does not depend on the row: decided once per lane.
    int shift = 0;
    if (words_ok) {
        const int sx0 = cx0 + wx0;
        shift = sx0 & 3;                                           // floor alignment, also for negative origins
        const int xs = sx0 - shift + 4 * lane;                     // scene column of this lane's word
        uint32_t mask = 0u;
        if (lane < kWinWords) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = 4 * lane + b - shift, xx = wx0 + c, sx = xs + b;
                if (c >= 0 && c < kWX && xx >= 0 && xx < 3 * p && sx >= 0 && sx < W) mask |= 0xffu << (8 * b);
            }
        }
        for (int r = warp; r < kWY; r += kThreads / 32) {
            const int yy = wy0 + r, sy = cy0 + yy;                 // row inside the 3p x 3p crop / the scene
            const bool row_ok = yy >= 0 && yy < 3 * p && sy >= 0 && sy < H;
            if (lane < kWinWords) {
                uint32_t w = 0u;
                if (row_ok && mask) w = __ldg(reinterpret_cast<const unsigned int*>(plane + static_cast<long long>(sy) * W + xs)) & mask;
                reinterpret_cast<uint32_t*>(win[r])[lane] = w;
            }
        }
    } else {
        // a lane owns window columns lane, lane + 32, ...
        constexpr int kCols = (kWX + 31) / 32;
        bool cok[kCols];
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
            const int c = lane + 32 * j, xx = wx0 + c, sx = cx0 + xx;
            cok[j] = c < kWX && xx >= 0 && xx < 3 * p && sx >= 0 && sx < W;
        }
        for (int r = warp; r < kWY; r += kThreads / 32) {
            const int yy = wy0 + r, sy = cy0 + yy;
            const bool row_ok = yy >= 0 && yy < 3 * p && sy >= 0 && sy < H;
            const unsigned char* __restrict__ src = plane + static_cast<long long>(row_ok ? sy : 0) * W + (cx0 + wx0);
#pragma unroll
            for (int j = 0; j < kCols; ++j) {
                const int c = lane + 32 * j;
                if (c < kWX) win[r][c] = (row_ok && cok[j]) ? src[c] : static_cast<unsigned char>(0);
            }
        }
    }
    __syncthreads();
    {
        // horizontal pass: a thread keeps ONE output column (kThreads is a multiple of kBX), so its row type, first
        // window column and weights are fixed; it walks down the window rows
        const int x = tid % kBX, ox = bx + x;
        const int row = ox == 0 ? 0 : (ox == p - 1 ? 2 : 1);
        const int c0 = 3 * x + offs[row] + 1 + shift;
        const float w0 = wts[row][0], w1 = wts[row][1], w2 = wts[row][2], w3 = wts[row][3], w4 = wts[row][4];
        for (int r = tid / kBX; r < kWY; r += kThreads / kBX) {
            const unsigned char* q = &win[r][c0];
            float acc = __fmul_rn(static_cast<float>(q[0]), w0);
            acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(q[1]), w1));
            acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(q[2]), w2));
            acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(q[3]), w3));
            acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(q[4]), w4));
            hbuf[r][x] = acc;
        }
    }
    __syncthreads();
    {
        const int x = tid % kBX, ox = bx + x;
        for (int y = tid / kBX; y < kBY; y += kThreads / kBX) {
            const int oy = by + y;
            if (ox < p && oy < p) {
                const int row = oy == 0 ? 0 : (oy == p - 1 ? 2 : 1);
                const int r0 = 3 * y + offs[row] + 1;
                float acc = __fmul_rn(hbuf[r0][x], wts[row][0]);
#pragma unroll
                for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(hbuf[r0 + k][x], wts[row][k]));
                const float r = rintf(acc);                       // torch.round: half to even
                out[((static_cast<size_t>(slot) * Cb + band) * p + oy) * p + ox] = static_cast<unsigned char>(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
            }
        }
    }
}

// torch's antialias weights (aten UpSampleKernel.cpp `_compute_indices_min_size_weights_aa`, bilinear triangle filter),
// float32 arithmetic throughout as aten does for float32 input, for input size 3p -> output size p.
void taps_for(int p, int i, int* xmin_out, int* n_out, float* w_out) {
    const float scale = static_cast<float>(3 * p) / static_cast<float>(p), support = scale;
    const float invscale = 1.0f / scale;
    const float center = scale * (static_cast<float>(i) + 0.5f);
    int xmin = static_cast<int>(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5f);
    if (xmax > 3 * p) xmax = 3 * p;
    const int n = xmax - xmin;
    volatile float tot = 0.f;
    for (int j = 0; j < n; ++j) {
        volatile float d = static_cast<float>(j + xmin) - center;
        volatile float v = (d + 0.5f) * invscale;
        if (v < 0.f) v = -v;
        w_out[j] = v < 1.0f ? 1.0f - v : 0.0f;
        tot = tot + w_out[j];
    }
    for (int j = 0; j < n; ++j) {
        volatile float q = w_out[j] / tot;
        w_out[j] = q;
    }
    *xmin_out = xmin;
    *n_out = n;
}

}  // namespace

int context_launch(const unsigned char* scene, int Cb, int H, int W, const int* tile_yx, const int* tile_slot, int n_tiles,
                   int p, unsigned char* out, cudaStream_t stream) {
    CVCS_REQUIRE(scene && tile_yx && out, "cvcs_tile_context: NULL argument");
    CVCS_REQUIRE(Cb >= 1 && H > 0 && W > 0 && n_tiles >= 0 && p >= 3, "cvcs_tile_context: bad shape (Cb %d, %d x %d, %d tiles, p %d >= 3)", Cb, H, W, n_tiles, p);
    if (n_tiles == 0) return CVCS_OK;
    CVCS_REQUIRE(static_cast<long long>(n_tiles) * Cb <= 65535, "cvcs_tile_context: n_tiles * Cb = %lld exceeds the grid's z range; split the call",
                 static_cast<long long>(n_tiles) * Cb);
    Taps t{};
    const int rows[3] = {0, 1, p - 1};
    for (int r = 0; r < 3; ++r) {
        float w[16];
        int xmin = 0;
        taps_for(p, rows[r], &xmin, &t.n[r], w);
        while (t.n[r] > 1 && w[t.n[r] - 1] == 0.0f) --t.n[r];      // aten's tap range ends on a zero weight: x * 0 adds nothing
        if (t.n[r] > 5) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_tile_context: %d non-zero taps (the input window holds 5)", t.n[r]);
        if (t.n[r] > kTapsMax) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_tile_context: %d taps", t.n[r]);
        t.off[r] = xmin - 3 * rows[r];
        for (int j = 0; j < t.n[r]; ++j) t.w[r][j] = w[j];
    }
    dim3 grid((p + kBX - 1) / kBX, (p + kBY - 1) / kBY, n_tiles * Cb);
    const int words_ok = (W % 4 == 0 && (static_cast<long long>(H) * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(scene) & 3u) == 0) ? 1 : 0;
    context_kernel<<<grid, kThreads, 0, stream>>>(scene, Cb, H, W, tile_yx, tile_slot, p, out, t, words_ok);
    CVCS_CUDA_OK(cudaGetLastError());
    return CVCS_OK;
}

}  // namespace cvcs
