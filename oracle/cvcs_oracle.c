/*
 * cvcs_oracle.c — CPU restatement of the reference's per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cvcs_b200/ may import, link or call this file; it
 * is the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs), never the thing measured as the product or shipped.
 *
 * The reference (theElandor/CVCS, pure Python) delegates the arithmetic of this path to
 * third-party libraries that are not in its tree; each function restates the published
 * semantics of the library call made at the cited reference call site
 * (/root/reference/source/scripts/...):
 *
 *   oracle_cross_entropy  nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') forward and
 *                         its autograd backward        utils.py:223-242, train.py:122,125
 *                         (torch pinned 2.3.1 in README.MD:19; container has 2.11)
 *   oracle_argmax         torch.max(x, dim) / torch.argmax indices: first maximal element, NaN
 *                         is maximal                   utils.py:90,158,504; esa.py:56
 *   oracle_confmat        torchmetrics MulticlassConfusionMatrix.update on index inputs
 *                         (unpinned, README.MD:28; not installed): drop target==ignore_index,
 *                         M[t,p] += 1 via bincount(t*C+p)   utils.py:77-78,93-94
 *   oracle_label_hist     Loader._get_class_count      dataset.py:346-358
 *   oracle_tile           torchvision crop (zero padding out of bounds) + .type(float32) +
 *                         v2.Normalize                 dataset.py:28-32; train.py:121; nets.py:339-342
 *   oracle_vote           torch.mode over a stack of index maps   utils.py:504-507
 *   oracle_colorize       GID15Converter.iconvert      converters.py:23-36
 *   oracle_stitch         tile re-assembly             inference.py:40-57, utils.py:146,154
 *
 * Pinning: tests/test_oracle.py checks every function against golden vectors generated in the
 * build container by running the reference's own Python (utils.py / dataset.py imported
 * unmodified through oracle/ref_shim.py) and the torch / torchvision calls it makes
 * (tests/golden/make_golden.py, outputs committed under tests/golden/).
 *
 * Softmax/CE arithmetic is carried out in double precision: it is the mathematical definition,
 * against which both torch's fp32 result and the CUDA kernel's fp32 result are compared with
 * the 1e-5 relative tolerance BASELINE.json states.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* torch.max / argmax rule: first maximal index; NaN counts as maximal, first NaN wins. */
static int better_f(float v, float best) { return v > best || (v != v && best == best); }

/*
 * logits: fp32, element (b, c, pix) at logits[b*image_stride + c*class_stride + pix*pixel_stride]
 *         (NCHW: class_stride = HW, pixel_stride = 1; NHWC: class_stride = 1, pixel_stride = C)
 * target: int64 [B*HW]
 * weight: nullable fp32[C]
 * sums:   out double[3] = { Σ v w[y] nll, Σ v w[y], #labels outside [0,C) and != ignore_index }
 * dlogits: nullable, same indexing as logits, receives v w[y] (softmax - onehot) / Σ v w[y]
 * returns 0, or 1 if an out-of-bounds label was met (torch raises "Target N is out of bounds.")
 */
int oracle_cross_entropy(const float* logits, const int64_t* target, const float* weight, int64_t ignore_index,
                         int64_t B, int64_t C, int64_t HW, int64_t image_stride, int64_t class_stride,
                         int64_t pixel_stride, double* sums, float* dlogits) {
    double num = 0.0, den = 0.0;
    int64_t bad = 0;
    /* pass 1: the 'mean' divisor and the loss */
    for (int64_t b = 0; b < B; ++b) {
        double num_b = 0.0, den_b = 0.0;
        for (int64_t i = 0; i < HW; ++i) {
            const int64_t t = target[b * HW + i];
            if (t == ignore_index) continue;
            if (t < 0 || t >= C) {
                ++bad;
                continue;
            }
            const float* x = logits + b * image_stride + i * pixel_stride;
            double m = -INFINITY;
            for (int64_t c = 0; c < C; ++c) {
                const double v = x[c * class_stride];
                if (v > m || v != v) m = v;
            }
            double s = 0.0;
            for (int64_t c = 0; c < C; ++c) s += exp((double)x[c * class_stride] - m);
            const double lse = m + log(s);
            const double w = weight ? (double)weight[t] : 1.0;
            num_b += w * (lse - (double)x[t * class_stride]);
            den_b += w;
        }
        num += num_b;
        den += den_b;
    }
    sums[0] = num;
    sums[1] = den;
    sums[2] = (double)bad;
    if (dlogits) {
        for (int64_t b = 0; b < B; ++b)
            for (int64_t i = 0; i < HW; ++i) {
                const int64_t t = target[b * HW + i];
                const float* x = logits + b * image_stride + i * pixel_stride;
                float* d = dlogits + b * image_stride + i * pixel_stride;
                if (t == ignore_index || t < 0 || t >= C) {
                    for (int64_t c = 0; c < C; ++c) d[c * class_stride] = 0.0f;
                    continue;
                }
                double m = -INFINITY;
                for (int64_t c = 0; c < C; ++c) {
                    const double v = x[c * class_stride];
                    if (v > m || v != v) m = v;
                }
                double s = 0.0;
                for (int64_t c = 0; c < C; ++c) s += exp((double)x[c * class_stride] - m);
                const double g = (weight ? (double)weight[t] : 1.0) / den;
                for (int64_t c = 0; c < C; ++c) {
                    const double p = exp((double)x[c * class_stride] - m) / s;
                    d[c * class_stride] = (float)(g * (p - (c == t ? 1.0 : 0.0)));
                }
            }
    }
    return bad ? 1 : 0;
}

void oracle_argmax(const float* logits, int64_t B, int64_t C, int64_t HW, int64_t image_stride,
                   int64_t class_stride, int64_t pixel_stride, int64_t* out) {
    for (int64_t b = 0; b < B; ++b)
        for (int64_t i = 0; i < HW; ++i) {
            const float* x = logits + b * image_stride + i * pixel_stride;
            float best = x[0];
            int64_t arg = 0;
            for (int64_t c = 1; c < C; ++c)
                if (better_f(x[c * class_stride], best)) {
                    best = x[c * class_stride];
                    arg = c;
                }
            out[b * HW + i] = arg;
        }
}

/* has_ignore == 0 mirrors ignore_index=None.  Returns #pairs with t or p outside [0,C)
 * (torchmetrics would fail on those: bincount longer than C*C). */
int64_t oracle_confmat(const int64_t* preds, const int64_t* target, int64_t n, int64_t C, int has_ignore,
                       int64_t ignore_index, int64_t* confmat) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t t = target[i], p = preds[i];
        if (has_ignore && t == ignore_index) continue;
        if (t < 0 || t >= C || p < 0 || p >= C) {
            ++bad;
            continue;
        }
        confmat[t * C + p] += 1;
    }
    return bad;
}

/* hist[C+2]: classes, then #(== ignore_index outside [0,C)), then #(out of bounds). */
void oracle_label_hist(const uint8_t* labels, int64_t n, int64_t C, int64_t ignore_index, int64_t* hist) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t t = labels[i];
        if (t < C) hist[t] += 1;
        else if (t == ignore_index) hist[C] += 1;
        else hist[C + 1] += 1;
    }
}

/*
 * scene u8 [Cb,H,W]; tile k has top-left (yx[2k], yx[2k+1]) and size th x tw; pixels outside the
 * scene read as 0 (torchvision.transforms.functional.crop pads with zeros).
 * out fp32 [n,Cb,th,tw] = (float(x) - mean[cb]) / std[cb] with fp32 IEEE ops (mean == NULL: cast).
 * labels nullable u8 [H,W] -> label_out u8 [n,th,tw].
 */
void oracle_tile(const uint8_t* scene, int64_t Cb, int64_t H, int64_t W, const int32_t* yx, int64_t n, int64_t th,
                 int64_t tw, const float* mean, const float* stdv, float* out, const uint8_t* labels,
                 uint8_t* label_out) {
    for (int64_t k = 0; k < n; ++k)
        for (int64_t y = 0; y < th; ++y)
            for (int64_t x = 0; x < tw; ++x) {
                const int64_t sy = (int64_t)yx[2 * k] + y, sx = (int64_t)yx[2 * k + 1] + x;
                const int inside = sy >= 0 && sy < H && sx >= 0 && sx < W;
                for (int64_t cb = 0; cb < Cb; ++cb) {
                    volatile float v = inside ? (float)scene[(cb * H + sy) * W + sx] : 0.0f;
                    if (mean) {
                        volatile float d = v - mean[cb]; /* volatile: no contraction / reassociation */
                        v = d / stdv[cb];
                    }
                    out[((k * Cb + cb) * th + y) * tw + x] = v;
                }
                if (labels) label_out[(k * th + y) * tw + x] = inside ? labels[sy * W + sx] : 0;
            }
}

/* torch.mode over dim 0 of [n_maps, n]: most frequent value, smallest value on ties. */
void oracle_vote(const int64_t* maps, int64_t n_maps, int64_t n, int64_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t best_v = 0, best_c = 0;
        for (int64_t a = 0; a < n_maps; ++a) {
            const int64_t va = maps[a * n + i];
            int64_t cnt = 0;
            for (int64_t b = 0; b < n_maps; ++b) cnt += maps[b * n + i] == va;
            if (cnt > best_c || (cnt == best_c && va < best_v)) {
                best_c = cnt;
                best_v = va;
            }
        }
        out[i] = best_v;
    }
}

/* iconvert: output starts as ones, pixels with label k in [0,C) take lut[k]. */
void oracle_colorize(const int64_t* idx, int64_t n, const float* lut, int64_t C, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = idx[i];
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (v >= 0 && v < C) ? lut[3 * v + k] : 1.0f;
    }
}

/* paste the centred ch x cw window of every th x tw tile at (yx[2k], yx[2k+1]); clip to the scene */
void oracle_stitch(const uint8_t* tiles, int64_t n, int64_t th, int64_t tw, const int32_t* yx, int64_t ch, int64_t cw,
                   uint8_t* scene, int64_t H, int64_t W) {
    /* torchvision CenterCrop: int(round(d / 2.0)), Python rounds halves to even */
    const int64_t ky = (th - ch) / 2, kx = (tw - cw) / 2;
    const int64_t oy = ((th - ch) % 2 == 0 || ky % 2 == 0) ? ky : ky + 1;
    const int64_t ox = ((tw - cw) % 2 == 0 || kx % 2 == 0) ? kx : kx + 1;
    for (int64_t k = 0; k < n; ++k)
        for (int64_t y = 0; y < ch; ++y)
            for (int64_t x = 0; x < cw; ++x) {
                const int64_t sy = (int64_t)yx[2 * k] + y, sx = (int64_t)yx[2 * k + 1] + x;
                if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
                scene[sy * W + sx] = tiles[(k * th + oy + y) * tw + ox + x];
            }
}

/* dataset.py:11-16 _get_context: crop(image, tly - p, tlx - p, 3p, 3p) — zeros outside the scene — then the
 * reference's resizer (dataset.py:65,131; `v2` there is torchvision.transforms, the v1 API): uint8 -> float32,
 * torch's antialiased bilinear interpolate, torch.round (half to even), back to uint8
 * (torchvision/transforms/_functional_tensor.py resize: _cast_squeeze_in / interpolate(antialias=True) / _cast_squeeze_out).
 * The antialiased kernel (aten/src/ATen/native/cpu/UpSampleKernel.cpp `_compute_indices_min_size_weights_aa` with the
 * triangle filter, `basic_loop_aa_horizontal/vertical<float>`), all in float32: per output index i of an axis scaled
 * by s = in/out (= 3): centre c = s (i + 1/2), support s, taps j in [max(0, int(c - s + 1/2)), min(in, int(c + s + 1/2))),
 * weight max(0, 1 - |(j - c + 1/2) / s|) normalised to sum 1; each pass is x0 w0 + x1 w1 + ... left to right;
 * horizontal pass first.  Pinned against the reference itself: tests/golden/context_cases.npz (equal except at exact .5
 * ties of the filtered value, where torch's own rounding depends on FMA contraction in its vectorised loop).
 * scene u8 [Cb,H,W]; yx i32 [n,2] patch origins; out u8 [n,Cb,p,p]; pre (nullable) float [n,Cb,p,p] value before rounding. */
static int ctx_taps(int64_t in, int64_t outn, int64_t i, int* xmin_out, float* w) {
    const float scale = (float)in / (float)outn, support = scale;
    const float invscale = 1.0f / scale;
    const float center = scale * ((float)i + 0.5f);
    int64_t xmin = (int64_t)(center - support + 0.5f), xmax = (int64_t)(center + support + 0.5f);
    if (xmin < 0) xmin = 0;
    if (xmax > in) xmax = in;
    volatile float tot = 0.0f;
    const int n = (int)(xmax - xmin);
    for (int j = 0; j < n; ++j) {
        volatile float d = (float)(j + xmin) - center;
        volatile float v = (d + 0.5f) * invscale;
        if (v < 0) v = -v;
        w[j] = v < 1.0f ? 1.0f - v : 0.0f;
        tot = tot + w[j];
    }
    for (int j = 0; j < n; ++j) {
        volatile float q = w[j] / tot;
        w[j] = q;
    }
    *xmin_out = (int)xmin;
    return n;
}

int oracle_context(const uint8_t* scene, int64_t Cb, int64_t H, int64_t W, const int32_t* yx, int64_t n, int64_t p,
                   uint8_t* out, float* pre) {
    const int64_t in = 3 * p;
    int* xmin = (int*)malloc(sizeof(int) * p);
    int* cnt = (int*)malloc(sizeof(int) * p);
    float* wf = (float*)malloc(sizeof(float) * p * 16);
    uint8_t* crop = (uint8_t*)malloc((size_t)in * in);
    float* hbuf = (float*)malloc(sizeof(float) * (size_t)in * p);
    if (!xmin || !cnt || !wf || !crop || !hbuf) return -1;
    for (int64_t i = 0; i < p; ++i) {
        cnt[i] = ctx_taps(in, p, i, &xmin[i], wf + 16 * i);
        if (cnt[i] > 16) return -2;
    }
    for (int64_t k = 0; k < n; ++k)
        for (int64_t cb = 0; cb < Cb; ++cb) {
            const int64_t cy0 = (int64_t)yx[2 * k] - p, cx0 = (int64_t)yx[2 * k + 1] - p;
            for (int64_t y = 0; y < in; ++y)
                for (int64_t x = 0; x < in; ++x) {
                    const int64_t sy = cy0 + y, sx = cx0 + x;
                    crop[y * in + x] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? scene[(cb * H + sy) * W + sx] : 0;
                }
            for (int64_t y = 0; y < in; ++y)
                for (int64_t i = 0; i < p; ++i) {
                    volatile float acc = (float)crop[y * in + xmin[i]] * wf[16 * i];
                    for (int j = 1; j < cnt[i]; ++j) {
                        volatile float t = (float)crop[y * in + xmin[i] + j] * wf[16 * i + j]; /* no contraction */
                        acc = acc + t;
                    }
                    hbuf[y * p + i] = acc;
                }
            for (int64_t i = 0; i < p; ++i)
                for (int64_t x = 0; x < p; ++x) {
                    volatile float acc = hbuf[(int64_t)xmin[i] * p + x] * wf[16 * i];
                    for (int j = 1; j < cnt[i]; ++j) {
                        volatile float t = hbuf[((int64_t)xmin[i] + j) * p + x] * wf[16 * i + j];
                        acc = acc + t;
                    }
                    const size_t o = (size_t)(((k * Cb + cb) * p + i) * p + x);
                    if (pre) pre[o] = acc;
                    float r = rintf(acc); /* default rounding mode: to nearest even, as torch.round */
                    out[o] = (uint8_t)(r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r));
                }
        }
    free(xmin); free(cnt); free(wf); free(crop); free(hbuf);
    return 0;
}
