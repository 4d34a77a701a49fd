/*
 * cvcs_b200.h — C-ABI of the B200-native per-pixel hot path of theElandor/CVCS.
 *
 * The reference (pure Python) has no FFI; its seam for this path is a handful of
 * PyTorch library calls.  Each entry point below replaces one of those call sites
 * (citations are into /root/reference/source/scripts/):
 *
 *   cvcs_label_hist       <- Loader._get_class_count            dataset.py:346-358
 *   cvcs_total_weight     <- the "mean" divisor of nn.CrossEntropyLoss (utils.py:230,238)
 *   cvcs_labels_prepare   <- mask.type(torch.long) + that divisor          train.py:122
 *   cvcs_ce_fused         <- crit(mask_pred, mask.long())       train.py:122, utils.py:120
 *                            loss.backward()                    train.py:125
 *                            torch.max(y_pred, dim=0)           utils.py:90
 *                            MulticlassConfusionMatrix.update   utils.py:93-94
 *   cvcs_eval_fused       <- torch.max + MulticlassConfusionMatrix.update x2   utils.py:88-94 (eval_model)
 *   cvcs_scale_inplace    <- autograd's grad_output * dlogits (only when grad_output != 1)
 *   cvcs_argmax           <- torch.argmax(..., dim=2)           utils.py:158,504  esa.py:56
 *   cvcs_confmat          <- MulticlassConfusionMatrix.update   utils.py:93-94 (index inputs)
 *   cvcs_tile_normalize   <- _get_cropped_data / crop           dataset.py:28-32,136-150
 *                            image.type(torch.float32)          train.py:121
 *                            SegformerMod.preprocessor          nets.py:339-342
 *   cvcs_vote             <- Ensemble majority vote (torch.mode) utils.py:499-507
 *   cvcs_colorize         <- GID15Converter.iconvert            converters.py:23-36
 *   cvcs_stitch           <- tile re-assembly                   inference.py:40-57
 *   cvcs_tile_context     <- _get_context (crop + v2.Resize)    dataset.py:11-16
 *   cvcs_host_*           <- the same calls on HOST buffers (copies inside), i.e. what a
 *                            non-torch caller binds; used for the end-to-end bench number.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No torch / C++ types.
 *   - every pointer named *_dev* or documented "device" is a CUDA device pointer owned by the
 *     caller (e.g. the PyTorch caching allocator).  Kernels never allocate, free or
 *     synchronise the host.  `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).
 *   - confusion matrices and histograms are ACCUMULATED INTO (caller zeroes them), which is
 *     the running-state behaviour of torchmetrics' update().
 *   - return value: 0 on success, negative cvcs_status otherwise; cvcs_last_error() returns
 *     a thread-local human-readable message for the last failing call on this thread.
 */
#ifndef CVCS_B200_H
#define CVCS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden; only the entry points below are exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define CVCS_ABI_VERSION 1

/* status codes */
enum cvcs_status {
    CVCS_OK = 0,
    CVCS_ERR_INVALID_ARG = -1,  /* bad dtype/layout/shape/NULL/alignment                  */
    CVCS_ERR_UNSUPPORTED = -2,  /* shape outside the supported envelope (e.g. C > 1024)    */
    CVCS_ERR_CUDA = -3,         /* a CUDA runtime call failed (message has the CUDA text)  */
    CVCS_ERR_WORKSPACE = -4     /* workspace too small                                     */
};

/* dtype tags */
enum cvcs_dtype {
    CVCS_F32 = 0,
    CVCS_BF16 = 1,
    CVCS_U8 = 2,
    CVCS_I64 = 3,
    CVCS_I32 = 4
};

/* logits layout */
enum cvcs_layout {
    CVCS_NCHW = 0, /* [B, C, H, W] contiguous (conv output)                */
    CVCS_NHWC = 1  /* [B, H, W, C] contiguous (torch channels_last memory) */
};

int cvcs_abi_version(void);
const char* cvcs_last_error(void);
/* number of SMs of the current device (grid sizing is done inside; exposed for the bench) */
int cvcs_sm_count(void);

/* Bytes of device scratch every kernel-launching entry point may be handed (one buffer can
 * be shared by all calls issued on one stream).  The buffer must be zero-filled ONCE after
 * allocation; kernels leave it zeroed again on exit. */
size_t cvcs_workspace_bytes(void);
/* Identity of the CUDA-graph capture `stream` is currently part of (cudaStreamGetCaptureInfo's sequence number, unique
 * per capture in the process), 0 when the stream is not capturing.  Lets a host layer hand ONE zeroed workspace to all
 * the calls of one captured graph (one memset node per graph instead of one per call, and no full dependency between
 * consecutive K1 launches, which would undo their programmatic overlap). */
int cvcs_stream_capture_id(void* stream, unsigned long long* id_out);

/* Sizes: every index on the path is 64-bit.  Entry points accept up to 2^33 pixels per call (B*H*W, or n for the
 * index-map kernels) and any scene size; tests/test_gpu_large.py runs batches of 2.2e9 pixels / 4.4e9 logit elements
 * and a scene of 2.2e9 bytes.  Larger inputs are refused with CVCS_ERR_INVALID_ARG ("too many pixels"). */

/* Process-wide tuning knobs (A/B measurements, path-coverage tests).  Every value selects
 * among CUDA implementations of the same entry point; none changes results. */
enum cvcs_option {
    CVCS_OPT_CE_PATH = 0,    /* cvcs_ce_fused variant: 0 auto, 1 TMA-staged (incl. the wide-C form, C > 21), 2 direct-load, 3 generic */
    CVCS_OPT_TMA_STAGES = 1, /* TMA-staged variant: pipeline depth (0 = as many as fit, max 8)        */
    CVCS_OPT_TMA_WAIT_HINT = 2, /* 0: mbarrier waits pass a long suspend-time hint (default), 1: no hint */
    CVCS_OPT_TMA_VECP = 3,   /* NCHW, C <= 8: pixels per consumer thread (0 = default; f32: 2|4, bf16: 4|8) */
    CVCS_OPT_TMA_CTAS = 4,   /* CTAs per SM the TMA variant sizes its stages for (0 = default 2; 1..4)      */
    CVCS_OPT_TILE_CTAS = 5,  /* K5: CTAs per SM of the persistent grid (0 = default; 1..8)                  */
    CVCS_OPT_RESERVE_SMS = 6, /* K1: SMs left free (e.g. for an NCCL kernel that must run concurrently); 0..32   */
    CVCS_OPT_PDL = 7,        /* K1 (TMA variant) launches with programmatic stream serialization (its prologue overlaps
                                the tail of the previous kernel in the stream; the kernel waits for that kernel before
                                it reads global memory, so results never change): 0 = default (on), 1 = on, 2 = off   */
    CVCS_OPT_L2_HINT = 8,    /* K1 (TMA variant) L2 eviction hints on its bulk copies: 0 = default (when a launch also stages
                                the next batch's labels they are kept in L2 for the next launch, which then reads them from
                                L2 instead of HBM); v in 1..8 = bit mask v - 1 (1 labels as above, 2 logits evict_first,
                                4 gradient stores evict_first); 1 = no hints                                              */
    CVCS_OPT_COUNT = 9
};
int cvcs_set_option(int option, int value);

/* ---- K4: label histogram ------------------------------------------------------------
 * hist_dev: nullable u64[C + 2], accumulated into.
 *   hist[c]     = #pixels with label c, 0 <= c < C   (including c == ignore_index)
 *   hist[C]     = #pixels equal to ignore_index when ignore_index is outside [0, C)
 *   hist[C + 1] = #pixels with an out-of-bounds label (neither of the above)
 * target dtype: CVCS_U8 or CVCS_I64.
 * total_weight_out_dev: nullable f64[2]; when given it is overwritten with
 *   {Σ_i v_i w[y_i], 1/Σ_i v_i w[y_i]} of THIS call's labels (weight_dev nullable = ones),
 *   so that K4 -> K1 chains on the stream without a host round trip. */
int cvcs_label_hist(const void* target_dev, int target_dtype, long long n_pixels, int C,
                    long long ignore_index, unsigned long long* hist_dev, const float* weight_dev,
                    double* total_weight_out_dev, void* workspace_dev, void* stream);

/* int64 labels (what the reference passes: mask.type(torch.long), train.py:122) in ONE pass over them:
 * total_weight_out_dev f64[2] = {Σ_i v_i w[y_i], 1/Σ} as above, and labels_u8_out_dev[i] = the label as a
 * byte (ignore_index -> 255, any other value outside [0, C) -> 254), so that K1 reads 1 B/px instead of
 * 8 B/px: call cvcs_ce_fused with target_dtype CVCS_U8 and ignore_index 255 afterwards.  Needs C <= 254. */
int cvcs_labels_prepare(const long long* target_dev, long long n_pixels, int C, long long ignore_index,
                        const float* weight_dev, double* total_weight_out_dev,
                        unsigned char* labels_u8_out_dev, void* workspace_dev, void* stream);

/* Σ_i v_i w[y_i] from a label histogram (e.g. after an all-reduce across ranks):
 * out_dev[0] = Σw (f64), out_dev[1] = 1/Σw.  weight_dev nullable (all ones).  Class
 * ignore_index (if inside [0, C)) is excluded. */
int cvcs_total_weight(const unsigned long long* hist_dev, const float* weight_dev, int C,
                      long long ignore_index, double* out_dev, void* stream);

/* ---- K1: fused softmax cross-entropy fwd (+bwd) + argmax + confusion matrix ------------
 * logits      [B,C,H,W] (layout NCHW) or [B,H,W,C] (NHWC); CVCS_F32 or CVCS_BF16
 * target      [B,H,W]  CVCS_U8 or CVCS_I64
 * weight_dev  nullable f32[C]
 * inv_total_weight_dev   nullable device f64*: 1/Σ v w[y] (cvcs_total_weight's out[1]); when
 *             NULL the host value `inv_total_weight` is used.  Only read if dlogits != NULL.
 * dlogits     nullable; same dtype/layout as logits. NULL -> forward-only (torch.no_grad()).
 * argmax      nullable; [B,H,W] CVCS_U8 (C <= 256) or CVCS_I64; first maximal index, NaN is
 *             maximal (torch.max semantics).
 * confmat_dev nullable u64[C*C], rows = target, cols = prediction, accumulated into; pixels
 *             with target == ignore_index are dropped.
 * loss_sums_dev  f64[3], OVERWRITTEN: {Σ v w[y] nll, Σ v w[y], #out-of-bounds labels}
 * loss_out_dev   nullable f32[1], overwritten with sums[0]/sums[1] (NaN if an
 *             out-of-bounds label was seen).
 */
int cvcs_ce_fused(const void* logits_dev, int logits_dtype, int layout, const void* target_dev,
                  int target_dtype, const float* weight_dev, long long ignore_index, int B, int C,
                  int H, int W, double inv_total_weight, const double* inv_total_weight_dev,
                  void* dlogits_dev, void* argmax_dev, int argmax_dtype,
                  unsigned long long* confmat_dev, double* loss_sums_dev, float* loss_out_dev,
                  void* workspace_dev, void* stream);

/* ---- K1 with the total weight computed inside the kernel, and exchanged across GPUs -----------------------------
 * nn.CrossEntropyLoss's 'mean' (utils.py:230,238) divides every gradient by Σ_i v_i w[y_i] of the whole batch — with
 * data-parallel ranks, of ALL ranks' batches (SURVEY §8e collective (1)).  cvcs_ce_fused_tw needs no cvcs_label_hist
 * launch and no all-reduce in front of it: the kernel's CTAs first sum the weights over the (u8) labels, meet at a
 * grid-wide barrier, and — when `xchg` spans several ranks — one thread stores this rank's sum into every peer's
 * exchange block over NVLink while every CTA waits for the peers' sums in its own block and adds them in rank order
 * (bit-identical total on every rank).  One launch per step; the collective is a handful of 8-byte peer stores.
 *   xchg                  nullable (single GPU); see below
 *   local_total_weight_dev  nullable f64[1]: this rank's Σ v·w[y] if a cvcs_label_hist / cvcs_labels_prepare launch
 *                         already produced it (e.g. one step ahead on another stream; any label dtype): the kernel then
 *                         skips its label pre-pass and grid barrier and only performs the exchange
 *   total_weight_out_dev  nullable f64[2], overwritten with {Σ (global), 1/Σ}
 *   next_target_dev       nullable u8 labels (16-byte aligned, next_n_pixels of them) of the NEXT batch: this launch also
 *                         sums the weights over them — in its prologue, while its first bulk loads are in flight — and
 *                         writes {Σ, 1/Σ} (this rank's) to next_total_weight_out_dev f64[2] when it ends.  Passing that
 *                         buffer as local_total_weight_dev of the next call pipelines the label pre-pass across
 *                         launches: one launch per step, no pre-pass on the critical path, no second stream.
 * Everything else as cvcs_ce_fused.  Shapes the TMA-staged kernel does not take (int64 labels, odd sizes, C > 21) run
 * as cvcs_label_hist + cvcs_ce_fused internally on one GPU (total_weight_out_dev required) and are refused with
 * CVCS_ERR_UNSUPPORTED across GPUs.  Every rank of an exchange must make the same sequence of calls. */
typedef struct cvcs_xchg cvcs_xchg;
int cvcs_ce_fused_tw(const void* logits_dev, int logits_dtype, int layout, const void* target_dev,
                     int target_dtype, const float* weight_dev, long long ignore_index, int B, int C,
                     int H, int W, cvcs_xchg* xchg, const double* local_total_weight_dev,
                     double* total_weight_out_dev, const void* next_target_dev, long long next_n_pixels,
                     double* next_total_weight_out_dev, void* dlogits_dev,
                     void* argmax_dev, int argmax_dtype, unsigned long long* confmat_dev,
                     double* loss_sums_dev, float* loss_out_dev, void* workspace_dev, void* stream);

/* Exchange handle: a small device block on this rank (cudaMalloc) plus the peers' blocks mapped into this process.
 * One process per GPU: create, pass cvcs_xchg_local_handle's 64 bytes (a cudaIpcMemHandle_t) to the other ranks with
 * whatever transport the host program has (torch.distributed all_gather_object in cvcs_b200.shard), open each peer's.
 * Several GPUs in one process: cvcs_xchg_set_peer with the other handle's cvcs_xchg_local_block (peer access enabled
 * by the caller).  cvcs_xchg_state reads {exchanges completed, time-outs/overruns seen}; cvcs_xchg_poke plays another
 * rank's part of exchange number `seq` (tests). */
int cvcs_xchg_create(cvcs_xchg** out, int world, int rank);
int cvcs_xchg_local_handle(cvcs_xchg* x, unsigned char* handle_out64);
int cvcs_xchg_open_peer(cvcs_xchg* x, int peer_rank, const unsigned char* handle64);
int cvcs_xchg_set_peer(cvcs_xchg* x, int peer_rank, void* block_dev);
void* cvcs_xchg_local_block(cvcs_xchg* x);
int cvcs_xchg_state(cvcs_xchg* x, unsigned long long* seq_out, unsigned long long* errors_out);
int cvcs_xchg_poke(cvcs_xchg* x, int as_rank, unsigned long long seq, double value, void* stream);
/* Pass-end sums without a collective library: buf_dev[0..n) (f64, n <= 2048) is replaced by its sum over all ranks of the
 * exchange, added in rank order (bit-identical on every rank); counts below 2^53 are exact.  One small kernel on
 * `stream`; every rank must call it the same number of times.  Single-rank handles return at once. */
int cvcs_xchg_allreduce_f64(cvcs_xchg* x, double* buf_dev, int n, void* stream);
int cvcs_xchg_destroy(cvcs_xchg* x);

/* ---- K1, metrics mode: argmax + confusion matrix straight from logits, no softmax / loss -----------
 * What utils.eval_model needs per tile (utils.py:88-94: torch.max + two MulticlassConfusionMatrix
 * updates): one read of the logits, the u8 / i64 argmax map (nullable) and the C x C update
 * (nullable, accumulated).  status_dev: nullable u64[1], += #labels outside [0, C) that are not
 * ignore_index (torchmetrics' validate_args, checked lazily by the caller). */
int cvcs_eval_fused(const void* logits_dev, int logits_dtype, int layout, const void* target_dev,
                    int target_dtype, long long ignore_index, int B, int C, int H, int W,
                    void* argmax_dev, int argmax_dtype, unsigned long long* confmat_dev,
                    unsigned long long* status_dev, void* workspace_dev, void* stream);

/* x[i] *= *scale_dev, in place (x: CVCS_F32 or CVCS_BF16).  Returns without touching x when *scale_dev == 1.0f. */
int cvcs_scale_inplace(void* x_dev, int dtype, long long n, const float* scale_dev, void* stream);

/* ---- K2: argmax over the class dimension -------------------------------------------- */
int cvcs_argmax(const void* logits_dev, int logits_dtype, int layout, int B, int C, int H, int W,
                void* out_dev, int out_dtype, void* stream);

/* ---- K3: confusion matrix from index maps --------------------------------------------
 * pred/target dtype: CVCS_U8 or CVCS_I64.  status_dev: nullable u64[1], += #pixels whose
 * (kept) target or prediction is outside [0, C). */
int cvcs_confmat(const void* pred_dev, int pred_dtype, const void* target_dev, int target_dtype,
                 long long n_pixels, int C, long long ignore_index,
                 unsigned long long* confmat_dev, unsigned long long* status_dev,
                 void* workspace_dev, void* stream);

/* ---- K5: tile gather + cast + per-band normalise ---------------------------------------
 * scene_dev   u8 [Cb, H, W] (CHW, as tv_tensors.Image gives)
 * tile_yx_dev i32 [n_tiles, 2] top-left (tly, tlx) of each tile; may be negative / overhang
 *             (zero fill, torchvision crop semantics)
 * tile_slot_dev nullable i32 [n_tiles]: tile i is written to out[tile_slot[i]] (default i), so that
 *             the tiles of several scenes land in one batch in the chunk's shuffled order
 * tile_h/w    output tile size
 * mean_dev/std_dev  nullable f32[Cb]; out = (float(x) - mean) / std, IEEE fp32 as
 *             v2.Normalize does; both NULL -> pure cast
 * out_dev     [n_tiles, Cb, tile_h, tile_w], out_dtype CVCS_U8 (only without mean/std),
 *             CVCS_F32 or CVCS_BF16
 * label_dev   nullable u8 [H, W] label scene; label_out_dev u8 or i64 [n_tiles, tile_h, tile_w]
 * hist_dev    nullable u64[hist_C + 2] label histogram of the emitted label tiles (same
 *             layout as cvcs_label_hist) fused into the same pass. */
int cvcs_tile_normalize(const unsigned char* scene_dev, int Cb, int H, int W,
                        const int* tile_yx_dev, const int* tile_slot_dev, int n_tiles, int tile_h,
                        int tile_w,
                        const float* mean_dev, const float* std_dev, void* out_dev, int out_dtype,
                        const unsigned char* label_dev, void* label_out_dev, int label_out_dtype,
                        unsigned long long* hist_dev, int hist_C, long long hist_ignore_index,
                        void* workspace_dev, void* stream);

/* ---- "next" rows ------------------------------------------------------------ */
/* N4: the context view of patches (dataset.py:11-16 _get_context): for each patch origin (tly, tlx) the 3p x 3p
 * neighbourhood crop(image, tly - p, tlx - p, 3p, 3p) (zeros outside the scene) reduced to p x p exactly as the
 * reference's v2.Resize(p) does on a uint8 tensor (antialiased bilinear, horizontal pass first, Pillow-style fixed
 * point) — bit-identical bytes.  scene u8 [Cb,H,W]; tile_yx i32 [n,2] PATCH origins; out u8 [n (or slots), Cb, p, p];
 * tile_slot as in cvcs_tile_normalize.  p >= 3, n_tiles * Cb <= 65535 per call. */
int cvcs_tile_context(const unsigned char* scene_dev, int Cb, int H, int W, const int* tile_yx_dev,
                      const int* tile_slot_dev, int n_tiles, int p, unsigned char* out_dev, void* stream);
/* N3: per-pixel majority vote over n_maps index maps (u8 or i64, [n_maps, n_pixels]);
 * ties -> smallest class index (torch.mode). */
int cvcs_vote(const void* maps_dev, int dtype, int n_maps, long long n_pixels, int C, void* out_dev,
              int out_dtype, void* stream);
/* N4: index map -> RGB via LUT.  lut_dev f32[C*3]; out f32 [n_pixels, 3]; labels outside
 * [0,C) get (1,1,1) (iconvert initialises with ones). */
int cvcs_colorize(const void* index_dev, int dtype, long long n_pixels, const float* lut_dev, int C,
                  float* out_dev, void* stream);
/* N2: paste tiles [n_tiles, tile_h, tile_w] (u8) into a scene-sized map [H, W] at tile_yx,
 * optionally taking only a centred crop_h x crop_w window of each tile (border correction). */
int cvcs_stitch(const unsigned char* tiles_dev, int n_tiles, int tile_h, int tile_w,
                const int* tile_yx_dev, int crop_h, int crop_w, unsigned char* scene_dev, int H,
                int W, void* stream);

/* ---- host-buffer entry points (what a non-torch caller binds) ------------------------------
 * A context owns device staging buffers, a stream pair and pinned bounce buffers; it is
 * created for a maximum problem size and reused.  All pointers below are HOST pointers
 * (pinned or pageable).  Copies, kernels and the final read-back all happen inside the
 * call; the call returns after the results are in host memory. */
typedef struct cvcs_host_ctx cvcs_host_ctx;

int cvcs_host_ctx_create(cvcs_host_ctx** out, int device, long long max_pixels, int max_C,
                         int logits_dtype);
int cvcs_host_ctx_destroy(cvcs_host_ctx* ctx);

/* Same semantics as cvcs_label_hist + cvcs_total_weight + cvcs_ce_fused on host buffers.
 * dlogits / argmax may be NULL (they then stay on the device, as in a training loop where
 * the model backward consumes them there).  confmat (u64[C*C]) is accumulated into.
 * loss_out f32[1]; loss_sums f64[3] nullable. */
int cvcs_host_ce_fused(cvcs_host_ctx* ctx, const void* logits, int logits_dtype, int layout,
                       const void* target, int target_dtype, const float* weight,
                       long long ignore_index, int B, int C, int H, int W, int want_grad,
                       void* dlogits, void* argmax, int argmax_dtype, unsigned long long* confmat,
                       float* loss_out, double* loss_sums);

/* Device pointers of the last cvcs_host_ce_fused results held by the context (valid until
 * the next call on the context): what = 0 dlogits, 1 argmax, 2 logits staging. */
void* cvcs_host_ctx_device_ptr(cvcs_host_ctx* ctx, int what);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* CVCS_B200_H */
