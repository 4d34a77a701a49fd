// abi.cu — the extern "C" surface declared in include/cvcs_b200.h, plus the host-buffer
// context (device staging, copy/compute streams) behind the cvcs_host_* entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ce_common.cuh"

namespace cvcs {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int g_options[CVCS_OPT_COUNT] = {0};
int get_option(int option) { return (option >= 0 && option < CVCS_OPT_COUNT) ? g_options[option] : 0; }

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// launchers implemented in the kernel translation units
int ce_fused_launch(const void*, int, int, const void*, int, const float*, long long, int, int, int, int, double,
                    const double*, void*, void*, int, unsigned long long*, double*, float*, void*, cudaStream_t,
                    unsigned long long* status = nullptr, int no_loss = 0, const TwRequest* tw = nullptr);
int label_hist_launch(const void*, int, long long, int, long long, unsigned long long*, const float*, double*, void*,
                      cudaStream_t);
int total_weight_launch(const unsigned long long*, const float*, int, long long, double*, cudaStream_t);
int labels_prepare_launch(const long long*, long long, int, long long, const float*, double*, unsigned char*, void*, cudaStream_t);
int scale_launch(void*, int, long long, const float*, cudaStream_t);
int argmax_launch(const void*, int, int, int, int, int, int, void*, int, cudaStream_t);
int confmat_launch(const void*, int, const void*, int, long long, int, long long, unsigned long long*,
                   unsigned long long*, void*, cudaStream_t);
int tile_launch(const unsigned char*, int, int, int, const int*, const int*, int, int, int, const float*, const float*, void*, int,
                const unsigned char*, void*, int, unsigned long long*, int, long long, void*, cudaStream_t);
int stitch_launch(const unsigned char*, int, int, int, const int*, int, int, unsigned char*, int, int, cudaStream_t);
int context_launch(const unsigned char*, int, int, int, const int*, const int*, int, int, unsigned char*, cudaStream_t);
int xchg_allreduce_launch(XchgRegion* const* peers, int world, int rank, double* buf, int n, cudaStream_t stream);
int vote_launch(const void*, int, int, long long, int, void*, int, cudaStream_t);
int colorize_launch(const void*, int, long long, const float*, int, float*, cudaStream_t);

}  // namespace cvcs

using namespace cvcs;

// ---- host-buffer context --------------------------------------------------------------------------
struct cvcs_host_ctx {
    int device;
    long long max_pixels;
    int max_C;
    int esize;
    cudaStream_t s_copy, s_comp, s_back;
    std::vector<cudaEvent_t> ev_in, ev_done;
    cudaEvent_t ev_labels;
    void *d_logits, *d_dlogits, *d_target, *d_argmax, *d_ws;
    float *d_weight, *d_loss;
    unsigned long long *d_conf, *d_hist;
    double *d_tw, *d_sums;
    // pinned result block: sums[3*B] | conf[C*C]
    double* h_sums;
    unsigned long long* h_conf;
    int max_chunks;
};

static void host_ctx_free(cvcs_host_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto e : c->ev_in) cudaEventDestroy(e);
    for (auto e : c->ev_done) cudaEventDestroy(e);
    if (c->ev_labels) cudaEventDestroy(c->ev_labels);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_back) cudaStreamDestroy(c->s_back);
    cudaFree(c->d_logits);
    cudaFree(c->d_dlogits);
    cudaFree(c->d_target);
    cudaFree(c->d_argmax);
    cudaFree(c->d_ws);
    cudaFree(c->d_weight);
    cudaFree(c->d_loss);
    cudaFree(c->d_conf);
    cudaFree(c->d_hist);
    cudaFree(c->d_tw);
    cudaFree(c->d_sums);
    if (c->h_sums) cudaFreeHost(c->h_sums);
    if (c->h_conf) cudaFreeHost(c->h_conf);
    delete c;
}

extern "C" {

int cvcs_abi_version(void) { return CVCS_ABI_VERSION; }
const char* cvcs_last_error(void) { return g_err; }
int cvcs_sm_count(void) { return num_sms(); }
size_t cvcs_workspace_bytes(void) { return kWorkspaceBytes; }

int cvcs_stream_capture_id(void* stream, unsigned long long* id_out) {
    CVCS_REQUIRE(id_out, "cvcs_stream_capture_id: NULL id_out");
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0ull;
    CVCS_CUDA_OK(cudaStreamGetCaptureInfo(reinterpret_cast<cudaStream_t>(stream), &status, &id));
    *id_out = status == cudaStreamCaptureStatusActive ? id : 0ull;
    return CVCS_OK;
}

int cvcs_set_option(int option, int value) {
    CVCS_REQUIRE(option >= 0 && option < CVCS_OPT_COUNT, "cvcs_set_option: unknown option %d", option);
    g_options[option] = value;
    return CVCS_OK;
}

int cvcs_label_hist(const void* target_dev, int target_dtype, long long n_pixels, int C, long long ignore_index,
                    unsigned long long* hist_dev, const float* weight_dev, double* total_weight_out_dev,
                    void* workspace_dev, void* stream) {
    return label_hist_launch(target_dev, target_dtype, n_pixels, C, ignore_index, hist_dev, weight_dev,
                             total_weight_out_dev, workspace_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_labels_prepare(const long long* target_dev, long long n_pixels, int C, long long ignore_index,
                        const float* weight_dev, double* total_weight_out_dev, unsigned char* labels_u8_out_dev,
                        void* workspace_dev, void* stream) {
    return labels_prepare_launch(target_dev, n_pixels, C, ignore_index, weight_dev, total_weight_out_dev, labels_u8_out_dev,
                                 workspace_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_total_weight(const unsigned long long* hist_dev, const float* weight_dev, int C, long long ignore_index,
                      double* out_dev, void* stream) {
    return total_weight_launch(hist_dev, weight_dev, C, ignore_index, out_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_ce_fused(const void* logits_dev, int logits_dtype, int layout, const void* target_dev, int target_dtype,
                  const float* weight_dev, long long ignore_index, int B, int C, int H, int W, double inv_total_weight,
                  const double* inv_total_weight_dev, void* dlogits_dev, void* argmax_dev, int argmax_dtype,
                  unsigned long long* confmat_dev, double* loss_sums_dev, float* loss_out_dev, void* workspace_dev,
                  void* stream) {
    return ce_fused_launch(logits_dev, logits_dtype, layout, target_dev, target_dtype, weight_dev, ignore_index, B, C, H,
                           W, inv_total_weight, inv_total_weight_dev, dlogits_dev, argmax_dev, argmax_dtype, confmat_dev,
                           loss_sums_dev, loss_out_dev, workspace_dev, static_cast<cudaStream_t>(stream));
}

// ---- Σw exchange handle: the local block + the peers' blocks mapped through CUDA IPC ------------------------------
struct cvcs_xchg {
    int world, rank, device;
    XchgBlock* peer[kXMaxRanks];
    bool opened[kXMaxRanks];      // mapped with cudaIpcOpenMemHandle (to be closed), as opposed to set directly
};

int cvcs_xchg_create(cvcs_xchg** out, int world, int rank) {
    CVCS_REQUIRE(out && world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, "cvcs_xchg_create: world %d rank %d (max %d ranks)",
                 world, rank, kXMaxRanks);
    *out = nullptr;
    cvcs_xchg* x = new cvcs_xchg();
    x->world = world;
    x->rank = rank;
    for (int q = 0; q < kXMaxRanks; ++q) { x->peer[q] = nullptr; x->opened[q] = false; }
    cudaError_t e = cudaGetDevice(&x->device);
    void* blk = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&blk, sizeof(XchgRegion));     // its own allocation: an IPC handle maps whole allocations
    if (e == cudaSuccess) e = cudaMemset(blk, 0, sizeof(XchgRegion));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (blk) cudaFree(blk);
        delete x;
        (void)cudaGetLastError();
        return set_error(CVCS_ERR_CUDA, "cvcs_xchg_create: %s", cudaGetErrorString(e));
    }
    x->peer[rank] = static_cast<XchgBlock*>(blk);
    *out = x;
    return CVCS_OK;
}

int cvcs_xchg_local_handle(cvcs_xchg* x, unsigned char* handle_out64) {
    CVCS_REQUIRE(x && handle_out64, "cvcs_xchg_local_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CVCS_CUDA_OK(cudaIpcGetMemHandle(&h, x->peer[x->rank]));
    memcpy(handle_out64, &h, 64);
    return CVCS_OK;
}

int cvcs_xchg_open_peer(cvcs_xchg* x, int peer_rank, const unsigned char* handle64) {
    CVCS_REQUIRE(x && handle64 && peer_rank >= 0 && peer_rank < x->world && peer_rank != x->rank, "cvcs_xchg_open_peer: bad peer rank %d", peer_rank);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CVCS_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer_rank] = static_cast<XchgBlock*>(ptr);
    x->opened[peer_rank] = true;
    return CVCS_OK;
}

int cvcs_xchg_set_peer(cvcs_xchg* x, int peer_rank, void* block_dev) {
    CVCS_REQUIRE(x && block_dev && peer_rank >= 0 && peer_rank < x->world && peer_rank != x->rank, "cvcs_xchg_set_peer: bad peer rank %d", peer_rank);
    x->peer[peer_rank] = static_cast<XchgBlock*>(block_dev);
    x->opened[peer_rank] = false;
    return CVCS_OK;
}

void* cvcs_xchg_local_block(cvcs_xchg* x) { return x ? x->peer[x->rank] : nullptr; }

int cvcs_xchg_state(cvcs_xchg* x, unsigned long long* seq_out, unsigned long long* errors_out) {
    CVCS_REQUIRE(x, "cvcs_xchg_state: NULL handle");
    XchgBlock* b = x->peer[x->rank];
    if (seq_out) CVCS_CUDA_OK(cudaMemcpy(seq_out, &b->seq, 8, cudaMemcpyDeviceToHost));
    if (errors_out) CVCS_CUDA_OK(cudaMemcpy(errors_out, &b->errors, 8, cudaMemcpyDeviceToHost));
    return CVCS_OK;
}

int cvcs_xchg_poke(cvcs_xchg* x, int as_rank, unsigned long long seq, double value, void* stream) {
    CVCS_REQUIRE(x && as_rank >= 0 && as_rank < x->world && seq > 0, "cvcs_xchg_poke: bad argument");
    XchgBlock* b = x->peer[x->rank];
    const int slot = static_cast<int>(seq % kXDepth);
    const unsigned int tag = static_cast<unsigned int>(seq);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CVCS_CUDA_OK(cudaMemcpyAsync(&b->slots[slot][as_rank], &value, 8, cudaMemcpyHostToDevice, st));
    CVCS_CUDA_OK(cudaMemcpyAsync(&b->flags[slot][as_rank], &tag, 4, cudaMemcpyHostToDevice, st));
    CVCS_CUDA_OK(cudaStreamSynchronize(st));     // the host values above live on this stack frame
    return CVCS_OK;
}

int cvcs_xchg_allreduce_f64(cvcs_xchg* x, double* buf_dev, int n, void* stream) {
    CVCS_REQUIRE(x && buf_dev && n >= 1 && n <= kXWideN, "cvcs_xchg_allreduce_f64: 1 <= n <= %d doubles (got %d)", kXWideN, n);
    if (x->world == 1) return CVCS_OK;
    XchgRegion* peers[kXMaxRanks];
    for (int q = 0; q < x->world; ++q) {
        CVCS_REQUIRE(x->peer[q], "cvcs_xchg_allreduce_f64: the exchange block of rank %d is not mapped", q);
        peers[q] = reinterpret_cast<XchgRegion*>(x->peer[q]);
    }
    return xchg_allreduce_launch(peers, x->world, x->rank, buf_dev, n, static_cast<cudaStream_t>(stream));
}

int cvcs_xchg_destroy(cvcs_xchg* x) {
    if (!x) return CVCS_OK;
    for (int q = 0; q < x->world; ++q)
        if (q != x->rank && x->opened[q] && x->peer[q]) cudaIpcCloseMemHandle(x->peer[q]);
    if (x->peer[x->rank]) cudaFree(x->peer[x->rank]);
    (void)cudaGetLastError();
    delete x;
    return CVCS_OK;
}

int cvcs_ce_fused_tw(const void* logits_dev, int logits_dtype, int layout, const void* target_dev, int target_dtype,
                     const float* weight_dev, long long ignore_index, int B, int C, int H, int W, cvcs_xchg* xchg,
                     const double* local_total_weight_dev, double* total_weight_out_dev, const void* next_target_dev,
                     long long next_n_pixels, double* next_total_weight_out_dev, void* dlogits_dev, void* argmax_dev, int argmax_dtype,
                     unsigned long long* confmat_dev, double* loss_sums_dev, float* loss_out_dev, void* workspace_dev,
                     void* stream) {
    TwRequest tw{};
    tw.tw_local = local_total_weight_dev;
    tw.next_target = next_target_dev;
    tw.next_n = next_n_pixels;
    tw.next_tw_out = next_total_weight_out_dev;
    tw.tw_out = total_weight_out_dev;
    tw.world = 1;
    tw.rank = 0;
    if (xchg && xchg->world > 1) {
        tw.world = xchg->world;
        tw.rank = xchg->rank;
        for (int q = 0; q < xchg->world; ++q) {
            CVCS_REQUIRE(xchg->peer[q], "cvcs_ce_fused_tw: the exchange block of rank %d is not mapped (cvcs_xchg_open_peer)", q);
            tw.peer[q] = xchg->peer[q];
        }
    }
    return ce_fused_launch(logits_dev, logits_dtype, layout, target_dev, target_dtype, weight_dev, ignore_index, B, C, H,
                           W, 0.0, nullptr, dlogits_dev, argmax_dev, argmax_dtype, confmat_dev, loss_sums_dev, loss_out_dev,
                           workspace_dev, static_cast<cudaStream_t>(stream), nullptr, 0, &tw);
}

int cvcs_eval_fused(const void* logits_dev, int logits_dtype, int layout, const void* target_dev, int target_dtype,
                    long long ignore_index, int B, int C, int H, int W, void* argmax_dev, int argmax_dtype,
                    unsigned long long* confmat_dev, unsigned long long* status_dev, void* workspace_dev, void* stream) {
    CVCS_REQUIRE(argmax_dev || confmat_dev, "cvcs_eval_fused: nothing to compute (argmax and confmat both NULL)");
    return ce_fused_launch(logits_dev, logits_dtype, layout, target_dev, target_dtype, nullptr, ignore_index, B, C, H, W,
                           0.0, nullptr, nullptr, argmax_dev, argmax_dtype, confmat_dev, nullptr, nullptr, workspace_dev,
                           static_cast<cudaStream_t>(stream), status_dev, 1);
}

int cvcs_scale_inplace(void* x_dev, int dtype, long long n, const float* scale_dev, void* stream) {
    return scale_launch(x_dev, dtype, n, scale_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_argmax(const void* logits_dev, int logits_dtype, int layout, int B, int C, int H, int W, void* out_dev,
                int out_dtype, void* stream) {
    return argmax_launch(logits_dev, logits_dtype, layout, B, C, H, W, out_dev, out_dtype,
                         static_cast<cudaStream_t>(stream));
}

int cvcs_confmat(const void* pred_dev, int pred_dtype, const void* target_dev, int target_dtype, long long n_pixels,
                 int C, long long ignore_index, unsigned long long* confmat_dev, unsigned long long* status_dev,
                 void* workspace_dev, void* stream) {
    return confmat_launch(pred_dev, pred_dtype, target_dev, target_dtype, n_pixels, C, ignore_index, confmat_dev,
                          status_dev, workspace_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_tile_normalize(const unsigned char* scene_dev, int Cb, int H, int W, const int* tile_yx_dev,
                        const int* tile_slot_dev, int n_tiles,
                        int tile_h, int tile_w, const float* mean_dev, const float* std_dev, void* out_dev,
                        int out_dtype, const unsigned char* label_dev, void* label_out_dev, int label_out_dtype,
                        unsigned long long* hist_dev, int hist_C, long long hist_ignore_index, void* workspace_dev,
                        void* stream) {
    return tile_launch(scene_dev, Cb, H, W, tile_yx_dev, tile_slot_dev, n_tiles, tile_h, tile_w, mean_dev, std_dev, out_dev, out_dtype,
                       label_dev, label_out_dev, label_out_dtype, hist_dev, hist_C, hist_ignore_index, workspace_dev,
                       static_cast<cudaStream_t>(stream));
}

int cvcs_tile_context(const unsigned char* scene_dev, int Cb, int H, int W, const int* tile_yx_dev, const int* tile_slot_dev,
                      int n_tiles, int p, unsigned char* out_dev, void* stream) {
    return context_launch(scene_dev, Cb, H, W, tile_yx_dev, tile_slot_dev, n_tiles, p, out_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_vote(const void* maps_dev, int dtype, int n_maps, long long n_pixels, int C, void* out_dev, int out_dtype,
              void* stream) {
    return vote_launch(maps_dev, dtype, n_maps, n_pixels, C, out_dev, out_dtype, static_cast<cudaStream_t>(stream));
}

int cvcs_colorize(const void* index_dev, int dtype, long long n_pixels, const float* lut_dev, int C, float* out_dev,
                  void* stream) {
    return colorize_launch(index_dev, dtype, n_pixels, lut_dev, C, out_dev, static_cast<cudaStream_t>(stream));
}

int cvcs_stitch(const unsigned char* tiles_dev, int n_tiles, int tile_h, int tile_w, const int* tile_yx_dev, int crop_h,
                int crop_w, unsigned char* scene_dev, int H, int W, void* stream) {
    return stitch_launch(tiles_dev, n_tiles, tile_h, tile_w, tile_yx_dev, crop_h, crop_w, scene_dev, H, W,
                         static_cast<cudaStream_t>(stream));
}

// ---- host-buffer entry points ---------------------------------------------------------------
int cvcs_host_ctx_create(cvcs_host_ctx** out, int device, long long max_pixels, int max_C, int logits_dtype) {
    CVCS_REQUIRE(out && max_pixels > 0 && max_C >= 1 && max_C <= 1024, "cvcs_host_ctx_create: bad argument");
    CVCS_REQUIRE(logits_dtype == CVCS_F32 || logits_dtype == CVCS_BF16, "cvcs_host_ctx_create: logits dtype tag %d", logits_dtype);
    *out = nullptr;
    CVCS_CUDA_OK(cudaSetDevice(device));
    cvcs_host_ctx* c = new cvcs_host_ctx();  // value-initialised: every pointer starts NULL
    c->s_copy = c->s_comp = c->s_back = nullptr;
    c->ev_labels = nullptr;
    c->d_logits = c->d_dlogits = c->d_target = c->d_argmax = c->d_ws = nullptr;
    c->d_weight = c->d_loss = nullptr;
    c->d_conf = c->d_hist = nullptr;
    c->d_tw = c->d_sums = nullptr;
    c->h_sums = nullptr;
    c->h_conf = nullptr;
    c->device = device;
    c->max_pixels = max_pixels;
    c->max_C = max_C;
    c->esize = logits_dtype == CVCS_F32 ? 4 : 2;
    c->max_chunks = 64;
    const size_t lbytes = static_cast<size_t>(max_pixels) * max_C * c->esize;
#define CTX_OK(expr)                                                                                        \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            host_ctx_free(c);                                                                               \
            return set_error(CVCS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));                \
        }                                                                                                   \
    } while (0)
    CTX_OK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CTX_OK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    CTX_OK(cudaStreamCreateWithFlags(&c->s_back, cudaStreamNonBlocking));
    CTX_OK(cudaEventCreateWithFlags(&c->ev_labels, cudaEventDisableTiming));
    c->ev_in.resize(c->max_chunks);
    c->ev_done.resize(c->max_chunks);
    for (int i = 0; i < c->max_chunks; ++i) {
        c->ev_in[i] = c->ev_done[i] = nullptr;
    }
    for (int i = 0; i < c->max_chunks; ++i) {
        CTX_OK(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        CTX_OK(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    CTX_OK(cudaMalloc(&c->d_logits, lbytes));
    CTX_OK(cudaMalloc(&c->d_dlogits, lbytes));
    CTX_OK(cudaMalloc(&c->d_target, static_cast<size_t>(max_pixels) * 8));
    CTX_OK(cudaMalloc(&c->d_argmax, static_cast<size_t>(max_pixels) * 8));
    CTX_OK(cudaMalloc(&c->d_ws, kWorkspaceBytes));
    CTX_OK(cudaMemset(c->d_ws, 0, kWorkspaceBytes));
    CTX_OK(cudaMalloc(&c->d_weight, sizeof(float) * max_C));
    CTX_OK(cudaMalloc(&c->d_loss, sizeof(float) * c->max_chunks));
    CTX_OK(cudaMalloc(&c->d_conf, sizeof(unsigned long long) * max_C * max_C));
    CTX_OK(cudaMalloc(&c->d_hist, sizeof(unsigned long long) * (max_C + 2)));
    CTX_OK(cudaMalloc(&c->d_tw, sizeof(double) * 2));
    CTX_OK(cudaMalloc(&c->d_sums, sizeof(double) * 3 * c->max_chunks));
    CTX_OK(cudaMallocHost(&c->h_sums, sizeof(double) * 3 * c->max_chunks));
    CTX_OK(cudaMallocHost(&c->h_conf, sizeof(unsigned long long) * max_C * max_C));
    CTX_OK(cudaDeviceSynchronize());
#undef CTX_OK
    *out = c;
    return CVCS_OK;
}

int cvcs_host_ctx_destroy(cvcs_host_ctx* ctx) {
    host_ctx_free(ctx);
    return CVCS_OK;
}

void* cvcs_host_ctx_device_ptr(cvcs_host_ctx* ctx, int what) {
    if (!ctx) return nullptr;
    switch (what) {
        case 0: return ctx->d_dlogits;
        case 1: return ctx->d_argmax;
        case 2: return ctx->d_logits;
        default: return nullptr;
    }
}

int cvcs_host_ce_fused(cvcs_host_ctx* c, const void* logits, int logits_dtype, int layout, const void* target,
                       int target_dtype, const float* weight, long long ignore_index, int B, int C, int H, int W,
                       int want_grad, void* dlogits, void* argmax, int argmax_dtype, unsigned long long* confmat,
                       float* loss_out, double* loss_sums) {
    CVCS_REQUIRE(c && logits && target && loss_out, "cvcs_host_ce_fused: NULL ctx/logits/target/loss_out");
    CVCS_REQUIRE(logits_dtype == CVCS_F32 || logits_dtype == CVCS_BF16, "cvcs_host_ce_fused: logits dtype tag %d", logits_dtype);
    CVCS_REQUIRE((logits_dtype == CVCS_F32 ? 4 : 2) == c->esize, "cvcs_host_ce_fused: context was created for another logits dtype");
    CVCS_REQUIRE(target_dtype == CVCS_U8 || target_dtype == CVCS_I64, "cvcs_host_ce_fused: target dtype tag %d", target_dtype);
    CVCS_REQUIRE(argmax_dtype == CVCS_U8 || argmax_dtype == CVCS_I64, "cvcs_host_ce_fused: argmax dtype tag %d", argmax_dtype);
    CVCS_REQUIRE(B > 0 && C >= 1 && H > 0 && W > 0, "cvcs_host_ce_fused: bad shape");
    const long long hw = static_cast<long long>(H) * W, n = hw * B;
    CVCS_REQUIRE(n <= c->max_pixels && C <= c->max_C, "cvcs_host_ce_fused: problem larger than the context (pixels %lld > %lld or C %d > %d)", n, c->max_pixels, C, c->max_C);
    CVCS_REQUIRE(!(dlogits && !want_grad), "cvcs_host_ce_fused: dlogits buffer given but want_grad == 0");
    CVCS_CUDA_OK(cudaSetDevice(c->device));

    const size_t tsize = target_dtype == CVCS_I64 ? 8 : 1;
    const size_t asize = argmax_dtype == CVCS_I64 ? 8 : 1;
    // chunk = a run of whole images, so that every chunk stays 16-byte aligned in all buffers
    int imgs_per_chunk = 1;
    while ((B + imgs_per_chunk - 1) / imgs_per_chunk > c->max_chunks) ++imgs_per_chunk;
    if ((hw * c->esize) % 16 != 0 || (hw * static_cast<long long>(tsize)) % 16 != 0 || (hw * static_cast<long long>(asize)) % 16 != 0)
        imgs_per_chunk = B;  // odd plane sizes: one chunk keeps base-pointer alignment
    const int n_chunks = (B + imgs_per_chunk - 1) / imgs_per_chunk;

    // labels (+ weights) first: Σw must be known before the first dlogit is written
    CVCS_CUDA_OK(cudaMemcpyAsync(c->d_target, target, static_cast<size_t>(n) * tsize, cudaMemcpyHostToDevice, c->s_comp));
    if (weight) CVCS_CUDA_OK(cudaMemcpyAsync(c->d_weight, weight, sizeof(float) * C, cudaMemcpyHostToDevice, c->s_comp));
    CVCS_CUDA_OK(cudaMemsetAsync(c->d_conf, 0, sizeof(unsigned long long) * C * C, c->s_comp));
    if (want_grad) {
        int rc = label_hist_launch(c->d_target, target_dtype, n, C, ignore_index, nullptr, weight ? c->d_weight : nullptr,
                                   c->d_tw, c->d_ws, c->s_comp);
        if (rc) return rc;
    }
    const size_t img_lbytes = static_cast<size_t>(hw) * C * c->esize;
    for (int k = 0; k < n_chunks; ++k) {
        const int b0 = k * imgs_per_chunk;
        const int nb = (b0 + imgs_per_chunk <= B) ? imgs_per_chunk : (B - b0);
        const size_t loff = img_lbytes * b0;
        CVCS_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(c->d_logits) + loff, static_cast<const char*>(logits) + loff,
                                     img_lbytes * nb, cudaMemcpyHostToDevice, c->s_copy));
        CVCS_CUDA_OK(cudaEventRecord(c->ev_in[k], c->s_copy));
        CVCS_CUDA_OK(cudaStreamWaitEvent(c->s_comp, c->ev_in[k], 0));
        int rc = ce_fused_launch(static_cast<char*>(c->d_logits) + loff, logits_dtype, layout,
                                 static_cast<char*>(c->d_target) + static_cast<size_t>(hw) * b0 * tsize, target_dtype,
                                 weight ? c->d_weight : nullptr, ignore_index, nb, C, H, W, 0.0,
                                 want_grad ? c->d_tw + 1 : nullptr,  // {Σw, 1/Σw}: K1 wants the reciprocal
                                 want_grad ? static_cast<char*>(c->d_dlogits) + loff : nullptr,
                                 static_cast<char*>(c->d_argmax) + static_cast<size_t>(hw) * b0 * asize, argmax_dtype,
                                 c->d_conf, c->d_sums + 3 * k, nullptr, c->d_ws, c->s_comp);
        if (rc) return rc;
        if (dlogits || argmax) {
            CVCS_CUDA_OK(cudaEventRecord(c->ev_done[k], c->s_comp));
            CVCS_CUDA_OK(cudaStreamWaitEvent(c->s_back, c->ev_done[k], 0));
            if (dlogits)
                CVCS_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(dlogits) + loff, static_cast<char*>(c->d_dlogits) + loff,
                                             img_lbytes * nb, cudaMemcpyDeviceToHost, c->s_back));
            if (argmax)
                CVCS_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(argmax) + static_cast<size_t>(hw) * b0 * asize,
                                             static_cast<char*>(c->d_argmax) + static_cast<size_t>(hw) * b0 * asize,
                                             static_cast<size_t>(hw) * nb * asize, cudaMemcpyDeviceToHost, c->s_back));
        }
    }
    CVCS_CUDA_OK(cudaMemcpyAsync(c->h_sums, c->d_sums, sizeof(double) * 3 * n_chunks, cudaMemcpyDeviceToHost, c->s_comp));
    CVCS_CUDA_OK(cudaMemcpyAsync(c->h_conf, c->d_conf, sizeof(unsigned long long) * C * C, cudaMemcpyDeviceToHost, c->s_comp));
    CVCS_CUDA_OK(cudaStreamSynchronize(c->s_comp));
    if (dlogits || argmax) CVCS_CUDA_OK(cudaStreamSynchronize(c->s_back));

    double a = 0.0, b = 0.0, bad = 0.0;
    for (int k = 0; k < n_chunks; ++k) {  // fixed order
        a += c->h_sums[3 * k];
        b += c->h_sums[3 * k + 1];
        bad += c->h_sums[3 * k + 2];
    }
    float l = static_cast<float>(a / b);
    if (bad > 0) {
        const unsigned int qnan = 0x7fc00000u;
        memcpy(&l, &qnan, 4);
    }
    *loss_out = l;
    if (loss_sums) {
        loss_sums[0] = a;
        loss_sums[1] = b;
        loss_sums[2] = bad;
    }
    if (confmat)
        for (int i = 0; i < C * C; ++i) confmat[i] += c->h_conf[i];
    return CVCS_OK;
}

}  // extern "C"
