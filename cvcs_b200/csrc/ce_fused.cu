// ce_fused.cu — K1 entry: argument validation and variant selection.
//   tma     : warp-specialised, cp.async.bulk-staged (ce_tma_impl.cuh) — primary path
//   direct  : register-resident 128-bit LDG/STG, NCHW (ce_direct.cu)
//   wide    : C > 21, NCHW: bulk-copy staged, a thread walks its pixel's classes in shared memory (ce_wide.cu)
//   generic : any C <= 1024 / any layout, one pixel per thread (ce_direct.cu)
// cvcs_set_option(CVCS_OPT_CE_PATH, ...) forces a variant for A/B measurements and path-coverage
// tests (all are CUDA; there is no CPU path).
#include <stdlib.h>
#include <string.h>

#include "ce_common.cuh"

namespace cvcs {

int ce_tma_launch_f32(const CeParams& p, int layout, cudaStream_t stream, bool* handled);
int ce_tma_launch_bf16(const CeParams& p, int layout, cudaStream_t stream, bool* handled);
int ce_wide_launch(const CeParams& p, int logits_dtype, cudaStream_t stream, bool* handled);   // ce_wide.cu: C > kMaxRegC, NCHW

// C beyond the register-resident range: the shared-memory class walk (same bulk-copy pipeline), when the layout and
// alignment allow bulk copies; *handled = false leaves the call to the generic kernel
static int try_wide(const CeParams& p, int logits_dtype, int layout, bool ptr16, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int forced = get_option(CVCS_OPT_CE_PATH);
    if (forced == 2 || forced == 3 || p.C <= kMaxRegC || layout != CVCS_NCHW || !ptr16 || p.hw % 16 != 0 || p.tw_mode != 0)
        return CVCS_OK;
    return ce_wide_launch(p, logits_dtype, stream, handled);
}

int label_hist_launch(const void*, int, long long, int, long long, unsigned long long*, const float*, double*, void*,
                      cudaStream_t);
int reciprocal_launch(const double* in, double* out2, cudaStream_t stream);   // out2 = {in[0], 1 / in[0]}

// tw_mode 1 on a shape the TMA-staged kernel does not take (odd sizes, int64 labels, C > 21, a forced variant): the
// same result from two launches — K4 (lean: Σ v·w[y] only) into tw_out, then K1 reading 1/Σ from there.  Single GPU
// only: the cross-GPU exchange lives in the TMA kernel's prologue.
static int tw_fallback(CeParams p, int logits_dtype, int layout, int target_dtype, cudaStream_t stream) {
    if (p.xworld > 1)
        return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_ce_fused_tw: the cross-GPU exchange needs a shape the TMA-staged kernel takes "
                         "(u8 labels, 16-byte aligned tensors, H*W %% 16 == 0, 2 <= C <= %d)", kMaxRegC);
    int rc = CVCS_OK;
    if (p.tw_mode == 2) {
        // Σ already on the device: only its reciprocal is missing
        CVCS_REQUIRE(p.tw_out, "cvcs_ce_fused_tw: total_weight_out_dev is needed for this shape (fallback path)");
        rc = reciprocal_launch(p.tw_local_dev, p.tw_out, stream);
    } else {
        CVCS_REQUIRE(p.tw_out, "cvcs_ce_fused_tw: total_weight_out_dev is needed for this shape (two-launch fallback)");
        rc = label_hist_launch(p.target, target_dtype, p.n_pixels, p.C, p.ignore_index, nullptr, p.weight, p.tw_out, p.ws, stream);
    }
    if (rc) return rc;
    if (p.next_target) {   // the next batch's sum as a launch of its own on this path
        rc = label_hist_launch(p.next_target, CVCS_U8, p.next_n, p.C, p.ignore_index, nullptr, p.weight, p.next_tw_out, p.ws, stream);
        if (rc) return rc;
        p.next_target = nullptr;
        p.next_tw_out = nullptr;
    }
    p.tw_mode = 0;
    p.inv_tw_dev = p.tw_out + 1;
    const int forced = get_option(CVCS_OPT_CE_PATH);
    const int esize = logits_dtype == CVCS_F32 ? 4 : 2;
    (void)esize;
    auto aligned = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const bool ptr16 = aligned(p.logits, 16) && aligned(p.dlogits, 16) && aligned(p.target, 16) && aligned(p.argmax, 16);
    bool handled = false;
    if (forced != 3 && p.C >= 2 && p.C <= kMaxRegC && ptr16) {
        const bool tma_ok = layout == CVCS_NCHW ? (p.hw % 16 == 0) : (p.n_pixels % 16 == 0);
        if (tma_ok && forced != 2) {
            rc = logits_dtype == CVCS_F32 ? ce_tma_launch_f32(p, layout, stream, &handled) : ce_tma_launch_bf16(p, layout, stream, &handled);
            if (handled) return rc;
        }
        if (layout == CVCS_NCHW && forced != 1) {
            const int vec = logits_dtype == CVCS_F32 ? 4 : (p.C <= 12 ? 8 : 4);
            if (p.hw % vec == 0) {
                p.n_items = p.n_pixels / vec;
                p.items_per_image = static_cast<unsigned int>(p.hw / vec);
                rc = ce_direct_launch(p, logits_dtype, vec, stream, &handled);
                if (handled) return rc;
            }
        }
    }
    rc = try_wide(p, logits_dtype, layout, ptr16, stream, &handled);
    if (handled || rc) return rc;
    return ce_generic_launch(p, logits_dtype, layout, stream);
}

int ce_fused_launch(const void* logits, int logits_dtype, int layout, const void* target, int target_dtype,
                    const float* weight, long long ignore_index, int B, int C, int H, int W,
                    double inv_total_weight, const double* inv_total_weight_dev, void* dlogits, void* argmax,
                    int argmax_dtype, unsigned long long* confmat, double* loss_sums, float* loss_out,
                    void* workspace, cudaStream_t stream, unsigned long long* status, int no_loss, const TwRequest* tw) {
    CVCS_REQUIRE(logits && target && (loss_sums || no_loss) && workspace, "cvcs_ce_fused: NULL logits/target/loss_sums/workspace");
    CVCS_REQUIRE(!(no_loss && dlogits), "cvcs_eval_fused: metrics mode has no gradients");
    CVCS_REQUIRE(logits_dtype == CVCS_F32 || logits_dtype == CVCS_BF16, "cvcs_ce_fused: logits dtype tag %d (want f32/bf16)", logits_dtype);
    CVCS_REQUIRE(layout == CVCS_NCHW || layout == CVCS_NHWC, "cvcs_ce_fused: layout %d", layout);
    CVCS_REQUIRE(target_dtype == CVCS_U8 || target_dtype == CVCS_I64,
                 "cvcs_ce_fused: expected target dtype Long (i64) or Byte (u8), got tag %d", target_dtype);
    CVCS_REQUIRE(!argmax || argmax_dtype == CVCS_U8 || argmax_dtype == CVCS_I64, "cvcs_ce_fused: argmax dtype tag %d", argmax_dtype);
    CVCS_REQUIRE(B > 0 && C >= 1 && H > 0 && W > 0, "cvcs_ce_fused: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    if (C > 1024) return set_error(CVCS_ERR_UNSUPPORTED, "cvcs_ce_fused: C=%d > 1024", C);
    CVCS_REQUIRE(!(argmax && argmax_dtype == CVCS_U8 && C > 256), "cvcs_ce_fused: u8 argmax needs C <= 256");

    const long long hw = static_cast<long long>(H) * W;
    const long long n_pixels = hw * B;
    CVCS_REQUIRE(n_pixels < (1ll << 33), "cvcs_ce_fused: too many pixels (%lld)", n_pixels);

    CeParams p{};
    p.logits = logits;
    p.target = target;
    p.weight = weight;
    p.dlogits = dlogits;
    p.argmax = argmax;
    p.confmat = confmat;
    p.inv_tw_dev = inv_total_weight_dev;
    p.inv_tw = inv_total_weight;
    p.loss_sums = loss_sums;
    p.loss_out = loss_out;
    p.ws = reinterpret_cast<Workspace*>(workspace);
    p.ignore_index = ignore_index;
    p.hw = hw;
    p.n_pixels = n_pixels;
    p.C = C;
    p.target_i64 = target_dtype == CVCS_I64;
    p.argmax_i64 = argmax_dtype == CVCS_I64;
    p.status = status;
    p.no_loss = no_loss;
    p.xworld = 1;
    if (tw && !dlogits && tw->next_target) {
        // forward-only call inside a pipelined sequence: the next batch's sum still has to appear — as a launch of its own
        CVCS_REQUIRE(tw->next_tw_out, "cvcs_ce_fused_tw: next_target needs next_total_weight_out_dev");
        int rc = label_hist_launch(tw->next_target, CVCS_U8, tw->next_n, C, ignore_index, nullptr, weight, tw->next_tw_out, workspace, stream);
        if (rc) return rc;
    }
    if (tw && dlogits) {
        // total weight computed by K1 itself (label pre-pass + grid barrier [+ exchange]); forward-only calls do not need it
        CVCS_REQUIRE(tw->world >= 1 && tw->world <= kXMaxRanks && tw->rank >= 0 && tw->rank < tw->world,
                     "cvcs_ce_fused_tw: bad exchange geometry (world %d, rank %d)", tw->world, tw->rank);
        if (tw->next_target) {
            CVCS_REQUIRE(tw->next_tw_out && tw->next_n >= 0 && (reinterpret_cast<uintptr_t>(tw->next_target) % 16) == 0,
                         "cvcs_ce_fused_tw: next_target needs next_total_weight_out_dev and 16-byte alignment");
            p.next_target = tw->next_target;
            p.next_n = tw->next_n;
            p.next_tw_out = tw->next_tw_out;
        }
        p.tw_mode = tw->tw_local ? 2 : 1;
        p.tw_local_dev = tw->tw_local;
        p.tw_out = tw->tw_out;
        p.xworld = tw->world;
        p.xrank = tw->rank;
        for (int q = 0; q < tw->world; ++q) p.xpeer[q] = tw->peer[q];
        p.inv_tw_dev = nullptr;
        p.inv_tw = 0.0;
    }

    const int forced = get_option(CVCS_OPT_CE_PATH);  // 0 auto, 1 tma, 2 direct, 3 generic
    const int esize = logits_dtype == CVCS_F32 ? 4 : 2;
    auto aligned = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const bool ptr16 = aligned(logits, 16) && aligned(dlogits, 16) && aligned(target, 16) && aligned(argmax, 16);
    bool handled = false;
    int rc = CVCS_OK;

    if (p.tw_mode != 0 && !(forced != 3 && C >= 2 && C <= kMaxRegC && ptr16)) return tw_fallback(p, logits_dtype, layout, target_dtype, stream);
    if (forced != 3 && C >= 2 && C <= kMaxRegC && ptr16) {
        // ---- TMA-staged: every bulk copy must be a multiple of 16 bytes
        const bool tma_ok = layout == CVCS_NCHW ? (hw % 16 == 0) : (n_pixels % 16 == 0);
        if (p.tw_mode != 0 && !(tma_ok && forced != 2 && forced != 3 && (target_dtype == CVCS_U8 || p.tw_mode == 2)))
            return tw_fallback(p, logits_dtype, layout, target_dtype, stream);
        if (tma_ok && forced != 2) {
            rc = logits_dtype == CVCS_F32 ? ce_tma_launch_f32(p, layout, stream, &handled)
                                          : ce_tma_launch_bf16(p, layout, stream, &handled);
            if (handled) return rc;
            if (p.tw_mode != 0) return tw_fallback(p, logits_dtype, layout, target_dtype, stream);
        }
        // ---- direct NCHW: f32 4 pixels/thread; bf16 8 pixels/thread up to C=12, else 4
        if (layout == CVCS_NCHW && forced != 1) {
            const int vec = logits_dtype == CVCS_F32 ? 4 : (C <= 12 ? 8 : 4);
            if (hw % vec == 0) {
                p.n_items = n_pixels / vec;
                p.items_per_image = static_cast<unsigned int>(hw / vec);
                rc = ce_direct_launch(p, logits_dtype, vec, stream, &handled);
                if (handled) return rc;
            }
        }
    }
    (void)esize;
    rc = try_wide(p, logits_dtype, layout, ptr16, stream, &handled);
    if (handled || rc) return rc;
    return ce_generic_launch(p, logits_dtype, layout, stream);
}

}  // namespace cvcs
