// Error reporting, version, and the fused per-batch entry points of the C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

namespace lgcn {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int propagate_fwd_impl(const lgcn_graph *, const float *, const float *, int, float *, float *, float *, size_t,
                       cudaStream_t);
int propagate_bwd_impl(const lgcn_graph *, const float *, int, const float *, const float *, const int32_t *, float,
                       float *, double *, float *, size_t, cudaStream_t);
int bpr_impl(const lgcn_graph *, const float *, const float *, const int64_t *, float *, int32_t *, float *,
             double *, bool, int, int, int, int, cudaStream_t);
int reg_value_impl(const lgcn_graph *, const float *, const float *, const int32_t *, double *, cudaStream_t);
int loss_finalize_impl(const double *, int64_t, float, float *, cudaStream_t);

}  // namespace lgcn

extern "C" const char *lgcn_last_error(void) { return lgcn::g_err; }
extern "C" int lgcn_version(void) { return 200; }

static int check_buffers(const lgcn_step_buffers *b, bool grad) {
    LGCN_REQUIRE(b && b->final_emb && b->rnorm && b->neg_count && b->accum, LGCN_E_INVALID,
                 "step buffers: null forward buffer");
    LGCN_REQUIRE(!grad || (b->grad_final && b->grad_e0 && b->trip_scratch), LGCN_E_INVALID,
                 "step buffers: null gradient buffer");
    return LGCN_OK;
}

// utils/train_test.py:88-96: zero_grad -> forward -> bpr_loss -> backward -> clip -> Adam.step
extern "C" int lgcn_train_step(const lgcn_graph *g, float *user_w, float *item_w, int num_layers,
                               const int64_t *neg, float bpr_coeff, const lgcn_adam *opt,
                               const lgcn_step_buffers *buf, float *loss_out, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && user_w && item_w && neg && opt, LGCN_E_INVALID, "train_step: null argument");
    int rc = check_buffers(buf, true);
    if (rc) return rc;
    LGCN_REQUIRE(g->num_triplets > 0, LGCN_E_INVALID,
                 "train_step: batch has no user->movie edge (the reference's loss would be NaN)");
    if ((rc = lgcn_step_begin(opt, buf->accum, stream))) return rc;
    if ((rc = propagate_fwd_impl(g, user_w, item_w, num_layers, buf->final_emb, buf->rnorm, buf->work,
                                 buf->work_bytes, st))) return rc;
    if ((rc = bpr_impl(g, buf->final_emb, buf->rnorm, neg, buf->grad_final, buf->neg_count, buf->trip_scratch,
                       buf->accum, true, 0, g->n_out_user_tasks, 0, g->num_users, st))) return rc;
    // regulariser gradient 2*coeff/(64 P) * cnt[r] * e0[r] is folded into the last backward layer
    const float reg_coef = 2.0f * bpr_coeff / (64.0f * (float)g->num_triplets);
    if ((rc = propagate_bwd_impl(g, buf->grad_final, num_layers, user_w, item_w, buf->neg_count, reg_coef,
                                 buf->grad_e0, buf->accum, buf->work, buf->work_bytes, st))) return rc;
    return lgcn_clip_adam(opt, user_w, item_w, g->num_users, g->num_nodes - g->num_users, buf->grad_e0,
                          buf->accum, g->num_triplets, bpr_coeff, loss_out, stream);
}

// evaluate() loss (utils/train_test.py:153-156): forward over the given edges + BPR value.
extern "C" int lgcn_eval_loss(const lgcn_graph *g, const float *user_w, const float *item_w, int num_layers,
                              const int64_t *neg, float bpr_coeff, const lgcn_step_buffers *buf,
                              float *loss_out, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && user_w && item_w && neg && loss_out, LGCN_E_INVALID, "eval_loss: null argument");
    int rc = check_buffers(buf, false);
    if (rc) return rc;
    LGCN_REQUIRE(g->num_triplets > 0, LGCN_E_INVALID, "eval_loss: no user->movie edge");
    if ((rc = lgcn_step_begin(nullptr, buf->accum, stream))) return rc;
    if ((rc = propagate_fwd_impl(g, user_w, item_w, num_layers, buf->final_emb, buf->rnorm, buf->work,
                                 buf->work_bytes, st))) return rc;
    if ((rc = bpr_impl(g, buf->final_emb, buf->rnorm, neg, nullptr, buf->neg_count, nullptr, buf->accum, false, 0,
                       g->n_out_user_tasks, 0, g->num_users, st))) return rc;
    if ((rc = reg_value_impl(g, user_w, item_w, buf->neg_count, buf->accum + 1, st))) return rc;
    return loss_finalize_impl(buf->accum, g->num_triplets, bpr_coeff, loss_out, st);
}
