// K6: clip_grad_norm_(max_norm=1) + dense Adam over both embedding tables, one streaming pass.
//
// Replaces torch.nn.utils.clip_grad_norm_(model.parameters(), 1) and optimizer.step()
// (/root/reference/utils/train_test.py:95-96; Adam(lr=1e-3) created at :236).  The update is
// torch.optim.Adam's (amsgrad=False, weight_decay=0, maximize=False):
//     m = m + (1-b1) (g - m);  v = b2 v + (1-b2) g^2
//     p = p - (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// and it is DENSE on purpose: rows with zero gradient still move through their momentum
// (SURVEY.md App. B #11).  HBM-streaming kernel: 4 reads + 3 writes of N*256 B.
#include "adam.cuh"

namespace lgcn {

__global__ void step_begin_kernel(int64_t *step, double *accum, int32_t *list_count) {
    if (step) step[0] += 1;
    accum[0] = 0.0; accum[1] = 0.0; accum[2] = 0.0; accum[3] = 0.0;
    if (list_count) list_count[0] = 0;
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float4 *__restrict__ user_w, float4 *__restrict__ item_w, size_t user_vec, size_t vec_begin,
                 size_t total_vec, const float4 *__restrict__ grad, float4 *__restrict__ m, float4 *__restrict__ v,
                 const double *__restrict__ accum, const int64_t *__restrict__ step, AdamHyper h, int64_t P,
                 float coeff, float *loss_out, int32_t *__restrict__ row_step) {
    const AdamScalars a = adam_scalars(h, step[0]);
    const int t32 = (int)step[0];
    const float clip = clip_coef(h, accum[2]);
    if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
        const double p = (double)P;
        loss_out[0] = (float)(-accum[0] / (10.0 * p) + (double)coeff * accum[1] / (64.0 * p));
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = vec_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
        float4 *pp = i < user_vec ? user_w + i : item_w + (i - user_vec);
        const float4 g4 = __ldcs(grad + i);
        float4 p4 = *pp, m4 = m[i], v4 = v[i];
        adam_vec(p4, m4, v4, g4, clip, a);
        *pp = p4; m[i] = m4; v[i] = v4;
        if (row_step && (i % D4) == 0) row_step[i / D4] = t32;     // every row is now at step t
    }
}

}  // namespace lgcn

extern "C" int lgcn_step_begin(const lgcn_adam *opt, double *accum, void *stream) {
    LGCN_REQUIRE(accum, LGCN_E_INVALID, "step_begin: null accum");
    lgcn::step_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt ? opt->step : nullptr, accum, nullptr);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_clip_adam_rows(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users,
                                   int64_t num_items, const float *grad, const double *accum,
                                   int64_t num_triplets, float bpr_coeff, float *loss_out, int64_t row_begin,
                                   int64_t row_end, void *stream) {
    LGCN_REQUIRE(opt && user_w && item_w && grad && accum && opt->m && opt->v && opt->step, LGCN_E_INVALID,
                 "clip_adam: null argument");
    LGCN_REQUIRE(row_begin >= 0 && row_end <= num_users + num_items && row_begin <= row_end, LGCN_E_INVALID,
                 "clip_adam: bad row range");
    const size_t user_vec = (size_t)num_users * lgcn::D4;
    const size_t vb = (size_t)row_begin * lgcn::D4, ve = (size_t)row_end * lgcn::D4;
    static int sms_of[64];                               // SM count, cached per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev >= 0 && dev < 64 ? dev : 0;
    if (sms_of[dev] == 0) {
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms_of[dev] <= 0) sms_of[dev] = 148;
    }
    const int sms = sms_of[dev];
    const int64_t want = (int64_t)((ve - vb + 255) / 256);
    const int grid = (int)(want < (int64_t)sms * 8 ? (want > 0 ? want : 1) : (int64_t)sms * 8);
    lgcn::clip_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(user_w), reinterpret_cast<float4 *>(item_w), user_vec, vb, ve,
        reinterpret_cast<const float4 *>(grad), reinterpret_cast<float4 *>(opt->m),
        reinterpret_cast<float4 *>(opt->v), accum, opt->step, lgcn::make_hyper(opt), num_triplets, bpr_coeff, loss_out,
        opt->row_step);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_clip_adam(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users,
                              int64_t num_items, const float *grad, const double *accum,
                              int64_t num_triplets, float bpr_coeff, float *loss_out, void *stream) {
    return lgcn_clip_adam_rows(opt, user_w, item_w, num_users, num_items, grad, accum, num_triplets, bpr_coeff,
                               loss_out, 0, num_users + num_items, stream);
}
