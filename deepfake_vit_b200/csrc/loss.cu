// CombinedLoss (src/training/losses.py:192-247) forward + gradient in one single-CTA kernel.
//   ce          = sum_i w[y_i] * l_i / sum_i w[y_i]            (nn.CrossEntropyLoss(weight))
//   focal       = mean_i alpha[y_i] * (1 - p_i)^2 * l_i         (FocalLoss, gamma = 2)
//   contrastive = mean_j  s_j * d_j^2 + (1 - s_j) * relu(1 - d_j)^2 over pairs (2j, 2j+1),
//                 d_j = || f_2j - f_2j+1 + 1e-6 ||_2,  s_j = [y_2j == y_2j+1]
//   total       = w_ce * ce + w_focal * focal + w_con * contrastive
// All fp32.  Bytes are negligible (B x (C + D) floats); the point is one launch instead of ~30.
#include "common.cuh"

namespace dfv {

__global__ void __launch_bounds__(256) combined_loss_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ targets,
                                                           const float* __restrict__ features,
                                                           const float* __restrict__ cw, float w_ce, float w_focal,
                                                           float w_con, float* __restrict__ losses,
                                                           float* __restrict__ dlogits, float* __restrict__ dfeat,
                                                           int B, int C, int D, const float* __restrict__ ce_norm) {
  pdl_prologue();
  __shared__ float red[3][8];
  __shared__ float s_wsum, s_ce_num, s_focal;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- pass 1: per-sample CE terms -> sums
  float wsum = 0.f, ce_num = 0.f, focal = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float* z = logits + (size_t)i * C;
    const int y = (int)targets[i];
    float mx = z[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
    const float l = (mx + logf(se)) - z[y];        // -log softmax[y]
    const float wy = cw ? cw[y] : 1.f;
    const float pt = expf(-l);
    wsum += wy;
    ce_num += wy * l;
    focal += wy * (1.f - pt) * (1.f - pt) * l;
  }
  wsum = warp_sum(wsum); ce_num = warp_sum(ce_num); focal = warp_sum(focal);
  if (lane == 0) { red[0][warp] = wsum; red[1][warp] = ce_num; red[2][warp] = focal; }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, b2 = 0.f, c2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b2 += red[1][w]; c2 += red[2][w]; }
    s_wsum = a; s_ce_num = b2; s_focal = c2 / (float)B;
  }
  __syncthreads();
  // the weighted mean's normaliser: the local sum of w[y_i], or the caller's (data-parallel exact form)
  const float inv_wsum = 1.f / (ce_norm != nullptr ? ce_norm[0] : s_wsum);

  // ---- gradient wrt logits
  if (dlogits) {
    for (int i = tid; i < B; i += blockDim.x) {
      const float* z = logits + (size_t)i * C;
      const int y = (int)targets[i];
      float mx = z[0];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
      const float lse = mx + logf(se);
      const float l = lse - z[y];
      const float wy = cw ? cw[y] : 1.f;
      const float pt = expf(-l);
      // d focal_i / d l = (1-pt)^2 + 2 (1-pt) pt l      (since d pt / d l = -pt)
      const float dfl = (1.f - pt) * (1.f - pt) + 2.f * (1.f - pt) * pt * l;
      const float coef = w_ce * wy * inv_wsum + w_focal * wy * dfl / (float)B;
      for (int c = 0; c < C; ++c) {
        const float sm = expf(z[c] - lse);
        dlogits[(size_t)i * C + c] = coef * (sm - (c == y ? 1.f : 0.f));   // d l / d z_c
      }
    }
  }

  // ---- contrastive over consecutive pairs: one warp per pair
  const int pairs = (features != nullptr && B >= 2) ? B / 2 : 0;
  float con = 0.f;
  for (int j = warp; j < pairs; j += (blockDim.x >> 5)) {
    const float* f1 = features + (size_t)(2 * j) * D;
    const float* f2 = f1 + D;
    float ss = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float df = (f1[k] - f2[k]) + 1e-6f;
      ss = fmaf(df, df, ss);
    }
    ss = warp_sum(ss);
    const float d = sqrtf(ss);
    const float same = targets[2 * j] == targets[2 * j + 1] ? 1.f : 0.f;
    const float hinge = fmaxf(1.f - d, 0.f);
    if (lane == 0) con += same * d * d + (1.f - same) * hinge * hinge;
    if (dfeat) {
      // d/d d = 2 same d - 2 (1-same) hinge ; d d / d f1 = (f1 - f2 + eps) / d
      const float g = d > 0.f ? (w_con / (float)pairs) * (2.f * same * d - 2.f * (1.f - same) * hinge) / d : 0.f;
      for (int k = lane; k < D; k += 32) {
        const float df = (f1[k] - f2[k]) + 1e-6f;
        dfeat[(size_t)(2 * j) * D + k] = g * df;
        dfeat[(size_t)(2 * j + 1) * D + k] = -g * df;
      }
    }
  }
  if (dfeat && (B & 1) && features) {   // the unpaired last sample gets no contrastive gradient
    for (int k = tid; k < D; k += blockDim.x) dfeat[(size_t)(B - 1) * D + k] = 0.f;
  }
  __syncthreads();
  if (lane == 0) red[0][warp] = con;
  __syncthreads();
  if (tid == 0) {
    float c2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c2 += red[0][w];
    const float ce = s_ce_num * inv_wsum;
    const float contrastive = pairs > 0 ? c2 / (float)pairs : 0.f;
    losses[0] = ce;
    losses[1] = s_focal;
    losses[2] = contrastive;
    float total = 0.f;
    if (w_ce > 0.f) total += w_ce * ce;
    if (w_focal > 0.f) total += w_focal * s_focal;
    if (pairs > 0 && w_con > 0.f) total += w_con * contrastive;
    losses[3] = total;
  }
}

// sum_i w[y_i] by ONE warp in a fixed order (the normaliser every rank contributes to the global weighted CE)
__global__ void __launch_bounds__(32) class_weight_sum_kernel(const long long* __restrict__ targets, const float* __restrict__ cw,
                                                             float* __restrict__ out, int B) {
  pdl_prologue();
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += 32) s += cw ? cw[(int)targets[i]] : 1.f;
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = s;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_class_weight_sum(const int64_t* targets, const float* class_weights, float* out, int B, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(targets && out && B > 0 && C > 1, "dfv_class_weight_sum: bad arguments");
  DFV_PDL((class_weight_sum_kernel), 1, 32, 0, as_stream(stream), (const long long*)targets, class_weights, out, B);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

extern "C" int dfv_combined_loss_fwd_bwd(const float* logits, const int64_t* targets, const float* features,
                                         const float* class_weights, float w_ce, float w_focal, float w_contrastive,
                                         float* losses, float* dlogits, float* dfeatures, int B, int C, int D,
                                         int* has_contrastive, dfv_stream_t stream) {
  return dfv_combined_loss_fwd_bwd_ex(logits, targets, features, class_weights, w_ce, w_focal, w_contrastive, losses, dlogits, dfeatures,
                                      B, C, D, has_contrastive, nullptr, stream);
}

extern "C" int dfv_combined_loss_fwd_bwd_ex(const float* logits, const int64_t* targets, const float* features,
                                            const float* class_weights, float w_ce, float w_focal, float w_contrastive,
                                            float* losses, float* dlogits, float* dfeatures, int B, int C, int D,
                                            int* has_contrastive, const float* ce_norm, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(logits && targets && losses, "dfv_combined_loss_fwd_bwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 1 && (features == nullptr || D > 0), "dfv_combined_loss_fwd_bwd: bad shape");
  const bool con = features != nullptr && B >= 2 && w_contrastive > 0.f;
  if (has_contrastive) *has_contrastive = con ? 1 : 0;
  // weights <= 0 switch a term off exactly as `self.weights[k] > 0` does (losses.py:216,222,228)
  ProfScope prof(PK_LOSS, 4.0 * B * (2.0 * C + 2.0 * D), 10.0 * B * (C + D), as_stream(stream));
  DFV_PDL((combined_loss_kernel), 1, 256, 0, as_stream(stream), logits, (const long long*)targets, con ? features : nullptr,
                                                        class_weights, w_ce > 0.f ? w_ce : 0.f,
                                                        w_focal > 0.f ? w_focal : 0.f, con ? w_contrastive : 0.f, losses,
                                                        dlogits, con ? dfeatures : nullptr, B, C, D, ce_norm);
  DFV_LAUNCH_CHECK();
  if (!con && dfeatures && features)
    DFV_CUDA(cudaMemsetAsync(dfeatures, 0, sizeof(float) * (size_t)B * D, as_stream(stream)));
  return DFV_OK;
}
