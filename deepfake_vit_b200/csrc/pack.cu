// Weight packing for inference: folds eval-mode BatchNorm into the preceding convolution / Linear and lays the
// result out as the inference kernels read it (dfv_blob_* slots, transposed classifier / attention matrices) --
// straight from the module's own fp32 torch-layout parameter storage (the dfv_train_index table), in a handful
// of launches.  Replaces the host-side `w * gamma / sqrt(var + eps)` arithmetic of the first round (hundreds of
// stock elementwise launches per re-pack) so that NO arithmetic of the path runs outside this library.
//
//   scale[r] = gamma[r] / sqrt(running_var[r] + eps)          (1 without BatchNorm)
//   dst[r][c] (or dst[c][r]) = src[r][c] * scale[r]
//   dst_bias[r] = lin_bias[r] * scale[r] + (beta[r] - running_mean[r] * scale[r])
#include <vector>

#include "common.cuh"

namespace dfv {

struct PackJob {
  const float* src;        // [rows][cols] row-major fp32
  const float *g, *b, *rm, *rv;   // BatchNorm over rows (all NULL: no BatchNorm)
  const float* lin_bias;   // [rows] or NULL
  void* dst;
  float* dst_bias;         // [rows] or NULL
  int rows, cols;
  int mode;                // 0: dst[r][c]   1: dst[c][r]   2: stem: dst[((kh*3+kw)*3+ci)][r] from c = ci*9 + kh*3 + kw
  int out_bf16;
  float eps;
};

constexpr int kPackJobsPerLaunch = 40;
struct PackBatch {
  PackJob job[kPackJobsPerLaunch];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackBatch batch) {
  const PackJob& j = batch.job[blockIdx.y];
  const long long total = (long long)j.rows * j.cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // walk the DESTINATION index so that stores coalesce
    int r, c;
    if (j.mode == 0) {
      r = (int)(i / j.cols);
      c = (int)(i % j.cols);
    } else {
      r = (int)(i % j.rows);
      const int cd = (int)(i / j.rows);
      if (j.mode == 2) {
        const int ci = cd % 3, kw = (cd / 3) % 3, kh = cd / 9;
        c = ci * 9 + kh * 3 + kw;
      } else {
        c = cd;
      }
    }
    float s = 1.0f;
    if (j.g) s = j.g[r] / sqrtf(j.rv[r] + j.eps);
    const float v = j.src[(size_t)r * j.cols + c] * s;
    if (j.out_bf16) reinterpret_cast<__nv_bfloat16*>(j.dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(j.dst)[i] = v;
  }
  if (j.dst_bias && blockIdx.x == 0) {
    for (int r = threadIdx.x; r < j.rows; r += blockDim.x) {
      float s = 1.0f, shift = 0.0f;
      if (j.g) {
        s = j.g[r] / sqrtf(j.rv[r] + j.eps);
        shift = j.b[r] - j.rm[r] * s;
      }
      j.dst_bias[r] = (j.lin_bias ? j.lin_bias[r] * s : 0.0f) + shift;
    }
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_pack_weights(const dfv_pack_args* a, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(a && a->params && a->blob, "dfv_pack_weights: null pointer");
  DFV_REQUIRE(valid_dtype(a->dtype), "dfv_pack_weights: bad dtype %d", a->dtype);
  DFV_REQUIRE(a->head_layers >= 0 && a->head_layers <= DFV_MAX_CLS_LAYERS, "dfv_pack_weights: bad classifier depth %d", a->head_layers);
  const int dtype = a->dtype;
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  const int stem_c = topo_stem_c(), head_c = topo_head_c();
  auto P = [&](int block, int kind) -> const float* { return a->params[dfv_train_index(block, kind)]; };
  auto W_ = [&](int block, int kind) -> void* { return const_cast<void*>(blob_ptr(a->blob, dtype, block, kind)); };
  std::vector<PackJob> jobs;
  auto add = [&](const float* src, const float* g, const float* b, const float* rm, const float* rv, const float* lin_bias, void* dst,
                 void* dst_bias, int rows, int cols, int mode, int out_bf16, float eps) {
    PackJob j;
    j.src = src; j.g = g; j.b = b; j.rm = rm; j.rv = rv; j.lin_bias = lin_bias;
    j.dst = dst; j.dst_bias = static_cast<float*>(dst_bias);
    j.rows = rows; j.cols = cols; j.mode = mode; j.out_bf16 = out_bf16; j.eps = eps;
    jobs.push_back(j);
  };
  const int wbf = dtype == DFV_BF16;
  const float eps = a->bn_eps;
  DFV_REQUIRE(P(-1, DFV_TG_STEM_W) && P(-1, DFV_TG_STEM_G) && P(-1, DFV_TG_STEM_B) && P(-1, DFV_TG_STEM_RM) && P(-1, DFV_TG_STEM_RV),
              "dfv_pack_weights: stem parameters missing");
  add(P(-1, DFV_TG_STEM_W), P(-1, DFV_TG_STEM_G), P(-1, DFV_TG_STEM_B), P(-1, DFV_TG_STEM_RM), P(-1, DFV_TG_STEM_RV), nullptr,
      W_(-1, DFV_W_STEM), W_(-1, DFV_W_STEM_BIAS), stem_c, 27, 2, 0, eps);
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    const int kk = b.kernel * b.kernel;
    DFV_REQUIRE(P(i, DFV_T_DW_W) && P(i, DFV_T_BN1_G) && P(i, DFV_T_BN1_RV) && P(i, DFV_T_SE_R_W) && P(i, DFV_T_SE_R_B) && P(i, DFV_T_SE_E_W) &&
                    P(i, DFV_T_SE_E_B) && P(i, DFV_T_PROJ_W) && P(i, DFV_T_BN2_G) && P(i, DFV_T_BN2_RV),
                "dfv_pack_weights: block %d parameters missing", i);
    if (b.has_expand) {
      DFV_REQUIRE(P(i, DFV_T_EXPAND_W) && P(i, DFV_T_BN0_G) && P(i, DFV_T_BN0_RV), "dfv_pack_weights: block %d expand parameters missing", i);
      add(P(i, DFV_T_EXPAND_W), P(i, DFV_T_BN0_G), P(i, DFV_T_BN0_B), P(i, DFV_T_BN0_RM), P(i, DFV_T_BN0_RV), nullptr, W_(i, DFV_W_EXPAND),
          W_(i, DFV_W_EXPAND_BIAS), b.c_mid, b.c_in, 0, wbf, eps);
    }
    add(P(i, DFV_T_DW_W), P(i, DFV_T_BN1_G), P(i, DFV_T_BN1_B), P(i, DFV_T_BN1_RM), P(i, DFV_T_BN1_RV), nullptr, W_(i, DFV_W_DW),
        W_(i, DFV_W_DW_BIAS), b.c_mid, kk, 1, 0, eps);
    add(P(i, DFV_T_SE_R_W), nullptr, nullptr, nullptr, nullptr, P(i, DFV_T_SE_R_B), W_(i, DFV_W_SE_REDUCE), W_(i, DFV_W_SE_REDUCE_BIAS),
        b.se_squeeze, b.c_mid, 0, 0, eps);
    add(P(i, DFV_T_SE_E_W), nullptr, nullptr, nullptr, nullptr, P(i, DFV_T_SE_E_B), W_(i, DFV_W_SE_EXPAND), W_(i, DFV_W_SE_EXPAND_BIAS),
        b.c_mid, b.se_squeeze, 1, 0, eps);
    add(P(i, DFV_T_PROJ_W), P(i, DFV_T_BN2_G), P(i, DFV_T_BN2_B), P(i, DFV_T_BN2_RM), P(i, DFV_T_BN2_RV), nullptr, W_(i, DFV_W_PROJECT),
        W_(i, DFV_W_PROJECT_BIAS), b.c_out, b.c_mid, 0, wbf, eps);
  }
  DFV_REQUIRE(P(-1, DFV_TG_HEAD_W) && P(-1, DFV_TG_HEAD_G) && P(-1, DFV_TG_HEAD_RV), "dfv_pack_weights: head parameters missing");
  add(P(-1, DFV_TG_HEAD_W), P(-1, DFV_TG_HEAD_G), P(-1, DFV_TG_HEAD_B), P(-1, DFV_TG_HEAD_RM), P(-1, DFV_TG_HEAD_RV), nullptr, W_(-1, DFV_W_HEAD),
      W_(-1, DFV_W_HEAD_BIAS), head_c, blk[n - 1].c_out, 0, wbf, eps);
  // classifier: Linear (+ BatchNorm1d) -> transposed folded weight [din][dout], folded bias
  for (int l = 0; l < a->head_layers; ++l) {
    DFV_REQUIRE(a->head_dims && a->head_w_t && a->head_b && a->head_w_t[l] && a->head_b[l], "dfv_pack_weights: classifier outputs missing");
    const float* w = a->params[dfv_train_cls_index(l, 0)];
    const float* bias = a->params[dfv_train_cls_index(l, 1)];
    DFV_REQUIRE(w && bias, "dfv_pack_weights: classifier layer %d parameters missing", l);
    const float* g = a->params[dfv_train_cls_index(l, 2)];
    add(w, g, g ? a->params[dfv_train_cls_index(l, 3)] : nullptr, g ? a->params[dfv_train_cls_index(l, 4)] : nullptr,
        g ? a->params[dfv_train_cls_index(l, 5)] : nullptr, bias, a->head_w_t[l], a->head_b[l], a->head_dims[l + 1], a->head_dims[l], 1, 0,
        a->cls_bn_eps);
  }
  if (a->ca_w2_t && a->ca_hidden > 0) {
    DFV_REQUIRE(P(-1, DFV_TG_CA_W2), "dfv_pack_weights: channel-attention weights missing");
    add(P(-1, DFV_TG_CA_W2), nullptr, nullptr, nullptr, nullptr, nullptr, a->ca_w2_t, nullptr, head_c, a->ca_hidden, 1, 0, eps);
  }
  cudaStream_t st = as_stream(stream);
  for (size_t j0 = 0; j0 < jobs.size(); j0 += kPackJobsPerLaunch) {
    PackBatch batch;
    const int nj = (int)std::min<size_t>(kPackJobsPerLaunch, jobs.size() - j0);
    long long max_elems = 0;
    for (int j = 0; j < nj; ++j) {
      batch.job[j] = jobs[j0 + j];
      max_elems = std::max(max_elems, (long long)jobs[j0 + j].rows * jobs[j0 + j].cols);
    }
    const unsigned gx = (unsigned)std::min<long long>(64, (max_elems + 2047) / 2048);
    pack_weights_kernel<<<dim3(gx ? gx : 1, nj), 256, 0, st>>>(batch);
    DFV_LAUNCH_CHECK();
  }
  return DFV_OK;
}
