// HybridAttention + global average pool, training path: a forward that saves what backward needs and
// the backward itself (landmark_attention.py:283-310 under autograd).  One CTA per image; the
// 12x12x1792 map is L2 resident.  All arithmetic fp32 (the autocast reference promotes here too).
//
// forward (per image; x = head activation [P][C], A = landmark map or 1):
//   x1 = x*A;  avg_c, max_c over p;  gc = sigmoid(W2 (relu(W1 avg) + relu(W1 max)));  x2 = x1*gc
//   sm_p, sx_p = mean / max over c of x2;  gp = sigmoid(conv7x7([sm, sx]));  f_c = mean_p x2*gp
// backward: see hybrid_attention_bwd_kernel.
#include "common.cuh"

namespace dfv {

struct AttnSaved {
  float* avg;      // [B][C]
  float* mx;       // [B][C]
  int* mx_idx;     // [B][C]   argmax position of x1 per channel
  float* hpre;     // [B][2][hidden]  W1 avg, W1 max (pre-ReLU)
  float* gate_c;   // [B][C]
  float* sp_mean;  // [B][P]
  float* sp_max;   // [B][P]
  int* sp_idx;     // [B][P]   argmax channel of x2 per position
  float* gate_p;   // [B][P]
};

// One CTA of 1024 threads per image.  The first version ran 256 threads with one thread per 8-channel vector walking all
// positions and one warp per row of W1: ~28 KB of map and 8 KB of weights in flight per SM, 226 us at batch 64 for 33 MB of
// map (5 % of the HBM roofline).  Now kTrPhases threads share a vector and take every kTrPhases-th position (partials meet in
// shared memory in a fixed order), 32 warps walk W1 / W2 with 14 / 4 rows x 128 B in flight each.
constexpr int kTrPhases = 4;
template <typename T>
__global__ void __launch_bounds__(1024) hybrid_attention_train_fwd_kernel(
    const T* __restrict__ fmap, const float* __restrict__ heat, const float* __restrict__ w1,
    const float* __restrict__ w2, const float* __restrict__ sa_w, float* __restrict__ features, AttnSaved sv, int H,
    int W, int C, int hidden, int use_channel, int use_spatial) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int HW = H * W;
  float* a_lm = sm;
  float* avg_c = a_lm + HW;
  float* max_c = avg_c + C;
  float* gate_c = max_c + C;
  float* hid = gate_c + C;
  float* sp_mean = hid + hidden;
  float* sp_max = sp_mean + HW;
  float* gate_p = sp_max + HW;
  float* red = gate_p + HW;                                       // [kTrPhases - 1][3][C]: partial sum / max / argmax of a position phase

  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const T* fb = fmap + (size_t)b * HW * C;
  const int CV = C >> 3;
  const float inv_hw = 1.0f / (float)HW;
  const int nvec = blockDim.x / kTrPhases;                        // 8-channel vectors per sweep
  const int ph = tid / nvec, lv = tid % nvec;

  for (int p = tid; p < HW; p += blockDim.x) a_lm[p] = heat ? heat[(size_t)b * HW + p] : 1.0f;
  __syncthreads();

  if (use_channel) {
    for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
      const int cv = cv0 + lv;
      float s[8], m[8];
      int mi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] = 0.f; m[e] = -INFINITY; mi[e] = 0; }
      if (cv < CV) {
#pragma unroll 4
        for (int p = ph; p < HW; p += kTrPhases) {
          float v[8];
          load8(fb + (size_t)p * C + cv * 8, v);
          const float a = a_lm[p];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x = v[e] * a;
            s[e] += x;
            if (x > m[e]) { m[e] = x; mi[e] = p; }
          }
        }
        if (ph > 0) {
          float* r = red + (size_t)(ph - 1) * 3 * C + cv * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) { r[e] = s[e]; r[C + e] = m[e]; r[2 * C + e] = __int_as_float(mi[e]); }
        }
      }
      __syncthreads();
      if (ph == 0 && cv < CV) {
#pragma unroll
        for (int q = 0; q < kTrPhases - 1; ++q) {                 // fixed order; the FIRST position of the maximum wins (torch's argmax)
          const float* r = red + (size_t)q * 3 * C + cv * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            s[e] += r[e];
            const float om = r[C + e];
            const int oi = __float_as_int(r[2 * C + e]);
            if (om > m[e] || (om == m[e] && oi < mi[e])) { m[e] = om; mi[e] = oi; }
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cv * 8 + e;
          avg_c[c] = s[e] * inv_hw;
          max_c[c] = m[e];
          sv.avg[(size_t)b * C + c] = s[e] * inv_hw;
          sv.mx[(size_t)b * C + c] = m[e];
          sv.mx_idx[(size_t)b * C + c] = mi[e];
        }
      }
      __syncthreads();
    }
    for (int j = warp; j < hidden; j += nwarps) {
      const float* wr = w1 + (size_t)j * C;
      float sa = 0.f, sx = 0.f;
#pragma unroll 14
      for (int c = lane; c < C; c += 32) {
        const float wv = __ldg(wr + c);
        sa = fmaf(wv, avg_c[c], sa);
        sx = fmaf(wv, max_c[c], sx);
      }
      sa = warp_sum(sa);
      sx = warp_sum(sx);
      if (lane == 0) {
        hid[j] = fmaxf(sa, 0.f) + fmaxf(sx, 0.f);
        sv.hpre[((size_t)b * 2 + 0) * hidden + j] = sa;
        sv.hpre[((size_t)b * 2 + 1) * hidden + j] = sx;
      }
    }
    __syncthreads();
    // W2 [C][hidden]: one warp per group of four rows, lanes along the row (coalesced; a thread per row read 32 sectors per request)
    for (int c0 = warp * 4; c0 < C; c0 += nwarps * 4) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = lane; j < hidden; j += 32) {
        const float h = hid[j];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + u < C) s[u] = fmaf(__ldg(w2 + (size_t)(c0 + u) * hidden + j), h, s[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] = warp_sum(s[u]);
      if (lane < 4 && c0 + lane < C) {
        const float g = sigmoid_exact(lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3]);
        gate_c[c0 + lane] = g;
        sv.gate_c[(size_t)b * C + c0 + lane] = g;
      }
    }
  } else {
    for (int c = tid; c < C; c += blockDim.x) gate_c[c] = 1.0f;
  }
  __syncthreads();

  if (use_spatial) {
    for (int p = warp; p < HW; p += nwarps) {
      const float a = a_lm[p];
      float s = 0.f, m = -INFINITY;
      int mi = 0;
      #pragma unroll 8
      for (int cv = lane; cv < CV; cv += 32) {
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = (v[e] * a) * gate_c[cv * 8 + e];
          s += x;
          if (x > m) { m = x; mi = cv * 8 + e; }
        }
      }
      s = warp_sum(s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
      }
      if (lane == 0) {
        sp_mean[p] = s / (float)C;
        sp_max[p] = m;
        sv.sp_mean[(size_t)b * HW + p] = s / (float)C;
        sv.sp_max[(size_t)b * HW + p] = m;
        sv.sp_idx[(size_t)b * HW + p] = mi;
      }
    }
    __syncthreads();
    for (int p = tid; p < HW; p += blockDim.x) {
      const int y = p / W, x = p % W;
      float s = 0.f;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = x + kx - 3;
          if (xx < 0 || xx >= W) continue;
          s = fmaf(sa_w[ky * 7 + kx], sp_mean[yy * W + xx], s);
          s = fmaf(sa_w[49 + ky * 7 + kx], sp_max[yy * W + xx], s);
        }
      }
      const float g = sigmoid_exact(s);
      gate_p[p] = g;
      sv.gate_p[(size_t)b * HW + p] = g;
    }
  } else {
    for (int p = tid; p < HW; p += blockDim.x) gate_p[p] = 1.0f;
  }
  __syncthreads();

  for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
    const int cv = cv0 + lv;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    if (cv < CV) {
      float gc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gc[e] = gate_c[cv * 8 + e];
#pragma unroll 4
      for (int p = ph; p < HW; p += kTrPhases) {
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
        const float a = a_lm[p], g = gate_p[p];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += ((v[e] * a) * gc[e]) * g;
      }
      if (ph > 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) red[(size_t)(ph - 1) * C + cv * 8 + e] = s[e];
      }
    }
    __syncthreads();
    if (ph == 0 && cv < CV) {
#pragma unroll
      for (int q = 0; q < kTrPhases - 1; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += red[(size_t)q * C + cv * 8 + e];
#pragma unroll
      for (int e = 0; e < 8; ++e) features[(size_t)b * C + cv * 8 + e] = s[e] * inv_hw;
    }
    __syncthreads();
  }
}

// Backward.  df: gradient wrt the pooled features [B][C].  Writes dx (gradient wrt the head activation,
// T), dA (gradient wrt the landmark map, fp32 [B][P], if heat != NULL), the per-image quantities the
// weight-gradient kernel reduces over the batch (dz [B][C], dhpre [B][2][hidden]) and adds the 7x7 conv
// weight gradient into dsa_w[98] atomically.
template <typename T>
__global__ void __launch_bounds__(512) hybrid_attention_bwd_kernel(
    const T* __restrict__ fmap, const float* __restrict__ heat, const float* __restrict__ w1,
    const float* __restrict__ w2, const float* __restrict__ sa_w, const float* __restrict__ df, AttnSaved sv,
    T* __restrict__ dx, float* __restrict__ dA, float* __restrict__ dz_out, float* __restrict__ dhpre_out,
    float* __restrict__ dsa_w, int H, int W, int C, int hidden, int use_channel, int use_spatial) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int HW = H * W;
  float* a_lm = sm;                 // [HW]
  float* gate_p = a_lm + HW;        // [HW]
  float* dt = gate_p + HW;          // [HW]
  float* dsm = dt + HW;             // [HW]
  float* dsx = dsm + HW;            // [HW]
  float* spm = dsx + HW;            // [HW]
  float* spx = spm + HW;            // [HW]
  int* sp_idx = reinterpret_cast<int*>(spx + HW);   // [HW]
  float* gate_c = reinterpret_cast<float*>(sp_idx + HW);   // [C]
  float* dfc = gate_c + C;          // [C]  df / P
  float* dzs = dfc + C;             // [C]
  float* davg = dzs + C;            // [C]
  float* dmx = davg + C;            // [C]
  int* mx_idx = reinterpret_cast<int*>(dmx + C);    // [C]
  float* part = reinterpret_cast<float*>(mx_idx + C);   // [4][hidden]
  float* dh = part + 4 * hidden;    // [2][hidden]
  float* pdA = dh + 2 * hidden;     // [warps][HW]  per-warp partial sums of dA (S6)

  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const T* fb = fmap + (size_t)b * HW * C;
  const int CV = C >> 3;
  const float inv_hw = 1.0f / (float)HW, inv_c = 1.0f / (float)C;

  for (int p = tid; p < HW; p += blockDim.x) {
    a_lm[p] = heat ? heat[(size_t)b * HW + p] : 1.0f;
    gate_p[p] = use_spatial ? sv.gate_p[(size_t)b * HW + p] : 1.0f;
    sp_idx[p] = use_spatial ? sv.sp_idx[(size_t)b * HW + p] : -1;
    spm[p] = use_spatial ? sv.sp_mean[(size_t)b * HW + p] : 0.f;
    spx[p] = use_spatial ? sv.sp_max[(size_t)b * HW + p] : 0.f;
    dt[p] = dsm[p] = dsx[p] = 0.f;
  }
  for (int i = tid; i < nwarps * HW; i += blockDim.x) pdA[i] = 0.f;
  for (int c = tid; c < C; c += blockDim.x) {
    gate_c[c] = use_channel ? sv.gate_c[(size_t)b * C + c] : 1.0f;
    mx_idx[c] = use_channel ? sv.mx_idx[(size_t)b * C + c] : -1;
    dfc[c] = df[(size_t)b * C + c] * inv_hw;
    davg[c] = dmx[c] = 0.f;
  }
  __syncthreads();

  if (use_spatial) {
    // S1: d gp[p] = sum_c (df_c / P) * x2[p][c];  dt = dgp * gp (1 - gp)
    for (int p = warp; p < HW; p += nwarps) {
      const float a = a_lm[p];
      float s = 0.f;
      #pragma unroll 8
      for (int cv = lane; cv < CV; cv += 32) {
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(dfc[cv * 8 + e], (v[e] * a) * gate_c[cv * 8 + e], s);
      }
      s = warp_sum(s);
      if (lane == 0) {
        const float g = gate_p[p];
        dt[p] = s * g * (1.f - g);
      }
    }
    __syncthreads();
    // S2: transpose of the 7x7 conv -> d sm, d sx; conv weight gradient
    for (int q = tid; q < HW; q += blockDim.x) {
      const int qy = q / W, qx = q % W;
      float s0 = 0.f, s1 = 0.f;
      for (int ky = 0; ky < 7; ++ky) {
        const int py = qy - ky + 3;
        if (py < 0 || py >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int px = qx - kx + 3;
          if (px < 0 || px >= W) continue;
          const float d = dt[py * W + px];
          s0 = fmaf(d, sa_w[ky * 7 + kx], s0);
          s1 = fmaf(d, sa_w[49 + ky * 7 + kx], s1);
        }
      }
      dsm[q] = s0;
      dsx[q] = s1;
    }
    // conv weight gradient: one warp per tap, lanes over positions (98 threads walking all positions one after the other kept the
    // other 400 waiting for ~3 k dependent instructions)
    for (int t = warp; t < 98; t += nwarps) {
      const int ch = t / 49, ky = (t % 49) / 7, kx = t % 7;
      const float* in = ch ? spx : spm;
      float s = 0.f;
      for (int p = lane; p < HW; p += 32) {
        const int yy = p / W + ky - 3, xx = p % W + kx - 3;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        s = fmaf(dt[p], in[yy * W + xx], s);
      }
      s = warp_sum(s);
      if (lane == 0) atomicAdd(dsa_w + t, s);
    }
    __syncthreads();
  }

  if (use_channel) {
    // S3: d gc[c] = sum_p dx2[p][c] * x1[p][c];  dz = dgc * gc (1 - gc)
    // two position phases per 8-channel vector (even / odd positions; the odd half's partial sums pass through `davg`, which S5
    // overwrites afterwards): one thread per vector left half of the CTA idle with 28 KB in flight
    {
      constexpr int kPh = 2;
      const int nvec = blockDim.x / kPh, ph = tid / nvec, lv = tid % nvec;
      for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
        const int cv = cv0 + lv;
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        if (cv < CV) {
          #pragma unroll 8
          for (int p = ph; p < HW; p += kPh) {
            float v[8];
            load8(fb + (size_t)p * C + cv * 8, v);
            const float a = a_lm[p], g = gate_p[p], dm = dsm[p] * inv_c, dxv = dsx[p];
            const int si = sp_idx[p];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = cv * 8 + e;
              const float dx2 = fmaf(dfc[c], g, dm) + (c == si ? dxv : 0.f);
              s[e] = fmaf(dx2, v[e] * a, s[e]);
            }
          }
          if (ph > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) davg[cv * 8 + e] = s[e];
          }
        }
        __syncthreads();
        if (ph == 0 && cv < CV) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = cv * 8 + e;
            const float g = gate_c[c];
            const float v = (s[e] + davg[c]) * g * (1.f - g);
            dzs[c] = v;
            dz_out[(size_t)b * C + c] = v;
          }
        }
      }
    }
    __syncthreads();
    // S4: d(ha + hm)[j] = sum_c dz[c] * W2[c][j]
    const int nq = min(4, (int)blockDim.x / hidden);
    if (tid < nq * hidden) {
      const int j = tid % hidden, qd = tid / hidden;
      const int c0 = (int)((long long)C * qd / nq), c1 = (int)((long long)C * (qd + 1) / nq);
      float s = 0.f;
      #pragma unroll 32
      for (int c = c0; c < c1; ++c) s = fmaf(dzs[c], __ldg(w2 + (size_t)c * hidden + j), s);      // 32 rows in flight per thread (8: 14 KB per SM)
      part[qd * hidden + j] = s;
    }
    __syncthreads();
    if (tid < hidden) {
      float s = 0.f;
      for (int qd = 0; qd < nq; ++qd) s += part[qd * hidden + tid];
      const float ha = sv.hpre[((size_t)b * 2 + 0) * hidden + tid], hm = sv.hpre[((size_t)b * 2 + 1) * hidden + tid];
      const float da = ha > 0.f ? s : 0.f, dm = hm > 0.f ? s : 0.f;
      dh[tid] = da;
      dh[hidden + tid] = dm;
      dhpre_out[((size_t)b * 2 + 0) * hidden + tid] = da;
      dhpre_out[((size_t)b * 2 + 1) * hidden + tid] = dm;
    }
    __syncthreads();
    // S5: d avg[c], d max[c]
    for (int c = tid; c < C; c += blockDim.x) {
      float sa = 0.f, sx = 0.f;
      #pragma unroll 28
      for (int j = 0; j < hidden; ++j) {
        const float wv = __ldg(w1 + (size_t)j * C + c);
        sa = fmaf(dh[j], wv, sa);
        sx = fmaf(dh[hidden + j], wv, sx);
      }
      davg[c] = sa * inv_hw;
      dmx[c] = sx;
    }
    __syncthreads();
  }

  // S6: dx1 = dx2 * gc + davg / P + [p == argmax_c] dmax;  dx = dx1 * A;  dA[p] = sum_c dx1 * x
  // thread = (8-channel vector, position phase): the five per-channel terms live in registers for the whole walk over the
  // positions and only the per-position scalars come from shared memory.  (One warp per position read five shared arrays per
  // ELEMENT at an 8-word lane stride: 57 % of the kernel's stall samples sat on those two lines, ~140 of 247 us.)  The
  // per-position sum over channels goes through one shuffle reduction per warp into pdA[warp][p], summed in a fixed order.
  {
    constexpr int kPh = 2, kPB = 4;
    const int nvec = blockDim.x / kPh, ph = tid / nvec, lv = tid % nvec;      // nvec is a multiple of 32: a warp has one phase
    T* dxb = dx + (size_t)b * HW * C;
    for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
      const int cv = cv0 + lv;
      const bool on = cv < CV;
      float k_df[8], k_gc[8], k_da[8], k_dm[8];
      int k_mi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = on ? cv * 8 + e : 0;
        k_df[e] = dfc[c]; k_gc[e] = gate_c[c]; k_da[e] = davg[c]; k_dm[e] = dmx[c]; k_mi[e] = mx_idx[c];
      }
      for (int p0 = ph; p0 < HW; p0 += kPh * kPB) {
        float v[kPB][8];
#pragma unroll
        for (int u = 0; u < kPB; ++u) {
          const int p = p0 + u * kPh;
          if (on && p < HW) load8(fb + (size_t)p * C + cv * 8, v[u]);
        }
#pragma unroll
        for (int u = 0; u < kPB; ++u) {
          const int p = p0 + u * kPh;
          if (p < HW) {                      // warp-uniform
            const float a = a_lm[p], g = gate_p[p], dm = dsm[p] * inv_c, dxv = dsx[p];
            const int si = sp_idx[p];
            float sA = 0.f;
            if (on) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = cv * 8 + e;
                const float dx2 = fmaf(k_df[e], g, dm) + (c == si ? dxv : 0.f);
                const float dx1 = fmaf(dx2, k_gc[e], k_da[e]) + (k_mi[e] == p ? k_dm[e] : 0.f);
                sA = fmaf(dx1, v[u][e], sA);
                o[e] = dx1 * a;
              }
              store8(dxb + (size_t)p * C + cv * 8, o);
            }
            sA = warp_sum(sA);
            if (lane == 0) pdA[warp * HW + p] += sA;
          }
        }
      }
    }
    __syncthreads();
    if (dA) {
      for (int p = tid; p < HW; p += blockDim.x) {
        float sA = 0.f;
        for (int w = 0; w < nwarps; ++w) sA += pdA[w * HW + p];
        dA[(size_t)b * HW + p] = sA;
      }
    }
  }
}

// Channel-attention weight gradients, reduced over the batch.  thread = (channel, group of kCaJ hidden units);
// grid = (C / 128, hidden / kCaJ).  (One thread per channel walking all hidden x B terms took 577 us.)
//   dW1[j][c] = sum_b dha[b][j] avg[b][c] + dhm[b][j] max[b][c];   dW2[c][j] = sum_b dz[b][c] (relu(ha) + relu(hm))[b][j]
constexpr int kCaJ = 4;
__global__ void __launch_bounds__(128) ca_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ dhpre,
                                                      const float* __restrict__ hpre, const float* __restrict__ avg,
                                                      const float* __restrict__ mx, float* __restrict__ dw1,
                                                      float* __restrict__ dw2, int B, int C, int hidden) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * kCaJ;
  if (c >= C) return;
  float s1[kCaJ], s2[kCaJ];
#pragma unroll
  for (int u = 0; u < kCaJ; ++u) { s1[u] = 0.f; s2[u] = 0.f; }
#pragma unroll 4
  for (int b = 0; b < B; ++b) {
    const float av = avg[(size_t)b * C + c], mv = mx[(size_t)b * C + c], zv = dz[(size_t)b * C + c];
#pragma unroll
    for (int u = 0; u < kCaJ; ++u) {
      const int j = min(j0 + u, hidden - 1);
      const float ha = __ldg(hpre + ((size_t)b * 2 + 0) * hidden + j), hm = __ldg(hpre + ((size_t)b * 2 + 1) * hidden + j);
      const float da = __ldg(dhpre + ((size_t)b * 2 + 0) * hidden + j), dm = __ldg(dhpre + ((size_t)b * 2 + 1) * hidden + j);
      s1[u] = fmaf(da, av, s1[u]);
      s1[u] = fmaf(dm, mv, s1[u]);
      s2[u] = fmaf(zv, fmaxf(ha, 0.f) + fmaxf(hm, 0.f), s2[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < kCaJ; ++u) {
    const int j = j0 + u;
    if (j < hidden) {
      dw1[(size_t)j * C + c] = s1[u];
      dw2[(size_t)c * hidden + j] = s2[u];
    }
  }
}

// Landmark heat-map backward -> d attention_weights[5].  Single CTA.
//   N = R / (max_g + 1e-8); A = clamp(N, 0.1, 1);  R = sum_i w_i g_i
__global__ void __launch_bounds__(256) heat_bwd_kernel(const float* __restrict__ lm, const float* __restrict__ w5,
                                                      const float* __restrict__ raw, const uint32_t* __restrict__ gmax,
                                                      const float* __restrict__ dA, float* __restrict__ dw5, int B, int H,
                                                      int W, float sx, float sy, float denom, int group) {
  pdl_prologue();
  __shared__ float red[8][6];
  __shared__ float s_dmax, s_ties;
  const int HW = H * W, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_groups = (B + group - 1) / group;
  float dw[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int gidx = 0; gidx < n_groups; ++gidx) {
    uint32_t key = gmax[gidx];
    key = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    const float mx = __uint_as_float(key);
    const float den = __fadd_rn(mx, 1e-8f);
    const int b0 = gidx * group, b1 = min(B, b0 + group);
    // pass A: d max and the number of maximal elements
    float dmax = 0.f, ties = 0.f;
    for (int i = b0 * HW + tid; i < b1 * HW; i += blockDim.x) {
      const float r = raw[i];
      const float n = __fdiv_rn(r, den);
      const float dn = (n >= 0.1f && n <= 1.0f) ? dA[i] : 0.f;
      dmax -= dn * r / (den * den);
      if (r == mx) ties += 1.f;
    }
    dmax = warp_sum(dmax);
    ties = warp_sum(ties);
    if (lane == 0) { red[warp][0] = dmax; red[warp][1] = ties; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[w][0]; t += red[w][1]; }
      s_dmax = a;
      s_ties = t > 0.f ? t : 1.f;
    }
    __syncthreads();
    const float dmax_each = s_dmax / s_ties;
    // pass B: dR and the weight gradient
    for (int i = b0 * HW + tid; i < b1 * HW; i += blockDim.x) {
      const int b = i / HW, pos = i % HW;
      const float yv = (float)(pos / W), xv = (float)(pos % W);
      const float r = raw[i];
      const float n = __fdiv_rn(r, den);
      const float dn = (n >= 0.1f && n <= 1.0f) ? dA[i] : 0.f;
      float dR = dn / den;
      if (r == mx) dR += dmax_each;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float lx = __fmul_rn(lm[(b * 5 + k) * 2 + 0], sx), ly = __fmul_rn(lm[(b * 5 + k) * 2 + 1], sy);
        const float ddx = __fsub_rn(xv, lx), ddy = __fsub_rn(yv, ly);
        const float d2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
        dw[k] = fmaf(dR, expf(__fdiv_rn(-d2, denom)), dw[k]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) dw[k] = warp_sum(dw[k]);
  if (lane == 0)
    for (int k = 0; k < 5; ++k) red[warp][k] = dw[k];
  __syncthreads();
  if (tid < 5) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][tid];
    dw5[tid] = s;
  }
}

static size_t attn_saved_floats(int B, int HW, int C, int hidden) {
  return (size_t)B * (4 * (size_t)C + 2 * (size_t)hidden + 4 * (size_t)HW);
}
static void attn_carve(AttnSaved* sv, float* base, int B, int HW, int C, int hidden) {
  float* p = base;
  sv->avg = p; p += (size_t)B * C;
  sv->mx = p; p += (size_t)B * C;
  sv->mx_idx = reinterpret_cast<int*>(p); p += (size_t)B * C;
  sv->gate_c = p; p += (size_t)B * C;
  sv->hpre = p; p += (size_t)B * 2 * hidden;
  sv->sp_mean = p; p += (size_t)B * HW;
  sv->sp_max = p; p += (size_t)B * HW;
  sv->sp_idx = reinterpret_cast<int*>(p); p += (size_t)B * HW;
  sv->gate_p = p;
}

}  // namespace dfv

using namespace dfv;

extern "C" {

size_t dfv_attention_saved_floats(int B, int H, int W, int C, int hidden) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || hidden < 0) return 0;
  return attn_saved_floats(B, H * W, C, hidden);
}

/* Train-mode forward: as dfv_hybrid_attention_fwd, with native torch weight layouts (ca_w2 = fc.2.weight
 * [C][hidden]) and `saved` (fp32, dfv_attention_saved_floats()) receiving what backward needs. */
int dfv_hybrid_attention_train_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2,
                                   const float* sa_w, float* features, float* saved, int dtype, int B, int H, int W, int C,
                                   int hidden, int use_channel, int use_spatial, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(fmap && features && saved, "dfv_hybrid_attention_train_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_hybrid_attention_train_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(!use_channel || (ca_w1 && ca_w2 && hidden > 0), "dfv_hybrid_attention_train_fwd: channel attention needs weights");
  DFV_REQUIRE(!use_spatial || sa_w, "dfv_hybrid_attention_train_fwd: spatial attention needs weights");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dfv_hybrid_attention_train_fwd: bad shape (C %% 8 == 0)");
  if (!use_channel) hidden = 0;
  const size_t smem = sizeof(float) * ((size_t)4 * H * W + 3 * (size_t)C + (size_t)hidden + (size_t)(kTrPhases - 1) * 3 * C);
  DFV_REQUIRE(smem <= 200 * 1024, "dfv_hybrid_attention_train_fwd: map too large for one CTA");
  cudaStream_t st = as_stream(stream);
  AttnSaved sv;
  attn_carve(&sv, saved, B, H * W, C, hidden);
  ProfScope prof(PK_ATTENTION, 2.0 * (double)B * H * W * C * dtype_size(dtype) + 4.0 * B * C, 8.0 * (double)B * H * W * C, st);
  if (dtype == DFV_BF16) {
    auto k = hybrid_attention_train_fwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 1024, smem, st, (const __nv_bfloat16*)fmap, heat, ca_w1, ca_w2, sa_w, features, sv, H, W, C, hidden, use_channel, use_spatial);
  } else {
    auto k = hybrid_attention_train_fwd_kernel<float>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 1024, smem, st, (const float*)fmap, heat, ca_w1, ca_w2, sa_w, features, sv, H, W, C, hidden, use_channel, use_spatial);
  }
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Backward of the above.  dfeatures: [B][C] fp32.  Outputs: dfmap [B][H*W][C] (activation dtype); dheat fp32
 * [B][H*W] (if heat != NULL); parameter gradients in torch layout: dca_w1 [hidden][C], dca_w2 [C][hidden]
 * (overwritten), dsa_w [2][7][7] (overwritten).  ws: fp32 [B*C + B*2*hidden]. */
int dfv_hybrid_attention_bwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2,
                             const float* sa_w, const float* dfeatures, const float* saved, void* dfmap, float* dheat,
                             float* dca_w1, float* dca_w2, float* dsa_w, float* ws, int dtype, int B, int H, int W, int C,
                             int hidden, int use_channel, int use_spatial, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(fmap && dfeatures && saved && dfmap && ws, "dfv_hybrid_attention_bwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_hybrid_attention_bwd: bad dtype %d", dtype);
  DFV_REQUIRE(!use_channel || (ca_w1 && ca_w2 && dca_w1 && dca_w2 && hidden > 0 && hidden <= 128),
              "dfv_hybrid_attention_bwd: channel attention needs weights / gradient buffers (hidden <= 128)");
  DFV_REQUIRE(!use_spatial || (sa_w && dsa_w), "dfv_hybrid_attention_bwd: spatial attention needs weights / gradient buffer");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dfv_hybrid_attention_bwd: bad shape");
  DFV_REQUIRE(!heat || dheat, "dfv_hybrid_attention_bwd: dheat missing");
  if (!use_channel) hidden = 0;
  const int HW = H * W;
  const size_t smem = sizeof(float) * ((size_t)(8 + 512 / 32) * HW + 6 * (size_t)C + 6 * (size_t)hidden);
  DFV_REQUIRE(smem <= 200 * 1024, "dfv_hybrid_attention_bwd: map too large for one CTA");
  cudaStream_t st = as_stream(stream);
  AttnSaved sv;
  attn_carve(&sv, const_cast<float*>(saved), B, HW, C, hidden);
  float* dz = ws;
  float* dhpre = ws + (size_t)B * C;
  if (use_spatial) DFV_CUDA(cudaMemsetAsync(dsa_w, 0, 98 * sizeof(float), st));
  ProfScope prof(PK_ATTENTION, 4.0 * (double)B * HW * C * dtype_size(dtype), 16.0 * (double)B * HW * C, st);
  if (dtype == DFV_BF16) {
    auto k = hybrid_attention_bwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 512, smem, st, (const __nv_bfloat16*)fmap, heat, ca_w1, ca_w2, sa_w, dfeatures, sv, (__nv_bfloat16*)dfmap, dheat, dz, dhpre,
                            dsa_w, H, W, C, hidden, use_channel, use_spatial);
  } else {
    auto k = hybrid_attention_bwd_kernel<float>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 512, smem, st, (const float*)fmap, heat, ca_w1, ca_w2, sa_w, dfeatures, sv, (float*)dfmap, dheat, dz, dhpre, dsa_w, H, W, C,
                            hidden, use_channel, use_spatial);
  }
  DFV_LAUNCH_CHECK();
  if (use_channel) {
    DFV_PDL((ca_wgrad_kernel), dim3((unsigned)((C + 127) / 128), (unsigned)((hidden + kCaJ - 1) / kCaJ)), 128, 0, st, dz, dhpre, sv.hpre, sv.avg, sv.mx,
            dca_w1, dca_w2, B, C, hidden);
    DFV_LAUNCH_CHECK();
  }
  return DFV_OK;
}

/* Landmark heat-map backward: d attention_weights[5] from dheat (gradient wrt the clamped map).
 * raw_ws / max_ws are the scratch buffers dfv_landmark_heatmap_fwd filled. */
int dfv_landmark_heatmap_bwd(const float* landmarks, const float* weights5, const float* raw_ws, const uint32_t* max_ws,
                             const float* dheat, float* dweights5, int B, int H, int W, float ref_size, float sigma,
                             int group, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(landmarks && weights5 && raw_ws && max_ws && dheat && dweights5, "dfv_landmark_heatmap_bwd: null pointer");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && ref_size > 0.f && sigma > 0.f, "dfv_landmark_heatmap_bwd: bad shape");
  if (group <= 0 || group > B) group = B;
  const float sx = (float)((double)W / (double)ref_size), sy = (float)((double)H / (double)ref_size);
  const float denom = (float)(2.0 * (double)sigma * (double)sigma);
  DFV_PDL((heat_bwd_kernel), 1, 256, 0, as_stream(stream), landmarks, weights5, raw_ws, max_ws, dheat, dweights5, B, H, W, sx, sy, denom, group);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
