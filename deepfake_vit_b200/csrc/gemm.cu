// Pointwise (1x1) convolution as a GEMM with fused epilogue:
//   out[M][N] = act((a[M][K] * gate[image][K]) . w[N][K]^T + bias[N]) + residual[M][N]
//
// bf16 path (pw_gemm_tc_kernel): persistent, warp-specialised tcgen05 kernel
//   warp 0      TMA producer: A tile 128 x 64 and B tile BN x 64 per k-block, 128B-swizzled
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16, fp32 accumulators
//               in TMEM (2 stages x 256 columns), owns TMEM alloc/dealloc
//   warps 2-5   SE-gate transform (project GEMMs only): multiply the landed A tile in shared
//               memory by the per-image channel gate, fence.proxy.async, hand to the MMA warp.
//               (x (.) s) . W is thereby applied on the operand path; the gated tensor is
//               never written to HBM.
//   warps 6-13  epilogue: tcgen05.ld -> +bias -> swish -> +residual -> bf16 -> global
// Most of these GEMMs are HBM-bound (K/N of 24..960: AI < 218 flop/B); only the 12x12-stage
// layers (K or N >= 1632, head 448x1792) are tensor-bound.
//
// fp32 path / debug (pw_gemm_simt_kernel): plain tiled FFMA kernel, exact fp32 arithmetic for
// the 1e-4 parity mode.
#include "common.cuh"

namespace dfv {

// ------------------------------------------------------------------------------------
// SIMT kernel
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float v) { d = __float2bfloat16_rn(v); }

template <typename T, bool kFast>
__global__ void __launch_bounds__(256) pw_gemm_simt_kernel(const T* __restrict__ a, const T* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          const T* __restrict__ a_scale, int rows_per_image,
                                                          const T* __restrict__ residual, T* __restrict__ out,
                                                          long long M, int K, int N, int act) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx / TK, kk = idx % TK;
      const long long m = m0 + r;
      const int k = k0 + kk;
      float va = 0.f, vw = 0.f;
      if (m < M && k < K) {
        va = to_f(a[(size_t)m * K + k]);
        if (a_scale) {
          va *= to_f(a_scale[(size_t)(m / rows_per_image) * K + k]);
          if constexpr (sizeof(T) == 2) va = __bfloat162float(__float2bfloat16_rn(va));  // as the tc path rounds
        }
      }
      if (n0 + r < N && k < K) vw = to_f(w[(size_t)(n0 + r) * K + k]);
      As[kk][r] = va;
      Ws[kk][r] = vw;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + bias[n];
      if (act == DFV_ACT_SILU) v = silu<kFast>(v);
      if (residual) v += to_f(residual[(size_t)m * N + n]);
      from_f(out[(size_t)m * N + n], v);
    }
  }
}

// ------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------
constexpr int kBM = 128;        // UMMA M (cta_group::1)
constexpr int kBK = 64;         // one 128-byte swizzle atom of bf16 along K
constexpr int kMaxStages = 12;
constexpr int kTcThreads = 576; // 18 warps
constexpr int kFirstWorker = 2; // warps 2..17: 16 workers
constexpr int kNumWorkers = 16;
constexpr uint32_t kTmemCols = 512;

struct TcParams {
  long long M;
  int K, N;
  int cw;             // output columns per epilogue warp (= BN / column parts), = nb * bw
  int nb, bw;         // TMA-store boxes per warp and their width (16 / 32 / 48 / 64 columns)
  int swz;            // XOR mask source of the staging swizzle: 3 = 128B, 2 = 64B, 1 = 32B, 0 = none
  int nbuf;           // output staging buffers per warp (1 or 2)
  int BN;             // N tile (multiple of 16, <= 256)
  int n_tiles_n;
  long long n_tiles;  // total tiles
  long long tiles_per_cta;   // contiguous tile range per CTA
  int b_res;          // 1: weight-stationary -- the CTA keeps ONE N tile of the weight (all of K) in shared memory
  int m_splits;       //    and walks tiles_per_cta consecutive M tiles; blockIdx = n_tile * m_splits + m_split
  long long n_tiles_m;
  int cl;             // streaming plans: 2 = CTA pair (cta_group::2).  The two CTAs of a cluster take consecutive M tiles of
                      // the SAME N tile; each loads its A tile and HALF of the weight tile's rows, the leader issues M = 256 MMAs
  int nsub;           // streaming plans with exactly two N tiles: 2 = both N tiles of an M tile share ONE A stage -- the stage carries A and the
                      // weight rows of both tiles, the MMA thread accumulates into the two TMEM buffers side by side, the A operand (and its
                      // gate transform) is fetched once instead of once per N tile.  The price: no accumulator double-buffering across tiles
  int rotate;         // streaming plans: 1 = every cluster starts its walk over the N tiles at a different tile (see tile_n)
  int k_blocks;
  int stages;
  int rows_per_image;
  int act;
};

struct __align__(8) TcBarriers {
  uint64_t full[kMaxStages];    // TMA landed
  uint64_t ready[kMaxStages];   // SE transform done
  uint64_t empty[kMaxStages];   // MMAs that read the stage completed
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t b_full;              // weight-stationary mode: the resident weight tile landed
  uint64_t res_full[16][2];     // per epilogue warp and staging buffer: the residual box landed (TMA)
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: 8-row core-matrix groups are 1024 B apart (SBO); LBO unused.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// Worker warps 2..17.  With an SE gate (project GEMMs) warps 2..9 are the operand transform
// (two groups of four alternating k-block stages) and warps 10..17 the epilogue; without one
// (expand / head GEMMs) all sixteen are epilogue warps.
// kPair: the CTA-pair (cta_group::2) variant -- a separate instantiation, because a kernel that contains cta_group::2
// instructions can only be launched as a cluster of two (a plain launch fails with "cluster misconfiguration").
template <bool kHasScale, int kAct, bool kRes, bool kPair, bool kDual = false>
__global__ void __launch_bounds__(kTcThreads, 1)
    pw_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                      const float* __restrict__ bias, const __nv_bfloat16* __restrict__ a_scale,
                      TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; do not rely on the dynamic-smem base
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool b_res = p.b_res != 0;
  const uint32_t a_bytes = kBM * kBK * 2;
  constexpr bool pair = kPair;                                             // CTA pair: this CTA stages half of the weight tile's rows
  const uint32_t b_bytes = (uint32_t)(pair ? p.BN / 2 : p.BN) * kBK * 2;
  constexpr bool dual = kDual;                                             // both N tiles of an M tile ride on one A stage (p.nsub == 2)
  const uint32_t stage_bytes = b_res ? a_bytes : a_bytes + (dual ? 2 : 1) * b_bytes;        // weight-stationary: stages carry A only
  unsigned char* tiles = smem;
  constexpr int kNumEpiW = kNumWorkers - (kHasScale ? 8 : 0);
  unsigned char* bres = tiles + (size_t)p.stages * stage_bytes;              // [k_blocks][BN x 64] resident weight tile
  unsigned char* staging = bres + (b_res ? (size_t)p.k_blocks * b_bytes : 0);   // [epilogue warp][nbuf][nb][32 rows x bw]
  const uint32_t warp_stage_bytes = (uint32_t)p.cw * 64;                     // 32 rows x cw columns x 2 B
  float* bias_sm = reinterpret_cast<float*>(staging + (size_t)kNumEpiW * p.nbuf * warp_stage_bytes);
  const int n_pad = p.n_tiles_n * p.BN;
  __nv_bfloat16* gate_sm = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(bias_sm) + (size_t)((n_pad * 4 + 15) / 16) * 16);
  const int k_pad = p.k_blocks * kBK;      // gate rows of the (at most two) images a tile touches: [2][k_pad]
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(reinterpret_cast<unsigned char*>(gate_sm) + (kHasScale ? (size_t)2 * k_pad * 2 : 0));
  // contiguous tile range per CTA (m-major): consecutive tiles stay inside one image for rows_per_image / 128 tiles.
  // Weight-stationary mode: one N tile, tiles_per_cta consecutive M tiles.
  // CTA-pair plans: the tile index counts PAIR tiles (two consecutive M tiles x one N tile), both CTAs walk the same
  // range, CTA rank r takes M tile (pair tile row) * 2 + r.
  const int cl = pair ? 2 : 1;
  const uint32_t crank = pair ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const long long t_begin = b_res ? (long long)(blockIdx.x % p.m_splits) * p.tiles_per_cta : (long long)(blockIdx.x / cl) * p.tiles_per_cta;
  const long long t_end = min(t_begin + p.tiles_per_cta, b_res ? p.n_tiles_m : p.n_tiles);
  const int nt_res = b_res ? (int)(blockIdx.x / p.m_splits) : 0;
  auto tile_m = [&](long long t) -> long long { return b_res ? t : (t / p.n_tiles_n) * cl + crank; };
  // Streaming plans walk the N tiles of an M tile in an order ROTATED by the M tile's index: CTAs run in near lockstep on
  // different M tiles, and without the rotation all of them fetch the same weight tile (the same few hundred L2 lines, i.e.
  // the same few L2 slices) at the same moment while the other slices idle.  The rotation is a function of the M tile only, so
  // the (M tile, N tile) pairs stay a bijection however the tile range is cut into CTAs.
  const bool rotate = !b_res && !dual && p.rotate != 0 && p.n_tiles_n > 1;
  auto tile_n = [&](long long t) -> int {
    if (b_res) return nt_res;
    const long long row = t / p.n_tiles_n;
    const int n0 = (int)(t - row * p.n_tiles_n);
    if (!rotate) return n0;
    const int n = n0 + (int)((uint32_t)row % (uint32_t)p.n_tiles_n);
    return n >= p.n_tiles_n ? n - p.n_tiles_n : n;
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kNumXform = kHasScale ? 8 : 0;
  constexpr int kNumEpi = kNumWorkers - kNumXform;

  // swish epilogue works on h = x / 2: keep bias / 2
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) bias_sm[i] = i < p.N ? (kAct == DFV_ACT_SILU ? 0.5f * bias[i] : bias[i]) : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->ready[s], pair ? 8 : 4);     // pair: the leader's MMA thread waits for both CTAs' transform groups
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], pair ? 2 * kNumEpi : kNumEpi);     // pair: both CTAs' epilogue warps release the leader
    }
    mbar_init(&bars->b_full, 1);
    for (int e = 0; e < 16; ++e) { mbar_init(&bars->res_full[e][0], 1); mbar_init(&bars->res_full[e][1], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (pair) tmem_alloc_pair(&bars->tmem_base, kTmemCols); else tmem_alloc(&bars->tmem_base, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();     // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (b_res && t_begin < t_end) {
        mbar_expect_tx(&bars->b_full, (uint32_t)p.k_blocks * b_bytes);
        for (int kb = 0; kb < p.k_blocks; ++kb) tma_load_2d(bres + (size_t)kb * b_bytes, &tm_b, &bars->b_full, kb * kBK, nt_res * p.BN);
      }
      for (long long t = t_begin; t < t_end; ++t) {
        const long long mt = tile_m(t);
        const int nt = tile_n(t);
        if (dual && nt != 0) continue;      // its operands arrived with the first N tile of this M tile
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1, 1);
          unsigned char* sa = tiles + (size_t)stage * stage_bytes;
          if (pair) {
            // this CTA's A tile and its half of the weight tile's rows (tm_b's box is BN / 2 rows).  Gated: the bytes are
            // counted on THIS CTA's barrier (its transform warps wait there, then arrive at the leader); ungated: on the
            // LEADER's barrier, which expects both CTAs' bytes and releases the MMA thread directly.
            const int brow = nt * p.BN + (int)crank * (p.BN / 2);
            if (kHasScale) {
              mbar_expect_tx(&bars->full[stage], stage_bytes);
              tma_load_2d(sa, &tm_a, &bars->full[stage], kb * kBK, (int)(mt * kBM));
              tma_load_2d(sa + a_bytes, &tm_b, &bars->full[stage], kb * kBK, brow);
              if (dual) tma_load_2d(sa + a_bytes + b_bytes, &tm_b, &bars->full[stage], kb * kBK, brow + p.BN);
            } else {
              if (leader) mbar_expect_tx(&bars->full[stage], 2 * stage_bytes);
              const uint32_t lbar = mapa_shared(&bars->full[stage], 0);
              tma_load_2d_pair(sa, &tm_a, lbar, kb * kBK, (int)(mt * kBM));
              tma_load_2d_pair(sa + a_bytes, &tm_b, lbar, kb * kBK, brow);
              if (dual) tma_load_2d_pair(sa + a_bytes + b_bytes, &tm_b, lbar, kb * kBK, brow + p.BN);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&bars->full[stage], stage_bytes);
          tma_load_2d(sa, &tm_a, &bars->full[stage], kb * kBK, (int)(mt * kBM));
          if (!b_res) tma_load_2d(sa + a_bytes, &tm_b, &bars->full[stage], kb * kBK, nt * p.BN);
          if (dual) tma_load_2d(sa + a_bytes + b_bytes, &tm_b, &bars->full[stage], kb * kBK, (nt + 1) * p.BN);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0 && leader) {      // CTA pair: the leader issues for both SMs
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128 (256 across a CTA pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((pair ? 2 * kBM : kBM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (b_res && t_begin < t_end) mbar_wait(&bars->b_full, 0, 6);
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (dual && as != 0) continue;      // second N tile of the M tile: accumulated together with the first (a CTA's range holds whole M tiles)
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1, 2);
        if (dual) mbar_wait(&bars->tmem_empty[1], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(kHasScale ? &bars->ready[stage] : &bars->full[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(b_res ? smem_u32(bres + (size_t)kb * b_bytes) : sa + a_bytes);
          const uint64_t db2 = make_sw128_desc(sa + a_bytes + b_bytes);
          const int k_left = p.K - kb * kBK;
          const int ksteps = k_left >= kBK ? kBK / 16 : (k_left + 15) / 16;
          if (pair) {
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              if (dual) umma_bf16_pair(d_tmem + 256, da + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit_pair(&bars->empty[stage]);      // frees the stage in both CTAs
          } else {
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              if (dual) umma_bf16(d_tmem + 256, da + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit(&bars->empty[stage]);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (pair) umma_commit_pair(&bars->tmem_full[as]); else umma_commit(&bars->tmem_full[as]);
        if (dual) { if (pair) umma_commit_pair(&bars->tmem_full[1]); else umma_commit(&bars->tmem_full[1]); }
      }
    }
  } else if (kHasScale && warp < kFirstWorker + kNumXform) {
    // ------------------------------------------------------------- SE-gate transform
    // Two groups of four warps take the even / the odd pipeline STAGES, so two stages are in transform at any time: one stage is a serial chain (barrier wait -> shared loads -> multiply ->
    // stores -> proxy fence -> arrival) of several hundred cycles, and with all eight warps on every stage that chain
    // -- not HBM -- paced the gated GEMMs.  Thread = one 128-byte tile row (eight 16-byte chunks, two halves of four).
    // The split is by stage index, not by k-block counter: a group then sees EVERY phase of the barriers it waits on.
    // (Alternating by counter with an odd stage count lets the other group consume the intermediate phase of a barrier;
    // a parity wait two phases behind passes at once on the stale phase -- seen as a rare wrong tile.)
    const int grp = (warp - kFirstWorker) >> 2;           // 0 / 1
    const int tx = threadIdx.x - kFirstWorker * 32;       // 0..255: gate staging is done by all transform threads together
    const int row = tx & 127;
    const bool smem_gate = p.rows_per_image >= kBM;       // a tile then touches at most two images
    const long long n_images = p.M / p.rows_per_image;
    const uint32_t tiles_s = smem_u32(tiles), gate_s = smem_u32(gate_sm);
    const uint32_t row_s = (uint32_t)row * 128u;
    uint32_t sw[8];                                       // swizzled chunk offsets of this row (tile-invariant)
#pragma unroll
    for (int c = 0; c < 8; ++c) sw[c] = (uint32_t)((c ^ (row & 7)) << 4);
    long long cached_img = -1;
    long long img_lo = t_begin < t_end ? (tile_m(t_begin) * kBM) / p.rows_per_image : 0;    // one division per CTA
    int stage = 0;                                        // stage / phase of the running k-block (all tiles)
    uint32_t phase = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      if (dual && tile_n(t) != 0) continue;                      // the A stages of this M tile were transformed with its first N tile
      const long long m0 = tile_m(t) * kBM;
      while (m0 >= (img_lo + 1) * p.rows_per_image) ++img_lo;   // tiles are visited in increasing m: at most a step or two
      const long long m = m0 + row;
      const bool valid = m < p.M;
      const int second = (smem_gate && m >= (img_lo + 1) * p.rows_per_image) ? 1 : 0;      // this row lies in image img_lo + 1
      if (smem_gate && img_lo != cached_img) {
        // stage the gate rows of images img_lo, img_lo + 1 (bf16 [2][k_pad]); the barrier before keeps
        // slower transform threads of the previous tile from reading rows that are being replaced
        asm volatile("bar.sync 3, 256;" ::: "memory");
        for (int i = tx; i < 2 * (p.K >> 3); i += 256) {
          const int r = i / (p.K >> 3), c = i % (p.K >> 3);
          if (img_lo + r < n_images)
            *reinterpret_cast<uint4*>(gate_sm + (size_t)r * k_pad + c * 8) =
                __ldg(reinterpret_cast<const uint4*>(a_scale + (size_t)(img_lo + r) * p.K + c * 8));
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        cached_img = img_lo;
      }
      const uint32_t grow_s = gate_s + (uint32_t)second * (uint32_t)k_pad * 2u;
      const __nv_bfloat16* grow_g = smem_gate ? nullptr : a_scale + (size_t)(valid ? m / p.rows_per_image : 0) * p.K;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        if ((stage & 1) != grp) {       // the other group's stage
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_wait(&bars->full[stage], phase, 4);
        if (valid) {
          const uint32_t arow = tiles_s + (uint32_t)stage * stage_bytes + row_s;
          const int k0 = kb * kBK;
          const int nchunk = min(8, (p.K - k0) >> 3);
          // x * gate in packed bf16 (HMUL2.BF16): both operands are bf16, as in the autocast
          // reference where sigmoid(se) is a bf16 tensor; one rounding, no unpack / repack.
          // All loads of a half first: a store between them would serialise (swizzled addresses alias).
          if (nchunk == 8 && smem_gate) {
            // the common case, branch-free: a full 64-column k-block, gate rows in shared memory
            const uint32_t gk = grow_s + (uint32_t)k0 * 2u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 u[4], gt[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u[i].x), "=r"(u[i].y), "=r"(u[i].z), "=r"(u[i].w) : "r"(arow + sw[h * 4 + i]));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(gt[i].x), "=r"(gt[i].y), "=r"(gt[i].z), "=r"(gt[i].w) : "r"(gk + (uint32_t)(h * 4 + i) * 16u));
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t vx = hmul2_bf16(u[i].x, gt[i].x), vy = hmul2_bf16(u[i].y, gt[i].y);
                const uint32_t vz = hmul2_bf16(u[i].z, gt[i].z), vw = hmul2_bf16(u[i].w, gt[i].w);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(arow + sw[h * 4 + i]), "r"(vx), "r"(vy), "r"(vz), "r"(vw) : "memory");
              }
            }
          } else {
            for (int c = 0; c < nchunk; ++c) {      // K tail / tiny images: one chunk at a time
              uint4 u, gt;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(arow + (uint32_t)((c ^ (row & 7)) << 4)));
              if (smem_gate)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(gt.x), "=r"(gt.y), "=r"(gt.z), "=r"(gt.w) : "r"(grow_s + (uint32_t)(k0 + c * 8) * 2u));
              else
                gt = __ldg(reinterpret_cast<const uint4*>(grow_g + k0 + c * 8));
              const uint32_t vx = hmul2_bf16(u.x, gt.x), vy = hmul2_bf16(u.y, gt.y), vz = hmul2_bf16(u.z, gt.z), vw = hmul2_bf16(u.w, gt.w);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(arow + (uint32_t)((c ^ (row & 7)) << 4)), "r"(vx), "r"(vy), "r"(vz), "r"(vw) : "memory");
            }
          }
        }
        fence_proxy_async();      // every writer fences, then ONE arrival per warp
        __syncwarp();
        if (lane == 0) {
          if (pair) mbar_arrive_cluster(mapa_shared(&bars->ready[stage], 0));      // the leader's MMA thread waits for both CTAs
          else mbar_arrive(&bars->ready[stage]);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue
    // TMEM -> registers -> (+bias, swish, +residual, bf16) -> swizzled staging tile in shared memory
    // -> TMA store (clips the M / N tails; full-line writes instead of 32 strided 16-byte stores).
    // Every epilogue warp is autonomous: it owns 32 rows (its TMEM lane quarter) x cw columns of the tile, stages
    // them in its own shared-memory slice and issues its own TMA stores -- no CTA-wide barrier per tile (the
    // all-warps barrier + single-leader store version spent half its samples waiting at the barrier on the
    // large-M / small-N layers).
    const int ew = warp - kFirstWorker - kNumXform;  // 0 .. kNumEpi-1
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int part = ew >> 2;                        // column part
    const int col_lo = part * p.cw;                  // first tile column of this warp
    const int nchunk = p.cw >> 4;                    // 16-column chunks per warp (<= 8)
    const uint32_t pitch = (uint32_t)p.bw * 2;
    const uint32_t row_off = (uint32_t)lane * pitch;
    const uint32_t xr = p.swz == 3 ? (uint32_t)(lane & 7) : (p.swz == 2 ? (uint32_t)((lane >> 1) & 3) : (p.swz == 1 ? (uint32_t)((lane >> 2) & 1) : 0u));
    const uint32_t box_bytes = 32u * pitch;
    unsigned char* my_stage = staging + (size_t)ew * p.nbuf * warp_stage_bytes;
    // one 8-column group: +bias, swish, +residual, bf16, swizzled 16-byte store into the warp's staging slice.
    // Issue slots are what this epilogue runs out of on the write-dominated expand layers (6x more outputs than
    // inputs: ~11 outputs per clock per SM at the HBM roofline), so the arithmetic is packed fp32 pairs (FFMA2 /
    // FADD2: bit-identical to two scalar operations) and the staging address needs no division (a warp owns at most
    // two store boxes).
    auto emit8 = [&](const uint32_t* v, int wcol, int n, unsigned char* stg) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_sm + n);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_sm + n + 4);
      const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        if constexpr (kAct == DFV_ACT_SILU) {
          const float2 h = ffma2(a, make_float2(0.5f, 0.5f), bb[j]);
          float2 th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(h.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(h.y));
          o[j] = ffma2(h, th, h);
        } else {
          o[j] = fadd2(a, bb[j]);
        }
      }
      const uint32_t box = (uint32_t)wcol >= (uint32_t)p.bw ? 1u : 0u;
      const uint32_t c8 = ((uint32_t)wcol - (box ? (uint32_t)p.bw : 0u)) >> 3;
      unsigned char* dst = stg + box * box_bytes + row_off + ((c8 ^ xr) << 4);
      if constexpr (kRes) {
        // the residual box was loaded by TMA into this very slot (same box shape and swizzle as the store): add in place
        const uint4 rres = *reinterpret_cast<const uint4*>(dst);
        o[0] = fadd2(o[0], make_float2(bf16_lo(rres.x), bf16_hi(rres.x)));
        o[1] = fadd2(o[1], make_float2(bf16_lo(rres.y), bf16_hi(rres.y)));
        o[2] = fadd2(o[2], make_float2(bf16_lo(rres.z), bf16_hi(rres.z)));
        o[3] = fadd2(o[3], make_float2(bf16_lo(rres.w), bf16_hi(rres.w)));
      }
      uint4 pk;
      pk.x = pack_bf16(o[0].x, o[0].y); pk.y = pack_bf16(o[1].x, o[1].y);
      pk.z = pack_bf16(o[2].x, o[2].y); pk.w = pack_bf16(o[3].x, o[3].y);
      *reinterpret_cast<uint4*>(dst) = pk;
    };
    int it = 0;
    int n_stores = 0;     // tiles THIS warp stored: picks the staging buffer (a warp whose column part lies beyond N on
                          // some N tiles skips those, so the tile counter's parity would reuse a buffer still being read)
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const long long m0 = tile_m(t) * kBM + q * 32;           // first row of this warp
      const int nt = tile_n(t);
      const int n_lo = nt * p.BN + col_lo;                     // first output column of this warp
      const bool active = col_lo < p.BN && n_lo < p.N && m0 < p.M;
      unsigned char* stg = my_stage + (size_t)(p.nbuf == 2 ? (n_stores & 1) : 0) * warp_stage_bytes;
      const int buf = p.nbuf == 2 ? (n_stores & 1) : 0;
      // the staging slice we are about to overwrite must have been read by this warp's earlier TMA stores
      if (lane == 0) {
        if (p.nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if constexpr (kRes) {
          if (active) {      // residual boxes -> the staging slice, by TMA (rows / columns beyond the tensor arrive as zeros)
            uint64_t* rb = &bars->res_full[ew][buf];
            int nbx = 0;
            for (int bx = 0; bx < p.nb; ++bx) nbx += (n_lo + bx * p.bw < p.N) ? 1 : 0;
            mbar_expect_tx(rb, (uint32_t)nbx * box_bytes);
            for (int bx = 0; bx < nbx; ++bx) tma_load_2d(stg + (size_t)bx * box_bytes, &tm_res, rb, n_lo + bx * p.bw, (int)m0);
          }
        }
      }
      __syncwarp();
      mbar_wait(&bars->tmem_full[as], aphase, 5);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256 + (uint32_t)col_lo;
      if (active) {
        if constexpr (kRes) mbar_wait(&bars->res_full[ew][buf], (uint32_t)((p.nbuf == 2 ? (n_stores >> 1) : n_stores) & 1), 7);
        // 16-column chunks, software-pipelined over two register sets: the tcgen05.ld of chunk c+1 is in flight while
        // chunk c goes through bias / swish / pack / staging store (the load latency was exposed once per 32 columns)
        int ncv = (p.N - n_lo + 15) >> 4;            // chunks of this warp that hold real columns
        if (ncv > nchunk) ncv = nchunk;
        uint32_t v0[16], v1[16];
        __syncwarp();
        tmem_ld16(tbase, v0);
        for (int cc = 0; cc < ncv; cc += 2) {
          const int wcol = cc * 16, n0 = n_lo + wcol;
          tmem_ld_wait16(v0);
          if (cc + 1 < ncv) tmem_ld16(tbase + (uint32_t)wcol + 16, v1);
          emit8(v0, wcol, n0, stg);
          emit8(v0 + 8, wcol + 8, n0 + 8, stg);
          if (cc + 1 < ncv) {
            tmem_ld_wait16(v1);
            if (cc + 2 < ncv) tmem_ld16(tbase + (uint32_t)wcol + 32, v0);
            emit8(v1, wcol + 16, n0 + 16, stg);
            emit8(v1 + 8, wcol + 24, n0 + 24, stg);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                       // TMEM stage drained (pair: tell the leader's MMA thread)
        if (pair) mbar_arrive_cluster(mapa_shared(&bars->tmem_empty[as], 0)); else mbar_arrive(&bars->tmem_empty[as]);
      }
      fence_proxy_async();                                   // staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0 && active) {
        for (int bx = 0; bx < p.nb; ++bx) {
          const int n = n_lo + bx * p.bw;
          if (n >= p.N) break;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(smem_u32(stg + (size_t)bx * box_bytes)), "r"(n), "r"((int)m0)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (active) ++n_stores;
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();     // no CTA leaves (or frees TMEM) while the pair's MMAs or barrier signals may still touch it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (pair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Tile plan of one GEMM: N tile, streaming vs weight-stationary, pipeline depth, grid.
static int plan_tc(TcParams& p, size_t& smem, long long& grid, long long M, int K, int N, bool scaled, int rows_per_image, int act,
                   const dfv_gemm_tuning* tuning = nullptr, int max_clusters = 0) {
  p.M = M;
  p.K = K;
  p.N = N;
  // N tile: BN = (column parts) x cw, cw in {16, 32, 48, 64, 96, 128} so that every epilogue warp owns whole
  // TMA-store boxes.  Streaming cost model: padded columns plus a fixed per-tile charge (A is re-fetched per N tile).
  const int parts = scaled ? 2 : 4;
  static const int cws[6] = {16, 32, 48, 64, 96, 128};
  long long best_cost = -1, best_pad = 0;
  p.BN = 0;
  for (int ci = 0; ci < 6; ++ci) {
    const int bn = parts * cws[ci];
    if (bn > 256) continue;
    const long long nt = (N + bn - 1) / bn, cost = nt * (bn + 96), pad = nt * bn;
    // ties: a gated GEMM repeats the SE transform of A per N tile -> fewer, wider tiles (1632 -> 272: 2 x 192 beats
    // 3 x 96 by 17%); otherwise fewer padded columns
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && (scaled || pad < best_pad))) {
      best_cost = cost;
      best_pad = pad;
      p.BN = bn;
      p.cw = cws[ci];
    }
  }
  // Weight-stationary candidates: the CTA keeps one N tile of the weight (all of K) in shared memory and streams A
  // only.  The streaming scheme re-fetches the weight tile for every 128 rows, and on many of these layers that
  // L2 -> SM traffic (not HBM) sets the pace: e.g. 272 -> 1632 at 12x12 moved 496 MB through the crossbar for 21 MB of
  // A and 0.9 MB of weight.  Estimated crossbar bytes decide; a caller-supplied dfv_gemm_tuning restricts the choice.
  const long long n_tm = (M + kBM - 1) / kBM;
  const int k_blocks = (K + kBK - 1) / kBK;
  const size_t budget = 226 * 1024;   // of the 227 KB a CTA may opt in to
  auto tail_bytes = [&](int bn) {
    const int ntn = (N + bn - 1) / bn;
    return align_up((size_t)ntn * bn * 4, 16) + (scaled ? (size_t)4 * k_blocks * kBK : 0) + sizeof(TcBarriers) + 64 + 1024;
  };
  p.b_res = 0;
  p.m_splits = 1;
  p.tiles_per_cta = 1;
  {
    const int force_res = tuning ? tuning->weight_stationary : -1, force_bn = tuning ? tuning->bn : 0;
    if (force_res == 0 && force_bn > 0) {
      for (int ci = 0; ci < 6; ++ci)
        if (parts * cws[ci] == force_bn) { p.BN = force_bn; p.cw = cws[ci]; }
    }
    const int ntn0 = (N + p.BN - 1) / p.BN;
    const double stream_bytes = (double)n_tm * ((double)ntn0 * kBM * K * 2 + (double)N * K * 2);
    // must be clearly better than streaming; a gated GEMM is charged its padded columns on both sides (it gains the
    // deeper A ring as well as the weight traffic)
    double best = force_res == 1 ? 1e300 : stream_bytes * (scaled ? (double)ntn0 * p.BN / N : 0.8);
    for (int ci = 0; ci < 6 && force_res != 0; ++ci) {
      const int bn = parts * cws[ci];
      if (bn > 256) continue;
      // gated GEMMs: one N tile only (the gate transform of A would be repeated per N tile) and a weight that leaves
      // room for a deep A ring -- their pace is set by stages in flight over the load + transform + MMA chain
      if (scaled && ((N + bn - 1) / bn != 1 || (size_t)k_blocks * bn * kBK * 2 > 96 * 1024)) continue;
      if (force_res == 1 && force_bn > 0 && bn != force_bn) continue;
      const size_t fixed = (size_t)k_blocks * bn * kBK * 2 + (size_t)bn * 256 + tail_bytes(bn);
      // A alone is 16 KB per k-block: the pipeline needs depth (>= 5 stages) to cover the L2 latency, and a narrow tile
      // (BN = 128: one k-block per 256 MMA cycles = 64 B/clk) asks more of the SM's L2 port than it delivers
      if (fixed + (force_res == 1 ? 2 : 5) * (size_t)kBM * kBK * 2 > budget) continue;
      const int ntn = (N + bn - 1) / bn;
      if (ntn > num_sms()) continue;
      if (bn < 192 && ntn > 1 && force_res != 1) continue;
      long long ms = num_sms() / ntn;
      if (ms > n_tm) ms = n_tm;
      const long long tpc = (n_tm + ms - 1) / ms;
      if (tpc < 4 && force_res != 1) continue;                       // too few tiles to amortise the weight load
      ms = (n_tm + tpc - 1) / tpc;
      const double quant = (double)(n_tm * ntn) / (double)(tpc * num_sms());          // SM-time actually used
      const double pad = (double)ntn * bn / N;                                        // padded output columns
      const double bytes = ((double)ntn * M * K * 2 + (double)ms * ntn * bn * K * 2) * pad / quant;
      if (bytes < best) {
        best = bytes;
        p.b_res = 1;
        p.BN = bn;
        p.cw = cws[ci];
        p.m_splits = (int)ms;
        p.tiles_per_cta = tpc;
      }
    }
  }
  p.n_tiles_n = (N + p.BN - 1) / p.BN;
  // Streaming plans re-fetch the weight tile for every 128 rows, and on the wide-K layers the bytes DELIVERED to each SM
  // (A + B stages, ~28 B/clk/SM measured), not HBM, set the pace.  A CTA pair (cta_group::2, M = 256 across two SMs)
  // stages only HALF of the weight tile's rows per SM: (16 + BN/8) KB instead of (16 + BN/4) KB per k-block, and more
  // stages fit.  (Sharing the weight stage by TMA multicast over a cluster instead -- every SM still receives the whole
  // tile -- was measured first: no gain, 0-3 % slower; L2 reads are not the limit.)
  p.cl = 1;
  if (!p.b_res) {
    int want = tuning && tuning->cluster != 0 ? tuning->cluster : ((size_t)p.BN * kBK * 2 >= (size_t)kBM * kBK && k_blocks >= 4 ? 2 : 1);
    if (want != 2 || (p.BN / 2) % 8 != 0 || p.BN % 32 != 0 || n_tm < 4) want = 1;
    p.cl = want;
  }
  p.rotate = (tuning && tuning->rotate < 0) ? 0 : 1;
  // Shared-A plans: a streaming GEMM with exactly two N tiles fetches (and, gated, transforms) every A stage twice.  What
  // paces the wide-K project GEMMs of the 12x12 stage is the byte rate INTO the SMs (L2 -> SM: 16 KB of A + the weight rows
  // per k-block and N tile; ~7-8 TB/s over the chip whether the bytes come from HBM or L2), so both N tiles ride on one A
  // stage and accumulate side by side in the two TMEM buffers: per k-block 16 + 2 x BN/8 KB instead of 2 x (16 + BN/8) KB.
  // Gated GEMMs only: their epilogue (bias + residual, no activation) is short, and without accumulator double-buffering
  // it is exposed once per M tile.
  p.nsub = 1;
  {
    const int want = tuning ? tuning->share_a : 0;
    const bool ok = !p.b_res && scaled && p.cl == 2 && p.n_tiles_n == 2 && 2 * p.BN <= (int)kTmemCols;      // (the kernel variant exists for gated CTA-pair plans)
    if (ok && (want == 1 || (want == 0 && k_blocks >= 8))) p.nsub = 2;
  }
  p.nb = p.cw > 64 ? 2 : 1;
  p.bw = p.cw / p.nb;
  p.swz = p.bw == 64 ? 3 : (p.bw == 32 ? 2 : (p.bw == 16 ? 1 : 0));
  p.n_tiles_m = n_tm;
  p.n_tiles = ((n_tm + p.cl - 1) / p.cl) * p.n_tiles_n;      // cluster tiles (= tiles when cl == 1)
  p.k_blocks = k_blocks;
  p.rows_per_image = rows_per_image > 0 ? rows_per_image : 1;
  p.act = act;
  const size_t a_stage = (size_t)kBM * kBK * 2;
  const size_t resident = p.b_res ? (size_t)k_blocks * p.BN * kBK * 2 : 0;
  const size_t staging = (size_t)p.BN * 256;            // all epilogue warps, one buffer each
  const size_t tail = tail_bytes(p.BN) + resident;
  size_t stage_bytes = 0;
  int stages = 0;
  for (;;) {
    const size_t b_stage = (size_t)(p.cl == 2 ? p.BN / 2 : p.BN) * kBK * 2;
    stage_bytes = p.b_res ? a_stage : a_stage + (size_t)p.nsub * b_stage;
    p.nbuf = (2 * staging + (p.b_res ? 5 : 3) * stage_bytes + tail <= 222 * 1024) ? 2 : 1;     // (the measured plans' rule)
    // gated GEMMs write little and wait long (load -> gate transform -> MMA per stage): a pipeline stage is worth more
    // than a second staging buffer unless six stages fit anyway
    if (scaled && p.nbuf == 2 && 2 * staging + 6 * stage_bytes + tail > budget) p.nbuf = 1;
    stages = (int)((budget - tail - p.nbuf * staging) / stage_bytes);
    if (p.nsub == 2 && stages < 3) { p.nsub = 1; continue; }      // a shared-A stage is up to 80 KB: without three of them, do not share
    break;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  DFV_REQUIRE(stages >= 2, "dfv_pw_gemm_fwd: tile does not fit shared memory (N=%d)", N);
  p.stages = stages;
  // keep one CTA per SM (each allocates all 512 TMEM columns): ask for > half of the SM's smem
  smem = (size_t)stages * stage_bytes + p.nbuf * staging + tail;
  if (smem < 120 * 1024) smem = 120 * 1024;
  if (p.b_res) {
    grid = (long long)p.m_splits * p.n_tiles_n;
  } else {
    long long slots = num_sms() / p.cl;                   // clusters that can be resident (refined by the launcher's occupancy query)
    if (max_clusters > 0 && p.cl > 1 && max_clusters < slots) slots = max_clusters;
    long long nc = p.n_tiles < slots ? p.n_tiles : slots;
    p.tiles_per_cta = (p.n_tiles + nc - 1) / nc;
    if (p.nsub == 2) p.tiles_per_cta += p.tiles_per_cta & 1;      // whole M tiles (both N tiles) per CTA
    nc = (p.n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    grid = nc * p.cl;
  }
  return DFV_OK;
}

/* Debug / documentation aid (host only): the tile plan of a bf16 tensor-core GEMM.
 * out[0..8] = BN, weight-stationary flag, pipeline stages, staging buffers, grid, tiles per CTA, smem bytes, N tiles,
 * CTAs per cluster (2 = CTA-pair plan; the grid is the planner's estimate before the launcher's occupancy query). */
extern "C" int dfv_gemm_plan_info(long long M, int K, int N, int scaled, int* out) {
  TcParams p;
  size_t smem = 0;
  long long grid = 0;
  if (!out || M <= 0 || K <= 0 || N <= 0 || K % 8 || N % 8) {
    set_error("dfv_gemm_plan_info: bad shape");
    return DFV_ERR_INVALID;
  }
  DFV_TRY(plan_tc(p, smem, grid, M, K, N, scaled != 0, 1, 0));
  out[0] = p.BN; out[1] = p.b_res; out[2] = p.stages; out[3] = p.nbuf; out[4] = (int)grid; out[5] = (int)p.tiles_per_cta;
  out[6] = (int)smem; out[7] = p.n_tiles_n; out[8] = p.cl; out[9] = p.nsub;
  return DFV_OK;
}

// Resident-cluster capacity of the device for one kernel variant / cluster size at (about) this much shared memory;
// queried once per (variant, cluster size) and cached.  0 = clusters of that size cannot be launched.
static int max_active_clusters(int kidx, int cl, size_t smem) {
  static int cache[8][5] = {};     // 0 = unknown, -1 = unsupported
  if (cache[kidx][cl] != 0) return cache[kidx][cl] < 0 ? 0 : cache[kidx][cl];
  const void* fn = nullptr;
  switch (kidx) {
    case 0: fn = (const void*)pw_gemm_tc_kernel<false, DFV_ACT_NONE, false, true>; break;
    case 1: fn = (const void*)pw_gemm_tc_kernel<false, DFV_ACT_NONE, true, true>; break;
    case 2: fn = (const void*)pw_gemm_tc_kernel<false, DFV_ACT_SILU, false, true>; break;
    case 3: fn = (const void*)pw_gemm_tc_kernel<false, DFV_ACT_SILU, true, true>; break;
    case 4: fn = (const void*)pw_gemm_tc_kernel<true, DFV_ACT_NONE, false, true>; break;
    case 5: fn = (const void*)pw_gemm_tc_kernel<true, DFV_ACT_NONE, true, true>; break;
    case 6: fn = (const void*)pw_gemm_tc_kernel<true, DFV_ACT_SILU, false, true>; break;
    default: fn = (const void*)pw_gemm_tc_kernel<true, DFV_ACT_SILU, true, true>; break;
  }
  int n = 0;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms() / cl * cl));
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
  } else {
    (void)cudaGetLastError();
  }
  cache[kidx][cl] = n > 0 ? n : -1;
  return n > 0 ? n : 0;
}

static int launch_tc(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                     const void* residual, void* out, long long M, int K, int N, int act, const dfv_gemm_tuning* tuning,
                     cudaStream_t st) {
  TcParams p;
  size_t smem = 0;
  long long grid = 0;
  DFV_TRY(plan_tc(p, smem, grid, M, K, N, a_scale != nullptr, rows_per_image, act, tuning));
  if (p.cl > 1) {
    // how many clusters of this shape the device can hold at once (GPCs with an odd SM count leave SMs unpaired): a
    // persistent grid larger than that would run its last clusters as a second wave
    const int kidx = (a_scale ? 4 : 0) + (act == DFV_ACT_SILU ? 2 : 0) + (residual ? 1 : 0);
    const int mc = max_active_clusters(kidx, p.cl, smem);
    if (mc <= 0) {
      dfv_gemm_tuning t1 = tuning ? *tuning : dfv_gemm_tuning{-1, 0, 0};
      t1.cluster = -1;
      DFV_TRY(plan_tc(p, smem, grid, M, K, N, a_scale != nullptr, rows_per_image, act, &t1));
    } else {
      DFV_TRY(plan_tc(p, smem, grid, M, K, N, a_scale != nullptr, rows_per_image, act, tuning, mc));
    }
  }

  CUtensorMap tm_a, tm_b, tm_out, tm_res;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)kBK, (uint32_t)kBM};
    DFV_TRY(make_tensor_map(&tm_a, DFV_BF16, 2, a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)kBK, (uint32_t)(p.b_res ? p.BN : p.BN / p.cl)};     // CTA pair: every CTA loads its half of the rows
    DFV_TRY(make_tensor_map(&tm_b, DFV_BF16, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {(uint32_t)p.bw, 32};
    const CUtensorMapSwizzle sw = p.swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (p.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (p.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    DFV_TRY(make_tensor_map(&tm_out, DFV_BF16, 2, out, dims, strides, box, sw));
    tm_res = tm_out;      // unused without a residual
    if (residual) DFV_TRY(make_tensor_map(&tm_res, DFV_BF16, 2, residual, dims, strides, box, sw));   // the residual: same geometry as the output
  }
  DFV_TRY(init_timeout_word_tu());
#define TC_LAUNCH(S_, A_, R_)                                                                                                   \
  do {                                                                                                                          \
    static thread_local bool configured = false;                                                                                \
    if (!configured) {                                                                                                          \
      DFV_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<S_, A_, R_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   \
      DFV_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<S_, A_, R_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   \
      DFV_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<S_, A_, R_, true, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   \
      configured = true;                                                                                                        \
    }                                                                                                                           \
    cudaLaunchConfig_t cfg = {};                                                                                                \
    cfg.gridDim = dim3((unsigned)grid);                                                                                         \
    cfg.blockDim = dim3(kTcThreads);                                                                                            \
    cfg.dynamicSmemBytes = smem;                                                                                                \
    cfg.stream = st;                                                                                                            \
    cudaLaunchAttribute cattr[1];                                                                                               \
    cattr[0].id = cudaLaunchAttributeClusterDimension;                                                                          \
    cattr[0].val.clusterDim.x = (unsigned)p.cl;                                                                                 \
    cattr[0].val.clusterDim.y = 1;                                                                                              \
    cattr[0].val.clusterDim.z = 1;                                                                                              \
    cfg.attrs = cattr;                                                                                                          \
    cfg.numAttrs = p.cl > 1 ? 1 : 0;                                                                                            \
    if (p.cl == 2 && p.nsub == 2)                                                                                               \
      DFV_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<S_, A_, R_, true, S_>, tm_a, tm_b, tm_out, tm_res, bias,              \
                                  (const __nv_bfloat16*)a_scale, p));                                                           \
    else if (p.cl == 2)                                                                                                         \
      DFV_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<S_, A_, R_, true>, tm_a, tm_b, tm_out, tm_res, bias,                  \
                                  (const __nv_bfloat16*)a_scale, p));                                                           \
    else                                                                                                                        \
      DFV_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<S_, A_, R_, false>, tm_a, tm_b, tm_out, tm_res, bias,                 \
                                  (const __nv_bfloat16*)a_scale, p));                                                           \
  } while (0)
  const bool silu_act = act == DFV_ACT_SILU;
  if (a_scale) {
    if (residual) { if (silu_act) TC_LAUNCH(true, DFV_ACT_SILU, true); else TC_LAUNCH(true, DFV_ACT_NONE, true); }
    else { if (silu_act) TC_LAUNCH(true, DFV_ACT_SILU, false); else TC_LAUNCH(true, DFV_ACT_NONE, false); }
  } else {
    if (residual) { if (silu_act) TC_LAUNCH(false, DFV_ACT_SILU, true); else TC_LAUNCH(false, DFV_ACT_NONE, true); }
    else { if (silu_act) TC_LAUNCH(false, DFV_ACT_SILU, false); else TC_LAUNCH(false, DFV_ACT_NONE, false); }
  }
#undef TC_LAUNCH
  DFV_LAUNCH_CHECK();
  if (debug_flags() & 32) {   // bisecting aid: attribute an asynchronous fault to this launch
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("tc GEMM faulted: %s M=%lld K=%d N=%d BN=%d b_res=%d stages=%d nbuf=%d cw=%d bw=%d scale=%d res=%d act=%d grid=%lld smem=%zu",
                cudaGetErrorString(e), M, K, N, p.BN, p.b_res, p.stages, p.nbuf, p.cw, p.bw, a_scale != nullptr,
                residual != nullptr, act, grid, smem);
      return DFV_ERR_CUDA;
    }
  }
  return DFV_OK;
}

// ---- row folding of thin 1x1 convolutions --------------------------------------------------------------
// A [M][K] activation with K <= 48 channels is a poor tensor-core / TMA operand: 128-row tiles carry 6-12 KB and the
// per-tile pipeline latency, not HBM, sets the pace (measured 1.4-2.7 TB/s on the 190x190 layers).  The SAME memory
// read as [M/f][f*K] times the block-diagonal weight diag(W, ..., W) gives the SAME output memory [M/f][f*N]:
// f x fewer, f x fatter tiles for f x the (idle) tensor FLOPs.  The SE gate and bias are tiled f times.
template <typename T>
__global__ void fold_weight_kernel(const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ wf,
                                   float* __restrict__ bf, int N, int K, int f) {
  pdl_prologue();
  const int total = f * N * f * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int col = i % (f * K), row = i / (f * K);
    const int a = row / N, n = row % N, b = col / K, k = col % K;
    T v = w[(size_t)n * K + k];
    if (a != b) v = T(0.f);
    wf[i] = v;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f * N; i += gridDim.x * blockDim.x) bf[i] = bias[i % N];
}
template <typename T>
__global__ void tile_gate_kernel(const T* __restrict__ g, T* __restrict__ gf, int B, int K, int f) {
  pdl_prologue();
  const int total = B * f * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / (f * K), k = (i % (f * K)) % K;
    gf[i] = g[(size_t)b * K + k];
  }
}

constexpr int kFoldWElems = 64 * 1024, kFoldBias = 1024, kFoldGate = 256;

static int pick_fold(int dtype, long long M, int K, int N, int rows_per_image, bool gated, int B) {
  if (dtype != DFV_BF16 || K > 48) return 1;
  for (int f = 4; f >= 2; f >>= 1) {
    if (M % f || (gated && rows_per_image % f)) continue;
    // folded width: one N tile, or (f = 4) an exact multiple of the 192-column tile (e.g. 24 -> 144: 576 = 3 x 192)
    const bool exact4 = f == 4 && !gated && f * N <= 768 && (f * N) % 192 == 0 && N % 48 == 0 && N < 192;
    if (f * N > (f == 4 ? 256 : 512) && !exact4) continue;
    if ((size_t)f * N * f * K > (size_t)kFoldWElems || f * N > kFoldBias || (gated && f * K > kFoldGate)) continue;
    return f;
  }
  return 1;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_pw_gemm_fwd(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                               const void* residual, void* out, int dtype, long long M, int K, int N, int act,
                               dfv_stream_t stream) {
  return dfv_pw_gemm_fwd_tuned(a, w, bias, a_scale, rows_per_image, residual, out, dtype, M, K, N, act, nullptr, stream);
}

extern "C" int dfv_pw_gemm_fwd_tuned(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                                     const void* residual, void* out, int dtype, long long M, int K, int N, int act,
                                     const dfv_gemm_tuning* tuning, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(a && w && bias && out, "dfv_pw_gemm_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_pw_gemm_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(M > 0 && K > 0 && N > 0, "dfv_pw_gemm_fwd: bad shape (M=%lld K=%d N=%d)", M, K, N);
  DFV_REQUIRE(dtype == DFV_F32 || (K % 8 == 0 && N % 8 == 0), "dfv_pw_gemm_fwd: bf16 needs K %% 8 == 0 and N %% 8 == 0 (K=%d N=%d)", K, N);
  DFV_REQUIRE(!a_scale || (rows_per_image > 0 && M % rows_per_image == 0), "dfv_pw_gemm_fwd: a_scale needs rows_per_image dividing M");
  DFV_REQUIRE(act == DFV_ACT_NONE || act == DFV_ACT_SILU, "dfv_pw_gemm_fwd: bad act %d", act);
  cudaStream_t st = as_stream(stream);
  const int dbg = debug_flags();
  const bool tc = dtype == DFV_BF16 && !((dbg & 8) && a_scale) && !((dbg & 16) && !a_scale);
  // algorithmic bytes: A read once, out written once, residual read once, weights once
  const double es = (double)dtype_size(dtype);
  ProfScope prof(tc ? (a_scale ? PK_PROJECT_GEMM : PK_EXPAND_GEMM) : PK_GEMM_SIMT,
                 es * ((double)M * K + (double)M * N + (residual ? (double)M * N : 0.0) + (double)N * K), 2.0 * (double)M * K * N, st);
  if (tc)
    return launch_tc(a, w, bias, a_scale, rows_per_image, residual, out, M, K, N, act, tuning, st);
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
  if (dtype == DFV_BF16)
    pw_gemm_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)w, bias,
                                                               (const __nv_bfloat16*)a_scale, rows_per_image, (const __nv_bfloat16*)residual,
                                                               (__nv_bfloat16*)out, M, K, N, act);
  else
    pw_gemm_simt_kernel<float, false><<<grid, 256, 0, st>>>((const float*)a, (const float*)w, bias, (const float*)a_scale, rows_per_image,
                                                         (const float*)residual, (float*)out, M, K, N, act);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Scratch of dfv_pw_conv_fwd: block-diagonal weights + tiled bias + tiled SE gate of a row-folded convolution. */
extern "C" size_t dfv_pw_fold_ws_bytes(int B) {
  return align_up((size_t)kFoldWElems * 2, 1024) + align_up((size_t)kFoldBias * 4, 1024) + align_up((size_t)(B > 0 ? B : 1) * kFoldGate * 2, 1024);
}

/* 1x1 convolution = dfv_pw_gemm_fwd, except that thin bf16 layers (K <= 48) run row-folded: the same memory read as
 * [M/f][f*K] against diag(W, ..., W) (see fold_weight_kernel).  fold_ws: dfv_pw_fold_ws_bytes(B) bytes, may be NULL
 * (no folding).  B = number of images (rows of the SE gate). */
extern "C" int dfv_pw_conv_fwd(const void* x, const void* w, const float* bias, const void* gate, int rows_per_image,
                               const void* residual, void* out, int dtype, int B, long long M, int K, int N, int act,
                               void* fold_ws, dfv_stream_t stream) {
  const int f = fold_ws ? pick_fold(dtype, M, K, N, rows_per_image, gate != nullptr, B) : 1;
  if (f == 1) return dfv_pw_gemm_fwd(x, w, bias, gate, rows_per_image, residual, out, dtype, M, K, N, act, stream);
  DFV_TRY(check_device());
  char* base = static_cast<char*>(fold_ws);
  __nv_bfloat16* fw = reinterpret_cast<__nv_bfloat16*>(base);
  float* fb = reinterpret_cast<float*>(base + align_up((size_t)kFoldWElems * 2, 1024));
  __nv_bfloat16* fg = reinterpret_cast<__nv_bfloat16*>(base + align_up((size_t)kFoldWElems * 2, 1024) + align_up((size_t)kFoldBias * 4, 1024));
  cudaStream_t st = as_stream(stream);
  DFV_PDL((fold_weight_kernel<__nv_bfloat16>), 16, 256, 0, st, (const __nv_bfloat16*)w, bias, fw, fb, N, K, f);
  DFV_LAUNCH_CHECK();
  if (gate) {
    DFV_PDL((tile_gate_kernel<__nv_bfloat16>), 32, 256, 0, st, (const __nv_bfloat16*)gate, fg, B, K, f);
    DFV_LAUNCH_CHECK();
  }
  return dfv_pw_gemm_fwd(x, fw, fb, gate ? fg : nullptr, rows_per_image / f, residual, out, dtype, M / f, f * K, f * N, act, stream);
}
