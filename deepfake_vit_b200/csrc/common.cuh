// Shared host/device helpers for libdfvit.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dfvit.h"

namespace dfv {

// ---------------------------------------------------------------- host: errors
void set_error(const char* fmt, ...);
int check_device();                       // DFV_OK or DFV_ERR_DEVICE (cached per device)
int num_sms();
void count_launch(int n = 1);
// Bisecting aid of DEBUG builds only (make DEBUG=1 -> -DDFV_DEBUG): DFV_DEBUG_FLAGS env (1 skip dwconv, 2 skip se,
// 4 skip stem, 8 simt project, 16 simt expand, 32 sync after tc GEMMs).  The product build has no environment
// switches and no process-global modes: the function is a compile-time 0.
#ifdef DFV_DEBUG
int debug_flags();
#else
constexpr int debug_flags() { return 0; }
#endif

#define DFV_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::dfv::set_error(__VA_ARGS__);              \
      return DFV_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define DFV_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ::dfv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return DFV_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define DFV_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) {                                                             \
      ::dfv::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return DFV_ERR_CUDA;                                                               \
    }                                                                                    \
    ::dfv::count_launch();                                                               \
  } while (0)

#define DFV_TRY(expr)            \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != DFV_OK) return rc_; \
  } while (0)

// Optional per-launch CUDA-event profiler (bench.py's roofline leg): every operator entry
// declares its kind and ALGORITHMIC bytes / flops; when enabled, start/stop events are recorded
// on the launching stream around the launch.
enum ProfKind { PK_STEM = 0, PK_EXPAND_GEMM, PK_DWCONV, PK_SE_GATE, PK_PROJECT_GEMM, PK_HEATMAP, PK_ATTENTION,
                PK_MLP_HEAD, PK_LOSS, PK_GEMM_SIMT, PK_BN, PK_WGRAD, PK_DWCONV_BWD, PK_NUM };
struct ProfScope {
  ProfScope(int kind, double bytes, double flops, cudaStream_t st);
  ~ProfScope();
  int idx;
  cudaStream_t st;
};

inline cudaStream_t as_stream(dfv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t dtype_size(int dtype) { return dtype == DFV_BF16 ? 2 : 4; }
inline bool valid_dtype(int dtype) { return dtype == DFV_F32 || dtype == DFV_BF16; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

unsigned int* timeout_device_ptr();      // host-mapped debug word (device address), api.cu

// ---------------------------------------------------------------- host: topology / blob (api.cu)
const dfv_block_info* topo_blocks(int* n);
int topo_stem_c();
int topo_head_c();
const void* blob_ptr(const void* blob, int dtype, int block, int kind);

// ---------------------------------------------------------------- host: TMA descriptors
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda link).
// dims/strides innermost first; strides in bytes for dims 1..rank-1.
int make_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// 8 consecutive channels <-> 8 fp32 registers.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float v[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void load8(const float* p, float v[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float v[8]) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// swish / SiLU.  EXACT: x / (1 + exp(-x)) with IEEE expf + division (fp32 parity mode).
// FAST: one MUFU.TANH: silu(x) = h*tanh(h) + h, h = x/2 (max rel err ~2^-11, below bf16 ulp).
// packed fp32 pair helpers (sm_100 FFMA2 / FADD2: two fp32 operations per issue slot)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
template <bool kFast>
__device__ __forceinline__ float silu(float x) {
  if constexpr (kFast) {
    float h = 0.5f * x, t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  } else {
    return x / (1.0f + expf(-x));
  }
}
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- mbarrier / TMA / tcgen05 PTX wrappers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) after ~2 s instead of hanging the GPU, after
// leaving a tag in host-mapped memory (dfv_last_timeout_word()) that says which wait starved.
static __device__ unsigned int* g_timeout_word = nullptr;   // one copy per translation unit (no -rdc)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (g_timeout_word) {
        *reinterpret_cast<volatile unsigned int*>(g_timeout_word) = 0x80000000u | (tag << 24) | (blockIdx.x & 0xffffff);
        __threadfence_system();
      }
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// CTA-pair (cta_group::2) load: the box lands in THIS CTA's shared memory, the bytes are counted on an mbarrier that may
// live in the peer CTA (`bar_cluster` is a shared::cluster address, see mapa_shared).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// arrive on an mbarrier anywhere in the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  // default semantics (.release.cta), as CUTLASS' ClusterBarrier::arrive(cta_id): the .release.cluster form costs a
  // cluster-scope memory barrier per arrival (measured: the gated pair GEMM 78 -> 122 us, top stall "membar")
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA pair (cta_group::2): one thread of the leader CTA issues an M = 256 MMA that runs on both SMs -- each CTA holds
// its 128 rows of A and HALF of the B tile's rows in its own shared memory (same offsets), and its 128 accumulator rows in
// its own TMEM.  The commit is delivered to the same-offset mbarrier in both CTAs.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, with the 16 destination registers of a tmem_ld16 tied to it: code that reads them cannot be scheduled
// above the wait (used where another tmem_ld16 is already in flight into a second register set).
__device__ __forceinline__ void tmem_ld_wait16(uint32_t v[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A training step is ~3000 mostly small launches; stream-ordered launches pay the full launch latency and the previous
// grid's tail at every boundary.  Kernels launched with launch_pdl() may be SCHEDULED while their predecessor still
// runs: pdl_prologue() first lets this grid's own successor do the same, then blocks until every prerequisite grid has
// completed and flushed its memory (griddepcontrol.wait), so data dependencies are exactly those of plain stream order.
// (Without the launch attribute both instructions are no-ops.)
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Point this translation unit's g_timeout_word at the process-wide host-mapped word.
static inline int init_timeout_word_tu() {
  static bool done = false;
  if (done) return DFV_OK;
  unsigned int* d = timeout_device_ptr();
  if (d) DFV_CUDA(cudaMemcpyToSymbol(g_timeout_word, &d, sizeof(d)));
  done = true;
  return DFV_OK;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__
}  // namespace dfv

#ifdef __CUDACC__
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-dependent-launch attribute (common.cuh)
#define DFV_PDL(kern, grid, block, smem, st, ...)                                                        \
  do {                                                                                                   \
    cudaError_t pe_ = ::dfv::launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__); \
    if (pe_ != cudaSuccess) {                                                                            \
      ::dfv::set_error("launch failed: %s (%s:%d)", cudaGetErrorString(pe_), __FILE__, __LINE__);        \
      return DFV_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)
#endif
