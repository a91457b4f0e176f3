// Error state, device checks, TMA descriptor encoding, EfficientNet-B4 topology and the
// folded-weight blob layout.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "common.cuh"

namespace dfv {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};   // process-wide: the backward pass runs on autograd's worker thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
#ifdef DFV_DEBUG
int debug_flags() {
  static int flags = -1;
  if (flags < 0) {
    const char* e = getenv("DFV_DEBUG_FLAGS");
    flags = e ? atoi(e) : 0;
  }
  return flags;
}
#endif

static unsigned int* g_timeout_host = nullptr;
static unsigned int* g_timeout_dev = nullptr;

// A host-mapped word the device writes to when a bounded mbarrier wait starves (readable even
// after the context is lost).
unsigned int* timeout_device_ptr() {
  if (g_timeout_dev) return g_timeout_dev;
  unsigned int* h = nullptr;
  if (cudaHostAlloc((void**)&h, sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess) return nullptr;
  *h = 0;
  unsigned int* d = nullptr;
  if (cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) return nullptr;
  g_timeout_host = h;
  g_timeout_dev = d;
  return d;
}

struct ProfRec {
  cudaEvent_t a, b;
  int kind;
  double bytes, flops;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(int kind, double bytes, double flops, cudaStream_t s) : idx(-1), st(s) {
  if (!g_prof_on) return;
  ProfRec r;
  r.kind = kind;
  r.bytes = bytes;
  r.flops = flops;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].b, st);
}

int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = DFV_ERR_DEVICE;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device available: libdfvit has no CPU path");
    return DFV_ERR_DEVICE;
  }
  if (dev == cached_dev) {
    if (cached_rc != DFV_OK) set_error("device %d is not sm_100: libdfvit is built for sm_100a only", dev);
    return cached_rc;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed");
    return DFV_ERR_CUDA;
  }
  cached_dev = dev;
  cached_rc = (major == 10) ? DFV_OK : DFV_ERR_DEVICE;
  if (cached_rc != DFV_OK)
    set_error("device %d has compute capability %d.x: libdfvit is built for sm_100a only", dev, major);
  return cached_rc;
}

int num_sms() {
  static thread_local int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return DFV_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapDataType dt = dtype == DFV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0);
    return DFV_ERR_CUDA;
  }
  return DFV_OK;
}

// ------------------------------------------------------------------------------------
// EfficientNet-B4 topology, derived the way efficientnet-pytorch 0.7.1 derives it
// (SURVEY.md Appendix A.1-A.3): width 1.4, depth 1.8, divisor 8, image size 380.
// ------------------------------------------------------------------------------------
struct Topology {
  dfv_block_info blocks[64];
  int n = 0;
  int stem_c = 0, head_c = 0;
};

static int round_filters(int f) {
  const double w = 1.4;
  const int div = 8;
  double scaled = f * w;
  int nf = (int)(scaled + div / 2.0) / div * div;
  if (nf < div) nf = div;
  if (nf < 0.9 * scaled) nf += div;
  return nf;
}
static int round_repeats(int r) { return (int)ceil(1.8 * r); }

static const Topology& topology() {
  static Topology t;
  if (t.n) return t;
  // {repeats, kernel, stride, expand, in, out}
  static const int base[7][6] = {{1, 3, 1, 1, 32, 16},  {2, 3, 2, 6, 16, 24},  {2, 5, 2, 6, 24, 40},
                                 {3, 3, 2, 6, 40, 80},  {3, 5, 1, 6, 80, 112}, {4, 5, 2, 6, 112, 192},
                                 {1, 3, 1, 6, 192, 320}};
  t.stem_c = round_filters(32);
  int size = 380;
  size = (size + 1) / 2;  // after the stride-2 stem
  for (int s = 0; s < 7; ++s) {
    int rep = round_repeats(base[s][0]);
    int cin = round_filters(base[s][4]), cout = round_filters(base[s][5]);
    for (int r = 0; r < rep; ++r) {
      dfv_block_info b;
      b.c_in = (r == 0) ? cin : cout;
      b.c_out = cout;
      b.c_mid = b.c_in * base[s][3];
      b.kernel = base[s][1];
      b.stride = (r == 0) ? base[s][2] : 1;
      int out = (size + b.stride - 1) / b.stride;
      int pad = (out - 1) * b.stride + b.kernel - size;
      if (pad < 0) pad = 0;
      b.pad_lo = pad / 2;
      b.pad_hi = pad - pad / 2;
      b.se_squeeze = b.c_in / 4 > 1 ? b.c_in / 4 : 1;
      b.has_expand = base[s][3] != 1;
      b.has_skip = (b.stride == 1 && b.c_in == b.c_out);
      t.blocks[t.n++] = b;
      size = out;
    }
  }
  t.head_c = round_filters(1280);
  return t;
}

// ------------------------------------------------------------------------------------
// Weight blob layout
// ------------------------------------------------------------------------------------
struct Slot {
  size_t off = 0, elems = 0;
  bool valid = false;
};
struct BlobLayout {
  Slot stem[2], head[2];
  Slot block[64][DFV_W_KINDS];
  size_t bytes = 0;
};

static const BlobLayout& blob_layout(int dtype) {
  static BlobLayout L[2];
  BlobLayout& l = L[dtype];
  if (l.bytes) return l;
  const Topology& t = topology();
  size_t off = 0;
  const size_t ts = dtype_size(dtype);
  auto put = [&](Slot& s, size_t elems, size_t esz) {
    off = align_up(off, 256);
    s.off = off;
    s.elems = elems;
    s.valid = true;
    off += elems * esz;
  };
  put(l.stem[0], 27 * (size_t)t.stem_c, 4);
  put(l.stem[1], t.stem_c, 4);
  for (int i = 0; i < t.n; ++i) {
    const dfv_block_info& b = t.blocks[i];
    if (b.has_expand) {
      put(l.block[i][DFV_W_EXPAND], (size_t)b.c_mid * b.c_in, ts);
      put(l.block[i][DFV_W_EXPAND_BIAS], b.c_mid, 4);
    }
    put(l.block[i][DFV_W_DW], (size_t)b.kernel * b.kernel * b.c_mid, 4);
    put(l.block[i][DFV_W_DW_BIAS], b.c_mid, 4);
    put(l.block[i][DFV_W_SE_REDUCE], (size_t)b.se_squeeze * b.c_mid, 4);
    put(l.block[i][DFV_W_SE_REDUCE_BIAS], b.se_squeeze, 4);
    put(l.block[i][DFV_W_SE_EXPAND], (size_t)b.se_squeeze * b.c_mid, 4);
    put(l.block[i][DFV_W_SE_EXPAND_BIAS], b.c_mid, 4);
    put(l.block[i][DFV_W_PROJECT], (size_t)b.c_out * b.c_mid, ts);
    put(l.block[i][DFV_W_PROJECT_BIAS], b.c_out, 4);
  }
  put(l.head[0], (size_t)t.head_c * t.blocks[t.n - 1].c_out, ts);
  put(l.head[1], t.head_c, 4);
  l.bytes = align_up(off, 256);
  return l;
}

const dfv_block_info* topo_blocks(int* n) {
  const Topology& t = topology();
  *n = t.n;
  return t.blocks;
}
int topo_stem_c() { return topology().stem_c; }
int topo_head_c() { return topology().head_c; }
const void* blob_ptr(const void* blob, int dtype, int block, int kind) {
  size_t off = 0, elems = 0;
  if (dfv_blob_slot(dtype, block, kind, &off, &elems) != DFV_OK) return nullptr;
  return static_cast<const char*>(blob) + off;
}

}  // namespace dfv

using namespace dfv;

extern "C" {

int dfv_version(void) { return 200; }
const char* dfv_last_error(void) { return g_error; }
int dfv_device_check(void) { return check_device(); }
long long dfv_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

unsigned int dfv_last_timeout_word(void) { return g_timeout_host ? *g_timeout_host : 0; }

int dfv_profile_enable(int on) {
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on = on != 0;
  return DFV_OK;
}
int dfv_profile_count(void) { return (int)g_prof.size(); }
int dfv_profile_get(int i, int* kind, double* bytes, double* flops, float* ms) {
  DFV_REQUIRE(i >= 0 && i < (int)g_prof.size() && kind && bytes && flops && ms, "dfv_profile_get: bad index %d", i);
  const ProfRec& r = g_prof[i];
  DFV_CUDA(cudaEventSynchronize(r.b));
  DFV_CUDA(cudaEventElapsedTime(ms, r.a, r.b));
  *kind = r.kind;
  *bytes = r.bytes;
  *flops = r.flops;
  return DFV_OK;
}

int dfv_b4_num_blocks(void) { return topology().n; }
int dfv_b4_block(int idx, dfv_block_info* out) {
  const Topology& t = topology();
  DFV_REQUIRE(idx >= 0 && idx < t.n && out, "dfv_b4_block: bad index %d", idx);
  *out = t.blocks[idx];
  return DFV_OK;
}
int dfv_b4_stem_channels(void) { return topology().stem_c; }
int dfv_b4_head_channels(void) { return topology().head_c; }

int dfv_b4_output_hw(int H, int W, int* Ho, int* Wo) {
  DFV_REQUIRE(H >= 32 && W >= 32 && Ho && Wo, "dfv_b4_output_hw: input must be at least 32x32");
  const Topology& t = topology();
  int h = (H + 1 - 3) / 2 + 1, w = (W + 1 - 3) / 2 + 1;  // stem: pad (0,1), k3, s2
  for (int i = 0; i < t.n; ++i) {
    const dfv_block_info& b = t.blocks[i];
    h = (h + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
    w = (w + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
  }
  *Ho = h;
  *Wo = w;
  return DFV_OK;
}

size_t dfv_blob_bytes(int dtype) { return valid_dtype(dtype) ? blob_layout(dtype).bytes : 0; }

int dfv_blob_slot(int dtype, int block, int kind, size_t* offset, size_t* elems) {
  DFV_REQUIRE(valid_dtype(dtype), "dfv_blob_slot: bad dtype %d", dtype);
  DFV_REQUIRE(kind >= 0 && kind < DFV_W_KINDS && offset && elems, "dfv_blob_slot: bad kind %d", kind);
  const BlobLayout& l = blob_layout(dtype);
  const Slot* s = nullptr;
  if (kind == DFV_W_STEM || kind == DFV_W_STEM_BIAS) {
    s = &l.stem[kind - DFV_W_STEM];
  } else if (kind == DFV_W_HEAD || kind == DFV_W_HEAD_BIAS) {
    s = &l.head[kind - DFV_W_HEAD];
  } else {
    DFV_REQUIRE(block >= 0 && block < topology().n, "dfv_blob_slot: bad block %d", block);
    s = &l.block[block][kind];
  }
  DFV_REQUIRE(s->valid, "dfv_blob_slot: block %d has no tensor of kind %d", block, kind);
  *offset = s->off;
  *elems = s->elems;
  return DFV_OK;
}

}  // extern "C"
