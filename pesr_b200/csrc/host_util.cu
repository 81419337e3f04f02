// Error reporting, launch counter, device query and the TMA tensor-map cache.
#include <cstdlib>
#include "host_util.cuh"

#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

namespace pesr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// Packed weights are written by pack_weights / cast16 launches.  A PDL consumer may fetch weights BEFORE its dependency
// wait only if a kernel that itself waited for the writer lies between them IN THE SAME STREAM (see conv_igemm.cu: the
// conv kernels trigger their dependents only after their own wait).  Tracked per stream (the training step runs the
// VGG branch and the Discriminator on two streams): a stream is "dirty" from a weight write on it until the next
// tensor-core launch on it.
static cudaStream_t g_dirty_streams[8];
static int g_ndirty = 0;
static std::mutex g_dirty_mu;
void note_weight_write(cudaStream_t stream) {
  std::lock_guard<std::mutex> g(g_dirty_mu);
  for (int i = 0; i < g_ndirty; i++)
    if (g_dirty_streams[i] == stream) return;
  if (g_ndirty < 8) g_dirty_streams[g_ndirty++] = stream;
  else g_dirty_streams[7] = stream;      // more than 8 dirty streams: never happens; stays conservative for this one
}
// Called by the tensor-core launchers: true = weights may be prefetched early by THIS launch; marks the stream clean.
bool weights_settled(cudaStream_t stream) {
  std::lock_guard<std::mutex> g(g_dirty_mu);
  for (int i = 0; i < g_ndirty; i++)
    if (g_dirty_streams[i] == stream) {
      g_dirty_streams[i] = g_dirty_streams[--g_ndirty];
      return false;
    }
  return true;
}
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("PESR_NO_PDL");
    g_pdl = (e && e[0] == '1') ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on; }

// Shared memory (bytes) the tensor-core kernels may claim per CTA.  The full 227 KB maximises pipeline depth; a smaller
// budget leaves room for the blocks of a concurrently running memory-bound kernel of another stream on the same SM
// (the training step runs the VGG branch beside the Discriminator).  PESR_CONV_SMEM_KB overrides (120..227).
int conv_smem_budget() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("PESR_CONV_SMEM_KB");
    int kb = e ? atoi(e) : 227;
    if (kb < 120) kb = 120;
    if (kb > 227) kb = 227;
    v = kb * 1024;
  }
  return v;
}

static int g_reserve_sms = -1;     // -1: from the environment (PESR_RESERVE_SMS)

void set_reserved_sms(int n) { g_reserve_sms = n < 0 ? 0 : n; }

// SMs the persistent tensor-core kernels size their grids to: the device's count minus the reserved ones
// (PESR_OPT_RESERVE_SMS), kept even so that CTA pairs fit.
int num_sms() {
  if (g_reserve_sms < 0) {
    const char* e = getenv("PESR_RESERVE_SMS");
    g_reserve_sms = e ? atoi(e) : 0;
    if (g_reserve_sms < 0) g_reserve_sms = 0;
  }
  int n = device_sms() - g_reserve_sms;
  n &= ~1;
  return n < 2 ? 2 : n;
}

int device_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// ------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver symbol; resolve it through the runtime so the library
// links without libcuda (this container has no GPU driver; the GPU box does).
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[20];
  bool operator<(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};

int get_tensor_map(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  return get_tensor_map_ex(out, base, dtype, 3, rank, dims, strides_bytes, box);
}

int get_tensor_map_ex(CUtensorMap* out, const void* base, int elem, int swizzle, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  static std::mutex mu;
  static std::map<MapKey, CUtensorMap> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.v[0] = (uint64_t)(uintptr_t)base;
  k.v[1] = ((uint64_t)elem << 8) | (uint64_t)rank | ((uint64_t)dev << 16) | ((uint64_t)swizzle << 32);
  for (int i = 0; i < rank; i++) k.v[2 + i] = dims[i];
  for (int i = 0; i < rank - 1; i++) k.v[8 + i] = strides_bytes[i];
  for (int i = 0; i < rank; i++) k.v[14 + i] = box[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(k);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled could not be resolved (no CUDA driver?)");
    return PESR_E_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < rank - 1; i++) gstr[i] = strides_bytes[i];
  CUtensorMap m;
  const CUtensorMapDataType et = elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : elem == PESR_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(&m, et, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): base %p rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
              (int)r, base, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return PESR_E_DRIVER;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache[k] = m;
  }
  *out = m;
  return 0;
}

// ------------------------------------------------------------------------------------------
// per-launch profiling
// ------------------------------------------------------------------------------------------
struct ProfRec {
  cudaEvent_t e0, e1;
  int kind;
  double flops;
};
static bool g_prof = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_event_pool;

bool profiling_enabled() { return g_prof; }

static cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void profile_begin(int kind, double flops, cudaStream_t stream) {
  ProfRec r;
  r.e0 = get_event();
  r.e1 = get_event();
  r.kind = kind;
  r.flops = flops;
  cudaEventRecord(r.e0, stream);
  g_prof_recs.push_back(r);
}

void profile_end(int kind, cudaStream_t stream) {
  (void)kind;
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, stream);
}

}  // namespace pesr

extern "C" void pesr_profile_enable(int on) {
  pesr::g_prof = on != 0;
}

// Synchronises, sums the recorded launches of `kind`, returns them and clears the records of that kind.
extern "C" int pesr_profile_read(int kind, double* total_ms, long long* launches, double* flops) {
  using namespace pesr;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("profile_read: sync failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  double ms = 0, fl = 0;
  long long n = 0;
  std::vector<ProfRec> keep;
  for (auto& r : g_prof_recs) {
    if (r.kind != kind) {
      keep.push_back(r);
      continue;
    }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms += t;
      fl += r.flops;
      n++;
    }
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof_recs.swap(keep);
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (flops) *flops = fl;
  return 0;
}

extern "C" const char* pesr_last_error(void) { return pesr::g_err; }
extern "C" int pesr_version(void) { return 100; }
extern "C" long long pesr_launch_count(int reset) {
  long long v = pesr::g_launches.load();
  if (reset) pesr::g_launches.store(0);
  return v;
}
extern "C" int pesr_sizeof(int which) {
  return which == 0 ? (int)sizeof(pesr_conv_desc) : which == 1 ? (int)sizeof(pesr_wgrad_desc) : -1;
}
