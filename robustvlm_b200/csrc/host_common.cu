// Host-side helpers shared by all kernels: driver entry point for tensor-map
// encoding (resolved at run time so the library links without libcuda and can be
// dlopen'ed on a machine with no GPU), SM count, error strings.
#include "gemm.cuh"
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/rvlm.h"

namespace rvlm {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return RVLM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) return RVLM_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : RVLM_ERR_TMAP;
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t s1, uint64_t s2, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return RVLM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((s1 * 2) & 15) || ((s2 * 2) & 15))
    return RVLM_ERR_ALIGN;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : RVLM_ERR_TMAP;
}

// Tensor map for TMA STORES of a 16-bit [rows, cols] row-major output (pitch ld elements): box =
// 32 rows x 32 columns (64-byte rows, 64B swizzle), the per-warp staging tile of the GEMM epilogue.
int make_tmap_2d_store16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return RVLM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15) || (cols & 31)) return RVLM_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : RVLM_ERR_TMAP;
}

int epi_tma_level() {   // (read per call: the tests cover every level)
  const char* e = getenv("RVLM_EPI_TMA");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t e0, e1; double flops, mma; const char* tag; bool gemm; int n, k; };
static int g_prof_on = 0;   // 0 off, 1 tcgen05 GEMM launches, 2 every bracketed kernel
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;
static std::string g_prof_breakdown = "{}";   // per-kernel JSON of the last collect()


GemmProfScope::GemmProfScope(cudaStream_t s, double flops, double mma_factor, const char* tag,
                             bool is_gemm, int n, int k)
    : st(s), slot(-1) {
  if (g_prof_on == 0 || (!is_gemm && g_prof_on < 2)) return;
  if (g_prof_used == g_prof.size()) {
    ProfRec r{};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    g_prof.push_back(r);
  }
  slot = (int)g_prof_used++;
  g_prof[slot].flops = flops;
  g_prof[slot].mma = flops * mma_factor;
  g_prof[slot].tag = tag;
  g_prof[slot].gemm = is_gemm;
  g_prof[slot].n = n;
  g_prof[slot].k = k;
  cudaEventRecord(g_prof[slot].e0, st);
}
GemmProfScope::~GemmProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, st);
}

bool gemm_use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RVLM_GEMM_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;   // default on (validated on B200: matches cuBLAS)
  }
  return v == 1;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

}  // namespace rvlm

extern "C" const char* rvlm_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case RVLM_ERR_INVAL: return "invalid argument";
    case RVLM_ERR_ALIGN: return "pointer or pitch not 16-byte aligned";
    case RVLM_ERR_DRIVER: return "cuTensorMapEncodeTiled entry point unavailable";
    case RVLM_ERR_TMAP: return "tensor map encode failed";
    case RVLM_ERR_NOMEM: return "workspace too small";
    case RVLM_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return code <= -1000 ? "CUDA runtime error (code = -1000 - cudaError_t)" : "unknown";
  }
}

extern "C" int rvlm_abi_version(void) { return RVLM_ABI_VERSION; }

extern "C" long long rvlm_launch_count(void) { return rvlm::g_launches.load(); }
extern "C" void rvlm_launch_count_reset(void) { rvlm::g_launches.store(0); }

extern "C" int rvlm_profile_gemm(int enable) {
  rvlm::g_prof_on = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  rvlm::g_prof_used = 0;
  return 0;
}

extern "C" int rvlm_profile_gemm_collect(double* total_ms, double* total_flops,
                                         long long* launches) {
  // totals over the tcgen05 GEMM launches only (the roofline's kernel); every bracketed kernel
  // goes into the per-tag breakdown
  double ms = 0.0, fl = 0.0;
  long long ng = 0;
  struct Agg { double ms = 0.0, flops = 0.0; long long n = 0; };
  std::map<std::string, Agg> agg;
  for (size_t i = 0; i < rvlm::g_prof_used; ++i) {
    cudaError_t e = cudaEventSynchronize(rvlm::g_prof[i].e1);
    if (e != cudaSuccess) return -(int)e - 1000;
    float t = 0.f;
    e = cudaEventElapsedTime(&t, rvlm::g_prof[i].e0, rvlm::g_prof[i].e1);
    if (e != cudaSuccess) return -(int)e - 1000;
    if (rvlm::g_prof[i].gemm) {
      ms += t;
      fl += rvlm::g_prof[i].flops;
      ++ng;
    }
    std::string key = rvlm::g_prof[i].tag ? rvlm::g_prof[i].tag : "?";
    if (rvlm::g_prof[i].gemm)   // one line per GEMM shape: epilogue / N / K
      key += " N=" + std::to_string(rvlm::g_prof[i].n) + " K=" + std::to_string(rvlm::g_prof[i].k);
    Agg& a = agg[key];
    a.ms += t;
    a.flops += rvlm::g_prof[i].flops;
    ++a.n;
  }
  std::string js = "{";
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.4f, \"launches\": %lld, \"gflop\": %.3f}",
             js.size() > 1 ? ", " : "", kv.first.c_str(), kv.second.ms, kv.second.n,
             kv.second.flops * 1e-9);
    js += buf;
  }
  js += "}";
  rvlm::g_prof_breakdown = js;
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = ng;
  rvlm::g_prof_used = 0;
  return 0;
}

extern "C" int rvlm_profile_breakdown(char* out, int cap) {
  const std::string& js = rvlm::g_prof_breakdown;
  if (out && cap > 0) {
    const size_t n = js.size() < (size_t)cap - 1 ? js.size() : (size_t)cap - 1;
    memcpy(out, js.data(), n);
    out[n] = 0;
  }
  return (int)js.size() + 1;
}
