// extern "C" surface of libdcvit.so (declared in include/dcvit.h) plus the host
// utilities shared by all launchers.
#include <atomic>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "host.h"

namespace dcv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// profiler
// ---------------------------------------------------------------------------------------------
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
struct ProfRec {
  cudaEvent_t a, b;
  int tag;
};
static std::vector<ProfRec> g_prof;
static const char* kProfNames[PT_COUNT] = {"gemm_nt",  "gemm_nn",   "gemm_tn",    "attn_fwd",   "attn_bwd_prep", "attn_bwd",
                                           "attn_bwd_fin", "ln_fwd", "ln_bwd",  "colsum",     "cast",          "im2col",
                                           "embed_gemm", "embed_misc", "tdl",    "cdl",        "embed_bwd",     "small"};

ProfScope::ProfScope(int tag, cudaStream_t s) : slot(-1), st(s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = tag;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  slot = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[slot].b, st);
}

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    // off unless DCV_PDL=1: measured with both sets of graphs in one process (tools/pdl_ab.py, JUMP-CP B=32) the step is
    // 4.6 % faster at C' = 1, unchanged at C' = 2 and 2-6 % SLOWER from C' = 4 up
    const char* e = getenv("DCV_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void debug_set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// side stream of the block backward
// ---------------------------------------------------------------------------------------------
// 0 = off, 1 = always, n > 1 = only for calls of at most n token rows; < 0 = not read yet (DCV_BWD_OVERLAP)
constexpr long long kBwdOverlapDefault = 1;
static std::atomic<long long> g_bwd_overlap{-1};
void debug_set_bwd_overlap(int on) { g_bwd_overlap.store(on < 0 ? -1 : on, std::memory_order_relaxed); }

SideBranch* side_branch(cudaStream_t main, int rows) {
  long long mode = g_bwd_overlap.load(std::memory_order_relaxed);
  if (mode < 0) {
    const char* e = getenv("DCV_BWD_OVERLAP");
    mode = (e != nullptr && e[0] >= '0' && e[0] <= '9') ? atoll(e) : kBwdOverlapDefault;
    g_bwd_overlap.store(mode, std::memory_order_relaxed);
  }
  if (mode == 0 || (mode > 1 && rows > mode) || g_prof_on.load(std::memory_order_relaxed)) return nullptr;
  static std::mutex mu;
  static SideBranch by_dev[64];
  static bool ready[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!ready[dev]) {
    // never create a stream while the caller is capturing (global capture mode rejects it): that call stays serial
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;
    }
    SideBranch sb;
    if (cudaStreamCreateWithFlags(&sb.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    bool ok = cudaEventCreateWithFlags(&sb.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&sb.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      return nullptr;
    }
    by_dev[dev] = sb;
    ready[dev] = true;
  }
  return &by_dev[dev];
}

int num_sms() {
  static std::atomic<int> sms_by_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int sms = sms_by_dev[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    sms_by_dev[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}

int ensure_smem_attr(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, uint64_t> done;  // kernel -> bit mask of devices that carry the attribute
  int dev = 0;
  DCV_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(kernel);
    if (it != done.end() && (it->second & bit)) return 0;
  }
  DCV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  std::lock_guard<std::mutex> lk(mu);
  done[kernel] |= bit;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor maps are pure functions of (pointer, geometry): the same ~60 maps recur every training step, so encoded
// descriptors are cached per host thread (cuTensorMapEncodeTiled costs ~1 us, several per launch).
struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, d2, s1, s2;
  uint32_t b0, b1, b2, kind;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b0 == o.b0 &&
           b1 == o.b1 && b2 == o.b2 && kind == o.kind;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.d0, k.d1, k.d2, k.s1, k.s2, (uint64_t)k.b0 << 32 | k.b1, (uint64_t)k.b2 << 32 | k.kind})
      h = (h ^ v) * 0x100000001B3ull;
    return static_cast<size_t>(h);
  }
};
static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapHash> g_tmap_cache;

static bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
static void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(k, m);
}

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  const TmapKey key{ptr, inner, outer, 0, row_stride_bytes, 0, box_inner, box_outer, 0, 1};
  if (tmap_lookup(key, m)) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_bytes & 15))
    return set_error(DCV_ERR_UNSUPPORTED, "TMA operand must be 16-byte aligned (ptr %p, row stride %llu)", ptr,
                     (unsigned long long)row_stride_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llux%llu box %ux%u) failed: %d",
                     (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer, (int)r);
  tmap_store(key, *m);
  return 0;
}

int make_tmap_2d(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer) {
  const TmapKey key{ptr, inner, outer, 0, row_stride_bytes, 0, box_inner, box_outer, 0, is_f32 ? 2u : 3u};
  if (tmap_lookup(key, m)) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_bytes & 15))
    return set_error(DCV_ERR_UNSUPPORTED, "TMA operand must be 16-byte aligned (ptr %p, row stride %llu)", ptr,
                     (unsigned long long)row_stride_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled(2d %s %llux%llu box %ux%u) failed: %d", is_f32 ? "f32" : "bf16",
                     (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer, (int)r);
  tmap_store(key, *m);
  return 0;
}

int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  const TmapKey key{ptr, d0, d1, d2, stride1_bytes, stride2_bytes, box0, box1, box2, 4};
  if (tmap_lookup(key, m)) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return set_error(DCV_ERR_UNSUPPORTED, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  tmap_store(key, *m);
  return 0;
}

int make_tmap_f32_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  const TmapKey key{ptr, d0, d1, d2, stride1_bytes, stride2_bytes, box0, box1, box2, 5};
  if (tmap_lookup(key, m)) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return set_error(DCV_ERR_UNSUPPORTED, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled(3d f32) failed: %d", (int)r);
  tmap_store(key, *m);
  return 0;
}

int make_tmap_f32_5d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t d4,
                     uint64_t s1, uint64_t s2, uint64_t s3, uint64_t s4, uint32_t b0, uint32_t b1, uint32_t b2,
                     uint32_t b3, uint32_t b4) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((s1 | s2 | s3 | s4) & 15))
    return set_error(DCV_ERR_UNSUPPORTED, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[5] = {d0, d1, d2, d3, d4};
  cuuint64_t strides[4] = {s1, s2, s3, s4};
  cuuint32_t box[5] = {b0, b1, b2, b3, b4};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DCV_ERR_CUDA, "cuTensorMapEncodeTiled(5d f32) failed: %d", (int)r);
  return 0;
}

}  // namespace dcv

using namespace dcv;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* dcv_last_error(void) { return g_err; }
int dcv_version(void) { return 1; }
long long dcv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dcv_gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue,
                const float* bias, void* out, void* out2, const float* resid, const void* aux, int ldo,
                void* stream) {
  if (!A || !B) return set_error(DCV_ERR_INVALID, "dcv_gemm_nt: null operand");
  return gemm_nt(A, lda, B, ldb, M, N, K, epilogue, bias, out, out2, resid, aux, ldo, false, ST(stream));
}

int dcv_gemm_nn(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, void* out,
                const void* aux, int ldo, void* stream) {
  if (!A || !B) return set_error(DCV_ERR_INVALID, "dcv_gemm_nn: null operand");
  return gemm_nt(A, lda, B, ldb, M, N, K, epilogue, nullptr, out, nullptr, nullptr, aux, ldo, true, ST(stream));
}

int dcv_gemm_nn_delta(const void* dY, int lda, const void* W, int ldb, int M, int N, int K, void* dO, const void* O,
                      float* delta, int L, void* stream) {
  if (!dY || !W || !dO || !O || !delta) return set_error(DCV_ERR_INVALID, "dcv_gemm_nn_delta: null pointer");
  return gemm_nt(dY, lda, W, ldb, M, N, K, 6 /* EPI_DELTA */, nullptr, dO, nullptr, nullptr, O, N, true, ST(stream), 0, 0,
                 nullptr, 0, delta, L);
}

int dcv_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
                int accumulate, int splits, void* stream) {
  if (!A || !B || !C) return set_error(DCV_ERR_INVALID, "dcv_gemm_tn: null operand");
  return gemm_tn(A, lda, B, ldb, M, Nout, Kout, C, ldc, accumulate, splits, ST(stream));
}

int dcv_attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, void* stream) {
  if (!qkv || !o || !lse2) return set_error(DCV_ERR_INVALID, "dcv_attn_fwd: null pointer");
  return attn_fwd(qkv, o, lse2, B, L, H, scale, ST(stream));
}

int dcv_attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
                 void* dqkv, int B, int L, int H, float scale, void* stream) {
  if (!qkv || !o || !dO || !lse2 || !delta || !dq_acc || !dqkv)
    return set_error(DCV_ERR_INVALID, "dcv_attn_bwd: null pointer");
  return attn_bwd(qkv, o, dO, lse2, delta, dq_acc, dqkv, B, L, H, scale, ST(stream));
}

int dcv_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M,
               int D, float eps, void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return set_error(DCV_ERR_INVALID, "dcv_ln_fwd: null pointer");
  return ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, ST(stream));
}

int dcv_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
               float* dres, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dres || !dx_bf16 || !dgamma || !dbeta)
    return set_error(DCV_ERR_INVALID, "dcv_ln_bwd: null pointer");
  return ln_bwd(dy, x, mean, rstd, gamma, dres, dx_bf16, dgamma, dbeta, dxsum, M, D, ST(stream));
}

int dcv_attn_bwd_fused(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
                       void* dqkv, float* dbias_qkv, int delta_ready, int B, int L, int H, float scale, void* stream) {
  if (!qkv || !o || !dO || !lse2 || !delta || !dq_acc || !dqkv)
    return set_error(DCV_ERR_INVALID, "dcv_attn_bwd_fused: null pointer");
  return attn_bwd(qkv, o, dO, lse2, delta, dq_acc, dqkv, B, L, H, scale, ST(stream), false, delta_ready != 0, dbias_qkv);
}

int dcv_colsum_bf16(const void* a, float* out, int M, int N, int lda, void* stream) {
  if (!a || !out) return set_error(DCV_ERR_INVALID, "dcv_colsum_bf16: null pointer");
  return colsum_bf16(a, out, M, N, lda, ST(stream));
}

int dcv_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (!src || !dst) return set_error(DCV_ERR_INVALID, "dcv_cast_f32_bf16: null pointer");
  return cast_f32_bf16(src, dst, n, ST(stream));
}

int dcv_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* clip, void* stream) {
  if (!p || !g || !m || !v) return set_error(DCV_ERR_INVALID, "dcv_adamw_step: null pointer");
  return adamw_step(p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, weight_decay, step, clip, ST(stream));
}

int dcv_optim_sched_step(dcv_optim_state* state, const dcv_sched* cfg, void* stream) {
  if (!state || !cfg) return set_error(DCV_ERR_INVALID, "dcv_optim_sched_step: null pointer");
  return optim_sched_step(state, *cfg, ST(stream));
}

int dcv_adamw_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float beta1, float beta2,
                       float eps, const dcv_optim_state* state, const float* clip, void* stream) {
  if (!p || !g || !m || !v || !state) return set_error(DCV_ERR_INVALID, "dcv_adamw_step_dev: null pointer");
  return adamw_step_dev(p, g, m, v, p_bf16, n, beta1, beta2, eps, state, clip, ST(stream));
}

int dcv_sumsq_f32(const float* g, long long n, float* out, void* stream) {
  if (!g || !out) return set_error(DCV_ERR_INVALID, "dcv_sumsq_f32: null pointer");
  return sumsq_f32(g, n, out, ST(stream));
}

int dcv_sgemm_small(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                    const float* bias, int accumulate, int M, int N, int K, void* stream) {
  if (!A || !B || !C) return set_error(DCV_ERR_INVALID, "dcv_sgemm_small: null pointer");
  return sgemm_small(A, lda, transA, B, ldb, transB, C, ldc, bias, accumulate, M, N, K, ST(stream));
}

int dcv_block_fwd(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a, void* stream) {
  if (!dims || !p || !a) return set_error(DCV_ERR_INVALID, "dcv_block_fwd: null struct");
  return block_fwd(*dims, *p, *a, ST(stream));
}

int dcv_block_bwd(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a,
                  const dcv_block_grads* g, const dcv_block_ws* ws, float* dres, void* dres_bf16,
                  float* dbias_prev, void* stream) {
  if (!dims || !p || !a || !g || !ws || !dres || !dres_bf16)
    return set_error(DCV_ERR_INVALID, "dcv_block_bwd: null struct / pointer");
  return block_bwd(*dims, *p, *a, *g, *ws, dres, dres_bf16, dbias_prev, ST(stream));
}

int dcv_block_fwd_cls(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a, void* stream) {
  if (!dims || !p || !a) return set_error(DCV_ERR_INVALID, "dcv_block_fwd_cls: null struct");
  return block_fwd_cls(*dims, *p, *a, ST(stream));
}

int dcv_block_bwd_cls(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a,
                      const dcv_block_grads* g, const dcv_block_ws* ws, float* dres_c, void* dres_c_bf16,
                      float* dres, void* dres_bf16, float* dbias_prev, void* stream) {
  if (!dims || !p || !a || !g || !ws || !dres_c || !dres_c_bf16 || !dres || !dres_bf16)
    return set_error(DCV_ERR_INVALID, "dcv_block_bwd_cls: null struct / pointer");
  return block_bwd_cls(*dims, *p, *a, *g, *ws, dres_c, dres_c_bf16, dres, dres_bf16, dbias_prev, ST(stream));
}

int dcv_embed_fwd(const dcv_embed_dims* dims, const dcv_embed_cfg* cfg, const dcv_embed_params* p, const void* x,
                  const int* idx, const int* gid, const dcv_embed_acts* a, void* stream) {
  if (!dims || !cfg || !p || !a) return set_error(DCV_ERR_INVALID, "dcv_embed_fwd: null struct");
  return embed_fwd(*dims, *cfg, *p, x, idx, gid, *a, ST(stream));
}

int dcv_embed_bwd(const dcv_embed_dims* dims, const dcv_embed_cfg* cfg, const dcv_embed_params* p, const int* gid,
                  const dcv_embed_acts* a, const dcv_embed_grads* g, const dcv_embed_ws* ws, const float* G,
                  const float* d_extra, void* stream) {
  if (!dims || !cfg || !p || !a || !g || !ws) return set_error(DCV_ERR_INVALID, "dcv_embed_bwd: null struct");
  return embed_bwd(*dims, *cfg, *p, gid, *a, *g, *ws, G, d_extra, ST(stream));
}

int dcv_head_fwd(const float* x_last, int B, int L, int D, const float* norm_w, const float* norm_b, float* feat,
                 float* mean, float* rstd, const float* head_w, const float* head_b, float* logits, int num_classes,
                 void* stream) {
  return head_fwd(x_last, B, L, D, norm_w, norm_b, feat, mean, rstd, head_w, head_b, logits, num_classes, ST(stream));
}

int dcv_head_bwd(const float* d_out, const float* x_last, int B, int L, int D, const float* norm_w,
                 const float* feat, const float* mean, const float* rstd, const float* head_w, int num_classes,
                 float* dfeat_ws, float* dres, void* dres_bf16, float* g_norm_w, float* g_norm_b, float* g_head_w,
                 float* g_head_b, float* dbias_last, void* stream) {
  return head_bwd(d_out, x_last, B, L, D, norm_w, feat, mean, rstd, head_w, num_classes, dfeat_ws, dres, dres_bf16,
                  g_norm_w, g_norm_b, g_head_w, g_head_b, dbias_last, ST(stream));
}

void dcv_debug_set_tn_desc(int lbo_bytes, int sbo_bytes) { debug_set_tn_desc(lbo_bytes, sbo_bytes); }

void dcv_debug_set_nt_cluster(int cm) { debug_set_nt_cluster(cm); }

int dcv_debug_attn_timeline(long long* buf) { return debug_attn_timeline(buf); }

void dcv_debug_set_pdl(int on) { debug_set_pdl(on); }
void dcv_debug_set_bwd_overlap(int on) { debug_set_bwd_overlap(on); }
void dcv_debug_set_embed_fused(int on) { debug_set_embed_fused(on); }
int dcv_debug_embed_timeline(long long* buf) { return debug_embed_timeline(buf); }

void dcv_debug_set_attn_mode(int fwd_mode, int bwd_mode) {
  if (fwd_mode >= 0) debug_set_attn_fwd_mode(fwd_mode);
  if (bwd_mode >= 0) debug_set_attn_bwd_mode(bwd_mode);
}

int dcv_profile_num_tags(void) { return PT_COUNT; }
const char* dcv_profile_tag_name(int tag) { return (tag >= 0 && tag < PT_COUNT) ? kProfNames[tag] : ""; }

int dcv_profile_start(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on.store(true);
  return 0;
}

int dcv_profile_stop(double* ms_by_tag, long long* launches_by_tag, int ntags) {
  g_prof_on.store(false);
  if (cudaDeviceSynchronize() != cudaSuccess) return set_error(DCV_ERR_CUDA, "dcv_profile_stop: device sync failed");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < ntags; ++i) {
    if (ms_by_tag) ms_by_tag[i] = 0.0;
    if (launches_by_tag) launches_by_tag[i] = 0;
  }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.tag < ntags) {
      if (ms_by_tag) ms_by_tag[r.tag] += ms;
      if (launches_by_tag) launches_by_tag[r.tag] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return 0;
}

}  // extern "C"
