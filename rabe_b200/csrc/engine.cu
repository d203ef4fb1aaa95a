// librabe_b200.so -- host side of the C ABI declared in include/rabe_b200.h: context, device
// scratch arena, host<->device staging, kernel launches.  No arithmetic happens on the host and
// there is no CPU fallback: every entry point fails with RB_ECUDA when CUDA is unavailable.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rabe_b200.h"
#include "kernels.cuh"
#include "coop_kernels.cuh"
#include "wide_kernels.cuh"
#include "kem_kernels.cuh"
#include "internal.h"

using namespace rb;

// ------------------------------------------------------------------------------------------
struct Block { char* p; size_t cap; };
struct Copyback { void* host; const void* dev; size_t bytes; };
struct ProfRec { const char* name; cudaEvent_t e0, e1; };

struct rb_ctx {
  int device;
  cudaStream_t own_stream, stream;
  int* d_err;
  int sticky;
  uint64_t launches;
  std::vector<Block> blocks;      // scratch arena (bump allocated, reset per API call)
  size_t cur, off;
  std::vector<Copyback> copybacks;
  bool host_io;                   // this call touched host buffers -> finish synchronously (unless async_host)
  bool must_sync;                 // this call builds a handle / returns a verdict: it completes before returning even in async mode
  cudaStream_t side[2];           // high-priority side streams: small kernels of a call overlap the big one
  cudaEvent_t ev_fork, ev_join[2];
  bool prof;                      // per-kernel CUDA-event timing (rb_ctx_profile)
  size_t rows_smem;               // dynamic shared memory reserved by k_ac17_enc_rows (occupancy cap, see rb_ac17_cp_encrypt_batch)
  int nest;                       // > 0 inside a fused scheme entry point: L0 calls share its arena and finish() once
  int pairing_mode;               // RB_PAIRING_AUTO / _THROUGHPUT / _LATENCY (rb_ctx_set_pairing_layout); 3 = hybrid (env only: two-lane Miller + six-lane final exp)
  cudaEvent_t ev_pair;            // recorded after the last launch of this context's latest call (the AUTO policy looks at the other contexts' events)
  bool ev_pair_used;
  bool async_host;                // host-buffer calls return after enqueueing (rb_ctx_set_async)
  int* h_err;                     // pinned copy of the device error flag for asynchronous host-buffer calls
  bool check_g2;                  // G2 inputs from the caller are tested for subgroup membership (rb_ctx_set_g2_subgroup_check)
  std::vector<ProfRec> prof_recs;
};

struct rb_table { rb_ctx* ctx; int kind; int W; int nwin; void* d; size_t bytes; size_t stride; };   // stride: entries per window (2^W, or 2^(W-1) + 32 for the signed-digit G1 tables)
struct rb_ac17_pk { rb_ctx* ctx; rb_table* g; rb_table* h_a[3]; rb_table* e[2]; };
struct rb_ac17_msk { rb_ctx* ctx; rb_table* g; rb_table* h; uint8_t* d_msk; Ac17MskConsts* consts; };
struct rb_msp { rb_ctx* ctx; uint32_t n1, n2; Fr* A; size_t n_pol; };   // n_pol > 1: one folded policy per batch item
struct rb_ac17_sk { rb_ctx* ctx; uint32_t n_k; uint8_t* d_k0; uint8_t* d_k; uint8_t* d_kp; MillerLine* lines; MillerLine* lines_unit; };   // lines_unit: every line divided by its l0 (null when some l0 is zero)
struct rb_share_plan { rb_ctx* ctx; uint32_t n_terms, n_leaves, n_coefs; ShareTerm* terms; uint32_t* leaf_offs; Fr* consts; };

enum { KIND_G1 = 1, KIND_G2 = 2, KIND_GT = 3 };

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { return (e_ == cudaErrorMemoryAllocation) ? RB_ENOMEM : RB_ECUDA; } } while (0)

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

namespace {

// Every live context of the process (the AUTO pairing-layout policy asks the OTHER contexts of the same device whether
// they have pairing work in flight).
std::mutex g_registry_mu;
std::vector<rb_ctx*> g_registry;
void registry_add(rb_ctx* c) { std::lock_guard<std::mutex> l(g_registry_mu); g_registry.push_back(c); }
void registry_remove(rb_ctx* c) {
  std::lock_guard<std::mutex> l(g_registry_mu);
  for (size_t i = 0; i < g_registry.size(); ++i) if (g_registry[i] == c) { g_registry.erase(g_registry.begin() + i); break; }
}
// bit 0: six-lane Miller kernels, bit 1: six-lane final exponentiation (wide.cuh); 0: the two-lane kernels (coop.cuh).
// AUTO: the six-lane kernels finish one batch sooner (everything in registers, three times the lanes per item); the
// two-lane kernels retire more batches per second once several are in flight (fewer instructions per product).  So a
// context takes the latency layout unless at least TWO other contexts of its device still have work (of any kind: an
// encrypt batch fills the SMs as well as a pairing batch does) queued or running.
int pairing_layout(rb_ctx* c) {
  if (c->pairing_mode == RB_PAIRING_THROUGHPUT) return 0;
  if (c->pairing_mode == RB_PAIRING_LATENCY) return 3;
  if (c->pairing_mode == 3) return 2;
  std::lock_guard<std::mutex> l(g_registry_mu);
  int busy = 0;                                          // ONE busy neighbour is usually this batch's own producer (encrypt -> decrypt)
  for (rb_ctx* o : g_registry) {
    if (o == c || o->device != c->device || !o->ev_pair_used) continue;
    const cudaError_t q = cudaEventQuery(o->ev_pair);
    cudaGetLastError();                                  // "not ready" is an answer, not an error to keep
    if (q == cudaErrorNotReady && ++busy >= 2) return 0;
  }
  return 3;
}
void pairing_mark(rb_ctx* c) { cudaEventRecord(c->ev_pair, c->stream); c->ev_pair_used = true; }

struct Guard {   // selects the context's device for the duration of a call
  int prev; bool ok;
  explicit Guard(rb_ctx* c) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (prev == c->device) || cudaSetDevice(c->device) == cudaSuccess;
  }
  ~Guard() { if (ok && prev >= 0) cudaSetDevice(prev); }
};

void arena_reset(rb_ctx* c) {
  c->copybacks.clear(); c->host_io = false;
  if (c->blocks.size() > 1) {          // consolidate: wait for users of the old blocks, then one big block
    cudaStreamSynchronize(c->stream);
    size_t total = 0;
    for (auto& b : c->blocks) { total += b.cap; cudaFree(b.p); }
    c->blocks.clear();
    char* p = nullptr;
    if (cudaMalloc(&p, total) == cudaSuccess) c->blocks.push_back({p, total});
  }
  c->cur = 0; c->off = 0;
}

// start of an API call: the scratch arena is recycled, except inside a fused scheme entry point
void begin_call(rb_ctx* c) {
  if (c->nest == 0) {
    cudaGetLastError();          // a stale, non-sticky runtime status (another library's cudaEventQuery ...) is not this call's failure
    arena_reset(c);
  }
}

void* arena_alloc(rb_ctx* c, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  if (c->cur < c->blocks.size() && c->off + bytes <= c->blocks[c->cur].cap) {
    void* r = c->blocks[c->cur].p + c->off; c->off += bytes; return r;
  }
  size_t cap = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
  char* p = nullptr;
  if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
  c->blocks.push_back({p, cap});
  c->cur = c->blocks.size() - 1; c->off = bytes;
  return p;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// input staging: device pointers pass through, host buffers are copied into the arena
template <class T> const T* stage_in(rb_ctx* c, const T* p, size_t bytes, int& st) {
  if (!p || bytes == 0) return p;
  if (is_device_ptr(p)) return p;
  void* d = arena_alloc(c, bytes);
  if (!d) { st = RB_ENOMEM; return nullptr; }
  if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = RB_ECUDA; return nullptr; }
  c->host_io = true;
  return static_cast<const T*>(d);
}
template <class T> T* stage_out(rb_ctx* c, T* p, size_t bytes, int& st) {
  if (!p || bytes == 0) return p;
  if (is_device_ptr(p)) return p;
  void* d = arena_alloc(c, bytes);
  if (!d) { st = RB_ENOMEM; return nullptr; }
  c->copybacks.push_back({p, d, bytes});
  c->host_io = true;
  return static_cast<T*>(d);
}

int map_flags(int flags) {
  if (flags & ERR_NOT_MEMBER) return RB_ENOTMEMBER;
  if (flags & ERR_POLICY) return RB_EPOLICY;
  return flags ? RB_EINVAL : RB_OK;
}

// end of an API call: copy results back / fetch the error flag when host buffers were involved
int finish(rb_ctx* c, int st) {
  if (c->nest > 0) return st;              // the enclosing entry point copies back / synchronises once
  pairing_mark(c);                         // any queued work of this context keeps the OTHER contexts of the device on the throughput layout
  if (st != RB_OK) { cudaStreamSynchronize(c->stream); return st; }
  { const cudaError_t le = cudaGetLastError(); if (le != cudaSuccess && le != cudaErrorNotReady) return RB_ECUDA; }
  if (!c->host_io) return RB_OK;
  for (auto& cb : c->copybacks)
    if (cudaMemcpyAsync(cb.host, cb.dev, cb.bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return RB_ECUDA;
  if (c->async_host && !c->must_sync) return RB_OK;          // results and status arrive with rb_ctx_sync() / rb_ctx_status()
  c->must_sync = false;
  int flags = 0;
  if (cudaMemcpyAsync(&flags, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return RB_ECUDA;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return RB_ECUDA;
  if (flags) { cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream); return map_flags(flags); }
  return RB_OK;
}

#define LAUNCH_ON(ctx, strm, kernel, grid, block, ...) do {                                          \
    ProfRec pr_{#kernel, nullptr, nullptr};                                                          \
    if ((ctx)->prof) { cudaEventCreate(&pr_.e0); cudaEventCreate(&pr_.e1); cudaEventRecord(pr_.e0, (strm)); } \
    kernel<<<(grid), (block), 0, (strm)>>>(__VA_ARGS__);                                              \
    (ctx)->launches++;                                                                                \
    if ((ctx)->prof) { cudaEventRecord(pr_.e1, (strm)); (ctx)->prof_recs.push_back(pr_); }             \
  } while (0)
#define LAUNCH(ctx, kernel, grid, block, ...) LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, __VA_ARGS__)

// Subgroup test of n caller-supplied G2 points (device bytes, `stride` apart): a failure raises the
// context's NOT_MEMBER flag, reported by finish() / rb_ctx_status() like every other membership error.
// Skipped inside fused entry points (their intermediates are the library's own) and when the caller
// declared its inputs trusted.
static void check_g2(rb_ctx* c, const uint8_t* d, size_t n, size_t stride = 128) {
  if (!c->check_g2 || c->nest > 0 || !d || n == 0) return;
  LAUNCH(c, k_g2_subgroup_check, grid_for(n, 64), 64, d, stride, n, c->d_err);
}

// product t = final exponentiation of the product of its Miller values (offs / fixed_count), times an optional Gt factor
static void launch_final_exp(rb_ctx* c, const Fp12* mil, const uint32_t* offs, uint32_t fixed_count, size_t n, const uint8_t* extra, uint8_t* out, int layout = -1) {
  if (layout < 0) layout = pairing_layout(c);
  pairing_mark(c);
  if (layout & 2) LAUNCH(c, k_final_exp_w6, w6_grid(n, RB_W6_BLOCK), RB_W6_BLOCK, mil, offs, fixed_count, n, extra, out, c->d_err);
  else LAUNCH(c, k_final_exp_co, grid_for(2 * n, RB_CO_FE_BLOCK), RB_CO_FE_BLOCK, mil, offs, fixed_count, n, extra, out, c->d_err);
}

// host-resident offset lists: non-decreasing and bounded by the index list they address
static bool offs_ok(const uint32_t* offs, size_t n_lists, size_t n_idx) {
  if (!offs || is_device_ptr(offs)) return true;
  for (size_t i = 0; i < n_lists; ++i) if (offs[i + 1] < offs[i]) return false;
  return offs[n_lists] <= n_idx;
}

// fork: side streams wait for everything enqueued so far on the main stream; join: the reverse.
static void fork_streams(rb_ctx* c) {
  cudaEventRecord(c->ev_fork, c->stream);
  for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(c->side[i], c->ev_fork, 0);
}
static void join_streams(rb_ctx* c) {
  for (int i = 0; i < 2; ++i) { cudaEventRecord(c->ev_join[i], c->side[i]); cudaStreamWaitEvent(c->stream, c->ev_join[i], 0); }
}

#ifndef RB_DEC_ITEM
#define RB_DEC_ITEM 0        // 1: the three decrypt terms of an item share one Miller accumulator (fewer, longer threads)
#endif
#ifndef RB_COOP_PAIRING
#define RB_COOP_PAIRING 1   // 0: one thread per Miller loop / final exponentiation (A/B comparisons)
#endif

#ifndef RB_C0_BLOCK
#define RB_C0_BLOCK 128      // threads per block of k_ac17_enc_c0
#endif
#ifndef RB_G1_M
#define RB_G1_M 24
#endif
constexpr int G1_M = RB_G1_M;   // most outputs per thread in the G1 fixed-base kernels (one inversion per thread)

// Outputs per thread for n outputs of a G1 fixed-base kernel: the grid is shaped to WHOLE waves of the kernel's resident
// threads -- 16 per thread by default, but e.g. the 786 432 outputs of a 4096 x 64-row encrypt are 1.3 waves at 16 (the
// second wave runs on a third of the SMs) and one wave at 21; a few thousand outputs are spread thin (2 per thread)
// instead of leaving most SMs idle behind 16-output threads.
template <class K> static int g1_outputs_per_thread(rb_ctx* c, K kernel, size_t smem, size_t n) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 2; }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || sms < 1) { cudaGetLastError(); sms = 148; }
  const size_t cap = (size_t)per_sm * sms * 128;                 // resident threads
  size_t waves = (n + 8 * cap) / (16 * cap);                     // round(n / (16 cap))
  if (waves < 1) waves = 1;
  size_t opt = (n + waves * cap - 1) / (waves * cap);
  if (opt < 2) opt = 2;
  if (opt > (size_t)G1_M) opt = G1_M;
  return (int)opt;
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

const char* rb_strerror(int s) {
  switch (s) {
    case RB_OK: return "ok";
    case RB_EINVAL: return "invalid argument";
    case RB_ENOTMEMBER: return "input is not a canonical field element / not on the curve";
    case RB_EPOLICY: return "inconsistent policy description";
    case RB_ECUDA: return "CUDA failure or no CUDA device";
    case RB_ENOMEM: return "out of device memory";
    default: return "unknown status";
  }
}
const char* rb_version(void) { return "rabe_b200 0.1 (sm_100a)"; }

int rb_ctx_create(int device, rb_ctx** out) {
  if (!out) return RB_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return RB_ECUDA; }
  CK(cudaSetDevice(device));
  rb_ctx* c = new (std::nothrow) rb_ctx();
  if (!c) return RB_ENOMEM;
  c->device = device; c->sticky = 0; c->launches = 0; c->cur = 0; c->off = 0; c->host_io = false; c->prof = false; c->nest = 0; c->check_g2 = true;
  c->async_host = false; c->must_sync = false; c->h_err = nullptr; c->ev_pair_used = false;
  {
    const char* e = getenv("RABE_B200_PAIRING");       // development override: co | w6 | hybrid
    c->pairing_mode = !e ? RB_PAIRING_AUTO : (strcmp(e, "co") == 0 ? RB_PAIRING_THROUGHPUT : (strcmp(e, "w6") == 0 ? RB_PAIRING_LATENCY : (strcmp(e, "hybrid") == 0 ? 3 : RB_PAIRING_AUTO)));
  }
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RB_ECUDA; }
  c->stream = c->own_stream;
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);       // hi = numerically lowest = greatest priority
    for (int i = 0; i < 2; ++i)
      if (cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, hi) != cudaSuccess) { delete c; return RB_ECUDA; }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
  }
  if (cudaMalloc(&c->d_err, sizeof(int)) != cudaSuccess || cudaMemset(c->d_err, 0, sizeof(int)) != cudaSuccess) { delete c; return RB_ECUDA; }
  cudaEventCreateWithFlags(&c->ev_pair, cudaEventDisableTiming);
  if (cudaMallocHost(&c->h_err, sizeof(int)) != cudaSuccess) { delete c; return RB_ECUDA; }
  *c->h_err = 0;
  registry_add(c);
  c->rows_smem = 0;
  if (const char* e = getenv("RABE_B200_ROWS_SMEM")) c->rows_smem = (size_t)atol(e);
  cudaFuncSetAttribute(k_ac17_enc_rows<G1_M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  {
    // The stack-heavy kernels keep Fq12 / point arrays in local memory: prefer the largest L1 over
    // shared memory they do not use (+3 % on the pipelined step; RABE_B200_MAX_L1=0 restores the default).
    const char* e = getenv("RABE_B200_MAX_L1");
    if (!e || atoi(e) != 0) {
      const int cv = cudaSharedmemCarveoutMaxL1;
      cudaFuncSetAttribute(k_ac17_dec_miller_pair_co, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_final_exp_co, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_miller_co, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_leaf_pair_co, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_leaf_fixed4_co, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_ac17_enc_rows<G1_M>, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_g1_mul_fixed<G1_M>, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_g1_gather_sum, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_ac17_enc_c0, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_g2_mul_fixed, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_gt_pow_fixed, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
      cudaFuncSetAttribute(k_gt_pow_var, cudaFuncAttributePreferredSharedMemoryCarveout, cv);
    }
  }
  cudaFuncSetAttribute(k_ac17_enc_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CP_PARTS * CP_ITEMS_PER_BLOCK * sizeof(Fp12)));
  // deep call chains (Fq12 routines are real functions): give local memory room
  cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
  // The fixed-base walks read 64-byte table entries at random addresses of a multi-GB table: ask L2 not to fetch
  // more than the 64 bytes a lookup uses (a hint; every other access pattern of the engine is compute-bound).
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
  cudaGetLastError();
  *out = c;
  return RB_OK;
}

void rb_ctx_destroy(rb_ctx* c) {
  if (!c) return;
  Guard g(c);
  registry_remove(c);
  cudaStreamSynchronize(c->stream);
  for (auto& b : c->blocks) cudaFree(b.p);
  cudaFree(c->d_err);
  cudaFreeHost(c->h_err);
  cudaEventDestroy(c->ev_pair);
  for (int i = 0; i < 2; ++i) { cudaStreamDestroy(c->side[i]); cudaEventDestroy(c->ev_join[i]); }
  cudaEventDestroy(c->ev_fork);
  cudaStreamDestroy(c->own_stream);
  delete c;
}

int rb_ctx_set_stream(rb_ctx* c, void* s) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  cudaStreamSynchronize(c->stream);
  c->stream = static_cast<cudaStream_t>(s);
  return RB_OK;
}
int rb_ctx_reset_stream(rb_ctx* c) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  cudaStreamSynchronize(c->stream);
  c->stream = c->own_stream;
  return RB_OK;
}
void* rb_ctx_get_stream(rb_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
int rb_ctx_sync(rb_ctx* c) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  CK(cudaStreamSynchronize(c->stream));
  return RB_OK;
}
int rb_ctx_status(rb_ctx* c) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  int flags = 0;
  CK(cudaMemcpyAsync(&flags, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (flags) cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
  return map_flags(flags);
}
uint64_t rb_ctx_launch_count(rb_ctx* c) { return c ? c->launches : 0; }
int rb_ctx_set_g2_subgroup_check(rb_ctx* c, int enable) {
  if (!c) return RB_EINVAL;
  c->check_g2 = enable != 0;
  return RB_OK;
}
int rb_g2_check_batch(rb_ctx* c, const uint8_t* q, size_t n) {
  if (!c || !q) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dq = stage_in(c, q, 128 * n, st);
  if (st == RB_OK) LAUNCH(c, k_g2_subgroup_check, grid_for(n, 64), 64, dq, (size_t)128, n, c->d_err);
  c->host_io = true; c->must_sync = true;   // a validation call always reports its verdict
  return finish(c, st);
}

int rb_ctx_profile(rb_ctx* c, int enable) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  c->prof_recs.clear();
  c->prof = enable != 0;
  return RB_OK;
}
int rb_ctx_profile_report(rb_ctx* c, char* out, size_t cap, size_t* needed) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  CK(cudaStreamSynchronize(c->stream));
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : c->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) return RB_ECUDA;
    std::string name(r.name);
    size_t lt = name.find('<');                   // "k_ac17_enc_rows<G1_M>" -> "k_ac17_enc_rows"
    if (lt != std::string::npos) name.resize(lt);
    auto& a = agg[name]; a.first += 1; a.second += ms;
  }
  std::string s = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %d, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    s += buf; first = false;
  }
  s += "}";
  if (needed) *needed = s.size() + 1;
  if (!out) return RB_OK;
  if (cap < s.size() + 1) return RB_EINVAL;
  memcpy(out, s.c_str(), s.size() + 1);
  return RB_OK;
}

// ---- element-wise -------------------------------------------------------------------------
static int fe_mul_batch(rb_ctx* c, bool fq, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 32 * n, st);
  const uint8_t* db = stage_in(c, b, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) {
    if (fq) LAUNCH(c, k_fe_mul<ModP>, grid_for(n, 128), 128, da, db, n, dout, c->d_err);
    else LAUNCH(c, k_fe_mul<ModR>, grid_for(n, 128), 128, da, db, n, dout, c->d_err);
  }
  return finish(c, st);
}
int rb_fq_mul_batch(rb_ctx* c, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) { return fe_mul_batch(c, true, a, b, n, out); }
int rb_fr_mul_batch(rb_ctx* c, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) { return fe_mul_batch(c, false, a, b, n, out); }

int rb_fq_mul_chain(rb_ctx* c, const uint8_t* a, const uint8_t* b, size_t n, int iters, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 32 * n, st);
  const uint8_t* db = stage_in(c, b, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) {
    // iters < 0 selects the single-chain variant (ILP 1), iters >= 2^20 the 4-chain variant: microbench only
    if (iters < 0) LAUNCH(c, k_fq_mul_chain<1>, grid_for(n, 128), 128, da, db, n, -iters, dout);
    else if (iters >= (1 << 20)) LAUNCH(c, k_fq_mul_chain<4>, grid_for(n, 128), 128, da, db, n, iters - (1 << 20), dout);
    else LAUNCH(c, k_fq_mul_chain<2>, grid_for(n, 128), 128, da, db, n, iters, dout);
  }
  return finish(c, st);
}

// ---- tables ---------------------------------------------------------------------------------
static int table_create(rb_ctx* c, int kind, const uint8_t* base, int W, rb_table** out) {
  if (!c || !base || !out) return RB_EINVAL;
  *out = nullptr;
  int maxw = (kind == KIND_G1) ? 26 : 16;         // G1: 2^24 x 11 windows x 64 B = 11.8 GB; 2^26 x 10 windows = 42.9 GB
  if (W < 4 || W > maxw) return RB_EINVAL;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int nwin = (256 + W - 1) / W;
  size_t esz = (kind == KIND_G1) ? sizeof(G1Affine) : (kind == KIND_G2 ? sizeof(G2Affine) : sizeof(Fp12));
  // G1 tables wider than 12 bits keep signed digits: d = 0 .. 2^(W-1) per window (fixed_base_mul_signed), half the memory
  const bool sgn = kind == KIND_G1 && W > 12;
  const size_t stride = sgn ? (((size_t)1 << (W - 1)) + TABLE_CHUNK) : ((size_t)1 << W);
  size_t entries = (size_t)nwin * stride;
  rb_table* t = new (std::nothrow) rb_table();
  if (!t) return RB_ENOMEM;
  t->ctx = c; t->kind = kind; t->W = W; t->nwin = nwin; t->bytes = entries * esz; t->d = nullptr; t->stride = stride;
  if (cudaMalloc(&t->d, t->bytes) != cudaSuccess) { delete t; cudaGetLastError(); return RB_ENOMEM; }
  int st = RB_OK;
  size_t bsz = (kind == KIND_G1) ? 64 : (kind == KIND_G2 ? 128 : 384);
  const uint8_t* dbase = stage_in(c, base, bsz, st);
  // decode + validate the base on the device with a one-thread kernel, then build
  if (st == RB_OK) {
    if (kind == KIND_G1) {
      G1Affine* tmp = (G1Affine*)arena_alloc(c, sizeof(G1Affine));
      if (!tmp) st = RB_ENOMEM;
      else {
        GatherArgs ga{dbase, nullptr, nullptr, 0, 1, 1, 0, dbase, 0};   // "sum" of just the extra point = decode
        LAUNCH(c, k_g1_gather_sum, 1, 32, ga, (size_t)1, tmp, (uint8_t*)nullptr, c->d_err);
        G1Affine hb;
        if (cudaMemcpyAsync(&hb, tmp, sizeof hb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) st = RB_ECUDA;
        else {
          LAUNCH(c, k_table_window_bases<Fp>, grid_for(nwin, 32), 32, hb, W, nwin, (G1Affine*)t->d, stride);
          if (W > 12) LAUNCH(c, k_table_fill_chunked<Fp>, grid_for(entries / TABLE_CHUNK, 64), 64, W, nwin, (G1Affine*)t->d, stride, stride);
          else LAUNCH(c, k_table_fill<Fp>, grid_for(entries, 128), 128, W, nwin, (G1Affine*)t->d);
        }
      }
    } else if (kind == KIND_G2) {
      uint8_t* tmpb = (uint8_t*)arena_alloc(c, 128 + sizeof(G2Affine));
      if (!tmpb) st = RB_ENOMEM;
      else {
        G2Affine* tmp = (G2Affine*)(tmpb + 128);
        check_g2(c, dbase, 1);
        LAUNCH(c, k_decode_g2, 1, 32, dbase, tmp, c->d_err);
        G2Affine hb;
        if (cudaMemcpyAsync(&hb, tmp, sizeof hb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) st = RB_ECUDA;
        else {
          LAUNCH(c, k_table_window_bases<Fp2>, grid_for(nwin, 32), 32, hb, W, nwin, (G2Affine*)t->d, stride);
          if (W > 12) LAUNCH(c, k_table_fill_chunked<Fp2>, grid_for(entries / TABLE_CHUNK, 64), 64, W, nwin, (G2Affine*)t->d, stride, stride);
          else LAUNCH(c, k_table_fill<Fp2>, grid_for(entries, 64), 64, W, nwin, (G2Affine*)t->d);
        }
      }
    } else {
      Fp12* tmp = (Fp12*)arena_alloc(c, sizeof(Fp12));
      if (!tmp) st = RB_ENOMEM;
      else {
        LAUNCH(c, k_decode_gt, 1, 32, dbase, tmp, c->d_err);
        LAUNCH(c, k_gt_table_window_bases, grid_for(nwin, 32), 32, tmp, W, nwin, (Fp12*)t->d);
        if (W > 12) LAUNCH(c, k_gt_table_fill_chunked, grid_for(entries / TABLE_CHUNK, 64), 64, W, nwin, (Fp12*)t->d);
        else LAUNCH(c, k_gt_table_fill, grid_for(entries, 64), 64, W, nwin, (Fp12*)t->d);
      }
    }
  }
  c->host_io = true; c->must_sync = true;   // table construction always completes before returning
  st = finish(c, st);
  if (st != RB_OK) { cudaFree(t->d); delete t; return st; }
  *out = t;
  return RB_OK;
}

int rb_g1_table_create(rb_ctx* c, const uint8_t* base, int W, rb_table** out) { return table_create(c, KIND_G1, base, W, out); }
int rb_g2_table_create(rb_ctx* c, const uint8_t* base, int W, rb_table** out) { return table_create(c, KIND_G2, base, W, out); }
int rb_gt_table_create(rb_ctx* c, const uint8_t* base, int W, rb_table** out) { return table_create(c, KIND_GT, base, W, out); }
void rb_table_destroy(rb_table* t) {
  if (!t) return;
  Guard g(t->ctx);
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->d);
  delete t;
}

// ---- fixed / variable base batches -----------------------------------------------------------
int rb_g1_mul_fixed_batch(rb_ctx* c, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out) {
  if (!c || !t || t->kind != KIND_G1 || !k || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dk = stage_in(c, k, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 64 * n, st);
  if (st == RB_OK) {
    const int opt = g1_outputs_per_thread(c, k_g1_mul_fixed<G1_M>, 0, n);
    size_t threads = (n + opt - 1) / opt;
    LAUNCH(c, k_g1_mul_fixed<G1_M>, grid_for(threads, 128), 128, (const G1Affine*)t->d, t->W, t->nwin, dk, n, dout, c->d_err, t->stride, opt);
  }
  return finish(c, st);
}
int rb_g2_mul_fixed_batch(rb_ctx* c, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out) {
  if (!c || !t || t->kind != KIND_G2 || !k || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dk = stage_in(c, k, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 128 * n, st);
  if (st == RB_OK) LAUNCH(c, k_g2_mul_fixed, grid_for(n, 128), 128, (const G2Affine*)t->d, TabSel{0, nullptr}, t->W, t->nwin, dk, n, dout, c->d_err);
  return finish(c, st);
}
int rb_gt_pow_fixed_batch(rb_ctx* c, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out) {
  if (!c || !t || t->kind != KIND_GT || !k || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dk = stage_in(c, k, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 384 * n, st);
  if (st == RB_OK) LAUNCH(c, k_gt_pow_fixed, grid_for(n, 64), 64, (const Fp12*)t->d, TabSel{0, nullptr}, t->W, t->nwin, dk, n, dout, c->d_err);
  return finish(c, st);
}

// element-wise binary operators; the operands of element i are a[(i / ai.div) % ai.mod], k[...] (OpIdx)
#define BINARY_EX(NAME, KERNEL, ABYTES, KBYTES, OBYTES, BLOCK)                                                          \
  static int NAME(rb_ctx* c, const uint8_t* a, OpIdx ai, size_t a_count, const uint8_t* k, OpIdx ki, size_t k_count, size_t n, uint8_t* out) { \
    if (!c || !a || !k || !out) return RB_EINVAL;                                                                        \
    if (n == 0) return RB_OK;                                                                                             \
    Guard g(c); if (!g.ok) return RB_ECUDA;                                                                               \
    begin_call(c);                                                                                                        \
    int st = RB_OK;                                                                                                       \
    const uint8_t* da = stage_in(c, a, (size_t)(ABYTES) * a_count, st);                                                   \
    const uint8_t* dk = stage_in(c, k, (size_t)(KBYTES) * k_count, st);                                                   \
    uint8_t* dout = stage_out(c, out, (size_t)(OBYTES) * n, st);                                                          \
    if (st == RB_OK) LAUNCH(c, KERNEL, grid_for(n, BLOCK), BLOCK, da, ai, dk, ki, n, dout, c->d_err);                     \
    return finish(c, st);                                                                                                 \
  }
BINARY_EX(g1_mul_var_ex, k_g1_mul_var, 64, 32, 64, 128)
static int g2_mul_var_ex(rb_ctx* c, const uint8_t* a, OpIdx ai, size_t a_count, const uint8_t* k, OpIdx ki, size_t k_count, size_t n, uint8_t* out) {
  if (!c || !a || !k || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, (size_t)128 * a_count, st);
  const uint8_t* dk = stage_in(c, k, (size_t)32 * k_count, st);
  uint8_t* dout = stage_out(c, out, (size_t)128 * n, st);
  if (st == RB_OK) { check_g2(c, da, a_count); LAUNCH(c, k_g2_mul_var, grid_for(n, 128), 128, da, ai, dk, ki, n, dout, c->d_err); }
  return finish(c, st);
}
BINARY_EX(gt_pow_var_ex, k_gt_pow_var, 384, 32, 384, 64)
static const OpIdx EACH = {1, 0};
int rb_g1_mul_var_batch(rb_ctx* c, const uint8_t* a, const uint8_t* k, size_t n, uint8_t* out) { return g1_mul_var_ex(c, a, EACH, n, k, EACH, n, n, out); }
int rb_g2_mul_var_batch(rb_ctx* c, const uint8_t* a, const uint8_t* k, size_t n, uint8_t* out) { return g2_mul_var_ex(c, a, EACH, n, k, EACH, n, n, out); }
int rb_gt_pow_var_batch(rb_ctx* c, const uint8_t* a, const uint8_t* k, size_t n, uint8_t* out) { return gt_pow_var_ex(c, a, EACH, n, k, EACH, n, n, out); }
static int gt_mul_ex(rb_ctx* c, const uint8_t* a, const uint8_t* b, OpIdx bi, size_t b_count, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 384 * n, st);
  const uint8_t* db = stage_in(c, b, 384 * b_count, st);
  uint8_t* dout = stage_out(c, out, 384 * n, st);
  if (st == RB_OK) LAUNCH(c, k_gt_mul, grid_for(n, 64), 64, da, db, bi, n, dout, c->d_err);
  return finish(c, st);
}
int rb_gt_mul_batch(rb_ctx* c, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) { return gt_mul_ex(c, a, b, EACH, n, n, out); }

int rb_gt_inverse_batch(rb_ctx* c, const uint8_t* a, size_t n, uint8_t* out) {
  if (!c || !a || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 384 * n, st);
  uint8_t* dout = stage_out(c, out, 384 * n, st);
  if (st == RB_OK) LAUNCH(c, k_gt_inverse, grid_for(n, 64), 64, da, n, dout, c->d_err);
  return finish(c, st);
}

int rb_g1_sum_gather_batch(rb_ctx* c, const uint8_t* points, size_t n_points, const uint32_t* idx, const uint32_t* offs,
                           size_t n_out, uint8_t* out) {
  if (!c || !points || !offs || !out || (!idx && n_points)) return RB_EINVAL;
  if (n_out == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  // the offsets live on the host or the device; the list length is offs[n_out]
  uint32_t total = 0;
  if (is_device_ptr(offs)) { CK(cudaMemcpyAsync(&total, offs + n_out, 4, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
  else {
    for (size_t i = 0; i < n_out; ++i) if (offs[i + 1] < offs[i]) return RB_EINVAL;
    total = offs[n_out];
    if (idx && !is_device_ptr(idx)) for (uint32_t i = 0; i < total; ++i) if (idx[i] >= n_points) return RB_EINVAL;
  }
  const uint8_t* dp = stage_in(c, points, 64 * n_points, st);
  const uint32_t* didx = stage_in(c, idx, 4 * (size_t)total, st);
  const uint32_t* doffs = stage_in(c, offs, 4 * (n_out + 1), st);
  uint8_t* dout = stage_out(c, out, 64 * n_out, st);
  if (st == RB_OK) {
    GatherArgs ga{dp, didx, doffs, total, 1, 1, 0, nullptr, 0};
    LAUNCH(c, k_g1_gather_sum, grid_for(n_out, 128), 128, ga, n_out, (G1Affine*)nullptr, dout, c->d_err);
  }
  return finish(c, st);
}

int rb_pairing_product_batch(rb_ctx* c, const uint8_t* P, const uint8_t* Q, const uint32_t* offs, size_t n_products, uint8_t* out) {
  if (!c || !P || !Q || !offs || !out) return RB_EINVAL;
  if (n_products == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  uint32_t total = 0;
  if (is_device_ptr(offs)) { CK(cudaMemcpyAsync(&total, offs + n_products, 4, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
  else {
    for (size_t i = 0; i < n_products; ++i) if (offs[i + 1] < offs[i]) return RB_EINVAL;
    total = offs[n_products];
  }
  const uint8_t* dP = stage_in(c, P, 64 * (size_t)total, st);
  const uint8_t* dQ = stage_in(c, Q, 128 * (size_t)total, st);
  const uint32_t* doffs = stage_in(c, offs, 4 * (n_products + 1), st);
  uint8_t* dout = stage_out(c, out, 384 * n_products, st);
  Fp12* mil = (Fp12*)arena_alloc(c, sizeof(Fp12) * (size_t)(total ? total : 1));
  if (!mil) st = RB_ENOMEM;
  if (st == RB_OK) {
    MillerArgs ma{nullptr, dP, dQ, 0, nullptr, nullptr};
    check_g2(c, dQ, total);
#if RB_COOP_PAIRING
    if (total) LAUNCH(c, k_miller_co, grid_for(2 * (size_t)total, RB_CO_BLOCK), RB_CO_BLOCK, ma, (size_t)total, mil, c->d_err);
    launch_final_exp(c, mil, doffs, 0u, n_products, nullptr, dout);
#else
    if (total) LAUNCH(c, k_miller, grid_for(total, RB_ML_BLOCK), RB_ML_BLOCK, ma, (size_t)total, mil, c->d_err);
    LAUNCH(c, k_final_exp, grid_for(n_products, RB_FE_BLOCK), RB_FE_BLOCK, mil, doffs, 0u, n_products, (const uint8_t*)nullptr, dout, c->d_err);
#endif
  }
  return finish(c, st);
}

// ---- AC17 -----------------------------------------------------------------------------------
void rb_ac17_pk_free(rb_ac17_pk* pk) {
  if (!pk) return;
  rb_table_destroy(pk->g);
  for (int i = 0; i < 3; ++i) rb_table_destroy(pk->h_a[i]);
  for (int i = 0; i < 2; ++i) rb_table_destroy(pk->e[i]);
  delete pk;
}
int rb_ac17_pk_load_ex(rb_ctx* c, const uint8_t* pkb, int g1_window, int g2_window, int gt_window, rb_ac17_pk** out) {
  if (!c || !pkb || !out) return RB_EINVAL;
  *out = nullptr;
  uint8_t host[RB_AC17_PK_BYTES];
  if (is_device_ptr(pkb)) { Guard g(c); CK(cudaMemcpy(host, pkb, sizeof host, cudaMemcpyDeviceToHost)); }
  else memcpy(host, pkb, sizeof host);
  rb_ac17_pk* pk = new (std::nothrow) rb_ac17_pk();
  if (!pk) return RB_ENOMEM;
  pk->ctx = c;
  int st = rb_g1_table_create(c, host, g1_window, &pk->g);
  for (int i = 0; i < 3 && st == RB_OK; ++i) st = rb_g2_table_create(c, host + 64 + 128 * i, g2_window, &pk->h_a[i]);
  for (int i = 0; i < 2 && st == RB_OK; ++i) st = rb_gt_table_create(c, host + 448 + 384 * i, gt_window, &pk->e[i]);
  if (st != RB_OK) { rb_ac17_pk_free(pk); return st; }
  *out = pk;
  return RB_OK;
}
int rb_ac17_pk_load(rb_ctx* c, const uint8_t* pkb, rb_ac17_pk** out) { return rb_ac17_pk_load_ex(c, pkb, 16, 8, 8, out); }

void rb_msp_free(rb_msp* m) {
  if (!m) return;
  Guard g(m->ctx);
  cudaStreamSynchronize(m->ctx->stream);
  cudaFree(m->A);
  delete m;
}
int rb_msp_load_batch(rb_ctx* c, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row, const uint8_t* h_col, size_t n_pol, rb_msp** out) {
  if (!c || !m || !h_row || !h_col || !out) return RB_EINVAL;
  *out = nullptr;
  if (n1 == 0 || n2 == 0 || n_pol == 0) return RB_EPOLICY;
  if (!is_device_ptr(m)) for (size_t i = 0; i < n_pol * n1 * n2; ++i) if (m[i] < -1 || m[i] > 1) return RB_EPOLICY;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  rb_msp* p = new (std::nothrow) rb_msp();
  if (!p) return RB_ENOMEM;
  p->ctx = c; p->n1 = n1; p->n2 = n2; p->A = nullptr; p->n_pol = n_pol;
  if (cudaMalloc(&p->A, sizeof(Fr) * n_pol * n1 * 6) != cudaSuccess) { delete p; cudaGetLastError(); return RB_ENOMEM; }
  int st = RB_OK;
  const int8_t* dm = stage_in(c, m, n_pol * n1 * n2, st);
  const uint8_t* dhr = stage_in(c, h_row, n_pol * n1 * 6 * 32, st);
  const uint8_t* dhc = stage_in(c, h_col, n_pol * n2 * 6 * 32, st);
  if (st == RB_OK) LAUNCH(c, k_ac17_fold_msp, grid_for(n_pol * n1 * 6, 128), 128, n1, n2, dm, dhr, dhc, p->A, c->d_err, n_pol);
  c->host_io = true; c->must_sync = true;
  st = finish(c, st);
  if (st != RB_OK) { cudaFree(p->A); delete p; return st; }
  *out = p;
  return RB_OK;
}
// Refolds a loaded handle in place from new matrices / hashes of the same shape: no allocation, stream-ordered.
int rb_msp_reload_batch(rb_ctx* c, rb_msp* p, const int8_t* m, const uint8_t* h_row, const uint8_t* h_col, int h_col_shared) {
  if (!c || !p || !m || !h_row || !h_col) return RB_EINVAL;
  const uint32_t n1 = p->n1, n2 = p->n2;
  const size_t n_pol = p->n_pol;
  // (matrix entries outside {-1, 0, 1} are flagged by the fold kernel -> RB_EPOLICY with the call's status: this entry point
  //  sits inside the encrypt step of per-item-policy batches and stays free of per-byte host work)
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const int8_t* dm = stage_in(c, m, n_pol * n1 * n2, st);
  const uint8_t* dhr = stage_in(c, h_row, n_pol * n1 * 6 * 32, st);
  const uint8_t* dhc = stage_in(c, h_col, (h_col_shared ? 1 : n_pol) * n2 * 6 * 32, st);
  if (st == RB_OK) LAUNCH(c, k_ac17_fold_msp, grid_for(n_pol * n1 * 6, 128), 128, n1, n2, dm, dhr, dhc, p->A, c->d_err, n_pol, h_col_shared ? 1 : 0);
  if (!is_device_ptr(m) || !is_device_ptr(h_row) || !is_device_ptr(h_col)) c->host_io = true;
  return finish(c, st);
}
int rb_msp_load(rb_ctx* c, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row, const uint8_t* h_col, rb_msp** out) {
  return rb_msp_load_batch(c, n1, n2, m, h_row, h_col, 1, out);
}

int rb_ac17_cp_encrypt_batch(rb_ctx* c, const rb_ac17_pk* pk, const rb_msp* msp, const uint8_t* s, const uint8_t* msg, size_t B,
                             uint8_t* c_0, uint8_t* cc, uint8_t* c_p) {
  if (!c || !pk || !msp || !s || !msg || !c_0 || !cc || !c_p) return RB_EINVAL;
  if (msp->n_pol > 1 && msp->n_pol != B) return RB_EPOLICY;      // per-item policies: one per batch item
  if (B == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint32_t rows3 = msp->n1 * 3;
  const size_t total = B * rows3;
  const uint8_t* ds = stage_in(c, s, 64 * B, st);
  const uint8_t* dmsg = stage_in(c, msg, 384 * B, st);
  uint8_t* dc0 = stage_out(c, c_0, 384 * B, st);
  uint8_t* dcc = stage_out(c, cc, 64 * total, st);
  uint8_t* dcp = stage_out(c, c_p, 384 * B, st);
  if (st == RB_OK) {
    // the two small, latency-bound kernels run on high-priority side streams under the big one
    fork_streams(c);
    {
      ProfRec pr_{"k_ac17_enc_cp", nullptr, nullptr};
      if (c->prof) { cudaEventCreate(&pr_.e0); cudaEventCreate(&pr_.e1); cudaEventRecord(pr_.e0, c->side[0]); }
      const int threads_cp = CP_PARTS * CP_ITEMS_PER_BLOCK;
      k_ac17_enc_cp<<<grid_for(B, CP_ITEMS_PER_BLOCK), threads_cp, threads_cp * sizeof(Fp12), c->side[0]>>>(
          (const Fp12*)pk->e[0]->d, (const Fp12*)pk->e[1]->d, pk->e[0]->W, pk->e[0]->nwin, ds, dmsg, B, dcp, c->d_err);
      c->launches++;
      if (c->prof) { cudaEventRecord(pr_.e1, c->side[0]); c->prof_recs.push_back(pr_); }
    }
    G2Tab3 tabs{{(const G2Affine*)pk->h_a[0]->d, (const G2Affine*)pk->h_a[1]->d, (const G2Affine*)pk->h_a[2]->d}};
    LAUNCH_ON(c, c->side[1], k_ac17_enc_c0, grid_for(3 * B, RB_C0_BLOCK), RB_C0_BLOCK, tabs, pk->h_a[0]->W, pk->h_a[0]->nwin, ds, B, dc0, c->d_err);
    const int opt = g1_outputs_per_thread(c, k_ac17_enc_rows<G1_M>, c->rows_smem, total);
    size_t threads = (total + opt - 1) / opt;
    {
      // Occupancy knob: dynamic shared memory the kernel never touches caps its resident blocks per
      // SM, leaving registers for the decrypt kernels of other batches that share the SM (the
      // pipeline does better with heterogeneous residents than with this kernel alone at full
      // occupancy).  RABE_B200_ROWS_SMEM overrides (bytes; development aid).
      ProfRec pr_{"k_ac17_enc_rows", nullptr, nullptr};
      if (c->prof) { cudaEventCreate(&pr_.e0); cudaEventCreate(&pr_.e1); cudaEventRecord(pr_.e0, c->stream); }
      k_ac17_enc_rows<G1_M><<<grid_for(threads, 128), 128, c->rows_smem, c->stream>>>((const G1Affine*)pk->g->d, pk->g->W, pk->g->nwin, msp->A, ds,
                                                                                      rows3, total, dcc, c->d_err,
                                                                                      msp->n_pol > 1 ? (size_t)rows3 * 2 : (size_t)0, pk->g->stride, opt);
      c->launches++;
      if (c->prof) { cudaEventRecord(pr_.e1, c->stream); c->prof_recs.push_back(pr_); }
    }
    join_streams(c);
  }
  return finish(c, st);
}

// thread (b, j), j < 3: the two pairs of decrypt term j -- e(-(k_p[j] + prod_h_j), c_0[b][j]) with a
// variable second argument and e(prod_g_j, k_0[j]) with a fixed one (precomputed lines) -- share one
// Miller accumulator (miller_pair): 3 Miller values per item instead of 6.
__global__ void __launch_bounds__(RB_ML_BLOCK, RB_PAIR_MINB) k_ac17_dec_miller_pair(const G1Affine* __restrict__ ph, int ph_per_item, const G1Affine* __restrict__ pg,
                                                              const uint8_t* __restrict__ c_0, const MillerLine* __restrict__ lines, size_t B,
                                                              Fp12* out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * B) return;
  size_t b = t / 3; int j = (int)(t % 3);
  Fp12 f;
  G1Affine pv = ph[(ph_per_item ? 3 * b : 0) + j];
  G2Affine q = load_g2_checked(c_0 + 128 * t, err);
  G1Affine pf = pg[t];
  const MillerLine* lj = lines + (size_t)j * MILLER_LINES;
  const bool hv = !(aff_is_inf(pv) || aff_is_inf(q)), hf = !aff_is_inf(pf);
  if (hv && hf) miller_pair(&f, &pv, &q, &pf, lj);
  else if (hv) miller_single(&f, &pv, &q);
  else if (hf) miller_fixed(&f, &pf, lj);
  else fp12_set_one(f);
  out[t] = f;
}
__global__ void k_miller_lines(const uint8_t* __restrict__ q_bytes, int n, MillerLine* lines, int* err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine q = load_g2_checked(q_bytes + 128 * i, err);
  if (aff_is_inf(q)) { flag_error(err, ERR_NOT_MEMBER); return; }       // an infinite k_0 member cannot come from keygen
  miller_lines_for(lines + (size_t)i * MILLER_LINES, &q);
}

// table t of a loaded key: out = in with every line divided by its l0 (one thread per table; once per key).  *degenerate is
// raised when some l0 is zero -- the handle then keeps only the general table.
__global__ void k_lines_normalize(const MillerLine* __restrict__ in, MillerLine* out, int n_tables, int* degenerate) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tables) return;
  if (!miller_lines_normalize(out + (size_t)t * MILLER_LINES, in + (size_t)t * MILLER_LINES, MILLER_LINES)) atomicOr(degenerate, 1);
}

static int ac17_decrypt_common(rb_ctx* c, const uint8_t* dk0, const uint8_t* dk, uint32_t n_k, const uint8_t* dkp, const MillerLine* lines, bool unit_lines,
                               const uint8_t* c_0, const uint8_t* cc, uint32_t n1, const uint8_t* c_p, size_t B, const uint32_t* ct_idx,
                               const uint32_t* ct_offs, size_t n_ct_idx, const uint32_t* sk_idx, const uint32_t* sk_offs, size_t n_sk_idx,
                               uint8_t* msg_out) {
  int st = RB_OK;
  const uint8_t* dc0 = stage_in(c, c_0, 384 * B, st);
  const uint8_t* dcc = stage_in(c, cc, 192 * (size_t)n1 * B, st);
  const uint8_t* dcp = stage_in(c, c_p, 384 * B, st);
  const uint32_t* dci = stage_in(c, ct_idx, 4 * n_ct_idx, st);
  const uint32_t* dco = stage_in(c, ct_offs, 4 * (B + 1), st);
  const uint32_t* dsi = stage_in(c, sk_idx, 4 * n_sk_idx, st);
  const uint32_t* dso = stage_in(c, sk_offs, 4 * (B + 1), st);
  uint8_t* dout = stage_out(c, msg_out, 384 * B, st);
  size_t n_h = sk_offs ? B : 1;
  G1Affine* ph = (G1Affine*)arena_alloc(c, sizeof(G1Affine) * 3 * n_h);
  G1Affine* pg = (G1Affine*)arena_alloc(c, sizeof(G1Affine) * 3 * B);
  Fp12* mil = (Fp12*)arena_alloc(c, sizeof(Fp12) * 3 * B);
  if (!ph || !pg || !mil) st = RB_ENOMEM;
  if (st == RB_OK) {
    check_g2(c, dc0, 3 * B);                // c_0 comes from the ciphertext: untrusted unless the caller says otherwise
    // the key-side sums (3 threads when the whole batch shares one pruned list) run beside the ciphertext-side sums
    fork_streams(c);
    GatherArgs gh{dk, dsi, dso, (uint32_t)n_sk_idx, 1, 3, 0, dkp, 1};
    LAUNCH_ON(c, c->side[0], k_g1_gather_sum, grid_for(3 * n_h, 128), 128, gh, n_h, ph, (uint8_t*)nullptr, c->d_err);
    GatherArgs gg{dcc, dci, dco, (uint32_t)n_ct_idx, 1, 3, (size_t)n1 * 3, nullptr, 0};
    LAUNCH(c, k_g1_gather_sum, grid_for(3 * B, 128), 128, gg, B, pg, (uint8_t*)nullptr, c->d_err);
    join_streams(c);
    if (!lines) {
      // no loaded key handle: the line tables of k_0 are built for this call (rb_ac17_sk_load keeps them)
      MillerLine* tmp = (MillerLine*)arena_alloc(c, sizeof(MillerLine) * 3 * MILLER_LINES);
      if (!tmp) return finish(c, RB_ENOMEM);
      LAUNCH(c, k_miller_lines, 1, 32, dk0, 3, tmp, c->d_err);
      lines = tmp; unit_lines = false;
    }
#if RB_COOP_PAIRING
    const int layout = pairing_layout(c);
    if (layout & 1) {
      // six lanes per ciphertext: its three terms on one accumulator, everything in registers (wide.cuh)
      LAUNCH(c, k_ac17_dec_item_w6, w6_grid(B, RB_W6_BLOCK), RB_W6_BLOCK, ph, sk_offs ? 1 : 0, pg, dc0, lines, unit_lines ? 1 : 0, B, mil, c->d_err);
      launch_final_exp(c, mil, nullptr, 1u, B, dcp, dout, layout);
      return finish(c, st);
    }
    // two threads per Miller loop / final exponentiation (coop.cuh)
#if RB_DEC_ITEM
    LAUNCH(c, k_ac17_dec_miller_item_co, grid_for(2 * B, RB_CO_BLOCK), RB_CO_BLOCK, ph, sk_offs ? 1 : 0, pg, dc0, lines, B, mil, c->d_err);
    LAUNCH(c, k_final_exp_co, grid_for(2 * B, RB_CO_FE_BLOCK), RB_CO_FE_BLOCK, mil, (const uint32_t*)nullptr, 1u, B, dcp, dout, c->d_err);
#else
    LAUNCH(c, k_ac17_dec_miller_pair_co, grid_for(2 * 3 * B, RB_CO_BLOCK), RB_CO_BLOCK, ph, sk_offs ? 1 : 0, pg, dc0, lines, unit_lines ? 1 : 0, B, mil, c->d_err);
    launch_final_exp(c, mil, nullptr, 3u, B, dcp, dout, layout);
#endif
#else
    LAUNCH(c, k_ac17_dec_miller_pair, grid_for(3 * B, RB_ML_BLOCK), RB_ML_BLOCK, ph, sk_offs ? 1 : 0, pg, dc0, lines, B, mil, c->d_err);
    LAUNCH(c, k_final_exp, grid_for(B, RB_FE_BLOCK), RB_FE_BLOCK, mil, (const uint32_t*)nullptr, 3u, B, dcp, dout, c->d_err);
#endif
  }
  return finish(c, st);
}

int rb_ac17_cp_decrypt_batch(rb_ctx* c, const uint8_t* k_0, const uint8_t* k, uint32_t n_k, const uint8_t* k_p, const uint8_t* c_0,
                             const uint8_t* cc, uint32_t n1, const uint8_t* c_p, size_t B, const uint32_t* ct_idx,
                             const uint32_t* ct_offs, size_t n_ct_idx, const uint32_t* sk_idx, const uint32_t* sk_offs,
                             size_t n_sk_idx, uint8_t* msg_out) {
  if (!c || !k_0 || !k || !k_p || !c_0 || !cc || !c_p || !msg_out || (!ct_idx && n_ct_idx) || (!sk_idx && n_sk_idx)) return RB_EINVAL;
  if (B == 0) return RB_OK;
  // index range checks for host-resident lists (device-resident lists are the caller's contract)
  if (ct_idx && !is_device_ptr(ct_idx)) for (size_t i = 0; i < n_ct_idx; ++i) if (ct_idx[i] >= n1) return RB_EINVAL;
  if (sk_idx && !is_device_ptr(sk_idx)) for (size_t i = 0; i < n_sk_idx; ++i) if (sk_idx[i] >= n_k) return RB_EINVAL;
  if (!offs_ok(ct_offs, B, n_ct_idx) || !offs_ok(sk_offs, B, n_sk_idx)) return RB_EINVAL;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dk0 = stage_in(c, k_0, 384, st);
  if (st == RB_OK) check_g2(c, dk0, 3);
  const uint8_t* dk = stage_in(c, k, 192 * (size_t)n_k, st);
  const uint8_t* dkp = stage_in(c, k_p, 192, st);
  if (st != RB_OK) return finish(c, st);
  return ac17_decrypt_common(c, dk0, dk, n_k, dkp, nullptr, false, c_0, cc, n1, c_p, B, ct_idx, ct_offs, n_ct_idx, sk_idx, sk_offs, n_sk_idx, msg_out);
}

void rb_ac17_sk_free(rb_ac17_sk* s) {
  if (!s) return;
  Guard g(s->ctx);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->d_k0); cudaFree(s->d_k); cudaFree(s->d_kp); cudaFree(s->lines); cudaFree(s->lines_unit);
  delete s;
}
int rb_ac17_sk_load(rb_ctx* c, const uint8_t* k_0, const uint8_t* k, uint32_t n_k, const uint8_t* k_p, rb_ac17_sk** out) {
  if (!c || !k_0 || !k || !k_p || !out || n_k == 0) return RB_EINVAL;
  *out = nullptr;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  rb_ac17_sk* s = new (std::nothrow) rb_ac17_sk();
  if (!s) return RB_ENOMEM;
  s->ctx = c; s->n_k = n_k; s->d_k0 = s->d_k = s->d_kp = nullptr; s->lines = nullptr; s->lines_unit = nullptr;
  int st = RB_OK;
  if (cudaMalloc(&s->d_k0, 384) != cudaSuccess || cudaMalloc(&s->d_k, 192 * (size_t)n_k) != cudaSuccess || cudaMalloc(&s->d_kp, 192) != cudaSuccess ||
      cudaMalloc(&s->lines, sizeof(MillerLine) * 3 * MILLER_LINES) != cudaSuccess ||
      cudaMalloc(&s->lines_unit, sizeof(MillerLine) * 3 * MILLER_LINES) != cudaSuccess) st = RB_ENOMEM;
  int* d_deg = (st == RB_OK) ? (int*)arena_alloc(c, sizeof(int)) : nullptr;
  if (st == RB_OK && (!d_deg || cudaMemsetAsync(d_deg, 0, sizeof(int), c->stream) != cudaSuccess)) st = RB_ENOMEM;
  auto up = [&](uint8_t* dst, const uint8_t* src, size_t n) {
    cudaMemcpyKind kind = is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (st == RB_OK && cudaMemcpyAsync(dst, src, n, kind, c->stream) != cudaSuccess) st = RB_ECUDA;
  };
  if (st == RB_OK) { up(s->d_k0, k_0, 384); up(s->d_k, k, 192 * (size_t)n_k); up(s->d_kp, k_p, 192); }
  if (st == RB_OK) {
    check_g2(c, s->d_k0, 3);
    LAUNCH(c, k_miller_lines, 1, 32, s->d_k0, 3, s->lines, c->d_err);
    // the tables the decrypt kernels walk: every line divided by its l0 (the cheaper line products of miller_pair / wide.cuh)
    LAUNCH(c, k_lines_normalize, 1, 32, s->lines, s->lines_unit, 3, d_deg);
  }
  c->host_io = true; c->must_sync = true;
  st = finish(c, st);
  if (st != RB_OK) { rb_ac17_sk_free(s); return st; }
  int deg = 1;
  if (cudaMemcpy(&deg, d_deg, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) deg = 1;
  if (deg) { cudaFree(s->lines_unit); s->lines_unit = nullptr; }        // some l0 is zero: the general tables serve this key
  *out = s;
  return RB_OK;
}
int rb_ac17_cp_decrypt_sk_batch(rb_ctx* c, const rb_ac17_sk* sk, const uint8_t* c_0, const uint8_t* cc, uint32_t n1, const uint8_t* c_p,
                                size_t B, const uint32_t* ct_idx, const uint32_t* ct_offs, size_t n_ct_idx, const uint32_t* sk_idx,
                                const uint32_t* sk_offs, size_t n_sk_idx, uint8_t* msg_out) {
  if (!c || !sk || !c_0 || !cc || !c_p || !msg_out || (!ct_idx && n_ct_idx) || (!sk_idx && n_sk_idx)) return RB_EINVAL;
  if (B == 0) return RB_OK;
  if (ct_idx && !is_device_ptr(ct_idx)) for (size_t i = 0; i < n_ct_idx; ++i) if (ct_idx[i] >= n1) return RB_EINVAL;
  if (sk_idx && !is_device_ptr(sk_idx)) for (size_t i = 0; i < n_sk_idx; ++i) if (sk_idx[i] >= sk->n_k) return RB_EINVAL;
  if (!offs_ok(ct_offs, B, n_ct_idx) || !offs_ok(sk_offs, B, n_sk_idx)) return RB_EINVAL;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  return ac17_decrypt_common(c, sk->d_k0, sk->d_k, sk->n_k, sk->d_kp, sk->lines_unit ? sk->lines_unit : sk->lines, sk->lines_unit != nullptr, c_0, cc, n1, c_p,
                             B, ct_idx, ct_offs, n_ct_idx, sk_idx, sk_offs, n_sk_idx, msg_out);
}

int rb_ac17_setup(rb_ctx* c, const uint8_t* rnd, uint8_t* pk, uint8_t* msk) {
  if (!c || !rnd || !pk || !msk) return RB_EINVAL;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* drnd = stage_in(c, rnd, 9 * 32, st);
  uint8_t* dpk = stage_out(c, pk, RB_AC17_PK_BYTES, st);
  uint8_t* dmsk = stage_out(c, msk, RB_AC17_MSK_BYTES, st);
  if (st == RB_OK) LAUNCH(c, k_ac17_setup, 1, 32, drnd, dpk, dmsk, c->d_err);
  return finish(c, st);
}

void rb_ac17_msk_free(rb_ac17_msk* m) {
  if (!m) return;
  rb_table_destroy(m->g);
  rb_table_destroy(m->h);
  { Guard g(m->ctx); cudaStreamSynchronize(m->ctx->stream); cudaFree(m->d_msk); cudaFree(m->consts); }
  delete m;
}
int rb_ac17_msk_load(rb_ctx* c, const uint8_t* mskb, rb_ac17_msk** out) {
  if (!c || !mskb || !out) return RB_EINVAL;
  *out = nullptr;
  uint8_t host[RB_AC17_MSK_BYTES];
  if (is_device_ptr(mskb)) { Guard g(c); CK(cudaMemcpy(host, mskb, sizeof host, cudaMemcpyDeviceToHost)); }
  else memcpy(host, mskb, sizeof host);
  rb_ac17_msk* m = new (std::nothrow) rb_ac17_msk();
  if (!m) return RB_ENOMEM;
  m->ctx = c; m->g = nullptr; m->h = nullptr; m->d_msk = nullptr; m->consts = nullptr;
  int st = rb_g1_table_create(c, host, 16, &m->g);
  if (st == RB_OK) st = rb_g2_table_create(c, host + 64, 8, &m->h);
  if (st == RB_OK) {
    Guard g(c);
    begin_call(c);
    if (cudaMalloc(&m->d_msk, sizeof host) != cudaSuccess || cudaMalloc(&m->consts, sizeof(Ac17MskConsts)) != cudaSuccess) st = RB_ENOMEM;
    else if (cudaMemcpyAsync(m->d_msk, host, sizeof host, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) st = RB_ECUDA;
    else {
      LAUNCH(c, k_ac17_msk_consts, 1, 32, m->d_msk, m->consts, c->d_err);
      c->host_io = true; c->must_sync = true;
      st = finish(c, st);
    }
  }
  if (st != RB_OK) { rb_ac17_msk_free(m); return st; }
  *out = m;
  return RB_OK;
}

int rb_ac17_cp_keygen_batch(rb_ctx* c, const rb_ac17_msk* msk, uint32_t n, const uint8_t* h_attr, const uint8_t* h_01, const uint8_t* rnd,
                            size_t B, uint8_t* k_0, uint8_t* k, uint8_t* k_p) {
  if (!c || !msk || !h_attr || !h_01 || !rnd || !k_0 || !k || !k_p) return RB_EINVAL;
  if (n == 0) return RB_EINVAL;                                        // "empty attributes!" ac17/mod.rs:197
  if (B == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dha = stage_in(c, h_attr, 192 * (size_t)n, st);
  const uint8_t* dh01 = stage_in(c, h_01, 192, st);
  const uint8_t* drnd = stage_in(c, rnd, 32 * (size_t)(n + 3) * B, st);
  uint8_t* dk0 = stage_out(c, k_0, 384 * B, st);
  uint8_t* dk = stage_out(c, k, 192 * (size_t)n * B, st);
  uint8_t* dkp = stage_out(c, k_p, 192 * B, st);
  const size_t rows = (size_t)(n + 1) * B;
  uint8_t* sc = (uint8_t*)arena_alloc(c, 96 * rows);
  uint8_t* sc_k0 = (uint8_t*)arena_alloc(c, 96 * B);
  uint8_t* pts = (uint8_t*)arena_alloc(c, 192 * rows);
  uint32_t* zero_idx = (uint32_t*)arena_alloc(c, 4);
  if (!sc || !sc_k0 || !pts || !zero_idx) st = RB_ENOMEM;
  if (st == RB_OK) {
    cudaMemsetAsync(zero_idx, 0, 4, c->stream);
    LAUNCH(c, k_ac17_keygen_scalars, grid_for(rows, 128), 128, msk->consts, n, dha, dh01, drnd, B, sc, sc_k0, c->d_err);
    size_t outs = rows * 3;
    const int opt = g1_outputs_per_thread(c, k_g1_mul_fixed<G1_M>, 0, outs);
    size_t threads = (outs + opt - 1) / opt;
    LAUNCH(c, k_g1_mul_fixed<G1_M>, grid_for(threads, 128), 128, (const G1Affine*)msk->g->d, msk->g->W, msk->g->nwin, sc, outs, pts, c->d_err, msk->g->stride, opt);
    LAUNCH(c, k_g2_mul_fixed, grid_for(3 * B, 128), 128, (const G2Affine*)msk->h->d, TabSel{0, nullptr}, msk->h->W, msk->h->nwin, sc_k0, 3 * B, dk0, c->d_err);
    if (cudaMemcpy2DAsync(dk, 192 * (size_t)n, pts, 192 * (size_t)(n + 1), 192 * (size_t)n, B, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) st = RB_ECUDA;
    // k_p[t] = g_k[t] + g*sc[key][n][t]   (ac17/mod.rs:247-260)
    GatherArgs ga{pts + 192 * (size_t)n, zero_idx, nullptr, 1, 1, 3, (size_t)(n + 1) * 3, msk->d_msk + 192, 0};
    LAUNCH(c, k_g1_gather_sum, grid_for(3 * B, 128), 128, ga, B, (G1Affine*)nullptr, dkp, c->d_err);
  }
  return finish(c, st);
}

// ---- Fr / group element-wise -------------------------------------------------------------------
static int fr_op_ex(rb_ctx* c, int op, const uint8_t* a, const uint8_t* b, OpIdx bi, size_t b_count, size_t n, uint8_t* out) {
  if (!c || !a || !out || op < 0 || op > 4 || (op <= 2 && !b)) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 32 * n, st);
  const uint8_t* db = (op <= 2) ? stage_in(c, b, 32 * b_count, st) : nullptr;
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) LAUNCH(c, k_fr_op, grid_for(n, 128), 128, op, da, db, bi, n, dout, c->d_err);
  return finish(c, st);
}
int rb_fr_op_batch(rb_ctx* c, int op, const uint8_t* a, const uint8_t* b, int b_is_scalar, size_t n, uint8_t* out) {
  return fr_op_ex(c, op, a, b, b_is_scalar ? OpIdx{1, 1} : EACH, b_is_scalar ? 1 : n, n, out);
}
static int g1_add_ex(rb_ctx* c, const uint8_t* a, const uint8_t* b, OpIdx bi, size_t b_count, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 64 * n, st);
  const uint8_t* db = stage_in(c, b, 64 * b_count, st);
  uint8_t* dout = stage_out(c, out, 64 * n, st);
  if (st == RB_OK) LAUNCH(c, k_g1_add, grid_for(n, 128), 128, da, db, bi, n, dout, c->d_err);
  return finish(c, st);
}
int rb_g1_add_batch(rb_ctx* c, const uint8_t* a, const uint8_t* b, int b_is_point, size_t n, uint8_t* out) {
  return g1_add_ex(c, a, b, b_is_point ? OpIdx{1, 1} : EACH, b_is_point ? 1 : n, n, out);
}
static int g2_add_ex(rb_ctx* c, const uint8_t* a, const uint8_t* b, OpIdx bi, size_t b_count, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 128 * n, st);
  const uint8_t* db = stage_in(c, b, 128 * b_count, st);
  uint8_t* dout = stage_out(c, out, 128 * n, st);
  if (st == RB_OK) { check_g2(c, da, n); check_g2(c, db, b_count); LAUNCH(c, k_g2_add, grid_for(n, 128), 128, da, db, bi, n, dout, c->d_err); }
  return finish(c, st);
}
int rb_g2_add_batch(rb_ctx* c, const uint8_t* a, const uint8_t* b, int b_is_point, size_t n, uint8_t* out) {
  return g2_add_ex(c, a, b, b_is_point ? OpIdx{1, 1} : EACH, b_is_point ? 1 : n, n, out);
}

// SHA3-256 -> Fr of n byte strings (hash/mod.rs:23-31); offs [n+1] byte offsets into data
static int sha3_fr_batch_impl(rb_ctx* c, const uint8_t* data, const uint32_t* offs, size_t n, uint8_t* out, int64_t data_len) {
  if (!c || !offs || !out || (!data && n)) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  uint32_t total = 0;
  if (is_device_ptr(offs)) {
    if (data_len >= 0) total = (uint32_t)data_len;        // caller states offs[n]: no device->host read, the call stays stream-ordered
    else { CK(cudaMemcpyAsync(&total, offs + n, 4, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
  } else {
    uint32_t bad = 0;                                       // branch-free: the compiler vectorises it (millions of labels per call)
    for (size_t i = 0; i < n; ++i) bad |= (uint32_t)(offs[i + 1] < offs[i]);
    if (bad) return RB_EINVAL;
    total = offs[n];
    if (data_len >= 0 && (uint64_t)data_len != total) return RB_EINVAL;
  }
  const uint8_t* dd = stage_in(c, data, total ? total : 1, st);
  const uint32_t* doffs = stage_in(c, offs, 4 * (n + 1), st);
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) LAUNCH(c, k_sha3_fr, grid_for(n, 128), 128, dd, doffs, n, dout);
  return finish(c, st);
}
int rb_sha3_fr_batch(rb_ctx* c, const uint8_t* data, const uint32_t* offs, size_t n, uint8_t* out) {
  return sha3_fr_batch_impl(c, data, offs, n, out, -1);
}
int rb_sha3_fr_batch_len(rb_ctx* c, const uint8_t* data, size_t data_len, const uint32_t* offs, size_t n, uint8_t* out) {
  if (data_len > 0xffffffffull) return RB_EINVAL;
  return sha3_fr_batch_impl(c, data, offs, n, out, (int64_t)data_len);
}

// ---- secret sharing ------------------------------------------------------------------------------
void rb_share_plan_free(rb_share_plan* p) {
  if (!p) return;
  Guard g(p->ctx);
  cudaStreamSynchronize(p->ctx->stream);
  cudaFree(p->terms); cudaFree(p->leaf_offs); cudaFree(p->consts);
  delete p;
}
int rb_share_plan_dims(const rb_share_plan* p, uint32_t* n_leaves, uint32_t* n_coefs) {
  if (!p) return RB_EINVAL;
  if (n_leaves) *n_leaves = p->n_leaves;
  if (n_coefs) *n_coefs = p->n_coefs;
  return RB_OK;
}
int rb_share_plan_create_raw(rb_ctx* c, const uint32_t* terms, uint32_t n_terms, const uint32_t* leaf_offs, uint32_t n_leaves,
                             uint32_t n_coefs, rb_share_plan** out) {
  if (!c || !leaf_offs || !out || n_leaves == 0 || (!terms && n_terms)) return RB_EINVAL;
  *out = nullptr;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  rb_share_plan* p = new (std::nothrow) rb_share_plan();
  if (!p) return RB_ENOMEM;
  p->ctx = c; p->n_terms = n_terms; p->n_leaves = n_leaves; p->n_coefs = n_coefs; p->terms = nullptr; p->leaf_offs = nullptr; p->consts = nullptr;
  size_t tb = sizeof(ShareTerm) * (size_t)(n_terms ? n_terms : 1);
  int st = RB_OK;
  if (cudaMalloc(&p->terms, tb) != cudaSuccess || cudaMalloc(&p->leaf_offs, 4 * (size_t)(n_leaves + 1)) != cudaSuccess ||
      cudaMalloc(&p->consts, sizeof(Fr) * (size_t)(n_terms ? n_terms : 1)) != cudaSuccess) st = RB_ENOMEM;
  if (st == RB_OK && n_terms && cudaMemcpyAsync(p->terms, terms, sizeof(ShareTerm) * (size_t)n_terms, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) st = RB_ECUDA;
  if (st == RB_OK && cudaMemcpyAsync(p->leaf_offs, leaf_offs, 4 * (size_t)(n_leaves + 1), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) st = RB_ECUDA;
  if (st == RB_OK && n_terms) LAUNCH(c, k_share_consts, grid_for(n_terms, 128), 128, p->terms, n_terms, p->consts);
  c->host_io = true; c->must_sync = true;
  st = finish(c, st);
  if (st != RB_OK) { rb_share_plan_free(p); return st; }
  *out = p;
  return RB_OK;
}
int rb_shares_batch(rb_ctx* c, const rb_share_plan* p, const uint8_t* secret, const uint8_t* coeffs, size_t B, uint8_t* shares) {
  if (!c || !p || !secret || !shares || (!coeffs && p->n_coefs)) return RB_EINVAL;
  if (B == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* ds = stage_in(c, secret, 32 * B, st);
  const uint8_t* dc = p->n_coefs ? stage_in(c, coeffs, 32 * B * p->n_coefs, st) : nullptr;
  uint8_t* dout = stage_out(c, shares, 32 * B * p->n_leaves, st);
  if (st == RB_OK) LAUNCH(c, k_shares, grid_for(B * p->n_leaves, 128), 128, p->terms, p->consts, p->leaf_offs, p->n_leaves, p->n_coefs, ds, dc, B, dout, c->d_err);
  return finish(c, st);
}
int rb_lagrange_raw(rb_ctx* c, const uint32_t* terms, uint32_t n_terms, const uint32_t* leaf_offs, uint32_t n_leaves, uint8_t* out) {
  if (!c || !leaf_offs || !out || n_leaves == 0 || (!terms && n_terms)) return RB_EINVAL;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const ShareTerm* dt = (const ShareTerm*)stage_in(c, terms, sizeof(ShareTerm) * (size_t)(n_terms ? n_terms : 0), st);
  const uint32_t* dlo = stage_in(c, leaf_offs, 4 * (size_t)(n_leaves + 1), st);
  uint8_t* dout = stage_out(c, out, 32 * (size_t)n_leaves, st);
  if (st == RB_OK) LAUNCH(c, k_lagrange_coeffs, grid_for(n_leaves, 128), 128, dt, dlo, n_leaves, dout);
  return finish(c, st);
}

int rb_ac17_kp_keygen_batch(rb_ctx* c, const rb_ac17_msk* msk, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row,
                            const uint8_t* h_col, const uint8_t* rnd, size_t B, uint8_t* k_0, uint8_t* k) {
  if (!c || !msk || !m || !h_row || !h_col || !rnd || !k_0 || !k) return RB_EINVAL;
  if (n1 == 0 || n2 == 0) return RB_EPOLICY;
  if (B == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const size_t per_key = 2 + (size_t)(n2 - 1) + n1;
  const int8_t* dm = stage_in(c, m, (size_t)n1 * n2, st);
  const uint8_t* dhr = stage_in(c, h_row, (size_t)n1 * 6 * 32, st);
  const uint8_t* dhc = stage_in(c, h_col, (size_t)n2 * 6 * 32, st);
  const uint8_t* drnd = stage_in(c, rnd, 32 * per_key * B, st);
  uint8_t* dk0 = stage_out(c, k_0, 384 * B, st);
  uint8_t* dk = stage_out(c, k, 192 * (size_t)n1 * B, st);
  const size_t rows = (size_t)n1 * B;
  uint8_t* sc = (uint8_t*)arena_alloc(c, 96 * rows);
  uint8_t* sc_k0 = (uint8_t*)arena_alloc(c, 96 * B);
  uint8_t* pts = (uint8_t*)arena_alloc(c, 192 * rows);
  if (!sc || !sc_k0 || !pts) st = RB_ENOMEM;
  if (st == RB_OK) {
    LAUNCH(c, k_ac17_kp_keygen_scalars, grid_for(rows, 128), 128, msk->consts, n1, n2, dm, dhr, dhc, drnd, B, sc, sc_k0, c->d_err);
    size_t outs = rows * 3;
    const int opt = g1_outputs_per_thread(c, k_g1_mul_fixed<G1_M>, 0, outs);
    size_t threads = (outs + opt - 1) / opt;
    LAUNCH(c, k_g1_mul_fixed<G1_M>, grid_for(threads, 128), 128, (const G1Affine*)msk->g->d, msk->g->W, msk->g->nwin, sc, outs, pts, c->d_err, msk->g->stride, opt);
    LAUNCH(c, k_g2_mul_fixed, grid_for(3 * B, 128), 128, (const G2Affine*)msk->h->d, TabSel{0, nullptr}, msk->h->W, msk->h->nwin, sc_k0, 3 * B, dk0, c->d_err);
    LAUNCH(c, k_ac17_kp_finish, grid_for(outs, 128), 128, pts, msk->d_msk + 192, n1, n2, dm, B, dk, c->d_err);
  }
  return finish(c, st);
}

// ---- KEM tail (utils/aes/mod.rs:10-55) ---------------------------------------------------------------------
static int kem_batch(rb_ctx* c, const uint8_t* gt, const uint8_t* nonce, const uint8_t* in, const uint32_t* offs, size_t B, int decrypt, uint8_t* out, int* ok) {
  if (!c || !gt || !in || !offs || !out || (!decrypt && !nonce) || (decrypt && !ok)) return RB_EINVAL;
  if (B == 0) return RB_OK;
  if (is_device_ptr(offs)) return RB_EINVAL;                    // the blob lengths size the staging copies: host offsets only
  for (size_t i = 0; i < B; ++i) if (offs[i + 1] < offs[i]) return RB_EINVAL;
  const size_t total = offs[B];
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dgt = stage_in(c, gt, 384 * B, st);
  const uint8_t* dn = decrypt ? nullptr : stage_in(c, nonce, 12 * B, st);
  const uint8_t* din = stage_in(c, in, total ? total : 1, st);
  const uint32_t* doffs = stage_in(c, offs, 4 * (B + 1), st);
  const size_t out_bytes = decrypt ? total : total + 28 * B;     // decrypt: plaintext b (28 bytes shorter than its blob) sits at the blob's offset
  uint8_t* dout = stage_out(c, out, out_bytes ? out_bytes : 1, st);
  int* dok = decrypt ? stage_out(c, ok, sizeof(int) * B, st) : nullptr;
  if (st == RB_OK) LAUNCH(c, k_kem_aes256gcm, grid_for(B, 128), 128, dgt, dn, din, doffs, B, decrypt, dout, dok);
  return finish(c, st);
}
int rb_kem_encrypt_batch(rb_ctx* c, const uint8_t* gt, const uint8_t* nonce, const uint8_t* data, const uint32_t* offs, size_t B, uint8_t* out) {
  return kem_batch(c, gt, nonce, data, offs, B, 0, out, nullptr);
}
int rb_kem_decrypt_batch(rb_ctx* c, const uint8_t* gt, const uint8_t* nonce_ct, const uint32_t* offs, size_t B, uint8_t* out, int* ok) {
  return kem_batch(c, gt, nullptr, nonce_ct, offs, B, 1, out, ok);
}

// ---- test hooks of the six-lane layer (internal.h; tests/test_gpu_wide.py) ----------------------------
int rb_dbg_wide_dot(rb_ctx* c, const uint8_t* xs, const uint8_t* ys, int K, size_t n, uint8_t* out) {
  if (!c || !xs || !ys || !out || K < 1 || K > 6) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* dx = stage_in(c, xs, 32 * n * K, st);
  const uint8_t* dy = stage_in(c, ys, 32 * n * K, st);
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) LAUNCH(c, k_dbg_wide_dot, grid_for(n, 128), 128, dx, dy, K, n, dout);
  return finish(c, st);
}
int rb_dbg_fq_sqr(rb_ctx* c, const uint8_t* a, const uint8_t* b, int mode, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out || mode < 0 || mode > 2) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 32 * n, st);
  const uint8_t* db = stage_in(c, b, 32 * n, st);
  uint8_t* dout = stage_out(c, out, 32 * n, st);
  if (st == RB_OK) LAUNCH(c, k_dbg_fq_sqr, grid_for(n, 128), 128, da, db, mode, n, dout);
  return finish(c, st);
}
int rb_dbg_w6_op(rb_ctx* c, int op, int arg, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  if (!c || !a || !b || !out) return RB_EINVAL;
  if (n == 0) return RB_OK;
  Guard g(c); if (!g.ok) return RB_ECUDA;
  begin_call(c);
  int st = RB_OK;
  const uint8_t* da = stage_in(c, a, 384 * n, st);
  const uint8_t* db = stage_in(c, b, 384 * n, st);
  uint8_t* dout = stage_out(c, out, 384 * n, st);
  if (st == RB_OK) LAUNCH(c, k_dbg_w6_op, w6_grid(n, RB_W6_BLOCK), RB_W6_BLOCK, op, arg, da, db, n, dout, c->d_err);
  return finish(c, st);
}
int rb_ctx_set_pairing_layout(rb_ctx* c, int mode) {
  if (!c || mode < RB_PAIRING_AUTO || mode > 3) return RB_EINVAL;      // 3 = hybrid (development)
  c->pairing_mode = mode;
  return RB_OK;
}
int rb_ctx_set_async(rb_ctx* c, int enable) {
  if (!c) return RB_EINVAL;
  Guard g(c);
  cudaStreamSynchronize(c->stream);
  c->async_host = enable != 0;
  return RB_OK;
}

#include "scheme_batch.inc"

}  // extern "C"
