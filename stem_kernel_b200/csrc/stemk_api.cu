// stem_kernel_b200/csrc/stemk_api.cu -- implementation of the C ABI in include/stemk.h.
//
// A context owns one CUDA device, one stream, the kernel constants and growable device scratch;
// a set owns the device copy of a compiled record set.  The Gram / cross / diagonal entry points
// (KernelMatrix::calculate / diagonal, common/kernel_matrix.cpp:485-754) are thin drivers over one
// pair-list evaluator.  There is deliberately no CPU path: without a device every entry point
// returns STEMK_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <memory>

#include "kernels.cuh"

using namespace stemk;


namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace

struct stemk_set {
  // host-side record headers and statistics (work model, scheduler, launch shapes); shared between the copies of a
  // set on several devices (stemk_set_clone), immutable after the upload
  std::shared_ptr<CompiledSet> hostp;
  CompiledSet& host;
  DevBuf blob;        // every array of the view in one allocation
  SetView view;       // device pointers
  int device = 0;
  double loop_gap = 0;      // the parameters the derived tables were built under
  uint32_t len_band = 0;
  size_t lay[kBlobArrays] = {};   // offset of every array of the view inside the blob (same order as make_view)
  stemk_set() : hostp(std::make_shared<CompiledSet>()), host(*hostp) {}
  explicit stemk_set(std::shared_ptr<CompiledSet> h) : hostp(std::move(h)), host(*hostp) {}
};

struct stemk_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  stemk_params params;
  KernelTables tables;
  cudaStream_t stream = nullptr;
  double* d_pair_tab = nullptr;
  double* d_subst = nullptr;
  unsigned long long* d_counter = nullptr;
  DevBuf blob_cache;                           // device image of the last freed set, handed to the next stemk_upload that fits
  DevBuf scratch, scratch_big, fold_scratch, carry, tmp_stem, tmp_str, idx_x, idx_y, vals, matrix, order, rowacc;
  DevBuf perm, offs, diag, selfv, diag_idx, diag_vals, diag_idx2, diag_vals2;
  DevBuf deal_x, deal_y, gathered, undealt;     // stemk_gram_multi: this device's share of the pair list; on device 0 the gather
  cudaEvent_t multi_ev = nullptr;
  void* upload_stage = nullptr;                 // pinned host image of the set being uploaded (kept for the next upload)
  size_t upload_stage_bytes = 0;
  cudaStream_t last_stream = nullptr;           // stream of the previous device-side call (scratch, queues and slabs are per context)
  cudaEvent_t order_ev = nullptr;
  void* stage[2] = {nullptr, nullptr};          // pinned host staging (copy_out / copy_in)
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  unsigned long long* d_bucket = nullptr;  // count[16] | start[16] | queue heads[16]
  int use_fast = 1;                        // stemk_set_option(STEMK_OPT_FORCE_GENERAL, 1) routes every pair to the general stem kernel
  FoldResult fold;                         // result of the last stemk_fold_bpp
  uint32_t fold_n = 0;
  int force_unstaged = 0;                  // stemk_set_option(STEMK_OPT_FORCE_UNSTAGED, 1): the general kernel's pairs all run unstaged
  int timing = 0;                          // stemk_set_option(STEMK_OPT_TIMING, 1): host-side breakdown of the calls on stderr
  std::string err;
  // stats
  uint64_t launches = 0;
  double stem_ms = 0, string_ms = 0, fold_ms = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // launch timings are resolved lazily (stemk_stats_get) so that the launch path never blocks the host
  struct Timed { cudaEvent_t a, b; int which; };
  std::vector<Timed> pending, free_events;
};

namespace {

// event pair around one kernel launch; `which` 0 = stem, 1 = string
stemk_ctx::Timed timed_begin(stemk_ctx* c, int which, cudaStream_t st) {
  stemk_ctx::Timed t{nullptr, nullptr, which};
  if (!c->free_events.empty()) { t = c->free_events.back(); c->free_events.pop_back(); t.which = which; }
  else { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
  cudaEventRecord(t.a, st);
  return t;
}
void timed_resolve(stemk_ctx* c) {
  for (auto& t : c->pending) {
    float ms = 0;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess)
      (t.which == 0 ? c->stem_ms : c->string_ms) += ms;
    c->free_events.push_back(t);
  }
  c->pending.clear();
}
void timed_end(stemk_ctx* c, stemk_ctx::Timed t, cudaStream_t st) {
  cudaEventRecord(t.b, st);
  c->pending.push_back(t);
  if (c->pending.size() >= 256) timed_resolve(c);
}

// The context's scratch buffers, work queues and DP slabs belong to ONE call at a time.  When a device-side call
// arrives on another stream than the previous one, the new stream first waits (on the device) for everything the
// previous stream was given, so two calls never share those buffers in flight.
cudaError_t order_after_previous(stemk_ctx* c, cudaStream_t st) {
  if (c->last_stream && c->last_stream != st) {
    if (!c->order_ev) { cudaError_t e = cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    if (cudaEventRecord(c->order_ev, c->last_stream) == cudaSuccess) {
      const cudaError_t e = cudaStreamWaitEvent(st, c->order_ev, 0);
      if (e != cudaSuccess) return e;
    } else {
      cudaGetLastError();                       // the caller destroyed that stream: its work was synchronised by then
    }
  }
  c->last_stream = st;
  return cudaSuccess;
}

int fail(stemk_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(stemk_ctx* c, cudaError_t e, const char* where) {
  return fail(c, STEMK_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
bool set_usable(const stemk_ctx* ctx, const stemk_set* s) {
  return s->device == ctx->device && s->loop_gap == ctx->params.loop_gap && s->len_band == ctx->params.len_band;
}
const char* kSetMismatch = "set was uploaded through a context with another device, loop gap or length band";
int no_device(stemk_ctx* c) {
  return fail(c, STEMK_ERR_CUDA, "host-only context: kernel values are computed on a CUDA device only (no CPU path)");
}
#define CU(call)                                                     \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);         \
  } while (0)

template <class T>
size_t place(size_t& off, const std::vector<T>& v) {
  off = (off + 255) & ~size_t(255);
  size_t at = off;
  off += v.size() * sizeof(T);
  return at;
}

// number of row slots (warps with private Q/G1 rows) and shared memory for a launch of x set against y set
void stem_config(const stemk_ctx* ctx, uint32_t nx, uint32_t ny, uint32_t ey, uint32_t lv, uint32_t* nslots, size_t* smem) {
  const size_t budget = std::min<size_t>(ctx->smem_optin, (size_t)226 * 1024);
  uint32_t best = 0;
  for (uint32_t w = (uint32_t)stem_warps_per_cta(); w >= 1; --w) if (stem_smem_bytes(w, nx, ny, ey, lv) <= budget) { best = w; break; }
  *nslots = best;
  *smem = best ? stem_smem_bytes(best, nx, ny, ey, lv) : 0;
}

// Records the staged general kernel takes when the largest records of a set do not fit its shared-memory carve-up with
// at least kGenMinSlots row slots: 56 B per node + 16 B per inner edge + 4 B per x row + two rows of 8 B per node and
// slot leave 8 slots at these sizes.  Larger records run on the unstaged kernel.
constexpr uint32_t kGenMinSlots = 4, kGenCapNx = 2048, kGenCapNy = 640, kGenCapEy = 5120;

}  // namespace

extern "C" {

const char* stemk_version(void) { return "stemk-b200 0.1 (sm_100a)"; }

int stemk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* stemk_last_error(const stemk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int stemk_create(stemk_ctx** out, const stemk_params* params, int device) {
  if (!out || !params) return fail(nullptr, STEMK_ERR_ARG, "null argument");
  *out = nullptr;
  if (params->kind < STEMK_SI_STEM || params->kind > STEMK_STR_NAIVE) return fail(nullptr, STEMK_ERR_ARG, "unknown kernel kind");
  if (device == STEMK_DEVICE_NONE) {
    // host-only context: record compilation and the work model, nothing that computes a kernel value
    stemk_ctx* c = new stemk_ctx;
    c->device = STEMK_DEVICE_NONE;
    c->params = *params;
    make_tables(*params, &c->tables);
    *out = c;
    return STEMK_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, STEMK_ERR_CUDA, "no CUDA device: this library has no CPU path");
  }
  if (device < 0 || device >= n) return fail(nullptr, STEMK_ERR_ARG, "device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  stemk_ctx* c = new stemk_ctx;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->params = *params;
  make_tables(*params, &c->tables);
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&c->d_pair_tab, sizeof(double) * 256) == cudaSuccess &&
            cudaMalloc((void**)&c->d_subst, sizeof(double) * 16) == cudaSuccess &&
            cudaMalloc((void**)&c->d_counter, sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc((void**)&c->d_bucket, sizeof(unsigned long long) * 48) == cudaSuccess &&
            cudaMemcpy(c->d_pair_tab, c->tables.pair_tab, sizeof(double) * 256, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(c->d_subst, c->tables.subst, sizeof(double) * 16, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
  if (!ok) {
    e = cudaGetLastError();
    stemk_destroy(c);
    return cuda_fail(nullptr, e, "context allocation");
  }
  *out = c;
  return STEMK_OK;
}

int stemk_set_option(stemk_ctx* ctx, int option, int value) {
  if (!ctx) return fail(ctx, STEMK_ERR_ARG, "null argument");
  switch (option) {
    case STEMK_OPT_FORCE_GENERAL: ctx->use_fast = value ? 0 : 1; return STEMK_OK;
    case STEMK_OPT_FORCE_UNSTAGED: ctx->force_unstaged = value ? 1 : 0; return STEMK_OK;
    case STEMK_OPT_TIMING: ctx->timing = value; return STEMK_OK;
    default: return fail(ctx, STEMK_ERR_ARG, "unknown option");
  }
}

void stemk_destroy(stemk_ctx* c) {
  if (!c) return;
  if (c->device == STEMK_DEVICE_NONE) { delete c; return; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->blob_cache, &c->scratch, &c->scratch_big, &c->fold_scratch, &c->carry, &c->tmp_stem, &c->tmp_str, &c->idx_x, &c->idx_y, &c->vals, &c->matrix, &c->order, &c->rowacc,
                    &c->perm, &c->offs, &c->diag, &c->selfv, &c->diag_idx, &c->diag_vals, &c->diag_idx2, &c->diag_vals2,
                    &c->deal_x, &c->deal_y, &c->gathered, &c->undealt}) b->release();
  if (c->multi_ev) cudaEventDestroy(c->multi_ev);
  if (c->order_ev) cudaEventDestroy(c->order_ev);
  if (c->upload_stage) cudaFreeHost(c->upload_stage);
  for (int k = 0; k < 2; ++k) { if (c->stage[k]) cudaFreeHost(c->stage[k]); if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]); }
  if (c->d_pair_tab) cudaFree(c->d_pair_tab);
  if (c->d_subst) cudaFree(c->d_subst);
  if (c->d_counter) cudaFree(c->d_counter);
  if (c->d_bucket) cudaFree(c->d_bucket);
  timed_resolve(c);
  for (auto& t : c->free_events) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

namespace {

// device pointers of a set from its blob and the array offsets recorded at upload
void make_view(stemk_set* s) {
  char* b = static_cast<char*>(s->blob.p);
  const size_t* o = s->lay;
  SetView& v = s->view;
  v.n_recs = (uint32_t)s->host.rec.size();
  v.rec = (const RecDev*)(b + o[0]); v.a = (const double*)(b + o[1]); v.el = (const double*)(b + o[2]);
  v.ql = (const double*)(b + o[3]); v.paths = (const double*)(b + o[4]); v.gapt = (const double*)(b + o[5]);
  v.bfreq = (const double*)(b + o[6]); v.len = (const uint32_t*)(b + o[7]); v.bcode = (const uint8_t*)(b + o[8]);
  v.coff = (const uint32_t*)(b + o[9]); v.cidx = (const uint32_t*)(b + o[10]); v.ce = (const double*)(b + o[11]);
  v.lev_off = (const uint32_t*)(b + o[12]); v.boff = (const uint32_t*)(b + o[13]); v.bab = (const uint8_t*)(b + o[14]);
  v.bfq = (const double*)(b + o[15]); v.ccode = (const uint8_t*)(b + o[16]); v.cw = (const double*)(b + o[17]);
  v.prof = (const float*)(b + o[18]); v.text = (const uint8_t*)(b + o[19]);
  v.up = (const double*)(b + o[20]); v.dn = (const double*)(b + o[21]); v.s2 = (const double*)(b + o[22]);
  v.nodei = (const NodeI*)(b + o[23]); v.c16 = (const uint16_t*)(b + o[24]); v.blk = (const uint32_t*)(b + o[25]);
  v.xnode = (const XNode*)(b + o[26]); v.lperm = (const uint32_t*)(b + o[27]);
}

}  // namespace

int stemk_upload(stemk_ctx* ctx, const stemk_seqset_desc* desc, stemk_set** out) {
  if (!ctx || !desc || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  *out = nullptr;
  const bool on_device = ctx->device != STEMK_DEVICE_NONE;
  if (on_device) CU(cudaSetDevice(ctx->device));
  stemk_set* s = new stemk_set;
  s->device = ctx->device;
  s->loop_gap = ctx->params.loop_gap;
  s->len_band = ctx->params.len_band;
  const int n_threads = (int)std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
  // The records are compiled in parallel and written straight into ONE pinned host image of the device blob (kept
  // by the context for the next upload), which then goes to the device with a single copy.
  cudaError_t stage_err = cudaSuccess;
  std::function<char*(size_t)> sink;
  if (on_device) sink = [&](size_t bytes) -> char* {
    if (bytes > ctx->upload_stage_bytes) {
      if (ctx->upload_stage) cudaFreeHost(ctx->upload_stage);
      ctx->upload_stage = nullptr; ctx->upload_stage_bytes = 0;
      stage_err = cudaHostAlloc(&ctx->upload_stage, bytes + bytes / 8, cudaHostAllocDefault);
      if (stage_err != cudaSuccess) { cudaGetLastError(); ctx->upload_stage = nullptr; return nullptr; }
      ctx->upload_stage_bytes = bytes + bytes / 8;
    }
    return static_cast<char*>(ctx->upload_stage);
  };
  std::string err;
  const auto t0 = std::chrono::steady_clock::now();
  try { err = compile_set(*desc, ctx->params.loop_gap, ctx->params.len_band, n_threads, ctx->timing != 0, &s->host, sink); }
  catch (const std::bad_alloc&) { delete s; return fail(ctx, STEMK_ERR_NOMEM, "out of host memory while compiling the record set"); }
  catch (const std::exception& ex) { delete s; return fail(ctx, STEMK_ERR_ARG, std::string("record set: ") + ex.what()); }
  if (stage_err != cudaSuccess) { delete s; return cuda_fail(ctx, stage_err, "pinned staging for the set upload"); }
  if (!err.empty()) { delete s; return fail(ctx, STEMK_ERR_ARG, err); }
  if (!on_device) { *out = s; return STEMK_OK; }
  const CompiledSet& h = s->host;
  for (int k = 0; k < kBlobArrays; ++k) s->lay[k] = h.blob_lay[k];
  const auto t1 = std::chrono::steady_clock::now();
  // A caller that uploads a set per call (upload, gram, free, ...) would otherwise allocate and free ~100 KB per record
  // of device memory every time; the image of the last freed set is kept and reused when it fits without waste.
  const size_t need = std::max<size_t>(h.blob_bytes, 256);
  if (ctx->blob_cache.p && ctx->blob_cache.bytes >= need && ctx->blob_cache.bytes <= need + need / 4) {
    s->blob = ctx->blob_cache;
    ctx->blob_cache = DevBuf();
  }
  cudaError_t e = s->blob.reserve(need);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->blob.p, ctx->upload_stage, h.blob_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { s->blob.release(); delete s; return cuda_fail(ctx, e, "set upload"); }
  if (ctx->timing)
    std::fprintf(stderr, "stemk_upload: compile %.1f ms, device allocation + one copy of %.1f MB %.1f ms\n",
                 std::chrono::duration<double, std::milli>(t1 - t0).count(), h.blob_bytes / 1e6,
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  make_view(s);
  *out = s;
  return STEMK_OK;
}

void stemk_set_free(stemk_ctx* ctx, stemk_set* s) {
  if (!s) return;
  if (ctx && ctx->device != STEMK_DEVICE_NONE) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (ctx && ctx->device != STEMK_DEVICE_NONE && s->device == ctx->device && s->blob.p && s->blob.bytes > ctx->blob_cache.bytes) {
    ctx->blob_cache.release();          // keep the larger of the two images for the next upload
    ctx->blob_cache = s->blob;
    s->blob = DevBuf();
  }
  s->blob.release();
  delete s;
}

// ---- one set on several devices -----------------------------------------------------------------------------
int stemk_set_clone(stemk_ctx* ctx, const stemk_set* src, stemk_set** out) {
  if (!ctx || !src || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  *out = nullptr;
  if (ctx->device == STEMK_DEVICE_NONE || src->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (src->loop_gap != ctx->params.loop_gap || src->len_band != ctx->params.len_band)
    return fail(ctx, STEMK_ERR_ARG, "stemk_set_clone: the destination context has another loop gap or length band");
  CU(cudaSetDevice(ctx->device));
  stemk_set* s = new stemk_set(src->hostp);   // record headers and statistics are shared, not copied
  s->device = ctx->device; s->loop_gap = src->loop_gap; s->len_band = src->len_band;
  std::memcpy(s->lay, src->lay, sizeof(s->lay));
  cudaError_t e = s->blob.reserve(std::max<size_t>(src->blob.bytes, 256));
  // device to device: over NVLink when the two devices are peers, staged by the driver otherwise
  if (e == cudaSuccess) e = cudaMemcpyPeerAsync(s->blob.p, ctx->device, src->blob.p, src->device, src->blob.bytes, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { s->blob.release(); delete s; return cuda_fail(ctx, e, "stemk_set_clone"); }
  make_view(s);
  *out = s;
  return STEMK_OK;
}

int stemk_upload_multi(stemk_ctx* const* ctxs, int n_ctx, const stemk_seqset_desc* desc, stemk_set** sets) {
  if (!ctxs || n_ctx < 1 || !desc || !sets) return fail(n_ctx > 0 && ctxs ? ctxs[0] : nullptr, STEMK_ERR_ARG, "null argument");
  for (int d = 0; d < n_ctx; ++d) sets[d] = nullptr;
  int rc = stemk_upload(ctxs[0], desc, &sets[0]);   // the records are compiled once ...
  for (int d = 1; rc == STEMK_OK && d < n_ctx; ++d) {
    rc = stemk_set_clone(ctxs[d], sets[0], &sets[d]);   // ... and reach the other devices device-to-device
    if (rc != STEMK_OK) fail(ctxs[0], rc, std::string("device ") + std::to_string(d) + ": " + stemk_last_error(ctxs[d]));
  }
  if (rc != STEMK_OK) for (int d = 0; d < n_ctx; ++d) { stemk_set_free(ctxs[d], sets[d]); sets[d] = nullptr; }
  return rc;
}

// A set as one byte string -- [header | host-side record headers and statistics | device blob] -- written into a
// DEVICE buffer, so that a multi-process driver can broadcast it (NCCL) instead of compiling the records once per
// rank.  Layout private to this library version.
}  // extern "C" (templates below)
namespace {
constexpr uint64_t kExportMagic = 0x53544d4b53455432ull;   // "STMKSET2"
struct ExportHeader {
  uint64_t magic, meta_bytes, blob_bytes;
  double loop_gap;
  uint32_t len_band, n_recs;
  uint64_t lay[kBlobArrays];
  uint32_t max_E4, max_fastN, n_fast, max_band_cnt, n_weighted, n_simple_cols, has_dag, max_N, max_L, max_E, max_nlev, pad_;
  uint64_t cnt[12];   // element counts of the meta arrays, in the order of meta_arrays()
};
template <class F>
void meta_arrays(CompiledSet& h, F&& f) {   // every host array that outlives the upload
  f(0, h.rec); f(1, h.n_nodes_all); f(2, h.n_edges_all); f(3, h.max_level_rows); f(4, h.band_cnt); f(5, h.deg_all);
  f(6, h.len); f(7, h.boff); f(8, h.text); f(9, h.cost_off); f(10, h.cost_pd); f(11, h.cost_pb);
}
size_t round256(size_t v) { return (v + 255) & ~size_t(255); }
}  // namespace
extern "C" {

uint64_t stemk_set_export_bytes(const stemk_set* s) {
  if (!s || s->device == STEMK_DEVICE_NONE) return 0;
  size_t meta = 0;
  meta_arrays(const_cast<CompiledSet&>(s->host), [&](int, auto& v) { meta += round256(v.size() * sizeof(v[0])); });
  return round256(sizeof(ExportHeader)) + meta + round256(s->blob.bytes);
}

int stemk_set_export(stemk_ctx* ctx, const stemk_set* s, void* d_dst, void* stream_) {
  if (!ctx || !s || !d_dst) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (!set_usable(ctx, s)) return fail(ctx, STEMK_ERR_ARG, kSetMismatch);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
  const CompiledSet& h = s->host;
  ExportHeader hd;
  std::memset(&hd, 0, sizeof(hd));
  hd.magic = kExportMagic; hd.blob_bytes = s->blob.bytes; hd.loop_gap = s->loop_gap; hd.len_band = s->len_band;
  hd.n_recs = (uint32_t)h.rec.size();
  for (int k = 0; k < kBlobArrays; ++k) hd.lay[k] = s->lay[k];
  hd.max_E4 = h.max_E4; hd.max_fastN = h.max_fastN; hd.n_fast = h.n_fast; hd.max_band_cnt = h.max_band_cnt;
  hd.n_weighted = h.n_weighted; hd.n_simple_cols = h.n_simple_cols; hd.has_dag = h.has_dag ? 1u : 0u;
  hd.max_N = h.max_N; hd.max_L = h.max_L; hd.max_E = h.max_E; hd.max_nlev = h.max_nlev;
  size_t meta = 0;
  meta_arrays(s->host, [&](int k, auto& v) { hd.cnt[k] = v.size(); meta += round256(v.size() * sizeof(v[0])); });
  hd.meta_bytes = meta;
  char* dst = static_cast<char*>(d_dst);
  // the header and the host arrays are pageable: the copies below return when the source has been read
  CU(cudaMemcpyAsync(dst, &hd, sizeof(hd), cudaMemcpyHostToDevice, st));
  size_t at = round256(sizeof(ExportHeader));
  cudaError_t e = cudaSuccess;
  meta_arrays(s->host, [&](int, auto& v) {
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpyAsync(dst + at, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice, st);
    at += round256(v.size() * sizeof(v[0]));
  });
  CU(e);
  CU(cudaMemcpyAsync(dst + at, s->blob.p, s->blob.bytes, cudaMemcpyDeviceToDevice, st));
  CU(cudaStreamSynchronize(st));
  return STEMK_OK;
}

int stemk_set_import(stemk_ctx* ctx, const void* d_src, uint64_t bytes, stemk_set** out) {
  if (!ctx || !d_src || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  *out = nullptr;
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (bytes < sizeof(ExportHeader)) return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: buffer too small");
  CU(cudaSetDevice(ctx->device));
  const char* src = static_cast<const char*>(d_src);
  ExportHeader hd;
  CU(cudaMemcpyAsync(&hd, src, sizeof(hd), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hd.magic != kExportMagic) return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: not a set exported by this library version");
  if (round256(sizeof(ExportHeader)) + hd.meta_bytes + round256(hd.blob_bytes) > bytes)
    return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: truncated buffer");
  if (hd.loop_gap != ctx->params.loop_gap || hd.len_band != ctx->params.len_band)
    return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: the set was compiled under another loop gap or length band");
  stemk_set* s = new stemk_set;
  s->device = ctx->device; s->loop_gap = hd.loop_gap; s->len_band = hd.len_band;
  for (int k = 0; k < kBlobArrays; ++k) s->lay[k] = (size_t)hd.lay[k];
  CompiledSet& h = s->host;
  h.max_E4 = hd.max_E4; h.max_fastN = hd.max_fastN; h.n_fast = hd.n_fast; h.max_band_cnt = hd.max_band_cnt;
  h.n_weighted = hd.n_weighted; h.n_simple_cols = hd.n_simple_cols; h.has_dag = hd.has_dag != 0;
  h.max_N = hd.max_N; h.max_L = hd.max_L; h.max_E = hd.max_E; h.max_nlev = hd.max_nlev;
  size_t at = round256(sizeof(ExportHeader));
  cudaError_t e = cudaSuccess;
  try {
    meta_arrays(h, [&](int k, auto& v) {
      v.resize((size_t)hd.cnt[k]);
      if (e == cudaSuccess && !v.empty()) e = cudaMemcpyAsync(v.data(), src + at, v.size() * sizeof(v[0]), cudaMemcpyDeviceToHost, ctx->stream);
      at += round256(v.size() * sizeof(v[0]));
    });
  } catch (const std::bad_alloc&) { delete s; return fail(ctx, STEMK_ERR_NOMEM, "stemk_set_import: out of host memory"); }
  if (e == cudaSuccess && at != round256(sizeof(ExportHeader)) + hd.meta_bytes) { delete s; return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: inconsistent header"); }
  if (e == cudaSuccess) e = s->blob.reserve(std::max<size_t>((size_t)hd.blob_bytes, 256));
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->blob.p, src + at, (size_t)hd.blob_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { s->blob.release(); delete s; return cuda_fail(ctx, e, "stemk_set_import"); }
  if (h.rec.size() != hd.n_recs) { s->blob.release(); delete s; return fail(ctx, STEMK_ERR_ARG, "stemk_set_import: inconsistent header"); }
  make_view(s);
  *out = s;
  return STEMK_OK;
}

uint32_t stemk_set_size(const stemk_set* s) { return s ? (uint32_t)s->host.rec.size() : 0; }

void stemk_set_stats(const stemk_set* s, uint32_t* n_nodes, uint32_t* n_edges, uint32_t* length) {
  if (!s) return;
  for (size_t i = 0; i < s->host.rec.size(); ++i) {
    if (n_nodes) n_nodes[i] = s->host.n_nodes_all[i];
    if (n_edges) n_edges[i] = s->host.n_edges_all[i];
    if (length) length[i] = s->host.rec[i].L;
  }
}

uint64_t stemk_set_device_bytes(const stemk_set* s) { return s && s->blob.p ? (uint64_t)std::max<size_t>(s->host.blob_bytes, 256) : 0; }

// -------------------------------------------------------------------------- pair evaluator
int stemk_pairs_device(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs, const uint32_t* d_xi,
                       const uint32_t* d_yi, double* d_out, void* stream_) {
  if (!ctx || !x || !y) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n_pairs == 0) return STEMK_OK;
  if (!d_xi || !d_yi || !d_out) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  if (!set_usable(ctx, x) || !set_usable(ctx, y)) return fail(ctx, STEMK_ERR_ARG, kSetMismatch);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
  CU(order_after_previous(ctx, st));
  const int kind = ctx->params.kind;
  const bool has_stem = kind_has_stem(kind), has_str = kind_has_string(kind);
  const bool combine = (has_stem && has_str) || kind == STEMK_LSU_STEM || kind == STEMK_LSU_STR;
  double* stem_out = d_out;
  double* str_out = d_out;
  if (combine) {
    if (has_stem) { CU(ctx->tmp_stem.reserve(n_pairs * sizeof(double))); stem_out = (double*)ctx->tmp_stem.p; }
    if (has_str) { CU(ctx->tmp_str.reserve(n_pairs * sizeof(double))); str_out = (double*)ctx->tmp_str.p; }
  }

  if (has_stem) {
    if (n_pairs > 0xffffffffull) return fail(ctx, STEMK_ERR_ARG, "more than 2^32 pairs in one call");
    // ---- classify: trivial pairs are finished, the others go to the general kernel (bucket 0), to the fast
    // kernel's size buckets (1..) or to the unstaged kernel (last bucket), each bucket keeping the caller's pair order
    static const uint32_t kCaps[kMaxFastBuckets] = {256, 320, 384, 448, 512, 640, 768, kFastMaxN};
    const bool any_fast = ctx->use_fast && x->host.n_fast > 0 && y->host.n_fast > 0;
    StemClassify C;
    C.X = x->view; C.Y = y->view; C.xi = d_xi; C.yi = d_yi; C.n_pairs = n_pairs; C.out = stem_out;
    C.count = ctx->d_bucket; C.start = ctx->d_bucket + 16;
    unsigned long long* heads = ctx->d_bucket + 32;
    C.n_caps = 0; C.allow_fast = any_fast;
    if (any_fast)
      for (int b = 0; b < kMaxFastBuckets; ++b) {
        C.caps[C.n_caps++] = kCaps[b];
        if (kCaps[b] >= y->host.max_fastN) break;
      }
    // general kernel: needed unless every record with a DAG is fast-eligible.  Its shared-memory carve-up is sized for
    // the largest records of the two sets; when those leave fewer than kGenMinSlots row slots it is sized for
    // kGenCap* instead and the pairs with a larger record run on the unstaged kernel.
    auto n_dag = [](const CompiledSet& h) { uint32_t n = 0; for (const RecDev& r : h.rec) n += r.N > 0; return n; };
    const bool need_general = !any_fast || x->host.n_fast < n_dag(x->host) || y->host.n_fast < n_dag(y->host);
    uint32_t gen_nx = std::max(1u, x->host.max_N), gen_ny = std::max(1u, y->host.max_N), gen_ey = std::max(1u, y->host.max_E);
    uint32_t gen_lv = std::max(1u, y->host.max_nlev), nslots = 0;
    size_t gen_smem = 0;
    C.big_bucket = -1; C.gen_nx = C.gen_ny = C.gen_ey = 0xffffffffu;
    bool run_staged = need_general;
    if (need_general) {
      stem_config(ctx, gen_nx, gen_ny, gen_ey, gen_lv, &nslots, &gen_smem);
      if (nslots < kGenMinSlots || ctx->force_unstaged) {
        C.big_bucket = 1 + C.n_caps;
        if (ctx->force_unstaged) { C.gen_nx = C.gen_ny = C.gen_ey = 0; run_staged = false; }
        else {
          gen_nx = std::min(gen_nx, kGenCapNx); gen_ny = std::min(gen_ny, kGenCapNy); gen_ey = std::min(gen_ey, kGenCapEy);
          gen_lv = std::min(gen_lv, gen_ny);   // a record has at most one sub-level per node
          C.gen_nx = gen_nx; C.gen_ny = gen_ny; C.gen_ey = gen_ey;
          stem_config(ctx, gen_nx, gen_ny, gen_ey, gen_lv, &nslots, &gen_smem);
          if (!nslots) return fail(ctx, STEMK_ERR_CUDA, "general stem kernel does not fit in shared memory");
        }
      }
    }
    CU(ctx->order.reserve(n_pairs * sizeof(uint32_t)));
    C.order = (uint32_t*)ctx->order.p;
    const int n_buckets = 1 + C.n_caps + (C.big_bucket >= 0 ? 1 : 0);
    CU(launch_classify(C, n_buckets, heads, st));
    ctx->launches += 3;

    if (run_staged) {
      int per_sm = stem_max_ctas_per_sm(gen_smem);
      if (per_sm < 1) return fail(ctx, STEMK_ERR_CUDA, "stem kernel does not fit on an SM");
      per_sm = std::min(per_sm, 4);
      const unsigned long long stride = (unsigned long long)gen_nx * ((gen_ny + 1u) & ~1u);
      // the G0 slabs of all resident CTAs: at most ~4 GB
      const size_t by_mem = std::max<size_t>(1, ((size_t)4 << 30) / (sizeof(double) * (size_t)stride));
      const int grid = (int)std::min(std::min<size_t>(n_pairs, (size_t)ctx->sm_count * per_sm), by_mem);
      CU(ctx->scratch.reserve(sizeof(double) * stride * grid));
      StemLaunch L;
      L.X = x->view; L.Y = y->view; L.xi = d_xi; L.yi = d_yi; L.n_pairs = n_pairs; L.out = stem_out;
      L.counter = heads + 0; L.scratch = (double*)ctx->scratch.p; L.scratch_stride = stride;
      L.pair_tab = ctx->d_pair_tab; L.len_band = ctx->params.len_band; L.nslots = nslots; L.nx_cap = gen_nx; L.ny_cap = gen_ny;
      L.ey_cap = gen_ey; L.lev_cap = gen_lv;
      L.order = C.order; L.n_items_dev = C.count + 0;
      stemk_ctx::Timed tm = timed_begin(ctx, 0, st);
      cudaError_t le = launch_stem(L, grid, gen_smem, st);
      timed_end(ctx, tm, st);
      CU(le);
      ctx->launches += 1;
    }
    if (C.big_bucket >= 0) {
      // unstaged kernel: one CTA per pair in flight, everything in the CTA's global scratch (sized for the largest
      // records of the two sets); as many CTAs as ~4 GB of scratch allow, at most one per SM
      const uint32_t bx = std::max(1u, x->host.max_N), by = std::max(1u, y->host.max_N);
      const unsigned long long stride = stem_unstaged_scratch_doubles(bx, by);
      const size_t by_mem = std::max<size_t>(1, ((size_t)4 << 30) / (sizeof(double) * (size_t)stride));
      const int grid = (int)std::min(std::min<size_t>(n_pairs, (size_t)ctx->sm_count), by_mem);
      CU(ctx->scratch_big.reserve(sizeof(double) * stride * grid));
      StemBigLaunch B;
      B.X = x->view; B.Y = y->view; B.xi = d_xi; B.yi = d_yi; B.out = stem_out; B.order = C.order;
      B.start = C.start; B.count = C.count; B.counter = heads + C.big_bucket; B.bucket = C.big_bucket;
      B.scratch = (double*)ctx->scratch_big.p; B.scratch_stride = stride; B.pair_tab = ctx->d_pair_tab;
      B.len_band = ctx->params.len_band; B.nx_cap = bx; B.ny_cap = by;
      stemk_ctx::Timed tm = timed_begin(ctx, 0, st);
      cudaError_t le = launch_stem_unstaged(B, grid, st);
      timed_end(ctx, tm, st);
      CU(le);
      ctx->launches += 1;
    }
    // the fast kernel only ever sees fast-eligible records: its row flags, slabs and level table are sized for those
    const uint32_t nx_cap = std::max(1u, x->host.max_fastN);
    const uint32_t lev_cap = std::max(1u, std::min(y->host.max_nlev, std::max(1u, y->host.max_fastN)));
    // ---- fast kernel, one launch per size bucket (shared memory and warps per CTA sized for the bucket)
    for (int b = 0; any_fast && b < C.n_caps; ++b) {
      const uint32_t ny_cap = std::min(C.caps[b], std::max(1u, y->host.max_fastN));
      const uint32_t lo = b ? C.caps[b - 1] : 0u;
      uint32_t e4_cap = 4, band_cap = 1;
      bool any = false;
      for (size_t r = 0; r < y->host.rec.size(); ++r) {
        const RecDev& ry = y->host.rec[r];
        if (!(ry.flags & REC_FAST) || ry.N <= lo || ry.N > C.caps[b]) continue;
        e4_cap = std::max(e4_cap, ry.e4); band_cap = std::max(band_cap, y->host.band_cnt[r]); any = true;
      }
      if (!any) continue;
      const size_t budget = std::min<size_t>(ctx->smem_optin, (size_t)227 * 1024) - 256;   // static shared memory
      // one CTA per SM with as many warps as the bucket's rows leave room for
      int best_w = 0; size_t best_smem = 0;
      for (int t = stem_fast_max_warps(ny_cap); t >= 2; --t) if (stem_fast_smem_bytes(t, nx_cap, ny_cap, e4_cap, lev_cap, band_cap) <= budget) { best_w = t; break; }
      if (!best_w) return fail(ctx, STEMK_ERR_NOMEM, "fast stem kernel: record does not fit in shared memory");
      best_smem = stem_fast_smem_bytes(best_w, nx_cap, ny_cap, e4_cap, lev_cap, band_cap);
      const int per_sm = stem_fast_ctas_per_sm(ny_cap, best_w, best_smem);
      if (per_sm < 1) return fail(ctx, STEMK_ERR_CUDA, "fast stem kernel does not fit on an SM");
      const int grid = (int)std::min<size_t>((n_pairs + kFastGroup - 1) / kFastGroup, (size_t)ctx->sm_count);
      const unsigned long long stride = (unsigned long long)kFastGroup * nx_cap * ((ny_cap + 1u) & ~1u);
      CU(ctx->scratch.reserve(sizeof(double) * stride * grid));
      CU(ctx->rowacc.reserve(sizeof(double) * (size_t)kFastGroup * nx_cap * grid));
      StemFastLaunch F;
      F.X = x->view; F.Y = y->view; F.xi = d_xi; F.yi = d_yi; F.out = stem_out; F.order = C.order;
      F.start = C.start; F.count = C.count; F.counter = heads + 1 + b; F.bucket = 1 + b;
      F.scratch = (double*)ctx->scratch.p; F.scratch_stride = stride; F.rowacc = (double*)ctx->rowacc.p; F.pair_tab = ctx->d_pair_tab;
      F.len_band = ctx->params.len_band; F.nx_cap = nx_cap; F.ny_cap = ny_cap; F.e4_cap = e4_cap; F.lev_cap = lev_cap;
      F.band_cap = band_cap; F.prof = nullptr;
#ifdef FAST_PROF
      unsigned long long* d_prof = nullptr;
      cudaMalloc((void**)&d_prof, 16 * sizeof(unsigned long long));
      cudaMemsetAsync(d_prof, 0, 16 * sizeof(unsigned long long), st);
      F.prof = d_prof;
#endif
      stemk_ctx::Timed tm = timed_begin(ctx, 0, st);
      cudaError_t le = launch_stem_fast(F, grid, best_w, best_smem, st);
      timed_end(ctx, tm, st);
      CU(le);
      ctx->launches += 1;
#ifdef FAST_PROF
      {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(d_prof);
        const double nr = h[5] ? (double)h[5] : 1.0;
        std::fprintf(stderr, "fast prof cap %u warps %d smem %zu band_cap %u: rows %llu | cycles per row: total %.0f A %.0f (flag wait %.0f) B1+Z %.0f B2 %.0f C %.0f | of all warp time: blocks %.1f%% group barrier %.1f%%\n",
                     ny_cap, best_w, best_smem, band_cap, h[5], h[0] / nr, h[1] / nr, h[6] / nr, h[2] / nr, h[3] / nr, h[4] / nr,
                     100.0 * h[0] / (h[9] ? (double)h[9] : 1.0), 100.0 * h[8] / (h[9] ? (double)h[9] : 1.0));
      }
#endif
    }
  }
  if (has_str) {
    const uint32_t ly_cap = std::max(1u, y->host.max_L), lx_cap = std::max(1u, x->host.max_L);
    // specialised kernels when the whole launch is uniform, the general one otherwise
    const size_t nxr = x->host.rec.size(), nyr = y->host.rec.size();
    const bool simple = x->host.n_simple_cols == nxr && y->host.n_simple_cols == nyr;
    const bool all_w = x->host.n_weighted == nxr && y->host.n_weighted == nyr;
    const bool none_w = x->host.n_weighted == 0 || y->host.n_weighted == 0;
    const int mode = kind == STEMK_STR_NAIVE ? 2 : (!simple ? 3 : (all_w ? 1 : (none_w ? 0 : 3)));
    int cw, tp;
    string_shape_for(ly_cap, mode, &cw, &tp);
    const int wpc = string_warps_per_cta();
    const int gpw = 32 / tp;                                   // pairs per warp
    const size_t want = (n_pairs + (size_t)wpc * gpw - 1) / ((size_t)wpc * gpw);
    const int grid = (int)std::min<size_t>(want, (size_t)ctx->sm_count * 12);
    const unsigned long long cstride = 2ull * (lx_cap + 2);
    CU(ctx->carry.reserve(sizeof(double) * cstride * grid * wpc * gpw));
    CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), st));
    StringLaunch L;
    L.X = x->view; L.Y = y->view; L.xi = d_xi; L.yi = d_yi; L.n_pairs = n_pairs; L.out = str_out;
    L.counter = ctx->d_counter; L.carry = (double*)ctx->carry.p; L.carry_stride = cstride; L.subst = ctx->d_subst;
    L.gap = ctx->params.gap; L.naive = kind == STEMK_STR_NAIVE; L.pow_cap = std::max(lx_cap, ly_cap) + 1;
    stemk_ctx::Timed tm = timed_begin(ctx, 1, st);
    cudaError_t le = launch_string(L, cw, tp, mode, grid, st);
    timed_end(ctx, tm, st);
    CU(le);
    ctx->launches += 1;
  }
  if (combine) {
    CU(launch_combine(kind, ctx->params.alpha, ctx->params.beta, has_stem ? stem_out : nullptr,
                      has_str ? str_out : nullptr, d_out, n_pairs, st));
    ctx->launches += 1;
  }
  return STEMK_OK;
}

// ---- host <-> device staging ------------------------------------------------------------------------------------
// Results travel through two pinned buffers of the context: the DMA of chunk c+1 runs while the host copies chunk c
// into the caller's (pageable) buffer; one cudaMemcpyAsync per chunk.
namespace {
constexpr size_t kStageBytes = (size_t)8 << 20;

int stage_reserve(stemk_ctx* ctx) {
  for (int k = 0; k < 2; ++k)
    if (!ctx->stage[k]) {
      CU(cudaHostAlloc(&ctx->stage[k], kStageBytes, cudaHostAllocDefault));
      CU(cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
    }
  return STEMK_OK;
}

// dst[0 .. bytes) <- device src, after everything queued on the context's stream; returns when dst is complete
int copy_out(stemk_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return STEMK_OK;
  if (int rc = stage_reserve(ctx)) return rc;
  const size_t n_chunks = (bytes + kStageBytes - 1) / kStageBytes;
  for (size_t c = 0; c <= n_chunks; ++c) {
    if (c < n_chunks) {
      const size_t off = c * kStageBytes, sz = std::min(kStageBytes, bytes - off);
      CU(cudaMemcpyAsync(ctx->stage[c & 1], static_cast<const char*>(d_src) + off, sz, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaEventRecord(ctx->stage_ev[c & 1], ctx->stream));
    }
    if (c > 0) {
      const size_t off = (c - 1) * kStageBytes, sz = std::min(kStageBytes, bytes - off);
      CU(cudaEventSynchronize(ctx->stage_ev[(c - 1) & 1]));
      std::memcpy(static_cast<char*>(dst) + off, ctx->stage[(c - 1) & 1], sz);
    }
  }
  return STEMK_OK;
}

// device dst <- host src through the pinned buffers (index lists of stemk_pairs)
int copy_in(stemk_ctx* ctx, void* d_dst, const void* src, size_t bytes) {
  if (bytes == 0) return STEMK_OK;
  if (int rc = stage_reserve(ctx)) return rc;
  const size_t n_chunks = (bytes + kStageBytes - 1) / kStageBytes;
  for (size_t c = 0; c < n_chunks; ++c) {
    const size_t off = c * kStageBytes, sz = std::min(kStageBytes, bytes - off);
    if (c >= 2) CU(cudaEventSynchronize(ctx->stage_ev[c & 1]));   // the DMA that last read this buffer
    std::memcpy(ctx->stage[c & 1], static_cast<const char*>(src) + off, sz);
    CU(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, ctx->stage[c & 1], sz, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->stage_ev[c & 1], ctx->stream));
  }
  return STEMK_OK;
}

}  // namespace

int stemk_pairs(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs, const uint32_t* xi,
                const uint32_t* yi, double* out) {
  if (!ctx || !x || !y) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n_pairs == 0) return STEMK_OK;
  if (!xi || !yi || !out) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  for (size_t k = 0; k < n_pairs; ++k)
    if (xi[k] >= x->host.rec.size() || yi[k] >= y->host.rec.size()) return fail(ctx, STEMK_ERR_ARG, "pair index out of range");
  CU(cudaSetDevice(ctx->device));
  CU(ctx->idx_x.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->idx_y.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->vals.reserve(n_pairs * sizeof(double)));
  if (int rc = copy_in(ctx, ctx->idx_x.p, xi, n_pairs * sizeof(uint32_t))) return rc;
  if (int rc = copy_in(ctx, ctx->idx_y.p, yi, n_pairs * sizeof(uint32_t))) return rc;
  int rc = stemk_pairs_device(ctx, x, y, n_pairs, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p,
                              (double*)ctx->vals.p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  return copy_out(ctx, out, ctx->vals.p, n_pairs * sizeof(double));
}

// -------------------------------------------------------------------------- Gram drivers
// Rough per-record size key used only to order the work queue (largest first).
static inline double size_key(const stemk_ctx* ctx, const CompiledSet& a, uint32_t i) {
  double c = 0;
  if (kind_has_stem(ctx->params.kind)) c += 2.0 * (double)a.n_nodes_all[i] * a.n_edges_all[i];
  if (kind_has_string(ctx->params.kind)) c += (double)a.rec[i].L * a.rec[i].L;
  return c;
}
static std::vector<uint32_t> size_order(const stemk_ctx* ctx, const CompiledSet& h, const uint32_t* subset, uint32_t n) {
  std::vector<uint32_t> perm(n);
  std::vector<double> key(n);
  for (uint32_t i = 0; i < n; ++i) { perm[i] = i; key[i] = size_key(ctx, h, subset ? subset[i] : i); }
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
  if (subset) for (uint32_t& v : perm) v = subset[v];
  return perm;
}

int stemk_gram(stemk_ctx* ctx, const stemk_set* train, int normalize, double* out) {
  if (!ctx || !train || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  const uint32_t n = (uint32_t)train->host.rec.size();
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n == 0) return STEMK_OK;
  CU(cudaSetDevice(ctx->device));
  const bool timing = ctx->timing != 0;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto tg0 = now();
  if (timing) cudaEventRecord(ctx->ev0, ctx->stream);
  // records sorted by size, big first: the queue then hands out the expensive pairs first.  The y-major pair list
  // itself (for every record b all partners a <= b: consecutive pairs share their y record, which the stem kernel
  // stages once per group of pairs) is written by the device from the permutation: row q owns perm[q] + 1 pairs.
  const std::vector<uint32_t> perm = size_order(ctx, train->host, nullptr, n);
  std::vector<unsigned long long> off(n);
  unsigned long long acc = 0;
  for (uint32_t q = 0; q < n; ++q) { off[q] = acc; acc += (unsigned long long)perm[q] + 1ull; }
  const size_t n_pairs = (size_t)n * (n + 1) / 2;
  CU(ctx->idx_x.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->idx_y.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->vals.reserve(n_pairs * sizeof(double)));
  CU(ctx->matrix.reserve((size_t)n * n * sizeof(double)));
  CU(ctx->perm.reserve(n * sizeof(uint32_t)));
  CU(ctx->offs.reserve(n * sizeof(unsigned long long)));
  CU(cudaMemcpyAsync(ctx->perm.p, perm.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->offs.p, off.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_gram_pairs((const uint32_t*)ctx->perm.p, (const unsigned long long*)ctx->offs.p, n, (uint32_t*)ctx->idx_x.p,
                       (uint32_t*)ctx->idx_y.p, ctx->stream));
  ctx->launches += 1;
  const auto tg1 = now();
  int rc = stemk_pairs_device(ctx, train, train, n_pairs, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p,
                              (double*)ctx->vals.p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  rc = stemk_assemble_device(ctx, n_pairs, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p,
                             (const double*)ctx->vals.p, n, normalize, (double*)ctx->matrix.p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  const auto tg2 = now();
  float span_ms = 0.0f;
  if (timing) {
    cudaEventRecord(ctx->ev1, ctx->stream);
    CU(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&span_ms, ctx->ev0, ctx->ev1);
  }
  const auto tg3 = now();
  rc = copy_out(ctx, out, ctx->matrix.p, (size_t)n * n * sizeof(double));
  if (rc != STEMK_OK) return rc;
  if (timing)
    std::fprintf(stderr, "stemk_gram: %u records: permutation + pair-list launch %.1f ms, enqueue %.1f ms, device %.1f ms (span of the call's work on the stream %.1f ms), D2H through pinned staging %.1f ms\n", n,
                 ms(tg0, tg1), ms(tg1, tg2), ms(tg2, tg3), (double)span_ms, ms(tg3, now()));
  return STEMK_OK;
}

// KernelMatrix::calculate over several devices of ONE process: the reference's MPI path (CalcTrainMatrix with
// `cnt % size == rank`, kernel_matrix.cpp:186-262; Ssend/Recv of every rank's values to rank 0, :495-526; normalisation
// of the gathered matrix, :560-571) with a device in the place of a rank.  One global y-major pair order, position k
// goes to device k % n_ctx (neighbouring pairs cost nearly the same: balanced to a fraction of a percent, and every
// device still runs its expensive pairs first); every device evaluates its share through stemk_pairs_device on its own
// stream, device 0 pulls the shares over NVLink (cudaMemcpyPeerAsync behind an event of the producing stream), puts
// them back into the global order, scatters with the mirror and normalises.
int stemk_gram_multi(stemk_ctx* const* ctxs, const stemk_set* const* sets, int n_ctx, int normalize, double* out) {
  stemk_ctx* ctx = (ctxs && n_ctx > 0) ? ctxs[0] : nullptr;
  if (!ctx || !sets || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  for (int d = 0; d < n_ctx; ++d) {
    if (!ctxs[d] || !sets[d]) return fail(ctx, STEMK_ERR_ARG, "null context or set");
    if (ctxs[d]->device == STEMK_DEVICE_NONE) return no_device(ctx);
    if (!set_usable(ctxs[d], sets[d])) return fail(ctx, STEMK_ERR_ARG, kSetMismatch);
    if (sets[d]->host.rec.size() != sets[0]->host.rec.size() || std::memcmp(&ctxs[d]->params, &ctx->params, sizeof(stemk_params)) != 0)
      return fail(ctx, STEMK_ERR_ARG, "stemk_gram_multi: the contexts must share their kernel parameters and the sets their records");
    for (int q = 0; q < d; ++q) if (ctxs[q]->device == ctxs[d]->device) return fail(ctx, STEMK_ERR_ARG, "stemk_gram_multi: one context per device");
  }
  if (n_ctx == 1) return stemk_gram(ctx, sets[0], normalize, out);
  const uint32_t n = (uint32_t)sets[0]->host.rec.size();
  if (n == 0) return STEMK_OK;
  const uint32_t W = (uint32_t)n_ctx;
  const std::vector<uint32_t> perm = size_order(ctx, sets[0]->host, nullptr, n);
  std::vector<unsigned long long> off(n);
  unsigned long long acc = 0;
  for (uint32_t q = 0; q < n; ++q) { off[q] = acc; acc += (unsigned long long)perm[q] + 1ull; }
  const size_t n_pairs = (size_t)n * (n + 1) / 2;
  const size_t per = (n_pairs + W - 1) / W;
  auto dfail = [&](int d, int rc) { if (d > 0) fail(ctx, rc, std::string("device ") + std::to_string(d) + ": " + stemk_last_error(ctxs[d])); return rc; };
  // ---- every device: the pair list from the permutation, its share of it, the values
  for (uint32_t d = 0; d < W; ++d) {
    stemk_ctx* c = ctxs[d];
    const size_t mine = n_pairs > d ? (n_pairs - d + W - 1) / W : 0;
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess && !c->multi_ev) e = cudaEventCreateWithFlags(&c->multi_ev, cudaEventDisableTiming);
    for (auto [buf, bytes] : {std::pair<DevBuf*, size_t>{&c->idx_x, n_pairs * sizeof(uint32_t)}, {&c->idx_y, n_pairs * sizeof(uint32_t)},
                              {&c->deal_x, per * sizeof(uint32_t)}, {&c->deal_y, per * sizeof(uint32_t)}, {&c->vals, per * sizeof(double)},
                              {&c->perm, n * sizeof(uint32_t)}, {&c->offs, n * sizeof(unsigned long long)}})
      if (e == cudaSuccess) e = buf->reserve(bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->perm.p, perm.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->offs.p, off.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_gram_pairs((const uint32_t*)c->perm.p, (const unsigned long long*)c->offs.p, n, (uint32_t*)c->idx_x.p, (uint32_t*)c->idx_y.p, c->stream);
    if (e == cudaSuccess) e = launch_deal_pairs((const uint32_t*)c->idx_x.p, (const uint32_t*)c->idx_y.p, n_pairs, d, W, (uint32_t*)c->deal_x.p, (uint32_t*)c->deal_y.p, c->stream);
    if (e != cudaSuccess) return dfail((int)d, cuda_fail(c, e, "stemk_gram_multi"));
    c->launches += 2;
    if (mine) {
      const int rc = stemk_pairs_device(c, sets[d], sets[d], mine, (const uint32_t*)c->deal_x.p, (const uint32_t*)c->deal_y.p, (double*)c->vals.p, c->stream);
      if (rc != STEMK_OK) return dfail((int)d, rc);
    }
    e = cudaEventRecord(c->multi_ev, c->stream);
    if (e != cudaSuccess) return dfail((int)d, cuda_fail(c, e, "stemk_gram_multi"));
  }
  // ---- device 0: gather, un-deal, assemble
  CU(cudaSetDevice(ctx->device));
  CU(ctx->gathered.reserve(per * W * sizeof(double)));
  CU(ctx->undealt.reserve(n_pairs * sizeof(double)));
  CU(ctx->matrix.reserve((size_t)n * n * sizeof(double)));
  for (uint32_t d = 0; d < W; ++d) {
    const size_t mine = n_pairs > d ? (n_pairs - d + W - 1) / W : 0;
    if (!mine) continue;
    if (d > 0) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[d]->device) == cudaSuccess && can) {
        const cudaError_t pe = cudaDeviceEnablePeerAccess(ctxs[d]->device, 0);
        if (pe != cudaSuccess) cudaGetLastError();   // already enabled
      }
      CU(cudaStreamWaitEvent(ctx->stream, ctxs[d]->multi_ev, 0));
    }
    CU(cudaMemcpyPeerAsync(static_cast<double*>(ctx->gathered.p) + (size_t)d * per, ctx->device, ctxs[d]->vals.p, ctxs[d]->device,
                           mine * sizeof(double), ctx->stream));
  }
  CU(launch_undeal((const double*)ctx->gathered.p, per, W, n_pairs, (double*)ctx->undealt.p, ctx->stream));
  ctx->launches += 1;
  int rc = stemk_assemble_device(ctx, n_pairs, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p,
                                 (const double*)ctx->undealt.p, n, normalize, (double*)ctx->matrix.p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  return copy_out(ctx, out, ctx->matrix.p, (size_t)n * n * sizeof(double));
}

int stemk_assemble_device(stemk_ctx* ctx, size_t n_pairs, const uint32_t* d_xi, const uint32_t* d_yi,
                          const double* d_vals, uint32_t n, int normalize, double* d_matrix, void* stream_) {
  if (!ctx) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n == 0) return STEMK_OK;
  if (!d_matrix || (n_pairs && (!d_xi || !d_yi || !d_vals))) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
  CU(order_after_previous(ctx, st));
  CU(launch_scatter_square(d_vals, d_xi, d_yi, n_pairs, d_matrix, n, st));
  ctx->launches += 1;
  if (normalize) { CU(launch_normalize_square(d_matrix, n, st)); ctx->launches += 2; }
  return STEMK_OK;
}

// k(x_i, x_i) for i in idx (or all) into the DEVICE vector d_out[n] at the records' own positions
static int diag_device(stemk_ctx* ctx, const stemk_set* set, const uint32_t* idx, uint32_t n_idx, double* d_out, DevBuf* d_idx,
                       DevBuf* d_vals) {
  const uint32_t n = (uint32_t)set->host.rec.size();
  const uint32_t m = idx ? n_idx : n;
  if (m == 0) return STEMK_OK;
  CU(d_idx->reserve(m * sizeof(uint32_t)));
  CU(d_vals->reserve(m * sizeof(double)));
  if (idx) CU(cudaMemcpyAsync(d_idx->p, idx, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  else CU(launch_iota((uint32_t*)d_idx->p, m, ctx->stream));
  int rc = stemk_pairs_device(ctx, set, set, m, (const uint32_t*)d_idx->p, (const uint32_t*)d_idx->p, (double*)d_vals->p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  CU(launch_scatter_vec((const double*)d_vals->p, (const uint32_t*)d_idx->p, m, d_out, ctx->stream));
  ctx->launches += 1 + (idx ? 0 : 1);
  return STEMK_OK;
}

int stemk_diag(stemk_ctx* ctx, const stemk_set* train, const uint32_t* sv_index, uint32_t n_sv, double* out) {
  if (!ctx || !train || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  const uint32_t n = (uint32_t)train->host.rec.size();
  if (n_sv && !sv_index) return fail(ctx, STEMK_ERR_ARG, "null sv_index");
  for (uint32_t k = 0; k < n_sv; ++k) if (sv_index[k] >= n) return fail(ctx, STEMK_ERR_ARG, "sv_index out of range");
  if (n == 0) return STEMK_OK;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->diag.reserve(n * sizeof(double)));
  int rc = diag_device(ctx, train, n_sv ? sv_index : nullptr, n_sv, (double*)ctx->diag.p, &ctx->diag_idx, &ctx->diag_vals);
  if (rc != STEMK_OK) return rc;
  if (n_sv == 0) return copy_out(ctx, out, ctx->diag.p, n * sizeof(double));
  // with an sv_index only those entries of `out` are written (KernelMatrix::diagonal, kernel_matrix.cpp:578-633)
  std::vector<double> v(n_sv);
  rc = copy_out(ctx, v.data(), ctx->diag_vals.p, n_sv * sizeof(double));
  if (rc != STEMK_OK) return rc;
  for (uint32_t k = 0; k < n_sv; ++k) out[sv_index[k]] = v[k];
  return STEMK_OK;
}

int stemk_cross(stemk_ctx* ctx, const stemk_set* test, const stemk_set* train, const uint32_t* sv_index, uint32_t n_sv,
                int normalize, double* out, double* self_out) {
  if (!ctx || !test || !train || !out) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  const uint32_t nt = (uint32_t)test->host.rec.size(), ns = (uint32_t)train->host.rec.size();
  if (n_sv && !sv_index) return fail(ctx, STEMK_ERR_ARG, "null sv_index");
  for (uint32_t k = 0; k < n_sv; ++k) if (sv_index[k] >= ns) return fail(ctx, STEMK_ERR_ARG, "sv_index out of range");
  if (nt == 0 || ns == 0) return STEMK_OK;
  CU(cudaSetDevice(ctx->device));
  // rows k(train_x, test_i): the train record is the FIRST argument (kernel_matrix.cpp:159,168).  Work order: big
  // test records first, big train records first inside each; the list is written by the device from the two
  // permutations.  With an sv_index the device matrix is compact (nt x n_sv, column c = sv_index[c]).
  const uint32_t nc = n_sv ? n_sv : ns;
  const std::vector<uint32_t> tperm = size_order(ctx, test->host, nullptr, nt);
  const std::vector<uint32_t> cperm = size_order(ctx, train->host, n_sv ? sv_index : nullptr, nc);
  const size_t n_pairs = (size_t)nt * nc;
  CU(ctx->idx_x.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->idx_y.reserve(n_pairs * sizeof(uint32_t)));
  CU(ctx->vals.reserve(n_pairs * sizeof(double)));
  CU(ctx->matrix.reserve(n_pairs * sizeof(double)));
  CU(ctx->perm.reserve(((size_t)nt + nc + ns + nc) * sizeof(uint32_t)));
  uint32_t* d_tperm = (uint32_t*)ctx->perm.p;
  uint32_t* d_cperm = d_tperm + nt;
  uint32_t* d_colof = d_cperm + nc;   // train record -> column of the compact matrix
  uint32_t* d_cols = d_colof + ns;    // column of the compact matrix -> train record
  CU(cudaMemcpyAsync(d_tperm, tperm.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_cperm, cperm.data(), nc * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<uint32_t> colof;
  if (n_sv) {
    colof.assign(ns, 0u);
    for (uint32_t c = 0; c < n_sv; ++c) colof[sv_index[c]] = c;
    CU(cudaMemcpyAsync(d_colof, colof.data(), ns * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_cols, sv_index, n_sv * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(launch_cross_pairs(d_tperm, d_cperm, nt, nc, (uint32_t*)ctx->idx_x.p, (uint32_t*)ctx->idx_y.p, ctx->stream));
  int rc = stemk_pairs_device(ctx, train, test, n_pairs, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p,
                              (double*)ctx->vals.p, ctx->stream);
  if (rc != STEMK_OK) return rc;
  CU(launch_scatter_cross((const double*)ctx->vals.p, (const uint32_t*)ctx->idx_x.p, (const uint32_t*)ctx->idx_y.p, n_pairs,
                          n_sv ? d_colof : nullptr, (double*)ctx->matrix.p, nc, ctx->stream));
  ctx->launches += 2;
  if (self_out || normalize) {
    CU(ctx->selfv.reserve(nt * sizeof(double)));
    rc = diag_device(ctx, test, nullptr, 0, (double*)ctx->selfv.p, &ctx->diag_idx, &ctx->diag_vals);
    if (rc != STEMK_OK) return rc;
  }
  if (normalize) {
    // kernel_matrix.cpp:735-748: out_ij /= sqrt(self_i * k(train_j, train_j)); with an sv_index only its columns exist
    CU(ctx->diag.reserve(ns * sizeof(double)));
    rc = diag_device(ctx, train, n_sv ? sv_index : nullptr, n_sv, (double*)ctx->diag.p, &ctx->diag_idx2, &ctx->diag_vals2);
    if (rc != STEMK_OK) return rc;
    CU(launch_normalize_cross((double*)ctx->matrix.p, nt, nc, nc, (const double*)ctx->selfv.p, (const double*)ctx->diag.p,
                              n_sv ? d_cols : nullptr, ctx->stream));
    ctx->launches += 1;
  }
  if (self_out) { rc = copy_out(ctx, self_out, ctx->selfv.p, nt * sizeof(double)); if (rc != STEMK_OK) return rc; }
  if (n_sv == 0) return copy_out(ctx, out, ctx->matrix.p, n_pairs * sizeof(double));
  // with an sv_index only those columns of `out` are written (static row calculate, kernel_matrix.cpp:635-697)
  std::vector<double> v(n_pairs);
  rc = copy_out(ctx, v.data(), ctx->matrix.p, n_pairs * sizeof(double));
  if (rc != STEMK_OK) return rc;
  for (uint32_t i = 0; i < nt; ++i)
    for (uint32_t c = 0; c < n_sv; ++c) out[(size_t)i * ns + sv_index[c]] = v[(size_t)i * n_sv + c];
  return STEMK_OK;
}

// -------------------------------------------------------------------------- work model
int stemk_pair_cost(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs, const uint32_t* xi,
                    const uint32_t* yi, double* cells, double* flops) {
  if (!ctx || !x || !y || !xi || !yi) return fail(ctx, STEMK_ERR_ARG, "null argument");
  const int kind = ctx->params.kind;
  const bool has_stem = kind_has_stem(kind), has_str = kind_has_string(kind) && kind != STEMK_STR_NAIVE;
  const CompiledSet& A = x->host;
  const CompiledSet& B = y->host;
  for (size_t k = 0; k < n_pairs; ++k)
    if (xi[k] >= A.rec.size() || yi[k] >= B.rec.size()) return fail(ctx, STEMK_ERR_ARG, "pair index out of range");
  const uint32_t band = ctx->params.len_band;
  auto one = [&](size_t k) {
    const uint32_t i = xi[k], j = yi[k];
    const RecDev& rx = A.rec[i];
    const RecDev& ry = B.rec[j];
    double c = 0, f = 0;
    if (has_stem) {
      // U_match / U_bf over in-band non-leaf node pairs, with the reference's full degrees (leaf edges included):
      // per x node, the y nodes with |len - len_x| <= band come from y's prefix sums over node length
      double um = 0, ub = 0;
      const double* pd = B.cost_pd.data() + B.cost_off[j];
      const double* pb = B.cost_pb.data() + B.cost_off[j];
      const uint32_t ny_len = (uint32_t)(B.cost_off[j + 1] - B.cost_off[j]) - 1;  // prefix arrays cover len < ny_len
      for (uint32_t a = 0; a < rx.N; ++a) {
        const uint32_t la = A.len[rx.node0 + a];
        uint32_t lo = 0, hi = ny_len;  // [lo, hi) in node length
        if (band != 0) { lo = la > band ? la - band : 0; hi = std::min<uint64_t>(ny_len, (uint64_t)la + band + 1); }
        if (lo >= hi) continue;
        um += (double)A.deg_all[rx.node0 + a] * (pd[hi] - pd[lo]);
        ub += (double)(A.boff[rx.boff0 + a + 1] - A.boff[rx.boff0 + a]) * (pb[hi] - pb[lo]);
      }
      c += (double)A.n_nodes_all[i] * B.n_nodes_all[j];
      f += 2 * um + 3 * ub + 3 * ((double)A.n_nodes_all[i] * B.n_edges_all[j] + (double)A.n_edges_all[i] * B.n_nodes_all[j]);
    }
    if (has_str) {
      const bool w = (rx.flags & REC_HAS_WEIGHT) && (ry.flags & REC_HAS_WEIGHT);
      if (!has_stem) c += (double)rx.L * ry.L;
      f += (w ? 9.0 : 7.0) * rx.L * ry.L;
    }
    if (kind == STEMK_STR_NAIVE) {
      double m = 0;
      const uint8_t* tx = A.text.data() + rx.col0;
      const uint8_t* ty = B.text.data() + ry.col0;
      uint32_t cnt_x[256] = {0};
      for (uint32_t a = 0; a < rx.L; ++a) ++cnt_x[tx[a]];
      for (uint32_t b = 0; b < ry.L; ++b) m += cnt_x[ty[b]];
      c += (double)rx.L * ry.L;
      f += 4.0 * rx.L * ry.L + 3.0 * m;
    }
    if (cells) cells[k] = c;
    if (flops) flops[k] = f;
  };
  const unsigned nt = n_pairs < 4096 ? 1u : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  if (nt <= 1) { for (size_t k = 0; k < n_pairs; ++k) one(k); }
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (size_t k = t; k < n_pairs; k += nt) one(k); });
    for (auto& w : th) w.join();
  }
  return STEMK_OK;
}

void stemk_stats_reset(stemk_ctx* ctx) { if (ctx) { if (ctx->device != STEMK_DEVICE_NONE) timed_resolve(ctx); ctx->launches = 0; ctx->stem_ms = ctx->string_ms = 0; } }
void stemk_stats_get(stemk_ctx* ctx, uint64_t* launches, double* stem_ms, double* string_ms) {
  if (!ctx) return;
  if (ctx->device != STEMK_DEVICE_NONE) timed_resolve(ctx);
  if (launches) *launches = ctx->launches;
  if (stem_ms) *stem_ms = ctx->stem_ms;
  if (string_ms) *string_ms = ctx->string_ms;
}

int stemk_bpla_pairs(stemk_ctx* ctx, const stemk_bpla_params* params, const stemk_bpla_set* x, const stemk_bpla_set* y,
                     size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  if (!ctx || !params || !x || !y) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n_pairs == 0) return STEMK_OK;
  if (!xi || !yi || !out) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  for (const stemk_bpla_set* s : {x, y}) {
    if (s->n_seqs && (!s->col_off || !s->profile)) return fail(ctx, STEMK_ERR_ARG, "null set array");
    if (!params->no_bp && s->n_seqs && (!s->p_left || !s->p_right || !s->p_unpair))
      return fail(ctx, STEMK_ERR_ARG, "base-pairing profiles missing (p_left / p_right / p_unpair)");
  }
  for (size_t k = 0; k < n_pairs; ++k)
    if (xi[k] >= x->n_seqs || yi[k] >= y->n_seqs) return fail(ctx, STEMK_ERR_ARG, "pair index out of range");
  CU(cudaSetDevice(ctx->device));
  std::string err;
  cudaError_t e = run_bpla(*params, *x, *y, n_pairs, xi, yi, out, nullptr, ctx->sm_count, ctx->smem_optin, ctx->stream, &err);
  if (e != cudaSuccess) return err.empty() ? cuda_fail(ctx, e, "BPLA kernel") : fail(ctx, STEMK_ERR_NOMEM, err);
  ctx->launches += 1;
  return STEMK_OK;
}

int stemk_bpla_gradients(stemk_ctx* ctx, const stemk_bpla_params* params, const stemk_bpla_set* x, const stemk_bpla_set* y,
                         size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* value, double* grad) {
  if (!ctx || !params || !x || !y) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (params->no_bp || params->sw)
    return fail(ctx, STEMK_ERR_ARG, "compute_gradients is defined for the base-pairing-profile sum kernel only (no_bp = sw = 0)");
  if (n_pairs == 0) return STEMK_OK;
  if (!xi || !yi || !value || !grad) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  for (const stemk_bpla_set* s : {x, y}) {
    if (s->n_seqs && (!s->col_off || !s->profile)) return fail(ctx, STEMK_ERR_ARG, "null set array");
    if (s->n_seqs && (!s->p_left || !s->p_right || !s->p_unpair))
      return fail(ctx, STEMK_ERR_ARG, "base-pairing profiles missing (p_left / p_right / p_unpair)");
  }
  for (size_t k = 0; k < n_pairs; ++k)
    if (xi[k] >= x->n_seqs || yi[k] >= y->n_seqs) return fail(ctx, STEMK_ERR_ARG, "pair index out of range");
  CU(cudaSetDevice(ctx->device));
  std::string err;
  cudaError_t e = run_bpla(*params, *x, *y, n_pairs, xi, yi, value, grad, ctx->sm_count, ctx->smem_optin, ctx->stream, &err);
  if (e != cudaSuccess) return err.empty() ? cuda_fail(ctx, e, "BPLA gradient kernel") : fail(ctx, STEMK_ERR_NOMEM, err);
  ctx->launches += 1;
  return STEMK_OK;
}

static int nstem_pairs_impl(stemk_ctx* ctx, const stemk_nstem_params* params, uint32_t band, const stemk_nstem_set* x,
                            const stemk_nstem_set* y, size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out,
                            const uint32_t* win_off, const uint32_t* c_low, const uint32_t* c_high);

int stemk_nstem_pairs(stemk_ctx* ctx, const stemk_nstem_params* params, const stemk_nstem_set* x, const stemk_nstem_set* y,
                      size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  return nstem_pairs_impl(ctx, params, 0, x, y, n_pairs, xi, yi, out, nullptr, nullptr, nullptr);
}

int stemk_nstem_pairs_banded(stemk_ctx* ctx, const stemk_nstem_params* params, uint32_t band, const stemk_nstem_set* x,
                             const stemk_nstem_set* y, size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  if (ctx && band == 0) return fail(ctx, STEMK_ERR_ARG, "band must be positive (band 0 is stemk_nstem_pairs)");
  return nstem_pairs_impl(ctx, params, band, x, y, n_pairs, xi, yi, out, nullptr, nullptr, nullptr);
}

int stemk_nstem_pairs_windows(stemk_ctx* ctx, const stemk_nstem_params* params, const stemk_nstem_set* x, const stemk_nstem_set* y,
                              size_t n_pairs, const uint32_t* xi, const uint32_t* yi, const uint32_t* win_off, const uint32_t* c_low,
                              const uint32_t* c_high, double* out) {
  if (ctx && (!win_off || !c_low || !c_high)) return fail(ctx, STEMK_ERR_ARG, "null window array");
  return nstem_pairs_impl(ctx, params, 1, x, y, n_pairs, xi, yi, out, win_off, c_low, c_high);
}

static int nstem_pairs_impl(stemk_ctx* ctx, const stemk_nstem_params* params, uint32_t band, const stemk_nstem_set* x,
                            const stemk_nstem_set* y, size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out,
                            const uint32_t* win_off, const uint32_t* c_low, const uint32_t* c_high) {
  if (!ctx || !params || !x || !y) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (n_pairs == 0) return STEMK_OK;
  if (!xi || !yi || !out) return fail(ctx, STEMK_ERR_ARG, "null buffer");
  for (const stemk_nstem_set* s : {x, y}) {
    if (s->n_seqs && (!s->off || !s->text)) return fail(ctx, STEMK_ERR_ARG, "null set array");
    if (params->bp_mode == 1 && s->n_seqs && (!s->bp_off || !s->bp)) return fail(ctx, STEMK_ERR_ARG, "base-pair probability tables missing");
  }
  if (params->bp_mode != 0 && params->bp_mode != 1) return fail(ctx, STEMK_ERR_ARG, "unknown bp_mode");
  for (size_t k = 0; k < n_pairs; ++k)
    if (xi[k] >= x->n_seqs || yi[k] >= y->n_seqs) return fail(ctx, STEMK_ERR_ARG, "pair index out of range");
  CU(cudaSetDevice(ctx->device));
  std::string err;
  if (win_off) {   // windows: one entry per row 0..lx of every pair, inside the second sequence
    for (size_t k = 0; k < n_pairs; ++k) {
      const uint32_t lx = x->off[xi[k] + 1] - x->off[xi[k]], ly = y->off[yi[k] + 1] - y->off[yi[k]];
      if (win_off[k + 1] < win_off[k] || win_off[k + 1] - win_off[k] != lx + 1u) return fail(ctx, STEMK_ERR_ARG, "window arrays: a pair needs lx + 1 entries");
      for (uint32_t q = win_off[k]; q < win_off[k + 1]; ++q)
        if (c_low[q] > c_high[q] || c_high[q] > ly) return fail(ctx, STEMK_ERR_ARG, "window arrays: need c_low <= c_high <= ly");
    }
  }
  cudaError_t e = run_nstem(*params, *x, *y, n_pairs, xi, yi, out, band, ctx->sm_count, ctx->smem_optin, ctx->stream, &err, win_off, c_low, c_high);
  if (e != cudaSuccess) return err.empty() ? cuda_fail(ctx, e, "naive stem kernel") : fail(ctx, STEMK_ERR_NOMEM, err);
  ctx->launches += 1;
  return STEMK_OK;
}

namespace {
// operator<<(std::ostream&, double) with default flags == printf("%g"): std::to_chars(general, precision 6) produces
// the same digits (it is specified as printf's %.6g in the C locale); non-finite values go through snprintf so
// that the spelling ("nan", "-nan", "inf") is glibc's.
inline char* put_g(char* p, double v) {
  if (!std::isfinite(v)) return p + std::snprintf(p, 32, "%g", v);
  auto r = std::to_chars(p, p + 32, v, std::chars_format::general, 6);
  return r.ptr;
}
inline char* put_u(char* p, uint32_t v) { return std::to_chars(p, p + 12, v).ptr; }
}  // namespace

// -------------------------------------------------------------------------- base-pair probabilities (front end)
void stemk_fold_model_default(stemk_fold_model* m) {
  if (!m) return;
  static const double stack[6][6] = {{-2.4, -3.3, -2.1, -1.4, -2.1, -2.1}, {-3.3, -3.4, -2.5, -1.5, -2.2, -2.4},
                                     {-2.1, -2.5, 1.3, -0.5, -1.4, -1.3},  {-1.4, -1.5, -0.5, 0.3, -0.6, -1.0},
                                     {-2.1, -2.2, -1.4, -0.6, -1.1, -0.9}, {-2.1, -2.4, -1.3, -1.0, -0.9, -1.3}};
  static const double hp[31] = {99, 99, 99, 5.7, 5.6, 5.6, 5.4, 5.9, 5.6, 6.4, 6.5, 6.6, 6.7, 6.78, 6.86, 6.94,
                                7.01, 7.07, 7.13, 7.19, 7.25, 7.3, 7.35, 7.4, 7.44, 7.49, 7.53, 7.57, 7.61, 7.65, 7.69};
  static const double bl[31] = {99, 3.8, 2.8, 3.2, 3.6, 4.0, 4.4, 4.59, 4.7, 4.8, 4.9, 5.0, 5.1, 5.2, 5.3, 5.4,
                                5.5, 5.6, 5.7, 5.8, 5.9, 6.0, 6.1, 6.2, 6.3, 6.4, 6.5, 6.6, 6.7, 6.8, 6.9};
  static const double il[31] = {99, 99, 4.1, 5.1, 1.7, 1.8, 2.0, 2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 2.9, 3.0,
                                3.1, 3.2, 3.3, 3.4, 3.5, 3.6, 3.7, 3.8, 3.9, 4.0, 4.1, 4.2, 4.3, 4.4, 4.5};
  std::memset(m, 0, sizeof(*m));
  m->temperature = 37.0;
  m->pf_scale = -1.0;
  for (int a = 1; a <= 6; ++a) for (int b = 1; b <= 6; ++b) m->stack[a][b] = stack[a - 1][b - 1];
  for (int u = 0; u <= 30; ++u) { m->hairpin[u] = hp[u]; m->bulge[u] = bl[u]; m->interior[u] = il[u]; }
  m->lxc = 1.07856;
  for (int t = 1; t <= 6; ++t)
    for (int a = 0; a < 5; ++a) {
      for (int b = 0; b < 5; ++b) {
        m->mismatch_h[t][a][b] = -(0.3 + 0.1 * ((t + 2 * a + 3 * b) % 9));
        m->mismatch_i[t][a][b] = -(0.1 * ((2 * t + a + 4 * b) % 8)) + (t > 2 ? 0.7 : 0.0);
      }
      m->dangle5[t][a] = a ? -(0.1 + 0.05 * ((t + 3 * a) % 6)) : 0.0;
      m->dangle3[t][a] = a ? -(0.2 + 0.1 * ((2 * t + a) % 7)) : 0.0;
    }
  m->ninio = 0.5; m->max_ninio = 3.0;
  m->terminal_au = 0.5;
  m->ml_closing = 3.4;
  for (int t = 0; t < 8; ++t) m->ml_intern[t] = 0.4 + (t > 2 ? 0.5 : 0.0);
  m->ml_base = 0.0;
}

int stemk_fold_bpp(stemk_ctx* ctx, const stemk_fold_model* model, uint32_t n_seqs, const uint64_t* seq_off, const char* text,
                   double cutoff, uint64_t* n_pairs_total, double* ensemble, double* dense) {
  if (!ctx || !model || (n_seqs && (!seq_off || (!text && seq_off[n_seqs])))) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  if (!(model->temperature > -273.0) || !std::isfinite(model->temperature)) return fail(ctx, STEMK_ERR_ARG, "fold model: temperature");
  CU(cudaSetDevice(ctx->device));
  std::string err;
  ctx->fold_n = 0;
  cudaError_t e;
  const auto t0 = std::chrono::steady_clock::now();
  try { e = run_fold(*model, n_seqs, seq_off, text, cutoff, dense != nullptr, ctx->sm_count, ctx->stream, &ctx->fold, &err,
                     &ctx->fold_scratch.p, &ctx->fold_scratch.bytes); }
  catch (const std::bad_alloc&) { return fail(ctx, STEMK_ERR_NOMEM, "stemk_fold_bpp: out of host memory"); }
  if (e != cudaSuccess) {
    if (err.empty()) return cuda_fail(ctx, e, "base-pair probability kernel");
    return fail(ctx, e == cudaErrorMemoryAllocation ? STEMK_ERR_NOMEM : STEMK_ERR_ARG, err);
  }
  ctx->fold_n = n_seqs;
  ctx->launches += 1;
  ctx->fold_ms += ctx->fold.kernel_ms;
  if (ctx->timing)
    std::fprintf(stderr, "stemk_fold_bpp: %u sequences, kernel %.1f ms, call %.1f ms\n", n_seqs, ctx->fold.kernel_ms,
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  if (n_pairs_total) *n_pairs_total = ctx->fold.pair_off.empty() ? 0 : ctx->fold.pair_off.back();
  if (ensemble) std::copy(ctx->fold.ensemble.begin(), ctx->fold.ensemble.end(), ensemble);
  if (dense) { std::copy(ctx->fold.dense.begin(), ctx->fold.dense.end(), dense); ctx->fold.dense.clear(); ctx->fold.dense.shrink_to_fit(); }
  return STEMK_OK;
}

double stemk_fold_last_ms(const stemk_ctx* ctx) { return ctx ? ctx->fold.kernel_ms : 0.0; }

int stemk_fold_fetch(stemk_ctx* ctx, uint64_t* pair_off, uint32_t* bi, uint32_t* bj, double* bp, double* unpaired) {
  if (!ctx) return fail(ctx, STEMK_ERR_ARG, "null argument");
  const FoldResult& r = ctx->fold;
  if (r.pair_off.size() != (size_t)ctx->fold_n + 1) return fail(ctx, STEMK_ERR_ARG, "stemk_fold_fetch: no result (call stemk_fold_bpp first)");
  if (pair_off) std::copy(r.pair_off.begin(), r.pair_off.end(), pair_off);
  if (bi) std::copy(r.bi.begin(), r.bi.end(), bi);
  if (bj) std::copy(r.bj.begin(), r.bj.end(), bj);
  if (bp) std::copy(r.bp.begin(), r.bp.end(), bp);
  if (unpaired) std::copy(r.unpaired.begin(), r.unpaired.end(), unpaired);
  return STEMK_OK;
}

size_t stemk_format_rows(const double* m, uint32_t n_rows, uint32_t n_cols, size_t ld, const char* const* labels,
                         uint32_t first_cnt, int n_threads, char* out, size_t cap) {
  if (!m || !labels || n_rows == 0) return 0;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<uint32_t>((uint32_t)n_threads, n_rows);
  std::vector<std::string> lines(n_rows);
  auto work = [&](int t) {
    std::vector<char> buf;
    for (uint32_t r = (uint32_t)t; r < n_rows; r += (uint32_t)n_threads) {
      const size_t ll = std::strlen(labels[r]);
      buf.resize(ll + 32 + (size_t)n_cols * 48);
      char* p = buf.data();
      std::memcpy(p, labels[r], ll); p += ll;
      *p++ = ' '; *p++ = '0'; *p++ = ':';
      p = put_u(p, first_cnt + r);
      *p++ = ' ';
      const double* row = m + (size_t)r * ld;
      for (uint32_t j = 0; j < n_cols; ++j) {
        p = put_u(p, j + 1);
        *p++ = ':';
        p = put_g(p, row[j]);
        *p++ = ' ';
      }
      *p++ = '\n';
      lines[r].assign(buf.data(), (size_t)(p - buf.data()));
    }
  };
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  size_t total = 0;
  for (const auto& l : lines) total += l.size();
  if (out && total <= cap) {
    char* p = out;
    for (const auto& l : lines) { std::memcpy(p, l.data(), l.size()); p += l.size(); }
  }
  return total;
}

size_t stemk_format_values(const double* v, size_t n, char* out, size_t cap) {
  if (!v) return 0;
  std::string s;
  s.reserve(n * 14);
  char b[40];
  for (size_t i = 0; i < n; ++i) { char* e = put_g(b, v[i]); *e++ = '\n'; s.append(b, (size_t)(e - b)); }
  if (out && s.size() <= cap) std::memcpy(out, s.data(), s.size());
  return s.size();
}

int stemk_fp64_peak(stemk_ctx* ctx, double seconds, double* tflops) {
  if (!ctx || !tflops) return fail(ctx, STEMK_ERR_ARG, "null argument");
  if (ctx->device == STEMK_DEVICE_NONE) return no_device(ctx);
  CU(cudaSetDevice(ctx->device));
  double* sink = nullptr;
  CU(cudaMalloc((void**)&sink, sizeof(double)));
  const int block = 256, grid = ctx->sm_count * 8, iters = 1 << 16;
  const double flop_per_launch = 2.0 * 8.0 * iters * (double)block * grid;
  double best = 0, spent = 0;
  for (int rep = 0; rep < 1000 && (rep < 3 || spent < seconds); ++rep) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    cudaError_t e = launch_fp64_peak(sink, grid, block, iters, ctx->stream);
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (e != cudaSuccess || cudaEventSynchronize(ctx->ev1) != cudaSuccess) { cudaFree(sink); return cuda_fail(ctx, cudaGetLastError(), "fp64 probe"); }
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    spent += ms * 1e-3;
    if (rep > 0) best = std::max(best, flop_per_launch / (ms * 1e-3) / 1e12);
  }
  cudaFree(sink);
  *tflops = best;
  return STEMK_OK;
}

}  // extern "C"
