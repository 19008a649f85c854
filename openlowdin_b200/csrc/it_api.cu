// it_api.cu -- C ABI (include/lowdin_it.h) of the B200 four-index AO->MO transformation.
//
// Host-side orchestration only: buffer management, the two-half plan, occupied batching,
// kernel launches on one CUDA stream, result download.  The arithmetic lives in it_kernels.cuh.
// Reference path being replaced: src/integralsTransformation/TransformIntegralsE.f90:821-1277,
// :1285-1837 (two-half), TransformIntegralsC.f90:141-471, :728-1168 (window/skip semantics),
// IntTransfD.cpp:125-181, :245-323 (full in-place transform behind the same style of C ABI).
#include "it_kernels.cuh"
#include "it_gemm_tma.cuh"
#include "it_list.cuh"
#include "it_eri.cuh"
#include "../../include/lowdin_it.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace lowdin;

namespace {

std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Species {
  int n = 0, ncols = 0;
  int64_t M = 0, ldc = 0;
  DevBuf C, Cs, pi, pj;  // Cs = C * 2^-53 (exact): B operand of the warp-specialised fused first quarter (bits_to_unscaled)
  DevBuf bs_shells, bs_fn, bs_expo, bs_coef;  // basis of the species (lowdin_it_set_basis): shells, functions (+ norma), primitives
  int bs_nbf = 0;
  std::vector<double> bs_norma;
};

struct AoSet {
  bool valid = false;
  AoSource src{};
  DevBuf data;
  DevBuf fa, fb;  // kind K: pair-vector factors [RANKK][M_a], [RANKK][M_b]
  // list storage (LOWDIN_IT_OPT_AO_LIST): the canonical list as uploaded, one segment per staged piece
  std::vector<DevBuf> segs;
  DevBuf seg_counts;      // entries kept in each segment (device, unsigned long long)
  int swapped = 0;
  bool sharded = false;   // data holds full M-vectors of the rank's own slabs only (uploaded on a communicator)
  void release_list() { for (auto &b : segs) b.release(); segs.clear(); }
};
constexpr int kMaxListSegs = 4096;

// NCCL, resolved lazily so that single-GPU use has no link-time dependency on it.
struct Id128 { char b[128]; };  // ncclUniqueId is passed by value (128 bytes)
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string &err) {
  if (g_nccl.lib) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) { err = "NCCL not found (dlopen libnccl.so.2)"; return false; }
#define NSYM(field, name) *(void **)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + name; return false; }
  NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommDestroy, "ncclCommDestroy")
  NSYM(GroupStart, "ncclGroupStart") NSYM(GroupEnd, "ncclGroupEnd") NSYM(Send, "ncclSend") NSYM(Recv, "ncclRecv") NSYM(AllReduce, "ncclAllReduce")
  NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
  g_nccl.lib = lib;
  return true;
}

// In-process rank group (lowdin_it_comm_init_local): the ranks are handles of ONE process, each driven by its own host thread;
// the all-to-all is a set of direct device-to-device (peer) copies, ordered by CUDA events, and the host threads meet at a barrier.
// The same collective semantics as the NCCL communicator, usable on a single device: the multi-rank logic (slab and slot
// division, chunk agreement, blocked layout) is then exercised by the GPU tests of a one-GPU box.
struct LocalGroup {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t gen = 0;
  int64_t agree[16] = {};
  const double *H[16] = {};
  int device[16] = {};
  cudaEvent_t ready[16] = {}, done[16] = {};
  bool failed = false;
  // false: a rank did not arrive within 5 minutes (it failed before the collective); the group is then unusable
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (failed) return false;
    const uint64_t g = gen;
    if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return true; }
    if (!cv.wait_for(lk, std::chrono::seconds(300), [&] { return gen != g || failed; }) || failed) { failed = true; cv.notify_all(); return false; }
    return true;
  }
  ~LocalGroup() { for (int i = 0; i < 16; ++i) { if (ready[i]) cudaEventDestroy(ready[i]); if (done[i]) cudaEventDestroy(done[i]); } }
};

// One process per GPU on ONE node: the all-to-all between the halves as peer-to-peer DMA instead of NCCL send/recv kernels.  The
// compute kernels of both halves are persistent, one CTA per SM with the shared memory full, so an NCCL kernel that overlaps them
// takes SMs a whole first-half launch was counting on (measured at 2 GPUs, profiles/r02v_*: ~1 s per pass of stall); copy engines
// take none.  Every rank PULLS the rows of its own slots out of every peer's chunk buffer (cudaIpc-mapped), ordered by interprocess
// events; the control plane (barrier, IPC handles) is a POSIX shared-memory segment named after the communicator's unique id.
struct NodeShm {
  std::atomic<int> count, gen, failed;
  struct Rank {
    cudaIpcMemHandle_t mem[2];
    int has[2];
    cudaIpcEventHandle_t ready[2], done[2];
  } r[16];
};
struct NodeLink {
  NodeShm *shm = nullptr;
  int n = 0;
  cudaEvent_t ready[2] = {}, done[2] = {};            // mine (interprocess)
  cudaEvent_t peer_ready[16][2] = {}, peer_done[16][2] = {};
  void *peer_mem[16][2] = {};                          // mapped for the duration of one pass
  bool barrier() {
    NodeShm &S = *shm;
    if (S.failed.load()) return false;
    const int g = S.gen.load();
    if (S.count.fetch_add(1) + 1 == n) { S.count.store(0); S.gen.fetch_add(1); return true; }
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0; S.gen.load() == g; ++spin) {
      if (S.failed.load()) return false;
      if ((spin & 1023) == 1023) {
        sched_yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) { S.failed.store(1); return false; }
      }
    }
    return true;
  }
  void unmap_all() {
    for (int g = 0; g < 16; ++g)
      for (int b = 0; b < 2; ++b)
        if (peer_mem[g][b]) { cudaIpcCloseMemHandle(peer_mem[g][b]); peer_mem[g][b] = nullptr; }
  }
  ~NodeLink() {
    unmap_all();
    for (int b = 0; b < 2; ++b) { if (ready[b]) cudaEventDestroy(ready[b]); if (done[b]) cudaEventDestroy(done[b]); }
    for (int g = 0; g < 16; ++g)
      for (int b = 0; b < 2; ++b) { if (peer_ready[g][b]) cudaEventDestroy(peer_ready[g][b]); if (peer_done[g][b]) cudaEventDestroy(peer_done[g][b]); }
    if (shm) munmap(shm, sizeof(NodeShm));
  }
};

}  // namespace

struct lowdin_it_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  Species sp[8];
  AoSet ao[8][8];
  // AO upload state
  int up_a = -1, up_b = -1, up_swapped = 0;
  DevBuf st[2], up_state;                    // two device staging buffers (copy of piece i+1 overlaps the scatter of piece i); {terminator, bad entry}
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {}, ev_scattered[2] = {};
  bool st_busy[2] = {false, false};
  int up_piece = 0;
  int64_t up_pos = 0;                        // entries pushed so far in this upload (error positions)
  int async_push = 0;                        // 1: pushed host buffers stay untouched until ao_end -> pushes do not wait for their copies
  size_t staging_bytes = (size_t)96 << 20;   // per staging buffer
  // workspaces
  DevBuf T1list, Cw;                         // list-driven first quarter: T1[slab][nu][f] of the pass, window rows Cw[mu][f]
  int ao_list = 0;                           // uploads keep the canonical list instead of the dense tensor (LOWDIN_IT_OPT_AO_LIST)
  DevBuf X, T1t, H, H2, OUT, T3, order, tab, sa, sb, ss, sf, blockcount, blockoff, sums, running, overflow, epsA, epsB, dtmp, agree;
  // results of the last lowdin_it_transform
  DevBuf r_i0, r_i1, r_i2, r_i3, r_v;
  int64_t count = 0;
  int res_conv = -1;
  std::vector<int64_t> res_pair;             // download mode: convention-order index (0-based, over ALL window pairs) of each of this rank's pairs
  std::vector<int64_t> res_seg;              // ... and the number of entries it kept (the rank's list is the concatenation of these segments)
  double timers[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double sink_bytes = 0;                     // bytes handed to the host sink by the last transform
  double launches = 0;
  cudaEvent_t ev[6] = {};
  // multi-GPU
  int rank = 0, nranks = 1;
  void *comm = nullptr;                      // NCCL communicator (one process per GPU)
  std::shared_ptr<LocalGroup> lgroup;        // or: in-process group of handles
  std::unique_ptr<NodeLink> node;            // NCCL form on one node: peer-to-peer DMA exchange (LOWDIN_IT_OPT_EXCHANGE_DMA)
  int exchange_dma = 1;                      // 1 = use it when the node link could be set up, 0 = NCCL send/recv
  bool node_mapped = false;                  // the peers' chunk buffers are mapped (inside the chunk loop of a pass)
  int slab_logB = 5;                         // block-cyclic distribution of the first half's slabs: blocks of 2^slab_logB slabs (it_kernels.cuh, slab_owner)
  void *sink_host[2] = {nullptr, nullptr};   // pinned host ring of the dense-block sink (lowdin_it_transform_stream_sink)
  size_t sink_host_cap = 0;
  size_t sink_group_bytes = (size_t)256 << 20;  // size of one dense result block handed to the sink (LOWDIN_IT_OPT_SINK_BLOCK_BYTES)
  cudaEvent_t ev_q4[2] = {}, ev_d2h[2] = {};
  DevBuf Hx, H2x, coltabx;                   // second set of chunk buffers: the exchange of chunk c overlaps the first half of chunk c + 1
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_fh[2] = {}, ev_ex[2] = {}, ev_sh[2] = {};
  std::vector<cudaEvent_t> chunk_ev;         // four timing events per chunk (first half begin/end, second half begin/end)
  int overlap_exchange = 1;                  // N > 1 ranks: double-buffered chunks, exchange on its own stream under the next chunk's first half: 0 never, 1 when memory allows wide chunks, 2 always
  DevBuf coltab;                             // column table of the chunk after the exchange (SRC_RECT_TABLE)
  DevBuf seg;                                // download mode: kept entries per window pair (convention order)
  size_t workspace_bytes = (size_t)1 << 30;  // target size of the X / T1t batch buffers
  int q1_variant = 5;                        // fused first quarter: 1 = shared-memory ring for the coefficient window, 2 = L1 path, no barrier, 3 = warp-specialised
  int gemm_variant = 2;                      // quarter-transform GEMM: 1 = cp.async ring (dgemm_tn_kernel), 2 = TMA + mbarrier persistent (dgemm_tma_kernel)
  int num_sms = 148;
  int q3_two_cta = 0;                        // third quarter: products with K <= this value run as two 4-warp CTAs per SM (0 = off; LOWDIN_IT_OPT_Q3_TWO_CTA)
  int gemm_tall = 0;                         // TMA GEMM: 192 x 64 tiles when the rows are a multiple of 192 plus a few (probe; LOWDIN_IT_OPT_GEMM_TALL)
  int split_row_tail = 1;                    // TMA GEMM: run the <= 80-row tail of a few-rows x many-columns product as a swapped second launch
  int stored_fused = 1;                      // stored AO tensors: 1 = fused unpack + first quarter (q1_load_ws5_kernel), 0 = expansion kernel + DMMA GEMM
  int q1_dbg = 0;                            // probe switches of the warp-specialised first quarter (Q1WsArgs::dbg)
  int q3_red = 0;                            // third-quarter accumulation: 0 = staged read-modify-write epilogue, 1 = red.global.add.f64 (EpiAccRed)
  int frag_perm = 1;                         // TMA kernels: 1 = conflict-free fragment-row permutation (it_gemm_tma.cuh, frag_row); validated on B200 in round 2 (bit-identical, +2.4 % per pass)
  int bench_gen = 1;                         // generator kind used by lowdin_it_kernel_bench kind 2
  int64_t chunk_cols_limit = 0;              // >0: cap on AO-pair columns per chunk (tests force many chunks with it)
  // per-kernel-category device timing (lowdin_it_set_profiling): CUDA event pairs around every launch
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_pool;
  struct ProfRec { int cat, e0, e1; };
  std::vector<ProfRec> prof_recs;
  int prof_used = 0;
  double prof_ms[8] = {0}, prof_cnt[8] = {0}, prof_work[8] = {0};
};

namespace {

int fail(lowdin_it_handle h, const std::string &msg) {
  if (h) h->err = msg; else g_create_error = msg;
  if (h && h->lgroup) {  // a rank of an in-process group that fails releases its peers from the next collective
    std::lock_guard<std::mutex> lk(h->lgroup->m);
    h->lgroup->failed = true;
    h->lgroup->cv.notify_all();
  }
  return 1;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(h, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

inline int64_t npairs(int64_t n) { return n * (n + 1) / 2; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: every call site keeps one bit per device
// (handles on several GPUs of one process each get the opt-in), set under a mutex (handles may live on different threads).
std::mutex g_attr_mutex;
template <class K>
cudaError_t ensure_dyn_smem(K kern, size_t smem, int device, uint64_t &done_mask) {
  std::lock_guard<std::mutex> lk(g_attr_mutex);
  const uint64_t bit = 1ull << (device & 63);
  if (done_mask & bit) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) done_mask |= bit;
  return e;
}
inline int64_t roundup2(int64_t x) { return (x + 1) & ~int64_t(1); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- per-category kernel timing ---------------------------------------------------------------
// categories: 0 expand(1st half) 1 Q1 2 Q2 3 expand(2nd half) 4 Q3 5 Q4 6 consume 7 exchange
void prof_drain(lowdin_it_handle h) {
  if (h->prof_recs.empty()) return;
  cudaStreamSynchronize(h->stream);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  for (auto &r : h->prof_recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->prof_pool[r.e0], h->prof_pool[r.e1]) == cudaSuccess) { h->prof_ms[r.cat] += ms; h->prof_cnt[r.cat] += 1; }
  }
  h->prof_recs.clear();
  h->prof_used = 0;
}
struct ProfScope {
  lowdin_it_handle h; int idx = -1; cudaStream_t st;
  ProfScope(lowdin_it_handle h_, int cat, double work, cudaStream_t st_ = nullptr) : h(h_), st(st_ ? st_ : h_->stream) {
    if (!h->prof_on) return;
    if (h->prof_used + 2 > 8192) prof_drain(h);
    while ((int)h->prof_pool.size() < h->prof_used + 2) { cudaEvent_t e; cudaEventCreate(&e); h->prof_pool.push_back(e); }
    idx = (int)h->prof_recs.size();
    h->prof_recs.push_back({cat, h->prof_used, h->prof_used + 1});
    h->prof_used += 2;
    h->prof_work[cat] += work;
    cudaEventRecord(h->prof_pool[h->prof_recs[idx].e0], st);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof_pool[h->prof_recs[idx].e1], st); }
};

// ---- GEMM dispatch ----------------------------------------------------------------------------
template <int BM, int BN, int WM, int WN, class Epi>
cudaError_t launch_gemm_cfg(lowdin_it_handle h, const GemmArgs &g, const Epi &epi) {
  constexpr int ST = 3;
  constexpr size_t smem = (size_t)ST * (BM + BN) * 20 * sizeof(double);
  auto kern = dgemm_tn_kernel<BM, BN, WM, WN, ST, Epi>;
  static uint64_t configured = 0;
  if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured); e != cudaSuccess) return e;
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM), 1);
  kern<<<grid, WM * WN * 32, smem, h->stream>>>(g, epi);
  h->launches += 1;
  return cudaGetLastError();
}

// ---- TMA + mbarrier persistent GEMM (it_gemm_tma.cuh) -----------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (TensorMapEncodeFn)p;
  }();
  return fn;
}
// [rows][K] FP64 operand, K contiguous, row stride ld doubles; box = 16 doubles x box_rows rows, 128-byte swizzle, zero fill
bool make_operand_map(CUtensorMap *m, const double *base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  TensorMapEncodeFn fn = tensor_map_encoder();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool tma_eligible(const GemmArgs &g) {
  return ((uintptr_t)g.A % 16 == 0) && ((uintptr_t)g.B % 16 == 0) && (g.lda % 2 == 0) && (g.ldb % 2 == 0) && g.K >= 1 &&
         g.strideA == 0 && g.strideB == 0 && tensor_map_encoder() != nullptr;
}

template <int BM, int BN, int WM, int WN, class Epi, int OCC = 1>
cudaError_t launch_gemm_tma_cfg(lowdin_it_handle h, const GemmArgs &g, const Epi &epi) {
  constexpr int TNW = BN / WN / 8;
  constexpr bool staged = Epi::kRowCoalesced && (BM / WM / 8 == 4);  // row-coalescing epilogue: 32-row warp tiles only
  constexpr size_t staging = staged ? (size_t)WM * WN * TNW * 8 * TMA_STAGE_LDM * 8 : 0;
  // OCC CTAs share the SM's 228 KB (1 KB of it reserved per CTA)
  constexpr int ST_RAW = (OCC == 1) ? (int)((200704 + (staged ? 26624 : 0) - staging) / ((BM + BN) * 128))
                                    : (int)((233472 / OCC - 1024 - 1024 - 128 - staging) / ((BM + BN) * 128));
  constexpr int ST = ST_RAW > 8 ? 8 : ST_RAW;
  static_assert(ST >= 2, "ring too short");
  constexpr size_t smem = tma_gemm_smem_bytes<BM, BN, ST>(staged ? WM * WN : 0, TNW);
  static_assert(smem <= 232448 / OCC, "shared memory per CTA");
  const int perm = h->frag_perm ? 1 : 0;
  auto kern = perm ? dgemm_tma_kernel<BM, BN, WM, WN, ST, Epi, true, OCC> : dgemm_tma_kernel<BM, BN, WM, WN, ST, Epi, false, OCC>;
  static uint64_t configured[2] = {0, 0};
  if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured[perm]); e != cudaSuccess) return e;
  CUtensorMap mapA, mapB;
  if (!make_operand_map(&mapA, g.A, g.M, g.K, g.lda, BM) || !make_operand_map(&mapB, g.B, g.N, g.K, g.ldb, BN)) return cudaErrorInvalidValue;
  const int64_t ntiles = ceil_div(g.M, BM) * ceil_div(g.N, BN);
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)OCC * h->num_sms);
  kern<<<grid, WM * WN * 32, smem, h->stream>>>(mapA, mapB, TmaGemmShape{g.M, g.N, g.K}, epi);
  h->launches += 1;
  return cudaGetLastError();
}

template <class Epi>
cudaError_t launch_gemm_one(lowdin_it_handle h, const GemmArgs &g, const Epi &epi, bool tma, bool tall = false) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  const int n = g.N;
  if constexpr (Epi::kSplitRowTail)
    if (tall && tma) return launch_gemm_tma_cfg<192, 64, 4, 2>(h, g, epi);  // 48 x 32 warp tiles: 10 fragment loads per 48 DMMAs
  if constexpr (Epi::kRowCoalesced) {
    // third-quarter accumulation with a short K (the rank-pc update of a chunk): per 128 x 80 tile the
    // read-modify-write of the accumulators takes as long as the DMMAs; two 4-warp CTAs per SM overlap the two
    // (LOWDIN_IT_OPT_Q3_TWO_CTA; same warp tiles, same arithmetic)
    if (tma && h->q3_two_cta && g.K <= h->q3_two_cta && n > 32) {
      const int64_t q80 = ceil_div(n, 80) * 80, q64 = ceil_div(n, 64) * 64;
      if (n > 64 && q80 <= q64) return launch_gemm_tma_cfg<64, 80, 2, 2, Epi, 2>(h, g, epi);
      return launch_gemm_tma_cfg<64, 64, 2, 2, Epi, 2>(h, g, epi);
    }
  }
#define LOWDIN_GEMM_CFG(BM, BN, WM, WN) (tma ? launch_gemm_tma_cfg<BM, BN, WM, WN>(h, g, epi) : launch_gemm_cfg<BM, BN, WM, WN>(h, g, epi))
  if (n <= 8) return LOWDIN_GEMM_CFG(256, 8, 8, 1);
  if (n <= 16) return LOWDIN_GEMM_CFG(256, 16, 8, 1);
  if (n <= 32) return LOWDIN_GEMM_CFG(256, 32, 8, 1);
  if (n <= 64) return LOWDIN_GEMM_CFG(128, 64, 4, 2);
  // pick the N tile with the least padding (ties -> the larger tile)
  const int64_t p128 = ceil_div(n, 128) * 128, p80 = ceil_div(n, 80) * 80, p64 = ceil_div(n, 64) * 64;
  // very wide outputs: padding is negligible, take the tile with the best DMMA : shared-load ratio.
  // (The row-coalescing epilogue works on 32-row warp tiles, which the 128x128 configuration does not have.)
  const bool allow128 = !(tma && Epi::kRowCoalesced);
  if (allow128 && (n >= 2048 || (p128 <= p80 && p128 <= p64))) return LOWDIN_GEMM_CFG(128, 128, 2, 4);
  if (p80 <= p64) return LOWDIN_GEMM_CFG(128, 80, 4, 2);
  return LOWDIN_GEMM_CFG(128, 64, 4, 2);
#undef LOWDIN_GEMM_CFG
}

template <class Epi>
cudaError_t launch_gemm(lowdin_it_handle h, const GemmArgs &g, const Epi &epi) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  const bool tma = (h->gemm_variant == 2) && tma_eligible(g);
  if constexpr (Epi::kSplitRowTail) {
    // Few rows x very many columns (second and fourth quarter: rows = the second-contracted window, e.g. 1350 virtuals):
    // 128-row tiles would pad 1350 to 1408 (4 % of the DMMAs on zeros).  The full 128-row tiles run as they are; the
    // row tail (<= 80 rows) runs as a second launch with the operands exchanged, so that the tail becomes the N
    // dimension and gets an 8..80-wide tile.  Only when the tail launch itself has >= 3 waves of 128-row tiles: measured
    // on B200 (profiles/r01f_variant_probe.log) 1350 x 89440 x 1500 gains 1.7 %, 1350 x 8400 x 1500 would lose 11 %.
    // Rows that are nearly a multiple of 192 (1350 virtuals = 7 x 192 + 6): 192 x 64 tiles leave a tail of a few rows instead of
    // 70 (LOWDIN_IT_OPT_GEMM_TALL).
    const int tail192 = g.M % 192, main192 = g.M - tail192;
    if (tma && h->gemm_tall && ceil_div(g.N, 64) >= 3 * (int64_t)h->num_sms && main192 >= 192 && tail192 <= 16 && (g.M % 128) > 16) {
      GemmArgs gm = g;
      gm.M = main192;
      cudaError_t e = launch_gemm_one(h, gm, epi, true, true);
      if (e != cudaSuccess || tail192 == 0) return e;
      GemmArgs gt{g.B, g.A + (int64_t)main192 * g.lda, g.N, tail192, g.K, g.ldb, g.lda, 0, 0};
      return launch_gemm_one(h, gt, EpiSwapped<Epi>{epi, main192}, tma_eligible(gt));
    }
    const int tail = g.M % 128, main = g.M - tail;
    if (tma && h->split_row_tail && ceil_div(g.N, 128) >= 3 * (int64_t)h->num_sms && main >= 128 && tail > 0 && tail <= 80) {
      GemmArgs gm = g;
      gm.M = main;
      cudaError_t e = launch_gemm_one(h, gm, epi, true);
      if (e != cudaSuccess) return e;
      GemmArgs gt{g.B, g.A + (int64_t)main * g.lda, g.N, tail, g.K, g.ldb, g.lda, 0, 0};
      return launch_gemm_one(h, gt, EpiSwapped<Epi>{epi, main}, tma_eligible(gt));
    }
  }
  return launch_gemm_one(h, g, epi, tma);
}

// ---- plan -------------------------------------------------------------------------------------
struct Half {
  int nc = 0;                 // basis size of the species contracted in this half
  const double *C = nullptr;  // device coefficients, column-major, ldc
  const double *Cs = nullptr; // the same scaled by 2^-53 (generated-source first quarter)
  int64_t ldc = 0;
  int lf = 1, nf = 0;         // first-contracted window (1-based start, count)
  int ls = 1, ns = 0;         // second-contracted window
  bool first_is_conv_second = true;  // first-contracted == the convention's second-listed index (q / s)
};

struct Plan {
  int a = 0, b = 0, conv = 0, symmetric = 0;
  bool intra = true;
  int win[8];
  Half h1, h2;
  int64_t nslabs1 = 0;     // slabs of the first half (M of the non-contracted species)
  AoSource src{};
  // all needed first pairs in convention order: window-relative (second, first) indices and orbital numbers ((i,j) or (p,q))
  std::vector<int> pairs_s, pairs_f, pairs_a, pairs_b;
  int max_slots_per_f = 0;  // largest number of needed pairs sharing one first-contracted index
};

// One occupied-batch pass: first-contracted indices [f0, f0+nfb).  Slots (the needed window pairs of the
// pass) are numbered f-major, so that the pairs sharing one first-contracted index are consecutive: they
// are finished together (fourth quarter + consumer), owned by one rank, and contain each other's
// exchange partners.  `order` lists the slots in the reference's loop order for the download.
struct PassTables {
  int f0 = 0, nfb = 0, nslots = 0;
  std::vector<int32_t> table, sa, sb, ss, sf, order;
  std::vector<int64_t> order_conv;  // convention-order index (over all window pairs of the plan) of order[k]
  std::vector<int> fbeg;  // [nfb+1] first slot of each f
};

int build_plan(lowdin_it_handle h, int a, int b, const int win[8], int conv, int symmetric, Plan &pl) {
  if (a < 0 || a > 7 || b < 0 || b > 7) return fail(h, "species slot out of range");
  if (!h->sp[a].n || !h->sp[b].n) return fail(h, "species not set");
  if (!h->ao[a][b].valid) return fail(h, "AO integrals for this species pair were not uploaded");
  if (conv != LOWDIN_IT_CONV_C && conv != LOWDIN_IT_CONV_E) return fail(h, "unknown output convention");
  pl.a = a; pl.b = b; pl.conv = conv; pl.symmetric = symmetric; pl.intra = (a == b);
  memcpy(pl.win, win, sizeof(int) * 8);
  const Species &A = h->sp[a], &B = h->sp[b];
  const int lim[4] = {A.ncols, A.ncols, B.ncols, B.ncols};
  for (int w = 0; w < 4; ++w) {
    if (win[2 * w] < 1) return fail(h, "window lower bound < 1");
    if (win[2 * w + 1] > lim[w]) return fail(h, "window upper bound exceeds the number of orbitals");
  }
  auto cnt = [&](int w) { return std::max(0, win[2 * w + 1] - win[2 * w] + 1); };
  // first pair (p|i , q|j) on species a; second pair (r|k , s|l) on species b.
  // The smaller window is contracted first (ties: the convention's second index, as in E.f90:1081).
  auto setup = [&](Half &hf, const Species &S, int w_first_listed, int w_second_listed) {
    hf.nc = S.n; hf.C = S.C.as<double>(); hf.Cs = S.Cs.as<double>(); hf.ldc = S.ldc;
    const int n1 = cnt(w_first_listed), n2 = cnt(w_second_listed);
    hf.first_is_conv_second = (n2 <= n1);
    const int wf = hf.first_is_conv_second ? w_second_listed : w_first_listed;
    const int ws = hf.first_is_conv_second ? w_first_listed : w_second_listed;
    hf.lf = win[2 * wf]; hf.nf = cnt(wf); hf.ls = win[2 * ws]; hf.ns = cnt(ws);
  };
  setup(pl.h1, A, 0, 1);
  setup(pl.h2, B, 2, 3);
  pl.nslabs1 = B.M;
  pl.src = h->ao[a][b].src;
  if (h->nranks > 1) {
    // the first half runs over the rank's own slabs (block-cyclic, fixed at upload): stored tensors must be the row-sharded ones
    if ((pl.src.kind == SRC_SYM_PACKED || pl.src.kind == SRC_RECT) && !h->ao[a][b].sharded)
      return fail(h, "the AO integrals of this species pair were uploaded before the communicator existed; upload them again");
    pl.src.logB = h->slab_logB; pl.src.G = h->nranks; pl.src.rank = h->rank;
  } else if (h->ao[a][b].sharded) return fail(h, "the AO integrals of this species pair are one rank's share of a communicator upload");
  // needed first pairs, convention order
  pl.pairs_s.clear(); pl.pairs_f.clear(); pl.pairs_a.clear(); pl.pairs_b.clear();
  std::vector<int> per_f(std::max(pl.h1.nf, 1), 0);
  for (int x = win[0]; x <= win[1]; ++x)
    for (int y = win[2]; y <= win[3]; ++y) {
      bool keep = (conv == LOWDIN_IT_CONV_E) ? (y <= x) : !(symmetric && y < x);
      if (!keep) continue;
      pl.pairs_a.push_back(x); pl.pairs_b.push_back(y);
      if (pl.h1.first_is_conv_second) { pl.pairs_s.push_back(x - win[0]); pl.pairs_f.push_back(y - win[2]); }
      else { pl.pairs_s.push_back(y - win[2]); pl.pairs_f.push_back(x - win[0]); }
      per_f[pl.pairs_f.back()]++;
    }
  pl.max_slots_per_f = 0;
  for (int c : per_f) pl.max_slots_per_f = std::max(pl.max_slots_per_f, c);
  return 0;
}

void build_pass(const Plan &pl, int f0, int nfb, PassTables &pt) {
  pt.f0 = f0; pt.nfb = nfb;
  const int ns = pl.h1.ns;
  pt.table.assign((size_t)std::max(ns, 1) * nfb, -1);
  std::vector<int> conv_of((size_t)std::max(ns, 1) * nfb, -1);  // (s,f) -> index in convention order
  for (size_t k = 0; k < pl.pairs_s.size(); ++k) {
    int f = pl.pairs_f[k] - f0;
    if (f < 0 || f >= nfb) continue;
    conv_of[(size_t)pl.pairs_s[k] * nfb + f] = (int)k;
  }
  pt.sa.clear(); pt.sb.clear(); pt.ss.clear(); pt.sf.clear(); pt.order.clear();
  pt.fbeg.assign(nfb + 1, 0);
  std::vector<std::pair<int, int>> byconv;  // (convention index, slot)
  int slot = 0;
  for (int f = 0; f < nfb; ++f) {
    pt.fbeg[f] = slot;
    for (int s_ = 0; s_ < ns; ++s_) {
      const int k = conv_of[(size_t)s_ * nfb + f];
      if (k < 0) continue;
      pt.table[(size_t)s_ * nfb + f] = slot;
      pt.sa.push_back(pl.pairs_a[k]); pt.sb.push_back(pl.pairs_b[k]);
      pt.ss.push_back(s_); pt.sf.push_back(f);
      byconv.push_back({k, slot});
      ++slot;
    }
  }
  pt.fbeg[nfb] = slot;
  pt.nslots = slot;
  std::sort(byconv.begin(), byconv.end());
  pt.order_conv.clear();
  for (auto &e : byconv) { pt.order.push_back(e.second); pt.order_conv.push_back(e.first); }
}

// The transform's workspaces (chunk buffers, accumulators, result blocks) grow to what the largest transform so far needed --
// up to all of the free HBM -- and are kept for the next one.  Calls that allocate INPUT storage (a stored AO tensor) give them
// back first: they are re-created on demand, the tensor is not.
void release_workspaces(lowdin_it_handle h) {
  cudaStreamSynchronize(h->stream);
  DevBuf *ws[] = {&h->X, &h->T1t, &h->H, &h->H2, &h->Hx, &h->H2x, &h->OUT, &h->T3, &h->T1list, &h->coltab, &h->coltabx,
                  &h->r_i0, &h->r_i1, &h->r_i2, &h->r_i3, &h->r_v, &h->dtmp};
  for (DevBuf *b : ws) b->release();
}

int upload_i32(lowdin_it_handle h, DevBuf &buf, const std::vector<int32_t> &v) {
  CK(buf.ensure(std::max<size_t>(v.size(), 1) * sizeof(int32_t)));
  if (!v.empty()) CK(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// ---- kernel launch helpers --------------------------------------------------------------------
int launch_expand(lowdin_it_handle h, const AoSource &src, int64_t slab0, int64_t bc, int n, int r0, int nrows, int c0, int ncols,
                  int64_t colbase, int ldx, double *X) {
  if (bc <= 0 || nrows <= 0 || ncols <= 0) return 0;
  if (bc > 65535) return fail(h, "slab batch exceeds gridDim.z");
  dim3 grid((unsigned)ceil_div(ldx, 32), (unsigned)ceil_div(nrows, 32), (unsigned)bc);
  if (n >= 65536) return fail(h, "basis too large for the 32-bit pair arithmetic of the expansion kernel");
  switch (src.kind) {
    case SRC_SYM_PACKED: expand_block_kernel<SRC_SYM_PACKED><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
    case SRC_RECT: expand_block_kernel<SRC_RECT><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
    case SRC_RECT_TABLE: expand_block_kernel<SRC_RECT_TABLE><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
    case SRC_HASH_SYM: expand_block_kernel<SRC_HASH_SYM><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
    case SRC_RANKK: expand_block_kernel<SRC_RANKK><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
    default: expand_block_kernel<SRC_HASH_RECT><<<grid, 256, 0, h->stream>>>(src, slab0, n, r0, nrows, c0, ncols, colbase, ldx, X); break;
  }
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

template <int TN>
cudaError_t launch_q1_gen_cfg(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc,
                              int nfb, double *T1t, int64_t ldt) {
  dim3 grid((unsigned)ceil_div(nc, 128), (unsigned)bc);
  if (h->q1_variant == 1) {  // coefficient window through a shared-memory ring
    constexpr int ST = 4;
    constexpr size_t smem = (size_t)ST * (TN * 8) * 20 * sizeof(double);
    auto kern = q1_gen_smem_kernel<TN, ST>;
    static uint64_t configured = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured); e != cudaSuccess) return e;
    kern<<<grid, 256, smem, h->stream>>>(src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
  } else {
    q1_gen_kernel<TN><<<grid, 256, 0, h->stream>>>(src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
  }
  h->launches += 1;
  return cudaGetLastError();
}

// warp-specialised variant (q1_gen_ws_kernel): generator warps + DMMA warps, coefficient window through TMA
template <int TN, int KIND, int GEN>
cudaError_t launch_q1_ws_cfg(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc,
                             int nfb, double *T1t, int64_t ldt) {
  constexpr int ST = 6;
  constexpr size_t smem = q1_ws_smem_bytes<TN, ST>();
  const bool v4 = (h->q1_variant == 4);  // 8 DMMA warps with register double-buffering + 4 vectorised generator warps
  const bool perm = h->frag_perm != 0;
  auto kern = v4 ? (perm ? q1_gen_ws2_kernel<TN, ST, KIND, GEN, true> : q1_gen_ws2_kernel<TN, ST, KIND, GEN, false>)
                 : (perm ? q1_gen_ws_kernel<TN, ST, KIND, GEN, true> : q1_gen_ws_kernel<TN, ST, KIND, GEN, false>);
  static uint64_t configured[4] = {0, 0, 0, 0};
  const int ci = (v4 ? 2 : 0) + (perm ? 1 : 0);
  if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured[ci]); e != cudaSuccess) return e;
  CUtensorMap mapB;
  if (!make_operand_map(&mapB, Cf, nfb, nc, ldc, TN * 8)) return cudaErrorInvalidValue;
  const int64_t ntiles = ceil_div(nc, 128) * (int64_t)bc;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, h->num_sms);
  Q1WsArgs q{slab0, src.logB, src.G, src.rank, bc, nc, nfb, (uint32_t)(KIND == SRC_HASH_SYM ? src.M : src.aux), src.seed, T1t, ldt, h->q1_dbg};
  kern<<<grid, v4 ? 384 : 512, smem, h->stream>>>(mapB, q);
  h->launches += 1;
  return cudaGetLastError();
}
// variant 5 (q1_gen_ws5_kernel): 256-row tiles, 32-row DMMA warps, setmaxnreg
template <int TN, int KIND, int GEN>
cudaError_t launch_q1_ws5_cfg(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc,
                              int nfb, double *T1t, int64_t ldt) {
  constexpr int ST = 5;
  constexpr size_t smem = q1_ws5_smem_bytes<TN, ST>();
  static_assert(smem <= 232448, "shared memory per CTA");
  auto kern = q1_gen_ws5_kernel<TN, ST, KIND, GEN>;
  static uint64_t configured = 0;
  if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured); e != cudaSuccess) return e;
  CUtensorMap mapB;
  if (!make_operand_map(&mapB, Cf, nfb, nc, ldc, TN * 8)) return cudaErrorInvalidValue;
  const int64_t ntiles = ceil_div(nc, 256) * (int64_t)bc;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, h->num_sms);
  Q1WsArgs q{slab0, src.logB, src.G, src.rank, bc, nc, nfb, (uint32_t)(KIND == SRC_HASH_SYM ? src.M : src.aux), src.seed, T1t, ldt, h->q1_dbg};
  kern<<<grid, 512, smem, h->stream>>>(mapB, q);
  h->launches += 1;
  return cudaGetLastError();
}
template <int TN>
cudaError_t launch_q1_ws(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc, int nfb,
                         double *T1t, int64_t ldt) {
  if (h->q1_variant == 5) {
    if (src.kind == SRC_HASH_SYM)
      return src.gen == 2 ? launch_q1_ws5_cfg<TN, SRC_HASH_SYM, 2>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt)
                          : launch_q1_ws5_cfg<TN, SRC_HASH_SYM, 1>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
    return src.gen == 2 ? launch_q1_ws5_cfg<TN, SRC_HASH_RECT, 2>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt)
                        : launch_q1_ws5_cfg<TN, SRC_HASH_RECT, 1>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
  }
  if (src.kind == SRC_HASH_SYM)
    return src.gen == 2 ? launch_q1_ws_cfg<TN, SRC_HASH_SYM, 2>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt)
                        : launch_q1_ws_cfg<TN, SRC_HASH_SYM, 1>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
  return src.gen == 2 ? launch_q1_ws_cfg<TN, SRC_HASH_RECT, 2>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt)
                      : launch_q1_ws_cfg<TN, SRC_HASH_RECT, 1>(h, src, slab0, bc, nc, Cf, ldc, nfb, T1t, ldt);
}

// fused generation + first quarter; window columns in groups of at most 64
// Cf: coefficient window [nfb][ldc]; Cfs: the same window of the 2^-53-scaled copy (needed by the warp-specialised variant)
int launch_q1_gen(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, const double *Cfs, int64_t ldc,
                  int nfb, double *T1t, int64_t ldt) {
  const bool ws = (h->q1_variant >= 3 && h->q1_variant <= 5) && Cfs && tensor_map_encoder() != nullptr && ((uintptr_t)Cfs % 16 == 0) && (ldc % 2 == 0);
  // Column groups of at most 64 (8 DMMA n-tiles), balanced in units of 8 columns: 80 -> 40 + 40, 150 -> 56 + 48 + 46.
  // (A 64 + 16 split runs its narrow launch at 12 TF/s: measured 19.9 TF/s for 80 columns against 24.3 for 40 + 40.)
  const int groups = (int)ceil_div(nfb, 64), units = (int)ceil_div(nfb, 8);
  for (int gi = 0, f = 0, w = 0; gi < groups && f < nfb; ++gi, f += w) {
    w = std::min(8 * (units / groups + (gi < units % groups ? 1 : 0)), nfb - f);
    const int tn = (int)ceil_div(w, 8);
    const double *cf = (ws ? Cfs : Cf) + (int64_t)f * ldc;
    double *out = T1t + (int64_t)f * bc * ldt;
    cudaError_t e = cudaSuccess;
#define LOWDIN_Q1_CFG(TN) (ws ? launch_q1_ws<TN>(h, src, slab0, bc, nc, cf, ldc, w, out, ldt) : launch_q1_gen_cfg<TN>(h, src, slab0, bc, nc, cf, ldc, w, out, ldt))
    switch (tn) {
      case 1: e = LOWDIN_Q1_CFG(1); break;
      case 2: e = LOWDIN_Q1_CFG(2); break;
      case 3: e = LOWDIN_Q1_CFG(3); break;
      case 4: e = LOWDIN_Q1_CFG(4); break;
      case 5: e = LOWDIN_Q1_CFG(5); break;
      case 6: e = LOWDIN_Q1_CFG(6); break;
      case 7: e = LOWDIN_Q1_CFG(7); break;
      default: e = LOWDIN_Q1_CFG(8); break;
    }
#undef LOWDIN_Q1_CFG
    CK(e);
  }
  return 0;
}

// Fused first quarter of STORED tensors (q1_load_ws5_kernel): the packed rows feed the DMMA warps directly, no dense slab.
template <int TN>
cudaError_t launch_q1_load_cfg(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc,
                               int nfb, double *T1t, int64_t ldt) {
  constexpr int ST = 5;
  constexpr size_t smem = q1_ws5_smem_bytes<TN, ST>();
  auto kern = q1_load_ws5_kernel<TN, ST>;
  static uint64_t configured = 0;
  if (cudaError_t e = ensure_dyn_smem(kern, smem, h->device, configured); e != cudaSuccess) return e;
  CUtensorMap mapB;
  if (!make_operand_map(&mapB, Cf, nfb, nc, ldc, TN * 8)) return cudaErrorInvalidValue;
  const int64_t ntiles = ceil_div(nc, 256) * (int64_t)bc;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, h->num_sms);
  Q1LoadArgs q{src.data, src.M, src.ld, slab0, bc, nc, nfb, T1t, ldt};
  kern<<<grid, 512, smem, h->stream>>>(mapB, q);
  h->launches += 1;
  return cudaGetLastError();
}
bool q1_load_eligible(lowdin_it_handle h, const AoSource &src, const double *Cf, int64_t ldc) {
  return h->stored_fused && (src.kind == SRC_SYM_PACKED || src.kind == SRC_RECT) && tensor_map_encoder() != nullptr &&
         ((uintptr_t)Cf % 16 == 0) && (ldc % 2 == 0);
}
int launch_q1_load(lowdin_it_handle h, const AoSource &src, int64_t slab0, int bc, int nc, const double *Cf, int64_t ldc, int nfb, double *T1t,
                   int64_t ldt) {
  if (src.kind != SRC_RECT) return fail(h, "internal: the fused stored first quarter reads rectangular rows");
  // window columns in balanced groups of at most 64, as for the generated sources (launch_q1_gen)
  const int groups = (int)ceil_div(nfb, 64), units = (int)ceil_div(nfb, 8);
  for (int gi = 0, f = 0, w = 0; gi < groups && f < nfb; ++gi, f += w) {
    w = std::min(8 * (units / groups + (gi < units % groups ? 1 : 0)), nfb - f);
    const int tn = (int)ceil_div(w, 8);
    const double *cf = Cf + (int64_t)f * ldc;
    double *out = T1t + (int64_t)f * bc * ldt;
    cudaError_t e = cudaSuccess;
#define LOWDIN_Q1L(TN) launch_q1_load_cfg<TN>(h, src, slab0, bc, nc, cf, ldc, w, out, ldt)
    switch (tn) {
      case 1: e = LOWDIN_Q1L(1); break;
      case 2: e = LOWDIN_Q1L(2); break;
      case 3: e = LOWDIN_Q1L(3); break;
      case 4: e = LOWDIN_Q1L(4); break;
      case 5: e = LOWDIN_Q1L(5); break;
      case 6: e = LOWDIN_Q1L(6); break;
      case 7: e = LOWDIN_Q1L(7); break;
      default: e = LOWDIN_Q1L(8); break;
    }
#undef LOWDIN_Q1L
    CK(e);
  }
  return 0;
}

// List-driven first quarter of ONE PASS (it_list.cuh): T1list[slab][nu][f] for every slab, from the resident canonical list.
int list_first_quarter(lowdin_it_handle h, const Plan &pl, const PassTables &pt) {
  const Half &hf = pl.h1;
  AoSet &S = h->ao[pl.a][pl.b];
  const int nc = hf.nc, nfb = pt.nfb, nfbp = (nfb + 3) & ~3;
  const int n_slab = h->sp[pl.b].n;
  const int64_t nloc = slabs_owned_below(pl.nslabs1, h->slab_logB, h->nranks, h->rank);  // this rank's slabs (all of them on one GPU)
  const size_t t1_bytes = std::max<size_t>((size_t)nloc * nc * nfbp, 1) * sizeof(double);
  CK(h->Cw.ensure((size_t)nc * nfbp * sizeof(double)));
  CK(h->T1list.ensure(t1_bytes));
  ProfScope ps(h, 1, 2.0 * (double)pl.nslabs1 * nc * (double)nc * nfb);
  list_window_kernel<<<(unsigned)ceil_div((int64_t)nc * nfbp, 256), 256, 0, h->stream>>>(hf.C, hf.ldc, nc, hf.lf - 1 + pt.f0, nfb, nfbp, h->Cw.as<double>());
  CK(cudaMemsetAsync(h->T1list.p, 0, t1_bytes, h->stream));
  for (size_t i = 0; i < S.segs.size(); ++i) {
    list_first_quarter_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(S.segs[i].as<ListEntry>(), S.seg_counts.as<unsigned long long>() + i,
                                                                   pl.intra ? 1 : 0, S.swapped, nc, n_slab, h->Cw.as<double>(), nfb, nfbp,
                                                                   h->T1list.as<double>(), h->slab_logB, h->nranks, h->rank);
    h->launches += 1;
  }
  CK(cudaGetLastError());
  return 0;
}

// First quarter (E.f90:1047-1090) of the batch of bc slabs starting at `slab`: T1t[f][z][mu] = sum_nu AO(slab+z; mu nu) C(nu, lf+f0+f)
int first_quarter_batch(lowdin_it_handle h, const Plan &pl, const PassTables &pt, int64_t slab, int64_t bc, int64_t ldx, int64_t ldt) {
  const Half &hf = pl.h1;
  const int nc = hf.nc, nfb = pt.nfb;
  const double *Cf = hf.C + (int64_t)(hf.lf - 1 + pt.f0) * hf.ldc;
  const double *Cfs = hf.Cs ? hf.Cs + (int64_t)(hf.lf - 1 + pt.f0) * hf.ldc : nullptr;
  if (pl.src.kind == SRC_LIST) {
    // the quarter-transformed slabs of the pass exist already (list_first_quarter): bring the batch into the operand layout
    const int nfbp = (nfb + 3) & ~3;
    dim3 grid((unsigned)ceil_div(nc, 32), (unsigned)ceil_div(nfb, 32), (unsigned)bc);
    ProfScope ps(h, 0, (double)bc * nc * nfb * 16.0);
    list_transpose_kernel<<<grid, 256, 0, h->stream>>>(h->T1list.as<double>(), slab, (int)bc, nc, nfb, nfbp, h->T1t.as<double>(), ldt);
    h->launches += 1;
    CK(cudaGetLastError());
  } else if (pl.src.kind == SRC_HASH_SYM || pl.src.kind == SRC_HASH_RECT) {
    // slab generation + first quarter in one kernel: the dense slab never exists
    ProfScope ps(h, 1, 2.0 * bc * nc * (double)nc * nfb);
    if (launch_q1_gen(h, pl.src, slab, (int)bc, nc, Cf, Cfs, hf.ldc, nfb, h->T1t.as<double>(), ldt)) return 1;
  } else if (q1_load_eligible(h, pl.src, Cf, hf.ldc)) {
    // stored tensor, fused: the M-vector of a slab feeds the DMMA warps directly, the symmetric N x N slab of E.f90:1047-1063 is
    // formed on the way into shared memory (q1_load_ws5_kernel) and never exists in HBM
    AoSource rows = pl.src;
    int64_t first = slab;
    if (pl.src.kind == SRC_SYM_PACKED) {
      // packed tensor: first the full M-vectors of the batch (both halves of the packed symmetry in whole sectors)
      const int64_t M = pl.src.M, ldr = roundup2(M);
      ProfScope ps(h, 0, (double)bc * 16.0 * (double)M);
      dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(bc, 32));
      complete_rows_kernel<<<grid, 256, 0, h->stream>>>(pl.src.data, M, slab, (int)bc, ldr, h->X.as<double>());
      h->launches += 1;
      CK(cudaGetLastError());
      rows = AoSource{SRC_RECT, h->X.as<double>(), M, ldr, bc, 0};
      first = 0;
    }
    ProfScope ps(h, 1, 2.0 * bc * nc * (double)nc * nfb);
    if (launch_q1_load(h, rows, first, (int)bc, nc, Cf, hf.ldc, nfb, h->T1t.as<double>(), ldt)) return 1;
  } else {
    {  // unpack (E.f90:1047-1063)
      ProfScope ps(h, 0, (double)bc * 8.0 * ((double)pl.src.M + (double)nc * nc));
      if (launch_expand(h, pl.src, slab, bc, nc, 0, nc, 0, nc, 0, (int)ldx, h->X.as<double>())) return 1;
    }
    // first quarter (E.f90:1081-1090): T1t[f][z][mu] = sum_nu X[z][mu][nu] C(nu, lf+f0+f)
    GemmArgs g{h->X.as<double>(), Cf, (int)(bc * nc), nfb, nc, ldx, hf.ldc, 0, 0};
    EpiQ1 epi{h->T1t.as<double>(), nc, (int)bc, ldt};
    ProfScope ps(h, 1, 2.0 * bc * nc * (double)nc * nfb);
    CK(launch_gemm(h, g, epi));
  }
  return 0;
}

// doubles of the X workspace one slab of the first half needs: the dense slab (expansion kernel + GEMM), the completed M-vector
// (packed tensor, fused first quarter), nothing (generated / list / rectangular rows read in place)
size_t x_per_slab(lowdin_it_handle h, const Plan &pl) {
  const int nc = pl.h1.nc;
  const bool stored = (pl.src.kind == SRC_SYM_PACKED || pl.src.kind == SRC_RECT);
  if (stored && q1_load_eligible(h, pl.src, pl.h1.C, pl.h1.ldc)) return pl.src.kind == SRC_SYM_PACKED ? (size_t)roundup2(pl.src.M) : 0;
  if (stored || pl.src.kind == SRC_RANKK) return (size_t)nc * roundup2(nc);
  return 0;
}
// slabs per batch of the first half: what the X / T1t workspaces hold, within the grid limits of the kernels
int64_t first_half_batch(lowdin_it_handle h, const Plan &pl, int nfb, int64_t count) {
  const int nc = pl.h1.nc;
  const int64_t ldt = roundup2(nc);
  const size_t per_slab = std::max(x_per_slab(h, pl), (size_t)nfb * ldt) * sizeof(double);
  int64_t B = std::max<int64_t>(1, (int64_t)(h->workspace_bytes / per_slab));
  B = std::min<int64_t>(B, count);
  B = std::min<int64_t>(B, std::max<int64_t>(1, (int64_t)(60000LL * 128 / nc)));  // gridDim.y limit of the stacked GEMM
  B = std::min<int64_t>(B, 65535);                                                 // gridDim limits of expansion / fused kernel
  return B;
}

// First half (E.f90:1043-1132) of `count` AO-pair slabs starting at slab `slab0`:
//   Hc[slot][col0 + (slab - slab0)] for every slot of the pass.
int first_half(lowdin_it_handle h, const Plan &pl, const PassTables &pt, int64_t slab0, int64_t count, double *Hc, int64_t ldh,
               int64_t col0, double tol) {
  const Half &hf = pl.h1;
  const int nc = hf.nc, nfb = pt.nfb;
  const int64_t ldx = roundup2(nc), ldt = roundup2(nc);
  const int64_t B = first_half_batch(h, pl, nfb, count);
  if (x_per_slab(h, pl)) CK(h->X.ensure((size_t)B * x_per_slab(h, pl) * sizeof(double)));
  CK(h->T1t.ensure((size_t)B * nfb * ldt * sizeof(double)));
  for (int64_t s = 0; s < count; s += B) {
    const int64_t bc = std::min<int64_t>(B, count - s);
    if (first_quarter_batch(h, pl, pt, slab0 + s, bc, ldx, ldt)) return 1;
    {  // second quarter (E.f90:1099-1110): T2[s][(f,z)] = sum_mu C(mu, ls+s) T1t[f][z][mu]  ->  Hc[slot(s,f)][col0+s'+z]
      GemmArgs g{hf.C + (int64_t)(hf.ls - 1) * hf.ldc, h->T1t.as<double>(), hf.ns, (int)(bc * nfb), nc, hf.ldc, ldt, 0, 0};
      EpiScatterH epi{Hc, ldh, col0 + s, h->tab.as<int32_t>(), nfb, (int)bc, tol};
      ProfScope ps(h, 2, 2.0 * bc * nc * (double)hf.ns * nfb);
      CK(launch_gemm(h, g, epi));
    }
  }
  return 0;
}

// A chunk of AO-pair slabs = the pairs (p,q>=p) of rows p in [p0,p1) of the second-half species' pair triangle.
struct Chunk { int p0, p1; int64_t base, width; };

// Partial second half (E.f90:1178-1239, third quarter only) of one chunk for slots [s_lo, s_hi):
//   T3[slot][kf][mu] += sum_nu Y_chunk(mu,nu) C(nu, kf),  Y_chunk = the symmetric N x N matrix of the slot's
//   half-transformed row restricted to pairs whose smaller index lies in [p0,p1).
// Two dense blocks carry it: W = Y[p0.., p0..p1) (all rows below, chunk columns) and V = Y[p0..p1), p1..) .
int second_half_partial(lowdin_it_handle h, const Plan &pl, const AoSource &hsrc, const Chunk &ck, int s_lo, int s_hi, double *T3,
                        int64_t ldt2) {
  const Half &h2 = pl.h2;
  const int n2 = h2.nc, nf2 = h2.nf;
  const int pc = ck.p1 - ck.p0, nrw = n2 - ck.p0, ncv = n2 - ck.p1;
  const int64_t ldw = roundup2(pc), ldv = roundup2(std::max(ncv, 1));
  const size_t per_slot = std::max((size_t)nrw * ldw, (size_t)pc * ldv) * sizeof(double);
  int64_t B = std::max<int64_t>(1, (int64_t)(h->workspace_bytes / per_slot));
  B = std::min<int64_t>(B, s_hi - s_lo);
  B = std::min<int64_t>(B, std::max<int64_t>(1, (int64_t)(60000LL * 128 / std::max(nrw, 1))));
  B = std::min<int64_t>(B, 65535);
  CK(h->X.ensure((size_t)B * nrw * ldw * sizeof(double)));
  if (ncv > 0) CK(h->T1t.ensure((size_t)B * pc * ldv * sizeof(double)));
  const double *C2f = h2.C + (int64_t)(h2.lf - 1) * h2.ldc;
  for (int64_t s = s_lo; s < s_hi; s += B) {
    const int64_t bs = std::min<int64_t>(B, s_hi - s);
    double *T3s = T3 + (int64_t)(s - s_lo) * nf2 * ldt2;
    {
      ProfScope ps(h, 3, (double)bs * 8.0 * ((double)ck.width + (double)nrw * pc + (double)pc * ncv));
      if (launch_expand(h, hsrc, s, bs, n2, ck.p0, nrw, ck.p0, pc, ck.base, (int)ldw, h->X.as<double>())) return 1;
      if (ncv > 0 && launch_expand(h, hsrc, s, bs, n2, ck.p0, pc, ck.p1, ncv, ck.base, (int)ldv, h->T1t.as<double>())) return 1;
    }
    {
      ProfScope ps(h, 4, 2.0 * bs * (double)nf2 * ((double)nrw * pc + (double)pc * ncv));
      GemmArgs g{h->X.as<double>(), C2f + ck.p0, (int)(bs * nrw), nf2, pc, ldw, h2.ldc, 0, 0};
      if (h->q3_red) CK(launch_gemm(h, g, EpiAccRed{T3s, nrw, ck.p0, nf2, ldt2}));
      else CK(launch_gemm(h, g, EpiAccT{T3s, nrw, ck.p0, nf2, ldt2}));
      if (ncv > 0) {
        GemmArgs g2{h->T1t.as<double>(), C2f + ck.p1, (int)(bs * pc), nf2, ncv, ldv, h2.ldc, 0, 0};
        if (h->q3_red) CK(launch_gemm(h, g2, EpiAccRed{T3s, pc, ck.p0, nf2, ldt2}));
        else CK(launch_gemm(h, g2, EpiAccT{T3s, pc, ck.p0, nf2, ldt2}));
      }
    }
  }
  return 0;
}

// How one chunk of AO-pair slabs [base, base + width) and the pass's slots are divided among G ranks:
// slabs block-cyclically (blocks of 2^logB slabs, block b on rank b % G -- it_kernels.cuh, slab_owner): rank r computes its
// own slabs of the chunk, which are the consecutive LOCAL slabs [loc_lo, loc_lo + cnt); wblk = the largest cnt over the ranks
// (row stride of every rank's H and of the exchanged blocks).  Slots by contiguous blocks of first-contracted indices.
void shard_plan(int nfb, const int *fbeg, int64_t base, int64_t width, int G, int rank, int logB, int *own, int64_t *wblk, int64_t *loc_lo,
                int64_t *cnt) {
  for (int r = 0; r <= G; ++r) own[r] = fbeg[(int)((int64_t)nfb * r / G)];
  int64_t w = 0;
  for (int r = 0; r < G; ++r) {
    const int64_t lo = slabs_owned_below(base, logB, G, r), c = slabs_owned_below(base + width, logB, G, r) - lo;
    w = std::max(w, c);
    if (r == rank) { *loc_lo = lo; *cnt = c; }
  }
  *wblk = w;
}

// Ranks size their batches from their OWN free memory, which differs by a few MB from rank to rank; everything that
// decides the shape of the exchange (occupied batch, chunk boundaries) must be the same number everywhere, so the ranks
// take the minimum (one 8-byte ncclAllReduce on the library's stream).  Single GPU: nothing happens.
int agree_min(lowdin_it_handle h, int64_t *v) {
  if (h->nranks <= 1) return 0;
  if (h->lgroup) {
    LocalGroup &L = *h->lgroup;
    L.agree[h->rank] = *v;
    if (!L.barrier()) return fail(h, "in-process group: a rank did not reach the collective");
    int64_t m = L.agree[0];
    for (int g = 1; g < L.n; ++g) m = std::min(m, L.agree[g]);
    if (!L.barrier()) return fail(h, "in-process group: a rank did not reach the collective");  // everyone has read before anyone writes the next value
    *v = m;
    return 0;
  }
  if (!h->comm) return fail(h, "multi-GPU transform without a communicator");
  CK(h->agree.ensure(sizeof(int64_t)));
  CK(cudaMemcpyAsync(h->agree.p, v, sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  const int rc = g_nccl.AllReduce(h->agree.p, h->agree.p, 1, /*ncclInt64*/ 4, /*ncclMin*/ 3, h->comm, h->stream);
  if (rc != 0) return fail(h, std::string("ncclAllReduce failed: ") + g_nccl.GetErrorString(rc));
  CK(cudaMemcpyAsync(v, h->agree.p, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

struct Consumer {
  int mode = 0;  // 0: compaction (download), 1: streaming reduce (+ optional host sink of every dense result block)
  lowdin_it_sink_fn sink = nullptr;
  void *sink_user = nullptr;
  double tol = 1e-10;
  const double *epsA = nullptr, *epsB = nullptr;  // device
  double lambda = 2.0;
};

int exchange_chunk(lowdin_it_handle h, const std::vector<int> &own, int64_t wblk, DevBuf &Hsrc, DevBuf &H2dst, cudaStream_t st);

int run_passes(lowdin_it_handle h, const Plan &pl, int occ_batch, int first_pass, int n_passes, const Consumer &cons,
               double sums_out[4]) {
  const Half &h1 = pl.h1, &h2 = pl.h2;
  const int nf_total = h1.nf;
  const int G = h->nranks;
  if (occ_batch <= 0 || occ_batch > nf_total) occ_batch = std::max(nf_total, 1);
  const int total_passes = (int)ceil_div(std::max(nf_total, 1), occ_batch);
  if (n_passes <= 0) { first_pass = 0; n_passes = total_passes; }
  if (first_pass < 0 || first_pass + n_passes > total_passes) return fail(h, "pass range out of bounds");
  for (int t = 0; t < 8; ++t) h->timers[t] = (t == 0 || t == 5) ? h->timers[t] : 0.0;
  h->launches = 0;
  if (cons.mode == 1) {
    CK(h->sums.ensure(4 * sizeof(double)));
    CK(cudaMemsetAsync(h->sums.p, 0, 4 * sizeof(double), h->stream));
  }
  h->count = 0;
  h->sink_bytes = 0;
  if (cons.mode == 0) {
    // An empty window (e.g. a one-function species under MP2: virtual window 2..1) is a valid transform with no
    // integrals: the reference writes a terminator-only moint.dat (C.f90:450-456, E.f90:1263-1268).
    h->res_conv = pl.conv;
    CK(h->overflow.ensure(sizeof(int)));
    CK(cudaMemsetAsync(h->overflow.p, 0, sizeof(int), h->stream));
  }
  if (pl.pairs_s.empty() || h2.nf <= 0 || h2.ns <= 0) {  // an empty window: no integrals, nothing to run
    if (cons.mode == 1 && sums_out) sums_out[0] = sums_out[1] = sums_out[2] = sums_out[3] = 0.0;
    h->timers[6] = 0; h->timers[7] = 0;
    return 0;
  }
  double flops = 0.0;
  const int n2 = h2.nc, nf2 = h2.nf, ns2 = h2.ns;
  const int64_t ldt2 = roundup2(n2);
  const int64_t per_out = (int64_t)ns2 * nf2;
  const double half_tol = (pl.conv == LOWDIN_IT_CONV_E) ? cons.tol : -1.0;

  for (int pass = first_pass; pass < first_pass + n_passes; ++pass) {
    PassTables pt;
    const int f0 = pass * occ_batch, nfb = std::min(occ_batch, nf_total - f0);
    if (nfb <= 0) continue;
    build_pass(pl, f0, nfb, pt);
    if (pt.nslots == 0) continue;
    if (upload_i32(h, h->tab, pt.table) || upload_i32(h, h->sa, pt.sa) || upload_i32(h, h->sb, pt.sb) ||
        upload_i32(h, h->ss, pt.ss) || upload_i32(h, h->sf, pt.sf) || upload_i32(h, h->order, pt.order))
      return 1;
    // slot ownership: rank r owns the slots of a contiguous block of first-contracted indices
    std::vector<int> own(G + 1, 0);
    { int64_t w_, a_, b_; shard_plan(nfb, pt.fbeg.data(), 0, 0, G, h->rank, h->slab_logB, own.data(), &w_, &a_, &b_); }
    const int s_lo = own[h->rank], s_hi = own[h->rank + 1], nmine = s_hi - s_lo;
    // ---- device memory of the pass: T3 accumulators (own slots) + one chunk of half-transformed rows ----
    const size_t t3_bytes = std::max<size_t>((size_t)std::max(nmine, 1) * nf2 * ldt2, 1) * sizeof(double);
    if (t3_bytes > h->T3.cap) {
      // the chunk buffers of an earlier transform may hold all the free memory (they take what is left): give them back before growing T3
      size_t fb = 0, tb = 0;
      CK(cudaMemGetInfo(&fb, &tb));
      if (fb < t3_bytes + ((size_t)1 << 30)) {
        CK(cudaStreamSynchronize(h->stream));
        for (DevBuf *b : {&h->H, &h->H2, &h->Hx, &h->H2x, &h->OUT, &h->T3}) b->release();
      }
    }
    CK(h->T3.ensure(t3_bytes));
    CK(cudaMemsetAsync(h->T3.p, 0, t3_bytes, h->stream));
    if (pl.src.kind == SRC_LIST) {
      if (list_first_quarter(h, pl, pt)) return 1;  // flops are counted with the chunks below (the algorithmic count is the dense one)
    }
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const bool download = (cons.mode == 0);
    double out_need = download ? (double)std::max(nmine, 1) * per_out * 8.0
                               : std::min(4.0e9, (double)std::max(nmine, 1) * per_out * 8.0);
    out_need = std::max(out_need, (double)pl.max_slots_per_f * per_out * 8.0);
    const bool sink_mode = (cons.mode == 1 && cons.sink != nullptr);
    // a host sink double-buffers its (smaller) groups: two groups of max(workspace, one f-block) must fit
    const double out_reserve = sink_mode ? 2.0 * std::min(out_need, std::max((double)h->sink_group_bytes, (double)pl.max_slots_per_f * per_out * 8.0)) : out_need;
    {
      // OUT is sized HERE, before the chunk buffers take what is left: a pass with a host sink needs two groups where the same pass
      // without one needed a single group, and the chunk buffers of that earlier pass may hold all the free memory (N_bf = 2000: one
      // f-block of results is 5.2 GB)
      const double gb = sink_mode ? std::min(out_need, std::max((double)h->sink_group_bytes, (double)pl.max_slots_per_f * per_out * 8.0)) : out_need;
      const int64_t cap_slots = download ? std::max(nmine, 1) : std::max<int64_t>(pl.max_slots_per_f, (int64_t)(gb / (per_out * 8.0)));
      const size_t need = std::max<size_t>((size_t)cap_slots * per_out, 1) * (sink_mode ? 2 : 1) * sizeof(double);
      if (need > h->OUT.cap) {
        if (free_b + h->OUT.cap < need + ((size_t)1 << 29)) {
          CK(cudaStreamSynchronize(h->stream));
          if (h->comm_stream) CK(cudaStreamSynchronize(h->comm_stream));
          for (DevBuf *b : {&h->H, &h->H2, &h->Hx, &h->H2x, &h->X, &h->T1t}) b->release();
        }
        CK(h->OUT.ensure(need));
        CK(cudaMemGetInfo(&free_b, &total_b));
      }
    }
    double avail = 0.92 * ((double)free_b + (double)h->H.cap + (double)h->H2.cap + (double)h->Hx.cap + (double)h->H2x.cap + (double)h->OUT.cap +
                           (double)h->X.cap + (double)h->T1t.cap) -
                   std::max(out_need, out_reserve) - 2.0 * (double)h->workspace_bytes - (double)((size_t)1 << 29);
    // bytes per chunk column: one rank holds H[all slots][its 1/G of the columns] and, after the exchange, H2[its slots][all
    // columns] (counted twice: headroom for NCCL's own buffers) -- the same 3/G of a full column pick_occ_batch assumes
    const double per_col = (G > 1) ? (double)pt.nslots * 8.0 / G + 2.0 * (double)nmine * 8.0 + 8.0 : (double)pt.nslots * 8.0;
    int64_t max_cols = (int64_t)std::max(avail / per_col, 0.0);
    if (agree_min(h, &max_cols)) return 1;  // same chunk boundaries on every rank (the exchange depends on them)
    // Overlapping the exchange takes two sets of both buffers, i.e. chunks 3/4 as wide -- more chunks, more read-modify-write passes
    // over T3.  Measured at N_bf = 1500 (profiles/r02n_*): with ~24 chunks per pass (2 GPUs) the narrower chunks cost more than the
    // hidden exchange gains (47.3 against 50.1 TF/s); with memory to spare (<= 12 chunks per pass) the exchange is what is left to
    // hide.  1 = decide per pass (from the agreed width, so every rank decides alike), 2 = always, 0 = never.
    bool use_overlap = false;
    if (G > 1 && h->overlap_exchange) {
      use_overlap = (h->overlap_exchange == 2) || ((double)max_cols * 12.0 >= (double)pl.nslabs1);
      if (use_overlap) max_cols = max_cols * 3 / 4;
    }
    if (max_cols < 2 * (int64_t)n2 && max_cols < pl.nslabs1) return fail(h, "not enough device memory for one chunk of half-transformed integrals; lower occ_batch");
    if (h->chunk_cols_limit > 0) max_cols = std::min<int64_t>(max_cols, h->chunk_cols_limit);
    // chunks: rows [p0,p1) of the pair triangle, even boundaries, as many rows as fit
    std::vector<Chunk> chunks;
    for (int p0 = 0; p0 < n2;) {
      int p1 = p0;
      int64_t width = 0;
      while (p1 < n2) {
        int step = std::min(2, n2 - p1);
        int64_t add = 0;
        for (int t = 0; t < step; ++t) add += n2 - (p1 + t);
        if (p1 > p0 && width + add > max_cols) break;
        width += add; p1 += step;
      }
      chunks.push_back({p0, p1, (int64_t)pair0(p0, p0, n2), width});
      p0 = p1;
    }
    float ms_first = 0, ms_exch = 0, ms_second = 0;
    // ---------------- the chunks: first half -> exchange -> third quarter ----------------
    // N > 1 ranks: two sets of chunk buffers; the exchange of chunk c runs on its own stream while the first half of chunk c + 1
    // computes (compute stream: FH(0) FH(1) SH(0) FH(2) SH(1) ...).
    const bool pipelined = use_overlap && chunks.size() > 1;
    if (G > 1 && !h->comm_stream) {
      CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_fh[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_ex[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_sh[i], cudaEventDisableTiming));
      }
    }
    cudaStream_t cs = pipelined ? h->comm_stream : h->stream;
    DevBuf *Hb[2] = {&h->H, &h->Hx}, *H2b[2] = {&h->H2, &h->H2x}, *tabb[2] = {&h->coltab, &h->coltabx};
    while (h->chunk_ev.size() < 4 * chunks.size() + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->chunk_ev.push_back(e); }
    std::vector<int64_t> c_ldh(chunks.size(), 0);
    auto stage_first = [&](size_t c) -> int {
      const Chunk &ck = chunks[c];
      const int bsel = pipelined ? (int)(c & 1) : 0;
      int64_t wblk, loc_lo, cnt;  // this rank's slabs of the chunk: local slabs [loc_lo, loc_lo + cnt) (G == 1: the chunk itself)
      shard_plan(nfb, pt.fbeg.data(), ck.base, ck.width, G, h->rank, h->slab_logB, own.data(), &wblk, &loc_lo, &cnt);
      const int64_t ldh = (G > 1) ? std::max<int64_t>(wblk, 1) : ck.width;
      c_ldh[c] = ldh;
      CK(Hb[bsel]->ensure(std::max<size_t>((size_t)pt.nslots * ldh, 1) * sizeof(double)));
      CK(cudaEventRecord(h->chunk_ev[4 * c], h->stream));
      // first half (E.f90:1043-1132) over this rank's share of the chunk's slabs
      if (cnt > 0 && first_half(h, pl, pt, loc_lo, cnt, Hb[bsel]->as<double>(), ldh, 0, half_tol)) return 1;
      flops += 2.0 * h1.nc * nfb * ((double)h1.nc + h1.ns) * (double)cnt;
      CK(cudaEventRecord(h->chunk_ev[4 * c + 1], h->stream));
      if (G > 1) {
        // exchange (the it2.tmp bucket file of E.f90:1121-1141, :1189-1203)
        if (pipelined) {
          CK(cudaEventRecord(h->ev_fh[bsel], h->stream));
          CK(cudaStreamWaitEvent(cs, h->ev_fh[bsel], 0));
          if (c >= 2) CK(cudaStreamWaitEvent(cs, h->ev_sh[bsel], 0));  // the third quarter of chunk c - 2 has read this H2
        }
        {
          ProfScope ps(h, 7, (double)pt.nslots * (double)cnt * 8.0, cs);
          if (exchange_chunk(h, own, ldh, *Hb[bsel], *H2b[bsel], cs)) return 1;
        }
        if (pipelined) CK(cudaEventRecord(h->ev_ex[bsel], cs));
      }
      return 0;
    };
    auto stage_second = [&](size_t c) -> int {
      const Chunk &ck = chunks[c];
      const int bsel = pipelined ? (int)(c & 1) : 0;
      const int64_t ldh = c_ldh[c];
      if (pipelined) CK(cudaStreamWaitEvent(h->stream, h->ev_ex[bsel], 0));
      CK(cudaEventRecord(h->chunk_ev[4 * c + 2], h->stream));
      AoSource hsrc{SRC_RECT, Hb[bsel]->as<double>(), ck.width, ldh, 0, 0};
      if (G > 1) {
        // column table: chunk column -> (sending rank's block, its local column)
        RankStarts st{};
        for (int r = 0; r < G; ++r) st.lo[r] = slabs_owned_below(ck.base, h->slab_logB, G, r);
        CK(tabb[bsel]->ensure((size_t)ck.width * sizeof(int64_t)));
        column_table_kernel<<<(unsigned)ceil_div(ck.width, 256), 256, 0, h->stream>>>(ck.base, ck.width, h->slab_logB, G, (int64_t)nmine, ldh, st,
                                                                                     tabb[bsel]->as<int64_t>());
        CK(cudaGetLastError());
        hsrc = AoSource{SRC_RECT_TABLE, H2b[bsel]->as<double>(), ck.width, ldh, (int64_t)nmine, 0, 0, reinterpret_cast<const double *>(tabb[bsel]->p)};
      }
      // third quarter of the chunk, accumulated into T3 (own slots)
      if (nmine > 0) {
        // rows of hsrc are numbered from this rank's first slot when the data came through the exchange
        int lo = s_lo, hi = s_hi;
        if (G > 1) { lo = 0; hi = nmine; }
        if (second_half_partial(h, pl, hsrc, ck, lo, hi, h->T3.as<double>(), ldt2)) return 1;
        const double nrw = n2 - ck.p0, pc = ck.p1 - ck.p0, ncv = n2 - ck.p1;
        flops += 2.0 * (double)nmine * nf2 * (nrw * pc + pc * ncv);
      }
      CK(cudaEventRecord(h->chunk_ev[4 * c + 3], h->stream));
      if (pipelined) CK(cudaEventRecord(h->ev_sh[bsel], h->stream));
      return 0;
    };
    bool dma = (G > 1 && h->node && h->exchange_dma && !h->lgroup);
    if (dma) {
      // the chunk buffers at their largest size of the pass, once (peers map them for the whole chunk loop), then the mappings
      NodeLink &N = *h->node;
      size_t hmax = 1;
      for (const Chunk &ck : chunks) {
        int64_t wblk, lo_, cnt_;
        shard_plan(nfb, pt.fbeg.data(), ck.base, ck.width, G, h->rank, h->slab_logB, own.data(), &wblk, &lo_, &cnt_);
        hmax = std::max(hmax, (size_t)pt.nslots * (size_t)std::max<int64_t>(wblk, 1));
      }
      NodeShm::Rank &me = N.shm->r[h->rank];
      int64_t mapped_ok = 1;
      for (int b = 0; b < 2; ++b) {
        me.has[b] = 0;
        if (b == 1 && !pipelined) continue;
        CK(Hb[b]->ensure(hmax * sizeof(double)));
        if (cudaIpcGetMemHandle(&me.mem[b], Hb[b]->p) == cudaSuccess) me.has[b] = 1; else mapped_ok = 0;
      }
      if (!N.barrier()) return fail(h, "node link: a rank did not reach the pass");
      for (int g = 0; g < G && mapped_ok; ++g) {
        if (g == h->rank) continue;
        for (int b = 0; b < (pipelined ? 2 : 1) && mapped_ok; ++b)
          if (!N.shm->r[g].has[b] || cudaIpcOpenMemHandle(&N.peer_mem[g][b], N.shm->r[g].mem[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) mapped_ok = 0;
      }
      cudaGetLastError();
      if (agree_min(h, &mapped_ok)) return 1;
      if (mapped_ok) h->node_mapped = true;
      else {  // a buffer could not be exported or mapped somewhere: every rank goes back to ncclSend/ncclRecv, for good
        N.unmap_all();
        if (!N.barrier()) return fail(h, "node link: a rank did not reach the pass");
        h->exchange_dma = 0;
        dma = false;
      }
    }
    cudaEvent_t ev_begin = h->chunk_ev[4 * chunks.size()], ev_end = h->chunk_ev[4 * chunks.size() + 1];
    CK(cudaEventRecord(ev_begin, h->stream));
    if (pipelined) {
      if (stage_first(0)) return 1;
      for (size_t c = 0; c < chunks.size(); ++c) {
        if (c + 1 < chunks.size() && stage_first(c + 1)) return 1;
        if (stage_second(c)) return 1;
      }
    } else {
      for (size_t c = 0; c < chunks.size(); ++c)
        if (stage_first(c) || stage_second(c)) return 1;
    }
    CK(cudaEventRecord(ev_end, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (G > 1) CK(cudaStreamSynchronize(cs));
    if (dma) {  // nobody keeps a peer's buffer mapped outside the chunk loop: buffers may be freed or grown between passes
      h->node->unmap_all();
      h->node_mapped = false;
      if (!h->node->barrier()) return fail(h, "node link: a rank did not finish the pass");
    }
    prof_drain(h);
    {
      float ms, total = 0;
      for (size_t c = 0; c < chunks.size(); ++c) {
        CK(cudaEventElapsedTime(&ms, h->chunk_ev[4 * c], h->chunk_ev[4 * c + 1])); ms_first += ms;
        CK(cudaEventElapsedTime(&ms, h->chunk_ev[4 * c + 2], h->chunk_ev[4 * c + 3])); ms_second += ms;
      }
      CK(cudaEventElapsedTime(&total, ev_begin, ev_end));
      ms_exch = std::max(0.0f, total - ms_first - ms_second);  // what the compute stream spent waiting for an exchange
    }
    // ---------------- fourth quarter (E.f90:1230-1239) + consumer, in groups of whole f-blocks ----------------
    CK(cudaEventRecord(h->ev[1], h->stream));
    float ms_consume = 0;
    if (nmine > 0) {
      const int fr_lo = (int)((int64_t)nfb * h->rank / G), fr_hi = (int)((int64_t)nfb * (h->rank + 1) / G);
      const bool sinking = (cons.mode == 1 && cons.sink != nullptr);
      // with a host sink the groups are smaller (256 MiB by default: finer pipelining of Q4, device-to-host copy and the host's consumer) and
      // OUT is double-buffered: the copy of group g overlaps the fourth quarter of group g + 1
      const double group_bytes = sinking ? std::min(out_need, std::max((double)h->sink_group_bytes, (double)pl.max_slots_per_f * per_out * 8.0)) : out_need;
      const int64_t out_slots_cap = download ? nmine : std::max<int64_t>(pl.max_slots_per_f, (int64_t)(group_bytes / (per_out * 8.0)));
      const size_t out_elems = std::max<size_t>((size_t)out_slots_cap * per_out, 1);
      CK(h->OUT.ensure(out_elems * (sinking ? 2 : 1) * sizeof(double)));
      int gi = 0, pending = -1;            // sink: group counter, ring slot whose block the host has not consumed yet
      lowdin_it_block blk_pending{};
      if (sinking) {
        if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
          if (!h->ev_q4[i]) CK(cudaEventCreateWithFlags(&h->ev_q4[i], cudaEventDisableTiming));
          if (!h->ev_d2h[i]) CK(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
        }
        if (h->sink_host_cap < out_elems * sizeof(double)) {
          for (int i = 0; i < 2; ++i) { if (h->sink_host[i]) cudaFreeHost(h->sink_host[i]); h->sink_host[i] = nullptr; }
          h->sink_host_cap = 0;
          for (int i = 0; i < 2; ++i) CK(cudaHostAlloc(&h->sink_host[i], out_elems * sizeof(double), cudaHostAllocDefault));
          h->sink_host_cap = out_elems * sizeof(double);
        }
      }
      auto flush_pending = [&]() -> int {   // hand the block whose copy was enqueued last to the host consumer
        if (pending < 0) return 0;
        CK(cudaEventSynchronize(h->ev_d2h[pending]));
        pending = -1;
        if (cons.sink(cons.sink_user, &blk_pending)) return fail(h, "the sink callback returned an error");
        return 0;
      };
      for (int fa = fr_lo; fa < fr_hi;) {
        int fb = fa + 1;
        while (fb < fr_hi && pt.fbeg[fb + 1] - pt.fbeg[fa] <= out_slots_cap) ++fb;
        const int g_lo = pt.fbeg[fa], g_hi = pt.fbeg[fb], gn = g_hi - g_lo;
        fa = fb;
        if (gn <= 0) continue;
        const double *T3g = h->T3.as<double>() + (int64_t)(g_lo - s_lo) * nf2 * ldt2;
        const int ring = sinking ? (gi & 1) : 0;
        double *OUT = h->OUT.as<double>() + (size_t)ring * out_elems;
        if (sinking && gi >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_d2h[ring], 0));  // the copy of group gi - 2 has left this buffer
        // batches bounded by the GEMM grid (N dimension = slots * nf2)
        const int64_t Bq = std::max<int64_t>(1, std::min<int64_t>(gn, (int64_t)2000000000 / std::max(nf2, 1) / 64));
        for (int64_t s = 0; s < gn; s += Bq) {
          const int64_t bs = std::min<int64_t>(Bq, gn - s);
          GemmArgs g{h2.C + (int64_t)(h2.ls - 1) * h2.ldc, T3g + s * nf2 * ldt2, ns2, (int)(bs * nf2), n2, h2.ldc, ldt2, 0, 0};
          ProfScope ps(h, 5, 2.0 * bs * n2 * (double)ns2 * nf2);
          CK(launch_gemm(h, g, EpiOut{OUT + s * per_out, ns2, nf2}));
        }
        flops += 2.0 * (double)gn * n2 * (double)ns2 * nf2;
        if (!download) {
          ReduceArgs ra{};
          ra.OUT = OUT; ra.nslots = gn; ra.ns2 = ns2; ra.nf2 = nf2; ra.slot_base = g_lo;
          ra.slot_s = h->ss.as<int32_t>() + g_lo; ra.slot_f = h->sf.as<int32_t>() + g_lo;
          ra.slot_table = h->tab.as<int32_t>(); ra.nfb = nfb;
          ra.orb_s1 = h1.ls; ra.orb_f1 = h1.lf + f0; ra.orb_s2 = h2.ls; ra.orb_f2 = h2.lf;
          ra.epsA = cons.epsA; ra.epsB = cons.epsB;
          ra.exchange = (pl.intra && h1.ls == h2.ls && h1.ns == h2.ns && h1.lf == h2.lf && h1.nf == h2.nf &&
                         pl.win[0] == pl.win[4] && pl.win[2] == pl.win[6]) ? 1 : 0;
          ra.lambda = cons.lambda; ra.tol = cons.tol;
          ProfScope ps(h, 6, (double)gn * per_out * 8.0);
          reduce_block_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(ra, h->sums.as<double>());
          h->launches += 1;
          CK(cudaGetLastError());
        }
        if (sinking) {
          // the previous block goes to the host consumer while this group's fourth quarter runs; it also frees its ring slot
          if (flush_pending()) return 1;
          CK(cudaEventRecord(h->ev_q4[ring], h->stream));
          CK(cudaStreamWaitEvent(h->copy_stream, h->ev_q4[ring], 0));
          CK(cudaMemcpyAsync(h->sink_host[ring], OUT, (size_t)gn * per_out * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
          CK(cudaEventRecord(h->ev_d2h[ring], h->copy_stream));
          blk_pending = lowdin_it_block{pl.conv, gn, ns2, nf2, h2.ls, h2.lf, h2.first_is_conv_second ? 1 : 0, pt.sa.data() + g_lo, pt.sb.data() + g_lo,
                                        static_cast<const double *>(h->sink_host[ring])};
          pending = ring;
          h->sink_bytes += (double)gn * per_out * sizeof(double);
          ++gi;
        }
      }
      if (sinking && flush_pending()) return 1;
    }
    CK(cudaEventRecord(h->ev[2], h->stream));
    if (download) { h->res_pair.clear(); h->res_seg.clear(); }
    if (nmine > 0 && download) {
      // this rank's window pairs in the reference's loop order, as slots relative to its first one (one GPU: all of them)
      std::vector<int32_t> own_order;
      for (size_t k = 0; k < pt.order.size(); ++k)
        if (pt.order[k] >= s_lo && pt.order[k] < s_hi) { own_order.push_back(pt.order[k] - s_lo); h->res_pair.push_back(pt.order_conv[k]); }
      if (upload_i32(h, h->order, own_order)) return 1;
      SelectArgs sa{};
      sa.OUT = h->OUT.as<double>(); sa.order = h->order.as<int32_t>(); sa.nslots_batch = nmine; sa.ns2 = ns2; sa.nf2 = nf2;
      sa.swap2 = h2.first_is_conv_second ? 0 : 1;
      sa.n_outer = std::max(0, pl.win[5] - pl.win[4] + 1); sa.n_inner = std::max(0, pl.win[7] - pl.win[6] + 1);
      sa.lo_outer = pl.win[4]; sa.lo_inner = pl.win[6];
      sa.conv = pl.conv; sa.symmetric = pl.symmetric; sa.intra = pl.intra ? 1 : 0;
      sa.slot_a = h->sa.as<int32_t>() + s_lo; sa.slot_b = h->sb.as<int32_t>() + s_lo;
      sa.nA = h->sp[pl.a].n; sa.nB = h->sp[pl.b].n; sa.tol = cons.tol;
      const int64_t ncand = (int64_t)nmine * sa.n_outer * sa.n_inner;
      const int64_t nblk = ceil_div(ncand, SEL_CHUNK);
      if (nblk > 0) {
        if (nblk > 2147483647LL) return fail(h, "too many candidates for one download; use the streaming form");
        CK(h->blockcount.ensure(nblk * sizeof(unsigned)));
        CK(h->blockoff.ensure(nblk * sizeof(int64_t)));
        CK(h->running.ensure(sizeof(unsigned long long)));
        CK(cudaMemsetAsync(h->running.p, 0, sizeof(unsigned long long), h->stream));
        select_count_kernel<<<(unsigned)nblk, SEL_THREADS, 0, h->stream>>>(sa, ncand, h->blockcount.as<unsigned>());
        select_scan_kernel<<<1, 1024, 0, h->stream>>>(h->blockcount.as<unsigned>(), nblk, h->blockoff.as<int64_t>(),
                                                      h->running.as<unsigned long long>());
        h->launches += 2;
        unsigned long long total = 0;
        CK(cudaMemcpyAsync(&total, h->running.p, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->count = (int64_t)total;
        const size_t n = std::max<size_t>(total, 1);
        EmitArgs ea{};
        CK(h->r_v.ensure(n * sizeof(double)));
        ea.o_v = h->r_v.as<double>(); ea.capacity = (int64_t)total; ea.overflow = h->overflow.as<int>();
        if (pl.conv == LOWDIN_IT_CONV_E) {
          CK(h->r_i0.ensure(n * sizeof(int64_t))); CK(h->r_i1.ensure(n * sizeof(int64_t)));
          ea.o_ij = h->r_i0.as<int64_t>(); ea.o_kl = h->r_i1.as<int64_t>();
        } else {
          CK(h->r_i0.ensure(n * sizeof(int32_t))); CK(h->r_i1.ensure(n * sizeof(int32_t)));
          CK(h->r_i2.ensure(n * sizeof(int32_t))); CK(h->r_i3.ensure(n * sizeof(int32_t)));
          ea.o_p = h->r_i0.as<int32_t>(); ea.o_q = h->r_i1.as<int32_t>(); ea.o_r = h->r_i2.as<int32_t>(); ea.o_s = h->r_i3.as<int32_t>();
        }
        select_emit_kernel<<<(unsigned)nblk, SEL_THREADS, 0, h->stream>>>(sa, ncand, h->blockoff.as<int64_t>(), ea);
        h->launches += 1;
        CK(cudaGetLastError());
        // entries kept per window pair: the segments a host merges the ranks' lists by (lowdin_it_result_segments)
        CK(h->seg.ensure((size_t)nmine * sizeof(unsigned long long)));
        select_segments_kernel<<<(unsigned)nmine, SEL_THREADS, 0, h->stream>>>(sa, h->seg.as<unsigned long long>());
        h->launches += 1;
        CK(cudaGetLastError());
        std::vector<unsigned long long> seg(nmine);
        CK(cudaMemcpyAsync(seg.data(), h->seg.p, (size_t)nmine * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->res_seg.assign(seg.begin(), seg.end());
      } else h->res_seg.assign(h->res_pair.size(), 0);
    }
    CK(cudaEventRecord(h->ev[3], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    prof_drain(h);
    {
      float ms;
      CK(cudaEventElapsedTime(&ms, h->ev[1], h->ev[2])); ms_second += ms;
      CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); ms_consume += ms;
    }
    h->timers[1] += ms_first * 1e-3; h->timers[2] += ms_exch * 1e-3; h->timers[3] += ms_second * 1e-3; h->timers[4] += ms_consume * 1e-3;
  }
  if (cons.mode == 1 && sums_out) {
    CK(cudaMemcpyAsync(sums_out, h->sums.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  h->timers[6] = flops;
  h->timers[7] = h->launches;
  return 0;
}

// All-to-all of one chunk of the half-transformed block between the halves.  Every rank holds
// H[slot][its own wblk columns of the chunk] for ALL slots; afterwards it holds, for ITS slots, all columns as G
// blocks [g][slot_local][wblk], which the chunk expansion reads in place through the chunk's column table (SRC_RECT_TABLE).
int exchange_chunk(lowdin_it_handle h, const std::vector<int> &own, int64_t wblk, DevBuf &Hsrc, DevBuf &H2dst, cudaStream_t st) {
  const int G = h->nranks;
  if (!h->comm && !h->lgroup) return fail(h, "multi-GPU transform without a communicator");
  const int mine = own[h->rank + 1] - own[h->rank];
  CK(H2dst.ensure(std::max<size_t>((size_t)std::max(mine, 1) * wblk * G, 1) * sizeof(double)));
  if (h->lgroup) {
    // in-process group: every rank PULLS its slots' rows from every peer's H with peer copies on its own stream
    LocalGroup &L = *h->lgroup;
    const int r = h->rank;
    L.H[r] = Hsrc.as<double>();
    CK(cudaEventRecord(L.ready[r], st));  // my first half of this chunk is complete behind this event
    if (!L.barrier()) return fail(h, "in-process group: a rank did not reach the exchange");
    for (int g = 0; g < G; ++g) {
      CK(cudaStreamWaitEvent(st, L.ready[g], 0));
      if (mine > 0 && wblk > 0)
        CK(cudaMemcpyPeerAsync(H2dst.as<double>() + (size_t)g * mine * wblk, h->device, L.H[g] + (size_t)own[r] * wblk, L.device[g],
                               (size_t)mine * wblk * sizeof(double), st));
    }
    CK(cudaEventRecord(L.done[r], st));
    if (!L.barrier()) return fail(h, "in-process group: a rank did not reach the exchange");
    for (int g = 0; g < G; ++g) CK(cudaStreamWaitEvent(st, L.done[g], 0));  // my H is rewritten only after every peer has pulled
    h->launches += 1;
    return 0;
  }
  if (h->node && h->node_mapped) {
    // one node, one process per GPU: every rank pulls the rows of its slots from every peer's (IPC-mapped) chunk buffer with copy
    // engines, starting at its right-hand neighbour so that the G pulls of a step hit G different sources
    NodeLink &N = *h->node;
    const int r = h->rank, b = (Hsrc.p == h->Hx.p) ? 1 : 0;
    CK(cudaEventRecord(N.ready[b], st));  // my first half of this chunk is complete behind this event
    if (!N.barrier()) return fail(h, "node link: a rank did not reach the exchange");
    for (int k = 1; k <= G; ++k) {
      const int g = (r + k) % G;
      if (g != r) CK(cudaStreamWaitEvent(st, N.peer_ready[g][b], 0));
      if (mine > 0 && wblk > 0) {
        const double *src = (g == r) ? Hsrc.as<double>() : static_cast<const double *>(N.peer_mem[g][b]);
        if (!src) return fail(h, "node link: a peer's chunk buffer is not mapped");
        CK(cudaMemcpyAsync(H2dst.as<double>() + (size_t)g * mine * wblk, src + (size_t)own[r] * wblk, (size_t)mine * wblk * sizeof(double),
                           cudaMemcpyDeviceToDevice, st));
      }
    }
    CK(cudaEventRecord(N.done[b], st));
    if (!N.barrier()) return fail(h, "node link: a rank did not reach the exchange");
    for (int g = 0; g < G; ++g)
      if (g != r) CK(cudaStreamWaitEvent(st, N.peer_done[g][b], 0));  // my H is rewritten only after every peer has pulled
    h->launches += 1;
    return 0;
  }
  int rc = g_nccl.GroupStart();
  for (int g = 0; g < G && rc == 0; ++g) {
    const size_t send_n = (size_t)(own[g + 1] - own[g]) * wblk;   // rows own[g]..own[g+1] of H (row stride wblk)
    const size_t recv_n = (size_t)mine * wblk;
    if (send_n) rc = g_nccl.Send(Hsrc.as<double>() + (size_t)own[g] * wblk, send_n * sizeof(double), /*ncclChar*/ 0, g, h->comm, st);
    if (rc == 0 && recv_n) rc = g_nccl.Recv(H2dst.as<double>() + (size_t)g * mine * wblk, recv_n * sizeof(double), 0, g, h->comm, st);
  }
  if (rc == 0) rc = g_nccl.GroupEnd();
  if (rc != 0) return fail(h, std::string("NCCL all-to-all failed: ") + g_nccl.GetErrorString(rc));
  h->launches += 1;
  return 0;
}

// Contexts behind the handle-less transformer-D compatible entry points: one per device, on the device selected by
// LOWDIN_IT_DEVICE or else the calling thread's current CUDA device (a multi-GPU host binds each process/thread to its GPU).
lowdin_it_handle g_dcompat_dev[64] = {};
thread_local lowdin_it_handle g_dcompat = nullptr;

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int lowdin_it_create(int device, lowdin_it_handle *out) {
  lowdin_it_handle h = nullptr;
  if (!out) return fail(nullptr, "null output handle");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, "device index out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return fail(nullptr, "this build targets sm_100a (B200) only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
  h = new lowdin_it_ctx();
  h->device = device;
  h->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, "cudaStreamCreate failed"); }
  for (auto &ev : h->ev) cudaEventCreate(&ev);
  *out = h;
  return 0;
}

int lowdin_it_destroy(lowdin_it_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->node.reset();
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (auto &s : h->sp) { s.C.release(); s.Cs.release(); s.pi.release(); s.pj.release(); s.bs_shells.release(); s.bs_fn.release(); s.bs_expo.release(); s.bs_coef.release(); }
  for (auto &row : h->ao) for (auto &a : row) { a.data.release(); a.fa.release(); a.fb.release(); a.release_list(); a.seg_counts.release(); }
  DevBuf *bufs[] = {&h->st[0], &h->st[1], &h->up_state, &h->T1list, &h->Cw, &h->coltab, &h->coltabx, &h->Hx, &h->H2x, &h->seg, &h->X, &h->T1t, &h->H, &h->H2, &h->OUT, &h->T3, &h->order, &h->tab, &h->sa, &h->sb,
                    &h->ss, &h->sf, &h->blockcount, &h->blockoff, &h->sums, &h->running, &h->overflow, &h->epsA, &h->epsB, &h->dtmp, &h->agree,
                    &h->r_i0, &h->r_i1, &h->r_i2, &h->r_i3, &h->r_v};
  for (DevBuf *b : bufs) b->release();
  for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : h->prof_pool) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) { if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); if (h->ev_scattered[i]) cudaEventDestroy(h->ev_scattered[i]); }
  for (int i = 0; i < 2; ++i) { if (h->sink_host[i]) cudaFreeHost(h->sink_host[i]); if (h->ev_q4[i]) cudaEventDestroy(h->ev_q4[i]); if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < 2; ++i) for (cudaEvent_t e : {h->ev_fh[i], h->ev_ex[i], h->ev_sh[i]}) if (e) cudaEventDestroy(e);
  for (auto &ev : h->chunk_ev) cudaEventDestroy(ev);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

const char *lowdin_it_last_error(lowdin_it_handle h) {
  if (h) return h->err.c_str();
  // no handle: the handle-less entry points (lowdin_it_create, lowdin_it_transform_all / _inter_all)
  if (g_create_error.empty() && g_dcompat && !g_dcompat->err.empty()) return g_dcompat->err.c_str();
  return g_create_error.c_str();
}

int lowdin_it_set_species(lowdin_it_handle h, int slot, int nao, const double *C, int ldc, int ncols) {
  if (!h) return 1;
  if (slot < 0 || slot > 7) return fail(h, "species slot out of range");
  if (nao <= 0 || ncols <= 0 || ldc < nao || !C) return fail(h, "bad coefficient matrix arguments");
  CK(cudaSetDevice(h->device));
  Species &S = h->sp[slot];
  S.n = nao; S.ncols = ncols; S.M = npairs(nao); S.ldc = roundup2(nao);
  CK(S.C.ensure((size_t)S.ldc * ncols * sizeof(double)));
  CK(cudaMemsetAsync(S.C.p, 0, (size_t)S.ldc * ncols * sizeof(double), h->stream));
  CK(cudaMemcpy2DAsync(S.C.p, S.ldc * sizeof(double), C, (size_t)ldc * sizeof(double), (size_t)nao * sizeof(double), ncols,
                       cudaMemcpyHostToDevice, h->stream));
  {
    const int64_t cnt = S.ldc * (int64_t)ncols;
    CK(S.Cs.ensure((size_t)cnt * sizeof(double)));
    scale_copy_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, h->stream>>>(S.C.as<double>(), S.Cs.as<double>(), cnt, 0x1p-53);
    CK(cudaGetLastError());
  }
  std::vector<int32_t> pi(S.M), pj(S.M);
  int64_t m = 0;
  for (int i = 0; i < nao; ++i) for (int j = i; j < nao; ++j) { pi[m] = i; pj[m] = j; ++m; }
  if (upload_i32(h, S.pi, pi) || upload_i32(h, S.pj, pj)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  for (int o = 0; o < 8; ++o) { h->ao[slot][o].valid = false; h->ao[o][slot].valid = false; }
  return 0;
}

int lowdin_it_ao_begin(lowdin_it_handle h, int a, int b, int swapped) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->sp[a].n || !h->sp[b].n) return fail(h, "ao_begin: species not set");
  CK(cudaSetDevice(h->device));
  AoSet &S = h->ao[a][b];
  const int64_t Ma = h->sp[a].M, Mb = h->sp[b].M;
  const size_t count = (a == b) ? (size_t)(Ma * (Ma + 1) / 2) : (size_t)(Ma * Mb);
  if (count * sizeof(double) > ((size_t)1 << 30)) release_workspaces(h);
  S.release_list();
  S.swapped = swapped;
  if (h->ao_list) {  // keep the list as it comes (16 bytes per integral); the first quarter is then list-driven (it_list.cuh)
    if (h->sp[a].n > 65535 || h->sp[b].n > 65535) return fail(h, "ao_begin: the resident list stores 16-bit AO indices");
    S.data.release();
    CK(S.seg_counts.ensure(kMaxListSegs * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(S.seg_counts.p, 0, kMaxListSegs * sizeof(unsigned long long), h->stream));
    S.src = AoSource{SRC_LIST, nullptr, Ma, 0, Mb, 0};
    S.sharded = false;
  } else if (h->nranks > 1) {
    // On a communicator a rank stores only the slabs it computes in the first half (block-cyclic, it_kernels.cuh slab_owner): the
    // full M_a-vector of each, i.e. 2/G of the packed tensor per rank.  Every rank is pushed the whole list and keeps its rows.
    const int64_t nloc = slabs_owned_below(Mb, h->slab_logB, h->nranks, h->rank);
    const size_t cnt = std::max<size_t>((size_t)nloc * Ma, 1);
    CK(S.data.ensure(cnt * sizeof(double)));
    CK(cudaMemsetAsync(S.data.p, 0, cnt * sizeof(double), h->stream));
    S.src = AoSource{SRC_RECT, S.data.as<double>(), Ma, Ma, nloc, 0};
    S.sharded = true;
  } else {
    CK(S.data.ensure(count * sizeof(double)));
    CK(cudaMemsetAsync(S.data.p, 0, count * sizeof(double), h->stream));  // C.f90:669-685 zero-initialises
    S.src = (a == b) ? AoSource{SRC_SYM_PACKED, S.data.as<double>(), Ma, 0, 0, 0} : AoSource{SRC_RECT, S.data.as<double>(), Ma, Ma, Mb, 0};
    S.sharded = false;
  }
  S.valid = false;
  h->up_a = a; h->up_b = b; h->up_swapped = swapped;
  h->up_pos = 0; h->up_piece = 0; h->st_busy[0] = h->st_busy[1] = false;
  CK(h->up_state.ensure(2 * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(h->up_state.p, 0xff, sizeof(unsigned long long), h->stream));                              // no terminator yet
  CK(cudaMemsetAsync(h->up_state.as<unsigned long long>() + 1, 0xff, sizeof(unsigned long long), h->stream));   // no bad entry yet
  CK(cudaEventRecord(h->ev[5], h->stream));
  h->timers[0] = 0;
  return 0;
}

namespace {
// One piece of an upload: `cnt` entries (a whole number of stacks of S entries, or one partial stack) whose five arrays start at
// host pointers p..v, stack t at + t*stride.  The bytes go to a staging buffer on the copy stream; terminator search, index check
// and scatter run on the library's stream behind them (it_kernels.cuh, scatter_stacks_kernel).
int push_piece(lowdin_it_handle h, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s, const double *v, int64_t S,
               int64_t nstk, int64_t cnt, bool raw_blocks, int64_t call_pos) {
  const int slot = h->up_piece & 1;
  ++h->up_piece;
  if (h->st_busy[slot]) CK(cudaStreamWaitEvent(h->copy_stream, h->ev_scattered[slot], 0));  // the staging buffer's previous piece is scattered
  uint8_t *st = h->st[slot].as<uint8_t>();
  StackView w{};
  if (raw_blocks) {  // contiguous blocks of 24 S bytes: one copy
    CK(cudaMemcpyAsync(st, p, (size_t)nstk * 24 * S, cudaMemcpyHostToDevice, h->copy_stream));
    w.p = (const int32_t *)st; w.q = w.p + S; w.r = w.p + 2 * S; w.s = w.p + 3 * S; w.v = (const double *)(st + 16 * S);
    w.stride_i = 6 * S; w.stride_v = 3 * S;
  } else {           // five separate arrays, one (partial) stack
    const int64_t sp = roundup2(cnt);
    CK(cudaMemcpyAsync(st, p, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(st + 4 * sp, q, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(st + 8 * sp, r, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(st + 12 * sp, s, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(st + 16 * sp, v, cnt * 8, cudaMemcpyHostToDevice, h->copy_stream));
    w.p = (const int32_t *)st; w.q = w.p + sp; w.r = w.p + 2 * sp; w.s = w.p + 3 * sp; w.v = (const double *)(st + 16 * sp);
    w.stride_i = 0; w.stride_v = 0;
  }
  w.S = S; w.total = cnt; w.pos0 = call_pos;
  CK(cudaEventRecord(h->ev_copied[slot], h->copy_stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
  const int na = h->sp[h->up_a].n, nb = h->sp[h->up_b].n;
  AoSet &AS = h->ao[h->up_a][h->up_b];
  unsigned long long *state = h->up_state.as<unsigned long long>();
  const unsigned grid = (unsigned)ceil_div(cnt, 256);
  find_terminator_kernel<<<grid, 256, 0, h->stream>>>(w, state);
  if (AS.src.kind == SRC_LIST) {
    if ((int)AS.segs.size() >= kMaxListSegs) return fail(h, "ao_push: too many list segments; push larger pieces");
    AS.segs.emplace_back();
    CK(AS.segs.back().ensure((size_t)cnt * sizeof(ListEntry)));
    append_list_kernel<<<grid, 256, 0, h->stream>>>(w, h->up_a == h->up_b, h->up_swapped, na, nb, AS.segs.back().as<ListEntry>(),
                                                    AS.seg_counts.as<unsigned long long>() + (AS.segs.size() - 1), state);
  } else {
    ScatterDst d{AS.data.as<double>(), h->up_a == h->up_b, h->up_swapped, na, nb, AS.sharded ? 1 : 0, h->slab_logB, h->nranks, h->rank};
    scatter_stacks_kernel<<<grid, 256, 0, h->stream>>>(w, d, state);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev_scattered[slot], h->stream));
  h->st_busy[slot] = true;
  if (!h->async_push) CK(cudaEventSynchronize(h->ev_copied[slot]));  // the caller may reuse its buffers
  return 0;
}

int push_begin_call(lowdin_it_handle h) {
  if (h->up_a < 0) return fail(h, "ao_push without ao_begin");
  CK(cudaSetDevice(h->device));
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_scattered[i], cudaEventDisableTiming));
    }
  }
  for (int i = 0; i < 2; ++i) CK(h->st[i].ensure(h->staging_bytes));
  // the terminator ends the stream of THIS call (one file, C.f90:279-280); the next call is the next file
  CK(cudaMemsetAsync(h->up_state.p, 0xff, sizeof(unsigned long long), h->stream));
  return 0;
}
}  // namespace

int lowdin_it_ao_push_stacks(lowdin_it_handle h, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                             const double *v, int64_t n) {
  if (!h) return 1;
  if (n < 0 || (n > 0 && (!p || !q || !r || !s || !v))) return fail(h, "ao_push_stacks: bad arguments");
  if (push_begin_call(h)) return 1;
  const int64_t cap = (int64_t)(h->staging_bytes / 24) - 2;
  for (int64_t o = 0; o < n; o += cap) {
    const int64_t cnt = std::min<int64_t>(cap, n - o);
    if (push_piece(h, p + o, q + o, r + o, s + o, v + o, cnt, 1, cnt, false, o)) return 1;
  }
  h->up_pos += n;
  return 0;
}

int lowdin_it_ao_push_blocks(lowdin_it_handle h, const void *blocks, int64_t nblocks, int stack_size) {
  if (!h) return 1;
  if (nblocks < 0 || stack_size < 1 || (nblocks > 0 && !blocks)) return fail(h, "ao_push_blocks: bad arguments");
  if ((size_t)stack_size * 24 > h->staging_bytes) return fail(h, "ao_push_blocks: stack larger than the staging buffer");
  if (push_begin_call(h)) return 1;
  const int64_t S = stack_size, per = (int64_t)(h->staging_bytes / (24 * (size_t)S));
  const uint8_t *src = (const uint8_t *)blocks;
  for (int64_t t = 0; t < nblocks; t += per) {
    const int64_t nb_ = std::min<int64_t>(per, nblocks - t);
    if (push_piece(h, (const int32_t *)(src + (size_t)t * 24 * S), nullptr, nullptr, nullptr, nullptr, S, nb_, nb_ * S, true, t * S)) return 1;
  }
  h->up_pos += nblocks * S;
  return 0;
}

int lowdin_it_ao_end(lowdin_it_handle h) {
  if (!h) return 1;
  if (h->up_a < 0) return fail(h, "ao_end without ao_begin");
  CK(cudaSetDevice(h->device));
  unsigned long long state[2] = {0, 0};
  CK(cudaMemcpyAsync(state, h->up_state.p, sizeof(state), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev[0], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev[5], h->ev[0]));
  h->timers[0] = ms * 1e-3;
  const int a = h->up_a, b = h->up_b;
  h->up_a = h->up_b = -1;
  if (state[1] != ~0ull)  // nothing of a list with a bad entry is trusted
    return fail(h, "AO stack entry " + std::to_string(state[1]) + " (1-based, counted within its push call) has an index outside the basis");
  h->ao[a][b].valid = true;
  return 0;
}

int lowdin_it_ao_set_generator(lowdin_it_handle h, int a, int b, int kind, uint64_t seed) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->sp[a].n || !h->sp[b].n) return fail(h, "ao_set_generator: species not set");
  if (kind != LOWDIN_IT_GEN_HASH && kind != LOWDIN_IT_GEN_FOLD) return fail(h, "unknown generator kind");
  AoSet &S = h->ao[a][b];
  S.data.release();
  S.release_list();
  S.sharded = false;
  S.src = (a == b) ? AoSource{SRC_HASH_SYM, nullptr, h->sp[a].M, 0, 0, seed, kind} : AoSource{SRC_HASH_RECT, nullptr, h->sp[a].M, 0, h->sp[b].M, seed, kind};
  S.valid = true;
  return 0;
}

int lowdin_it_ao_set_rankk(lowdin_it_handle h, int a, int b, int K, const double *La, const double *Lb) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->sp[a].n || !h->sp[b].n) return fail(h, "ao_set_rankk: species not set");
  if (K < 1 || K > RANKK || !La || (a != b && !Lb)) return fail(h, "ao_set_rankk: 1 <= K <= 8 factor sets required");
  CK(cudaSetDevice(h->device));
  AoSet &S = h->ao[a][b];
  S.data.release();
  S.release_list();
  S.sharded = false;
  const int64_t Ma = h->sp[a].M, Mb = h->sp[b].M;
  CK(S.fa.ensure((size_t)RANKK * Ma * sizeof(double)));
  CK(cudaMemsetAsync(S.fa.p, 0, (size_t)RANKK * Ma * sizeof(double), h->stream));
  CK(cudaMemcpyAsync(S.fa.p, La, (size_t)K * Ma * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const double *fb = S.fa.as<double>();
  if (a != b) {
    CK(S.fb.ensure((size_t)RANKK * Mb * sizeof(double)));
    CK(cudaMemsetAsync(S.fb.p, 0, (size_t)RANKK * Mb * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(S.fb.p, Lb, (size_t)K * Mb * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    fb = S.fb.as<double>();
  }
  CK(cudaStreamSynchronize(h->stream));
  S.src = AoSource{SRC_RANKK, S.fa.as<double>(), Ma, 0, (a == b) ? Ma : Mb, 0, LOWDIN_IT_GEN_RANKK, fb};
  S.valid = true;
  return 0;
}

int lowdin_it_ao_materialize(lowdin_it_handle h, int a, int b) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->ao[a][b].valid) return fail(h, "ao_materialize: AO set not available");
  CK(cudaSetDevice(h->device));
  AoSet &S = h->ao[a][b];
  const AoSource src = S.src;
  if (src.kind != SRC_HASH_SYM && src.kind != SRC_HASH_RECT && src.kind != SRC_RANKK) return fail(h, "ao_materialize: the AO set is not a generated one");
  const bool intra = (a == b);
  const int64_t Ma = h->sp[a].M, Mb = h->sp[b].M;
  release_workspaces(h);
  const bool shard = h->nranks > 1;  // on a communicator: the rank's own rows only, as an upload would leave them
  const int64_t nrows = shard ? slabs_owned_below(Mb, h->slab_logB, h->nranks, h->rank) : Mb;
  const size_t count = shard ? std::max<size_t>((size_t)nrows * Ma, 1) : (intra ? (size_t)(Ma * (Ma + 1) / 2) : (size_t)(Ma * Mb));
  DevBuf fresh;  // the generator's factor arrays (kind K) stay alive until the fill is done
  CK(fresh.ensure(count * sizeof(double)));
  AoSource gsrc = src;
  if (shard) { gsrc.logB = h->slab_logB; gsrc.G = h->nranks; gsrc.rank = h->rank; }
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(nrows, (int64_t)h->num_sms * 16));
  double *dst = fresh.as<double>();
  switch (src.kind) {
    case SRC_HASH_SYM: materialize_kernel<SRC_HASH_SYM><<<grid, 256, 0, h->stream>>>(gsrc, 1, Ma, nrows, dst); break;
    case SRC_HASH_RECT: materialize_kernel<SRC_HASH_RECT><<<grid, 256, 0, h->stream>>>(gsrc, 0, Ma, nrows, dst); break;
    default: materialize_kernel<SRC_RANKK><<<grid, 256, 0, h->stream>>>(gsrc, intra ? 1 : 0, Ma, nrows, dst); break;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  S.data.release();
  S.data = fresh;
  S.sharded = shard;
  S.src = shard ? AoSource{SRC_RECT, dst, Ma, Ma, nrows, 0} : (intra ? AoSource{SRC_SYM_PACKED, dst, Ma, 0, 0, 0} : AoSource{SRC_RECT, dst, Ma, Ma, Mb, 0});
  return 0;
}

// ---- row f4: the AO integrals evaluated on the device (it_eri.cuh) ---------------------------------------------
int lowdin_it_set_basis(lowdin_it_handle h, int slot, int nshells, const lowdin_it_shell *shells, const double *exponents,
                        const double *coefficients) {
  if (!h) return 1;
  if (slot < 0 || slot > 7 || !h->sp[slot].n) return fail(h, "set_basis: species not set (lowdin_it_set_species first)");
  if (nshells < 1 || !shells || !exponents || !coefficients) return fail(h, "set_basis: null argument");
  CK(cudaSetDevice(h->device));
  Species &S = h->sp[slot];
  EriHostBasis hb;
  if (!eri_prepare_basis(nshells, shells, exponents, coefficients, hb)) return fail(h, "set_basis: " + hb.err);
  const std::vector<EriShell> &sh = hb.sh;
  const std::vector<EriFunction> &fn = hb.fn;
  const std::vector<double> &expo = hb.expo, &coef = hb.coef;
  const int nprim = (int)expo.size();
  S.bs_norma.clear();
  for (const EriFunction &f : fn) S.bs_norma.push_back(f.norma);
  if ((int)fn.size() != S.n) return fail(h, "set_basis: the shells hold " + std::to_string(fn.size()) + " Cartesian functions, the species has " + std::to_string(S.n));
  CK(S.bs_shells.ensure(sh.size() * sizeof(EriShell)));
  CK(S.bs_fn.ensure(fn.size() * sizeof(EriFunction)));
  CK(S.bs_expo.ensure(nprim * sizeof(double)));
  CK(S.bs_coef.ensure(nprim * sizeof(double)));
  CK(cudaMemcpyAsync(S.bs_shells.p, sh.data(), sh.size() * sizeof(EriShell), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(S.bs_fn.p, fn.data(), fn.size() * sizeof(EriFunction), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(S.bs_expo.p, expo.data(), nprim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(S.bs_coef.p, coef.data(), nprim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  S.bs_nbf = (int)fn.size();
  return 0;
}

int lowdin_it_basis_norma(lowdin_it_handle h, int slot, double *norma) {
  if (!h) return 1;
  if (slot < 0 || slot > 7 || !h->sp[slot].bs_nbf || !norma) return fail(h, "basis_norma: basis not set");
  std::copy(h->sp[slot].bs_norma.begin(), h->sp[slot].bs_norma.end(), norma);
  return 0;
}

int lowdin_it_ao_compute(lowdin_it_handle h, int a, int b) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->sp[a].bs_nbf || !h->sp[b].bs_nbf) return fail(h, "ao_compute: basis not set (lowdin_it_set_basis)");
  if (h->ao_list) return fail(h, "ao_compute fills the stored tensor; unset LOWDIN_IT_OPT_AO_LIST");
  CK(cudaSetDevice(h->device));
  AoSet &S = h->ao[a][b];
  const bool intra = (a == b);
  const int64_t Ma = h->sp[a].M, Mb = h->sp[b].M;
  release_workspaces(h);
  S.release_list();
  const bool shard = h->nranks > 1;  // on a communicator: the rows this rank owns in the first half, whole M-vectors (as an upload leaves them)
  const int64_t nrows = shard ? slabs_owned_below(Mb, h->slab_logB, h->nranks, h->rank) : Mb;
  const size_t count = shard ? std::max<size_t>((size_t)nrows * Ma, 1) : (intra ? (size_t)(Ma * (Ma + 1) / 2) : (size_t)(Ma * Mb));
  CK(S.data.ensure(count * sizeof(double)));
  double *dst = S.data.as<double>();
  const Species &A = h->sp[a], &B = h->sp[b];
  EriBasis ba{A.bs_shells.as<EriShell>(), A.bs_fn.as<EriFunction>(), A.bs_expo.as<double>(), A.bs_coef.as<double>(), A.bs_nbf};
  EriBasis bb{B.bs_shells.as<EriShell>(), B.bs_fn.as<EriFunction>(), B.bs_expo.as<double>(), B.bs_coef.as<double>(), B.bs_nbf};
  EriFillArgs fa{ba, bb, shard ? 2 : (intra ? 0 : 1), Ma, nrows, h->slab_logB, h->nranks, h->rank, intra ? 1 : 0, dst};
  if (ceil_div(Ma, ERI_FILL_THREADS) > 65535) return fail(h, "ao_compute: basis too large for one launch");
  cudaEvent_t e0 = h->ev[5], e1 = h->ev[0];
  CK(cudaEventRecord(e0, h->stream));
  if (nrows > 0) {
    eri_fill_kernel<<<eri_fill_grid(fa), ERI_FILL_THREADS, 0, h->stream>>>(fa);
    h->launches += 1;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->timers[0] = ms * 1e-3;  // reported like an upload
  S.swapped = 0;
  S.sharded = shard;
  S.src = shard ? AoSource{SRC_RECT, dst, Ma, Ma, nrows, 0} : (intra ? AoSource{SRC_SYM_PACKED, dst, Ma, 0, 0, 0} : AoSource{SRC_RECT, dst, Ma, Ma, Mb, 0});
  S.valid = true;
  return 0;
}

/* the stored AO tensor of a pair back on the host (single rank): packed M(M+1)/2 (intra) or [M_b][M_a] (inter) */
int lowdin_it_ao_download(lowdin_it_handle h, int a, int b, double *out, int64_t capacity) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->ao[a][b].valid) return fail(h, "ao_download: AO set not available");
  const AoSet &S = h->ao[a][b];
  if (S.sharded || (S.src.kind != SRC_SYM_PACKED && S.src.kind != SRC_RECT)) return fail(h, "ao_download: the AO set is not a whole stored tensor");
  const int64_t Ma = h->sp[a].M, Mb = h->sp[b].M;
  const int64_t count = (a == b) ? Ma * (Ma + 1) / 2 : Ma * Mb;
  if (!out || capacity < count) return fail(h, "ao_download: buffer too small");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(out, S.data.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lowdin_it_transform(lowdin_it_handle h, int a, int b, const int win[8], int conv, int symmetric, double drop_tol) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  Plan pl;
  if (build_plan(h, a, b, win, conv, symmetric, pl)) return 1;
  Consumer cons; cons.mode = 0; cons.tol = drop_tol;
  // download mode keeps the dense result block of every window pair (of this rank's share of them, on a communicator) and the
  // third-quarter accumulators: check that they fit
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const double need = (double)pl.pairs_s.size() / h->nranks * ((double)pl.h2.ns * pl.h2.nf + (double)pl.h2.nf * roundup2(pl.h2.nc)) * 8.0;
  if (need > 0.8 * ((double)free_b + (double)h->T3.cap + (double)h->OUT.cap + (double)h->H.cap))
    return fail(h, "result block does not fit in device memory; use lowdin_it_transform_stream");
  return run_passes(h, pl, 0, 0, 0, cons, nullptr);
}

int lowdin_it_result_count(lowdin_it_handle h, int64_t *count) {
  if (!h || !count) return 1;
  *count = h->count;
  return 0;
}

int lowdin_it_result_segments(lowdin_it_handle h, int64_t *npairs, int64_t *pair_index, int64_t *kept) {
  if (!h || !npairs) return 1;
  *npairs = (int64_t)h->res_pair.size();
  if (pair_index) memcpy(pair_index, h->res_pair.data(), h->res_pair.size() * sizeof(int64_t));
  if (kept) memcpy(kept, h->res_seg.data(), h->res_seg.size() * sizeof(int64_t));
  return 0;
}

int lowdin_it_download_pairs(lowdin_it_handle h, int64_t *ij, int64_t *kl, double *v) {
  if (!h) return 1;
  if (h->res_conv != LOWDIN_IT_CONV_E) return fail(h, "last transform did not use the E (pair id) convention");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (h->count) {
    CK(cudaMemcpyAsync(ij, h->r_i0.p, h->count * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(kl, h->r_i1.p, h->count * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(v, h->r_v.p, h->count * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->timers[5] = ms * 1e-3;
  int ov = 0;
  if (h->overflow.p) CK(cudaMemcpy(&ov, h->overflow.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (ov) return fail(h, "result buffer overflow");
  return 0;
}

int lowdin_it_download_quads(lowdin_it_handle h, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v) {
  if (!h) return 1;
  if (h->res_conv != LOWDIN_IT_CONV_C) return fail(h, "last transform did not use the C (quad) convention");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (h->count) {
    CK(cudaMemcpyAsync(p, h->r_i0.p, h->count * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(q, h->r_i1.p, h->count * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(r, h->r_i2.p, h->count * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(s, h->r_i3.p, h->count * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(v, h->r_v.p, h->count * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->timers[5] = ms * 1e-3;
  int ov = 0;
  if (h->overflow.p) CK(cudaMemcpy(&ov, h->overflow.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (ov) return fail(h, "result buffer overflow");
  return 0;
}

// Occupied batch (first-window values per pass) from a cost model of what depends on it, everything else (second and fourth
// quarter, bytes exchanged) being the same for every choice:
//  * first quarter: the fused kernels take the window columns in balanced groups of <= 64 (launch_q1_gen) and a group of w columns
//    runs at eff(w) of the tensor rate (kernels alone on B200, profiles/r02i_q_probe.log: 16 -> 14.0, 32 -> 23.7, 40 -> 25.8,
//    48 -> 26.9, 56 -> 28.2, 64 -> 28.9 TFLOP/s);
//  * third quarter: its flops at ~33 TFLOP/s plus one read-modify-write sweep of the accumulators T3 per chunk at ~6 TB/s, the number
//    of chunks following from the memory the batch leaves for the chunk buffers;
//  * a stored AO tensor is read again by every pass (row completion + fused first quarter: ~24 M_a bytes per slab).
// More passes are not a cost in themselves: 150 occupied orbitals on TWO GPUs run better as 56 + 56 + 38 (wide chunks, third quarter
// ~28 TFLOP/s) than as 110 + 40 (measured: third quarter 14.3 TFLOP/s, profiles/r02t_bench_n1500_g2.json) or 80 + 70 (first quarter in
// groups of 40 and 32 columns, 24.2 TFLOP/s, profiles/r02n_*).  Pure function of its arguments: every rank decides alike.
int occ_batch_model(int nf, int qmax, int G, double spf, int nf2, int n2, int n1, double nslabs1, double Ma, double avail, bool stored) {
  static const double eff[9] = {1.0, 7.5, 14.0, 19.5, 23.7, 25.8, 26.9, 28.2, 28.9};  // TFLOP/s per group of 8 u columns
  auto q1_cost = [&](int nfb) {  // seconds of first quarter per slab of the pass, times 1e12 / (2 n1^2)
    const int groups = (int)ceil_div(nfb, 64), units = (int)ceil_div(nfb, 8);
    double c = 0.0;
    for (int gi = 0; gi < groups; ++gi) {
      const int u = units / groups + (gi < units % groups ? 1 : 0);
      if (u > 0) c += 8.0 * u / eff[std::min(u, 8)];
    }
    return c;
  };
  const double ldt2 = (double)roundup2(n2), slabs_own = nslabs1 / G;
  auto pass_cost = [&](int q) {
    const double t3 = spf * q * nf2 * ldt2 * 8.0 / G;
    const double per_col = (G > 1) ? spf * q * 8.0 / G + 2.0 * (spf * q / G) * 8.0 + 8.0 : spf * q * 8.0;
    const double cols = std::max(std::min((avail - t3) / per_col, nslabs1), 2.0 * n2);
    const double nchunks = std::max(1.0, nslabs1 / cols);
    const double t_q1 = slabs_own * 2.0 * n1 * (double)n1 * q1_cost(q) * 1e-12;
    const double t_q3 = 2.0 * (spf * q / G) * nf2 * (double)n2 * n2 / 33e12 + nchunks * t3 / 6e12;
    const double t_ao = stored ? slabs_own * Ma * 24.0 / 5e12 : 0.0;
    return t_q1 + t_q3 + t_ao;
  };
  int best_q = std::max(1, std::min(qmax, nf));
  double best = -1.0;
  for (int q = std::max(1, std::min(qmax, nf)); q >= 1; --q) {  // from the largest batch down: ties go to fewer passes
    const double c = (nf / q) * pass_cost(q) + (nf % q ? pass_cost(nf % q) : 0.0);
    if (best < 0.0 || c < best * (1.0 - 1.5e-2)) { best = c; best_q = q; }  // a smaller batch must win by more than the model's noise
  }
  return best_q;
}

static int pick_occ_batch(lowdin_it_handle h, const Plan &pl, int requested, int *used) {
  const int nf = std::max(pl.h1.nf, 1);
  if (requested > 0) { *used = std::min(requested, nf); return 0; }
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const int G = h->nranks;
  const double spf = std::max(pl.max_slots_per_f, 1);
  const double per_out = (double)pl.h2.ns * pl.h2.nf * 8.0;
  // results buffer: 4 GB of groups, or two f-blocks when one alone is larger (the host sink double-buffers its groups, and the same
  // occupied batch must serve the pass with and without a sink)
  const double out_need = std::max(std::min(4.0e9, spf * per_out * nf), std::min(2.0, (double)nf) * spf * per_out);
  double avail = 0.92 * ((double)free_b + (double)h->H.cap + (double)h->OUT.cap + (double)h->H2.cap + (double)h->Hx.cap + (double)h->H2x.cap + (double)h->T3.cap +
                         (double)h->X.cap + (double)h->T1t.cap) - 2.0 * (double)h->workspace_bytes - out_need - (double)((size_t)1 << 30);
  // the same figure on every rank (collective when occ_batch == 0 on a communicator: every rank must make this call)
  int64_t avail64 = (int64_t)std::max(avail, 0.0);
  if (agree_min(h, &avail64)) return 1;
  avail = (double)avail64;
  // per first-window value: third-quarter accumulators of its slots (own share) + a chunk of at least 8 pair rows of H
  const double t3_per_f = spf * (double)pl.h2.nf * (double)roundup2(pl.h2.nc) * 8.0 / G;
  const double cols_min = (double)std::min<int64_t>(pl.nslabs1, 8LL * pl.h2.nc);
  const double hc_per_f = spf * 8.0 * cols_min * (G > 1 ? 3.0 / G : 1.0);
  const double list_per_f = (pl.src.kind == SRC_LIST) ? (double)pl.nslabs1 * pl.h1.nc * 8.0 / G : 0.0;  // T1list[own slab][nu][f]
  const int qmax = (int)std::max(1.0, std::min((double)nf, avail / (t3_per_f + hc_per_f + list_per_f)));
  const bool stored = (pl.src.kind == SRC_SYM_PACKED || pl.src.kind == SRC_RECT);
  *used = occ_batch_model(nf, qmax, G, spf, pl.h2.nf, pl.h2.nc, pl.h1.nc, (double)pl.nslabs1, (double)h->sp[pl.a].M, avail, stored);
  return 0;
}

int lowdin_it_stream_num_passes(lowdin_it_handle h, int a, int b, const int win[8], int conv, int occ_batch, int *n_passes,
                                int *occ_batch_used) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  Plan pl;
  if (build_plan(h, a, b, win, conv, 0, pl)) return 1;
  int used = 0;
  if (pick_occ_batch(h, pl, occ_batch, &used)) return 1;
  if (n_passes) *n_passes = (int)ceil_div(std::max(pl.h1.nf, 1), used);
  if (occ_batch_used) *occ_batch_used = used;
  return 0;
}

int lowdin_it_transform_stream(lowdin_it_handle h, int a, int b, const int win[8], int conv, double drop_tol, int occ_batch,
                               int first_pass, int n_passes, const double *epsA, const double *epsB, double lambda, double sums[4]) {
  return lowdin_it_transform_stream_sink(h, a, b, win, conv, drop_tol, occ_batch, first_pass, n_passes, epsA, epsB, lambda, sums, nullptr, nullptr);
}

int lowdin_it_transform_stream_sink(lowdin_it_handle h, int a, int b, const int win[8], int conv, double drop_tol, int occ_batch,
                                    int first_pass, int n_passes, const double *epsA, const double *epsB, double lambda, double sums[4],
                                    lowdin_it_sink_fn sink, void *user) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  Plan pl;
  if (build_plan(h, a, b, win, conv, 0, pl)) return 1;
  int used = 0;
  if (pick_occ_batch(h, pl, occ_batch, &used)) return 1;
  Consumer cons; cons.mode = 1; cons.tol = drop_tol; cons.lambda = lambda; cons.sink = sink; cons.sink_user = user;
  if (epsA) {
    const int na = h->sp[a].ncols, nb = h->sp[b].ncols;
    CK(h->epsA.ensure(na * sizeof(double)));
    CK(cudaMemcpyAsync(h->epsA.p, epsA, na * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    cons.epsA = h->epsA.as<double>();
    if (a != b) {
      if (!epsB) return fail(h, "epsB required for an inter-species energy sum");
      CK(h->epsB.ensure(nb * sizeof(double)));
      CK(cudaMemcpyAsync(h->epsB.p, epsB, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      cons.epsB = h->epsB.as<double>();
    } else cons.epsB = cons.epsA;
  }
  return run_passes(h, pl, used, first_pass, n_passes, cons, sums);
}

int lowdin_it_set_option(lowdin_it_handle h, int option, int64_t value) {
  if (!h) return 1;
  switch (option) {
    case LOWDIN_IT_OPT_WORKSPACE_BYTES:
      if (value < (1 << 12)) return fail(h, "workspace too small");
      h->workspace_bytes = (size_t)value; return 0;
    case LOWDIN_IT_OPT_CHUNK_COLS:
      if (value < 0) return fail(h, "negative chunk column limit");
      h->chunk_cols_limit = value; return 0;
    case LOWDIN_IT_OPT_Q1_VARIANT:
      if (value < 1 || value > 5) return fail(h, "q1 variant must be 1..5");
      h->q1_variant = (int)value; return 0;
    case LOWDIN_IT_OPT_GEMM_VARIANT:
      if (value != 1 && value != 2) return fail(h, "gemm variant must be 1 or 2");
      if (value == 2 && !tensor_map_encoder()) return fail(h, "cuTensorMapEncodeTiled is not available from this driver");
      h->gemm_variant = (int)value; return 0;
    case LOWDIN_IT_OPT_SPLIT_ROW_TAIL:
      h->split_row_tail = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_FRAG_PERM:
      h->frag_perm = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_SLAB_BLOCK_LOG:
      if (value < 0 || value > 20) return fail(h, "slab block size must be 2^0 .. 2^20");
      for (auto &row : h->ao) for (auto &a : row) if (a.sharded) return fail(h, "the slab block size cannot change while row-sharded AO integrals are resident");
      h->slab_logB = (int)value; return 0;
    case LOWDIN_IT_OPT_AO_LIST:
      h->ao_list = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_GEMM_TALL:
      h->gemm_tall = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_EXCHANGE_DMA:
      h->exchange_dma = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_Q3_TWO_CTA:
      if (value < 0 || value > 4096) return fail(h, "LOWDIN_IT_OPT_Q3_TWO_CTA: 0 (off) or the largest K handled by the two-CTA kernels");
      h->q3_two_cta = (int)value; return 0;
    case LOWDIN_IT_OPT_SINK_BLOCK_BYTES: {
      // block size of the host sink; the pinned two-slot ring is allocated HERE (page-locking gigabytes takes seconds when eight
      // processes of a box do it at once: a caller does it once, outside its timed region)
      if (value < (1 << 12)) return fail(h, "sink block too small");
      CK(cudaSetDevice(h->device));
      h->sink_group_bytes = (size_t)value;
      if (h->sink_host_cap < (size_t)value) {
        for (int i = 0; i < 2; ++i) { if (h->sink_host[i]) cudaFreeHost(h->sink_host[i]); h->sink_host[i] = nullptr; }
        h->sink_host_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaHostAlloc(&h->sink_host[i], (size_t)value, cudaHostAllocDefault));
        h->sink_host_cap = (size_t)value;
      }
      return 0;
    }
    case LOWDIN_IT_OPT_OVERLAP_EXCHANGE:
      if (value < 0 || value > 2) return fail(h, "overlap option must be 0, 1 or 2");
      h->overlap_exchange = (int)value; return 0;
    case LOWDIN_IT_OPT_STORED_FUSED:
      h->stored_fused = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_Q1_DEBUG:
      h->q1_dbg = (int)value; return 0;
    case LOWDIN_IT_OPT_Q3_RED:
      h->q3_red = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_ASYNC_PUSH:
      h->async_push = value ? 1 : 0; return 0;
    case LOWDIN_IT_OPT_STAGING_BYTES:
      if (value < (1 << 16)) return fail(h, "staging buffer too small");
      if (h->up_a >= 0) return fail(h, "staging size cannot change during an upload");
      h->staging_bytes = (size_t)value; h->st[0].release(); h->st[1].release(); return 0;
    case LOWDIN_IT_OPT_BENCH_GEN:
      if (value != 1 && value != 2) return fail(h, "generator kind must be 1 or 2");
      h->bench_gen = (int)value; return 0;
    default: return fail(h, "unknown option");
  }
}

int lowdin_it_set_profiling(lowdin_it_handle h, int on) {
  if (!h) return 1;
  h->prof_on = on != 0;
  for (int c = 0; c < 8; ++c) h->prof_ms[c] = h->prof_cnt[c] = h->prof_work[c] = 0;
  return 0;
}

int lowdin_it_kernel_stats(lowdin_it_handle h, double ms[8], double launches[8], double work[8]) {
  if (!h) return 1;
  for (int c = 0; c < 8; ++c) { ms[c] = h->prof_ms[c]; launches[c] = h->prof_cnt[c]; work[c] = h->prof_work[c]; }
  return 0;
}

int lowdin_it_timers(lowdin_it_handle h, double out[8]) {
  if (!h || !out) return 1;
  memcpy(out, h->timers, sizeof(double) * 8);
  return 0;
}

// ---- transformer-D compatible entry points --------------------------------------------------
static int dcompat_ctx() {
  g_create_error.clear();
  int dev = 0;
  if (const char *e = getenv("LOWDIN_IT_DEVICE")) dev = atoi(e);
  else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev < 0 || dev >= 64) return fail(nullptr, "device index out of range");
  std::lock_guard<std::mutex> lk(g_attr_mutex);
  if (!g_dcompat_dev[dev] && lowdin_it_create(dev, &g_dcompat_dev[dev])) return 1;
  g_dcompat = g_dcompat_dev[dev];
  g_dcompat->err.clear();
  return 0;
}

int lowdin_it_transform_all(const double *coeff, double *ints, int nao) {
  if (dcompat_ctx()) return 1;
  lowdin_it_handle h = g_dcompat;
  if (!coeff || !ints || nao <= 0) return fail(h, "bad arguments");
  if (lowdin_it_set_species(h, 0, nao, coeff, nao, nao)) return 1;
  const int64_t M = npairs(nao), cnt = M * (M + 1) / 2;
  CK(h->dtmp.ensure(cnt * sizeof(double)));
  CK(cudaMemcpyAsync(h->dtmp.p, ints, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  AoSet &S = h->ao[0][0];
  CK(S.data.ensure(cnt * sizeof(double)));
  repack_d_intra_kernel<<<(unsigned)ceil_div(M * M, 256), 256, 0, h->stream>>>(S.data.as<double>(), h->dtmp.as<double>(),
                                                                             h->sp[0].pi.as<int32_t>(), h->sp[0].pj.as<int32_t>(), M);
  CK(cudaGetLastError());
  S.src = AoSource{SRC_SYM_PACKED, S.data.as<double>(), M, 0, 0, 0};
  S.valid = true;
  const int win[8] = {1, nao, 1, nao, 1, nao, 1, nao};
  // E convention with a negative tolerance keeps every (ij,kl), ij and kl in ijmap/klmap order,
  // which for the full window is exactly D's lower-triangular pair numbering.
  if (lowdin_it_transform(h, 0, 0, win, LOWDIN_IT_CONV_E, 0, -1.0)) return 1;
  if (h->count != M * M) return fail(h, "internal: full transform did not produce M*M values");
  pack_d_lower_kernel<<<(unsigned)ceil_div(M * M, 256), 256, 0, h->stream>>>(h->r_v.as<double>(), h->dtmp.as<double>(), M);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ints, h->dtmp.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lowdin_it_transform_inter_all(const double *coeff, const double *ocoeff, double *ints, int nao, int onao) {
  if (dcompat_ctx()) return 1;
  lowdin_it_handle h = g_dcompat;
  if (!coeff || !ocoeff || !ints || nao <= 0 || onao <= 0) return fail(h, "bad arguments");
  if (lowdin_it_set_species(h, 0, nao, coeff, nao, nao)) return 1;
  if (lowdin_it_set_species(h, 1, onao, ocoeff, onao, onao)) return 1;
  const int64_t Ma = npairs(nao), Mb = npairs(onao), cnt = Ma * Mb;
  CK(h->dtmp.ensure(cnt * sizeof(double)));
  CK(cudaMemcpyAsync(h->dtmp.p, ints, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  AoSet &S = h->ao[0][1];
  CK(S.data.ensure(cnt * sizeof(double)));
  repack_d_inter_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, h->stream>>>(S.data.as<double>(), h->dtmp.as<double>(),
                                                                           h->sp[0].pi.as<int32_t>(), h->sp[0].pj.as<int32_t>(),
                                                                           h->sp[1].pi.as<int32_t>(), h->sp[1].pj.as<int32_t>(), Ma, Mb);
  CK(cudaGetLastError());
  S.src = AoSource{SRC_RECT, S.data.as<double>(), Ma, Ma, Mb, 0};
  S.valid = true;
  const int win[8] = {1, nao, 1, nao, 1, onao, 1, onao};
  if (lowdin_it_transform(h, 0, 1, win, LOWDIN_IT_CONV_E, 0, -1.0)) return 1;
  if (h->count != cnt) return fail(h, "internal: full inter transform did not produce Ma*Mb values");
  CK(cudaMemcpyAsync(ints, h->r_v.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- multi-GPU --------------------------------------------------------------------------------
int lowdin_it_shard_plan(int nfb, const int *fbeg, int64_t chunk_base, int64_t chunk_width, int nranks, int rank, int log_block, int *own,
                         int64_t *wblk, int64_t *loc_lo, int64_t *count) {
  if (nfb < 1 || !fbeg || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || log_block < 0 || log_block > 20 || chunk_base < 0 ||
      chunk_width < 0 || !own || !wblk || !loc_lo || !count)
    return 1;
  shard_plan(nfb, fbeg, chunk_base, chunk_width, nranks, rank, log_block, own, wblk, loc_lo, count);
  return 0;
}
int lowdin_it_exchange_is_dma(lowdin_it_handle h) { return (h && h->nranks > 1 && h->node && h->exchange_dma && !h->lgroup) ? 1 : 0; }
int lowdin_it_occ_batch_model(int n_first, int q_max, int nranks, int64_t slots_per_first, int n_first2, int nao2, int nao1, int64_t nslabs,
                              int64_t npairs1, double avail_bytes, int stored) {
  return occ_batch_model(n_first, q_max, std::max(nranks, 1), (double)slots_per_first, n_first2, nao2, nao1, (double)nslabs, (double)npairs1, avail_bytes,
                         stored != 0);
}
int lowdin_it_slab_owner(int64_t slab, int nranks, int log_block) { return nranks > 1 ? slab_owner(slab, log_block, nranks) : 0; }
int64_t lowdin_it_slab_local(int64_t slab, int nranks, int log_block) { return nranks > 1 ? slab_local(slab, log_block, nranks) : slab; }
int64_t lowdin_it_slab_global(int64_t local, int nranks, int rank, int log_block) { return slab_global(local, log_block, nranks, rank); }
int64_t lowdin_it_exchanged_offset(int64_t row, int64_t slab, int64_t chunk_base, int64_t wblk, int64_t rows, int nranks, int log_block) {
  if (nranks <= 1) return row * wblk + (slab - chunk_base);
  const int r = slab_owner(slab, log_block, nranks);
  return ((int64_t)r * rows + row) * wblk + (slab_local(slab, log_block, nranks) - slabs_owned_below(chunk_base, log_block, nranks, r));
}

int lowdin_it_comm_unique_id(char id[128]) {
  std::string err;
  if (!load_nccl(err)) return fail(nullptr, err);
  int rc = g_nccl.GetUniqueId(id);
  if (rc) return fail(nullptr, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  return 0;
}

namespace {
uint64_t fnv1a(const void *p, size_t n) {
  uint64_t x = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { x ^= ((const unsigned char *)p)[i]; x *= 1099511628211ull; }
  return x;
}
// Node link of an NCCL communicator whose ranks share one host (NodeLink above).  Collective; never an error: when any step fails
// on any rank (different hosts, no shared /dev/shm, IPC refused) every rank keeps the NCCL exchange.
void setup_node_link(lowdin_it_handle h, const char id[128]) {
  const int G = h->nranks, r = h->rank;
  char host[256] = {0};
  gethostname(host, 255);
  const int64_t hh = (int64_t)(fnv1a(host, strlen(host)) & 0x3fffffffffffffffull);
  int64_t mn = hh, mx = -hh;
  if (agree_min(h, &mn) || agree_min(h, &mx) || mn != -mx) return;
  char name[64];
  snprintf(name, sizeof name, "/lowdin_it_%016llx", (unsigned long long)fnv1a(id, 128));
  NodeShm *S = nullptr;
  int64_t ok = 1;
  if (r == 0) {
    shm_unlink(name);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, sizeof(NodeShm)) != 0) ok = 0;
    if (ok) { void *m = mmap(nullptr, sizeof(NodeShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); if (m == MAP_FAILED) ok = 0; else { S = (NodeShm *)m; memset(m, 0, sizeof(NodeShm)); } }
    if (fd >= 0) close(fd);
  }
  if (agree_min(h, &ok)) return;  // also the barrier behind which the segment exists
  if (ok && r != 0) {
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) ok = 0;
    else { void *m = mmap(nullptr, sizeof(NodeShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); if (m == MAP_FAILED) ok = 0; else S = (NodeShm *)m; close(fd); }
  }
  if (agree_min(h, &ok)) ok = 0;
  if (r == 0) shm_unlink(name);  // the mappings keep the segment alive
  if (!ok) { if (S) munmap(S, sizeof(NodeShm)); return; }
  std::unique_ptr<NodeLink> N(new NodeLink);
  N->shm = S; N->n = G;
  int64_t ok2 = 1;
  for (int b = 0; b < 2 && ok2; ++b) {
    if (cudaEventCreateWithFlags(&N->ready[b], cudaEventDisableTiming | cudaEventInterprocess) != cudaSuccess ||
        cudaEventCreateWithFlags(&N->done[b], cudaEventDisableTiming | cudaEventInterprocess) != cudaSuccess ||
        cudaIpcGetEventHandle(&S->r[r].ready[b], N->ready[b]) != cudaSuccess || cudaIpcGetEventHandle(&S->r[r].done[b], N->done[b]) != cudaSuccess)
      ok2 = 0;
  }
  if (agree_min(h, &ok2)) ok2 = 0;  // every rank has published its event handles (or none goes on)
  for (int g = 0; g < G && ok2; ++g) {
    if (g == r) continue;
    for (int b = 0; b < 2 && ok2; ++b)
      if (cudaIpcOpenEventHandle(&N->peer_ready[g][b], S->r[g].ready[b]) != cudaSuccess ||
          cudaIpcOpenEventHandle(&N->peer_done[g][b], S->r[g].done[b]) != cudaSuccess)
        ok2 = 0;
  }
  if (agree_min(h, &ok2)) ok2 = 0;
  cudaGetLastError();  // a refused IPC call must not poison the next CK()
  if (ok2) h->node = std::move(N);
}
}  // namespace

int lowdin_it_comm_init(lowdin_it_handle h, int rank, int nranks, const char id[128]) {
  if (!h) return 1;
  if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return fail(h, "bad rank / nranks (1..16 ranks)");
  CK(cudaSetDevice(h->device));
  if (nranks == 1) { h->rank = 0; h->nranks = 1; return 0; }
  std::string err;
  if (!load_nccl(err)) return fail(h, err);
  Id128 uid; memcpy(uid.b, id, 128);
  int rc = g_nccl.CommInitRank(&h->comm, nranks, uid, rank);
  if (rc) return fail(h, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
  h->rank = rank; h->nranks = nranks;
  setup_node_link(h, id);
  return 0;
}

// First QUARTER only: out[f][z][mu] = sum_nu AO(slab0+z; mu nu) C(nu, f_first+f) for nf MO indices starting at orbital f_first (1-based)
// and nslabs AO-pair slabs, through whichever first-quarter kernel the AO set's storage selects (expansion + DMMA GEMM, fused
// generator, list-driven scatter).  With C = identity this is the AO tensor itself: the bit-exact test of the index work.
int lowdin_it_debug_first_quarter(lowdin_it_handle h, int a, int b, int f_first, int nf, int64_t slab0, int nslabs, double *out) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->sp[a].n || !h->sp[b].n || !h->ao[a][b].valid) return fail(h, "debug_first_quarter: AO set not available");
  const Species &A = h->sp[a];
  if (f_first < 1 || nf < 1 || f_first + nf - 1 > A.ncols || !out) return fail(h, "debug_first_quarter: bad orbital range");
  Plan pl;
  pl.a = a; pl.b = b; pl.intra = (a == b); pl.src = h->ao[a][b].src; pl.nslabs1 = h->sp[b].M;
  pl.h1.nc = A.n; pl.h1.C = A.C.as<double>(); pl.h1.Cs = A.Cs.as<double>(); pl.h1.ldc = A.ldc; pl.h1.lf = f_first; pl.h1.nf = nf;
  if (slab0 < 0 || nslabs < 1 || slab0 + nslabs > pl.nslabs1) return fail(h, "debug_first_quarter: slab range out of bounds");
  PassTables pt;
  pt.f0 = 0; pt.nfb = nf;
  const int nc = A.n;
  const int64_t ldx = roundup2(nc), ldt = roundup2(nc);
  if (pl.src.kind == SRC_LIST && list_first_quarter(h, pl, pt)) return 1;
  const int64_t B = first_half_batch(h, pl, nf, nslabs);
  if (x_per_slab(h, pl)) CK(h->X.ensure((size_t)B * x_per_slab(h, pl) * sizeof(double)));
  CK(h->T1t.ensure((size_t)B * nf * ldt * sizeof(double)));
  for (int64_t s = 0; s < nslabs; s += B) {
    const int64_t bc = std::min<int64_t>(B, nslabs - s);
    if (first_quarter_batch(h, pl, pt, slab0 + s, bc, ldx, ldt)) return 1;
    for (int f = 0; f < nf; ++f)  // T1t[f][z][mu] of the batch -> out[f][s+z][mu]
      CK(cudaMemcpy2DAsync(out + ((size_t)f * nslabs + s) * nc, (size_t)nc * 8, h->T1t.as<double>() + (size_t)f * bc * ldt, (size_t)ldt * 8,
                           (size_t)nc * 8, (size_t)bc, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int lowdin_it_comm_init_local(lowdin_it_handle *handles, int nranks) {
  if (!handles || nranks < 1 || nranks > 16) return fail(nullptr, "comm_init_local: 1..16 handles required");
  for (int r = 0; r < nranks; ++r) if (!handles[r]) return fail(nullptr, "comm_init_local: null handle");
  if (nranks == 1) { handles[0]->rank = 0; handles[0]->nranks = 1; handles[0]->lgroup.reset(); return 0; }
  auto L = std::make_shared<LocalGroup>();
  L->n = nranks;
  for (int r = 0; r < nranks; ++r) {
    lowdin_it_handle h = handles[r];
    if (h->comm) return fail(h, "comm_init_local: the handle already belongs to an NCCL communicator");
    CK(cudaSetDevice(h->device));
    L->device[r] = h->device;
    CK(cudaEventCreateWithFlags(&L->ready[r], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&L->done[r], cudaEventDisableTiming));
    for (int o = 0; o < r; ++o)
      if (handles[o]->device != h->device) {  // direct peer access where the topology has it (the copies work either way)
        int can = 0;
        cudaDeviceCanAccessPeer(&can, h->device, handles[o]->device);
        if (can) { cudaDeviceEnablePeerAccess(handles[o]->device, 0); cudaGetLastError(); }
        cudaSetDevice(handles[o]->device);
        cudaDeviceCanAccessPeer(&can, handles[o]->device, h->device);
        if (can) { cudaDeviceEnablePeerAccess(h->device, 0); cudaGetLastError(); }
        cudaSetDevice(h->device);
      }
  }
  for (int r = 0; r < nranks; ++r) { handles[r]->rank = r; handles[r]->nranks = nranks; handles[r]->lgroup = L; }
  return 0;
}

// ---- single-process multi-GPU: the calls a one-process host (the reference's is one) makes for a whole group ----------------
}  // extern "C"
namespace {
// Merge the ranks' lists into the reference's loop order: window pair k (convention order) is one segment of its owner's list.
template <class Copy>
int merge_group_lists(lowdin_it_handle *hs, int n, Copy copy_segment) {
  std::vector<size_t> next(n, 0), off(n, 0);
  int64_t out = 0;
  for (;;) {
    int best = -1;
    int64_t kbest = 0;
    for (int r = 0; r < n; ++r)
      if (next[r] < hs[r]->res_pair.size() && (best < 0 || hs[r]->res_pair[next[r]] < kbest)) { best = r; kbest = hs[r]->res_pair[next[r]]; }
    if (best < 0) break;
    const int64_t len = hs[best]->res_seg[next[best]];
    copy_segment(best, (int64_t)off[best], out, len);
    off[best] += len; out += len; ++next[best];
  }
  return 0;
}
}  // namespace
extern "C" {

int lowdin_it_group_transform(lowdin_it_handle *hs, int n, int a, int b, const int win[8], int conv, int symmetric, double drop_tol) {
  if (!hs || n < 1) return 1;
  if (n == 1) return lowdin_it_transform(hs[0], a, b, win, conv, symmetric, drop_tol);
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < n; ++r) th.emplace_back([&, r] { rc[r] = lowdin_it_transform(hs[r], a, b, win, conv, symmetric, drop_tol); });
  for (auto &t : th) t.join();
  for (int r = 0; r < n; ++r) if (rc[r]) { if (r) hs[0]->err = hs[r]->err; return 1; }
  return 0;
}

int lowdin_it_group_result_count(lowdin_it_handle *hs, int n, int64_t *count) {
  if (!hs || n < 1 || !count) return 1;
  *count = 0;
  for (int r = 0; r < n; ++r) *count += hs[r]->count;
  return 0;
}

int lowdin_it_group_download_pairs(lowdin_it_handle *hs, int n, int64_t *ij, int64_t *kl, double *v) {
  if (!hs || n < 1) return 1;
  if (n == 1) return lowdin_it_download_pairs(hs[0], ij, kl, v);
  std::vector<std::vector<int64_t>> tij(n), tkl(n);
  std::vector<std::vector<double>> tv(n);
  for (int r = 0; r < n; ++r) {
    const size_t c = std::max<size_t>((size_t)hs[r]->count, 1);
    tij[r].resize(c); tkl[r].resize(c); tv[r].resize(c);
    if (lowdin_it_download_pairs(hs[r], tij[r].data(), tkl[r].data(), tv[r].data())) { if (r) hs[0]->err = hs[r]->err; return 1; }
  }
  return merge_group_lists(hs, n, [&](int r, int64_t src, int64_t dst, int64_t len) {
    memcpy(ij + dst, tij[r].data() + src, len * sizeof(int64_t));
    memcpy(kl + dst, tkl[r].data() + src, len * sizeof(int64_t));
    memcpy(v + dst, tv[r].data() + src, len * sizeof(double));
  });
}

int lowdin_it_group_download_quads(lowdin_it_handle *hs, int n, int32_t *p, int32_t *q, int32_t *r_, int32_t *s_, double *v) {
  if (!hs || n < 1) return 1;
  if (n == 1) return lowdin_it_download_quads(hs[0], p, q, r_, s_, v);
  std::vector<std::vector<int32_t>> t0(n), t1(n), t2(n), t3(n);
  std::vector<std::vector<double>> tv(n);
  for (int r = 0; r < n; ++r) {
    const size_t c = std::max<size_t>((size_t)hs[r]->count, 1);
    t0[r].resize(c); t1[r].resize(c); t2[r].resize(c); t3[r].resize(c); tv[r].resize(c);
    if (lowdin_it_download_quads(hs[r], t0[r].data(), t1[r].data(), t2[r].data(), t3[r].data(), tv[r].data())) { if (r) hs[0]->err = hs[r]->err; return 1; }
  }
  return merge_group_lists(hs, n, [&](int r, int64_t src, int64_t dst, int64_t len) {
    memcpy(p + dst, t0[r].data() + src, len * sizeof(int32_t));
    memcpy(q + dst, t1[r].data() + src, len * sizeof(int32_t));
    memcpy(r_ + dst, t2[r].data() + src, len * sizeof(int32_t));
    memcpy(s_ + dst, t3[r].data() + src, len * sizeof(int32_t));
    memcpy(v + dst, tv[r].data() + src, len * sizeof(double));
  });
}

// ---- stand-alone kernel timing / debug --------------------------------------------------------
// First half only (E.f90:1043-1132) of slabs [slab0, slab0 + nslabs): out[k][z] = half-transformed value of the k-th window pair
// (convention order: ijmap order for E) and slab slab0 + z, with E's |t| <= drop_tol -> 0.  Parity checks of the first half at
// sizes where no CPU can run the whole transform (bench.py compares a sample of slabs with the oracle at N_bf = 1500).
int lowdin_it_debug_first_half(lowdin_it_handle h, int a, int b, const int win[8], int conv, double drop_tol, int64_t slab0, int nslabs,
                               double *out, int64_t *npairs_out) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  Plan pl;
  if (build_plan(h, a, b, win, conv, 0, pl)) return 1;
  if (slab0 < 0 || nslabs < 1 || slab0 + nslabs > pl.nslabs1) return fail(h, "debug_first_half: slab range out of bounds");
  PassTables pt;
  build_pass(pl, 0, std::max(pl.h1.nf, 1), pt);
  if (npairs_out) *npairs_out = pt.nslots;
  if (pt.nslots == 0 || !out) return 0;
  if (upload_i32(h, h->tab, pt.table)) return 1;
  CK(h->H.ensure((size_t)pt.nslots * nslabs * sizeof(double)));
  CK(cudaMemsetAsync(h->H.p, 0, (size_t)pt.nslots * nslabs * sizeof(double), h->stream));
  if (first_half(h, pl, pt, slab0, nslabs, h->H.as<double>(), nslabs, 0, conv == LOWDIN_IT_CONV_E ? drop_tol : -1.0)) return 1;
  std::vector<double> tmp((size_t)pt.nslots * nslabs);
  CK(cudaMemcpyAsync(tmp.data(), h->H.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < pt.nslots; ++k) memcpy(out + (size_t)k * nslabs, tmp.data() + (size_t)pt.order[k] * nslabs, (size_t)nslabs * sizeof(double));
  return 0;
}

int lowdin_it_kernel_bench(lowdin_it_handle h, int kind, int64_t m, int64_t n, int64_t k, int iters, double *ms_per_launch, double *check) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  if (iters < 1) iters = 1;
  cudaEvent_t e0 = h->ev[0], e1 = h->ev[1];
  float ms = 0;
  if (kind == 0) {  // slab expansion of n slabs of an m-function hash tensor
    const int nc = (int)m; const int64_t ldx = roundup2(nc);
    CK(h->X.ensure((size_t)n * nc * ldx * sizeof(double)));
    AoSource src{(int)k /* source kind */, nullptr, npairs(nc), npairs(nc), npairs(nc), 12345, 1};
    if (src.kind == SRC_SYM_PACKED || src.kind == SRC_RECT) {
      const int64_t M = npairs(nc);
      const size_t cnt = (src.kind == SRC_SYM_PACKED) ? (size_t)(M * (M + 1) / 2) : (size_t)(n * M);
      CK(h->H.ensure(cnt * sizeof(double)));
      CK(cudaMemsetAsync(h->H.p, 0, cnt * sizeof(double), h->stream));
      src.data = h->H.as<double>();
    }
    if (launch_expand(h, src, 0, n, nc, 0, nc, 0, nc, 0, (int)ldx, h->X.as<double>())) return 1;
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < iters; ++i)
      if (launch_expand(h, src, 0, n, nc, 0, nc, 0, nc, 0, (int)ldx, h->X.as<double>())) return 1;
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (check) CK(cudaMemcpy(check, h->X.p, sizeof(double), cudaMemcpyDeviceToHost));
  } else if (kind == 2) {  // fused slab generation + first quarter: m = basis size, n = window columns, k = slabs per launch
    const int nc = (int)m, nfb = (int)n, bc = (int)k;
    const int64_t ldc = roundup2(nc), ldt = roundup2(nc);
    CK(h->X.ensure((size_t)nfb * ldc * sizeof(double)));
    CK(h->T1t.ensure((size_t)bc * nfb * ldt * sizeof(double)));
    CK(cudaMemsetAsync(h->X.p, 0x3f, (size_t)nfb * ldc * sizeof(double), h->stream));
    AoSource src{SRC_HASH_SYM, nullptr, npairs(nc), 0, 0, 12345, h->bench_gen};
    if (launch_q1_gen(h, src, 0, bc, nc, h->X.as<double>(), h->X.as<double>(), ldc, nfb, h->T1t.as<double>(), ldt)) return 1;
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < iters; ++i)
      if (launch_q1_gen(h, src, 0, bc, nc, h->X.as<double>(), h->X.as<double>(), ldc, nfb, h->T1t.as<double>(), ldt)) return 1;
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (check) CK(cudaMemcpy(check, h->T1t.p, sizeof(double), cudaMemcpyDeviceToHost));
  } else if (kind == 3) {  // probe: TMA GEMM with ONE warp per scheduler (64x128 tile, 4 warps) -- how far a single warp feeds the DMMA pipe
    const int64_t lda = roundup2(k);
    CK(h->X.ensure((size_t)m * lda * sizeof(double)));
    CK(h->T1t.ensure((size_t)n * lda * sizeof(double)));
    CK(h->OUT.ensure((size_t)m * n * sizeof(double)));
    CK(cudaMemsetAsync(h->X.p, 0x3f, (size_t)m * lda * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->T1t.p, 0x3f, (size_t)n * lda * sizeof(double), h->stream));
    GemmArgs g{h->X.as<double>(), h->T1t.as<double>(), (int)m, (int)n, (int)k, lda, lda, 0, 0};
    EpiPlain epi{h->OUT.as<double>(), n, 0};
    CK((launch_gemm_tma_cfg<64, 128, 2, 2>(h, g, epi)));
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < iters; ++i) CK((launch_gemm_tma_cfg<64, 128, 2, 2>(h, g, epi)));
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (check) CK(cudaMemcpy(check, h->OUT.p, sizeof(double), cudaMemcpyDeviceToHost));
  } else if (kind == 4 || kind == 5) {  // third-quarter shape: (m/1500 slots x 1500 rows) x n x k accumulated into T3[slot][n][1500]; 4 = staged RMW, 5 = red
    const int nrows = 1500;
    const int64_t lda = roundup2(k), slots = std::max<int64_t>(1, m / nrows), ldt = nrows;
    m = slots * nrows;
    CK(h->X.ensure((size_t)m * lda * sizeof(double)));
    CK(h->T1t.ensure((size_t)n * lda * sizeof(double)));
    CK(h->OUT.ensure((size_t)slots * n * ldt * sizeof(double)));
    CK(cudaMemsetAsync(h->X.p, 0x3f, (size_t)m * lda * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->T1t.p, 0x3f, (size_t)n * lda * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->OUT.p, 0, (size_t)slots * n * ldt * sizeof(double), h->stream));
    GemmArgs g{h->X.as<double>(), h->T1t.as<double>(), (int)m, (int)n, (int)k, lda, lda, 0, 0};
    auto run = [&]() -> cudaError_t {
      return kind == 5 ? launch_gemm(h, g, EpiAccRed{h->OUT.as<double>(), nrows, 0, (int)n, ldt})
                       : launch_gemm(h, g, EpiAccT{h->OUT.as<double>(), nrows, 0, (int)n, ldt});
    };
    CK(run());
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < iters; ++i) CK(run());
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (check) CK(cudaMemcpy(check, h->OUT.p, sizeof(double), cudaMemcpyDeviceToHost));
  } else {  // DGEMM m x n x k on generated operands
    const int64_t lda = roundup2(k);
    CK(h->X.ensure((size_t)m * lda * sizeof(double)));
    CK(h->T1t.ensure((size_t)n * lda * sizeof(double)));
    CK(h->OUT.ensure((size_t)m * n * sizeof(double)));
    CK(cudaMemsetAsync(h->X.p, 0x3f, (size_t)m * lda * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->T1t.p, 0x3f, (size_t)n * lda * sizeof(double), h->stream));
    GemmArgs g{h->X.as<double>(), h->T1t.as<double>(), (int)m, (int)n, (int)k, lda, lda, 0, 0};
    EpiPlain epi{h->OUT.as<double>(), n, 0};
    CK(launch_gemm(h, g, epi));
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < iters; ++i) CK(launch_gemm(h, g, epi));
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (check) CK(cudaMemcpy(check, h->OUT.p, sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (ms_per_launch) *ms_per_launch = ms / iters;
  return 0;
}

// C[m][n] = sum_k A[m][k] B[n][k] on host arrays (row-major, K contiguous): parity tests of the DMMA kernel alone.
int lowdin_it_debug_gemm(lowdin_it_handle h, const double *A, const double *B, double *C, int m, int n, int k) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  const int64_t ld = roundup2(k);
  CK(h->X.ensure((size_t)m * ld * sizeof(double)));
  CK(h->T1t.ensure((size_t)n * ld * sizeof(double)));
  CK(h->OUT.ensure((size_t)m * n * sizeof(double)));
  CK(cudaMemcpy2DAsync(h->X.p, ld * 8, A, (size_t)k * 8, (size_t)k * 8, m, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->T1t.p, ld * 8, B, (size_t)k * 8, (size_t)k * 8, n, cudaMemcpyHostToDevice, h->stream));
  GemmArgs g{h->X.as<double>(), h->T1t.as<double>(), m, n, k, ld, ld, 0, 0};
  EpiPlain epi{h->OUT.as<double>(), n, 0};
  CK(launch_gemm(h, g, epi));
  CK(cudaMemcpyAsync(C, h->OUT.p, (size_t)m * n * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// Dense N x N expansion of `nb` slabs starting at slab0 of the (a,b) AO set, to host X[nb][n][n].
int lowdin_it_debug_expand(lowdin_it_handle h, int a, int b, int64_t slab0, int nb, double *X) {
  if (!h) return 1;
  if (a < 0 || a > 7 || b < 0 || b > 7 || !h->ao[a][b].valid) return fail(h, "debug_expand: AO set not available");
  if (h->ao[a][b].src.kind == SRC_LIST) return fail(h, "debug_expand: the AO set is a resident list (no dense slabs exist)");
  CK(cudaSetDevice(h->device));
  const int nc = h->sp[a].n; const int64_t ldx = roundup2(nc);
  CK(h->X.ensure((size_t)nb * nc * ldx * sizeof(double)));
  if (launch_expand(h, h->ao[a][b].src, slab0, nb, nc, 0, nc, 0, nc, 0, (int)ldx, h->X.as<double>())) return 1;
  CK(cudaMemcpy2DAsync(X, (size_t)nc * 8, h->X.p, ldx * 8, (size_t)nc * 8, (size_t)nb * nc, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // extern "C"
