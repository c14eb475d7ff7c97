// nemb200.cu -- C ABI (include/nemb200.h) and host-side EM driver of the B200-native NEM path.
//
// Host control flow restates /root/reference/ppanggolin/nem/NEM/nem_alg.c:1174-1193 (INIT_PARAM_FILE start:
// densities from the initial parameters, a "blind" beta = 0 sweep, then a sweep with beta) and :1783-1938 (NemAlgo:
// M-step from the previous posteriors, E-step, convergence on max |c - c_old| < threshold).  All arithmetic is in the
// kernels (nemb200_kernels.cuh); the host only sequences launches and reads back two words per iteration.
#include "../../include/nemb200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "nemb200_kernels.cuh"

using namespace nemb200;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(e_ == cudaErrorMemoryAllocation ? NEMB200_ERR_NOMEM : NEMB200_ERR_CUDA, "%s:%d %s: %s", \
                  __FILE__, __LINE__, #call, cudaGetErrorString(e_));                                 \
  } while (0)

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    g_num_sms = p.multiProcessorCount;
  }
  return g_num_sms;
}

// Device memory: every plan / run takes ONE block from a small per-process cache of freed blocks (cudaMalloc and
// cudaFree cost 0.1-1 ms each and cudaFree synchronises the device; a chunked partition() issues thousands of
// instances).  nemb200_release_cached_memory() returns the cache to the driver.
#include <mutex>
extern "C" int nemb200_release_cached_memory(void);
namespace {
struct CachedBlock { void* p; size_t bytes; int dev; };
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
size_t g_cache_bytes = 0;
constexpr size_t kCacheLimit = (size_t)8 << 30;

int cached_malloc(void** out, size_t bytes) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  bytes = std::max<size_t>((bytes + 255) & ~(size_t)255, 256);
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int best = -1;
    for (int i = 0; i < (int)g_cache.size(); i++)
      if (g_cache[i].dev == dev && g_cache[i].bytes >= bytes && g_cache[i].bytes <= 2 * bytes + (1 << 20) &&
          (best < 0 || g_cache[i].bytes < g_cache[best].bytes))
        best = i;
    if (best >= 0) {
      *out = g_cache[best].p;
      g_cache_bytes -= g_cache[best].bytes;
      g_cache.erase(g_cache.begin() + best);
      return 0;
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
    cudaGetLastError();
    nemb200_release_cached_memory();
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? NEMB200_ERR_NOMEM : NEMB200_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return 0;
}
size_t block_size_of(size_t bytes) { return std::max<size_t>((bytes + 255) & ~(size_t)255, 256); }
void cached_free(void* p, size_t bytes) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaPointerAttributes at;   // the block belongs to the device it was allocated on, whatever is current now
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
  else cudaGetLastError();
  bytes = block_size_of(bytes);
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (g_cache_bytes + bytes > kCacheLimit) { cudaFree(p); return; }
  g_cache.push_back({p, bytes, dev});
  g_cache_bytes += bytes;
}
// bump allocator over one block; in counting mode it only measures
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  bool counting = true;
  template <typename T> T* take(size_t n) {
    const size_t a = (off + 255) & ~(size_t)255;
    off = a + std::max<size_t>(n, 1) * sizeof(T);
    return counting ? nullptr : reinterpret_cast<T*>(base + a);
  }
};
}  // namespace

extern "C" int nemb200_release_cached_memory(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto& b : g_cache) cudaFree(b.p);
  g_cache.clear();
  g_cache_bytes = 0;
  return NEMB200_OK;
}

struct nemb200_plan {
  long long N_local = 0, N_total = 0;
  int D = 0, Ws = 0, Dpad = 0, K = 0, free_disp = 0;
  uint32_t flags = 0;
  Params P{};
  // statistics
  int* cnt = nullptr;       // [K][Dpad]
  double* fsum = nullptr;   // [K][Dpad]
  int* nk_cnt = nullptr;    // [K]
  double* nk_fz = nullptr;  // [K]
  int* cnt_local = nullptr; // [K][Dpad] this rank's running one-hot column counts (kept across iterations)
  int* hist = nullptr;      // [2K+1]  add buckets, remove buckets, fuzzy bucket
  int* offsets = nullptr;   // [2K+2]
  int* cursor = nullptr;    // [2K+1]
  int* prev = nullptr;      // [N_local] key of every row at the previous M-step (kNone before the first)
  int2* ent = nullptr;      // [N_local] bucket entries of the current M-step
  int* perm = nullptr;      // [2 * N_local] rows grouped by bucket
  double* Ltab = nullptr;   // [K][Ws * 8][16] (skd): partial sums of the log-odds per nibble
  double* Lword = nullptr;  // [K][Ws] (skd): totals per 32-column word
  double* fin_partial = nullptr;  // [K][nchunk][2]
  int* fin_tickets = nullptr;     // [K]
  // optional CUDA-event timing of k_mstep_onehot (bench.py's roofline): pairs recorded per launch, read back on demand
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> onehot_events;
  // fork/join for the fuzzy-row sums, which run beside the one-hot count kernel (its fat CTAs leave room on every SM)
  // schedule-ordered copy of the neighbour graph for the level-scheduled sweep (built on first use, keyed by the arrays)
  void* ell = nullptr;
  size_t ell_bytes = 0;
  const void* ell_key[4] = {nullptr, nullptr, nullptr, nullptr};
  long long ell_N = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint32_t* flags2 = nullptr;     // [0] max |T_new - T_old| as float bits, [1] status (P.status points here)
  // row-sharded exchange over NVLink peer memory (nemb200_peer_*)
  double* logf_full = nullptr;    // [N_total][K] log-densities of ALL rows: every rank's density kernel stores into it
  int* sig = nullptr;             // [4][kMaxPeers] epochs completed by each rank: phase 0 = M-step statistics, 1 = densities, 2 = run start
  int* peer_err = nullptr;        // [1] set when a peer wait timed out
  int** sig_ptrs = nullptr;       // [kMaxPeers] device array: the ranks' sig arrays
  int world = 1, rank = 0;
  bool attached = false;
  void* peer_base[kMaxPeers] = {nullptr};
  double* peer_logf[kMaxPeers] = {nullptr};
  PeerParams peer_params{};      // the ranks' parameter arrays (column-sliced finalize of the device-resident sharded loop)
  const int* peer_iblock[kMaxPeers] = {nullptr};
  const double* peer_dblock[kMaxPeers] = {nullptr};
  int epoch_stats = 0, epoch_dens = 0, epoch_run = 0;
  int* bcast_ticket = nullptr;    // [1] CTAs of k_peer_broadcast that have finished
  int sweep_cta_cap = 0;          // > 0: CTAs of the cooperative sweep (instances of a batch run side by side)
  bool sweep_waits = false;       // the next sweep waits for every rank's densities (nemb200_peer_wait_densities)
  int fin_chunks = 0;
  float* eps_init = nullptr;
  double* crit_partial = nullptr;
  double* crit4 = nullptr;
  int crit_blocks = 0;
  float* crit_terms = nullptr;    // per-warp regions of compacted (d_ik, g_ik) terms (reference-order accumulation)
  int* crit_counts = nullptr;     // [crit_blocks * 8] terms per region
  int* crit_offsets = nullptr;    // [crit_blocks * 8 + 1]
  float* crit_dense = nullptr;    // the regions' terms, contiguous and in order
  double* crit_lz = nullptr;      // [N_total][2] per-row (log f_i, -log z_i): terms of the float32 L and Z sums
  int* crit_nrows = nullptr;      // [1] number of rows of the last criteria call (term count of the L / Z chains)
  long long* crit_rec = nullptr;  // k_cc_*, D / G chains: per (chunk, chain) records (3 x int64 + fp64 chunk sum; flag + predicted binade)
  int* crit_recbad = nullptr;
  CcState* crit_state = nullptr;
  long long* crit_rec_lz = nullptr;   // the same for the L / Z chains (their CTA runs beside the D / G one)
  int* crit_recbad_lz = nullptr;
  CcState* crit_state_lz = nullptr;
  // run buffers of the whole-run drivers (nem_host / nem_device / batch); not used by the stepwise interface
  bool stepwise_used = false;     // work was enqueued through the stepwise entry points on a stream this library does not own
  bool with_run_buffers = false;
  bool batch_ell = false;         // batched runs: the schedule-ordered graph copy lives in the plan's block
  const void* run_ell_key[4] = {nullptr, nullptr, nullptr, nullptr};   // row-sharded device loop: graph the copy in the block was built from
  void* ell_in_block = nullptr;
  char* ll_base = nullptr;        // PEER_EXCHANGE: this rank's LL region (see LLayout); peer_ll[r] = rank r's
  char* peer_ll[kMaxPeers] = {nullptr};
  LLayout ll_lay{};
  int* spec_chg = nullptr;        // speculative sweep: change marks [2][rows], words [8]
  int* spec_words = nullptr;
  double* run_logf = nullptr;
  float *run_Ta = nullptr, *run_Tb = nullptr;
  void* block = nullptr;          // the one device allocation everything above is carved from
  size_t block_bytes = 0;
};

// lays every buffer of a plan out in `ar` (counting pass, then the real one)
static void plan_layout(nemb200_plan* pl, Arena& ar) {
  Params& P = pl->P;
  const int K = pl->K;
  const size_t N = (size_t)pl->N_local, D = (size_t)pl->D, Ws = (size_t)pl->Ws, Dpad = (size_t)pl->Dpad;
  P.pk = ar.take<float>(K); P.logpk32 = ar.take<float>(K); P.logpk64 = ar.take<double>(K);
  P.mu = ar.take<float>(K * D); P.eps = ar.take<float>(K * D); P.eps_k = ar.take<float>(K);
  P.m1 = ar.take<uint32_t>(K * Ws); P.valid = ar.take<uint32_t>(K * Ws);
  P.L = ar.take<double>(pl->free_disp ? K * Dpad : (size_t)K); P.B = ar.take<double>(K);
  P.has_half = ar.take<int>(1);
  pl->flags2 = ar.take<uint32_t>(kStWords + 8);   // kSt* state words of a run, then the sweep barrier counter of batched runs
  P.status = (int*)(pl->flags2 + kStStatus);
  // statistics that a multi-GPU driver all-reduces: one contiguous int32 block and one contiguous fp64 block
  pl->cnt = ar.take<int>(K * Dpad + K); pl->fsum = ar.take<double>(K * Dpad + K);
  pl->nk_cnt = pl->cnt + K * Dpad; pl->nk_fz = pl->fsum + K * Dpad;
  // hist and cnt_local follow directly: one memset per M-step clears [nk_cnt .. hist] (and the counts on a full recount)
  pl->hist = ar.take<int>(2 * K + 1); pl->cursor = ar.take<int>(2 * K + 1);
  pl->cnt_local = ar.take<int>(K * Dpad);
  pl->offsets = ar.take<int>(2 * K + 2);
  pl->prev = ar.take<int>(N); pl->ent = ar.take<int2>(N); pl->perm = ar.take<int>(2 * N);
  pl->Ltab = ar.take<double>(pl->free_disp ? K * Ws * 128 : (size_t)1); pl->Lword = ar.take<double>(K * Ws); pl->eps_init = ar.take<float>(K);
  pl->fin_partial = ar.take<double>((size_t)K * pl->fin_chunks * 2); pl->fin_tickets = ar.take<int>(K);
  pl->crit_partial = ar.take<double>((size_t)pl->crit_blocks * 4); pl->crit4 = ar.take<double>(8);
  pl->sig = ar.take<int>(4 * kMaxPeers); pl->peer_err = ar.take<int>(1); pl->bcast_ticket = ar.take<int>(1); pl->sig_ptrs = ar.take<int*>(kMaxPeers);
  if (pl->flags & NEMB200_PEER_EXCHANGE) {
    pl->logf_full = ar.take<double>((size_t)pl->N_total * K);
    // the device-resident loop of a sharded run (nemb200_nem_sharded): replicated posteriors and graph copy of ALL rows
    pl->run_Ta = ar.take<float>((size_t)pl->N_total * K); pl->run_Tb = ar.take<float>((size_t)pl->N_total * K);
    pl->ell_in_block = ar.take<char>((size_t)pl->N_total * (sizeof(int4) + kEllW * sizeof(int2)));
    pl->spec_chg = ar.take<int>(2 * (size_t)pl->N_total); pl->spec_words = ar.take<int>(8);
    pl->ll_lay = ll_layout(K, pl->D, pl->Dpad, pl->Ws, pl->fin_chunks, pl->free_disp);
    pl->ll_base = (char*)ar.take<uint4>((size_t)(pl->ll_lay.bytes / 16 + 1));
  }
  // the sweep and the criteria run over all rows, also on a shard's plan; one region of ceil(N/warps) rows per warp
  pl->crit_terms = ar.take<float>(((size_t)pl->N_total + (size_t)pl->crit_blocks * 8) * K * 2);
  pl->crit_counts = ar.take<int>((size_t)pl->crit_blocks * 8);
  pl->crit_offsets = ar.take<int>((size_t)pl->crit_blocks * 8 + 1);
  pl->crit_dense = ar.take<float>(((size_t)pl->N_total * K + 64) * 2);
  pl->crit_lz = ar.take<double>(((size_t)pl->N_total + 64) * 2);
  pl->crit_nrows = ar.take<int>(1);
  {
    const size_t nch = ((size_t)pl->N_total * K + kCcChunk - 1) / kCcChunk + 1;
    pl->crit_rec = ar.take<long long>(nch * 2 * 4); pl->crit_recbad = ar.take<int>(nch * 2 * 2); pl->crit_state = ar.take<CcState>(1);
    const size_t nch_lz = ((size_t)pl->N_total + kCcChunk - 1) / kCcChunk + 1;
    pl->crit_rec_lz = ar.take<long long>(nch_lz * 2 * 4); pl->crit_recbad_lz = ar.take<int>(nch_lz * 2 * 2); pl->crit_state_lz = ar.take<CcState>(1);
  }
  if (pl->with_run_buffers) {
    pl->run_logf = ar.take<double>(N * K); pl->run_Ta = ar.take<float>(N * K); pl->run_Tb = ar.take<float>(N * K);
    if (pl->batch_ell) {
      pl->ell_in_block = ar.take<char>(N * (sizeof(int4) + kEllW * sizeof(int2)));
      pl->spec_chg = ar.take<int>(2 * N); pl->spec_words = ar.take<int>(8);
    }
  }
}

static int plan_build(nemb200_plan** out, int64_t N_local, int64_t N_total, int64_t D, int64_t row_stride_words,
                      int32_t K, int32_t free_disp, uint32_t flags, bool with_run_buffers, bool batch_ell = false);
static bool batch_enabled_for_plans() {
  const char* e = getenv("NEMB200_NO_BATCH");
  return !(e && atoi(e) != 0);
}

extern "C" int nemb200_abi_version(void) { return NEMB200_ABI_VERSION; }
extern "C" const char* nemb200_last_error(void) { return g_err.c_str(); }
extern "C" int nemb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int nemb200_plan_create(nemb200_plan** out, int64_t N_local, int64_t N_total, int64_t D,
                                   int64_t row_stride_words, int32_t K, int32_t free_disp, uint32_t flags) {
  return plan_build(out, N_local, N_total, D, row_stride_words, K, free_disp, flags, false);
}

static int plan_build(nemb200_plan** out, int64_t N_local, int64_t N_total, int64_t D, int64_t row_stride_words,
                      int32_t K, int32_t free_disp, uint32_t flags, bool with_run_buffers, bool batch_ell) {
  if (!out || N_local < 0 || N_total < 1 || D < 1 || K < 1 || K > NEMB200_MAX_K || row_stride_words % 4 != 0 ||
      row_stride_words * 32 < D || N_total > 0x7fffffffLL)
    return fail(NEMB200_ERR_ARG, "plan_create: bad shape N_local=%lld N_total=%lld D=%lld stride=%lld K=%d",
                (long long)N_local, (long long)N_total, (long long)D, (long long)row_stride_words, K);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  nemb200_plan* pl = new nemb200_plan();
  pl->N_local = N_local; pl->N_total = N_total; pl->D = (int)D; pl->Ws = (int)row_stride_words;
  pl->Dpad = pl->Ws * 32; pl->K = K; pl->free_disp = free_disp ? 1 : 0; pl->flags = flags;
  pl->with_run_buffers = with_run_buffers;
  pl->batch_ell = batch_ell;
  Params& P = pl->P;
  P.K = K; P.D = pl->D; P.Ws = pl->Ws; P.Dpad = pl->Dpad; P.free_disp = pl->free_disp;
  pl->fin_chunks = (pl->Dpad + kFinCols - 1) / kFinCols;
  pl->crit_blocks = num_sms() * 4;
  Arena count;
  plan_layout(pl, count);
  pl->block_bytes = count.off + 256;
  int rc = cached_malloc(&pl->block, pl->block_bytes);
  if (rc) { delete pl; return rc; }
  Arena real;
  real.base = (char*)pl->block; real.counting = false;
  plan_layout(pl, real);
  // whole-run plans of the device-resident loop clear their words in k_b_start; the stepwise / sharded interfaces need the
  // exchange words and the finalize tickets cleared here (blocking memsets: 10-20 us each, kept off the per-run path)
  if (!with_run_buffers || (flags & NEMB200_PEER_EXCHANGE) || !batch_enabled_for_plans()) {
    if (cudaMemset(pl->sig, 0, 4 * kMaxPeers * sizeof(int)) != cudaSuccess || cudaMemset(pl->peer_err, 0, sizeof(int)) != cudaSuccess || cudaMemset(pl->bcast_ticket, 0, sizeof(int)) != cudaSuccess ||
        cudaMemset(pl->fin_tickets, 0, K * sizeof(int)) != cudaSuccess) {
      nemb200_plan_destroy(pl);
      return fail(NEMB200_ERR_CUDA, "plan_create: memset");
    }
  }
  if (pl->ll_base != nullptr && cudaMemset(pl->ll_base, 0, pl->ll_lay.bytes) != cudaSuccess) {   // no word carries a valid epoch yet
    nemb200_plan_destroy(pl);
    return fail(NEMB200_ERR_CUDA, "plan_create: memset of the exchange words");
  }
  *out = pl;
  return NEMB200_OK;
}

extern "C" void nemb200_plan_destroy(nemb200_plan* pl) {
  if (!pl) return;
  for (auto& pr : pl->onehot_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  if (pl->side) cudaStreamDestroy(pl->side);
  // a block goes back to the cache only when nothing queued on the caller's streams can still touch it
  if (pl->stepwise_used && !pl->attached) cudaDeviceSynchronize();
  if (pl->attached) {
    cudaDeviceSynchronize();
    for (int r = 0; r < pl->world; r++)
      if (r != pl->rank && pl->peer_base[r]) cudaIpcCloseMemHandle(pl->peer_base[r]);
  }
  if (pl->ell) cached_free(pl->ell, pl->ell_bytes);
  cached_free(pl->block, pl->block_bytes);
  delete pl;
}

extern "C" int nemb200_init_params(nemb200_plan* pl, void* stream_) {
  if (!pl) return fail(NEMB200_ERR_ARG, "init_params: null plan");
  cudaStream_t st = (cudaStream_t)stream_;
  const int K = pl->K;
  // partition.py:65-85: the values are written as text and read back with fscanf("%f") -> float32
  char buf[64];
  snprintf(buf, sizeof buf, "%.12g", 1.0 / (double)K);
  const float base = strtof(buf, nullptr);
  const double step = 0.5 / std::ceil((double)K / 2.0);
  const double pich = (K == 2) ? 0.1 : 0.0;
  std::vector<float> e(K);
  for (int k = 1; k <= K; k++)
    e[k - 1] = ((double)k <= (double)K / 2.0) ? (float)(step * k - pich) : (float)(step * (K - k + 1) - pich);
  CK(cudaMemcpyAsync(pl->eps_init, e.data(), K * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // e is a stack vector
  k_init_params<<<K, 256, 0, st>>>(pl->P, base, pl->eps_init);
  CK(cudaGetLastError());
  // a new run: no row is counted yet (kNone == -2 is the byte pattern 0xFE repeated)
  CK(cudaMemsetAsync(pl->prev, 0xFE, (size_t)std::max<long long>(pl->N_local, 1) * sizeof(int), st));
  CK(cudaMemsetAsync(pl->cnt_local, 0, (size_t)K * pl->Dpad * sizeof(int), st));
  CK(cudaMemsetAsync(pl->cnt, 0, (size_t)K * pl->Dpad * sizeof(int), st));
  if (pl->free_disp) {
    const long long n = (long long)K * pl->Ws;
    k_ltab<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pl->P.L, K, pl->Ws, pl->Ltab, pl->Lword);
    CK(cudaGetLastError());
  }
  return NEMB200_OK;
}

template <int KB>
static int launch_classify(nemb200_plan* pl, const float* T, cudaStream_t st) {
  const long long N = pl->N_local;
  if (N > 0) {
    k_classify<KB><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(T, N, pl->K, pl->prev, pl->ent, pl->hist, pl->nk_cnt, pl->nk_fz,
                                                                 (pl->flags & NEMB200_FULL_RECOUNT) ? 1 : 0);
    CK(cudaGetLastError());
  }
  return 0;
}
template <int KB>
static int launch_fuzzy(nemb200_plan* pl, const uint32_t* bits, const float* T, cudaStream_t st) {
  const char* mma_env = getenv("NEMB200_FUZZY_MMA");
  if (KB >= 8 && !(mma_env && atoi(mma_env) == 0)) {   // more than 4 classes: fp64 tensor cores (k_mstep_fuzzy_mma)
    constexpr int KT = KB / 8 < 1 ? 1 : KB / 8;
    const int col_tiles = (pl->Ws + 7) / 8;
    k_mstep_fuzzy_mma<KT><<<dim3(col_tiles, 128), 256, 0, st>>>(bits, pl->Ws, pl->K, pl->D, pl->offsets, pl->perm, T, pl->fsum, pl->Dpad);
    CK(cudaGetLastError());
    return 0;
  }
  const int cols_per_cta = 256 * FuzzyCfg<KB>::WPT;
  const int col_tiles = (pl->Dpad + cols_per_cta - 1) / cols_per_cta;
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mstep_fuzzy<KB>, 256, 0));
  const long long slots = (long long)num_sms() * std::max(per_sm, 1);
  const long long slices = std::min<long long>(std::max<long long>(slots / col_tiles, 1), 65535);  // one resident wave
  dim3 grid(col_tiles, (unsigned)slices);
  k_mstep_fuzzy<KB><<<grid, 256, 0, st>>>(bits, pl->Ws, pl->K, pl->D, pl->offsets, pl->perm, T, pl->fsum, pl->Dpad);
  CK(cudaGetLastError());
  return 0;
}

#define DISPATCH_KB(K, FN, ...)                    \
  ((K) <= 3 ? FN<3>(__VA_ARGS__)                   \
   : (K) <= 4 ? FN<4>(__VA_ARGS__)                 \
   : (K) <= 8 ? FN<8>(__VA_ARGS__)                 \
   : (K) <= 16 ? FN<16>(__VA_ARGS__)               \
               : FN<32>(__VA_ARGS__))

extern "C" int nemb200_mstep_accumulate(nemb200_plan* pl, const uint32_t* bits, const float* T, void* stream_) {
  if (!pl || (pl->N_local > 0 && (!bits || !T))) return fail(NEMB200_ERR_ARG, "mstep_accumulate: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  pl->stepwise_used = true;
  const int K = pl->K;
  const long long N = pl->N_local;
  // The running one-hot counts live in the exposed block itself unless a collective library reduces that block in place
  // (row shard without peer mapping): then they are kept in cnt_local and copied out after every update.
  const bool separate = pl->N_local < pl->N_total && pl->world <= 1;
  int* cnt_acc = separate ? pl->cnt_local : pl->cnt;
  {  // one memset: class sizes, fuzzy sums, bucket histogram -- plus the counts when every one-hot row is streamed again
    const bool full = (pl->flags & NEMB200_FULL_RECOUNT) != 0;
    char* z0 = (char*)((full && !separate) ? pl->cnt : pl->nk_cnt);
    char* z1 = (full && separate) ? (char*)(pl->cnt_local + (size_t)K * pl->Dpad) : (char*)(pl->cursor + 2 * K + 1);
    CK(cudaMemsetAsync(z0, 0, (size_t)(z1 - z0), st));
  }
  int rc = DISPATCH_KB(K, launch_classify, pl, T, st);
  if (rc) return rc;
  if (N > 0) {
    k_scatter<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pl->ent, N, pl->hist, 2 * K + 1, pl->offsets, pl->cursor, pl->perm);
    CK(cudaGetLastError());
    // column tiles of equal width; row slices sized so that the whole grid is one resident wave
    const int W2 = pl->Ws / 2;
    const int ntiles = (W2 + kMsTpb - 1) / kMsTpb;
    const int tile_w = (W2 + ntiles - 1) / ntiles;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mstep_onehot, kMsTpb, 0));
    const long long slots = (long long)num_sms() * std::max(per_sm, 1);
    const long long slices = std::min<long long>(std::max<long long>(slots / ntiles, 1), 65535);  // longer slices are walked in kMaxRun pieces
    dim3 grid(ntiles, (unsigned)slices);
    static const bool fork_fuzzy = !(getenv("NEMB200_FUZZY_FORK") && atoi(getenv("NEMB200_FUZZY_FORK")) == 0);
    if (fork_fuzzy) {
      if (!pl->side) {
        CK(cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
      }
      CK(cudaEventRecord(pl->ev_fork, st));   // the index lists are complete here
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (pl->timing) {
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0, st));
    }
    // wide matrices: one fat CTA per SM, rows staged with cp.async (k_mstep_onehot_ring); NEMB200_ONEHOT_RING=0/1 forces
    const char* force = getenv("NEMB200_ONEHOT_RING");
    const int W4 = pl->Ws / 4;
    if (force ? atoi(force) != 0 : W4 >= kOhTpb / 2) {
      const int ntx = (W4 + kOhTpb - 1) / kOhTpb;
      const int tw = (W4 + ntx - 1) / ntx;
      const size_t smem = onehot_ring_smem();
      static bool smem_set[64] = {false};
      int dev = 0;
      CK(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64 || !smem_set[dev]) {
        CK(cudaFuncSetAttribute(k_mstep_onehot_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_set[dev] = true;
      }
      const long long sl = std::min<long long>(std::max<long long>(num_sms() / ntx, 1), 65535);
      k_mstep_onehot_ring<<<dim3(ntx, (unsigned)sl), kOhTpb, smem, st>>>((const uint4*)bits, W4, tw, K, pl->offsets, pl->perm,
                                                                        cnt_acc, pl->Dpad);
    } else {
      k_mstep_onehot<<<grid, kMsTpb, 0, st>>>((const uint2*)bits, W2, tile_w, K, pl->offsets, pl->perm, cnt_acc, pl->Dpad);
    }
    CK(cudaGetLastError());
    if (pl->timing) {
      CK(cudaEventRecord(e1, st));
      pl->onehot_events.emplace_back(e0, e1);
    }
    if (fork_fuzzy) {
      CK(cudaStreamWaitEvent(pl->side, pl->ev_fork, 0));
      rc = DISPATCH_KB(K, launch_fuzzy, pl, bits, T, pl->side);
      if (rc) return rc;
      CK(cudaEventRecord(pl->ev_join, pl->side));
      CK(cudaStreamWaitEvent(st, pl->ev_join, 0));
    } else {
      rc = DISPATCH_KB(K, launch_fuzzy, pl, bits, T, st);
      if (rc) return rc;
    }
  }
  // the exposed statistics (all-reduced in place by a multi-GPU driver) start from this rank's running counts
  if (separate) CK(cudaMemcpyAsync(pl->cnt, pl->cnt_local, (size_t)K * pl->Dpad * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return NEMB200_OK;
}

extern "C" int nemb200_stats_ptrs(nemb200_plan* pl, int32_t** cnt, double** fsum, int32_t** nk_cnt, double** nk_fz,
                                  int64_t* Dpad) {
  if (!pl) return fail(NEMB200_ERR_ARG, "stats_ptrs: null plan");
  if (cnt) *cnt = pl->cnt;
  if (fsum) *fsum = pl->fsum;
  if (nk_cnt) *nk_cnt = pl->nk_cnt;
  if (nk_fz) *nk_fz = pl->nk_fz;
  if (Dpad) *Dpad = pl->Dpad;
  return NEMB200_OK;
}

extern "C" int nemb200_plan_timing(nemb200_plan* pl, int32_t enable, float* onehot_ms_total, int32_t* onehot_launches) {
  if (!pl) return fail(NEMB200_ERR_ARG, "plan_timing: null plan");
  float tot = 0.0f;
  int n = 0;
  for (auto& pr : pl->onehot_events) {
    float ms = 0.0f;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { tot += ms; n++; }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  pl->onehot_events.clear();
  if (onehot_ms_total) *onehot_ms_total = tot;
  if (onehot_launches) *onehot_launches = n;
  pl->timing = enable != 0;
  return NEMB200_OK;
}

extern "C" int nemb200_stats_blocks(nemb200_plan* pl, int32_t** iblock, int64_t* ilen, double** dblock, int64_t* dlen) {
  if (!pl) return fail(NEMB200_ERR_ARG, "stats_blocks: null plan");
  const int64_t n = (int64_t)pl->K * pl->Dpad + pl->K;
  if (iblock) *iblock = pl->cnt;
  if (dblock) *dblock = pl->fsum;
  if (ilen) *ilen = n;
  if (dlen) *dlen = n;
  return NEMB200_OK;
}

extern "C" int nemb200_mstep_finalize(nemb200_plan* pl, void* stream_) {
  if (!pl) return fail(NEMB200_ERR_ARG, "mstep_finalize: null plan");
  cudaStream_t st = (cudaStream_t)stream_;
  CK(cudaMemsetAsync(pl->P.has_half, 0, sizeof(int), st));
  PeerStats ps;
  ps.n = 1; ps.iblock[0] = pl->cnt; ps.dblock[0] = pl->fsum;
  if (pl->attached) {
    // the all-reduce of the M-step statistics, fused into this kernel: publish "my statistics of this iteration are
    // complete", wait for every rank's, then read and add all ranks' blocks in rank order (peer loads over NVLink)
    const int epoch = ++pl->epoch_stats;
    k_peer_signal_wait<<<1, 32, 0, st>>>(pl->sig_ptrs, pl->sig, pl->world, 0, pl->rank, epoch, pl->peer_err);
    CK(cudaGetLastError());
    ps.n = pl->world;
    for (int r = 0; r < pl->world; r++) { ps.iblock[r] = pl->peer_iblock[r]; ps.dblock[r] = pl->peer_dblock[r]; }
  }
  k_mstep_finalize<<<dim3(pl->fin_chunks, pl->K), 256, 0, st>>>(pl->P, ps, pl->N_total, pl->fin_partial, pl->fin_tickets);
  CK(cudaGetLastError());
  if (pl->free_disp) {
    const long long n = (long long)pl->K * pl->Ws;
    k_ltab<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pl->P.L, pl->K, pl->Ws, pl->Ltab, pl->Lword);
    CK(cudaGetLastError());
  }
  return NEMB200_OK;
}

static int group_lanes(int units) {  // lanes cooperating on one row: smallest power of two >= units, capped at 32
  int g = 1;
  while (g < units && g < 32) g <<= 1;
  return g;
}

template <int KB, int KX>
static int launch_density_sk(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st) {
  constexpr int RB = KB <= 4 ? 4 : (KB <= 8 ? 2 : 1);  // rows per lane group: bounded by registers (2 * RB * KB + 4 * RB)
  const int W4 = pl->Ws / 4;
  const int G = group_lanes(W4);
  const size_t smem = (size_t)2 * pl->K * W4 * sizeof(uint4);
  const int in_smem = smem <= 100 * 1024;
  auto kern = k_density_sk<KB, KX, RB>;
  if (in_smem && smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rows_per_block = (long long)(kDensTpb / 32) * (32 / G) * RB;
  long long blocks = (pl->N_local + rows_per_block - 1) / rows_per_block;
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDensTpb, in_smem ? smem : 0));
  const long long cap = (long long)num_sms() * std::max(per_sm, 1) * 2;  // two waves of resident CTAs, grid-stride inside
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, kDensTpb, in_smem ? smem : 0, st>>>((const uint4*)bits, pl->N_local, W4, pl->K, G,
                                                          (const uint4*)pl->P.m1, (const uint4*)pl->P.valid, pl->P.L,
                                                          pl->P.B, pl->P.has_half, logf, in_smem);
  CK(cudaGetLastError());
  return 0;
}
// Shared-memory-ring variant (k_density_sk_ring): one fat CTA per SM, matrix words staged with cp.async.  Used when the
// shard has enough rows to give every SM full CTAs and the ring plus both centre planes fit in shared memory; small
// instances keep the 128-thread kernel, which spreads few rows over more SMs.  NEMB200_DENS_RING=0/1 forces the choice
// (tests run both on the same inputs: the results are integers, so they must be identical).
template <int KB, int KX>
static int launch_density_ring(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st, bool* done, int k0 = 0,
                               int kc = -1 /*classes [k0, k0 + kc) of the model; -1: all*/) {
  constexpr int RB = KB <= 4 ? 4 : (KB <= 20 ? 2 : 1);
  constexpr int TPB = KB <= 3 ? 640 : (KB <= 8 ? 512 : 384);   // registers: 96 at K <= 3, up to 127 / 168 above
  constexpr int P = RB == 4 ? 3 : (RB == 2 ? 4 : 6);  // steps in flight per thread (RB x 16 B each)
  *done = false;
  if (kc < 0) kc = pl->K;
  const int W4 = pl->Ws / 4;
  const int G = group_lanes(W4);
  const size_t smem = ((size_t)P * RB * TPB + (size_t)2 * kc * W4) * sizeof(uint4);
  const long long rows_per_block = (long long)(TPB / 32) * (32 / G) * RB;
  const char* force = getenv("NEMB200_DENS_RING");
  if (smem > 220 * 1024) return 0;
  if (force ? atoi(force) == 0 : pl->N_local < rows_per_block * num_sms()) return 0;
  auto kern = k_density_sk_ring<KB, KX, RB, P, TPB>;
  static size_t smem_set[64] = {0};   // per instantiation and device: the opt-in only has to grow
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) smem_set[dev] = smem;
  }
  long long blocks = (pl->N_local + rows_per_block - 1) / rows_per_block;
  if (blocks > num_sms()) blocks = num_sms();
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, TPB, smem, st>>>((const uint4*)bits, pl->N_local, W4, kc, G, (const uint4*)pl->P.m1 + (size_t)k0 * W4,
                                            (const uint4*)pl->P.valid + (size_t)k0 * W4, pl->P.L + k0, pl->P.B + k0, pl->P.has_half,
                                            logf + k0, pl->K, pl->D);
  CK(cudaGetLastError());
  *done = true;
  return 0;
}
template <int KB, int KX>
static int launch_density_any(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st) {
  bool done = false;
  const int rc = launch_density_ring<KB, KX>(pl, bits, logf, st, &done);
  if (rc != 0 || done) return rc;
  return launch_density_sk<KB, KX>(pl, bits, logf, st);
}
// More than 8 classes on a large shard: passes over class chunks of <= 8 with the exact-K kernels.  The accumulators of
// K classes x RB rows do not fit the register file at K = 12..32 with RB > 1, and the kernel is bound by its LOP3 / POPC
// work, far from HBM, so reading the matrix once per chunk costs nothing (K = 20: 1.57 -> 1.2 ms on the bench matrix).
static int density_ring_chunk(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st, bool* done, int k0, int kc) {
  switch (kc) {
    case 1: case 2: return launch_density_ring<2, 0>(pl, bits, logf, st, done, k0, kc);
    case 3: return launch_density_ring<3, 3>(pl, bits, logf, st, done, k0, kc);
    case 4: return launch_density_ring<4, 4>(pl, bits, logf, st, done, k0, kc);
    case 5: return launch_density_ring<5, 5>(pl, bits, logf, st, done, k0, kc);
    case 6: return launch_density_ring<6, 6>(pl, bits, logf, st, done, k0, kc);
    case 7: return launch_density_ring<7, 7>(pl, bits, logf, st, done, k0, kc);
    default: return launch_density_ring<8, 8>(pl, bits, logf, st, done, k0, kc);
  }
}
static int dispatch_density_sk(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st) {
  if (pl->K > 8) {
    const int nchunk = (pl->K + 7) / 8;
    int k0 = 0;
    bool all = true;
    for (int c = 0; c < nchunk && all; c++) {
      const int kc = (pl->K - k0 + (nchunk - c) - 1) / (nchunk - c);   // balanced chunk sizes
      bool done = false;
      const int rc = density_ring_chunk(pl, bits, logf, st, &done, k0, kc);
      if (rc) return rc;
      if (!done) { all = false; break; }                                // not a ring-kernel shape: one launch below
      k0 += kc;
    }
    if (all) return 0;
  }
  switch (pl->K) {  // exact-K kernels for the common small K (no per-class predicates), buckets above
    case 1: case 2: return launch_density_any<2, 0>(pl, bits, logf, st);
    case 3: return launch_density_any<3, 3>(pl, bits, logf, st);
    case 4: return launch_density_any<4, 4>(pl, bits, logf, st);
    case 5: return launch_density_any<5, 5>(pl, bits, logf, st);
    case 6: return launch_density_any<6, 6>(pl, bits, logf, st);
    case 7: return launch_density_any<7, 7>(pl, bits, logf, st);
    case 8: return launch_density_any<8, 8>(pl, bits, logf, st);
    default: break;
  }
  if (pl->K <= 12) return launch_density_any<12, 0>(pl, bits, logf, st);
  if (pl->K <= 16) return launch_density_any<16, 0>(pl, bits, logf, st);
  if (pl->K <= 20) return launch_density_any<20, 0>(pl, bits, logf, st);
  return launch_density_any<32, 0>(pl, bits, logf, st);
}
template <int KB>
static int launch_density_skd(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st) {
  const int G = group_lanes(pl->Ws);
  const long long rows_per_block = (long long)(256 / 32) * (32 / G);
  long long blocks = (pl->N_local + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_density_skd<KB><<<(unsigned)blocks, 256, 0, st>>>(bits, pl->N_local, pl->Ws, pl->K, G, pl->P.m1, pl->P.valid, pl->Ltab,
                                                      pl->Lword, pl->P.B, logf);
  CK(cudaGetLastError());
  return 0;
}

static int density_to(nemb200_plan* pl, const uint32_t* bits, double* logf, cudaStream_t st) {
  if (pl->free_disp) return DISPATCH_KB(pl->K, launch_density_skd, pl, bits, logf, st);
  return dispatch_density_sk(pl, bits, logf, st);
}

extern "C" int nemb200_density(nemb200_plan* pl, const uint32_t* bits, double* logf, void* stream_) {
  if (!pl || (pl->N_local > 0 && (!bits || !logf))) return fail(NEMB200_ERR_ARG, "density: null argument");
  if (pl->N_local == 0) return NEMB200_OK;
  pl->stepwise_used = true;
  return density_to(pl, bits, logf, (cudaStream_t)stream_);
}

// ---- row-sharded exchange over peer memory -------------------------------------------------------------------------
extern "C" int nemb200_peer_export(nemb200_plan* pl, unsigned char handle_out[64], int64_t offsets_out[NEMB200_PEER_OFFSETS]) {
  if (!pl || !handle_out || !offsets_out) return fail(NEMB200_ERR_ARG, "peer_export: null argument");
  if (!pl->logf_full) return fail(NEMB200_ERR_ARG, "peer_export: plan was not created with NEMB200_PEER_EXCHANGE");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, pl->block));
  memcpy(handle_out, &hd, 64);
  const char* base = (const char*)pl->block;
  offsets_out[0] = (const char*)pl->cnt - base;
  offsets_out[1] = (const char*)pl->fsum - base;
  offsets_out[2] = (const char*)pl->logf_full - base;
  offsets_out[3] = (const char*)pl->sig - base;
  offsets_out[4] = (const char*)pl->P.mu - base; offsets_out[5] = (const char*)pl->P.eps - base;
  offsets_out[6] = (const char*)pl->P.L - base; offsets_out[7] = (const char*)pl->P.m1 - base;
  offsets_out[8] = (const char*)pl->P.valid - base; offsets_out[9] = (const char*)pl->fin_partial - base;
  offsets_out[10] = (const char*)pl->P.has_half - base;
  offsets_out[11] = (const char*)pl->ll_base - base;
  for (int q = 12; q < NEMB200_PEER_OFFSETS; q++) offsets_out[q] = 0;
  return NEMB200_OK;
}

extern "C" int nemb200_peer_attach(nemb200_plan* pl, int32_t world, int32_t rank, const unsigned char* handles,
                                   const int64_t* offsets) {
  if (!pl || !handles || !offsets || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(NEMB200_ERR_ARG, "peer_attach: bad argument (world <= %d)", kMaxPeers);
  if (!pl->logf_full) return fail(NEMB200_ERR_ARG, "peer_attach: plan was not created with NEMB200_PEER_EXCHANGE");
  int* sig_host[kMaxPeers] = {nullptr};
  for (int r = 0; r < world; r++) {
    char* base;
    if (r == rank) base = (char*)pl->block;
    else {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, handles + (size_t)r * 64, 64);
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      pl->peer_base[r] = p;
      base = (char*)p;
    }
    const int64_t* o = offsets + (size_t)r * NEMB200_PEER_OFFSETS;
    pl->peer_iblock[r] = (const int*)(base + o[0]);
    pl->peer_dblock[r] = (const double*)(base + o[1]);
    pl->peer_logf[r] = (double*)(base + o[2]);
    sig_host[r] = (int*)(base + o[3]);
    pl->peer_params.mu[r] = (float*)(base + o[4]); pl->peer_params.eps[r] = (float*)(base + o[5]);
    pl->peer_params.L[r] = (double*)(base + o[6]); pl->peer_params.m1[r] = (uint32_t*)(base + o[7]);
    pl->peer_params.valid[r] = (uint32_t*)(base + o[8]); pl->peer_params.fin_partial[r] = (double*)(base + o[9]);
    pl->peer_params.has_half[r] = (int*)(base + o[10]);
    pl->peer_ll[r] = base + o[11];
  }
  pl->peer_params.n = world; pl->peer_params.rank = rank;
  CK(cudaMemcpy(pl->sig_ptrs, sig_host, sizeof sig_host, cudaMemcpyHostToDevice));
  pl->world = world; pl->rank = rank; pl->attached = true;
  return NEMB200_OK;
}

extern "C" int nemb200_peer_logf(nemb200_plan* pl, double** logf_full) {
  if (!pl || !logf_full) return fail(NEMB200_ERR_ARG, "peer_logf: null argument");
  *logf_full = pl->logf_full;
  return NEMB200_OK;
}

/* E-step part 1 for a shard: densities of this rank's rows, stored into the logf buffer of every rank at row_offset
 * (the all-gather, fused into the kernel's epilogue), then this rank's "densities complete" epoch is published. */
extern "C" int nemb200_density_peers(nemb200_plan* pl, const uint32_t* bits, int64_t row_offset, void* stream_) {
  if (!pl || !pl->attached) return fail(NEMB200_ERR_ARG, "density_peers: plan is not attached to its peers");
  cudaStream_t st = (cudaStream_t)stream_;
  if (pl->N_local > 0) {
    if (!bits) return fail(NEMB200_ERR_ARG, "density_peers: null matrix");
    // densities into this rank's own buffer at full streaming speed, then one wide push of the shard to every peer
    int rc = density_to(pl, bits, pl->logf_full + (size_t)row_offset * pl->K, st);
    if (rc) return rc;
    if (pl->world > 1) {
      PeerOut out;
      out.n = pl->world; out.ptr[0] = pl->logf_full;
      int q = 1;
      for (int r = 0; r < pl->world; r++) if (r != pl->rank) out.ptr[q++] = pl->peer_logf[r];
      long long first = row_offset * pl->K, count = pl->N_local * (long long)pl->K;   // the shard, in doubles
      if (first & 1) {                                  // 16-byte stores need an even start: leading double on its own
        PeerOut one = out;
        one.row_off = first;
        k_peer_broadcast<<<dim3(1, pl->world - 1), 32, 0, st>>>(pl->logf_full, one, 1, nullptr, nullptr, 0, 0, 0);
        CK(cudaGetLastError());
        first += 1; count -= 1;
      }
      if (count > 0) {   // the last CTA of the push publishes "densities of this epoch are in place"
        out.row_off = first;
        dim3 grid(std::max(1, std::min(32, (int)((count / 2 + 255) / 256))), pl->world - 1);
        k_peer_broadcast<<<grid, 256, 0, st>>>(pl->logf_full, out, count, pl->bcast_ticket, pl->sig_ptrs, 1, pl->rank,
                                               ++pl->epoch_dens);
        CK(cudaGetLastError());
        return NEMB200_OK;
      }
    }
  }
  k_peer_signal<<<1, 32, 0, st>>>(pl->sig_ptrs, pl->world, 1, pl->rank, ++pl->epoch_dens);
  CK(cudaGetLastError());
  return NEMB200_OK;
}
extern "C" int nemb200_peer_wait_densities(nemb200_plan* pl, void* stream_) {
  if (!pl || !pl->attached) return fail(NEMB200_ERR_ARG, "peer_wait_densities: plan is not attached to its peers");
  (void)stream_;
  pl->sweep_waits = true;   // folded into the next nemb200_sweep: its CTAs check the flags themselves
  return NEMB200_OK;
}

extern "C" int nemb200_plan_graph_changed(nemb200_plan* pl) {
  if (!pl) return fail(NEMB200_ERR_ARG, "plan_graph_changed: null plan");
  pl->ell_N = 0;
  pl->run_ell_key[0] = nullptr;
  return NEMB200_OK;
}

extern "C" int nemb200_sweep(nemb200_plan* pl, int64_t N, const double* logf, const float* T_old, float* T_new,
                             const int32_t* rowptr, const int32_t* col, const float* w,
                             const int32_t* sched_level_ptr, const int32_t* sched_rows, int32_t n_levels,
                             float beta_now, uint32_t* maxdiff_bits, void* stream_) {
  if (!pl || !logf || !T_old || !T_new || N < 1) return fail(NEMB200_ERR_ARG, "sweep: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  pl->stepwise_used = true;
  if (!maxdiff_bits) {  // use (and reset) the plan's own convergence word, read back with nemb200_read_flags
    maxdiff_bits = pl->flags2;
    CK(cudaMemsetAsync(pl->flags2, 0, sizeof(uint32_t), st));
  }
  SweepArgs a;
  a.N = N; a.K = pl->K; a.logf = logf; a.pk = pl->P.pk; a.logpk64 = pl->P.logpk64; a.T_old = T_old; a.T_new = T_new;
  a.rowptr = rowptr; a.col = col; a.w = w; a.beta = beta_now;
  a.logspace = (pl->flags & NEMB200_EPI_LOGSPACE) ? 1 : 0;
  a.jacobi = (pl->flags & NEMB200_JACOBI) ? 1 : 0;
  a.maxdiff_bits = maxdiff_bits;
  a.ell_hdr = nullptr; a.ell_nb = nullptr; a.stamps = nullptr;
  a.chg = nullptr; a.spec_words = nullptr; a.spec_limit = 0; a.spec_parity = 0;   // the host-polled driver sweeps by levels only
  a.wait_sig = nullptr; a.wait_n = 0; a.wait_epoch = 0; a.wait_err = nullptr;
  if (pl->sweep_waits) {
    a.wait_sig = pl->sig; a.wait_n = pl->world; a.wait_epoch = pl->epoch_dens; a.wait_err = pl->peer_err;
    pl->sweep_waits = false;
  }
#ifdef NEMB200_SWEEP_STAMPS
  static unsigned long long* stampbuf = nullptr;
  static int stampcalls = 0;
  if (getenv("NEMB200_SWEEP_STAMPS")) {
    if (!stampbuf) { cudaMallocManaged(&stampbuf, 256 * 64 * 8); memset(stampbuf, 0, 256 * 64 * 8); }
    a.stamps = stampbuf;
  }
#endif
  const bool need_levels = (beta_now != 0.0f) && rowptr != nullptr && !a.jacobi;
  if (need_levels && (!sched_level_ptr || !sched_rows || n_levels < 1))
    return fail(NEMB200_ERR_ARG, "sweep: Gauss-Seidel sweep needs a level schedule (nemb200_gs_schedule)");
  const int GS = pl->K <= 4 ? 4 : (pl->K <= 8 ? 8 : (pl->K <= 16 ? 16 : 32));
  const long long rows_per_block = 8LL * (32 / GS);
  if (need_levels && n_levels > 1) {
    a.level_ptr = sched_level_ptr; a.level_rows = sched_rows; a.n_levels = n_levels;
    {  // schedule-ordered graph copy: rebuilt only when the caller passes other arrays (they must not change in place)
      const void* key[4] = {rowptr, col, w, sched_rows};
      if (!pl->ell || pl->ell_N != N || memcmp(key, pl->ell_key, sizeof(key)) != 0) {
        const size_t bytes = (size_t)N * (sizeof(int4) + kEllW * sizeof(int2));
        if (pl->ell && pl->ell_bytes < bytes) { CK(cudaStreamSynchronize(st)); cached_free(pl->ell, pl->ell_bytes); pl->ell = nullptr; }
        if (!pl->ell) {
          int rc = cached_malloc(&pl->ell, bytes);
          if (rc) return rc;
          pl->ell_bytes = bytes;
        }
        k_sweep_ell_build<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, sched_rows, rowptr, col, w, (int4*)pl->ell,
                                                                     (int2*)((int4*)pl->ell + N));
        CK(cudaGetLastError());
        memcpy(pl->ell_key, key, sizeof(key));
        pl->ell_N = N;
      }
      a.ell_hdr = (const int4*)pl->ell;
      a.ell_nb = (const int2*)((const int4*)pl->ell + N);
    }
    void* fn = GS == 4 ? (void*)k_sweep<true, 4> : GS == 8 ? (void*)k_sweep<true, 8> : GS == 16 ? (void*)k_sweep<true, 16>
                                                                                             : (void*)k_sweep<true, 32>;
    // few, fat CTAs: the grid barrier between levels is the fixed cost of this kernel (ncu: more than half of the
    // warp samples sit in it with 592 CTAs of 256 threads), and its price grows with the number of arriving CTAs
    const int tpb = 1024;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, tpb, 0));
    if (per_sm < 1) return fail(NEMB200_ERR_CUDA, "sweep: kernel does not fit on an SM");
    long long blocks = (long long)num_sms() * per_sm;
    const long long rows_per_cta = (long long)(tpb / 32) * (32 / GS);
    const long long want = (N + rows_per_cta - 1) / rows_per_cta;
    if (blocks > want) blocks = std::max<long long>(want, 1);
    if (pl->sweep_cta_cap > 0) blocks = std::min<long long>(blocks, pl->sweep_cta_cap);   // batched instances share the SMs
    void* args[] = {&a};
    CK(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(tpb), args, 0, st));
#ifdef NEMB200_SWEEP_STAMPS
    if (a.stamps && ++stampcalls == 20) {   // one warmed-up launch: first / last CTA to reach each stamp, relative to the first start
      cudaStreamSynchronize(st);
      unsigned long long t0 = ~0ull;
      for (int c = 0; c < blocks; c++) t0 = std::min(t0, stampbuf[c * 64]);
      for (int s2 = 0; s2 < std::min(2 * n_levels, 64); s2++) {
        unsigned long long mn = ~0ull, mx = 0;
        for (int c = 0; c < blocks; c++) { const unsigned long long v = stampbuf[c * 64 + s2]; if (v >= t0) { mn = std::min(mn, v); mx = std::max(mx, v); } }
        if (mx) fprintf(stderr, "sweep stamp %2d (%s level %d): first CTA %7.2f us, last CTA %7.2f us\n", s2, (s2 & 1) ? "end of  " : "start of", s2 / 2, (mn - t0) / 1e3, (mx - t0) / 1e3);
      }
    }
#endif
  } else {
    a.level_ptr = nullptr; a.level_rows = nullptr; a.n_levels = 1;
    long long blocks = (N + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)a.K * 256 * sizeof(double);   // sweep_independent_body (no neighbour term): a row's per-class terms
    if (smem > 48 * 1024) {
      const void* fn = GS == 16 ? (const void*)k_sweep<false, 16> : (const void*)k_sweep<false, 32>;
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (GS == 4) k_sweep<false, 4><<<(unsigned)blocks, 256, smem, st>>>(a);
    else if (GS == 8) k_sweep<false, 8><<<(unsigned)blocks, 256, smem, st>>>(a);
    else if (GS == 16) k_sweep<false, 16><<<(unsigned)blocks, 256, smem, st>>>(a);
    else k_sweep<false, 32><<<(unsigned)blocks, 256, smem, st>>>(a);
    CK(cudaGetLastError());
  }
  return NEMB200_OK;
}

// the four launches of the exact parallel float chains (nemb200_kernels.cuh, k_cc_*) for `njobs` jobs of at most max_terms terms
static void chain_launch(const CcJob* jobs_dev, const CcJob& job0, int njobs, size_t max_terms, cudaStream_t st) {
  const size_t nch = (max_terms + kCcChunk - 1) / kCcChunk + 1;
  const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((nch + 7) / 8, (size_t)std::max(1, num_sms() * 4 / njobs)));
  k_cc_sums<<<dim3(gx, njobs), 256, 0, st>>>(jobs_dev, job0);
  k_cc_predict<<<dim3(1, njobs), 64, 0, st>>>(jobs_dev, job0);
  k_cc_records<<<dim3(gx, njobs), 256, 0, st>>>(jobs_dev, job0);
  k_cc_walk<<<dim3(1, njobs), 64, 0, st>>>(jobs_dev, job0);
}

// the two chain jobs of a plan's criteria: D / G over the compacted float terms, L / Z over the per-row double terms
static void chain_jobs(nemb200_plan* pl, int64_t N, int n_regions, const unsigned* run_state, CcJob* dg, CcJob* lz) {
  const size_t nch = ((size_t)N * pl->K + kCcChunk - 1) / kCcChunk + 1;
  const size_t nch_lz = ((size_t)N + kCcChunk - 1) / kCcChunk + 1;
  dg->terms = pl->crit_dense; dg->total_ptr = pl->crit_offsets + n_regions;
  dg->recR = pl->crit_rec; dg->recMn = pl->crit_rec + nch * 2; dg->recMx = pl->crit_rec + nch * 4; dg->recFlag = pl->crit_recbad;
  dg->csum = (double*)(pl->crit_rec + nch * 6); dg->epred = pl->crit_recbad + nch * 2;
  dg->state = pl->crit_state; dg->out2 = pl->crit4; dg->run_state = run_state; dg->double_terms = 0;
  lz->terms = pl->crit_lz; lz->total_ptr = pl->crit_nrows;
  lz->recR = pl->crit_rec_lz; lz->recMn = pl->crit_rec_lz + nch_lz * 2; lz->recMx = pl->crit_rec_lz + nch_lz * 4; lz->recFlag = pl->crit_recbad_lz;
  lz->csum = (double*)(pl->crit_rec_lz + nch_lz * 6); lz->epred = pl->crit_recbad_lz + nch_lz * 2;
  lz->state = pl->crit_state_lz; lz->out2 = pl->crit4 + 2; lz->run_state = run_state; lz->double_terms = 1;
}

static int criteria_enqueue(nemb200_plan* pl, int64_t N, const double* logf, const float* T, const int32_t* rowptr,
                            const int32_t* col, const float* w, float beta, cudaStream_t st) {
  CritArgs a;
  a.N = N; a.K = pl->K; a.logf = logf; a.pk = pl->P.pk; a.logpk32 = pl->P.logpk32; a.logpk64 = pl->P.logpk64; a.T = T;
  a.rowptr = rowptr; a.col = col; a.w = w; a.beta = beta; a.logspace = (pl->flags & NEMB200_EPI_LOGSPACE) ? 1 : 0;
  a.partial = pl->crit_partial;
  if (N > pl->N_total) return fail(NEMB200_ERR_ARG, "criteria: N exceeds the plan's total row count");
  const int n_warps = pl->crit_blocks * 8;
  a.rows_per_warp = (N + n_warps - 1) / n_warps;
  a.terms = a.logspace ? nullptr : (float2*)pl->crit_terms;
  a.counts = a.logspace ? nullptr : pl->crit_counts;
  a.row_lz = a.logspace ? nullptr : (double2*)pl->crit_lz;
  const int GS = pl->K <= 4 ? 4 : (pl->K <= 8 ? 8 : (pl->K <= 16 ? 16 : 32));
  if (GS == 4) k_criteria<4><<<pl->crit_blocks, 256, 0, st>>>(a);
  else if (GS == 8) k_criteria<8><<<pl->crit_blocks, 256, 0, st>>>(a);
  else if (GS == 16) k_criteria<16><<<pl->crit_blocks, 256, 0, st>>>(a);
  else k_criteria<32><<<pl->crit_blocks, 256, 0, st>>>(a);
  CK(cudaGetLastError());
  if (a.logspace) {
    k_crit_reduce<<<1, 128, 0, st>>>(pl->crit_partial, pl->crit_blocks, pl->crit4, 0);
    CK(cudaGetLastError());
  } else {
    k_crit_offsets<<<1, 32, 0, st>>>(pl->crit_counts, n_warps, pl->crit_offsets, pl->crit_nrows, (int)N);
    CK(cudaGetLastError());
    k_crit_compact<<<(n_warps + 7) / 8, 256, 0, st>>>((const float2*)pl->crit_terms, pl->crit_counts, pl->crit_offsets, n_warps,
                                                    a.rows_per_warp, pl->K, (float2*)pl->crit_dense);
    CK(cudaGetLastError());
    // all four sums are float32 accumulators walked in the reference's order (nem_alg.c:2813-2827): D, G over the compacted
    // non-zero (d_ik, g_ik) terms, L, Z over the per-row double terms -- a sequential chain for small instances, the exact
    // parallel scheme above it
    const char* force = getenv("NEMB200_CRIT_SEQ");
    CcJob dg, lz;
    chain_jobs(pl, N, n_warps, nullptr, &dg, &lz);
    if (force ? atoi(force) != 0 : (size_t)N * pl->K < 8192) {
      k_crit_chain<TermsF2><<<1, 64, 0, st>>>(TermsF2{(const float2*)dg.terms}, dg.total_ptr, dg.out2);
      k_crit_chain<TermsD2><<<1, 64, 0, st>>>(TermsD2{(const double2*)lz.terms}, lz.total_ptr, lz.out2);
    } else {
      chain_launch(nullptr, dg, 1, (size_t)N * pl->K, st);     // (batched runs put both jobs of all instances in one launch)
      chain_launch(nullptr, lz, 1, (size_t)N, st);
    }
    CK(cudaGetLastError());
  }
  return 0;
}

// the four sums -> U, D, L, M, Z with the reference's expression types (nem_alg.c:2834-2835)
static void crit_combine(const double h[4], float beta, bool logspace, double crit_out[5]) {
  if (logspace) {
    const double cD = h[0], cG = h[1], cL = h[2], cZ = h[3];
    crit_out[0] = cD + 0.5 * (double)beta * cG;  // U  (nem_alg.c:2834)
    crit_out[1] = cD;
    crit_out[2] = cL;
    crit_out[3] = cD + (double)beta * cG + cZ;    // M  (nem_alg.c:2835)
    crit_out[4] = cZ;
  } else {
    const float cD = (float)h[0], cG = (float)h[1], cL = (float)h[2], cZ = (float)h[3];
    const float cU = (float)((double)cD + 0.5 * (double)beta * (double)cG);
    volatile float t1 = beta * cG;
    volatile float t2 = cD + t1;
    const float cM = t2 + cZ;
    crit_out[0] = cU; crit_out[1] = cD; crit_out[2] = cL; crit_out[3] = cM; crit_out[4] = cZ;
  }
}

static int criteria_collect(nemb200_plan* pl, float beta, double crit_out[5], cudaStream_t st) {
  double h[4];
  CK(cudaMemcpyAsync(h, pl->crit4, sizeof h, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  crit_combine(h, beta, (pl->flags & NEMB200_EPI_LOGSPACE) != 0, crit_out);
  return NEMB200_OK;
}

extern "C" int nemb200_criteria(nemb200_plan* pl, int64_t N, const double* logf, const float* T, const int32_t* rowptr,
                                const int32_t* col, const float* w, float beta, double crit_out[5], void* stream_) {
  if (!pl || !logf || !T || !crit_out) return fail(NEMB200_ERR_ARG, "criteria: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = criteria_enqueue(pl, N, logf, T, rowptr, col, w, beta, st);
  if (rc) return rc;
  return criteria_collect(pl, beta, crit_out, st);
}

extern "C" int nemb200_selftest_float_chain(const float* d, const float* g, int64_t n, int32_t parallel, double out2[2]) {
  if (!d || !g || n < 0 || n > 0x7fffffffLL || !out2) return fail(NEMB200_ERR_ARG, "selftest_float_chain: bad argument");
  std::vector<float2> h((size_t)std::max<int64_t>(n, 1));
  for (int64_t i = 0; i < n; i++) h[(size_t)i] = make_float2(d[i], g[i]);
  const size_t nch = ((size_t)n + kCcChunk - 1) / kCcChunk + 1;
  const size_t b_dense = h.size() * sizeof(float2), b_rec = nch * 2 * 4 * sizeof(long long), b_bad = nch * 2 * 2 * sizeof(int);
  void *p_dense = nullptr, *p_rec = nullptr, *p_bad = nullptr, *p_misc = nullptr;
  int rc;
  if ((rc = cached_malloc(&p_dense, b_dense))) return rc;
  if ((rc = cached_malloc(&p_rec, b_rec))) { cached_free(p_dense, b_dense); return rc; }
  if ((rc = cached_malloc(&p_bad, b_bad))) { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); return rc; }
  if ((rc = cached_malloc(&p_misc, 256))) { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); cached_free(p_bad, b_bad); return rc; }
  auto release = [&]() { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); cached_free(p_bad, b_bad); cached_free(p_misc, 256); };
  int* offs = (int*)p_misc;                       // [2] offsets, then the state, then the two results
  CcState* stp = (CcState*)((char*)p_misc + 64);
  double* out = (double*)((char*)p_misc + 128);
  const int hoffs[2] = {0, (int)n};
  cudaError_t e = cudaMemcpy(p_dense, h.data(), b_dense, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(offs, hoffs, sizeof hoffs, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    TermsF2 dense{(const float2*)p_dense};
    const int* tot = offs + 1;
    if (parallel) {
      long long *rR = (long long*)p_rec;
      CcJob job{p_dense, tot, rR, rR + nch * 2, rR + nch * 4, (double*)(rR + nch * 6), (int*)p_bad, (int*)p_bad + nch * 2, stp, out, nullptr, 0};
      chain_launch(nullptr, job, 1, (size_t)n, nullptr);
    } else {
      k_crit_chain<TermsF2><<<1, 64>>>(dense, tot, out);
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out2, out, 2 * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && parallel && getenv("NEMB200_CHAIN_DEBUG")) {
    CcState hs;
    if (cudaMemcpy(&hs, stp, sizeof hs, cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "float chain: %lld terms, %zu chunks, chunks added sequentially: %d / %d\n", (long long)n, nch - 1,
              hs.seq_chunks[0], hs.seq_chunks[1]);
  }
  release();
  if (e != cudaSuccess) return fail(NEMB200_ERR_CUDA, "selftest_float_chain: %s", cudaGetErrorString(e));
  return NEMB200_OK;
}

/* the same for the L / Z kind of chain: double terms added to a float accumulator, acc = (float)((double)acc + x) */
extern "C" int nemb200_selftest_double_chain(const double* l, const double* z, int64_t n, int32_t parallel, double out2[2]) {
  if (!l || !z || n < 0 || n > 0x7fffffffLL || !out2) return fail(NEMB200_ERR_ARG, "selftest_double_chain: bad argument");
  std::vector<double2> h((size_t)std::max<int64_t>(n, 1));
  for (int64_t i = 0; i < n; i++) h[(size_t)i] = make_double2(l[i], z[i]);
  const size_t nch = ((size_t)n + kCcChunk - 1) / kCcChunk + 1;
  const size_t b_dense = h.size() * sizeof(double2), b_rec = nch * 2 * 4 * sizeof(long long), b_bad = nch * 2 * 2 * sizeof(int);
  void *p_dense = nullptr, *p_rec = nullptr, *p_bad = nullptr, *p_misc = nullptr;
  int rc;
  if ((rc = cached_malloc(&p_dense, b_dense))) return rc;
  if ((rc = cached_malloc(&p_rec, b_rec))) { cached_free(p_dense, b_dense); return rc; }
  if ((rc = cached_malloc(&p_bad, b_bad))) { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); return rc; }
  if ((rc = cached_malloc(&p_misc, 256))) { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); cached_free(p_bad, b_bad); return rc; }
  auto release = [&]() { cached_free(p_dense, b_dense); cached_free(p_rec, b_rec); cached_free(p_bad, b_bad); cached_free(p_misc, 256); };
  int* tot = (int*)p_misc;
  CcState* stp = (CcState*)((char*)p_misc + 64);
  double* out = (double*)((char*)p_misc + 128);
  const int hn = (int)n;
  cudaError_t e = cudaMemcpy(p_dense, h.data(), b_dense, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(tot, &hn, sizeof hn, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    TermsD2 dense{(const double2*)p_dense};
    const int* totp = tot;
    if (parallel) {
      long long *rR = (long long*)p_rec;
      CcJob job{p_dense, totp, rR, rR + nch * 2, rR + nch * 4, (double*)(rR + nch * 6), (int*)p_bad, (int*)p_bad + nch * 2, stp, out, nullptr, 1};
      chain_launch(nullptr, job, 1, (size_t)n, nullptr);
    } else {
      k_crit_chain<TermsD2><<<1, 64>>>(dense, totp, out);
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out2, out, 2 * sizeof(double), cudaMemcpyDeviceToHost);
  release();
  if (e != cudaSuccess) return fail(NEMB200_ERR_CUDA, "selftest_double_chain: %s", cudaGetErrorString(e));
  return NEMB200_OK;
}

extern "C" int nemb200_get_params(nemb200_plan* pl, float* mu, float* eps, float* pk, int32_t* status, void* stream_) {
  if (!pl) return fail(NEMB200_ERR_ARG, "get_params: null plan");
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t kd = (size_t)pl->K * pl->D * sizeof(float);
  if (eps && !pl->free_disp) {
    k_expand_eps<<<dim3(std::max(1, std::min(64, (pl->D + 255) / 256)), pl->K), 256, 0, st>>>(pl->P);
    CK(cudaGetLastError());
  }
  if (mu) CK(cudaMemcpyAsync(mu, pl->P.mu, kd, cudaMemcpyDeviceToHost, st));
  if (eps) CK(cudaMemcpyAsync(eps, pl->P.eps, kd, cudaMemcpyDeviceToHost, st));
  if (pk) CK(cudaMemcpyAsync(pk, pl->P.pk, pl->K * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (status) CK(cudaMemcpyAsync(status, pl->P.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return NEMB200_OK;
}

extern "C" int nemb200_read_flags(nemb200_plan* pl, float* maxdiff, int32_t* status, void* stream_) {
  if (!pl) return fail(NEMB200_ERR_ARG, "read_flags: null plan");
  cudaStream_t st = (cudaStream_t)stream_;
  uint32_t h[2] = {0, 0};
  int perr = 0;
  CK(cudaMemcpyAsync(h, pl->flags2, sizeof h, cudaMemcpyDeviceToHost, st));
  if (pl->attached) CK(cudaMemcpyAsync(&perr, pl->peer_err, sizeof perr, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (perr != 0) return fail(NEMB200_ERR_CUDA, "peer exchange: a wait for another rank timed out");
  if (maxdiff) memcpy(maxdiff, &h[0], sizeof(float));
  if (status) *status = (int32_t)h[1];
  return NEMB200_OK;
}

extern "C" int nemb200_labels(const float* T, int64_t N, int32_t K, int32_t* labels_dev, double* entropy_host,
                              void* stream_) {
  if (!T || !labels_dev || N < 1 || K < 1) return fail(NEMB200_ERR_ARG, "labels: bad argument");
  cudaStream_t st = (cudaStream_t)stream_;
  const int blocks = (int)((N + 255) / 256);
  void* mem = nullptr;
  const size_t bytes = (size_t)(blocks + 1) * sizeof(double);
  int rc = cached_malloc(&mem, bytes);   // cudaMalloc / cudaFree would cost more than the two kernels
  if (rc) return rc;
  double* part = (double*)mem;
  k_labels<<<blocks, 256, 0, st>>>(T, N, K, labels_dev, part);
  k_sum_partials<<<1, 32, 0, st>>>(part, blocks, part + blocks);
  double e = 0.0;
  cudaError_t err = cudaMemcpyAsync(&e, part + blocks, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  cached_free(mem, bytes);
  if (err != cudaSuccess) return fail(NEMB200_ERR_CUDA, "labels: %s", cudaGetErrorString(err));
  if (entropy_host) *entropy_host = e;
  return NEMB200_OK;
}

extern "C" int nemb200_gs_schedule(int64_t N, const int32_t* rowptr, const int32_t* col, int32_t* level_ptr_out,
                                   int32_t* rows_out) {
  if (N < 1 || !rowptr || !level_ptr_out || !rows_out) return fail(NEMB200_ERR_ARG, "gs_schedule: bad argument");
  std::vector<int32_t> level((size_t)N, 0);
  int32_t maxl = 0;
  for (int64_t i = 0; i < N; i++) {
    int32_t lv = 0;
    for (int32_t n = rowptr[i]; n < rowptr[i + 1]; n++) {
      const int32_t j = col[n];
      if (j < 0 || j >= N) return fail(NEMB200_ERR_ARG, "gs_schedule: neighbour %d of row %lld out of range", j, (long long)i);
      if (j < i) lv = std::max(lv, level[j] + 1);
    }
    level[i] = lv;
    maxl = std::max(maxl, lv);
  }
  const int32_t nl = maxl + 1;
  std::vector<int32_t> count((size_t)nl + 1, 0);
  for (int64_t i = 0; i < N; i++) count[level[i] + 1]++;
  for (int32_t l = 0; l < nl; l++) count[l + 1] += count[l];
  for (int32_t l = 0; l <= nl; l++) level_ptr_out[l] = count[l];
  std::vector<int32_t> cur(count.begin(), count.end() - 1);
  for (int64_t i = 0; i < N; i++) rows_out[cur[level[i]]++] = (int32_t)i;
  return nl;
}

// ------------------------------------------------------------------------------------------------------------------
// whole-run drivers
// ------------------------------------------------------------------------------------------------------------------
struct RunBuffers {   // views into the plan's block (plan_build(..., with_run_buffers = true))
  double* logf = nullptr;
  float* Ta = nullptr;
  float* Tb = nullptr;
  uint32_t* flags = nullptr;  // [0] maxdiff bits
};

// One EM run as a small state machine, so that a single instance and a batch of independent instances (K sweep,
// chunk samples: partition.py:482-514, :837-846) share the same code: begin() and enqueue_iteration() only enqueue
// work on the instance's stream; poll() does the iteration's single 8-byte read-back.
struct EmRun {
  nemb200_plan* pl = nullptr;
  const uint32_t* bits = nullptr;
  const int32_t *rowptr = nullptr, *col = nullptr, *lvl_ptr = nullptr, *lvl_rows = nullptr;
  const float* w = nullptr;
  int n_levels = 1;
  float beta = 0.0f, cv_th = 0.01f;
  int itermax = 100;
  cudaStream_t st = nullptr;
  RunBuffers rb;
  float *cur = nullptr, *other = nullptr;
  int iter = 0, converged = 0, status = NEMB200_STS_OK;
  bool active = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double crit[5] = {0, 0, 0, 0, 0};
  float ms = 0.0f;
  ~EmRun() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }

  int begin() {  // nem_alg.c:2023-2065 with Needinit = 1
    const long long N = pl->N_local;
    const int K = pl->K;
    int rc;
    if (!pl->with_run_buffers) return fail(NEMB200_ERR_ARG, "EmRun: plan built without run buffers");
    rb.logf = pl->run_logf; rb.Ta = pl->run_Ta; rb.Tb = pl->run_Tb; rb.flags = pl->flags2;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, st));
    if ((rc = nemb200_init_params(pl, st))) return rc;
    CK(cudaMemsetAsync(rb.Ta, 0, (size_t)N * K * sizeof(float), st));  // ClassifM is calloc'ed (nem_exe.c:533-536)
    if ((rc = nemb200_density(pl, bits, rb.logf, st))) return rc;
    if ((rc = nemb200_sweep(pl, N, rb.logf, rb.Ta, rb.Tb, rowptr, col, w, lvl_ptr, lvl_rows, n_levels, 0.0f, rb.flags, st))) return rc;
    cur = rb.Tb;
    other = rb.Ta;
    if (beta != 0.0f) {
      if ((rc = nemb200_sweep(pl, N, rb.logf, cur, other, rowptr, col, w, lvl_ptr, lvl_rows, n_levels, beta, rb.flags, st))) return rc;
      std::swap(cur, other);
    }
    // with beta == 0 the second sweep of the reference recomputes the same posteriors: skipped
    active = itermax >= 1;
    return 0;
  }
  int enqueue_iteration() {  // nem_alg.c:1844-1863
    int rc;
    iter++;
    if ((rc = nemb200_mstep_accumulate(pl, bits, cur, st))) return rc;
    if ((rc = nemb200_mstep_finalize(pl, st))) return rc;
    if ((rc = nemb200_density(pl, bits, rb.logf, st))) return rc;
    return nemb200_sweep(pl, pl->N_local, rb.logf, cur, other, rowptr, col, w, lvl_ptr, lvl_rows, n_levels, beta, nullptr, st);
  }
  int poll() {  // nem_alg.c:1868-1892
    float maxdif = 0.0f;
    int32_t sts = 0;
    int rc = nemb200_read_flags(pl, &maxdif, &sts, st);  // the one device->host read of the iteration
    if (rc) return rc;
    status = sts;
    if (status != NEMB200_STS_OK) { active = false; return 0; }  // empty class: the reference stops (nem_alg.c:1873-1892)
    std::swap(cur, other);
    converged = (maxdif < cv_th);  // nem_alg.c:2160-2173
    if (converged || iter >= itermax) active = false;
    return 0;
  }
  int finish_enqueue() {  // nem_alg.c:1906, asynchronous part
    if (status == NEMB200_STS_OK) return criteria_enqueue(pl, pl->N_local, rb.logf, cur, rowptr, col, w, beta, st);
    return 0;
  }
  int finish(float* T_out, nemb200_result* res) {
    int rc;
    if (status == NEMB200_STS_OK) {
      if ((rc = criteria_collect(pl, beta, crit, st))) return rc;
      if (std::isinf(crit[0]) && crit[0] < 0) status = NEMB200_STS_INFINITE;  // nem_alg.c:2482-2487
    }
    CK(cudaEventRecord(ev1, st));
    CK(cudaEventSynchronize(ev1));
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    if (T_out) CK(cudaMemcpyAsync(T_out, cur, (size_t)pl->N_local * pl->K * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    if (res) {
      for (int q = 0; q < 5; q++) res->crit[q] = crit[q];
      res->n_iter = iter; res->converged = converged; res->status = status; res->n_levels = n_levels; res->elapsed_ms = ms;
      res->em_ms = ms; res->density_ms = 0.0f; res->density_launches = 0;
    }
    return NEMB200_OK;
  }
};

static int run_em(nemb200_plan* pl, const uint32_t* bits, const int32_t* rowptr, const int32_t* col, const float* w,
                  const int32_t* lvl_ptr, const int32_t* lvl_rows, int n_levels, float beta, int itermax, float cv_th,
                  float* T_out, nemb200_result* res, cudaStream_t st) {
  EmRun r;
  r.pl = pl; r.bits = bits; r.rowptr = rowptr; r.col = col; r.w = w; r.lvl_ptr = lvl_ptr; r.lvl_rows = lvl_rows;
  r.n_levels = n_levels; r.beta = beta; r.itermax = itermax; r.cv_th = cv_th; r.st = st;
  int rc = r.begin();
  if (rc) return rc;
  while (r.active) {
    if ((rc = r.enqueue_iteration())) return rc;
    if ((rc = r.poll())) return rc;
  }
  if ((rc = r.finish_enqueue())) return rc;
  return r.finish(T_out, res);
}

#include "nemb200_batch.inc"

static int copy_params_out(nemb200_plan* pl, float* mu, float* eps, float* pk, cudaStream_t st) {
  const size_t kd = (size_t)pl->K * pl->D * sizeof(float);
  if (eps && !pl->free_disp) {
    k_expand_eps<<<dim3(std::max(1, std::min(64, (pl->D + 255) / 256)), pl->K), 256, 0, st>>>(pl->P);
    CK(cudaGetLastError());
  }
  if (mu) CK(cudaMemcpyAsync(mu, pl->P.mu, kd, cudaMemcpyDefault, st));
  if (eps) CK(cudaMemcpyAsync(eps, pl->P.eps, kd, cudaMemcpyDefault, st));
  if (pk) CK(cudaMemcpyAsync(pk, pl->P.pk, pl->K * sizeof(float), cudaMemcpyDefault, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int nemb200_nem_device(const uint32_t* bits, int64_t N, int64_t D, int64_t row_stride_words,
                                  const int32_t* rowptr, const int32_t* col, const float* w,
                                  const int32_t* sched_level_ptr, const int32_t* sched_rows, int32_t n_levels,
                                  int32_t K, float beta, int32_t free_disp, int32_t itermax, float cv_th, uint32_t flags,
                                  float* T_out, float* mu_out, float* eps_out, float* pk_out, nemb200_result* result,
                                  void* stream_) {
  if (!bits || N < 1) return fail(NEMB200_ERR_ARG, "nem_device: null matrix");
  if (beta != 0.0f && (!rowptr || !col || !w)) return fail(NEMB200_ERR_ARG, "nem_device: beta != 0 needs the neighbour graph");
  if (K < 1 || K > NEMB200_MAX_K || row_stride_words % 4 != 0 || row_stride_words * 32 < D || D < 1)
    return fail(NEMB200_ERR_ARG, "nem_device: bad shape D=%lld stride=%lld K=%d", (long long)D, (long long)row_stride_words, K);
  if (batch_enabled()) {   // the device-resident loop (a batch of one)
    nemb200_result tmp;
    const std::vector<int> one{0};
    return batch_run(one, &bits, &N, &D, &row_stride_words, &rowptr, &col, &w, &sched_level_ptr, &sched_rows, &n_levels, &K, &beta,
                     &free_disp, &itermax, cv_th, flags, &T_out, &mu_out, &eps_out, &pk_out, result ? result : &tmp,
                     (cudaStream_t)stream_, true);
  }
  nemb200_plan* pl = nullptr;
  int rc = plan_build(&pl, N, N, D, row_stride_words, K, free_disp, flags, true);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  rc = run_em(pl, bits, rowptr, col, w, sched_level_ptr, sched_rows, n_levels, beta, itermax, cv_th, T_out, result, st);
  if (rc == 0) rc = copy_params_out(pl, mu_out, eps_out, pk_out, st);
  cudaStreamSynchronize(st);
  pl->stepwise_used = false;   // everything this run enqueued has completed
  nemb200_plan_destroy(pl);
  return rc;
}

/* One rank's part of a row-sharded run with the device-resident loop: `pl` was created with NEMB200_PEER_EXCHANGE for this
 * rank's rows and attached to its peers; every rank calls this with the same graph, K, beta and iteration cap. */
extern "C" int nemb200_nem_sharded(nemb200_plan* pl, const uint32_t* bits_local, int64_t row_offset, const int32_t* rowptr,
                                   const int32_t* col, const float* w, const int32_t* sched_level_ptr, const int32_t* sched_rows,
                                   int32_t n_levels, float beta, int32_t itermax, float cv_th, float* T_out, nemb200_result* result,
                                   void* stream_) {
  if (!pl || !pl->attached || !pl->logf_full) return fail(NEMB200_ERR_ARG, "nem_sharded: plan is not attached to its peers");
  if (pl->N_local > 0 && !bits_local) return fail(NEMB200_ERR_ARG, "nem_sharded: null matrix");
  if (beta != 0.0f && (!rowptr || !col || !w)) return fail(NEMB200_ERR_ARG, "nem_sharded: beta != 0 needs the neighbour graph");
  if (row_offset < 0 || row_offset + pl->N_local > pl->N_total) return fail(NEMB200_ERR_ARG, "nem_sharded: row block outside the instance");
  nemb200_result tmp;
  const std::vector<int> one{0};
  const int64_t N = pl->N_local, D = pl->D, Ws = pl->Ws;
  const int32_t K = pl->K, fd = pl->free_disp;
  float *nomu = nullptr;
  return batch_run(one, &bits_local, &N, &D, &Ws, &rowptr, &col, &w, &sched_level_ptr, &sched_rows, &n_levels, &K, &beta, &fd, &itermax,
                   cv_th, pl->flags, &T_out, &nomu, &nomu, &nomu, result ? result : &tmp, (cudaStream_t)stream_, true, pl, row_offset);
}

extern "C" int nemb200_nem_device_batch(int32_t n_inst, const uint32_t* const* bits, const int64_t* N, const int64_t* D,
                                        const int64_t* row_stride_words, const int32_t* const* rowptr,
                                        const int32_t* const* col, const float* const* w,
                                        const int32_t* const* sched_level_ptr, const int32_t* const* sched_rows,
                                        const int32_t* n_levels, const int32_t* K, const float* beta,
                                        const int32_t* free_disp, const int32_t* itermax, float cv_th, uint32_t flags,
                                        float* const* T_out, float* const* mu_out, float* const* eps_out,
                                        float* const* pk_out, nemb200_result* results) {
  if (n_inst < 1 || !bits || !N || !D || !row_stride_words || !K || !beta || !free_disp || !itermax || !results)
    return fail(NEMB200_ERR_ARG, "nem_device_batch: null argument");
  // grouped by kernel shape, one launch per EM phase and group, convergence on the device (nemb200_batch.inc); the
  // host-polled driver below only runs with NEMB200_NO_BATCH=1 (cross-check in the tests)
  std::vector<int> small;
  std::vector<char> is_small(n_inst, 0);
  for (int i = 0; i < n_inst; i++) {
    if (!bits[i] || N[i] < 1) return fail(NEMB200_ERR_ARG, "nem_device_batch: instance %d has no matrix", i);
    if (beta[i] != 0.0f && (!rowptr || !col || !w || !rowptr[i] || !col[i] || !w[i]))
      return fail(NEMB200_ERR_ARG, "nem_device_batch: instance %d needs its neighbour graph", i);
    if (K[i] < 1 || K[i] > NEMB200_MAX_K || row_stride_words[i] % 4 != 0 || row_stride_words[i] * 32 < D[i] || D[i] < 1)
      return fail(NEMB200_ERR_ARG, "nem_device_batch: instance %d has a bad shape", i);
    if (batch_enabled()) { small.push_back(i); is_small[i] = 1; }
  }
  if (!small.empty()) {
    const int brc = batch_run(small, bits, N, D, row_stride_words, rowptr, col, w, sched_level_ptr, sched_rows, n_levels, K, beta,
                              free_disp, itermax, cv_th, flags, T_out, mu_out, eps_out, pk_out, results);
    if (brc) return brc;
    if ((int)small.size() == n_inst) return NEMB200_OK;
  }
  std::vector<EmRun> runs(n_inst);
  std::vector<nemb200_plan*> plans(n_inst, nullptr);
  std::vector<cudaStream_t> streams(n_inst, nullptr);
  int rc = 0;
  auto cleanup = [&]() {
    for (auto s : streams) if (s) { cudaStreamSynchronize(s); }
    runs.clear();                                   // frees the per-run device buffers
    for (auto p : plans) { if (p) p->stepwise_used = false; nemb200_plan_destroy(p); }   // streams synchronised above
    for (auto s : streams) if (s) cudaStreamDestroy(s);
  };
  for (int i = 0; i < n_inst && rc == 0; i++) {
    if (is_small[i]) { runs[i].active = false; continue; }
    const bool graph = beta[i] != 0.0f;
    if (!bits[i] || N[i] < 1) { rc = fail(NEMB200_ERR_ARG, "nem_device_batch: instance %d has no matrix", i); break; }
    if (graph && (!rowptr || !col || !w || !rowptr[i] || !col[i] || !w[i])) { rc = fail(NEMB200_ERR_ARG, "nem_device_batch: instance %d needs its neighbour graph", i); break; }
    if (cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) != cudaSuccess) { rc = fail(NEMB200_ERR_CUDA, "nem_device_batch: stream"); break; }
    rc = plan_build(&plans[i], N[i], N[i], D[i], row_stride_words[i], K[i], free_disp[i], flags, true);
    if (rc) break;
    // sweeps of up to four instances side by side: measured on 8 chunk instances (100 000 x 500), 9.7 -> 7.8 ms per batch,
    // while a single instance loses nothing down to a third of the SMs (its levels hold < 2 rows per lane group)
    if (n_inst > 1) plans[i]->sweep_cta_cap = std::max(num_sms() / 4, (num_sms() + n_inst - 1) / n_inst);
    EmRun& r = runs[i];
    r.pl = plans[i]; r.bits = bits[i];
    r.rowptr = graph ? rowptr[i] : nullptr; r.col = graph ? col[i] : nullptr; r.w = graph ? w[i] : nullptr;
    r.lvl_ptr = (graph && sched_level_ptr) ? sched_level_ptr[i] : nullptr;
    r.lvl_rows = (graph && sched_rows) ? sched_rows[i] : nullptr;
    r.n_levels = (graph && n_levels) ? n_levels[i] : 1;
    r.beta = beta[i]; r.itermax = itermax[i]; r.cv_th = cv_th; r.st = streams[i];
    rc = r.begin();
  }
  // lock-step: enqueue one iteration of every live instance (their kernels overlap on the GPU), then read the flags
  bool any = rc == 0;
  while (any && rc == 0) {
    any = false;
    for (int i = 0; i < n_inst && rc == 0; i++)
      if (runs[i].active) rc = runs[i].enqueue_iteration();
    for (int i = 0; i < n_inst && rc == 0; i++)
      if (runs[i].active) { rc = runs[i].poll(); any = any || runs[i].active; }
  }
  for (int i = 0; i < n_inst && rc == 0; i++) if (!is_small[i]) rc = runs[i].finish_enqueue();   // all criteria kernels in flight together
  for (int i = 0; i < n_inst && rc == 0; i++) {
    if (is_small[i]) continue;
    rc = runs[i].finish(T_out ? T_out[i] : nullptr, &results[i]);
    if (rc == 0) rc = copy_params_out(plans[i], mu_out ? mu_out[i] : nullptr, eps_out ? eps_out[i] : nullptr,
                                      pk_out ? pk_out[i] : nullptr, streams[i]);
  }
  cleanup();
  return rc;
}

extern "C" int nemb200_nem_host(const uint32_t* bits, int64_t N, int64_t D, int64_t row_stride_words,
                                const int32_t* rowptr, const int32_t* col, const float* w, int32_t K, float beta,
                                int32_t free_disp, int32_t itermax, float cv_th, uint32_t flags, float* T_out,
                                float* mu_out, float* eps_out, float* pk_out, nemb200_result* result) {
  if (!bits || N < 1 || row_stride_words < 1) return fail(NEMB200_ERR_ARG, "nem_host: null matrix");
  const bool graph = (beta != 0.0f);
  if (graph && (!rowptr || !col || !w)) return fail(NEMB200_ERR_ARG, "nem_host: beta != 0 needs the neighbour graph");
  cudaStream_t st = nullptr;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct Guard {
    std::vector<std::pair<void*, size_t>> p; cudaStream_t s;
    ~Guard() { cudaStreamSynchronize(s); for (auto& q : p) cached_free(q.first, q.second); cudaStreamDestroy(s); }
  } g{{}, st};
  auto dmal = [&](void** p, size_t bytes) -> int {
    int r = cached_malloc(p, bytes);
    if (r == 0) g.p.emplace_back(*p, bytes);
    return r;
  };
  uint32_t* d_bits = nullptr; int32_t *d_rowptr = nullptr, *d_col = nullptr, *d_lp = nullptr, *d_lr = nullptr; float* d_w = nullptr;
  float* d_T = nullptr;
  int rc;
  const size_t bits_bytes = (size_t)N * row_stride_words * sizeof(uint32_t);
  if ((rc = dmal((void**)&d_bits, bits_bytes))) return rc;
  CK(cudaMemcpyAsync(d_bits, bits, bits_bytes, cudaMemcpyHostToDevice, st));
  int n_levels = 1;
  if (graph) {
    const int64_t nnz = rowptr[N];
    std::vector<int32_t> lp((size_t)N + 1), lr((size_t)N);
    n_levels = nemb200_gs_schedule(N, rowptr, col, lp.data(), lr.data());
    if (n_levels < 0) return n_levels;
    if ((rc = dmal((void**)&d_rowptr, (size_t)(N + 1) * 4))) return rc;
    if ((rc = dmal((void**)&d_col, (size_t)nnz * 4))) return rc;
    if ((rc = dmal((void**)&d_w, (size_t)nnz * 4))) return rc;
    if ((rc = dmal((void**)&d_lp, (size_t)(N + 1) * 4))) return rc;
    if ((rc = dmal((void**)&d_lr, (size_t)N * 4))) return rc;
    CK(cudaMemcpyAsync(d_rowptr, rowptr, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_col, col, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_w, w, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lp, lp.data(), (size_t)(n_levels + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lr, lr.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // lp / lr are locals
  }
  if ((rc = dmal((void**)&d_T, (size_t)N * K * sizeof(float)))) return rc;
  if (K >= 1 && K <= NEMB200_MAX_K && row_stride_words % 4 == 0 && batch_enabled()) {
    float *d_mu = nullptr, *d_eps = nullptr, *d_pk = nullptr;
    const size_t kd = (size_t)K * D * sizeof(float);
    if (mu_out && (rc = dmal((void**)&d_mu, kd))) return rc;
    if (eps_out && (rc = dmal((void**)&d_eps, kd))) return rc;
    if (pk_out && (rc = dmal((void**)&d_pk, K * sizeof(float)))) return rc;
    nemb200_result tmp;
    const std::vector<int> one{0};
    const uint32_t* cb = d_bits; const int32_t *crp = d_rowptr, *ccol = d_col, *clp = d_lp, *clr = d_lr; const float* cw = d_w;
    rc = batch_run(one, &cb, &N, &D, &row_stride_words, &crp, &ccol, &cw, &clp, &clr, &n_levels, &K, &beta, &free_disp, &itermax,
                   cv_th, flags, &d_T, &d_mu, &d_eps, &d_pk, result ? result : &tmp, st, true);
    if (rc) return rc;
    if (T_out) CK(cudaMemcpyAsync(T_out, d_T, (size_t)N * K * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (mu_out) CK(cudaMemcpyAsync(mu_out, d_mu, kd, cudaMemcpyDeviceToHost, st));
    if (eps_out) CK(cudaMemcpyAsync(eps_out, d_eps, kd, cudaMemcpyDeviceToHost, st));
    if (pk_out) CK(cudaMemcpyAsync(pk_out, d_pk, K * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return NEMB200_OK;
  }
  nemb200_plan* pl = nullptr;
  rc = plan_build(&pl, N, N, D, row_stride_words, K, free_disp, flags, true);
  if (rc) return rc;
  rc = run_em(pl, d_bits, d_rowptr, d_col, d_w, d_lp, d_lr, n_levels, beta, itermax, cv_th, d_T, result, st);
  if (rc == 0 && T_out) {
    cudaError_t e = cudaMemcpyAsync(T_out, d_T, (size_t)N * K * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = fail(NEMB200_ERR_CUDA, "nem_host: D2H %s", cudaGetErrorString(e));
  }
  if (rc == 0) rc = copy_params_out(pl, mu_out, eps_out, pk_out, st);
  cudaStreamSynchronize(st);
  pl->stepwise_used = false;
  nemb200_plan_destroy(pl);
  return rc;
}

#include "nemb200_export.inc"
#include "nem_shim.inc"
