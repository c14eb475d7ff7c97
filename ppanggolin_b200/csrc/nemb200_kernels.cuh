// nemb200_kernels.cuh -- sm_100a kernels of the NEM partitioning hot path.
//
// What each kernel restates (paths relative to /root/reference/ppanggolin/nem/):
//   k_density_sk / k_density_skd   NEM/nem_mod.c:618-689 (DensBernoulli) for all rows x classes, closed form:
//                                  log f_k(x_i) = sum_d log(1-e_kd) - sum_{d: x_id != mu_kd} log((1-e_kd)/e_kd)
//   k_sweep                        NEM/nem_alg.c:2412-2489 + :2630-2701 (in-place posterior sweep) via level schedule
//   k_classify/k_scatter/k_mstep_* NEM/nem_mod.c:1270-1312, :1317-1474, :1641-1699 (sizes, medians, inertia) as
//                                  column sums S1[k][d] = sum_i c_ik x_id over the bit-packed matrix
//   k_mstep_finalize               NEM/nem_mod.c:1020-1076 / :1131-1170 (dispersion), :455-459 (proportions)
//   k_criteria                     NEM/nem_alg.c:2764-2843
//   k_labels                       partition.py:206-231 on nem_exe.c:1682's "%5.3f"
//
// All of it is HBM-bound integer/bit work plus small fp64 epilogues; no tensor cores on purpose (K <= 32).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <climits>

namespace nemb200 {
namespace cg = cooperative_groups;

constexpr double kEps = 1e-20;  // NEM/nem_typ.h:65 EPSILON
// state words of a run (device): convergence word, status (NEMB200_STS_*), and the bookkeeping of the device-resident loop
enum { kStMaxdiff = 0, kStStatus = 1, kStActive = 2, kStIter = 3, kStConverged = 4, kStWords = 8 };
constexpr int kFuzzy = -1;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  // the matrix is read once per pass: keep it out of L1, let L2 decide
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// Parameter block (device memory, owned by a plan)
// ------------------------------------------------------------------------------------------------------------------
struct Params {
  int K, D, Ws, Dpad, free_disp;
  float* pk;        // [K]   proportions (float32 like ModelParaT.Prop_K)
  float* logpk32;   // [K]   (float) log((double)pk)           nem_alg.c:2350-2352
  double* logpk64;  // [K]
  float* mu;        // [K][D]   {0, 0.5, 1}
  float* eps;       // [K][D]
  float* eps_k;     // [K]   "sk_": the shared dispersion of each class
  uint32_t* m1;     // [K][Ws]  bit set where mu == 1
  uint32_t* valid;  // [K][Ws]  bit set where mu != 0.5 (a 0.5 centre never mismatches: nem_mod.c:656-657)
  double* L;        // sk_: [K]        log((1-e)/e);  skd: [K][Dpad]   (+inf where e <= 1e-20: null density on mismatch)
  double* B;        // [K]   sum_d log(1-e_kd) over columns with e > 1e-20
  int* has_half;    // [1]   any mu == 0.5
  int* status;      // [1]   NEMB200_STS_*
};

// ------------------------------------------------------------------------------------------------------------------
// Row-sharded multi-GPU exchange over NVLink peer memory (no collective library on the data path).
// Stepwise interface (host-polled driver, nem/sharded.py):
//  * k_peer_broadcast pushes the shard of log-densities a rank has just computed into the logf buffer of EVERY rank
//    (16-byte peer stores), which is the all-gather the replicated Gauss-Seidel sweep needs;
//  * k_mstep_finalize reads the column statistics of every rank (peer loads) and adds them in rank order, which is the
//    all-reduce of the M-step, fused into the consuming kernel (same order on every rank -> identical parameters);
//  * two flag barriers per iteration order those accesses: k_peer_signal_wait in front of k_mstep_finalize; the last
//    CTA of k_peer_broadcast signals the densities, and the sweep kernel's CTAs wait for them (peer_wait_lane).
// Device-resident loop (nemb200_nem_sharded): the M-step exchange is k_b_finalize_ll ({payload, epoch} words, no fence, no
// flag: see LLayout), the all-gather of the log-densities is the epilogue of k_b_density_ring.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct PeerOut {            // k_peer_broadcast: ptr[1..n-1] = the peers' buffers; row_off = first double of the shard
  double* ptr[kMaxPeers];
  int n;
  long long row_off;
};
struct PeerParams {         // column-sliced finalize of a row-sharded run: where a rank writes its column chunks' results
  float* mu[kMaxPeers]; float* eps[kMaxPeers]; double* L[kMaxPeers];
  uint32_t* m1[kMaxPeers]; uint32_t* valid[kMaxPeers];
  double* fin_partial[kMaxPeers]; int* has_half[kMaxPeers];
  int n, rank;
};
struct PeerStats {          // where k_mstep_finalize reads: the statistic blocks of the n ranks
  const int* iblock[kMaxPeers];     // cnt[K][Dpad] followed by nk_cnt[K]
  const double* dblock[kMaxPeers];  // fsum[K][Dpad] followed by nk_fz[K]
  int n;
};
struct FinSync {            // the flag barriers of a row-sharded M-step, folded into the finalize kernel (device-resident loop)
  int* const* sig_ptrs; const int* sig; int* err; int* ticket;
  int world, rank, epoch;
  unsigned long long* dbg;   // NEMB200_TRACE: %globaltimer stamps of the launch's phases (null otherwise)
};
__device__ __forceinline__ void dbg_stamp(unsigned long long* dbg, int slot) {
  if (dbg != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[slot] = t;
  }
}

// ---- "LL" exchange words for the row-sharded M-step ------------------------------------------------------------------
// A flag barrier over NVLink costs a system-scope fence on the writer (measured ~12 us with stores to 7 peers in flight)
// plus a dependent peer load round trip (~5.5 us) on the reader.  Here every value travels as an 8-byte word
// {payload, epoch} written with one store: the receiver polls ITS OWN memory until the epoch of the current iteration
// shows up, so there is no fence, no separate flag and no peer load -- one NVLink one-way trip per exchange step.
// fp64 payloads travel as two such words (low half, high half).  Bit 31 of the epoch word of a count says "the fp64
// sum of this column is zero and is not sent" (most columns once the run has few fuzzy rows).
struct LLayout {            // byte offsets inside every rank's LL region (identical on all ranks)
  unsigned long long in_c, in_f, nk_c, nk_f, out_m1, out_valid, out_part, out_L, out_eps, bytes;
};
__host__ __device__ inline LLayout ll_layout(int K, int D, int Dpad, int Ws, int fin_chunks, int free_disp) {
  LLayout l;
  unsigned long long o = 0;
  auto take = [&](unsigned long long n) { const unsigned long long at = o; o += (n + 255ull) & ~255ull; return at; };
  const unsigned long long KD = (unsigned long long)K * Dpad;
  l.in_c = take(8ull * kMaxPeers * KD);
  l.in_f = take(16ull * kMaxPeers * KD);
  l.nk_c = take(8ull * kMaxPeers * 32);
  l.nk_f = take(16ull * kMaxPeers * 32);
  l.out_m1 = take(8ull * K * Ws);
  l.out_valid = take(8ull * K * Ws);
  l.out_part = take(16ull * K * fin_chunks * 2);
  l.out_L = take(free_disp ? 16ull * KD : 0ull);
  l.out_eps = take(free_disp ? 8ull * K * D : 0ull);
  l.bytes = o;
  return l;
}
__device__ __forceinline__ void ll_put(uint2* p, uint32_t v, uint32_t e) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(e) : "memory");
}
__device__ __forceinline__ void ll_put_f64(uint4* p, double v, uint32_t e) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)b), "r"(e), "r"((uint32_t)(b >> 32)), "r"(e)
               : "memory");
}
__device__ __forceinline__ uint2 ll_peek(const uint2* p) {
  uint2 r;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint4 ll_peek16(const uint4* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
constexpr long long kLlSpins = 1LL << 24;   // bounded: a stalled peer shows up as an error, never as a hung GPU
// word of the current epoch (own memory); the epoch word comes back with its flag bit
__device__ __forceinline__ uint2 ll_get(const uint2* p, uint32_t e, int* err) {
  uint2 r = ll_peek(p);
  long long spins = 0;
  while ((r.y & 0x7fffffffu) != e) {
    if (++spins > kLlSpins) { atomicExch(err, 3); break; }
    r = ll_peek(p);
  }
  return r;
}
__device__ __forceinline__ double ll_get_f64(const uint4* p, uint32_t e, int* err) {
  uint4 r = ll_peek16(p);
  long long spins = 0;
  while (r.y != e || r.w != e) {
    if (++spins > kLlSpins) { atomicExch(err, 3); break; }
    r = ll_peek16(p);
  }
  return __longlong_as_double((long long)(((unsigned long long)r.z << 32) | r.x));
}

// push this rank's freshly computed shard of log-densities (contiguous rows of its own buffer) into the same rows of
// every peer's buffer: 16-byte coalesced peer stores, one peer per blockIdx.y.  With `ticket` != null the last CTA to
// finish (all stores fenced at system scope) also publishes the epoch in every rank's signal words: the "densities
// ready" signal of the exchange, without a launch of its own.
__global__ void __launch_bounds__(256)
k_peer_broadcast(const double* __restrict__ own, PeerOut out /*ptr[1..n-1] = peers*/, long long n_doubles /*shard size*/,
                 int* ticket, int* const* peer_sig, int phase, int rank, int epoch) {
  const int peer = blockIdx.y + 1;
  if (peer < out.n) {
    const long long off = out.row_off;                  // in doubles, even (K * row_off rounded by the caller)
    const double2* src = reinterpret_cast<const double2*>(own + off);
    double2* dst = reinterpret_cast<double2*>(out.ptr[peer] + off);
    const long long n2 = n_doubles / 2;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (long long)gridDim.x * blockDim.x) dst[t] = src[t];
    if ((n_doubles & 1) && blockIdx.x == 0 && threadIdx.x == 0) out.ptr[peer][off + n_doubles - 1] = own[off + n_doubles - 1];
  }
  if (ticket == nullptr) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = gridDim.x * gridDim.y;
    if (atomicAdd(ticket, 1) == total - 1) {
      *ticket = 0;
      __threadfence_system();
      for (int r = 0; r < out.n; r++) {
        volatile int* dst = peer_sig[r] + phase * kMaxPeers + rank;
        *dst = epoch;
      }
      __threadfence_system();
    }
  }
}

// every rank owns sig[2][kMaxPeers] ints: sig[phase][r] = last epoch rank r has completed for that phase
__global__ void k_peer_signal(int* const* peer_sig /*[n] device pointers to the ranks' sig arrays*/, int n, int phase, int rank,
                              int epoch) {
  if (threadIdx.x < n && blockIdx.x == 0) {
    __threadfence_system();
    volatile int* dst = peer_sig[threadIdx.x] + phase * kMaxPeers + rank;
    *dst = epoch;
    __threadfence_system();
  }
}
__device__ __forceinline__ void peer_wait_lane(const int* sig, int n, int phase, int epoch, int* error) {
  if (threadIdx.x < n) {
    const volatile int* src = sig + phase * kMaxPeers + threadIdx.x;
    long long spins = 0;
    while (*src < epoch) {
      __nanosleep(200);
      if (++spins > (1LL << 27)) { atomicExch(error, 2); break; }   // ~half a minute: a stalled peer, never a hung GPU
    }
    __threadfence_system();
  }
}
// publish this rank's epoch, then wait for every rank's: the flag barrier in front of k_mstep_finalize in one launch
__global__ void k_peer_signal_wait(int* const* peer_sig, const int* sig, int n, int phase, int rank, int epoch, int* error) {
  if (blockIdx.x != 0) return;
  if (threadIdx.x < n) {
    __threadfence_system();
    volatile int* dst = peer_sig[threadIdx.x] + phase * kMaxPeers + rank;
    *dst = epoch;
    __threadfence_system();
  }
  peer_wait_lane(sig, n, phase, epoch, error);
}
// ------------------------------------------------------------------------------------------------------------------
// E-step part 1: log densities.  A group of G lanes walks RB rows at once so that each class's centre words are
// fetched from shared memory once per RB matrix words.
// ------------------------------------------------------------------------------------------------------------------
// carry-save adder on 32 bit-columns at once: a + b + c = 2 * hi + lo
// Written as two explicit LOP3 (majority 0xE8, parity 0x96): left to the compiler, the mismatch XORs feeding b and c get
// folded into the adder as five 3-input LOP3 per adder instead of 2 + 2 (seen in the SASS; the ALU pipe is what bounds
// the density kernel once the loads are staged through shared memory).
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h, l;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(h) : "r"(a), "r"(b), "r"(c));
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(l) : "r"(a), "r"(b), "r"(c));
  hi = h; lo = l;
}

// Pipe balance (measured with ncu on B200): one POPC per word and class saturates the quarter-rate XU pipe (80 % XU at
// 53 % DRAM); a full carry-save tree moves the bottleneck to the ALU pipe instead.  Middle ground used here, per 4
// mismatch words and class: 2 carry-save adders (4 LOP3, ALU) folding into a running `ones` word + 2 POPC of the
// weight-2 carries (XU).  Hamming distance = 2 * sum popc(carries) + popc(ones).
//
// KX > 0: K is the compile-time constant KX (no per-class predicates); KX == 0: runtime K <= KB.
template <int KB, int KX, int RB, bool HALF, typename PlanePtr>
__device__ __forceinline__ void density_sk_body(const uint4* __restrict__ bits, long long N, int W4, int Krt, int G,
                                                PlanePtr m1, PlanePtr valid, const double* __restrict__ L,
                                                const double* __restrict__ B, double* __restrict__ logf) {
  const int K = KX > 0 ? KX : Krt;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int grp = lane / G;
  const int groups_per_warp = 32 / G;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long rows_per_warp = (long long)groups_per_warp * RB;

  for (long long base = warp_global * rows_per_warp; base < N; base += nwarps * rows_per_warp) {
    const long long r0 = base + (long long)grp * RB;
    const uint4* rowp[RB];   // running pointers (advanced by G per step): keeps the 64-bit row arithmetic out of the loop
#pragma unroll
    for (int r = 0; r < RB; r++) rowp[r] = bits + (r0 + r < N ? r0 + r : N - 1) * W4 + sub;  // clamp: tail rows re-read the last row
    int h2[RB][KB];
    uint32_t ones[RB][KB];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
      for (int k = 0; k < KB; k++) { h2[r][k] = 0; ones[r][k] = 0; }

    PlanePtr pm = m1 + sub, pv = valid + sub;
    // software pipeline: the matrix words of step t+1 are requested before the words of step t are consumed
    uint4 xn[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) { xn[r] = make_uint4(0, 0, 0, 0); if (sub < W4) { xn[r] = ld_stream(rowp[r]); rowp[r] += G; } }
#pragma unroll 1
    for (int c = sub; c < W4; c += G) {
      uint4 x[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) x[r] = xn[r];
      if (c + G < W4) {
#pragma unroll
        for (int r = 0; r < RB; r++) { xn[r] = ld_stream(rowp[r]); rowp[r] += G; }
      }
#pragma unroll
      for (int k = 0; k < KB; k++) {
        if (KX > 0 || k < K) {
          const uint4 a = pm[k * W4];
          uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
          if (HALF) v = pv[k * W4];
#pragma unroll
          for (int r = 0; r < RB; r++) {
            const uint32_t m0 = (x[r].x ^ a.x) & v.x, m1w = (x[r].y ^ a.y) & v.y;
            const uint32_t m2 = (x[r].z ^ a.z) & v.z, m3 = (x[r].w ^ a.w) & v.w;
            uint32_t c1, c2;
            csa(c1, ones[r][k], ones[r][k], m0, m1w);
            csa(c2, ones[r][k], ones[r][k], m2, m3);
            h2[r][k] += __popc(c1) + __popc(c2);
          }
        }
      }
      pm += G; pv += G;
    }
    int h[RB][KB];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
      for (int k = 0; k < KB; k++) h[r][k] = 2 * h2[r][k] + __popc(ones[r][k]);
    // reduce over the G lanes of the group
    for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int r = 0; r < RB; r++)
#pragma unroll
        for (int k = 0; k < KB; k++) h[r][k] += __shfl_xor_sync(0xffffffffu, h[r][k], off);
    }
    // lane `sub` writes classes sub, sub+G, ...
#pragma unroll
    for (int k = 0; k < KB; k++) {
      if (k < K && (k & (G - 1)) == sub) {
        const double Lk = L[k], Bk = B[k];
#pragma unroll
        for (int r = 0; r < RB; r++) {
          if (r0 + r < N) {
            // e_k <= 1e-20: L = +inf; a mismatch gives the reference's null density (nem_mod.c:663-665, :684-685)
            const double v = (h[r][k] == 0) ? Bk : (Bk - (double)h[r][k] * Lk);
            logf[(r0 + r) * K + k] = v;
          }
        }
      }
    }
  }
}

constexpr int kDensTpb = 128;

template <int KB, int KX, int RB>
__device__ __forceinline__ void
density_sk_cta(const uint4* __restrict__ bits, long long N, int W4, int K, int G, const uint4* __restrict__ m1g,
               const uint4* __restrict__ validg, const double* __restrict__ L, const double* __restrict__ B,
               const int* __restrict__ has_half, double* __restrict__ logf, int planes_in_smem) {
  extern __shared__ uint4 smem4[];
  const bool half = (*has_half != 0);   // any centre at 0.5 (exact median tie): needs the `valid` plane
  if (planes_in_smem) {
    const int n = K * W4;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      smem4[t] = m1g[t];
      if (half) smem4[n + t] = validg[t];
    }
    __syncthreads();
    const uint4* sm1 = smem4;            // shared-memory pointers: LDS, not generic loads
    const uint4* sval = smem4 + n;
    if (half) density_sk_body<KB, KX, RB, true>(bits, N, W4, K, G, sm1, sval, L, B, logf);
    else      density_sk_body<KB, KX, RB, false>(bits, N, W4, K, G, sm1, sval, L, B, logf);
  } else {
    if (half) density_sk_body<KB, KX, RB, true>(bits, N, W4, K, G, m1g, validg, L, B, logf);
    else      density_sk_body<KB, KX, RB, false>(bits, N, W4, K, G, m1g, validg, L, B, logf);
  }
}

template <int KB, int KX, int RB>
__global__ void __launch_bounds__(kDensTpb)
k_density_sk(const uint4* __restrict__ bits, long long N, int W4, int K, int G, const uint4* __restrict__ m1g,
             const uint4* __restrict__ validg, const double* __restrict__ L, const double* __restrict__ B,
             const int* __restrict__ has_half, double* __restrict__ logf, int planes_in_smem) {
  density_sk_cta<KB, KX, RB>(bits, N, W4, K, G, m1g, validg, L, B, has_half, logf, planes_in_smem);
}

// ---- shared-memory ring variant -------------------------------------------------------------------------------------
// Measured on B200 (tools/probes/stream_probe.cu): a read-only pass over the matrix needs >= 64 KB in flight per SM to
// reach the HBM ceiling (7.0 TB/s), and ncu shows the register-prefetch kernel above stalled on the long scoreboard.
// Here every thread keeps P steps of its own matrix words in flight with cp.async (LDGSTS, 16 B, L1 bypass) into
// private shared-memory slots -- no registers are held by loads in flight, no barrier is needed (a thread only reads
// what it copied itself), and the pipeline runs across row batches.  One fat CTA per SM shares one copy of the centre
// planes.  Arithmetic is identical to density_sk_body.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ const uint4* ptr_plus16(const uint4* base, uint32_t idx) {   // base + idx (IMAD.WIDE: FMA pipe)
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(idx), "l"((unsigned long long)base));
  return reinterpret_cast<const uint4*>(r);
}

// PC (with KB = KX = 2): the three-class case whose first centre is all ones and whose last centre is all zeros -- the
// persistent and the cloud class of nearly every iteration of a K = 3 run.  Their mismatch counts are D - popc(x) and
// popc(x), so only two accumulations run per matrix word instead of three (plane 0 is a zero plane: "mismatches" against
// it are the row's set bits; plane 1 is the middle class), which takes the kernel from ALU-bound to HBM-bound; the
// epilogue writes the three log-densities (L, B of the three classes; Dcols = number of genome columns).
template <int KB, int KX, int RB, int P, int TPB, bool HALF, bool PC = false>
__device__ __forceinline__ void density_sk_ring(const uint4* __restrict__ bits, long long N, int W4, int Krt, int G,
                                                const uint4* m1, const uint4* valid, uint4* ring /*[P][RB][TPB]*/,
                                                const double* __restrict__ L, const double* __restrict__ B,
                                                double* __restrict__ logf, int Kstride /*classes per row of logf*/, int Dcols = 0) {
  const int K = KX > 0 ? KX : Krt;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int grp = lane / G;
  const long long warp_global = (long long)blockIdx.x * (TPB >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (TPB >> 5);
  const long long rows_per_warp = (long long)(32 / G) * RB;
  const long long stride_rows = nwarps * rows_per_warp;
  const int nsteps = (W4 + G - 1) / G;
  const int my_steps = (W4 - sub + G - 1) / G;   // steps in which this lane has a matrix word (0 if sub >= W4)
  constexpr uint32_t kStage = RB * TPB * 16;

  const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(ring) + threadIdx.x * 16;
  const char* myslot = reinterpret_cast<const char*>(ring + threadIdx.x);

  // issue cursor: runs P-1 items (one item = RB x 16 B of this thread) ahead of the consume cursor, across row batches
  long long ir0 = warp_global * rows_per_warp + (long long)grp * RB;
  const uint4* ibr[RB];
  int isteps = 0;       // my_steps while the cursor is on an existing batch, 0 afterwards
  auto set_batch = [&]() {
    isteps = ir0 < N ? my_steps : 0;
#pragma unroll
    for (int r = 0; r < RB; r++) ibr[r] = bits + (ir0 + r < N ? ir0 + r : N - 1) * W4 + sub;   // tail rows re-read the last row
  };
  set_batch();
  uint32_t icol = 0;
  int istep = 0;
  int sj = 0;           // stage the next issue goes to; the item consumed in the same iteration sits in stage sj + 1 (mod P)
  auto issue = [&]() {
    if (istep < isteps) {
      const uint32_t dst = slot0 + sj * kStage;
#pragma unroll
      for (int r = 0; r < RB; r++) cp_async16(dst + r * (TPB * 16), ptr_plus16(ibr[r], icol));
    }
    cp_async_commit();
    sj = (sj + 1 == P) ? 0 : sj + 1;
    icol += G;
    if (++istep == nsteps) {
      istep = 0; icol = 0;
      ir0 += stride_rows;
      set_batch();
    }
  };
#pragma unroll
  for (int s = 0; s < P - 1; s++) issue();

  for (long long base = warp_global * rows_per_warp; base < N; base += stride_rows) {
    const long long r0 = base + (long long)grp * RB;
    int h2[RB][KB];
    uint32_t ones[RB][KB];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
      for (int k = 0; k < KB; k++) { h2[r][k] = 0; ones[r][k] = 0; }
    const uint4 *pm = m1 + sub, *pv = valid + sub;
#pragma unroll 1
    for (int step = 0; step < nsteps; step++) {
      issue();                                // fills stage sj-1; afterwards sj is the stage of the item consumed now
      cp_async_wait<P - 1>();
      if (step < my_steps) {
        const char* st = myslot + sj * kStage;
        uint4 x[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) x[r] = *reinterpret_cast<const uint4*>(st + r * (TPB * 16));
#pragma unroll
        for (int k = 0; k < KB; k++) {
          if (KX > 0 || k < K) {
            const uint4 a = pm[k * W4];
            uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (HALF) v = pv[k * W4];
#pragma unroll
            for (int r = 0; r < RB; r++) {
              const uint32_t m0 = (x[r].x ^ a.x) & v.x, m1w = (x[r].y ^ a.y) & v.y;
              const uint32_t m2 = (x[r].z ^ a.z) & v.z, m3 = (x[r].w ^ a.w) & v.w;
              uint32_t c1, c2;
              csa(c1, ones[r][k], ones[r][k], m0, m1w);
              csa(c2, ones[r][k], ones[r][k], m2, m3);
              h2[r][k] += __popc(c1) + __popc(c2);
            }
          }
        }
      }
      pm += G; pv += G;
    }
    int h[RB][KB];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
      for (int k = 0; k < KB; k++) h[r][k] = 2 * h2[r][k] + __popc(ones[r][k]);
    for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int r = 0; r < RB; r++)
#pragma unroll
        for (int k = 0; k < KB; k++) h[r][k] += __shfl_xor_sync(0xffffffffu, h[r][k], off);
    }
    if (PC) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if ((k & (G - 1)) == sub) {
          const double Lk = L[k], Bk = B[k];
#pragma unroll
          for (int r = 0; r < RB; r++) {
            if (r0 + r < N) {
              const int hk = k == 0 ? Dcols - h[r][0] : (k == 1 ? h[r][KB - 1] : h[r][0]);
              logf[(r0 + r) * Kstride + k] = (hk == 0) ? Bk : (Bk - (double)hk * Lk);
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KB; k++) {
        if (k < K && (k & (G - 1)) == sub) {
          const double Lk = L[k], Bk = B[k];
#pragma unroll
          for (int r = 0; r < RB; r++) {
            if (r0 + r < N) {
              const double v = (h[r][k] == 0) ? Bk : (Bk - (double)h[r][k] * Lk);   // see density_sk_body
              logf[(r0 + r) * Kstride + k] = v;
            }
          }
        }
      }
    }
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ uint32_t columns_mask(int w, int D) {   // bits of 32-column word w that are genome columns
  const int n = D - w * 32;
  return n >= 32 ? 0xffffffffu : (n > 0 ? ((1u << n) - 1u) : 0u);
}

template <int KB, int KX, int RB, int P, int TPB>
__device__ __forceinline__ void
density_sk_ring_cta(const uint4* __restrict__ bits, long long N, int W4, int K, int G, const uint4* __restrict__ m1g,
                    const uint4* __restrict__ validg, const double* __restrict__ L, const double* __restrict__ B,
                    const int* __restrict__ has_half, double* __restrict__ logf, int Kstride, int D) {
  // K classes starting at the pointers given (a launch may cover a chunk of the model's classes); logf rows hold Kstride
  extern __shared__ uint4 smem4[];
  const bool half = (*has_half != 0);
  uint4* ring = smem4;                       // [P][RB][TPB]
  uint4* sm1 = smem4 + P * RB * TPB;         // [K][W4]
  uint4* sval = sm1 + K * W4;                // [K][W4]
  const int n = K * W4;
  for (int t = threadIdx.x; t < n; t += TPB) {
    sm1[t] = m1g[t];
    if (half) sval[t] = validg[t];
  }
  if (KX == 3 && Kstride == 3 && D > 0) {
    // first centre all ones, last centre all zeros, no centre at 0.5?  (block-uniform decision)
    bool mine = !half;
    for (int t = threadIdx.x; t < W4 && mine; t += TPB) {
      const uint4 a = m1g[t], c = m1g[2 * W4 + t];
      mine = a.x == columns_mask(4 * t, D) && a.y == columns_mask(4 * t + 1, D) && a.z == columns_mask(4 * t + 2, D) &&
             a.w == columns_mask(4 * t + 3, D) && (c.x | c.y | c.z | c.w) == 0u;
    }
    if (__syncthreads_and(mine)) {
      for (int t = threadIdx.x; t < W4; t += TPB) sm1[t] = make_uint4(0u, 0u, 0u, 0u);   // plane 0: the row's own bits
      __syncthreads();
      density_sk_ring<2, 2, RB, P, TPB, false, true>(bits, N, W4, 2, G, sm1, sval, ring, L, B, logf, Kstride, D);
      return;
    }
  }
  __syncthreads();
  if (half) density_sk_ring<KB, KX, RB, P, TPB, true>(bits, N, W4, K, G, sm1, sval, ring, L, B, logf, Kstride);
  else      density_sk_ring<KB, KX, RB, P, TPB, false>(bits, N, W4, K, G, sm1, sval, ring, L, B, logf, Kstride);
}

template <int KB, int KX, int RB, int P, int TPB>
__global__ void __launch_bounds__(TPB, 1)
k_density_sk_ring(const uint4* __restrict__ bits, long long N, int W4, int K, int G, const uint4* __restrict__ m1g,
                  const uint4* __restrict__ validg, const double* __restrict__ L, const double* __restrict__ B,
                  const int* __restrict__ has_half, double* __restrict__ logf, int Kstride, int D) {
  density_sk_ring_cta<KB, KX, RB, P, TPB>(bits, N, W4, K, G, m1g, validg, L, B, has_half, logf, Kstride, D);
}

// free dispersion ("skd"): per-column log-odds, logf_ik = B_k - sum over mismatching columns d of L_kd.  The sum walks
// the non-zero NIBBLES of the mismatch word -- or of its complement, subtracted from the word's total, whichever has
// fewer set bits -- with one lookup per nibble in a table of the 16 partial sums of every nibble (k_ltab, rebuilt after
// each M-step: K x D/4 x 16 doubles, 48 KB for a 500-genome chunk instance).  Rows that match their class or mismatch
// almost everywhere (the common cases) cost 0-2 lookups per word, mixed rows at most 8.  +inf entries (e_kd <= 1e-20)
// propagate to the reference's null density; a word whose total is infinite is summed directly.
// SMEM: the CTA first copies the instance's tables (K x Ws x 128 doubles: 48 KB at K = 3, 500 genomes), plane words and word
// totals into dynamic shared memory and looks them up there.  ncu on the global-memory version at 16 x (100 000 x 500): the
// kernel is bound by L1 sector throughput -- every lookup instruction touches 32 different sectors -- not by latency; shared
// memory serves the same scattered 8-byte reads at a few bank conflicts each.  Same tables, same order of additions:
// identical sums.
template <int KB, bool SMEM = false>
__device__ __forceinline__ void
density_skd_body(const uint32_t* __restrict__ bits, long long N, int Ws, int K, int G,
                 const uint32_t* __restrict__ m1, const uint32_t* __restrict__ valid, const double* __restrict__ Ltab /*[K][Ws*8][16]*/,
                 const double* __restrict__ Lword /*[K][Ws]*/, const double* __restrict__ B, double* __restrict__ logf) {
  extern __shared__ double s_skd[];
  const double* tabs = Ltab;
  const double* lwords = Lword;
  const uint32_t* pm1 = m1;
  const uint32_t* pvalid = valid;
  if (SMEM) {
    const int kw = K * Ws;
    double* s_tab = s_skd;                                       // [K][Ws][128]
    double* s_lw = s_tab + (size_t)kw * 128;                     // [K][Ws]
    uint32_t* s_m1 = reinterpret_cast<uint32_t*>(s_lw + kw);     // [K][Ws]
    uint32_t* s_va = s_m1 + kw;
    const double2* src2 = reinterpret_cast<const double2*>(Ltab);
    double2* dst2 = reinterpret_cast<double2*>(s_tab);
    for (int t = threadIdx.x; t < kw * 64; t += blockDim.x) dst2[t] = __ldg(src2 + t);
    for (int t = threadIdx.x; t < kw; t += blockDim.x) { s_lw[t] = Lword[t]; s_m1[t] = m1[t]; s_va[t] = valid[t]; }
    __syncthreads();
    tabs = s_tab; lwords = s_lw; pm1 = s_m1; pvalid = s_va;
  }
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int grp = lane / G;
  const int groups_per_warp = 32 / G;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = warp_global * groups_per_warp; base < N; base += nwarps * groups_per_warp) {
    const long long row = base + grp;
    double s[KB];
#pragma unroll
    for (int k = 0; k < KB; k++) s[k] = 0.0;
    if (row < N) {
      for (int c = sub; c < Ws; c += G) {
        const uint32_t x = ld_stream32(bits + row * Ws + c);
#pragma unroll
        for (int k = 0; k < KB; k++) {
          if (k < K) {
            const uint32_t mm = (x ^ pm1[k * Ws + c]) & pvalid[k * Ws + c];   // mismatching columns
            const double* tab = tabs + ((size_t)k * Ws + c) * 128;
            const double lw = lwords[(size_t)k * Ws + c];      // all 32 columns of the word (padding and mu = 0.5 included)
            const bool direct = __popc(mm) <= 16 || isinf(lw);
            const uint32_t it = direct ? mm : ~mm;              // the complement holds matches, mu = 0.5 and padding columns
            // the non-zero nibbles in ascending order, as predicated lookups: eight independent loads and one chain of adds
            // per class instead of a data-dependent loop (lanes of a warp hold different words; a zero nibble adds nothing)
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const uint32_t nib = (it >> (4 * q)) & 15u;
              if (nib != 0u) acc += SMEM ? tab[q * 16 + nib] : __ldg(tab + q * 16 + nib);
            }
            if (mm != 0u) s[k] += direct ? acc : lw - acc;
          }
        }
      }
    }
    for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < KB; k++) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    }
#pragma unroll
    for (int k = 0; k < KB; k++)
      if (k < K && (k & (G - 1)) == sub && row < N) logf[row * K + k] = B[k] - s[k];
  }
}

template <int KB>
__global__ void __launch_bounds__(256)
k_density_skd(const uint32_t* __restrict__ bits, long long N, int Ws, int K, int G, const uint32_t* __restrict__ m1,
              const uint32_t* __restrict__ valid, const double* __restrict__ Ltab, const double* __restrict__ Lword,
              const double* __restrict__ B, double* __restrict__ logf) {
  density_skd_body<KB>(bits, N, Ws, K, G, m1, valid, Ltab, Lword, B, logf);
}

// ------------------------------------------------------------------------------------------------------------------
// E-step part 2: posterior sweep.  One warp per row, lane k = class k.  Rows of one level are independent; levels are
// separated by a grid-wide barrier (cooperative launch) so that neighbours with a smaller index are read from T_new
// and the others from T_old -- exactly what the reference's in-place sweep over rows 0..N-1 sees.
// ------------------------------------------------------------------------------------------------------------------
struct SweepArgs {
  long long N;
  int K;
  const double* logf;
  const float* pk;
  const double* logpk64;
  const float* T_old;
  float* T_new;
  const int* rowptr;
  const int* col;
  const float* w;
  const int* level_ptr;  // may be null: single level over all rows
  const int* level_rows;
  int n_levels;
  float beta;
  int logspace;
  int jacobi;
  unsigned* maxdiff_bits;
  const int4* ell_hdr;   // [N] {row, degree, CSR offset, 0} in schedule order (level-scheduled sweeps)
  const int2* ell_nb;    // [N][kEllW] (neighbour, weight bits), zero padded
  const int* wait_sig;   // row-sharded runs: the rank's signal words; the sweep starts when every rank's densities are in
  int wait_n, wait_epoch;
  int* wait_err;
  unsigned long long* stamps;   // profiling builds (-DNEMB200_SWEEP_STAMPS): per CTA, %globaltimer at level start / end
  // speculative sweep (see sweep_spec): per-row "changed in the previous / this round" marks and the sweep's bookkeeping
  int* chg;                     // [2][N], may be null: level-scheduled sweep only
  int* spec_words;              // [8]: 0..2 rotating "any row changed this round" flags, 3..4 rows changed by the last two sweeps
  int spec_limit;               // speculative when the previous sweep changed at most this many rows
  int spec_parity;              // this sweep adds its count to spec_words[3 + parity] and reads the other one
};

// One group of GS lanes (GS = power of two >= K) updates one row; lane `sub` of the group owns class `sub`.
//
// The sweep is latency, not bandwidth (timed with %globaltimer stamps per level: ~2 us of grid barrier plus 2-4 us of
// dependent L2 round trips per level, whatever the level's size).  Per row the chain is schedule -> rowptr -> (col, w)
// -> neighbour posteriors, and only the last link depends on what other rows wrote.  So
//  * the graph is copied once per (graph, schedule) into schedule order with the first 8 neighbours inline
//    (k_sweep_ell_build: header {row, degree, CSR offset} + 8 x (j, w) per slot), which makes the graph side of a row
//    ONE round trip, its address known from the slot number alone;
//  * that round trip is taken BEFORE the barrier that precedes the row's level (SweepPre, spread over the lanes of the
//    group: lane s keeps neighbours s, s + GS, ... and hands them out by shuffle);
//  * after the barrier the neighbour posteriors, the row's log-densities and its old posterior are loaded together,
//    and the neighbour terms are added in list order (float, no FMA: nem_alg.c:2954-2962).
constexpr int kEllW = 8;          // neighbours held inline per slot; longer lists continue in the CSR arrays
template <int GS> struct SweepCfg {
  static constexpr int NS = GS >= kEllW ? 1 : kEllW / GS;   // neighbour slots per lane
};
template <int GS> struct SweepPre {
  long long i;
  int n0, deg;
  int j[SweepCfg<GS>::NS];
  float w[SweepCfg<GS>::NS];
  bool valid;
};

__global__ void __launch_bounds__(256)
k_sweep_ell_build(long long N, const int* __restrict__ level_rows, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const float* __restrict__ w, int4* __restrict__ hdr, int2* __restrict__ nb) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const int i = level_rows[idx];
  const int n0 = rowptr[i], deg = rowptr[i + 1] - n0;
  hdr[idx] = make_int4(i, deg, n0, 0);
#pragma unroll
  for (int u = 0; u < kEllW; u++)
    nb[idx * kEllW + u] = u < deg ? make_int2(col[n0 + u], __float_as_int(w[n0 + u])) : make_int2(0, 0);
}

// graph side of the row in schedule slot `idx` (ELL copy) ...
template <int GS>
__device__ __forceinline__ SweepPre<GS> sweep_preload_slot(const SweepArgs& a, long long idx, bool valid, int sub) {
  constexpr int NS = SweepCfg<GS>::NS;
  SweepPre<GS> p;
  p.valid = valid;
  const long long slot = valid ? idx : 0;
  const int4 h = a.ell_hdr[slot];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int u = s * GS + sub;
    int2 e = make_int2(0, 0);
    if (u < kEllW) e = a.ell_nb[slot * kEllW + u];        // address independent of the header: one round trip in all
    p.j[s] = e.x; p.w[s] = __int_as_float(e.y);
  }
  p.i = valid ? h.x : 0;
  p.deg = (valid && a.beta != 0.0f) ? h.y : 0;
  p.n0 = h.z;
  return p;
}
// ... or of row i straight from the CSR arrays (single-level sweeps, no ELL copy)
template <int GS>
__device__ __forceinline__ SweepPre<GS> sweep_preload_csr(const SweepArgs& a, long long i, bool valid, int sub) {
  constexpr int NS = SweepCfg<GS>::NS;
  SweepPre<GS> p;
  p.valid = valid;
  p.i = valid ? i : 0;
  p.n0 = 0; p.deg = 0;
#pragma unroll
  for (int s = 0; s < NS; s++) { p.j[s] = 0; p.w[s] = 0.0f; }
  if (valid && a.rowptr != nullptr && a.beta != 0.0f) {
    p.n0 = a.rowptr[p.i];
    p.deg = a.rowptr[p.i + 1] - p.n0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int u = s * GS + sub;
      if (u < p.deg && u < kEllW) { p.j[s] = a.col[p.n0 + u]; p.w[s] = a.w[p.n0 + u]; }
    }
  }
  return p;
}

// MODE 0: the row's update as part of the in-order sweep (neighbours j < i from T_new); the posterior change against T_old
//         goes into wmax.  Returns whether the row's posterior differs from T_old (any class).
// MODE 1: speculative first round: every neighbour from T_old.  Returns whether the row's posterior differs from T_old.
// MODE 2: speculative repair round: neighbours j < i from T_new.  Returns whether it differs from the row's previous T_new.
template <int GS, int MODE = 0>
__device__ __forceinline__ bool sweep_finish(const SweepArgs& a, const SweepPre<GS>& p, int sub, float& wmax) {
  const int K = a.K;
  const int k = sub < K ? sub : K - 1;
  const long long i = p.i;
  const double lf = a.logf[(size_t)i * K + k];
  const float told = MODE == 2 ? __ldcg(a.T_new + (size_t)i * K + k) : a.T_old[(size_t)i * K + k];
  const bool all_old = a.jacobi || MODE == 1;
  float ctx = 0.0f;
  // deg is the same in all lanes of the group; groups of one warp may differ, the shuffles below are group-local and
  // executed by every lane (invalid rows have deg = 0 but still take part)
  const int dmax = min(kEllW, __reduce_max_sync(0xffffffffu, p.deg));
  if (dmax > 0) {
    float c[kEllW], wt[kEllW];
#pragma unroll
    for (int u = 0; u < kEllW; u++) {
      c[u] = 0.0f; wt[u] = 0.0f;
      if (u < dmax) {                                   // warp-uniform
        const int jj = __shfl_sync(0xffffffffu, p.j[u / GS], u % GS, GS);
        wt[u] = __shfl_sync(0xffffffffu, p.w[u / GS], u % GS, GS);
        if (u < p.deg) {
          const float* src = (!all_old && jj < i) ? a.T_new : a.T_old;
          c[u] = __ldcg(src + (size_t)jj * K + k);      // L2 read: T_new rows are written by other SMs in this launch
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kEllW; u++)
      if (u < p.deg) ctx = __fadd_rn(ctx, __fmul_rn(wt[u], c[u]));
    for (int n = p.n0 + kEllW; n < p.n0 + p.deg; n++) {  // rows with more than kEllW neighbours
      const int j = a.col[n];
      const float* src = (!all_old && j < i) ? a.T_new : a.T_old;
      ctx = __fadd_rn(ctx, __fmul_rn(a.w[n], __ldcg(src + (size_t)j * K + k)));
    }
  }
  float cnew;
  if (!a.logspace) {
    // reference arithmetic: nem_mod.c:677-679, nem_alg.c:2367, :2665-2698
    const float lf32 = (float)lf;
    const double f = (lf == -INFINITY) ? 0.0 : exp((double)lf32);
    double num = (double)a.pk[k] * f;
    num = num * exp((double)a.beta * (double)ctx);
    if (sub >= K) num = 0.0;
    double cum = 0.0;
    for (int kk = 0; kk < K; kk++) cum = cum + __shfl_sync(0xffffffffu, num, kk, GS);  // sequential, like the reference
    if (cum > 0) {
      if (cum > kEps) { const double invZ = 1 / cum; cnew = (float)(invZ * num); }
      else { const double invZ = 1 / (cum / kEps); cnew = (float)(invZ * (num / kEps)); }
    } else {
      cnew = (float)(1.0 / K);
    }
  } else {
    double av = a.logpk64[k] + lf + (double)a.beta * (double)ctx;
    if (sub >= K) av = -INFINITY;
    double m = av;
#pragma unroll
    for (int off = GS >> 1; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off, GS));
    const double e = (sub < K && m > -INFINITY) ? exp(av - m) : 0.0;
    double sum = 0.0;
    for (int kk = 0; kk < K; kk++) sum = sum + __shfl_sync(0xffffffffu, e, kk, GS);
    cnew = (m == -INFINITY) ? (float)(1.0 / K) : (float)(e / sum);
  }
  bool differs = false;
  if (p.valid && sub < K) {
    __stcg(a.T_new + (size_t)i * K + sub, cnew);
    differs = __float_as_uint(cnew) != __float_as_uint(told);
    if (MODE == 0) {
      float dif = cnew - told;
      if (dif < 0) dif = -dif;
      wmax = fmaxf(wmax, dif);
    }
  }
  // any class of the row (the GS lanes of the group)
  const unsigned m = __ballot_sync(0xffffffffu, differs);
  const int lane = threadIdx.x & 31;
  return ((m >> (lane & ~(GS - 1))) & (GS == 32 ? 0xffffffffu : ((1u << GS) - 1u))) != 0u;
}

// ---- speculative sweep -------------------------------------------------------------------------------------------------
// In a converging run almost every posterior is saturated (exactly one-hot in float32) and does not change from one
// iteration to the next, although every row formally depends on its lower-indexed neighbours.  The level-scheduled sweep
// still pays a grid barrier and two dependent L2 round trips for each of the graph's levels (65-73 us for 200 000 rows in
// 12 levels).  Speculation: round 0 updates ALL rows at once from the old posteriors (one level's worth of latency);
// a row whose lower-indexed neighbour came out different from what the row used is repaired in the next round with the
// neighbours' current values, and so on until a round changes nothing.  Dependencies only point to lower indices, so the
// fixed point is unique and is the in-order sweep's result, bit for bit (a row's last repair evaluates the same
// expression on the same inputs); rounds are bounded by the schedule depth, and in a converging run two or three rounds
// that touch a handful of rows replace the twelve levels.  A racing read of a posterior that is being rewritten in the
// same round is harmless: the writer's change mark sends the reader into the next round.
template <int GS, class Bar>
__device__ __forceinline__ void sweep_spec(const SweepArgs& a, const Bar& bar, float& wmax, int& n_changed) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (GS - 1), grp = lane / GS;
  constexpr int RPW = 32 / GS;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gtotal = nwarps * RPW;
  int* chg_prev = a.chg;
  int* chg_cur = a.chg + a.N;
  volatile int* words = a.spec_words;
  // round 0: every row from the old posteriors
  {
    bool any = false;
    long long idx0 = warp_global * RPW;
    if (idx0 < a.N) {
      SweepPre<GS> cur = sweep_preload_slot<GS>(a, idx0 + grp, idx0 + grp < a.N, sub);
      for (; idx0 < a.N; idx0 += gtotal) {
        const long long nx = idx0 + gtotal;
        SweepPre<GS> nxt = cur;
        if (nx < a.N) nxt = sweep_preload_slot<GS>(a, nx + grp, nx + grp < a.N, sub);
        const bool d = sweep_finish<GS, 1>(a, cur, sub, wmax);
        if (cur.valid && sub == 0) __stcg(chg_prev + cur.i, d ? 1 : 0);
        any = any || (cur.valid && d);
        cur = nxt;
      }
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr((int*)&words[0], 1);
  }
  bar.sync();
  // repair rounds
  for (int r = 1;; r++) {
    const int prev_flag = __ldcg((const int*)&words[(r - 1) % 3]);
    if (blockIdx.x == 0 && threadIdx.x == 0) words[(r + 1) % 3] = 0;     // used again two rounds from now
    if (!prev_flag) break;
    bool any = false;
    for (long long idx0 = warp_global * RPW; idx0 < a.N; idx0 += gtotal) {
      const long long idx = idx0 + grp;
      SweepPre<GS> p = sweep_preload_slot<GS>(a, idx, idx < a.N, sub);
      // dirty: a lower-indexed neighbour changed in the previous round (the lanes of the group hold the inline neighbours)
      bool mine = false;
#pragma unroll
      for (int q = 0; q < SweepCfg<GS>::NS; q++) {
        const int u = q * GS + sub;
        if (p.valid && u < p.deg && u < kEllW && p.j[q] < p.i) mine = mine || (__ldcg(chg_prev + p.j[q]) != 0);
      }
      if (p.valid && sub == 0)
        for (int n = p.n0 + kEllW; n < p.n0 + p.deg; n++) { const int j = a.col[n]; if (j < p.i) mine = mine || (__ldcg(chg_prev + j) != 0); }
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      const bool dirty = ((bal >> (lane & ~(GS - 1))) & (GS == 32 ? 0xffffffffu : ((1u << GS) - 1u))) != 0u;
      bool d = false;
      if (bal != 0u) {                                                    // warp-uniform: someone in the warp repairs a row
        SweepPre<GS> q = p;
        q.valid = p.valid && dirty;
        if (!q.valid) q.deg = 0;
        d = sweep_finish<GS, 2>(a, q, sub, wmax) && q.valid;
      }
      if (p.valid && sub == 0) __stcg(chg_cur + p.i, d ? 1 : 0);
      any = any || d;
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr((int*)&words[r % 3], 1);
    int* t = chg_prev; chg_prev = chg_cur; chg_cur = t;
    bar.sync();
  }
  // the sweep's change against the previous posteriors (nem_alg.c:2160-2173) and how many rows it changed
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.N; t += (long long)gridDim.x * blockDim.x) {
    bool differs = false;
    for (int k = 0; k < a.K; k++) {
      const float n = __ldcg(a.T_new + (size_t)t * a.K + k), o = a.T_old[(size_t)t * a.K + k];
      wmax = fmaxf(wmax, fabsf(n - o));
      differs = differs || (__float_as_uint(n) != __float_as_uint(o));
    }
    n_changed += differs ? 1 : 0;
  }
}

// Barrier between the levels of one sweep: the whole grid (single instance, cooperative launch) ...
struct GridBar {
  __device__ __forceinline__ void sync() const { __threadfence(); cg::this_grid().sync(); }
};
// ... or the CTAs of ONE instance of a batched launch (blockIdx.z = instance; all CTAs of the launch are co-resident --
// cooperative launch -- so spinning is safe).  Monotonic counter: every CTA adds one per barrier and waits until the
// count reaches the next multiple of the instance's CTA count; the counter is reset between launches by k_b_poll.
struct InstBar {
  unsigned* ctr;
  unsigned nctas;
  __device__ __forceinline__ void sync() const {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned t = atomicAdd(ctr, 1u);
      const unsigned target = (t / nctas + 1u) * nctas;
      unsigned v;
      do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
      __threadfence();
    }
    __syncthreads();
  }
};

// Sweep without a neighbour term (beta = 0: the K-selection sweep of partition.py:482-514 and the blind start sweep of every
// run, nem_alg.c:2043-2046): rows are independent, so a THREAD owns a row and walks its classes -- every lane busy and no
// shuffles, where the lane-per-class layout spends K dependent double shuffles per row on the reference's sequential sum
// (at K = 20: 12 of 32 lanes idle).  The per-class terms wait in shared memory ([class][thread]: conflict-free) for the
// row's sum, which keeps the kernel at ~50 registers (held in registers they cost 154 and the occupancy with them: 122 us
// per sweep of 4 x 50 000 rows at K = 17..20, ncu: 2 warps per scheduler, issue slots 20 % busy).  Same expressions in the
// same order as sweep_finish with ctx = 0 (x * exp(0) = x and x + 0.0 = x exactly): identical bits.
// Dynamic shared memory: K * blockDim.x doubles.
template <int KB>
__device__ __forceinline__ void sweep_independent_body(const SweepArgs& a) {
  extern __shared__ double s_terms[];
  const int K = a.K;
  double* mine = s_terms + threadIdx.x;
  const int pitch = blockDim.x;
  float wmax = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
    const double* lfp = a.logf + (size_t)i * K;
    double cum = 0.0, m = -INFINITY;
    if (!a.logspace) {
#pragma unroll 4
      for (int k = 0; k < K; k++) {
        const double lf = lfp[k];
        const float lf32 = (float)lf;
        const double f = (lf == -INFINITY) ? 0.0 : exp((double)lf32);
        const double v = (double)a.pk[k] * f;
        mine[k * pitch] = v;
        cum = cum + v;
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < K; k++) {
        const double av = a.logpk64[k] + lfp[k];
        mine[k * pitch] = av;
        m = fmax(m, av);
      }
#pragma unroll 4
      for (int k = 0; k < K; k++) {
        const double e = (m > -INFINITY) ? exp(mine[k * pitch] - m) : 0.0;
        mine[k * pitch] = e;
        cum = cum + e;
      }
    }
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      const double v = mine[k * pitch];
      float cnew;
      if (!a.logspace) {
        if (cum > 0) {
          if (cum > kEps) { const double invZ = 1 / cum; cnew = (float)(invZ * v); }
          else { const double invZ = 1 / (cum / kEps); cnew = (float)(invZ * (v / kEps)); }
        } else {
          cnew = (float)(1.0 / K);
        }
      } else {
        cnew = (m == -INFINITY) ? (float)(1.0 / K) : (float)(v / cum);
      }
      const float told = a.T_old[(size_t)i * K + k];
      __stcg(a.T_new + (size_t)i * K + k, cnew);
      float dif = cnew - told;
      if (dif < 0) dif = -dif;
      wmax = fmaxf(wmax, dif);
    }
  }
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if ((threadIdx.x & 31) == 0 && wmax > 0.0f) atomicMax(a.maxdiff_bits, __float_as_uint(wmax));
}

constexpr int kSpecVeryDeep = 64;

template <bool COOP, int GS, class Bar>
__device__ __forceinline__ void sweep_body(const SweepArgs& a, const Bar& bar) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (GS - 1), grp = lane / GS;
  constexpr int RPW = 32 / GS;  // rows per warp
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float wmax = 0.0f;
  int n_changed = 0;            // rows whose posterior differs from T_old (counted by the lane with sub == 0)
  auto fin = [&](const SweepPre<GS>& p) { if (sweep_finish<GS>(a, p, sub, wmax) && p.valid && sub == 0) n_changed++; };
  if (a.wait_sig != nullptr) {     // every CTA checks the flags itself (phase 1 = densities): no wait kernel in front
    peer_wait_lane(a.wait_sig, a.wait_n, 1, a.wait_epoch, a.wait_err);
    __syncthreads();
  }
  if (!COOP && (a.beta == 0.0f || a.rowptr == nullptr) && a.chg == nullptr) {   // no neighbour term: a thread per row
    sweep_independent_body<GS>(a);
    return;
  }
  bool spec = false;
  // speculation pays where the level-scheduled sweep is latency (few passes of the grid per level); a batch whose instances
  // each own a small share of the SMs is throughput-bound, and the repair rounds' scans of all rows would only add to it
  // (a schedule of kSpecVeryDeep levels or more is the exception: a barrier and an underfilled pass per level cost more
  //  than a few rounds over all rows -- a family order that follows the genomes has hundreds of thin levels)
  if (COOP && a.level_ptr != nullptr && a.chg != nullptr && a.n_levels > 3 && (a.N <= 8 * nwarps * RPW || a.n_levels >= kSpecVeryDeep))
    spec = __ldcg(a.spec_words + 3 + (a.spec_parity ^ 1)) <= a.spec_limit;     // the previous sweep changed few rows (grid-uniform)
  if (spec) {
    sweep_spec<GS>(a, bar, wmax, n_changed);
  } else if (a.level_ptr == nullptr) {
    for (long long i0 = warp_global * RPW; i0 < a.N; i0 += nwarps * RPW) {
      const long long i = i0 + grp;
      const SweepPre<GS> p = sweep_preload_csr<GS>(a, i, i < a.N, sub);
      fin(p);
    }
  } else {
    // Levels are separated by grid-wide barriers; a run of consecutive small levels (deep, thin tails of the schedule)
    // is instead walked by CTA 0 alone with block barriers, so the run costs one grid barrier instead of one per level.
    constexpr int kSmall = 256;
    const int block_warp = threadIdx.x >> 5, block_warps = blockDim.x >> 5;
    const long long gidx = warp_global * RPW + grp, gtotal = nwarps * RPW;   // this lane group among all groups
    SweepPre<GS> p0, p1;          // the group's first two rows of the coming grid-wide level, fetched before its barrier
    bool have_pre = false;
    int L = 0;
#ifdef NEMB200_SWEEP_STAMPS
    auto stamp = [&](int slot) {
      if (a.stamps && threadIdx.x == 0 && slot < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.stamps[blockIdx.x * 64 + slot] = t;
      }
    };
#else
    auto stamp = [](int) {};
#endif
    stamp(0);
    int b = a.level_ptr[0], e = a.level_ptr[1];
    while (L < a.n_levels) {
      int e_next = (L + 1 < a.n_levels) ? a.level_ptr[L + 2] : e;   // asked for early: needed only at the end of the level
      if (COOP && e - b <= kSmall) {
        while (true) {
          if (blockIdx.x == 0) {
            for (long long idx0 = b + (long long)block_warp * RPW; idx0 < e; idx0 += (long long)block_warps * RPW) {
              const long long idx = idx0 + grp;
              const SweepPre<GS> p = sweep_preload_slot<GS>(a, idx, idx < e, sub);
              fin(p);
            }
            __syncthreads();
          }
          L++; b = e; e = e_next;
          if (L >= a.n_levels) break;
          e_next = (L + 1 < a.n_levels) ? a.level_ptr[L + 2] : e;
          if (e - b > kSmall) break;
        }
      } else {
        long long idx = b + gidx;
        if (COOP && have_pre) {
          fin(p0);
          if (b + gtotal < e) fin(p1);   // grid-uniform condition
          idx += 2 * gtotal;
        }
        // the rest of the level, software-pipelined: the graph side of the group's next row (one L2 round trip that does
        // not depend on anything this sweep writes) is in flight while the current row's posteriors are gathered
        long long idx0 = idx - grp;                                    // the warp's first slot (warp-uniform loop)
        if (idx0 < e) {
          SweepPre<GS> cur = sweep_preload_slot<GS>(a, idx0 + grp, idx0 + grp < e, sub);
          for (; idx0 < e; idx0 += gtotal) {
            const long long nx = idx0 + gtotal;
            SweepPre<GS> nxt = cur;
            if (nx < e) nxt = sweep_preload_slot<GS>(a, nx + grp, nx + grp < e, sub);
            fin(cur);
            cur = nxt;
          }
        }
        L++; b = e; e = e_next;
      }
      have_pre = false;
      if (COOP && L < a.n_levels) {
        if (e - b > kSmall) {             // graph side of the next level: independent of what this level wrote
          p0 = sweep_preload_slot<GS>(a, b + gidx, b + gidx < e, sub);
          if (b + gtotal < e) p1 = sweep_preload_slot<GS>(a, b + gidx + gtotal, b + gidx + gtotal < e, sub);
          have_pre = true;
        }
        stamp(2 * L - 1);
        bar.sync();
        stamp(2 * L);
      }
    }
    stamp(2 * a.n_levels - 1);
  }
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if (lane == 0 && wmax > 0.0f) atomicMax(a.maxdiff_bits, __float_as_uint(wmax));  // non-negative floats order as uints
  if (a.chg != nullptr) {           // rows this sweep changed: the next sweep's choice between levels and speculation
    for (int off = 16; off > 0; off >>= 1) n_changed += __shfl_xor_sync(0xffffffffu, n_changed, off);
    if (lane == 0 && n_changed > 0) atomicAdd(a.spec_words + 3 + a.spec_parity, n_changed);
  }
}

template <bool COOP, int GS>
__global__ void __launch_bounds__(COOP ? 1024 : 256) k_sweep(SweepArgs a) { sweep_body<COOP, GS>(a, GridBar()); }

// ------------------------------------------------------------------------------------------------------------------
// M-step.  Rows whose posterior is exactly one-hot in float32 (the vast majority after the first iteration) only add
// their bit pattern to one class; they are summed with bit-sliced carry-save counters (k_mstep_onehot), the remaining
// fuzzy rows take the fp64 path (k_mstep_fuzzy).
//
// The integer column counts are kept across EM iterations and updated incrementally: a row whose one-hot class did not
// change since the previous M-step contributes nothing new, so only rows that entered a class ("add" bucket k) or
// left one ("remove" bucket K + k) are streamed again.  Integer arithmetic makes this bit-identical to a full recount;
// the first M-step of a run (prev = kNone everywhere) is the full pass.  Fuzzy rows (bucket 2K) are always recomputed.
// Buckets are materialised as index lists: k_classify (bucket histogram) -> k_scatter (prefix sums + scatter).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kNone = -2;  // prev key of a row that is not counted anywhere yet

template <int KB>
__device__ __forceinline__ void
classify_body(const float* __restrict__ T, long long N, int K, int* __restrict__ prev /*[N] in/out: key of last M-step*/,
              int2* __restrict__ ent /*[N]: (remove bucket | -1, add/fuzzy bucket | -1)*/, int* __restrict__ hist /*[2K+1]*/,
              int* __restrict__ nk_cnt /*[K]*/, double* __restrict__ nk_fz /*[K]*/, int ignore_prev /*full recount*/) {
  __shared__ int s_hist[2 * KB + 1];
  __shared__ int s_nk[KB];
  __shared__ double s_fz[KB];
  for (int t = threadIdx.x; t < 2 * KB + 1; t += blockDim.x) s_hist[t] = 0;
  for (int t = threadIdx.x; t < KB; t += blockDim.x) { s_fz[t] = 0.0; s_nk[t] = 0; }
  __syncthreads();
  // one row per thread and pass; a grid smaller than the rows takes several passes (the tail of the sweep kernel)
  for (long long base = (long long)blockIdx.x * blockDim.x; base < N; base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + threadIdx.x;
    bool fuzzy = false;
    if (i < N) {
      int ones = 0, zeros = 0, arg = 0;
#pragma unroll
      for (int k = 0; k < KB; k++) {
        if (k < K) {
          const float t = T[(size_t)i * K + k];
          if (t == 1.0f) { ones++; arg = k; }
          else if (t == 0.0f) zeros++;
        }
      }
      int key;
      if (ones == 1 && zeros == K - 1) { key = arg; atomicAdd(&s_nk[key], 1); }
      else { key = K; fuzzy = true; }
      const int old = ignore_prev ? kNone : prev[i];
      int e0 = -1, e1 = -1;
      if (key == K) e1 = 2 * K;                   // fuzzy: always recomputed
      if (old != key) {
        if (old >= 0 && old < K) e0 = K + old;    // leaves class `old`
        if (key < K) e1 = key;                     // enters class `key`
        if (!ignore_prev) prev[i] = key;
      }
      ent[i] = make_int2(e0, e1);
      if (e0 >= 0) atomicAdd(&s_hist[e0], 1);
      if (e1 >= 0) atomicAdd(&s_hist[e1], 1);
    }
    // class sizes of the fuzzy rows: warp reduction first (in the K sweep nearly every row is fuzzy and per-thread
    // shared-memory fp64 atomics serialise), one shared atomic per warp and class
    if (__any_sync(0xffffffffu, fuzzy)) {
#pragma unroll
      for (int k = 0; k < KB; k++) {
        if (k < K) {
          double v = fuzzy ? (double)T[(size_t)i * K + k] : 0.0;
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&s_fz[k], v);
        }
      }
    }
  }
  __syncthreads();
  for (int t2 = threadIdx.x; t2 < 2 * K + 1; t2 += blockDim.x)
    if (s_hist[t2]) atomicAdd(&hist[t2], s_hist[t2]);
  for (int t2 = threadIdx.x; t2 < K; t2 += blockDim.x) {
    if (s_nk[t2]) atomicAdd(&nk_cnt[t2], s_nk[t2]);
    if (s_fz[t2] != 0.0) atomicAdd(&nk_fz[t2], s_fz[t2]);
  }
}

template <int KB>
__global__ void __launch_bounds__(256)
k_classify(const float* __restrict__ T, long long N, int K, int* __restrict__ prev, int2* __restrict__ ent, int* __restrict__ hist,
           int* __restrict__ nk_cnt, double* __restrict__ nk_fz, int ignore_prev) {
  classify_body<KB>(T, N, K, prev, ent, hist, nk_cnt, nk_fz, ignore_prev);
}

// rows -> perm, grouped by bucket.  One atomic per (warp, bucket present in the warp) instead of one per entry: the
// handful of cursors is shared by all rows, and per-row atomics on them serialise (measured 70 us for 200 000 rows).
__device__ __forceinline__ void scatter_one(int key, long long i, int lane, const int* s_off, int* __restrict__ fill,
                                            int* __restrict__ perm) {
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int leader = __ffs(peers) - 1;
  const int rank = __popc(peers & ((1u << lane) - 1u));
  int base = 0;
  if (key >= 0 && lane == leader) base = s_off[key] + atomicAdd(&fill[key], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (key >= 0) perm[base + rank] = (int)i;
}

// offsets[b] = first slot of bucket b in perm (prefix sums of the histogram, b = 0..nb; offsets[nb] = total): every CTA
// derives them itself (nb <= 65), CTA 0 also publishes them for the count kernels; `fill` (zeroed by the caller) counts
// the slots handed out per bucket.
__device__ __forceinline__ void
scatter_body(const int2* __restrict__ ent, long long N, const int* __restrict__ hist, int nb, int* __restrict__ offsets /*[nb+1]*/,
             int* __restrict__ fill /*[nb]*/, int* __restrict__ perm) {
  __shared__ int s_off[2 * NEMB200_MAX_K + 2];
  if (threadIdx.x < 32) {   // warp scan over <= 65 buckets, 3 per lane
    const int lane = threadIdx.x;
    int h[3], tot = 0;
#pragma unroll
    for (int q = 0; q < 3; q++) { const int b = lane * 3 + q; h[q] = b < nb ? hist[b] : 0; tot += h[q]; }
    int incl = tot;
    for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
    int acc = incl - tot;
#pragma unroll
    for (int q = 0; q < 3; q++) { const int b = lane * 3 + q; if (b <= nb) s_off[b] = acc; acc += h[q]; }
  }
  __syncthreads();
  if (blockIdx.x == 0) for (int b = threadIdx.x; b <= nb; b += blockDim.x) offsets[b] = s_off[b];
  const int lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < N; base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + threadIdx.x;
    int2 e = make_int2(-1, -1);
    if (i < N) e = ent[i];
    if (__any_sync(0xffffffffu, e.x >= 0)) scatter_one(e.x, i, lane, s_off, fill, perm);   // "leaves a class": rare, never on a full recount
    if (__any_sync(0xffffffffu, e.y >= 0)) scatter_one(e.y, i, lane, s_off, fill, perm);
  }
}

__global__ void __launch_bounds__(256)
k_scatter(const int2* __restrict__ ent, long long N, const int* __restrict__ hist, int nb, int* __restrict__ offsets,
          int* __restrict__ fill, int* __restrict__ perm) {
  scatter_body(ent, N, hist, nb, offsets, fill, perm);
}

constexpr int kPlanes = 12;
constexpr int kMsTpb = 128;    // threads per CTA; each owns one 64-bit matrix word (64 genomes) of every row
constexpr int kBatch = 16;     // rows per Harley-Seal block

__device__ __forceinline__ uint2 ld_stream64(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// 16 words -> the weight-16 carry, rippled into planes 4..11 (Harley-Seal)
__device__ __forceinline__ void hs16(uint32_t (&p)[kPlanes], const uint32_t (&x)[kBatch]) {
  uint32_t t2a, t2b, t2c, t2d, t2e, t2f, t2g, t2h, t4a, t4b, t4c, t4d, t8a, t8b, c16;
  csa(t2a, p[0], p[0], x[0], x[1]);
  csa(t2b, p[0], p[0], x[2], x[3]);
  csa(t2c, p[0], p[0], x[4], x[5]);
  csa(t2d, p[0], p[0], x[6], x[7]);
  csa(t2e, p[0], p[0], x[8], x[9]);
  csa(t2f, p[0], p[0], x[10], x[11]);
  csa(t2g, p[0], p[0], x[12], x[13]);
  csa(t2h, p[0], p[0], x[14], x[15]);
  csa(t4a, p[1], p[1], t2a, t2b);
  csa(t4b, p[1], p[1], t2c, t2d);
  csa(t4c, p[1], p[1], t2e, t2f);
  csa(t4d, p[1], p[1], t2g, t2h);
  csa(t8a, p[2], p[2], t4a, t4b);
  csa(t8b, p[2], p[2], t4c, t4d);
  csa(c16, p[3], p[3], t8a, t8b);
  uint32_t carry = c16;
#pragma unroll
  for (int j = 4; j < kPlanes; j++) {
    const uint32_t nc = p[j] & carry;
    p[j] ^= carry;
    carry = nc;
  }
}

__device__ __forceinline__ void load_batch(uint32_t (&lo)[kBatch], uint32_t (&hi)[kBatch], const uint2* __restrict__ colp,
                                           const int* s_rows, int r, int W2) {
#pragma unroll
  for (int q = 0; q < kBatch; q += 4) {
    const int4 rows = *reinterpret_cast<const int4*>(s_rows + r + q);
    const uint2 v0 = ld_stream64(colp + (size_t)rows.x * W2);
    const uint2 v1 = ld_stream64(colp + (size_t)rows.y * W2);
    const uint2 v2 = ld_stream64(colp + (size_t)rows.z * W2);
    const uint2 v3 = ld_stream64(colp + (size_t)rows.w * W2);
    lo[q] = v0.x; hi[q] = v0.y; lo[q + 1] = v1.x; hi[q + 1] = v1.y;
    lo[q + 2] = v2.x; hi[q + 2] = v2.y; lo[q + 3] = v3.x; hi[q + 3] = v3.y;
  }
}

// One thread owns one 64-bit word column (64 genomes) of a column tile; blockIdx.y owns slice y of the add/remove
// entries [offsets[0], offsets[2K]) -- equal shares, so every CTA streams the same number of rows and the grid is sized
// by the host to one resident wave (no tail).  A slice that straddles bucket boundaries flushes its counters once per
// bucket: +count for an "add" bucket, -count for a "remove" bucket.  Rows are streamed in batches of 16 with the next
// batch's loads issued before the current batch is summed; a bucket run longer than kMaxRun rows is flushed in pieces
// (12 counter planes hold up to 4095).
constexpr int kMaxRun = 4080;

__device__ __forceinline__ void
mstep_onehot_body(const uint2* __restrict__ bits2, int W2, int tile_w, int K, const int* __restrict__ offsets,
                  const int* __restrict__ perm, int* __restrict__ cnt /*[K][Dpad]*/, int Dpad) {
  __shared__ __align__(16) int s_rows[kMaxRun];
  __shared__ int s_cnt[kMsTpb * 33];
  const int total = offsets[2 * K];                       // add + remove entries
  // few entries (steady state of the incremental update): fewer, fuller slices -- the per-run counter flush is the
  // fixed cost of a CTA
  const long long S = min((long long)gridDim.y, max(1LL, ((long long)total + 63) / 64));
  if (blockIdx.y >= S) return;
  const int lo = (int)((long long)total * blockIdx.y / S), hi = (int)((long long)total * (blockIdx.y + 1) / S);
  if (lo >= hi) return;
  const int word = blockIdx.x * tile_w + threadIdx.x;
  const bool active = threadIdx.x < tile_w && word < W2;
  const uint2* colp = bits2 + (active ? word : 0);

  for (int b = 0; b < 2 * K; b++) {
    const int b0 = max(lo, offsets[b]), b1 = min(hi, offsets[b + 1]);
    for (int start = b0; start < b1; start += kMaxRun) {
      const int len = min(kMaxRun, b1 - start);
      __syncthreads();                                    // s_rows / s_cnt of the previous run are no longer read
      for (int t = threadIdx.x; t < len; t += blockDim.x) s_rows[t] = perm[start + t];
      __syncthreads();
      uint32_t plo[kPlanes], phi[kPlanes];
#pragma unroll
      for (int j = 0; j < kPlanes; j++) { plo[j] = 0; phi[j] = 0; }
      if (active) {
        const int nb = len / kBatch;
        uint32_t alo[kBatch], ahi[kBatch], blo[kBatch], bhi[kBatch];
        if (nb > 0) load_batch(alo, ahi, colp, s_rows, 0, W2);
        for (int q = 0; q < nb; q += 2) {
          if (q + 1 < nb) load_batch(blo, bhi, colp, s_rows, (q + 1) * kBatch, W2);
          hs16(plo, alo);
          hs16(phi, ahi);
          if (q + 1 < nb) {
            if (q + 2 < nb) load_batch(alo, ahi, colp, s_rows, (q + 2) * kBatch, W2);
            hs16(plo, blo);
            hs16(phi, bhi);
          }
        }
        for (int r = nb * kBatch; r < len; r++) {  // tail rows: ripple through all planes
          const uint2 v = ld_stream64(colp + (size_t)s_rows[r] * W2);
          uint32_t c0 = v.x, c1 = v.y;
#pragma unroll
          for (int j = 0; j < kPlanes; j++) {
            const uint32_t n0 = plo[j] & c0, n1 = phi[j] & c1;
            plo[j] ^= c0; phi[j] ^= c1;
            c0 = n0; c1 = n1;
          }
        }
      }
      // expand the planes into integer counts, transpose through shared memory, add to global with coalesced REDs
      const int cls = b < K ? b : b - K;
      const int sign = b < K ? 1 : -1;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        if (half) __syncthreads();
#pragma unroll 4
        for (int bit = 0; bit < 32; bit++) {
          int c = 0;
#pragma unroll
          for (int j = 0; j < kPlanes; j++) c += (int)(((half ? phi[j] : plo[j]) >> bit) & 1u) << j;
          s_cnt[threadIdx.x * 33 + bit] = c;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < tile_w * 32; t += blockDim.x) {
          const int c = s_cnt[(t >> 5) * 33 + (t & 31)];
          const int col = (blockIdx.x * tile_w + (t >> 5)) * 64 + half * 32 + (t & 31);
          if (c != 0 && col < Dpad) atomicAdd(&cnt[(size_t)cls * Dpad + col], sign * c);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kMsTpb)
k_mstep_onehot(const uint2* __restrict__ bits2, int W2, int tile_w, int K, const int* __restrict__ offsets,
               const int* __restrict__ perm, int* __restrict__ cnt, int Dpad) {
  mstep_onehot_body(bits2, W2, tile_w, K, offsets, perm, cnt, Dpad);
}

// ---- shared-memory ring variant for wide matrices -------------------------------------------------------------------
// ncu on the kernel above: long-scoreboard stalls 10 per issue, ALU 38 % -- the loads in flight (16 x 8 B per thread,
// held in registers) limit it, and a read-only pass needs >= 64 KB in flight per SM to reach the HBM ceiling
// (tools/probes/stream_probe.cu; the same probe shows 1-D bulk copies of 3 KB row segments capped at 6.0 TB/s by their
// per-copy cost, so the staging is done with per-thread cp.async instead).  One fat CTA per SM owns a column tile of
// up to kOhTpb 128-bit words; a thread copies its own word of 8 listed rows per group into private shared-memory
// slots, kOhGroups groups in flight (~100 KB per SM), no barrier in the streaming loop.  Counting (Harley-Seal over 8
// rows, 12 bit-sliced planes per 32 columns) and flushing are the same integer arithmetic, so the counts are identical.
constexpr int kOhTpb = 416;       // 13 warps: one 128-bit word column each
constexpr int kOhRows = 8;        // rows per group
constexpr int kOhGroups = 3;      // groups in the ring (kOhGroups - 1 in flight while one is summed)

// 8 words -> the weight-8 carry, rippled into planes 3..11
__device__ __forceinline__ void hs8(uint32_t (&p)[kPlanes], const uint32_t (&x)[kOhRows]) {
  uint32_t t2a, t2b, t2c, t2d, t4a, t4b, c8;
  csa(t2a, p[0], p[0], x[0], x[1]);
  csa(t2b, p[0], p[0], x[2], x[3]);
  csa(t2c, p[0], p[0], x[4], x[5]);
  csa(t2d, p[0], p[0], x[6], x[7]);
  csa(t4a, p[1], p[1], t2a, t2b);
  csa(t4b, p[1], p[1], t2c, t2d);
  csa(c8, p[2], p[2], t4a, t4b);
  uint32_t carry = c8;
#pragma unroll
  for (int j = 3; j < kPlanes; j++) {
    const uint32_t nc = p[j] & carry;
    p[j] ^= carry;
    carry = nc;
  }
}
__device__ __forceinline__ void ripple1(uint32_t (&p)[kPlanes], uint32_t c) {
#pragma unroll
  for (int j = 0; j < kPlanes; j++) {
    const uint32_t n = p[j] & c;
    p[j] ^= c;
    c = n;
  }
}

// 12 bit-sliced planes of one 32-column word -> 32 counts (<= 4095) as 16-bit values.  Four planes at a time are
// interleaved into nibbles (phase b holds columns 4i + b), bytes are assembled from the nibbles of two plane groups and
// PRMT picks the (low byte, high byte) pair of each column: ~6 instructions per column instead of 36 for a bit loop.
constexpr int kOhPitch16 = 130;   // 16-bit counts per thread row in shared memory: 128 + 2 (odd number of 32-bit words)
__device__ __forceinline__ void planes_to_counts16(const uint32_t (&p)[kPlanes], unsigned short* dst /*[32]*/) {
#pragma unroll
  for (int b = 0; b < 4; b++) {
    uint32_t a[3] = {0u, 0u, 0u};
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x = p[4 * g + j];
        const uint32_t sh = b >= j ? (x >> (b - j)) : (x << (j - b));
        a[g] |= sh & (0x11111111u << j);
      }
    // nibble i of a[g] = bits 4g..4g+3 of the count of column 4i + b
    const uint32_t lo01 = (a[0] & 0x0F0F0F0Fu) | ((a[1] << 4) & 0xF0F0F0F0u);   // byte k: bits 0..7 of column 8k + b
    const uint32_t hi01 = ((a[0] >> 4) & 0x0F0F0F0Fu) | (a[1] & 0xF0F0F0F0u);   // byte k: bits 0..7 of column 8k + 4 + b
    const uint32_t lo2 = a[2] & 0x0F0F0F0Fu, hi2 = (a[2] >> 4) & 0x0F0F0F0Fu;    // byte k: bits 8..11
#pragma unroll
    for (int k = 0; k < 4; k++) {
      dst[8 * k + b] = (unsigned short)__byte_perm(lo01, lo2, k | ((4 + k) << 4));
      dst[8 * k + 4 + b] = (unsigned short)__byte_perm(hi01, hi2, k | ((4 + k) << 4));
    }
  }
}

__host__ __device__ inline size_t onehot_ring_smem() {
  return (size_t)kOhGroups * kOhRows * kOhTpb * 16 + (size_t)kMaxRun * 4;
}

__device__ __forceinline__ void
mstep_onehot_ring_body(const uint4* __restrict__ bits4, int W4, int tile_w /*<= kOhTpb*/, int K, const int* __restrict__ offsets,
                       const int* __restrict__ perm, int* __restrict__ cnt /*[K][Dpad]*/, int Dpad, int slice_rows = 256) {
  extern __shared__ uint4 smem4[];
  uint4* ring = smem4;                                            // [kOhGroups][kOhRows][kOhTpb]
  int* s_rows = reinterpret_cast<int*>(ring + kOhGroups * kOhRows * kOhTpb);
  unsigned short* s_cnt16 = reinterpret_cast<unsigned short*>(ring);   // the flush reuses the (drained) ring
  constexpr uint32_t kGroupBytes = kOhRows * kOhTpb * 16;

  const int total = offsets[2 * K];
  // A CTA's counter flush (one RED per non-zero column of its tile and bucket, ~53 000 of them) costs as much as streaming
  // ~200 rows; measured on the benchmark instance, 32 to 256 rows per slice give 41-44 us per incremental launch (the
  // listed rows are scattered over the matrix, one 6 KB row per page), 64 being the best.
  // slices of at most `slice_rows` rows, but no fewer slices than CTAs while a slice keeps 64 rows (a shard's full count of
  // 25 000 rows: 148 slices of 169 rows instead of 98 of 256; a few thousand changed rows: 64-row slices)
  const int eff_rows = min(slice_rows, max(64, (total + (int)gridDim.y - 1) / (int)gridDim.y));
  const long long S = min((long long)gridDim.y, max(1LL, ((long long)total + eff_rows - 1) / eff_rows));
  if (blockIdx.y >= S) return;
  const int lo = (int)((long long)total * blockIdx.y / S), hi = (int)((long long)total * (blockIdx.y + 1) / S);
  if (lo >= hi) return;
  const int word0 = blockIdx.x * tile_w;
  const int nw = min(tile_w, W4 - word0);                         // 128-bit words of this tile
  const bool active = (int)threadIdx.x < nw;
  const uint4* colp = bits4 + word0 + (active ? threadIdx.x : 0);
  const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(ring) + threadIdx.x * 16;
  const char* myslot = reinterpret_cast<const char*>(ring + threadIdx.x);

  for (int b = 0; b < 2 * K; b++) {
    const int b0 = max(lo, offsets[b]), b1 = min(hi, offsets[b + 1]);
    for (int start = b0; start < b1; start += kMaxRun) {
      const int len = min(kMaxRun, b1 - start);
      __syncthreads();                                    // the previous run's flush no longer reads s_cnt / s_rows
      for (int t = threadIdx.x; t < len; t += kOhTpb) s_rows[t] = perm[start + t];
      __syncthreads();
      const int ngroups = (len + kOhRows - 1) / kOhRows;  // the last one may hold fewer rows
      uint32_t pl[4][kPlanes];
#pragma unroll
      for (int w = 0; w < 4; w++)
#pragma unroll
        for (int j = 0; j < kPlanes; j++) pl[w][j] = 0;

      int gi = 0, sj = 0;                                 // next group to issue, and the ring slot it goes to
      auto issue = [&]() {
        if (gi < ngroups && active) {
          const int r0 = gi * kOhRows;
          const uint32_t dst = slot0 + sj * kGroupBytes;
          if (r0 + kOhRows <= len) {
#pragma unroll
            for (int j = 0; j < kOhRows; j += 4) {
              const int4 rows = *reinterpret_cast<const int4*>(s_rows + r0 + j);
              cp_async16(dst + (j + 0) * (kOhTpb * 16), colp + (size_t)rows.x * W4);
              cp_async16(dst + (j + 1) * (kOhTpb * 16), colp + (size_t)rows.y * W4);
              cp_async16(dst + (j + 2) * (kOhTpb * 16), colp + (size_t)rows.z * W4);
              cp_async16(dst + (j + 3) * (kOhTpb * 16), colp + (size_t)rows.w * W4);
            }
          } else {
            for (int j = 0; r0 + j < len; j++) cp_async16(dst + j * (kOhTpb * 16), colp + (size_t)s_rows[r0 + j] * W4);
          }
        }
        cp_async_commit();
        gi++;
        sj = (sj + 1 == kOhGroups) ? 0 : sj + 1;
      };
#pragma unroll
      for (int s = 0; s < kOhGroups - 1; s++) issue();
      for (int g = 0; g < ngroups; g++) {
        issue();                                          // afterwards sj is the slot of group g
        cp_async_wait<kOhGroups - 1>();
        if (active) {
          const char* src = myslot + sj * kGroupBytes;
          const int rows_in = min(kOhRows, len - g * kOhRows);
          if (rows_in == kOhRows) {
            uint32_t x0[kOhRows], x1[kOhRows], x2[kOhRows], x3[kOhRows];
#pragma unroll
            for (int j = 0; j < kOhRows; j++) {
              const uint4 v = *reinterpret_cast<const uint4*>(src + j * (kOhTpb * 16));
              x0[j] = v.x; x1[j] = v.y; x2[j] = v.z; x3[j] = v.w;
            }
            hs8(pl[0], x0);
            hs8(pl[1], x1);
            hs8(pl[2], x2);
            hs8(pl[3], x3);
          } else {
            for (int j = 0; j < rows_in; j++) {           // tail rows of the run: ripple through all planes
              const uint4 v = *reinterpret_cast<const uint4*>(src + j * (kOhTpb * 16));
              ripple1(pl[0], v.x); ripple1(pl[1], v.y); ripple1(pl[2], v.z); ripple1(pl[3], v.w);
            }
          }
        }
      }
      cp_async_wait<0>();
      __syncthreads();                                    // every thread is done with the ring: reuse it as s_cnt
      // expand the planes into 16-bit counts in shared memory (thread-major, odd word pitch), then walk the tile's
      // columns in order: coalesced REDs to the global counters
      const int cls = b < K ? b : b - K;
      const int sign = b < K ? 1 : -1;
      unsigned short* mine = s_cnt16 + threadIdx.x * kOhPitch16;
#pragma unroll
      for (int w = 0; w < 4; w++) planes_to_counts16(pl[w], mine + w * 32);
      __syncthreads();
      for (int t = threadIdx.x; t < nw * 128; t += kOhTpb) {
        const int c = s_cnt16[(t >> 7) * kOhPitch16 + (t & 127)];
        const int col = word0 * 128 + t;
        if (c != 0 && col < Dpad) atomicAdd(&cnt[(size_t)cls * Dpad + col], sign * c);
      }
    }
  }
}

__global__ void __launch_bounds__(kOhTpb, 1)
k_mstep_onehot_ring(const uint4* __restrict__ bits4, int W4, int tile_w, int K, const int* __restrict__ offsets,
                    const int* __restrict__ perm, int* __restrict__ cnt, int Dpad) {
  mstep_onehot_ring_body(bits4, W4, tile_w, K, offsets, perm, cnt, Dpad);
}

// Fuzzy rows: fsum[k][d] += T[i][k] for every set bit (i, d) of a fuzzy row.  A thread owns bit `lane` of WPT consecutive
// matrix words (columns 32 apart) with fp64 accumulators per class; a warp's 32 lanes look at the same words (one
// 4/8/16-byte broadcast load per row), so all-zero groups are skipped with a uniform branch and a row's posteriors
// are fetched once for WPT words.  blockIdx.y owns an equal share of the fuzzy list (the host sizes the grid to the
// resident CTA slots: in the K-selection sweep at K = 20 nearly every row is fuzzy and this kernel is the M-step).  The
// share's row ids and posteriors (converted to fp64 once) are staged in shared memory, so the matrix loads of
// consecutive rows are independent.  The kernel is instruction-bound (8 warps x n_fuzzy x Ws / WPT row visits).
constexpr int kFzRows = 64;
template <int KB> struct FuzzyCfg { static constexpr int WPT = KB <= 4 ? 4 : (KB <= 8 ? 2 : 1); };

template <int KB>
__device__ __forceinline__ void
mstep_fuzzy_body(const uint32_t* __restrict__ bits, int Ws, int K, int D, const int* __restrict__ offsets,
                 const int* __restrict__ perm, const float* __restrict__ T, double* __restrict__ fsum /*[K][Dpad]*/,
                 int Dpad) {
  constexpr int KP = (KB + 1) & ~1;   // even: the staged posteriors are read as double2
  constexpr int WPT = FuzzyCfg<KB>::WPT;
  constexpr int XW = 8 * WPT / 4;     // uint4 per row that the CTA's 8 warps look at (contiguous in the row)
  __shared__ __align__(16) double s_t[kFzRows][KP];
  __shared__ uint4 s_x[kFzRows][XW];
  const int f0 = offsets[2 * K], f1 = offsets[2 * K + 1];   // the fuzzy bucket
  const int nf = f1 - f0;
  const long long S = min((long long)gridDim.y, max(1LL, ((long long)nf + kFzRows - 1) / kFzRows));
  if (blockIdx.y >= S) return;
  const int lo = f0 + (int)((long long)nf * blockIdx.y / S), hi = f0 + (int)((long long)nf * (blockIdx.y + 1) / S);
  if (lo >= hi) return;
  const int bit = threadIdx.x & 31;
  const int word0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * WPT;   // Ws is a multiple of 4 >= WPT
  const bool in_range = word0 < Ws;
  const int cta_word0 = blockIdx.x * 8 * WPT;
  double acc[WPT][KB];
#pragma unroll
  for (int w = 0; w < WPT; w++)
#pragma unroll
    for (int k = 0; k < KB; k++) acc[w][k] = 0.0;
  for (int base = lo; base < hi; base += kFzRows) {
    const int n = min(kFzRows, hi - base);
    __syncthreads();
    // the batch's posteriors and matrix words in one round of independent loads (row id first, then both)
    for (int t = threadIdx.x; t < n * (KP + XW); t += blockDim.x) {
      if (t < n * KP) {
        const int r = t / KP, k = t - r * KP;
        const int row = perm[base + r];
        s_t[r][k] = k < K ? (double)T[(size_t)row * K + k] : 0.0;
      } else {
        const int u = t - n * KP;
        const int r = u / XW, q = u - r * XW;
        const int row = perm[base + r];
        const int wq = cta_word0 + 4 * q;
        s_x[r][q] = wq < Ws ? __ldg(reinterpret_cast<const uint4*>(bits + (size_t)row * Ws + wq)) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncthreads();
    const uint32_t* myx = reinterpret_cast<const uint32_t*>(&s_x[0][0]) + (threadIdx.x >> 5) * WPT;
#pragma unroll 4
    for (int r = 0; r < n; r++) {
      uint32_t x[WPT];
      const uint32_t* src = myx + r * (XW * 4);
      if (WPT == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        x[0] = v.x; x[1 % WPT] = v.y; x[2 % WPT] = v.z; x[3 % WPT] = v.w;
      } else if (WPT == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(src);
        x[0] = v.x; x[1 % WPT] = v.y;
      } else {
        x[0] = *src;
      }
      uint32_t any = 0u;
#pragma unroll
      for (int w = 0; w < WPT; w++) any |= x[w];
      if (WPT == 1) {
        if (any != 0u && ((x[0] >> bit) & 1u)) {
          const double2* tr = reinterpret_cast<const double2*>(&s_t[r][0]);
#pragma unroll
          for (int k2 = 0; k2 < KP / 2; k2++) {
            if (2 * k2 < K) {
              const double2 v = tr[k2];
              acc[0][2 * k2] += v.x;
              if (2 * k2 + 1 < KB) acc[0][2 * k2 + 1] += v.y;
            }
          }
        }
      } else if (any != 0u) {              // warp-uniform: the 32 lanes of a warp share the words
        double t[KP];
        const double2* tr = reinterpret_cast<const double2*>(&s_t[r][0]);
#pragma unroll
        for (int k2 = 0; k2 < KP / 2; k2++) {
          t[2 * k2] = 0.0; t[2 * k2 + 1] = 0.0;
          if (2 * k2 < K) { const double2 v = tr[k2]; t[2 * k2] = v.x; t[2 * k2 + 1] = v.y; }
        }
#pragma unroll
        for (int w = 0; w < WPT; w++) {
          if ((x[w] >> bit) & 1u) {
#pragma unroll
            for (int k = 0; k < KB; k++)
              if (k < K) acc[w][k] += t[k];
          }
        }
      }
    }
  }
#pragma unroll
  for (int w = 0; w < WPT; w++) {
    const int d = (word0 + w) * 32 + bit;
    if (in_range && d < D) {
#pragma unroll
      for (int k = 0; k < KB; k++)
        if (k < K && acc[w][k] != 0.0) atomicAdd(&fsum[(size_t)k * Dpad + d], acc[w][k]);
    }
  }
}

template <int KB>
__global__ void __launch_bounds__(256)
k_mstep_fuzzy(const uint32_t* __restrict__ bits, int Ws, int K, int D, const int* __restrict__ offsets,
              const int* __restrict__ perm, const float* __restrict__ T, double* __restrict__ fsum, int Dpad) {
  mstep_fuzzy_body<KB>(bits, Ws, K, D, offsets, perm, T, fsum, Dpad);
}

// ---- fuzzy-row sums on the FP64 tensor cores ---------------------------------------------------------------------------
// fsum[k][d] = sum_i T[i][k] x[i][d] over the fuzzy rows is a (K x n_fuzzy) . (n_fuzzy x D) product; in the K-selection
// sweep nearly every row is fuzzy and K goes up to 20, where the predicated-add kernel above issues K fp64 adds per
// (row, column) whether the bit is set or not -- 4x the fp64 pipe's bound (measured: 392 us per iteration for the four
// K = 17..20 candidates of BASELINE configs[2], half of the whole sweep).  DMMA (mma.sync m8n8k4, f64: the only FP64
// tensor-core instruction -- tcgen05 has no f64 kind) keeps the exact fp64 accumulation: A = 8 classes x 4 rows of
// posteriors, B = 4 rows x 8 columns of matrix bits expanded to 0.0 / 1.0, C = 8 x 8 fp64.  A warp owns one 32-column
// word (4 column tiles) and all class tiles; rows are staged in shared memory 64 at a time like in the kernel above.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int KT /*class tiles of 8*/>
__device__ __forceinline__ void
mstep_fuzzy_mma_body(const uint32_t* __restrict__ bits, int Ws, int K, int D, const int* __restrict__ offsets,
                     const int* __restrict__ perm, const float* __restrict__ T, double* __restrict__ fsum /*[K][Dpad]*/, int Dpad) {
  constexpr int KP = KT * 8 + ((KT & 1) ? 12 : 4);   // row pitch in doubles, = 4 mod 16: the 16 lanes of a half-warp hit 16 banks
  constexpr int XW = 8;                              // 32-bit words per row that the CTA's 8 warps look at
  __shared__ __align__(16) double s_t[kFzRows][KP];
  __shared__ __align__(16) uint32_t s_x[kFzRows][XW];
  const int f0 = offsets[2 * K], f1 = offsets[2 * K + 1];   // the fuzzy bucket
  const int nf = f1 - f0;
  const long long S = min((long long)gridDim.y, max(1LL, ((long long)nf + kFzRows - 1) / kFzRows));
  if (blockIdx.y >= S) return;
  const int lo = f0 + (int)((long long)nf * blockIdx.y / S), hi = f0 + (int)((long long)nf * (blockIdx.y + 1) / S);
  if (lo >= hi) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;            // fragment coordinates: g = class (A) / column (B) index, t4 = row of the quad
  const int cta_word0 = blockIdx.x * XW;
  const int word = cta_word0 + wid;
  const int ktiles = min(KT, (K + 7) >> 3);
  double acc[KT][4][2];
#pragma unroll
  for (int mt = 0; mt < KT; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
  for (int base = lo; base < hi; base += kFzRows) {
    const int n = min(kFzRows, hi - base);
    const int n4 = (n + 3) & ~3;
    __syncthreads();
    for (int t = threadIdx.x; t < n4 * (KT * 8 + XW); t += blockDim.x) {
      if (t < n4 * KT * 8) {
        const int r = t / (KT * 8), k = t - r * (KT * 8);
        double v = 0.0;
        if (r < n && k < K) v = (double)T[(size_t)perm[base + r] * K + k];
        s_t[r][k] = v;                                // rows past n are zero: they add nothing
      } else {
        const int u = t - n4 * KT * 8;
        const int r = u / XW, q = u - r * XW;
        const int wq = cta_word0 + q;
        s_x[r][q] = (r < n && wq < Ws) ? __ldg(bits + (size_t)perm[base + r] * Ws + wq) : 0u;
      }
    }
    __syncthreads();
    if (word < Ws) {
#pragma unroll 2
      for (int r0 = 0; r0 < n4; r0 += 4) {
        const uint32_t x = s_x[r0 + t4][wid];
        if (__any_sync(0xffffffffu, x != 0u)) {        // an all-zero quad of this word adds nothing
          double b[4];
#pragma unroll
          for (int nt = 0; nt < 4; nt++) b[nt] = ((x >> (nt * 8 + g)) & 1u) ? 1.0 : 0.0;
#pragma unroll
          for (int mt = 0; mt < KT; mt++) {
            if (mt < ktiles) {
              const double a = s_t[r0 + t4][mt * 8 + g];
#pragma unroll
              for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a, b[nt]);
            }
          }
        }
      }
    }
  }
  if (word < Ws) {
#pragma unroll
    for (int mt = 0; mt < KT; mt++) {
      const int k = mt * 8 + g;                         // C fragment: row = class g, columns 2 t4, 2 t4 + 1 of the tile
      if (mt < ktiles && k < K) {
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int d = word * 32 + nt * 8 + t4 * 2 + c;
            const double v = acc[mt][nt][c];
            if (d < D && v != 0.0) atomicAdd(&fsum[(size_t)k * Dpad + d], v);
          }
        }
      }
    }
  }
}

template <int KT>
__global__ void __launch_bounds__(256)
k_mstep_fuzzy_mma(const uint32_t* __restrict__ bits, int Ws, int K, int D, const int* __restrict__ offsets,
                  const int* __restrict__ perm, const float* __restrict__ T, double* __restrict__ fsum, int Dpad) {
  mstep_fuzzy_mma_body<KT>(bits, Ws, K, D, offsets, perm, T, fsum, Dpad);
}

// Parameters from the column statistics.  grid = (column chunks, K); every CTA handles kFinCols columns of one class:
// centres (weighted median of a 0/1 column, nem_mod.c:1417-1474), inertia (nem_mod.c:1641-1699), centre bit planes and,
// for "skd", the per-column dispersion and log-odds.  The last CTA of a class to finish (ticket counter) adds the chunk
// partials in chunk order (deterministic) and derives the class scalars: proportion (nem_mod.c:455-459), the shared
// dispersion of "sk_" (nem_mod.c:1053-1068, MISSING_IGNORE) and the E-step constants L_k, B_k.
constexpr int kFinCols = 256;     // one column per thread: the kernel is a chain of dependent memory round trips, many small CTAs hide them

__device__ __forceinline__ void finalize_class_scalars(const Params& P, int k, double nk, int nchunk, const double* __restrict__ partial,
                                                       long long N_total, int lane);

// size of class k over all ranks (rank order: the same value on every rank); every peer load in flight at once
__device__ __forceinline__ double class_size_all_ranks(const PeerStats& ps, size_t KD, int k) {
  int ni[kMaxPeers];
  double nf[kMaxPeers];
#pragma unroll
  for (int q = 0; q < kMaxPeers; q++) {
    ni[q] = 0; nf[q] = 0.0;
    if (q < ps.n) { ni[q] = ps.iblock[q][KD + k]; nf[q] = ps.dblock[q][KD + k]; }
  }
  long long nk_i = 0;
  double nk_f = 0.0;
#pragma unroll
  for (int q = 0; q < kMaxPeers; q++) { nk_i += ni[q]; nk_f += nf[q]; }
  return (double)nk_i + nk_f;
}

// With `pp` and `fs` (row-sharded run with NEMB200_LL=0: the flag-barrier variant kept for comparison with k_b_finalize_ll)
// the column chunks are dealt out to the ranks: after the first flag barrier a rank reduces only its chunks (1 / n of the
// peer loads) and writes their centres, planes, dispersions and partial sums into EVERY rank's parameter block (peer
// stores); after the second barrier the last CTA derives the class scalars.
__device__ __forceinline__ void
mstep_finalize_body(const Params& P, const PeerStats& ps, long long N_total, double* __restrict__ partial /*[K][nchunk][2]*/,
                    int* __restrict__ tickets /*[K]*/, const PeerParams* pp = nullptr, const FinSync* fs = nullptr) {
  __shared__ double s_red[8][2];
  __shared__ int s_last;
  const int k = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int nw = pp != nullptr ? pp->n : 1;                 // ranks that receive this CTA's results
  const int D = P.D, Dpad = P.Dpad, Ws = P.Ws;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t KD = (size_t)P.K * Dpad;
  unsigned long long* dbg = (fs != nullptr && (int)blockIdx.x == fs->rank && blockIdx.y == 0 && threadIdx.x == 0) ? fs->dbg : nullptr;
  dbg_stamp(dbg, 0);
  if (fs != nullptr) {
    // the all-reduce's flag barrier (phase 0): "my statistics of this iteration are complete" -- they are, the count kernels
    // precede this launch -- by the first CTA; every CTA waits for all ranks' before it loads their statistics
    if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < fs->world) {
      __threadfence_system();
      volatile int* dst = fs->sig_ptrs[threadIdx.x] + 0 * kMaxPeers + fs->rank;
      *dst = fs->epoch;
      __threadfence_system();
    }
    peer_wait_lane(fs->sig, fs->world, 0, fs->epoch, fs->err);
    __syncthreads();
  }
  dbg_stamp(dbg, 1);
  const double nk = class_size_all_ranks(ps, KD, k);
  const bool empty = !(nk > kEps);   // empty class: nem_mod.c:1399-1403 -> STS_W_EMPTYCLASS, parameters kept
  if (empty && threadIdx.x == 0 && chunk == 0) atomicMax(P.status, 1);
  const bool mine = !empty && !(pp != nullptr && chunk % pp->n != pp->rank);   // (else: another rank's chunk)
  if (fs == nullptr && !mine) return;
  if (mine) {
  double iner_sum = 0.0, b_sum = 0.0;
  int any_half = 0;
  // every statistic of the CTA's columns is requested before the first result is stored (peer loads: one round trip)
  int ci[kFinCols / 256][kMaxPeers];
  double cf[kFinCols / 256][kMaxPeers];
#pragma unroll
  for (int j = 0; j < kFinCols / 256; j++) {
    const int d = chunk * kFinCols + j * 256 + threadIdx.x;
#pragma unroll
    for (int q = 0; q < kMaxPeers; q++) {
      ci[j][q] = 0; cf[j][q] = 0.0;
      if (d < D && q < ps.n) { ci[j][q] = ps.iblock[q][(size_t)k * Dpad + d]; cf[j][q] = ps.dblock[q][(size_t)k * Dpad + d]; }
    }
  }
#pragma unroll
  for (int j = 0; j < kFinCols / 256; j++) {
    const int d = chunk * kFinCols + j * 256 + threadIdx.x;
    float m = 0.0f;
    double iner = 0.0;
    if (d < D) {
      long long c_i = 0;
      double c_f = 0.0;
#pragma unroll
      for (int q = 0; q < kMaxPeers; q++) { c_i += ci[j][q]; c_f += cf[j][q]; }   // added in rank order
      const double s1 = (double)c_i + c_f;
      if (s1 >= 0.0) dbg_stamp(dbg, 2);
      const double s0 = nk - s1;
      // Weighted median of a 0/1 column (nem_mod.c:1417-1474): 0 if the weight of the zero rows exceeds half the class
      // size, 1 if it falls short, midway (0.5) on a tie.  The reference compares float32 sums, in which weights below
      // the float32 resolution of the class size are absorbed -- near-ties become exact ties there -- so the comparison
      // is made at that resolution.
      const float w0 = (float)s0, half = (float)nk * 0.5f;
      if (w0 > half) { m = 0.0f; iner = s1; }
      else if (w0 < half) { m = 1.0f; iner = s0; }
      else { m = 0.5f; iner = 0.5 * nk; any_half = 1; }   // tie: midway between a 0 and a 1 (nem_mod.c:1463-1471)
      if (pp == nullptr) P.mu[(size_t)k * D + d] = m;
      else for (int q = 0; q < nw; q++) pp->mu[q][(size_t)k * D + d] = m;
      if (P.free_disp) {
        const float e = (float)(iner / nk);                // nem_mod.c:1162-1163
        double Lv;
        if (e > kEps) { Lv = log((double)((1.0f - e) / e)); b_sum += log((double)(1.0f - e)); }  // nem_mod.c:660
        else Lv = INFINITY;
        if (pp == nullptr) { P.eps[(size_t)k * D + d] = e; P.L[(size_t)k * Dpad + d] = Lv; }
        else for (int q = 0; q < nw; q++) { pp->eps[q][(size_t)k * D + d] = e; pp->L[q][(size_t)k * Dpad + d] = Lv; }
      }
    } else if (P.free_disp && d < Dpad) {
      if (pp == nullptr) P.L[(size_t)k * Dpad + d] = 0.0;
      else for (int q = 0; q < nw; q++) pp->L[q][(size_t)k * Dpad + d] = 0.0;
    }
    iner_sum += iner;
    const unsigned b1 = __ballot_sync(0xffffffffu, m == 1.0f);
    const unsigned bv = __ballot_sync(0xffffffffu, m != 0.5f);
    if (lane == 0 && (d >> 5) < Ws) {
      if (pp == nullptr) { P.m1[(size_t)k * Ws + (d >> 5)] = b1; P.valid[(size_t)k * Ws + (d >> 5)] = bv; }
      else for (int q = 0; q < nw; q++) { pp->m1[q][(size_t)k * Ws + (d >> 5)] = b1; pp->valid[q][(size_t)k * Ws + (d >> 5)] = bv; }
    }
  }
  if (__any_sync(0xffffffffu, any_half) && lane == 0) {
    if (pp == nullptr) atomicOr(P.has_half, 1);
    else for (int q = 0; q < nw; q++) atomicOr(pp->has_half[q], 1);
  }
  for (int off = 16; off > 0; off >>= 1) {
    iner_sum += __shfl_xor_sync(0xffffffffu, iner_sum, off);
    b_sum += __shfl_xor_sync(0xffffffffu, b_sum, off);
  }
  if (lane == 0) { s_red[wid][0] = iner_sum; s_red[wid][1] = b_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int q = 0; q < 8; q++) { t0 += s_red[q][0]; t1 += s_red[q][1]; }
    if (pp != nullptr) {                                   // every rank receives the partials; the scalars follow after the barrier
      for (int q = 0; q < nw; q++) {
        pp->fin_partial[q][((size_t)k * nchunk + chunk) * 2 + 0] = t0;
        pp->fin_partial[q][((size_t)k * nchunk + chunk) * 2 + 1] = t1;
      }
      __threadfence_system();
      s_last = 0;
    } else {
      partial[((size_t)k * nchunk + chunk) * 2 + 0] = t0;
      partial[((size_t)k * nchunk + chunk) * 2 + 1] = t1;
      __threadfence();
      s_last = (atomicAdd(&tickets[k], 1) == nchunk - 1);
    }
  }
  __syncthreads();
  if (pp == nullptr) {
    if (!s_last || wid != 0) return;
    __threadfence();
    finalize_class_scalars(P, k, nk, nchunk, partial, N_total, lane);
    if (lane == 0) tickets[k] = 0;  // ready for the next M-step
    return;
  }
  dbg_stamp(dbg, 3);
  __threadfence_system();
  dbg_stamp(dbg, 4);
  }  // mine
  if (fs == nullptr) return;
  // Row-sharded: the last CTA of the launch to get here (every chunk of this rank is written on every rank) publishes
  // "my chunks of this iteration are everywhere" (phase 3), waits for every rank's, and derives the class scalars from
  // the partial sums of all chunks -- identical on every rank: same values, same order.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int total = (int)gridDim.x * P.K;
    s_last = (atomicAdd(fs->ticket, 1) == total - 1);
    if (s_last) *fs->ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  unsigned long long* dbg2 = threadIdx.x == 0 ? fs->dbg : nullptr;
  dbg_stamp(dbg2, 5);
  if ((int)threadIdx.x < fs->world) {
    __threadfence_system();
    volatile int* dst = fs->sig_ptrs[threadIdx.x] + 3 * kMaxPeers + fs->rank;
    *dst = fs->epoch;
    __threadfence_system();
  }
  dbg_stamp(dbg2, 6);
  peer_wait_lane(fs->sig, fs->world, 3, fs->epoch, fs->err);
  __syncthreads();
  dbg_stamp(dbg2, 7);
  for (int kk = wid; kk < P.K; kk += (int)(blockDim.x >> 5)) {
    const double nkk = class_size_all_ranks(ps, KD, kk);
    if (kk == 0 && nkk >= 0.0 && lane == 0) dbg_stamp(fs->dbg, 8);
    if (nkk > kEps) finalize_class_scalars(P, kk, nkk, nchunk, partial, N_total, lane);
  }
  __syncthreads();
  dbg_stamp(dbg2, 9);
}

// class scalars from the chunk partials (added in chunk order: deterministic): proportion (nem_mod.c:455-459), the shared
// dispersion of "sk_" (nem_mod.c:1053-1068, MISSING_IGNORE) and the E-step constants L_k, B_k.  One warp.
__device__ __forceinline__ void finalize_class_scalars(const Params& P, int k, double nk, int nchunk, const double* __restrict__ partial,
                                                       long long N_total, int lane) {
  const int D = P.D;
  // chunk partials added in chunk order (deterministic); the loads of 32 chunks are in flight together
  double tot_iner = 0.0, tot_b = 0.0;
  for (int c0 = 0; c0 < nchunk; c0 += 32) {
    const int c = c0 + lane;
    double v0 = 0.0, v1 = 0.0;
    if (c < nchunk) {
      v0 = __ldcg(&partial[((size_t)k * nchunk + c) * 2 + 0]);
      v1 = __ldcg(&partial[((size_t)k * nchunk + c) * 2 + 1]);
    }
    const int m = min(32, nchunk - c0);
    for (int j = 0; j < m; j++) {
      tot_iner += __shfl_sync(0xffffffffu, v0, j);
      tot_b += __shfl_sync(0xffffffffu, v1, j);
    }
  }
  if (lane != 0) return;
  const float pkf = (float)(nk / (double)N_total);       // nem_mod.c:455-459
  P.pk[k] = pkf;
  const double lp = ((double)pkf > kEps) ? log((double)pkf) : -INFINITY;  // nem_alg.c:2350-2356
  P.logpk32[k] = (float)lp;
  P.logpk64[k] = lp;
  if (!P.free_disp) {
    const float e = (float)(tot_iner / ((double)D * nk));
    P.eps_k[k] = e;                                       // expanded to eps[k][:] on demand (k_expand_eps)
    if (e > kEps) {
      P.L[k] = log((double)((1.0f - e) / e));            // float ratio, double log (nem_mod.c:660)
      P.B[k] = (double)D * log((double)(1.0f - e));
    } else {
      P.L[k] = INFINITY;
      P.B[k] = 0.0;
    }
  } else {
    P.B[k] = tot_b;
  }
}

__global__ void __launch_bounds__(256)
k_mstep_finalize(Params P, PeerStats ps, long long N_total, double* __restrict__ partial, int* __restrict__ tickets) {
  mstep_finalize_body(P, ps, N_total, partial, tickets);
}

// "sk_": the dispersion is one scalar per class; the K x D table PPanGGOLiN reads from .mf is materialised on demand
__global__ void k_expand_eps(Params P) {
  const int k = blockIdx.y;
  const float e = P.eps_k[k];
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < P.D; d += gridDim.x * blockDim.x) P.eps[(size_t)k * P.D + d] = e;
}

// nibble tables of k_density_skd: Ltab[k][q][v] = sum of L_kd over the set bits b of v, d = 4 q + b  (q over Dpad / 4),
// and the totals of every 32-column word.  One thread per (class, word).
__device__ __forceinline__ void ltab_body(const double* __restrict__ Lkd, int K, int Ws, double* __restrict__ Ltab,
                                          double* __restrict__ Lword) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (k, word): [K][Dpad] is contiguous
  if (t >= (long long)K * Ws) return;
  double total = 0.0;
  for (int q = 0; q < 8; q++) {
    const double* l = Lkd + t * 32 + q * 4;
    const double l0 = l[0], l1 = l[1], l2 = l[2], l3 = l[3];
    double* out = Ltab + (t * 8 + q) * 16;
#pragma unroll
    for (int v = 0; v < 16; v++) {
      double a = 0.0;
      if (v & 1) a += l0;
      if (v & 2) a += l1;
      if (v & 4) a += l2;
      if (v & 8) a += l3;
      out[v] = a;
    }
    total += ((l0 + l1) + l2) + l3;
  }
  Lword[t] = total;
}

__global__ void k_ltab(const double* __restrict__ Lkd, int K, int Ws, double* __restrict__ Ltab, double* __restrict__ Lword) {
  ltab_body(Lkd, K, Ws, Ltab, Lword);
}

// initial parameters: partition.py:65-85 (values) + nem_exe.c:1023-1040 (last proportion by subtraction)
__device__ __forceinline__ void init_params_body(const Params& P, float base_prop, const float* __restrict__ eps_k /*[K] host-computed*/) {
  const int K = P.K, D = P.D, Ws = P.Ws, Dpad = P.Dpad;
  const int k = blockIdx.x;
  const bool one = (double)(k + 1) <= (double)K / 2.0;
  const float e = eps_k[k];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    P.mu[(size_t)k * D + d] = one ? 1.0f : 0.0f;
    P.eps[(size_t)k * D + d] = e;
  }
  for (int c = threadIdx.x; c < Ws; c += blockDim.x) {
    uint32_t m = 0;
    if (one) {
      const int lo = c * 32;
      const int n = D - lo;
      m = n >= 32 ? 0xffffffffu : (n > 0 ? ((1u << n) - 1u) : 0u);
    }
    P.m1[(size_t)k * Ws + c] = m;
    P.valid[(size_t)k * Ws + c] = 0xffffffffu;
  }
  const double Lv = (e > kEps) ? log((double)((1.0f - e) / e)) : INFINITY;
  const double l1 = (e > kEps) ? log((double)(1.0f - e)) : 0.0;
  if (P.free_disp) {
    for (int d = threadIdx.x; d < Dpad; d += blockDim.x) P.L[(size_t)k * Dpad + d] = d < D ? Lv : 0.0;
  }
  if (threadIdx.x == 0) {
    float pK = 1.0f;
    for (int q = 0; q < K - 1; q++) pK = pK - base_prop;
    const float pkf = (k < K - 1) ? base_prop : pK;
    P.pk[k] = pkf;
    const double lp = ((double)pkf > kEps) ? log((double)pkf) : -INFINITY;
    P.logpk32[k] = (float)lp;
    P.logpk64[k] = lp;
    if (!P.free_disp) P.L[k] = Lv;
    P.eps_k[k] = e;
    P.B[k] = (double)D * l1;
    if (k == 0) { *P.has_half = 0; *P.status = 0; }
  }
}

__global__ void k_init_params(Params P, float base_prop, const float* __restrict__ eps_k) { init_params_body(P, base_prop, eps_k); }

// ------------------------------------------------------------------------------------------------------------------
// Criteria (nem_alg.c:2764-2843), once per run.
//
// k_criteria<GS>: every warp owns a contiguous range of rows and walks it 32/GS rows at a time (a group of GS lanes per
// row, lane = class).  L = sum log f_i and Z = -sum log z_i are summed in fp64 (per-block partials -> k_crit_reduce).
// D and G are the two criteria PPanGGOLiN's "log-likelihood" M is made of, and the reference accumulates them in
// float32, sequentially over rows then classes (nem_alg.c:2813-2823): at N = 20 000 that rounding alone moves M by
// 2e-5 relative (measured with the oracle), more than the parity bar.  In reference-arithmetic mode each warp therefore
// also writes its non-zero (d_ik, g_ik) terms, in (row, class) order, into its own region (adding 0.0f is the identity, so
// the zero terms -- all but about one per row -- can be dropped), and k_crit_chain re-adds the regions in order.
// ------------------------------------------------------------------------------------------------------------------
struct CritArgs {
  long long N;
  int K;
  const double* logf;
  const float* pk;
  const float* logpk32;
  const double* logpk64;
  const float* T;
  const int* rowptr;
  const int* col;
  const float* w;
  float beta;
  int logspace;
  double* partial;   // [gridDim.x][4]  D, G, L, Z
  float2* terms;     // reference arithmetic: per-warp regions of compacted (d_ik, g_ik), region w at w * rows_per_warp * K
  int* counts;       // [total warps] terms in each region
  long long rows_per_warp;
  double2* row_lz;   // reference arithmetic: per row (log f_i, -log z_i), the terms of the float32 L and Z sums
};

template <int GS>
__device__ __forceinline__ void criteria_body(const CritArgs& a) {
  __shared__ double s_part[8][4];
  constexpr int RPW = 32 / GS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane & (GS - 1), grp = lane / GS;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  const int K = a.K;
  const int k = sub < K ? sub : K - 1;
  const long long row_lo = warp_global * a.rows_per_warp;
  const long long row_hi = min(a.N, row_lo + a.rows_per_warp);
  float2* region = a.terms ? a.terms + (size_t)row_lo * K : nullptr;
  int cnt = 0;
  double accD = 0, accG = 0, accL = 0, accZ = 0;
  // the neighbour lists are read eight entries at a time (all column / weight loads in flight, then all posterior
  // gathers, then the adds in list order), and the list bounds of the group's next row are requested a row ahead
  int nx0 = 0, nx1 = 0;
  if (a.rowptr != nullptr && row_lo + grp < row_hi) { nx0 = a.rowptr[row_lo + grp]; nx1 = a.rowptr[row_lo + grp + 1]; }
  for (long long i0 = row_lo; i0 < row_hi; i0 += RPW) {
    const long long i = i0 + grp;
    const bool valid = i < row_hi;
    const long long ii = valid ? i : row_lo;
    const float cik = a.T[(size_t)ii * K + k];
    const double lf = a.logf[(size_t)ii * K + k];
    float pik = 0.0f;
    if (a.rowptr != nullptr) {
      const int n0 = nx0, n1 = valid ? nx1 : nx0;
      if (i + RPW < row_hi) { nx0 = a.rowptr[i + RPW]; nx1 = a.rowptr[i + RPW + 1]; }
      for (int nb = n0; nb < n1; nb += 8) {
        int cc[8];
        float ww[8], tv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { cc[u] = 0; ww[u] = 0.0f; if (nb + u < n1) { cc[u] = a.col[nb + u]; ww[u] = a.w[nb + u]; } }
#pragma unroll
        for (int u = 0; u < 8; u++) { tv[u] = 0.0f; if (nb + u < n1) tv[u] = a.T[(size_t)cc[u] * K + k]; }
#pragma unroll
        for (int u = 0; u < 8; u++) if (nb + u < n1) pik = __fadd_rn(pik, __fmul_rn(ww[u], tv[u]));
      }
    }
    double dik = 0.0, gik = 0.0, term_l, zterm;
    if (!a.logspace) {
      const float lf32 = (lf == -INFINITY) ? -FLT_MAX : (float)lf;
      const double f = (lf == -INFINITY) ? 0.0 : exp((double)lf32);
      const double pkf = (double)a.pk[k] * f;
      const float lpf32 = __fadd_rn(a.logpk32[k], lf32);
      if (cik > FLT_MIN) {
        dik = (double)(float)((double)cik * ((double)lpf32 - log((double)cik)));
        gik = (double)__fmul_rn(cik, pik);
      }
      term_l = pkf;
      zterm = exp((double)__fmul_rn(a.beta, pik));
    } else {
      const double av = a.logpk64[k] + lf;
      if (cik > FLT_MIN) {
        dik = (double)cik * (av - log((double)cik));
        gik = (double)cik * (double)pik;
      }
      term_l = av;
      zterm = exp((double)a.beta * (double)pik);
    }
    if (sub >= K || !valid) { dik = 0; gik = 0; zterm = 0; term_l = a.logspace ? -INFINITY : 0.0; }
    double rowD = 0, rowG = 0, rowL, rowZ;
    if (!a.logspace) {
      double fi = 0.0;
      float zi = 0.0f;
      for (int kk = 0; kk < K; kk++) {
        fi = fi + __shfl_sync(0xffffffffu, term_l, kk, GS);
        zi = (float)((double)zi + __shfl_sync(0xffffffffu, zterm, kk, GS));
        rowD += __shfl_sync(0xffffffffu, dik, kk, GS);
        rowG += __shfl_sync(0xffffffffu, gik, kk, GS);
      }
      rowL = log(fi);
      rowZ = -log((double)zi);
      if (region != nullptr) {      // ordered compaction of the non-zero terms: lanes are in (row, class) order
        const bool nz = valid && sub < K && (dik != 0.0 || gik != 0.0);
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) region[cnt + __popc(m & ((1u << lane) - 1u))] = make_float2((float)dik, (float)gik);
        cnt += __popc(m);
      }
    } else {
      double m = term_l;
#pragma unroll
      for (int off = GS >> 1; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off, GS));
      double sum = 0.0, zi = 0.0;
      const double e = (sub < K && valid && m > -INFINITY) ? exp(term_l - m) : 0.0;
      for (int kk = 0; kk < K; kk++) {
        sum += __shfl_sync(0xffffffffu, e, kk, GS);
        zi += __shfl_sync(0xffffffffu, zterm, kk, GS);
        rowD += __shfl_sync(0xffffffffu, dik, kk, GS);
        rowG += __shfl_sync(0xffffffffu, gik, kk, GS);
      }
      rowL = (m > -INFINITY) ? m + log(sum) : -INFINITY;
      rowZ = -log(zi);
    }
    if (valid && sub == 0) {
      accD += rowD; accG += rowG; accL += rowL; accZ += rowZ;
      if (a.row_lz != nullptr) a.row_lz[i] = make_double2(rowL, rowZ);
    }
  }
  if (a.counts != nullptr && lane == 0) a.counts[warp_global] = cnt;
  // the group leaders' partial sums -> warp -> block (fixed order)
  for (int off = 16; off > 0; off >>= 1) {
    accD += __shfl_xor_sync(0xffffffffu, accD, off);
    accG += __shfl_xor_sync(0xffffffffu, accG, off);
    accL += __shfl_xor_sync(0xffffffffu, accL, off);
    accZ += __shfl_xor_sync(0xffffffffu, accZ, off);
  }
  if (lane == 0) { s_part[wid][0] = accD; s_part[wid][1] = accG; s_part[wid][2] = accL; s_part[wid][3] = accZ; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += s_part[q][threadIdx.x];
    a.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
}

template <int GS>
__global__ void __launch_bounds__(256) k_criteria(CritArgs a) { criteria_body<GS>(a); }

// exclusive prefix sum of the region counts (one warp; a few thousand regions)
__device__ __forceinline__ void crit_offsets_body(const int* __restrict__ counts, int n, int* __restrict__ offsets /*[n+1]*/,
                                                  int* __restrict__ n_rows_out, int n_rows) {
  const int lane = threadIdx.x;
  if (lane == 0) *n_rows_out = n_rows;
  // lane l owns the contiguous regions [l seg, (l + 1) seg): its total (independent loads), a scan over the 32 totals,
  // then the lane's own running sums -- two passes over the counts instead of a chain of n / 32 dependent steps
  const int seg = (n + 31) / 32;
  const int b = min(n, lane * seg), e = min(n, b + seg);
  int tot = 0;
#pragma unroll 8
  for (int r = b; r < e; r++) tot += counts[r];
  int incl = tot;
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  int run = incl - tot;
#pragma unroll 8
  for (int r = b; r < e; r++) { const int v = counts[r]; offsets[r] = run; run += v; }
  if (lane == 31) offsets[n] = incl;
}

__global__ void __launch_bounds__(32) k_crit_offsets(const int* __restrict__ counts, int n, int* __restrict__ offsets,
                                                     int* __restrict__ n_rows_out, int n_rows) {
  crit_offsets_body(counts, n, offsets, n_rows_out, n_rows);
}

// regions -> one dense array, order preserved (one warp per region)
__device__ __forceinline__ void
crit_compact_body(const float2* __restrict__ terms, const int* __restrict__ counts, const int* __restrict__ offsets, int n_regions,
                  long long rows_per_warp, int K, float2* __restrict__ dense) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_regions) return;
  const int lane = threadIdx.x & 31;
  const int n = counts[r];
  const float2* src = terms + (size_t)r * rows_per_warp * K;
  float2* dst = dense + offsets[r];
  for (int t = lane; t < n; t += 32) dst[t] = src[t];
}

__global__ void __launch_bounds__(256)
k_crit_compact(const float2* __restrict__ terms, const int* __restrict__ counts, const int* __restrict__ offsets, int n_regions,
               long long rows_per_warp, int K, float2* __restrict__ dense) {
  crit_compact_body(terms, counts, offsets, n_regions, rows_per_warp, K, dense);
}

// ---- the same float32 sequential sums, computed in parallel and still bit-exact ------------------------------------------
// A dependent chain of float adds runs at one add per 2-3 ns: 3.2 ms for the 10^6 terms of a K = 20 instance, more than
// its ten EM iterations.  But while the running sum s stays inside one binade [2^E, 2^(E+1)), every partial sum is a
// multiple of u = ulp(s) = 2^(E-23), and fl(s + x) = s + rn_u(x): the rounded addend rn_u(x) (x rounded to the nearest
// multiple of u) does not depend on s -- except on an exact tie, where round-to-even looks at s.  So, for the binade the
// accumulator is in, the integers r_i = rn_u(x_i) / u of the terms can be computed in parallel, summed per chunk with
// integer prefix sums (which also give the smallest and largest partial sum inside the chunk), and the state advances
// over whole chunks by integer addition for as long as the partial sums stay inside the binade and no tie occurs.
// Two kinds of chains.  D and G (nem_alg.c:2813-2823) add FLOAT terms to a float accumulator: fl32(s + x).  L and Z
// (nem_alg.c:2826-2827: `cL = cL + log(fi)`, `cZ = cZ - log(zi)`) add DOUBLE terms: the sum is formed in double and stored
// to the float accumulator, fl32(fl64(s + x)) -- at 50 000 rows the drift of that float accumulator is 1.2e-5 of M
// (measured against the reference on BASELINE configs[2]), so it is reproduced as well.  For double terms the integer
// shortcut also excludes terms within 2^-26 u of a rounding tie (where the double rounding could matter).
struct TermsF2 {
  const float2* p;
  static constexpr bool kDouble = false;
  __device__ __forceinline__ void get(int t, double& a, double& b) const { const float2 v = p[t]; a = (double)v.x; b = (double)v.y; }
  static __device__ __forceinline__ float add(float acc, double x) { return __fadd_rn(acc, (float)x); }
};
struct TermsD2 {
  const double2* p;
  static constexpr bool kDouble = true;
  __device__ __forceinline__ void get(int t, double& a, double& b) const { const double2 v = p[t]; a = v.x; b = v.y; }
  static __device__ __forceinline__ float add(float acc, double x) { return (float)__dadd_rn((double)acc, x); }
};

constexpr int kCcChunk = 256;     // terms per chunk: a chunk that cannot be skipped (crossing, tie) is added term by term
struct CcState { float s[2]; int cur[2]; int rounds; int seq_chunks[2]; };   // per chain (0: D, 1: G): exact running sum, next chunk

__device__ __forceinline__ bool cc_binade(float s, int& E) {   // normal, non-zero, finite
  if (!(fabsf(s) >= 1.17549435e-38f) || !isfinite(s)) return false;
  E = ilogbf(s);
  return true;
}

// rn_u(x) / u for u = 2^Eu, as an integer, by integer arithmetic on the bits of x (a double; float terms are exact in it):
// x = +-M 2^e with a 53-bit M.  `needseq` is raised for what the integer shortcut must not decide: an exact rounding tie
// (round-to-even looks at the running sum), for double terms also anything within 2^-26 u of a tie (where the double
// rounding of fl32(fl64(s + x)) could matter), a term of 2^30 u or more, an infinity or a NaN.
__device__ __forceinline__ long long cc_round_units(double x, int Eu, bool double_terms, bool& needseq) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
  const int ex = (int)((bits >> 52) & 0x7ffu);
  const unsigned long long frac = bits & 0xfffffffffffffULL;
  if (ex == 0x7ff) { needseq = true; return 0; }
  if (ex == 0 && frac == 0) return 0;
  const unsigned long long M = ex ? (frac | (1ULL << 52)) : frac;
  const int e = (ex ? ex : 1) - 1075;                    // x = +-M 2^e
  if (ex - 1023 - Eu >= 30) { needseq = true; return 0; }
  const int sh = e - Eu;
  long long r;
  if (sh >= 0) {
    r = (long long)(M << sh);
  } else {
    const int sdown = -sh;
    if (sdown >= 63) return 0;                           // |x| < u / 2^9: rounds to zero, far from a tie
    const unsigned long long fl = M >> sdown, rem = M & ((1ULL << sdown) - 1ULL), half = 1ULL << (sdown - 1);
    const unsigned long long dist = rem > half ? rem - half : half - rem;
    const unsigned long long window = double_terms ? (half >> 25) : 0ULL;
    if (dist <= window && (dist == 0 || double_terms)) needseq = true;
    r = (long long)(fl + (rem > half ? 1ULL : 0ULL));
  }
  return (bits >> 63) ? -r : r;
}

// The scheme as it runs (four small launches, every one over all chains of a batch of instances -- blockIdx.y = job):
//   k_cc_sums     per chunk of 512 terms its fp64 sum                                              (all SMs)
//   k_cc_predict  prefix sums of those -> for every chunk the binade the float accumulator will be in when it gets there
//                 (the fp64 prefix differs from the float state by ~1e-6 relative: wrong only next to a binade crossing)
//   k_cc_records  per chunk, for the predicted u: integer total, extreme partial sums, flags         (all SMs, once per chunk)
//   k_cc_walk     one warp per chain walks the chunks with the EXACT state: 32 chunks per step by integer prefix sums
//                 while the records were made for the state's binade and the partial sums stay inside it; any other
//                 chunk (a crossing, a tie, a misprediction, a zero / subnormal state) is added term by term
// In the worst case it degrades to plain sequential adds, never to another result (tests compare it bit for bit with
// k_crit_chain and a numpy loop on adversarial inputs).
struct CcJob {
  const void* terms;            // float2* (D, G: compacted terms) or double2* (L, Z: one pair per row)
  const int* total_ptr;         // number of terms
  long long* recR; long long* recMn; long long* recMx;   // [nchunks][2]
  double* csum;                 // [nchunks][2] fp64 chunk sums
  int* recFlag;                 // [nchunks][2]; bit 0: needs the sequential path, bit 1: all terms zero
  int* epred;                   // [nchunks][2] binade the records were made for (INT_MIN: none)
  CcState* state;
  double* out2;                 // the two sums
  const unsigned* run_state;    // batched runs: kSt* words of the instance (a failed run is skipped); may be null
  int double_terms;             // 0: float2 terms (D, G), 1: double2 terms (L, Z)
};

__device__ __forceinline__ bool cc_job(const CcJob* __restrict__ jobs, const CcJob& job0, CcJob& job) {
  job = jobs != nullptr ? jobs[blockIdx.y] : job0;
  return !(job.run_state != nullptr && job.run_state[kStStatus] != 0u);
}

template <class Terms>
__device__ __forceinline__ void cc_sums_body(const CcJob& job) {
  Terms dense{reinterpret_cast<decltype(Terms::p)>(job.terms)};
  const int total = *job.total_ptr;
  const int nchunks = (total + kCcChunk - 1) / kCcChunk;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  for (int c = warp; c < nchunks; c += nwarps) {
    double sa = 0.0, sb = 0.0;
    const int t_end = min(total, (c + 1) * kCcChunk);
#pragma unroll 4
    for (int t = c * kCcChunk + lane; t < t_end; t += 32) { double a, b; dense.get(t, a, b); sa += a; sb += b; }
    for (int off = 16; off > 0; off >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, off); sb += __shfl_xor_sync(0xffffffffu, sb, off); }
    if (lane == 0) { job.csum[(size_t)c * 2] = sa; job.csum[(size_t)c * 2 + 1] = sb; }
  }
}
__global__ void __launch_bounds__(256) k_cc_sums(const CcJob* __restrict__ jobs, CcJob job0) {
  CcJob job;
  if (!cc_job(jobs, job0, job)) return;
  if (job.double_terms) cc_sums_body<TermsD2>(job); else cc_sums_body<TermsF2>(job);
}

// one warp per chain: exclusive fp64 prefix of the chunk sums -> predicted binade at the start of every chunk
__global__ void __launch_bounds__(64) k_cc_predict(const CcJob* __restrict__ jobs, CcJob job0) {
  CcJob job;
  if (!cc_job(jobs, job0, job)) return;
  const int total = *job.total_ptr;
  const int nchunks = (total + kCcChunk - 1) / kCcChunk;
  const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
  double carry = 0.0;
  for (int c0 = 0; c0 < nchunks; c0 += 32) {
    const int c = c0 + lane;
    const double v = c < nchunks ? job.csum[(size_t)c * 2 + ch] : 0.0;
    double incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const double o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    const double start = carry + incl - v;
    if (c < nchunks) {
      const float sf = (float)start;
      int E = INT_MIN;
      if (fabsf(sf) >= 1.17549435e-38f && isfinite(sf)) E = ilogbf(sf);
      job.epred[(size_t)c * 2 + ch] = E;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

template <class Terms>
__device__ __forceinline__ void cc_records_body(const CcJob& job) {
  Terms dense{reinterpret_cast<decltype(Terms::p)>(job.terms)};
  const int total = *job.total_ptr;
  const int nchunks = (total + kCcChunk - 1) / kCcChunk;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  for (int c = warp; c < nchunks; c += nwarps) {
    // lane l owns the 16 consecutive terms [16 l, 16 l + 16) of the chunk: a lane-local running sum with its extremes,
    // then one scan over the 32 lane totals
    constexpr int kPer = kCcChunk / 32;
    const int t_end = min(total, (c + 1) * kCcChunk);
    const int t_lane = c * kCcChunk + lane * kPer;
    double vx[kPer], vy[kPer];
#pragma unroll
    for (int j = 0; j < kPer; j++) { vx[j] = 0.0; vy[j] = 0.0; if (t_lane + j < t_end) dense.get(t_lane + j, vx[j], vy[j]); }
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      const int E = job.epred[(size_t)c * 2 + ch];
      const bool okc = E != INT_MIN;
      bool nz = false, needseq = false;
      long long pre = 0, lmn = LLONG_MAX, lmx = LLONG_MIN;
#pragma unroll
      for (int j = 0; j < kPer; j++) {
        const double x = ch ? vy[j] : vx[j];
        nz = nz || (x != 0.0);
        if (okc) {
          pre += cc_round_units(x, E - 23, Terms::kDouble, needseq);     // u = 2^(E - 23)
          if (t_lane + j < t_end) { lmn = min(lmn, pre); lmx = max(lmx, pre); }
        }
      }
      int flag = 2;
      if (__any_sync(0xffffffffu, nz)) flag &= ~2;
      if (!okc || __any_sync(0xffffffffu, needseq)) flag |= 1;
      long long incl = pre;                                              // inclusive scan of the lane totals
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const long long o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
      const long long base = incl - pre;
      long long lo = (lmn == LLONG_MAX) ? LLONG_MAX : base + lmn, hi = (lmx == LLONG_MIN) ? LLONG_MIN : base + lmx;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
      }
      const long long run = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 0) {
        const size_t i = (size_t)c * 2 + ch;
        job.recR[i] = run; job.recMn[i] = lo; job.recMx[i] = hi; job.recFlag[i] = flag;
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_cc_records(const CcJob* __restrict__ jobs, CcJob job0) {
  CcJob job;
  if (!cc_job(jobs, job0, job)) return;
  if (job.double_terms) cc_records_body<TermsD2>(job); else cc_records_body<TermsF2>(job);
}

template <class Terms>
__device__ __forceinline__ void cc_walk_body(const CcJob& job) {
  __shared__ double s_x[2][kCcChunk];
  Terms dense{reinterpret_cast<decltype(Terms::p)>(job.terms)};
  const int total = *job.total_ptr;
  const int nchunks = (total + kCcChunk - 1) / kCcChunk;
  const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
  float sacc = 0.0f;
  int cur = 0, nseq = 0;
  // the records of the 32 chunks after the current step are requested while the current step is evaluated
  struct Rec { long long R, mn, mx; int fl, Ep; };
  auto load = [&](int c) {
    Rec r{0, 0, 0, 1, INT_MIN};
    if (c < nchunks) {
      const size_t i = (size_t)c * 2 + ch;
      r.R = job.recR[i]; r.mn = job.recMn[i]; r.mx = job.recMx[i]; r.fl = job.recFlag[i]; r.Ep = job.epred[i];
    }
    return r;
  };
  Rec nxt = load(lane);
  int nxt_at = 0;
  while (cur < nchunks) {
    int Ec = 0;
    const bool have = cc_binade(sacc, Ec);
    const int c = cur + lane;
    const Rec rec = (nxt_at == cur) ? nxt : load(c);
    nxt = load(c + 32);
    nxt_at = cur + 32;
    int nvalid;
    if (have) {
      const long long S0 = (long long)scalbn((double)sacc, 23 - Ec);   // exact integer, 2^23 <= |S0| < 2^24
      long long R = rec.R;
      int fl = rec.fl;
      if (rec.Ep != Ec) { R = 0; fl |= 1; }                              // records made for another u: not usable
      long long incl = R;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const long long o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
      const long long Sstart = S0 + incl - R;
      const long long lo = Sstart + rec.mn, hi = Sstart + rec.mx;
      bool okc = !(fl & 1) && c < nchunks;
      if (fl & 2) okc = c < nchunks;                                     // an all-zero chunk leaves any state alone
      else if (S0 > 0) okc = okc && lo >= (1LL << 23) && hi <= (1LL << 24) - 1;
      else             okc = okc && hi <= -(1LL << 23) && lo >= -((1LL << 24) - 1);
      const unsigned m = __ballot_sync(0xffffffffu, !okc);
      nvalid = m ? __ffs(m) - 1 : 32;
      if (nvalid > 0) {
        const long long Send = S0 + __shfl_sync(0xffffffffu, incl, nvalid - 1);
        sacc = (float)scalbn((double)Send, Ec - 23);                     // exact: |Send| < 2^24, same binade
      }
    } else {                                                             // zero / subnormal / non-finite state: only all-zero chunks are free
      const bool z = c < nchunks && (rec.fl & 2);
      const unsigned m = __ballot_sync(0xffffffffu, !z);
      nvalid = m ? __ffs(m) - 1 : 32;
    }
    cur += nvalid;
    if (nvalid == 32 || cur >= nchunks) continue;
    {   // chunk `cur` term by term: the warp stages it in shared memory, one lane runs the dependent chain of adds
      const int t_begin = cur * kCcChunk, t_end = min(total, t_begin + kCcChunk);
      const int count = t_end - t_begin;
#pragma unroll
      for (int b = 0; b < kCcChunk / 32; b++) {
        const int t = t_begin + b * 32 + lane;
        double v = 0.0;
        if (t < t_end) { double a, bb; dense.get(t, a, bb); v = ch ? bb : a; }
        s_x[ch][b * 32 + lane] = v;
      }
      __syncwarp();
      float acc = sacc;
      if (lane == 0) {
        int q = 0;
        for (; q + 8 <= count; q += 8) {
          double x[8];
#pragma unroll
          for (int u = 0; u < 8; u++) x[u] = s_x[ch][q + u];
#pragma unroll
          for (int u = 0; u < 8; u++) acc = Terms::add(acc, x[u]);
        }
        for (; q < count; q++) acc = Terms::add(acc, s_x[ch][q]);
      }
      sacc = __shfl_sync(0xffffffffu, acc, 0);
      __syncwarp();
      cur++; nseq++;
    }
  }
  if (lane == 0) {
    job.out2[ch] = (double)sacc;
    if (job.state != nullptr) { job.state->s[ch] = sacc; job.state->cur[ch] = cur; job.state->seq_chunks[ch] = nseq; job.state->rounds = 1; }
  }
}
__global__ void __launch_bounds__(64) k_cc_walk(const CcJob* __restrict__ jobs, CcJob job0) {
  CcJob job;
  if (!cc_job(jobs, job0, job)) return;
  if (job.double_terms) cc_walk_body<TermsD2>(job); else cc_walk_body<TermsF2>(job);
}

// D (warp 0) and G (warp 1) re-accumulated in float32 in the reference's order.  The warp loads 32 terms at a time
// (coalesced, one block ahead of the adds) and hands them to the dependent chain of adds through shuffles.
template <class Terms>
__global__ void __launch_bounds__(64)
k_crit_chain(Terms dense, const int* __restrict__ total_ptr, double* __restrict__ out2) {
  const int lane = threadIdx.x & 31, which = threadIdx.x >> 5;   // 0: D (L), 1: G (Z)
  const int total = *total_ptr;
  auto fetch = [&](int t) -> double {
    if (t < total) { double a, b; dense.get(t, a, b); return which == 0 ? a : b; }
    return 0.0;
  };
  float acc = 0.0f;
  double x = fetch(lane), x1 = fetch(32 + lane);
  for (int t0 = 0; t0 < total; t0 += 32) {
    const double x2 = fetch(t0 + 64 + lane);      // two blocks ahead of the chain
    const int m = min(32, total - t0);
    for (int q = 0; q < m; q++) acc = Terms::add(acc, __shfl_sync(0xffffffffu, x, q));
    x = x1; x1 = x2;
  }
  if (lane == 0) out2[which] = (double)acc;
}

// fp64 sum of n values (stride apart) in index order: a warp loads 32 at a time, lane 0's adds take them by shuffle --
// the order of a sequential loop without its load latency per element
__device__ __forceinline__ double ordered_sum(const double* __restrict__ p, int n, int stride) {
  const int lane = threadIdx.x & 31;
  double t = 0.0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const double v = i < n ? p[(size_t)i * stride] : 0.0;
    const int m = min(32, n - i0);
    for (int j = 0; j < m; j++) t += __shfl_sync(0xffffffffu, v, j);
  }
  return t;
}

// launched with 4 warps: out4[w] = fp64 sum of the block partials of criterion w (first = 0: all four; first = 2: only L and Z)
__global__ void k_crit_reduce(const double* __restrict__ partial, int nblocks, double* __restrict__ out4, int first) {
  const int w = threadIdx.x >> 5;
  if (blockIdx.x != 0 || w >= 4 || w < first) return;
  const double t = ordered_sum(partial + w, nblocks, 4);
  if ((threadIdx.x & 31) == 0) out4[w] = t;
}

// ------------------------------------------------------------------------------------------------------------------
// Labels + entropy from "%5.3f"-rounded posteriors (partition.py:206-231, nem_exe.c:1682).
// v * 1000 is exact in double for a float32 v (24 + 10 significant bits), so rint() (ties to even) reproduces printf.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_labels(const float* __restrict__ T, long long N, int K, int* __restrict__ labels, double* __restrict__ ent_partial) {
  __shared__ double s_e[8];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double ent = 0.0;
  if (i < N) {
    double mx = -1.0;
    int arg = -1, nmax = 0;
    for (int k = 0; k < K; k++) {
      const double v = rint((double)T[(size_t)i * K + k] * 1000.0) / 1000.0;
      if (v > 0) ent += log(v) * v;
      if (v > mx) { mx = v; arg = k; nmax = 1; }
      else if (v == mx) nmax++;
    }
    labels[i] = (nmax > 1 || mx < 0.5) ? -1 : arg;
  }
  for (int off = 16; off > 0; off >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, off);
  if ((threadIdx.x & 31) == 0) s_e[threadIdx.x >> 5] = ent;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += s_e[q];
    ent_partial[blockIdx.x] = t;
  }
}

__global__ void k_sum_partials(const double* __restrict__ p, int n, double* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const double t = ordered_sum(p, n, 1);
  if (threadIdx.x == 0) *out = t;
}


// ------------------------------------------------------------------------------------------------------------------
// Batched, device-resident EM for small instances (the K-selection sweep of partition.py:482-514 and the chunk samples of
// partition.py:837-846: thousands of instances of <= 100 000 x 500).  One launch per EM phase covers ALL instances of a
// group (blockIdx.z = instance, operands from a descriptor array in device memory), the convergence test of
// nem_alg.c:1868-1892 runs on the device (k_b_poll) and every kernel of a later iteration returns at once for an
// instance that is no longer active -- so the host enqueues iterations without ever waiting for one, and only looks at an
// "any instance still active" word one iteration late.  The arithmetic is the single-instance kernels' (same bodies).
// ------------------------------------------------------------------------------------------------------------------

struct BatchInst {
  const uint32_t* bits;        // this rank's rows of the matrix
  long long N;                 // rows of the instance: the sweep and the criteria run over all of them (replicated)
  long long N_loc, row_off;    // this rank's row block [row_off, row_off + N_loc) (N_loc == N, row_off == 0 without sharding)
  // row-sharded runs (one process per GPU): the two exchange steps of an iteration over NVLink peer memory
  int world, rank, epoch0;     // epochs of run-local iteration s: epoch0 + s (the same on every rank)
  int* const* sig_ptrs; const int* sig; int* peer_err; int* bcast_ticket;
  PeerStats ps;                // the ranks' statistic blocks (ps.n == 1 without sharding)
  PeerOut push;                // the ranks' log-density buffers
  PeerParams pp;               // the ranks' parameter blocks (column-sliced finalize)
  double* logf_loc;            // = logf + row_off * K
  int D, Ws, Dpad, K, free_disp;
  const int* rowptr; const int* col; const float* w;
  const int* level_ptr; const int* level_rows; int n_levels;
  int4* ell_hdr; int2* ell_nb;
  float beta, cv_th;
  int itermax, c0;             // c0: posterior buffer that holds the start posteriors (nem_alg.c:2023-2065)
  int logspace, jacobi, full_recount;
  Params P;
  int* cnt; double* fsum; int* nk_cnt; double* nk_fz; int* hist; int* offsets; int* cursor; int* prev; int2* ent; int* perm;
  double* Ltab; double* Lword; double* fin_partial; int* fin_tickets;
  double* logf;
  float* T[2];
  unsigned* state;             // kSt* words; P.status aliases state + kStStatus
  unsigned* barrier;           // InstBar counter of the sweep
  int* chg; int* spec_words;   // speculative sweep: per-row change marks [2][N], bookkeeping words [8] (null: level sweeps only)
  int spec_deep;               // schedules of at least this many levels are always swept speculatively (0: never forced)
  char* zero_from; long long zero_bytes;   // statistics cleared before every M-step
  float base_prop;
  float eps_init[NEMB200_MAX_K];
  // criteria (nem_alg.c:2764-2843), computed once at the end of the run
  float* logpk32_unused;       // (P carries logpk32 / logpk64)
  double* crit_partial; float* crit_terms; int* crit_counts; int* crit_offsets; float* crit_dense; double* crit_lz;
  int* crit_nrows; double* crit4;
  // outputs of the run (device pointers of the caller, may be null)
  float* T_out; float* mu_out; float* eps_out; float* pk_out;
  unsigned long long* dbg;     // NEMB200_TRACE: in-kernel phase stamps (null otherwise)
  char* ll[kMaxPeers];         // row-sharded: every rank's LL region (ll[rank] = this rank's own), laid out by `lay`
  LLayout lay;
};

__device__ __forceinline__ int b_final_buffer(const BatchInst& I) {
  // a failed iteration does not replace the posteriors (nem_alg.c:1873-1892)
  return (I.c0 + (int)I.state[kStIter] - (I.state[kStStatus] != 0u ? 1 : 0)) & 1;
}

__device__ __forceinline__ bool b_active(const BatchInst& I) { return I.state[kStActive] != 0u; }

// start of a run: initial parameters, "nothing counted yet", zero posteriors (ClassifM is calloc'ed, nem_exe.c:533-536)
__global__ void __launch_bounds__(256) k_b_start(const BatchInst* __restrict__ B) {
  const BatchInst& I = B[blockIdx.z];
  if (blockIdx.x < (unsigned)I.K) {
    Params P = I.P;
    // init_params_body works per class with blockIdx.x = class
    init_params_body(P, I.base_prop, I.eps_init);
  }
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long t = tid; t < I.N_loc; t += nth) I.prev[t] = kNone;
  const long long nkd = (long long)I.K * I.Dpad;
  for (long long t = tid; t < nkd; t += nth) I.cnt[t] = 0;
  const long long ntk = I.N * I.K;
  float* T0 = I.T[0];
  for (long long t = tid; t < ntk; t += nth) T0[t] = 0.0f;
  for (long long t = tid; t < I.zero_bytes / 4; t += nth) reinterpret_cast<int*>(I.zero_from)[t] = 0;
  if (tid == 0) {
    I.state[kStMaxdiff] = 0u; I.state[kStActive] = I.itermax >= 1 ? 1u : 0u; I.state[kStIter] = 0u; I.state[kStConverged] = 0u;
    *I.barrier = 0u;
    if (I.spec_words != nullptr) {
      for (int q = 0; q < 8; q++) I.spec_words[q] = 0;
      I.spec_words[3] = 0x3fffffff;      // "the previous sweep changed everything": the first iteration sweeps by levels
    }
  }
  if (tid < I.K) I.fin_tickets[tid] = 0;
}

__global__ void __launch_bounds__(256) k_b_ltab(const BatchInst* __restrict__ B, int gate) {
  const BatchInst& I = B[blockIdx.z];
  if (gate && !b_active(I)) return;
  ltab_body(I.P.L, I.K, I.Ws, I.Ltab, I.Lword);
}

__global__ void __launch_bounds__(256) k_b_ell_build(const BatchInst* __restrict__ B) {
  const BatchInst& I = B[blockIdx.z];
  if (I.ell_hdr == nullptr) return;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I.N) return;
  const int i = I.level_rows[idx];
  const int n0 = I.rowptr[i], deg = I.rowptr[i + 1] - n0;
  I.ell_hdr[idx] = make_int4(i, deg, n0, 0);
#pragma unroll
  for (int u = 0; u < kEllW; u++)
    I.ell_nb[idx * kEllW + u] = u < deg ? make_int2(I.col[n0 + u], __float_as_int(I.w[n0 + u])) : make_int2(0, 0);
}

template <int KB>
__global__ void __launch_bounds__(256) k_b_classify(const BatchInst* __restrict__ B, int s) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I) || (long long)blockIdx.x * 256 >= I.N_loc) return;
  classify_body<KB>(I.T[(I.c0 + s - 1) & 1] + I.row_off * I.K, I.N_loc, I.K, I.prev, I.ent, I.hist, I.nk_cnt, I.nk_fz, I.full_recount);
}

__global__ void __launch_bounds__(256) k_b_scatter(const BatchInst* __restrict__ B) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I) || (long long)blockIdx.x * 256 >= I.N_loc) return;
  scatter_body(I.ent, I.N_loc, I.hist, 2 * I.K + 1, I.offsets, I.cursor, I.perm);
}

__global__ void __launch_bounds__(kMsTpb) k_b_onehot(const BatchInst* __restrict__ B, int tile_w) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I)) return;
  mstep_onehot_body(reinterpret_cast<const uint2*>(I.bits), I.Ws / 2, tile_w, I.K, I.offsets, I.perm, I.cnt, I.Dpad);
}

template <int KB>
__global__ void __launch_bounds__(256) k_b_fuzzy(const BatchInst* __restrict__ B, int s) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I)) return;
  mstep_fuzzy_body<KB>(I.bits, I.Ws, I.K, I.D, I.offsets, I.perm, I.T[(I.c0 + s - 1) & 1] + I.row_off * I.K, I.fsum, I.Dpad);
}

template <int KT>
__global__ void __launch_bounds__(256) k_b_fuzzy_mma(const BatchInst* __restrict__ B, int s) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I)) return;
  mstep_fuzzy_mma_body<KT>(I.bits, I.Ws, I.K, I.D, I.offsets, I.perm, I.T[(I.c0 + s - 1) & 1] + I.row_off * I.K, I.fsum, I.Dpad);
}

// row-sharded runs (world > 1): the two flag barriers of the M-step and the class scalars are part of this launch (FinSync)
__global__ void __launch_bounds__(256) k_b_finalize(const BatchInst* __restrict__ B, int s) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I) || blockIdx.y >= (unsigned)I.K) return;
  if (I.world > 1) {
    FinSync fs;
    fs.sig_ptrs = I.sig_ptrs; fs.sig = I.sig; fs.err = I.peer_err; fs.ticket = I.fin_tickets;   // (the per-class tickets are unused here)
    fs.world = I.world; fs.rank = I.rank; fs.epoch = I.epoch0 + s; fs.dbg = I.dbg;
    mstep_finalize_body(I.P, I.ps, I.N, I.fin_partial, I.fin_tickets, &I.pp, &fs);
  } else {
    mstep_finalize_body(I.P, I.ps, I.N, I.fin_partial, I.fin_tickets);
  }
}

// Row-sharded M-step in ONE launch without fences: the column-sliced all-reduce of the statistics and the all-gather of
// the parameters both travel as LL words (see LLayout).  grid.x <= the number of co-resident CTAs (every CTA's sends of
// phase 1 are issued before anything waits); work items = (class, chunk of kFinCols columns), chunk c is reduced by rank
// c % world.
//   phase 1  every rank sends its counts / fuzzy sums of every item to the item's owner, and its class sizes to everyone
//   phase 2  the owner adds the ranks' statistics in rank order (same arithmetic as mstep_finalize_body), sends the
//            centre plane words, [skd: dispersions and log-odds] and the chunk's partial sums to every rank
//   phase 3  every rank takes the words of every item into its plain parameter arrays; the last CTA derives the class
//            scalars from the chunk partials (identical on every rank: same values, same order)
__global__ void __launch_bounds__(256, 4) k_b_finalize_ll(const BatchInst* __restrict__ B, int s) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I)) return;
  __shared__ int s_nki[kMaxPeers][NEMB200_MAX_K];
  __shared__ double s_nkf[kMaxPeers][NEMB200_MAX_K];
  __shared__ double s_nk[NEMB200_MAX_K];
  __shared__ double s_red[8][2];
  __shared__ int s_last;
  const Params& P = I.P;
  const int K = I.K, D = I.D, Dpad = I.Dpad, Ws = I.Ws, W = I.world, me = I.rank;
  const int nchunk = (Dpad + kFinCols - 1) / kFinCols;
  const int items = K * nchunk;
  const uint32_t ep = (uint32_t)(I.epoch0 + s);
  const size_t KD = (size_t)K * Dpad;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const LLayout& L = I.lay;
  unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? I.dbg : nullptr;
  dbg_stamp(dbg, 0);
  // ---- phase 1
  if (blockIdx.x == 0 && (int)threadIdx.x < K) {
    const int k = threadIdx.x;
    const int ni = I.nk_cnt[k];
    const double nf = I.nk_fz[k];
    for (int q = 0; q < W; q++) {
      ll_put(reinterpret_cast<uint2*>(I.ll[q] + L.nk_c) + me * 32 + k, (uint32_t)ni, ep);
      ll_put_f64(reinterpret_cast<uint4*>(I.ll[q] + L.nk_f) + me * 32 + k, nf, ep);
    }
  }
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int k = it / nchunk, c = it - k * nchunk;
    const int d = c * kFinCols + threadIdx.x;
    if (d < D) {
      char* dst = I.ll[c % W];
      const size_t at = (size_t)me * KD + (size_t)k * Dpad + d;
      const int ci = __ldcg(I.cnt + (size_t)k * Dpad + d);
      const double cf = __ldcg(I.fsum + (size_t)k * Dpad + d);
      if (cf != 0.0) ll_put_f64(reinterpret_cast<uint4*>(dst + L.in_f) + at, cf, ep);
      ll_put(reinterpret_cast<uint2*>(dst + L.in_c) + at, (uint32_t)ci, cf != 0.0 ? ep : (ep | 0x80000000u));
    }
  }
  dbg_stamp(dbg, 1);
  // class sizes over all ranks (rank order: the same value on every rank)
  if ((int)threadIdx.x < W * K) {
    const int r = threadIdx.x / K, k = threadIdx.x - r * K;
    s_nki[r][k] = (int)ll_get(reinterpret_cast<const uint2*>(I.ll[me] + L.nk_c) + r * 32 + k, ep, I.peer_err).x;
    s_nkf[r][k] = ll_get_f64(reinterpret_cast<const uint4*>(I.ll[me] + L.nk_f) + r * 32 + k, ep, I.peer_err);
  }
  __syncthreads();
  if ((int)threadIdx.x < K) {
    long long ni = 0;
    double nf = 0.0;
    for (int r = 0; r < W; r++) { ni += s_nki[r][threadIdx.x]; nf += s_nkf[r][threadIdx.x]; }
    const double nk = (double)ni + nf;
    s_nk[threadIdx.x] = nk;
    if (!(nk > kEps) && blockIdx.x == 0) atomicMax(P.status, 1);   // empty class: nem_mod.c:1399-1403, parameters kept
  }
  __syncthreads();
  dbg_stamp(dbg, 2);
  // ---- phase 2: the chunks this rank reduces
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int k = it / nchunk, c = it - k * nchunk;
    if (c % W != me) continue;
    const double nk = s_nk[k];
    if (!(nk > kEps)) continue;
    const int d = c * kFinCols + threadIdx.x;
    float m = 0.0f;
    double iner = 0.0, b_add = 0.0;
    if (d < D) {
      const uint2* inc = reinterpret_cast<const uint2*>(I.ll[me] + L.in_c) + (size_t)k * Dpad + d;
      const uint4* inf = reinterpret_cast<const uint4*>(I.ll[me] + L.in_f) + (size_t)k * Dpad + d;
      uint2 cw[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; r++) cw[r] = r < W ? ll_peek(inc + (size_t)r * KD) : make_uint2(0u, ep | 0x80000000u);
      long long c_i = 0;
      double c_f = 0.0;
#pragma unroll
      for (int r = 0; r < kMaxPeers; r++) {           // added in rank order
        if (r < W) {
          if ((cw[r].y & 0x7fffffffu) != ep) cw[r] = ll_get(inc + (size_t)r * KD, ep, I.peer_err);
          c_i += (int)cw[r].x;
          if (!(cw[r].y & 0x80000000u)) c_f += ll_get_f64(inf + (size_t)r * KD, ep, I.peer_err);
        }
      }
      const double s1 = (double)c_i + c_f;
      const double s0 = nk - s1;
      const float w0 = (float)s0, half = (float)nk * 0.5f;   // see mstep_finalize_body
      if (w0 > half) { m = 0.0f; iner = s1; }
      else if (w0 < half) { m = 1.0f; iner = s0; }
      else { m = 0.5f; iner = 0.5 * nk; }
      if (P.free_disp) {
        const float e = (float)(iner / nk);
        double Lv;
        if (e > kEps) { Lv = log((double)((1.0f - e) / e)); b_add = log((double)(1.0f - e)); }
        else Lv = INFINITY;
        for (int q = 0; q < W; q++) {
          ll_put(reinterpret_cast<uint2*>(I.ll[q] + L.out_eps) + (size_t)k * D + d, __float_as_uint(e), ep);
          ll_put_f64(reinterpret_cast<uint4*>(I.ll[q] + L.out_L) + (size_t)k * Dpad + d, Lv, ep);
        }
      }
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, m == 1.0f);
    const unsigned bv = __ballot_sync(0xffffffffu, m != 0.5f);
    if (lane == 0 && (d >> 5) < Ws) {
      for (int q = 0; q < W; q++) {
        ll_put(reinterpret_cast<uint2*>(I.ll[q] + L.out_m1) + (size_t)k * Ws + (d >> 5), b1, ep);
        ll_put(reinterpret_cast<uint2*>(I.ll[q] + L.out_valid) + (size_t)k * Ws + (d >> 5), bv, ep);
      }
    }
    double iner_sum = iner, b_sum = b_add;
    for (int off = 16; off > 0; off >>= 1) {
      iner_sum += __shfl_xor_sync(0xffffffffu, iner_sum, off);
      b_sum += __shfl_xor_sync(0xffffffffu, b_sum, off);
    }
    __syncthreads();                                   // (s_red of the previous item is consumed)
    if (lane == 0) { s_red[wid][0] = iner_sum; s_red[wid][1] = b_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int q = 0; q < 8; q++) { t0 += s_red[q][0]; t1 += s_red[q][1]; }
      for (int q = 0; q < W; q++) {
        uint4* o = reinterpret_cast<uint4*>(I.ll[q] + L.out_part) + ((size_t)k * nchunk + c) * 2;
        ll_put_f64(o, t0, ep);
        ll_put_f64(o + 1, t1, ep);
      }
    }
  }
  dbg_stamp(dbg, 3);
  // ---- phase 3: every item's results into the plain parameter arrays of this rank
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int k = it / nchunk, c = it - k * nchunk;
    if (!(s_nk[k] > kEps)) continue;
    const int d = c * kFinCols + threadIdx.x;
    const int w = d >> 5;
    unsigned b1 = 0u, bv = 0xffffffffu;
    if (w < Ws) {
      if (lane == 0) {
        b1 = ll_get(reinterpret_cast<const uint2*>(I.ll[me] + L.out_m1) + (size_t)k * Ws + w, ep, I.peer_err).x;
        bv = ll_get(reinterpret_cast<const uint2*>(I.ll[me] + L.out_valid) + (size_t)k * Ws + w, ep, I.peer_err).x;
        P.m1[(size_t)k * Ws + w] = b1;
        P.valid[(size_t)k * Ws + w] = bv;
      }
      b1 = __shfl_sync(0xffffffffu, b1, 0);
      bv = __shfl_sync(0xffffffffu, bv, 0);
    }
    if (d < D) {
      const bool is_half = !((bv >> lane) & 1u);
      P.mu[(size_t)k * D + d] = is_half ? 0.5f : (((b1 >> lane) & 1u) ? 1.0f : 0.0f);
      if (is_half) atomicOr(P.has_half, 1);
      if (P.free_disp) {
        P.eps[(size_t)k * D + d] = __uint_as_float(ll_get(reinterpret_cast<const uint2*>(I.ll[me] + L.out_eps) + (size_t)k * D + d, ep, I.peer_err).x);
        P.L[(size_t)k * Dpad + d] = ll_get_f64(reinterpret_cast<const uint4*>(I.ll[me] + L.out_L) + (size_t)k * Dpad + d, ep, I.peer_err);
      }
    } else if (P.free_disp && d < Dpad) {
      P.L[(size_t)k * Dpad + d] = 0.0;
    }
    if (threadIdx.x < 2) {
      const uint4* o = reinterpret_cast<const uint4*>(I.ll[me] + L.out_part) + ((size_t)k * nchunk + c) * 2 + threadIdx.x;
      I.fin_partial[((size_t)k * nchunk + c) * 2 + threadIdx.x] = ll_get_f64(o, ep, I.peer_err);
    }
  }
  dbg_stamp(dbg, 4);
  // ---- class scalars by the last CTA to finish
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = (atomicAdd(I.fin_tickets, 1) == (int)gridDim.x - 1);
    if (s_last) *I.fin_tickets = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int kk = wid; kk < K; kk += (int)(blockDim.x >> 5))
    if (s_nk[kk] > kEps) finalize_class_scalars(P, kk, s_nk[kk], nchunk, I.fin_partial, I.N, lane);
  __syncthreads();
  dbg_stamp(threadIdx.x == 0 ? I.dbg : nullptr, 5);
}

template <int KB, int KX, int RB>
__global__ void __launch_bounds__(kDensTpb) k_b_density_sk(const BatchInst* __restrict__ B, int G, int gate, int planes_in_smem) {
  const BatchInst& I = B[blockIdx.z];
  if (gate && !b_active(I)) return;
  density_sk_cta<KB, KX, RB>(reinterpret_cast<const uint4*>(I.bits), I.N_loc, I.Ws / 4, I.K, G, reinterpret_cast<const uint4*>(I.P.m1),
                             reinterpret_cast<const uint4*>(I.P.valid), I.P.L, I.P.B, I.P.has_half, I.logf_loc, planes_in_smem);
}

// large matrices: the streaming (shared-memory ring) kernels, classes [k0, k0 + kc) per launch.  Row-sharded runs: the
// all-gather of the log-densities is the kernel's epilogue -- a CTA stores the rows it computed (they are still in L2) into
// every peer's buffer with coalesced stores -- and (push_s >= 0: the launch of the last class chunk) the last CTA of the
// launch to finish publishes the epoch of iteration push_s (phase 1) that the sweep kernels of all ranks wait for.
// (Storing every value to the peers as it is computed -- 8-byte stores 24 bytes apart -- was measured at 92 us per launch on
// 8 GPUs against 53 us for this epilogue: NVLink wants full sectors.)
template <int KB, int KX, int RB, int P, int TPB>
__global__ void __launch_bounds__(TPB, 1) k_b_density_ring(const BatchInst* __restrict__ B, int G, int gate, int k0, int kc, int push_s) {
  const BatchInst& I = B[blockIdx.z];
  if (gate && !b_active(I)) return;
  const int W4 = I.Ws / 4;
  density_sk_ring_cta<KB, KX, RB, P, TPB>(reinterpret_cast<const uint4*>(I.bits), I.N_loc, W4, kc, G,
                                          reinterpret_cast<const uint4*>(I.P.m1) + (size_t)k0 * W4,
                                          reinterpret_cast<const uint4*>(I.P.valid) + (size_t)k0 * W4, I.P.L + k0, I.P.B + k0,
                                          I.P.has_half, I.logf_loc + k0, I.K, I.D);
  if (push_s < 0 || I.world <= 1) return;     // (the last class chunk's launch pushes whole rows: all classes)
  __syncthreads();
  {
    const long long rows_cta = (long long)(TPB / 32) * (32 / G) * RB;     // the contiguous row block of one pass of this CTA
    const long long stride = rows_cta * gridDim.x;
    for (long long base = (long long)blockIdx.x * rows_cta; base < I.N_loc; base += stride) {
      const long long first = (I.row_off + base) * I.K;
      const int cnt = (int)(min(rows_cta, I.N_loc - base) * I.K);
      const double* src = I.logf + first;
      for (int q = 1; q < I.push.n; q++) {
        double* dst = I.push.ptr[q] + first;
        for (int t = threadIdx.x; t < cnt; t += TPB) dst[t] = __ldcg(src + t);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(I.bcast_ticket, 1) == (int)gridDim.x - 1) {
      *I.bcast_ticket = 0;
      __threadfence_system();
      for (int r = 0; r < I.world; r++) {
        volatile int* dst = I.sig_ptrs[r] + 1 * kMaxPeers + I.rank;
        *dst = I.epoch0 + push_s;
      }
      __threadfence_system();
    }
  }
}

__global__ void __launch_bounds__(kOhTpb, 1) k_b_onehot_ring(const BatchInst* __restrict__ B, int tile_w, int slice_rows) {
  const BatchInst& I = B[blockIdx.z];
  if (!b_active(I)) return;
  mstep_onehot_ring_body(reinterpret_cast<const uint4*>(I.bits), I.Ws / 4, tile_w, I.K, I.offsets, I.perm, I.cnt, I.Dpad, slice_rows);
}

template <int KB, bool SMEM>
__global__ void __launch_bounds__(256) k_b_density_skd(const BatchInst* __restrict__ B, int G, int gate) {
  const BatchInst& I = B[blockIdx.z];
  if (gate && !b_active(I)) return;
  density_skd_body<KB, SMEM>(I.bits, I.N_loc, I.Ws, I.K, G, I.P.m1, I.P.valid, I.Ltab, I.Lword, I.P.B, I.logf_loc);
}

// row-sharded runs, start of a run: no rank may overwrite the others' log-densities (push) while they still read them for
// the criteria of the previous run -- every rank reports "my previous run is complete" and waits for all (phase 2)
__global__ void __launch_bounds__(32) k_b_run_barrier(const BatchInst* __restrict__ B) {
  const BatchInst& I = B[blockIdx.z];
  if (I.world <= 1) return;
  if (threadIdx.x < I.world) {
    __threadfence_system();
    volatile int* dst = I.sig_ptrs[threadIdx.x] + 2 * kMaxPeers + I.rank;
    *dst = I.epoch0;
    __threadfence_system();
  }
  peer_wait_lane(I.sig, I.world, 2, I.epoch0, I.peer_err);
}
// row-sharded runs, E-step: push this rank's fresh log-densities into every peer's buffer; the last CTA publishes the epoch
__global__ void __launch_bounds__(256) k_b_push_densities(const BatchInst* __restrict__ B, int s, int gate) {
  const BatchInst& I = B[blockIdx.z];
  if (I.world <= 1 || (gate && !b_active(I))) return;
  const int peer = blockIdx.y + 1;
  const long long first = I.row_off * I.K, count = I.N_loc * (long long)I.K;   // the shard, in doubles
  if (peer < I.push.n && count > 0) {
    const double* src = I.logf + first;
    double* dst = I.push.ptr[peer] + first;
    const long long head = (first & 1) ? 1 : 0;                                   // 16-byte stores need an even start
    if (head && blockIdx.x == 0 && threadIdx.x == 0) dst[0] = src[0];
    const double2* s2 = reinterpret_cast<const double2*>(src + head);
    double2* d2 = reinterpret_cast<double2*>(dst + head);
    const long long n2 = (count - head) / 2;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (long long)gridDim.x * blockDim.x) d2[t] = s2[t];
    if (((count - head) & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[count - 1] = src[count - 1];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = gridDim.x * gridDim.y;
    if (atomicAdd(I.bcast_ticket, 1) == total - 1) {
      *I.bcast_ticket = 0;
      __threadfence_system();
      for (int r = 0; r < I.world; r++) {
        volatile int* dst = I.sig_ptrs[r] + 1 * kMaxPeers + I.rank;
        *dst = I.epoch0 + s;
      }
      __threadfence_system();
    }
  }
}

// nem_alg.c:1868-1892 after the sweep of iteration s, for one instance: convergence / failure / iteration cap (one thread)
__device__ __forceinline__ int b_poll_decide(const BatchInst& I, int s, bool reset_barrier) {
  const float maxdif = __uint_as_float(I.state[kStMaxdiff]);
  const int status = (int)I.state[kStStatus];
  I.state[kStIter] = (unsigned)s;
  int still = 1;
  if (status != 0) still = 0;                                       // empty class: the reference stops (nem_alg.c:1873-1892)
  else {
    const int conv = maxdif < I.cv_th;                              // nem_alg.c:2160-2173
    I.state[kStConverged] = (unsigned)conv;
    if (conv || s >= I.itermax) still = 0;
  }
  I.state[kStMaxdiff] = 0u;
  if (reset_barrier) *I.barrier = 0u;
  if (I.spec_words != nullptr) I.spec_words[3 + ((s + 1) & 1)] = 0;
  if (!still) I.state[kStActive] = 0u;
  return still;
}
// the last instance of a group to report publishes "instances still active" into the host's mapped word (the ticket counter
// is the scratch word's high half)
__device__ __forceinline__ void b_poll_publish(int still, int n, volatile int* any_active, int* group_active) {
  __threadfence();
  const int prevv = atomicAdd(group_active, (still ? 1 : 0) + (1 << 16));
  if ((prevv >> 16) == n - 1) {
    const int total = (prevv & 0xffff) + (still ? 1 : 0);
    *group_active = 0;
    __threadfence_system();
    *any_active = total;
    __threadfence_system();
  }
}

// Tail of a cooperative sweep launch; replaces three launches of the next iteration: the statistics of the next M-step are
// cleared, [s >= 1: the convergence decision, k_b_poll's,] and the rows of the fresh posteriors are classified and scattered
// into the bucket lists (k_b_classify / k_b_scatter) -- separated by the instance's barrier instead of kernel boundaries.
// Not inlined: its registers must not weigh on the sweep's loops (the kernel runs 1024 threads per CTA, 64 registers each).
template <int GS>
__device__ __noinline__ void b_sweep_tail(const BatchInst* __restrict__ B, int s, volatile int* any_active, int* __restrict__ group_active) {
  const BatchInst& I = B[blockIdx.z];
  InstBar bar{I.barrier, gridDim.x};
  // the M-step statistics are not read by anything between the last finalize and the next count kernels
  {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long t = tid; t < I.zero_bytes / 4; t += nth) reinterpret_cast<int*>(I.zero_from)[t] = 0;
    if (tid == 0) *I.P.has_half = 0;
  }
  bar.sync();        // every row's posterior, the sweep's max |dT| and the cleared statistics are visible
  if (s >= 1 && blockIdx.x == 0 && threadIdx.x == 0) {
    const int still = b_poll_decide(I, s, false);     // (the barrier counter is monotonic across the launches of a run)
    b_poll_publish(still, (int)gridDim.z, any_active, group_active);
  }
  if (s == -1 && blockIdx.x == 0 && threadIdx.x == 0 && I.spec_words != nullptr) I.spec_words[4] = 0;   // (a speculative start sweep counted there)
  // classification of the fresh posteriors for the next M-step (an instance that just stopped does it for nothing);
  // buffer read by the M-step of iteration s + 1: (c0 + s) & 1, after the start sweeps: c0
  const int cur = s >= 1 ? ((I.c0 + s) & 1) : I.c0;
  classify_body<GS>(I.T[cur] + I.row_off * I.K, I.N_loc, I.K, I.prev, I.ent, I.hist, I.nk_cnt, I.nk_fz, I.full_recount);
  bar.sync();
  scatter_body(I.ent, I.N_loc, I.hist, 2 * I.K + 1, I.offsets, I.cursor, I.perm);
}

// s >= 1: the sweep of EM iteration s (reads the posteriors of iteration s - 1); s == 0 / -1: the two start sweeps of
// nem_alg.c:2043-2054 (blind beta = 0 from the zero posteriors in T[0] into T[1]; then with beta from T[1] into T[0]).
template <bool COOP, int GS>
__global__ void __launch_bounds__(COOP ? 1024 : 256) k_b_sweep(const BatchInst* __restrict__ B, int s, int tail,
                                                                 volatile int* any_active, int* __restrict__ group_active) {
  const BatchInst& I = B[blockIdx.z];
  if (s >= 1 && !b_active(I)) {      // a finished instance still reports to the group's "any active" word
    if (COOP && tail && blockIdx.x == 0 && threadIdx.x == 0) b_poll_publish(0, (int)gridDim.z, any_active, group_active);
    return;
  }
  const bool skip = s == -1 && I.beta == 0.0f;   // an instance without a graph (in a group with one): no second start sweep
  if (skip && !(COOP && tail)) return;
  int from, to;
  float beta = I.beta;
  if (s >= 1) { from = (I.c0 + s - 1) & 1; to = from ^ 1; }
  else if (s == 0) { from = 0; to = 1; beta = 0.0f; }
  else { from = 1; to = 0; }
  SweepArgs a;
  a.N = I.N; a.K = I.K; a.logf = I.logf; a.pk = I.P.pk; a.logpk64 = I.P.logpk64; a.T_old = I.T[from]; a.T_new = I.T[to];
  a.rowptr = I.rowptr; a.col = I.col; a.w = I.w; a.beta = beta; a.logspace = I.logspace; a.jacobi = I.jacobi;
  a.maxdiff_bits = I.state + (s >= 1 ? kStMaxdiff : kStWords - 1);   // the start sweeps are not convergence-tested
  a.ell_hdr = I.ell_hdr; a.ell_nb = I.ell_nb; a.stamps = nullptr;
  // a deep schedule (a family order that follows the genomes: hundreds of thin levels, up to one row per level) is swept
  // speculatively whatever the previous sweep changed, the second start sweep included
  const bool deep = I.spec_deep > 0 && I.n_levels >= I.spec_deep;
  a.chg = (s >= 1 || (deep && s == -1)) ? I.chg : nullptr; a.spec_words = I.spec_words;
  a.spec_limit = deep ? 0x7fffffff : (int)(I.N / 32); a.spec_parity = s & 1;
  a.wait_sig = nullptr; a.wait_n = 0; a.wait_epoch = 0; a.wait_err = nullptr;
  if (I.world > 1 && s >= 0) {        // the densities of this sweep's epoch, pushed by every rank (the start sweep with beta
    a.wait_sig = I.sig; a.wait_n = I.world; a.wait_epoch = I.epoch0 + s; a.wait_err = I.peer_err;   // reuses those of s = 0)
  }
  const bool levels = COOP && beta != 0.0f && I.rowptr != nullptr && !I.jacobi && I.n_levels > 1;
  a.level_ptr = levels ? I.level_ptr : nullptr; a.level_rows = levels ? I.level_rows : nullptr; a.n_levels = levels ? I.n_levels : 1;
  InstBar bar{I.barrier, gridDim.x};
  if (!skip) sweep_body<COOP, GS>(a, bar);
  if (COOP && tail) b_sweep_tail<GS>(B, s, any_active, group_active);
}

// nem_alg.c:1868-1892 on the device, after the sweep of iteration s: convergence / failure / iteration cap per instance,
// the statistics of the next M-step cleared, and "is any instance of the group still running" for the host
__global__ void __launch_bounds__(256) k_b_poll(const BatchInst* __restrict__ B, int s, volatile int* any_active /*[1] mapped host word*/,
                                                int* __restrict__ group_active /*[1] device scratch*/) {
  const BatchInst& I = B[blockIdx.z];
  const bool was = b_active(I);
  if (was) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long t = tid; t < I.zero_bytes / 4; t += nth) reinterpret_cast<int*>(I.zero_from)[t] = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) *I.P.has_half = 0;
  }
  __syncthreads();
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int still = was ? b_poll_decide(I, s, true) : 0;
  b_poll_publish(still, (int)gridDim.z, any_active, group_active);
}

// criteria of all instances of a group (nem_alg.c:1906): per-row terms, then the float32 sums in the reference's order
template <int GS>
__global__ void __launch_bounds__(256) k_b_criteria(const BatchInst* __restrict__ B) {
  const BatchInst& I = B[blockIdx.z];
  if (I.state[kStStatus] != 0u) return;
  CritArgs a;
  a.N = I.N; a.K = I.K; a.logf = I.logf; a.pk = I.P.pk; a.logpk32 = I.P.logpk32; a.logpk64 = I.P.logpk64;
  a.T = I.T[b_final_buffer(I)];
  a.rowptr = I.rowptr; a.col = I.col; a.w = I.w; a.beta = I.beta; a.logspace = I.logspace;
  a.partial = I.crit_partial;
  const int n_warps = gridDim.x * 8;
  a.rows_per_warp = (I.N + n_warps - 1) / n_warps;
  a.terms = I.logspace ? nullptr : reinterpret_cast<float2*>(I.crit_terms);
  a.counts = I.logspace ? nullptr : I.crit_counts;
  a.row_lz = I.logspace ? nullptr : reinterpret_cast<double2*>(I.crit_lz);
  criteria_body<GS>(a);
}
// reference arithmetic: region offsets (one warp per instance) ...
__global__ void __launch_bounds__(32) k_b_crit_offsets(const BatchInst* __restrict__ B, int n_regions) {
  const BatchInst& I = B[blockIdx.z];
  if (I.state[kStStatus] != 0u || I.logspace) return;
  crit_offsets_body(I.crit_counts, n_regions, I.crit_offsets, I.crit_nrows, (int)I.N);
}
// ... and the compaction of the regions' terms
__global__ void __launch_bounds__(256) k_b_crit_compact(const BatchInst* __restrict__ B, int n_regions) {
  const BatchInst& I = B[blockIdx.z];
  if (I.state[kStStatus] != 0u || I.logspace) return;
  crit_compact_body(reinterpret_cast<const float2*>(I.crit_terms), I.crit_counts, I.crit_offsets, n_regions,
                    (I.N + n_regions - 1) / n_regions, I.K, reinterpret_cast<float2*>(I.crit_dense));
}
// log-space mode: fp64 sums of the block partials, in block order
__global__ void __launch_bounds__(128) k_b_crit_reduce(const BatchInst* __restrict__ B, int nblocks) {
  const BatchInst& I = B[blockIdx.z];
  if (I.state[kStStatus] != 0u || !I.logspace) return;
  const int w = threadIdx.x >> 5;
  const double t = ordered_sum(I.crit_partial + w, nblocks, 4);
  if ((threadIdx.x & 31) == 0) I.crit4[w] = t;
}

// results of all instances of a group in one launch: final posteriors, parameters, and the scalar outcome
// (out_scalars[inst] = {n_iter, converged, status, c_final})
__global__ void __launch_bounds__(256) k_b_collect(const BatchInst* __restrict__ B, double* __restrict__ out_scalars /*[n][8]*/) {
  const BatchInst& I = B[blockIdx.z];
  const int n_iter = (int)I.state[kStIter], status = (int)I.state[kStStatus];
  const int cfin = b_final_buffer(I);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (tid == 0) {
    double* o = out_scalars + (size_t)blockIdx.z * 8;
    o[0] = n_iter; o[1] = (double)I.state[kStConverged]; o[2] = status; o[3] = cfin;
    for (int q = 0; q < 4; q++) o[4 + q] = status == 0 ? I.crit4[q] : 0.0;
  }
  if (I.T_out != nullptr) {
    const float* src = I.T[cfin];
    const long long n = I.N * I.K;
    for (long long t = tid; t < n; t += nth) I.T_out[t] = src[t];
  }
  const long long kd = (long long)I.K * I.D;
  if (I.mu_out != nullptr) for (long long t = tid; t < kd; t += nth) I.mu_out[t] = I.P.mu[t];
  if (I.eps_out != nullptr)
    for (long long t = tid; t < kd; t += nth) I.eps_out[t] = I.free_disp ? I.P.eps[t] : I.P.eps_k[t / I.D];
  if (I.pk_out != nullptr && tid < I.K) I.pk_out[tid] = I.P.pk[tid];
}

}  // namespace nemb200
