// nemb200_export.cuh -- device-side export of NEM instances from the array form of a pangenome.
//
// What it restates: write_nem_input_files (/root/reference/ppanggolin/nem/partition.py:344-436) -- for a subset of genomes
// the presence rows of the families that occur in it (:381-391), the neighbour lists with weights
// round(gene pairs in the subset / #genomes, 4) (:393-414), the max-degree cut (:415-433), sum of kept weights / 2 (:416,
// :436) -- plus what the C reader then keeps of the .nei file (NEM/nem_exe.c:1395-1451: the first nv indices and the nv
// non-zero weights), and the level schedule of the in-place sweep (nem_alg.c:2441-2476, see nemb200_gs_schedule).
// The reference does this in Python for every chunk sample (~1.2 us per matrix cell); here a batch of samples is derived
// in a handful of launches (blockIdx.y = sample) from arrays that stay resident in HBM.
#pragma once
#include "nemb200_kernels.cuh"

namespace nemb200 {

struct ExpPan {               // the whole pangenome (device pointers)
  long long F, G, E, A;
  int Wg;                     // words per presence row
  const uint32_t* presence;   // [F][Wg]
  const int* adj_ptr;         // [F+1]
  const int* adj_other;       // [A]
  const int* adj_edge;        // [A]
  const int* cov_ptr;         // [E+1] or null: co-presence model
  const int* cov_org;         // [M]
  const int* cov_cnt;         // [M]
  const int* edge_fam;        // [E][2]
};

struct ExpSample {            // one genome subset and its outputs (device pointers; capacity F rows / A entries)
  const int* sel;             // [n_sel] genome columns, ascending; null = all genomes in order (n_sel = G, W = Wg)
  int n_sel, W;               // W: words per instance row
  uint32_t* mask;             // [Wg] scratch: bit set for every selected genome
  uint32_t* tmp;              // [F][W] scratch: gathered rows of all families
  int* present;               // [F] scratch -> rowmap (exclusive scan)
  int* cov;                   // [E] scratch: gene pairs of every edge in the subset
  int* nv_fam;                // [F] scratch: kept neighbours per family
  double* ew_part;            // [ew_blocks] scratch
  uint32_t* bits;             // [N][W] out
  int* fam_rows;              // [N] out: family of every instance row
  int* rowptr;                // [N+1] out
  int* col;                   // [nnz] out
  float* w;                   // [nnz] out
  int* level;                 // [N] scratch
  int* level_ptr;             // [N+1] out (n_levels + 1 used)
  int* level_rows;            // [N] out
  int* level_cur;             // [N+1] scratch
  unsigned* barrier;          // [1] scratch
  int* changed;               // [2] scratch
  int* scalars;               // [4] out: N, nnz, n_levels, -
  double* edges_weight;       // [1] out
};

// ---- block-wide exclusive scan helper: one CTA of 1024 threads walks n items, 4 per thread and step ----------------
__device__ __forceinline__ int block_scan_excl(const int* in, int* out /*may alias in*/, long long n) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  constexpr int IT = 4;
  for (long long base = 0; base < n; base += (long long)blockDim.x * IT) {
    const long long i0 = base + (long long)threadIdx.x * IT;
    int v[IT], tot = 0;
#pragma unroll
    for (int q = 0; q < IT; q++) { v[q] = (i0 + q < n) ? in[i0 + q] : 0; tot += v[q]; }
    int incl = tot;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int wv = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
      int wi = wv;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, off); if (lane >= off) wi += t; }
      s_warp[lane] = wi - wv;    // exclusive warp offsets
      if (lane == 31) s_warp[31] = wi - wv;
    }
    __syncthreads();
    int acc = s_carry + s_warp[wid] + incl - tot;
#pragma unroll
    for (int q = 0; q < IT; q++) { if (i0 + q < n) out[i0 + q] = acc; acc += v[q]; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = acc;
    __syncthreads();
  }
  return s_carry;
}

// selected-genome bit mask (explicit coverage tables look genomes up in it)
__global__ void __launch_bounds__(256) k_exp_mask(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.Wg; t += gridDim.x * blockDim.x) s.mask[t] = 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) *s.barrier = 0u;
}
__global__ void __launch_bounds__(256) k_exp_mask_set(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < s.n_sel; t += gridDim.x * blockDim.x) {
    const int g = s.sel ? s.sel[t] : t;
    atomicOr(&s.mask[g >> 5], 1u << (g & 31));
  }
}

// column gather: row f of the instance matrix = the selected genome bits of presence row f, in `sel` order
// (partition.py:381-389); one warp per family, lane = output bit
__global__ void __launch_bounds__(256) k_exp_gather(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  if (s.sel == nullptr) {   // all genomes in column order: the rows are the presence rows (s.tmp aliases them)
    for (long long f = warp; f < P.F; f += nwarps) {
      unsigned any = 0u;
      for (int w = lane; w < P.Wg; w += 32) any |= P.presence[f * P.Wg + w];
      any = __reduce_or_sync(0xffffffffu, any);
      if (lane == 0) s.present[f] = any != 0u;
    }
    return;
  }
  for (long long f = warp; f < P.F; f += nwarps) {
    const uint32_t* row = P.presence + f * P.Wg;
    unsigned any = 0u;
    for (int w = 0; w < s.W; w++) {
      const int c = w * 32 + lane;
      int bit = 0;
      if (c < s.n_sel) { const int g = s.sel[c]; bit = (row[g >> 5] >> (g & 31)) & 1u; }
      const unsigned word = __ballot_sync(0xffffffffu, bit);
      any |= word;
      if (lane == 0) s.tmp[f * s.W + w] = word;
    }
    if (lane == 0) s.present[f] = any != 0u;
  }
}

// present -> rowmap (in place), N
__global__ void __launch_bounds__(1024) k_exp_scan_rows(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  // the flags are needed again after the scan: keep them in nv_fam's sign?  No: rowmap[f + 1] - rowmap[f] recovers them.
  const int total = block_scan_excl(s.present, s.present, P.F);
  if (threadIdx.x == 0) s.scalars[0] = total;
}

__device__ __forceinline__ bool exp_is_present(const ExpSample& s, long long f, long long F, int N) {
  const int a = s.present[f];
  const int b = (f + 1 < F) ? s.present[f + 1] : N;
  return b > a;
}

__global__ void __launch_bounds__(256) k_exp_compact(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  const int N = s.scalars[0];
  const long long n = P.F * s.W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long f = t / s.W;
    const int w = (int)(t - f * s.W);
    if (exp_is_present(s, f, P.F, N)) {
      const int r = s.present[f];
      s.bits[(long long)r * s.W + w] = s.tmp[t];
      if (w == 0) s.fam_rows[r] = (int)f;
    }
  }
}

// gene pairs of every edge inside the subset (partition.py:399-407)
__global__ void __launch_bounds__(256) k_exp_cov(ExpPan P, const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  if (P.cov_ptr == nullptr) {   // co-presence model: one pair in genome g iff both families are present in g
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < P.E; e += (long long)gridDim.x * blockDim.x) {
      const uint32_t* ra = s.tmp + (long long)P.edge_fam[2 * e] * s.W;
      const uint32_t* rb = s.tmp + (long long)P.edge_fam[2 * e + 1] * s.W;
      int c = 0;
      for (int w = 0; w < s.W; w++) c += __popc(ra[w] & rb[w]);
      s.cov[e] = c;
    }
  } else {                      // explicit (edge, genome, pairs) table: one warp per edge
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long e = warp; e < P.E; e += nwarps) {
      int c = 0;
      for (int m = P.cov_ptr[e] + lane; m < P.cov_ptr[e + 1]; m += 32) {
        const int g = P.cov_org[m];
        if ((s.mask[g >> 5] >> (g & 31)) & 1u) c += P.cov_cnt[m];
      }
      for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
      if (lane == 0) s.cov[e] = c;
    }
  }
}

// per family: neighbours with coverage, max-degree cut, entries the C reader keeps, contribution to the total weight
constexpr int kExpEwBlocks = 256;
__global__ void __launch_bounds__(256) k_exp_degrees(ExpPan P, const ExpSample* __restrict__ S, float sm_degree,
                                                     const float* __restrict__ lut, int lut_len) {
  __shared__ double s_red[8];
  const ExpSample& s = S[blockIdx.y];
  const int N = s.scalars[0];
  double ew = 0.0;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < P.F; f += (long long)gridDim.x * blockDim.x) {
    int nv = 0;
    if (exp_is_present(s, f, P.F, N)) {
      int n_nei = 0, nz = 0;
      double dist = 0.0;
      for (int a = P.adj_ptr[f]; a < P.adj_ptr[f + 1]; a++) {
        const int c = s.cov[P.adj_edge[a]];
        if (c > 0) {
          n_nei++;
          dist += (double)c / (double)s.n_sel;               // partition.py:408
          const float wv = c < lut_len ? lut[c] : (float)((double)c / (double)s.n_sel);
          if (wv != 0.0f) nz++;
        }
      }
      if (n_nei > 0 && (float)n_nei < sm_degree) { nv = nz; ew += dist; }   // partition.py:415-416
    }
    s.nv_fam[f] = nv;
  }
  for (int off = 16; off > 0; off >>= 1) ew += __shfl_xor_sync(0xffffffffu, ew, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ew;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; q++) t += s_red[q];
    s.ew_part[blockIdx.x] = t;
  }
}

// nv per instance row -> rowptr; total -> nnz; sum of the block partials (fixed order) / 2 -> edges_weight
__global__ void __launch_bounds__(1024) k_exp_scan_nnz(ExpPan P, const ExpSample* __restrict__ S, int ew_blocks) {
  const ExpSample& s = S[blockIdx.y];
  const int N = s.scalars[0];
  // nv of the instance rows, in row order: gather through fam_rows into level (scratch), then scan into rowptr
  for (int r = threadIdx.x; r < N; r += blockDim.x) s.level[r] = s.nv_fam[s.fam_rows[r]];
  __syncthreads();
  const int total = block_scan_excl(s.level, s.rowptr, N);
  if (threadIdx.x == 0) {
    s.rowptr[N] = total;
    s.scalars[1] = total;
    double t = 0.0;
    for (int b = 0; b < ew_blocks; b++) t += s.ew_part[b];
    *s.edges_weight = t / 2.0;
  }
}

// neighbour lists: the first nv kept indices, and the nv non-zero weights in order (nem_exe.c:1395-1451)
__global__ void __launch_bounds__(256) k_exp_fill(ExpPan P, const ExpSample* __restrict__ S, const float* __restrict__ lut, int lut_len) {
  const ExpSample& s = S[blockIdx.y];
  const int N = s.scalars[0];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
    const int f = s.fam_rows[r];
    const int nv = s.nv_fam[f];
    if (nv == 0) continue;
    const int base = s.rowptr[r];
    int ti = 0, tw = 0;
    for (int a = P.adj_ptr[f]; a < P.adj_ptr[f + 1]; a++) {
      const int c = s.cov[P.adj_edge[a]];
      if (c <= 0) continue;
      if (ti < nv) s.col[base + ti] = s.present[P.adj_other[a]];
      ti++;
      const float wv = c < lut_len ? lut[c] : (float)((double)c / (double)s.n_sel);
      if (wv != 0.0f) { s.w[base + tw] = wv; tw++; }
    }
  }
}

// level(i) = 1 + max level(j) over neighbours j < i: relaxation passes until nothing changes (depth of the schedule + 1
// passes), all CTAs of a sample between instance barriers; then a counting sort of the rows by level
__global__ void __launch_bounds__(256) k_exp_levels(const ExpSample* __restrict__ S) {
  const ExpSample& s = S[blockIdx.y];
  const int N = s.scalars[0];
  InstBar bar{s.barrier, gridDim.x};
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < N; i += nth) s.level[i] = 0;
  for (int i = tid; i <= N; i += nth) s.level_cur[i] = 0;
  if (tid == 0) { s.changed[0] = 0; s.changed[1] = 0; }
  bar.sync();
  int pass = 0;
  while (true) {
    int any = 0;
    for (int i = tid; i < N; i += nth) {
      int lv = 0;
      for (int n = s.rowptr[i]; n < s.rowptr[i + 1]; n++) {
        const int j = s.col[n];
        if (j < i) lv = max(lv, __ldcg(&s.level[j]) + 1);
      }
      if (lv != s.level[i]) { __stcg(&s.level[i], lv); any = 1; }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicExch(&s.changed[pass & 1], 1);
    bar.sync();
    const int ch = __ldcg(&s.changed[pass & 1]);
    if (tid == 0) s.changed[(pass + 1) & 1] = 0;
    if (!ch) break;
    pass++;
    if (pass > N + 1) break;
    bar.sync();   // the reset of the other flag is ordered before its next use
  }
  // histogram of the levels (level_cur[l + 1] counts), scan by CTA 0, scatter
  for (int i = tid; i < N; i += nth) atomicAdd(&s.level_cur[s.level[i] + 1], 1);
  bar.sync();
  if (blockIdx.x == 0) {
    // sequential-in-chunks scan by one warp is enough: the number of levels is small
    __shared__ int s_nl;
    if (threadIdx.x == 0) {
      int acc = 0, nl = 0;
      for (int l = 0; l <= N; l++) {
        const int c = s.level_cur[l];
        acc += c;
        s.level_ptr[l] = acc;
        s.level_cur[l] = acc;        // running cursor of level l = start of level l
        if (l >= 1 && c > 0) nl = l;
        if (acc == N && l >= 1) { nl = l; break; }
      }
      s_nl = nl;
      s.scalars[2] = nl;
    }
    __syncthreads();
  }
  bar.sync();
  for (int i = tid; i < N; i += nth) {
    const int l = s.level[i];
    const int pos = atomicAdd(&s.level_cur[l], 1);
    s.level_rows[pos] = i;
  }
}

// Chunk votes (partition.py:784-802, validate_family).  Step 1, per finished chunk: the label rule of partition.py:206-231
// on the posteriors, reduced to the partition letter that is voted with (P / S / C), scattered to a vector over ALL families
// (-1 = family not in the chunk).  Step 2, per round: every family walks the chunks of the round in sample order and applies
// the reference's rule vote by vote -- a family occurs at most once per chunk, so this is the reference's sequential loop.
__global__ void __launch_bounds__(256)
k_chunk_codes(const float* __restrict__ T, int N, int K, const int* __restrict__ fam_rows, signed char* __restrict__ codes /*[F], -1 filled*/) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  double mx = -1.0;
  int arg = -1, nmax = 0;
  for (int k = 0; k < K; k++) {
    const double v = rint((double)T[(size_t)r * K + k] * 1000.0) / 1000.0;
    if (v > mx) { mx = v; arg = k; nmax = 1; }
    else if (v == mx) nmax++;
  }
  const int lab = (nmax > 1 || mx < 0.5) ? -1 : arg;
  codes[fam_rows[r]] = (signed char)(lab == 0 ? 0 : (lab == K - 1 ? 2 : 1));     // first letter: P / S (S1.., S_) / C
}

__global__ void __launch_bounds__(256)
k_apply_votes(const signed char* __restrict__ codes /*[n_chunks][F]*/, int n_chunks, long long F, int* __restrict__ votes /*[F][4] P S C U*/,
              unsigned char* __restrict__ validated, int* __restrict__ n_validated, double ratio /*n_org / chunk_size*/, int n_org) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int v0 = votes[f * 4], v1 = votes[f * 4 + 1], v2 = votes[f * 4 + 2], v3 = votes[f * 4 + 3];
  bool val = validated[f] != 0, newly = false;
  for (int c = 0; c < n_chunks; c++) {
    const int code = codes[(size_t)c * F + f];
    if (code < 0) continue;
    if (code == 0) v0++; else if (code == 1) v1++; else v2++;
    const int sum = v0 + v1 + v2 + v3;
    const int m = max(max(v0, v1), max(v2, v3));
    if ((((double)sum > ratio) && ((double)m >= (double)sum * 0.5)) || sum > n_org) {
      if (!val) {
        if ((double)m < (double)sum * 0.5) v3 = n_org;
        val = true; newly = true;
      }
    }
  }
  votes[f * 4] = v0; votes[f * 4 + 1] = v1; votes[f * 4 + 2] = v2; votes[f * 4 + 3] = v3;
  if (newly) { validated[f] = 1; atomicAdd(n_validated, 1); }
}

__global__ void __launch_bounds__(256) k_set_mask_bits(const int* __restrict__ sel, int n, uint32_t* __restrict__ mask) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) atomicOr(&mask[sel[t] >> 5], 1u << (sel[t] & 31));
}

// popcount statistics of a genome subset for the rarefaction curves (rarefaction.py:657-684 counts them with gmpy2
// bitarrays): per family the number of selected genomes that hold it
__global__ void __launch_bounds__(256) k_subset_counts(ExpPan P, const uint32_t* __restrict__ mask /*[Wg]*/, int* __restrict__ counts /*[F]*/) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long f = warp; f < P.F; f += nwarps) {
    int c = 0;
    for (int w = lane; w < P.Wg; w += 32) c += __popc(P.presence[f * P.Wg + w] & mask[w]);
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) counts[f] = c;
  }
}

}  // namespace nemb200
