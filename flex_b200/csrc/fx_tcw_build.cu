// fx_tcw_build.cu -- builder of the tensor-window format (FX_FMT_TCW): the B200-native tile format.
//
// Flex's thesis (mat.cu:1345-1518 tiles, :718-1050 diagonal tiling) is that after reordering many
// rows of a panel share columns, and that those shared columns should be served from on-chip memory
// once per panel.  On a B200 the on-chip consumer of choice is the tensor core, so the format is:
//   per 128-row panel p : S_p = the (at most W) columns with the most nz in the panel, each with
//                         at least T of them, ascending                      -> tc_cols / tc_ncol
//   window part         : the nz whose column is in S_p, grouped by (panel, 32-column chunk of S_p) and
//                         ordered by (row, slot) inside a chunk, as (word of (row, slot&31) in the K-major A tile, value):
//                         one chunk is one K=32 step of the tensor kernel, which scatters its entries with
//                         all threads at once                                -> win_cptr/code/val
//   remainder           : every other nz as an ordinary CSR                  -> rest_rowptr/col/val
// The remainder then goes through the ASpT builder unchanged (fx_aspt_build.cu), the window part is
// multiplied by k_spmm_tc (fx_tc_kernel.cuh).  oracle/tcw.py restates the selection rule
// on the CPU; tests compare every array bit for bit.
//
// Selection rule (deterministic, all integer):  cnt_p[c] = nz of panel p in column c;  candidates =
// {c : cnt >= T'} where T' = T, doubled until at most CAND_CAP candidates remain (at most 6 doublings,
// else the panel gets no window);  order candidates by (cnt descending, column ascending) and cut them
// into chunks of 32 (one K=32 step of the tensor kernel), at most W/32 of them;  S_j = sum over chunk j
// of (cnt - 1) is the number of B-row fetches the chunk saves the gather path;  keep the leading chunks
// with S_j >= CHUNK_COST (what staging + 12 MMAs cost, in the same unit);  the panel keeps its window
// only if  G_p = sum(S_j - CHUNK_COST) >= MIN_GAIN  (the tc_out round trip of the panel);  the whole
// matrix keeps its windows only if  sum_p G_p >= MIN_TOTAL  (an extra kernel launch and its tail).
// Kept columns are listed ascending.  Costs measured on a B200 at k = 128: DESIGN.md section 5.
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>

#include <algorithm>

#include "fx_common.cuh"
#include "fx_scan.cuh"
#include "fx_tc_kernel.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int BH = 128;
constexpr int CAND_CAP = 8192;
// A column counter of k_tcw_select: bits 0-7 the count (at most 128: columns are unique within a row), bits 8-13 the
// "registered in doubling round r" flags, bits 14-30 the EPOCH of the panel that last touched it (a counter of an older
// epoch reads as zero, so the kernel never walks the panel a third time to zero what it touched), bit 31 = MARK | slot.
constexpr unsigned CNT_MASK = 0xFFu;
constexpr int FLAG_SHIFT = 8, EPOCH_SHIFT = 14;
constexpr unsigned EPOCH_MAX = (1u << (31 - EPOCH_SHIFT)) - 1;
constexpr unsigned MARK = 0x80000000u;
constexpr int MAX_CH = 128;  // W <= 4096

__device__ __forceinline__ void cmpswap64(unsigned long long& a, unsigned long long& b) {
  if (a > b) { unsigned long long t = a; a = b; b = t; }
}

// normalised bitonic network over shared memory (every comparator moves the minimum down, so slots
// >= n act as +inf padding and any n sorts correctly)
__device__ void tcw_sort_u64(unsigned long long* a, int n) {
  int N = 1;
  while (N < n) N <<= 1;
  for (int k = 2; k <= N; k <<= 1) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      int j = i ^ (k - 1);
      if (j > i && j < n) cmpswap64(a[i], a[j]);
    }
    __syncthreads();
    for (int s = k >> 2; s > 0; s >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int j = i ^ s;
        if (j > i && j < n) cmpswap64(a[i], a[j]);
      }
      __syncthreads();
    }
  }
}

// The listed columns of ONE panel (at most W) in a small open-addressing table in shared memory: key = column, slot = its
// position in the ascending list.  Pass 2 of k_tcw_select and both passes of k_tcw_split ask "is this nz's column listed, and
// where" once per nz; through this table the question costs one or two shared-memory reads instead of a dependent global
// load from the CTA's n-sized counter array (k_tcw_split 0.95 -> 0.xx ms on Reddit-shape).  HT = power of two >= 2 W.
constexpr unsigned HEMPTY = 0xFFFFFFFFu;
__host__ __device__ inline int tcw_hash_size(int W) { int h = 64; while (h < 2 * W) h <<= 1; return h; }
__device__ __forceinline__ unsigned tcw_hash(unsigned c, int hbits) { return (c * 2654435761u) >> (32 - hbits); }
__device__ __forceinline__ void tcw_hash_clear(unsigned* hk, int HT) {
  for (int i = threadIdx.x; i < HT; i += blockDim.x) hk[i] = HEMPTY;
}
__device__ __forceinline__ void tcw_hash_insert(unsigned* hk, unsigned short* hs, int HT, int hbits, unsigned c, int slot) {
  unsigned h = tcw_hash(c, hbits);
  while (atomicCAS(&hk[h], HEMPTY, c) != HEMPTY) h = (h + 1) & (unsigned)(HT - 1);
  hs[h] = (unsigned short)slot;
}
// MARK | slot when the column is listed, 0 otherwise
__device__ __forceinline__ unsigned tcw_hash_find(const unsigned* hk, const unsigned short* hs, int HT, int hbits, unsigned c) {
  unsigned h = tcw_hash(c, hbits);
  for (;;) {
    const unsigned k = hk[h];
    if (k == c) return MARK | hs[h];
    if (k == HEMPTY) return 0u;
    h = (h + 1) & (unsigned)(HT - 1);
  }
}

// Pass 1 of k_tcw_select counts a panel's columns.  A panel of Reddit-shape has ~13 k nz in ~5 k distinct columns: they fit an
// open-addressing table in shared memory (HT2 slots: key = column, value = the same count | flag word the global counters
// hold), so the count costs two shared-memory atomics per nz instead of two L2 atomics into the CTA's n-sized array.  A panel
// whose distinct columns pass 3/4 of the table (hub panels of hub-first orderings) falls back to the global counters.
// (Measured: k_tcw_select 0.63 -> 0.61 ms on Reddit-shape, tPre of yelp-shape 1.43 -> 1.21 ms.  The counting pass is not what
// the kernel waits for any more: its stall samples are 28 % barrier + 22 % fixed-latency waits, the ~70 block barriers per
// panel of the two bitonic sorts -- a partial sort of the 256 best candidates is the next step.)
constexpr int HT2_BITS = 14, HT2 = 1 << HT2_BITS;
__device__ __forceinline__ int tcw_h2_upsert(unsigned* hk, unsigned c, int* distinct) {
  unsigned h = (c * 2654435761u) >> (32 - HT2_BITS);
  for (int probe = 0; probe < 96; ++probe) {
    const unsigned k = *reinterpret_cast<volatile unsigned*>(hk + h);
    if (k == c) return (int)h;
    if (k == HEMPTY) {
      const unsigned prev = atomicCAS(&hk[h], HEMPTY, c);
      if (prev == HEMPTY) { atomicAdd(distinct, 1); return (int)h; }
      if (prev == c) return (int)h;
    }
    h = (h + 1) & (unsigned)(HT2 - 1);
  }
  return -1;
}
__device__ __forceinline__ int tcw_h2_find(const unsigned* hk, unsigned c) {  // the column is in the table
  unsigned h = (c * 2654435761u) >> (32 - HT2_BITS);
  while (hk[h] != c) h = (h + 1) & (unsigned)(HT2 - 1);
  return (int)h;
}

__global__ void k_tcw_pad(const uint32_t* __restrict__ rowptr, int row0, int nloc, int nr, int ne, int* __restrict__ csr_v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nr) csr_v[i] = i <= nloc ? (int)(rowptr[row0 + i] - rowptr[row0]) : ne;
}

// stats: [3] panels dropped by the candidate cap, [5] total net gain (the rest is filled by k_tcw_panels)
__global__ void __launch_bounds__(1024) k_tcw_select(const int* __restrict__ csr_v, const uint32_t* __restrict__ col, int npanel,
                                                    int ncols, int T, int W, int min_gain, int chunk_cost,
                                                    unsigned* __restrict__ cnt_all,
                                                    int* __restrict__ tc_cols, int* __restrict__ tc_ncol,
                                                    int* __restrict__ win_len, int* __restrict__ chunk_len,
                                                    unsigned long long* __restrict__ stats, int use_h2) {
  extern __shared__ unsigned long long skeys[];  // CAND_CAP, then the listed-column table (HT keys, HT slots)
  const int HT = tcw_hash_size(W), hbits = 31 - __clz(HT);
  unsigned* hk = reinterpret_cast<unsigned*>(skeys + CAND_CAP);
  unsigned* h2k = hk + HT;        // [HT2] pass-1 table: columns
  unsigned* h2v = h2k + HT2;      // [HT2] their count | flag words
  unsigned short* hs = reinterpret_cast<unsigned short*>(use_h2 ? h2v + HT2 : hk + HT);  // (no pass-1 table without use_h2)
  __shared__ int s_nc, s_distinct, s_over;
  __shared__ int hist[MAX_CH];
  __shared__ int s_ns;
  __shared__ long long s_gain;
  __shared__ int s_rp[BH + 1], s_wl[BH];
  const int CH = W / 32;
  unsigned* cnt = cnt_all + (size_t)blockIdx.x * ncols;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  __shared__ int s_panel;
  unsigned epoch = 0;
  // Panels are handed out by a counter (stats[6]): a panel that holds hub rows takes several times the average, and a fixed
  // stride left the kernel waiting for the CTA that drew the most of them.  Every panel's output depends on the panel alone.
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_panel = (int)atomicAdd(&stats[6], 1ull);
    __syncthreads();
    const int p = s_panel;
    if (p >= npanel) break;
    const int lb = csr_v[p * BH], ub = csr_v[(p + 1) * BH];
    if (threadIdx.x == 0) s_nc = 0;
    if (++epoch > EPOCH_MAX) {  // the epoch field is full: zero this CTA's counters and start over
      for (int i = threadIdx.x; i < ncols; i += blockDim.x) cnt[i] = 0u;
      epoch = 1;
    }
    const unsigned ebase = epoch << EPOCH_SHIFT;
    __syncthreads();
    // pass 1, in shared memory where the panel's distinct columns fit the table
    bool in_smem = use_h2 && ub - lb <= 4 * HT2;
    if (in_smem) {
      for (int i = threadIdx.x; i < HT2; i += blockDim.x) { h2k[i] = HEMPTY; h2v[i] = 0u; }
      if (threadIdx.x == 0) { s_distinct = 0; s_over = 0; }
      __syncthreads();
      for (int e0 = lb + threadIdx.x; e0 < ub; e0 += 4 * blockDim.x) {
        unsigned c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = e0 + j * (int)blockDim.x < ub ? col[e0 + j * (int)blockDim.x] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c[j] != 0xFFFFFFFFu) {
            const int slot = tcw_h2_upsert(h2k, c[j], &s_distinct);
            if (slot < 0) { s_over = 1; continue; }
            const unsigned old = atomicAdd(&h2v[slot], 1u) & CNT_MASK;
            if ((int)old == T - 1) {
              const int pos = atomicAdd(&s_nc, 1);
              if (pos < CAND_CAP) skeys[pos] = c[j];
            }
          }
        if (s_distinct > HT2 * 3 / 4) s_over = 1;
        if (s_over) break;
      }
      __syncthreads();
      if (s_over) {  // start over on the global counters
        in_smem = false;
        __syncthreads();
        if (threadIdx.x == 0) s_nc = 0;
        __syncthreads();
      }
    }
    // count of a column / set a round flag, wherever the panel was counted
    auto count_of = [&](unsigned c) { return in_smem ? (h2v[tcw_h2_find(h2k, c)] & CNT_MASK) : (cnt[c] & CNT_MASK); };
    auto flag_or = [&](unsigned c, unsigned bit) { return in_smem ? atomicOr(&h2v[tcw_h2_find(h2k, c)], bit) : atomicOr(&cnt[c], bit); };
    // pass 1 on the global counters: count; the T-th hit of a column registers it.  Four independent (column load, atomic)
    // pairs per thread and step: the kernel waits on exactly these round trips (long scoreboard 54 % of its stall samples)
    for (int e0 = lb + threadIdx.x; e0 < (in_smem ? lb : ub); e0 += 4 * blockDim.x) {
      unsigned c[4], old[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = e0 + j * (int)blockDim.x < ub ? col[e0 + j * (int)blockDim.x] : 0xFFFFFFFFu;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c[j] != 0xFFFFFFFFu) atomicMax(&cnt[c[j]], ebase);  // a counter of an older epoch starts over (no return value needed)
#pragma unroll
      for (int j = 0; j < 4; ++j) old[j] = c[j] != 0xFFFFFFFFu ? (atomicAdd(&cnt[c[j]], 1u) & CNT_MASK) : 0xFFFFFFFFu;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((int)old[j] == T - 1 && c[j] != 0xFFFFFFFFu) {
          const int pos = atomicAdd(&s_nc, 1);
          if (pos < CAND_CAP) skeys[pos] = c[j];
        }
    }
    __syncthreads();
    int nc = s_nc;
    int Tcur = T;
    for (int round = 0; nc > CAND_CAP && round < 6; ++round) {  // too many candidates: double the bar
      Tcur *= 2;
      __syncthreads();
      if (threadIdx.x == 0) s_nc = 0;
      __syncthreads();
      const unsigned bit = 1u << (FLAG_SHIFT + round);
      for (int e = lb + threadIdx.x; e < ub; e += blockDim.x) {
        const unsigned c = col[e];
        if ((int)count_of(c) >= Tcur) {
          const unsigned old = flag_or(c, bit);
          if (!(old & bit)) {
            const int pos = atomicAdd(&s_nc, 1);
            if (pos < CAND_CAP) skeys[pos] = c;
          }
        }
      }
      __syncthreads();
      nc = s_nc;
    }
    int ns = 0;
    if (nc > 0 && nc <= CAND_CAP) {
      for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        const unsigned c = (unsigned)skeys[i];
        skeys[i] = ((unsigned long long)(0xFFFFFFFFu - count_of(c)) << 32) | c;
      }
      __syncthreads();
      tcw_sort_u64(skeys, nc);
      const int m = nc < W ? nc : W, mch = (m + 31) / 32;
      // S_j per chunk of 32 candidates (one warp per chunk), then the leading chunks that pay
      for (int j = warp; j < mch; j += nwarp) {
        const int i = j * 32 + lane;
        int sj = i < m ? (int)(0xFFFFFFFFu - (unsigned)(skeys[i] >> 32)) - 1 : 0;
        sj = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), sj, cg::plus<int>());
        if (lane == 0) hist[j] = sj;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int j = 0;
        long long g = 0;
        while (j < mch && hist[j] >= chunk_cost) { g += hist[j] - chunk_cost; ++j; }
        s_ns = g >= min_gain ? min(m, 32 * j) : 0;
        s_gain = g >= min_gain ? g : 0;
      }
      __syncthreads();
      ns = s_ns;
      const long long captured = s_gain;
      __syncthreads();
      if (ns > 0) {
        for (int i = threadIdx.x; i < ns; i += blockDim.x) skeys[i] &= 0xFFFFFFFFull;  // keep the column only
        __syncthreads();
        tcw_sort_u64(skeys, ns);
        tcw_hash_clear(hk, HT);
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += blockDim.x) tcw_hash_insert(hk, hs, HT, hbits, (unsigned)skeys[i], i);
        if (threadIdx.x == 0) atomicAdd(&stats[5], (unsigned long long)captured);  // total net gain
      }
    } else if (nc > CAND_CAP && threadIdx.x == 0) {
      atomicAdd(&stats[3], 1ull);
    }
    for (int i = threadIdx.x; i < W; i += blockDim.x) tc_cols[(size_t)p * W + i] = i < ns ? (int)(unsigned)skeys[i] : -1;
    if (threadIdx.x == 0) tc_ncol[p] = ns;
    for (int i = threadIdx.x; i < CH; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // pass 2: window nz of every row and of every 32-column chunk of the list.  All threads stride over the panel's nz
    // (a warp per row left the CTA waiting for the warp that drew a hub row: barrier 30 % of the stall samples); the row
    // of a nz is found in the panel's 129 row pointers in shared memory.
    for (int i = threadIdx.x; i <= BH; i += blockDim.x) { s_rp[i] = csr_v[p * BH + i]; if (i < BH) s_wl[i] = 0; }
    __syncthreads();
    if (ns > 0) {
      for (int e0 = lb + threadIdx.x; e0 < ub; e0 += 4 * blockDim.x) {
        unsigned f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = e0 + j * (int)blockDim.x < ub ? col[e0 + j * (int)blockDim.x] : HEMPTY;
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = f[j] != HEMPTY ? tcw_hash_find(hk, hs, HT, hbits, f[j]) : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (f[j] >> 31) {
            const int e = e0 + j * (int)blockDim.x;
            int lo = 0, hi = BH;  // last row whose first nz is <= e
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_rp[mid] <= e) lo = mid; else hi = mid; }
            atomicAdd(&s_wl[lo], 1);
            atomicAdd(&hist[(f[j] & 0xFFFFu) >> 5], 1);
          }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BH; i += blockDim.x) win_len[p * BH + i] = s_wl[i];
    __syncthreads();
    for (int i = threadIdx.x; i < CH; i += blockDim.x) chunk_len[(size_t)p * CH + i] = hist[i];
    // counts of this panel are stale for the next one by their epoch (nothing to undo), and the whole scratch is cleared
    // once at the end of the build
    __syncthreads();
  }
}

// MIN_TOTAL gate: windows that together do not pay for a second kernel are dropped, on the device
__global__ void k_tcw_gate(const unsigned long long* __restrict__ stats, long long min_total, int npanel, int nr, int W,
                           int* __restrict__ tc_cols, int* __restrict__ tc_ncol, int* __restrict__ win_len,
                           int* __restrict__ chunk_len) {
  if ((long long)stats[5] >= min_total) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)npanel * W) tc_cols[i] = -1;
  if (i < npanel) tc_ncol[i] = 0;
  if (i < nr) win_len[i] = 0;
  if (i < (long long)npanel * (W / 32)) chunk_len[i] = 0;
}

// exclusive scan by one CTA: out[n] = total
__global__ void __launch_bounds__(1024) k_tcw_scan(const int* __restrict__ in, int n, int* __restrict__ out) {
  fx::cta_exclusive_scan(in, n, out);
}

__global__ void k_tcw_rowptr(const int* __restrict__ csr_v, const int* __restrict__ win_rowptr, int nloc,
                             uint32_t* __restrict__ rest_rowptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nloc) rest_rowptr[i] = (uint32_t)(csr_v[i] - win_rowptr[i]);
}

// ascending list of the panels that have a window; tc_slot[p] = position in that list or -1
__global__ void __launch_bounds__(1024) k_tcw_panels(const int* __restrict__ tc_ncol, const int* __restrict__ win_rowptr, int npanel,
                                                     int nr, int* __restrict__ tc_panels, int* __restrict__ tc_slot,
                                                     unsigned long long* __restrict__ stats) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < npanel; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < npanel && tc_ncol[i] > 0) ? 1 : 0;
    const int inc = cg::inclusive_scan(warp, v);
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ws = warp_sum[threadIdx.x];
      const int wi = cg::inclusive_scan(warp, ws);
      warp_sum[threadIdx.x] = wi - ws;
    }
    __syncthreads();
    const int incl = carry_s + warp_sum[threadIdx.x >> 5] + inc;
    if (i < npanel) {
      tc_slot[i] = v ? incl - 1 : -1;
      if (v) tc_panels[incl - 1] = i;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  // stats: [0] window nz, [1] listed columns, [2] panels with a window
  long long cols = 0;
  for (int i = threadIdx.x; i < npanel; i += blockDim.x) cols += tc_ncol[i];
  cols = cg::reduce(warp, cols, cg::plus<long long>());
  if ((threadIdx.x & 31) == 0 && cols) atomicAdd(&stats[1], (unsigned long long)cols);
  if (threadIdx.x == 0) { stats[2] = (unsigned long long)carry_s; stats[0] = (unsigned long long)win_rowptr[nr]; }
}

// split of every row into its window part (chunk-major inside the panel) and its remainder (row order)
__global__ void __launch_bounds__(1024) k_tcw_split(const int* __restrict__ csr_v, const uint32_t* __restrict__ col,
                                                   const float* __restrict__ val, const int* __restrict__ win_rowptr,
                                                   const int* __restrict__ win_cptr, const int* __restrict__ tc_cols,
                                                   const int* __restrict__ tc_ncol, int npanel, int ncols, int W,
                                                   unsigned* __restrict__ cnt_all, uint16_t* __restrict__ win_code,
                                                   float* __restrict__ win_val, uint32_t* __restrict__ rest_col,
                                                   float* __restrict__ rest_val, unsigned long long* __restrict__ sched) {
  extern __shared__ int off2[];  // [CH][128] nz of (chunk, row), then their exclusive scan; then the listed-column table
  __shared__ int wsum[32];
  __shared__ int s_rp[BH + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int CH = W / 32, n2 = CH * BH;
  const int HT = tcw_hash_size(W), hbits = 31 - __clz(HT);
  unsigned* hk = reinterpret_cast<unsigned*>(off2 + n2);
  int* s_cnt = reinterpret_cast<int*>(hk + HT);  // [nwarp][CH] window nz of a step per warp and chunk (long rows)
  int* s_run = s_cnt + nwarp * CH;               // [CH] window nz of the long row placed so far
  unsigned short* hs = reinterpret_cast<unsigned short*>(s_run + CH);
  __shared__ int s_rest[32], s_rrest, s_next;
  constexpr int LONG_ROW = 512;
  __shared__ int s_panel;
  for (;;) {  // panels by a counter, as in k_tcw_select
    __syncthreads();
    if (threadIdx.x == 0) s_panel = (int)atomicAdd(sched, 1ull);
    __syncthreads();
    const int p = s_panel;
    if (p >= npanel) break;
    const int ns = tc_ncol[p];
    const int* list = tc_cols + (size_t)p * W;
    if (ns > 0) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) off2[i] = 0;
      tcw_hash_clear(hk, HT);
    }
    for (int i = threadIdx.x; i <= BH; i += blockDim.x) s_rp[i] = csr_v[p * BH + i];
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += blockDim.x) tcw_hash_insert(hk, hs, HT, hbits, (unsigned)list[i], i);
    __syncthreads();
    if (ns > 0) {
      // nz of every (chunk, row): all threads stride over the panel's nz, the row of a nz comes from the panel's row pointers
      const int lb = s_rp[0], ub = s_rp[BH];
      for (int e0 = lb + threadIdx.x; e0 < ub; e0 += 4 * blockDim.x) {
        unsigned f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = e0 + j * (int)blockDim.x < ub ? col[e0 + j * (int)blockDim.x] : HEMPTY;
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = f[j] != HEMPTY ? tcw_hash_find(hk, hs, HT, hbits, f[j]) : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (f[j] >> 31) {
            const int e = e0 + j * (int)blockDim.x;
            int lo = 0, hi = BH;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_rp[mid] <= e) lo = mid; else hi = mid; }
            atomicAdd(&off2[((f[j] & 0xFFFFu) >> 5) * BH + lo], 1);
          }
      }
      __syncthreads();
      // exclusive scan of off2 in (chunk, row) order
      const int ipt = (n2 + (int)blockDim.x - 1) / (int)blockDim.x;
      const int b = threadIdx.x * ipt, e = min(n2, b + ipt);
      int sum = 0;
      for (int i = b; i < e; ++i) sum += off2[i];
      auto w32 = cg::tiled_partition<32>(cg::this_thread_block());
      const int inc = cg::inclusive_scan(w32, sum);
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (threadIdx.x < 32) {
        const int ws = threadIdx.x < nwarp ? wsum[threadIdx.x] : 0;
        const int wi = cg::inclusive_scan(w32, ws);
        if (threadIdx.x < nwarp) wsum[threadIdx.x] = wi - ws;
      }
      __syncthreads();
      int run = wsum[warp] + inc - sum;
      for (int i = b; i < e; ++i) { const int v = off2[i]; off2[i] = run; run += v; }
      __syncthreads();
    }
    const int wbase = win_cptr[(size_t)p * CH];
    // Placing pass.  A warp per row left the CTA waiting for the warp that drew a hub row (Reddit-shape: the longest warp of a
    // panel does 3.6 times the average work): rows of LONG_ROW nz or more are placed by the whole CTA, 32 nz per warp and
    // step, the ranks of a step carried across warps through per-warp counts in shared memory; the other rows are handed
    // to the warps by a counter.
    for (int r = 0; r < BH; ++r) {
      const int lo = s_rp[r], hi = s_rp[r + 1];
      if (hi - lo < LONG_ROW) continue;
      const int row = p * BH + r;
      for (int i = threadIdx.x; i < CH; i += blockDim.x) s_run[i] = 0;
      if (threadIdx.x == 0) s_rrest = lo - win_rowptr[row];
      unsigned c_n = 0;
      float v_n = 0.f;
      if (lo + (int)threadIdx.x < hi) { c_n = col[lo + threadIdx.x]; v_n = val[lo + threadIdx.x]; }
      for (int e0 = lo; e0 < hi; e0 += (int)blockDim.x) {
        const int e = e0 + threadIdx.x;
        const bool valid = e < hi;
        const unsigned c = c_n, f = (valid && ns > 0) ? tcw_hash_find(hk, hs, HT, hbits, c_n) : 0u;
        const float v = v_n;
        if (e + (int)blockDim.x < hi) { c_n = col[e + blockDim.x]; v_n = val[e + blockDim.x]; }
        const bool isw = (f >> 31) != 0u;
        const unsigned mw = __ballot_sync(0xffffffffu, isw), mr = __ballot_sync(0xffffffffu, valid && !isw);
        for (int i = lane; i < CH; i += 32) s_cnt[warp * CH + i] = 0;
        __syncwarp();
        int ch = -1, sl = 0;
        unsigned peers = 0;
        if (isw) {
          sl = (int)(f & 0xFFFFu);
          ch = sl >> 5;
          peers = __match_any_sync(mw, ch);
          if ((peers & lt) == 0u) s_cnt[warp * CH + ch] = __popc(peers);
        }
        if (lane == 0) s_rest[warp] = __popc(mr);
        __syncthreads();  // (also orders the s_run / s_rrest updates of the previous step)
        if (isw) {
          int rank = s_run[ch] + __popc(peers & lt);
          for (int w2 = 0; w2 < warp; ++w2) rank += s_cnt[w2 * CH + ch];
          const int pos = wbase + off2[ch * BH + r] + rank;
          win_code[pos] = (uint16_t)fxtc::tile_word(r, sl & 31);
          win_val[pos] = v;
        } else if (valid) {
          int pos = s_rrest + __popc(mr & lt);
          for (int w2 = 0; w2 < warp; ++w2) pos += s_rest[w2];
          rest_col[pos] = c;
          rest_val[pos] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < CH; i += blockDim.x) {
          int tsum = 0;
          for (int w2 = 0; w2 < nwarp; ++w2) tsum += s_cnt[w2 * CH + i];
          s_run[i] += tsum;
        }
        if (threadIdx.x == 0) {
          int tsum = 0;
          for (int w2 = 0; w2 < nwarp; ++w2) tsum += s_rest[w2];
          s_rrest += tsum;
        }
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    for (;;) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_next, 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= BH) break;
      const int row = p * BH + r;
      const int lo = s_rp[r], hi = s_rp[r + 1];
      if (hi - lo >= LONG_ROW) continue;
      int ro = lo - win_rowptr[row];
      int last_ch = -1, last_cnt = 0;
      // the (column, value) loads of the next step are requested before this step's ranks are worked out
      unsigned c_n = 0;
      float v_n = 0.f;
      if (lo + lane < hi) { c_n = col[lo + lane]; v_n = val[lo + lane]; }
      for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < hi;
        const unsigned c = c_n, f = (valid && ns > 0) ? tcw_hash_find(hk, hs, HT, hbits, c_n) : 0u;
        const float v = v_n;
        if (e + 32 < hi) { c_n = col[e + 32]; v_n = val[e + 32]; }
        const bool isw = (f >> 31) != 0u;
        const unsigned mw = __ballot_sync(0xffffffffu, isw), mr = __ballot_sync(0xffffffffu, valid && !isw);
        int ch = -1, pc = 0;
        if (isw) {
          const int s = (int)(f & 0xFFFFu);
          ch = s >> 5;
          const unsigned peers = __match_any_sync(mw, ch);
          pc = __popc(peers);
          const int rank = __popc(peers & lt) + (ch == last_ch ? last_cnt : 0);
          const int pos = wbase + off2[ch * BH + r] + rank;
          win_code[pos] = (uint16_t)fxtc::tile_word(r, s & 31);
          win_val[pos] = v;
        } else if (valid) {
          const int pos = ro + __popc(mr & lt);
          rest_col[pos] = c;
          rest_val[pos] = v;
        }
        ro += __popc(mr);
        if (mw) {  // nz are in ascending slot order: only the last chunk of this batch can continue
          const int hl = 31 - __clz(mw);
          const int ch_h = __shfl_sync(0xffffffffu, ch, hl), pc_h = __shfl_sync(0xffffffffu, pc, hl);
          last_cnt = (ch_h == last_ch ? last_cnt : 0) + pc_h;
          last_ch = ch_h;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

namespace fx {

size_t tcw_arena_bytes(const fx_tiles* t) {
  const fx_aspt_dev& a = t->aspt;
  const fx_tcw_dev& w = t->tcw;
  const int64_t ne = t->nnz_local, nr = a.nr, npanel = a.npanel;
  size_t bytes = 0;
  auto add = [&](size_t b) { bytes += Arena::pad(b) + 256; };
  add(sizeof(int) * (size_t)npanel * w.W);
  for (int i = 0; i < 3; ++i) add(sizeof(int) * (npanel + 2));
  for (int i = 0; i < 4; ++i) add(sizeof(int) * (nr + 2));
  for (int i = 0; i < 2; ++i) add(sizeof(int) * ((size_t)npanel * (w.W / 32) + 2));
  add(sizeof(uint16_t) * (ne + 2));
  add(sizeof(float) * (ne + 2) * 2);
  add(sizeof(uint32_t) * (ne + 2));
  add(sizeof(float) * (size_t)nr * t->k);
  add(sizeof(unsigned long long) * 8);
  return bytes;
}

int tcw_carve(fx_tiles* t) {
  const fx_aspt_dev& a = t->aspt;
  fx_tcw_dev& w = t->tcw;
  const int64_t ne = t->nnz_local, nr = a.nr, npanel = a.npanel;
  Arena& A = t->arena;
  w.tc_cols = A.take<int>((size_t)npanel * w.W);
  w.tc_ncol = A.take<int>(npanel + 2);
  w.tc_slot = A.take<int>(npanel + 2);
  w.tc_panels = A.take<int>(npanel + 2);
  w.csr_v = A.take<int>(nr + 2);
  w.win_len = A.take<int>(nr + 2);
  w.win_rowptr = A.take<int>(nr + 2);
  w.rest_rowptr = A.take<uint32_t>(nr + 2);
  w.chunk_len = A.take<int>((size_t)npanel * (w.W / 32) + 2);
  w.win_cptr = A.take<int>((size_t)npanel * (w.W / 32) + 2);
  w.win_code = A.take<uint16_t>(ne + 2);
  w.win_val = A.take<float>(ne + 2);
  w.rest_val = A.take<float>(ne + 2);
  w.rest_col = A.take<uint32_t>(ne + 2);
  w.tc_out = A.take<float>((size_t)nr * t->k);
  w.stats = A.take<unsigned long long>(8);
  if (!w.stats || !w.tc_out) { set_error("tcw arena carve overflow"); return FX_ERR_NOMEM; }
  return FX_OK;
}

// Runs in front of aspt_build inside the timed region of fx_build.
int tcw_build(fx_tiles* t, cudaStream_t s) {
  fx_aspt_dev& a = t->aspt;
  fx_tcw_dev& w = t->tcw;
  const fx_matrix* m = t->mat;
  const int nloc = t->row_end - t->row_begin;
  const int ne = (int)t->nnz_local;
  const uint32_t* col = m->col_dev + m->rowptr[t->row_begin];
  const float* val = m->val_dev + m->rowptr[t->row_begin];
  FX_CUDA(cudaMemsetAsync(w.stats, 0, sizeof(unsigned long long) * 8, s));
  k_tcw_pad<<<ceil_div(a.nr + 1, 256), 256, 0, s>>>(m->rowptr_dev, t->row_begin, nloc, a.nr, ne, w.csr_v);
  FX_LAUNCH_CHECK();
  static SmemAttr select_attr;
  const size_t hash_bytes = (size_t)tcw_hash_size(w.W) * (sizeof(unsigned) + sizeof(unsigned short));
  // the pass-1 table needs the whole SM's shared memory: only with one CTA per SM (the default, fx_common.cuh:build_threads)
  static const int h2_env = getenv("FLEX_SELECT_SMEM") ? atoi(getenv("FLEX_SELECT_SMEM")) : 1;
  // (and only while it fits beside the candidate keys and the listed-column table: W = 4096 needs 48 KB for the latter)
  const size_t select_base = CAND_CAP * sizeof(unsigned long long) + hash_bytes;
  const int use_h2 = (h2_env && a.G <= sm_count_of_current_device() && select_base + (size_t)HT2 * 8 <= 216 * 1024) ? 1 : 0;
  const size_t select_smem = select_base + (use_h2 ? (size_t)HT2 * 8 : 0);
  if (int rc = select_attr.ensure(k_tcw_select, select_smem)) return rc;
  k_tcw_select<<<a.G, build_threads(), select_smem, s>>>(w.csr_v, col, a.npanel, (int)m->n, w.T, w.W, w.min_gain,
                                                                      w.chunk_cost, a.cnt_scratch, w.tc_cols, w.tc_ncol, w.win_len, w.chunk_len, w.stats, use_h2);
  FX_LAUNCH_CHECK();
  {
    const long long cells = std::max<long long>((long long)a.npanel * w.W, a.nr);
    k_tcw_gate<<<ceil_div(cells, 256), 256, 0, s>>>(w.stats, w.min_total, a.npanel, a.nr, w.W, w.tc_cols, w.tc_ncol, w.win_len,
                                                     w.chunk_len);
    FX_LAUNCH_CHECK();
  }
  FX_CUDA(device_exclusive_scan(a.spec_sort_tmp, a.spec_sort_tmp_bytes, w.win_len, a.nr, w.win_rowptr, s));
  k_tcw_scan<<<1, 1024, 0, s>>>(w.chunk_len, a.npanel * (w.W / 32), w.win_cptr);
  FX_LAUNCH_CHECK();
  k_tcw_rowptr<<<ceil_div(nloc + 1, 256), 256, 0, s>>>(w.csr_v, w.win_rowptr, nloc, w.rest_rowptr);
  FX_LAUNCH_CHECK();
  k_tcw_panels<<<1, 1024, 0, s>>>(w.tc_ncol, w.win_rowptr, a.npanel, a.nr, w.tc_panels, w.tc_slot, w.stats);
  FX_LAUNCH_CHECK();
  // the split keeps no per-CTA counters, so its grid is its own: two 512-thread CTAs per SM (one of 1024 was slower, 0.74 vs
  // 0.58 ms on Reddit-shape: its long-row steps synchronise the whole CTA); panels are handed out by a counter
  const int split_ctas = std::min(2 * sm_count_of_current_device(), std::max(1, a.npanel));
  const size_t split_smem = sizeof(int) * (size_t)(w.W / 32) * BH + hash_bytes + sizeof(int) * (size_t)(w.W / 32) * 33;
  static SmemAttr split_attr;
  if (int rc = split_attr.ensure(k_tcw_split, split_smem)) return rc;
  k_tcw_split<<<split_ctas, 512, split_smem, s>>>(w.csr_v, col, val, w.win_rowptr, w.win_cptr, w.tc_cols, w.tc_ncol, a.npanel, (int)m->n,
                                           w.W, a.cnt_scratch, w.win_code, w.win_val, w.rest_col, w.rest_val, w.stats + 7);
  FX_LAUNCH_CHECK();
  // k_tcw_select leaves stale (epoch, count) words behind; the ASpT builder expects zeros (a streaming memset: 46 us for the
  // 276 MB of Reddit-shape, against the ~0.3 ms the third pass over every panel's nz cost)
  FX_CUDA(cudaMemsetAsync(a.cnt_scratch, 0, sizeof(unsigned) * (size_t)a.G * (size_t)m->n, s));
  FX_CUDA(cudaMemcpyAsync(t->stats_host, w.stats, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
  FX_CUDA(cudaStreamSynchronize(s));
  w.win_nnz = (long long)t->stats_host[0];
  w.ncols_listed = (long long)t->stats_host[1];
  w.ntc = (int)t->stats_host[2];
  w.dropped = (int)t->stats_host[3];
  w.net_gain = (long long)t->stats_host[5];
  // the remainder is what the ASpT builder sees
  t->src_rowptr = w.rest_rowptr;
  t->src_row0 = 0;
  t->src_col = w.rest_col;
  t->src_val = w.rest_val;
  a.ne = ne - (int)w.win_nnz;
  return FX_OK;
}

}  // namespace fx

template <class T>
static int tcw_d2h(std::vector<T>& dst, const void* src, size_t count) {
  dst.resize(count);
  if (count) FX_CUDA(cudaMemcpy(dst.data(), src, sizeof(T) * count, cudaMemcpyDeviceToHost));
  return FX_OK;
}

extern "C" int fx_tiles_export_tcw(fx_tiles* t, fx_tcw_arrays* o) {
  FX_REQUIRE(t && o && t->format == FX_FMT_TCW, FX_ERR_ARG, "fx_tiles_export_tcw: not a tensor-window handle");
  fx_tcw_dev& w = t->tcw;
  const fx_aspt_dev& a = t->aspt;
  FX_CUDA(cudaDeviceSynchronize());
  const int nloc = t->row_end - t->row_begin;
  const size_t rest = (size_t)(t->nnz_local - w.win_nnz);
  int rc;
  if ((rc = tcw_d2h(w.h_cols, w.tc_cols, (size_t)a.npanel * w.W))) return rc;
  if ((rc = tcw_d2h(w.h_ncol, w.tc_ncol, a.npanel))) return rc;
  if ((rc = tcw_d2h(w.h_win_cptr, w.win_cptr, (size_t)a.npanel * (w.W / 32) + 1))) return rc;
  if ((rc = tcw_d2h(w.h_win_code, w.win_code, (size_t)w.win_nnz))) return rc;
  if ((rc = tcw_d2h(w.h_win_val, w.win_val, (size_t)w.win_nnz))) return rc;
  if ((rc = tcw_d2h(w.h_rest_rowptr, w.rest_rowptr, (size_t)nloc + 1))) return rc;
  if ((rc = tcw_d2h(w.h_rest_col, w.rest_col, rest))) return rc;
  if ((rc = tcw_d2h(w.h_rest_val, w.rest_val, rest))) return rc;
  o->n = nloc; o->nr = a.nr; o->npanel = a.npanel; o->W = w.W; o->T = w.T; o->min_gain = w.min_gain;
  o->ntc = w.ntc; o->dropped = w.dropped; o->chunk_cost = w.chunk_cost; o->min_total = w.min_total; o->net_gain = w.net_gain;
  o->win_nnz = w.win_nnz; o->rest_nnz = (int64_t)rest;
  o->tc_cols = w.h_cols.data(); o->tc_ncol = w.h_ncol.data(); o->win_cptr = w.h_win_cptr.data();
  o->win_code = w.h_win_code.data(); o->win_val = w.h_win_val.data();
  o->rest_rowptr = w.h_rest_rowptr.data(); o->rest_col = w.h_rest_col.data(); o->rest_val = w.h_rest_val.data();
  return FX_OK;
}

extern "C" int fx_tiles_tcw_info(const fx_tiles* t, int64_t out[8]) {
  FX_REQUIRE(t && out && t->format == FX_FMT_TCW, FX_ERR_ARG, "fx_tiles_tcw_info: not a tensor-window handle");
  const fx_tcw_dev& w = t->tcw;
  out[0] = t->aspt.npanel; out[1] = w.ntc; out[2] = w.win_nnz; out[3] = t->nnz_local - w.win_nnz;
  out[4] = w.ncols_listed; out[5] = w.net_gain; out[6] = w.W; out[7] = w.T;
  return FX_OK;
}
