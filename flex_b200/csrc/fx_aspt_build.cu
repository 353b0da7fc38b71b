// fx_aspt_build.cu -- GPU builder of the ASpT dense/sparse tile format.
//
// Produces, bit for bit, the layout the reference's pre-process section builds
// (aspt/sspmm_128.cu:1207-1333 with kernels :831-1087) under the canonical tie-breaks
// (heavy columns take slots in ascending column order; a row's nz keep their column order
// inside each group), but by a different route that suits a B200:
//   * the reference sorts every panel's nz by column (bb_segsort #1) to find heavy columns and
//     then sorts every row by tile id (bb_segsort #2); here heavy columns are found by counting
//     into an L2-resident per-CTA column counter array (one atomic per nz), only the few heavy
//     columns of a panel are sorted (shared-memory bitonic network), and rows are partitioned by
//     a warp-level stable counting scatter (match_any + popc) -- no full sort of the nz at all;
//   * no allocation, no host round trip except one 64-byte read of the statistics at the end
//     (the reference does ~20 cudaMalloc/cudaFree and 5 blocking copies inside tPre).
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>

#include <cub/device/device_radix_sort.cuh>

#include "fx_common.cuh"
#include "fx_scan.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int BH = 128;          // panel height            aspt/sspmm_128.cu:33
constexpr int THRESHOLD = 16;    // heavy column threshold  :32
constexpr int SC_SIZE = 2048;    // detect histogram        :46
constexpr int STHRESHOLD = 512;  // long-row chunk          :44
constexpr int SPARSE_KEY = 30000;  // :890
constexpr int OCC_SIZE = 1024;   // MCSR_CNT_SIZE :895
constexpr int SORT_SMEM_CAP = 8192;  // heavy columns sorted in shared memory up to this many

// ---- K0: padded local row pointer ---------------------------------------------------------
__global__ void k_pad_rowptr(const uint32_t* __restrict__ rowptr, int row0, int nloc, int nr, int ne,
                             int* __restrict__ csr_v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nr) csr_v[i] = i <= nloc ? (int)(rowptr[row0 + i] - rowptr[row0]) : ne;
}

// ---- K1: dense_block_detect (:831-868) ----------------------------------------------------
__global__ void __launch_bounds__(1024) k_detect(const int* __restrict__ csr_v, const uint32_t* __restrict__ col,
                                                int min_occ, int* __restrict__ chk,
                                                unsigned long long* __restrict__ stats) {
  __shared__ int hist[SC_SIZE];
  __shared__ int total;
  const int p = blockIdx.x;
  for (int i = threadIdx.x; i < SC_SIZE; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int lb = csr_v[p * BH], ub = csr_v[(p + 1) * BH];
  for (int i = lb + threadIdx.x; i < ub; i += blockDim.x) atomicAdd(&hist[col[i] & (SC_SIZE - 1)], 1);
  __syncthreads();
  int r = 0;
  for (int i = threadIdx.x; i < SC_SIZE; i += blockDim.x) r += hist[i] >= THRESHOLD;
  r = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), r, cg::plus<int>());
  if ((threadIdx.x & 31) == 0 && r) atomicAdd(&total, r);
  __syncthreads();
  if (threadIdx.x == 0) {
    int f = total >= min_occ;
    chk[p] = f;
    if (f) atomicOr(&stats[4], 1ull);
  }
}

// ---- K2: heavy columns, slot depths, tile count and per-nz tile id of flagged panels -----
// (mcsr_cnt_calc :896-925 + key2_marking :929-981 without the panel sort)
__device__ __forceinline__ void cmpswap(unsigned long long& a, unsigned long long& b) {
  if (a > b) { unsigned long long t = a; a = b; b = t; }
}

// normalised bitonic network: every comparator moves the minimum to the lower index, so indices
// >= n behave as +inf padding and arbitrary n sorts correctly.
__device__ void block_sort_u64(unsigned long long* a, int n) {
  int N = 1;
  while (N < n) N <<= 1;
  for (int k = 2; k <= N; k <<= 1) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      int j = i ^ (k - 1);
      if (j > i && j < n) cmpswap(a[i], a[j]);
    }
    __syncthreads();
    for (int s = k >> 2; s > 0; s >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int j = i ^ s;
        if (j > i && j < n) cmpswap(a[i], a[j]);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024) k_heavy(const int* __restrict__ csr_v, const uint32_t* __restrict__ col,
                                               const int* __restrict__ chk, int npanel, int BW, int min_occ,
                                               int ncols, unsigned* __restrict__ cnt_all,
                                               unsigned long long* __restrict__ heavy_all, int* __restrict__ nheavy,
                                               int* __restrict__ tcount, uint16_t* __restrict__ key2) {
  extern __shared__ unsigned long long skeys[];  // SORT_SMEM_CAP
  __shared__ int occ[OCC_SIZE];
  __shared__ int gstart[256];
  __shared__ int s_nh, s_tp;
  unsigned* cnt = cnt_all + (size_t)blockIdx.x * ncols;
  const int wmask = BW - 1;
  for (int p = blockIdx.x; p < npanel; p += gridDim.x) {
    if (!chk[p]) continue;  // uniform
    const int lb = csr_v[p * BH], ub = csr_v[(p + 1) * BH];
    unsigned long long* heavy = heavy_all + (lb / THRESHOLD + p);
    if (threadIdx.x == 0) { s_nh = 0; s_tp = 0; }
    for (int i = threadIdx.x; i < OCC_SIZE; i += blockDim.x) occ[i] = 0;
    __syncthreads();
    // pass 1: count columns; the 16th hit of a column registers it as heavy
    for (int e = lb + threadIdx.x; e < ub; e += blockDim.x) {
      unsigned c = col[e];
      unsigned old = atomicAdd(&cnt[c], 1u);
      if (old == THRESHOLD - 1) {
        int pos = atomicAdd(&s_nh, 1);
        heavy[pos] = ((unsigned long long)(c & wmask) << 32) | c;
      }
    }
    __syncthreads();
    const int nh = s_nh;
    // sort heavy columns by (slot width, column): ascending column inside each slot = canonical
    unsigned long long* keys = heavy;
    if (nh <= SORT_SMEM_CAP) {
      for (int i = threadIdx.x; i < nh; i += blockDim.x) skeys[i] = heavy[i];
      keys = skeys;
    }
    __syncthreads();
    block_sort_u64(keys, nh);
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      int w = (int)(keys[i] >> 32);
      if (i == 0 || (int)(keys[i - 1] >> 32) != w) gstart[w] = i;
    }
    __syncthreads();
    // occ'[d+1] = #slots with more than d heavy columns (:915-916); occ'[0] = BW
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      int d = i - gstart[(int)(keys[i] >> 32)];
      if (d + 1 < OCC_SIZE) atomicAdd(&occ[d + 1], 1);
    }
    if (threadIdx.x == 0) occ[0] = BW;
    __syncthreads();
    for (int t = threadIdx.x; t < OCC_SIZE - 1; t += blockDim.x)
      if (occ[t] >= min_occ && occ[t + 1] < min_occ) s_tp = t;  // :921-922 (unique t: occ is non-increasing)
    __syncthreads();
    const int tp = s_tp;
    // publish (column, depth); mark the columns that made it into a tile in the counter array
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      unsigned long long kk = keys[i];
      unsigned c = (unsigned)kk;
      int d = i - gstart[(int)(kk >> 32)];
      if (d < tp) cnt[c] = 0x80000000u | (unsigned)d;
      heavy[i] = ((unsigned long long)(unsigned)d << 32) | c;  // read as int2 {c, depth}
    }
    if (threadIdx.x == 0) { nheavy[p] = nh; tcount[p] = tp; }
    __syncthreads();
    // pass 2: tile id of every nz (key2, :971-978); pass 3 restores the counters to zero
    if (tp > 0) {
      for (int e = lb + threadIdx.x; e < ub; e += blockDim.x) {
        unsigned x = cnt[col[e]];
        key2[e] = (x & 0x80000000u) ? (uint16_t)(x & 0xffffu) : (uint16_t)SPARSE_KEY;
      }
      __syncthreads();
    }
    for (int e = lb + threadIdx.x; e < ub; e += blockDim.x) cnt[col[e]] = 0u;
    __syncthreads();
  }
}

// ---- K3: mcsr_cnt prefix (host loop :1260-1266 moved to the device) ------------------------
__global__ void __launch_bounds__(1024) k_scan_tcount(const int* __restrict__ tcount, int npanel,
                                                      int* __restrict__ mcsr_cnt,
                                                      unsigned long long* __restrict__ stats) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  __shared__ int maxtp;
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  if (threadIdx.x == 0) { carry_s = 0; mcsr_cnt[0] = 0; maxtp = 0; }
  __syncthreads();
  int mymax = 0;
  for (int base = 0; base < npanel; base += blockDim.x) {
    int i = base + threadIdx.x;
    int t = i < npanel ? tcount[i] : 0;
    mymax = max(mymax, t);
    int v = i < npanel ? t + 1 : 0;
    int inc = cg::inclusive_scan(warp, v);
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int ws = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0;
      int wi = cg::inclusive_scan(warp, ws);
      warp_sum[threadIdx.x] = wi - ws;
    }
    __syncthreads();
    int total = carry_s + warp_sum[threadIdx.x >> 5] + inc;
    if (i < npanel) mcsr_cnt[i + 1] = total;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = total;
    __syncthreads();
  }
  mymax = cg::reduce(warp, mymax, cg::greater<int>());
  if ((threadIdx.x & 31) == 0) atomicMax(&maxtp, mymax);
  __syncthreads();
  if (threadIdx.x == 0) { stats[3] = (unsigned long long)(carry_s - npanel); stats[5] = (unsigned long long)maxtp; }
}

// ordered compaction of the panel ids into "has dense tiles" / "has none" (SpMM launches one grid each)
__global__ void __launch_bounds__(1024) k_panel_lists(const int* __restrict__ tcount, int npanel,
                                                      int* __restrict__ plain, int* __restrict__ tiled,
                                                      unsigned long long* __restrict__ stats) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < npanel; base += blockDim.x) {
    int i = base + threadIdx.x;
    int f = i < npanel ? (tcount[i] > 0) : 0;
    int inc = cg::inclusive_scan(warp, f);
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int ws = warp_sum[threadIdx.x];
      int wi = cg::inclusive_scan(warp, ws);
      warp_sum[threadIdx.x] = wi - ws;
    }
    __syncthreads();
    int incl = carry_s + warp_sum[threadIdx.x >> 5] + inc;  // tiled panels among [0, i]
    if (i < npanel) {
      if (f) tiled[incl - 1] = i;
      else plain[i - incl] = i;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) stats[6] = (unsigned long long)carry_s;
}

// ---- work list of one SpMM launch: panels (all, or those of `plist`) cut into parts by their handled nz ----
constexpr int WL_MAX_PARTS = 16;
// mode 0: every panel; 1: the panels of `plist` = plist_plain; 2: plist_tiled (their counts come from stats[6])
__global__ void __launch_bounds__(1024) k_worklist(const int* __restrict__ plist, int mode, const int* __restrict__ csr_v,
                                                   const int* __restrict__ spec_off, int nr, int npanel_all,
                                                   int2* __restrict__ wl, int cap, const unsigned long long* __restrict__ stats,
                                                   unsigned long long* __restrict__ count, int nclass) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  const int n_tiled = (int)stats[6];
  const int npan = mode == 0 ? npanel_all : (mode == 1 ? npanel_all - n_tiled : n_tiled);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  // a part should hold about twice the average panel's handled nz (everything but the 512-chunks)
  const long long handled_all = (long long)csr_v[nr] - (long long)STHRESHOLD * spec_off[nr];
  const long long mean = handled_all / (npanel_all > 0 ? npanel_all : 1);
  const long long target = mean * 2 > 4096 ? mean * 2 : 4096;
  // nclass > 1: the list is emitted heaviest class first (work per entry >= 1.5, 1, 0.5 times the mean panel, the rest;
  // panel order inside a class), so that the grid ends on short CTAs instead of on whatever the last panels hold
  for (int cls = 0; cls < nclass; ++cls)
  for (int base = 0; base < npan; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int parts = 0, p = 0;
    if (i < npan) {
      p = plist ? plist[i] : i;
      const long long h = (long long)(csr_v[(p + 1) * BH] - csr_v[p * BH]) -
                          (long long)STHRESHOLD * (spec_off[(p + 1) * BH] - spec_off[p * BH]);
      parts = (int)((h + target - 1) / target);
      parts = parts < 1 ? 1 : (parts > WL_MAX_PARTS ? WL_MAX_PARTS : parts);
      if (nclass > 1) {
        const long long w2 = 2 * h / parts;  // twice the work of one entry
        const int c = w2 >= 3 * mean ? 0 : (w2 >= 2 * mean ? 1 : (w2 >= mean ? 2 : 3));
        if ((c < nclass ? c : nclass - 1) != cls) parts = 0;
      }
    }
    const int inc = cg::inclusive_scan(warp, parts);
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ws = warp_sum[threadIdx.x];
      const int wi = cg::inclusive_scan(warp, ws);
      warp_sum[threadIdx.x] = wi - ws;
    }
    __syncthreads();
    const int incl = carry_s + warp_sum[threadIdx.x >> 5] + inc;
    for (int q = 0; q < parts; ++q) {
      const int o = incl - parts + q;
      if (o < cap) wl[o] = make_int2(p, q | (parts << 8));
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = (unsigned long long)(carry_s < cap ? carry_s : cap);
}

// ---- K4: slot lists, per-row group offsets, stable partition of the nz ---------------------
// (key2_marking's list writes :958-959,946-949; bb_segsort#2 :1282; fill_mcsre :1006-1038;
//  porting :1040-1048; the length statistics of cal_vari :1050-1073)
constexpr int FILL_WARPS = 8;
__global__ void __launch_bounds__(FILL_WARPS * 32) k_fill(
    const int* __restrict__ csr_v, const uint32_t* __restrict__ col, const float* __restrict__ val,
    const int* __restrict__ tcount, const int* __restrict__ mcsr_cnt, const int2* __restrict__ heavy_all,
    const int* __restrict__ nheavy, const uint16_t* __restrict__ key2, int npanel, int BW, int ne,
    int* __restrict__ mcsr_e, int* __restrict__ mcsr_list, int* __restrict__ baddr, int* __restrict__ saddr,
    int* __restrict__ perm, int* __restrict__ csr_e, float* __restrict__ csr_ev, int* __restrict__ spec_cnt,
    unsigned long long* __restrict__ stats) {
  __shared__ int hist[FILL_WARPS][OCC_SIZE];
  const int p = blockIdx.x;
  const int tp = tcount[p], cnt0 = mcsr_cnt[p], delta = tp + 1, g0 = cnt0 - p;
  const int wmask = BW - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (tp > 0) {
    for (int i = threadIdx.x; i < tp * BW; i += blockDim.x) mcsr_list[(size_t)g0 * BW + i] = -1;  // memset -1 :1128
    for (int i = threadIdx.x; i < tp; i += blockDim.x) { baddr[g0 + i] = p; saddr[g0 + i] = i; }
    __syncthreads();
    const int2* heavy = heavy_all + (csr_v[p * BH] / THRESHOLD + p);
    const int nh = nheavy[p];
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      int2 h = heavy[i];
      if (h.y < tp) mcsr_list[(size_t)(g0 + h.y) * BW + (h.x & wmask)] = h.x;
    }
  }
  long long s1 = 0, s2 = 0;
  int chunks = 0;
  if (tp == 0) {
    // No dense tile in this panel: the nz keep their order.  One flat, coalesced copy by the whole CTA (a warp per row made
    // the CTA wait for the warp that drew a hub row), then the per-row words.
    const int lb = csr_v[p * BH], ub = csr_v[(p + 1) * BH];
    for (int e = lb + threadIdx.x; e < ub; e += blockDim.x) { perm[e] = e; csr_e[e] = (int)col[e]; csr_ev[e] = val[e]; }
    for (int r = threadIdx.x; r < BH; r += blockDim.x) {
      const int row = p * BH + r;
      const int rs = csr_v[row], len = csr_v[row + 1] - rs;
      mcsr_e[cnt0 * BH + r] = rs;
      spec_cnt[row] = len / STHRESHOLD;
      s1 += len; s2 += (long long)len * len; chunks += len / STHRESHOLD;
    }
  }
  for (int r = tp == 0 ? BH : warp; r < BH; r += FILL_WARPS) {
    const int row = p * BH + r;
    const int rs = csr_v[row], re = csr_v[row + 1];
    const int base = cnt0 * BH + r * delta;
    int len;
    {
      int* h = hist[warp];
      for (int g = lane; g < delta; g += 32) h[g] = 0;
      __syncwarp();
      for (int e = rs + lane; e < re; e += 32) {
        int g = key2[e];
        g = g == SPARSE_KEY ? tp : g;
        atomicAdd(&h[g], 1);
      }
      __syncwarp();
      // exclusive scan of the group sizes -> group starts
      int carry = 0;
      for (int g0i = 0; g0i < delta; g0i += 32) {
        int g = g0i + lane;
        int v = g < delta ? h[g] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        int ex = carry + inc - v;
        if (g < delta) { h[g] = ex; mcsr_e[base + g] = rs + ex; }
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
      len = (re - rs) - h[tp];
      __syncwarp();
      // stable scatter, 32 nz at a time
      for (int e0 = rs; e0 < re; e0 += 32) {
        int e = e0 + lane;
        bool act = e < re;
        int g = 0xffff;
        uint32_t c = 0; float v = 0.f;
        if (act) { g = key2[e]; g = g == SPARSE_KEY ? tp : g; c = col[e]; v = val[e]; }
        unsigned peers = __match_any_sync(0xffffffffu, g);
        int rank = __popc(peers & ((1u << lane) - 1u));
        int pos = 0;
        if (act) pos = rs + h[g] + rank;
        __syncwarp();
        if (act && rank == 0) h[g] += __popc(peers);
        __syncwarp();
        if (act) { perm[pos] = e; csr_e[pos] = (int)c; csr_ev[pos] = v; }
      }
    }
    if (lane == 0) {
      spec_cnt[row] = len / STHRESHOLD;
      s1 += len; s2 += (long long)len * len; chunks += len / STHRESHOLD;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    chunks += __shfl_xor_sync(0xffffffffu, chunks, o);
  }
  if (lane == 0) {
    if (s1) atomicAdd(&stats[0], (unsigned long long)s1);
    if (s2) atomicAdd(&stats[1], (unsigned long long)s2);
    if (chunks) atomicAdd(&stats[2], (unsigned long long)chunks);
  }
  if (p == npanel - 1 && threadIdx.x == 0) mcsr_e[(size_t)BH * mcsr_cnt[npanel]] = ne;  // :1297
}

// no dense tile anywhere (:1224-1230): mcsr_cnt[p]=p, mcsr_e aliases csr_v, nz arrays alias the CSR
__global__ void k_nodense(const int* __restrict__ csr_v, int npanel, int nr, int* __restrict__ mcsr_cnt,
                          int* __restrict__ tcount, int* __restrict__ spec_cnt,
                          unsigned long long* __restrict__ stats) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= npanel) { mcsr_cnt[i] = i; tcount[i] = 0; }
  long long s1 = 0, s2 = 0;
  int chunks = 0;
  if (i < nr) {
    int len = csr_v[i + 1] - csr_v[i];
    spec_cnt[i] = len / STHRESHOLD;
    s1 = len; s2 = (long long)len * len; chunks = len / STHRESHOLD;
  }
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  s1 = cg::reduce(warp, s1, cg::plus<long long>());
  s2 = cg::reduce(warp, s2, cg::plus<long long>());
  chunks = cg::reduce(warp, chunks, cg::plus<int>());
  if ((threadIdx.x & 31) == 0) {
    if (s1) atomicAdd(&stats[0], (unsigned long long)s1);
    if (s2) atomicAdd(&stats[1], (unsigned long long)s2);
    if (chunks) atomicAdd(&stats[2], (unsigned long long)chunks);
  }
}

// ---- K5: special lists (make_special :1076-1087) in canonical (row-ascending) order ---------
__global__ void k_fill_special(const int* __restrict__ spec_cnt, const int* __restrict__ spec_off, int nr,
                               int* __restrict__ special, int* __restrict__ special2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr) return;
  int c = spec_cnt[i], o = spec_off[i];
  for (int j = 0; j < c; ++j) { special[o + j] = i; special2[o + j] = STHRESHOLD * j; }
}

// L2-residency scheduling of the 512-chunks (the idea of the reference's segment re-ordering, mat.cu:366
// dfsSegs / :527 sliWinSegs: run next to each other what reads the same rows of B).  A chunk is 512 consecutive nz of one
// long row, i.e. a contiguous range of COLUMNS; in row order, neighbouring CTAs of k_spmm_special_cta read unrelated ranges
// and every row of B comes from DRAM several times per SpMM.  The chunks are executed in the order of the first column they
// read instead (key below, radix-sorted), so the CTAs in flight at any time share one slice of B in L2.  partial[] stays
// indexed by chunk id: only the execution order changes, the result is bit-identical.
__global__ void k_special_keys(const int* __restrict__ special, const int* __restrict__ special2, const int* __restrict__ spec_off,
                               const int* __restrict__ mcsr_cnt, const int* __restrict__ mcsr_e, const int* __restrict__ csr_e,
                               int nr, int cap, unsigned* __restrict__ keys, int* __restrict__ iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  iota[i] = i;
  unsigned key = 0xFFFFFFFFu;
  if (i < spec_off[nr]) {
    const int row = special[i], off = special2[i];
    const int p = row / BH, r = row % BH;
    const int cnt0 = mcsr_cnt[p], delta = mcsr_cnt[p + 1] - cnt0;
    const int nch = spec_off[row + 1] - spec_off[row];
    const int lo = mcsr_e[cnt0 * BH + (r + 1) * delta] - nch * STHRESHOLD + off;
    key = (unsigned)csr_e[lo];
  }
  keys[i] = key;
}

}  // namespace

namespace fx {

static int sm_count() {
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  return sm;
}

static int heavy_ctas(int npanel) {
  const int sm = sm_count();
  int g = build_threads() > 512 ? sm : 2 * sm;  // one 1024-thread CTA per SM (fx_common.cuh:build_threads), or two of 512
  // FLEX_BUILD_CTAS: persistent CTAs of the counting kernels.  Every CTA owns one counter per column (4 B x ncols): 138 MB on
  // Reddit-shape with one CTA per SM (276 MB with two, as in round 1)
  static const int g_env = getenv("FLEX_BUILD_CTAS") ? atoi(getenv("FLEX_BUILD_CTAS")) : 0;
  if (g_env > 0) g = g_env;
  return npanel < g ? (npanel > 0 ? npanel : 1) : g;
}

int aspt_carve(fx_tiles* t, int64_t ncols, size_t extra_bytes) {
  fx_aspt_dev& a = t->aspt;
  const int64_t ne = a.ne, nr = a.nr, npanel = a.npanel;
  const int BW = a.BW, min_occ = BW * 3 / 4;
  a.G = heavy_ctas((int)npanel);
  a.list_cap_tiles = (int)(ne / ((int64_t)min_occ * THRESHOLD) + 1);
  a.mcsr_e_cap = (int)((int64_t)BH * (a.list_cap_tiles + npanel) + 2);
  a.special_cap = (int)(ne / STHRESHOLD + 1);
  a.partial_cap_floats = (size_t)a.special_cap * (size_t)t->k;
  size_t bytes = 0;
  auto add = [&](size_t b) { bytes += Arena::pad(b) + 256; };
  add(sizeof(int) * (nr + 2));                       // csr_v
  for (int i = 0; i < 6; ++i) add(sizeof(int) * (npanel + 2));  // chk, cnt, tcount, nheavy, 2 panel lists
  add(sizeof(uint16_t) * (ne + 2));                  // key2
  add(sizeof(int2) * (ne / THRESHOLD + npanel + 2));  // heavy
  add(sizeof(unsigned) * (size_t)a.G * ncols);       // counters
  add(sizeof(int) * (size_t)a.mcsr_e_cap);
  add(sizeof(int) * (size_t)a.list_cap_tiles * BW);
  add(sizeof(int) * (size_t)a.list_cap_tiles * 2);
  add(sizeof(int) * (ne + 2) * 2);                   // perm, csr_e
  add(sizeof(float) * (ne + 2));                     // csr_ev
  add(sizeof(int) * (nr + 2) * 2);                   // spec_cnt, spec_off
  add(sizeof(int) * (size_t)a.special_cap * 2);
  add(sizeof(int) * (size_t)a.special_cap * 4);  // spec_order, spec_iota, spec_keys, spec_keys_out
  cub::DeviceRadixSort::SortPairs(nullptr, a.spec_sort_tmp_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, a.special_cap);
  a.spec_sort_tmp_bytes = std::max(a.spec_sort_tmp_bytes, device_scan_tmp_bytes((int)nr + 1));  // shared with the row scans
  add(a.spec_sort_tmp_bytes);
  add(sizeof(unsigned long long) * 16);
  add(sizeof(float) * a.partial_cap_floats);
  a.wl_cap = (int)(npanel * 2 + ne / 4096 + 32);
  for (int i = 0; i < 3; ++i) add(sizeof(int2) * (size_t)a.wl_cap);
  int rc = t->arena.reserve(bytes + extra_bytes);
  if (rc != FX_OK) return rc;
  Arena& A = t->arena;
  a.csr_v = A.take<int>(nr + 2);
  a.mcsr_chk = A.take<int>(npanel + 2);
  a.mcsr_cnt = A.take<int>(npanel + 2);
  a.tcount = A.take<int>(npanel + 2);
  a.nheavy = A.take<int>(npanel + 2);
  a.plist_plain = A.take<int>(npanel + 2);
  a.plist_tiled = A.take<int>(npanel + 2);
  a.key2 = A.take<uint16_t>(ne + 2);
  a.heavy = A.take<int2>(ne / THRESHOLD + npanel + 2);
  a.cnt_scratch = A.take<unsigned>((size_t)a.G * ncols);
  a.mcsr_e = A.take<int>(a.mcsr_e_cap);
  a.mcsr_list = A.take<int>((size_t)a.list_cap_tiles * BW);
  a.baddr = A.take<int>(a.list_cap_tiles);
  a.saddr = A.take<int>(a.list_cap_tiles);
  a.perm = A.take<int>(ne + 2);
  a.csr_e = A.take<int>(ne + 2);
  a.csr_ev = A.take<float>(ne + 2);
  a.spec_cnt = A.take<int>(nr + 2);
  a.spec_off = A.take<int>(nr + 2);
  a.special = A.take<int>(a.special_cap);
  a.special2 = A.take<int>(a.special_cap);
  a.spec_order = A.take<int>(a.special_cap);
  a.spec_iota = A.take<int>(a.special_cap);
  a.spec_keys = A.take<unsigned>(a.special_cap);
  a.spec_keys_out = A.take<unsigned>(a.special_cap);
  a.spec_sort_tmp = A.take<char>(a.spec_sort_tmp_bytes);
  a.stats = A.take<unsigned long long>(16);
  a.wl_all = A.take<int2>(a.wl_cap);
  a.wl_plain = A.take<int2>(a.wl_cap);
  a.wl_tiled = A.take<int2>(a.wl_cap);
  a.partial = A.take<float>(a.partial_cap_floats);
  if (!a.partial || !a.stats) { set_error("arena carve overflow"); return FX_ERR_NOMEM; }
  // the column counters must start at zero; every build leaves them zeroed again
  FX_CUDA(cudaMemset(a.cnt_scratch, 0, sizeof(unsigned) * (size_t)a.G * ncols));
  return FX_OK;
}

// Everything between the two events of fx_build: kernels + one 64-byte D2H.
int aspt_build(fx_tiles* t, cudaStream_t s) {
  fx_aspt_dev& a = t->aspt;
  const fx_matrix* m = t->mat;
  const int BW = a.BW, min_occ = BW * 3 / 4;
  const uint32_t* col = t->src_rowptr ? t->src_col : m->col_dev + m->rowptr[t->row_begin];
  const float* val = t->src_rowptr ? t->src_val : m->val_dev + m->rowptr[t->row_begin];
  const uint32_t* src_rowptr = t->src_rowptr ? t->src_rowptr : m->rowptr_dev;
  const int src_row0 = t->src_rowptr ? t->src_row0 : t->row_begin;
  const int nloc = t->row_end - t->row_begin;
  FX_CUDA(cudaMemsetAsync(a.stats, 0, sizeof(unsigned long long) * 16, s));
  k_pad_rowptr<<<ceil_div(a.nr + 1, 256), 256, 0, s>>>(src_rowptr, src_row0, nloc, a.nr, a.ne, a.csr_v);
  FX_LAUNCH_CHECK();
  k_detect<<<a.npanel, build_threads(), 0, s>>>(a.csr_v, col, min_occ, a.mcsr_chk, a.stats);
  FX_LAUNCH_CHECK();
  // the one decision the host has to take (reference: d_flag copy :1217-1224)
  FX_CUDA(cudaMemcpyAsync(t->stats_host, a.stats, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
  FX_CUDA(cudaStreamSynchronize(s));
  a.any_flag = t->stats_host[4] != 0;
  int* mcsr_e = a.mcsr_e;
  if (!a.any_flag) {
    a.aliased = true;
    a.mcsr_e_use = a.csr_v;
    a.csr_e_use = reinterpret_cast<const int*>(col);
    a.csr_ev_use = val;
    k_nodense<<<ceil_div(a.nr + 1, 256), 256, 0, s>>>(a.csr_v, a.npanel, a.nr, a.mcsr_cnt, a.tcount, a.spec_cnt,
                                                         a.stats);
    FX_LAUNCH_CHECK();
  } else {
    a.aliased = false;
    a.mcsr_e_use = a.mcsr_e;
    a.csr_e_use = a.csr_e;
    a.csr_ev_use = a.csr_ev;
    FX_CUDA(cudaMemsetAsync(a.tcount, 0, sizeof(int) * (a.npanel + 1), s));
    static SmemAttr heavy_attr;
    if (int rc = heavy_attr.ensure(k_heavy, SORT_SMEM_CAP * sizeof(unsigned long long))) return rc;
    k_heavy<<<a.G, build_threads(), SORT_SMEM_CAP * sizeof(unsigned long long), s>>>(
        a.csr_v, col, a.mcsr_chk, a.npanel, BW, min_occ, (int)m->n, a.cnt_scratch,
        reinterpret_cast<unsigned long long*>(a.heavy), a.nheavy, a.tcount, a.key2);
    FX_LAUNCH_CHECK();
    k_scan_tcount<<<1, 1024, 0, s>>>(a.tcount, a.npanel, a.mcsr_cnt, a.stats);
    FX_LAUNCH_CHECK();
    k_panel_lists<<<1, 1024, 0, s>>>(a.tcount, a.npanel, a.plist_plain, a.plist_tiled, a.stats);
    FX_LAUNCH_CHECK();
    k_fill<<<a.npanel, FILL_WARPS * 32, 0, s>>>(a.csr_v, col, val, a.tcount, a.mcsr_cnt, a.heavy, a.nheavy, a.key2,
                                                a.npanel, BW, a.ne, mcsr_e, a.mcsr_list, a.baddr, a.saddr, a.perm,
                                                a.csr_e, a.csr_ev, a.spec_cnt, a.stats);
    FX_LAUNCH_CHECK();
  }
  FX_CUDA(device_exclusive_scan(a.spec_sort_tmp, a.spec_sort_tmp_bytes, a.spec_cnt, a.nr, a.spec_off, s));
  k_fill_special<<<ceil_div(a.nr, 256), 256, 0, s>>>(a.spec_cnt, a.spec_off, a.nr, a.special, a.special2);
  FX_LAUNCH_CHECK();
  {  // execution order of the chunks: by the first column they read
    k_special_keys<<<ceil_div(a.special_cap, 256), 256, 0, s>>>(a.special, a.special2, a.spec_off, a.mcsr_cnt, a.mcsr_e_use, a.csr_e_use,
                                                              a.nr, a.special_cap, a.spec_keys, a.spec_iota);
    FX_LAUNCH_CHECK();
    size_t tmp = a.spec_sort_tmp_bytes;
    FX_CUDA(cub::DeviceRadixSort::SortPairs(a.spec_sort_tmp, tmp, a.spec_keys, a.spec_keys_out, a.spec_iota, a.spec_order, a.special_cap,
                                            0, 32, s));  // padding keys (0xFFFFFFFF) sort to the end
  }
  // Heaviest entries first only when the grid is a few waves long (at most 6 panels per SM; the row kernel runs 3 CTAs per SM): there the
  // tail is what counts (flickr-shape, 698 panels: 0.087 -> 0.073 ms); on long grids the panel order is worth more, because
  // neighbouring panels share B rows in L2 (yelp-shape 0.576 -> 0.593 ms, Amazon-shape 6.55 -> 7.34 ms; Reddit-shape equal).
  // (tried on long grids: cutting the panels of the last wave in four and of the wave before in two, to shorten the tail
  // (12 % of the row kernel's SM time on Reddit-shape) -- each part repeats the panel prologue and it only costs:
  // Reddit-shape 0.536 -> 0.577 ms, yelp-shape 0.577 -> 0.614, Amazon-shape 6.58 -> 6.76)
  static const int wl_env = getenv("FLEX_WL_CLASSES") ? std::max(1, std::min(4, atoi(getenv("FLEX_WL_CLASSES")))) : 0;
  const int wl_classes = wl_env ? wl_env : (a.npanel <= 6 * sm_count() ? 4 : 1);
  k_worklist<<<1, 1024, 0, s>>>(nullptr, 0, a.csr_v, a.spec_off, a.nr, a.npanel, a.wl_all, a.wl_cap, a.stats, a.stats + 8, wl_classes);
  FX_LAUNCH_CHECK();
  if (a.any_flag) {  // the tiled launch and the plain launch walk their own panel lists (counts are on the device)
    k_worklist<<<1, 1024, 0, s>>>(a.plist_plain, 1, a.csr_v, a.spec_off, a.nr, a.npanel, a.wl_plain, a.wl_cap, a.stats, a.stats + 9, wl_classes);
    FX_LAUNCH_CHECK();
    k_worklist<<<1, 1024, 0, s>>>(a.plist_tiled, 2, a.csr_v, a.spec_off, a.nr, a.npanel, a.wl_tiled, a.wl_cap, a.stats, a.stats + 10, wl_classes);
    FX_LAUNCH_CHECK();
  }
  FX_CUDA(cudaMemcpyAsync(t->stats_host, a.stats, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost, s));
  FX_CUDA(cudaStreamSynchronize(s));
  a.S1 = (long long)t->stats_host[0];
  a.S2 = (long long)t->stats_host[1];
  a.special_p = (int)t->stats_host[2];
  a.num_dense = a.any_flag ? (int)t->stats_host[3] : 0;
  a.max_tp = a.any_flag ? (int)t->stats_host[5] : 0;
  a.n_tiled = a.any_flag ? (int)t->stats_host[6] : 0;
  a.n_plain = a.npanel - a.n_tiled;
  a.n_wl_all = (int)t->stats_host[8];
  a.n_wl_plain = a.any_flag ? (int)t->stats_host[9] : 0;
  a.n_wl_tiled = a.any_flag ? (int)t->stats_host[10] : 0;
  a.avg = a.nr ? (double)a.S1 / a.nr : 0;                       // :1226 / :1300
  a.vari = a.nr ? (double)a.S2 / a.nr - a.avg * a.avg : 0;      // Σ(len-avg)²/nr, exact sums
  const int nc = a.n;
  a.regime = (nc > 0 && a.ne / nc < 6 && a.vari < 40) ? 0 : (a.vari < 200 ? 1 : 2);  // :1355,1368,1381
  return FX_OK;
}

}  // namespace fx
