// fx_flex_build.cu -- builders of the Flex formats (mat.cu) and the SpMM kernels that consume them.
//
//   FX_FMT_TILE   csr2flex_Rmajor / csr2flex_Cmajor   mat.cu:1345-1518   GPU builder
//   FX_FMT_SEG    csr2seg_Cmajor + SM buckets         mat.cu:1192-1269, 1118-1162   GPU builder
//   FX_FMT_PILLAR csr2_DiagTiling                     mat.cu:680-903    host round 1 (serial chain) + GPU rounds 2-3
//
// GPU builders: one cooperative tile of TM lanes per row panel, lane i owning row i's cursor; a step
// of the reference's sequential sweep (next flexible-origin tile / next distinct column) becomes a
// min-reduction + ballot over the TM lanes.  Two passes (count, then write at scanned offsets), no
// host round trip except the totals.  Layouts are bit-identical to the reference's.
//
// SpMM: k_spmm_panel_acc streams a panel's nz in the stored order and accumulates the panel's tm rows
// of C in shared memory (one store per C element, deterministic -- the reference's v10-v35 kernels
// use fp32 atomics for rows split across segments).  k_spmm_alpha is the pillar-format kernel
// (alpha_w_atomic_spmm_v36, flex.cu:4010-4124): SM-affine pillar queues + shared balance queue,
// store or atomicAdd by the MSB of segVoMap.
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <numeric>

#include "fx_common.cuh"
#include "fx_flex.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr unsigned NOCOL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// F2/F3: row-panel segmentation
// ---------------------------------------------------------------------------------------------
struct SegOut {
  unsigned *alpha_rowPtr, *alpha_colIdx, *pillar_rowPtr, *segVoMap;
  float* alpha_vals;
  unsigned *segPtr, *segNzRCIdx, *segVoMapPad;
  float *segVals, *segNzCV;
  int* seg_rowPtr;
};

template <int TM, bool WRITE>
__global__ void __launch_bounds__(128) k_seg(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col,
                                             const float* __restrict__ val, const int* __restrict__ vo_mp, int m,
                                             int npanels, int nnz_limit, int* __restrict__ segs_per_panel,
                                             const int* __restrict__ seg_off, SegOut o) {
  auto tile = cg::tiled_partition<TM>(cg::this_thread_block());
  const int lane = tile.thread_rank();
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) / TM;
  if (p >= npanels) return;
  const int rowStart = p * TM, rowEnd = min(m, rowStart + TM), rows = rowEnd - rowStart;
  const int dif = (int)(0.1 * nnz_limit);
  const bool has_row = lane < rows;
  const int row = rowStart + lane;
  const unsigned rs = has_row ? rowptr[row] : 0, re = has_row ? rowptr[row + 1] : 0;
  unsigned cur = rs, prev = rs;
  const unsigned panel_base = rowptr[rowStart];
  int emitted = 0, nnzInSeg = 0, nseg = 0, atom = 0;
  int remaining = (int)(rowptr[rowEnd] - panel_base);
  const int s0 = WRITE ? seg_off[p] : 0;
  // the lane's current (column, value) and the pair after it: the next pair is requested when the current one is taken.
  // (tried: the lane's next 32 columns in shared memory, refilled 32 loads at a time, and a bulk step when a single row is
  // left -- both slower: a step of the sweep is a ~350-clk chain of dependent shuffles and votes, not a wait for a load, and
  // the rows of a panel all reach to the last columns, so a hub row is never "alone".  The sweep of the panel that holds the
  // longest row IS the kernel's duration: 18 k steps on Reddit-shape.)
  unsigned c = cur < re ? col[cur] : NOCOL, cn = cur + 1 < re ? col[cur + 1] : NOCOL;
  float v = (WRITE && cur < re) ? val[cur] : 0.f, vn = (WRITE && cur + 1 < re) ? val[cur + 1] : 0.f;
  while (remaining > 0 || nnzInSeg > 0) {
    if (remaining > 0) {
      const unsigned j = cg::reduce(tile, c, cg::less<unsigned>());
      const bool take = c == j;  // a lane without nz holds NOCOL, never the minimum while nz remain
      const unsigned bal = tile.ballot(take);
      if (take) {
        if (WRITE) {  // C-major stream: position in consumption order
          const unsigned pos = panel_base + emitted + nnzInSeg + __popc(bal & ((1u << lane) - 1u));
          o.segNzRCIdx[2 * (size_t)pos] = lane;
          o.segNzRCIdx[2 * (size_t)pos + 1] = c;
          o.segVals[pos] = v;
        }
        ++cur; ++atom;
        c = cn; v = vn;
        cn = cur + 1 < re ? col[cur + 1] : NOCOL;
        if (WRITE) vn = cur + 1 < re ? val[cur + 1] : 0.f;
      }
      const int took = __popc(bal);
      nnzInSeg += took; remaining -= took;
    }
    if ((remaining == 0 && nnzInSeg) || (nnz_limit - nnzInSeg) <= dif || nnzInSeg > nnz_limit) {  // mat.cu:1235
      if (WRITE) {
        const int s = s0 + nseg;
        const int cnt = (int)(cur - prev);
        const int ex = cg::exclusive_scan(tile, cnt);
        const unsigned seg_base = panel_base + emitted;
        const size_t rbase = (size_t)s0 * TM + (size_t)nseg * rows;  // every earlier panel has TM rows
        if (has_row) {
          o.alpha_rowPtr[rbase + lane] = seg_base + ex;
          for (int q = 0; q < cnt; ++q) {
            const unsigned cc = col[prev + q];
            const float vv = val[prev + q];
            o.alpha_colIdx[seg_base + ex + q] = cc;
            o.alpha_vals[seg_base + ex + q] = vv;
            o.segNzCV[2 * (size_t)(seg_base + ex + q)] = (float)cc;
            o.segNzCV[2 * (size_t)(seg_base + ex + q) + 1] = vv;
          }
          const unsigned v = (unsigned)vo_mp[row];
          o.segVoMap[rbase + lane] = atom < (int)(re - rs) ? (v | 0x80000000u) : v;  // mat.cu:1252-1260
        }
        o.segVoMapPad[(size_t)s * TM + lane] = has_row ? ((unsigned)vo_mp[row] | (atom < (int)(re - rs) ? 0x80000000u : 0u))
                                                       : 0x7fffffffu;
        o.seg_rowPtr[(size_t)s * (TM + 1) + lane] = has_row ? ex : nnzInSeg;
        if (lane == 0) {
          o.seg_rowPtr[(size_t)s * (TM + 1) + TM] = nnzInSeg;
          o.pillar_rowPtr[s] = (unsigned)rbase;
          o.segPtr[s] = seg_base;
        }
      }
      emitted += nnzInSeg;
      nnzInSeg = 0; atom = 0; prev = cur;
      ++nseg;
    }
  }
  if (!WRITE && lane == 0) segs_per_panel[p] = nseg;
}

// exclusive scan of a small int array by one CTA (npanels up to ~1M): out[n] = total
__global__ void __launch_bounds__(1024) k_scan_int(const int* __restrict__ in, int n, int* __restrict__ out) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  auto warp = cg::tiled_partition<32>(cg::this_thread_block());
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
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
    if (i < n) out[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

__global__ void k_seg_tail(SegOut o, int nsegs, int rows_total, unsigned nnz) {
  o.alpha_rowPtr[rows_total] = nnz;
  o.pillar_rowPtr[nsegs] = (unsigned)rows_total;
  o.segPtr[nsegs] = nnz;
}

// ---------------------------------------------------------------------------------------------
// F5: pillar format, rounds 2 and 3 of csr2_DiagTiling (mat.cu:771-903)
// ---------------------------------------------------------------------------------------------
// Round 1 (host, fx_flex_host.cu) leaves the diagonal blocks (pstart) and one byte per column (listed).  Round 2 gives every row
// the nz whose column lies in the column window of the row's SM (the rows of its 64 blocks) and is listed -- independent per
// row: one warp per row counts, a scan places, one warp per row writes.  Round 3 cuts what is left into column-major
// segments per tm-row panel, the k_seg sweep above over a filtered row (a nz is "claimed" iff round 2 took it, so the
// reference's per-row hash sets reduce to re-evaluating the round-2 predicate).
struct PillarIn {
  const unsigned *rowptr, *col;
  const float* val;
  const int* vo_mp;
  const unsigned char* listed;
  const int* pstart;  // [wpw+1] strictly increasing, pstart[wpw] == m
  int m, wpw;
};

struct PillarOut {
  unsigned *alpha_rowPtr, *alpha_colIdx, *pillar_rowPtr, *segVoMap;
  float* alpha_vals;
  unsigned* rest_col;  // the nz round 2 did not take, row by row: row r at [rowptr[r] - alpha_rowPtr[r], rowptr[r+1] - alpha_rowPtr[r+1])
  float* rest_val;
};

// column window of the SM that owns `row` (mat.cu:773-779): the rows of the group of 64 blocks holding the row's block
__device__ __forceinline__ void sm_window(const PillarIn& a, int row, unsigned& cs, unsigned& ce) {
  int lo = 0, hi = a.wpw;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.pstart[mid] <= row) lo = mid; else hi = mid;
  }
  const int g = lo & ~63;
  cs = (unsigned)a.pstart[g];
  ce = (unsigned)a.pstart[min(g + 64, a.wpw)];
}

__device__ __forceinline__ bool r2_takes(const PillarIn& a, unsigned c, unsigned cs, unsigned ce) {
  return c >= cs && c < ce && a.listed[c];
}

__global__ void __launch_bounds__(256) k_pillar_count(PillarIn a, PillarOut o) {
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row > a.m) return;
  if (row == a.m) { if (lane == 0) o.alpha_rowPtr[a.m] = 0; return; }
  unsigned cs, ce;
  sm_window(a, row, cs, ce);
  const unsigned rs = a.rowptr[row], re = a.rowptr[row + 1];
  int cnt = 0;
  for (unsigned e = rs + lane; e < re; e += 32) cnt += r2_takes(a, a.col[e], cs, ce);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) {
    o.alpha_rowPtr[row] = (unsigned)cnt;  // scanned in place
    const unsigned v = (unsigned)a.vo_mp[row];
    o.segVoMap[row] = cnt < (int)(re - rs) ? (v | 0x80000000u) : v;  // mat.cu:826-833
  }
}

__global__ void __launch_bounds__(256) k_pillar_fill(PillarIn a, PillarOut o, int* __restrict__ err) {
  const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (gt < (size_t)a.wpw) {  // assert( nnz_p_warp[i] ), mat.cu:835
    if (o.alpha_rowPtr[a.pstart[gt + 1]] == o.alpha_rowPtr[a.pstart[gt]]) atomicExch(err, 1 + (int)gt);
  }
  const int row = (int)(gt >> 5), lane = threadIdx.x & 31;
  if (row >= a.m) return;
  unsigned cs, ce;
  sm_window(a, row, cs, ce);
  const unsigned rs = a.rowptr[row], re = a.rowptr[row + 1];
  unsigned base = o.alpha_rowPtr[row], rbase = rs - base;
  for (unsigned e0 = rs; e0 < re; e0 += 32) {
    const unsigned e = e0 + lane;
    unsigned c = 0;
    bool t = false;
    if (e < re) { c = a.col[e]; t = r2_takes(a, c, cs, ce); }
    const unsigned bal = __ballot_sync(0xffffffffu, t), balr = __ballot_sync(0xffffffffu, e < re && !t);
    if (t) {
      const unsigned pos = base + __popc(bal & ((1u << lane) - 1u));
      o.alpha_colIdx[pos] = c;
      o.alpha_vals[pos] = a.val[e];
    } else if (e < re) {  // left for round 3: compacted, so that its sweep reads a plain CSR
      const unsigned pos = rbase + __popc(balr & ((1u << lane) - 1u));
      o.rest_col[pos] = c;
      o.rest_val[pos] = a.val[e];
    }
    base += __popc(bal);
    rbase += __popc(balr);
  }
}

// Round 1 on the GPU, for the matrices every BASELINE shape is (diagonal stored in every row, structure symmetric near the
// diagonal).  Round 1 (mat.cu:706-759) grows block i from its first row s until the nz inside the square [s, j) x [s, j) exceed
// thr: with the diagonal present the walk of row j (mat.cu:718-727) counts the row's nz with s <= col <= j, and the column
// sweep (mat.cu:729-743) counts the rows kk in [s, j) that hold column j -- by symmetry the row's nz with s <= col < j.  So the
// end e(s) of a block depends on s alone and is computed for EVERY s in parallel (a thread per s, at most thr + 1 rows each:
// every row adds at least its diagonal); the chain s -> e(s) over the 64 n_sm blocks is then followed on the host over one
// downloaded array.  Anything else (a row without its diagonal, an nz within thr + 1 of the diagonal without its transpose)
// goes to the host's round 1 (fx_flex_host.cu), which follows the reference statement by statement.
__global__ void k_diag_pos(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col, int m, int* __restrict__ dpos,
                           int* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  unsigned lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (col[mid] < (unsigned)r) lo = mid + 1; else hi = mid; }
  const bool ok = lo < rowptr[r + 1] && col[lo] == (unsigned)r;
  dpos[r] = ok ? (int)lo : -1;
  if (!ok) atomicExch(&flags[1], 1);
}

__device__ __forceinline__ bool row_has(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col, unsigned r, unsigned c) {
  unsigned lo = rowptr[r], hi = rowptr[r + 1];
  const unsigned end = hi;
  while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (col[mid] < c) lo = mid + 1; else hi = mid; }
  return lo < end && col[lo] == c;
}

// every nz (r, c) with 0 < |r - c| <= band has its transpose (only these can lie inside a diagonal square)
__global__ void k_band_sym(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col, const int* __restrict__ dpos, int m,
                           int band, int* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m || dpos[r] < 0) return;
  const int d = dpos[r];
  for (int e = d - 1; e >= (int)rowptr[r] && (int)col[e] >= r - band; --e)
    if (!row_has(rowptr, col, col[e], (unsigned)r)) { atomicExch(&flags[2], 1); return; }
  for (int e = d + 1; e < (int)rowptr[r + 1] && (int)col[e] <= r + band; ++e)
    if (!row_has(rowptr, col, col[e], (unsigned)r)) { atomicExch(&flags[2], 1); return; }
}

__global__ void k_diag_ends(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col, const int* __restrict__ dpos, int m,
                            int thr, int* __restrict__ e_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  int cnt = 0, j = s;
  while (j < m && cnt <= thr) {  // mat.cu:716
    unsigned lo = rowptr[j], hi = (unsigned)dpos[j];
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (col[mid] < (unsigned)s) lo = mid + 1; else hi = mid; }
    const int a = dpos[j] - (int)lo + 1;  // nz of row j with s <= col <= j
    cnt += 2 * a - 1;                     // + the rows of [s, j) that hold column j
    ++j;
  }
  e_out[s] = j;
}

// listed[c] = 1 iff some nz (r, c) has r and c in the same diagonal block (mat.cu:722,737)
__global__ void __launch_bounds__(256) k_diag_listed(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col,
                                                     const int* __restrict__ pstart, int wpw, int m, unsigned char* __restrict__ listed) {
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  int lo = 0, hi = wpw;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pstart[mid] <= row) lo = mid; else hi = mid; }
  const unsigned cs = (unsigned)pstart[lo], ce = (unsigned)pstart[lo + 1];
  for (unsigned e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) {
    const unsigned c = col[e];
    if (c >= cs && c < ce) listed[c] = 1;
  }
}

// csr2seg_Cmajor (mat.cu:1192-1269) over the nz round 2 left, nnz_limit = 128 (mat.cu:875 passes the member default)
template <int TM, bool WRITE>
__global__ void __launch_bounds__(128) k_pseg(PillarIn a, int npanels, int nnz_limit, int* __restrict__ segs_per_panel,
                                              int* __restrict__ nz_per_panel, const int* __restrict__ seg_off,
                                              const int* __restrict__ nz_off, PillarOut o) {
  auto tile = cg::tiled_partition<TM>(cg::this_thread_block());
  const int lane = tile.thread_rank();
  const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / TM);
  if (p >= npanels) return;
  const int rowStart = p * TM, rowEnd = min(a.m, rowStart + TM), rows = rowEnd - rowStart;
  const int dif = (int)(0.1 * nnz_limit);
  const bool has_row = lane < rows;
  const int row = rowStart + lane;
  // the row's remainder after round 2, compacted by k_pillar_fill; `full` = the whole row's length (MSB rule, mat.cu:1252)
  unsigned rs = 0, re = 0;
  int full = 0;
  if (has_row) {
    const unsigned g0 = a.rowptr[row], g1 = a.rowptr[row + 1];
    rs = g0 - o.alpha_rowPtr[row];
    re = g1 - o.alpha_rowPtr[row + 1];
    full = (int)(g1 - g0);
  }
  const unsigned* __restrict__ rcol = o.rest_col;
  const float* __restrict__ rval = o.rest_val;
  unsigned cur = rs, prev = rs;
  // the lane's current column and the one after it (see k_seg)
  unsigned c = cur < re ? rcol[cur] : NOCOL;
  unsigned cn = cur + 1 < re ? rcol[cur + 1] : NOCOL;
  int emitted = 0, nnzInSeg = 0, nseg = 0, atom = 0;
  const unsigned r2 = WRITE ? o.alpha_rowPtr[a.m] : 0;
  const int s0 = WRITE ? seg_off[p] : 0;
  const unsigned nz0 = WRITE ? r2 + (unsigned)nz_off[p] : 0;
  bool left = tile.any(cur < re);
  while (left || nnzInSeg > 0) {
    if (left) {
      const unsigned j = cg::reduce(tile, c, cg::less<unsigned>());
      const bool take = c == j;  // a lane without nz holds NOCOL, which is never the minimum while any nz is left
      if (take) {
        ++cur; ++atom;
        c = cn;
        cn = cur + 1 < re ? rcol[cur + 1] : NOCOL;
      }
      nnzInSeg += __popc(tile.ballot(take));
      left = tile.any(cur < re);
    }
    if ((!left && nnzInSeg) || (nnz_limit - nnzInSeg) <= dif || nnzInSeg > nnz_limit) {  // mat.cu:1235
      if (WRITE) {
        const int ex = cg::exclusive_scan(tile, atom);
        const unsigned seg_base = nz0 + emitted;
        const size_t rbase = (size_t)s0 * TM + (size_t)nseg * rows;  // every earlier panel has TM rows
        if (has_row) {
          o.alpha_rowPtr[a.m + 1 + rbase + lane] = seg_base + ex + atom;  // end of the virtual row; its start is the entry before
          unsigned w = seg_base + ex;
          for (unsigned e = prev; e < cur; ++e, ++w) {
            o.alpha_colIdx[w] = rcol[e];
            o.alpha_vals[w] = rval[e];
          }
          const unsigned v = (unsigned)a.vo_mp[row];
          o.segVoMap[a.m + rbase + lane] = atom < full ? (v | 0x80000000u) : v;  // mat.cu:1252-1260
        }
        if (lane == 0) o.pillar_rowPtr[a.wpw + s0 + nseg] = (unsigned)(a.m + rbase);
      }
      emitted += nnzInSeg;
      nnzInSeg = 0; atom = 0; prev = cur;
      ++nseg;
    }
  }
  if (!WRITE && lane == 0) { segs_per_panel[p] = nseg; nz_per_panel[p] = emitted; }
  if (!WRITE && p == npanels - 1 && lane == 0) { segs_per_panel[npanels] = 0; nz_per_panel[npanels] = 0; }
}

__global__ void k_pillar_tail(PillarOut o, int wpw, int nsegs3, unsigned rows_total) { o.pillar_rowPtr[wpw + nsegs3] = rows_total; }

// ---------------------------------------------------------------------------------------------
// F1: flexible-origin tiles
// ---------------------------------------------------------------------------------------------
struct TileOut {
  unsigned *tileRowPtr, *tileNnz, *tileColIdx;
  int *nnzTile, *bitMap, *rcOffset;
  float* newVals;
};

template <int TM, bool WRITE>
__global__ void __launch_bounds__(128) k_tile(const unsigned* __restrict__ rowptr, const unsigned* __restrict__ col,
                                              const float* __restrict__ val, int m, int n, int npanels, int tn,
                                              int cmajor, int* __restrict__ tiles_per_panel,
                                              const int* __restrict__ tile_off, TileOut o) {
  auto tile = cg::tiled_partition<TM>(cg::this_thread_block());
  const int lane = tile.thread_rank();
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) / TM;
  if (p >= npanels) return;
  const int rowStart = p * TM, rowEnd = min(m, rowStart + TM), rows = rowEnd - rowStart;
  const bool has_row = lane < rows;
  const unsigned rs = has_row ? rowptr[rowStart + lane] : 0, re = has_row ? rowptr[rowStart + lane + 1] : 0;
  unsigned cur = rs;
  unsigned pos = rowptr[rowStart];
  const unsigned pend = rowptr[rowEnd];
  int nt = 0;
  const int t0 = WRITE ? tile_off[p] : 0;
  while (pos < pend) {
    const unsigned c0 = cur < re ? col[cur] : NOCOL;
    const unsigned left = cg::reduce(tile, c0, cg::less<unsigned>());
    const unsigned right = min(left + (unsigned)tn, (unsigned)n);
    // this row's entries inside [left,right)
    unsigned e = cur;
    unsigned bits = 0;
    while (e < re && col[e] < right) { bits |= 1u << (col[e] - left); ++e; }
    const int cnt = (int)(e - cur);
    const int total = cg::reduce(tile, cnt, cg::plus<int>());
    if (WRITE) {
      unsigned bm = bits;
      for (int o2 = TM / 2; o2 > 0; o2 >>= 1) bm |= tile.shfl_xor(bm, o2);
      if (!cmajor) {
        const int ex = cg::exclusive_scan(tile, cnt);
        for (int q = 0; q < cnt; ++q) {
          o.rcOffset[pos + ex + q] = (lane << 16) | (int)(col[cur + q] - left);
          o.newVals[pos + ex + q] = val[cur + q];
        }
      } else {
        // (column, row) order: entries of earlier columns first, then earlier rows of the same column
        unsigned before = 0;  // running count of entries in columns < current
        unsigned q = cur;
        for (int ci = 0; ci < tn; ++ci) {
          const bool mine = q < e && col[q] == left + ci;
          const unsigned bal = tile.ballot(mine);
          if (mine) {
            const unsigned w = pos + before + __popc(bal & ((1u << lane) - 1u));
            o.rcOffset[w] = (lane << 16) | ci;
            o.newVals[w] = val[q];
            ++q;
          }
          before += __popc(bal);
        }
      }
      if (lane == 0) {
        const int t = t0 + nt;
        o.nnzTile[t] = total;
        o.bitMap[t] = (int)bm;
        o.tileNnz[t] = pos;  // prefix: tileNnz[t] = nz before tile t
        o.tileColIdx[t] = left;
      }
    }
    cur = e;
    pos += total;
    ++nt;
  }
  if (!WRITE && lane == 0) tiles_per_panel[p] = nt;
}

__global__ void k_tile_tail(TileOut o, const int* __restrict__ tile_off, int npanels, unsigned nnz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= npanels) o.tileRowPtr[i] = (unsigned)tile_off[i];
  if (i == 0) o.tileNnz[tile_off[npanels]] = nnz;
}

// ---------------------------------------------------------------------------------------------
// SpMM over a panel-ordered nz stream with shared-memory accumulators (tile + seg formats)
// ---------------------------------------------------------------------------------------------
// MODE 0: tile format  (tileRowPtr/tileNnz/tileColIdx/rcOffset/newVals)
// MODE 1: seg format   (per panel: segments seg_off[p]..seg_off[p+1], stream segNzRCIdx/segVals)
struct AccArgs {
  const unsigned *tileRowPtr, *tileNnz, *tileColIdx;
  const int* rcOffset;
  const float* vals;
  const int* seg_off;
  const unsigned *segPtr, *segNzRCIdx;
  const int* vo_mp;  // output row = vo_mp ? vo_mp[row] : row   (C in ORIGINAL vertex order, flex.cu:994)
  const float* B;
  float* C;
  int m, npanels, k, tm;
};

template <int KC, int MODE>
__global__ void __launch_bounds__(256) k_spmm_panel_acc(AccArgs a) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR, NWK = 8 * RPW;
  extern __shared__ __align__(16) float sacc[];  // [NWK][tm][KC]
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int wk = (threadIdx.x >> 5) * RPW + (threadIdx.x & 31) / LPR;
  const int p = blockIdx.x * NWK + wk;
  const int kc0 = blockIdx.y * KC;
  const unsigned k4 = a.k / 4;
  const bool col_ok = kc0 / 4 + sl < (int)k4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  if (p >= a.npanels) return;
  float4* acc = reinterpret_cast<float4*>(sacc + (size_t)wk * a.tm * KC) + sl;  // row r at acc[r*LPR]
  for (int r = 0; r < a.tm; ++r) acc[r * LPR] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* B4 = reinterpret_cast<const float4*>(a.B) + c4;
  unsigned lo, hi;
  if (MODE == 0) { lo = a.tileNnz[a.tileRowPtr[p]]; hi = a.tileNnz[a.tileRowPtr[p + 1]]; }
  else { lo = a.segPtr[a.seg_off[p]]; hi = a.segPtr[a.seg_off[p + 1]]; }
  unsigned tcur = MODE == 0 ? a.tileRowPtr[p] : 0;
  for (unsigned e0 = lo; e0 < hi; e0 += LPR) {
    const unsigned e = e0 + sl;
    int r = 0; unsigned c = 0; float v = 0.f;
    if (e < hi) {
      v = a.vals[e];
      if (MODE == 0) {
        const int rc = a.rcOffset[e];
        unsigned t = tcur;
        while (a.tileNnz[t + 1] <= e) ++t;  // tiles are short: a few steps at most
        r = rc >> 16;
        c = a.tileColIdx[t] + (unsigned)(rc & 0xffff);
      } else {
        r = (int)a.segNzRCIdx[2 * (size_t)e];
        c = a.segNzRCIdx[2 * (size_t)e + 1];
      }
    }
    if (MODE == 0) {  // advance the shared tile cursor to the chunk's last entry
      const unsigned elast = min(e0 + LPR, hi) - 1;
      while (a.tileNnz[tcur + 1] <= elast) ++tcur;
    }
    const int cnt = (int)min((unsigned)LPR, hi - e0);
    // the B rows of GB nz are requested before their FMAs (one load in flight per worker left this kernel waiting on L2:
    // 68 stalled warps per issue); a column repeated by the next nz of the stream is an L1 hit
    constexpr int GB = LPR < 8 ? LPR : 8;
    for (int j0 = 0; j0 < cnt; j0 += GB) {
      float4 bb[GB];
      int rr[GB];
      float vv[GB];
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        const int j = (j0 + u) & (LPR - 1);
        const unsigned cc = tile.shfl(c, j);
        rr[u] = tile.shfl(r, j);
        vv[u] = tile.shfl(v, j);
        bb[u] = j0 + u < cnt ? __ldg(B4 + (size_t)cc * k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < GB; ++u)
        if (j0 + u < cnt) {
          float4 x = acc[rr[u] * LPR];
          x.x = fmaf(vv[u], bb[u].x, x.x); x.y = fmaf(vv[u], bb[u].y, x.y); x.z = fmaf(vv[u], bb[u].z, x.z); x.w = fmaf(vv[u], bb[u].w, x.w);
          acc[rr[u] * LPR] = x;
        }
    }
  }
  if (col_ok)
    for (int r = 0; r < a.tm; ++r) {
      const int row = p * a.tm + r;
      if (row < a.m) {
        const int orow = a.vo_mp ? a.vo_mp[row] : row;
        reinterpret_cast<float4*>(a.C)[(size_t)orow * k4 + c4] = acc[r * LPR];
      }
    }
}

// ---------------------------------------------------------------------------------------------
// pillar format kernel (alpha_w_atomic_spmm_v36 semantics)
// ---------------------------------------------------------------------------------------------
struct AlphaArgs {
  const unsigned *alpha_rowPtr, *alpha_colIdx, *pillar_rowPtr, *pillarIdx, *segVoMap;
  const float* alpha_vals;
  unsigned* counter;  // n_sm+1, zeroed before every launch (flex.cu:5058)
  const float* B;
  float* C;
  int n_sm, k;
};

template <int KC>
__global__ void __launch_bounds__(256) k_spmm_alpha(AlphaArgs a) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR;
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int lane = threadIdx.x & 31, sub = lane / LPR;
  const int kc0 = blockIdx.y * KC;
  const unsigned k4 = a.k / 4;
  const bool col_ok = kc0 / 4 + sl < (int)k4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  const float4* B4 = reinterpret_cast<const float4*>(a.B) + c4;
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const int q_own = (int)(smid % (unsigned)a.n_sm);
  // A warp pops one pillar at a time: first from its SM's own queue, then from the shared balance queue (v36's order,
  // flex.cu:4010), then it sweeps every other SM's queue.  The sweep is what makes the result independent of CTA
  // placement: the reference relies on a CTA landing on every SM, which CUDA does not promise (a second feature chunk,
  // opts.n_sm above the SM count, or another kernel on the GPU left queues undrained and rows of C silently zero).
  // (The sweep looks at 32 queues per step -- a lane per queue reads its counter from L2 -- and visits only those with
  // unclaimed pillars: walking all n_sm queues one by one cost every warp ~150 dependent round trips, 0.3 of the 0.5 ms
  // of this kernel on flickr-shape.)
  unsigned sweep_mask = 0;
  int sweep_next = 0, sweep_base = 0;
  for (int phase = 0;; ++phase) {
    int q;
    if (phase == 0) q = q_own;
    else if (phase == 1) q = a.n_sm;
    else {
      while (sweep_mask == 0u && sweep_next < a.n_sm) {
        const int cand = sweep_next + lane;
        bool rem = false;
        if (cand < a.n_sm && cand != q_own) {
          const unsigned size = a.pillarIdx[cand + 1] - a.pillarIdx[cand];
          rem = __ldcg(&a.counter[cand * gridDim.y + blockIdx.y]) < size;
        }
        sweep_mask = __ballot_sync(0xffffffffu, rem);
        sweep_base = sweep_next;
        sweep_next += 32;
      }
      if (sweep_mask == 0u) break;
      q = sweep_base + (__ffs(sweep_mask) - 1);
      sweep_mask &= sweep_mask - 1u;
    }
    const unsigned qbeg = a.pillarIdx[q], qend = a.pillarIdx[q + 1];
    if (qbeg == qend) continue;
    while (true) {
      unsigned pil = 0;
      if (lane == 0) pil = qbeg + atomicAdd(&a.counter[q * gridDim.y + blockIdx.y], 1u);
      pil = __shfl_sync(0xffffffffu, pil, 0);
      if (pil >= qend) break;
      const unsigned r0 = a.pillar_rowPtr[pil], r1 = a.pillar_rowPtr[pil + 1];
      // the row pointers and output maps of up to 32 rows of the pillar in one coalesced load each (they were two dependent
      // loads per row in front of the row's nz)
      unsigned rp_l = 0, vm_l = 0, rbase = r0;
      for (unsigned rb = r0; rb < r1; rb += RPW) {
        if (rb == r0 || rb + RPW > rbase + 31) {
          rbase = rb;
          rp_l = rbase + lane <= r1 ? a.alpha_rowPtr[rbase + lane] : 0u;
          vm_l = rbase + lane < r1 ? a.segVoMap[rbase + lane] : 0u;
        }
        const unsigned r = rb + sub;
        const bool act = r < r1;
        const int li = (int)(r - rbase);  // <= 31 - 1: the entry after the row's is in the same load (or the row is inactive)
        const unsigned lo_s = __shfl_sync(0xffffffffu, rp_l, li & 31), hi_s = __shfl_sync(0xffffffffu, rp_l, (li + 1) & 31);
        const unsigned vm_s = __shfl_sync(0xffffffffu, vm_l, li & 31);
        const unsigned lo = act ? lo_s : 0, hi = act ? hi_s : 0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned e0 = lo; e0 < hi; e0 += LPR) {
          const unsigned e = e0 + sl;
          unsigned off = 0; float v = 0.f;
          if (e < hi) { off = a.alpha_colIdx[e] * k4; v = a.alpha_vals[e]; }
          const int cnt = (int)min((unsigned)LPR, hi - e0);
          constexpr int GB = LPR < 8 ? LPR : 8;  // B rows requested before their FMAs
          for (int j0 = 0; j0 < cnt; j0 += GB) {
            float4 bb[GB];
            float vv[GB];
#pragma unroll
            for (int u = 0; u < GB; ++u) {
              const int j = (j0 + u) & (LPR - 1);
              const unsigned o = tile.shfl(off, j);
              vv[u] = tile.shfl(v, j);
              bb[u] = j0 + u < cnt ? __ldg(B4 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < GB; ++u)
              if (j0 + u < cnt) {
                acc.x = fmaf(vv[u], bb[u].x, acc.x); acc.y = fmaf(vv[u], bb[u].y, acc.y);
                acc.z = fmaf(vv[u], bb[u].z, acc.z); acc.w = fmaf(vv[u], bb[u].w, acc.w);
              }
          }
        }
        if (act && col_ok) {
          const unsigned vm = vm_s;
          float* dst = a.C + (size_t)(vm & 0x7fffffffu) * a.k + (size_t)c4 * 4;
          if (vm & 0x80000000u) {  // the row has nz in other pillars too: accumulate (flex.cu:4108-4118)
            if (hi > lo) { atomicAdd(dst, acc.x); atomicAdd(dst + 1, acc.y); atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w); }
          } else {
            *reinterpret_cast<float4*>(dst) = acc;
          }
        }
      }
    }
  }
}

int pick_kc(int k) { return k <= 32 ? 32 : (k <= 64 ? 64 : 128); }

template <class T>
int d2h(std::vector<T>& dst, const void* src, size_t count) {
  dst.resize(count);
  if (count) FX_CUDA(cudaMemcpy(dst.data(), src, sizeof(T) * count, cudaMemcpyDeviceToHost));
  return FX_OK;
}

}  // namespace

namespace fx {

// ---- round 1 of csr2_DiagTiling lives in fx_flex_host.cu ----
int diag_round1_host(int M, int nnz, const uint32_t* rowptr, const uint32_t* col, int n_sm, std::vector<int>& tile_width,
                     int& warps_with_weights, std::vector<uint8_t>& listed);

static int n_sm_of(const fx_tiles* t) {
  return t->opts.n_sm > 0 ? t->opts.n_sm : sm_count_of_current_device();
}

int flex_carve(fx_tiles* t) {
  fx_flex_dev& f = t->flex;
  const fx_matrix* m = t->mat;
  const int64_t n = m->n, nnz = m->nnz;
  f.tm = t->opts.tm > 0 ? t->opts.tm : 4;
  f.tn = t->opts.tn > 0 ? t->opts.tn : 4;
  f.cmajor = t->opts.cmajor;
  f.nnz_limit = t->opts.nnz_limit > 0 ? t->opts.nnz_limit : 128;
  f.n_sm = n_sm_of(t);
  f.m = (int)n; f.nnz = (int)nnz;
  f.npanels = (int)((n + f.tm - 1) / f.tm);
  FX_REQUIRE(f.tm == 2 || f.tm == 4 || f.tm == 8 || f.tm == 16, FX_ERR_ARG, "tm must be 2, 4, 8 or 16 (tileConfs, flex.cu:4146-4152)");
  FX_REQUIRE(t->format != FX_FMT_TILE || (f.tn >= 1 && f.tn <= 32), FX_ERR_ARG, "tn must be in [1,32] (bitMap is 32 bits)");
  FX_REQUIRE(t->row_begin == 0 && t->row_end == n, FX_ERR_UNSUPPORTED, "Flex formats are built for the whole matrix");
  // the reference asserts that no row is empty (mat.cu:1207, :1359)
  for (int64_t r = 0; r < n; ++r)
    FX_REQUIRE(m->rowptr[r] < m->rowptr[r + 1], FX_ERR_FORMAT, "row %lld is empty: the Flex builders need every row non-empty (mat.cu:1207,1359)", (long long)r);
  size_t bytes = 0;
  auto add = [&](size_t b) { bytes += Arena::pad(b) + 256; };
  add(sizeof(int) * (f.npanels + 2) * 2);
  if (t->format == FX_FMT_TILE) {
    for (int i = 0; i < 4; ++i) add(sizeof(int) * (nnz + 2));  // tileNnz, tileColIdx, nnzTile, bitMap (<= nnz tiles)
    add(sizeof(int) * (nnz + 2)); add(sizeof(float) * (nnz + 2)); add(sizeof(int) * (f.npanels + 2));
  } else if (t->format == FX_FMT_SEG) {
    // segments: every cut but the last of a panel holds >= 116 nz  =>  nsegs <= nnz/116 + npanels
    f.seg_cap = (int)(nnz / (f.nnz_limit - (int)(0.1 * f.nnz_limit) ) + f.npanels + 2);
    const size_t rows_cap = (size_t)f.seg_cap * f.tm + 2;
    add(sizeof(int) * rows_cap * 2);          // alpha_rowPtr, segVoMap
    add(sizeof(int) * (nnz + 2)); add(sizeof(float) * (nnz + 2));  // alpha_colIdx, alpha_vals
    add(sizeof(int) * (f.seg_cap + 2) * 2);   // pillar_rowPtr, segPtr
    add(sizeof(int) * 2 * (nnz + 2)); add(sizeof(float) * (nnz + 2)); add(sizeof(float) * 2 * (nnz + 2));
    add(sizeof(int) * rows_cap); add(sizeof(int) * ((size_t)f.seg_cap * (f.tm + 1) + 2));
    add(sizeof(int) * (f.n_sm + 2) * 2);
  } else if (t->format == FX_FMT_PILLAR) {
    f.partitions = 64 * f.n_sm;  // warps_per_sm * n_sm, mat.cu:688-700
    f.seg_cap = (int)(nnz / (128 - 12) + f.npanels + 2);  // round 3: every cut but the last of a panel holds >= 116 nz
    const size_t rows_cap = (size_t)n + 1 + (size_t)f.seg_cap * f.tm + 2;
    add(sizeof(int) * (f.npanels + 2) * 2);   // nzcnt, nzoff
    add((size_t)n + 4); add(sizeof(int) * (f.partitions + 2));   // listed, pstart
    add(sizeof(int) * (n + 2)); add(sizeof(int) * (n + 2));       // dpos, block ends (round 1 on the GPU)
    add(sizeof(int) * rows_cap); add(sizeof(int) * rows_cap);     // alpha_rowPtr, segVoMap
    add(sizeof(int) * (nnz + 2)); add(sizeof(float) * (nnz + 2)); // alpha_colIdx, alpha_vals
    add(sizeof(int) * (nnz + 2)); add(sizeof(float) * (nnz + 2)); // rest_col, rest_val (what round 2 leaves for round 3)
    add(sizeof(int) * ((size_t)f.partitions + f.seg_cap + 2));    // pillar_rowPtr
    add(sizeof(int) * (f.n_sm + 2)); add(sizeof(int) * (f.n_sm + 1) * 16); add(sizeof(int) * 4);  // pillarIdx, counter, perr
    f.scan_tmp_bytes = 0;
    FX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, f.scan_tmp_bytes, (int*)nullptr, (int*)nullptr, (int)std::max<int64_t>(n + 1, f.npanels + 1)));
    add(f.scan_tmp_bytes);
  }
  int rc = t->arena.reserve(bytes + 1024);
  if (rc != FX_OK) return rc;
  Arena& A = t->arena;
  f.count = A.take<int>(f.npanels + 2);
  f.off = A.take<int>(f.npanels + 2);
  if (t->format == FX_FMT_TILE) {
    f.tileNnz = A.take<unsigned>(nnz + 2); f.tileColIdx = A.take<unsigned>(nnz + 2);
    f.nnzTile = A.take<int>(nnz + 2); f.bitMap = A.take<int>(nnz + 2);
    f.rcOffset = A.take<int>(nnz + 2); f.newVals = A.take<float>(nnz + 2);
    f.tileRowPtr = A.take<unsigned>(f.npanels + 2);
    if (!f.tileRowPtr) { set_error("arena overflow"); return FX_ERR_NOMEM; }
  } else if (t->format == FX_FMT_SEG) {
    const size_t rows_cap = (size_t)f.seg_cap * f.tm + 2;
    f.alpha_rowPtr = A.take<unsigned>(rows_cap); f.segVoMap = A.take<unsigned>(rows_cap);
    f.alpha_colIdx = A.take<unsigned>(nnz + 2); f.alpha_vals = A.take<float>(nnz + 2);
    f.pillar_rowPtr = A.take<unsigned>(f.seg_cap + 2); f.segPtr = A.take<unsigned>(f.seg_cap + 2);
    f.segNzRCIdx = A.take<unsigned>(2 * (nnz + 2)); f.segVals = A.take<float>(nnz + 2);
    f.segNzCV = A.take<float>(2 * (nnz + 2));
    f.segVoMapPad = A.take<unsigned>(rows_cap); f.seg_rowPtr = A.take<int>((size_t)f.seg_cap * (f.tm + 1) + 2);
    f.next_seg = A.take<int>(f.n_sm + 2); f.grouped_tailSeg = A.take<int>(f.n_sm + 2);
    if (!f.grouped_tailSeg) { set_error("arena overflow"); return FX_ERR_NOMEM; }
  } else if (t->format == FX_FMT_PILLAR) {
    const size_t rows_cap = (size_t)n + 1 + (size_t)f.seg_cap * f.tm + 2;
    f.nzcnt = A.take<int>(f.npanels + 2); f.nzoff = A.take<int>(f.npanels + 2);
    f.listed = A.take<unsigned char>(n + 4); f.pstart = A.take<int>(f.partitions + 2);
    f.dpos = A.take<int>(n + 2); f.ends = A.take<int>(n + 2);
    f.alpha_rowPtr = A.take<unsigned>(rows_cap); f.segVoMap = A.take<unsigned>(rows_cap);
    f.alpha_colIdx = A.take<unsigned>(nnz + 2); f.alpha_vals = A.take<float>(nnz + 2);
    f.rest_col = A.take<unsigned>(nnz + 2); f.rest_val = A.take<float>(nnz + 2);
    f.pillar_rowPtr = A.take<unsigned>((size_t)f.partitions + f.seg_cap + 2);
    f.pillarIdx = A.take<unsigned>(f.n_sm + 2); f.counter = A.take<unsigned>((f.n_sm + 1) * 16); f.perr = A.take<int>(4);
    f.scan_tmp = A.take<char>(f.scan_tmp_bytes);
    if (!f.scan_tmp) { set_error("arena overflow"); return FX_ERR_NOMEM; }
  }
  return FX_OK;
}

template <bool WRITE>
static int launch_seg(const fx_tiles* t, cudaStream_t s, SegOut o) {
  const fx_flex_dev& f = t->flex;
  const fx_matrix* m = t->mat;
  const int threads = 128;
  const int grid = ceil_div((long long)f.npanels * f.tm, threads);
#define FX_SEG(TM) k_seg<TM, WRITE><<<grid, threads, 0, s>>>(m->rowptr_dev, m->col_dev, m->val_dev, m->vo_mp_dev, f.m, f.npanels, f.nnz_limit, f.count, f.off, o)
  switch (f.tm) { case 2: FX_SEG(2); break; case 4: FX_SEG(4); break; case 8: FX_SEG(8); break; default: FX_SEG(16); }
#undef FX_SEG
  FX_LAUNCH_CHECK();
  return FX_OK;
}

template <bool WRITE>
static int launch_tile(const fx_tiles* t, cudaStream_t s, TileOut o) {
  const fx_flex_dev& f = t->flex;
  const fx_matrix* m = t->mat;
  const int threads = 128;
  const int grid = ceil_div((long long)f.npanels * f.tm, threads);
#define FX_TILE(TM) k_tile<TM, WRITE><<<grid, threads, 0, s>>>(m->rowptr_dev, m->col_dev, m->val_dev, f.m, f.m, f.npanels, f.tn, f.cmajor, f.count, f.off, o)
  switch (f.tm) { case 2: FX_TILE(2); break; case 4: FX_TILE(4); break; case 8: FX_TILE(8); break; default: FX_TILE(16); }
#undef FX_TILE
  FX_LAUNCH_CHECK();
  return FX_OK;
}

int flex_build(fx_tiles* t, cudaStream_t s) {
  fx_flex_dev& f = t->flex;
  if (t->format == FX_FMT_SEG) {
    SegOut o{f.alpha_rowPtr, f.alpha_colIdx, f.pillar_rowPtr, f.segVoMap, f.alpha_vals, f.segPtr, f.segNzRCIdx,
             f.segVoMapPad, f.segVals, f.segNzCV, f.seg_rowPtr};
    int rc = launch_seg<false>(t, s, o);
    if (rc) return rc;
    k_scan_int<<<1, 1024, 0, s>>>(f.count, f.npanels, f.off);
    FX_LAUNCH_CHECK();
    rc = launch_seg<true>(t, s, o);
    if (rc) return rc;
    // totals + the per-panel counts come to the host: the F4 bucket walk is a 149-step greedy
    f.h_count.resize(f.npanels);
    FX_CUDA(cudaMemcpyAsync(t->stats_host, f.off + f.npanels, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(f.h_count.data(), f.count, sizeof(int) * f.npanels, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    f.nsegs = *reinterpret_cast<int*>(t->stats_host);
    const int last_rows = f.m - (f.npanels - 1) * f.tm;
    f.rows_total = f.nsegs * f.tm - (f.npanels ? f.h_count[f.npanels - 1] * (f.tm - last_rows) : 0);
    k_seg_tail<<<1, 1, 0, s>>>(o, f.nsegs, f.rows_total, (unsigned)f.nnz);
    FX_LAUNCH_CHECK();
    // F4 (mat.cu:1118-1162, row_based_split)
    f.h_next.assign(f.n_sm + 1, 0); f.h_tail.assign(f.n_sm + 1, 0);
    {
      const int segload = f.nsegs / f.n_sm;
      int head = 0, tail = 0, panel = 0;
      for (int i = 0; i < f.n_sm; ++i) {
        f.h_next[i] = head;
        if (panel < f.npanels) {
          int cur = f.h_count[panel];
          tail = head + cur;
          while (++panel < f.npanels) {
            if (f.h_count[panel] + cur > segload) break;
            cur += f.h_count[panel];
            tail += f.h_count[panel];
          }
        }
        f.h_tail[i] = std::min(f.nsegs, tail);
        head = std::min(f.nsegs, tail);
      }
      f.h_next[f.n_sm] = head;
      f.h_tail[f.n_sm] = f.nsegs;
    }
    FX_CUDA(cudaMemcpyAsync(f.next_seg, f.h_next.data(), sizeof(int) * (f.n_sm + 1), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpyAsync(f.grouped_tailSeg, f.h_tail.data(), sizeof(int) * (f.n_sm + 1), cudaMemcpyHostToDevice, s));
    return FX_OK;
  }
  if (t->format == FX_FMT_TILE) {
    TileOut o{f.tileRowPtr, f.tileNnz, f.tileColIdx, f.nnzTile, f.bitMap, f.rcOffset, f.newVals};
    int rc = launch_tile<false>(t, s, o);
    if (rc) return rc;
    k_scan_int<<<1, 1024, 0, s>>>(f.count, f.npanels, f.off);
    FX_LAUNCH_CHECK();
    rc = launch_tile<true>(t, s, o);
    if (rc) return rc;
    k_tile_tail<<<ceil_div(f.npanels + 1, 256), 256, 0, s>>>(o, f.off, f.npanels, (unsigned)f.nnz);
    FX_LAUNCH_CHECK();
    FX_CUDA(cudaMemcpyAsync(t->stats_host, f.off + f.npanels, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    f.ntiles = *reinterpret_cast<int*>(t->stats_host);
    return FX_OK;
  }
  if (t->format == FX_FMT_PILLAR) {
    // Round 1 of csr2_DiagTiling grows the diagonal blocks one row at a time, each block starting where the previous one ended
    // (mat.cu:706-759): a serial dependence over the whole diagonal, kept on the host (inside tPre, like the reference's
    // whole builder).  Rounds 2 and 3 run here on the GPU from the block boundaries and one byte per column.
    const fx_matrix* m = t->mat;
    std::vector<int> tile_width;
    std::vector<uint8_t> listed;
    int wpw = 0, rc = FX_OK;
    int* sh = reinterpret_cast<int*>(t->stats_host);
    // round 1 on the GPU where its preconditions hold (see k_diag_ends), else on the host
    const int nnz_p_diagonal_tile = std::max(32, (int)(0.3f * (float)m->rowptr[f.m]) / f.partitions);  // mat.cu:690-704
    const int thr = (int)(0.85 * nnz_p_diagonal_tile);
    static const int r1_env = getenv("FLEX_PILLAR_ROUND1") ? atoi(getenv("FLEX_PILLAR_ROUND1")) : -1;  // 0 = host, 1 = GPU if possible
    bool on_gpu = false;
    FX_CUDA(cudaMemsetAsync(f.perr, 0, sizeof(int) * 4, s));
    if (r1_env != 0 && f.nnz > 0) {
      k_diag_pos<<<ceil_div(f.m, 256), 256, 0, s>>>(m->rowptr_dev, m->col_dev, f.m, f.dpos, f.perr);
      FX_LAUNCH_CHECK();
      k_band_sym<<<ceil_div(f.m, 256), 256, 0, s>>>(m->rowptr_dev, m->col_dev, f.dpos, f.m, thr + 1, f.perr);
      FX_LAUNCH_CHECK();
      FX_CUDA(cudaMemcpyAsync(sh, f.perr, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
      FX_CUDA(cudaStreamSynchronize(s));
      on_gpu = sh[1] == 0 && sh[2] == 0;
    }
    std::vector<int> pstart;
    if (on_gpu) {
      k_diag_ends<<<ceil_div(f.m, 128), 128, 0, s>>>(m->rowptr_dev, m->col_dev, f.dpos, f.m, thr, f.ends);
      FX_LAUNCH_CHECK();
      std::vector<int> ends((size_t)f.m);
      FX_CUDA(cudaMemcpyAsync(ends.data(), f.ends, sizeof(int) * (size_t)f.m, cudaMemcpyDeviceToHost, s));
      FX_CUDA(cudaStreamSynchronize(s));
      tile_width.assign(f.partitions, 0);
      int r0 = 0;
      for (int i = 0; i < f.partitions && r0 < f.m; ++i) { tile_width[i] = ends[r0] - r0; r0 += tile_width[i]; ++wpw; }
      FX_REQUIRE(r0 == f.m, FX_ERR_FORMAT, "alpha is too small: diagonal blocks cover %d of %d rows (assert mat.cu:759)", r0, f.m);
      FX_REQUIRE(f.partitions != wpw, FX_ERR_FORMAT, "no idle warp left for the balance queue (division by zero at mat.cu:862)");
    } else {
      // a matrix created from device arrays (fx_csr_from_device) has no host copy: its structure comes down once, inside tPre
      std::vector<uint32_t> h_rowptr, h_col;
      const uint32_t *hr = m->rowptr.data(), *hc = m->col.data();
      if (m->col.empty() && f.nnz > 0) {
        if ((rc = d2h(h_rowptr, m->rowptr_dev, (size_t)f.m + 1)) || (rc = d2h(h_col, m->col_dev, (size_t)f.nnz))) return rc;
        hr = h_rowptr.data(); hc = h_col.data();
      }
      rc = diag_round1_host(f.m, f.nnz, hr, hc, f.n_sm, tile_width, wpw, listed);
      if (rc) return rc;
    }
    pstart.assign(wpw + 1, 0);
    for (int i = 0; i < wpw; ++i) pstart[i + 1] = pstart[i] + tile_width[i];
    FX_REQUIRE(pstart[wpw] == f.m, FX_ERR_FORMAT, "diagonal blocks do not cover every row (assert mat.cu:853)");
    FX_CUDA(cudaMemcpyAsync(f.pstart, pstart.data(), sizeof(int) * (wpw + 1), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpyAsync(f.pillar_rowPtr, pstart.data(), sizeof(int) * (wpw + 1), cudaMemcpyHostToDevice, s));
    if (on_gpu) {
      FX_CUDA(cudaMemsetAsync(f.listed, 0, (size_t)f.m, s));
      k_diag_listed<<<ceil_div((long long)f.m * 32, 256), 256, 0, s>>>(m->rowptr_dev, m->col_dev, f.pstart, wpw, f.m, f.listed);
      FX_LAUNCH_CHECK();
    } else {
      FX_CUDA(cudaMemcpyAsync(f.listed, listed.data(), (size_t)f.m, cudaMemcpyHostToDevice, s));
    }
    FX_CUDA(cudaMemsetAsync(f.perr, 0, sizeof(int) * 4, s));
    f.round1_on_gpu = on_gpu;
    PillarIn in{m->rowptr_dev, m->col_dev, m->val_dev, m->vo_mp_dev, f.listed, f.pstart, f.m, wpw};
    PillarOut o{f.alpha_rowPtr, f.alpha_colIdx, f.pillar_rowPtr, f.segVoMap, f.alpha_vals, f.rest_col, f.rest_val};
    // round 2 (mat.cu:771-835)
    const int rgrid = ceil_div(((long long)f.m + 1) * 32, 256);
    k_pillar_count<<<rgrid, 256, 0, s>>>(in, o);
    FX_LAUNCH_CHECK();
    FX_CUDA(cub::DeviceScan::ExclusiveSum(f.scan_tmp, f.scan_tmp_bytes, (int*)f.alpha_rowPtr, (int*)f.alpha_rowPtr, f.m + 1, s));
    k_pillar_fill<<<rgrid, 256, 0, s>>>(in, o, f.perr);
    FX_LAUNCH_CHECK();
    // round 3 (mat.cu:871-878): what is left, as column-major segments per tm-row panel, into the shared balance queue
    const int threads = 128;
    const int grid = ceil_div((long long)f.npanels * f.tm, threads);
#define FX_PSEG(TM, W) k_pseg<TM, W><<<grid, threads, 0, s>>>(in, f.npanels, 128, f.count, f.nzcnt, f.off, f.nzoff, o)
    switch (f.tm) { case 2: FX_PSEG(2, false); break; case 4: FX_PSEG(4, false); break; case 8: FX_PSEG(8, false); break; default: FX_PSEG(16, false); }
    FX_LAUNCH_CHECK();
    FX_CUDA(cub::DeviceScan::ExclusiveSum(f.scan_tmp, f.scan_tmp_bytes, f.count, f.off, f.npanels + 1, s));
    FX_CUDA(cub::DeviceScan::ExclusiveSum(f.scan_tmp, f.scan_tmp_bytes, f.nzcnt, f.nzoff, f.npanels + 1, s));
    switch (f.tm) { case 2: FX_PSEG(2, true); break; case 4: FX_PSEG(4, true); break; case 8: FX_PSEG(8, true); break; default: FX_PSEG(16, true); }
#undef FX_PSEG
    FX_LAUNCH_CHECK();
    FX_CUDA(cudaMemcpyAsync(sh + 0, f.off + f.npanels, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(sh + 1, f.count + (f.npanels - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(sh + 2, f.alpha_rowPtr + f.m, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(sh + 3, f.nzoff + f.npanels, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(sh + 4, f.perr, sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    const int nsegs3 = sh[0], last_cnt = sh[1];
    f.r2_nnz = sh[2];
    FX_REQUIRE(sh[4] == 0, FX_ERR_FORMAT, "pillar %d without nz (assert mat.cu:835)", sh[4] - 1);
    FX_REQUIRE(f.r2_nnz + sh[3] == f.nnz, FX_ERR_FORMAT, "pillar format lost nz (assert mat.cu:895)");
    FX_REQUIRE(nsegs3 <= f.seg_cap, FX_ERR_FORMAT, "pillar format: more balance segments than reserved");
    const int last_rows = f.m - (f.npanels - 1) * f.tm;
    f.rows_total = f.m + nsegs3 * f.tm - last_cnt * (f.tm - last_rows);
    f.nsegs = wpw + nsegs3;
    k_pillar_tail<<<1, 1, 0, s>>>(o, wpw, nsegs3, (unsigned)f.rows_total);
    FX_LAUNCH_CHECK();
    auto& P = f.ph;
    P = fx_flex_dev::PillarHost();
    for (int i = 0; i < wpw; i += 64) P.alpha_pillarIdx.push_back((unsigned)i);              // mat.cu:831-833
    while ((int)P.alpha_pillarIdx.size() <= f.n_sm) P.alpha_pillarIdx.push_back((unsigned)wpw);  // mat.cu:835
    P.alpha_pillarIdx.push_back((unsigned)f.nsegs);                                          // mat.cu:881
    FX_REQUIRE((int)P.alpha_pillarIdx.size() == f.n_sm + 2, FX_ERR_FORMAT, "alpha_pillarIdx has %zu entries (assert mat.cu:899)", P.alpha_pillarIdx.size());
    P.n_segs = f.nsegs;
    P.warps_with_weights = wpw;
    P.empty_wp_p = (1 - (float)wpw / f.partitions) * 100;
    P.band_nz_p = (float)f.r2_nnz / m->rowptr[f.m] * 100;
    FX_CUDA(cudaMemcpyAsync(f.pillarIdx, P.alpha_pillarIdx.data(), sizeof(unsigned) * (f.n_sm + 2), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaStreamSynchronize(s));
    return FX_OK;
  }
  set_error("flex_build: unknown format");
  return FX_ERR_ARG;
}

int flex_spmm(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s) {
  const fx_flex_dev& f = t->flex;
  const fx_matrix* m = t->mat;
  FX_REQUIRE(k % 4 == 0, FX_ERR_UNSUPPORTED, "Flex-format kernels need k %% 4 == 0 (vec4 kernels, SURVEY 8b)");
  const int KC = pick_kc(k);
  const int kchunks = ceil_div(k, KC);
  if (t->format == FX_FMT_PILLAR) {
    AlphaArgs a{f.alpha_rowPtr, f.alpha_colIdx, f.pillar_rowPtr, f.pillarIdx, f.segVoMap, f.alpha_vals, f.counter, B, C, f.n_sm, k};
    FX_REQUIRE(kchunks <= 16, FX_ERR_UNSUPPORTED, "k too large for the pillar kernel's counters");
    FX_CUDA(cudaMemsetAsync(f.counter, 0, sizeof(unsigned) * (f.n_sm + 1) * kchunks, s));  // flex.cu:5058
    FX_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)f.m * k, s));                    // flex.cu:5057 (atomics)
    dim3 grid(f.n_sm * 8, kchunks);
    if (KC == 32) k_spmm_alpha<32><<<grid, 256, 0, s>>>(a);
    else if (KC == 64) k_spmm_alpha<64><<<grid, 256, 0, s>>>(a);
    else k_spmm_alpha<128><<<grid, 256, 0, s>>>(a);
    FX_LAUNCH_CHECK();
    return FX_OK;
  }
  AccArgs a{};
  a.B = B; a.C = C; a.m = f.m; a.npanels = f.npanels; a.k = k; a.tm = f.tm;
  a.vo_mp = m->info.order != FX_ORDER_OVO ? m->vo_mp_dev : nullptr;
  const int RPW = 32 / (KC / 4), NWK = 8 * RPW;
  const size_t smem = (size_t)NWK * f.tm * KC * sizeof(float);
  dim3 grid(ceil_div(f.npanels, NWK), kchunks);
#define FX_ACC(KCV, MODE)                                                                                       \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      FX_CUDA(cudaFuncSetAttribute(k_spmm_panel_acc<KCV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_spmm_panel_acc<KCV, MODE><<<grid, 256, smem, s>>>(a);                                                     \
  } while (0)
  if (t->format == FX_FMT_TILE) {
    a.tileRowPtr = f.tileRowPtr; a.tileNnz = f.tileNnz; a.tileColIdx = f.tileColIdx; a.rcOffset = f.rcOffset; a.vals = f.newVals;
    if (KC == 32) FX_ACC(32, 0); else if (KC == 64) FX_ACC(64, 0); else FX_ACC(128, 0);
  } else {
    a.seg_off = f.off; a.segPtr = f.segPtr; a.segNzRCIdx = f.segNzRCIdx; a.vals = f.segVals;
    if (KC == 32) FX_ACC(32, 1); else if (KC == 64) FX_ACC(64, 1); else FX_ACC(128, 1);
  }
#undef FX_ACC
  FX_LAUNCH_CHECK();
  return FX_OK;
}

void flex_release(fx_tiles*) {}  // every array of the Flex formats lives in the handle's arena

}  // namespace fx

extern "C" int fx_tiles_export_tile(fx_tiles* t, fx_tile_arrays* o) {
  FX_REQUIRE(t && o && t->format == FX_FMT_TILE, FX_ERR_ARG, "fx_tiles_export_tile: not a tile-format handle");
  fx_flex_dev& f = t->flex;
  FX_CUDA(cudaDeviceSynchronize());
  int rc;
  if ((rc = d2h(f.e_u32[0], f.tileRowPtr, f.npanels + 1)) || (rc = d2h(f.e_u32[1], f.tileNnz, f.ntiles + 1)) ||
      (rc = d2h(f.e_i32[0], f.nnzTile, f.ntiles)) || (rc = d2h(f.e_i32[1], f.bitMap, f.ntiles)) ||
      (rc = d2h(f.e_u32[2], f.tileColIdx, f.ntiles)) || (rc = d2h(f.e_i32[2], f.rcOffset, f.nnz)) ||
      (rc = d2h(f.e_f32[0], f.newVals, f.nnz)))
    return rc;
  o->m = f.m; o->tm = f.tm; o->tn = f.tn; o->ntiles = f.ntiles; o->npanels = f.npanels; o->nnz = f.nnz; o->cmajor = f.cmajor;
  o->tileRowPtr = f.e_u32[0].data(); o->tileNnz = f.e_u32[1].data(); o->nnzTile = f.e_i32[0].data();
  o->bitMap = f.e_i32[1].data(); o->tileColIdx = f.e_u32[2].data(); o->rcOffset = f.e_i32[2].data();
  o->newVals = f.e_f32[0].data();
  return FX_OK;
}

extern "C" int fx_tiles_export_seg(fx_tiles* t, fx_seg_arrays* o) {
  FX_REQUIRE(t && o && t->format == FX_FMT_SEG, FX_ERR_ARG, "fx_tiles_export_seg: not a seg-format handle");
  fx_flex_dev& f = t->flex;
  FX_CUDA(cudaDeviceSynchronize());
  int rc;
  const size_t R = f.rows_total, S = f.nsegs, N = f.nnz;
  if ((rc = d2h(f.e_u32[0], f.alpha_rowPtr, R + 1)) || (rc = d2h(f.e_u32[1], f.alpha_colIdx, N)) ||
      (rc = d2h(f.e_u32[2], f.pillar_rowPtr, S + 1)) || (rc = d2h(f.e_u32[3], f.segVoMap, R)) ||
      (rc = d2h(f.e_f32[0], f.alpha_vals, N)) || (rc = d2h(f.e_u32[4], f.segPtr, S + 1)) ||
      (rc = d2h(f.e_u32[5], f.segNzRCIdx, 2 * N)) || (rc = d2h(f.e_u32[6], f.segVoMapPad, S * f.tm)) ||
      (rc = d2h(f.e_f32[1], f.segVals, N)) || (rc = d2h(f.e_f32[2], f.segNzCV, 2 * N)) ||
      (rc = d2h(f.e_i32[0], f.seg_rowPtr, S * (f.tm + 1))))
    return rc;
  o->m = f.m; o->tm = f.tm; o->nnz = f.nnz; o->nsegs = f.nsegs; o->rows_total = f.rows_total; o->npanels = f.npanels; o->n_sm = f.n_sm;
  o->alpha_rowPtr = f.e_u32[0].data(); o->alpha_colIdx = f.e_u32[1].data(); o->alpha_pillar_rowPtr = f.e_u32[2].data();
  o->segVoMap = f.e_u32[3].data(); o->alpha_vals = f.e_f32[0].data(); o->segs_per_panel = f.h_count.data();
  o->segPtr = f.e_u32[4].data(); o->segNzRCIdx = f.e_u32[5].data(); o->segVoMapPad = f.e_u32[6].data();
  o->segVals = f.e_f32[1].data(); o->segNzCV = f.e_f32[2].data(); o->seg_rowPtr = f.e_i32[0].data();
  o->next_seg = f.h_next.data(); o->grouped_tailSeg = f.h_tail.data();
  return FX_OK;
}

extern "C" int fx_tiles_export_pillar(fx_tiles* t, fx_pillar_arrays* o) {
  FX_REQUIRE(t && o && t->format == FX_FMT_PILLAR, FX_ERR_ARG, "fx_tiles_export_pillar: not a pillar-format handle");
  fx_flex_dev& f = t->flex;
  auto& P = f.ph;
  FX_CUDA(cudaDeviceSynchronize());
  int rc;
  const size_t R = f.rows_total, S = f.nsegs, N = f.nnz;
  if ((rc = d2h(P.alpha_rowPtr, f.alpha_rowPtr, R + 1)) || (rc = d2h(P.alpha_colIdx, f.alpha_colIdx, N)) ||
      (rc = d2h(P.alpha_pillar_rowPtr, f.pillar_rowPtr, S + 1)) || (rc = d2h(P.segVoMap, f.segVoMap, R)) ||
      (rc = d2h(P.alpha_vals, f.alpha_vals, N)))
    return rc;
  o->m = f.m; o->nnz = f.nnz; o->n_sm = f.n_sm; o->n_segs = P.n_segs; o->rows_total = f.rows_total;
  o->warps_with_weights = P.warps_with_weights;
  o->alpha_rowPtr = P.alpha_rowPtr.data(); o->alpha_colIdx = P.alpha_colIdx.data();
  o->alpha_pillar_rowPtr = P.alpha_pillar_rowPtr.data(); o->alpha_pillarIdx = P.alpha_pillarIdx.data();
  o->segVoMap = P.segVoMap.data(); o->alpha_vals = P.alpha_vals.data();
  o->empty_wp_p = P.empty_wp_p; o->band_nz_p = P.band_nz_p;
  o->round1_on_gpu = f.round1_on_gpu ? 1 : 0;
  return FX_OK;
}
