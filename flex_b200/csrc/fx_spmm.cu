// fx_spmm.cu -- the SpMM kernels (C = A*B, fp32, B/C row-major) for sm_100a.
//
// Work mapping (contrast: the reference's kernels use one lane per output column, scalar 4-byte
// B loads and fp32 atomics into a pre-zeroed C -- aspt/sspmm_128.cu:321-427,703-827):
//   * a row of C is owned by LPR = KC/4 lanes; each lane keeps one float4 of the row in registers,
//     so a B row is fetched with one 128-bit load per lane (512 B per warp instruction at k=128);
//   * (col,val) pairs are read coalesced, 32 nz at a time, staged per worker in shared memory as (B-row offset, value)
//     and read back four at a time with broadcast LDS.128 (the 512-chunk kernel broadcasts with warp shuffles);
//   * B rows carry an evict_last L2 policy, the streams read or written once bypass L1 and carry evict_first;
//   * a panel's dense tiles are staged in shared memory once per CTA by the TMA unit
//     (cp.async.bulk global->shared, completion on an mbarrier), one 16-byte-aligned row of B per
//     bulk copy, and read back with 128-bit shared loads;
//   * every C element is written exactly once with a plain 128-bit store (no atomics, no memset):
//     the 512-nz chunks of very long rows (the reference's "special" lists, :1076-1087) are reduced
//     into a scratch buffer by a first kernel and folded in, in a fixed order, by the row's owner.
#include <cooperative_groups.h>

#include <type_traits>

#include "fx_common.cuh"
#include "fx_tc_kernel.cuh"
#include "fx_gemm_kernel.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int BH = 128;
constexpr int STHRESHOLD = 512;
constexpr int PANEL_WARPS = 8;
constexpr int MAX_TS = 4;  // dense tiles of one panel resident in shared memory at a time

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_arrive(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 residency: the step touches B (re-read per nz, the only reused data) next to streams read or written once
// (A's columns and values, tc_out, the chunk partials, C).  B rows are loaded with an evict_last policy, the
// streams bypass L1 and carry evict_first, so B stays in L2 when the streams pass through (`hints` = 0 gives
// every access the normal priority: FLEX_HINTS=0).
struct Policies { uint64_t keep, stream; };
__device__ __forceinline__ Policies make_policies(int hints) {
  Policies p;
  if (hints) {
    // (tried: evict_last on a fixed fraction of B only, evict_first on the rest -- 0.75 / 0.5 / 0.25 are each worse than the
    // one before: Reddit-shape 0.556 -> 0.573 / 0.593 / 0.621 ms, yelp-shape 0.583 -> 0.631 / 0.678 / 0.719)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  } else {
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p.keep));
    p.stream = p.keep;
  }
  return p;
}
__device__ __forceinline__ float4 ldg4_keep(const float4* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// B row `off` (in float4 units, < 2^32: fx_spmm refuses larger n*k) from this lane's base pointer: one IMAD.WIDE.U32 per
// address (written as C++ pointer arithmetic the compiler re-derives the base from its parts and spends four
// instructions per load on the 64-bit sum)
__device__ __forceinline__ float4 ldg4_keep_at(const float4* base, unsigned off, uint64_t pol) {
  const float4* p;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(p) : "r"(off), "l"(base));
  return ldg4_keep(p, pol);
}
__device__ __forceinline__ float4 ldg4_stream(const float4* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_stream(const int* p, uint64_t pol) {
  int v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p, uint64_t pol) {
  float v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg4_stream(float4* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// acc += v * b with two packed FFMA2 (fma.rn.f32x2, sm_100+): same rounding as four fmaf, half the
// issue slots -- the panel kernel is issue-bound, not FMA-bound.
__device__ __forceinline__ void fma4(float4& a, float v, const float4& b) {
  unsigned long long vv, b01, b23, a01, a23;
  asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b01) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b23) : "f"(b.z), "f"(b.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(a.z), "f"(a.w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a01) : "l"(vv), "l"(b01));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a23) : "l"(vv), "l"(b23));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(a01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.z), "=f"(a.w) : "l"(a23));
}

// acc += sum over nz [lo,hi) of val * B[col,:]  with B read from global memory (L1/L2 path).
// `B4` already points at this lane's float4 of row 0; a row is `k4` float4 wide.
template <int LPR, class Tile>
__device__ __forceinline__ void accum_global(const Tile& tile, int lo, int hi, const int* __restrict__ ce,
                                             const float* __restrict__ cv, const float4* __restrict__ B4,
                                             unsigned k4, float4& acc, const Policies& pol) {
  const int sl = tile.thread_rank();
  for (int e0 = lo; e0 < hi; e0 += LPR) {
    const int e = e0 + sl;
    unsigned off = 0;
    float v = 0.f;
    if (e < hi) { off = (unsigned)ldg_stream(ce + e, pol.stream) * k4; v = ldg_stream(cv + e, pol.stream); }
    const int cnt = min(LPR, hi - e0);
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
      const unsigned o0 = tile.shfl(off, j), o1 = tile.shfl(off, j + 1), o2 = tile.shfl(off, j + 2),
                     o3 = tile.shfl(off, j + 3);
      const float v0 = tile.shfl(v, j), v1 = tile.shfl(v, j + 1), v2 = tile.shfl(v, j + 2), v3 = tile.shfl(v, j + 3);
      const float4 b0 = ldg4_keep_at(B4, o0, pol.keep), b1 = ldg4_keep_at(B4, o1, pol.keep), b2 = ldg4_keep_at(B4, o2, pol.keep),
                   b3 = ldg4_keep_at(B4, o3, pol.keep);
      fma4(acc, v0, b0); fma4(acc, v1, b1); fma4(acc, v2, b2); fma4(acc, v3, b3);
    }
    for (; j < cnt; ++j) {
      const unsigned o = tile.shfl(off, j);
      const float vv = tile.shfl(v, j);
      fma4(acc, vv, ldg4_keep_at(B4, o, pol.keep));
    }
  }
}

struct PanelArgs {
  const int* mcsr_cnt;
  const int* mcsr_e;
  const int* mcsr_list;
  const int* csr_e;
  const float* csr_ev;
  const int* spec_off;   // nullptr: rows are never split
  const int* spec_order; // execution order of the 512-chunks (chunk ids sorted by first column), nullptr = list order
  const float* partial;  // [chunk][k] partial sums of the 512-chunks
  const float* B;
  float* C;
  int npanel, nloc, k, BW, TS;  // k = row stride of B, C, partial and tc_out (floats)
  int width;                    // feature columns computed, from the B/C pointers on (== k except in column-chunk launches)
  int split;  // CTAs per panel (>1 when the shard has too few panels to fill the GPU); cut at row boundaries
  const int2* wl;  // per-CTA work list (panel, part | parts << 8) of the row kernel; nullptr = blockIdx.x / split
  // FX_FMT_TCW: products of the panels' tensor windows, added when a row is stored
  const float* tc_out;  // [ntc][128][k]
  const int* tc_slot;   // [npanel] position in tc_out or -1; nullptr = no windows
  int hints;            // L2 eviction priorities (make_policies)
  int team_row;         // rows of at least this many handled nz go to a team of four warps; 0 = by width (k_spmm_rows)
  int tiles_ok;         // bit 31 of a staged offset is free to flag "this B row is in the shared-memory tiles"
  int p_lo, p_hi;       // only the panels in [p_lo, p_hi) are multiplied (row groups of the host pipeline, fx_spmm_host)
};

// ---- long-row chunks: partial[i,:] = sum over the i-th 512-nz chunk --------------------------
template <int KC>
__global__ void __launch_bounds__(PANEL_WARPS * 32) k_spmm_special(PanelArgs a, const int* __restrict__ special,
                                                                  const int* __restrict__ special2, int special_p,
                                                                  float* __restrict__ partial) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR;
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int item = (blockIdx.x * PANEL_WARPS + (threadIdx.x >> 5)) * RPW + ((threadIdx.x & 31) / LPR);
  const int kc0 = blockIdx.y * KC;
  if (item >= special_p) return;  // uniform over the tile
  const unsigned k4 = a.k / 4;
  const bool col_ok = kc0 / 4 + sl < a.width / 4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  const int row = special[item], off = special2[item];
  const int p = row / BH, r = row % BH;
  if (p < a.p_lo || p >= a.p_hi) return;  // uniform over the tile
  const int cnt0 = a.mcsr_cnt[p], delta = a.mcsr_cnt[p + 1] - cnt0;
  // the row's nch chunks are its LAST nch*512 nz (the panel kernel streams everything before them)
  const int nch = a.spec_off[row + 1] - a.spec_off[row];
  const int lo = a.mcsr_e[cnt0 * BH + (r + 1) * delta] - nch * STHRESHOLD + off;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const Policies pol = make_policies(a.hints);
  accum_global<LPR>(tile, lo, lo + STHRESHOLD, a.csr_e, a.csr_ev, reinterpret_cast<const float4*>(a.B) + c4, k4, acc, pol);
  if (col_ok) reinterpret_cast<float4*>(partial)[(size_t)item * k4 + c4] = acc;
}

// Few chunks (small matrices with a handful of long rows): one CTA per chunk, its workers take equal
// slices of the 512 nz and are summed in worker order through shared memory, so that a short list of
// chunks still fills the machine (one warp streaming 512 nz alone is pure DRAM latency).
// (tried: metadata staged in shared memory and read with broadcast LDS.128 as in k_spmm_rows instead of two SHFL per nz --
// the shuffles are 29 % of this kernel's L1 data-pipe wavefronts -- same time: 0.5676 vs 0.5669 ms on Reddit-shape k=128)
template <int KC>
__global__ void __launch_bounds__(PANEL_WARPS * 32) k_spmm_special_cta(PanelArgs a, const int* __restrict__ special,
                                                                      const int* __restrict__ special2,
                                                                      float* __restrict__ partial) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR, NWK = PANEL_WARPS * RPW, SLICE = STHRESHOLD / NWK;
  __shared__ __align__(16) float red[NWK][KC];
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int wk = (threadIdx.x >> 5) * RPW + (threadIdx.x & 31) / LPR;
  // chunks run in the order of the columns they read (fx_aspt_build.cu:k_special_keys); partial[] is indexed by chunk id
  const int item = a.spec_order ? a.spec_order[blockIdx.x] : (int)blockIdx.x, kc0 = blockIdx.y * KC;
  const unsigned k4 = a.k / 4;
  const bool col_ok = kc0 / 4 + sl < a.width / 4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  const int row = special[item], off = special2[item];
  const int p = row / BH, r = row % BH;
  if (p < a.p_lo || p >= a.p_hi) return;  // uniform over the CTA
  const int cnt0 = a.mcsr_cnt[p], delta = a.mcsr_cnt[p + 1] - cnt0;
  const int nch = a.spec_off[row + 1] - a.spec_off[row];
  const int lo = a.mcsr_e[cnt0 * BH + (r + 1) * delta] - nch * STHRESHOLD + off + wk * SLICE;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const Policies pol = make_policies(a.hints);
  accum_global<LPR>(tile, lo, lo + SLICE, a.csr_e, a.csr_ev, reinterpret_cast<const float4*>(a.B) + c4, k4, acc, pol);
  reinterpret_cast<float4*>(red[wk])[sl] = acc;
  __syncthreads();
  if (wk == 0 && col_ok) {
    float4 t = reinterpret_cast<const float4*>(red[0])[sl];
    for (int w2 = 1; w2 < NWK; ++w2) {
      const float4 x = reinterpret_cast<const float4*>(red[w2])[sl];
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
    }
    reinterpret_cast<float4*>(partial)[(size_t)item * k4 + c4] = t;
  }
}

// ---- row-grab kernel: one CTA per 128-row panel, workers take whole rows from a shared counter ----
// The panel's handled nz are every row's [dense groups | sparse tail], i.e. everything except the
// 512-chunks k_spmm_special takes from the END of long sparse groups.  A worker (LPR lanes = one row of
// C, each lane a float4 of features) owns whole rows: it grabs the next row of the panel from a
// shared-memory counter (long rows in a first pass, so none starts when the others run out; the longest
// rows are shared by teams of four warps, see the three passes below), streams the row's nz in chunks of
// 32 -- metadata staged as offsets and values in a per-worker shared buffer and read back four nz per
// broadcast LDS.128, B rows fetched with 128-bit loads, FFMA2 -- starting from the row's tensor-window
// product, adds the row's 512-chunk partials and stores the row once: every C element is written exactly
// once, in a fixed summation order, without atomics.
// (Round-1 history: an nz-balanced variant -- workers take equal slices of the panel's nz stream, rows
// looked up per nz through a prefix table, shared rows reduced through shared memory -- needed ~3x the
// instructions per nz on short rows: 0.729 ms vs 0.634 ms on Reddit-shape)
// TILES=false: no dynamic shared memory beyond the staging buffers, the SM keeps its L1.
// TILES=true : the first TS dense tiles of the panel are TMA-staged in shared memory; nz of those
//              tiles read B from there, all other nz (further tiles, sparse tail) from L1/L2.
// G = B rows requested back to back before the first FMA of a group (a multiple of 4, at most LPR): the loaded
// rows sit in registers, so G*4 registers per thread are the kernel's in-flight buffer.  At 40 registers (three
// 16-warp CTAs per SM) the compiler cannot hold more than four rows; G = 16 needs the 64-80 register launch
// configurations (launch_aspt).
template <int KC, int WARPS, bool TILES, int MINB, int G>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_spmm_rows(PanelArgs a, const int* __restrict__ plist) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR;  // NW = WARPS * RPW workers per CTA
  // rows of >= TEAM_ROW nz are shared by a team of four warps, rows of >= LONG_ROW are taken before the short ones
  constexpr int LONG_ROW = 96;
  const int TEAM_ROW = a.team_row > 0 ? a.team_row : ((1024 / (4 * RPW)) > LONG_ROW ? 1024 / (4 * RPW) : LONG_ROW);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // dynamic shared memory: [TILES: TS*BW*KC floats] [sbuf: NW * 2 buffers * (32 offsets + 32 values)] [team_acc: NW * KC floats]
  float* stile = reinterpret_cast<float*>(smem_raw);  // [TS][BW][KC]
  uint32_t* sbuf = reinterpret_cast<uint32_t*>(reinterpret_cast<float*>(smem_raw) + (TILES ? (size_t)a.TS * a.BW * KC : 0));
  __shared__ int P[BH + 1], RS[BH];
  __shared__ int next_row, next_row2, next_team;
  __shared__ int long_list[BH], n_long;  // rows of this CTA with >= TEAM_ROW handled nz (any order)
  __shared__ int team_row[WARPS / 4];
  __shared__ uint64_t bar;
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LPR;
  const int w = warp * RPW + sub;
  int p, part_q, split;
  if (a.wl) {
    const int2 e = a.wl[blockIdx.x];
    p = e.x; part_q = e.y & 0xff; split = e.y >> 8;
  } else {
    const int pslot = blockIdx.x / a.split;
    part_q = blockIdx.x % a.split; split = a.split;
    p = plist ? plist[pslot] : pslot;
  }
  if (p < a.p_lo || p >= a.p_hi) return;  // uniform over the CTA
  const int kc0 = blockIdx.y * KC;
  const int cnt0 = a.mcsr_cnt[p], delta = a.mcsr_cnt[p + 1] - cnt0;
  const int ntres = TILES ? min(delta - 1, a.TS) : 0;
  const unsigned k4 = a.k / 4;
  const bool col_ok = kc0 / 4 + sl < a.width / 4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  const int kw = min(KC, a.width - kc0);
  const float4* B4 = reinterpret_cast<const float4*>(a.B) + c4;
  float4* C4 = reinterpret_cast<float4*>(a.C) + c4;
  const int BW = a.BW;
  const int tslot = a.tc_slot ? a.tc_slot[p] : -1;  // position of the panel's tensor-window product, -1 = none

  if (TILES && ntres > 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 32);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
      const int* list = a.mcsr_list + (size_t)(cnt0 - p) * BW;
      const int nslot = ntres * BW;
      uint32_t bytes = 0;
      for (int i = lane; i < nslot; i += 32) bytes += list[i] >= 0 ? (uint32_t)kw * 4u : 0u;
      mbar_expect_tx_arrive(&bar, bytes);
      for (int i = lane; i < nslot; i += 32) {
        const int c = list[i];
        if (c >= 0) tma_bulk_g2s(stile + (size_t)i * KC, a.B + (size_t)c * a.k + kc0, (uint32_t)kw * 4u, &bar);
      }
    }
  }
  for (int r = threadIdx.x; r < BH; r += blockDim.x) {
    const int base = cnt0 * BH + r * delta;
    const int rs = a.mcsr_e[base], re = a.mcsr_e[base + delta];
    int nch = 0;
    if (a.spec_off) nch = a.spec_off[p * BH + r + 1] - a.spec_off[p * BH + r];
    RS[r] = rs;
    P[r + 1] = re - rs - nch * STHRESHOLD;  // handled length; turned into a prefix below when the panel is split
  }
  __syncthreads();
  int rlo = 0, rhi = BH;
  if (split > 1) {  // this CTA's share of the panel: rows cut where the nz stream crosses q/split of its length
    if (warp == 0) {
      int v0 = P[4 * lane + 1], v1 = P[4 * lane + 2], v2 = P[4 * lane + 3], v3 = P[4 * lane + 4];
      int s4 = v0 + v1 + v2 + v3, inc = s4;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      int ex = inc - s4;
      if (lane == 0) P[0] = 0;
      P[4 * lane + 1] = ex + v0; P[4 * lane + 2] = ex + v0 + v1; P[4 * lane + 3] = ex + v0 + v1 + v2; P[4 * lane + 4] = ex + s4;
    }
    __syncthreads();
    const int Tall = P[BH];
    const int tlo = (int)((long long)Tall * part_q / split), thi = (int)((long long)Tall * (part_q + 1) / split);
    int lo = 0, hi = BH;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid] < tlo) lo = mid + 1; else hi = mid; }
    rlo = part_q == 0 ? 0 : lo;
    lo = 0; hi = BH;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid] < thi) lo = mid + 1; else hi = mid; }
    rhi = part_q == split - 1 ? BH : lo;
  }
  if (threadIdx.x == 0) { next_team = 0; next_row = rlo; next_row2 = rlo; n_long = 0; }
  __syncthreads();
  for (int r = rlo + threadIdx.x; r < rhi; r += blockDim.x)
    if ((split > 1 ? P[r + 1] - P[r] : P[r + 1]) >= TEAM_ROW) long_list[atomicAdd(&n_long, 1)] = r;
  __syncthreads();
  if (TILES && ntres > 0) mbar_wait(&bar, 0);

  const float4* S4 = reinterpret_cast<const float4*>(stile) + sl;
  const Policies pol = make_policies(a.hints);
  auto bload = [&](unsigned o) -> float4 {
    if (TILES && (o & 0x80000000u)) return S4[o & 0x7fffffffu];
    return ldg4_keep_at(B4, o, pol.keep);
  };
  // A worker stages CH = 32 nz per chunk whatever its width: at k = 32 (8 lanes per row) a lane brings in four nz, so the
  // chunk's fixed cost (metadata request, staging, the tile barrier) is paid per 32 nz and not per 8
  constexpr int CH = 32, EPL = CH / LPR;
  uint32_t* sb0 = sbuf + (size_t)w * 4 * CH;  // [2 buffers][offsets CH | values CH]
  int buf = 0;
  // one group: M*4 B rows requested back to back, then their FMAs in nz order
  auto group = [&](auto mtag, const uint32_t* so, float4& acc) {
    constexpr int M = decltype(mtag)::value;
    float4 b[4 * M];
#pragma unroll
    for (int q = 0; q < M; ++q) {
      const uint4 o = *reinterpret_cast<const uint4*>(so + 4 * q);
      b[4 * q] = bload(o.x); b[4 * q + 1] = bload(o.y); b[4 * q + 2] = bload(o.z); b[4 * q + 3] = bload(o.w);
    }
#pragma unroll
    for (int q = 0; q < M; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(so + CH + 4 * q);
      fma4(acc, v.x, b[4 * q]); fma4(acc, v.y, b[4 * q + 1]); fma4(acc, v.z, b[4 * q + 2]); fma4(acc, v.w, b[4 * q + 3]);
    }
  };
  // Three passes over the panel's rows.  The longest rows (>= TEAM_ROW nz: 256 at k = 128, 96 at k = 32) first, each by a
  // TEAM of four warps (U = 4 * RPW workers): worker u takes the row's chunks u, u + U, ..., the partial sums meet in shared
  // memory and worker 0 adds them in worker order and stores the row.  One worker alone streams a 500-nz row as 16 chunks of
  // dependent memory round trips (~90 us, longer than the rest of its panel takes), which made those rows the critical path
  // of their CTAs (yelp-shape k=32: 0.44 -> 0.28 ms).  Then single workers take whole rows from shared counters, rows of
  // >= LONG_ROW nz before the short ones so that none starts when the others run out of rows.
  constexpr int GM = (G < CH ? G : CH) / 4;  // quads per full group
  constexpr int TW = 4, U = TW * RPW;
  static_assert(WARPS % TW == 0, "teams of four warps");
  int pass = 0;
  auto grab = [&](int& Lr) -> int {  // next row of this pass (its handled length in Lr), -1 when the panel is done
    for (;;) {
      int r = 0;
      if (sl == 0) r = atomicAdd(pass == 0 ? &next_row : &next_row2, 1);
      r = tile.shfl(r, 0);
      if (r >= rhi) {
        if (pass == 1) return -1;
        pass = 1;
        continue;
      }
      Lr = split > 1 ? P[r + 1] - P[r] : P[r + 1];
      if (Lr < TEAM_ROW && (Lr >= LONG_ROW) == (pass == 0)) return r;
    }
  };
  // stage the chunk of row rr that starts at nz i: lane sl brings in nz sl, sl + LPR, ... ; entries past the end repeat the
  // chunk's last nz with value 0, so every staged entry is a valid B row and the tail group is padded without branches
  auto stage = [&](int rr, int i, int cnt, uint32_t* sb) {
    int cc[EPL];
    float vv[EPL];
    const int e0 = RS[rr] + i;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int e = e0 + min(sl + j * LPR, cnt - 1);
      cc[j] = ldg_stream(a.csr_e + e, pol.stream);
      vv[j] = ldg_stream(a.csr_ev + e, pol.stream);
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      unsigned off = (unsigned)cc[j] * k4;
      if (TILES && ntres > 0) {
        const int e = e0 + min(sl + j * LPR, cnt - 1);
        const int base = cnt0 * BH + rr * delta;
        if (e < a.mcsr_e[base + ntres]) {
          int g = 0;
          for (int b = 1; b < ntres; ++b) g += e >= a.mcsr_e[base + b];
          off = 0x80000000u | (unsigned)((g * BW + (cc[j] & (BW - 1))) * (KC / 4));
        }
      }
      sb[sl + j * LPR] = off;
      sb[CH + sl + j * LPR] = sl + j * LPR < cnt ? __float_as_uint(vv[j]) : 0u;
    }
  };
  // acc += chunks first, first + stride, ... of row r (handled length L)
  auto run_chunks = [&](int r, int L, int first, int stride, float4& acc) {
    for (int i = first * CH; i < L; i += stride * CH) {
      const int cnt = min(CH, L - i);
      uint32_t* sb = sb0 + buf * 2 * CH;
      stage(r, i, cnt, sb);
      tile.sync();
      int q4 = (cnt + 3) >> 2;  // quads of this chunk, the last one padded
      const uint32_t* so = sb;
      for (; q4 >= GM; q4 -= GM, so += 4 * GM) group(std::integral_constant<int, GM>{}, so, acc);
      if (GM > 1) {
        if (GM > 2 && q4 >= 3) group(std::integral_constant<int, (GM > 2 ? 3 : 1)>{}, so, acc);
        else if (q4 == 2) group(std::integral_constant<int, 2>{}, so, acc);
        else if (q4 == 1) group(std::integral_constant<int, 1>{}, so, acc);
      }
      buf ^= 1;
    }
  };
  // acc (+ the row's 512-chunk partials, in chunk order) -> C, every element written once
  auto finish_row = [&](int r, float4 acc) {
    const int row = p * BH + r;
    if (a.spec_off) {
      const int so = a.spec_off[row], nch = a.spec_off[row + 1] - so;
      const float4* P4 = reinterpret_cast<const float4*>(a.partial) + (size_t)so * k4 + c4;
      for (int c = 0; c < nch; ++c) {
        const float4 pp = ldg4_stream(P4 + (size_t)c * k4, pol.stream);
        acc.x += pp.x; acc.y += pp.y; acc.z += pp.z; acc.w += pp.w;
      }
    }
    if (col_ok && row < a.nloc) stg4_stream(C4 + (size_t)row * k4, acc, pol.stream);
  };
  // the accumulator starts from the row's tensor-window product (requested first, needed by the first FMA: its latency
  // runs under the metadata and B requests instead of ending the row)
  auto start_acc = [&](int r) -> float4 {
    if (tslot >= 0) return ldg4_stream(reinterpret_cast<const float4*>(a.tc_out) + ((size_t)tslot * BH + r) * k4 + c4, pol.stream);
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };

  if (n_long > 0) {  // uniform over the CTA
    const int team = warp / TW, u = (warp % TW) * RPW + sub;
    float4* tacc = reinterpret_cast<float4*>(sbuf + (size_t)WARPS * RPW * 4 * CH) + (size_t)team * U * LPR;  // [U][LPR]
    const int nl = n_long;
    for (;;) {
      if (warp % TW == 0 && lane == 0) team_row[team] = atomicAdd(&next_team, 1);
      __syncwarp();  // the team barrier is taken by whole warps
      asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TW * 32) : "memory");
      const int li = team_row[team];
      if (li >= nl) break;
      const int r = long_list[li];
      const int L = split > 1 ? P[r + 1] - P[r] : P[r + 1];
      float4 acc = u == 0 ? start_acc(r) : make_float4(0.f, 0.f, 0.f, 0.f);
      run_chunks(r, L, u, U, acc);
      if (u != 0) tacc[u * LPR + sl] = acc;
      __syncwarp();  // the team barrier is taken by whole warps
      asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TW * 32) : "memory");
      if (u == 0) {
        for (int q = 1; q < U; ++q) {
          const float4 x = tacc[q * LPR + sl];
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        finish_row(r, acc);
      }
      // tacc and team_row are rewritten only after the next bar.sync pair, which worker 0 joins after its reads
    }
  }
  // (tried: taking the next row and requesting its first metadata before this row's B requests go out -- 0.597 vs
  // 0.564 ms on Reddit-shape k=128 at the same 40 registers; rows of different workers already overlap)
  int L = 0;
  for (int r = grab(L); r >= 0; r = grab(L)) {
    float4 acc = start_acc(r);
    run_chunks(r, L, 0, 1, acc);
    finish_row(r, acc);
  }
}

template <int KC>
__global__ void __launch_bounds__(256) k_spmm_csr_vec(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                                                      const float* __restrict__ val, int nrows,
                                                      const float* __restrict__ B, float* __restrict__ C, int k) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR;
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + ((threadIdx.x & 31) / LPR);
  const int c4 = blockIdx.y * LPR + sl;
  const unsigned k4 = k / 4;
  if (row >= nrows) return;
  const bool col_ok = c4 < (int)k4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const Policies pol = make_policies(1);
  accum_global<LPR>(tile, (int)rowptr[row], (int)rowptr[row + 1], reinterpret_cast<const int*>(col), val,
                    reinterpret_cast<const float4*>(B) + (col_ok ? c4 : 0), k4, acc, pol);
  if (col_ok) reinterpret_cast<float4*>(C)[(size_t)row * k4 + c4] = acc;
}

__global__ void __launch_bounds__(256) k_spmm_csr_scalar(const uint32_t* __restrict__ rowptr,
                                                         const uint32_t* __restrict__ col,
                                                         const float* __restrict__ val, int nrows,
                                                         const float* __restrict__ B, float* __restrict__ C, int k) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.y * 32 + lane;
  if (row >= nrows) return;
  const int lo = (int)rowptr[row], hi = (int)rowptr[row + 1];
  float acc = 0.f;
  for (int e0 = lo; e0 < hi; e0 += 32) {
    const int e = e0 + lane;
    size_t off = 0;
    float v = 0.f;
    if (e < hi) { off = (size_t)col[e] * (size_t)k; v = val[e]; }
    const int cnt = min(32, hi - e0);
    for (int q = 0; q < cnt; ++q) {
      const size_t o = __shfl_sync(0xffffffffu, off, q);
      const float vv = __shfl_sync(0xffffffffu, v, q);
      if (j < k) acc = fmaf(vv, __ldg(B + o + j), acc);
    }
  }
  if (j < k) C[(size_t)row * k + j] = acc;
}

// shadow_b[r,:] = B[map[r],:] (gather) or out[map[r],:] = in[r,:] (scatter)  -- flex.cu:276-306
__global__ void k_permute_rows(const int32_t* __restrict__ map, long long n, int k, const float* __restrict__ src,
                               float* __restrict__ dst, bool scatter) {
  const int vec = (k % 4 == 0) ? 4 : 1;
  const long long per_row = k / vec;
  const long long total = n * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row, c = i % per_row;
    const long long rs = scatter ? r : map[r], rd = scatter ? map[r] : r;
    if (vec == 4)
      reinterpret_cast<float4*>(dst)[rd * per_row + c] = __ldg(reinterpret_cast<const float4*>(src) + rs * per_row + c);
    else
      dst[rd * k + c] = __ldg(src + rs * k + c);
  }
}

int pick_kc(int k) { return k <= 32 ? 32 : (k <= 64 ? 64 : 128); }

}  // namespace

namespace fx {

int spmm_csr(const uint32_t* rowptr, const uint32_t* col, const float* val, int64_t nrows, const float* B, float* C,
             int k, cudaStream_t s) {
  if (nrows == 0) return FX_OK;
  if (k % 4 != 0) {
    dim3 grid(ceil_div(nrows, 8), ceil_div(k, 32));
    k_spmm_csr_scalar<<<grid, 256, 0, s>>>(rowptr, col, val, (int)nrows, B, C, k);
    FX_LAUNCH_CHECK();
    return FX_OK;
  }
  const int KC = pick_kc(k);
  dim3 grid(ceil_div(nrows, 8 * (32 / (KC / 4))), ceil_div(k, KC));
  if (KC == 32) k_spmm_csr_vec<32><<<grid, 256, 0, s>>>(rowptr, col, val, (int)nrows, B, C, k);
  else if (KC == 64) k_spmm_csr_vec<64><<<grid, 256, 0, s>>>(rowptr, col, val, (int)nrows, B, C, k);
  else k_spmm_csr_vec<128><<<grid, 256, 0, s>>>(rowptr, col, val, (int)nrows, B, C, k);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

// Launch configurations of the row kernel: (warps per CTA, CTAs per SM the register cap is set for, rows in flight per group)
static int rows_cfg() {
  static int c = -1;
  if (c < 0) {
    const char* e = getenv("FLEX_ROWS_CFG");
    c = e ? atoi(e) : 0;
    if (c < 0 || c > 3) c = 0;
  }
  return c;
}

template <int KC, int WARPS, int MINB, int G, bool TILES>
static int launch_one(PanelArgs a, const int* plist, int npan, const int2* wl, int nwl, int kchunks, size_t tile_smem,
                      cudaStream_t s) {
  constexpr int NW = WARPS * (32 / (KC / 4));
  // + per-worker offset / value staging (2 x 32 nz) + the teams' partial sums of long rows (one row of C per worker)
  const size_t smem = tile_smem + (size_t)NW * 4 * 32 * sizeof(uint32_t) + (size_t)NW * KC * sizeof(float);
  static const bool no_wl = getenv("FLEX_NO_WORKLIST") != nullptr;
  // a uniform split (few panels, or FLEX_SPLIT) keeps the blockIdx mapping; otherwise the build's work list
  a.wl = (a.split == 1 && wl && nwl > 0 && !no_wl) ? wl : nullptr;
  dim3 grid(a.wl ? nwl : npan * a.split, kchunks);
  static SmemAttr attr;
  if (int rc = attr.ensure(k_spmm_rows<KC, WARPS, TILES, MINB, G>, smem)) return rc;
  k_spmm_rows<KC, WARPS, TILES, MINB, G><<<grid, WARPS * 32, smem, s>>>(a, plist);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

template <int KC, int WARPS, int MINB, int G>
static int launch_panels(const fx_aspt_dev& d, const PanelArgs& a, int kchunks, cudaStream_t s) {
  // Shared-memory staging of dense tiles only pays when a large share of the nz sits in them: the
  // tile memory (64 KB per tile at k=128) comes out of the SM's L1, which serves the sparse nz.
  // Measured on Reddit-shape (6 % of nz in tiles): 0.855 ms with the tiled CTAs, 0.759 ms without.
  const char* tiles_env = getenv("FLEX_TILES");  // "0" never, "1" always, unset = by density
  const double dense_frac = d.ne > 0 ? (double)(d.ne - d.S1) / d.ne : 0.0;
  const bool use_tiles = (tiles_env ? atoi(tiles_env) != 0 : dense_frac >= 0.25) && a.tiles_ok;
  if (!use_tiles || d.n_tiled == 0)  // every panel through the L1 path (dense groups are ordinary nz there)
    return launch_one<KC, WARPS, MINB, G, false>(a, nullptr, d.npanel, d.wl_all, d.n_wl_all, kchunks, 0, s);
  if (d.n_plain > 0) {
    const int rc = launch_one<KC, WARPS, MINB, G, false>(a, d.n_tiled ? d.plist_plain : nullptr, d.n_plain, d.wl_plain, d.n_wl_plain, kchunks, 0, s);
    if (rc != FX_OK) return rc;
  }
  if (d.n_tiled > 0)
    return launch_one<KC, WARPS, MINB, G, true>(a, d.plist_tiled, d.n_tiled, d.wl_tiled, d.n_wl_tiled, kchunks, (size_t)a.TS * a.BW * KC * sizeof(float), s);
  return FX_OK;
}

// fx_spmm_kernel_times: events recorded between the kernels of one SpMM (null outside that call)
static thread_local cudaEvent_t* g_prof = nullptr;
static int prof_mark(int i, cudaStream_t s) {
  if (g_prof) FX_CUDA(cudaEventRecord(g_prof[i], s));
  return FX_OK;
}

template <int KC>
static int launch_aspt(const fx_tiles* t, const PanelArgs& a, int special_p, cudaStream_t s) {
  const fx_aspt_dev& d = t->aspt;
  const int kchunks = ceil_div(a.width, KC);
  if (special_p > 0) {
    constexpr int RPW = 32 / (KC / 4);
    static const int cta_thr = getenv("FLEX_SPECIAL_CTA") ? atoi(getenv("FLEX_SPECIAL_CTA")) : 0x7fffffff;
    if (special_p < cta_thr) {  // default: always (0.734 vs 0.758 ms on Reddit-shape, 0.126 vs 0.184 ms on a 1/8 shard)
      dim3 g(special_p, kchunks);
      k_spmm_special_cta<KC><<<g, PANEL_WARPS * 32, 0, s>>>(a, d.special, d.special2, d.partial);
    } else {
      dim3 g(ceil_div(special_p, PANEL_WARPS * RPW), kchunks);
      k_spmm_special<KC><<<g, PANEL_WARPS * 32, 0, s>>>(a, d.special, d.special2, special_p, d.partial);
    }
    FX_LAUNCH_CHECK();
  }
  if (int rc = prof_mark(2, s)) return rc;
  // Measured on Reddit-shape k=128 (ms per SpMM): 0.569 default; 0.576 <16,2,16>; 0.577 <16,2,8>; 0.582 <8,6,8>; fewer, fatter
  // warps lose (<8,4,16> 0.589, <8,3,12> 0.617, <8,3,16> 0.631, <8,2,16> at 101 registers 0.689): the per-row latency chain
  // (grab, metadata, B rounds, window product, store) is covered by warps, not by loads in flight per warp
  switch (rows_cfg()) {
    case 1: return launch_panels<KC, 16, 2, 16>(d, a, kchunks, s);  // 64 registers, 32 warps per SM
    case 2: return launch_panels<KC, 16, 2, 8>(d, a, kchunks, s);
    case 3: return launch_panels<KC, 8, 6, 8>(d, a, kchunks, s);    // 40 registers, 48 warps per SM in small CTAs
    default: return launch_panels<KC, 16, 3, 8>(d, a, kchunks, s);  // 40 registers, 48 warps per SM
  }
}

template <int N>
static int launch_tc(const fxtc::TcArgs& ta, int ntc, cudaStream_t s) {
  const size_t smem = fxtc::tc_smem_bytes<N>(ta.W);
  static SmemAttr attr;
  if (int rc = attr.ensure(fxtc::k_spmm_tc<N>, smem)) return rc;
  dim3 grid(ntc, ceil_div(ta.width, N));
  fxtc::k_spmm_tc<N><<<grid, 256, smem, s>>>(ta);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

int spmm_aspt(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s, int width, int p_lo, int p_hi) {
  if (width <= 0) width = k;
  const fx_aspt_dev& d = t->aspt;
  if (d.npanel == 0) return FX_OK;
  if (p_hi < 0 || p_hi > d.npanel) p_hi = d.npanel;
  if (p_lo < 0) p_lo = 0;
  if (p_lo >= p_hi) return FX_OK;
  // float4 offsets of B rows are 32-bit in the kernels (both entry points, fx_spmm and fx_spmm_host, come through here)
  FX_REQUIRE((int64_t)t->mat->n * k / 4 < (1ll << 32), FX_ERR_UNSUPPORTED, "n*k too large for 32-bit float4 offsets");
  const int KC = pick_kc(width);
  PanelArgs a;
  a.tc_out = nullptr; a.tc_slot = nullptr; a.wl = nullptr;
  a.p_lo = p_lo; a.p_hi = p_hi;
  static const int hints = getenv("FLEX_HINTS") ? atoi(getenv("FLEX_HINTS")) : 1;
  a.hints = hints;
  static const int team_row = getenv("FLEX_TEAM_ROW") ? atoi(getenv("FLEX_TEAM_ROW")) : 0;
  a.team_row = team_row;
  a.tiles_ok = (int64_t)t->mat->n * k / 4 < (1ll << 31);
  if (int rc = prof_mark(0, s)) return rc;
  if (t->format == FX_FMT_TCW && t->tcw.ntc > 0) {  // tensor windows first; the panel kernel adds them in
    const fx_tcw_dev& w = t->tcw;
    fxtc::TcArgs ta;
    ta.win_cptr = w.win_cptr; ta.win_code = w.win_code; ta.win_val = w.win_val;
    ta.tc_panels = w.tc_panels; ta.tc_cols = w.tc_cols; ta.tc_ncol = w.tc_ncol;
    ta.B = B; ta.out = w.tc_out; ta.k = k; ta.width = width; ta.W = w.W;
    ta.hints = getenv("FLEX_TC_HINTS") ? atoi(getenv("FLEX_TC_HINTS")) : hints;
    ta.p_lo = p_lo; ta.p_hi = p_hi;
    const int rc = KC == 32 ? launch_tc<32>(ta, w.ntc, s) : (KC == 64 ? launch_tc<64>(ta, w.ntc, s) : launch_tc<128>(ta, w.ntc, s));
    if (rc != FX_OK) return rc;
    a.tc_out = w.tc_out; a.tc_slot = w.tc_slot;
  }
  if (int rc = prof_mark(1, s)) return rc;
  a.mcsr_cnt = d.mcsr_cnt; a.mcsr_e = d.mcsr_e_use; a.mcsr_list = d.mcsr_list;
  a.csr_e = d.csr_e_use; a.csr_ev = d.csr_ev_use;
  static const bool no_special = getenv("FLEX_NO_SPECIAL") != nullptr;
  const int special_p = no_special ? 0 : d.special_p;
  // the 512-chunk partial sums live in a scratch sized for the build's k (fx_aspt_build.cu: special_cap * k floats): a
  // wider call would write past it.  fx_spmm routes such calls to the CSR kernel; this is the backstop.
  FX_REQUIRE((size_t)special_p * (size_t)k <= d.partial_cap_floats, FX_ERR_ARG,
             "fx_spmm: k = %d is larger than the k = %d the tiles were built for", k, t->k);
  a.spec_off = special_p > 0 ? d.spec_off : nullptr;
  static const bool spec_sched = !(getenv("FLEX_SPEC_ORDER") && atoi(getenv("FLEX_SPEC_ORDER")) == 0);
  a.spec_order = spec_sched ? d.spec_order : nullptr;
  a.partial = d.partial;
  a.B = B; a.C = C;
  a.npanel = d.npanel; a.nloc = t->row_end - t->row_begin; a.k = k; a.width = width; a.BW = d.BW;
  {  // fewer panels than SMs: several CTAs per panel.  Otherwise one CTA per work-list entry -- the row kernel
     // balances inside a panel by itself, and more CTAs only repeat its prologue (measured: flickr-shape 0.103 ms
     // unsplit vs 0.121 split in two; 1/8 Reddit-shape shards 0.136 / 0.129 / 0.141 / 0.146 ms for 1 / 2 / 3 / 4)
    const int sm = sm_count_of_current_device();
    const char* e = getenv("FLEX_SPLIT");
    const int kch = ceil_div(width, KC);
    int sp = e ? atoi(e) : (d.npanel * kch >= sm ? 1 : ceil_div(2 * sm, d.npanel * kch));
    a.split = sp < 1 ? 1 : (sp > 8 ? 8 : sp);
  }
  const size_t tile_bytes = (size_t)d.BW * KC * sizeof(float);
  int ts = d.max_tp;
  if (ts > MAX_TS) ts = MAX_TS;
  while (ts > 1 && ts * tile_bytes > 160 * 1024) --ts;  // leave room for the per-worker buffers
  a.TS = ts > 0 ? ts : 1;
  if (d.max_tp == 0) a.TS = 0;
  const int rc = KC == 32 ? launch_aspt<32>(t, a, special_p, s) : (KC == 64 ? launch_aspt<64>(t, a, special_p, s) : launch_aspt<128>(t, a, special_p, s));
  if (rc != FX_OK) return rc;
  return prof_mark(3, s);
}

// One SpMM with events between its kernels: ms[0] = tensor-window kernel, ms[1] = 512-chunk kernel, ms[2] = row kernel,
// ms[3] = the whole step (ASpT / tensor-window formats; a kernel that does not run reports 0).
int spmm_aspt_times(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s, float ms[4]) {
  cudaEvent_t ev[4] = {};
  for (int i = 0; i < 4; ++i) FX_CUDA(cudaEventCreate(&ev[i]));
  g_prof = ev;
  const int rc = spmm_aspt(t, B, C, k, s);
  g_prof = nullptr;
  int out = rc;
  if (rc == FX_OK) {
    if (cudaEventSynchronize(ev[3]) != cudaSuccess) out = FX_ERR_CUDA;
    for (int i = 0; i < 3 && out == FX_OK; ++i)
      if (cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]) != cudaSuccess) out = FX_ERR_CUDA;
    if (out == FX_OK && cudaEventElapsedTime(&ms[3], ev[0], ev[3]) != cudaSuccess) out = FX_ERR_CUDA;
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  if (out == FX_ERR_CUDA && rc == FX_OK) set_error("fx_spmm_kernel_times: event timing failed: %s", cudaGetErrorString(cudaGetLastError()));
  return out;
}

template <int N>
static int launch_gemm(const fxtc::GemmArgs& g, cudaStream_t s) {
  const size_t smem = fxtc::gemm_smem_bytes<N>();
  static SmemAttr attr;
  if (int rc = attr.ensure(fxtc::k_gemm_xw<N>, smem)) return rc;
  dim3 grid(ceil_div(g.rows, fxtc::TC_BH), ceil_div(g.c, N));
  fxtc::k_gemm_xw<N><<<grid, 256, smem, s>>>(g);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

// out[rows x c] = X[rows x k] * W[k x c] on tcgen05 (3xTF32), fx_gemm_kernel.cuh
int gemm_xw(const float* X, const float* W, float* out, int64_t rows, int k, int c, cudaStream_t s) {
  if (rows == 0) return FX_OK;
  FX_REQUIRE(k % 4 == 0 && c % 4 == 0 && k > 0 && c > 0, FX_ERR_UNSUPPORTED, "AXW: k and c must be multiples of 4");
  fxtc::GemmArgs g{X, W, out, (int)rows, k, c};
  const int N = c <= 32 ? 32 : (c <= 64 ? 64 : 128);
  return N == 32 ? launch_gemm<32>(g, s) : (N == 64 ? launch_gemm<64>(g, s) : launch_gemm<128>(g, s));
}

int permute_rows(const int32_t* map, int64_t n, int k, const float* src, float* dst, bool scatter, cudaStream_t s) {
  if (n == 0) return FX_OK;
  const int sm = sm_count_of_current_device();
  k_permute_rows<<<sm * 8, 256, 0, s>>>(map, n, k, src, dst, scatter);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

}  // namespace fx
