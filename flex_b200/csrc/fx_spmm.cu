// fx_spmm.cu -- the SpMM kernels (C = A*B, fp32, B/C row-major) for sm_100a.
//
// Work mapping (contrast: the reference's kernels use one lane per output column, scalar 4-byte
// B loads and fp32 atomics into a pre-zeroed C -- aspt/sspmm_128.cu:321-427,703-827):
//   * a row of C is owned by LPR = KC/4 lanes; each lane keeps one float4 of the row in registers,
//     so a B row is fetched with one 128-bit load per lane (512 B per warp instruction at k=128);
//   * (col,val) pairs are read coalesced, LPR at a time, and broadcast with warp shuffles;
//   * a panel's dense tiles are staged in shared memory once per CTA by the TMA unit
//     (cp.async.bulk global->shared, completion on an mbarrier), one 16-byte-aligned row of B per
//     bulk copy, and read back with 128-bit shared loads;
//   * every C element is written exactly once with a plain 128-bit store (no atomics, no memset):
//     the 512-nz chunks of very long rows (the reference's "special" lists, :1076-1087) are reduced
//     into a scratch buffer by a first kernel and folded in, in a fixed order, by the row's owner.
#include <cooperative_groups.h>

#include "fx_common.cuh"

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

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ void fma4(float4& a, float v, const float4& b) {
  a.x = fmaf(v, b.x, a.x); a.y = fmaf(v, b.y, a.y); a.z = fmaf(v, b.z, a.z); a.w = fmaf(v, b.w, a.w);
}

// acc += sum over nz [lo,hi) of val * B[col,:]  with B read from global memory (L1/L2 path).
// `B4` already points at this lane's float4 of row 0; a row is `k4` float4 wide.
template <int LPR, class Tile>
__device__ __forceinline__ void accum_global(const Tile& tile, int lo, int hi, const int* __restrict__ ce,
                                             const float* __restrict__ cv, const float4* __restrict__ B4,
                                             unsigned k4, float4& acc) {
  const int sl = tile.thread_rank();
  for (int e0 = lo; e0 < hi; e0 += LPR) {
    const int e = e0 + sl;
    unsigned off = 0;
    float v = 0.f;
    if (e < hi) { off = (unsigned)ce[e] * k4; v = cv[e]; }
    const int cnt = min(LPR, hi - e0);
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
      const unsigned o0 = tile.shfl(off, j), o1 = tile.shfl(off, j + 1), o2 = tile.shfl(off, j + 2),
                     o3 = tile.shfl(off, j + 3);
      const float v0 = tile.shfl(v, j), v1 = tile.shfl(v, j + 1), v2 = tile.shfl(v, j + 2), v3 = tile.shfl(v, j + 3);
      const float4 b0 = ldg4(B4 + o0), b1 = ldg4(B4 + o1), b2 = ldg4(B4 + o2), b3 = ldg4(B4 + o3);
      fma4(acc, v0, b0); fma4(acc, v1, b1); fma4(acc, v2, b2); fma4(acc, v3, b3);
    }
    for (; j < cnt; ++j) {
      const unsigned o = tile.shfl(off, j);
      const float vv = tile.shfl(v, j);
      fma4(acc, vv, ldg4(B4 + o));
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
  const float* partial;  // [chunk][k] partial sums of the 512-chunks
  const float* B;
  float* C;
  int npanel, nloc, k, BW, TS;
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
  const bool col_ok = kc0 / 4 + sl < (int)k4;
  const int c4 = col_ok ? kc0 / 4 + sl : 0;
  const int row = special[item], off = special2[item];
  const int p = row / BH, r = row % BH;
  const int cnt0 = a.mcsr_cnt[p], delta = a.mcsr_cnt[p + 1] - cnt0;
  const int lo = a.mcsr_e[cnt0 * BH + (r + 1) * delta - 1] + off;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  accum_global<LPR>(tile, lo, lo + STHRESHOLD, a.csr_e, a.csr_ev, reinterpret_cast<const float4*>(a.B) + c4, k4, acc);
  if (col_ok) reinterpret_cast<float4*>(partial)[(size_t)item * k4 + c4] = acc;
}

// ---- panel kernel: one CTA per 128-row panel -------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(PANEL_WARPS * 32) k_spmm_panel(PanelArgs a) {
  constexpr int LPR = KC / 4, RPW = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stile = reinterpret_cast<float*>(smem_raw);  // [TS][BW][KC]
  __shared__ uint64_t bar;
  auto tile = cg::tiled_partition<LPR>(cg::this_thread_block());
  const int sl = tile.thread_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LPR;
  const int p = blockIdx.x, kc0 = blockIdx.y * KC;
  const int cnt0 = a.mcsr_cnt[p], delta = a.mcsr_cnt[p + 1] - cnt0, tp = delta - 1;
  const unsigned k4 = a.k / 4;
  const int c4 = kc0 / 4 + sl;
  const bool col_ok = c4 < (int)k4;
  const int kw = min(KC, a.k - kc0);  // floats of this k-chunk that exist
  const float4* B4 = reinterpret_cast<const float4*>(a.B) + (col_ok ? c4 : 0);
  const int BW = a.BW, TS = a.TS;

  if (tp > 0 && threadIdx.x == 0) {
    mbar_init(&bar, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tp > 0) __syncthreads();

  const int rounds = tp > 0 ? (tp + TS - 1) / TS : 1;
  for (int rd = 0; rd < rounds; ++rd) {
    const int r0 = rd * TS;
    const int ntile = tp > 0 ? min(TS, tp - r0) : 0;
    if (ntile > 0) {
      if (rd > 0) __syncthreads();  // everybody is done reading the previous round's tiles
      if (warp == 0) {
        // producer: one bulk copy per occupied slot; each lane announces its own byte count first
        const int* list = a.mcsr_list + (size_t)(cnt0 - p + r0) * BW;
        const int nslot = ntile * BW;
        uint32_t bytes = 0;
        for (int i = lane; i < nslot; i += 32) bytes += list[i] >= 0 ? (uint32_t)kw * 4u : 0u;
        mbar_expect_tx_arrive(&bar, bytes);
        for (int i = lane; i < nslot; i += 32) {
          const int c = list[i];
          if (c >= 0) tma_bulk_g2s(stile + (size_t)i * KC, a.B + (size_t)c * a.k + kc0, (uint32_t)kw * 4u, &bar);
        }
      }
      mbar_wait(&bar, rd & 1);
    }
    const bool last = rd == rounds - 1;
    for (int it = 0; it < BH / (PANEL_WARPS * RPW); ++it) {
      const int r = it * (PANEL_WARPS * RPW) + warp * RPW + sub;
      const int row = p * BH + r;
      const int base = cnt0 * BH + r * delta;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rd > 0 && col_ok && row < a.nloc) acc = reinterpret_cast<const float4*>(a.C)[(size_t)row * k4 + c4];
      int sp_lo = 0, sp_hi = 0, nch = 0;
      if (last) {
        sp_lo = a.mcsr_e[base + tp];
        sp_hi = a.mcsr_e[base + delta];
        if (a.spec_off) {
          nch = (sp_hi - sp_lo) / STHRESHOLD;
          if (nch > 0 && col_ok) {  // fold the chunk partials in, in chunk order
            const float4* P4 = reinterpret_cast<const float4*>(a.partial) + (size_t)a.spec_off[row] * k4 + c4;
            for (int c = 0; c < nch; ++c) {
              const float4 pp = P4[(size_t)c * k4];
              acc.x += pp.x; acc.y += pp.y; acc.z += pp.z; acc.w += pp.w;
            }
          }
        }
      }
      if (ntile > 0) {
        // dense groups r0..r0+ntile-1 of this row are one contiguous nz range
        int bnd[MAX_TS + 1];
#pragma unroll
        for (int g = 0; g <= MAX_TS; ++g) bnd[g] = g <= ntile ? a.mcsr_e[base + r0 + g] : 0x7fffffff;
        const int lo = bnd[0], hi = a.mcsr_e[base + r0 + ntile];
        for (int e0 = lo; e0 < hi; e0 += LPR) {
          const int e = e0 + sl;
          unsigned off = 0;
          float v = 0.f;
          if (e < hi) {
            int g = 0;
#pragma unroll
            for (int q = 1; q < MAX_TS; ++q) g += e >= bnd[q];
            off = (unsigned)(g * BW + (a.csr_e[e] & (BW - 1))) * (KC / 4);
            v = a.csr_ev[e];
          }
          const int cnt = min(LPR, hi - e0);
          const float4* S4 = reinterpret_cast<const float4*>(stile) + sl;
          int j = 0;
          for (; j + 4 <= cnt; j += 4) {
            const unsigned o0 = tile.shfl(off, j), o1 = tile.shfl(off, j + 1), o2 = tile.shfl(off, j + 2),
                           o3 = tile.shfl(off, j + 3);
            const float v0 = tile.shfl(v, j), v1 = tile.shfl(v, j + 1), v2 = tile.shfl(v, j + 2),
                        v3 = tile.shfl(v, j + 3);
            const float4 b0 = S4[o0], b1 = S4[o1], b2 = S4[o2], b3 = S4[o3];
            fma4(acc, v0, b0); fma4(acc, v1, b1); fma4(acc, v2, b2); fma4(acc, v3, b3);
          }
          for (; j < cnt; ++j) {
            const unsigned o = tile.shfl(off, j);
            const float vv = tile.shfl(v, j);
            fma4(acc, vv, S4[o]);
          }
        }
      }
      if (last) accum_global<LPR>(tile, sp_lo + nch * STHRESHOLD, sp_hi, a.csr_e, a.csr_ev, B4, k4, acc);
      if (col_ok && row < a.nloc) reinterpret_cast<float4*>(a.C)[(size_t)row * k4 + c4] = acc;
    }
  }
}

// ---- plain CSR kernels (FX_FMT_CSR; also the fallback for k not divisible by 4) --------------
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
  accum_global<LPR>(tile, (int)rowptr[row], (int)rowptr[row + 1], reinterpret_cast<const int*>(col), val,
                    reinterpret_cast<const float4*>(B) + (col_ok ? c4 : 0), k4, acc);
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

template <int KC>
static int launch_aspt(const fx_tiles* t, const PanelArgs& a, int special_p, cudaStream_t s) {
  const fx_aspt_dev& d = t->aspt;
  const int kchunks = ceil_div(a.k, KC);
  if (special_p > 0) {
    constexpr int RPW = 32 / (KC / 4);
    dim3 g(ceil_div(special_p, PANEL_WARPS * RPW), kchunks);
    k_spmm_special<KC><<<g, PANEL_WARPS * 32, 0, s>>>(a, d.special, d.special2, special_p, d.partial);
    FX_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)a.TS * a.BW * KC * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FX_CUDA(cudaFuncSetAttribute(k_spmm_panel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid(d.npanel, kchunks);
  k_spmm_panel<KC><<<grid, PANEL_WARPS * 32, smem, s>>>(a);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

int spmm_aspt(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s) {
  const fx_aspt_dev& d = t->aspt;
  if (d.npanel == 0) return FX_OK;
  const int KC = pick_kc(k);
  PanelArgs a;
  a.mcsr_cnt = d.mcsr_cnt; a.mcsr_e = d.mcsr_e_use; a.mcsr_list = d.mcsr_list;
  a.csr_e = d.csr_e_use; a.csr_ev = d.csr_ev_use;
  a.spec_off = d.special_p > 0 ? d.spec_off : nullptr;
  a.partial = d.partial;
  a.B = B; a.C = C;
  a.npanel = d.npanel; a.nloc = t->row_end - t->row_begin; a.k = k; a.BW = d.BW;
  const size_t tile_bytes = (size_t)d.BW * KC * sizeof(float);
  int ts = d.max_tp;
  if (ts > MAX_TS) ts = MAX_TS;
  while (ts > 1 && ts * tile_bytes > 200 * 1024) --ts;
  a.TS = ts > 0 ? ts : 1;
  if (d.max_tp == 0) a.TS = 0;
  if (KC == 32) return launch_aspt<32>(t, a, d.special_p, s);
  if (KC == 64) return launch_aspt<64>(t, a, d.special_p, s);
  return launch_aspt<128>(t, a, d.special_p, s);
}

int permute_rows(const int32_t* map, int64_t n, int k, const float* src, float* dst, bool scatter, cudaStream_t s) {
  if (n == 0) return FX_OK;
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  k_permute_rows<<<sm * 8, 256, 0, s>>>(map, n, k, src, dst, scatter);
  FX_LAUNCH_CHECK();
  return FX_OK;
}

}  // namespace fx
