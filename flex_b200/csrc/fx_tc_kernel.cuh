// fx_tc_kernel.cuh -- tensor-core path for panels with a genuinely dense column window (sm_100a).
//
// For a 128-row panel p and its densest 256-column window [win, win+256), D_p = A_p[:, win..] * B[win.., :]
// is a 128 x 256 x N contraction.  It runs on the 5th-generation tensor cores: tcgen05.mma
// (kind::tf32, cta_group::1, M=128, N, K=8 per instruction) issued by one thread, operands in shared
// memory described by UMMA descriptors, the 128 x N fp32 accumulator in TMEM, read back with
// tcgen05.ld.  fp32 accuracy is kept with the 3xTF32 split: x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi);  A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi  (the dropped lo*lo term is 2^-22 relative).
//
// Shared-memory operand layouts (no swizzle, "interleaved" canonical UMMA layouts, cf. CUTLASS
// cute/atom/mma_traits_sm100.hpp:165-199):
//   A chunk [128 rows x 64 k], K-major : core matrix = 8 rows x 16 B;  byte(r,kk) =
//        (r/8)*SBO_A + (kk/4)*128 + (r%8)*16 + (kk%4)*4,  LBO_A = 128, SBO_A = 2048
//   B chunk [N x 64 k], K-major (transposed while staging): the same formula with r = feature n.
//        MN-major tf32 operands only work with the 128B_BASE32B swizzle on this part (probed with
//        scripts/tc_probe.cu: every other layout type reads zeros), so B is transposed instead.
// One K=8 step spans two K-adjacent core matrices of either operand (start += 256 B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxtc {

constexpr int TC_BH = 128;   // panel height
constexpr int TC_W = 256;    // window width
constexpr int TC_KCH = 64;   // columns of the window staged per step

struct TcArgs {
  const unsigned* col;   // raw CSR nz arrays of the shard
  const float* val;
  const int* tc_panels;  // [ntc] panels that take the tensor path
  const int* tc_win;     // [npanel] first column of the panel's window
  const int* tc_lo;      // [rows] nz range of each row inside its panel's window
  const int* tc_hi;
  const float* B;        // [ncols x k]
  float* out;            // [ntc][128][k]
  int k, ncols;
};

#ifdef FX_TC_DEBUG
__device__ int g_tc_sleep;
#endif
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp:98-113): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
  // version=1 [46,48), base_offset 0, lbo_mode 0, layout_type 0 (no swizzle) [61,64)
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

template <int N>
__device__ __forceinline__ uint32_t make_idesc() {
  // InstrDescriptor (mma_sm100_desc.hpp:412-434): D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10),
  // A and B K-major (0<<15, 0<<16), N>>3 at [17,23), M>>4 at [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// grid = (ntc, k / N), block = 256.  Dynamic shared memory: 2*128*64*4 + 2*64*N*4 bytes.
template <int N>
__global__ void __launch_bounds__(256, 1) k_spmm_tc(TcArgs a) {
  constexpr int TM_COLS = N < 32 ? 32 : N;
  constexpr uint32_t LBO_A = 128, SBO_A = (TC_KCH / 4) * 128, LBO_B = 128, SBO_B = (TC_KCH / 4) * 128;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  float* Ahi = reinterpret_cast<float*>(tc_smem);
  float* Alo = Ahi + TC_BH * TC_KCH;
  float* Bhi = Alo + TC_BH * TC_KCH;
  float* Blo = Bhi + TC_KCH * N;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.x, panel = a.tc_panels[slot], n0 = blockIdx.y * N;
  const int win = a.tc_win[panel];

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc<N>();
#ifdef FX_TC_DEBUG
  if (warp < 4) {  // sentinel 7.0 in every accumulator cell
    for (int c0 = 0; c0 < N; c0 += 8) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      const uint32_t s7 = __float_as_uint(7.0f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(s7) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) printf("tmem base = 0x%08x idesc = 0x%08x panel %d win %d\n", tmem, idesc, panel, win);
#endif

  // each of the first 128 threads walks its row's window nz once, left to right
  const int row = panel * TC_BH + tid;
  int cur = 0, row_hi = 0;
  if (tid < TC_BH) { cur = a.tc_lo[row]; row_hi = a.tc_hi[row]; }
  uint32_t phase = 0;

  for (int ch = 0; ch < TC_W / TC_KCH; ++ch) {
    const int c0 = win + ch * TC_KCH;
    // zero both A tiles (contiguous)
    for (int i = tid; i < 2 * TC_BH * TC_KCH / 4; i += 256) reinterpret_cast<float4*>(Ahi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // stage, split and transpose the B chunk (rows c0..c0+63 of B, columns n0..n0+N) to K-major: a warp reads
    // 128-byte runs of four consecutive B rows, each lane keeps one feature and stores 4 k as one 16-byte word
    for (int idx = tid; idx < (TC_KCH / 4) * N; idx += 256) {
      const int n = idx % N, kq = idx / N;
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gcol = c0 + kq * 4 + j;
        x[j] = gcol < a.ncols ? __ldg(a.B + (size_t)gcol * a.k + n0 + n) : 0.f;
      }
      float4 h, l;
      h.x = to_tf32(x[0]); h.y = to_tf32(x[1]); h.z = to_tf32(x[2]); h.w = to_tf32(x[3]);
      l.x = to_tf32(x[0] - h.x); l.y = to_tf32(x[1] - h.y); l.z = to_tf32(x[2] - h.z); l.w = to_tf32(x[3] - h.w);
      const uint32_t off = (uint32_t)(n >> 3) * SBO_B + (uint32_t)kq * LBO_B + (uint32_t)(n & 7) * 16u;
      *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Bhi) + off) = h;
      *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Blo) + off) = l;
    }
    __syncthreads();  // zeros are in place before the scatter
    if (tid < TC_BH) {
      const int cend = c0 + TC_KCH;
      while (cur < row_hi) {
        const int c = (int)a.col[cur];
        if (c >= cend) break;
        const int kk = c - c0;
        const float v = a.val[cur];
        const float h = to_tf32(v), l = to_tf32(v - h);
        const uint32_t off = (uint32_t)(tid >> 3) * SBO_A + (uint32_t)(kk >> 2) * LBO_A + (uint32_t)(tid & 7) * 16u + (uint32_t)(kk & 3) * 4u;
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ahi) + off) = h;
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Alo) + off) = l;
        ++cur;
      }
    }
    // generic-proxy writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
#ifdef FX_TC_DEBUG
    if (a.ncols < 0) continue;  // debug: no MMA at all -> the sentinel must come back
    if (g_tc_sleep & 4) {  // every operand word = 1.0 whatever the layout
      for (int i = tid; i < (2 * TC_BH * TC_KCH + 2 * TC_KCH * N); i += 256) Ahi[i] = 1.0f;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    if ((g_tc_sleep & 8) && tid == 0 && ch == 0 && blockIdx.x == 0 && blockIdx.y == 0)
      printf("Ahi[0..3] %g %g %g %g  Ahi[32] %g Bhi[0..3] %g %g %g %g saddr A %x B %x\n", Ahi[0], Ahi[1], Ahi[2], Ahi[3], Ahi[32], Bhi[0],
             Bhi[1], Bhi[2], Bhi[3], smem_addr(Ahi), smem_addr(Bhi));
    if (g_tc_sleep & 2) {  // a single MMA per chunk, no accumulation
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma_tf32(tmem, make_desc(smem_addr(Ahi), LBO_A, SBO_A), make_desc(smem_addr(Bhi), LBO_B, SBO_B), idesc, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
      }
      mbar_wait_parity(&bar, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      continue;
    }
#endif
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ahi = smem_addr(Ahi), alo = smem_addr(Alo), bhi = smem_addr(Bhi), blo = smem_addr(Blo);
#pragma unroll
      for (int ks = 0; ks < TC_KCH / 8; ++ks) {
        const uint64_t dah = make_desc(ahi + ks * 2 * LBO_A, LBO_A, SBO_A), dal = make_desc(alo + ks * 2 * LBO_A, LBO_A, SBO_A);
        const uint64_t dbh = make_desc(bhi + ks * 2 * LBO_B, LBO_B, SBO_B), dbl = make_desc(blo + ks * 2 * LBO_B, LBO_B, SBO_B);
        mma_tf32(tmem, dah, dbh, idesc, (ch | ks) ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dal, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    }
#ifdef FX_TC_DEBUG
    if (g_tc_sleep & 1) __nanosleep(200000);
#endif
    mbar_wait_parity(&bar, phase);  // MMAs of this chunk are done: the tiles may be overwritten
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // epilogue: TMEM lane = panel row; warps 0..3 own lanes 32w..32w+31
  if (warp < 4) {
    const int r = warp * 32 + lane;
    float* dst = a.out + ((size_t)slot * TC_BH + r) * a.k + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c0 + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

template <int N>
inline size_t tc_smem_bytes() { return (size_t)2 * TC_BH * TC_KCH * 4 + (size_t)2 * TC_KCH * N * 4; }

}  // namespace fxtc
