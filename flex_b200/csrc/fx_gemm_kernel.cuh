// fx_gemm_kernel.cuh -- the dense factor of the reference's AXW experiment (cusp.cu:3-208, main.cu:22-79: B = X*W by
// cublasSgemm then C = A*B by cuSPARSE, and the other order; dead code there) on the 5th-generation tensor cores.
//
// out[rows x c] = X[rows x k] * W[k x c], all row-major fp32.  One CTA per 128 rows (x one per N output features): for every
// 32-wide slice of k the X tile goes into the K-major A operand as it lies (a row of X IS k-contiguous), the 32 rows of W of the
// slice are transposed into the K-major B operand exactly as the window kernel stages rows of B (fx_tc_kernel.cuh), both split
// hi = tf32(x), lo = x - hi, and one thread issues the 12 tcgen05.mma of the 3xTF32 scheme (hi*hi + hi*lo + lo*hi) into a
// 128 x N fp32 accumulator in TMEM; tcgen05.ld -> per-warp slab -> 128-byte row pieces of `out`.  fp32 accuracy: the dropped
// lo*lo term is 2^-22 relative.  The next slice's values are requested before the current slice's MMAs are issued.
#pragma once
#include "fx_tc_kernel.cuh"

namespace fxtc {

struct GemmArgs {
  const float* X;   // [rows x k]
  const float* W;   // [k x c]
  float* out;       // [rows x c]
  int rows, k, c;
};

template <int N>
__global__ void __launch_bounds__(256, 3) k_gemm_xw(GemmArgs a) {
  constexpr int TM_COLS = N < 32 ? 32 : N;
  constexpr uint32_t LBO = 128, SBO = (TC_KCH / 4) * 128;
  constexpr uint32_t SBO_B = SBO + 16, B_BYTES = (N / 8) * SBO_B;
  extern __shared__ __align__(1024) unsigned char gm_smem[];
  unsigned char* Ahi = gm_smem;
  unsigned char* Alo = Ahi + TC_BH * TC_KCH * 4;
  unsigned char* Bhi = Alo + TC_BH * TC_KCH * 4;
  unsigned char* Blo = Bhi + B_BYTES;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TC_BH, n0 = blockIdx.y * N;
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
  const int nchunk = (a.k + TC_KCH - 1) / TC_KCH;
  // A: thread t owns the 16-byte pieces t, t+256, t+512, t+768 of the 128 x 32 tile (piece = row r, k-quad kq): eight
  // neighbouring threads read one 128-byte row segment of X.  B: the 4 x 4 unit (W rows kq*4.., features fq*4..) of the window kernel.
  constexpr int UNITS = 2 * N;
  const int fq = tid % (N / 4), kq = tid / (N / 4);
  const bool unit_ok = tid < UNITS && n0 + fq * 4 < a.c;
  float4 ax[4], bx[4];
  auto prefetch = [&](int ch) {
    const int k0 = ch * TC_KCH;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 256, r = idx >> 3, kk = k0 + (idx & 7) * 4;
      ax[q] = (row0 + r < a.rows && kk < a.k) ? __ldg(reinterpret_cast<const float4*>(a.X + (size_t)(row0 + r) * a.k + kk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int wr = k0 + kq * 4 + j;
      bx[j] = (unit_ok && wr < a.k) ? __ldg(reinterpret_cast<const float4*>(a.W + (size_t)wr * a.c + n0 + fq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  prefetch(0);
  uint32_t phase = 0;
  for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 256, r = idx >> 3, kq8 = idx & 7;
      float4 h, l;
      h.x = to_tf32(ax[q].x); h.y = to_tf32(ax[q].y); h.z = to_tf32(ax[q].z); h.w = to_tf32(ax[q].w);
      l.x = to_tf32(ax[q].x - h.x); l.y = to_tf32(ax[q].y - h.y); l.z = to_tf32(ax[q].z - h.z); l.w = to_tf32(ax[q].w - h.w);
      const uint32_t off = (uint32_t)(r >> 3) * SBO + (uint32_t)kq8 * LBO + (uint32_t)(r & 7) * 16u;
      *reinterpret_cast<float4*>(Ahi + off) = h;
      *reinterpret_cast<float4*>(Alo + off) = l;
    }
    if (tid < UNITS) {
      const float xs[4][4] = {{bx[0].x, bx[1].x, bx[2].x, bx[3].x}, {bx[0].y, bx[1].y, bx[2].y, bx[3].y},
                              {bx[0].z, bx[1].z, bx[2].z, bx[3].z}, {bx[0].w, bx[1].w, bx[2].w, bx[3].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = fq * 4 + i;
        float4 h, l;
        h.x = to_tf32(xs[i][0]); h.y = to_tf32(xs[i][1]); h.z = to_tf32(xs[i][2]); h.w = to_tf32(xs[i][3]);
        l.x = to_tf32(xs[i][0] - h.x); l.y = to_tf32(xs[i][1] - h.y); l.z = to_tf32(xs[i][2] - h.z); l.w = to_tf32(xs[i][3] - h.w);
        const uint32_t off = (uint32_t)(n >> 3) * SBO_B + (uint32_t)kq * LBO + (uint32_t)(n & 7) * 16u;
        *reinterpret_cast<float4*>(Bhi + off) = h;
        *reinterpret_cast<float4*>(Blo + off) = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (ch + 1 < nchunk) prefetch(ch + 1);
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ahi = smem_addr(Ahi), alo = smem_addr(Alo), bhi = smem_addr(Bhi), blo = smem_addr(Blo);
#pragma unroll
      for (int ks = 0; ks < TC_KCH / 8; ++ks) {
        const uint64_t dah = make_desc(ahi + ks * 2 * LBO, LBO, SBO), dal = make_desc(alo + ks * 2 * LBO, LBO, SBO);
        const uint64_t dbh = make_desc(bhi + ks * 2 * LBO, LBO, SBO_B), dbl = make_desc(blo + ks * 2 * LBO, LBO, SBO_B);
        mma_tf32(tmem, dah, dbh, idesc, (ch | ks) ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dal, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    }
    mbar_wait_parity(&bar, phase);  // the slice's MMAs are done: the tiles may be overwritten
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  {  // epilogue: as k_spmm_tc (warp w reads TMEM lanes 32(w%4).., warps 0-3 the lower half of the features, 4-7 the upper half)
    constexpr int HALF = N >= 64 ? N / 2 : N;
    const int cbeg = (N >= 64 && warp >= 4) ? HALF : 0;
    const bool active = N >= 64 || warp < 4;
    constexpr int SLAB_STRIDE = 36;
    float* slab = reinterpret_cast<float*>(gm_smem) + warp * (32 * SLAB_STRIDE);
    __syncthreads();
    if (active && nchunk > 0) {
      const int r0 = (warp & 3) * 32;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + HALF; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)r0 << 16) + (uint32_t)c0;
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
          *reinterpret_cast<uint4*>(slab + lane * SLAB_STRIDE + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int col = c0 + (lane & 7) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          const uint4 x = *reinterpret_cast<const uint4*>(slab + rr * SLAB_STRIDE + (lane & 7) * 4);
          if (n0 + col < a.c && row0 + r0 + rr < a.rows)
            *reinterpret_cast<uint4*>(a.out + (size_t)(row0 + r0 + rr) * a.c + n0 + col) = x;
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

template <int N>
inline size_t gemm_smem_bytes() {
  return (size_t)2 * TC_BH * TC_KCH * 4 + (size_t)2 * (N / 8) * ((TC_KCH / 4) * 128 + 16);
}

}  // namespace fxtc
