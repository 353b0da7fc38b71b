// fx_tc_kernel.cuh -- tensor-core path for the heavy columns of a row panel (sm_100a).
//
// For a 128-row panel p the builder (fx_tcw_build.cu) picks up to W columns that several rows of
// the panel share.  With S_p that column list, D_p = A_p[:, S_p] * B[S_p, :] is a dense
// 128 x |S_p| x N contraction in which every B row is fetched ONCE per panel instead of once per
// nz (the panel kernel of fx_spmm.cu is bound by exactly that L2->SM gather traffic).  It runs on
// the 5th-generation tensor cores: tcgen05.mma (kind::tf32, cta_group::1, M=128, N, K=8 per
// instruction) issued by one thread, operands in shared memory described by UMMA descriptors, the
// 128 x N fp32 accumulator in TMEM, read back with tcgen05.ld.  fp32 accuracy is kept with the
// 3xTF32 split: x = hi + lo with hi = tf32(x), lo = tf32(x - hi);
// A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi  (the dropped lo*lo term is 2^-22 relative).
//
// Shared-memory operand layouts (no swizzle, "interleaved" canonical UMMA layout, cf. CUTLASS
// cute/atom/mma_traits_sm100.hpp:165-199), both operands K-major:
//   A chunk [128 rows x 32 k] : core matrix = 8 rows x 16 B;  byte(r,kk) =
//        (r/8)*SBO + (kk/4)*128 + (r%8)*16 + (kk%4)*4,  LBO = 128, SBO = 1024
//   B chunk [N x 32 k] (B rows transposed while staging): the same formula with r = feature n.
//        MN-major tf32 operands only work with the 128B_BASE32B swizzle on this part (probed with
//        scripts/tc_probe.cu: every other layout type reads zeros), so B is transposed instead.
// One K=8 step spans two K-adjacent core matrices of either operand (start += 256 B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxtc {

constexpr int TC_BH = 128;  // panel height
constexpr int TC_KCH = 32;  // window columns staged per step

struct TcArgs {
  const int* win_cptr;       // [npanel*(W/32)+1] window nz of each (panel, 32-column chunk)
  const uint16_t* win_code;  // (row in panel << 5) | (position in the chunk)
  const float* win_val;
  const int* tc_panels;      // [ntc] panels that have a window
  const int* tc_cols;        // [npanel][W] column list of each panel, -1 padded
  const int* tc_ncol;        // [npanel]
  const float* B;            // [ncols x k]
  float* out;                // [ntc][128][k]
  int k, W;
};

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

__device__ __forceinline__ uint32_t a_offset(int r, int kk, uint32_t SBO) {
  return (uint32_t)(r >> 3) * SBO + (uint32_t)(kk >> 2) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kk & 3) * 4u;
}

// grid = (ntc, ceil(k / N)), block = 256.  Dynamic shared memory: 2*128*32*4 + 2*N*32*4 + 4*W bytes, so three
// CTAs share an SM and the staging of one overlaps the MMAs of another.
template <int N>
__global__ void __launch_bounds__(256, 3) k_spmm_tc(TcArgs a) {
  constexpr int TM_COLS = N < 32 ? 32 : N;
  constexpr uint32_t LBO = 128, SBO = (TC_KCH / 4) * 128;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  float* Ahi = reinterpret_cast<float*>(tc_smem);
  float* Alo = Ahi + TC_BH * TC_KCH;
  float* Bhi = Alo + TC_BH * TC_KCH;
  float* Blo = Bhi + TC_KCH * N;
  int* scols = reinterpret_cast<int*>(Blo + TC_KCH * N);  // [W] the panel's column list
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.x, panel = a.tc_panels[slot], n0 = blockIdx.y * N;
  const int ncol = a.tc_ncol[panel];
  const int* cols = a.tc_cols + (size_t)panel * a.W;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * TC_BH * TC_KCH / 4; i += 256) reinterpret_cast<float4*>(Ahi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < a.W; i += 256) scols[i] = i < ncol ? __ldg(cols + i) : -1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc<N>();

  uint32_t phase = 0;
  const int nchunk = (ncol + TC_KCH - 1) / TC_KCH;
  const int* cptr = a.win_cptr + (size_t)panel * (a.W / TC_KCH);

  // Software pipeline: the global loads of chunk ch+1 (B rows through the column list, the first
  // nz of every thread) are issued into registers right after the MMAs of chunk ch, so their L2
  // latency runs under the tensor pipe; conversion and the shared-memory stores follow the wait.
  constexpr int IT = (TC_KCH / 4) * N / 256;  // B items (4 k of one feature) per thread: idx = tid + it*256
  constexpr int NE = 4;                        // nz per thread kept in registers; longer chunks loop
  static_assert(IT >= 1, "N too small for 256 threads");
  float bx[IT][4];
  uint32_t ecode[NE];
  float eval[NE];
  int eb = 0, ee = 0;
  auto prefetch = [&](int ch) {
    const int s0 = ch * TC_KCH;
    eb = __ldg(cptr + ch); ee = __ldg(cptr + ch + 1);
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int idx = tid + it * 256, n = idx % N, kq = idx / N;
      const int4 gc = *reinterpret_cast<const int4*>(scols + s0 + kq * 4);
      const bool nok = n0 + n < a.k;
      bx[it][0] = (gc.x >= 0 && nok) ? __ldg(a.B + (size_t)gc.x * a.k + n0 + n) : 0.f;
      bx[it][1] = (gc.y >= 0 && nok) ? __ldg(a.B + (size_t)gc.y * a.k + n0 + n) : 0.f;
      bx[it][2] = (gc.z >= 0 && nok) ? __ldg(a.B + (size_t)gc.z * a.k + n0 + n) : 0.f;
      bx[it][3] = (gc.w >= 0 && nok) ? __ldg(a.B + (size_t)gc.w * a.k + n0 + n) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = eb + tid + q * 256;
      ecode[q] = e < ee ? (uint32_t)a.win_code[e] : 0xFFFFFFFFu;
      eval[q] = e < ee ? a.win_val[e] : 0.f;
    }
  };
  if (nchunk > 0) prefetch(0);

  for (int ch = 0; ch < nchunk; ++ch) {
    // B rows of the chunk, split and transposed to K-major (a warp read 128-byte runs of four B
    // rows; each lane keeps one feature and stores its 4 k as one 16-byte word)
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int idx = tid + it * 256, n = idx % N, kq = idx / N;
      float4 h, l;
      h.x = to_tf32(bx[it][0]); h.y = to_tf32(bx[it][1]); h.z = to_tf32(bx[it][2]); h.w = to_tf32(bx[it][3]);
      l.x = to_tf32(bx[it][0] - h.x); l.y = to_tf32(bx[it][1] - h.y); l.z = to_tf32(bx[it][2] - h.z); l.w = to_tf32(bx[it][3] - h.w);
      const uint32_t off = (uint32_t)(n >> 3) * SBO + (uint32_t)kq * LBO + (uint32_t)(n & 7) * 16u;
      *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Bhi) + off) = h;
      *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Blo) + off) = l;
    }
    // every thread has cleared its words of the previous chunk before anyone writes the A tiles again
    // (two chunks reuse the same (row, k) words)
    if (ch > 0) __syncthreads();
    // scatter the chunk's nz into the (zeroed) A tiles, all threads at once
    uint32_t ccode[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      ccode[q] = ecode[q];
      if (ecode[q] != 0xFFFFFFFFu) {
        const float h = to_tf32(eval[q]), l = to_tf32(eval[q] - h);
        const uint32_t off = a_offset((int)(ecode[q] >> 5), (int)(ecode[q] & 31u), SBO);
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ahi) + off) = h;
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Alo) + off) = l;
      }
    }
    const int ceb = eb, cee = ee;
    for (int e = ceb + tid + NE * 256; e < cee; e += 256) {
      const uint32_t code = a.win_code[e];
      const float v = a.win_val[e];
      const float h = to_tf32(v), l = to_tf32(v - h);
      const uint32_t off = a_offset((int)(code >> 5), (int)(code & 31u), SBO);
      *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ahi) + off) = h;
      *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Alo) + off) = l;
    }
    if (ch + 1 < nchunk) prefetch(ch + 1);  // in flight under the barrier, the MMAs and their wait
    // generic-proxy writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ahi = smem_addr(Ahi), alo = smem_addr(Alo), bhi = smem_addr(Bhi), blo = smem_addr(Blo);
#pragma unroll
      for (int ks = 0; ks < TC_KCH / 8; ++ks) {
        const uint64_t dah = make_desc(ahi + ks * 2 * LBO, LBO, SBO), dal = make_desc(alo + ks * 2 * LBO, LBO, SBO);
        const uint64_t dbh = make_desc(bhi + ks * 2 * LBO, LBO, SBO), dbl = make_desc(blo + ks * 2 * LBO, LBO, SBO);
        mma_tf32(tmem, dah, dbh, idesc, (ch | ks) ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dal, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    }
    mbar_wait_parity(&bar, phase);  // MMAs of this chunk are done: the tiles may be overwritten
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // clear only the A words this chunk set
    if (ch + 1 < nchunk) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        if (ccode[q] != 0xFFFFFFFFu) {
          const uint32_t off = a_offset((int)(ccode[q] >> 5), (int)(ccode[q] & 31u), SBO);
          *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ahi) + off) = 0.f;
          *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Alo) + off) = 0.f;
        }
      }
      for (int e = ceb + tid + NE * 256; e < cee; e += 256) {
        const uint32_t code = a.win_code[e];
        const uint32_t off = a_offset((int)(code >> 5), (int)(code & 31u), SBO);
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ahi) + off) = 0.f;
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Alo) + off) = 0.f;
      }
    }
  }

  // epilogue: TMEM lane = panel row; warp w reads lanes 32(w%4).., warps 0-3 the lower half of the
  // features and warps 4-7 the upper half
  {
    const int r = (warp & 3) * 32 + lane;
    float* dst = a.out + ((size_t)slot * TC_BH + r) * a.k + n0;
    constexpr int HALF = N >= 64 ? N / 2 : N;
    const int cbeg = (N >= 64 && warp >= 4) ? HALF : 0;
    const bool active = N >= 64 || warp < 4;
    if (active) {
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + HALF; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
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
          if (n0 + c0 + j < a.k)
            *reinterpret_cast<float4*>(dst + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

template <int N>
inline size_t tc_smem_bytes(int W) { return (size_t)2 * TC_BH * TC_KCH * 4 + (size_t)2 * TC_KCH * N * 4 + (size_t)W * 4; }

}  // namespace fxtc
