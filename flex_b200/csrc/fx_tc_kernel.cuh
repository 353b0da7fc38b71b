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
//   B chunk [N x 32 k] (B rows transposed while staging): the same formula with r = feature n (SBO = 1040
//        in k_spmm_tc: see there).
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
  const uint16_t* win_code;  // word of the nz inside the chunk's 128 x 32 A tile (tile_word(row, k), below)
  const float* win_val;
  const int* tc_panels;      // [ntc] panels that have a window
  const int* tc_cols;        // [npanel][W] column list of each panel, -1 padded
  const int* tc_ncol;        // [npanel]
  const float* B;            // [ncols x k]
  float* out;                // [ntc][128][k]
  int k, W;   // k = row stride of B and out (floats)
  int width;  // feature columns computed, from the B/out pointers on
  int hints;  // L2 eviction priorities on (FLEX_HINTS)
  int p_lo, p_hi;  // only the panels in [p_lo, p_hi) are multiplied (row groups of the host pipeline, fx_spmm_host)
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

// round to nearest (ties away) tf32 = cvt.rna.tf32.f32 for finite x, done on the bit pattern: the cvt instruction is
// emulated on sm_100a (VIADD 0x1000, an Inf/NaN test, SEL, LOP3 -- four instructions per value, and a chunk converts
// 16 B values + its nz per thread twice: 42 % of the kernel's instructions were this conversion)
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"  // suspend-time hint: the spin was 11 % of the kernel's instructions
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t}" ::"r"(smem_addr(bar)), "r"(parity), "r"(0x989680u) : "memory");
}

// 4-byte word of element (r, kk) in a K-major 128 x 32 tile with LBO = 128, SBO = 1024 bytes: the builder
// stores this number per window nz, so the scatter is one shift and two stores
__host__ __device__ __forceinline__ uint32_t tile_word(int r, int kk) {
  return ((uint32_t)(r >> 3) << 8) | ((uint32_t)(kk >> 2) << 5) | ((uint32_t)(r & 7) << 2) | (uint32_t)(kk & 3);
}

// grid = (ntc, ceil(k / N)), block = 256.  Dynamic shared memory: 2*128*32*4 + 2*N*32*4 + 4*W bytes, so three
// CTAs share an SM and the staging of one overlaps the MMAs of another.
// (tried: leaving the low part unrounded -- the tensor core reads its upper 19 bits -- for two instructions less per value:
// same time, so the rounding stays.  The kernel is bound by its per-chunk chain of barriers, fence and MMA wait, not by
// instruction issue: dropping 17 % of its instructions moved the step by 0.3 %.)
template <int N>
__global__ void __launch_bounds__(256, 3) k_spmm_tc(TcArgs a) {
  constexpr int TM_COLS = N < 32 ? 32 : N;
  constexpr uint32_t LBO = 128, SBO = (TC_KCH / 4) * 128;
  // B tiles: 16 extra bytes between 8-feature groups, so that the 16-byte stores of a warp whose lanes hold
  // four consecutive features each (two lanes per group) fall on all 32 banks
  constexpr uint32_t SBO_B = SBO + 16, B_BYTES = (N / 8) * SBO_B;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  float* Ahi = reinterpret_cast<float*>(tc_smem);
  float* Alo = Ahi + TC_BH * TC_KCH;
  unsigned char* Bhi = reinterpret_cast<unsigned char*>(Alo + TC_BH * TC_KCH);
  unsigned char* Blo = Bhi + B_BYTES;
  int* scols = reinterpret_cast<int*>(Blo + B_BYTES);  // [W] the panel's column list
  int* scptr = scols + a.W;                            // [W/32 + 1] nz range of every chunk
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.x, panel = a.tc_panels[slot], n0 = blockIdx.y * N;
  if (panel < a.p_lo || panel >= a.p_hi) return;  // uniform over the CTA, before any allocation
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
  for (int i = tid; i <= a.W / TC_KCH; i += 256) scptr[i] = __ldg(a.win_cptr + (size_t)panel * (a.W / TC_KCH) + i);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc<N>();

  uint32_t phase = 0;
  const int nchunk = (ncol + TC_KCH - 1) / TC_KCH;
  // L2 priorities as in the row kernel: B rows evict_last, the window's nz stream and the product evict_first
  uint64_t pol_keep, pol_stream;
  if (a.hints) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  } else {
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
    pol_stream = pol_keep;
  }
  auto ldB = [&](const float4* p) {
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_keep));
    return v;
  };

  // Software pipeline: the global loads of chunk ch+1 (B rows through the column list, the first nz
  // of every thread) are issued into registers before the MMAs of chunk ch, so their L2 latency runs
  // under the barrier, the tensor pipe and its wait; conversion and the shared-memory stores follow.
  // A thread owns one 4 x 4 unit of the chunk: B rows kq*4..+3, features fq*4..+3 (one 16-byte load per
  // row, a warp reads whole 512-byte rows), transposed in registers into four 16-byte words of 4 k.
  constexpr int UNITS = 2 * N;  // (32 / 4) * (N / 4)
  constexpr int NE = 4;         // nz per thread kept in registers; longer chunks loop
  const int fq = tid % (N / 4), kq = tid / (N / 4);
  const bool unit_ok = tid < UNITS && n0 + fq * 4 < a.width;
  const float4* Bq = reinterpret_cast<const float4*>(a.B + n0 + fq * 4);
  const size_t k4 = (size_t)a.k / 4;
  float4 bx[4];
  uint32_t ecode[NE];
  float eval[NE];
  int eb = 0, ee = 0;
  auto prefetch = [&](int ch) {
    eb = scptr[ch]; ee = scptr[ch + 1];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    bx[0] = z; bx[1] = z; bx[2] = z; bx[3] = z;
    if (tid < UNITS) {
      const int4 gc = *reinterpret_cast<const int4*>(scols + ch * TC_KCH + kq * 4);
      if (unit_ok) {
        if (gc.x >= 0) bx[0] = ldB(Bq + (size_t)gc.x * k4);
        if (gc.y >= 0) bx[1] = ldB(Bq + (size_t)gc.y * k4);
        if (gc.z >= 0) bx[2] = ldB(Bq + (size_t)gc.z * k4);
        if (gc.w >= 0) bx[3] = ldB(Bq + (size_t)gc.w * k4);
      }
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
        Ahi[ecode[q]] = h;
        Alo[ecode[q]] = l;
      }
    }
    const int ceb = eb, cee = ee;
    for (int e = ceb + tid + NE * 256; e < cee; e += 256) {
      const uint32_t code = a.win_code[e];
      const float v = a.win_val[e];
      const float h = to_tf32(v), l = to_tf32(v - h);
      Ahi[code] = h;
      Alo[code] = l;
    }
    // generic-proxy writes -> visible to the tensor core's async proxy.  The fence is a full membar: it waits
    // for every outstanding load of the thread, so the prefetch is issued AFTER it and stays in flight under
    // the barrier, the MMAs and their wait
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
    mbar_wait_parity(&bar, phase);  // MMAs of this chunk are done: the tiles may be overwritten
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // clear only the A words this chunk set
    if (ch + 1 < nchunk) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        if (ccode[q] != 0xFFFFFFFFu) { Ahi[ccode[q]] = 0.f; Alo[ccode[q]] = 0.f; }
      }
      for (int e = ceb + tid + NE * 256; e < cee; e += 256) {
        const uint32_t code = a.win_code[e];
        Ahi[code] = 0.f;
        Alo[code] = 0.f;
      }
    }
  }

  // epilogue: TMEM lane = panel row; warp w reads lanes 32(w%4).., warps 0-3 the lower half of the features and warps
  // 4-7 the upper half.  tcgen05.ld hands a thread 32 consecutive features of ITS row, so storing from there writes 16
  // bytes to 32 different rows per instruction (32 L1 wavefronts; the epilogue was ~30 % of the kernel's L1 data-pipe
  // time).  Each warp instead turns its 32 x 32 block through a padded slab of the (now idle) operand memory and stores
  // four whole 128-byte row pieces per instruction.
  {
    constexpr int HALF = N >= 64 ? N / 2 : N;
    const int cbeg = (N >= 64 && warp >= 4) ? HALF : 0;
    const bool active = N >= 64 || warp < 4;
    constexpr int SLAB_STRIDE = 36;  // floats per staged row: 16-byte aligned, rows 4 banks apart
    float* slab = reinterpret_cast<float*>(tc_smem) + warp * (32 * SLAB_STRIDE);
    __syncthreads();  // every thread is past the last MMA wait: the operand tiles are free
    if (active) {
      const int r0 = (warp & 3) * 32;
      float* dst0 = a.out + ((size_t)slot * TC_BH + r0) * a.k + n0;
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
          if (n0 + col < a.width)
            asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(dst0 + (size_t)rr * a.k + col),
                         "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w), "l"(pol_stream) : "memory");
        }
        __syncwarp();  // the slab is rewritten by the next 32 features
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

template <int N>
inline size_t tc_smem_bytes(int W) {
  return (size_t)2 * TC_BH * TC_KCH * 4 + (size_t)2 * (N / 8) * ((TC_KCH / 4) * 128 + 16) + (size_t)W * 4 + (size_t)(W / TC_KCH + 1) * 4;
}

}  // namespace fxtc
