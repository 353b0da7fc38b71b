// Micro-probe for tcgen05.mma operand plumbing: all shared memory = 1.0, one MMA, print D[lane 0..][col 0..3].
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/tc_probe.cu -o scripts/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe {
  uint32_t idesc;
  uint32_t desc_hi_a, desc_hi_b;  // upper 32 bits of the smem descriptors
  uint32_t lbo_a, lbo_b;          // >>4 units, go to bits 16..29
  int kind;                       // 0 tf32, 1 f16 (bf16 data)
  int form;                       // 0 mask form, 1 plain form
  int elect;                      // 1: issue from elect.sync lane of warp 0
  int fill;                       // 0 generic stores, 1 bulk async copy from global
  const float* ones;              // global buffer of ones (fp32 or bf16 pattern)
  float* out;                     // [128][4]
};

__global__ void __launch_bounds__(128, 1) k_probe(Probe p) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int BYTES = 128 * 1024;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tbase)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  {  // sentinel
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t s7 = __float_as_uint(7.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(s7) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (p.fill == 0) {
    const uint32_t w = p.kind == 0 ? __float_as_uint(1.0f) : 0x3f803f80u;
    for (int i = tid; i < BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = w;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  } else if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar2)), "r"(BYTES) : "memory");
    for (int off = 0; off < BYTES; off += 32768)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(sm + off)),
                   "l"(reinterpret_cast<const unsigned char*>(p.ones) + off), "r"(32768), "r"(saddr(&bar2))
                   : "memory");
  }
  if (p.fill == 1) {
    asm volatile("{\n\t.reg .pred q;\n\tPW:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n\t@q bra PD;\n\tbra PW;\n\tPD:\n\t}" ::"r"(saddr(&bar2)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint64_t da = (uint64_t)((saddr(sm) >> 4) & 0x3fff) | ((uint64_t)p.lbo_a << 16) | ((uint64_t)p.desc_hi_a << 32);
  const uint64_t db = (uint64_t)((saddr(sm + 65536) >> 4) & 0x3fff) | ((uint64_t)p.lbo_b << 16) | ((uint64_t)p.desc_hi_b << 32);
  bool issuer = tid == 0;
  if (p.elect) {
    uint32_t pred = 0;
    if (warp == 0)
      asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    issuer = pred != 0;
  }
  if (issuer) {
    if (p.kind == 0) {
      if (p.form == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(0u), "r"(0u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(0u) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(&bar)) : "memory");
  }
  __syncwarp();
  asm volatile("{\n\t.reg .pred q;\n\tQW:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n\t@q bra QD;\n\tbra QW;\n\tQD:\n\t}" ::"r"(saddr(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[4];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 4; ++j) p.out[tid * 4 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

static int run(const char* name, Probe p, float* dout, const float* dones_f32, const float* dones_bf16) {
  p.out = dout;
  p.ones = p.kind == 0 ? dones_f32 : dones_bf16;
  CK(cudaMemset(dout, 0xff, 128 * 4 * 4));
  k_probe<<<1, 128, 128 * 1024>>>(p);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s : CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  float h[128 * 4];
  CK(cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost));
  printf("%-44s : lane0 %g %g  lane1 %g lane37 %g lane127 %g %g\n", name, h[0], h[1], h[4], h[37 * 4], h[127 * 4], h[127 * 4 + 3]);
  return 0;
}

int main() {
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  float *dout, *df, *db;
  CK(cudaMalloc(&dout, 128 * 4 * 4)); CK(cudaMalloc(&df, 128 * 1024)); CK(cudaMalloc(&db, 128 * 1024));
  {
    static uint32_t h[32768];
    for (auto& x : h) x = 0x3f800000u;
    CK(cudaMemcpy(df, h, sizeof h, cudaMemcpyHostToDevice));
    for (auto& x : h) x = 0x3f803f80u;
    CK(cudaMemcpy(db, h, sizeof h, cudaMemcpyHostToDevice));
  }
  const uint32_t V = 1u << 14;  // version bit 46 in the upper word
  const uint32_t tf_kk = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);            // A,B K-major
  const uint32_t tf_kmn = tf_kk | (1u << 16);                                                      // B MN-major
  const uint32_t bf_kk = (1u << 4) | (1u << 7) | (1u << 10) | (16u << 17) | (8u << 24);            // bf16 K-major
  Probe base{};
  base.idesc = tf_kmn; base.desc_hi_a = V | 128; base.desc_hi_b = V | 8; base.lbo_a = 8; base.lbo_b = 256;
  Probe p;
  p = base; run("tf32 K/MN mask tid0 (as the kernel)", p, dout, df, db);
  p = base; p.form = 1; run("tf32 K/MN plain form", p, dout, df, db);
  p = base; p.elect = 1; run("tf32 K/MN elect.sync", p, dout, df, db);
  p = base; p.fill = 1; run("tf32 K/MN bulk-async fill", p, dout, df, db);
  p = base; p.idesc = tf_kk; p.desc_hi_b = V | 128; p.lbo_b = 8; run("tf32 K/K", p, dout, df, db);
  p = base; p.desc_hi_a = 128; p.desc_hi_b = 8; run("tf32 K/MN version bit off", p, dout, df, db);
  p = base; p.idesc = tf_kk; p.desc_hi_a = V | 64 | (2u << 29); p.desc_hi_b = V | 64 | (2u << 29); p.lbo_a = 1; p.lbo_b = 1;
  run("tf32 K/K swizzle128 (SBO 1024)", p, dout, df, db);
  p = base; p.kind = 1; p.idesc = bf_kk; p.desc_hi_a = V | 128; p.desc_hi_b = V | 128; p.lbo_a = 8; p.lbo_b = 8;
  run("bf16 K/K no swizzle", p, dout, df, db);
  p = base; p.kind = 1; p.idesc = bf_kk; p.desc_hi_a = V | 64 | (2u << 29); p.desc_hi_b = V | 64 | (2u << 29); p.lbo_a = 1; p.lbo_b = 1;
  run("bf16 K/K swizzle128", p, dout, df, db);
  p = base; p.kind = 1; p.idesc = bf_kk; p.desc_hi_a = V | 64 | (2u << 29); p.desc_hi_b = V | 64 | (2u << 29); p.lbo_a = 1; p.lbo_b = 1; p.fill = 1;
  run("bf16 K/K swizzle128 bulk fill", p, dout, df, db);
  for (uint32_t lt : {0u, 1u, 2u, 4u, 6u}) {
    char nm[64];
    p = base; p.desc_hi_b = V | 64 | (lt << 29); p.lbo_b = 64;
    snprintf(nm, sizeof nm, "tf32 K/MN layout_type %u lbo=sbo=1024", lt); run(nm, p, dout, df, db);
    p = base; p.desc_hi_b = V | 8 | (lt << 29); p.lbo_b = 8;
    snprintf(nm, sizeof nm, "tf32 K/MN layout_type %u lbo=sbo=128", lt); run(nm, p, dout, df, db);
    p = base; p.desc_hi_b = V | 0 | (lt << 29); p.lbo_b = 0;
    snprintf(nm, sizeof nm, "tf32 K/MN layout_type %u lbo=sbo=0", lt); run(nm, p, dout, df, db);
  }
  {
    const uint32_t bf_kmn = bf_kk | (1u << 16);
    p = base; p.kind = 1; p.idesc = bf_kmn; p.desc_hi_a = V | 128; p.desc_hi_b = V | 8; p.lbo_a = 8; p.lbo_b = 256;
    run("bf16 K/MN no swizzle", p, dout, df, db);
    const uint32_t tf_mnk = tf_kk | (1u << 15);
    p = base; p.idesc = tf_mnk; p.desc_hi_a = V | 8; p.lbo_a = 256; p.desc_hi_b = V | 128; p.lbo_b = 8;
    run("tf32 MN/K no swizzle", p, dout, df, db);
  }
  return 0;
}
