// Probe for the next optimisation of conv3d_tc_kernel: feeding the A operand of
// tcgen05.mma from TENSOR MEMORY instead of shared memory.
//
// Why: ncu (profiles/r1_ncu_summary_final.md) shows the W=128 stem kernel bound by the L1/shared-memory data pipe -- the
// tensor core's operand reads take 49 % of it and the loader's STS of the hi/lo split another large share.  With A in TMEM
// (written by the loader warps with tcgen05.st, which does not touch shared memory) every MMA reads only the 3 KB weight
// tile from shared memory instead of 4 KB + 3 KB, and the 2 x 16 KB STS per input row disappear.
//
// What it checks (nothing of this could be verified without a GPU; build: nvcc -arch=sm_100a -o tmem_a_probe tmem_a_probe.cu):
//   1. layout: A (128 x 32 fp32) stored with tcgen05.st.32x32b (thread = row = TMEM lane, columns = K) and consumed by
//      `tcgen05.mma.cta_group::1.kind::tf32 [d], [a_tmem], b_desc, idesc, p` gives D = A . B^T (max error vs host printed);
//   2. rate: cycles per M128 x N96 x K8 MMA with A from shared memory vs A from TMEM (the former measured 55-56 in
//      tools/tc_probe.cu; the latter should drop towards the 48-clk tensor-pipe time if the shared-memory read was the limiter).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define CHECK(x)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);              \
      exit(1);                                                                                     \
    }                                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
// A operand in tensor memory: 128 lanes (rows of A) x K 32-bit columns starting at a_tmem
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int N = 96, K = 32;
constexpr uint32_t A_COL = 256;      // A tile lives in TMEM columns [256, 288)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// MODE 0: correctness (A from TMEM);  MODE 1: rate with A from shared memory;  MODE 2: rate with A from TMEM.
// The issue loop is warp-uniform with an elected lane, as in the product kernels (an `if (t == 0)` body makes ptxas wrap every
// descriptor in R2UR broadcast loops and the measurement becomes issue-bound: 117 clk/MMA in the first version of this probe).
template <int MODE>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d,
                                             long long* cycles) {
  constexpr int mode = MODE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;                    // 128 rows x 128 B, SWIZZLE_128B
  uint8_t* sb = smem + 128 * 128;        // 96 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 128 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operands into shared memory (K-major rows of 32 floats, 128B swizzle: 16-byte chunk c of row r at r*128 + ((c ^ (r&7))<<4))
  for (int i = t; i < 128 * 8; i += 128) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<float4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const float4*>(a)[i];
  }
  for (int i = t; i < N * 8; i += 128) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<float4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const float4*>(b)[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *slot;
  // A into tensor memory: thread t owns row t (= lane t; a warp may only touch its own 32-lane quadrant)
  {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + A_COL;
    for (int c0 = 0; c0 < K; c0 += 8) tmem_st8(lane_base + c0, a + t * K + c0);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t idesc = idesc_tf32(128, N);
  const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
  if (warp == 0) {
    long long t0 = 0;
    if (mode == 0) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) mma_ts(tmem, tmem + A_COL + ks * 8, db + 2 * ks, idesc, ks > 0);
        commit(bar);
      }
    } else {
      t0 = clock64();
      for (int it = 0; it < 2000; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < K / 8; ++ks) {
            if (mode == 1) mma_ss(tmem, da + 2 * ks, db + 2 * ks, idesc, 1);
            else mma_ts(tmem, tmem + A_COL + ks * 8, db + 2 * ks, idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (t == 0 && mode != 0) *cycles = clock64() - t0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (mode == 0) {
    const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(row + c0, v);
      for (int i = 0; i < 16; ++i) d[t * N + c0 + i] = v[i];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- exploratory: where does tcgen05.cp put shared-memory bytes?  smem holds float(i) at float index i; after
// `tcgen05.cp.cta_group::1.<shape> [tmem], desc` every thread reads its lane's first 8 columns back and a few lanes are
// printed, for several descriptor settings (swizzle mode, leading / stride byte offsets).  Picks the layout for staging A
// tiles with ONE bulk copy per tile instead of per-thread tcgen05.st (and instead of 3 smem reads per k-step by the MMA).
__global__ void __launch_bounds__(128) cp_probe(int shape, uint64_t desc_hi_bits, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~static_cast<uintptr_t>(1023));
  float* src = reinterpret_cast<float*>(smem);                       // 16 KB = 4096 floats
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  for (int i = t; i < 4096; i += 128) src[i] = (float)i;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *slot;
  {   // clear the destination columns so untouched cells are recognisable (-1)
    float m1[8] = {-1.f, -1.f, -1.f, -1.f, -1.f, -1.f, -1.f, -1.f};
    tmem_st8(tmem + ((uint32_t)(warp * 32) << 16), m1);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (t == 0) {
    const uint64_t desc = (uint64_t)((smem_u32(src) & 0x3FFFF) >> 4) | desc_hi_bits;
    if (shape == 0) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem), "l"(desc) : "memory");
    else asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tmem), "l"(desc) : "memory");
    commit(bar);
  }
  if (warp == 0) {
    __syncwarp();
    mbar_wait(bar, 0);
  }
  fence_before();
  __syncthreads();
  fence_after();
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 8; ++i) out[t * 8 + i] = v[i];
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float tf32_trunc(float a) {
  uint32_t u;
  memcpy(&u, &a, 4);
  u &= 0xffffe000u;
  memcpy(&a, &u, 4);
  return a;
}

int main() {
  float *ha = (float*)malloc(128 * K * 4), *hb = (float*)malloc(N * K * 4), *hd = (float*)malloc(128 * N * 4);
  srand(1);
  for (int i = 0; i < 128 * K; ++i) ha[i] = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
  for (int i = 0; i < N * K; ++i) hb[i] = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
  float *da, *db, *dd;
  long long* dc;
  CHECK(cudaMalloc(&da, 128 * K * 4));
  CHECK(cudaMalloc(&db, N * K * 4));
  CHECK(cudaMalloc(&dd, 128 * N * 4));
  CHECK(cudaMalloc(&dc, 8));
  CHECK(cudaMemcpy(da, ha, 128 * K * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(db, hb, N * K * 4, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 2 * 128 * 128 + 64;
  CHECK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CHECK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CHECK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<0><<<1, 128, smem>>>(da, db, dd, dc);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(hd, dd, 128 * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)ha[m * K + k] * hb[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - hd[m * N + n]));
    }
  printf("A-from-TMEM  D = A.B^T  max_abs_err = %.3e  %s   (d[0][0..3] = %g %g %g %g)\n", maxerr, maxerr < 1e-4 ? "OK" : "WRONG LAYOUT",
         hd[0], hd[1], hd[2], hd[3]);
  for (int mode = 1; mode <= 2; ++mode) {
    long long cyc = 0;
    if (mode == 1) probe<1><<<1, 128, smem>>>(da, db, dd, dc);
    else probe<2><<<1, 128, smem>>>(da, db, dd, dc);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    printf("rate  A from %s:  %.1f cycles per M128 x N%d x K8 MMA\n", mode == 1 ? "shared memory" : "tensor memory ", (double)cyc / 8000.0, N);
  }
  // ---- tcgen05.cp layout exploration
  float* dout;
  CHECK(cudaMalloc(&dout, 128 * 8 * 4));
  float hout[128 * 8];
  CHECK(cudaFuncSetAttribute(cp_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 16384 + 64));
  struct Cfg { const char* name; int shape; uint64_t hi; };
  const uint64_t V1 = (uint64_t)1 << 46;
  auto lbo = [](uint64_t b) { return (uint64_t)(b >> 4) << 16; };
  auto sbo = [](uint64_t b) { return (uint64_t)(b >> 4) << 32; };
  const Cfg cfgs[] = {
      {"128x256b swizzle=none LBO=16   SBO=256 ", 0, V1 | lbo(16) | sbo(256)},
      {"128x256b swizzle=none LBO=128  SBO=256 ", 0, V1 | lbo(128) | sbo(256)},
      {"128x256b swizzle=none LBO=256  SBO=128 ", 0, V1 | lbo(256) | sbo(128)},
      {"128x256b swizzle=32B  LBO=16   SBO=256 ", 0, V1 | lbo(16) | sbo(256) | ((uint64_t)6 << 61)},
      {"128x256b swizzle=128B LBO=16   SBO=1024", 0, V1 | lbo(16) | sbo(1024) | ((uint64_t)2 << 61)},
      {"128x128b swizzle=none LBO=16   SBO=128 ", 1, V1 | lbo(16) | sbo(128)},
      {"128x128b swizzle=128B LBO=16   SBO=1024", 1, V1 | lbo(16) | sbo(1024) | ((uint64_t)2 << 61)},
  };
  for (const Cfg& c : cfgs) {
    cp_probe<<<1, 128, 1024 + 16384 + 64>>>(c.shape, c.hi, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("cp %s: kernel failed: %s\n", c.name, cudaGetErrorString(e));
      return 1;                                   // a fault poisons the context: stop here, rerun with the entry removed
    }
    CHECK(cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost));
    printf("cp %s: float index found at (lane, col 0..7)\n", c.name);
    const int lanes[] = {0, 1, 2, 7, 8, 9, 31, 32, 64, 127};
    for (int l : lanes) {
      printf("   lane %3d:", l);
      for (int i = 0; i < 8; ++i) printf(" %6.0f", hout[l * 8 + i]);
      printf("\n");
    }
  }
  return 0;
}
