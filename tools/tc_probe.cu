// tcgen05 feasibility probe for the tensor-core Conv3d path (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu -lcuda && ./tc_probe
// Checks, on a real B200:
//   1. K-major SWIZZLE_128B descriptors + kind::tf32 MMA (M=128, N=32, K=32) against a CPU reference;
//   2. starting the A descriptor at an UNALIGNED row (row offset 1, 2 of a 136-row tile) -- the kw tap shift of an
//      implicit-GEMM conv -- for each candidate value of the descriptor's base_offset field;
//   3. 3xTF32 (hi*hi + hi*lo + lo*hi) accuracy against an fp64 reference;
//   4. MMA issue rate for N = 32 / 64 / 128 / 256 with A and B in shared memory (is small N A-read bound?).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {   // K-major SWIZZLE_64B, 512-byte 8-row atoms
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address
  d |= (uint64_t)1 << 16;                                  // LBO (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;        // SBO: stride between 8-row groups
  d |= (uint64_t)1 << 46;                                  // version = 1 (Blackwell)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int kRowsA = 136;  // 128 + 8 halo rows
struct Params {
  float* out;        // [128][64]
  long long* cycles; // [8]
  int mode;          // 0: gemm 1xTF32, 1: 3xTF32 stacked, 2: timing
  int row_off;       // start row of A
  int base_off;      // descriptor base_offset to use
  int N;             // MMA N
  int reps;
};

// smem: Ahi[136][32] | Alo[136][32] | B[256][32] (Bhi rows 0..N-1, Blo rows N..2N-1) ; all 1024-aligned
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* Ahi = (float*)base;
  float* Alo = (float*)(base + 17 * 1024);                  // 136*128 = 17408 = 17 KB
  float* Bs = (float*)(base + 34 * 1024);                   // 256*128 = 32 KB
  uint64_t* bars = (uint64_t*)(base + 66 * 1024);
  uint32_t* tmem_slot = (uint32_t*)(base + 66 * 1024 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], kRowsA * 128 + 256 * 128);
    tma_load_2d(Ahi, &mapA, &bars[0], 0, 0);
    tma_load_2d(Bs, &mapB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  // lo split, position preserving: lo = a - trunc_tf32(a)
  for (int i = tid; i < kRowsA * 32; i += 128) {
    const float a = Ahi[i];
    const float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    Alo[i] = a - hi;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int N = p.N;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a_hi = smem_u32(Ahi) + p.row_off * 128, a_lo = smem_u32(Alo) + p.row_off * 128;
    const uint32_t b_hi = smem_u32(Bs), b_lo = smem_u32(Bs) + N * 128;
    if (p.mode == 0) {
      const uint32_t idesc = make_idesc(128, N);
      for (int k = 0; k < 4; ++k)
        mma_tf32(tmem, make_desc(a_hi + 32 * k, 1024, p.base_off), make_desc(b_hi + 32 * k, 1024, 0), idesc, k > 0);
    } else if (p.mode == 1) {
      // D[:, 0:N]  = Alo*Bhi + Ahi*Bhi ; D[:, N:2N] = Ahi*Blo  (second MMA uses the stacked [Bhi|Blo] operand, N' = 2N)
      const uint32_t idN = make_idesc(128, N), id2N = make_idesc(128, 2 * N);
      for (int k = 0; k < 4; ++k) mma_tf32(tmem, make_desc(a_lo + 32 * k, 1024, p.base_off), make_desc(b_hi + 32 * k, 1024, 0), idN, k > 0);
      // columns N..2N-1 are fresh in the first stacked MMA only when accumulate=0 for them; clear them with a lo*lo-free trick:
      // run the stacked MMA with accumulate=1 on columns 0..N-1 requires columns N..2N-1 zeroed first -> issue Ahi*Blo separately
      for (int k = 0; k < 4; ++k)
        mma_tf32(tmem + N, make_desc(a_hi + 32 * k, 1024, p.base_off), make_desc(b_lo + 32 * k, 1024, 0), idN, k > 0);
      for (int k = 0; k < 4; ++k) mma_tf32(tmem, make_desc(a_hi + 32 * k, 1024, p.base_off), make_desc(b_hi + 32 * k, 1024, 0), idN, 1);
      (void)id2N;
    } else if (p.mode == 2) {
      const uint32_t idesc = make_idesc(128, N);
      t0 = clock64();
      for (int r = 0; r < p.reps; ++r)
        for (int k = 0; k < 4; ++k)
          mma_tf32(tmem, make_desc(a_hi + 32 * k, 1024, 0), make_desc(b_hi + 32 * k, 1024, 0), idesc, (r | k) > 0);
    } else if (p.mode == 3) {   // SWIZZLE_64B operands (64-byte rows), timing only
      const uint32_t idesc = make_idesc(128, N);
      t0 = clock64();
      for (int r = 0; r < p.reps; ++r)
        for (int k = 0; k < 4; ++k)
          mma_tf32(tmem, make_desc64(a_hi + 32 * (k & 1)), make_desc64(b_hi + 32 * (k & 1)), idesc, (r | k) > 0);
    } else {                    // mode 4: SWIZZLE_128B, MMAs round-robin over 4 accumulator tiles
      const uint32_t idesc = make_idesc(128, N);
      t0 = clock64();
      for (int r = 0; r < p.reps; ++r)
        for (int k = 0; k < 4; ++k)
          mma_tf32(tmem + (k & 3) * 128, make_desc(a_hi + 32 * k, 1024, 0), make_desc(b_hi + 32 * k, 1024, 0), idesc, r > 0);
    }
    mma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  if (tid == 0 && p.mode >= 2) {
    t1 = clock64();
    p.cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) p.out[(size_t)tid * 64 + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* ptr, int rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {32, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t el[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

static float tf32_trunc(float a) {
  uint32_t u;
  memcpy(&u, &a, 4);
  u &= 0xffffe000u;
  memcpy(&a, &u, 4);
  return a;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)fnp;
  const int NB = 256;
  float *hA = (float*)malloc(kRowsA * 32 * 4), *hB = (float*)malloc(NB * 32 * 4), *hO = (float*)malloc(128 * 64 * 4);
  srand(1);
  for (int i = 0; i < kRowsA * 32; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  // B rows 0..31: full fp32 weights W; for 3xTF32: rows 0..31 = hi(W), rows 32..63 = lo(W) (set per test)
  float W[32 * 32];
  for (int i = 0; i < 32 * 32; ++i) W[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dO;
  long long* dC;
  CK(cudaMalloc(&dA, kRowsA * 32 * 4));
  CK(cudaMalloc(&dB, NB * 32 * 4));
  CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMalloc(&dC, 64));
  CK(cudaMemcpy(dA, hA, kRowsA * 32 * 4, cudaMemcpyHostToDevice));
  CUtensorMap mapA = make_map(fn, dA, kRowsA), mapB = make_map(fn, dB, NB);
  const size_t smem = 68 * 1024 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  auto run = [&](Params p) {
    p.out = dO;
    p.cycles = dC;
    CK(cudaMemset(dO, 0, 128 * 64 * 4));
    probe_kernel<<<1, 128, smem>>>(mapA, mapB, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  kernel failed: %s\n", cudaGetErrorString(e));
      exit(2);
    }
    CK(cudaMemcpy(hO, dO, 128 * 64 * 4, cudaMemcpyDeviceToHost));
  };

  // ---- test 5 (./tc_probe acc): how does the TMEM accumulator round?  The same K = 32 product block is accumulated `reps` times
  // into one accumulator (4 MMAs of K = 8 each per rep).  With all-positive operands the exact answer is reps * (A.B); round-to-
  // nearest accumulation leaves a zero-mean error growing like sqrt(#MMAs) * 2^-25, truncation (RZ) a NEGATIVE bias growing like
  // #MMAs * 2^-25.  Second pass: random-sign operands (error relative to the RMS magnitude of the result).
  if (argc > 1 && !strcmp(argv[1], "acc")) {
    for (int pass = 0; pass < 2; ++pass) {
      float* hA2 = (float*)malloc(kRowsA * 32 * 4);
      for (int i = 0; i < kRowsA * 32; ++i) hA2[i] = pass == 0 ? fabsf(hA[i]) + 0.25f : hA[i];
      for (int i = 0; i < NB * 32; ++i) hB[i] = 0.f;
      for (int i = 0; i < 32 * 32; ++i) hB[i] = pass == 0 ? fabsf(W[i]) + 0.25f : W[i];
      CK(cudaMemcpy(dA, hA2, kRowsA * 32 * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, hB, NB * 32 * 4, cudaMemcpyHostToDevice));
      for (int reps : {1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048}) {
        Params p{};
        p.mode = 2, p.N = 32, p.reps = reps;
        run(p);
        double sum_rel = 0, max_rel = 0, rms = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)tf32_trunc(hA2[m * 32 + k]) * (double)tf32_trunc(hB[n * 32 + k]);
            rms += ref * ref;
          }
        rms = sqrt(rms / (128 * 32));
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)tf32_trunc(hA2[m * 32 + k]) * (double)tf32_trunc(hB[n * 32 + k]);
            const double denom = pass == 0 ? ref * reps : rms * reps;
            const double rel = ((double)hO[m * 64 + n] - ref * reps) / denom;
            sum_rel += (pass == 0) ? rel : rel * (ref > 0 ? 1.0 : -1.0);      // signed towards/away from zero
            max_rel = fmax(max_rel, fabs(rel));
          }
        printf("acc %s reps=%4d (%5d MMAs): mean signed rel err %+.3e (x 2^-24 = %+.2f), max |rel| %.3e\n",
               pass == 0 ? "positive" : "random  ", reps, 4 * reps, sum_rel / (128 * 32), sum_rel / (128 * 32) / 5.96e-8, max_rel);
      }
      free(hA2);
    }
    return 0;
  }
  // ---- test 1 + 2: 1xTF32 GEMM with row offsets / base_offset candidates
  for (int i = 0; i < NB * 32; ++i) hB[i] = 0.f;
  for (int i = 0; i < 32 * 32; ++i) hB[i] = W[i];
  CK(cudaMemcpy(dB, hB, NB * 32 * 4, cudaMemcpyHostToDevice));
  for (int off = 0; off <= 3; ++off) {
    for (int bo = 0; bo < 8; ++bo) {
      if (off == 0 && bo > 0) break;
      Params p{};
      p.mode = 0, p.row_off = off, p.base_off = bo, p.N = 32, p.reps = 1;
      run(p);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) ref += (double)tf32_trunc(hA[(m + off) * 32 + k]) * (double)tf32_trunc(W[n * 32 + k]);
          maxerr = fmax(maxerr, fabs(ref - hO[m * 64 + n]));
        }
      printf("gemm1x row_off=%d base_offset=%d max_abs_err=%.3e %s\n", off, bo, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  }
  // ---- test 3: 3xTF32 accuracy
  {
    for (int i = 0; i < 32 * 32; ++i) {
      float hi = tf32_trunc(W[i]);
      hB[i] = hi;
      hB[32 * 32 + i] = W[i] - hi;
    }
    CK(cudaMemcpy(dB, hB, NB * 32 * 4, cudaMemcpyHostToDevice));
    Params p{};
    p.mode = 1, p.row_off = 0, p.base_off = 0, p.N = 32, p.reps = 1;
    run(p);
    double maxerr3 = 0, maxerr1 = 0, maxfp32 = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0, r1 = 0;
        float f32 = 0.f;
        for (int k = 0; k < 32; ++k) {
          ref += (double)hA[m * 32 + k] * (double)W[n * 32 + k];
          r1 += (double)tf32_trunc(hA[m * 32 + k]) * (double)tf32_trunc(W[n * 32 + k]);
          f32 = fmaf(hA[m * 32 + k], W[n * 32 + k], f32);
        }
        const float got = hO[m * 64 + n] + hO[m * 64 + 32 + n];
        maxerr3 = fmax(maxerr3, fabs(ref - got));
        maxerr1 = fmax(maxerr1, fabs(ref - r1));
        maxfp32 = fmax(maxfp32, fabs(ref - (double)f32));
      }
    printf("3xTF32 max_abs_err=%.3e   (1xTF32 would be %.3e, fp32 FMA chain %.3e)\n", maxerr3, maxerr1, maxfp32);
  }
  // ---- test 4: MMA rate vs N
  for (int N : {32, 64, 128, 256}) {
    Params p{};
    p.mode = 2, p.N = N, p.reps = 2000;
    run(p);
    long long cyc;
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    printf("rate N=%3d: %.1f cycles per M128xN%dxK8 MMA  (ideal N/2 = %d) -> %.0f MAC/clk\n", N, (double)cyc / (2000 * 4), N, N / 2,
           128.0 * N * 8 / ((double)cyc / (2000 * 4)));
  }
  for (int mode : {3, 4})
    for (int N : {32, 64, 96, 128}) {
      Params p{};
      p.mode = mode, p.N = N, p.reps = 2000;
      run(p);
      long long cyc;
      CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
      printf("rate mode=%d (%s) N=%3d: %.1f cycles per MMA\n", mode, mode == 3 ? "SW64 operands" : "SW128, 4 accumulators round-robin", N,
             (double)cyc / (2000 * 4));
    }
  {
    Params p{};
    p.mode = 2, p.N = 96, p.reps = 2000;
    run(p);
    long long cyc;
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    printf("rate mode=2 (SW128, one accumulator) N= 96: %.1f cycles per MMA\n", (double)cyc / (2000 * 4));
  }
  return 0;
}
