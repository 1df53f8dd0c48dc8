// Tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM Conv3d 3x3x3, stride 1, pad 1, for the full-resolution layers of the
// hourglass aggregation (W = 128 output columns = one UMMA M tile).  fp32-accurate through 3xTF32 operand splitting:
//
//      a*b  ~=  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi ,   x_hi = x with the low 13 mantissa bits cleared, x_lo = x - x_hi
//
// (the kind::tf32 MMA reads only the top 19 bits of each fp32 operand, so a_hi needs no copy; a_lo is produced in shared
// memory by the converter warps; b_hi / b_lo are split once on the host).  Measured on B200 with tools/tc_probe.cu:
// max error 2.6e-6 against an fp64 reference where an fp32 FMA chain has 1.7e-6 and plain TF32 5.2e-3 -- i.e. this
// path keeps the north star's 1e-3 px EPE bar that plain TF32/BF16 tensor-core math breaks.
//
// Replaces the same reference modules as conv3d.cu (convbn_3d + ReLU of gwcnet/hourglass.py:5-16,
// gwcnet_disp_processor.py:40-81; conv3d_bn(_relu) psmnet/submodule.py:68-83,160-177) for layers with W == 128.
//
// GEMM mapping.  Activations are channels-last (B, D, H, W, Cin), so an A tile "128 voxels of one image row x 16 input
// channels" is K-major with 64-byte rows and lands in shared memory with one TMA box (SWIZZLE_64B).
//   * the three kw taps are NOT realised by shifting A (that would need halo columns and unaligned tiles); instead the
//     three weight slices are stacked along N -- one MMA of N = 3*Cout produces P_kw[m] = A[m] . B_kw for kw = 0,1,2 --
//     and the epilogue forms D[m] = P_0[m-1] + P_1[m] + P_2[m+1] with warp shuffles (zero padding at m = -1 / 128 is
//     implicit because a row tile spans the whole image width).  N = 96 also lifts the MMA above the ~54-cycle floor
//     that small-N tf32 MMAs hit (tools/tc_probe.cu: N=32 and N=64 both take 54.5 cycles).
//   * kh taps: output row t of the block needs input rows t-1, t, t+1: input rows stream through a shared-memory ring
//     and each staged row feeds up to three accumulator tiles.
//   * kd taps and Cin chunks are phases of one work item; the weights of a phase (3 kh x 3*Cout rows x 16 ch, hi+lo)
//     are double-buffered.
// Work item = (image b, output plane d, block of 5 output rows); 5 accumulator tiles x 96 columns = 480 TMEM columns.
//
// Warp roles (192 threads, 1 CTA/SM, persistent): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator),
// warps 2-5 = lo-split converters and epilogue (TMEM -> registers -> BN/residual/ReLU -> global).
#include "common.cuh"

namespace osb {

constexpr int TC_W = 128;          // image width handled (UMMA M)
constexpr int TC_KC = 16;          // input channels per phase (64-byte K-major rows, SWIZZLE_64B)
constexpr int TC_TILES = 5;        // output rows (accumulator tiles) per work item
constexpr int TC_ROWS = TC_TILES + 2;
constexpr int TC_STAGES = 4;       // A-row ring depth
constexpr int TC_ROW_BYTES = TC_W * TC_KC * 4;     // 8192: one staged input row (hi or lo)
constexpr int TC_THREADS = 192;

struct TcParams {
  const float* scale;
  const float* shift;
  const float* residual;
  float* y;
  int B, D, H, Cin, Cout;
  int act;
  int out_ndhwc, res_ndhwc;
  int items, hblocks;
};

// ------------------------------------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, SWIZZLE_64B shared-memory matrix descriptor: 8-row atoms of 512 bytes (SBO), version 1.
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory carve-up (all TMA / UMMA buffers 1024-byte aligned).
struct TcSmem {
  static constexpr int A_HI = 0;                                        // [STAGES][8192]
  static constexpr int A_LO = A_HI + TC_STAGES * TC_ROW_BYTES;          // [STAGES][8192]
  static constexpr int B_OFF = A_LO + TC_STAGES * TC_ROW_BYTES;         // [2][ hi: 3 kh x N3 rows x 64 B | lo: same ]
};

template <int COUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
    conv3d_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  constexpr int N3 = 3 * COUT;                      // kw-stacked MMA N
  constexpr int B_KH_BYTES = N3 * TC_KC * 4;        // one kh slice of a phase's weights (hi or lo)
  constexpr int B_HALF = 3 * B_KH_BYTES;            // hi (or lo) part of a phase
  constexpr int B_PHASE = 2 * B_HALF;
  static_assert(B_KH_BYTES % 1024 == 0, "weight slices must stay 1024-byte aligned");
  static_assert(TC_TILES * N3 <= 512, "accumulators exceed TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_hi = smem + TcSmem::A_HI;
  uint8_t* a_lo = smem + TcSmem::A_LO;
  uint8_t* b_buf = smem + TcSmem::B_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + 2 * B_PHASE);
  uint64_t* a_full = bars;                          // [STAGES] TMA -> converters
  uint64_t* a_ready = bars + TC_STAGES;             // [STAGES] converters -> MMA
  uint64_t* a_empty = bars + 2 * TC_STAGES;         // [STAGES] MMA -> TMA
  uint64_t* b_full = bars + 3 * TC_STAGES;          // [2]
  uint64_t* b_empty = b_full + 2;                   // [2]
  uint64_t* acc_full = b_empty + 2;                 // [1] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;               // [1] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* xchg = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][2][16] shuffle boundary exchange
  float* s_scale = xchg + 4 * 2 * 16;                      // [COUT]
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = p.Cin / TC_KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_ready[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full[0], 1), mbar_init(&b_full[1], 1);
    mbar_init(&b_empty[0], 1), mbar_init(&b_empty[1], 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int c = threadIdx.x; c < COUT; c += blockDim.x) {
    s_scale[c] = p.scale ? p.scale[c] : 1.f;
    s_shift[c] = p.shift ? p.shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---------------------------------------------------------------------------------------------- TMA producer
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_x);
      tma_prefetch_desc(&map_w);
      uint32_t rowc = 0, phc = 0;                    // running counters: ring slot / weight buffer + parity
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        const int hb = it % p.hblocks;
        const int d = (it / p.hblocks) % p.D;
        const int b = it / (p.hblocks * p.D);
        const int h0 = hb * TC_TILES;
        for (int kd = 0; kd < 3; ++kd) {
          const int din = d + kd - 1;
          if (din < 0 || din >= p.D) continue;
          for (int ch = 0; ch < nchunk; ++ch) {
            const uint32_t j = phc & 1, jpar = (phc >> 1) & 1;
            mbar_wait(&b_empty[j], jpar ^ 1);
            mbar_arrive_expect_tx(&b_full[j], B_PHASE);
            // weight tensor rows: [half(hi,lo)][kd][chunk][kh][N3]
            for (int half = 0; half < 2; ++half)
              for (int kh = 0; kh < 3; ++kh)
                tma_load_2d(b_buf + j * B_PHASE + half * B_HALF + kh * B_KH_BYTES, &map_w, &b_full[j], 0,
                            (((half * 3 + kd) * nchunk + ch) * 3 + kh) * N3);
            ++phc;
            for (int r = 0; r < TC_ROWS; ++r) {
              const uint32_t s = rowc % TC_STAGES, par = (rowc / TC_STAGES) & 1;
              mbar_wait(&a_empty[s], par ^ 1);
              mbar_arrive_expect_tx(&a_full[s], TC_ROW_BYTES);
              // x viewed as (Cin, W, H, B*D): rows outside [0,H) are zero-filled by TMA = the conv's padding
              tma_load_4d(a_hi + s * TC_ROW_BYTES, &map_x, &a_full[s], ch * TC_KC, 0, h0 - 1 + r, b * p.D + din);
              ++rowc;
            }
          }
        }
      }
    }
  }
  // ---------------------------------------------------------------------------------------------- MMA issuer
  else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, N3);
      uint32_t rowc = 0, phc = 0, itc = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++itc) {
        const int hb = it % p.hblocks;
        const int d = (it / p.hblocks) % p.D;
        const int ntiles = min(TC_TILES, p.H - hb * TC_TILES);
        mbar_wait(acc_empty, (itc & 1) ^ 1);          // epilogue of the previous item has drained TMEM
        tc_fence_after();
        uint32_t started = 0;
        for (int kd = 0; kd < 3; ++kd) {
          const int din = d + kd - 1;
          if (din < 0 || din >= p.D) continue;
          for (int ch = 0; ch < nchunk; ++ch) {
            const uint32_t j = phc & 1, jpar = (phc >> 1) & 1;
            mbar_wait(&b_full[j], jpar);
            tc_fence_after();
            const uint32_t bh = smem_u32(b_buf + j * B_PHASE), bl = bh + B_HALF;
            for (int r = 0; r < TC_ROWS; ++r) {
              const uint32_t s = rowc % TC_STAGES, par = (rowc / TC_STAGES) & 1;
              mbar_wait(&a_ready[s], par);
              tc_fence_after();
              const uint32_t ah = smem_u32(a_hi + s * TC_ROW_BYTES), al = smem_u32(a_lo + s * TC_ROW_BYTES);
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                const int t = r - kh;                 // output row tile fed by input row r through tap kh
                if (t < 0 || t >= ntiles) continue;
                const uint32_t acc = tmem + t * N3;
#pragma unroll
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
                  const uint64_t dah = desc_sw64(ah + ks * 32), dal = desc_sw64(al + ks * 32);
                  const uint64_t dbh = desc_sw64(bh + kh * B_KH_BYTES + ks * 32), dbl = desc_sw64(bl + kh * B_KH_BYTES + ks * 32);
                  mma_tf32(acc, dal, dbh, idesc, (started >> t) & 1);   // small terms first
                  mma_tf32(acc, dah, dbl, idesc, 1);
                  mma_tf32(acc, dah, dbh, idesc, 1);
                  started |= 1u << t;
                }
              }
              mma_commit(&a_empty[s]);                // ring slot reusable once these MMAs have read it
              ++rowc;
            }
            mma_commit(&b_empty[j]);
            ++phc;
          }
        }
        mma_commit(acc_full);
      }
    }
  }
  // ---------------------------------------------------------------------------------------------- converters + epilogue
  else {
    const int ct = threadIdx.x - 64;                 // 0..127
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;                     // voxel (image column) owned in the epilogue
    uint32_t rowc = 0, itc = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++itc) {
      const int hb = it % p.hblocks;
      const int d = (it / p.hblocks) % p.D;
      const int b = it / (p.hblocks * p.D);
      const int h0 = hb * TC_TILES;
      const int ntiles = min(TC_TILES, p.H - h0);
      // ---- lo split of every staged row of this item
      for (int kd = 0; kd < 3; ++kd) {
        const int din = d + kd - 1;
        if (din < 0 || din >= p.D) continue;
        for (int ch = 0; ch < nchunk; ++ch) {
          for (int r = 0; r < TC_ROWS; ++r) {
            const uint32_t s = rowc % TC_STAGES, par = (rowc / TC_STAGES) & 1;
            mbar_wait(&a_full[s], par);
            const float4* src = reinterpret_cast<const float4*>(a_hi + s * TC_ROW_BYTES);
            float4* dst = reinterpret_cast<float4*>(a_lo + s * TC_ROW_BYTES);
#pragma unroll
            for (int i = 0; i < TC_ROW_BYTES / 16 / 128; ++i) {
              const float4 a = src[ct + i * 128];
              float4 l;
              l.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xffffe000u);
              l.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xffffe000u);
              l.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xffffe000u);
              l.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xffffe000u);
              dst[ct + i * 128] = l;                  // position preserving: the swizzle of the hi tile carries over
            }
            fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(&a_ready[s]);
            ++rowc;
          }
        }
      }
      // ---- epilogue: D[m] = P0[m-1] + P1[m] + P2[m+1]
      mbar_wait(acc_full, itc & 1);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) {
        const int h = h0 + t;
        const size_t vox = (((size_t)b * p.D + d) * p.H + h) * TC_W + m;           // NDHWC voxel index
        const size_t plane = (size_t)p.D * p.H * TC_W;                             // NCDHW channel stride
        const size_t ncdhw0 = (size_t)b * COUT * plane + ((size_t)d * p.H + h) * TC_W + m;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + t * N3;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          float p0[16], p1[16], p2[16];
          tmem_ld16(trow + 0 * COUT + c0, p0);
          tmem_ld16(trow + 1 * COUT + c0, p1);
          tmem_ld16(trow + 2 * COUT + c0, p2);
          // lanes at the warp edges need the neighbour warp's values
          float* myx = xchg + (q * 2) * 16;
          named_bar_sync(1, 128);                     // previous use of xchg finished
          if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 16; ++i) myx[i] = p0[i];
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) myx[16 + i] = p2[i];
          }
          named_bar_sync(1, 128);
          float out[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float left = __shfl_up_sync(0xffffffffu, p0[i], 1);
            float right = __shfl_down_sync(0xffffffffu, p2[i], 1);
            if (lane == 0) left = (q > 0) ? xchg[((q - 1) * 2) * 16 + i] : 0.f;          // m-1 of the previous quadrant
            if (lane == 31) right = (q < 3) ? xchg[((q + 1) * 2) * 16 + 16 + i] : 0.f;   // m+1 of the next quadrant
            out[i] = (left + p1[i]) + right;
            out[i] = fmaf(out[i], s_scale[c0 + i], s_shift[c0 + i]);
          }
          if (p.residual) {
            if (p.res_ndhwc) {
              const float4* rp = reinterpret_cast<const float4*>(p.residual + vox * COUT + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 rv = __ldg(rp + i);
                out[4 * i] += rv.x, out[4 * i + 1] += rv.y, out[4 * i + 2] += rv.z, out[4 * i + 3] += rv.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) out[i] += __ldg(p.residual + ncdhw0 + (size_t)(c0 + i) * plane);
            }
          }
          if (p.act == OSB_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) out[i] = fmaxf(out[i], 0.f);
          } else if (p.act == OSB_ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 16; ++i) out[i] = out[i] > 0.f ? out[i] : 0.01f * out[i];
          }
          if (p.out_ndhwc) {
            float4* yp = reinterpret_cast<float4*>(p.y + vox * COUT + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) yp[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) p.y[ncdhw0 + (size_t)(c0 + i) * plane] = out[i];   // 128-byte rows per warp
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------- layout conversion kernels
// NCDHW -> NDHWC through a 32x32 shared-memory transpose (both sides coalesced).
__global__ void __launch_bounds__(256) ncdhw_to_ndhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, size_t vol) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const size_t v0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const size_t v = v0 + tx;
    tile[i][tx] = (c < C && v < vol) ? __ldg(x + ((size_t)b * C + c) * vol + v) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const size_t v = v0 + i;
    const int c = c0 + tx;
    if (c < C && v < vol) y[((size_t)b * vol + v) * C + c] = tile[tx][i];
  }
}

// host-side tensor-map factory for up to 4 dims (fp32, SWIZZLE_64B)
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool encode_sw64(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box) {
  static EncodeTiledFn4 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return false;
    }
    fn = reinterpret_cast<EncodeTiledFn4>(ptr);
  }
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (SWIZZLE_64B, rank %d) failed with CUresult %d", rank, (int)r);
    return false;
  }
  return true;
}

template <int COUT>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t stream) {
  constexpr int N3 = 3 * COUT;
  const size_t smem = 1024 + 2 * (size_t)TC_STAGES * TC_ROW_BYTES + 2 * (size_t)(2 * 3 * N3 * TC_KC * 4) + 256 + 4 * 2 * 16 * 4 +
                      2 * COUT * 4;
  auto kernel = conv3d_tc_kernel<COUT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3d_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return OSB_ECUDA;
    }
    configured = true;
  }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    sms = 148;
  }
  const int grid = p.items < sms ? p.items : sms;   // persistent: one CTA per SM (it owns all 512 TMEM columns)
  kernel<<<grid, TC_THREADS, smem, stream>>>(mx, mw, p);
  count_launch();
  return check_launch("conv3d_tc_kernel");
}

}  // namespace osb

extern "C" {

int osb_conv3d_tc_supported(int Cin, int Cout, int W, int stride) {
  return (W == osb::TC_W && stride == 1 && Cin % osb::TC_KC == 0 && Cin >= osb::TC_KC && (Cout == 32)) ? 1 : 0;
}

int osb_ncdhw_to_ndhwc(const float* x, float* y, int B, int C, int D, int H, int W, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(x && y, "ncdhw_to_ndhwc: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0 && B <= 65535, "ncdhw_to_ndhwc: bad shape");
  const size_t vol = (size_t)D * H * W;
  dim3 grid((unsigned)((vol + 31) / 32), (C + 31) / 32, B);
  ncdhw_to_ndhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, C, vol);
  count_launch();
  return check_launch("ncdhw_to_ndhwc_kernel");
}

int osb_conv3d_k3_tc_fwd(const float* x_ndhwc, const float* w_split, const float* scale, const float* shift,
                         const float* residual, float* y, int B, int Cin, int Cout, int D, int H, int W, int act,
                         int out_ndhwc, int res_ndhwc, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(x_ndhwc && w_split && y, "conv3d_k3_tc: null pointer");
  OSB_REQUIRE(B > 0 && D > 0 && H > 0, "conv3d_k3_tc: empty shape");
  OSB_REQUIRE(osb_conv3d_tc_supported(Cin, Cout, W, 1), "conv3d_k3_tc: unsupported shape Cin=%d Cout=%d W=%d (needs W=128, Cin%%16==0, Cout=32)",
              Cin, Cout, W);
  OSB_REQUIRE(act >= 0 && act <= 2, "conv3d_k3_tc: unknown activation %d", act);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(x_ndhwc) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_split) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
              "conv3d_k3_tc: pointers must be 16-byte aligned");
  const int nchunk = Cin / TC_KC, N3 = 3 * Cout;
  CUtensorMap mx, mw;
  {  // x: (B*D, H, W, Cin) channels-last; box = 16 channels x 128 columns x 1 row x 1 plane
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
    const cuuint32_t box[4] = {TC_KC, TC_W, 1, 1};
    if (!encode_sw64(&mx, x_ndhwc, 4, dims, strides, box)) return OSB_ECUDA;
  }
  {  // weights: rows [half][kd][chunk][kh][kw*Cout + co], 16 input channels per row
    const cuuint64_t rows = (cuuint64_t)2 * 3 * nchunk * 3 * N3;
    const cuuint64_t dims[2] = {TC_KC, rows};
    const cuuint64_t strides[1] = {TC_KC * 4};
    const cuuint32_t box[2] = {TC_KC, (cuuint32_t)N3};
    if (!encode_sw64(&mw, w_split, 2, dims, strides, box)) return OSB_ECUDA;
  }
  TcParams p{};
  p.scale = scale, p.shift = shift, p.residual = residual, p.y = y;
  p.B = B, p.D = D, p.H = H, p.Cin = Cin, p.Cout = Cout, p.act = act;
  p.out_ndhwc = out_ndhwc, p.res_ndhwc = res_ndhwc;
  p.hblocks = (H + TC_TILES - 1) / TC_TILES;
  const long long items = (long long)B * D * p.hblocks;
  OSB_REQUIRE(items < (1ll << 31), "conv3d_k3_tc: too many work items");
  p.items = (int)items;
  return launch_tc<32>(mx, mw, p, (cudaStream_t)stream);
}
}
