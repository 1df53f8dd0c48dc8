// Tensor-core (tcgen05) ConvTranspose3d k=3, stride 2, padding 1, output_padding 1 with ONE WORK ITEM PER OUTPUT-PARITY QUAD:
// the variant of conv3d_tcdc.cu for the layer that dominated the transposed convs, conv6 64 -> 32 at 1/8 -> 1/4 resolution
//   gwcnet/hourglass.py:35-41 (conv6), psmnet/psmnet_cost_processor.py:99-106.
// conv3d_tcdc.cu gives every output parity class (od & 1, oh & 1) its own work item and stages the input unit of every (kd, kh)
// tap pair again for each of them: 1 + 2 + 2 + 4 = 9 unit stagings (global loads, fp32 -> [hi | lo] fp16 conversion, swizzled
// stores, one mbarrier round trip each) per group of four class tiles, 3 MMAs per staged unit -- the kernel is bound by its
// loaders, not by the tensor pipe (12 % busy).  Here an item is (image b, INPUT plane i, tile of R input
// rows j0..) and owns the four class tiles it feeds, od in {2i, 2i+1} x oh parity {0, 1}, as four TMEM accumulators:
//     unit                      MMA groups (class = (od & 1) * 2 + (oh & 1); tap (kd, kh))
//     U0  plane i,   rows j0    c0 (1,1)   c1 (1,2)   c2 (2,1)   c3 (2,2)
//     U1  plane i,   rows j0+1             c1 (1,0)              c3 (2,0)
//     U2  plane i+1, rows j0                          c2 (0,1)   c3 (0,2)        (plane i+1 exists)
//     U3  plane i+1, rows j0+1                                   c3 (0,0)        (plane i+1 exists)
// -- the same 9 tap groups from 4 staged units (2.25x fewer loads / conversions / hand-shakes, up to 12 MMAs per unit).
// Per dimension, output index o gathers tap k from input (o + 1 - k) / 2: o = 2m -> k = 1 (input m); o = 2m+1 -> k = 2 (input m)
// and k = 0 (input m+1); in w the three kw slices are stacked along N as in conv3d_tcdc.cu ([W1 | W2 | W0], even output column =
// P1[m], odd = P2[m] + P0[m+1]).  A phase is one 16-channel chunk; its nine (kd, kh) weight slices (54 KB for Cout = 32) arrive by
// three 1-D TMA bulk copies into one of two buffer sets.  Epilogue, operand staging, 3xFP16 split: tc_common.cuh / conv3d_tcdc.cu.
#include "tc_common.cuh"

namespace osb {

struct TcdqParams {
  const float* x;          // (B, D, H, W, Cin) channels-last
  const void* w;           // fp16 [3 kd][Cin/16][3 kh][3*Cout (kw order 1,2,0)][16 hi | 16 lo]  (ops.pack_tc_deconv_weight)
  const float* scale;
  const float* shift;
  const float* residual;   // (B, 2D, 2H, 2W, Cout) channels-last or nullptr
  float* y;                // (B, 2D, 2H, 2W, Cout) channels-last
  int B, D, H, Cin;        // INPUT extent D x H x W
  int act;
  float kappa;             // expected round-towards-zero loss per accumulating MMA (tc_common.cuh)
  unsigned int* overflow;  // sticky fp16-range flag (tc_common.cuh)
  int items, rtiles;       // work items, row tiles per plane
};

template <int COUT, int W>     // W = INPUT width
struct TcdqCfg {
  static constexpr int KC = 16;
  static constexpr int R = 128 / W;                         // input rows per M tile
  static constexpr int ROWB = KC * 4;                       // bytes per K-major operand row: [16 fp16 hi | 16 fp16 lo]
  static constexpr int UNIT_BYTES = 128 * ROWB;
  static constexpr int N3 = 3 * COUT;
  static constexpr int NCLS = 4;                            // accumulators: (od parity, row parity)
  static constexpr int B_SLICE = N3 * ROWB;                 // one (kd, kh) weight slice
  static constexpr int B_SET = 9 * B_SLICE;                 // the nine slices of one chunk
  static constexpr int NLW = 5;                             // loader warps (1-4 and 10), units round-robin
  static constexpr int OTHER_SMEM = 1024 + 1024 + 2 * 4 * 2 * 32 * 4 + 3 * COUT * 4 + TP_BYTES;
  static constexpr int STAGES = (232448 - OTHER_SMEM - 2 * B_SET) / UNIT_BYTES < 10 ? (232448 - OTHER_SMEM - 2 * B_SET) / UNIT_BYTES : 10;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + STAGES * UNIT_BYTES;
  static constexpr int BAR_OFF = B_OFF + 2 * B_SET;
  static constexpr int THREADS = 32 + 128 + 128 + 64;       // MMA | A loaders | epilogue | weight producer + 5th loader (11 warps)
  static constexpr size_t SMEM = 1024 + (size_t)BAR_OFF + 1024 + 2 * 4 * 2 * 32 * 4 + 3 * COUT * 4 + TP_BYTES;
  static_assert(STAGES >= NLW, "the ring must hold at least one unit per loader warp");
  static_assert(SMEM <= 232448, "shared memory budget of one CTA exceeded");
  static_assert(NCLS * N3 <= 512, "accumulators exceed TMEM");
  static_assert(B_SLICE % 1024 == 0 && UNIT_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned");
  static_assert(N3 % 16 == 0 && N3 <= 256 && W >= 32 && 128 % W == 0, "invalid tile shape");
};

template <int COUT, int W>
__global__ void __launch_bounds__(TcdqCfg<COUT, W>::THREADS, 1) conv3d_tcdq_kernel(const TcdqParams p) {
  using C = TcdqCfg<COUT, W>;
  constexpr int KC = C::KC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_buf = smem + C::A_OFF;
  uint8_t* b_buf = smem + C::B_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* a_ready = bars;                         // [STAGES] loaders -> MMA        (32 arrivals: one warp)
  uint64_t* a_empty = a_ready + C::STAGES;          // [STAGES] MMA -> loaders        (tcgen05.commit)
  uint64_t* b_full = a_empty + C::STAGES;           // [2]      weight producer -> MMA (expect_tx + TMA bytes)
  uint64_t* b_empty = b_full + 2;                   // [2]      MMA -> weight producer (tcgen05.commit)
  uint64_t* acc_full = b_empty + 2;                 // [4]      MMA -> epilogue, one per class accumulator
  uint64_t* acc_empty = acc_full + C::NCLS;         // [4]      epilogue -> MMA       (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + C::NCLS);
  float* xchg = reinterpret_cast<float*>(smem + C::BAR_OFF + 1024);   // [2][4 quadrants][2 sides][32]
  float* s_scale = xchg + 2 * 4 * 2 * 32;
  float* s_shift = s_scale + COUT;
  float* zeros = s_shift + COUT;
  float* tpose = zeros + COUT;                      // [4 warps][32][TP_STRIDE] transpose tiles of the epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = p.Cin / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&a_ready[s], 32);
      mbar_init(&a_empty[s], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&b_full[k], 1);
      mbar_init(&b_empty[k], 1);
    }
    for (int c = 0; c < C::NCLS; ++c) {
      mbar_init(&acc_full[c], 1);
      mbar_init(&acc_empty[c], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int c = threadIdx.x; c < COUT; c += blockDim.x) {
    s_scale[c] = p.scale ? p.scale[c] : 1.f;
    s_shift[c] = p.shift ? p.shift[c] : 0.f;
    zeros[c] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // item -> (b, input plane i, first input row j0); row tiles vary fastest
  auto decode = [&](int it, int& b, int& i, int& j0) {
    j0 = (it % p.rtiles) * C::R;
    it /= p.rtiles;
    i = it % p.D;
    b = it / p.D;
  };

  // ---------------------------------------------------------------------------------------------- MMA issuer
  if (warp == 0) {
    const uint32_t idesc = idesc_f16(128, C::N3);
    const uint64_t dbase = desc_sw64_base();
    const uint32_t b16 = (smem_u32(b_buf) & 0x3FFFF) >> 4;
    uint32_t unitc = 0, phc = 0, itc = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++itc) {
      int b, i, j0;
      decode(it, b, i, j0);
      const bool has_next = i + 1 < p.D;
      uint32_t started = 0;
      for (int ch = 0; ch < nchunk; ++ch, ++phc) {
        const uint32_t set = phc & 1;
        mbar_wait(&b_full[set], (phc >> 1) & 1);
        tc_fence_after();
        // one tap group: unit in ring slot `slot` times weight slice (kd, kh) into the accumulator of class c
        auto group = [&](uint32_t slot, int c, int kd, int kh) {
          const uint32_t accum = (started >> c) & 1;
          if (!accum) {                                 // first touch of this class in this item: the previous item's epilogue
            mbar_wait(&acc_empty[c], (itc & 1) ^ 1);    // must have drained its accumulator
            tc_fence_after();
            started |= 1u << c;
          }
          if (elect_one()) {
            const uint64_t da0 = dbase | (uint64_t)((smem_u32(a_buf + slot * C::UNIT_BYTES) & 0x3FFFF) >> 4);
            const uint32_t acc = tmem + c * C::N3;
            const uint64_t db0 = dbase | (uint64_t)(b16 + ((set * 9 + kd * 3 + kh) * C::B_SLICE) / 16);
            mma_f16(acc, da0 + 2, db0, idesc, accum);   // a_lo * b_hi first (small terms first); LO offset = KC / 8 = 2
            mma_f16(acc, da0, db0 + 2, idesc, 1);       // a_hi * b_lo
            mma_f16(acc, da0, db0, idesc, 1);           // a_hi * b_hi
          }
          __syncwarp();
        };
        auto unit_begin = [&]() -> uint32_t {
          const uint32_t slot = unitc % C::STAGES, par = (unitc / C::STAGES) & 1;
          mbar_wait(&a_ready[slot], par);
          tc_fence_after();
          return slot;
        };
        auto unit_end = [&](uint32_t slot) {
          if (elect_one()) mma_commit(&a_empty[slot]);
          __syncwarp();
          ++unitc;
        };
        uint32_t s = unit_begin();                      // U0: plane i, rows j0
        group(s, 0, 1, 1), group(s, 1, 1, 2), group(s, 2, 2, 1), group(s, 3, 2, 2);
        unit_end(s);
        s = unit_begin();                               // U1: plane i, rows j0 + 1
        group(s, 1, 1, 0), group(s, 3, 2, 0);
        unit_end(s);
        if (has_next) {
          s = unit_begin();                             // U2: plane i + 1, rows j0
          group(s, 2, 0, 1), group(s, 3, 0, 2);
          unit_end(s);
          s = unit_begin();                             // U3: plane i + 1, rows j0 + 1
          group(s, 3, 0, 0);
          unit_end(s);
        }
        if (elect_one()) {
          mma_commit(&b_empty[set]);                    // this chunk's slices are free once these MMAs have read them
          if (ch == nchunk - 1) {
#pragma unroll
            for (int c = 0; c < C::NCLS; ++c) mma_commit(&acc_full[c]);
          }
        }
        __syncwarp();
      }
    }
  }
  // ---------------------------------------------------------------------------------------------- A-unit loaders
  // Units round-robin over the NLW loader warps (unit u -> warp u % NLW, ring slot u % STAGES); each warp enumerates only its own.
  else if (warp < 5 || warp == 10) {
    const int lw = warp < 5 ? warp - 1 : 4;
    constexpr int CPR = KC / 4;                      // fp32 16-byte chunks per voxel of the K chunk
    constexpr int VPL = 32 / CPR;                    // voxels covered by one warp-wide LDG.128
    constexpr int NLD = 128 / VPL;                   // loads per lane per unit
    static_assert(W % VPL == 0, "a load instruction must not straddle image rows");
    const int v0 = lane_voxel<KC>(lane), c = lane % CPR;   // permuted voxel order: conflict-free STS.64 (tc_common.cuh)
    float amax = 0.f;
    uint32_t ubase = 0;                              // global index of the current phase's first unit
    int first = lw;                                  // this warp's first local unit index in the current phase
    for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
      int b, i, j0;
      decode(it, b, i, j0);
      const int upp = (i + 1 < p.D) ? 4 : 2;         // units per chunk phase
      for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll 1
        for (int j = first; j < upp; j += C::NLW) {  // local unit j: plane i + (j >> 1), rows j0 + (j & 1) ...
          const int h_first = j0 + (j & 1);
          const float* base = p.x + ((((size_t)b * p.D + i + (j >> 1)) * p.H + h_first) * (size_t)W + v0) * p.Cin + ch * KC + c * 4;
          float4 v[NLD];
#pragma unroll
          for (int l = 0; l < NLD; ++l) {
            const int hin = h_first + (VPL * l) / W;
            const size_t off = ((size_t)((VPL * l) / W) * W + (size_t)((VPL * l) % W)) * p.Cin;
            v[l] = hin < p.H ? __ldg(reinterpret_cast<const float4*>(base + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const uint32_t u = ubase + j, slot = u % C::STAGES, par = (u / C::STAGES) & 1;
          mbar_wait_relaxed(&a_empty[slot], par ^ 1);
          uint8_t* tile = a_buf + slot * C::UNIT_BYTES;
#pragma unroll
          for (int l = 0; l < NLD; ++l) stage_f16_split<KC>(tile, v0 + VPL * l, c, v[l], amax);
          fence_proxy_async();
          mbar_arrive(&a_ready[slot]);
        }
        ubase += upp;
        first = (first + C::NLW - upp % C::NLW) % C::NLW;
      }
    }
    tc_report_overflow(p.overflow, amax);
  }
  // ---------------------------------------------------------------------------------------------- epilogue
  else if (warp < 9) {
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;                     // operand row owned by this thread
    const int rr = m / W, wcol = m % W;              // input row inside the tile, input column
    const bool has_right_q = (((q + 1) * 32) % W) != 0;   // the next quadrant continues the same image row
    const int Do = 2 * p.D, Ho = 2 * p.H;
    constexpr int Wo = 2 * W;
    uint32_t itc = 0, exc = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++itc) {
      int b, i, j0;
      decode(it, b, i, j0);
      const bool has_next = i + 1 < p.D;
      const int j = j0 + rr;
      const bool live = j < p.H;
#pragma unroll 1
      for (int c = 0; c < C::NCLS; ++c) {
        const int eo = c >> 1, ph = c & 1;
        const int od = 2 * i + eo, oh = 2 * j + ph;
        // tap groups this class accumulated: (kd taps: even plane 1, odd plane 1 or 2) x (kh taps: 1 or 2), each chunks x 3 MMAs
        const float corr = 1.f + p.kappa * (float)((eo && has_next ? 2 : 1) * (ph + 1) * nchunk * 3);
        const size_t vox = (((size_t)b * Do + od) * Ho + oh) * Wo + 2 * wcol;        // NDHWC index of the EVEN output voxel
        if (live && p.residual) {                       // pull the residual towards L2 while the accumulators are still filling
          const float* rp = p.residual + vox * COUT;
#pragma unroll
          for (int k = 0; k < 2 * COUT; k += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + k));
        }
        mbar_wait_relaxed(&acc_full[c], itc & 1);
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + c * C::N3;
#pragma unroll 1
        for (int cg = 0; cg < COUT; cg += 32) {
          // accumulator column groups: raw[0] = E (kw=1), raw[1] = P2 (kw=2), raw[2] = P0 (kw=0)
          uint32_t raw[3][32];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) tmem_ld16_nowait(trow + kw * COUT + cg + c0, &raw[kw][c0]);
          tmem_ld_wait();
          if (cg + 32 >= COUT) {                        // whole accumulator in registers: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&acc_empty[c]);
          }
          float* xb = xchg + (exc & 1) * (4 * 2 * 32);
          ++exc;
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) xb[(q * 2) * 32 + k] = __uint_as_float(raw[2][k]);
          }
          named_bar_sync(1, 128);
          const float* xr = has_right_q ? xb + ((q + 1) * 2) * 32 : zeros;
          float ev[32], od_[32];
#pragma unroll
          for (int i0 = 0; i0 < 32; i0 += 4) {          // neighbour values loaded unconditionally, merged with selects (no branches)
            const float4 r4 = *reinterpret_cast<const float4*>(xr + i0);
            const float re[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = i0 + k;
              float right = __shfl_down_sync(0xffffffffu, __uint_as_float(raw[2][e]), 1);   // P0 of input column m+1
              right = (lane == 31) ? re[k] : right;                                        // zero beyond the last input column
              ev[e] = __uint_as_float(raw[0][e]) * corr;
              od_[e] = (__uint_as_float(raw[1][e]) + right) * corr;
            }
          }
          if (live) {
            // lane k owns output voxels (vox0 + 2k) and (vox0 + 2k + 1): two transposes with a 2-voxel lane stride
            float* y0 = p.y + (vox - 2 * lane) * COUT + cg;
            const float* r0 = p.residual ? p.residual + (vox - 2 * lane) * COUT + cg : nullptr;
            store_ndhwc_chunk32(tpose + q * TP_WARP_FLOATS, lane, ev, y0, r0, 2 * COUT, s_scale + cg, s_shift + cg, p.act);
            store_ndhwc_chunk32(tpose + q * TP_WARP_FLOATS, lane, od_, y0 + COUT, r0 ? r0 + COUT : nullptr, 2 * COUT, s_scale + cg,
                                s_shift + cg, p.act);
          }
        }
      }
    }
  }
  // ---------------------------------------------------------------------------------------------- weight-slice producer
  // One elected lane streams the nine pre-swizzled (kd, kh) slices of a chunk -- three contiguous runs of 3 * B_SLICE in the
  // [kd][chunk][kh] pack -- into the two buffer sets, one phase ahead of the MMAs.
  else if (warp == 9) {
    if (elect_one()) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
      uint32_t phc = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        for (int ch = 0; ch < nchunk; ++ch, ++phc) {
          const uint32_t set = phc & 1;
          mbar_wait_relaxed(&b_empty[set], ((phc >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&b_full[set], C::B_SET);
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
            bulk_g2s(b_buf + (set * 9 + kd * 3) * C::B_SLICE, wsrc + ((size_t)kd * nchunk + ch) * 3 * C::B_SLICE, 3 * C::B_SLICE,
                     &b_full[set]);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// Launcher used by conv3d_tcdc.cu's C entry point for the shapes instantiated here (channels-last output and residual only).
int launch_tcdq(const float* x, const void* w, const float* scale, const float* shift, const float* residual, float* y, int B,
                int Cin, int Cout, int D, int H, int W, int act, cudaStream_t stream) {
  if (!(Cout == 32 && W == 64 && Cin % 16 == 0 && Cin >= 16)) return -1;
  using C = TcdqCfg<32, 64>;
  TcdqParams p{};
  p.x = x, p.w = w, p.scale = scale, p.shift = shift, p.residual = residual, p.y = y;
  p.B = B, p.D = D, p.H = H, p.Cin = Cin, p.act = act;
  p.kappa = rz_kappa(), p.overflow = tc_overflow_flag();
  if (!p.overflow) return OSB_ECUDA;
  auto kernel = conv3d_tcdq_kernel<32, 64>;
  static PerDeviceFlag configured;
  if (!configured.here()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) {
      set_error("conv3d_tcdq: cannot reserve %zu bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return OSB_ECUDA;
    }
    configured.here() = true;
  }
  p.rtiles = (H + C::R - 1) / C::R;
  const long long items = (long long)B * D * p.rtiles;
  OSB_REQUIRE(items < (1ll << 31), "conv3d_tcdq: too many work items");
  p.items = (int)items;
  const int sms = sm_count();
  const int grid = p.items < sms ? p.items : sms;
  kernel<<<grid, C::THREADS, C::SMEM, stream>>>(p);
  count_launch();
  return check_launch("conv3d_tcdq_kernel");
}

}  // namespace osb
