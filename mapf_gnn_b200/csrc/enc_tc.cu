// Conv stack of the observation encoder on the 5th-generation tensor cores (reference
// models/framework_gnn.py:55-99,156-163: 3 x {conv3x3 pad 1 + BatchNorm(eval, folded) + ReLU} per agent image [2,5,5]).
//
// A 3x3 convolution over 16 channels is a GEMM with N = 16 per tap -- too thin for tcgen05, whose 128 x 16 x 8 MMA is
// bound by reading its A operand from shared memory.  Here each layer is ONE dense GEMM per 4 agent images,
//   D[pixel, (ky, co)] = sum_{kx, ci} act[pixel + kx - 1, ci] * W[co, ci, ky, kx]        128 x 48 x 48
// i.e. the horizontal taps sit in K (the A operand holds the pixel and its left / right neighbour, zero at the image
// border) and the filter rows sit in N; what is left for the epilogue is the vertical "col2im",
//   out[pixel, co] = bias[co] + D[pixel, (1, co)] + D[pixel - 5, (0, co)] + D[pixel + 5, (2, co)].
// The GEMM runs as 3xTF32 tcgen05.mma with the ACTIVATIONS AS THE A OPERAND IN TENSOR MEMORY (TS mode): lane = pixel and
// warp = image, so every neighbour exchange is a warp shuffle (64 per thread and layer: 32 for the vertical sums, 32 for
// the horizontal copies) and a layer's output never touches shared memory -- the epilogue threads read the 48 partial
// sums of their pixel from TMEM (tcgen05.ld), add the neighbours' shares, apply bias and ReLU, split the result and its
// two horizontal neighbours into TF32 hi / lo and write them straight back to TMEM (tcgen05.st) as the next layer's A
// operand.  The B operand (BN-folded filters, hi | lo, canonical no-swizzle K-major tiles, 40 KB for three layers) stays
// resident in shared memory; the tensor core reads only B from shared memory (1.5 KB per 24-cycle MMA).
//
// One M tile = 128 TMEM lanes = 4 agent images x 32 lanes (25 pixels + 7 zero lanes).  A CTA holds three independent
// warpgroups; each walks its own M tiles through all layers with its own 144 TMEM columns (D 48 | A hi 48 | A lo 48), its
// own mbarrier and named barrier, started a third of a layer apart so that one group's MMAs run under the others'
// epilogues.  The last layer's epilogue emits the flattened activations as the pre-split, pre-tiled A operand of the
// encoder FC (mapf_tc_fc_presplit) with the K index ordered pixel-major (k = pixel * 16 + channel); the four images of a
// tile meet in shared memory so that the global stores are whole 64-byte runs.  The FC's B operand is packed with the
// same K permutation (PackedLayout::tc_fc_pm).
#include <stdlib.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kEtGroups = 3;                    // epilogue warpgroups per CTA
constexpr int kEtThreads = 128 * kEtGroups + 32;   // + one warp whose lane 0 issues every tcgen05.mma
constexpr int kEtCout = 16;                     // channels of every conv layer this kernel is built for
constexpr int kEtN = 3 * kEtCout;               // GEMM N: (filter row, co)
constexpr int kEtBlk = kEtN * 8;                // floats of one (K step, hi|lo) block of the B operand
constexpr int kEtMaxK = 3 * kEtCout;            // widest A operand: (kx, ci)
constexpr uint32_t kEtColAHi = kEtN, kEtColALo = kEtN + kEtMaxK, kEtSetCols = kEtN + 2 * kEtMaxK;   // 48 | 48 | 48
constexpr int kEtOutPitch = 8 * 16 + 4;         // floats per pixel in the output staging buffer (8 runs x 64 B + 16 B)
constexpr int kEtOutFloats = kFovCells * kEtOutPitch;

struct EncTcArgs {
  int num_conv;
  int cin[MAPF_MAX_CONV];
  int w_off[MAPF_MAX_CONV];       // float offset of the layer's B blocks inside the shared-memory weight image
  int64_t bias_off[MAPF_MAX_CONV];// float offset of the layer's folded bias inside `packed`
  int w_floats;                   // floats of the B-operand image (all layers)
  int64_t rows;
  const float* states;
  const float* packed;            // whole packed buffer (biases)
  const float* wtc;               // packed + L.tc_conv
  float* ap;                      // pre-split FC operand out
  int nch;                        // 8-wide K chunks of the FC (fc_in / 8)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T (TS mode): A = 128 lanes x 8 TF32 columns at a_tmem, issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(uint32_t(accumulate))
      : "memory");
}

// TF32 hi / lo of v; hi rounded to nearest, lo = v - hi exact in fp32 (the tensor core drops its low 13 bits: a
// relative 2^-21 of v, the size of the lo*lo term 3xTF32 drops anyway)
__device__ __forceinline__ void split_fast(float v, uint32_t& hi, uint32_t& lo) {
  const float h = tc::tf32_rn(v);
  hi = __float_as_uint(h);
  lo = __float_as_uint(v - h);
}

// B operand of the conv GEMMs from the BN-folded filters of pack_weights_kernel ([Cout/4][Cin][9][4]):
// per layer, per 8-wide K step: hi block | lo block, each [n/8][k/4][n%8][k%4] with n = ky*16 + co, k = kx*Cin + ci
__global__ void pack_tc_conv_kernel(mapf_net_params p, PackedLayout L, float* __restrict__ packed) {
  int64_t base = 0;
  for (int l = 0; l < p.num_conv; ++l) {
    const int cin = p.channels[l], ks = (3 * cin + 7) >> 3;
    const int64_t n = int64_t(ks) * 2 * kEtBlk;
    for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
      int r = int(idx);
      const int step = r / (2 * kEtBlk);
      r -= step * 2 * kEtBlk;
      const int is_lo = r >= kEtBlk;
      if (is_lo) r -= kEtBlk;
      const int k4i = r & 3, n8 = (r >> 2) & 7, kq = (r >> 5) & 1, ng = r >> 6;
      const int nn = ng * 8 + n8, k = step * 8 + kq * 4 + k4i;
      const int ky = nn / kEtCout, co = nn % kEtCout;
      float v = 0.0f;
      if (k < 3 * cin) {
        const int kx = k / cin, ci = k - kx * cin;
        v = packed[L.conv_w[l] + ((int64_t(co >> 2) * cin + ci) * 9 + ky * 3 + kx) * 4 + (co & 3)];
      }
      float hi, lo;
      tc::split1(v, hi, lo);
      packed[L.tc_conv + base + idx] = is_lo ? lo : hi;
    }
    base += n;
  }
}

#ifdef MAPF_ENC_DEBUG
__device__ long long* g_etc_dbg = nullptr;   // block 0: [layer iteration < 16][warp < 12][6 stamps]
#define ETC_STAMP(i)                                                                                   \
  do {                                                                                                 \
    if (g_etc_dbg && blockIdx.x == 0 && lane == 0 && dbg_it < 16) g_etc_dbg[(dbg_it * 12 + warp) * 6 + (i)] = clock64(); \
  } while (0)
#else
#define ETC_STAMP(i)
#endif

__global__ void __launch_bounds__(kEtThreads, 1) encoder_tc_kernel(EncTcArgs g) {
  extern __shared__ __align__(128) float s_w[];          // B-operand image of all layers, then the output staging buffers
  __shared__ __align__(16) float s_bias[MAPF_MAX_CONV][kEtCout];
  __shared__ __align__(8) uint64_t s_bar[kEtGroups];     // MMAs of the group's current layer complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t s_ready[kEtGroups];   // the group's A operand is in TMEM and its D columns are free
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, quad = warp & 3;
  float* const s_out = s_w + g.w_floats + wg * kEtOutFloats;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    for (int i = 0; i < kEtGroups; ++i) {
      tc::mbar_init(&s_bar[i], 1);
      tc::mbar_init(&s_ready[i], 4);     // one arrival per warp of the group
    }
    tc::fence_mbar_init();
  }
  for (int i = tid; i < (g.w_floats >> 2); i += kEtThreads)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(g.wtc) + i);
  if (tid < g.num_conv * kEtCout) s_bias[tid / kEtCout][tid % kEtCout] = __ldg(g.packed + g.bias_off[tid / kEtCout] + tid % kEtCout);
  tc::fence_proxy_async();      // the filters are read by the tensor core (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  constexpr uint32_t idesc = tc::idesc_tf32(128, kEtN);
  const int64_t ntiles = (g.rows + 3) >> 2;
  const int64_t stride = int64_t(gridDim.x) * kEtGroups;

  if (wg == kEtGroups) {
    // ===== MMA issuer: ONE thread issues the MMAs of all groups, group after group.  The tensor pipe runs them in
    // issue order, so group 0's layer completes first and its epilogue runs under the MMAs of groups 1 and 2: the
    // groups stay a third of a round apart instead of contending for the pipe (and then for the issue slots) in step.
    if (lane == 0) {
      const uint32_t w_base = tc::smem_u32(s_w);
      uint32_t ph = 0;
      for (int64_t t0 = int64_t(blockIdx.x) * kEtGroups; t0 < ntiles; t0 += stride) {
        for (int l = 0; l < g.num_conv; ++l, ph ^= 1u) {
          const uint32_t wl = w_base + uint32_t(g.w_off[l]) * 4u;
          const int ksteps = (3 * g.cin[l] + 7) >> 3;
          const uint64_t bh0 = tc::smem_desc(wl, 128, 256), bl0 = tc::smem_desc(wl + kEtBlk * 4, 128, 256);
          for (int gi = 0; gi < kEtGroups; ++gi) {
            if (t0 + gi >= ntiles) break;
            const uint32_t tm = s_tmem + uint32_t(gi) * kEtSetCols;
            tc::mbar_wait(&s_ready[gi], ph);
            tc::fence_after_sync();
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t o = uint64_t(ks) * ((2 * kEtBlk * 4) >> 4);   // descriptor address field counts 16 bytes
              const uint32_t ah = tm + kEtColAHi + uint32_t(ks) * 8, al = tm + kEtColALo + uint32_t(ks) * 8;
              mma_tf32_ts(tm, al, bh0 + o, idesc, ks > 0);     // small terms first
              mma_tf32_ts(tm, ah, bl0 + o, idesc, true);
              mma_tf32_ts(tm, ah, bh0 + o, idesc, true);
            }
            tc::mma_commit(&s_bar[gi]);
          }
        }
      }
    }
    __syncwarp();
  } else {
  const uint32_t tmem = s_tmem + uint32_t(wg) * kEtSetCols;             // this warpgroup's columns, lane 0
  const uint32_t tlane = tmem + (uint32_t(quad * 32) << 16);            // this warp's 32 lanes
  uint64_t* const bar = &s_bar[wg];
  uint32_t phase = 0;

  // lane = pixel (y, x) of this warp's image
  const int py = lane / kFov, px = lane - py * kFov;
  const bool pix_ok = lane < kFovCells;
  const float m_up = (pix_ok && py > 0) ? 1.0f : 0.0f, m_dn = (pix_ok && py < kFov - 1) ? 1.0f : 0.0f;
  const uint32_t v_ok = pix_ok ? 0xFFFFFFFFu : 0u;
  const uint32_t v_lf = (pix_ok && px > 0) ? 0xFFFFFFFFu : 0u, v_rt = (pix_ok && px < kFov - 1) ? 0xFFFFFFFFu : 0u;

  auto load_obs = [&](int64_t tile, float& v0, float& v1) {
    const int64_t img = tile * 4 + quad;
    v0 = 0.0f;
    v1 = 0.0f;
    if (tile < ntiles && img < g.rows && pix_ok) {
      const float* s = g.states + img * kFovFloats + lane;
      v0 = __ldg(s);
      v1 = __ldg(s + kFovCells);
    }
  };
  auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); };
  // v of this pixel's left / right neighbour (zero outside the image)
  auto left_of = [&](float v) { return __uint_as_float(__float_as_uint(__shfl_up_sync(0xffffffffu, v, 1)) & v_lf); };
  auto right_of = [&](float v) { return __uint_as_float(__float_as_uint(__shfl_down_sync(0xffffffffu, v, 1)) & v_rt); };

#ifdef MAPF_ENC_DEBUG
  int dbg_it = 0;
#endif
  int64_t tile = int64_t(blockIdx.x) * kEtGroups + wg;
  float o0, o1;
  load_obs(tile, o0, o1);
  for (; tile < ntiles; tile += stride) {
    // ---- layer-0 A operand: k = kx*2 + c, K padded to 8 ----
    {
      uint32_t hi[8], lo[8];
      split_fast(left_of(o0), hi[0], lo[0]);
      split_fast(left_of(o1), hi[1], lo[1]);
      split_fast(o0, hi[2], lo[2]);
      split_fast(o1, hi[3], lo[3]);
      split_fast(right_of(o0), hi[4], lo[4]);
      split_fast(right_of(o1), hi[5], lo[5]);
      hi[6] = hi[7] = lo[6] = lo[7] = 0u;
      tmem_st8(tlane + kEtColAHi, hi);
      tmem_st8(tlane + kEtColALo, lo);
    }
    load_obs(tile + stride, o0, o1);     // next tile's pixels arrive while this one is computed
    const int64_t img0 = tile * 4;

    for (int l = 0; l < g.num_conv; ++l) {
      ETC_STAMP(0);
      tmem_st_wait();                    // this lane's A row is in TMEM (and its reads of D are long done)
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_ready[wg]);
      ETC_STAMP(1);
      tc::mbar_wait(bar, phase);
      phase ^= 1u;
      tc::fence_after_sync();
      ETC_STAMP(2);

      // ---- epilogue: out[p][co] = bias[co] + D[p][(1,co)] + D[p-5][(0,co)] + D[p+5][(2,co)] ----
      uint32_t r[3][kEtCout];
#pragma unroll
      for (int u = 0; u < 3; ++u) tmem_ld16(tlane + uint32_t(u * kEtCout), r[u]);
      tc::tmem_ld_wait();
      ETC_STAMP(3);
      float v[kEtCout];
      {
        const float4* b4 = reinterpret_cast<const float4*>(s_bias[l]);
#pragma unroll
        for (int q = 0; q < kEtCout / 4; ++q) {
          const float4 b = b4[q];
          v[4 * q] = b.x; v[4 * q + 1] = b.y; v[4 * q + 2] = b.z; v[4 * q + 3] = b.w;
        }
      }
#pragma unroll
      for (int c = 0; c < kEtCout; ++c) {
        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r[0][c]), kFov);
        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c]), kFov);
        float a = v[c] + __uint_as_float(r[1][c]);
        a = fmaf(up, m_up, a);
        a = fmaf(dn, m_dn, a);
        v[c] = __uint_as_float(__float_as_uint(fmaxf(a, 0.0f)) & v_ok);   // ReLU; lanes that are not pixels hold zero
      }
      ETC_STAMP(4);
      if (l + 1 < g.num_conv) {
        // next layer's A operand: k = kx*16 + ci -> (left neighbour | this pixel | right neighbour), hi and lo
        uint32_t h[kEtCout], o[kEtCout];
#pragma unroll
        for (int c = 0; c < kEtCout; ++c) split_fast(left_of(v[c]), h[c], o[c]);
        tmem_st16(tlane + kEtColAHi, h);
        tmem_st16(tlane + kEtColALo, o);
#pragma unroll
        for (int c = 0; c < kEtCout; ++c) split_fast(v[c], h[c], o[c]);
        tmem_st16(tlane + kEtColAHi + kEtCout, h);
        tmem_st16(tlane + kEtColALo + kEtCout, o);
#pragma unroll
        for (int c = 0; c < kEtCout; ++c) split_fast(right_of(v[c]), h[c], o[c]);
        tmem_st16(tlane + kEtColAHi + 2 * kEtCout, h);
        tmem_st16(tlane + kEtColALo + 2 * kEtCout, o);
      } else {
        // flattened activations, k = pixel*16 + c, as Ap [row/128][k/8][hi|lo][(row%128)/8][(k%8)/4][row%8][k%4]:
        // staged per pixel as 8 runs (chunk half, hi|lo, k quad) of 4 images x 16 bytes, then stored as 64-byte runs
        if (pix_ok) {
          float* dst = s_out + lane * kEtOutPitch + quad * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 hi4, lo4;
            tc::split4(make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]), hi4, lo4);
            *reinterpret_cast<float4*>(dst + (((q >> 1) * 2 + 0) * 2 + (q & 1)) * 16) = hi4;
            *reinterpret_cast<float4*>(dst + (((q >> 1) * 2 + 1) * 2 + (q & 1)) * 16) = lo4;
          }
        }
        wg_sync();
        float* const base = g.ap + (img0 >> 7) * int64_t(g.nch) * 2048 + ((img0 & 127) >> 3) * 64 + (img0 & 7) * 4;
        const int wt = tid & 127;
#pragma unroll
        for (int it = 0; it < (kFovCells * 32 + 127) / 128; ++it) {
          const int f = it * 128 + wt;
          const int p = f >> 5, run = (f >> 2) & 7, piece = f & 3;
          if (p < kFovCells && img0 + piece < g.rows) {
            const float4 val = *reinterpret_cast<const float4*>(s_out + p * kEtOutPitch + run * 16 + piece * 4);
            float* d = base + int64_t(2 * p + (run >> 2)) * 2048 + ((run >> 1) & 1) * 1024 + (run & 1) * 32 + piece * 4;
            *reinterpret_cast<float4*>(d) = val;
          }
        }
      }
      ETC_STAMP(5);
#ifdef MAPF_ENC_DEBUG
      ++dbg_it;
#endif
    }
  }
  }  // epilogue groups
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(s_tmem, 512);
}

}  // namespace

#ifdef MAPF_ENC_DEBUG
extern "C" int mapf_enc_tc_set_debug(long long* p) {
  return cudaMemcpyToSymbol(g_etc_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#endif

// every conv layer 16 channels wide, input 2 planes, FC input 400 = 50 whole K chunks
extern "C" int mapf_encoder_tc_supported(const mapf_net_params* p) {
  static_assert(kEtCout == 16 && kEtBlk == 48 * 8, "mapf_packed_layout sizes the tc_conv region with these");
  if (mapf_check_params(p) != MAPF_OK) return 0;
  return mapf_enc_tc_dims_ok(*p) ? 1 : 0;
}

// called by mapf_pack_weights behind pack_weights_kernel (same stream): needs the BN-folded filters in place
int mapf_encoder_tc_pack(const mapf_net_params* p, float* packed, mapf_stream_t stream) {
  const PackedLayout L = mapf_packed_layout(*p);
  pack_tc_conv_kernel<<<32, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p, L, packed);
  const int rc = mapf_tc_pack_b_pixel_major(p->fc_w, p->channels[p->num_conv], p->fc_out, packed + L.tc_fc_pm, stream);
  if (rc != MAPF_OK) return rc;
  return mapf_launch_status();
}

// Conv stack -> a->x = pre-split A operand of the FC in pixel-major K order (pair it with packed + tc_fc_pm).
extern "C" int mapf_encoder_tc_conv_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->x, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_encoder_tc_supported(p), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(a->packed, 16) && mapf_aligned(a->x, 128) && mapf_aligned(a->states, 4), MAPF_ERR_ALIGN);
  const PackedLayout L = mapf_packed_layout(*p);
  EncTcArgs g{};
  g.num_conv = p->num_conv;
  int off = 0;
  for (int l = 0; l < p->num_conv; ++l) {
    g.cin[l] = p->channels[l];
    g.w_off[l] = off;
    off += ((3 * p->channels[l] + 7) >> 3) * 2 * kEtBlk;
    g.bias_off[l] = L.conv_b[l];
  }
  g.w_floats = off;
  g.rows = a->rows;
  g.states = a->states;
  g.packed = a->packed;
  g.wtc = a->packed + L.tc_conv;
  g.ap = a->x;
  g.nch = (p->fc_in + 7) >> 3;
  const size_t smem = size_t(off + kEtGroups * kEtOutFloats) * sizeof(float);
  MAPF_REQUIRE(smem <= 200 * 1024, MAPF_ERR_UNSUPPORTED);
  static int cache[16] = {0};
  if (!mapf_ensure_smem(encoder_tc_kernel, smem, cache)) return MAPF_ERR_CUDA;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (a->rows + 3) >> 2;
  const int64_t want = (ntiles + kEtGroups - 1) / kEtGroups;
  const int grid = int(want < sms ? want : sms);
  encoder_tc_kernel<<<grid, kEtThreads, smem, static_cast<cudaStream_t>(stream)>>>(g);
  return mapf_launch_status();
}
