// Conv stack of the observation encoder on the 5th-generation tensor cores (reference
// models/framework_gnn.py:55-99,156-163: 3 x {conv3x3 pad 1 + BatchNorm(eval, folded) + ReLU} per agent image [2,5,5]).
//
// A 3x3 convolution over 16 channels is a GEMM with N = 16 per tap -- too thin for tcgen05, whose 128 x 16 x 8 MMA is
// bound by reading its A operand.  Here the convolution is turned around:
//   P[pixel, (tap, co)] = sum_ci act[pixel, ci] * W[co, ci, tap]          one 128 x 144 x Cin GEMM (all 9 taps at once)
//   out[pixel, co]      = bias[co] + sum_tap P[pixel + shift(tap), (tap, co)]   9 shifted adds ("col2im")
// The GEMM runs as 3xTF32 tcgen05.mma with the ACTIVATIONS AS THE A OPERAND IN TENSOR MEMORY (TS mode): the layer's
// output never touches shared memory -- the epilogue threads read the 144 partial sums of their pixel from TMEM
// (tcgen05.ld), fetch the neighbours' partial sums with warp shuffles (lane = pixel, warp = image, so every shift stays
// inside one warp), add bias, apply ReLU, split the result into TF32 hi / lo and write it straight back to TMEM
// (tcgen05.st) as the next layer's A operand.  The B operand (BN-folded filters, hi | lo, canonical no-swizzle K-major
// tiles, 46 KB for three layers) stays resident in shared memory.
//
// One M tile = 128 TMEM lanes = 4 agent images x 32 lanes (25 pixels + 7 zero lanes).  A CTA holds two independent
// warpgroups; each walks its own M tiles through all layers with its own TMEM columns (D 144 + A 32), its own mbarrier
// and its own named barrier, so one group's MMAs run under the other group's epilogue.  The last layer's epilogue
// writes the flattened activations as the pre-split, pre-tiled A operand of the encoder FC (mapf_tc_fc_presplit) with
// the K index ordered pixel-major (k = pixel * 16 + channel: 16 contiguous k per thread); the FC's B operand is packed
// with the same permutation (PackedLayout::tc_fc_pm).
#include <stdlib.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kEtThreads = 256;                 // two warpgroups
constexpr int kEtCout = 16;                     // channels of every conv layer this kernel is built for
constexpr int kEtN = 9 * kEtCout;               // GEMM N: (tap, co)
constexpr int kEtBlk = kEtN * 8;                // floats of one (K step, hi|lo) block of the B operand
constexpr uint32_t kEtColA = 144, kEtWgCols = 256;   // per warpgroup: D [0,144), A hi [144,144+16), A lo [160,176)
constexpr uint32_t kEtColALo = 160;

struct EncTcArgs {
  int num_conv;
  int ksteps[MAPF_MAX_CONV];      // 8-wide K steps of each layer (Cin padded to a multiple of 8)
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

// B operand of the conv GEMMs from the BN-folded filters of pack_weights_kernel ([Cout/4][Cin][9][4]):
// per layer, per 8-wide K step: hi block | lo block, each [n/8][k/4][n%8][k%4] with n = tap*16 + co, k = ci
__global__ void pack_tc_conv_kernel(mapf_net_params p, PackedLayout L, float* __restrict__ packed) {
  int64_t base = 0;
  for (int l = 0; l < p.num_conv; ++l) {
    const int cin = p.channels[l], ks = (cin + 7) >> 3;
    const int64_t n = int64_t(ks) * 2 * kEtBlk;
    for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
      int r = int(idx);
      const int step = r / (2 * kEtBlk);
      r -= step * 2 * kEtBlk;
      const int is_lo = r >= kEtBlk;
      if (is_lo) r -= kEtBlk;
      const int k4i = r & 3, n8 = (r >> 2) & 7, kq = (r >> 5) & 1, ng = r >> 6;
      const int nn = ng * 8 + n8, ci = step * 8 + kq * 4 + k4i;
      const int tap = nn / kEtCout, co = nn % kEtCout;
      float v = 0.0f;
      if (ci < cin) v = packed[L.conv_w[l] + ((int64_t(co >> 2) * cin + ci) * 9 + tap) * 4 + (co & 3)];
      float hi, lo;
      tc::split1(v, hi, lo);
      packed[L.tc_conv + base + idx] = is_lo ? lo : hi;
    }
    base += n;
  }
}

#ifdef MAPF_ENC_DEBUG
__device__ long long* g_etc_dbg = nullptr;   // block 0: [layer iteration < 16][warp][6 stamps]
#define ETC_STAMP(i)                                                                                   \
  do {                                                                                                 \
    if (g_etc_dbg && blockIdx.x == 0 && lane == 0 && dbg_it < 16) g_etc_dbg[(dbg_it * 8 + warp) * 6 + (i)] = clock64(); \
  } while (0)
#else
#define ETC_STAMP(i)
#endif

__global__ void __launch_bounds__(kEtThreads, 1) encoder_tc_kernel(EncTcArgs g) {
  extern __shared__ __align__(128) float s_w[];          // B-operand image of all layers
  __shared__ __align__(16) float s_bias[MAPF_MAX_CONV][kEtCout];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, quad = warp & 3;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    tc::mbar_init(&s_bar[0], 1);
    tc::mbar_init(&s_bar[1], 1);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < (g.w_floats >> 2); i += kEtThreads)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(g.wtc) + i);
  if (tid < g.num_conv * kEtCout) s_bias[tid / kEtCout][tid % kEtCout] = __ldg(g.packed + g.bias_off[tid / kEtCout] + tid % kEtCout);
  tc::fence_proxy_async();      // the filters are read by the tensor core (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem + uint32_t(wg) * kEtWgCols;               // this warpgroup's columns, lane 0
  const uint32_t tlane = tmem + (uint32_t(quad * 32) << 16);            // this warp's 32 lanes
  uint64_t* const bar = &s_bar[wg];
  uint32_t phase = 0;

  // lane = pixel (y, x) of this warp's image; mask[t] = 1 when the tap's neighbour pixel exists
  const int py = lane / kFov, px = lane - py * kFov;
  const bool pix_ok = lane < kFovCells;
  float mask[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
    mask[t] = (pix_ok && yy >= 0 && yy < kFov && xx >= 0 && xx < kFov) ? 1.0f : 0.0f;
  }
  const uint32_t vmask = pix_ok ? 0xFFFFFFFFu : 0u;

  constexpr uint32_t idesc = tc::idesc_tf32(128, kEtN);
  const uint32_t w_base = tc::smem_u32(s_w);
  const int64_t ntiles = (g.rows + 3) >> 2;
  const int64_t stride = int64_t(gridDim.x) * 2;

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

#ifdef MAPF_ENC_DEBUG
  int dbg_it = 0;
#endif
  int64_t tile = int64_t(blockIdx.x) * 2 + wg;
  float o0, o1;
  load_obs(tile, o0, o1);
  for (; tile < ntiles; tile += stride) {
    // ---- layer-0 A operand: channels (0, 1) of this pixel, K padded to 8 ----
    {
      float h0, l0, h1, l1;
      tc::split1(o0, h0, l0);
      tc::split1(o1, h1, l1);
      uint32_t hi[8] = {__float_as_uint(h0), __float_as_uint(h1), 0, 0, 0, 0, 0, 0};
      uint32_t lo[8] = {__float_as_uint(l0), __float_as_uint(l1), 0, 0, 0, 0, 0, 0};
      tmem_st8(tlane + kEtColA, hi);
      tmem_st8(tlane + kEtColALo, lo);
    }
    load_obs(tile + stride, o0, o1);     // next tile's pixels arrive while this one is computed
    const int64_t img = tile * 4 + quad;

    for (int l = 0; l < g.num_conv; ++l) {
      ETC_STAMP(0);
      tmem_st_wait();
      tc::fence_before_sync();
      wg_sync();                         // every lane's A rows are in TMEM, every read of D is done
      ETC_STAMP(1);
      if (quad == 0 && lane == 0) {
        tc::fence_after_sync();
        const uint32_t wl = w_base + uint32_t(g.w_off[l]) * 4u;
        for (int ks = 0; ks < g.ksteps[l]; ++ks) {
          const uint64_t bh = tc::smem_desc(wl + uint32_t(ks) * (2 * kEtBlk * 4), 128, 256);
          const uint64_t bl = tc::smem_desc(wl + uint32_t(ks) * (2 * kEtBlk * 4) + kEtBlk * 4, 128, 256);
          const uint32_t ah = tmem + kEtColA + uint32_t(ks) * 8, al = tmem + kEtColALo + uint32_t(ks) * 8;
          mma_tf32_ts(tmem, al, bh, idesc, ks > 0);     // small terms first
          mma_tf32_ts(tmem, ah, bl, idesc, true);
          mma_tf32_ts(tmem, ah, bh, idesc, true);
        }
        tc::mma_commit(bar);
      }
      __syncwarp();
      tc::mbar_wait(bar, phase);
      phase ^= 1u;
      tc::fence_after_sync();
      ETC_STAMP(2);

      // ---- epilogue: out[p][co] = bias[co] + sum_tap P[p + shift(tap)][tap][co] ----
      float2 acc[kEtCout / 2];
      {
        const float4* b4 = reinterpret_cast<const float4*>(s_bias[l]);
#pragma unroll
        for (int q = 0; q < kEtCout / 4; ++q) {
          const float4 b = b4[q];
          acc[2 * q] = make_float2(b.x, b.y);
          acc[2 * q + 1] = make_float2(b.z, b.w);
        }
      }
#pragma unroll
      for (int tg = 0; tg < 3; ++tg) {       // three taps (one filter row) per TMEM round trip
        uint32_t r[3][kEtCout];
#pragma unroll
        for (int u = 0; u < 3; ++u) tmem_ld16(tlane + uint32_t((tg * 3 + u) * kEtCout), r[u]);
        tc::tmem_ld_wait();
        if (tg == 0) ETC_STAMP(3);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int t = tg * 3 + u;
          const int sh = (t / 3 - 1) * kFov + (t % 3 - 1);   // lane offset of the contributing pixel
#pragma unroll
          for (int c = 0; c < kEtCout; c += 2) {
            float a = __uint_as_float(r[u][c]), b = __uint_as_float(r[u][c + 1]);
            if (sh > 0) {
              a = __shfl_down_sync(0xffffffffu, a, sh);
              b = __shfl_down_sync(0xffffffffu, b, sh);
            } else if (sh < 0) {
              a = __shfl_up_sync(0xffffffffu, a, -sh);
              b = __shfl_up_sync(0xffffffffu, b, -sh);
            }
            mapf_ffma2(acc[c >> 1], mask[t], make_float2(a, b));
          }
        }
      }
      ETC_STAMP(4);
      // ReLU; lanes that are not pixels hold zero (they are the zero rows of the next GEMM)
      float hi[kEtCout], lo[kEtCout];
#pragma unroll
      for (int c = 0; c < kEtCout; ++c) {
        const float a = (c & 1) ? acc[c >> 1].y : acc[c >> 1].x;
        const float v = __uint_as_float(__float_as_uint(fmaxf(a, 0.0f)) & vmask);
        tc::split1(v, hi[c], lo[c]);
      }
      if (l + 1 < g.num_conv) {
        uint32_t uh[kEtCout], ul[kEtCout];
#pragma unroll
        for (int c = 0; c < kEtCout; ++c) {
          uh[c] = __float_as_uint(hi[c]);
          ul[c] = __float_as_uint(lo[c]);
        }
        tmem_st16(tlane + kEtColA, uh);
        tmem_st16(tlane + kEtColALo, ul);
      } else if (pix_ok && img < g.rows) {
        // flattened activations, k = pixel*16 + c, as Ap [row/128][k/8][hi|lo][(row%128)/8][(k%8)/4][row%8][k%4]
        float* dst = g.ap + (img >> 7) * int64_t(g.nch) * 2048 + int64_t(2 * lane) * 2048 + ((img & 127) >> 3) * 64 +
                     (img & 7) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float* d = dst + (q >> 1) * 2048 + (q & 1) * 32;
          *reinterpret_cast<float4*>(d) = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          *reinterpret_cast<float4*>(d + 1024) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
      }
      ETC_STAMP(5);
#ifdef MAPF_ENC_DEBUG
      ++dbg_it;
#endif
    }
  }
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
  static_assert(kEtCout == 16 && kEtBlk == 144 * 8, "mapf_packed_layout sizes the tc_conv region with these");
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
    g.ksteps[l] = (p->channels[l] + 7) >> 3;
    g.w_off[l] = off;
    off += g.ksteps[l] * 2 * kEtBlk;
    g.bias_off[l] = L.conv_b[l];
  }
  g.w_floats = off;
  g.rows = a->rows;
  g.states = a->states;
  g.packed = a->packed;
  g.wtc = a->packed + L.tc_conv;
  g.ap = a->x;
  g.nch = (p->fc_in + 7) >> 3;
  const size_t smem = size_t(off) * sizeof(float);
  MAPF_REQUIRE(smem <= 200 * 1024, MAPF_ERR_UNSUPPORTED);
  static int cache[16] = {0};
  if (!mapf_ensure_smem(encoder_tc_kernel, smem, cache)) return MAPF_ERR_CUDA;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (a->rows + 3) >> 2;
  const int64_t want = (ntiles + 1) >> 1;
  const int grid = int(want < sms ? want : sms);
  encoder_tc_kernel<<<grid, kEtThreads, smem, static_cast<cudaStream_t>(stream)>>>(g);
  return mapf_launch_status();
}
