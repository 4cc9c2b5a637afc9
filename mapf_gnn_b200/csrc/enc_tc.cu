// Conv stack of the observation encoder on the 5th-generation tensor cores (reference
// models/framework_gnn.py:55-99,156-163: 3 x {conv3x3 pad 1 + BatchNorm(eval, folded) + ReLU} per agent image [2,5,5]).
//
// A 3x3 convolution over 16 channels is a GEMM with N = 16 per tap -- too thin for tcgen05, whose 128 x 16 x 8 MMA is
// bound by reading its A operand from shared memory.  Here each layer is ONE dense GEMM per 4 agent images,
//   D[pixel, (ky, co)] = sum_{kx, ci} act[pixel + kx - 1, ci] * W[co, ci, ky, kx]        128 x 48 x 48
// i.e. the horizontal taps sit in K (the A operand holds the pixel and its left / right neighbour, zero at the image
// border) and the filter rows sit in N; what is left for the epilogue is the vertical "col2im",
//   out[pixel, co] = bias[co] + D[pixel, (1, co)] + D[pixel above, (0, co)] + D[pixel below, (2, co)].
// The ACTIVATIONS ARE THE A OPERAND IN TENSOR MEMORY (TS mode): lane = pixel and warp = image, so every neighbour
// exchange is a warp shuffle (64 per thread and layer) and a layer's output never touches shared memory -- the epilogue
// threads read the 48 partial sums of their pixel from TMEM (tcgen05.ld), add the neighbours' shares, apply bias and
// ReLU, and write the result and its two horizontal neighbours straight back to TMEM (tcgen05.st) as the next layer's A
// operand.  The B operand (BN-folded filters) stays resident in shared memory (21 KB for three layers).
//
// fp32 accuracy from 16-bit MMAs.  One thread issues a tcgen05.mma only about every 100 cycles however small the MMA is,
// so the number of MMAs per layer is what matters: kind::f16 covers K = 16 per instruction (TF32: 8) and halves the A
// operand's TMEM columns.  Every operand is split into two fp16 numbers, x * 2^-e = hi + lo (11 + 11 significant bits,
// the same as the TF32 split of tc_gemm.cuh) and the layer is 3 products per K step, lo*hi + hi*lo + hi*hi, accumulated
// in fp32 = 9 MMAs.  The power-of-two scales keep fp16's range out of the picture: the filters of a layer share one
// static scale (largest |w| -> [2^9, 2^10), mapf_pack_weights), the activations of an image share a dynamic one
// (largest activation of the image -> [2^9, 2^10), one warp max-reduction per layer); the product of the two inverse
// scales rides on the multiplier of the epilogue's first FMA.  Entries far below their scale's maximum are rounded at
// 2^-25 of it -- 2^-34 of the image's largest activation.
//
// The baseline network's first layer has 32 channels (framework_baseline.py): encoder_tc_kernel<32> runs it as two
// 16-channel GEMM rounds on the same operand, gives the 32 outputs one scale and a 3 x 32-entry operand (144 TMEM
// columns per set, three warpgroups instead of five); layer 1 then contracts over six K steps.
//
// One M tile = 128 TMEM lanes = 4 agent images x 32 lanes (5 rows of 5 pixels, one zero lane between rows, so the
// horizontal exchange needs no border masks).  A CTA holds five independent warpgroups; each walks its own M tiles
// through all layers with its own 96 TMEM columns (D 48 | A hi 24 | A lo 24), its own mbarrier and named barrier; one
// thread per group issues its MMAs in a fixed order (reproducible results), and the groups start a fraction of a layer
// apart (a chain of mbarriers: group g starts when group g - 1 has issued its first MMAs) so that one group's MMA round
// trip falls under the others' epilogues.  The last layer's epilogue emits the
// flattened activations as the fp16-split, pre-tiled A operand of the encoder FC (mapf_tc_fc_presplit_f16: one scale per
// image, its inverse stored behind the operand) with the K index ordered pixel-major (k = pixel * 16 + channel); the
// four images of a tile meet in shared memory so that the global stores are whole 64-byte runs.  The FC's B operand is
// packed with the same K permutation (PackedLayout::tc_fc_pm).  The kernel is launched with programmatic stream
// serialization: everything in front of mapf_pdl_wait() touches only the packed weights.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kEtCout = 16;                     // channels per GEMM: a 32-channel first layer is two of them
constexpr int kEtN = 3 * kEtCout;               // GEMM N: (filter row, co)
constexpr int kEtBlk = kEtN * 16 / 2;           // floats (= pairs of halfs) of one (K step, hi|lo) block of the B operand
constexpr int kEtHdr = 4;                       // floats in front of the B blocks: the layers' inverse filter scales
constexpr uint32_t kEtColAHi = kEtN;            // TMEM columns of a set: D 48 | A hi 24*C1/16 | A lo 24*C1/16
constexpr int kEtOutPitch = 4 * 16 + 4;         // words per pixel in the output staging buffer (4 runs x 64 B + 16 B)
constexpr int kEtOutFloats = kFovCells * kEtOutPitch;
constexpr int kEtMaxC = 32;                     // widest layer (the first one of the baseline network)
// C1 = channels of the first layer (16: the GNN networks, 32: the baseline); every later layer has 16.  The operand of
// layer 1 has 3*C1 K entries, so a set needs 96 or 144 TMEM columns and a CTA holds five or three warpgroups.
template <int C1> struct EtCfg {
  static constexpr int kGroups = C1 == 16 ? 5 : 3;
  static constexpr int kThreads = 128 * kGroups;
  static constexpr uint32_t kAW = 24 * (C1 / 16);
  static constexpr uint32_t kColALo = kEtColAHi + kAW;
  static constexpr uint32_t kSetCols = kEtColAHi + 2 * kAW;
};
static_assert(MAPF_MAX_CONV <= kEtHdr, "one inverse scale per layer");

struct EncTcArgs {
  int num_conv;
  int cin[MAPF_MAX_CONV];
  int w_off[MAPF_MAX_CONV];       // float offset of the layer's B blocks inside the shared-memory weight image
  int64_t bias_off[MAPF_MAX_CONV];// float offset of the layer's folded bias inside `packed`
  int w_floats;                   // floats of the weight image (header + all layers)
  int64_t rows;
  const float* states;
  const float* packed;            // whole packed buffer (biases)
  const float* wtc;               // packed + L.tc_conv
  float* ap;                      // pre-split FC operand out
  float* row_inv;                 // inverse scale of every FC operand row out
  int nch;                        // 16-wide K chunks of the FC (fc_in / 16)
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// wait for the group's MMAs (development variants, tools/enc_tc_variants.py): 0 = suspended try_wait (20 us hint),
// 1 = polling test_wait, 2 = try_wait with a 1 us hint
#ifndef MAPF_ETC_WAIT
#define MAPF_ETC_WAIT 0
#endif
__device__ __forceinline__ void etc_mbar_wait(uint64_t* bar, uint32_t parity) {
#if MAPF_ETC_WAIT == 0
  tc::mbar_wait(bar, parity);
#else
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
#if MAPF_ETC_WAIT == 1
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity), "r"(1000u) : "memory");
#endif
    if (ok) return;
  }
  asm volatile("trap;");
#endif
}

// D[tmem] (+)= A[tmem] * B[smem]^T (TS mode): A = 128 lanes x 8 columns (16 fp16, two per column) at a_tmem
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(uint32_t(accumulate))
      : "memory");
}

// (a, b) -> fp16 pair hi (round to nearest) and fp16 pair lo = (a, b) - hi; a sits in the low half (lower K index)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 back = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// The same split of (a * sc, b * sc) with the two multiplies and the two subtractions as ONE packed instruction each
// (fma.rn.f32x2 with the scalar-broadcast multiplicand: the same IEEE operations, half the issue slots)
__device__ __forceinline__ void split_h2_scaled(float a, float b, float sc, uint32_t& hi, uint32_t& lo) {
  float2 p = make_float2(0.0f, 0.0f);
  mapf_ffma2(p, sc, make_float2(a, b));              // (a, b) * sc  (+ 0: exact)
  const __half2 h = __floats2half2_rn(p.x, p.y);
  mapf_ffma2(p, -1.0f, __half22float2(h));           // (a, b) * sc - hi  (exact: the difference fits fp32)
  const __half2 l = __floats2half2_rn(p.x, p.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// power-of-two scale that brings a maximum magnitude `amax` into [2^9, 2^10), and its inverse
__device__ __forceinline__ void pow2_scale(float amax, float& scale, float& inv) {
  int eb = int(__float_as_uint(amax) >> 23) & 0xFF;
  eb = eb < 32 ? 32 : (eb > 240 ? 240 : eb);
  scale = __uint_as_float(uint32_t(263 - eb) << 23);
  inv = __uint_as_float(uint32_t(eb - 9) << 23);
}

// B operand of the conv GEMMs from the BN-folded filters of pack_weights_kernel ([Cout/4][Cin][9][4]); block l packs
// layer l.  Region: [inverse filter scale of each layer (kEtHdr floats)] then per layer, per 16-wide K step: hi block |
// lo block of fp16 pairs, each [n/8][k/8][n%8][k%8] with n = ky*16 + co, k = kx*Cin + ci (canonical no-swizzle K-major)
__global__ void pack_tc_conv_kernel(mapf_net_params p, PackedLayout L, float* __restrict__ packed) {
  __shared__ float s_red[256];
  const int l = blockIdx.x;
  const int cin = p.channels[l], cout = p.channels[l + 1], nw = cout * cin * 9;
  float m = 0.0f;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) m = fmaxf(m, fabsf(packed[L.conv_w[l] + i]));
  s_red[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + s]);
    __syncthreads();
  }
  float scale, inv;
  pow2_scale(s_red[0], scale, inv);
  if (threadIdx.x == 0) packed[L.tc_conv + l] = inv;
  int64_t base = kEtHdr;
  for (int j = 0; j < l; ++j) base += int64_t(p.channels[j + 1] / kEtCout) * ((3 * p.channels[j] + 15) >> 4) * 2 * kEtBlk;
  const int ks = (3 * cin + 15) >> 4;
  __half* out = reinterpret_cast<__half*>(packed + L.tc_conv + base);
  const int per_half = ks * 2 * kEtBlk * 2;            // halfs of one 16-channel output group: [K step][hi|lo][48 x 16]
  const int n = (cout / kEtCout) * per_half;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int grp = idx / per_half;
    int r = idx - grp * per_half;
    const int step = r / (4 * kEtBlk);
    r -= step * 4 * kEtBlk;
    const int is_lo = r >= 2 * kEtBlk;
    if (is_lo) r -= 2 * kEtBlk;
    const int k8i = r & 7, n8 = (r >> 3) & 7, kq = (r >> 6) & 1, ng = r >> 7;
    const int nn = ng * 8 + n8, k = step * 16 + kq * 8 + k8i;
    const int ky = nn / kEtCout, co = grp * kEtCout + nn % kEtCout;
    float v = 0.0f;
    if (k < 3 * cin) {
      const int kx = k / cin, ci = k - kx * cin;
      v = packed[L.conv_w[l] + ((int64_t(co >> 2) * cin + ci) * 9 + ky * 3 + kx) * 4 + (co & 3)] * scale;
    }
    const __half hi = __float2half_rn(v);
    out[idx] = is_lo ? __float2half_rn(v - __half2float(hi)) : hi;
  }
}

#ifdef MAPF_ENC_DEBUG
__device__ long long* g_etc_dbg = nullptr;   // block 0: [layer iteration < 16][warp < 20][10 stamps]
#define ETC_STAMP(i)                                                                                   \
  do {                                                                                                 \
    if (g_etc_dbg && blockIdx.x == 0 && lane == 0 && dbg_it < 16) g_etc_dbg[(dbg_it * 20 + warp) * 10 + (i)] = clock64(); \
  } while (0)
#else
#define ETC_STAMP(i)
#endif

template <int C1>
__global__ void __launch_bounds__(EtCfg<C1>::kThreads, 1) encoder_tc_kernel(EncTcArgs g) {
  constexpr int kEtGroups = EtCfg<C1>::kGroups, kEtThreads = EtCfg<C1>::kThreads;
  constexpr uint32_t kEtColALo = EtCfg<C1>::kColALo, kEtSetCols = EtCfg<C1>::kSetCols;
  extern __shared__ __align__(128) float s_w[];          // weight image (header + B blocks), then the output staging buffers
  __shared__ __align__(16) float s_bias[MAPF_MAX_CONV][kEtMaxC];
  __shared__ __align__(8) uint64_t s_bar[kEtGroups];     // MMAs of the group's current layer complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t s_go[kEtGroups];      // start-up chain: group g's first MMAs are issued / complete
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_lw[MAPF_MAX_CONV], s_lnm[MAPF_MAX_CONV];   // per layer: byte offset of its B blocks, 16-wide K steps
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, quad = warp & 3;
  float* const s_out = s_w + g.w_floats + wg * kEtOutFloats;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    for (int i = 0; i < kEtGroups; ++i) { tc::mbar_init(&s_bar[i], 1); tc::mbar_init(&s_go[i], 1); }
    tc::fence_mbar_init();
  }
  for (int i = tid; i < (g.w_floats >> 2); i += kEtThreads)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(g.wtc) + i);
  if (tid < g.num_conv) {   // (indexing the kernel-parameter arrays by a runtime layer number would go through local memory)
    s_lw[tid] = uint32_t(kEtHdr + g.w_off[tid]) * 4u;
    s_lnm[tid] = uint32_t((3 * g.cin[tid] + 15) >> 4);
  }
  if (tid < g.num_conv * kEtMaxC) {
    const int l = tid / kEtMaxC, c = tid % kEtMaxC;
    s_bias[l][c] = c < (l == 0 ? C1 : kEtCout) ? __ldg(g.packed + g.bias_off[l] + c) : 0.0f;
  }
  tc::fence_proxy_async();      // the filters are read by the tensor core (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem + uint32_t(wg) * kEtSetCols;             // this warpgroup's columns, lane 0
  const uint32_t tlane = tmem + (uint32_t(quad * 32) << 16);            // this warp's 32 lanes
  uint64_t* const bar = &s_bar[wg];
  uint32_t phase = 0;
  // warp-uniform copies for the MMA-issuing warp (lane 0's values, which ptxas can keep in uniform registers)
  const bool issuer_warp = __shfl_sync(0xffffffffu, quad, 0) == 0;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

  // lanes of a warp = one image, rows six lanes apart: lane 1 + 6y + x holds pixel (y, x); lanes 0, 6, 12, 18, 24, 30, 31 hold
  // zeros, so the left / right neighbour of a border pixel is a zero lane and the horizontal exchange needs no masks
  const int py = (lane + 5) / 6 - 1, px = lane - 1 - 6 * py;             // lane 0 -> py = -1... guarded by pix_ok
  const bool pix_ok = lane >= 1 && lane <= 29 && px >= 0 && px < kFov;
  const int pidx = pix_ok ? py * kFov + px : 0;
  const float m_up = (pix_ok && py > 0) ? 1.0f : 0.0f, m_dn = (pix_ok && py < kFov - 1) ? 1.0f : 0.0f;
  constexpr int kRow = kFov + 1;                                         // lanes between vertically adjacent pixels

  constexpr uint32_t idesc = tc::idesc_f16(128, kEtN);
  const uint32_t w_base_u = __shfl_sync(0xffffffffu, tc::smem_u32(s_w), 0);
  const int64_t ntiles = (g.rows + 3) >> 2;
  const int64_t stride = int64_t(gridDim.x) * kEtGroups;

  auto load_obs = [&](int64_t tile, float& v0, float& v1) {
    const int64_t img = tile * 4 + quad;
    v0 = 0.0f;
    v1 = 0.0f;
    if (tile < ntiles && img < g.rows && pix_ok) {
      const float* s = g.states + img * kFovFloats + pidx;
      v0 = mapf_ld_dep(s);            // written by the env kernel in front: not __ldg (common.cuh)
      v1 = mapf_ld_dep(s + kFovCells);
    }
  };
  auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); };
  // operand word of this pixel's left / right neighbour: the neighbouring lane (a zero lane outside the image; lanes 0
  // and 31, whose shuffle source does not exist, get their own zero back)
  auto left_of = [&](uint32_t w) { return __shfl_up_sync(0xffffffffu, w, 1); };
  auto right_of = [&](uint32_t w) { return __shfl_down_sync(0xffffffffu, w, 1); };
  // maximum over the warp of a non-negative float (its bit pattern orders like an unsigned integer): one REDUX
  auto warp_max = [&](float m) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))); };

#ifdef MAPF_ENC_DEBUG
  int dbg_it = 0;
#endif
  int64_t tile = int64_t(blockIdx.x) * kEtGroups + wg;
  float o0, o1;
  mapf_pdl_wait();      // everything above touched only the packed weights; the observations come from the env kernel
  mapf_pdl_launch();
  load_obs(tile, o0, o1);
  // Start the groups a fraction of a layer apart (one group's MMA wait then falls under the others' epilogues) as a
  // chain of events instead of timed waits: group g starts once group g - 1 has issued its first MMAs -- one operand
  // build + issue (~1 100 clocks) per link, the spacing the tuned clock64() offsets of round 1 had converged to
  // (35.4 us against 35.6 us at 20 480 images; 12.6 against 12.3 us at 5 920).
  if (wg > 0 && tile < ntiles) tc::mbar_wait(&s_go[wg - 1], 0);
  if (tile >= ntiles && (tid & 127) == 0) tc::mbar_arrive(&s_go[wg]);     // a group without work releases its successor
  bool first_issue = true;
  for (; tile < ntiles; tile += stride) {
    float inv_a;     // inverse of the scale of the A operand now in TMEM (one per image)
    // ---- layer-0 A operand: k = kx*2 + c, K padded to 16 ----
    {
      float sc;
      pow2_scale(warp_max(fmaxf(fabsf(o0), fabsf(o1))), sc, inv_a);
      uint32_t hi[8], lo[8];
      split_h2(o0 * sc, o1 * sc, hi[1], lo[1]);
      hi[0] = left_of(hi[1]);  lo[0] = left_of(lo[1]);
      hi[2] = right_of(hi[1]); lo[2] = right_of(lo[1]);
#pragma unroll
      for (int i = 3; i < 8; ++i) hi[i] = lo[i] = 0u;
      tmem_st8(tlane + kEtColAHi, hi);
      tmem_st8(tlane + kEtColALo, lo);
    }
    load_obs(tile + stride, o0, o1);     // next tile's pixels arrive while this one is computed
    const int64_t img0 = tile * 4;

    // MMA rounds of a tile: layer 0 once per group of 16 output channels (twice for the 32-channel first layer of the
    // baseline network: the second round re-uses D after the first round's partial sums have been read), then one per layer
    constexpr int kHalves = C1 / kEtCout;
    float v0[kEtCout];            // C1 == 32: the first 16 channels of layer 0, kept while the second round runs
    float amax0 = 0.0f;
    for (int rd = 0; rd < g.num_conv + kHalves - 1; ++rd) {
      const int l = rd < kHalves ? 0 : rd - (kHalves - 1);
      const int half = rd < kHalves ? rd : 0;
      ETC_STAMP(0);
      tmem_st_wait();
      tc::fence_before_sync();
      wg_sync();                         // every lane's A row is in TMEM, every read of D is done
      ETC_STAMP(1);
      if (issuer_warp) {
        // ONE thread issues the layer's MMAs in a fixed order: the accumulation order -- and with it every bit of the
        // result -- is the same on every run.  (Dealing the MMAs over the group's four warps shortens this phase by
        // a quarter, but the pipe then accumulates in arrival order and results differ in the last bit from run to run.)
        // The whole warp enters this (warp-uniform) branch and elect.sync picks the issuing lane: every operand below is
        // then a uniform-register value and ptxas emits plain UTCHMMAs -- behind a divergent `lane == 0` branch it wraps
        // each MMA in an ELECT / R2UR.BROADCAST loop that costs ~100 clocks per instruction.
        tc::fence_after_sync();
        ETC_STAMP(6);
        const int ksteps = int(s_lnm[l]);
        const uint32_t wl = w_base_u + s_lw[l] + uint32_t(half * ksteps) * (2 * kEtBlk * 4);
        if (tc::elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t bh = tc::smem_desc(wl + uint32_t(ks) * (2 * kEtBlk * 4), 128, 256);
            const uint64_t bl = tc::smem_desc(wl + uint32_t(ks) * (2 * kEtBlk * 4) + kEtBlk * 4, 128, 256);
            const uint32_t ah = tmem_u + kEtColAHi + uint32_t(ks) * 8, al = tmem_u + kEtColALo + uint32_t(ks) * 8;
#ifdef MAPF_ETC_L0_NOLO
            // layer 0: the observations are small integers (exact in fp16), their lo part is identically zero
            if (l > 0) mma_f16_ts(tmem_u, al, bh, idesc, ks > 0);
            mma_f16_ts(tmem_u, ah, bl, idesc, l > 0 || ks > 0);
#else
            mma_f16_ts(tmem_u, al, bh, idesc, ks > 0);     // small terms first
            mma_f16_ts(tmem_u, ah, bl, idesc, true);
#endif
            mma_f16_ts(tmem_u, ah, bh, idesc, true);
          }
          ETC_STAMP(7);
          tc::mma_commit(bar);
          if (first_issue) tc::mbar_arrive(&s_go[wg]);     // the next group may start
          ETC_STAMP(8);
        }
        first_issue = false;
      }
      __syncwarp();
      etc_mbar_wait(bar, phase);
      phase ^= 1u;
      tc::fence_after_sync();
      ETC_STAMP(2);

      // ---- epilogue: out[p][co] = bias[co] + f * (D[p][(1,co)] + D[p-5][(0,co)] + D[p+5][(2,co)]) ----
      uint32_t r[3][kEtCout];
#pragma unroll
      for (int u = 0; u < 3; ++u) tmem_ld16(tlane + uint32_t(u * kEtCout), r[u]);
      tc::tmem_ld_wait();
      ETC_STAMP(3);
      const float f = inv_a * s_w[l];    // undo the activation scale of this image and the filter scale of this layer
      const float f_up = m_up * f, f_dn = m_dn * f;
      float v[kEtCout];
      {
        const float4* b4 = reinterpret_cast<const float4*>(s_bias[l] + half * kEtCout);
#pragma unroll
        for (int q = 0; q < kEtCout / 4; ++q) {
          const float4 b = b4[q];
          v[4 * q] = b.x; v[4 * q + 1] = b.y; v[4 * q + 2] = b.z; v[4 * q + 3] = b.w;
        }
      }
      float amax = 0.0f;
      // two channels per packed FMA (fma.rn.f32x2: the same two IEEE FMAs, one issue slot)
#pragma unroll
      for (int c = 0; c < kEtCout; c += 2) {
        const float2 up = make_float2(__shfl_up_sync(0xffffffffu, __uint_as_float(r[0][c]), kRow),
                                      __shfl_up_sync(0xffffffffu, __uint_as_float(r[0][c + 1]), kRow));
        const float2 dn = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c]), kRow),
                                      __shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c + 1]), kRow));
        float2 a = make_float2(v[c], v[c + 1]);
        mapf_ffma2(a, f, make_float2(__uint_as_float(r[1][c]), __uint_as_float(r[1][c + 1])));
        mapf_ffma2(a, f_up, up);
        mapf_ffma2(a, f_dn, dn);
        v[c] = fmaxf(a.x, 0.0f);                                        // ReLU
        v[c + 1] = fmaxf(a.y, 0.0f);
        amax = fmaxf(amax, fmaxf(v[c], v[c + 1]));
      }
      amax = pix_ok ? amax : 0.0f;   // lanes that are not pixels hold finite garbage: they are zeroed by their scale below
      ETC_STAMP(4);
      if (kHalves == 2 && rd == 0) {
        // first half of the 32-channel layer: keep it in registers, the operand is written when both halves are known
#pragma unroll
        for (int c = 0; c < kEtCout; ++c) v0[c] = v[c];
        amax0 = amax;
      } else if (l + 1 < g.num_conv) {
        // next layer's A operand: k = kx*C + ci -> (left neighbour | this pixel | right neighbour), fp16 hi and lo
        constexpr int kW = kEtCout / 2;                 // operand words (fp16 pairs) per 16 channels
        const bool wide = kHalves == 2 && l == 0;       // this layer has 32 output channels (v0 | v)
        const int nw = wide ? 2 * kW : kW;              // words per copy
        float sc;
        pow2_scale(warp_max(wide ? fmaxf(amax, amax0) : amax), sc, inv_a);
        sc = pix_ok ? sc : 0.0f;     // zero rows for the lanes between the image rows
        uint32_t h[kW], o[kW], t[kW];
#pragma unroll
        for (int part = 0; part < kHalves; ++part) {
          if (part == 1 && !wide) break;
          const int cb = wide ? part * kW : 0;          // first operand word of this group of 16 channels
#pragma unroll
          for (int c = 0; c < kEtCout; c += 2) {
            const float a0 = (wide && part == 0) ? v0[c] : v[c], a1 = (wide && part == 0) ? v0[c + 1] : v[c + 1];
            split_h2_scaled(a0, a1, sc, h[c >> 1], o[c >> 1]);
          }
          tmem_st8(tlane + kEtColAHi + nw + cb, h);
          tmem_st8(tlane + kEtColALo + nw + cb, o);
          // the neighbours' operand words are the neighbours' own splits: shuffle the packed fp16 pairs
#pragma unroll
          for (int j = 0; j < kW; ++j) t[j] = left_of(h[j]);
          tmem_st8(tlane + kEtColAHi + cb, t);
#pragma unroll
          for (int j = 0; j < kW; ++j) t[j] = left_of(o[j]);
          tmem_st8(tlane + kEtColALo + cb, t);
#pragma unroll
          for (int j = 0; j < kW; ++j) t[j] = right_of(h[j]);
          tmem_st8(tlane + kEtColAHi + 2 * nw + cb, t);
#pragma unroll
          for (int j = 0; j < kW; ++j) t[j] = right_of(o[j]);
          tmem_st8(tlane + kEtColALo + 2 * nw + cb, t);
        }
      } else {
        // flattened activations, k = pixel*16 + c, as the fp16-split A operand of the FC: Ap [row/128][k/16 = pixel]
        // [hi|lo][(row%128)/8][(k%16)/8][row%8][k%8] with one power-of-two scale per image (its inverse goes to the row
        // scale array behind the operand).  Staged per pixel as 4 runs (hi|lo, k half) of 4 images x 16 bytes, then
        // stored as 64-byte runs.
        float sc, inv_row;
        pow2_scale(warp_max(amax), sc, inv_row);
        if (pix_ok) {
          uint32_t h[8], o[8];
#pragma unroll
          for (int c = 0; c < kEtCout; c += 2) split_h2_scaled(v[c], v[c + 1], sc, h[c >> 1], o[c >> 1]);
          uint32_t* dst = reinterpret_cast<uint32_t*>(s_out) + pidx * kEtOutPitch + quad * 4;
          *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(dst + 16) = make_uint4(h[4], h[5], h[6], h[7]);
          *reinterpret_cast<uint4*>(dst + 32) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(dst + 48) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        if (lane == 0 && img0 + quad < g.rows) g.row_inv[img0 + quad] = inv_row;
        wg_sync();
        float* const base = g.ap + (img0 >> 7) * int64_t(g.nch) * 2048 + ((img0 & 127) >> 3) * 64 + (img0 & 7) * 4;
        const int wt = tid & 127;
#pragma unroll
        for (int it = 0; it < (kFovCells * 16 + 127) / 128; ++it) {
          const int fi = it * 128 + wt;
          const int p = fi >> 4, run = (fi >> 2) & 3, piece = fi & 3;     // run = (hi|lo, k half)
          if (p < kFovCells && img0 + piece < g.rows) {
            const float4 val = *reinterpret_cast<const float4*>(s_out + p * kEtOutPitch + run * 16 + piece * 4);
            float* d = base + int64_t(p) * 2048 + (run >> 1) * 1024 + (run & 1) * 32 + piece * 4;
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
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(s_tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// ONE 3x3 convolution layer (16 output channels, 2 or 16 input channels) of the TRAINING step on the same tensor-core
// formulation: forward convolutions (pre-BatchNorm output u = conv + bias; the per-slot statistics are a separate pass)
// and data gradients (the same convolution with mirrored, transposed filters) of train.py:87,96.  Activations live in
// global memory as [rows][C*25] (C,H,W order): lane = pixel, so a warp reads / writes 25 contiguous floats per channel.
// Per tile of 4 images: operand into TMEM (fp16 hi / lo split, one power-of-two scale per image) -> 3 (or 9) MMAs ->
// vertical shifted adds -> store.  Five warpgroups per CTA with private TMEM sets, as in encoder_tc_kernel.
// ---------------------------------------------------------------------------------------------------------
struct ConvTcArgs {
  int64_t rows;
  const float* in;     // [rows][CIN*25]
  const float* wtc;    // [kEtHdr floats: inverse filter scale][K steps x (hi block | lo block)] (pack_tc_conv_raw_kernel)
  const float* bias;   // [16] or NULL
  float* out;          // [rows][16*25]
  // per-slot BatchNorm statistics of the output (forward layers; NULL for data gradients): fp64 sums [N][stat_c][2]
  double* stats;
  int slots, stat_c;   // N agent slots (row = b * N + n), channel pitch of `stats`
  // BatchNorm + ReLU of the input on load (bn.stats != NULL; CIN = 16, slots * 16 <= kEtBnTab): see MapfConvBnIn
  MapfConvBnIn bn;
};
constexpr int kEtBnTab = 256;
constexpr float kEtBnEps = 1e-5f;   // nn.BatchNorm2d default (framework_gnn.py:64)

// raw filters [Cout][Cin][3][3] (mode 0) or their data-gradient form W'[o][i][ky][kx] = W[i][o][2-ky][2-kx] (mode 1,
// o = forward input channel) -> B operand blocks for 16 output channels; one block, scale from the largest |w|
__global__ void pack_tc_conv_raw_kernel(const float* __restrict__ w, int cin_fwd, int cout_fwd, int mode,
                                        float* __restrict__ out) {
  __shared__ float s_red[256];
  mapf_pdl_wait();
  mapf_pdl_launch();
  const int nw = cout_fwd * cin_fwd * 9;
  float m = 0.0f;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) m = fmaxf(m, fabsf(mapf_ld_dep(w + i)));
  s_red[threadIdx.x] = m;
  __syncthreads();
  for (int st = blockDim.x / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + st]);
    __syncthreads();
  }
  float scale, inv;
  pow2_scale(s_red[0], scale, inv);
  if (blockIdx.x == 0 && threadIdx.x < kEtHdr) out[threadIdx.x] = threadIdx.x == 0 ? inv : 0.0f;
  const int cin = mode == 0 ? cin_fwd : cout_fwd;      // input channels of THIS convolution (its output has 16)
  const int ks = (3 * cin + 15) >> 4;
  __half* o = reinterpret_cast<__half*>(out + kEtHdr);
  const int n = ks * 2 * kEtBlk * 2;
  // (every block recomputes the scale -- a few loads per thread -- and converts its share of the operand)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    int r = idx;
    const int step = r / (4 * kEtBlk);
    r -= step * 4 * kEtBlk;
    const int is_lo = r >= 2 * kEtBlk;
    if (is_lo) r -= 2 * kEtBlk;
    const int k8i = r & 7, n8 = (r >> 3) & 7, kq = (r >> 6) & 1, ng = r >> 7;
    const int nn = ng * 8 + n8, k = step * 16 + kq * 8 + k8i;
    const int ky = nn / kEtCout, co = nn % kEtCout;
    float v = 0.0f;
    if (k < 3 * cin) {
      const int kx = k / cin, ci = k - kx * cin;
      v = (mode == 0 ? mapf_ld_dep(w + ((co * cin_fwd + ci) * 3 + ky) * 3 + kx)
                     : mapf_ld_dep(w + ((ci * cin_fwd + co) * 3 + (2 - ky)) * 3 + (2 - kx))) * scale;
    }
    const __half hi = __float2half_rn(v);
    o[idx] = is_lo ? __float2half_rn(v - __half2float(hi)) : hi;
  }
}

template <int CIN>
__global__ void __launch_bounds__(EtCfg<16>::kThreads, 1) conv_tc_layer_kernel(ConvTcArgs g) {
  constexpr int kGroups = EtCfg<16>::kGroups, kThreads = EtCfg<16>::kThreads;
  constexpr uint32_t kColALo = EtCfg<16>::kColALo, kSetCols = EtCfg<16>::kSetCols;
  constexpr int kSteps = (3 * CIN + 15) / 16;
  constexpr int kWFloats = kEtHdr + kSteps * 2 * kEtBlk;
  __shared__ __align__(128) float s_w[kWFloats];
  __shared__ __align__(16) float s_b[kEtCout];
  __shared__ __align__(8) uint64_t s_bar[kGroups];
  __shared__ __align__(8) uint64_t s_go[kGroups];      // start-up chain (see encoder_tc_kernel)
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float2 s_tab[CIN == 16 ? kEtBnTab : 1];   // (slot, channel) scale / shift of the input's BatchNorm (g.bn.stats only)
  extern __shared__ __align__(16) double s_stat[];     // [warp][slot][32]: a warp's private per-slot sums (g.stats only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, quad = warp & 3;
  const bool do_stats = g.stats != nullptr;
  double* const my_stat = s_stat + warp * g.slots * 32 + lane;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) { tc::mbar_init(&s_bar[i], 1); tc::mbar_init(&s_go[i], 1); }
    tc::fence_mbar_init();
  }
  if (do_stats)
    for (int i = tid; i < (kThreads / 32) * g.slots * 32; i += kThreads) s_stat[i] = 0.0;
  mapf_pdl_wait();      // filters (packed by the kernel in front), input activations and bias all come from earlier kernels
  mapf_pdl_launch();
  for (int i = tid; i < kWFloats / 4; i += kThreads)
    reinterpret_cast<float4*>(s_w)[i] = mapf_ld_dep(reinterpret_cast<const float4*>(g.wtc) + i);
  if (tid < kEtCout) s_b[tid] = g.bias != nullptr ? mapf_ld_dep(g.bias + tid) : 0.0f;
  const bool bn_in = CIN == 16 && g.bn.stats != nullptr;
  if (CIN == 16 && bn_in) {
    // the arithmetic of bn_finalize_apply4_kernel (train_kernels.cu), so that both forms of the step agree bit for bit
    const int N = g.slots, C = CIN;
    for (int i = tid; i < N * C; i += kThreads) {
      const int n = i / C, c = i - n * C;
      const double s1 = mapf_ld_dep(g.bn.stats + (int64_t(n) * g.bn.stat_c + c) * 2);
      const double s2 = mapf_ld_dep(g.bn.stats + (int64_t(n) * g.bn.stat_c + c) * 2 + 1);
      const double mean = s1 / g.bn.count;
      double var = s2 / g.bn.count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const double invstd = 1.0 / sqrt(var + double(kEtBnEps));
      const float w = mapf_ld_dep(g.bn.gamma + c) * float(invstd);
      const float sh = mapf_ld_dep(g.bn.beta + c) - float(mean) * w;
      s_tab[i] = make_float2(w, sh);
      if (blockIdx.x == 0) {
        g.bn.meaninv[i * 2] = float(mean);
        g.bn.meaninv[i * 2 + 1] = float(invstd);
        g.bn.scale_shift[i * 2] = w;
        g.bn.scale_shift[i * 2 + 1] = sh;
      }
    }
    if (blockIdx.x == 0 && tid < C) {
      const int c = tid;
      float rm = g.bn.running_mean ? g.bn.running_mean[c] : 0.0f, rv = g.bn.running_var ? g.bn.running_var[c] : 0.0f;
      for (int n = 0; n < N; ++n) {
        const double s1 = mapf_ld_dep(g.bn.stats + (int64_t(n) * g.bn.stat_c + c) * 2);
        const double s2 = mapf_ld_dep(g.bn.stats + (int64_t(n) * g.bn.stat_c + c) * 2 + 1);
        const double mean = s1 / g.bn.count;
        double var = s2 / g.bn.count - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float unbiased = float(g.bn.count > 1.0 ? var * g.bn.count / (g.bn.count - 1.0) : var);
        rm = (1.0f - g.bn.momentum) * rm + g.bn.momentum * float(mean);
        rv = (1.0f - g.bn.momentum) * rv + g.bn.momentum * unbiased;
      }
      if (g.bn.running_mean) g.bn.running_mean[c] = rm;
      if (g.bn.running_var) g.bn.running_var[c] = rv;
      if (c == 0 && g.bn.batches != nullptr) *g.bn.batches += N;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem + uint32_t(wg) * kSetCols;
  const uint32_t tlane = tmem + (uint32_t(quad * 32) << 16);
  uint64_t* const bar = &s_bar[wg];
  uint32_t phase = 0;
  const bool issuer_warp = __shfl_sync(0xffffffffu, quad, 0) == 0;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t w_base_u = __shfl_sync(0xffffffffu, tc::smem_u32(s_w), 0);

  const int py = (lane + 5) / 6 - 1, px = lane - 1 - 6 * py;
  const bool pix_ok = lane >= 1 && lane <= 29 && px >= 0 && px < kFov;
  const int pidx = pix_ok ? py * kFov + px : 0;
  const float m_up = (pix_ok && py > 0) ? 1.0f : 0.0f, m_dn = (pix_ok && py < kFov - 1) ? 1.0f : 0.0f;
  constexpr int kRow = kFov + 1;
  constexpr uint32_t idesc = tc::idesc_f16(128, kEtN);
  const int64_t ntiles = (g.rows + 3) >> 2;
  const int64_t stride = int64_t(gridDim.x) * kGroups;

  auto load_in = [&](int64_t tile, float (&x)[CIN]) {
    const int64_t img = tile * 4 + quad;
    const bool live = tile < ntiles && img < g.rows && pix_ok;
    const float* src = g.in + img * (CIN * kFovCells) + pidx;
#pragma unroll
    for (int c = 0; c < CIN; ++c) x[c] = live ? mapf_ld_dep(src + c * kFovCells) : 0.0f;
  };
  auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); };
  auto left_of = [&](uint32_t w) { return __shfl_up_sync(0xffffffffu, w, 1); };
  auto right_of = [&](uint32_t w) { return __shfl_down_sync(0xffffffffu, w, 1); };
  auto warp_max = [&](float m) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))); };

  // tiles are dealt to the CTAs first, then to the groups of a CTA: a small batch (fewer tiles than 5 x SMs) spreads over
  // all SMs with one or two busy groups each instead of filling a fifth of the SMs with five
  int64_t tile = int64_t(wg) * gridDim.x + blockIdx.x;
  float x[CIN];
  load_in(tile, x);
  if (wg > 0 && tile < ntiles) tc::mbar_wait(&s_go[wg - 1], 0);        // start-up chain of the groups (see encoder_tc_kernel)
  if (tile >= ntiles && (tid & 127) == 0) tc::mbar_arrive(&s_go[wg]);
  bool first_issue = true;
  for (; tile < ntiles; tile += stride) {
    float inv_a;
    if (CIN == 16 && bn_in) {
      // x holds the pre-BatchNorm values of this pixel: normalise, ReLU, and leave the activations for the backward pass
      const int64_t im = tile * 4 + quad;
      if (im < g.rows && pix_ok) {
        const float4* tab = reinterpret_cast<const float4*>(s_tab + int(im % g.slots) * CIN);
        float* ao = g.bn.act_out + im * (CIN * kFovCells) + pidx;
#pragma unroll
        for (int c = 0; c < CIN; c += 2) {
          const float4 t = tab[c >> 1];
          x[c] = fmaxf(fmaf(x[c], t.x, t.y), 0.0f);
          x[c + 1] = fmaxf(fmaf(x[c + 1], t.z, t.w), 0.0f);
          ao[c * kFovCells] = x[c];
          ao[(c + 1) * kFovCells] = x[c + 1];
        }
      }
    }
    {
      float amax = 0.0f;
#pragma unroll
      for (int c = 0; c < CIN; ++c) amax = fmaxf(amax, fabsf(x[c]));
      float sc;
      pow2_scale(warp_max(amax), sc, inv_a);
      sc = pix_ok ? sc : 0.0f;
      if (CIN == 2) {
        uint32_t hi[8], lo[8];
        split_h2(x[0] * sc, x[CIN - 1] * sc, hi[1], lo[1]);
        hi[0] = left_of(hi[1]);  lo[0] = left_of(lo[1]);
        hi[2] = right_of(hi[1]); lo[2] = right_of(lo[1]);
#pragma unroll
        for (int i = 3; i < 8; ++i) hi[i] = lo[i] = 0u;
        tmem_st8(tlane + kEtColAHi, hi);
        tmem_st8(tlane + kColALo, lo);
      } else {
        constexpr int kW = 8;
        uint32_t h[kW], o[kW], t[kW];
#pragma unroll
        for (int c = 0; c < 16; c += 2) split_h2_scaled(x[c % CIN], x[(c + 1) % CIN], sc, h[c >> 1], o[c >> 1]);
        tmem_st8(tlane + kEtColAHi + kW, h);
        tmem_st8(tlane + kColALo + kW, o);
#pragma unroll
        for (int j = 0; j < kW; ++j) t[j] = left_of(h[j]);
        tmem_st8(tlane + kEtColAHi, t);
#pragma unroll
        for (int j = 0; j < kW; ++j) t[j] = left_of(o[j]);
        tmem_st8(tlane + kColALo, t);
#pragma unroll
        for (int j = 0; j < kW; ++j) t[j] = right_of(h[j]);
        tmem_st8(tlane + kEtColAHi + 2 * kW, t);
#pragma unroll
        for (int j = 0; j < kW; ++j) t[j] = right_of(o[j]);
        tmem_st8(tlane + kColALo + 2 * kW, t);
      }
    }
    const int64_t img = tile * 4 + quad;
    load_in(tile + stride, x);          // next tile's activations arrive while this one is computed
    tmem_st_wait();
    tc::fence_before_sync();
    wg_sync();
    if (issuer_warp) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kSteps; ++ks) {
          const uint32_t wl = w_base_u + uint32_t(kEtHdr * 4) + uint32_t(ks) * (2 * kEtBlk * 4);
          const uint64_t bh = tc::smem_desc(wl, 128, 256);
          const uint64_t bl = tc::smem_desc(wl + kEtBlk * 4, 128, 256);
          const uint32_t ah = tmem_u + kEtColAHi + uint32_t(ks) * 8, al = tmem_u + kColALo + uint32_t(ks) * 8;
          mma_f16_ts(tmem_u, al, bh, idesc, ks > 0);     // small terms first; fixed order: reproducible
          mma_f16_ts(tmem_u, ah, bl, idesc, true);
          mma_f16_ts(tmem_u, ah, bh, idesc, true);
        }
        tc::mma_commit(bar);
        if (first_issue) tc::mbar_arrive(&s_go[wg]);
      }
      first_issue = false;
    }
    __syncwarp();
    etc_mbar_wait(bar, phase);
    phase ^= 1u;
    tc::fence_after_sync();

    uint32_t r[3][kEtCout];
#pragma unroll
    for (int u = 0; u < 3; ++u) tmem_ld16(tlane + uint32_t(u * kEtCout), r[u]);
    tc::tmem_ld_wait();
    const float f = inv_a * s_w[0];
    const float f_up = m_up * f, f_dn = m_dn * f;
    float* dst = g.out + img * (kEtCout * kFovCells) + pidx;
    const bool live = img < g.rows && pix_ok;
    float sv[kEtCout];            // this pixel's outputs (statistics)
#pragma unroll
    for (int c = 0; c < kEtCout; c += 2) {      // two channels per packed FMA (fma.rn.f32x2)
      const float2 up = make_float2(__shfl_up_sync(0xffffffffu, __uint_as_float(r[0][c]), kRow),
                                    __shfl_up_sync(0xffffffffu, __uint_as_float(r[0][c + 1]), kRow));
      const float2 dn = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c]), kRow),
                                    __shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c + 1]), kRow));
      float2 a = make_float2(s_b[c], s_b[c + 1]);
      mapf_ffma2(a, f, make_float2(__uint_as_float(r[1][c]), __uint_as_float(r[1][c + 1])));
      mapf_ffma2(a, f_up, up);
      mapf_ffma2(a, f_dn, dn);
      if (live) { dst[c * kFovCells] = a.x; dst[(c + 1) * kFovCells] = a.y; }
      sv[c] = live ? a.x : 0.0f;
      sv[c + 1] = live ? a.y : 0.0f;
    }
    if (do_stats) {
      // Per-slot BatchNorm sums (framework_gnn.py:160-163: statistics per agent slot): a warp = one image = one slot.
      // Recursive halving over the lanes (8 + 4 + 2 + 1 shuffles, then one across the two half warps): lanes L and L + 16
      // hold the image's total of channel L.  The squares are summed in fp64 -- the variance is a difference of large
      // numbers (E[x^2] - mean^2) and must not inherit fp32 rounding -- the plain sums in fp32 (25 terms, as ever).
      double sq[kEtCout];
#pragma unroll
      for (int c = 0; c < kEtCout; ++c) sq[c] = double(sv[c]) * double(sv[c]);
#pragma unroll
      for (int h = kEtCout / 2; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = upper ? sv[i] : sv[h + i];
          const float keep = upper ? sv[h + i] : sv[i];
          sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
          const double dsend = upper ? sq[i] : sq[h + i];
          const double dkeep = upper ? sq[h + i] : sq[i];
          sq[i] = dkeep + __shfl_xor_sync(0xffffffffu, dsend, h);
        }
      }
      const float s1 = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 16);
      const double s2 = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 16);
      // lane L < 16 adds the plain sum of channel L, lane L >= 16 the sum of squares of channel L - 16
      if (img < g.rows) my_stat[int(img % g.slots) * 32] += lane < 16 ? double(s1) : s2;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(s_tmem, 512);
  if (do_stats) {
    for (int i = tid; i < g.slots * 32; i += kThreads) {
      double t = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) t += s_stat[w * g.slots * 32 + i];
      const int n = i >> 5, L = i & 31;
      atomicAdd(g.stats + (int64_t(n) * g.stat_c + (L & 15)) * 2 + (L >> 4), t);
    }
  }
}

// per-slot BatchNorm batch statistics of a conv output u [rows][C*25], rows = b * N + n (framework_gnn.py:160-163: the
// reference runs the conv stack once per agent slot, so the statistics are per slot over B*25 values): fp64 sums
// [N][stat_c][2] accumulated with one atomic pair per (CTA, channel).  grid (ceil(B/64), N), one warp per channel.
__global__ void __launch_bounds__(512) bn_stats_kernel(const float* __restrict__ u, int B, int N, int C, int stat_c,
                                                       double* __restrict__ stats) {
  mapf_pdl_wait();
  mapf_pdl_launch();
  const int n = blockIdx.y, b0 = blockIdx.x * 64, lane = threadIdx.x & 31;
  const int nb = B - b0 < 64 ? B - b0 : 64;
  for (int c = threadIdx.x >> 5; c < C; c += blockDim.x >> 5) {
    float s1 = 0.0f;
    double s2 = 0.0;
    if (lane < kFovCells) {
      const float* p = u + (int64_t(b0) * N + n) * (C * kFovCells) + c * kFovCells + lane;
      float q1 = 0.0f, q2 = 0.0f;
      for (int b = 0; b < nb; ++b) {
        const float v = mapf_ld_dep(p + int64_t(b) * N * (C * kFovCells));
        q1 += v;
        q2 = fmaf(v, v, q2);
        if ((b & 15) == 15) { s2 += double(q2); q2 = 0.0f; }
      }
      s1 = q1;
      s2 += double(q2);
    }
    double t1 = double(s1), t2 = s2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) {
      atomicAdd(stats + (int64_t(n) * stat_c + c) * 2, t1);
      atomicAdd(stats + (int64_t(n) * stat_c + c) * 2 + 1, t2);
    }
  }
}

}  // namespace

// One training-step convolution layer on the tensor cores (declared in train_kernels.cu): in [rows][cin*25] -> out
// [rows][16*25] = conv3x3(in, w) + bias, w raw [cout_fwd][cin_fwd][3][3]; mode 1 = data gradient (mirrored, transposed
// filters: the input then has cout_fwd channels and the output cin_fwd = 16).  wtc_scratch: kEtHdr + 3*2*384 floats.
// stats != NULL: per-slot BatchNorm sums of the output are accumulated (B, N = batch and agent slots, rows = B*N).
int mapf_conv_tc_layer_floats() { return kEtHdr + 3 * 2 * kEtBlk; }
bool mapf_conv_tc_layer_ok(int cin_fwd, int cout_fwd, int mode) {
  const int cin = mode == 0 ? cin_fwd : cout_fwd, cout = mode == 0 ? cout_fwd : cin_fwd;
  return cout == kEtCout && (cin == 2 || cin == 16);
}
// filter bank -> the layer kernel's fp16 hi/lo operand blocks (depends on the parameters only: callers may run it on
// another stream than the layer itself)
int mapf_conv_tc_layer_pack(const float* w, int cin_fwd, int cout_fwd, int mode, float* wtc_scratch, mapf_stream_t stream) {
  MAPF_REQUIRE(w && wtc_scratch, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_conv_tc_layer_ok(cin_fwd, cout_fwd, mode), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(wtc_scratch, 16), MAPF_ERR_ALIGN);
  const int cin = mode == 0 ? cin_fwd : cout_fwd;
  mapf_launch_pdl(pack_tc_conv_raw_kernel, dim3(cin == 2 ? 3 : 9), dim3(256), 0, static_cast<cudaStream_t>(stream), w, cin_fwd,
                  cout_fwd, mode, wtc_scratch);
  return mapf_launch_status();
}
// `packed` != 0: wtc_scratch already holds mapf_conv_tc_layer_pack's output for these filters
bool mapf_conv_tc_layer_bn_ok(int cin_fwd, int cout_fwd, int N) {
  return mapf_conv_tc_layer_ok(cin_fwd, cout_fwd, 0) && cin_fwd == 16 && N * 16 <= kEtBnTab;
}
// bn != NULL (mode 0, mapf_conv_tc_layer_bn_ok): `in` is the pre-BatchNorm output of the layer in front, see MapfConvBnIn
int mapf_conv_tc_layer(const float* in, const float* w, const float* bias, int cin_fwd, int cout_fwd, int mode, int B,
                       int N, float* wtc_scratch, int packed, float* out, double* stats, int stat_c, const MapfConvBnIn* bn,
                       mapf_stream_t stream) {
  MAPF_REQUIRE(in && w && wtc_scratch && out, MAPF_ERR_NULL);
  if (bn != nullptr) {
    MAPF_REQUIRE(mode == 0 && mapf_conv_tc_layer_bn_ok(cin_fwd, cout_fwd, N), MAPF_ERR_UNSUPPORTED);
    MAPF_REQUIRE(bn->stats && bn->gamma && bn->beta && bn->meaninv && bn->scale_shift && bn->act_out, MAPF_ERR_NULL);
  }
  MAPF_REQUIRE(mapf_conv_tc_layer_ok(cin_fwd, cout_fwd, mode), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(wtc_scratch, 16), MAPF_ERR_ALIGN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cin = mode == 0 ? cin_fwd : cout_fwd;
  if (!packed) {
    const int prc = mapf_conv_tc_layer_pack(w, cin_fwd, cout_fwd, mode, wtc_scratch, stream);
    if (prc != MAPF_OK) return prc;
  }
  ConvTcArgs g{};
  g.rows = int64_t(B) * N;
  g.in = in; g.wtc = wtc_scratch; g.bias = bias; g.out = out;
  g.slots = N;
  if (bn != nullptr) g.bn = *bn;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (g.rows + 3) >> 2;
  const int grid = int(ntiles < sms ? ntiles : sms);

  // statistics in the conv epilogue while the warps' private tables fit shared memory (N <= 16); else a second pass
  const bool fused_stats = stats != nullptr && N <= 16;
  size_t smem = 0;
  if (fused_stats) {
    g.stats = stats; g.slots = N; g.stat_c = stat_c;
    smem = sizeof(double) * size_t(EtCfg<16>::kThreads / 32) * N * 32;
  }
  auto launch = [&](auto kernel, int* cache) -> int {
    if (smem > 40 * 1024 && !mapf_ensure_smem(kernel, smem, cache)) return MAPF_ERR_CUDA;
    mapf_launch_pdl(kernel, dim3(grid), dim3(EtCfg<16>::kThreads), smem, s, g);
    return MAPF_OK;
  };
  static int c2[16] = {0}, c16[16] = {0};
  const int rc = cin == 2 ? launch(conv_tc_layer_kernel<2>, c2) : launch(conv_tc_layer_kernel<16>, c16);
  if (rc != MAPF_OK) return rc;
  if (stats != nullptr && !fused_stats)
    mapf_launch_pdl(bn_stats_kernel, dim3((B + 63) / 64, N), dim3(512), 0, s, out, B, N, kEtCout, stat_c, stats);
  return mapf_launch_status();
}

#ifdef MAPF_ENC_DEBUG
extern "C" int mapf_enc_tc_set_debug(long long* p) {
  return cudaMemcpyToSymbol(g_etc_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#endif

// first conv layer 16 or 32 channels wide, every later one 16, input 2 planes, FC input 400 = 25 whole K chunks
extern "C" int mapf_encoder_tc_supported(const mapf_net_params* p) {
  static_assert(kEtCout == 16 && kEtBlk == 48 * 8 && kEtHdr == 4 && kEtMaxC == 32,
                "mapf_packed_layout / mapf_enc_tc_dims_ok size the tc_conv region with these");
  if (mapf_check_params(p) != MAPF_OK) return 0;
  return mapf_enc_tc_dims_ok(*p) ? 1 : 0;
}

// called by mapf_pack_weights behind pack_weights_kernel (same stream): needs the BN-folded filters in place
int mapf_encoder_tc_pack(const mapf_net_params* p, float* packed, mapf_stream_t stream) {
  const PackedLayout L = mapf_packed_layout(*p);
  pack_tc_conv_kernel<<<p->num_conv, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p, L, packed);
  const int rc = mapf_tc_pack_b_f16_pixel_major(p->fc_w, p->channels[p->num_conv], p->fc_out, packed + L.tc_fc_pm, stream);
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
    off += (p->channels[l + 1] / kEtCout) * ((3 * p->channels[l] + 15) >> 4) * 2 * kEtBlk;
    g.bias_off[l] = L.conv_b[l];
  }
  off += kEtHdr;
  g.w_floats = off;
  g.rows = a->rows;
  g.states = a->states;
  g.packed = a->packed;
  g.wtc = a->packed + L.tc_conv;
  g.ap = a->x;
  g.nch = (p->fc_in + 15) >> 4;
  g.row_inv = a->x + ((a->rows + 127) >> 7) * int64_t(g.nch) * 2048;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (a->rows + 3) >> 2;
  auto launch = [&](auto cfg, auto kernel, int* cache) -> int {
    constexpr int kGroups = decltype(cfg)::kGroups, kThreads = decltype(cfg)::kThreads;
    const size_t smem = size_t(off + kGroups * kEtOutFloats) * sizeof(float);
    if (smem > 200 * 1024) return MAPF_ERR_UNSUPPORTED;
    if (!mapf_ensure_smem(kernel, smem, cache)) return MAPF_ERR_CUDA;
    const int64_t want = (ntiles + kGroups - 1) / kGroups;
    const int grid = int(want < sms ? want : sms);
    mapf_launch_pdl(kernel, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), g);
    return MAPF_OK;
  };
  static int cache16[16] = {0}, cache32[16] = {0};
  const int rc = p->channels[1] == 16 ? launch(EtCfg<16>{}, encoder_tc_kernel<16>, cache16)
                                      : launch(EtCfg<32>{}, encoder_tc_kernel<32>, cache32);
  if (rc != MAPF_OK) return rc;
  return mapf_launch_status();
}
