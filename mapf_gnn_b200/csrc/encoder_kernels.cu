// Eval-mode observation encoder (reference models/framework_gnn.py:156-166):
// per agent image [2,5,5]: 3 x {conv3x3 pad 1 + BatchNorm(eval, folded) + ReLU} -> flatten (C,H,W) ->
// Linear + ReLU, for all B*N agent images of a batch.
//
// Design (B200): a PERSISTENT kernel, one 256-thread CTA per SM.  Every weight the encoder needs
// (BN-folded conv filters 20-30 KB, transposed FC matrix 100 KB) is copied into shared memory once
// per CTA and stays there while the CTA walks over tiles of 32 agent images, so the steady state
// touches HBM/L2 only for the 200 B image and the 256 B feature row of each agent.  Inside a tile
// lane = agent and warp = a pair of output channels: each thread keeps a 2 x 25 accumulator block
// in registers, reads the 25 input pixels of one input channel with conflict-free LDS (activations
// are stored [channel*25 + pixel][agent]) and the 9 x 2 filter taps with broadcast LDS.64, and the
// 3x3 border handling is resolved at compile time (169 instead of 225 taps per channel pair).
// The FC runs as a 32 x F x fc_in register-tiled GEMM (4 agents x 4 outputs per thread, K split
// over the two halves of the CTA).  Everything is fp32 FFMA (bit-exact argmax requirement).
#include <stdlib.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kEncAgents = 32;
constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;

struct EncoderDims {
  int num_conv;
  int channels[MAPF_MAX_CONV + 1];
  int fc_in, fc_out;
  int w_floats;      // floats of the packed prefix resident in shared memory
  int fc_in_smem;    // 1: the FC matrix is part of that prefix
  int buf0_floats;   // ping buffer 0: image staging, outputs of odd layers, FC split-K scratch
  int buf1_floats;   // ping buffer 1: outputs of even layers
};

#ifdef MAPF_ENC_DEBUG
__device__ long long* g_enc_dbg = nullptr;   // [tile][phase 0..7][warp][arrive, leave] clocks of block 0
#endif

// One 3x3 / stride 1 / zero-pad 1 convolution (BN folded) + ReLU over the tile's 32 images.
// Input element (ci, p) of this lane's image is at in[ci*in_cs + p*in_ps + in_off]; out is
// [Cout*25][32].  w is the packed [Cout/4][Cin][9][4] filter bank in shared memory.
__device__ __forceinline__ void conv3x3_relu_pair(const float* __restrict__ in, int in_cs, int in_ps, int in_off,
                                                  float* __restrict__ out, const float* __restrict__ w,
                                                  const float* __restrict__ bias, int cin, int cout, int warp,
                                                  int lane, int out_ps = kEncAgents) {
  for (int pr = warp; pr < (cout >> 1); pr += kEncWarps) {
    float2 acc[kFovCells];   // (channel 2pr, channel 2pr+1) per pixel: one FFMA2 updates both
    {
      const float2 b = *reinterpret_cast<const float2*>(bias + 2 * pr);
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) acc[p] = b;
    }
    // float2 view: element [(cog*cin + ci)*9 + t]*2 + h with cog = pr/2, h = pr&1
    const float2* wp = reinterpret_cast<const float2*>(w) + (int64_t(pr >> 1) * cin * 9) * 2 + (pr & 1);
#ifdef MAPF_ENC_DEBUG
    const long long t_la = clock64();
#endif
#pragma unroll 2
    for (int ci = 0; ci < cin; ++ci) {
      float v[kFovCells];
      const float* src = in + ci * in_cs + in_off;
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) v[p] = src[p * in_ps];
      float2 wv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wv[t] = wp[(ci * 9 + t) * 2];
      // taps outermost: consecutive FFMA2 hit 25 different accumulator pairs (no back-to-back dependency)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int y = 0; y < kFov; ++y)
#pragma unroll
            for (int x = 0; x < kFov; ++x) {
              const int yy = y + ky - 1, xx = x + kx - 1;
              if (yy >= 0 && yy < kFov && xx >= 0 && xx < kFov)   // resolved at compile time
                mapf_ffma2(acc[y * kFov + x], v[yy * kFov + xx], wv[ky * 3 + kx]);
            }
    }
#ifdef MAPF_ENC_DEBUG
    const long long t_lb = clock64();
#endif
#pragma unroll
    for (int p = 0; p < kFovCells; ++p) {
      out[((2 * pr) * kFovCells + p) * out_ps + lane] = fmaxf(acc[p].x, 0.0f);
      out[((2 * pr + 1) * kFovCells + p) * out_ps + lane] = fmaxf(acc[p].y, 0.0f);
    }
#ifdef MAPF_ENC_DEBUG
    if (g_enc_dbg && blockIdx.x == 0 && lane == 0 && cin > 2) {
      g_enc_dbg[950 + warp * 3] = t_la; g_enc_dbg[951 + warp * 3] = t_lb; g_enc_dbg[952 + warp * 3] = clock64();
    }
#endif
  }
}

// Same layer for a HALF tile of 16 agent images (the ragged last round of the persistent loop, see mapf_encoder_fwd):
// lane = (agent = lane & 15, h = lane >> 4); the two half-warps split the INPUT channels of the warp's output pair,
// exchange their partial sums with one shuffle per pixel, and half h finishes channel 2*pr + h.  Per-warp FFMA2 count is
// half that of a full tile, so a half tile costs about half a tile.  Activations are [channel][25][16] with 16 floats of
// padding from channel `mid` on, which puts the two half-warps on disjoint banks.
__device__ __forceinline__ void conv3x3_relu_half(const float* __restrict__ in, int in_cs, int in_ps, int in_off,
                                                  int in_pad, float* __restrict__ out, int out_ps, int out_mid,
                                                  int out_pad, const float* __restrict__ w,
                                                  const float* __restrict__ bias, int cin, int cout, int warp,
                                                  int lane) {
  const int ag = lane & 15, h = lane >> 4;
  const int cmid = (cin + 1) >> 1;
  for (int pr = warp; pr < (cout >> 1); pr += kEncWarps) {
    float2 acc[kFovCells];
    {
      float2 b = *reinterpret_cast<const float2*>(bias + 2 * pr);
      if (h) b = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) acc[p] = b;
    }
    const float2* wp = reinterpret_cast<const float2*>(w) + (int64_t(pr >> 1) * cin * 9) * 2 + (pr & 1);
#ifdef MAPF_ENC_DEBUG
    const long long t_a = clock64();
#endif
#pragma unroll 2
    for (int i = 0; i < cmid; ++i) {
      const int ci = h * cmid + i;
      const bool live = ci < cin;                       // odd channel counts: the upper half runs one dummy round
      const int cc = live ? ci : cin - 1;
      float v[kFovCells];
      const float* src = in + cc * in_cs + (cc >= cmid ? in_pad : 0) + in_off;
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) v[p] = src[p * in_ps];
      float2 wv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wv[t] = live ? wp[(cc * 9 + t) * 2] : make_float2(0.0f, 0.0f);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int y = 0; y < kFov; ++y)
#pragma unroll
            for (int x = 0; x < kFov; ++x) {
              const int yy = y + ky - 1, xx = x + kx - 1;
              if (yy >= 0 && yy < kFov && xx >= 0 && xx < kFov)
                mapf_ffma2(acc[y * kFov + x], v[yy * kFov + xx], wv[ky * 3 + kx]);
            }
    }
#ifdef MAPF_ENC_DEBUG
    const long long t_b = clock64();
#endif
    const int co = 2 * pr + h;
    float* dst = out + co * kFovCells * out_ps + (co >= out_mid ? out_pad : 0) + ag;
#pragma unroll
    for (int p = 0; p < kFovCells; ++p) {
      const float give = h ? acc[p].x : acc[p].y;       // the partner's channel
      const float got = __shfl_xor_sync(0xffffffffu, give, 16);
      dst[p * out_ps] = fmaxf((h ? acc[p].y : acc[p].x) + got, 0.0f);
    }
#ifdef MAPF_ENC_DEBUG
    if (g_enc_dbg && blockIdx.x == 0 && lane == 0) {
      g_enc_dbg[900 + warp * 3] = t_a; g_enc_dbg[901 + warp * 3] = t_b; g_enc_dbg[902 + warp * 3] = clock64();
    }
#endif
  }
}

// Development hooks (MAPF_B200_NVCC_EXTRA=-DMAPF_ENC_DEBUG, tools/enc_dbg.py): per-phase, per-warp barrier clocks.
#ifdef MAPF_ENC_DEBUG
#define ENC_SYNC(phase)                                                                                    \
  do {                                                                                                     \
    long long t_arr__ = clock64();                                                                         \
    __syncthreads();                                                                                       \
    if (g_enc_dbg && blockIdx.x == 0 && lane == 0 && dbg_tile < 8) {                                       \
      g_enc_dbg[((dbg_tile * 8 + (phase)) * 8 + warp) * 2] = t_arr__;                                      \
      g_enc_dbg[((dbg_tile * 8 + (phase)) * 8 + warp) * 2 + 1] = clock64();                                \
    }                                                                                                      \
  } while (0)
#else
#define ENC_SYNC(phase) __syncthreads()
#endif

// kMode 0: FC with the matrix resident in shared memory; 1: FC with the matrix read through L1 (it did not fit);
// 2: no FC -- the flattened conv output leaves the kernel as the pre-split (TF32 hi | lo), pre-tiled A operand of the
// tensor-core FC (mapf_tc_fc_presplit): x = Ap [row/128][k/8][hi|lo][(row%128)/8][(k%8)/4][row%8][k%4].
constexpr int kSplitPs = 40;   // activation row pitch of the last layer in mode 2 (= 8 mod 32: conflict-free 8x4 reads)

template <int kMode>
__global__ void __launch_bounds__(kEncThreads, 1) encoder_fwd_kernel(EncoderDims d, PackedLayout L, int64_t rows,
                                                                     int64_t nfull, int nhalf,
                                                                     const float* __restrict__ states,
                                                                     const float* __restrict__ packed,
                                                                     float* __restrict__ x) {
  extern __shared__ __align__(16) float s_enc[];
  constexpr bool kFcSmem = kMode == 0, kSplit = kMode == 2;
  constexpr int kLastPs = kSplit ? kSplitPs : kEncAgents;
  float* const s_w = s_enc;                     // packed prefix: conv filters/biases (+ FC matrix and bias)
  float* const s_b0 = s_enc + d.w_floats;
  float* const s_b1 = s_b0 + d.buf0_floats;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // Weights -> shared memory by TMA bulk copies issued by one thread: the conv filters (first barrier) are needed
  // before conv1 of the first item, the FC matrix (second barrier) only at its FC, so most of the 120 KB lands while
  // the first convolutions run.
  __shared__ __align__(8) uint64_t s_wbar[2];
  const uint32_t conv_bytes = uint32_t(L.fc_wt) * 4u, all_bytes = uint32_t(d.w_floats) * 4u;
  if (tid == 0) {
    tc::mbar_init(&s_wbar[0], 1);
    tc::mbar_init(&s_wbar[1], 1);
    tc::fence_mbar_init();
    constexpr uint32_t kChunk = 16 * 1024;
    tc::mbar_expect_tx(&s_wbar[0], conv_bytes);
    for (uint32_t o = 0; o < conv_bytes; o += kChunk)
      tc::bulk_g2s(reinterpret_cast<char*>(s_w) + o, reinterpret_cast<const char*>(packed) + o,
                   conv_bytes - o < kChunk ? conv_bytes - o : kChunk, &s_wbar[0]);
    tc::mbar_arrive(&s_wbar[0]);
    tc::mbar_expect_tx(&s_wbar[1], all_bytes - conv_bytes);
    for (uint32_t o = conv_bytes; o < all_bytes; o += kChunk)
      tc::bulk_g2s(reinterpret_cast<char*>(s_w) + o, reinterpret_cast<const char*>(packed) + o,
                   all_bytes - o < kChunk ? all_bytes - o : kChunk, &s_wbar[1]);
    tc::mbar_arrive(&s_wbar[1]);
  }
  __syncthreads();   // barrier words initialised before anyone polls them

  mapf_pdl_wait();      // the weight copies above touch only the packed buffer; the observations come from the env kernel
  mapf_pdl_launch();
  float* const s_st = s_b1 + d.buf1_floats;     // [32][50] image staging, filled by cp.async one item ahead

  // Work items of this CTA: full tiles b, b + G, ... (< nfull) of 32 agent images, then half tile b (< nhalf) of 16.
  // Item ids: [0, nfull) full tiles, nfull + h half tiles (rows nfull*32 + 16 h ...).
  const int64_t first = blockIdx.x < nfull ? int64_t(blockIdx.x) : (int(blockIdx.x) < nhalf ? nfull + blockIdx.x : -1);
  auto next_item = [&](int64_t it) -> int64_t {
    if (it < nfull) {
      if (it + gridDim.x < nfull) return it + gridDim.x;
      return int(blockIdx.x) < nhalf ? nfull + blockIdx.x : -1;
    }
    return -1;
  };
  auto item_rows = [&](int64_t it, int64_t& row0, int& nrows) {
    const bool full = it < nfull;
    row0 = full ? it * kEncAgents : nfull * kEncAgents + (it - nfull) * (kEncAgents / 2);
    const int cap = full ? kEncAgents : kEncAgents / 2;
    nrows = (rows - row0) < cap ? int(rows - row0) : cap;
  };

  // stage an item's images as they lie in memory ([agent][50]); rows past the end are zero-filled
  auto prefetch = [&](int64_t it) {
    if (it >= 0) {
      int64_t row0;
      int nrows;
      item_rows(it, row0, nrows);
      const int nfl = nrows * kFovFloats;
      const float* src = states + row0 * kFovFloats;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nq = nfl >> 2;
        for (int q = tid; q < nq; q += kEncThreads) {
          const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s_st + 4 * q));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + 4 * q));
        }
        for (int i = (nq << 2) + tid; i < nfl; i += kEncThreads) s_st[i] = mapf_ld_dep(src + i);
      } else {
        for (int i = tid; i < nfl; i += kEncThreads) s_st[i] = mapf_ld_dep(src + i);
      }
      for (int i = nfl + tid; i < kEncAgents * kFovFloats; i += kEncThreads) s_st[i] = 0.0f;
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  prefetch(first);

#ifdef MAPF_ENC_DEBUG
  int dbg_tile = 0;
  if (g_enc_dbg && blockIdx.x == 0 && tid == 0) g_enc_dbg[1023] = clock64();
#endif
  for (int64_t item = first; item >= 0; item = next_item(item)) {
    int64_t row0;
    int nrows;
    item_rows(item, row0, nrows);
    asm volatile("cp.async.wait_group 0;\n" ::);
    ENC_SYNC(0);
    tc::mbar_wait(&s_wbar[0], 0);   // conv filters resident (returns at once after the first item)

    bool in1 = true;  // current activations live in s_b1
    if (item < nfull) {
      // layer 0 reads the staging layout [agent][ci*25 + p]; later layers read [ci*25 + p][agent]
      conv3x3_relu_pair(s_st, kFovCells, 1, lane * kFovFloats, s_b1, s_w + L.conv_w[0], s_w + L.conv_b[0],
                        d.channels[0], d.channels[1], warp, lane, d.num_conv == 1 ? kLastPs : kEncAgents);
      ENC_SYNC(1);
      prefetch(next_item(item));   // the staging buffer is free: the next images land during conv2 / conv3 / FC
      for (int l = 1; l < d.num_conv; ++l) {
        conv3x3_relu_pair(in1 ? s_b1 : s_b0, kFovCells * kEncAgents, kEncAgents, lane, in1 ? s_b0 : s_b1,
                          s_w + L.conv_w[l], s_w + L.conv_b[l], d.channels[l], d.channels[l + 1], warp, lane,
                          l == d.num_conv - 1 ? kLastPs : kEncAgents);
        in1 = !in1;
        ENC_SYNC(1 + l);
      }
    } else {
      // half tile: intermediate activations [channel][25][16] (+16 from the next layer's split channel on); the last
      // layer writes the FC layout [fc_in][32] (columns 16..31 are never stored)
      constexpr int kHalf = kEncAgents / 2;
      const int ag = lane & 15;
      for (int l = 0; l < d.num_conv; ++l) {
        const bool last = l == d.num_conv - 1;
        const float* src = l == 0 ? s_st : (in1 ? s_b1 : s_b0);
        float* dst = l == 0 ? s_b1 : (in1 ? s_b0 : s_b1);
        conv3x3_relu_half(src, l == 0 ? kFovCells : kFovCells * kHalf, l == 0 ? 1 : kHalf,
                          l == 0 ? ag * kFovFloats : ag, l == 0 ? 0 : kHalf, dst, last ? kLastPs : kHalf,
                          last ? (1 << 30) : (d.channels[l + 1] + 1) >> 1, kHalf, s_w + L.conv_w[l], s_w + L.conv_b[l],
                          d.channels[l], d.channels[l + 1], warp, lane);
        if (l > 0) in1 = !in1;
        ENC_SYNC(1 + l);
      }
    }
    if constexpr (kSplit) {
      // flatten (C,H,W) -> TF32 hi / lo halves of the FC's A operand, 8 rows x 4 k core matrices: one warp store =
      // one 128-byte core matrix; element (k, image) sits at act[k * 40 + image], so lane (r8, k4) reads bank 8 k4 + r8
      const float* act = in1 ? s_b1 : s_b0;
      const int nch = (d.fc_in + 7) >> 3;
      float* base = x + (row0 >> 7) * int64_t(nch) * 2048;
      const int rg0 = int(row0 & 127) >> 3;
      // lane = (row group, row in group): it gathers the 4 k of one core-matrix row (4 conflict-free LDS: for a fixed
      // k the warp reads 32 consecutive images) and stores them as one 16-byte piece each of the hi and the lo tile,
      // so a warp instruction writes four whole 128-byte core matrices.  Full tile: one (chunk, k half) per warp
      // step; half tile (two row groups): lanes 16-31 take the other k half of the same chunk.
      const bool full_item = item < nfull;
      const int r8 = lane & 7;
      const int rgl = full_item ? lane >> 3 : (lane >> 3) & 1;
      const int steps = full_item ? nch * 2 : nch;
#pragma unroll 2
      for (int u = warp; u < steps; u += kEncWarps) {
        const int c = full_item ? u >> 1 : u;
        const int kq = full_item ? u & 1 : lane >> 4;
        const int k = c * 8 + kq * 4;
        const float* src = act + k * kSplitPs + rgl * 8 + r8;
        float4 v;
        v.x = k + 0 < d.fc_in ? src[0] : 0.0f;
        v.y = k + 1 < d.fc_in ? src[kSplitPs] : 0.0f;
        v.z = k + 2 < d.fc_in ? src[2 * kSplitPs] : 0.0f;
        v.w = k + 3 < d.fc_in ? src[3 * kSplitPs] : 0.0f;
        float4 hi, lo;
        tc::split4(v, hi, lo);
        float* dst = base + int64_t(c) * 2048 + ((rg0 + rgl) * 2 + kq) * 32 + r8 * 4;
        *reinterpret_cast<float4*>(dst) = hi;
        *reinterpret_cast<float4*>(dst + 1024) = lo;
      }
      ENC_SYNC(4);   // the activation buffers are rewritten by the next item
    } else {
    tc::mbar_wait(&s_wbar[1], 0);           // FC matrix resident
    const float* act = in1 ? s_b1 : s_b0;   // [fc_in][32]
    float* scratch = in1 ? s_b0 : s_b1;     // free buffer: split-K partials [128 threads][16]

    // compressMLP: X[32, F] = relu(act^T Wt + b); thread tile 4 agents x 8 outputs (16 FFMA2 per 3 LDS.128), K split
    // four ways over the CTA, partial sums meet in the free ping buffer
    const int kg = tid >> 6, t6 = tid & 63;
    const int aq = t6 & 7, jq = t6 >> 3;
    const int kq = d.fc_in >> 2;
    const int k0 = kg * kq, k1 = kg == 3 ? d.fc_in : k0 + kq;
    const int wstride = d.fc_out >> 2;
    for (int jb = 0; jb < d.fc_out; jb += 64) {
      const int j = jb + jq * 8;
      const bool jlive = j < d.fc_out;
      float2 acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int h = 0; h < 4; ++h) acc[r][h] = make_float2(0.0f, 0.0f);
      if (jlive) {
        const float4* wps = reinterpret_cast<const float4*>(s_w + L.fc_wt + j);
        const float4* wpg = reinterpret_cast<const float4*>(packed + L.fc_wt + j);
        const float4* ap = reinterpret_cast<const float4*>(act + aq * 4);
#pragma unroll 4
        for (int i = k0; i < k1; ++i) {
          const float4 a4 = ap[i * (kEncAgents / 4)];
          const float4 w0 = kFcSmem ? wps[i * wstride] : __ldg(wpg + i * wstride);
          const float4 w1 = kFcSmem ? wps[i * wstride + 1] : __ldg(wpg + i * wstride + 1);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
          const float2 wp0 = make_float2(w0.x, w0.y), wp1 = make_float2(w0.z, w0.w);
          const float2 wp2 = make_float2(w1.x, w1.y), wp3 = make_float2(w1.z, w1.w);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            mapf_ffma2(acc[r][0], av[r], wp0);
            mapf_ffma2(acc[r][1], av[r], wp1);
            mapf_ffma2(acc[r][2], av[r], wp2);
            mapf_ffma2(acc[r][3], av[r], wp3);
          }
        }
      }
      float4* sc = reinterpret_cast<float4*>(scratch);
      if (kg > 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          sc[((kg - 1) * 8 + r * 2 + 0) * 64 + t6] = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
          sc[((kg - 1) * 8 + r * 2 + 1) * 64 + t6] = make_float4(acc[r][2].x, acc[r][2].y, acc[r][3].x, acc[r][3].y);
        }
      }
      ENC_SYNC(4 + (jb >> 6) * 2);
      if (kg == 0 && jlive) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(packed + L.fc_b + j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(packed + L.fc_b + j) + 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int a = aq * 4 + r;
          float4 lo = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
          float4 hi = make_float4(acc[r][2].x, acc[r][2].y, acc[r][3].x, acc[r][3].y);
#pragma unroll
          for (int g2 = 0; g2 < 3; ++g2) {
            const float4 p0 = sc[(g2 * 8 + r * 2 + 0) * 64 + t6], p1 = sc[(g2 * 8 + r * 2 + 1) * 64 + t6];
            lo.x += p0.x; lo.y += p0.y; lo.z += p0.z; lo.w += p0.w;
            hi.x += p1.x; hi.y += p1.y; hi.z += p1.z; hi.w += p1.w;
          }
          if (a < nrows) {
            float* dst = x + (row0 + a) * d.fc_out + j;
            *reinterpret_cast<float4*>(dst) = make_float4(fmaxf(lo.x + b0.x, 0.0f), fmaxf(lo.y + b0.y, 0.0f),
                                                          fmaxf(lo.z + b0.z, 0.0f), fmaxf(lo.w + b0.w, 0.0f));
            *reinterpret_cast<float4*>(dst + 4) = make_float4(fmaxf(hi.x + b1.x, 0.0f), fmaxf(hi.y + b1.y, 0.0f),
                                                              fmaxf(hi.z + b1.z, 0.0f), fmaxf(hi.w + b1.w, 0.0f));
          }
        }
      }
      ENC_SYNC(5 + (jb >> 6) * 2);  // scratch / staging buffer is rewritten by the next block or tile
    }
    }  // !kSplit
#ifdef MAPF_ENC_DEBUG
    ++dbg_tile;
#endif
  }
}

// Conv stack only (the FC runs on the tensor cores, tc_gemm.cu): NG independent groups of 256 threads per CTA, each
// walking its own tiles of 32 agent images with its own ping-pong buffers and its own named barrier, so that the
// load / barrier phases of one group overlap the FFMA phases of the other (16 warps per SM).
template <int NG>
__global__ void __launch_bounds__(kEncThreads * NG, 1) encoder_conv_kernel(EncoderDims d, PackedLayout L, int64_t rows,
                                                                           const float* __restrict__ states,
                                                                           const float* __restrict__ packed,
                                                                           float* __restrict__ act) {
  extern __shared__ __align__(16) float s_enc[];
  float* const s_w = s_enc;
  const int grp = threadIdx.x / kEncThreads, tid = threadIdx.x % kEncThreads;
  float* const s_b0 = s_enc + d.w_floats + grp * (d.buf0_floats + d.buf1_floats);
  float* const s_b1 = s_b0 + d.buf0_floats;
  const int warp = tid >> 5, lane = tid & 31;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kEncThreads) : "memory"); };

  for (int i = threadIdx.x; i < (d.w_floats >> 2); i += kEncThreads * NG)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(packed) + i);
  __syncthreads();

  const int64_t ntiles = (rows + kEncAgents - 1) / kEncAgents;
  for (int64_t tile = int64_t(blockIdx.x) * NG + grp; tile < ntiles; tile += int64_t(gridDim.x) * NG) {
    const int64_t row0 = tile * kEncAgents;
    const int nrows = (rows - row0) < kEncAgents ? int(rows - row0) : kEncAgents;
    {
      const int nfl = nrows * kFovFloats;
      const float* src = states + row0 * kFovFloats;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nq = nfl >> 2;
        for (int q = tid; q < nq; q += kEncThreads)
          reinterpret_cast<float4*>(s_b0)[q] = mapf_ld_dep(reinterpret_cast<const float4*>(src) + q);
        for (int i = (nq << 2) + tid; i < nfl; i += kEncThreads) s_b0[i] = mapf_ld_dep(src + i);
      } else {
        for (int i = tid; i < nfl; i += kEncThreads) s_b0[i] = mapf_ld_dep(src + i);
      }
      for (int i = nfl + tid; i < kEncAgents * kFovFloats; i += kEncThreads) s_b0[i] = 0.0f;
    }
    group_sync();
    // the last layer writes its [fc_in][32] block with a padded stride so that the row-major copy-out below
    // reads shared memory conflict-free (lanes walk the feature index)
    const bool single = d.num_conv == 1;
    conv3x3_relu_pair(s_b0, kFovCells, 1, lane * kFovFloats, s_b1, s_w + L.conv_w[0], s_w + L.conv_b[0],
                      d.channels[0], d.channels[1], warp, lane, single ? kEncAgents + 1 : kEncAgents);
    group_sync();
    bool in1 = true;
    for (int l = 1; l < d.num_conv; ++l) {
      const bool last = l == d.num_conv - 1;
      conv3x3_relu_pair(in1 ? s_b1 : s_b0, kFovCells * kEncAgents, kEncAgents, lane, in1 ? s_b0 : s_b1,
                        s_w + L.conv_w[l], s_w + L.conv_b[l], d.channels[l], d.channels[l + 1], warp, lane,
                        last ? kEncAgents + 1 : kEncAgents);
      in1 = !in1;
      group_sync();
    }
    {  // flatten (C,H,W) -> act[row][fc_in], coalesced: warp w copies rows w, w+8, ...
      const float* src = in1 ? s_b1 : s_b0;
      for (int r = warp; r < nrows; r += kEncWarps) {
        float* dst = act + (row0 + r) * d.fc_in;
        for (int i = lane; i < d.fc_in; i += 32) dst[i] = src[i * (kEncAgents + 1) + r];
      }
    }
    group_sync();   // the buffers are rewritten by the next tile
  }
}

}  // namespace

#ifdef MAPF_ENC_DEBUG
extern "C" int mapf_enc_set_debug(long long* p) {
  return cudaMemcpyToSymbol(g_enc_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#endif

static int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return 148;
    cached = n;
  }
  return cached;
}

extern "C" int mapf_encoder_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->x, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(a->packed, 16) && mapf_aligned(a->x, 16) && mapf_aligned(a->states, 4), MAPF_ERR_ALIGN);
  MAPF_REQUIRE((p->fc_in & 3) == 0 && (p->fc_out & 7) == 0, MAPF_ERR_UNSUPPORTED);
  const PackedLayout L = mapf_packed_layout(*p);
  EncoderDims d;
  d.num_conv = p->num_conv;
  for (int l = 0; l <= MAPF_MAX_CONV; ++l) d.channels[l] = l <= p->num_conv ? p->channels[l] : 0;
  d.fc_in = p->fc_in;
  d.fc_out = p->fc_out;
  int b0 = kFovFloats, b1 = 0;
  for (int l = 0; l < p->num_conv; ++l) {
    const int out = kFovCells * p->channels[l + 1];
    if (l & 1) b0 = max(b0, out); else b1 = max(b1, out);
  }
  b0 = max(b0, 192);  // FC split-K scratch: 3 partial groups x 64 threads x 32 floats = 6144 floats = 192 x 32
  b1 = max(b1, 192);
  d.buf0_floats = b0 * kEncAgents;
  d.buf1_floats = b1 * kEncAgents;
  const size_t act_bytes = size_t(d.buf0_floats + d.buf1_floats + kEncAgents * kFovFloats) * sizeof(float);
  const size_t limit = 227 * 1024 - 64;   // minus the kernel's static shared memory (two mbarriers)
  d.fc_in_smem = (size_t(L.gf_wt) * sizeof(float) + act_bytes <= limit) ? 1 : 0;
  d.w_floats = int(d.fc_in_smem ? L.gf_wt : L.fc_wt);
  const size_t smem = size_t(d.w_floats) * sizeof(float) + act_bytes;
  MAPF_REQUIRE(smem <= limit, MAPF_ERR_UNSUPPORTED);
  // Persistent grid of G = min(tiles, SMs) CTAs.  When the last round of 32-image tiles would leave at least half of
  // the CTAs idle, that round is cut into 16-image half tiles (one per CTA, about half the time of a full tile).
  const int64_t ntiles = (a->rows + kEncAgents - 1) / kEncAgents;
  const int sms = sm_count();
  int64_t nfull = ntiles;
  int nhalf = 0;
  {
    const int64_t rounds = ntiles / sms, rem = ntiles - rounds * sms;
    if (rem > 0 && 2 * rem <= sms) {
      nfull = rounds * sms;
      nhalf = int((a->rows - nfull * kEncAgents + kEncAgents / 2 - 1) / (kEncAgents / 2));
    }
  }
  const int grid = int(nfull >= sms ? sms : (nfull > nhalf ? nfull : nhalf));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d.fc_in_smem) {
    static int cache_a[16] = {0};
    if (!mapf_ensure_smem(encoder_fwd_kernel<0>, smem, cache_a)) return MAPF_ERR_CUDA;
    mapf_launch_pdl(encoder_fwd_kernel<0>, dim3(grid), dim3(kEncThreads), smem, s, d, L, a->rows, nfull, nhalf, a->states, a->packed, a->x);
  } else {
    static int cache_b[16] = {0};
    if (!mapf_ensure_smem(encoder_fwd_kernel<1>, smem, cache_b)) return MAPF_ERR_CUDA;
    mapf_launch_pdl(encoder_fwd_kernel<1>, dim3(grid), dim3(kEncThreads), smem, s, d, L, a->rows, nfull, nhalf, a->states, a->packed, a->x);
  }
  return mapf_launch_status();
}

// Conv stack of the encoder with the flattened output written as the pre-split A operand of mapf_tc_fc_presplit
// (framework_gnn.py:160-163; the Linear + ReLU of :93 then runs on the tensor cores).  a->x must hold
// mapf_tc_presplit_a_size(rows, fc_in) floats.
extern "C" int mapf_encoder_conv_split_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->x, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(a->packed, 16) && mapf_aligned(a->x, 128) && mapf_aligned(a->states, 4), MAPF_ERR_ALIGN);
  const PackedLayout L = mapf_packed_layout(*p);
  EncoderDims d;
  d.num_conv = p->num_conv;
  for (int l = 0; l <= MAPF_MAX_CONV; ++l) d.channels[l] = l <= p->num_conv ? p->channels[l] : 0;
  d.fc_in = p->fc_in;
  d.fc_out = p->fc_out;
  int b0 = kFovFloats, b1 = 0;
  for (int l = 0; l < p->num_conv; ++l) {
    const int pitch = l == p->num_conv - 1 ? kSplitPs : kEncAgents;
    const int out = (kFovCells * p->channels[l + 1] * pitch + kEncAgents - 1) / kEncAgents;   // in 32-float rows
    if (l & 1) b0 = max(b0, out); else b1 = max(b1, out);
  }
  d.buf0_floats = b0 * kEncAgents;
  d.buf1_floats = b1 * kEncAgents;
  d.fc_in_smem = 0;
  d.w_floats = int(L.fc_wt);
  const size_t smem = size_t(d.w_floats + d.buf0_floats + d.buf1_floats + kEncAgents * kFovFloats) * sizeof(float);
  MAPF_REQUIRE(smem <= 227 * 1024 - 64, MAPF_ERR_UNSUPPORTED);
  const int64_t ntiles = (a->rows + kEncAgents - 1) / kEncAgents;
  const int sms = sm_count();
  int64_t nfull = ntiles;
  int nhalf = 0;
  {
    const int64_t rounds = ntiles / sms, rem = ntiles - rounds * sms;
    if (rem > 0 && 2 * rem <= sms) {
      nfull = rounds * sms;
      nhalf = int((a->rows - nfull * kEncAgents + kEncAgents / 2 - 1) / (kEncAgents / 2));
    }
  }
  const int grid = int(nfull >= sms ? sms : (nfull > nhalf ? nfull : nhalf));
  static int cache_c[16] = {0};
  if (!mapf_ensure_smem(encoder_fwd_kernel<2>, smem, cache_c)) return MAPF_ERR_CUDA;
  mapf_launch_pdl(encoder_fwd_kernel<2>, dim3(grid), dim3(kEncThreads), smem, static_cast<cudaStream_t>(stream), d, L,
                  a->rows, nfull, nhalf, a->states, a->packed, a->x);
  return mapf_launch_status();
}

// Conv stack of the encoder only: act[rows, fc_in] = flatten(relu(bn(conv(...)))) (framework_gnn.py:160-163); the
// Linear + ReLU that follows runs as a tensor-core GEMM (mapf_tc_gemm).
extern "C" int mapf_encoder_conv_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->x, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(a->packed, 16) && mapf_aligned(a->x, 16) && mapf_aligned(a->states, 4), MAPF_ERR_ALIGN);
  const PackedLayout L = mapf_packed_layout(*p);
  EncoderDims d;
  d.num_conv = p->num_conv;
  for (int l = 0; l <= MAPF_MAX_CONV; ++l) d.channels[l] = l <= p->num_conv ? p->channels[l] : 0;
  d.fc_in = p->fc_in;
  d.fc_out = p->fc_out;
  int b0 = kFovFloats * kEncAgents, b1 = 0;
  for (int l = 0; l < p->num_conv; ++l) {
    const int stride = l == p->num_conv - 1 ? kEncAgents + 1 : kEncAgents;   // padded copy-out layout for the last layer
    const int out = kFovCells * p->channels[l + 1] * stride;
    if (l & 1) b0 = max(b0, out); else b1 = max(b1, out);
  }
  d.buf0_floats = (b0 + 3) & ~3;
  d.buf1_floats = (b1 + 3) & ~3;
  d.fc_in_smem = 0;
  d.w_floats = int(L.fc_wt);
  const size_t wbytes = size_t(d.w_floats) * sizeof(float);
  const size_t gbytes = size_t(d.buf0_floats + d.buf1_floats) * sizeof(float);
  const size_t limit = 227 * 1024;
  MAPF_REQUIRE(wbytes + gbytes <= limit, MAPF_ERR_UNSUPPORTED);
  const int ng = 1;   // two independent 256-thread groups per CTA measured slower (worse tile quantisation)
  const size_t smem = wbytes + ng * gbytes;
  const int64_t ntiles = (a->rows + kEncAgents - 1) / kEncAgents;
  const int64_t want = (ntiles + ng - 1) / ng;
  const int grid = int(want < sm_count() ? want : sm_count());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ng == 2) {
    if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(encoder_conv_kernel<2>, size_t(smem), cache__); }())
      return MAPF_ERR_CUDA;
    encoder_conv_kernel<2><<<grid, kEncThreads * 2, smem, s>>>(d, L, a->rows, a->states, a->packed, a->x);
  } else {
    if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(encoder_conv_kernel<1>, size_t(smem), cache__); }())
      return MAPF_ERR_CUDA;
    encoder_conv_kernel<1><<<grid, kEncThreads, smem, s>>>(d, L, a->rows, a->states, a->packed, a->x);
  }
  return mapf_launch_status();
}
