// Imitation-training step on the GPU (reference semantics: train.py:82-100):
// train-mode forward with per-agent-slot BatchNorm statistics, backward through head / graph filter /
// encoder, cross-entropy, global-norm clip + Adam.  New decomposition; nothing here is a port.
//
// Activations live in HBM between layers (grid-wide BatchNorm statistics force a kernel boundary per
// conv layer); rows are r = b*N + n.  Statistics are accumulated in fp64 (block-reduced, then one
// atomicAdd(double) per (slot, channel)), matching the double accumulators of the reference's ATen
// CPU BatchNorm.  Weight-gradient reductions over the batch use split-K GEMMs / persistent CTAs with
// one fp32 atomicAdd per CTA and output element.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr float kBnEpsT = 1e-5f;

// ---------------------------------------------------------------------------------------------
// workspace layout (offsets in floats)
// ---------------------------------------------------------------------------------------------
}  // namespace
// enc_tc.cu: one conv layer of the training step on the tensor cores
int mapf_conv_tc_layer_floats();
bool mapf_conv_tc_layer_ok(int cin_fwd, int cout_fwd, int mode);
int mapf_conv_tc_layer_pack(const float* w, int cin_fwd, int cout_fwd, int mode, float* wtc_scratch, mapf_stream_t stream);
bool mapf_conv_tc_layer_bn_ok(int cin_fwd, int cout_fwd, int N);
int mapf_conv_tc_layer(const float* in, const float* w, const float* bias, int cin_fwd, int cout_fwd, int mode, int B,
                       int N, float* wtc_scratch, int packed, float* out, double* stats, int stat_c, const MapfConvBnIn* bn,
                       mapf_stream_t stream);
namespace {

struct TrainLayout {
  int64_t stat_fwd, stat_bwd;                 // fp64 [L][N][Cmax][2] each (offset in floats, 8-byte aligned)
  int64_t meaninv[MAPF_MAX_CONV];             // [N][C][2]  (mean, invstd)
  int64_t scsh[MAPF_MAX_CONV];                // [N][C][2]  (gamma*invstd, beta - mean*gamma*invstd)
  int64_t coef[MAPF_MAX_CONV];                // [N][C][4]  backward coefficients
  int64_t convpk[MAPF_MAX_CONV];              // [Cout/4][Cin][9][4]
  int64_t convpk_bwd[MAPF_MAX_CONV];          // flipped / transposed filters for the data gradient
  int64_t wt;                                 // [(K(+1))F][G]
  int64_t u[MAPF_MAX_CONV], a[MAPF_MAX_CONV]; // pre-BN conv outputs, post BN+ReLU activations [rows][C*25]
  int64_t x, z, y;                            // [rows][F], [rows][KT], [rows][G]
  int64_t dy, dz, dx, da0, da1, da2;
  int64_t fcpk;                               // compressMLP weight, packed for the tensor-core GEMM (mapf_tc_pack_b)
  int64_t dwpart;                             // [grid <= 296][max cin*cout*9 + cmax] per-CTA weight-gradient partials
  int64_t convtc[MAPF_MAX_CONV], convtc_bwd[MAPF_MAX_CONV];   // tensor-core filter operands (enc_tc.cu mapf_conv_tc_layer)
  int64_t wtt, gfpk_bwd;                      // W_eff^T [KT][G] and its tensor-core packs, one per 128 rows of KT (dz = dy W_eff)
  int64_t fcwt, fcpk_bwd;                     // fc_w^T [fc_in][F] and its packs, one per 128 rows of fc_in (da = dx W_fc)
  int64_t total;
  int cmax;
};

__host__ __device__ inline int64_t r4(int64_t v) { return (v + 3) & ~int64_t(3); }
inline int64_t mapf_min64(int64_t a, int64_t b) { return a < b ? a : b; }

TrainLayout train_layout(const mapf_net_params& p, int B, int N, int message) {
  TrainLayout T{};
  const int64_t rows = int64_t(B) * N;
  int cmax = 0;
  for (int l = 1; l <= p.num_conv; ++l) cmax = max(cmax, p.channels[l]);
  T.cmax = cmax;
  const int KT = (p.taps + (message ? 1 : 0)) * p.fc_out;
  int64_t off = 0;
  auto take = [&](int64_t n) { const int64_t o = off; off += r4(n); return o; };
  T.stat_fwd = take(int64_t(2) * p.num_conv * N * cmax * 2);
  T.stat_bwd = take(int64_t(2) * p.num_conv * N * cmax * 2);
  for (int l = 0; l < p.num_conv; ++l) {
    const int64_t cin = p.channels[l], cout = p.channels[l + 1];
    T.meaninv[l] = take(int64_t(N) * cout * 2);
    T.scsh[l] = take(int64_t(N) * cout * 2);
    T.coef[l] = take(int64_t(N) * cout * 4);
    T.convpk[l] = take(cout * cin * 9);
    T.convpk_bwd[l] = take(cout * r4(cin) * 9);
    T.u[l] = take(rows * cout * 25);
    T.a[l] = take(rows * cout * 25);
  }
  T.wt = take(p.taps > 0 ? (int64_t(KT) * p.node_dim > mapf_tc_b_floats(p.node_dim, KT) ? int64_t(KT) * p.node_dim
                                                                                  : mapf_tc_b_floats(p.node_dim, KT))
                         : 0);
  T.x = take(rows * p.fc_out);
  T.z = take(p.taps > 0 ? rows * KT : 0);
  T.y = take(p.taps > 0 ? rows * p.node_dim : 0);
  T.dy = take(p.taps > 0 ? rows * p.node_dim : 0);
  T.dz = take(p.taps > 0 ? rows * KT : 0);
  T.dx = take(rows * p.fc_out);
  T.da0 = take(rows * cmax * 25);
  T.da1 = take(rows * cmax * 25);
  T.da2 = take(rows * cmax * 25);
  {
    int64_t maxp = 0;
    for (int l = 0; l < p.num_conv; ++l) maxp = max(maxp, int64_t(p.channels[l]) * p.channels[l + 1]);
    const int64_t grid = mapf_min64((rows + 7) / 8, 2 * 148);      // (kCwRowsSmall rows per partial at small batches)
    T.dwpart = take(2 * grid * (maxp * 9 + cmax));     // two sets: consecutive layers' weight gradients run on two lanes
  }
  T.fcpk = take(p.fc_out <= 128 ? mapf_tc_b_floats(p.fc_out, p.fc_in) : 0);
  for (int l = 0; l < p.num_conv; ++l) {
    T.convtc[l] = take(mapf_conv_tc_layer_floats());
    T.convtc_bwd[l] = take(mapf_conv_tc_layer_floats());
  }
  T.wtt = take(p.taps > 0 ? int64_t(KT) * p.node_dim : 0);
  T.gfpk_bwd = take(p.taps > 0 ? int64_t((KT + 127) / 128) * mapf_tc_b_floats(128, p.node_dim) : 0);
  T.fcwt = take(int64_t(p.fc_in) * p.fc_out);
  T.fcpk_bwd = take(int64_t((p.fc_in + 127) / 128) * mapf_tc_b_floats(128, p.fc_out));
  T.total = off;
  return T;
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// filters -> [Cout'/4][Cin'][9][4].  flip = 0: plain regrouping of w [Cout][Cin][9].  flip = 1: the
// data-gradient filter bank, Cout' = Cin (padded to 4), Cin' = Cout, tap mirrored.
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int flip, float* __restrict__ out) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int co2 = flip ? int(r4(cin)) : cout, ci2 = flip ? cout : cin;
  const int total = co2 * ci2 * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int r = idx;
    const int c4 = r & 3; r >>= 2;
    const int t = r % 9; r /= 9;
    const int ci = r % ci2;
    const int co = (r / ci2) * 4 + c4;
    float v;
    if (!flip) v = w[(co * cin + ci) * 9 + t];
    else v = co < cin ? w[(ci * cin + co) * 9 + (8 - t)] : 0.0f;
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// conv3x3 over one agent slot: grid (ceil(B/32), N), lane = batch item, warp = 4 output channels.  The 32 input images
// are staged transposed ([cin*25][33], conflict-free both ways), the packed filters sit in shared memory, and every lane
// stores its 4 finished channels (100 contiguous floats of its row) straight from registers, so a CTA needs
// (cin*25*33 + cout*cin*9) floats and three of them fit an SM.
// ---------------------------------------------------------------------------------------------
// 4-byte asynchronous global -> shared copy: no register staging, so a thread keeps its whole share of a tile in flight
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst_smem))), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

constexpr int kCtThreads = 128;
constexpr int kCtPad = 33;

__global__ void __launch_bounds__(kCtThreads) conv_train_kernel(const float* __restrict__ in, int cin, int cout,
                                                                int B, int N, const float* __restrict__ wpk,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ out, double* __restrict__ stats,
                                                                int stat_c) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  extern __shared__ __align__(16) float s_ct[];
  float* s_w = s_ct;                                 // [cout/4][cin][9][4]
  float* s_in = s_ct + ((cout * cin * 9 + 3) & ~3);  // [cin*25][33]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, b0 = blockIdx.x * 32;
  const int nb = min(32, B - b0);
  const int in_len = cin * kFovCells, out_len = cout * kFovCells;
  for (int i = tid; i < cout * cin * 9; i += kCtThreads) s_w[i] = mapf_ld_dep(wpk + i);
  {
    // staging: every element is its own 4-byte cp.async (coalesced reads, transposing conflict-free writes), all of a
    // thread's ~100 copies in flight at once
    // (warp -> rows, lane -> elements: no index division in the loop)
    for (int r = warp; r < 32; r += kCtThreads / 32) {
      const float* src = in + (int64_t(b0 + r) * N + n) * in_len;
      if (r < nb) {
        for (int i = lane; i < in_len; i += 32) cp_async4(s_in + i * kCtPad + r, src + i);
      } else {
        for (int i = lane; i < in_len; i += 32) s_in[i * kCtPad + r] = 0.0f;
      }
    }
    cp_async_wait_all();
  }
  __syncthreads();
  for (int cog = warp; cog < (cout >> 2); cog += kCtThreads / 32) {
    float2 acc[2][kFovCells];   // channel pairs (4cog, 4cog+1), (4cog+2, 4cog+3): one FFMA2 each per tap
    {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias != nullptr) b = mapf_ld_dep(reinterpret_cast<const float4*>(bias) + cog);
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) { acc[0][p] = make_float2(b.x, b.y); acc[1][p] = make_float2(b.z, b.w); }
    }
    const float4* wp = reinterpret_cast<const float4*>(s_w) + cog * cin * 9;
#pragma unroll 1
    for (int ci = 0; ci < cin; ++ci) {
      float v[kFovCells];
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) v[p] = s_in[(ci * kFovCells + p) * kCtPad + lane];
      float4 wv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wv[t] = wp[ci * 9 + t];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int y = 0; y < kFov; ++y)
#pragma unroll
            for (int x = 0; x < kFov; ++x) {
              const int yy = y + ky - 1, xx = x + kx - 1;
              if (yy >= 0 && yy < kFov && xx >= 0 && xx < kFov) {
                const float a = v[yy * kFov + xx];
                const float4 ww = wv[ky * 3 + kx];
                mapf_ffma2(acc[0][y * kFov + x], a, make_float2(ww.x, ww.y));
                mapf_ffma2(acc[1][y * kFov + x], a, make_float2(ww.z, ww.w));
              }
            }
    }
    // the lane's row holds its 4 channels as 100 contiguous floats (16-byte aligned: row and channel offsets are
    // multiples of 100 floats)
    float4* dst = reinterpret_cast<float4*>(out + (int64_t(b0 + lane) * N + n) * out_len + cog * 4 * kFovCells);
    float flat[4 * kFovCells];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int p = 0; p < kFovCells; ++p) flat[c * kFovCells + p] = (c & 1) ? acc[c >> 1][p].y : acc[c >> 1][p].x;
    if (lane < nb) {
#pragma unroll
      for (int q = 0; q < kFovCells; ++q) dst[q] = make_float4(flat[4 * q], flat[4 * q + 1], flat[4 * q + 2], flat[4 * q + 3]);
    }
    if (stats != nullptr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float s1 = 0.0f;
        double s2 = 0.0;
#pragma unroll
        for (int p = 0; p < kFovCells; ++p) {
          const float val = flat[c * kFovCells + p];
          s1 += val;
          s2 += double(val) * double(val);
        }
        const double t1 = warp_sum(lane < nb ? double(s1) : 0.0);
        const double t2 = warp_sum(lane < nb ? s2 : 0.0);
        if (lane == 0) {
          double* dstat = stats + (int64_t(n) * stat_c + cog * 4 + c) * 2;
          atomicAdd(dstat, t1);
          atomicAdd(dstat + 1, t2);
        }
      }
    }
  }
}

// batch statistics -> (mean, invstd) per (slot, channel); N sequential running-stat updates
// (nn.BatchNorm2d momentum rule with the unbiased variance), framework_gnn.py:160-163
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int N, int C, int stat_c, double count,
                                   float momentum, float* __restrict__ meaninv, float* running_mean,
                                   float* running_var, int64_t* batches, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale_shift) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float rm = running_mean ? running_mean[c] : 0.0f, rv = running_var ? running_var[c] : 0.0f;
    for (int n = 0; n < N; ++n) {
      const double s1 = stats[(int64_t(n) * stat_c + c) * 2], s2 = stats[(int64_t(n) * stat_c + c) * 2 + 1];
      const double mean = s1 / count;
      double var = s2 / count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const double invstd = 1.0 / sqrt(var + double(kBnEpsT));
      meaninv[(n * C + c) * 2] = float(mean);
      meaninv[(n * C + c) * 2 + 1] = float(invstd);
      const float w = gamma[c] * float(invstd);
      scale_shift[(n * C + c) * 2] = w;
      scale_shift[(n * C + c) * 2 + 1] = beta[c] - float(mean) * w;
      const float unbiased = float(count > 1.0 ? var * count / (count - 1.0) : var);
      rm = (1.0f - momentum) * rm + momentum * float(mean);
      rv = (1.0f - momentum) * rv + momentum * unbiased;
    }
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
  }
  if (c == 0 && batches != nullptr) *batches += N;
}

__global__ void bn_apply_kernel(const float* __restrict__ u, const float* __restrict__ scale_shift, int N, int C,
                                int64_t planes, float* __restrict__ a) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  // coalesced, one element per thread and iteration; the (slot, channel) coefficients come from a tiny L1-resident table
  const int64_t total = planes * kFovCells;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t pl = uint32_t(idx / kFovCells);
    const uint32_t c = pl % uint32_t(C), n = (pl / uint32_t(C)) % uint32_t(N);
    const float2 ws = mapf_ld_dep(reinterpret_cast<const float2*>(scale_shift) + (n * C + c));
    a[idx] = fmaxf(fmaf(u[idx], ws.x, ws.y), 0.0f);
  }
}

// 128-bit variant for C*25 % 4 == 0 (a row's C*25 floats are 16-byte aligned, so a float4 never leaves its row)
__global__ void bn_apply4_kernel(const float4* __restrict__ u, const float* __restrict__ scale_shift, int N, int C,
                                 int64_t rows, float4* __restrict__ a) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const uint32_t row4 = uint32_t(C * kFovCells) >> 2;
  const int64_t total = rows * row4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = idx / row4;
    const uint32_t o = uint32_t(idx - row * row4) * 4u, n = uint32_t(row % N);
    const float2* tab = reinterpret_cast<const float2*>(scale_shift) + n * C;
    const float4 v = u[idx];
    const float2 w0 = mapf_ld_dep(tab + o / kFovCells), w1 = mapf_ld_dep(tab + (o + 1) / kFovCells);
    const float2 w2 = mapf_ld_dep(tab + (o + 2) / kFovCells), w3 = mapf_ld_dep(tab + (o + 3) / kFovCells);
    a[idx] = make_float4(fmaxf(fmaf(v.x, w0.x, w0.y), 0.0f), fmaxf(fmaf(v.y, w1.x, w1.y), 0.0f),
                         fmaxf(fmaf(v.z, w2.x, w2.y), 0.0f), fmaxf(fmaf(v.w, w3.x, w3.y), 0.0f));
  }
}

// bn_finalize + bn_apply4 in one launch for small tables (N*C <= kBnFusedMax): every CTA derives the (slot, channel)
// scale / shift pairs from the fp64 sums into shared memory (the same arithmetic as bn_finalize_kernel); CTA 0 also
// stores the tables the backward pass reads and advances the running statistics.  One launch less on the chain of a
// latency-bound small-batch step.
constexpr int kBnFusedMax = 256;
__global__ void __launch_bounds__(256) bn_finalize_apply4_kernel(
    const float4* __restrict__ u, const double* __restrict__ stats, int N, int C, int stat_c, double count, float momentum,
    float* __restrict__ meaninv, float* running_mean, float* running_var, int64_t* batches, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ scale_shift, int64_t rows, float4* __restrict__ a) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  __shared__ float2 s_tab[kBnFusedMax];
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    const int n = i / C, c = i - n * C;
    const double s1 = mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2), s2 = mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2 + 1);
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double invstd = 1.0 / sqrt(var + double(kBnEpsT));
    const float w = gamma[c] * float(invstd);
    s_tab[i] = make_float2(w, beta[c] - float(mean) * w);
    if (blockIdx.x == 0) {
      meaninv[i * 2] = float(mean);
      meaninv[i * 2 + 1] = float(invstd);
      scale_shift[i * 2] = w;
      scale_shift[i * 2 + 1] = beta[c] - float(mean) * w;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < C) {
    const int c = threadIdx.x;
    float rm = running_mean ? running_mean[c] : 0.0f, rv = running_var ? running_var[c] : 0.0f;
    for (int n = 0; n < N; ++n) {
      const double s1 = mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2), s2 = mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2 + 1);
      const double mean = s1 / count;
      double var = s2 / count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const float unbiased = float(count > 1.0 ? var * count / (count - 1.0) : var);
      rm = (1.0f - momentum) * rm + momentum * float(mean);
      rv = (1.0f - momentum) * rv + momentum * unbiased;
    }
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
    if (c == 0 && batches != nullptr) *batches += N;
  }
  __syncthreads();
  const uint32_t row4 = uint32_t(C * kFovCells) >> 2;
  const int64_t total = rows * row4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = idx / row4;
    const uint32_t o = uint32_t(idx - row * row4) * 4u, n = uint32_t(row % N);
    const float2* tab = s_tab + n * C;
    const float4 v = mapf_ld_dep(u + idx);
    const float2 w0 = tab[o / kFovCells], w1 = tab[(o + 1) / kFovCells], w2 = tab[(o + 2) / kFovCells], w3 = tab[(o + 3) / kFovCells];
    a[idx] = make_float4(fmaxf(fmaf(v.x, w0.x, w0.y), 0.0f), fmaxf(fmaf(v.y, w1.x, w1.y), 0.0f),
                         fmaxf(fmaf(v.z, w2.x, w2.y), 0.0f), fmaxf(fmaf(v.w, w3.x, w3.y), 0.0f));
  }
}

// sums over (batch, pixels) per (slot, channel) of dy and dy*xhat, dy = da * [a > 0].  The ReLU mask is recomputed from
// the pre-BatchNorm value with the forward pass's own scale / shift and fmaf (a > 0 <=> fmaf(u, scale, shift) > 0, bit for
// bit what bn_apply stored), so the backward kernels do not read the activations at all.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ scale_shift,
                                                            const float* __restrict__ u,
                                                            const float* __restrict__ meaninv, int B, int N, int C,
                                                            int stat_c, double* __restrict__ stats) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y, b0 = blockIdx.x * 32;
  const int nb = min(32, B - b0);
  for (int c = warp; c < C; c += 8) {
    const float mean = meaninv[(n * C + c) * 2], invstd = meaninv[(n * C + c) * 2 + 1];
    const float sc = scale_shift[(n * C + c) * 2], sh = scale_shift[(n * C + c) * 2 + 1];
    float s1 = 0.0f, s2 = 0.0f;
    if (lane < kFovCells) {
      for (int r = 0; r < nb; ++r) {
        const int64_t idx = ((int64_t(b0 + r) * N + n) * C + c) * kFovCells + lane;
        const float uv = u[idx];
        const float dy = fmaf(uv, sc, sh) > 0.0f ? da[idx] : 0.0f;
        s1 += dy;
        s2 = fmaf(dy, (uv - mean) * invstd, s2);
      }
    }
    const double t1 = warp_sum(double(s1)), t2 = warp_sum(double(s2));
    if (lane == 0) {
      atomicAdd(stats + (int64_t(n) * stat_c + c) * 2, t1);
      atomicAdd(stats + (int64_t(n) * stat_c + c) * 2 + 1, t2);
    }
  }
}

// per (slot, channel): k0 = gamma*invstd, m1 = mean(dy), m2 = mean(dy*xhat)   (fp64 sums -> fp32 coefficients)
__global__ void bn_bwd_coef_kernel(const double* __restrict__ stats, const float* __restrict__ meaninv,
                                   const float* __restrict__ gamma, int N, int C, int stat_c, double count,
                                   float* __restrict__ coef) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  coef[i * 4 + 0] = gamma[c] * meaninv[i * 2 + 1];
  coef[i * 4 + 1] = float(stats[(int64_t(n) * stat_c + c) * 2] / count);
  coef[i * 4 + 2] = float(stats[(int64_t(n) * stat_c + c) * 2 + 1] / count);
  coef[i * 4 + 3] = 0.0f;
}

// du = gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)), in place over da (coalesced, element per thread)
__global__ void bn_bwd_apply_kernel(float* __restrict__ da, const float* __restrict__ scale_shift, const float* __restrict__ u,
                                    const float* __restrict__ meaninv, const float* __restrict__ coef, int N, int C,
                                    int64_t planes) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int64_t total = planes * kFovCells;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t pl = uint32_t(idx / kFovCells);
    const uint32_t c = pl % uint32_t(C), n = (pl / uint32_t(C)) % uint32_t(N);
    const float2 mi = mapf_ld_dep(reinterpret_cast<const float2*>(meaninv) + (n * C + c));
    const float4 k = mapf_ld_dep(reinterpret_cast<const float4*>(coef) + (n * C + c));
    const float2 ss = mapf_ld_dep(reinterpret_cast<const float2*>(scale_shift) + (n * C + c));
    const float uv = u[idx];
    const float dy = fmaf(uv, ss.x, ss.y) > 0.0f ? da[idx] : 0.0f;
    const float xhat = (uv - mi.x) * mi.y;
    da[idx] = k.x * (dy - k.y - xhat * k.z);
  }
}

__global__ void bn_bwd_apply4_kernel(float4* __restrict__ da, const float* __restrict__ scale_shift, const float4* __restrict__ u,
                                     const float* __restrict__ meaninv, const float* __restrict__ coef, int N, int C,
                                     int64_t rows) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const uint32_t row4 = uint32_t(C * kFovCells) >> 2;
  const int64_t total = rows * row4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = idx / row4;
    const uint32_t o = uint32_t(idx - row * row4) * 4u, n = uint32_t(row % N);
    const float4 dv = da[idx], uv = u[idx];
    const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (o + j) / kFovCells;
      const float2 mi = mapf_ld_dep(reinterpret_cast<const float2*>(meaninv) + (n * C + c));
      const float4 k = mapf_ld_dep(reinterpret_cast<const float4*>(coef) + (n * C + c));
      const float2 ss = mapf_ld_dep(reinterpret_cast<const float2*>(scale_shift) + (n * C + c));
      const float dy = fmaf(uu[j], ss.x, ss.y) > 0.0f ? dd[j] : 0.0f;
      const float xhat = (uu[j] - mi.x) * mi.y;
      r[j] = k.x * (dy - k.y - xhat * k.z);
    }
    da[idx] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// bn_bwd_coef + bn_bwd_apply4 in one launch for small tables (N*C <= kBnFusedMax): the coefficients go to shared memory
__global__ void __launch_bounds__(256) bn_bwd_coef_apply4_kernel(
    float4* __restrict__ da, const float* __restrict__ scale_shift, const float4* __restrict__ u, const float* __restrict__ meaninv,
    const double* __restrict__ stats, const float* __restrict__ gamma, int N, int C, int stat_c, double count, int64_t rows,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  if (blockIdx.x == 0 && threadIdx.x < C) {
    // BatchNorm parameter gradients (bn_param_grads_kernel's arithmetic): d gamma = sum dy * xhat, d beta = sum dy over all slots
    const int c = threadIdx.x;
    double g = 0.0, b = 0.0;
    for (int n = 0; n < N; ++n) {
      b += mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2);
      g += mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2 + 1);
    }
    dgamma[c] = float(g);
    dbeta[c] = float(b);
  }
  __shared__ float4 s_k[kBnFusedMax];
  __shared__ float2 s_mi[kBnFusedMax], s_ss[kBnFusedMax];
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    const int n = i / C, c = i - n * C;
    const float2 mi = mapf_ld_dep(reinterpret_cast<const float2*>(meaninv) + i);
    s_mi[i] = mi;
    s_ss[i] = mapf_ld_dep(reinterpret_cast<const float2*>(scale_shift) + i);
    s_k[i] = make_float4(gamma[c] * mi.y, float(mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2) / count),
                         float(mapf_ld_dep(stats + (int64_t(n) * stat_c + c) * 2 + 1) / count), 0.0f);
  }
  __syncthreads();
  const uint32_t row4 = uint32_t(C * kFovCells) >> 2;
  const int64_t total = rows * row4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = idx / row4;
    const uint32_t o = uint32_t(idx - row * row4) * 4u, n = uint32_t(row % N);
    const float4 dv = mapf_ld_dep(da + idx), uv = u[idx];
    const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (o + j) / kFovCells;
      const float2 mi = s_mi[n * C + c], ss = s_ss[n * C + c];
      const float4 k = s_k[n * C + c];
      const float dy = fmaf(uu[j], ss.x, ss.y) > 0.0f ? dd[j] : 0.0f;
      const float xhat = (uu[j] - mi.x) * mi.y;
      r[j] = k.x * (dy - k.y - xhat * k.z);
    }
    da[idx] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// bn_bwd_reduce for C*25 % 4 == 0: one CTA per (32 batch rows, slot); thread t < 2*row4 owns the float4 column
// t % row4 of the rows of its parity, so every warp reads whole 512-byte pieces of a row; a float4 touches at most two
// channels, whose partial sums meet in shared memory before one fp64 atomic per (CTA, channel).
__global__ void __launch_bounds__(256) bn_bwd_reduce4_kernel(const float4* __restrict__ da, const float* __restrict__ scale_shift,
                                                             const float4* __restrict__ u,
                                                             const float* __restrict__ meaninv, int B, int N, int C,
                                                             int stat_c, double* __restrict__ stats, int rows_per_cta) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  __shared__ float s_sum[64 * 2];
  const int n = blockIdx.y, b0 = blockIdx.x * rows_per_cta;
  const int nb = min(rows_per_cta, B - b0);
  const int row4 = (C * kFovCells) >> 2;
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) s_sum[i] = 0.0f;
  __syncthreads();
  const int groups = blockDim.x / row4;                 // row-parity groups that fit the CTA (2 for C = 16)
  const int grp = threadIdx.x / row4, f4 = threadIdx.x - grp * row4;
  if (grp < groups) {
    const int o = f4 * 4;
    int ch[4];
    float mean[4], inv[4], sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ch[j] = (o + j) / kFovCells;
      mean[j] = meaninv[(n * C + ch[j]) * 2];
      inv[j] = meaninv[(n * C + ch[j]) * 2 + 1];
      sc[j] = scale_shift[(n * C + ch[j]) * 2];
      sh[j] = scale_shift[(n * C + ch[j]) * 2 + 1];
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int r = grp; r < nb; r += groups) {
      const int64_t idx = (int64_t(b0 + r) * N + n) * row4 + f4;
      const float4 dv = da[idx], uv = u[idx];
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dy = fmaf(uu[j], sc[j], sh[j]) > 0.0f ? dd[j] : 0.0f;
        s1[j] += dy;
        s2[j] = fmaf(dy, (uu[j] - mean[j]) * inv[j], s2[j]);
      }
    }
    // elements of the same channel first (a float4 spans at most two channels)
    float c0s1 = s1[0], c0s2 = s2[0], c1s1 = 0.f, c1s2 = 0.f;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      if (ch[j] == ch[0]) { c0s1 += s1[j]; c0s2 += s2[j]; } else { c1s1 += s1[j]; c1s2 += s2[j]; }
    }
    atomicAdd(&s_sum[ch[0] * 2], c0s1);
    atomicAdd(&s_sum[ch[0] * 2 + 1], c0s2);
    if (ch[3] != ch[0]) {
      atomicAdd(&s_sum[ch[3] * 2], c1s1);
      atomicAdd(&s_sum[ch[3] * 2 + 1], c1s2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x)
    atomicAdd(stats + (int64_t(n) * stat_c + (i >> 1)) * 2 + (i & 1), double(s_sum[i]));
}

__global__ void bn_param_grads_kernel(const double* __restrict__ stats, int N, int C, int stat_c,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, b = 0.0;
  for (int n = 0; n < N; ++n) {
    b += stats[(int64_t(n) * stat_c + c) * 2];
    g += stats[(int64_t(n) * stat_c + c) * 2 + 1];
  }
  dgamma[c] = float(g);
  dbeta[c] = float(b);
}

// dW[co][ci][tap] = sum_{rows, pixels} du[r][co][p] * in[r][ci][p + tap], db[co] = sum du
// Persistent over tiles of 32 rows; thread = (co, ci) pair with its 9 tap sums in registers.  Rows are staged in PAIRS
// ([16][len][2]) so that one LDS.64 fetches the same element of two rows and one packed FFMA2 advances the even-row
// and the odd-row partial sum of a tap together: half the shared-memory wavefronts and half the FMA issue slots of the
// scalar form.
constexpr int kCwThreads = 256;
constexpr int kCwRows = 32;       // rows per tile when the batch fills the SMs with such tiles
constexpr int kCwRowsSmall = 8;   // ... and when it does not (a CTA walks its row pairs serially: 15 us per 32-row tile)

template <int SLOTS>
__global__ void __launch_bounds__(kCwThreads, SLOTS == 1 ? 2 : 1)
conv_bwd_weight_kernel(const float* __restrict__ du, const float* __restrict__ in, int cin, int cout, int64_t rows,
                       float* __restrict__ part, int tile_rows) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  extern __shared__ __align__(16) float s_cw[];
  const int in_len = cin * kFovCells, out_len = cout * kFovCells;
  float2* s_du = reinterpret_cast<float2*>(s_cw);                          // [16][out_len] (even row, odd row)
  float2* s_in = reinterpret_cast<float2*>(s_cw) + (tile_rows / 2) * out_len;   // [tile_rows / 2][in_len]
  const int tid = threadIdx.x;
  const int npairs = cin * cout;
  const bool small = npairs < kCwThreads;
  const int rstep = small ? kCwThreads / npairs : 1;
  const int rstart = small ? tid / npairs : 0;
  const bool active = small ? tid < npairs * rstep : true;
  float2 acc[SLOTS][9];
  float2 accb[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    accb[s] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[s][t] = make_float2(0.f, 0.f);
  }
  auto stage = [&](const float* __restrict__ src, int len, int64_t r0, int nr, float2* dst) {
    // element k of row 2 rp + h -> dst[rp][k].{x,y}: 4-byte cp.async each (coalesced reads), all in flight at once
    // (warp -> rows, lane -> elements: no index division in the loop)
    float* d1 = reinterpret_cast<float*>(dst);
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < tile_rows; r += kCwThreads / 32) {
      float* drow = d1 + int64_t(r >> 1) * len * 2 + (r & 1);
      const float* srow = src + (r0 + r) * len;
      if (r < nr) {
        for (int k = lane; k < len; k += 32) cp_async4(drow + 2 * k, srow + k);
      } else {
        for (int k = lane; k < len; k += 32) drow[2 * k] = 0.0f;
      }
    }
  };
  const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * tile_rows;
    const int nr = (rows - r0) < tile_rows ? int(rows - r0) : tile_rows;
    __syncthreads();
    stage(du, out_len, r0, nr, s_du);
    stage(in, in_len, r0, nr, s_in);
    cp_async_wait_all();
    __syncthreads();
    if (!active) continue;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int pp = small ? tid % npairs : tid + s * kCwThreads;
      if (pp >= npairs || (small && s > 0)) break;
      const int co = pp / cin, ci = pp - co * cin;
      for (int rp = rstart; rp < tile_rows / 2; rp += rstep) {
        float2 d[kFovCells], v[kFovCells];
#pragma unroll
        for (int p = 0; p < kFovCells; ++p) {
          d[p] = s_du[rp * out_len + co * kFovCells + p];
          v[p] = s_in[rp * in_len + ci * kFovCells + p];
        }
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
                  mapf_ffma2_pp(acc[s][ky * 3 + kx], d[y * kFov + x], v[yy * kFov + xx]);
              }
        if (ci == 0) {
#pragma unroll
          for (int p = 0; p < kFovCells; ++p) { accb[s].x += d[p].x; accb[s].y += d[p].y; }
        }
      }
    }
  }
  // Per-CTA partial sums go to part[blockIdx.x][npairs*9 + cout] (fixed order inside the CTA); a second kernel adds the
  // CTAs in index order: no atomics (2 300 same-address atomics per filter tap cost more than the convolution itself
  // when cin*cout is small) and a reproducible gradient.
  float* mine = part + int64_t(blockIdx.x) * (npairs * 9 + cout);
  if (small) {
    float* s_red = s_cw;                       // [rstep][npairs][10]
    __syncthreads();
    if (active) {
      float* dst = s_red + (int64_t(rstart) * npairs + tid % npairs) * 10;
#pragma unroll
      for (int t = 0; t < 9; ++t) dst[t] = acc[0][t].x + acc[0][t].y;
      dst[9] = accb[0].x + accb[0].y;
    }
    __syncthreads();
    if (tid < npairs) {
      float sum[10];
#pragma unroll
      for (int t = 0; t < 10; ++t) sum[t] = 0.0f;
      for (int g = 0; g < rstep; ++g)
#pragma unroll
        for (int t = 0; t < 10; ++t) sum[t] += s_red[(int64_t(g) * npairs + tid) * 10 + t];
      const int co = tid / cin, ci = tid - co * cin;
#pragma unroll
      for (int t = 0; t < 9; ++t) mine[tid * 9 + t] = sum[t];
      if (ci == 0) mine[npairs * 9 + co] = sum[9];
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int pp = tid + s * kCwThreads;
    if (pp >= npairs) break;
    const int co = pp / cin, ci = pp - co * cin;
#pragma unroll
    for (int t = 0; t < 9; ++t) mine[pp * 9 + t] = acc[s][t].x + acc[s][t].y;
    if (ci == 0) mine[npairs * 9 + co] = accb[s].x + accb[s].y;
  }
}

// dw[e] += sum over CTAs of part[g][e] in a fixed order (the last `cout` entries are the bias gradient): a CTA owns 32
// outputs, its 8 warps take the partials g = w, w + 8, ... (coalesced 128-byte rows), and the 8 warp sums meet in
// shared memory in warp order.
__global__ void __launch_bounds__(256) conv_bwd_weight_reduce_kernel(const float* __restrict__ part, int grid, int nw,
                                                                     int cout, float* __restrict__ dw,
                                                                     float* __restrict__ db) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  __shared__ float s_sum[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane, tot = nw + cout;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  if (e < tot) {
    int g = warp;
    for (; g + 24 < grid; g += 32) {
      a0 += mapf_ld_dep(part + int64_t(g) * tot + e);
      a1 += mapf_ld_dep(part + int64_t(g + 8) * tot + e);
      a2 += mapf_ld_dep(part + int64_t(g + 16) * tot + e);
      a3 += mapf_ld_dep(part + int64_t(g + 24) * tot + e);
    }
    for (; g < grid; g += 8) a0 += mapf_ld_dep(part + int64_t(g) * tot + e);
  }
  s_sum[warp][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (warp == 0 && e < tot) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_sum[w][lane];
    if (e < nw) dw[e] += sum;
    else if (db != nullptr) db[e - nw] += sum;
  }
}

// ---------------------------------------------------------------------------------------------
// generic fp32 GEMM  C[M,N] (+)= A[M,K] B[K,N]  with arbitrary element strides
//   epilogue: + bias[n], ReLU, * [mask > 0]; split-K accumulates with atomicAdd into zeroed C
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  int M, N, K;
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  const float* bias;
  const float* mask; int64_t ldmask;
  int relu;
  int splits;
};
constexpr int kGemmTM = 128, kGemmTN = 64, kGemmK = 16, kGemmThreads = 256;

// 128 x 64 tile, 8 x 4 outputs per thread kept as FFMA2 pairs along n, global -> register prefetch of the next K slab
// while the current one is multiplied out of shared memory.
__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(GemmArgs g) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  __shared__ __align__(16) float As[kGemmK][kGemmTM + 4];
  __shared__ __align__(16) float Bs[kGemmK][kGemmTN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kGemmTM, n0 = blockIdx.x * kGemmTN;
  const int kper = ((g.K + g.splits - 1) / g.splits + kGemmK - 1) / kGemmK * kGemmK;
  const int kbeg = blockIdx.z * kper, kend = min(g.K, kbeg + kper);
  float2 acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);
  const bool a_kfast = g.sak == 1, b_nfast = g.sbn == 1;
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * kGemmThreads;
      int m, k;
      if (a_kfast) { m = e >> 4; k = e & 15; } else { m = e & 127; k = e >> 7; }
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < kend) ? mapf_ld_dep(g.A + gm * g.sam + gk * g.sak) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * kGemmThreads;
      int n, kb;
      if (b_nfast) { n = e & 63; kb = e >> 6; } else { n = e >> 4; kb = e & 15; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[i] = (gn < g.N && gkb < kend) ? mapf_ld_dep(g.B + gkb * g.sbk + gn * g.sbn) : 0.0f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * kGemmThreads;
      if (a_kfast) As[e & 15][e >> 4] = ra[i]; else As[e >> 7][e & 127] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * kGemmThreads;
      if (b_nfast) Bs[e >> 6][e & 63] = rb[i]; else Bs[e & 15][e >> 4] = rb[i];
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += kGemmK) {
    stash();
    __syncthreads();
    if (k0 + kGemmK < kend) fetch(k0 + kGemmK);
#pragma unroll
    for (int k = 0; k < kGemmK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 blo = make_float2(b4.x, b4.y), bhi = make_float2(b4.z, b4.w);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mapf_ffma2(acc[i][0], av[i], blo);
        mapf_ffma2(acc[i][1], av[i], bhi);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    const float vals[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = vals[j];
      if (g.splits > 1) {
        atomicAdd(g.C + m * g.ldc + n, v);
      } else {
        if (g.bias) v += mapf_ld_dep(g.bias + n);
        if (g.relu) v = fmaxf(v, 0.0f);
        if (g.mask) v = mapf_ld_dep(g.mask + m * g.ldmask + n) > 0.0f ? v : 0.0f;
        g.C[m * g.ldc + n] = v;
      }
    }
  }
}

int gemm(cudaStream_t s, int M, int N, int K, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
         int64_t sbn, float* C, int64_t ldc, const float* bias = nullptr, int relu = 0, const float* mask = nullptr,
         int64_t ldmask = 0, int splits = 1) {
  GemmArgs g{M, N, K, A, sam, sak, B, sbk, sbn, C, ldc, bias, mask, ldmask, relu, splits};
  dim3 grid((N + kGemmTN - 1) / kGemmTN, (M + kGemmTM - 1) / kGemmTM, splits);
  mapf_launch_pdl(sgemm_kernel, dim3(grid), dim3(kGemmThreads), 0, s, g);
  return mapf_launch_status();
}

// number of K splits for a reduction over `rows` with a small output
int splits_for(int64_t rows, int tiles) {
  int s = int((rows + 127) / 128);               // >= 128 rows of the reduction per split
  const int cap = max(1, 592 / max(1, tiles));   // ~4 CTAs per SM in total
  return max(1, min(s, cap));
}

// out[c] = sum_r in[r][c]   (C <= 256)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, int64_t rows, int C,
                                                     float* __restrict__ out) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  __shared__ float s_part[256];
  const int tid = threadIdx.x;
  const int per = 256 / C;                  // row lanes per block pass
  const int c = tid % C, rl = tid / C;
  float s = 0.0f;
  if (rl < per) {
    // 8 independent loads in flight per thread (a serial loop pays one L2 round trip per row)
    const int64_t step = int64_t(gridDim.x) * per;
    int64_t r = blockIdx.x * int64_t(per) + rl;
    float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (; r + 7 * step < rows; r += 8 * step) {
#pragma unroll
      for (int q = 0; q < 8; ++q) part[q] += mapf_ld_dep(in + (r + q * step) * C + c);
    }
    for (; r < rows; r += step) part[0] += mapf_ld_dep(in + r * C + c);
    s = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
  }
  s_part[tid] = rl < per ? s : 0.0f;
  __syncthreads();
  if (tid < C) {
    float t = 0.0f;
    for (int q = 0; q < per; ++q) t += s_part[q * C + tid];
    atomicAdd(out + tid, t);
  }
}

// ---------------------------------------------------------------------------------------------
// graph-filter recurrence backward: dZ [rows][KT] -> dX [rows][F] (one CTA per episode)
//   x_k = x_{k-1} S  =>  dx_{k-1}[i] += sum_j S[i][j] dx_k[j];  skip: Xr = raw reshape of X[B,F,N]
// then dh = dX * [x > 0] (ReLU of the encoder FC)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) recurrence_bwd_kernel(const float* __restrict__ dz,
                                                             const float* __restrict__ gso,
                                                             const float* __restrict__ x, int N, int F, int K,
                                                             int has_skip, int add_self_loop, float* __restrict__ dx,
                                                             int64_t dx_sb, int64_t dx_sn, int64_t dx_sf) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  extern __shared__ __align__(16) float s_rb[];
  const int KT = (K + has_skip) * F;
  float* sm = s_rb;                 // [N][N]
  float* dinv = sm + N * N;         // [N]
  float* d = dinv + ((N + 3) & ~3); // [N][KT]
  const int tid = threadIdx.x, e = blockIdx.x;
  for (int idx = tid; idx < N * N; idx += 128) {
    const int i = idx / N, j = idx - i * N;
    float v = mapf_ld_dep(gso + int64_t(e) * N * N + idx);
    if (add_self_loop && i == j) v += 1.0f;
    sm[idx] = v;
  }
  for (int idx = tid; idx < N * KT; idx += 128) d[idx] = mapf_ld_dep(dz + int64_t(e) * N * KT + idx);
  __syncthreads();
  if (tid < N) {
    float deg = 0.0f;
    for (int j = 0; j < N; ++j) deg += sm[tid * N + j];
    dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;
  }
  __syncthreads();
  for (int idx = tid; idx < N * N; idx += 128) {
    const int i = idx / N, j = idx - i * N;
    sm[idx] = (dinv[i] * sm[idx]) * dinv[j];
  }
  __syncthreads();
  for (int k = K - 1; k >= 1; --k) {
    for (int idx = tid; idx < N * F; idx += 128) {
      const int i = idx / F, f = idx - i * F;
      float s = 0.0f;
      for (int j = 0; j < N; ++j) s = fmaf(sm[i * N + j], d[j * KT + k * F + f], s);
      d[i * KT + (k - 1) * F + f] += s;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < N * F; idx += 128) {
    const int i = idx / F, f = idx - i * F;
    float v = d[i * KT + f];
    if (has_skip) {
      const int q = f * N + i;  // X[b,f,n] sits at flat f*N+n = n'*F + f' of the reshaped [N,F] view
      v += d[(q / F) * KT + K * F + (q % F)];
    }
    if (x != nullptr && !(mapf_ld_dep(x + (int64_t(e) * N + i) * F + f) > 0.0f)) v = 0.0f;
    dx[int64_t(e) * dx_sb + int64_t(i) * dx_sn + int64_t(f) * dx_sf] = v;
  }
}

// [B,C,N] -> [B*N, C]
__global__ void bcn_to_rows_kernel(const float* __restrict__ in, int B, int C, int N, float* __restrict__ out) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int64_t total = int64_t(B) * C * N;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int64_t r = idx / C;
    const int n = int(r % N);
    const int64_t b = r / N;
    out[idx] = mapf_ld_dep(in + (b * C + c) * N + n);
  }
}

// ---------------------------------------------------------------------------------------------
// loss and optimiser
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_loss_kernel(mapf_ce_loss_args a) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int A = a.num_actions;
  float loss = 0.0f;
  if (r < a.rows) {
    float v[kMaxActions];
    float mx = -INFINITY;
    for (int q = 0; q < A; ++q) { v[q] = a.logits[r * A + q]; mx = fmaxf(mx, v[q]); }
    float sum = 0.0f;
    for (int q = 0; q < A; ++q) { v[q] = expf(v[q] - mx); sum += v[q]; }
    const int t = a.targets[r];
    const float inv = 1.0f / sum, invn = 1.0f / float(a.rows);
    for (int q = 0; q < A; ++q) a.dlogits[r * A + q] = (v[q] * inv - (q == t ? 1.0f : 0.0f)) * invn;
    loss = (logf(sum) + mx - a.logits[r * A + t]) * invn;
  }
  loss = warp_sum(loss);
  __shared__ float s_l[8];
  if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += s_l[w];
    atomicAdd(a.loss, t);
  }
}

// Both reductions below in ONE block where the input is small (64 k parameters; up to 2 048 rows): the result is stored,
// not accumulated -- no memset node in front of the kernel on the step's dependent chain
// (~4 us each at batch 128) and a fixed summation order.
constexpr int kOneBlockThreads = 1024;
__global__ void __launch_bounds__(kOneBlockThreads) ce_loss_one_block_kernel(mapf_ce_loss_args a) {
  mapf_pdl_wait();
  mapf_pdl_launch();
  const int A = a.num_actions;
  const float invn = 1.0f / float(a.rows);
  float loss = 0.0f;
  for (int64_t r = threadIdx.x; r < a.rows; r += kOneBlockThreads) {
    float v[kMaxActions];
    float mx = -INFINITY;
    for (int q = 0; q < A; ++q) { v[q] = a.logits[r * A + q]; mx = fmaxf(mx, v[q]); }
    float sum = 0.0f;
    for (int q = 0; q < A; ++q) { v[q] = expf(v[q] - mx); sum += v[q]; }
    const int t = a.targets[r];
    const float inv = 1.0f / sum;
    for (int q = 0; q < A; ++q) a.dlogits[r * A + q] = (v[q] * inv - (q == t ? 1.0f : 0.0f)) * invn;
    loss += (logf(sum) + mx - a.logits[r * A + t]) * invn;
  }
  loss = warp_sum(loss);
  __shared__ float s_l[kOneBlockThreads / 32];
  if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < kOneBlockThreads / 32; ++w) t += s_l[w];
    *a.loss = t;
  }
}

__global__ void __launch_bounds__(kOneBlockThreads) sumsq_one_block_kernel(const float* __restrict__ g, int64_t n,
                                                                            double* __restrict__ out, int32_t* step_dev) {
  mapf_pdl_wait();
  mapf_pdl_launch();
  if (step_dev != nullptr && threadIdx.x == 0) *step_dev += 1;   // read by the next kernel
  double s = 0.0;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(g) & 15) == 0 ? n >> 2 : 0;
  for (int64_t i = threadIdx.x; i < n4; i += kOneBlockThreads) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += double(v.x) * double(v.x) + double(v.y) * double(v.y) + double(v.z) * double(v.z) + double(v.w) * double(v.w);
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += kOneBlockThreads) s += double(g[i]) * double(g[i]);
  s = warp_sum(s);
  __shared__ double s_s[kOneBlockThreads / 32];
  if ((threadIdx.x & 31) == 0) s_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kOneBlockThreads / 32; ++w) t += s_s[w];
    *out = t;
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out,
                                                    int32_t* step_dev) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  if (step_dev != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;   // read by the next kernel
  double s = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    s += double(g[i]) * double(g[i]);
  s = warp_sum(s);
  __shared__ double s_s[8];
  if ((threadIdx.x & 31) == 0) s_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_s[w];
    atomicAdd(out, t);
  }
}

// One-shot gradient all-reduce over NVLink peer memory, fused with the gradient-norm reduction (47.8 k - 64.2 k
// floats, pure latency).  Every rank publishes its flat gradient in a peer-mapped buffer; this kernel (same launch on every
// rank, no NCCL call, capturable in a CUDA graph) signals "my gradient of epoch e is complete" into every peer's flag
// array, waits for all peers' signals, then every CTA sums its slice of ALL ranks' buffers in rank order (identical
// bits on every rank), scales by 1 / world, writes the averaged gradient to the local buffer clip_adam_kernel reads and
// accumulates its squared norm.  Two gradient buffers alternate by step parity: a rank that races ahead writes the other
// buffer, and cannot reach the step after that before every peer has signalled the next epoch, i.e. finished reading.
struct PeerReduceArgs {
  int world, rank;
  int64_t n;                 // floats, multiple of 4
  const float* grads[8];     // peer-mapped gradient buffers of this step's parity ([rank] = own)
  int32_t* flags[8];         // peer-mapped flag blocks: int32[8] arrival epochs + [8] = the owner's epoch counter
  float* out;                // local averaged gradient
  double* sumsq;
  int32_t* step_dev;
};

__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {   // never from a stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256) peer_reduce_sumsq_kernel(PeerReduceArgs a) {
  mapf_pdl_wait();   // the backward kernels in front have completed: the local gradient is final
  mapf_pdl_launch();
  int32_t* const mine = a.flags[a.rank];
  const int epoch = ld_acquire_sys(mine + 8) + 1;   // incremented by the clip_adam kernel behind this one
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0 && a.step_dev != nullptr) *a.step_dev += 1;   // read by the next kernel
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(a.flags[threadIdx.x] + a.rank, epoch);
    }
  }
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {
      __nanosleep(64);
      if (clock64() - t0 > 20000000000LL) asm volatile("trap;");   // ~10 s: a missing peer must not hang the GPU
    }
  }
  __syncthreads();
  const float inv = 1.0f / float(a.world);
  double s = 0.0;
  for (int64_t i = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) * 4; i < a.n; i += int64_t(gridDim.x) * blockDim.x * 4) {
    float4 acc = ld_peer_v4(a.grads[0] + i);
    for (int r = 1; r < a.world; ++r) {
      const float4 v = ld_peer_v4(a.grads[r] + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    *reinterpret_cast<float4*>(a.out + i) = acc;
    s += double(acc.x) * acc.x + double(acc.y) * acc.y + double(acc.z) * acc.z + double(acc.w) * acc.w;
  }
  s = warp_sum(s);
  __shared__ double s_s[8];
  if ((threadIdx.x & 31) == 0) s_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_s[w];
    atomicAdd(a.sumsq, t);
  }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(mapf_clip_adam_args a, float bc1, float bc2_sqrt,
                                                        int32_t* epoch_inc) {
  mapf_pdl_wait();   // programmatic dependent launch: this grid may be resident before its predecessor completes
  mapf_pdl_launch();
  if (epoch_inc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *epoch_inc += 1;   // peer all-reduce epoch
  if (a.step_dev != nullptr) {   // device-resident step count (CUDA-graph replay): bias corrections computed here
    const double t = double(*a.step_dev);
    bc1 = float(1.0 - pow(double(a.beta1), t));
    bc2_sqrt = float(sqrt(1.0 - pow(double(a.beta2), t)));
  }
  if (a.lr_dev != nullptr) a.lr = *a.lr_dev;
  const float total = float(sqrt(*a.scratch));
  const float coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);   // clip_grad_norm_ (train.py:97)
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.grad_norm = total;
  const float step_size = a.lr / bc1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < a.n; i += int64_t(gridDim.x) * blockDim.x) {
    const float p = a.params[i];
    float g = a.grads[i] * coef;
    a.grads[i] = g;
    g = fmaf(a.weight_decay, p, g);                                 // Adam L2 weight decay
    const float m = a.beta1 * a.exp_avg[i] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.exp_avg_sq[i] + (1.0f - a.beta2) * g * g;
    a.exp_avg[i] = m;
    a.exp_avg_sq[i] = v;
    a.params[i] = p - step_size * (m / (sqrtf(v) / bc2_sqrt + a.eps));
  }
}

// d act = (dlogits W) * [act > 0]: the ReLU-masked input gradient of the action head (actionMLP, framework_gnn.py:171-181
// backward); rows x D outputs, A <= 8 actions.  Memory-bound: one float4 of D per thread.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                                       const float* __restrict__ mask, int64_t rows, int D, int A,
                                                       float* __restrict__ out) {
  extern __shared__ __align__(16) float s_hw[];   // [A][D]
  mapf_pdl_wait();
  mapf_pdl_launch();
  for (int i = threadIdx.x; i < A * D; i += blockDim.x) s_hw[i] = mapf_ld_dep(w + i);
  __syncthreads();
  const int d4 = D >> 2;
  const int64_t total = rows * d4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / d4;
    const int c = int(idx - r * d4) << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < A; ++a) {
      const float g = mapf_ld_dep(dl + r * A + a);
      const float4 wv = *reinterpret_cast<const float4*>(s_hw + a * D + c);
      acc.x = fmaf(g, wv.x, acc.x); acc.y = fmaf(g, wv.y, acc.y); acc.z = fmaf(g, wv.z, acc.z); acc.w = fmaf(g, wv.w, acc.w);
    }
    const float4 m = mapf_ld_dep(reinterpret_cast<const float4*>(mask + r * D + c));
    acc.x = m.x > 0.0f ? acc.x : 0.0f; acc.y = m.y > 0.0f ? acc.y : 0.0f;
    acc.z = m.z > 0.0f ? acc.z : 0.0f; acc.w = m.w > 0.0f ? acc.w : 0.0f;
    *reinterpret_cast<float4*>(out + r * D + c) = acc;
  }
}

int launch_conv_train(cudaStream_t s, const float* in, int cin, int cout, int B, int N, const float* wpk,
                      const float* bias, float* out, double* stats, int stat_c) {
  const size_t smem = sizeof(float) * (size_t(cin) * kFovCells * kCtPad + ((size_t(cout) * cin * 9 + 3) & ~size_t(3)));
  if (smem > 220 * 1024) return MAPF_ERR_UNSUPPORTED;
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(conv_train_kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  dim3 grid((B + 31) / 32, N);
  mapf_launch_pdl(conv_train_kernel, dim3(grid), dim3(kCtThreads), smem, s, in, cin, cout, B, N, wpk, bias, out, stats, stat_c);
  return mapf_launch_status();
}

#define MAPF_TRY(expr)              \
  do {                              \
    const int rc__ = (expr);        \
    if (rc__ != MAPF_OK) return rc__; \
  } while (0)

int check_train(const mapf_net_params* p, int B, int N) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(B > 0 && N > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 64 && N <= 65535, MAPF_ERR_UNSUPPORTED);
  for (int l = 0; l < p->num_conv; ++l)
    MAPF_REQUIRE(p->channels[l] * p->channels[l + 1] <= 4 * kCwThreads, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->fc_out <= 256 && p->num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  return MAPF_OK;
}

// ---------------------------------------------------------------------------------------------
// Auxiliary lanes.  A training step is a chain of ~55 short launches; at batch 128 each is latency-bound and
// the chain length IS the step time.  Two classes of launches hang off the chain without feeding it: the weight
// packs (depend on the parameters only) and the weight / bias gradients (nothing downstream reads them).  They
// run on two auxiliary streams per host thread and device, ordered against the caller's stream by events --
// under stream capture the same calls become parallel branches of the graph.  The lanes are created on the
// first un-captured call; a first call made under capture stays on the caller's stream.
// ---------------------------------------------------------------------------------------------
struct Lanes {
  cudaStream_t lane[2] = {nullptr, nullptr};
  cudaEvent_t ev[40] = {};
  bool ready = false;
};

Lanes* lanes_for(cudaStream_t main, int flags) {
  if (flags & MAPF_TRAIN_SINGLE_LANE) return nullptr;
  thread_local Lanes tl[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  Lanes& L = tl[dev];
  if (!L.ready) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
      (void)cudaGetLastError();
      return nullptr;
    }
    for (auto& ln : L.lane)
      if (cudaStreamCreateWithFlags(&ln, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : L.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    L.ready = true;
  }
  return &L;
}

// Event bookkeeping of one forward / backward call: mark(stream) = "everything issued on `stream` so far",
// wait(stream, mark) orders later work on `stream` behind it.  With lanes == nullptr everything is one stream
// and both are no-ops.
struct LaneOrder {
  Lanes* lanes;
  int used = 0;
  bool ok = true;
  explicit LaneOrder(Lanes* l) : lanes(l) {}
  int mark(cudaStream_t from) {
    if (!lanes) return -1;
    if (used >= int(sizeof(lanes->ev) / sizeof(lanes->ev[0]))) { ok = false; return -1; }
    ok = ok && cudaEventRecord(lanes->ev[used], from) == cudaSuccess;
    return used++;
  }
  void wait(cudaStream_t to, int m) {
    if (!lanes || m < 0) return;
    ok = ok && cudaStreamWaitEvent(to, lanes->ev[m], 0) == cudaSuccess;
  }
};

// The backward pass' parameter-only operands -- transposed weights and their tensor-core packs for the activation-gradient
// GEMMs, mirrored / transposed filter operands for the data-gradient convolutions -- on stream `sp`.  Called from the
// FORWARD call's pack lane under MAPF_TRAIN_PREPARED (the packs are then long complete when the backward chain needs them:
// at batch 128 the first activation-gradient GEMM otherwise waits ~15 us for them), else from the backward call itself.
struct BwdPackMarks { int gf = -1, fc = -1, conv[MAPF_MAX_CONV]; };
int pack_backward_operands(const mapf_net_params* p, int message, int flags, const TrainLayout& T, float* ws, cudaStream_t sp,
                           LaneOrder* ord, BwdPackMarks* m) {
  const int L = p->num_conv, F = p->fc_out, fin = p->fc_in, Gd = p->node_dim, K = p->taps;
  const int head_in = K > 0 ? Gd : F;
  const int KT = (K + (message ? 1 : 0)) * F;
  const bool tc_bwd = !(flags & MAPF_TRAIN_FFMA_BWD) && (head_in & 3) == 0 && (F & 3) == 0 && (fin & 3) == 0 &&
                      (K == 0 || (Gd & 3) == 0) && F <= 128 && Gd <= 128;
  const bool conv_tc = !(flags & MAPF_TRAIN_FFMA_CONV);
  if (tc_bwd) {
    if (K > 0) {
      // dz = dy W_eff (and the W2 columns behind): W_eff^T rows are the packed B operand
      MAPF_TRY(mapf_pack_graph_weights(p->gf_w, message ? p->gf_w2 : nullptr, F, K, Gd, ws + T.wtt, sp));
      MAPF_TRY(mapf_tc_pack_b_tiles(ws + T.wtt, Gd, KT, ws + T.gfpk_bwd, sp));
      if (ord) m->gf = ord->mark(sp);
    }
    MAPF_TRY(mapf_pack_graph_weights(p->fc_w, nullptr, fin, 1, F, ws + T.fcwt, sp));   // fc_w^T [fc_in][F]
    MAPF_TRY(mapf_tc_pack_b_tiles(ws + T.fcwt, F, fin, ws + T.fcpk_bwd, sp));
    if (ord) m->fc = ord->mark(sp);
  }
  for (int l = L - 1; l >= 1; --l) {
    const int cin = p->channels[l], cout = p->channels[l + 1];
    if (conv_tc && mapf_conv_tc_layer_ok(cin, cout, 1)) {
      MAPF_TRY(mapf_conv_tc_layer_pack(p->conv_w[l], cin, cout, 1, ws + T.convtc_bwd[l], sp));
    } else {
      const int cin_pad = int(r4(cin));
      MAPF_REQUIRE(cin_pad == cin, MAPF_ERR_UNSUPPORTED);
      mapf_launch_pdl(pack_conv_kernel, dim3((cin_pad * cout * 9 + 255) / 256), dim3(256), 0, sp, p->conv_w[l], cout, cin, 1, ws + T.convpk_bwd[l]);
    }
    if (ord) m->conv[l] = ord->mark(sp);
  }
  return MAPF_OK;
}

}  // namespace

extern "C" int64_t mapf_train_workspace_size(const mapf_net_params* p, int32_t batch, int32_t agents) {
  if (check_train(p, batch, agents) != MAPF_OK) return -1;
  return train_layout(*p, batch, agents, 1).total;
}

extern "C" int mapf_train_forward(const mapf_net_params* p, const mapf_train_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_TRY(check_train(p, a->batch, a->agents));
  MAPF_REQUIRE(a->states && a->workspace && a->logits && p->fc_w && p->fc_b && p->act_w && p->act_b, MAPF_ERR_NULL);
  MAPF_REQUIRE(p->taps == 0 || (a->gso && p->gf_w), MAPF_ERR_NULL);
  MAPF_REQUIRE(!a->message || p->gf_w2, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_aligned(a->workspace, 16), MAPF_ERR_ALIGN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = a->batch, N = a->agents, L = p->num_conv;
  const int64_t rows = int64_t(B) * N;
  const TrainLayout T = train_layout(*p, B, N, 1);
  float* ws = a->workspace;
  double* stat = reinterpret_cast<double*>(ws + T.stat_fwd);
  const double count = double(B) * kFovCells;
  const int F = p->fc_out, fin = p->fc_in, A = p->num_actions;
  const int KT = (p->taps + (a->message ? 1 : 0)) * F;
  // tensor-core convolutions (fp16-split operands, enc_tc.cu) wherever the layer shape allows; FFMA2 kernel otherwise
  const bool conv_tc = !(a->flags & MAPF_TRAIN_FFMA_CONV);
  const bool fc_tc = F <= 128 && (fin & 3) == 0 && !(a->flags & MAPF_TRAIN_FFMA_FC);
  const bool graph_tc = p->taps > 0 && N <= 64 && p->node_dim <= 128 && (p->node_dim & 15) == 0 && (F & 3) == 0 &&
                        mapf_aligned(p->act_w, 16) && !(a->flags & MAPF_TRAIN_FFMA_GRAPH);
  // weight packs (they depend on the parameters only) and the zeroing of the BatchNorm sums on the pack lane, in the
  // order the chain consumes them
  LaneOrder ord(lanes_for(s, a->flags));
  cudaStream_t sp = ord.lanes ? ord.lanes->lane[1] : s;
  ord.wait(sp, ord.mark(s));
  const bool prepare = (a->flags & MAPF_TRAIN_PREPARED) != 0;
  // (the backward call's sums sit right behind the forward's: one memset zeroes both)
  static_assert(sizeof(double) == 2 * sizeof(float), "stat blocks are sized in floats");
  MAPF_REQUIRE(!prepare || T.stat_bwd == T.stat_fwd + int64_t(2) * L * N * T.cmax * 2, MAPF_ERR_UNSUPPORTED);
  if (cudaMemsetAsync(stat, 0, sizeof(double) * size_t(2) * L * N * T.cmax * (prepare ? 2 : 1), sp) != cudaSuccess) return MAPF_ERR_CUDA;
  int m_conv[MAPF_MAX_CONV];
  for (int l = 0; l < L; ++l) {
    const int cin = p->channels[l], cout = p->channels[l + 1];
    MAPF_REQUIRE(p->conv_w[l] && p->conv_b[l] && p->bn_gamma[l] && p->bn_beta[l], MAPF_ERR_NULL);
    if (conv_tc && mapf_conv_tc_layer_ok(cin, cout, 0))
      MAPF_TRY(mapf_conv_tc_layer_pack(p->conv_w[l], cin, cout, 0, ws + T.convtc[l], sp));
    else
      mapf_launch_pdl(pack_conv_kernel, dim3((cout * cin * 9 + 255) / 256), dim3(256), 0, sp, p->conv_w[l], cout, cin, 0, ws + T.convpk[l]);
    m_conv[l] = ord.mark(sp);
  }
  // 3xTF32 tcgen05 operand of the current compressMLP weights (re-packed every step: 100 KB)
  if (fc_tc) MAPF_TRY(mapf_tc_pack_b(p->fc_w, fin, nullptr, 0, F, ws + T.fcpk, sp));
  const int m_fc = ord.mark(sp);
  // the backward call's operands: behind the packs the forward chain needs first, in front of the graph-filter pack whose
  // mark doubles as the join of the lane
  if (prepare) MAPF_TRY(pack_backward_operands(p, a->message, a->flags, T, ws, sp, nullptr, nullptr));
  if (graph_tc)
    MAPF_TRY(mapf_tc_pack_b(p->gf_w, p->taps * F, a->message ? p->gf_w2 : nullptr, a->message ? F : 0, p->node_dim,
                            ws + T.wt, sp));
  else if (p->taps > 0)
    MAPF_TRY(mapf_pack_graph_weights(p->gf_w, a->message ? p->gf_w2 : nullptr, F, p->taps, p->node_dim, ws + T.wt, sp));
  const int m_gf = ord.mark(sp);

  // BatchNorm + ReLU of layer l - 1 applied by layer l's tensor-core kernel on load (one launch less per layer on the
  // step's dependent chain, one pass over the activations less): the last layer keeps its own pass (the FC reads a_L)
  bool bn_pending = false;
  for (int l = 0; l < L; ++l) {
    const int cin = p->channels[l], cout = p->channels[l + 1];
    double* st = stat + int64_t(l) * N * T.cmax * 2;
    ord.wait(s, m_conv[l]);
    if (conv_tc && mapf_conv_tc_layer_ok(cin, cout, 0)) {
      MapfConvBnIn bn{};
      if (bn_pending) {
        bn.stats = stat + int64_t(l - 1) * N * T.cmax * 2;
        bn.gamma = p->bn_gamma[l - 1]; bn.beta = p->bn_beta[l - 1];
        bn.count = count; bn.momentum = a->momentum; bn.stat_c = T.cmax;
        bn.meaninv = ws + T.meaninv[l - 1]; bn.scale_shift = ws + T.scsh[l - 1];
        bn.running_mean = const_cast<float*>(p->bn_mean[l - 1]); bn.running_var = const_cast<float*>(p->bn_var[l - 1]);
        bn.batches = a->bn_batches[l - 1];
        bn.act_out = ws + T.a[l - 1];
      }
      MAPF_TRY(mapf_conv_tc_layer(l == 0 ? a->states : (bn_pending ? ws + T.u[l - 1] : ws + T.a[l - 1]), p->conv_w[l], p->conv_b[l],
                                  cin, cout, 0, B, N, ws + T.convtc[l], 1, ws + T.u[l], st, T.cmax, bn_pending ? &bn : nullptr, stream));
      bn_pending = false;
    } else {
      MAPF_TRY(launch_conv_train(s, l == 0 ? a->states : ws + T.a[l - 1], cin, cout, B, N, ws + T.convpk[l],
                                 p->conv_b[l], ws + T.u[l], st, T.cmax));
    }
    const int64_t planes = rows * cout;
    const bool vec4 = (cout * kFovCells) % 4 == 0;
    if (l + 1 < L && conv_tc && !(a->flags & MAPF_TRAIN_BN_PASS) && mapf_conv_tc_layer_bn_ok(cout, p->channels[l + 2], N)) {
      bn_pending = true;      // the next layer's kernel normalises on load
      continue;
    }
    if (vec4 && N * cout <= kBnFusedMax) {
      // batch statistics -> scale / shift -> BatchNorm + ReLU in one launch
      mapf_launch_pdl(bn_finalize_apply4_kernel, dim3(unsigned(mapf_min64((planes * kFovCells / 4 + 255) / 256, 148 * 16))), dim3(256), 0, s,
          reinterpret_cast<const float4*>(ws + T.u[l]), st, N, cout, T.cmax, count, a->momentum, ws + T.meaninv[l],
          const_cast<float*>(p->bn_mean[l]), const_cast<float*>(p->bn_var[l]), a->bn_batches[l], p->bn_gamma[l], p->bn_beta[l],
          ws + T.scsh[l], rows, reinterpret_cast<float4*>(ws + T.a[l]));
      continue;
    }
    mapf_launch_pdl(bn_finalize_kernel, dim3((cout + 63) / 64), dim3(64), 0, s, st, N, cout, T.cmax, count, a->momentum, ws + T.meaninv[l],
                                                      const_cast<float*>(p->bn_mean[l]),
                                                      const_cast<float*>(p->bn_var[l]), a->bn_batches[l], p->bn_gamma[l],
                                                      p->bn_beta[l], ws + T.scsh[l]);
    if (vec4)
      mapf_launch_pdl(bn_apply4_kernel, dim3(unsigned(mapf_min64((planes * kFovCells / 4 + 255) / 256, 148 * 16))), dim3(256), 0, s, 
          reinterpret_cast<const float4*>(ws + T.u[l]), ws + T.scsh[l], N, cout, rows, reinterpret_cast<float4*>(ws + T.a[l]));
    else
      mapf_launch_pdl(bn_apply_kernel, dim3(unsigned(mapf_min64((planes * kFovCells + 255) / 256, 148 * 16))), dim3(256), 0, s, ws + T.u[l], ws + T.scsh[l], N,
                                                                                        cout, planes, ws + T.a[l]);
  }
  MAPF_TRY(mapf_launch_status());
  // compressMLP: x = relu(a_L Wfc^T + b)
  ord.wait(s, m_fc);
  if (fc_tc) {
    MAPF_TRY(mapf_tc_gemm(int32_t(rows), F, fin, ws + T.a[L - 1], fin, ws + T.fcpk, ws + T.x, F, p->fc_b, 1, stream));
  } else {
    MAPF_TRY(gemm(s, int(rows), F, fin, ws + T.a[L - 1], fin, 1, p->fc_w, 1, fin, ws + T.x, F, p->fc_b, 1));
  }
  ord.wait(s, m_gf);         // also the join of the pack lane
  MAPF_REQUIRE(ord.ok, MAPF_ERR_CUDA);
  if (p->taps == 0) return gemm(s, int(rows), A, F, ws + T.x, F, 1, p->act_w, 1, F, a->logits, A, p->act_b, 0);
  mapf_graph_filter_fwd_args g{};
  g.batch = B; g.agents = N; g.in_dim = F; g.out_dim = p->node_dim; g.taps = p->taps;
  g.add_self_loop = a->message ? 0 : 1;
  g.has_skip = a->message ? 1 : 0;
  g.relu = 1;
  g.num_actions = A;
  g.x = ws + T.x; g.x_sb = int64_t(N) * F; g.x_sn = F; g.x_sf = 1;
  g.gso = a->gso;
  g.wt = ws + T.wt;
  g.y = ws + T.y; g.y_sb = int64_t(N) * p->node_dim; g.y_sn = p->node_dim; g.y_sg = 1;
  g.z = ws + T.z;
  g.act_w = p->act_w; g.act_b = p->act_b;
  g.logits = a->logits; g.actions = nullptr;
  if (graph_tc) {
    // filter taps Z (saved for the weight gradient) + 3xTF32 tcgen05 mix with relu(Y) stored for the backward pass
    MAPF_TRY(mapf_graph_taps(&g, stream));
    return mapf_tc_mix_head(int32_t(rows), p->node_dim, KT, ws + T.z, KT, ws + T.wt, p->act_w, p->act_b, A, a->logits,
                            nullptr, ws + T.y, stream);
  }
  return mapf_graph_filter_fwd(&g, stream);
}

extern "C" int mapf_train_backward(const mapf_net_params* p, const mapf_train_bwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_TRY(check_train(p, a->batch, a->agents));
  MAPF_REQUIRE(a->states && a->workspace && a->dlogits, MAPF_ERR_NULL);
  const mapf_net_grads& G = a->grads;
  MAPF_REQUIRE(G.fc_w && G.fc_b && G.act_w && G.act_b, MAPF_ERR_NULL);
  MAPF_REQUIRE(p->taps == 0 || (a->gso && G.gf_w), MAPF_ERR_NULL);
  MAPF_REQUIRE(!a->message || G.gf_w2, MAPF_ERR_NULL);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = a->batch, N = a->agents, L = p->num_conv;
  const int F = p->fc_out, fin = p->fc_in, A = p->num_actions, Gd = p->node_dim, K = p->taps;
  const int64_t rows = int64_t(B) * N;
  const TrainLayout T = train_layout(*p, B, N, 1);
  float* ws = a->workspace;
  const int head_in = K > 0 ? Gd : F;
  // Lanes: `s` carries the activation-gradient chain; `sw` the weight / bias gradients (read the chain's tensors, feed
  // nothing downstream); `sp` the transposed-weight packs (depend on the parameters only).
  LaneOrder ord(lanes_for(s, a->flags));
  cudaStream_t sw = ord.lanes ? ord.lanes->lane[0] : s;
  cudaStream_t sp = ord.lanes ? ord.lanes->lane[1] : s;
  {
    const int m0 = ord.mark(s);
    ord.wait(sw, m0);
    ord.wait(sp, m0);
  }
  // Zero the accumulated gradients: adjacent buffers (a trainer keeps every gradient as a view of ONE flat tensor) are
  // merged into one memset -- twelve separate memset nodes were a 40 us chain at the head of the weight-gradient lane,
  // the lane that finishes last at small batches.  The BatchNorm gradients are overwritten later on the same lane
  // (bn_param_grads_kernel): they are only swept along where they bridge two ranges.
  struct Range { float* p; int64_t n; bool needed; };
  Range rg[6 + 4 * MAPF_MAX_CONV];
  int nrg = 0;
  auto want = [&](float* ptr, int64_t n, bool needed) { if (ptr != nullptr && n > 0) rg[nrg++] = Range{ptr, n, needed}; };
  want(G.act_w, int64_t(A) * head_in, true); want(G.act_b, A, true); want(G.fc_w, int64_t(F) * fin, true); want(G.fc_b, F, true);
  if (K > 0) want(G.gf_w, int64_t(F) * K * Gd, true);
  if (a->message) want(G.gf_w2, int64_t(F) * Gd, true);
  for (int l = 0; l < L; ++l) {
    MAPF_REQUIRE(G.conv_w[l] && G.conv_b[l] && G.bn_gamma[l] && G.bn_beta[l], MAPF_ERR_NULL);
    want(G.conv_w[l], int64_t(p->channels[l]) * p->channels[l + 1] * 9, true);
    want(G.conv_b[l], p->channels[l + 1], true);
    // (written by bn_param_grads_kernel later on this lane: swept along where they bridge two ranges -- but not where the
    // chain's own bn_bwd_coef_apply4_kernel writes them, which is not ordered behind this lane's memsets)
    const bool on_chain = (p->channels[l + 1] * kFovCells) % 4 == 0 && N * p->channels[l + 1] <= kBnFusedMax && p->channels[l + 1] <= 256;
    if (!on_chain) {
      want(G.bn_gamma[l], p->channels[l + 1], false);
      want(G.bn_beta[l], p->channels[l + 1], false);
    }
  }
  for (int i = 1; i < nrg; ++i)          // insertion sort by address (at most 18 entries)
    for (int j = i; j > 0 && rg[j].p < rg[j - 1].p; --j) { const Range t = rg[j]; rg[j] = rg[j - 1]; rg[j - 1] = t; }
  bool ok = true;
  for (int i = 0; i < nrg;) {
    int j = i;
    float* end = rg[i].p + rg[i].n;
    bool needed = rg[i].needed;
    while (j + 1 < nrg && rg[j + 1].p == end) { ++j; end += rg[j].n; needed = needed || rg[j].needed; }
    if (needed) ok = ok && cudaMemsetAsync(rg[i].p, 0, sizeof(float) * size_t(end - rg[i].p), sw) == cudaSuccess;
    i = j + 1;
  }
  ord.wait(sp, ord.mark(sw));      // (the pack lane also carries every other layer's weight gradient: behind the zeroing)
  double* stat = reinterpret_cast<double*>(ws + T.stat_bwd);
  const bool prepared = (a->flags & MAPF_TRAIN_PREPARED) != 0;     // packs and zeroed sums come from the forward call
  if (!prepared) ok = ok && cudaMemsetAsync(stat, 0, sizeof(double) * size_t(2) * L * N * T.cmax, sp) == cudaSuccess;
  if (!ok) return MAPF_ERR_CUDA;
  const int m_stat = prepared ? -1 : ord.mark(sp);

  const float* head_act = K > 0 ? ws + T.y : ws + T.x;
  const int KF = K * F, KT = (K + (a->message ? 1 : 0)) * F;
  // Tensor-core path (default): weight gradients = mapf_tc_gemm_tn (contraction over the rows, split K), activation
  // gradients = mapf_tc_gemm against transposed-weight packs, head input gradient = head_bwd_kernel.
  // MAPF_TRAIN_FFMA_BWD keeps the FFMA2 sgemm kernels (development switch).
  const bool tc_bwd = !(a->flags & MAPF_TRAIN_FFMA_BWD) && (head_in & 3) == 0 && (F & 3) == 0 && (fin & 3) == 0 &&
                      (K == 0 || (Gd & 3) == 0) && F <= 128 && Gd <= 128;
  const bool conv_tc = !(a->flags & MAPF_TRAIN_FFMA_CONV);
  // three activation-gradient buffers in rotation: a layer's data-gradient convolution writes the buffer read two layers up,
  // so the chain never waits for the weight-gradient kernel of the layer just above (which reads the buffer in between)
  float* da = ws + T.da0;
  float* da_next = ws + T.da1;
  float* da_spare = ws + T.da2;
  // C[rows, Ncols] = Asrc[rows, Kd] . Wt^T with Wt [Ncols, Kd] row-major (a transposed weight): tiles of 128 output
  // columns packed on the pack lane, one GEMM launch over all tiles on the chain
  auto nn_gemm = [&](const float* Asrc, int Kd, int Ncols, const float* packs, float* Cout, int64_t ldc) -> int {
    return mapf_tc_gemm_tiles(int32_t(rows), Ncols, Kd, Asrc, Kd, packs, Cout, ldc, nullptr, 0, stream);
  };
  // ---- pack lane ----
  BwdPackMarks pk;
  for (int l = 0; l < MAPF_MAX_CONV; ++l) pk.conv[l] = -1;
  if (!prepared) MAPF_TRY(pack_backward_operands(p, a->message, a->flags, T, ws, sp, &ord, &pk));
  const int m_gfpk = pk.gf, m_fcpk = pk.fc;
  const int* m_convpk = pk.conv;

  // ---- actionMLP: db = colsum(dl); dW[a][g] = sum_r dl[r][a] act[r][g]; d act = (dl Wa) * [act > 0] ----
  mapf_launch_pdl(colsum_kernel, dim3(int(mapf_min64((rows + 255) / 256, 296))), dim3(256), 0, sw, a->dlogits, rows, A, G.act_b);
  if (tc_bwd) {
    MAPF_TRY(mapf_tc_gemm_tn(A, head_in, rows, a->dlogits, A, head_act, head_in, G.act_w, head_in, 0, 1, sw));
    const size_t hsm = sizeof(float) * size_t(A) * head_in;
    float* dhead = K > 0 ? ws + T.dy : ws + T.dx;
    mapf_launch_pdl(head_bwd_kernel, dim3(unsigned(mapf_min64((rows * (head_in / 4) + 255) / 256, 148 * 8))), dim3(256), hsm, s,
                    a->dlogits, p->act_w, head_act, rows, head_in, A, dhead);
  } else {
    MAPF_TRY(gemm(sw, A, head_in, int(rows), a->dlogits, 1, A, head_act, head_in, 1, G.act_w, head_in, nullptr, 0, nullptr,
                  0, splits_for(rows, (head_in + 63) / 64)));
    if (K > 0)
      MAPF_TRY(gemm(s, int(rows), Gd, A, a->dlogits, A, 1, p->act_w, Gd, 1, ws + T.dy, Gd, nullptr, 0, ws + T.y, Gd));
    else
      MAPF_TRY(gemm(s, int(rows), F, A, a->dlogits, A, 1, p->act_w, F, 1, ws + T.dx, F, nullptr, 0, ws + T.x, F));
  }
  // ---- graph filter ----
  if (K > 0) {
    ord.wait(sw, ord.mark(s));        // dy
    // dW_eff[g][j] = sum_r dy[r][g] z[r][j]  -- row-major [G][K*F] IS the memory of GFL.0.W [F,K,G] (gnn.py:114-116)
    if (tc_bwd) {
      MAPF_TRY(mapf_tc_gemm_tn(Gd, KF, rows, ws + T.dy, Gd, ws + T.z, KT, G.gf_w, KF, 0, 1, sw));
      if (a->message) MAPF_TRY(mapf_tc_gemm_tn(Gd, F, rows, ws + T.dy, Gd, ws + T.z + KF, KT, G.gf_w2, F, 0, 1, sw));
      ord.wait(s, m_gfpk);
      MAPF_TRY(nn_gemm(ws + T.dy, Gd, KT, ws + T.gfpk_bwd, ws + T.dz, KT));
    } else {
      MAPF_TRY(gemm(sw, Gd, KF, int(rows), ws + T.dy, 1, Gd, ws + T.z, KT, 1, G.gf_w, KF, nullptr, 0, nullptr, 0,
                    splits_for(rows, ((Gd + 127) / 128) * ((KF + 63) / 64))));
      if (a->message)
        MAPF_TRY(gemm(sw, Gd, F, int(rows), ws + T.dy, 1, Gd, ws + T.z + KF, KT, 1, G.gf_w2, F, nullptr, 0, nullptr, 0,
                      splits_for(rows, ((Gd + 127) / 128) * ((F + 63) / 64))));
      // dz[r][j] = sum_g dy[r][g] W_eff[g][j]
      MAPF_TRY(gemm(s, int(rows), KF, Gd, ws + T.dy, Gd, 1, p->gf_w, KF, 1, ws + T.dz, KT));
      if (a->message) MAPF_TRY(gemm(s, int(rows), F, Gd, ws + T.dy, Gd, 1, p->gf_w2, F, 1, ws + T.dz + KF, KT));
    }
    const size_t smem = sizeof(float) * (size_t(N) * N + ((N + 3) & ~3) + size_t(N) * KT);
    MAPF_REQUIRE(smem <= 220 * 1024, MAPF_ERR_UNSUPPORTED);
    if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(recurrence_bwd_kernel, size_t(smem), cache__); }())
      return MAPF_ERR_CUDA;
    mapf_launch_pdl(recurrence_bwd_kernel, dim3(B), dim3(128), smem, s, ws + T.dz, a->gso, ws + T.x, N, F, K, a->message ? 1 : 0,
                                               a->message ? 0 : 1, ws + T.dx, int64_t(N) * F, F, 1);
  }
  // ---- compressMLP ----
  ord.wait(sw, ord.mark(s));          // dx
  mapf_launch_pdl(colsum_kernel, dim3(int(mapf_min64((rows + 255) / 256, 296))), dim3(256), 0, sw, ws + T.dx, rows, F, G.fc_b);
  if (tc_bwd) {
    MAPF_TRY(mapf_tc_gemm_tn(F, fin, rows, ws + T.dx, F, ws + T.a[L - 1], fin, G.fc_w, fin, 0, 1, sw));
    ord.wait(s, m_fcpk);
    MAPF_TRY(nn_gemm(ws + T.dx, F, fin, ws + T.fcpk_bwd, da, fin));
  } else {
    MAPF_TRY(gemm(sw, F, fin, int(rows), ws + T.dx, 1, F, ws + T.a[L - 1], fin, 1, G.fc_w, fin, nullptr, 0, nullptr, 0,
                  splits_for(rows, ((F + 127) / 128) * ((fin + 63) / 64))));
    MAPF_TRY(gemm(s, int(rows), fin, F, ws + T.dx, F, 1, p->fc_w, fin, 1, da, fin));
  }
  // ---- conv stack ----
  const double count = double(B) * kFovCells;
  ord.wait(s, m_stat);
  int m_dw_prev = -1, m_dw_prev2 = -1; // weight-gradient kernels of the two layers above: the one two layers up reads the
                                      // buffer this layer's data gradient is about to overwrite
  for (int l = L - 1; l >= 0; --l) {
    const int cin = p->channels[l], cout = p->channels[l + 1];
    double* st = stat + int64_t(l) * N * T.cmax * 2;
    dim3 grid((B + 31) / 32, N);
    const bool vec4 = (cout * kFovCells) % 4 == 0 && cout * kFovCells / 4 <= 256 && cout <= 64;
    // small batches are latency-bound: fewer rows per CTA (a CTA walks its rows serially) until the grid fills the SMs
    const int rpc = int64_t(B) * N >= 32 * 2 * 148 ? 32 : (int64_t(B) * N >= 8 * 2 * 148 ? 8 : 4);
    if (vec4)
      mapf_launch_pdl(bn_bwd_reduce4_kernel, dim3((B + rpc - 1) / rpc, N), dim3(256), 0, s, reinterpret_cast<const float4*>(da), ws + T.scsh[l],
                                                 reinterpret_cast<const float4*>(ws + T.u[l]), ws + T.meaninv[l], B, N, cout,
                                                 T.cmax, st, rpc);
    else
      mapf_launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(256), 0, s, da, ws + T.scsh[l], ws + T.u[l], ws + T.meaninv[l], B, N, cout, T.cmax, st);
    const int64_t planes = rows * cout;
    bool bn_grads_done = false;
    if ((cout * kFovCells) % 4 == 0 && N * cout <= kBnFusedMax) {
      mapf_launch_pdl(bn_bwd_coef_apply4_kernel, dim3(unsigned(mapf_min64((planes * kFovCells / 4 + 255) / 256, 148 * 16))), dim3(256), 0, s,
          reinterpret_cast<float4*>(da), ws + T.scsh[l], reinterpret_cast<const float4*>(ws + T.u[l]),
          ws + T.meaninv[l], st, p->bn_gamma[l], N, cout, T.cmax, count, rows, G.bn_gamma[l], G.bn_beta[l]);
      bn_grads_done = true;
    } else {
      mapf_launch_pdl(bn_bwd_coef_kernel, dim3((N * cout + 127) / 128), dim3(128), 0, s, st, ws + T.meaninv[l], p->bn_gamma[l], N, cout, T.cmax, count,
                                                               ws + T.coef[l]);
      if ((cout * kFovCells) % 4 == 0)
        mapf_launch_pdl(bn_bwd_apply4_kernel, dim3(unsigned(mapf_min64((planes * kFovCells / 4 + 255) / 256, 148 * 16))), dim3(256), 0, s, 
            reinterpret_cast<float4*>(da), ws + T.scsh[l], reinterpret_cast<const float4*>(ws + T.u[l]),
            ws + T.meaninv[l], ws + T.coef[l], N, cout, rows);
      else
        mapf_launch_pdl(bn_bwd_apply_kernel, dim3(unsigned(mapf_min64((planes * kFovCells + 255) / 256, 148 * 16))), dim3(256), 0, s, 
            da, ws + T.scsh[l], ws + T.u[l], ws + T.meaninv[l], ws + T.coef[l], N, cout, planes);
    }
    // weight gradients of consecutive layers on alternating lanes: each starts when its du is final instead of behind the
    // layer above (a batch of 128 fills 80 CTAs per layer)
    cudaStream_t swl = ((L - 1 - l) & 1) ? sp : sw;
    ord.wait(swl, ord.mark(s));       // du_l (in `da`) and the BatchNorm sums
    if (!bn_grads_done)
      mapf_launch_pdl(bn_param_grads_kernel, dim3((cout + 63) / 64), dim3(64), 0, swl, st, N, cout, T.cmax, G.bn_gamma[l], G.bn_beta[l]);
    {
      const int tile_rows = (rows + kCwRows - 1) / kCwRows >= 4 * 148 ? kCwRows : kCwRowsSmall;
      // (at least the [rstep][npairs][10] reduction scratch of the kernel's small-layer path: <= 256 x 10 floats)
      const size_t smem = std::max(sizeof(float) * size_t(tile_rows) * (cin + cout) * kFovCells, sizeof(float) * size_t(kCwThreads) * 10);
      MAPF_REQUIRE(smem <= 220 * 1024, MAPF_ERR_UNSUPPORTED);
      MAPF_REQUIRE(cin * cout <= 4 * kCwThreads, MAPF_ERR_UNSUPPORTED);
      const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
      const int gridw = int(mapf_min64(ntiles, 2 * 148));
      const float* lin = l == 0 ? a->states : ws + T.a[l - 1];
      int64_t maxp = 0;
      for (int q = 0; q < L; ++q) maxp = max(maxp, int64_t(p->channels[q]) * p->channels[q + 1]);
      float* part = ws + T.dwpart + (((L - 1 - l) & 1) ? mapf_min64((rows + 7) / 8, 2 * 148) * (maxp * 9 + T.cmax) : 0);
      if (cin * cout <= kCwThreads) {
        if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(conv_bwd_weight_kernel<1>, size_t(smem), cache__); }())
          return MAPF_ERR_CUDA;
        mapf_launch_pdl(conv_bwd_weight_kernel<1>, dim3(gridw), dim3(kCwThreads), smem, swl, da, lin, cin, cout, rows, part, tile_rows);
      } else {
        if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(conv_bwd_weight_kernel<4>, size_t(smem), cache__); }())
          return MAPF_ERR_CUDA;
        mapf_launch_pdl(conv_bwd_weight_kernel<4>, dim3(gridw), dim3(kCwThreads), smem, swl, da, lin, cin, cout, rows, part, tile_rows);
      }
      const int nw = cin * cout * 9;
      mapf_launch_pdl(conv_bwd_weight_reduce_kernel, dim3((nw + cout + 31) / 32), dim3(256), 0, swl, part, gridw, nw, cout, G.conv_w[l], G.conv_b[l]);
    }
    const int m_dw = ord.mark(swl);
    if (l > 0) {
      ord.wait(s, m_dw_prev2);
      ord.wait(s, m_convpk[l]);
      // data gradient = the same 3x3 convolution with mirrored, transposed filters
      if (conv_tc && mapf_conv_tc_layer_ok(cin, cout, 1)) {
        MAPF_TRY(mapf_conv_tc_layer(da, p->conv_w[l], nullptr, cin, cout, 1, B, N, ws + T.convtc_bwd[l], 1, da_next, nullptr, 0, nullptr,
                                    stream));
      } else {
        MAPF_TRY(launch_conv_train(s, da, cout, cin, B, N, ws + T.convpk_bwd[l], nullptr, da_next, nullptr, 0));
      }
      float* t = da; da = da_next; da_next = da_spare; da_spare = t;
    }
    m_dw_prev2 = m_dw_prev;
    m_dw_prev = m_dw;
  }
  ord.wait(s, ord.mark(sw));          // join both lanes
  ord.wait(s, ord.mark(sp));
  MAPF_REQUIRE(ord.ok, MAPF_ERR_CUDA);
  return mapf_launch_status();
}

extern "C" int mapf_ce_loss(const mapf_ce_loss_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->logits && a->targets && a->loss && a->dlogits, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0 && a->num_actions > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->rows <= 2 * kOneBlockThreads) {      // (one block walks 5 120 rows in 9.6 us, three memset-ed blocks' worth in 4)
    mapf_launch_pdl(ce_loss_one_block_kernel, dim3(1), dim3(kOneBlockThreads), 0, s, *a);
    return mapf_launch_status();
  }
  if (cudaMemsetAsync(a->loss, 0, sizeof(float), s) != cudaSuccess) return MAPF_ERR_CUDA;
  mapf_launch_pdl(ce_loss_kernel, dim3(unsigned((a->rows + 255) / 256)), dim3(256), 0, s, *a);
  return mapf_launch_status();
}

extern "C" int mapf_clip_adam(const mapf_clip_adam_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->grad_norm && a->scratch, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->n > 0 && (a->step >= 1 || a->step_dev != nullptr), MAPF_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = int(mapf_min64((a->n + 255) / 256, 148));
  if (a->n <= (int64_t(1) << 18)) {
    mapf_launch_pdl(sumsq_one_block_kernel, dim3(1), dim3(kOneBlockThreads), 0, s, a->grads, a->n, static_cast<double*>(a->scratch),
                    static_cast<int32_t*>(a->step_dev));
  } else {
    if (cudaMemsetAsync(a->scratch, 0, sizeof(double), s) != cudaSuccess) return MAPF_ERR_CUDA;
    mapf_launch_pdl(sumsq_kernel, dim3(blocks), dim3(256), 0, s, a->grads, a->n, a->scratch, a->step_dev);
  }
  const double bc1 = 1.0 - pow(double(a->beta1), double(a->step));
  const double bc2 = 1.0 - pow(double(a->beta2), double(a->step));
  mapf_launch_pdl(clip_adam_kernel, dim3(blocks), dim3(256), 0, s, *a, float(bc1), float(sqrt(bc2)),
                  static_cast<int32_t*>(nullptr));
  return mapf_launch_status();
}

extern "C" int mapf_peer_reduce_clip_adam(const mapf_peer_reduce_args* r, const mapf_clip_adam_args* a,
                                          mapf_stream_t stream) {
  MAPF_REQUIRE(r && a && a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->grad_norm && a->scratch, MAPF_ERR_NULL);
  MAPF_REQUIRE(r->world >= 1 && r->world <= 8 && r->rank >= 0 && r->rank < r->world, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->n > 0 && (a->n & 3) == 0 && (a->step >= 1 || a->step_dev != nullptr), MAPF_ERR_SHAPE);
  PeerReduceArgs k{};
  k.world = r->world; k.rank = r->rank; k.n = a->n;
  for (int i = 0; i < r->world; ++i) {
    MAPF_REQUIRE(r->peer_grads[i] && r->peer_flags[i], MAPF_ERR_NULL);
    MAPF_REQUIRE(mapf_aligned(r->peer_grads[i], 16), MAPF_ERR_ALIGN);
    k.grads[i] = r->peer_grads[i];
    k.flags[i] = r->peer_flags[i];
  }
  MAPF_REQUIRE(mapf_aligned(a->grads, 16), MAPF_ERR_ALIGN);
  k.out = a->grads; k.sumsq = a->scratch; k.step_dev = a->step_dev;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(a->scratch, 0, sizeof(double), s) != cudaSuccess) return MAPF_ERR_CUDA;
  const int blocks = int(mapf_min64((a->n / 4 + 255) / 256, 148));
  mapf_launch_pdl(peer_reduce_sumsq_kernel, dim3(blocks), dim3(256), 0, s, k);
  const double bc1 = 1.0 - pow(double(a->beta1), double(a->step));
  const double bc2 = 1.0 - pow(double(a->beta2), double(a->step));
  const int ablocks = int(mapf_min64((a->n + 255) / 256, 148));
  mapf_launch_pdl(clip_adam_kernel, dim3(ablocks), dim3(256), 0, s, *a, float(bc1), float(sqrt(bc2)),
                  r->peer_flags[r->rank] + 8);
  return mapf_launch_status();
}

// Peer-mapped device buffers (CUDA IPC): the only entry points of the library that allocate.
extern "C" int mapf_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle) {
  MAPF_REQUIRE(ptr && handle, MAPF_ERR_NULL);
  MAPF_REQUIRE(bytes > 0, MAPF_ERR_SHAPE);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, size_t(bytes)) != cudaSuccess) { (void)cudaGetLastError(); return MAPF_ERR_CUDA; }
  if (cudaMemset(p, 0, size_t(bytes)) != cudaSuccess) { (void)cudaGetLastError(); cudaFree(p); return MAPF_ERR_CUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { (void)cudaGetLastError(); cudaFree(p); return MAPF_ERR_CUDA; }
  memcpy(handle, &h, 64);
  *ptr = p;
  return MAPF_OK;
}
extern "C" int mapf_peer_open(const uint8_t* handle, void** ptr) {
  MAPF_REQUIRE(ptr && handle, MAPF_ERR_NULL);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); return MAPF_ERR_CUDA; }
  return MAPF_OK;
}
extern "C" int mapf_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) { (void)cudaGetLastError(); return MAPF_ERR_CUDA; }
  return MAPF_OK;
}
extern "C" int mapf_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) { (void)cudaGetLastError(); return MAPF_ERR_CUDA; }
  return MAPF_OK;
}

extern "C" int64_t mapf_graph_filter_bwd_workspace_size(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim,
                                                        int32_t taps, int32_t has_skip) {
  const int64_t rows = int64_t(batch) * agents, KT = int64_t(taps + (has_skip ? 1 : 0)) * in_dim;
  return r4(KT * out_dim) + 2 * r4(rows * KT) + r4(rows * out_dim);
}

extern "C" int mapf_graph_filter_bwd(const mapf_graph_filter_bwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->w && a->dy && a->workspace && a->dx && a->dw, MAPF_ERR_NULL);
  MAPF_REQUIRE((a->w2 == nullptr) == (a->dw2 == nullptr), MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0 && a->in_dim > 0 && a->out_dim > 0 && a->taps > 0, MAPF_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = a->batch, N = a->agents, F = a->in_dim, G = a->out_dim, K = a->taps;
  const int skip = a->w2 ? 1 : 0, KF = K * F, KT = (K + skip) * F;
  const int64_t rows = int64_t(B) * N;
  float* wt = a->workspace;
  float* z = wt + r4(int64_t(KT) * G);
  float* dz = z + r4(rows * KT);
  float* dyr = dz + r4(rows * KT);
  // recompute the filter taps z with the forward kernel
  MAPF_TRY(mapf_pack_graph_weights(a->w, a->w2, F, K, G, wt, stream));
  mapf_launch_pdl(bcn_to_rows_kernel, dim3(unsigned(mapf_min64((rows * G + 255) / 256, 148 * 16))), dim3(256), 0, s, a->dy, B, G, N, dyr);
  mapf_graph_filter_fwd_args g{};
  g.batch = B; g.agents = N; g.in_dim = F; g.out_dim = G; g.taps = K;
  g.add_self_loop = a->add_self_loop; g.has_skip = skip; g.relu = 0; g.num_actions = 0;
  g.x = a->x; g.x_sb = int64_t(F) * N; g.x_sn = 1; g.x_sf = N;
  g.gso = a->gso; g.wt = wt;
  g.y = nullptr;   // only the saved filter taps z are needed
  g.z = z;
  MAPF_TRY(mapf_graph_filter_fwd(&g, stream));
  if (cudaMemsetAsync(a->dw, 0, sizeof(float) * size_t(KF) * G, s) != cudaSuccess) return MAPF_ERR_CUDA;
  MAPF_TRY(gemm(s, G, KF, int(rows), dyr, 1, G, z, KT, 1, a->dw, KF, nullptr, 0, nullptr, 0,
                splits_for(rows, ((G + 127) / 128) * ((KF + 63) / 64))));
  MAPF_TRY(gemm(s, int(rows), KF, G, dyr, G, 1, a->w, KF, 1, dz, KT));
  if (skip) {
    if (cudaMemsetAsync(a->dw2, 0, sizeof(float) * size_t(F) * G, s) != cudaSuccess) return MAPF_ERR_CUDA;
    MAPF_TRY(gemm(s, G, F, int(rows), dyr, 1, G, z + KF, KT, 1, a->dw2, F, nullptr, 0, nullptr, 0,
                  splits_for(rows, ((G + 127) / 128) * ((F + 63) / 64))));
    MAPF_TRY(gemm(s, int(rows), F, G, dyr, G, 1, a->w2, F, 1, dz + KF, KT));
  }
  const size_t smem = sizeof(float) * (size_t(N) * N + ((N + 3) & ~3) + size_t(N) * KT);
  MAPF_REQUIRE(smem <= 220 * 1024, MAPF_ERR_UNSUPPORTED);
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(recurrence_bwd_kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  mapf_launch_pdl(recurrence_bwd_kernel, dim3(B), dim3(128), smem, s, dz, a->gso, nullptr, N, F, K, skip, a->add_self_loop, a->dx, int64_t(F) * N,
                                             1, N);
  return mapf_launch_status();
}
