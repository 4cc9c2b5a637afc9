// fp32-accurate tensor-core GEMM (3xTF32 on tcgen05, accumulators in TMEM):
//   D[M,N] = A[M,K] . B[N,K]^T        A row-major fp32 (K contiguous), B pre-split / pre-tiled (mapf_tc_pack_b)
// followed by a fused epilogue:
//   EPI_STORE : C = act(D + bias)                                   (encoder FC, compressMLP framework_gnn.py:93)
//   EPI_HEAD  : y = relu(D); logits = y Wa^T + ba; first-max argmax (graph-filter mix gnn.py:117-118 + ReLU +
//               actionMLP framework_gnn.py:171-181 + np.argmax evaluate.py:238)
//
// Warp-specialised, one 256-thread CTA per 128-row tile, six shared-memory stages of KC floats of K:
//   warps 0-5  producers : warp w owns stage w and the chunks c = w, w+6, ...: its lanes load the A chunk (fp32,
//                          global, all loads of the chunk in flight at once), split it into TF32 hi / lo
//                          (round-to-nearest) and store both in the canonical no-swizzle K-major UMMA layout; lane 0
//                          also issues the TMA bulk copy (cp.async.bulk, SASS UBLKCP) of the pre-tiled hi|lo B chunk;
//                          the warp arrives on the stage's `full` mbarrier (the bulk copy through its byte count)
//   warps 6,7  MMA issuers: one elected lane each (even / odd chunks, each with its own TMEM accumulator) waits
//                          `full`, issues 3 tcgen05.mma (128 x N x 8) per 8-wide K slice (lo*hi, hi*lo, hi*hi) and
//                          commits them to the stage's `empty` mbarrier.  Two issuers because the MMAs are small
//                          (32-64 cycles of tensor work): one thread's wait/issue/commit loop (~500 cycles) would pace
//                          the whole pipeline
//   all warps  epilogue  : after the final commit warp w reads TMEM lanes 32(w%4).. (= 32 output rows, one per thread)
//                          and column half w/4 with tcgen05.ld; EPI_STORE transposes through shared memory so that
//                          global stores are 128-byte coalesced; EPI_HEAD needs one exchange between the two halves
// The tensor core's fp32 accumulation loses a little precision per accumulate, so consecutive chunks rotate over
// NACC independent TMEM accumulators that are summed in registers in the epilogue.
#include <cuda_fp16.h>
#include <cstdio>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcProdWarps = 6;      // warps 0-5, one shared-memory stage each
constexpr int kTcM = 128;

enum { EPI_STORE = 0, EPI_HEAD = 1 };

struct TcGemmArgs {
  int M, N, K;
  const float* A; int64_t lda;
  const float* Bp;     // pre-tiled B: [chunk][hi | lo][NT/8][KC/4][8][4]
  int64_t bp_tile;     // EPI_STORE with N > NT: floats between the packs of consecutive NT-column tiles (grid.y), else 0
  // EPI_STORE
  float* C; int64_t ldc;
  const float* bias;
  int relu;
  // EPI_HEAD
  const float* act_w;  // [A, N]
  const float* act_b;  // [A]
  int num_actions;
  float* logits;       // [M, A]
  int32_t* actions;    // [M] or NULL
  float* y;            // optional [M, N] post-ReLU features or NULL
  // fp16-split presplit GEMM
  const float* row_scale;   // [M] inverse scale of each A row
  const float* b_hdr;       // [0] inverse scale of B
};

// B [N,K0] (+ [N,K1] appended along K) row-major fp32 -> chunked canonical hi / lo blocks
__global__ void tc_pack_b_kernel(const float* __restrict__ b0, int K0, const float* __restrict__ b1, int K1, int N,
                                 int NT, int KC, float* __restrict__ out) {
  const int K = K0 + K1;
  const int nchunks = (K + KC - 1) / KC;
  const int64_t per_chunk = int64_t(2) * NT * KC;
  const int64_t total = per_chunk * nchunks;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx / per_chunk);
    int64_t r = idx - c * per_chunk;
    const int is_lo = r >= int64_t(NT) * KC;
    if (is_lo) r -= int64_t(NT) * KC;
    // canonical: [(row/8)][(k/4)][row%8][k%4]
    const int k4i = int(r & 3), r8 = int((r >> 2) & 7);
    const int rest = int(r >> 5);
    const int kq = rest % (KC >> 2), rg = rest / (KC >> 2);
    const int row = rg * 8 + r8, k = c * KC + kq * 4 + k4i;
    float v = 0.0f;
    if (row < N && k < K) v = k < K0 ? b0[int64_t(row) * K0 + k] : b1[int64_t(row) * K1 + (k - K0)];
    float hi, lo;
    tc::split1(v, hi, lo);
    out[idx] = is_lo ? lo : hi;
  }
}

// fp16-split, scaled, pixel-major B operand (see mapf_tc_pack_b_f16_pixel_major); one block
__global__ void tc_pack_b_f16_kernel(const float* __restrict__ b, int C, int N, int NT, float* __restrict__ out) {
  __shared__ float s_red[1024];
  const int K = C * kFovCells;
  float m = 0.0f;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) m = fmaxf(m, fabsf(b[i]));
  s_red[threadIdx.x] = m;
  __syncthreads();
  for (int st = blockDim.x / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + st]);
    __syncthreads();
  }
  int eb = int(__float_as_uint(s_red[0]) >> 23) & 0xFF;
  eb = eb < 32 ? 32 : (eb > 240 ? 240 : eb);
  const float scale = __uint_as_float(uint32_t(263 - eb) << 23);
  if (threadIdx.x < 4) out[threadIdx.x] = threadIdx.x == 0 ? __uint_as_float(uint32_t(eb - 9) << 23) : 0.0f;
  __half* o = reinterpret_cast<__half*>(out + 4);
  const int nch = (K + 15) >> 4, per_chunk = 2 * NT * 16;
  for (int idx = threadIdx.x; idx < nch * per_chunk; idx += blockDim.x) {
    const int c = idx / per_chunk;
    int r = idx - c * per_chunk;
    const int is_lo = r >= NT * 16;
    if (is_lo) r -= NT * 16;
    const int k8i = r & 7, n8 = (r >> 3) & 7, kq = (r >> 6) & 1, ng = r >> 7;
    const int row = ng * 8 + n8, k = c * 16 + kq * 8 + k8i;
    float v = 0.0f;
    if (row < N && k < K) v = b[int64_t(row) * K + (k % C) * kFovCells + k / C] * scale;
    const __half hi = __float2half_rn(v);
    o[idx] = is_lo ? __float2half_rn(v - __half2float(hi)) : hi;
  }
}

// Epilogue store of a warp-private 32 x 32 block (lane = output row, v[j] = column j) to row-major global memory through the
// warp's transpose scratch t (kTcTP floats).  vec (rows 16-byte aligned, N % 4 == 0): 8 STS.128 + 8 LDS.128 + 8 STG.128, every
// store instruction writing four 128-byte row segments; otherwise 32 STS + 32 (LDS + STG.32), one row segment each.
// (The 32-bit form cost 1.7 us per block -- 2 to 6.6 us of a 640-row GEMM's 11 to 15 us.)
constexpr int kTcTP = 32 * 36;
__device__ __forceinline__ void tc_store_block32(float* t, const float (&v)[32], float* __restrict__ C, int64_t ldc, int row0,
                                                 int col0, int M, int N, bool vec, int lane) {
  if (vec) {
    float4* t4 = reinterpret_cast<float4*>(t + lane * 36);
#pragma unroll
    for (int j = 0; j < 8; ++j) t4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int rr = lane >> 3, c4 = (lane & 7) * 4;
    float* dst = C + int64_t(row0 + rr) * ldc + col0 + c4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float4 o = *reinterpret_cast<const float4*>(t + (it * 4 + rr) * 36 + c4);
      if (row0 + it * 4 + rr < M && col0 + c4 < N) *reinterpret_cast<float4*>(dst + int64_t(it * 4) * ldc) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[lane * 33 + j] = v[j];
    __syncwarp();
    const int n = col0 + lane;
#pragma unroll 4
    for (int r = 0; r < 32; ++r)
      if (row0 + r < M && n < N) C[int64_t(row0 + r) * ldc + n] = t[r * 33 + lane];
  }
  __syncwarp();
}

// NP = producer warps = shared-memory stages (6: two CTAs per SM; 4 with NACC = 1: three CTAs per SM -- the variant the
// training step uses, whose 320 row tiles at B = 8192 then fit one wave instead of 296 + 24); NACC = MMA issuers.
// NTL = column tiles of NT per CTA (EPI_STORE, NACC = 1): the A chunk is loaded and split ONCE and multiplied with NTL
// B tiles into NTL accumulators -- wide outputs (the training step's da = dx W_fc, 400 columns) re-read A per pair of
// tiles instead of per tile.
template <int NT, int KC, int NACC, int EPI, int NP = kTcProdWarps, int NTL = 1>
__global__ void __launch_bounds__(kTcThreads) tc_gemm_kernel(TcGemmArgs g) {
  static_assert(NTL == 1 || (NACC == 1 && EPI == EPI_STORE), "several column tiles per CTA: store epilogue, one accumulator each");
  constexpr int NSTAGE = NP;
  constexpr uint32_t kColsNeeded = NT * NACC * NTL;
  constexpr uint32_t kCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128
                             : kColsNeeded <= 256 ? 256 : 512;
  constexpr int kABytes = kTcM * KC * 4, kBBytes = NT * KC * 4;
  constexpr int kStageBytes = 2 * kABytes + NTL * 2 * kBBytes;
  constexpr int KQ = KC / 4;
  extern __shared__ __align__(128) uint8_t s_tc[];
  __shared__ __align__(8) uint64_t s_full[NSTAGE], s_empty[NSTAGE], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_head[EPI == EPI_HEAD ? kMaxActions * NT + kMaxActions : NT * NTL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef MAPF_TCG_STAMP
  __shared__ unsigned long long s_ts[16], s_tw[8][4];
#define TCG_WSTAMP(i) do { if (lane == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); s_tw[warp][i] = t__; } } while (0)
#define TCG_STAMP(i) do { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); s_ts[i] = t__; } while (0)
#define TCG_NOW(v) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v))
  unsigned long long tq0 = 0, tq1 = 0, tq2 = 0, acc_a = 0, acc_b = 0, acc_c = 0;
  if (tid == 0) TCG_STAMP(0);
#else
#define TCG_STAMP(i) do {} while (0)
#define TCG_NOW(v) do {} while (0)
#define TCG_WSTAMP(i) do {} while (0)
#endif
  const int m0 = blockIdx.x * kTcM, n0 = blockIdx.y * NT * NTL;
  // column tiles this CTA really has (the last CTA of a row of tiles may own fewer than NTL)
  const int my_tiles = NTL == 1 ? 1 : min(NTL, (g.N - n0 + NT - 1) / NT);

  if (warp == 0) tc::tmem_alloc(&s_tmem, kCols);
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      tc::mbar_init(&s_full[s], 32);
      tc::mbar_init(&s_empty[s], 1);
    }
    tc::mbar_init(&s_done, NACC);
    tc::fence_mbar_init();
  }
  // programmatic dependent launch: the CTA is resident, its TMEM columns and barriers set up, before the kernel in front
  // has drained; every global read (A, the packed B, bias / head weights) stays behind this wait
  mapf_pdl_wait();
  mapf_pdl_launch();
  if (EPI == EPI_HEAD) {
    for (int i = tid; i < g.num_actions * NT; i += kTcThreads) {
      const int q = i / NT, n = i - q * NT;
      s_head[i] = n < g.N ? __ldg(g.act_w + int64_t(q) * g.N + n) : 0.0f;
    }
    if (tid < g.num_actions) s_head[kMaxActions * NT + tid] = __ldg(g.act_b + tid);
  } else {
    for (int i = tid; i < NT * NTL; i += kTcThreads) s_head[i] = (g.bias != nullptr && n0 + i < g.N) ? __ldg(g.bias + n0 + i) : 0.0f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (tid == 0) TCG_STAMP(1);
  const int nchunks = (g.K + KC - 1) / KC;
  // warp-uniform copies for the MMA issuers: the whole issuing warp runs its (warp-uniform) branch and elect.sync picks the
  // lane, so every MMA operand is a uniform-register value and ptxas emits plain UTCHMMAs instead of wrapping each one in
  // an ELECT / R2UR.BROADCAST loop (~100 clocks per MMA behind a divergent `lane == 0` branch)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

  if (warp_u < NP) {
    // ===== producers: warp w owns stage w =====
    const int s = warp;
    uint8_t* a_hi = s_tc + s * kStageBytes;
    uint8_t* a_lo = a_hi + kABytes;
    uint8_t* b_hi = a_lo + kABytes;
    constexpr int kPerLane = kTcM * KQ / 32;
    // kPerLane independent 128-bit loads of one chunk in flight per lane
    auto load_chunk = [&](int c, float4* v) {
#pragma unroll
      for (int it = 0; it < kPerLane; ++it) {
        const int idx = lane + it * 32;
        const int r8 = idx & 7, k4 = (idx >> 3) % KQ, rg = idx / (8 * KQ);
        const int grow = m0 + rg * 8 + r8, gk = c * KC + k4 * 4;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nchunks && grow < g.M) {
          const float* p = g.A + int64_t(grow) * g.lda;
          if (gk + 3 < g.K) {
            v[it] = mapf_ld_dep(reinterpret_cast<const float4*>(p + gk));
          } else {
            if (gk + 0 < g.K) v[it].x = mapf_ld_dep(p + gk + 0);
            if (gk + 1 < g.K) v[it].y = mapf_ld_dep(p + gk + 1);
            if (gk + 2 < g.K) v[it].z = mapf_ld_dep(p + gk + 2);
          }
        }
      }
    };
    float4 v[kPerLane], vn[kPerLane];
    load_chunk(warp, v);
    uint32_t ph = 0;
    for (int c = warp; c < nchunks; c += NP, ph ^= 1) {
      load_chunk(c + NP, vn);   // next chunk's loads overlap this chunk's wait / convert / store
      TCG_NOW(tq0);
      tc::mbar_wait(&s_empty[s], ph ^ 1);
      TCG_NOW(tq1);
      if (lane == 0) {
        tc::mbar_expect_tx(&s_full[s], uint32_t(my_tiles) * 2 * kBBytes);
        for (int t = 0; t < my_tiles; ++t)
          tc::bulk_g2s(b_hi + t * 2 * kBBytes, g.Bp + (int64_t(blockIdx.y) * NTL + t) * g.bp_tile + int64_t(c) * (2 * NT * KC),
                       2 * kBBytes, &s_full[s]);
      }
#pragma unroll
      for (int it = 0; it < kPerLane; ++it) {
        const int idx = lane + it * 32;
        const int r8 = idx & 7, k4 = (idx >> 3) % KQ, rg = idx / (8 * KQ);
        float4 h, l;
        tc::split4(v[it], h, l);
        const uint32_t off = tc::canon_off(rg * 8 + r8, k4 * 4, KC);
        *reinterpret_cast<float4*>(a_hi + off) = h;
        *reinterpret_cast<float4*>(a_lo + off) = l;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&s_full[s]);
      if (tid == 0 && c == 0) TCG_STAMP(2);
      TCG_NOW(tq2);
#pragma unroll
      for (int it = 0; it < kPerLane; ++it) v[it] = vn[it];
#ifdef MAPF_TCG_STAMP
      {
        float sink = 0.f;
#pragma unroll
        for (int it = 0; it < kPerLane; ++it) sink += v[it].x;
        if (sink == 123.456f) s_ts[15] = 1;      // (forces the loads to have landed before the next stamp)
        unsigned long long t3; TCG_NOW(t3);
        acc_a += tq1 - tq0; acc_b += tq2 - tq1; acc_c += t3 - tq2;
        if (tid == 0) { s_ts[10] = acc_a; s_ts[11] = acc_b; s_ts[12] = acc_c; }
      }
#endif
    }
  } else if (warp_u < NP + NACC) {
    // ===== MMA issuers (NACC warps behind the producers: chunks c = w, w + NACC, ... into accumulator w) =====
    // A single thread issues every tcgen05.mma, so its instruction stream is the critical path: descriptors are
    // built once (stage 0) and advanced by adding the stage / K-slice byte offset to the 14-bit address field.
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_tf32(kTcM, NT);
      constexpr uint32_t sbo = KQ * 128, lbo = 128;
      const uint32_t base = tc::smem_u32(s_tc);
      const uint64_t dah0 = tc::smem_desc(base, lbo, sbo);
      const uint64_t dal0 = tc::smem_desc(base + kABytes, lbo, sbo);
      const uint64_t dbh0 = tc::smem_desc(base + 2 * kABytes, lbo, sbo);
      const uint64_t dbl0 = tc::smem_desc(base + 2 * kABytes + kBBytes, lbo, sbo);
      static_assert(NACC == 1 || (NACC == 2 && (NSTAGE % 2) == 0), "issuer w owns accumulator w and the stages of its parity");
      const int which = warp_u - NP;
      int s = which;
      uint32_t ph = 0;
      const uint32_t accsel = which;
      for (int c = which; c < nchunks; c += NACC) {
        TCG_NOW(tq0);
        tc::mbar_wait(&s_full[s], ph);
        tc::fence_after_sync();
        TCG_NOW(tq1);
        if (c == 0) TCG_STAMP(3);
        const uint64_t soff = uint64_t(uint32_t(s) * (kStageBytes >> 4));
        const bool first = c < NACC;   // first touch of this accumulator overwrites
#pragma unroll
        for (int t = 0; t < NTL; ++t) {
          if (t < my_tiles) {
            const uint32_t acc = tmem_u + (accsel * NTL + t) * NT;
            const uint64_t bo = uint64_t(t * ((2 * kBBytes) >> 4));
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
              const uint64_t o = soff + uint64_t(ks * 16);   // one 8-wide K slice = 2 core matrices = 256 bytes
              tc::mma_tf32(acc, dal0 + o, dbh0 + o + bo, idesc, !(first && ks == 0));   // small terms first
              tc::mma_tf32(acc, dah0 + o, dbl0 + o + bo, idesc, true);
              tc::mma_tf32(acc, dah0 + o, dbh0 + o + bo, idesc, true);
            }
          }
        }
        tc::mma_commit(&s_empty[s]);
#ifdef MAPF_TCG_STAMP
        TCG_NOW(tq2);
        acc_a += tq1 - tq0; acc_b += tq2 - tq1;
#endif
        s += NACC;
        if (s >= NSTAGE) { s -= NSTAGE; ph ^= 1; }
      }
      tc::mma_commit(&s_done);
      if (which == 0) TCG_STAMP(4);
#ifdef MAPF_TCG_STAMP
      if (which == 0) { s_ts[8] = acc_a; s_ts[9] = acc_b; }
#endif
    }
    __syncwarp();   // lanes 1-31 park here instead of spinning on s_done next to the issuing lane
  }

  {
    // ===== epilogue (all 8 warps): warp w owns TMEM lanes [32(w%4), +32) = rows m0 + 32(w%4) + lane and the column
    // half w/4.  The operand stages are free once s_done fires and are reused as transpose / exchange scratch. =====
    tc::mbar_wait(&s_done, 0);
    tc::fence_after_sync();
    if (tid == 0) TCG_STAMP(5);
    TCG_WSTAMP(0);
    const int quad = warp & 3, half = warp >> 2;
    const int row = m0 + quad * 32 + lane;
    const int nacc_used = nchunks < NACC ? nchunks : NACC;
    constexpr int kHalfCols = NT / 2;
    float* s_scr = reinterpret_cast<float*>(s_tc);
    // 128-bit epilogue stores when every output row is 16-byte aligned
    const bool vec_c = EPI == EPI_STORE ? ((g.ldc & 3) == 0 && (g.N & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0)
                                        : ((g.N & 3) == 0 && (reinterpret_cast<uintptr_t>(g.y) & 15) == 0);
    float part[kMaxActions];
#pragma unroll
    for (int q = 0; q < kMaxActions; ++q) part[q] = 0.0f;
#pragma unroll
    for (int tl = 0; tl < NTL; ++tl) {
    if (tl >= my_tiles) break;
#pragma unroll
    for (int cbl = 0; cbl < kHalfCols; cbl += 32) {
      const int cb = tl * NT + half * kHalfCols + cbl;      // column inside this CTA's NT * NTL (TMEM column with NACC = 1)
      uint32_t raw[NACC][32];
#pragma unroll
      for (int a = 0; a < NACC; ++a)
        if (a < nacc_used) tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + a * NT * NTL + cb, raw[a]);   // warp-uniform
      tc::tmem_ld_wait();
      if (tl == 0 && cbl == 0) TCG_WSTAMP(1);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[0][j]);
#pragma unroll
      for (int a = 1; a < NACC; ++a)
        if (a < nacc_used) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(raw[a][j]);
        }
      if (EPI == EPI_STORE) {
        // bias / ReLU, then a warp-private 32x32 transpose through shared memory so that every global store
        // instruction writes 128 contiguous bytes of one output row
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += s_head[cb + j];
          if (g.relu) v[j] = fmaxf(v[j], 0.0f);
        }
        if (tl == 0 && cbl == 0) TCG_WSTAMP(2);
        tc_store_block32(s_scr + warp * kTcTP, v, g.C, g.ldc, m0 + quad * 32, n0 + cb, g.M, g.N, vec_c, lane);
        if (tl == 0 && cbl == 0) TCG_WSTAMP(3);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        if (g.y != nullptr) tc_store_block32(s_scr + warp * kTcTP, v, g.y, g.N, m0 + quad * 32, cb, g.M, g.N, vec_c, lane);
#pragma unroll
        for (int q = 0; q < kMaxActions; ++q) {
          if (q < g.num_actions) {
            const float4* wq = reinterpret_cast<const float4*>(s_head + q * NT + cb);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w4 = wq[j >> 2];
              part[q] = fmaf(v[j], w4.x, part[q]);
              part[q] = fmaf(v[j + 1], w4.y, part[q]);
              part[q] = fmaf(v[j + 2], w4.z, part[q]);
              part[q] = fmaf(v[j + 3], w4.w, part[q]);
            }
          }
        }
      }
    }
    }
    if (EPI == EPI_HEAD) {
      // the two column halves of a row meet in shared memory (after the y transposes are done with it)
      float* ex = s_scr + 8 * kTcTP;   // [128 rows][kMaxActions]
      if (half == 1) {
#pragma unroll
        for (int q = 0; q < kMaxActions; ++q) ex[(quad * 32 + lane) * kMaxActions + q] = part[q];
      }
      __syncthreads();
      if (half == 0 && row < g.M) {
        float best = 0.0f;
        int arg = 0;
#pragma unroll
        for (int q = 0; q < kMaxActions; ++q) {
          if (q < g.num_actions) {
            const float o = part[q] + ex[(quad * 32 + lane) * kMaxActions + q] + s_head[kMaxActions * NT + q];
            g.logits[int64_t(row) * g.num_actions + q] = o;
            if (q == 0 || o > best) { best = o; arg = q; }   // first max wins (np.argmax)
          }
        }
        if (g.actions != nullptr) g.actions[row] = arg;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kCols);
#ifdef MAPF_TCG_STAMP
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    TCG_STAMP(6);
    printf("tcg<%d,%d,%d> K=%d: setup %llu first-full %llu first-mma-wake %llu last-commit %llu done-wake %llu end %llu ns\n", NT, NP, NTL, g.K,
           s_ts[1] - s_ts[0], s_ts[2] - s_ts[0], s_ts[3] - s_ts[0], s_ts[4] - s_ts[0], s_ts[5] - s_ts[0], s_ts[6] - s_ts[0]);
    printf("   K=%d issuer 0: waiting %llu issuing %llu ns | producer 0: empty-wait %llu convert+store %llu load-wait %llu ns\n", g.K,
           s_ts[8], s_ts[9], s_ts[10], s_ts[11], s_ts[12]);
    if (EPI == EPI_STORE)
      for (int w = 0; w < 8; ++w)
        printf("   K=%d warp %d: wake %llu ld %llu sts %llu stores %llu\n", g.K, w, s_tw[w][0] - s_ts[0], s_tw[w][1] - s_ts[0], s_tw[w][2] - s_ts[0], s_tw[w][3] - s_ts[0]);
  }
#endif
}

template <int NT, int KC, int NACC, int EPI, int NP = kTcProdWarps, int NTL = 1>
int launch_tc(const TcGemmArgs& g, cudaStream_t s) {
  constexpr size_t smem = size_t(NP) * (2 * kTcM * KC * 4 + NTL * 2 * NT * KC * 4);
  static_assert(smem >= 8 * kTcTP * 4 + 128 * kMaxActions * 4, "the epilogue transposes through the operand stages");
  auto kernel = tc_gemm_kernel<NT, KC, NACC, EPI, NP, NTL>;
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  dim3 grid((g.M + kTcM - 1) / kTcM, g.bp_tile ? (g.N + NT * NTL - 1) / (NT * NTL) : 1);
  mapf_launch_pdl(kernel, grid, dim3(kTcThreads), smem, s, g);
  return mapf_launch_status();
}

// ---------------------------------------------------------------------------------------------------------
// Same GEMM with the A operand ALREADY split and tiled by its producer (the encoder's conv kernel writes
// Ap [row/128][k/8][hi|lo][(row%128)/8][(k%8)/4][row%8][k%4]): no conversion warps, the whole operand pipeline is TMA.
//   warp 0 lane 0   : per stage of kPsChunks 32-byte K chunks, one bulk copy for A (contiguous in Ap) and one for B
//   warps 1.. lane 0: warp w issues the 3 tcgen05.mma (lo*hi, hi*lo, hi*hi) of chunk w-1 of every stage into its own
//                     TMEM accumulator (kPsChunks issuers: a thread issues about one MMA per 100 cycles)
// Stages are small and many (2 chunks x 4 stages, 96 KB, two CTAs per SM): the first MMAs start after 24 KB have landed and
// the producer refills at that granularity -- 10.7 -> 9.4 us at 20 480 rows against 4 chunks x 2 stages.
//   all 8 warps   : epilogue = bias, ReLU, transpose through shared memory, 128-byte row stores
// ---------------------------------------------------------------------------------------------------------
constexpr int kPsChunks = 2, kPsStages = 4;

template <int NT, bool F16>
__global__ void __launch_bounds__(kTcThreads) tc_presplit_kernel(TcGemmArgs g) {
  constexpr int kABytes = kTcM * 8 * 4, kBBytes = NT * 8 * 4;
  constexpr int kStageA = kPsChunks * 2 * kABytes, kStageB = kPsChunks * 2 * kBBytes;
  constexpr int kStageBytes = kStageA + kStageB;
  constexpr uint32_t kCols = kPsChunks * NT <= 256 ? 256 : 512;
  extern __shared__ __align__(128) uint8_t s_tc[];
  __shared__ __align__(8) uint64_t s_full[kPsStages], s_empty[kPsStages], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[NT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcM;
  if (warp == 0) tc::tmem_alloc(&s_tmem, kCols);
  if (tid == 0) {
    for (int s = 0; s < kPsStages; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], kPsChunks); }
    tc::mbar_init(&s_done, kPsChunks);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < NT; i += kTcThreads) s_bias[i] = (g.bias != nullptr && i < g.N) ? __ldg(g.bias + i) : 0.0f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  // warp-uniform copies for the MMA issuers: the whole issuing warp runs its (warp-uniform) branch and elect.sync picks the
  // lane, so every MMA operand is a uniform-register value and ptxas emits plain UTCHMMAs instead of wrapping each one in
  // an ELECT / R2UR.BROADCAST loop (~100 clocks per MMA behind a divergent `lane == 0` branch)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  mapf_pdl_wait();      // the A operand (and its row scales) come from the conv kernel in front
  mapf_pdl_launch();
  const int nch = F16 ? (g.K + 15) >> 4 : (g.K + 7) >> 3;   // a chunk is 32 bytes of K per row: 8 TF32 or 16 fp16
  const int nsc = (nch + kPsChunks - 1) / kPsChunks;

  if (warp_u == 0) {
    if (lane == 0) {
      const float* a_src = g.A + int64_t(blockIdx.x) * nch * (2 * kTcM * 8);   // (same bytes per chunk in both modes)
      for (int sc = 0; sc < nsc; ++sc) {
        const int s = sc % kPsStages;
        const uint32_t ph = uint32_t(sc / kPsStages) & 1u;
        tc::mbar_wait(&s_empty[s], ph ^ 1u);
        const int c0 = sc * kPsChunks;
        const int nc = nch - c0 < kPsChunks ? nch - c0 : kPsChunks;
        uint8_t* st = s_tc + s * kStageBytes;
        tc::mbar_expect_tx(&s_full[s], uint32_t(nc) * (2 * kABytes + 2 * kBBytes));
        tc::bulk_g2s(st, a_src + int64_t(c0) * (2 * kTcM * 8), uint32_t(nc) * 2 * kABytes, &s_full[s]);
        tc::bulk_g2s(st + kStageA, g.Bp + int64_t(c0) * (2 * NT * 8), uint32_t(nc) * 2 * kBBytes, &s_full[s]);
        tc::mbar_arrive(&s_full[s]);
      }
    }
    __syncwarp();
  } else if (warp_u <= kPsChunks) {
    // issuer warp w (1..kPsChunks) owns chunk w-1 of every stage and TMEM accumulator w-1: one thread issues only about
    // one MMA per 100 cycles, the issuers work in parallel; each accumulator is filled by one thread in a fixed order (reproducible)
    if (tc::elect_one()) {
      constexpr uint32_t idesc = F16 ? tc::idesc_f16(kTcM, NT) : tc::idesc_tf32(kTcM, NT);
      constexpr uint32_t sbo = 256, lbo = 128;
      const uint32_t base = tc::smem_u32(s_tc);
      const int ci = warp_u - 1;
      const uint32_t acc = tmem_u + uint32_t(ci) * NT;
      for (int sc = 0; sc < nsc; ++sc) {
        const int s = sc % kPsStages;
        const uint32_t ph = uint32_t(sc / kPsStages) & 1u;
        tc::mbar_wait(&s_full[s], ph);
        tc::fence_after_sync();
        const int c = sc * kPsChunks + ci;
        if (c < nch) {
          const uint32_t sa = base + uint32_t(s) * kStageBytes, sb = sa + kStageA;
          const uint64_t dah = tc::smem_desc(sa + uint32_t(ci) * 2 * kABytes, lbo, sbo);
          const uint64_t dal = tc::smem_desc(sa + uint32_t(ci) * 2 * kABytes + kABytes, lbo, sbo);
          const uint64_t dbh = tc::smem_desc(sb + uint32_t(ci) * 2 * kBBytes, lbo, sbo);
          const uint64_t dbl = tc::smem_desc(sb + uint32_t(ci) * 2 * kBBytes + kBBytes, lbo, sbo);
          if (F16) {
            tc::mma_f16(acc, dal, dbh, idesc, sc > 0);
            tc::mma_f16(acc, dah, dbl, idesc, true);
            tc::mma_f16(acc, dah, dbh, idesc, true);
          } else {
            tc::mma_tf32(acc, dal, dbh, idesc, sc > 0);   // first touch of an accumulator overwrites; small terms first
            tc::mma_tf32(acc, dah, dbl, idesc, true);
            tc::mma_tf32(acc, dah, dbh, idesc, true);
          }
        }
        tc::mma_commit(&s_empty[s]);
      }
      tc::mma_commit(&s_done);
    }
    __syncwarp();
  }

  {
    tc::mbar_wait(&s_done, 0);
    tc::fence_after_sync();
    const int quad = warp & 3, half = warp >> 2;
    constexpr int kHalfCols = NT / 2;
    float* s_scr = reinterpret_cast<float*>(s_tc);
    const int nacc = nch < kPsChunks ? nch : kPsChunks;   // accumulators that were written
    const bool vec_c = (g.ldc & 3) == 0 && (g.N & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0;
#pragma unroll
    for (int cbl = 0; cbl < kHalfCols; cbl += 32) {
      const int cb = half * kHalfCols + cbl;
      uint32_t r0[32], r1[32];
      float v[32];
      tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + cb, r0);
      if (nacc > 1) tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + NT + cb, r1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + (nacc > 1 ? __uint_as_float(r1[j]) : 0.0f);
      if (nacc > 2) {
        tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + 2 * NT + cb, r0);
        if (nacc > 3) tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + 3 * NT + cb, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r0[j]) + (nacc > 3 ? __uint_as_float(r1[j]) : 0.0f);
      }
      float rs = 1.0f;   // fp16-split operands: undo the row's and the matrix' power-of-two scales
      if (F16) {
        const int row = m0 + quad * 32 + lane;
        rs = (row < g.M ? mapf_ld_dep(g.row_scale + row) : 0.0f) * __ldg(g.b_hdr);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], rs, s_bias[cb + j]);
        if (g.relu) v[j] = fmaxf(v[j], 0.0f);
      }
      tc_store_block32(s_scr + warp * kTcTP, v, g.C, g.ldc, m0 + quad * 32, cb, g.M, g.N, vec_c, lane);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

template <int NT, bool F16>
int launch_presplit(const TcGemmArgs& g, cudaStream_t s) {
  constexpr size_t smem = size_t(kPsStages) * kPsChunks * (2 * kTcM * 8 * 4 + 2 * NT * 8 * 4);
  static_assert(smem >= 8 * kTcTP * 4, "the epilogue transposes through the operand stages");
  auto kernel = tc_presplit_kernel<NT, F16>;
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  mapf_launch_pdl(kernel, dim3((g.M + kTcM - 1) / kTcM), dim3(kTcThreads), smem, s, g);
  return mapf_launch_status();
}

}  // namespace

// tile shapes: N <= 64 -> (NT 64, 4 accumulators); N <= 128 -> (NT 128, 2 accumulators); KC = 8 floats per stage
static inline int tc_nt(int N) { return N <= 64 ? 64 : 128; }
constexpr int kTcKC = 8;

extern "C" int64_t mapf_tc_packed_b_size(int32_t N, int32_t K) {
  if (N <= 0 || N > 128 || K <= 0) return -1;
  const int64_t nchunks = (K + kTcKC - 1) / kTcKC;
  return nchunks * 2 * tc_nt(N) * kTcKC;
}

extern "C" int mapf_tc_pack_b(const float* b0, int32_t K0, const float* b1, int32_t K1, int32_t N, float* out,
                              mapf_stream_t stream) {
  MAPF_REQUIRE(b0 && out && (K1 == 0 || b1), MAPF_ERR_NULL);
  MAPF_REQUIRE(N > 0 && N <= 128 && K0 > 0 && K1 >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(out, 16), MAPF_ERR_ALIGN);
  const int64_t total = mapf_tc_packed_b_size(N, K0 + K1);
  tc_pack_b_kernel<<<int((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(b0, K0, b1, K1, N, tc_nt(N),
                                                                                           kTcKC, out);
  return mapf_launch_status();
}

// Encoder FC matrix [N, 25*C] for the fp16-split GEMM behind enc_tc.cu: K index re-ordered pixel-major (k' = pixel*C + c
// <- k = c*25 + pixel, the order in which the conv kernel emits the flattened activations), scaled by a power of two
// that brings the largest |w| into [2^9, 2^10), split into fp16 hi / lo.  out = [inverse scale, 0, 0, 0] then per
// 16-wide K chunk: hi block | lo block, each [n/8][k/8][n%8][k%8] (canonical no-swizzle K-major, NT rows).
extern "C" int64_t mapf_tc_packed_b_f16_size(int32_t N, int32_t K) {
  if (N <= 0 || N > 128 || K <= 0) return -1;
  return 4 + int64_t((K + 15) / 16) * tc_nt(N) * 16;
}

extern "C" int mapf_tc_pack_b_f16_pixel_major(const float* b, int32_t C, int32_t N, float* out, mapf_stream_t stream) {
  MAPF_REQUIRE(b && out, MAPF_ERR_NULL);
  MAPF_REQUIRE(N > 0 && N <= 128 && C > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(out, 16), MAPF_ERR_ALIGN);
  tc_pack_b_f16_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(b, C, N, tc_nt(N), out);
  return mapf_launch_status();
}

// C[M,N] = act(A B^T + bias) with both operands fp16-split: Ap = the conv kernel's output [row/128][k/16][hi|lo]
// [(row%128)/8][(k%16)/8][row%8][k%8] (fp16) followed, at float offset ceil(M/128)*ceil(K/16)*2048, by the inverse scale
// of every row; Bp from mapf_tc_pack_b_f16_pixel_major.
extern "C" int mapf_tc_fc_presplit_f16(int32_t M, int32_t N, int32_t K, const float* Ap, const float* Bp, float* C,
                                       int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(Ap && Bp && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(Ap, 128) && mapf_aligned(Bp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = Ap; g.Bp = Bp + 4;
  g.row_scale = Ap + int64_t((M + kTcM - 1) / kTcM) * ((K + 15) / 16) * 2048;
  g.b_hdr = Bp;
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_presplit<64, true>(g, s);
  return launch_presplit<128, true>(g, s);
}

// C[M,N] = act(A[M,K] B^T + bias): fp32 in / out, 3xTF32 on the tensor cores; Bp from mapf_tc_pack_b.
extern "C" int mapf_tc_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C,
                            int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(A && Bp && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((lda & 3) == 0 && mapf_aligned(A, 16) && mapf_aligned(Bp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.Bp = Bp;
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64, kTcKC, 2, EPI_STORE>(g, s);
  return launch_tc<128, kTcKC, 2, EPI_STORE>(g, s);
}

// Same GEMM with 4 stages and one accumulator (three CTAs per SM): the training step's activation-gradient and
// forward GEMMs (K <= 400), where the number of row tiles is close to the number of CTA slots.
extern "C" int mapf_tc_gemm_w3(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C,
                               int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(A && Bp && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((lda & 3) == 0 && mapf_aligned(A, 16) && mapf_aligned(Bp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.Bp = Bp;
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64, kTcKC, 1, EPI_STORE, 4>(g, s);
  return launch_tc<128, kTcKC, 1, EPI_STORE, 4>(g, s);
}

// Wide outputs: B [N, K] with any N, packed as ceil(N / 128) tiles of 128 rows (mapf_tc_b_floats(128, K) floats apart),
// and the GEMM over all tiles in ONE launch (grid.y = tile): the training step's dz = dy W_eff and da = dx W_fc.
extern "C" int mapf_tc_pack_b_tiles(const float* b, int32_t K, int32_t N, float* out, mapf_stream_t stream) {
  MAPF_REQUIRE(b && out, MAPF_ERR_NULL);
  MAPF_REQUIRE(N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_aligned(out, 16), MAPF_ERR_ALIGN);
  const int64_t tile = mapf_tc_packed_b_size(128, K);
  for (int n0 = 0, t = 0; n0 < N; n0 += 128, ++t) {
    const int nn = N - n0 < 128 ? N - n0 : 128;
    tc_pack_b_kernel<<<int((tile + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        b + int64_t(n0) * K, K, nullptr, 0, nn, 128, kTcKC, out + int64_t(t) * tile);
  }
  return mapf_launch_status();
}

extern "C" int mapf_tc_gemm_tiles(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C,
                                  int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(A && Bp && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE((lda & 3) == 0 && mapf_aligned(A, 16) && mapf_aligned(Bp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.Bp = Bp; g.bp_tile = mapf_tc_packed_b_size(128, K);
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  if (N <= 128) return launch_tc<128, kTcKC, 1, EPI_STORE, 4>(g, static_cast<cudaStream_t>(stream));
  return launch_tc<128, kTcKC, 1, EPI_STORE, 4, 2>(g, static_cast<cudaStream_t>(stream));   // pairs of tiles share the A operand
}

// logits / greedy actions = head(relu(Z Wp^T)): Z [M, K] row-major filter taps, Wp = packed [W | W2] (mapf_tc_pack_b).
extern "C" int mapf_tc_mix_head(int32_t M, int32_t N, int32_t K, const float* Z, int64_t ldz, const float* Wp,
                                const float* act_w, const float* act_b, int32_t num_actions, float* logits,
                                int32_t* actions, float* y, mapf_stream_t stream) {
  MAPF_REQUIRE(Z && Wp && act_w && act_b && logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128 && (N & 15) == 0 && num_actions > 0 && num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((ldz & 3) == 0 && mapf_aligned(Z, 16) && mapf_aligned(Wp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = Z; g.lda = ldz; g.Bp = Wp;
  g.act_w = act_w; g.act_b = act_b; g.num_actions = num_actions;
  g.logits = logits; g.actions = actions; g.y = y;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64, kTcKC, 2, EPI_HEAD>(g, s);
  return launch_tc<128, kTcKC, 2, EPI_HEAD>(g, s);
}

// Floats of the pre-split A operand [ceil(M/128)][ceil(K/8)][hi|lo][128 x 8] that mapf_encoder_conv_split_fwd writes
extern "C" int64_t mapf_tc_presplit_a_size(int64_t M, int32_t K) {
  if (M <= 0 || K <= 0) return -1;
  return ((M + kTcM - 1) / kTcM) * int64_t((K + 7) / 8) * (2 * kTcM * 8);
}

// C[M,N] = act(A B^T + bias) with A given pre-split / pre-tiled (Ap, see mapf_tc_presplit_a_size) and Bp from
// mapf_tc_pack_b: the Linear + ReLU of the encoder (compressMLP, framework_gnn.py:93) behind the conv kernel.
extern "C" int mapf_tc_fc_presplit(int32_t M, int32_t N, int32_t K, const float* Ap, const float* Bp, float* C,
                                   int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(Ap && Bp && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(Ap, 128) && mapf_aligned(Bp, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = Ap; g.Bp = Bp;
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_presplit<64, false>(g, s);
  return launch_presplit<128, false>(g, s);
}
