// fp64-accurate triangular GEMM on the INT8 tensor cores (Ozaki error-free splitting, tcgen05 kind::i8).
//
//   A = Linv · K*  is needed to ~2^-46 relative to |row scale|·|K* scale| for the 1e-9·σ_f² variance bar.
//   Each fp64 operand is split into S = 6 balanced base-256 digits (int8 in [-128, 127]) under a power-of-two
//   scale (per row of Linv; one global scale for K*):   x = 2^e · Σ_p d_p · 2^(-8p)   (48 bits kept).
//   Products of digit matrices are EXACT in the int32 TMEM accumulators, and all pairs with the same
//   p + q = r share one accumulator T_r, so
//        A[n,t] = 2^(e_n + f) · Σ_{r=2..R} 2^(-8r) · T_r[n,t],        R = 7  (21 digit products);
//   |T_r| <= 6 · K · 2^14 < 2^31 for K <= 16384.
//   TMEM holds 512 columns = four 128x128 int32 accumulators, so each row-block runs two passes:
//        pass LO: r = 6,7   (11 products, digits 1..6 of both operands), kept as fp64 in registers,
//        pass HI: r = 2..5  (10 products, digits 1..4), then scale, square and column-reduce.
//   Operands are pre-packed in the UMMA no-swizzle K-major core-matrix layout, so each pipeline stage is two
//   contiguous 1-D bulk-TMA copies.  Warp roles: 8 epilogue warps, 1 TMA producer, 1 MMA issuer.
#pragma once
#include "common.cuh"
#include "kernel_fn.cuh"
#include "umma.cuh"
#include <cfloat>

namespace tb {
namespace oz {

constexpr int S = 6;                         // digits per operand
constexpr int TILE = 128 * KST;              // one digit tile: 128 rows x 64 k-bytes = 8 KB
constexpr int HI_DIG = 4;                    // pass HI uses digits 1..4, pass LO digits 1..6
constexpr int STAGES_HI = 3, STAGES_LO = 2;
constexpr int STAGE_BYTES_HI = 2 * HI_DIG * TILE;   // 64 KB
constexpr int STAGE_BYTES_LO = 2 * S * TILE;        // 112 KB
constexpr size_t SMEM_BYTES = (size_t)STAGES_LO * STAGE_BYTES_LO + 256;   // 224 KB + barriers
constexpr int DIGIT_BITS = 48;               // v = rint(x / 2^e * 2^48) = Σ d_p 256^(6-p)


// balanced base-256 digits of v (|v| <= 2^46, so the top digit stays below 128): d[0] most significant
__device__ __forceinline__ void digits7(long long v, int d[S]) {
#pragma unroll
  for (int p = S - 1; p >= 0; --p) {
    int lo = (int)(((v + 128) & 255) - 128);
    d[p] = lo;
    v = (v - lo) >> 8;
  }
}

// The same 6 digits without a carry chain: v = Σ d_p 256^(6-p) with d_p in [-128,127]  <=>  the ordinary base-256 digits of
// u = v + Σ 128·256^i are d_p + 128, so the int8 digits are the bytes of (u ^ 0x808080808080); byte 0 = least significant = d_6.
__device__ __forceinline__ void digit_bytes6(long long v, uint32_t& lo, uint32_t& hi) {
  const unsigned long long K = 0x0000808080808080ULL;
  const unsigned long long w = ((unsigned long long)v + K) ^ K;
  lo = (uint32_t)w;
  hi = (uint32_t)(w >> 32);
}
// scatter the six digit bytes of one element into the six digit planes: element index JJ (0..15) within the lane's 16-byte rows
template <int JJ>
__device__ __forceinline__ void scatter_digits(uint32_t (&pk)[S][4], uint32_t lo, uint32_t hi) {
  // plane p (0 = most significant digit d_1) takes byte (5 - p) of w
  pk[0][JJ >> 2] = put_byte<JJ & 3, 1>(pk[0][JJ >> 2], hi);
  pk[1][JJ >> 2] = put_byte<JJ & 3, 0>(pk[1][JJ >> 2], hi);
  pk[2][JJ >> 2] = put_byte<JJ & 3, 3>(pk[2][JJ >> 2], lo);
  pk[3][JJ >> 2] = put_byte<JJ & 3, 2>(pk[3][JJ >> 2], lo);
  pk[4][JJ >> 2] = put_byte<JJ & 3, 1>(pk[4][JJ >> 2], lo);
  pk[5][JJ >> 2] = put_byte<JJ & 3, 0>(pk[5][JJ >> 2], lo);
}

__device__ __forceinline__ void scatter_digits_rt(uint32_t (&pk)[S][4], int jj, uint32_t lo, uint32_t hi) {
  switch (jj) {  // jj is a compile-time constant after unrolling: the switch folds away
    case 0: scatter_digits<0>(pk, lo, hi); break;
    case 1: scatter_digits<1>(pk, lo, hi); break;
    case 2: scatter_digits<2>(pk, lo, hi); break;
    case 3: scatter_digits<3>(pk, lo, hi); break;
    case 4: scatter_digits<4>(pk, lo, hi); break;
    case 5: scatter_digits<5>(pk, lo, hi); break;
    case 6: scatter_digits<6>(pk, lo, hi); break;
    case 7: scatter_digits<7>(pk, lo, hi); break;
    case 8: scatter_digits<8>(pk, lo, hi); break;
    case 9: scatter_digits<9>(pk, lo, hi); break;
    case 10: scatter_digits<10>(pk, lo, hi); break;
    case 11: scatter_digits<11>(pk, lo, hi); break;
    case 12: scatter_digits<12>(pk, lo, hi); break;
    case 13: scatter_digits<13>(pk, lo, hi); break;
    case 14: scatter_digits<14>(pk, lo, hi); break;
    default: scatter_digits<15>(pk, lo, hi); break;
  }
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(IDESC), "r"(acc)
               : "memory");
}


// ------------------------------------------------------------------------------------------------
// once per BO step: digit tiles of Linv.  grid = (stage kc, row-block I), 256 threads.
//   rowscale[n] = 2^e_n with 2^e_n > 2 max_k |Linv[n,k]|  (so |x|/2^e < 1/2 and the top digit fits)
// ------------------------------------------------------------------------------------------------
__global__ void linv_rowscale_kernel(const double* __restrict__ Linv, int64_t N, int64_t rows, double* __restrict__ rowscale) {
  const int64_t n = blockIdx.x;
  double mx = 0.0;
  if (n < N)
    for (int64_t k = threadIdx.x; k <= n; k += blockDim.x) mx = fmax(mx, fabs(Linv[n + k * N]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sm[w]);
    int e = 0;
    if (mx > 0.0) {
      frexp(mx, &e);  // mx = m 2^e, m in [0.5, 1)
      e += 2;         // |x| / 2^e < 1/4
    }
    if (n < rows) rowscale[n] = ldexp(1.0, e);
  }
}

__global__ void linv_digits_kernel(const double* __restrict__ Linv, int64_t N, const double* __restrict__ rowscale,
                                   int8_t* __restrict__ AS) {
  const int I = blockIdx.y, kc = blockIdx.x;
  if (kc >= 2 * (I + 1)) return;
  int8_t* dst = AS + (a_stage_offset(I) + kc) * (int64_t)(S * TILE);
  for (int e = threadIdx.x; e < 128 * KST; e += blockDim.x) {
    const int r = e % 128, kin = e / 128;  // r fastest: column-major source is contiguous in n
    const int64_t n = (int64_t)I * 128 + r, k = (int64_t)kc * KST + kin;
    long long v = 0;
    if (n < N && k <= n) v = __double2ll_rn(Linv[n + k * N] / rowscale[n] * 281474976710656.0);  // 2^48
    int d[S];
    digits7(v, d);
    const int off = (r >> 3) * SBO + (kin >> 4) * LBO + (r & 7) * 16 + (kin & 15);
#pragma unroll
    for (int p = 0; p < S; ++p) dst[p * TILE + off] = (int8_t)d[p];
  }
}

// K^-1 (gradient path): symmetric matrix given by its lower triangle (column-major, cusolver potri); full rows.
__device__ __forceinline__ double sym_at(const double* __restrict__ A, int64_t N, int64_t n, int64_t k) {
  return n >= k ? A[n + k * N] : A[k + n * N];
}
__global__ void sym_rowscale_kernel(const double* __restrict__ A, int64_t N, int64_t rows, double* __restrict__ rowscale) {
  const int64_t n = blockIdx.x;
  double mx = 0.0;
  if (n < N)
    for (int64_t k = threadIdx.x; k < N; k += blockDim.x) mx = fmax(mx, fabs(sym_at(A, N, n, k)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sm[w]);
    int e = 0;
    if (mx > 0.0) {
      frexp(mx, &e);
      e += 2;
    }
    if (n < rows) rowscale[n] = ldexp(1.0, e);
  }
}
__global__ void sym_digits_kernel(const double* __restrict__ A, int64_t N, int nst, const double* __restrict__ rowscale,
                                  int8_t* __restrict__ AS) {
  const int I = blockIdx.y, kc = blockIdx.x;
  int8_t* dst = AS + ((int64_t)I * nst + kc) * (int64_t)(S * TILE);
  for (int e = threadIdx.x; e < 128 * KST; e += blockDim.x) {
    const int r = e % 128, kin = e / 128;
    const int64_t n = (int64_t)I * 128 + r, k = (int64_t)kc * KST + kin;
    long long v = 0;
    if (n < N && k < N) v = __double2ll_rn(sym_at(A, N, n, k) / rowscale[n] * 281474976710656.0);
    int d[S];
    digits7(v, d);
    const int off = (r >> 3) * SBO + (kin >> 4) * LBO + (r & 7) * 16 + (kin & 15);
#pragma unroll
    for (int p = 0; p < S; ++p) dst[p * TILE + off] = (int8_t)d[p];
  }
}

// ------------------------------------------------------------------------------------------------
// K* digit tiles + posterior mean.  grid = candidate tiles x (512 / blockDim); warp w (0..15 within a tile) owns candidates
// [8w, 8w+8); lane l <-> (candidate l % 8, 16-wide k chunk l / 8): every digit store of a warp is 512
// contiguous bytes (four adjacent core matrices).
// ------------------------------------------------------------------------------------------------
template <int KIND, int DP>
__global__ void __launch_bounds__(512, 2)
kstar_digits_kernel(const double* __restrict__ Xs, const double* __restrict__ alpha, const double* __restrict__ Xc,
                    const double* __restrict__ inv_ls, int N, int nst, int D, int64_t M, double variance,
                    double inv_bscale_2p48, double mean_const, int8_t* __restrict__ BS, double* __restrict__ mean_out) {
  // CTA size is free (any multiple of 32 dividing 512): each warp owns one 8-candidate row group of a tile
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), w = wg & 15, tile_id = wg >> 4;
  const int cl = lane & 7, ch = lane >> 3;
  const int t_local = w * 8 + cl;
  const int64_t t = (int64_t)tile_id * 128 + t_local;
  const bool valid = t < M;
  double xc[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) xc[d] = (valid && d < D) ? Xc[t * D + d] * inv_ls[d] : 0.0;
  int8_t* tile = BS + (int64_t)tile_id * nst * (S * TILE) + w * SBO + ch * LBO + cl * 16;
  // the 64 training rows (+ alpha) of a stage are staged through shared memory with cp.async, double-buffered: ncu showed the
  // kernel waiting on L1/L2 latency of these warp-broadcast loads (long_scoreboard was the top stall)
  __shared__ __align__(16) double xs_s[2][KST * DP];
  __shared__ __align__(16) double al_s[2][KST];
  auto stage_load = [&](int kc, int buf) {
    const double* src = Xs + (int64_t)kc * KST * DP;
    for (int e = threadIdx.x; e < KST * DP / 2; e += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&xs_s[buf][2 * e])), "l"(src + 2 * e) : "memory");
    for (int e = threadIdx.x; e < KST / 2; e += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&al_s[buf][2 * e])), "l"(alpha + (int64_t)kc * KST + 2 * e)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_load(0, 0);
  double macc = 0.0;
  for (int kc = 0; kc < nst; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nst) {
      stage_load(kc + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    uint32_t pk[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p) pk[p][0] = pk[p][1] = pk[p][2] = pk[p][3] = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int kl = ch * 16 + j, k = kc * KST + kl;
      const double* xr = &xs_s[buf][kl * DP];
      double r2 = 0.0;
#pragma unroll
      for (int d = 0; d < DP; d += 2) {
        const double2 v = *reinterpret_cast<const double2*>(xr + d);
        double d0 = xc[d] - v.x, d1 = xc[d + 1] - v.y;
        r2 = fma(d0, d0, r2);
        r2 = fma(d1, d1, r2);
      }
      const double kval = (valid && k < N) ? kernel_from_r2<KIND>(r2, variance) : 0.0;
      macc = fma(kval, al_s[buf][kl], macc);
      uint32_t wl, wh;
      digit_bytes6(__double2ll_rn(kval * inv_bscale_2p48), wl, wh);
      scatter_digits_rt(pk, j, wl, wh);
    }
#pragma unroll
    for (int p = 0; p < S; ++p)
      *reinterpret_cast<uint4*>(tile + (int64_t)kc * (S * TILE) + p * TILE) = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
    __syncthreads();  // everyone is done with xs_s[buf] before the next iteration's prefetch overwrites it
  }
  macc += __shfl_xor_sync(0xffffffffu, macc, 8);
  macc += __shfl_xor_sync(0xffffffffu, macc, 16);
  if (ch == 0) mean_out[(int64_t)tile_id * 128 + t_local] = macc + mean_const;
}

// all MMAs of one pipeline stage, fully unrolled at compile time: per MMA two 32-bit adds on precomputed descriptors
// (the issuing thread shares its scheduler with co-resident K*-generation warps, so its instruction count matters)
__device__ __forceinline__ void umma_i8_desc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(IDESC), "r"(acc)
      : "memory");
}
// pass order inside a CTA: LO HI | HI LO | LO HI ... so that consecutive row-blocks meet on the same stage geometry
__device__ __forceinline__ bool pass_is_lo(int i, int j) { return ((i & 1) == 0) == (j == 0); }

template <int NDIG, int RLO, int RHI>
__device__ __forceinline__ void issue_stage(uint32_t tmem, uint32_t stage_base, uint32_t not_first_kc) {
  // descriptor = [hi: SBO | version][lo: LBO | start>>4]
  constexpr uint32_t HI = ((SBO >> 4) & 0x3FFF) | (1u << 14);
  const uint32_t a0 = ((stage_base >> 4) & 0x3FFF) | (((LBO >> 4) & 0x3FFF) << 16);
  const uint32_t b0 = a0 + ((NDIG * TILE) >> 4);
#pragma unroll
  for (int p = 1; p <= NDIG; ++p)
#pragma unroll
    for (int q = 1; q <= NDIG; ++q) {
      const int r = p + q;
      if (r < RLO || r > RHI) continue;
      const uint32_t acc = tmem + (uint32_t)(r - RLO) * 128u;
      const bool first_pair = (p == (r - NDIG > 1 ? r - NDIG : 1));
#pragma unroll
      for (int kk = 0; kk < KST / 32; ++kk) {
        const uint32_t flag = (first_pair && kk == 0) ? not_first_kc : 1u;
        umma_i8_desc(acc, a0 + (((p - 1) * TILE + kk * 2 * (int)LBO) >> 4), b0 + (((q - 1) * TILE + kk * 2 * (int)LBO) >> 4), HI, flag);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// the GEMM: grid = (candidate tiles, G); partial[g][t] = Σ_{rows n of group g} A[n,t]^2
// ------------------------------------------------------------------------------------------------
// OZ_SUMSQ: partial column sums of A^2 (variance path).  OZ_STORE: A itself, fp64, candidate-major [t][lda] (joint path)
enum { OZ_SUMSQ = 0, OZ_STORE = 1 };
// EW = number of epilogue warps (8: 64 accumulator columns each; 4 = 128 columns each was measured slower and is not instantiated)
template <int EPI, int EW>
__device__ __forceinline__ void trigemm_i8_body(const int8_t* __restrict__ AS, const int8_t* __restrict__ BS,
                                                const double* __restrict__ rowscale, int NB, int nst, int G, int64_t McPad,
                                                double out_scale, int npass, int full_rows, double* __restrict__ partial,
                                                double* __restrict__ Aplain, int64_t lda) {
  // full_rows = 0: lower-triangular left factor (Linv): row-block I spans stages [0, 2(I+1)), packed triangularly.
  // full_rows = 1: dense square left factor (K^-1, gradient path): every row-block spans all nst stages, offset I*nst.
  // npass = 2: full fp64 accuracy (LO + HI passes, 21 digit products).  npass = 1: HI pass only (digits 1..4, 10 products,
  // ~2^-28 of the operand scales) — used for fp32 models, whose tolerance it exceeds by orders of magnitude.
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES_LO * STAGE_BYTES_LO);
  uint64_t* full_hi = bars;            // [3]
  uint64_t* empty_hi = bars + 3;       // [3]
  uint64_t* full_lo = bars + 6;        // [2]
  uint64_t* empty_lo = bars + 8;       // [2]
  uint64_t* acc_full = bars + 10;      // MMA -> epilogue (and producer: all MMAs of the pass retired)
  uint64_t* acc_empty = bars + 11;     // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, tile = blockIdx.y;  // g fastest: co-resident CTAs share few candidate tiles -> K* digits stay in L2

  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(&full_hi[s], 1); mbar_init(&empty_hi[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&full_lo[s], 1); mbar_init(&empty_lo[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EW);
    fence_barrier_init();
  }
  if (warp == EW) {  // TMEM allocation: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int8_t* bTile = BS + (int64_t)tile * nst * (S * TILE);

  if (warp == EW) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st_hi = 0, st_lo = 0, n = 0, seen = 0;  // seen = completions of acc_full already observed
      uint32_t ph_hi = 0, ph_lo = 0;
      bool prev_lo = false;
      for (int i = 0;; ++i) {
        const int I = serpentine_rowblock(i, g, G);
        if (I >= NB) break;
        const int nk = full_rows ? nst : min(2 * (I + 1), nst);
        const int8_t* aRow = AS + (full_rows ? (int64_t)I * nst : a_stage_offset(I)) * (int64_t)(S * TILE);
        for (int j = 0; j < npass; ++j, ++n) {
          const bool lo = npass == 2 && pass_is_lo(i, j);
          // the two pass types lay different stage geometries over the same bytes: on a type change wait until every MMA
          // of the previous pass has retired (completion #n of acc_full); same-type passes just keep streaming
          // (completions are consumed one by one: a parity wait cannot tell completion #k from #k+2)
          if (n > 0 && lo != prev_lo)
            for (; seen < n; ++seen) mbar_wait(acc_full, (uint32_t)(seen & 1));
          prev_lo = lo;
          for (int kc = 0; kc < nk; ++kc) {
            const int8_t* a = aRow + (int64_t)kc * (S * TILE);
            const int8_t* b = bTile + (int64_t)kc * (S * TILE);
            if (lo) {
              mbar_wait(&empty_lo[st_lo], ph_lo ^ 1);
              unsigned char* dst = smem + (size_t)st_lo * STAGE_BYTES_LO;
              mbar_expect_tx(&full_lo[st_lo], STAGE_BYTES_LO);
              bulk_g2s(dst, a, S * TILE, &full_lo[st_lo]);
              bulk_g2s(dst + S * TILE, b, S * TILE, &full_lo[st_lo]);
              if (++st_lo == STAGES_LO) { st_lo = 0; ph_lo ^= 1; }
            } else {
              mbar_wait(&empty_hi[st_hi], ph_hi ^ 1);
              unsigned char* dst = smem + (size_t)st_hi * STAGE_BYTES_HI;
              mbar_expect_tx(&full_hi[st_hi], STAGE_BYTES_HI);
              bulk_g2s(dst, a, HI_DIG * TILE, &full_hi[st_hi]);
              bulk_g2s(dst + HI_DIG * TILE, b, HI_DIG * TILE, &full_hi[st_hi]);
              if (++st_hi == STAGES_HI) { st_hi = 0; ph_hi ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == EW + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int st_hi = 0, st_lo = 0, n = 0;
      uint32_t ph_hi = 0, ph_lo = 0;
      for (int i = 0;; ++i) {
        const int I = serpentine_rowblock(i, g, G);
        if (I >= NB) break;
        const int nk = full_rows ? nst : min(2 * (I + 1), nst);
        for (int j = 0; j < npass; ++j, ++n) {
          const bool lo = npass == 2 && pass_is_lo(i, j);
          if (n > 0) {  // accumulators must have been read out by the epilogue warps (completion #n of acc_empty)
            mbar_wait(acc_empty, (uint32_t)((n - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          for (int kc = 0; kc < nk; ++kc) {
            if (lo) {
              mbar_wait(&full_lo[st_lo], ph_lo);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              issue_stage<S, 6, 7>(tmem, smem_u32(smem + (size_t)st_lo * STAGE_BYTES_LO), kc != 0 ? 1u : 0u);
              umma_commit(&empty_lo[st_lo]);
              if (++st_lo == STAGES_LO) { st_lo = 0; ph_lo ^= 1; }
            } else {
              mbar_wait(&full_hi[st_hi], ph_hi);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              issue_stage<HI_DIG, 2, 5>(tmem, smem_u32(smem + (size_t)st_hi * STAGE_BYTES_HI), kc != 0 ? 1u : 0u);
              umma_commit(&empty_hi[st_hi]);
              if (++st_hi == STAGES_HI) { st_hi = 0; ph_hi ^= 1; }
            }
          }
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // warp w reads TMEM lanes [32 (w%4), +32) (rows) and columns [64 (w/4), +64) of every accumulator.
    // Per pass: TMEM -> fp64 partial sums (cheap), release the accumulators immediately, and only on the second pass of a
    // row-block do the expensive part (scale, square, cross-lane column sums) while the tensor pipe already runs on.
    const int lq = warp & 3, ch = warp >> 2;
    constexpr int CW = 512 / EW;  // accumulator columns per epilogue warp
    const uint32_t lane_base = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * CW);
    double vacc[CW];
    double colsum[CW / 16];  // even lanes: column ch*CW + h*16 + (lane >> 1)
#pragma unroll
    for (int h = 0; h < CW / 16; ++h) colsum[h] = 0.0;
    int n = 0;
    for (int i = 0;; ++i) {
      const int I = serpentine_rowblock(i, g, G);
      if (I >= NB) break;
      const double rs = rowscale[(int64_t)I * 128 + lq * 32 + lane] * out_scale;
      for (int j = 0; j < npass; ++j, ++n) {
        const bool lo = npass == 2 && pass_is_lo(i, j);
        mbar_wait(acc_full, (uint32_t)(n & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lo) {
#pragma unroll
          for (int h = 0; h < CW / 16; ++h) {
            uint32_t t7[16], t6[16];
            tmem_ld16(lane_base + 1 * 128 + h * 16, t7);
            tmem_ld16(lane_base + 0 * 128 + h * 16, t6);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              double v = (double)(int)t7[c] * 0x1p-56;
              v = fma((double)(int)t6[c], 0x1p-48, v);
              vacc[h * 16 + c] = (j == 0) ? v : vacc[h * 16 + c] + v;
            }
          }
        } else {
#pragma unroll
          for (int h = 0; h < CW / 16; ++h) {
            uint32_t t5[16], t4[16], t3[16], t2[16];
            tmem_ld16(lane_base + 3 * 128 + h * 16, t5);
            tmem_ld16(lane_base + 2 * 128 + h * 16, t4);
            tmem_ld16(lane_base + 1 * 128 + h * 16, t3);
            tmem_ld16(lane_base + 0 * 128 + h * 16, t2);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              double v = (double)(int)t5[c] * 0x1p-40;
              v = fma((double)(int)t4[c], 0x1p-32, v);
              v = fma((double)(int)t3[c], 0x1p-24, v);
              v = fma((double)(int)t2[c], 0x1p-16, v);
              vacc[h * 16 + c] = (j == 0) ? v : vacc[h * 16 + c] + v;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);  // accumulators are free again: the next pass starts now
        if (j == npass - 1 && EPI == OZ_STORE) {
          // A[n,t] = rowscale * 2^f * v, stored candidate-major: the 32 lanes of a warp write 32 consecutive rows (256 B)
          const int64_t nrow = (int64_t)I * 128 + lq * 32 + lane;
          double* dstA = Aplain + ((int64_t)tile * 128 + ch * CW) * lda + nrow;
#pragma unroll
          for (int c = 0; c < CW; ++c) dstA[(int64_t)c * lda] = vacc[c] * rs;
        }
        if (j == npass - 1 && EPI == OZ_SUMSQ) {
          // A[n,t] = rowscale * 2^f * v; column sums of A^2 over the warp's 32 rows by recursive halving
          // (16 + 8 + 4 + 2 + 2 shuffles per 16 columns instead of 160)
#pragma unroll
          for (int h = 0; h < CW / 16; ++h) {
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const double v = vacc[h * 16 + c] * rs;
              a[c] = v * v;
            }
            double b8[8], b4[4], b2[2];
            bool up = (lane & 16) != 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const double mine = up ? a[8 + c] : a[c], theirs = up ? a[c] : a[8 + c];
              b8[c] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
            }
            up = (lane & 8) != 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double mine = up ? b8[4 + c] : b8[c], theirs = up ? b8[c] : b8[4 + c];
              b4[c] = mine + __shfl_xor_sync(0xffffffffu, theirs, 8);
            }
            up = (lane & 4) != 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const double mine = up ? b4[2 + c] : b4[c], theirs = up ? b4[c] : b4[2 + c];
              b2[c] = mine + __shfl_xor_sync(0xffffffffu, theirs, 4);
            }
            up = (lane & 2) != 0;
            double e = (up ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, up ? b2[0] : b2[1], 2);
            e += __shfl_xor_sync(0xffffffffu, e, 1);
            colsum[h] += e;  // lane holds column h*16 + (lane >> 1) (both lanes of a pair hold the same sum)
          }
        }
      }
    }
    if (EPI == OZ_SUMSQ) {
    // combine the four row-quarters (warps lq = 0..3 of the same column half) through shared memory
    // (every MMA has retired, so the stage buffers are free to hold the 4 KB of partial column sums)
    double (*redbuf)[CW] = reinterpret_cast<double (*)[CW]>(smem);
    if ((lane & 1) == 0) {
#pragma unroll
      for (int h = 0; h < CW / 16; ++h) redbuf[warp][h * 16 + (lane >> 1)] = colsum[h];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EW * 32));
    for (int col = threadIdx.x; col < 128; col += EW * 32) {  // warps of one column group are ch*4 .. ch*4+3
      const int cg = col / CW, cc = col % CW, wb = cg * 4;
      partial[(int64_t)g * McPad + (int64_t)tile * 128 + col] =
          redbuf[wb][cc] + redbuf[wb + 1][cc] + redbuf[wb + 2][cc] + redbuf[wb + 3][cc];
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == EW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int EPI, int EW>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
trigemm_i8_kernel(const int8_t* __restrict__ AS, const int8_t* __restrict__ BS, const double* __restrict__ rowscale, int NB,
                  int nst, int G, int64_t McPad, double out_scale, int npass, int full_rows, double* __restrict__ partial,
                  double* __restrict__ Aplain, int64_t lda) {
  trigemm_i8_body<EPI, EW>(AS, BS, rowscale, NB, nst, G, McPad, out_scale, npass, full_rows, partial, Aplain, lda);
}
// same kernel capped at 128 registers / thread (41 K registers per CTA): leaves room for three 128-thread K*-generation CTAs
// on the same SM when the generation of the next chunk runs concurrently on the low-priority stream (TB_OZ_OVERLAP=1)
template <int EPI, int EW>
__global__ void __maxnreg__(128)
trigemm_i8_lowreg_kernel(const int8_t* __restrict__ AS, const int8_t* __restrict__ BS, const double* __restrict__ rowscale, int NB,
                         int nst, int G, int64_t McPad, double out_scale, int npass, int full_rows, double* __restrict__ partial,
                         double* __restrict__ Aplain, int64_t lda) {
  trigemm_i8_body<EPI, EW>(AS, BS, rowscale, NB, nst, G, McPad, out_scale, npass, full_rows, partial, Aplain, lda);
}

}  // namespace oz
}  // namespace tb
