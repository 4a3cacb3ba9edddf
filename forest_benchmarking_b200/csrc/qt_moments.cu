// The step BEFORE the tomography path: raw shot data -> ExperimentResult.expectation / std_err.
//
// shots_to_obs_moments (observable_estimation.py:804-853): a bitarray [n_shots, n_qubits] of 0/1 bytes, the
// columns of the observable's qubits, eigenvalue product prod(1 - 2 b) per shot, then mean and variance of the
// mean (optionally through the Beta(n+ + 1, n- + 1) posterior).  Batched over settings: bits[B, n_shots, n_qubits].
// calibrate_observable_estimates' arithmetic (:1033-1049 + ratio_variance :1052-1090) is the elementwise kernel below.
//
// Integer / byte work, HBM-bound: one byte per shot and qubit is read once, two doubles per setting are written.
// For n_qubits in {1, 2, 4, 8} the whole array is read as one flat stream of aligned 16-byte words by 8 / 16 / 32
// lanes per setting (a setting's byte range may start anywhere: the bytes of neighbouring settings in its first /
// last word are masked off), and the per-shot parity is folded inside the 32-bit words (SWAR) before one popcount
// per word.  The other widths up to 16 (and misaligned 2 / 4 / 8-column arrays) read the same stream, compact every word
// to one bit per byte and take the per-shot parities with a sliding XOR (moments_stream_kernel); settings of more than
// 16 columns stage the setting through shared memory (warp per setting) and XOR the selected columns.
#include "qt_common.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

// per-u32 pattern with 0x01 in byte i if column (first_col + i) % Q of the bitarray is selected
__device__ __forceinline__ unsigned col_pattern(unsigned colmask, int first_col, int Q) {
  unsigned pat = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) pat |= ((colmask >> ((first_col + i) % Q)) & 1u) << (8 * i);
  return pat;
}

// n_shots is common to the batch: its reciprocals come in as kernel arguments (MomInv), so the epilogue has no division
struct MomInv {
  double inv_n;        // 1 / S
  double inv_ab;       // 1 / (S + 2)
  double inv_ab2_ab1;  // 1 / ((S + 2)^2 (S + 3))
};

__device__ __forceinline__ void moments_epilogue(unsigned long long n_minus, long long S, bool identity, double coeff,
                                                 int prior, MomInv iv, double* mean, double* var) {
  if (identity) {  // identity term (:826-827)
    *mean = coeff;
    *var = 0.0;
    return;
  }
  const double nm = (double)n_minus, np_ = (double)(S - (long long)n_minus);
  if (prior) {  // Beta(n+ + 1, n- + 1) mean / variance, then bit -> Pauli moments (:838-847, utils.py:446-458)
    const double a = np_ + 1.0, b = nm + 1.0;
    *mean = (2.0 * (a * iv.inv_ab) - 1.0) * coeff;
    *var = 4.0 * (a * b * iv.inv_ab2_ab1) * coeff * coeff;
  } else {  // np.mean / np.var of the +-coeff values, variance of the mean (:849-851)
    const double m = coeff * (np_ - nm) * iv.inv_n;
    const double dp = coeff - m, dm = -coeff - m;
    *mean = m;
    *var = (np_ * dp * dp + nm * dm * dm) * iv.inv_n * iv.inv_n;
  }
}

// Number of shots with eigenvalue product -1 inside one aligned 16-byte word.  Every byte is 0 or 1, so the four
// 32-bit words are first packed into bit planes 0..3 of one word (z = x0 + 2 x1 + 4 x2 + 8 x3, no carries), masked
// with the column pattern replicated over the planes, and the per-shot XOR fold and the popcount run once.
template <int Q>
__device__ __forceinline__ unsigned swar_count(uint4 w, unsigned pat0, unsigned pat1) {
  if (Q == 8) {  // a shot is two words with different column patterns
    const unsigned y0 = (w.x & pat0) ^ (w.y & pat1), y1 = (w.z & pat0) ^ (w.w & pat1);
    unsigned z = y0 + 2u * y1;
    z ^= z >> 16;
    z ^= z >> 8;
    return __popc(z & 0x3u);
  }
  unsigned z = (w.x + 2u * w.y + 4u * w.z + 8u * w.w) & (pat0 * 15u);
  if (Q == 1) return __popc(z);
  if (Q == 2) {
    z ^= z >> 8;
    return __popc(z & 0x000f000fu);
  }
  z ^= z >> 16;
  z ^= z >> 8;
  return __popc(z & 0xfu);
}

// 16-byte load that never touches memory outside the tensor [tlo, thi): a word that straddles either end is assembled
// from byte loads (only the first word of the first setting and the last word of the last one can do so)
__device__ __forceinline__ uint4 ld16_inside(long long a, long long tlo, long long thi) {
  if (a >= tlo && a + 16 <= thi) return *reinterpret_cast<const uint4*>(a);
  unsigned w[4] = {0u, 0u, 0u, 0u};
  for (int i = 0; i < 16; ++i)
    if (a + i >= tlo && a + i < thi) w[i >> 2] |= (unsigned)(*reinterpret_cast<const unsigned char*>(a + i)) << (8 * (i & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// the bytes of the 16-byte word at flat address e that lie outside [lo, hi) are cleared
__device__ __forceinline__ uint4 clip_word(uint4 w, long long e, long long lo, long long hi) {
  const int b0 = (int)max(lo - e, 0LL), b1 = (int)min(hi - e, 16LL);       // valid bytes [b0, b1)
  const unsigned en = ((1u << b1) - 1u) & ~((1u << b0) - 1u);             // 16 byte-enable bits
  auto expand = [](unsigned nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xffu; };
  w.x &= expand(en & 15u);
  w.y &= expand((en >> 4) & 15u);
  w.z &= expand((en >> 8) & 15u);
  w.w &= expand((en >> 12) & 15u);
  return w;
}

// G lanes per setting (8 / 16 / 32, chosen from the setting's size so that a lane has ~8 independent 16-byte loads in
// flight).  Words that lie entirely inside the setting's byte range take the unmasked loop (4 loads issued before the
// first use); the at most two edge words shared with the neighbouring settings are masked byte-wise.
template <int Q, int G>
__global__ void __launch_bounds__(256)  // 43 registers, 5 blocks per SM; capping at 32 registers measured slower
   
    moments_swar_kernel(int64_t B, int64_t S, const unsigned char* __restrict__ bits,
                        const uint32_t* __restrict__ colmask, const double* __restrict__ coeff, int prior, MomInv iv,
                        double* __restrict__ mean, double* __restrict__ var) {
  static_assert(Q == 1 || Q == 2 || Q == 4 || Q == 8, "shots must tile a 32-bit word");
  const int gl = threadIdx.x % G;
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ngrp = ((int64_t)gridDim.x * blockDim.x) / G;
  const long long base = (long long)reinterpret_cast<uintptr_t>(bits);  // flat addresses: the alignment of `bits` is free
  const int64_t rounds = (B + ngrp - 1) / ngrp;  // every lane of a warp runs the same number of rounds (shuffles below)
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t b = grp + it * ngrp;
    const bool live = b < B;
    const unsigned cm = live ? (colmask[b] & ((1u << Q) - 1u)) : 0u;
    const long long lo = base + (live ? b : 0) * S * Q, hi = live ? lo + S * Q : lo;
    const unsigned pat0 = col_pattern(cm, 0, Q), pat1 = (Q == 8) ? col_pattern(cm, 4, Q) : pat0;
    unsigned cnt = 0;
    const long long A0 = (lo + 15) & ~15LL, A1 = hi & ~15LL, stride = 16LL * G;
    long long a = A0 + 16LL * gl;
    for (; a + 3 * stride < A1; a += 4 * stride) {
      const uint4 w0 = *reinterpret_cast<const uint4*>(a), w1 = *reinterpret_cast<const uint4*>(a + stride);
      const uint4 w2 = *reinterpret_cast<const uint4*>(a + 2 * stride), w3 = *reinterpret_cast<const uint4*>(a + 3 * stride);
      cnt += swar_count<Q>(w0, pat0, pat1) + swar_count<Q>(w1, pat0, pat1) + swar_count<Q>(w2, pat0, pat1) +
             swar_count<Q>(w3, pat0, pat1);
    }
    for (; a < A1; a += stride) {
      cnt += swar_count<Q>(*reinterpret_cast<const uint4*>(a), pat0, pat1);
    }
    // edge words
    const long long E1 = lo & ~15LL, E2 = hi & ~15LL;
    const bool need1 = live && (lo & 15), need2 = live && (hi & 15) && (E2 != E1 || !(lo & 15));
    if ((need1 && gl == 0) || (need2 && gl == (G > 1 ? 1 : 0))) {
      const long long e = (need1 && gl == 0) ? E1 : E2;
      cnt += swar_count<Q>(clip_word(ld16_inside(e, base, base + B * S * Q), e, lo, hi), pat0, pat1);
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (live && gl == 0) moments_epilogue(tot, S, cm == 0, coeff[b], prior, iv, mean + b, var + b);
  }
}

// Widths that do not tile a 32-bit word (3, 5, 6, 7, 9 .. 16 columns; also 2 / 4 / 8 columns when `bits` is not
// aligned to the shot width): the same flat stream of aligned 16-byte words, but a
// shot may straddle two words and its position inside a word changes from word to word.  Each word is first COMPACTED
// to one bit per byte (bytes are 0/1: (x * 0x01020408) >> 24 packs four of them), masked with the column pattern --
// periodic in the stream with period Q, so one 32-bit constant per setting shifted by the word's phase (a - lo) mod Q --
// and joined with the compacted previous word (the neighbouring lane's, by shuffle).  A sliding XOR over Q bits then
// holds, at every shot-END position, the parity of that shot; the end positions are again a periodic mask.  About 35
// integer instructions per 16 bytes, no shared memory, no alignment requirement on `bits`.
template <int Q>
__device__ __forceinline__ unsigned window_xor(unsigned x) {  // bit u of the result = x[u] ^ x[u-1] ^ .. ^ x[u-Q+1]
  const unsigned w2 = x ^ (x << 1), w4 = w2 ^ (w2 << 2), w8 = w4 ^ (w4 << 4);  // windows of 2, 4, 8 bits
  unsigned acc = 0;
  int off = 0;
  if (Q & 16) { acc ^= w8 ^ (w8 << 8); off += 16; }
  if (Q & 8) { acc ^= w8 << off; off += 8; }
  if (Q & 4) { acc ^= w4 << off; off += 4; }
  if (Q & 2) { acc ^= w2 << off; off += 2; }
  if (Q & 1) { acc ^= x << off; }
  return acc;
}
__device__ __forceinline__ unsigned compact16(uint4 w) {  // bit t = low bit of byte t of the 16-byte word
  const unsigned m = 0x01010101u, k = 0x01020408u;
  const unsigned c0 = ((w.x & m) * k) >> 24, c1 = ((w.y & m) * k) >> 24, c2 = ((w.z & m) * k) >> 24, c3 = ((w.w & m) * k) >> 24;
  return c0 | (c1 << 4) | (c2 << 8) | (c3 << 12);
}
template <int Q>
struct StreamPat {
  static constexpr unsigned rep() {
    unsigned r = 0;
    for (int k = 0; k * Q < 32; ++k) r |= 1u << (k * Q);
    return r;
  }
  static constexpr unsigned REP = rep();            // 1 at the positions u = 0 (mod Q)
  static constexpr unsigned END = rep() << (Q - 1); // 1 at the positions u = Q - 1 (mod Q): the last byte of a shot
};

template <int Q, int G>
__global__ void __launch_bounds__(256)
    moments_stream_kernel(int64_t B, int64_t S, const unsigned char* __restrict__ bits,
                          const uint32_t* __restrict__ colmask, const double* __restrict__ coeff, int prior, MomInv iv,
                          double* __restrict__ mean, double* __restrict__ var) {
  static_assert(Q >= 2 && Q <= 16, "a shot reaches at most 15 bytes back into the previous 16-byte word");
  const int gl = threadIdx.x % G;
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ngrp = ((int64_t)gridDim.x * blockDim.x) / G;
  const long long base = (long long)reinterpret_cast<uintptr_t>(bits);
  const int64_t rounds = (B + ngrp - 1) / ngrp;
  // the same trip count for every group of the warp (shuffles inside): the most words a setting can touch, in fours
  const int max_words = (int)((S * Q + 15 + 15) >> 4);
  const int trips = ((max_words + G - 1) / G + 3) / 4;
  constexpr int INC = (16 * G) % Q;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t b = grp + it * ngrp;
    const bool live = b < B;
    const unsigned cm = live ? (colmask[b] & ((1u << Q) - 1u)) : 0u;
    const long long lo = base + (live ? b : 0) * S * Q, hi = live ? lo + S * Q : lo;
    const unsigned cpat = cm * StreamPat<Q>::REP;  // column pattern along the stream (bits beyond 32 are never used)
    const long long W0 = lo & ~15LL;
    const int nwords = live ? (int)(((hi - 1 - W0) >> 4) + 1) : 0;
    int ph = (int)((W0 + 16LL * gl - lo) % Q);     // phase of byte 0 of this lane's first word (may be before lo)
    if (ph < 0) ph += Q;
    unsigned cnt = 0, carry = 0;
    for (int k4 = 0; k4 < trips; ++k4) {
      uint4 v[4];
      unsigned en[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = (4 * k4 + u) * G + gl;
        const long long a = W0 + 16LL * w;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        en[u] = 0u;
        if (w < nwords) {
          if (a >= lo && a + 16 <= hi) {
            v[u] = *reinterpret_cast<const uint4*>(a);
            en[u] = 0xffffu;
          } else {  // the first / last word of the setting: shared with its neighbours (or the end of the tensor)
            v[u] = ld16_inside(a, base, base + B * S * Q);
            const int b0 = (int)max(lo - a, 0LL), b1 = (int)min(hi - a, 16LL);
            en[u] = ((1u << b1) - 1u) & ~((1u << b0) - 1u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned c = compact16(v[u]) & (cpat >> ph) & en[u];
        unsigned prev = __shfl_up_sync(0xffffffffu, c, 1, G);
        if (gl == 0) prev = carry;                      // the word before lane 0's is lane G-1's of the previous step
        carry = __shfl_sync(0xffffffffu, c, G - 1, G);
        const unsigned y = window_xor<Q>(prev | (c << 16));
        cnt += __popc((y >> 16) & (StreamPat<Q>::END >> ph) & en[u]);
        ph += INC;
        if (ph >= Q) ph -= Q;
      }
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (live && gl == 0) moments_epilogue(tot, S, cm == 0, coeff[b], prior, iv, mean + b, var + b);
  }
}

// Any width up to 32 columns.  One warp per setting: the setting's bytes are staged through shared memory in chunks of
// whole shots with coalesced, aligned 16-byte loads (4 in flight per lane), then each lane takes the parity of the
// selected columns of its shots out of shared memory (widths up to 8: word reads + funnel shift + popcount).
constexpr int MOM_CHUNK = 2048;  // payload bytes per chunk (+ up to 30 bytes of alignment slack)
__global__ void __launch_bounds__(256)
    moments_bytes_kernel(int64_t B, int64_t S, int Q, const unsigned char* __restrict__ bits,
                         const uint32_t* __restrict__ colmask, const double* __restrict__ coeff, int prior, MomInv iv,
                         double* __restrict__ mean, double* __restrict__ var) {
  __shared__ __align__(16) unsigned char stage_all[8][MOM_CHUNK + 32];
  unsigned char* stage = stage_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const long long base = (long long)reinterpret_cast<uintptr_t>(bits);
  const int64_t shots_per_chunk = MOM_CHUNK / Q;
  for (int64_t b = warp; b < B; b += nwarps) {
    const unsigned cm = colmask[b] & ((Q >= 32) ? 0xffffffffu : ((1u << Q) - 1u));
    const long long lo = base + b * S * Q;
    unsigned pat_lo = 0, pat_hi = 0;  // 0x01 in byte i when column i (lo) / column 4 + i (hi) is selected
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pat_lo |= ((cm >> i) & 1u) << (8 * i);
      pat_hi |= ((cm >> (4 + i)) & 1u) << (8 * i);
    }
    unsigned cnt = 0;
    for (int64_t s0 = 0; s0 < S && cm; s0 += shots_per_chunk) {
      const int64_t ns = min(shots_per_chunk, S - s0);
      const long long c0 = lo + s0 * Q, c1 = c0 + ns * Q, a0 = c0 & ~15LL;
      const int nwords = (int)((c1 - a0 + 15) >> 4);  // <= (MOM_CHUNK + 30) / 16
      for (int w = lane; w < nwords; w += 128) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (w + 32 * u < nwords) v[u] = ld16_inside(a0 + 16LL * (w + 32 * u), base, base + B * S * Q);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (w + 32 * u < nwords) *reinterpret_cast<uint4*>(stage + 16 * (w + 32 * u)) = v[u];
      }
      __syncwarp();
      const int off = (int)(c0 - a0);
      if (Q <= 8) {
        // a shot is at most 8 bytes: two / three aligned 32-bit reads, funnel-shifted to the shot's first byte
        const unsigned* sw = reinterpret_cast<const unsigned*>(stage);
        for (int s = lane; s < ns; s += 32) {
          const int o = off + s * Q, i = o >> 2, sh = (o & 3) * 8;
          const unsigned w0 = sw[i], w1 = sw[i + 1];
          unsigned t = __funnelshift_r(w0, w1, sh) & pat_lo;
          if (Q > 4) t ^= __funnelshift_r(w1, sw[i + 2], sh) & pat_hi;
          cnt += __popc(t) & 1u;
        }
      } else {
        for (int s = lane; s < ns; s += 32) {
          unsigned par = 0;
          for (unsigned m = cm; m; m &= m - 1) par ^= stage[off + s * Q + (__ffs(m) - 1)];
          cnt += par & 1u;
        }
      }
      __syncwarp();
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) moments_epilogue(tot, S, cm == 0, coeff[b], prior, iv, mean + b, var + b);
  }
}

// corrected mean / variance of calibrate_observable_estimates: mean / cal_mean and ratio_variance
__global__ void calibrate_kernel(int64_t B, const double* __restrict__ mean, const double* __restrict__ var,
                                 const double* __restrict__ cal_mean, const double* __restrict__ cal_var,
                                 double* __restrict__ out_mean, double* __restrict__ out_var) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = mean[i], va = var[i], b = cal_mean[i], vb = cal_var[i];
    const double b2 = b * b;
    out_mean[i] = a / b;
    out_var[i] = va / b2 + (a * a * vb) / (b2 * b2);
  }
}

extern "C" int qt_shots_to_obs_moments_batch(int64_t B, int64_t n_shots, int n_qubits, const uint8_t* bits,
                                             const uint32_t* col_mask, const double* coeff, int use_beta_prior,
                                             double* mean_out, double* var_out, void* stream) {
  QT_REQUIRE(n_qubits >= 1 && n_qubits <= 32, "qt_shots_to_obs_moments_batch: n_qubits=%d out of range 1..32", n_qubits);
  QT_REQUIRE(n_shots >= 1, "qt_shots_to_obs_moments_batch: n_shots must be positive");
  if (B == 0) return QT_OK;
  QT_REQUIRE(bits && col_mask && coeff && mean_out && var_out, "qt_shots_to_obs_moments_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)std::min<int64_t>((B + 7) / 8, (int64_t)QT_NUM_SMS * 8);
  const double sn = (double)n_shots;
  const MomInv iv{1.0 / sn, 1.0 / (sn + 2.0), 1.0 / ((sn + 2.0) * (sn + 2.0) * (sn + 3.0))};
// lanes per setting from the number of 16-byte words of one setting
#define SWAR(Q)                                                                                                     \
  do {                                                                                                              \
    const int64_t words = (n_shots * Q + 15) / 16;                                                                  \
    if (words >= 256) {                                                                                             \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 7) / 8, (int64_t)QT_NUM_SMS * 8);                    \
      moments_swar_kernel<Q, 32><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,     \
                                                         mean_out, var_out);                                        \
    } else if (words >= 128) {                                                                                      \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 15) / 16, (int64_t)QT_NUM_SMS * 8);                  \
      moments_swar_kernel<Q, 16><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,     \
                                                         mean_out, var_out);                                        \
    } else {                                                                                                        \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 31) / 32, (int64_t)QT_NUM_SMS * 8);                  \
      moments_swar_kernel<Q, 8><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,      \
                                                        mean_out, var_out);                                         \
    }                                                                                                               \
  } while (0)
#define STREAM(Q)                                                                                                   \
  do {                                                                                                              \
    const int64_t words = (n_shots * Q + 15) / 16;                                                                  \
    if (words >= 256) {                                                                                             \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 7) / 8, (int64_t)QT_NUM_SMS * 8);                    \
      moments_stream_kernel<Q, 32><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,   \
                                                           mean_out, var_out);                                      \
    } else if (words >= 128) {                                                                                      \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 15) / 16, (int64_t)QT_NUM_SMS * 8);                  \
      moments_stream_kernel<Q, 16><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,   \
                                                           mean_out, var_out);                                      \
    } else {                                                                                                        \
      const unsigned blocks = (unsigned)std::min<int64_t>((B + 31) / 32, (int64_t)QT_NUM_SMS * 8);                  \
      moments_stream_kernel<Q, 8><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, iv,    \
                                                          mean_out, var_out);                                       \
    }                                                                                                               \
  } while (0)
  // the SWAR kernels need every shot aligned to its own width in flat addresses; the stream kernel does not
  const bool pow2 = n_qubits == 1 || n_qubits == 2 || n_qubits == 4 || n_qubits == 8;
  const bool aligned = reinterpret_cast<uintptr_t>(bits) % (uintptr_t)n_qubits == 0;
  if (n_qubits >= 2 && n_qubits <= 16 && !(pow2 && aligned)) {
    switch (n_qubits) {
      case 2: STREAM(2); break;
      case 3: STREAM(3); break;
      case 4: STREAM(4); break;
      case 5: STREAM(5); break;
      case 6: STREAM(6); break;
      case 7: STREAM(7); break;
      case 8: STREAM(8); break;
      case 9: STREAM(9); break;
      case 10: STREAM(10); break;
      case 11: STREAM(11); break;
      case 12: STREAM(12); break;
      case 13: STREAM(13); break;
      case 14: STREAM(14); break;
      case 15: STREAM(15); break;
      default: STREAM(16); break;
    }
    return qt_check_launch("moments_stream_kernel");
  }
  const int width = aligned ? n_qubits : 0;
  switch (width) {
    case 1: SWAR(1); break;
    case 2: SWAR(2); break;
    case 4: SWAR(4); break;
    case 8: SWAR(8); break;
    default:
      moments_bytes_kernel<<<blocks, 256, 0, st>>>(B, n_shots, n_qubits, bits, col_mask, coeff, use_beta_prior, iv,
                                                   mean_out, var_out);
  }
#undef SWAR
#undef STREAM
  return qt_check_launch("moments_kernel");
}

extern "C" int qt_calibrate_estimates_batch(int64_t B, const double* mean, const double* var, const double* cal_mean,
                                            const double* cal_var, double* mean_out, double* var_out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(mean && var && cal_mean && cal_var && mean_out && var_out, "qt_calibrate_estimates_batch: null argument");
  const unsigned blocks = (unsigned)std::min<int64_t>((B + 255) / 256, (int64_t)QT_NUM_SMS * 8);
  calibrate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(B, mean, var, cal_mean, cal_var, mean_out, var_out);
  return qt_check_launch("calibrate_kernel");
}
