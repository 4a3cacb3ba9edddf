// The step BEFORE the tomography path: raw shot data -> ExperimentResult.expectation / std_err.
//
// shots_to_obs_moments (observable_estimation.py:804-853): a bitarray [n_shots, n_qubits] of 0/1 bytes, the
// columns of the observable's qubits, eigenvalue product prod(1 - 2 b) per shot, then mean and variance of the
// mean (optionally through the Beta(n+ + 1, n- + 1) posterior).  Batched over settings: bits[B, n_shots, n_qubits].
// calibrate_observable_estimates' arithmetic (:1033-1049 + ratio_variance :1052-1090) is the elementwise kernel below.
//
// Integer / byte work, HBM-bound: one byte per shot and qubit is read once, two doubles per setting are written.
// One warp per setting.  For n_qubits in {1, 2, 4, 8} the whole array is read as one flat stream of aligned 16-byte
// words (a setting's byte range may start anywhere: the bytes of neighbouring settings in its first / last word are
// masked off), and the per-shot parity is folded inside the 32-bit words (SWAR) before one popcount per word.
// Other widths take the byte-per-lane kernel.
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

// bytes of the u32 at flat byte address `a` that lie inside [lo, hi)
__device__ __forceinline__ unsigned range_mask(long long a, long long lo, long long hi) {
  unsigned m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (a + i >= lo && a + i < hi) m |= 0xffu << (8 * i);
  return m;
}

__device__ __forceinline__ void moments_epilogue(unsigned long long n_minus, long long S, bool identity, double coeff,
                                                 int prior, double* mean, double* var) {
  if (identity) {  // identity term (:826-827)
    *mean = coeff;
    *var = 0.0;
    return;
  }
  const double nm = (double)n_minus, np_ = (double)(S - (long long)n_minus), n = (double)S;
  if (prior) {  // Beta(n+ + 1, n- + 1) mean / variance, then bit -> Pauli moments (:838-847, utils.py:446-458)
    const double a = np_ + 1.0, b = nm + 1.0, ab = a + b;
    const double bm = a / ab, bv = a * b / (ab * ab * (ab + 1.0));
    *mean = (2.0 * bm - 1.0) * coeff;
    *var = 4.0 * bv * coeff * coeff;
  } else {  // np.mean / np.var of the +-coeff values, variance of the mean (:849-851)
    const double m = coeff * (np_ - nm) / n;
    const double dp = coeff - m, dm = -coeff - m;
    *mean = m;
    *var = (np_ * dp * dp + nm * dm * dm) / n / n;
  }
}

template <int Q>
__global__ void __launch_bounds__(256)
    moments_swar_kernel(int64_t B, int64_t S, const unsigned char* __restrict__ bits,
                        const uint32_t* __restrict__ colmask, const double* __restrict__ coeff, int prior,
                        double* __restrict__ mean, double* __restrict__ var) {
  static_assert(Q == 1 || Q == 2 || Q == 4 || Q == 8, "shots must tile a 32-bit word");
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const long long base = (long long)reinterpret_cast<uintptr_t>(bits);  // flat addresses: the alignment of `bits` is free
  for (int64_t b = warp; b < B; b += nwarps) {
    const unsigned cm = colmask[b] & ((Q >= 32) ? 0xffffffffu : ((1u << Q) - 1u));
    const long long lo = base + b * S * Q, hi = lo + S * Q;
    const unsigned pat0 = col_pattern(cm, 0, Q), pat1 = col_pattern(cm, 4, Q);
    unsigned cnt = 0;
    for (long long a = (lo & ~15LL) + 16LL * lane; a < hi; a += 16LL * 32) {
      const uint4 w = *reinterpret_cast<const uint4*>(a);
      unsigned x0 = w.x & pat0, x1 = w.y & ((Q == 8) ? pat1 : pat0), x2 = w.z & pat0, x3 = w.w & ((Q == 8) ? pat1 : pat0);
      if (a < lo || a + 16 > hi) {  // first / last word of the setting: drop the neighbours' bytes
        x0 &= range_mask(a, lo, hi);
        x1 &= range_mask(a + 4, lo, hi);
        x2 &= range_mask(a + 8, lo, hi);
        x3 &= range_mask(a + 12, lo, hi);
      }
      if (Q == 1) {
        cnt += __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3);
      } else if (Q == 2) {
        x0 ^= x0 >> 8; x1 ^= x1 >> 8; x2 ^= x2 >> 8; x3 ^= x3 >> 8;
        cnt += __popc(x0 & 0x00010001u) + __popc(x1 & 0x00010001u) + __popc(x2 & 0x00010001u) + __popc(x3 & 0x00010001u);
      } else if (Q == 4) {
        x0 ^= x0 >> 16; x1 ^= x1 >> 16; x2 ^= x2 >> 16; x3 ^= x3 >> 16;
        x0 ^= x0 >> 8; x1 ^= x1 >> 8; x2 ^= x2 >> 8; x3 ^= x3 >> 8;
        cnt += (x0 & 1u) + (x1 & 1u) + (x2 & 1u) + (x3 & 1u);
      } else {
        unsigned y0 = x0 ^ x1, y1 = x2 ^ x3;
        y0 ^= y0 >> 16; y1 ^= y1 >> 16;
        y0 ^= y0 >> 8; y1 ^= y1 >> 8;
        cnt += (y0 & 1u) + (y1 & 1u);
      }
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) moments_epilogue(tot, S, cm == 0, coeff[b], prior, mean + b, var + b);
  }
}

// any width up to 32 columns: one shot per lane, only the selected columns are read
__global__ void __launch_bounds__(256)
    moments_bytes_kernel(int64_t B, int64_t S, int Q, const unsigned char* __restrict__ bits,
                         const uint32_t* __restrict__ colmask, const double* __restrict__ coeff, int prior,
                         double* __restrict__ mean, double* __restrict__ var) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    const unsigned cm = colmask[b] & ((Q >= 32) ? 0xffffffffu : ((1u << Q) - 1u));
    const unsigned char* src = bits + b * S * Q;
    unsigned cnt = 0;
    for (int64_t s = lane; s < S; s += 32) {
      unsigned par = 0;
      for (unsigned m = cm; m; m &= m - 1) par ^= src[s * Q + (__ffs(m) - 1)];
      cnt += par & 1u;
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) moments_epilogue(tot, S, cm == 0, coeff[b], prior, mean + b, var + b);
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
#define SWAR(Q)                                                                                                    \
  moments_swar_kernel<Q><<<blocks, 256, 0, st>>>(B, n_shots, bits, col_mask, coeff, use_beta_prior, mean_out, var_out)
  // the SWAR kernels need every shot aligned to its own width in flat addresses
  const int width = (reinterpret_cast<uintptr_t>(bits) % (uintptr_t)n_qubits == 0) ? n_qubits : 0;
  switch (width) {
    case 1: SWAR(1); break;
    case 2: SWAR(2); break;
    case 4: SWAR(4); break;
    case 8: SWAR(8); break;
    default:
      moments_bytes_kernel<<<blocks, 256, 0, st>>>(B, n_shots, n_qubits, bits, col_mask, coeff, use_beta_prior,
                                                   mean_out, var_out);
  }
#undef SWAR
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
