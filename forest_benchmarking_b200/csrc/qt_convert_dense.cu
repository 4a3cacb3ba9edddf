// superop <-> Pauli-Liouville as the reference writes it: two DENSE complex matrix products with the 4^n x 4^n
// change-of-basis matrix (operator_tools/superoperator_transformations.py:253-264, 301-312, 374-438),
//     PL = c2p S c2p^dagger d = (1/d) G S G^dagger,     S = p2c PL p2c^dagger / d = (1/d) G^dagger PL G,
//     G = p2c^dagger,  G[i][c d + r] = conj(P_i[r][c])   (column-stacking vec, :33-51; entries in {0, +-1, +-i}),
// on the FP64 tensor path: mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05 has no FP64 kind).  A complex product is four
// real DMMA products on planar (re / im) operands held in shared memory.
//
// This is the "genuine dense contraction" variant the north_star allows tensor cores for.  It exists to be MEASURED
// against the Kronecker-factored butterfly kernels of qt_convert.cu (16 d^6 real FLOPs per matrix here, 4 n d^4 additions
// there): profiles/r02_ptm_dense_vs_butterfly.md has the timings and the tensor-pipe utilisation; the butterfly wins at
// every n and stays the default (variant 0 of qt_superop2pl_batch_variant).  n = 2, 3 only: at n = 1 a 4 x 4 product
// does not fill one 8 x 8 x 4 DMMA tile, at n >= 4 the dense form is 256x / 1024x more arithmetic than the factored one.
#include "qt_common.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

namespace {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int N>
struct DenseCfg {
  static constexpr int L = 1 << (2 * N);
  static constexpr int LDP = L + 4;            // == 4 (mod 16) doubles: both fragment access patterns are conflict-free
  static constexpr int TPD = L / 16;           // 16 x 16 warp tiles per dimension
  static constexpr int WPM = TPD * TPD;        // warps per matrix
  static constexpr int NW = 16;                // warps per block
  static constexpr int MPB = NW / WPM;         // matrices in flight per block
  static constexpr int PLANE = L * LDP;        // doubles per plane
  static constexpr size_t smem = sizeof(double) * PLANE * (2 + 4 * MPB);  // G re/im + (X re/im, T re/im) per matrix
};

// One 16 x 16 complex output tile of C = op_a(A) * op_b(B), K = L, on one warp (2 x 2 DMMA tiles of 8 x 8).
//   A_MODE 0: A[i][k] at i*LDP + k          A_MODE 1: conj(M[k][i]) at k*LDP + i
//   B_MODE 0: B[k][j] at k*LDP + j          B_MODE 1: conj(M[j][k]) at j*LDP + k
// acc[ti][tj][re/im][2]: lane holds C[8 ti + lane/4][8 tj + 2 (lane%4) + {0, 1}].
template <int L, int LDP, int A_MODE, int B_MODE>
__device__ __forceinline__ void warp_tile_cgemm(const double* __restrict__ Are, const double* __restrict__ Aim,
                                                const double* __restrict__ Bre, const double* __restrict__ Bim,
                                                int row0, int col0, int lane, double (&acc)[2][2][2][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ti = 0; ti < 2; ++ti)
#pragma unroll
    for (int tj = 0; tj < 2; ++tj)
#pragma unroll
      for (int c = 0; c < 2; ++c) acc[ti][tj][c][0] = acc[ti][tj][c][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < L; k0 += 4) {
    double ar[2], ai[2], nai[2], br[2], bi[2];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) {
      const int i = row0 + 8 * ti + g, k = k0 + t;
      const int off = A_MODE == 0 ? i * LDP + k : k * LDP + i;
      ar[ti] = Are[off];
      ai[ti] = A_MODE == 0 ? Aim[off] : -Aim[off];
      nai[ti] = -ai[ti];
    }
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      const int j = col0 + 8 * tj + g, k = k0 + t;
      const int off = B_MODE == 0 ? k * LDP + j : j * LDP + k;
      br[tj] = Bre[off];
      bi[tj] = B_MODE == 0 ? Bim[off] : -Bim[off];
    }
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) {
        dmma884(acc[ti][tj][0][0], acc[ti][tj][0][1], ar[ti], br[tj]);   // re += Ar Br
        dmma884(acc[ti][tj][0][0], acc[ti][tj][0][1], nai[ti], bi[tj]);  // re -= Ai Bi
        dmma884(acc[ti][tj][1][0], acc[ti][tj][1][1], ar[ti], bi[tj]);   // im += Ar Bi
        dmma884(acc[ti][tj][1][0], acc[ti][tj][1][1], ai[ti], br[tj]);   // im += Ai Br
      }
  }
}

template <int N, bool FWD>
__global__ void __launch_bounds__(512) pl_dense_dmma_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  using C = DenseCfg<N>;
  constexpr int L = C::L, LDP = C::LDP, D = 1 << N;
  extern __shared__ __align__(16) double sm[];
  double* Gre = sm;
  double* Gim = Gre + C::PLANE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / C::WPM, wig = warp % C::WPM;  // matrix slot of this warp, warp index inside the slot
  const int gt = (warp % C::WPM) * 32 + lane, gnt = C::WPM * 32;  // thread index / count inside the slot
  double* Xre = Gim + C::PLANE + (size_t)grp * 4 * C::PLANE;
  double* Xim = Xre + C::PLANE;
  double* Tre = Xim + C::PLANE;
  double* Tim = Tre + C::PLANE;
  auto group_sync = [&]() {
    if constexpr (C::WPM == 1) __syncwarp();
    else __syncthreads();
  };
  // G[i][c d + r] = conj(P_i[r][c]),  P_i[r][c] = [r ^ c == x] i^{|x & z|} (-1)^{|z & c|}
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
    const int i = e / L, v = e % L, c = v / D, r = v % D;
    const int x = pauli_xmask(i, N), z = pauli_zmask(i, N);
    double re = 0.0, im = 0.0;
    if ((r ^ c) == x) {
      const int ph = __popc(x & z) & 3;
      const double sg = (__popc(z & c) & 1) ? -1.0 : 1.0;
      re = (ph == 0) ? sg : (ph == 2 ? -sg : 0.0);
      im = (ph == 1) ? sg : (ph == 3 ? -sg : 0.0);
    }
    Gre[i * LDP + v] = re;
    Gim[i * LDP + v] = -im;  // conj
  }
  __syncthreads();
  const double scale = 1.0 / (double)D;
  const int row0 = (wig / C::TPD) * 16, col0 = (wig % C::TPD) * 16;
  const int g = lane >> 2, t = lane & 3;
  const int64_t n_iter = (B + (int64_t)gridDim.x * C::MPB - 1) / ((int64_t)gridDim.x * C::MPB);
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t b = (it * gridDim.x + blockIdx.x) * C::MPB + grp;
    const bool live = b < B;  // every warp keeps the block-wide barriers of the 64 x 64 case
    if (live) {
      const cplx* src = in + b * L * L;
      for (int e = gt; e < L * L; e += gnt) {
        const cplx v = src[e];
        Xre[(e / L) * LDP + e % L] = v.x;
        Xim[(e / L) * LDP + e % L] = v.y;
      }
    }
    group_sync();
    double acc[2][2][2][2];
    if (live) {
      // T = G X (FWD)  or  G^dagger X (!FWD)
      if (FWD) warp_tile_cgemm<L, LDP, 0, 0>(Gre, Gim, Xre, Xim, row0, col0, lane, acc);
      else warp_tile_cgemm<L, LDP, 1, 0>(Gre, Gim, Xre, Xim, row0, col0, lane, acc);
#pragma unroll
      for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) {
          const int off = (row0 + 8 * ti + g) * LDP + col0 + 8 * tj + 2 * t;
          *reinterpret_cast<double2*>(Tre + off) = make_double2(acc[ti][tj][0][0], acc[ti][tj][0][1]);
          *reinterpret_cast<double2*>(Tim + off) = make_double2(acc[ti][tj][1][0], acc[ti][tj][1][1]);
        }
    }
    group_sync();
    if (live) {
      // out = T G^dagger (FWD)  or  T G (!FWD), scaled by 1/d
      if (FWD) warp_tile_cgemm<L, LDP, 0, 1>(Tre, Tim, Gre, Gim, row0, col0, lane, acc);
      else warp_tile_cgemm<L, LDP, 0, 0>(Tre, Tim, Gre, Gim, row0, col0, lane, acc);
      cplx* dst = out + b * L * L;
#pragma unroll
      for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) {
          const int r = row0 + 8 * ti + g, c = col0 + 8 * tj + 2 * t;
          double4 v = make_double4(acc[ti][tj][0][0] * scale, acc[ti][tj][1][0] * scale, acc[ti][tj][0][1] * scale,
                                   acc[ti][tj][1][1] * scale);
          *reinterpret_cast<double4*>(dst + (size_t)r * L + c) = v;  // two consecutive complex numbers
        }
    }
    group_sync();
  }
}

template <int N, bool FWD>
int launch_dense(int64_t B, const void* in, void* out, cudaStream_t st) {
  using C = DenseCfg<N>;
  QT_CUDA(cudaFuncSetAttribute(pl_dense_dmma_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
  const int64_t blocks = std::min<int64_t>((B + C::MPB - 1) / C::MPB, (int64_t)QT_NUM_SMS);
  pl_dense_dmma_kernel<N, FWD><<<(unsigned)blocks, 512, C::smem, st>>>(B, (const cplx*)in, (cplx*)out);
  return qt_check_launch("pl_dense_dmma_kernel");
}

}  // namespace

extern "C" int qt_superop2pl_batch(int n, int64_t B, const void* superop, void* pl_out, void* workspace, void* stream);
extern "C" int qt_pl2superop_batch(int n, int64_t B, const void* pl, void* superop_out, void* workspace, void* stream);

extern "C" int qt_superop_pl_batch_variant(int n, int64_t B, const void* in, void* out, void* workspace, int forward,
                                           int variant, void* stream) {
  if (variant == QT_PL_VARIANT_BUTTERFLY)
    return forward ? qt_superop2pl_batch(n, B, in, out, workspace, stream)
                   : qt_pl2superop_batch(n, B, in, out, workspace, stream);
  QT_REQUIRE(variant == QT_PL_VARIANT_DENSE_DMMA, "qt_superop_pl_batch_variant: unknown variant %d", variant);
  if (B == 0) return QT_OK;
  QT_REQUIRE(in && out && in != out, "qt_superop_pl_batch_variant: bad arguments (out-of-place)");
  if (n != 2 && n != 3) {
    qt_set_error("the dense FP64-MMA PTM variant exists for n = 2, 3 only (got %d)", n);
    return QT_ERR_UNSUPPORTED;
  }
  if (B == 0) return QT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 2) return forward ? launch_dense<2, true>(B, in, out, st) : launch_dense<2, false>(B, in, out, st);
  return forward ? launch_dense<3, true>(B, in, out, st) : launch_dense<3, false>(B, in, out, st);
}
