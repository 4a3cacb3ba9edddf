// Batched superoperator representation changes (HBM-bound streaming kernels).
//
//   kraus2choi        operator_tools/superoperator_transformations.py:159-182   sum_k vec(K) vec(K)^dagger
//   kraus2superop     :100-145                                                  sum_k conj(K) (x) K
//   choi <-> superop  :267-277, :351-361     reshape [d]*4, swapaxes(0,3) -- a pure index permutation
//   superop <-> PTM   :253-264, :301-312     (1/d) F S F^dagger and (1/d) F^dagger R F, where
//        F[a, r] = conj(vec(P_a)[r]) is the computational->Pauli change of basis the reference rebuilds
//        densely on every call (:374-438).  F is a Kronecker power: per qubit a 4-point ADD-ONLY
//        butterfly on the bit pair (i_q, j_q) of the vec index r = j*d + i:
//            I: u00+u11   X: u01+u10   Y: i(u01-u10)   Z: u00-u11
//        followed by a bit un-interleave to the canonical base-4 Pauli index.
//
// All matrices are [B, d^2, d^2] interleaved complex128 row-major.  Reads and writes are fully coalesced
// 16-byte accesses; permutations and butterflies go through shared memory.
#include "qt_common.cuh"
#include "qt_pauli.cuh"
#include "../../include/qtomo.h"

static constexpr int TILE_ELEMS = 4096;  // 64 KB of complex128 per block pass

// ---------------------------------------------------------------------------------------------
// kraus2choi / kraus2superop
// ---------------------------------------------------------------------------------------------
// MODE 0: choi[r, c] = sum_k v_k[r] conj(v_k[c]),  v_k[j*d + i] = K_k[i, j]
// MODE 1: superop[(i1,i2),(j1,j2)] = sum_k conj(K_k[i1, j1]) K_k[i2, j2]
template <int MODE>
__global__ void kraus_outer_kernel(int d, int nk, int64_t B, const cplx* __restrict__ kraus, cplx* __restrict__ out,
                                   int items_per_block, int rows_per_tile) {
  extern __shared__ __align__(16) cplx ks[];  // [items_per_block][nk][d*d]  (vec order for MODE 0)
  const int d2 = d * d;
  const int tiles = (d2 + rows_per_tile - 1) / rows_per_tile;
  const int64_t unit = blockIdx.x;
  const int64_t b0 = (unit / tiles) * items_per_block;
  const int row0 = (int)(unit % tiles) * rows_per_tile;
  const int nb = (int)min((int64_t)items_per_block, B - b0);
  // stage the Kraus operators of these items (coalesced read, transposed into vec order for MODE 0)
  for (int e = threadIdx.x; e < nb * nk * d2; e += blockDim.x) {
    const int within = e % d2, which = e / d2;
    cplx v = kraus[b0 * nk * d2 + e];
    if (MODE == 0) {
      const int i = within / d, j = within % d;
      ks[which * d2 + j * d + i] = v;
    } else {
      ks[which * d2 + within] = v;
    }
  }
  __syncthreads();
  const int nrows = min(rows_per_tile, d2 - row0);
  for (int e = threadIdx.x; e < nb * nrows * d2; e += blockDim.x) {
    const int c = e % d2;
    const int rr = (e / d2) % nrows;
    const int bi = e / (d2 * nrows);
    const int r = row0 + rr;
    const cplx* kb = ks + bi * nk * d2;
    cplx acc = cmake(0.0, 0.0);
    if (MODE == 0) {
      for (int k = 0; k < nk; ++k) cfma_conj(acc, kb[k * d2 + r], kb[k * d2 + c]);
    } else {
      const int i1 = r / d, i2 = r % d, j1 = c / d, j2 = c % d;
      for (int k = 0; k < nk; ++k) cfma_conj(acc, kb[k * d2 + i2 * d + j2], kb[k * d2 + i1 * d + j1]);
    }
    out[((b0 + bi) * d2 + r) * d2 + c] = acc;
  }
}

static int launch_kraus_outer(int mode, int d, int nk, int64_t B, const void* kraus, void* out, cudaStream_t st) {
  const int d2 = d * d;
  const int64_t d4 = (int64_t)d2 * d2;
  int ipb = 1, rpt = d2;
  if (d4 <= TILE_ELEMS) ipb = (int)(TILE_ELEMS / d4);
  else rpt = max(1, 4 * TILE_ELEMS / d2);
  // keep the staged Kraus operators within 96 KB
  while (ipb > 1 && (size_t)ipb * nk * d2 * sizeof(cplx) > 96 * 1024) ipb /= 2;
  const size_t smem = (size_t)ipb * nk * d2 * sizeof(cplx);
  QT_REQUIRE(smem <= 200 * 1024, "kraus2choi/superop: %d Kraus operators of dimension %d exceed shared memory", nk, d);
  const int tiles = (d2 + rpt - 1) / rpt;
  const int64_t units = ((B + ipb - 1) / ipb) * tiles;
  if (mode == 0) {
    QT_CUDA(cudaFuncSetAttribute(kraus_outer_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kraus_outer_kernel<0><<<(unsigned)units, 256, smem, st>>>(d, nk, B, (const cplx*)kraus, (cplx*)out, ipb, rpt);
  } else {
    QT_CUDA(cudaFuncSetAttribute(kraus_outer_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kraus_outer_kernel<1><<<(unsigned)units, 256, smem, st>>>(d, nk, B, (const cplx*)kraus, (cplx*)out, ipb, rpt);
  }
  return qt_check_launch("kraus_outer_kernel");
}

extern "C" int qt_kraus2choi_batch(int d, int n_kraus, int64_t B, const void* kraus, void* choi_out, void* stream) {
  QT_REQUIRE(d >= 1 && d <= 32 && n_kraus >= 1 && kraus && choi_out, "qt_kraus2choi_batch: bad arguments");
  if (B == 0) return QT_OK;
  return launch_kraus_outer(0, d, n_kraus, B, kraus, choi_out, (cudaStream_t)stream);
}

extern "C" int qt_kraus2superop_batch(int d, int n_kraus, int64_t B, const void* kraus, void* superop_out,
                                      void* stream) {
  QT_REQUIRE(d >= 1 && d <= 32 && n_kraus >= 1 && kraus && superop_out, "qt_kraus2superop_batch: bad arguments");
  if (B == 0) return QT_OK;
  return launch_kraus_outer(1, d, n_kraus, B, kraus, superop_out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// choi <-> superop reshuffle:  out[(i0,i1),(i2,i3)] = in[(i3,i1),(i2,i0)]
// A work unit is (item, i1, chunk of i2): the d x C2 x d block in[(i3,i1),(i2,i0)] is read with i0
// fastest (runs of d contiguous elements, C2*d when C2 == d), transposed in shared memory and written
// with i3 fastest.  Several units per block when a unit is smaller than the 64 KB tile.
// ---------------------------------------------------------------------------------------------
__global__ void reshuffle_kernel(int d, int c2, int64_t n_units, const cplx* __restrict__ in,
                                 cplx* __restrict__ out, int units_per_block) {
  extern __shared__ __align__(16) cplx tile[];
  const int d2 = d * d;
  const int chunks = d / c2;
  const int unit_elems = d * c2 * d;
  const int row_stride = c2 * d + 1;  // +1 element of padding: conflict-free transposed reads
  const int unit_smem = d * row_stride;
  const int64_t u0 = (int64_t)blockIdx.x * units_per_block;
  const int nu = (int)min((int64_t)units_per_block, n_units - u0);
  for (int e = threadIdx.x; e < nu * unit_elems; e += blockDim.x) {
    const int ul = e / unit_elems, w = e % unit_elems;
    const int i0 = w % d, i2l = (w / d) % c2, i3 = w / (d * c2);
    const int64_t u = u0 + ul;
    const int chunk = (int)(u % chunks), i1 = (int)((u / chunks) % d);
    const int64_t b = u / ((int64_t)chunks * d);
    const int i2 = chunk * c2 + i2l;
    tile[ul * unit_smem + i3 * row_stride + i2l * d + i0] =
        in[(b * d2 + (i3 * d + i1)) * d2 + i2 * d + i0];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nu * unit_elems; e += blockDim.x) {
    const int ul = e / unit_elems, w = e % unit_elems;
    const int i3 = w % d, i2l = (w / d) % c2, i0 = w / (d * c2);
    const int64_t u = u0 + ul;
    const int chunk = (int)(u % chunks), i1 = (int)((u / chunks) % d);
    const int64_t b = u / ((int64_t)chunks * d);
    const int i2 = chunk * c2 + i2l;
    out[(b * d2 + (i0 * d + i1)) * d2 + i2 * d + i3] = tile[ul * unit_smem + i3 * row_stride + i2l * d + i0];
  }
}

extern "C" int qt_choi_superop_reshuffle_batch(int d, int64_t B, const void* in, void* out, void* stream) {
  QT_REQUIRE(d >= 1 && d <= 32 && (d & (d - 1)) == 0 && in && out && in != out,
             "qt_choi_superop_reshuffle_batch: bad arguments (d power of two <= 32, out-of-place)");
  if (B == 0) return QT_OK;
  int c2 = d;
  while ((int64_t)d * c2 * d > TILE_ELEMS) c2 /= 2;
  const int unit_elems = d * c2 * d;
  const int upb = max(1, TILE_ELEMS / unit_elems);
  const int64_t n_units = B * d * (d / c2);
  const size_t smem = (size_t)upb * d * (c2 * d + 1) * sizeof(cplx);
  QT_CUDA(cudaFuncSetAttribute(reshuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (n_units + upb - 1) / upb;
  reshuffle_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(d, c2, n_units, (const cplx*)in,
                                                                           (cplx*)out, upb);
  return qt_check_launch("reshuffle_kernel");
}

// ---------------------------------------------------------------------------------------------
// superop <-> Pauli-Liouville (PTM)
// ---------------------------------------------------------------------------------------------
// Fused kernel for n <= 3: whole matrices in shared memory.
// FWD: out = (1/d) F S F^dagger, rows/cols permuted to canonical Pauli order.
// !FWD: out = (1/d) F^dagger R F, input rows/cols gathered from canonical Pauli order.
template <bool FWD>
__global__ void pl_fused_kernel(int n, int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out,
                                int items_per_block) {
  extern __shared__ __align__(16) cplx buf[];
  const int L = 1 << (2 * n);
  const int ld = L + 1;  // padded leading dimension
  const int msz = L * ld;
  const int64_t b0 = (int64_t)blockIdx.x * items_per_block;
  const int nb = (int)min((int64_t)items_per_block, B - b0);
  const double scale = 1.0 / (double)(1 << n);
  for (int e = threadIdx.x; e < nb * L * L; e += blockDim.x) {
    const int bi = e / (L * L), r = (e / L) % L, c = e % L;
    const cplx v = in[b0 * L * L + e];
    if (FWD) buf[bi * msz + r * ld + c] = v;
    else buf[bi * msz + pauli_to_pos(r, n) * ld + pauli_to_pos(c, n)] = v;
  }
  __syncthreads();
  // along columns index (within each row): right factor.  FWD: Y = X F^dagger -> conj butterfly.
  pauli_butterfly_smem<FWD, true>(buf, n, nb * L, /*vstride*/ ld, /*estride*/ 1, threadIdx.x, blockDim.x);
  // along row index (for each column): left factor.  vector v = (item, column)
  for (int bi = 0; bi < nb; ++bi)
    pauli_butterfly_smem<FWD, false>(buf + bi * msz, n, L, /*vstride*/ 1, /*estride*/ ld, threadIdx.x, blockDim.x);
  for (int e = threadIdx.x; e < nb * L * L; e += blockDim.x) {
    const int bi = e / (L * L), r = (e / L) % L, c = e % L;
    cplx v;
    if (FWD) v = buf[bi * msz + pauli_to_pos(r, n) * ld + pauli_to_pos(c, n)];
    else v = buf[bi * msz + r * ld + c];
    out[b0 * L * L + e] = cscale(v, scale);
  }
}

// Two-pass path for n >= 4 (matrix larger than shared memory).
// Pass ROWS: transform along the contiguous (column-index) axis, `rows_per_block` full rows per block.
template <bool FWD>
__global__ void pl_rows_kernel(int n, int64_t total_rows, const cplx* __restrict__ in, cplx* __restrict__ out,
                               int rows_per_block) {
  extern __shared__ __align__(16) cplx buf[];
  const int L = 1 << (2 * n);
  const int ld = L + 1;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int nr = (int)min((int64_t)rows_per_block, total_rows - r0);
  for (int e = threadIdx.x; e < nr * L; e += blockDim.x) {
    const int rr = e / L, c = e % L;
    const cplx v = in[r0 * L + e];
    buf[rr * ld + (FWD ? c : pauli_to_pos(c, n))] = v;
  }
  __syncthreads();
  pauli_butterfly_smem<FWD, true>(buf, n, nr, ld, 1, threadIdx.x, blockDim.x);
  for (int e = threadIdx.x; e < nr * L; e += blockDim.x) {
    const int rr = e / L, c = e % L;
    out[r0 * L + e] = buf[rr * ld + (FWD ? pauli_to_pos(c, n) : c)];
  }
}
// Pass COLS: transform along the row-index axis for a panel of `w` columns (all L rows) of one item,
// row permutation folded into the global addressing; applies the 1/d scale.
template <bool FWD>
__global__ void pl_cols_kernel(int n, int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, int w) {
  extern __shared__ __align__(16) cplx buf[];
  const int L = 1 << (2 * n);
  const int panels = L / w;
  const int64_t b = blockIdx.x / panels;
  const int c0 = (int)(blockIdx.x % panels) * w;
  const int ld = w + 1;
  const double scale = 1.0 / (double)(1 << n);
  const cplx* src = in + b * L * L;
  cplx* dst = out + b * L * L;
  for (int e = threadIdx.x; e < L * w; e += blockDim.x) {
    const int r = e / w, c = e % w;
    buf[(FWD ? r : pauli_to_pos(r, n)) * ld + c] = src[(int64_t)r * L + c0 + c];
  }
  __syncthreads();
  pauli_butterfly_smem<FWD, false>(buf, n, w, /*vstride*/ 1, /*estride*/ ld, threadIdx.x, blockDim.x);
  for (int e = threadIdx.x; e < L * w; e += blockDim.x) {
    const int r = e / w, c = e % w;
    dst[(int64_t)r * L + c0 + c] = cscale(buf[(FWD ? pauli_to_pos(r, n) : r) * ld + c], scale);
  }
}

template <bool FWD>
static int launch_pl(int n, int64_t B, const void* in, void* out, void* workspace, cudaStream_t st) {
  const int L = 1 << (2 * n);
  if (n <= 3) {
    const int ipb = max(1, TILE_ELEMS / (L * L));
    const size_t smem = (size_t)ipb * L * (L + 1) * sizeof(cplx);
    QT_CUDA(cudaFuncSetAttribute(pl_fused_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pl_fused_kernel<FWD><<<(unsigned)((B + ipb - 1) / ipb), 256, smem, st>>>(n, B, (const cplx*)in, (cplx*)out, ipb);
    return qt_check_launch("pl_fused_kernel");
  }
  QT_REQUIRE(workspace, "superop<->pauli_liouville with n >= 4 needs a workspace of B*16^n*16 bytes");
  const int rpb = max(1, TILE_ELEMS / L);
  const size_t smem_r = (size_t)rpb * (L + 1) * sizeof(cplx);
  QT_CUDA(cudaFuncSetAttribute(pl_rows_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
  const int64_t total_rows = B * L;
  pl_rows_kernel<FWD><<<(unsigned)((total_rows + rpb - 1) / rpb), 256, smem_r, st>>>(n, total_rows, (const cplx*)in,
                                                                                     (cplx*)workspace, rpb);
  int rc = qt_check_launch("pl_rows_kernel");
  if (rc) return rc;
  const int w = max(1, 2 * TILE_ELEMS / L);  // n=4: 32 columns (512 B runs), n=5: 8 columns (128 B runs)
  const size_t smem_c = (size_t)L * (w + 1) * sizeof(cplx);
  QT_CUDA(cudaFuncSetAttribute(pl_cols_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  pl_cols_kernel<FWD><<<(unsigned)(B * (L / w)), 256, smem_c, st>>>(n, B, (const cplx*)workspace, (cplx*)out, w);
  return qt_check_launch("pl_cols_kernel");
}

extern "C" int qt_superop2pl_batch(int n, int64_t B, const void* superop, void* pl_out, void* workspace,
                                   void* stream) {
  QT_REQUIRE(n >= 1 && n <= 5 && superop && pl_out, "qt_superop2pl_batch: bad arguments");
  if (B == 0) return QT_OK;
  return launch_pl<true>(n, B, superop, pl_out, workspace, (cudaStream_t)stream);
}

extern "C" int qt_pl2superop_batch(int n, int64_t B, const void* pl, void* superop_out, void* workspace,
                                   void* stream) {
  QT_REQUIRE(n >= 1 && n <= 5 && pl && superop_out, "qt_pl2superop_batch: bad arguments");
  if (B == 0) return QT_OK;
  return launch_pl<false>(n, B, pl, superop_out, workspace, (cudaStream_t)stream);
}
