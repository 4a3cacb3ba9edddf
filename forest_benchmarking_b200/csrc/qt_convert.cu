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

#include <algorithm>

// elements per block of the small-n tiles: 16 KB tiles keep 8 blocks resident per SM (cf. proj_tp_kernel)
#ifndef QT_PL_TILE
#define QT_PL_TILE 1024
#endif
#ifndef QT_RESHUFFLE_TILE
#define QT_RESHUFFLE_TILE 1024
#endif

static constexpr int TILE_ELEMS = 4096;  // 64 KB of complex128 per block pass
#ifndef QT_PL3_REGISTER_KERNEL
#define QT_PL3_REGISTER_KERNEL 1  // n = 3 superop <-> PTM: register-resident radix-16 kernel (0: shared-memory stages)
#endif
#ifndef QT_PL_FUSED_PASSES
#define QT_PL_FUSED_PASSES 1  // n = 4, 5: both passes in one launch, intermediate in an L2-resident ring (0: two kernels)
#endif

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

// kraus2choi fast path for d = 2^LOGD: a block produces 4096 output elements (64 KB), consecutive threads write
// consecutive elements, and all index math is compile-time shifts and masks.
template <int LOGD>
struct K2cCfg {
  static constexpr int D = 1 << LOGD, D2 = D * D;
  static constexpr int64_t O = (int64_t)D2 * D2;
  static constexpr int OUT_PER_BLOCK = (D2 >= 1024) ? 16384 : 4096;  // n = 5: amortise the 32 KB Kraus staging
  static constexpr int IPB = (O >= OUT_PER_BLOCK) ? 1 : (int)(OUT_PER_BLOCK / O);
  static constexpr int ROWS = (O >= OUT_PER_BLOCK) ? OUT_PER_BLOCK / D2 : D2;
  static constexpr int RCHUNKS = D2 / ROWS;
  static constexpr int UNITS = IPB * ROWS * (D2 / 4);
};

template <int LOGD>
__global__ void __launch_bounds__(256) kraus2choi_kernel(int nk, int64_t B, const cplx* __restrict__ kraus,
                                                         cplx* __restrict__ out) {
  using C = K2cCfg<LOGD>;
  constexpr int D = C::D, D2 = C::D2, ROWS = C::ROWS;
  extern __shared__ __align__(16) cplx ks[];  // [IPB][nk][D2] in vec order: v_k[j*D + i] = K_k[i][j]
  const int64_t b0 = (int64_t)(blockIdx.x / C::RCHUNKS) * C::IPB;
  const int row0 = (int)(blockIdx.x % C::RCHUNKS) * ROWS;
  const int nb = (int)min((int64_t)C::IPB, B - b0);
  for (int e = threadIdx.x; e < nb * nk * D2; e += 256) {
    const int within = e % D2, which = e / D2;
    if (D >= 16) {
      // transpose on the (L2-resident) read side: a transposed shared-memory write would put 32 lanes on one bank
      ks[e] = kraus[b0 * nk * D2 + which * D2 + (within % D) * D + within / D];
    } else {
      ks[which * D2 + (within % D) * D + within / D] = kraus[b0 * nk * D2 + e];
    }
  }
  __syncthreads();
  // consecutive threads -> consecutive output elements (fully coalesced 16-byte stores)
  constexpr int PER = C::IPB * ROWS * D2 / 256;
#pragma unroll 4
  for (int k = 0; k < PER; ++k) {
    const int e = threadIdx.x + k * 256;
    const int c = e % D2, rr = (e / D2) % ROWS, bi = e / (ROWS * D2);
    if (bi < nb) {
      const int r = row0 + rr;
      const cplx* kb = ks + bi * nk * D2;
      cplx acc = cmake(0.0, 0.0);
      for (int q = 0; q < nk; ++q) cfma_conj(acc, kb[q * D2 + r], kb[q * D2 + c]);
      out[((b0 + bi) * D2 + r) * D2 + c] = acc;
    }
  }
}

template <int LOGD>
static int launch_kraus2choi_fast(int nk, int64_t B, const void* kraus, void* out, cudaStream_t st) {
  using C = K2cCfg<LOGD>;
  const size_t smem = sizeof(cplx) * C::IPB * nk * C::D2;
  QT_CUDA(cudaFuncSetAttribute(kraus2choi_kernel<LOGD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = ((B + C::IPB - 1) / C::IPB) * C::RCHUNKS;
  kraus2choi_kernel<LOGD><<<(unsigned)blocks, 256, smem, st>>>(nk, B, (const cplx*)kraus, (cplx*)out);
  return qt_check_launch("kraus2choi_kernel");
}

static int launch_kraus_outer(int mode, int d, int nk, int64_t B, const void* kraus, void* out, cudaStream_t st) {
  const int d2 = d * d;
  if (mode == 0 && (d & (d - 1)) == 0 && d >= 2) {
    const int ipb = (d2 * d2 >= 4096) ? 1 : 4096 / (d2 * d2);  // K2cCfg::IPB
    if ((size_t)ipb * nk * d2 * sizeof(cplx) <= 96 * 1024) {
      switch (d) {
        case 2: return launch_kraus2choi_fast<1>(nk, B, kraus, out, st);
        case 4: return launch_kraus2choi_fast<2>(nk, B, kraus, out, st);
        case 8: return launch_kraus2choi_fast<3>(nk, B, kraus, out, st);
        case 16: return launch_kraus2choi_fast<4>(nk, B, kraus, out, st);
        case 32: return launch_kraus2choi_fast<5>(nk, B, kraus, out, st);
        default: break;
      }
    }
  }
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
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(d >= 1 && d <= 32 && n_kraus >= 1 && kraus && choi_out, "qt_kraus2choi_batch: bad arguments");
  return launch_kraus_outer(0, d, n_kraus, B, kraus, choi_out, (cudaStream_t)stream);
}

extern "C" int qt_kraus2superop_batch(int d, int n_kraus, int64_t B, const void* kraus, void* superop_out,
                                      void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(d >= 1 && d <= 32 && n_kraus >= 1 && kraus && superop_out, "qt_kraus2superop_batch: bad arguments");
  return launch_kraus_outer(1, d, n_kraus, B, kraus, superop_out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// choi <-> superop reshuffle:  out[(i0,i1),(i2,i3)] = in[(i3,i1),(i2,i0)]   (d = 2^LOGD)
// A work unit is (item, i1, chunk of C2 values of i2): the d x C2 x d block in[(i3,i1),(i2,i0)] is read with
// (i2,i0) fastest (runs of C2*d contiguous elements), transposed through padded shared memory and written with
// (i2,i3) fastest.  Everything is a compile-time power of two: index math is shifts and masks.
// ---------------------------------------------------------------------------------------------
template <int LOGD>
struct ReshuffleCfg {
  static constexpr int D = 1 << LOGD;
  static constexpr int TILE = (LOGD <= 2) ? QT_RESHUFFLE_TILE : 2048;  // elements per block
  static constexpr int C2 = (TILE / (D * D) >= D) ? D : (TILE / (D * D) >= 1 ? TILE / (D * D) : 1);
  static constexpr int UNIT = D * C2 * D;                             // elements per unit
  static constexpr int UPB = (TILE / UNIT >= 1) ? TILE / UNIT : 1;    // units per block
  static constexpr int RS = C2 * D + 1;                               // padded row stride in shared memory
  static constexpr int CHUNKS = D / C2;
  static constexpr size_t smem = sizeof(cplx) * UPB * D * RS;
  static constexpr int NT = 256;
};

template <int LOGD>
__global__ void __launch_bounds__(256) reshuffle_kernel(int64_t n_units, const cplx* __restrict__ in,
                                                        cplx* __restrict__ out) {
  using C = ReshuffleCfg<LOGD>;
  constexpr int D = C::D, C2 = C::C2, UNIT = C::UNIT, RS = C::RS, CHUNKS = C::CHUNKS;
  extern __shared__ __align__(16) cplx tile[];
  const int64_t u0 = (int64_t)blockIdx.x * C::UPB;
  const int nu = (int)min((int64_t)C::UPB, n_units - u0);
  constexpr int PER = (C::UPB * UNIT + C::NT - 1) / C::NT;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int e = threadIdx.x + k * C::NT;
    if (e < nu * UNIT) {
      const int ul = e / UNIT, w = e % UNIT;
      const int i0 = w % D, i2l = (w / D) % C2, i3 = w / (D * C2);
      const int64_t u = u0 + ul;
      const int chunk = (int)(u % CHUNKS), i1 = (int)((u / CHUNKS) % D);
      const int64_t b = u / (CHUNKS * D);
      const int i2 = chunk * C2 + i2l;
      tile[ul * (D * RS) + i3 * RS + i2l * D + i0] = in[((b * D + i3) * D + i1) * (D * D) + i2 * D + i0];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int e = threadIdx.x + k * C::NT;
    if (e < nu * UNIT) {
      const int ul = e / UNIT, w = e % UNIT;
      const int i3 = w % D, i2l = (w / D) % C2, i0 = w / (D * C2);
      const int64_t u = u0 + ul;
      const int chunk = (int)(u % CHUNKS), i1 = (int)((u / CHUNKS) % D);
      const int64_t b = u / (CHUNKS * D);
      const int i2 = chunk * C2 + i2l;
      out[((b * D + i0) * D + i1) * (D * D) + i2 * D + i3] = tile[ul * (D * RS) + i3 * RS + i2l * D + i0];
    }
  }
}

template <int LOGD>
static int launch_reshuffle(int64_t B, const void* in, void* out, cudaStream_t st) {
  using C = ReshuffleCfg<LOGD>;
  const int64_t n_units = B * C::D * C::CHUNKS;
  const int64_t blocks = (n_units + C::UPB - 1) / C::UPB;
  reshuffle_kernel<LOGD><<<(unsigned)blocks, C::NT, C::smem, st>>>(n_units, (const cplx*)in, (cplx*)out);
  return qt_check_launch("reshuffle_kernel");
}

extern "C" int qt_choi_superop_reshuffle_batch(int d, int64_t B, const void* in, void* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(d >= 2 && d <= 32 && (d & (d - 1)) == 0 && in && out && in != out,
             "qt_choi_superop_reshuffle_batch: bad arguments (d = 2, 4, 8, 16 or 32, out-of-place)");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: return launch_reshuffle<1>(B, in, out, st);
    case 4: return launch_reshuffle<2>(B, in, out, st);
    case 8: return launch_reshuffle<3>(B, in, out, st);
    case 16: return launch_reshuffle<4>(B, in, out, st);
    default: return launch_reshuffle<5>(B, in, out, st);
  }
}

// ---------------------------------------------------------------------------------------------
// superop <-> Pauli-Liouville (PTM)
// ---------------------------------------------------------------------------------------------
// FWD : out = (1/d) F S F^dagger, rows/cols delivered in canonical Pauli order.
// !FWD: out = (1/d) F^dagger R F, rows/cols of the input are in canonical Pauli order.
// F = kron_q F_q; F_q is a 4-point add-only butterfly on the bit pair (j_q, i_q) of the vec index.  The inner
// (column-index) transform and the outer (row-index) transform commute and so do different qubits, so the
// work is cut into two kinds of shared-memory passes:
//   pass A: tiles of 4^RQN full rows (all combinations of the first RQN row qubits): all N inner stages +
//           the outer stages of those RQN qubits.  n <= 3: RQN = N, the whole matrix is one tile -> one pass.
//   pass B: tiles of 4^(N-RQN) rows x W columns: the remaining outer stages, the row permutation and the scale.
// Between the passes the matrix lives in `workspace` with rows in "position" order (row index bits
// (j_0..j_{n-1}, i_0..i_{n-1}); Pauli digits of not-yet-transformed qubits sit at their (j, i) bit positions).
// Row runs are >= 256 B in every pass; every shared buffer row is padded by one element.
template <int N>
struct PlCfg {
  static constexpr int L = 1 << (2 * N);
  static constexpr int RQN = (N <= 3) ? N : (N == 4 ? 2 : 1);     // row qubits handled by pass A
  static constexpr bool RADIX16 = (N == 2);                       // two butterfly stages per shared-memory pass
  static constexpr int TRA = 1 << (2 * RQN);                      // tile rows of pass A
  static constexpr int IPB = (QT_PL_TILE / (TRA * L) >= 1) ? QT_PL_TILE / (TRA * L) : 1;  // tiles per block (small n)
  static constexpr int LDA = L + 1;
  static constexpr size_t smem_a = sizeof(cplx) * IPB * TRA * LDA;
  static constexpr int NTA = (TRA * L * IPB >= 4096) ? 512 : 256;
  static constexpr int TRB = 1 << (2 * (N - RQN));                // tile rows of pass B
  static constexpr int W = (N == 4) ? 256 : 16;                   // tile columns of pass B (n = 4: full rows)
  static constexpr int LDB = W + 1;
  static constexpr size_t smem_b = sizeof(cplx) * TRB * LDB;
  static constexpr int NTB = (TRB * W >= 4096) ? 512 : 256;
};

template <int N, bool FWD>
__global__ void __launch_bounds__(PlCfg<N>::NTA) pl_pass_a_kernel(int64_t n_tiles, const cplx* __restrict__ in,
                                                                  cplx* __restrict__ out) {
  using C = PlCfg<N>;
  constexpr int L = C::L, RQN = C::RQN, TRA = C::TRA, LDA = C::LDA, NT = C::NTA;
  constexpr int REST = 1 << (2 * (N - RQN));  // combinations of the row qubits NOT in this tile
  constexpr bool SINGLE = (RQN == N);
  extern __shared__ __align__(16) cplx buf[];
  const int64_t t0 = (int64_t)blockIdx.x * C::IPB;
  const int nt_here = (int)min((int64_t)C::IPB, n_tiles - t0);
  const double scale = SINGLE ? 1.0 / (double)(1 << N) : 1.0;
  // load: tile row t <-> Pauli-digit row index idx = t * REST + fixed  (position row = pauli_to_pos(idx))
  for (int e = threadIdx.x; e < nt_here * TRA * L; e += NT) {
    const int c = e % L, t = (e / L) % TRA, tl = e / (L * TRA);
    const int64_t tile_id = t0 + tl;
    const int64_t b = tile_id / REST;
    const int idx = t * REST + (int)(tile_id % REST);
    const int row = FWD ? pauli_to_pos(idx, N) : idx;
    const cplx v = in[(b * L + row) * L + c];
    buf[(tl * TRA + t) * LDA + (FWD ? c : pauli_to_pos(c, N))] = v;
  }
  __syncthreads();
  // N inner stages (right factor F^dagger: conjugated butterfly on column bits (N-1-q, 2N-1-q)) and RQN outer
  // stages (left factor F on the tile-row digit q).  Radix 16 (two stages per pass over shared memory, 16
  // elements per work item in registers) where it measured faster (n = 2), radix 4 otherwise
  // (profiles/r01_bench_convert_v4.json vs _v5.json).  The tiles of a block are stacked along the row index.
  if constexpr (C::RADIX16) {
    {
      constexpr int NS = N + RQN;
      const int n_el = nt_here * TRA * L;
      auto lo_of = [](int st) { return st < N ? N - 1 - st : 2 * N + 2 * (RQN - 1 - (st - N)); };
      auto hi_of = [](int st) { return st < N ? 2 * N - 1 - st : 2 * N + 2 * (RQN - 1 - (st - N)) + 1; };
#pragma unroll
      for (int st = 0; st + 1 < NS; st += 2) {
        const bool in1 = st < N, in2 = st + 1 < N;
        if (in1 && in2)
          flat_stage2<FWD, true, true>(buf, n_el / 16, 2 * N, LDA, lo_of(st), hi_of(st), lo_of(st + 1), hi_of(st + 1),
                                       threadIdx.x, NT);
        else if (in1)
          flat_stage2<FWD, true, false>(buf, n_el / 16, 2 * N, LDA, lo_of(st), hi_of(st), lo_of(st + 1), hi_of(st + 1),
                                        threadIdx.x, NT);
        else
          flat_stage2<FWD, false, false>(buf, n_el / 16, 2 * N, LDA, lo_of(st), hi_of(st), lo_of(st + 1), hi_of(st + 1),
                                         threadIdx.x, NT);
        __syncthreads();
      }
      if (NS & 1) {
        constexpr int st = NS - 1;
        if (st < N) flat_stage1<FWD, true>(buf, n_el / 4, 2 * N, LDA, lo_of(st), hi_of(st), threadIdx.x, NT);
        else flat_stage1<FWD, false>(buf, n_el / 4, 2 * N, LDA, lo_of(st), hi_of(st), threadIdx.x, NT);
        __syncthreads();
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) {
      // consecutive threads walk the ROWS (odd leading dimension: conflict-free) when a warp's worth exists;
      // walking the row would put the 4 partners of lanes 0 and 4 on the same banks (2-way conflict)
      if (TRA * C::IPB >= 8)
        bfly_stage<FWD, true, true>(buf, 2 * N, N - 1 - q, 2 * N - 1 - q, nt_here * TRA, LDA, 1, threadIdx.x, NT);
      else
        bfly_stage<FWD, true, false>(buf, 2 * N, N - 1 - q, 2 * N - 1 - q, nt_here * TRA, LDA, 1, threadIdx.x, NT);
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < RQN; ++q) {
      for (int tl = 0; tl < nt_here; ++tl)
        bfly_stage<FWD, false, true>(buf + tl * TRA * LDA, 2 * RQN, 2 * (RQN - 1 - q), 2 * (RQN - 1 - q) + 1, L, 1, LDA,
                                     threadIdx.x, NT);
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < nt_here * TRA * L; e += NT) {
    const int c = e % L, t = (e / L) % TRA, tl = e / (L * TRA);
    const int64_t tile_id = t0 + tl;
    const int64_t b = tile_id / REST;
    const int idx = t * REST + (int)(tile_id % REST);
    // FWD: columns leave in Pauli order; rows in Pauli order if this is the only pass, else position order.
    // !FWD: columns are back in position order; rows leave in position order.
    const int row = FWD ? (SINGLE ? idx : pauli_to_pos(idx, N)) : pauli_to_pos(idx, N);
    const cplx v = buf[(tl * TRA + t) * LDA + (FWD ? pauli_to_pos(c, N) : c)];
    out[(b * L + row) * L + c] = SINGLE ? cscale(v, scale) : v;
  }
}

template <int N, bool FWD>
__global__ void __launch_bounds__(PlCfg<N>::NTB) pl_pass_b_kernel(int64_t B, const cplx* __restrict__ in,
                                                                  cplx* __restrict__ out) {
  using C = PlCfg<N>;
  constexpr int L = C::L, RQN = C::RQN, TRB = C::TRB, W = C::W, LDB = C::LDB, NT = C::NTB;
  constexpr int TOPS = 1 << (2 * RQN), PANELS = L / W;
  extern __shared__ __align__(16) cplx buf[];
  const int64_t tile_id = blockIdx.x;
  const int panel = (int)(tile_id % PANELS);
  const int top = (int)((tile_id / PANELS) % TOPS);
  const int64_t b = tile_id / ((int64_t)PANELS * TOPS);
  const int c0 = panel * W;
  const double scale = 1.0 / (double)(1 << N);
  const cplx* src = in + b * L * L;
  cplx* dst = out + b * L * L;
  for (int e = threadIdx.x; e < TRB * W; e += NT) {
    const int cc = e % W, t = e / W;
    const int idx = top * TRB + t;
    buf[t * LDB + cc] = src[(int64_t)pauli_to_pos(idx, N) * L + c0 + cc];
  }
  __syncthreads();
  // outer stages of the remaining qubits q = RQN..N-1: digit q of the tile row t sits at bits
  // log2(W) + 2(N-1-q), +1 of the flat index t * W + cc
  if constexpr (C::RADIX16) {
    {
      constexpr int NS = N - RQN;
      constexpr int WB = (W == 256) ? 8 : (W == 32 ? 5 : 4);
      auto lo_of = [](int st) { return WB + 2 * (N - 1 - (RQN + st)); };
#pragma unroll
      for (int st = 0; st + 1 < NS; st += 2) {
        flat_stage2<FWD, false, false>(buf, TRB * W / 16, WB, LDB, lo_of(st), lo_of(st) + 1, lo_of(st + 1),
                                       lo_of(st + 1) + 1, threadIdx.x, NT);
        __syncthreads();
      }
      if (NS & 1) {
        flat_stage1<FWD, false>(buf, TRB * W / 4, WB, LDB, lo_of(NS - 1), lo_of(NS - 1) + 1, threadIdx.x, NT);
        __syncthreads();
      }
    }
  } else {
#pragma unroll
    for (int q = RQN; q < N; ++q) {
      bfly_stage<FWD, false, true>(buf, 2 * (N - RQN), 2 * (N - 1 - q), 2 * (N - 1 - q) + 1, W, 1, LDB, threadIdx.x, NT);
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < TRB * W; e += NT) {
    const int cc = e % W, t = e / W;
    const int idx = top * TRB + t;
    const int row = FWD ? idx : pauli_to_pos(idx, N);
    dst[(int64_t)row * L + c0 + cc] = cscale(buf[t * LDB + cc], scale);
  }
}

// ---------------------------------------------------------------------------------------------
// n = 3: register-resident radix-16 passes.
// pl_pass_a_kernel<3> runs its six radix-4 stages in shared memory: 14 passes over a 64 KB tile per matrix, which made
// it shared-memory-bandwidth bound at 0.53 of the HBM roof (profiles/r01_ncu_ptm3_kernel.md).  Here a 256-thread block
// keeps the matrix in REGISTERS, 16 elements per thread, chosen so that two complete stages run without any exchange:
//     pass 1  row stages q = 0, 1   (registers <-> bits 11..8 of the flat index e = row * 64 + col)
//     pass 2  row stage  q = 2 and column stage q = 2   (bits 7, 6, 1, 0)
//     pass 3  column stages q = 0, 1   (bits 5..2)
// with ONE 64 KB shared buffer used twice as a transposition stage (4 passes over shared memory instead of 14).
// Rows and columns are held in base-4 DIGIT order (digit q of an index at bits 5-2q, 4-2q = (j_q, i_q)), which after
// the forward butterflies IS the canonical Pauli order; the computational ("position") order of the other side is
// applied as a bit permutation on the global address, so both the loads and the stores are full 32-byte sectors.
// The shared buffer is addressed through a linear (XOR) swizzle that makes both access patterns of each transposition
// conflict-free: bank group = 3 parity bits, one of which toggles for every bit a quarter-warp varies.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pl3_bits(int e, int b) { return (e >> b) & 1; }
// transposition 1 (written in the pass-1 mapping, read in the pass-2 mapping)
template <bool FWD>
__device__ __forceinline__ int pl3_swz1(int e) {
  const int e0 = pl3_bits(e, 0), e1 = pl3_bits(e, 1), e2 = pl3_bits(e, 2), e3 = pl3_bits(e, 3), e4 = pl3_bits(e, 4);
  if (FWD)  // writers vary (e0, e2, e4), readers (e2, e3, e4)
    return ((e >> 5) << 5) | (e1 << 4) | (e3 << 3) | (e4 << 2) | (e2 << 1) | (e0 ^ e3);
  // !FWD: writers vary (e0, e1, e2), readers (e2, e3, e4)
  return ((e >> 5) << 5) | (e4 << 4) | (e3 << 3) | (e2 << 2) | ((e1 ^ e4) << 1) | (e0 ^ e3);
}
// transposition 2 (written in the pass-2 mapping: (e2, e3, e4) vary; read in the pass-3 mapping: (e0, e1, e6) vary)
__device__ __forceinline__ int pl3_swz2(int e) {
  const int e0 = pl3_bits(e, 0), e1 = pl3_bits(e, 1), e2 = pl3_bits(e, 2), e3 = pl3_bits(e, 3), e4 = pl3_bits(e, 4);
  const int e5 = pl3_bits(e, 5), e6 = pl3_bits(e, 6);
  return ((e >> 7) << 7) | (e5 << 6) | (e6 << 5) | (e1 << 4) | (e0 << 3) | ((e4 ^ e6) << 2) | ((e3 ^ e1) << 1) | (e2 ^ e0);
}
// two stages on the 16 registers: the first on register bits (3, 2), the second on bits (1, 0)
template <bool FWD, bool CONJ_HI, bool CONJ_LO>
__device__ __forceinline__ void pl3_two_stages(cplx (&u)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) bfly4<FWD, CONJ_HI>(u[b], u[4 + b], u[8 + b], u[12 + b]);
#pragma unroll
  for (int a = 0; a < 4; ++a) bfly4<FWD, CONJ_LO>(u[4 * a], u[4 * a + 1], u[4 * a + 2], u[4 * a + 3]);
}

template <bool FWD>
__global__ void __launch_bounds__(256, 2) pl3_reg_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  constexpr int N = 3, L = 64;
  extern __shared__ __align__(16) cplx buf[];  // 4096 elements
  const int tid = threadIdx.x;
  // thread part of the flat index e in the three mappings (the register part is OR-ed / XOR-ed in per element)
  const int col1 = tid & 63;                                     // pass 1: the global column this thread reads
  const int t1 = (FWD ? pos_to_pauli(col1, N) : col1) | ((tid >> 6) << 6);
  const int t2 = ((tid & 15) << 2) | ((tid >> 4) << 8);
  const int t3 = (tid & 3) | ((tid >> 2) << 6);
  const int s1w = pl3_swz1<FWD>(t1), s1r = pl3_swz1<FWD>(t2), s2w = pl3_swz2(t2), s2r = pl3_swz2(t3);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* src = in + b * (L * L);
    cplx* dst = out + b * (L * L);
    cplx u[16];
    // ---- pass 1: registers <-> row digits 0, 1 (e bits 11..8); lanes <-> 32 consecutive global columns ----
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int tr = (r << 2) | (tid >> 6);
      u[r] = src[(FWD ? pauli_to_pos(tr, N) : tr) * L + col1];
    }
    pl3_two_stages<FWD, false, false>(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[s1w ^ pl3_swz1<FWD>(r << 8)] = u[r];
    __syncthreads();
    // ---- pass 2: registers <-> e bits (7, 6) = row digit 2 and (1, 0) = column digit 2 ----
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = buf[s1r ^ pl3_swz1<FWD>(((r >> 2) << 6) | (r & 3))];
    __syncthreads();
    pl3_two_stages<FWD, false, true>(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[s2w ^ pl3_swz2(((r >> 2) << 6) | (r & 3))] = u[r];
    __syncthreads();
    // ---- pass 3: registers <-> e bits 5..2 = column digits 0, 1; lanes <-> 4 consecutive columns x 8 rows ----
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = buf[s2r ^ pl3_swz2(r << 2)];
    __syncthreads();
    pl3_two_stages<FWD, true, true>(u);
    const int tr = tid >> 2;
    const int64_t rowoff = (int64_t)(FWD ? tr : pauli_to_pos(tr, N)) * L;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int tc = (r << 2) | (tid & 3);
      dst[rowoff + (FWD ? tc : pauli_to_pos(tc, N))] = cscale(u[r], 0.125);
    }
  }
}

// n = 1: a 4 x 4 matrix is 16 registers of ONE thread -- the same two butterfly stages (row digit on register bits
// (3, 2), column digit on bits (1, 0)) with no exchange at all; 128 matrices per block go through a shared-memory
// transposition so that both the loads and the stores are contiguous 16-byte lanes.  The generic shared-memory kernel
// ran this size at 0.47 of the HBM roof.
template <bool FWD>
__global__ void __launch_bounds__(128) pl1_reg_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  constexpr int N = 1;
  __shared__ cplx tile[16 * QT_TS];
  const int tid = threadIdx.x;
  for (int64_t b0 = (int64_t)blockIdx.x * 128; b0 < B; b0 += (int64_t)gridDim.x * 128) {
    const int nb = (int)min((int64_t)128, B - b0);
    for (int e = tid; e < nb * 16; e += 128) tile[(e % 16) * QT_TS + (e / 16)] = in[b0 * 16 + e];
    __syncthreads();
    if (tid < nb) {
      cplx u[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int tr = r >> 2, tc = r & 3;  // digit (Pauli) order; the superoperator side is in position order
        u[r] = tile[((FWD ? pauli_to_pos(tr, N) : tr) * 4 + (FWD ? pauli_to_pos(tc, N) : tc)) * QT_TS + tid];
      }
      pl3_two_stages<FWD, false, true>(u);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int tr = r >> 2, tc = r & 3;
        tile[((FWD ? tr : pauli_to_pos(tr, N)) * 4 + (FWD ? tc : pauli_to_pos(tc, N))) * QT_TS + tid] = cscale(u[r], 0.5);
      }
    }
    __syncthreads();
    for (int e = tid; e < nb * 16; e += 128) out[b0 * 16 + e] = tile[(e % 16) * QT_TS + (e / 16)];
    __syncthreads();
  }
}

template <bool FWD>
static int launch_pl3_reg(int64_t B, const void* in, void* out, cudaStream_t st) {
  const size_t smem = sizeof(cplx) * 4096;
  QT_CUDA(cudaFuncSetAttribute(pl3_reg_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 2 * 4);
  pl3_reg_kernel<FWD><<<(unsigned)blocks, 256, smem, st>>>(B, (const cplx*)in, (cplx*)out);
  return qt_check_launch("pl3_reg_kernel");
}

// ---------------------------------------------------------------------------------------------
// n = 4, 5: the same register-resident passes for the two HBM passes of the factored transform.
//   pass A tile = 4^RQN rows x all 4^n columns = 4096 elements (n = 4: 16 x 256, n = 5: 4 x 1024): flat index
//   e = (tile_row << 2n) | column_digits, six stages on the bit pairs (11,10) (9,8) | (7,6) (1,0) | (5,4) (3,2) exactly
//   like pl3_reg_kernel (pairs at or above bit 2n are row stages, the others column stages), same swizzles.
//   pass B, n = 4: 16 rows x 256 columns, two row stages: one register pass, no shared memory at all.
//   pass B, n = 5: 256 rows x 16 columns, four row stages: two register passes around one transposition.
// ---------------------------------------------------------------------------------------------
// global-memory access flavours: the two-kernel path uses plain accesses; the fused path streams the matrix in and
// out (evict-first) and keeps the intermediate in L2 (written with .cg, read back with .cg: never through a stale L1 line)
struct PlPlain {
  static __device__ __forceinline__ cplx ld_in(const cplx* p) { return *p; }
  static __device__ __forceinline__ void st_out(cplx* p, cplx v) { *p = v; }
  static __device__ __forceinline__ cplx ld_ws(const cplx* p) { return *p; }
  static __device__ __forceinline__ void st_ws(cplx* p, cplx v) { *p = v; }
};
struct PlFused {
  static __device__ __forceinline__ cplx ld_in(const cplx* p) { return __ldcs(p); }
  static __device__ __forceinline__ void st_out(cplx* p, cplx v) { __stcs(p, v); }
  static __device__ __forceinline__ cplx ld_ws(const cplx* p) { return __ldcg(p); }
  static __device__ __forceinline__ void st_ws(cplx* p, cplx v) { __stcg(p, v); }
};

// one pass-A tile: src = the input matrix, dst = the intermediate matrix (rows in position order)
template <int N, bool FWD, class P>
__device__ __forceinline__ void pl_tile_a(const cplx* __restrict__ src, cplx* __restrict__ dst, int fixed, cplx* buf,
                                          int tid) {
  using C = PlCfg<N>;
  constexpr int L = C::L, RQN = C::RQN, CB = 2 * N;
  constexpr int REST = 1 << (2 * (N - RQN));
  static_assert(2 * RQN + CB == 12, "a pass-A tile is 4096 elements");
  const int t1 = FWD ? pos_to_pauli(tid, 4) : tid;  // e bits 7..0: the four trailing column digits
  const int t2 = ((tid & 15) << 2) | ((tid >> 4) << 8);
  const int t3 = (tid & 3) | ((tid >> 2) << 6);
  const int s1w = pl3_swz1<FWD>(t1), s1r = pl3_swz1<FWD>(t2), s2w = pl3_swz2(t2), s2r = pl3_swz2(t3);
  cplx u[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int e = t1 | (r << 8);
    const int tc = e & (L - 1), idx = (e >> CB) * REST + fixed;
    u[r] = P::ld_in(src + (int64_t)(FWD ? pauli_to_pos(idx, N) : idx) * L + (FWD ? pauli_to_pos(tc, N) : tc));
  }
  pl3_two_stages<FWD, (10 < CB), (8 < CB)>(u);
#pragma unroll
  for (int r = 0; r < 16; ++r) buf[s1w ^ pl3_swz1<FWD>(r << 8)] = u[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) u[r] = buf[s1r ^ pl3_swz1<FWD>(((r >> 2) << 6) | (r & 3))];
  __syncthreads();
  pl3_two_stages<FWD, (6 < CB), true>(u);
#pragma unroll
  for (int r = 0; r < 16; ++r) buf[s2w ^ pl3_swz2(((r >> 2) << 6) | (r & 3))] = u[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) u[r] = buf[s2r ^ pl3_swz2(r << 2)];
  __syncthreads();
  pl3_two_stages<FWD, true, true>(u);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int e = t3 | (r << 2);
    const int tc = e & (L - 1), idx = (e >> CB) * REST + fixed;
    // rows leave in position order (pass B expects them there); FWD columns are now canonical, !FWD columns positional
    P::st_ws(dst + (int64_t)pauli_to_pos(idx, N) * L + (FWD ? tc : pauli_to_pos(tc, N)), u[r]);
  }
}

// one pass-B tile.  n = 4: tile = `top` (row digits 0, 1), thread = column, registers = the 16 rows (row digits 2, 3).
// n = 5: tile = (top = row digit 0, panel of 16 columns), 256 rows; e = (tile_row << 4) | column.
template <int N, bool FWD, class P>
__device__ __forceinline__ void pl_tile_b(const cplx* __restrict__ src, cplx* __restrict__ dst, int tile, cplx* buf,
                                          int tid) {
  constexpr int L = 1 << (2 * N);
  cplx u[16];
  if constexpr (N == 4) {
    const int top = tile, c = tid;
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = P::ld_ws(src + (int64_t)pauli_to_pos(top * 16 + r, N) * L + c);
    pl3_two_stages<FWD, false, false>(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int idx = top * 16 + r;
      P::st_out(dst + (int64_t)(FWD ? idx : pauli_to_pos(idx, N)) * L + c, cscale(u[r], 1.0 / 16.0));
    }
  } else {
    constexpr int PANELS = L / 16;
    const int panel = tile % PANELS, top = tile / PANELS;
    const int t2 = (tid & 15) | ((tid >> 4) << 8);  // pass 2: threads <-> e bits 3..0 and 11..8
    src += panel * 16;
    dst += panel * 16;
#pragma unroll
    for (int r = 0; r < 16; ++r) {  // pass 1: registers <-> e bits 11..8 (row digits 1, 2); threads <-> e bits 7..0
      const int e = tid | (r << 8);
      u[r] = P::ld_ws(src + (int64_t)pauli_to_pos(top * 256 + (e >> 4), N) * L + (e & 15));
    }
    pl3_two_stages<FWD, false, false>(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[tid | (r << 8)] = u[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = buf[t2 | (r << 4)];  // pass 2: registers <-> e bits 7..4 (row digits 3, 4)
    __syncthreads();
    pl3_two_stages<FWD, false, false>(u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int e = t2 | (r << 4);
      const int idx = top * 256 + (e >> 4);
      P::st_out(dst + (int64_t)(FWD ? idx : pauli_to_pos(idx, N)) * L + (e & 15), cscale(u[r], 1.0 / 32.0));
    }
  }
}

template <int N>
struct PlFuseCfg {
  static constexpr int L = 1 << (2 * N);
  static constexpr int TA = 1 << (2 * (N - PlCfg<N>::RQN));       // pass-A tiles per matrix
  static constexpr int TB = (N == 4) ? 16 : 4 * (L / 16);         // pass-B tiles per matrix
  static constexpr int LAG = (N == 4) ? 12 : 1;                   // matrices between a matrix's A tiles and its B tiles
  static constexpr int RING = (N == 4) ? 24 : 3;                  // intermediate matrices kept (L2-resident ring)
};

template <int N, bool FWD>
__global__ void __launch_bounds__(256, 2) pl_pass_a_reg_kernel(int64_t n_tiles, const cplx* __restrict__ in,
                                                               cplx* __restrict__ out) {
  constexpr int L = PlCfg<N>::L, REST = PlFuseCfg<N>::TA;
  extern __shared__ __align__(16) cplx buf[];
  for (int64_t tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    const int64_t b = tile_id / REST;
    pl_tile_a<N, FWD, PlPlain>(in + b * (int64_t)L * L, out + b * (int64_t)L * L, (int)(tile_id % REST), buf, threadIdx.x);
  }
}

template <int N, bool FWD>
__global__ void __launch_bounds__(256, 2) pl_pass_b_reg_kernel(int64_t n_tiles, const cplx* __restrict__ in,
                                                               cplx* __restrict__ out) {
  constexpr int L = PlCfg<N>::L, TB = PlFuseCfg<N>::TB;
  extern __shared__ __align__(16) cplx buf[];
  for (int64_t tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    const int64_t b = tile_id / TB;
    pl_tile_b<N, FWD, PlPlain>(in + b * (int64_t)L * L, out + b * (int64_t)L * L, (int)(tile_id % TB), buf, threadIdx.x);
  }
}

// Both passes in ONE launch with the intermediate in an L2-resident ring (n = 4: 24 x 1 MB, n = 5: 3 x 16.8 MB).
// Persistent blocks pull work items from a global counter; the item order is, per matrix index m,
//     [pass-A tiles of matrix m]  [pass-B tiles of matrix m - LAG],
// so a B tile is normally dequeued long after the A tiles it depends on have finished.  Dependencies are enforced anyway:
// a B tile waits until adone[m] == TA, an A tile that re-uses a ring slot waits until bdone[m - RING] == TB.  Every wait
// is on items that were dequeued EARLIER, and the grid never exceeds the number of co-resident blocks, so there is no
// deadlock.  Intermediate lines are written and read with .cg (L2 only), the matrix itself streams with evict-first hints:
// the second HBM pass of the two-kernel path (profiles/r01_traffic_convert.md: 2.0x the algorithmic traffic) disappears.
template <int N, bool FWD>
__global__ void __launch_bounds__(256, 2) pl_fused_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out,
                                                          cplx* __restrict__ ring, int ring_slots,
                                                          unsigned long long* __restrict__ queue, int* __restrict__ adone,
                                                          int* __restrict__ bdone) {
  using F = PlFuseCfg<N>;
  constexpr int L = F::L, TA = F::TA, TB = F::TB, LAG = F::LAG, PER = TA + TB;
  extern __shared__ __align__(16) cplx buf[];
  __shared__ long long item;
  const int tid = threadIdx.x;
  const long long n_items = (long long)(B + LAG) * PER;
  auto wait_for = [&](int* counter, int want) {
    if (tid == 0) {
      while (atomicAdd(counter, 0) < want) __nanosleep(200);
      __threadfence();
    }
    __syncthreads();
  };
  auto signal = [&](int* counter) {
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(counter, 1);
  };
  while (true) {
    if (tid == 0) item = (long long)atomicAdd(queue, 1ULL);
    __syncthreads();
    const long long k = item;
    __syncthreads();
    if (k >= n_items) break;
    const int64_t g = k / PER;
    const int w = (int)(k % PER);
    if (w < TA) {
      const int64_t m = g;
      if (m >= B) continue;
      if (m >= ring_slots) wait_for(bdone + (m - ring_slots), TB);  // the slot's previous tenant has been consumed
      pl_tile_a<N, FWD, PlFused>(in + m * (int64_t)L * L, ring + (m % ring_slots) * (int64_t)L * L, w, buf, tid);
      signal(adone + m);
    } else {
      const int64_t m = g - LAG;
      if (m < 0 || m >= B) continue;
      wait_for(adone + m, TA);
      pl_tile_b<N, FWD, PlFused>(ring + (m % ring_slots) * (int64_t)L * L, out + m * (int64_t)L * L, w - TA, buf, tid);
      signal(bdone + m);
    }
  }
}

template <int N, bool FWD>
static int launch_pl_reg_two_pass(int64_t B, const void* in, void* out, void* workspace, cudaStream_t st) {
  using C = PlCfg<N>;
  using F = PlFuseCfg<N>;
  QT_REQUIRE(workspace, "superop<->pauli_liouville with n >= 4 needs a workspace of B*16^n*16 bytes");
  const size_t smem = sizeof(cplx) * 4096;
  const int64_t mat_bytes = (int64_t)C::L * C::L * sizeof(cplx);
#if QT_PL_FUSED_PASSES
  // fused path: the caller's workspace (B matrices) holds the ring and, in its last matrix, the queue + the counters
  if (B - 1 > F::LAG && 256 + 8 * B <= mat_bytes) {  // ring_slots > LAG: every wait is on an EARLIER item
    const int ring_slots = (int)std::min<int64_t>(B - 1, F::RING);
    unsigned char* tail = (unsigned char*)workspace + (B - 1) * mat_bytes;
    unsigned long long* queue = (unsigned long long*)tail;
    int* adone = (int*)(tail + 256);
    int* bdone = adone + B;
    QT_CUDA(cudaMemsetAsync(tail, 0, 256 + 8 * B, st));
    QT_CUDA(cudaFuncSetAttribute(pl_fused_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    QT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pl_fused_kernel<N, FWD>, 256, smem));
    QT_REQUIRE(per_sm >= 1, "pl_fused_kernel does not fit on this device");
    const int64_t grid = std::min<int64_t>((int64_t)QT_NUM_SMS * per_sm, (B + F::LAG) * (F::TA + F::TB));
    pl_fused_kernel<N, FWD><<<(unsigned)grid, 256, smem, st>>>(B, (const cplx*)in, (cplx*)out, (cplx*)workspace,
                                                              ring_slots, queue, adone, bdone);
    return qt_check_launch("pl_fused_kernel");
  }
#endif
  const int64_t cap = (int64_t)QT_NUM_SMS * 2 * 8;
  const int64_t tiles_a = B * F::TA, tiles_b = B * F::TB;
  QT_CUDA(cudaFuncSetAttribute(pl_pass_a_reg_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pl_pass_a_reg_kernel<N, FWD><<<(unsigned)std::min(tiles_a, cap), 256, smem, st>>>(tiles_a, (const cplx*)in,
                                                                                   (cplx*)workspace);
  int rc = qt_check_launch("pl_pass_a_reg_kernel");
  if (rc) return rc;
  QT_CUDA(cudaFuncSetAttribute(pl_pass_b_reg_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pl_pass_b_reg_kernel<N, FWD><<<(unsigned)std::min(tiles_b, cap), 256, smem, st>>>(tiles_b, (const cplx*)workspace,
                                                                                   (cplx*)out);
  return qt_check_launch("pl_pass_b_reg_kernel");
}

template <int N, bool FWD>
static int launch_pl_n(int64_t B, const void* in, void* out, void* workspace, cudaStream_t st) {
  using C = PlCfg<N>;
  constexpr int REST = 1 << (2 * (N - C::RQN));
#if QT_PL3_REGISTER_KERNEL
  if constexpr (N == 1) {
    const int64_t blocks = std::min<int64_t>((B + 127) / 128, (int64_t)QT_NUM_SMS * 16);
    pl1_reg_kernel<FWD><<<(unsigned)blocks, 128, 0, st>>>(B, (const cplx*)in, (cplx*)out);
    return qt_check_launch("pl1_reg_kernel");
  }
  if constexpr (N == 3) return launch_pl3_reg<FWD>(B, in, out, st);
  if constexpr (N >= 4) return launch_pl_reg_two_pass<N, FWD>(B, in, out, workspace, st);
#endif
  const int64_t n_tiles = B * REST;
  const int64_t blocks_a = (n_tiles + C::IPB - 1) / C::IPB;
  QT_CUDA(cudaFuncSetAttribute(pl_pass_a_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_a));
  if (C::RQN == N) {
    pl_pass_a_kernel<N, FWD><<<(unsigned)blocks_a, C::NTA, C::smem_a, st>>>(n_tiles, (const cplx*)in, (cplx*)out);
    return qt_check_launch("pl_pass_a_kernel");
  }
  QT_REQUIRE(workspace, "superop<->pauli_liouville with n >= 4 needs a workspace of B*16^n*16 bytes");
  pl_pass_a_kernel<N, FWD><<<(unsigned)blocks_a, C::NTA, C::smem_a, st>>>(n_tiles, (const cplx*)in, (cplx*)workspace);
  int rc = qt_check_launch("pl_pass_a_kernel");
  if (rc) return rc;
  QT_CUDA(cudaFuncSetAttribute(pl_pass_b_kernel<N, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_b));
  const int64_t blocks_b = B * (1 << (2 * C::RQN)) * (C::L / C::W);
  pl_pass_b_kernel<N, FWD><<<(unsigned)blocks_b, C::NTB, C::smem_b, st>>>(B, (const cplx*)workspace, (cplx*)out);
  return qt_check_launch("pl_pass_b_kernel");
}

template <bool FWD>
static int launch_pl(int n, int64_t B, const void* in, void* out, void* workspace, cudaStream_t st) {
  switch (n) {
    case 1: return launch_pl_n<1, FWD>(B, in, out, workspace, st);
    case 2: return launch_pl_n<2, FWD>(B, in, out, workspace, st);
    case 3: return launch_pl_n<3, FWD>(B, in, out, workspace, st);
    case 4: return launch_pl_n<4, FWD>(B, in, out, workspace, st);
    default: return launch_pl_n<5, FWD>(B, in, out, workspace, st);
  }
}

extern "C" int qt_superop2pl_batch(int n, int64_t B, const void* superop, void* pl_out, void* workspace,
                                   void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(n >= 1 && n <= 5 && superop && pl_out, "qt_superop2pl_batch: bad arguments");
  return launch_pl<true>(n, B, superop, pl_out, workspace, (cudaStream_t)stream);
}

extern "C" int qt_pl2superop_batch(int n, int64_t B, const void* pl, void* superop_out, void* workspace,
                                   void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(n >= 1 && n <= 5 && pl && superop_out, "qt_pl2superop_batch: bad arguments");
  return launch_pl<false>(n, B, pl, superop_out, workspace, (cudaStream_t)stream);
}
