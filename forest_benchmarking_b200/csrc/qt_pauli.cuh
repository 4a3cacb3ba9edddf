// Per-qubit 4-point butterflies of the computational <-> Pauli change of basis (see qt_convert.cu header),
// shared by the conversion kernels and the process-tomography kernel.
#pragma once
#include "qt_common.cuh"
#include "qt_eigh.cuh"  // SyncWarp / SyncBlock

// position p (2n bits: j bits high, i bits low; qubit 0 most significant in each half) <-> canonical
// Pauli index with digit code (hi=j_q, lo=i_q): I=(0,0) X=(0,1) Y=(1,0) Z=(1,1) after the butterfly.
__host__ __device__ __forceinline__ int pos_to_pauli(int p, int n) {
  int idx = 0;
  for (int q = 0; q < n; ++q) {
    const int lo = (p >> (n - 1 - q)) & 1, hi = (p >> (2 * n - 1 - q)) & 1;
    const int code = 2 * hi + lo;                   // 0:I 1:X 2:Y 3:Z
    idx |= code << (2 * (n - 1 - q));
  }
  return idx;
}
__host__ __device__ __forceinline__ int pauli_to_pos(int idx, int n) {
  int p = 0;
  for (int q = 0; q < n; ++q) {
    const int code = (idx >> (2 * (n - 1 - q))) & 3;
    p |= (code & 1) << (n - 1 - q);
    p |= (code >> 1) << (2 * n - 1 - q);
  }
  return p;
}

// One 4-point butterfly.  FWD: computational -> Pauli (F, or conj(F) when CONJ); !FWD: the adjoint.
template <bool FWD, bool CONJ>
__device__ __forceinline__ void bfly4(cplx& u00, cplx& u01, cplx& u10, cplx& u11) {
  // argument order: (hi,lo) = (j,i) bit pair -> u[j i]; u01 means j=0,i=1, i.e. matrix element U[i=1][j=0].
  // F row a, entry (i,j) = conj(sigma_a[i,j]).  sigma_y[i=0,j=1] = -i, sigma_y[i=1,j=0] = +i.
  if (FWD) {
    const cplx a = cadd(u00, u11), z = csub(u00, u11);
    const cplx x = cadd(u01, u10);
    // Y: conj(sy[1,0]) u(i=1,j=0) + conj(sy[0,1]) u(i=0,j=1) = -i*u01 + i*u10   (u01 = (j=0,i=1))
    cplx y = csub(u10, u01);
    y = CONJ ? cmake(y.y, -y.x) : cmake(-y.y, y.x);  // (+i or -i) * (u10 - u01)
    u00 = a; u01 = x; u10 = y; u11 = z;
  } else {
    // adjoint: u(i,j) = sum_a sigma_a[i,j] v_a   (or its conjugate)
    const cplx vi = u00, vx = u01, vy = u10, vz = u11;
    cplx iy = CONJ ? cmake(vy.y, -vy.x) : cmake(-vy.y, vy.x);  // (+i or -i) * vy
    u00 = cadd(vi, vz);
    u11 = csub(vi, vz);
    // element (i=1,j=0) = vx + sy[1,0] vy = vx + i vy ; element (i=0,j=1) = vx - i vy
    u01 = cadd(vx, iy);
    u10 = csub(vx, iy);
  }
}

// In-place transform of `count` vectors of length L = 4^n held in shared memory.
// Vector v, element p at  buf[v * vstride + p * estride].
template <bool FWD, bool CONJ, class Sync = SyncBlock>
__device__ void pauli_butterfly_smem(cplx* buf, int n, int count, int vstride, int estride, int tid, int nt) {
  const int L = 1 << (2 * n);
  const int quarter = L >> 2;
  for (int q = 0; q < n; ++q) {
    const int lo_bit = n - 1 - q, hi_bit = 2 * n - 1 - q;
    for (int w = tid; w < count * quarter; w += nt) {
      const int v = w / quarter;
      int r = w % quarter;
      // insert zero bits at lo_bit and hi_bit
      int p = ((r >> lo_bit) << (lo_bit + 1)) | (r & ((1 << lo_bit) - 1));
      p = ((p >> hi_bit) << (hi_bit + 1)) | (p & ((1 << hi_bit) - 1));
      cplx* base = buf + v * vstride;
      const int lo = 1 << lo_bit, hi = 1 << hi_bit;
      cplx u00 = base[p * estride], u01 = base[(p | lo) * estride];
      cplx u10 = base[(p | hi) * estride], u11 = base[(p | hi | lo) * estride];
      bfly4<FWD, CONJ>(u00, u01, u10, u11);
      base[p * estride] = u00;
      base[(p | lo) * estride] = u01;
      base[(p | hi) * estride] = u10;
      base[(p | hi | lo) * estride] = u11;
    }
    Sync::sync();
  }
}


// One butterfly stage on the bit pair (lo_bit < hi_bit) of the element index of `count` vectors of
// 2^len_bits elements (element e of vector v at buf[v * vstride + e * estride]).  VEC_FASTEST: consecutive
// threads walk the vector id first (use when vstride == 1 so that a warp touches consecutive addresses).
// No trailing synchronisation.
template <bool FWD, bool CONJ, bool VEC_FASTEST>
__device__ __forceinline__ void bfly_stage(cplx* buf, int len_bits, int lo_bit, int hi_bit, int count, int vstride,
                                           int estride, int tid, int nt) {
  const int quarter = 1 << (len_bits - 2);
  const int lo = 1 << lo_bit, hi = 1 << hi_bit;
  for (int w = tid; w < count * quarter; w += nt) {
    const int v = VEC_FASTEST ? w % count : w / quarter;
    const int r = VEC_FASTEST ? w / count : w % quarter;
    int p = ((r >> lo_bit) << (lo_bit + 1)) | (r & (lo - 1));  // insert zero bits at lo_bit and hi_bit
    p = ((p >> hi_bit) << (hi_bit + 1)) | (p & (hi - 1));
    cplx* base = buf + v * vstride;
    cplx u00 = base[p * estride], u01 = base[(p | lo) * estride];
    cplx u10 = base[(p | hi) * estride], u11 = base[(p | hi | lo) * estride];
    bfly4<FWD, CONJ>(u00, u01, u10, u11);
    base[p * estride] = u00;
    base[(p | lo) * estride] = u01;
    base[(p | hi) * estride] = u10;
    base[(p | hi | lo) * estride] = u11;
  }
}

// ---------------------------------------------------------------------------------------------
// Flat-index butterflies for the PTM kernels.  A shared tile of 2^total_bits elements is addressed by a flat
// index f whose low `colbits` bits are the column and whose remaining bits are the row of the (padded, leading
// dimension ld) buffer.  A stage acts on one bit pair (lo, hi) of f.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int flat_addr(int f, int colbits, int ld) {
  return (f >> colbits) * ld + (f & ((1 << colbits) - 1));
}
__device__ __forceinline__ int insert_zero_bit(int x, int pos) {
  return ((x >> pos) << (pos + 1)) | (x & ((1 << pos) - 1));
}

template <bool FWD, bool CONJ>
__device__ __forceinline__ void flat_stage1(cplx* buf, int n_items4, int colbits, int ld, int lo, int hi, int tid,
                                            int nt) {
  for (int w = tid; w < n_items4; w += nt) {
    const int f = insert_zero_bit(insert_zero_bit(w, lo), hi);  // lo < hi
    const int a00 = flat_addr(f, colbits, ld), a01 = flat_addr(f | (1 << lo), colbits, ld);
    const int a10 = flat_addr(f | (1 << hi), colbits, ld), a11 = flat_addr(f | (1 << hi) | (1 << lo), colbits, ld);
    cplx u00 = buf[a00], u01 = buf[a01], u10 = buf[a10], u11 = buf[a11];
    bfly4<FWD, CONJ>(u00, u01, u10, u11);
    buf[a00] = u00;
    buf[a01] = u01;
    buf[a10] = u10;
    buf[a11] = u11;
  }
}

// Two stages in one pass over shared memory: 16 elements per work item held in registers (radix 16).
template <bool FWD, bool CONJ1, bool CONJ2>
__device__ __forceinline__ void flat_stage2(cplx* buf, int n_items16, int colbits, int ld, int lo1, int hi1, int lo2,
                                            int hi2, int tid, int nt) {
  // the four bit positions in ascending order (for the zero-bit insertion)
  int p0 = lo1, p1 = hi1, p2 = lo2, p3 = hi2, t;
  if (p0 > p2) { t = p0; p0 = p2; p2 = t; }
  if (p1 > p3) { t = p1; p1 = p3; p3 = t; }
  if (p0 > p1) { t = p0; p0 = p1; p1 = t; }
  if (p2 > p3) { t = p2; p2 = p3; p3 = t; }
  if (p1 > p2) { t = p1; p1 = p2; p2 = t; }
  for (int w = tid; w < n_items16; w += nt) {
    const int f = insert_zero_bit(insert_zero_bit(insert_zero_bit(insert_zero_bit(w, p0), p1), p2), p3);
    cplx u[4][4];
    int addr[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int g = f | ((a >> 1) << hi1) | ((a & 1) << lo1) | ((b >> 1) << hi2) | ((b & 1) << lo2);
        addr[a][b] = flat_addr(g, colbits, ld);
        u[a][b] = buf[addr[a][b]];
      }
#pragma unroll
    for (int b = 0; b < 4; ++b) bfly4<FWD, CONJ1>(u[0][b], u[1][b], u[2][b], u[3][b]);
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4<FWD, CONJ2>(u[a][0], u[a][1], u[a][2], u[a][3]);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) buf[addr[a][b]] = u[a][b];
  }
}
