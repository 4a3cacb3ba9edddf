// Batched state distance measures: one pair (rho_b, sigma_b) per warp.
//
//   trace_distance  distance_measures.py:100-114  0.5 * ||rho - sigma||_1 with NumPy's INDUCED 1-norm
//                   (max column abs-sum) -- the reference's actual behaviour, pinned by its own test
//                   tests/test_distance_measures.py:73-82.  Pure streaming, HBM-bound.
//   trace_distance_nuclear   the textbook 0.5 * sum |eig(rho - sigma)| (extra, not a reference function)
//   fidelity        distance_measures.py:64-84 + calculational.py:77-91
//                   (tr sqrt( sqrt(rho) sigma sqrt(rho) ))^2 = (sum_i sqrt(max(lambda_i, 0)))^2
//   purity          distance_measures.py:14-37    tr(rho^2)
#include "qt_common.cuh"
#include "qt_eigh.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

// Small states (d = 2, 4: fewer than 32 elements) are packed IPW = 32 / d^2 pairs per warp, one element per lane, so
// that every lane of every load is used; larger ones take one pair per warp with the element loop unrolled.
template <int D>
__global__ void trace_distance_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                      double* __restrict__ out) {
  constexpr int DD = D * D;
  constexpr int IPW = (DD < 32) ? 32 / DD : 1;   // pairs per warp
  constexpr int GL = (DD < 32) ? DD : 32;        // lanes per pair
  const int lane = threadIdx.x & 31, gl = lane % GL;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t b = warp * IPW + lane / GL;
  const bool live = b < B;
  const cplx* r = rho + (live ? b : 0) * DD;
  const cplx* s = sigma + (live ? b : 0) * DD;
  double colsum = 0.0;  // lane owns column gl % D, rows gl / D + k * (GL / D)
  if (D <= 32) {
#pragma unroll
    for (int e = gl; e < DD; e += GL) {
      const cplx a = r[e], c = s[e];
      colsum += sqrt(cabs2(csub(a, c)));
    }
#pragma unroll
    for (int o = GL / 2; o >= D; o >>= 1) colsum += __shfl_xor_sync(0xffffffffu, colsum, o);
    double m = colsum;
#pragma unroll
    for (int o = (D < 32 ? D : 32) / 2; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (live && gl == 0) out[b] = 0.5 * m;
  }
}

// tr(rho rho) = sum_ij rho_ij rho_ji (real part).  The transposed partner of an element is fetched without a second,
// uncoalesced global read: by shuffle when a state fits one warp pass (d <= 4), from a padded shared-memory copy
// of the state otherwise.
template <int D>
__global__ void purity_kernel(int64_t B, const cplx* __restrict__ rho, double* __restrict__ out) {
  constexpr int DD = D * D;
  constexpr int IPW = (DD < 32) ? 32 / DD : 1;
  constexpr int GL = (DD < 32) ? DD : 32;
  constexpr int WPB = (D >= 32) ? 2 : 8, LD = D + 1;
  const int lane = threadIdx.x & 31, gl = lane % GL, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * WPB + wib;
  if constexpr (DD < 32) {
    // a warp moves only 32 elements per step: persistent warps, four steps' loads in flight, so that the kernel is
    // bound by HBM and not by how fast blocks can be issued (0.67 of the roof with one step per warp)
    const int i = gl / D, j = gl % D, src = (lane - gl) + j * D + i;
    const int64_t nwarps = (int64_t)gridDim.x * WPB, steps = (B + IPW - 1) / IPW;
    for (int64_t s0 = warp * 4; s0 < steps; s0 += nwarps * 4) {
      cplx a[4];
      int64_t bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        bb[u] = (s0 + u) * IPW + lane / GL;
        a[u] = (bb[u] < B) ? rho[bb[u] * DD + gl] : cmake(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cplx c;
        c.x = __shfl_sync(0xffffffffu, a[u].x, src);
        c.y = __shfl_sync(0xffffffffu, a[u].y, src);
        double acc = a[u].x * c.x - a[u].y * c.y;
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (bb[u] < B && gl == 0) out[bb[u]] = acc;
      }
    }
    return;
  }
  const int64_t b = warp * IPW + lane / GL;
  const bool live = b < B;
  const cplx* r = rho + (live ? b : 0) * DD;
  double acc = 0.0;
  if constexpr (DD >= 32) {
    __shared__ cplx tile[WPB][D * LD];
    cplx* t = tile[wib];
#pragma unroll
    for (int e = lane; e < DD; e += 32) t[(e / D) * LD + e % D] = r[e];
    __syncwarp();
#pragma unroll
    for (int e = lane; e < DD; e += 32) {
      const int i = e / D, j = e % D;
      const cplx a = t[i * LD + j], c = t[j * LD + i];
      acc += a.x * c.x - a.y * c.y;
    }
  }
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && gl == 0) out[b] = acc;
}

// shared memory per warp: 3 matrices (leading dimension LD = D + 1 for D >= 8: the Jacobi block updates and the
// products read columns, which an unpadded power-of-two row stride maps onto the same banks) + eigenvalues +
// Jacobi scratch
template <int D>
struct FidSmem {
  static constexpr int LD = (D >= 8) ? D + 1 : D;
  static constexpr int MP = D * LD;
  static constexpr size_t bytes =
      (sizeof(cplx) * MP * 3 + sizeof(double) * (D + JacobiScratch<D>::doubles) + 15) / 16 * 16;
};

// Fast path of the fidelity for a numerically positive-definite rho: rho = L L^dagger (Cholesky), and the eigenvalues of
// L^dagger sigma L are those of sqrt(rho) sigma sqrt(rho) (similar matrices), so ONE values-only eigendecomposition
// replaces eigh(rho) with vectors + sqrtm + the second eigh.  Two matrices of shared memory per warp instead of three
// and no eigenvector registers: 1.5x the resident warps of fidelity_kernel.  A pivot that is not safely positive
// (rank-deficient or non-PSD input, where the reference's clamping in sqrtm_psd matters) writes FID_FLAG and leaves the
// item to fidelity_kernel, which runs the reference's own sequence on the flagged items.
constexpr double FID_FLAG = -1.0;  // a fidelity is never negative
template <int D>
struct FidFastSmem {
  static constexpr int LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  static constexpr size_t bytes = (sizeof(cplx) * MP * 2 + sizeof(double) * (D + JacobiScratch<D>::doubles) + 15) / 16 * 16;
};

template <int D>
__global__ void __launch_bounds__(256, (D == 16) ? 3 : 1) fidelity_fast_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                     double* __restrict__ out) {
  constexpr int DD = D * D, LD = FidFastSmem<D>::LD, MP = FidFastSmem<D>::MP, PER = (DD + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidFastSmem<D>::bytes * wib);
  cplx* W = A + MP;
  double* ev = reinterpret_cast<double*>(W + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  const cplx* r = rho + b * DD;
  const cplx* s = sigma + b * DD;
  // scipy.linalg.eigh reads the lower triangle only: factor the Hermitian matrix it sees
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = r[e];
    if (i == j) v.y = 0.0;
    if (i >= j) A[i * LD + j] = v;
  }
  __syncwarp();
  double dmax = 0.0;
  for (int k = 0; k < D; ++k) dmax = fmax(dmax, A[k * LD + k].x);
  const double thr = 1e-10 * dmax;
  bool ok = dmax > 0.0;
  for (int j = 0; j < D && ok; ++j) {
    const double piv = A[j * LD + j].x;
    if (!(piv > thr)) {
      ok = false;
      break;
    }
    const double inv = rsqrt(piv);
    __syncwarp();
    for (int i = j + lane; i < D; i += 32) A[i * LD + j] = (i == j) ? cmake(piv * inv, 0.0) : cscale(A[i * LD + j], inv);
    __syncwarp();
    const int nt = D - j - 1;
    for (int e = lane; e < nt * nt; e += 32) {
      const int i = j + 1 + e / nt, k = j + 1 + e % nt;
      if (k <= i) {
        const cplx li = A[i * LD + j], lk = A[k * LD + j];
        cplx v = A[i * LD + k];
        v.x -= li.x * lk.x + li.y * lk.y;  // v -= li * conj(lk)
        v.y -= li.y * lk.x - li.x * lk.y;
        if (i == k) v.y = 0.0;
        A[i * LD + k] = v;
      }
    }
    __syncwarp();
  }
  if (!ok) {
    if (lane == 0) out[b] = FID_FLAG;
    return;
  }
  // W = sigma L  (L lower triangular in A)
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = j; k < D; ++k) cfma(acc, s[i * D + k], A[k * LD + j]);
    W[i * LD + j] = acc;
  }
  __syncwarp();
  // Y = L^dagger W: lower triangle (incl. diagonal) held in registers until every lane has finished reading L, then
  // written over A together with its conjugate mirror -- the Hermitian matrix an eigensolver reading the lower
  // triangle would see
  cplx y[PER];
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int e = lane + 32 * t, i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    if (e < DD && i >= j)
      for (int k = i; k < D; ++k) cfma(acc, cconj(A[k * LD + i]), W[k * LD + j]);
    y[t] = acc;
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int e = lane + 32 * t, i = e / D, j = e % D;
    if (e < DD && i >= j) {
      cplx v = y[t];
      if (i == j) v.y = 0.0;
      A[i * LD + j] = v;
      if (i > j) A[j * LD + i] = cconj(v);
    }
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, false, LD>(A, nullptr, ev, ev + D, lane);
  double acc = 0.0;
  for (int k = lane; k < D; k += 32) acc += sqrt(fmax(ev[k], 0.0));
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc * acc;
}

// ---------------------------------------------------------------------------------------------
// fidelity, d = 4, 8, 16: D lanes per pair (lane = one row / column held in registers), 32 / D pairs in flight per
// warp, 32 pairs per warp in all.  Per pair: Cholesky rho = L L^dagger (rows in registers, the finished column
// broadcast through shared memory), W = sigma L and Y = L^dagger W (columns in registers), Householder reduction
// of Y to a real symmetric tridiagonal matrix (rows in registers; only the diagonal d and the SQUARED off-diagonal
// e^2 = |x|^2 are kept -- the phases of the complex off-diagonal do not change the spectrum).  After 32 pairs the
// warp switches layout: lane j runs the square-root-free QL iteration (Pal-Walker-Kahan, the algorithm behind
// LAPACK's dsterf) on pair j's (d, e^2), kept in a lane-interleaved shared-memory slab, and writes
// (sum_k sqrt(max(ev_k, 0)))^2.  4.4 k warp instructions per pair (22 % of them the QL phase) against ~29 k for the
// warp-per-pair Jacobi kernel above (ncu: profiles/r02_ncu_fidelity_tri_final.md).  Pairs whose rho fails the pivot test, or whose QL does not converge, get FID_FLAG and are redone by
// fidelity_kernel with the reference's own sequence.
// ---------------------------------------------------------------------------------------------
// Block shape of the d = 16 instance: 3 blocks of 4 warps per SM at 168 registers.  Measured alternatives (2^18 pairs,
// profiles/r02_ubench_fidelity_tri.txt): 16 warps/SM at 128 registers 3.38 ms vs 3.28 ms; block-wide barriers that keep
// the warps of an SM on the same code (instruction-cache misses 48.7 M -> 0.4 M) 3.5-3.6 ms; sigma staged by cp.async
// 3.5 ms; all loops rolled over a shared-memory matrix (20 KB of code instead of 108 KB) 4.9 ms.
#ifndef FID_TRI_LB_T
#define FID_TRI_LB_T 128
#endif
#ifndef FID_TRI_LB_B
#define FID_TRI_LB_B 3
#endif
#ifndef FID_TRI_WPB_N
#define FID_TRI_WPB_N 4
#endif
constexpr int FID_TRI_WPB = FID_TRI_WPB_N;
template <int D>
struct FidTriSmem {
  static constexpr int G = 32 / D, MP = D * (D + 1) / 2;  // L packed by rows: L[k][i] at k (k + 1) / 2 + i, i <= k
  // per warp: G Cholesky factors, G x 2 Householder vectors, the (d, e^2) slab of 32 pairs
  static constexpr size_t bytes = sizeof(cplx) * (G * MP + 2 * 32) + sizeof(double) * 2 * D * 32;
};

template <int D>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = D / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int D>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
  for (int o = D / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// eigenvalues of the symmetric tridiagonal (d, e2) of one lane's pair, in place in d; false if not converged.
// d[k] and e2[k] live at stride ST doubles (lane-interleaved: no bank conflicts whatever k each lane is at).
template <int D, int ST = 32>
__device__ __forceinline__ bool pwk_ql(double* __restrict__ d, double* __restrict__ e2) {
  constexpr double EPS2 = 1.232595164407831e-32;  // (2^-53)^2
  for (int l = 0; l < D; ++l) {
    int it = 0;
    while (true) {
      int m = l;
      while (m < D - 1 && !(e2[m * ST] <= EPS2 * fabs(d[m * ST] * d[(m + 1) * ST]))) ++m;
      if (m == l) break;
      if (++it > 40) return false;
      double p = d[l * ST];
      const double rte = sqrt(e2[l * ST]);
      const double sg = (d[(l + 1) * ST] - p) / (2.0 * rte);
      const double rr = fabs(sg) > 1e100 ? fabs(sg) : sqrt(fma(sg, sg, 1.0));
      const double shift = p - rte / (sg + copysign(rr, sg));
      double c = 1.0, s = 0.0, gamma = d[m * ST] - shift;
      p = gamma * gamma;
      for (int i = m - 1; i >= l; --i) {
        const double bb = e2[i * ST], r = p + bb;
        if (i != m - 1) e2[(i + 1) * ST] = s * r;
        const double oldc = c, ir = fast_rcp(r);
        c = p * ir;
        s = bb * ir;
        const double oldgam = gamma, al = d[i * ST];
        gamma = c * (al - shift) - s * oldgam;
        d[(i + 1) * ST] = oldgam + (al - gamma);
        p = (c != 0.0) ? gamma * gamma * fast_rcp(c) : oldc * bb;
      }
      e2[l * ST] = s * p;
      d[l * ST] = shift + gamma;
    }
  }
  return true;
}

template <int D>
__global__ void __launch_bounds__((D == 16) ? FID_TRI_LB_T : 32 * FID_TRI_WPB, (D == 16) ? FID_TRI_LB_B : 4)
    fidelity_tri_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma, double* __restrict__ out) {
  constexpr int DD = D * D, G = FidTriSmem<D>::G, MP = FidTriSmem<D>::MP, ROUNDS = 32 / G;
  auto lidx = [](int k, int i) { return k * (k + 1) / 2 + i; };
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = lane % D, g = lane / D, glane0 = g * D;
  cplx* Lg = reinterpret_cast<cplx*>(smem_raw + FidTriSmem<D>::bytes * wib) + g * MP;
  cplx* Ug = reinterpret_cast<cplx*>(smem_raw + FidTriSmem<D>::bytes * wib) + G * MP + g * D;
  cplx* Wg = Ug + 32;
  double* td = reinterpret_cast<double*>(reinterpret_cast<cplx*>(smem_raw + FidTriSmem<D>::bytes * wib) + G * MP + 64);
  double* te = td + D * 32;
  const int64_t b0 = ((int64_t)blockIdx.x * FID_TRI_WPB + wib) * 32;
  if (b0 >= B) return;

#pragma unroll 1
  for (int t = 0; t < ROUNDS; ++t) {
    if (b0 + t * G >= B) break;  // warp-uniform
    const int slot = t * G + g;
    const int64_t b = min(b0 + slot, B - 1);  // a group past the end recomputes the last pair; its slot is never read
    const cplx* rp = rho + b * DD;
    const cplx* sp = sigma + b * DD;
    {  // next round's rows on their way into L2 while this round computes (row r = D * 16 bytes)
      const int64_t bn = min(b + G, B - 1);
      const char* pr = reinterpret_cast<const char*>(rho + bn * DD + r * D);
      const char* ps = reinterpret_cast<const char*>(sigma + bn * DD + r * D);
#pragma unroll
      for (int o = 0; o < D * 16; o += 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + o));
      }
    }
    // ---- Cholesky, lane r owns row r (scipy's eigh reads the lower triangle only: so does this) ----
    cplx a[D];
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = rp[r * D + c];
    const double dmax = group_max<D>(rp[r * D + r].x);
    const double thr = 1e-10 * dmax;
    bool ok = dmax > 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double piv = __shfl_sync(0xffffffffu, a[j].x, glane0 + j);
      ok = ok && (piv > thr);
      const double inv = ok ? fast_rsqrt(piv) : 0.0;
      cplx l = (r == j) ? cmake(piv * inv, 0.0) : cscale(a[j], inv);
      a[j] = l;
      if (r >= j) Lg[lidx(r, j)] = l;  // triangular numbers: distinct banks within each quarter warp
      __syncwarp();
#pragma unroll
      for (int k = j + 1; k < D; ++k) {
        const cplx lk = Lg[lidx(k, j)];
        a[k].x = fma(-l.y, lk.y, fma(-l.x, lk.x, a[k].x));  // a[k] -= l * conj(lk)
        a[k].y = fma(l.x, lk.y, fma(-l.y, lk.x, a[k].y));
      }
    }
    // ---- W = sigma L, Y = L^dagger W: lane r owns COLUMN r ----
    cplx y[D];
    {
      cplx lc[D], w[D];
#pragma unroll
      for (int m = 0; m < D; ++m) lc[m] = (m >= r) ? Lg[lidx(m, r)] : cmake(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < D; ++m) cfma(acc, sp[k * D + m], lc[m]);
        w[k] = acc;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = i; k < D; ++k) cfma(acc, cconj(Lg[lidx(k, i)]), w[k]);
        y[i] = acc;
      }
    }
    // row r of the Hermitian matrix an eigensolver reading the lower triangle would see: A[r][c] = conj(Y[c][r])
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = (c == r) ? cmake(y[c].x, 0.0) : cconj(y[c]);
    // ---- Householder tridiagonalisation, lane r owns row r ----
#pragma unroll
    for (int k = 0; k < D - 2; ++k) {
      const cplx x = (r > k) ? a[k] : cmake(0.0, 0.0);
      Ug[r] = x;
      const double n2 = group_sum<D>(cabs2(x));
      __syncwarp();
      const cplx alpha = Ug[k + 1];
      if (r == 0) te[k * 32 + slot] = n2;
      const double aa = cabs2(alpha);
      // MUFU seeds + Newton steps (normal, positive arguments only): a column that is zero to 1e-290 is left alone
      const bool live = n2 > 1e-290, has_a = aa > 1e-290;
      const double rn = live ? fast_rsqrt(n2) : 0.0, ra = has_a ? fast_rsqrt(aa) : 0.0;
      const double xn = n2 * rn, an = aa * ra;
      const cplx ph = has_a ? cscale(alpha, ra) : cmake(1.0, 0.0);
      const cplx u1 = cscale(ph, an + xn);  // u = x - gamma e1, gamma = -ph |x|: H x = gamma e1, H = I - beta u u^dagger
      const double beta = live ? rn * fast_rcp(xn + an) : 0.0;
      const cplx u = (r == k + 1) ? u1 : ((r > k + 1) ? x : cmake(0.0, 0.0));
      cplx p = cmake(0.0, 0.0);
      cfma(p, a[k + 1], u1);
#pragma unroll
      for (int c = k + 2; c < D; ++c) cfma(p, a[c], Ug[c]);
      p = cscale(p, beta);
      const double kk = 0.5 * beta * group_sum<D>(u.x * p.x + u.y * p.y);
      const cplx w2 = cmake(p.x - kk * u.x, p.y - kk * u.y);
      Wg[r] = w2;
      __syncwarp();
#pragma unroll
      for (int c = k + 1; c < D; ++c) {  // A -= u w^dagger + w u^dagger
        const cplx uc = (c == k + 1) ? u1 : Ug[c];
        const cplx wc = Wg[c];
        a[c].x = fma(-w2.y, uc.y, fma(-w2.x, uc.x, fma(-u.y, wc.y, fma(-u.x, wc.x, a[c].x))));
        a[c].y = fma(w2.x, uc.y, fma(-w2.y, uc.x, fma(u.x, wc.y, fma(-u.y, wc.x, a[c].y))));
      }
      __syncwarp();
    }
    double dr = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c)
      if (c == r) dr = a[c].x;
    const double elast = __shfl_sync(0xffffffffu, cabs2(a[D - 2]), glane0 + D - 1);
    td[r * 32 + slot] = (r == 0 && !ok) ? __longlong_as_double(0x7ff8000000000000LL) : dr;
    if (r == 0) te[(D - 2) * 32 + slot] = elast;
    __syncwarp();
  }
  // ---- lane j: eigenvalues of pair b0 + j ----
  const int64_t b = b0 + lane;
  if (b < B) {
    double* d = td + lane;
    double res = FID_FLAG;
    if (d[0] == d[0] && pwk_ql<D>(d, te + lane)) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) acc += sqrt(fmax(d[k * 32], 0.0));
      res = acc * acc;
    }
    out[b] = res;
  }
}

// ---------------------------------------------------------------------------------------------
// fidelity, d = 32 (n = 5): the same Cholesky -> L^dagger sigma L -> tridiagonal -> QL sequence with ONE pair per warp
// (lane r owns row / column r) and the matrix in shared memory: at this size the register-resident rows of
// fidelity_tri_kernel would need 128 registers per matrix row and ~400 KB of unrolled code, so every loop over the
// dimension is rolled here (the trade measured at d = 16 in profiles/r02_ubench_fidelity_tri.txt; against the
// values-only Jacobi kernel, which needs ~10 sweeps at d = 32, it is still a large net gain).  The warp's 32 pairs are
// processed in batches of SLOTS pairs followed by one QL phase on SLOTS lanes: a smaller (d, e^2) slab buys resident warps.
// Measured on 32768 pairs (scripts/ubench_fid.cu 5 32768): 16 slots, one warp per block 4.54 ms; 2 / 4 warps per block
// 4.77 / 4.75 ms; 8 slots 5.0-5.3 ms; 32 slots 7.07 ms; the Jacobi kernel it replaces 15.8 ms.
// ---------------------------------------------------------------------------------------------
template <int SLOTS>
struct FidTri32Smem {
  static constexpr int D = 32, LD = D + 1, MP = D * LD;
  static constexpr size_t bytes = sizeof(cplx) * (MP + 2 * D) + sizeof(double) * 2 * D * SLOTS;
};
#ifndef FID_TRI32_WPB_N
#define FID_TRI32_WPB_N 1
#endif
#ifndef FID_TRI32_SLOTS_N
#define FID_TRI32_SLOTS_N 16
#endif
#ifndef FID_TRI32_MINB
#define FID_TRI32_MINB 8
#endif
constexpr int FID_TRI32_WPB = FID_TRI32_WPB_N, FID_TRI32_SLOTS = FID_TRI32_SLOTS_N;

// y[0..IM) += conj(L[k][0..IM)) * (sigma[k][.] . L[., r]) for k = k0..k1-1 (L[k][i] = 0 for i > k: IM >= k1 suffices)
template <int IM>
__device__ __forceinline__ void fid_tri32_wy(cplx (&y)[32], const cplx* __restrict__ sp, const cplx* __restrict__ Ag,
                                             int r, int k0, int k1) {
  constexpr int D = 32, LD = 33;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    cplx w0 = cmake(0.0, 0.0), w1 = cmake(0.0, 0.0);  // two chains
#pragma unroll 8
    for (int m = 0; m < D; m += 2) {  // sigma[k][m]: every lane reads the same 16 bytes (broadcast); L[m][r]: lane-contiguous
      cfma(w0, sp[k * D + m], Ag[m * LD + r]);
      cfma(w1, sp[k * D + m + 1], Ag[(m + 1) * LD + r]);
    }
    const cplx wk = cadd(w0, w1);
#pragma unroll
    for (int i = 0; i < IM; ++i) cfma(y[i], cconj(Ag[k * LD + i]), wk);
  }
}

template <int SLOTS>
__global__ void __launch_bounds__(32 * FID_TRI32_WPB, FID_TRI32_MINB)
    fidelity_tri32_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma, double* __restrict__ out) {
  constexpr int D = 32, DD = D * D, LD = FidTri32Smem<SLOTS>::LD, MP = FidTri32Smem<SLOTS>::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = threadIdx.x & 31, wib = threadIdx.x >> 5;
  cplx* Ag = reinterpret_cast<cplx*>(smem_raw + FidTri32Smem<SLOTS>::bytes * wib);
  cplx* Ug = Ag + MP;
  cplx* Wg = Ug + D;
  double* td = reinterpret_cast<double*>(Wg + D);
  double* te = td + D * SLOTS;
  const int64_t b0 = ((int64_t)blockIdx.x * FID_TRI32_WPB + wib) * 32;
  if (b0 >= B) return;
#pragma unroll 1
  for (int batch = 0; batch < 32 / SLOTS; ++batch) {
    const int64_t bb0 = b0 + batch * SLOTS;
    if (bb0 >= B) break;
#pragma unroll 1
    for (int slot = 0; slot < SLOTS; ++slot) {
      if (bb0 + slot >= B) break;  // warp-uniform; the slots past the end are never read
      const int64_t b = bb0 + slot;
      const cplx* rp = rho + b * DD;
      const cplx* sp = sigma + b * DD;
      if (b + 1 < B) {  // the next pair's rows on their way into L2 (row r = 512 bytes)
        const char* pr = reinterpret_cast<const char*>(rho + (b + 1) * DD + r * D);
        const char* ps = reinterpret_cast<const char*>(sigma + (b + 1) * DD + r * D);
#pragma unroll
        for (int o = 0; o < D * 16; o += 128) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + o));
        }
      }
      // ---- rho -> shared memory, coalesced (scipy's eigh reads the lower triangle only: so does everything below) ----
#pragma unroll 8
      for (int e = r; e < DD; e += 32) Ag[(e >> 5) * LD + (e & 31)] = rp[e];
      __syncwarp();
      const double dmax = warp_max(Ag[r * LD + r].x);
      const double thr = 1e-10 * dmax;
      bool ok = dmax > 0.0;
      // ---- Cholesky in place; rows above the diagonal of a column are written as zeros ----
#pragma unroll 1
      for (int j = 0; j < D; ++j) {
        const double piv = Ag[j * LD + j].x;
        ok = ok && (piv > thr);
        const double inv = ok ? fast_rsqrt(piv) : 0.0;
        cplx l = cscale(Ag[r * LD + j], inv);
        if (r == j) l = cmake(piv * inv, 0.0);
        if (r < j) l = cmake(0.0, 0.0);
        Ag[r * LD + j] = l;
        __syncwarp();
#pragma unroll 4
        for (int k = D - 1; k > j; --k) {  // row r of the trailing block (rows r < k hold scratch that is zeroed later)
          const cplx lk = Ag[k * LD + j];
          cplx v = Ag[r * LD + k];
          v.x = fma(-l.y, lk.y, fma(-l.x, lk.x, v.x));  // v -= l * conj(lk)
          v.y = fma(l.x, lk.y, fma(-l.y, lk.x, v.y));
          Ag[r * LD + k] = v;
        }
        __syncwarp();
      }
      // ---- Y = L^dagger sigma L, lane r accumulates COLUMN r ----
      cplx y[D];
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = cmake(0.0, 0.0);
      fid_tri32_wy<D / 2>(y, sp, Ag, r, 0, D / 2);
      fid_tri32_wy<D>(y, sp, Ag, r, D / 2, D);
      __syncwarp();
      // row r of the Hermitian matrix an eigensolver reading the lower triangle would see: A[r][c] = conj(Y[c][r])
#pragma unroll
      for (int c = 0; c < D; ++c) Ag[r * LD + c] = (c == r) ? cmake(y[c].x, 0.0) : cconj(y[c]);
      __syncwarp();
      // ---- Householder tridiagonalisation in place, lane r owns row r ----
#pragma unroll 1
      for (int k = 0; k < D - 2; ++k) {
        const cplx x = (r > k) ? Ag[r * LD + k] : cmake(0.0, 0.0);
        const double n2 = warp_sum(cabs2(x));
        const cplx alpha = cmake(__shfl_sync(0xffffffffu, x.x, k + 1), __shfl_sync(0xffffffffu, x.y, k + 1));
        if (r == 0) te[k * SLOTS + slot] = n2;
        const double aa = cabs2(alpha);
        const bool live = n2 > 1e-290, has_a = aa > 1e-290;
        const double rn = live ? fast_rsqrt(n2) : 0.0, ra = has_a ? fast_rsqrt(aa) : 0.0;
        const double xn = n2 * rn, an = aa * ra;
        const cplx ph = has_a ? cscale(alpha, ra) : cmake(1.0, 0.0);
        // u = x - gamma e1 with gamma = -ph |x|:  H = I - beta u u^dagger is unitary and H x = gamma e1
        const cplx u = (r == k + 1) ? cscale(ph, an + xn) : x;
        const double beta = live ? rn * fast_rcp(xn + an) : 0.0;
        Ug[r] = u;
        __syncwarp();
        cplx p0 = cmake(0.0, 0.0), p1 = cmake(0.0, 0.0);
        int c = D - 1;
        for (; c > k + 1; c -= 2) {
          cfma(p0, Ag[r * LD + c], Ug[c]);
          cfma(p1, Ag[r * LD + c - 1], Ug[c - 1]);
        }
        if (c == k + 1) cfma(p0, Ag[r * LD + c], Ug[c]);
        const cplx p = cscale(cadd(p0, p1), beta);
        const double kk = 0.5 * beta * warp_sum(u.x * p.x + u.y * p.y);
        const cplx w2 = cmake(fma(-kk, u.x, p.x), fma(-kk, u.y, p.y));
        Wg[r] = w2;
        __syncwarp();
#pragma unroll 3
        for (int c2 = D - 1; c2 > k; --c2) {  // A -= u w^dagger + w u^dagger
          const cplx uc = Ug[c2], wc = Wg[c2];
          cplx v = Ag[r * LD + c2];
          v.x = fma(-w2.y, uc.y, fma(-w2.x, uc.x, fma(-u.y, wc.y, fma(-u.x, wc.x, v.x))));
          v.y = fma(w2.x, uc.y, fma(-w2.y, uc.x, fma(u.x, wc.y, fma(-u.y, wc.x, v.y))));
          Ag[r * LD + c2] = v;
        }
        __syncwarp();
      }
      const double dr = Ag[r * LD + r].x;
      const double elast = cabs2(Ag[(D - 1) * LD + D - 2]);
      td[r * SLOTS + slot] = (r == 0 && !ok) ? __longlong_as_double(0x7ff8000000000000LL) : dr;
      if (r == 0) te[(D - 2) * SLOTS + slot] = elast;
      __syncwarp();
    }
    // ---- lane j < SLOTS: eigenvalues of pair bb0 + j ----
    const int64_t b = bb0 + r;
    if (r < SLOTS && b < B) {
      double* d = td + r;
      double res = FID_FLAG;
      if (d[0] == d[0] && pwk_ql<D, SLOTS>(d, te + r)) {
        double acc = 0.0;
#pragma unroll 4
        for (int k = 0; k < D; ++k) acc += sqrt(fmax(d[k * SLOTS], 0.0));
        res = acc * acc;
      }
      out[b] = res;
    }
    __syncwarp();
  }
}

// MODE 0: fidelity.  MODE 1: nuclear-norm trace distance.
template <int D, int MODE>
__global__ void fidelity_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                double* __restrict__ out, int only_flagged) {
  constexpr int DD = D * D, LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidSmem<D>::bytes * wib);
  cplx* V = A + MP;
  cplx* W = V + MP;
  double* ev = reinterpret_cast<double*>(W + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  if (only_flagged && out[b] != FID_FLAG) return;  // fidelity_fast_kernel already produced this item
  const cplx* r = rho + b * DD;
  const cplx* s = sigma + b * DD;
  if constexpr (MODE == 1) {
    for (int e = lane; e < DD; e += 32) {
      // Hermitian part of rho - sigma
      const int i = e / D, j = e % D;
      const cplx d1 = csub(r[e], s[e]), d2 = csub(r[j * D + i], s[j * D + i]);
      A[i * LD + j] = cmake(0.5 * (d1.x + d2.x), 0.5 * (d1.y - d2.y));
    }
    __syncwarp();
    jacobi_eigh<D, 32, SyncWarp, false, LD>(A, nullptr, ev, ev + D, lane);
    double acc = 0.0;
    for (int k = lane; k < D; k += 32) acc += fabs(ev[k]);
    acc = warp_sum(acc);
    if (lane == 0) out[b] = 0.5 * acc;
    return;
  }
  // scipy.linalg.eigh reads the lower triangle only: build the Hermitian matrix it sees
  auto load_rho = [&]() {
    for (int e = lane; e < DD; e += 32) {  // coalesced read; the lower triangle is mirrored into the upper one
      const int i = e / D, j = e % D;
      cplx v = r[e];
      if (i == j) v.y = 0.0;
      if (i >= j) {
        A[i * LD + j] = v;
        if (i > j) A[j * LD + i] = cconj(v);
      }
    }
    __syncwarp();
  };
  load_rho();
  jacobi_eigh<D, 32, SyncWarp, true, LD>(A, V, ev, ev + D, lane);
  // S = V sqrt(max(ev,0)) V^dagger  -> A
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) {
      const double w = sqrt(fmax(ev[k], 0.0));
      cfma_conj(acc, cscale(V[i * LD + k], w), V[j * LD + k]);
    }
    A[i * LD + j] = acc;
  }
  __syncwarp();
  // W = S sigma
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma(acc, A[i * LD + k], s[k * D + j]);
    W[i * LD + j] = acc;
  }
  __syncwarp();
  // V = W S, then the Hermitian matrix eigh would see (lower triangle)
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma(acc, W[i * LD + k], A[k * LD + j]);
    V[i * LD + j] = acc;
  }
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = (i >= j) ? V[i * LD + j] : cconj(V[j * LD + i]);
    if (i == j) v.y = 0.0;
    W[i * LD + j] = v;
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, false, LD>(W, nullptr, ev, ev + D, lane);
  double acc = 0.0;
  for (int k = lane; k < D; k += 32) acc += sqrt(fmax(ev[k], 0.0));
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc * acc;
}

// project_state_matrix_to_physical (operator_tools/project_state_matrix.py:6-52), one state per warp:
// rho / tr(rho) -> eigh (lower triangle, like scipy) -> if any eigenvalue is negative, zero the smallest ones and
// spread their mass evenly over the rest (water filling) -> V diag(l') V^dagger.  A matrix that is already PSD
// is returned as rho / tr(rho) unchanged, exactly like the reference.
template <int D>
__global__ void project_state_kernel(int64_t B, const cplx* __restrict__ rho, cplx* __restrict__ out) {
  constexpr int DD = D * D, LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidSmem<D>::bytes * wib);
  cplx* V = A + MP;
  cplx* R = V + MP;  // rho / tr(rho), kept for the already-physical case
  double* ev = reinterpret_cast<double*>(R + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  const cplx* r = rho + b * DD;
  cplx tr = cmake(0.0, 0.0);
  for (int k = lane; k < D; k += 32) tr = cadd(tr, r[k * D + k]);
  tr.x = warp_sum(tr.x);
  tr.y = warp_sum(tr.y);
  const double n2 = cabs2(tr);
  const cplx itr = cmake(tr.x / n2, -tr.y / n2);
  for (int e = lane; e < DD; e += 32) R[(e / D) * LD + e % D] = cmul(r[e], itr);
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = (i >= j) ? R[i * LD + j] : cconj(R[j * LD + i]);
    if (i == j) v.y = 0.0;
    A[i * LD + j] = v;
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, true, LD>(A, V, ev, ev + D, lane);
  double mn = 1e300;
  for (int k = 0; k < D; ++k) mn = fmin(mn, ev[k]);
  cplx* dst = out + b * DD;
  if (mn >= 0.0) {
    for (int e = lane; e < DD; e += 32) dst[e] = R[(e / D) * LD + e % D];
    return;
  }
  // rank of each eigenvalue in DESCENDING order (ties broken by index), then the water-filling scan
  double* desc = ev + D;      // the Jacobi scratch (3 D + 32 doubles) is free now
  double* wgt = ev + 2 * D;   // new eigenvalue of eigenpair k
  __syncwarp();
  for (int k = lane; k < D; k += 32) {
    int rk = 0;
    for (int j = 0; j < D; ++j) rk += (ev[j] > ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
    desc[rk] = ev[k];
  }
  __syncwarp();
  int i = D;
  double acc = 0.0;
  while (i > 1 && desc[i - 1] + acc / (double)i < 0.0) {
    acc += desc[i - 1];
    --i;
  }
  const double shift = acc / (double)i;  // the i largest eigenvalues are kept and shifted
  for (int k = lane; k < D; k += 32) {
    int rk = 0;
    for (int j = 0; j < D; ++j) rk += (ev[j] > ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
    wgt[k] = (rk < i) ? ev[k] + shift : 0.0;
  }
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int a = e / D, c = e % D;
    cplx s = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma_conj(s, cscale(V[a * LD + k], wgt[k]), V[c * LD + k]);
    dst[e] = s;
  }
}

template <int D>
static int launch_project_state(int64_t B, const void* rho, void* out, cudaStream_t st) {
  const size_t per_warp = FidSmem<D>::bytes;
  int wpb = (int)max((size_t)1, min((size_t)8, (size_t)(112 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  QT_CUDA(cudaFuncSetAttribute(project_state_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_state_kernel<D><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho, (cplx*)out);
  return qt_check_launch("project_state_kernel");
}

// Hilbert-Schmidt inner product tr(A^dagger B) per pair (distance_measures.py:198-216); the kernel behind
// entanglement_fidelity / process_fidelity (:271-375).  Pure streaming: one warp per pair for small matrices, one
// block per pair otherwise.  out[b] is complex.
__global__ void hs_inner_kernel(int64_t elems, int64_t B, const cplx* __restrict__ a, const cplx* __restrict__ b,
                                cplx* __restrict__ out, int block_per_pair) {
  __shared__ double red[2][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  if (block_per_pair == 2) {
    // fewer than 32 elements per matrix (a power of two): 32 / elems pairs per warp, one element per lane
    // persistent warps, four steps' loads in flight (see purity_kernel)
    const int gl_n = (int)elems, ipw = 32 / gl_n, gl = lane % gl_n;
    const int64_t nwarps = (int64_t)gridDim.x * wpb, steps = (B + ipw - 1) / ipw;
    for (int64_t s0 = ((int64_t)blockIdx.x * wpb + wib) * 4; s0 < steps; s0 += nwarps * 4) {
      cplx av[4], bv[4];
      int64_t pp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pp[u] = (s0 + u) * ipw + lane / gl_n;
        const bool live = pp[u] < B;
        av[u] = live ? a[pp[u] * elems + gl] : cmake(0.0, 0.0);
        bv[u] = live ? b[pp[u] * elems + gl] : cmake(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cplx acc = cmake(0.0, 0.0);
        cfma(acc, cconj(av[u]), bv[u]);
        for (int o = gl_n / 2; o > 0; o >>= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (pp[u] < B && gl == 0) out[pp[u]] = acc;
      }
    }
    return;
  }
  if (!block_per_pair) {
    const int64_t p = (int64_t)blockIdx.x * wpb + wib;
    if (p >= B) return;
    cplx acc = cmake(0.0, 0.0);
    for (int64_t e = lane; e < elems; e += 32) cfma(acc, cconj(a[p * elems + e]), b[p * elems + e]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) out[p] = acc;
    return;
  }
  for (int64_t p = blockIdx.x; p < B; p += gridDim.x) {
    cplx acc = cmake(0.0, 0.0);
    for (int64_t e = threadIdx.x; e < elems; e += blockDim.x) cfma(acc, cconj(a[p * elems + e]), b[p * elems + e]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    __syncthreads();
    if (lane == 0) {
      red[0][wib] = acc.x;
      red[1][wib] = acc.y;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      cplx t = cmake(0.0, 0.0);
      for (int w = 0; w < wpb; ++w) t = cadd(t, cmake(red[0][w], red[1][w]));
      out[p] = t;
    }
  }
}

template <int D>
static int launch_td(int64_t B, const void* rho, const void* sigma, double* out, cudaStream_t st) {
  const int wpb = 8;
  constexpr int IPW_TD = (D * D < 32) ? 32 / (D * D) : 1;
  const int64_t warps_td = (B + IPW_TD - 1) / IPW_TD;
  trace_distance_kernel<D><<<(unsigned)((warps_td + wpb - 1) / wpb), 32 * wpb, 0, st>>>(B, (const cplx*)rho,
                                                                                 (const cplx*)sigma, out);
  return qt_check_launch("trace_distance_kernel");
}
template <int D>
static int launch_purity(int64_t B, const void* rho, double* out, cudaStream_t st) {
  constexpr int IPW_P = (D * D < 32) ? 32 / (D * D) : 1;
  constexpr int WPB_P = (D >= 32) ? 2 : 8;  // must match purity_kernel
  const int64_t warps_p = (B + IPW_P - 1) / IPW_P;
  int64_t blocks_p = (warps_p + WPB_P - 1) / WPB_P;
  if (D * D < 32) blocks_p = std::min<int64_t>((blocks_p + 3) / 4, (int64_t)QT_NUM_SMS * 32);  // persistent, 4 steps per trip
  purity_kernel<D><<<(unsigned)blocks_p, 32 * WPB_P, 0, st>>>(B, (const cplx*)rho, out);
  return qt_check_launch("purity_kernel");
}
template <int D, int MODE>
static int launch_fid(int64_t B, const void* rho, const void* sigma, double* out, cudaStream_t st) {
  const size_t per_warp = FidSmem<D>::bytes;
  int wpb = (int)max((size_t)1, min((size_t)8, (size_t)(112 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  int only_flagged = 0;
  if constexpr (MODE == 0 && D >= 4) {
    if constexpr (D <= 16) {
      const size_t smem_tri = FidTriSmem<D>::bytes * FID_TRI_WPB;
      QT_CUDA(cudaFuncSetAttribute(fidelity_tri_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tri));
      const int64_t per_block = 32 * FID_TRI_WPB;
      fidelity_tri_kernel<D><<<(unsigned)((B + per_block - 1) / per_block), 32 * FID_TRI_WPB, smem_tri, st>>>(
          B, (const cplx*)rho, (const cplx*)sigma, out);
      int rc = qt_check_launch("fidelity_tri_kernel");
      if (rc) return rc;
      QT_CUDA(cudaFuncSetAttribute(fidelity_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      fidelity_kernel<D, MODE><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho,
                                                                                        (const cplx*)sigma, out, 1);
      return qt_check_launch("fidelity_kernel");
    }
    if constexpr (D == 32) {
      const size_t smem_tri = FidTri32Smem<FID_TRI32_SLOTS>::bytes * FID_TRI32_WPB;
      QT_CUDA(cudaFuncSetAttribute(fidelity_tri32_kernel<FID_TRI32_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem_tri));
      const int64_t per_block = 32 * FID_TRI32_WPB;
      fidelity_tri32_kernel<FID_TRI32_SLOTS><<<(unsigned)((B + per_block - 1) / per_block), 32 * FID_TRI32_WPB, smem_tri, st>>>(
          B, (const cplx*)rho, (const cplx*)sigma, out);
      int rc = qt_check_launch("fidelity_tri32_kernel");
      if (rc) return rc;
      QT_CUDA(cudaFuncSetAttribute(fidelity_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      fidelity_kernel<D, MODE><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho,
                                                                                        (const cplx*)sigma, out, 1);
      return qt_check_launch("fidelity_kernel");
    }
    const size_t per_fast = FidFastSmem<D>::bytes;
    const int wf = (int)max((size_t)1, min((size_t)8, (size_t)(74 * 1024) / per_fast));
    QT_CUDA(cudaFuncSetAttribute(fidelity_fast_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_fast * wf)));
    fidelity_fast_kernel<D><<<(unsigned)((B + wf - 1) / wf), 32 * wf, per_fast * wf, st>>>(B, (const cplx*)rho,
                                                                                          (const cplx*)sigma, out);
    int rc = qt_check_launch("fidelity_fast_kernel");
    if (rc) return rc;
    only_flagged = 1;
  }
  QT_CUDA(cudaFuncSetAttribute(fidelity_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fidelity_kernel<D, MODE><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho,
                                                                                    (const cplx*)sigma, out, only_flagged);
  return qt_check_launch("fidelity_kernel");
}

#define DISPATCH_D(n, CALL)                      \
  switch (n) {                                   \
    case 1: return CALL(2);                      \
    case 2: return CALL(4);                      \
    case 3: return CALL(8);                      \
    case 4: return CALL(16);                     \
    case 5: return CALL(32);                     \
    default:                                     \
      qt_set_error("n=%d out of range 1..5", n); \
      return QT_ERR_ARG;                         \
  }

extern "C" int qt_trace_distance_batch(int n, int64_t B, const void* rho, const void* sigma, double* out,
                                       void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_trace_distance_batch: null argument");
#define CALL(D) launch_td<D>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_trace_distance_nuclear_batch(int n, int64_t B, const void* rho, const void* sigma, double* out,
                                               void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_trace_distance_nuclear_batch: null argument");
#define CALL(D) launch_fid<D, 1>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_fidelity_batch(int n, int64_t B, const void* rho, const void* sigma, double* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_fidelity_batch: null argument");
#define CALL(D) launch_fid<D, 0>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_purity_batch(int n, int64_t B, const void* rho, double* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && out, "qt_purity_batch: null argument");
#define CALL(D) launch_purity<D>(B, rho, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_project_state_batch(int n, int64_t B, const void* rho, void* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && out, "qt_project_state_batch: null argument");
#define CALL(D) launch_project_state<D>(B, rho, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_hs_inner_batch(int64_t rows, int64_t cols, int64_t B, const void* a, const void* b, void* out,
                                 void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rows > 0 && cols > 0 && a && b && out, "qt_hs_inner_batch: bad arguments");
  const int64_t elems = rows * cols;
  const bool packed = elems < 32 && (elems & (elems - 1)) == 0;
  const int block_per_pair = packed ? 2 : (elems > 1024 ? 1 : 0);
  const int64_t blocks = packed ? std::min<int64_t>((B + 32 * (32 / elems) - 1) / (32 * (32 / elems)), (int64_t)QT_NUM_SMS * 32)
                                : (block_per_pair ? std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 8) : (B + 7) / 8);
  hs_inner_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(elems, B, (const cplx*)a, (const cplx*)b,
                                                                      (cplx*)out, block_per_pair);
  return qt_check_launch("hs_inner_kernel");
}
