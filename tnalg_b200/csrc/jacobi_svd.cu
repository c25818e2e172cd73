// One-sided (Hestenes) Jacobi SVD in FP64  (a7 'svd' branch, a11, a12).
//
// The rows of a work matrix W are rotated pairwise until they are mutually orthogonal:
//   m >= n : W = A^T (n rows of length m), the accumulated rotations give V^T, U = normalised rows
//   m <  n : W = A   (m rows of length n), the accumulated rotations give U^T, V^T = normalised rows
// so every dot product and every rotation streams over contiguous memory (coalesced, HBM/L2-bound).
// Blocked schedule (default): the rows are cut into blocks of w rows; one launch per round of a round-robin (chess
// tournament) over the BLOCKS, one CTA per block pair.  The CTA stages its 2w rows of W and of the rotation accumulator
// in shared memory with 1-D bulk copies (cp.async.bulk + mbarrier), rotates every row pair between the two blocks there
// (w rounds of w disjoint pairs, 16/w warps per pair, warp-shuffle reductions for |p|^2, |q|^2, p.q) and bulk-stores
// the rows back: each global round trip serves w^2 rotations instead of one and a sweep needs rows/w - 1 launches
// instead of rows - 1.  Pairs inside a block are rotated in the first round of every sweep.  w is the largest power of
// two whose 2w rows fit 200 KB of shared memory (w = 8 for 512 rows with vectors, 4 for 1024, 2 for 2048); rows too
// long for that use the unblocked kernel (one CTA per row pair, rows in global memory / L2).
// The sweep loop stops when a whole sweep applied no rotation.  Singular values are sorted on the host (k doubles), the gather of the k_keep leading triplets is a
// kernel.  Jacobi gives high relative accuracy for the small Schmidt values the entanglement spectrum needs.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "host_math.h"

namespace tn {

constexpr int kJacThreads = 256;

__device__ __forceinline__ double warp_sum_j(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// W: rows x len (row pitch ldw), Acc: rows x rows_acc (row pitch lda).  round in [0, rows-1), rows even.
__global__ void __launch_bounds__(kJacThreads) jacobi_round_kernel(double* __restrict__ W, long long ldw, int len,
                                                                   double* __restrict__ Acc, long long lda, int acc_len,
                                                                   int rows, int round, double tol, unsigned* n_rot) {
  const int i = blockIdx.x;  // pair index
  const int nm1 = rows - 1;
  int p, q;
  if (i == 0) {
    p = round % nm1;
    q = nm1;
  } else {
    p = (round + i) % nm1;
    q = (round + nm1 - i) % nm1;
  }
  if (p > q) { int tmp = p; p = q; q = tmp; }
  double* wp = W + (long long)p * ldw;
  double* wq = W + (long long)q * ldw;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int e = threadIdx.x; e < len; e += kJacThreads) {
    const double x = wp[e], y = wq[e];
    a += x * x;
    b += y * y;
    g += x * y;
  }
  a = warp_sum_j(a); b = warp_sum_j(b); g = warp_sum_j(g);
  __shared__ double red[3][kJacThreads / 32];
  __shared__ double cs[2];
  __shared__ int do_rot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = g; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double A_ = 0, B_ = 0, G_ = 0;
#pragma unroll
    for (int k = 0; k < kJacThreads / 32; ++k) { A_ += red[0][k]; B_ += red[1][k]; G_ += red[2][k]; }
    int rot = (A_ > 0.0 && B_ > 0.0 && fabs(G_) > tol * sqrt(A_) * sqrt(B_)) ? 1 : 0;
    if (rot) {
      double c, s;
      jacobi_rotation(A_, B_, G_, &c, &s);
      cs[0] = c; cs[1] = s;
      atomicAdd(n_rot, 1u);
    }
    do_rot = rot;
  }
  __syncthreads();
  if (!do_rot) return;
  const double c = cs[0], s = cs[1];
  for (int e = threadIdx.x; e < len; e += kJacThreads) {
    const double x = wp[e], y = wq[e];
    wp[e] = c * x - s * y;
    wq[e] = s * x + c * y;
  }
  if (Acc) {
    double* ap = Acc + (long long)p * lda;
    double* aq = Acc + (long long)q * lda;
    for (int e = threadIdx.x; e < acc_len; e += kJacThreads) {
      const double x = ap[e], y = aq[e];
      ap[e] = c * x - s * y;
      aq[e] = s * x + c * y;
    }
  }
}


// ---- blocked variant: 2w rows resident in shared memory ----
constexpr int kBlkThreads = 512;
constexpr int kBlkWarps = kBlkThreads / 32;
constexpr size_t kBlkSmemBudget = 200 * 1024;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// chess-tournament pair i (0 <= i < rows/2) of round `round` (0 <= round < rows-1), rows even
__host__ __device__ __forceinline__ void tournament_pair(int rows, int round, int i, int* p, int* q) {
  const int nm1 = rows - 1;
  int a, b;
  if (i == 0) {
    a = round % nm1;
    b = nm1;
  } else {
    a = (round + i) % nm1;
    b = (round + nm1 - i) % nm1;
  }
  *p = a < b ? a : b;
  *q = a < b ? b : a;
}

// W: rp x ldw, Acc: rp x lda (or null), rp = nb*w rows, nb even.  Launch nb/2 CTAs for block round `round` in
// [0, nb-1).  full != 0: also rotate the pairs inside each block (first round of a sweep).
__global__ void __launch_bounds__(kBlkThreads, 1) jacobi_block_kernel(double* __restrict__ W, int ldw, int len, double* __restrict__ Acc, int lda,
                                                                      int w, int nb, int round, int full, double tol, unsigned* n_rot) {
  extern __shared__ __align__(128) double sm[];
  const int acc_len = Acc ? lda : 0;
  const int rowlen = ldw + acc_len;
  double* part = sm + (size_t)2 * w * rowlen;                       // [w pairs][kBlkWarps][3]
  uint64_t* bar = reinterpret_cast<uint64_t*>(part + w * kBlkWarps * 3);
  int* any_rot = reinterpret_cast<int*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bi, bj;
  tournament_pair(nb, round, blockIdx.x, &bi, &bj);
  auto grow = [&](int r) { return (long long)(r < w ? bi * w + r : bj * w + (r - w)); };

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *any_rot = 0;
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)(2 * w * rowlen * sizeof(double)))
                   : "memory");
    __syncwarp();
    for (int r = lane; r < 2 * w; r += 32) {
      const long long gr = grow(r);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sm + (size_t)r * rowlen)),
                   "l"(W + gr * ldw), "r"((unsigned)(ldw * sizeof(double))), "r"(smem_addr(bar))
                   : "memory");
      if (acc_len)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(sm + (size_t)r * rowlen + ldw)),
                     "l"(Acc + gr * lda), "r"((unsigned)(lda * sizeof(double))), "r"(smem_addr(bar))
                     : "memory");
    }
  }
  asm volatile(
      "{\n.reg .pred p;\nTN_JWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@!p bra TN_JWAIT_%=;\n}\n" ::"r"(smem_addr(bar))
      : "memory");

  const int wpp = kBlkWarps / w;          // warps per row pair
  const int pr = warp / wpp, sw = warp % wpp;
  const int n_rounds = full ? 2 * w - 1 : w;
  unsigned my_rot = 0;
  for (int r = 0; r < n_rounds; ++r) {
    int p, q;
    if (full) {
      tournament_pair(2 * w, r, pr, &p, &q);
    } else {
      p = pr;
      q = w + (pr + r) % w;
    }
    double* rp_ = sm + (size_t)p * rowlen;
    double* rq_ = sm + (size_t)q * rowlen;
    // rows are 16-byte aligned and zero-padded to an even length: double2 accesses, four independent loads in flight
    double2* const p2 = reinterpret_cast<double2*>(rp_);
    double2* const q2 = reinterpret_cast<double2*>(rq_);
    const int stride = 32 * wpp, e0 = sw * 32 + lane;
    double a = 0.0, b = 0.0, g = 0.0;
    {
      const int n2 = ldw >> 1;
      int e = e0;
      for (; e + 3 * stride < n2; e += 4 * stride) {
        double2 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = p2[e + u * stride]; y[u] = q2[e + u * stride]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a += x[u].x * x[u].x + x[u].y * x[u].y;
          b += y[u].x * y[u].x + y[u].y * y[u].y;
          g += x[u].x * y[u].x + x[u].y * y[u].y;
        }
      }
      for (; e < n2; e += stride) {
        const double2 x = p2[e], y = q2[e];
        a += x.x * x.x + x.y * x.y;
        b += y.x * y.x + y.y * y.y;
        g += x.x * y.x + x.y * y.y;
      }
    }
    a = warp_sum_j(a); b = warp_sum_j(b); g = warp_sum_j(g);
    if (wpp > 1) {  // CTA-uniform
      if (lane == 0) {
        double* o = part + (pr * kBlkWarps + sw) * 3;
        o[0] = a; o[1] = b; o[2] = g;
      }
      __syncthreads();
      a = b = g = 0.0;
      for (int k = 0; k < wpp; ++k) {  // fixed order: every warp of the pair gets bit-identical sums
        const double* o = part + (pr * kBlkWarps + k) * 3;
        a += o[0]; b += o[1]; g += o[2];
      }
    }
    if (a > 0.0 && b > 0.0 && g * g > (tol * tol) * a * b) {  // warp-uniform; |g| > tol |p| |q| without the square roots
      double c, s;
      jacobi_rotation(a, b, g, &c, &s);
      const int n2 = rowlen >> 1;
      int e = e0;
      for (; e + 3 * stride < n2; e += 4 * stride) {
        double2 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = p2[e + u * stride]; y[u] = q2[e + u * stride]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          p2[e + u * stride] = make_double2(c * x[u].x - s * y[u].x, c * x[u].y - s * y[u].y);
          q2[e + u * stride] = make_double2(s * x[u].x + c * y[u].x, s * x[u].y + c * y[u].y);
        }
      }
      for (; e < n2; e += stride) {
        const double2 x = p2[e], y = q2[e];
        p2[e] = make_double2(c * x.x - s * y.x, c * x.y - s * y.y);
        q2[e] = make_double2(s * x.x + c * y.x, s * x.y + c * y.y);
      }
      if (sw == 0 && lane == 0) { ++my_rot; *any_rot = 1; }
    }
    __syncthreads();
  }
  if (my_rot) atomicAdd(n_rot, my_rot);
  if (*any_rot && warp == 0) {  // write the rows back (generic-proxy smem writes -> async-proxy reads need the fence)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = lane; r < 2 * w; r += 32) {
      const long long gr = grow(r);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(W + gr * ldw), "r"(smem_addr(sm + (size_t)r * rowlen)),
                   "r"((unsigned)(ldw * sizeof(double)))
                   : "memory");
      if (acc_len)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Acc + gr * lda),
                     "r"(smem_addr(sm + (size_t)r * rowlen + ldw)), "r"((unsigned)(lda * sizeof(double)))
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}


// ---- register-resident blocked variant ----
// The 2W rows of a block pair live in REGISTERS: thread t of 512 holds columns t, t + 512, ... (C per row) of every row, so
// a rotation is thread-local and needs no data movement at all; only the 3W dot products of a round are reduced across the
// CTA (packed reduce-scatter with warp shuffles, 9 shuffles per 2 pairs, then one shared-memory hop across the 16 warps).
// The pair schedule is unrolled at compile time so that every register index is static.  W x C = 16: (16,1) rows up to 512
// doubles (work row + accumulator row), (8,2) up to 1024, (4,4) up to 2048.  Only (8,2) and (4,4) are instantiated: rows of at
// most 512 doubles run faster on the shared-memory kernel with w = 16.
constexpr int kRegThreads = 512;
constexpr int kRegWarps = kRegThreads / 32;

// position k of the circle-method cycle over the 2W - 1 moving rows: 1, 2, .., W-1, 2W-1, 2W-2, .., W
template <int W>
__host__ __device__ constexpr int reg_cycle(int k) {
  return k < W - 1 ? k + 1 : (2 * W - 1) - (k - (W - 1));
}

template <int W, int C, bool FULL>
__global__ void __launch_bounds__(kRegThreads, 1) jacobi_block_reg_kernel(double* __restrict__ Wm, int ldw, double* __restrict__ Acc, int lda,
                                                                          int nb, int round, double tol, unsigned* n_rot) {
  constexpr int R = 2 * W;
  static_assert(W * C == 16 && W >= 2 && W % 2 == 0, "register tile: 32 doubles per thread");
  __shared__ double part[kRegWarps][W * 4];
  __shared__ double cs[W][2];
  __shared__ int any_rot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int acc_len = Acc ? lda : 0;
  int bi, bj;
  tournament_pair(nb, round, blockIdx.x, &bi, &bj);

  double reg[R][C];
  double msk[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int e = tid + kRegThreads * c;
    msk[c] = e < ldw ? 1.0 : 0.0;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long gr = (long long)(r < W ? bi * W + r : bj * W + (r - W));
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = tid + kRegThreads * c;
      double v = 0.0;
      if (e < ldw) v = Wm[gr * ldw + e];
      else if (e - ldw < acc_len) v = Acc[gr * lda + (e - ldw)];
      reg[r][c] = v;
    }
  }
  if (tid == 0) any_rot = 0;
  unsigned my_rot = 0;
  const int n_wcols = (ldw + kRegThreads - 1) / kRegThreads;
  constexpr int NR = FULL ? R - 1 : W;
  // The pairs are always (i, W + i); between rounds the ROWS move through the registers instead (circle method: row 0
  // fixed, the others advance one place along 1, 2, .., W-1, 2W-1, 2W-2, .., W; inter-block schedule: the second block
  // shifts by one).  After NR rounds every row is back in its place.  The round body is therefore identical code and the
  // loop stays rolled: the fully unrolled schedule spent 25 % of its samples on instruction-cache misses (ncu no_inst).
#pragma unroll 1
  for (int r = 0; r < NR; ++r) {
    // ---- partial dot products, two pairs (8 padded values) at a time, packed warp reduce-scatter ----
#pragma unroll
    for (int g2 = 0; g2 < W / 2; ++g2) {
      double v[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = 2 * g2 + h, q = W + 2 * g2 + h;
        double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (c < n_wcols) {  // CTA-uniform: columns beyond the work matrix hold the accumulator
            const double x = reg[p][c] * msk[c], y = reg[q][c] * msk[c];
            a += x * x;
            b += y * y;
            g += x * y;
          }
        }
        v[4 * h + 0] = a; v[4 * h + 1] = b; v[4 * h + 2] = g; v[4 * h + 3] = 0.0;
      }
      // halving steps (offsets 16, 8, 4): lane keeps the upper half when its bit is set
#pragma unroll
      for (int st = 0; st < 3; ++st) {
        const int o = 16 >> st, n = 4 >> st;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double send = up ? v[i] : v[i + n];
          const double keep = up ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      if ((lane & 3) == 0) part[warp][8 * g2 + (lane >> 2)] = v[0];   // value index = lane >> 2
    }
    __syncthreads();
    // ---- warp i < W: totals of pair i over the 16 warps, rotation parameters ----
    if (warp < W) {
      double a = 0.0, b = 0.0, g = 0.0;
      if (lane < kRegWarps) {
        a = part[lane][4 * warp + 0];
        b = part[lane][4 * warp + 1];
        g = part[lane][4 * warp + 2];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      a = __shfl_sync(0xffffffffu, a, 0); b = __shfl_sync(0xffffffffu, b, 0); g = __shfl_sync(0xffffffffu, g, 0);
      double c = 1.0, s = 0.0;
      if (a > 0.0 && b > 0.0 && g * g > (tol * tol) * a * b) {
        jacobi_rotation(a, b, g, &c, &s);
        if (lane == 0) { ++my_rot; any_rot = 1; }
      }
      if (lane == 0) { cs[warp][0] = c; cs[warp][1] = s; }
    }
    __syncthreads();
    // ---- thread-local rotations ----
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int p = i, q = W + i;
      const double c = cs[i][0], s = cs[i][1];
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        const double x = reg[p][cc], y = reg[q][cc];
        reg[p][cc] = c * x - s * y;
        reg[q][cc] = s * x + c * y;
      }
    }
    // ---- move the rows to their places of the next round ----
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
      if (FULL) {
        constexpr int NC = R - 1;                                   // cycle 1, 2, .., W-1, 2W-1, .., W
        const double last = reg[reg_cycle<W>(NC - 1)][cc];
#pragma unroll
        for (int k = NC - 1; k >= 1; --k) reg[reg_cycle<W>(k)][cc] = reg[reg_cycle<W>(k - 1)][cc];
        reg[reg_cycle<W>(0)][cc] = last;
      } else {
        const double first = reg[W][cc];
#pragma unroll
        for (int k = 0; k < W - 1; ++k) reg[W + k][cc] = reg[W + k + 1][cc];
        reg[R - 1][cc] = first;
      }
    }
    // `part` / `cs` are rewritten only after the next round's first barrier / second barrier respectively
  }
  if (my_rot) atomicAdd(n_rot, my_rot);
  __syncthreads();
  if (any_rot) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long gr = (long long)(r < W ? bi * W + r : bj * W + (r - W));
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int e = tid + kRegThreads * c;
        if (e < ldw) Wm[gr * ldw + e] = reg[r][c];
        else if (e - ldw < acc_len) Acc[gr * lda + (e - ldw)] = reg[r][c];
      }
    }
  }
}

// out (rows_out x cols_out, pitch ldo) = in^T, tiled through shared memory
__global__ void transpose_kernel(const double* __restrict__ in, int rows_in, int cols_in, double* __restrict__ out, long long ldo) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int gi = by + r, gj = bx + threadIdx.x;
    if (gi < rows_in && gj < cols_in) tile[r][threadIdx.x] = in[(long long)gi * cols_in + gj];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int gi = bx + r, gj = by + threadIdx.x;  // out[gi][gj] = in[gj][gi]
    if (gi < cols_in && gj < rows_in) out[(long long)gi * ldo + gj] = tile[threadIdx.x][r];
  }
}

__global__ void set_identity_kernel(double* M, int rows, long long ld) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * ld; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / ld, c = i % ld;
    M[i] = (r == c) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(kJacThreads) row_norms_kernel(const double* __restrict__ W, long long ldw, int len, double* __restrict__ sig) {
  const double* w = W + (long long)blockIdx.x * ldw;
  double a = 0.0;
  for (int e = threadIdx.x; e < len; e += kJacThreads) a += w[e] * w[e];
  a = warp_sum_j(a);
  __shared__ double red[kJacThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < kJacThreads / 32; ++k) s += red[k];
    sig[blockIdx.x] = sqrt(s);
  }
}

// dst row jj (length len, pitch ldd) = scale(jj) * src row perm[jj];  scale = 1/sig[perm[jj]] when normalise != 0
__global__ void gather_rows_kernel(const double* __restrict__ src, long long lds, const int* __restrict__ perm, const double* __restrict__ sig,
                                   int normalise, int len, double* __restrict__ dst, long long ldd) {
  const int jj = blockIdx.x, r = perm[jj];
  double sc = 1.0;
  if (normalise) sc = sig[r] > 1e-300 ? 1.0 / sig[r] : 0.0;
  for (int e = threadIdx.x; e < len; e += blockDim.x) dst[(long long)jj * ldd + e] = sc * src[(long long)r * lds + e];
}

// dst (len x k, row-major) column jj = scale(jj) * src row perm[jj]   (transposing gather)
__global__ void gather_cols_kernel(const double* __restrict__ src, long long lds, const int* __restrict__ perm, const double* __restrict__ sig,
                                   int normalise, int len, int k, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int j0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int jj = j0 + r, e = e0 + threadIdx.x;
    double v = 0.0;
    if (jj < k && e < len) {
      const int row = perm[jj];
      double sc = 1.0;
      if (normalise) sc = sig[row] > 1e-300 ? 1.0 / sig[row] : 0.0;
      v = sc * src[(long long)row * lds + e];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int e = e0 + r, jj = j0 + threadIdx.x;
    if (e < len && jj < k) dst[(long long)e * k + jj] = tile[threadIdx.x][r];
  }
}

__global__ void gather_sig_kernel(const double* sig, const int* perm, int k, double* S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) S[i] = sig[perm[i]];
}

}  // namespace tn

using namespace tn;

struct SvdGeom {
  int rows, len, w, rp;   // w = block rows of the blocked schedules (0: unblocked kernel), rp = padded row count
  long long ldw;
  size_t smem;
  bool in_registers;      // register-resident kernel (w in {16, 8, 4}) instead of the shared-memory one
};

static size_t block_smem_bytes(int w, long long rowlen) {
  return sizeof(double) * ((size_t)2 * w * rowlen + (size_t)w * kBlkWarps * 3) + 16;
}

static SvdGeom svd_geom(int m, int n) {
  SvdGeom g;
  g.rows = std::min(m, n);
  g.len = std::max(m, n);
  g.ldw = (long long)(g.len + 1) / 2 * 2;
  g.w = 0;
  g.rp = (g.rows + 1) / 2 * 2;
  g.smem = 0;
  g.in_registers = false;
  const char* e = getenv("TNALG_SVD_UNBLOCKED");
  if (e && e[0] == '1') return g;
  const char* e2 = getenv("TNALG_SVD_SMEM");
  if (!(e2 && e2[0] == '1')) {
    // register-resident blocks: 2w rows x 512*(16/w) columns per CTA, work row + accumulator row side by side
    // (rows short enough for w = 16 are faster on the shared-memory kernel: measured 6.8 vs 8.9 ms at 512 x 256)
    for (int w = 8; w >= 4 && g.ldw + g.rp > kRegThreads; w >>= 1) {
      const int rp = (g.rows + 2 * w - 1) / (2 * w) * (2 * w);
      if (g.ldw + rp <= (long long)kRegThreads * (16 / w)) {
        g.w = w;
        g.rp = rp;
        g.in_registers = true;
        return g;
      }
    }
  }
  int w0 = 16;
  while (w0 > 1 && w0 >= g.rows) w0 >>= 1;
  for (int w = w0; w >= 1; w >>= 1) {
    const int rp = (g.rows + 2 * w - 1) / (2 * w) * (2 * w);
    const size_t bytes = block_smem_bytes(w, g.ldw + rp);
    if (bytes <= kBlkSmemBudget) {
      // measured on B200 (profiles/r01_svd.md): with w < 4 the staging cost per launch outweighs the saved launches
      if (w < 4 && g.rows > 64) break;
      g.w = w;
      g.rp = rp;
      g.smem = bytes;
      break;
    }
  }
  return g;
}

extern "C" size_t tn_svd_workspace_bytes(int m, int n) {
  const SvdGeom g = svd_geom(m, n);
  const size_t rp = (size_t)(g.rows + 31) / 32 * 32 + 32;  // covers the padding of every kernel variant (env switches)
  return align_up(sizeof(double) * rp * (size_t)g.ldw) + align_up(sizeof(double) * rp * rp) + align_up(sizeof(double) * rp) +
         align_up(sizeof(int) * rp) + align_up(sizeof(unsigned)) + 1024;
}

extern "C" int tn_svd_jacobi(const double* A, int m, int n, int k_keep, double* U, double* S, double* Vt, int* sweeps_out,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && S && m > 0 && n > 0, "tn_svd_jacobi: bad arguments");
  const SvdGeom geo = svd_geom(m, n);
  const int rows = geo.rows, rp = geo.rp, len = geo.len;
  TN_REQUIRE(k_keep >= 1 && k_keep <= rows, "tn_svd_jacobi: k_keep=%d not in 1..%d", k_keep, rows);
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_svd_jacobi: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_svd_workspace_bytes(m, n)) {
    set_error("tn_svd_jacobi: workspace %zu < %zu bytes", workspace_bytes, tn_svd_workspace_bytes(m, n));
    return TN_ERR_WORKSPACE;
  }
  const long long ldw = (long long)(len + 1) / 2 * 2;
  Carver cw(workspace, workspace_bytes);
  double* W = cw.take<double>((size_t)rp * ldw);
  double* Acc = cw.take<double>((size_t)rp * rp);
  double* sig = cw.take<double>(rp);
  int* perm = cw.take<int>(rp);
  unsigned* n_rot = cw.take<unsigned>(1);
  TN_REQUIRE(W && Acc && sig && perm && n_rot, "tn_svd_jacobi: workspace carve failed");
  const bool tall = m >= n;  // W = A^T
  const bool need_acc = tall ? (Vt != nullptr) : (U != nullptr);

  TN_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)rp * ldw, stream));
  if (tall) {
    dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, stream>>>(A, m, n, W, ldw);
    TN_LAUNCHED();
  } else {
    TN_CUDA(cudaMemcpy2DAsync(W, sizeof(double) * ldw, A, sizeof(double) * n, sizeof(double) * n, m, cudaMemcpyDeviceToDevice, stream));
  }
  if (need_acc) {
    set_identity_kernel<<<std::min(1024, (rp * rp + 255) / 256), 256, 0, stream>>>(Acc, rp, rp);
    TN_LAUNCHED();
  }
  if (geo.w > 0 && !geo.in_registers) {
    static bool configured = false;
    if (!configured) {
      TN_CUDA(cudaFuncSetAttribute(jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
  }
  const double tol = std::max(1e-15, std::sqrt((double)len) * 2.2e-16);
  int sweeps = 0;
  const int max_sweeps = 80;  // QR-preconditioned inputs (ops.CudaBackend.svd) need ~8; raw ill-conditioned ones 20-50
  bool converged = rp < 2;
  std::vector<double> hs(rp);
  auto row_norms = [&]() -> int {
    row_norms_kernel<<<rp, kJacThreads, 0, stream>>>(W, ldw, len, sig);
    TN_LAUNCHED();
    TN_CUDA(cudaMemcpyAsync(hs.data(), sig, sizeof(double) * rp, cudaMemcpyDeviceToHost, stream));
    TN_CUDA(cudaStreamSynchronize(stream));
    return TN_OK;
  };
  auto run_sweeps = [&]() -> int {
    converged = rp < 2;
    while (!converged && sweeps < max_sweeps) {
      TN_CUDA(cudaMemsetAsync(n_rot, 0, sizeof(unsigned), stream));
      if (geo.in_registers) {
        const int nb = rp / geo.w;
        double* acc = need_acc ? Acc : nullptr;
        for (int round = 0; round < nb - 1; ++round) {
          const bool full = round == 0;
#define TN_JREG(WW, CC)                                                                                                              \
  if (full) jacobi_block_reg_kernel<WW, CC, true><<<nb / 2, kRegThreads, 0, stream>>>(W, (int)ldw, acc, rp, nb, round, tol, n_rot); \
  else jacobi_block_reg_kernel<WW, CC, false><<<nb / 2, kRegThreads, 0, stream>>>(W, (int)ldw, acc, rp, nb, round, tol, n_rot)
          if (geo.w == 8) { TN_JREG(8, 2); }
          else { TN_JREG(4, 4); }
#undef TN_JREG
          TN_LAUNCHED();
        }
      } else if (geo.w > 0) {
        const int nb = rp / geo.w;
        for (int round = 0; round < nb - 1; ++round) {
          jacobi_block_kernel<<<nb / 2, kBlkThreads, geo.smem, stream>>>(W, (int)ldw, len, need_acc ? Acc : nullptr, rp, geo.w, nb, round,
                                                                         round == 0 ? 1 : 0, tol, n_rot);
          TN_LAUNCHED();
        }
      } else {
        for (int round = 0; round < rp - 1; ++round) {
          jacobi_round_kernel<<<rp / 2, kJacThreads, 0, stream>>>(W, ldw, len, need_acc ? Acc : nullptr, rp, rp, rp, round, tol, n_rot);
          TN_LAUNCHED();
        }
      }
      unsigned h_rot = 0;
      TN_CUDA(cudaMemcpyAsync(&h_rot, n_rot, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
      TN_CUDA(cudaStreamSynchronize(stream));
      ++sweeps;
      converged = (h_rot == 0);
    }
    return TN_OK;
  };
  // (A shortcut that left pairs of round-off-level rows un-rotated in truncating calls was tried and removed: excluding
  // pairs breaks the convergence of the cyclic process -- a mid-size row keeps being re-rotated against two tiny rows that
  // are never made orthogonal to each other -- and it saved no sweeps on real two-site wavefunctions.)
  TN_CHECK(run_sweeps());
  TN_CHECK(row_norms());
  if (sweeps_out) *sweeps_out = sweeps;
  std::vector<int> hp(rows);
  std::iota(hp.begin(), hp.end(), 0);
  std::stable_sort(hp.begin(), hp.end(), [&](int x, int y) { return hs[x] > hs[y]; });
  TN_CUDA(cudaMemcpyAsync(perm, hp.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, stream));
  gather_sig_kernel<<<(k_keep + 255) / 256, 256, 0, stream>>>(sig, perm, k_keep, S);
  TN_LAUNCHED();
  dim3 tb(32, 8);
  if (tall) {
    // U (m x k_keep) columns = normalised rows of W; Vt (k_keep x n) rows = rows of Acc
    if (U) {
      dim3 grid((k_keep + 31) / 32, (m + 31) / 32);
      gather_cols_kernel<<<grid, tb, 0, stream>>>(W, ldw, perm, sig, 1, m, k_keep, U);
      TN_LAUNCHED();
    }
    if (Vt) {
      gather_rows_kernel<<<k_keep, 256, 0, stream>>>(Acc, rp, perm, sig, 0, n, Vt, n);
      TN_LAUNCHED();
    }
  } else {
    // U (m x k_keep) columns = rows of Acc (U = Acc^T); Vt rows = normalised rows of W
    if (U) {
      dim3 grid((k_keep + 31) / 32, (m + 31) / 32);
      gather_cols_kernel<<<grid, tb, 0, stream>>>(Acc, rp, perm, sig, 0, m, k_keep, U);
      TN_LAUNCHED();
    }
    if (Vt) {
      gather_rows_kernel<<<k_keep, 256, 0, stream>>>(W, ldw, perm, sig, 1, n, Vt, n);
      TN_LAUNCHED();
    }
  }
  TN_CUDA(cudaStreamSynchronize(stream));  // hp / hs are host temporaries
  if (!converged) {
    set_error("tn_svd_jacobi: not converged after %d sweeps", sweeps);
    return TN_ERR_NOCONV;
  }
  return TN_OK;
}
