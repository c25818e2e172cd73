// Effective-Hamiltonian plan + matvec (a1/a2, MPSClass.py:684-776) and batched environment update
// (a5, TensorBasicModule.py:530-619), both expressed as chain-GEMM launches.
//
// matvec:   out = c_id psi + c_h [ M psi + HL psi + sum_k LS_k op_k psi            (left stage,  TN_NN)
//                                  + psi HR^T + sum_k op_k psi RS_k^T               (right stage, TN_NT)
//                                  + sum_i c_i (XL_i psi) XR_i^T ]                  (left stage writes Phi_i = c_i XL_i psi,
//                                                                                    right stage consumes it)
// HBM layout: psi is (a,d,b) C-order; the left stage sees it as a (a, d*b) matrix whose column index is (s,b),
// the right stage as a (a*d, b) matrix whose row index is (a,s) -- no transpose or reshape is ever materialised,
// the d x d site operators are applied while MMA fragments are built.
#include <algorithm>
#include <cstring>
#include <vector>

#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "vector_ops.cuh"

using namespace tn;

struct tn_effh_plan {
  int a, d, b;
  long long n;
  int rank, world;
  // row-sliced plan (multi-GPU 'rows' sharding): this plan computes rows [row_begin, row_begin + a_out) of the (a, d*b)
  // output from the FULL input psi; a_out == a and row_begin == 0 otherwise
  bool rows;
  int row_begin, a_out;
  long long off, n_out;  // element offset of the slice inside psi, slice length
  // d > TN_MAX_LOADPATH_DIM (two-site window of spin-1, d = 9): the site operators are applied by one element-wise pass per
  // operator into pre_buf before the GEMMs, which then run without an operator (launch d = 1)
  bool preop;
  int n_pre_l, n_pre_r;
  std::vector<SiteOp> pre_ops;
  double* pre_buf;
  SiteOp M;
  bool has_M;
  bool haveA, haveB;
  GemmLaunch LA, LB;
  GemmSchedule SA, SB;
  ProblemDev *probA, *probB;
  LinkDev *linkA, *linkB;
  double* phi;
  int n_phi;
  double alg_flops, exec_flops;
  // TMA staging (chain_gemm_tma.cu): tensor maps of the fixed operands; psi maps are encoded per call
  bool tmaA, tmaB;
  TmaMap* maps;
  // small plans without crossing terms: the two stages are independent and latency bound, so they run concurrently
  bool overlapAB;
};

// side stream + events for the concurrent stages (one device per process)
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream() {
  static SideStream x;
  if (!x.s) {
    if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &x;
}

static bool owns(int idx, int rank, int world) { return world <= 1 || idx % world == rank; }

extern "C" size_t tn_effh_plan_workspace_bytes(int a, int d, int b, int n_ls, int n_rs, int n_x) {
  size_t n = (size_t)a * d * b;
  return align_up(sizeof(ProblemDev) * (size_t)(1 + n_x)) + align_up(sizeof(LinkDev) * (size_t)(1 + n_ls + n_x)) +
         align_up(sizeof(ProblemDev) * (size_t)(1 + n_rs + n_x)) + align_up(sizeof(LinkDev) * (size_t)(1 + n_rs + n_x)) +
         align_up(sizeof(double) * n * (size_t)std::max(n_x, 0)) + align_up(sizeof(TmaMap) * (size_t)(3 * (2 + n_ls + n_rs + 2 * n_x))) +
         (d > kMaxD ? align_up(sizeof(double) * n * (size_t)(n_ls + n_rs) + 8) : 0) + 1024;
}

static void set_link(LinkDev& L, const double* A, const double* B, int a_dyn, int b_dyn, const double* op, int d) {
  std::memset(&L, 0, sizeof(L));
  L.A = A;
  L.B = B;
  L.a_dyn = a_dyn;
  L.b_dyn = b_dyn;
  L.has_op = op != nullptr;
  if (op)
    for (int i = 0; i < d * d; ++i) L.op[i] = op[i];
}

static int plan_create_impl(tn_effh_plan** out_plan, int a, int d, int b, const double* HL, const double* HR,
                            const double* M, int n_ls, const double* const* LS, const double* ls_op, int n_rs,
                            const double* const* RS, const double* rs_op, int n_x, const double* const* XL,
                            const double* const* XR, const double* x_coeff, int rank, int world, int row_begin, int row_count,
                            void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool rows = row_count >= 0;
  if (rows) TN_REQUIRE(row_begin >= 0 && row_count > 0 && row_begin + row_count <= a, "tn_effh_plan_create_rows: bad row slice [%d, %d) of %d", row_begin, row_begin + row_count, a);
  const int a_out = rows ? row_count : a;
  const long long roff = rows ? (long long)row_begin * a : 0;  // element offset of the row slice inside an (a, a) left block
  TN_REQUIRE(out_plan, "tn_effh_plan_create: null plan pointer");
  TN_REQUIRE(a > 0 && b > 0 && d >= 1 && d <= kMaxPhys, "tn_effh_plan_create: bad shape (%d,%d,%d)", a, d, b);
  const bool preop = d > kMaxD;
  const int dl = preop ? 1 : d;  // physical dimension the GEMM launches see
  TN_REQUIRE(n_ls >= 0 && n_rs >= 0 && n_x >= 0, "tn_effh_plan_create: negative counts");
  TN_REQUIRE(n_ls == 0 || (LS && ls_op), "tn_effh_plan_create: LS/ls_op missing");
  TN_REQUIRE(n_rs == 0 || (RS && rs_op), "tn_effh_plan_create: RS/rs_op missing");
  TN_REQUIRE(n_x == 0 || (XL && XR && x_coeff), "tn_effh_plan_create: crossing arrays missing");
  TN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "tn_effh_plan_create: bad rank/world %d/%d", rank, world);
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_effh_plan_create: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_effh_plan_workspace_bytes(a, d, b, n_ls, n_rs, n_x)) {
    set_error("tn_effh_plan_create: workspace %zu < %zu bytes", workspace_bytes, tn_effh_plan_workspace_bytes(a, d, b, n_ls, n_rs, n_x));
    return TN_ERR_WORKSPACE;
  }
  tn_effh_plan* P = new tn_effh_plan();
  P->a = a; P->d = d; P->b = b; P->n = (long long)a * d * b;
  P->rank = rank; P->world = world;
  P->rows = rows; P->row_begin = rows ? row_begin : 0; P->a_out = a_out;
  P->off = rows ? (long long)row_begin * d * b : 0; P->n_out = (long long)a_out * d * b;
  P->has_M = M != nullptr;
  std::memset(&P->M, 0, sizeof(P->M));
  if (M) for (int i = 0; i < d * d; ++i) P->M.m[i] = M[i];

  Carver cw(workspace, workspace_bytes);
  P->probA = cw.take<ProblemDev>(1 + n_x);
  P->linkA = cw.take<LinkDev>(1 + n_ls + n_x);
  P->probB = cw.take<ProblemDev>(1 + n_rs + n_x);
  P->linkB = cw.take<LinkDev>(1 + n_rs + n_x);
  P->phi = cw.take<double>((size_t)P->n_out * std::max(n_x, 0) + 1);
  P->maps = cw.take<TmaMap>((size_t)(3 * (2 + n_ls + n_rs + 2 * n_x)));
  P->tmaA = P->tmaB = false;
  P->preop = preop; P->n_pre_l = preop ? n_ls : 0; P->n_pre_r = preop ? n_rs : 0;
  P->pre_buf = preop ? cw.take<double>((size_t)P->n * (size_t)(n_ls + n_rs) + 1) : nullptr;
  if (preop) {
    P->pre_ops.resize((size_t)(n_ls + n_rs));
    for (int k = 0; k < n_ls + n_rs; ++k) {
      const double* src = k < n_ls ? ls_op + (size_t)k * d * d : rs_op + (size_t)(k - n_ls) * d * d;
      std::memset(&P->pre_ops[k], 0, sizeof(SiteOp));
      for (int i = 0; i < d * d; ++i) P->pre_ops[k].m[i] = src[i];
    }
  }
  if (!P->probA || !P->linkA || !P->probB || !P->linkB || !P->phi || !P->maps || (preop && !P->pre_buf)) {
    delete P;
    set_error("tn_effh_plan_create: workspace carve failed");
    return TN_ERR_WORKSPACE;
  }

  std::vector<ProblemDev> pa, pb;
  std::vector<LinkDev> la, lb;
  LinkDev L;
  // ---- left stage: chain into `out`, then one problem per owned crossing term ----
  int idx = 0;
  int chain_links = 0;
  if (HL && owns(idx, rank, world)) { set_link(L, HL + roff, nullptr, 0, 1, nullptr, d); la.push_back(L); ++chain_links; }
  ++idx;
  for (int k = 0; k < n_ls; ++k, ++idx) {
    TN_REQUIRE(LS[k], "tn_effh_plan_create: LS[%d] is null", k);
    if (owns(idx, rank, world)) {
      if (preop) set_link(L, LS[k] + roff, P->pre_buf + (size_t)k * P->n, 0, 0, nullptr, d);  // LS_k . (op_k psi), op_k psi precomputed
      else set_link(L, LS[k] + roff, nullptr, 0, 1, ls_op + (size_t)k * d * d, d);
      la.push_back(L);
      ++chain_links;
    }
  }
  if (chain_links) {
    ProblemDev q{};
    q.C = nullptr; q.alpha = 1.0; q.link_begin = 0; q.link_count = chain_links; q.accumulate = 1; q.c_dyn = 1;
    pa.push_back(q);
  }
  // right stage links on psi
  if (HR && owns(idx, rank, world)) { set_link(L, nullptr, HR, 1, 0, nullptr, d); lb.push_back(L); }
  ++idx;
  for (int k = 0; k < n_rs; ++k, ++idx) {
    TN_REQUIRE(RS[k], "tn_effh_plan_create: RS[%d] is null", k);
    if (owns(idx, rank, world)) {
      if (preop) set_link(L, P->pre_buf + (size_t)(n_ls + k) * P->n + P->off, RS[k], 0, 0, nullptr, d);
      else set_link(L, nullptr, RS[k], 1, 0, rs_op + (size_t)k * d * d, d);
      lb.push_back(L);
    }
  }
  int n_phi = 0;
  for (int i = 0; i < n_x; ++i, ++idx) {
    TN_REQUIRE(XL[i] && XR[i], "tn_effh_plan_create: crossing term %d has a null matrix", i);
    if (!owns(idx, rank, world)) continue;
    double* phi_i = P->phi + (size_t)n_phi * P->n_out;
    set_link(L, XL[i] + roff, nullptr, 0, 1, nullptr, d);
    ProblemDev q{};
    q.C = phi_i; q.alpha = x_coeff[i]; q.link_begin = (int)la.size(); q.link_count = 1; q.accumulate = 0; q.c_dyn = 0;
    la.push_back(L);
    pa.push_back(q);
    set_link(L, phi_i, XR[i], 0, 0, nullptr, d);
    lb.push_back(L);
    ++n_phi;
  }
  P->n_phi = n_phi;
  P->haveA = !pa.empty();
  P->haveB = !lb.empty();
  // independent stages (no crossing term feeds the right stage) of a small site: both add into `out` with atomics and run
  // on two streams (chi = 256 chain: 2 x 22 us of latency-bound launches per matvec overlap)
  P->overlapAB = P->haveA && P->haveB && n_phi == 0 && P->n_out <= (1LL << 19) && !deterministic_mode() && !getenv("TNALG_NO_OVERLAP");
  if (P->overlapAB)
    for (auto& q : pa) q.shared_out = 1;
  // concurrent stages in the small-tile configuration: each is cut for half of the CTA slots, which halves the number of
  // partial tiles combined with FP64 atomics and of pipeline fills compared with two full-size grids queuing behind each other
  static const int overlap_div = getenv("TNALG_OVERLAP_GRID_DIV") ? atoi(getenv("TNALG_OVERLAP_GRID_DIV")) : 2;
  const int grid_div = P->overlapAB ? std::max(1, overlap_div) : 1;
  const double* fake_psi = reinterpret_cast<const double*>(uintptr_t(256));  // alignment stand-in for scheduling
  const double* fake_slice = fake_psi + (P->off & 1);                         // the right stage reads psi + off
  const bool use_tma = tma_available() && !getenv("TNALG_NO_TMA");
  std::vector<TmaMap> hmaps;
  if (P->haveA) {
    P->LA = GemmLaunch{TN_NN, a_out, d * b, a, dl, a, d * b, d * b, (int)pa.size(), (int)la.size(), deterministic_mode() ? 1 : 0, grid_div};
    int st = gemm_plan_schedule(P->LA, pa.data(), la.data(), fake_psi, fake_psi, &P->SA);
    if (st != TN_OK) { delete P; return st; }
    if (use_tma && P->SA.config == 0 && P->SA.aligned16) {  // left stage: A = environment matrix (a x a), B = psi (per call)
      P->tmaA = true;
      for (auto& l : la) {
        l.a_map = (int)hmaps.size();
        hmaps.emplace_back();
        if (tma_encode_2d(&hmaps.back(), l.A, a_out, a, a, tma_box_rows_a()) != TN_OK) { P->tmaA = false; break; }
        if (!l.b_dyn) {  // pre-applied operator: the B operand is a fixed buffer of the plan
          l.b_map = (int)hmaps.size();
          hmaps.emplace_back();
          if (tma_encode_3d(&hmaps.back(), l.B, a, dl, (long long)(d / dl) * b, (long long)d * b) != TN_OK) { P->tmaA = false; break; }
        }
      }
    }
    TN_CUDA(cudaMemcpyAsync(P->probA, pa.data(), sizeof(ProblemDev) * pa.size(), cudaMemcpyHostToDevice, stream));
    TN_CUDA(cudaMemcpyAsync(P->linkA, la.data(), sizeof(LinkDev) * la.size(), cudaMemcpyHostToDevice, stream));
  }
  if (P->haveB) {
    // The right stage is one long chain (K_R + n_x links) into `out`.  It is cut into sub-problems of kChunk links that
    // all add into `out` with FP64 atomics: tiles are then visited chunk by chunk, every CTA of the grid works on the same
    // few links at the same time and their operands (Phi_i, XR_i) are shared through L2 instead of being re-read from HBM
    // by every tile (17.9 GB -> ~1.5 GB of DRAM reads per launch at chi = 1024, profiles/r01_ncu_matvec.md).
    // The chunk size is chosen so that (chunks x tiles) fills whole waves of the grid: the whole-tile round-robin schedule
    // then keeps all CTAs inside one chunk; a stream-K split would spread them over all chunks again.
    const int nl = (int)lb.size();
    int kChunk = 3;
    if (deterministic_mode()) {
      kChunk = nl;  // one problem, one CTA per output tile, fixed summation order (no FP64 atomics)
    } else {
      const int BMt = kTileBM, BNt = kTileBN;  // tile of the large configuration
      const int BMe = dl > 1 ? (BMt / dl) * dl : BMt;
      const long long tiles = (long long)((a_out * d + BMe - 1) / BMe) * ((b + BNt - 1) / BNt);
      const long long G = (long long)kTileCtas * sm_count();
      double best = -1.0;
      for (int ch = 4; ch >= 1; --ch) {
        const long long total = tiles * ((nl + ch - 1) / ch);
        const double eff = (double)total / (double)(((total + G - 1) / G) * G);
        if (eff >= 0.93) { kChunk = ch; break; }
        if (eff > best) { best = eff; kChunk = ch; }
      }
    }
    for (int l0 = 0; l0 < nl; l0 += kChunk) {
      ProblemDev q{};
      q.C = nullptr; q.alpha = 1.0; q.link_begin = l0; q.link_count = std::min(kChunk, nl - l0); q.accumulate = 1; q.c_dyn = 1;
      q.shared_out = (nl > kChunk || P->overlapAB) ? 1 : 0;
      pb.push_back(q);
    }
    P->LB = GemmLaunch{TN_NT, a_out * d, b, b, dl, b, b, b, (int)pb.size(), nl, deterministic_mode() ? 1 : 0, grid_div};
    int st = gemm_plan_schedule(P->LB, pb.data(), lb.data(), fake_slice, fake_psi, &P->SB);
    if (st != TN_OK) { delete P; return st; }
    if (use_tma && P->SB.config == 0 && P->SB.aligned16) {  // right stage: A = psi (per call) or Phi_i, B = environment matrix (b x b)
      P->tmaB = true;
      for (auto& l : lb) {
        if (!l.a_dyn) {
          l.a_map = (int)hmaps.size();
          hmaps.emplace_back();
          if (tma_encode_2d(&hmaps.back(), l.A, (long long)a_out * d, b, b, tma_box_rows_a()) != TN_OK) { P->tmaB = false; break; }
        }
        l.b_map = (int)hmaps.size();
        hmaps.emplace_back();
        if (tma_encode_2d(&hmaps.back(), l.B, b, b, b, tma_box_rows_b_nt()) != TN_OK) { P->tmaB = false; break; }
      }
    }
    TN_CUDA(cudaMemcpyAsync(P->probB, pb.data(), sizeof(ProblemDev) * pb.size(), cudaMemcpyHostToDevice, stream));
    TN_CUDA(cudaMemcpyAsync(P->linkB, lb.data(), sizeof(LinkDev) * lb.size(), cudaMemcpyHostToDevice, stream));
  }
  if (!hmaps.empty() && (P->tmaA || P->tmaB))
    TN_CUDA(cudaMemcpyAsync(P->maps, hmaps.data(), sizeof(TmaMap) * hmaps.size(), cudaMemcpyHostToDevice, stream));
  const double KL = (HL ? 1 : 0) + n_ls, KR = (HR ? 1 : 0) + n_rs;
  P->alg_flops = 2.0 * a * d * b * ((double)a * (KL + n_x) + (double)b * (KR + n_x));
  P->exec_flops = 2.0 * a_out * d * b * ((double)a * (double)la.size() + (double)b * (double)lb.size());
  *out_plan = P;
  return TN_OK;
}

extern "C" int tn_effh_plan_create(tn_effh_plan** out_plan, int a, int d, int b, const double* HL, const double* HR,
                                   const double* M, int n_ls, const double* const* LS, const double* ls_op, int n_rs,
                                   const double* const* RS, const double* rs_op, int n_x, const double* const* XL,
                                   const double* const* XR, const double* x_coeff, int rank, int world, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  return plan_create_impl(out_plan, a, d, b, HL, HR, M, n_ls, LS, ls_op, n_rs, RS, rs_op, n_x, XL, XR, x_coeff, rank, world, 0, -1,
                          workspace, workspace_bytes, stream_);
}

extern "C" int tn_effh_plan_create_rows(tn_effh_plan** out_plan, int a, int d, int b, const double* HL, const double* HR,
                                        const double* M, int n_ls, const double* const* LS, const double* ls_op, int n_rs,
                                        const double* const* RS, const double* rs_op, int n_x, const double* const* XL,
                                        const double* const* XR, const double* x_coeff, int row_begin, int row_count,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
  TN_REQUIRE(row_count > 0, "tn_effh_plan_create_rows: empty row slice");
  return plan_create_impl(out_plan, a, d, b, HL, HR, M, n_ls, LS, ls_op, n_rs, RS, rs_op, n_x, XL, XR, x_coeff, 0, 1, row_begin,
                          row_count, workspace, workspace_bytes, stream_);
}

extern "C" int tn_effh_plan_rows(const tn_effh_plan* P, int* row_begin, int* row_count) {
  TN_REQUIRE(P, "tn_effh_plan_rows: null plan");
  if (row_begin) *row_begin = P->row_begin;
  if (row_count) *row_count = P->a_out;
  return P->rows ? 1 : 0;
}

extern "C" int tn_effh_matvec(tn_effh_plan* P, const double* psi_in, double* psi_out, double c_id, double c_h, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(P && psi_in && psi_out && psi_in != psi_out, "tn_effh_matvec: bad arguments");
  const bool need16 = (P->haveA && P->SA.aligned16) || (P->haveB && P->SB.aligned16);
  if (need16)
    TN_REQUIRE(((reinterpret_cast<uintptr_t>(psi_in) | reinterpret_cast<uintptr_t>(psi_out)) & 15) == 0,
               "tn_effh_matvec: psi_in/psi_out must be 16-byte aligned for this plan");
  // identity + on-site part (rank 0 only when sharded); doubles as the zero-initialisation of `out`
  // a row-sliced plan reads the full psi_in and writes rows [row_begin, row_begin + a_out) into psi_out (the slice buffer)
  const bool lead = P->rank == 0;
  const double* psi_slice = psi_in + P->off;
  TN_CHECK(launch_site_op_axpby(psi_out, psi_slice, P->a_out, P->d, P->b, lead ? c_id : 0.0, (lead && P->has_M) ? c_h : 0.0, P->M, stream));
  TmaMap psi_a, psi_b;  // psi as the A operand of the right stage / the (k, s, y) B operand of the left stage
  const int dl = P->preop ? 1 : P->d;
  if (P->tmaA) TN_CHECK(tma_encode_3d(&psi_b, psi_in, P->a, dl, (long long)(P->d / dl) * P->b, (long long)P->d * P->b));
  if (P->preop) {  // op_k psi for the '1_s_0' links (all rows: B operand of the left stage) and the '0_s_1' links (this slice)
    for (int k = 0; k < P->n_pre_l; ++k)
      TN_CHECK(launch_site_op_axpby(P->pre_buf + (size_t)k * P->n, psi_in, P->a, P->d, P->b, 0.0, 1.0, P->pre_ops[k], stream));
    for (int k = 0; k < P->n_pre_r; ++k)
      TN_CHECK(launch_site_op_axpby(P->pre_buf + (size_t)(P->n_pre_l + k) * P->n + P->off, psi_slice, P->a_out, P->d, P->b, 0.0, 1.0,
                                    P->pre_ops[P->n_pre_l + k], stream));
  }
  if (P->tmaB) TN_CHECK(tma_encode_2d(&psi_a, psi_slice, (long long)P->a_out * P->d, P->b, P->b, tma_box_rows_a()));
  SideStream* side = P->overlapAB ? side_stream() : nullptr;
  cudaStream_t streamB = stream;
  if (side) {  // fork: the right stage runs on the side stream after the init kernel
    TN_CUDA(cudaEventRecord(side->fork, stream));
    TN_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    streamB = side->s;
  }
  if (side && P->haveB) {
    if (P->tmaB)
      TN_CHECK(gemm_launch_tma(P->LB, P->SB, P->probB, P->linkB, P->maps, psi_a, psi_b, psi_slice, psi_out, c_h, streamB));
    else
      TN_CHECK(gemm_launch(P->LB, P->SB, P->probB, P->linkB, psi_slice, psi_out, c_h, streamB));
    TN_CUDA(cudaEventRecord(side->join, streamB));
  }
  if (P->haveA) {
    if (P->SA.split && P->n_phi > 0) TN_CUDA(cudaMemsetAsync(P->phi, 0, sizeof(double) * (size_t)P->n_out * P->n_phi, stream));
    if (P->tmaA)
      TN_CHECK(gemm_launch_tma(P->LA, P->SA, P->probA, P->linkA, P->maps, psi_a, psi_b, psi_in, psi_out, c_h, stream));
    else
      TN_CHECK(gemm_launch(P->LA, P->SA, P->probA, P->linkA, psi_in, psi_out, c_h, stream));
  }
  if (side) {
    TN_CUDA(cudaStreamWaitEvent(stream, side->join, 0));  // join
  } else if (P->haveB) {
    if (P->tmaB)
      TN_CHECK(gemm_launch_tma(P->LB, P->SB, P->probB, P->linkB, P->maps, psi_a, psi_b, psi_slice, psi_out, c_h, stream));
    else
      TN_CHECK(gemm_launch(P->LB, P->SB, P->probB, P->linkB, psi_slice, psi_out, c_h, stream));
  }
  return TN_OK;
}

extern "C" int tn_effh_plan_flops(const tn_effh_plan* P, double* algorithmic, double* executed) {
  TN_REQUIRE(P, "tn_effh_plan_flops: null plan");
  if (algorithmic) *algorithmic = P->alg_flops;
  if (executed) *executed = P->exec_flops;
  return TN_OK;
}

extern "C" int tn_effh_plan_uses_tma(const tn_effh_plan* P) { return P ? ((P->haveA && P->tmaA) ? 1 : 0) | ((P->haveB && P->tmaB) ? 2 : 0) : 0; }

extern "C" int tn_effh_plan_destroy(tn_effh_plan* P) {
  delete P;
  return TN_OK;
}

// internal accessors for the Lanczos driver
namespace tn {
long long plan_dim(const tn_effh_plan* P) { return P->n; }
bool plan_is_rows(const tn_effh_plan* P) { return P->rows; }
long long plan_slice_offset(const tn_effh_plan* P) { return P->off; }
long long plan_slice_len(const tn_effh_plan* P) { return P->n_out; }
int plan_world(const tn_effh_plan* P) { return P->world; }
}

// ------------------------------------------------------------------------------------------------
// environment update
// ------------------------------------------------------------------------------------------------
extern "C" size_t tn_env_update_workspace_bytes(int a, int d, int b, int n_out, int n_links) {
  size_t n = (size_t)a * d * b;
  return align_up(sizeof(double) * n * (size_t)std::max(n_out, 1)) + 2 * align_up(sizeof(ProblemDev) * (size_t)std::max(n_out, 1)) +
         align_up(sizeof(LinkDev) * (size_t)std::max(n_links, 1)) + align_up(sizeof(LinkDev) * (size_t)std::max(n_out, 1)) + 1024;
}

extern "C" int tn_env_update(int dir, const double* T, int a, int d, int b, int n_out, double* const* E_out,
                             const int* out_link_begin, const double* const* link_E, const double* link_op,
                             const int* link_has_op, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(dir == 0 || dir == 1, "tn_env_update: dir must be 0 (l2r) or 1 (r2l)");
  TN_REQUIRE(T && a > 0 && b > 0 && d >= 1 && d <= kMaxD, "tn_env_update: bad tensor (%d,%d,%d)", a, d, b);
  TN_REQUIRE(n_out > 0 && E_out && out_link_begin && link_E && link_op && link_has_op, "tn_env_update: null arrays");
  const int n_links = out_link_begin[n_out];
  TN_REQUIRE(out_link_begin[0] == 0 && n_links > 0, "tn_env_update: bad link offsets");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_env_update: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_env_update_workspace_bytes(a, d, b, n_out, n_links)) {
    set_error("tn_env_update: workspace %zu < %zu bytes", workspace_bytes, tn_env_update_workspace_bytes(a, d, b, n_out, n_links));
    return TN_ERR_WORKSPACE;
  }
  const size_t n = (size_t)a * d * b;
  const int e_in = dir == 0 ? a : b;    // dimension of the incoming environments
  const int e_out = dir == 0 ? b : a;   // dimension of the outgoing environments
  Carver cw(workspace, workspace_bytes);
  double* X = cw.take<double>(n * n_out);
  ProblemDev* p1d = cw.take<ProblemDev>(n_out);
  ProblemDev* p2d = cw.take<ProblemDev>(n_out);
  LinkDev* l1d = cw.take<LinkDev>(n_links);
  LinkDev* l2d = cw.take<LinkDev>(n_out);
  TN_REQUIRE(X && p1d && p2d && l1d && l2d, "tn_env_update: workspace carve failed");

  std::vector<ProblemDev> p1, p2;
  std::vector<LinkDev> l1, l2;
  for (int j = 0; j < n_out; ++j) {
    TN_REQUIRE(E_out[j], "tn_env_update: E_out[%d] is null", j);
    TN_REQUIRE(out_link_begin[j + 1] > out_link_begin[j], "tn_env_update: output %d has no links", j);
    double* Xj = X + n * j;
    // links whose incoming environment is the identity reduce to a site operator acting on T
    SiteOp idsum{};
    bool any_id = false;
    int first_gemm = (int)l1.size();
    for (int l = out_link_begin[j]; l < out_link_begin[j + 1]; ++l) {
      const double* op = link_op + (size_t)l * d * d;
      if (link_E[l] == nullptr) {
        any_id = true;
        for (int s = 0; s < d; ++s)
          for (int sp = 0; sp < d; ++sp) idsum.m[s * d + sp] += link_has_op[l] ? op[s * d + sp] : (s == sp ? 1.0 : 0.0);
      } else {
        LinkDev L;
        if (dir == 0) set_link(L, link_E[l], T, 0, 0, link_has_op[l] ? op : nullptr, d);   // X += E . (op T)
        else          set_link(L, T, link_E[l], 0, 0, link_has_op[l] ? op : nullptr, d);   // Y += (op T) . E^T
        l1.push_back(L);
      }
    }
    if (any_id) TN_CHECK(launch_site_op_axpby(Xj, T, a, d, b, 0.0, 1.0, idsum, stream));
    const int n_gemm = (int)l1.size() - first_gemm;
    if (n_gemm > 0) {
      ProblemDev q{};
      q.C = Xj; q.alpha = 1.0; q.link_begin = first_gemm; q.link_count = n_gemm; q.accumulate = any_id ? 1 : 0; q.c_dyn = 0;
      p1.push_back(q);
    }
    ProblemDev q2{};
    q2.C = E_out[j]; q2.alpha = 1.0; q2.link_begin = (int)l2.size(); q2.link_count = 1; q2.accumulate = 0; q2.c_dyn = 0;
    p2.push_back(q2);
    LinkDev L2;
    if (dir == 0) set_link(L2, T, Xj, 0, 0, nullptr, 1);   // E' = T^T X      (TN)
    else          set_link(L2, T, Xj, 0, 0, nullptr, 1);   // E' = T . Y^T    (NT, d = 1)
    l2.push_back(L2);
  }
  (void)e_in;
  if (!p1.empty()) {
    // heavy chains first so that the whole-tile schedule balances
    GemmLaunch L1 = dir == 0 ? GemmLaunch{TN_NN, a, d * b, a, d, a, d * b, d * b, (int)p1.size(), (int)l1.size(), 1}
                             : GemmLaunch{TN_NT, a * d, b, b, d, b, b, b, (int)p1.size(), (int)l1.size(), 1};
    std::stable_sort(p1.begin(), p1.end(), [](const ProblemDev& x, const ProblemDev& y) { return x.link_count > y.link_count; });
    GemmSchedule S1;
    TN_CHECK(gemm_plan_schedule(L1, p1.data(), l1.data(), nullptr, nullptr, &S1));
    TN_CUDA(cudaMemcpyAsync(p1d, p1.data(), sizeof(ProblemDev) * p1.size(), cudaMemcpyHostToDevice, stream));
    TN_CUDA(cudaMemcpyAsync(l1d, l1.data(), sizeof(LinkDev) * l1.size(), cudaMemcpyHostToDevice, stream));
    TN_CHECK(gemm_launch(L1, S1, p1d, l1d, nullptr, nullptr, 1.0, stream));
  }
  GemmLaunch L2 = dir == 0 ? GemmLaunch{TN_TN, e_out, e_out, a * d, 1, b, b, e_out, (int)p2.size(), (int)l2.size(), 1}
                           : GemmLaunch{TN_NT, e_out, e_out, d * b, 1, d * b, d * b, e_out, (int)p2.size(), (int)l2.size(), 1};
  GemmSchedule S2;
  TN_CHECK(gemm_plan_schedule(L2, p2.data(), l2.data(), nullptr, nullptr, &S2));
  TN_CUDA(cudaMemcpyAsync(p2d, p2.data(), sizeof(ProblemDev) * p2.size(), cudaMemcpyHostToDevice, stream));
  TN_CUDA(cudaMemcpyAsync(l2d, l2.data(), sizeof(LinkDev) * l2.size(), cudaMemcpyHostToDevice, stream));
  return gemm_launch(L2, S2, p2d, l2d, nullptr, nullptr, 1.0, stream);
}
