// One- and two-body expectation values on a centre-orthogonal MPS (a10: observation_s1 / observation_s1_s2,
// MPSClass.py:857-909), the C-ABI form SURVEY.md 8b names: a batch of terms in, host doubles out.
// Every term is one left-to-right chain of environment transfers (tn_env_update, a5) from min(first site, centre) to
// max(last site, centre) -- the sites left of that range are left-orthonormal, the sites right of it right-orthonormal, so
// both ends close with the identity -- followed by a trace.  (The Python layer's expect_products shares chain prefixes
// between the terms of a batch; this entry point is the plain per-term form for callers that bind the library directly.)
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace tn;

namespace {
struct Term {
  int n;              // operators in the term (1 or 2)
  int site[2];
  const double* op[2];  // [host] d*d each
};

size_t expect_ws_bytes(const int* dims, int L, int d, int n_terms) {
  size_t chi2 = 1, wmax = 0;
  for (int s = 0; s < L; ++s) {
    chi2 = std::max<size_t>(chi2, (size_t)dims[s + 1] * dims[s + 1]);
    wmax = std::max(wmax, tn_env_update_workspace_bytes(dims[s], d, dims[s + 1], 1, 1));
  }
  return 2 * align_up(sizeof(double) * chi2) + align_up(wmax) + align_up(sizeof(double) * (size_t)std::max(n_terms, 1)) + 1024;
}

int run_terms(const double* const* mps, const int* dims, int L, int d, int center, const std::vector<Term>& terms, double* out,
              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  TN_REQUIRE(mps && dims && out && L >= 1 && d >= 1 && d <= kMaxD && center >= 0 && center < L, "tn_expect: bad arguments");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_expect: workspace must be 256-byte aligned");
  const int n_terms = (int)terms.size();
  if (workspace_bytes < expect_ws_bytes(dims, L, d, n_terms)) {
    set_error("tn_expect: workspace %zu < %zu bytes", workspace_bytes, expect_ws_bytes(dims, L, d, n_terms));
    return TN_ERR_WORKSPACE;
  }
  size_t chi2 = 1, wmax = 0;
  for (int s = 0; s < L; ++s) {
    TN_REQUIRE(mps[s] && dims[s] > 0 && dims[s + 1] > 0, "tn_expect: bad tensor %d", s);
    chi2 = std::max<size_t>(chi2, (size_t)dims[s + 1] * dims[s + 1]);
    wmax = std::max(wmax, tn_env_update_workspace_bytes(dims[s], d, dims[s + 1], 1, 1));
  }
  Carver cw(workspace, workspace_bytes);
  double* E[2] = {cw.take<double>(chi2), cw.take<double>(chi2)};
  char* wenv = cw.take<char>(wmax);
  double* results = cw.take<double>((size_t)std::max(n_terms, 1));
  TN_REQUIRE(E[0] && E[1] && wenv && results, "tn_expect: workspace carve failed");
  std::vector<double> ident((size_t)d * d, 0.0);
  for (int tI = 0; tI < n_terms; ++tI) {
    const Term& tm = terms[tI];
    for (int k = 0; k < tm.n; ++k) TN_REQUIRE(tm.site[k] >= 0 && tm.site[k] < L && tm.op[k], "tn_expect: term %d has a bad site / operator", tI);
    TN_REQUIRE(tm.n == 1 || tm.site[0] < tm.site[1], "tn_expect: term %d needs site1 < site2", tI);
    const int first = std::min(tm.site[0], center), last = std::max(tm.site[tm.n - 1], center);
    int cur = 0;
    bool have = false;  // false: the incoming environment is the identity
    for (int s = first; s <= last; ++s) {
      const double* op = nullptr;
      for (int k = 0; k < tm.n; ++k)
        if (tm.site[k] == s) op = tm.op[k];
      double* outp[1] = {E[cur ^ 1]};
      const int begin[2] = {0, 1};
      const double* linkE[1] = {have ? E[cur] : nullptr};
      const int has_op[1] = {op ? 1 : 0};
      TN_CHECK(tn_env_update(0, mps[s], dims[s], d, dims[s + 1], 1, outp, begin, linkE, op ? op : ident.data(), has_op, wenv, wmax, stream));
      cur ^= 1;
      have = true;
    }
    TN_CHECK(tn_trace(E[cur], dims[last + 1], results + tI, stream));
  }
  if (n_terms > 0) TN_CUDA(cudaMemcpyAsync(out, results, sizeof(double) * n_terms, cudaMemcpyDeviceToHost, stream));
  TN_CUDA(cudaStreamSynchronize(stream));
  return TN_OK;
}
}  // namespace

extern "C" size_t tn_expect_workspace_bytes(const int* dims, int L, int d, int n_terms) {
  if (!dims || L < 1) return 0;
  return expect_ws_bytes(dims, L, d, n_terms);
}

extern "C" int tn_expect_1body(const double* const* mps, const int* dims, int L, int d, int center, int n_terms, const int* sites,
                               const double* ops, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  TN_REQUIRE(n_terms >= 0 && (n_terms == 0 || (sites && ops)), "tn_expect_1body: null arrays");
  std::vector<Term> terms((size_t)n_terms);
  for (int i = 0; i < n_terms; ++i) terms[i] = Term{1, {sites[i], 0}, {ops + (size_t)i * d * d, nullptr}};
  return run_terms(mps, dims, L, d, center, terms, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int tn_expect_2body(const double* const* mps, const int* dims, int L, int d, int center, int n_terms, const int* site1,
                               const int* site2, const double* ops1, const double* ops2, double* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  TN_REQUIRE(n_terms >= 0 && (n_terms == 0 || (site1 && site2 && ops1 && ops2)), "tn_expect_2body: null arrays");
  std::vector<Term> terms((size_t)n_terms);
  for (int i = 0; i < n_terms; ++i)
    terms[i] = Term{2, {site1[i], site2[i]}, {ops1 + (size_t)i * d * d, ops2 + (size_t)i * d * d}};
  return run_terms(mps, dims, L, d, center, terms, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
