/*
 * tnalg_b200 -- C ABI of the B200 (sm_100a) finite-size DMRG hot path of ranshiju/T-Nalg.
 *
 * The reference is pure Python and has no FFI: the seam this ABI replaces is the set of numpy/scipy call
 * sites inside MPSClass.MpsOpenBoundaryClass and TensorBasicModule.  Each entry point
 * cites the reference interface it stands in for (file:line relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, a negative tn_status on failure; tn_last_error() gives the text
 *     (thread-local).  Shape / alignment / null-pointer mistakes are errors, never undefined behaviour.
 *   - all `double*` arguments marked [dev] are device pointers to C-contiguous (row-major) float64;
 *     [host] arguments are ordinary host memory read before the call returns.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous with
 *     respect to the host unless stated otherwise.
 *   - no hidden device allocation: scratch is passed by the caller; *_workspace_bytes() says how much.
 *   - MPS tensors are (a, d, b) = (left bond, physical, right bond), flattened in C order -- the layout of
 *     mps[n] in MPSClass.py:66-72.  Environment matrices are indexed [bra, ket] and act as E . psi
 *     (MPSClass.py:757-775 passes E.T to absorb_matrix2tensor, which contracts the first index).
 */
#ifndef TNALG_B200_H
#define TNALG_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* physical dimension of a site (or of the combined two-site window): d = 2 (spin-1/2), d = 3 (spin-1,
 * Parameters.py:440-446), d = 4 (two spin-1/2 sites), d = 9 (two spin-1 sites).  Site operators of dimension up to
 * TN_MAX_LOADPATH_DIM are applied inside the GEMM operand path; larger ones (effective-Hamiltonian plans only) are applied by
 * one element-wise pass per operator before the GEMM. */
#define TN_MAX_PHYS_DIM 9
#define TN_MAX_LOADPATH_DIM 4

typedef enum {
  TN_OK = 0,
  TN_ERR_INVALID = -1,   /* bad argument (shape, null pointer, alignment) */
  TN_ERR_CUDA = -2,      /* CUDA runtime error, text in tn_last_error() */
  TN_ERR_WORKSPACE = -3, /* workspace too small */
  TN_ERR_NOCONV = -4,    /* iteration limit reached (result is still written) */
  TN_ERR_DEVICE = -5,    /* not an sm_100 device */
  TN_ERR_NUMERIC = -6    /* numerical failure (NaN/Inf reached an iteration); outputs are NOT written */
} tn_status;

const char* tn_last_error(void);
int tn_version(void);
/* sm_count, compute capability of the current device; fails with TN_ERR_DEVICE when it is not sm_100. */
int tn_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched since tn_launch_count_reset() (bench.py "gpu_launches"). */
long long tn_launch_count(void);
void tn_launch_count_reset(void);
/* on != 0: bit-reproducible mode -- no stream-K split, no FP64-atomic combination of partial tiles, no concurrent matvec
 * stages (every output tile is summed by one CTA in a fixed order).  Returns the previous setting.  Default off. */
int tn_set_deterministic(int on);
/* issue-rate ceiling of the FP64 tensor pipe (DMMA.8x8x4 from registers on every SM), TFLOP/s, measured now on the current
 * device; scratch: [dev] 8 * 512 * sm_count bytes.  Blocking (a few ms). */
int tn_measure_dmma_peak(double* tflops_out /* [host] */, void* scratch /* [dev] */, size_t scratch_bytes, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Chain GEMM: the FP64 tensor-core (DMMA) contraction engine every tensor-network contraction below is built
 * from.   C_p (+)= alpha_p * sum_{l in links(p)}  opA_l(A_l) . opB_l(B_l)      for p = 0 .. n_problems-1
 * All problems of one call share M, N, K and the leading dimensions.
 *   mode TN_NN : A_l is (M,K) row-major, B_l is (K,N) row-major with N = d*Ny and column index (s,y); a link's
 *                d x d operator acts on s:   B'[k,(s,y)] = sum_s' op[s,s'] B[k,(s',y)]
 *                (E . (op T) with T an (a,d,b) tensor: absorb_matrix2tensor bonds 0+1, TensorBasicModule.py:387-424)
 *   mode TN_NT : A_l is (M,K) row-major with M = Mx*d and row index (x,s), B_l is (N,K) row-major;
 *                A'[(x,s),k] = sum_s' op[s,s'] A[(x,s'),k]            ((op T) . E^T, bonds 1+2)
 *   mode TN_TN : A_l is (K,M) row-major, B_l is (K,N) row-major, no operator  (T^T . X, the closing GEMM of
 *                bound_vec_operator_left2right, TensorBasicModule.py:560-568)
 * A link with has_op == 0 uses the identity.  d = 1 disables the operator machinery.
 * deterministic != 0 forbids split-K (one CTA owns each output tile, fixed summation order).
 * Otherwise the K range of a tile may be split over CTAs (stream-K) and partial tiles are combined with FP64
 * atomics; the result then differs between runs at the 1e-16 relative level.
 * accumulate == 0: C is overwritten; != 0: the result is added to the existing C.
 * -------------------------------------------------------------------------------------------------- */
enum { TN_NN = 0, TN_NT = 1, TN_TN = 2 };

typedef struct {
  const double* A; /* [dev] */
  const double* B; /* [dev] */
  double op[TN_MAX_LOADPATH_DIM * TN_MAX_LOADPATH_DIM]; /* row-major d x d (d <= TN_MAX_LOADPATH_DIM), used when has_op != 0 */
  int has_op;
  int reserved;
} tn_link;

typedef struct {
  double* C; /* [dev] (M,N) row-major, leading dimension ldc */
  double alpha;
  int link_begin; /* index of the first link of this problem in the links array */
  int link_count;
  int accumulate;
  int reserved;
} tn_problem;

size_t tn_chain_gemm_workspace_bytes(int n_problems, int n_links);
int tn_chain_gemm(int mode, int M, int N, int K, int d, int lda, int ldb, int ldc,
                  const tn_problem* problems /* [host] */, int n_problems,
                  const tn_link* links /* [host] */, int n_links, int deterministic,
                  void* workspace /* [dev] */, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Environment update through one site tensor (a5: bound_vec_operator_left2right / right2left,
 * TensorBasicModule.py:530-619), batched over every operator that crosses the bond and fused with the sum over
 * the terms that close on this site (MPSClass.py:474-530 keeps one matrix per term instead):
 *   dir = 0 (left to right):  E_out[j][b,b'] = sum_{l in links(j)} sum conj(T[a,s,b]) E_l[a,a'] op_l[s,s'] T[a',s',b']
 *   dir = 1 (right to left):  E_out[j][a,a'] = sum_{l in links(j)} sum conj(T[a,s,b]) E_l[b,b'] op_l[s,s'] T[a',s',b']
 * link_E[l] == NULL means the identity (the far side of the bond is orthonormal); link_has_op[l] == 0 means the
 * identity on the physical index.  Coefficients are folded into link_op.  Deterministic.
 * -------------------------------------------------------------------------------------------------- */
size_t tn_env_update_workspace_bytes(int a, int d, int b, int n_out, int n_links);
int tn_env_update(int dir, const double* T /* [dev] (a,d,b) */, int a, int d, int b, int n_out,
                  double* const* E_out /* [host] array of n_out [dev] pointers */,
                  const int* out_link_begin /* [host] n_out+1 prefix offsets into the link arrays */,
                  const double* const* link_E /* [host] n_links [dev] pointers or NULL */,
                  const double* link_op /* [host] n_links * d*d */, const int* link_has_op /* [host] n_links */,
                  void* workspace /* [dev] */, size_t workspace_bytes, void* stream);

/* out[i] = sum_t coeffs[t] * xs[t][i]  (the sums classify_and_update_env keeps per key, MPSClass.py:684-733) */
int tn_lincomb(double* out /* [dev] */, long long n, int n_terms, const double* const* xs /* [host] of [dev] */,
               const double* coeffs /* [host] */, void* stream);

/* out = op (d x d, [host]) applied on the physical index of T (a,d,b): out[a,s,b] = sum_s' op[s,s'] T[a,s',b]
 * (absorb_matrix2tensor(T, op.T, 1), TensorBasicModule.py:415-421) */
int tn_apply_site_op(double* out /* [dev] */, const double* T /* [dev] */, int a, int d, int b,
                     const double* op /* [host] d*d */, void* stream);

/* result[0] = sum_i x[i]*y[i]; deterministic two-stage reduction; result is a [dev] double.
 * tn_trace: result[0] = sum_i E[i,i] of an (n,n) matrix (np.trace in MPSClass.py:872,909). */
int tn_dot(const double* x, const double* y, long long n, double* result /* [dev] */, void* workspace /* [dev] */,
           size_t workspace_bytes, void* stream);
size_t tn_dot_workspace_bytes(long long n);
int tn_trace(const double* E /* [dev] (n,n) */, int n, double* result /* [dev] */, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Effective-Hamiltonian plan for one site (a1/a2): the groups of MpsOpenBoundaryClass.opt_env
 * (MPSClass.py:684-733) with every coefficient already folded in:
 *   HL  '1_0_0' (a,a) or NULL          HR  '0_0_1' (b,b) or NULL          M  '0_s_0' summed (d,d) [host] or NULL
 *   LS[k], ls_op[k]  '1_s_0': (LS_k (x) op_k)      RS[k], rs_op[k]  '0_s_1': (op_k (x) RS_k)
 *   XL[i], XR[i], x_coeff[i]  '1_0_1': c_i (XL_i (x) 1 (x) XR_i)
 * H_eff psi = HL psi + psi HR^T + M psi + sum_k LS_k op_k psi + sum_k op_k psi RS_k^T + sum_i c_i XL_i psi XR_i^T.
 * tn_effh_matvec:  out = c_id * psi + c_h * H_eff psi ; the reference handle
 * update_tensor_eigs_f_handle_optimized (MPSClass.py:755-776) is (c_id, c_h) = (1, -tau).
 * Executed as two chain-GEMM launches (left stage TN_NN, right stage TN_NT) plus one element-wise init kernel;
 * the crossing intermediates XL_i psi live in the plan workspace.
 * rank/world shard the links round-robin (multi-GPU): the caller all-reduces `out`; the identity
 * and on-site parts are applied on rank 0 only.
 * tn_effh_plan_create_rows builds the other multi-GPU decomposition: the plan computes rows [row_begin, row_begin +
 * row_count) of the (a, d*b) output -- every term, balanced for any number of terms -- from the FULL input psi;
 * tn_effh_matvec then reads psi_in (a,d,b) and writes the (row_count,d,b) slice to psi_out.  No reduction is needed: the
 * exchange step becomes an all-gather of the slices of the next Krylov vector (tn_lanczos_lm1 with a tn_comm).
 * -------------------------------------------------------------------------------------------------- */
typedef struct tn_effh_plan tn_effh_plan;

size_t tn_effh_plan_workspace_bytes(int a, int d, int b, int n_ls, int n_rs, int n_x);
int tn_effh_plan_create(tn_effh_plan** plan, int a, int d, int b, const double* HL /* [dev] */,
                        const double* HR /* [dev] */, const double* M /* [host] d*d or NULL */, int n_ls,
                        const double* const* LS /* [host] of [dev] */, const double* ls_op /* [host] n_ls*d*d */,
                        int n_rs, const double* const* RS, const double* rs_op, int n_x,
                        const double* const* XL, const double* const* XR, const double* x_coeff /* [host] */,
                        int rank, int world, void* workspace /* [dev] */, size_t workspace_bytes, void* stream);
int tn_effh_plan_create_rows(tn_effh_plan** plan, int a, int d, int b, const double* HL, const double* HR, const double* M,
                             int n_ls, const double* const* LS, const double* ls_op, int n_rs, const double* const* RS,
                             const double* rs_op, int n_x, const double* const* XL, const double* const* XR,
                             const double* x_coeff, int row_begin, int row_count, void* workspace /* [dev] */,
                             size_t workspace_bytes, void* stream);
/* returns 1 for a row-sliced plan (and its slice), 0 for a full / term-sharded plan, negative on error */
int tn_effh_plan_rows(const tn_effh_plan* plan, int* row_begin, int* row_count);
int tn_effh_matvec(tn_effh_plan* plan, const double* psi_in /* [dev] */, double* psi_out /* [dev] */, double c_id,
                   double c_h, void* stream);
/* algorithmic flop of one matvec: 2*a*d*b*[a*(K_L+n_x) + b*(K_R+n_x)], and executed flop. */
int tn_effh_plan_flops(const tn_effh_plan* plan, double* algorithmic, double* executed);
/* bit 0: the left stage runs the TMA-staged kernel, bit 1: the right stage does (else the cp.async kernel). */
int tn_effh_plan_uses_tma(const tn_effh_plan* plan);
int tn_effh_plan_destroy(tn_effh_plan* plan);

/* --------------------------------------------------------------------------------------------------
 * Device-resident Lanczos (a8): the eigenpair of (1 - tau*H_eff) of largest magnitude, i.e. what
 * eigsh(LinearOperator, k=1, which='LM', v0, tol) returns at MPSClass.py:801-805.  Thick-restart Lanczos on
 * H_eff itself (same Krylov space, no 1e-4 shift cancellation) with full CGS2 re-orthogonalisation; all scalars,
 * the projected tridiagonal eigenproblem and the convergence test stay on the device; the host reads one flag
 * per restart cycle.  Converged when  tau*|beta_m u_m| <= tol * |1 - tau*theta|  (ARPACK's criterion applied to
 * the shifted operator).  ncv = Krylov dimension per cycle (ARPACK default 20), max_restarts cycles at most.
 * Multi-GPU (one exchange per Lanczos step, stream-ordered, issued by the library):
 *   - term-sharded plan + comm: ncclAllReduce of the partial H psi (or the caller's `allreduce` callback, comm == NULL);
 *   - row-sliced plan + comm:   every rank keeps only its slice of each Krylov vector; ncclAllGather of the slices of v_j
 *     before the matvec, dot products as two <= (ncv+2)-number all-reduces per step; v0 and vec_out are full vectors,
 *     bit-identical on all ranks.
 * Blocking: returns after the result is available.  lambda/resid/n_matvec are [host] outputs.
 * TN_ERR_NOCONV: iteration limit, the best Ritz pair IS written.  TN_ERR_NUMERIC: nothing is written.
 * -------------------------------------------------------------------------------------------------- */
typedef int (*tn_allreduce_fn)(double* buf /* [dev] */, long long count, void* user, void* stream);
typedef struct tn_comm tn_comm;

size_t tn_lanczos_workspace_bytes(long long n, int ncv);
size_t tn_lanczos_workspace_bytes_sharded(const tn_effh_plan* plan, const tn_comm* comm, int ncv);
int tn_lanczos_lm1(tn_effh_plan* plan, double tau, const double* v0 /* [dev] n */, double tol, int ncv,
                   int max_restarts, double* lambda_out, double* vec_out /* [dev] n */, int* n_matvec_out,
                   double* resid_out, tn_allreduce_fn allreduce, void* allreduce_user, tn_comm* comm,
                   void* workspace /* [dev] */, size_t workspace_bytes, void* stream);

/* The same solver for a caller-supplied operator y = H x on device vectors (the `lin_map` of eigs_fh,
 * Eigs_Module_sjr.py:18-58; also the full-space Hamiltonian of the exact-diagonalisation cross-check,
 * library/EDspinClass.py:69-77).  `locked`: n_locked orthonormal vectors (pitch ld_locked) the search is kept orthogonal
 * to -- eigenpairs beyond the first are computed by deflation. */
typedef int (*tn_matvec_fn)(const double* x /* [dev] n */, double* y /* [dev] n */, void* user, void* stream);
int tn_lanczos_generic(tn_matvec_fn matvec, void* user, long long n, double tau, const double* v0 /* [dev] */, double tol,
                       int ncv, int max_restarts, const double* locked /* [dev] or NULL */, int n_locked,
                       long long ld_locked, double* lambda_out, double* vec_out /* [dev] */, int* n_matvec_out,
                       double* resid_out, void* workspace /* [dev] */, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Communicator.  NCCL is bound at run time (dlopen libnccl.so.2; a process that already carries one,
 * e.g. torch's, shares it).  tn_comm_init adopts a caller-owned ncclComm_t; tn_comm_unique_id + tn_comm_init_rank create
 * one (the 128-byte id travels through the caller's own bootstrap, e.g. a torch.distributed broadcast).
 * All collectives are in place on [dev] float64 buffers and stream-ordered.
 * -------------------------------------------------------------------------------------------------- */
int tn_comm_unique_id(char* id128 /* [host] 128 bytes out */);
int tn_comm_init_rank(tn_comm** comm, const char* id128 /* [host] */, int rank, int world);
int tn_comm_init(tn_comm** comm, void* nccl_comm /* ncclComm_t */, int rank, int world);
int tn_comm_rank(const tn_comm* comm);
int tn_comm_world(const tn_comm* comm);
long long tn_comm_collectives(const tn_comm* comm); /* collectives issued so far through this handle */
/* ... of which ran as kernels storing into the peer window (one IPC-mapped buffer per rank over NVLink: the small all-reduces
 * and the Krylov-vector all-gathers of the row-sliced Lanczos) instead of NCCL launches; 0 when the box does not allow CUDA IPC
 * between the ranks or TNALG_NO_PEER is set */
long long tn_comm_peer_collectives(const tn_comm* comm);
int tn_comm_destroy(tn_comm* comm);
int tn_comm_allreduce_sum(tn_comm* comm, double* buf /* [dev] */, long long count, void* stream);
int tn_comm_allgather(tn_comm* comm, const double* send /* [dev] count_per_rank */, double* recv /* [dev] world*count_per_rank */,
                      long long count_per_rank, void* stream);
/* n broadcasts in ONE grouped launch (the outgoing operators of a sharded environment update) */
int tn_comm_broadcast_many(tn_comm* comm, double* const* bufs /* [host] of [dev] */, const long long* counts /* [host] */,
                           const int* roots /* [host] */, int n, void* stream);

/* --------------------------------------------------------------------------------------------------
 * One-sided Jacobi SVD (a7 'svd' branch, a11, a12): A (m,n) row-major = U diag(S) Vt, singular values sorted
 * in decreasing order, k = min(m,n) triplets of which the first k_keep are written (truncation to chi,
 * library/MPSClass.py:186-247,1676-1686).  U is (m,k_keep), S (k_keep), Vt (k_keep,n); U or Vt may be NULL.
 * Blocking.  sweeps_out [host] receives the number of Jacobi sweeps.  Callers should pass the transposed triangular factor
 * of a QR factorisation (Drmac-Veselic preconditioning, tnalg_b200/ops.py:svd): ~8 sweeps and high relative accuracy of the
 * small singular values, versus 20-50 sweeps on a raw ill-conditioned matrix.
 * -------------------------------------------------------------------------------------------------- */
size_t tn_svd_workspace_bytes(int m, int n);
int tn_svd_jacobi(const double* A /* [dev] */, int m, int n, int k_keep, double* U /* [dev] */, double* S /* [dev] */,
                  double* Vt /* [dev] */, int* sweeps_out, void* workspace /* [dev] */, size_t workspace_bytes,
                  void* stream);

/* --------------------------------------------------------------------------------------------------
 * Observables (a10: observation_s1 / observation_s1_s2, MPSClass.py:857-909; observe_magnetization :911, observe_bond_energy :923):
 * <op(site)> and <op1(site1) op2(site2)>, site1 < site2, on an MPS whose sites < center are left- and sites > center
 * right-orthonormal.  mps: [host] array of L [dev] tensors, tensor s of shape (dims[s], d, dims[s+1]); operators row-major d x d
 * [host], one (pair) per term; out: [host] n_terms doubles.  Blocking.  Each term is one chain of tn_env_update transfers and a
 * trace (the Python layer batches terms that share chain prefixes; this is the per-term form for direct bindings).
 * -------------------------------------------------------------------------------------------------- */
size_t tn_expect_workspace_bytes(const int* dims /* [host] L+1 */, int L, int d, int n_terms);
int tn_expect_1body(const double* const* mps, const int* dims, int L, int d, int center, int n_terms, const int* sites /* [host] */,
                    const double* ops /* [host] n_terms*d*d */, double* out, void* workspace /* [dev] */, size_t workspace_bytes,
                    void* stream);
int tn_expect_2body(const double* const* mps, const int* dims, int L, int d, int center, int n_terms, const int* site1,
                    const int* site2, const double* ops1, const double* ops2, double* out, void* workspace /* [dev] */,
                    size_t workspace_bytes, void* stream);

/* Symmetric eigenproblem on the same Jacobi kernels (the north_star's "Jacobi SVD/eigh"): A (n,n) symmetric -> w (n) ascending,
 * V (n,n) row-major with eigenvector j in column j.  The dense local solve of eig_way = 0 (MPSClass.py:792-794). Blocking. */
size_t tn_eigh_workspace_bytes(int n);
int tn_eigh_jacobi(const double* A /* [dev] */, int n, double* w /* [dev] */, double* V /* [dev] */, int* sweeps_out,
                   void* workspace /* [dev] */, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Householder QR of the gauge moves (a7: np.linalg.qr in left2right/right2left_decompose_tensor,
 * TensorBasicModule.py:342-345,378-380; every step of the sweep DMRG_anyH.py:47-64 moves the centre with it).
 * A (m,n) = Q (m,k) R (k,n), k = min(m,n); LAPACK's reflector convention, so R matches np.linalg.qr including row signs for
 * full-rank input; Q^T Q = 1 to rounding for any conditioning.  Blocked (panel 32): the panel is factored by one thread-block
 * cluster with one distributed-shared-memory reduction per column, trailing updates and the explicit Q run on FP64 tensor
 * cores.  Deterministic.  trans_in: the input is stored as A^T (n,m); trans_q: Q is written as Q^T (k,m).
 * tn_qr_l2r / tn_qr_r2l are the two gauge moves on an MPS tensor T (a,d,b):
 *   l2r: T.reshape(a*d, b) = Q R        -> Q_out (a,d,k), R_out (k,b)
 *   r2l: T.reshape(a, d*b)^T = Q R      -> Q_out (k,d,b) (= Q^T), R_out (k,a)
 * -------------------------------------------------------------------------------------------------- */
size_t tn_qr_workspace_bytes(int m, int n);
int tn_qr_householder(const double* A /* [dev] */, int m, int n, int trans_in, double* Q /* [dev] */, int trans_q,
                      double* R /* [dev] */, void* workspace /* [dev] */, size_t workspace_bytes, void* stream);
int tn_qr_l2r(const double* T /* [dev] (a,d,b) */, int a, int d, int b, double* Q_out, double* R_out, void* workspace,
              size_t workspace_bytes, void* stream);
int tn_qr_r2l(const double* T /* [dev] (a,d,b) */, int a, int d, int b, double* Q_out, double* R_out, void* workspace,
              size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Exact-diagonalisation cross-checker (library/EDspinClass.py:69-77 project_all_hamilt and
 * algorithms/ExactDiagonalizationAlgo.py:12-24 exact_ground_state): the full-space operator on a d^L vector (site 0 slowest)
 *   out = c_id v + c_h sum_n h[h_index[n]] on sites (p1[n], p2[n]),   hs: n_h matrices (d^2 x d^2), index (s_p1, s_p2),
 * and its ground state through the device-resident Lanczos (dominant eigenpair of 1 - tau*H, like eigsh(..., which='LM')).
 * -------------------------------------------------------------------------------------------------- */
size_t tn_ed_workspace_bytes(int L, int d, int n_terms, int n_h, int ncv);
int tn_ed_apply(double* out /* [dev] */, const double* v /* [dev] */, int L, int d, int n_terms, const int* p1 /* [host] */,
                const int* p2 /* [host] */, const int* h_index /* [host] */, const double* hs /* [host] */, int n_h, double c_id,
                double c_h, void* workspace /* [dev] */, size_t workspace_bytes, void* stream);
int tn_ed_ground_state(int L, int d, int n_terms, const int* p1, const int* p2, const int* h_index, const double* hs, int n_h,
                       double tau, const double* v0 /* [dev] */, double tol, int ncv, int max_restarts, double* lambda_out,
                       double* vec_out /* [dev] */, int* n_matvec_out, double* resid_out, void* workspace /* [dev] */,
                       size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNALG_B200_H */
