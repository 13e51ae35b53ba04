/* ipm_b200.h -- C ABI of libipm_b200.so: the B200 (sm_100a) hot path of a log-barrier interior-point engine
 * and a batched ADMM Lasso, drop-in behind the Python API of fdeguire03/InteriorPoint-GPU.
 *
 * The reference has NO native boundary: its "GPU arm" is CuPy calls selected by `if self.use_gpu:` inside the
 * Python classes (e.g. FunctionManager.py:122-125, NewtonSolver.py:285-313).  Each entry point below names the
 * reference call sites (file:line, relative to the reference repo) whose arithmetic it replaces; the host-side
 * binding is ctypes (interiorpoint-gpu_b200/_abi.py), see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to FP64 data unless stated otherwise; matrices are row-major with an
 *     explicit leading dimension in ELEMENTS.  Operands of ipm_gemm_tn_f64 (and everything built on it) need a
 *     16-byte aligned base and an even leading dimension (TMA row stride); the engine pads ld to 16 doubles.
 *   - symmetric matrices: only the UPPER triangle (col >= row) is read / written.  Cholesky is H = U^T U.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream); nothing here
 *     synchronises or throws.  Workspace comes from the caller (*_ws_doubles()); the only device memory the library
 *     owns is small per-(device, stream) scratch allocated on first use and kept for the life of the process: the
 *     stream-K partials of ipm_gemm_tn_f64 (#SMs x 128 KB), the progress counters of the persistent solves and of
 *     ipm_potrf_upper_dag_f64 (a few KB).  So the FIRST call on a stream must not happen inside a CUDA graph capture.
 *   - return value: IPM_OK, or a negative status.  Numerical failure is NOT a status: ipm_potrf_upper_f64 writes
 *     LAPACK-style `info` to device memory, the line searches write a `stuck` flag (the reference signals these
 *     with LinAlgError / success_flag=False, NewtonSolver.py:130-131,314-330).
 */
#ifndef IPM_B200_H
#define IPM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define IPM_OK 0
#define IPM_ERR_ARG (-1)       /* invalid size / alignment / null pointer */
#define IPM_ERR_CUDA (-2)      /* CUDA runtime error, text via ipm_last_cuda_error() */
#define IPM_ERR_NO_DEVICE (-3) /* no sm_100 device or driver entry point */

/* ---- library status ------------------------------------------------------------------------------------ */
int ipm_abi_version(void);
int ipm_device_ok(void);                       /* IPM_OK iff the current device is compute capability 10.x */
const char* ipm_last_cuda_error(void);
unsigned long long ipm_launch_count(void);     /* kernels launched by this library in this process */
/* Watchdog of the device-side spin waits (persistent kernels whose CTAs -- or whose peers on other GPUs -- wait for
 * each other: tile-DAG Cholesky, stream-K GEMM, triangular solves, the peer-memory Hessian exchange).  A wait that
 * exceeds the limit records a code (1 potrf-dag, 2 stream-K, 3/4 peer Hessian, 5 trsv, 6 lasso, 7 peer potrf, 8 INT8 Hessian), every
 * other wait of the process then returns at once, the kernels finish on garbage in bounded time, and potrf reports
 * info = -1.  Read it after synchronising the stream; 0 = no fault.  The reference has no counterpart (a failed CuPy
 * kernel raises on the next call). */
unsigned int ipm_device_fault(void);
void ipm_clear_device_fault(void);
unsigned int ipm_set_spin_limit(unsigned int mcycles);  /* limit in units of 2^20 SM cycles; returns the old value */

/* ---- dense contraction core (TMA + FP64 DMMA) ----------------------------------------------------------- */
/* D = beta*D + alpha * A^T diag(w) B.   A: K x M (lda), B: K x N (ldb), w: K or NULL, D: M x N (ldd).
 * upper = 0: all of D;  1: M == N, only tiles/elements with col >= row (SYRK);  2: all tiles, elements col >= row.
 * Replaces cp.matmul(C.T, (inv_s**2)[:,None]*C)  FunctionManager.py:301-312, 564-576, 801-813 (Hessian),
 * the per-cone Hessian loop FunctionManager.py:1119-1144, 1396-1422, cp.matmul(A, A11_inv_AT)
 * NewtonSolverInfeasibleStart.py:426,474,780,796 (Schur) and Qinv @ (u - alpha) LassoSolver.py:245-249. */
int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha, double beta,
                    double* D, int ldd, int M, int N, int K, int upper, void* stream);

/* ---- barrier Hessian on the INT8 tensor pipe (tcgen05.mma.kind::i8, FP64-accurate by error-free slicing) ------ */
/* H (upper tiles, n x n, ldh) = beta*H + C^T diag(w) C for C: m x n (ldc) and w >= 0 -- the contraction of
 * ipm_gemm_tn_f64(C, C, w, upper = 1), FunctionManager.py:301-312, 564-576, 801-813 -- computed as slices*(slices+1)/2
 * exact INT8 x INT8 -> INT32 products of `slices` 7-bit digits per entry of diag(sqrt w) C (csrc/hess_i8.cu).  slices = 8
 * matches the FP64 DMMA kernel to ~2e-15 of sum_k |x_ki x_kj|; 1 <= slices <= 8, m <= 65408.  ws: 256-byte aligned
 * device memory of ipm_hess_i8_ws_bytes(m, n, slices) bytes (0 = unsupported shape), set up ONCE by ipm_hess_i8_prepare
 * (synchronises the stream) and then private to calls with the same (m, n, slices). */
long long ipm_hess_i8_ws_bytes(int m, int n, int slices);
int ipm_hess_i8_prepare(void* ws, int m, int n, int slices, void* stream);
int ipm_hess_i8_f64(const double* C, int ldc, int m, int n, const double* w, double beta, double* H, int ldh, int slices,
                    void* ws, void* stream);
/* Row-sharded variant: the partial C_r^T diag(w) C_r of this rank's m local rows, filed tile by tile under the owners in
 * this rank's exchange buffer and announced through the owners' flags (same peer arrays, slots and epoch as
 * ipm_syrk_scatter_f64 below; no local addend); to be followed by ipm_hess_reduce_bcast_pull_f64, in which the owners
 * pull the tiles over NVLink. */
int ipm_hess_i8_scatter_f64(const double* C, int ldc, int m, int n, const double* w, int slices, void* ws,
                            void* const* peer_inbox, void* const* peer_flags, int me, int R, int slots,
                            unsigned int epoch, void* stream);

/* ---- HBM-streaming level-1/2 ---------------------------------------------------------------------------- */
/* y = alpha * M x + beta * y   (M: rows x cols).  np.matmul(C, x) FunctionManager.py:123-125, 432-434, 939-941;
 * np.matmul(P, x) :694-699, 759-761; A @ x NewtonSolverInfeasibleStart.py:199-205. */
int ipm_gemv_n_f64(const double* M, int ld, int rows, int cols, const double* x, double* y, double alpha, double beta,
                   void* stream);
/* Y[v] = alpha * M^T V[v] + beta * Y[v],  v < nv <= 2 (deterministic two-stage column sums).
 * np.matmul(C.T, inv_slacks) FunctionManager.py:256-262, 527-529, 568-570; A.T @ v
 * NewtonSolverInfeasibleStart.py:197-203. */
long long ipm_gemv_t_ws_doubles(int rows, int cols, int nv);
int ipm_gemv_t_f64(const double* M, int ld, int rows, int cols, const double* V, int nv, int ldv, double* Y, int ldy,
                   double alpha, double beta, double* ws, long long ws_doubles, void* stream);
/* ---- row-sharded Hessian over peer memory (no reference counterpart; BASELINE north_star / SURVEY 8(e)) ------- */
/* Partial Hessian alpha * C_r' diag(w) C_r of this rank's K local rows, scattered tile by tile (128 x 128, upper
 * triangle, tile t owned by rank t % R) into the owners' inboxes over NVLink, with a system-scope arrival flag per
 * (tile, source).  peer_inbox / peer_flags: HOST arrays of R device pointers (symmetric allocations):
 * inbox [R sources][slots][128*128] doubles, flags [slots][R] u32.  epoch != 0, identical on all ranks. */
int ipm_syrk_scatter_f64(const double* C, int ldc, const double* w, int n, int K, double alpha, const double* base,
                         int ldb, void* const* peer_inbox, void* const* peer_flags, int me, int R, int slots,
                         unsigned int epoch, void* stream);
/* For every owned tile: wait for the R partials, add them in rank order (+ tP * P), write the final tile into every
 * rank's H (upper part) and bump that rank's completion counter; then park the stream until this rank's own counter
 * reaches done_target (callers add #tiles per step).  All ranks end with bit-identical H. */
int ipm_hess_reduce_bcast_f64(const double* inbox, const unsigned int* flags, void* const* peer_H,
                              void* const* peer_done, int ldh, int n, int me, int R, int slots, unsigned int epoch,
                              unsigned int done_target, const double* P, int ldp, double tP, void* stream);
/* The same after ipm_hess_i8_scatter_f64: the partial tiles are read from the buffers of the ranks that computed them
 * (peer_inbox: R device pointers, host array) with coalesced loads over NVLink instead of being pushed. */
int ipm_hess_reduce_bcast_pull_f64(void* const* peer_inbox, const unsigned int* flags, void* const* peer_H,
                                   void* const* peer_done, int ldh, int n, int me, int R, int slots, unsigned int epoch,
                                   unsigned int done_target, const double* P, int ldp, double tP, void* stream);

/* ---- L2 residency ----------------------------------------------------------------------------------------- */
/* Mark [base, base + bytes) as persisting in L2 for kernels launched on `stream` (bytes == 0: clear).  *ratio_out
 * receives the configured hit ratio (carve-out / window).  No counterpart in the reference: the B200's 126 MB L2
 * holds the whole ADMM state (84 MB at K = 4096) that LassoSolver.py:240-337 re-reads every iteration. */
int ipm_l2_persist(const void* base, unsigned long long bytes, double* ratio_out, void* stream);

/* ---- sparse-aware variants (MIPLIB `.npy` LPs, testSolver.py:278-300, >99 % zeros) ------------- */
/* y = alpha * S x + beta * y for S in CSR (int32 rowptr[rows+1], col, val).  Same call sites as ipm_gemv_n_f64 with
 * CSR(C); with CSR(C^T) it replaces ipm_gemv_t_f64 (deterministic, no atomics). */
int ipm_csr_gemv_f64(const int* rowptr, const int* col, const double* val, int rows, const double* x, double* y,
                     double alpha, double beta, void* stream);
/* H[out_i[e]][out_j[e]] += sum_{k in [segptr[e], segptr[e+1])} w[seg_row[k]] * seg_prod[k]   for e < nout: the
 * structurally non-zero upper-triangle entries of C^T diag(w) C (segments precomputed on the host, ascending row
 * order).  Replaces the Hessian SYRK of FunctionManager.py:301-312 for sparse C. */
int ipm_sparse_syrk_f64(int nout, const int* segptr, const int* seg_row, const double* seg_prod, const int* out_i,
                        const int* out_j, const double* w, double* H, int ld, void* stream);
/* out[k] = a[k] . b[k], k < npairs <= 8; a, b, n are HOST arrays of device pointers / lengths.
 * gradf.dot(x), gradf.dot(xstep) NewtonSolver.py:129,168; c.dot(x) FunctionManager.py:159. */
int ipm_dots_f64(int npairs, const double* const* a, const double* const* b, const int* n, double* out, void* stream);
/* y += (*a_dev) * x with the scalar on the device.  x += step_size * xstep NewtonSolver.py:103. */
int ipm_axpy_dev_f64(int n, const double* a_dev, const double* x, double* y, void* stream);
/* out = x + (*a_dev) * dx.  next_x = x + step_size * xstep NewtonSolver.py:167,182. */
int ipm_trial_point_f64(int n, const double* a_dev, const double* x, const double* dx, double* out, void* stream);
/* out = ca*a + cb*b + cc*c (b, c may be NULL). */
int ipm_lincomb3_f64(int n, double ca, const double* a, double cb, const double* b, double cc, const double* c,
                     double* out, void* stream);
/* op 0: out = alpha*a*b;  1: out = alpha*a/b;  2: out = alpha/a.   -H_inv * gradf NewtonSolver.py:417-418. */
int ipm_vec_op_f64(int op, int n, const double* a, const double* b, double* out, double alpha, void* stream);
/* out[i][j] = s*in[i][j] (+ dshift on the diagonal).  Qinv_cache *= -m*rho LassoSolver.py:218. */
int ipm_scale_shift_f64(const double* in, int ldi, double* out, int ldo, int rows, int cols, double s, double dshift,
                        void* stream);

/* ---- conjugate-gradient Newton solves (linear_solve_method="cg") ----------------------------------------- */
/* scipy.sparse.linalg.cg semantics (rtol on ||b||, at most maxiter steps, last iterate returned) for (sign*H) x = b with
 * H dense symmetric; no host synchronisation (the scalars and the convergence flag stay on the device).
 * NewtonSolverCG NewtonSolver.py:365-400: cg(-H, gradf, x0, maxiter); PhaseOne.py:137-150: cg(hess, -grad, x0=[x, s]). */
int ipm_symmetrize_upper_f64(double* H, int ld, int n, void* stream);   /* H[j][i] = H[i][j], j > i */
long long ipm_cg_ws_doubles(int n);
/* x0 = (x.g < 0) ? -(x.g) x / (x.Hx) : 0 and Hx0 = H x0 (NewtonSolver.py:379-383); Hx = H x from ipm_gemv_n_f64. */
int ipm_cg_descent_x0_f64(int n, const double* x, const double* g, const double* Hx, double* x0, double* Hx0,
                          void* stream);
/* x: x0 in / solution out; Hx0: H x0 or NULL for x0 = 0; ws[3n+2] != 0 when converged, ws[3n+3] = iterations. */
int ipm_cg_solve_f64(const double* H, int ld, int n, const double* b, double* x, const double* Hx0, double sign,
                     int maxiter, double rtol, double* ws, void* stream);

/* ---- factorisation and triangular solves ---------------------------------------------------------------- */
/* In-place H = U^T U on the upper triangle; *info_dev (device int) = 0 or the 1-based index of the first
 * non-positive pivot.  Uses a library-owned high-priority side stream for the panel chain (joined before return).  scipy.linalg.cho_factor / cp.linalg.cholesky NewtonSolver.py:286,303;
 * NewtonSolverInfeasibleStart.py:398,426,455,473,780,795; LassoSolver.py:160,178. */
int ipm_potrf_upper_f64(double* H, int ld, int n, int* info_dev, void* stream);
/* Same contract, ONE persistent launch: left-looking tile DAG with device-side flags (csrc/chol.cu, namespace dag).
 * Used by ipm_potrf_upper_f64 itself for n >= 6144 (IPM_POTRF_DAG=1: every admissible size, =0: never); sizes it does not cover
 * (n <= 256, n > 32768) go through the stream-ordered code.  One factorisation at a time per stream. */
int ipm_potrf_upper_dag_f64(double* H, int ld, int n, int* info_dev, void* stream);
/* The same factorisation distributed over the R <= 8 GPUs of a node (north_star: "replicated or 2D-block-cyclic
 * factorisation"): block column j belongs to rank j % R; the owner pushes every finished 32-row step into ALL R copies
 * of the matrix over NVLink and release-stores the progress counter of every rank, so loads and waits stay local and
 * every rank ends with the whole factor.  Called by every rank at the same point of its stream.  peer_H / peer_info /
 * peer_prog: R device pointers each (host arrays) to every rank's matrix copy, info word and progress counters
 * (ipm_potrf_peer_prog_words() 64-bit words, zeroed once); epoch != 0, new for every call, identical on all ranks;
 * max_ctas = 0 (one CTA per SM).  n <= 256: IPM_ERR_ARG (factor replicated).  NewtonSolver.py:286,303. */
int ipm_potrf_peer_prog_words(void);
int ipm_potrf_upper_peer_f64(void* const* peer_H, int ld, int n, void* const* peer_info, void* const* peer_prog, int me,
                             int R, unsigned int epoch, int max_ctas, void* stream);
/* H = U^T U in place AND B <- U^{-T} B (B: n x p) in ONE persistent launch: the right-hand sides are extra block columns
 * of the tile DAG.  cho_factor(H) + the forward half of cho_solve(L1, A.T) NewtonSolverInfeasibleStart.py:398-426
 * (the Schur complement is then Y'Y).  Sizes the tile-DAG kernel does not take: potrf followed by trsm. */
int ipm_potrf_trsm_upper_f64(double* H, int ld, int n, double* B, int ldb, int p, int* info_dev, void* stream);
/* b <- U^{-T} b (trans = 1) or U^{-1} b (trans = 0), in place; ws is unused (kept for ABI stability, may be NULL).
 * One persistent launch; not re-entrant per device (two concurrent solves on different streams of one device would
 * share the block flags).  cho_solve / solve_triangular NewtonSolver.py:287-313. */
int ipm_trsv_upper_f64(const double* U, int ld, int n, double* b, int trans, double* ws, void* stream);
/* B <- U^{-T} B, B: n x p.  cho_solve(L1, A.T) NewtonSolverInfeasibleStart.py:399-411,460-465;
 * cho_solve(L, I) LassoSolver.py:163-188. */
int ipm_trsm_upper_t_f64(const double* U, int ldu, int n, double* B, int ldb, int p, void* stream);

/* ---- barrier evaluation: linear inequalities (LP / QP / phase-I) ---------------------------------------- */
/* Slacks [m rows | ub | lb], reciprocals, SYRK weights w = inv^2, bound diagonal, reductions
 * red_out[5] = {sum log(s+1e-15), min s, sum inv, sum inv^2, #(s<0)}.  s_ptr: phase-I slack variable (device) or
 * NULL.  hdiag_guard g: diagonal 1/(s+g)^2.  update_slacks_fxn FunctionManager.py:118-149, 429-449;
 * newton_objective :208-230, 484-507; inv_slacks :243-247. */
long long ipm_lin_barrier_ws_doubles(void);
int ipm_lin_barrier_eval_f64(int m, int n, const double* Cx, const double* d, const double* x, const double* ub,
                             const double* lb, const double* s_ptr, int phase1, double hdiag_guard, double* slacks,
                             double* inv, double* w, double* hdiag, double* red_out, double* ws, void* stream);
/* g = t*lin - inv_lb + inv_ub + CtInv (main) or the phase-I gradient / Hessian border.
 * gradient FunctionManager.py:232-265, 509-545, 741-781; hess_xs :568-587. */
int ipm_lin_grad_f64(int n, double t, const double* lin, const double* CtInv, const double* inv_ub,
                     const double* inv_lb, int phase1, const double* suminv, const double* CtW, double* g, double* hxs,
                     void* stream);
/* H[i][i] += hdiag[i] + shift; optional border column H[i][n] = border[i], H[n][n] = *hss + shift.
 * FunctionManager.py:314-322, 589-607; add_psd_conditioning NewtonSolver.py:269-275. */
int ipm_hess_finish_f64(double* H, int ld, int n, const double* hdiag, const double* border, const double* hss,
                        double shift, void* stream);
/* H(upper) = t * P(upper) (P NULL: zero).  self.hess = self.t * self.P FunctionManager.py:797, 1117. */
int ipm_scale_copy_upper_f64(double* H, int ldh, const double* P, int ldp, int n, double t, void* stream);

/* ---- barrier evaluation: second-order cones ------------------------------------------------------------- */
/* Cone slacks s_i = rhs_i^2 - |lhs_i|^2 (+ *s_ptr), SYRK weights, coefficients of the per-cone gradient rows and
 * the reductions (merged with red_bounds[5] from the bound part).  FunctionManager.py:933-994, 1258-1262. */
int ipm_cone_eval_f64(int M, const int* cone_off, int ktot, const double* lhs, const double* rhs, const double* s_ptr,
                      double guard, int tail_off, double* slacks, double* inv_c, double* wts, double* coefA,
                      double* coefC, double* plog, double* pinv, const double* red_bounds, double* red_out,
                      void* stream);
/* G[i] = sum_{r in cone i} coefA[r] A[r] + coefC[i] c_i = 2/(s_i+eps) (A_i' lhs_i - c_i rhs_i).
 * FunctionManager.py:1076-1093, 1124-1137, 1341-1358, 1401-1413. */
int ipm_cone_grad_rows_f64(int M, int n, const int* cone_off, const double* A, int lda, const double* Cc, int ldc,
                           const double* coefA, const double* coefC, double* G, int ldg, void* stream);
/* Quadratic of every cone slack (and linear of every cone rhs) along x + a*dx. */
int ipm_cone_ls_coeffs_f64(int M, const int* cone_off, const double* lhs, const double* rhs, const double* dlhs,
                           const double* drhs, const double* ds_ptr, int tail_off, double* p1, double* p2,
                           void* stream);

/* ---- line searches (device side; `table` = 1, beta, beta^2, ... down to the first entry < 1e-13) --------- */
/* Feasibility back-off: *kmax = first table index keeping every slack >= 0.  NewtonSolver.py:170-183;
 * NewtonSolverInfeasibleStart.py:179-193. */
int ipm_ls_feas_lin_f64(int m, int n, const double* slacks, const double* Cdx, const double* dz, int has_ub,
                        int has_lb, int phase1, const double* table, int len, double* p1_out, int* kmax,
                        void* stream);
int ipm_ls_feas_poly_f64(int count, const double* s0, const double* p1, const double* p2, const double* table,
                         int len, int* kmax, int reset, void* stream);
/* Armijo loop with the reference's semantics (slope g.x, frozen barrier term, lagging trial point).
 * out[5] = {step, stuck, index, frozen log-sum, trials}.  NewtonSolver.py:185-206.  textbook != 0: the plain Armijo
 * rule of the stand-alone phase-I (slope g.dx = terms[4], barrier re-evaluated per trial; PhaseOne.py:187-218).
 * update_slacks_every > 0 (NewtonSolver.py:196-202): the barrier term is refreshed every so many trials -- from the slack
 * polynomial, or, when L_direct is given (cones), by the host: the call returns stuck = 4 with *kmax = index of the
 * lagging point; evaluate the barrier there into *L_direct and call again with resume = 1. */
int ipm_ls_armijo_f64(int nc, const double* s0, const double* p1, const double* p2, const double* table, int len,
                      int* kmax, const double* sumlog, const double* terms, double t, double alpha,
                      int update_slacks_every, const double* L_direct, const double* nneg, int textbook, int resume,
                      double* out, void* stream);
/* out[0] = sum log(s(table[*kmax]) + 1e-15) over this rank's slack entries (row-sharded problems all-reduce it and
 * pass the result as L_direct). */
int ipm_ls_logsum_f64(int nc, const double* s0, const double* p1, const double* p2, const double* table, int len,
                      const int* kmax, double* out, void* stream);
/* Residual-norm search of the infeasible-start method.  out[6] = {step, stuck, index, r0, r(step), trials}.
 * NewtonSolverInfeasibleStart.py:207-273.  update_slacks_every > 0 (:249-255): when a refresh is due the call returns
 * stuck = 4 with *kmax = the trial index; re-evaluate the barrier gradient there into u0 and call again with resume = 1. */
int ipm_ls_residual_f64(int n, int p, const double* r0d, const double* u0, const double* u1, const double* q0,
                        const double* q1, const double* table, int len, int* kmax, double alpha, const double* nneg,
                        int update_slacks_every, int resume, double* out, void* stream);
int ipm_table_lookup_f64(const double* table, int len, const int* kmax, double* out, void* stream);

/* ---- batched ADMM Lasso ---------------------------------------------------------------------------------- */
/* One ADMM iteration for K problems: x = bA + Q~ z; alpha+ = prox(x + u, eta); u+ = u + x - alpha+;
 * z_out = u+ - alpha+; on request the four squared norms of the stop test.  LassoSolver.py:245-253, 273-298,
 * 517-543. */
long long ipm_lasso_partials_doubles(int n, int K);
int ipm_lasso_admm_step_f64(const double* Qt, int ldq, int n, int K, const double* bA, const double* eta, double rho,
                            double* alpha, double* u, const double* z_in, double* z_out, int ld, int add_bias,
                            int positive, int want_norms, double* partials, double* norms_out, void* stream);
/* `n_iters` ADMM iterations in ONE persistent launch (csrc/lasso_multi.cu): the loop body of LassoSolver.py:240-337
 * including the batch-coupled stop test :273-298, evaluated on the device after the last iteration when want_norms:
 *   ||x - alpha+|| < stop_mult + eps_rel ||alpha+||  and  ||rho (alpha+ - alpha)|| < stop_mult + eps_rel rho ||u+||.
 * z0 / z1: the two buffers of z = u - alpha (after `it` iterations in total the current one is z[it & 1]).
 * ws: device workspace of ipm_lasso_steps_ws_bytes() bytes, zeroed by the caller at the start of a solve;
 * ((int*)ws)[0] = 1 once the test held (later launches are then no-ops), ((int*)ws)[1] = iterations performed; the
 * four squared norms of the last test sit at byte offset ipm_lasso_steps_norms_offset().  alpha is valid after every
 * launch. */
long long ipm_lasso_steps_ws_bytes(int n, int K);
long long ipm_lasso_steps_norms_offset(int n, int K);
int ipm_lasso_admm_steps_f64(const double* Qt, int ldq, int n, int K, const double* bA, const double* eta, double rho,
                             double* alpha, double* u, double* z0, double* z1, int ld, int add_bias, int positive,
                             int n_iters, int want_norms, double stop_mult, double eps_rel, void* ws, void* stream);
/* f_c = 1/(2m)|R[:,c]|^2 + reg_c * |alpha[1:,c]|_1 with R = A alpha - b.  LassoSolver.py:314-325. */
int ipm_lasso_objective_f64(const double* R, int ldr, int m, const double* alpha, int lda, int n, int K,
                            const double* reg, int add_bias, int positive, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPM_B200_H */
