/* qmprs_b200 -- C ABI of the B200 (sm_100a) kernels behind the qmprs MPS hot path.
 *
 * Conventions
 *   - every matrix is row-major complex128 (interleaved re,im; 16 bytes per element),
 *     leading dimensions are in elements;
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, a cudaError_t (>0) on a CUDA failure, <0 on a
 *     workspace/argument error.  The Python host maps non-zero to RuntimeError.
 *
 * Each entry point names the reference interface it replaces.  The reference
 * (Qualition/qmprs) is pure Python: the "FFI" it binds today is numpy/scipy LAPACK and
 * quimb's tensor routines, reached from the cited lines.
 */
#ifndef QMPRS_B200_H
#define QMPRS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- dense linear algebra ------------------------------------------------------ */

/* C = alpha*op(A)*B + beta*C, op(A) = A (trans_a = 0) or A^H with A stored k x m (trans_a = 1).  Replaces quimb tensordot/tensor_contract reached from
 * qmprs/primitives/mps.py:270 (to_dense), :451-453 (compress), :968-971 (gate_split_).
 * batch > 1 strides the three operands. */
int qm_zgemm(int m, int n, int k, double alpha_re, double alpha_im, const void* A, long long lda,
             const void* B, long long ldb, double beta_re, double beta_im, void* C, long long ldc,
             int batch, long long strideA, long long strideB, long long strideC, int trans_a, void* stream);

/* Thin SVD A = U diag(S) Vh by blocked one-sided Jacobi.  Replaces numpy/LAPACK zgesdd
 * behind quimb tensor_split: mps.py:242 (from_dense), :451-453, :928-931, :968-971;
 * sequential.py:443.  S is double[k], sorted descending, k = min(m,n).  U/Vh may be NULL.
 * info_host: host int[2] = {sweeps, converged} (may be NULL).
 * flags: QM_SVD_BACKMULT -- matrices with more than 32 short vectors only rotate W (no [W | I] extension
 * that accumulates the rotations); the factor on the long side comes from the converged rows as before and the
 * other one from a single ZGEMM against the input (U = A Z^H Sigma^-1 resp. Vh = Sigma^-1 U^H A): a third less
 * tensor work per sweep on square matrices.  Column j of that factor is accurate to eps * sigma_max / sigma_j, which
 * is what the MPS path needs (it multiplies the factor by sigma_j or sqrt(sigma_j) right away); the default (0)
 * keeps both factors orthonormal to rounding whatever the spectrum. */
#define QM_SVD_BACKMULT 1
long long qm_svd_work_bytes(int m, int n);
int qm_svd(int m, int n, const void* A, long long lda, void* U, long long ldu, void* S, void* Vh,
           long long ldvh, void* work, long long work_bytes, double tol, int max_sweeps, int* info_host,
           int flags, void* stream);

/* Sync-free SVD for CUDA-graph capture: exactly fixed_sweeps sweeps are enqueued (kernels return
 * immediately once a device-side flag says the iteration converged); mismatch[0] (int) is set to 1
 * if it had not converged, in which case the caller re-runs the problem through qm_svd. */
int qm_svd_static(int m, int n, const void* A, long long lda, void* U, long long ldu, void* S, void* Vh,
                  long long ldvh, void* work, long long work_bytes, double tol, int fixed_sweeps, void* mismatch,
                  int flags, void* stream);

/* Small problems (min(m,n) <= 64 and the work matrix within one SM's shared memory, e.g. every matrix of a
 * 12-qubit / chi=64 register): `batch` independent thin SVDs of one shape, ONE CTA each, the whole Jacobi
 * iteration inside the kernel -- no workspace, no host synchronisation, CUDA-graph capturable.  Same outputs and
 * flags as qm_svd; strides in elements between consecutive problems (0 for batch = 1); U / Vh may be NULL.
 * mismatch (optional int[1]) is set to 1 if a problem has not converged after max_sweeps sweeps.
 * qm_svd_small_fits: 1 if the shape is supported (else qm_svd_small returns -3).
 * Same reference call sites as qm_svd (numpy.linalg.svd behind quimb tensor_split: mps.py:242, :451-453, :881,
 * :928-931, :968-971) wherever min(m, n) <= 64. */
int qm_svd_small_fits(int m, int n, int flags);
int qm_svd_small(int m, int n, const void* A, long long lda, long long strideA, void* U, long long ldu,
                 long long strideU, void* S, long long strideS, void* Vh, long long ldvh, long long strideVh,
                 double tol, int max_sweeps, int flags, int batch, void* mismatch, void* stream);

/* out (cols x rows) = in^T (optionally conjugated): the coalesced reshape/transposed-store kernels
 * that stream the statevector in the TT-SVD (quimb from_dense reshapes, mps.py:242), exposed for
 * tests and HBM-bandwidth measurement (32*rows*cols bytes per call). */
int qm_transpose(void* out, long long ldo, const void* in, long long ldi, long long rows, long long cols, int conj,
                 void* stream);

/* Householder QR (LAPACK zgeqr2 layout), explicit thin Q, and R with non-negative
 * diagonal.  Replaces quimb qr_stabilized behind left_canonize / right_canonize /
 * tensor_compress_bond: mps.py:396-398, :451-453. */
int qm_qr(int m, int n, void* A, long long lda, void* tau, void* stream);
/* Blocked forms (compact WY, panels of 32 columns, LAPACK zgeqrf / zlarft / zlarfb): the panel by the column kernels,
 * everything to its right by three ZGEMMs on the FP64 tensor cores.  Same output layout as qm_qr / qm_qr_formq.
 * work: qm_qr_work_bytes(m, n) bytes of device scratch. */
long long qm_qr_work_bytes(int m, int n);
int qm_qr_blocked(int m, int n, void* A, long long lda, void* tau, void* work, long long work_bytes, void* stream);
int qm_qr_formq_blocked(int m, int k, const void* A, long long lda, const void* tau, void* Q, long long ldq, void* work,
                        long long work_bytes, void* stream);
int qm_qr_formq(int m, int k, const void* A, long long lda, const void* tau, void* Q, long long ldq, void* stream);
int qm_qr_finish(int m, int n, const void* A, long long lda, void* R, long long ldr, void* Q, long long ldq,
                 void* stream);

/* ---- MPS bookkeeping ----------------------------------------------------------- */

/* Rank selection of quimb _trim_and_renorm_svd_result: mode 0 'rel' (compress: mps.py:247, :451-453, :881),
 * 1 'rsum2' (+renorm; from_dense mps.py:242, gate_split_ mps.py:928-931, :968-971).
 * out_rank: int[1], out_f: double[1] (renormalisation factor). */
int qm_trim(const void* S, int k, double cutoff, int mode, int max_bond, void* out_rank, void* out_f, void* stream);

/* out = in with rows (mode 1) or columns (mode 2) scaled by (S*f)^(half_power ? 1/2 : 1);
 * mode 0 copies.  absorb='both' (mps.py:242, :968-971) / 'left' (mps.py:451-453, :881) of tensor_split.
 * f may be NULL. */
int qm_scale_copy(void* out, long long ldo, const void* in, long long ldi, int rows, int cols, const void* S,
                  const void* f, int mode, int half_power, void* stream);

/* theta[(l,oi),(oj,r)] = sum M[(oi,oj),(pi,pj)] X[(l,pi),(pj,r)], M = G or G^H, in place
 * on the (2l x 2r) matrix X.  gate_split_ contraction, mps.py:928-931, :968-971. */
int qm_theta_gate(void* X, int l, int r, const void* G, int dagger, void* stream);

/* B[l,o,r] = sum_p M[o,p] B[l,p,r] in place.  gate_(contract=True), mps.py:913-917, :953-957. */
int qm_site_gate(void* B, int l, int r, const void* G, int dagger, void* stream);

/* chi=2 truncation bookkeeping for one bond (mps.py:881): picks n <= 2 by the 'rel'
 * cutoff, applies the canonical row-phase rule, writes the site tensor rows Csite[2][4],
 * the projector Vsel[4][2] and bond[0] = n.  squared = 1: S holds eigenvalues of T^H L T (squared
 * singular values); squared = 2: Vh points at the 4x4 Hermitian matrix T^H L T itself and the kernel
 * diagonalises it (S unused); then ambiguous[0] is set to 1 when s_1 <= ambiguous_rel * s_0, i.e. when the
 * squared formulation cannot resolve the rank decision / second vector and the caller must redo
 * the layer with the QR-based path. */
int qm_chi2_select(const void* S, const void* Vh, long long ldvh, double cutoff, double tie, void* Csite,
                   void* Vsel, void* bond, int squared, double ambiguous_rel, void* ambiguous, void* stream);
int qm_chi2_first(const void* T0, void* Csite, void* stream);

/* Isometry -> unitary completion for all sites of a chi=2 MPS
 * (_generate_{first,two,last}_site_unitary + generate_unitary_layer, mps.py:565-847).
 * C: [N][8] padded site tensors, bond: int[N-1]; gates: [N][16], kinds: int[N]
 * (2 = two-qubit gate on (i,i+1), 1 = one-qubit gate), bad: int[1] unitarity flag. */
int qm_complete_unitaries(const void* C, const void* bond, int n_sites, void* gates, void* kinds, void* bad,
                          double sign_tol, void* stream);

/* ---- fused bookkeeping for small registers (bonds <= 64; one kernel instead of a group of launches: a batch of
 * small states replayed from CUDA graphs is bound by the number of kernel nodes) -------------------------------- */

/* qm_trim + qm_expect_ints + 2 x qm_scale_copy: rank by the cutoff (mode 0 'rel' + max_bond, singular values absorbed
 * to the left: left = U S, right = Vh; mode 1 'rsum2' + Frobenius renormalisation, sqrt absorbed on both sides),
 * outputs in the shapes of the ASSUMED rank expect_rank (left: m x expect, right: expect x n, contiguous);
 * mismatch[0] = 1 if the data give another rank.  quimb _trim_and_renorm_svd_result + absorb behind mps.py:242,
 * :451-453, :968-971. */
int qm_split_absorb(const void* U, long long ldu, const void* S, const void* Vh, long long ldvh, int m, int n, int k,
                    double cutoff, int mode, int max_bond, int expect_rank, void* left, void* right, void* mismatch,
                    void* stream);

/* qm_zgemm + qm_theta_gate: X (2l x 2r) = M (A A2), A: (l,2,b), A2: (b,2,r), M = G or G^H (mps.py:968-971). */
int qm_theta_small(const void* A, const void* A2, int l, int b, int r, const void* G, int dagger, void* X, void* stream);

/* chi=2 truncation (mps.py:881), fast path through left environments: qm_chi2_env: Lout (r x r) =
 * sum_p B[:,p,:]^H Lprev B[:,p,:] (Lprev NULL = identity; B: (l,2,r)); qm_chi2_bond: one bond -- M = L T, H = T^H M,
 * rank <= 2 selection with the canonical phase rule (as qm_chi2_select, squared = 2), Tout (l0 x 4) = Bprev (T Vsel). */
int qm_chi2_env(const void* Lprev, const void* B, int l, int r, void* Lout, void* stream);
int qm_chi2_bond(const void* L, int b, const void* T, const void* Bprev, int l0, double cutoff, double tie,
                 double ambiguous_rel, void* Csite, void* bond, void* ambiguous, void* Tout, void* stream);

/* <0..0|psi> as the product of the p = 0 slices of all sites in one launch (mps.py:1020-1039).  sites: HOST array of
 * device pointers to the (l,2,r) tensors; dims: HOST int[n_sites+1] bond sizes; out: complex[1]; with tol >= 0 the
 * early-break test |f - 1| <= tol (sequential.py:390) must NOT fire, else mismatch[0] = 1. */
int qm_zero_overlap(const void* const* sites, const int* dims, int n_sites, double tol, void* out, void* mismatch,
                    void* stream);

/* out (r,2,l) = in (l,2,r) with the bond axes swapped: mirror image of a site tensor, used
 * for the left-handed canonicalize/compress variants (mps.py:396, :451-453 with mode="left"). */
int qm_reverse3(void* out, const void* in, int l, int r, void* stream);

/* Speculative static-shape execution (graphs.py): ranks / block structure / early break are assumed
 * and validated on the device; mismatch[0] = 1 sends the state back through the eager path.
 * qm_expect_ints: vals[i] must equal expect[i] (device array) or `scalar` when expect is NULL.
 * qm_expect_not_close: the early-break test |f - 1| <= tol (sequential.py:390) must not fire. */
int qm_expect_ints(const void* vals, const void* expect, int n, int scalar, void* mismatch, void* stream);

/* Batch forms of the small-register kernels: `batch` same-shape problems in ONE launch (the lock-step lanes of the
 * CUDA-graph batch path advance W states together: one graph node per step instead of W; no counterpart in the
 * reference, which prepares one state per call -- sequential.py:509-541 is run once per state).  `strides`: HOST array
 * with the element stride between consecutive problems of every pointer argument, in argument order; scalar stride
 * arguments likewise.  mismatch / bad / ambiguous: int vectors indexed by the problem. */
int qm_split_absorb_batch(const void* U, long long ldu, const void* S, const void* Vh, long long ldvh, int m, int n,
                          int k, double cutoff, int mode, int max_bond, int expect_rank, void* left, void* right,
                          void* mismatch, int batch, const long long* strides /* U S Vh left right */, void* stream);
int qm_theta_small_batch(const void* A, const void* A2, int l, int b, int r, const void* G, int dagger, void* X,
                         int batch, const long long* strides /* A A2 G X */, void* stream);
int qm_chi2_env_batch(const void* Lprev, const void* B, int l, int r, void* Lout, int batch,
                      const long long* strides /* Lprev B Lout */, void* stream);
int qm_chi2_bond_batch(const void* L, int b, const void* T, const void* Bprev, int l0, double cutoff, double tie,
                       double ambiguous_rel, void* Csite, void* bond, void* ambiguous, void* Tout, int batch,
                       const long long* strides /* L T Bprev Csite bond ambiguous Tout */, void* stream);
int qm_zero_overlap_batch(const void* const* sites, const int* dims, int n_sites, double tol, void* out, void* mismatch,
                          int batch, const long long* strides /* one per site tensor */, void* stream);
int qm_site_gate_batch(void* B, int l, int r, const void* G, int dagger, int batch, long long strideB, long long strideG,
                       void* stream);
int qm_chi2_first_batch(const void* T0, void* Csite, int batch, long long strideT, long long strideC, void* stream);
/* C [batch][n_sites][8], gates [batch][n_sites][16], kinds [batch][n_sites], bad [batch]: dense per problem */
int qm_complete_unitaries_batch(const void* C, const void* bond, int n_sites, void* gates, void* kinds, void* bad,
                                double sign_tol, int batch, long long stride_bond, void* stream);
int qm_expect_ints_batch(const void* vals, const void* expect, int n, int scalar, void* mismatch, int batch,
                         long long stride_vals, void* stream);
int qm_expect_not_close(const void* f, double tol, void* mismatch, void* stream);

/* ---- vectors ------------------------------------------------------------------- */
int qm_conj_scale_copy(void* out, const void* in, long long n, int conj, double scale, void* stream);
/* out[0..1] = sum conj(a_i) b_i (re, im).  Reproducible: per-CTA partials are kept in out[2..] and added in
 * fixed order by a second kernel, so `out` must hold qm_vdot_out_doubles() doubles.
 * mps.norm() / mps.normalize() (mps.py:285, :302-310), the Ket normalisation (base.py:96-97) and the final
 * <psi|circuit> (README.md:66); qm_div_sqrt: x <- x / sqrt(nrm2[0]). */
int qm_vdot_out_doubles(void);
int qm_vdot(const void* a, const void* b, long long n, void* out /* double[qm_vdot_out_doubles()] */, void* stream);
int qm_div_sqrt(void* x, long long n, const void* nrm2 /* double[1] */, void* stream);

/* ---- dense statevector path (optimisation sweeps) -------------------------------- */

/* x <- op(G) x on site (kind 1) or sites (site,site+1) (kind 2); op 0 G, 1 G^H, 2 G^T.
 * quimb `TensorNetwork @ Tensor`, sequential.py:460, :496. */
int qm_apply_gate(void* x, int n_sites, int site, int kind, const void* G, int op, void* stream);

/* c <- gates applied in order to |0..0>.  sites/kinds: HOST int arrays.
 * sequential.py:215-292 + to_dense :443-447. */
int qm_circuit_state(void* c, int n_sites, const void* gates, const int* sites, const int* kinds, int n_gates,
                     void* stream);

/* One environment sweep, gates updated in place (sequential.py:452-505).
 * work: qm_sweep_work_bytes() bytes.  envs: optional [n_gates][16] record of each E. */
long long qm_sweep_work_bytes(void);
int qm_sweep(void* c, void* tbar, int n_sites, void* gates, const int* sites, const int* kinds, int n_gates,
             void* work, void* envs, void* stream);

/* Stored-intermediates variant: qm_circuit_states keeps every c_k = g_{k-1}..g_0|0> in HBM
 * (cs: (n_gates+1) * 2^N amplitudes) and qm_sweep_stored runs the same sweep with one fused
 * pass per gate (tbar update by the previous new gate + environment reduction).  vwarm (optional,
 * [n_gates][16], zero-initialised by the caller before the first sweep) carries each gate's right
 * singular vectors from one sweep to the next as the starting point of the 4x4 Jacobi polar. */
int qm_circuit_states(void* cs, int n_sites, const void* gates, const int* sites, const int* kinds, int n_gates,
                      void* stream);
int qm_sweep_stored(const void* cs, void* tbar, int n_sites, void* gates, const int* sites, const int* kinds,
                    int n_gates, void* work, void* envs, void* vwarm, void* stream);

/* Small registers (n_sites <= 12, n_gates <= 256): ALL `num_sweeps` optimisation sweeps of `batch`
 * independent states in one launch, one CTA per state with both dense vectors and the gates in
 * shared memory (same arithmetic as qm_circuit_state + qm_sweep per sweep; sequential.py:443-505,
 * 509-541).  targets: [batch][2^N] (not conjugated); gates: [batch][n_gates][16], updated in place;
 * sites_dev / kinds_dev: DEVICE int[n_gates], one schedule for the batch; envs: optional
 * [batch][n_gates][16], environments of the last sweep.  psis / overlaps (optional): overlaps[batch][2] (double)
 * receives <psi_b| circuit_b |0..0> / |psi_b| for the final gates (README.md:66), psi_b = psis[b] ([batch][2^N]) or the
 * target when psis is NULL; num_sweeps = 0 is then allowed.  Returns -3 if the state does not fit. */
int qm_sweeps_small(const void* targets, int n_sites, void* gates, const int* sites_dev, const int* kinds_dev,
                    int n_gates, int num_sweeps, int batch, void* envs, const void* psis, void* overlaps, void* stream);

/* Large registers: ALL `num_sweeps` optimisation sweeps (forward circuit states + backward environment
 * sweep, sequential.py:509-541, :443-505) in one persistent cooperative launch, one CTA per SM and one grid
 * barrier per gate-step; every CTA reduces the partial environments and updates the gate redundantly (bit-identical),
 * so the new gate needs no broadcast.  cs: (n_gates+1) * 2^N amplitudes of scratch (stored circuit states); tbar:
 * 2^N scratch; target: dense target (not conjugated); gates: [n_gates][16], updated in place; sites_dev /
 * kinds_dev: DEVICE int[n_gates]; work: qm_sweeps_persist_work_bytes(n_gates) bytes; envs: optional
 * [n_gates][16], environments of the last sweep.  Returns -3 if the grid cannot be co-scheduled. */
long long qm_sweeps_persist_work_bytes(int n_gates);
int qm_sweeps_persist(void* cs, void* tbar, const void* target, int n_sites, void* gates, const int* sites_dev,
                      const int* kinds_dev, int n_gates, int num_sweeps, void* work, void* envs, void* stream);

/* Library identification: returns the compiled architecture number (100 for sm_100a). */
int qm_version(void);

/* Programmatic dependent launch of the dependent kernel chains (default on; env QM_PDL=0 turns it off).
 * Returns the previous setting. */
int qm_set_pdl(int on);

/* Kernel launches issued by this library since load (every launch is counted). */
long long qm_launch_count(void);

/* Optional per-kernel-class CUDA-event profiler (bench.py roofline leg): between begin
 * and end every launch is bracketed by events on its stream.  total_ms / count: host
 * arrays of qm_prof_num_classes() entries. */
int qm_prof_num_classes(void);
const char* qm_prof_class_name(int cls);
int qm_prof_begin(void);
int qm_prof_end(double* total_ms, long long* count);
/* algorithmic work per class since qm_prof_begin: flops (zgemm, svd_*, qr_apply) or bytes (gate, env_polar) */
int qm_prof_work_get(double* work);

#ifdef __cplusplus
}
#endif
#endif /* QMPRS_B200_H */
