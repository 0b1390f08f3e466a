/* ipf_b200.h -- C-ABI of the B200-native LsqDB assembly + lsqfit solve.
 *
 * Drop-in boundary for IPFitting.jl's least-squares hot path.
 * Plain pointers and sizes only; no C++ exceptions cross this boundary; every
 * entry point returns an int32 status (0 = IPF_OK, negative = error) and the
 * message of the last error is available from ipf_last_error().
 *
 * The Julia shim (julia/IPFittingB200.jl) replaces exactly two regions of the
 * reference with ccalls into this library:
 *   (1) the threaded assembly loop   src/lsq_db.jl:181-185
 *         tfor_observations(configs, (n,okey,cfg,lck) -> safe_append!(db,lck,cfg,okey))
 *       i.e. eval_obs/vec_obs/block placement   src/lsq_db.jl:268-278,
 *       src/datatypes.jl:28,40,65 (energy/forces/virial of an IPBasis)
 *       ->  ipf_assemble
 *   (2) weighting, column scaling and the QR solve
 *         src/lsq.jl:306-307 (Y = W.*Y; lmul!(Diagonal(W), Psi)),
 *         src/lsq.jl:450-473 (Psi*pinv(P); qr!; cond(R); c = qr \ Y; rel_rms),
 *         src/lsq.jl:593     (c = D_inv * c)
 *       ->  ipf_solve_begin / ipf_solve_gram / ipf_solve_factor / ipf_solve_finish
 *   plus the error table of src/lsqerrors.jl:225-266 evaluated from Psi*c
 *       ->  ipf_residual_stats
 *
 * Ownership: the caller owns every host pointer; the library never retains a
 * host pointer past return.  Device objects are opaque handles with explicit
 * destroy calls.  Threading: one ipf_ctx = one GPU = one host thread at a time
 * (entry points are not re-entrant on the same ctx; distinct ctxs are independent).
 * Multi-GPU: one ctx per GPU (one process per GPU, or several ctxs in one process, each driven by its own
 * host thread); the only exchange step is the sum of the Gram blocks of a CholeskyQR pass (plus the n-vector A^T r of a
 * refinement step and two residual scalars in ipf_solve_finish).  With a communicator attached (ipf_comm_init)
 * ipf_solve_gram / ipf_solve_finish perform it themselves: ncclAllReduce on the ctx stream, no host synchronisation.
 * Without one the caller may sum the exposed Gram buffer by any other means, after ipf_solve_set_csne(S, 0), and must sum
 * the residual scalars of ipf_solve_finish itself.
 */
#ifndef IPF_B200_H
#define IPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPF_OK 0
#define IPF_EINVAL (-1)      /* bad argument */
#define IPF_ECUDA (-2)       /* CUDA runtime error (see ipf_last_error) */
#define IPF_ENOMEM (-3)      /* device allocation failed */
#define IPF_ENAN_Y (-4)      /* "NaN detected in observations vector"   src/lsq.jl:261 */
#define IPF_ENAN_W (-5)      /* "NaN detected in weights vector"        src/lsq.jl:262 */
#define IPF_ENAN_PSI (-6)    /* "discovered NaNs in LSQ system matrix"  src/lsq.jl:285-292 */
#define IPF_ESINGULAR (-7)   /* factorisation broke down (rank deficient beyond repair) */
#define IPF_ESTATE (-8)      /* entry points called out of order */
#define IPF_ENCCL (-9)       /* NCCL error / libnccl.so.2 not loadable (see ipf_last_error) */

#define IPF_MAX_VARS 10      /* (N-1) + (N-1)(N-2)/2 for N = 5 */
#define IPF_MAX_DEGREE 20
#define IPF_MAX_ATOMS 4096   /* per configuration */
#define IPF_MAX_SOLVE_COLS 3072    /* columns of a system handed to ipf_solve* (shared-memory bound of the n x n factor kernels) */
#define IPF_MAX_PREDICT_COLS 6144  /* columns of a matrix handed to ipf_predict / ipf_residual_stats */

typedef struct ipf_ctx ipf_ctx;
typedef struct ipf_basis ipf_basis;
typedef struct ipf_matrix ipf_matrix;
typedef struct ipf_solver ipf_solver;
typedef struct ipf_configs ipf_configs;

/* One run of basis functions sharing body-order and descriptor.
 * Monomial m belongs to local basis function mono_b[m] (non-decreasing) and has
 * exponents mono_exp[m*IPF_MAX_VARS + v] over the variables
 * (x_1..x_{N-1}, w_12, w_13, .., w_23, ..); a basis function is the sum of its
 * monomials (an S_{N-1} orbit).  transform 0: x = exp(-a (r/r0 - 1)); 1: x = (r0/r)^a.
 * cutoff: CosCut(rc1, rc2).  Replaces the NBodyIPs objects built by
 * nbpolys(N, BondAngleDesc(transform, CosCut(rc1, rc2)), deg)   README.md:34-47.
 * Species (JuLIP Atoms.Z, src/prototypes.jl:19-25): zc > 0 makes the run species-resolved: it acts on the clusters
 * whose centre atom has atomic number zc and whose N-1 neighbours have exactly the atomic numbers zs[0] <= .. <= zs[N-2]
 * (legs assigned to the variable slots in that order; the monomial sums need only be invariant under exchanges of
 * equal-species slots) and is zero on every other cluster.  zc = 0: species-agnostic, every cluster (zs ignored).
 * Species-resolved runs are evaluated by the table-driven kernel.
 * Environment (NBodyIPs EnvIP; the `env` of the README's W5Bdeg12env database, README.md:65): env_t > 0 multiplies the
 * site function by n_i^env_t, n_i = sum_j fenv(r_ij) over the neighbours of the centre atom, fenv = CosCut(env_rc1,
 * env_rc2) with env_rc2 <= rc2; the forces and virials include the derivative of n_i.  env_t = 0: plain function.
 * Environment-dependent runs are evaluated by the table-driven kernel. */
typedef struct {
    int32_t N;
    int32_t nb;
    int32_t nmono;
    int32_t transform;
    double a, r0, rc1, rc2;
    const int32_t* mono_b;
    const uint8_t* mono_exp;
    int32_t zc;
    int32_t zs[4];
    int32_t env_t;
    double env_rc1, env_rc2;
} ipf_block_spec;

/* ---- context ---------------------------------------------------------- */
int32_t ipf_init(int32_t device, ipf_ctx** ctx);
int32_t ipf_destroy(ipf_ctx* ctx);
const char* ipf_last_error(ipf_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t ipf_launch_count(ipf_ctx* ctx);
/* the CUDA stream all work of this ctx is issued on (a cudaStream_t) */
void* ipf_stream(ipf_ctx* ctx);
/* per-kernel device timing (CUDA events on the ctx stream), used by bench.py for the
 * roofline numbers.  Names: "neighbour_list", "pair_record", "site_deriv_2b|3b|4b|5b",
 * "k4_build", "gram", "cholesky", "trsm", "residual". */
/* kernel-path selection for tests and profiling.  on = 0: automatic (default); 1: only the table-driven kernels;
 * 2: specialised kernels, but the 3-body block always takes the HBM-scratch + gather path (basis_3b.cu) instead of the
 * configuration-resident kernel (basis_3b_cfg.cu) */
int32_t ipf_set_force_generic(ipf_ctx* ctx, int32_t on);
/* on = 1: ipf_assemble / ipf_assemble_configs return as soon as the last kernels are queued on the ctx stream
 * (no final synchronisation); every later call on this ctx is ordered behind them, and a device fault is
 * reported by the next synchronising call.  Lets host-side work of the caller (collect_observations, ...) overlap
 * the tail of the assembly.  Default 0: synchronous, errors reported by the call itself. */
int32_t ipf_set_async(ipf_ctx* ctx, int32_t on);
int32_t ipf_profile_enable(ipf_ctx* ctx, int32_t on);
int32_t ipf_profile_reset(ipf_ctx* ctx);
int32_t ipf_profile_get(ipf_ctx* ctx, const char* name, double* ms, int64_t* count);

/* ---- multi-GPU communicator (comm.cu) -----------------------------------
 * The reference has no distributed path (only a commented-out DistributedArrays stub, src/lsq_db.jl:202-228); this is
 * the exchange step SURVEY.md section 8(e) defines.  Rank 0 calls ipf_comm_unique_id and hands the 128 bytes to the
 * other ranks by any transport (Julia: Distributed / MPI / a file); every rank then calls ipf_comm_init (collective).
 * From then on, on this ctx:
 *   ipf_solve_gram    returns the Gram block SUMMED over the ranks (identical bits on every rank),
 *   ipf_solve_finish  returns ssr, ssy summed over the ranks,
 *   ipf_comm_allreduce_f64 sums a small host vector (error-table sums) over the ranks.
 * NCCL is loaded at run time (libnccl.so.2); IPF_ENCCL if it is missing. */
#define IPF_UNIQUE_ID_BYTES 128
int32_t ipf_comm_unique_id(ipf_ctx* ctx, uint8_t* id /* [IPF_UNIQUE_ID_BYTES] */);
int32_t ipf_comm_init(ipf_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id);
int32_t ipf_comm_destroy(ipf_ctx* ctx);
/* paused = 1: the solve stages on this ctx stay rank-local (no exchange inside ipf_solve_gram / finish / residual) until
 * paused = 0; ipf_comm_allreduce_f64 always exchanges.  Used by the TSQR variant: every rank factors its own rows, the
 * n x (n+1) blocks [R_p | Q_p^T y] are exchanged once and the stacked system is factored again (src/lsq.jl:468). */
int32_t ipf_comm_pause(ipf_ctx* ctx, int32_t paused);
int32_t ipf_comm_size(const ipf_ctx* ctx);
int32_t ipf_comm_rank(const ipf_ctx* ctx);
int32_t ipf_comm_allreduce_f64(ipf_ctx* ctx, double* x, int64_t n);

/* ---- basis ------------------------------------------------------------ */
int32_t ipf_basis_create(ipf_ctx* ctx, int32_t nblocks, const ipf_block_spec* blocks,
                         ipf_basis** basis);
int32_t ipf_basis_destroy(ipf_ctx* ctx, ipf_basis* basis);
int64_t ipf_basis_length(const ipf_basis* basis);

/* ---- device-resident column-major matrix (db.Psi, src/prototypes.jl:82-87) ---- */
int32_t ipf_matrix_create(ipf_ctx* ctx, int64_t nrows, int64_t ncols, ipf_matrix** M); /* zero filled, src/lsq_db.jl:249 */
int32_t ipf_matrix_destroy(ipf_ctx* ctx, ipf_matrix* M);
int32_t ipf_matrix_to_host(ipf_ctx* ctx, const ipf_matrix* M, double* host, int64_t ld);
int32_t ipf_matrix_from_host(ipf_ctx* ctx, ipf_matrix* M, const double* host, int64_t ld);
int64_t ipf_matrix_rows(const ipf_matrix* M);
int64_t ipf_matrix_cols(const ipf_matrix* M);
int64_t ipf_matrix_ld(const ipf_matrix* M);
void* ipf_matrix_device_ptr(const ipf_matrix* M);

/* ---- neighbour list (for the bit-exact tests; K1) ----------------------
 * All (i, j, S) with |X_j + S*C - X_i| < rcut, (i,j,S) != (i,i,0), in
 * lexicographic (i, j, Sx, Sy, Sz) order.  cell: row-major 3x3, rows = lattice
 * vectors.  Writes at most cap entries; *n receives the total.  Replaces
 * JuLIP.neighbourlist(at, rcut) as used inside energy/forces/virial
 * (src/datatypes.jl:28,40,65) and src/preconditioners.jl:56. */
int32_t ipf_neighbours(ipf_ctx* ctx, int32_t nat, const double* pos, const double* cell,
                       const uint8_t* pbc, double rcut, int64_t cap, int64_t* n,
                       int32_t* I, int32_t* J, int32_t* S /* [3*cap] */, double* R /* [3*cap] or NULL */);

/* ---- assembly (K1-K3) --------------------------------------------------
 * Fills Psi[rows of every observation of every config, :] with
 * vec_obs(okey, eval_obs(okey, basis, cfg))  (src/lsq_db.jl:268-278).
 *   atom_off[ncfg+1]  prefix offsets into pos/Z
 *   pos[3*natoms]     atom-major xyz;  cell[9*ncfg] row-major, rows = lattice vectors
 *   pbc[3*ncfg];  Z[natoms] (atomic numbers; read by species-resolved blocks, may be NULL otherwise)
 *   rowE/rowF/rowV[ncfg]: 1-based first row of the E / F / V block of each config,
 *       0 = observation absent (rows come from _alloc_lsq_matrix, src/lsq_db.jl:237-250,
 *       so any Dict key order is honoured).  Block lengths are 1 / 3*nat / 6.
 * Psi must have ipf_basis_length(basis) columns.  The E, F and V blocks of a
 * config are produced by ONE fused evaluation (the reference evaluates them as
 * three separate tasks, src/obsiter.jl:87-92). */
int32_t ipf_assemble(ipf_ctx* ctx, const ipf_basis* basis, int64_t ncfg,
                     const int64_t* atom_off, const double* pos, const double* cell,
                     const uint8_t* pbc, const int32_t* Z, const int64_t* rowE,
                     const int64_t* rowF, const int64_t* rowV, ipf_matrix* Psi);

/* The same in two steps, with the configurations kept resident on the device
 * (repeated assembly, benchmarking with inputs already in HBM). */
int32_t ipf_configs_create(ipf_ctx* ctx, int64_t ncfg, const int64_t* atom_off, const double* pos,
                           const double* cell, const uint8_t* pbc, const int32_t* Z,
                           const int64_t* rowE, const int64_t* rowF, const int64_t* rowV,
                           ipf_configs** cfgs);
int32_t ipf_configs_destroy(ipf_ctx* ctx, ipf_configs* cfgs);
int32_t ipf_assemble_configs(ipf_ctx* ctx, const ipf_basis* basis, const ipf_configs* cfgs,
                             ipf_matrix* Psi);

/* ---- solve (K4-K7) -----------------------------------------------------
 * Weighted, column-scaled least squares on the device-resident matrix:
 *     minimise | diag(W[I]) (Psi[I, J] diag(Dinv) c' - Y[I]) |,   c = Dinv .* c'
 *   Y, W [nrows of Psi]: observations (Vref already subtracted) and row weights
 *       from collect_observations (src/lsq.jl:74-109); checked for NaN (:261-262).
 *   Irows [nIrows] 0-based row selection or NULL = all (the host builds it as
 *       (W != 0) & (Icfg in Itrain), src/lsq.jl:270-283);  Icols likewise (Ibasis).
 *   Dinv [nIcols] diagonal of pinv(P) (src/lsq.jl:450-458) or NULL = identity.
 * Psi is never modified (the reference copies too, src/lsq.jl:300).
 *
 * Staged form (multi-GPU: one ctx per GPU holding its own rows):
 *     ipf_solve_begin
 *     repeat { ipf_solve_gram  -> device pointer to the local (n+1)x(n+1) Gram of [A | y]
 *              [host: all-reduce(sum) that buffer across ranks]
 *              ipf_solve_factor -> done? }
 *     ipf_solve_finish -> c, R, local residual sums
 * ipf_solve is the single-GPU one-shot.  rel_rms = |A c' - y| / |y| on the
 * weighted, scaled system (src/lsq.jl:473). */
int32_t ipf_solve_begin(ipf_ctx* ctx, const ipf_matrix* Psi, const double* Y, const double* W,
                        const int64_t* Irows, int64_t nIrows, const int64_t* Icols, int64_t nIcols,
                        const double* Dinv, ipf_solver** S);
/* as ipf_solve_begin, every vector argument being a DEVICE pointer */
int32_t ipf_solve_begin_dev(ipf_ctx* ctx, const ipf_matrix* Psi, const double* dY, const double* dW,
                            const int64_t* dIrows, int64_t nIrows, const int64_t* dIcols,
                            int64_t nIcols, const double* dDinv, ipf_solver** S);
/* Force preconditioner (_forceprecon, src/lsq.jl:184-234; P from e.g. AdjacancyPrecon,
 * src/preconditioners.jl:54-74).  Block b replaces the rows [row0[b], row0[b]+m[b]) of the weighted system
 * (positions in the SELECTED row numbering, all columns and y) by rtP_b \ rows, where
 *   solver 0 ("chol"): P holds P_b, rtP_b = cholesky(P_b).L          (src/lsq.jl:215-221)
 *   solver 1 ("sqrt"): P holds sqrt(P_b) (symmetric), applied as its inverse   (src/lsq.jl:222-223)
 * P: packed column-major m[b] x m[b] blocks at P[Poff[b]], Plen doubles in total.  The scalar row weight of a
 * force block commutes with rtP \ ., so this is called after ipf_solve_begin and before the first ipf_solve_gram. */
int32_t ipf_solve_precon(ipf_ctx* ctx, ipf_solver* S, int64_t nblocks, const int64_t* row0, const int32_t* m,
                         const int64_t* Poff, const double* P, int64_t Plen, int32_t solver);
int32_t ipf_solve_shape(const ipf_solver* S, int64_t* M, int64_t* n);
/* the weighted system itself (get_lsq_system, src/lsq.jl:245-311): Aw [M x n] column-major, yw [M] */
int32_t ipf_solve_system_to_host(ipf_ctx* ctx, const ipf_solver* S, double* Aw, int64_t ld, double* yw);
/* G_dev: device pointer to the (n+1) x (n+1) Gram block of [A | y] (summed over the ranks when a communicator is
 * attached; the call returns without waiting for the device) */
int32_t ipf_solve_gram(ipf_ctx* ctx, ipf_solver* S, void** G_dev, int64_t* n1);
int32_t ipf_solve_factor(ipf_ctx* ctx, ipf_solver* S, int32_t* done);
/* c [n]; R_out [n*n] column-major upper factor of the scaled system, or NULL;
 * ssr, ssy: sums |y - A c'|^2 and |y|^2 over the rows of this ctx -- over ALL ranks when a communicator is attached
 * (otherwise sum them across ranks before taking the ratio) */
int32_t ipf_solve_finish(ipf_ctx* ctx, ipf_solver* S, double* c, double* R_out, double* ssr, double* ssy);
/* Regulariser rows (`_regularise!`, src/lsq.jl:163-181): append nextra rows [P | q] to the weighted system (after
 * the row weighting, before the first ipf_solve_gram).  P: column-major nextra x n (ldP), columns = the selected
 * basis functions, unscaled (the solver applies Dinv); q: nextra targets or NULL (zeros).  In a multi-GPU solve
 * exactly one rank appends them. */
int32_t ipf_solve_append_rows(ipf_ctx* ctx, ipf_solver* S, int64_t nextra, const double* P, int64_t ldP, const double* q);
/* After the factorisation is complete (done = 1): qty[n] = Q^T y for the factor R that ipf_solve_finish returns
 * (A0 = Q R), and the residual sums of ANY coefficient vector c[n] given in the scaled basis (before D_inv):
 * ssr = |y - A0 c|^2, ssy = |y|^2 (local rows).  They carry the `:svd` and `:rrqr` branches of lsqfit
 * (src/lsq.jl:482-497), whose small dense work (SVD / pivoted QR of the n x n factor) stays on the host. */
int32_t ipf_solve_qty(ipf_ctx* ctx, ipf_solver* S, double* qty);
int32_t ipf_solve_residual(ipf_ctx* ctx, ipf_solver* S, const double* c, double* ssr, double* ssy);
/* on = 0: never finish on the corrected semi-normal equations (whose refinement steps need the sum of A^T r over all
 * ranks).  REQUIRED when the ranks' Gram blocks are summed by the caller instead of an attached communicator: the
 * library then only sees local rows in ipf_solve_finish.  Default on. */
int32_t ipf_solve_set_csne(ipf_solver* S, int32_t on);
/* cond(R) (src/lsq.jl:469) of the factor ipf_solve_finish returns, estimated on the device by power / inverse iteration
 * (a lower bound, tight to a few per cent); call after ipf_solve_finish.  The exact value is an SVD of R_out on the host. */
int32_t ipf_solve_cond(ipf_ctx* ctx, ipf_solver* S, double* kappa);
/* Diagnostics of a solve: Gram passes, orthogonality defect |A^T A - I|_F seen by the last pass, largest diagonal shift
 * used, kappa_hat = (max / min diagonal of the equilibrated Cholesky factor)^2 of the last pass, and the number of
 * refinement steps of the corrected semi-normal equations (0: the last pass was orthonormal to 0.1). */
int32_t ipf_solve_info(const ipf_solver* S, int32_t* npasses, double* defect, double* shift,
                       double* kappa_hat, double* refine_steps);
int32_t ipf_solve_destroy(ipf_ctx* ctx, ipf_solver* S);
int32_t ipf_solve(ipf_ctx* ctx, const ipf_matrix* Psi, const double* Y, const double* W,
                  const int64_t* Irows, int64_t nIrows, const int64_t* Icols, int64_t nIcols,
                  const double* Dinv, double* c, double* R_out, double* rel_rms);

/* yhat [nrows] = Psi * c  (c has one entry per column of Psi).  The fit values
 * add_fits! stores (src/lsqerrors.jl:62-77) for a potential that is linear in
 * the basis, without re-evaluating it (README.md:114-116). */
int32_t ipf_predict(ipf_ctx* ctx, const ipf_matrix* Psi, const double* c, double* yhat);

/* Error table sums (src/lsqerrors.jl:225-266): for every row r with group[r] = g >= 0
 *   sse[g] += (werr[r] (Yraw[r] - (Psi c)[r] - Yoff[r]))^2,  ssy[g] += (werr[r] Yraw[r])^2, cnt[g] += 1
 * Yoff = prediction of Vref (may be NULL); werr = err_weighthook (src/datatypes.jl:30,67). */
int32_t ipf_residual_stats(ipf_ctx* ctx, const ipf_matrix* Psi, const double* c, const double* Yraw,
                           const double* Yoff, const double* werr, const int32_t* group,
                           int32_t ngroups, double* sse, double* ssy, int64_t* cnt);

#ifdef __cplusplus
}
#endif
#endif /* IPF_B200_H */
