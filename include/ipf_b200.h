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
 * Multi-GPU: one process (ctx) per GPU; the only exchange step is the sum of the
 * Gram blocks between ipf_solve_gram and ipf_solve_factor, done by the host
 * (torch.distributed/NCCL all-reduce on the device pointer the library exposes).
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

#define IPF_MAX_VARS 10      /* (N-1) + (N-1)(N-2)/2 for N = 5 */
#define IPF_MAX_DEGREE 20
#define IPF_MAX_ATOMS 4096   /* per configuration */

typedef struct ipf_ctx ipf_ctx;
typedef struct ipf_basis ipf_basis;
typedef struct ipf_matrix ipf_matrix;
typedef struct ipf_solve ipf_solve;

/* One run of basis functions sharing body-order and descriptor.
 * Monomial m belongs to local basis function mono_b[m] (non-decreasing) and has
 * exponents mono_exp[m*IPF_MAX_VARS + v] over the variables
 * (x_1..x_{N-1}, w_12, w_13, .., w_23, ..); a basis function is the sum of its
 * monomials (an S_{N-1} orbit).  transform 0: x = exp(-a (r/r0 - 1)); 1: x = (r0/r)^a.
 * cutoff: CosCut(rc1, rc2).  Replaces the NBodyIPs objects built by
 * nbpolys(N, BondAngleDesc(transform, CosCut(rc1, rc2)), deg)   README.md:34-47. */
typedef struct {
    int32_t N;
    int32_t nb;
    int32_t nmono;
    int32_t transform;
    double a, r0, rc1, rc2;
    const int32_t* mono_b;
    const uint8_t* mono_exp;
} ipf_block_spec;

/* ---- context ---------------------------------------------------------- */
int32_t ipf_init(int32_t device, ipf_ctx** ctx);
int32_t ipf_destroy(ipf_ctx* ctx);
const char* ipf_last_error(ipf_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t ipf_launch_count(ipf_ctx* ctx);
/* the CUDA stream all work of this ctx is issued on (a cudaStream_t) */
void* ipf_stream(ipf_ctx* ctx);

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
 *   pbc[3*ncfg];  Z[natoms] (atomic numbers; currently informational)
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

#ifdef __cplusplus
}
#endif
#endif /* IPF_B200_H */
