// ipf_common.cuh -- shared internals of libipf_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ipf_b200.h"

// per-kernel device timing (CUDA events on the ctx stream), off by default
struct ProfEntry {
    double ms = 0.0;
    int64_t count = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// bump allocator for per-chunk device temporaries: no cudaMalloc/cudaFree in steady state
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, need = 0;
    std::vector<void*> spill;       // overflow allocations of the current cycle
};

struct ipf_ctx {
    Arena arena;                        // assembly temporaries (per chunk)
    Arena solve_arena;                  // workspaces of the (single) active solver
    Arena chunk_arena;                  // configuration chunks of the host-pointer ipf_assemble (transient)
    int active_solvers = 0;
    // pair-record cache key (see ipf_pair_records)
    int64_t nl_serial = 0;           // neighbour lists built so far (identity of a list for the record cache)
    int64_t rec_nl = -1; int64_t rec_npairs = -1; int rec_transform = -1, rec_maxdeg = -1, rec_rs = 0;
    double rec_a = 0, rec_r0 = 0, rec_rc1 = 0, rec_rc2 = 0; const double* rec_ptr = nullptr;
    double* matrix_cache = nullptr;     // one recycled design-matrix allocation (avoids cudaMalloc/cudaFree per LsqDB)
    size_t matrix_cache_bytes = 0;
    cudaStream_t stream2 = nullptr;     // second stream: gather of batch n overlaps the moments of batch n+1
    bool profiling = false;
    bool force_generic = false;
    bool async_assemble = false;    // ipf_assemble*: no final stream synchronisation
    bool force_scratch3b = false;   // 3-body: HBM-scratch path even when the configuration-resident kernel applies
    std::map<std::string, ProfEntry> prof;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;
    // reusable device scratch (grown on demand, freed with the ctx)
    static constexpr int kScratchSlots = 12;
    void* scratch[kScratchSlots] = {};
    size_t scratch_bytes[kScratchSlots] = {};
    // combine tables of the 4-body fast path, keyed by the BlockDev they were built for (basis_4b.cu)
    std::map<const void*, void*> tabs4b;
    // multi-GPU: NCCL communicator attached by ipf_comm_init (comm.cu); nranks == 1: no exchange
    int gram_tiles_n1 = -1;             // Gram tile table currently on the device (gram.cu)
    const void* gram_tiles_ptr = nullptr;
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    bool comm_paused = false;           // ipf_comm_pause: the solve stages run rank-local (first level of a TSQR solve)
};

struct ipf_matrix {
    int64_t nrows = 0, ncols = 0, ld = 0;
    double* d = nullptr;
    size_t bytes = 0;       // size of the allocation behind d
};

struct BlockDev {
    int N = 0, nb = 0, nmono = 0, transform = 0, maxdeg = 0;
    double a = 1, r0 = 1, rc1 = 0, rc2 = 0;
    int64_t col0 = 0;
    int32_t* mono_start = nullptr;   // device [nb+1]
    uint8_t* mono_exp = nullptr;     // device [nmono][IPF_MAX_VARS]
    std::vector<int32_t> h_mono_start;
    std::vector<uint8_t> h_mono_exp;
    // species-resolved block: centre species, sorted neighbour species, and for every variable slot s that is the first
    // of its species the exponent table with slot s moved to position 0 (the kernels pin the own leg to slot 0)
    int zc = 0, zs[4] = {0, 0, 0, 0};
    uint8_t* mono_exp_var[4] = {nullptr, nullptr, nullptr, nullptr};
    // environment-dependent block: site function times n_i^env_t, n_i = sum_j CosCut(env_rc1, env_rc2)(r_ij)
    int env_t = 0;
    double env_rc1 = 0.0, env_rc2 = 0.0;
};

struct ipf_basis {
    std::vector<BlockDev> blocks;
    int64_t ncols = 0;
};

int ipf_set_error(ipf_ctx* ctx, int code, const char* fmt, ...);

#define IPF_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ipf_set_error((ctx), e__ == cudaErrorMemoryAllocation ? IPF_ENOMEM : IPF_ECUDA, \
                                 "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define IPF_CHECK(call)                \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != IPF_OK) return rc__; \
    } while (0)

// RAII timing scope: records an event pair around the launches issued inside it
struct ProfScope {
    ipf_ctx* ctx;
    ProfEntry* e = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    ProfScope(ipf_ctx* c, const char* name, cudaStream_t stream = nullptr) : ctx(c), st(stream ? stream : c->stream) {
        if (!c->profiling) return;
        e = &c->prof[name];
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    ~ProfScope() {
        if (!e) return;
        cudaEventRecord(e1, st);
        e->pending.emplace_back(e0, e1);
        e->count++;
    }
};

// grow-only scratch slot
int ipf_scratch(ipf_ctx* ctx, int slot, size_t bytes, void** out);
// arena: alloc is stream-ordered-safe on ctx->stream; reset recycles everything handed out so far
int ipf_arena_alloc(ipf_ctx* ctx, size_t bytes, void** out);
int ipf_arena_reset(ipf_ctx* ctx);
int ipf_arena_alloc_in(ipf_ctx* ctx, Arena& a, size_t bytes, void** out);
int ipf_arena_reset_in(ipf_ctx* ctx, Arena& a);

template <class T>
struct ABuf {          // arena-backed array (never freed individually)
    T* p = nullptr;
    int alloc(ipf_ctx* ctx, size_t count) {
        void* q = nullptr;
        int rc = ipf_arena_alloc(ctx, sizeof(T) * (count ? count : 1), &q);
        p = (T*)q;
        return rc;
    }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t count) {
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        return cudaMalloc((void**)&p, sizeof(T) * (count ? count : 1));
    }
};

// device array that is either owned (cudaMalloc/cudaFree) or carved from an arena (never freed individually)
template <class T>
struct FlexBuf {
    T* p = nullptr;
    bool owned = false;
    ~FlexBuf() { if (p && owned) cudaFree(p); }
    FlexBuf() = default;
    FlexBuf(const FlexBuf&) = delete;
    FlexBuf& operator=(const FlexBuf&) = delete;
    int alloc(ipf_ctx* ctx, Arena* arena, size_t count) {
        if (p && owned) cudaFree(p);
        p = nullptr;
        if (arena) {
            owned = false;
            void* q = nullptr;
            int rc = ipf_arena_alloc_in(ctx, *arena, sizeof(T) * (count ? count : 1), &q);
            p = (T*)q;
            return rc;
        }
        owned = true;
        cudaError_t e = cudaMalloc((void**)&p, sizeof(T) * (count ? count : 1));
        if (e != cudaSuccess) { p = nullptr; return ipf_set_error(ctx, IPF_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
        return IPF_OK;
    }
};

// ---- neighbour list of a chunk of configurations (device resident, arena backed) -------
struct NeighList {
    int64_t serial = -1;      // unique per built list (ipf_ctx::nl_serial)
    int64_t nsites = 0;       // atoms in the chunk
    int64_t npairs = 0;
    int32_t maxpairs_cfg = 0; // max pairs of one config
    int32_t max_nbrs = 0;     // max neighbours of one site
    double rcut = 0;
    ABuf<int64_t> site_off;       // [nsites+1] pair offsets (chunk-global)
    ABuf<int32_t> pi;             // [npairs] centre atom, local index in its config
    ABuf<int32_t> pj;             // [npairs] neighbour atom, local index in its config
    ABuf<int32_t> psite;          // [npairs] centre atom, chunk-global site index
    ABuf<int64_t> pbeg;           // [npairs] first pair of the centre atom's list (= site_off[psite])
    ABuf<int32_t> pcnt;           // [npairs] length of that list
    ABuf<int32_t> pS;             // [npairs*3] image shift
    ABuf<double> pR;              // [npairs*3] bond vector R_ij = X_j + S C - X_i
    ABuf<int32_t> prev;           // [npairs] chunk-global index of the reverse pair (j,i,-S)
    ABuf<int64_t> cfg_pair_off;   // [ncfg+1] device copy of the per-config pair offsets
    std::vector<int64_t> h_cfg_pair_off;   // [ncfg+1] host copy
    std::vector<int64_t> h_site_off;       // [nsites+1] host copy (only filled while profiling)
};

// Config chunk resident on the device
struct ChunkDev {
    int64_t ncfg = 0, natoms = 0;
    FlexBuf<int64_t> atom_off;   // [ncfg+1], chunk-local
    FlexBuf<double> pos;         // [3*natoms]
    FlexBuf<double> cell;        // [9*ncfg]
    FlexBuf<uint8_t> pbc;        // [3*ncfg]
    FlexBuf<int64_t> rowE, rowF, rowV;   // [ncfg] 1-based, 0 = absent
    FlexBuf<int32_t> site_cfg;           // [natoms] config of every atom
    FlexBuf<int32_t> Z;                  // [natoms] atomic numbers (zeros when the caller passed none)
    bool has_Z = false;
    std::vector<int64_t> h_atom_off;
};

int ipf_build_neighbours(ipf_ctx* ctx, const ChunkDev& ch, double rcut, NeighList& nl);
int ipf_filter_neighbours(ipf_ctx* ctx, const ChunkDev& ch, const NeighList& big, double rcut, NeighList& nl);
int ipf_assemble_block_generic(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch,
                               const NeighList& nl, ipf_matrix* Psi);
int ipf_assemble_block_2b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled);
int ipf_assemble_block_4b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled);
void ipf_basis4b_forget(ipf_ctx* ctx, const ipf_basis* B);      // drop the tables of a basis that is being destroyed
void ipf_basis4b_forget_ctx(ipf_ctx* ctx);
int ipf_comm_allreduce_dev(ipf_ctx* ctx, double* buf, size_t n);    // in-place sum over the ranks on the ctx stream
int ipf_assemble_block_3b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled);

// ---- device helpers ---------------------------------------------------------
// distance transform x(r) and cutoff fc(r) of a descriptor with their r-derivatives (builder-defined maths,
// basis.py: transform 0 = exp(-a (r/r0 - 1)), 1 = (r0/r)^a;  CosCut(rc1, rc2)); one definition for every kernel
__device__ __forceinline__ void ipf_descriptor(double r, int transform, double a, double r0, double rc1, double rc2,
                                               double& x, double& dx, double& fc, double& dfc) {
    constexpr double kPiD = 3.14159265358979323846;
    if (transform == 0) {
        x = exp(-a * (r / r0 - 1.0));
        dx = -a / r0 * x;
    } else {
        x = pow(r0 / r, a);
        dx = -a * x / r;
    }
    if (r <= rc1) { fc = 1.0; dfc = 0.0; }
    else if (r >= rc2) { fc = 0.0; dfc = 0.0; }
    else {
        const double w = rc2 - rc1, s = (r - rc1) / w;
        fc = 0.5 * (1.0 + cos(kPiD * s));
        dfc = -0.5 * kPiD / w * sin(kPiD * s);
    }
}
__device__ __forceinline__ double ipf_bond_length(double Rx, double Ry, double Rz) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(Rx, Rx), __dmul_rn(Ry, Ry)), __dmul_rn(Rz, Rz)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
