// capi.cu -- extern "C" entry points of libipf_b200 (see include/ipf_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <cmath>
#include <new>

#include "ipf_common.cuh"

int ipf_set_error(ipf_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int ipf_scratch(ipf_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->stream2) IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
            cudaFree(ctx->scratch[slot]);
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 4;
        IPF_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return IPF_OK;
}

int ipf_arena_alloc(ipf_ctx* ctx, size_t bytes, void** out) { return ipf_arena_alloc_in(ctx, ctx->arena, bytes, out); }
int ipf_arena_reset(ipf_ctx* ctx) { return ipf_arena_reset_in(ctx, ctx->arena); }

int ipf_arena_alloc_in(ipf_ctx* ctx, Arena& a, size_t bytes, void** out) {
    bytes = (bytes + 255) & ~(size_t)255;
    a.need += bytes;
    if (a.off + bytes <= a.cap) {
        *out = a.base + a.off;
        a.off += bytes;
        return IPF_OK;
    }
    void* p = nullptr;
    IPF_CUDA(ctx, cudaMalloc(&p, bytes));
    a.spill.push_back(p);
    *out = p;
    return IPF_OK;
}

int ipf_arena_reset_in(ipf_ctx* ctx, Arena& a) {
    if (!a.spill.empty() || a.need > a.cap) {
        IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->stream2) IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
        for (void* p : a.spill) cudaFree(p);
        a.spill.clear();
        if (a.need > a.cap) {
            if (a.base) cudaFree(a.base);
            a.base = nullptr;
            a.cap = 0;
            const size_t want = a.need + a.need / 8 + (1 << 20);
            IPF_CUDA(ctx, cudaMalloc((void**)&a.base, want));
            a.cap = want;
        }
    }
    a.off = 0;
    a.need = 0;
    return IPF_OK;
}

extern "C" {

int32_t ipf_init(int32_t device, ipf_ctx** out) {
    if (!out) return IPF_EINVAL;
    *out = nullptr;
    ipf_ctx* ctx = new (std::nothrow) ipf_ctx();
    if (!ctx) return IPF_ENOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        int sm = 0;
        e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess && sm > 0) ctx->sm_count = sm;
    }
    if (e != cudaSuccess) {
        // leave the ctx alive so that the caller can read the message
        ipf_set_error(ctx, IPF_ECUDA, "ipf_init(device=%d): %s", device, cudaGetErrorString(e));
        *out = ctx;
        return IPF_ECUDA;
    }
    *out = ctx;
    return IPF_OK;
}

int32_t ipf_destroy(ipf_ctx* ctx) {
    if (!ctx) return IPF_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    for (int k = 0; k < ipf_ctx::kScratchSlots; ++k)
        if (ctx->scratch[k]) cudaFree(ctx->scratch[k]);
    ipf_basis4b_forget_ctx(ctx);
    ipf_comm_destroy(ctx);
    for (void* p : ctx->arena.spill) cudaFree(p);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    if (ctx->matrix_cache) cudaFree(ctx->matrix_cache);
    for (void* p : ctx->chunk_arena.spill) cudaFree(p);
    if (ctx->chunk_arena.base) cudaFree(ctx->chunk_arena.base);
    for (void* p : ctx->solve_arena.spill) cudaFree(p);
    if (ctx->solve_arena.base) cudaFree(ctx->solve_arena.base);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    delete ctx;
    return IPF_OK;
}

// 1: always use the table-driven kernels (tests compare the two paths)
int32_t ipf_set_force_generic(ipf_ctx* ctx, int32_t on) {
    if (!ctx) return IPF_EINVAL;
    ctx->force_generic = on == 1;
    ctx->force_scratch3b = on == 2;
    return IPF_OK;
}

int32_t ipf_set_async(ipf_ctx* ctx, int32_t on) {
    if (!ctx) return IPF_EINVAL;
    ctx->async_assemble = on != 0;
    return IPF_OK;
}

int32_t ipf_profile_enable(ipf_ctx* ctx, int32_t on) {
    if (!ctx) return IPF_EINVAL;
    ctx->profiling = on != 0;
    return IPF_OK;
}

int32_t ipf_profile_reset(ipf_ctx* ctx) {
    if (!ctx) return IPF_EINVAL;
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->prof) {
        for (auto& ev : kv.second.pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    }
    ctx->prof.clear();
    return IPF_OK;
}

int32_t ipf_profile_get(ipf_ctx* ctx, const char* name, double* ms, int64_t* count) {
    if (!ctx || !name) return IPF_EINVAL;
    auto it = ctx->prof.find(name);
    if (ms) *ms = 0.0;
    if (count) *count = 0;
    if (it == ctx->prof.end()) return IPF_OK;
    cudaStreamSynchronize(ctx->stream);
    ProfEntry& e = it->second;
    for (auto& ev : e.pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) e.ms += t;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    e.pending.clear();
    if (ms) *ms = e.ms;
    if (count) *count = e.count;
    return IPF_OK;
}

const char* ipf_last_error(ipf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t ipf_launch_count(ipf_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* ipf_stream(ipf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---------------------------------------------------------------- basis
int32_t ipf_basis_create(ipf_ctx* ctx, int32_t nblocks, const ipf_block_spec* blocks, ipf_basis** out) {
    if (!ctx || !out || nblocks <= 0 || !blocks) return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: bad arguments");
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    ipf_basis* B = new (std::nothrow) ipf_basis();
    if (!B) return IPF_ENOMEM;
    int64_t col = 0;
    for (int k = 0; k < nblocks; ++k) {
        const ipf_block_spec& s = blocks[k];
        const int nx = s.N - 1, nv = nx + nx * (nx - 1) / 2;
        if (s.N < 2 || s.N > 5 || s.nb <= 0 || s.nmono < s.nb || !s.mono_b || !s.mono_exp ||
            !(s.rc1 > 0.0 && s.rc2 > s.rc1) || !(s.r0 > 0.0) || (s.transform != 0 && s.transform != 1)) {
            delete B;
            return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d has invalid parameters", k);
        }
        BlockDev b;
        b.N = s.N; b.nb = s.nb; b.nmono = s.nmono; b.transform = s.transform;
        b.a = s.a; b.r0 = s.r0; b.rc1 = s.rc1; b.rc2 = s.rc2; b.col0 = col;
        if (s.zc < 0) { delete B; return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: negative species", k); }
        b.zc = s.zc;
        if (s.env_t < 0 || (s.env_t > 0 && !(s.env_rc1 > 0.0 && s.env_rc2 > s.env_rc1 && s.env_rc2 <= s.rc2))) {
            delete B;
            return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: environment cutoff must satisfy 0 < rc1 < rc2 <= descriptor cutoff", k);
        }
        b.env_t = s.env_t; b.env_rc1 = s.env_t > 0 ? s.env_rc1 : 0.0; b.env_rc2 = s.env_t > 0 ? s.env_rc2 : 0.0;
        for (int t = 0; t < 4; ++t) b.zs[t] = (s.zc > 0 && t < nx) ? s.zs[t] : 0;
        for (int t = 0; s.zc > 0 && t < nx; ++t)
            if (b.zs[t] <= 0 || (t > 0 && b.zs[t] < b.zs[t - 1])) {
                delete B;
                return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: neighbour species must be positive and sorted", k);
            }
        b.h_mono_start.assign(s.nb + 1, 0);
        b.h_mono_exp.assign(s.mono_exp, s.mono_exp + (size_t)s.nmono * IPF_MAX_VARS);
        int prevb = 0;
        for (int m = 0; m < s.nmono; ++m) {
            const int bb = s.mono_b[m];
            if (bb < prevb || bb >= s.nb || bb > prevb + 1) {
                delete B;
                return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: mono_b must be non-decreasing without gaps", k);
            }
            prevb = bb;
            b.h_mono_start[bb + 1]++;
            int deg = 0;
            for (int v = 0; v < IPF_MAX_VARS; ++v) {
                const int e = s.mono_exp[(size_t)m * IPF_MAX_VARS + v];
                if (v >= nv && e != 0) {
                    delete B;
                    return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: exponent on unused variable", k);
                }
                deg = std::max(deg, e);
            }
            b.maxdeg = std::max(b.maxdeg, deg);
        }
        if (b.maxdeg > IPF_MAX_DEGREE) {
            delete B;
            return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: degree > %d", IPF_MAX_DEGREE);
        }
        for (int i = 0; i < s.nb; ++i) {
            if (b.h_mono_start[i + 1] == 0) {
                delete B;
                return ipf_set_error(ctx, IPF_EINVAL, "ipf_basis_create: block %d: basis function %d has no monomial", k, i);
            }
            b.h_mono_start[i + 1] += b.h_mono_start[i];
        }
        B->blocks.push_back(std::move(b));
        col += s.nb;
    }
    B->ncols = col;
    for (auto& b : B->blocks) {
        cudaError_t e = cudaMalloc((void**)&b.mono_start, sizeof(int32_t) * (b.nb + 1));
        if (e == cudaSuccess) e = cudaMalloc((void**)&b.mono_exp, (size_t)b.nmono * IPF_MAX_VARS);
        if (e == cudaSuccess) e = cudaMemcpy(b.mono_start, b.h_mono_start.data(), sizeof(int32_t) * (b.nb + 1), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(b.mono_exp, b.h_mono_exp.data(), (size_t)b.nmono * IPF_MAX_VARS, cudaMemcpyHostToDevice);
        // species-resolved block: one exponent table per own-leg slot (first slot of every species), the slot moved to
        // position 0 and the pair variables renumbered accordingly
        const int nx = b.N - 1;
        for (int s0 = 0; e == cudaSuccess && b.zc > 0 && s0 < nx; ++s0) {
            if (s0 > 0 && b.zs[s0] == b.zs[s0 - 1]) continue;
            int order[4], k = 0;
            order[k++] = s0;
            for (int t = 0; t < nx; ++t) if (t != s0) order[k++] = t;
            auto pair_index = [&](int a, int c) {          // lexicographic index of the pair (a < c) among nx slots
                int idx = 0;
                for (int u = 0; u < a; ++u) idx += nx - 1 - u;
                return idx + (c - a - 1);
            };
            std::vector<uint8_t> tab((size_t)b.nmono * IPF_MAX_VARS, 0);
            for (int m = 0; m < b.nmono; ++m) {
                const uint8_t* src = &b.h_mono_exp[(size_t)m * IPF_MAX_VARS];
                uint8_t* dst = &tab[(size_t)m * IPF_MAX_VARS];
                for (int i = 0; i < nx; ++i) dst[i] = src[order[i]];
                for (int i = 0; i < nx; ++i)
                    for (int j = i + 1; j < nx; ++j) {
                        const int a = std::min(order[i], order[j]), c = std::max(order[i], order[j]);
                        dst[nx + pair_index(i, j)] = src[nx + pair_index(a, c)];
                    }
            }
            e = cudaMalloc((void**)&b.mono_exp_var[s0], tab.size());
            if (e == cudaSuccess) e = cudaMemcpy(b.mono_exp_var[s0], tab.data(), tab.size(), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            ipf_basis_destroy(ctx, B);
            return ipf_set_error(ctx, IPF_ECUDA, "ipf_basis_create: %s", cudaGetErrorString(e));
        }
    }
    *out = B;
    return IPF_OK;
}

int32_t ipf_basis_destroy(ipf_ctx* ctx, ipf_basis* B) {
    if (!B) return IPF_OK;
    if (ctx) { cudaSetDevice(ctx->device); ipf_basis4b_forget(ctx, B); }
    for (auto& b : B->blocks) {
        if (b.mono_start) cudaFree(b.mono_start);
        if (b.mono_exp) cudaFree(b.mono_exp);
        for (int t = 0; t < 4; ++t) if (b.mono_exp_var[t]) cudaFree(b.mono_exp_var[t]);
    }
    delete B;
    return IPF_OK;
}

int64_t ipf_basis_length(const ipf_basis* B) { return B ? B->ncols : 0; }

// ---------------------------------------------------------------- matrix
int32_t ipf_matrix_create(ipf_ctx* ctx, int64_t nrows, int64_t ncols, ipf_matrix** out) {
    if (!ctx || !out || nrows < 0 || ncols < 0) return ipf_set_error(ctx, IPF_EINVAL, "ipf_matrix_create: bad arguments");
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    ipf_matrix* M = new (std::nothrow) ipf_matrix();
    if (!M) return IPF_ENOMEM;
    M->nrows = nrows; M->ncols = ncols; M->ld = nrows;
    const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(1, nrows * ncols);
    if (ctx->matrix_cache && ctx->matrix_cache_bytes >= bytes && ctx->matrix_cache_bytes <= 2 * bytes + (1 << 20)) {
        M->d = ctx->matrix_cache;
        M->bytes = ctx->matrix_cache_bytes;
        ctx->matrix_cache = nullptr;
        ctx->matrix_cache_bytes = 0;
    } else {
        cudaError_t e = cudaMalloc((void**)&M->d, bytes);
        if (e != cudaSuccess) {
            delete M;
            return ipf_set_error(ctx, IPF_ENOMEM, "ipf_matrix_create(%lld x %lld): %s", (long long)nrows, (long long)ncols, cudaGetErrorString(e));
        }
        M->bytes = bytes;
    }
    IPF_CUDA(ctx, cudaMemsetAsync(M->d, 0, bytes, ctx->stream));
    IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = M;
    return IPF_OK;
}

int32_t ipf_matrix_destroy(ipf_ctx* ctx, ipf_matrix* M) {
    if (!M) return IPF_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    if (M->d) {
        if (ctx && M->bytes > ctx->matrix_cache_bytes) {       // keep the largest buffer for the next LsqDB
            if (ctx->matrix_cache) cudaFree(ctx->matrix_cache);
            ctx->matrix_cache = M->d;
            ctx->matrix_cache_bytes = M->bytes;
        } else {
            cudaFree(M->d);
        }
    }
    delete M;
    return IPF_OK;
}

int32_t ipf_matrix_to_host(ipf_ctx* ctx, const ipf_matrix* M, double* host, int64_t ld) {
    if (!ctx || !M || !host || ld < M->nrows) return ipf_set_error(ctx, IPF_EINVAL, "ipf_matrix_to_host: bad arguments");
    if (M->nrows == 0 || M->ncols == 0) return IPF_OK;
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    IPF_CUDA(ctx, cudaMemcpy2DAsync(host, sizeof(double) * ld, M->d, sizeof(double) * M->ld,
                                    sizeof(double) * M->nrows, M->ncols, cudaMemcpyDeviceToHost, ctx->stream));
    IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IPF_OK;
}

int32_t ipf_matrix_from_host(ipf_ctx* ctx, ipf_matrix* M, const double* host, int64_t ld) {
    if (!ctx || !M || !host || ld < M->nrows) return ipf_set_error(ctx, IPF_EINVAL, "ipf_matrix_from_host: bad arguments");
    if (M->nrows == 0 || M->ncols == 0) return IPF_OK;
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    IPF_CUDA(ctx, cudaMemcpy2DAsync(M->d, sizeof(double) * M->ld, host, sizeof(double) * ld,
                                    sizeof(double) * M->nrows, M->ncols, cudaMemcpyHostToDevice, ctx->stream));
    IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IPF_OK;
}

int64_t ipf_matrix_rows(const ipf_matrix* M) { return M ? M->nrows : 0; }
int64_t ipf_matrix_cols(const ipf_matrix* M) { return M ? M->ncols : 0; }
int64_t ipf_matrix_ld(const ipf_matrix* M) { return M ? M->ld : 0; }
void* ipf_matrix_device_ptr(const ipf_matrix* M) { return M ? (void*)M->d : nullptr; }

}  // extern "C"

// ---------------------------------------------------------------- chunk upload
static int upload_chunk(ipf_ctx* ctx, int64_t c0, int64_t c1, const int64_t* atom_off, const double* pos,
                        const double* cell, const uint8_t* pbc, const int64_t* rowE, const int64_t* rowF,
                        const int64_t* rowV, ChunkDev& ch, Arena* arena = nullptr, const int32_t* Z = nullptr) {
    cudaStream_t st = ctx->stream;
    const int64_t ncfg = c1 - c0;
    const int64_t a0 = atom_off[c0], natoms = atom_off[c1] - a0;
    ch.ncfg = ncfg; ch.natoms = natoms;
    ch.h_atom_off.resize(ncfg + 1);
    for (int64_t c = 0; c <= ncfg; ++c) ch.h_atom_off[c] = atom_off[c0 + c] - a0;
    std::vector<int64_t> zeros(ncfg, 0);
    IPF_CHECK(ch.atom_off.alloc(ctx, arena, ncfg + 1));
    IPF_CHECK(ch.pos.alloc(ctx, arena, 3 * natoms));
    IPF_CHECK(ch.cell.alloc(ctx, arena, 9 * ncfg));
    IPF_CHECK(ch.pbc.alloc(ctx, arena, 3 * ncfg));
    IPF_CHECK(ch.rowE.alloc(ctx, arena, ncfg));
    IPF_CHECK(ch.rowF.alloc(ctx, arena, ncfg));
    IPF_CHECK(ch.rowV.alloc(ctx, arena, ncfg));
    IPF_CHECK(ch.site_cfg.alloc(ctx, arena, natoms));
    std::vector<int32_t> h_site_cfg((size_t)natoms);
    for (int64_t c = 0; c < ncfg; ++c)
        for (int64_t s = ch.h_atom_off[c]; s < ch.h_atom_off[c + 1]; ++s) h_site_cfg[s] = (int32_t)c;
    if (natoms) IPF_CUDA(ctx, cudaMemcpyAsync(ch.site_cfg.p, h_site_cfg.data(), sizeof(int32_t) * natoms, cudaMemcpyHostToDevice, st));
    IPF_CHECK(ch.Z.alloc(ctx, arena, natoms));
    ch.has_Z = Z != nullptr;
    if (natoms) {
        if (Z) IPF_CUDA(ctx, cudaMemcpyAsync(ch.Z.p, Z + a0, sizeof(int32_t) * natoms, cudaMemcpyHostToDevice, st));
        else IPF_CUDA(ctx, cudaMemsetAsync(ch.Z.p, 0, sizeof(int32_t) * natoms, st));
    }
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.atom_off.p, ch.h_atom_off.data(), sizeof(int64_t) * (ncfg + 1), cudaMemcpyHostToDevice, st));
    if (natoms) IPF_CUDA(ctx, cudaMemcpyAsync(ch.pos.p, pos + 3 * a0, sizeof(double) * 3 * natoms, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.cell.p, cell + 9 * c0, sizeof(double) * 9 * ncfg, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.pbc.p, pbc + 3 * c0, 3 * ncfg, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.rowE.p, rowE ? rowE + c0 : zeros.data(), sizeof(int64_t) * ncfg, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.rowF.p, rowF ? rowF + c0 : zeros.data(), sizeof(int64_t) * ncfg, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaMemcpyAsync(ch.rowV.p, rowV ? rowV + c0 : zeros.data(), sizeof(int64_t) * ncfg, cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaStreamSynchronize(st));
    return IPF_OK;
}

static int validate_configs(ipf_ctx* ctx, int64_t ncfg, const int64_t* atom_off, const double* pos,
                            const double* cell) {
    if (atom_off[0] != 0) return ipf_set_error(ctx, IPF_EINVAL, "atom_off[0] must be 0");
    for (int64_t c = 0; c < ncfg; ++c) {
        const int64_t n = atom_off[c + 1] - atom_off[c];
        if (n < 0 || n > IPF_MAX_ATOMS)
            return ipf_set_error(ctx, IPF_EINVAL, "config %lld has %lld atoms (limit %d)", (long long)c, (long long)n, IPF_MAX_ATOMS);
        const double* C = cell + 9 * c;
        const double det = C[0] * (C[4] * C[8] - C[5] * C[7]) - C[1] * (C[3] * C[8] - C[5] * C[6]) + C[2] * (C[3] * C[7] - C[4] * C[6]);
        if (!(std::fabs(det) > 0.0) || !std::isfinite(det))
            return ipf_set_error(ctx, IPF_EINVAL, "config %lld has a singular cell", (long long)c);
    }
    const int64_t nat = atom_off[ncfg];
    for (int64_t k = 0; k < 3 * nat; ++k)
        if (!std::isfinite(pos[k])) return ipf_set_error(ctx, IPF_EINVAL, "non-finite position");
    return IPF_OK;
}

extern "C" {

int32_t ipf_neighbours(ipf_ctx* ctx, int32_t nat, const double* pos, const double* cell, const uint8_t* pbc,
                       double rcut, int64_t cap, int64_t* n, int32_t* I, int32_t* J, int32_t* S, double* R) {
    if (!ctx || nat < 0 || !pos || !cell || !pbc || !n || !(rcut > 0.0))
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_neighbours: bad arguments");
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t off[2] = {0, nat};
    IPF_CHECK(validate_configs(ctx, 1, off, pos, cell));
    ChunkDev ch;
    IPF_CHECK(upload_chunk(ctx, 0, 1, off, pos, cell, pbc, nullptr, nullptr, nullptr, ch));
    NeighList nl;
    IPF_CHECK(ipf_arena_reset(ctx));
    IPF_CHECK(ipf_build_neighbours(ctx, ch, rcut, nl));
    *n = nl.npairs;
    const int64_t m = std::min<int64_t>(cap, nl.npairs);
    if (m > 0) {
        // same stream as the kernels that produce the list (a non-blocking stream does not order against stream 0)
        cudaStream_t st = ctx->stream;
        if (I) IPF_CUDA(ctx, cudaMemcpyAsync(I, nl.pi.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
        if (J) IPF_CUDA(ctx, cudaMemcpyAsync(J, nl.pj.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
        if (S) IPF_CUDA(ctx, cudaMemcpyAsync(S, nl.pS.p, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost, st));
        if (R) IPF_CUDA(ctx, cudaMemcpyAsync(R, nl.pR.p, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
    }
    IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IPF_OK;
}

// Device-resident configuration set: the chunks ipf_assemble works through.
struct ipf_configs {
    std::vector<ChunkDev*> chunks;
    int64_t ncfg = 0, natoms = 0, max_row = 0;
    ~ipf_configs() { for (auto* c : chunks) delete c; }
};

static int32_t configs_create_impl(ipf_ctx* ctx, int64_t ncfg, const int64_t* atom_off, const double* pos,
                                   const double* cell, const uint8_t* pbc, const int32_t* Z, const int64_t* rowE,
                                   const int64_t* rowF, const int64_t* rowV, ipf_configs** out, Arena* arena);

int32_t ipf_configs_create(ipf_ctx* ctx, int64_t ncfg, const int64_t* atom_off, const double* pos,
                           const double* cell, const uint8_t* pbc, const int32_t* Z, const int64_t* rowE,
                           const int64_t* rowF, const int64_t* rowV, ipf_configs** out) {
    return configs_create_impl(ctx, ncfg, atom_off, pos, cell, pbc, Z, rowE, rowF, rowV, out, nullptr);
}

static int32_t configs_create_impl(ipf_ctx* ctx, int64_t ncfg, const int64_t* atom_off, const double* pos,
                                   const double* cell, const uint8_t* pbc, const int32_t* Z, const int64_t* rowE,
                                   const int64_t* rowF, const int64_t* rowV, ipf_configs** out, Arena* arena) {
    if (!ctx || !out || ncfg < 0 || !atom_off || (ncfg && (!cell || !pbc || !rowE || !rowF || !rowV)))
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_configs_create: bad arguments");
    *out = nullptr;
    if (ncfg && !pos && atom_off[ncfg] > 0) return ipf_set_error(ctx, IPF_EINVAL, "ipf_configs_create: pos is NULL");
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ncfg) IPF_CHECK(validate_configs(ctx, ncfg, atom_off, pos, cell));
    ipf_configs* cf = new (std::nothrow) ipf_configs();
    if (!cf) return IPF_ENOMEM;
    cf->ncfg = ncfg;
    cf->natoms = ncfg ? atom_off[ncfg] : 0;
    for (int64_t c = 0; c < ncfg; ++c) {
        const int64_t nat = atom_off[c + 1] - atom_off[c];
        const int64_t r[3] = {rowE[c], rowF[c], rowV[c]};
        const int64_t len[3] = {1, 3 * nat, 6};
        for (int k = 0; k < 3; ++k) {
            if (r[k] < 0) { delete cf; return ipf_set_error(ctx, IPF_EINVAL, "config %lld: negative row", (long long)c); }
            if (r[k] > 0) cf->max_row = std::max(cf->max_row, r[k] - 1 + len[k]);
        }
    }
    // chunks of configurations: bounded scratch per launch
    const int64_t max_atoms_chunk = 1 << 16;
    int64_t c0 = 0;
    while (c0 < ncfg) {
        int64_t c1 = c0 + 1;
        while (c1 < ncfg && atom_off[c1 + 1] - atom_off[c0] <= max_atoms_chunk) ++c1;
        ChunkDev* ch = new (std::nothrow) ChunkDev();
        if (!ch) { delete cf; return IPF_ENOMEM; }
        cf->chunks.push_back(ch);
        int rc = upload_chunk(ctx, c0, c1, atom_off, pos, cell, pbc, rowE, rowF, rowV, *ch, arena, Z);
        if (rc != IPF_OK) { delete cf; return rc; }
        c0 = c1;
    }
    *out = cf;
    return IPF_OK;
}

int32_t ipf_configs_destroy(ipf_ctx* ctx, ipf_configs* cf) {
    if (!cf) return IPF_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    delete cf;
    return IPF_OK;
}

int32_t ipf_assemble_configs(ipf_ctx* ctx, const ipf_basis* basis, const ipf_configs* cf, ipf_matrix* Psi) {
    if (!ctx || !basis || !cf || !Psi) return ipf_set_error(ctx, IPF_EINVAL, "ipf_assemble_configs: bad arguments");
    if (Psi->ncols != basis->ncols)
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_assemble: Psi has %lld columns, basis has %lld",
                             (long long)Psi->ncols, (long long)basis->ncols);
    if (cf->max_row > Psi->nrows)
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_assemble: row block out of range (%lld > %lld rows)",
                             (long long)cf->max_row, (long long)Psi->nrows);
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    // distinct cutoffs, largest first: its neighbour list is built from the positions, the lists of the smaller
    // cutoffs are sub-lists of it (ipf_filter_neighbours; IPF_NL_REBUILD=1 builds every list from scratch)
    std::vector<double> cutoffs;
    for (const BlockDev& b : basis->blocks)
        if (std::find(cutoffs.begin(), cutoffs.end(), b.rc2) == cutoffs.end()) cutoffs.push_back(b.rc2);
    std::sort(cutoffs.begin(), cutoffs.end(), std::greater<double>());
    static const bool nl_rebuild = getenv("IPF_NL_REBUILD") != nullptr;
    for (const ChunkDev* ch : cf->chunks) {
        IPF_CHECK(ipf_arena_reset(ctx));
        NeighList big;
        for (size_t ic = 0; ic < cutoffs.size(); ++ic) {
            const double rc = cutoffs[ic];
            NeighList sub;
            if (ic == 0) IPF_CHECK(ipf_build_neighbours(ctx, *ch, rc, big));
            else if (nl_rebuild) IPF_CHECK(ipf_build_neighbours(ctx, *ch, rc, sub));
            else IPF_CHECK(ipf_filter_neighbours(ctx, *ch, big, rc, sub));
            const NeighList& nl = ic == 0 ? big : sub;
            for (size_t k2 = 0; k2 < basis->blocks.size(); ++k2)
                if (basis->blocks[k2].rc2 == rc) {
                    bool handled = false;
                    if (basis->blocks[k2].zc > 0 || basis->blocks[k2].env_t > 0) {
                        // species-resolved / environment-dependent run: table-driven kernel (the specialised kernels
                        // assume every cluster and the plain site function)
                        if (basis->blocks[k2].zc > 0 && !ch->has_Z)
                            return ipf_set_error(ctx, IPF_EINVAL, "a species-resolved basis block needs the atomic numbers Z");
                        IPF_CHECK(ipf_assemble_block_generic(ctx, basis->blocks[k2], *ch, nl, Psi));
                        continue;
                    }
                    if (!ctx->force_generic) IPF_CHECK(ipf_assemble_block_2b(ctx, basis->blocks[k2], *ch, nl, Psi, &handled));
                    if (!ctx->force_generic && !handled) IPF_CHECK(ipf_assemble_block_3b(ctx, basis->blocks[k2], *ch, nl, Psi, &handled));
                    if (!ctx->force_generic && !handled) IPF_CHECK(ipf_assemble_block_4b(ctx, basis->blocks[k2], *ch, nl, Psi, &handled));
                    if (!handled) IPF_CHECK(ipf_assemble_block_generic(ctx, basis->blocks[k2], *ch, nl, Psi));
                }
        }
    }
    if (!ctx->async_assemble) IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IPF_OK;
}

int32_t ipf_assemble(ipf_ctx* ctx, const ipf_basis* basis, int64_t ncfg, const int64_t* atom_off,
                     const double* pos, const double* cell, const uint8_t* pbc, const int32_t* Z,
                     const int64_t* rowE, const int64_t* rowF, const int64_t* rowV, ipf_matrix* Psi) {
    if (!ctx || !basis || !Psi) return ipf_set_error(ctx, IPF_EINVAL, "ipf_assemble: bad arguments");
    ipf_configs* cf = nullptr;
    // transient chunks: carved from a recycled arena (no cudaMalloc/cudaFree per call)
    IPF_CHECK(ipf_arena_reset_in(ctx, ctx->chunk_arena));
    IPF_CHECK(configs_create_impl(ctx, ncfg, atom_off, pos, cell, pbc, Z, rowE, rowF, rowV, &cf, &ctx->chunk_arena));
    const int rc = ipf_assemble_configs(ctx, basis, cf, Psi);
    // the chunks live in the recycled arena (stream-ordered reuse): only the host-side descriptors go away
    if (ctx->async_assemble && rc == IPF_OK) delete cf;
    else ipf_configs_destroy(ctx, cf);
    return rc;
}

}  // extern "C"
