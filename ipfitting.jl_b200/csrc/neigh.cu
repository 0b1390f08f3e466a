// neigh.cu -- K1: per-configuration linked-cell neighbour build with periodic images.
//
// Replaces JuLIP.neighbourlist(at, rcut) as called inside energy/forces/virial
// (reference call sites src/datatypes.jl:28,40,65; src/preconditioners.jl:56).
// Output is the canonical CSR: for every centre atom i all (j, S) with
// |X_j + S C - X_i| < rcut, (j,S) != (i,0), sorted by (j, Sx, Sy, Sz).
// The accept test uses individually rounded operations in the same order as the
// oracle (oracle/ipf_oracle.c: bond()), so the list is bit-exact, including the
// borderline pairs.  Candidate generation (binning, images) is free to differ.
#include <cub/device/device_scan.cuh>

#include "ipf_common.cuh"

namespace {

constexpr int NC_MAX = 8;                      // cells per direction
constexpr int NCELL_MAX = NC_MAX * NC_MAX * NC_MAX;
constexpr int SHIFT_BIAS = 512;

struct GridInfo {          // per config
    int ncell[3];
    int m[3];              // search range in cells
    double Ci[9];          // inverse cell (s = X * Ci)
};

__device__ __forceinline__ void inv3(const double* C, double* Ci) {
    double a = C[0], b = C[1], c = C[2], d = C[3], e = C[4], f = C[5], g = C[6], h = C[7], k = C[8];
    double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    double id = 1.0 / det;
    Ci[0] = (e * k - f * h) * id; Ci[1] = (c * h - b * k) * id; Ci[2] = (b * f - c * e) * id;
    Ci[3] = (f * g - d * k) * id; Ci[4] = (a * k - c * g) * id; Ci[5] = (c * d - a * f) * id;
    Ci[6] = (d * h - e * g) * id; Ci[7] = (b * g - a * h) * id; Ci[8] = (a * e - b * d) * id;
}

// one CTA per config: grid geometry, wrap counts, cell ids, cell-sorted atom list
__global__ void nl_bin_kernel(const int64_t* __restrict__ atom_off, const double* __restrict__ pos,
                              const double* __restrict__ cell, const uint8_t* __restrict__ pbc,
                              double rcut, GridInfo* __restrict__ grid, int32_t* __restrict__ cell_start,
                              int32_t* __restrict__ cell_atoms, int32_t* __restrict__ atom_cell,
                              int32_t* __restrict__ atom_wrap) {
    const int c = blockIdx.x;
    const int64_t a0 = atom_off[c];
    const int nat = (int)(atom_off[c + 1] - a0);
    __shared__ GridInfo g;
    __shared__ int cnt[NCELL_MAX + 1];
    if (threadIdx.x == 0) {
        const double* C = cell + 9 * c;
        inv3(C, g.Ci);
        for (int a = 0; a < 3; ++a) {
            if (!pbc[3 * c + a]) { g.ncell[a] = 1; g.m[a] = 0; continue; }
            double n2 = g.Ci[a] * g.Ci[a] + g.Ci[3 + a] * g.Ci[3 + a] + g.Ci[6 + a] * g.Ci[6 + a];
            double h = 1.0 / sqrt(n2);
            int nc = (int)floor(h / (0.5 * rcut));
            nc = nc < 1 ? 1 : (nc > NC_MAX ? NC_MAX : nc);
            g.ncell[a] = nc;
            g.m[a] = (int)floor(rcut / (h / nc)) + 1;
        }
        grid[c] = g;
    }
    for (int k = threadIdx.x; k <= NCELL_MAX; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nat; i += blockDim.x) {
        const double* X = pos + 3 * (a0 + i);
        int cid = 0;
        for (int a = 0; a < 3; ++a) {
            int n = 0, ci = 0;
            if (pbc[3 * c + a]) {
                double s = X[0] * g.Ci[a] + X[1] * g.Ci[3 + a] + X[2] * g.Ci[6 + a];
                double fl = floor(s);
                n = (int)fl;
                ci = (int)((s - fl) * g.ncell[a]);
                if (ci >= g.ncell[a]) ci = g.ncell[a] - 1;
                if (ci < 0) ci = 0;
            }
            atom_wrap[3 * (a0 + i) + a] = n;
            cid = cid * g.ncell[a] + ci;
        }
        atom_cell[a0 + i] = cid;
        atomicAdd(&cnt[cid], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        const int nctot = g.ncell[0] * g.ncell[1] * g.ncell[2];
        for (int k = 0; k < nctot; ++k) { int v = cnt[k]; cnt[k] = acc; cell_start[(int64_t)c * (NCELL_MAX + 1) + k] = acc; acc += v; }
        for (int k = nctot; k <= NCELL_MAX; ++k) cell_start[(int64_t)c * (NCELL_MAX + 1) + k] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nat; i += blockDim.x) {
        int slot = atomicAdd(&cnt[atom_cell[a0 + i]], 1);
        cell_atoms[a0 + slot] = i;
    }
}

__device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// exact bond arithmetic (mirrors oracle bond())
__device__ __forceinline__ double bond(const double* __restrict__ Xi, const double* __restrict__ Xj,
                                       const double* __restrict__ C, int sx, int sy, int sz, double* R) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double sh = __dadd_rn(__dadd_rn(__dmul_rn((double)sx, C[d]), __dmul_rn((double)sy, C[3 + d])),
                              __dmul_rn((double)sz, C[6 + d]));
        R[d] = __dadd_rn(__dsub_rn(Xj[d], Xi[d]), sh);
    }
    return __dadd_rn(__dadd_rn(__dmul_rn(R[0], R[0]), __dmul_rn(R[1], R[1])), __dmul_rn(R[2], R[2]));
}

// the lattice shift of bond() alone, and the bond for a given shift: the same rounded operations in the same order,
// so that the shift of a neighbouring cell can be computed once and reused for every atom in it
__device__ __forceinline__ void bond_shift(const double* __restrict__ C, int sx, int sy, int sz, double* sh) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
        sh[d] = __dadd_rn(__dadd_rn(__dmul_rn((double)sx, C[d]), __dmul_rn((double)sy, C[3 + d])),
                          __dmul_rn((double)sz, C[6 + d]));
}
__device__ __forceinline__ double bond_with_shift(const double* __restrict__ Xi, const double* __restrict__ Xj,
                                                  const double* __restrict__ sh, double* R) {
#pragma unroll
    for (int d = 0; d < 3; ++d) R[d] = __dadd_rn(__dsub_rn(Xj[d], Xi[d]), sh[d]);
    return __dadd_rn(__dadd_rn(__dmul_rn(R[0], R[0]), __dmul_rn(R[1], R[1])), __dmul_rn(R[2], R[2]));
}

// NL_SUB threads per centre atom (they split the (x, y) columns of neighbouring cells: four times the threads, the
// kernel is latency bound).  FILL = false: count per (atom, sub-thread); true: write (key, S, R) at the sub-thread's
// offset, in discovery order (nl_sort_kernel puts every list into its canonical order afterwards).
constexpr int NL_SUB = 4;
template <bool FILL>
__global__ void nl_pairs_kernel(int64_t nsites, const int32_t* __restrict__ site_cfg,
                                const int64_t* __restrict__ atom_off, const double* __restrict__ pos,
                                const double* __restrict__ cell, double rc2,
                                const GridInfo* __restrict__ grid, const int32_t* __restrict__ cell_start,
                                const int32_t* __restrict__ cell_atoms, const int32_t* __restrict__ atom_cell,
                                const int32_t* __restrict__ atom_wrap, int64_t* __restrict__ counts,
                                const int64_t* __restrict__ site_off, unsigned long long* __restrict__ keys,
                                double* __restrict__ Rtmp, unsigned long long* __restrict__ maxn) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = gt / NL_SUB;
    const int sub = (int)(gt - s * NL_SUB);
    if (s >= nsites) return;
    const int c = site_cfg[s];
    const int64_t a0 = atom_off[c];
    const int i = (int)(s - a0);
    const GridInfo g = grid[c];
    const double* C = cell + 9 * c;
    const double* Xi = pos + 3 * s;
    const int32_t* cs = cell_start + (int64_t)c * (NCELL_MAX + 1);
    int cid = atom_cell[s];
    const int cz = cid % g.ncell[2]; cid /= g.ncell[2];
    const int cy = cid % g.ncell[1];
    const int cx = cid / g.ncell[1];
    const int nix = atom_wrap[3 * s], niy = atom_wrap[3 * s + 1], niz = atom_wrap[3 * s + 2];
    int64_t n = 0;
    const int64_t o0 = FILL ? site_off[NL_SUB * s + sub] : 0;      // site_off = offsets of the (atom, sub-thread) pieces here
    const int ny = 2 * g.m[1] + 1, nxy = (2 * g.m[0] + 1) * ny;
    for (int oxy = sub; oxy < nxy; oxy += NL_SUB) {
        const int ox = oxy / ny - g.m[0], oy = oxy % ny - g.m[1];
        const int tx = cx + ox, sx = floordiv(tx, g.ncell[0]), ccx = tx - sx * g.ncell[0];
        {
            const int ty = cy + oy, sy = floordiv(ty, g.ncell[1]), ccy = ty - sy * g.ncell[1];
            for (int oz = -g.m[2]; oz <= g.m[2]; ++oz) {
                const int tz = cz + oz, sz = floordiv(tz, g.ncell[2]), ccz = tz - sz * g.ncell[2];
                const int cc = (ccx * g.ncell[1] + ccy) * g.ncell[2] + ccz;
                const int q0 = cs[cc], q1 = cs[cc + 1];
                if (q0 == q1) continue;
                // shift of this cell image for atoms wrapped like the centre atom (the common case)
                double sh0[3];
                bond_shift(C, sx, sy, sz, sh0);
                for (int q = q0; q < q1; ++q) {
                    const int j = cell_atoms[a0 + q];
                    const int Sx = sx - atom_wrap[3 * (a0 + j)] + nix;
                    const int Sy = sy - atom_wrap[3 * (a0 + j) + 1] + niy;
                    const int Sz = sz - atom_wrap[3 * (a0 + j) + 2] + niz;
                    if (j == i && Sx == 0 && Sy == 0 && Sz == 0) continue;
                    double R[3];
                    double r2;
                    if (Sx == sx && Sy == sy && Sz == sz) r2 = bond_with_shift(Xi, pos + 3 * (a0 + j), sh0, R);
                    else r2 = bond(Xi, pos + 3 * (a0 + j), C, Sx, Sy, Sz, R);
                    if (r2 < rc2) {
                        if (FILL) {
                            keys[o0 + n] = ((unsigned long long)(unsigned)j << 32) |
                                           ((unsigned long long)(Sx + SHIFT_BIAS) << 20) |
                                           ((unsigned long long)(Sy + SHIFT_BIAS) << 10) |
                                           (unsigned long long)(Sz + SHIFT_BIAS);
                            Rtmp[3 * (o0 + n)] = R[0]; Rtmp[3 * (o0 + n) + 1] = R[1]; Rtmp[3 * (o0 + n) + 2] = R[2];
                        }
                        ++n;
                    }
                }
            }
        }
    }
    if (!FILL) counts[NL_SUB * s + sub] = n;
}

// site offsets from the offsets of the (atom, sub-thread) pieces; maxn = longest list
__global__ void nl_site_off_kernel(int64_t nsites, const int64_t* __restrict__ sub_off, int64_t* __restrict__ site_off,
                                   unsigned long long* __restrict__ maxn) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned cnt = 0;
    if (s <= nsites) {
        site_off[s] = sub_off[NL_SUB * s];
        if (s < nsites) cnt = (unsigned)(sub_off[NL_SUB * (s + 1)] - sub_off[NL_SUB * s]);
    }
    cnt = __reduce_max_sync(0xffffffffu, cnt);          // one atomic per warp
    if ((threadIdx.x & 31) == 0 && cnt) atomicMax(maxn, (unsigned long long)cnt);
}

__global__ void cfg_offsets_kernel(int64_t ncfg, const int64_t* __restrict__ atom_off,
                                   const int64_t* __restrict__ site_off, int64_t* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= ncfg) out[c] = site_off[atom_off[c]];
}

// thread per pair: rank within the site by key -> canonical order
__global__ void nl_sort_kernel(int64_t npairs, int64_t nsites, const int64_t* __restrict__ site_off,
                               const int32_t* __restrict__ site_cfg, const int64_t* __restrict__ atom_off,
                               const unsigned long long* __restrict__ keys, const double* __restrict__ Rtmp,
                               int32_t* __restrict__ pi, int32_t* __restrict__ pj, int32_t* __restrict__ psite,
                               int64_t* __restrict__ pbeg, int32_t* __restrict__ pcnt,
                               int32_t* __restrict__ pS, double* __restrict__ pR,
                               unsigned long long* __restrict__ keys_sorted) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    // site of pair p: binary search in site_off
    int64_t lo = 0, hi = nsites;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (site_off[mid] <= p) lo = mid; else hi = mid; }
    const int64_t s = lo, o0 = site_off[s], o1 = site_off[s + 1];
    const unsigned long long k = keys[p];
    int rank = 0;
    for (int64_t q = o0; q < o1; ++q) rank += (keys[q] < k);
    const int64_t dst = o0 + rank;
    keys_sorted[dst] = k;
    pi[dst] = (int32_t)(s - atom_off[site_cfg[s]]);
    pj[dst] = (int32_t)(k >> 32);
    psite[dst] = (int32_t)s;
    pbeg[dst] = o0;
    pcnt[dst] = (int32_t)(o1 - o0);
    pS[3 * dst] = (int)((k >> 20) & 1023) - SHIFT_BIAS;
    pS[3 * dst + 1] = (int)((k >> 10) & 1023) - SHIFT_BIAS;
    pS[3 * dst + 2] = (int)(k & 1023) - SHIFT_BIAS;
    pR[3 * dst] = Rtmp[3 * p]; pR[3 * dst + 1] = Rtmp[3 * p + 1]; pR[3 * dst + 2] = Rtmp[3 * p + 2];
}

// thread per pair (i,j,S): locate (j,i,-S) in site j's sorted list
__global__ void nl_reverse_kernel(int64_t npairs, const int64_t* __restrict__ site_off,
                                  const int32_t* __restrict__ site_cfg_of_pair_site,
                                  const int64_t* __restrict__ atom_off, int64_t nsites,
                                  const int32_t* __restrict__ pi, const int32_t* __restrict__ pj,
                                  const int32_t* __restrict__ pS, const unsigned long long* __restrict__ keys_sorted,
                                  int32_t* __restrict__ prev) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    int64_t lo = 0, hi = nsites;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (site_off[mid] <= p) lo = mid; else hi = mid; }
    const int64_t s = lo;
    const int64_t a0 = atom_off[site_cfg_of_pair_site[s]];
    const int64_t sj = a0 + pj[p];
    const unsigned long long want = ((unsigned long long)(unsigned)pi[p] << 32) |
                                    ((unsigned long long)(-pS[3 * p] + SHIFT_BIAS) << 20) |
                                    ((unsigned long long)(-pS[3 * p + 1] + SHIFT_BIAS) << 10) |
                                    (unsigned long long)(-pS[3 * p + 2] + SHIFT_BIAS);
    int64_t l = site_off[sj], h = site_off[sj + 1];
    int32_t found = -1;
    while (l < h) {
        int64_t mid = (l + h) >> 1;
        unsigned long long km = keys_sorted[mid];
        if (km == want) { found = (int32_t)mid; break; }
        if (km < want) l = mid + 1; else h = mid;
    }
    prev[p] = found;
}

// ---- derive the list of a smaller cutoff from an existing one -------------------------------------------
__global__ void nlf_flag_kernel(int64_t npairs, const double* __restrict__ pR, double rc2, int32_t* __restrict__ keep) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > npairs) return;
    if (p == npairs) { keep[p] = 0; return; }
    const double R0 = pR[3 * p], R1 = pR[3 * p + 1], R2 = pR[3 * p + 2];
    // same individually rounded sum and comparison as pair_r2() / nl_pairs_kernel
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(R0, R0), __dmul_rn(R1, R1)), __dmul_rn(R2, R2));
    keep[p] = r2 < rc2 ? 1 : 0;
}

// new site / config offsets = new index of the old offsets; [ncfg+1] = max neighbours of one site
__global__ void nlf_offsets_kernel(int64_t nsites, int64_t ncfg, const int64_t* __restrict__ site_off_old,
                                   const int64_t* __restrict__ cfg_off_old, const int32_t* __restrict__ newidx,
                                   int64_t* __restrict__ site_off, int64_t* __restrict__ cfg_off) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= nsites) site_off[s] = newidx[site_off_old[s]];
    if (s <= ncfg) cfg_off[s] = newidx[cfg_off_old[s]];
    if (s < nsites) {
        const int64_t cnt = (int64_t)newidx[site_off_old[s + 1]] - newidx[site_off_old[s]];
        atomicMax((unsigned long long*)(cfg_off + ncfg + 1), (unsigned long long)cnt);
    }
}

__global__ void nlf_compact_kernel(int64_t npairs, const int32_t* __restrict__ keep, const int32_t* __restrict__ newidx,
                                   const int64_t* __restrict__ site_off,
                                   const int32_t* __restrict__ pi0, const int32_t* __restrict__ pj0,
                                   const int32_t* __restrict__ psite0, const int32_t* __restrict__ pS0,
                                   const double* __restrict__ pR0, const int32_t* __restrict__ prev0,
                                   int32_t* __restrict__ pi, int32_t* __restrict__ pj, int32_t* __restrict__ psite,
                                   int64_t* __restrict__ pbeg, int32_t* __restrict__ pcnt, int32_t* __restrict__ pS,
                                   double* __restrict__ pR, int32_t* __restrict__ prev) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs || !keep[p]) return;
    const int64_t q = newidx[p];
    const int32_t s = psite0[p];
    pi[q] = pi0[p]; pj[q] = pj0[p]; psite[q] = s;
    pbeg[q] = site_off[s];
    pcnt[q] = (int32_t)(site_off[s + 1] - site_off[s]);
    pS[3 * q] = pS0[3 * p]; pS[3 * q + 1] = pS0[3 * p + 1]; pS[3 * q + 2] = pS0[3 * p + 2];
    pR[3 * q] = pR0[3 * p]; pR[3 * q + 1] = pR0[3 * p + 1]; pR[3 * q + 2] = pR0[3 * p + 2];
    const int32_t rv = prev0[p];
    prev[q] = rv >= 0 ? newidx[rv] : -1;          // the reverse pair has the same length: it is kept too
}

}  // namespace

// Neighbour list of cutoff rcut <= big.rcut as the sub-list of `big` (same chunk): identical, entry for entry,
// to ipf_build_neighbours(ctx, ch, rcut, nl) -- the accept test is re-evaluated on the stored bond vectors with
// the same rounded operations, and the order within a site is inherited -- at a fraction of the cost.
int ipf_filter_neighbours(ipf_ctx* ctx, const ChunkDev& ch, const NeighList& big, double rcut, NeighList& nl) {
    const int64_t ncfg = ch.ncfg, nsites = ch.natoms;
    nl.serial = ctx->nl_serial++;
    nl.nsites = nsites;
    nl.rcut = rcut;
    nl.npairs = 0;
    nl.maxpairs_cfg = 0;
    nl.max_nbrs = 0;
    nl.h_cfg_pair_off.assign(ncfg + 2, 0);
    nl.h_site_off.clear();
    cudaStream_t st = ctx->stream;
    IPF_CHECK(nl.site_off.alloc(ctx, nsites + 1));
    IPF_CHECK(nl.cfg_pair_off.alloc(ctx, ncfg + 2));
    if (nsites == 0 || big.npairs == 0) {
        IPF_CUDA(ctx, cudaMemsetAsync(nl.site_off.p, 0, sizeof(int64_t) * (nsites + 1), st));
        IPF_CUDA(ctx, cudaMemsetAsync(nl.cfg_pair_off.p, 0, sizeof(int64_t) * (ncfg + 2), st));
        IPF_CHECK(nl.pi.alloc(ctx, 0)); IPF_CHECK(nl.pj.alloc(ctx, 0)); IPF_CHECK(nl.psite.alloc(ctx, 0));
        IPF_CHECK(nl.pbeg.alloc(ctx, 0)); IPF_CHECK(nl.pcnt.alloc(ctx, 0)); IPF_CHECK(nl.pS.alloc(ctx, 0));
        IPF_CHECK(nl.pR.alloc(ctx, 0)); IPF_CHECK(nl.prev.alloc(ctx, 0));
        return IPF_OK;
    }
    ProfScope prof_(ctx, "neighbour_list");
    const int64_t np0 = big.npairs;
    ABuf<int32_t> keep, newidx;
    IPF_CHECK(keep.alloc(ctx, np0 + 1));
    IPF_CHECK(newidx.alloc(ctx, np0 + 1));
    const unsigned nbp = (unsigned)((np0 + 1 + 255) / 256);
    nlf_flag_kernel<<<nbp, 256, 0, st>>>(np0, big.pR.p, rcut * rcut, keep.p);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, keep.p, newidx.p, (int)(np0 + 1), st);
    void* tmp = nullptr;
    IPF_CHECK(ipf_arena_alloc(ctx, tmp_bytes, &tmp));
    IPF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, keep.p, newidx.p, (int)(np0 + 1), st));
    IPF_CUDA(ctx, cudaMemsetAsync(nl.cfg_pair_off.p + ncfg + 1, 0, sizeof(int64_t), st));
    const int64_t nmax = std::max(nsites, ncfg) + 1;
    nlf_offsets_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(nsites, ncfg, big.site_off.p, big.cfg_pair_off.p,
                                                                      newidx.p, nl.site_off.p, nl.cfg_pair_off.p);
    ctx->launches += 3;
    IPF_CUDA(ctx, cudaMemcpyAsync(nl.h_cfg_pair_off.data(), nl.cfg_pair_off.p, sizeof(int64_t) * (ncfg + 2),
                                  cudaMemcpyDeviceToHost, st));
    if (ctx->profiling) {
        nl.h_site_off.assign(nsites + 1, 0);
        IPF_CUDA(ctx, cudaMemcpyAsync(nl.h_site_off.data(), nl.site_off.p, sizeof(int64_t) * (nsites + 1),
                                      cudaMemcpyDeviceToHost, st));
    }
    IPF_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t npairs = nl.h_cfg_pair_off[ncfg];
    nl.max_nbrs = (int32_t)nl.h_cfg_pair_off[ncfg + 1];
    nl.npairs = npairs;
    int64_t mx = 0;
    for (int64_t c = 0; c < ncfg; ++c) mx = std::max<int64_t>(mx, nl.h_cfg_pair_off[c + 1] - nl.h_cfg_pair_off[c]);
    nl.maxpairs_cfg = (int32_t)mx;
    IPF_CHECK(nl.pi.alloc(ctx, npairs));
    IPF_CHECK(nl.pj.alloc(ctx, npairs));
    IPF_CHECK(nl.psite.alloc(ctx, npairs));
    IPF_CHECK(nl.pbeg.alloc(ctx, npairs));
    IPF_CHECK(nl.pcnt.alloc(ctx, npairs));
    IPF_CHECK(nl.pS.alloc(ctx, 3 * npairs));
    IPF_CHECK(nl.pR.alloc(ctx, 3 * npairs));
    IPF_CHECK(nl.prev.alloc(ctx, npairs));
    if (npairs == 0) return IPF_OK;
    nlf_compact_kernel<<<(unsigned)((np0 + 255) / 256), 256, 0, st>>>(np0, keep.p, newidx.p, nl.site_off.p, big.pi.p, big.pj.p,
                                                                     big.psite.p, big.pS.p, big.pR.p, big.prev.p, nl.pi.p,
                                                                     nl.pj.p, nl.psite.p, nl.pbeg.p, nl.pcnt.p, nl.pS.p,
                                                                     nl.pR.p, nl.prev.p);
    ctx->launches++;
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}

int ipf_build_neighbours(ipf_ctx* ctx, const ChunkDev& ch, double rcut, NeighList& nl) {
    const int64_t ncfg = ch.ncfg, nsites = ch.natoms;
    nl.serial = ctx->nl_serial++;
    nl.nsites = nsites;
    nl.rcut = rcut;
    nl.npairs = 0;
    nl.maxpairs_cfg = 0;
    nl.h_cfg_pair_off.assign(ncfg + 2, 0);
    nl.h_site_off.clear();
    nl.max_nbrs = 0;
    cudaStream_t st = ctx->stream;
    IPF_CHECK(nl.site_off.alloc(ctx, nsites + 1));
    IPF_CHECK(nl.cfg_pair_off.alloc(ctx, ncfg + 2));     // [ncfg+1] = max neighbours of one site
    if (nsites == 0) {
        IPF_CUDA(ctx, cudaMemsetAsync(nl.site_off.p, 0, sizeof(int64_t), st));
        IPF_CUDA(ctx, cudaMemsetAsync(nl.cfg_pair_off.p, 0, sizeof(int64_t) * (ncfg + 2), st));
        return IPF_OK;
    }
    ProfScope prof_(ctx, "neighbour_list");
    ABuf<int32_t> cell_start, cell_atoms, atom_cell, atom_wrap;
    ABuf<GridInfo> grid;
    ABuf<int64_t> counts, sub_off;
    IPF_CHECK(cell_start.alloc(ctx, (size_t)ncfg * (NCELL_MAX + 1)));
    IPF_CHECK(cell_atoms.alloc(ctx, nsites));
    IPF_CHECK(atom_cell.alloc(ctx, nsites));
    IPF_CHECK(atom_wrap.alloc(ctx, 3 * nsites));
    IPF_CHECK(grid.alloc(ctx, ncfg));
    IPF_CHECK(counts.alloc(ctx, NL_SUB * nsites + 1));
    IPF_CHECK(sub_off.alloc(ctx, NL_SUB * nsites + 1));
    const int32_t* site_cfg = ch.site_cfg.p;
    nl_bin_kernel<<<(unsigned)ncfg, 128, 0, st>>>(ch.atom_off.p, ch.pos.p, ch.cell.p, ch.pbc.p, rcut,
                                                  grid.p, cell_start.p, cell_atoms.p, atom_cell.p, atom_wrap.p);
    ctx->launches++;
    const double rc2 = rcut * rcut;
    const unsigned nb_sites = (unsigned)((NL_SUB * nsites + 127) / 128);
    IPF_CUDA(ctx, cudaMemsetAsync(counts.p + NL_SUB * nsites, 0, sizeof(int64_t), st));
    IPF_CUDA(ctx, cudaMemsetAsync(nl.cfg_pair_off.p + ncfg + 1, 0, sizeof(int64_t), st));
    nl_pairs_kernel<false><<<nb_sites, 128, 0, st>>>(nsites, site_cfg, ch.atom_off.p, ch.pos.p, ch.cell.p, rc2,
                                                     grid.p, cell_start.p, cell_atoms.p, atom_cell.p, atom_wrap.p,
                                                     counts.p, nullptr, nullptr, nullptr, nullptr);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.p, sub_off.p, (int)(NL_SUB * nsites + 1), st);
    void* tmp = nullptr;
    IPF_CHECK(ipf_arena_alloc(ctx, tmp_bytes, &tmp));
    IPF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts.p, sub_off.p, (int)(NL_SUB * nsites + 1), st));
    nl_site_off_kernel<<<(unsigned)((nsites + 256) / 256), 256, 0, st>>>(nsites, sub_off.p, nl.site_off.p,
                                                                        (unsigned long long*)(nl.cfg_pair_off.p + ncfg + 1));
    ctx->launches += 2;
    // per-config pair offsets (ncfg+1 values) to the host; gives the total too
    cfg_offsets_kernel<<<(unsigned)((ncfg + 256) / 256), 256, 0, st>>>(ncfg, ch.atom_off.p, nl.site_off.p, nl.cfg_pair_off.p);
    ctx->launches++;
    IPF_CUDA(ctx, cudaMemcpyAsync(nl.h_cfg_pair_off.data(), nl.cfg_pair_off.p, sizeof(int64_t) * (ncfg + 2),
                                  cudaMemcpyDeviceToHost, st));
    if (ctx->profiling) {
        nl.h_site_off.assign(nsites + 1, 0);
        IPF_CUDA(ctx, cudaMemcpyAsync(nl.h_site_off.data(), nl.site_off.p, sizeof(int64_t) * (nsites + 1),
                                      cudaMemcpyDeviceToHost, st));
    }
    IPF_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t npairs = nl.h_cfg_pair_off[ncfg];
    nl.max_nbrs = (int32_t)nl.h_cfg_pair_off[ncfg + 1];
    nl.npairs = npairs;
    int64_t mx = 0;
    for (int64_t c = 0; c < ncfg; ++c) mx = std::max<int64_t>(mx, nl.h_cfg_pair_off[c + 1] - nl.h_cfg_pair_off[c]);
    nl.maxpairs_cfg = (int32_t)mx;
    IPF_CHECK(nl.pi.alloc(ctx, npairs));
    IPF_CHECK(nl.pj.alloc(ctx, npairs));
    IPF_CHECK(nl.psite.alloc(ctx, npairs));
    IPF_CHECK(nl.pbeg.alloc(ctx, npairs));
    IPF_CHECK(nl.pcnt.alloc(ctx, npairs));
    IPF_CHECK(nl.pS.alloc(ctx, 3 * npairs));
    IPF_CHECK(nl.pR.alloc(ctx, 3 * npairs));
    IPF_CHECK(nl.prev.alloc(ctx, npairs));
    if (npairs == 0) return IPF_OK;
    ABuf<unsigned long long> keys, keys_sorted;
    ABuf<double> Rtmp;
    IPF_CHECK(keys.alloc(ctx, npairs));
    IPF_CHECK(keys_sorted.alloc(ctx, npairs));
    IPF_CHECK(Rtmp.alloc(ctx, 3 * npairs));
    nl_pairs_kernel<true><<<nb_sites, 128, 0, st>>>(nsites, site_cfg, ch.atom_off.p, ch.pos.p, ch.cell.p, rc2,
                                                    grid.p, cell_start.p, cell_atoms.p, atom_cell.p, atom_wrap.p,
                                                    nullptr, sub_off.p, keys.p, Rtmp.p, nullptr);
    const unsigned nb_pairs = (unsigned)((npairs + 255) / 256);
    nl_sort_kernel<<<nb_pairs, 256, 0, st>>>(npairs, nsites, nl.site_off.p, site_cfg, ch.atom_off.p, keys.p,
                                             Rtmp.p, nl.pi.p, nl.pj.p, nl.psite.p, nl.pbeg.p, nl.pcnt.p, nl.pS.p, nl.pR.p, keys_sorted.p);
    nl_reverse_kernel<<<nb_pairs, 256, 0, st>>>(npairs, nl.site_off.p, site_cfg, ch.atom_off.p, nsites,
                                                nl.pi.p, nl.pj.p, nl.pS.p, keys_sorted.p, nl.prev.p);
    ctx->launches += 3;
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
