# IPFittingB200.jl -- ccall shim that puts libipf_b200.so behind IPFitting.jl's own API.
#
# NOT EXECUTED in the build container (no Julia toolchain there); written against
# include/ipf_b200.h and the reference sources it patches:
#   (1) src/lsq_db.jl:181-185   tfor_observations(... safe_append! ...)      -> ipf_assemble
#   (2) src/lsq.jl:306-307, 450-473, 593   weights, P-scaling, qr!, cond, \  -> ipf_solve
#   (3) src/lsq.jl:642-648  add_fits_serial! + rmse                          -> ipf_residual_stats
# Everything else (Dict weights, row allocation, combine, fitinfo) is the unchanged reference.
module IPFittingB200

using IPFitting, JuLIP, LinearAlgebra
using IPFitting: Dat, LsqDB, observations, observation, weighthook, err_weighthook
using IPFitting.DB: _alloc_lsq_matrix, matrows
using IPFitting.Lsq: collect_observations, _fix_weights!

const LIB = get(ENV, "IPF_B200_LIB", "libipf_b200.so")
const MAX_VARS = 10

struct BlockSpec            # mirrors ipf_block_spec
   N::Int32; nb::Int32; nmono::Int32; transform::Int32
   a::Float64; r0::Float64; rc1::Float64; rc2::Float64
   mono_b::Ptr{Int32}; mono_exp::Ptr{UInt8}
   zc::Int32; zs::NTuple{4, Int32}     # species-resolved run: centre species, sorted neighbour species (zc = 0: any)
   env_t::Int32; env_rc1::Float64; env_rc2::Float64   # environment factor n_i^env_t, n_i = sum_j CosCut(env_rc1, env_rc2)(r_ij)
end

mutable struct Ctx
   h::Ptr{Cvoid}
   function Ctx(device::Integer = 0)
      ref = Ref{Ptr{Cvoid}}(C_NULL)
      rc = ccall((:ipf_init, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, ref)
      ctx = new(ref[])
      rc == 0 || error(lasterror(ctx))
      finalizer(c -> ccall((:ipf_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
   end
end
lasterror(ctx::Ctx) = unsafe_string(ccall((:ipf_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
check(ctx::Ctx, rc) = (rc == 0 || error(lasterror(ctx)); nothing)   # NaN codes carry the reference's messages

const CTX = Ref{Union{Nothing, Ctx}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Ctx(parse(Int, get(ENV, "LOCAL_RANK", "0")))); CTX[])

"""
`basisdict(basis)`: the serialised form of a bond-angle polynomial basis,

    Dict("__id__" => "IPFittingB200.NBodyBasis",
         "descs"  => [Dict("kind" => 0|1, "a" => a, "r0" => r0, "rc1" => rc1, "rc2" => rc2), ...],
         "polys"  => [Dict("N" => N, "desc" => index into descs (0-based), "orbit" => [[e_1, ..., e_nv], ...]), ...])

one entry of "polys" per basis function (column of Ψ), "orbit" = the exponent tuples of its S_{N-1} orbit over the
variables (x_1..x_{N-1}, w_12, w_13, .., w_23, ..), optional "zc" / "zs" (centre species and sorted neighbour species of a
species-resolved function); kind 0: x = exp(-a (r/r0 - 1)), 1: x = (r0/r)^a; cutoff
CosCut(rc1, rc2).  It is what `NBodyBasis.write_dict` of ipfitting.jl_b200/basis.py emits and what the `basis` entry of
the `_info.json` written by either side holds.  For an NBodyIPs basis the method below reads the same information from
the basis elements: `bodyorder(b)`, the descriptor (`b.D.transform`, `b.D.cutoff`) and the exponent tuples of the
invariant polynomial (`b.t`, NBodyIPs `NBPoly`); a basis that is already such a Dict passes through.
"""
basisdict(basis::AbstractDict) = basis
function basisdict(basis::AbstractVector)
   descs = Any[]; polys = Any[]
   for b in basis
      D = b.D                                                # BondAngleDesc(transform, cutoff)
      d = Dict("kind" => transform_kind(D.transform), "a" => transform_a(D.transform), "r0" => transform_r0(D.transform),
               "rc1" => D.cutoff.rc1, "rc2" => D.cutoff.rc2)
      i = findfirst(isequal(d), descs)
      i === nothing && (push!(descs, d); i = length(descs))
      push!(polys, Dict("N" => bodyorder(b), "desc" => i - 1, "orbit" => [collect(Int, t) for t in orbit_exponents(b)]))
   end
   return Dict("__id__" => "IPFittingB200.NBodyBasis", "descs" => descs, "polys" => polys)
end
# The three accessors a maintainer maps onto NBodyIPs' SpaceTransform (the README's "exp(- (r/r0 - 1.0))" is kind 0,
# a = 1; "(r0/r)^p" is kind 1, a = p) and the expansion of an NBPoly's invariant tuples into monomial exponents.
function transform_kind end; function transform_a end; function transform_r0 end; function orbit_exponents end

"""
`blockspecs(basis) -> (specs::Vector{BlockSpec}, keep)`: consecutive basis functions with the same body-order and
descriptor form one `ipf_block_spec` (the C side evaluates a block per kernel pass); `keep` holds the arrays the
specs point into and must be GC-preserved across `ipf_basis_create`.  Same grouping as `NBodyBasis.__init__` /
`_compile` of ipfitting.jl_b200/basis.py.
"""
function blockspecs(basis)
   Bd = basisdict(basis)
   descs, polys = Bd["descs"], Bd["polys"]
   specs = BlockSpec[]; keep = Any[]
   start = 1
   while start <= length(polys)
      p0 = polys[start]
      stop = start
      while stop < length(polys) && polys[stop + 1]["N"] == p0["N"] && polys[stop + 1]["desc"] == p0["desc"] &&
            get(polys[stop + 1], "zc", 0) == get(p0, "zc", 0) && get(polys[stop + 1], "zs", Int[]) == get(p0, "zs", Int[]) &&
            get(polys[stop + 1], "env_t", 0) == get(p0, "env_t", 0) && get(polys[stop + 1], "env_rc", []) == get(p0, "env_rc", [])
         stop += 1
      end
      nv = (p0["N"] - 1) + (p0["N"] - 1) * (p0["N"] - 2) ÷ 2
      mono_b = Int32[]; mono_exp = UInt8[]
      for k in start:stop, m in polys[k]["orbit"]
         length(m) == nv || error("monomial has the wrong number of variables")
         push!(mono_b, Int32(k - start))                        # local basis index, non-decreasing
         append!(mono_exp, UInt8.(m)); append!(mono_exp, zeros(UInt8, MAX_VARS - nv))   # row-major [nmono, MAX_VARS]
      end
      d = descs[p0["desc"] + 1]
      push!(keep, mono_b); push!(keep, mono_exp)
      push!(specs, BlockSpec(Int32(p0["N"]), Int32(stop - start + 1), Int32(length(mono_b)), Int32(d["kind"]),
                             Float64(d["a"]), Float64(d["r0"]), Float64(d["rc1"]), Float64(d["rc2"]),
                             pointer(mono_b), pointer(mono_exp),
                             Int32(get(p0, "zc", 0)), ntuple(i -> Int32(get(get(p0, "zs", Int[]), i, 0)), 4),
                             Int32(get(p0, "env_t", 0)), Float64(get(p0, "env_rc", [0.0, 0.0])[1]), Float64(get(p0, "env_rc", [0.0, 0.0])[2])))
      start = stop + 1
   end
   return specs, keep
end

# ---------------------------------------------------------------- multi-GPU: one Julia process per GPU
"""
`attach_communicator(ctx, rank, nranks, bcast)`: creates the library's NCCL communicator over `nranks` processes (one
GPU each).  `bcast(bytes, root)` is any byte broadcast of the host program (`MPI.Bcast!`, a shared file, ...): it only
carries the 128-byte NCCL unique id of rank 0.  Afterwards `ipf_solve*` on this ctx sums the Gram blocks (and the
residual sums) over all ranks on the ctx stream: every rank assembles the rows of its own configurations and calls
`solve_qr` with them; all ranks receive the same coefficients.  No torch, no MPI inside the library.
"""
function attach_communicator(ctx::Ctx, rank::Integer, nranks::Integer, bcast)
   uid = zeros(UInt8, 128)
   rank == 0 && check(ctx, ccall((:ipf_comm_unique_id, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), ctx.h, uid))
   bcast(uid, 0)
   check(ctx, ccall((:ipf_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx.h, nranks, rank, uid))
end

# ---------------------------------------------------------------- (1) assembly
function IPFitting.DB.LsqDB(dbpath::AbstractString, basis, configs::AbstractVector{Dat};
                            verbose = true, maxnthreads = 0, device = true)
   Ψ = _alloc_lsq_matrix(configs, basis)          # unchanged: assigns rows in Dict key order
   db = LsqDB(basis, configs, Ψ, dbpath)
   ctx = context()
   specs, keep = blockspecs(basis)
   hb = Ref{Ptr{Cvoid}}(C_NULL)
   GC.@preserve keep specs begin
      check(ctx, ccall((:ipf_basis_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{BlockSpec}, Ref{Ptr{Cvoid}}),
                       ctx.h, length(specs), specs, hb))
   end
   ncfg = length(configs)
   atom_off = cumsum(vcat(0, [length(c.at) for c in configs]))
   pos  = reduce(hcat, [mat(positions(c.at)) for c in configs])          # 3 x natoms, column-major = atom-major xyz
   cell = reduce(hcat, [reshape(Matrix(JuLIP.cell(c.at))', 9) for c in configs])   # rows = lattice vectors
   pbc  = reduce(vcat, [UInt8.(collect(JuLIP.pbc(c.at))) for c in configs])
   Z    = reduce(vcat, [Int32.(c.at.Z) for c in configs])
   row(c, k) = haskey(c.D, k) ? Int64(first(matrows(c, k))) : Int64(0)  # 1-based first row, 0 = absent
   rowE = [row(c, "E") for c in configs]; rowF = [row(c, "F") for c in configs]; rowV = [row(c, "V") for c in configs]
   hm = Ref{Ptr{Cvoid}}(C_NULL)
   check(ctx, ccall((:ipf_matrix_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    ctx.h, size(Ψ, 1), size(Ψ, 2), hm))
   check(ctx, ccall((:ipf_assemble, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int32},
                     Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}),
                    ctx.h, hb[], ncfg, Int64.(atom_off), pos, cell, pbc, Z, rowE, rowF, rowV, hm[]))
   # materialise on the host (db.Ψ is a Julia Matrix); keep hm[] in a side table to skip the H2D in lsqfit
   check(ctx, ccall((:ipf_matrix_to_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64),
                    ctx.h, hm[], db.Ψ, size(Ψ, 1)))
   DEVICE_PSI[objectid(db)] = hm[]
   ccall((:ipf_basis_destroy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, hb[])
   dbpath != "" && IPFitting.DB.flush(db)
   return db
end
const DEVICE_PSI = Dict{UInt, Ptr{Cvoid}}()

# ---------------------------------------------------------------- (2) solve: replaces src/lsq.jl:436-473,593
function solve_qr(db::LsqDB, Y::Vector{Float64}, W::Vector{Float64}, Irows::Vector{Int}, Icols, Dinv)
   ctx = context()
   hm = get(DEVICE_PSI, objectid(db), C_NULL)
   if hm == C_NULL                       # db loaded from disk: upload once
      r = Ref{Ptr{Cvoid}}(C_NULL)
      check(ctx, ccall((:ipf_matrix_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}), ctx.h, size(db.Ψ)..., r))
      check(ctx, ccall((:ipf_matrix_from_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), ctx.h, r[], db.Ψ, size(db.Ψ, 1)))
      hm = DEVICE_PSI[objectid(db)] = r[]
   end
   n = Icols isa Colon ? size(db.Ψ, 2) : length(Icols)
   c = zeros(n); R = zeros(n, n); rel = Ref{Float64}(0.0)
   ir = Int64.(Irows .- 1)                                            # 0-based at the ABI
   ic = Icols isa Colon ? C_NULL : Int64.(collect(Icols) .- 1)
   check(ctx, ccall((:ipf_solve, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Int64}, Int64,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                    ctx.h, hm, Y, W, ir, length(ir), ic, n, Dinv === nothing ? C_NULL : Dinv, c, R, rel))
   return c, cond(UpperTriangular(R)), rel[]          # c already multiplied by D_inv (src/lsq.jl:593)
end
# In IPFitting.Lsq.lsqfit the :qr branch becomes
#     Y, W, Icfg = collect_observations(db, weights, Vref)                      (unchanged)
#     Irows = intersect(findall(W .!= 0) |> sort, findall(in.(Icfg, Ref(Itrain))))   (unchanged, src/lsq.jl:277)
#     c, κ, rel_rms = IPFittingB200.solve_qr(db, Y, W, Irows, Ibasis, diag(pinv(P)))
# and, per SURVEY Q8, `c` is scattered into a full-length vector before `combine(db.basis, c)`.

end # module
