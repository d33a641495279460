# FDFDB200.jl -- the `ccall` binding a FDFD.jl maintainer adds to route the assembly + linear-solve hot path to
# libfdfd_b200.so (include/fdfd_b200.h).  Include it after `src/solver/modulation.jl` in `src/FDFD.jl`; with
# ENV["FDFD_SOLVER"] == "b200" (same convention as solver.jl:10) the three L3 entry points below replace the stock
# methods, signatures unchanged.  NOTE: Julia is not installed in the build image of this repository, so this file
# has never been executed there; the identical C symbols are exercised from Python (tests/, ctypes).
module FDFDB200

using ..FDFD: Grid, Device, ModulatedDevice, AbstractDevice, Polarization, TM, TE, FieldTM, FieldTE,
              setup_mode!, Float
import ..FDFD: solve, eigenfrequency

const LIB = get(ENV, "FDFD_B200_LIB", "libfdfd_b200")

struct CGrid            # fdfd_grid_t  (mirrors Grid{2}, src/grid.jl:7-13)
    Nx::Int64; Ny::Int64; Npml_x::Int64; Npml_y::Int64
    x0::Float64; x1::Float64; y0::Float64; y1::Float64; L0::Float64
end
CGrid(g::Grid{2}) = CGrid(g.N[1], g.N[2], g.Npml[1], g.Npml[2], g.bounds[1][1], g.bounds[2][1],
                          g.bounds[1][2], g.bounds[2][2], g.L₀)

mutable struct COpts    # fdfd_solve_opts_t
    solver::Int32; precond::Int32; tol::Float64; maxit::Int32; mg_precision::Int32; mg_cycle::Int32
    mg_wdepth::Int32; mg_nu1::Int32; mg_nu2::Int32; mg_coarse_sweeps::Int32
    mg_beta::Float64; mg_wjac::Float64; mg_wline::Float64; check_every::Int32; verbose::Int32
    mg_shift_growth::Float64; mg_max_levels::Int32; use_graph::Int32; concurrency::Int32; ml_spec::Int32
    COpts() = (o = new(); ccall((:fdfd_default_opts, LIB), Cvoid, (Ref{COpts},), o); o)
end

mutable struct CInfo    # fdfd_info_t
    iters::Int32; flag::Int32; relres::Float64; setup_ms::Float64; solve_ms::Float64; total_ms::Float64
    launches::Int64; restarts::Int32; mg_levels::Int32
    CInfo() = new(0, 0, 0.0, 0.0, 0.0, 0.0, 0, 0, 0)
end

struct CInfoV           # the same 56 bytes as an isbits struct: a Vector{CInfoV} is a contiguous C array of fdfd_info_t
    iters::Int32; flag::Int32; relres::Float64; setup_ms::Float64; solve_ms::Float64; total_ms::Float64
    launches::Int64; restarts::Int32; mg_levels::Int32
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        st = ccall((:fdfd_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                   parse(Int, get(ENV, "FDFD_B200_DEVICE", "0")), C_NULL, CTX)
        st == 0 || error("fdfd_b200: ", unsafe_string(ccall((:fdfd_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    return CTX[]
end
check(st) = st == 0 || error("fdfd_b200: ", unsafe_string(ccall((:fdfd_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
enabled() = haskey(ENV, "FDFD_SOLVER") && lowercase(ENV["FDFD_SOLVER"]) == "b200"

"solve(d::Device, pol) -- drop-in for src/solver/driven.jl:4-59"
function solve_b200(d::Device{2}, pol::Polarization=TM)
    g = CGrid(d.grid); (Nx, Ny) = size(d.grid); Nω = length(d.ω); N = Nx * Ny
    opts = COpts()
    # ONE ccall for the whole sweep (the reference loops `for i in eachindex(d.ω)`, driven.jl:11): the library stages ϵᵣ once and
    # solves up to opts.concurrency frequencies at the same time on its worker streams -- the path bench.py measures.
    # Mode sources depend on ω (driven.jl:15-19): one source per frequency (src_per_omega = 1); otherwise d.src is shared.
    ϵ = ComplexF64.(d.ϵᵣ)                                                     # Array{Complex} is boxed: convert (SURVEY §9)
    permode = length(d.modes) > 0
    srcs = Array{ComplexF64}(undef, Nx, Ny, permode ? Nω : 1)
    if permode
        for i in eachindex(d.ω)
            d.src = zeros(Complex, size(d.grid))                              # driven.jl:15
            for mode in d.modes                                               # mode source stays in Julia (<=100 unknowns)
                setup_mode!(d, TM, d.ω[i], mode.neff, mode.pt, mode.dir, mode.width)
            end
            srcs[:, :, i] = ComplexF64.(d.src)
        end
    else
        srcs[:, :, 1] = ComplexF64.(d.src)
    end
    ωs = Float64.(d.ω)
    out = Array{ComplexF64}(undef, Nx, Ny, 3, Nω); infos = Vector{CInfoV}(undef, Nω)
    GC.@preserve ϵ srcs out ωs infos check(ccall((:fdfd_solve_driven, LIB), Cint,
        (Ptr{Cvoid}, Ref{CGrid}, Cint, Cint, Ptr{Float64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Ref{COpts},
         Ptr{ComplexF64}, Ptr{CInfoV}),
        ctx(), g, Int32(pol), Nω, ωs, ϵ, srcs, permode ? 1 : 0, opts, out, infos))
    fields = pol == TM ? Array{FieldTM}(undef, Nω) : Array{FieldTE}(undef, Nω)
    for i in eachindex(d.ω)
        @info "fdfd_b200: ω[$i]: $(infos[i].iters) iterations, relres $(infos[i].relres), $(infos[i].solve_ms) ms"
        fields[i] = pol == TM ? FieldTM(d.grid, d.ω[i], out[:, :, :, i]) : FieldTE(d.grid, d.ω[i], out[:, :, :, i])  # data.jl:56,71
    end
    Nω == 1 && return fields[1]                                               # driven.jl:57-58
    return fields
end

"solve(d::ModulatedDevice) -- drop-in for src/solver/modulation.jl:35-119"
function solve_b200(d::ModulatedDevice{2})
    g = CGrid(d.grid); (Nx, Ny) = size(d.grid); Nω = length(d.ω); nf = 2 * d.nsidebands + 1
    fields = Array{FieldTM}(undef, Nω, nf); opts = COpts()
    for i in eachindex(d.ω)
        ω = d.ω[i]; ωn = ω .+ d.Ω * (-d.nsidebands:1:d.nsidebands)
        length(d.modes) > 0 && (d.src = zeros(Complex, size(d.grid)))
        for mode in d.modes
            setup_mode!(d, TM, ω, mode.neff, mode.pt, mode.dir, mode.width)
        end
        ϵ = ComplexF64.(d.ϵᵣ); Δϵ = ComplexF64.(d.Δϵᵣ); src = ComplexF64.(d.src)
        out = Array{ComplexF64}(undef, Nx, Ny, 3, nf); info = CInfo()
        GC.@preserve ϵ Δϵ src out check(ccall((:fdfd_solve_modulated, LIB), Cint,
            (Ptr{Cvoid}, Ref{CGrid}, Float64, Float64, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64},
             Ref{COpts}, Ptr{ComplexF64}, Ref{CInfo}),
            ctx(), g, ω, d.Ω, d.nsidebands, d.sharedpml, ϵ, Δϵ, src, opts, out, info))
        for j = 1:nf
            fields[i, j] = FieldTM(d.grid, ωn[j], out[:, :, :, j])
        end
    end
    return fields
end

const WHICH = Dict(:LM => 0, :LR => 1, :SR => 2, :LI => 3, :SI => 4)

"eigenfrequency(d, pol, nev; which) -- drop-in for src/solver/eigen.jl:69-115"
function eigenfrequency_b200(d::AbstractDevice{2}, pol::Polarization, nev::Int; which::Symbol=:LM)
    g = CGrid(d.grid); (Nx, Ny) = size(d.grid)
    ϵ = ComplexF64.(d.ϵᵣ); ω = Array{ComplexF64}(undef, nev); out = Array{ComplexF64}(undef, Nx, Ny, 3, nev)
    opts = COpts(); info = CInfo()
    GC.@preserve ϵ ω out check(ccall((:fdfd_eigenfrequency, LIB), Cint,
        (Ptr{Cvoid}, Ref{CGrid}, Cint, Float64, Cint, Cint, Cint, Ptr{ComplexF64}, Ref{COpts}, Ptr{ComplexF64},
         Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), g, Int32(pol), d.ω[1], nev, WHICH[which], 0, ϵ, opts, ω, out, info))
    F = pol == TM ? FieldTM : FieldTE
    return (ω, [F(d.grid, ω[i], out[:, :, :, i]) for i = 1:nev])
end

# ---- one large grid split into row slabs over several GPUs (include/fdfd_b200.h, "row slabs"; csrc/slab.cu) -------------
# One Julia process per GPU (e.g. MPI.jl or Distributed.jl workers).  Rank 0 makes the 128-byte NCCL id, the host program
# broadcasts it (MPI.Bcast!, a file, ...), every rank creates its communicator and calls solve_slab_b200 with the SAME
# device; each rank gets back its own rows of FieldTM.data.  The halo exchange and the Krylov allreduce run inside the
# library over NCCL (NVLink), never through Julia.
nccl_unique_id() = (id = zeros(UInt8, 128); check(ccall((:fdfd_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)

function slab_comm(id::Vector{UInt8}, nranks::Int, rank::Int)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fdfd_comm_create_nccl, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}), ctx(), nranks, rank, id, c))
    return c[]
end

function slab_rows(g::Grid{2}, nranks::Int, rank::Int)
    y0 = Ref{Int64}(0); n = Ref{Int64}(0)
    check(ccall((:fdfd_slab_rows, LIB), Cint, (Ref{CGrid}, Cint, Cint, Ref{Int64}, Ref{Int64}), CGrid(g), nranks, rank, y0, n))
    return (y0[] + 1):(y0[] + n[])          # 1-based row range owned by `rank`
end

"solve(d, TM) of one frequency with the grid cut into `nranks` row slabs; returns the (Nx, nrows, 3) rows of this rank"
function solve_slab_b200(d::Device{2}, comm::Ptr{Cvoid}, nranks::Int, rank::Int)
    rows = slab_rows(d.grid, nranks, rank); (Nx, _) = size(d.grid)
    ϵ = ComplexF64.(d.ϵᵣ[:, rows]); src = ComplexF64.(d.src[:, rows]); out = Array{ComplexF64}(undef, Nx, length(rows), 3)
    opts = COpts(); info = CInfo()
    GC.@preserve ϵ src out check(ccall((:fdfd_solve_driven_slab, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CGrid}, Float64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{COpts}, Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), comm, CGrid(d.grid), d.ω[1], ϵ, src, opts, out, info))
    @info "fdfd_b200 slab solve" rank iters=info.iters relres=info.relres
    return out
end

"solve(d::ModulatedDevice) of one frequency on row slabs (csrc/slab_multi.cu); returns this rank's (Nx, nrows, 3, nf) rows"
function solve_slab_b200(d::ModulatedDevice{2}, comm::Ptr{Cvoid}, nranks::Int, rank::Int)
    rows = slab_rows(d.grid, nranks, rank); (Nx, _) = size(d.grid); nf = 2 * d.nsidebands + 1
    ϵ = ComplexF64.(d.ϵᵣ[:, rows]); Δϵ = ComplexF64.(d.Δϵᵣ[:, rows]); src = ComplexF64.(d.src[:, rows])
    out = Array{ComplexF64}(undef, Nx, length(rows), 3, nf); opts = COpts(); info = CInfo()
    GC.@preserve ϵ Δϵ src out check(ccall((:fdfd_solve_modulated_slab, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CGrid}, Float64, Float64, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64},
         Ref{COpts}, Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), comm, CGrid(d.grid), d.ω[1], d.Ω, d.nsidebands, d.sharedpml, ϵ, Δϵ, src, opts, out, info))
    return out
end

"eigenfrequency(d, TM, nev; which) on row slabs; returns (ω, this rank's (Nx, nrows, 3, nev) rows of the mode fields)"
function eigenfrequency_slab_b200(d::AbstractDevice{2}, nev::Int, comm::Ptr{Cvoid}, nranks::Int, rank::Int; which::Symbol=:LM)
    rows = slab_rows(d.grid, nranks, rank); (Nx, _) = size(d.grid)
    ϵ = ComplexF64.(d.ϵᵣ[:, rows]); ω = Array{ComplexF64}(undef, nev); out = Array{ComplexF64}(undef, Nx, length(rows), 3, nev)
    opts = COpts(); info = CInfo()
    GC.@preserve ϵ ω out check(ccall((:fdfd_eigenfrequency_slab, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CGrid}, Cint, Float64, Cint, Cint, Cint, Ptr{ComplexF64}, Ref{COpts}, Ptr{ComplexF64},
         Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), comm, CGrid(d.grid), Int32(TM), d.ω[1], nev, WHICH[which], 0, ϵ, opts, ω, out, info))
    return (ω, out)
end

# ---- the reference's own seam: dolinearsolve(A, b, matrixsym) (src/solver/solver.jl:4-41) for callers that assemble their own
# matrix -- the chi-3 outer loops (nonlinear.jl:69,97,120) and the 2-D eigenmode (eigen.jl:32-66).  In solver.jl the maintainer
# adds one branch next to the "pardiso" one (:19):   elseif solver == "b200"; x = FDFDB200.dolinearsolve_b200(A, b)
# BiCGSTAB + Jacobi on a SELL-32 image of A (csrc/linsolve.cu): a compatibility path; the L3 entry points above are the fast ones.
using SparseArrays: SparseMatrixCSC
function dolinearsolve_b200(A::SparseMatrixCSC, b::AbstractVector; tol::Float64=1e-10, maxit::Int=100000)
    n = size(A, 1); n == size(A, 2) == length(b) || error("dolinearsolve_b200: dimension mismatch")
    colptr = Vector{Int64}(A.colptr); rowval = Vector{Int64}(A.rowval); nzval = Vector{ComplexF64}(A.nzval)
    bb = Vector{ComplexF64}(b); x = Vector{ComplexF64}(undef, n)
    opts = COpts(); opts.tol = tol; opts.maxit = maxit; info = CInfo()
    GC.@preserve colptr rowval nzval bb x check(ccall((:fdfd_dolinearsolve_csc, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Ref{COpts}, Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), n, colptr, rowval, nzval, 1, bb, opts, x, info))
    @info "fdfd_b200 dolinearsolve: $(info.iters) iterations, relres $(info.relres), $(info.total_ms) ms"
    return x
end

# With the grid and frequency the matrix was assembled on (nonlinear.jl has both in scope at :58-69 and passes them down to
# _doborn): a matrix that IS the TM operator of that grid -- the first solve and every Born step -- runs on the multigrid path
# (fdfd_dolinearsolve_csc_grid checks A*v against the matrix-free operator on the device first; anything else falls back).
function dolinearsolve_b200(A::SparseMatrixCSC, b::AbstractVector, grid::Grid{2}, ω::Real; tol::Float64=1e-10, maxit::Int=20000)
    n = size(A, 1); n == size(A, 2) == length(b) || error("dolinearsolve_b200: dimension mismatch")
    colptr = Vector{Int64}(A.colptr); rowval = Vector{Int64}(A.rowval); nzval = Vector{ComplexF64}(A.nzval)
    bb = Vector{ComplexF64}(b); x = Vector{ComplexF64}(undef, n)
    opts = COpts(); opts.tol = tol; opts.maxit = maxit; info = CInfo()
    GC.@preserve colptr rowval nzval bb x check(ccall((:fdfd_dolinearsolve_csc_grid, LIB), Cint,
        (Ptr{Cvoid}, Ref{CGrid}, Float64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Ref{COpts},
         Ptr{ComplexF64}, Ref{CInfo}),
        ctx(), CGrid(grid), Float64(ω), n, colptr, rowval, nzval, 1, bb, opts, x, info))
    @info "fdfd_b200 dolinearsolve ($(info.mg_levels > 0 ? "multigrid" : "generic") path): $(info.iters) iterations, relres $(info.relres)"
    return x
end

end # module
