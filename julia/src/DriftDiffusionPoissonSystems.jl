# Drop-in Julia (>= 1.6) front end for the B200 hot path.
#
# Same module name and exported names as the reference package
# (Stivanification/DriftDiffusionPoissonSystems.jl, src/DriftDiffusionPoissonSystems.jl:6-7), so
# `using DriftDiffusionPoissonSystems` keeps working.  The host side here only
#   * reads / writes meshes,
#   * turns `bddata` into Dirichlet node lists and Neumann edge loads,
#   * samples the user's spatial closures once (centroids, edge midpoints, boundary nodes),
# and then hands plain arrays to libddps_b200.so through `ccall` (C ABI: include/ddps.h).  All
# per-iteration work (assembly, elimination, PCG, Newton and Gummel loops) runs on the GPU.
#
# NOTE: no Julia toolchain exists in the build image, so this file could not be executed there; the
# Python mirror (driftdiffusionpoissonsystems.jl_b200/api.py) performs the identical call sequence and
# is what the test-suite drives.  Keep the two in sync.
module DriftDiffusionPoissonSystems

using DelimitedFiles

export Mesh, read_mesh, construct_mesh, construct_mesh4, construct_mesh8, solve_ddpe, solve_semlin_poisson,
       aquire_boundary, SinhRHS, PolyRHS, derivative, calculate_current, aquire_boundary_separate_edges, get_gradients,
       ConstFn, AffineFn, StepX, StepY, comm_unique_id, Comm

const LIB = get(ENV, "DDPS_B200_LIB", joinpath(@__DIR__, "..", "..", "driftdiffusionpoissonsystems.jl_b200", "libddps_b200.so"))

mutable struct Mesh
    nodes::Matrix{Float64}      # 2 x nq
    edges::Matrix{Int64}        # 4 x ne  [n1; n2; phys; geom]
    elements::Matrix{Int64}     # 5 x nme [n1; n2; n3; phys; geom]
end

# ---------------------------------------------------------------- registered u-dependent functors
"f(x,y,u) = h(x,y) + alpha*sinh(beta*u); also an ordinary callable"
struct SinhRHS{F}
    h::F
    alpha::Float64
    beta::Float64
end
SinhRHS(h; alpha = -1.0, beta = 1.0) = SinhRHS(h, Float64(alpha), Float64(beta))
(f::SinhRHS)(x, y, u) = f.h(x, y) + f.alpha * sinh(f.beta * u)

"f(x,y,u) = sum_k c_k(x,y) u^k (degree <= 4)"
struct PolyRHS
    coeffs::Vector{Function}
end
(f::PolyRHS)(x, y, u) = sum(c(x, y) * u^(k - 1) for (k, c) in enumerate(f.coeffs))

derivative(f::SinhRHS) = (x, y, u) -> f.alpha * f.beta * cosh(f.beta * u)
derivative(f::PolyRHS) = (x, y, u) -> sum((k - 1) * c(x, y) * u^(k - 2) for (k, c) in enumerate(f.coeffs) if k > 1; init = 0.0)

# ---------------------------------------------------------------- registered SPATIAL functors (include/ddps.h: ddps_spatial_kind)
# Coefficient closures the library samples on the DEVICE (ddps_coeff_functor) -- no host sampling, no upload.  Each is also
# an ordinary callable (x,y)->value, so user code can evaluate it like the closure it replaces; any other closure is sampled
# on the host (broadcast) and uploaded with ddps_coeff_set.
abstract type SpatialFn <: Function end
struct ConstFn <: SpatialFn; c::Float64; end
struct AffineFn <: SpatialFn; a::Float64; bx::Float64; by::Float64; end
struct StepX <: SpatialFn; x0::Float64; left::Float64; right::Float64; end
struct StepY <: SpatialFn; y0::Float64; below::Float64; above::Float64; end
(f::ConstFn)(x, y) = f.c
(f::AffineFn)(x, y) = (f.a + f.bx * x) + f.by * y
(f::StepX)(x, y) = x > f.x0 ? f.right : f.left
(f::StepY)(x, y) = y > f.y0 ? f.above : f.below
spatial_kind(::ConstFn) = 1; spatial_kind(::AffineFn) = 2; spatial_kind(::StepX) = 3; spatial_kind(::StepY) = 4
spatial_params(f::ConstFn) = [f.c]; spatial_params(f::AffineFn) = [f.a, f.bx, f.by]
spatial_params(f::StepX) = [f.x0, f.left, f.right]; spatial_params(f::StepY) = [f.y0, f.below, f.above]

# ---------------------------------------------------------------- C ABI plumbing (field order = include/ddps.h)
struct SolverOpts
    lin_rtol::Cdouble; newton_eta::Cdouble; lin_max_it::Int64
    max_newton::Int32; max_gummel::Int32; check_every::Int32; warm_start::Int32; verbose::Int32; lin_cap_ok::Int32
    no_alt_traversal::Int32; warm_start_v::Int32; no_sym_scaling::Int32; renumber::Int32
    precond::Int32        # 0 = Jacobi-PCG (default), 1 = CG preconditioned with the smoothed-aggregation multigrid V-cycle
    amg_no_lag::Int32     # 1: recompute the multigrid coarse operators for every assembled operator (default 0: lagged)
    newton_damping::Int32 # k > 0: opt-in backtracking, the Newton step is halved up to k times until ||M*V-b||_inf decreases
    reserved0::Int32
    lin_etol::Cdouble     # error-based (energy-norm) stop tolerance of the linear solves (default 1e-10; < 0: residual rules only)
end

struct Stats
    outer_iters::Int32; newton_total::Int32; pcg_total::Int64; spmv_total::Int64; kernel_launches::Int64
    control::Cdouble; t_setup_ms::Cdouble; t_solve_ms::Cdouble; t_asm_ms::Cdouble; t_pcg_ms::Cdouble; t_wall_ms::Cdouble
    status::Int32; n_negative_area::Int32
    amg_levels::Int32; amg_reuses::Int32; t_amg_setup_ms::Cdouble; newton_halvings::Int32; reserved::NTuple{3,Int32}
end

# Multi-GPU: one Julia process per GPU.  `id` = comm_unique_id() on rank 0, broadcast by the application's own plumbing
# (MPI.jl, Distributed, a file); pass `comm = Comm(id, rank, nranks)` to the solvers.  Every rank passes the same mesh and
# gets the full solution back (the library partitions the nodes by coordinate bisection, include/ddps.h).
struct Comm
    id::Vector{UInt8}     # 128 bytes (ncclUniqueId)
    rank::Int
    nranks::Int
end
function comm_unique_id()
    id = zeros(UInt8, 128)
    ccall((:ddps_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error("ddps: ddps_comm_unique_id failed")
    id
end

lasterr(h) = unsafe_string(ccall((:ddps_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
chk(h, rc) = rc == 0 ? nothing : error("ddps: " * lasterr(h))

# Preconditioner used by every solve of this module (include/ddps.h: DDPS_PRECOND_JACOBI = 0, DDPS_PRECOND_AMG = 1);
# switch with `DriftDiffusionPoissonSystems.PRECOND[] = 1`.
const PRECOND = Ref{Int32}(0)
# Any other field of ddps_solver_opts, by name: `DriftDiffusionPoissonSystems.OPTS[:newton_damping] = 8`, `OPTS[:renumber] = 1`,
# `OPTS[:warm_start_v] = 1`, `OPTS[:lin_etol] = 1e-11`, ... (defaults: ddps_default_opts)
const OPTS = Dict{Symbol,Any}()

function with_ctx(f; device = 0, verbose = true, comm::Union{Nothing,Comm} = nothing)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ddps_ctx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
    rc == 0 || error("ddps: " * lasterr(C_NULL))
    h = ref[]
    try
        o = Ref{SolverOpts}()
        ccall((:ddps_default_opts, LIB), Cvoid, (Ref{SolverOpts},), o)
        v = o[]
        val(name) = name == :verbose ? Int32(verbose) : name == :precond ? PRECOND[] : get(OPTS, name, getfield(v, name))
        o[] = SolverOpts((convert(fieldtype(SolverOpts, n), val(n)) for n in fieldnames(SolverOpts))...)
        chk(h, ccall((:ddps_set_opts, LIB), Cint, (Ptr{Cvoid}, Ref{SolverOpts}), h, o))
        # the communicator must exist before the mesh is set (the mesh is partitioned across the ranks there)
        comm === nothing || chk(h, ccall((:ddps_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), h, comm.id, comm.rank, comm.nranks))
        return f(h)
    finally
        ccall((:ddps_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    end
end

set_mesh(h, m::Mesh) = chk(h, ccall((:ddps_mesh_set, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Int64}, Int64),
                                    h, m.nodes, size(m.nodes, 2), m.elements, size(m.elements, 2)))
set_dirichlet(h, field, n, gD) = chk(h, ccall((:ddps_dirichlet_set, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Cdouble}, Int64),
                                             h, field, Vector{Int64}(n), Vector{Float64}(gD), length(n)))
set_coeff(h, slot, a) = (v = vec(Array{Float64}(a)); chk(h, ccall((:ddps_coeff_set, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64), h, slot, v, length(v))))
set_functor(h, kind, p) = chk(h, ccall((:ddps_functor_set, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint), h, kind, Vector{Float64}(p), length(p)))
# coefficient slot from a closure: registered spatial functors are sampled by the library on the device, anything else is
# broadcast here over the sampling points (xs, ys) like the reference does (src:254,467,341-343) and uploaded
set_coeff_fn(h, slot, fn::SpatialFn, xs, ys; square = false) =
    (p = spatial_params(fn); chk(h, ccall((:ddps_coeff_functor, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint), h, slot, spatial_kind(fn), p, length(p))))
set_coeff_fn(h, slot, fn, xs, ys; square = false) = set_coeff(h, slot, square ? fn.(xs, ys) .^ 2 : fn.(xs, ys))

# ---------------------------------------------------------------- host-side preprocessing
function split_bddata(bddata)
    size(bddata, 1) == 3 || error("bddata must be 3 x n")
    neu = findall(==('N'), bddata[2, :]); diri = findall(==('D'), bddata[2, :])
    length(neu) + length(diri) == size(bddata, 2) ||
        error("Could not understand boundary file. Not all parts of the boundary have BC assigned.")
    neu, diri
end

function dirichlet_nodes(mesh::Mesh, bddata)
    _, diri = split_bddata(bddata)
    nodes = Int64[]; vals = Float64[]
    for j in diri
        sel = findall(==(bddata[1, j]), mesh.edges[4, :])
        seg = unique(vec(mesh.edges[1:2, sel]))
        length(seg) == length(sel) + 1 || error("boundary segment $(bddata[1, j]) is not a single open polyline")
        append!(nodes, seg)
        append!(vals, Float64[bddata[3, j](mesh.nodes[1, k], mesh.nodes[2, k]) for k in seg])
    end
    keep = unique(i -> (nodes[i], vals[i]), eachindex(nodes))
    nodes[keep], vals[keep]
end

function neumann_edges(mesh::Mesh, bddata)
    neu, _ = split_bddata(bddata)
    n1 = Int64[]; n2 = Int64[]; load = Float64[]
    for j in neu
        for e in findall(==(bddata[1, j]), mesh.edges[4, :])
            a, b = mesh.edges[1, e], mesh.edges[2, e]
            pa, pb = mesh.nodes[:, a], mesh.nodes[:, b]
            push!(n1, a); push!(n2, b)
            push!(load, hypot(pa[1] - pb[1], pa[2] - pb[2]) * bddata[3, j]((pa[1] + pb[1]) / 2, (pa[2] + pb[2]) / 2) / 2)
        end
    end
    n1, n2, load
end

function aquire_boundary(mesh::Mesh, bddata)
    n1, n2, _ = neumann_edges(mesh, bddata)
    n, gD = dirichlet_nodes(mesh, bddata)
    n1, n2, gD, n
end

corner(mesh, k) = mesh.nodes[:, mesh.elements[k, :]]
centroids(mesh) = (corner(mesh, 1) .+ corner(mesh, 2) .+ corner(mesh, 3)) ./ 3
function midpoints(mesh)   # 3 x nme per coordinate, order m12, m23, m31
    q = (corner(mesh, 1), corner(mesh, 2), corner(mesh, 3))
    mx = permutedims(hcat((q[1][1, :] .+ q[2][1, :]) ./ 2, (q[2][1, :] .+ q[3][1, :]) ./ 2, (q[3][1, :] .+ q[1][1, :]) ./ 2))
    my = permutedims(hcat((q[1][2, :] .+ q[2][2, :]) ./ 2, (q[2][2, :] .+ q[3][2, :]) ./ 2, (q[3][2, :] .+ q[1][2, :]) ./ 2))
    mx, my
end

# ---------------------------------------------------------------- public solvers
function solve_ddpe(mesh::Mesh, rec_mode::Symbol, Vbddata::Array, ubddata::Array, vbddata::Array, lambda::Function,
                    delta::Float64, tau_p::Float64, tau_n::Float64, mu_p::Function, mu_n::Function, C::Function,
                    tol = 10.0^(-9); device = 0, verbose = true, comm = nothing)
    rec_mode in (:shockley, :zero) || error("rec_mode must be :shockley or :zero")
    n, gD = dirichlet_nodes(mesh, Vbddata)
    _, gDu = dirichlet_nodes(mesh, ubddata)
    _, gDv = dirichlet_nodes(mesh, vbddata)
    (length(gDu) == length(n) && length(gDv) == length(n)) || throw(DimensionMismatch("u/v bddata must mark the same segments 'D' as the V bddata"))
    c = centroids(mesh); mx, my = midpoints(mesh)
    nq = size(mesh.nodes, 2)
    V = zeros(nq); u = zeros(nq); v = zeros(nq)
    with_ctx(; device = device, verbose = verbose, comm = comm) do h
        set_mesh(h, mesh)
        set_dirichlet(h, 0, n, gD); set_dirichlet(h, 1, n, gDu); set_dirichlet(h, 2, n, gDv)
        set_coeff_fn(h, 0, lambda, c[1, :], c[2, :]; square = true)      # the library squares a device-sampled lambda itself
        set_coeff_fn(h, 1, mu_p, c[1, :], c[2, :]); set_coeff_fn(h, 2, mu_n, c[1, :], c[2, :])
        set_coeff_fn(h, 3, C, mx, my)
        st = Ref{Stats}()
        chk(h, ccall((:ddps_solve_ddpe, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Stats}),
                     h, rec_mode == :shockley ? 1 : 0, delta, tau_p, tau_n, tol, V, u, v, st))
    end
    V, u, v
end

function solve_semlin_poisson(mesh, A::Array, bddata::Array, f, g, tol = 10.0^(-14); device = 0, verbose = true, comm = nothing)
    f isa Union{SinhRHS,PolyRHS} || error("f must be a registered functor (SinhRHS / PolyRHS): arbitrary closures cannot run on the device and there is no CPU fallback")
    n, gD = dirichlet_nodes(mesh, bddata)
    n1, n2, load = neumann_edges(mesh, bddata)
    nq = size(mesh.nodes, 2)
    nl = zeros(nq)
    for (a, b, l) in zip(n1, n2, load); nl[a] += l; nl[b] += l; end
    mx, my = midpoints(mesh)
    tag = float.(mesh.elements[5, :])
    V = zeros(nq)
    with_ctx(; device = device, verbose = verbose, comm = comm) do h
        set_mesh(h, mesh)
        set_coeff(h, 4, Float64[A[1](t) for t in tag]); set_coeff(h, 5, Float64[A[4](t) for t in tag])
        set_dirichlet(h, 3, n, gD)
        set_coeff(h, 7, nl)
        if f isa SinhRHS
            set_coeff(h, 6, f.h.(mx, my)); set_functor(h, 1, [f.alpha, f.beta])
        else
            for (k, c) in enumerate(f.coeffs)
                set_coeff(h, 16 + k - 1, c.(mx, my)); set_coeff(h, 32 + k - 1, c.(mesh.nodes[1, :], mesh.nodes[2, :]))
            end
            set_functor(h, 2, [length(f.coeffs) - 1])
        end
        st = Ref{Stats}()
        chk(h, ccall((:ddps_solve_semlin, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Ref{Stats}), h, tol, V, st))
    end
    V
end

# ---------------------------------------------------------------- terminal current (host post-processing)
function aquire_boundary_separate_edges(mesh::Mesh, Endpoints::Array, bddata::Array)
    n1, n2, gD, n = aquire_boundary(mesh, bddata)
    _, diri = split_bddata(bddata)
    idx = reduce(vcat, [findall(==(bddata[1, j]), mesh.edges[4, :]) for j in diri]; init = Int[])
    X(e, k) = mesh.nodes[1, mesh.edges[k, e]]; Y(e, k) = mesh.nodes[2, mesh.edges[k, e]]
    xpar = [e for e in idx if isapprox(Y(e, 1) - Y(e, 2), 0)]
    ypar = [e for e in idx if isapprox(X(e, 1) - X(e, 2), 0)]
    length(xpar) + length(ypar) == length(idx) || error("Dirichlet edges must be parallel to the axes")
    left = [e for e in ypar if isapprox(Endpoints[1, 1], X(e, 1))]; right = [e for e in ypar if isapprox(Endpoints[3, 1], X(e, 1))]
    bottom = [e for e in xpar if isapprox(Endpoints[1, 2], Y(e, 1))]; top = [e for e in xpar if isapprox(Endpoints[3, 2], Y(e, 1))]
    length(left) + length(right) + length(bottom) + length(top) == length(idx) || error("Dirichlet edges must lie on the rectangle")
    n1, n2, gD, n, mesh.edges[:, left], mesh.edges[:, bottom], mesh.edges[:, right], mesh.edges[:, top]
end

function get_gradients(uu::Array{Float64,2}, vv::Array{Float64,2}, ww::Array{Float64,2}, ar::Array{Float64,1})
    [0.5 .* uu[2, :] ./ ar 0.5 .* vv[2, :] ./ ar 0.5 .* ww[2, :] ./ ar], [-0.5 .* uu[1, :] ./ ar -0.5 .* vv[1, :] ./ ar -0.5 .* ww[1, :] ./ ar]
end

function calculate_current(mesh::Mesh, rec_mode::Symbol, Endpoints::Array, Vbddata::Array, ubddata::Array, vbddata::Array,
                           lambda::Function, delta::Float64, tau_p::Float64, tau_n::Float64, mu_p::Function, mu_n::Function,
                           C::Function, tol = 10.0^(-9); kwargs...)
    V, u, v = solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lambda, delta, tau_p, tau_n, mu_p, mu_n, C, tol; kwargs...)
    _, _, _, _, eL, eB, eR, eT = aquire_boundary_separate_edges(mesh, Endpoints, Vbddata)
    q1, q2, q3 = corner(mesh, 1), corner(mesh, 2), corner(mesh, 3)
    uu, vv, ww = q2 .- q3, q3 .- q1, q1 .- q2
    ar = (uu[1, :] .* vv[2, :] .- uu[2, :] .* vv[1, :]) ./ 2
    gradx, grady = get_gradients(uu, vv, ww, ar)
    first = Dict{Tuple{Int64,Int64},Int}()                      # edge -> first triangle (replaces the O(ne*nme) findfirst)
    for e in size(mesh.elements, 2):-1:1, (a, b) in ((1, 2), (2, 3), (1, 3))
        p, q = minmax(mesh.elements[a, e], mesh.elements[b, e]); first[(p, q)] = e
    end
    gl = (-sqrt(3 / 5), 0.0, sqrt(3 / 5)); wt = (5 / 9, 8 / 9, 5 / 9)
    I = 0.0
    for (edges, fixed, alongx, g, sgn) in ((eL, Endpoints[1, 1], false, gradx, -1.0), (eB, Endpoints[1, 2], true, grady, -1.0),
                                           (eR, Endpoints[3, 1], false, gradx, 1.0), (eT, Endpoints[3, 2], true, grady, 1.0))
        for i in 1:size(edges, 2)
            Vb = Vbddata[3, edges[4, i]]
            c = mesh.nodes[alongx ? 1 : 2, edges[1:2, i]]; a, b = minimum(c), maximum(c)
            ind = first[minmax(edges[1, i], edges[2, i])]
            gu = sum(u[mesh.elements[1:3, ind]] .* g[ind, :]); gv = sum(v[mesh.elements[1:3, ind]] .* g[ind, :])
            for (t, w) in zip(gl, wt)
                s = (t + 1) * (b - a) / 2 + a
                x, y = alongx ? (s, fixed) : (fixed, s)
                I += sgn * w * (mu_n(x, y) * exp(Vb(x, y)) * gu - mu_p(x, y) * exp(-Vb(x, y)) * gv) * (b - a) / 2
            end
        end
    end
    I, V, u, v
end

# ---------------------------------------------------------------- mesh I/O (host only)
function read_mesh(path)
    raw = readdlm(path)
    r = 1
    tok(i) = raw[i, 1]
    tok(r) == "\$MeshFormat" || error("not a gmsh file"); raw[r + 1, 1] == 2.2 || error("only MSH 2.2"); r += 3
    if tok(r) == "\$PhysicalNames"; r += Int(raw[r + 1, 1]) + 3; end
    tok(r) == "\$Nodes" || error("missing \$Nodes")
    nq = Int(raw[r + 1, 1])
    all(raw[r + 2:r + 1 + nq, 4] .== 0) || error("mesh is not planar")
    nodes = permutedims(Float64.(raw[r + 2:r + 1 + nq, 2:3])); r += nq + 3
    tok(r) == "\$Elements" || error("missing \$Elements")
    nall = Int(raw[r + 1, 1]); rows = r + 2:r + 1 + nall
    kinds = Int.(raw[rows, 2])
    all(k -> k == 1 || k == 2, kinds) || error("Unknown type of elements used! Please only use edges and triangular elements!")
    all(raw[rows, 3] .== 2) || error("elements must carry exactly two tags")
    le = rows[kinds .== 1]; lt = rows[kinds .== 2]
    edges = permutedims(Int64.(raw[le, [6, 7, 4, 5]])); elements = permutedims(Int64.(raw[lt, [6, 7, 8, 4, 5]]))
    r += nall + 3
    r <= size(raw, 1) && @warn "There are lines in the mesh file that are not needed and cannot be taken into account."
    Mesh(nodes, edges, elements)
end

function write_geo(pts, h, meshname)
    np = length(pts)
    open("$meshname.geo", "w") do io
        println(io, "h = $h;\n ")
        for (k, (x, y)) in enumerate(pts); println(io, "Point($k)={$x, $y, 0, h};"); end
        for k in 1:np; println(io, "Line($k) = {$k, $(k % np + 1)};"); end
        ids = join(1:np, ",")
        println(io, "Line Loop(1) = {$ids};\nPlane Surface(2) ={1};\nPhysical Surface(2) = {2};")
        println(io, "Physical Line(1) = {$(join(1:np, ", "))};")
        print(io, "Coherence;")
    end
end

construct_mesh4(E::Array, h::Float64, meshname::AbstractString) = write_geo([(E[k, 1], E[k, 2]) for k in 1:4], h, meshname)
construct_mesh8(E::Array, D::Array, h::Float64, meshname::AbstractString) =
    write_geo([(E[1, 1], E[1, 2]), (E[2, 1], E[2, 2]), (E[2, 1], D[2, 1]), (E[2, 1], D[2, 2]), (E[3, 1], E[3, 2]),
               (E[4, 1], E[4, 2]), (E[4, 1], D[1, 2]), (E[4, 1], D[1, 1])], h, meshname)
const construct_mesh = construct_mesh8   # README/test scripts call this name

end # module
