# Weno5GPU.jl -- Julia host side of libweno5mhd (B200 / sm_100a).
#
# Drop-in for the hot path of jdonnert/julia's Weno5_2D.jl / Weno5_1D.jl: the functions below
# keep the reference's names, argument meaning and array layout
#     q   :: Array{Float64,3} (Nx,Ny,9)  [rho,Mx,My,Mz,Bx,By,Bz,S,E]      (Weno5_2D.jl:35,2006)
#     bxb, byb :: Array{Float64,2} (Nx,Ny)  face-centred Bx / By            (Weno5_2D.jl:37,1669)
#     1-D: q :: Array{Float64,2} (N,9)                                        (Weno5_1D.jl:37)
# and forward to the C ABI of include/weno5mhd.h with `ccall`.  Julia arrays are column-major
# Float64, which is exactly the layout the library expects, so pointers are passed unchanged.
#
# The reference reads its configuration from module-level constants (Weno5_2D.jl:7-29); here it
# is an explicit `Weno5Params` value.  Written for Julia >= 1.0 (the reference itself is Julia 0.6
# and does not parse on a current Julia).  There is no `julia` binary in the build container, so
# this file is exercised only where Julia exists; its Python twin julia_b200/solver.py binds the
# same entry points and is what the test-suite drives.
module Weno5GPU

export Weno5Params, Solver, upload!, download, boundaries!, timestep, classical_RK4_step, set_es_switch!, es_switch_count, es_switch_mask, comm_halo_mode,
       pinned_array, free_pinned, pipe_overlapped, arr2img,
       classical_RK4_step!, advance!, compute_divB, compute_global_vmax, interpolateB_center2face,
       state2conserved, WENO5_2D, WENO5_1D, Pipe, pipe_step!,
       Weno5Params3D, Solver3D, WENO5_3D

const NBnd = 3                                    # Weno5_2D.jl:9
const BC_OUTFLOW, BC_PERIODIC = Cint(0), Cint(1)

const libweno5 = get(ENV, "WENO5MHD_LIB", joinpath(@__DIR__, "..", "libweno5mhd.so"))

# mirrors `weno5_params` (include/weno5mhd.h)
struct Weno5Params
    nx::Cint; ny::Cint
    dx::Cdouble; dy::Cdouble
    gamma::Cdouble; cfl::Cdouble; eps::Cdouble
    bc_x::Cint; bc_y::Cint
    es_roundtrip::Cint
end

"Reference defaults: gam = 5/3, courFac = 0.8, eps = 1e-6 (Weno5_2D.jl:25-29), outflow boundaries."
Weno5Params(nx, ny; xsize=1.0, ysize=1.0, gamma=5/3, cfl=0.8, eps=1e-6,
            bc_x=BC_OUTFLOW, bc_y=BC_OUTFLOW, es_roundtrip=true) =
    Weno5Params(nx, ny, xsize/nx, ny > 1 ? ysize/ny : 1.0, gamma, cfl, eps, bc_x, bc_y, es_roundtrip ? 1 : 0)

struct Weno5Error <: Exception
    code::Int
    msg::String
end

function check(rc::Integer)
    rc == 0 && return nothing
    throw(Weno5Error(rc, unsafe_string(ccall((:weno5_last_error, libweno5), Cstring, ()))))
end

"One GPU, one y-slab; the state stays resident in HBM between calls."
mutable struct Solver
    h::Ptr{Cvoid}
    p::Weno5Params
    ny_local::Int
    function Solver(p::Weno5Params; device::Integer=0, ny_local::Integer=p.ny, j_offset::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:weno5_create, libweno5), Cint, (Ref{Weno5Params}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                    p, device, ny_local, j_offset, h))
        s = new(h[], p, ny_local)
        finalizer(x -> (x.h != C_NULL && ccall((:weno5_destroy, libweno5), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), s)
        return s
    end
end

is1d(s::Solver) = s.p.ny == 1
dims(s::Solver) = (Int(s.p.nx) + 2NBnd, is1d(s) ? 1 : s.ny_local + 2NBnd)

# ---- multi-GPU (one Julia process per GPU; ship the id with Distributed / MPI)
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:weno5_comm_unique_id, libweno5), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(s::Solver, rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ccall((:weno5_comm_init, libweno5), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), s.h, rank, nranks, id))
"0: no communicator, 1: NCCL send/recv + all-reduce, 2: ghost rows stored straight into the neighbours' planes over peer memory, 3: also the dt all-reduce over peer memory."
comm_halo_mode(s::Solver) = Int(ccall((:weno5_comm_halo_mode, libweno5), Cint, (Ptr{Cvoid},), s.h))

# ---- state transfer
function upload!(s::Solver, q::Array{Float64,3}, bxb::Array{Float64,2}, byb::Array{Float64,2})
    @assert size(q) == (dims(s)..., 9) && size(bxb) == dims(s) && size(byb) == dims(s)
    check(ccall((:weno5_upload, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s.h, q, bxb, byb))
end
"q only: the face fields are derived on the device as interpolateB_center2face does (Weno5_2D.jl:1669)."
upload!(s::Solver, q::Array{Float64,3}) =
    check(ccall((:weno5_upload, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s.h, q, C_NULL, C_NULL))
upload!(s::Solver, q::Array{Float64,2}) =
    check(ccall((:weno5_upload_1d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, q))

function download(s::Solver)
    Nx, Ny = dims(s)
    if is1d(s)
        q = Array{Float64}(undef, Nx, 9)
        check(ccall((:weno5_download_1d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, q))
        return q
    end
    q = Array{Float64}(undef, Nx, Ny, 9); bxb = Array{Float64}(undef, Nx, Ny); byb = similar(bxb)
    check(ccall((:weno5_download, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s.h, q, bxb, byb))
    return q, bxb, byb
end

# ---- the hot path (reference names)
"boundaries!(q, bxb, byb) on the resident state (Weno5_2D.jl:1684)."
boundaries!(s::Solver) = check(ccall((:weno5_boundaries, libweno5), Cint, (Ptr{Cvoid},), s.h))

"timestep(q) -> (dt, vxmax, vymax) (Weno5_2D.jl:1930); 1-D: dt = courFac*dx/vmax (Weno5_1D.jl:55)."
function timestep(s::Solver)
    dt = Ref(0.0); vx = Ref(0.0); vy = Ref(0.0)
    check(ccall((:weno5_timestep, libweno5), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ref{Float64}), s.h, dt, vx, vy))
    return dt[], vx[], vy[]
end
compute_global_vmax(s::Solver) = (r = timestep(s); is1d(s) ? r[2] : (r[2], r[3]))

"classical_RK4_step on the resident state (Weno5_2D.jl:125 / Weno5_1D.jl:152)."
classical_RK4_step!(s::Solver, dt::Real) = check(ccall((:weno5_rk4_step, libweno5), Cint, (Ptr{Cvoid}, Cdouble), s.h, dt))

"""
Dual-energy mode: run the E sweeps and the entropy-variable S sweeps (weno5_flux_difference_S, Weno5_2D.jl:361)
every stage and let ES_Switch (:1734) choose per cell with the criterion the reference wrote and disabled
(:1740): |S(E) - S| >= threshold takes the E state.  `enable = false` (0) is the shipped 2-D behaviour.
1-D grids: `enable = 2` is Weno5_1D.jl as shipped -- both branches swept, energy state kept (:252-254), the flags
(`idx`, :245) available through `es_switch_mask`; `enable = true` (1) drops that override.
"""
set_es_switch!(s::Solver, enable::Union{Bool,Integer}=true; threshold::Real=1e-5) =
    check(ccall((:weno5_set_es_switch, libweno5), Cint, (Ptr{Cvoid}, Cint, Cdouble), s.h, Cint(enable), threshold))
"1.0 where |S(E) - S| >= threshold in the last stage; `find(mask .> 0)` is the reference's `idx` (Weno5_1D.jl:245)."
function es_switch_mask(s::Solver)
    Nx, Ny = dims(s)
    mask = is1d(s) ? zeros(Nx) : zeros(Nx, Ny)
    check(ccall((:weno5_es_switch_mask, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, mask))
    return mask
end
function es_switch_count(s::Solver)
    n = Ref{Culonglong}(0)
    check(ccall((:weno5_es_switch_count, libweno5), Cint, (Ptr{Cvoid}, Ref{Culonglong}), s.h, n))
    return Int(n[])
end

"""
Arr2Img(q[:,:,component+1], cmap; range, log) (JColorMaps.jl:92) evaluated on the resident state: the colour
indices come from the GPU, the lookup `cmap[rimg]` happens here.  component 0..8 = q's components in the
reference's order, 9 = bxb, 10 = byb, 11 = compute_divB.  Returns (img, (lo, hi)).
"""
function arr2img(s::Solver, component::Integer, cmap::AbstractVector; range=(0.0, 0.0), log::Bool=false)
    nx, ny = dims(s)
    rimg = Array{Int16}(undef, ny, nx)
    rng = zeros(2)
    check(ccall((:weno5_arr2img, libweno5), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cint, Cint, Ptr{Int16}, Ptr{Float64}),
                s.h, component, range[1], range[2], log ? 1 : 0, length(cmap), rimg, rng))
    return cmap[rimg], (rng[1], rng[2])
end

"The driver loop of WENO5_2D()/WENO5_1D() kept on the device; returns (t, steps_done)."
function advance!(s::Solver, nsteps::Integer; tmax::Real=0.0, t::Real=0.0)
    tt = Ref(Float64(t)); nd = Ref{Cint}(0)
    check(ccall((:weno5_advance, libweno5), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{Float64}, Ref{Cint}), s.h, nsteps, tmax, tt, nd))
    return tt[], Int(nd[])
end

function compute_divB(s::Solver)
    d = Array{Float64}(undef, dims(s)...)
    check(ccall((:weno5_divb, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, d))
    return d
end

# ---- reference-shaped functional forms: host arrays in, fresh host arrays out
"(q4, bxb4, byb4) = classical_RK4_step(p, q, bxb, byb, dt)   -- Weno5_2D.jl:125-317"
function classical_RK4_step(p::Weno5Params, q::Array{Float64,3}, bxb::Array{Float64,2}, byb::Array{Float64,2}, dt::Float64)
    q4 = similar(q); bxb4 = similar(bxb); byb4 = similar(byb)
    check(ccall((:weno5_classical_rk4_step_2d, libweno5), Cint,
                (Ref{Weno5Params}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                p, q, bxb, byb, dt, q4, bxb4, byb4))
    return q4, bxb4, byb4
end
"q4 = classical_RK4_step(p, q, dt)   -- Weno5_1D.jl:152-236 (first return value)"
function classical_RK4_step(p::Weno5Params, q::Array{Float64,2}, dt::Float64)
    q4 = similar(q)
    check(ccall((:weno5_classical_rk4_step_1d, libweno5), Cint, (Ref{Weno5Params}, Ptr{Float64}, Cdouble, Ptr{Float64}), p, q, dt, q4))
    return q4
end

"""
Pipelined host -> host classical_RK4_step (weno5_pipe_*): the grid streams through the GPU in y-slabs
so that upload, compute and download overlap.  `pipe_step!(pipe, q, bxb, byb, dt, q4, bxb4, byb4)`
fills the preallocated outputs (no byte of which may lie in an input) and returns timestep(q4).
The copies overlap the compute only for page-locked arrays: take them from `pinned_array` (weno5_host_alloc);
ordinary `Array`s are pageable -- correct results, serialised copies (`pipe_overlapped(pipe)` tells which it was).
"""
mutable struct Pipe
    h::Ptr{Cvoid}
    function Pipe(p::Weno5Params; device::Integer=0, rows_per_slab::Integer=128)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:weno5_pipe_create, libweno5), Cint, (Ref{Weno5Params}, Cint, Cint, Ref{Ptr{Cvoid}}), p, device, rows_per_slab, h))
        s = new(h[])
        finalizer(x -> (x.h != C_NULL && ccall((:weno5_pipe_destroy, libweno5), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), s)
        return s
    end
end

function pipe_step!(p::Pipe, q::Array{Float64,3}, bxb::Array{Float64,2}, byb::Array{Float64,2}, dt::Real,
                    q4::Array{Float64,3}, bxb4::Array{Float64,2}, byb4::Array{Float64,2})
    dtn = Ref(0.0)
    check(ccall((:weno5_pipe_step, libweno5), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                p.h, q, bxb, byb, dt, q4, bxb4, byb4, dtn))
    return dtn[]
end

"1 if the last pipe_step! got page-locked arrays (its copies overlapped the compute)."
pipe_overlapped(p::Pipe) = ccall((:weno5_pipe_overlapped, libweno5), Cint, (Ptr{Cvoid},), p.h) == 1

"Page-locked Float64 array of the given shape (weno5_host_alloc); release it with `free_pinned` (weno5_host_free)."
function pinned_array(dims::Integer...)
    ptr = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:weno5_host_alloc, libweno5), Cint, (Ref{Ptr{Cvoid}}, Culonglong), ptr, 8 * prod(dims)))
    return unsafe_wrap(Array, Ptr{Float64}(ptr[]), dims; own=false)
end
free_pinned(a::Array{Float64}) = ccall((:weno5_host_free, libweno5), Cint, (Ptr{Cvoid},), pointer(a))

function interpolateB_center2face(p::Weno5Params, q::Array{Float64,3})
    s = Solver(p); upload!(s, q); _, bxb, byb = download(s); return bxb, byb
end

"state2conserved(rho,vx,vy,vz,Bx,By,Bz,P) -> 9-vector (Weno5_2D.jl:2043-2053); host arithmetic."
function state2conserved(rho, vx, vy, vz, Bx, By, Bz, P; gam=5/3)
    v2 = vx^2 + vy^2 + vz^2; B2 = Bx^2 + By^2 + Bz^2
    E = P/(gam-1) + 0.5*(rho*v2 + B2); S = P/rho^(gam-1)
    return [rho, rho*vx, rho*vy, rho*vz, Bx, By, Bz, S, E]
end

"Driver loop of Weno5_2D.jl:33-110: upload, advance, report t / dt / max divB, download."
function WENO5_2D(p::Weno5Params, q, bxb, byb; nstep::Integer=typemax(Int32), tmax::Real=0.1)
    s = Solver(p); upload!(s, q, bxb, byb)
    t, n = advance!(s, nstep; tmax=tmax)
    println("t = $t after $n steps; Max DivB = ", maximum(abs.(compute_divB(s))))
    return download(s)..., t
end

"Driver loop of Weno5_1D.jl:33-148 (without the plotting)."
function WENO5_1D(p::Weno5Params, q; tmax::Real=0.2)
    s = Solver(p); upload!(s, q)
    t, n = advance!(s, typemax(Int32); tmax=tmax)
    return download(s), t, n
end

# ---- 3-D extension (include/weno5mhd.h "3-D extension"; SURVEY 8 f4).  The reference has no 3-D solver: its
# classical_RK4_step docstring (Weno5_2D.jl:113-123) says the sweep functions are the same in every dimension, and
# these wrappers continue the 2-D names to q :: Array{Float64,4} (Nx,Ny,Nz,9), bxb/byb/bzb :: Array{Float64,3}.

# mirrors `weno5_params3d`
struct Weno5Params3D
    nx::Cint; ny::Cint; nz::Cint
    dx::Cdouble; dy::Cdouble; dz::Cdouble
    gamma::Cdouble; cfl::Cdouble; eps::Cdouble
    bc_x::Cint; bc_y::Cint; bc_z::Cint
    es_roundtrip::Cint
end

Weno5Params3D(nx, ny, nz; xsize=1.0, ysize=1.0, zsize=1.0, gamma=5/3, cfl=0.8, eps=1e-6,
              bc_x=BC_OUTFLOW, bc_y=BC_OUTFLOW, bc_z=BC_OUTFLOW, es_roundtrip=true) =
    Weno5Params3D(nx, ny, nz, xsize/nx, ysize/ny, zsize/nz, gamma, cfl, eps, bc_x, bc_y, bc_z, es_roundtrip ? 1 : 0)

mutable struct Solver3D
    h::Ptr{Cvoid}
    p::Weno5Params3D
    function Solver3D(p::Weno5Params3D; device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:weno5_create_3d, libweno5), Cint, (Ref{Weno5Params3D}, Cint, Ref{Ptr{Cvoid}}), p, device, h))
        s = new(h[], p)
        finalizer(x -> (x.h != C_NULL && ccall((:weno5_destroy_3d, libweno5), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), s)
        return s
    end
end

dims(s::Solver3D) = (Int(s.p.nx) + 2NBnd, Int(s.p.ny) + 2NBnd, Int(s.p.nz) + 2NBnd)

function upload!(s::Solver3D, q::Array{Float64,4}, bxb::Array{Float64,3}, byb::Array{Float64,3}, bzb::Array{Float64,3})
    @assert size(q) == (dims(s)..., 9) && size(bxb) == size(byb) == size(bzb) == dims(s)
    check(ccall((:weno5_upload_3d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.h, q, bxb, byb, bzb))
end
"q only: the three face fields are derived on the device (interpolateB_center2face, Weno5_2D.jl:1669)."
upload!(s::Solver3D, q::Array{Float64,4}) =
    check(ccall((:weno5_upload_3d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.h, q, C_NULL, C_NULL, C_NULL))

function download(s::Solver3D)
    d = dims(s)
    q = Array{Float64}(undef, d..., 9); bxb = Array{Float64}(undef, d...); byb = similar(bxb); bzb = similar(bxb)
    check(ccall((:weno5_download_3d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.h, q, bxb, byb, bzb))
    return q, bxb, byb, bzb
end

"timestep(q) -> (dt, (vxmax, vymax, vzmax)): Weno5_2D.jl:1930 with a third term."
function timestep(s::Solver3D)
    dt = Ref(0.0); vm = zeros(3)
    check(ccall((:weno5_timestep_3d, libweno5), Cint, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), s.h, dt, vm))
    return dt[], (vm[1], vm[2], vm[3])
end

"classical_RK4_step on the resident state (Weno5_2D.jl:125 with a z sweep and three face fields)."
classical_RK4_step!(s::Solver3D, dt::Real) =
    check(ccall((:weno5_rk4_step_3d, libweno5), Cint, (Ptr{Cvoid}, Cdouble), s.h, dt))

function advance!(s::Solver3D, nsteps::Integer; tmax::Real=0.0, t::Real=0.0)
    tt = Ref(Float64(t)); nd = Ref{Cint}(0)
    check(ccall((:weno5_advance_3d, libweno5), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{Float64}, Ref{Cint}), s.h, nsteps, tmax, tt, nd))
    return tt[], Int(nd[])
end

function compute_divB(s::Solver3D)
    d = Array{Float64}(undef, dims(s)...)
    check(ccall((:weno5_divb_3d, libweno5), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, d))
    return d
end

"(q4, bxb4, byb4, bzb4) = classical_RK4_step(q0, bxb0, byb0, bzb0, dt): host arrays in, host arrays out."
function classical_RK4_step(p::Weno5Params3D, q::Array{Float64,4}, bxb::Array{Float64,3}, byb::Array{Float64,3},
                            bzb::Array{Float64,3}, dt::Float64)
    q4 = similar(q); bx4 = similar(bxb); by4 = similar(byb); bz4 = similar(bzb)
    check(ccall((:weno5_classical_rk4_step_3d, libweno5), Cint,
                (Ref{Weno5Params3D}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), p, q, bxb, byb, bzb, dt, q4, bx4, by4, bz4))
    return q4, bx4, by4, bz4
end

"The driver loop of WENO5_2D (Weno5_2D.jl:33-110) for a volume: state resident on the GPU, dt recomputed every step."
function WENO5_3D(p::Weno5Params3D, q, bxb, byb, bzb; nstep::Integer=typemax(Int32) >> 12, tmax::Real=0.1)
    s = Solver3D(p)
    upload!(s, q, bxb, byb, bzb)
    t, n = advance!(s, nstep; tmax=tmax)
    return (download(s)..., t, n)
end

end # module
