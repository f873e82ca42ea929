# QPControlB200.jl -- thin Julia host shim over libqpcontrol_b200.so (C ABI: include/qpcontrol_b200.h).
#
# SOURCE ONLY: Julia is not installed in the build container, so this file has never been executed there.  It is the
# binding a QPControl.jl maintainer would load next to the reference: the same exported names and argument meaning as
# reference src/QPControl.jl:5-33 (MomentumBasedController, addtask!, addcontact!, regularize!, setdesired!, disable!,
# the task types, ContactPoint, StandingController), with the numerics behind `ccall`.  Differences forced by batching:
# the controller functors take matrices (one column per robot instance) and `t` is dropped from the batched call
# because neither reference functor reads it (src/lowlevel/momentum.jl:41-81, src/highlevel/standing.jl:58-89).
module QPControlB200

using LinearAlgebra
using StaticArrays
using RigidBodyDynamics
import RigidBodyDynamics: Mechanism, RigidBody, Joint, bodies, joints, tree_joints, joint_to_predecessor,
    predecessor, successor, root_body, spatial_inertia, num_positions, num_velocities

export MomentumBasedController, StandingController, ContactPoint, OSQPSettings, QPSolveFailure,
    SpatialAccelerationTask, AngularAccelerationTask, LinearAccelerationTask, PointAccelerationTask,
    JointAccelerationTask, MomentumRateTask, LinearMomentumRateTask,
    addtask!, addcontact!, regularize!, setdesired!, disable!, checkstatus, warmstart!, resetwarmstart!, simulate!,
    simulateplant!, bindse3pd!, InterpPiece, rotationpiece, vectorpiece

const LIB = Ref{String}(get(ENV, "QPCONTROL_B200_LIB", "libqpcontrol_b200"))

# ---- qpc_settings (include/qpcontrol_b200.h): the OSQPSettings.* attributes the reference sets -------------------
struct OSQPSettings
    rho::Cdouble; sigma::Cdouble; alpha::Cdouble
    eps_abs::Cdouble; eps_rel::Cdouble; eps_prim_inf::Cdouble; eps_dual_inf::Cdouble
    adaptive_rho_tolerance::Cdouble
    max_iter::Int32; scaling::Int32; adaptive_rho::Int32; adaptive_rho_interval::Int32; check_termination::Int32
    reserved::NTuple{3,Int32}
end
function OSQPSettings(; kwargs...)
    s = Ref{OSQPSettings}()
    ccall((:qpc_default_settings, LIB[]), Cvoid, (Ref{OSQPSettings},), s)
    d = Dict(n => getfield(s[], n) for n in fieldnames(OSQPSettings))
    for (k, v) in kwargs
        d[k] = v
    end
    OSQPSettings((d[n] for n in fieldnames(OSQPSettings))...)
end

# reference src/exceptions.jl:1-11
struct QPSolveFailure <: Exception
    instance::Int
    status::Int32
end
Base.showerror(io::IO, e::QPSolveFailure) = print(io, "QP solve failed for instance $(e.instance): status $(e.status)")
# reference src/lowlevel/momentum.jl:83-91: OPTIMAL (1) and ALMOST_OPTIMAL (2) are accepted
function checkstatus(status::AbstractVector{Int32})
    i = findfirst(s -> !(s == 1 || s == 2), status)
    i === nothing || throw(QPSolveFailure(i, status[i]))
    nothing
end

lasterror() = unsafe_string(ccall((:qpc_last_error, LIB[]), Cstring, ()))
check(code, what) = code < 0 ? error("$what failed ($code): $(lasterror())") : code

# ---- RigidBodyDynamics.Mechanism -> flat tree (qpc_mechanism_create) ----------------------------------------------
# Bodies are numbered in tree_joints order (a joint is identified with its successor body); -1 is the world.
struct FlatMechanism
    handle::Ptr{Cvoid}
    mechanism::Mechanism{Float64}
    bodyindex::Dict{RigidBody{Float64},Int32}
end

function jointkind(j::Joint)
    jt = joint_type(j)
    jt isa Revolute && return Int32(0), Vector(jt.axis)
    jt isa Prismatic && return Int32(1), Vector(jt.axis)
    jt isa QuaternionFloating && return Int32(2), zeros(3)
    jt isa Fixed && return Int32(3), zeros(3)
    error("unsupported joint type $(typeof(jt)) (remove_fixed_tree_joints! first; Planar/SPQuatFloating are not built)")
end

function FlatMechanism(mechanism::Mechanism{Float64})
    js = tree_joints(mechanism)
    nb = length(js)
    index = Dict{RigidBody{Float64},Int32}(root_body(mechanism) => Int32(-1))
    parent = Vector{Int32}(undef, nb); jtype = similar(parent)
    axis = zeros(3, nb); XR = zeros(9, nb); Xp = zeros(3, nb); mass = zeros(nb); com = zeros(3, nb); J = zeros(9, nb)
    for (i, j) in enumerate(js)
        body = successor(j, mechanism)
        index[body] = Int32(i - 1)
        parent[i] = index[predecessor(j, mechanism)]
        jtype[i], axis[:, i] = jointkind(j)
        # fixed transform: frame before the joint -> predecessor's default frame
        T = joint_to_predecessor(j)
        XR[:, i] = vec(permutedims(Matrix(rotation(T))))   # row-major
        Xp[:, i] = translation(T)
        # spatial inertia about the body's default frame origin (frame after the joint)
        I = spatial_inertia(body)
        mass[i] = I.mass
        com[:, i] = I.mass > 0 ? Vector(I.cross_part ./ I.mass) : zeros(3)
        J[:, i] = vec(permutedims(Matrix(I.moment)))
    end
    g = Vector(mechanism.gravitational_acceleration.v)
    h = ccall((:qpc_mechanism_create, LIB[]), Ptr{Cvoid},
              (Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
               Ptr{Cdouble}, Ptr{Cdouble}), nb, parent, jtype, axis, XR, Xp, mass, com, J, g)
    h == C_NULL && error("qpc_mechanism_create: $(lasterror())")
    FlatMechanism(h, mechanism, index)
end

# ---- tasks (reference src/tasks.jl): host-side records; `desired` is the setdesired! default -------------------------
abstract type AbstractMotionTask end
const KIND = Dict(:spatial => 0, :angular => 1, :linear => 2, :point => 3, :joint => 4, :momentum => 5, :linmom => 6)
mutable struct PathTask <: AbstractMotionTask           # Spatial / Angular / Linear / Point acceleration tasks
    kind::Int32; base::RigidBody{Float64}; body::RigidBody{Float64}; frame::RigidBody{Float64}
    point::Vector{Float64}; desired::Vector{Float64}; index::Int32
end
SpatialAccelerationTask(m::Mechanism, path; frame=path.target) =                      # tasks.jl:8-18
    PathTask(0, path.source, path.target, frame, zeros(3), zeros(6), -1)
AngularAccelerationTask(m::Mechanism, path; frame=path.target) =                      # tasks.jl:52-62
    PathTask(1, path.source, path.target, frame, zeros(3), zeros(3), -1)
LinearAccelerationTask(m::Mechanism, path; frame=path.target) =                       # tasks.jl:91-101
    PathTask(2, path.source, path.target, frame, zeros(3), zeros(3), -1)
PointAccelerationTask(m::Mechanism, path, point::Point3D) =                           # tasks.jl:131-142
    PathTask(3, path.source, path.target, path.source, Vector(point.v), zeros(3), -1)
mutable struct JointAccelerationTask <: AbstractMotionTask                            # tasks.jl:173-183
    joint::Joint{Float64}; desired::Vector{Float64}; index::Int32
end
JointAccelerationTask(j::Joint) = JointAccelerationTask(j, zeros(num_velocities(j)), -1)
mutable struct MomentumTask <: AbstractMotionTask                                     # tasks.jl:192-262
    kind::Int32; desired::Vector{Float64}; index::Int32
end
MomentumRateTask(m::Mechanism, centroidalframe) = MomentumTask(5, zeros(6), -1)
LinearMomentumRateTask(m::Mechanism, centroidalframe) = MomentumTask(6, zeros(3), -1)

dimension(t::AbstractMotionTask) = length(t.desired)
# reference setdesired!: frame checks are the caller's responsibility here (the reference @framechecks, tasks.jl:24-26)
setdesired!(t::AbstractMotionTask, desired) = (t.desired .= vec(collect(desired)); t)

# ---- contacts (reference src/contacts.jl:27-73) --------------------------------------------------------------------
mutable struct ContactPoint{N}
    body::RigidBody{Float64}; position::Vector{Float64}; normal::Vector{Float64}; mu::Float64
    weight::Base.RefValue{Float64}; maxnormalforce::Base.RefValue{Float64}; index::Int32
end
disable!(p::ContactPoint) = (p.maxnormalforce[] = 0.0; p)                             # contacts.jl:72
isenabled(p::ContactPoint) = p.maxnormalforce[] > 0                                   # contacts.jl:73

# ---- MomentumBasedController{N} (reference src/lowlevel/momentum.jl:1-34) ---------------------------------------------
mutable struct MomentumBasedController{N}
    flat::FlatMechanism
    handle::Ptr{Cvoid}
    tasks::Vector{AbstractMotionTask}
    contacts::Vector{ContactPoint{N}}
    device::Int32
    initialized::Bool
end

function MomentumBasedController{N}(mechanism::Mechanism{Float64}, optimizer::OSQPSettings;
                                    floatingjoint=nothing, device::Integer=0) where {N}
    flat = FlatMechanism(mechanism)
    fb = floatingjoint === nothing ? Int32(-1) : flat.bodyindex[successor(floatingjoint, mechanism)]
    h = ccall((:qpc_controller_create, LIB[]), Ptr{Cvoid}, (Ptr{Cvoid}, Int32, Int32, Ref{OSQPSettings}),
              flat.handle, N, fb, optimizer)
    h == C_NULL && error("qpc_controller_create: $(lasterror())")
    c = MomentumBasedController{N}(flat, h, AbstractMotionTask[], ContactPoint{N}[], device, false)
    finalizer(c) do x
        ccall((:qpc_controller_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), x.handle)
        ccall((:qpc_mechanism_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), x.flat.handle)
    end
    c
end

bodyid(c::MomentumBasedController, b) = c.flat.bodyindex[b]

# addtask!(controller, task[, weight]) -- momentum.jl:99-117: no weight = hard constraint, number = w e'e, matrix = e'We
function addtask!(c::MomentumBasedController, task::AbstractMotionTask, weight=nothing)
    mode, w, W = weight === nothing ? (0, 0.0, C_NULL) : weight isa Number ? (1, Float64(weight), C_NULL) :
                 (2, 0.0, Matrix{Float64}(permutedims(weight)))
    args = if task isa PathTask
        (task.kind, bodyid(c, task.base), bodyid(c, task.body), bodyid(c, task.frame), task.point, Int32(-1))
    elseif task isa JointAccelerationTask
        (Int32(4), Int32(-1), Int32(-1), Int32(-1), zeros(3), bodyid(c, successor(task.joint, c.flat.mechanism)))
    else
        (task.kind, Int32(-1), Int32(-1), Int32(-1), zeros(3), Int32(-1))
    end
    task.index = check(ccall((:qpc_add_task, LIB[]), Cint,
                             (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Cdouble}, Int32, Int32, Cdouble, Ptr{Cdouble}),
                             c.handle, args[1], args[2], args[3], args[4], args[5], args[6], mode, w, W), "qpc_add_task")
    push!(c.tasks, task)
    task
end

# regularize!(controller, joint, weight) -- momentum.jl:128-131
regularize!(c::MomentumBasedController, joint::Joint, weight) =
    check(ccall((:qpc_regularize, LIB[]), Cint, (Ptr{Cvoid}, Int32, Cdouble), c.handle,
                bodyid(c, successor(joint, c.flat.mechanism)), weight), "qpc_regularize")

# addcontact!(controller, body, position, normal, mu) -- momentum.jl:142-148, contacts.jl:38-69
function addcontact!(c::MomentumBasedController{N}, body::RigidBody, position::Point3D, normal::FreeVector3D, mu) where {N}
    p = ContactPoint{N}(body, Vector(position.v), Vector(normal.v), mu, Ref(0.0), Ref(0.0), -1)
    p.index = check(ccall((:qpc_add_contact, LIB[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                          c.handle, bodyid(c, body), p.position, p.normal, mu), "qpc_add_contact")
    push!(c.contacts, p)
    p
end

# The reference re-evaluates every Parameter on each `solve!` (tasks.jl:40 `desired`, contacts.jl:54-57 `weight[]`,
# `maxnormalforce[]`, `disable!`), so `setdesired!` / `disable!` / weight changes made BETWEEN ticks must reach the device:
# syncdefaults! pushes the current values before every tick (a few dozen setter calls; the library uploads the program
# once, at the next tick, only if something changed).
function syncdefaults!(c::MomentumBasedController)
    for p in c.contacts
        check(ccall((:qpc_set_contact_params, LIB[]), Cint, (Ptr{Cvoid}, Int32, Cdouble, Cdouble), c.handle, p.index,
                    p.weight[], p.maxnormalforce[]), "qpc_set_contact_params")
    end
    for t in c.tasks
        check(ccall((:qpc_set_task_desired, LIB[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}), c.handle, t.index,
                    t.desired), "qpc_set_task_desired")
    end
end

# lazy initialize! (momentum.jl:43-46,150-156): pushes the mutable defaults and freezes the program on the device
function initialize!(c::MomentumBasedController)
    syncdefaults!(c)
    check(ccall((:qpc_finalize, LIB[]), Cint, (Ptr{Cvoid}, Int32), c.handle, c.device), "qpc_finalize")
    c.initialized = true
end

struct qpc_batch_in
    q::Ptr{Cdouble}; v::Ptr{Cdouble}; desired::Ptr{Cdouble}; desired_stride::Int64
    contact_weight::Ptr{Cdouble}; contact_maxnormalforce::Ptr{Cdouble}; contact_stride::Int64
    task_weight::Ptr{Cdouble}; task_weight_stride::Int64            # per-tick Parameter weights (momentum.jl:107-110)
    contact_geometry::Ptr{Cdouble}; contact_geometry_stride::Int64  # per-tick position / normal / mu (contacts.jl:39)
    task_weight_matrix::Ptr{Cdouble}; task_weight_matrix_stride::Int64  # per-tick matrix weights (momentum.jl:113-117)
    time::Ptr{Cdouble}; time_stride::Int64                          # the functor's t (se3pdcontroller.jl:13); 0 = one value
end
qpc_batch_in(q, v, desired, dstride, cw, cm, cstride) =
    qpc_batch_in(q, v, desired, dstride, cw, cm, cstride, C_NULL, 0, C_NULL, 0, C_NULL, 0, C_NULL, 0)
qpc_batch_in(q, v, desired, dstride, cw, cm, cstride, tw, twstride, cg, cgstride) =
    qpc_batch_in(q, v, desired, dstride, cw, cm, cstride, tw, twstride, cg, cgstride, C_NULL, 0, C_NULL, 0)
qpc_batch_in(q, v, desired, dstride, cw, cm, cstride, tw, twstride, cg, cgstride, time::Ptr{Cdouble}) =
    qpc_batch_in(q, v, desired, dstride, cw, cm, cstride, tw, twstride, cg, cgstride, C_NULL, 0, time, 0)
struct qpc_batch_out
    tau::Ptr{Cdouble}; vdot::Ptr{Cdouble}; wrench::Ptr{Cdouble}; status::Ptr{Int32}; iters::Ptr{Int32}
    residuals::Ptr{Cdouble}; factorizations::Ptr{Int32}
end

# The control tick for B instances: (controller)(tau, t, x) of momentum.jl:41-81 with one COLUMN per instance
# (Julia is column-major, the C ABI wants [B][n] row-major: a (n x B) Matrix is exactly that memory).
# tau: nv x B (overwritten), q: nq x B, v: nv x B.  Returns (vdot, wrenches[6, ncontacts, B], status).
# taskweight (ntasks x B) and contactgeometry (7 x ncontacts x B: position, normal, mu) carry Parameter-valued task
# weights and contact frames per instance.
function (c::MomentumBasedController)(tau::Matrix{Float64}, t::Number, q::Matrix{Float64}, v::Matrix{Float64};
                                      maxnormalforce::Union{Nothing,Matrix{Float64}}=nothing,
                                      weight::Union{Nothing,Matrix{Float64}}=nothing,
                                      taskweight::Union{Nothing,Matrix{Float64}}=nothing,
                                      contactgeometry::Union{Nothing,Array{Float64,3}}=nothing, check::Bool=true)
    c.initialized || initialize!(c)
    syncdefaults!(c)   # setdesired! / disable! / weight changes since the last tick (the reference re-reads them in solve!)
    B = size(q, 2)
    nc = length(c.contacts)
    vdot = similar(v); wrench = zeros(6, nc, B); status = zeros(Int32, B); iters = zeros(Int32, B); res = zeros(2, B)
    tref = Float64[t]   # the functor's t: read by the device-side SE3PDControllers (bindse3pd!)
    GC.@preserve tau q v vdot wrench status iters res maxnormalforce weight taskweight contactgeometry tref begin
        bin = qpc_batch_in(pointer(q), pointer(v), C_NULL, 0,
                           weight === nothing ? C_NULL : pointer(weight),
                           maxnormalforce === nothing ? C_NULL : pointer(maxnormalforce), nc,
                           taskweight === nothing ? C_NULL : pointer(taskweight),
                           taskweight === nothing ? 0 : size(taskweight, 1),
                           contactgeometry === nothing ? C_NULL : pointer(contactgeometry), 7 * nc, pointer(tref))
        bout = qpc_batch_out(pointer(tau), pointer(vdot), pointer(wrench), pointer(status), pointer(iters),
                             pointer(res), C_NULL)
        QPControlB200.check(ccall((:qpc_solve_batch, LIB[]), Cint,
                                  (Ptr{Cvoid}, Int64, Ref{qpc_batch_in}, Ref{qpc_batch_out}, Int32, Ptr{Cvoid}),
                                  c.handle, B, bin, bout, 0, C_NULL), "qpc_solve_batch")
    end
    check && checkstatus(status)
    vdot, wrench, status
end

# ---- SE3PDController on the device (SURVEY.md 8(f) rank 3; reference src/lowlevel/se3pdcontroller.jl:1-18) ---------------
# One Interpolated piece of an SE3Trajectory component (interpolated.jl:1-60), mirror of qpc_interp_piece.
struct InterpPiece
    break_start::Cdouble; x0::Cdouble; xf::Cdouble
    y0::NTuple{4,Cdouble}; dy::NTuple{3,Cdouble}; angle::Cdouble
    coeffs::NTuple{6,Cdouble}; ncoeffs::Int32; reserved::Int32
end
_coeffs(c) = (ntuple(i -> i <= length(c) ? Float64(c[i]) : 0.0, 6), Int32(length(c)))
# rotation piece from quaternions (w, x, y, z): axis / angle of y0 \ yf (interpolated.jl:75-78)
function rotationpiece(x0, xf, y0::NTuple{4,Float64}, yf::NTuple{4,Float64}; coeffs=Float64[], break_start=0.0)
    w1, x1, y1, z1 = y0; w2, x2, y2, z2 = yf
    d = (w1 * w2 + x1 * x2 + y1 * y2 + z1 * z2, w1 * x2 - x1 * w2 - y1 * z2 + z1 * y2,
         w1 * y2 + x1 * z2 - y1 * w2 - z1 * x2, w1 * z2 - x1 * y2 + y1 * x2 - z1 * w2)      # conj(y0) * yf
    d = d[1] < 0 ? map(-, d) : d
    s = sqrt(d[2]^2 + d[3]^2 + d[4]^2)
    axis = s > 0 ? (d[2] / s, d[3] / s, d[4] / s) : (1.0, 0.0, 0.0)
    c, n = _coeffs(coeffs)
    InterpPiece(break_start, x0, xf, y0, axis, 2atan(s, d[1]), c, n, 0)
end
function vectorpiece(x0, xf, y0, yf; coeffs=Float64[], break_start=0.0)
    c, n = _coeffs(coeffs)
    InterpPiece(break_start, x0, xf, (y0[1], y0[2], y0[3], 0.0), (yf[1] - y0[1], yf[2] - y0[2], yf[3] - y0[3]), 0.0, c, n, 0)
end
# SE3PDController(base, body, trajectory, weight, gains) bound to the SpatialAccelerationTask `task` it drives: from then on
# the device evaluates setdesired!(task, controller(t, state)) every tick.  gains: (K_ang, D_ang, K_lin, D_lin) 3x3 matrices.
function bindse3pd!(c::MomentumBasedController, task::SpatialAccelerationTask, base::RigidBody, body::RigidBody,
                    gains::NTuple{4,Matrix{Float64}}, angular::Vector{InterpPiece}, linear::Vector{InterpPiece};
                    angular_break_end::Float64=NaN, linear_break_end::Float64=NaN)
    K = vcat((vec(permutedims(g)) for g in gains)...)   # row-major 3x3 blocks
    id = ccall((:qpc_add_se3pd, LIB[]), Cint,
               (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Cdouble}, Int32, Ptr{InterpPiece}, Int32, Cdouble, Int32, Ptr{InterpPiece},
                Int32, Cdouble), c.handle, task.index, bodyid(c, base), bodyid(c, body), K,
               length(angular), angular, isnan(angular_break_end) ? 0 : 1, isnan(angular_break_end) ? 0.0 : angular_break_end,
               length(linear), linear, isnan(linear_break_end) ? 0 : 1, isnan(linear_break_end) ? 0.0 : linear_break_end)
    check(id, "qpc_add_se3pd")
    id
end

# ---- sequential ticks (SURVEY.md 8(f) rank 1) -----------------------------------------------------------------------------
# The reference solves every tick in one OSQP workspace, so each `solve!` (momentum.jl:58) starts from the previous
# tick's iterates and adapted rho.  warmstart!(controller, true) gives every batch column that behaviour.
function warmstart!(c::MomentumBasedController, on::Bool=true)
    c.initialized || initialize!(c)
    check(ccall((:qpc_set_warm_start, LIB[]), Cint, (Ptr{Cvoid}, Int32), c.handle, on ? 1 : 0), "qpc_set_warm_start")
end
resetwarmstart!(c::MomentumBasedController) =
    check(ccall((:qpc_reset_warm_start, LIB[]), Cint, (Ptr{Cvoid},), c.handle), "qpc_reset_warm_start")

# `nsteps` closed-loop control ticks of period dt for B robots on the device: the batched counterpart of
# simulate(state, T, PeriodicController(tau, dt, controller)) (notebooks/Standing controller.ipynb:202-214) under the
# contact model the controller itself assumes.  q (nq x B) and v (nv x B) are advanced in place.
function simulate!(c::MomentumBasedController, q::Matrix{Float64}, v::Matrix{Float64}, dt::Float64, nsteps::Integer;
                   check::Bool=true)
    c.initialized || initialize!(c)
    syncdefaults!(c)
    B = size(q, 2)
    nc = length(c.contacts)
    tau = zeros(size(v)); vdot = similar(v); wrench = zeros(6, nc, B); status = zeros(Int32, B)
    GC.@preserve q v tau vdot wrench status begin
        bin = qpc_batch_in(C_NULL, C_NULL, C_NULL, 0, C_NULL, C_NULL, 0)
        bout = qpc_batch_out(pointer(tau), pointer(vdot), pointer(wrench), pointer(status), C_NULL, C_NULL, C_NULL)
        QPControlB200.check(ccall((:qpc_step_batch, LIB[]), Cint,
                                  (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ref{qpc_batch_in}, Ref{qpc_batch_out},
                                   Cdouble, Int32, Int32, Ptr{Cvoid}),
                                  c.handle, B, q, v, bin, bout, dt, nsteps, 0, C_NULL), "qpc_step_batch")
    end
    check && checkstatus(status)
    tau, vdot, wrench, status
end

# The same loop with a PLANT between the control ticks (qpc_simulate_batch): forward dynamics vd = M^-1 (tau - c + J'f) under
# a soft ground contact at z = groundz, `substeps` integration steps per control period, tau held in between
# (PeriodicController).  stiffness / damping per contact point, mu of the ground.
struct qpc_contact_model
    stiffness::Cdouble; damping::Cdouble; mu::Cdouble; v_eps::Cdouble; ground_z::Cdouble
end
function simulateplant!(c::MomentumBasedController, q::Matrix{Float64}, v::Matrix{Float64}, dt::Float64, nticks::Integer;
                        groundz::Float64=0.0, substeps::Integer=8, stiffness=5e4, damping=1e3, mu=0.8, v_eps=1e-2,
                        check::Bool=true)
    c.initialized || initialize!(c)
    syncdefaults!(c)
    B = size(q, 2)
    nc = length(c.contacts)
    tau = zeros(size(v)); vdot = similar(v); wrench = zeros(6, nc, B); status = zeros(Int32, B)
    GC.@preserve q v tau vdot wrench status begin
        bin = qpc_batch_in(C_NULL, C_NULL, C_NULL, 0, C_NULL, C_NULL, 0)
        bout = qpc_batch_out(pointer(tau), pointer(vdot), pointer(wrench), pointer(status), C_NULL, C_NULL, C_NULL)
        plant = qpc_contact_model(stiffness, damping, mu, v_eps, groundz)
        QPControlB200.check(ccall((:qpc_simulate_batch, LIB[]), Cint,
                                  (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ref{qpc_batch_in}, Ref{qpc_batch_out},
                                   Ref{qpc_contact_model}, Cdouble, Int32, Int32, Int32, Ptr{Cvoid}),
                                  c.handle, B, q, v, bin, bout, plant, dt, substeps, nticks, 0, C_NULL), "qpc_simulate_batch")
    end
    check && checkstatus(status)
    tau, vdot, wrench, status
end

# ---- StandingController (reference src/highlevel/standing.jl:18-56); the PD laws of :58-85 run on the device ---------
struct StandingController{N}
    lowlevel::MomentumBasedController{N}
end
function StandingController(lowlevel::MomentumBasedController{N}, feet::AbstractVector{<:RigidBody}, pelvis::RigidBody,
                            nominalstate::MechanismState; joint_regularization=0.05, linear_momentum_weight=1.0,
                            comgains=(10.0, 2sqrt(10.0)), pelvisgains=(20.0, 2sqrt(20.0)), jointgains=(100.0, 20.0),
                            comref=center_of_mass(nominalstate).v - SVector(0, 0, 0.05)) where {N}
    m = lowlevel.flat.mechanism
    world = root_body(m)
    for j in tree_joints(m)
        regularize!(lowlevel, j, joint_regularization)                                     # standing.jl:35
    end
    for foot in feet
        addtask!(lowlevel, SpatialAccelerationTask(m, path(m, world, foot)))              # standing.jl:37-38
    end
    linmom = addtask!(lowlevel, LinearMomentumRateTask(m, nothing), linear_momentum_weight) # standing.jl:40-41
    pelvistask = addtask!(lowlevel, AngularAccelerationTask(m, path(m, world, pelvis)))   # standing.jl:43-44
    onpaths = Set(j for foot in feet for j in collect(path(m, world, foot)))
    jts = JointAccelerationTask[]
    for j in tree_joints(m)                                                               # standing.jl:46-49
        if joint_type(j) isa Revolute && !(j in onpaths)
            push!(jts, addtask!(lowlevel, JointAccelerationTask(j)))
        end
    end
    nj = length(jts)
    jtask = Int32[t.index for t in jts]
    jbody = Int32[bodyid(lowlevel, successor(t.joint, m)) for t in jts]
    kp = fill(Float64(jointgains[1]), nj); kd = fill(Float64(jointgains[2]), nj)
    qref = Float64[configuration(nominalstate, t.joint)[1] for t in jts]
    check(ccall((:qpc_standing_setup, LIB[]), Cint,
                (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                 Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}),
                lowlevel.handle, linmom.index, pelvistask.index, bodyid(lowlevel, pelvis), nj, jtask, jbody, kp, kd, qref,
                comgains[1], comgains[2], pelvisgains[1], pelvisgains[2], Vector{Float64}(comref)), "qpc_standing_setup")
    StandingController{N}(lowlevel)
end
(c::StandingController)(tau::Matrix{Float64}, t::Number, q::Matrix{Float64}, v::Matrix{Float64}; kwargs...) =
    c.lowlevel(tau, t, q, v; kwargs...)                                                   # standing.jl:87

end # module
