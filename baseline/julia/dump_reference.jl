# dump_reference.jl -- run on a machine that HAS Julia (>= 1.1) and the pinned Manifest of
# ryanelandt/PressureFieldContact.jl; it cannot run in the build container (no Julia there).
#
#   julia --project=/path/to/PressureFieldContact.jl baseline/julia/dump_reference.jl OUT_DIR [boxes pencil vol_vol]
#
# Writes OUT_DIR/<scene>.json ("pfc-reference-dump-1").  Copy the files to tests/golden/julia/ of this
# repository; tests/test_reference_dump.py then rebuilds each scene through the C ABI FROM THE DUMP
# ALONE (meshes, flattened trees, instructions) and checks, for every sampled state and instruction,
#   candidate-pair lists (bit-exact, same order)          m.TT_Cache                (src/obb/tree_types.jl:32-50)
#   traction points                                       tm.bodyBodyCache.TractionCache (src/mechanism_scenario.jl:51-58)
#   wrench in r2                                          yes_contact!              (src/contact_algorithms_friction.jl:50-72,119-143)
#   bristle state derivative                              tm.ṡ
#   generalized contact force                             tm.f_generalized          (src/contact_algorithms_non_friction.jl:60-68)
# of the CPU oracle AND of the CUDA library against what the reference computed here -- the step
# that pins the oracle to the real reference (SURVEY.md H8 / section 8c).
#
# What is dumped per scene
#   meshes[]        name, kind, body id, Ebar, points, 1-based connectivity, eps, and the bin_BB_Tree
#                   flattened PRE-ORDER (node 0 = root): box c / e / R (column-major 9), child links,
#                   leaf id (0-based primitive, -1 for internal nodes = the reference's id == -9999)
#   instructions[]  mesh ids (0-based, after the reference's Tri/Tet ordering rule), chi, friction
#                   model + parameters, quadrature rule, bristle id
#   samples[]       states x = [q; v; s] along a short Radau run (and the initial state); for each one
#                   the boundary arrays of every instruction (x_r2_r1.mat, twist_r2_r1_r2), the pair
#                   list, the TractionCache, the wrench, plus s-dot, f_generalized and calcXd!'s x-dot
#   trajectory      t[k], x[k] of N Radau steps from the script's initial state (state parity over N steps)
using Printf
using StaticArrays
using LinearAlgebra
using RigidBodyDynamics
using Rotations: RotZ
using PressureFieldContact
using PressureFieldContact.Geometry
using PressureFieldContact.Binary_BB_Trees
const PFC = PressureFieldContact

set_zero_subnormals(true)
LinearAlgebra.BLAS.set_num_threads(1)

# ---- a minimal JSON writer (no package needed); floats with 17 significant digits round-trip ---------------
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isfinite(x) ? @sprintf("%.17g", x) : "null"
jval(x::Number) = jnum(x)
jval(x::Bool) = x ? "true" : "false"
jval(x::Nothing) = "null"
jval(x::AbstractString) = "\"" * escape_string(x) * "\""
jval(x::Union{AbstractVector,Tuple}) = "[" * join((jval(v) for v in x), ",") * "]"
jval(x::AbstractMatrix) = jval(vec(x))                      # column-major
jval(d::AbstractDict) = "{" * join((jval(string(k)) * ":" * jval(v) for (k, v) in d), ",") * "}"
jval(p::Vector{<:Pair}) = "{" * join((jval(string(k)) * ":" * jval(v) for (k, v) in p), ",") * "}"   # ordered object

# ---- bin_BB_Tree -> pre-order arrays (src/obb/tree_types.jl:1-16) ---------------------------------------------
function flatten_tree(tree)
    c = Vector{Vector{Float64}}(); e = Vector{Vector{Float64}}(); R = Vector{Vector{Float64}}()
    left = Int[]; right = Int[]; leaf = Int[]
    function visit(t)
        k = length(left)                    # 0-based index of this node
        push!(c, collect(t.box.c)); push!(e, collect(t.box.e)); push!(R, collect(vec(SMatrix{3,3,Float64,9}(t.box.R))))
        push!(left, -1); push!(right, -1)
        if t.id != -9999                    # leaf: id = 1-based primitive index
            push!(leaf, t.id - 1)
        else
            push!(leaf, -1)
            left[k + 1] = visit(t.node_1)
            right[k + 1] = visit(t.node_2)
        end
        return k
    end
    visit(tree)
    return ["c" => c, "e" => e, "R" => R, "left" => left, "right" => right, "leaf_id" => leaf]
end

function dump_mesh(mc)
    em = mc.mesh
    is_tet = em.tet !== nothing
    return ["name" => mc.name, "kind" => is_tet ? "tet" : "tri", "body_id" => Int(mc.BodyID),
            "Ebar" => mc.c_prop === nothing ? nothing : mc.c_prop.Ē,
            "point" => [collect(p) for p in em.point],
            "prim" => is_tet ? [collect(Int.(t)) for t in em.tet] : [collect(Int.(t)) for t in em.tri],   # 1-based
            "eps" => is_tet ? em.ϵ : nothing,
            "tree" => flatten_tree(mc.tree)]
end

function dump_instruction(ci)
    fm = ci.FrictionModel
    n_quad_rule = length(ci.quad.w) == 1 ? 1 : 2
    if fm isa PFC.Bristle
        return ["id_1" => Int(ci.id_1) - 1, "id_2" => Int(ci.id_2) - 1, "chi" => ci.χ, "model" => 1,
                "params" => [fm.τ, fm.k̄, fm.μs, fm.μd, fm.magic], "n_quad_rule" => n_quad_rule, "bristle_id" => Int(fm.BristleID) - 1]
    else
        return ["id_1" => Int(ci.id_1) - 1, "id_2" => Int(ci.id_2) - 1, "chi" => ci.χ, "model" => 0,
                "params" => [fm.μs, fm.μd, fm.v_c], "n_quad_rule" => n_quad_rule, "bristle_id" => -1]
    end
end

# One evaluation of forceAllElasticIntersections! (src/contact_algorithms_non_friction.jl:60-84), instruction by
# instruction with the package's own functions, recording every intermediate.
function dump_sample(m, x::Vector{Float64})
    tm = m.float
    copyto!(tm, x)                                       # src/extensions.jl:21-31
    PFC.refreshJacobians!(m, tm)
    tm.f_generalized .= 0.0
    per_ins = []
    for c_ins in m.ContactInstructions
        PFC.calcTriTetIntersections!(m, c_ins)
        n_pair = length(m.TT_Cache)
        pairs = [[m.TT_Cache.vc[i][1] - 1, m.TT_Cache.vc[i][2] - 1] for i = 1:n_pair]
        PFC.refreshBodyBodyCache!(m, tm, c_ins)           # boundary arrays in mode Float64 (also when there is no pair)
        b = tm.bodyBodyCache
        rec = Pair{String,Any}["X_r2_r1" => collect(vec(b.x_r²_r¹.mat)), "x_rw_r2" => collect(vec(b.x_rʷ_r².mat)),
               "twist_r2" => collect(PFC.as_static_vector(b.twist_r²_r¹_r²)), "pairs" => pairs]
        contact = false
        wrench = zeros(6)
        trac = []
        if n_pair != 0
            PFC.integrate_over!(b, m.TT_Cache)
            trac = [vcat(collect(b.TractionCache[i].n̂), collect(b.TractionCache[i].r_cart), b.TractionCache[i].dA, b.TractionCache[i].p)
                    for i = 1:length(b.TractionCache)]
            if !isempty(b.TractionCache)
                w = PFC.yes_contact!(c_ins.FrictionModel, tm, c_ins)
                PFC.addGeneralizedForcesThirdLaw!(w, tm, c_ins)
                wrench = collect(PFC.as_static_vector(w))
                contact = true
            end
        end
        contact || PFC.no_contact!(c_ins.FrictionModel, tm, c_ins)
        push!(rec, "traction" => trac); push!(rec, "contact" => contact); push!(rec, "wrench_r2" => wrench)
        push!(per_ins, rec)
    end
    f_gen = copy(tm.f_generalized[1:tm.nv])
    sdot = copy(tm.ṡ.parent)
    xx = zeros(length(x))
    PFC.calcXd!(xx, x, m)                                 # the whole right-hand side (src/contact_algorithms_non_friction.jl:18-38)
    return ["x" => x, "per_instruction" => per_ins, "f_generalized" => f_gen, "sdot" => sdot, "xdot" => xx]
end

function dump_scene(name::String, m, out_dir::String; n_steps::Int, h_max::Float64, n_sample::Int)
    x0 = get_state(m)
    rr = Radau_for_MechanismScenario(m)
    rr.step.h_max = h_max
    data_time, data_state = integrate_scenario_radau(rr, t_final=1.0e9, max_steps=n_steps)
    n_row = length(data_time)
    rows = unique(vcat(1, round.(Int, range(1, stop=n_row, length=min(n_sample, n_row)))))
    samples = [dump_sample(m, data_state[r, :]) for r in rows]
    g = m.float.state.mechanism.gravitational_acceleration.v
    doc = ["format" => "pfc-reference-dump-1", "scene" => name, "julia" => string(VERSION),
           "nq" => num_positions(m.float.state.mechanism), "nv" => num_velocities(m.float.state.mechanism), "n_bristle" => length(m.bristle_ids),
           "gravity" => collect(g),
           "meshes" => [dump_mesh(m.MeshCache[k]) for k in m.mesh_ids],
           "instructions" => [dump_instruction(ci) for ci in m.ContactInstructions],
           "sample_rows" => rows .- 1,
           "samples" => samples,
           "trajectory" => ["h_max" => h_max, "t" => data_time, "x" => [data_state[r, :] for r = 1:n_row]]]
    open(joinpath(out_dir, name * ".json"), "w") do io
        write(io, jval(doc))
    end
    println("wrote ", joinpath(out_dir, name * ".json"), ": ", n_row, " rows, ", length(samples), " samples")
end

include(joinpath(@__DIR__, "scenes_reference.jl"))   # scene_boxes(), scene_vol_vol(), scene_pencil()

function main(args)
    out_dir = length(args) >= 1 ? args[1] : "reference_dump"
    which = length(args) >= 2 ? args[2:end] : ["boxes", "vol_vol", "pencil"]
    mkpath(out_dir)
    ("boxes" in which) && dump_scene("boxes", scene_boxes(), out_dir, n_steps=1000, h_max=0.05, n_sample=40)
    ("vol_vol" in which) && dump_scene("vol_vol", scene_vol_vol(), out_dir, n_steps=300, h_max=0.05, n_sample=20)
    ("pencil" in which) && dump_scene("pencil", scene_pencil(), out_dir, n_steps=1000, h_max=0.01, n_sample=40)
end

main(ARGS)
