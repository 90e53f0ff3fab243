# bench_reference.jl -- the reference's own CPU timing of the hot path, for a machine that HAS Julia
# (>= 1.1) and the pinned Manifest of ryanelandt/PressureFieldContact.jl (the build container has no Julia,
# so bench.py reports a C++ port instead and labels it so).
#
#   julia --project=/path/to/PressureFieldContact.jl baseline/julia/bench_reference.jl [n_env] [dump_dir]
#
# Times forceAllElasticIntersections!(m, m.float) ONLY (src/contact_algorithms_non_friction.jl:60-68; no mass matrix,
# no Cholesky), single thread like the reference's own tests (test/runtests.jl:13-14), after JIT warm-up:
#   C1  the test/boxes.jl scene at sampled settled-stack states      -> evals/s  (the unit of bench.py's `value`)
#   C3  n_env such states one after the other (the reference has no batching: the time is n_env x C1)
# The states come from a dump written by dump_reference.jl (dump_dir/boxes.json, key "samples") when given,
# otherwise from a short Radau run here.  Prints one JSON line shaped like bench.py's cpu_baseline object:
#   {"cpu_baseline": {"value": ..., "unit": "evals/s", "cores": 1, "kind": "reference", "sample": "..."}}
using Printf
using Statistics
using StaticArrays
using LinearAlgebra
using RigidBodyDynamics
using PressureFieldContact
const PFC = PressureFieldContact

set_zero_subnormals(true)
LinearAlgebra.BLAS.set_num_threads(1)

include(joinpath(@__DIR__, "scenes_reference.jl"))   # scene_boxes(), shared with dump_reference.jl

function sample_states(m, n::Int)
    rr = Radau_for_MechanismScenario(m)
    rr.step.h_max = 0.05
    data_time, data_state = integrate_scenario_radau(rr, t_final=5.0, max_steps=1000)
    n_row = length(data_time)
    rows = round.(Int, range(max(1, n_row ÷ 4), stop=n_row, length=n))   # after the drop: the stack is in contact
    return [data_state[r, :] for r in rows]
end

function time_evals(m, states, n_rep::Int)
    tm = m.float
    t = Float64[]
    for rep = 1:n_rep, x in states
        copyto!(tm, x)                                   # state vector -> MechanismState (not timed as contact work, but part of every call)
        t0 = time_ns()
        PFC.forceAllElasticIntersections!(m, tm)
        push!(t, (time_ns() - t0) * 1.0e-9)
    end
    return t
end

function main(args)
    n_env = length(args) >= 1 ? parse(Int, args[1]) : 4096
    m = scene_boxes()
    states = sample_states(m, 64)
    time_evals(m, states, 2)                             # JIT warm-up
    t = time_evals(m, states, max(1, cld(n_env, length(states))))
    med = median(t)
    @printf("{\"cpu_baseline\": {\"value\": %.6g, \"unit\": \"evals/s\", \"cores\": 1, \"kind\": \"reference\", \"sample\": \"%d evaluations of forceAllElasticIntersections! on test/boxes.jl states, median %.3g us, Julia %s\"}, \"c3_seconds_for_%d_envs\": %.6g}\n",
            1.0 / med, length(t), med * 1.0e6, string(VERSION), n_env, sum(t) * n_env / length(t))
end

main(ARGS)
