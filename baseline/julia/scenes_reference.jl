# scenes_reference.jl -- the reference's example scenes, built with the reference's own API exactly as its scripts do
# (test/boxes.jl:18-45, test/test_vol_vol.jl:2-31, test/pencil.jl:176-236 with is_bristle = true and no controller).
# Included by dump_reference.jl and bench_reference.jl; needs `using PressureFieldContact, RigidBodyDynamics, StaticArrays`.
using Rotations: RotZ
using PressureFieldContact.Geometry

# ---- the scenes, as the reference's own scripts build them -------------------------------------------------------
function scene_boxes()                                    # test/boxes.jl:18-45
    box_rad = 0.05
    c_prop = ContactProperties(Ē=1.0e6)
    i_c = InertiaProperties(400.0)
    i_r = InertiaProperties(400.0, d=box_rad)
    eM_r = as_tri_eMesh(eMesh_box(box_rad))
    eM_c = as_tet_eMesh(eMesh_box(box_rad))
    m = MechanismScenario()
    nt_plane = add_contact!(m, "plane", as_tet_eMesh(eMesh_half_plane()), c_prop=c_prop)
    b1 = add_body_contact!(m, "box_1", eM_r, i_prop=i_r)
    b2 = add_body_contact!(m, "box_2", eM_c, i_prop=i_c, c_prop=c_prop)
    b3 = add_body_contact!(m, "box_3", eM_r, i_prop=i_r)
    b4 = add_body_contact!(m, "box_4", eM_c, i_prop=i_c, c_prop=c_prop)
    add_friction_regularize!(m, nt_plane.id, b1.id, μd=0.0, χ=2.2, n_quad_rule=2)
    add_friction_regularize!(m, b1.id, b2.id, μd=0.2, χ=0.2, n_quad_rule=2)
    add_friction_regularize!(m, b2.id, b3.id, μd=0.2, χ=0.2, n_quad_rule=2)
    add_friction_regularize!(m, b3.id, b4.id, μd=0.2, χ=0.2, n_quad_rule=2)
    finalize!(m)
    set_state_spq!(m, b1.joint, trans=SVector(0.0, 0.0,  2 * box_rad), w=SVector(0.0, 0.0, 1.0))
    set_state_spq!(m, b2.joint, trans=SVector(0.0, 0.0,  5 * box_rad), w=SVector(0.0, 0.0, 2.0))
    set_state_spq!(m, b3.joint, trans=SVector(0.0, 0.0,  8 * box_rad), w=SVector(0.0, 0.0, 3.0))
    set_state_spq!(m, b4.joint, trans=SVector(0.0, 0.0, 11 * box_rad), w=SVector(0.0, 0.0, 4.0))
    return m
end

function scene_vol_vol()                                  # test/test_vol_vol.jl:2-31 (tet-tet, frictionless)
    box_rad = 0.05
    c_prop = ContactProperties(Ē=1.0e6)
    m = MechanismScenario()
    nt_plane = add_contact!(m, "plane", as_tet_eMesh(eMesh_half_plane()), c_prop=c_prop)
    b1 = add_body_contact!(m, "box_1", as_tet_eMesh(eMesh_box(box_rad)), i_prop=InertiaProperties(400.0), c_prop=c_prop)
    add_friction_regularize!(m, nt_plane.id, b1.id, μd=0.0, χ=0.0, n_quad_rule=2)
    finalize!(m)
    set_state_spq!(m, b1.joint, trans=SVector(0.0, 0.0, 2 * box_rad), w=SVector(0.0, 0.0, 1.14))
    return m
end

# test/pencil.jl needs a MeshCat Visualizer and its controller; here only its contact scene with is_bristle = true, no
# controller (the arm falls under gravity onto the pencil): every instruction kind of configuration C2 is exercised.
function scene_pencil()                                   # test/pencil.jl:176-236
    k̄ = 8.0e4; τ = 0.01; magic = 1.0e-2; v_tol = 1.0e-5
    pad_rad = 0.0035; penci_length = 0.16; penci_rad = 0.0035
    m = MechanismScenario()
    mech = m.float.state.mechanism
    c_prop = ContactProperties(Ē=1.0e6)
    i_pad = InertiaProperties(16000.0)
    i_rigid = InertiaProperties(400.0, d=penci_rad)
    eM_plane = eMesh_half_plane()
    transform!(eM_plane, 0.6)
    eM_pad_n = as_tet_eMesh(eMesh_sphere(pad_rad, 4))
    transform!(eM_pad_n, SVector(0.0, -(pad_rad + penci_rad), 0.0))
    transform!(eM_pad_n, SMatrix{3,3,Float64,9}(2.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 2.0))
    eM_pad_p = as_tet_eMesh(eMesh_sphere(pad_rad, 4))
    transform!(eM_pad_p, SVector(0.0, +(pad_rad + penci_rad), 0.0))
    transform!(eM_pad_p, SMatrix{3,3,Float64,9}(2.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 2.0))
    eM_rev_y = as_tri_eMesh(eMesh_box(pad_rad * SVector(1, 7, 1), pad_rad * SVector(0, -4, 8)))
    eM_penci = as_tri_eMesh(create_swept_mesh(f_swept_triv, [0.0, 0.013, penci_length], [0.0, penci_rad, penci_rad], 12, true, rot_half=true))
    transform!(eM_penci, SVector(0.0, 0.0, penci_rad))
    nt_plane = add_contact!(m, "plane", as_tet_eMesh(eM_plane), c_prop=c_prop)
    mii_base = MeshInertiaInfo(1 .* ones(SMatrix{3,3,Float64,9}), zeros(SVector{3,Float64}), 1.0, NaN)
    nt_tra_z = add_body_from_inertia!(mech, "tra_z", mii_base, joint=Prismatic(SVector(0.0, 0.0, 1.0)))
    nt_rev_y = add_body_from_inertia!(mech, "rev_y", mii_base, joint=Revolute(SVector(0.0, 1.0, 0.0)), body=nt_tra_z.body)
    add_contact!(m, "rev_y", eM_rev_y, body=nt_rev_y.body)
    nt_pad_n = add_body_contact!(m, "pad_n", eM_pad_n, c_prop=c_prop, i_prop=i_pad, joint=Prismatic(SVector(0.0, +1.0, 0.0)), body=nt_rev_y.body)
    nt_pad_p = add_body_contact!(m, "pad_p", eM_pad_p, c_prop=c_prop, i_prop=i_pad, joint=Prismatic(SVector(0.0, -1.0, 0.0)), body=nt_rev_y.body)
    nt_penci = add_body_contact!(m, "name", eM_penci, i_prop=i_rigid)
    add_friction_bristle!(m, nt_penci.id, nt_pad_n.id, μd=0.5, χ=0.6, k̄=k̄, magic=magic, τ=τ)
    add_friction_bristle!(m, nt_penci.id, nt_pad_p.id, μd=0.5, χ=0.6, k̄=k̄, magic=magic, τ=τ)
    add_friction_regularize!(m, nt_penci.id, nt_plane.id, μd=0.5, χ=0.6, v_tol=v_tol)
    add_friction_regularize!(m, nt_pad_n.id, nt_pad_p.id, μd=0.0, χ=0.6, v_tol=v_tol)
    finalize!(m)
    set_configuration!(m, nt_tra_z.joint, [0.004])        # arm low enough for the pads to straddle the pencil
    set_configuration!(m, nt_rev_y.joint, [0.0])
    set_state_spq!(m, nt_penci.joint, trans=SVector(penci_length * -0.5, 0.0, 0.0), rot=RotZ(-pi / 2))
    return m
end

