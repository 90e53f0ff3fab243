// oracle_capi.cpp -- extern "C" surface of the CPU oracle, for ctypes (tests/, bench.py's
// cpu_baseline and --impl reference legs, __graft_entry__.smoke()).  TEST INFRASTRUCTURE ONLY.
// The scene-level entry points deliberately mirror include/pfc.h so that parity tests feed the
// oracle and the CUDA library the very same arrays.
#include "pfc_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <thread>

// Runs body(e) for e in [0, n) on n_threads std::threads (dynamic chunks of 8); the reference is
// single-threaded, threads are only used for the "all host cores" CPU baseline.
template <class F> static void parallel_for(int64_t n, int n_threads, F body) {
    if (n_threads <= 1 || n < 16) { for (int64_t e = 0; e < n; ++e) body(e); return; }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(8);
            if (b >= n) break;
            int64_t e_end = std::min<int64_t>(n, b + 8);
            for (int64_t e = b; e < e_end; ++e) body(e);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

using namespace orc;
typedef Dual<6> D6;

extern "C" {

// ---- unit-level entry points (pinned by the reference's own unit tests) ------------------------
int orc_sat(const double* a_c, const double* a_e, const double* a_R, const double* b_c, const double* b_e, const double* b_R,
            const double* R_a_b, const double* t_a_b) {
    OBB a, b;
    for (int k = 0; k < 3; ++k) { a.c[k] = a_c[k]; a.e[k] = a_e[k]; b.c[k] = b_c[k]; b.e[k] = b_e[k]; }
    for (int k = 0; k < 9; ++k) { a.R.m[k] = a_R[k]; b.R.m[k] = b_R[k]; }
    M3<double> R; V3<double> t;
    for (int k = 0; k < 9; ++k) R.m[k] = R_a_b[k];
    for (int k = 0; k < 3; ++k) t[k] = t_a_b[k];
    return BB_BB_intersect<double>(R, t, a, b) ? 1 : 0;
}

// zeta_in: n_in x 4 (row per vertex); zeta_out: 8 x 4.  Returns vertex count, or -1/-2 on the
// reference's error() paths.
int orc_clip_in_tet_coordinates(int n_in, const double* zeta_in, double* zeta_out) {
    Poly<4, double> p;
    p.n = n_in;
    for (int k = 0; k < n_in && k < 8; ++k) for (int i = 0; i < 4; ++i) p.v[k][i] = zeta_in[4 * k + i];
    ClipStatus st;
    Poly<4, double> r = clip_in_tet_coordinates(p, st);
    if (st.non_finite) return -1;
    if (st.bad_arity) return -2;
    for (int k = 0; k < r.n; ++k) for (int i = 0; i < 4; ++i) zeta_out[4 * k + i] = r.v[k][i];
    return r.n;
}

// plane: 4, tet: 4x4 col-major ([v;1] per column); out: 4 x 3.
int orc_clip_plane_tet(const double* plane, const double* tet, double* out) {
    V4<double> pl; M4<double> t;
    for (int k = 0; k < 4; ++k) pl[k] = plane[k];
    for (int k = 0; k < 16; ++k) t.m[k] = tet[k];
    Poly<3, double> r = clip_plane_tet(pl, t);
    for (int k = 0; k < r.n; ++k) for (int i = 0; i < 3; ++i) out[3 * k + i] = r.v[k][i];
    return r.n;
}

double orc_poly_centroid(int n, const double* v, const double* nhat, double* c_out) {
    Poly<3, double> p; p.n = n;
    for (int k = 0; k < n; ++k) for (int i = 0; i < 3; ++i) p.v[k][i] = v[3 * k + i];
    V3<double> nn = mk3<double>(nhat[0], nhat[1], nhat[2]);
    auto r = poly_centroid(p, nn);
    for (int i = 0; i < 3; ++i) c_out[i] = r.second[i];
    return r.first;
}

void orc_zero_small_coordinates(int n, double* zeta) {
    Poly<4, double> p; p.n = n;
    for (int k = 0; k < 8; ++k) for (int i = 0; i < 4; ++i) p.v[k][i] = (k < n) ? zeta[4 * k + i] : 0.0;
    Poly<4, double> r = zero_small_coordinates(p);
    int cnt = (n <= 4) ? ((n < 4) ? n : 4) : n;
    for (int k = 0; k < cnt; ++k) for (int i = 0; i < 4; ++i) zeta[4 * k + i] = r.v[k][i];
}

double orc_calc_clamped_piecewise(double x, double x_1, double x_2, double y_1, double y_2) { return calc_clamped_piecewise(x, x_1, x_2, y_1, y_2); }

void orc_traction_regularized(double v_c, double mu_s, double mu_d, const double* vel_t, double p_dA, double* out) {
    Regularized r = make_regularized(v_c, mu_s, mu_d);
    V3<double> o = traction(r, mk3<double>(vel_t[0], vel_t[1], vel_t[2]), p_dA);
    for (int i = 0; i < 3; ++i) out[i] = o[i];
}
void orc_traction_bristle(double tau, double k_bar, double mu_s, double mu_d, double magic, const double* Ts, double p_dA, double* out) {
    Bristle b = make_bristle(0, tau, k_bar, mu_s, mu_d, magic);
    V3<double> o = traction(b, mk3<double>(Ts[0], Ts[1], Ts[2]), p_dA);
    for (int i = 0; i < 3; ++i) out[i] = o[i];
}

// K: 6x6 row-major (symmetric; upper triangle read).  Outputs Sinv(6), Kbar_sqrt_inv(36 row-major).
void orc_decompose_K(const double* K, double magic, double* Sinv, double* Kbar_sqrt_inv) {
    SpatialStiffness<double> s;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) s.K[i][j] = K[6 * i + j];
    decompose_K(s, magic);
    for (int i = 0; i < 6; ++i) { Sinv[i] = s.Sinv[i]; for (int j = 0; j < 6; ++j) Kbar_sqrt_inv[6 * i + j] = s.Kbar_sqrt_inv[i][j]; }
}
// Dual-6 flavour: every scalar is 7 doubles (value, 6 partials).
void orc_decompose_K_dual6(const double* K7, double magic, double* Sinv7, double* Kbar_sqrt_inv7) {
    SpatialStiffness<D6> s;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) std::memcpy((void*)&s.K[i][j], K7 + 7 * (6 * i + j), sizeof(D6));
    decompose_K(s, magic);
    for (int i = 0; i < 6; ++i) {
        std::memcpy(Sinv7 + 7 * i, &s.Sinv[i], sizeof(D6));
        for (int j = 0; j < 6; ++j) std::memcpy(Kbar_sqrt_inv7 + 7 * (6 * i + j), &s.Kbar_sqrt_inv[i][j], sizeof(D6));
    }
}

void orc_inv44(const double* A, double* out) {
    M4<double> a; for (int k = 0; k < 16; ++k) a.m[k] = A[k];
    M4<double> r = inv44(a);
    for (int k = 0; k < 16; ++k) out[k] = r.m[k];
}
double orc_tet_volume(const double* v) {
    V3<double> p[4];
    for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) p[k][i] = v[3 * k + i];
    return tet_volume(p[0], p[1], p[2], p[3]);
}
static void put_obb(const OBB& o, double* c, double* e, double* R) {
    for (int k = 0; k < 3; ++k) { c[k] = o.c[k]; e[k] = o.e[k]; }
    for (int k = 0; k < 9; ++k) R[k] = o.R.m[k];
}
void orc_fit_tri_obb(const double* v, double* c, double* e, double* R) {
    V3<double> p[3];
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) p[k][i] = v[3 * k + i];
    put_obb(fit_tri_obb(p), c, e, R);
}
int orc_fit_tet_obb(const double* v, const double* eps, double* c, double* e, double* R) {
    V3<double> p[4];
    for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) p[k][i] = v[3 * k + i];
    OBB o;
    if (!fit_tet_obb(p, eps, o)) return -1;
    put_obb(o, c, e, R);
    return 0;
}
void orc_merge_obb(const double* ac, const double* ae, const double* aR, const double* bc, const double* be, const double* bR, double* c,
                   double* e, double* R) {
    OBB a, b;
    for (int k = 0; k < 3; ++k) { a.c[k] = ac[k]; a.e[k] = ae[k]; b.c[k] = bc[k]; b.e[k] = be[k]; }
    for (int k = 0; k < 9; ++k) { a.R.m[k] = aR[k]; b.R.m[k] = bR[k]; }
    put_obb(merge_obb(a, b), c, e, R);
}

// normal_wrench_cop + calc_patch_spatial_stiffness! on an explicit traction list (8 doubles per
// point: n(3) r(3) dA p).  Outputs cop(3), normal wrench(6: ang, lin), K (36 row-major, full).
void orc_patch_stiffness(int64_t n_points, const double* traction8, double k_bar, double* cop_out, double* wrench_out, double* K_out) {
    BodyBodyCache<double> b;
    for (int64_t k = 0; k < n_points; ++k) {
        const double* t = traction8 + 8 * k;
        b.traction.push_back(TractionCache<double>{mk3<double>(t[0], t[1], t[2]), mk3<double>(t[3], t[4], t[5]), t[6], t[7]});
    }
    V3<double> ang, lin;
    V3<double> cop = normal_wrench_cop(b, ang, lin);
    Bristle BF = make_bristle(0, 1.0, k_bar, 1.0, 1.0, 1.0);
    calc_patch_spatial_stiffness(b, BF, cop);
    for (int i = 0; i < 3; ++i) { cop_out[i] = cop[i]; wrench_out[i] = ang[i]; wrench_out[3 + i] = lin[i]; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) K_out[6 * i + j] = b.stiff.K[i][j];
}

// ---- scene-level entry points (mirror include/pfc.h) ---------------------------------------------
void* orc_scene_create() { return new Scene(); }
void orc_scene_destroy(void* h) { delete static_cast<Scene*>(h); }

int orc_add_mesh(void* h, int kind, int64_t n_point, const double* xyz, int64_t n_prim, const int32_t* idx, const double* eps, double Ebar,
                 int64_t n_node, const double* node_c, const double* node_e, const double* node_R, const int32_t* node_left,
                 const int32_t* node_right, const int32_t* node_leaf_id) {
    Scene& sc = *static_cast<Scene*>(h);
    Mesh m;
    m.kind = kind;
    m.point.resize(n_point);
    for (int64_t k = 0; k < n_point; ++k) for (int i = 0; i < 3; ++i) m.point[k][i] = xyz[3 * k + i];
    int w = (kind == 0) ? 3 : 4;
    m.idx.assign(idx, idx + w * n_prim);
    if (kind == 1) m.eps.assign(eps, eps + n_point);
    m.Ebar = Ebar;
    m.tree.box.resize(n_node);
    for (int64_t k = 0; k < n_node; ++k) {
        for (int i = 0; i < 3; ++i) { m.tree.box[k].c[i] = node_c[3 * k + i]; m.tree.box[k].e[i] = node_e[3 * k + i]; }
        for (int i = 0; i < 9; ++i) m.tree.box[k].R.m[i] = node_R[9 * k + i];
    }
    m.tree.left.assign(node_left, node_left + n_node);
    m.tree.right.assign(node_right, node_right + n_node);
    m.tree.leaf_id.assign(node_leaf_id, node_leaf_id + n_node);
    m.tree.root = 0;
    sc.mesh.push_back(std::move(m));
    return int(sc.mesh.size()) - 1;
}

// params: regularized {mu_s, mu_d, v_c}; bristle {tau, k_bar, mu_s, mu_d, magic}
int orc_add_instruction(void* h, int mesh_1, int mesh_2, double chi, int model, const double* params, int n_quad_rule) {
    Scene& sc = *static_cast<Scene*>(h);
    if (mesh_1 < 0 || mesh_2 < 0 || mesh_1 >= int(sc.mesh.size()) || mesh_2 >= int(sc.mesh.size())) return -1;
    if (sc.mesh[mesh_2].kind != 1) return -1;  // id_2 is always a Tet mesh (mechanism_scenario.jl:399-416)
    if (n_quad_rule < 1 || n_quad_rule > 2) return -1;
    Instruction ci{};
    ci.id_1 = mesh_1; ci.id_2 = mesh_2; ci.chi = chi; ci.model = model;
    ci.quad = getTriQuadRule(n_quad_rule);
    if (model == 0) ci.reg = make_regularized(params[2], params[0], params[1]);
    else ci.bri = make_bristle(sc.n_bristle++, params[0], params[1], params[2], params[3], params[4]);
    sc.ins.push_back(ci);
    return int(sc.ins.size()) - 1;
}

int orc_n_bristle(void* h) { return static_cast<Scene*>(h)->n_bristle; }
int64_t orc_n_visited(void* h) { return static_cast<Scene*>(h)->n_visited; }

// Layouts as in pfc.h: X_r2_r1 [env][ins][16] col-major; twist [env][ins][6] (ang, lin);
// s/sdot [env][bristle][6]; wrench [env][ins][6] (ang, lin); n_pairs/flags [env][ins].
// keep != 0 retains every (env, ins) pair list and traction list for orc_get_pairs / orc_get_traction.
int orc_eval_f64(void* h, int64_t n_env, const double* X, const double* twist, const double* s, double* wrench, double* sdot,
                 int64_t* n_pairs, int32_t* flags, int n_threads, int keep) {
    Scene& sc = *static_cast<Scene*>(h);
    const int n_ins = int(sc.ins.size());
    const int nb = sc.n_bristle;
    if (keep) { sc.last_pairs.assign(n_env * n_ins, {}); sc.last_traction.assign(n_env * n_ins, {}); }
    std::atomic<int64_t> visited_total{0};
    parallel_for(n_env, n_threads, [&](int64_t e) {
        std::vector<std::pair<int32_t, int32_t>> pairs;
        std::vector<double> trac;
        int64_t visited = 0;
        for (int k = 0; k < n_ins; ++k) {
            const int64_t ei = e * n_ins + k;
            int f = force_single_elastic_intersection<double>(sc, sc.ins[k], X + 16 * ei, X + 16 * ei, twist + 6 * ei, s ? s + 6 * nb * e : nullptr,
                                                              wrench + 6 * ei, sdot ? sdot + 6 * nb * e : nullptr, pairs, keep ? &trac : nullptr,
                                                              visited);
            if (n_pairs) n_pairs[ei] = int64_t(pairs.size());
            if (flags) flags[ei] = f;
            if (keep) { sc.last_pairs[ei] = pairs; sc.last_traction[ei] = trac; }
        }
        visited_total += visited;
    });
    sc.n_visited = visited_total.load();
    return 0;
}

// Dual-6: X_bp is the Float64 transform used for the broad phase (R8/H6); X7, twist7, s7 are
// value+6 partials per scalar.
int orc_eval_dual6(void* h, int64_t n_env, const double* X_bp, const double* X7, const double* twist7, const double* s7, double* wrench7,
                   double* sdot7, int64_t* n_pairs, int32_t* flags, int n_threads) {
    Scene& sc = *static_cast<Scene*>(h);
    const int n_ins = int(sc.ins.size());
    const int nb = sc.n_bristle;
    parallel_for(n_env, n_threads, [&](int64_t e) {
        std::vector<std::pair<int32_t, int32_t>> pairs;
        int64_t visited = 0;
        for (int k = 0; k < n_ins; ++k) {
            const int64_t ei = e * n_ins + k;
            int f = force_single_elastic_intersection<D6>(sc, sc.ins[k], X_bp + 16 * ei, reinterpret_cast<const D6*>(X7) + 16 * ei,
                                                          reinterpret_cast<const D6*>(twist7) + 6 * ei,
                                                          s7 ? reinterpret_cast<const D6*>(s7) + 6 * nb * e : nullptr,
                                                          reinterpret_cast<D6*>(wrench7) + 6 * ei,
                                                          sdot7 ? reinterpret_cast<D6*>(sdot7) + 6 * nb * e : nullptr, pairs, nullptr, visited);
            if (n_pairs) n_pairs[ei] = int64_t(pairs.size());
            if (flags) flags[ei] = f;
        }
    });
    return 0;
}

// ALGORITHMIC work of the reference algorithm for one batch, counted with the instrumented scalar:
// out[0] broad-phase FLOPs, out[1] narrow-phase + friction FLOPs, out[2] node pairs visited,
// out[3] candidate pairs, out[4] traction points (quadrature points with p > 0).
int orc_count_work(void* h, int64_t n_env, const double* X, const double* twist, const double* s, int64_t* out) {
    Scene& sc = *static_cast<Scene*>(h);
    const int n_ins = int(sc.ins.size());
    const int nb = sc.n_bristle;
    for (int k = 0; k < 5; ++k) out[k] = 0;
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (int64_t e = 0; e < n_env; ++e) {
        for (int k = 0; k < n_ins; ++k) {
            const int64_t ei = e * n_ins + k;
            Counted Xc[16], tw[6], w[6];
            std::vector<Counted> sv(6 * std::max(nb, 1)), sd(6 * std::max(nb, 1));
            for (int i = 0; i < 16; ++i) Xc[i] = Counted(X[16 * ei + i]);
            for (int i = 0; i < 6; ++i) tw[i] = Counted(twist[6 * ei + i]);
            for (int i = 0; i < 6 * nb; ++i) sv[i] = Counted(s[6 * nb * e + i]);
            flop_counter() = 0;
            int64_t broad = 0, visited = 0, points = 0;
            force_single_elastic_intersection<Counted>(sc, sc.ins[k], X + 16 * ei, Xc, tw, sv.data(), w, sd.data(), pairs, nullptr, visited, &broad, &points);
            out[0] += broad;
            out[1] += flop_counter() - broad;
            out[2] += visited;
            out[3] += int64_t(pairs.size());
            out[4] += points;
        }
    }
    return 0;
}

// Partial regularized wrench of one (env = 0) instruction over the slice of its candidate pairs that rank
// `rank` of `world` owns under the product's sharding rule (contiguous runs of 256-pair chunks).  Used by the
// world_size-2 gloo test of the multi-GPU host logic: the sum over ranks must equal the full wrench.
int orc_eval_slice_regularized(void* h, const double* X, const double* twist, int ins, int rank, int world, double* wrench_out, int64_t* n_pairs_out) {
    Scene& sc = *static_cast<Scene*>(h);
    const Instruction& ci = sc.ins.at(ins);
    if (ci.model != 0) return -1;
    const Mesh& m1 = sc.mesh[ci.id_1];
    const Mesh& m2 = sc.mesh[ci.id_2];
    M4<double> xf; for (int k = 0; k < 16; ++k) xf.m[k] = X[16 * ins + k];
    M4<double> x12 = inv_transform(xf);
    std::vector<std::pair<int32_t, int32_t>> pairs;
    int64_t visited = 0;
    tree_tree_intersect<double>(pairs, visited, rot_of(x12), mk3<double>(x12(0, 3), x12(1, 3), x12(2, 3)), m1.tree, m1.tree.root, m2.tree, m2.tree.root);
    const int64_t n = int64_t(pairs.size());
    const int64_t n_units = (n + 255) / 256;
    const int64_t lo = std::min<int64_t>(n, (n_units * rank / world) * 256), hi = std::min<int64_t>(n, (n_units * (rank + 1) / world) * 256);
    BodyBodyCache<double> b;
    b.mesh_1 = &m1; b.mesh_2 = &m2;
    b.x_r2_r1 = xf; b.x_r1_r2 = x12;
    for (int k = 0; k < 6; ++k) b.twist_r2_r1_r2[k] = twist[6 * ins + k];
    b.chi = ci.chi; b.Ebar = m2.Ebar; b.quad = ci.quad;
    for (int64_t k = lo; k < hi; ++k) {
        if (m1.kind == 0) integrate_over_tri_tet(pairs[k].first, pairs[k].second, b);
        else integrate_over_tet_tet(pairs[k].first, pairs[k].second, b);
    }
    V6<double> w = yes_contact_regularized(ci.reg, b);
    for (int k = 0; k < 6; ++k) wrench_out[k] = w[k];
    if (n_pairs_out) *n_pairs_out = n;
    return 0;
}

int64_t orc_get_pairs(void* h, int64_t env, int ins, int32_t* pairs, int64_t cap) {
    Scene& sc = *static_cast<Scene*>(h);
    const int n_ins = int(sc.ins.size());
    const auto& v = sc.last_pairs.at(env * n_ins + ins);
    for (int64_t k = 0; k < int64_t(v.size()) && k < cap; ++k) { pairs[2 * k] = v[k].first; pairs[2 * k + 1] = v[k].second; }
    return int64_t(v.size());
}
int64_t orc_get_traction(void* h, int64_t env, int ins, double* out, int64_t cap_points) {
    Scene& sc = *static_cast<Scene*>(h);
    const int n_ins = int(sc.ins.size());
    const auto& v = sc.last_traction.at(env * n_ins + ins);
    int64_t n = int64_t(v.size()) / 8;
    for (int64_t k = 0; k < n && k < cap_points; ++k) for (int i = 0; i < 8; ++i) out[8 * k + i] = v[8 * k + i];
    return n;
}

int orc_max_threads() { unsigned n = std::thread::hardware_concurrency(); return n ? int(n) : 1; }


// Small kernels of the path behind one entry point, for the restated unit tests of the reference (tests/test_oracle_units.py):
//   0 weightPoly(p1, p2, w1, w2)            in: p1[3] p2[3] w1 w2                 out: r[3]       (src/math_kernel/utility.jl:21-26)
//   1 vec_sub_vec_proj(v, n)                in: v[3] n[3]                         out: r[3]       (vector_projections.jl:2-7)
//   2 a_dot_one_pad_b(a, b)                 in: a[4] b[3]                         out: r[1]       (vector_projections.jl:9-13)
//   3 triangle kernels                      in: v1[3] v2[3] v3[3]                 out: area, centroid[3], normal[3]   (geometry_kernel.jl:3-10)
//   4 getTriQuadRule(n)                     in: n                                 out: n_point, w[3], zeta[3][3]      (src/clip/quadrature.jl:21-41)
//   5 basic_dh algebra                      in: R[9] (row-major) t[3] p[3]        out: dh*p [3], inv(dh)*(dh*p) [3], (dh*dh)*p [3]   (basic_dh.jl)
int orc_kat(int which, const double* in, double* out) {
    if (which == 0) {
        V3<double> r = weightPoly(mk3<double>(in[0], in[1], in[2]), mk3<double>(in[3], in[4], in[5]), in[6], in[7]);
        for (int i = 0; i < 3; ++i) out[i] = r[i];
    } else if (which == 1) {
        V3<double> r = vec_sub_vec_proj(mk3<double>(in[0], in[1], in[2]), mk3<double>(in[3], in[4], in[5]));
        for (int i = 0; i < 3; ++i) out[i] = r[i];
    } else if (which == 2) {
        V4<double> a; for (int i = 0; i < 4; ++i) a[i] = in[i];
        out[0] = a_dot_one_pad_b(a, mk3<double>(in[4], in[5], in[6]));
    } else if (which == 3) {
        const V3<double> v1 = mk3<double>(in[0], in[1], in[2]), v2 = mk3<double>(in[3], in[4], in[5]), v3 = mk3<double>(in[6], in[7], in[8]);
        const V3<double> n = triangleNormal(v1, v2, v3), c = centroid3(v1, v2, v3);
        out[0] = triangle_area(v1, v2, v3, n);
        for (int i = 0; i < 3; ++i) { out[1 + i] = c[i]; out[4 + i] = n[i]; }
    } else if (which == 4) {
        const TriQuadRule q = getTriQuadRule(int(in[0]));
        out[0] = q.n;
        for (int k = 0; k < 3; ++k) { out[1 + k] = q.w[k]; for (int i = 0; i < 3; ++i) out[4 + 3 * k + i] = q.zeta[k][i]; }
    } else if (which == 5) {
        M3<double> R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = in[3 * i + j];
        const V3<double> t = mk3<double>(in[9], in[10], in[11]);
        const M4<double> dh = basic_dh<double>(R, t);
        V4<double> p; for (int i = 0; i < 3; ++i) p[i] = in[12 + i]; p[3] = 1.0;
        const V4<double> q1 = mul4v<double, double, double>(dh, p);
        const V4<double> q2 = mul4v<double, double, double>(inv_transform(dh), q1);
        const V4<double> q3 = mul4v<double, double, double>(mul44<double, double, double>(dh, dh), p);
        for (int i = 0; i < 3; ++i) { out[i] = q1[i]; out[3 + i] = q2[i]; out[6 + i] = q3[i]; }
    } else return -1;
    return 0;
}

}  // extern "C"
