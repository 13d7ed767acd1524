// TEST INFRASTRUCTURE (oracle/_ref): driver that runs the UNMODIFIED reference
// (headers + the six .cpp of libtfel, compiled where they lie under /root/reference)
// against the stand-in third-party headers in oracle/shim/, and dumps what the
// reference itself produces for the hot path: mesh, boundary submesh, DOF maps,
// Dirichlet sets, the CSR / RHS it hands to PETSc (recorded by the fake
// petscksp.h), and phase timings. It is the checker, never the product:
// only tests/, gen_golden.py, __graft_entry__.build() and bench.py's reference
// arm may build or execute it.
//
// Problem set-ups follow the reference's own programs:
//   poisson   README.md:66-120 (P1/P2/P1-bubble, qf5pT, homogeneous Dirichlet, f = 1)
//   poissong  same mesh, g(x) Dirichlet data, Gaussian source of examples/poisson.cpp:4-14
//             and a reaction term c(x) u v (exercises free_function in a bilinear form)
//   tet       test/tetrahedron_cube.cpp:16-46 (P1 with qfSym4pTet, P1-bubble with qfSym35pTet)
//   stokes    test/stokes_2d_p2_p1.cpp:56-106 (with "atol": the test as shipped passes
//             "abstol" and throws at solver construction)
//   ns        one Newton system of test/navier_stokes_2d_p2_p1.cpp:292-366 with
//             synthetic previous iterates
//   laplace2  test/composite_fe.cpp:82-131 (two decoupled Laplacians)
//   quadrature  dumps every quadrature table of src/core/quadrature.cpp
//   export    the output side (fes.hpp:464-504, export.hpp): nodal interpolants written by the reference's own
//             exporters into the directory --out (the files are the fixtures of tests/golden/export)
#include <iostream>
#include <fstream>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>

#include "core/mesh.hpp"
#include "core/fe.hpp"
#include "core/fes.hpp"
#include "core/quadrature.hpp"
#include "core/form.hpp"
#include "core/composite_fe.hpp"
#include "core/composite_fes.hpp"
#include "core/composite_form.hpp"
#include "core/projector.hpp"
#include "core/export.hpp"

extern "C" {
  // CPU Krylov restatement (oracle/tfel_oracle.c)
  int orc_gmres_ilu(int n, const int* rowptr, const int* colind, const double* val,
                    const double* b, double* x, int restart, int ilu_levels,
                    double rtol, double atol, double dtol, int maxits,
                    double* final_rnorm);
  int orc_cg_jacobi(int n, const int* rowptr, const int* colind, const double* val,
                    const double* b, double* x, double rtol, double atol, int maxits,
                    double* final_rnorm);
}

namespace {
  typedef std::chrono::steady_clock clk;
  double seconds_since(clk::time_point t0) {
    return std::chrono::duration<double>(clk::now() - t0).count();
  }

  // ---- dump container: [u32 name_len][name][u8 dtype][u32 ndim][u64 dims...][raw data]
  // dtype: 0 = uint32, 1 = int32, 2 = float64
  struct dumper {
    std::ofstream out;
    bool hash_only;
    std::ostringstream meta;
    bool first;
    dumper(const std::string& path, bool hash_only) : hash_only(hash_only), first(true) {
      if (not hash_only and not path.empty()) out.open(path.c_str(), std::ios::binary);
    }
    // position-weighted multiplicative hash, reproducible with numpy uint64 arithmetic:
    //   H = sum_i (a_i + 1) * (2 i + 1) * 0x9E3779B97F4A7C15  (mod 2^64)
    template<typename T>
    static std::uint64_t hash_int(const T* p, std::size_t n) {
      std::uint64_t h(0);
      for (std::size_t i(0); i < n; ++i)
        h += (static_cast<std::uint64_t>(static_cast<std::uint32_t>(p[i])) + 1ull) * (2ull * i + 1ull) * 0x9E3779B97F4A7C15ull;
      return h;
    }
    void note(const std::string& key, const std::string& json_value) {
      meta << (first ? "" : ", ") << "\"" << key << "\": " << json_value;
      first = false;
    }
    template<typename T> void note_num(const std::string& key, T v) {
      std::ostringstream s; s.precision(17); s << v; note(key, s.str());
    }
    template<typename T>
    void put(const std::string& name, std::uint8_t dtype, const T* data, const std::vector<std::uint64_t>& dims) {
      std::uint64_t n(1); for (auto d : dims) n *= d;
      if (dtype != 2) {
        std::ostringstream s; s << "\"" << hash_int(data, n) << "\"";
        note("hash_" + name, s.str());
      } else {
        double s1(0.0), s2(0.0), sa(0.0);
        for (std::uint64_t i(0); i < n; ++i) { const double v(static_cast<double>(data[i])); s1 += v; s2 += v * v; sa += std::fabs(v); }
        note_num("sum_" + name, s1); note_num("sumsq_" + name, s2); note_num("sumabs_" + name, sa);
      }
      note_num("len_" + name, n);
      if (hash_only or not out.is_open()) return;
      std::uint32_t nl(name.size()), nd(dims.size());
      out.write(reinterpret_cast<const char*>(&nl), 4);
      out.write(name.data(), nl);
      out.write(reinterpret_cast<const char*>(&dtype), 1);
      out.write(reinterpret_cast<const char*>(&nd), 4);
      for (auto d : dims) out.write(reinterpret_cast<const char*>(&d), 8);
      out.write(reinterpret_cast<const char*>(data), n * sizeof(T));
    }
    void put_u32(const std::string& name, const std::vector<unsigned int>& v, std::vector<std::uint64_t> dims) { put(name, 0, v.data(), dims); }
    void put_i32(const std::string& name, const std::vector<int>& v, std::vector<std::uint64_t> dims) { put(name, 1, v.data(), dims); }
    void put_f64(const std::string& name, const std::vector<double>& v, std::vector<std::uint64_t> dims) { put(name, 2, v.data(), dims); }
  };

  template<typename mesh_type>
  void dump_mesh(dumper& dmp, const mesh_type& m) {
    const std::size_t V(m.get_vertex_number()), K(m.get_cell_number());
    const std::size_t dim(m.get_embedding_space_dimension());
    const std::size_t nv(m.get_cells().get_size(1));
    std::vector<double> vertices(V * dim);
    std::copy(&m.get_vertices().at(0, 0), &m.get_vertices().at(0, 0) + V * dim, vertices.begin());
    std::vector<unsigned int> cells(K * nv);
    std::copy(&m.get_cells().at(0, 0), &m.get_cells().at(0, 0) + K * nv, cells.begin());
    dmp.put_f64("vertices", vertices, {V, dim});
    dmp.put_u32("cells", cells, {K, nv});
    std::vector<double> vol(K), jmt(K * dim * dim);
    for (std::size_t k(0); k < K; ++k) {
      vol[k] = m.get_cell_volume(k);
      const array<double>& j(m.get_jmt(k));
      for (std::size_t s(0); s < dim; ++s)
        for (std::size_t t(0); t < dim; ++t)
          jmt[(k * dim + s) * dim + t] = j.at(s, t);
    }
    dmp.put_f64("cell_volume", vol, {K});
    dmp.put_f64("jmt", jmt, {K, dim, dim});
    // face-neighbour table of mesh<cell>::compute_cell_neighbours (mesh.hpp:301-332), -1 on the boundary
    std::vector<int> nb(K * nv);
    for (std::size_t k(0); k < K; ++k)
      for (std::size_t j(0); j < nv; ++j) nb[k * nv + j] = static_cast<int>(m.get_cell_neighbour(k, j));
    dmp.put_i32("cell_neighbours", nb, {K, nv});
  }

  template<typename submesh_type>
  void dump_submesh(dumper& dmp, const std::string& prefix, const submesh_type& dm) {
    const std::size_t F(dm.get_cell_number());
    const std::size_t nv(F ? dm.get_cells().get_size(1) : 0);
    std::vector<unsigned int> el(F * nv), el_id(F), sd_id(F);
    for (std::size_t k(0); k < F; ++k) {
      for (std::size_t i(0); i < nv; ++i) el[k * nv + i] = dm.get_cells().at(k, i);
      el_id[k] = dm.get_parent_cell_id(k);
      sd_id[k] = dm.get_subdomain_id(k);
    }
    dmp.put_u32(prefix + "_cells", el, {F, nv});
    dmp.put_u32(prefix + "_parent_cell", el_id, {F});
    dmp.put_u32(prefix + "_parent_subdomain", sd_id, {F});
  }

  template<typename fes_type>
  void dump_fes(dumper& dmp, const std::string& prefix, const fes_type& fes, std::size_t K) {
    const std::size_t nd(fes_type::fe_type::n_dof_per_element);
    const std::size_t N(fes.get_dof_number());
    std::vector<unsigned int> dof_map(K * nd), g2l(N * 2);
    for (std::size_t k(0); k < K; ++k)
      for (std::size_t i(0); i < nd; ++i)
        dof_map[k * nd + i] = fes.get_dof(k, i);
    for (std::size_t i(0); i < N; ++i) {
      g2l[2 * i] = fes.get_dof_element(i);
      g2l[2 * i + 1] = fes.get_dof_local_id(i);
    }
    dmp.put_u32(prefix + "dof_map", dof_map, {K, nd});
    dmp.put_u32(prefix + "g2l", g2l, {N, 2});
    std::vector<unsigned int> ids; std::vector<double> vals;
    for (const auto& kv : fes.get_dirichlet_dof_values()) { ids.push_back(kv.first); vals.push_back(kv.second); }
    dmp.put_u32(prefix + "dirichlet_dof", ids, {ids.size()});
    dmp.put_f64(prefix + "dirichlet_value", vals, {vals.size()});
    dmp.note_num(prefix + "n_dof", N);
  }

  void dump_captured_system(dumper& dmp) {
    oracle_shim::capture_t& c(oracle_shim::capture());
    if (not c.A) return;
    const std::size_t N(c.A->n_row), nnz(c.A->colind.size());
    dmp.put_i32("csr_rowptr", c.A->rowptr, {N + 1});
    dmp.put_i32("csr_colind", c.A->colind, {nnz});
    dmp.put_f64("csr_val", c.A->val, {nnz});
    dmp.put_f64("rhs", c.b->v, {c.b->v.size()});
    std::size_t zeros(0);
    for (double v : c.A->val) zeros += (v == 0.0);
    dmp.note_num("nnz", nnz); dmp.note_num("n_row", N); dmp.note_num("n_exact_zero", zeros);
  }

  struct run_options {
    std::string problem, fe;
    std::size_t n;
    std::string out;
    bool hash_only;
    std::string solve;   // "", "gmres", "cg"
    double solve_rtol;
    unsigned n_thread;
    unsigned repeat;     // poisson: passes of (a += ..., b += ..., a.solve(b, s)) on one mesh / space (bench.py's reference arm)
    run_options() : n(4), hash_only(false), solve_rtol(1e-8), n_thread(std::thread::hardware_concurrency()), repeat(1) {}
  };

  std::vector<double> g_solution;
  int g_its(0);
  double g_rnorm(0.0), g_solve_s(0.0);

  void install_solver(const run_options& opt) {
    oracle_shim::capture_t& c(oracle_shim::capture());
    if (opt.solve.empty()) { c.solve_hook = nullptr; return; }
    const bool cg(opt.solve == "cg");
    const double rtol(opt.solve_rtol);
    c.solve_hook = [cg, rtol](KSP ksp, Mat A, Vec b, Vec x) -> int {
      const auto t0(clk::now());
      int its(0);
      if (cg)
        its = orc_cg_jacobi(A->n_row, A->rowptr.data(), A->colind.data(), A->val.data(),
                            b->v.data(), x->v.data(), rtol, ksp->atol, ksp->maxits > 0 ? 1000000 : 0, &g_rnorm);
      else
        its = orc_gmres_ilu(A->n_row, A->rowptr.data(), A->colind.data(), A->val.data(),
                            b->v.data(), x->v.data(), ksp->restart, ksp->pc.levels,
                            rtol, ksp->atol, ksp->dtol, ksp->maxits, &g_rnorm);
      g_solve_s = seconds_since(t0);
      g_its = its;
      g_solution = x->v;
      return its;
    };
  }

  dictionary solver_parameters() {
    // README.md:100-106 (the authoritative key set, solver.hpp:125-133)
    return dictionary()
      .set("maxits",  2000u)
      .set("restart", 1000u)
      .set("rtol",    1.e-8)
      .set("atol",    1.e-50)
      .set("dtol",    1.e20)
      .set("ilufill", 2u);
  }

  void finish(dumper& dmp, const run_options& opt) {
    dump_captured_system(dmp);
    if (not opt.solve.empty()) {
      dmp.put_f64("solution", g_solution, {g_solution.size()});
      dmp.note_num("solve_its", g_its); dmp.note_num("solve_rnorm", g_rnorm); dmp.note_num("t_solve_s", g_solve_s);
    }
    std::cout << "{" << dmp.meta.str() << "}" << std::endl;
  }

  double one(const double*) { return 1.0; }

  // examples/poisson.cpp:4-14
  double gaussian_src(const double* x) {
    const double sigma(0.1);
    const double r(std::sqrt(std::pow(x[0] - 0.5, 2) + std::pow(x[1] - 0.5, 2)));
    return std::exp(- r * r / (2.0 * sigma * sigma));
  }
  double g_bc(const double* x) { return 1.0 + x[0] + 2.0 * x[1] + x[0] * x[1]; }
  double c_reac(const double* x) { return 1.0 + x[0] * x[0] + 0.5 * x[1]; }

  template<typename fe_type>
  void run_poisson(const run_options& opt, bool general) {
    using cell_type = cell::triangle;
    using mesh_type = fe_mesh<cell_type>;
    using fes_type = finite_element_space<fe_type>;
    using quad_type = quad::triangle::qf5pT;
    dumper dmp(opt.out, opt.hash_only);
    const std::size_t n(opt.n);

    auto t0(clk::now());
    mesh_type m(gen_square_mesh(1.0, 1.0, n, n));
    dmp.note_num("t_mesh_s", seconds_since(t0));
    t0 = clk::now();
    submesh<cell_type> dm(m.get_boundary_submesh());
    dmp.note_num("t_boundary_s", seconds_since(t0));
    t0 = clk::now();
    fes_type fes(m);
    dmp.note_num("t_fes_s", seconds_since(t0));
    t0 = clk::now();
    if (general) fes.add_dirichlet_boundary(dm, g_bc);
    else fes.add_dirichlet_boundary(dm);
    dmp.note_num("t_dirichlet_s", seconds_since(t0));

    // bench.py's reference arm: repeat - 1 extra passes of the hot path on the same mesh and space, each timed like the
    // last one (which is also the one that is dumped / hashed below)
    std::ostringstream passes;
    for (unsigned pass(0); pass + 1 < opt.repeat; ++pass) {
      auto p0(clk::now());
      bilinear_form<fes_type, fes_type> ar(fes, fes); {
        auto u(ar.get_trial_function());
        auto v(ar.get_test_function());
        if (general)
          ar += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + c_reac * (u * v), m);
        else
          ar += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v), m);
      }
      const double tb(seconds_since(p0));
      p0 = clk::now();
      linear_form<fes_type> br(fes, 0, opt.n_thread); {
        auto v(br.get_test_function());
        if (general) br += integrate<quad_type>(gaussian_src * v, m);
        else br += integrate<quad_type>(one * v, m);
      }
      const double tl(seconds_since(p0));
      install_solver(opt);
      solver::petsc::gmres_ilu sr(solver_parameters());
      p0 = clk::now();
      const typename fes_type::element ur(ar.solve(br, sr));
      const double ts_call(seconds_since(p0));
      passes << (pass ? "," : "") << "[" << tb << "," << tl << "," << g_solve_s << "," << ts_call << "," << g_its << "]";
    }
    if (opt.repeat > 1) dmp.note("passes_bilinear_linear_solve_solvecall_its", "[" + passes.str() + "]");

    t0 = clk::now();
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      auto u(a.get_trial_function());
      auto v(a.get_test_function());
      if (general)
        a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + c_reac * (u * v), m);
      else
        a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v), m);
    }
    const double t_bilinear(seconds_since(t0));
    dmp.note_num("t_bilinear_s", t_bilinear);
    dmp.note_num("cells_per_s_bilinear", m.get_cell_number() / t_bilinear);

    t0 = clk::now();
    linear_form<fes_type> b(fes, 0, opt.n_thread); {
      auto v(b.get_test_function());
      if (general) b += integrate<quad_type>(gaussian_src * v, m);
      else b += integrate<quad_type>(one * v, m);
    }
    dmp.note_num("t_linear_s", seconds_since(t0));
    dmp.note_num("n_thread_linear", opt.n_thread);

    if (not opt.hash_only) {
      dump_mesh(dmp, m);
      dump_submesh(dmp, "boundary", dm);
    }
    dump_fes(dmp, "", fes, m.get_cell_number());
    {
      const std::size_t N(fes.get_dof_number());
      std::vector<double> coef(N);
      std::copy(&b.get_coefficients().at(0), &b.get_coefficients().at(0) + N, coef.begin());
      dmp.put_f64("linear_form", coef, {N});
    }

    install_solver(opt);
    solver::petsc::gmres_ilu s(solver_parameters());
    t0 = clk::now();
    const typename fes_type::element u(a.solve(b, s));
    dmp.note_num("t_solve_call_s", seconds_since(t0));
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }

  double tet_u_bc(const double* x) { return x[0]*x[0] + x[1]*x[1] + x[2]*x[2]; }
  double tet_f(const double*) { return -6.0; }

  template<typename fe_type, typename quad_type>
  void run_tet(const run_options& opt) {
    using cell_type = cell::tetrahedron;
    using fes_type = finite_element_space<fe_type>;
    dumper dmp(opt.out, opt.hash_only);
    const std::size_t n(opt.n);

    auto t0(clk::now());
    fe_mesh<cell_type> m(gen_cube_mesh(1.0, 1.0, 1.0, n, n, n));
    dmp.note_num("t_mesh_s", seconds_since(t0));
    t0 = clk::now();
    submesh<cell_type> dm(m.get_boundary_submesh());
    dmp.note_num("t_boundary_s", seconds_since(t0));
    t0 = clk::now();
    fes_type fes(m);
    dmp.note_num("t_fes_s", seconds_since(t0));
    t0 = clk::now();
    fes.add_dirichlet_boundary(dm, tet_u_bc);
    dmp.note_num("t_dirichlet_s", seconds_since(t0));

    t0 = clk::now();
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      const auto u(a.get_trial_function());
      const auto v(a.get_test_function());
      a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + d<3>(u) * d<3>(v), m);
    }
    const double t_bilinear(seconds_since(t0));
    dmp.note_num("t_bilinear_s", t_bilinear);
    dmp.note_num("cells_per_s_bilinear", m.get_cell_number() / t_bilinear);

    t0 = clk::now();
    linear_form<fes_type> f(fes, 0, opt.n_thread); {
      const auto v(f.get_test_function());
      f += integrate<quad_type>(tet_f * v, m);
    }
    dmp.note_num("t_linear_s", seconds_since(t0));

    // test/tetrahedron_cube.cpp:56-58
    const double int_x2(integrate<quad_type>(make_expr(std::function<double(const double*)>([](const double* x){ return x[1]; })), m));
    dmp.note_num("integral_x2", int_x2);

    if (not opt.hash_only) {
      dump_mesh(dmp, m);
      dump_submesh(dmp, "boundary", dm);
    }
    dump_fes(dmp, "", fes, m.get_cell_number());
    {
      const std::size_t N(fes.get_dof_number());
      std::vector<double> coef(N);
      std::copy(&f.get_coefficients().at(0), &f.get_coefficients().at(0) + N, coef.begin());
      dmp.put_f64("linear_form", coef, {N});
    }
    install_solver(opt);
    solver::petsc::gmres_ilu s(solver_parameters());
    const typename fes_type::element u(a.solve(f, s));
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }

  // test/stokes_2d_p2_p1.cpp:36-54
  double st_f0(const double*) { return 0.0; }
  double st_f1(const double*) { return 0.0; }
  double st_u0_bv(const double* x) { return x[1] < 0.0001 ? x[0] * (1.0 - x[0]) : 0.0; }
  double st_u1_bv(const double* x) { return x[0] < 0.0001 ? x[1] * (1.0 - x[1]) : 0.0; }
  double st_p_v(const double*) { return 0.0; }

  template<typename cfes_type, typename u_fe, typename p_fe>
  void dump_cfes(dumper& dmp, const cfes_type& fes, std::size_t K) {
    dump_fes(dmp, "c0_", fes.template get_finite_element_space<0>(), K);
    dump_fes(dmp, "c1_", fes.template get_finite_element_space<1>(), K);
    dump_fes(dmp, "c2_", fes.template get_finite_element_space<2>(), K);
  }

  void run_stokes(const run_options& opt) {
    using cell_type = cell::triangle;
    using u_fe_type = cell_type::fe::lagrange_p2;
    using p_fe_type = cell_type::fe::lagrange_p1;
    using quad_type = quad::triangle::qf5pT;
    dumper dmp(opt.out, opt.hash_only);
    const std::size_t n(opt.n);

    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    submesh<cell_type, cell::point> pinned_pressure_point(m.get_point_submesh(n * n / 2 + n / 2));

    using fe_type = composite_finite_element<u_fe_type, u_fe_type, p_fe_type>;
    using fes_type = composite_finite_element_space<fe_type>;
    auto t0(clk::now());
    fes_type fes(m);
    dmp.note_num("t_fes_s", seconds_since(t0));

    t0 = clk::now();
    fes.add_dirichlet_boundary<0>(dm, st_u0_bv);
    fes.add_dirichlet_boundary<1>(dm, st_u1_bv);
    fes.add_dirichlet_boundary<2>(pinned_pressure_point, st_p_v);
    dmp.note_num("t_dirichlet_s", seconds_since(t0));

    t0 = clk::now();
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      auto v0(a.get_test_function<0>());
      auto v1(a.get_test_function<1>());
      auto q (a.get_test_function<2>());
      auto u0(a.get_trial_function<0>());
      auto u1(a.get_trial_function<1>());
      auto p (a.get_trial_function<2>());
      a += integrate<quad_type>(    d<1>(u0) * d<1>(v0) + d<2>(u0) * d<2>(v0)
                                  + d<1>(u1) * d<1>(v1) + d<2>(u1) * d<2>(v1)
                                  + p * (d<1>(v0) + d<2>(v1))
                                  + q * (d<1>(u0) + d<2>(u1))
                                , m);
    }
    const double t_bilinear(seconds_since(t0));
    dmp.note_num("t_bilinear_s", t_bilinear);
    dmp.note_num("cells_per_s_bilinear", m.get_cell_number() / t_bilinear);

    t0 = clk::now();
    linear_form<fes_type> f(fes); {
      auto v0(f.get_test_function<0>());
      auto v1(f.get_test_function<1>());
      f += integrate<quad_type>(st_f0 * v0 + st_f1 * v1, m);
    }
    dmp.note_num("t_linear_s", seconds_since(t0));

    if (not opt.hash_only) {
      dump_mesh(dmp, m);
      dump_submesh(dmp, "boundary", dm);
      dump_submesh(dmp, "pin", pinned_pressure_point);
    }
    dump_cfes<fes_type, u_fe_type, p_fe_type>(dmp, fes, m.get_cell_number());
    {
      const std::size_t N(fes.get_total_dof_number());
      std::vector<double> coef(N);
      std::copy(&f.get_coefficients().at(0), &f.get_coefficients().at(0) + N, coef.begin());
      dmp.put_f64("linear_form", coef, {N});
    }
    install_solver(opt);
    solver::petsc::gmres_ilu s(solver_parameters());
    const fes_type::element x(a.solve(f, s));
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }

  // test/navier_stokes_2d_p2_p1.cpp:13-31 (regularised driven south and east boundaries)
  double ns_f0(const double*) { return 0.0; }
  double ns_f1(const double*) { return 0.0; }
  double ns_u0_bv(const double* x) { return x[1] < 0.00001 ? 4.0 * x[0] * (1.0 - x[0]) : 0.0; }
  double ns_u1_bv(const double* x) { return x[0] < 0.00001 ? 4.0 * x[1] * (1.0 - x[1]) : 0.0; }
  double ns_p_v(const double*) { return 0.0; }
  // synthetic smooth "previous iterates" (not in the reference; any P2 field will do)
  double ns_prev0(const double* x) { return std::sin(3.0 * x[0]) * std::cos(2.0 * x[1]) + 0.25; }
  double ns_prev1(const double* x) { return x[0] * x[1] - 0.5 * x[1] * x[1] + 0.1; }
  double ns_old0(const double* x) { return 0.5 * x[0] - x[1] * x[1]; }
  double ns_old1(const double* x) { return std::cos(x[0] + 2.0 * x[1]); }

  void run_ns(const run_options& opt) {
    using cell_type = cell::triangle;
    using u_fe_type = cell_type::fe::lagrange_p2;
    using p_fe_type = cell_type::fe::lagrange_p1;
    using quad_type = quad::triangle::qf5pT;
    dumper dmp(opt.out, opt.hash_only);
    const std::size_t n(opt.n);

    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    submesh<cell_type, cell::point> pinned_pressure_point(m.get_point_submesh(n * n / 2 + n / 2));

    using fe_type = composite_finite_element<u_fe_type, u_fe_type, p_fe_type>;
    using fes_type = composite_finite_element_space<fe_type>;
    fes_type fes(m);
    fes.add_dirichlet_boundary<0>(dm, ns_u0_bv);
    fes.add_dirichlet_boundary<1>(dm, ns_u1_bv);
    fes.add_dirichlet_boundary<2>(pinned_pressure_point, ns_p_v);

    const std::size_t Nu(fes.get_dof_number<0>());
    array<double> c_prev0{Nu}, c_prev1{Nu}, c_old0{Nu}, c_old1{Nu};
    for (std::size_t i(0); i < Nu; ++i) {
      const array<double> x(fes.get_dof_space_coordinate<0>(i));
      c_prev0.at(i) = ns_prev0(&x.at(0, 0));
      c_prev1.at(i) = ns_prev1(&x.at(0, 0));
      c_old0.at(i) = ns_old0(&x.at(0, 0));
      c_old1.at(i) = ns_old1(&x.at(0, 0));
    }
    finite_element_space<u_fe_type>::element
      v0_h(fes.get_finite_element_space<0>(), c_prev0),
      v1_h(fes.get_finite_element_space<1>(), c_prev1),
      u0_h(fes.get_finite_element_space<0>(), c_old0),
      u1_h(fes.get_finite_element_space<1>(), c_old1);

    const double reynolds(5000.0);
    const double time_step(0.5);

    auto v0(make_expr<u_fe_type>(v0_h));
    auto v1(make_expr<u_fe_type>(v1_h));

    auto t0(clk::now());
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      auto w0(a.get_test_function<0>());
      auto w1(a.get_test_function<1>());
      auto q (a.get_test_function<2>());
      auto vn0(a.get_trial_function<0>());
      auto vn1(a.get_trial_function<1>());
      auto pn (a.get_trial_function<2>());
      // test/navier_stokes_2d_p2_p1.cpp:319-338
      a += integrate<quad_type>(vn0 * w0 + vn1 * w1 +
                                time_step * (vn0 * d<1>(v0) * w0 +
                                             vn0 * d<1>(v1) * w1 +
                                             vn1 * d<2>(v0) * w0 +
                                             vn1 * d<2>(v1) * w1) +
                                time_step * (v0 * d<1>(vn0) * w0 +
                                             v0 * d<1>(vn1) * w1 +
                                             v1 * d<2>(vn0) * w0 +
                                             v1 * d<2>(vn1) * w1) +
                                (time_step / reynolds) * (d<1>(vn0) * d<1>(w0) + d<2>(vn0) * d<2>(w0) +
                                                          d<1>(vn1) * d<1>(w1) + d<2>(vn1) * d<2>(w1)) +
                                time_step * pn * (d<1>(w0) + d<2>(w1)) +
                                time_step * q * (d<1>(vn0) + d<2>(vn1))
                                , m);
    }
    const double t_bilinear(seconds_since(t0));
    dmp.note_num("t_bilinear_s", t_bilinear);
    dmp.note_num("cells_per_s_bilinear", m.get_cell_number() / t_bilinear);

    t0 = clk::now();
    linear_form<fes_type> f(fes); {
      auto w0(f.get_test_function<0>());
      auto w1(f.get_test_function<1>());
      auto u0(make_expr<u_fe_type>(u0_h));
      auto u1(make_expr<u_fe_type>(u1_h));
      // test/navier_stokes_2d_p2_p1.cpp:354-363
      f += integrate<quad_type>( u0 * w0 + u1 * w1 +
                                 time_step * (ns_f0 * w0 +
                                              ns_f1 * w1) +
                                 time_step * (v0 * d<1>(v0) * w0 +
                                              v0 * d<1>(v1) * w1 +
                                              v1 * d<2>(v0) * w0 +
                                              v1 * d<2>(v1) * w1)
                                 , m);
    }
    dmp.note_num("t_linear_s", seconds_since(t0));

    // rank-0 integral used by the Newton stopping test (:405-414)
    const double norm0(std::sqrt(integrate<quad_type>((make_expr<u_fe_type>(v0_h) - make_expr<u_fe_type>(u0_h)) *
                                                      (make_expr<u_fe_type>(v0_h) - make_expr<u_fe_type>(u0_h)), m)));
    dmp.note_num("l2_prev0_minus_old0", norm0);

    if (not opt.hash_only) {
      dump_mesh(dmp, m);
      dump_submesh(dmp, "boundary", dm);
      dump_submesh(dmp, "pin", pinned_pressure_point);
      std::vector<double> tmp(Nu);
      std::copy(&c_prev0.at(0), &c_prev0.at(0) + Nu, tmp.begin()); dmp.put_f64("coef_prev0", tmp, {Nu});
      std::copy(&c_prev1.at(0), &c_prev1.at(0) + Nu, tmp.begin()); dmp.put_f64("coef_prev1", tmp, {Nu});
      std::copy(&c_old0.at(0), &c_old0.at(0) + Nu, tmp.begin()); dmp.put_f64("coef_old0", tmp, {Nu});
      std::copy(&c_old1.at(0), &c_old1.at(0) + Nu, tmp.begin()); dmp.put_f64("coef_old1", tmp, {Nu});
    }
    dump_cfes<fes_type, u_fe_type, p_fe_type>(dmp, fes, m.get_cell_number());
    {
      const std::size_t N(fes.get_total_dof_number());
      std::vector<double> coef(N);
      std::copy(&f.get_coefficients().at(0), &f.get_coefficients().at(0) + N, coef.begin());
      dmp.put_f64("linear_form", coef, {N});
    }
    install_solver(opt);
    solver::petsc::gmres_ilu s(solver_parameters());
    const fes_type::element x(a.solve(f, s));
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }

  template<typename Q>
  void dump_quadrature(dumper& dmp, const std::string& name) {
    const std::size_t nq(Q::n_point), dim(Q::n_point_space_dimension);
    std::vector<double> x(nq * dim), w(nq);
    for (std::size_t q(0); q < nq; ++q) {
      w[q] = Q::w[q];
      for (std::size_t s(0); s < dim; ++s) x[q * dim + s] = Q::x[q][s];
    }
    dmp.put_f64(name + "_x", x, {nq, dim});
    dmp.put_f64(name + "_w", w, {nq});
  }

  template<typename FE>
  void dump_fe_at(dumper& dmp, const std::string& name, const double* pts, std::size_t npts) {
    // phi / dphi tables at arbitrary reference points, for FE-table parity
    const std::size_t nd(FE::n_dof_per_element), dim(FE::cell_type::n_dimension);
    std::vector<double> v(npts * nd * (dim + 1)), nodes(nd * dim);
    for (std::size_t q(0); q < npts; ++q)
      for (std::size_t i(0); i < nd; ++i) {
        v[(q * nd + i) * (dim + 1)] = FE::phi(i, pts + q * dim);
        for (std::size_t s(0); s < dim; ++s)
          v[(q * nd + i) * (dim + 1) + 1 + s] = FE::dphi(s, i, pts + q * dim);
      }
    for (std::size_t i(0); i < nd; ++i)
      for (std::size_t s(0); s < dim; ++s) nodes[i * dim + s] = FE::x[i][s];
    dmp.put_f64(name + "_values", v, {npts, nd, dim + 1});
    dmp.put_f64(name + "_nodes", nodes, {nd, dim});
  }

  void run_tables(const run_options& opt) {
    dumper dmp(opt.out, opt.hash_only);
    dump_quadrature<quad::edge::gauss1>(dmp, "edge_gauss1");
    dump_quadrature<quad::edge::gauss2>(dmp, "edge_gauss2");
    dump_quadrature<quad::edge::gauss3>(dmp, "edge_gauss3");
    dump_quadrature<quad::edge::gauss4>(dmp, "edge_gauss4");
    dump_quadrature<quad::edge::gauss5>(dmp, "edge_gauss5");
    dump_quadrature<quad::triangle::qf1pT>(dmp, "tri_qf1pT");
    dump_quadrature<quad::triangle::qf2pT>(dmp, "tri_qf2pT");
    dump_quadrature<quad::triangle::qf5pT>(dmp, "tri_qf5pT");
    dump_quadrature<quad::triangle::qf1pTlump>(dmp, "tri_qf1pTlump");
    dump_quadrature<quad::tetrahedron::qf1pTet>(dmp, "tet_qf1pTet");
    dump_quadrature<quad::tetrahedron::qf4pTet>(dmp, "tet_qf4pTet");
    dump_quadrature<quad::tetrahedron::qf5pTet>(dmp, "tet_qf5pTet");
    dump_quadrature<quad::tetrahedron::qfSym1pTet>(dmp, "tet_qfSym1pTet");
    dump_quadrature<quad::tetrahedron::qfSym4pTet>(dmp, "tet_qfSym4pTet");
    dump_quadrature<quad::tetrahedron::qfSym10pTet>(dmp, "tet_qfSym10pTet");
    dump_quadrature<quad::tetrahedron::qfSym20pTet>(dmp, "tet_qfSym20pTet");
    dump_quadrature<quad::tetrahedron::qfSym35pTet>(dmp, "tet_qfSym35pTet");
    dump_quadrature<quad::tetrahedron::qfSym56pTet>(dmp, "tet_qfSym56pTet");
    const double tri_pts[] = {0.2, 0.3,  0.0, 0.0,  1.0, 0.0,  0.0, 1.0,  1.0/3.0, 1.0/3.0,  0.61, 0.17};
    const double tet_pts[] = {0.1, 0.2, 0.3,  0.0, 0.0, 0.0,  1.0, 0.0, 0.0,  0.0, 1.0, 0.0,  0.0, 0.0, 1.0,  0.25, 0.25, 0.25,  0.4, 0.05, 0.31};
    dump_fe_at<cell::triangle::fe::lagrange_p1>(dmp, "tri_p1", tri_pts, 6);
    dump_fe_at<cell::triangle::fe::lagrange_p1_bubble>(dmp, "tri_p1b", tri_pts, 6);
    dump_fe_at<cell::triangle::fe::lagrange_p2>(dmp, "tri_p2", tri_pts, 6);
    dump_fe_at<cell::tetrahedron::fe::lagrange_p1>(dmp, "tet_p1", tet_pts, 7);
    dump_fe_at<cell::tetrahedron::fe::lagrange_p1_bubble>(dmp, "tet_p1b", tet_pts, 7);
    std::cout << "{" << dmp.meta.str() << "}" << std::endl;
  }

  double l2_f0(const double* x) { return 1.0 + x[0]; }
  double l2_f1(const double* x) { return std::sin(3.0 * x[1]); }

  void run_laplace2(const run_options& opt) {
    // test/composite_fe.cpp:82-131
    using cell_type = cell::triangle;
    using fe_0_type = cell::triangle::fe::lagrange_p1;
    using fe_1_type = cell::triangle::fe::lagrange_p1;
    using fe_type = composite_finite_element<fe_0_type, fe_1_type>;
    using fes_type = composite_finite_element_space<fe_type>;
    dumper dmp(opt.out, opt.hash_only);
    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, opt.n, opt.n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    fes_type fes(m);
    fes.add_dirichlet_boundary<0>(dm);
    fes.add_dirichlet_boundary<1>(dm);
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      const auto u_0(a.get_trial_function<0>());
      const auto u_1(a.get_trial_function<1>());
      const auto v_0(a.get_test_function<0>());
      const auto v_1(a.get_test_function<1>());
      a += integrate<quad::triangle::qf5pT>(d<1>(u_0) * d<1>(v_0) + d<2>(u_0) * d<2>(v_0)
                                            + d<1>(u_1) * d<1>(v_1) + d<2>(u_1) * d<2>(v_1), m);
    }
    linear_form<fes_type> f(fes); {
      const auto v_0(f.get_test_function<0>());
      const auto v_1(f.get_test_function<1>());
      f += integrate<quad::triangle::qf5pT>(l2_f0 * v_0 + l2_f1 * v_1, m);
    }
    if (not opt.hash_only) { dump_mesh(dmp, m); dump_submesh(dmp, "boundary", dm); }
    dump_fes(dmp, "c0_", fes.get_finite_element_space<0>(), m.get_cell_number());
    dump_fes(dmp, "c1_", fes.get_finite_element_space<1>(), m.get_cell_number());
    {
      const std::size_t N(fes.get_total_dof_number());
      std::vector<double> coef(N);
      std::copy(&f.get_coefficients().at(0), &f.get_coefficients().at(0) + N, coef.begin());
      dmp.put_f64("linear_form", coef, {N});
    }
    install_solver(opt);
    solver::petsc::gmres_ilu s(solver_parameters());
    const fes_type::element x(a.solve(f, s));
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }
}

namespace {
  // a1 / a2 / a4 at full size: vertices, cells, geometry, face neighbours and the boundary submesh of the generators
  template<typename cell_type>
  void run_mesh(const run_options& opt, mesh<cell_type>&& generated) {
    dumper dmp(opt.out, opt.hash_only);
    auto t0(clk::now());
    fe_mesh<cell_type> m(std::move(generated));
    dmp.note_num("t_fe_mesh_s", seconds_since(t0));
    t0 = clk::now();
    submesh<cell_type> dm(m.get_boundary_submesh());
    dmp.note_num("t_boundary_s", seconds_since(t0));
    dump_mesh(dmp, m);
    dump_submesh(dmp, "boundary", dm);
    dmp.note_num("n_cells", m.get_cell_number());
    finish(dmp, opt);
  }

  double export_g(const double* x) { return 1.0 + x[0] + 2.0 * x[1] + x[0] * x[1]; }
  double export_h(const double* x) { return x[0] * x[0] - 0.5 * x[1]; }
  double export_t(const double* x) { return 1.0 + x[0] - x[1] + 2.0 * x[2]; }

  // every exporter entry point on small meshes: scalar per node / per element, a two-component vector per node, P2
  // vertex data, the plain ASCII dump, a tetrahedral geometry and a transient series
  void run_export(const run_options& opt) {
    const std::string dir(opt.out);
    {
      using cell_type = cell::triangle;
      using mesh_type = fe_mesh<cell_type>;
      using p1 = cell_type::fe::lagrange_p1;
      using p2 = cell_type::fe::lagrange_p2;
      mesh_type m(gen_square_mesh(1.0, 1.0, opt.n, opt.n));
      finite_element_space<p1> fes1(m);
      finite_element_space<p2> fes2(m);
      const auto g1(projector::lagrange<p1>(export_g, fes1));
      const auto h1(projector::lagrange<p1>(export_h, fes1));
      const auto g2(projector::lagrange<p2>(export_g, fes2));
      exporter::ensight6(dir + "/square", to_mesh_vertex_data<p1>(g1), "g", to_mesh_cell_data<p1>(h1), "hc",
                         to_mesh_vertex_data<p2>(g2), "g2");
      const auto dg(to_mesh_vertex_data<p1>(g1));
      const auto dh(to_mesh_vertex_data<p1>(h1));
      exporter::ensight6(dir + "/vector", mesh_data<double, mesh_type>(m, mesh_data_kind::vertex, dg, dh), "gh");
      exporter::ascii<mesh_type>(dir + "/square_vertex.dat", to_mesh_vertex_data<p1>(g1));
      exporter::ascii<mesh_type>(dir + "/square_cell.dat", to_mesh_cell_data<p2>(g2));
      exporter::ensight6_geometry(dir + "/geometry_only", m);
      exporter::ensight6_transient<mesh_type> series(dir + "/series", m, "g");
      series.export_time_step(0.0, to_mesh_vertex_data<p1>(g1));
      series.export_time_step(0.25, to_mesh_vertex_data<p1>(h1));
    }
    {
      using cell_type = cell::tetrahedron;
      using mesh_type = fe_mesh<cell_type>;
      using p1 = cell_type::fe::lagrange_p1;
      mesh_type m(gen_cube_mesh(1.0, 1.0, 1.0, 2, 2, 2));
      finite_element_space<p1> fes(m);
      const auto t1(projector::lagrange<p1>(export_t, fes));
      exporter::ensight6(dir + "/cube", to_mesh_vertex_data<p1>(t1), "t", to_mesh_cell_data<p1>(t1), "tc");
    }
  }
}

int main(int argc, char* argv[]) {
  try {
    run_options opt;
    for (int i(1); i < argc; ++i) {
      const std::string a(argv[i]);
      if (a == "--problem") opt.problem = argv[++i];
      else if (a == "--fe") opt.fe = argv[++i];
      else if (a == "--n") opt.n = std::stoul(argv[++i]);
      else if (a == "--out") opt.out = argv[++i];
      else if (a == "--hash-only") opt.hash_only = true;
      else if (a == "--repeat") opt.repeat = static_cast<unsigned>(std::stoul(argv[++i]));
      else if (a == "--solve") opt.solve = argv[++i];
      else if (a == "--solve-rtol") opt.solve_rtol = std::stod(argv[++i]);
      else if (a == "--threads") opt.n_thread = std::stoul(argv[++i]);
      else throw std::string("unknown argument ") + a;
    }
    if (opt.problem == "poisson" or opt.problem == "poissong") {
      const bool general(opt.problem == "poissong");
      if (opt.fe == "p1" or opt.fe.empty()) run_poisson<cell::triangle::fe::lagrange_p1>(opt, general);
      else if (opt.fe == "p2") run_poisson<cell::triangle::fe::lagrange_p2>(opt, general);
      else if (opt.fe == "p1b") run_poisson<cell::triangle::fe::lagrange_p1_bubble>(opt, general);
      else throw std::string("unknown fe");
    } else if (opt.problem == "tet") {
      if (opt.fe == "p1" or opt.fe.empty()) run_tet<cell::tetrahedron::fe::lagrange_p1, quad::tetrahedron::qfSym4pTet>(opt);
      else if (opt.fe == "p1b") run_tet<cell::tetrahedron::fe::lagrange_p1_bubble, quad::tetrahedron::qfSym35pTet>(opt);
      else throw std::string("unknown fe");
    } else if (opt.problem == "stokes") run_stokes(opt);
    else if (opt.problem == "ns") run_ns(opt);
    else if (opt.problem == "laplace2") run_laplace2(opt);
    else if (opt.problem == "tables") run_tables(opt);
    else if (opt.problem == "export") run_export(opt);
    else if (opt.problem == "mesh") run_mesh<cell::triangle>(opt, gen_square_mesh(1.0, 1.0, opt.n, opt.n));
    else if (opt.problem == "meshcube") run_mesh<cell::tetrahedron>(opt, gen_cube_mesh(1.0, 1.0, 1.0, opt.n, opt.n, opt.n));
    else throw std::string("unknown problem");
  } catch (const std::string& e) {
    std::cerr << "ref_driver: " << e << std::endl;
    return 1;
  }
  return 0;
}
