// test/laplacian_transient.cpp of the reference on the CUDA path: implicit Euler for c_t - lap c = f on the unit square,
// P1, homogeneous Dirichlet data, the bilinear form assembled once and solved at every step, the right-hand side reassembled
// per step from the previous iterate (a coefficient field) and host functions of x, every step exported with
// exporter::ensight6_transient. Differences from the reference program: the solver line (solver::cuda::cg_jacobi), the
// dictionary key "atol" (the reference passes "abstol" and throws at solver construction), the mesh size and
// step count taken from the command line, and a check against the exact solution c(x, t) = (1 - t) sin(pi x) sin(pi y),
// which implicit Euler reproduces in time (it is linear in t) up to the O(h^2) error in space.
#include <cmath>

#include <tfel/tfel.hpp>
#include <tfel/tfel_export.hpp>

double initial_condition(const double* x) { return std::sin(M_PI * x[0]) * std::sin(M_PI * x[1]); }
double bc(const double*) { return 0.0; }

int main(int argc, char* argv[]) {
  try {
    typedef cell::triangle cell_type;
    typedef cell::triangle::fe::lagrange_p1 fe_type;
    typedef finite_element_space<fe_type> fes_type;
    typedef finite_element_space<fe_type>::element element_type;
    typedef quad::triangle::qf5pT q_type;
    using mesh_type = fe_mesh<cell_type>;

    const std::size_t n(argc > 1 ? std::stoul(argv[1]) : 100);
    const std::size_t n_step(argc > 2 ? std::stoul(argv[2]) : 100);
    const std::string out_dir(argc > 3 ? argv[3] : ".");
    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());

    fes_type fes(m, dm, bc);
    const element_type c_init(projector::l2<fe_type, q_type>(initial_condition, fes));
    exporter::ensight6(out_dir + "/initial_condition", to_mesh_vertex_data<fe_type>(c_init), "c_init");

    const std::size_t M(100);
    const double delta_t(2.0 / M);
    const double diffusivity(1.0);

    bilinear_form<fes_type, fes_type> a(fes, fes); {
      const auto u(a.get_trial_function());
      const auto v(a.get_test_function());
      a += integrate<q_type>(diffusivity * (d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v)), m);
      a += integrate<q_type>((1.0 / delta_t) * u * v, m);
    }

    element_type c(c_init);

    exporter::ensight6_transient<mesh_type> ens(out_dir + "/solution", m, "c");
    ens.export_time_step(0.0, to_mesh_vertex_data<fe_type>(c));

    dictionary p(dictionary()
                 .set("maxits",  2000u)
                 .set("restart", 1000u)
                 .set("rtol",    1.e-10)
                 .set("atol",    1.e-50)
                 .set("dtol",    1.e20)
                 .set("ilufill", 2u));
    solver::cuda::cg_jacobi s(p);

    for (std::size_t k(0); k < n_step; ++k) {
      linear_form<fes_type> f(fes); {
        const auto v(f.get_test_function());
        f += integrate<q_type>((1.0 / delta_t) * make_expr<fe_type>(c) * v
                               - initial_condition * v
                               + (1.0 - (k + 1) * delta_t) * 2 * M_PI * M_PI * (initial_condition * v), m);
      }
      c = a.solve(f, s);
      ens.export_time_step((k + 1) * delta_t, to_mesh_vertex_data<fe_type>(c));
    }

    // c(., t) = (1 - t) c_0 up to the spatial error
    const double t_end(n_step * delta_t);
    double err(0.0), scale(0.0);
    for (std::size_t i(0); i < fes.get_dof_number(); ++i) {
      err = std::max(err, std::fabs(c.get_coefficients().at(i) - (1.0 - t_end) * c_init.get_coefficients().at(i)));
      scale = std::max(scale, std::fabs(c_init.get_coefficients().at(i)));
    }
    std::cout << "laplacian_transient: n " << n << " steps " << n_step << " t " << t_end << " max |c - (1 - t) c_0| " << err
              << " (max |c_0| " << scale << ")" << std::endl;
    if (not (err <= 20.0 / (n * n) * scale)) throw std::string("transient solution is off the exact one");
    std::cout << "laplacian_transient: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "laplacian_transient: " << e << std::endl;
    return 1;
  }
  return 0;
}
