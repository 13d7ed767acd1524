// test/stokes_2d_p2_p1.cpp:36-148 of the reference (with "atol") on the CUDA path:
// composite P2-P2-P1 space, two-sided driven cavity data, pressure pinned at one vertex, GMRES + ILU(0).
//   stokes_2d_p2_p1 n [solution.bin [rtol]]
#include <tfel/tfel.hpp>

#include <fstream>

double f0(const double*) { return 0.0; }
double f1(const double*) { return 0.0; }
double u0_bv(const double* x) { return x[1] < 0.0001 ? x[0] * (1.0 - x[0]) : 0.0; }
double u1_bv(const double* x) { return x[0] < 0.0001 ? x[1] * (1.0 - x[1]) : 0.0; }
double p_v(const double*) { return 0.0; }

int main(int argc, char* argv[]) {
  try {
    if (argc < 2)
      throw std::string("wrong number of arguments. An single integer describing the number of subdivisions is expected.");
    const std::size_t n(std::stoi(argv[1]));

    using cell_type = cell::triangle;
    using u_fe_type = cell_type::fe::lagrange_p2;
    using p_fe_type = cell_type::fe::lagrange_p1;
    using quad_type = quad::triangle::qf5pT;

    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    submesh<cell_type, cell::point> pinned_pressure_point(m.get_point_submesh(n * n / 2 + n / 2));

    using fe_type = composite_finite_element<u_fe_type, u_fe_type, p_fe_type>;
    using fes_type = composite_finite_element_space<fe_type>;
    fes_type fes(m);

    fes.add_dirichlet_boundary<0>(dm, u0_bv);
    fes.add_dirichlet_boundary<1>(dm, u1_bv);
    fes.add_dirichlet_boundary<2>(pinned_pressure_point, p_v);

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

    linear_form<fes_type> f(fes); {
      auto v0(f.get_test_function<0>());
      auto v1(f.get_test_function<1>());

      f += integrate<quad_type>(    f0 * v0
                                  + f1 * v1
                                , m);
    }

    dictionary param(dictionary()
                     .set("maxits",  2000u)
                     .set("restart", 1000u)
                     .set("rtol",    argc > 3 ? std::stod(argv[3]) : 1.e-8)
                     .set("atol",    1.e-50)
                     .set("dtol",    1.e20)
                     .set("ilufill", 2u));
    solver::cuda::gmres_ilu s(param);   // ILU(ilufill = 2), like solver::petsc::gmres_ilu (solver.hpp:162-163)
    dictionary report;
    const fes_type::element x(a.solve(f, s, &report));

    const auto p(x.get_component<2>());
    std::cout.precision(10);
    std::cout << "dofs " << fes.get_total_dof_number() << " its " << report.get<unsigned>("its")
              << " converged " << report.get<bool>("converged") << " p_dofs " << p.get_coefficients().get_size(0) << std::endl;
    if (argc > 2) {
      std::ofstream out(argv[2], std::ios::binary);
      out.write(reinterpret_cast<const char*>(x.get_coefficients().get_data()), sizeof(double) * x.get_coefficients().get_size(0));
    }
  } catch (const std::string& e) {
    std::cerr << "stokes_2d_p2_p1: " << e << std::endl;
    return 1;
  }
  return 0;
}
