// The reference's README program (README.md:66-120), unchanged except for the solver plugin:
// solver::cuda::cg_jacobi instead of solver::petsc::gmres_ilu, and the export replaced by a binary dump.
//   poisson [n] [solution.bin]
#include <tfel/tfel.hpp>

#include <algorithm>
#include <fstream>

double f(const double* x);

int main(int argc, char* argv[]) {
  try {
    const std::size_t n = argc > 1 ? std::stoul(argv[1]) : 10;
    using cell_type = cell::triangle;
    using mesh_type = fe_mesh<cell_type>;
    using fe_type = cell_type::fe::lagrange_p1;
    using fes_type = finite_element_space<fe_type>;
    using quad_type = quad::triangle::qf5pT;

    mesh_type m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());

    fes_type fes(m);
    fes.add_dirichlet_boundary(dm);

    bilinear_form<fes_type, fes_type> a(fes, fes); {
      auto u(a.get_trial_function());
      auto v(a.get_test_function());

      a += integrate<quad_type>(d<1>(u) * d<1>(v) +
                                d<2>(u) * d<2>(v)
                                , m);
    }

    linear_form<fes_type> b(fes); {
      auto v(b.get_test_function());

      b += integrate<quad_type>(f * v, m);
    }

    dictionary param(dictionary()
                     .set("maxits",  20000u)
                     .set("restart", 1000u)
                     .set("rtol",    1.e-8)
                     .set("atol",    1.e-50)
                     .set("dtol",    1.e20)
                     .set("ilufill", 2u));
    solver::cuda::cg_jacobi s(param);

    dictionary report;
    const fes_type::element u(a.solve(b, s, &report));

    const array<double>& c(u.get_coefficients());
    double mx(0.0);
    for (std::size_t i(0); i < c.get_size(0); ++i) mx = std::max(mx, c.at(i));
    std::cout.precision(10);
    std::cout << "dofs " << fes.get_dof_number() << " dirichlet " << fes.get_dirichlet_dof_number()
              << " its " << report.get<unsigned>("its") << " max_u " << mx << std::endl;
    if (argc > 2) {
      std::ofstream out(argv[2], std::ios::binary);
      out.write(reinterpret_cast<const char*>(c.get_data()), sizeof(double) * c.get_size(0));
    }
  } catch (const std::string& e) {
    std::cerr << "poisson: " << e << std::endl;
    return 1;
  }
  return 0;
}

double f(const double* x) {
  return 1.0;
}
