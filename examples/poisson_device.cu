// examples/poisson.cpp of the reference (its source term f is the Gaussian of examples/poisson.cpp:4-14) with the source
// as DEVICE code: the same program, compiled by nvcc, `f` a __device__ functor instead of a host function pointer. A
// kernel of this translation unit evaluates it at the quadrature points of every cell when b += ... runs
// (tfel_device.cuh: dev(f)); nothing of the integrand is evaluated on the host, at any mesh size.
//   poisson_device [n] [linear_form.bin] [solution.bin]
#include <tfel/tfel.hpp>
#include <tfel/tfel_device.cuh>

#include <algorithm>
#include <chrono>
#include <fstream>

struct gaussian_source {   // examples/poisson.cpp:4-14
  double sigma;
  __device__ double operator()(const double* x) const {
    const double r(sqrt(pow(x[0] - 0.5, 2.0) + pow(x[1] - 0.5, 2.0)));
    return exp(-r * r / (2.0 * sigma * sigma));
  }
};

int main(int argc, char* argv[]) {
  try {
    const std::size_t n = argc > 1 ? std::stoul(argv[1]) : 100;
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

    linear_form<fes_type> b(fes);
    double assemble_ms(0.0);
    for (int pass(0); pass < 2; ++pass) {   // second pass timed (the first one loads the kernels)
      b.clear();
      tfelcu_ctx_synchronize(tfel_cuda::context());
      const auto t0(std::chrono::steady_clock::now());
      auto v(b.get_test_function());
      b += integrate<quad_type>(dev(gaussian_source{0.1}) * v, m);
      tfelcu_ctx_synchronize(tfel_cuda::context());
      assemble_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
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
    std::cout << "poisson_device: n " << n << " cells " << m.get_cell_number() << " dofs " << fes.get_dof_number()
              << " linear form with the device source " << assemble_ms << " ms, host evaluations of integrand leaves "
              << tfel_cuda::host_callback_evaluations() << ", its " << report.get<unsigned>("its") << " max_u " << mx << std::endl;
    if (tfel_cuda::host_callback_evaluations() != 0) throw std::string("an integrand leaf was evaluated on the host");
    if (argc > 2) {
      const array<double> bc(b.get_coefficients());
      std::ofstream out(argv[2], std::ios::binary);
      out.write(reinterpret_cast<const char*>(bc.get_data()), sizeof(double) * bc.get_size(0));
    }
    if (argc > 3) {
      std::ofstream out(argv[3], std::ios::binary);
      out.write(reinterpret_cast<const char*>(c.get_data()), sizeof(double) * c.get_size(0));
    }
  } catch (const std::string& e) {
    std::cerr << "poisson_device: " << e << std::endl;
    return 1;
  }
  return 0;
}
