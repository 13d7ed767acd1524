// Coefficient functions as __device__ functors (include/tfel/tfel_device.cuh): the general Poisson problem of the parity
// tests (-lap u + c(x) u = f(x), u = g on the boundary; source of examples/poisson.cpp:4-14 of the reference) assembled
// three times -- with host callbacks (tabulated on the host, the fallback that reports itself), with device functors
// tabulated on the device up front (device_expr) and with device functors evaluated when the form is assembled (dev(f):
// a kernel of THIS translation unit launched by the library on its stream, nothing is evaluated on the host) -- and
// compared.
#include <tfel/tfel.hpp>
#include <tfel/tfel_device.cuh>

#include <algorithm>

double src_host(const double* x) {
  const double sigma(0.1);
  const double r(std::sqrt(std::pow(x[0] - 0.5, 2) + std::pow(x[1] - 0.5, 2)));
  return std::exp(-r * r / (2.0 * sigma * sigma));
}
double reac_host(const double* x) { return 1.0 + x[0] * x[0] + 0.5 * x[1]; }
double g_bc(const double* x) { return 1.0 + x[0] + 2.0 * x[1] + x[0] * x[1]; }

struct src_dev {
  double sigma;
  __device__ double operator()(const double* x) const {
    const double r(sqrt(pow(x[0] - 0.5, 2.0) + pow(x[1] - 0.5, 2.0)));
    return exp(-r * r / (2.0 * sigma * sigma));
  }
};
struct reac_dev {
  __device__ double operator()(const double* x) const { return 1.0 + x[0] * x[0] + 0.5 * x[1]; }
};
__device__ double reac_plain(const double* x) { return 1.0 + x[0] * x[0] + 0.5 * x[1]; }
TFEL_DEVICE_FUNCTION(reac_plain);

int main(int argc, char* argv[]) {
  try {
    const std::size_t n = argc > 1 ? std::stoul(argv[1]) : 48;
    using cell_type = cell::triangle;
    using fe_type = cell_type::fe::lagrange_p2;
    using fes_type = finite_element_space<fe_type>;
    using quad_type = quad::triangle::qf5pT;

    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    fes_type fes(m);
    fes.add_dirichlet_boundary(dm, g_bc);

    dictionary param(dictionary().set("maxits", 20000u).set("restart", 1000u).set("rtol", 1.e-12)
                     .set("atol", 1.e-50).set("dtol", 1.e20).set("ilufill", 2u));
    std::vector<array<double>> sol;
    unsigned long long host_calls[3] = {0, 0, 0};
    for (int variant(0); variant < 3; ++variant) {
      const unsigned long long calls_before(tfel_cuda::host_callback_evaluations());
      bilinear_form<fes_type, fes_type> a(fes, fes);
      linear_form<fes_type> b(fes);
      auto u(a.get_trial_function()); auto v(a.get_test_function()); auto w(b.get_test_function());
      if (variant == 0) {
        a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + reac_host * (u * v), m);
        b += integrate<quad_type>(src_host * w, m);
      } else if (variant == 1) {
        a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + device_expr<quad_type>(reac_dev(), m) * (u * v), m);
        b += integrate<quad_type>(device_expr<quad_type>(src_dev{0.1}, m) * w, m);
      } else {
        a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v) + dev(reac_plain_functor()) * (u * v), m);
        b += integrate<quad_type>(dev(src_dev{0.1}) * w, m);
      }
      host_calls[variant] = tfel_cuda::host_callback_evaluations() - calls_before;
      solver::cuda::cg_jacobi s(param);
      sol.push_back(a.solve(b, s).get_coefficients());
    }
    std::cout << "device_functor: host evaluations of integrand leaves: callbacks " << host_calls[0] << ", device_expr "
              << host_calls[1] << ", dev " << host_calls[2] << std::endl;
    if (host_calls[0] == 0 or host_calls[1] != 0 or host_calls[2] != 0) throw std::string("unexpected host evaluation counts");
    double diff(0.0), nrm(0.0);
    for (std::size_t i(0); i < fes.get_dof_number(); ++i) {
      diff = std::max(diff, std::max(std::fabs(sol[0].at(i) - sol[1].at(i)), std::fabs(sol[0].at(i) - sol[2].at(i))));
      nrm = std::max(nrm, std::fabs(sol[0].at(i)));
    }
    std::cout << "device_functor: dofs " << fes.get_dof_number() << " max |u_host_callback - u_device_functor| " << diff
              << " (max |u| " << nrm << ")" << std::endl;
    if (not (diff <= 1e-10 * nrm)) throw std::string("host-callback and device-functor assemblies disagree");
    std::cout << "device_functor: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "device_functor: " << e << std::endl;
    return 1;
  }
  return 0;
}
