// Callers on either side of the path (projector.hpp:12-97, fes.hpp:219-405): nodal interpolation, L2 projection, the
// vector-space algebra of discrete functions and their use as coefficient fields of the next form. The coefficients
// stay in device memory from a solve through the algebra into the next += and the next solve; the program checks that
// no host <-> device copy of coefficients happened on that chain, then writes every coefficient vector to
// projectors.bin (tests/test_cpp_surface.py compares them with the oracle).
#include <tfel/tfel.hpp>

#include <fstream>

double wave(const double* x) { return std::sin(3.0 * x[0]) * std::cos(2.0 * x[1]) + 0.25; }
double bump(const double* x) { return x[0] * (1.0 - x[0]) * x[1] + 0.5 * x[1] * x[1]; }

template <typename fe_type>
void run(const fe_mesh<cell::triangle>& m, std::ofstream& out) {
  using fes_type = finite_element_space<fe_type>;
  using quad_type = quad::triangle::qf5pT;
  fes_type fes(m);
  const std::size_t nd(fes.get_dof_number());
  auto put = [&](const typename fes_type::element& e) {
    out.write(reinterpret_cast<const char*>(e.get_coefficients().get_data()), sizeof(double) * nd);
  };

  const typename fes_type::element a(projector::lagrange<fe_type>(wave, fes));
  const typename fes_type::element b(projector::lagrange<fe_type>(bump, fes));
  const typename fes_type::element pa(projector::l2<fe_type, quad_type>(wave, fes));

  // device chain: solve -> algebra -> coefficient field of the next linear form -> solve
  const std::size_t before(tfel_cuda::coefficient_transfers());
  const typename fes_type::element pb(projector::l2<fe_type, quad_type>(xfun::x<0>() * (1.0 - xfun::x<0>()) * xfun::x<1>() + 0.5 * xfun::x<1>() * xfun::x<1>(), fes));
  const typename fes_type::element w(2.0 * pa - pb / 3.0 + pa);
  const typename fes_type::element pw(projector::l2<fe_type, quad_type>(make_expr<fe_type>(w) + d<1>(make_expr<fe_type>(pb)), fes));
  const std::size_t moved(tfel_cuda::coefficient_transfers() - before);
  std::cout << "fe " << fe_type::id << " dofs " << nd << " coefficient transfers on the device chain: " << moved << std::endl;
  if (moved != 0) throw std::string("coefficients left the device between solve, algebra and the next +=");

  const typename fes_type::element mix(0.5 * a + b * 3.0 - a / 4.0);
  put(a); put(b); put(pa); put(pb); put(w); put(pw); put(mix);
}

int main(int argc, char** argv) {
  try {
    const std::size_t n(argc > 1 ? std::stoul(argv[1]) : 16);
    fe_mesh<cell::triangle> m(gen_square_mesh(1.0, 1.0, n, n));
    std::ofstream out("projectors.bin", std::ios::binary);
    run<cell::triangle::fe::lagrange_p1>(m, out);
    run<cell::triangle::fe::lagrange_p2>(m, out);
    std::cout << "projectors: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "projectors: " << e << std::endl;
    return 1;
  }
  return 0;
}
