// Checks of the host surface that need no reference program: a USER-WRITTEN solver plugin deriving from
// solver::basic_solver (solver.hpp:29-40) receives the reference's CRS (solver.cpp:71-96) and solves on the host;
// coefficient fields, device-evaluated functions of x, rank-0 integrals (test/fe_derivative_form.cpp:4-26,
// test/quadrature_2d.cpp:10-18) and the error convention (std::string thrown by value).
#include <tfel/tfel.hpp>

#include <algorithm>
#include <cassert>

// Jacobi-preconditioned CG written against the plugin interface only
class host_cg : public solver::basic_solver {
public:
  void set_operator(const sparse_matrix& m) { a = m; }
  bool solve(const array<double>& rhs, array<double>& x, dictionary& report) {
    const std::size_t n(a.get_row_number());
    std::vector<double> r(n), z(n), p(n), q(n), dinv(n);
    for (std::size_t i(0); i < n; ++i) {
      double dd(1.0);
      bool identity(a.row[i + 1] - a.row[i] == 1);
      for (int e(a.row[i]); e < a.row[i + 1]; ++e) if (a.col[e] == int(i)) dd = a.val[e];
      dinv[i] = 1.0 / dd;
      x.at(i) = identity ? rhs.at(i) : 0.0;
    }
    auto mult = [&](const double* in, std::vector<double>& out) {
      for (std::size_t i(0); i < n; ++i) { double s(0.0); for (int e(a.row[i]); e < a.row[i + 1]; ++e) s += a.val[e] * in[a.col[e]]; out[i] = s; }
    };
    mult(x.get_data(), q);
    double bb(0.0), rz(0.0), rr(0.0);
    for (std::size_t i(0); i < n; ++i) { r[i] = rhs.at(i) - q[i]; bb += rhs.at(i) * rhs.at(i); z[i] = dinv[i] * r[i]; p[i] = z[i]; rz += r[i] * z[i]; rr += r[i] * r[i]; }
    unsigned its(0);
    while (rr > 1e-20 * bb and its < 10000) {
      mult(p.data(), q);
      double pq(0.0); for (std::size_t i(0); i < n; ++i) pq += p[i] * q[i];
      const double alpha(rz / pq);
      double rz_new(0.0); rr = 0.0;
      for (std::size_t i(0); i < n; ++i) { x.at(i) += alpha * p[i]; r[i] -= alpha * q[i]; z[i] = dinv[i] * r[i]; rz_new += r[i] * z[i]; rr += r[i] * r[i]; }
      const double beta(rz_new / rz); rz = rz_new;
      for (std::size_t i(0); i < n; ++i) p[i] = z[i] + beta * p[i];
      ++its;
    }
    report.set("its", its);
    return true;
  }
private:
  sparse_matrix a;
};

double one(const double*) { return 1.0; }
double sin_sum(const double* x) { return std::sin(x[0]) + std::sin(x[1]); }
double half_y2(const double* x) { return 0.5 * x[1] * x[1]; }

int main() {
  try {
    using cell_type = cell::triangle;
    using fe_type = cell_type::fe::lagrange_p1;
    using fes_type = finite_element_space<fe_type>;
    using quad_type = quad::triangle::qf5pT;
    const std::size_t n(24);
    fe_mesh<cell_type> m(gen_square_mesh(1.0, 1.0, n, n));
    submesh<cell_type> dm(m.get_boundary_submesh());
    {
      // face neighbours (mesh.hpp:301-332): symmetric, and -1 exactly on the 4n boundary edges
      std::size_t on_boundary(0);
      for (std::size_t k(0); k < m.get_cell_number(); ++k)
        for (std::size_t j(0); j < 3; ++j) {
          const int other(static_cast<int>(m.get_cell_neighbour(k, j)));
          if (other < 0) { ++on_boundary; continue; }
          bool back(false);
          for (std::size_t jj(0); jj < 3; ++jj) back = back or static_cast<std::size_t>(m.get_cell_neighbour(other, jj)) == k;
          if (not back) throw std::string("cell neighbour table is not symmetric");
        }
      if (on_boundary != dm.get_cell_number() or on_boundary != 4 * n) throw std::string("cell neighbour table: wrong boundary count");
    }
    fes_type fes(m);
    fes.add_dirichlet_boundary(dm);
    bilinear_form<fes_type, fes_type> a(fes, fes); {
      auto u(a.get_trial_function()); auto v(a.get_test_function());
      a += integrate<quad_type>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v), m);
    }
    linear_form<fes_type> b(fes); { auto v(b.get_test_function()); b += integrate<quad_type>(one * v, m); }
    dictionary param(dictionary().set("maxits", 2000u).set("restart", 1000u).set("rtol", 1.e-12)
                     .set("atol", 1.e-50).set("dtol", 1.e20).set("ilufill", 2u));
    solver::cuda::cg_jacobi s_dev(param);
    host_cg s_host;
    dictionary r1, r2;
    const fes_type::element u_dev(a.solve(b, s_dev, &r1));
    const fes_type::element u_host(a.solve(b, s_host, &r2));
    double diff(0.0), nrm(0.0);
    for (std::size_t i(0); i < fes.get_dof_number(); ++i) {
      const double e(u_dev.get_coefficients().at(i) - u_host.get_coefficients().at(i));
      diff += e * e; nrm += u_host.get_coefficients().at(i) * u_host.get_coefficients().at(i);
    }
    std::cout << "plugin: device its " << r1.get<unsigned>("its") << " host its " << r2.get<unsigned>("its")
              << " rel diff " << std::sqrt(diff / nrm) << std::endl;
    if (not (std::sqrt(diff / nrm) < 1e-9)) throw std::string("user plugin and device solver disagree");

    // rank-0 integrals: host callback, device-evaluated function, P2 coefficient field and its derivative
    fe_mesh<cell_type> mpi(gen_square_mesh(M_PI, M_PI, 40, 40));
    const double i1(integrate<quad_type>(make_expr(sin_sum), mpi));
    const double i2(integrate<quad_type>(static_cast<expression<false>>(xfun::sin(xfun::x<0>()) + xfun::sin(xfun::x<1>())), mpi));
    std::cout.precision(14);
    std::cout << "int sin x + sin y: " << i1 << " " << i2 << " (4 pi = " << 4 * M_PI << ")" << std::endl;
    if (std::fabs(i1 - 4 * M_PI) > 1e-9 or std::fabs(i2 - 4 * M_PI) > 1e-9) throw std::string("quadrature_2d known answer missed");

    using p2_type = cell_type::fe::lagrange_p2;
    using p2_fes_type = finite_element_space<p2_type>;
    fe_mesh<cell_type> m4(gen_square_mesh(1.0, 1.0, 4, 4));
    p2_fes_type fes2(m4);
    // nodal interpolation of 0.5 y^2 through the Dirichlet machinery: every dof of the boundary ... use dof coordinates
    std::size_t nd(fes2.get_dof_number());
    std::vector<unsigned int> dofs(nd); for (std::size_t i(0); i < nd; ++i) dofs[i] = i;
    std::vector<double> xy(2 * nd);
    tfel_cuda::ck(tfelcu_fes_dof_coordinates(fes2.handle(), dofs.data(), nd, xy.data()));
    array<double> coef{nd};
    for (std::size_t i(0); i < nd; ++i) coef.at(i) = half_y2(&xy[2 * i]);
    const p2_fes_type::element w(fes2, coef);
    const double j1(integrate<quad_type>(make_expr<p2_type>(w), m4));
    const double j2(integrate<quad_type>(d<2>(make_expr<p2_type>(w)), m4));
    std::cout << "int 0.5 y^2 = " << j1 << " (1/6), int d_y = " << j2 << " (1/2)" << std::endl;
    if (std::fabs(j1 - 1.0 / 6.0) > 1e-13 or std::fabs(j2 - 0.5) > 1e-13) throw std::string("fe_derivative_form known answer missed");

    // projectors (projector.hpp:12-97, test/l2_projection.cpp): 0.5 y^2 is in P2, so both projections reproduce it
    {
      const p2_fes_type::element wl(projector::lagrange<p2_type>(half_y2, fes2));
      const p2_fes_type::element w2(projector::l2<p2_type, quad_type>(half_y2, fes2));
      double e1(0.0), e2(0.0);
      for (std::size_t i(0); i < nd; ++i) {
        e1 = std::max(e1, std::fabs(wl.get_coefficients().at(i) - coef.at(i)));
        e2 = std::max(e2, std::fabs(w2.get_coefficients().at(i) - coef.at(i)));
      }
      std::cout << "projector::lagrange max diff " << e1 << ", projector::l2 max diff " << e2 << std::endl;
      if (e1 > 0.0 or e2 > 1e-6) throw std::string("projectors do not reproduce a P2 function");
    }

    // error convention: missing dictionary key -> std::string (solver.hpp:132)
    bool thrown(false);
    try { solver::cuda::cg_jacobi bad(dictionary().set("maxits", 10u).set("abstol", 1e-50)); }
    catch (const std::string& e) { thrown = true; std::cout << "expected error: " << e << std::endl; }
    if (not thrown) throw std::string("missing keys were accepted");
    std::cout << "plugin_and_forms: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "plugin_and_forms: " << e << std::endl;
    return 1;
  }
  return 0;
}
