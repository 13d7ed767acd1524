// test/navier_stokes_2d_p2_p1.cpp:236-448 of the reference (`navier_stokes(n)`) on the CUDA path: P2-P2-P1 composite
// space, regularised lid-driven cavity data (:22-28), implicit Euler + Newton iteration with a NEW bilinear_form and
// solver per Newton step, stopping on the relative L2 norm of the velocity step (rank-0 integrate). Differences from the
// reference program: solver::cuda::gmres_ilu (GMRES(restart) + ILU(ilufill), same dictionary) instead of
// solver::petsc::gmres_ilu, "atol" (the reference's "abstol" makes its own solver constructor throw), the
// number of time steps is an argument, no EnSight export. The constructor of the per-Newton-step bilinear_form and the
// solver set-up are timed as well: after the first Newton step both find their structure (CSR pattern, slot map, SpMV plan,
// ILU symbolic phase) parked by the objects of the previous step.
//   navier_stokes_2d_p2_p1 n [time_steps [solution.bin]]
#include <tfel/tfel.hpp>

#include <chrono>
#include <fstream>

double f0(const double*) { return 0.0; }
double f1(const double*) { return 0.0; }
double u0_bv(const double* x) { return x[1] < 0.00001 ? 4.0 * x[0] * (1.0 - x[0]) : 0.0; }
double u1_bv(const double* x) { return x[0] < 0.00001 ? 4.0 * x[1] * (1.0 - x[1]) : 0.0; }
double p_v(const double*) { return 0.0; }

struct timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double tic() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

int main(int argc, char* argv[]) {
  try {
    if (argc < 2) throw std::string("usage: navier_stokes_2d_p2_p1 n [time_steps [solution.bin]]");
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

    array<double> u_comp{fes.get_dof_number<0>()};
    array<double> p_comp{fes.get_dof_number<2>()};
    u_comp.fill(0.0);
    p_comp.fill(0.0);
    finite_element_space<u_fe_type>::element
      u0(fes.get_finite_element_space<0>(), u_comp),
      u1(fes.get_finite_element_space<1>(), u_comp);
    finite_element_space<p_fe_type>::element
      p(fes.get_finite_element_space<2>(), p_comp);

    fes_type::element x(fes, u0, u1, p);
    fes_type::element xp(x);

    const double reynolds(5000.0);
    const double end_time(20.0);
    const double time_step(0.5);
    const std::size_t end_time_step(argc > 2 ? std::stoul(argv[2]) : std::size_t(std::ceil(end_time / time_step)));
    const std::size_t newton_max_step(10);
    const double newton_rtol(1.0e-8);

    double bilinear_form_elapsed_time(0.0), linear_form_elapsed_time(0.0), linear_solve_elapsed_time(0.0);
    double total_bilinear(0.0), total_linear(0.0), total_solve(0.0);
    double total_ctor(0.0), total_setup(0.0), total_symbolic(0.0), total_krylov(0.0);
    double first_ctor(0.0), first_setup(0.0), first_symbolic(0.0);
    std::size_t newton_steps(0), krylov_its(0), not_converged(0);
    const timer whole_run;

    double time(0.0);
    for (std::size_t k(0); k < end_time_step; ++k) {
      std::cout << "time step #" << k << "(time " << time << ")" << std::endl;
      time += time_step;

      bool newton_done(false);
      for (std::size_t j(0); j < newton_max_step and not newton_done; ++j) {
        auto v0_h(xp.template get_component<0>());
        auto v1_h(xp.template get_component<1>());

        auto v0(make_expr<u_fe_type>(v0_h));
        auto v1(make_expr<u_fe_type>(v1_h));

        const timer t_ctor;
        bilinear_form<fes_type, fes_type> a(fes, fes);
        tfelcu_ctx_synchronize(tfel_cuda::context());
        const double form_ctor_elapsed_time(t_ctor.tic()); {
          timer t;

          auto w0(a.get_test_function<0>());
          auto w1(a.get_test_function<1>());
          auto q (a.get_test_function<2>());

          auto vn0(a.get_trial_function<0>());
          auto vn1(a.get_trial_function<1>());
          auto pn (a.get_trial_function<2>());

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
          tfelcu_ctx_synchronize(tfel_cuda::context());
          bilinear_form_elapsed_time = t.tic();
        }

        linear_form<fes_type> f(fes); {
          timer t;

          auto w0(f.get_test_function<0>());
          auto w1(f.get_test_function<1>());

          auto u0_h(x.template get_component<0>());
          auto u1_h(x.template get_component<1>());

          auto u0(make_expr<u_fe_type>(u0_h));
          auto u1(make_expr<u_fe_type>(u1_h));

          f += integrate<quad_type>( u0 * w0 + u1 * w1 +

                                     time_step * (f0 * w0 +
                                                  f1 * w1) +

                                     time_step * (v0 * d<1>(v0) * w0 +
                                                  v0 * d<1>(v1) * w1 +
                                                  v1 * d<2>(v0) * w0 +
                                                  v1 * d<2>(v1) * w1)
                                     , m);
          tfelcu_ctx_synchronize(tfel_cuda::context());
          linear_form_elapsed_time = t.tic();
        }

        dictionary param(dictionary()
                         .set("maxits",  2000u)
                         .set("restart", 1000u)
                         .set("rtol",    1.e-8)
                         .set("atol",    1.e-50)
                         .set("dtol",    1.e20)
                         .set("ilufill", 2u));
        solver::cuda::gmres_ilu s(param);

        timer t;
        dictionary report;
        auto xpp(a.solve(f, s, &report));
        linear_solve_elapsed_time = t.tic();
        krylov_its += report.get<unsigned>("its");
        if (not report.get<bool>("converged")) { ++not_converged; std::cout << "warning: linear solve did not converge" << std::endl; }
        const double setup_ms(1e3 * report.get<double>("t_setup_s")), symbolic_ms(1e3 * report.get<double>("t_symbolic_s")),
                     krylov_ms(1e3 * report.get<double>("t_solve_s"));

        auto vn0_h(xpp.template get_component<0>());
        auto vn1_h(xpp.template get_component<1>());

        const double
          newton_step_0_norm(std::sqrt(integrate<quad_type>((make_expr<u_fe_type>(vn0_h) - make_expr<u_fe_type>(v0_h)) *
                                                            (make_expr<u_fe_type>(vn0_h) - make_expr<u_fe_type>(v0_h))
                                                            , m))),
          newton_step_1_norm(std::sqrt(integrate<quad_type>((make_expr<u_fe_type>(vn1_h) - make_expr<u_fe_type>(v1_h)) *
                                                            (make_expr<u_fe_type>(vn1_h) - make_expr<u_fe_type>(v1_h))
                                                            , m)));
        const double
          velocity_0_norm(std::sqrt(integrate<quad_type>(make_expr<u_fe_type>(vn0_h) * make_expr<u_fe_type>(vn0_h), m))),
          velocity_1_norm(std::sqrt(integrate<quad_type>(make_expr<u_fe_type>(vn1_h) * make_expr<u_fe_type>(vn1_h), m)));

        const double
          newton_relative_step_0_norm(newton_step_0_norm / velocity_0_norm),
          newton_relative_step_1_norm(newton_step_1_norm / velocity_1_norm);
        std::cout << "newton step #" << j << ": relative step norm (" << newton_relative_step_0_norm << ", "
                  << newton_relative_step_1_norm << "), elapsed ms (bilinear, linear, solve): ("
                  << bilinear_form_elapsed_time << ", " << linear_form_elapsed_time << ", " << linear_solve_elapsed_time
                  << "), krylov its " << report.get<unsigned>("its") << "; form constructor " << form_ctor_elapsed_time
                  << " ms, solver set-up " << setup_ms << " ms (ILU symbolic " << symbolic_ms << " ms), GMRES " << krylov_ms
                  << " ms" << std::endl;

        newton_done = (newton_relative_step_0_norm < newton_rtol and
                       newton_relative_step_1_norm < newton_rtol);

        xp = xpp;
        total_bilinear += bilinear_form_elapsed_time; total_linear += linear_form_elapsed_time; total_solve += linear_solve_elapsed_time;
        if (newton_steps == 0) { first_ctor = form_ctor_elapsed_time; first_setup = setup_ms; first_symbolic = symbolic_ms; }
        else { total_ctor += form_ctor_elapsed_time; total_setup += setup_ms; total_symbolic += symbolic_ms; }
        total_krylov += krylov_ms;
        ++newton_steps;
      }

      x = xp;
    }
    std::cout << "summary: n " << n << " cells " << m.get_cell_number() << " dofs " << fes.get_total_dof_number()
              << " time_steps " << end_time_step << " newton_steps " << newton_steps << " krylov_its " << krylov_its
              << " ms_per_newton_step (bilinear, linear, solve) (" << total_bilinear / newton_steps << ", "
              << total_linear / newton_steps << ", " << total_solve / newton_steps << ")" << std::endl;
    {
      const double run_ms(whole_run.tic());
      const double later(newton_steps > 1 ? double(newton_steps - 1) : 1.0);
      std::cout.precision(6);
      std::cout << "json: {\"n\": " << n << ", \"cells\": " << m.get_cell_number() << ", \"dofs\": " << fes.get_total_dof_number()
                << ", \"time_steps\": " << end_time_step << ", \"newton_steps\": " << newton_steps << ", \"krylov_its\": " << krylov_its
                << ", \"linear_solves_not_converged\": " << not_converged
                << ", \"ms_per_time_step\": " << run_ms / end_time_step
                << ", \"cells_per_s_whole_run\": " << double(m.get_cell_number()) * newton_steps / (run_ms * 1e-3)
                << ", \"per_newton_step_ms\": {\"bilinear_assembly\": " << total_bilinear / newton_steps
                << ", \"linear_assembly\": " << total_linear / newton_steps << ", \"solve_call\": " << total_solve / newton_steps
                << ", \"gmres\": " << total_krylov / newton_steps << "}"
                << ", \"structure_ms\": {\"first_step\": {\"form_constructor\": " << first_ctor << ", \"solver_setup\": " << first_setup
                << ", \"ilu_symbolic\": " << first_symbolic << "}, \"later_steps_mean\": {\"form_constructor\": " << total_ctor / later
                << ", \"solver_setup\": " << total_setup / later << ", \"ilu_symbolic\": " << total_symbolic / later << "}}}" << std::endl;
    }
    if (argc > 3) {
      std::ofstream out(argv[3], std::ios::binary);
      out.write(reinterpret_cast<const char*>(xp.get_coefficients().get_data()), sizeof(double) * xp.get_coefficients().get_size(0));
    }
  } catch (const std::string& e) {
    std::cerr << "navier_stokes_2d_p2_p1: " << e << std::endl;
    return 1;
  }
  return 0;
}
