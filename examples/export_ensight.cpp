// The output side of the path on the CUDA headers: to_mesh_vertex_data / to_mesh_cell_data (fes.hpp:464-504) and the
// EnSight 6 / ASCII exporters (export.hpp). Writes, into the directory given as argument, the same set of files
// oracle/ref_driver.cpp --problem export writes with the UNMODIFIED reference (tests/golden/export/*): the GPU test
// compares them byte for byte.
#include <tfel/tfel.hpp>
#include <tfel/tfel_export.hpp>

double export_g(const double* x) { return 1.0 + x[0] + 2.0 * x[1] + x[0] * x[1]; }
double export_h(const double* x) { return x[0] * x[0] - 0.5 * x[1]; }
double export_t(const double* x) { return 1.0 + x[0] - x[1] + 2.0 * x[2]; }

int main(int argc, char* argv[]) {
  try {
    const std::string dir(argc > 1 ? argv[1] : "export");
    const std::size_t n(argc > 2 ? std::stoul(argv[2]) : 3);
    {
      using cell_type = cell::triangle;
      using mesh_type = fe_mesh<cell_type>;
      using p1 = cell_type::fe::lagrange_p1;
      using p2 = cell_type::fe::lagrange_p2;
      mesh_type m(gen_square_mesh(1.0, 1.0, n, n));
      finite_element_space<p1> fes1(m);
      finite_element_space<p2> fes2(m);
      const auto g1(projector::lagrange<p1>(export_g, fes1));
      const auto h1(projector::lagrange<p1>(export_h, fes1));
      const auto g2(projector::lagrange<p2>(export_g, fes2));
      {
        // element algebra (fes.hpp:314-405): (2 g - g) / 1 + (h - h) is g again, coefficient for coefficient
        const auto again((2.0 * g1 - g1) / 1.0 + (h1 - h1).transformed([](double c) { return 3.0 * c; }));
        for (std::size_t i(0); i < fes1.get_dof_number(); ++i)
          if (again.get_coefficients().at(i) != g1.get_coefficients().at(i)) throw std::string("element algebra is wrong");
      }
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
    std::cout << "export_ensight: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "export_ensight: " << e << std::endl;
    return 1;
  }
  return 0;
}
