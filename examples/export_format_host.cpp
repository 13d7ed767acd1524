// Host-only check of the exporters' file layout (no device needed): the writers of <tfel/tfel_export.hpp> only ask a mesh
// for get_vertices / get_cells / get_vertex_number / get_cell_number, so a plain host mesh with the vertices and cells of
// gen_square_mesh (mesh.cpp:19-51) and nodal values computed here reproduces part of tests/golden/export/* -- the files
// the UNMODIFIED reference wrote -- byte for byte.
#include <tfel/tfel.hpp>
#include <tfel/tfel_export.hpp>

struct host_square_mesh {
  using cell_type = cell::triangle;
  host_square_mesh(std::size_t n) : vertices{(n + 1) * (n + 1), 2}, cells{2 * n * n, 3} {
    for (std::size_t j(0); j <= n; ++j)
      for (std::size_t i(0); i <= n; ++i) {
        vertices.at(j * (n + 1) + i, 0) = (1.0 / n) * i;
        vertices.at(j * (n + 1) + i, 1) = (1.0 / n) * j;
      }
    for (std::size_t j(0); j < n; ++j)
      for (std::size_t i(0); i < n; ++i) {
        const unsigned int v00(j * (n + 1) + i), v10(v00 + 1), v01(v00 + n + 1), v11(v01 + 1);
        const std::size_t k(2 * (n * j + i));
        cells.at(k, 0) = v00; cells.at(k, 1) = v10; cells.at(k, 2) = v11;
        cells.at(k + 1, 0) = v00; cells.at(k + 1, 1) = v01; cells.at(k + 1, 2) = v11;
      }
  }
  std::size_t get_vertex_number() const { return vertices.get_size(0); }
  std::size_t get_cell_number() const { return cells.get_size(0); }
  const array<double>& get_vertices() const { return vertices; }
  const array<unsigned int>& get_cells() const { return cells; }
  array<double> vertices;
  array<unsigned int> cells;
};

double export_g(const double* x) { return 1.0 + x[0] + 2.0 * x[1] + x[0] * x[1]; }
double export_h(const double* x) { return x[0] * x[0] - 0.5 * x[1]; }

int main(int argc, char* argv[]) {
  try {
    const std::string dir(argc > 1 ? argv[1] : "export");
    using data_type = mesh_data<double, host_square_mesh>;
    host_square_mesh m(3);
    data_type g(m, mesh_data_kind::vertex), h(m, mesh_data_kind::vertex), hc(m, mesh_data_kind::cell);
    for (std::size_t v(0); v < m.get_vertex_number(); ++v) {
      g.value(v, 0) = export_g(&m.vertices.at(v, 0));
      h.value(v, 0) = export_h(&m.vertices.at(v, 0));
    }
    const double phi[3] = {1.0 - 1.0 / 3.0 - 1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0};   // P1 basis at the barycentre (fe.hpp:260-317)
    for (std::size_t k(0); k < m.get_cell_number(); ++k) {
      double value(0.0);
      for (std::size_t n(0); n < 3; ++n) value += h.value(m.cells.at(k, n), 0) * phi[n];
      hc.value(k, 0) = value;
    }
    exporter::ensight6(dir + "/square", g, "g", hc, "hc", g, "g2");   // P2 vertex data == P1 vertex data for a nodal interpolant
    exporter::ensight6(dir + "/vector", data_type(m, mesh_data_kind::vertex, g, h), "gh");
    exporter::ascii<host_square_mesh>(dir + "/square_vertex.dat", g);
    exporter::ensight6_geometry(dir + "/geometry_only", m);
    {
      exporter::ensight6_transient<host_square_mesh> series(dir + "/series", m, "g");
      series.export_time_step(0.0, g);
      series.export_time_step(0.25, h);
    }
    std::cout << "export_format_host: ok" << std::endl;
  } catch (const std::string& e) {
    std::cerr << "export_format_host: " << e << std::endl;
    return 1;
  }
  return 0;
}
