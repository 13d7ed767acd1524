// <tfel/tfel_export.hpp> -- the output side of the path: mesh_data, to_mesh_vertex_data / to_mesh_cell_data and the
// EnSight 6 / ASCII exporters with the reference's spellings and file layout (src/core/mesh_data.hpp:215-330,
// fes.hpp:464-504, export.hpp:13-395), so that programs ending in
//     exporter::ensight6("poisson", to_mesh_vertex_data<fe_type>(u), "u");
// keep working on the CUDA path. The per-vertex / per-cell values are gathered on the device
// (tfelcu_fes_vertex_values / tfelcu_fes_cell_values); the files are written by a buffered formatter on the host.
// Byte-identical to the reference's files for the same data (tests/golden/export_*: written by the unmodified reference).
#ifndef TFEL_TFEL_EXPORT_HPP
#define TFEL_TFEL_EXPORT_HPP

#include <cstdio>
#include <sstream>
#include <utility>

#include "tfel.hpp"

enum class mesh_data_kind { cell, vertex };   // mesh_data.hpp:215

// values attached to the cells or to the vertices of a mesh, n_component per entity (mesh_data.hpp:241-330; the
// expression algebra of that header is out of scope)
template <typename value_t, typename mesh_t>
class mesh_data {
public:
  using mesh_type = mesh_t;
  using cell_type = typename mesh_t::cell_type;
  using value_type = value_t;

  mesh_data(const mesh_t& m, mesh_data_kind t, std::size_t n_component = 1ul)
    : msh(&m), type(t), values{entities(m, t), n_component} {}
  mesh_data(const mesh_t& m, mesh_data_kind t, array<value_t> v) : msh(&m), type(t), values(std::move(v)) {}
  // several one-component data side by side (mesh_data.hpp:269-273), e.g. the components of a velocity
  template <typename... more_t>
  mesh_data(const mesh_t& m, mesh_data_kind t, const mesh_data& first, const more_t&... more)
    : msh(&m), type(t), values{entities(m, t), 1 + sizeof...(more_t)} {
    const mesh_data* parts[] = {&first, &more...};
    for (std::size_t c(0); c < 1 + sizeof...(more_t); ++c) {
      if (parts[c]->get_component_number() != 1 or parts[c]->values.get_size(0) != values.get_size(0))
        throw std::string("mesh_data: components must be one-component data of the same kind");
      for (std::size_t k(0); k < values.get_size(0); ++k) values.at(k, c) = parts[c]->values.at(k, 0);
    }
  }

  const value_t& value(std::size_t k, std::size_t n) const { return values.at(k, n); }
  value_t& value(std::size_t k, std::size_t n) { return values.at(k, n); }
  const array<value_t>& get_values() const { return values; }
  std::size_t get_component_number() const { return values.get_size(1); }
  mesh_data_kind get_kind() const { return type; }
  const mesh_t& get_mesh() const { return *msh; }
  value_t evaluate(std::size_t k, const double*, std::size_t component) const {   // mesh_data.hpp:299-311
    if (type != mesh_data_kind::cell) throw std::string("mesh_data::evaluate: evaluation of vertex data is not implemented.");
    return values.at(k, component);
  }

private:
  static std::size_t entities(const mesh_t& m, mesh_data_kind t) {
    return t == mesh_data_kind::cell ? m.get_cell_number() : m.get_vertex_number();
  }
  const mesh_t* msh;
  mesh_data_kind type;
  array<value_t> values;
};

// fes.hpp:483-504: the coefficient of the vertex dofs, per vertex (elements whose first dim + 1 dofs sit on the vertices)
template <typename fe_type>
mesh_data<double, fe_mesh<typename fe_type::cell_type>>
to_mesh_vertex_data(const typename finite_element_space<fe_type>::element& v) {
  const finite_element_space<fe_type>& fes(v.get_finite_element_space());
  const fe_mesh<typename fe_type::cell_type>& m(fes.get_mesh());
  array<double> values{m.get_vertex_number(), 1};
  tfel_cuda::ck(tfelcu_fes_vertex_values(fes.handle(), v.get_coefficients().get_data(), values.get_data()));
  return mesh_data<double, fe_mesh<typename fe_type::cell_type>>(m, mesh_data_kind::vertex, std::move(values));
}

// fes.hpp:464-481: the discrete function at the barycentre of every cell
template <typename fe_type>
mesh_data<double, fe_mesh<typename fe_type::cell_type>>
to_mesh_cell_data(const typename finite_element_space<fe_type>::element& v) {
  const finite_element_space<fe_type>& fes(v.get_finite_element_space());
  const fe_mesh<typename fe_type::cell_type>& m(fes.get_mesh());
  array<double> values{m.get_cell_number(), 1};
  tfel_cuda::ck(tfelcu_fes_cell_values(fes.handle(), v.get_coefficients().get_data(), values.get_data()));
  return mesh_data<double, fe_mesh<typename fe_type::cell_type>>(m, mesh_data_kind::cell, std::move(values));
}

namespace exporter {
namespace detail {

// buffered text file: formatted numbers are appended to a 1 MB buffer and flushed with fwrite (the field layouts are the
// reference's iostream manipulators: setw(8) right integers, setw(12) setprecision(5) scientific reals)
class text_file {
public:
  explicit text_file(const std::string& path) : f(std::fopen(path.c_str(), "w")) {
    if (not f) throw std::string("failed to open ") + path + " for output";
    buf.reserve(1u << 20);
  }
  text_file(const text_file&) = delete;
  ~text_file() { flush(); std::fclose(f); }
  void text(const std::string& s) { buf += s; room(); }
  void integer8(unsigned long long v) { char t[32]; buf.append(t, std::snprintf(t, sizeof t, "%8llu", v)); room(); }
  void real12(double v) { char t[40]; buf.append(t, std::snprintf(t, sizeof t, "%12.5e", v)); room(); }
  void general12(double v) { char t[40]; buf.append(t, std::snprintf(t, sizeof t, "%.12g", v)); room(); }
private:
  void room() { if (buf.size() > (1u << 20) - 64) flush(); }
  void flush() { if (not buf.empty()) std::fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
  std::FILE* f;
  std::string buf;
};

template <typename cell_t> struct ensight_cell_name;                                  // export.hpp:87-95
template <> struct ensight_cell_name<cell::triangle> { static const char* value() { return "tria3"; } };
template <> struct ensight_cell_name<cell::tetrahedron> { static const char* value() { return "tetra4"; } };

inline const char* variable_kind(mesh_data_kind t, std::size_t n_component) {          // export.hpp:55-72
  if (n_component == 1) return t == mesh_data_kind::cell ? "scalar per element" : "scalar per node";
  if (n_component == 2 or n_component == 3) return t == mesh_data_kind::cell ? "vector per element" : "vector per node";
  throw std::string("unsupported component number");
}

// <name>.geom: node coordinates padded to three columns, then one part with 1-based connectivity (export.hpp:97-131)
template <typename mesh_t>
void write_geometry(const std::string& path, const mesh_t& m) {
  using cell_type = typename mesh_t::cell_type;
  const array<double> vertices(m.get_vertices());
  const array<unsigned int> cells(m.get_cells());
  text_file out(path);
  out.text("mesh\n\nnode id off\nelement id off\ncoordinates\n");
  out.integer8(vertices.get_size(0)); out.text("\n");
  const std::size_t dim(vertices.get_size(0) ? vertices.get_size(1) : 0);
  for (std::size_t v(0); v < vertices.get_size(0); ++v) {
    for (std::size_t c(0); c < 3; ++c) out.real12(c < dim ? vertices.at(v, c) : 0.0);
    out.text("\n");
  }
  out.text("part"); out.integer8(1); out.text("\npart1\n");
  out.text(ensight_cell_name<cell_type>::value()); out.text("\n");
  out.integer8(cells.get_size(0)); out.text("\n");
  for (std::size_t k(0); k < cells.get_size(0); ++k) {
    for (std::size_t n(0); n < static_cast<std::size_t>(cell_type::dim + 1); ++n) out.integer8(cells.at(k, n) + 1ull);
    out.text("\n");
  }
}

// <name>.var.<variable>: six reals per line, two- and three-component data padded to three (export.hpp:140-195)
template <typename mesh_t>
void write_variable(const std::string& path, const mesh_data<double, mesh_t>& data, const std::string& var_name) {
  const std::size_t nc(data.get_component_number());
  if (nc < 1 or nc > 3) throw std::string("write_static_variable_file: unsupported component number");
  const std::size_t width(nc == 1 ? 1 : 3);
  text_file out(path);
  out.text(var_name);
  if (data.get_kind() == mesh_data_kind::cell) {
    out.text("\npart"); out.integer8(1); out.text("\n");
    out.text(ensight_cell_name<typename mesh_t::cell_type>::value());
  }
  const std::size_t n(data.get_values().get_rank() ? data.get_values().get_size(0) : 0);
  for (std::size_t k(0); k < n; ++k)
    for (std::size_t c(0); c < width; ++c) {
      if ((k * width + c) % 6 == 0) out.text("\n");
      out.real12(c < nc ? data.value(k, c) : 0.0);
    }
}

inline void collect(const std::string&, std::vector<std::string>&, std::vector<std::string>&, std::vector<std::string>&) {}

template <typename mesh_t, typename... more_t>
void collect(const std::string& base, std::vector<std::string>& files, std::vector<std::string>& names,
             std::vector<std::string>& kinds, const mesh_data<double, mesh_t>& data, const std::string& var_name,
             more_t&&... more) {
  const std::string path(base + ".var." + var_name);
  write_variable(path, data, var_name);
  files.push_back(path); names.push_back(var_name);
  kinds.push_back(variable_kind(data.get_kind(), data.get_component_number()));
  collect(base, files, names, kinds, std::forward<more_t>(more)...);
}

template <typename mesh_t, typename... more_t>
const mesh_t& first_mesh(const mesh_data<double, mesh_t>& data, more_t&&...) { return data.get_mesh(); }

}  // namespace detail

// export.hpp:13-52: one line per vertex (coordinates, components) or per cell (barycentre, components), 12 significant digits
template <typename mesh_t>
void ascii(const std::string& filename, const mesh_data<double, mesh_t>& data) {
  const mesh_t& m(data.get_mesh());
  const array<double> vertices(m.get_vertices());
  const std::size_t dim(mesh_t::cell_type::dim);
  detail::text_file out(filename);
  if (data.get_kind() == mesh_data_kind::vertex) {
    for (std::size_t v(0); v < m.get_vertex_number(); ++v) {
      for (std::size_t c(0); c < dim; ++c) { out.general12(vertices.at(v, c)); out.text(" "); }
      for (std::size_t i(0); i < data.get_component_number(); ++i) { out.general12(data.value(v, i)); out.text(" "); }
      out.text("\n");
    }
  } else {
    const array<unsigned int> cells(m.get_cells());
    const double weight(1.0 / (dim + 1));
    for (std::size_t k(0); k < m.get_cell_number(); ++k) {
      for (std::size_t c(0); c < dim; ++c) {   // cell.hpp:300-312
        double bc(0.0);
        for (std::size_t n(0); n <= dim; ++n) bc += vertices.at(cells.at(k, n), c) * weight;
        out.general12(bc); out.text(" ");
      }
      for (std::size_t i(0); i < data.get_component_number(); ++i) { out.general12(data.value(k, i)); out.text(" "); }
      out.text("\n");
    }
  }
}

// export.hpp:246-273: <filename>.case, <filename>.geom and one <filename>.var.<name> per (data, name) pair
template <typename... pairs_t>
void ensight6(const std::string& filename, pairs_t&&... data_and_names) {
  static_assert(sizeof...(pairs_t) % 2 == 0 and sizeof...(pairs_t) >= 2, "ensight6(filename, data, \"name\", ...)");
  const std::string geometry(filename + ".geom");
  std::vector<std::string> files, names, kinds;
  detail::collect(filename, files, names, kinds, std::forward<pairs_t>(data_and_names)...);
  {
    detail::text_file out(filename + ".case");
    out.text("FORMAT\ntype: ensight\n\nGEOMETRY\nmodel: " + geometry + "\n\nVARIABLE\n");
    for (std::size_t i(0); i < names.size(); ++i) out.text(kinds[i] + ": " + names[i] + " " + files[i] + "\n");
  }
  detail::write_geometry(geometry, detail::first_mesh(data_and_names...));
}

// export.hpp:275-297
template <typename mesh_t>
void ensight6_geometry(const std::string& filename, const mesh_t& m) {
  const std::string geometry(filename + ".geom");
  {
    detail::text_file out(filename + ".case");
    out.text("FORMAT\ntype: ensight\n\nGEOMETRY\nmodel: " + geometry + "\n\nVARIABLE\n");
  }
  detail::write_geometry(geometry, m);
}

// export.hpp:299-392: one variable over time; the case file is rewritten after every step and by the destructor
template <typename mesh_t>
class ensight6_transient {
public:
  using cell_type = typename mesh_t::cell_type;
  ensight6_transient(const std::string& filename, const mesh_t& m, const std::string& variable_name)
    : filename(filename), variable_name(variable_name), kind(mesh_data_kind::cell), n_component(0) {
    detail::write_geometry(filename + ".geom", m);
  }
  ~ensight6_transient() { try { write_case(); } catch (...) {} }
  void export_time_step(double time, const mesh_data<double, mesh_t>& data) {
    char step[16];
    std::snprintf(step, sizeof step, "%06zu", times.size());
    detail::write_variable(filename + ".var." + variable_name + "." + step, data, variable_name);
    times.push_back(time);
    kind = data.get_kind(); n_component = data.get_component_number();
    write_case();
  }
private:
  void write_case() {
    detail::text_file out(filename + ".case");
    out.text("FORMAT\ntype: ensight\n\nGEOMETRY\nmodel: 1 " + filename + ".geom\n\nVARIABLE\n");
    out.text(std::string(detail::variable_kind(kind, n_component ? n_component : 1)) + ": 1 " + variable_name + " " + filename +
             ".var." + variable_name + ".******\n\n");
    std::ostringstream n; n << times.size();
    out.text("TIME\ntime set: 1\nnumber of steps: " + n.str() + "\nfilename start number: 0\nfilename increment: 1\ntime values: ");
    for (double t : times) { out.real12(t); out.text("\n"); }
  }
  std::string filename, variable_name;
  mesh_data_kind kind;
  std::size_t n_component;
  std::vector<double> times;
};

}  // namespace exporter

#endif  // TFEL_TFEL_EXPORT_HPP
