// <tfel/tfel.hpp> -- host-side C++14 face of libtfel_cuda: TFEL's user surface with the same spellings
// (reference: src/tfel.hpp:5-29 umbrella header, README.md:66-120 usage), implemented on top of the C ABI in
// include/tfel_cuda.h. Nothing here computes on the host: meshes, spaces, forms and solves live on the B200.
//
//   fe_mesh<cell::triangle> m(gen_square_mesh(1.0, 1.0, n, n));
//   submesh<cell::triangle> dm(m.get_boundary_submesh());
//   finite_element_space<cell::triangle::fe::lagrange_p1> fes(m);  fes.add_dirichlet_boundary(dm);
//   bilinear_form<fes_type, fes_type> a(fes, fes);
//   { auto u(a.get_trial_function()); auto v(a.get_test_function());
//     a += integrate<quad::triangle::qf5pT>(d<1>(u) * d<1>(v) + d<2>(u) * d<2>(v), m); }
//   linear_form<fes_type> b(fes); { auto v(b.get_test_function()); b += integrate<quad_type>(f * v, m); }
//   solver::cuda::cg_jacobi s(param);                 // next to solver::petsc::gmres_ilu (solver.hpp:122-210)
//   const fes_type::element u(a.solve(b, s));
//
// Errors are thrown by value as std::string, the reference's convention (solver.hpp:132, dictionary.hpp:46,50).
// Expression templates (expression.hpp:14-29, 198-250, 315-413) are lowered on the host to the sum-of-products form the
// C ABI takes; host callbacks double(*)(const double*) (free_function, expression.hpp:31-58) are tabulated at the
// quadrature points, device-evaluated functions of x are built from xfun::x<i>() and arithmetic.
#ifndef TFEL_TFEL_HPP
#define TFEL_TFEL_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../tfel_cuda.h"

namespace tfel_cuda {

inline void ck(int code) {
  if (code != 0) throw std::string(tfelcu_last_error());
}

// one context per process (device from TFEL_CUDA_DEVICE, default 0), created on first use
inline tfelcu_ctx* context() {
  struct holder {
    tfelcu_ctx* c = nullptr;
    holder() {
      const char* e = std::getenv("TFEL_CUDA_DEVICE");
      ck(tfelcu_ctx_create(e ? std::atoi(e) : 0, &c));
    }
    ~holder() { tfelcu_ctx_destroy(c); }
  };
  static holder h;
  return h.c;
}

// a plain device array (tfelcu_array_*): coefficient vectors of discrete functions live here between a solve, the algebra
// on elements and the next += that uses them as coefficient fields
class device_array {
public:
  explicit device_array(std::size_t count) : n(count) { ck(tfelcu_array_create(context(), count, &p)); }
  device_array(const device_array&) = delete;
  device_array& operator=(const device_array&) = delete;
  ~device_array() { if (p) tfelcu_array_destroy(context(), p); }
  double* p = nullptr;
  std::size_t n = 0;
};

}  // namespace tfel_cuda

// ---- array<T> (thomashilke/spikes array.hpp subset: what crosses the solver plugin boundary) -------------
template <typename T>
class array {
public:
  array() {}
  array(std::initializer_list<std::size_t> extents) : size(extents), data(count(), T()) {}
  std::size_t get_rank() const { return size.size(); }
  std::size_t get_size(std::size_t d) const { return size[d]; }
  std::size_t get_element_number() const { return data.size(); }
  T* get_data() { return data.data(); }
  const T* get_data() const { return data.data(); }
  void set_data(const T* p) { for (std::size_t i(0); i < data.size(); ++i) data[i] = p[i]; }
  void fill(const T& v) { for (auto& x : data) x = v; }
  T& at(std::size_t i) { return data[i]; }
  const T& at(std::size_t i) const { return data[i]; }
  T& at(std::size_t i, std::size_t j) { return data[i * size[1] + j]; }
  const T& at(std::size_t i, std::size_t j) const { return data[i * size[1] + j]; }
  // any rank, row-major (spikes array.hpp: at(i, j, k, ...))
  template <typename... more_t>
  T& at(std::size_t i, std::size_t j, std::size_t k, more_t... more) { return data[offset({i, j, k, static_cast<std::size_t>(more)...})]; }
  template <typename... more_t>
  const T& at(std::size_t i, std::size_t j, std::size_t k, more_t... more) const { return data[offset({i, j, k, static_cast<std::size_t>(more)...})]; }
private:
  std::size_t offset(std::initializer_list<std::size_t> idx) const {
    std::size_t o(0), d(0);
    for (std::size_t i : idx) { o = o * size[d] + i; ++d; }
    return o;
  }
  std::size_t count() const { std::size_t n(1); for (auto s : size) n *= s; return size.empty() ? 0 : n; }
  std::vector<std::size_t> size;
  std::vector<T> data;
};

namespace tfel_cuda {

// Coefficients of a discrete function: on the device, on the host, or both (whichever side was asked for last is
// fetched lazily). Device arrays are immutable once shared: arithmetic produces new ones.
inline std::size_t& coefficient_transfers() { static std::size_t n(0); return n; }   // host <-> device copies so far

class coefficients {
public:
  coefficients() {}
  explicit coefficients(array<double> c) : host(std::move(c)), host_valid(true), count(host.get_element_number()) {}
  explicit coefficients(std::shared_ptr<device_array> d) : dev(std::move(d)), count(dev ? dev->n : 0) {}
  std::size_t size() const { return count; }
  const array<double>& on_host() const {
    if (not host_valid) {
      host = array<double>{count};
      if (count) ck(tfelcu_array_copy(context(), host.get_data(), dev->p, count));
      host_valid = true; ++coefficient_transfers();
    }
    return host;
  }
  const std::shared_ptr<device_array>& on_device() const {
    if (not dev) {
      dev = std::make_shared<device_array>(count);
      if (count) ck(tfelcu_array_copy(context(), dev->p, host.get_data(), count));
      ++coefficient_transfers();
    }
    return dev;
  }
  // a x + b y on the device (y may be null)
  static coefficients lincomb(double a, const coefficients& x, double b, const coefficients* y) {
    std::shared_ptr<device_array> out(std::make_shared<device_array>(x.size()));
    ck(tfelcu_array_lincomb(context(), x.size(), a, x.on_device()->p, b, y ? y->on_device()->p : nullptr, out->p));
    return coefficients(out);
  }
  // [offset, offset + n) as a new device array
  coefficients slice(std::size_t offset, std::size_t n) const {
    std::shared_ptr<device_array> out(std::make_shared<device_array>(n));
    if (n) ck(tfelcu_array_copy(context(), out->p, on_device()->p + offset, n));
    return coefficients(out);
  }
private:
  mutable array<double> host;
  mutable bool host_valid = false;
  mutable std::shared_ptr<device_array> dev;
  std::size_t count = 0;
};

}  // namespace tfel_cuda

// ---- dictionary (src/core/dictionary.hpp:8-94) --------------------------------------------------------------------------
class dictionary {
public:
  template <typename T>
  dictionary& set(const std::string& key, const T& v) { put(key, v); return *this; }
  template <typename T>
  T get(const std::string& key) const {
    auto it(items.find(key));
    if (it == items.end()) throw std::string("dictionary::get: key ") + key + " does not exist";
    return convert(it->second, static_cast<T*>(nullptr), key);
  }
  bool key_exists(const std::string& key) const { return items.count(key) != 0; }
  bool keys_exist(std::initializer_list<std::string> keys) const {
    for (const auto& k : keys) if (not key_exists(k)) return false;
    return true;
  }
  void print(std::ostream& out = std::cout) const {
    for (const auto& kv : items) {
      out << kv.first << ": ";
      switch (kv.second.type) {
        case 0: out << kv.second.d; break; case 1: out << kv.second.u; break; case 2: out << kv.second.i; break;
        case 3: out << kv.second.s; break; default: out << (kv.second.b ? "true" : "false");
      }
      out << std::endl;
    }
  }
private:
  struct value { int type; double d; unsigned u; int i; std::string s; bool b; };
  void put(const std::string& k, double v) { items[k] = value{0, v, 0u, 0, "", false}; }
  void put(const std::string& k, unsigned v) { items[k] = value{1, 0.0, v, 0, "", false}; }
  void put(const std::string& k, int v) { items[k] = value{2, 0.0, 0u, v, "", false}; }
  void put(const std::string& k, const std::string& v) { items[k] = value{3, 0.0, 0u, 0, v, false}; }
  void put(const std::string& k, const char* v) { items[k] = value{3, 0.0, 0u, 0, v, false}; }
  void put(const std::string& k, bool v) { items[k] = value{4, 0.0, 0u, 0, "", v}; }
  static std::string bad(const std::string& key) { return std::string("dictionary::get: wrong type for key ") + key; }
  static double convert(const value& v, double*, const std::string& key) { if (v.type != 0) throw bad(key); return v.d; }
  static unsigned convert(const value& v, unsigned*, const std::string& key) { if (v.type != 1) throw bad(key); return v.u; }
  static int convert(const value& v, int*, const std::string& key) { if (v.type != 2) throw bad(key); return v.i; }
  static std::string convert(const value& v, std::string*, const std::string& key) { if (v.type != 3) throw bad(key); return v.s; }
  static bool convert(const value& v, bool*, const std::string& key) { if (v.type != 4) throw bad(key); return v.b; }
  std::map<std::string, value> items;
};

// ---- cells, finite elements, quadrature tags (cell.hpp, fe.hpp, quadrature.hpp) ---------------------------------------
namespace cell {
struct point { static const int n_vertex = 1; };
struct edge { static const int n_vertex = 2; };
struct triangle {
  static const int dim = 2;
  using boundary_cell_type = edge;
  struct fe {
    struct lagrange_p1 { using cell_type = triangle; static const int id = TFELCU_FE_P1; };          // fe.hpp:260-317
    struct lagrange_p1_bubble { using cell_type = triangle; static const int id = TFELCU_FE_P1_BUBBLE; };  // :319-387
    struct lagrange_p2 { using cell_type = triangle; static const int id = TFELCU_FE_P2; };          // :389-457
  };
};
struct tetrahedron {
  static const int dim = 3;
  using boundary_cell_type = triangle;
  struct fe {
    struct lagrange_p1 { using cell_type = tetrahedron; static const int id = TFELCU_FE_P1; };        // fe.hpp:494-562
    struct lagrange_p1_bubble { using cell_type = tetrahedron; static const int id = TFELCU_FE_P1_BUBBLE; };  // :564-654
    struct lagrange_p2 { using cell_type = tetrahedron; static const int id = TFELCU_FE_P2; };        // new (cell.hpp:948-952)
  };
};
}  // namespace cell

namespace quad {
namespace triangle {
struct qf1pT { static const int id = TFELCU_QUAD_TRI_QF1PT; };
struct qf2pT { static const int id = TFELCU_QUAD_TRI_QF2PT; };
struct qf5pT { static const int id = TFELCU_QUAD_TRI_QF5PT; };
struct qf1pTlump { static const int id = TFELCU_QUAD_TRI_QF1PTLUMP; };
}
namespace tetrahedron {
struct qf1pTet { static const int id = TFELCU_QUAD_TET_QF1PTET; };
struct qf4pTet { static const int id = TFELCU_QUAD_TET_QF4PTET; };
struct qf5pTet { static const int id = TFELCU_QUAD_TET_QF5PTET; };
struct qfSym1pTet { static const int id = TFELCU_QUAD_TET_SYM1; };
struct qfSym4pTet { static const int id = TFELCU_QUAD_TET_SYM4; };
struct qfSym10pTet { static const int id = TFELCU_QUAD_TET_SYM10; };
struct qfSym20pTet { static const int id = TFELCU_QUAD_TET_SYM20; };
struct qfSym35pTet { static const int id = TFELCU_QUAD_TET_SYM35; };
struct qfSym56pTet { static const int id = TFELCU_QUAD_TET_SYM56; };
}
}  // namespace quad

// ---- meshes (mesh.hpp:142-158, 355-364, 470-521; mesh.cpp:19-135) ------------------------------------------------------
template <typename cell_t, typename sub_cell_t = typename cell_t::boundary_cell_type>
class submesh;

template <typename cell_t>
class mesh {
public:
  using cell_type = cell_t;
  mesh(const double* vertices, std::size_t n_vertices, const unsigned int* cells, std::size_t n_cells) {
    tfel_cuda::ck(tfelcu_mesh_create(tfel_cuda::context(), cell_t::dim, vertices, n_vertices, cells, n_cells, &h));
  }
  explicit mesh(tfelcu_mesh* handle) : h(handle) {}
  mesh(mesh&& o) noexcept : h(o.h) { o.h = nullptr; }
  mesh(const mesh&) = delete;
  mesh& operator=(const mesh&) = delete;
  ~mesh() { if (h) tfelcu_mesh_destroy(h); }
  std::size_t get_embedding_space_dimension() const { return cell_t::dim; }
  std::size_t get_vertex_number() const { std::size_t v(0); tfel_cuda::ck(tfelcu_mesh_info(h, nullptr, &v, nullptr)); return v; }
  std::size_t get_cell_number() const { std::size_t k(0); tfel_cuda::ck(tfelcu_mesh_info(h, nullptr, nullptr, &k)); return k; }
  array<double> get_vertices() const {
    array<double> v{get_vertex_number(), static_cast<std::size_t>(cell_t::dim)};
    tfel_cuda::ck(tfelcu_mesh_download(h, v.get_data(), nullptr));
    return v;
  }
  array<unsigned int> get_cells() const {
    array<unsigned int> c{get_cell_number(), static_cast<std::size_t>(cell_t::dim + 1)};
    tfel_cuda::ck(tfelcu_mesh_download(h, nullptr, c.get_data()));
    return c;
  }
  // mesh.hpp:214-216 (the reference returns the int table entry as std::size_t: -1 wraps, kept)
  std::size_t get_cell_neighbour(std::size_t k, std::size_t face_id) const {
    if (neighbours.get_rank() == 0) {
      neighbours = array<int>{get_cell_number(), static_cast<std::size_t>(cell_t::dim + 1)};
      tfel_cuda::ck(tfelcu_mesh_neighbours(h, neighbours.get_data()));
    }
    return neighbours.at(k, face_id);
  }
  tfelcu_mesh* handle() const { return h; }
protected:
  tfelcu_mesh* h = nullptr;
  mutable array<int> neighbours;   // compute_cell_neighbours (mesh.hpp:301-332), fetched from the device on first use
};

template <typename cell_t>
class fe_mesh : public mesh<cell_t> {
public:
  fe_mesh(mesh<cell_t>&& m) : mesh<cell_t>(std::move(m)) {}
  // per-cell |K| and J^{-T} (cell.hpp:474-502, 799-837); computed on device, returned for inspection only
  array<double> get_cell_volumes() const {
    const std::size_t K(this->get_cell_number());
    array<double> vol{K}, jmt{K, static_cast<std::size_t>(cell_t::dim * cell_t::dim)};
    tfel_cuda::ck(tfelcu_mesh_geometry(this->h, vol.get_data(), jmt.get_data()));
    return vol;
  }
  submesh<cell_t> get_boundary_submesh() const;                                   // mesh.hpp:485-521
  submesh<cell_t, cell::point> get_point_submesh(std::size_t k) const;            // mesh.hpp:470-476
};

template <typename cell_t, typename sub_cell_t>
class submesh {
public:
  using cell_type = sub_cell_t;
  submesh(const fe_mesh<cell_t>& m, bool whole, array<unsigned int> el) : parent(&m), whole_boundary(whole), elements(std::move(el)) {}
  std::size_t get_cell_number() const { return elements.get_size(0); }
  const array<unsigned int>& get_cells() const { return elements; }
  const fe_mesh<cell_t>& get_mesh() const { return *parent; }
  bool is_whole_boundary() const { return whole_boundary; }
private:
  const fe_mesh<cell_t>* parent;
  bool whole_boundary;
  array<unsigned int> elements;   // [F][n_vertex]
};

template <typename cell_t>
submesh<cell_t> fe_mesh<cell_t>::get_boundary_submesh() const {
  std::size_t F(0);
  tfel_cuda::ck(tfelcu_mesh_boundary(this->h, &F));
  array<unsigned int> el{F, static_cast<std::size_t>(cell_t::dim)};
  tfel_cuda::ck(tfelcu_mesh_boundary_download(this->h, el.get_data(), nullptr, nullptr));
  return submesh<cell_t>(*this, true, std::move(el));
}

template <typename cell_t>
submesh<cell_t, cell::point> fe_mesh<cell_t>::get_point_submesh(std::size_t k) const {
  const array<unsigned int> c(this->get_cells());
  array<unsigned int> el{1, 1};
  el.at(0, 0) = c.at(k, 0);
  return submesh<cell_t, cell::point>(*this, false, std::move(el));
}

inline mesh<cell::triangle> gen_square_mesh(double x_1, double x_2, unsigned int n_1, unsigned int n_2) {
  tfelcu_mesh* h(nullptr);
  tfel_cuda::ck(tfelcu_mesh_gen_square(tfel_cuda::context(), x_1, x_2, n_1, n_2, &h));
  return mesh<cell::triangle>(h);
}

inline mesh<cell::tetrahedron> gen_cube_mesh(double x_1, double x_2, double x_3, unsigned int n_1, unsigned int n_2, unsigned int n_3) {
  tfelcu_mesh* h(nullptr);
  tfel_cuda::ck(tfelcu_mesh_gen_cube(tfel_cuda::context(), x_1, x_2, x_3, n_1, n_2, n_3, &h));
  return mesh<cell::tetrahedron>(h);
}

// ---- finite element spaces (fes.hpp:18-149, composite_fes.hpp:41-122) -----------------------------------------------
template <typename fe_t>
class finite_element_space {
public:
  using fe_type = fe_t;
  using cell_type = typename fe_t::cell_type;
  static const std::size_t n_component = 1;

  // fes.hpp:219-364: coefficient vector of a discrete function. The coefficients stay on the device between a solve,
  // the algebra below (fused vector kernels, tfelcu_array_lincomb) and their use as coefficient fields of the next +=
  // (make_expr); get_coefficients() / evaluate() fetch a host copy on first use.
  class element {
  public:
    element(const finite_element_space& f, array<double> c) : fes(&f), coef(std::move(c)) {}
    element(const finite_element_space& f, tfel_cuda::coefficients c) : fes(&f), coef(std::move(c)) {}
    const finite_element_space& get_finite_element_space() const { return *fes; }
    const array<double>& get_coefficients() const { return coef.on_host(); }
    const tfel_cuda::coefficients& get_storage() const { return coef; }
    // fes.hpp:247-255: sum_n coefficients[dof(k, n)] phi_n(x_hat) (the dof map is fetched from the device once per space)
    double evaluate(std::size_t k, const double* x_hat) const {
      const array<unsigned int>& dm(fes->cached_dof_map());
      const array<double>& c(coef.on_host());
      double table[10 * 4];
      int nd(0);
      tfel_cuda::ck(tfelcu_fe_tabulate(cell_type::dim, fe_t::id, x_hat, table, &nd));
      double value(0.0);
      for (int n(0); n < nd; ++n) value += c.at(dm.at(k, n)) * table[n * (cell_type::dim + 1)];
      return value;
    }
    // an arbitrary host function of the coefficients (host round trip: the function is host code)
    element& transform(const std::function<double(double)>& f) {
      array<double> c(coef.on_host());
      for (std::size_t i(0); i < c.get_element_number(); ++i) c.at(i) = f(c.at(i));
      coef = tfel_cuda::coefficients(std::move(c));
      return *this;
    }
    element transformed(const std::function<double(double)>& f) const { element r(*this); r.transform(f); return r; }
    // coefficient-wise algebra of fes.hpp:314-359 (discrete functions of one space are a vector space), on the device
    element& operator*=(double s) { coef = tfel_cuda::coefficients::lincomb(s, coef, 0.0, nullptr); return *this; }
    element& operator/=(double s) { coef = tfel_cuda::coefficients::lincomb(1.0 / s, coef, 0.0, nullptr); return *this; }
    element& operator+=(const element& o) { same_space(o); coef = tfel_cuda::coefficients::lincomb(1.0, coef, 1.0, &o.coef); return *this; }
    element& operator-=(const element& o) { same_space(o); coef = tfel_cuda::coefficients::lincomb(1.0, coef, -1.0, &o.coef); return *this; }
    // fes.hpp:366-405 (hidden friends: the reference's free templates cannot deduce fe from a nested type)
    friend element operator+(element a, const element& b) { a += b; return a; }
    friend element operator-(element a, const element& b) { a -= b; return a; }
    friend element operator*(double s, element a) { a *= s; return a; }
    friend element operator*(element a, double s) { a *= s; return a; }
    friend element operator/(element a, double s) { a /= s; return a; }
  private:
    void same_space(const element& o) const {
      if (fes != o.fes) throw std::string("operations between elements of different finite element spaces are not supported.");
    }
    const finite_element_space* fes;
    tfel_cuda::coefficients coef;
  };

  explicit finite_element_space(const fe_mesh<cell_type>& m) : msh(&m) {
    tfel_cuda::ck(tfelcu_fes_create(tfel_cuda::context(), m.handle(), fe_t::id, &h));
  }
  // fes.hpp:84-98: space + Dirichlet data on a submesh in one go
  template <typename sm_t>
  finite_element_space(const fe_mesh<cell_type>& m, const sm_t& dm, double value = 0.0) : finite_element_space(m) {
    add_dirichlet_boundary(dm, value);
  }
  template <typename sm_t>
  finite_element_space(const fe_mesh<cell_type>& m, const sm_t& dm, double (*f)(const double*)) : finite_element_space(m) {
    add_dirichlet_boundary(dm, f);
  }
  template <typename sm_t>
  finite_element_space(const fe_mesh<cell_type>& m, const sm_t& dm, const std::function<double(const double*)>& f)
    : finite_element_space(m) {
    add_dirichlet_boundary(dm, f);
  }
  finite_element_space(const finite_element_space&) = delete;
  ~finite_element_space() { if (h) tfelcu_fes_destroy(h); }

  std::size_t get_dof_number() const { std::size_t n(0); tfel_cuda::ck(tfelcu_fes_info(h, nullptr, &n, nullptr)); return n; }
  const fe_mesh<cell_type>& get_mesh() const { return *msh; }
  std::size_t get_dirichlet_dof_number() const { std::size_t n(0); tfel_cuda::ck(tfelcu_fes_dirichlet_count(h, &n)); return n; }
  array<unsigned int> get_dof_map() const {
    int nd(0); tfel_cuda::ck(tfelcu_fes_info(h, nullptr, nullptr, &nd));
    array<unsigned int> dm{msh->get_cell_number(), static_cast<std::size_t>(nd)};
    tfel_cuda::ck(tfelcu_fes_download_dof_map(h, dm.get_data()));
    return dm;
  }

  // add_dirichlet_boundary(dm [, value | function of the dof coordinate]), fes.hpp:100-149
  template <typename sm_t>
  void add_dirichlet_boundary(const sm_t& dm, double value = 0.0) {
    select(dm);
    tfel_cuda::ck(tfelcu_fes_add_dirichlet_const(h, value));
  }
  template <typename sm_t>
  void add_dirichlet_boundary(const sm_t& dm, const std::function<double(const double*)>& f) {
    const std::size_t n(select(dm));
    if (n == 0) return;
    const std::size_t dim(cell_type::dim);
    std::vector<unsigned int> dofs(n);
    std::vector<double> x(n * dim), values(n);
    tfel_cuda::ck(tfelcu_fes_boundary_dofs_get(h, dofs.data(), x.data()));
    for (std::size_t i(0); i < n; ++i) values[i] = f(&x[i * dim]);
    tfel_cuda::ck(tfelcu_fes_add_dirichlet(h, dofs.data(), values.data(), n));
  }
  template <typename sm_t>
  void add_dirichlet_boundary(const sm_t& dm, double (*f)(const double*)) {
    add_dirichlet_boundary(dm, std::function<double(const double*)>(f));
  }

  std::size_t get_dof(std::size_t k, std::size_t n) const { return cached_dof_map().at(k, n); }   // fes.hpp:151-153
  const array<unsigned int>& cached_dof_map() const {
    if (dof_map_cache.get_rank() == 0) dof_map_cache = get_dof_map();
    return dof_map_cache;
  }
  tfelcu_fes* handle() const { return h; }
private:
  mutable array<unsigned int> dof_map_cache;
  template <typename sm_t>
  std::size_t select(const sm_t& dm) {
    std::size_t n(0);
    if (dm.is_whole_boundary())
      tfel_cuda::ck(tfelcu_fes_boundary_dofs(h, nullptr, 0, 0, &n));
    else
      tfel_cuda::ck(tfelcu_fes_boundary_dofs(h, dm.get_cells().get_data(), dm.get_cell_number(),
                                             static_cast<int>(dm.get_cells().get_size(1)), &n));
    return n;
  }
  const fe_mesh<cell_type>* msh;
  tfelcu_fes* h = nullptr;
};

template <typename... fe_pack>
struct composite_finite_element {
  static const std::size_t n_component = sizeof...(fe_pack);
  using fe_list = std::tuple<fe_pack...>;
  using cell_type = typename std::tuple_element<0, fe_list>::type::cell_type;
};

template <typename cfe_t>
class composite_finite_element_space;

template <typename... fe_pack>
class composite_finite_element_space<composite_finite_element<fe_pack...>> {
public:
  using cfe_type = composite_finite_element<fe_pack...>;
  using cell_type = typename cfe_type::cell_type;
  static const std::size_t n_component = sizeof...(fe_pack);
  template <std::size_t n>
  using component_space = finite_element_space<typename std::tuple_element<n, std::tuple<fe_pack...>>::type>;

  class element {   // composite_fes.hpp:152-300; coefficients [u0 | u1 | p] on the device like the scalar element's
  public:
    element(const composite_finite_element_space& f, array<double> c) : fes(&f), coef(std::move(c)) {}
    element(const composite_finite_element_space& f, tfel_cuda::coefficients c) : fes(&f), coef(std::move(c)) {}
    // from one element per component (composite_fes.hpp:160-175)
    template <typename first_t, typename... rest_t,
              typename = typename std::enable_if<!std::is_same<typename std::decay<first_t>::type, array<double>>::value and
                                                 !std::is_same<typename std::decay<first_t>::type, tfel_cuda::coefficients>::value>::type>
    element(const composite_finite_element_space& f, const first_t& c0, const rest_t&... cs) : fes(&f) {
      std::shared_ptr<tfel_cuda::device_array> all(std::make_shared<tfel_cuda::device_array>(f.get_total_dof_number()));
      std::size_t at(0);
      append(*all, at, c0, cs...);
      coef = tfel_cuda::coefficients(all);
    }
    const array<double>& get_coefficients() const { return coef.on_host(); }
    const tfel_cuda::coefficients& get_storage() const { return coef; }
    template <std::size_t n>
    typename component_space<n>::element get_component() const {
      const component_space<n>& sp(fes->template get_finite_element_space<n>());
      return typename component_space<n>::element(sp, coef.slice(fes->get_component_offset(n), sp.get_dof_number()));
    }
  private:
    void append(tfel_cuda::device_array&, std::size_t&) {}
    template <typename first_t, typename... rest_t>
    void append(tfel_cuda::device_array& all, std::size_t& at, const first_t& c0, const rest_t&... cs) {
      const tfel_cuda::coefficients& c(c0.get_storage());
      if (c.size()) tfel_cuda::ck(tfelcu_array_copy(tfel_cuda::context(), all.p + at, c.on_device()->p, c.size()));
      at += c.size();
      append(all, at, cs...);
    }
    const composite_finite_element_space* fes;
    tfel_cuda::coefficients coef;
  };

  explicit composite_finite_element_space(const fe_mesh<cell_type>& m)
    : msh(&m), spaces(std::unique_ptr<finite_element_space<fe_pack>>(new finite_element_space<fe_pack>(m))...) {}

  template <std::size_t n>
  const component_space<n>& get_finite_element_space() const { return *std::get<n>(spaces); }
  template <std::size_t n, typename sm_t, typename... rest>
  void add_dirichlet_boundary(const sm_t& dm, rest... value) { std::get<n>(spaces)->add_dirichlet_boundary(dm, value...); }

  std::size_t get_total_dof_number() const { return get_component_offset(n_component); }
  template <std::size_t n>
  std::size_t get_dof_number() const { return std::get<n>(spaces)->get_dof_number(); }
  std::size_t get_component_offset(std::size_t c) const {   // blocks concatenated: [u0 | u1 | p] (composite_bilinear_form.hpp:41-62)
    std::vector<std::size_t> n;
    dof_numbers(n, std::make_index_sequence<n_component>());
    std::size_t off(0);
    for (std::size_t i(0); i < c; ++i) off += n[i];
    return off;
  }
  void handles(std::vector<tfelcu_fes*>& out) const { collect(out, std::make_index_sequence<n_component>()); }
  const fe_mesh<cell_type>& get_mesh() const { return *msh; }
private:
  template <std::size_t... i>
  void collect(std::vector<tfelcu_fes*>& out, std::index_sequence<i...>) const { out = {std::get<i>(spaces)->handle()...}; }
  template <std::size_t... i>
  void dof_numbers(std::vector<std::size_t>& out, std::index_sequence<i...>) const { out = {std::get<i>(spaces)->get_dof_number()...}; }
  const fe_mesh<cell_type>* msh;
  std::tuple<std::unique_ptr<finite_element_space<fe_pack>>...> spaces;
};

namespace tfel_cuda {
template <typename fes_t> struct space_traits;
template <typename fe_t>
struct space_traits<finite_element_space<fe_t>> {
  static void handles(const finite_element_space<fe_t>& f, std::vector<tfelcu_fes*>& out) { out = {f.handle()}; }
  static std::size_t total(const finite_element_space<fe_t>& f) { return f.get_dof_number(); }
};
template <typename cfe_t>
struct space_traits<composite_finite_element_space<cfe_t>> {
  static void handles(const composite_finite_element_space<cfe_t>& f, std::vector<tfelcu_fes*>& out) { f.handles(out); }
  static std::size_t total(const composite_finite_element_space<cfe_t>& f) { return f.get_total_dof_number(); }
};
}  // namespace tfel_cuda

// ---- expressions (expression.hpp) -------------------------------------------------------------------------------------
namespace tfel_cuda {

struct factor_desc {
  int kind = 0;                                      // TFELCU_FACTOR_* ; 0 = host callback to be tabulated
  int d = 0;
  std::shared_ptr<void> device_table;                // TABLE: double[K][nq] on the device (tfel_device.cuh functors)
  int table_quad = 0;                                // TABLE: quadrature formula the table was evaluated for
  const tfelcu_fes* fes = nullptr;                   // FIELD
  coefficients data;                                 // FIELD coefficients (device resident)
  std::function<double(const double*)> fn;           // host callback (free_function / std_function)
  std::vector<tfelcu_op> prog;                       // PROGRAM
  tfelcu_tabulate_fn device_fn = nullptr;            // FUNCTION: launcher of the kernel instantiated in the user's nvcc TU
  std::shared_ptr<void> device_closure;              // FUNCTION: the functor object in device memory (tfel_device.cuh)
};

// Host callbacks are the FALLBACK for integrands: a free_function leaf compiled by a host compiler cannot run on the GPU,
// so it is evaluated here at every quadrature point of every cell and uploaded (expression.hpp:31-58 evaluates it in the
// cell loop). It reports itself: the counter below, and once per process a note on stderr. Device alternatives without
// any host evaluation: __device__ functors (tfel_device.cuh: dev(f) * v, compiled by nvcc) or xfun expressions.
inline unsigned long long& host_callback_evaluations() { static unsigned long long n(0); return n; }
inline void note_host_callbacks(unsigned long long n) {
  host_callback_evaluations() += n;
  static bool said(false);
  if (not said and not std::getenv("TFEL_CUDA_QUIET")) {
    said = true;
    std::cerr << "tfel_cuda: an integrand contains a host function (free_function): " << n << " evaluations on the host for this "
                 "integral; use a __device__ functor (tfel/tfel_device.cuh) or xfun for device-side evaluation" << std::endl;
  }
}


struct monomial {
  double c = 1.0;
  int test_comp = -1, test_d = 0, trial_comp = -1, trial_d = 0;
  std::vector<int> factors;
};

}  // namespace tfel_cuda

// has_form == false: rank-0 expression (coefficients only); true: contains a test and / or a trial function
template <bool has_form>
class expression {
public:
  std::vector<tfel_cuda::monomial> terms;
  std::vector<tfel_cuda::factor_desc> factors;
};

namespace tfel_cuda {

template <bool a>
inline int import_factors(const expression<a>& src, std::vector<factor_desc>& dst) {
  const int base(static_cast<int>(dst.size()));
  dst.insert(dst.end(), src.factors.begin(), src.factors.end());
  return base;
}

// c * (one finite element field, possibly differentiated) and nothing else
template <bool a>
inline bool is_plain_field(const expression<a>& x) {
  return x.terms.size() == 1 and x.terms[0].test_comp < 0 and x.terms[0].trial_comp < 0 and x.terms[0].factors.size() == 1 and
         x.factors[x.terms[0].factors[0]].kind == TFELCU_FACTOR_FIELD;
}

template <bool r, bool a, bool b>
inline expression<r> add(const expression<a>& x, const expression<b>& y, double sign) {
  expression<r> out;
  if (is_plain_field(x) and is_plain_field(y)) {
    // fields of one space are a vector space: u_h +- v_h is the field of the summed coefficients, evaluated pointwise
    // like the reference does (no cancellation between expanded products in (u_h - v_h) * (u_h - v_h))
    const factor_desc& fx(x.factors[x.terms[0].factors[0]]);
    const factor_desc& fy(y.factors[y.terms[0].factors[0]]);
    if (fx.fes == fy.fes and fx.d == fy.d and fx.data.size() == fy.data.size()) {
      factor_desc f(fx);
      f.data = coefficients::lincomb(x.terms[0].c, fx.data, sign * y.terms[0].c, &fy.data);   // on the device
      out.factors.push_back(f);
      monomial m; m.factors.push_back(0); out.terms.push_back(m);
      return out;
    }
  }
  out.factors = x.factors; out.terms = x.terms;
  const int base(import_factors(y, out.factors));
  for (monomial m : y.terms) { m.c *= sign; for (auto& f : m.factors) f += base; out.terms.push_back(m); }
  return out;
}

template <bool r, bool a, bool b>
inline expression<r> mul(const expression<a>& x, const expression<b>& y) {
  expression<r> out;
  out.factors = x.factors;
  const int base(import_factors(y, out.factors));
  for (const monomial& p : x.terms)
    for (const monomial& q : y.terms) {
      monomial m(p);
      m.c = p.c * q.c;
      if (q.test_comp >= 0) {
        if (p.test_comp >= 0) throw std::string("expression: product of two test functions");
        m.test_comp = q.test_comp; m.test_d = q.test_d;
      }
      if (q.trial_comp >= 0) {
        if (p.trial_comp >= 0) throw std::string("expression: product of two trial functions");
        m.trial_comp = q.trial_comp; m.trial_d = q.trial_d;
      }
      for (int f : q.factors) m.factors.push_back(f + base);
      out.terms.push_back(m);
    }
  return out;
}

inline expression<false> constant(double c) { expression<false> e; monomial m; m.c = c; e.terms.push_back(m); return e; }
inline expression<false> callback(std::function<double(const double*)> f) {
  expression<false> e; factor_desc d; d.kind = 0; d.fn = std::move(f); e.factors.push_back(d);
  monomial m; m.factors.push_back(0); e.terms.push_back(m); return e;
}

}  // namespace tfel_cuda

template <bool a, bool b> expression<a || b> operator+(const expression<a>& x, const expression<b>& y) { return tfel_cuda::add<a || b>(x, y, 1.0); }
template <bool a, bool b> expression<a || b> operator-(const expression<a>& x, const expression<b>& y) { return tfel_cuda::add<a || b>(x, y, -1.0); }
template <bool a, bool b> expression<a || b> operator*(const expression<a>& x, const expression<b>& y) { return tfel_cuda::mul<a || b>(x, y); }
template <bool a> expression<a> operator-(const expression<a>& x) { return tfel_cuda::mul<a>(tfel_cuda::constant(-1.0), x); }
template <bool a> expression<a> operator*(double c, const expression<a>& x) { return tfel_cuda::mul<a>(tfel_cuda::constant(c), x); }   // expression.hpp:376-413
template <bool a> expression<a> operator*(const expression<a>& x, double c) { return tfel_cuda::mul<a>(x, tfel_cuda::constant(c)); }
template <bool a> expression<a> operator+(double c, const expression<a>& x) { return tfel_cuda::add<a>(tfel_cuda::constant(c), x, 1.0); }
template <bool a> expression<a> operator+(const expression<a>& x, double c) { return tfel_cuda::add<a>(x, tfel_cuda::constant(c), 1.0); }
template <bool a> expression<a> operator*(double (*f)(const double*), const expression<a>& x) { return tfel_cuda::mul<a>(tfel_cuda::callback(f), x); }   // free_function, expression.hpp:392-399
template <bool a> expression<a> operator*(const expression<a>& x, double (*f)(const double*)) { return tfel_cuda::mul<a>(x, tfel_cuda::callback(f)); }
inline expression<false> make_expr(double (*f)(const double*)) { return tfel_cuda::callback(f); }
inline expression<false> make_expr(const std::function<double(const double*)>& f) { return tfel_cuda::callback(f); }
inline expression<false> make_expr(double c) { return tfel_cuda::constant(c); }

// coefficient field: finite_element_function (expression.hpp:89-139)
template <typename fe_t>
expression<false> make_expr(const typename finite_element_space<fe_t>::element& u) {
  expression<false> e;
  tfel_cuda::factor_desc d;
  d.kind = TFELCU_FACTOR_FIELD; d.d = 0; d.fes = u.get_finite_element_space().handle();
  d.data = u.get_storage();   // shares the device array of the element
  e.factors.push_back(d);
  tfel_cuda::monomial m; m.factors.push_back(0); e.terms.push_back(m);
  return e;
}

// d<k>(expr): derivative along x_k with the product rule (expression.hpp:315-373); constants and functions of x that are
// not finite element fields are not differentiable here
template <std::size_t k, bool a>
expression<a> d(const expression<a>& x) {
  expression<a> out;
  out.factors = x.factors;
  for (const tfel_cuda::monomial& m : x.terms) {
    bool any(false);
    if (m.test_comp >= 0) {
      if (m.test_d != 0) throw std::string("d<k>: second derivatives are not available");
      tfel_cuda::monomial t(m); t.test_d = static_cast<int>(k); out.terms.push_back(t); any = true;
    }
    if (m.trial_comp >= 0) {
      if (m.trial_d != 0) throw std::string("d<k>: second derivatives are not available");
      tfel_cuda::monomial t(m); t.trial_d = static_cast<int>(k); out.terms.push_back(t); any = true;
    }
    for (std::size_t i(0); i < m.factors.size(); ++i) {
      const tfel_cuda::factor_desc& f(out.factors[m.factors[i]]);
      if (f.kind != TFELCU_FACTOR_FIELD) throw std::string("d<k>: only finite element fields and test / trial functions can be differentiated");
      if (f.d != 0) throw std::string("d<k>: second derivatives are not available");
      tfel_cuda::factor_desc g(f); g.d = static_cast<int>(k);
      out.factors.push_back(g);
      tfel_cuda::monomial t(m); t.factors[i] = static_cast<int>(out.factors.size()) - 1; out.terms.push_back(t); any = true;
    }
    (void)any;   // a constant differentiates to nothing
  }
  return out;
}

// functions of x evaluated ON DEVICE (no host callback): xfun::x<0>() * xfun::x<0>() + 0.5 * xfun::x<1>() ...
namespace xfun {
class fun {
public:
  std::vector<tfelcu_op> prog;
  operator expression<false>() const {
    expression<false> e; tfel_cuda::factor_desc d; d.kind = TFELCU_FACTOR_PROGRAM; d.prog = prog; e.factors.push_back(d);
    tfel_cuda::monomial m; m.factors.push_back(0); e.terms.push_back(m); return e;
  }
};
inline fun leaf(int op, double imm) { fun f; tfelcu_op o; o.op = op; o.pad = 0; o.imm = imm; f.prog.push_back(o); return f; }
inline fun join(const fun& a, const fun& b, int op) {
  fun f(a); f.prog.insert(f.prog.end(), b.prog.begin(), b.prog.end());
  tfelcu_op o; o.op = op; o.pad = 0; o.imm = 0.0; f.prog.push_back(o); return f;
}
inline fun unary(const fun& a, int op) { fun f(a); tfelcu_op o; o.op = op; o.pad = 0; o.imm = 0.0; f.prog.push_back(o); return f; }
template <int i> fun x() { return leaf(TFELCU_OP_COORD, i); }
inline fun c(double v) { return leaf(TFELCU_OP_CONST, v); }
inline fun operator+(const fun& a, const fun& b) { return join(a, b, TFELCU_OP_ADD); }
inline fun operator-(const fun& a, const fun& b) { return join(a, b, TFELCU_OP_SUB); }
inline fun operator*(const fun& a, const fun& b) { return join(a, b, TFELCU_OP_MUL); }
inline fun operator/(const fun& a, const fun& b) { return join(a, b, TFELCU_OP_DIV); }
inline fun operator+(const fun& a, double b) { return a + c(b); }
inline fun operator+(double a, const fun& b) { return c(a) + b; }
inline fun operator-(const fun& a, double b) { return a - c(b); }
inline fun operator-(double a, const fun& b) { return c(a) - b; }
inline fun operator*(double a, const fun& b) { return c(a) * b; }
inline fun operator*(const fun& a, double b) { return a * c(b); }
inline fun operator/(const fun& a, double b) { return a / c(b); }
inline fun operator-(const fun& a) { return unary(a, TFELCU_OP_NEG); }
inline fun sqrt(const fun& a) { return unary(a, TFELCU_OP_SQRT); }
inline fun exp(const fun& a) { return unary(a, TFELCU_OP_EXP); }
inline fun sin(const fun& a) { return unary(a, TFELCU_OP_SIN); }
inline fun cos(const fun& a) { return unary(a, TFELCU_OP_COS); }
inline fun abs(const fun& a) { return unary(a, TFELCU_OP_ABS); }
inline fun pow(const fun& a, double p) { return join(a, c(p), TFELCU_OP_POW); }
inline fun less(const fun& a, const fun& b) { return join(a, b, TFELCU_OP_LT); }
}  // namespace xfun
template <bool a> expression<a> operator*(const xfun::fun& f, const expression<a>& x) { return tfel_cuda::mul<a>(static_cast<expression<false>>(f), x); }
template <bool a> expression<a> operator*(const expression<a>& x, const xfun::fun& f) { return tfel_cuda::mul<a>(x, static_cast<expression<false>>(f)); }

// ---- integrate<quad>(expr, mesh) (form.hpp:71-139) --------------------------------------------------------------------
template <bool has_form>
struct integration_proxy {
  int quad_id;
  expression<has_form> integrand;
  tfelcu_mesh* mesh;
};

namespace tfel_cuda {

// C arrays of a lowered integrand; host callbacks are tabulated at the quadrature points of every cell
class lowered {
public:
  template <bool a>
  lowered(const expression<a>& e, int quad_id, tfelcu_mesh* mesh) {
    std::vector<double> xq;
    int nq(0), dim(0);
    std::size_t K(0);
    // the same coefficient (same field + derivative, same callback, same program) appears once in the C arrays
    std::vector<int> remap(e.factors.size(), -1);
    std::vector<std::size_t> kept;
    for (std::size_t i(0); i < e.factors.size(); ++i) {
      const factor_desc& f(e.factors[i]);
      for (std::size_t k(0); k < kept.size() and remap[i] < 0; ++k) {
        const factor_desc& g(e.factors[kept[k]]);
        bool same(f.kind == g.kind and f.d == g.d);
        if (same and f.kind == TFELCU_FACTOR_FIELD) same = f.fes == g.fes and f.data.on_device().get() == g.data.on_device().get();
        else if (same and f.kind == TFELCU_FACTOR_TABLE) same = f.device_table.get() == g.device_table.get();
        else if (same and f.kind == TFELCU_FACTOR_FUNCTION) same = f.device_fn == g.device_fn and f.device_closure.get() == g.device_closure.get();
        else if (same and f.kind == TFELCU_FACTOR_PROGRAM) {
          same = f.prog.size() == g.prog.size();
          for (std::size_t t(0); same and t < f.prog.size(); ++t) same = f.prog[t].op == g.prog[t].op and f.prog[t].imm == g.prog[t].imm;
        } else if (same) {
          typedef double (*fn_t)(const double*);
          const fn_t* pf(f.fn.template target<fn_t>()); const fn_t* pg(g.fn.template target<fn_t>());
          same = pf and pg and *pf == *pg;
        }
        if (same) remap[i] = static_cast<int>(k);
      }
      if (remap[i] < 0) { remap[i] = static_cast<int>(kept.size()); kept.push_back(i); }
    }
    factors.resize(kept.size());
    for (std::size_t i(0); i < kept.size(); ++i) {
      const factor_desc& f(e.factors[kept[i]]);
      tfelcu_factor& o(factors[i]);
      o.kind = f.kind; o.d = f.d; o.fes = f.fes; o.data = nullptr; o.ops = nullptr; o.n_ops = 0; o.pad = 0;
      o.fn = nullptr; o.closure = nullptr;
      if (f.kind == TFELCU_FACTOR_FIELD) o.data = f.data.on_device()->p;   // a device pointer: nothing is uploaded
      else if (f.kind == TFELCU_FACTOR_FUNCTION) { o.fn = f.device_fn; o.closure = f.device_closure.get(); }
      else if (f.kind == TFELCU_FACTOR_PROGRAM) { o.ops = f.prog.data(); o.n_ops = static_cast<int>(f.prog.size()); }
      else if (f.kind == TFELCU_FACTOR_TABLE) {   // evaluated on the device by a user functor (tfel_device.cuh)
        if (f.table_quad != quad_id) throw std::string("integrate: a device function was tabulated for another quadrature formula");
        o.data = static_cast<const double*>(f.device_table.get());
      }
      else {   // free_function (expression.hpp:31-58): f(x_q) for every cell and quadrature point
        if (xq.empty()) {
          ck(tfelcu_quadrature_size(quad_id, &nq, &dim));
          ck(tfelcu_mesh_info(mesh, nullptr, nullptr, &K));
          xq.resize(K * nq * dim);
          ck(tfelcu_mesh_quadrature_points(mesh, quad_id, xq.data()));
        }
        tables.emplace_back(K * nq);
        std::vector<double>& t(tables.back());
        for (std::size_t p(0); p < K * nq; ++p) t[p] = f.fn(&xq[p * dim]);
        note_host_callbacks(K * nq);
        o.kind = TFELCU_FACTOR_TABLE; o.data = t.data();
      }
    }
    for (const monomial& m : e.terms) {
      // terms with a zero constant are kept: the reference adds their (zero) contributions too
      if (m.factors.size() > 5) throw std::string("integrand: more than 5 coefficient factors in one product");
      tfelcu_term t;
      t.test_comp = m.test_comp; t.test_d = m.test_d; t.trial_comp = m.trial_comp; t.trial_d = m.trial_d;
      t.c = m.c; t.n_factors = static_cast<int>(m.factors.size());
      for (int k(0); k < 5; ++k) t.factor[k] = k < t.n_factors ? remap[m.factors[k]] : 0;
      terms.push_back(t);
    }
  }
  std::vector<tfelcu_factor> factors;
  std::vector<tfelcu_term> terms;
private:
  std::vector<std::vector<double>> tables;
};

}  // namespace tfel_cuda

template <typename quad_t, bool a, typename mesh_t>
integration_proxy<true> integrate(const expression<a>& e, const mesh_t& m,
                                  typename std::enable_if<a, int>::type = 0) {
  return integration_proxy<true>{quad_t::id, e, m.handle()};
}

// rank-0 integral: sum_k |K| sum_q w_q f(x_q) (form.hpp:87-125)
template <typename quad_t, bool a, typename mesh_t>
double integrate(const expression<a>& e, const mesh_t& m, typename std::enable_if<!a, int>::type = 0) {
  tfel_cuda::lowered L(e, quad_t::id, m.handle());
  double result(0.0);
  tfel_cuda::ck(tfelcu_integrate(tfel_cuda::context(), m.handle(), quad_t::id, L.factors.data(), static_cast<int>(L.factors.size()),
                                 L.terms.data(), static_cast<int>(L.terms.size()), &result));
  return result;
}
template <typename quad_t, typename mesh_t>
double integrate(double (*f)(const double*), const mesh_t& m) { return integrate<quad_t>(make_expr(f), m); }

// ---- sparse matrix and the solver plugin interface (solver.hpp:29-40, 236-283) ----------------------------------------
class sparse_matrix {
public:
  // host CRS exactly as solver.cpp:71-96 hands it to PETSc: rows ascending, columns ascending, explicit zeros kept
  std::vector<int> row, col;
  std::vector<double> val;
  std::size_t n_row = 0, n_col = 0;
  tfelcu_matrix* device = nullptr;     // set when the matrix also lives on the device (zero-copy set_operator)
  std::size_t get_row_number() const { return n_row; }
  std::size_t get_column_number() const { return n_col; }
};

namespace solver {

class basic_solver {   // solver.hpp:29-40
public:
  virtual ~basic_solver() {}
  virtual void set_operator(const sparse_matrix& m) = 0;
  virtual bool solve(const array<double>& rhs, array<double>& x, dictionary& report) = 0;
};

namespace cuda {

// Krylov solvers on the B200, same dictionary keys as solver::petsc::gmres_ilu (solver.hpp:125-133):
// maxits, restart, rtol, atol, dtol, ilufill. ilufill is the level of fill of the incomplete factorisation
// (PCFactorSetLevels, solver.hpp:163) and is honoured by the ILU-preconditioned solvers; fixed_ilu_levels >= 0 pins it
// (gmres_ilu0) whatever the dictionary says.
class krylov : public basic_solver {
public:
  krylov(const dictionary& params, int method, int precond, int block_size = 0, int fixed_ilu_levels = -1) {
    if (not params.keys_exist({"maxits", "restart", "rtol", "atol", "dtol", "ilufill"}))
      throw std::string("solver::cuda: missing key(s) in the dictionary");   // solver.hpp:125-133
    tfelcu_solver_params p;
    p.method = method; p.precond = precond;
    p.maxits = params.get<unsigned>("maxits"); p.restart = params.get<unsigned>("restart");
    p.rtol = params.get<double>("rtol"); p.atol = params.get<double>("atol"); p.dtol = params.get<double>("dtol");
    p.block_size = block_size; p.check_every = 0; p.spmv_kernel = TFELCU_SPMV_AUTO; p.profile_every = 0;
    p.ilu_levels = 0;
    if (precond == TFELCU_PC_ILU)
      p.ilu_levels = fixed_ilu_levels >= 0 ? fixed_ilu_levels : static_cast<int>(params.get<unsigned>("ilufill"));
    p.pad = 0;
    tfel_cuda::ck(tfelcu_solver_create(tfel_cuda::context(), &p, &h));
  }
  krylov(const krylov&) = delete;
  ~krylov() { if (h) tfelcu_solver_destroy(h); }

  void set_operator(const sparse_matrix& m) {
    n = m.get_row_number();
    if (m.device) tfel_cuda::ck(tfelcu_solver_set_operator(h, m.device));
    else tfel_cuda::ck(tfelcu_solver_set_operator_csr(h, n, m.row.data(), m.col.data(), m.val.data()));
  }
  bool solve(const array<double>& rhs, array<double>& x, dictionary& report) {
    tfelcu_solver_report r;
    tfel_cuda::ck(tfelcu_solver_solve(h, rhs.get_data(), x.get_data(), &r));
    fill(report, r);
    return r.converged == 1;
  }
  tfelcu_solver* handle() const { return h; }
  static void fill(dictionary& report, const tfelcu_solver_report& r) {
    report.set("its", static_cast<unsigned>(r.its)).set("rnorm", r.rnorm).set("bnorm", r.bnorm)
          .set("converged", r.converged == 1).set("reason", static_cast<int>(r.converged))
          .set("t_setup_s", r.t_setup_s).set("t_solve_s", r.t_solve_s)
          .set("t_precond_setup_s", r.t_precond_setup_s).set("t_symbolic_s", r.t_symbolic_s)
          .set("restart_used", static_cast<unsigned>(r.restart_used))
          .set("ilufill_used", static_cast<int>(r.ilu_levels)).set("ilu_nnz", static_cast<double>(r.ilu_nnz))
          .set("ilu_zero_pivot_row", static_cast<int>(r.ilu_zero_pivot_row));
  }
private:
  tfelcu_solver* h = nullptr;
  std::size_t n = 0;
};

struct cg_jacobi : krylov { explicit cg_jacobi(const dictionary& p) : krylov(p, TFELCU_METHOD_CG, TFELCU_PC_JACOBI) {} };
struct cg_block_jacobi : krylov {
  explicit cg_block_jacobi(const dictionary& p, int block_size = 4) : krylov(p, TFELCU_METHOD_CG, TFELCU_PC_BLOCK_JACOBI, block_size) {}
};
// GMRES(restart) with modified Gram-Schmidt, left-preconditioned by ILU(ilufill) in natural ordering: the device
// counterpart of solver::petsc::gmres_ilu (solver.hpp:139-163)
struct gmres_ilu : krylov { explicit gmres_ilu(const dictionary& p) : krylov(p, TFELCU_METHOD_GMRES, TFELCU_PC_ILU) {} };
// the same with the level of fill pinned to 0 whatever "ilufill" says
struct gmres_ilu0 : krylov { explicit gmres_ilu0(const dictionary& p) : krylov(p, TFELCU_METHOD_GMRES, TFELCU_PC_ILU, 0, 0) {} };
struct cg_ilu : krylov { explicit cg_ilu(const dictionary& p) : krylov(p, TFELCU_METHOD_CG, TFELCU_PC_ILU) {} };
struct gmres_jacobi : krylov { explicit gmres_jacobi(const dictionary& p) : krylov(p, TFELCU_METHOD_GMRES, TFELCU_PC_JACOBI) {} };

}  // namespace cuda

#ifdef TFEL_CUDA_AS_PETSC
// Opt-in (define TFEL_CUDA_AS_PETSC before including this header): programs written against the reference keep their
// `solver::petsc::gmres_ilu s(p);` line (solver.hpp:122-210) and get the device GMRES(restart) with modified Gram-Schmidt
// and ILU(ilufill) -- same required dictionary keys, so a program that passes "abstol" instead of "atol" throws here
// exactly as it does in the reference, and the same level of fill as PCFactorSetLevels(pc, ilufill).
namespace petsc { using gmres_ilu = cuda::gmres_ilu; }
#endif
}  // namespace solver

// ---- forms (bilinear_form.hpp, linear_form.hpp, composite_*_form.hpp) -------------------------------------------------
template <typename fes_t>
class linear_form {
public:
  explicit linear_form(const fes_t& f) : fes(&f) {
    std::vector<tfelcu_fes*> hs;
    tfel_cuda::space_traits<fes_t>::handles(f, hs);
    tfel_cuda::ck(tfelcu_vector_create(tfel_cuda::context(), static_cast<int>(hs.size()), hs.data(), &h));
  }
  linear_form(const linear_form&) = delete;
  ~linear_form() { if (h) tfelcu_vector_destroy(h); }

  expression<true> get_test_function() const { return test(0); }
  template <std::size_t n> expression<true> get_test_function() const { return test(static_cast<int>(n)); }

  void operator+=(const integration_proxy<true>& p) {   // linear_form.hpp:20-142
    for (const auto& m : p.integrand.terms)
      if (m.test_comp < 0 or m.trial_comp >= 0) throw std::string("linear_form: every term needs exactly one test function");
    tfel_cuda::lowered L(p.integrand, p.quad_id, p.mesh);
    tfel_cuda::ck(tfelcu_vector_assemble(h, p.quad_id, L.factors.data(), static_cast<int>(L.factors.size()),
                                         L.terms.data(), static_cast<int>(L.terms.size())));
  }
  void clear() { tfel_cuda::ck(tfelcu_vector_clear(h)); }
  array<double> get_coefficients() const {
    std::size_t n(0); tfel_cuda::ck(tfelcu_vector_size(h, &n));
    array<double> c{n};
    tfel_cuda::ck(tfelcu_vector_download(h, c.get_data()));
    return c;
  }
  tfelcu_vector* handle() const { return h; }
private:
  static expression<true> test(int comp) {
    expression<true> e; tfel_cuda::monomial m; m.test_comp = comp; e.terms.push_back(m); return e;
  }
  const fes_t* fes;
  tfelcu_vector* h = nullptr;
};

template <typename test_fes_t, typename trial_fes_t>
class bilinear_form {
public:
  // sizes the matrix, builds the CSR pattern and performs clear() with the Dirichlet dofs registered so far
  // (bilinear_form.hpp:9-19): register boundary conditions before constructing the form
  bilinear_form(const test_fes_t& te, const trial_fes_t& tr) : test_fes(&te), trial_fes(&tr) {
    std::vector<tfelcu_fes*> a, b;
    tfel_cuda::space_traits<test_fes_t>::handles(te, a);
    tfel_cuda::space_traits<trial_fes_t>::handles(tr, b);
    tfel_cuda::ck(tfelcu_matrix_create(tfel_cuda::context(), static_cast<int>(a.size()), a.data(),
                                       static_cast<int>(b.size()), b.data(), &h));
  }
  bilinear_form(const bilinear_form&) = delete;
  ~bilinear_form() { if (h) tfelcu_matrix_destroy(h); }

  expression<true> get_test_function() const { return placeholder(0, true); }
  expression<true> get_trial_function() const { return placeholder(0, false); }
  template <std::size_t n> expression<true> get_test_function() const { return placeholder(static_cast<int>(n), true); }
  template <std::size_t n> expression<true> get_trial_function() const { return placeholder(static_cast<int>(n), false); }

  void operator+=(const integration_proxy<true>& p) {   // bilinear_form.hpp:23-110, composite_bilinear_form.hpp:123-184
    for (const auto& m : p.integrand.terms)
      if (m.test_comp < 0 or m.trial_comp < 0) throw std::string("bilinear_form: every term needs one test and one trial function");
    tfel_cuda::lowered L(p.integrand, p.quad_id, p.mesh);
    tfel_cuda::ck(tfelcu_matrix_assemble(h, p.quad_id, L.factors.data(), static_cast<int>(L.factors.size()),
                                         L.terms.data(), static_cast<int>(L.terms.size())));
  }
  void clear() { tfel_cuda::ck(tfelcu_matrix_clear(h)); }   // bilinear_form.hpp:159-165

  // the matrix as the reference would hand it to a solver plugin (solver.cpp:71-96)
  sparse_matrix get_matrix(bool with_host_copy = true) const {
    sparse_matrix m;
    std::size_t nnz(0);
    tfel_cuda::ck(tfelcu_matrix_info(h, &m.n_row, &m.n_col, &nnz));
    m.device = h;
    if (with_host_copy) {
      m.row.resize(m.n_row + 1); m.col.resize(nnz); m.val.resize(nnz);
      tfel_cuda::ck(tfelcu_matrix_download_csr(h, m.row.data(), m.col.data(), m.val.data()));
    }
    return m;
  }

  // bilinear_form.hpp:122-157: f <- b with the Dirichlet rows replaced by their values; s.set_operator(a); s.solve(f, x)
  typename trial_fes_t::element solve(const linear_form<test_fes_t>& form, solver::basic_solver& s, dictionary* report = nullptr) const {
    std::size_t n_row(0), n_col(0);
    tfel_cuda::ck(tfelcu_matrix_info(h, &n_row, &n_col, nullptr));
    dictionary local;
    dictionary& rep(report ? *report : local);
    solver::cuda::krylov* device_solver(dynamic_cast<solver::cuda::krylov*>(&s));
    if (device_solver) {   // everything stays on the device, the solution included
      tfelcu_solver_report r;
      std::shared_ptr<tfel_cuda::device_array> xd(std::make_shared<tfel_cuda::device_array>(n_col));
      tfel_cuda::ck(tfelcu_matrix_solve(h, form.handle(), device_solver->handle(), xd->p, &r));
      solver::cuda::krylov::fill(rep, r);
      return typename trial_fes_t::element(*trial_fes, tfel_cuda::coefficients(xd));
    }
    array<double> x{n_col};
    {                      // any other plugin: the reference's host protocol
      array<double> f{n_row};
      tfel_cuda::ck(tfelcu_matrix_make_rhs(h, form.handle(), f.get_data()));
      sparse_matrix m(get_matrix(true));
      m.device = nullptr;
      s.set_operator(m);
      s.solve(f, x, rep);
    }
    return typename trial_fes_t::element(*trial_fes, std::move(x));
  }
  tfelcu_matrix* handle() const { return h; }
private:
  static expression<true> placeholder(int comp, bool test) {
    expression<true> e; tfel_cuda::monomial m;
    if (test) m.test_comp = comp; else m.trial_comp = comp;
    e.terms.push_back(m); return e;
  }
  const test_fes_t* test_fes;
  const trial_fes_t* trial_fes;
  tfelcu_matrix* h = nullptr;
};

// ---- projectors (projector.hpp:12-97): callers on either side of the path (initial conditions, coefficient fields) ------
namespace projector {

// L2 projection: the mass-matrix solve through the assembly + solve path (projector.hpp:12-46), CG + Jacobi on device
template <typename fe_type, typename quadrature_type>
typename finite_element_space<fe_type>::element
l2(const expression<false>& expr, const finite_element_space<fe_type>& fes) {
  typedef finite_element_space<fe_type> fes_type;
  bilinear_form<fes_type, fes_type> a(fes, fes); {
    auto u(a.get_trial_function());
    auto v(a.get_test_function());
    a += integrate<quadrature_type>(u * v, fes.get_mesh());
  }
  linear_form<fes_type> f(fes); {
    auto v(f.get_test_function());
    f += integrate<quadrature_type>(expr * v, fes.get_mesh());
  }
  dictionary p(dictionary()
               .set("maxits",  2000u)
               .set("restart", 1000u)
               .set("rtol",    1.e-8)
               .set("atol",    1.e-50)
               .set("dtol",    1.e20)
               .set("ilufill", 2u));
  solver::cuda::cg_jacobi s(p);
  return a.solve(f, s);
}

template <typename fe_type, typename quadrature_type>
typename finite_element_space<fe_type>::element
l2(const std::function<double(const double*)>& fun, const finite_element_space<fe_type>& fes) {
  return l2<fe_type, quadrature_type>(make_expr(fun), fes);
}

// nodal interpolation (projector.hpp:48-63): the function at the space coordinate of every dof (fes.hpp:190-202)
template <typename fe_type>
typename finite_element_space<fe_type>::element
lagrange(const std::function<double(const double*)>& fun, const finite_element_space<fe_type>& fes) {
  const std::size_t n(fes.get_dof_number());
  const std::size_t dim(fe_type::cell_type::dim);
  std::vector<unsigned int> dofs(n);
  for (std::size_t i(0); i < n; ++i) dofs[i] = static_cast<unsigned int>(i);
  std::vector<double> x(n * dim);
  tfel_cuda::ck(tfelcu_fes_dof_coordinates(fes.handle(), dofs.data(), n, x.data()));
  array<double> coefficients{n};
  for (std::size_t i(0); i < n; ++i) coefficients.at(i) = fun(&x[i * dim]);
  return typename finite_element_space<fe_type>::element(fes, coefficients);
}

}  // namespace projector

// the output side (mesh_data, to_mesh_vertex_data / to_mesh_cell_data, exporter::*): part of the umbrella header like
// the reference's export.hpp / mesh_data.hpp
#include "tfel_export.hpp"

#endif  // TFEL_TFEL_HPP
