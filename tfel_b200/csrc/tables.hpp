// Reference-element data of the product: Lagrange basis functions and quadrature formulas.
// Host-side tabulation; kernels receive the small tables and stage them in shared memory.
//
// Reference: src/core/fe.hpp (triangle P1 :260-317, P1 bubble :319-387, P2 :389-457;
// tetrahedron P1 :494-562, P1 bubble :564-654), src/core/quadrature.cpp:52-367.
// The tetrahedron P2 element does not exist in the reference (cell.hpp:948-952); it is defined
// here with the reference's conventions: vertex dofs first, then one dof per edge in the local
// edge order (01)(02)(03)(12)(13)(23) of cell.hpp:646-663.
#pragma once

#include <cmath>
#include <vector>

#include "common.cuh"

namespace tfelcu {

constexpr int kMaxDim = 3;
constexpr int kMaxNd = 10;   // dofs per cell (tet P2)
constexpr int kMaxNq = 56;   // qfSym56pTet

// local vertices of edge j (cell.hpp:362-373 for triangles, :646-663 for tetrahedra)
inline void local_edge_vertices(int dim, int j, int& a, int& b) {
  static const int tri[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  static const int tet[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  if (dim == 2) { a = tri[j][0]; b = tri[j][1]; }
  else { a = tet[j][0]; b = tet[j][1]; }
}

inline int n_sub_entities(int dim, int sd) {   // cell.hpp:341-344, 625-628
  if (sd == 0) return dim + 1;
  if (sd == dim) return 1;
  if (sd == 1) return dim == 2 ? 3 : 6;
  return 4;  // triangles of a tetrahedron
}

inline int fe_ndof(int dim, int fe) {
  const int nv = dim + 1, ne = n_sub_entities(dim, 1);
  switch (fe) {
    case TFELCU_FE_P1: return nv;
    case TFELCU_FE_P1_BUBBLE: return nv + 1;
    case TFELCU_FE_P2: return nv + ne;
  }
  fail("unknown finite element id");
}

// dofs attached to one entity of kind sd (n_dof[] of fe.hpp)
inline int fe_dofs_on(int dim, int fe, int sd) {
  if (sd == 0) return 1;
  if (fe == TFELCU_FE_P1_BUBBLE && sd == dim) return 1;
  if (fe == TFELCU_FE_P2 && sd == 1) return 1;
  return 0;
}

// value and reference gradient of every basis function at a reference point:
// out[i * (dim + 1) + 0] = phi_i, out[i * (dim + 1) + 1 + s] = d phi_i / d xhat_s
inline void fe_eval_hat(int dim, int fe, const double* xh, double* out) {
  const int nv = dim + 1, w = dim + 1;
  double L[4], dL[4][3];
  L[0] = 1.0;
  for (int s = 0; s < dim; ++s) { L[0] -= xh[s]; L[1 + s] = xh[s]; }
  for (int v = 0; v < nv; ++v)
    for (int s = 0; s < dim; ++s) dL[v][s] = v == 0 ? -1.0 : (v == s + 1 ? 1.0 : 0.0);
  if (fe == TFELCU_FE_P2) {
    for (int v = 0; v < nv; ++v) {
      out[v * w] = L[v] * (2.0 * L[v] - 1.0);
      for (int s = 0; s < dim; ++s) out[v * w + 1 + s] = (4.0 * L[v] - 1.0) * dL[v][s];
    }
    const int ne = n_sub_entities(dim, 1);
    for (int j = 0; j < ne; ++j) {
      int a, b; local_edge_vertices(dim, j, a, b);
      double* o = out + (nv + j) * w;
      o[0] = 4.0 * L[a] * L[b];
      for (int s = 0; s < dim; ++s) o[1 + s] = 4.0 * (dL[a][s] * L[b] + L[a] * dL[b][s]);
    }
    return;
  }
  for (int v = 0; v < nv; ++v) {
    out[v * w] = L[v];
    for (int s = 0; s < dim; ++s) out[v * w + 1 + s] = dL[v][s];
  }
  if (fe == TFELCU_FE_P1_BUBBLE) {
    const double scale = dim == 2 ? 27.0 : 256.0;   // fe.hpp:380, :645
    double* o = out + nv * w;
    double all = 1.0;
    for (int v = 0; v < nv; ++v) all *= L[v];
    o[0] = scale * all;
    for (int s = 0; s < dim; ++s) {
      double sum = 0.0;
      for (int v = 0; v < nv; ++v) {
        double term = dL[v][s];
        for (int u = 0; u < nv; ++u) if (u != v) term *= L[u];
        sum += term;
      }
      o[1 + s] = scale * sum;
    }
  }
}

inline void fe_nodes(int dim, int fe, double* x /* [nd][dim] */) {
  const int nv = dim + 1, nd = fe_ndof(dim, fe);
  for (int i = 0; i < nd * dim; ++i) x[i] = 0.0;
  for (int v = 1; v < nv; ++v) x[v * dim + v - 1] = 1.0;
  if (fe == TFELCU_FE_P1_BUBBLE)
    for (int s = 0; s < dim; ++s) x[nv * dim + s] = 1.0 / nv;
  if (fe == TFELCU_FE_P2)
    for (int j = 0; j < n_sub_entities(dim, 1); ++j) {
      int a, b; local_edge_vertices(dim, j, a, b);
      for (int s = 0; s < dim; ++s) x[(nv + j) * dim + s] = 0.5 * (x[a * dim + s] + x[b * dim + s]);
    }
}

struct Quadrature {
  int dim = 0, nq = 0;
  std::vector<double> x, w;   // x[nq][dim], weights sum to 1 (cell volume applied separately)
};

namespace detail {
// barycentric orbit of a symmetric tetrahedron rule: multiset {b0, b1, b2, b3} with multiplicities
// given by kind (1: centroid, 4: {a,a,a,b}, 6: {a,a,b,b}, 12: {a,a,b,c}); weight per point
struct Orbit { int kind; double a, b, c, w; };

inline void expand(const Orbit* orbits, int n, Quadrature& q) {
  q.dim = 3;
  for (int o = 0; o < n; ++o) {
    const Orbit& r = orbits[o];
    double m[4];
    switch (r.kind) {
      case 1: m[0] = m[1] = m[2] = m[3] = r.a; break;
      case 4: m[0] = m[1] = m[2] = r.a; m[3] = r.b; break;
      case 6: m[0] = m[1] = r.a; m[2] = m[3] = r.b; break;
      default: m[0] = m[1] = r.a; m[2] = r.b; m[3] = r.c; break;
    }
    std::vector<double> pts;
    for (int p = 0; p < 24; ++p) {             // all orderings of 4 slots, duplicates dropped
      int idx[4] = {0, 1, 2, 3};
      int code = p;
      for (int s = 0; s < 3; ++s) {            // factorial number system
        const int radix = 4 - s, pick = code % radix; code /= radix;
        const int t = idx[s + pick];
        for (int u = s + pick; u > s; --u) idx[u] = idx[u - 1];
        idx[s] = t;
      }
      const double c0 = m[idx[0]], c1 = m[idx[1]], c2 = m[idx[2]], c3 = m[idx[3]];
      bool dup = false;
      for (size_t e = 0; e < pts.size(); e += 4)
        dup = dup || (pts[e] == c0 && pts[e + 1] == c1 && pts[e + 2] == c2 && pts[e + 3] == c3);
      if (dup) continue;
      pts.insert(pts.end(), {c0, c1, c2, c3});
      q.x.insert(q.x.end(), {c0, c1, c2});
      q.w.push_back(r.w);
    }
  }
  q.nq = static_cast<int>(q.w.size());
}
}  // namespace detail

inline Quadrature make_quadrature(int id) {
  using detail::Orbit;
  Quadrature q;
  auto tri = [&](std::initializer_list<double> xs, std::initializer_list<double> ws) {
    q.dim = 2; q.x.assign(xs); q.w.assign(ws); q.nq = static_cast<int>(q.w.size());
  };
  switch (id) {
    case TFELCU_QUAD_TRI_QF1PT: tri({1.0 / 3.0, 1.0 / 3.0}, {1.0}); break;
    case TFELCU_QUAD_TRI_QF2PT:
      tri({0.5, 0.5, 0.5, 0.0, 0.0, 0.5}, {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0}); break;
    case TFELCU_QUAD_TRI_QF1PTLUMP:
      tri({0.0, 0.0, 1.0, 0.0, 0.0, 1.0}, {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0}); break;
    case TFELCU_QUAD_TRI_QF5PT: {
      // degree-5 seven point rule (quadrature.cpp:62-76 evaluates the same closed forms at run time)
      const double r = std::sqrt(15.0);
      const double lo = (6.0 - r) / 21.0, lo2 = (9.0 + 2.0 * r) / 21.0;
      const double hi = (6.0 + r) / 21.0, hi2 = (9.0 - 2.0 * r) / 21.0;
      const double wl = (155.0 - r) / 1200.0, wh = (155.0 + r) / 1200.0;
      tri({1.0 / 3.0, 1.0 / 3.0, lo, lo, lo, lo2, lo2, lo, hi, hi, hi, hi2, hi2, hi},
          {0.225, wl, wl, wl, wh, wh, wh});
      break;
    }
    case TFELCU_QUAD_TET_QF1PTET:
    case TFELCU_QUAD_TET_SYM1: { const Orbit o[] = {{1, 0.25, 0, 0, 1.0}}; detail::expand(o, 1, q); break; }
    case TFELCU_QUAD_TET_QF5PTET: {
      q.dim = 3; q.nq = 5;
      q.x = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.25, 0.25, 0.25};
      q.w = {1.0 / 20.0, 1.0 / 20.0, 1.0 / 20.0, 1.0 / 20.0, 16.0 / 20.0};
      break;
    }
    case TFELCU_QUAD_TET_QF4PTET: {
      const Orbit o[] = {{4, 0.138196601125011, 0.585410196624969, 0, 0.25}}; detail::expand(o, 1, q); break;
    }
    case TFELCU_QUAD_TET_SYM4: {
      const Orbit o[] = {{4, 0.1381966011250110, 0.5854101966249680, 0, 0.25}}; detail::expand(o, 1, q); break;
    }
    case TFELCU_QUAD_TET_SYM10: {
      const Orbit o[] = {{4, 0.0738349017262234, 0.7784952948213300, 0, 0.0476331348432089},
                         {6, 0.0937556561159491, 0.4062443438840510, 0, 0.1349112434378610}};
      detail::expand(o, 2, q); break;
    }
    case TFELCU_QUAD_TET_SYM20: {
      const Orbit o[] = {{4, 0.0323525947272439, 0.9029422158182680, 0, 0.0070670747944695},
                         {12, 0.0603604415251421, 0.2626825838877790, 0.6165965330619370, 0.0469986689718877},
                         {4, 0.3097693042728620, 0.0706920871814129, 0, 0.1019369182898680}};
      detail::expand(o, 3, q); break;
    }
    case TFELCU_QUAD_TET_SYM35: {
      const Orbit o[] = {{4, 0.0267367755543735, 0.9197896733368800, 0, 0.0021900463965388},
                         {12, 0.0391022406356488, 0.1740356302468940, 0.7477598884818090, 0.0143395670177665},
                         {6, 0.0452454000155172, 0.4547545999844830, 0, 0.0250305395686746},
                         {12, 0.2232010379623150, 0.0504792790607720, 0.5031186450145980, 0.0479839333057554},
                         {1, 0.25, 0, 0, 0.0931745731195340}};
      detail::expand(o, 5, q); break;
    }
    case TFELCU_QUAD_TET_SYM56: {
      const Orbit o[] = {{4, 0.0149520651530592, 0.9551438045408220, 0, 0.0010373112336140},
                         {12, 0.0340960211962615, 0.1518319491659370, 0.7799760084415400, 0.0096016645399480},
                         {12, 0.0462051504150017, 0.3549340560639790, 0.5526556431060170, 0.0164493976798232},
                         {12, 0.2281904610687610, 0.0055147549744775, 0.5381043228880020, 0.0153747766513310},
                         {12, 0.3523052600879940, 0.0992057202494530, 0.1961837595745600, 0.0293520118375230},
                         {4, 0.1344783347929940, 0.5965649956210170, 0, 0.0366291366405108}};
      detail::expand(o, 6, q); break;
    }
    default: fail("unknown quadrature id");
  }
  return q;
}

}  // namespace tfelcu
