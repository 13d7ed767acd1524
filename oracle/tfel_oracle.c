/* TEST INFRASTRUCTURE (oracle) -- see tfel_oracle.h for the rules on who may use this.
 *
 * Plain-C restatement of the reference's (thomashilke/tfel) assembly + solve path.
 * Every function cites the reference file:line it follows. Compiled with
 * -ffp-contract=off so that the arithmetic is the reference's (x86-64, no FMA).
 */
#include "tfel_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------- */
/* Finite elements: src/core/fe.hpp                                           */
/* ------------------------------------------------------------------------- */

int orc_fe_ndof(int dim, int fe) {
  /* fe.hpp:263 (tri P1: 3), :337 (tri P1 bubble: 4), :392 (tri P2: 6),
   * :497 (tet P1: 4), :588 (tet P1 bubble: 5); tet P2 (new): 4 vertices + 6 edges */
  if (dim == 2) return fe == ORC_FE_P1 ? 3 : fe == ORC_FE_P1_BUBBLE ? 4 : fe == ORC_FE_P2 ? 6 : -1;
  if (dim == 3) return fe == ORC_FE_P1 ? 4 : fe == ORC_FE_P1_BUBBLE ? 5 : fe == ORC_FE_P2 ? 10 : -1;
  return -1;
}

void orc_fe_dofs_per_entity(int dim, int fe, int* out) {
  /* n_dof[] arrays: fe.hpp:264,338,393,498,589 */
  for (int s = 0; s <= dim; ++s) out[s] = 0;
  out[0] = 1;
  if (fe == ORC_FE_P1_BUBBLE) out[dim] = 1;
  if (fe == ORC_FE_P2) out[1] = 1;
}

/* local edge j of a simplex -> its two local vertices
 * triangle: cell.hpp:362-373; tetrahedron: cell.hpp:646-663 */
static void local_edge(int dim, int j, int* a, int* b) {
  if (dim == 2) {
    if (j < 2) { *a = 0; *b = j + 1; } else { *a = 1; *b = 2; }
  } else {
    if (j < 3) { *a = 0; *b = 1 + j; }
    else if (j < 5) { *a = 1; *b = j - 1; }
    else { *a = 2; *b = 3; }
  }
}

/* local triangle j of a tetrahedron: cell.hpp:665-681 */
static void local_triangle(int j, int* v) {
  if (j < 2) { v[0] = 0; v[1] = 1; v[2] = 2 + j; }
  else { v[0] = j - 2; v[1] = 2; v[2] = 3; }
}

static int n_entities(int dim, int sd) {
  /* n_subdomain(): cell.hpp:341-344 (3,3,1), :625-628 (4,6,4,1) */
  static const int tri[3] = {3, 3, 1};
  static const int tet[4] = {4, 6, 4, 1};
  return dim == 2 ? tri[sd] : tet[sd];
}

void orc_fe_nodes(int dim, int fe, double* nodes) {
  /* x[][] arrays: fe.hpp:271-273, 345-348, 401-402, 505-508, 596-600 */
  const int nd = orc_fe_ndof(dim, fe);
  memset(nodes, 0, sizeof(double) * (size_t)nd * dim);
  for (int i = 1; i <= dim; ++i) nodes[i * dim + (i - 1)] = 1.0;
  if (fe == ORC_FE_P1_BUBBLE)
    for (int s = 0; s < dim; ++s) nodes[(dim + 1) * dim + s] = 1.0 / (dim + 1);
  if (fe == ORC_FE_P2) {
    const int ne = n_entities(dim, 1);
    for (int j = 0; j < ne; ++j) {
      int a, b; local_edge(dim, j, &a, &b);
      for (int s = 0; s < dim; ++s)
        nodes[(dim + 1 + j) * dim + s] = 0.5 * (nodes[a * dim + s] + nodes[b * dim + s]);
    }
  }
}

static void fe_eval(int dim, int fe, const double* x, double* out /* [nd][dim+1] */) {
  const int w = dim + 1;
  double lam[4], dlam[4][3];
  lam[0] = 1.0;
  for (int s = 0; s < dim; ++s) { lam[0] -= x[s]; lam[s + 1] = x[s]; }
  for (int i = 0; i <= dim; ++i)
    for (int s = 0; s < dim; ++s)
      dlam[i][s] = (i == 0) ? -1.0 : (i - 1 == s ? 1.0 : 0.0);

  if (fe == ORC_FE_P1 || fe == ORC_FE_P1_BUBBLE) {
    /* fe.hpp:306-316 (tri), :547-562 (tet) */
    for (int i = 0; i <= dim; ++i) {
      out[i * w] = lam[i];
      for (int s = 0; s < dim; ++s) out[i * w + 1 + s] = dlam[i][s];
    }
    if (fe == ORC_FE_P1_BUBBLE) {
      /* fe.hpp:380-386 (27 l0 l1 l2), :645-653 (256 l0 l1 l2 l3) */
      const double scale = dim == 2 ? 27.0 : 256.0;
      double prod = 1.0;
      for (int i = 0; i <= dim; ++i) prod *= lam[i];
      out[(dim + 1) * w] = scale * prod;
      for (int s = 0; s < dim; ++s) {
        double acc = 0.0;
        for (int i = 0; i <= dim; ++i) {
          double t = dlam[i][s];
          for (int l = 0; l <= dim; ++l) if (l != i) t *= lam[l];
          acc += t;
        }
        out[(dim + 1) * w + 1 + s] = scale * acc;
      }
    }
    return;
  }
  /* P2: vertex functions l_i (2 l_i - 1), edge functions 4 l_a l_b
   * (triangle: fe.hpp:436-456 written in x, y; tetrahedron: new element) */
  for (int i = 0; i <= dim; ++i) {
    out[i * w] = lam[i] * (2.0 * lam[i] - 1.0);
    for (int s = 0; s < dim; ++s) out[i * w + 1 + s] = (4.0 * lam[i] - 1.0) * dlam[i][s];
  }
  const int ne = n_entities(dim, 1);
  for (int j = 0; j < ne; ++j) {
    int a, b; local_edge(dim, j, &a, &b);
    const int r = dim + 1 + j;
    out[r * w] = 4.0 * lam[a] * lam[b];
    for (int s = 0; s < dim; ++s)
      out[r * w + 1 + s] = 4.0 * (dlam[a][s] * lam[b] + lam[a] * dlam[b][s]);
  }
}

void orc_fe_tabulate(int dim, int fe, int npts, const double* pts, double* out) {
  /* fe_value_manager.hpp:135-157 (values_hat[q][i][0..dim]) */
  const int nd = orc_fe_ndof(dim, fe);
  for (int q = 0; q < npts; ++q)
    fe_eval(dim, fe, pts + (size_t)q * dim, out + (size_t)q * nd * (dim + 1));
}

/* ------------------------------------------------------------------------- */
/* Quadrature: src/core/quadrature.cpp                                        */
/* ------------------------------------------------------------------------- */

typedef struct { int kind; double a, b, c, w; } tet_orbit;
/* kind 1: centroid; 4: {a,a,a,b}; 6: {a,a,b,b}; 12: {a,a,b,c}. Literals are the ones the
 * reference lists (quadrature.cpp:117-367, symmetric rules); expansion order is ours. */
static const tet_orbit sym4[] = {{4, 0.1381966011250110, 0.5854101966249680, 0, 1.0 / 4.0}};
static const tet_orbit sym10[] = {
  {4, 0.0738349017262234, 0.7784952948213300, 0, 0.0476331348432089},
  {6, 0.0937556561159491, 0.4062443438840510, 0, 0.1349112434378610}};
static const tet_orbit sym20[] = {
  {4, 0.0323525947272439, 0.9029422158182680, 0, 0.0070670747944695},
  {12, 0.0603604415251421, 0.2626825838877790, 0.6165965330619370, 0.0469986689718877},
  {4, 0.3097693042728620, 0.0706920871814129, 0, 0.1019369182898680}};
static const tet_orbit sym35[] = {
  {4, 0.0267367755543735, 0.9197896733368800, 0, 0.0021900463965388},
  {12, 0.0391022406356488, 0.1740356302468940, 0.7477598884818090, 0.0143395670177665},
  {6, 0.0452454000155172, 0.4547545999844830, 0, 0.0250305395686746},
  {12, 0.2232010379623150, 0.0504792790607720, 0.5031186450145980, 0.0479839333057554},
  {1, 0.25, 0, 0, 0.0931745731195340}};
static const tet_orbit sym56[] = {
  {4, 0.0149520651530592, 0.9551438045408220, 0, 0.0010373112336140},
  {12, 0.0340960211962615, 0.1518319491659370, 0.7799760084415400, 0.0096016645399480},
  {12, 0.0462051504150017, 0.3549340560639790, 0.5526556431060170, 0.0164493976798232},
  {12, 0.2281904610687610, 0.0055147549744775, 0.5381043228880020, 0.0153747766513310},
  {12, 0.3523052600879940, 0.0992057202494530, 0.1961837595745600, 0.0293520118375230},
  {4, 0.1344783347929940, 0.5965649956210170, 0, 0.0366291366405108}};

static int expand_tet_orbits(const tet_orbit* orb, int n_orb, double* x, double* w) {
  int n = 0;
  for (int o = 0; o < n_orb; ++o) {
    double t[4];
    const tet_orbit* r = &orb[o];
    if (r->kind == 1) { t[0] = t[1] = t[2] = t[3] = r->a; }
    else if (r->kind == 4) { t[0] = t[1] = t[2] = r->a; t[3] = r->b; }
    else if (r->kind == 6) { t[0] = t[1] = r->a; t[2] = t[3] = r->b; }
    else { t[0] = t[1] = r->a; t[2] = r->b; t[3] = r->c; }
    /* all distinct arrangements of the multiset t over 4 slots; keep slots 0..2 */
    double seen[24][4]; int n_seen = 0;
    for (int p0 = 0; p0 < 4; ++p0) for (int p1 = 0; p1 < 4; ++p1) for (int p2 = 0; p2 < 4; ++p2) for (int p3 = 0; p3 < 4; ++p3) {
      if (p0 == p1 || p0 == p2 || p0 == p3 || p1 == p2 || p1 == p3 || p2 == p3) continue;
      const double cand[4] = {t[p0], t[p1], t[p2], t[p3]};
      int dup = 0;
      for (int s = 0; s < n_seen && !dup; ++s)
        dup = cand[0] == seen[s][0] && cand[1] == seen[s][1] && cand[2] == seen[s][2] && cand[3] == seen[s][3];
      if (dup) continue;
      memcpy(seen[n_seen++], cand, sizeof(cand));
      if (x) { x[n * 3] = cand[0]; x[n * 3 + 1] = cand[1]; x[n * 3 + 2] = cand[2]; w[n] = r->w; }
      ++n;
    }
  }
  return n;
}

int orc_quadrature(const char* name, double* x, double* w, int* dim) {
  if (!strcmp(name, "qf1pT")) {              /* quadrature.cpp:52-53 */
    if (dim) *dim = 2;
    if (x) { x[0] = 1.0 / 3.0; x[1] = 1.0 / 3.0; w[0] = 1.0; }
    return 1;
  }
  if (!strcmp(name, "qf2pT")) {              /* quadrature.cpp:56-59 */
    if (dim) *dim = 2;
    if (x) {
      const double p[6] = {0.5, 0.5, 0.5, 0.0, 0.0, 0.5};
      memcpy(x, p, sizeof(p)); w[0] = w[1] = w[2] = 1.0 / 3.0;
    }
    return 3;
  }
  if (!strcmp(name, "qf1pTlump")) {          /* quadrature.cpp:79-82 */
    if (dim) *dim = 2;
    if (x) {
      const double p[6] = {0.0, 0.0, 1.0, 0.0, 0.0, 1.0};
      memcpy(x, p, sizeof(p)); w[0] = w[1] = w[2] = 1.0 / 3.0;
    }
    return 3;
  }
  if (!strcmp(name, "qf5pT")) {              /* quadrature.cpp:62-76, runtime sqrt */
    if (dim) *dim = 2;
    if (x) {
      const double s = sqrt(15.0);
      const double a = (6.0 - s) / 21.0, A = (9.0 + 2.0 * s) / 21.0;
      const double b = (6.0 + s) / 21.0, B = (9.0 - 2.0 * s) / 21.0;
      const double p[14] = {1.0 / 3.0, 1.0 / 3.0, a, a, a, A, A, a, b, b, b, B, B, b};
      memcpy(x, p, sizeof(p));
      w[0] = 0.225;
      w[1] = w[2] = w[3] = (155.0 - s) / 1200.0;
      w[4] = w[5] = w[6] = (155.0 + s) / 1200.0;
    }
    return 7;
  }
  if (dim) *dim = 3;
  if (!strcmp(name, "qf1pTet") || !strcmp(name, "qfSym1pTet")) {  /* quadrature.cpp:88-89,113-114 */
    if (x) { x[0] = x[1] = x[2] = 0.25; w[0] = 1.0; }
    return 1;
  }
  if (!strcmp(name, "qf5pTet")) {            /* quadrature.cpp:99-108 */
    if (x) {
      memset(x, 0, sizeof(double) * 15);
      x[3] = 1.0; x[7] = 1.0; x[11] = 1.0; x[12] = x[13] = x[14] = 0.25;
      w[0] = w[1] = w[2] = w[3] = 1.0 / 20.0; w[4] = 16.0 / 20.0;
    }
    return 5;
  }
  if (!strcmp(name, "qf4pTet")) {            /* quadrature.cpp:92-96 */
    static const tet_orbit o[] = {{4, 0.138196601125011, 0.585410196624969, 0, 1.0 / 4.0}};
    return expand_tet_orbits(o, 1, x, w);
  }
  if (!strcmp(name, "qfSym4pTet")) return expand_tet_orbits(sym4, 1, x, w);
  if (!strcmp(name, "qfSym10pTet")) return expand_tet_orbits(sym10, 2, x, w);
  if (!strcmp(name, "qfSym20pTet")) return expand_tet_orbits(sym20, 3, x, w);
  if (!strcmp(name, "qfSym35pTet")) return expand_tet_orbits(sym35, 5, x, w);
  if (!strcmp(name, "qfSym56pTet")) return expand_tet_orbits(sym56, 6, x, w);
  return -1;
}

/* ------------------------------------------------------------------------- */
/* Meshes: src/core/mesh.cpp                                                  */
/* ------------------------------------------------------------------------- */

void orc_gen_square_mesh(double x1, double x2, unsigned n1, unsigned n2, double* vertices, unsigned* cells) {
  /* mesh.cpp:19-51 */
  for (unsigned j = 0; j <= n2; ++j)
    for (unsigned i = 0; i <= n1; ++i) {
      vertices[2 * ((size_t)j * (n1 + 1) + i)] = x1 / n1 * i;
      vertices[2 * ((size_t)j * (n1 + 1) + i) + 1] = x2 / n2 * j;
    }
  for (unsigned j = 0; j < n2; ++j)
    for (unsigned i = 0; i < n1; ++i) {
      unsigned* e = cells + 6 * ((size_t)n1 * j + i);
      const unsigned v00 = j * (n1 + 1) + i, v10 = v00 + 1, v01 = v00 + n1 + 1, v11 = v01 + 1;
      e[0] = v00; e[1] = v10; e[2] = v11;
      e[3] = v00; e[4] = v01; e[5] = v11;
    }
}

void orc_gen_cube_mesh(double x1, double x2, double x3, unsigned n1, unsigned n2, unsigned n3,
                       double* vertices, unsigned* cells) {
  /* mesh.cpp:54-135: five tetrahedra per cube, two orientations on (i+j+k) parity */
  const double hx = x1 / (double)n1, hy = x2 / (double)n2, hz = x3 / (double)n3;
  size_t id = 0;
  for (unsigned k = 0; k <= n3; ++k)
    for (unsigned j = 0; j <= n2; ++j)
      for (unsigned i = 0; i <= n1; ++i, ++id) {
        vertices[3 * id] = i * hx; vertices[3 * id + 1] = j * hy; vertices[3 * id + 2] = k * hz;
      }
#define VID(kk, jj, ii) ((unsigned)(((size_t)(kk) * (n2 + 1) + (jj)) * (n1 + 1) + (ii)))
  /* corner c = 4*dk + 2*dj + di */
  static const int even[5][4] = {{0, 1, 2, 4}, {1, 4, 5, 7}, {1, 2, 3, 7}, {2, 4, 6, 7}, {1, 2, 4, 7}};
  static const int odd[5][4]  = {{0, 1, 3, 5}, {0, 2, 3, 6}, {0, 4, 5, 6}, {3, 5, 6, 7}, {0, 3, 5, 6}};
  for (unsigned k = 0; k < n3; ++k)
    for (unsigned j = 0; j < n2; ++j)
      for (unsigned i = 0; i < n1; ++i) {
        unsigned* e = cells + 20 * (((size_t)k * n2 + j) * n1 + i);
        const int (*tab)[4] = ((i + j + k) % 2 == 0) ? even : odd;
        for (int t = 0; t < 5; ++t)
          for (int v = 0; v < 4; ++v) {
            const int c = tab[t][v];
            e[4 * t + v] = VID(k + ((c >> 2) & 1), j + ((c >> 1) & 1), i + (c & 1));
          }
      }
#undef VID
}

/* ------------------------------------------------------------------------- */
/* Geometry: cell.hpp:474-502 (triangle), :799-837 (tetrahedron)              */
/* ------------------------------------------------------------------------- */

static void cell_geometry(int dim, const double* vertices, const unsigned* cell, double* vol, double* jmt) {
  double J[3][3]; /* J[r][c] = v_{c+1}[r] - v_0[r] */
  for (int c = 0; c < dim; ++c)
    for (int r = 0; r < dim; ++r)
      J[r][c] = vertices[(size_t)cell[c + 1] * dim + r] - vertices[(size_t)cell[0] * dim + r];
  if (dim == 2) {
    const double det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
    *vol = 1.0 / 2.0 * fabs(det);
    /* inverse, then transpose (cell.hpp:498-499) */
    const double inv00 = J[1][1] / det, inv01 = -J[0][1] / det, inv10 = -J[1][0] / det, inv11 = J[0][0] / det;
    jmt[0] = inv00; jmt[1] = inv10; jmt[2] = inv01; jmt[3] = inv11;
  } else {
    const double a[3] = {J[0][0], J[1][0], J[2][0]}, b[3] = {J[0][1], J[1][1], J[2][1]}, c[3] = {J[0][2], J[1][2], J[2][2]};
    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
    const double det = cx * c[0] + cy * c[1] + cz * c[2];
    *vol = fabs(1.0 / 6.0 * det);
    double inv[3][3];
    inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    for (int s = 0; s < 3; ++s) for (int t = 0; t < 3; ++t) jmt[s * 3 + t] = inv[t][s];
  }
}

void orc_geometry(int dim, const double* vertices, const unsigned* cells, size_t K, double* vol, double* jmt) {
  for (size_t k = 0; k < K; ++k)
    cell_geometry(dim, vertices, cells + k * (dim + 1), vol + k, jmt + k * dim * dim);
}

/* ------------------------------------------------------------------------- */
/* Entities ("subdomains"): subdomain.hpp:36-43 ordering = (size, lexicographic) */
/* ------------------------------------------------------------------------- */

typedef struct { unsigned v[4]; unsigned cell; unsigned local; } ent_key;

static int ent_cmp(const void* pa, const void* pb) {
  const ent_key* a = (const ent_key*)pa; const ent_key* b = (const ent_key*)pb;
  for (int i = 0; i < 4; ++i) if (a->v[i] != b->v[i]) return a->v[i] < b->v[i] ? -1 : 1;
  return 0;
}
static int ent_cmp_stable(const void* pa, const void* pb) {
  const int c = ent_cmp(pa, pb);
  if (c) return c;
  const ent_key* a = (const ent_key*)pa; const ent_key* b = (const ent_key*)pb;
  if (a->cell != b->cell) return a->cell < b->cell ? -1 : 1;
  return a->local < b->local ? -1 : (a->local > b->local);
}

/* vertices of local entity (sd, j) of a simplex with dim+1 vertices */
static int local_entity(int dim, int sd, int j, int* lv) {
  if (sd == 0) { lv[0] = j; return 1; }
  if (sd == dim) { for (int i = 0; i <= dim; ++i) lv[i] = i; return dim + 1; }
  if (sd == 1) { local_edge(dim, j, &lv[0], &lv[1]); return 2; }
  local_triangle(j, lv); return 3; /* dim == 3, sd == 2 */
}

static void make_key(ent_key* key, const unsigned* cell, const int* lv, int n) {
  for (int i = 0; i < 4; ++i) key->v[i] = 0;
  for (int i = 0; i < n; ++i) key->v[i] = cell[lv[i]];
  /* pad so that the (size, lexicographic) order only compares same-size keys: unused slots = 0
   * and all keys of one kind have the same size */
}

/* sorted unique list of entities of kind sd (get_subdomain_list: cell.hpp:383-435, 692-764) */
static size_t entity_list(int dim, int sd, const unsigned* cells, size_t K, ent_key** out) {
  const int ne = n_entities(dim, sd);
  ent_key* keys = (ent_key*)malloc(sizeof(ent_key) * K * ne + 1);
  size_t n = 0;
  for (size_t k = 0; k < K; ++k)
    for (int j = 0; j < ne; ++j) {
      int lv[4]; const int m = local_entity(dim, sd, j, lv);
      make_key(&keys[n], cells + k * (dim + 1), lv, m);
      keys[n].cell = (unsigned)k; keys[n].local = (unsigned)j;
      ++n;
    }
  qsort(keys, n, sizeof(ent_key), ent_cmp_stable);
  size_t u = 0;
  for (size_t i = 0; i < n; ++i)
    if (u == 0 || ent_cmp(&keys[u - 1], &keys[i]) != 0) keys[u++] = keys[i];
  *out = keys;
  return u;
}

static size_t entity_rank(const ent_key* list, size_t n, const ent_key* key) {
  /* std::lower_bound (fes.hpp:52-55) */
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (ent_cmp(&list[mid], key) < 0) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* ------------------------------------------------------------------------- */
/* Boundary submesh: mesh.hpp:485-521                                         */
/* ------------------------------------------------------------------------- */

size_t orc_boundary_submesh(int dim, const unsigned* cells, size_t K,
                            unsigned** el, unsigned** el_id, unsigned** sd_id) {
  const int sd = dim - 1, ne = n_entities(dim, sd);
  ent_key* keys = (ent_key*)malloc(sizeof(ent_key) * K * ne + 1);
  size_t n = 0;
  for (size_t k = 0; k < K; ++k)
    for (int j = 0; j < ne; ++j) {
      int lv[4]; const int m = local_entity(dim, sd, j, lv);
      make_key(&keys[n], cells + k * (dim + 1), lv, m);
      keys[n].cell = (unsigned)k; keys[n].local = (unsigned)j;
      ++n;
    }
  /* std::map<face, count>: the stored key is the first (k, j) inserted, i.e. smallest k */
  qsort(keys, n, sizeof(ent_key), ent_cmp_stable);
  size_t F = 0;
  for (size_t i = 0; i < n;) {
    size_t j = i + 1;
    while (j < n && ent_cmp(&keys[i], &keys[j]) == 0) ++j;
    if (j - i == 1) ++F;
    i = j;
  }
  *el = (unsigned*)malloc(sizeof(unsigned) * (F * dim + 1));
  *el_id = (unsigned*)malloc(sizeof(unsigned) * (F + 1));
  *sd_id = (unsigned*)malloc(sizeof(unsigned) * (F + 1));
  size_t f = 0;
  for (size_t i = 0; i < n;) {
    size_t j = i + 1;
    while (j < n && ent_cmp(&keys[i], &keys[j]) == 0) ++j;
    if (j - i == 1) {
      for (int s = 0; s < dim; ++s) (*el)[f * dim + s] = keys[i].v[s];
      (*el_id)[f] = keys[i].cell; (*sd_id)[f] = keys[i].local;
      ++f;
    }
    i = j;
  }
  free(keys);
  return F;
}

/* Face neighbours: mesh.hpp:301-332. The cell loop runs in ascending k; the first cell that meets a face is
 * remembered, every later one links itself to the first and the first to itself (the last one wins there). */
void orc_cell_neighbours(int dim, const unsigned* cells, size_t K, int* nb) {
  const int sd = dim - 1, ne = n_entities(dim, sd);
  for (size_t i = 0; i < K * (size_t)(dim + 1); ++i) nb[i] = -1;
  ent_key* keys = (ent_key*)malloc(sizeof(ent_key) * K * ne + 1);
  size_t n = 0;
  for (size_t k = 0; k < K; ++k)
    for (int j = 0; j < ne; ++j) {
      int lv[4]; const int m = local_entity(dim, sd, j, lv);
      make_key(&keys[n], cells + k * (dim + 1), lv, m);
      keys[n].cell = (unsigned)k; keys[n].local = (unsigned)j;
      ++n;
    }
  qsort(keys, n, sizeof(ent_key), ent_cmp_stable);   /* equal faces stay in ascending (k, j) order */
  for (size_t i = 0; i < n;) {
    size_t j = i + 1;
    while (j < n && ent_cmp(&keys[i], &keys[j]) == 0) ++j;
    for (size_t t = i + 1; t < j; ++t) {
      nb[(size_t)keys[t].cell * (dim + 1) + keys[t].local] = (int)keys[i].cell;
      nb[(size_t)keys[i].cell * (dim + 1) + keys[i].local] = (int)keys[t].cell;
    }
    i = j;
  }
  free(keys);
}

/* ------------------------------------------------------------------------- */
/* Finite element space: fes.hpp:18-82                                        */
/* ------------------------------------------------------------------------- */

size_t orc_fes_build(int dim, int fe, const unsigned* cells, size_t K, unsigned* dof_map) {
  int per[4]; orc_fe_dofs_per_entity(dim, fe, per);
  const int nd = orc_fe_ndof(dim, fe);
  size_t global_off = 0; int local_off = 0;
  for (int sd = 0; sd <= dim; ++sd) {
    if (!per[sd]) continue;
    ent_key* list; const size_t n = entity_list(dim, sd, cells, K, &list);
    const int hat_m = per[sd], hat_n = n_entities(dim, sd);
    for (size_t k = 0; k < K; ++k)
      for (int hj = 0; hj < hat_n; ++hj) {
        int lv[4]; ent_key key; const int m = local_entity(dim, sd, hj, lv);
        make_key(&key, cells + k * (dim + 1), lv, m);
        const size_t j = entity_rank(list, n, &key);
        for (int hi = 0; hi < hat_m; ++hi)
          dof_map[k * nd + (hj * hat_m + hi) + local_off] = (unsigned)((j * hat_m + hi) + global_off);
      }
    global_off += (size_t)hat_m * n;
    local_off += hat_m * hat_n;
    free(list);
  }
  return global_off;
}

void orc_fes_g2l(const unsigned* dof_map, size_t K, int nd, size_t N, unsigned* g2l) {
  /* fes.hpp:76-81: the last (k, n) that maps to the dof wins */
  memset(g2l, 0, sizeof(unsigned) * 2 * N);
  for (size_t k = 0; k < K; ++k)
    for (int n = 0; n < nd; ++n) {
      g2l[2 * (size_t)dof_map[k * nd + n]] = (unsigned)k;
      g2l[2 * (size_t)dof_map[k * nd + n] + 1] = (unsigned)n;
    }
}

static int cmp_u32(const void* a, const void* b) {
  const unsigned x = *(const unsigned*)a, y = *(const unsigned*)b;
  return x < y ? -1 : (x > y);
}

size_t orc_fes_dirichlet_dofs(int dim, int fe, const unsigned* cells, size_t K,
                              const unsigned* bcells, size_t F, int nvb, unsigned** dofs) {
  /* fes.hpp:100-122: sub-entities of kind sd < n_subdomain_type of the DEFAULT boundary cell
   * type (dim-1 simplex => kinds 0..dim-1) of every boundary cell; dof = rank * m + i + offset */
  int per[4]; orc_fe_dofs_per_entity(dim, fe, per);
  size_t cap = 16, n = 0;
  unsigned* out = (unsigned*)malloc(sizeof(unsigned) * cap);
  size_t global_off = 0;
  for (int sd = 0; sd < dim; ++sd) {
    if (!per[sd]) continue;
    ent_key* list; const size_t nl = entity_list(dim, sd, cells, K, &list);
    const int m = sd + 1; /* vertices per sub-entity */
    for (size_t f = 0; f < F; ++f) {
      /* all m-subsets of the boundary cell's vertices; for the kind equal to the boundary
       * cell itself the reference inserts all nvb vertices (cell.hpp:173-184, 393-435) */
      int idx[4];
      if (nvb < m) continue;  /* e.g. the "edge" of a point submesh: never found in the edge list */
      for (int i = 0; i < m; ++i) idx[i] = i;
      for (;;) {
        ent_key key; for (int i = 0; i < 4; ++i) key.v[i] = 0;
        for (int i = 0; i < m; ++i) key.v[i] = bcells[f * nvb + idx[i]];
        const size_t j = entity_rank(list, nl, &key);
        if (j < nl && ent_cmp(&list[j], &key) == 0) {
          for (int hi = 0; hi < per[sd]; ++hi) {
            if (n == cap) { cap *= 2; out = (unsigned*)realloc(out, sizeof(unsigned) * cap); }
            out[n++] = (unsigned)((j * per[sd] + hi) + global_off);
          }
        }
        int p = m - 1;
        while (p >= 0 && idx[p] == nvb - m + p) --p;
        if (p < 0) break;
        ++idx[p];
        for (int q = p + 1; q < m; ++q) idx[q] = idx[q - 1] + 1;
      }
    }
    global_off += (size_t)per[sd] * nl;
    free(list);
  }
  qsort(out, n, sizeof(unsigned), cmp_u32);
  size_t u = 0;
  for (size_t i = 0; i < n; ++i) if (u == 0 || out[u - 1] != out[i]) out[u++] = out[i];
  *dofs = out;
  return u;
}

static void map_point(int dim, const double* vertices, const unsigned* cell, const double* xh, double* x) {
  /* cell.hpp:456-468 (triangle), :783-797 (tetrahedron): v0 + sum_m xh[m] (v_{m+1} - v0) */
  for (int n = 0; n < dim; ++n) {
    if (dim == 2) {
      x[n] = vertices[(size_t)cell[0] * 2 + n]
        + xh[0] * (vertices[(size_t)cell[1] * 2 + n] - vertices[(size_t)cell[0] * 2 + n])
        + xh[1] * (vertices[(size_t)cell[2] * 2 + n] - vertices[(size_t)cell[0] * 2 + n]);
    } else {
      double acc = vertices[(size_t)cell[0] * 3 + n];
      for (int m = 0; m < 3; ++m)
        acc += xh[m] * (vertices[(size_t)cell[m + 1] * 3 + n] - vertices[(size_t)cell[0] * 3 + n]);
      x[n] = acc;
    }
  }
}

void orc_dof_coordinates(int dim, int fe, const double* vertices, const unsigned* cells,
                         const unsigned* g2l, const unsigned* dofs, size_t n, double* x) {
  /* fes.hpp:190-202 */
  double nodes[10 * 3];
  orc_fe_nodes(dim, fe, nodes);
  for (size_t i = 0; i < n; ++i) {
    const unsigned k = g2l[2 * (size_t)dofs[i]], ln = g2l[2 * (size_t)dofs[i] + 1];
    map_point(dim, vertices, cells + (size_t)k * (dim + 1), nodes + ln * dim, x + i * dim);
  }
}

/* ------------------------------------------------------------------------- */
/* Sparse pattern: solver.hpp:259-265 (std::map COO), solver.cpp:71-96 (map -> CRS) */
/* ------------------------------------------------------------------------- */

static int cmp_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y);
}

static size_t space_total(const orc_space* s, size_t* offsets) {
  size_t tot = 0;
  for (int c = 0; c < s->n_comp; ++c) { if (offsets) offsets[c] = tot; tot += s->n_dof[c]; }
  return tot;
}

size_t orc_pattern(size_t K, const orc_space* te, const orc_space* tr,
                   const unsigned char* dirichlet_row, int** rowptr, int** colind) {
  size_t off_te[8], off_tr[8];
  const size_t N = space_total(te, off_te);
  space_total(tr, off_tr);
  const int dim_te = te->dim, dim_tr = tr->dim;
  const int nc_te = te->n_comp, nc_tr = tr->n_comp;
  size_t pairs_per_cell = 0;
  int nd_te[8], nd_tr[8];
  for (int m = 0; m < nc_te; ++m) nd_te[m] = orc_fe_ndof(dim_te, te->fe[m]);
  for (int n = 0; n < nc_tr; ++n) nd_tr[n] = orc_fe_ndof(dim_tr, tr->fe[n]);
  for (int m = 0; m < nc_te; ++m) for (int n = 0; n < nc_tr; ++n) pairs_per_cell += (size_t)nd_te[m] * nd_tr[n];
  size_t n_dir = 0;
  for (size_t i = 0; i < N; ++i) n_dir += dirichlet_row ? dirichlet_row[i] : 0;
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (K * pairs_per_cell + n_dir + 1));
  size_t n = 0;
  /* bilinear_form.hpp:159-165 / composite :282-299: identity rows of Dirichlet dofs */
  for (size_t i = 0; i < N; ++i)
    if (dirichlet_row && dirichlet_row[i]) keys[n++] = ((uint64_t)i << 32) | (uint64_t)i;
  /* bilinear_form.hpp:103-108,183-186: every (i, j) of every cell unless the row is Dirichlet */
  {
    size_t o_te = 0;
    for (int m = 0; m < nc_te; ++m) {
      size_t o_tr = 0;
      for (int c = 0; c < nc_tr; ++c) {
        for (size_t k = 0; k < K; ++k)
          for (int i = 0; i < nd_te[m]; ++i) {
            const uint64_t row = (uint64_t)te->dof_map[m][k * nd_te[m] + i] + o_te;
            if (dirichlet_row && dirichlet_row[row]) continue;
            for (int j = 0; j < nd_tr[c]; ++j)
              keys[n++] = (row << 32) | ((uint64_t)tr->dof_map[c][k * nd_tr[c] + j] + o_tr);
          }
        o_tr += tr->n_dof[c];
      }
      o_te += te->n_dof[m];
    }
  }
  qsort(keys, n, sizeof(uint64_t), cmp_u64);
  size_t u = 0;
  for (size_t i = 0; i < n; ++i) if (u == 0 || keys[u - 1] != keys[i]) keys[u++] = keys[i];
  *rowptr = (int*)calloc(N + 1, sizeof(int));
  *colind = (int*)malloc(sizeof(int) * (u + 1));
  for (size_t e = 0; e < u; ++e) {
    (*rowptr)[(keys[e] >> 32) + 1] += 1;
    (*colind)[e] = (int)(keys[e] & 0xffffffffu);
  }
  for (size_t i = 0; i < N; ++i) (*rowptr)[i + 1] += (*rowptr)[i];
  free(keys);
  return u;
}

/* ------------------------------------------------------------------------- */
/* Assembly                                                                   */
/* ------------------------------------------------------------------------- */

static size_t csr_find(const int* rowptr, const int* colind, size_t row, int col) {
  int lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (colind[mid] < col) lo = mid + 1; else hi = mid;
  }
  return (size_t)lo;
}

typedef struct {
  int dim, nq;
  double xq[56 * 3], wq[56];
} quad_t;

static int load_quad(const char* name, quad_t* q) {
  q->nq = orc_quadrature(name, q->xq, q->wq, &q->dim);
  return q->nq;
}

/* fe_value_manager.hpp:103-127: phi[q][i][1+s] = sum_t jmt[s][t] * phi_hat[q][i][1+t] */
static void push_forward(int dim, int nq, int nd, const double* hat, const double* jmt, double* out) {
  const int w = dim + 1;
  for (int q = 0; q < nq; ++q)
    for (int i = 0; i < nd; ++i) {
      const double* h = hat + ((size_t)q * nd + i) * w;
      double* o = out + ((size_t)q * nd + i) * w;
      o[0] = h[0];
      for (int s = 0; s < dim; ++s) {
        double acc = 0.0;
        for (int t = 0; t < dim; ++t) acc += jmt[s * dim + t] * h[1 + t];
        o[1 + s] = acc;
      }
    }
}

#define ORC_MAX_COMP 4
#define ORC_MAX_ND 10

void orc_assemble_bilinear(int dim, const double* vertices, const unsigned* cells, size_t K,
                           const orc_space* te, const orc_space* tr,
                           const char* quadrature, const orc_term* terms, int n_terms,
                           const unsigned char* dirichlet_row, int composite,
                           const int* rowptr, const int* colind, double* val) {
  quad_t Q; load_quad(quadrature, &Q);
  const int nq = Q.nq, w = dim + 1;
  const int nc_te = te->n_comp, nc_tr = tr->n_comp;
  size_t off_te[8], off_tr[8]; space_total(te, off_te); space_total(tr, off_tr);
  int nd_te[ORC_MAX_COMP], nd_tr[ORC_MAX_COMP];
  double* hat_te[ORC_MAX_COMP]; double* hat_tr[ORC_MAX_COMP];
  double* val_te[ORC_MAX_COMP]; double* val_tr[ORC_MAX_COMP];
  for (int m = 0; m < nc_te; ++m) {
    nd_te[m] = orc_fe_ndof(dim, te->fe[m]);
    hat_te[m] = (double*)malloc(sizeof(double) * nq * nd_te[m] * w);
    val_te[m] = (double*)malloc(sizeof(double) * nq * nd_te[m] * w);
    orc_fe_tabulate(dim, te->fe[m], nq, Q.xq, hat_te[m]);
  }
  for (int n = 0; n < nc_tr; ++n) {
    nd_tr[n] = orc_fe_ndof(dim, tr->fe[n]);
    hat_tr[n] = (double*)malloc(sizeof(double) * nq * nd_tr[n] * w);
    val_tr[n] = (double*)malloc(sizeof(double) * nq * nd_tr[n] * w);
    orc_fe_tabulate(dim, tr->fe[n], nq, Q.xq, hat_tr[n]);
  }
  for (size_t k = 0; k < K; ++k) {
    double vol, jmt[9];
    cell_geometry(dim, vertices, cells + k * (dim + 1), &vol, jmt);
    for (int m = 0; m < nc_te; ++m) push_forward(dim, nq, nd_te[m], hat_te[m], jmt, val_te[m]);
    for (int n = 0; n < nc_tr; ++n) push_forward(dim, nq, nd_tr[n], hat_tr[n], jmt, val_tr[n]);
    for (int m = 0; m < nc_te; ++m)
      for (int n = 0; n < nc_tr; ++n) {
        double a_el[ORC_MAX_ND][ORC_MAX_ND];
        memset(a_el, 0, sizeof(a_el));
        for (int q = 0; q < nq; ++q)
          for (int i = 0; i < nd_te[m]; ++i) {
            const size_t row = (size_t)te->dof_map[m][k * nd_te[m] + i] + off_te[m];
            for (int j = 0; j < nd_tr[n]; ++j) {
              const double* psi = val_te[m] + ((size_t)q * nd_te[m] + i) * w;
              const double* phi = val_tr[n] + ((size_t)q * nd_tr[n] + j) * w;
              double f = 0.0;
              for (int t = 0; t < n_terms; ++t) {
                if (terms[t].test_comp != m || terms[t].trial_comp != n) continue;
                const double c = terms[t].table ? terms[t].c * terms[t].table[k * nq + q] : terms[t].c;
                f += c * psi[terms[t].test_d] * phi[terms[t].trial_d];
              }
              if (composite) {
                /* composite_bilinear_form.hpp:98-117: one add per (q, i, j) */
                if (!(dirichlet_row && dirichlet_row[row])) {
                  const size_t col = (size_t)tr->dof_map[n][k * nd_tr[n] + j] + off_tr[n];
                  val[csr_find(rowptr, colind, row, (int)col)] += vol * Q.wq[q] * f;
                }
              } else {
                a_el[i][j] += Q.wq[q] * f;  /* bilinear_form.hpp:94-98 */
              }
            }
          }
        if (!composite)
          for (int i = 0; i < nd_te[m]; ++i) {
            const size_t row = (size_t)te->dof_map[m][k * nd_te[m] + i] + off_te[m];
            if (dirichlet_row && dirichlet_row[row]) continue;  /* bilinear_form.hpp:183-186 */
            for (int j = 0; j < nd_tr[n]; ++j) {
              const size_t col = (size_t)tr->dof_map[n][k * nd_tr[n] + j] + off_tr[n];
              val[csr_find(rowptr, colind, row, (int)col)] += a_el[i][j] * vol;  /* :103-108 */
            }
          }
      }
  }
  for (int m = 0; m < nc_te; ++m) { free(hat_te[m]); free(val_te[m]); }
  for (int n = 0; n < nc_tr; ++n) { free(hat_tr[n]); free(val_tr[n]); }
}

void orc_assemble_linear(int dim, const double* vertices, const unsigned* cells, size_t K,
                         const orc_space* te, const char* quadrature,
                         const orc_term* terms, int n_terms, int composite, double* b) {
  /* linear_form.hpp:67-142 (scalar: rhs_el += w_q * vol * f, then ordered scatter :55-63),
   * composite_linear_form.hpp:60-76 (one add per (q, i)) */
  quad_t Q; load_quad(quadrature, &Q);
  const int nq = Q.nq, w = dim + 1;
  const int nc_te = te->n_comp;
  size_t off_te[8]; space_total(te, off_te);
  int nd_te[ORC_MAX_COMP]; double* hat_te[ORC_MAX_COMP]; double* val_te[ORC_MAX_COMP];
  for (int m = 0; m < nc_te; ++m) {
    nd_te[m] = orc_fe_ndof(dim, te->fe[m]);
    hat_te[m] = (double*)malloc(sizeof(double) * nq * nd_te[m] * w);
    val_te[m] = (double*)malloc(sizeof(double) * nq * nd_te[m] * w);
    orc_fe_tabulate(dim, te->fe[m], nq, Q.xq, hat_te[m]);
  }
  for (size_t k = 0; k < K; ++k) {
    double vol, jmt[9];
    cell_geometry(dim, vertices, cells + k * (dim + 1), &vol, jmt);
    for (int m = 0; m < nc_te; ++m) {
      push_forward(dim, nq, nd_te[m], hat_te[m], jmt, val_te[m]);
      double rhs_el[ORC_MAX_ND];
      memset(rhs_el, 0, sizeof(rhs_el));
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < nd_te[m]; ++i) {
          const double* psi = val_te[m] + ((size_t)q * nd_te[m] + i) * w;
          double f = 0.0;
          for (int t = 0; t < n_terms; ++t) {
            if (terms[t].test_comp != m) continue;
            const double c = terms[t].table ? terms[t].c * terms[t].table[k * nq + q] : terms[t].c;
            f += c * psi[terms[t].test_d];
          }
          if (composite)
            b[(size_t)te->dof_map[m][k * nd_te[m] + i] + off_te[m]] += vol * Q.wq[q] * f;
          else
            rhs_el[i] += Q.wq[q] * vol * f;
        }
      if (!composite)
        for (int i = 0; i < nd_te[m]; ++i)
          b[(size_t)te->dof_map[m][k * nd_te[m] + i] + off_te[m]] += rhs_el[i];
    }
  }
  for (int m = 0; m < nc_te; ++m) { free(hat_te[m]); free(val_te[m]); }
}

double orc_integrate(int dim, const double* vertices, const unsigned* cells, size_t K,
                     const char* quadrature, const double* table) {
  /* form.hpp:87-125 */
  quad_t Q; load_quad(quadrature, &Q);
  double result = 0.0;
  for (size_t k = 0; k < K; ++k) {
    double vol, jmt[9];
    cell_geometry(dim, vertices, cells + k * (dim + 1), &vol, jmt);
    double el = 0.0;
    for (int q = 0; q < Q.nq; ++q) el += Q.wq[q] * table[k * Q.nq + q];
    result += el * vol;
  }
  return result;
}

void orc_quadrature_points(int dim, const double* vertices, const unsigned* cells, size_t K,
                           const char* quadrature, double* xq) {
  quad_t Q; load_quad(quadrature, &Q);
  for (size_t k = 0; k < K; ++k)
    for (int q = 0; q < Q.nq; ++q)
      map_point(dim, vertices, cells + k * (dim + 1), Q.xq + q * dim, xq + (k * Q.nq + q) * dim);
}

void orc_field_at_quadrature(int dim, int fe, const double* vertices, const unsigned* cells, size_t K,
                             const unsigned* dof_map, const double* coef, int d,
                             const char* quadrature, double* out) {
  /* expression.hpp:114-128: sum_i coef[dof(k,i)] * phi[q][i][d] */
  quad_t Q; load_quad(quadrature, &Q);
  const int nq = Q.nq, nd = orc_fe_ndof(dim, fe), w = dim + 1;
  double* hat = (double*)malloc(sizeof(double) * nq * nd * w);
  double* val = (double*)malloc(sizeof(double) * nq * nd * w);
  orc_fe_tabulate(dim, fe, nq, Q.xq, hat);
  for (size_t k = 0; k < K; ++k) {
    double vol, jmt[9];
    cell_geometry(dim, vertices, cells + k * (dim + 1), &vol, jmt);
    push_forward(dim, nq, nd, hat, jmt, val);
    for (int q = 0; q < nq; ++q) {
      double acc = 0.0;
      for (int i = 0; i < nd; ++i) acc += coef[dof_map[k * nd + i]] * val[((size_t)q * nd + i) * w + d];
      out[k * nq + q] = acc;
    }
  }
  free(hat); free(val);
}

/* ------------------------------------------------------------------------- */
/* Linear algebra + Krylov (restates PETSc 3.7 semantics named at solver.hpp:139-163) */
/* ------------------------------------------------------------------------- */

void orc_spmv(int n, const int* rowptr, const int* colind, const double* val, const double* x, double* y) {
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) acc += val[e] * x[colind[e]];
    y[i] = acc;
  }
}

static double dot(int n, const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

int orc_cg_jacobi(int n, const int* rowptr, const int* colind, const double* val,
                  const double* b, double* x, double rtol, double atol, int maxits, double* final_rnorm) {
  /* Jacobi-preconditioned conjugate gradients on the reference's row-replaced Dirichlet system:
   * x0 = b on identity rows (the lift), 0 elsewhere, so r0 vanishes on those rows and the
   * iteration runs on the symmetric interior block. Stop: ||r||_2 <= max(rtol ||b||_2, atol). */
  double* r = (double*)malloc(sizeof(double) * n); double* z = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n); double* q = (double*)malloc(sizeof(double) * n);
  double* dinv = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    double d = 1.0; int identity = (rowptr[i + 1] - rowptr[i] == 1 && colind[rowptr[i]] == i && val[rowptr[i]] == 1.0);
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) if (colind[e] == i) d = val[e];
    dinv[i] = 1.0 / d;
    x[i] = identity ? b[i] : 0.0;
  }
  orc_spmv(n, rowptr, colind, val, x, q);
  for (int i = 0; i < n; ++i) r[i] = b[i] - q[i];
  const double bnorm = sqrt(dot(n, b, b));
  const double target = fmax(rtol * bnorm, atol);
  for (int i = 0; i < n; ++i) { z[i] = dinv[i] * r[i]; p[i] = z[i]; }
  double rz = dot(n, r, z), rnorm = sqrt(dot(n, r, r));
  int it = 0;
  while (rnorm > target && it < maxits) {
    orc_spmv(n, rowptr, colind, val, p, q);
    const double alpha = rz / dot(n, p, q);
    for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
    for (int i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
    const double rz_new = dot(n, r, z);
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    rnorm = sqrt(dot(n, r, r));
    ++it;
  }
  if (final_rnorm) *final_rnorm = rnorm;
  free(r); free(z); free(p); free(q); free(dinv);
  return it;
}

/* ILU(k) with level-of-fill on a CSR matrix with sorted columns, natural ordering
 * (PCILU + PCFactorSetLevels, solver.hpp:162-163). Returns pattern + factors. */
typedef struct { int n; int* rowptr; int* colind; int* diag; double* lu; } ilu_t;

static void ilu_free(ilu_t* f) { free(f->rowptr); free(f->colind); free(f->diag); free(f->lu); }

static int ilu_build(int n, const int* rowptr, const int* colind, const double* val, int levels, ilu_t* f) {
  f->n = n;
  f->rowptr = (int*)calloc((size_t)n + 1, sizeof(int));
  f->diag = (int*)malloc(sizeof(int) * (size_t)n);
  size_t cap = (size_t)rowptr[n] * (levels > 0 ? 3 : 1) + 16, nnz = 0;
  f->colind = (int*)malloc(sizeof(int) * cap);
  int* lev = (int*)malloc(sizeof(int) * cap);
  /* symbolic: row-merge with levels */
  int* w_lev = (int*)malloc(sizeof(int) * (size_t)n);
  int* w_next = (int*)malloc(sizeof(int) * ((size_t)n + 1));  /* sorted linked list over columns */
  for (int i = 0; i < n; ++i) w_lev[i] = -1;
  for (int i = 0; i < n; ++i) {
    /* row i starts here; known before the merge so that the U part of row i - 1 ends at the right place */
    f->rowptr[i] = (int)nnz;
    /* load row i */
    int head = n; /* sentinel */
    int prev = -1;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const int c = colind[e];
      w_lev[c] = 0;
      if (prev < 0) head = c; else w_next[prev] = c;
      prev = c;
    }
    int has_diag = (w_lev[i] == 0);
    if (prev >= 0) w_next[prev] = n;
    if (!has_diag) {
      /* insert diagonal with level 0 (PETSc requires a diagonal; FE matrices always have it) */
      int* link = &head; int c = head;
      while (c < i) { link = &w_next[c]; c = w_next[c]; }
      w_next[i] = c; *link = i; w_lev[i] = 0;
    }
    for (int k = head; k < i; k = w_next[k]) {
      const int lk = w_lev[k];
      if (lk > levels) continue;
      /* merge U part of row k */
      int pos = k;
      for (int e = f->diag[k] + 1; e < f->rowptr[k + 1]; ++e) {
        const int c = f->colind[e];
        const int nl = lk + lev[e] + 1;
        if (nl > levels) continue;
        if (w_lev[c] >= 0) { if (nl < w_lev[c]) w_lev[c] = nl; continue; }
        /* insert c after pos in sorted order */
        int p = pos;
        while (w_next[p] < c) p = w_next[p];
        w_next[c] = w_next[p]; w_next[p] = c; w_lev[c] = nl;
        pos = c;
      }
    }
    /* store row */
    f->rowptr[i] = (int)nnz;
    for (int k = head; k < n;) {
      if (nnz + 1 >= cap) { cap *= 2; f->colind = (int*)realloc(f->colind, sizeof(int) * cap); lev = (int*)realloc(lev, sizeof(int) * cap); }
      if (w_lev[k] <= levels) {
        if (k == i) f->diag[i] = (int)nnz;
        f->colind[nnz] = k; lev[nnz] = w_lev[k]; ++nnz;
      }
      const int nx = w_next[k];
      w_lev[k] = -1;
      k = nx;
    }
  }
  f->rowptr[n] = (int)nnz;
  free(w_lev); free(w_next); free(lev);
  /* numeric: IKJ variant */
  f->lu = (double*)calloc(nnz + 1, sizeof(double));
  double* wv = (double*)calloc((size_t)n, sizeof(double));
  int status = 0;
  for (int i = 0; i < n; ++i) {
    for (int e = f->rowptr[i]; e < f->rowptr[i + 1]; ++e) wv[f->colind[e]] = 0.0;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) wv[colind[e]] = val[e];
    for (int e = f->rowptr[i]; e < f->diag[i]; ++e) {
      const int k = f->colind[e];
      const double l = wv[k] / f->lu[f->diag[k]];
      wv[k] = l;
      if (l == 0.0) continue;
      /* only positions present in row i's pattern are updated */
      int ei = e + 1;
      for (int ek = f->diag[k] + 1; ek < f->rowptr[k + 1]; ++ek) {
        const int c = f->colind[ek];
        while (ei < f->rowptr[i + 1] && f->colind[ei] < c) ++ei;
        if (ei < f->rowptr[i + 1] && f->colind[ei] == c) wv[c] -= l * f->lu[ek];
      }
    }
    for (int e = f->rowptr[i]; e < f->rowptr[i + 1]; ++e) f->lu[e] = wv[f->colind[e]];
    if (f->lu[f->diag[i]] == 0.0) status = i + 1;
  }
  free(wv);
  return status;
}

static void ilu_solve(const ilu_t* f, const double* b, double* x) {
  const int n = f->n;
  for (int i = 0; i < n; ++i) {
    double acc = b[i];
    for (int e = f->rowptr[i]; e < f->diag[i]; ++e) acc -= f->lu[e] * x[f->colind[e]];
    x[i] = acc;
  }
  for (int i = n - 1; i >= 0; --i) {
    double acc = x[i];
    for (int e = f->diag[i] + 1; e < f->rowptr[i + 1]; ++e) acc -= f->lu[e] * x[f->colind[e]];
    x[i] = acc / f->lu[f->diag[i]];
  }
}

/* pattern of ILU(levels) (row-merge with the sum rule lev = lev_ik + lev_kj + 1, entries above `levels` dropped):
 * out_rowptr has n + 1 entries, *out_colind is malloc'ed (orc_free). Returns the number of stored entries. */
size_t orc_iluk_pattern(int n, const int* rowptr, const int* colind, int levels, int* out_rowptr, int** out_colind) {
  ilu_t f;
  double* ones = (double*)malloc(sizeof(double) * ((size_t)rowptr[n] + 1));
  for (int e = 0; e < rowptr[n]; ++e) ones[e] = 1.0;
  ilu_build(n, rowptr, colind, ones, levels, &f);
  free(ones);
  const size_t nnz = (size_t)f.rowptr[n];
  memcpy(out_rowptr, f.rowptr, sizeof(int) * ((size_t)n + 1));
  *out_colind = (int*)malloc(sizeof(int) * (nnz + 1));
  memcpy(*out_colind, f.colind, sizeof(int) * nnz);
  ilu_free(&f);
  return nnz;
}

/* ILU(levels): pattern and factors together (what PCILU + PCFactorSetLevels builds, solver.hpp:162-163). out_rowptr has
 * n + 1 entries; *out_colind / *out_lu are malloc'ed (orc_free). Returns the number of stored entries. */
size_t orc_iluk(int n, const int* rowptr, const int* colind, const double* val, int levels, int* out_rowptr,
                int** out_colind, double** out_lu) {
  ilu_t f;
  ilu_build(n, rowptr, colind, val, levels, &f);
  const size_t nnz = (size_t)f.rowptr[n];
  memcpy(out_rowptr, f.rowptr, sizeof(int) * ((size_t)n + 1));
  *out_colind = (int*)malloc(sizeof(int) * (nnz + 1));
  *out_lu = (double*)malloc(sizeof(double) * (nnz + 1));
  memcpy(*out_colind, f.colind, sizeof(int) * nnz);
  memcpy(*out_lu, f.lu, sizeof(double) * nnz);
  ilu_free(&f);
  return nnz;
}

int orc_ilu0(int n, const int* rowptr, const int* colind, const double* val, double* lu) {
  ilu_t f;
  const int st = ilu_build(n, rowptr, colind, val, 0, &f);
  /* with levels = 0 and a full diagonal the pattern equals the input pattern */
  if (f.rowptr[n] != rowptr[n]) { ilu_free(&f); return -1; }
  memcpy(lu, f.lu, sizeof(double) * (size_t)rowptr[n]);
  ilu_free(&f);
  return st;
}

int orc_gmres_ilu(int n, const int* rowptr, const int* colind, const double* val,
                  const double* b, double* x, int restart, int ilu_levels,
                  double rtol, double atol, double dtol, int maxits, double* final_rnorm) {
  /* Left-preconditioned restarted GMRES with modified Gram-Schmidt, x0 = 0, convergence on the
   * preconditioned residual norm: ||M^-1 r|| <= max(rtol ||M^-1 b||, atol); divergence when
   * > dtol ||M^-1 b|| (KSPSetTolerances, solver.hpp:150-156). */
  ilu_t M;
  ilu_build(n, rowptr, colind, val, ilu_levels, &M);
  if (restart > maxits) restart = maxits;
  if (restart < 1) restart = 1;
  const int m = restart;
  double** Vc = (double**)calloc((size_t)m + 1, sizeof(double*));  /* basis columns, allocated on demand */
#define VCOL(j) (Vc[j] ? Vc[j] : (Vc[j] = (double*)malloc(sizeof(double) * (size_t)n)))
  double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
  double* cs = (double*)malloc(sizeof(double) * m); double* sn = (double*)malloc(sizeof(double) * m);
  double* g = (double*)malloc(sizeof(double) * (m + 1)); double* y = (double*)malloc(sizeof(double) * m);
  double* t = (double*)malloc(sizeof(double) * n); double* u = (double*)malloc(sizeof(double) * n);
  memset(x, 0, sizeof(double) * n);
  ilu_solve(&M, b, u);
  const double bnorm = sqrt(dot(n, u, u));
  const double target = fmax(rtol * bnorm, atol);
  int its = 0; double rnorm = bnorm;
  while (its < maxits) {
    orc_spmv(n, rowptr, colind, val, x, t);
    for (int i = 0; i < n; ++i) t[i] = b[i] - t[i];
    double* v0 = VCOL(0);
    ilu_solve(&M, t, v0);
    rnorm = sqrt(dot(n, v0, v0));
    if (rnorm <= target || rnorm > dtol * bnorm) break;
    for (int i = 0; i < n; ++i) v0[i] /= rnorm;
    memset(g, 0, sizeof(double) * (m + 1)); g[0] = rnorm;
    int k = 0;
    for (; k < m && its < maxits; ++k) {
      double* vk1 = VCOL(k + 1);
      orc_spmv(n, rowptr, colind, val, Vc[k], t);
      ilu_solve(&M, t, vk1);
      for (int j = 0; j <= k; ++j) {  /* modified Gram-Schmidt (solver.hpp:155) */
        const double h = dot(n, vk1, Vc[j]);
        H[(size_t)j * m + k] = h;
        for (int i = 0; i < n; ++i) vk1[i] -= h * Vc[j][i];
      }
      const double hn = sqrt(dot(n, vk1, vk1));
      H[(size_t)(k + 1) * m + k] = hn;
      if (hn != 0.0) for (int i = 0; i < n; ++i) vk1[i] /= hn;
      for (int j = 0; j < k; ++j) {
        const double a = H[(size_t)j * m + k], bb = H[(size_t)(j + 1) * m + k];
        H[(size_t)j * m + k] = cs[j] * a + sn[j] * bb;
        H[(size_t)(j + 1) * m + k] = -sn[j] * a + cs[j] * bb;
      }
      const double a = H[(size_t)k * m + k], bb = H[(size_t)(k + 1) * m + k];
      const double rr = hypot(a, bb);
      cs[k] = rr == 0.0 ? 1.0 : a / rr; sn[k] = rr == 0.0 ? 0.0 : bb / rr;
      H[(size_t)k * m + k] = rr; H[(size_t)(k + 1) * m + k] = 0.0;
      g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
      ++its;
      rnorm = fabs(g[k + 1]);
      if (rnorm <= target || hn == 0.0) { ++k; break; }
    }
    for (int i = k - 1; i >= 0; --i) {
      double acc = g[i];
      for (int j = i + 1; j < k; ++j) acc -= H[(size_t)i * m + j] * y[j];
      y[i] = acc / H[(size_t)i * m + i];
    }
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < n; ++i) x[i] += y[j] * Vc[j][i];
    if (rnorm <= target) break;
  }
  if (final_rnorm) *final_rnorm = rnorm;
  for (int j = 0; j <= m; ++j) free(Vc[j]);
  free(Vc); free(H);
#undef VCOL
  free(cs); free(sn); free(g); free(y); free(t); free(u);
  ilu_free(&M);
  return its;
}
