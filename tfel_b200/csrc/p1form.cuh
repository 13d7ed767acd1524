// Constant-coefficient forms on P1 triangles in closed form (shared by tile.cu and p1rows.cu).
// On a P1 cell the gradients of the basis are constant: ghat_0 = (-1, -1), ghat_1 = (1, 0), ghat_2 = (0, 1) on the
// reference triangle (fe.hpp:260-317), so the element matrix of  sum_t c_t D^a psi_i D^b phi_j  is
//     K_ij |K|,  K_ij = W ghat_i^T G ghat_j + mhat_i (g0r . ghat_j) + (gr0 . ghat_i) mhat_j + c00 Mhat_ij
// with G = Jmt^T C Jmt, g0r = Jmt^T c0r, gr0 = Jmt^T cr0 and the reference integrals W = sum_q w_q,
// mhat_i = sum_q w_q phihat_i, Mhat_ij = sum_q w_q phihat_i phihat_j of the form's quadrature formula -- the same numbers
// the generic contraction of element.cuh reads from its pre-integrated reference tensor.
#pragma once

#include "element.cuh"

namespace tfelcu {

struct P1Form {
  int has00, has0r, hasr0, hasrr;
  double c00, c0r[2], cr0[2], C[2][2];
  double wsum;          // sum_q w_q
  double mhat[3];       // sum_q w_q phi_i
  double Mhat[9];       // sum_q w_q phi_i phi_j
  double b_c[3];        // fused linear form: sum_t c_t mhat_i of ITS quadrature
  int b_uniform;        // b_c[0] == b_c[1] == b_c[2] bit for bit (any symmetric formula): no select per cell
  // isotropic shape: C = c I (Laplacian-like) and, if there is a value x value term, a reference mass matrix with one
  // diagonal and one off-diagonal value (any symmetric quadrature formula): K_ij = (c W / 4|K|) e_i . e_j + c00 |K| M_ij
  // with e_i the edge opposite to vertex i -- no Jacobian inverse, nothing that depends on the local numbering
  int iso;
  double iso_k;         // 0.5 c W
  double iso_md, iso_mo;   // c00 Mhat_ii, c00 Mhat_ij (i != j)
};

// fb: the bilinear form (constant coefficients, P1 x P1 on triangles); fbv: a constant-coefficient linear form without
// derivatives that rides along, or nullptr
inline P1Form make_p1_form(const FormBuild& fb, const FormBuild* fbv) {
  P1Form F;
  const FormDev& A = fb.F;
  F.has00 = A.has00; F.has0r = A.has0r; F.hasr0 = A.hasr0; F.hasrr = A.hasrr;
  F.c00 = A.c00;
  for (int d = 0; d < 2; ++d) { F.c0r[d] = A.c0r[d]; F.cr0[d] = A.cr0[d]; for (int e = 0; e < 2; ++e) F.C[d][e] = A.C[d][e]; }
  // ref[((a W + b) 3 + i) 3 + j] = sum_q w_q D^a phihat_i D^b phihat_j, W = 3 (value, d/dxhat_0, d/dxhat_1)
  const double* ref = fb.tables.data() + A.off_ref;
  for (int e = 0; e < 9; ++e) F.Mhat[e] = ref[e];
  for (int i = 0; i < 3; ++i) F.mhat[i] = ref[((0 * 3 + 1) * 3 + i) * 3 + 1];   // mhat_i ghat_1[0], ghat_1 = (1, 0)
  F.wsum = ref[((1 * 3 + 1) * 3 + 1) * 3 + 1];                                 // ghat_1[0]^2 sum_q w_q
  F.iso = 0; F.iso_k = 0.0; F.iso_md = 0.0; F.iso_mo = 0.0;
  if (A.hasrr && !A.has0r && !A.hasr0 && A.C[0][0] == A.C[1][1] && A.C[0][1] == 0.0 && A.C[1][0] == 0.0) {
    bool mass_ok = true;
    if (A.has00) {
      const double d = F.Mhat[0], o = F.Mhat[1];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) mass_ok = mass_ok && F.Mhat[i * 3 + j] == (i == j ? d : o);
      if (mass_ok) { F.iso_md = A.c00 * d; F.iso_mo = A.c00 * o; }
    }
    if (mass_ok) { F.iso = 1; F.iso_k = 0.5 * A.C[0][0] * F.wsum; }
  }
  F.b_c[0] = F.b_c[1] = F.b_c[2] = 0.0;
  if (fbv) {
    const FormDev& G = fbv->F;   // sum_t c_t psi_i integrated on the reference cell: mhat[0][i] of its reference table
    double c = 0.0;
    for (int t = 0; t < G.n_terms; ++t) c += G.terms[t].c;
    for (int i = 0; i < 3; ++i) F.b_c[i] = c * fbv->tables[(size_t)G.off_ref + i];
  }
  F.b_uniform = (F.b_c[0] == F.b_c[1] && F.b_c[1] == F.b_c[2]) ? 1 : 0;
  return F;
}

// whole element matrix (without |K|)
__device__ __forceinline__ void p1_element(const P1Form& F, const double (&jmt)[2][2], double (&K)[9]) {
#pragma unroll
  for (int e = 0; e < 9; ++e) K[e] = 0.0;
  if (F.has00) {
#pragma unroll
    for (int e = 0; e < 9; ++e) K[e] = F.c00 * F.Mhat[e];
  }
  if (F.has0r) {
    const double g1 = F.c0r[0] * jmt[0][0] + F.c0r[1] * jmt[1][0], g2 = F.c0r[0] * jmt[0][1] + F.c0r[1] * jmt[1][1];
    const double g0 = -(g1 + g2);
#pragma unroll
    for (int i = 0; i < 3; ++i) { K[i * 3 + 0] += F.mhat[i] * g0; K[i * 3 + 1] += F.mhat[i] * g1; K[i * 3 + 2] += F.mhat[i] * g2; }
  }
  if (F.hasr0) {
    const double g1 = F.cr0[0] * jmt[0][0] + F.cr0[1] * jmt[1][0], g2 = F.cr0[0] * jmt[0][1] + F.cr0[1] * jmt[1][1];
    const double g0 = -(g1 + g2);
#pragma unroll
    for (int j = 0; j < 3; ++j) { K[0 * 3 + j] += g0 * F.mhat[j]; K[1 * 3 + j] += g1 * F.mhat[j]; K[2 * 3 + j] += g2 * F.mhat[j]; }
  }
  if (F.hasrr) {
    double CJ[2][2], G[2][2];
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int u = 0; u < 2; ++u) CJ[d][u] = F.C[d][0] * jmt[0][u] + F.C[d][1] * jmt[1][u];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int u = 0; u < 2; ++u) G[s2][u] = (jmt[0][s2] * CJ[0][u] + jmt[1][s2] * CJ[1][u]) * F.wsum;
    // the same order of additions as the generic contraction (s outer, u inner)
    K[1 * 3 + 1] += G[0][0]; K[1 * 3 + 2] += G[0][1]; K[2 * 3 + 1] += G[1][0]; K[2 * 3 + 2] += G[1][1];
    K[1 * 3 + 0] += -G[0][0] - G[0][1];
    K[2 * 3 + 0] += -G[1][0] - G[1][1];
    K[0 * 3 + 1] += -G[0][0] - G[1][0];
    K[0 * 3 + 2] += -G[0][1] - G[1][1];
    K[0 * 3 + 0] += ((G[0][0] + G[0][1]) + G[1][0]) + G[1][1];
  }
}

// row i of the element matrix (without |K|): (gi0, gi1) = ghat_i, selected without branches
__device__ __forceinline__ void p1_element_row(const P1Form& F, const double (&jmt)[2][2], int i, double (&K)[3]) {
  const double gi0 = i == 0 ? -1.0 : (i == 1 ? 1.0 : 0.0), gi1 = i == 0 ? -1.0 : (i == 2 ? 1.0 : 0.0);
  K[0] = K[1] = K[2] = 0.0;
  const double mi = i == 0 ? F.mhat[0] : (i == 1 ? F.mhat[1] : F.mhat[2]);
  if (F.has00) {
#pragma unroll
    for (int j = 0; j < 3; ++j) K[j] = F.c00 * (i == 0 ? F.Mhat[j] : (i == 1 ? F.Mhat[3 + j] : F.Mhat[6 + j]));
  }
  if (F.has0r) {
    const double g1 = F.c0r[0] * jmt[0][0] + F.c0r[1] * jmt[1][0], g2 = F.c0r[0] * jmt[0][1] + F.c0r[1] * jmt[1][1];
    K[0] -= mi * (g1 + g2); K[1] += mi * g1; K[2] += mi * g2;
  }
  if (F.hasr0) {
    const double g1 = F.cr0[0] * jmt[0][0] + F.cr0[1] * jmt[1][0], g2 = F.cr0[0] * jmt[0][1] + F.cr0[1] * jmt[1][1];
    const double gi = gi0 * g1 + gi1 * g2;
#pragma unroll
    for (int j = 0; j < 3; ++j) K[j] += gi * F.mhat[j];
  }
  if (F.hasrr) {
    // t = Jmt ghat_i (a vector of the physical space), r = W Jmt^T C^T ... : K_ij = W (Jmt ghat_i)^T C (Jmt ghat_j)
    const double t0 = jmt[0][0] * gi0 + jmt[0][1] * gi1, t1 = jmt[1][0] * gi0 + jmt[1][1] * gi1;   // grad psi_i
    const double u0 = (t0 * F.C[0][0] + t1 * F.C[1][0]) * F.wsum, u1 = (t0 * F.C[0][1] + t1 * F.C[1][1]) * F.wsum;   // C^T grad psi_i
    const double k1 = u0 * jmt[0][0] + u1 * jmt[1][0];   // . grad phi_1 = Jmt (1, 0)
    const double k2 = u0 * jmt[0][1] + u1 * jmt[1][1];   // . grad phi_2 = Jmt (0, 1)
    K[1] += k1; K[2] += k2; K[0] -= k1 + k2;
  }
}

}  // namespace tfelcu
