// P2 (quadratic Lagrange) element matrices in closed form (included inside an anonymous namespace by dfe_general.cu).
//
// The reference has P1 elements only; P2 is an unchecked item of its roadmap (reference README.md:139-143, SURVEY §8(f)
// N4), so there is no upstream arithmetic to reproduce bit for bit: these kernels are checked against the quadrature-based
// oracle (oracle/oracle_p2.py) and against closed-form solutions.  Conventions (shared with the oracle): 1-D element
// [left, right, mid], 2-D element [v0, v1, v2, m01, m12, m20]; affine geometry read from the vertices; load F = M f with the
// consistent mass matrix; triangles with area < 1e-15 contribute nothing (solver.py:120-121).
//
// 1-D, h = x_right - x_left:   K0 = 1/(3h) [[7, 1, -8], [1, 7, -8], [-8, -8, 16]],   M = h/30 [[4, -1, 2], [-1, 4, 2], [2, 2, 16]].
// 2-D, with S_ij = (b_i b_j + c_i c_j) / (4 area) (the P1 element matrix at kappa = 1, solver.py:125-139) and edges
// e_3 = (0, 1), e_4 = (1, 2), e_5 = (2, 0):
//   K0(v_i, v_i) = S_ii,  K0(v_i, v_j) = -S_ij / 3,
//   K0(v_i, e) = 4/3 S_ik if e = (i, k) or (k, i), 0 if e is the edge opposite to v_i,
//   K0((a, b), (c, d)) = 4/3 [S_bd w(a, c) + S_bc w(a, d) + S_ad w(b, c) + S_ac w(b, d)],  w(x, y) = 2 if x == y else 1,
//   M = area/180 * T (table below).
#pragma once

__device__ __forceinline__ double p2_line_k0(int i, int j) {   // times 1 / (3 h)
  if (i == 2 && j == 2) return 16.0;
  if (i == 2 || j == 2) return -8.0;
  return i == j ? 7.0 : 1.0;
}
__device__ __forceinline__ double p2_line_m(int i, int j) {    // times h / 30
  if (i == 2 && j == 2) return 16.0;
  if (i == 2 || j == 2) return 2.0;
  return i == j ? 4.0 : -1.0;
}

struct P2Tri {
  double S[3][3];   // P1 element matrix at kappa = 1
  double area;
  bool keep;        // area >= 1e-15
};

__device__ __forceinline__ P2Tri p2_tri_geom(const double (&x)[3], const double (&y)[3]) {
  P2Tri T;
  const double b[3] = {y[1] - y[2], y[2] - y[0], y[0] - y[1]};
  const double c[3] = {x[2] - x[1], x[0] - x[2], x[1] - x[0]};
  T.area = 0.5 * fabs(c[2] * b[1] - c[1] * b[2]);
  T.keep = !(T.area < 1e-15);
  const double r = T.keep ? 1.0 / (4.0 * T.area) : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T.S[i][j] = (b[i] * b[j] + c[i] * c[j]) * r;
  return T;
}

// entry (I, J) of K0 for compile-time indices
template <int I, int J>
__device__ __forceinline__ double p2_tri_k0(const P2Tri& T) {
  constexpr int EA[3] = {0, 1, 2}, EB[3] = {1, 2, 0};   // endpoints of edge 3 + k
  if constexpr (I < 3 && J < 3) {
    return I == J ? T.S[I][I] : -T.S[I][J] * (1.0 / 3.0);
  } else if constexpr (I < 3 || J < 3) {
    constexpr int v = I < 3 ? I : J, e = (I < 3 ? J : I) - 3;
    if constexpr (EA[e] == v) return (4.0 / 3.0) * T.S[v][EB[e]];
    else if constexpr (EB[e] == v) return (4.0 / 3.0) * T.S[v][EA[e]];
    else return 0.0;
  } else {
    constexpr int a = EA[I - 3], b = EB[I - 3], c = EA[J - 3], d = EB[J - 3];
    return (4.0 / 3.0) * (T.S[b][d] * (a == c ? 2.0 : 1.0) + T.S[b][c] * (a == d ? 2.0 : 1.0) +
                          T.S[a][d] * (b == c ? 2.0 : 1.0) + T.S[a][c] * (b == d ? 2.0 : 1.0));
  }
}
template <int I>
__device__ __forceinline__ void p2_tri_k0_row(const P2Tri& T, double (&row)[6]) {
  row[0] = p2_tri_k0<I, 0>(T);
  row[1] = p2_tri_k0<I, 1>(T);
  row[2] = p2_tri_k0<I, 2>(T);
  row[3] = p2_tri_k0<I, 3>(T);
  row[4] = p2_tri_k0<I, 4>(T);
  row[5] = p2_tri_k0<I, 5>(T);
}
__device__ __forceinline__ void p2_tri_k0_row(const P2Tri& T, int loc, double (&row)[6]) {
  switch (loc) {
    case 0: p2_tri_k0_row<0>(T, row); break;
    case 1: p2_tri_k0_row<1>(T, row); break;
    case 2: p2_tri_k0_row<2>(T, row); break;
    case 3: p2_tri_k0_row<3>(T, row); break;
    case 4: p2_tri_k0_row<4>(T, row); break;
    default: p2_tri_k0_row<5>(T, row); break;
  }
}
// row `loc` of the mass matrix, times 180 / area
__device__ __forceinline__ void p2_tri_m_row(int loc, double (&row)[6]) {
  if (loc < 3) {
#pragma unroll
    for (int j = 0; j < 3; ++j) row[j] = j == loc ? 6.0 : -1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) row[3 + k] = (k == (loc + 1) % 3) ? -4.0 : 0.0;   // the edge opposite to vertex loc
  } else {
    const int k = loc - 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) row[j] = (k == (j + 1) % 3) ? -4.0 : 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l) row[3 + l] = l == k ? 32.0 : 16.0;
  }
}
