// 3x3 normal equations on the device: minimum-norm least-squares solution with the decision rule of the host solver
// (mcre/lsm.py:solve_normal_equations), shared by the exposure regression of linear products (irc_solve.cu) and the
// Longstaff-Schwartz backward induction (lsm.cu).  Takes the place of torch.linalg.lstsq (driver gelsy) of
// controller.py:368-374.
#pragma once

namespace mcre {

// Eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi rotations: a -> diagonal, v -> eigenvectors (columns).
__device__ inline void jacobi3(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// Minimum-norm least-squares solution of G c = rhs (G: Gram matrix of [1, u, u^2] from the sums m[0..4] = sum u^k).
__device__ inline void solve_normal_equations_dev(const double *m, const double *rhs, double *c) {
  double G[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[i][j] = m[i + j];
  c[0] = c[1] = c[2] = 0.0;
  bool finite = true;
  for (int k = 0; k < 5; ++k) finite = finite && isfinite(m[k]);
  double d[3], scale[3];
  for (int i = 0; i < 3; ++i) { d[i] = sqrt(fmax(G[i][i], 0.0)); scale[i] = d[i] > 0.0 ? d[i] : 1.0; }
  if (!finite || d[0] == 0.0) return;
  if (rhs[0] == 0.0 && rhs[1] == 0.0 && rhs[2] == 0.0) return;   // zero solution in every branch
  bool full = d[1] > 0.0 && d[2] > 0.0;
  double Gs[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Gs[i][j] = G[i][j] / (scale[i] * scale[j]);
  if (full) {
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) a[i][j] = Gs[i][j];
    jacobi3(a, v);
    const double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
    const double lmin = fmin(fabs(a[0][0]), fmin(fabs(a[1][1]), fabs(a[2][2])));
    full = lmin > 1e-10 * lmax;
  }
  if (full) {
    // Gaussian elimination with partial pivoting on the equilibrated system
    double A[3][4];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) A[i][j] = Gs[i][j];
      A[i][3] = rhs[i] / scale[i];
    }
    for (int k = 0; k < 3; ++k) {
      int piv = k;
      for (int i = k + 1; i < 3; ++i)
        if (fabs(A[i][k]) > fabs(A[piv][k])) piv = i;
      if (piv != k)
        for (int j = 0; j < 4; ++j) { const double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
      for (int i = k + 1; i < 3; ++i) {
        const double f = A[i][k] / A[k][k];
        for (int j = k; j < 4; ++j) A[i][j] -= f * A[k][j];
      }
    }
    double y[3];
    for (int i = 2; i >= 0; --i) {
      double s = A[i][3];
      for (int j = i + 1; j < 3; ++j) s -= A[i][j] * y[j];
      y[i] = s / A[i][i];
    }
    for (int i = 0; i < 3; ++i) c[i] = y[i] / scale[i];
    return;
  }
  // rank deficient (constant regressor): minimum norm in the unscaled coefficients, like gelsy
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = G[i][j];
  jacobi3(a, v);
  const double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
  for (int e = 0; e < 3; ++e) {
    const double lam = a[e][e];
    if (!(lam > 1e-10 * lmax)) continue;
    const double proj = (v[0][e] * rhs[0] + v[1][e] * rhs[1] + v[2][e] * rhs[2]) / lam;
    for (int i = 0; i < 3; ++i) c[i] += v[i][e] * proj;
  }
}

}  // namespace mcre
