// Host-side eigen-decomposition for the MlModel native half.
// Replaces the LAPACK calls behind diagonalize_sym (lib/mlmodel.c:163-197, dsyev) and
// diagonalize_gtr (lib/mlmodel.c:208-262, dgeev + dgetrf/dgetri) with a self-contained
// cyclic Jacobi solver: n <= 64, called once per model, never on the GPU hot loop.
// Eigenvector scaling/order is solver-specific, so parity is defined
// on P(t) = U exp(Dt) Ui, not on U/D themselves.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "phylo_engine.h"

namespace {

// Cyclic Jacobi on a symmetric n*n matrix A (row-major, destroyed). V's COLUMNS are the
// eigenvectors. Returns false if it does not converge.
bool jacobi(std::vector<double> &A, int n, std::vector<double> &V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-60 || off <= 1e-34 * diag) return true;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p,q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p,q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  return false;
}

// eigenvalues ascending (dsyev's order), eigenvector columns permuted alongside
void sort_eigen(std::vector<double> &lam, std::vector<double> &V, int n) {
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return lam[a] < lam[b]; });
  std::vector<double> l2(n), V2((size_t)n * n);
  for (int c = 0; c < n; ++c) {
    l2[c] = lam[idx[c]];
    for (int r = 0; r < n; ++r) V2[(size_t)r * n + c] = V[(size_t)r * n + idx[c]];
  }
  lam.swap(l2);
  V.swap(V2);
}

// stationary distribution of a rate matrix: pi Q = 0, sum pi = 1 (Gaussian elimination with
// partial pivoting on Q^T with the last equation replaced by the normalisation)
bool stationary(const double *Q, int n, std::vector<double> &pi) {
  std::vector<double> M((size_t)n * (n + 1), 0.0);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) M[(size_t)r * (n + 1) + c] = Q[(size_t)c * n + r];
  for (int c = 0; c < n; ++c) M[(size_t)(n - 1) * (n + 1) + c] = 1.0;
  M[(size_t)(n - 1) * (n + 1) + n] = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * (n + 1) + col]) > std::fabs(M[(size_t)piv * (n + 1) + col])) piv = r;
    if (std::fabs(M[(size_t)piv * (n + 1) + col]) < 1e-300) return false;
    if (piv != col)
      for (int c = 0; c <= n; ++c) std::swap(M[(size_t)piv * (n + 1) + c], M[(size_t)col * (n + 1) + c]);
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = M[(size_t)r * (n + 1) + col] / M[(size_t)col * (n + 1) + col];
      if (f == 0.0) continue;
      for (int c = col; c <= n; ++c) M[(size_t)r * (n + 1) + c] -= f * M[(size_t)col * (n + 1) + c];
    }
  }
  pi.resize(n);
  for (int i = 0; i < n; ++i) pi[i] = M[(size_t)i * (n + 1) + n] / M[(size_t)i * (n + 1) + i];
  return true;
}

}  // namespace

extern "C" int phylo_diagonalize_sym(double *Q, double *D, int n) {
  if (!Q || !D || n < 1) return PHYLO_ERR_ARG;
  for (int i = 0; i < n * n; ++i)
    if (std::isnan(Q[i])) return PHYLO_ERR_NUMERIC;  // lib/mlModel.ml:557-568 guards NaN too
  std::vector<double> A(Q, Q + (size_t)n * n), V;
  // dsyev reads the upper triangle only (uplo='U', lib/mlmodel.c:166): symmetrise from it
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) A[(size_t)j * n + i] = A[(size_t)i * n + j];
  if (!jacobi(A, n, V)) return PHYLO_ERR_NUMERIC;
  std::vector<double> lam(n);
  for (int i = 0; i < n; ++i) lam[i] = A[(size_t)i * n + i];
  sort_eigen(lam, V, n);
  // the OCaml side sees eigenvectors as the ROWS of U (lib/mlmodel.c:155-158)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) Q[(size_t)r * n + c] = V[(size_t)c * n + r];
  std::memset(D, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; ++i) D[(size_t)i * n + i] = lam[i];
  return PHYLO_OK;
}

extern "C" int phylo_diagonalize_gtr(double *Q, double *D, double *Ui, int n) {
  if (!Q || !D || !Ui || n < 1) return PHYLO_ERR_ARG;
  for (int i = 0; i < n * n; ++i)
    if (std::isnan(Q[i])) return PHYLO_ERR_NUMERIC;
  // A rate matrix with real spectrum that is "similar to a symmetric matrix" (the
  // reference's stated precondition, lib/mlModel.ml:65-72, Keilson 1979) is time-reversible:
  // B = Pi^1/2 Q Pi^-1/2 is symmetric. Diagonalise B and map back.
  std::vector<double> pi;
  bool reversible = stationary(Q, n, pi);
  if (reversible) {
    double scale = 0.0;
    for (int i = 0; i < n * n; ++i) scale = std::max(scale, std::fabs(Q[i]));
    for (int i = 0; i < n && reversible; ++i) {
      if (!(pi[i] > 0.0)) reversible = false;
      for (int j = i + 1; j < n && reversible; ++j)
        if (std::fabs(pi[i] * Q[(size_t)i * n + j] - pi[j] * Q[(size_t)j * n + i]) > 1e-9 * scale)
          reversible = false;
    }
  }
  std::vector<double> A((size_t)n * n), V, lam(n), sq(n, 1.0);
  if (reversible) {
    for (int i = 0; i < n; ++i) sq[i] = std::sqrt(pi[i]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        // average the two mirror entries so B is exactly symmetric
        const double bij = sq[i] * Q[(size_t)i * n + j] / sq[j];
        const double bji = sq[j] * Q[(size_t)j * n + i] / sq[i];
        A[(size_t)i * n + j] = 0.5 * (bij + bji);
      }
  } else {
    // not a reversible generator: accept only an exactly symmetric matrix
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        if (std::fabs(Q[(size_t)i * n + j] - Q[(size_t)j * n + i]) > 1e-12 * (1.0 + std::fabs(Q[(size_t)i * n + j])))
          return PHYLO_ERR_NUMERIC;  // reference: "Imaginary eigenvalues" / QR failure class
        A[(size_t)i * n + j] = Q[(size_t)i * n + j];
      }
  }
  if (!jacobi(A, n, V)) return PHYLO_ERR_NUMERIC;
  for (int i = 0; i < n; ++i) lam[i] = A[(size_t)i * n + i];
  sort_eigen(lam, V, n);
  // Q = U D Ui with U = Pi^-1/2 V, Ui = V^T Pi^1/2 (row-major)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      Q[(size_t)r * n + c] = V[(size_t)r * n + c] / sq[r];
      Ui[(size_t)r * n + c] = V[(size_t)c * n + r] * sq[c];
    }
  std::memset(D, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; ++i) D[(size_t)i * n + i] = lam[i];
  return PHYLO_OK;
}
