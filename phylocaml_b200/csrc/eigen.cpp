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

// ---- general real matrices with a real spectrum (non-reversible generators) ----------------
// The reference hands any rate matrix to dgeev and only rejects complex eigenvalues
// (lib/mlmodel.c:246-250); Const / m_file / m_custom models (lib/mlModel.ml:473-520) can be
// non-reversible. Path for those: Householder reduction to Hessenberg form, single-shift QR with
// Givens rotations to a real Schur form T (upper triangular when the spectrum is real; an
// irreducible 2x2 block with a negative discriminant means complex eigenvalues -> rejected like
// the reference), eigenvectors of T by back-substitution, V = Z X, Ui = V^-1 by Gauss-Jordan.
// n <= 64, once per model.

// A (row-major n*n) -> Hessenberg H in place, Z accumulates the transformations: A = Z H Z^T
void hessenberg(std::vector<double> &A, int n, std::vector<double> &Z) {
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
  std::vector<double> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double alpha = 0.0;
    for (int i = k + 1; i < n; ++i) alpha += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    alpha = std::sqrt(alpha);
    if (alpha == 0.0) continue;
    if (A[(size_t)(k + 1) * n + k] > 0.0) alpha = -alpha;
    double vn = 0.0;
    for (int i = 0; i < n; ++i) v[i] = 0.0;
    v[k + 1] = A[(size_t)(k + 1) * n + k] - alpha;
    for (int i = k + 2; i < n; ++i) v[i] = A[(size_t)i * n + k];
    for (int i = k + 1; i < n; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    // A <- (I - 2 v v^T / vn) A (I - 2 v v^T / vn), Z <- Z (I - 2 v v^T / vn)
    for (int c = 0; c < n; ++c) {
      double d = 0.0;
      for (int i = k + 1; i < n; ++i) d += v[i] * A[(size_t)i * n + c];
      d = 2.0 * d / vn;
      for (int i = k + 1; i < n; ++i) A[(size_t)i * n + c] -= d * v[i];
    }
    for (int r = 0; r < n; ++r) {
      double d = 0.0, dz = 0.0;
      for (int i = k + 1; i < n; ++i) { d += A[(size_t)r * n + i] * v[i]; dz += Z[(size_t)r * n + i] * v[i]; }
      d = 2.0 * d / vn; dz = 2.0 * dz / vn;
      for (int i = k + 1; i < n; ++i) { A[(size_t)r * n + i] -= d * v[i]; Z[(size_t)r * n + i] -= dz * v[i]; }
    }
    for (int i = k + 2; i < n; ++i) A[(size_t)i * n + k] = 0.0;
  }
}

// Hessenberg H -> upper triangular T in place by shifted QR steps (Givens), Z updated so that
// Z_in H Z_in^T = Z_out T Z_out^T. Returns 0 ok, 1 complex eigenvalues, 2 no convergence.
int schur_real(std::vector<double> &H, int n, std::vector<double> &Z) {
  double norm = 0.0;
  for (size_t i = 0; i < H.size(); ++i) norm = std::max(norm, std::fabs(H[i]));
  if (norm == 0.0) return 0;
  const double eps = 2.2e-16;
  int hi = n - 1, iter = 0;
  std::vector<double> cs(n), sn(n);
  while (hi > 0) {
    // deflate negligible subdiagonals
    int lo = hi;
    while (lo > 0) {
      const double s = std::fabs(H[(size_t)(lo - 1) * n + lo - 1]) + std::fabs(H[(size_t)lo * n + lo]);
      if (std::fabs(H[(size_t)lo * n + lo - 1]) <= eps * (s == 0.0 ? norm : s)) { H[(size_t)lo * n + lo - 1] = 0.0; break; }
      --lo;
    }
    if (lo == hi) { --hi; iter = 0; continue; }
    // trailing 2x2 block of the active window
    const double a = H[(size_t)(hi - 1) * n + hi - 1], b = H[(size_t)(hi - 1) * n + hi], c = H[(size_t)hi * n + hi - 1],
                 d = H[(size_t)hi * n + hi];
    const double tr = a + d, det = a * d - b * c, disc = 0.25 * tr * tr - det;
    if (lo == hi - 1 && disc < 0.0) {
      // isolated 2x2 block: complex pair unless the discriminant is rounding noise
      if (-disc > 1e-14 * (0.25 * tr * tr + std::fabs(det) + norm * norm * 1e-16)) return 1;
    }
    if (++iter > 60 * n) return disc < 0.0 ? 1 : 2;
    double mu;
    if (disc >= 0.0) {  // Wilkinson: the eigenvalue of the block closer to d
      const double r = std::sqrt(disc), l1 = 0.5 * tr + r, l2 = 0.5 * tr - r;
      mu = std::fabs(l1 - d) < std::fabs(l2 - d) ? l1 : l2;
    } else {
      mu = d;
    }
    if (iter % 11 == 10) mu += 0.75 * (std::fabs(c) + std::fabs(H[(size_t)(hi - 1) * n + (hi >= 2 ? hi - 2 : 0)]));  // exceptional shift
    // QR step on rows/cols lo..hi: H - mu I = Q R, H <- R Q + mu I (applied to the full matrix)
    for (int i = lo; i <= hi; ++i) H[(size_t)i * n + i] -= mu;
    for (int k = lo; k < hi; ++k) {
      const double x = H[(size_t)k * n + k], y = H[(size_t)(k + 1) * n + k];
      const double r = std::hypot(x, y);
      double cg = 1.0, sg = 0.0;
      if (r != 0.0) { cg = x / r; sg = y / r; }
      cs[k] = cg; sn[k] = sg;
      for (int col = k; col < n; ++col) {  // rows k, k+1
        const double u = H[(size_t)k * n + col], w = H[(size_t)(k + 1) * n + col];
        H[(size_t)k * n + col] = cg * u + sg * w;
        H[(size_t)(k + 1) * n + col] = -sg * u + cg * w;
      }
    }
    for (int k = lo; k < hi; ++k) {
      const double cg = cs[k], sg = sn[k];
      for (int row = 0; row <= std::min(hi, k + 2); ++row) {  // columns k, k+1
        const double u = H[(size_t)row * n + k], w = H[(size_t)row * n + k + 1];
        H[(size_t)row * n + k] = cg * u + sg * w;
        H[(size_t)row * n + k + 1] = -sg * u + cg * w;
      }
      for (int row = 0; row < n; ++row) {
        const double u = Z[(size_t)row * n + k], w = Z[(size_t)row * n + k + 1];
        Z[(size_t)row * n + k] = cg * u + sg * w;
        Z[(size_t)row * n + k + 1] = -sg * u + cg * w;
      }
    }
    for (int i = lo; i <= hi; ++i) H[(size_t)i * n + i] += mu;
  }
  return 0;
}

// right eigenvectors (columns of V) and eigenvalues of a general real matrix with real spectrum
int eigen_general(const double *Q, int n, std::vector<double> &lam, std::vector<double> &V) {
  std::vector<double> H(Q, Q + (size_t)n * n), Z;
  hessenberg(H, n, Z);
  const int rc = schur_real(H, n, Z);
  if (rc != 0) return rc;
  double norm = 0.0;
  for (size_t i = 0; i < H.size(); ++i) norm = std::max(norm, std::fabs(H[i]));
  const double tiny = std::max(norm, 1e-300) * 2.2e-16;
  lam.resize(n);
  for (int i = 0; i < n; ++i) lam[i] = H[(size_t)i * n + i];
  // eigenvectors of the triangular T: X[k][k] = 1, back-substitution above the diagonal
  std::vector<double> X((size_t)n * n, 0.0);
  for (int k = 0; k < n; ++k) {
    X[(size_t)k * n + k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      double sum = 0.0;
      for (int j = i + 1; j <= k; ++j) sum += H[(size_t)i * n + j] * X[(size_t)j * n + k];
      double den = H[(size_t)i * n + i] - lam[k];
      if (std::fabs(den) < tiny) den = tiny;  // repeated eigenvalue: the perturbation LAPACK's dtrevc uses
      X[(size_t)i * n + k] = -sum / den;
    }
  }
  V.assign((size_t)n * n, 0.0);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      double sacc = 0.0;
      for (int j = 0; j <= c; ++j) sacc += Z[(size_t)r * n + j] * X[(size_t)j * n + c];
      V[(size_t)r * n + c] = sacc;
    }
  for (int c = 0; c < n; ++c) {  // unit columns
    double nn = 0.0;
    for (int r = 0; r < n; ++r) nn += V[(size_t)r * n + c] * V[(size_t)r * n + c];
    nn = std::sqrt(nn);
    if (!(nn > 0.0) || !std::isfinite(nn)) return 2;
    for (int r = 0; r < n; ++r) V[(size_t)r * n + c] /= nn;
  }
  return 0;
}

// Gauss-Jordan inverse with partial pivoting (dgetrf/dgetri's job, lib/mlmodel.c:92-124)
bool invert(const std::vector<double> &M, int n, double *inv) {
  std::vector<double> W((size_t)n * 2 * n, 0.0);
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) W[(size_t)r * 2 * n + c] = M[(size_t)r * n + c];
    W[(size_t)r * 2 * n + n + r] = 1.0;
  }
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (std::fabs(W[(size_t)r * 2 * n + col]) > std::fabs(W[(size_t)piv * 2 * n + col])) piv = r;
    const double p = W[(size_t)piv * 2 * n + col];
    if (std::fabs(p) < 1e-13) return false;  // eigenvectors (unit columns) not independent: defective matrix
    if (piv != col)
      for (int c = 0; c < 2 * n; ++c) std::swap(W[(size_t)piv * 2 * n + c], W[(size_t)col * 2 * n + c]);
    for (int c = 0; c < 2 * n; ++c) W[(size_t)col * 2 * n + c] /= p;
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = W[(size_t)r * 2 * n + col];
      if (f == 0.0) continue;
      for (int c = 0; c < 2 * n; ++c) W[(size_t)r * 2 * n + c] -= f * W[(size_t)col * 2 * n + c];
    }
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) inv[(size_t)r * n + c] = W[(size_t)r * 2 * n + n + c];
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
    bool symmetric = true;
    for (int i = 0; i < n && symmetric; ++i)
      for (int j = 0; j < n; ++j)
        if (std::fabs(Q[(size_t)i * n + j] - Q[(size_t)j * n + i]) > 1e-12 * (1.0 + std::fabs(Q[(size_t)i * n + j]))) {
          symmetric = false;
          break;
        }
    if (!symmetric) {
      // a non-reversible generator (Const / m_file / m_custom, lib/mlModel.ml:473-520): general
      // real eigen-solver; complex eigenvalues are an error exactly as in lib/mlmodel.c:248-250
      std::vector<double> lam_g, Vg;
      if (eigen_general(Q, n, lam_g, Vg) != 0) return PHYLO_ERR_NUMERIC;
      sort_eigen(lam_g, Vg, n);
      std::vector<double> inv((size_t)n * n);
      if (!invert(Vg, n, inv.data())) return PHYLO_ERR_NUMERIC;
      std::memcpy(Q, Vg.data(), sizeof(double) * (size_t)n * n);  // U: columns are the eigenvectors
      std::memcpy(Ui, inv.data(), sizeof(double) * (size_t)n * n);
      std::memset(D, 0, sizeof(double) * (size_t)n * n);
      for (int i = 0; i < n; ++i) D[(size_t)i * n + i] = lam_g[i];
      return PHYLO_OK;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = Q[(size_t)i * n + j];
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
