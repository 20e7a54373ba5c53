// es_kernel_host.cpp -- plan-time, double precision, host side:
//   * exponential-of-semicircle kernel parameters (finufft's published choices; the reference
//     reaches them through finufft.Plan(..., eps) at
//     src/mrinufft/operators/interfaces/finufft.py:43-50),
//   * piecewise polynomial (one polynomial per tap) for the device Horner evaluator,
//   * deapodisation vectors (-1)^k / phihat(k) from the kernel's continuous Fourier transform.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* b200_last_error(void) { return g_err; }

int next235even(int n) {
  if (n < 2) n = 2;
  if (n & 1) ++n;
  for (;; n += 2) {
    int m = n;
    while (m % 2 == 0) m /= 2;
    while (m % 3 == 0) m /= 3;
    while (m % 5 == 0) m /= 5;
    if (m == 1) return n;
  }
}

void es_kernel_params(double eps, double sigma, int* w_out, double* beta_out) {
  int w;
  double s = sigma < 1.001 ? 1.001 : sigma;
  if (sigma == 2.0)
    w = (int)std::ceil(std::log10(10.0 / eps));
  else
    w = (int)std::ceil(-std::log(eps) / (M_PI * std::sqrt(1.0 - 1.0 / s)));
  if (w < 2) w = 2;
  if (w > B200_MAX_W) w = B200_MAX_W;
  double bon;
  if (sigma == 2.0) {
    bon = 2.30;
    if (w == 2) bon = 2.20;
    if (w == 3) bon = 2.26;
    if (w == 4) bon = 2.38;
  } else {
    bon = 0.97 * M_PI * (1.0 - 1.0 / (2.0 * s));
  }
  *w_out = w;
  *beta_out = bon * w;
}

static inline double es_phi(double x, int w, double beta) {
  double a = 1.0 - (2.0 * x / w) * (2.0 * x / w);
  if (a < 0.0) return 0.0;
  return std::exp(beta * (std::sqrt(a) - 1.0));
}

// Chebyshev interpolant of degree `deg` on z in [-1,1] for tap i, converted to monomials.
static void fit_tap(int w, double beta, int deg, int i, std::vector<double>* mono) {
  const int n = deg + 1;
  std::vector<double> f(n), ck(n);
  for (int j = 0; j < n; ++j) {
    double z = std::cos(M_PI * (j + 0.5) / n);
    double x = 0.5 * (z + 1.0 - w) + i;
    f[j] = es_phi(x, w, beta);
  }
  for (int k = 0; k < n; ++k) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += f[j] * std::cos(k * M_PI * (j + 0.5) / n);
    ck[k] = 2.0 * s / n;
  }
  ck[0] *= 0.5;
  // Chebyshev series -> monomial coefficients through T_{k+1} = 2 z T_k - T_{k-1}
  std::vector<double> tkm1(n, 0.0), tk(n, 0.0), tkp1(n, 0.0);
  mono->assign(n, 0.0);
  tkm1[0] = 1.0;  // T0
  for (int m = 0; m < n; ++m) (*mono)[m] += ck[0] * tkm1[m];
  if (n > 1) {
    tk[1] = 1.0;  // T1
    for (int m = 0; m < n; ++m) (*mono)[m] += ck[1] * tk[m];
  }
  for (int k = 2; k < n; ++k) {
    for (int m = 0; m < n; ++m) tkp1[m] = -tkm1[m] + (m > 0 ? 2.0 * tk[m - 1] : 0.0);
    for (int m = 0; m < n; ++m) (*mono)[m] += ck[k] * tkp1[m];
    tkm1 = tk;
    tk = tkp1;
  }
}

static double fit_error(int w, double beta, int deg, const std::vector<std::vector<double>>& mono) {
  double err = 0;
  const int ns = 400;
  for (int i = 0; i < w; ++i)
    for (int s = 0; s < ns; ++s) {
      double z = -1.0 + 2.0 * (s + 0.5) / ns;
      double acc = mono[i][deg];
      for (int k = deg - 1; k >= 0; --k) acc = acc * z + mono[i][k];
      double x = 0.5 * (z + 1.0 - w) + i;
      double e = std::fabs(acc - es_phi(x, w, beta));
      if (e > err) err = e;
    }
  return err;
}

void es_fit_polynomial(int w, double beta, double eps, KernelTables* out) {
  // smallest degree whose max abs error (peak of phi is 1) is below the float32 floor or
  // 2% of the requested tolerance; the sqrt end-point singularity makes convergence slow
  // past ~5e-8, so stop there.
  double tol = std::fmax(0.02 * eps, 6e-8);
  std::vector<std::vector<double>> mono(w);
  int deg = 4;
  for (; deg <= B200_MAX_DEG; ++deg) {
    for (int i = 0; i < w; ++i) fit_tap(w, beta, deg, i, &mono[i]);
    if (fit_error(w, beta, deg, mono) <= tol) break;
  }
  if (deg > B200_MAX_DEG) {
    deg = B200_MAX_DEG;
    for (int i = 0; i < w; ++i) fit_tap(w, beta, deg, i, &mono[i]);
  }
  out->w = w;
  out->deg = deg;
  out->beta = beta;
  out->poly.assign((size_t)(deg + 1) * w, 0.f);
  for (int k = 0; k <= deg; ++k)
    for (int i = 0; i < w; ++i) out->poly[(size_t)k * w + i] = (float)mono[i][k];
}

// Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n.
static void gauss_legendre(int n, std::vector<double>* x, std::vector<double>* wt) {
  x->resize(n);
  wt->resize(n);
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
    double pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = z;
      for (int k = 2; k <= n; ++k) {
        double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      pp = n * (z * p1 - p0) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    // recompute derivative at the converged node
    double p0 = 1.0, p1 = z;
    for (int k = 2; k <= n; ++k) {
      double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
      p0 = p1;
      p1 = p2;
    }
    pp = n * (z * p1 - p0) / (z * z - 1.0);
    (*x)[i] = z;
    (*wt)[i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

// phihat(k) = int phi(x) cos(2 pi k x / nf) dx with x = (w/2) sin(theta) (removes the
// square-root end-point singularity), Gauss-Legendre on theta in [0, pi/2].
void es_deapod_vector(int n, int nf, int w, double beta, std::vector<float>* out) {
  const int nq = 96;
  std::vector<double> t, wt;
  gauss_legendre(nq, &t, &wt);
  std::vector<double> f(nq), sx(nq);
  const double h = 0.5 * w;
  for (int q = 0; q < nq; ++q) {
    double th = (t[q] + 1.0) * (M_PI / 4.0);
    f[q] = std::exp(beta * (std::cos(th) - 1.0)) * std::cos(th) * h * wt[q] * (M_PI / 4.0);
    sx[q] = h * std::sin(th);
  }
  out->resize(n);
  for (int i = 0; i < n; ++i) {
    int k = i - n / 2;
    double s = 0;
    for (int q = 0; q < nq; ++q) s += f[q] * std::cos(2.0 * M_PI * k * sx[q] / nf);
    double ph = 2.0 * s;
    double sign = (k % 2 == 0) ? 1.0 : -1.0;
    (*out)[i] = (float)(sign / ph);
  }
}
