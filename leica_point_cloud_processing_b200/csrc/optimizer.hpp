// optimizer.hpp - host-side BFGS with Fletcher line search, the outer-loop optimiser of GICP.
//
// BASELINE.json north_star keeps the BFGS / line-search loop in C++ on the host so that convergence follows the
// reference's optimiser: pcl::GeneralizedIterativeClosestPoint::estimateRigidTransformationBFGS driving PCL's
// BFGS<> (registration/bfgs.h), itself a port of GSL's vector_bfgs2 minimiser and linear_minimize.c
// (reached from gicp_.align(), reference src/GICPAlignment.cpp:96,116).
//
// Difference in structure, not in arithmetic: every probe of the line function runs ONE fused GPU evaluation that
// returns f and the full gradient, and the probe is memoised by its step length, so the "f then df at the same
// alpha" pattern of the line search costs one kernel launch instead of two passes over the pairs.
#pragma once
#include <cmath>
#include <functional>
#include <limits>

namespace gicpb {

enum class BfgsStatus { kRunning = -1, kSuccess = 0, kNoProgress = 1, kEvalFailed = 2 };

class Bfgs6 {
 public:
  // evaluates f and g[6] at x[6]; returns false if the evaluation could not be carried out
  using Eval = std::function<bool(const double* x, double* f, double* g)>;

  struct Options {
    double rho = 0.01, sigma = 0.01, tau1 = 9.0, tau2 = 0.05, tau3 = 0.5, step_size = 1.0;
    int order = 3, bracket_iters = 100, section_iters = 100;
  };

  explicit Bfgs6(Eval eval) : eval_(std::move(eval)) {}
  Options opt;
  long evaluations = 0;

  bool init(const double* x) {
    for (int i = 0; i < 6; ++i) x0_[i] = x[i];
    if (!evaluate(x0_, &f_, g0_)) return false;
    delta_f_ = 0.0;
    g0norm_ = norm(g0_);
    for (int i = 0; i < 6; ++i) p_[i] = g0_[i] * -1 / g0norm_;
    pnorm_ = norm(p_);
    fp0_ = -g0norm_;
    remember(0.0, x0_, f_, g0_);
    return true;
  }

  // one BFGS iteration; x is updated in place on success
  BfgsStatus step(double* x) {
    const double f0 = f_;
    if (pnorm_ == 0.0 || g0norm_ == 0.0 || fp0_ == 0.0) return BfgsStatus::kNoProgress;
    double alpha1;
    if (delta_f_ < 0) {
      const double del = std::max(-delta_f_, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
      alpha1 = std::min(1.0, 2.0 * del / (-fp0_));
    } else {
      alpha1 = std::fabs(opt.step_size);
    }
    double alpha = 0.0;
    const BfgsStatus ls = line_search(alpha1, &alpha);
    if (ls != BfgsStatus::kSuccess) return ls;

    if (!probe(alpha)) return BfgsStatus::kEvalFailed;  // memoised: the accepted point was just evaluated
    double g1[6];
    for (int i = 0; i < 6; ++i) {
      x[i] = cache_x_[i];
      g1[i] = cache_g_[i];
    }
    f_ = cache_f_;
    delta_f_ = f_ - f0;

    // memoryless BFGS direction: p' = g1 - A dx - B dg
    double dx[6], dg[6];
    for (int i = 0; i < 6; ++i) {
      dx[i] = x[i] - x0_[i];
      dg[i] = g1[i] - g0_[i];
    }
    const double dxg = dot(dx, g1), dgg = dot(dg, g1), dxdg = dot(dx, dg), dgnorm = norm(dg);
    double A = 0.0, B = 0.0;
    if (dxdg != 0) {
      B = dxg / dxdg;
      A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
    }
    for (int i = 0; i < 6; ++i) p_[i] = -A * dx[i];
    for (int i = 0; i < 6; ++i) p_[i] += g1[i];
    for (int i = 0; i < 6; ++i) p_[i] += -B * dg[i];
    for (int i = 0; i < 6; ++i) {
      g0_[i] = g1[i];
      x0_[i] = x[i];
    }
    g0norm_ = norm(g0_);
    pnorm_ = norm(p_);
    const double dir = (dot(p_, g1) > 0) ? -1.0 : 1.0;
    for (int i = 0; i < 6; ++i) p_[i] *= dir / pnorm_;
    pnorm_ = norm(p_);
    fp0_ = dot(p_, g0_);
    remember(0.0, x0_, f_, g0_);  // alpha = 0 of the new line is the current point
    return BfgsStatus::kSuccess;
  }

  double gradient_norm() const { return g0norm_; }
  double value() const { return f_; }
  bool eval_failed() const { return eval_failed_; }

 private:
  static double dot(const double* a, const double* b) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += a[i] * b[i];
    return s;
  }
  static double norm(const double* a) { return std::sqrt(dot(a, a)); }

  bool evaluate(const double* x, double* f, double* g) {
    ++evaluations;
    if (!eval_(x, f, g)) {
      eval_failed_ = true;
      return false;
    }
    return true;
  }
  void remember(double alpha, const double* x, double f, const double* g) {
    cache_alpha_ = alpha;
    cache_f_ = f;
    for (int i = 0; i < 6; ++i) {
      cache_x_[i] = x[i];
      cache_g_[i] = g[i];
    }
    cache_df_ = dot(cache_g_, p_);
  }
  // phi(alpha) = f(x0 + alpha p), phi'(alpha) = g . p ; one fused evaluation per distinct alpha
  bool probe(double alpha) {
    if (alpha == cache_alpha_) return true;
    double xa[6], fa, ga[6];
    for (int i = 0; i < 6; ++i) xa[i] = x0_[i] + alpha * p_[i];
    if (!evaluate(xa, &fa, ga)) return false;
    remember(alpha, xa, fa, ga);
    return true;
  }

  static int solve_quadratic(double a, double b, double c, double* x0, double* x1) {
    if (a == 0) {
      if (b == 0) return 0;
      *x0 = -c / b;
      return 1;
    }
    const double disc = b * b - 4 * a * c;
    if (disc > 0) {
      if (b == 0) {
        const double r = std::sqrt(-c / a);
        *x0 = -r;
        *x1 = r;
      } else {
        const double sgnb = (b > 0 ? 1 : -1);
        const double temp = -0.5 * (b + sgnb * std::sqrt(disc));
        const double r1 = temp / a, r2 = c / temp;
        *x0 = r1 < r2 ? r1 : r2;
        *x1 = r1 < r2 ? r2 : r1;
      }
      return 2;
    }
    if (disc == 0) {
      *x0 = *x1 = -0.5 * b / a;
      return 2;
    }
    return 0;
  }
  static double poly3(double c0, double c1, double c2, double c3, double z) { return c0 + z * (c1 + z * (c2 + z * c3)); }
  static double interp_quadratic(double f0, double fp0, double f1, double zl, double zh) {
    const double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
    const double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
    const double c = 2 * (f1 - f0 - fp0);
    double zmin = zl, fmin = fl;
    if (fh < fmin) {
      zmin = zh;
      fmin = fh;
    }
    if (c > 0) {
      const double z = -fp0 / c;
      if (z > zl && z < zh) {
        const double f = f0 + z * (fp0 + z * (f1 - f0 - fp0));
        if (f < fmin) zmin = z;
      }
    }
    return zmin;
  }
  static double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh) {
    const double eta = 3 * (f1 - f0) - 2 * fp0 - fp1;
    const double xi = fp0 + fp1 - 2 * (f1 - f0);
    double zmin = zl, fmin = poly3(f0, fp0, eta, xi, zl);
    auto consider = [&](double z) {
      const double y = poly3(f0, fp0, eta, xi, z);
      if (y < fmin) {
        zmin = z;
        fmin = y;
      }
    };
    consider(zh);
    double z0 = 0, z1 = 0;
    const int n = solve_quadratic(3 * xi, 2 * eta, fp0, &z0, &z1);
    if (n >= 1 && z0 > zl && z0 < zh) consider(z0);
    if (n == 2 && z1 > zl && z1 < zh) consider(z1);
    return zmin;
  }
  double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax) const {
    double zmin = (xmin - a) / (b - a), zmax = (xmax - a) / (b - a);
    if (zmin > zmax) std::swap(zmin, zmax);
    const double z = (opt.order > 2 && !std::isnan(fpb)) ? interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zmin, zmax)
                                                         : interp_quadratic(fa, fpa * (b - a), fb, zmin, zmax);
    return a + z * (b - a);
  }

  // GSL linear_minimize.c `minimize`: bracketing then sectioning with Fletcher's rho / sigma tests
  BfgsStatus line_search(double alpha1, double* alpha_new) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (!probe(0.0)) return BfgsStatus::kEvalFailed;
    const double f0 = cache_f_, fp0 = cache_df_;
    double falpha, fpalpha, falpha_prev = f0, fpalpha_prev = fp0;
    double alpha = alpha1, alpha_prev = 0.0;
    double a = 0.0, b = alpha, fa = f0, fb = 0.0, fpa = fp0, fpb = 0.0;
    int i = 0;
    while (i++ < opt.bracket_iters) {
      if (!probe(alpha)) return BfgsStatus::kEvalFailed;
      falpha = cache_f_;
      if (falpha > f0 + alpha * opt.rho * fp0 || falpha >= falpha_prev) {
        a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
        b = alpha; fb = falpha; fpb = nan;
        break;
      }
      fpalpha = cache_df_;
      if (std::fabs(fpalpha) <= -opt.sigma * fp0) {
        *alpha_new = alpha;
        return BfgsStatus::kSuccess;
      }
      if (fpalpha >= 0) {
        a = alpha; fa = falpha; fpa = fpalpha;
        b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
        break;
      }
      const double delta = alpha - alpha_prev;
      const double alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha,
                                            alpha + delta, alpha + opt.tau1 * delta);
      alpha_prev = alpha;
      falpha_prev = falpha;
      fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    while (i++ < opt.section_iters) {
      const double delta = b - a;
      alpha = interpolate(a, fa, fpa, b, fb, fpb, a + opt.tau2 * delta, b - opt.tau3 * delta);
      if (!probe(alpha)) return BfgsStatus::kEvalFailed;
      falpha = cache_f_;
      if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return BfgsStatus::kNoProgress;
      if (falpha > f0 + opt.rho * alpha * fp0 || falpha >= fa) {
        b = alpha; fb = falpha; fpb = nan;
      } else {
        fpalpha = cache_df_;
        if (std::fabs(fpalpha) <= -opt.sigma * fp0) {
          *alpha_new = alpha;
          return BfgsStatus::kSuccess;
        }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a; fb = fa; fpb = fpa;
        }
        a = alpha; fa = falpha; fpa = fpalpha;
      }
    }
    return BfgsStatus::kSuccess;
  }

  Eval eval_;
  bool eval_failed_ = false;
  double x0_[6] = {0}, g0_[6] = {0}, p_[6] = {0};
  double f_ = 0, delta_f_ = 0, g0norm_ = 0, pnorm_ = 0, fp0_ = 0;
  double cache_alpha_ = 0, cache_f_ = 0, cache_df_ = 0, cache_x_[6] = {0}, cache_g_[6] = {0};
};

}  // namespace gicpb
