// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the CAMB/CosmoMC hot path, used only by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs as the checker / reported CPU baseline.
// Every routine cites the reference file:line it follows (paths relative to the reference root).
// Loop order and expressions follow the Fortran so that integer artefacts (grid sizes, l-sets,
// table indices) are bit-identical and FP sums differ only by compiler reassociation.
// Build with -ffp-contract=off (see oracle/Makefile) so no FMA contraction changes int() truncations.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static const double pi = 3.14159265358979323846264338328;  // camb/constants.f90:11
static const double twopi = 2 * pi, fourpi = 4 * pi;
static const double spl_large = 1.e40;  // camb/modules.f90:227
static const int lmin = 2;              // camb/modules.f90:218
static const int lmax_extrap_highl = 8000;  // camb/modules.f90:234

// ---------------------------------------------------------------- Ranges (camb/utils.F90:8-485)
static const int Max_Ranges = 100;
static const double RangeTol = 0.1;  // utils.F90:13

struct Region {  // utils.F90:16-23
  int start_index = 0, steps = 0;
  bool IsLog = false;
  double Low = 0, High = 0, delta = 0, delta_max = 0, delta_min = 0;
};

struct Regions {  // utils.F90:25-35 ; points/dpoints are 1-based in Fortran, 0-based here
  int count = 0, npoints = 0;
  double Lowest = 0, Highest = 0;
  Region R[Max_Ranges + 1];  // 1-based use
  bool has_dpoints = false;
  std::vector<double> points, dpoints;
};

inline void Ranges_Init(Regions& R) {  // utils.F90:39-67
  R.points.clear(); R.dpoints.clear();
  R.count = 0; R.npoints = 0; R.has_dpoints = false;
}

// utils.F90:81-111 ; returns the 1-based Fortran index
inline int Ranges_IndexOf(const Regions& Reg, double tau) {
  for (int i = 1; i <= Reg.count; i++) {
    const Region& A = Reg.R[i];
    if (tau < A.High && tau >= A.Low) {
      if (A.IsLog) return A.start_index + (int)(std::log(tau / A.Low) / A.delta);
      return A.start_index + (int)((tau - A.Low) / A.delta);
    }
  }
  if (tau >= Reg.Highest) return Reg.npoints;
  throw std::runtime_error("Ranges_IndexOf: value out of range");
}

inline void Ranges_Getdpoints(Regions& Reg, bool half_ends = true) {  // utils.F90:151-176
  int n = Reg.npoints;
  Reg.dpoints.assign(n, 0.0);
  auto& p = Reg.points;
  for (int i = 2; i <= n - 1; i++) Reg.dpoints[i - 1] = (p[i] - p[i - 2]) / 2;
  if (half_ends) {
    Reg.dpoints[0] = (p[1] - p[0]) / 2;
    Reg.dpoints[n - 1] = (p[n - 1] - p[n - 2]) / 2;
  } else {
    Reg.dpoints[0] = (p[1] - p[0]);
    Reg.dpoints[n - 1] = (p[n - 1] - p[n - 2]);
  }
}

inline void Ranges_GetArray(Regions& Reg, bool want_dpoints = true) {  // utils.F90:114-148
  Reg.has_dpoints = want_dpoints;
  Reg.points.assign(Reg.npoints, 0.0);
  int ix = 0;
  for (int i = 1; i <= Reg.count; i++) {
    const Region& A = Reg.R[i];
    for (int j = 0; j <= A.steps - 1; j++) {
      ix++;
      if (A.IsLog) Reg.points[ix - 1] = A.Low * std::exp(j * A.delta);
      else Reg.points[ix - 1] = A.Low + A.delta * j;
    }
  }
  ix++;
  Reg.points[ix - 1] = Reg.Highest;
  if (ix != Reg.npoints) throw std::runtime_error("Ranges_GetArray: ERROR");
  if (Reg.has_dpoints) Ranges_Getdpoints(Reg);
}

inline void Ranges_Add(Regions& Reg, double t_start, double t_end, int nstep, bool WantLog = false) {
  // utils.F90:206-466
  static thread_local Region NewRegions[Max_Ranges + 1];
  double EndPoints[Max_Ranges * 2 + 2];
  double RequestDelta[Max_Ranges + 1];
  double delta;
  if (WantLog) delta = std::log(t_end / t_start) / nstep;
  else delta = (t_end - t_start) / nstep;
  if (t_end <= t_start) throw std::runtime_error("Ranges_Add: end must be larger than start");
  if (nstep <= 0) throw std::runtime_error("Ranges_Add: nstep must be > 0");
  if (Reg.count >= Max_Ranges) throw std::runtime_error("Ranges_Add: Increase Max_Ranges");

  for (int i = 1; i <= Reg.count; i++) NewRegions[i] = Reg.R[i];
  int nreg = Reg.count + 1;
  {
    Region& A = NewRegions[nreg];
    A = Region();
    A.Low = t_start; A.High = t_end; A.delta = delta; A.steps = nstep; A.IsLog = WantLog;
  }
  // end points in order (utils.F90:251-294)
  int ix = 0;
  for (int i = 1; i <= nreg; i++) {
    const Region& A = NewRegions[i];
    if (ix == 0) {
      ix = 1; EndPoints[ix] = A.Low;
      ix = 2; EndPoints[ix] = A.High;
    } else {
      int ixin = ix;
      for (int j = 1; j <= ixin; j++) {
        if (A.Low < EndPoints[j]) {
          for (int m = ix + 1; m >= j + 1; m--) EndPoints[m] = EndPoints[m - 1];
          EndPoints[j] = A.Low;
          ix = ix + 1;
          break;
        }
      }
      if (ixin == ix) {
        ix++; EndPoints[ix] = A.Low;
        ix++; EndPoints[ix] = A.High;
      } else {
        ixin = ix;
        for (int j = 1; j <= ixin; j++) {
          if (A.High < EndPoints[j]) {
            for (int m = ix + 1; m >= j + 1; m--) EndPoints[m] = EndPoints[m - 1];
            EndPoints[j] = A.High;
            ix = ix + 1;
            break;
          }
        }
        if (ixin == ix) { ix++; EndPoints[ix] = A.High; }
      }
    }
  }
  // remove duplicate points (utils.F90:296-304)
  {
    int ixin = ix;
    ix = 1;
    for (int i = 2; i <= ixin; i++)
      if (EndPoints[i] != EndPoints[ix]) { ix++; EndPoints[ix] = EndPoints[i]; }
  }
  Reg.Lowest = EndPoints[1];
  Reg.Highest = EndPoints[ix];
  Reg.count = 0;
  double max_delta = Reg.Highest - Reg.Lowest;

  for (int i = 1; i <= ix - 1; i++) {  // utils.F90:314-387
    Region& A = Reg.R[i];
    A.Low = EndPoints[i];
    A.High = EndPoints[i + 1];
    delta = max_delta;
    A.IsLog = false;
    for (int j = 1; j <= nreg; j++) {
      const Region& N = NewRegions[j];
      if (A.Low >= N.Low && A.Low < N.High) {
        if (N.IsLog) {
          if (A.IsLog) {
            delta = std::min(delta, N.delta);
          } else {
            double min_log_step = A.Low * (std::exp(N.delta) - 1);
            if (min_log_step < delta) {
              double max_log_step = A.High * (1 - std::exp(-N.delta));
              if (delta < max_log_step) delta = min_log_step;
              else { A.IsLog = true; delta = N.delta; }
            }
          }
        } else {
          if (A.IsLog) {
            double max_log_step = A.High * (1 - std::exp(-delta));
            if (N.delta < max_log_step) {
              double min_log_step = A.Low * (std::exp(delta) - 1);
              if (min_log_step < N.delta) { A.IsLog = false; delta = min_log_step; }
              else delta = -std::log(1 - N.delta / A.High);
            }
          } else {
            delta = std::min(delta, N.delta);
          }
        }
      }
    }
    double Diff;
    if (A.IsLog) Diff = std::log(A.High / A.Low);
    else Diff = A.High - A.Low;
    if (delta >= Diff) { A.delta = Diff; A.steps = 1; }
    else {
      A.steps = std::max(1, (int)(Diff / delta + 1.0 - RangeTol));
      A.delta = Diff / A.steps;
    }
    Reg.count++;
    RequestDelta[Reg.count] = delta;
    if (A.IsLog) {
      if (A.steps == 1) { A.delta_min = A.High - A.Low; A.delta_max = A.delta_min; }
      else {
        A.delta_min = A.Low * (std::exp(A.delta) - 1);
        A.delta_max = A.High * (1 - std::exp(-A.delta));
      }
    } else { A.delta_max = A.delta; A.delta_min = A.delta; }
  }

  // get rid of tiny regions (utils.F90:390-441)
  ix = Reg.count;
  for (int i = ix; i >= 1; i--) {
    Region& A = Reg.R[i];
    if (A.steps == 1) {
      double Diff = A.High - A.Low, min_request, max_request;
      if (A.IsLog) {
        min_request = A.Low * (std::exp(RequestDelta[i]) - 1);
        max_request = A.High * (1 - std::exp(-RequestDelta[i]));
      } else { min_request = RequestDelta[i]; max_request = min_request; }
      if (i != Reg.count) {
        Region& L = Reg.R[i + 1];
        if (RequestDelta[i] >= A.delta && Diff <= L.delta_min && L.delta_min <= max_request) {
          L.Low = A.Low;
          if (Diff > L.delta_min * RangeTol) L.steps = L.steps + 1;
          if (L.IsLog) L.delta = std::log(L.High / L.Low) / L.steps;
          else L.delta = (L.High - L.Low) / L.steps;
          for (int m = i; m <= Reg.count - 1; m++) Reg.R[m] = Reg.R[m + 1];
          Reg.count--;
          continue;
        }
      }
      if (i != 1) {
        Region& L = Reg.R[i - 1];
        if (RequestDelta[i] >= A.delta && Diff <= L.delta_max && L.delta_max <= min_request) {
          L.High = A.High;
          if (Diff > L.delta_max * RangeTol) L.steps = L.steps + 1;
          if (L.IsLog) L.delta = std::log(L.High / L.Low) / L.steps;
          else L.delta = (L.High - L.Low) / L.steps;
          for (int m = i; m <= Reg.count - 1; m++) Reg.R[m] = Reg.R[m + 1];
          Reg.count--;
        }
      }
    }
  }
  // start indices (utils.F90:444-464)
  int nsteps = 1;
  for (int i = 1; i <= Reg.count; i++) {
    Region& A = Reg.R[i];
    A.start_index = nsteps;
    nsteps += A.steps;
    if (A.IsLog) {
      if (A.steps == 1) { A.delta_min = A.High - A.Low; A.delta_max = A.delta_min; }
      else {
        A.delta_min = A.Low * (std::exp(A.delta) - 1);
        A.delta_max = A.High * (1 - std::exp(-A.delta));
      }
    } else { A.delta_max = A.delta; A.delta_min = A.delta; }
  }
  Reg.npoints = nsteps;
}

inline void Ranges_Add_delta(Regions& Reg, double t_start, double t_end, double t_approx_delta,
                             bool WantLog = false) {  // utils.F90:179-203
  if (t_end <= t_start) throw std::runtime_error("Ranges_Add_delta: end must be larger than start");
  if (t_approx_delta <= 0) throw std::runtime_error("Ranges_Add_delta: delta must be > 0");
  int n;
  if (WantLog) n = std::max(1, (int)(std::log(t_end / t_start) / t_approx_delta + 1.0 - RangeTol));
  else n = std::max(1, (int)((t_end - t_start) / t_approx_delta + 1.0 - RangeTol));
  Ranges_Add(Reg, t_start, t_end, n, WantLog);
}

// ---------------------------------------------------------------- spline (camb/subroutines.f90:253-296)
inline void spline(const double* x, const double* y, int n, double d11, double d1n, double* d2) {
  std::vector<double> u(n);
  double d1r = (y[1] - y[0]) / (x[1] - x[0]), d1l;
  if (d11 > .99e30) { d2[0] = 0; u[0] = 0; }
  else { d2[0] = -0.5; u[0] = (3. / (x[1] - x[0])) * (d1r - d11); }
  for (int i = 1; i <= n - 2; i++) {
    d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    double xxdiv = 1. / (x[i + 1] - x[i - 1]);
    double sig = (x[i] - x[i - 1]) * xxdiv;
    double xp = 1. / (sig * d2[i - 1] + 2.);
    d2[i] = (sig - 1.) * xp;
    u[i] = (6. * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
  }
  d1l = d1r;
  double qn, un;
  if (d1n > .99e30) { qn = 0; un = 0; }
  else { qn = 0.5; un = (3. / (x[n - 1] - x[n - 2])) * (d1n - d1l); }
  d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.);
  for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// ---------------------------------------------------------------- rombint (camb/subroutines.f90:117-176)
template <class F>
inline double rombint(F f, double a, double b, double tol) {
  const int MAXITER = 20, MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(a) + f(b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20, g0 = 0, g1, fourj;
  int i = 0;
  for (;;) {
    i++;
    if (i > MAXITER || (i > 5 && std::fabs(error) < tol)) break;
    g0 = 0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(a + (k + k - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    int jmax = std::min(i, MAXJ);
    fourj = 1;
    for (int j = 1; j <= jmax; j++) {
      fourj = 4 * fourj;
      g1 = g0 + (g0 - g[j]) / (fourj - 1);
      g[j] = g0;
      g0 = g1;
    }
    if (std::fabs(g0) > tol) error = 1 - gmax / g0;
    else error = gmax;
    gmax = g0;
    g[jmax + 1] = g0;
  }
  return g0;
}

// ---------------------------------------------------------------- l sampling (camb/modules.f90:791-950)
struct AccuracyOpts {
  double AccuracyBoost = 1, lSampleBoost = 1, scale = 1;
  bool HighAccuracyDefault = true, AccurateReionization = true, use_spline_template = true, flat = true;
};

inline int nint_(double x) { return (int)std::lround(x); }  // Fortran nint: round half away from zero

inline std::vector<int> initlval(int max_l, const AccuracyOpts& o = AccuracyOpts()) {
  std::vector<int> ls(4001 + 16, 0);  // 1-based, lmax_arr=4000
  double Ascale = o.scale / o.lSampleBoost;
  int lind = 0, lvar, step, top, bot;
  auto finish = [&](int n) { return std::vector<int>(ls.begin() + 1, ls.begin() + 1 + n); };
  if (o.lSampleBoost >= 50) {
    for (lvar = lmin; lvar <= max_l; lvar++) ls[++lind] = lvar;
    return finish(lind);
  }
  for (lvar = lmin; lvar <= 10; lvar++) ls[++lind] = lvar;
  if (o.AccurateReionization) {
    if (o.lSampleBoost > 1) { for (lvar = 11; lvar <= 37; lvar += 1) ls[++lind] = lvar; }
    else { for (lvar = 11; lvar <= 37; lvar += 2) ls[++lind] = lvar; }
    step = std::max(nint_(5 * Ascale), 2);
    bot = 40;
    top = bot + step * 10;
  } else {
    if (o.lSampleBoost > 1) { for (lvar = 11; lvar <= 15; lvar++) ls[++lind] = lvar; }
    else { ls[++lind] = 12; ls[++lind] = 15; }
    step = std::max(nint_(10 * Ascale), 3);
    bot = 15 + std::max(step / 2, 2);
    top = bot + step * 7;
  }
  for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
  // (Log_lvalues = .false. branch, modules.f90:871-945)
  step = std::max(nint_(20 * Ascale), 4);
  bot = ls[lind] + step;
  top = bot + step * 2;
  for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
  if (ls[lind] >= max_l) {
    for (lvar = lind; lvar >= 1; lvar--) if (ls[lvar] <= max_l) break;
    lind = lvar;
    if (ls[lind] < max_l) { lind++; ls[lind] = max_l; }
  } else {
    step = std::max(nint_(25 * Ascale), 4);
    bot = ls[lind] + step;
    top = bot + step;
    for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
    if (ls[lind] >= max_l) {
      for (lvar = lind; lvar >= 1; lvar--) if (ls[lvar] <= max_l) break;
      lind = lvar;
      if (ls[lind] < max_l) { lind++; ls[lind] = max_l; }
    } else {
      if (o.HighAccuracyDefault && !o.use_spline_template) step = std::max(nint_(42 * Ascale), 7);
      else step = std::max(nint_(50 * Ascale), 7);
      bot = ls[lind] + step;
      top = std::min(5000, max_l);
      for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
      if (max_l > 5000) {
        step = std::max(nint_(400 * Ascale), 50);
        lvar = ls[lind];
        for (;;) {
          lvar = lvar + step;
          if (lvar > max_l) break;
          ls[++lind] = lvar;
          step = nint_(step * 1.5);
        }
      }
      if (ls[lind] != max_l) { lind++; ls[lind] = max_l; }
      if (!o.flat) ls[lind - 1] = (int)(max_l + ls[lind - 2]) / 2;
    }
  }
  return finish(lind);
}

// ---------------------------------------------------------------- bjl (camb/bessels.f90:132-275)
inline double bjl(int L, double X) {
  const double LN2 = 0.6931471805599453094, ONEMLN2 = 0.30685281944005469058277;
  const double PID2 = 1.5707963267948966192313217, PID4 = 0.78539816339744830961566084582;
  const double ROOTPI12 = 21.269446210866192327578;
  const double GAMMA1 = 2.6789385347077476336556, GAMMA2 = 1.3541179394264004169452;
  double JL;
  double AX = std::fabs(X), AX2 = AX * AX;
  if (L < 7) {
    if (L == 0) {
      if (AX < 1e-1) JL = 1 - AX2 / 6 * (1 - AX2 / 20);
      else JL = std::sin(AX) / AX;
    } else if (L == 1) {
      if (AX < 2e-1) JL = AX / 3 * (1 - AX2 / 10 * (1 - AX2 / 28));
      else JL = (std::sin(AX) / AX - std::cos(AX)) / AX;
    } else if (L == 2) {
      if (AX < 3e-1) JL = AX2 / 15 * (1 - AX2 / 14 * (1 - AX2 / 36));
      else JL = (-3.0 * std::cos(AX) / AX - std::sin(AX) * (1 - 3 / AX2)) / AX;
    } else if (L == 3) {
      if (AX < 4e-1) JL = AX * AX2 / 105 * (1 - AX2 / 18 * (1 - AX2 / 44));
      else JL = (std::cos(AX) * (1 - 15 / AX2) - std::sin(AX) * (6 - 15 / AX2) / AX) / AX;
    } else if (L == 4) {
      if (AX < 6e-1) JL = AX2 * AX2 / 945 * (1 - AX2 / 22 * (1 - AX2 / 52));
      else JL = (std::sin(AX) * (1 - (45 - 105 / AX2) / AX2) + std::cos(AX) * (10 - 105 / AX2) / AX) / AX;
    } else if (L == 5) {
      if (AX < 1) JL = AX2 * AX2 * AX / 10395 * (1 - AX2 / 26 * (1 - AX2 / 60));
      else JL = (std::sin(AX) * (15 - (420 - 945 / AX2) / AX2) / AX -
                 std::cos(AX) * (1 - (105 - 945.0 / AX2) / AX2)) / AX;
    } else {
      if (AX < 1) JL = AX2 * AX2 * AX2 / 135135 * (1 - AX2 / 30 * (1 - AX2 / 68));
      else JL = (std::sin(AX) * (-1 + (210 - (4725 - 10395 / AX2) / AX2) / AX2) +
                 std::cos(AX) * (-21 + (1260 - 10395 / AX2) / AX2) / AX) / AX;
    }
  } else {
    double NU = 0.5 + L, NU2 = NU * NU;
    if (AX < 1e-40) {
      JL = 0;
    } else if ((AX2 / L) < 5e-1) {
      JL = std::exp(L * std::log(AX / NU) - LN2 + NU * ONEMLN2 - (1 - (1 - 3.5 / NU2) / NU2 / 30) / 12 / NU) /
           NU * (1 - AX2 / (4 * NU + 4) * (1 - AX2 / (8 * NU + 16) * (1 - AX2 / (12 * NU + 36))));
    } else if (((double)L * (double)L / AX) < 5e-1) {
      double BETA = AX - PID2 * (L + 1);
      JL = (std::cos(BETA) * (1 - (NU2 - 0.25) * (NU2 - 2.25) / 8 / AX2 * (1 - (NU2 - 6.25) * (NU2 - 12.25) / 48 / AX2)) -
            std::sin(BETA) * (NU2 - 0.25) / 2 / AX *
                (1 - (NU2 - 2.25) * (NU2 - 6.25) / 24 / AX2 * (1 - (NU2 - 12.25) * (NU2 - 20.25) / 80 / AX2))) / AX;
    } else {
      // Fortran: NU**0.325 with single-precision literal 0.325 promoted to double; 1.31/1.48 likewise
      double L3 = std::pow(NU, (double)0.325f);
      if (AX < NU - (double)1.31f * L3) {
        double COSB = NU / AX, SX = std::sqrt(NU2 - AX2), COTB = NU / SX, SECB = AX / NU;
        double BETA = std::log(COSB + SX / AX);
        double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        double EXPTERM = ((2 + 3 * SEC2B) * COT3B / 24 -
                          ((4 + SEC2B) * SEC2B * COT6B / 16 +
                           ((16 - (1512 + (3654 + 375 * SEC2B) * SEC2B) * SEC2B) * COT3B / 5760 +
                            (32 + (288 + (232 + 13 * SEC2B) * SEC2B) * SEC2B) * SEC2B * COT6B / 128 / NU) * COT6B / NU) / NU) / NU;
        JL = std::sqrt(COTB * COSB) / (2 * NU) * std::exp(-NU * BETA + NU / COTB - EXPTERM);
      } else if (AX > NU + (double)1.48f * L3) {
        double COSB = NU / AX, SX = std::sqrt(AX2 - NU2), COTB = NU / SX, SECB = AX / NU;
        double BETA = std::acos(COSB);
        double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        double TRIGARG = NU / COTB - NU * BETA - PID4 -
                         ((2.0 + 3.0 * SEC2B) * COT3B / 24 +
                          (16 - (1512 + (3654 + 375 * SEC2B) * SEC2B) * SEC2B) * COT3B * COT6B / 5760 / NU2) / NU;
        double EXPTERM = ((4 + SEC2B) * SEC2B * COT6B / 16 -
                          (32 + (288 + (232 + 13 * SEC2B) * SEC2B) * SEC2B) * SEC2B * COT6B * COT6B / 128 / NU2) / NU2;
        JL = std::sqrt(COTB * COSB) / NU * std::exp(-EXPTERM) * std::cos(TRIGARG);
      } else {
        double BETA = AX - NU, BETA2 = BETA * BETA, SX = 6 / AX, SX2 = SX * SX;
        double SECB = std::pow(SX, 0.3333333333333333), SEC2B = SECB * SECB;
        JL = (GAMMA1 * SECB + BETA * GAMMA2 * SEC2B - (BETA2 / 18 - 1.0 / 45) * BETA * SX * SECB * GAMMA1 -
              ((BETA2 - 1) * BETA2 / 36 + 1.0 / 420) * SX * SEC2B * GAMMA2 +
              (((BETA2 / 1620 - 7.0 / 3240) * BETA2 + 1.0 / 648) * BETA2 - 1.0 / 8100) * SX2 * SECB * GAMMA1 +
              (((BETA2 / 4536 - 1.0 / 810) * BETA2 + 19.0 / 11340) * BETA2 - 13.0 / 28350) * BETA * SX2 * SEC2B * GAMMA2 -
              ((((BETA2 / 349920 - 1.0 / 29160) * BETA2 + 71.0 / 583200) * BETA2 - 121.0 / 874800) * BETA2 +
               7939.0 / 224532000) * BETA * SX2 * SX * SECB * GAMMA1) * std::sqrt(SX) / ROOTPI12;
      }
    }
  }
  if (X < 0 && (L % 2) != 0) JL = -JL;
  return JL;
}

// ---------------------------------------------------------------- flat Bessel table (camb/bessels.f90:50-120)
static const double xlimmin = 35., xlimfrac = 0.05;  // bessels.f90:24

struct BesselTable {
  Regions BessRanges;
  int num_xx = 0, nl = 0;
  std::vector<int> l;
  std::vector<double> ajl, ajlpr;  // column-major [j*num_xx + i]
};

inline void GenerateBessels(BesselTable& T, const std::vector<int>& ls, double max_eta_k, double AccuracyBoost = 1) {
  int kmaxfile = (int)(max_eta_k) + 1;  // bessels.f90:64
  Ranges_Init(T.BessRanges);
  Ranges_Add_delta(T.BessRanges, 0., 1., 0.01);
  Ranges_Add_delta(T.BessRanges, 1., 5., 0.1);
  Ranges_Add_delta(T.BessRanges, 5., 25., 0.2);
  Ranges_Add_delta(T.BessRanges, 25., 150., 0.5 / AccuracyBoost);
  Ranges_Add_delta(T.BessRanges, 150., (double)kmaxfile, 0.8 / AccuracyBoost);
  Ranges_GetArray(T.BessRanges, false);
  int num_xx = T.BessRanges.npoints;
  T.num_xx = num_xx; T.nl = (int)ls.size(); T.l = ls;
  T.ajl.assign((size_t)num_xx * T.nl, 0.0);
  T.ajlpr.assign((size_t)num_xx * T.nl, 0.0);
  const double* xs = T.BessRanges.points.data();
#pragma omp parallel for schedule(static)
  for (int j = 0; j < T.nl; j++) {
    double* a = &T.ajl[(size_t)j * num_xx];
    int lj = ls[j];
    for (int i = 0; i < num_xx; i++) {
      double x = xs[i];
      double xlim = xlimfrac * lj;
      xlim = std::max(xlim, xlimmin);
      xlim = lj - xlim;
      if (x > xlim) {
        if ((lj == 3 && x <= 0.2) || (lj > 3 && x < 0.5) || (lj > 5 && x < 1.0)) a[i] = 0;
        else a[i] = bjl(lj, x);
      } else a[i] = 0;
    }
    spline(xs, a, num_xx, spl_large, spl_large, &T.ajlpr[(size_t)j * num_xx]);
  }
}

}  // namespace orc
