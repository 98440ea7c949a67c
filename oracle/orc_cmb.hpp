// ORACLE (test infrastructure, NOT product code) — CMB part of the hot path:
// grids, source interpolation, line-of-sight projection, k-contraction, l-interpolation,
// correlation-function lensing, CosmoMC unit conversion.  See orc_core.hpp header.
#pragma once
#include "orc_core.hpp"

namespace orc {

// ---- time steps: camb/modules.f90:2994-3027 (SetTimeSteps) -------------------------------------
// Thermal-history scalars (taurst, taurend, dtaurec, reionisation window) are INPUTS here: they come
// from inithermo on the reference path.
inline void SetTimeSteps(Regions& TimeSteps, double taurst, double taurend, double dtaurec, double tau0,
                         bool WantTensors, bool Reionization, double reion_tau_start,
                         double reion_tau_complete, double AccuracyBoost = 1) {
  Ranges_Init(TimeSteps);
  Ranges_Add_delta(TimeSteps, taurst, taurend, dtaurec);
  double dtau0, Maxtau = tau0;
  if (WantTensors) dtau0 = std::max(taurst / 40, Maxtau / 2000. / AccuracyBoost);
  else dtau0 = Maxtau / 500. / AccuracyBoost;
  Ranges_Add_delta(TimeSteps, taurend, tau0, dtau0);
  if (Reionization) {
    int nri0 = (int)(50 * AccuracyBoost);  // reionization.f90:104-113
    Ranges_Add(TimeSteps, reion_tau_start, reion_tau_complete, nri0);
  }
  Ranges_GetArray(TimeSteps, true);
}

// dtaurec as set by InitVars + inithermo: camb/cmbmain.f90:742-745 ; modules.f90:2910-2915
inline double dtaurec_value(double qmax, double taurst, bool WantTensors, double AccuracyBoost = 1) {
  double dtaurec = 4 / qmax / AccuracyBoost;
  if (WantTensors) return std::min(dtaurec, taurst / 160) / AccuracyBoost;
  return std::min(dtaurec, taurst / 40) / AccuracyBoost;
}

// ---- source k grid: camb/cmbmain.f90:794-849 (SetkValuesForSources), flat ----------------------
struct SourceKOpts {
  bool WantScalars = true, WantTensors = false, Reionization = true, AccuratePolarization = true,
       AccurateReionization = true, HighAccuracyDefault = true, Want_CMB = true;
  int maximum_l = 2650, l_smooth_sample = 3000;
  double AccuracyBoost = 1;
};
inline void SetkValuesForSources(Regions& Evolve_q, double tau0, double taurst, double maximum_qeta,
                                 const SourceKOpts& o = SourceKOpts()) {
  double chi0 = tau0;
  double qmax = maximum_qeta / tau0, qmin = 0.1 / tau0 / o.AccuracyBoost;  // cmbmain.f90:731-733
  double dlnk0, dkn1, dkn2;
  if (o.WantScalars && o.Reionization && o.AccuratePolarization) dlnk0 = 2. / 10 / o.AccuracyBoost;
  else dlnk0 = 5. / 10 / o.AccuracyBoost;
  if (o.AccurateReionization) dlnk0 = dlnk0 / 2;
  dkn1 = 0.6 / taurst / o.AccuracyBoost;
  dkn2 = 0.9 / taurst / o.AccuracyBoost;
  if (o.HighAccuracyDefault) dkn2 = dkn2 / (double)1.2f;
  if (o.WantTensors) { dkn1 = dkn1 * 0.8; dlnk0 = dlnk0 / 2; dkn2 = dkn2 * 0.85; }
  double qmax_log = dkn1 / dlnk0;
  double q_switch = (double)(2 * 6.3f) / taurst;
  double q_cmb = 2 * o.l_smooth_sample / chi0 * o.AccuracyBoost;
  if (o.Want_CMB && o.maximum_l > 5000 && o.AccuratePolarization) q_cmb = q_cmb * (double)1.4f;
  double dksmooth = q_cmb / 2 / (o.AccuracyBoost * o.AccuracyBoost);
  if (o.Want_CMB) dksmooth = dksmooth / 6;
  Ranges_Init(Evolve_q);
  Ranges_Add_delta(Evolve_q, qmin, qmax_log, dlnk0, true);
  Ranges_Add_delta(Evolve_q, qmax_log, std::min(qmax, q_switch), dkn1);
  if (qmax > q_switch) {
    Ranges_Add_delta(Evolve_q, q_switch, std::min(q_cmb, qmax), dkn2);
    if (qmax > q_cmb) {
      dksmooth = std::log(1 + dksmooth / q_cmb);
      Ranges_Add_delta(Evolve_q, q_cmb, qmax, dksmooth, true);
    }
  }
  Ranges_GetArray(Evolve_q, false);
}

// ---- integration q grid: camb/cmbmain.f90:1221-1293 (SetkValuesForInt), flat --------------------
inline void SetkValuesForInt(Regions& q, double tau0, double maximum_qeta, int maximum_l,
                             bool HighAccuracyDefault = true, double AccuracyBoost = 1) {
  double chi0 = tau0, r = 1;
  double qmax = maximum_qeta / tau0, qmin = 0.1 / tau0 / AccuracyBoost;
  double max_bessels_etak = maximum_qeta;  // cmbmain.f90:245
  double qmax_int = std::min(qmax, max_bessels_etak / tau0);
  double IntSampleBoost = AccuracyBoost;
  Ranges_Init(q);
  int lognum = nint_(10 * IntSampleBoost);
  double dlnk1 = 1. / lognum;
  int no = nint_(600 * IntSampleBoost);
  double dk0 = 1.8 / r / chi0 / IntSampleBoost;
  double dk = 3. / r / chi0 / IntSampleBoost;
  if (HighAccuracyDefault) dk = dk / (double)1.6f;
  double k_max_log = lognum * dk0;
  double k_max_0 = no * dk0;
  double dk2 = (double)0.04f / IntSampleBoost;
  Ranges_Add_delta(q, qmin, k_max_log, dlnk1, true);
  Ranges_Add_delta(q, k_max_log, std::min(qmax_int, k_max_0), dk0);
  if (qmax_int > k_max_0) {
    double max_k_dk = std::max(3000, 2 * maximum_l) / tau0;
    Ranges_Add_delta(q, k_max_0, std::min(qmax_int, max_k_dk), dk);
    if (qmax_int > max_k_dk) Ranges_Add_delta(q, max_k_dk, qmax_int, dk2);
  }
  Ranges_GetArray(q, true);  // Init_ClTransfer, modules.f90:1110
}

// ---- one parameter point's projection inputs ---------------------------------------------------
struct ProjInput {
  double tau0 = 0;
  int n_tau = 0, n_k = 0, n_src = 3;
  const Regions* TimeSteps = nullptr;  // points + dpoints (1..n_tau)
  const double* k_src = nullptr;       // Evolve_q%points [n_k]
  const double* Src = nullptr;         // Fortran Src(k, s, tau): index k + n_k*(s + n_src*tau)
  bool WantTensors = false;
  double maximum_qeta = 14000;
  double AccuracyBoost = 1;
  bool HighAccuracyDefault = true;
};

// camb/cmbmain.f90:1207-1218 InitSourceInterpolation
inline void InitSourceInterpolation(const ProjInput& in, std::vector<double>& ddSrc) {
  ddSrc.assign((size_t)in.n_k * in.n_src * in.n_tau, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < in.n_tau; i++)
    for (int j = 0; j < in.n_src; j++) {
      size_t off = (size_t)in.n_k * (j + (size_t)in.n_src * i);
      spline(in.k_src, in.Src + off, in.n_k, spl_large, spl_large, &ddSrc[off]);
    }
}

struct ProjCounters { long long triples = 0; };  // instrumented unit-of-work counts

// camb/cmbmain.f90:478-498 SourceToTransfers -> 1295-1374 InterpolateSources -> 1387-1420
// DoSourceIntegration -> 1440-1562 DoFlatIntegration (flat, scalar with lensing source or tensor).
// Delta layout follows Fortran Delta_p_l_k(s, j, q): index s + n_src*(j + nl*q_ix).
inline void SourceToTransfers(const ProjInput& in, const std::vector<double>& ddSrc, const BesselTable& B,
                              const std::vector<int>& ls, double q_val, int q_ix, double* Delta,
                              ProjCounters* cnt = nullptr) {
  const int nt = in.n_tau, ns = in.n_src, nk = in.n_k, nl = (int)ls.size();
  const Regions& TS = *in.TimeSteps;
  const double* tp = TS.points.data();  // tp[i-1] = TimeSteps%points(i)
  std::vector<double> Source_q((size_t)(nt + 1) * ns, 0.0);  // Source_q(i,s) -> [(i)*ns + s], i 1-based
  auto SQ = [&](int i, int s) -> double& { return Source_q[(size_t)i * ns + s]; };
  // IntegrationVars_Init (cmbmain.f90:1377-1384): entries 1, nt-1, nt are zero (allocate does not
  // zero the rest but InterpolateSources overwrites 2..nt).
  // InterpolateSources
  int klo = 1;
  while ((q_val > in.k_src[klo]) && (klo < (nk - 1))) klo++;  // points(klo+1) -> k_src[klo]
  int khi = klo + 1;
  double ho = in.k_src[khi - 1] - in.k_src[klo - 1];
  double a0 = (in.k_src[khi - 1] - q_val) / ho;
  double b0 = (q_val - in.k_src[klo - 1]) / ho;
  double ho2o6 = ho * ho / 6;
  double a03 = (a0 * a0 * a0 - a0);
  double b03 = (b0 * b0 * b0 - b0);
  double max_etak_tensor = in.AccuracyBoost * in.maximum_qeta / 10;  // cmbmain.f90:747
  int step = 2;
  for (int i = 2; i <= nt; i++) {
    double xf = q_val * (in.tau0 - tp[i - 1]);
    bool ok;
    if (in.WantTensors) ok = (q_val * tp[i - 1] < max_etak_tensor) && xf > 1.e-8;
    else ok = xf > 1.e-8;  // WantLateTime = .true. (DoLensing), cmbmain.f90:134,1350
    if (ok) {
      step = i;
      for (int s = 0; s < ns; s++) {
        size_t base = (size_t)nk * (s + (size_t)ns * (i - 1));
        SQ(i, s) = a0 * in.Src[base + klo - 1] + b0 * in.Src[base + khi - 1] +
                   (a03 * ddSrc[base + klo - 1] + b03 * ddSrc[base + khi - 1]) * ho2o6;
      }
    } else
      for (int s = 0; s < ns; s++) SQ(i, s) = 0;
  }
  int SourceSteps = step;
  // NOTE: InterpolateSources overwrites entries nt-1 and nt that IntegrationVars_Init zeroed
  // (cmbmain.f90:485 is called BEFORE :491), so only entry 1 stays zero.
  for (int s = 0; s < ns; s++) SQ(1, s) = 0;

  // DoSourceIntegration (flat): cmbmain.f90:1392,1402-1407
  double nu = q_val;  // CP%r = 1
  int llmax = nint_(nu * in.tau0);
  if (llmax < 15) llmax = 17;
  else llmax = nint_(nu * (in.tau0 + 6 * pi / nu));

  // DoFlatIntegration
  std::vector<double> aa(SourceSteps + 1), fac(SourceSteps + 1);
  std::vector<int> bes_index(SourceSteps + 1);
  const double* bx = B.BessRanges.points.data();
  for (int j = 1; j <= SourceSteps; j++) {
    double xf = std::fabs(q_val * (in.tau0 - tp[j - 1]));
    int bes_ix = Ranges_IndexOf(B.BessRanges, xf);
    bes_index[j] = bes_ix;
    fac[j] = bx[bes_ix] - bx[bes_ix - 1];
    aa[j] = (bx[bes_ix] - xf) / fac[j];
    fac[j] = fac[j] * fac[j] * aa[j] / 6;
  }
  const int nxx = B.num_xx;
  for (int j = 1; j <= nl; j++) {
    int l = ls[j - 1];
    if (l > llmax) return;
    double xlim = xlimfrac * l;
    xlim = std::max(xlim, xlimmin);
    xlim = l - xlim;
    double xlmax1 = 80 * l * in.AccuracyBoost;
    double tmin = in.tau0 - xlmax1 / q_val;
    tmin = std::max(tp[1], tmin);
    double tmax = in.tau0 - xlim / q_val;
    tmax = std::min(in.tau0, tmax);
    tmin = std::max(tp[1], tmin);
    if (tmax < tp[1]) break;
    double sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double* ajl = &B.ajl[(size_t)(j - 1) * nxx];
    const double* ajlpr = &B.ajlpr[(size_t)(j - 1) * nxx];
    bool DoInt = true;
    if (!in.WantTensors) {
      double qmax_int = std::max(850, l) * 3 * in.AccuracyBoost / in.tau0;
      if (in.HighAccuracyDefault) qmax_int = qmax_int * (double)1.2f;
      DoInt = q_val < qmax_int;
    }
    if (DoInt) {
      int n1 = Ranges_IndexOf(TS, tmin), n2 = std::min(SourceSteps, Ranges_IndexOf(TS, tmax));
      for (int n = n1; n <= n2; n++) {
        double a2 = aa[n];
        int bes_ix = bes_index[n];
        double J_l = a2 * ajl[bes_ix - 1] +
                     (1 - a2) * (ajl[bes_ix] - ((a2 + 1) * ajlpr[bes_ix - 1] + (2 - a2) * ajlpr[bes_ix]) * fac[n]);
        J_l = J_l * TS.dpoints[n - 1];
        sums[0] = sums[0] + SQ(n, 0) * J_l;
        sums[1] = sums[1] + SQ(n, 1) * J_l;
        sums[2] = sums[2] + SQ(n, 2) * J_l;
      }
      if (cnt && n2 >= n1) cnt->triples += (n2 - n1 + 1);
    }
    bool UseLimber = l > 400 * std::sqrt(in.AccuracyBoost);  // cmbmain.f90:1434
    if ((!DoInt || UseLimber) && !in.WantTensors) {
      double xf = in.tau0 - (l + 0.5) / q_val;
      if (xf < TS.Highest && xf > TS.Lowest) {
        int n = Ranges_IndexOf(TS, xf);
        xf = (xf - tp[n - 1]) / (tp[n] - tp[n - 1]);
        sums[2] = (SQ(n, 2) * (1 - xf) + xf * SQ(n + 1, 2)) * std::sqrt(pi / 2 / (l + 0.5)) / q_val;
      } else sums[2] = 0;
    }
    for (int s = 0; s < ns; s++) Delta[s + (size_t)ns * ((j - 1) + (size_t)nl * q_ix)] += sums[s];
  }
}

// ---- initial power: camb/power_tilt.f90:114-169 ; source/Calculator_CAMB.f90:839-877 -------------
struct InitPower {
  double As = 2.1e-9, ns = 0.96, nrun = 0, nrunrun = 0, r = 0, nt = 0, ntrun = 0;
  double k_0_scalar = 0.05, k_0_tensor = 0.05;
  int tensor_parameterization = 1;  // 1 = indeptilt, 2 = rpivot, 3 = AT (power_tilt.f90:36-37)
};
inline double ScalarPower(const InitPower& P, double k) {
  double lnrat = std::log(k / P.k_0_scalar);
  return P.As * std::exp(lnrat * (P.ns - 1 + lnrat * (P.nrun / 2 + P.nrunrun / 6 * lnrat)));
}
inline double TensorPower(const InitPower& P, double k) {
  double lnrat = std::log(k / P.k_0_tensor);
  double k_dep = std::exp(lnrat * (P.nt + P.ntrun / 2 * lnrat));
  if (P.tensor_parameterization == 1) return P.r * P.As * k_dep;
  if (P.tensor_parameterization == 2) return P.r * ScalarPower(P, P.k_0_tensor) * k_dep;
  return 0;
}
// CosmoMC -> CAMB mapping incl. inflation consistency (Calculator_CAMB.f90:839-877).
// cmc[] = {logA-derived As (already cl_norm*As i.e. absolute), ns, nrun, nrunrun, r, nt, ntrun}
inline InitPower SetCAMBInitPower(double As, double ns, double nrun, double nrunrun, double r, double nt,
                                  double ntrun, bool inflation_consistency, double pivot_k,
                                  double tensor_pivot_k) {
  InitPower P;
  P.k_0_scalar = pivot_k; P.k_0_tensor = tensor_pivot_k;
  P.tensor_parameterization = (tensor_pivot_k != pivot_k) ? 2 : 1;
  P.As = As; P.r = r; P.ns = ns; P.nrun = nrun; P.nrunrun = nrunrun;
  if (inflation_consistency) {
    P.nt = -r / 8 * (2 - ns - r / 8);
    P.ntrun = r / 8 * (r / 8 + ns - 1);
  } else { P.nt = nt; P.ntrun = ntrun; }
  return P;
}

// ---- k contraction: camb/cmbmain.f90:2132-2264 CalcScalCls (flat, no 2D array, limber_phiphi=0) --
// iCl layout: iCl[j + nl*X], X = 0..5 = C_Temp,C_E,C_Cross,C_Phi,C_PhiTemp,C_PhiE
inline void CalcScalCls(const double* Delta, int ns, const std::vector<int>& ls, const Regions& q,
                        const InitPower& P, double ALens, double* iCl) {
  int nl = (int)ls.size(), nq = q.npoints;
  std::vector<double> pows(nq), dlnks(nq);
  for (int i = 0; i < nq; i++) {
    dlnks[i] = q.dpoints[i] / q.points[i];
    pows[i] = ScalarPower(P, q.points[i]);
  }
  for (int X = 0; X < 6; X++) for (int j = 0; j < nl; j++) iCl[j + nl * X] = 0;
  for (int j = 0; j < nl; j++) {
    double ell = ls[j];
    double c[6] = {0, 0, 0, 0, 0, 0};
    for (int qi = 0; qi < nq; qi++) {
      double dlnk = dlnks[qi], apowers = pows[qi];
      const double* D = Delta + (size_t)ns * (j + (size_t)nl * qi);
      c[0] = c[0] + apowers * D[0] * D[0] * dlnk;
      c[1] = c[1] + apowers * D[1] * D[1] * dlnk;
      c[2] = c[2] + apowers * D[0] * D[1] * dlnk;
      if (ns > 2) {
        c[3] = c[3] + apowers * D[2] * D[2] * dlnk;
        c[4] = c[4] + apowers * D[2] * D[0] * dlnk;
        c[5] = c[5] + apowers * D[2] * D[1] * dlnk;
      }
    }
    double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    double dbletmp = (ell * (ell + 1)) / twopi * fourpi;  // OutputDenominator = twopi
    iCl[j + nl * 0] = c[0] * dbletmp;
    iCl[j + nl * 1] = c[1] * dbletmp * ctnorm;
    iCl[j + nl * 2] = c[2] * dbletmp * std::sqrt(ctnorm);
    if (ns > 2) {
      iCl[j + nl * 3] = ALens * c[3] * fourpi * ell * ell * ell * ell;
      iCl[j + nl * 4] = std::sqrt(ALens) * c[4] * fourpi * ell * ell * ell;
      iCl[j + nl * 5] = std::sqrt(ALens) * c[5] * fourpi * ell * ell * ell * std::sqrt(ctnorm);
    }
  }
}

// camb/cmbmain.f90:2344-2397 CalcTensCls (flat). iCl[j + nl*X], X = 0..3 = CT_Temp,CT_E,CT_B,CT_Cross
inline void CalcTensCls(const double* Delta, int ns, const std::vector<int>& ls, const Regions& q,
                        const InitPower& P, double* iCl) {
  int nl = (int)ls.size(), nq = q.npoints;
  for (int j = 0; j < nl; j++) {
    double c[4] = {0, 0, 0, 0};
    for (int qi = 0; qi < nq; qi++) {
      double measure = q.dpoints[qi] / q.points[qi];
      double apowert = TensorPower(P, q.points[qi]);
      const double* D = Delta + (size_t)ns * (j + (size_t)nl * qi);
      c[0] += apowert * D[0] * D[0] * measure;
      c[1] += apowert * D[1] * D[1] * measure;
      c[2] += apowert * D[2] * D[2] * measure;
      c[3] += apowert * D[0] * D[1] * measure;
    }
    int l = ls[j];
    double ctnorm = (l * l - 1) * (double)((l + 2) * l);
    double dbletmp = (l * (l + 1)) / twopi * pi / 4;
    iCl[j + nl * 0] = c[0] * dbletmp * ctnorm;
    if (l == 1) dbletmp = 0;
    iCl[j + nl * 1] = c[1] * dbletmp;
    iCl[j + nl * 2] = c[2] * dbletmp;
    iCl[j + nl * 3] = c[3] * dbletmp * std::sqrt(ctnorm);
  }
}

// ---- l interpolation: camb/modules.f90:952-1029 ---------------------------------------------------
// all_Cl indexed by l (array of size >= l(max_ind)+1); only lmin..l(max_ind) written.
inline void InterpolateClArr(const std::vector<int>& ls, const double* iCl, double* all_Cl, int max_ind) {
  int l0 = (int)ls.size();
  std::vector<double> ddCl(l0), xl(l0);
  for (int i = 0; i < l0; i++) xl[i] = ls[i];
  spline(xl.data(), iCl, max_ind, 1.e30, 1.e30, ddCl.data());
  int llo = 1;
  for (int il = lmin; il <= ls[max_ind - 1]; il++) {
    int xi = il;
    if ((xi > ls[llo]) && (llo < max_ind)) llo++;  // lSet%l(llo+1) -> ls[llo]
    int lhi = llo + 1;
    double ho = ls[lhi - 1] - ls[llo - 1];
    double a0 = (ls[lhi - 1] - xi) / ho;
    double b0 = (xi - ls[llo - 1]) / ho;
    all_Cl[il] = a0 * iCl[llo - 1] + b0 * iCl[lhi - 1] +
                 ((a0 * a0 * a0 - a0) * ddCl[llo - 1] + (b0 * b0 * b0 - b0) * ddCl[lhi - 1]) * ho * ho / 6;
  }
}

// highL template: tmpl[X][l], X=0..3 = C_Temp,C_E,C_Cross,C_Phi (modules.f90:1162-1185)
struct HighLTemplate { std::vector<double> cl[4]; };  // each size lmax_extrap_highl+1

inline void InterpolateClArrTemplated(const std::vector<int>& ls, const double* iCl, double* all_Cl, int max_ind,
                                      int template_index /*1-based, <=0 for none*/, const HighLTemplate* T) {
  if (T && template_index >= 1 && template_index <= 3) {
    int maxdelta = max_ind;
    while (ls[maxdelta - 1] > lmax_extrap_highl) maxdelta--;
    std::vector<double> DeltaCL(ls.size(), 0.0);
    const std::vector<double>& t = T->cl[template_index - 1];
    for (int i = 0; i < maxdelta; i++) DeltaCL[i] = iCl[i] - t[ls[i]];
    InterpolateClArr(ls, DeltaCL.data(), all_Cl, maxdelta);
    for (int il = lmin; il <= ls[maxdelta - 1]; il++) all_Cl[il] = all_Cl[il] + t[il];
    if (maxdelta < max_ind) {
      std::vector<double> tmpall(ls[max_ind - 1] + 1, 0.0);
      InterpolateClArr(ls, iCl, tmpall.data(), max_ind);
      for (int il = ls[maxdelta - 3]; il <= ls[max_ind - 1]; il++) all_Cl[il] = tmpall[il];
    }
    return;
  }
  InterpolateClArr(ls, iCl, all_Cl, max_ind);
}

// ---- lensing: camb/lensing.f90:94-518 (CorrFuncFullSky + CorrFuncFullSkyImpl) ---------------------
struct LensOpts {
  double AccuracyBoost = 1;
  bool HighAccuracyDefault = true, AccurateBB = false;
  int lensed_convolution_margin = 100;
};
inline int lens_lmax_lensed(const std::vector<int>& ls, int Max_l, const LensOpts& o = LensOpts()) {
  int ix = (int)ls.size() - 1;  // lSamp%l0-1 (1-based)
  while (ls[ix - 1] > Max_l - o.lensed_convolution_margin) ix--;
  return ls[ix - 1];
}

// Cl_scalar[X][l] for l=lmin..Max_l, X = 0..3 (C_Temp,C_E,C_Cross,C_Phi); out Cl_lensed[Y][l], Y=0..3
// (CT_Temp,CT_E,CT_B,CT_Cross), l=lmin..lmax_lensed.  Returns lmax_lensed.
inline int CorrFuncFullSky(const std::vector<int>& lsamp, int Max_l, const double* const Cl_scalar[4],
                           const HighLTemplate& T, double* const Cl_lensed[4], const LensOpts& o = LensOpts()) {
  int lmax_extrap = Max_l - o.lensed_convolution_margin + 450;
  if (o.HighAccuracyDefault) lmax_extrap += 300;
  lmax_extrap = std::min(lmax_extrap_highl, lmax_extrap);
  const int lmax = std::max(lmax_extrap, Max_l);
  const int lmax_lensed = lens_lmax_lensed(lsamp, Max_l, o);

  int npoints = (int)(Max_l * 2 * o.AccuracyBoost);
  bool short_integral_range = !o.AccurateBB;
  double dtheta = pi / npoints;
  if (Max_l > 3500) dtheta = dtheta / (double)1.3f;
  int apodize_point_width = nint_((double)0.003f / dtheta);
  npoints = (int)(pi / dtheta);
  double range_fac;
  if (short_integral_range) {
    range_fac = std::max(1., 32 / o.AccuracyBoost);
    npoints = (int)(npoints / range_fac);
  } else range_fac = 1;
  int interp_fac = std::max(1, std::min(nint_(10 / o.AccuracyBoost), (int)(range_fac * 2) - 1));

  std::vector<double> ls_(lmax + 2), lfacs(lmax + 1), lfacs2(lmax + 1), lrootfacs(lmax + 1), theta_cut(lmax + 1);
  std::vector<int> lsi(lmax + 2);
  int jmax = 0;
  for (int l = lmin; l <= lmax; l++) {
    if (l <= 15 || ((l - 15) % interp_fac) == interp_fac / 2) { jmax++; lsi[jmax] = l; }
    lfacs[l] = (double)(l * (l + 1));
    lfacs2[l] = (double)((l + 2) * (l - 1));
    lrootfacs[l] = std::sqrt(lfacs[l] * lfacs2[l]);
  }
  for (int l = 2; l <= lmax; l++) theta_cut[l] = 0.244949 / std::sqrt(3. * lfacs[l] - 8.);
  std::vector<double> roots_(lmax + 6);
  double* roots = roots_.data() + 1;  // roots[-1..lmax+4]
  roots[-1] = 0;
  for (int l = 0; l <= lmax + 4; l++) roots[l] = std::sqrt((double)l);

  std::vector<double> Cphil3(lmax + 1), CTT(lmax + 1), CTE(lmax + 1), CEE(lmax + 1);
  for (int l = lmin; l <= Max_l; l++) {
    Cphil3[l] = Cl_scalar[3][l] * (2 * l + 1) * (l + 1) / ((double)l * (double)l * (double)l) / (4 * pi);
    double fac = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
    CTT[l] = Cl_scalar[0][l] * fac;
    CEE[l] = Cl_scalar[1][l] * fac;
    CTE[l] = Cl_scalar[2][l] * fac;
  }
  if (lmax > Max_l) {
    int l = Max_l;
    double sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
    double fac2 = CTT[Max_l] / (sc * T.cl[0][Max_l]);
    double fac = Cphil3[Max_l] / (sc * T.cl[3][Max_l]);
    for (l = Max_l + 1; l <= lmax; l++) {
      sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
      Cphil3[l] = T.cl[3][l] * fac * sc;
      CTT[l] = T.cl[0][l] * fac2 * sc;
      CEE[l] = T.cl[1][l] * fac2 * sc;
      CTE[l] = T.cl[2][l] * fac2 * sc;
    }
  }
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  std::vector<double> lens_contrib((size_t)4 * (lmax_lensed + 1) * nth, 0.0);

#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* lc = &lens_contrib[(size_t)4 * (lmax_lensed + 1) * tid];
    std::vector<double> P(lmax + 1), dP(lmax + 1), d_11(lmax + 1), d_m11(lmax + 1), d_22(lmax + 1), d_2m2(lmax + 1),
        d_20(lmax + 1);
    std::vector<double> cc((size_t)(jmax + 1) * 4);
#pragma omp for schedule(static)
    for (int i = 1; i <= npoints - 1; i++) {
      double theta = i * dtheta;
      double x = std::cos(theta), sinth = std::sin(theta), halfsinth = sinth / 2;
      double pmm = 1, pmmp1 = x, Cg2 = 0, sigmasq = 0;
      for (int l = 2; l <= lmax; l++) {
        P[l] = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) / l;
        dP[l] = l * (pmmp1 - x * P[l]) / (sinth * sinth);
        pmm = pmmp1;
        pmmp1 = P[l];
        double llp1 = lfacs[l];
        double fac1 = (1 - x), fac2 = (1 + x), fac = fac1 / fac2;
        d_11[l] = fac1 * dP[l] / llp1 + P[l];
        d_m11[l] = fac2 * dP[l] / llp1 - P[l];
        sigmasq = sigmasq + (1 - d_11[l]) * Cphil3[l];
        Cg2 = Cg2 + d_m11[l] * Cphil3[l];
        d_22[l] = (((4 * x - 8) / fac2 + llp1) * P[l] + 4 * fac * (fac2 + (x - 2) / llp1) * dP[l]) / lfacs2[l];
        if (theta > theta_cut[l])
          d_2m2[l] = ((llp1 - (4 * x + 8) / fac1) * P[l] + 4 / fac * (-fac1 + (x + 2) / llp1) * dP[l]) / lfacs2[l];
        else
          d_2m2[l] = lfacs[l] * lfacs2[l] * theta * theta * theta * theta *
                     (1. / 384. - (3. * lfacs[l] - 8.) / 23040. * theta * theta);
        d_20[l] = (2 * x * dP[l] - llp1 * P[l]) / lrootfacs[l];
      }
      for (int j = 1; j <= jmax; j++) {
        int l = lsi[j];
        double fac1 = (1 - x), fac2 = (1 + x), llp1 = lfacs[l];
        double rootllp1 = roots[l] * roots[l + 1];
        double rootfac1 = roots[l + 2] * roots[l - 1];
        double rootfac2 = roots[l + 3] * roots[l - 2];
        double dm11 = d_m11[l], d11 = d_11[l];
        double d2m2 = d_2m2[l], d22 = d_22[l], d20 = d_20[l];
        double d1m2 = sinth / rootfac1 * (dP[l] - 2 / fac1 * dm11);
        double d12 = sinth / rootfac1 * (dP[l] - 2 / fac2 * d11);
        double d1m3 = 0, d2m3 = 0, d3m3 = 0, d13 = 0, d23 = 0, d33 = 0;
        if (l >= 3) {
          double sinfac = 4 / sinth;
          d1m3 = (-(x + 0.5) * d1m2 * sinfac - lfacs2[l] * dm11 / rootfac1) / rootfac2;
          d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) / rootfac2;
          d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) / rootfac2;
          d13 = ((x - 0.5) * d12 * sinfac - lfacs2[l] * d11 / rootfac1) / rootfac2;
          d23 = (-fac1 * d22 * sinfac + rootfac1 * d12) / rootfac2;
          d33 = (-(x - 1.5) * d23 * sinfac - rootfac1 * d13) / rootfac2;
        }
        (void)d33;
        double d04 = 0, d2m4 = 0, d4m4 = 0, rootfac3 = 0;
        if (l >= 4) {
          rootfac3 = roots[l - 3] * roots[l + 4];
          d04 = ((-llp1 + (18 * x * x + 6) / (sinth * sinth)) * d20 - 6 * x * lfacs2[l] * dP[l] / lrootfacs[l]) /
                (rootfac2 * rootfac3);
          d2m4 = (-(6 * x + 4) * d2m3 / sinth - rootfac2 * d2m2) / rootfac3;
          d4m4 = (-7 / 5. * (llp1 - 6) * d2m2 + 12 / 5. * (-llp1 + (9 * x + 26) / fac1) * d3m3) / (llp1 - 12);
        }
        double X000 = std::exp(-llp1 * sigmasq / 4);
        double X022 = X000 * (1 + sigmasq);
        double X220 = lrootfacs[l] / 4 * X000;
        double X121 = -0.5 * rootfac1 * X000;
        double X132 = -0.5 * rootfac2 * X000;
        double X242 = 0.25 * rootfac2 * rootfac3 * X022;
        double dX000 = -llp1 / 4 * X000;
        double dX022 = (1 - llp1 / 4) * X022;
        fac1 = dX000 * dX000;
        double fac3 = X220 * X220;
        double Cg2sq = Cg2 * Cg2;
        double fac = ((X000 * X000 - 1) + Cg2sq * fac1) * P[l] + Cg2sq * fac3 * d2m2 + 8 / llp1 * fac1 * Cg2 * dm11;
        cc[(size_t)j * 4 + 0] = CTT[l] * fac;
        fac2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
        fac = 2 * Cg2 * X121 * X132 * d13 + fac2 * d22 + Cg2sq * X242 * X220 * d04;
        cc[(size_t)j * 4 + 1] = CEE[l] * fac;
        fac = (fac3 * P[l] + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) +
              fac2 * d2m2;
        cc[(size_t)j * 4 + 2] = CEE[l] * fac;
        fac = (X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
              Cg2sq * (X220 / 2 * d2m4 * X242 + (fac3 / 2 + dX022 * dX000) * d20);
        cc[(size_t)j * 4 + 3] = CTE[l] * fac;
      }
      double corr[4];
      for (int k = 0; k < 4; k++) {
        double s1 = 0, s2 = 0;
        for (int j = 1; j <= 14; j++) s1 += cc[(size_t)j * 4 + k];
        for (int j = 15; j <= jmax; j++) s2 += cc[(size_t)j * 4 + k];
        corr[k] = s1 + interp_fac * s2;
      }
      if (short_integral_range && i > npoints - apodize_point_width * 3) {
        // single-precision apodisation factor: integer**2 / real(...) in lensing.f90:436
        int d = i - npoints + apodize_point_width * 3;
        float ap = std::exp(-(float)(d * d) / (float)(2 * apodize_point_width * apodize_point_width));
        for (int k = 0; k < 4; k++) corr[k] = corr[k] * ap;
      }
      for (int l = lmin; l <= lmax_lensed; l++) {
        lc[0 + 4 * l] += corr[0] * P[l] * sinth;
        double T2 = corr[1] * d_22[l], T4 = corr[2] * d_2m2[l];
        lc[1 + 4 * l] += (T2 + T4) * halfsinth;
        lc[2 + 4 * l] += (T2 - T4) * halfsinth;
        lc[3 + 4 * l] += corr[3] * d_20[l] * sinth;
      }
    }
  }
  for (int l = lmin; l <= lmax_lensed; l++) {
    double fac = l * (l + 1) / twopi * dtheta * 2 * pi;
    double s[4] = {0, 0, 0, 0};
    for (int t = 0; t < nth; t++)
      for (int k = 0; k < 4; k++) s[k] += lens_contrib[(size_t)4 * (lmax_lensed + 1) * t + k + 4 * l];
    Cl_lensed[0][l] = s[0] * fac + Cl_scalar[0][l];
    Cl_lensed[1][l] = s[1] * fac + Cl_scalar[1][l];
    Cl_lensed[2][l] = s[2] * fac;
    Cl_lensed[3][l] = s[3] * fac + Cl_scalar[2][l];
  }
  return lmax_lensed;
}

}  // namespace orc
