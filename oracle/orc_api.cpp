// ORACLE (test infrastructure, NOT product code): C entry points over the restatement headers,
// loaded by oracle/pyoracle.py through ctypes.  Only tests/, smoke() and bench.py's CPU-baseline
// legs may load this library.
#include "orc_core.hpp"
#include "orc_cmb.hpp"
#include "orc_like.hpp"
#include "orc_bg.hpp"
#include "orc_thermo.hpp"
#include "orc_nonlin.hpp"
#include <memory>

using namespace orc;

#define ORC_TRY try {
#define ORC_CATCH                                   \
  }                                                 \
  catch (const std::exception& e) {                 \
    std::fprintf(stderr, "oracle: %s\n", e.what()); \
    return -1;                                      \
  }

extern "C" {

// Generic Ranges driver: ops[i] = {kind(0=Add_delta,1=Add), start, end, delta_or_nstep, islog}
// Outputs: npoints, points, dpoints (half ends), region descriptors reg[count][6] =
// {start_index, steps, IsLog, Low, High, delta}.  Returns count or -1.
int orc_ranges_build(int nops, const double* ops, int max_points, int* npoints, double* points, double* dpoints,
                     double* reg) {
  ORC_TRY
  Regions R;
  Ranges_Init(R);
  for (int i = 0; i < nops; i++) {
    const double* o = ops + 5 * i;
    if (o[0] == 0) Ranges_Add_delta(R, o[1], o[2], o[3], o[4] != 0);
    else Ranges_Add(R, o[1], o[2], (int)o[3], o[4] != 0);
  }
  Ranges_GetArray(R, true);
  *npoints = R.npoints;
  if (R.npoints > max_points) return -2;
  for (int i = 0; i < R.npoints; i++) { points[i] = R.points[i]; dpoints[i] = R.dpoints[i]; }
  for (int i = 1; i <= R.count; i++) {
    double* r = reg + 6 * (i - 1);
    r[0] = R.R[i].start_index; r[1] = R.R[i].steps; r[2] = R.R[i].IsLog; r[3] = R.R[i].Low; r[4] = R.R[i].High;
    r[5] = R.R[i].delta;
  }
  return R.count;
  ORC_CATCH
}

// index lookups on a Ranges built from ops (1-based Fortran indices out)
int orc_ranges_indexof(int nops, const double* ops, int n, const double* x, int* idx) {
  ORC_TRY
  Regions R;
  Ranges_Init(R);
  for (int i = 0; i < nops; i++) {
    const double* o = ops + 5 * i;
    if (o[0] == 0) Ranges_Add_delta(R, o[1], o[2], o[3], o[4] != 0);
    else Ranges_Add(R, o[1], o[2], (int)o[3], o[4] != 0);
  }
  Ranges_GetArray(R, false);
  for (int i = 0; i < n; i++) idx[i] = Ranges_IndexOf(R, x[i]);
  return 0;
  ORC_CATCH
}

int orc_initlval(int max_l, double lSampleBoost, int AccurateReionization, int* l_out, int max_n) {
  ORC_TRY
  AccuracyOpts o;
  o.lSampleBoost = lSampleBoost;
  o.AccurateReionization = AccurateReionization != 0;
  std::vector<int> ls = initlval(max_l, o);
  if ((int)ls.size() > max_n) return -2;
  for (size_t i = 0; i < ls.size(); i++) l_out[i] = ls[i];
  return (int)ls.size();
  ORC_CATCH
}

void orc_spline(const double* x, const double* y, int n, double d11, double d1n, double* d2) {
  spline(x, y, n, d11, d1n, d2);
}

void orc_bjl(int n, const int* L, const double* x, double* out) {
  for (int i = 0; i < n; i++) out[i] = bjl(L[i], x[i]);
}

// ---- Bessel table handle ------------------------------------------------------------------------
struct OrcBessel { BesselTable T; };
void* orc_bessel_create(int nl, const int* ls, double max_eta_k) {
  try {
    auto* b = new OrcBessel();
    GenerateBessels(b->T, std::vector<int>(ls, ls + nl), max_eta_k);
    return b;
  } catch (const std::exception& e) { std::fprintf(stderr, "oracle: %s\n", e.what()); return nullptr; }
}
void orc_bessel_destroy(void* h) { delete (OrcBessel*)h; }
int orc_bessel_numxx(void* h) { return ((OrcBessel*)h)->T.num_xx; }
void orc_bessel_get(void* h, double* x, double* ajl, double* ajlpr) {
  BesselTable& T = ((OrcBessel*)h)->T;
  std::memcpy(x, T.BessRanges.points.data(), sizeof(double) * T.num_xx);
  std::memcpy(ajl, T.ajl.data(), sizeof(double) * T.ajl.size());
  std::memcpy(ajlpr, T.ajlpr.data(), sizeof(double) * T.ajlpr.size());
}

// ---- per-point grids ----------------------------------------------------------------------------
int orc_time_steps(double taurst, double taurend, double tau0, double maximum_qeta, int want_tensors,
                   double reion_start, double reion_complete, int max_n, double* tau, double* dtau) {
  ORC_TRY
  Regions TS;
  double dtaurec = dtaurec_value(maximum_qeta / tau0, taurst, want_tensors != 0);
  SetTimeSteps(TS, taurst, taurend, dtaurec, tau0, want_tensors != 0, reion_start > 0, reion_start, reion_complete);
  if (TS.npoints > max_n) return -2;
  for (int i = 0; i < TS.npoints; i++) { tau[i] = TS.points[i]; dtau[i] = TS.dpoints[i]; }
  return TS.npoints;
  ORC_CATCH
}

int orc_source_k(double tau0, double taurst, double maximum_qeta, int want_tensors, int maximum_l, int max_n,
                 double* k) {
  ORC_TRY
  Regions E;
  SourceKOpts o;
  o.WantTensors = want_tensors != 0;
  o.WantScalars = !o.WantTensors;
  o.maximum_l = maximum_l;
  SetkValuesForSources(E, tau0, taurst, maximum_qeta, o);
  if (E.npoints > max_n) return -2;
  for (int i = 0; i < E.npoints; i++) k[i] = E.points[i];
  return E.npoints;
  ORC_CATCH
}

int orc_q_grid(double tau0, double maximum_qeta, int maximum_l, int max_n, double* q, double* dq) {
  ORC_TRY
  Regions Q;
  SetkValuesForInt(Q, tau0, maximum_qeta, maximum_l);
  if (Q.npoints > max_n) return -2;
  for (int i = 0; i < Q.npoints; i++) { q[i] = Q.points[i]; dq[i] = Q.dpoints[i]; }
  return Q.npoints;
  ORC_CATCH
}

// ---- projection: Delta[q][l][s] (C order) for one point ---------------------------------------------
// src layout: [tau][s][k] (C order) == Fortran Src(k,s,tau).  TimeSteps rebuilt from the thermo scalars so
// that Ranges_IndexOf is available (as in the reference).  Returns n_q (Delta must hold max_q*nl*n_src).
long long orc_project(void* bessel, int nl, const int* ls, double tau0, double taurst, double taurend,
                      double reion_start, double reion_complete, double maximum_qeta, int maximum_l,
                      int want_tensors, int n_k, const double* k_src, int n_src, const double* src, int max_q,
                      double* q_out, double* dq_out, double* Delta, long long* triples) {
  ORC_TRY
  BesselTable& B = ((OrcBessel*)bessel)->T;
  std::vector<int> lsv(ls, ls + nl);
  Regions TS, Q;
  double dtaurec = dtaurec_value(maximum_qeta / tau0, taurst, want_tensors != 0);
  SetTimeSteps(TS, taurst, taurend, dtaurec, tau0, want_tensors != 0, reion_start > 0, reion_start, reion_complete);
  SetkValuesForInt(Q, tau0, maximum_qeta, maximum_l);
  if (Q.npoints > max_q) return -2;
  ProjInput in;
  in.tau0 = tau0; in.n_tau = TS.npoints; in.n_k = n_k; in.n_src = n_src; in.TimeSteps = &TS;
  in.k_src = k_src; in.Src = src; in.WantTensors = want_tensors != 0; in.maximum_qeta = maximum_qeta;
  std::vector<double> ddSrc;
  InitSourceInterpolation(in, ddSrc);
  size_t nD = (size_t)Q.npoints * nl * n_src;
  for (size_t i = 0; i < nD; i++) Delta[i] = 0;
  long long tot = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot)
  for (int qi = 0; qi < Q.npoints; qi++) {
    ProjCounters c;
    SourceToTransfers(in, ddSrc, B, lsv, Q.points[qi], qi, Delta, &c);
    tot += c.triples;
  }
  for (int i = 0; i < Q.npoints; i++) { q_out[i] = Q.points[i]; dq_out[i] = Q.dpoints[i]; }
  if (triples) *triples = tot;
  return Q.npoints;
  ORC_CATCH
}

// initpower = {As, ns, nrun, nrunrun, r, nt, ntrun, pivot_k, tensor_pivot_k, inflation_consistency}
static InitPower ip_from(const double* p) {
  return SetCAMBInitPower(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[9] != 0, p[7], p[8]);
}

void orc_scalar_power(const double* initpower, int n, const double* k, double* out) {
  InitPower P = ip_from(initpower);
  for (int i = 0; i < n; i++) out[i] = ScalarPower(P, k[i]);
}
void orc_tensor_power(const double* initpower, int n, const double* k, double* out) {
  InitPower P = ip_from(initpower);
  for (int i = 0; i < n; i++) out[i] = TensorPower(P, k[i]);
}

// iCl[X][j] (C order: X major), X=0..5 scalar / 0..3 tensor
int orc_calc_cls(int tensors, int n_q, const double* q, const double* dq, int nl, const int* ls, int n_src,
                 const double* Delta, const double* initpower, double ALens, double* iCl) {
  ORC_TRY
  Regions Q;
  Q.npoints = n_q;
  Q.points.assign(q, q + n_q);
  Q.dpoints.assign(dq, dq + n_q);
  std::vector<int> lsv(ls, ls + nl);
  InitPower P = ip_from(initpower);
  if (tensors) CalcTensCls(Delta, n_src, lsv, Q, P, iCl);
  else CalcScalCls(Delta, n_src, lsv, Q, P, ALens, iCl);
  return 0;
  ORC_CATCH
}

// tmpl: [4][lmax_extrap_highl+1] or null.  out[l] for l = 0..l(max_ind)
int orc_interp_cl(int nl, const int* ls, const double* iCl, int max_ind, int template_index, const double* tmpl,
                  double* out) {
  ORC_TRY
  std::vector<int> lsv(ls, ls + nl);
  HighLTemplate T;
  if (tmpl)
    for (int X = 0; X < 4; X++) T.cl[X].assign(tmpl + (size_t)X * (lmax_extrap_highl + 1),
                                               tmpl + (size_t)(X + 1) * (lmax_extrap_highl + 1));
  InterpolateClArrTemplated(lsv, iCl, out, max_ind, template_index, tmpl ? &T : nullptr);
  return 0;
  ORC_CATCH
}

// cl_scalar: [4][Max_l+1] (TT,EE,TE,PP=l^4 C_phi); out: [4][lmax_lensed+1] (TT,EE,BB,TE). returns lmax_lensed
int orc_lens_cls(int nl, const int* ls, int Max_l, const double* cl_scalar, const double* tmpl, double* out,
                 int out_stride) {
  ORC_TRY
  std::vector<int> lsv(ls, ls + nl);
  HighLTemplate T;
  for (int X = 0; X < 4; X++) T.cl[X].assign(tmpl + (size_t)X * (lmax_extrap_highl + 1),
                                             tmpl + (size_t)(X + 1) * (lmax_extrap_highl + 1));
  const double* in[4];
  double* o[4];
  for (int X = 0; X < 4; X++) { in[X] = cl_scalar + (size_t)X * (Max_l + 1); o[X] = out + (size_t)X * out_stride; }
  return CorrFuncFullSky(lsv, Max_l, in, T, o);
  ORC_CATCH
}

// same with the accuracy knobs of the reference exposed (AccuracyBoost, accurate_BB: camb/lensing.f90:163-189), to
// check that the restatement CONVERGES to the reference's independent implementation (pycamb correlations.py)
int orc_lens_cls_opts(int nl, const int* ls, int Max_l, const double* cl_scalar, const double* tmpl, double* out,
                      int out_stride, double AccuracyBoost, int AccurateBB) {
  ORC_TRY
  std::vector<int> lsv(ls, ls + nl);
  HighLTemplate T;
  for (int X = 0; X < 4; X++) T.cl[X].assign(tmpl + (size_t)X * (lmax_extrap_highl + 1),
                                             tmpl + (size_t)(X + 1) * (lmax_extrap_highl + 1));
  const double* in[4];
  double* o[4];
  for (int X = 0; X < 4; X++) { in[X] = cl_scalar + (size_t)X * (Max_l + 1); o[X] = out + (size_t)X * out_stride; }
  LensOpts lo;
  lo.AccuracyBoost = AccuracyBoost;
  lo.AccurateBB = AccurateBB != 0;
  return CorrFuncFullSky(lsv, Max_l, in, T, o, lo);
  ORC_CATCH
}

}  // extern "C"

#include "orc_like_api.inc"
#include "orc_batch.inc"

extern "C" {
// ---- background (orc_bg.hpp): bg[16] layout documented there.  DA, H [nz]; extras[3] = tau0, age/Gyr, CosmomcTheta
static orc::NuTable g_nu_table;
int orc_background(const double* bg, int nz, const double* z, double* DA, double* H, double* extras) {
  ORC_TRY
  g_nu_table.init();
  Background B;
  B.set(bg, &g_nu_table);
  for (int i = 0; i < nz; i++) { DA[i] = B.AngularDiameterDistance(z[i]); H[i] = B.Hofz(z[i]); }
  if (extras) { extras[0] = B.tau0(); extras[1] = B.age_gyr(); extras[2] = B.CosmomcTheta(); }
  return 0;
  ORC_CATCH
}
int orc_nu_table(double* r1, double* dr1, double* dlnam) {
  ORC_TRY
  g_nu_table.init();
  for (int i = 0; i < NuTable::nrhopn; i++) { r1[i] = g_nu_table.r1[i + 1]; dr1[i] = g_nu_table.dr1[i + 1]; }
  *dlnam = g_nu_table.dlnam;
  return 0;
  ORC_CATCH
}
}  // extern "C"

extern "C" {
// ---- thermal history (orc_thermo.hpp).  in[8] = yhe, zre (used when optical_depth = 0), optical_depth (> 0: zre by
// bisection, Reionization_zreFromOptDepth), max_eta_k, want_tensors, transfer_kmax [h/Mpc; 0: WantTransfer = F],
// AccuracyBoost, reserved.  out[32] = tau0, taurst, taurend, reion tau_start, tau_complete, dtaurec, tau_maxvis, zre,
// z_star, z_drag, actual_opt_depth, status, derived[13] from [12], RECFAST derivative evaluations at [25].
// tables (optional) [4][20000]: xe, dotmu, emmu, cs2 of inithermo.
int orc_thermo(const double* bg, const double* in, double* out, double* tables) {
  ORC_TRY
  g_nu_table.init();
  Background B;
  B.set(bg, &g_nu_table);
  Thermo T;
  ThermoOut o = T.run(B, in[0], in[1], in[2], in[3], in[4] != 0, in[5], in[6] > 0 ? in[6] : 1.0);
  for (int i = 0; i < 32; i++) out[i] = 0;
  out[0] = o.tau0; out[1] = o.taurst; out[2] = o.taurend; out[3] = o.tau_start; out[4] = o.tau_complete; out[5] = o.dtaurec;
  out[6] = o.tau_maxvis; out[7] = o.zre; out[8] = o.z_star; out[9] = o.z_drag; out[10] = o.actual_opt_depth; out[11] = o.status;
  for (int i = 0; i < 13; i++) out[12 + i] = o.derived[i];
  out[25] = (double)T.rec.n_fcn;
  if (tables && o.status == 0) {
    for (int i = 0; i < Thermo::nthermo; i++) {
      tables[i] = T.xe[i + 1]; tables[Thermo::nthermo + i] = T.dotmu[i + 1];
      tables[2 * Thermo::nthermo + i] = T.emmu[i + 1]; tables[3 * Thermo::nthermo + i] = T.cs2[i + 1];
    }
  }
  return 0;
  ORC_CATCH
}
// xe of RECFAST alone at the given scale factors (Recombination_xe)
int orc_recfast_xe(const double* bg, double yhe, int n, const double* a, double* xe) {
  ORC_TRY
  g_nu_table.init();
  Background B;
  B.set(bg, &g_nu_table);
  Recfast R;
  R.init(B, yhe);
  for (int i = 0; i < n; i++) xe[i] = R.xe(a[i]);
  return 0;
  ORC_CATCH
}
}  // extern "C"

extern "C" {
// ---- non-linear lensing rescale and sigma_8 (orc_nonlin.hpp).  par[6] = h, omm0 (= omegac + omegab + omegan), omegav,
// fnu (= omegan / omm0), w, wa.  transfer [nz][nkt], kh [nkt], z [nz] (descending redshift = ascending time), tautf [nz].
// Outputs: sigma8 [nz], ratio [nz][nkt], spec [nz][3] = rknl, rneff, rncur; src [n_tau][3][n_k] (optional) is rescaled in place.
int orc_nonlinear(const double* initpower, const double* par, int nkt, int nz, const double* kh, const double* z,
                  const double* transfer, double* sigma8, double* ratio, double* spec, int n_k, const double* k, int n_tau,
                  const double* tau, const double* tautf, double* src) {
  ORC_TRY
  InitPower IP = ip_from(initpower);
  auto ps = [&IP](double kk) { return ScalarPower(IP, kk); };
  const double h = par[0];
  if (sigma8) sigma_R(nkt, nz, kh, transfer, h, 8.0, ps, sigma8);
  MatterPower PK;
  PK.from_transfer(nkt, nz, kh, transfer, z, h, ps);
  Halofit HF;
  HF.omm0 = par[1]; HF.omegav = par[2]; HF.fnu = par[3]; HF.w_lam = par[4]; HF.wa = par[5];
  const int err = HF.ratios(PK, spec);
  if (ratio)
    for (int i = 0; i < nz; i++)
      for (int j = 0; j < nkt; j++) ratio[(size_t)i * nkt + j] = PK.nonlin_ratio[i][j];
  if (src) make_nonlinear_sources(PK, n_k, k, h, n_tau, tau, tautf, src);
  return err;
  ORC_CATCH
}
// MatterPowerData_k at arbitrary k/h (first redshift of the table) - for the known-answer tests
int orc_matter_power_at(const double* initpower, double h, int nkt, const double* kh, const double* transfer, int n,
                        const double* kq, double* out) {
  ORC_TRY
  InitPower IP = ip_from(initpower);
  auto ps = [&IP](double kk) { return ScalarPower(IP, kk); };
  MatterPower PK;
  const double z0 = 0;
  PK.from_transfer(nkt, 1, kh, transfer, &z0, h, ps);
  for (int i = 0; i < n; i++) out[i] = PK.at(kq[i], 0);
  return 0;
  ORC_CATCH
}
}  // extern "C"
