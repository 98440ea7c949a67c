// ORACLE (test infrastructure, NOT product code): CPU restatement of CAMB's background functions and of the
// background-only likelihoods' theory vectors.  Only tests/, smoke() and bench.py's CPU-baseline legs use it.
//
//   CAMBParams_Set (densities)                 camb/modules.f90:258-446
//   MassiveNu: Nu_init / nuRhoPres / Nu_rho    camb/modules.f90:1532-1718 ; splini/splder/splint subroutines.f90:6-50,342-364
//   dtauda                                     camb/equations.f90:69-100
//   DeltaTime / ComovingRadialDistance / AngularDiameterDistance / Hofz / BAO_D_v / dsound_da / CosmomcTheta
//                                              camb/modules.f90:519-751
//   DeltaPhysicalTimeGyr                       camb/modules.f90:558-572
#pragma once
#include "orc_core.hpp"

namespace orc {

namespace cst {  // camb/constants.f90:15-60
constexpr double pi = 3.1415926535897932384626433832795;
constexpr double c = 2.99792458e8;
constexpr double G = 6.6738e-11;
constexpr double sigma_boltz = 5.6704e-8;
constexpr double Gyr = 3.1556926e16;
constexpr double Mpc = 3.085678e22;
constexpr double kappa = 8. * pi * G;
}  // namespace cst

// ---- splini / splder / splint (camb/subroutines.f90:6-50, 342-364) ----
inline void splini(std::vector<double>& g, int n) {
  g.assign(n + 1, 0.0);  // 1-based
  g[1] = 0;
  for (int i = 2; i <= n; i++) g[i] = 1 / (4. - g[i - 1]);
}
inline void splder(const std::vector<double>& y, std::vector<double>& dy, int n, const std::vector<double>& g) {
  std::vector<double> f(n + 1);
  dy.assign(n + 1, 0.0);
  const int n1 = n - 1;
  f[1] = (-10. * y[1] + 15. * y[2] - 6. * y[3] + y[4]) / 6.;
  f[n] = (10. * y[n] - 15. * y[n1] + 6. * y[n - 2] - y[n - 3]) / 6.;
  for (int i = 2; i <= n1; i++) f[i] = g[i] * (3. * (y[i + 1] - y[i - 1]) - f[i - 1]);
  dy[n] = f[n];
  for (int i = n1; i >= 1; i--) dy[i] = f[i] - g[i] * dy[i + 1];
}
inline double splint(const double* y /*1-based*/, int n) {
  const int n1 = n - 1;
  const double dy1 = 0.;
  const double dyn = (11. * y[n] - 18. * y[n1] + 9. * y[n - 2] - 2. * y[n - 3]) / 6.;
  double z = 0.5 * (y[1] + y[n]) + (dy1 - dyn) / 12.;
  double s = 0;
  for (int i = 2; i <= n1; i++) s += y[i];
  return z + s;
}

// ---- MassiveNu tables (camb/modules.f90:1484-1610) ----
struct NuTable {
  static constexpr int nrhopn = 2000;
  static constexpr double am_min = 0.01, am_max = 600.;
  double const_ = 7. / 120 * cst::pi * cst::pi * cst::pi * cst::pi;
  double const2 = 5. / 7 / (cst::pi * cst::pi);
  double zeta3 = 1.2020569031595942853997, zeta5 = 1.0369277551433699263313;
  double am_minp = am_min * 1.1, am_maxp = am_max * 0.9;
  double dlnam = 0;
  std::vector<double> r1, dr1;  // 1-based
  void nuRhoPres(double am, double& rhonu, double& pnu) const {  // :1612-1646
    const double qmax = 30.;
    const int nq = 100;
    double dum1[nq + 2], dum2[nq + 2];
    const double adq = qmax / nq;
    dum1[1] = 0; dum2[1] = 0;
    for (int i = 1; i <= nq; i++) {
      const double q = i * adq;
      const double aq = am / q;
      const double v = 1. / std::sqrt(1. + aq * aq);
      const double aqdn = adq * q * q * q / (std::exp(q) + 1.);
      dum1[i + 1] = aqdn / v;
      dum2[i + 1] = aqdn * v;
    }
    rhonu = splint(dum1, nq + 1);
    pnu = splint(dum2, nq + 1);
    rhonu = (rhonu + dum1[nq + 1] / adq) / const_;
    pnu = (pnu + dum2[nq + 1] / adq) / const_ / 3.;
  }
  void init() {
    if (!r1.empty()) return;
    r1.assign(nrhopn + 1, 0.0);
    dlnam = -(std::log(am_min / am_max)) / (nrhopn - 1);
    for (int i = 1; i <= nrhopn; i++) {
      const double am = am_min * std::exp((i - 1) * dlnam);
      double rhonu, pnu;
      nuRhoPres(am, rhonu, pnu);
      r1[i] = std::log(rhonu);
    }
    std::vector<double> g;
    splini(g, nrhopn);
    splder(r1, dr1, nrhopn, g);
  }
  double Nu_rho(double am) const {  // :1687-1718
    if (am <= am_minp) return 1. + const2 * am * am;
    if (am >= am_maxp) return 3 / (2 * const_) * (zeta3 * am + (15 * zeta5) / 2 / am);
    double d = std::log(am / am_min) / dlnam + 1.;
    const int i = (int)d;
    d = d - i;
    const double rhonu = r1[i] + d * (dr1[i] + d * (3. * (r1[i + 1] - r1[i]) - 2. * dr1[i] - dr1[i + 1] +
                                                    d * (dr1[i] + dr1[i + 1] + 2. * (r1[i] - r1[i + 1]))));
    return std::exp(rhonu);
  }
};

// bg[16] = H0, omegab, omegac, omegan, omegav, w_lam, tcmb, nu_massless_degeneracy, n_eigenstates,
//          nu_mass_degeneracies[3], nu_mass_fractions[3], rdrag
struct Background {
  double H0, omegab, omegac, omegan, omegav, omegak, w_lam, tcmb;
  double grhom, grhog, grhor, grhoc, grhob, grhov, grhok, grhornomass, grhormass[3], nu_masses[3];
  int n_eig;
  bool flat, closed;
  double r, Ksign;
  const NuTable* nu;
  void set(const double* bg, const NuTable* table) {  // CAMBParams_Set, modules.f90:335-375 ; Nu_init :1545-1548
    nu = table;
    H0 = bg[0]; omegab = bg[1]; omegac = bg[2]; omegan = bg[3]; omegav = bg[4]; w_lam = bg[5]; tcmb = bg[6];
    const double nu_massless_degeneracy = bg[7];
    n_eig = (int)bg[8];
    omegak = 1 - (omegab + omegac + omegav + omegan);
    flat = std::fabs(omegak) <= 5e-7;
    closed = omegak < -5e-7;
    if (flat) { r = 1; Ksign = 0; }
    else {
      const double t = (cst::c / 1000) / H0;
      const double curv = -omegak / (t * t);
      Ksign = curv > 0 ? 1. : -1.;
      r = 1. / std::sqrt(std::fabs(curv));
    }
    grhom = 3 * H0 * H0 / (cst::c * cst::c) * 1000 * 1000;
    grhog = cst::kappa / (cst::c * cst::c) * 4 * cst::sigma_boltz / (cst::c * cst::c * cst::c) * (tcmb * tcmb * tcmb * tcmb) *
            (cst::Mpc * cst::Mpc);
    grhor = 7. / 8 * std::pow(4. / 11, 4. / 3) * grhog;
    grhornomass = grhor * nu_massless_degeneracy;
    for (int i = 0; i < 3; i++) { grhormass[i] = 0; nu_masses[i] = 0; }
    for (int i = 0; i < n_eig; i++) {
      grhormass[i] = grhor * bg[9 + i];
      nu_masses[i] = nu->const_ / (1.5 * nu->zeta3) * grhom / grhor * omegan * bg[12 + i] / bg[9 + i];
    }
    grhoc = grhom * omegac; grhob = grhom * omegab; grhov = grhom * omegav; grhok = grhom * omegak;
  }
  double dtauda(double a) const {  // equations.f90:69-100
    const double a2 = a * a;
    double grhoa2 = grhok * a2 + (grhoc + grhob) * a + grhog + grhornomass;
    if (w_lam == -1.) grhoa2 = grhoa2 + grhov * a2 * a2;
    else grhoa2 = grhoa2 + grhov * std::pow(a, 1 - 3 * w_lam);
    for (int i = 0; i < n_eig; i++) grhoa2 = grhoa2 + nu->Nu_rho(a * nu_masses[i]) * grhormass[i];
    return std::sqrt(3 / grhoa2);
  }
  double rofchi(double chi) const {  // modules.f90:494-508
    if (closed) return std::sin(chi);
    if (!flat) return std::sinh(chi);
    return chi;
  }
  double DeltaTime(double a1, double a2, double atol = 1e-4 / 1000) const {
    return rombint([this](double a) { return dtauda(a); }, a1, a2, atol);
  }
  double ComovingRadialDistance(double z) const { return DeltaTime(1 / (1 + z), 1.); }
  double AngularDiameterDistance(double z) const { return r / (1 + z) * rofchi(ComovingRadialDistance(z) / r); }
  double Hofz(double z) const {
    const double a = 1 / (1 + z);
    return 1 / (a * a * dtauda(a));
  }
  double BAO_D_v(double z) const {
    const double ADD = AngularDiameterDistance(z) * (1. + z);
    return std::pow(ADD * ADD * z / Hofz(z), 1. / 3.);
  }
  double tau0() const { return DeltaTime(0., 1.); }
  double age_gyr() const {  // DeltaPhysicalTimeGyr(0,1)
    return rombint([this](double a) { return dtauda(a) * a; }, 0., 1., 1e-4) * cst::Mpc / cst::c / cst::Gyr;
  }
  double CosmomcTheta() const {  // modules.f90:729-751
    const double h = H0 / 100.0;
    const double ombh2 = omegab * h * h, omdmh2 = (omegac + omegan) * h * h;
    const double zstar = 1048 * (1 + 0.00124 * std::pow(ombh2, -0.738)) *
                         (1 + (0.0783 * std::pow(ombh2, -0.238) / (1 + 39.5 * std::pow(ombh2, 0.763))) *
                                  std::pow(omdmh2 + ombh2, 0.560 / (1 + 21.1 * std::pow(ombh2, 1.81))));
    const double astar = 1 / (1 + zstar);
    const double atol = (double)1e-6f;
    const double rs = rombint(
        [this, h](double a) {
          const double R = 3.0e4 * a * omegab * (h * h);
          const double cs = 1.0 / std::sqrt(3 * (1 + R));
          return dtauda(a) * cs;
        },
        1e-8, astar, atol);
    const double DA = AngularDiameterDistance(zstar) / astar;
    return rs / DA;
  }
};

}  // namespace orc
