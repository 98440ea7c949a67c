// ORACLE (test infrastructure, NOT product code): CPU restatement of CAMB's thermal history, the step
// right before the hot path for every slow point.  Only tests/, smoke() and bench.py's CPU-baseline legs use it.
//
//   dverk (Verner 6(5) Runge-Kutta, default options)       camb/subroutines.f90:370-1128
//   RECFAST 1.5.2: Recombination_init / GET_INIT / ION      camb/recfast.f90:460-1010 ; Recombination_xe :434-456
//   Reionization (tanh in (1+z)^1.5, second He reionisation) camb/reionization.f90:61-99,139-199,256-315
//   rombint2                                                camb/subroutines.f90:52-114
//   inithermo, find_z, optdepth, dragoptdepth, ddamping_da  camb/modules.f90:2682-2992, 3086-3180
//   dsound_da_exact                                         camb/modules.f90:698-707
//   GetTauStart, dtaurec                                    camb/cmbmain.f90:629-655, 742-768
//   theta -> H0 bisection, zre from tau                     source/CosmologyParameterizations.f90:114-187
//
// Pinned (tests/test_thermo_oracle.py) to the derived-parameter block of the reference's golden
// data/base_plikHM_TTTEEE_lowl_lowE.minimum: zrei, zstar, rstar, 100 thetastar, DAstar, zdrag, rdrag, kd, 100 thetad,
// zeq, keq, 100 thetaeq, 100 thetarseq and H0 from 100 theta_MC.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "orc_bg.hpp"
#include "orc_core.hpp"

namespace orc {

namespace cst {  // camb/constants.f90:15-60 (the ones recombination needs)
constexpr double h_P = 6.62606896e-34;
constexpr double sigma_thomson = 6.6524616e-29;
constexpr double k_B = 1.3806504e-23;
constexpr double m_p = 1.672621637e-27;
constexpr double m_H = 1.673575e-27;
constexpr double m_e = 9.10938215e-31;
constexpr double mass_ratio_He_H = 3.9715;
constexpr double MPC_in_sec = Mpc / c;
constexpr double barssc0 = k_B / m_p / (c * c);
constexpr double a_rad = 8. * pi * pi * pi * pi * pi * k_B * k_B * k_B * k_B / 15 / (c * c * c) / (h_P * h_P * h_P);
constexpr double Compton_CT = MPC_in_sec * (8.0 / 3.0) * (sigma_thomson / (m_e * c)) * a_rad;
}  // namespace cst

// ---- rombint2 (camb/subroutines.f90:52-114): Romberg with a minimum number of steps ----
template <class F>
inline double rombint2(F f, double a, double b, double tol, int maxit, int minsteps) {
  const int MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(a) + f(b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20, g0 = 0, g1, fourj;
  int i = 0;
  for (;;) {
    i++;
    if (i > maxit || ((i > 5 && std::fabs(error) < tol) && nint > minsteps)) break;
    g0 = 0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(a + (k + k - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    const int jmax = std::min(i, MAXJ);
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

// ---- dverk (camb/subroutines.f90:370-1128) with every option c(1..9) = 0, as RECFAST calls it.  The work array c
// persists between calls (ind = 1 on the first call, 3 afterwards): c[14] trial step, c[19] error estimate, c[23]
// successive failures, c[20] / c[21] end of the previous interval.
struct Dverk {
  double c[25];
  int ind = 1;
  Dverk() { for (double& v : c) v = 0; }
  // integrates y' = fcn(x, y) from x to xend; n <= 4.  returns false on the error exits (ind < 0)
  template <class F>
  bool run(int n, F fcn, double& x, double* y, double xend, double tol) {
    double w[10][4];  // w[j][k] = w(k+1, j)
    double temp;
    if (ind == 3) {
      if (c[21] != 0 && (x != c[20] || xend == c[20])) return false;
      c[21] = 0;
    } else {  // ind == 1
      for (int k = 1; k <= 9; k++) c[k] = 0;
      c[10] = std::ldexp(1.0, -56);
      c[11] = 1e-35;
      c[20] = x;
      for (int k = 21; k <= 24; k++) c[k] = 0;
    }
    for (;;) {  // 99999
      if (ind != 6) {
        fcn(x, y, w[1]);
        c[24] = c[24] + 1;
      }
      // 105: hmin
      c[13] = c[3];
      if (c[3] == 0) {
        temp = 0;
        for (int k = 0; k < n; k++) temp = std::max(temp, std::fabs(y[k]));
        c[12] = std::min(temp, 1.0);
        c[13] = 10 * std::max(c[11], c[10] * std::max(c[12] / tol, std::fabs(x)));
      }
      c[15] = 1;  // scale (c(5) = 0)
      c[16] = 2;  // hmax (c(6) = 0 and c(5) = 0)
      if (c[13] > c[16]) { ind = -2; return false; }
      if (ind <= 2) {
        c[14] = c[16] * std::pow(tol, 1.0 / 6.0);
      } else if (c[23] <= 1) {
        temp = 2 * c[14];
        if (tol < std::pow(2.0 / 0.9, 6) * c[19]) temp = 0.9 * std::pow(tol / c[19], 1.0 / 6.0) * c[14];
        c[14] = std::max(temp, 0.5 * c[14]);
      } else {
        c[14] = 0.5 * c[14];
      }
      c[14] = std::min(c[14], c[16]);
      c[14] = std::max(c[14], c[13]);
      // 1111
      if (c[14] >= std::fabs(xend - x)) {
        c[14] = std::fabs(xend - x);
        c[17] = xend;
      } else {
        c[14] = std::min(c[14], 0.5 * std::fabs(xend - x));
        c[17] = x + std::copysign(c[14], xend - x);
      }
      c[18] = c[17] - x;
      temp = c[18] / 1398169080000.0;
      for (int k = 0; k < n; k++) w[9][k] = y[k] + temp * w[1][k] * 233028180000.0;
      fcn(x + c[18] / 6.0, w[9], w[2]);
      for (int k = 0; k < n; k++) w[9][k] = y[k] + temp * (w[1][k] * 74569017600.0 + w[2][k] * 298276070400.0);
      fcn(x + c[18] * (4.0 / 15.0), w[9], w[3]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (w[1][k] * 1165140900000.0 - w[2][k] * 3728450880000.0 + w[3][k] * 3495422700000.0);
      fcn(x + c[18] * (2.0 / 3.0), w[9], w[4]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (-w[1][k] * 3604654659375.0 + w[2][k] * 12816549900000.0 - w[3][k] * 9284716546875.0 +
                                 w[4][k] * 1237962206250.0);
      fcn(x + c[18] * (5.0 / 6.0), w[9], w[5]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (w[1][k] * 3355605792000.0 - w[2][k] * 11185352640000.0 + w[3][k] * 9172628850000.0 -
                                 w[4][k] * 427218330000.0 + w[5][k] * 482505408000.0);
      fcn(x + c[18], w[9], w[6]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (-w[1][k] * 770204740536.0 + w[2][k] * 2311639545600.0 - w[3][k] * 1322092233000.0 -
                                 w[4][k] * 453006781920.0 + w[5][k] * 326875481856.0);
      fcn(x + c[18] / 15.0, w[9], w[7]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (w[1][k] * 2845924389000.0 - w[2][k] * 9754668000000.0 + w[3][k] * 7897110375000.0 -
                                 w[4][k] * 192082660000.0 + w[5][k] * 400298976000.0 + w[7][k] * 201586000000.0);
      fcn(x + c[18], w[9], w[8]);
      for (int k = 0; k < n; k++)
        w[9][k] = y[k] + temp * (w[1][k] * 104862681000.0 + w[3][k] * 545186250000.0 + w[4][k] * 446637345000.0 +
                                 w[5][k] * 188806464000.0 + w[7][k] * 15076875000.0 + w[8][k] * 97599465000.0);
      c[24] = c[24] + 7;
      for (int k = 0; k < n; k++)
        w[2][k] = (w[1][k] * 8738556750.0 + w[3][k] * 9735468750.0 - w[4][k] * 9709507500.0 + w[5][k] * 8582112000.0 +
                   w[6][k] * 95329710000.0 - w[7][k] * 15076875000.0 - w[8][k] * 97599465000.0) / 1398169080000.0;
      temp = 0;
      for (int k = 0; k < n; k++) temp = std::max(temp, std::fabs(w[2][k]) / std::max(1.0, std::fabs(y[k])));
      c[19] = temp * c[14] * c[15];
      ind = 5;
      if (c[19] > tol) ind = 6;
      // 2222
      if (ind != 6) {
        x = c[17];
        for (int k = 0; k < n; k++) y[k] = w[9][k];
        c[22] = c[22] + 1;
        c[23] = 0;
        if (x == xend) {
          ind = 3;
          c[20] = xend;
          c[21] = 1;
          return true;
        }
      } else {
        c[23] = c[23] + 1;
        if (!(c[14] > c[13])) { ind = -3; return false; }
      }
    }
  }
};

// ---- RECFAST (camb/recfast.f90) ----
struct Recfast {
  static constexpr int Nz = 10000;
  static constexpr double zinitial = 1e4, zfinal = 0;
  static constexpr double delta_z = (zinitial - zfinal) / Nz;
  // RECDATA (recfast.f90:207-254)
  static constexpr double Lambda = 8.2245809, Lambda_He = 51.3, L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6;
  static constexpr double L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7, L_He_2s = 1.66277434e7, L_He_2p = 1.71134891e7;
  static constexpr double A2P_s = 1.798287e9, A2P_t = 177.58, L_He_2Pt = 1.690871466e7, L_He_2St = 1.5985597526e7;
  static constexpr double L_He2St_ion = 3.8454693845e6, sigma_He_2Ps = 1.436289e-22, sigma_He_2Pt = 1.484872e-22;
  static constexpr double bigH = 100.0e3 / cst::Mpc, not4 = cst::mass_ratio_He_H;
  // fudges (recfast.f90:266-281): RECFAST_Hswitch = T -> fudge 1.125 ; Heswitch 6 ; fudge_He 0.86
  double RECFAST_fudge = 1.105 + 0.02, RECFAST_fudge_He = 0.86;
  int RECFAST_Heswitch = 6;
  bool RECFAST_Hswitch = true;
  static constexpr double AGauss1 = -0.14, AGauss2 = 0.079, zGauss1 = 7.28, zGauss2 = 6.73, wGauss1 = 0.18, wGauss2 = 0.33;

  const Background* bg = nullptr;
  double Tnow, HO, OmegaK, OmegaT, z_eq, mu_H, mu_T, fHe, Nnow, fu, H_frac;
  double Lalpha, Lalpha_He, DeltaB, CDB, DeltaB_He, CDB_He, CB1, CB1_He1, CB1_He2, CR, CK, CK_He, CL, CL_He, CT, Bfact;
  std::vector<double> zrec, xrec, dxrec;
  double recombination_saha_z = 0;
  long n_fcn = 0;

  void ION(double z, const double* y, double* f) {  // recfast.f90:778-1010
    n_fcn++;
    using namespace cst;
    const double a_PPB = 4.309, b_PPB = -0.6166, c_PPB = 0.6703, d_PPB = 0.5300;
    const double a_VF = std::pow(10.0, -16.744), b_VF = 0.711, T_0 = std::pow(10.0, 0.477121), T_1 = std::pow(10.0, 5.114);
    const double a_trip = std::pow(10.0, -16.306), b_trip = 0.761;
    const double x_H = y[0], x_He = y[1], x = x_H + fHe * x_He, Tmat = y[2];
    const double n = Nnow * ((1 + z) * (1 + z) * (1 + z)), n_He = fHe * Nnow * ((1 + z) * (1 + z) * (1 + z));
    const double Trad = Tnow * (1 + z);
    const double Hz = 1 / bg->dtauda(1 / (1 + z)) * ((1 + z) * (1 + z)) / MPC_in_sec;
    const double Rdown = 1e-19 * a_PPB * std::pow(Tmat / 1e4, b_PPB) / (1 + c_PPB * std::pow(Tmat / 1e4, d_PPB));
    const double Rup = Rdown * std::pow(CR * Tmat, 1.5) * std::exp(-CDB / Tmat);
    const double sq_0 = std::sqrt(Tmat / T_0), sq_1 = std::sqrt(Tmat / T_1);
    double Rdown_He = a_VF / (sq_0 * std::pow(1 + sq_0, 1 - b_VF));
    Rdown_He = Rdown_He / std::pow(1 + sq_1, 1 + b_VF);
    double Rup_He = Rdown_He * std::pow(CR * Tmat, 1.5) * std::exp(-CDB_He / Tmat);
    Rup_He = 4 * Rup_He;
    const double He_Boltz = (Bfact / Tmat > 680) ? std::exp(680.0) : std::exp(Bfact / Tmat);
    double K;
    if (!RECFAST_Hswitch) K = CK / Hz;
    else {
      const double l1 = (std::log(1 + z) - zGauss1) / wGauss1, l2 = (std::log(1 + z) - zGauss2) / wGauss2;
      K = CK / Hz * (1.0 + AGauss1 * std::exp(-(l1 * l1)) + AGauss2 * std::exp(-(l2 * l2)));
    }
    double Rdown_trip = a_trip / (sq_0 * std::pow(1 + sq_0, 1.0 - b_trip));
    Rdown_trip = Rdown_trip / std::pow(1 + sq_1, 1 + b_trip);
    double Rup_trip = Rdown_trip * std::exp(-h_P * c * L_He2St_ion / (k_B * Tmat));
    Rup_trip = Rup_trip * std::pow(CR * Tmat, 1.5) * (4.0 / 3.0);
    const int Heflag = (x_He < 5e-9 || x_He > 0.98) ? 0 : RECFAST_Heswitch;
    double K_He, CfHe_t = 0;
    if (Heflag == 0) K_He = CK_He / Hz;
    else {
      const double tauHe_s = A2P_s * CK_He * 3 * n_He * (1 - x_He) / Hz;
      const double pHe_s = (1 - std::exp(-tauHe_s)) / tauHe_s;
      K_He = 1 / (A2P_s * pHe_s * 3 * n_He * (1 - x_He));
      if ((Heflag == 2 || Heflag >= 5) && x_H < 0.9999999) {
        double Doppler = 2 * k_B * Tmat / (m_H * not4 * c * c);
        Doppler = c * L_He_2p * std::sqrt(Doppler);
        const double gamma_2Ps = 3 * A2P_s * fHe * (1 - x_He) * c * c /
                                 (std::sqrt(cst::pi) * sigma_He_2Ps * 8 * cst::pi * Doppler * (1 - x_H)) / ((c * L_He_2p) * (c * L_He_2p));
        const double pb = 0.36, qb = RECFAST_fudge_He;
        const double AHcon = A2P_s / (1 + pb * std::pow(gamma_2Ps, qb));
        K_He = 1 / ((A2P_s * pHe_s + AHcon) * 3 * n_He * (1 - x_He));
      }
      if (Heflag >= 3) {
        double tauHe_t = A2P_t * n_He * (1 - x_He) * 3;
        tauHe_t = tauHe_t / (8 * cst::pi * Hz * (L_He_2Pt * L_He_2Pt * L_He_2Pt));
        const double pHe_t = (1 - std::exp(-tauHe_t)) / tauHe_t;
        const double CL_PSt = h_P * c * (L_He_2Pt - L_He_2St) / k_B;
        if (Heflag == 3 || Heflag == 5 || x_H > 0.99999) {
          CfHe_t = A2P_t * pHe_t * std::exp(-CL_PSt / Tmat);
          CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
        } else {
          double Doppler = 2 * k_B * Tmat / (m_H * not4 * c * c);
          Doppler = c * L_He_2Pt * std::sqrt(Doppler);
          const double gamma_2Pt = 3 * A2P_t * fHe * (1 - x_He) * c * c /
                                   (std::sqrt(cst::pi) * sigma_He_2Pt * 8 * cst::pi * Doppler * (1 - x_H)) /
                                   ((c * L_He_2Pt) * (c * L_He_2Pt));
          const double pb = 0.66, qb = 0.9;
          const double AHcon = A2P_t / (1 + pb * std::pow(gamma_2Pt, qb)) / 3;
          CfHe_t = (A2P_t * pHe_t + AHcon) * std::exp(-CL_PSt / Tmat);
          CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
        }
      }
    }
    const double timeTh = (1 / (CT * (Trad * Trad * Trad * Trad))) * (1 + x + fHe) / x;
    const double timeH = 2. / (3. * HO * std::pow(1 + z, 1.5));
    if (x_H > 0.99) f[0] = 0;
    else if (x_H > 0.985) {
      f[0] = (x * x_H * n * Rdown - Rup * (1 - x_H) * std::exp(-CL / Tmat)) / (Hz * (1 + z));
      recombination_saha_z = z;
    } else {
      f[0] = ((x * x_H * n * Rdown - Rup * (1 - x_H) * std::exp(-CL / Tmat)) * (1 + K * Lambda * n * (1 - x_H))) /
             (Hz * (1 + z) * (1 / fu + K * Lambda * n * (1 - x_H) / fu + K * Rup * n * (1 - x_H)));
    }
    if (x_He < 1e-15) f[1] = 0;
    else {
      f[1] = ((x * x_He * n * Rdown_He - Rup_He * (1 - x_He) * std::exp(-CL_He / Tmat)) *
              (1 + K_He * Lambda_He * n_He * (1 - x_He) * He_Boltz)) /
             (Hz * (1 + z) * (1 + K_He * (Lambda_He + Rup_He) * n_He * (1 - x_He) * He_Boltz));
      if (Heflag >= 3)
        f[1] = f[1] + (x * x_He * n * Rdown_trip - (1 - x_He) * 3 * Rup_trip * std::exp(-h_P * c * L_He_2St / (k_B * Tmat))) *
                          CfHe_t / (Hz * (1 + z));
    }
    if (timeTh < H_frac * timeH) {
      const double dHdz = (HO * HO / 2 / Hz) * (4 * ((1 + z) * (1 + z) * (1 + z)) / (1 + z_eq) * OmegaT +
                                                3 * OmegaT * ((1 + z) * (1 + z)) + 2 * OmegaK * (1 + z));
      const double epsilon = Hz * (1 + x + fHe) / (CT * (Trad * Trad * Trad) * x);
      f[2] = Tnow + epsilon * ((1 + fHe) / (1 + fHe + x)) * ((f[0] + fHe * f[1]) / x) - epsilon * dHdz / Hz +
             3.0 * epsilon / (1 + z);
    } else {
      f[2] = CT * (Trad * Trad * Trad * Trad) * x / (1 + x + fHe) * (Tmat - Trad) / (Hz * (1 + z)) + 2 * Tmat / (1 + z);
    }
  }

  // Recombination_init (recfast.f90:460-722); Omegav only enters the curvature of the Tmat approximation
  void init(const Background& B, double yp) {
    using namespace cst;
    bg = &B;
    const double OmegaB = B.omegab, OmegaC = B.omegac, OmegaV = B.omegav;
    Tnow = B.tcmb;
    OmegaT = OmegaC + OmegaB;
    OmegaK = 1 - OmegaT - OmegaV;
    const double H = B.H0 / 100;
    HO = H * bigH;
    mu_H = 1 / (1 - yp);
    mu_T = not4 / (not4 - (not4 - 1) * yp);
    fHe = yp / (not4 * (1 - yp));
    Nnow = 3 * HO * HO * OmegaB / (8 * cst::pi * G * mu_H * m_H);
    const double fnu = (21.0 / 8.0) * std::pow(4.0 / 11.0, 4.0 / 3.0);
    z_eq = (3 * (HO * c) * (HO * c) / (8 * cst::pi * G * a_rad * (1 + fnu) * (Tnow * Tnow * Tnow * Tnow))) * (OmegaB + OmegaC);
    z_eq = z_eq - 1;
    Lalpha = 1 / L_H_alpha;
    Lalpha_He = 1 / L_He_2p;
    DeltaB = h_P * c * (L_H_ion - L_H_alpha);
    CDB = DeltaB / k_B;
    DeltaB_He = h_P * c * (L_He1_ion - L_He_2s);
    CDB_He = DeltaB_He / k_B;
    CB1 = h_P * c * L_H_ion / k_B;
    CB1_He1 = h_P * c * L_He1_ion / k_B;
    CB1_He2 = h_P * c * L_He2_ion / k_B;
    CR = 2 * cst::pi * (m_e / h_P) * (k_B / h_P);
    CK = (Lalpha * Lalpha * Lalpha) / (8 * cst::pi);
    CK_He = (Lalpha_He * Lalpha_He * Lalpha_He) / (8 * cst::pi);
    CL = c * h_P / (k_B * Lalpha);
    CL_He = c * h_P / (k_B / L_He_2s);
    CT = Compton_CT / MPC_in_sec;
    Bfact = h_P * c * (L_He_2p - L_He_2s) / k_B;
    H_frac = 1e-3;
    fu = RECFAST_fudge;
    double z = zinitial, y[4];
    y[2] = Tnow * (1 + z);
    y[3] = y[2];
    double x_H0, x_He0, x0;
    {  // GET_INIT at z = zinitial > 8000 (recfast.f90:725-774)
      x_H0 = 1; x_He0 = 1; x0 = 1 + 2 * fHe;
    }
    y[0] = x_H0; y[1] = x_He0;
    Dverk dv;
    zrec.assign(Nz, 0); xrec.assign(Nz, 0); dxrec.assign(Nz, 0);
    const double tol = 1e-5;
    auto fcn = [this](double zz, const double* yy, double* ff) { ION(zz, yy, ff); };
    for (int i = 1; i <= Nz; i++) {
      double zstart = zinitial - (double)(i - 1) * delta_z;
      const double zend = zinitial - (double)i * delta_z;
      z = zend;
      if (zend > 8000) {
        x_H0 = 1; x_He0 = 1; x0 = 1 + 2 * fHe;
        y[0] = x_H0; y[1] = x_He0; y[2] = Tnow * (1 + z); y[3] = y[2];
      } else if (z > 5000) {
        x_H0 = 1; x_He0 = 1;
        double rhs = std::exp(1.5 * std::log(CR * Tnow / (1 + z)) - CB1_He2 / (Tnow * (1 + z))) / Nnow;
        rhs = rhs * 1;
        x0 = 0.5 * (std::sqrt((rhs - 1 - fHe) * (rhs - 1 - fHe) + 4 * (1 + 2 * fHe) * rhs) - (rhs - 1 - fHe));
        y[0] = x_H0; y[1] = x_He0; y[2] = Tnow * (1 + z); y[3] = y[2];
      } else if (z > 3500) {
        x_H0 = 1; x_He0 = 1; x0 = x_H0 + fHe * x_He0;
        y[0] = x_H0; y[1] = x_He0; y[2] = Tnow * (1 + z); y[3] = y[2];
      } else if (y[1] > 0.99) {
        x_H0 = 1;
        double rhs = std::exp(1.5 * std::log(CR * Tnow / (1 + z)) - CB1_He1 / (Tnow * (1 + z))) / Nnow;
        rhs = rhs * 4;
        x_He0 = 0.5 * (std::sqrt((rhs - 1) * (rhs - 1) + 4 * (1 + fHe) * rhs) - (rhs - 1));
        x0 = x_He0;
        x_He0 = (x0 - 1) / fHe;
        y[0] = x_H0; y[1] = x_He0; y[2] = Tnow * (1 + z); y[3] = y[2];
      } else if (y[0] > 0.99) {
        const double rhs = std::exp(1.5 * std::log(CR * Tnow / (1 + z)) - CB1 / (Tnow * (1 + z))) / Nnow;
        x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4 * rhs) - rhs);
        dv.run(3, fcn, zstart, y, zend, tol);
        y[0] = x_H0;
        x0 = y[0] + fHe * y[1];
        y[3] = y[2];
      } else {
        dv.run(3, fcn, zstart, y, zend, tol);
        x0 = y[0] + fHe * y[1];
      }
      zrec[i - 1] = zend;
      xrec[i - 1] = x0;
    }
    spline(zrec.data(), xrec.data(), Nz, 1.0e40, 1.0e40, dxrec.data());
  }

  double xe(double a) const {  // Recombination_xe (recfast.f90:434-456)
    const double z = 1 / a - 1;
    if (z >= zrec[0]) return xrec[0];
    if (z <= zrec[Nz - 1]) return xrec[Nz - 1];
    const double zst = (zinitial - z) / delta_z;
    const int ihi = (int)zst, ilo = ihi + 1;   // 1-based indices of the reference
    const double az = zst - (int)zst, bz = 1 - az;
    return az * xrec[ilo - 1] + bz * xrec[ihi - 1] +
           ((az * az * az - az) * dxrec[ilo - 1] + (bz * bz * bz - bz) * dxrec[ihi - 1]) / 6;
  }
};

// ---- Reionization (camb/reionization.f90) ----
struct Reion {
  bool Reionization = true;
  double redshift = 10, delta_redshift = 0.5, fraction = -1;
  double helium_redshift = 3.5, helium_delta_redshift = 0.5, helium_redshiftstart = 5.0;
  double tau_start = 0, tau_complete = 0, akthom = 0, fHe = 0, WindowVarMid = 0, WindowVarDelta = 0;
  static constexpr double zexp = 1.5, maxz = 50, tol = 1e-5;
  static constexpr bool include_helium_fullreion = true;
  void set_for_zre() {  // Reionization_SetParamsForZre
    WindowVarMid = std::pow(1 + redshift, zexp);
    WindowVarDelta = zexp * std::pow(1 + redshift, zexp - 1) * delta_redshift;
  }
  double xe(double a, double xstart = 0) const {  // Reionization_xe
    double xod = (WindowVarMid - 1 / std::pow(a, zexp)) / WindowVarDelta;
    double tgh = xod > 100 ? 1.0 : std::tanh(xod);
    double r = (fraction - xstart) * (tgh + 1) / 2 + xstart;
    if (include_helium_fullreion && a > (1 / (1 + helium_redshiftstart))) {
      xod = (1 + helium_redshift - 1 / a) / helium_delta_redshift;
      tgh = xod > 100 ? 1.0 : std::tanh(xod);
      r = r + fHe * (tgh + 1) / 2;
    }
    return r;
  }
  double opt_depth(const Background& B) const {  // Reionization_GetOptDepth
    return rombint2([&](double z) { const double a = 1 / (1 + z); return xe(a) * akthom * B.dtauda(a); }, 0.0, maxz, tol, 20,
                    (int)std::lround(maxz / delta_redshift * 5));
  }
  // Reionization_Init with use_optical_depth = F (CosmoMC hands CAMB the redshift), or = T (GetZreFromTau)
  bool init(const Background& B, double yhe, double akthom_, double tau0, double zre, double optical_depth) {
    akthom = akthom_;
    fHe = yhe / (cst::mass_ratio_He_H * (1 - yhe));
    tau_start = tau0; tau_complete = tau0;
    Reionization = true;
    const bool use_od = optical_depth > 0;
    redshift = zre;
    if ((use_od && optical_depth < 0.001) || (!use_od && redshift < 0.001)) Reionization = false;
    if (!Reionization) return true;
    if (fraction == -1) fraction = 1 + fHe;
    if (use_od) {  // Reionization_zreFromOptDepth: bisection on the redshift
      double try_b = 0, try_t = maxz, tau = 0;
      int i = 0;
      for (;;) {
        i++;
        redshift = (try_t + try_b) / 2;
        set_for_zre();
        tau = opt_depth(B);
        if (tau > optical_depth) try_t = redshift; else try_b = redshift;
        if (std::fabs(try_b - try_t) < 2e-3) break;
        if (i > 100) return false;
      }
      if (std::fabs(tau - optical_depth) > 0.002) return false;
    }
    set_for_zre();
    const double astart = 1.0 / (1.0 + redshift + delta_redshift * 8);
    tau_start = std::max(0.05, rombint([&](double a) { return B.dtauda(a); }, 0.0, astart, 1e-3));
    tau_complete = std::min(tau0, tau_start + rombint([&](double a) { return B.dtauda(a); }, astart,
                                                      1.0 / (1.0 + std::max(0.0, redshift - delta_redshift * 8)), 1e-3));
    return true;
  }
};

// ---- inithermo (camb/modules.f90:2682-2992) and what CAMBParams_Set / cmbmain do just before it ----
struct ThermoOut {
  double tau0 = 0, taurst = 0, taurend = 0, tau_start = 0, tau_complete = 0, dtaurec = 0, tau_maxvis = 0, zre = 0;
  double z_star = 0, z_drag = 0, actual_opt_depth = 0;
  // ThermoDerivedParams: age, zstar, rstar, thetastar, DAstar, zdrag, rdrag, kD, thetaD, zEQ, kEQ, thetaEQ, theta_rs_EQ
  double derived[13] = {0};
  int status = 0;
};

struct Thermo {
  static constexpr int nthermo = 20000;
  std::vector<double> tb, cs2, xe, dotmu, sdotmu, emmu, scaleFactor;  // 1-based
  Recfast rec;
  Reion reion;
  double akthom = 0, r_drag0 = 0, adotrad = 0, Nnow = 0;
  const Background* bg = nullptr;

  double doptdepth_dz(double z) const { const double a = 1 / (1 + z); return rec.xe(a) * akthom * bg->dtauda(a); }
  double optdepth(double z) const { return rombint2([this](double zz) { return doptdepth_dz(zz); }, 0.0, z, 1e-5, 20, 100); }
  double dragoptdepth(double z) const {
    return rombint2([this](double zz) { const double a = 1 / (1 + zz); return doptdepth_dz(zz) / r_drag0 / a; }, 0.0, z, 1e-5, 20, 100);
  }
  template <class F>
  static double find_z(F func, bool& ok) {  // modules.f90:3148-3178
    double try1 = 0, try2 = 10000, diff = 10, avg = 0;
    int i = 0;
    ok = true;
    while (diff > 1e-3) {
      i++;
      if (i == 100) { ok = false; return 0; }
      diff = func(try2) - func(try1);
      avg = 0.5 * (try2 + try1);
      if (func(avg) > 1) try2 = avg; else try1 = avg;
    }
    return avg;
  }

  // transfer_kmax_h: P%Transfer%kmax in h/Mpc when WantTransfer (CosmoMC: 5 with use_nonlinear_lensing, else 1.0), 0 = no transfer
  ThermoOut run(const Background& B, double yhe, double zre, double optical_depth, double max_eta_k, bool want_tensors,
                double transfer_kmax_h, double AccuracyBoost = 1.0) {
    using namespace cst;
    ThermoOut out;
    bg = &B;
    // CAMBParams_Set (modules.f90:376-400)
    double grhormass_sum = 0;
    for (int i = 0; i < B.n_eig; i++) grhormass_sum += B.grhormass[i];
    adotrad = std::sqrt((B.grhog + B.grhornomass + grhormass_sum) / 3);
    Nnow = B.omegab * (1 - yhe) * B.grhom * c * c / kappa / m_H / (Mpc * Mpc);
    akthom = sigma_thomson * Nnow * Mpc;
    const double tau0 = B.tau0();
    out.tau0 = tau0;
    if (!reion.init(B, yhe, akthom, tau0, zre, optical_depth)) { out.status = 1; return out; }
    out.zre = reion.redshift;
    out.tau_start = reion.tau_start; out.tau_complete = reion.tau_complete;
    // cmbmain set-up (cmbmain.f90:729-768), flat
    const double qmax = max_eta_k / tau0;
    double dtaurec = 4 / qmax / AccuracyBoost;
    double maxq = qmax;
    if (transfer_kmax_h > 0) maxq = std::max(qmax, transfer_kmax_h * (B.H0 / 100));
    double taumin = 0.001 / maxq;   // GetTauStart
    taumin = std::min(taumin, 0.1);
    if (B.n_eig > 0) {
      double mx = 0;
      for (int i = 0; i < B.n_eig; i++) mx = std::max(mx, B.nu_masses[i]);
      taumin = std::min(taumin, 1e-3 / mx / adotrad);
    }
    // ---- inithermo ----
    rec.init(B, yhe);
    tb.assign(nthermo + 1, 0); cs2.assign(nthermo + 1, 0); xe.assign(nthermo + 1, 0); dotmu.assign(nthermo + 1, 0);
    sdotmu.assign(nthermo + 1, 0); emmu.assign(nthermo + 1, 0); scaleFactor.assign(nthermo + 1, 0);
    double actual_opt_depth = 0, z_star = 0, z_drag = 0;
    int ncount = 0;
    const double thomc0 = Compton_CT * (B.tcmb * B.tcmb * B.tcmb * B.tcmb);
    r_drag0 = 3.0 / 4.0 * B.omegab * B.grhom / B.grhog;
    const double tauminn = 0.05 * taumin;
    const double dlntau = std::log(tau0 / tauminn) / (nthermo - 1);
    double last_dotmu = 0;
    double tau01 = tauminn, adot0 = adotrad, a0 = adotrad * tauminn;
    tb[1] = B.tcmb / a0;
    const double xe0 = 1, x1 = 0, x2 = 1;
    xe[1] = xe0 + 0.25 * yhe / (1 - yhe) * (x1 + 2 * x2);
    double barssc = barssc0 * (1 - 0.75 * yhe + (1 - yhe) * xe[1]);
    cs2[1] = 4. / 3. * barssc * tb[1];
    dotmu[1] = xe[1] * akthom / (a0 * a0);
    sdotmu[1] = 0;
    for (int i = 2; i <= nthermo; i++) {
      const double tau = tauminn * std::exp((i - 1) * dlntau);
      const double dtau = tau - tau01;
      double a = a0 + adot0 * dtau;
      scaleFactor[i] = a;
      const double a2 = a * a;
      const double adot = 1 / B.dtauda(a);
      a = a0 + 2 * dtau / (1 / adot0 + 1 / adot);
      const double tg0 = B.tcmb / a0;
      const double ahalf = 0.5 * (a0 + a), adothalf = 0.5 * (adot0 + adot);
      const double fe = (1 - yhe) * xe[i - 1] / (1 - 0.75 * yhe + (1 - yhe) * xe[i - 1]);
      const double thomc = thomc0 * fe / adothalf / (ahalf * ahalf * ahalf);
      const double etc = std::exp(-thomc * (a - a0));
      const double a2t = a0 * a0 * (tb[i - 1] - tg0) * etc - B.tcmb / thomc * (1 - etc);
      tb[i] = B.tcmb / a + a2t / (a * a);
      if (reion.Reionization && tau > reion.tau_start) {
        if (ncount == 0) ncount = i - 1;
        xe[i] = reion.xe(a, xe[ncount]);
        // CP%AccurateReionization and CP%DerivedParameters (both set by CosmoMC)
        dotmu[i] = (rec.xe(a) - xe[i]) * akthom / a2;
        if (last_dotmu != 0) actual_opt_depth = actual_opt_depth - 2 * dtau / (1 / dotmu[i] + 1 / last_dotmu);
        last_dotmu = dotmu[i];
      } else {
        xe[i] = rec.xe(a);
      }
      const double dtbdla = -2 * tb[i] - thomc * adothalf / adot * (a * tb[i] - B.tcmb);
      barssc = barssc0 * (1 - 0.75 * yhe + (1 - yhe) * xe[i]);
      cs2[i] = barssc * tb[i] * (1 - dtbdla / tb[i] / 3);
      dotmu[i] = xe[i] * akthom / a2;
      if (tau < 0.001) sdotmu[i] = 0;
      else sdotmu[i] = sdotmu[i - 1] + 2 * dtau / (1 / dotmu[i] + 1 / dotmu[i - 1]);
      a0 = a; tau01 = tau; adot0 = adot;
    }
    for (int j1 = 1; j1 <= nthermo; j1++) {
      if (sdotmu[j1] - sdotmu[nthermo] < -69) emmu[j1] = 1e-30;
      else {
        emmu[j1] = std::exp(sdotmu[j1] - sdotmu[nthermo]);
        if (z_star == 0) {
          if (sdotmu[nthermo] - sdotmu[j1] - actual_opt_depth < 1) {
            double t1 = 1 - (sdotmu[nthermo] - sdotmu[j1] - actual_opt_depth);
            t1 = t1 * (1 / dotmu[j1] + 1 / dotmu[j1 - 1]) / 2;
            z_star = 1 / (scaleFactor[j1] - t1 / B.dtauda(scaleFactor[j1])) - 1;
          }
        }
      }
    }
    int iv = 0, ns;
    double vfi = 0, cf1, maxvis = 0, taurst = 0, taurend = 0, tau_maxvis = 0;
    if (ncount == 0) { cf1 = 1; ns = nthermo; }
    else { cf1 = std::exp(sdotmu[nthermo] - sdotmu[ncount]); ns = ncount; }
    for (int j1 = 1; j1 <= ns; j1++) {
      const double vis = emmu[j1] * dotmu[j1];
      const double tau = tauminn * std::exp((j1 - 1) * dlntau);
      vfi = vfi + vis * cf1 * dlntau * tau;
      if (iv == 0 && vfi > 1.0e-7 / AccuracyBoost) {
        taurst = 9. / 10. * tau;
        iv = 1;
      } else if (iv == 1) {
        if (vis > maxvis) { maxvis = vis; tau_maxvis = tau; }
        if (vfi > 0.995) { taurend = tau; iv = 2; break; }
      }
    }
    if (iv != 2) { out.status = 2; return out; }
    if (want_tensors) dtaurec = std::min(dtaurec, taurst / 160) / AccuracyBoost;
    else dtaurec = std::min(dtaurec, taurst / 40) / AccuracyBoost;
    if (reion.Reionization) taurend = std::min(taurend, reion.tau_start);
    out.taurst = taurst; out.taurend = taurend; out.dtaurec = dtaurec; out.tau_maxvis = tau_maxvis;
    out.actual_opt_depth = actual_opt_depth;
    bool ok = true;
    if (z_star == 0) z_star = find_z([this](double z) { return optdepth(z); }, ok);
    if (!ok) { out.status = 3; return out; }
    z_drag = find_z([this](double z) { return dragoptdepth(z); }, ok);
    if (!ok) { out.status = 3; return out; }
    out.z_star = z_star; out.z_drag = z_drag;
    // derived parameters (modules.f90:2936-2952)
    auto dsound = [&B](double a) {
      const double R = 3 * B.grhob * a / (4 * B.grhog);
      return B.dtauda(a) * (1.0 / std::sqrt(3 * (1 + R)));
    };
    double rs = rombint(dsound, 1e-8, 1 / (z_star + 1), 1e-6);
    const double DA = B.AngularDiameterDistance(z_star) / (1 / (z_star + 1));
    double* d = out.derived;
    d[0] = B.age_gyr();
    d[1] = z_star; d[2] = rs; d[3] = 100 * rs / DA; d[4] = DA / 1000; d[5] = z_drag;
    rs = rombint(dsound, 1e-8, 1 / (z_drag + 1), 1e-6);
    d[6] = rs;
    auto ddamping = [this, &B](double a) {
      const double R = r_drag0 * a;
      return (R * R + 16 * (1 + R) / 15) / ((1 + R) * (1 + R)) * B.dtauda(a) * (a * a) / (rec.xe(a) * akthom);
    };
    d[7] = std::sqrt(1.0 / (rombint(ddamping, 1e-8, 1 / (z_star + 1), 1e-6) / 6));
    d[8] = 100 * cst::pi / d[7] / DA;
    const double z_eq = (B.grhob + B.grhoc) / (B.grhog + B.grhornomass + grhormass_sum) - 1;
    d[9] = z_eq;
    const double a_eq = 1 / (1 + z_eq);
    d[10] = 1 / (a_eq * B.dtauda(a_eq));
    d[11] = 100 * B.DeltaTime(0, a_eq) / DA;   // timeOfz(z_eq)
    d[12] = 100 * rombint(dsound, 1e-8, a_eq, 1e-6) / DA;
    return out;
  }
};

}  // namespace orc
