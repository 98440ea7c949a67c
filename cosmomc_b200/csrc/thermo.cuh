// K10 — thermal history batched over parameter points: the step CAMB runs right before the hot path for
// every slow point, and the source of r_drag / z_star / theta_star / z_re and of the time-grid scalars
// (tau0, taurst, taurend, reionisation start / end) that the projection kernels take as input.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   camb/subroutines.f90:370-1128   dverk (Verner 6(5) pair, default options, state kept between the 10^4 intervals)
//   camb/recfast.f90:460-1010       Recombination_init, GET_INIT, ION (RECFAST 1.5.2: H fudge 1.125 + double Gaussian,
//                                   He switch 6, fudge_He 0.86) ; :434-456 Recombination_xe
//   camb/subroutines.f90:253-296    spline (natural) ; :52-114 rombint2 ; :117-176 rombint
//   camb/reionization.f90:61-99,139-199,256-315  Reionization_xe / _Init / _zreFromOptDepth / _GetOptDepth
//   camb/modules.f90:376-400        Nnow, akthom, adotrad ; :2682-2992 inithermo ; :3086-3180 optdepth, dragoptdepth,
//                                   find_z, ddamping_da ; :698-707 dsound_da_exact
//   camb/cmbmain.f90:629-655,742-768  GetTauStart, dtaurec
//   source/CosmologyParameterizations.f90:134-176  theta -> H0 bisection (thermo_theta_kernel)
//
// Decomposition: the recombination ODE and the 20 000-step opacity recursion are sequential in time and three
// equations wide, so the parallel axis is the parameter point: one thread per point, per-point tables stored
// point-minor ([sample][point]) so that the lanes of a warp - which walk the same sample index - touch consecutive
// addresses.  Latency-bound by construction (~3.1e4 dependent derivative evaluations per point); what the GPU buys is
// 10^4-10^5 points in flight.
#pragma once
#include "background.cuh"
#include "common.cuh"

namespace cb200 {

constexpr int TH_NZ = 10000;       // RECFAST output redshifts (recfast.f90:263)
constexpr int TH_NTHERMO = 20000;  // inithermo samples (modules.f90:2602)
constexpr int TH_NOUT = 32;        // doubles per point in the result row
constexpr int TH_NIN = 8;          // doubles per point in the input row

namespace thc {  // camb/constants.f90
constexpr double h_P = 6.62606896e-34, sigma_thomson = 6.6524616e-29, k_B = 1.3806504e-23, m_p = 1.672621637e-27;
constexpr double m_H = 1.673575e-27, m_e = 9.10938215e-31, not4 = 3.9715;
constexpr double MPC_in_sec = bgc::Mpc / bgc::c;
constexpr double barssc0 = k_B / m_p / (bgc::c * bgc::c);
constexpr double a_rad = 8. * bgc::pi * bgc::pi * bgc::pi * bgc::pi * bgc::pi * k_B * k_B * k_B * k_B / 15 /
                         (bgc::c * bgc::c * bgc::c) / (h_P * h_P * h_P);
constexpr double Compton_CT = MPC_in_sec * (8.0 / 3.0) * (sigma_thomson / (m_e * bgc::c)) * a_rad;
// RECDATA (recfast.f90:207-254)
constexpr double Lambda = 8.2245809, Lambda_He = 51.3, L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6;
constexpr double L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7, L_He_2s = 1.66277434e7, L_He_2p = 1.71134891e7;
constexpr double A2P_s = 1.798287e9, A2P_t = 177.58, L_He_2Pt = 1.690871466e7, L_He_2St = 1.5985597526e7;
constexpr double L_He2St_ion = 3.8454693845e6, sigma_He_2Ps = 1.436289e-22, sigma_He_2Pt = 1.484872e-22;
constexpr double bigH = 100.0e3 / bgc::Mpc;
constexpr double zinitial = 1e4, delta_z = 1.0;
}  // namespace thc

struct ThermoScratch {  // per-chunk work arrays, point-minor: element (i, point) at [i * P + point]
  int P;
  double* xrec;    // [TH_NZ]
  double* dxrec;   // [TH_NZ]
  double* work;    // [TH_NZ] spline sweep
  double* dotmu;   // [TH_NTHERMO]
  double* sdotmu;  // [TH_NTHERMO]
  double* sfac;    // [TH_NTHERMO] scale factor of the first Friedmann estimate (scaleFactor of inithermo)
};

struct RecConst {  // Recombination_init's constants for one point
  double Tnow, HO, OmegaK, OmegaT, z_eq, fHe, Nnow, fu, fudge_He;
  double CDB, CDB_He, CB1, CB1_He1, CB1_He2, CR, CK, CK_He, CL, CL_He, CT, Bfact;
  int Heswitch;
};

__device__ inline RecConst rec_const(const double* bg, double yp) {
  using namespace thc;
  RecConst R;
  const double OmegaB = bg[1], OmegaC = bg[2], OmegaV = bg[4];
  R.Tnow = bg[6];
  R.OmegaT = OmegaC + OmegaB;
  R.OmegaK = 1 - R.OmegaT - OmegaV;
  const double H = bg[0] / 100;
  R.HO = H * bigH;
  const double mu_H = 1 / (1 - yp);
  R.fHe = yp / (not4 * (1 - yp));
  R.Nnow = 3 * R.HO * R.HO * OmegaB / (8 * bgc::pi * bgc::G * mu_H * m_H);
  const double fnu = (21.0 / 8.0) * pow(4.0 / 11.0, 4.0 / 3.0);
  R.z_eq = (3 * (R.HO * bgc::c) * (R.HO * bgc::c) / (8 * bgc::pi * bgc::G * a_rad * (1 + fnu) * (R.Tnow * R.Tnow * R.Tnow * R.Tnow))) *
           (OmegaB + OmegaC);
  R.z_eq = R.z_eq - 1;
  const double Lalpha = 1 / L_H_alpha, Lalpha_He = 1 / L_He_2p;
  R.CDB = h_P * bgc::c * (L_H_ion - L_H_alpha) / k_B;
  R.CDB_He = h_P * bgc::c * (L_He1_ion - L_He_2s) / k_B;
  R.CB1 = h_P * bgc::c * L_H_ion / k_B;
  R.CB1_He1 = h_P * bgc::c * L_He1_ion / k_B;
  R.CB1_He2 = h_P * bgc::c * L_He2_ion / k_B;
  R.CR = 2 * bgc::pi * (m_e / h_P) * (k_B / h_P);
  R.CK = (Lalpha * Lalpha * Lalpha) / (8 * bgc::pi);
  R.CK_He = (Lalpha_He * Lalpha_He * Lalpha_He) / (8 * bgc::pi);
  R.CL = bgc::c * h_P / (k_B * Lalpha);
  R.CL_He = bgc::c * h_P / (k_B / L_He_2s);
  R.CT = Compton_CT / MPC_in_sec;
  R.Bfact = h_P * bgc::c * (L_He_2p - L_He_2s) / k_B;
  R.fu = 1.105 + 0.02;   // RECFAST_fudge_default2 (RECFAST_Hswitch = T)
  R.fudge_He = 0.86;
  R.Heswitch = 6;
  return R;
}

// ION (recfast.f90:778-1010): derivatives of x_H, x_He, T_mat with respect to redshift
__device__ inline void rec_ion(const RecConst& R, const BgPoint& P, const BgTables& T, double z, const double* y, double* f) {
  using namespace thc;
  const double c = bgc::c, pi = bgc::pi;
  const double a_PPB = 4.309, b_PPB = -0.6166, c_PPB = 0.6703, d_PPB = 0.5300;
  // 10^-16.744, 10^0.477121, 10^5.114, 10^-16.306 (the reference evaluates the powers at run time)
  const double a_VF = 1.80301774085957e-17, b_VF = 0.711, T_0 = 2.999998240459423, T_1 = 130016.95780332899;
  const double a_trip = 4.9431068698683435e-17, b_trip = 0.761;
  const double AGauss1 = -0.14, AGauss2 = 0.079, zGauss1 = 7.28, zGauss2 = 6.73, wGauss1 = 0.18, wGauss2 = 0.33;
  const double fHe = R.fHe, Tnow = R.Tnow;
  const double x_H = y[0], x_He = y[1], x = x_H + fHe * x_He, Tmat = y[2];
  const double zp = 1 + z;
  const double n = R.Nnow * (zp * zp * zp), n_He = fHe * R.Nnow * (zp * zp * zp);
  const double Trad = Tnow * zp;
  const double Hz = 1 / dtauda(P, T, 1 / zp) * (zp * zp) / MPC_in_sec;
  // powers with real exponents as exp(b log x) on shared logarithms, x^1.5 as x sqrt(x): a third of the instructions of
  // nine pow() calls per evaluation, the same value to ~1e-15 (this chain is 3.1e4 evaluations long and latency-bound)
  const double lt4 = log(Tmat / 1e4);
  const double Rdown = 1e-19 * a_PPB * exp(b_PPB * lt4) / (1 + c_PPB * exp(d_PPB * lt4));
  const double crt = R.CR * Tmat;
  const double crt15 = crt * sqrt(crt);
  const double Rup = Rdown * crt15 * exp(-R.CDB / Tmat);
  const double sq_0 = sqrt(Tmat / T_0), sq_1 = sqrt(Tmat / T_1);
  const double l0 = log(1 + sq_0), l1s = log(1 + sq_1);
  double Rdown_He = a_VF / (sq_0 * exp((1 - b_VF) * l0));
  Rdown_He = Rdown_He / exp((1 + b_VF) * l1s);
  double Rup_He = Rdown_He * crt15 * exp(-R.CDB_He / Tmat);
  Rup_He = 4 * Rup_He;
  const double He_Boltz = (R.Bfact / Tmat > 680) ? exp(680.0) : exp(R.Bfact / Tmat);
  const double l1 = (log(zp) - zGauss1) / wGauss1, l2 = (log(zp) - zGauss2) / wGauss2;
  const double K = R.CK / Hz * (1.0 + AGauss1 * exp(-(l1 * l1)) + AGauss2 * exp(-(l2 * l2)));
  double Rdown_trip = a_trip / (sq_0 * exp((1.0 - b_trip) * l0));
  Rdown_trip = Rdown_trip / exp((1 + b_trip) * l1s);
  double Rup_trip = Rdown_trip * exp(-h_P * c * L_He2St_ion / (k_B * Tmat));
  Rup_trip = Rup_trip * crt15 * (4.0 / 3.0);
  const int Heflag = (x_He < 5e-9 || x_He > 0.98) ? 0 : R.Heswitch;
  double K_He, CfHe_t = 0;
  if (Heflag == 0) K_He = R.CK_He / Hz;
  else {
    const double tauHe_s = A2P_s * R.CK_He * 3 * n_He * (1 - x_He) / Hz;
    const double pHe_s = (1 - exp(-tauHe_s)) / tauHe_s;
    K_He = 1 / (A2P_s * pHe_s * 3 * n_He * (1 - x_He));
    if ((Heflag == 2 || Heflag >= 5) && x_H < 0.9999999) {
      double Doppler = 2 * k_B * Tmat / (m_H * not4 * c * c);
      Doppler = c * L_He_2p * sqrt(Doppler);
      const double gamma_2Ps = 3 * A2P_s * fHe * (1 - x_He) * c * c / (sqrt(pi) * sigma_He_2Ps * 8 * pi * Doppler * (1 - x_H)) /
                               ((c * L_He_2p) * (c * L_He_2p));
      const double AHcon = A2P_s / (1 + 0.36 * exp(R.fudge_He * log(gamma_2Ps)));
      K_He = 1 / ((A2P_s * pHe_s + AHcon) * 3 * n_He * (1 - x_He));
    }
    if (Heflag >= 3) {
      double tauHe_t = A2P_t * n_He * (1 - x_He) * 3;
      tauHe_t = tauHe_t / (8 * pi * Hz * (L_He_2Pt * L_He_2Pt * L_He_2Pt));
      const double pHe_t = (1 - exp(-tauHe_t)) / tauHe_t;
      const double CL_PSt = h_P * c * (L_He_2Pt - L_He_2St) / k_B;
      if (Heflag == 3 || Heflag == 5 || x_H > 0.99999) {
        CfHe_t = A2P_t * pHe_t * exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      } else {
        double Doppler = 2 * k_B * Tmat / (m_H * not4 * c * c);
        Doppler = c * L_He_2Pt * sqrt(Doppler);
        const double gamma_2Pt = 3 * A2P_t * fHe * (1 - x_He) * c * c / (sqrt(pi) * sigma_He_2Pt * 8 * pi * Doppler * (1 - x_H)) /
                                 ((c * L_He_2Pt) * (c * L_He_2Pt));
        const double AHcon = A2P_t / (1 + 0.66 * exp(0.9 * log(gamma_2Pt))) / 3;
        CfHe_t = (A2P_t * pHe_t + AHcon) * exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      }
    }
  }
  const double timeTh = (1 / (R.CT * (Trad * Trad * Trad * Trad))) * (1 + x + fHe) / x;
  const double timeH = 2. / (3. * R.HO * (zp * sqrt(zp)));
  if (x_H > 0.99) f[0] = 0;
  else if (x_H > 0.985) f[0] = (x * x_H * n * Rdown - Rup * (1 - x_H) * exp(-R.CL / Tmat)) / (Hz * zp);
  else
    f[0] = ((x * x_H * n * Rdown - Rup * (1 - x_H) * exp(-R.CL / Tmat)) * (1 + K * Lambda * n * (1 - x_H))) /
           (Hz * zp * (1 / R.fu + K * Lambda * n * (1 - x_H) / R.fu + K * Rup * n * (1 - x_H)));
  if (x_He < 1e-15) f[1] = 0;
  else {
    f[1] = ((x * x_He * n * Rdown_He - Rup_He * (1 - x_He) * exp(-R.CL_He / Tmat)) *
            (1 + K_He * Lambda_He * n_He * (1 - x_He) * He_Boltz)) /
           (Hz * zp * (1 + K_He * (Lambda_He + Rup_He) * n_He * (1 - x_He) * He_Boltz));
    if (Heflag >= 3)
      f[1] = f[1] + (x * x_He * n * Rdown_trip - (1 - x_He) * 3 * Rup_trip * exp(-h_P * c * L_He_2St / (k_B * Tmat))) * CfHe_t / (Hz * zp);
  }
  if (timeTh < 1e-3 * timeH) {
    const double dHdz = (R.HO * R.HO / 2 / Hz) * (4 * (zp * zp * zp) / (1 + R.z_eq) * R.OmegaT + 3 * R.OmegaT * (zp * zp) + 2 * R.OmegaK * zp);
    const double epsilon = Hz * (1 + x + fHe) / (R.CT * (Trad * Trad * Trad) * x);
    f[2] = Tnow + epsilon * ((1 + fHe) / (1 + fHe + x)) * ((f[0] + fHe * f[1]) / x) - epsilon * dHdz / Hz + 3.0 * epsilon / zp;
  } else {
    f[2] = R.CT * (Trad * Trad * Trad * Trad) * x / (1 + x + fHe) * (Tmat - Trad) / (Hz * zp) + 2 * Tmat / zp;
  }
}

// dverk with every option at its default (c(1..9) = 0), state kept across calls as the reference keeps cw(24) and ind
struct DverkState {
  double h_trial, err_est, fails, x_last;  // c(14), c(19), c(23), c(20)
  int ind, done;                           // ind = 1 before the first call, 3 after a completed one ; c(21)
};

__device__ inline bool rec_dverk(DverkState& S, const RecConst& R, const BgPoint& P, const BgTables& T, double& x, double* y,
                                 double xend, double tol) {
  double w1[3], w2[3], w3[3], w4[3], w5[3], w6[3], w7[3], w8[3], w9[3];
  if (S.ind == 3) {
    if (S.done && (x != S.x_last || xend == S.x_last)) return false;
    S.done = 0;
  } else {
    S.x_last = x; S.done = 0; S.fails = 0; S.err_est = 0; S.h_trial = 0;
  }
  int ind = S.ind;
  for (;;) {
    if (ind != 6) rec_ion(R, P, T, x, y, w1);
    double temp = fmax(fmax(fabs(y[0]), fabs(y[1])), fabs(y[2]));
    const double c12 = fmin(temp, 1.0);
    const double hmin = 10 * fmax(1e-35, ldexp(1.0, -56) * fmax(c12 / tol, fabs(x)));
    const double hmax = 2;
    if (hmin > hmax) return false;
    if (ind <= 2) S.h_trial = hmax * pow(tol, 1.0 / 6.0);
    else if (S.fails <= 1) {
      temp = 2 * S.h_trial;
      if (tol < pow(2.0 / 0.9, 6.0) * S.err_est) temp = 0.9 * pow(tol / S.err_est, 1.0 / 6.0) * S.h_trial;
      S.h_trial = fmax(temp, 0.5 * S.h_trial);
    } else S.h_trial = 0.5 * S.h_trial;
    S.h_trial = fmin(S.h_trial, hmax);
    S.h_trial = fmax(S.h_trial, hmin);
    double xtrial;
    if (S.h_trial >= fabs(xend - x)) { S.h_trial = fabs(xend - x); xtrial = xend; }
    else { S.h_trial = fmin(S.h_trial, 0.5 * fabs(xend - x)); xtrial = x + copysign(S.h_trial, xend - x); }
    const double h = xtrial - x;
    temp = h / 1398169080000.0;
#pragma unroll
    for (int k = 0; k < 3; k++) w9[k] = y[k] + temp * w1[k] * 233028180000.0;
    rec_ion(R, P, T, x + h / 6.0, w9, w2);
#pragma unroll
    for (int k = 0; k < 3; k++) w9[k] = y[k] + temp * (w1[k] * 74569017600.0 + w2[k] * 298276070400.0);
    rec_ion(R, P, T, x + h * (4.0 / 15.0), w9, w3);
#pragma unroll
    for (int k = 0; k < 3; k++) w9[k] = y[k] + temp * (w1[k] * 1165140900000.0 - w2[k] * 3728450880000.0 + w3[k] * 3495422700000.0);
    rec_ion(R, P, T, x + h * (2.0 / 3.0), w9, w4);
#pragma unroll
    for (int k = 0; k < 3; k++)
      w9[k] = y[k] + temp * (-w1[k] * 3604654659375.0 + w2[k] * 12816549900000.0 - w3[k] * 9284716546875.0 + w4[k] * 1237962206250.0);
    rec_ion(R, P, T, x + h * (5.0 / 6.0), w9, w5);
#pragma unroll
    for (int k = 0; k < 3; k++)
      w9[k] = y[k] + temp * (w1[k] * 3355605792000.0 - w2[k] * 11185352640000.0 + w3[k] * 9172628850000.0 - w4[k] * 427218330000.0 +
                             w5[k] * 482505408000.0);
    rec_ion(R, P, T, x + h, w9, w6);
#pragma unroll
    for (int k = 0; k < 3; k++)
      w9[k] = y[k] + temp * (-w1[k] * 770204740536.0 + w2[k] * 2311639545600.0 - w3[k] * 1322092233000.0 - w4[k] * 453006781920.0 +
                             w5[k] * 326875481856.0);
    rec_ion(R, P, T, x + h / 15.0, w9, w7);
#pragma unroll
    for (int k = 0; k < 3; k++)
      w9[k] = y[k] + temp * (w1[k] * 2845924389000.0 - w2[k] * 9754668000000.0 + w3[k] * 7897110375000.0 - w4[k] * 192082660000.0 +
                             w5[k] * 400298976000.0 + w7[k] * 201586000000.0);
    rec_ion(R, P, T, x + h, w9, w8);
#pragma unroll
    for (int k = 0; k < 3; k++)
      w9[k] = y[k] + temp * (w1[k] * 104862681000.0 + w3[k] * 545186250000.0 + w4[k] * 446637345000.0 + w5[k] * 188806464000.0 +
                             w7[k] * 15076875000.0 + w8[k] * 97599465000.0);
    temp = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double e = (w1[k] * 8738556750.0 + w3[k] * 9735468750.0 - w4[k] * 9709507500.0 + w5[k] * 8582112000.0 +
                        w6[k] * 95329710000.0 - w7[k] * 15076875000.0 - w8[k] * 97599465000.0) / 1398169080000.0;
      temp = fmax(temp, fabs(e) / fmax(1.0, fabs(y[k])));
    }
    S.err_est = temp * S.h_trial;
    ind = (S.err_est > tol) ? 6 : 5;
    if (ind != 6) {
      x = xtrial;
#pragma unroll
      for (int k = 0; k < 3; k++) y[k] = w9[k];
      S.fails = 0;
      if (x == xend) {
        S.ind = 3; S.x_last = xend; S.done = 1;
        return true;
      }
    } else {
      S.fails = S.fails + 1;
      if (!(S.h_trial > hmin)) return false;
    }
  }
}

// Recombination_xe (recfast.f90:434-456) from this point's table
__device__ __forceinline__ double rec_xe(const ThermoScratch& W, int pt, double a) {
  const double z = 1 / a - 1;
  const size_t P = (size_t)W.P;
  if (z >= thc::zinitial - thc::delta_z) return W.xrec[pt];
  if (z <= 0) return W.xrec[(size_t)(TH_NZ - 1) * P + pt];
  const double zst = (thc::zinitial - z) / thc::delta_z;
  const int ihi = (int)zst, ilo = ihi + 1;
  const double az = zst - (int)zst, bz = 1 - az;
  return az * W.xrec[(size_t)(ilo - 1) * P + pt] + bz * W.xrec[(size_t)(ihi - 1) * P + pt] +
         ((az * az * az - az) * W.dxrec[(size_t)(ilo - 1) * P + pt] + (bz * bz * bz - bz) * W.dxrec[(size_t)(ihi - 1) * P + pt]) / 6;
}

// rombint / rombint2 (subroutines.f90:52-176) for a callable; minsteps < 0: plain rombint
template <class F>
__device__ inline double th_rombint(F f, double a, double b, double tol, int maxit = 20, int minsteps = -1) {
  const int MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(a) + f(b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20, g0 = 0;
  int i = 0;
  for (;;) {
    i++;
    if (i > maxit || ((i > 5 && fabs(error) < tol) && nint > minsteps)) break;
    g0 = 0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(a + (k + k - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    const int jmax = min(i, MAXJ);
    double fourj = 1;
#pragma unroll
    for (int j = 1; j <= MAXJ; j++) {
      if (j <= jmax) {
        fourj = 4 * fourj;
        const double g1 = g0 + (g0 - g[j]) / (fourj - 1);
        g[j] = g0;
        g0 = g1;
      }
    }
    if (fabs(g0) > tol) error = 1 - gmax / g0;
    else error = gmax;
    gmax = g0;
#pragma unroll
    for (int j = 1; j <= MAXJ + 1; j++)
      if (j == jmax + 1) g[j] = g0;
  }
  return g0;
}

struct ReionDev {  // camb/reionization.f90
  double redshift, delta_redshift, fraction, fHe, mid, delta, helium_redshift, helium_delta, helium_start;
  __device__ void set_for_zre() {
    mid = pow(1 + redshift, 1.5);
    delta = 1.5 * pow(1 + redshift, 0.5) * delta_redshift;
  }
  __device__ double xe(double a, double xstart) const {
    double xod = (mid - 1 / pow(a, 1.5)) / delta;
    double tgh = xod > 100 ? 1.0 : tanh(xod);
    double r = (fraction - xstart) * (tgh + 1) / 2 + xstart;
    if (a > (1 / (1 + helium_start))) {
      xod = (1 + helium_redshift - 1 / a) / helium_delta;
      tgh = xod > 100 ? 1.0 : tanh(xod);
      r = r + fHe * (tgh + 1) / 2;
    }
    return r;
  }
};

// ------------------------------------------------------------------------------------------------------------------
// Kernel A: RECFAST for every point of the chunk -> xrec, dxrec (point-minor).  tin[pt][TH_NIN] = yhe, zre, optical
// depth (> 0: z_re by bisection), max_eta_k, want_tensors, transfer kmax [h/Mpc], AccuracyBoost, reserved.
__global__ void __launch_bounds__(64) thermo_recfast_kernel(int np, const double* __restrict__ bg, const double* __restrict__ tin,
                                                            BgTables T, ThermoScratch W, int* __restrict__ status) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= np) return;
  const double* b = bg + (size_t)pt * NBG;
  const BgPoint P = bg_point(b);
  const double yp = tin[(size_t)pt * TH_NIN];
  const RecConst R = rec_const(b, yp);
  const size_t PS = (size_t)W.P;
  const double Tnow = R.Tnow, fHe = R.fHe;
  double y[3] = {1.0, 1.0, Tnow * (1 + thc::zinitial)};
  DverkState S;
  S.ind = 1; S.done = 0; S.fails = 0; S.err_est = 0; S.h_trial = 0; S.x_last = 0;
  int bad = 0;
  for (int i = 1; i <= TH_NZ; i++) {
    double zstart = thc::zinitial - (double)(i - 1) * thc::delta_z;
    const double zend = thc::zinitial - (double)i * thc::delta_z;
    const double z = zend;
    double x0;
    if (zend > 8000) {
      x0 = 1 + 2 * fHe;
      y[0] = 1; y[1] = 1; y[2] = Tnow * (1 + z);
    } else if (z > 5000) {
      double rhs = exp(1.5 * log(R.CR * Tnow / (1 + z)) - R.CB1_He2 / (Tnow * (1 + z))) / R.Nnow;
      x0 = 0.5 * (sqrt((rhs - 1 - fHe) * (rhs - 1 - fHe) + 4 * (1 + 2 * fHe) * rhs) - (rhs - 1 - fHe));
      y[0] = 1; y[1] = 1; y[2] = Tnow * (1 + z);
    } else if (z > 3500) {
      x0 = 1 + fHe;
      y[0] = 1; y[1] = 1; y[2] = Tnow * (1 + z);
    } else if (y[1] > 0.99) {
      double rhs = exp(1.5 * log(R.CR * Tnow / (1 + z)) - R.CB1_He1 / (Tnow * (1 + z))) / R.Nnow;
      rhs = rhs * 4;
      x0 = 0.5 * (sqrt((rhs - 1) * (rhs - 1) + 4 * (1 + fHe) * rhs) - (rhs - 1));
      y[0] = 1; y[1] = (x0 - 1) / fHe; y[2] = Tnow * (1 + z);
    } else if (y[0] > 0.99) {
      const double rhs = exp(1.5 * log(R.CR * Tnow / (1 + z)) - R.CB1 / (Tnow * (1 + z))) / R.Nnow;
      const double x_H0 = 0.5 * (sqrt(rhs * rhs + 4 * rhs) - rhs);
      if (!rec_dverk(S, R, P, T, zstart, y, zend, 1e-5)) bad = 1;
      y[0] = x_H0;
      x0 = y[0] + fHe * y[1];
    } else {
      if (!rec_dverk(S, R, P, T, zstart, y, zend, 1e-5)) bad = 1;
      x0 = y[0] + fHe * y[1];
    }
    W.xrec[(size_t)(i - 1) * PS + pt] = x0;
    if (bad) break;
  }
  if (bad) { status[pt] = 2; return; }   // error_recombination
  // natural spline of xrec over the redshifts zrec(i) = zinitial - i (subroutines.f90:253-296); u in W.work
  {
    double* d2 = W.dxrec + pt;
    double* u = W.work + pt;
    const double* yv = W.xrec + pt;
    auto X = [](int i) { return thc::zinitial - (double)i * thc::delta_z; };  // x(i), 1-based
    double d1r = (yv[PS] - yv[0]) / (X(2) - X(1)), d1l;
    d2[0] = 0; u[0] = 0;
    double d2p = 0, up = 0;
    for (int i = 2; i <= TH_NZ - 1; i++) {
      d1l = d1r;
      d1r = (yv[(size_t)i * PS] - yv[(size_t)(i - 1) * PS]) / (X(i + 1) - X(i));
      const double xxdiv = 1 / (X(i + 1) - X(i - 1));
      const double sig = (X(i) - X(i - 1)) * xxdiv;
      const double xp = 1 / (sig * d2p + 2);
      d2p = (sig - 1) * xp;
      up = (6 * (d1r - d1l) * xxdiv - sig * up) * xp;
      d2[(size_t)(i - 1) * PS] = d2p;
      u[(size_t)(i - 1) * PS] = up;
    }
    double nxt = (0.0 - 0.0 * up) / (0.0 * d2p + 1);   // d2(n) with qn = un = 0
    d2[(size_t)(TH_NZ - 1) * PS] = nxt;
    for (int i = TH_NZ - 1; i >= 1; i--) {
      nxt = d2[(size_t)(i - 1) * PS] * nxt + u[(size_t)(i - 1) * PS];
      d2[(size_t)(i - 1) * PS] = nxt;
    }
  }
  status[pt] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel B: reionisation set-up, inithermo's opacity recursion, z_star / z_drag and the derived parameters.
// out[pt][TH_NOUT] = tau0, taurst, taurend, reion tau_start, tau_complete, dtaurec, tau_maxvis, zre, z_star, z_drag,
// actual_opt_depth, status, derived[13] (age, zstar, rstar, 100 thetastar, DAstar, zdrag, rdrag, kD, 100 thetaD, zEQ, kEQ,
// 100 thetaEQ, 100 theta_rs_EQ).
__global__ void __launch_bounds__(64) thermo_init_kernel(int np, const double* __restrict__ bg, const double* __restrict__ tin,
                                                         BgTables T, ThermoScratch W, int* __restrict__ status,
                                                         double* __restrict__ out) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= np) return;
  double* o = out + (size_t)pt * TH_NOUT;
  for (int i = 0; i < TH_NOUT; i++) o[i] = 0;
  if (status[pt] != 0) { o[11] = status[pt]; return; }
  const double* b = bg + (size_t)pt * NBG;
  const double* in = tin + (size_t)pt * TH_NIN;
  const BgPoint P = bg_point(b);
  const size_t PS = (size_t)W.P;
  const double yhe = in[0], max_eta_k = in[3], AccuracyBoost = in[6] > 0 ? in[6] : 1.0;
  const bool want_tensors = in[4] != 0;
  const double H0 = b[0], omegab = b[1], omegac = b[2], tcmb = b[6];
  // CAMBParams_Set (modules.f90:335-400)
  const double grhom = 3 * H0 * H0 / (bgc::c * bgc::c) * 1000 * 1000;
  const double grhog = bgc::kappa / (bgc::c * bgc::c) * 4 * bgc::sigma_boltz / (bgc::c * bgc::c * bgc::c) * (tcmb * tcmb * tcmb * tcmb) *
                       (bgc::Mpc * bgc::Mpc);
  const double grhob = grhom * omegab, grhoc = grhom * omegac;
  double grhormass_sum = 0, nu_mass_max = 0;
  for (int i = 0; i < P.n_eig; i++) { grhormass_sum += P.grhormass[i]; nu_mass_max = fmax(nu_mass_max, P.nu_masses[i]); }
  const double grho_rad = P.grhog_nm + grhormass_sum;   // grhog + grhornomass + sum(grhormass)
  const double adotrad = sqrt(grho_rad / 3);
  const double Nnow = omegab * (1 - yhe) * grhom * bgc::c * bgc::c / bgc::kappa / thc::m_H / (bgc::Mpc * bgc::Mpc);
  const double akthom = thc::sigma_thomson * Nnow * bgc::Mpc;
  const double tau0 = bg_rombint<0>(P, T, 0.0, 1.0, 1e-4 / 1000);
  auto dtau_da = [&](double a) { return dtauda(P, T, a); };
  // ---- Reionization_Init ----
  ReionDev RI;
  RI.delta_redshift = 0.5; RI.helium_redshift = 3.5; RI.helium_delta = 0.5; RI.helium_start = 5.0;
  RI.fHe = yhe / (thc::not4 * (1 - yhe));
  RI.fraction = 1 + RI.fHe;
  RI.redshift = in[1];
  const double optical_depth = in[2];
  bool reionization = true;
  double tau_start = tau0, tau_complete = tau0;
  if ((optical_depth > 0 && optical_depth < 0.001) || (!(optical_depth > 0) && RI.redshift < 0.001)) reionization = false;
  if (reionization) {
    if (optical_depth > 0) {  // Reionization_zreFromOptDepth
      double try_b = 0, try_t = 50, tau = 0;
      int it = 0, fail = 0;
      for (;;) {
        it++;
        RI.redshift = (try_t + try_b) / 2;
        RI.set_for_zre();
        tau = th_rombint([&](double z) { const double a = 1 / (1 + z); return RI.xe(a, 0.0) * akthom * dtauda(P, T, a); }, 0.0, 50.0, 1e-5, 20,
                         (int)lround(50.0 / RI.delta_redshift * 5));
        if (tau > optical_depth) try_t = RI.redshift; else try_b = RI.redshift;
        if (fabs(try_b - try_t) < 2e-3) break;
        if (it > 100) { fail = 1; break; }
      }
      if (fail || fabs(tau - optical_depth) > 0.002) { o[11] = 1; status[pt] = 1; return; }   // error_reionization
    }
    RI.set_for_zre();
    const double astart = 1.0 / (1.0 + RI.redshift + RI.delta_redshift * 8);
    tau_start = fmax(0.05, th_rombint(dtau_da, 0.0, astart, 1e-3));
    tau_complete = fmin(tau0, tau_start + th_rombint(dtau_da, astart, 1.0 / (1.0 + fmax(0.0, RI.redshift - RI.delta_redshift * 8)), 1e-3));
  }
  // ---- cmbmain set-up (cmbmain.f90:729-768), flat ----
  const double qmax = max_eta_k / tau0;
  double dtaurec = 4 / qmax / AccuracyBoost;
  double maxq = qmax;
  if (in[5] > 0) maxq = fmax(qmax, in[5] * (H0 / 100));
  double taumin = fmin(0.001 / maxq, 0.1);
  if (P.n_eig > 0) taumin = fmin(taumin, 1e-3 / nu_mass_max / adotrad);
  // ---- inithermo (modules.f90:2682-2992) ----
  const double thomc0 = thc::Compton_CT * (tcmb * tcmb * tcmb * tcmb);
  const double r_drag0 = 3.0 / 4.0 * omegab * grhom / grhog;
  const double tauminn = 0.05 * taumin;
  const double dlntau = log(tau0 / tauminn) / (TH_NTHERMO - 1);
  double actual_opt_depth = 0, last_dotmu = 0;
  int ncount = 0;
  double xe_ncount = 0;
  double tau01 = tauminn, adot0 = adotrad, a0 = adotrad * tauminn;
  double tb_prev = tcmb / a0;
  double xe_prev = 1.0 + 0.25 * yhe / (1 - yhe) * (0.0 + 2 * 1.0);
  double dotmu_prev = xe_prev * akthom / (a0 * a0), sdotmu_prev = 0;
  W.dotmu[pt] = dotmu_prev; W.sdotmu[pt] = 0; W.sfac[pt] = 0;
  for (int i = 2; i <= TH_NTHERMO; i++) {
    const double tau = tauminn * exp((i - 1) * dlntau);
    const double dtau = tau - tau01;
    double a = a0 + adot0 * dtau;
    W.sfac[(size_t)(i - 1) * PS + pt] = a;
    const double a2 = a * a;
    const double adot = 1 / dtauda(P, T, a);
    a = a0 + 2 * dtau / (1 / adot0 + 1 / adot);
    const double tg0 = tcmb / a0;
    const double ahalf = 0.5 * (a0 + a), adothalf = 0.5 * (adot0 + adot);
    const double fe = (1 - yhe) * xe_prev / (1 - 0.75 * yhe + (1 - yhe) * xe_prev);
    const double thomc = thomc0 * fe / adothalf / (ahalf * ahalf * ahalf);
    const double etc = exp(-thomc * (a - a0));
    const double a2t = a0 * a0 * (tb_prev - tg0) * etc - tcmb / thomc * (1 - etc);
    const double tb = tcmb / a + a2t / (a * a);
    double xe;
    if (reionization && tau > tau_start) {
      if (ncount == 0) { ncount = i - 1; xe_ncount = xe_prev; }
      xe = RI.xe(a, xe_ncount);
      const double dm = (rec_xe(W, pt, a) - xe) * akthom / a2;   // AccurateReionization and DerivedParameters
      if (last_dotmu != 0) actual_opt_depth = actual_opt_depth - 2 * dtau / (1 / dm + 1 / last_dotmu);
      last_dotmu = dm;
    } else {
      xe = rec_xe(W, pt, a);
    }
    const double dotmu = xe * akthom / a2;
    const double sdotmu = (tau < 0.001) ? 0.0 : sdotmu_prev + 2 * dtau / (1 / dotmu + 1 / dotmu_prev);
    W.dotmu[(size_t)(i - 1) * PS + pt] = dotmu;
    W.sdotmu[(size_t)(i - 1) * PS + pt] = sdotmu;
    a0 = a; tau01 = tau; adot0 = adot; tb_prev = tb; xe_prev = xe; dotmu_prev = dotmu; sdotmu_prev = sdotmu;
  }
  const double sd_last = sdotmu_prev;
  // z_star from the first sample with optical depth below one (AccurateReionization branch, modules.f90:2838-2844)
  double z_star = 0;
  for (int j1 = 2; j1 <= TH_NTHERMO && z_star == 0; j1++) {
    const double sd = W.sdotmu[(size_t)(j1 - 1) * PS + pt];
    if (!(sd - sd_last < -69) && sd_last - sd - actual_opt_depth < 1) {
      double t1 = 1 - (sd_last - sd - actual_opt_depth);
      t1 = t1 * (1 / W.dotmu[(size_t)(j1 - 1) * PS + pt] + 1 / W.dotmu[(size_t)(j1 - 2) * PS + pt]) / 2;
      const double sf = W.sfac[(size_t)(j1 - 1) * PS + pt];
      z_star = 1 / (sf - t1 / dtauda(P, T, sf)) - 1;
    }
  }
  // start / end of recombination and the maximum of the visibility (modules.f90:2853-2888)
  int iv = 0;
  double vfi = 0, maxvis = 0, taurst = 0, taurend = 0, tau_maxvis = 0;
  const int ns = ncount == 0 ? TH_NTHERMO : ncount;
  const double cf1 = ncount == 0 ? 1.0 : exp(sd_last - W.sdotmu[(size_t)(ncount - 1) * PS + pt]);
  for (int j1 = 1; j1 <= ns; j1++) {
    const double sd = W.sdotmu[(size_t)(j1 - 1) * PS + pt];
    const double emmu = (sd - sd_last < -69) ? 1e-30 : exp(sd - sd_last);
    const double vis = emmu * W.dotmu[(size_t)(j1 - 1) * PS + pt];
    const double tau = tauminn * exp((j1 - 1) * dlntau);
    vfi = vfi + vis * cf1 * dlntau * tau;
    if (iv == 0 && vfi > 1.0e-7 / AccuracyBoost) { taurst = 9. / 10. * tau; iv = 1; }
    else if (iv == 1) {
      if (vis > maxvis) { maxvis = vis; tau_maxvis = tau; }
      if (vfi > 0.995) { taurend = tau; iv = 2; break; }
    }
  }
  if (iv != 2) { o[11] = 1; status[pt] = 1; return; }
  dtaurec = fmin(dtaurec, taurst / (want_tensors ? 160 : 40)) / AccuracyBoost;
  if (reionization) taurend = fmin(taurend, tau_start);
  // z_star (if the table did not give it) and z_drag: optical depth one (modules.f90:3086-3178)
  auto doptdepth_dz = [&](double z) { const double a = 1 / (1 + z); return rec_xe(W, pt, a) * akthom * dtauda(P, T, a); };
  auto optdepth = [&](double z) { return th_rombint(doptdepth_dz, 0.0, z, 1e-5, 20, 100); };
  auto dragoptdepth = [&](double z) {
    return th_rombint([&](double zz) { return doptdepth_dz(zz) / r_drag0 * (1 + zz); }, 0.0, z, 1e-5, 20, 100);
  };
  bool ok = true;
  // find_z (modules.f90:3148-3178).  The reference evaluates func(try2), func(try1) and func(avg) in every pass; the
  // end points are always an earlier midpoint (or the initial bracket), so their values are carried along instead of
  // being integrated again - same numbers, a third of the Romberg integrals.
  auto find_z = [&](auto func) {
    double try1 = 0, try2 = 10000, diff = 10, avg = 0;
    double f1 = func(try1), f2 = func(try2);
    int it = 0;
    while (diff > 1e-3) {
      it++;
      if (it == 100) { ok = false; return 0.0; }
      diff = f2 - f1;
      avg = 0.5 * (try2 + try1);
      const double fa = func(avg);
      if (fa > 1) { try2 = avg; f2 = fa; } else { try1 = avg; f1 = fa; }
    }
    return avg;
  };
  if (z_star == 0) z_star = find_z(optdepth);
  const double z_drag = find_z(dragoptdepth);
  if (!ok) { o[11] = 1; status[pt] = 1; return; }
  // ---- derived parameters (modules.f90:2936-2952) ----
  auto dsound = [&](double a) {
    const double Rr = 3 * grhob * a / (4 * grhog);
    return dtauda(P, T, a) * (1.0 / sqrt(3 * (1 + Rr)));
  };
  double rs = th_rombint(dsound, 1e-8, 1 / (z_star + 1), 1e-6);
  const double DA = bg_angular_diameter_distance(P, T, z_star) / (1 / (z_star + 1));
  o[0] = tau0; o[1] = taurst; o[2] = taurend; o[3] = tau_start; o[4] = tau_complete; o[5] = dtaurec; o[6] = tau_maxvis;
  o[7] = reionization ? RI.redshift : 0.0; o[8] = z_star; o[9] = z_drag; o[10] = actual_opt_depth; o[11] = 0;
  double* d = o + 12;
  d[0] = bg_rombint<1>(P, T, 0.0, 1.0, 1e-4) * bgc::Mpc / bgc::c / bgc::Gyr;
  d[1] = z_star; d[2] = rs; d[3] = 100 * rs / DA; d[4] = DA / 1000; d[5] = z_drag;
  d[6] = th_rombint(dsound, 1e-8, 1 / (z_drag + 1), 1e-6);
  d[7] = sqrt(1.0 / (th_rombint(
                         [&](double a) {
                           const double Rr = r_drag0 * a;
                           return (Rr * Rr + 16 * (1 + Rr) / 15) / ((1 + Rr) * (1 + Rr)) * dtauda(P, T, a) * (a * a) / (rec_xe(W, pt, a) * akthom);
                         },
                         1e-8, 1 / (z_star + 1), 1e-6) /
                     6));
  d[8] = 100 * bgc::pi / d[7] / DA;
  const double z_eq = (grhob + grhoc) / grho_rad - 1;
  d[9] = z_eq;
  const double a_eq = 1 / (1 + z_eq);
  d[10] = 1 / (a_eq * dtauda(P, T, a_eq));
  d[11] = 100 * bg_rombint<0>(P, T, 0.0, a_eq, 1e-4 / 1000) / DA;
  d[12] = 100 * th_rombint(dsound, 1e-8, a_eq, 1e-6) / DA;
}

// ------------------------------------------------------------------------------------------------------------------
// theta -> H0 (source/CosmologyParameterizations.f90:134-176): bisection on H0 in [H0_min, H0_max] until successive
// CosmomcTheta values differ by < 1e-7.  cos[pt][8] = ombh2, omch2, omnuh2, omk, w, 100 theta_MC, H0_min, H0_max;
// neutrino split nu[pt][7] = massless degeneracy, n_eig, degeneracies[3] ... as in bg[7..15); writes bg rows with the
// solved H0 (H0 = 0: theta out of range, the point is rejected as the reference does).
__device__ inline void th_fill_bg(double* b, const double* cs, const double* nu, double H0, double tcmb) {
  const double h2 = (H0 / 100) * (H0 / 100);
  const double omb = cs[0] / h2, omc = cs[1] / h2, omnu = cs[2] / h2;
  const double omdm = (cs[1] + cs[2]) / h2;
  b[0] = H0; b[1] = omb; b[2] = omc; b[3] = omnu; b[4] = 1 - cs[3] - omb - omdm; b[5] = cs[4]; b[6] = tcmb;
  for (int i = 0; i < 8; i++) b[7 + i] = nu[i];
  if (omnu == 0) b[8] = 0;
}
__device__ inline double th_cosmomc_theta(const double* b, const BgTables& T) {  // modules.f90:729-751
  const BgPoint P = bg_point(b);
  const double h = b[0] / 100.0;
  const double ombh2 = b[1] * h * h, omdmh2 = (b[2] + b[3]) * h * h;
  const double zstar = 1048 * (1 + 0.00124 * pow(ombh2, -0.738)) *
                       (1 + (0.0783 * pow(ombh2, -0.238) / (1 + 39.5 * pow(ombh2, 0.763))) * pow(omdmh2 + ombh2, 0.560 / (1 + 21.1 * pow(ombh2, 1.81))));
  const double astar = 1 / (1 + zstar);
  const double rs = bg_rombint<2>(P, T, 1e-8, astar, (double)1e-6f);
  const double DA = bg_angular_diameter_distance(P, T, zstar) / astar;
  return rs / DA;
}
__global__ void thermo_theta_kernel(int np, const double* __restrict__ cs, const double* __restrict__ nu, double tcmb, BgTables T,
                                    double* __restrict__ bg) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= np) return;
  const double* c = cs + (size_t)pt * 8;
  const double* n = nu + (size_t)pt * 8;
  double* b = bg + (size_t)pt * NBG;
  const double rd = b[15];
  const double DAt = c[5] / 100;
  double try_b = c[6], try_t = c[7];
  th_fill_bg(b, c, n, try_b, tcmb);
  const double D_b = th_cosmomc_theta(b, T);
  th_fill_bg(b, c, n, try_t, tcmb);
  const double D_t = th_cosmomc_theta(b, T);
  double H0 = 0;
  if (!(DAt < D_b || DAt > D_t)) {
    double lasttry = -1;
    for (int it = 0; it < 200; it++) {
      H0 = (try_b + try_t) / 2;
      th_fill_bg(b, c, n, H0, tcmb);
      const double D_try = th_cosmomc_theta(b, T);
      if (D_try < DAt) try_b = (try_b + try_t) / 2; else try_t = (try_b + try_t) / 2;
      if (fabs(D_try - lasttry) < 1e-7) break;
      lasttry = D_try;
    }
  }
  if (H0 == 0) { for (int i = 0; i < NBG; i++) b[i] = 0; }
  b[15] = rd;
}

}  // namespace cb200
