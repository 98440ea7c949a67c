// Spherical-Bessel interpolation table of the flat line-of-sight projection, built on the device once per
// (l-set, kmax) and kept L2-resident.  Reference behaviour: camb/bessels.f90:50-120 (table + natural
// cubic-spline second derivatives in x) and :132-275 (the j_l(x) evaluator and its asymptotic regimes).
//
// Layout: node table  bes[i][lp]  (double2 {j_l(x_i), j_l''-spline(x_i)}), x index slow, multipole fast
// (padded to a multiple of 32) so that a warp whose lanes are consecutive multipoles reads 512 contiguous
// bytes per node.
#pragma once
#include "common.cuh"

namespace cb200 {

__device__ double sph_bessel_j(int L, double X) {
  const double LN2 = 0.6931471805599453094, ONEMLN2 = 0.30685281944005469058277;
  const double PID2 = 1.5707963267948966192313217, PID4 = 0.78539816339744830961566084582;
  const double ROOTPI12 = 21.269446210866192327578;
  const double GAMMA1 = 2.6789385347077476336556, GAMMA2 = 1.3541179394264004169452;
  const double ax = fabs(X), ax2 = ax * ax;
  double jl;
  if (L < 7) {
    const double s = sin(ax), c = cos(ax);
    switch (L) {
      case 0: jl = (ax < 1e-1) ? 1 - ax2 / 6 * (1 - ax2 / 20) : s / ax; break;
      case 1: jl = (ax < 2e-1) ? ax / 3 * (1 - ax2 / 10 * (1 - ax2 / 28)) : (s / ax - c) / ax; break;
      case 2:
        jl = (ax < 3e-1) ? ax2 / 15 * (1 - ax2 / 14 * (1 - ax2 / 36)) : (-3.0 * c / ax - s * (1 - 3 / ax2)) / ax;
        break;
      case 3:
        jl = (ax < 4e-1) ? ax * ax2 / 105 * (1 - ax2 / 18 * (1 - ax2 / 44))
                         : (c * (1 - 15 / ax2) - s * (6 - 15 / ax2) / ax) / ax;
        break;
      case 4:
        jl = (ax < 6e-1) ? ax2 * ax2 / 945 * (1 - ax2 / 22 * (1 - ax2 / 52))
                         : (s * (1 - (45 - 105 / ax2) / ax2) + c * (10 - 105 / ax2) / ax) / ax;
        break;
      case 5:
        jl = (ax < 1) ? ax2 * ax2 * ax / 10395 * (1 - ax2 / 26 * (1 - ax2 / 60))
                      : (s * (15 - (420 - 945 / ax2) / ax2) / ax - c * (1 - (105 - 945.0 / ax2) / ax2)) / ax;
        break;
      default:
        jl = (ax < 1) ? ax2 * ax2 * ax2 / 135135 * (1 - ax2 / 30 * (1 - ax2 / 68))
                      : (s * (-1 + (210 - (4725 - 10395 / ax2) / ax2) / ax2) +
                         c * (-21 + (1260 - 10395 / ax2) / ax2) / ax) / ax;
    }
  } else {
    const double nu = 0.5 + L, nu2 = nu * nu;
    if (ax < 1e-40) {
      jl = 0;
    } else if ((ax2 / L) < 5e-1) {  // x << l : series
      jl = exp(L * log(ax / nu) - LN2 + nu * ONEMLN2 - (1 - (1 - 3.5 / nu2) / nu2 / 30) / 12 / nu) / nu *
           (1 - ax2 / (4 * nu + 4) * (1 - ax2 / (8 * nu + 16) * (1 - ax2 / (12 * nu + 36))));
    } else if (((double)L * (double)L / ax) < 5e-1) {  // x >> l^2 : trigonometric asymptote
      const double beta = ax - PID2 * (L + 1);
      jl = (cos(beta) * (1 - (nu2 - 0.25) * (nu2 - 2.25) / 8 / ax2 * (1 - (nu2 - 6.25) * (nu2 - 12.25) / 48 / ax2)) -
            sin(beta) * (nu2 - 0.25) / 2 / ax *
                (1 - (nu2 - 2.25) * (nu2 - 6.25) / 24 / ax2 * (1 - (nu2 - 12.25) * (nu2 - 20.25) / 80 / ax2))) / ax;
    } else {
      // the reference writes the exponents/thresholds as single-precision literals (0.325, 1.31, 1.48)
      const double l3 = pow(nu, (double)0.325f);
      if (ax < nu - (double)1.31f * l3) {  // below the turning point: Debye, exponentially small
        const double cosb = nu / ax, sx = sqrt(nu2 - ax2), cotb = nu / sx, secb = ax / nu;
        const double beta = log(cosb + sx / ax);
        const double cot3b = cotb * cotb * cotb, cot6b = cot3b * cot3b, sec2b = secb * secb;
        const double expterm =
            ((2 + 3 * sec2b) * cot3b / 24 -
             ((4 + sec2b) * sec2b * cot6b / 16 +
              ((16 - (1512 + (3654 + 375 * sec2b) * sec2b) * sec2b) * cot3b / 5760 +
               (32 + (288 + (232 + 13 * sec2b) * sec2b) * sec2b) * sec2b * cot6b / 128 / nu) * cot6b / nu) / nu) / nu;
        jl = sqrt(cotb * cosb) / (2 * nu) * exp(-nu * beta + nu / cotb - expterm);
      } else if (ax > nu + (double)1.48f * l3) {  // above the turning point: Debye, oscillatory
        const double cosb = nu / ax, sx = sqrt(ax2 - nu2), cotb = nu / sx, secb = ax / nu;
        const double beta = acos(cosb);
        const double cot3b = cotb * cotb * cotb, cot6b = cot3b * cot3b, sec2b = secb * secb;
        const double trigarg = nu / cotb - nu * beta - PID4 -
                               ((2.0 + 3.0 * sec2b) * cot3b / 24 +
                                (16 - (1512 + (3654 + 375 * sec2b) * sec2b) * sec2b) * cot3b * cot6b / 5760 / nu2) / nu;
        const double expterm = ((4 + sec2b) * sec2b * cot6b / 16 -
                                (32 + (288 + (232 + 13 * sec2b) * sec2b) * sec2b) * sec2b * cot6b * cot6b / 128 / nu2) / nu2;
        jl = sqrt(cotb * cosb) / nu * exp(-expterm) * cos(trigarg);
      } else {  // transition region: Airy-type expansion
        const double beta = ax - nu, beta2 = beta * beta, sx = 6 / ax, sx2 = sx * sx;
        const double secb = pow(sx, 0.3333333333333333), sec2b = secb * secb;
        jl = (GAMMA1 * secb + beta * GAMMA2 * sec2b - (beta2 / 18 - 1.0 / 45) * beta * sx * secb * GAMMA1 -
              ((beta2 - 1) * beta2 / 36 + 1.0 / 420) * sx * sec2b * GAMMA2 +
              (((beta2 / 1620 - 7.0 / 3240) * beta2 + 1.0 / 648) * beta2 - 1.0 / 8100) * sx2 * secb * GAMMA1 +
              (((beta2 / 4536 - 1.0 / 810) * beta2 + 19.0 / 11340) * beta2 - 13.0 / 28350) * beta * sx2 * sec2b * GAMMA2 -
              ((((beta2 / 349920 - 1.0 / 29160) * beta2 + 71.0 / 583200) * beta2 - 121.0 / 874800) * beta2 +
               7939.0 / 224532000) * beta * sx2 * sx * secb * GAMMA1) * sqrt(sx) / ROOTPI12;
      }
    }
  }
  if (X < 0 && (L & 1)) jl = -jl;
  return jl;
}

// values: thread per (x index, multipole); zero below the x-cut (bessels.f90:94-110)
__global__ void bessel_values_kernel(int num_xx, int nl, int nlp, const double* __restrict__ x,
                                     const int* __restrict__ ls, double2* __restrict__ bes) {
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_xx || j >= nlp) return;
  double v = 0;
  if (j < nl) {
    const int l = ls[j];
    const double xi = x[i];
    double xlim = 0.05 * l;
    xlim = fmax(xlim, 35.0);
    xlim = l - xlim;
    if (xi > xlim) {
      bool tiny = (l == 3 && xi <= 0.2) || (l > 3 && xi < 0.5) || (l > 5 && xi < 1.0);
      if (!tiny) v = sph_bessel_j(l, xi);
    }
  }
  bes[(size_t)i * nlp + j] = make_double2(v, 0.0);
}

// natural cubic-spline second derivatives along x for each multipole (subroutines.f90:253-296):
// a first-order linear recurrence — one thread per multipole, scratch in global memory.
__global__ void bessel_spline_kernel(int num_xx, int nl, int nlp, const double* __restrict__ x,
                                     double2* __restrict__ bes, double* __restrict__ scratch /*[nl][num_xx]*/) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nl) return;
  double* u = scratch + (size_t)j * num_xx;
  auto Y = [&](int i) { return bes[(size_t)i * nlp + j].x; };
  auto D2 = [&](int i) -> double& { return bes[(size_t)i * nlp + j].y; };
  double d1r = (Y(1) - Y(0)) / (x[1] - x[0]), d1l;
  D2(0) = 0; u[0] = 0;
  double d2p = 0, up = 0;  // previous d2/u kept in registers: the chain never waits on global memory
  for (int i = 1; i <= num_xx - 2; i++) {
    d1l = d1r;
    d1r = (Y(i + 1) - Y(i)) / (x[i + 1] - x[i]);
    double xxdiv = 1. / (x[i + 1] - x[i - 1]);
    double sig = (x[i] - x[i - 1]) * xxdiv;
    double xp = 1. / (sig * d2p + 2.);
    d2p = (sig - 1.) * xp;
    up = (6. * (d1r - d1l) * xxdiv - sig * up) * xp;
    D2(i) = d2p;
    u[i] = up;
  }
  D2(num_xx - 1) = 0;  // natural end: (un - qn*u)/(qn*d2+1) with qn = un = 0
  double nxt = 0;
  for (int i = num_xx - 2; i >= 0; i--) { nxt = D2(i) * nxt + u[i]; D2(i) = nxt; }
}

}  // namespace cb200
