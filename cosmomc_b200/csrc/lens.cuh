// l-interpolation of the sampled C_l (K3) and curved-sky correlation-function lensing (K4), then the
// CosmoMC unit conversion.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   K3  camb/cmbmain.f90:2447-2492 InterpolateCls ; camb/modules.f90:952-1029 InterpolateClArr(Templated)
//   K4  camb/lensing.f90:94-518 CorrFuncFullSky / CorrFuncFullSkyImpl
//       source/Calculator_CAMB.f90:349-463 CAMBCalc_SetPowersFromCAMB
//
// B200 design for K4: the theta grid and the multipole grid of the lensing integrals do not depend on the
// parameter point, so every Legendre / Wigner-d recurrence of the reference is evaluated ONCE at handle
// creation (lens_tables_kernel) and the per-point work becomes
//   (1) sigma^2(theta), C_gl,2(theta) = [B x l] x [l x theta]           FP64 tensor-pipe GEMM
//   (2) the non-perturbative correlation sums over ~343 sampled l       element-wise kernel
//   (3) lensed-minus-unlensed C_l = [B x theta] x [theta x l] (x4)      FP64 tensor-pipe GEMM
// instead of 164 x 3300-long sequential recurrences per point.
#pragma once
#include "common.cuh"

namespace cb200 {

// ------------------------------------------------------------------------------------------------ K3
struct InterpParams {
  int np, nl, max_l, LS, nspec, templated;
  const double* icl;     // [np][6][PROJ_LP]
  const int* ls;         // [nl]
  const int* llo_of_l;   // [max_l+1] 1-based lower sample index used by the reference's running search
  const double* tmpl;    // [4][8001] TT,EE,TE,PP or null
  double* cl;            // [np][nspec][LS]
};

__global__ void __launch_bounds__(192) interp_cls_kernel(InterpParams p, int lp_stride_icl) {
  __shared__ double y[6][128], dd[6][128];
  __shared__ int sl[128];
  const int lp = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < p.nl; i += blockDim.x) sl[i] = p.ls[i];
  __syncthreads();
  if (warp < p.nspec) {
    const int X = warp;
    const bool tm = p.templated && X < 3;
    for (int i = lane; i < p.nl; i += 32) {
      double v = p.icl[((size_t)lp * 6 + X) * lp_stride_icl + i];
      if (tm) v -= p.tmpl[(size_t)X * 8001 + sl[i]];
      y[X][i] = v;
    }
    __syncwarp();
    if (lane == 0) {
      // natural cubic spline in l (camb/subroutines.f90:253-296)
      const int n = p.nl;
      double* d2 = dd[X];
      double u[128];
      double d1r = (y[X][1] - y[X][0]) / (double)(sl[1] - sl[0]), d1l;
      d2[0] = 0; u[0] = 0;
      for (int i = 1; i <= n - 2; i++) {
        d1l = d1r;
        d1r = (y[X][i + 1] - y[X][i]) / (double)(sl[i + 1] - sl[i]);
        double xxdiv = 1. / (double)(sl[i + 1] - sl[i - 1]);
        double sig = (double)(sl[i] - sl[i - 1]) * xxdiv;
        double xp = 1. / (sig * d2[i - 1] + 2.);
        d2[i] = (sig - 1.) * xp;
        u[i] = (6. * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
      }
      d2[n - 1] = 0;
      for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
    }
  }
  __syncthreads();
  const int tot = p.nspec * (p.max_l + 1);
  for (int idx = tid; idx < tot; idx += blockDim.x) {
    const int X = idx / (p.max_l + 1), l = idx % (p.max_l + 1);
    double val = 0;
    if (l >= 2) {
      const int llo = p.llo_of_l[l], lhi = llo + 1;
      const double ho = sl[lhi - 1] - sl[llo - 1];
      const double a0 = (sl[lhi - 1] - l) / ho, b0 = (l - sl[llo - 1]) / ho;
      val = a0 * y[X][llo - 1] + b0 * y[X][lhi - 1] +
            ((a0 * a0 * a0 - a0) * dd[X][llo - 1] + (b0 * b0 * b0 - b0) * dd[X][lhi - 1]) * ho * ho / 6;
      if (p.templated && X < 3) val += p.tmpl[(size_t)X * 8001 + l];
    }
    p.cl[((size_t)lp * p.nspec + X) * p.LS + l] = val;
  }
}

// ------------------------------------------------------------------------------------------------ K4 tables
struct LensGeom {
  int max_l, lmax, lmax_lensed, npoints, NTH, NTHP, jmax, interp_fac, NLL;
  double dtheta;
};

// one thread per theta sample: every recurrence of lensing.f90:256-367, written once into
//   A1  [lmax-1][2*NTHP]   (1-d_11) | d_-11                        (B operand of GEMM 1, k = l-2)
//   M   [4][NTHP][NLL]     P sin, d22 sin/2, d2-2 sin/2, d20 sin   (B operands of GEMM 3, n = l-2)
//   tab [jmax][12][NTHP]   P,dm11,d11,d2m2,d22,d20,d13,d04,d4m4,d3m3,d1m3,d2m4 at the sampled l
__global__ void lens_tables_kernel(LensGeom g, const int* __restrict__ jidx_of_l, double* __restrict__ A1,
                                   double* __restrict__ M, double* __restrict__ tab) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;  // theta index 0..NTH-1  (i = it+1)
  if (it >= g.NTH) return;
  const double theta = (it + 1) * g.dtheta;
  const double x = cos(theta), sinth = sin(theta), halfsinth = sinth / 2;
  double pmm = 1, pmmp1 = x;
  const double fac1 = (1 - x), fac2 = (1 + x), fac = fac1 / fac2;
  for (int l = 2; l <= g.lmax; l++) {
    const double P = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) / l;
    const double dP = l * (pmmp1 - x * P) / (sinth * sinth);
    pmm = pmmp1;
    pmmp1 = P;
    const double llp1 = (double)(l * (l + 1));
    const double lf2 = (double)((l + 2) * (l - 1));
    const double lroot = sqrt(llp1 * lf2);
    const double d11 = fac1 * dP / llp1 + P;
    const double dm11 = fac2 * dP / llp1 - P;
    A1[(size_t)(l - 2) * 2 * g.NTHP + it] = 1 - d11;
    A1[(size_t)(l - 2) * 2 * g.NTHP + g.NTHP + it] = dm11;
    const double d22 = (((4 * x - 8) / fac2 + llp1) * P + 4 * fac * (fac2 + (x - 2) / llp1) * dP) / lf2;
    const double theta_cut = 0.244949 / sqrt(3. * llp1 - 8.);
    double d2m2;
    if (theta > theta_cut)
      d2m2 = ((llp1 - (4 * x + 8) / fac1) * P + 4 / fac * (-fac1 + (x + 2) / llp1) * dP) / lf2;
    else
      d2m2 = llp1 * lf2 * theta * theta * theta * theta * (1. / 384. - (3. * llp1 - 8.) / 23040. * theta * theta);
    const double d20 = (2 * x * dP - llp1 * P) / lroot;
    if (l <= g.lmax_lensed) {
      const size_t o = (size_t)it * g.NLL + (l - 2);
      const size_t ms = (size_t)g.NTHP * g.NLL;
      M[o] = P * sinth;
      M[ms + o] = d22 * halfsinth;
      M[2 * ms + o] = d2m2 * halfsinth;
      M[3 * ms + o] = d20 * sinth;
    }
    const int j = jidx_of_l[l];
    if (j >= 0) {
      const double rootfac1 = sqrt((double)(l + 2)) * sqrt((double)(l - 1));
      const double rootfac2 = sqrt((double)(l + 3)) * sqrt((double)(l - 2));
      const double d1m2 = sinth / rootfac1 * (dP - 2 / fac1 * dm11);
      const double d12 = sinth / rootfac1 * (dP - 2 / fac2 * d11);
      double d1m3 = 0, d2m3 = 0, d3m3 = 0, d13 = 0, d23 = 0;
      if (l >= 3) {
        const double sinfac = 4 / sinth;
        d1m3 = (-(x + 0.5) * d1m2 * sinfac - lf2 * dm11 / rootfac1) / rootfac2;
        d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) / rootfac2;
        d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) / rootfac2;
        d13 = ((x - 0.5) * d12 * sinfac - lf2 * d11 / rootfac1) / rootfac2;
        d23 = (-fac1 * d22 * sinfac + rootfac1 * d12) / rootfac2;
      }
      (void)d23;
      double d04 = 0, d2m4 = 0, d4m4 = 0;
      if (l >= 4) {
        const double rootfac3 = sqrt((double)(l - 3)) * sqrt((double)(l + 4));
        d04 = ((-llp1 + (18 * x * x + 6) / (sinth * sinth)) * d20 - 6 * x * lf2 * dP / lroot) / (rootfac2 * rootfac3);
        d2m4 = (-(6 * x + 4) * d2m3 / sinth - rootfac2 * d2m2) / rootfac3;
        d4m4 = (-7 / 5. * (llp1 - 6) * d2m2 + 12 / 5. * (-llp1 + (9 * x + 26) / fac1) * d3m3) / (llp1 - 12);
      }
      double* t = tab + (size_t)j * 12 * g.NTHP + it;
      t[0 * g.NTHP] = P;    t[1 * g.NTHP] = dm11; t[2 * g.NTHP] = d11;  t[3 * g.NTHP] = d2m2;
      t[4 * g.NTHP] = d22;  t[5 * g.NTHP] = d20;  t[6 * g.NTHP] = d13;  t[7 * g.NTHP] = d04;
      t[8 * g.NTHP] = d4m4; t[9 * g.NTHP] = d3m3; t[10 * g.NTHP] = d1m3; t[11 * g.NTHP] = d2m4;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K4 (0)
// C_l -> (2l+1)/(4pi) weighted inputs with the template tail (lensing.f90:211-241).
// cin [np][4][LL]: 0 Cphil3, 1 CTT, 2 CEE, 3 CTE for l = 0..lmax
__global__ void lens_prep_kernel(int np, LensGeom g, int LS, int LL, const double* __restrict__ cl /*[np][6][LS]*/,
                                 const double* __restrict__ tmpl /*[4][8001]*/, double* __restrict__ cin) {
  const int lp = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np || l > g.lmax) return;
  const double* c = cl + (size_t)lp * 6 * LS;
  double o[4] = {0, 0, 0, 0};
  if (l >= 2) {
    const double sc = (2 * l + 1) / (4 * kPi) * 2 * kPi / (l * (l + 1));
    if (l <= g.max_l) {
      o[0] = c[3 * (size_t)LS + l] * (2 * l + 1) * (l + 1) / ((double)l * (double)l * (double)l) / (4 * kPi);
      o[1] = c[0 * (size_t)LS + l] * sc;
      o[2] = c[1 * (size_t)LS + l] * sc;
      o[3] = c[2 * (size_t)LS + l] * sc;
    } else {
      const int L = g.max_l;
      const double scL = (2 * L + 1) / (4 * kPi) * 2 * kPi / (L * (L + 1));
      const double cttL = c[0 * (size_t)LS + L] * scL;
      const double cphL = c[3 * (size_t)LS + L] * (2 * L + 1) * (L + 1) / ((double)L * (double)L * (double)L) / (4 * kPi);
      const double f2 = cttL / (scL * tmpl[0 * 8001 + L]);
      const double f = cphL / (scL * tmpl[3 * 8001 + L]);
      o[0] = tmpl[3 * 8001 + l] * f * sc;
      o[1] = tmpl[0 * 8001 + l] * f2 * sc;
      o[2] = tmpl[1 * 8001 + l] * f2 * sc;
      o[3] = tmpl[2 * 8001 + l] * f2 * sc;
    }
  }
  double* d = cin + (size_t)lp * 4 * LL;
#pragma unroll
  for (int k = 0; k < 4; k++) d[(size_t)k * LL + l] = o[k];
}

// ------------------------------------------------------------------------------------------------ K4 (2)
// correlation-function sums over the sampled multipoles (lensing.f90:312-436).  Thread = theta sample,
// PB points per CTA so each table element fetched from L2 is used PB times.
#ifndef CB200_LENS_PB
#define CB200_LENS_PB 2   // points per CTA of lens_corr_kernel (1 024 points per launch: 1 / 2 / 3 / 4 -> 2.23 / 1.85 / 2.35 / 2.04 us per point for the lensing phase)
#endif
constexpr int LENS_PB = CB200_LENS_PB;

struct LensCorrParams {
  int np, LL;
  LensGeom g;
  const double* sc;     // [np][2*NTHP] sigma^2 | Cg2
  const double* cin;    // [np][4][LL]
  const double* tab;    // [jmax][12][NTHP]
  const int* lj;        // [jmax] sampled multipoles
  const double* apod;   // [NTHP] apodisation weight (0 on padding)
  double* corr;         // [np][4][NTHP]
};

__global__ void __launch_bounds__(192) lens_corr_kernel(LensCorrParams p) {
  extern __shared__ double sm[];  // [LENS_PB][3][jmax] CTT,CEE,CTE at sampled l
  const LensGeom& g = p.g;
  const int it = threadIdx.x;
  const int lp0 = blockIdx.x * LENS_PB;
  for (int idx = threadIdx.x; idx < LENS_PB * 3 * g.jmax; idx += blockDim.x) {
    const int b = idx / (3 * g.jmax), r = idx % (3 * g.jmax), k = r / g.jmax, j = r % g.jmax;
    const int lp = lp0 + b;
    sm[idx] = (lp < p.np) ? p.cin[((size_t)lp * 4 + 1 + k) * p.LL + p.lj[j]] : 0.0;
  }
  __syncthreads();
  if (it >= g.NTHP) return;
  double sig[LENS_PB], cg2[LENS_PB], s1[LENS_PB][4], s2[LENS_PB][4];
#pragma unroll
  for (int b = 0; b < LENS_PB; b++) {
    const int lp = min(lp0 + b, p.np - 1);
    sig[b] = p.sc[(size_t)lp * 2 * g.NTHP + it];
    cg2[b] = p.sc[(size_t)lp * 2 * g.NTHP + g.NTHP + it];
#pragma unroll
    for (int k = 0; k < 4; k++) s1[b][k] = s2[b][k] = 0;
  }
  const bool live = it < g.NTH;
  for (int j = 0; j < g.jmax && live; j++) {
    const int l = p.lj[j];
    const double* t = p.tab + (size_t)j * 12 * g.NTHP + it;
    const double P = t[0], dm11 = t[g.NTHP], d11 = t[2 * g.NTHP], d2m2 = t[3 * g.NTHP], d22 = t[4 * g.NTHP],
                 d20 = t[5 * g.NTHP], d13 = t[6 * g.NTHP], d04 = t[7 * g.NTHP], d4m4 = t[8 * g.NTHP],
                 d3m3 = t[9 * g.NTHP], d1m3 = t[10 * g.NTHP], d2m4 = t[11 * g.NTHP];
    const double llp1 = (double)(l * (l + 1));
    const double lroot = sqrt(llp1 * (double)((l + 2) * (l - 1)));
    const double rootllp1 = sqrt((double)l) * sqrt((double)(l + 1));
    const double rootfac1 = sqrt((double)(l + 2)) * sqrt((double)(l - 1));
    const double rootfac2 = sqrt((double)(l + 3)) * sqrt((double)(l - 2));
    const double rootfac3 = (l >= 4) ? sqrt((double)(l - 3)) * sqrt((double)(l + 4)) : 0.0;
#pragma unroll
    for (int b = 0; b < LENS_PB; b++) {
      const double sigmasq = sig[b], Cg2 = cg2[b];
      const double X000 = exp(-llp1 * sigmasq / 4);
      const double X022 = X000 * (1 + sigmasq);
      const double X220 = lroot / 4 * X000;
      const double X121 = -0.5 * rootfac1 * X000;
      const double X132 = -0.5 * rootfac2 * X000;
      const double X242 = 0.25 * rootfac2 * rootfac3 * X022;
      const double dX000 = -llp1 / 4 * X000;
      const double dX022 = (1 - llp1 / 4) * X022;
      const double f1 = dX000 * dX000, f3 = X220 * X220, Cg2sq = Cg2 * Cg2;
      const double ctt = sm[(b * 3 + 0) * g.jmax + j], cee = sm[(b * 3 + 1) * g.jmax + j],
                   cte = sm[(b * 3 + 2) * g.jmax + j];
      double fac = ((X000 * X000 - 1) + Cg2sq * f1) * P + Cg2sq * f3 * d2m2 + 8 / llp1 * f1 * Cg2 * dm11;
      const double c0 = ctt * fac;
      const double f2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
      fac = 2 * Cg2 * X121 * X132 * d13 + f2 * d22 + Cg2sq * X242 * X220 * d04;
      const double c1 = cee * fac;
      fac = (f3 * P + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) + f2 * d2m2;
      const double c2 = cee * fac;
      fac = (X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
            Cg2sq * (X220 / 2 * d2m4 * X242 + (f3 / 2 + dX022 * dX000) * d20);
      const double c3 = cte * fac;
      if (j < 14) { s1[b][0] += c0; s1[b][1] += c1; s1[b][2] += c2; s1[b][3] += c3; }
      else { s2[b][0] += c0; s2[b][1] += c1; s2[b][2] += c2; s2[b][3] += c3; }
    }
  }
  const double ap = p.apod[it];
#pragma unroll
  for (int b = 0; b < LENS_PB; b++) {
    const int lp = lp0 + b;
    if (lp >= p.np) break;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double c = live ? (s1[b][k] + g.interp_fac * s2[b][k]) * ap : 0.0;
      p.corr[((size_t)lp * 4 + k) * g.NTHP + it] = c;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K4 (4)
// lensed spectra (lensing.f90:499-510) + CosmoMC units and tails (Calculator_CAMB.f90:349-463)
struct FinishParams {
  int np, LS, NLL, lmax_out, lmax_computed_cl, lmax_lensed, n_highl, lmax_tensor, have_tensor, LST;
  double dtheta;
  const double* cl;        // [np][6][LS] unlensed
  const double* lcon;      // [np][4][NLL]  GEMM-3 results (TT, 22, 2-2, TE)
  const double* cl_tensor; // [np or 1][4][LST] or null
  int tensor_shared;       // 1: one tensor set per batch (point index 0)
  const double* highl;     // [4][n_highl] lensed template TT,EE,BB,TE
  const double* aphiphi;   // [np] or null
  double* cl_lensed;       // [np][4][LS] dimensionless
  double* cls_out;         // [np][5][lmax_out+1] TT,TE,EE,BB,PP
  // reference SAVE semantics (`real(mcp) :: highL_norm = 0`, Calculator_CAMB.f90:358,398-399): the lensed-only TT of the
  // FIRST point ever evaluated fixes the tail normalisation of every later point and call.  Device scalar, written once
  // by highl_norm_kernel ahead of lens_finish_kernel; null: every point normalises its own tail.
  double* norm_dev;
};

__device__ __forceinline__ double lensed_tt_units(const FinishParams& p, int lp, int ll) {
  const double cons = (2.7255 * 1e6) * (2.7255 * 1e6);
  const double fac = ll * (ll + 1) / kTwoPi * p.dtheta * 2 * kPi;
  return cons * (p.lcon[(size_t)lp * 4 * p.NLL + (ll - 2)] * fac + p.cl[(size_t)lp * 6 * p.LS + ll]);
}

__global__ void highl_norm_kernel(FinishParams p) {
  if (threadIdx.x || blockIdx.x || *p.norm_dev != 0) return;
  const int lmx = min(p.lmax_computed_cl, p.lmax_out);
  if (lmx < p.n_highl) *p.norm_dev = lensed_tt_units(p, 0, lmx) / p.highl[lmx];
}

__global__ void lens_finish_kernel(FinishParams p) {
  const int lp = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= p.np) return;
  const double cons = (2.7255 * 1e6) * (2.7255 * 1e6);
  const double* c = p.cl + (size_t)lp * 6 * p.LS;
  const int lmx = min(p.lmax_computed_cl, p.lmax_out);
  auto lensed = [&](int ll, double o[4]) {
    const double fac = ll * (ll + 1) / kTwoPi * p.dtheta * 2 * kPi;
    const double* lc = p.lcon + (size_t)lp * 4 * p.NLL + (ll - 2);
    const double t = lc[0], a = lc[p.NLL], b = lc[2 * (size_t)p.NLL], x = lc[3 * (size_t)p.NLL];
    o[0] = t * fac + c[ll];
    o[1] = (a + b) * fac + c[(size_t)p.LS + ll];
    o[2] = (a - b) * fac;
    o[3] = x * fac + c[2 * (size_t)p.LS + ll];
  };
  if (l <= p.lmax_lensed && l < p.LS) {
    double o[4] = {0, 0, 0, 0};
    if (l >= 2) lensed(l, o);
#pragma unroll
    for (int k = 0; k < 4; k++) p.cl_lensed[((size_t)lp * 4 + k) * p.LS + l] = o[k];
  }
  if (l > p.lmax_out || !p.cls_out) return;
  double* out = p.cls_out + (size_t)lp * 5 * (p.lmax_out + 1);
  double v[5] = {0, 0, 0, 0, 0};  // TT,TE,EE,BB,PP
  if (l >= 2) {
    if (l <= lmx) {
      double o[4];
      lensed(l, o);
      v[0] = cons * o[0]; v[1] = cons * o[3]; v[2] = cons * o[1]; v[3] = cons * o[2];
      if (p.have_tensor && l <= min(lmx, p.lmax_tensor)) {
        const double* ct = p.cl_tensor + (size_t)(p.tensor_shared ? 0 : lp) * 4 * p.LST;
        v[0] += cons * ct[l]; v[2] += cons * ct[(size_t)p.LST + l]; v[3] += cons * ct[2 * (size_t)p.LST + l];
        v[1] += cons * ct[3 * (size_t)p.LST + l];
      }
      // real(l+1)**2/l**2 is a default (single precision) REAL expression in the reference
      const float ratio = ((float)(l + 1) * (float)(l + 1)) / (float)(l * l);
      v[4] = c[3 * (size_t)p.LS + l] * (double)ratio / kTwoPi * (p.aphiphi ? p.aphiphi[lp] : 1.0);
    } else if (l < p.n_highl) {
      double norm = p.norm_dev ? *p.norm_dev : 0.0;
      if (!p.norm_dev) {
        double o[4];
        lensed(lmx, o);
        norm = cons * o[0] / p.highl[lmx];
      }
      v[0] = norm * p.highl[l];
      v[2] = norm * p.highl[(size_t)p.n_highl + l];
      v[3] = norm * p.highl[2 * (size_t)p.n_highl + l];
      v[1] = norm * p.highl[3 * (size_t)p.n_highl + l];
    }
  }
#pragma unroll
  for (int k = 0; k < 5; k++) out[(size_t)k * (p.lmax_out + 1) + l] = v[k];
}

// rms deflection angle (Calculator_CAMB.f90:440-449) and the reject flags (:239-256); one warp per point
__global__ void derived_status_kernel(int np, int LS, int lmax_out, int max_l, const double* __restrict__ cl,
                                      const double* __restrict__ cls_out, double* __restrict__ derived /*[np][4]*/,
                                      int* __restrict__ status, const double* __restrict__ initpower /*tensors: [np][10]*/) {
  const int lp = blockIdx.x, lane = threadIdx.x;
  if (lp >= np) return;
  double rms = 0;
  if (max_l >= 2000) {
    const double* cp = cl + ((size_t)lp * 6 + 3) * LS;
    for (int L = 2 + lane; L <= 2000; L += 32) {
      const float ratio = ((float)(L + 1) * (float)(L + 1)) / (float)(L * L);
      rms += cp[L] * (double)ratio / kTwoPi * (L + 0.5) / (L * (L + 1));
    }
    rms = warp_sum(rms);
  }
  int bad = 0;
  if (cls_out) {
    const double* o = cls_out + (size_t)lp * 5 * (lmax_out + 1);
    for (int l = 2 + lane; l <= lmax_out; l += 32) {
      const double tt = o[l], te = o[(size_t)(lmax_out + 1) + l], ee = o[2 * (size_t)(lmax_out + 1) + l],
                   bb = o[3 * (size_t)(lmax_out + 1) + l];
      if (tt < 0 || ee < 0 || bb < 0) bad = 1;
      if (isnan(tt) || isnan(te) || isnan(ee) || isnan(bb)) bad = 1;
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    if (derived) {
      derived[(size_t)lp * 4 + 0] = sqrt(rms) * 180 / kPi * 60;
      double r02 = 0, rbb = 0, at = 0;
      if (initpower) {  // tensor-derived ratios, Calculator_CAMB.f90:451-455
        const double* ip = initpower + (size_t)lp * 10;
        r02 = tensor_power_dev(ip, 0.002) / scalar_power_dev(ip, 0.002);
        at = tensor_power_dev(ip, ip[8]);
        rbb = tensor_power_dev(ip, 0.01) / scalar_power_dev(ip, 0.01);
      }
      derived[(size_t)lp * 4 + 1] = r02; derived[(size_t)lp * 4 + 2] = rbb; derived[(size_t)lp * 4 + 3] = at;
    }
    if (status) status[lp] = bad ? 1 : 0;
  }
}

}  // namespace cb200
