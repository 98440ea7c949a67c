// ORACLE (test infrastructure, NOT product code) — CosmoMC side of the hot path:
// SetPowersFromCAMB unit conversion and the likelihoods.  See orc_core.hpp header.
#pragma once
#include "orc_core.hpp"
#include "orc_cmb.hpp"

namespace orc {

static const double COBE_CMBTemp = 2.7255;  // camb/constants.f90 ; Calculator_CAMB.f90:356

// ---- source/Calculator_CAMB.f90:349-463 CAMBCalc_SetPowersFromCAMB ---------------------------------
// Output Cls in CosmoMC order TT, TE, EE, BB, PP each [lmax_out+1] (D_l in muK^2; PP=[L(L+1)]^2 C/2pi).
// Cl_lensed[4][..] = TT,EE,BB,TE from CorrFuncFullSky; Cl_scalar_phi = l^4 C_phi (C_Phi of CAMB).
// cl_lmax[5] = max l wanted per spectrum (0 = not wanted); lmax_computed_cl = CosmoSettings%lmax_computed_cl.
// highL_lensed[4][..] lensed template (TT,EE,BB,TE as in data/HighL_lensedCls.dat), used above lmax_computed_cl.
// highL_norm: in/out; the reference keeps it in a SAVEd local (set on the first call only, :358,396-397);
// pass *highL_norm = 0 to recompute.
inline void SetPowersFromCAMB(const double* const Cl_lensed[4], const double* Cl_scalar_phi,
                              const double* const Cl_tensor[4], int lmax_tensor, bool compute_tensors,
                              int lmax_computed_cl, const int cl_lmax[5], const double* const highL_lensed[4],
                              double Aphiphi, double* highL_norm, double* const out[5], int lmax_out,
                              double* rms_deflect) {
  const double cons = (COBE_CMBTemp * 1e6) * (COBE_CMBTemp * 1e6);
  // (i,j) loop order of the reference: (1,1)=TT, (2,2)=EE,(2,1)=TE, (3,3)=BB ... map to lensed index
  struct M { int out, idxT; };
  const M order[4] = {{0, 0}, {2, 1}, {1, 3}, {3, 2}};  // TT ; EE ; TE ; BB  (idxT: 0 TT,1 EE,2 BB,3 TE)
  for (int m = 0; m < 4; m++) {
    int lmaxCL = cl_lmax[order[m].out];
    int lmx = std::min(lmax_computed_cl, lmaxCL);
    if (lmx == 0) continue;
    double* CL = out[order[m].out];
    int t = order[m].idxT;
    for (int l = 0; l <= lmax_out; l++) CL[l] = 0;
    for (int l = 2; l <= lmx; l++) CL[l] = cons * Cl_lensed[t][l];
    if (lmax_computed_cl < lmaxCL) {
      if (*highL_norm == 0) *highL_norm = CL[lmx] / highL_lensed[t][lmx];
      for (int l = lmx + 1; l <= lmaxCL; l++) CL[l] = *highL_norm * highL_lensed[t][l];
    }
    if (compute_tensors) {
      int lt = std::min(lmx, lmax_tensor);
      for (int l = 2; l <= lt; l++) CL[l] = CL[l] + cons * Cl_tensor[t][l];
    }
  }
  {
    int lmx = std::min(lmax_computed_cl, cl_lmax[4]);
    if (lmx != 0) {
      double* CL = out[4];
      for (int l = 0; l <= lmax_out; l++) CL[l] = 0;
      for (int l = 2; l <= lmx; l++) {
        // real(l+1)**2/l**2 is single precision in the reference (default REAL); restated as such
        float ratio = ((float)(l + 1) * (float)(l + 1)) / (float)(l * l);
        CL[l] = Cl_scalar_phi[l] * (double)ratio / twopi * Aphiphi;
      }
    }
  }
  if (rms_deflect) {
    double rms = 0;
    for (int L = 2; L <= 2000; L++) {
      float ratio = ((float)(L + 1) * (float)(L + 1)) / (float)(L * L);
      rms = rms + Cl_scalar_phi[L] * (double)ratio / twopi * (L + 0.5) / (L * (L + 1));
    }
    *rms_deflect = std::sqrt(rms) * 180 / pi * 60;
  }
}

// ---- source/Matrix_utils_new.f90:2033-2047 Matrix_QuadForm (vecT * (Mat * vec), symmetric Mat) -----
inline double Matrix_QuadForm(const double* Mat, const double* vec, int n) {
  double tot = 0;
  for (int i = 0; i < n; i++) {
    double s = 0;
    for (int j = 0; j < n; j++) s += Mat[(size_t)i * n + j] * vec[j];
    tot += vec[i] * s;
  }
  return tot;
}

// ---- source/CMB.f90:305-329 TPlikLiteLikelihood_LogLike ----------------------------------------------
// cls[3][..] = D_l TT, TE, EE indexed by l.  blmin/blmax are absolute l (already + plmin), weights[l]
// already multiplied by 2pi/(l(l+1)) (CMB.f90:225-232).  used bins: nb[3] bins per spectrum, first nb[i]
// of the common bin table.
inline double PlikLite_LogLike(const double* const cls[3], const int nb[3], const int* blmin, const int* blmax,
                               const double* weights, const double* invcov, const double* X_data, double calPlanck) {
  int nused = nb[0] + nb[1] + nb[2];
  std::vector<double> cl(nused), d(nused);
  int ix = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < nb[i]; j++) {
      double s = 0;
      for (int l = blmin[j]; l <= blmax[j]; l++) s += cls[i][l] * weights[l];
      cl[ix++] = s;
    }
  for (int i = 0; i < nused; i++) cl[i] = cl[i] / (calPlanck * calPlanck);
  for (int i = 0; i < nused; i++) d[i] = X_data[i] - cl[i];
  return Matrix_QuadForm(invcov, d.data(), nused) / 2;
}

// ---- symmetric eigen-decomposition used for Matrix_Diagonalize (DSYEV, Matrix_utils_new.f90:361-383) -
// LAPACK is a third-party dependency of the reference (unpinned version); restated here by
// cyclic Jacobi.  Eigenvalues ascending, eigenvectors in columns of U (row-major U[i*n+j]).
inline void sym_eigen(std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& U) {
  U.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) U[(size_t)i * n + i] = 1;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    double diag = 0;
    for (int i = 0; i < n; i++) diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
    if (off <= 1e-32 * diag || off == 0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[(size_t)p * n + q];
        if (apq == 0) continue;
        double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        double tau = (aqq - app) / (2 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
        double c = 1 / std::sqrt(1 + t * t), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double ukp = U[(size_t)k * n + p], ukq = U[(size_t)k * n + q];
          U[(size_t)k * n + p] = c * ukp - s * ukq;
          U[(size_t)k * n + q] = s * ukp + c * ukq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[(size_t)i * n + i];
  // sort ascending
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return w[a] < w[b]; });
  std::vector<double> w2(n), U2((size_t)n * n);
  for (int j = 0; j < n; j++) {
    w2[j] = w[idx[j]];
    for (int i = 0; i < n; i++) U2[(size_t)i * n + j] = U[(size_t)i * n + idx[j]];
  }
  w = w2; U = U2;
}

// ---- source/CMBlikes.f90:861-914 CMBLikes_Transform (no COffset) ------------------------------------
// C (n x n, symmetric, row-major) is replaced by C_f^{1/2} U g(D) U^T C_f^{1/2}.
inline void CMBLikes_Transform(std::vector<double>& C, const double* Chat, const double* CfHalf, int n) {
  std::vector<double> U, Diag, A = C;
  sym_eigen(A, n, Diag, U);
  // Rot = U^T Chat U
  std::vector<double> T((size_t)n * n), Rot((size_t)n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += U[(size_t)k * n + i] * Chat[(size_t)k * n + j];
      T[(size_t)i * n + j] = s;
    }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += T[(size_t)i * n + k] * U[(size_t)k * n + j];
      Rot[(size_t)i * n + j] = s;
    }
  std::vector<double> roots(n);
  for (int i = 0; i < n; i++) roots[i] = std::sqrt(Diag[i]);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) Rot[(size_t)i * n + j] /= roots[i];
    for (int j = 0; j < n; j++) Rot[(size_t)j * n + i] /= roots[i];
  }
  // Rot = U Rot U^T
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += Rot[(size_t)i * n + k] * U[(size_t)j * n + k];
      T[(size_t)i * n + j] = s;
    }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += U[(size_t)i * n + k] * T[(size_t)k * n + j];
      Rot[(size_t)i * n + j] = s;
    }
  std::vector<double> D2, V;
  sym_eigen(Rot, n, D2, V);
  for (int i = 0; i < n; i++) {
    double v = std::sqrt(2 * std::max(0.0, D2[i] - std::log(D2[i]) - 1));
    D2[i] = (D2[i] - 1 >= 0) ? std::fabs(v) : -std::fabs(v);  // sign(a,b)
  }
  // U = CfHalf * V ; C = (U * diag) * U^T
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += CfHalf[(size_t)i * n + k] * V[(size_t)k * n + j];
      T[(size_t)i * n + j] = s;
    }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += T[(size_t)i * n + k] * D2[k] * T[(size_t)j * n + k];
      C[(size_t)i * n + j] = s;
    }
}

// ---- generic binned CMBLikes (source/CMBlikes.f90:1165-1256 with 981-995,1295-1325) in dense form ------
// The .dataset parsing (ReadIni, CMBlikes.f90:371-859) is host data-prep; this takes its result:
//  map_cl[ncl_req][nL]   theory map cross spectra (after AdaptTheoryForMaps), l = lmin..lmax
//  W[nbins][ncl][nL]     bin windows mapped onto output (used) cl index (zero where no window)
//  Wc, fid_corr          optional correction windows [nbins][ncl][ncl_req][nL] folded: see api
// To keep the oracle literal but compact, the api passes binned theory pieces; see orc_like_api.inc.
struct CMBLikesData {
  int nmaps = 0, ncl = 0, nbins = 0, ncl_used = 0, like_approx = 2;  // 1 = HL, 2 = fid gaussian
  std::vector<int> cl_use_index;                                   // [ncl_used] 0-based into vecp
  std::vector<double> NoiseM, ChatM, sqrt_fiducial;                // [nbins][nmaps*nmaps] (NoiseM may be empty)
  std::vector<double> inv_covariance;                              // [(nbins*ncl_used)^2]
};

// binnedC: [nbins][ncl] binned theory in "elements" order (i>=j lower triangle rows, CMBlikes.f90:916-929)
inline double CMBLikes_chisq(const CMBLikesData& D, const double* binnedC) {
  int n = D.nmaps;
  std::vector<double> bigX((size_t)D.nbins * D.ncl_used);
  for (int bin = 0; bin < D.nbins; bin++) {
    std::vector<double> C((size_t)n * n), vecp(D.ncl);
    int ix = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        C[(size_t)i * n + j] = binnedC[(size_t)bin * D.ncl + ix];
        C[(size_t)j * n + i] = C[(size_t)i * n + j];
        ix++;
      }
    if (!D.NoiseM.empty())
      for (int k = 0; k < n * n; k++) C[k] += D.NoiseM[(size_t)bin * n * n + k];
    if (D.like_approx == 1) {
      CMBLikes_Transform(C, &D.ChatM[(size_t)bin * n * n], &D.sqrt_fiducial[(size_t)bin * n * n], n);
    } else {
      for (int k = 0; k < n * n; k++) C[k] -= D.ChatM[(size_t)bin * n * n + k];
    }
    ix = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) vecp[ix++] = C[(size_t)i * n + j];
    for (int k = 0; k < D.ncl_used; k++) bigX[(size_t)bin * D.ncl_used + k] = vecp[D.cl_use_index[k]];
  }
  return Matrix_QuadForm(D.inv_covariance.data(), bigX.data(), D.nbins * D.ncl_used);
}

}  // namespace orc
