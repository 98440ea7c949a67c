// Line-of-sight projection of the source functions against the spherical-Bessel table, fused with the
// P(k)-weighted k-contraction (kernels K0, K1, K2 of DESIGN.md).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   K0  camb/cmbmain.f90:1207-1218  InitSourceInterpolation  (natural spline of Src along k)
//   K1  camb/cmbmain.f90:478-498    SourceToTransfers
//       camb/cmbmain.f90:1295-1374  InterpolateSources
//       camb/cmbmain.f90:1387-1420  DoSourceIntegration (flat)
//       camb/cmbmain.f90:1440-1562  DoFlatIntegration incl. the Limber switch for the lensing source
//   K2  camb/cmbmain.f90:2132-2264  CalcScalCls, :2344-2397 CalcTensCls (flat)
//
// Parallel decomposition (B200): one CTA per (parameter point, block of Q consecutive wavenumbers).
// Lanes of a warp are consecutive sampled multipoles, so the Bessel node table (multipole fastest) is read
// 512 contiguous bytes per node; NS warp groups split the conformal-time axis and are reduced through
// shared memory; the transfer functions Delta_l(q) never leave registers: the CTA emits its partial
// sum over q of P(q) Delta_a Delta_b dlnq, and a small second kernel adds the partials in a fixed order.
#pragma once
#include "common.cuh"

namespace cb200 {

constexpr int PROJ_Q = 8;      // wavenumbers per CTA
constexpr int PROJ_NS = 4;     // warp groups splitting the time axis
constexpr int PROJ_SLAB = 32;  // time samples per group per pass
constexpr int PROJ_LW = 3;     // multipole warps (96 padded multipoles)
constexpr int PROJ_LP = 32 * PROJ_LW;

struct PointView {  // resident per-point inputs (device pointers), strides NT / NK / NQ
  int NT, NK, NQ, NSRC;
  const double* thermo;  // [P][5] tau0, taurst, taurend, reion_start, reion_complete
  const int* n_tau;
  const double* tau;     // [P][NT]
  const double* dtau;    // [P][NT]
  const LinSegs* tseg;   // [P]
  const int* n_k;
  const double* ksrc;    // [P][NK]
  const int* n_q;
  const double* q;       // [P][NQ]
  const double* dq;      // [P][NQ]
  const double* src;     // [P][NT][NSRC][NK]
};

// ------------------------------------------------------------------------------------------------ K0
// Spline set-up that depends on the k grid only (per point): xxdiv, sig, xp, gam of the tridiagonal sweep.
__global__ void spline_setup_kernel(PointView v, int p0, int np, double* __restrict__ coef /*[np][5][NK]*/) {
  int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np) return;
  const int pt = p0 + lp, nk = v.n_k[pt];
  const double* x = v.ksrc + (size_t)pt * v.NK;
  double* c = coef + (size_t)lp * 5 * v.NK;
  double gam = 0;
  c[3 * v.NK + 0] = 0;
  for (int i = 0; i + 1 < nk; i++) c[4 * v.NK + i] = 1. / (x[i + 1] - x[i]);  // used by the tiled kernel only
  for (int i = 1; i <= nk - 2; i++) {
    double xxdiv = 1. / (x[i + 1] - x[i - 1]);
    double sig = (x[i] - x[i - 1]) * xxdiv;
    double xp = 1. / (sig * gam + 2.);
    gam = (sig - 1.) * xp;
    c[0 * v.NK + i] = xxdiv;
    c[1 * v.NK + i] = sig;
    c[2 * v.NK + i] = xp;
    c[3 * v.NK + i] = gam;
  }
}

// one thread per (point, tau, source) row: forward sweep for u, backward sweep for the second derivatives
__global__ void source_spline_kernel(PointView v, int p0, int np, const double* __restrict__ coef,
                                     double* __restrict__ ddsrc /*[np][NT][NSRC][NK]*/) {
  const long long rows_per_pt = (long long)v.NT * v.NSRC;
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows_per_pt * np) return;
  const int lp = (int)(r / rows_per_pt);
  const int row = (int)(r % rows_per_pt);
  const int pt = p0 + lp;
  const int it = row / v.NSRC;
  if (it >= v.n_tau[pt]) return;
  const int nk = v.n_k[pt];
  const double* x = v.ksrc + (size_t)pt * v.NK;
  const double* y = v.src + ((size_t)pt * rows_per_pt + row) * v.NK;
  double* d2 = ddsrc + ((size_t)lp * rows_per_pt + row) * v.NK;
  const double* c = coef + (size_t)lp * 5 * v.NK;
  double d1r = (y[1] - y[0]) / (x[1] - x[0]), d1l, u = 0;
  d2[0] = 0;
  for (int i = 1; i <= nk - 2; i++) {
    d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    u = (6. * (d1r - d1l) * c[0 * v.NK + i] - c[1 * v.NK + i] * u) * c[2 * v.NK + i];
    d2[i] = u;
  }
  double nxt = 0;
  d2[nk - 1] = 0;
  for (int i = nk - 2; i >= 1; i--) {
    nxt = c[3 * v.NK + i] * nxt + d2[i];
    d2[i] = nxt;
  }
  d2[0] = 0;  // gam(0) = 0, u(0) = 0
}

// Tiled K0: the rows of one point are contiguous ([tau][source][k]), so a CTA stages SPL_ROWS consecutive rows in
// shared memory with coalesced 16-byte loads, lets one thread per row run the two sequential sweeps out of shared
// memory (row stride NK+1 words: conflict-free), and writes the second derivatives back coalesced.  The one-thread-
// per-row kernel above reads and writes global memory with a 1.8 KB stride between lanes (0.65 TB/s measured).
// 1/(x[i+1]-x[i]) comes from the set-up kernel (a multiply instead of the reference's division: values, not indices).
constexpr int SPL_ROWS = 32;
constexpr int SPL_THREADS = 128;
__global__ void __launch_bounds__(SPL_THREADS) source_spline_tiled_kernel(PointView v, int p0, int np,
                                                                           const double* __restrict__ coef,
                                                                           double* __restrict__ ddsrc) {
  extern __shared__ __align__(16) double spl_smem[];
  const int NK = v.NK, NKP = NK + 1;
  double* sc = spl_smem;                 // [5][NK] sweep coefficients + reciprocal spacings
  double* sy = spl_smem + 5 * NK;        // [SPL_ROWS][NKP]
  const int rows_per_pt = v.NT * v.NSRC;
  const int tiles_per_pt = (rows_per_pt + SPL_ROWS - 1) / SPL_ROWS;
  const int lp = blockIdx.x / tiles_per_pt, tile = blockIdx.x - lp * tiles_per_pt;
  if (lp >= np) return;
  const int pt = p0 + lp;
  const int row0 = tile * SPL_ROWS;
  const int nrow_pt = v.n_tau[pt] * v.NSRC;          // rows beyond the point's time samples are never read
  if (row0 >= nrow_pt) return;
  const int nrows = min(SPL_ROWS, nrow_pt - row0);
  const int nk = v.n_k[pt];
  const double* c = coef + (size_t)lp * 5 * NK;
  const double* y = v.src + ((size_t)pt * rows_per_pt + row0) * NK;
  double* d2 = ddsrc + ((size_t)lp * rows_per_pt + row0) * NK;
  const int tid = threadIdx.x;
  for (int e = tid; e < 5 * NK; e += SPL_THREADS) sc[e] = c[e];
  // NK is a multiple of 2: 16-byte loads of the contiguous block, scattered into the padded rows
  const double2* y2 = reinterpret_cast<const double2*>(y);
  const int n2 = nrows * NK / 2;
  for (int e = tid; e < n2; e += SPL_THREADS) {
    const double2 val = __ldg(y2 + e);
    const int r = (2 * e) / NK, k = 2 * e - r * NK;
    sy[r * NKP + k] = val.x;
    sy[r * NKP + k + 1] = val.y;
  }
  __syncthreads();
  if (tid < nrows) {
    double* yr = sy + tid * NKP;
    double y0 = yr[0], y1 = yr[1];
    double d1r = (y1 - y0) * sc[4 * NK + 0], d1l, u = 0;
    yr[0] = 0;
    for (int i = 1; i <= nk - 2; i++) {
      const double y2v = yr[i + 1];
      d1l = d1r;
      d1r = (y2v - y1) * sc[4 * NK + i];
      u = (6. * (d1r - d1l) * sc[0 * NK + i] - sc[1 * NK + i] * u) * sc[2 * NK + i];
      yr[i] = u;          // y[i] is no longer needed
      y1 = y2v;
    }
    double nxt = 0;
    yr[nk - 1] = 0;
    for (int i = nk - 2; i >= 1; i--) {
      nxt = sc[3 * NK + i] * nxt + yr[i];
      yr[i] = nxt;
    }
    for (int i = nk; i < NK; i++) yr[i] = 0;
  }
  __syncthreads();
  double2* d22 = reinterpret_cast<double2*>(d2);
  for (int e = tid; e < n2; e += SPL_THREADS) {
    const int r = (2 * e) / NK, k = 2 * e - r * NK;
    d22[e] = make_double2(sy[r * NKP + k], sy[r * NKP + k + 1]);
  }
}

// ------------------------------------------------------------------------------------------------ K1
struct ProjParams {
  PointView v;
  int p0;                 // first resident point of this chunk; blockIdx.y is the chunk-local point
  int nl, num_xx, NQB;
  int tensors;
  double max_eta_k;       // maximum_qeta (scalar: Max_eta_k, tensor: Max_eta_k_tensor)
  const double* ddsrc;    // [chunk][NT][NSRC][NK]
  const double* bx;       // [num_xx] Bessel abscissae
  const double2* bes;     // [num_xx][PROJ_LP]
  const double* initpower;  // [chunk][10]
  double* part;           // [chunk][NQB][6][PROJ_LP]
  double* delta;          // optional [chunk][NQ][PROJ_LP][3]
  unsigned long long* triples;  // optional work counter
  LinSegs bseg;
  int ls[PROJ_LP];
};

struct __align__(16) ProjMeta {  // per (wavenumber, time sample): interpolation weights + weighted sources
  double a, fac;
  double s0, s1;
  double s2;
  int i0, pad;
};

struct ProjQ {  // per wavenumber of the CTA
  double q, w, a0, b0, a03h, b03h, ho2o6;
  int klo, steps, valid, pad;
};

__device__ __forceinline__ double scalar_power_dev(const double* ip, double k) {
  // camb/power_tilt.f90:114-133
  double lnrat = log(k / ip[7]);
  return ip[0] * exp(lnrat * (ip[1] - 1 + lnrat * (ip[2] / 2 + ip[3] / 6 * lnrat)));
}
__device__ __forceinline__ double tensor_power_dev(const double* ip, double k) {
  // camb/power_tilt.f90:136-162 with the CosmoMC mapping source/Calculator_CAMB.f90:839-877
  double r = ip[4], ns = ip[1], nt, ntrun;
  if (ip[9] != 0) {
    nt = -r / 8 * (2 - ns - r / 8);
    ntrun = r / 8 * (r / 8 + ns - 1);
  } else { nt = ip[5]; ntrun = ip[6]; }
  double lnrat = log(k / ip[8]);
  double k_dep = exp(lnrat * (nt + ntrun / 2 * lnrat));
  if (ip[8] != ip[7]) return r * scalar_power_dev(ip, ip[8]) * k_dep;  // tensor_param_rpivot
  return r * ip[0] * k_dep;                                            // tensor_param_indeptilt
}

__device__ __forceinline__ double interp_source(const double* __restrict__ S, const double* __restrict__ D,
                                                const ProjQ& c) {
  return c.a0 * S[c.klo - 1] + c.b0 * S[c.klo] + (c.a03h * D[c.klo - 1] + c.b03h * D[c.klo]) * c.ho2o6;
}

template <int Q, int NS, int SLAB>
__global__ void __launch_bounds__(32 * PROJ_LW * NS, 1) project_kernel(const ProjParams p) {
  constexpr int NTHR = 32 * PROJ_LW * NS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ProjMeta* meta = reinterpret_cast<ProjMeta*>(smem_raw);                      // [NS][SLAB][Q]
  double* red = reinterpret_cast<double*>(smem_raw);                           // reused in the epilogue
  constexpr size_t META_BYTES = sizeof(ProjMeta) * NS * SLAB * Q;
  constexpr size_t RED_BYTES = sizeof(double) * (NS > 1 ? (NS - 1) : 1) * Q * 3 * PROJ_LP;
  constexpr size_t OFF_Q = (META_BYTES > RED_BYTES ? META_BYTES : RED_BYTES);
  ProjQ* qc = reinterpret_cast<ProjQ*>(smem_raw + OFF_Q);                      // [Q]
  __shared__ int s_nhi;

  const PointView& v = p.v;
  const int lp = blockIdx.y, pt = p.p0 + lp, qb = blockIdx.x;
  const int nq = v.n_q[pt];
  const int q0 = qb * Q;
  if (q0 >= nq) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lw = warp % PROJ_LW, grp = warp / PROJ_LW;
  const int j = lw * 32 + lane;
  const bool lvalid = j < p.nl;
  const int l = lvalid ? p.ls[j] : 0;

  const int nt = v.n_tau[pt], nk = v.n_k[pt];
  const double tau0 = v.thermo[(size_t)pt * 5];
  const double* tau = v.tau + (size_t)pt * v.NT;
  const double* dtau = v.dtau + (size_t)pt * v.NT;
  const double* ksrc = v.ksrc + (size_t)pt * v.NK;
  const LinSegs& tseg = v.tseg[pt];
  const size_t row_stride = (size_t)v.NK;                 // between sources
  const size_t tau_stride = (size_t)v.NSRC * v.NK;        // between time samples
  const double* src = v.src + (size_t)pt * v.NT * tau_stride;
  const double* dds = p.ddsrc + (size_t)lp * v.NT * tau_stride;
  const double* ip = p.initpower + (size_t)lp * 10;

  // ---- per-wavenumber constants (InterpolateSources set-up, cmbmain.f90:1307-1320) ----
  if (tid < Q) {
    ProjQ c;
    const int qi = q0 + tid;
    c.valid = qi < nq;
    c.pad = 0;
    if (c.valid) {
      const double qv = v.q[(size_t)pt * v.NQ + qi];
      const double dqv = v.dq[(size_t)pt * v.NQ + qi];
      c.q = qv;
      c.w = (p.tensors ? tensor_power_dev(ip, qv) : scalar_power_dev(ip, qv)) * (dqv / qv);
      int klo = 1;
      while ((qv > ksrc[klo]) && (klo < (nk - 1))) klo++;
      c.klo = klo;
      const double ho = ksrc[klo] - ksrc[klo - 1];
      c.a0 = (ksrc[klo] - qv) / ho;
      c.b0 = (qv - ksrc[klo - 1]) / ho;
      c.ho2o6 = ho * ho / 6;
      c.a03h = (c.a0 * c.a0 * c.a0 - c.a0);
      c.b03h = (c.b0 * c.b0 * c.b0 - c.b0);
      // last time sample with a non-zero interpolated source (cmbmain.f90:1325-1361)
      const double max_etak_tensor = p.max_eta_k / 10;
      int step = 2;
      for (int i = nt; i >= 2; i--) {
        double xf = __dmul_rn(qv, __dsub_rn(tau0, tau[i - 1]));
        bool ok = xf > 1.e-8;
        if (p.tensors) ok = ok && (__dmul_rn(qv, tau[i - 1]) < max_etak_tensor);
        if (ok) { step = i; break; }
      }
      c.steps = step;
    } else {
      c.q = 1; c.w = 0; c.klo = 1; c.a0 = c.b0 = c.a03h = c.b03h = c.ho2o6 = 0; c.steps = 0;
    }
    qc[tid] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int i = 0; i < Q; i++) m = max(m, qc[i].steps);
    s_nhi = m;
  }

  // ---- per (wavenumber, multipole) integration window (DoSourceIntegration/DoFlatIntegration) ----
  int n1[Q], n2[Q];
  unsigned reached = 0, doint = 0;
#pragma unroll
  for (int qq = 0; qq < Q; qq++) {
    n1[qq] = 0x7fffffff; n2[qq] = 0;
    const ProjQ& c = qc[qq];
    if (!c.valid || !lvalid) continue;
    const double qv = c.q;
    int llmax = (int)llround(__dmul_rn(qv, tau0));
    if (llmax < 15) llmax = 17;
    else llmax = (int)llround(__dmul_rn(qv, __dadd_rn(tau0, __ddiv_rn(6 * kPi, qv))));
    if (l > llmax) continue;
    double xlim = 0.05 * l;
    xlim = fmax(xlim, 35.0);
    xlim = l - xlim;
    const double tau2 = tau[1];
    double tmin = __dsub_rn(tau0, __ddiv_rn((double)(80 * l), qv));
    tmin = fmax(tau2, tmin);
    double tmax = __dsub_rn(tau0, __ddiv_rn(xlim, qv));
    tmax = fmin(tau0, tmax);
    if (tmax < tau2) continue;
    reached |= 1u << qq;
    bool di = true;
    if (!p.tensors) {
      double qmax_int = __ddiv_rn((double)(max(850, l) * 3), tau0);
      qmax_int = __dmul_rn(qmax_int, (double)1.2f);
      di = qv < qmax_int;
    }
    if (!di) continue;
    doint |= 1u << qq;
    n1[qq] = lin_index_of(tseg, tmin);
    n2[qq] = min(c.steps, lin_index_of(tseg, tmax));
  }

  double acc[Q][3];
#pragma unroll
  for (int qq = 0; qq < Q; qq++) acc[qq][0] = acc[qq][1] = acc[qq][2] = 0.0;
  unsigned long long my_triples = 0;

  __syncthreads();
  const int n_hi = s_nhi;

  // ---- sweep over conformal time ----
  // starts at n = 1: IndexOf(TimeSteps, tmin) can truncate to 1; Source_q(1,:) = 0 there (cmbmain.f90:1380)
  for (int n_base = 1; n_base <= n_hi; n_base += NS * SLAB) {
    __syncthreads();
    for (int idx = tid; idx < NS * SLAB * Q; idx += NTHR) {
      const int qq = idx % Q, r = idx / Q;
      const int n = n_base + r;
      const ProjQ& c = qc[qq];
      ProjMeta m;
      m.pad = 0;
      if (c.valid && n <= c.steps) {
        const double t = tau[n - 1];
        const double x = fabs(__dmul_rn(c.q, __dsub_rn(tau0, t)));
        int bi = lin_index_of(p.bseg, x);
        bi = min(bi, p.num_xx - 1);
        const double x1 = p.bx[bi], x0 = p.bx[bi - 1];
        double fac = __dsub_rn(x1, x0);
        const double a = __ddiv_rn(__dsub_rn(x1, x), fac);
        fac = __ddiv_rn(__dmul_rn(__dmul_rn(fac, fac), a), 6.0);
        const double dt = dtau[n - 1];
        const double* S = src + (size_t)(n - 1) * tau_stride;
        const double* D = dds + (size_t)(n - 1) * tau_stride;
        m.a = a; m.fac = fac; m.i0 = bi - 1;
        m.s0 = m.s1 = m.s2 = 0;
        if (n >= 2) {
          m.s0 = interp_source(S, D, c) * dt;
          m.s1 = interp_source(S + row_stride, D + row_stride, c) * dt;
          m.s2 = interp_source(S + 2 * row_stride, D + 2 * row_stride, c) * dt;
        }
      } else {
        m.a = 0; m.fac = 0; m.s0 = m.s1 = m.s2 = 0; m.i0 = 0;
      }
      meta[r * Q + qq] = m;
    }
    __syncthreads();
    const int n_first = n_base + grp * SLAB;
    const ProjMeta* mg = meta + (size_t)grp * SLAB * Q;
#pragma unroll 1
    for (int nn = 0; nn < SLAB; nn++) {
      const int n = n_first + nn;
      if (n > n_hi) break;
#pragma unroll
      for (int qq = 0; qq < Q; qq++) {
        const bool act = (n >= n1[qq]) && (n <= n2[qq]);
        if (!__any_sync(0xffffffffu, act)) continue;
        const double2* m2 = reinterpret_cast<const double2*>(&mg[nn * Q + qq]);
        const double2 af = m2[0], s01 = m2[1], s2i = m2[2];
        const int i0 = __double2loint(s2i.y);
        const double2* row = p.bes + (size_t)i0 * PROJ_LP + j;
        const double2 nd0 = __ldg(row), nd1 = __ldg(row + PROJ_LP);
        const double a2 = af.x;
        // cubic-spline evaluation of j_l between the two nodes (cmbmain.f90:1515-1516)
        double J = a2 * nd0.x + (1 - a2) * (nd1.x - ((a2 + 1) * nd0.y + (2 - a2) * nd1.y) * af.y);
        if (act) {
          acc[qq][0] += s01.x * J;
          acc[qq][1] += s01.y * J;
          acc[qq][2] += s2i.x * J;
          if (p.triples) my_triples++;
        }
      }
    }
  }

  // ---- reduce the time-axis groups ----
  __syncthreads();
  if (NS > 1) {
    if (grp > 0) {
      double* dst = red + (size_t)(grp - 1) * Q * 3 * PROJ_LP;
#pragma unroll
      for (int qq = 0; qq < Q; qq++)
#pragma unroll
        for (int s = 0; s < 3; s++) dst[(qq * 3 + s) * PROJ_LP + j] = acc[qq][s];
    }
    __syncthreads();
    if (grp == 0) {
      for (int g = 0; g < NS - 1; g++) {
        const double* srcp = red + (size_t)g * Q * 3 * PROJ_LP;
#pragma unroll
        for (int qq = 0; qq < Q; qq++)
#pragma unroll
          for (int s = 0; s < 3; s++) acc[qq][s] += srcp[(qq * 3 + s) * PROJ_LP + j];
      }
    }
  }
  if (p.triples) {
    unsigned long long t = my_triples;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(p.triples, t);
  }
  if (grp != 0) return;

  // ---- Limber value of the lensing source (cmbmain.f90:1546-1556) and the partial k-contraction ----
  double cl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int qq = 0; qq < Q; qq++) {
    const ProjQ& c = qc[qq];
    if (!c.valid) continue;
    if (!p.tensors && lvalid && ((reached >> qq) & 1u)) {
      const bool use_limber = l > 400;
      if (!((doint >> qq) & 1u) || use_limber) {
        double xf = __dsub_rn(tau0, __ddiv_rn((double)l + 0.5, c.q));
        double s3 = 0;
        if (xf < tseg.highest && xf > tau[0]) {
          const int n = lin_index_of(tseg, xf);
          xf = __ddiv_rn(__dsub_rn(xf, tau[n - 1]), __dsub_rn(tau[n], tau[n - 1]));
          double sa = 0, sb = 0;
          if (n >= 2 && n <= c.steps)
            sa = interp_source(src + (size_t)(n - 1) * tau_stride + 2 * row_stride,
                               dds + (size_t)(n - 1) * tau_stride + 2 * row_stride, c);
          if (n + 1 >= 2 && n + 1 <= c.steps)
            sb = interp_source(src + (size_t)n * tau_stride + 2 * row_stride,
                               dds + (size_t)n * tau_stride + 2 * row_stride, c);
          s3 = (sa * (1 - xf) + xf * sb) * sqrt(kPi / 2 / ((double)l + 0.5)) / c.q;
        }
        acc[qq][2] = s3;
      }
    }
    const double d0 = acc[qq][0], d1 = acc[qq][1], d2 = acc[qq][2];
    if (p.delta) {
      double* dp = p.delta + (((size_t)lp * v.NQ + (q0 + qq)) * PROJ_LP + j) * 3;
      dp[0] = d0; dp[1] = d1; dp[2] = d2;
    }
    const double w = c.w;
    if (p.tensors) {
      cl[0] += w * d0 * d0; cl[1] += w * d1 * d1; cl[2] += w * d2 * d2; cl[3] += w * d0 * d1;
    } else {
      cl[0] += w * d0 * d0; cl[1] += w * d1 * d1; cl[2] += w * d0 * d1;
      cl[3] += w * d2 * d2; cl[4] += w * d2 * d0; cl[5] += w * d2 * d1;
    }
  }
  double* pp = p.part + (((size_t)lp * p.NQB + qb) * 6) * PROJ_LP + j;
#pragma unroll
  for (int X = 0; X < 6; X++) pp[(size_t)X * PROJ_LP] = cl[X];
}

// ------------------------------------------------------------------------------------------------ K2
// fixed-order sum of the per-CTA partials + the l-dependent normalisations (cmbmain.f90:2222-2257, :2388-2393)
__global__ void contract_reduce_kernel(int np, int p0, const int* __restrict__ n_q, int Q, int NQB, int nl,
                                       const int* __restrict__ ls, int tensors, const double* __restrict__ alens,
                                       const double* __restrict__ part, double* __restrict__ icl /*[np][6][PROJ_LP]*/,
                                       int nq_shared /* > 0: every point uses this wavenumber count */) {
  int lp = blockIdx.y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np || t >= 6 * PROJ_LP) return;
  const int X = t / PROJ_LP, j = t % PROJ_LP;
  double s = 0;
  if (j < nl) {
    const int nb = ((nq_shared > 0 ? nq_shared : n_q[p0 + lp]) + Q - 1) / Q;
    const double* pp = part + (((size_t)lp * NQB) * 6 + X) * PROJ_LP + j;
    for (int b = 0; b < nb; b++) s += pp[(size_t)b * 6 * PROJ_LP];
    const double ell = ls[j];
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    if (!tensors) {
      const double dbletmp = (ell * (ell + 1)) / kTwoPi * kFourPi;
      const double AL = alens ? alens[lp] : 1.0;
      switch (X) {
        case 0: s = s * dbletmp; break;
        case 1: s = s * dbletmp * ctnorm; break;
        case 2: s = s * dbletmp * sqrt(ctnorm); break;
        case 3: s = AL * s * kFourPi * ell * ell * ell * ell; break;
        case 4: s = sqrt(AL) * s * kFourPi * ell * ell * ell; break;
        default: s = sqrt(AL) * s * kFourPi * ell * ell * ell * sqrt(ctnorm); break;
      }
    } else {
      double dbletmp = (ell * (ell + 1)) / kTwoPi * kPi / 4;
      switch (X) {
        case 0: s = s * dbletmp * ctnorm; break;
        case 1: case 2: s = (ls[j] == 1) ? 0.0 : s * dbletmp; break;
        case 3: s = s * dbletmp * sqrt(ctnorm); break;
        default: s = 0; break;
      }
    }
  }
  icl[((size_t)lp * 6 + X) * PROJ_LP + j] = s;
}

// ---- shared-transfer contraction (one source point, many initial-power points): operands of the DMMA GEMM ----
// D2[q][X][l] = Delta_a Delta_b of CalcScalCls / CalcTensCls (cmbmain.f90:2195-2215, :2377-2383)
__global__ void delta_products_kernel(int nq, int tensors, const double* __restrict__ delta /*[nq][PROJ_LP][3]*/,
                                      double* __restrict__ d2 /*[nq][6][PROJ_LP]*/) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * PROJ_LP) return;
  const int q = t / PROJ_LP, j = t % PROJ_LP;
  const double* d = delta + (size_t)t * 3;
  const double d0 = d[0], d1 = d[1], dd2 = d[2];
  double* o = d2 + (size_t)q * 6 * PROJ_LP + j;
  if (tensors) {
    o[0] = d0 * d0; o[PROJ_LP] = d1 * d1; o[2 * PROJ_LP] = dd2 * dd2; o[3 * PROJ_LP] = d0 * d1;
    o[4 * PROJ_LP] = 0; o[5 * PROJ_LP] = 0;
  } else {
    o[0] = d0 * d0; o[PROJ_LP] = d1 * d1; o[2 * PROJ_LP] = d0 * d1;
    o[3 * PROJ_LP] = dd2 * dd2; o[4 * PROJ_LP] = dd2 * d0; o[5 * PROJ_LP] = dd2 * d1;
  }
}
// W[pt][q] = P(q; initpower_pt) dq/q
__global__ void power_weights_kernel(int np, int nq, int tensors, const double* __restrict__ q,
                                     const double* __restrict__ dq, const double* __restrict__ initpower,
                                     double* __restrict__ W, int ldw) {
  const int lp = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np || i >= nq) return;
  const double* ip = initpower + (size_t)lp * 10;
  const double qv = q[i];
  W[(size_t)lp * ldw + i] = (tensors ? tensor_power_dev(ip, qv) : scalar_power_dev(ip, qv)) * (dq[i] / qv);
}

}  // namespace cb200
