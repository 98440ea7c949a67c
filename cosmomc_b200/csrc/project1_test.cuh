// TEST KERNEL (built only with -DCB200_TEST_KERNELS into libcosmob200_test.so; NOT part of libcosmob200.so).
// K1 generation 1 - CTA = (point, 8 wavenumbers), lanes = multipoles, Bessel table nodes gathered straight from L2.
// Superseded by project4.cuh (628 -> 122 us/point); kept as an independent decomposition of the same reference
// routines (camb/cmbmain.f90:1295-1374, 1387-1420, 1440-1562) that the parity tests cross-check the product kernel
// against.
#pragma once
#include "common.cuh"
#include "project.cuh"

namespace cb200 {

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

}  // namespace cb200
