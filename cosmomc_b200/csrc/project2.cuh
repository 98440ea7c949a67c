// K1 v2 — line-of-sight projection with a sliding window of the Bessel table staged in shared memory.
//
// Same reference behaviour as project.cuh (camb/cmbmain.f90:478-498, 1295-1374, 1387-1420, 1440-1562 and the
// partial k-contraction of :2132-2264); only the parallel decomposition differs.
//
// Why: the v1 kernel gathers two 16-byte table nodes per (q, l, tau) triple straight from L2 and is bound by
// L2 latency (ncu, profiles/r01_project_v1.txt: long-scoreboard stalls, 16% issue utilisation).  The table
// argument x = q (tau0 - tau) moves by <= ~1.9 table rows between consecutive wavenumbers, so for a block of
// QC consecutive wavenumbers the rows touched during a few time steps form a contiguous window of ~100-190
// rows.  This kernel keeps that window in a shared-memory ring (cp.async fills, 512 B per row = 32 multipoles
// x {j_l, j_l''}) and slides it along the time axis: every table row is fetched from L2 once per CTA pass
// instead of once per triple, and the inner loop reads conflict-free LDS.128 instead of L2.
//
// CTA = (parameter point, block of QC = NW*QW wavenumbers, chunk of 32 sampled multipoles); warp = QW
// wavenumbers; lane = multipole.  A slab is S consecutive time samples: QW*S = 32 (q, tau) pairs per warp,
// whose interpolation weights and k-interpolated, dtau-weighted sources are computed one pair per lane.
#pragma once
#include "common.cuh"
#include "project.cuh"

namespace cb200 {

constexpr int W2_NW = 8;                 // warps per CTA
constexpr int W2_QW = 8;                 // wavenumbers per warp
constexpr int W2_S = 4;                  // time samples per slab
constexpr int W2_QC = W2_NW * W2_QW;     // wavenumbers per CTA
constexpr int W2_R = 184;                // ring capacity in table rows
static_assert(W2_QW * W2_S == 32, "one (q, tau) pair per lane");

struct __align__(16) ProjMeta2 {
  double a, fac;
  double s0, s1;
  double s2;
  int off0, off1;  // ring mode: element offsets (double2 units) of rows i0, i0+1 in the ring; direct mode: i0, -
};

struct ProjQ2 {
  double q, w, a0, b0, a03h, b03h, ho2o6;
  int klo, steps, valid, pad;
};

struct Proj2Params {
  PointView v;
  int p0, nl, num_xx, NQB2, tensors;
  double max_eta_k;
  const double* ddsrc;
  const double* bx;
  const double2* bes3;      // [3][num_xx][32]
  const double* initpower;
  double* part;             // [chunk][NQB2][6][PROJ_LP]
  double* delta;            // optional [chunk][NQ][PROJ_LP][3]
  unsigned long long* triples;
  unsigned long long* ring_stats;  // optional [4]: slabs, direct slabs, rows loaded, pairs
  LinSegs bseg;
  int ls[PROJ_LP];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double interp_source2(const double* __restrict__ S, const double* __restrict__ D,
                                                 const ProjQ2& c) {
  return c.a0 * S[c.klo - 1] + c.b0 * S[c.klo] + (c.a03h * D[c.klo - 1] + c.b03h * D[c.klo]) * c.ho2o6;
}

__global__ void __launch_bounds__(32 * W2_NW, 2) project2_kernel(const Proj2Params p) {
  constexpr int NW = W2_NW, QW = W2_QW, S = W2_S, QC = W2_QC, R = W2_R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* ring = reinterpret_cast<double2*>(smem_raw);                                  // [R][32]
  ProjMeta2* meta = reinterpret_cast<ProjMeta2*>(smem_raw + sizeof(double2) * R * 32);   // [NW][32]
  ProjQ2* qc = reinterpret_cast<ProjQ2*>(smem_raw + sizeof(double2) * R * 32 + sizeof(ProjMeta2) * NW * 32);
  __shared__ int s_wmin[2][NW], s_wmax[2][NW], s_wn1[NW][QW], s_wn2[NW][QW];
  __shared__ int s_nlo, s_nhi, s_rlo, s_rhi;

  const PointView& v = p.v;
  const int lp = blockIdx.z, pt = p.p0 + lp, qb = blockIdx.x, chunk = blockIdx.y;
  const int nq = v.n_q[pt];
  const int q0 = qb * QC;
  if (q0 >= nq) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = chunk * 32 + lane;
  const bool lvalid = j < p.nl;
  const int l = lvalid ? p.ls[j] : 0;
  if (chunk * 32 >= p.nl) return;

  const int nt = v.n_tau[pt], nk = v.n_k[pt];
  const double tau0 = v.thermo[(size_t)pt * 5];
  const double* tau = v.tau + (size_t)pt * v.NT;
  const double* dtau = v.dtau + (size_t)pt * v.NT;
  const double* ksrc = v.ksrc + (size_t)pt * v.NK;
  const LinSegs& tseg = v.tseg[pt];
  const size_t row_stride = (size_t)v.NK;
  const size_t tau_stride = (size_t)v.NSRC * v.NK;
  const double* src = v.src + (size_t)pt * v.NT * tau_stride;
  const double* dds = p.ddsrc + (size_t)lp * v.NT * tau_stride;
  const double* ip = p.initpower + (size_t)lp * 10;
  const double2* bes = p.bes3 + (size_t)chunk * p.num_xx * 32;

  // ---- per-wavenumber constants ----
  if (tid < QC) {
    ProjQ2 c;
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
  if (tid == 0) { s_nlo = 0x7fffffff; s_nhi = 0; s_rlo = 1; s_rhi = 0; }
  __syncthreads();

  // ---- integration window per (wavenumber, multipole) ----
  const ProjQ2* wq = qc + warp * QW;
  int n1[QW], n2[QW];
  unsigned reached = 0, doint = 0;
#pragma unroll
  for (int qq = 0; qq < QW; qq++) {
    n1[qq] = 0x7fffffff; n2[qq] = 0;
    const ProjQ2& c = wq[qq];
    if (c.valid && lvalid) {
      const double qv = c.q;
      int llmax = (int)llround(__dmul_rn(qv, tau0));
      if (llmax < 15) llmax = 17;
      else llmax = (int)llround(__dmul_rn(qv, __dadd_rn(tau0, __ddiv_rn(6 * kPi, qv))));
      if (l <= llmax) {
        double xlim = 0.05 * l;
        xlim = fmax(xlim, 35.0);
        xlim = l - xlim;
        const double tau2 = tau[1];
        double tmin = __dsub_rn(tau0, __ddiv_rn((double)(80 * l), qv));
        tmin = fmax(tau2, tmin);
        double tmax = __dsub_rn(tau0, __ddiv_rn(xlim, qv));
        tmax = fmin(tau0, tmax);
        if (!(tmax < tau2)) {
          reached |= 1u << qq;
          bool di = true;
          if (!p.tensors) {
            double qmax_int = __ddiv_rn((double)(max(850, l) * 3), tau0);
            qmax_int = __dmul_rn(qmax_int, (double)1.2f);
            di = qv < qmax_int;
          }
          if (di) {
            doint |= 1u << qq;
            n1[qq] = lin_index_of(tseg, tmin);
            n2[qq] = min(c.steps, lin_index_of(tseg, tmax));
          }
        }
      }
    }
    // union over the chunk's multipoles: which (q, tau) pairs this warp has to visit at all
    int a = n1[qq], b = n2[qq];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) {
      s_wn1[warp][qq] = a; s_wn2[warp][qq] = b;
      if (a <= b) { atomicMin(&s_nlo, a); atomicMax(&s_nhi, b); }
    }
  }
  double acc[QW][3];
#pragma unroll
  for (int qq = 0; qq < QW; qq++) acc[qq][0] = acc[qq][1] = acc[qq][2] = 0.0;
  unsigned long long my_triples = 0, st_slabs = 0, st_direct = 0, st_rows = 0, st_pairs = 0;

  __syncthreads();
  const int n_lo = s_nlo, n_hi = s_nhi;

  // ---- sweep over conformal time in slabs of S samples ----
  ProjMeta2* wmeta = meta + warp * 32;
  int par = 0;
  for (int n_base = n_lo; n_base <= n_hi; n_base += S, par ^= 1) {
    // (1) one (q, tau) pair per lane
    int i0 = -1;
    {
      const int qq = lane / S, n = n_base + (lane % S);
      const ProjQ2& c = wq[qq];
      ProjMeta2 m;
      m.a = 0; m.fac = 0; m.s0 = m.s1 = m.s2 = 0; m.off0 = -1; m.off1 = 0;
      if (c.valid && n <= c.steps && n >= s_wn1[warp][qq] && n <= s_wn2[warp][qq]) {
        const double t = tau[n - 1];
        const double x = fabs(__dmul_rn(c.q, __dsub_rn(tau0, t)));
        int bi = lin_index_of(p.bseg, x);
        bi = min(bi, p.num_xx - 1);
        const double x1 = p.bx[bi], x0 = p.bx[bi - 1];
        double fac = __dsub_rn(x1, x0);
        const double a = __ddiv_rn(__dsub_rn(x1, x), fac);
        fac = __ddiv_rn(__dmul_rn(__dmul_rn(fac, fac), a), 6.0);
        const double dt = dtau[n - 1];
        const double* Sp = src + (size_t)(n - 1) * tau_stride;
        const double* Dp = dds + (size_t)(n - 1) * tau_stride;
        m.a = a; m.fac = fac;
        if (n >= 2) {  // Source_q(1,:) is forced to zero (IntegrationVars_Init, cmbmain.f90:1380)
          m.s0 = interp_source2(Sp, Dp, c) * dt;
          m.s1 = interp_source2(Sp + row_stride, Dp + row_stride, c) * dt;
          m.s2 = interp_source2(Sp + 2 * row_stride, Dp + 2 * row_stride, c) * dt;
        }
        i0 = bi - 1;
        m.off0 = i0;
        m.off1 = i0 % R;
      }
      wmeta[lane] = m;
    }
    int rmin = (i0 >= 0) ? i0 : 0x7fffffff, rmax = (i0 >= 0) ? i0 + 1 : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
      rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if (lane == 0) { s_wmin[par][warp] = rmin; s_wmax[par][warp] = rmax; }
    __syncthreads();  // previous slab fully consumed by every warp; window bounds visible
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int w = 0; w < NW; w++) { lo = min(lo, s_wmin[par][w]); hi = max(hi, s_wmax[par][w]); }
    if (hi < 0) continue;  // nothing active in this slab (CTA-uniform)
    const bool direct = (hi - lo + 1) > R;
    if (!direct) {
      // (2) slide the ring: fetch the rows of [lo, hi] that are not resident
      int rlo = s_rlo, rhi = s_rhi;
      int la, lb, la2 = 1, lb2 = 0;
      if (rlo > rhi || hi < rlo - 1 || lo > rhi + 1) { la = lo; lb = hi; rlo = lo; rhi = hi; }
      else {
        la = lo; lb = min(hi, rlo - 1);       // below the resident range
        la2 = max(lo, rhi + 1); lb2 = hi;     // above it
        const int nlo = min(lo, rlo), nhi = max(hi, rhi);
        if (lo < rlo) { rlo = nlo; rhi = min(nhi, nlo + R - 1); }
        else { rhi = nhi; rlo = max(nlo, nhi - R + 1); }
      }
      for (int e = tid; e < (lb - la + 1) * 32; e += 32 * NW) {
        const int row = la + (e >> 5), ll = e & 31;
        cp_async16(&ring[(row % R) * 32 + ll], &bes[(size_t)row * 32 + ll]);
      }
      for (int e = tid; e < (lb2 - la2 + 1) * 32; e += 32 * NW) {
        const int row = la2 + (e >> 5), ll = e & 31;
        cp_async16(&ring[(row % R) * 32 + ll], &bes[(size_t)row * 32 + ll]);
      }
      if (p.ring_stats && tid == 0) st_rows += max(0, lb - la + 1) + max(0, lb2 - la2 + 1);
      cp_async_wait_all();
      __syncthreads();
      if (tid == 0) { s_rlo = rlo; s_rhi = rhi; }
    }
    if (p.ring_stats && tid == 0) { st_slabs++; st_direct += direct ? 1 : 0; }

    // (3) accumulate: 32 (q, tau) pairs per warp, lanes = multipoles
#pragma unroll
    for (int pr = 0; pr < 32; pr++) {
      const int qq = pr / S, n = n_base + (pr % S);
      const double2* m2 = reinterpret_cast<const double2*>(&wmeta[pr]);
      const double2 s2i = m2[2];
      const int r0 = __double2loint(s2i.y);
      if (r0 < 0) continue;  // warp-uniform
      const bool act = (n >= n1[qq]) && (n <= n2[qq]);
      const double2 af = m2[0], s01 = m2[1];
      double2 nd0, nd1;
      if (!direct) {
        const int sl0 = __double2hiint(s2i.y);
        const int sl1 = (sl0 + 1 == R) ? 0 : sl0 + 1;
        nd0 = ring[sl0 * 32 + lane];
        nd1 = ring[sl1 * 32 + lane];
      } else {
        nd0 = __ldg(&bes[(size_t)r0 * 32 + lane]);
        nd1 = __ldg(&bes[(size_t)(r0 + 1) * 32 + lane]);
      }
      const double a2 = af.x;
      const double J = a2 * nd0.x + (1 - a2) * (nd1.x - ((a2 + 1) * nd0.y + (2 - a2) * nd1.y) * af.y);
      if (act) {
        acc[qq][0] += s01.x * J;
        acc[qq][1] += s01.y * J;
        acc[qq][2] += s2i.x * J;
        if (p.triples) my_triples++;
      }
      if (p.ring_stats && lane == 0) st_pairs++;
    }
  }

  if (p.triples) {
    unsigned long long t = my_triples;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(p.triples, t);
  }
  if (p.ring_stats) {
    if (tid == 0) { atomicAdd(p.ring_stats + 0, st_slabs); atomicAdd(p.ring_stats + 1, st_direct); atomicAdd(p.ring_stats + 2, st_rows); }
    if (lane == 0) atomicAdd(p.ring_stats + 3, st_pairs);
  }

  // ---- Limber value of the lensing source and the partial k-contraction over this warp's wavenumbers ----
  double cl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int qq = 0; qq < QW; qq++) {
    const ProjQ2& c = wq[qq];
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
            sa = interp_source2(src + (size_t)(n - 1) * tau_stride + 2 * row_stride,
                                dds + (size_t)(n - 1) * tau_stride + 2 * row_stride, c);
          if (n + 1 >= 2 && n + 1 <= c.steps)
            sb = interp_source2(src + (size_t)n * tau_stride + 2 * row_stride,
                                dds + (size_t)n * tau_stride + 2 * row_stride, c);
          s3 = (sa * (1 - xf) + xf * sb) * sqrt(kPi / 2 / ((double)l + 0.5)) / c.q;
        }
        acc[qq][2] = s3;
      }
    }
    const double d0 = acc[qq][0], d1 = acc[qq][1], d2 = acc[qq][2];
    if (p.delta) {
      double* dp = p.delta + (((size_t)lp * v.NQ + (q0 + warp * QW + qq)) * PROJ_LP + j) * 3;
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
  // cross-warp sum in a fixed order through shared memory (the ring is free now)
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem_raw);  // [NW][6][32]
#pragma unroll
  for (int X = 0; X < 6; X++) red[(warp * 6 + X) * 32 + lane] = cl[X];
  __syncthreads();
  if (warp == 0) {
    double* pp = p.part + (((size_t)lp * p.NQB2 + qb) * 6) * PROJ_LP + j;
#pragma unroll
    for (int X = 0; X < 6; X++) {
      double s = 0;
      for (int w = 0; w < NW; w++) s += red[(w * 6 + X) * 32 + lane];
      pp[(size_t)X * PROJ_LP] = s;
    }
  }
}

constexpr size_t W2_SMEM = sizeof(double2) * W2_R * 32 + sizeof(ProjMeta2) * W2_NW * 32 + sizeof(ProjQ2) * W2_QC;

// re-layout of the node table for the windowed kernel: [row][PROJ_LP] -> [chunk][row][32]
__global__ void bessel_relayout_kernel(int num_xx, const double2* __restrict__ bes, double2* __restrict__ bes3) {
  const int i = blockIdx.x, t = threadIdx.x;  // t < PROJ_LP
  if (i >= num_xx || t >= PROJ_LP) return;
  bes3[((size_t)(t >> 5) * num_xx + i) * 32 + (t & 31)] = bes[(size_t)i * PROJ_LP + t];
}

}  // namespace cb200
