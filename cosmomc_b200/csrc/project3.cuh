// K1 v3 — windowed line-of-sight projection, quarter-warp pair layout.
//
// Same reference behaviour as project.cuh / project2.cuh (camb/cmbmain.f90:478-498, 1295-1374, 1387-1420,
// 1440-1562, partial k-contraction of :2132-2264).  Same sliding shared-memory window of the Bessel table and
// the same software-pipelined (q, tau)-pair metadata as v2; what changes is the lane mapping of the inner loop.
//
// Why (ncu + in-kernel phase counters, profiles/r01_project_v2_*): v2 is bound by shared-memory wavefronts,
// not by FP64 or L2 — a warp-wide shared load costs one LSU wavefront per quarter-warp (8 lanes x 16 B) even
// when all 32 lanes read the same address, so broadcasting the 44 bytes of per-pair metadata to a whole warp
// costs 11 wavefronts on top of the 8 wavefronts of table nodes: 19 wavefronts per (pair, 32 multipoles).
// Here each QUARTER-warp works on a different (q, tau) pair — the four wavenumbers of the warp at one time
// sample — and every lane covers four multipoles (l-slots i, i+8, i+16, i+24 of the 32-multipole chunk).  The
// metadata loads now deliver four pairs' worth per instruction (11 wavefronts per FOUR pairs), the table loads
// are unchanged (8 per pair), i.e. 10.75 wavefronts per (pair, 32 multipoles) instead of 19, and table loads of
// octets whose multipoles are all outside the integration window are predicated off.
#pragma once
#include "common.cuh"
#include "project.cuh"

namespace cb200 {

#ifndef CB200_W3_USKIP
#define CB200_W3_USKIP 0
#endif
#ifndef CB200_W3_S
#define CB200_W3_S 8
#endif
#ifndef CB200_W3_NW
#define CB200_W3_NW 8
#endif
#ifndef CB200_W3_R
#define CB200_W3_R 176
#endif
#ifndef CB200_W3_MINB
#define CB200_W3_MINB 2
#endif
constexpr int W3_NW = CB200_W3_NW;   // warps per CTA
constexpr int W3_QW = 4;             // wavenumbers per warp = one per quarter-warp
constexpr int W3_S = CB200_W3_S;     // time samples per slab
constexpr int W3_QC = W3_NW * W3_QW; // wavenumbers per CTA
constexpr int W3_R = CB200_W3_R;     // ring capacity in table rows (+1 mirror row)
constexpr int W3_NP = W3_QW * W3_S;  // (q, tau) pairs per warp per slab, one per lane during metadata set-up
constexpr int W3_LK = 4;             // multipoles per lane
static_assert(W3_NP <= 32, "at most one (q, tau) pair per lane");

struct __align__(16) ProjMeta3 {
  double2 af[32];   // a = (x_{i+1}-x)/h ; fac = h^2 a / 6
  double2 s01[32];  // k-interpolated sources x dtau (temperature, E)
  double s2[32];    // lensing-potential source x dtau
  int off[32];      // byte offset of node row i0 in the ring
  int i0[32];       // table row (-1: pair not visited)
};

struct ProjQ3 {
  double q, w, a0, b0, a03h, b03h, ho2o6;
  int klo, steps, valid, pad;
};

struct Proj3Params {
  PointView v;
  int p0, nl, num_xx, NQB, tensors;
  double max_eta_k;
  const double* ddsrc;
  const double* bx;
  const double2* bes3;      // [3][num_xx][32]
  const double* initpower;
  double* part;             // [chunk][NQB][6][PROJ_LP]
  double* delta;            // optional [chunk][NQ][PROJ_LP][3]
  unsigned long long* triples;
  unsigned long long* ring_stats;  // optional [16]
  int qc_rt;                       // wavenumbers per block actually used (<= W3_QC; 0 = W3_QC): matches the caller's block size
  const unsigned char* need;       // optional [chunk][NQB]: run only the flagged (point, wavenumber block) pairs
  LinSegs bseg;
  int ls[PROJ_LP];
};

__device__ __forceinline__ void cp_async16_3(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}

template <bool COUNT>
__device__ __forceinline__ void project3_body(const Proj3Params& p, const int lp, const int qb, const int chunk) {
  constexpr int NW = W3_NW, QW = W3_QW, S = W3_S, QC = W3_QC, R = W3_R, NP = W3_NP, LK = W3_LK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* ring = reinterpret_cast<double2*>(smem_raw);                                        // [R+1][32]
  ProjMeta3* meta = reinterpret_cast<ProjMeta3*>(smem_raw + sizeof(double2) * (R + 1) * 32);   // [NW]
  ProjQ3* qc = reinterpret_cast<ProjQ3*>(smem_raw + sizeof(double2) * (R + 1) * 32 + sizeof(ProjMeta3) * NW);
  __shared__ int s_wmin[2][NW], s_wmax[2][NW], s_wn1[NW][QW], s_wn2[NW][QW];
  __shared__ int s_nlo, s_nhi, s_rlo, s_rhi;

  const PointView& v = p.v;
  const int pt = p.p0 + lp;
  const int nq = v.n_q[pt];
  const int qcr = p.qc_rt > 0 ? p.qc_rt : QC;
  const int q0 = qb * qcr;
  if (q0 >= nq || chunk * 32 >= p.nl) return;
  if (p.need && !p.need[(size_t)lp * p.NQB + qb]) return;  // fallback pass behind project4_kernel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qr = lane >> 3, li = lane & 7;  // quarter-warp = wavenumber slot ; octet position

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

  // ---- per-wavenumber constants (InterpolateSources set-up, cmbmain.f90:1307-1320) ----
  if (tid < QC) {
    ProjQ3 c;
    const int qi = q0 + tid;
    c.valid = (qi < nq) && (tid < qcr);
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

  // ---- integration windows of this lane's wavenumber (quarter) and its LK multipoles ----
  const ProjQ3* wq = qc + warp * QW;
  const ProjQ3& myq = wq[qr];
  int n1[LK], n2[LK], ll[LK];
  unsigned reached = 0, doint = 0;
  int un1 = 0x7fffffff, un2 = 0;
#pragma unroll
  for (int k = 0; k < LK; k++) {
    n1[k] = 0x7fffffff; n2[k] = 0;
    const int j = chunk * 32 + li + 8 * k;
    const bool lvalid = j < p.nl;
    const int l = lvalid ? p.ls[j] : 0;
    ll[k] = l;
    if (myq.valid && lvalid) {
      const double qv = myq.q;
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
          reached |= 1u << k;
          bool di = true;
          if (!p.tensors) {
            double qmax_int = __ddiv_rn((double)(max(850, l) * 3), tau0);
            qmax_int = __dmul_rn(qmax_int, (double)1.2f);
            di = qv < qmax_int;
          }
          if (di) {
            doint |= 1u << k;
            n1[k] = lin_index_of(tseg, tmin);
            n2[k] = min(myq.steps, lin_index_of(tseg, tmax));
          }
        }
      }
    }
    un1 = min(un1, n1[k]);
    un2 = max(un2, n2[k]);
  }
  // union over the quarter's 32 multipoles: which (q, tau) pairs have to be visited at all
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    un1 = min(un1, __shfl_xor_sync(0xffffffffu, un1, o));
    un2 = max(un2, __shfl_xor_sync(0xffffffffu, un2, o));
  }
  if (li == 0) {
    s_wn1[warp][qr] = un1; s_wn2[warp][qr] = un2;
    if (un1 <= un2) { atomicMin(&s_nlo, un1); atomicMax(&s_nhi, un2); }
  }
  double acc[LK][3];
#pragma unroll
  for (int k = 0; k < LK; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0.0;
  unsigned long long my_triples = 0, st_slabs = 0, st_direct = 0, st_rows = 0, st_pairs = 0;
  long long ck_pro = 0, ck_pre = 0, ck_bar = 0, ck_ring = 0, ck_cmp = 0, ck_fin = 0, ck_t0 = clock64(), ck_t;
#define CK3(var) do { if (COUNT) { ck_t = clock64(); var += ck_t - ck_t0; ck_t0 = ck_t; } } while (0)

  __syncthreads();
  const int n_lo = s_nlo, n_hi = s_nhi;

  // ---- sweep over conformal time in slabs of S samples, software-pipelined metadata (see project2.cuh) ----
  ProjMeta3& wm = meta[warp];
  const unsigned char* ring_bytes = reinterpret_cast<const unsigned char*>(ring) + li * 16;
  const int pq = lane % QW, pn = lane / QW;  // metadata set-up: lane = nn*QW + qq
  const ProjQ3& pc = wq[pq];
  const bool lane_has_pair = (lane < NP) && pc.valid;
  const int pw1 = lane_has_pair ? max(s_wn1[warp][pq], 1) : 0x7fffffff;
  const int pw2 = lane_has_pair ? min(s_wn2[warp][pq], pc.steps) : 0;

  double f_tau = 0, f_dtau = 0, f_s[3][4];
  bool f_valid = false;
  auto prefetch = [&](int nb) {
    const int n = nb + pn;
    f_valid = (n >= pw1) && (n <= pw2);
    if (f_valid) {
      f_tau = __ldg(tau + n - 1);
      f_dtau = __ldg(dtau + n - 1);
      const double* Sp = src + (size_t)(n - 1) * tau_stride + (pc.klo - 1);
      const double* Dp = dds + (size_t)(n - 1) * tau_stride + (pc.klo - 1);
#pragma unroll
      for (int sI = 0; sI < 3; sI++) {
        f_s[sI][0] = __ldg(Sp + sI * row_stride);
        f_s[sI][1] = __ldg(Sp + sI * row_stride + 1);
        f_s[sI][2] = __ldg(Dp + sI * row_stride);
        f_s[sI][3] = __ldg(Dp + sI * row_stride + 1);
      }
    }
  };
  unsigned vmask = 0;
  auto finish = [&](int nb, int par) {
    const int n = nb + pn;
    int i0 = -1;
    double ma = 0, mfac = 0, ms0 = 0, ms1 = 0, ms2 = 0;
    int moff = 0;
    if (f_valid) {
      const double x = fabs(__dmul_rn(pc.q, __dsub_rn(tau0, f_tau)));
      double x0, x1, inv_h;
      int bi = lin_locate(p.bseg, x, x0, x1, inv_h);
      if (bi > p.num_xx - 1) { bi = p.num_xx - 1; x0 = p.bx[bi - 1]; x1 = p.bx[bi]; inv_h = 1.0 / (x1 - x0); }
      // interpolation weights (values, not indices): reciprocal multiplies instead of the reference's divisions
      const double fac = x1 - x0;
      ma = (x1 - x) * inv_h;
      mfac = fac * fac * ma * (1.0 / 6.0);
      if (n >= 2) {  // Source_q(1,:) is forced to zero (IntegrationVars_Init, cmbmain.f90:1380)
        ms0 = (pc.a0 * f_s[0][0] + pc.b0 * f_s[0][1] + (pc.a03h * f_s[0][2] + pc.b03h * f_s[0][3]) * pc.ho2o6) * f_dtau;
        ms1 = (pc.a0 * f_s[1][0] + pc.b0 * f_s[1][1] + (pc.a03h * f_s[1][2] + pc.b03h * f_s[1][3]) * pc.ho2o6) * f_dtau;
        ms2 = (pc.a0 * f_s[2][0] + pc.b0 * f_s[2][1] + (pc.a03h * f_s[2][2] + pc.b03h * f_s[2][3]) * pc.ho2o6) * f_dtau;
      }
      i0 = bi - 1;
      moff = (i0 % R) * 512;
    }
    wm.af[lane] = make_double2(ma, mfac);
    wm.s01[lane] = make_double2(ms0, ms1);
    wm.s2[lane] = ms2;
    wm.off[lane] = moff;
    wm.i0[lane] = i0;
    vmask = __ballot_sync(0xffffffffu, i0 >= 0);
    int rmin = (i0 >= 0) ? i0 : 0x7fffffff, rmax = (i0 >= 0) ? i0 + 1 : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
      rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if (lane == 0) { s_wmin[par][warp] = rmin; s_wmax[par][warp] = rmax; }
  };

  int par = 0;
  if (n_lo <= n_hi) { prefetch(n_lo); finish(n_lo, 0); }
  CK3(ck_pro);
  for (int n_base = n_lo; n_base <= n_hi; n_base += S, par ^= 1) {
    const unsigned vm = vmask;
    if (n_base + S <= n_hi) prefetch(n_base + S);
    CK3(ck_pre);
    __syncthreads();  // previous slab fully consumed by every warp; this slab's window bounds visible
    CK3(ck_bar);
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int w = 0; w < NW; w++) { lo = min(lo, s_wmin[par][w]); hi = max(hi, s_wmax[par][w]); }
    const bool direct = (hi - lo + 1) > R;
    if (hi >= 0 && !direct) {
      int rlo = s_rlo, rhi = s_rhi;
      int la, lb, la2 = 1, lb2 = 0;
      if (rlo > rhi || hi < rlo - 1 || lo > rhi + 1) { la = lo; lb = hi; rlo = lo; rhi = hi; }
      else {
        la = lo; lb = min(hi, rlo - 1);
        la2 = max(lo, rhi + 1); lb2 = hi;
        const int nlo = min(lo, rlo), nhi = max(hi, rhi);
        if (lo < rlo) { rlo = nlo; rhi = min(nhi, nlo + R - 1); }
        else { rhi = nhi; rlo = max(nlo, nhi - R + 1); }
      }
      for (int e = tid; e < (lb - la + 1) * 32; e += 32 * NW) {
        const int row = la + (e >> 5), l32 = e & 31, slot = row % R;
        cp_async16_3(&ring[slot * 32 + l32], &bes[(size_t)row * 32 + l32]);
        if (slot == 0) cp_async16_3(&ring[R * 32 + l32], &bes[(size_t)row * 32 + l32]);
      }
      for (int e = tid; e < (lb2 - la2 + 1) * 32; e += 32 * NW) {
        const int row = la2 + (e >> 5), l32 = e & 31, slot = row % R;
        cp_async16_3(&ring[slot * 32 + l32], &bes[(size_t)row * 32 + l32]);
        if (slot == 0) cp_async16_3(&ring[R * 32 + l32], &bes[(size_t)row * 32 + l32]);
      }
      if (COUNT && p.ring_stats && tid == 0) st_rows += max(0, lb - la + 1) + max(0, lb2 - la2 + 1);
      asm volatile("cp.async.wait_all;\n" ::: "memory");
      __syncthreads();
      if (tid == 0) { s_rlo = rlo; s_rhi = rhi; }
    }
    if (COUNT && p.ring_stats && tid == 0 && hi >= 0) { st_slabs++; st_direct += direct ? 1 : 0; }
    CK3(ck_ring);

    // accumulate: one time sample per step; quarter-warp r works on pair (q_r, n), lane covers LK multipoles
#pragma unroll
    for (int nn = 0; nn < S; nn++) {
      if (((vm >> (nn * QW)) & ((1u << QW) - 1u)) == 0u) continue;  // warp-uniform
      const int n = n_base + nn;
      const int pr = nn * QW + qr;
      const double2 af = wm.af[pr];
      const double2 s01 = wm.s01[pr];
      const double s2 = wm.s2[pr];
      bool act[LK];
      double2 N0[LK], N1[LK];
#pragma unroll
      for (int k = 0; k < LK; k++) {
        act[k] = (n >= n1[k]) && (n <= n2[k]);
        N0[k] = make_double2(0.0, 0.0);
        N1[k] = make_double2(0.0, 0.0);
      }
      if (!direct) {
        const unsigned char* rp = ring_bytes + wm.off[pr];
#pragma unroll
        for (int k = 0; k < LK; k++)
          if (act[k]) {
            N0[k] = *reinterpret_cast<const double2*>(rp + k * 128);
            N1[k] = *reinterpret_cast<const double2*>(rp + k * 128 + 512);
          }
      } else {
        const int r0 = max(wm.i0[pr], 0);
#pragma unroll
        for (int k = 0; k < LK; k++)
          if (act[k]) {
            N0[k] = __ldg(&bes[(size_t)r0 * 32 + li + 8 * k]);
            N1[k] = __ldg(&bes[(size_t)(r0 + 1) * 32 + li + 8 * k]);
          }
      }
      // cubic-spline value of j_l between the two nodes (cmbmain.f90:1515-1516), weights expanded:
      //   J = (a j0 + b j1) + (g0 p0 + g1 p1),  b = 1-a, g0 = -b fac (a+1), g1 = -b fac (2-a)
      const double a2 = af.x, b2 = 1 - a2, t = -(b2 * af.y);
      const double g0 = t * (a2 + 1), g1 = t * (2 - a2);
#pragma unroll
      for (int k = 0; k < LK; k++) {
#if CB200_W3_USKIP
        if (!__any_sync(0xffffffffu, act[k])) continue;  // no lane of the warp integrates this octet at this step
#endif
        const double Jv = fma(g1, N1[k].y, fma(g0, N0[k].y, fma(b2, N1[k].x, a2 * N0[k].x)));  // 0 when !act
        acc[k][0] = fma(s01.x, Jv, acc[k][0]);
        acc[k][1] = fma(s01.y, Jv, acc[k][1]);
        acc[k][2] = fma(s2, Jv, acc[k][2]);
        if (COUNT && p.triples && act[k]) my_triples++;
      }
      if (COUNT && p.ring_stats && li == 0 && ((vm >> pr) & 1u)) st_pairs++;
    }
    CK3(ck_cmp);
    __syncwarp();
    if (n_base + S <= n_hi) finish(n_base + S, par ^ 1);
    CK3(ck_fin);
  }

  if (COUNT && p.triples) {
    unsigned long long t = my_triples;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(p.triples, t);
  }
  if (COUNT && p.ring_stats) {
    if (tid == 0) { atomicAdd(p.ring_stats + 0, st_slabs); atomicAdd(p.ring_stats + 1, st_direct); atomicAdd(p.ring_stats + 2, st_rows); }
    if (li == 0) atomicAdd(p.ring_stats + 3, st_pairs);
    if (lane == 0) {
      atomicAdd(p.ring_stats + 4, (unsigned long long)ck_pro); atomicAdd(p.ring_stats + 5, (unsigned long long)ck_pre);
      atomicAdd(p.ring_stats + 6, (unsigned long long)ck_bar); atomicAdd(p.ring_stats + 7, (unsigned long long)ck_ring);
      atomicAdd(p.ring_stats + 8, (unsigned long long)ck_cmp); atomicAdd(p.ring_stats + 9, (unsigned long long)ck_fin);
    }
  }

  // ---- Limber value of the lensing source (cmbmain.f90:1546-1556) and the partial k-contraction ----
  double cl[LK][6];
#pragma unroll
  for (int k = 0; k < LK; k++) {
#pragma unroll
    for (int X = 0; X < 6; X++) cl[k][X] = 0.0;
    const int l = ll[k];
    const int j = chunk * 32 + li + 8 * k;
    if (!myq.valid) continue;
    if (!p.tensors && j < p.nl && ((reached >> k) & 1u)) {
      const bool use_limber = l > 400;
      if (!((doint >> k) & 1u) || use_limber) {
        double xf = __dsub_rn(tau0, __ddiv_rn((double)l + 0.5, myq.q));
        double s3 = 0;
        if (xf < tseg.highest && xf > tau[0]) {
          const int n = lin_index_of(tseg, xf);
          xf = __ddiv_rn(__dsub_rn(xf, tau[n - 1]), __dsub_rn(tau[n], tau[n - 1]));
          double sa = 0, sb = 0;
          const double* S2p = src + 2 * row_stride + (myq.klo - 1);
          const double* D2p = dds + 2 * row_stride + (myq.klo - 1);
          if (n >= 2 && n <= myq.steps) {
            const double* a = S2p + (size_t)(n - 1) * tau_stride;
            const double* d = D2p + (size_t)(n - 1) * tau_stride;
            sa = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
          }
          if (n + 1 >= 2 && n + 1 <= myq.steps) {
            const double* a = S2p + (size_t)n * tau_stride;
            const double* d = D2p + (size_t)n * tau_stride;
            sb = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
          }
          s3 = (sa * (1 - xf) + xf * sb) * sqrt(kPi / 2 / ((double)l + 0.5)) / myq.q;
        }
        acc[k][2] = s3;
      }
    }
    const double d0 = acc[k][0], d1 = acc[k][1], d2 = acc[k][2];
    if (p.delta) {
      double* dp = p.delta + (((size_t)lp * v.NQ + (q0 + warp * QW + qr)) * PROJ_LP + j) * 3;
      dp[0] = d0; dp[1] = d1; dp[2] = d2;
    }
    const double w = myq.w;
    if (p.tensors) {
      cl[k][0] = w * d0 * d0; cl[k][1] = w * d1 * d1; cl[k][2] = w * d2 * d2; cl[k][3] = w * d0 * d1;
    } else {
      cl[k][0] = w * d0 * d0; cl[k][1] = w * d1 * d1; cl[k][2] = w * d0 * d1;
      cl[k][3] = w * d2 * d2; cl[k][4] = w * d2 * d0; cl[k][5] = w * d2 * d1;
    }
  }
  // sum over the warp's four wavenumbers (quarters), then over warps in a fixed order through shared memory
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem_raw);  // [NW][6][32]
#pragma unroll
  for (int k = 0; k < LK; k++)
#pragma unroll
    for (int X = 0; X < 6; X++) {
      double s = cl[k][X];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (qr == 0) red[(warp * 6 + X) * 32 + li + 8 * k] = s;
    }
  __syncthreads();
  if (warp == 0) {
    double* pp = p.part + (((size_t)lp * p.NQB + qb) * 6) * PROJ_LP + chunk * 32 + lane;
#pragma unroll
    for (int X = 0; X < 6; X++) {
      double s = 0;
      for (int w = 0; w < NW; w++) s += red[(w * 6 + X) * 32 + lane];
      pp[(size_t)X * PROJ_LP] = s;
    }
  }
#undef CK3
}

template <bool COUNT>
__global__ void __launch_bounds__(32 * W3_NW, CB200_W3_MINB) project3_kernel(const Proj3Params p) {
  project3_body<COUNT>(p, blockIdx.z, blockIdx.x, blockIdx.y);
}

// The pass behind project4_kernel: ONE CTA per (point, multipole chunk) walks the point's wavenumber blocks and runs the
// flagged ones (in practice the first block only).  Launching project3_kernel over every block and letting the unflagged
// CTAs exit cost 0.75 ms per 512 points for 186 000 empty CTAs - twice the work of the flagged ones.
template <bool COUNT>
__global__ void __launch_bounds__(32 * W3_NW, CB200_W3_MINB) project3_sweep_kernel(const Proj3Params p) {
  const int lp = blockIdx.z, chunk = blockIdx.y;
  for (int qb = 0; qb < p.NQB; qb++) {
    if (!p.need[(size_t)lp * p.NQB + qb]) continue;   // CTA-uniform
    project3_body<COUNT>(p, lp, qb, chunk);
    __syncthreads();
  }
}

constexpr size_t W3_SMEM = sizeof(double2) * (W3_R + 1) * 32 + sizeof(ProjMeta3) * W3_NW + sizeof(ProjQ3) * W3_QC;

// re-layout of the node table for the chunked kernel (project3_kernel, the fallback pass of project4_kernel): [row][PROJ_LP] -> [chunk][row][32]
__global__ void bessel_relayout_kernel(int num_xx, const double2* __restrict__ bes, double2* __restrict__ bes3) {
  const int i = blockIdx.x, t = threadIdx.x;  // t < PROJ_LP
  if (i >= num_xx || t >= PROJ_LP) return;
  bes3[((size_t)(t >> 5) * num_xx + i) * 32 + (t & 31)] = bes[(size_t)i * PROJ_LP + t];
}

}  // namespace cb200
