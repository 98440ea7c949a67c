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

#ifndef CB200_W2_QW
#define CB200_W2_QW 4
#endif
#ifndef CB200_W2_S
#define CB200_W2_S 8
#endif
#ifndef CB200_W2_NW
#define CB200_W2_NW 8
#endif
#ifndef CB200_W2_R
#define CB200_W2_R 176
#endif
#ifndef CB200_W2_MINB
#define CB200_W2_MINB 2
#endif
constexpr int W2_NW = CB200_W2_NW;       // warps per CTA
constexpr int W2_QW = CB200_W2_QW;       // wavenumbers per warp
constexpr int W2_S = CB200_W2_S;         // time samples per slab
constexpr int W2_QC = W2_NW * W2_QW;     // wavenumbers per CTA
constexpr int W2_R = CB200_W2_R;         // ring capacity in table rows (+1 mirror row)
constexpr int W2_NP = W2_QW * W2_S;      // (q, tau) pairs per warp per slab, one per lane
static_assert(W2_NP <= 32, "at most one (q, tau) pair per lane");

// Per-warp, per-slab metadata of the 32 (q, tau) pairs, structure-of-arrays so that the broadcast reads of the
// inner loop are as narrow as possible: a warp-wide shared load costs one LSU wavefront per 4 bytes per lane
// even when every lane reads the same address (ncu: 4 wavefronts per broadcast LDS.128), so bytes matter.
struct __align__(16) ProjMeta2 {
  double2 af[32];   // a = (x_{i+1}-x)/h ; fac = h^2 a / 6
  double2 s01[32];  // k-interpolated sources x dtau (temperature, E)
  double s2[32];    // lensing-potential source x dtau
  int off[32];      // byte offset of node row i0 in the ring
  int i0[32];       // table row (-1: pair not visited)
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

template <bool COUNT>
__global__ void __launch_bounds__(32 * W2_NW, CB200_W2_MINB) project2_kernel(const Proj2Params p) {
  constexpr int NW = W2_NW, QW = W2_QW, S = W2_S, QC = W2_QC, R = W2_R, NP = W2_NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* ring = reinterpret_cast<double2*>(smem_raw);                                        // [R+1][32], slot R mirrors slot 0
  ProjMeta2* meta = reinterpret_cast<ProjMeta2*>(smem_raw + sizeof(double2) * (R + 1) * 32);   // [NW]
  ProjQ2* qc = reinterpret_cast<ProjQ2*>(smem_raw + sizeof(double2) * (R + 1) * 32 + sizeof(ProjMeta2) * NW);
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
  long long ck_pro = 0, ck_pre = 0, ck_bar = 0, ck_ring = 0, ck_cmp = 0, ck_fin = 0, ck_t0 = clock64(), ck_t;
#define CK(var) do { if (COUNT) { ck_t = clock64(); var += ck_t - ck_t0; ck_t0 = ck_t; } } while (0)

  __syncthreads();
  const int n_lo = s_nlo, n_hi = s_nhi;

  // ---- sweep over conformal time in slabs of S samples, software-pipelined ----
  // The per-pair metadata of slab k+1 (table row, spline weights, k-interpolated dtau-weighted sources) needs
  // 14 global values per lane; they are fetched BEFORE the accumulation of slab k and consumed after it, so
  // their L2/HBM latency hides behind the FP64 work instead of stalling every warp at the slab barrier.
  ProjMeta2& wm = meta[warp];
  const unsigned char* ring_bytes = reinterpret_cast<const unsigned char*>(ring) + lane * 16;
  const int pq = lane % QW, pn = lane / QW;            // this lane's pair within a slab
  const ProjQ2& pc = wq[pq];
  const bool lane_has_pair = (lane < NP) && pc.valid;
  const int pw1 = lane_has_pair ? max(s_wn1[warp][pq], 1) : 0x7fffffff;
  const int pw2 = lane_has_pair ? min(s_wn2[warp][pq], pc.steps) : 0;

  double f_tau = 0, f_dtau = 0, f_s[3][4];
  bool f_valid = false;
  auto prefetch = [&](int nb) {   // issue the loads of slab starting at nb
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
  auto finish = [&](int nb, int par) {   // turn the prefetched values into the slab's metadata
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
  CK(ck_pro);
  for (int n_base = n_lo; n_base <= n_hi; n_base += S, par ^= 1) {
    const unsigned vm = vmask;                 // pairs of THIS slab (metadata already in wm)
    if (n_base + S <= n_hi) prefetch(n_base + S);
    CK(ck_pre);
    __syncthreads();  // previous slab fully consumed by every warp; this slab's window bounds visible
    CK(ck_bar);
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int w = 0; w < NW; w++) { lo = min(lo, s_wmin[par][w]); hi = max(hi, s_wmax[par][w]); }
    const bool direct = (hi - lo + 1) > R;
    if (hi >= 0 && !direct) {
      // slide the ring: fetch the rows of [lo, hi] that are not resident
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
        const int row = la + (e >> 5), ll = e & 31, slot = row % R;
        cp_async16(&ring[slot * 32 + ll], &bes[(size_t)row * 32 + ll]);
        if (slot == 0) cp_async16(&ring[R * 32 + ll], &bes[(size_t)row * 32 + ll]);
      }
      for (int e = tid; e < (lb2 - la2 + 1) * 32; e += 32 * NW) {
        const int row = la2 + (e >> 5), ll = e & 31, slot = row % R;
        cp_async16(&ring[slot * 32 + ll], &bes[(size_t)row * 32 + ll]);
        if (slot == 0) cp_async16(&ring[R * 32 + ll], &bes[(size_t)row * 32 + ll]);
      }
      if (COUNT && p.ring_stats && tid == 0) st_rows += max(0, lb - la + 1) + max(0, lb2 - la2 + 1);
      cp_async_wait_all();
      __syncthreads();
      if (tid == 0) { s_rlo = rlo; s_rhi = rhi; }
    }
    if (COUNT && p.ring_stats && tid == 0 && hi >= 0) { st_slabs++; st_direct += direct ? 1 : 0; }
    CK(ck_ring);

    // accumulate: S groups (one time sample each) of QW wavenumbers; lanes = multipoles.  Within a group the
    // QW accumulator sets are independent, so the FP64 chains of different pairs overlap.
#ifndef CB200_W2_SKIP_COMPUTE
#pragma unroll
    for (int nn = 0; nn < S; nn++) {
      if (((vm >> (nn * QW)) & ((1u << QW) - 1u)) == 0u) continue;  // warp-uniform
      const int n = n_base + nn;
      // stage 1: all loads of the group (metadata broadcasts + two table nodes per pair)
      double2 AF[QW], N0[QW], N1[QW], S01[QW];
      double S2[QW];
#pragma unroll
      for (int qq = 0; qq < QW; qq++) {
        const int pr = nn * QW + qq;
        AF[qq] = wm.af[pr];
        S01[qq] = wm.s01[pr];
        S2[qq] = wm.s2[pr];
        if (!direct) {
          const unsigned char* rp = ring_bytes + wm.off[pr];
          N0[qq] = *reinterpret_cast<const double2*>(rp);
          N1[qq] = *reinterpret_cast<const double2*>(rp + 512);
        } else {
          const int r0 = max(wm.i0[pr], 0);
          N0[qq] = __ldg(&bes[(size_t)r0 * 32 + lane]);
          N1[qq] = __ldg(&bes[(size_t)(r0 + 1) * 32 + lane]);
        }
      }
      // stage 2: cubic-spline value of j_l between the two nodes (cmbmain.f90:1515-1516), weights expanded so
      // the dependent FP64 chain after the table loads is three deep:  J = (a j0 + b j1) + (g0 p0 + g1 p1)
      double Jv[QW];
#pragma unroll
      for (int qq = 0; qq < QW; qq++) {
        const double a2 = AF[qq].x, b2 = 1 - a2, t = -(b2 * AF[qq].y);
        const double g0 = t * (a2 + 1), g1 = t * (2 - a2);
        const double v = (a2 * N0[qq].x + b2 * N1[qq].x) + (g0 * N0[qq].y + g1 * N1[qq].y);
        const bool act = (n >= n1[qq]) && (n <= n2[qq]);
        Jv[qq] = act ? v : 0.0;
        if (COUNT) {
          if (p.triples && act) my_triples++;
          if (p.ring_stats && lane == 0 && ((vm >> (nn * QW + qq)) & 1u)) st_pairs++;
        }
      }
      // stage 3: the three source accumulations
#pragma unroll
      for (int qq = 0; qq < QW; qq++) {
        acc[qq][0] += S01[qq].x * Jv[qq];
        acc[qq][1] += S01[qq].y * Jv[qq];
        acc[qq][2] += S2[qq] * Jv[qq];
      }
    }
#endif
    CK(ck_cmp);
    // metadata of the next slab (its loads were issued before the barrier above)
    __syncwarp();
    if (n_base + S <= n_hi) finish(n_base + S, par ^ 1);
    CK(ck_fin);
  }

  if (COUNT && p.triples) {
    unsigned long long t = my_triples;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(p.triples, t);
  }
  if (COUNT && p.ring_stats) {
    if (tid == 0) { atomicAdd(p.ring_stats + 0, st_slabs); atomicAdd(p.ring_stats + 1, st_direct); atomicAdd(p.ring_stats + 2, st_rows); }
    if (lane == 0) atomicAdd(p.ring_stats + 3, st_pairs);
    if (lane == 0) {
      atomicAdd(p.ring_stats + 4, (unsigned long long)ck_pro); atomicAdd(p.ring_stats + 5, (unsigned long long)ck_pre);
      atomicAdd(p.ring_stats + 6, (unsigned long long)ck_bar); atomicAdd(p.ring_stats + 7, (unsigned long long)ck_ring);
      atomicAdd(p.ring_stats + 8, (unsigned long long)ck_cmp); atomicAdd(p.ring_stats + 9, (unsigned long long)ck_fin);
    }
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

constexpr size_t W2_SMEM = sizeof(double2) * (W2_R + 1) * 32 + sizeof(ProjMeta2) * W2_NW + sizeof(ProjQ2) * W2_QC;

}  // namespace cb200
