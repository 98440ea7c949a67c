// K1 v4 — windowed line-of-sight projection: ALL multipoles per quarter-warp, warp-specialised producer/consumer.
//
// Same reference behaviour as project.cuh / project2.cuh / project3.cuh (camb/cmbmain.f90:478-498, 1295-1374,
// 1387-1420, 1440-1562, partial k-contraction of :2132-2264).  What changes against v3 (ncu of v3: no unit is
// saturated — 2.86e7 shared-memory wavefronts and 1.28e8 warp instructions per point, 25 % occupancy, two
// barriers per slab, metadata and ring filled three times because a CTA only covers 32 of the 88 multipoles):
//   * one CTA = (point, 32 wavenumbers) and ALL sampled multipoles: a quarter-warp still owns one (q, tau)
//     pair per step, but each lane now carries LK = ceil(n_l / 8) multipoles (l-slots li + 8k), so the 40 B of
//     per-pair metadata are read once per 88 multipoles instead of once per 32, the pair metadata is computed
//     once (not once per multipole chunk) and the ring is filled once per wavenumber block;
//   * warp specialisation: 12 CONSUMER warps (6 wavenumber groups x 2 time halves: the two warps of a group take
//     alternate time samples of a slab, every lane all octets; CB200_W4_TSPLIT=0: even / odd octets instead) do nothing
//     but the accumulation (shared-memory loads + FP64 FMAs).  4 PRODUCER warps: warp 0 keeps the ring filled with
//     TMA bulk copies (cp.async.bulk, one per run of table rows, byte count on the slab's `full` mbarrier); warps 1-3 compute
//     the (q, tau)-pair metadata of the coming slabs (table row, spline weight, k-interpolated d-tau-weighted
//     sources).  The source rows a slab needs (S time samples x 3 sources x {Src, ddSrc}, cut to the 8 source
//     wavenumbers that bracket the block's wavenumbers) are staged in shared memory by two 2-D TMA tensor copies of
//     the ring warp, so the metadata threads read them with LDS instead of 12 global loads per pair through running
//     64-bit pointers (122.2 -> 118.2 us/point; a separate pre-pass kernel that wrote S(q,tau) dtau for every pair
//     made this kernel 110.4 us/point but cost 8.6 us/point of HBM writes itself).  The two producer roles run in separate loops
//     and meet only the consumers (full / empty mbarriers per metadata buffer): the metadata runs up to NST slabs
//     ahead even while the ring warp waits for room, and there is no CTA-wide barrier inside the time loop.
//     setmaxnreg moves registers from the producer warps to the consumers;
//   * the ring holds whole table rows [row][n_l] (16 B nodes {j_l, j_l''}); the rows of a slab are known
//     analytically (x = q (tau0 - tau) is monotone in q and tau) and are tabulated once per CTA; a slab's rows are
//     fetched as soon as the highest row still needed by the oldest unreleased slab leaves room for them;
//   * the active multipoles of a pair are one run of l-slots, tracked by the producer and shipped with the metadata;
//     octet batches with no active lane anywhere in the warp are skipped (no loads, no FP64 issue);
//   * wavenumber blocks whose table window exceeds the ring are flagged and left to project3_kernel (host launches
//     it right after, on the flagged blocks only).
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "common.cuh"
#include "project.cuh"
#include "project3.cuh"

namespace cb200 {

#ifndef CB200_W4_GROUPS
#define CB200_W4_GROUPS 0
#endif
#ifndef CB200_W4_S
#define CB200_W4_S 4
#endif
#ifndef CB200_W4_NQG
#define CB200_W4_NQG 6
#endif
constexpr int W4_NQG = CB200_W4_NQG;  // wavenumber groups (4 wavenumbers = one warp's quarter-warps)
#ifndef CB200_W4_OSPLIT
#define CB200_W4_OSPLIT 0  // 1: on top of the time split, even / odd octets go to different warps (4 warps per wavenumber group)
#endif
constexpr int W4_NCW = 2 * W4_NQG * (1 + CB200_W4_OSPLIT);  // consumer warps: wavenumber group x time half (x octet parity)
constexpr int W4_QC = 4 * W4_NQG;     // wavenumbers per CTA
constexpr int W4_S = CB200_W4_S;      // time samples per slab
constexpr int W4_NPAIR = W4_QC * W4_S;
#ifndef CB200_W4_NPW
#define CB200_W4_NPW 4
#endif
constexpr int W4_NPW = CB200_W4_NPW;  // producer warps
constexpr int W4_NT = 32 * (W4_NCW + W4_NPW);  // threads per CTA
#ifndef CB200_W4_MG
#define CB200_W4_MG 1
#endif
constexpr int W4_MG = CB200_W4_MG;              // metadata groups: group g computes the slabs t = g (mod MG)
constexpr int W4_MW = (W4_NPW - 1) / W4_MG;     // warps per metadata group (producer warp 0 keeps the ring filled)
constexpr int w4_pnn() {  // largest divisor of S that the producer threads cover in one sweep
  int best = 1;
  for (int d = 1; d <= W4_S; d++)
    if (W4_S % d == 0 && d * W4_QC <= 32 * W4_MW) best = d;
  return best;
}
constexpr int W4_PNN = w4_pnn();               // time samples covered by one sweep of the metadata threads
constexpr int W4_PPT = W4_S / W4_PNN;          // (q, tau) pairs per producer thread: same q, samples W4_PNN apart
static_assert(W4_PNN >= 1 && W4_S % W4_PNN == 0, "slab shape");
#ifndef CB200_W4_MS
#define CB200_W4_MS 1
#endif
constexpr int W4_MS = CB200_W4_MS;  // slabs per metadata iteration: their independent chains overlap in one thread
static_assert(W4_MS == 1 || W4_PPT == 1, "several slabs per metadata iteration need one pair per thread and slab");
#ifndef CB200_W4_CREG
#define CB200_W4_CREG 128   // (136 / 104 until the two-pairs-per-step consumers: with those the metadata warps pace the kernel and
#endif                      //  do better with their full share of registers: 113.2 -> 111.8 us/point)
#ifndef CB200_W4_PREG
#define CB200_W4_PREG 128
#endif
#define CB200_STR2(x) #x
#define CB200_STR(x) CB200_STR2(x)
constexpr int W4_SMEM_TOTAL = 227 * 1024;

struct __align__(16) Proj4Rec {  // what the metadata threads compute for a (q, tau) pair
  double a;   // spline weight of node row i0: (x1 - x) / (x1 - x0)
  int off;    // byte offset of node row i0 in the ring (a multiple of 128) | stretch of the abscissa grid (low 3 bits)
  int jr;     // active multipole slots [jlo, jhi]: jlo | (jhi + 1) << 8 ; none: 127
};

struct Proj4Params {
  // 2-D TMA descriptors of the resident sources [point][tau][source][k] and of the chunk's second derivatives, both
  // seen as [rows = (point, tau, source)][k]: a slab's S x 3 consecutive rows x W4_KSP wavenumbers are ONE box
  const CUtensorMap* tmaps;   // device memory: [0] sources, [1] second derivatives (written by the host before the launch)
  PointView v;
  int p0, nl, num_xx, NQB, tensors;
  int R, rb;                // ring capacity in rows (+1 mirror row), row stride in bytes (= 128 * octets)
  double max_eta_k;
  const double* ddsrc;
  double h2o6[8];           // (x1 - x0)^2 / 6 of every stretch of the Bessel abscissa grid
  const double* bx;
  const double2* bes;       // [num_xx][rb / 16]: rows packed at the ring's row stride
  const double* initpower;
  double* part;             // [chunk][NQB][6][PROJ_LP]
  double* delta;            // optional [chunk][NQ][PROJ_LP][3]
  double* raw;              // [chunk][NQB][wavenumber group][3 LK][32 lanes]: the time integrals as the consumers leave them
  ProjQ3* qcg;              // [chunk][NQB][QC]: the per-wavenumber constants of the block (for project4_finish_kernel)
  unsigned* flg;            // [chunk][NQB][consumer warp][lane]: reached | doint << 16, one bit per octet of the lane
  unsigned long long* triples;
  unsigned long long* ring_stats;  // optional [16]
  unsigned char* fallback;  // [chunk][NQB]: 1 = a slab of this block needs more table rows than the ring holds
  LinSegs bseg;
  // the last (coarsest) stretch of the Bessel abscissa grid, where almost every x = q (tau0 - tau) lies: scalar copies
  // so that the producers reach them as immediate constant-bank operands (filled by w4_set_last_stretch)
  double bl_lo, bl_hi, bl_step, bl_inv;
  int bl_first, bl_nseg;
  int ls[PROJ_LP];
};

inline void w4_set_last_stretch(Proj4Params& p) {
  const LinSegs& g = p.bseg;
  const int r = g.n - 1;
  p.bl_lo = g.seg[r][0]; p.bl_hi = g.seg[r][1]; p.bl_step = g.seg[r][2]; p.bl_inv = g.inv_step[r];
  p.bl_first = (int)g.seg[r][3]; p.bl_nseg = g.npoints - p.bl_first;
  for (int i = 0; i < 8; i++) {
    // spacing of the stretch as the table holds it: lo + step * 1 - lo, rounded like the abscissae themselves
    const double h = i < g.n ? (g.seg[i][0] + g.seg[i][2] * 1.0) - g.seg[i][0] : 0.0;
    p.h2o6[i] = h * h / 6;
  }
}

#ifndef CB200_W4_LEANMETA
#define CB200_W4_LEANMETA 1  // producers: division-free table lookup in the last stretch, window thresholds kept in registers
#endif
#ifndef CB200_W4_OCTSKIP
#define CB200_W4_OCTSKIP 0   // consumers: inside a batch of KB octets, an octet with no active lane in the warp loads nothing
#endif

// lin_locate_desc for the Bessel abscissae, with a fast path for the last stretch.  The index is the truncation of the
// true IEEE quotient (x - lo) / step (camb/utils.F90:81-111).  t = (x - lo) * (1/step) differs from that quotient by a few
// ulp (< 1e-11 for t < 2e4), so the two truncations agree unless t lies within 1e-7 of an integer - only then is the
// division carried out.  Everything else (x0, x1 rebuilt without FMA) is lin_locate_desc's arithmetic.
__device__ __forceinline__ int w4_locate(const Proj4Params& p, double v, double& x0, double& x1, double& inv_h, int& rs) {
  if (v < p.bl_hi && v >= p.bl_lo) {
    const double d = __dsub_rn(v, p.bl_lo);
    const double t = __dmul_rn(d, p.bl_inv);
    int j = (int)t;
    const double fr = __dsub_rn(t, (double)j);
    if (fr < 1e-7 || fr > 1 - 1e-7) j = (int)(__ddiv_rn(d, p.bl_step));
    x0 = __dadd_rn(p.bl_lo, __dmul_rn(p.bl_step, (double)j));
    x1 = (j + 1 < p.bl_nseg) ? __dadd_rn(p.bl_lo, __dmul_rn(p.bl_step, (double)(j + 1))) : p.bl_hi;
    inv_h = p.bl_inv;
    rs = p.bseg.n - 1;
    return p.bl_first + j;
  }
  // the finer stretches (x < 150): lin_locate_desc's arithmetic, also reporting the stretch
  const LinSegs& g = p.bseg;
#pragma unroll 1
  for (int r = g.n - 1; r >= 0; r--) {
    const double lo = g.seg[r][0], hi = g.seg[r][1];
    if (v < hi && v >= lo) {
      const double step = g.seg[r][2];
      const int first = (int)g.seg[r][3];
      const int j = (int)(__ddiv_rn(__dsub_rn(v, lo), step));
      const int nseg = ((r + 1 < g.n) ? (int)g.seg[r + 1][3] : g.npoints) - first;
      x0 = __dadd_rn(lo, __dmul_rn(step, (double)j));
      x1 = (j + 1 < nseg) ? __dadd_rn(lo, __dmul_rn(step, (double)(j + 1))) : hi;
      inv_h = g.inv_step[r];
      rs = r;
      return first + j;
    }
  }
  x0 = x1 = g.highest;
  inv_h = 0;
  rs = g.n - 1;
  return g.npoints;
}

#ifndef CB200_W4_PAIR2
#define CB200_W4_PAIR2 1   // 1: a consumer warp takes BOTH of its time samples of a slab in one step (S = 4, time-split): one prologue
                           //    per two pairs, octet-granular batches whose two chains belong to the two samples
#endif
#ifndef CB200_W4_WUNROLL
#define CB200_W4_WUNROLL 1   // unroll of the window set-up loop (its FP64 divisions are one dependent chain per octet)
#endif
constexpr int W4_WUNROLL = CB200_W4_WUNROLL;
#ifndef CB200_W4_KB
#define CB200_W4_KB 2   // octets per batch of loads in flight (time-split consumers: 1: 136.3 us/point, 2: 132.8, 3: 137.3, 4: 145.6)
#endif
#ifndef CB200_W4_UNSAFE_NORINGWAIT
#define CB200_W4_UNSAFE_NORINGWAIT 0  // timing experiment only: WRONG results
#endif
#ifndef CB200_W4_PBAL
#define CB200_W4_PBAL 0   // 1: every producer warp (the ring warp too) computes the metadata of ONE time sample of the slab
#endif
constexpr bool W4_PBAL = CB200_W4_PBAL != 0;
static_assert(!W4_PBAL || (W4_S == W4_NPW && W4_QC <= 32 && W4_MG == 1 && W4_MS == 1), "balanced producers: one time sample per producer warp");
// (last measured with the final consumers: 157.7 us/point against 109.6, and its C_l differ from the default build's at 1e-6 -
//  the variant has not followed the source staging; kept as a record of the experiment, not as an option to ship)
#ifndef CB200_W4_UNSAFE_FREEMETA
#define CB200_W4_UNSAFE_FREEMETA 0    // timing experiment only: WRONG results
#endif
#ifndef CB200_W4_UNSAFE_ABLATE
#define CB200_W4_UNSAFE_ABLATE 0      // timing experiments only: WRONG results
#endif
#ifndef CB200_W4_PREDLOAD
#define CB200_W4_PREDLOAD 0
#endif
#ifndef CB200_W4_TSPLIT
#define CB200_W4_TSPLIT 1  // 1: the two consumer warps of a wavenumber group split the slab's time samples, not the octets
#endif
constexpr bool W4_TS = CB200_W4_TSPLIT != 0;
constexpr bool W4_OS = CB200_W4_OSPLIT != 0;      // octet parity split on top of the time split
constexpr bool W4_OSPL = W4_OS || !W4_TS;         // a consumer lane holds every second octet
static_assert(!W4_OS || W4_TS, "the octet-parity split is defined on top of the time split");
#ifndef CB200_W4_UNROLL
#define CB200_W4_UNROLL 1
#endif
constexpr int W4_UNROLL = CB200_W4_UNROLL;
#ifndef CB200_W4_NST
#define CB200_W4_NST 4
#endif
constexpr int W4_NST = CB200_W4_NST;  // metadata buffers = slabs the producer may run ahead of the consumers
constexpr int W4_MB = W4_NPAIR * 40;   // one metadata buffer: {a, off, jr} 16 B + {S_T, S_E} dtau 16 B + S_phi dtau 8 B per (q, tau) pair
// staging of the raw source rows of a slab (ring warp): S time samples x 3 sources x {Src, ddSrc} rows of W4_KSP
// consecutive source wavenumbers (the few that bracket the block's integration wavenumbers)
constexpr int W4_KSP = 8;   // doubles per staged row
#ifndef CB200_W4_SPD
#define CB200_W4_SPD 2
#endif
// the copies of slab t + SPD are issued when slab t's metadata buffer is free (the consumers have released slab t - NST,
// so the metadata threads are done with every stage up to that slab's): NST + SPD stages keep a refill off live rows
constexpr int W4_SPD = CB200_W4_SPD, W4_NSR = W4_NST + W4_SPD;
constexpr int W4_SRAW_STAGE = W4_S * 3 * 2 * W4_KSP * 8;   // bytes: [Src | ddSrc][S x 3 rows][KSP]
static_assert((W4_S * 3 * W4_KSP * 8) % 128 == 0, "TMA box destinations are 128-byte aligned");
constexpr size_t W4_META_BYTES = (size_t)W4_NST * W4_MB + (size_t)W4_NSR * W4_SRAW_STAGE;
constexpr size_t W4_QC_BYTES = sizeof(ProjQ3) * W4_QC;
constexpr size_t W4_MISC_BYTES = 768;
// octets per ring row for a multipole set of noct octets (the kernel's template instances)
inline int w4_octets(int noct) { return noct <= 6 ? 6 : noct <= 11 ? 11 : 12; }
inline size_t w4_slab_table_bytes(int NT) { return (size_t)8 * ((NT + W4_S - 1) / W4_S + 2); }
// per-(wavenumber, multipole slot) integration windows {n1, n2} as 2 x u16, rows padded to an odd word count
inline size_t w4_wtab_bytes(int LK) { return (size_t)4 * W4_QC * (8 * LK + 1); }
// ring rows that fit for a row stride of rb bytes (one extra mirror row)
inline int w4_ring_rows(int LK, int NT) {
  return (int)((W4_SMEM_TOTAL - W4_META_BYTES - W4_QC_BYTES - W4_MISC_BYTES - w4_slab_table_bytes(NT) - w4_wtab_bytes(LK)) /
               (size_t)(LK * 128)) - 1;
}
inline size_t w4_smem_bytes(int LK, int NT, int R) {
  return (size_t)(R + 1) * LK * 128 + W4_META_BYTES + W4_QC_BYTES + W4_MISC_BYTES + w4_slab_table_bytes(NT) + w4_wtab_bytes(LK);
}

__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "W4_WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
      " @p bra W4_DONE_%=;\n"
      " bra W4_WAIT_%=;\n"
      "W4_DONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
// TMA tensor copy of a 2-D box global -> shared (coordinates in elements, innermost first; out-of-range elements read 0)
__device__ __forceinline__ void tma_box_g2s(void* dst, const CUtensorMap* tm, int c0, int c1, unsigned long long* b) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(b))
               : "memory");
}
// predicated 16-byte shared load: the destination keeps its (undefined) previous contents when !pred
__device__ __forceinline__ double2 lds128_if(unsigned saddr, bool pred) {
  double2 r;
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p ld.shared.v2.f64 {%0, %1}, [%2];\n}\n"
               : "=d"(r.x), "=d"(r.y)
               : "r"(saddr), "r"((unsigned)pred));
  return r;
}

// unpredicated 16-byte shared load.  The consumers load the nodes of every lane of an octet batch that has an active
// lane anywhere in the warp (the address is always inside the ring; the value of an inactive lane is discarded by a
// select): a predicated load leaves its destination live from the top of the loop, which costs a register pair per
// load for the whole loop instead of per batch.
__device__ __forceinline__ double2 lds128(unsigned saddr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(r.x), "=d"(r.y) : "r"(saddr));
  return r;
}

#ifndef CB200_W4_FIN_MINB
#define CB200_W4_FIN_MINB 3   // CTAs of project4_finish_kernel per SM (register cap 56; 2: 70 registers, 2.86 ms per 512 points against 2.48)
#endif
#ifndef CB200_W4_SPLIT_EPI
#define CB200_W4_SPLIT_EPI 1   // 1: Limber values and the partial k-contraction run in project4_finish_kernel, not in this CTA's tail
#endif

// last time sample a wavenumber integrates to (SourceSteps of IntegrationVars_Init, cmbmain.f90:1387-1400)
__device__ __forceinline__ int w4_steps(const Proj4Params& p, double qv, int nt, double tau0, const double* tau) {
  const double max_etak_tensor = p.max_eta_k / 10;
  int step = 2;
  for (int i = nt; i >= 2; i--) {
    double xf = __dmul_rn(qv, __dsub_rn(tau0, tau[i - 1]));
    bool ok = xf > 1.e-8;
    if (p.tensors) ok = ok && (__dmul_rn(qv, tau[i - 1]) < max_etak_tensor);
    if (ok) { step = i; break; }
  }
  return step;
}

// per-wavenumber constants (InterpolateSources set-up, cmbmain.f90:1307-1320)
__device__ __forceinline__ ProjQ3 w4_q_consts(const Proj4Params& p, int pt, int qi, int nq, const double* ip) {
  const PointView& v = p.v;
  const int nt = v.n_tau[pt], nk = v.n_k[pt];
  const double tau0 = v.thermo[(size_t)pt * 5];
  const double* tau = v.tau + (size_t)pt * v.NT;
  const double* ksrc = v.ksrc + (size_t)pt * v.NK;
  ProjQ3 c;
  c.valid = qi < nq;
  c.pad = 0;
  if (c.valid) {
    const double qv = v.q[(size_t)pt * v.NQ + qi];
    const double dqv = v.dq[(size_t)pt * v.NQ + qi];
    c.q = qv;
    c.w = (p.tensors ? tensor_power_dev(ip, qv) : scalar_power_dev(ip, qv)) * (dqv / qv);
    // klo = first index in [1, nk-1] with qv <= ksrc[klo] (else nk-1): the reference's linear scan
    // (cmbmain.f90:1313-1316) as a binary search - same result on the ascending source grid, 8 loads instead of ~100
    int klo = nk - 1;
    {
      int lo_ = 1, hi_ = nk - 1;
      while (lo_ < hi_) {
        const int mid = (lo_ + hi_) >> 1;
        if (qv > ksrc[mid]) lo_ = mid + 1; else hi_ = mid;
      }
      klo = lo_;
    }
    c.klo = klo;
    const double ho = ksrc[klo] - ksrc[klo - 1];
    c.a0 = (ksrc[klo] - qv) / ho;
    c.b0 = (qv - ksrc[klo - 1]) / ho;
    c.ho2o6 = ho * ho / 6;
    c.a03h = (c.a0 * c.a0 * c.a0 - c.a0);
    c.b03h = (c.b0 * c.b0 * c.b0 - c.b0);
    c.steps = w4_steps(p, qv, nt, tau0, tau);
  } else {
    c.q = 1; c.w = 0; c.klo = 1; c.a0 = c.b0 = c.a03h = c.b03h = c.ho2o6 = 0; c.steps = 0;
  }
  return c;
}

// integration window [n1, n2] of one (wavenumber, multipole) and the two flags the epilogue needs (cmbmain.f90:1387-1420,
// 1440-1470): reached = the window is entered at all, doint = the time integral is formed (else Limber only)
// highest multipole a wavenumber contributes to (cmbmain.f90:1402-1408): a constant of the wavenumber, taken out of the
// per-multipole window so that a lane derives it once instead of once per octet
__device__ __forceinline__ int w4_llmax(double qv, double tau0) {
  int llmax = (int)llround(__dmul_rn(qv, tau0));
  if (llmax < 15) llmax = 17;
  else llmax = (int)llround(__dmul_rn(qv, __dadd_rn(tau0, __ddiv_rn(6 * kPi, qv))));
  return llmax;
}
__device__ __forceinline__ void w4_window(const Proj4Params& p, const ProjQ3& myq, int llmax, bool lvalid, int l, double tau0,
                                          const double* tau, const LinSegs& tseg, int& n1, int& n2, bool& reached, bool& doint) {
  n1 = 0; n2 = 0; reached = false; doint = false;
  if (myq.valid && lvalid) {
    const double qv = myq.q;
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
        reached = true;
        bool di = true;
        if (!p.tensors) {
          double qmax_int = __ddiv_rn((double)(max(850, l) * 3), tau0);
          qmax_int = __dmul_rn(qmax_int, (double)1.2f);
          di = qv < qmax_int;
        }
        if (di) {
          doint = true;
          n1 = lin_index_of(tseg, tmin);
          n2 = min(myq.steps, lin_index_of(tseg, tmax));
        } else {
          n1 = n2 = 0x7fff;
        }
      }
    }
  }
}

template <int LK, bool COUNT, int KLIM>
__global__ void __launch_bounds__(W4_NT, 1) project4_kernel(const Proj4Params p) {
  constexpr int NCW = W4_NCW, NQG = W4_NQG, QC = W4_QC, S = W4_S, NPAIR = W4_NPAIR;
  // octets per consumer lane.  W4_TS: every octet (k), the warp pair of a wavenumber group shares the slab's time
  // samples (lh = 0: samples 0, 2, ..; lh = 1: samples 1, 3, ..) and the two partial sums meet in the epilogue: the
  // per-pair metadata is read once per 8 LK multipoles instead of once per 4 LK.  Otherwise the pair splits the octets
  // (k -> 2 k + lh).  K2 = octets that need the lensing-potential accumulator: for a scalar run every multipole above
  // 400 takes the Limber value in the epilogue (cmbmain.f90:1546-1556), so the host passes KLIM < LK when all octets
  // >= KLIM lie above 400 (88-multipole set: KLIM = 6) and their third accumulator is never formed.
  // W4_OS: both at once - 4 warps per wavenumber group, each with half of the time samples and the octets of one parity
  // (more warps with fewer accumulators each: the loop is latency-bound, not bandwidth-bound).
  constexpr int LKH = W4_OSPL ? (LK + 1) / 2 : LK;
  constexpr int K2 = !W4_TS ? LKH : (!W4_OS ? KLIM : (KLIM == LK ? LKH : KLIM / 2));
  static_assert(W4_TS || KLIM == LK, "the Limber cut of the accumulators needs the time-split consumers");
  static_assert(!W4_OS || KLIM == LK || KLIM % 2 == 0, "octet-parity split: the Limber cut must fall on an even octet");
#define W4_OCT(k) (W4_OSPL ? 2 * (k) + op : (k))
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = p.R;
  constexpr int rb = LK * 128;  // ring row stride in bytes
  unsigned char* ring = smem_raw;                                            // [(R+1)][rb]
  unsigned char* meta_base = smem_raw + (size_t)(R + 1) * rb;                // 2 x {af[NPAIR], s01[NPAIR], rec[NPAIR]}
  ProjQ3* qc = reinterpret_cast<ProjQ3*>(meta_base + W4_META_BYTES);         // [QC]
  int* s_q1 = reinterpret_cast<int*>(meta_base + W4_META_BYTES + W4_QC_BYTES);  // [2][QC]: first visited sample, per consumer half
  int* s_q2 = s_q1 + 2 * QC;                                                 // [2][QC]: last visited sample
  int* s_misc = s_q2 + 2 * QC;                                               // [2]: block needs the fallback
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_misc + 4);  // (s_misc[2]: block needs the fallback)  // full[NST], empty[NST]
  constexpr int NJ = 8 * LK, NJP = NJ + 1;
  unsigned* s_wtab = reinterpret_cast<unsigned*>(meta_base + W4_META_BYTES + W4_QC_BYTES + W4_MISC_BYTES);  // [QC][NJP]
  int2* s_win = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(s_wtab) + (size_t)4 * QC * NJP);  // [slab] rows lo, hi

  const PointView& v = p.v;
  const int lp = blockIdx.y, pt = p.p0 + lp, qb = blockIdx.x;
  const int nq = v.n_q[pt];
  const int q0 = qb * QC;
  if (q0 >= nq) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool consumer = warp < NCW;
  const int qr = lane >> 3, li = lane & 7;   // quarter-warp = wavenumber slot ; octet position
  const int wg = warp % NQG, lh = (warp / NQG) & 1;  // consumer: wavenumber group, half (time samples or octets)
  const int op = W4_OS ? (warp / (2 * NQG)) & 1 : (W4_TS ? 0 : lh);  // octet parity of this warp (when octets are split)

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
  const int noct = (p.nl + 7) >> 3;  // octets that hold multipoles

  // ---- per-wavenumber constants (InterpolateSources set-up, cmbmain.f90:1307-1320) ----
  // The full set (source-grid bracket by binary search, spline weights, primordial power: a chain of dependent global
  // loads) is computed by 24 lanes of the ring warp WHILE the consumers derive their integration windows, which only need
  // q and the last time sample: one CTA-wide barrier instead of two, and the two latency chains overlap.
  if (tid >= 32 * NCW && tid < 32 * NCW + QC) {
    const int t = tid - 32 * NCW;
    const ProjQ3 c = w4_q_consts(p, pt, q0 + t, nq, ip);
    qc[t] = c;
#if CB200_W4_SPLIT_EPI
    p.qcg[((size_t)lp * p.NQB + qb) * QC + t] = c;
#endif
  }
  if (tid == 32 * NCW + 31) {
    s_misc[2] = 0;
    for (int i = 0; i < W4_NST; i++) {
      mbar_init(s_bar + i, 32 * (1 + W4_MW)); // full: ring warp + one metadata group arrive (+ the bytes of the bulk copies)
      mbar_init(s_bar + W4_NST + i, NCW);    // empty: one lane per consumer warp
    }
    for (int i = 0; i < W4_NSR; i++) mbar_init(s_bar + 32 + i, 1);  // raw source rows of a slab landed
  }

  // ---- integration windows of this lane's wavenumber (quarter) and its LK multipoles (consumer warps) ----
  const int myqi = wg * 4 + qr;
  ProjQ3 myq;   // the three fields the windows need; the full record is in qc[] after the barrier
  myq.valid = consumer && (q0 + myqi < nq);
  myq.q = myq.valid ? v.q[(size_t)pt * v.NQ + q0 + myqi] : 1.0;
  myq.steps = myq.valid ? w4_steps(p, myq.q, nt, tau0, tau) : 0;
  // Integration windows [n1, n2] of every (wavenumber, multipole) go to shared memory; both bounds fall with l
  // (tmin and tmax do, and the index lookup is monotone), so for a given (q, tau) pair the active multipoles are
  // ONE run of l-slots [jlo, jhi], which the producer tracks incrementally and ships with the pair metadata.
  // Never-active entries keep that structure: cut from above (l > llmax, window not reached, padding, invalid q)
  // = {0, 0}; cut from below (q >= qmax_int: Limber only) = {0x7fff, 0x7fff}.
  unsigned win[COUNT ? LKH : 1];  // COUNT builds re-derive the mask from the exact windows and compare
  unsigned reached = 0, doint = 0;
  constexpr int KS = (LKH + 1) / 2;  // W4_TS: half 0 owns the windows / the epilogue of the octets below KS, half 1 the others
  if (consumer) {
    int un1 = 0x7fffffff, un2 = 0;
    const int my_llmax = myq.valid ? w4_llmax(myq.q, tau0) : 0;
    // each half derives the windows of its own octets only (they meet in s_wtab / the atomics below); COUNT builds
    // need every window in registers.  Not unrolled when split: the body is ~150 instructions
    const int kb = (W4_TS && !COUNT && lh == 1) ? KS : 0, ke = (W4_TS && !COUNT && lh == 0) ? KS : LKH;
#pragma unroll(COUNT || !W4_TS ? LKH : W4_WUNROLL)
    for (int k = kb; k < ke; k++) {
      int n1, n2;
      bool w_reached, w_doint;
      const int j = li + 8 * W4_OCT(k);
      const bool lvalid = j < p.nl;
      const int l = lvalid ? p.ls[j] : 0;
      w4_window(p, myq, my_llmax, lvalid, l, tau0, tau, tseg, n1, n2, w_reached, w_doint);
      if (w_reached) reached |= 1u << k;
      if (w_doint) doint |= 1u << k;
      if (j < NJ) s_wtab[myqi * NJP + j] = (unsigned)n1 | ((unsigned)n2 << 16);
      const bool live = (n2 >= n1) && n1 > 0 && n1 < 0x7fff;
      if (live) {
        un1 = min(un1, n1);
        un2 = max(un2, n2);
      }
      if (COUNT) win[k] = live ? ((unsigned)n1 | ((unsigned)(n2 - n1) << 16)) : 0x7fffu;
    }
#if CB200_W4_SPLIT_EPI
    p.flg[(((size_t)lp * p.NQB + qb) * NCW + warp) * 32 + lane] = reached | (doint << 16);
#endif
    // union over the quarter's multipoles: which (q, tau) pairs have to be visited at all
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      un1 = min(un1, __shfl_xor_sync(0xffffffffu, un1, o));
      un2 = max(un2, __shfl_xor_sync(0xffffffffu, un2, o));
    }
    if (li == 0) {  // per half; the halves (and the block's union) meet after the barrier
      s_q1[lh * QC + myqi] = un1 <= un2 ? un1 : 0x7fffffff;
      s_q2[lh * QC + myqi] = un1 <= un2 ? un2 : 0;
    }
  }
  __syncthreads();
  static_assert(W4_QC <= 32, "the block's union of visited samples is reduced by one warp-wide shuffle tree");
  int n_lo = 0x7fffffff, n_hi = 0;
  if (lane < QC) { n_lo = min(s_q1[lane], s_q1[QC + lane]); n_hi = max(s_q2[lane], s_q2[QC + lane]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_lo = min(n_lo, __shfl_xor_sync(0xffffffffu, n_lo, o));
    n_hi = max(n_hi, __shfl_xor_sync(0xffffffffu, n_hi, o));
  }
  const int nslab = (n_lo <= n_hi) ? (n_hi - n_lo) / S + 1 : 0;

  // ---- analytic table-row window of every slab (superset of the rows its visited pairs touch):
  //      x = q (tau0 - tau) decreases with tau and grows with q, the block's wavenumbers are ascending ----
  {
    const int nvalid = min(QC, nq - q0);
    const double q_first = qc[0].q, q_last = qc[nvalid - 1].q;
    for (int t = tid; t < nslab; t += W4_NT) {
      const int nb = n_lo + t * S;
      const int na = max(nb, 1), nz = min(min(nb + S - 1, n_hi), nt);
      const double xhi = fabs(__dmul_rn(q_last, __dsub_rn(tau0, tau[na - 1])));
      const double xlo = fabs(__dmul_rn(q_first, __dsub_rn(tau0, tau[nz - 1])));
      double x0, x1, ih;
      int bhi = lin_locate(p.bseg, xhi, x0, x1, ih);
      int blo = lin_locate(p.bseg, xlo, x0, x1, ih);
      bhi = min(bhi, p.num_xx - 1);
      blo = min(blo, p.num_xx - 1);
      s_win[t] = make_int2(max(blo - 1, 0), bhi);  // rows i0 .. i0 + 1
      if (bhi - max(blo - 1, 0) + 1 > R) s_misc[2] = 1;
    }
  }
  __syncthreads();
  // A slab that needs more table rows than the ring holds (in practice only the first, log-spaced wavenumber
  // block): the whole block is left to the chunked kernel (project3.cuh), which the host launches on the flagged
  // blocks right after this kernel.
  if (p.fallback) {
    if (tid == 0) p.fallback[(size_t)lp * p.NQB + qb] = (unsigned char)(s_misc[2] != 0);
  }
  if (s_misc[2] != 0) return;

  unsigned long long my_triples = 0, st_slabs = 0, st_rows = 0, st_late = 0, st_mismatch = 0;
  long long st_wait_all = 0, st_wait_prod = 0, st_wait_ring = 0;
  long long ck_a = 0, ck_b = 0, ck_c = 0, ck_t0 = clock64(), ck_t;
#define CK4(var) do { if (COUNT) { ck_t = clock64(); var += ck_t - ck_t0; ck_t0 = ck_t; } } while (0)

  // Register re-balancing: the four producer warps (one warp group, one warp per SM sub-partition) give registers
  // back, the consumers take them.  The pool is what the CTA was launched with (64 K / threads, per thread), so
  // NPW x (launch - PREG) must cover NCW x (CREG - launch).
  static_assert(W4_NPW % 4 == 0 && W4_NCW % 4 == 0 && W4_NPW * (65536 / W4_NT / 8 * 8 - CB200_W4_PREG) >= W4_NCW * (CB200_W4_CREG - 65536 / W4_NT / 8 * 8), "setmaxnreg works on aligned groups of 4 warps");
  if (!consumer) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " CB200_STR(CB200_W4_PREG) ";\n");
    // =========================================== PRODUCER ===========================================
    // producer warp 0 keeps the ring filled (TMA bulk copies); the other producer warps compute the pair metadata.
    // The two roles only meet the consumers (full / empty barriers), never each other: the metadata runs ahead as far
    // as the NST buffers allow even while the ring warp waits for room.
    const bool ring_warp = (warp == NCW);
    // W4_PBAL: producer warp pw owns time sample pw of every slab (lane = wavenumber), the ring warp included: the
    // metadata instructions are spread over the four SM sub-partitions instead of three
    const int m_tid = W4_PBAL ? (warp - NCW) * QC + lane : tid - 32 * (NCW + 1);   // metadata thread index
    const int m_grp = (ring_warp || W4_PBAL) ? 0 : min(m_tid / (32 * W4_MW), W4_MG);  // == MG: spare warp (exits)
    const int m_pair = m_tid - m_grp * 32 * W4_MW;   // pair of the slab owned by this (metadata) thread
    const bool m_live = W4_PBAL ? lane < QC : (!ring_warp && m_pair < QC * W4_PNN);  // spare threads only take part in the barriers
    if (!ring_warp && m_grp >= W4_MG) return;
    const int m_qi = m_live ? m_pair % QC : 0, m_nn = W4_PBAL ? warp - NCW : m_pair / QC;
    const ProjQ3& pc = qc[m_qi];  // read from shared memory where needed: the producer runs on few registers
    const int pw1 = (pc.valid && m_live) ? max(min(s_q1[m_qi], s_q1[QC + m_qi]), 1) : 0x7fffffff;  // (ring warp: never valid)
    const int pw2 = pc.valid ? min(max(s_q2[m_qi], s_q2[QC + m_qi]), pc.steps) : 0;
    constexpr int PPT = W4_PPT, PNN = W4_PNN, MS = W4_MS, NU = PPT * MS;
    double f_tau[NU], f_dtau[NU];
    bool f_valid[NU];
    // Conformal times of the coming slab(s), loaded one iteration ahead.  (The sources themselves no longer pass through
    // these threads: source_q_kernel has interpolated them to the integration wavenumbers, the ring warp copies a
    // slab's values straight into the metadata buffer.)
    int pf_n = n_lo + m_grp * MS * S + m_nn;                         // time sample of pair u = 0 at the next call
    auto prefetch = [&]() {
#pragma unroll
      for (int u = 0; u < NU; u++) {
        const int du = (u / PPT) * S + (u % PPT) * PNN;   // slab u / PPT of the iteration (compile-time constant)
        const int n = pf_n + du;
        f_valid[u] = (n >= pw1) && (n <= pw2);
        if (f_valid[u] && !CB200_W4_UNSAFE_FREEMETA) {
          f_tau[u] = __ldg(tau + n - 1);
          f_dtau[u] = (n >= 2) ? __ldg(dtau + n - 1) : 0.0;  // Source_q(1,:) is forced to zero (IntegrationVars_Init, cmbmain.f90:1380)
        }
      }
      pf_n += W4_MG * MS * S;
    };
    // ---- InterpolateSources (cmbmain.f90:1295-1374) for the slab's (q, tau) pairs.  The metadata threads used to fetch
    // their 12 source values with global loads (one slab ahead, 24 registers, running 64-bit pointers): with the table
    // lookup that made ~320 instructions per pair and the three metadata warps paced the kernel.  Now the S x 3
    // consecutive source rows of a slab, cut down to the W4_KSP source wavenumbers that bracket the block's integration
    // wavenumbers, are staged by TWO TMA tensor copies of the ring warp (one box of Src, one of ddSrc) as soon as the
    // slab's metadata buffer is free, and the metadata threads evaluate the spline out of shared memory.  Blocks whose
    // wavenumbers span more source intervals than a staged row holds (two log-spaced blocks per point) keep the loads
    // from global memory.
    const int r_nvalid = min(QC, nq - q0);
    const int r_k0 = (qc[0].klo - 1) & ~1;   // even: TMA wants the box to start on a 16-byte boundary (odd: illegal instruction)
    const bool r_wide = qc[r_nvalid - 1].klo - r_k0 + 1 > W4_KSP;
    unsigned char* sraw = meta_base + (size_t)W4_NST * W4_MB;   // [NSR][Src | ddSrc][S * 3][KSP] doubles
    unsigned long long* s_sbar = s_bar + 32;                     // [NSR] "raw rows landed"
    auto stage_sources = [&](int t) {  // issue the copies of slab t (ring warp)
      if (r_wide || t >= nslab) return;
      if (lane == 0) {
        const int st = t % W4_NSR;
        asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(s_sbar + st)),
                     "r"((unsigned)W4_SRAW_STAGE) : "memory");
        const int n0 = n_lo + t * S - 1;   // first time sample of the slab (0-based); rows past the grid are never used
        unsigned char* d = sraw + (size_t)st * W4_SRAW_STAGE;
        tma_box_g2s(d, p.tmaps, r_k0, (pt * v.NT + n0) * 3, s_sbar + st);
        tma_box_g2s(d + W4_SRAW_STAGE / 2, p.tmaps + 1, r_k0, (lp * v.NT + n0) * 3, s_sbar + st);
      }
    };
    // this metadata thread's wavenumber: spline weights of InterpolateSources (cmbmain.f90:1317-1320), ho^2/6 folded in
    const double i_a0 = pc.a0, i_b0 = pc.b0, i_a3 = pc.a03h * pc.ho2o6, i_b3 = pc.b03h * pc.ho2o6;
    const int i_ko = pc.klo - 1 - r_k0;      // column of source wavenumber klo - 1 in a staged row
    // rows [a, b] of the table -> ring.  The table is packed at the ring's row stride and slot = row mod R, so a run of
    // rows is ONE TMA bulk copy (two when it wraps around the end of the ring).  Slot 0 is mirrored behind slot R-1
    // (a pair reads rows i0 and i0 + 1 at a fixed offset): a run that continues through slot 0 simply copies one row
    // more behind slot R-1; a run that starts at slot 0 mirrors that row with a copy of its own.  The bytes are
    // accounted on the slab's `full` barrier, so nobody waits for them but the consumers.
    auto fetch_rows = [&](int a, int b, unsigned long long* bar) {
      if (lane == 0) {
        const int n = b - a + 1;
        const int slot = a % R;
        const int n1 = min(n, R - slot);   // rows up to the end of the ring
        const bool wraps = n > n1;
        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.bes) + (size_t)a * rb;
        const unsigned bytes1 = (unsigned)(n1 + (wraps ? 1 : 0)) * rb;
        const unsigned bytes2 = wraps ? (unsigned)(n - n1) * rb : 0u;
        mbar_expect_tx(bar, bytes1 + bytes2 + (slot == 0 ? rb : 0));
        bulk_g2s(ring + (size_t)slot * rb, g, bytes1, bar);
        if (wraps) bulk_g2s(ring, g + (size_t)n1 * rb, bytes2, bar);
        if (slot == 0) bulk_g2s(ring + (size_t)R * rb, g, rb, bar);
        if (COUNT && p.ring_stats) st_rows += n;
      }
    };
    int rlo = 0x7fffffff;      // lowest resident row (identical in every producer thread by construction)
    int released = -1;         // slabs <= released have been released by every consumer warp
    int jlo[NU], jhi[NU];      // active multipole slots of this thread's pairs (both only ever move down)
    const unsigned* wrow = s_wtab + m_qi * NJP;
    // thresholds of the next moves: n2 of slot jhi, n1 of slot jlo - 1 (reloaded only when the slot moves)
    int w_hi[NU], w_lo[NU];
#pragma unroll
    for (int u = 0; u < NU; u++) {
      jlo[u] = NJ; jhi[u] = NJ - 1;
      w_hi[u] = (int)(wrow[NJ - 1] >> 16); w_lo[u] = (int)(wrow[NJ - 1] & 0xffffu);
    }
    const float rinvR = 1.0f / (float)R;
    // ring duty for slab t (ring warp): wait until the rows of slab t may overwrite their slots, then fetch them
    auto ring_step = [&](int t, bool wait_own) {
      const int par = t % W4_NST;
      // the barrier phase of slab t - NST must be over before anything is signalled on it again
      if (t >= W4_NST) {
        if (wait_own) mbar_wait(s_bar + W4_NST + par, ((t - W4_NST) / W4_NST) & 1);
        released = max(released, t - W4_NST);
      }
      CK4(ck_c);
      stage_sources(t + W4_SPD);
      const int2 w = s_win[t];
      // The windows slide towards lower rows (x falls with tau).  Rows [w.x, ...] may overwrite ring slots only
      // if the highest row still needed by the oldest slab not yet released stays within R rows of w.x.
      while (released + 1 < t && s_win[released + 1].y - w.x + 1 > R) {
        const int r = released + 1;
        mbar_wait(s_bar + W4_NST + (r % W4_NST), (r / W4_NST) & 1);
        released = r;
        if (COUNT && p.ring_stats && lane == 0) st_late++;
      }
      CK4(ck_c);
      const int f_hi = (rlo <= w.y) ? rlo - 1 : w.y;  // rows >= rlo are resident (fetched for earlier slabs)
      if (w.x <= f_hi) fetch_rows(w.x, f_hi, s_bar + par);
      rlo = min(rlo, w.x);
      if (COUNT && p.ring_stats && lane == 0) st_slabs++;
      CK4(ck_b);
    };
    if (ring_warp && !W4_PBAL) {
#pragma unroll 1
      for (int t = 0; t < W4_SPD; t++) stage_sources(t);
      for (int t = 0; t < nslab; t++) {
        ring_step(t, true);
        mbar_arrive(s_bar + (t % W4_NST));
      }
    }
    if (ring_warp && W4_PBAL) {
#pragma unroll 1
      for (int t = 0; t < W4_SPD; t++) stage_sources(t);
    }
    const bool m_role = W4_PBAL || !ring_warp;   // this warp computes metadata
    if (m_role && m_grp * MS < nslab) prefetch();
    CK4(ck_a);
    for (int t = m_grp * MS; m_role && t < nslab; t += W4_MG * MS) {
      // metadata buffers are free once the consumers have released the slabs that used them NST slabs ago
#pragma unroll
      for (int us = 0; us < MS; us++) {
        const int tu = t + us;
        if (tu < nslab && tu >= W4_NST) mbar_wait(s_bar + W4_NST + (tu % W4_NST), ((tu - W4_NST) / W4_NST) & 1);
      }
      CK4(ck_c);
      if (W4_PBAL && ring_warp) ring_step(t, false);
      // ---- metadata of this thread's pairs ----
#pragma unroll
      for (int u = 0; u < NU; u++) {
        const int tu = t + u / PPT;
        const int n = n_lo + tu * S + m_nn + (u % PPT) * PNN;
        const int pidx = m_pair + (u % PPT) * QC * PNN;
        const int par = tu % W4_NST;
        Proj4Rec* m_rec = reinterpret_cast<Proj4Rec*>(meta_base + (size_t)par * W4_MB);
        int moff = 0, jr = 127;
        double ma = 0;
        // slots still inside their window at time sample n: n2 >= n from above, n1 <= n from below
        while (jhi[u] >= 0 && w_hi[u] < n) { jhi[u]--; w_hi[u] = (int)(wrow[max(jhi[u], 0)] >> 16); }
        while (jlo[u] > 0 && w_lo[u] <= n) { jlo[u]--; w_lo[u] = (int)(wrow[max(jlo[u] - 1, 0)] & 0xffffu); }
#if CB200_W4_UNSAFE_FREEMETA
        if (f_valid[u]) {  // timing experiment only (WRONG results): the consumers' floor with metadata that costs nothing
          ma = 0.5;
          moff = ((n * 7 + m_qi * 3) % R) * rb;
          if (jlo[u] <= jhi[u]) jr = jlo[u] | ((jhi[u] + 1) << 8);
        }
        if (false) {
#else
        if (f_valid[u]) {
#endif
          const double x = fabs(__dmul_rn(pc.q, __dsub_rn(tau0, f_tau[u])));
          double x0, x1, inv_h;
          int rs;
          int bi = w4_locate(p, x, x0, x1, inv_h, rs);
          if (bi > p.num_xx - 1) { bi = p.num_xx - 1; x0 = p.bx[bi - 1]; x1 = p.bx[bi]; inv_h = 1.0 / (x1 - x0); rs = p.bseg.n - 1; }
          // interpolation weight (a value, not an index): reciprocal multiply instead of the reference's division.  The
          // cubic term's (x1 - x0)^2 / 6 is a constant of the stretch: the consumers take it from p.h2o6[rs]
          ma = (x1 - x) * inv_h;
          // ring slot of row bi - 1 without an integer division: float estimate of the quotient, corrected by one
          int sl = (bi - 1) - R * (int)((float)(bi - 1) * rinvR);
          sl += (sl < 0) ? R : 0;
          sl -= (sl >= R) ? R : 0;
          moff = sl * rb + rs;   // rb is a multiple of 128: the stretch rides in the low bits
          if (jlo[u] <= jhi[u]) jr = jlo[u] | ((jhi[u] + 1) << 8);
        }
        if (m_live && tu < nslab) {
          Proj4Rec rec; rec.a = ma; rec.off = moff; rec.jr = jr;
          m_rec[pidx] = rec;
          // the three sources at this wavenumber and time sample, weighted by dtau
          double o[3] = {0.0, 0.0, 0.0};
          const int stg_i = tu % W4_NSR;
          if (!r_wide) mbar_wait(s_sbar + stg_i, (tu / W4_NSR) & 1);
          if (f_valid[u] && !CB200_W4_UNSAFE_FREEMETA) {
            if (!r_wide) {
              const double* stg = reinterpret_cast<const double*>(sraw + (size_t)stg_i * W4_SRAW_STAGE) + ((m_nn + (u % PPT) * PNN) * 3) * W4_KSP + i_ko;
#pragma unroll
              for (int sI = 0; sI < 3; sI++) {
                const double* Sp = stg + sI * W4_KSP;
                const double* Dp = Sp + S * 3 * W4_KSP;
                o[sI] = (i_a0 * Sp[0] + i_b0 * Sp[1] + (i_a3 * Dp[0] + i_b3 * Dp[1])) * f_dtau[u];
              }
            } else {
#pragma unroll
              for (int sI = 0; sI < 3; sI++) {
                const double* Sp = src + ((size_t)(n - 1) * v.NSRC + sI) * v.NK + (pc.klo - 1);
                const double* Dp = dds + ((size_t)(n - 1) * v.NSRC + sI) * v.NK + (pc.klo - 1);
                o[sI] = (i_a0 * __ldg(Sp) + i_b0 * __ldg(Sp + 1) + (i_a3 * __ldg(Dp) + i_b3 * __ldg(Dp + 1))) * f_dtau[u];
              }
            }
          }
#if CB200_W4_UNSAFE_FREEMETA
          o[0] = o[1] = o[2] = 1.0;
#endif
          reinterpret_cast<double2*>(meta_base + (size_t)par * W4_MB + NPAIR * 16)[pidx] = make_double2(o[0], o[1]);
          reinterpret_cast<double*>(meta_base + (size_t)par * W4_MB + NPAIR * 32)[pidx] = o[2];
        }
      }
      // publish first, then issue the global loads of the next slab (their issue would only delay the consumers)
#pragma unroll
      for (int us = 0; us < MS; us++)
        if (t + us < nslab) mbar_arrive(s_bar + ((t + us) % W4_NST));
      CK4(ck_c);
      if (t + W4_MG * MS < nslab) prefetch();
      CK4(ck_a);
    }
    if (COUNT && p.ring_stats) {
      if (ring_warp && lane == 0) {
        atomicAdd(p.ring_stats + 0, st_slabs);
        atomicAdd(p.ring_stats + 2, st_rows); atomicAdd(p.ring_stats + 3, st_late);
      }
      if (lane == 0) {  // producer: metadata + prefetch, ring issue, waits
        atomicAdd(p.ring_stats + 5, (unsigned long long)ck_a); atomicAdd(p.ring_stats + 7, (unsigned long long)ck_b);
        atomicAdd(p.ring_stats + 9, (unsigned long long)ck_c);
      }
    }
    return;
  }

  // ============================================= CONSUMER =============================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 " CB200_STR(CB200_W4_CREG) ";\n");
  double acc[LKH][2], acc2[K2];
#pragma unroll
  for (int k = 0; k < LKH; k++) acc[k][0] = acc[k][1] = 0.0;
#pragma unroll
  for (int k = 0; k < K2; k++) acc2[k] = 0.0;
  const unsigned ring_lane = smem_u32(ring) + li * 16 + (W4_OSPL ? op * 128 : 0);  // shared-space address
  const int lc = li + (W4_OSPL ? 8 * op : 0);
  constexpr int OSH = W4_OSPL ? 4 : 3, ORND = W4_OSPL ? 15 : 7, OSTR = W4_OSPL ? 256 : 128;  // l-slot stride 16 or 8 per k
  CK4(ck_a);
  for (int t = 0; t < nslab; t++) {
    const int par = t % W4_NST;
    const int n_base = n_lo + t * S;
    long long tw0 = 0;
    if (COUNT) tw0 = clock64();
    mbar_wait(s_bar + par, (t / W4_NST) & 1);
    if (COUNT && p.ring_stats && tid == 0) {
      const long long tw1 = clock64(), ta = (long long)s_bar[16 + par], tr = (long long)s_bar[24 + par];
      st_wait_all += tw1 - tw0;
      if (ta > tw0) st_wait_prod += ta - tw0;   // part of the wait spent before the last metadata thread arrived
      if (tr > tw0) st_wait_ring += tr - tw0;   // ... before the ring warp had issued this slab's copies
      s_bar[16 + par] = 0;
    }
    CK4(ck_b);
    const unsigned char* mb = meta_base + (size_t)par * W4_MB;
    const Proj4Rec* m_rec = reinterpret_cast<const Proj4Rec*>(mb);
    const double2* m_s01 = reinterpret_cast<const double2*>(mb + NPAIR * 16);
    const double* m_s2 = reinterpret_cast<const double*>(mb + NPAIR * 32);
#if CB200_W4_PAIR2
    {
      static_assert(!CB200_W4_PAIR2 || (W4_TS && !W4_OS && W4_S == 4), "PAIR2: time-split consumers, 4-sample slabs");
      // the warp's two time samples of this slab (lh, lh + 2) in ONE step: their metadata loads, weight chains and the warp-wide
      // OR of the masks overlap, and a batch is one octet of both pairs (4 node loads in flight, two independent chains).  The
      // sums are added in the order of the one-pair-per-step loop (sample lh first), so the results are bit-identical.
      const int pr0 = lh * QC + myqi, pr1 = pr0 + 2 * QC;
      const Proj4Rec rec0 = m_rec[pr0], rec1 = m_rec[pr1];
      const double2 sA = m_s01[pr0], sB = m_s01[pr1];
      const double s2A = m_s2[pr0], s2B = m_s2[pr1];
      auto mask_of = [&](const Proj4Rec& rec) {
        const int klo = max((int)((rec.jr & 0xff) + ORND - lc) >> OSH, 0);
        const int khi1 = ((rec.jr >> 8) + ORND - lc) >> OSH;
        return ((1u << khi1) - 1u) & ~((1u << klo) - 1u);
      };
      const unsigned mA = mask_of(rec0), mB = mask_of(rec1);
      if (COUNT) {
        unsigned mxA = 0, mxB = 0;
#pragma unroll
        for (int k = 0; k < LKH; k++) {
          const unsigned tA = (unsigned)(n_base + lh) - (win[k] & 0xffffu), tB = tA + 2u;
          mxA |= (tA <= (win[k] >> 16)) ? (1u << k) : 0u;
          mxB |= (tB <= (win[k] >> 16)) ? (1u << k) : 0u;
        }
        if (mxA != mA) st_mismatch++;
        if (mxB != mB) st_mismatch++;
        if (p.triples) my_triples += __popc(mA) + __popc(mB);
      }
      const double aA = rec0.a, bA = 1 - aA, tA2 = -(bA * (aA * p.h2o6[rec0.off & 7]));
      const double g0A = tA2 * (aA + 1), g1A = tA2 * (2 - aA);
      const double aB = rec1.a, bB = 1 - aB, tB2 = -(bB * (aB * p.h2o6[rec1.off & 7]));
      const double g0B = tB2 * (aB + 1), g1B = tB2 * (2 - aB);
      const unsigned rpA = ring_lane + (rec0.off & ~127), rpB = ring_lane + (rec1.off & ~127);
      const unsigned U = __reduce_or_sync(0xffffffffu, mA | mB);
#pragma unroll
      for (int k = 0; k < LKH; k++) {
        if (!((U >> k) & 1u)) continue;   // warp-uniform
        const double2 A0 = lds128(rpA + k * OSTR), A1 = lds128(rpA + k * OSTR + rb);
        const double2 B0 = lds128(rpB + k * OSTR), B1 = lds128(rpB + k * OSTR + rb);
        double JA = fma(g1A, A1.y, fma(g0A, A0.y, fma(bA, A1.x, aA * A0.x)));
        double JB = fma(g1B, B1.y, fma(g0B, B0.y, fma(bB, B1.x, aB * B0.x)));
        JA = ((mA >> k) & 1u) ? JA : 0.0;
        JB = ((mB >> k) & 1u) ? JB : 0.0;
        acc[k][0] = fma(sB.x, JB, fma(sA.x, JA, acc[k][0]));
        acc[k][1] = fma(sB.y, JB, fma(sA.y, JA, acc[k][1]));
        if (k < K2) acc2[k < K2 ? k : 0] = fma(s2B, JB, fma(s2A, JA, acc2[k < K2 ? k : 0]));
      }
    }
#else
    // quarter-warp r works on pair (q_r, n); every lane covers LK multipoles
#pragma unroll W4_UNROLL
    for (int nn = W4_TS ? lh : 0; nn < S; nn += W4_TS ? 2 : 1) {
      const int n = n_base + nn;
      const int pr = nn * QC + myqi;
      const Proj4Rec rec = m_rec[pr];
      const double2 s01 = m_s01[pr];
      const double s2v = m_s2[pr];
      // this lane's octets k with jlo <= lc + 16 k <= jhi (lc = li + 8 lh: the lane's first l-slot)
      const int klo = max((int)((rec.jr & 0xff) + ORND - lc) >> OSH, 0);
      const int khi1 = ((rec.jr >> 8) + ORND - lc) >> OSH;  // khi + 1 >= 0
      const unsigned m = ((1u << khi1) - 1u) & ~((1u << klo) - 1u);
      if (COUNT) {
        unsigned mx = 0;
#pragma unroll
        for (int k = 0; k < LKH; k++) {
          const unsigned tt = (unsigned)n - (win[k] & 0xffffu);
          mx |= (tt <= (win[k] >> 16)) ? (1u << k) : 0u;
        }
        if (mx != m) st_mismatch++;
      }
      // cubic-spline value of j_l between the two nodes (cmbmain.f90:1515-1516), weights expanded:
      //   J = a j0 + b j1 + g0 p0 + g1 p1,  b = 1-a, g0 = -b fac (a+1), g1 = -b fac (2-a)
      const double a2 = rec.a, b2 = 1 - a2, t2 = -(b2 * (a2 * p.h2o6[rec.off & 7]));
      const double g0 = t2 * (a2 + 1), g1 = t2 * (2 - a2);
      if (COUNT && p.triples) my_triples += __popc(m);
      // octets in batches of KB: the loads of a batch are in flight together, and a batch with no active lane
      // anywhere in the warp is skipped altogether (no loads, no FP64 issue)
      constexpr int KB = CB200_W4_KB;
      {
        unsigned rp = ring_lane + (rec.off & ~127);
        const unsigned U = __reduce_or_sync(0xffffffffu, m);  // octets with an active lane anywhere in the warp
        // octets [kf, kf + cnt) of a batch: loads first (in flight together), then the arithmetic
        auto octets = [&](const int kf, const int cnt) {
          double2 N0[KB], N1[KB];
#pragma unroll
          for (int kk = 0; kk < KB; kk++) {
            const int k = kf + kk;
            if (kk < cnt && k < LKH) {
              const bool act = (m >> k) & 1u;
#if CB200_W4_PREDLOAD
              N0[kk] = lds128_if(rp + k * OSTR, act);
              N1[kk] = lds128_if(rp + k * OSTR + rb, act);
#elif CB200_W4_UNSAFE_ABLATE == 1   // timing experiment only (WRONG results): no table loads
              (void)act;
              N0[kk] = make_double2(a2, g0);
              N1[kk] = make_double2(b2, g1);
#else
              (void)act;
              N0[kk] = lds128(rp + k * OSTR);
              N1[kk] = lds128(rp + k * OSTR + rb);
#endif
            }
          }
#pragma unroll
          for (int kk = 0; kk < KB; kk++) {
            const int k = kf + kk;
            if (kk < cnt && k < LKH) {
              const bool act = (m >> k) & 1u;
#if CB200_W4_UNSAFE_ABLATE == 2     // timing experiment only (WRONG results): loads kept, one FP64 add per load instead of 7 FMAs
              (void)act;
              acc[k][0] += N0[kk].x + N1[kk].y;
              acc[k][1] += N0[kk].y + N1[kk].x;
#else
              double Jv = fma(g1, N1[kk].y, fma(g0, N0[kk].y, fma(b2, N1[kk].x, a2 * N0[kk].x)));
              Jv = act ? Jv : 0.0;
              acc[k][0] = fma(s01.x, Jv, acc[k][0]);
              acc[k][1] = fma(s01.y, Jv, acc[k][1]);
              if (k < K2) acc2[k < K2 ? k : 0] = fma(s2v, Jv, acc2[k < K2 ? k : 0]);
#endif
            }
          }
        };
#pragma unroll
        for (int k0 = 0; k0 < LKH; k0 += KB) {
          const unsigned ub = (U >> k0) & ((1u << KB) - 1u);
          if (!ub) continue;  // warp-uniform: no loads, no FP64 issue
#if CB200_W4_OCTSKIP
          if (KB == 2 && k0 + 1 < LKH) {
            // a run of active multipoles ends inside half of the batches it touches: the idle octet loads nothing
            if (ub == 3u) octets(k0, 2);
            else if (ub == 1u) octets(k0, 1);
            else octets(k0 + 1, 1);
          } else {
            octets(k0, KB);
          }
#else
          octets(k0, KB);
#endif
        }
      }
    }
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(s_bar + W4_NST + par);
    CK4(ck_c);
  }

  if (COUNT && p.triples) {
    unsigned long long t = my_triples;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(p.triples, t);
  }
  if (COUNT && p.ring_stats && st_mismatch) atomicAdd(p.ring_stats + 1, st_mismatch);
  if (COUNT && p.ring_stats && tid == 0) { atomicAdd(p.ring_stats + 10, (unsigned long long)st_wait_all); atomicAdd(p.ring_stats + 11, (unsigned long long)st_wait_prod); atomicAdd(p.ring_stats + 12, (unsigned long long)st_wait_ring); }
  if (COUNT && p.ring_stats && lane == 0) {  // consumer: prologue, wait for the producer, accumulate
    atomicAdd(p.ring_stats + 4, (unsigned long long)ck_a); atomicAdd(p.ring_stats + 6, (unsigned long long)ck_b);
    atomicAdd(p.ring_stats + 8, (unsigned long long)ck_c);
  }
#undef CK4
#if CB200_W4_SPLIT_EPI
  // ---- the two time halves of a wavenumber group add their sums (through the now free ring), and the half that owns an octet
  //      writes its three time integrals to global memory.  The Limber values and the partial k-contraction - ~25 k clocks of
  //      dependent global loads, divisions and CTA-wide barriers at the end of every CTA (12 % of the consumers' time in the
  //      ncu source page, with nothing else resident on the SM) - run in project4_finish_kernel at full occupancy.
  static_assert(!CB200_W4_SPLIT_EPI || (W4_TS && !W4_OS), "the split epilogue is written for the time-split consumers");
  asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W4_NCW) : "memory");
  {
    unsigned e_tid, e_bx, e_by;
    asm volatile("mov.u32 %0, %%tid.x;\n" : "=r"(e_tid));
    asm volatile("mov.u32 %0, %%ctaid.x;\n" : "=r"(e_bx));
    asm volatile("mov.u32 %0, %%ctaid.y;\n" : "=r"(e_by));
    const int tid = (int)e_tid, lane = tid & 31, warp = tid >> 5;
    const int wg = warp % NQG, lh = (warp / NQG) & 1;
    constexpr int NE = 3 * LKH;
    static_assert(32 * 1024 + (size_t)NQG * NE * 32 * 8 <= 128 * 1024, "exchange scratch inside the ring");
    double* xs = reinterpret_cast<double*>(smem_raw + 32 * 1024) + (size_t)wg * NE * 32 + lane;  // [group][entry][lane]
#pragma unroll
    for (int k = 0; k < LKH; k++) {
      if ((k < KS) == (lh == 1)) {   // not mine to finish
        xs[(3 * k) * 32] = acc[k][0];
        xs[(3 * k + 1) * 32] = acc[k][1];
        if (k < K2) xs[(3 * k + 2) * 32] = acc2[k < K2 ? k : 0];
      }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W4_NCW) : "memory");
    double* rw = p.raw + ((((size_t)e_by * p.NQB + e_bx) * NQG + wg) * NE) * 32 + lane;
#pragma unroll
    for (int k = 0; k < LKH; k++) {
      if ((k < KS) == (lh == 0)) {
        rw[(3 * k) * 32] = acc[k][0] + xs[(3 * k) * 32];
        rw[(3 * k + 1) * 32] = acc[k][1] + xs[(3 * k + 1) * 32];
        if (k < K2) rw[(3 * k + 2) * 32] = acc2[k < K2 ? k : 0] + xs[(3 * k + 2) * 32];
      }
    }
  }
#else
  // all consumer warps are done with the ring: it is reused for the contraction partials
  asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W4_NCW) : "memory");

  // The block's context is re-derived here from laundered special-register reads, so that none of it has to stay
  // in registers across the time loop (the loop needs every register for accumulators and loads in flight).
  {
  unsigned e_tid, e_bx, e_by;
  asm volatile("mov.u32 %0, %%tid.x;\n" : "=r"(e_tid));
  asm volatile("mov.u32 %0, %%ctaid.x;\n" : "=r"(e_bx));
  asm volatile("mov.u32 %0, %%ctaid.y;\n" : "=r"(e_by));
  const int tid = (int)e_tid, lane = tid & 31, warp = tid >> 5;
  const int lp = (int)e_by, pt = p.p0 + lp, qb = (int)e_bx, q0 = qb * QC;
  const int qr = lane >> 3, li = lane & 7, wg = warp % NQG, lh = (warp / NQG) & 1;
  const int op = W4_OS ? (warp / (2 * NQG)) & 1 : (W4_TS ? 0 : lh);
  const double tau0 = v.thermo[(size_t)pt * 5];
  const double* tau = v.tau + (size_t)pt * v.NT;
  const LinSegs& tseg = v.tseg[pt];
  const size_t row_stride = (size_t)v.NK;
  const size_t tau_stride = (size_t)v.NSRC * v.NK;
  const double* src = v.src + (size_t)pt * v.NT * tau_stride;
  const double* dds = p.ddsrc + (size_t)lp * v.NT * tau_stride;
  const int noct = (p.nl + 7) >> 3;
  const int myqi = wg * 4 + qr;
  const ProjQ3& myq = reinterpret_cast<const ProjQ3*>(smem_raw + (size_t)(p.R + 1) * rb + W4_META_BYTES)[myqi];
  // ---- Limber value of the lensing source (cmbmain.f90:1546-1556) and the partial k-contraction ----
  double* red = reinterpret_cast<double*>(smem_raw);  // [wavenumber group][6][PROJ_LP]
  if (W4_TS) {
    // the two time halves of a wavenumber group meet: each hands the sums of the octets the OTHER half finishes over
    // through the (now free) ring, so the epilogue (Limber values, partial k-contraction) is shared between them
    constexpr int NE = 3 * LKH;
    constexpr int NGX = NQG * (W4_OS ? 2 : 1);  // warp pairs that exchange sums
    static_assert((size_t)NQG * 6 * PROJ_LP * 8 <= 32 * 1024 && 32 * 1024 + (size_t)NGX * NE * 32 * 8 <= 128 * 1024, "epilogue scratch inside the ring");
    double* xs = reinterpret_cast<double*>(smem_raw + 32 * 1024) + (size_t)(W4_OS ? 2 * wg + op : wg) * NE * 32 + lane;  // [pair][entry][lane]
#pragma unroll
    for (int k = 0; k < LKH; k++) {
      if ((k < KS) == (lh == 1)) {   // not mine to finish
        xs[(3 * k) * 32] = acc[k][0];
        xs[(3 * k + 1) * 32] = acc[k][1];
        if (k < K2) xs[(3 * k + 2) * 32] = acc2[k < K2 ? k : 0];
      }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W4_NCW) : "memory");
#pragma unroll
    for (int k = 0; k < LKH; k++) {
      if ((k < KS) == (lh == 0)) {
        acc[k][0] += xs[(3 * k) * 32];
        acc[k][1] += xs[(3 * k + 1) * 32];
        if (k < K2) acc2[k < K2 ? k : 0] += xs[(3 * k + 2) * 32];
      }
    }
  }
  {
#pragma unroll
  for (int k = 0; k < LKH; k++) {
    if (W4_TS && (k < KS) != (lh == 0)) continue;  // finished by the other half
    double cl[6];
#pragma unroll
    for (int X = 0; X < 6; X++) cl[X] = 0.0;
    const int j = li + 8 * W4_OCT(k);
    const int l = (j < p.nl) ? p.ls[j] : 0;
    if (W4_OCT(k) >= LK) continue;  // octet outside the row (odd LK)
    double d2 = (k < K2) ? acc2[k < K2 ? k : 0] : 0.0;
    if (myq.valid && W4_OCT(k) < noct) {
      if (!p.tensors && j < p.nl && ((reached >> k) & 1u)) {
        const bool use_limber = l > 400;
        if (!((doint >> k) & 1u) || use_limber) {
          double xf = __dsub_rn(tau0, __ddiv_rn((double)l + 0.5, myq.q));
          double s3 = 0;
          if (xf < tseg.highest && xf > tau[0]) {
            const int n = lin_index_of(tseg, xf);
            xf = __ddiv_rn(__dsub_rn(xf, tau[n - 1]), __dsub_rn(tau[n], tau[n - 1]));
            double sa2 = 0, sb2 = 0;
            const double* S2p = src + 2 * row_stride + (myq.klo - 1);
            const double* D2p = dds + 2 * row_stride + (myq.klo - 1);
            if (n >= 2 && n <= myq.steps) {
              const double* a = S2p + (size_t)(n - 1) * tau_stride;
              const double* d = D2p + (size_t)(n - 1) * tau_stride;
              sa2 = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
            }
            if (n + 1 >= 2 && n + 1 <= myq.steps) {
              const double* a = S2p + (size_t)n * tau_stride;
              const double* d = D2p + (size_t)n * tau_stride;
              sb2 = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
            }
            s3 = (sa2 * (1 - xf) + xf * sb2) * sqrt(kPi / 2 / ((double)l + 0.5)) / myq.q;
          }
          d2 = s3;
        }
      }
      const double d0 = acc[k][0], d1 = acc[k][1];
      if (p.delta) {
        double* dp = p.delta + (((size_t)lp * v.NQ + (q0 + myqi)) * PROJ_LP + j) * 3;
        dp[0] = d0; dp[1] = d1; dp[2] = d2;
      }
      const double w = myq.w;
      if (p.tensors) {
        cl[0] = w * d0 * d0; cl[1] = w * d1 * d1; cl[2] = w * d2 * d2; cl[3] = w * d0 * d1;
      } else {
        cl[0] = w * d0 * d0; cl[1] = w * d1 * d1; cl[2] = w * d0 * d1;
        cl[3] = w * d2 * d2; cl[4] = w * d2 * d0; cl[5] = w * d2 * d1;
      }
    }
    // sum over the warp's four wavenumbers (quarters); warps are added below in a fixed order
#pragma unroll
    for (int X = 0; X < 6; X++) {
      double s = cl[X];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (qr == 0) red[((size_t)wg * 6 + X) * PROJ_LP + j] = s;
    }
  }
  }
  asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W4_NCW) : "memory");
  {
    double* pp = p.part + (((size_t)lp * p.NQB + qb) * 6) * PROJ_LP;
    for (int e = tid; e < 6 * PROJ_LP; e += 32 * NCW) {
      const int X = e / PROJ_LP, j = e - X * PROJ_LP;
      double s = 0;
      if (j < 8 * LK) {
#pragma unroll
        for (int w = 0; w < NQG; w++) s += red[((size_t)w * 6 + X) * PROJ_LP + j];
      }
      pp[e] = s;
    }
  }
  }
#endif
#undef W4_OCT
}

// Second half of the projection (split off the tail of project4_kernel): Limber value of the lensing source
// (cmbmain.f90:1546-1556), Delta_l(q) output, partial k-contraction of a wavenumber block (cmbmain.f90:2132-2264) in the
// fixed order the in-kernel epilogue used (the four wavenumbers of a warp by two shuffle steps, then the groups in order), so
// the C_l are bit-identical.  Same thread layout as the consumers: warp = (wavenumber group, half), quarter = wavenumber,
// lane = l-slot li + 8 k; half 0 finishes the octets below KS, half 1 the others.  The block's wavenumber constants and the
// reached / doint bits come from the projection kernel's prologue; the octets of a lane are independent chains (unrolled),
// and five such CTAs share an SM, so the dependent loads and divisions of the Limber branch overlap instead of stalling an SM.
template <int LK, int KLIM>
__global__ void __launch_bounds__(32 * W4_NCW, CB200_W4_FIN_MINB) project4_finish_kernel(const Proj4Params p) {
  constexpr int NCW = W4_NCW, NQG = W4_NQG, QC = W4_QC;
  constexpr int LKH = LK, K2 = KLIM, KS = (LKH + 1) / 2, NE = 3 * LKH;
  __shared__ double red[NQG * 6 * PROJ_LP];
  __shared__ double limfac[PROJ_LP];   // sqrt(pi / 2 / (l + 1/2)) of every l-slot: once per CTA instead of once per thread and octet
  const PointView& v = p.v;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lp = blockIdx.y, pt = p.p0 + lp, qb = blockIdx.x, q0 = qb * QC;
  const int nq = v.n_q[pt];
  if (q0 >= nq) return;
  if (p.fallback && p.fallback[(size_t)lp * p.NQB + qb]) return;   // this block was left to project3_kernel
  const int qr = lane >> 3, li = lane & 7, wg = warp % NQG, lh = (warp / NQG) & 1;
  const double tau0 = v.thermo[(size_t)pt * 5];
  const double* tau = v.tau + (size_t)pt * v.NT;
  const LinSegs& tseg = v.tseg[pt];
  const size_t row_stride = (size_t)v.NK;
  const size_t tau_stride = (size_t)v.NSRC * v.NK;
  const double* src = v.src + (size_t)pt * v.NT * tau_stride;
  const double* dds = p.ddsrc + (size_t)lp * v.NT * tau_stride;
  const int noct = (p.nl + 7) >> 3;
  const int myqi = wg * 4 + qr;
  const ProjQ3 myq = p.qcg[((size_t)lp * p.NQB + qb) * QC + myqi];
  const unsigned fl = p.flg[(((size_t)lp * p.NQB + qb) * NCW + warp) * 32 + lane];
  const double* rw = p.raw + ((((size_t)lp * p.NQB + qb) * NQG + wg) * NE) * 32 + lane;
  if (tid < PROJ_LP) limfac[tid] = tid < p.nl ? sqrt(kPi / 2 / ((double)p.ls[tid] + 0.5)) : 0.0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < LKH; k++) {
    if ((k < KS) != (lh == 0)) continue;   // finished by the other half (warp-uniform)
    double cl[6];
#pragma unroll
    for (int X = 0; X < 6; X++) cl[X] = 0.0;
    const int j = li + 8 * k;
    const bool lvalid = j < p.nl;
    const int l = lvalid ? p.ls[j] : 0;
    const double d0 = rw[(3 * k) * 32], d1 = rw[(3 * k + 1) * 32];
    double d2 = (k < K2) ? rw[(3 * k + 2) * 32] : 0.0;
    if (myq.valid && k < noct) {
      const bool reached = (fl >> k) & 1u, doint = (fl >> (16 + k)) & 1u;
      if (!p.tensors && lvalid && reached) {
        const bool use_limber = l > 400;
        if (!doint || use_limber) {
          double xf = __dsub_rn(tau0, __ddiv_rn((double)l + 0.5, myq.q));
          double s3 = 0;
          if (xf < tseg.highest && xf > tau[0]) {
            const int n = lin_index_of(tseg, xf);
            xf = __ddiv_rn(__dsub_rn(xf, tau[n - 1]), __dsub_rn(tau[n], tau[n - 1]));
            double sa2 = 0, sb2 = 0;
            const double* S2p = src + 2 * row_stride + (myq.klo - 1);
            const double* D2p = dds + 2 * row_stride + (myq.klo - 1);
            if (n >= 2 && n <= myq.steps) {
              const double* a = S2p + (size_t)(n - 1) * tau_stride;
              const double* d = D2p + (size_t)(n - 1) * tau_stride;
              sa2 = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
            }
            if (n + 1 >= 2 && n + 1 <= myq.steps) {
              const double* a = S2p + (size_t)n * tau_stride;
              const double* d = D2p + (size_t)n * tau_stride;
              sb2 = myq.a0 * a[0] + myq.b0 * a[1] + (myq.a03h * d[0] + myq.b03h * d[1]) * myq.ho2o6;
            }
            s3 = (sa2 * (1 - xf) + xf * sb2) * limfac[j] / myq.q;
          }
          d2 = s3;
        }
      }
      if (p.delta) {
        double* dp = p.delta + (((size_t)lp * v.NQ + (q0 + myqi)) * PROJ_LP + j) * 3;
        dp[0] = d0; dp[1] = d1; dp[2] = d2;
      }
      const double w = myq.w;
      if (p.tensors) {
        cl[0] = w * d0 * d0; cl[1] = w * d1 * d1; cl[2] = w * d2 * d2; cl[3] = w * d0 * d1;
      } else {
        cl[0] = w * d0 * d0; cl[1] = w * d1 * d1; cl[2] = w * d0 * d1;
        cl[3] = w * d2 * d2; cl[4] = w * d2 * d0; cl[5] = w * d2 * d1;
      }
    }
    // sum over the warp's four wavenumbers (quarters); the groups are added below in a fixed order
#pragma unroll
    for (int X = 0; X < 6; X++) {
      double sm = cl[X];
      sm += __shfl_xor_sync(0xffffffffu, sm, 8);
      sm += __shfl_xor_sync(0xffffffffu, sm, 16);
      if (qr == 0) red[((size_t)wg * 6 + X) * PROJ_LP + j] = sm;
    }
  }
  __syncthreads();
  double* pp = p.part + (((size_t)lp * p.NQB + qb) * 6) * PROJ_LP;
  for (int e = tid; e < 6 * PROJ_LP; e += 32 * NCW) {
    const int X = e / PROJ_LP, j = e - X * PROJ_LP;
    double sm = 0;
    if (j < 8 * LK) {
#pragma unroll
      for (int w = 0; w < NQG; w++) sm += red[((size_t)w * 6 + X) * PROJ_LP + j];
    }
    pp[e] = sm;
  }
}

// host side: the template instances in use (octets per row, octets that keep the lensing-potential accumulator)
template <int LK, int KLIM>
inline void w4_launch(bool count, dim3 grid, size_t smem, cudaStream_t s, const Proj4Params& pp) {
  if (count) project4_kernel<LK, true, KLIM><<<grid, W4_NT, smem, s>>>(pp);
  else project4_kernel<LK, false, KLIM><<<grid, W4_NT, smem, s>>>(pp);
#if CB200_W4_SPLIT_EPI
  project4_finish_kernel<LK, KLIM><<<grid, 32 * W4_NCW, 0, s>>>(pp);
#endif
}
// doubles of Proj4Params::raw per (point, wavenumber block)
inline size_t w4_raw_doubles(int LK) { return (size_t)W4_NQG * 3 * LK * 32; }
template <int LK, int KLIM>
inline cudaError_t w4_set_smem_attr() {
  cudaError_t e = cudaFuncSetAttribute(project4_kernel<LK, false, KLIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, W4_SMEM_TOTAL);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(project4_kernel<LK, true, KLIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, W4_SMEM_TOTAL);
}
constexpr int W4_KLIM11 = W4_TS ? 6 : 11;  // 88-multipole scalar set: octets 6..10 are l >= 700 (Limber for the lensing source)

}  // namespace cb200
