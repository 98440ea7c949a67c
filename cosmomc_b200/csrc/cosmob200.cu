// cosmob200 — handle, orchestration and the extern "C" boundary declared in include/cosmob200.h.
// Everything numerical runs in the hand-written sm_100a kernels of project.cuh / lens.cuh / like.cuh /
// dgemm.cuh / bessel.cuh; there is no CPU fallback: without a CUDA device cb200_create fails.
#include "../../include/cosmob200.h"
#include "common.cuh"
#include "grids.hpp"
#include "bessel.cuh"
#include "dgemm.cuh"
#include "project.cuh"
#ifdef CB200_TEST_KERNELS   // superseded generations of K1, only in libcosmob200_test.so (cross-checks of the parity tests)
#include "project1_test.cuh"
#include "project2_test.cuh"
#endif
#include "project3.cuh"
#include "project4.cuh"
#include "lens.cuh"
#include "like.cuh"
#include "background.cuh"
#include "sn.cuh"
#include "bk.cuh"
#include "thermo.cuh"
#include "nonlin.cuh"

#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>

using namespace cb200;

namespace {

constexpr int kTmplL = 8001;

struct KindSet {  // per perturbation type (0 scalar, 1 tensor): multipole set + Bessel table
  bool active = false;
  int max_l = 0;
  double max_eta_k = 0;
  std::vector<int> ls;
  SampleGrid bgrid;
  LinSegs bseg{};
  int num_xx = 0;
  DevBuf<int> d_ls;
  DevBuf<double> d_bx;
  DevBuf<double2> d_bes;
  DevBuf<double2> d_bes3;  // [chunk][row][32] layout for the windowed projection
  DevBuf<double2> d_bes4;  // [row][8 * lk4]: rows packed at the ring's row stride, so that a run of rows is ONE bulk copy (kernel 4)
  int lk4 = 0;             // octets per packed row
  DevBuf<int> d_llo;  // [max_l+1]
};

struct PointStore {  // resident per-point inputs
  int cap = 0, NT = 0, NK = 0, NQ = 0;
  DevBuf<double> thermo, tau, dtau, ksrc, q, dq, src;
  DevBuf<int> n_tau, n_k, n_q;
  DevBuf<LinSegs> tseg;
  std::vector<int> h_nq, h_ntau;  // host mirrors for launch shapes
  PointView view() const {
    PointView v;
    v.NT = NT; v.NK = NK; v.NQ = NQ; v.NSRC = 3;
    v.thermo = thermo.p; v.n_tau = n_tau.p; v.tau = tau.p; v.dtau = dtau.p; v.tseg = tseg.p;
    v.n_k = n_k.p; v.ksrc = ksrc.p; v.n_q = n_q.p; v.q = q.p; v.dq = dq.p; v.src = src.p;
    return v;
  }
};

struct LikeEntry {
  int type = 0;  // 1 plik-lite, 2 cmblikes, 3 BAO / MGS, 4 HST, 5 supernovae (JLA / Pantheon)
  int cal_index = -1;
  // background likelihoods: segment of the handle's redshift list
  int z_off = 0, nz = 0;
  BaoParams bao{};
  DevBuf<double> bao_invcov, bao_prob;
  double hst_H0 = 0, hst_err = 1, hst_zeff = 0, hst_ang = 0;
  // supernovae
  SnData sn{};
  DevBuf<double> sn_cols, sn_A1, sn_A2, sn_cov[6], sn_vinv;
  int sn_ia = -1, sn_ib = -1, sn_ld = 0, sn_nr = 0;
  double sn_sum_vinv = 0;
  // plik-lite
  int nused = 0, lmax_w = 0;
  DevBuf<int> bin_spec, bin_lo, bin_hi;
  DevBuf<double> weights, x_data, invcov;
  // cmblikes
  int nmaps = 0, ncl = 0, nbins = 0, ncl_used = 0, like_approx = 2;
  double log_cal_prior = -1;
  DevBuf<double> Wt_cmb, Wt_pp, offset, noise, chat, sqrt_fid;
  DevBuf<int> cl_use;
  bool has_noise = false, has_sqrt_fid = false;
  // BK foreground model attached to a cmblikes entry (TBK_planck extends TCMBLikes)
  bool has_fg = false;
  BkParams bk{};
  DevBuf<double> bk_nu, bk_R, bk_dnu, bk_lnnu, bk_w, bk_fgW;   // bk_fgW: [l][nbins * ncl]
  DevBuf<int> bk_lrange;                         // [nbins][2] first / last multipole with a non-zero window
};

enum Phase { PH_SPLINE = 0, PH_PROJECT, PH_CONTRACT, PH_INTERP, PH_LENS, PH_LIKE, PH_BG, PH_COUNT };

}  // namespace

struct cb200_handle {
  cb200_config cfg{};
  cb200_info info{};
  std::string err;
  cudaStream_t stream = nullptr;
  // uploads run on their own stream so that the host->device copy of one block of points overlaps the kernels of
  // the previous one; ev_upload orders `stream` behind the latest upload, `busy` lists the point ranges that
  // in-flight cb200_powers calls still read (an upload into such a range waits for that call)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_upload = nullptr;
  bool upload_pending = false, async_upload = false;
  // results (C_l, derived, status) of cb200_powers go back on a third stream when option "async_results" is on: the
  // call returns with the copy in flight, cb200_sync / cb200_loglike_batch close it
  cudaStream_t d2h_stream = nullptr;
  bool async_results = false, d2h_pending = false;
  std::vector<cudaEvent_t> d2h_events;
  // pinned staging of the per-point grids built on the host (two sets: set i is rebuilt while set i-1 is in flight),
  // so that cb200_upload_sources never blocks on a pageable copy queued behind the previous block's source DMA
  struct UploadStage {
    unsigned char* pin = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; bool used = false;
  } stage[2];
  int stage_next = 0;
  // change mask of cb200_eval_batch: what each resident slot's spectra were computed from
  std::vector<double> ev_key;         // [max_points][12] initpower[10], alens, aphiphi
  std::vector<unsigned> ev_epoch;     // [max_points] source-upload epoch + 1 the spectra belong to (0 = none)
  std::vector<unsigned> src_epoch;    // [max_points] bumped by every upload into the slot
  long long ev_powers = 0, ev_reused = 0;
  DevBuf<double> w_packed;            // device landing buffer of a packed source block (cb200_upload_sources_packed)
  DevBuf<long long> w_packed_off;     // [npts] offset of each point's block in it (doubles)
  struct BusyRange { int kind, first, npts; cudaEvent_t done; };
  std::vector<BusyRange> busy;
  KindSet kind[2];
  PointStore store[2];
  bool have_templates = false;
  DevBuf<double> d_tmpl, d_highl;
  int n_highl = 0;
  // lensing tables
  LensGeom lg{};
  DevBuf<int> d_jidx, d_lj;
  DevBuf<double> d_A1, d_M, d_tab, d_apod;
  std::vector<int> lj;
  // chunk work buffers
  int chunk = 0, LS = 0, LL = 0, LST = 0;
  DevBuf<double> w_coef, w_ddsrc, w_part, w_icl, w_cl, w_cin, w_sc, w_corr, w_lcon, w_initpower, w_alens, w_aphi;
  DevBuf<double> w_delta, w_clt, w_delta_sh[2], w_d2[2], w_pw, r_icl_t, r_clt;
  bool keep_transfers = false;
  int last_chunk_p0 = 0, last_chunk_np = 0;
  DevBuf<unsigned long long> d_triples;
  bool count_triples = false, ring_stats = false;
  int proj_kernel = 4;
  int sn_preassemble = 1;  // V(alpha, beta) assembled by its own coalesced kernel ahead of the Cholesky
  int sn_chol_warps = 0;       // warps per CTA of sn_chol2_kernel (8: two CTAs per SM, 4: four, 0: by launch size)
  int sn_chunk = 1024;         // points per launch of the supernova kernels
  int sn_chol_kernel_gen = 2;  // 2: sn_chol2_kernel (A fragments from global, column panel double-buffered), 1: sn_chol_kernel
  int spline_kernel = 2;  // 1: one thread per row straight from global memory, 2: tiled through shared memory
  DevBuf<unsigned long long> d_ring_stats;
  DevBuf<ProjQ3> w_qcg;              // [chunk][NQB][QC] wavenumber constants of kernel 4's blocks
  DevBuf<unsigned> w_flg;            // [chunk][NQB][consumer warp][lane] reached / doint bits
  DevBuf<double> w_raw;              // [chunk][NQB of kernel 4][group][3 LK][32]: kernel 4's time integrals (split epilogue)
  DevBuf<unsigned char> w_fallback;  // [chunk][NQB] blocks left to the chunked kernel by project4_kernel
  DevBuf<CUtensorMap> d_tmaps[2];
  DevBuf<double> w_bk_shapes, w_bk_scal, w_bk_G;   // BK15 foregrounds in GEMM form: l-shapes, map scalings, band powers
  int bk_scalar_foregrounds = 0;                    // option "bk_scalar_foregrounds": force the per-point scalar kernel
  DevBuf<double> w_nl_in, w_nl_work;             // non-linear lensing: inputs / power tables, ratios, sigma_8
  DevBuf<int> w_nl_status;
  DevBuf<double> w_th_work, w_th_in, w_th_out;   // thermal history: work tables [sample][point], input / result rows
  DevBuf<int> w_th_status;    // per perturbation type: TMA descriptors of {sources, second derivatives}
  // resident outputs
  DevBuf<double> r_cl_lensed, r_cls_out, r_derived, r_icl, r_cl;
  DevBuf<int> r_status;
  DevBuf<double> d_highl_norm;  // device scalar, 0 until the first evaluated point sets it (highl_norm_first_call)
  // likelihoods
  std::vector<std::unique_ptr<LikeEntry>> likes;
  DevBuf<double> w_resid, w_T, w_bc, w_bp, w_bigx, w_quad, w_ll, w_total, w_nuis, w_cls_in, w_binned;
  // background: massive-neutrino table, resident bg vectors, redshift list of the registered likelihoods
  DevBuf<double> d_nu_r1, d_nu_dr1, r_bg, d_bgz, w_da, w_hz, w_snW, w_bg_in, w_bg_z, w_bg_out;
  DevBuf<int> w_snbad;
  double nu_dlnam = 0;
  std::vector<double> bg_z;
  bool bg_z_dirty = false;
  int n_bg_likes = 0, n_cmb_likes = 0;
  BgTables bg_tables() const { return BgTables{d_nu_r1.p, d_nu_dr1.p, nu_dlnam}; }
  // timing
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[PH_COUNT];
  std::vector<cudaEvent_t> ev_pool;
  long long n_launches = 0;
  cudaEvent_t t_a = nullptr, t_b = nullptr;

  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    CB_CUDA(cudaEventCreate(&e));
    return e;
  }
  struct Scope {
    cb200_handle* h; int ph; cudaEvent_t a, b;
    Scope(cb200_handle* h_, int ph_) : h(h_), ph(ph_) {
      a = h->get_event(); b = h->get_event();
      cudaEventRecord(a, h->stream);
    }
    ~Scope() { cudaEventRecord(b, h->stream); h->ev[ph].push_back({a, b}); }
  };
};

namespace {

int fail(cb200_handle* h, const std::string& msg, int code = -1) {
  if (h) h->err = msg;
  return code;
}

#define CB_API_BEGIN try {
#define CB_API_END(h)                                               \
  }                                                                 \
  catch (const std::exception& e) { return fail((h), e.what()); }   \
  catch (...) { return fail((h), "unknown error"); }

LinSegs to_linsegs(const SampleGrid& g) {
  LinSegs s{};
  if (g.seg.size() > 8) throw std::runtime_error("grid has more than 8 segments");
  s.n = (int)g.seg.size();
  s.highest = g.highest;
  s.npoints = g.npoints;
  for (int i = 0; i < s.n; i++) {
    if (g.seg[i].is_log) throw std::runtime_error("log segment in a grid that must be linear");
    s.seg[i][0] = g.seg[i].lo; s.seg[i][1] = g.seg[i].hi; s.seg[i][2] = g.seg[i].step; s.seg[i][3] = g.seg[i].first;
    s.inv_step[i] = 1.0 / g.seg[i].step;
  }
  return s;
}

// lower sample index used by the running search of InterpolateClArr (camb/modules.f90:969-975)
std::vector<int> make_llo(const std::vector<int>& ls, int max_l) {
  std::vector<int> llo_of(max_l + 1, 1);
  int llo = 1, max_ind = (int)ls.size();
  for (int il = 2; il <= ls[max_ind - 1] && il <= max_l; il++) {
    if ((il > ls[llo]) && (llo < max_ind)) llo++;
    llo_of[il] = llo;
  }
  return llo_of;
}

void build_kind(cb200_handle* h, int k, int max_l, double max_eta_k) {
  KindSet& K = h->kind[k];
  K.active = true;
  K.max_l = max_l;
  K.max_eta_k = max_eta_k;
  K.ls = make_l_samples(max_l);
  if ((int)K.ls.size() > PROJ_LP) throw std::runtime_error("more than 96 sampled multipoles is not supported");
  make_bessel_x(K.bgrid, max_eta_k);
  K.bseg = to_linsegs(K.bgrid);
  K.num_xx = K.bgrid.npoints;
  K.d_ls.upload(K.ls, h->stream);
  K.d_bx.upload(K.bgrid.x, h->stream);
  K.d_bes.alloc((size_t)K.num_xx * PROJ_LP);
  std::vector<int> llo = make_llo(K.ls, max_l);
  K.d_llo.upload(llo, h->stream);
  DevBuf<double> scratch;
  scratch.alloc((size_t)K.ls.size() * K.num_xx);
  dim3 blk(64, 4), grd((K.num_xx + 63) / 64, (PROJ_LP + 3) / 4);
  bessel_values_kernel<<<grd, blk, 0, h->stream>>>(K.num_xx, (int)K.ls.size(), PROJ_LP, K.d_bx.p, K.d_ls.p, K.d_bes.p);
  CB_LAUNCH_CHECK();
  bessel_spline_kernel<<<((int)K.ls.size() + 31) / 32, 32, 0, h->stream>>>(K.num_xx, (int)K.ls.size(), PROJ_LP,
                                                                           K.d_bx.p, K.d_bes.p, scratch.p);
  CB_LAUNCH_CHECK();
  K.d_bes3.alloc((size_t)PROJ_LW * K.num_xx * 32);
  bessel_relayout_kernel<<<K.num_xx, PROJ_LP, 0, h->stream>>>(K.num_xx, K.d_bes.p, K.d_bes3.p);
  CB_LAUNCH_CHECK();
  {
    const int noct = ((int)K.ls.size() + 7) / 8;
    K.lk4 = w4_octets(noct);
    K.d_bes4.alloc((size_t)K.num_xx * 8 * K.lk4);
    CB_CUDA(cudaMemcpy2DAsync(K.d_bes4.p, (size_t)K.lk4 * 128, K.d_bes.p, sizeof(double2) * PROJ_LP, (size_t)K.lk4 * 128,
                              K.num_xx, cudaMemcpyDeviceToDevice, h->stream));
  }
  CB_CUDA(cudaStreamSynchronize(h->stream));
  h->n_launches += 3;
}

void build_lensing(cb200_handle* h) {
  // geometry of the correlation-function integrals (camb/lensing.f90:94-101,153-189)
  const KindSet& K = h->kind[0];
  LensGeom& g = h->lg;
  const int Max_l = K.max_l;
  int lmax_extrap = Max_l - 100 + 450 + 300;  // HighAccuracyDefault
  lmax_extrap = std::min(8000, lmax_extrap);
  g.max_l = Max_l;
  g.lmax = std::max(lmax_extrap, Max_l);
  int ix = (int)K.ls.size() - 1;
  while (K.ls[ix - 1] > Max_l - 100) ix--;
  g.lmax_lensed = K.ls[ix - 1];
  int npoints = (int)(Max_l * 2 * 1.0);
  double dtheta = kPi / npoints;
  if (Max_l > 3500) dtheta = dtheta / (double)1.3f;
  const int apw = round_half_away((double)0.003f / dtheta);
  npoints = (int)(kPi / dtheta);
  double range_fac = 1;
  const bool short_range = !h->cfg.accurate_bb;
  if (short_range) { range_fac = std::max(1., 32 / 1.0); npoints = (int)(npoints / range_fac); }
  g.npoints = npoints;
  g.dtheta = dtheta;
  g.interp_fac = std::max(1, std::min(round_half_away(10 / 1.0), (int)(range_fac * 2) - 1));
  g.NTH = npoints - 1;
  g.NTHP = ((g.NTH + 31) / 32) * 32;
  if (g.NTHP > 192 && short_range) throw std::runtime_error("unexpected theta count");
  std::vector<int> jidx(g.lmax + 1, -1);
  h->lj.clear();
  for (int l = 2; l <= g.lmax; l++)
    if (l <= 15 || ((l - 15) % g.interp_fac) == g.interp_fac / 2) { jidx[l] = (int)h->lj.size(); h->lj.push_back(l); }
  g.jmax = (int)h->lj.size();
  g.NLL = ((g.lmax_lensed - 1 + 3) / 4) * 4;
  std::vector<double> apod(g.NTHP, 0.0);
  for (int it = 0; it < g.NTH; it++) {
    const int i = it + 1;
    double w = 1;
    if (short_range && i > npoints - apw * 3) {
      const int d = i - npoints + apw * 3;
      // single-precision factor, as the reference's integer**2/real(...) expression (lensing.f90:436)
      w = (double)std::exp(-(float)(d * d) / (float)(2 * apw * apw));
    }
    apod[it] = w;
  }
  h->d_jidx.upload(jidx, h->stream);
  h->d_lj.upload(h->lj, h->stream);
  h->d_apod.upload(apod, h->stream);
  h->d_A1.alloc((size_t)(g.lmax - 1) * 2 * g.NTHP);
  h->d_M.alloc((size_t)4 * g.NTHP * g.NLL);
  h->d_tab.alloc((size_t)g.jmax * 12 * g.NTHP);
  h->d_A1.zero(h->stream); h->d_M.zero(h->stream); h->d_tab.zero(h->stream);
  lens_tables_kernel<<<(g.NTH + 31) / 32, 32, 0, h->stream>>>(g, h->d_jidx.p, h->d_A1.p, h->d_M.p, h->d_tab.p);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaStreamSynchronize(h->stream));
  h->n_launches += 1;
}

void parallel_for(int n, const std::function<void(int, int)>& fn) {
  int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
  nt = std::min(nt, std::max(1, n / 64));
  if (nt <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> errs(nt);
  for (int t = 0; t < nt; t++) {
    int a = (int)((long long)n * t / nt), b = (int)((long long)n * (t + 1) / nt);
    th.emplace_back([&, a, b, t]() {
      try { fn(a, b); } catch (...) { errs[t] = std::current_exception(); }
    });
  }
  for (auto& x : th) x.join();
  for (auto& e : errs) if (e) std::rethrow_exception(e);
}

void ensure_store(cb200_handle* h, int k) {
  PointStore& S = h->store[k];
  if (S.cap) return;
  S.cap = h->cfg.max_points;
  if (k == 0) { S.NT = h->cfg.n_tau_max; S.NK = h->cfg.n_k_max; S.NQ = h->cfg.n_q_max; }
  else { S.NT = h->cfg.n_tau_max_tensor; S.NK = h->cfg.n_k_max_tensor; S.NQ = h->cfg.n_q_max_tensor; }
  const size_t P = S.cap;
  S.thermo.alloc(P * 5); S.tau.alloc(P * S.NT); S.dtau.alloc(P * S.NT); S.ksrc.alloc(P * S.NK);
  S.q.alloc(P * S.NQ); S.dq.alloc(P * S.NQ); S.n_tau.alloc(P); S.n_k.alloc(P); S.n_q.alloc(P); S.tseg.alloc(P);
  S.src.alloc(P * S.NT * 3 * S.NK);
  S.h_nq.assign(P, 0); S.h_ntau.assign(P, 0);
}

void ensure_work(cb200_handle* h) {
  if (h->w_icl.p) return;
  const bool tens = h->cfg.compute_tensors != 0;
  const int C = h->chunk, NQ = std::max(h->cfg.n_q_max, tens ? h->cfg.n_q_max_tensor : 0);
  const int NK = std::max(h->cfg.n_k_max, tens ? h->cfg.n_k_max_tensor : 0);
  const size_t NTK = std::max((size_t)h->cfg.n_tau_max * h->cfg.n_k_max,
                              tens ? (size_t)h->cfg.n_tau_max_tensor * h->cfg.n_k_max_tensor : (size_t)0);
  const int NQB = (NQ + PROJ_Q - 1) / PROJ_Q;  // v1 needs the larger partial buffer
  h->w_coef.alloc((size_t)C * 5 * NK);
  h->w_ddsrc.alloc((size_t)C * NTK * 3);
  h->w_part.alloc((size_t)C * NQB * 6 * PROJ_LP);
  h->w_fallback.alloc((size_t)C * NQB);
  h->w_fallback.zero(h->stream);
#if CB200_W4_SPLIT_EPI
  // time integrals of kernel 4 on their way to project4_finish_kernel: 50 KB per (point, wavenumber block), 6.2 MB per point
  h->w_raw.alloc((size_t)C * ((NQ + W4_QC - 1) / W4_QC) * w4_raw_doubles(12));
  h->w_qcg.alloc((size_t)C * ((NQ + W4_QC - 1) / W4_QC) * W4_QC);
  h->w_flg.alloc((size_t)C * ((NQ + W4_QC - 1) / W4_QC) * W4_NCW * 32);
#endif
  if (tens) {
    h->w_clt.alloc((size_t)C * 4 * h->LST);
    h->r_icl_t.alloc((size_t)h->cfg.max_points * 6 * PROJ_LP);
    h->r_clt.alloc((size_t)h->cfg.max_points * 4 * h->LST);
  }
  h->w_icl.alloc((size_t)C * 6 * PROJ_LP);
  h->w_cl.alloc((size_t)C * 6 * h->LS);
  h->w_cin.alloc((size_t)C * 4 * h->LL);
  h->w_sc.alloc((size_t)C * 2 * h->lg.NTHP);
  h->w_corr.alloc((size_t)C * 4 * h->lg.NTHP);
  h->w_lcon.alloc((size_t)C * 4 * h->lg.NLL);
  h->w_initpower.alloc((size_t)C * 10);
  h->w_alens.alloc(C); h->w_aphi.alloc(C);
  const size_t P = h->cfg.max_points;
  h->r_cl_lensed.alloc(P * 4 * h->LS);
  h->r_cls_out.alloc(P * 5 * (h->cfg.lmax_out + 1));
  h->r_derived.alloc(P * 4);
  h->r_status.alloc(P);
  h->r_icl.alloc(P * 6 * PROJ_LP);
  h->r_cl.alloc(P * 6 * h->LS);
  h->d_triples.alloc(1);
  h->d_triples.zero(h->stream);
  h->d_ring_stats.alloc(16);
  h->d_ring_stats.zero(h->stream);
}

}  // namespace

extern "C" {

void cb200_default_config(cb200_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->struct_size = (int)sizeof(cb200_config);
  c->device = 0;
  c->lmax_computed_cl = 2500;
  c->cmb_lensing = 1;
  c->use_lensing_potential = 1;
  c->use_nonlinear_lensing = 1;
  c->compute_tensors = 0;
  c->lmax_tensor = 600;
  c->accurate_bb = 0;
  c->k_eta_max_scalar = -1;
  c->accuracy_level = 1;
  c->lmax_out = 2508;
  c->highl_norm_first_call = 0;
  c->max_points = 1024;
  c->chunk_points = 0;
  c->n_tau_max = 768; c->n_k_max = 256; c->n_q_max = 3072;
  c->n_tau_max_tensor = 2304; c->n_k_max_tensor = 128; c->n_q_max_tensor = 1024;
}

int cb200_config_size(void) { return (int)sizeof(cb200_config); }
int cb200_abi_version(void) { return CB200_VERSION; }

int cb200_create(const cb200_config* cfg, cb200_handle** out) {
  if (!cfg || !out) return -1;
  *out = nullptr;
  if (cfg->struct_size != (int)sizeof(cb200_config)) {
    std::fprintf(stderr, "cosmob200: cb200_config mirror out of step with include/cosmob200.h (caller %d bytes, library %d)\n",
                 cfg->struct_size, (int)sizeof(cb200_config));
    return -3;
  }
  std::unique_ptr<cb200_handle> h(new cb200_handle());
  h->cfg = *cfg;
  cb200_config& c = h->cfg;
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw std::runtime_error("cosmob200: no CUDA device available (there is no CPU fallback)");
    CB_CUDA(cudaSetDevice(c.device));
    CB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {  // the upload stream's unpack kernel and the download stream's copies must not queue behind the thousands of pending
       // CTAs of a projection launch on the compute stream: highest priority for both
      int lo = 0, hi = 0;
      CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CB_CUDA(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi));
      CB_CUDA(cudaStreamCreateWithPriority(&h->d2h_stream, cudaStreamNonBlocking, hi));
    }
    CB_CUDA(cudaEventCreateWithFlags(&h->ev_upload, cudaEventDisableTiming));
    if (c.accuracy_level != 1) throw std::runtime_error("only accuracy_level = 1 is supported");
    if (!c.cmb_lensing) throw std::runtime_error("only CMB_lensing = T is supported");
    if (c.n_tau_max <= 0) c.n_tau_max = 768;
    if (c.n_k_max <= 0) c.n_k_max = 256;
    if (c.n_q_max <= 0) c.n_q_max = 3072;
    if (c.n_tau_max_tensor <= 0) c.n_tau_max_tensor = 2304;
    if (c.n_k_max_tensor <= 0) c.n_k_max_tensor = 128;
    if (c.n_q_max_tensor <= 0) c.n_q_max_tensor = 1024;
    if (c.max_points <= 0) c.max_points = 1024;
    if (c.lmax_out <= 0) c.lmax_out = c.lmax_computed_cl;
    {  // massive-neutrino density table of the background functions (camb/modules.f90:1532-1610)
      std::vector<double> r1, dr1;
      build_nu_table(r1, dr1, h->nu_dlnam);
      h->d_nu_r1.upload(r1, h->stream);
      h->d_nu_dr1.upload(dr1, h->stream);
      CB_CUDA(cudaStreamSynchronize(h->stream));
      CB_CUDA(cudaFuncSetAttribute(sn_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
      CB_CUDA(cudaFuncSetAttribute(sn_chol2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Cfg<8>::SMEM));
      CB_CUDA(cudaFuncSetAttribute(sn_chol2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Cfg<4>::SMEM));
    }
    h->chunk = c.chunk_points > 0 ? c.chunk_points : std::min(c.max_points, 1024);
    h->chunk = std::min(h->chunk, c.max_points);
    c.chunk_points = h->chunk;
    h->info.max_points = c.max_points; h->info.chunk_points = h->chunk;
    if (c.lmax_computed_cl <= 0) {  // background-only handle (parameterization = background; no CMB theory)
      *out = h.release();
      return 0;
    }
    // CAMBCalc_InitCAMBParams (source/Calculator_CAMB.f90:749-754, 802-820)
    int max_l = c.lmax_computed_cl + 100 + 50;
    double max_eta_k = max_l * 2;
    max_eta_k = std::max(std::min(max_l, 3000) * 2.5 * c.accuracy_level, max_eta_k);
    if (c.use_lensing_potential || c.use_nonlinear_lensing) max_eta_k = std::max(max_eta_k, 14000 * c.accuracy_level);
    if (c.k_eta_max_scalar > 0) max_eta_k = c.k_eta_max_scalar;
    double max_eta_k_tensor = c.lmax_tensor * 5. / 2;
    if (c.compute_tensors) max_eta_k = std::max(max_eta_k, max_eta_k_tensor);  // CAMBParams_Set, modules.f90:285
    build_kind(h.get(), 0, max_l, max_eta_k);
    if (c.compute_tensors) build_kind(h.get(), 1, c.lmax_tensor, max_eta_k_tensor);
    build_lensing(h.get());
    if (std::min(c.lmax_computed_cl, c.lmax_out) > h->lg.lmax_lensed)
      throw std::runtime_error("lmax_computed_cl exceeds lmax_lensed");
    h->LS = ((max_l + 1 + 3) / 4) * 4;
    h->LL = ((h->lg.lmax + 1 + 3) / 4) * 4;
    h->LST = ((c.lmax_tensor + 1 + 3) / 4) * 4;
    cb200_info& I = h->info;
    I.max_l = max_l; I.max_eta_k = (int)max_eta_k; I.max_l_tensor = c.lmax_tensor; I.max_eta_k_tensor = (int)max_eta_k_tensor;
    I.n_lsamp = (int)h->kind[0].ls.size(); I.n_lsamp_tensor = (int)h->kind[1].ls.size();
    I.num_xx = h->kind[0].num_xx; I.num_xx_tensor = h->kind[1].num_xx; I.lmax_lensed = h->lg.lmax_lensed; I.lens_lmax = h->lg.lmax;
    I.lens_npoints = h->lg.npoints; I.lens_jmax = h->lg.jmax;
    I.n_tau_max = c.n_tau_max; I.n_k_max = c.n_k_max; I.n_q_max = c.n_q_max; I.max_points = c.max_points;
    I.chunk_points = h->chunk;
    const char* ct = std::getenv("CB200_COUNT_TRIPLES");
    h->count_triples = ct && ct[0] == '1';
#ifdef CB200_TEST_KERNELS
    CB_CUDA(cudaFuncSetAttribute(project_kernel<PROJ_Q, PROJ_NS, PROJ_SLAB>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CB_CUDA(cudaFuncSetAttribute(project2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM));
    CB_CUDA(cudaFuncSetAttribute(project2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM));
#endif
    CB_CUDA(cudaFuncSetAttribute(project3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
    CB_CUDA(cudaFuncSetAttribute(project3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
    CB_CUDA(cudaFuncSetAttribute(project3_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
    CB_CUDA(cudaFuncSetAttribute(project3_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
    CB_CUDA(cudaFuncSetAttribute(source_spline_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CB_CUDA((w4_set_smem_attr<6, 6>()));
    CB_CUDA((w4_set_smem_attr<11, 11>()));
    CB_CUDA((w4_set_smem_attr<11, W4_KLIM11>()));
    CB_CUDA((w4_set_smem_attr<12, 12>()));
    const char* pk = std::getenv("CB200_PROJ_KERNEL");
    if (pk && pk[0] >= '1' && pk[0] <= '4') h->proj_kernel = pk[0] - '0';
  } catch (const std::exception& e) {
    std::fprintf(stderr, "cb200_create: %s\n", e.what());
    return -1;
  }
  *out = h.release();
  return 0;
}

void cb200_destroy(cb200_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
  for (auto e : h->d2h_events) cudaEventDestroy(e);
  for (auto& st : h->stage) { if (st.pin) cudaFreeHost(st.pin); if (st.done) cudaEventDestroy(st.done); }
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->ev_upload) cudaEventDestroy(h->ev_upload);
  for (auto& b : h->busy) cudaEventDestroy(b.done);
  for (auto& v : h->ev) for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* cb200_last_error(const cb200_handle* h) { return h ? h->err.c_str() : "null handle"; }

int cb200_get_info(const cb200_handle* h, cb200_info* info) {
  if (!h || !info) return -1;
  *info = h->info;
  return 0;
}

int cb200_get_lsamples(const cb200_handle* h, int kind, int* l, int* n) {
  if (!h || kind < 0 || kind > 1 || !h->kind[kind].active) return -1;
  const auto& ls = h->kind[kind].ls;
  if (l) for (size_t i = 0; i < ls.size(); i++) l[i] = ls[i];
  if (n) *n = (int)ls.size();
  return 0;
}

int cb200_set_templates(cb200_handle* h, const double* unl, const double* lensed, int n_l) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  h->d_tmpl.upload(unl, (size_t)4 * kTmplL, h->stream);
  if (lensed && n_l > 0) { h->d_highl.upload(lensed, (size_t)4 * n_l, h->stream); h->n_highl = n_l; }
  CB_CUDA(cudaStreamSynchronize(h->stream));
  h->have_templates = true;
  return 0;
  CB_API_END(h)
}

int cb200_make_q_grid(const cb200_handle* h, int kind, double tau0, int max_n, double* q, double* dq, int* n) {
  if (!h || !h->kind[kind].active) return -1;
  try {
    SampleGrid g;
    make_q_grid(g, tau0, h->kind[kind].max_eta_k, h->kind[kind].max_l);
    *n = g.npoints;
    if (g.npoints > max_n) return -2;
    std::copy(g.x.begin(), g.x.end(), q);
    std::copy(g.dx.begin(), g.dx.end(), dq);
    return 0;
  } catch (...) { return -1; }
}

int cb200_make_time_steps(const cb200_handle* h, int kind, double tau0, double taurst, double taurend, double rs,
                          double rc, int max_n, double* tau, double* dtau, int* n) {
  if (!h || !h->kind[kind].active) return -1;
  try {
    SampleGrid g;
    make_time_steps(g, taurst, taurend, tau0, h->kind[kind].max_eta_k, kind == 1, rs, rc);
    *n = g.npoints;
    if (g.npoints > max_n) return -2;
    std::copy(g.x.begin(), g.x.end(), tau);
    std::copy(g.dx.begin(), g.dx.end(), dtau);
    return 0;
  } catch (...) { return -1; }
}

int cb200_make_source_k(const cb200_handle* h, int kind, double tau0, double taurst, int max_n, double* k, int* n) {
  if (!h || !h->kind[kind].active) return -1;
  try {
    SampleGrid g;
    make_source_k(g, tau0, taurst, h->kind[kind].max_eta_k, kind == 1, h->kind[kind].max_l);
    *n = g.npoints;
    if (g.npoints > max_n) return -2;
    std::copy(g.x.begin(), g.x.end(), k);
    return 0;
  } catch (...) { return -1; }
}

int cb200_grid_build(int nops, const double* ops, int max_n, double* x, double* dx, int* n, int n_query,
                     const double* query, int* index_out) {
  try {
    SampleGrid g;
    for (int i = 0; i < nops; i++) {
      const double* o = ops + 5 * i;
      if (o[0] == 0) g.add_spacing(o[1], o[2], o[3], o[4] != 0);
      else g.add_count(o[1], o[2], (int)o[3], o[4] != 0);
    }
    g.materialise(true);
    *n = g.npoints;
    if (g.npoints > max_n) return -2;
    std::copy(g.x.begin(), g.x.end(), x);
    std::copy(g.dx.begin(), g.dx.end(), dx);
    for (int i = 0; i < n_query; i++) index_out[i] = g.index_of(query[i]);
    return 0;
  } catch (...) { return -1; }
}

int cb200_get_bessel_table(const cb200_handle* hc, int kind, double* x, double* ajl, double* ajlpr) {
  cb200_handle* h = const_cast<cb200_handle*>(hc);
  if (!h || !h->kind[kind].active) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  const KindSet& K = h->kind[kind];
  std::vector<double2> tmp((size_t)K.num_xx * PROJ_LP);
  CB_CUDA(cudaMemcpy(tmp.data(), K.d_bes.p, tmp.size() * sizeof(double2), cudaMemcpyDeviceToHost));
  const int nl = (int)K.ls.size();
  for (int i = 0; i < K.num_xx; i++) {
    x[i] = K.bgrid.x[i];
    for (int j = 0; j < nl; j++) {
      ajl[(size_t)j * K.num_xx + i] = tmp[(size_t)i * PROJ_LP + j].x;
      ajlpr[(size_t)j * K.num_xx + i] = tmp[(size_t)i * PROJ_LP + j].y;
    }
  }
  return 0;
  CB_API_END(h)
}

}  // extern "C"

namespace {

// scatter of a packed source block [pt][n_tau][3][n_k] (exact sizes, as CAMB holds Src(k, s, tau)) into the padded
// resident layout [pt][NT][3][NK]; one warp per (point, time sample, source) row, coalesced both ways
__global__ void unpack_sources_kernel(int npts, int first, int NT, int NK, const int* __restrict__ n_tau,
                                      const int* __restrict__ n_k, const long long* __restrict__ off,
                                      const double* __restrict__ packed, double* __restrict__ src) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (tau, source) row of the point
  const int i = blockIdx.y, lane = threadIdx.x & 31;
  if (i >= npts || row >= NT * 3) return;
  const int nt = n_tau[first + i], nk = n_k[first + i];
  const int t = row / 3;
  double* dst = src + ((size_t)(first + i) * NT * 3 + row) * NK;
  if (t < nt) {
    const double* from = packed + off[i] + (size_t)row * nk;
    for (int k = lane; k < nk; k += 32) dst[k] = from[k];
    for (int k = nk + lane; k < NK; k += 32) dst[k] = 0.0;
  } else {
    for (int k = lane; k < NK; k += 32) dst[k] = 0.0;
  }
}

// src_mode 0: padded host buffer, 1: padded device buffer, 2: packed host buffer, 3: no sources (grids only)
int upload_impl(cb200_handle* h, int kind, int first, int npts, const double* thermo, const int* n_tau_in, const int* n_k,
                const double* k, const double* src, int src_mode) {
  if (kind < 0 || kind > 1 || !h->kind[kind].active) return fail(h, "upload_sources: kind not configured");
  if (first < 0 || npts <= 0 || first + npts > h->cfg.max_points) return fail(h, "upload_sources: point range exceeds max_points");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  ensure_store(h, kind);
  PointStore& S = h->store[kind];
  const KindSet& K = h->kind[kind];
  cudaStream_t s = h->copy_stream;
  // an in-flight cb200_powers call that still reads this point range has to finish first
  for (size_t i = 0; i < h->busy.size();) {
    cb200_handle::BusyRange& br = h->busy[i];
    if (cudaEventQuery(br.done) == cudaSuccess) { cudaEventDestroy(br.done); h->busy.erase(h->busy.begin() + i); continue; }
    if (br.kind == kind && br.first < first + npts && first < br.first + br.npts) CB_CUDA(cudaStreamWaitEvent(s, br.done, 0));
    i++;
  }
  // per-point staging layout (bytes), 16-byte aligned sections
  const size_t per_pt = sizeof(double) * (5 + 2 * (size_t)S.NT + 2 * (size_t)S.NQ + S.NK) + sizeof(int) * 4 +
                        sizeof(LinSegs) + sizeof(long long);
  const int SB = 512;   // points per staging pass
  const size_t per = (size_t)S.NT * 3 * S.NK;
  long long packed_done = 0;   // doubles of the caller's packed buffer consumed so far
  for (int b0 = 0; b0 < npts; b0 += SB) {
    const int nb = std::min(SB, npts - b0);
    cb200_handle::UploadStage& st = h->stage[h->stage_next];
    h->stage_next ^= 1;
    const size_t need = per_pt * SB + 256;
    if (st.bytes < need) {
      if (st.pin) { if (st.used) CB_CUDA(cudaEventSynchronize(st.done)); CB_CUDA(cudaFreeHost(st.pin)); }
      CB_CUDA(cudaHostAlloc((void**)&st.pin, need, cudaHostAllocDefault));
      st.bytes = need;
      if (!st.done) CB_CUDA(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
      st.used = false;
    }
    if (st.used) CB_CUDA(cudaEventSynchronize(st.done));   // the copies issued from this set two passes ago
    unsigned char* base = st.pin;
    auto carve = [&](size_t bytes) { unsigned char* r = base; base += (bytes + 15) / 16 * 16; return r; };
    double* p_thermo = (double*)carve(sizeof(double) * 5 * SB);
    double* p_tau = (double*)carve(sizeof(double) * (size_t)S.NT * SB);
    double* p_dtau = (double*)carve(sizeof(double) * (size_t)S.NT * SB);
    double* p_q = (double*)carve(sizeof(double) * (size_t)S.NQ * SB);
    double* p_dq = (double*)carve(sizeof(double) * (size_t)S.NQ * SB);
    double* p_k = (double*)carve(sizeof(double) * (size_t)S.NK * SB);
    int* p_ntau = (int*)carve(sizeof(int) * SB);
    int* p_nq = (int*)carve(sizeof(int) * SB);
    int* p_nk = (int*)carve(sizeof(int) * SB);
    LinSegs* p_tseg = (LinSegs*)carve(sizeof(LinSegs) * SB);
    long long* p_off = (long long*)carve(sizeof(long long) * SB);
    parallel_for(nb, [&](int a, int b) {
      SampleGrid gt, gq;
      for (int i = a; i < b; i++) {
        const double* th = thermo + (size_t)(b0 + i) * 5;
        make_time_steps(gt, th[1], th[2], th[0], K.max_eta_k, kind == 1, th[3], th[4]);
        make_q_grid(gq, th[0], K.max_eta_k, K.max_l);
        const int nki = n_k[b0 + i];
        if (gt.npoints > S.NT) throw std::runtime_error("n_tau exceeds n_tau_max");
        if (gq.npoints > S.NQ) throw std::runtime_error("n_q exceeds n_q_max");
        if (nki > S.NK || nki < 4) throw std::runtime_error("n_k out of range");
        if (n_tau_in && n_tau_in[b0 + i] != gt.npoints)
          throw std::runtime_error("packed sources: n_tau differs from the time-step grid the thermal-history scalars give");
        p_ntau[i] = gt.npoints; p_nq[i] = gq.npoints; p_nk[i] = nki;
        std::copy(th, th + 5, p_thermo + (size_t)i * 5);
        double* t = p_tau + (size_t)i * S.NT; double* dt = p_dtau + (size_t)i * S.NT;
        std::copy(gt.x.begin(), gt.x.end(), t); std::fill(t + gt.npoints, t + S.NT, 0.0);
        std::copy(gt.dx.begin(), gt.dx.end(), dt); std::fill(dt + gt.npoints, dt + S.NT, 0.0);
        double* qq = p_q + (size_t)i * S.NQ; double* dqq = p_dq + (size_t)i * S.NQ;
        std::copy(gq.x.begin(), gq.x.end(), qq); std::fill(qq + gq.npoints, qq + S.NQ, 0.0);
        std::copy(gq.dx.begin(), gq.dx.end(), dqq); std::fill(dqq + gq.npoints, dqq + S.NQ, 0.0);
        std::copy(k + (size_t)(b0 + i) * S.NK, k + (size_t)(b0 + i + 1) * S.NK, p_k + (size_t)i * S.NK);
        p_tseg[i] = to_linsegs(gt);
      }
    });
    const size_t f = (size_t)first + b0;
    CB_CUDA(cudaMemcpyAsync(S.thermo.p + f * 5, p_thermo, sizeof(double) * nb * 5, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.tau.p + f * S.NT, p_tau, sizeof(double) * (size_t)nb * S.NT, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.dtau.p + f * S.NT, p_dtau, sizeof(double) * (size_t)nb * S.NT, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.q.p + f * S.NQ, p_q, sizeof(double) * (size_t)nb * S.NQ, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.dq.p + f * S.NQ, p_dq, sizeof(double) * (size_t)nb * S.NQ, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.ksrc.p + f * S.NK, p_k, sizeof(double) * (size_t)nb * S.NK, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.n_tau.p + f, p_ntau, sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.n_q.p + f, p_nq, sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.n_k.p + f, p_nk, sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(S.tseg.p + f, p_tseg, sizeof(LinSegs) * nb, cudaMemcpyHostToDevice, s));
    for (int i = 0; i < nb; i++) { S.h_nq[f + i] = p_nq[i]; S.h_ntau[f + i] = p_ntau[i]; }
    if (h->src_epoch.size() < (size_t)h->cfg.max_points) h->src_epoch.assign(h->cfg.max_points, 0);
    for (int i = 0; i < nb; i++) h->src_epoch[f + i]++;
    if (src_mode == 0 || src_mode == 1) {
      CB_CUDA(cudaMemcpyAsync(S.src.p + f * per, src + (size_t)b0 * per, sizeof(double) * per * nb,
                              src_mode == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    } else if (src_mode == 2) {
      long long tot = 0;
      for (int i = 0; i < nb; i++) { p_off[i] = tot; tot += (long long)p_ntau[i] * 3 * p_nk[i]; }
      h->w_packed.alloc((size_t)SB * per);
      h->w_packed_off.alloc(SB);
      CB_CUDA(cudaMemcpyAsync(h->w_packed_off.p, p_off, sizeof(long long) * nb, cudaMemcpyHostToDevice, s));
      CB_CUDA(cudaMemcpyAsync(h->w_packed.p, src + packed_done, sizeof(double) * (size_t)tot, cudaMemcpyHostToDevice, s));
      dim3 grid((S.NT * 3 + 7) / 8, nb);
      unpack_sources_kernel<<<grid, 256, 0, s>>>(nb, (int)f, S.NT, S.NK, S.n_tau.p, S.n_k.p, h->w_packed_off.p,
                                                 h->w_packed.p, S.src.p);
      CB_LAUNCH_CHECK();
      h->n_launches += 1;
      packed_done += tot;
    }
    CB_CUDA(cudaEventRecord(st.done, s));
    st.used = true;
  }
  CB_CUDA(cudaEventRecord(h->ev_upload, s));
  h->upload_pending = true;
  // default: the caller may reuse its buffers as soon as this returns.  With option "async_upload" the call
  // returns with the source copy in flight (the buffer must stay untouched until cb200_sync or any call that
  // returns results); the next cb200_powers is ordered behind it on the device.
  if (!h->async_upload) CB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

}  // namespace

extern "C" {

int cb200_upload_sources(cb200_handle* h, int kind, int first, int npts, const double* thermo, const int* n_k,
                         const double* k, const double* src, int src_is_device) {
  if (!h) return -1;
  CB_API_BEGIN
  return upload_impl(h, kind, first, npts, thermo, nullptr, n_k, k, src, !src ? 3 : (src_is_device ? 1 : 0));
  CB_API_END(h)
}

int cb200_upload_sources_packed(cb200_handle* h, int kind, int first, int npts, const double* thermo, const int* n_tau,
                                const int* n_k, const double* k, const double* src_packed) {
  if (!h) return -1;
  CB_API_BEGIN
  if (!n_tau || !src_packed) return fail(h, "upload_sources_packed: n_tau and src_packed are required");
  return upload_impl(h, kind, first, npts, thermo, n_tau, n_k, k, src_packed, 2);
  CB_API_END(h)
}

int cb200_keep_transfers(cb200_handle* h, int on) {
  if (!h) return -1;
  h->keep_transfers = on != 0;
  return 0;
}

}  // extern "C"

namespace {

struct ProjLaunch { int q_per_block, nqb_total; };

// 2-D TMA descriptor of a [rows][NK] array of doubles with a box of (W4_S * 3 rows) x (W4_KSP wavenumbers): the raw
// source rows of one slab of the projection kernel.  cuTensorMapEncodeTiled is taken from the driver at run time.
CUtensorMap make_source_tensor_map(const double* base, size_t rows, int NK) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available in this driver");
    enc = reinterpret_cast<encode_fn>(fn);
  }
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)NK, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)NK * sizeof(double)};   // bytes between rows (a multiple of 16: NK is even)
  const cuuint32_t box[2] = {(cuuint32_t)W4_KSP, (cuuint32_t)(W4_S * 3)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return tm;
}

// K0 + K1 + K2 for one chunk of points of one perturbation type: resident sources -> sampled C_l in w_icl
// (source spline, line-of-sight projection fused with the partial k-contraction, fixed-order reduction + l-norms)
void project_chunk(cb200_handle* h, int kind, int p0, int np, bool have_alens, double* d_delta) {
  cudaStream_t s = h->stream;
  PointStore& S = h->store[kind];
  const KindSet& K = h->kind[kind];
  const int nl = (int)K.ls.size();
  PointView v = S.view();
  int nq_max = 0;
  for (int i = 0; i < np; i++) nq_max = std::max(nq_max, S.h_nq[p0 + i]);
  {  // K0
    cb200_handle::Scope sc(h, PH_SPLINE);
    spline_setup_kernel<<<(np + 63) / 64, 64, 0, s>>>(v, p0, np, h->w_coef.p);
    CB_LAUNCH_CHECK();
    const long long rows = (long long)np * S.NT * 3;
    if (h->spline_kernel == 2 && S.NK % 2 == 0) {
      const int tiles = (S.NT * 3 + SPL_ROWS - 1) / SPL_ROWS;
      const size_t smem = sizeof(double) * ((size_t)5 * S.NK + (size_t)SPL_ROWS * (S.NK + 1));
      source_spline_tiled_kernel<<<(unsigned)(tiles * np), SPL_THREADS, smem, s>>>(v, p0, np, h->w_coef.p, h->w_ddsrc.p);
    } else {
      source_spline_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(v, p0, np, h->w_coef.p, h->w_ddsrc.p);
    }
    CB_LAUNCH_CHECK();
    h->n_launches += 2;
  }
  ProjLaunch pl{PROJ_Q, (S.NQ + PROJ_Q - 1) / PROJ_Q};
  {  // K1
    cb200_handle::Scope sc(h, PH_PROJECT);
#ifdef CB200_TEST_KERNELS
    if (h->proj_kernel == 1) {
      ProjParams pp;
      pp.v = v; pp.p0 = p0; pp.nl = nl; pp.num_xx = K.num_xx; pp.NQB = pl.nqb_total; pp.tensors = kind;
      pp.max_eta_k = K.max_eta_k; pp.ddsrc = h->w_ddsrc.p; pp.bx = K.d_bx.p; pp.bes = K.d_bes.p;
      pp.initpower = h->w_initpower.p; pp.part = h->w_part.p;
      pp.delta = d_delta;
      pp.triples = h->count_triples ? h->d_triples.p : nullptr;
      pp.bseg = K.bseg;
      for (int i = 0; i < PROJ_LP; i++) pp.ls[i] = i < nl ? K.ls[i] : 0;
      constexpr size_t META = sizeof(ProjMeta) * PROJ_NS * PROJ_SLAB * PROJ_Q;
      constexpr size_t RED = sizeof(double) * (PROJ_NS - 1) * PROJ_Q * 3 * PROJ_LP;
      const size_t smem = std::max(META, RED) + sizeof(ProjQ) * PROJ_Q;
      dim3 grid((nq_max + PROJ_Q - 1) / PROJ_Q, np);
      project_kernel<PROJ_Q, PROJ_NS, PROJ_SLAB><<<grid, 32 * PROJ_LW * PROJ_NS, smem, s>>>(pp);
    } else
#endif
    if (h->proj_kernel == 4) {
      pl.q_per_block = W4_QC;
      pl.nqb_total = (S.NQ + W4_QC - 1) / W4_QC;
      Proj4Params pp;
      pp.v = v; pp.p0 = p0; pp.nl = nl; pp.num_xx = K.num_xx; pp.NQB = pl.nqb_total; pp.tensors = kind;
      static_assert(W3_QC >= W4_QC, "the fallback pass uses the same wavenumber blocks");
      const int noct = (nl + 7) / 8;
      const int LK = w4_octets(noct);
      if (LK != K.lk4) throw std::runtime_error("packed Bessel table built for another multipole count");
      pp.rb = LK * 128;
      pp.R = std::min(w4_ring_rows(LK, S.NT), 4096);
      pp.max_eta_k = K.max_eta_k; pp.ddsrc = h->w_ddsrc.p; pp.bx = K.d_bx.p; pp.bes = K.d_bes4.p;
      if (S.NK % 2) throw std::runtime_error("n_k_max must be even");
      pp.initpower = h->w_initpower.p; pp.part = h->w_part.p;
      pp.delta = d_delta;
      pp.triples = h->count_triples ? h->d_triples.p : nullptr;
      pp.ring_stats = h->ring_stats ? h->d_ring_stats.p : nullptr;
      pp.fallback = h->w_fallback.p;
      pp.raw = h->w_raw.p; pp.qcg = h->w_qcg.p; pp.flg = h->w_flg.p;
      pp.bseg = K.bseg;
      w4_set_last_stretch(pp);
      if (!h->d_tmaps[kind].p) {  // descriptors of this perturbation type's source arrays (fixed buffers: built once)
        CUtensorMap tm[2] = {make_source_tensor_map(S.src.p, (size_t)S.cap * S.NT * 3, S.NK),
                             make_source_tensor_map(h->w_ddsrc.p, (size_t)h->chunk * S.NT * 3, S.NK)};
        h->d_tmaps[kind].alloc(2);
        CB_CUDA(cudaMemcpyAsync(h->d_tmaps[kind].p, tm, sizeof(tm), cudaMemcpyHostToDevice, s));
        CB_CUDA(cudaStreamSynchronize(s));   // tm lives on this stack frame
      }
      pp.tmaps = h->d_tmaps[kind].p;
      for (int i = 0; i < PROJ_LP; i++) pp.ls[i] = i < nl ? K.ls[i] : 0;
      dim3 grid((nq_max + W4_QC - 1) / W4_QC, np);
      const bool cnt = h->count_triples || h->ring_stats;
      const size_t smem = w4_smem_bytes(LK, S.NT, pp.R);
      // octets from klim on only hold multipoles above 400: in a scalar run their lensing-potential sum is replaced by the
      // Limber value (cmbmain.f90:1546-1556), so the kernel instance without those accumulators can be used
      bool limber_tail = !kind && LK == 11 && W4_KLIM11 < 11;
      for (int j = 8 * W4_KLIM11; limber_tail && j < nl; j++) limber_tail = K.ls[j] > 400;
      if (LK == 6) w4_launch<6, 6>(cnt, grid, smem, s, pp);
      else if (LK == 11 && limber_tail) w4_launch<11, W4_KLIM11>(cnt, grid, smem, s, pp);
      else if (LK == 11) w4_launch<11, 11>(cnt, grid, smem, s, pp);
      else w4_launch<12, 12>(cnt, grid, smem, s, pp);
      CB_LAUNCH_CHECK();
#if CB200_W4_SPLIT_EPI
      h->n_launches += 1;   // project4_finish_kernel
#endif
      // fallback pass: blocks whose table window does not fit the ring (the first, log-spaced wavenumber block)
      Proj3Params p3;
      p3.v = v; p3.p0 = p0; p3.nl = nl; p3.num_xx = K.num_xx; p3.NQB = pl.nqb_total; p3.tensors = kind;
      p3.max_eta_k = K.max_eta_k; p3.ddsrc = h->w_ddsrc.p; p3.bx = K.d_bx.p; p3.bes3 = K.d_bes3.p;
      p3.initpower = h->w_initpower.p; p3.part = h->w_part.p;
      p3.delta = d_delta;
      p3.triples = h->count_triples ? h->d_triples.p : nullptr;
      p3.ring_stats = nullptr;
      p3.need = h->w_fallback.p;
      p3.qc_rt = W4_QC;
      p3.bseg = K.bseg;
      for (int i = 0; i < PROJ_LP; i++) p3.ls[i] = i < nl ? K.ls[i] : 0;
      dim3 grid3(1, (nl + 31) / 32, np);
      if (h->count_triples) project3_sweep_kernel<true><<<grid3, 32 * W3_NW, W3_SMEM, s>>>(p3);
      else project3_sweep_kernel<false><<<grid3, 32 * W3_NW, W3_SMEM, s>>>(p3);
      h->n_launches += 1;
    } else if (h->proj_kernel == 3) {
      pl.q_per_block = W3_QC;
      pl.nqb_total = (S.NQ + W3_QC - 1) / W3_QC;
      Proj3Params pp;
      pp.v = v; pp.p0 = p0; pp.nl = nl; pp.num_xx = K.num_xx; pp.NQB = pl.nqb_total; pp.tensors = kind;
      pp.max_eta_k = K.max_eta_k; pp.ddsrc = h->w_ddsrc.p; pp.bx = K.d_bx.p; pp.bes3 = K.d_bes3.p;
      pp.initpower = h->w_initpower.p; pp.part = h->w_part.p;
      pp.delta = d_delta;
      pp.triples = h->count_triples ? h->d_triples.p : nullptr;
      pp.ring_stats = h->ring_stats ? h->d_ring_stats.p : nullptr;
      pp.need = nullptr; pp.qc_rt = 0;
      pp.bseg = K.bseg;
      for (int i = 0; i < PROJ_LP; i++) pp.ls[i] = i < nl ? K.ls[i] : 0;
      dim3 grid((nq_max + W3_QC - 1) / W3_QC, (nl + 31) / 32, np);
      if (h->count_triples || h->ring_stats) project3_kernel<true><<<grid, 32 * W3_NW, W3_SMEM, s>>>(pp);
      else project3_kernel<false><<<grid, 32 * W3_NW, W3_SMEM, s>>>(pp);
    } else {
#ifdef CB200_TEST_KERNELS
      pl.q_per_block = W2_QC;
      pl.nqb_total = (S.NQ + W2_QC - 1) / W2_QC;
      Proj2Params pp;
      pp.v = v; pp.p0 = p0; pp.nl = nl; pp.num_xx = K.num_xx; pp.NQB2 = pl.nqb_total; pp.tensors = kind;
      pp.max_eta_k = K.max_eta_k; pp.ddsrc = h->w_ddsrc.p; pp.bx = K.d_bx.p; pp.bes3 = K.d_bes3.p;
      pp.initpower = h->w_initpower.p; pp.part = h->w_part.p;
      pp.delta = d_delta;
      pp.triples = h->count_triples ? h->d_triples.p : nullptr;
      pp.ring_stats = h->ring_stats ? h->d_ring_stats.p : nullptr;
      pp.bseg = K.bseg;
      for (int i = 0; i < PROJ_LP; i++) pp.ls[i] = i < nl ? K.ls[i] : 0;
      dim3 grid((nq_max + W2_QC - 1) / W2_QC, (nl + 31) / 32, np);
      if (h->count_triples || h->ring_stats) project2_kernel<true><<<grid, 32 * W2_NW, W2_SMEM, s>>>(pp);
      else project2_kernel<false><<<grid, 32 * W2_NW, W2_SMEM, s>>>(pp);
#else
      throw std::runtime_error("projection kernels 1 and 2 are test kernels: load libcosmob200_test.so");
#endif
    }
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  {  // K2
    cb200_handle::Scope sc(h, PH_CONTRACT);
    dim3 grid((6 * PROJ_LP + 127) / 128, np);
    contract_reduce_kernel<<<grid, 128, 0, s>>>(np, p0, S.n_q.p, pl.q_per_block, pl.nqb_total, nl, K.d_ls.p, kind,
                                               (have_alens && kind == 0) ? h->w_alens.p : nullptr, h->w_part.p,
                                               h->w_icl.p, 0);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
}

// K3: sampled C_l (w_icl) -> all-l spectra
void interp_chunk(cb200_handle* h, int kind, int np, double* d_cl, int LSk) {
  cb200_handle::Scope sc(h, PH_INTERP);
  const KindSet& K = h->kind[kind];
  InterpParams ip;
  ip.np = np; ip.nl = (int)K.ls.size(); ip.max_l = K.max_l; ip.LS = LSk; ip.nspec = kind ? 4 : 6;
  ip.templated = kind ? 0 : 1;
  ip.icl = h->w_icl.p; ip.ls = K.d_ls.p; ip.llo_of_l = K.d_llo.p; ip.tmpl = h->d_tmpl.p; ip.cl = d_cl;
  interp_cls_kernel<<<np, 192, 0, h->stream>>>(ip, PROJ_LP);
  CB_LAUNCH_CHECK();
  h->n_launches += 1;
}

// K4 + units: unlensed w_cl (+ optional tensor spectra) -> lensed C_l, CosmoMC-unit Cls, derived, status at slot p0
void lens_chunk(cb200_handle* h, int p0, int np, const double* d_cl_tensor, int tensor_shared, bool have_aphi) {
  cudaStream_t s = h->stream;
  const KindSet& K = h->kind[0];
  const LensGeom& g = h->lg;
  const int lmax_out = h->cfg.lmax_out;
  cb200_handle::Scope sc(h, PH_LENS);
  dim3 gp((g.lmax + 1 + 127) / 128, np);
  lens_prep_kernel<<<gp, 128, 0, s>>>(np, g, h->LS, h->LL, h->w_cl.p, h->d_tmpl.p, h->w_cin.p);
  CB_LAUNCH_CHECK();
  // (1) sigma^2 | Cg2 = Cphil3[2..lmax] x A1
  dgemm(s, false, false, np, 2 * g.NTHP, g.lmax - 1, 1.0, h->w_cin.p + 2, 4 * h->LL, h->d_A1.p, 2 * g.NTHP,
        h->w_sc.p, 2 * g.NTHP, &h->n_launches);
  LensCorrParams cp;
  cp.np = np; cp.LL = h->LL; cp.g = g; cp.sc = h->w_sc.p; cp.cin = h->w_cin.p; cp.tab = h->d_tab.p;
  cp.lj = h->d_lj.p; cp.apod = h->d_apod.p; cp.corr = h->w_corr.p;
  const size_t csm = sizeof(double) * LENS_PB * 3 * g.jmax;
  lens_corr_kernel<<<(np + LENS_PB - 1) / LENS_PB, g.NTHP, csm, s>>>(cp);
  CB_LAUNCH_CHECK();
  // (3) four correlation -> multipole transforms
  const size_t ms = (size_t)g.NTHP * g.NLL;
  for (int k4 = 0; k4 < 4; k4++)
    dgemm(s, false, false, np, g.lmax_lensed - 1, g.NTHP, 1.0, h->w_corr.p + (size_t)k4 * g.NTHP, 4 * g.NTHP,
          h->d_M.p + k4 * ms, g.NLL, h->w_lcon.p + (size_t)k4 * g.NLL, 4 * g.NLL, &h->n_launches);
  FinishParams fp;
  fp.np = np; fp.LS = h->LS; fp.NLL = g.NLL; fp.lmax_out = lmax_out; fp.lmax_computed_cl = h->cfg.lmax_computed_cl;
  fp.lmax_lensed = g.lmax_lensed; fp.n_highl = h->n_highl;
  fp.lmax_tensor = h->cfg.lmax_tensor; fp.have_tensor = d_cl_tensor ? 1 : 0; fp.LST = h->LST;
  fp.dtheta = g.dtheta; fp.cl = h->w_cl.p; fp.lcon = h->w_lcon.p; fp.cl_tensor = d_cl_tensor; fp.tensor_shared = tensor_shared;
  fp.highl = h->d_highl.p; fp.aphiphi = have_aphi ? h->w_aphi.p : nullptr;
  fp.cl_lensed = h->r_cl_lensed.p + (size_t)p0 * 4 * h->LS;
  fp.cls_out = h->r_cls_out.p + (size_t)p0 * 5 * (lmax_out + 1);
  fp.norm_dev = nullptr;
  if (h->cfg.highl_norm_first_call && h->n_highl > 0) {
    if (!h->d_highl_norm.p) { h->d_highl_norm.alloc(1); h->d_highl_norm.zero(s); }
    fp.norm_dev = h->d_highl_norm.p;
    highl_norm_kernel<<<1, 32, 0, s>>>(fp);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  const int lspan = std::max(h->LS, lmax_out + 1);
  dim3 gf((lspan + 127) / 128, np);
  lens_finish_kernel<<<gf, 128, 0, s>>>(fp);
  CB_LAUNCH_CHECK();
  derived_status_kernel<<<np, 32, 0, s>>>(np, h->LS, lmax_out, K.max_l, h->w_cl.p, fp.cls_out,
                                          h->r_derived.p + (size_t)p0 * 4, h->r_status.p + p0,
                                          h->cfg.compute_tensors ? h->w_initpower.p : nullptr);
  CB_LAUNCH_CHECK();
  h->n_launches += 4;
  // keep intermediates resident for parity read-backs
  CB_CUDA(cudaMemcpyAsync(h->r_cl.p + (size_t)p0 * 6 * h->LS, h->w_cl.p, sizeof(double) * np * 6 * h->LS,
                          cudaMemcpyDeviceToDevice, s));
}

// close the window of asynchronous result copies (option "async_results")
void finish_results(cb200_handle* h) {
  if (!h->d2h_pending) return;
  CB_CUDA(cudaStreamSynchronize(h->d2h_stream));
  for (auto e : h->d2h_events) cudaEventDestroy(e);
  h->d2h_events.clear();
  h->d2h_pending = false;
}

int powers_finish(cb200_handle* h, int first, int npts, double* cls_out, double* derived_out, int* status) {
  cudaStream_t s = h->stream;
  const int lmax_out = h->cfg.lmax_out;
  if (h->async_results && (cls_out || derived_out || status)) {
    // results travel on their own stream behind an event of the compute stream; nothing blocks the host
    cudaEvent_t e;
    CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->d2h_events.push_back(e);
    CB_CUDA(cudaEventRecord(e, s));
    cudaStream_t d = h->d2h_stream;
    CB_CUDA(cudaStreamWaitEvent(d, e, 0));
    if (cls_out)
      CB_CUDA(cudaMemcpyAsync(cls_out, h->r_cls_out.p + (size_t)first * 5 * (lmax_out + 1),
                              sizeof(double) * npts * 5 * (lmax_out + 1), cudaMemcpyDeviceToHost, d));
    if (derived_out)
      CB_CUDA(cudaMemcpyAsync(derived_out, h->r_derived.p + (size_t)first * 4, sizeof(double) * npts * 4,
                              cudaMemcpyDeviceToHost, d));
    if (status)
      CB_CUDA(cudaMemcpyAsync(status, h->r_status.p + first, sizeof(int) * npts, cudaMemcpyDeviceToHost, d));
    h->d2h_pending = true;
    return 0;
  }
  if (cls_out)
    CB_CUDA(cudaMemcpyAsync(cls_out, h->r_cls_out.p + (size_t)first * 5 * (lmax_out + 1),
                            sizeof(double) * npts * 5 * (lmax_out + 1), cudaMemcpyDeviceToHost, s));
  if (derived_out)
    CB_CUDA(cudaMemcpyAsync(derived_out, h->r_derived.p + (size_t)first * 4, sizeof(double) * npts * 4,
                            cudaMemcpyDeviceToHost, s));
  if (status)
    CB_CUDA(cudaMemcpyAsync(status, h->r_status.p + first, sizeof(int) * npts, cudaMemcpyDeviceToHost, s));
  if (cls_out || derived_out || status) CB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

void upload_chunk_params(cb200_handle* h, int c0, int np, const double* initpower, const double* alens,
                         const double* aphiphi) {
  cudaStream_t s = h->stream;
  CB_CUDA(cudaMemcpyAsync(h->w_initpower.p, initpower + (size_t)c0 * 10, sizeof(double) * 10 * np, cudaMemcpyHostToDevice, s));
  if (alens) CB_CUDA(cudaMemcpyAsync(h->w_alens.p, alens + c0, sizeof(double) * np, cudaMemcpyHostToDevice, s));
  if (aphiphi) CB_CUDA(cudaMemcpyAsync(h->w_aphi.p, aphiphi + c0, sizeof(double) * np, cudaMemcpyHostToDevice, s));
}

}  // namespace

extern "C" {

int cb200_powers(cb200_handle* h, int first, int npts, const double* initpower, const double* alens,
                 const double* aphiphi, double* cls_out, double* derived_out, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (!h->kind[0].active) return fail(h, "powers: background-only handle");
  if (!h->have_templates) return fail(h, "powers: call cb200_set_templates first");
  for (int i = std::max(first, 0); i < first + npts && i < (int)h->ev_epoch.size(); i++) h->ev_epoch[i] = 0;
  PointStore& S = h->store[0];
  if (!S.cap || first < 0 || npts <= 0 || first + npts > S.cap) return fail(h, "powers: point range not resident");
  const bool tens = h->cfg.compute_tensors != 0;
  if (tens && !h->store[1].cap) return fail(h, "powers: compute_tensors but no tensor sources uploaded");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  ensure_work(h);
  cudaStream_t s = h->stream;
  if (h->keep_transfers) h->w_delta.alloc((size_t)h->chunk * S.NQ * PROJ_LP * 3);
  if (h->upload_pending) { CB_CUDA(cudaStreamWaitEvent(s, h->ev_upload, 0)); h->upload_pending = false; }

  for (int c0 = 0; c0 < npts; c0 += h->chunk) {
    const int np = std::min(h->chunk, npts - c0);
    const int p0 = first + c0;
    h->last_chunk_p0 = p0; h->last_chunk_np = np;
    upload_chunk_params(h, c0, np, initpower, alens, aphiphi);
    project_chunk(h, 0, p0, np, alens != nullptr, h->keep_transfers ? h->w_delta.p : nullptr);
    CB_CUDA(cudaMemcpyAsync(h->r_icl.p + (size_t)p0 * 6 * PROJ_LP, h->w_icl.p, sizeof(double) * np * 6 * PROJ_LP,
                            cudaMemcpyDeviceToDevice, s));
    interp_chunk(h, 0, np, h->w_cl.p, h->LS);
    if (tens) {  // tensor pass of CAMB_GetResults (camb/camb.f90:106-237): same stages on the tensor sources
      project_chunk(h, 1, p0, np, false, nullptr);
      CB_CUDA(cudaMemcpyAsync(h->r_icl_t.p + (size_t)p0 * 6 * PROJ_LP, h->w_icl.p, sizeof(double) * np * 6 * PROJ_LP,
                              cudaMemcpyDeviceToDevice, s));
      interp_chunk(h, 1, np, h->w_clt.p, h->LST);
      CB_CUDA(cudaMemcpyAsync(h->r_clt.p + (size_t)p0 * 4 * h->LST, h->w_clt.p, sizeof(double) * np * 4 * h->LST,
                              cudaMemcpyDeviceToDevice, s));
    }
    lens_chunk(h, p0, np, tens ? h->w_clt.p : nullptr, 0, aphiphi != nullptr);
  }
  {  // the sources of [first, first + npts) are in use until here
    cb200_handle::BusyRange br{0, first, npts, nullptr};
    CB_CUDA(cudaEventCreateWithFlags(&br.done, cudaEventDisableTiming));
    CB_CUDA(cudaEventRecord(br.done, s));
    h->busy.push_back(br);
    if (tens) {
      cb200_handle::BusyRange bt{1, first, npts, nullptr};
      CB_CUDA(cudaEventCreateWithFlags(&bt.done, cudaEventDisableTiming));
      CB_CUDA(cudaEventRecord(bt.done, s));
      h->busy.push_back(bt);
    }
  }
  return powers_finish(h, first, npts, cls_out, derived_out, status);
  CB_API_END(h)
}

// Semi-slow step with SHARED transfer functions: one resident source point, many initial-power points
// (CosmoMC calls GetNewPowerData without GetNewTransferData when only the InitPower block moved,
// source/CalcLike_Cosmology.f90:73-85; BK15 chains with fixed cosmology, batch3/BK15only.ini).  The transfer
// functions Delta_l(q) are projected once; the k-contraction of the whole batch is then the dense product
//   iCl[pt][X][l] = sum_q  P(q; pt) dq/q  x  Delta_a Delta_b (q, l)          (FP64 tensor-pipe GEMM)
int cb200_powers_shared(cb200_handle* h, int src_point, int first, int npts, const double* initpower,
                        const double* alens, const double* aphiphi, double* cls_out, double* derived_out, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (!h->kind[0].active) return fail(h, "powers_shared: background-only handle");
  for (int i = std::max(first, 0); i < first + npts && i < (int)h->ev_epoch.size(); i++) h->ev_epoch[i] = 0;
  if (!h->have_templates) return fail(h, "powers_shared: call cb200_set_templates first");
  PointStore& S = h->store[0];
  if (!S.cap || src_point < 0 || src_point >= S.cap) return fail(h, "powers_shared: source point not resident");
  if (first < 0 || npts <= 0 || first + npts > h->cfg.max_points) return fail(h, "powers_shared: output range exceeds max_points");
  const bool tens = h->cfg.compute_tensors != 0;
  if (tens && !h->store[1].cap) return fail(h, "powers_shared: compute_tensors but no tensor sources uploaded");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  ensure_work(h);
  cudaStream_t s = h->stream;
  if (h->upload_pending) { CB_CUDA(cudaStreamWaitEvent(s, h->ev_upload, 0)); h->upload_pending = false; }
  // ---- transfer functions of the source point, once per perturbation type
  std::vector<double> ip1(10, 0.0);
  ip1[0] = 1; ip1[1] = 1; ip1[7] = 0.05; ip1[8] = 0.05;
  CB_CUDA(cudaMemcpyAsync(h->w_initpower.p, ip1.data(), sizeof(double) * 10, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaStreamSynchronize(s));
  const int nk = tens ? 2 : 1;
  for (int kind = 0; kind < nk; kind++) {
    PointStore& Sk = h->store[kind];
    const int nq = Sk.h_nq[src_point];
    h->w_delta_sh[kind].alloc((size_t)Sk.NQ * PROJ_LP * 3);
    h->w_d2[kind].alloc((size_t)Sk.NQ * 6 * PROJ_LP);
    project_chunk(h, kind, src_point, 1, false, h->w_delta_sh[kind].p);
    cb200_handle::Scope sc(h, PH_CONTRACT);
    delta_products_kernel<<<(nq * PROJ_LP + 127) / 128, 128, 0, s>>>(nq, kind, h->w_delta_sh[kind].p, h->w_d2[kind].p);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  for (int c0 = 0; c0 < npts; c0 += h->chunk) {
    const int np = std::min(h->chunk, npts - c0);
    const int p0 = first + c0;
    h->last_chunk_p0 = p0; h->last_chunk_np = np;
    upload_chunk_params(h, c0, np, initpower, alens, aphiphi);
    for (int kind = 0; kind < nk; kind++) {
      PointStore& Sk = h->store[kind];
      const KindSet& K = h->kind[kind];
      const int nq = Sk.h_nq[src_point];
      {
        cb200_handle::Scope sc(h, PH_CONTRACT);
        h->w_pw.alloc((size_t)h->chunk * Sk.NQ);
        dim3 gw((nq + 127) / 128, np);
        power_weights_kernel<<<gw, 128, 0, s>>>(np, nq, kind, Sk.q.p + (size_t)src_point * Sk.NQ,
                                                 Sk.dq.p + (size_t)src_point * Sk.NQ, h->w_initpower.p, h->w_pw.p, Sk.NQ);
        CB_LAUNCH_CHECK();
        dgemm(s, false, false, np, 6 * PROJ_LP, nq, 1.0, h->w_pw.p, Sk.NQ, h->w_d2[kind].p, 6 * PROJ_LP, h->w_part.p,
              6 * PROJ_LP, &h->n_launches);
        dim3 grid((6 * PROJ_LP + 127) / 128, np);
        contract_reduce_kernel<<<grid, 128, 0, s>>>(np, 0, Sk.n_q.p, 1 << 30, 1, (int)K.ls.size(), K.d_ls.p, kind,
                                                   (alens && kind == 0) ? h->w_alens.p : nullptr, h->w_part.p,
                                                   h->w_icl.p, nq);
        CB_LAUNCH_CHECK();
        h->n_launches += 2;
      }
      if (kind == 0) {
        CB_CUDA(cudaMemcpyAsync(h->r_icl.p + (size_t)p0 * 6 * PROJ_LP, h->w_icl.p, sizeof(double) * np * 6 * PROJ_LP,
                                cudaMemcpyDeviceToDevice, s));
        interp_chunk(h, 0, np, h->w_cl.p, h->LS);
      } else {
        CB_CUDA(cudaMemcpyAsync(h->r_icl_t.p + (size_t)p0 * 6 * PROJ_LP, h->w_icl.p, sizeof(double) * np * 6 * PROJ_LP,
                                cudaMemcpyDeviceToDevice, s));
        interp_chunk(h, 1, np, h->w_clt.p, h->LST);
        CB_CUDA(cudaMemcpyAsync(h->r_clt.p + (size_t)p0 * 4 * h->LST, h->w_clt.p, sizeof(double) * np * 4 * h->LST,
                                cudaMemcpyDeviceToDevice, s));
      }
    }
    lens_chunk(h, p0, np, tens ? h->w_clt.p : nullptr, 0, aphiphi != nullptr);
  }
  return powers_finish(h, first, npts, cls_out, derived_out, status);
  CB_API_END(h)
}

int cb200_debug_fetch(cb200_handle* h, int what, int point, int max_n, double* out, int* n) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  CB_CUDA(cudaStreamSynchronize(h->stream));
  const PointStore& S = h->store[0];
  const KindSet& K = h->kind[0];
  const int nl = (int)K.ls.size();
  auto fetch = [&](const double* d, size_t cnt) {
    if ((long long)cnt > max_n) throw std::runtime_error("debug_fetch: buffer too small");
    CB_CUDA(cudaMemcpy(out, d, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    *n = (int)cnt;
  };
  switch (what) {
    case 0: {
      std::vector<double> t((size_t)6 * PROJ_LP);
      CB_CUDA(cudaMemcpy(t.data(), h->r_icl.p + (size_t)point * 6 * PROJ_LP, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
      if (6 * nl > max_n) return fail(h, "debug_fetch: buffer too small");
      for (int X = 0; X < 6; X++) for (int j = 0; j < nl; j++) out[X * nl + j] = t[(size_t)X * PROJ_LP + j];
      *n = 6 * nl;
      break;
    }
    case 1: {
      std::vector<double> t((size_t)6 * h->LS);
      CB_CUDA(cudaMemcpy(t.data(), h->r_cl.p + (size_t)point * 6 * h->LS, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
      const int w = K.max_l + 1;
      if (6 * w > max_n) return fail(h, "debug_fetch: buffer too small");
      for (int X = 0; X < 6; X++) for (int l = 0; l < w; l++) out[X * w + l] = t[(size_t)X * h->LS + l];
      *n = 6 * w;
      break;
    }
    case 2: {
      std::vector<double> t((size_t)4 * h->LS);
      CB_CUDA(cudaMemcpy(t.data(), h->r_cl_lensed.p + (size_t)point * 4 * h->LS, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
      const int w = K.max_l + 1;
      if (4 * w > max_n) return fail(h, "debug_fetch: buffer too small");
      for (int X = 0; X < 4; X++) for (int l = 0; l < w; l++) out[X * w + l] = (l <= h->lg.lmax_lensed) ? t[(size_t)X * h->LS + l] : 0.0;
      *n = 4 * w;
      break;
    }
    case 3: {
      if (!h->keep_transfers || !h->w_delta.p) return fail(h, "debug_fetch: transfers not kept");
      // valid for points of the last processed chunk only
      const int local = point - h->last_chunk_p0;
      if (local < 0 || local >= h->last_chunk_np) return fail(h, "debug_fetch: point not in the last chunk");
      fetch(h->w_delta.p + (size_t)local * S.NQ * PROJ_LP * 3, (size_t)S.h_nq[point] * PROJ_LP * 3);
      break;
    }
    case 8: {  // tensor iCl [4][n_lsamp_tensor]
      if (!h->r_icl_t.p) return fail(h, "debug_fetch: no tensor spectra");
      const int nlt = (int)h->kind[1].ls.size();
      std::vector<double> t((size_t)6 * PROJ_LP);
      CB_CUDA(cudaMemcpy(t.data(), h->r_icl_t.p + (size_t)point * 6 * PROJ_LP, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
      if (4 * nlt > max_n) return fail(h, "debug_fetch: buffer too small");
      for (int X = 0; X < 4; X++) for (int j = 0; j < nlt; j++) out[X * nlt + j] = t[(size_t)X * PROJ_LP + j];
      *n = 4 * nlt;
      break;
    }
    case 9: {  // Cl_tensor [4][lmax_tensor+1] (dimensionless) TT, EE, BB, TE
      if (!h->r_clt.p) return fail(h, "debug_fetch: no tensor spectra");
      std::vector<double> t((size_t)4 * h->LST);
      CB_CUDA(cudaMemcpy(t.data(), h->r_clt.p + (size_t)point * 4 * h->LST, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
      const int w = h->cfg.lmax_tensor + 1;
      if (4 * w > max_n) return fail(h, "debug_fetch: buffer too small");
      for (int X = 0; X < 4; X++) for (int l = 0; l < w; l++) out[X * w + l] = t[(size_t)X * h->LST + l];
      *n = 4 * w;
      break;
    }
    case 4: fetch(S.q.p + (size_t)point * S.NQ, S.h_nq[point]); break;
    case 5: fetch(S.dq.p + (size_t)point * S.NQ, S.h_nq[point]); break;
    case 6: fetch(S.tau.p + (size_t)point * S.NT, S.h_ntau[point]); break;
    case 7: fetch(S.dtau.p + (size_t)point * S.NT, S.h_ntau[point]); break;
    case 10: {  // the resident lensing-potential source of the point, [n_tau][NK] (rows of source 3)
      const size_t nt = (size_t)S.h_ntau[point];
      if ((long long)(nt * S.NK) > max_n) return fail(h, "debug_fetch: buffer too small");
      CB_CUDA(cudaMemcpy2D(out, sizeof(double) * S.NK, S.src.p + ((size_t)point * S.NT * 3 + 2) * S.NK, sizeof(double) * 3 * S.NK,
                           sizeof(double) * S.NK, nt, cudaMemcpyDeviceToHost));
      *n = (int)(nt * S.NK);
      break;
    }
    default: return fail(h, "debug_fetch: unknown selector");
  }
  return 0;
  CB_API_END(h)
}

// ---------------------------------------------------------------------------------------------- likelihoods
int cb200_like_add_pliklite(cb200_handle* h, const int* nb, int nbins_tab, const int* blmin, const int* blmax,
                            const double* weights, int lmax_w, const double* invcov, const double* x_data,
                            int cal_index, int* like_id) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  std::unique_ptr<LikeEntry> L(new LikeEntry());
  L->type = 1; L->cal_index = cal_index; L->lmax_w = lmax_w;
  std::vector<int> spec, lo, hi;
  for (int i = 0; i < 3; i++) {
    if (nb[i] > nbins_tab) return fail(h, "pliklite: nb exceeds bin table");
    for (int j = 0; j < nb[i]; j++) {
      if (blmax[j] > std::min(lmax_w, h->cfg.lmax_out)) return fail(h, "pliklite: bin exceeds lmax_out / weights");
      spec.push_back(i); lo.push_back(blmin[j]); hi.push_back(blmax[j]);
    }
  }
  L->nused = (int)spec.size();
  L->bin_spec.upload(spec, h->stream); L->bin_lo.upload(lo, h->stream); L->bin_hi.upload(hi, h->stream);
  L->weights.upload(weights, (size_t)lmax_w + 1, h->stream);
  L->x_data.upload(x_data, L->nused, h->stream);
  L->invcov.upload(invcov, (size_t)L->nused * L->nused, h->stream);
  CB_CUDA(cudaStreamSynchronize(h->stream));
  if (like_id) *like_id = (int)h->likes.size();
  h->likes.push_back(std::move(L));
  h->n_cmb_likes++;
  return 0;
  CB_API_END(h)
}

int cb200_like_add_cmblikes(cb200_handle* h, int nmaps, int nbins, int ncl_used, const int* cl_use_index,
                            int like_approx, int lmax_w, const double* W, const double* offset, const double* noise,
                            const double* chat, const double* sqrt_fid, const double* invcov, double log_cal_prior,
                            int cal_index, int* like_id) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  if (nmaps > CMBL_MAXMAPS) return fail(h, "cmblikes: too many maps");
  if (lmax_w > h->cfg.lmax_out) return fail(h, "cmblikes: window lmax exceeds lmax_out");
  if (like_approx == 1 && !sqrt_fid) return fail(h, "cmblikes: HL needs sqrt_fid");
  std::unique_ptr<LikeEntry> L(new LikeEntry());
  L->type = 2; L->cal_index = cal_index; L->lmax_w = lmax_w; L->nmaps = nmaps; L->ncl = nmaps * (nmaps + 1) / 2;
  L->nbins = nbins; L->ncl_used = ncl_used; L->like_approx = like_approx; L->log_cal_prior = log_cal_prior;
  const int nb = nbins * L->ncl, LO = h->cfg.lmax_out + 1;
  // transpose the windows into GEMM B operands: Wt_cmb [(X,l)][nb] for X = TT,TE,EE,BB and Wt_pp [l][nb]
  std::vector<double> wc((size_t)4 * LO * nb, 0.0), wp((size_t)LO * nb, 0.0);
  for (int b = 0; b < nb; b++)
    for (int X = 0; X < 5; X++)
      for (int l = 0; l <= lmax_w; l++) {
        const double v = W[((size_t)b * 5 + X) * (lmax_w + 1) + l];
        if (X < 4) wc[((size_t)X * LO + l) * nb + b] = v;
        else wp[(size_t)l * nb + b] = v;
      }
  L->Wt_cmb.upload(wc, h->stream); L->Wt_pp.upload(wp, h->stream);
  L->offset.upload(offset, nb, h->stream);
  const size_t nm = (size_t)nbins * nmaps * nmaps;
  if (noise) { L->noise.upload(noise, nm, h->stream); L->has_noise = true; }
  L->chat.upload(chat, nm, h->stream);
  if (sqrt_fid) { L->sqrt_fid.upload(sqrt_fid, nm, h->stream); L->has_sqrt_fid = true; }
  L->cl_use.upload(cl_use_index, ncl_used, h->stream);
  const size_t nx = (size_t)nbins * ncl_used;
  L->invcov.upload(invcov, nx * nx, h->stream);
  CB_CUDA(cudaStreamSynchronize(h->stream));
  if (like_id) *like_id = (int)h->likes.size();
  h->likes.push_back(std::move(L));
  h->n_cmb_likes++;
  return 0;
  CB_API_END(h)
}

int cb200_like_set_bk_foregrounds(cb200_handle* h, int like_id, int nmaps, const int* map_field, const int* bc_class,
                                  const int* bp_offset, const double* bp_nu, const double* bp_R, const double* bp_dnu,
                                  const double* th_dust, const double* th_sync, const double* nu_bar, double fpivot_dust,
                                  double fpivot_sync, const double* fpivot_dust_decorr, const double* fpivot_sync_decorr,
                                  int lform_dust, int lform_sync, int lmin, int lmax, const double* fgW, int nuis_offset) {
  if (!h) return -1;
  CB_API_BEGIN
  if (like_id < 0 || like_id >= (int)h->likes.size() || h->likes[like_id]->type != 2)
    return fail(h, "bk_foregrounds: like_id is not a cmblikes likelihood");
  LikeEntry& L = *h->likes[like_id];
  if (nmaps != L.nmaps || nmaps > BK_MAXMAPS) return fail(h, "bk_foregrounds: map count mismatch");
  if (lmax + 1 > BK_MAXL || lmax != L.lmax_w || lmin < 1) return fail(h, "bk_foregrounds: multipole range not supported");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  BkParams& b = L.bk;
  b.nmaps = nmaps; b.nbins = L.nbins; b.ncl = L.ncl; b.lmin = lmin; b.lmax = lmax; b.nuis_off = nuis_offset;
  b.lform_dust = lform_dust; b.lform_sync = lform_sync;
  b.fpivot_dust = fpivot_dust; b.fpivot_sync = fpivot_sync;
  for (int k = 0; k < 2; k++) { b.fp_dust_decorr[k] = fpivot_dust_decorr[k]; b.fp_sync_decorr[k] = fpivot_sync_decorr[k]; }
  for (int i = 0; i < nmaps; i++) {
    b.field[i] = map_field[i]; b.bc_class[i] = bc_class[i]; b.bp_off[i] = bp_offset[i];
    b.th_dust[i] = th_dust[i]; b.th_sync[i] = th_sync[i]; b.nu_bar[i] = nu_bar[i];
  }
  b.bp_off[nmaps] = bp_offset[nmaps];
  const size_t nb = bp_offset[nmaps];
  L.bk_nu.upload(bp_nu, nb, h->stream); L.bk_R.upload(bp_R, nb, h->stream); L.bk_dnu.upload(bp_dnu, nb, h->stream);
  {  // ln(nu) and R dnu of every bandpass sample: the kernel's nu^(3+beta) becomes one exp
    std::vector<double> ln((size_t)nb), w((size_t)nb);
    for (int k = 0; k < nb; k++) { ln[k] = std::log(bp_nu[k]); w[k] = bp_dnu[k] * bp_R[k]; }
    L.bk_lnnu.upload(ln, h->stream); L.bk_w.upload(w, h->stream);
    CB_CUDA(cudaStreamSynchronize(h->stream));
  }
  {  // windows transposed to [l][bin * ncl] (coalesced over the kernel's threads) + the multipole support of every bin
    const int NTc = L.nbins * L.ncl, LW = lmax + 1;
    std::vector<double> wt((size_t)LW * NTc);
    std::vector<int> lr((size_t)L.nbins * 2);
    for (int b = 0; b < L.nbins; b++) {
      int lo = LW, hi = -1;
      for (int c = 0; c < L.ncl; c++)
        for (int l = 0; l < LW; l++) {
          const double w = fgW[((size_t)b * L.ncl + c) * LW + l];
          wt[(size_t)l * NTc + (size_t)b * L.ncl + c] = w;
          if (w != 0.0) { lo = std::min(lo, l); hi = std::max(hi, l); }
        }
      lr[2 * b] = std::max(lo, lmin); lr[2 * b + 1] = std::min(hi, lmax);
    }
    L.bk_fgW.upload(wt, h->stream);
    L.bk_lrange.upload(lr, h->stream);
    CB_CUDA(cudaStreamSynchronize(h->stream));
  }
  CB_CUDA(cudaStreamSynchronize(h->stream));
  L.has_fg = true;
  return 0;
  CB_API_END(h)
}

// ---------------------------------------------------------------------------------- background + its likelihoods
int cb200_set_background(cb200_handle* h, int first, int npts, const double* bg) {
  if (!h) return -1;
  CB_API_BEGIN
  if (first < 0 || npts <= 0 || first + npts > h->cfg.max_points) return fail(h, "set_background: point range exceeds max_points");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  h->r_bg.alloc((size_t)h->cfg.max_points * NBG);
  CB_CUDA(cudaMemcpyAsync(h->r_bg.p + (size_t)first * NBG, bg, sizeof(double) * npts * NBG, cudaMemcpyHostToDevice, h->stream));
  CB_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
  CB_API_END(h)
}

int cb200_background(cb200_handle* h, int npts, const double* bg, int nz, const double* z, double* DA, double* H,
                     double* scalars) {
  if (!h) return -1;
  CB_API_BEGIN
  if (npts <= 0 || nz < 0) return fail(h, "background: bad sizes");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  h->w_bg_in.alloc((size_t)npts * NBG);
  CB_CUDA(cudaMemcpyAsync(h->w_bg_in.p, bg, sizeof(double) * npts * NBG, cudaMemcpyHostToDevice, s));
  cb200_handle::Scope sc(h, PH_BG);
  if (nz > 0) {
    h->w_bg_z.alloc(nz);
    h->w_bg_out.alloc((size_t)2 * npts * nz);
    CB_CUDA(cudaMemcpyAsync(h->w_bg_z.p, z, sizeof(double) * nz, cudaMemcpyHostToDevice, s));
    const long long nthr = (long long)npts * nz;
    bg_distance_kernel<<<(unsigned)((nthr + 127) / 128), 128, 0, s>>>(npts, nz, h->w_bg_in.p, h->w_bg_z.p, h->bg_tables(),
                                                                     DA ? h->w_bg_out.p : nullptr,
                                                                     H ? h->w_bg_out.p + (size_t)npts * nz : nullptr);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
    if (DA) CB_CUDA(cudaMemcpyAsync(DA, h->w_bg_out.p, sizeof(double) * npts * nz, cudaMemcpyDeviceToHost, s));
    if (H) CB_CUDA(cudaMemcpyAsync(H, h->w_bg_out.p + (size_t)npts * nz, sizeof(double) * npts * nz, cudaMemcpyDeviceToHost, s));
  }
  if (scalars) {
    h->w_quad.alloc((size_t)npts * 3);
    bg_scalars_kernel<<<(npts + 31) / 32, 96, 0, s>>>(npts, h->w_bg_in.p, h->bg_tables(), h->w_quad.p);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
    CB_CUDA(cudaMemcpyAsync(scalars, h->w_quad.p, sizeof(double) * npts * 3, cudaMemcpyDeviceToHost, s));
  }
  CB_CUDA(cudaStreamSynchronize(s));
  return 0;
  CB_API_END(h)
}


// ---- non-linear lensing rescale and sigma_8 (nonlin.cuh) ----
int cb200_nonlinear_lensing(cb200_handle* h, int first, int npts, const double* initpower, const double* cosmo, int n_kt,
                            int n_z, const double* kh, const double* z, const double* transfer, const double* tautf,
                            int rescale_sources, double* sigma8, double* ratio, double* spec, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (npts <= 0 || n_kt < 4 || n_z < 2 || n_z > NL_MAXZ || !initpower || !cosmo || !kh || !z || !transfer)
    return fail(h, "nonlinear_lensing: bad arguments");
  if (rescale_sources && (!tautf || first < 0 || first + npts > h->cfg.max_points || h->cfg.lmax_computed_cl <= 0))
    return fail(h, "nonlinear_lensing: rescaling needs resident sources and the conformal times of the transfer redshifts");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  cb200_handle::Scope sc(h, PH_SPLINE);
  const size_t N = (size_t)npts, nkz = (size_t)n_z * n_kt;
  // inputs: initpower [10], cosmo [6], kh [n_kt], transfer [n_z][n_kt], tautf [n_z] per point, then z [n_z]
  const size_t o_ip = 0, o_cs = o_ip + N * 10, o_kh = o_cs + N * 6, o_tr = o_kh + N * n_kt, o_tf = o_tr + N * nkz,
               o_z = o_tf + N * n_z, n_in = o_z + n_z;
  h->w_nl_in.alloc(n_in);
  double* di = h->w_nl_in.p;
  CB_CUDA(cudaMemcpyAsync(di + o_ip, initpower, sizeof(double) * N * 10, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(di + o_cs, cosmo, sizeof(double) * N * 6, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(di + o_kh, kh, sizeof(double) * N * n_kt, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(di + o_tr, transfer, sizeof(double) * N * nkz, cudaMemcpyHostToDevice, s));
  if (tautf) CB_CUDA(cudaMemcpyAsync(di + o_tf, tautf, sizeof(double) * N * n_z, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(di + o_z, z, sizeof(double) * n_z, cudaMemcpyHostToDevice, s));
  // work: logkh [n_kt], matpower, ddmat, ratio [n_z][n_kt], spec [n_z][3], sigma8 [n_z] per point
  const size_t w_lk = 0, w_mp = w_lk + N * n_kt, w_dd = w_mp + N * nkz, w_ra = w_dd + N * nkz, w_sp = w_ra + N * nkz,
               w_s8 = w_sp + N * n_z * 3, n_w = w_s8 + N * n_z;
  h->w_nl_work.alloc(n_w);
  h->w_nl_status.alloc(N);
  CB_CUDA(cudaMemsetAsync(h->w_nl_status.p, 0, sizeof(int) * N, s));
  double* dw = h->w_nl_work.p;
  NlParams P;
  P.np = npts; P.n_kt = n_kt; P.n_z = n_z;
  P.initpower = di + o_ip; P.cosmo = di + o_cs; P.kh = di + o_kh; P.z = di + o_z; P.transfer = di + o_tr;
  P.logkh = dw + w_lk; P.matpower = dw + w_mp; P.ddmat = dw + w_dd; P.ratio = dw + w_ra; P.spec = dw + w_sp; P.sigma8 = dw + w_s8;
  P.status = h->w_nl_status.p;
  nl_sigma8_kernel<<<(unsigned)((N * n_z + 127) / 128), 128, 0, s>>>(P);
  CB_LAUNCH_CHECK();
  nl_halofit_kernel<<<(unsigned)(N * n_z), NL_THREADS, sizeof(double) * 4 * n_kt, s>>>(P);
  CB_LAUNCH_CHECK();
  h->n_launches += 2;
  if (rescale_sources) {
    PointStore& S = h->store[0];
    PointView v = S.view();
    dim3 grid((S.NK + 127) / 128, npts);
    nl_rescale_kernel<<<grid, 128, 0, s>>>(v, first, npts, n_kt, n_z, P.cosmo, P.ratio, di + o_tf, S.src.p);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  if (sigma8) CB_CUDA(cudaMemcpyAsync(sigma8, P.sigma8, sizeof(double) * N * n_z, cudaMemcpyDeviceToHost, s));
  if (ratio) CB_CUDA(cudaMemcpyAsync(ratio, P.ratio, sizeof(double) * N * nkz, cudaMemcpyDeviceToHost, s));
  if (spec) CB_CUDA(cudaMemcpyAsync(spec, P.spec, sizeof(double) * N * n_z * 3, cudaMemcpyDeviceToHost, s));
  if (status) CB_CUDA(cudaMemcpyAsync(status, h->w_nl_status.p, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  return 0;
  CB_API_END(h)
}

// ---- thermal history (thermo.cuh) ----
int cb200_thermo(cb200_handle* h, int npts, const double* bg, const double* thermo_in, double* thermo_out, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (npts <= 0 || !bg || !thermo_in || !thermo_out) return fail(h, "thermo: bad arguments");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  cb200_handle::Scope sc(h, PH_BG);
  const int C = std::min(npts, 32768);   // points per pass: 720 KB of work tables per point (23.6 GB at 32 768 = the threads one B200 holds at 218 registers)
  h->w_th_work.alloc((size_t)C * (3 * TH_NZ + 3 * TH_NTHERMO));
  h->w_bg_in.alloc((size_t)C * NBG);
  h->w_th_in.alloc((size_t)C * TH_NIN);
  h->w_th_out.alloc((size_t)C * TH_NOUT);
  h->w_th_status.alloc(C);
  std::vector<int> st(C);
  for (int c0 = 0; c0 < npts; c0 += C) {
    const int np = std::min(C, npts - c0);
    ThermoScratch W;
    W.P = np;
    W.xrec = h->w_th_work.p; W.dxrec = W.xrec + (size_t)np * TH_NZ; W.work = W.dxrec + (size_t)np * TH_NZ;
    W.dotmu = W.work + (size_t)np * TH_NZ; W.sdotmu = W.dotmu + (size_t)np * TH_NTHERMO; W.sfac = W.sdotmu + (size_t)np * TH_NTHERMO;
    CB_CUDA(cudaMemcpyAsync(h->w_bg_in.p, bg + (size_t)c0 * NBG, sizeof(double) * np * NBG, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(h->w_th_in.p, thermo_in + (size_t)c0 * TH_NIN, sizeof(double) * np * TH_NIN, cudaMemcpyHostToDevice, s));
    thermo_recfast_kernel<<<(np + 63) / 64, 64, 0, s>>>(np, h->w_bg_in.p, h->w_th_in.p, h->bg_tables(), W, h->w_th_status.p);
    CB_LAUNCH_CHECK();
    thermo_init_kernel<<<(np + 63) / 64, 64, 0, s>>>(np, h->w_bg_in.p, h->w_th_in.p, h->bg_tables(), W, h->w_th_status.p, h->w_th_out.p);
    CB_LAUNCH_CHECK();
    h->n_launches += 2;
    CB_CUDA(cudaMemcpyAsync(thermo_out + (size_t)c0 * TH_NOUT, h->w_th_out.p, sizeof(double) * np * TH_NOUT, cudaMemcpyDeviceToHost, s));
    if (status) CB_CUDA(cudaMemcpyAsync(status + c0, h->w_th_status.p, sizeof(int) * np, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
  CB_API_END(h)
}

int cb200_theta_to_background(cb200_handle* h, int npts, const double* cosmo, const double* nu, double tcmb, double* bg) {
  if (!h) return -1;
  CB_API_BEGIN
  if (npts <= 0 || !cosmo || !nu || !bg) return fail(h, "theta_to_background: bad arguments");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  cb200_handle::Scope sc(h, PH_BG);
  h->w_bg_in.alloc((size_t)npts * NBG);
  h->w_th_in.alloc((size_t)npts * 16);
  CB_CUDA(cudaMemcpyAsync(h->w_bg_in.p, bg, sizeof(double) * npts * NBG, cudaMemcpyHostToDevice, s));   // keeps rdrag (bg[15])
  CB_CUDA(cudaMemcpyAsync(h->w_th_in.p, cosmo, sizeof(double) * npts * 8, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(h->w_th_in.p + (size_t)npts * 8, nu, sizeof(double) * npts * 8, cudaMemcpyHostToDevice, s));
  thermo_theta_kernel<<<(npts + 63) / 64, 64, 0, s>>>(npts, h->w_th_in.p, h->w_th_in.p + (size_t)npts * 8, tcmb, h->bg_tables(), h->w_bg_in.p);
  CB_LAUNCH_CHECK();
  h->n_launches += 1;
  CB_CUDA(cudaMemcpyAsync(bg, h->w_bg_in.p, sizeof(double) * npts * NBG, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  return 0;
  CB_API_END(h)
}

static int add_bg_z(cb200_handle* h, const double* z, int n) {
  const int off = (int)h->bg_z.size();
  h->bg_z.insert(h->bg_z.end(), z, z + n);
  h->bg_z_dirty = true;
  return off;
}

int cb200_like_add_bao(cb200_handle* h, int kind, int num_bao, const int* type, const double* z, const double* obs,
                       const double* invcov, double rs_rescale, double fixed_rs, const double* alpha_prob, int n_alpha,
                       int* like_id) {
  if (!h) return -1;
  CB_API_BEGIN
  if (num_bao < 1 || num_bao > BAO_MAXN) return fail(h, "bao: num_bao out of range");
  if (kind == 1 && (!alpha_prob || n_alpha < 399)) return fail(h, "bao: MGS needs its chi^2(alpha) table");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  std::unique_ptr<LikeEntry> L(new LikeEntry());
  L->type = 3;
  L->bao.kind = kind; L->bao.num_bao = num_bao; L->bao.rs_rescale = rs_rescale; L->bao.fixed_rs = fixed_rs;
  L->bao.n_alpha = n_alpha;
  for (int i = 0; i < num_bao; i++) {
    L->bao.type[i] = type ? type[i] : 2;
    L->bao.z[i] = z[i];
    L->bao.obs[i] = obs ? obs[i] : 0.0;
    if (kind == 0 && (L->bao.type[i] == 6 || L->bao.type[i] == 9 || L->bao.type[i] < 1 || L->bao.type[i] > 10))
      return fail(h, "bao: measurement type not supported (f_sigma8 / dilation need power spectra)");
  }
  if (invcov) L->bao_invcov.upload(invcov, (size_t)num_bao * num_bao, h->stream);
  if (alpha_prob) L->bao_prob.upload(alpha_prob, n_alpha, h->stream);
  L->z_off = add_bg_z(h, z, num_bao);
  L->nz = num_bao;
  CB_CUDA(cudaStreamSynchronize(h->stream));
  if (like_id) *like_id = (int)h->likes.size();
  h->likes.push_back(std::move(L));
  h->n_bg_likes++;
  return 0;
  CB_API_END(h)
}

int cb200_like_add_hst(cb200_handle* h, double H0, double H0_err, double zeff, double angconversion, int* like_id) {
  if (!h) return -1;
  CB_API_BEGIN
  std::unique_ptr<LikeEntry> L(new LikeEntry());
  L->type = 4; L->hst_H0 = H0; L->hst_err = H0_err; L->hst_zeff = zeff; L->hst_ang = angconversion;
  if (zeff > 0) { L->z_off = add_bg_z(h, &zeff, 1); L->nz = 1; }
  if (like_id) *like_id = (int)h->likes.size();
  h->likes.push_back(std::move(L));
  h->n_bg_likes++;
  return 0;
  CB_API_END(h)
}

int cb200_like_add_sn(cb200_handle* h, int nsn, const double* cols, const double* A1, const double* A2, int twoscriptm,
                      const double* const* cov, int alpha_index, int beta_index, int* like_id) {
  if (!h) return -1;
  CB_API_BEGIN
  if (nsn < 4) return fail(h, "sn: too few supernovae");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  std::unique_ptr<LikeEntry> L(new LikeEntry());
  L->type = 5; L->sn_ia = alpha_index; L->sn_ib = beta_index;
  L->sn_cols.upload(cols, (size_t)11 * nsn, s);
  std::vector<double> ones(nsn, 1.0), zeros(nsn, 0.0);
  L->sn_A1.upload(twoscriptm ? A1 : ones.data(), nsn, s);
  L->sn_A2.upload(twoscriptm ? A2 : zeros.data(), nsn, s);
  SnData& S = L->sn;
  S.nsn = nsn; S.twoscriptm = twoscriptm ? 1 : 0;
  const double* c = L->sn_cols.p;
  S.zcmb = c; S.zhel = c + nsn; S.mag = c + 2 * nsn; S.stretch = c + 3 * nsn; S.colour = c + 4 * nsn;
  S.pre_vars = c + 5 * nsn; S.stretch_var = c + 6 * nsn; S.colour_var = c + 7 * nsn; S.cov_ms = c + 8 * nsn;
  S.cov_mc = c + 9 * nsn; S.cov_sc = c + 10 * nsn;
  S.A1 = L->sn_A1.p; S.A2 = L->sn_A2.p;
  int ncov = 0;
  for (int m = 0; m < 6; m++) {
    S.cov[m] = nullptr;
    if (cov && cov[m]) { L->sn_cov[m].upload(cov[m], (size_t)nsn * nsn, s); S.cov[m] = L->sn_cov[m].p; ncov++; }
  }
  if (ncov == 0) return fail(h, "sn: diagonal-only errors are not supported (the shipped datasets all have covariances)");
  S.alphabeta = (ncov > 1 || !S.cov[0]) ? 1 : 0;  // alphabeta_covmat, supernovae_JLA.f90:735-741
  L->sn_nr = twoscriptm ? 3 : 2;
  L->sn_ld = (nsn + 3) / 4 * 4;
  if (!S.alphabeta) {
    if (twoscriptm) return fail(h, "sn: two-scriptM fit with a nuisance-independent covariance is not supported");
    // factor V once with the identity riding along: rows n.. of W become (L^-1)^T, then V^-1 = U U^T
    const int n = nsn, ld = L->sn_ld;
    DevBuf<double> W;
    W.alloc((size_t)2 * n * ld);
    W.zero(s);
    std::vector<double> eye((size_t)n * ld, 0.0);
    for (int i = 0; i < n; i++) eye[(size_t)i * ld + i] = 1.0;
    CB_CUDA(cudaMemcpyAsync(W.p + (size_t)n * ld, eye.data(), sizeof(double) * eye.size(), cudaMemcpyHostToDevice, s));
    DevBuf<int> bad;
    bad.alloc(1);
    CholParams cp;
    cp.n = n; cp.nr = n; cp.ld = ld; cp.np = 1; cp.assemble = 1; cp.S = S; cp.nuis = nullptr; cp.n_nuis = 0;
    cp.ia = -1; cp.ib = -1; cp.W = W.p; cp.pt_stride = (size_t)2 * n * ld; cp.status = bad.p;
    sn_chol_kernel<<<1, 256, CH_SMEM, s>>>(cp);
    CB_LAUNCH_CHECK();
    L->sn_vinv.alloc((size_t)n * n);
    dgemm(s, true, false, n, n, n, 1.0, W.p + (size_t)n * ld, ld, W.p + (size_t)n * ld, ld, L->sn_vinv.p, n, &h->n_launches);
    int hb = 0;
    std::vector<double> vi((size_t)n * n);
    CB_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaMemcpyAsync(vi.data(), L->sn_vinv.p, sizeof(double) * vi.size(), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (hb) return fail(h, "sn: covariance matrix is not positive definite");
    // amarg_E = sum of the inverse, upper triangle doubled (supernovae_JLA.f90:1123-1126)
    double E = 0;
    for (int i = 0; i < n; i++) {
      double su = 0;
      for (int j = 0; j < i; j++) su += vi[(size_t)j * n + i];
      E = E + vi[(size_t)i * n + i] + 2.0 * su;
    }
    L->sn_sum_vinv = E;
    h->n_launches += 1;
  }
  L->z_off = add_bg_z(h, cols, nsn);  // zcmb
  L->nz = nsn;
  CB_CUDA(cudaStreamSynchronize(s));
  if (like_id) *like_id = (int)h->likes.size();
  h->likes.push_back(std::move(L));
  h->n_bg_likes++;
  return 0;
  CB_API_END(h)
}

static int loglike_device(cb200_handle* h, int bg_first, int npts, const double* d_cls, const int* d_status,
                          const double* nuisance, int n_nuis, double* loglikes, double* total, int* status) {
  cudaStream_t s = h->stream;
  const int nlike = (int)h->likes.size();
  if (nlike == 0) return fail(h, "loglike: no likelihood registered");
  if (h->n_cmb_likes > 0 && !d_cls) return fail(h, "loglike: CMB likelihoods registered but no Cls resident");
  if (h->n_bg_likes > 0 && (bg_first < 0 || !h->r_bg.p)) return fail(h, "loglike: background likelihoods need cb200_set_background");
  const int LO = h->cfg.lmax_out + 1;
  h->w_ll.alloc((size_t)npts * nlike);
  h->w_total.alloc(npts);
  h->w_quad.alloc(npts);
  h->w_nuis.alloc((size_t)npts * std::max(1, n_nuis));
  if (n_nuis > 0) CB_CUDA(cudaMemcpyAsync(h->w_nuis.p, nuisance, sizeof(double) * npts * n_nuis, cudaMemcpyHostToDevice, s));
  const int nzt = (int)h->bg_z.size();
  const double* d_bg = h->n_bg_likes > 0 ? h->r_bg.p + (size_t)bg_first * NBG : nullptr;
  if (h->n_bg_likes > 0 && nzt > 0) {  // K5: D_A(z), H(z) at every redshift the registered likelihoods need
    cb200_handle::Scope sc(h, PH_BG);
    if (h->bg_z_dirty) { h->d_bgz.upload(h->bg_z, s); h->bg_z_dirty = false; }
    h->w_da.alloc((size_t)npts * nzt); h->w_hz.alloc((size_t)npts * nzt);
    const long long nthr = (long long)npts * nzt;
    bg_distance_kernel<<<(unsigned)((nthr + 127) / 128), 128, 0, s>>>(npts, nzt, d_bg, h->d_bgz.p, h->bg_tables(),
                                                                     h->w_da.p, h->w_hz.p);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  {
    cb200_handle::Scope sc(h, PH_LIKE);
    for (int li = 0; li < nlike; li++) {
      LikeEntry& L = *h->likes[li];
      if (L.cal_index >= n_nuis) return fail(h, "loglike: calibration index outside the nuisance vector");
      if (L.type == 1) {
        h->w_resid.alloc((size_t)npts * L.nused);
        h->w_T.alloc((size_t)npts * L.nused);
        PlikBinParams bp;
        bp.np = npts; bp.nused = L.nused; bp.lmax_out = h->cfg.lmax_out; bp.n_nuis = n_nuis; bp.cal_index = L.cal_index;
        bp.cls = d_cls; bp.bin_spec = L.bin_spec.p; bp.bin_lo = L.bin_lo.p; bp.bin_hi = L.bin_hi.p;
        bp.weights = L.weights.p; bp.x_data = L.x_data.p; bp.nuis = h->w_nuis.p; bp.resid = h->w_resid.p;
        dim3 grid((L.nused + 127) / 128, npts);
        plik_bin_kernel<<<grid, 128, 0, s>>>(bp);
        CB_LAUNCH_CHECK();
        dgemm(s, false, false, npts, L.nused, L.nused, 1.0, h->w_resid.p, L.nused, L.invcov.p, L.nused, h->w_T.p,
              L.nused, &h->n_launches);
        rowdot_kernel<<<(npts + 3) / 4, 128, 0, s>>>(npts, L.nused, h->w_T.p, h->w_resid.p, 0.5, h->w_ll.p + li, nlike, 0);
        CB_LAUNCH_CHECK();
        h->n_launches += 2;
      } else if (L.type == 2) {
        const int nb = L.nbins * L.ncl, nx = L.nbins * L.ncl_used;
        h->w_bc.alloc((size_t)npts * nb); h->w_bp.alloc((size_t)npts * nb);
        h->w_bigx.alloc((size_t)npts * nx); h->w_T.alloc((size_t)npts * std::max(nx, 1));
        h->w_binned.alloc((size_t)npts * nb);
        dgemm(s, false, false, npts, nb, 4 * LO, 1.0, d_cls, 5 * LO, L.Wt_cmb.p, nb, h->w_bc.p, nb, &h->n_launches);
        dgemm(s, false, false, npts, nb, LO, 1.0, d_cls + (size_t)4 * LO, 5 * LO, L.Wt_pp.p, nb, h->w_bp.p, nb,
              &h->n_launches);
        if (L.has_fg) {
          if (L.bk.nuis_off + 16 > n_nuis) return fail(h, "loglike: BK foreground parameters outside the nuisance vector");
          BkParams bk = L.bk;
          bk.np = npts; bk.n_nuis = n_nuis; bk.nuis = h->w_nuis.p; bk.binned = h->w_bc.p;
          bk.bp_nu = L.bk_nu.p; bk.bp_R = L.bk_R.p; bk.bp_dnu = L.bk_dnu.p; bk.bp_lnnu = L.bk_lnnu.p; bk.bp_w = L.bk_w.p; bk.fgW = L.bk_fgW.p; bk.lrange = L.bk_lrange.p;
          // frequency decorrelation makes the foreground of a map pair non-separable in l: the scalar kernel handles it;
          // otherwise (Delta_dust = Delta_sync = 1 for every point, the BK15 baseline) the sum is one GEMM for the batch
          bool decorr = false;
          for (int i = 0; i < npts && !decorr; i++) {
            const double* d = nuisance + (size_t)i * n_nuis + L.bk.nuis_off;
            decorr = std::fabs(d[10] - 1) > 1e-5 || std::fabs(d[11] - 1) > 1e-5;
          }
          if (decorr || h->bk_scalar_foregrounds) {
            bk_foreground_kernel<<<npts, 256, 0, s>>>(bk);
            CB_LAUNCH_CHECK();
            h->n_launches += 1;
          } else {
            const int LW = bk.lmax + 1;
            h->w_bk_shapes.alloc((size_t)npts * 3 * LW); h->w_bk_scal.alloc((size_t)npts * 3 * BK_MAXMAPS);
            h->w_bk_G.alloc((size_t)npts * 3 * nb);
            bk_shapes_kernel<<<npts, 256, 0, s>>>(bk, h->w_bk_shapes.p, h->w_bk_scal.p);
            CB_LAUNCH_CHECK();
            dgemm(s, false, false, 3 * npts, nb, LW, 1.0, h->w_bk_shapes.p, LW, L.bk_fgW.p, nb, h->w_bk_G.p, nb, &h->n_launches);
            dim3 gc((nb + 127) / 128, npts);
            bk_combine_kernel<<<gc, 128, 0, s>>>(bk, h->w_bk_G.p, h->w_bk_scal.p);
            CB_LAUNCH_CHECK();
            h->n_launches += 2;
          }
        }
        CmbLikesBinParams cp;
        cp.np = npts; cp.nmaps = L.nmaps; cp.ncl = L.ncl; cp.nbins = L.nbins; cp.ncl_used = L.ncl_used;
        cp.like_approx = L.like_approx; cp.n_nuis = n_nuis; cp.cal_index = L.cal_index;
        cp.bc = h->w_bc.p; cp.bp = h->w_bp.p; cp.offset = L.offset.p; cp.noise = L.has_noise ? L.noise.p : nullptr;
        cp.chat = L.chat.p; cp.sqrt_fid = L.has_sqrt_fid ? L.sqrt_fid.p : nullptr; cp.cl_use = L.cl_use.p;
        cp.nuis = h->w_nuis.p; cp.bigx = h->w_bigx.p; cp.binned_out = h->w_binned.p;
        cmblikes_bin_kernel<<<(unsigned)(((long long)npts * L.nbins + CMBL_THREADS - 1) / CMBL_THREADS), CMBL_THREADS, 0, s>>>(cp);
        CB_LAUNCH_CHECK();
        dgemm(s, false, false, npts, nx, nx, 1.0, h->w_bigx.p, nx, L.invcov.p, nx, h->w_T.p, nx, &h->n_launches);
        rowdot_kernel<<<(npts + 3) / 4, 128, 0, s>>>(npts, nx, h->w_T.p, h->w_bigx.p, 1.0, h->w_quad.p, 1, 0);
        CB_LAUNCH_CHECK();
        cmblikes_final_kernel<<<(npts + 127) / 128, 128, 0, s>>>(npts, h->w_quad.p, h->w_nuis.p, n_nuis, L.cal_index,
                                                                 L.log_cal_prior, h->w_ll.p + li, nlike);
        CB_LAUNCH_CHECK();
        h->n_launches += 3;
      } else if (L.type == 3) {
        BaoParams bp = L.bao;
        bp.np = npts; bp.nz_total = nzt; bp.z_off = L.z_off; bp.bg = d_bg; bp.DA = h->w_da.p; bp.H = h->w_hz.p;
        bp.invcov = L.bao_invcov.p; bp.alpha_prob = L.bao_prob.p; bp.out = h->w_ll.p + li; bp.out_stride = nlike;
        bao_kernel<<<(npts + 127) / 128, 128, 0, s>>>(bp);
        CB_LAUNCH_CHECK();
        h->n_launches += 1;
      } else if (L.type == 4) {
        hst_kernel<<<(npts + 127) / 128, 128, 0, s>>>(npts, d_bg, h->w_da.p, nzt, L.z_off, L.hst_zeff, L.hst_ang,
                                                     L.hst_H0, L.hst_err, h->w_ll.p + li, nlike);
        CB_LAUNCH_CHECK();
        h->n_launches += 1;
      } else if (L.type == 5) {
        if (L.sn_ia >= n_nuis || L.sn_ib >= n_nuis) return fail(h, "loglike: alpha/beta index outside the nuisance vector");
        const int n = L.sn.nsn;
        if (L.sn.alphabeta) {
          // K6: per-point covariance, blocked DMMA Cholesky with ride-along right-hand sides
          const int nr = L.sn_nr, ld = L.sn_ld;
          const size_t pst = (size_t)(n + nr) * ld;
          // points per launch: independent of the C_l chunk (4.4 MB of work matrix per point).  One launch of many CTAs keeps
          // every SM busy to the end (CTAs are scheduled as others finish): 1 024 points in one launch 9.6 us/point against
          // 10.8 in four launches of 256 (0.86 of a wave each)
          const int sub = std::min(npts, std::max(h->chunk, h->sn_chunk));
          h->w_snW.alloc((size_t)sub * pst);
          h->w_snbad.alloc(sub);
          for (int a = 0; a < npts; a += sub) {
            const int m = std::min(sub, npts - a);
            sn_prep_kernel<<<m, 256, 0, s>>>(L.sn, m, h->w_da.p + (size_t)a * nzt, nzt, L.z_off,
                                             h->w_nuis.p + (size_t)a * n_nuis, n_nuis, L.sn_ia, L.sn_ib,
                                             h->w_snW.p + (size_t)n * ld, pst, ld, 1);
            CB_LAUNCH_CHECK();
            const bool pre = h->sn_preassemble != 0;
            if (pre) {
              dim3 ga((n + 7) / 8, m);
              sn_assemble_kernel<<<ga, 256, 0, s>>>(L.sn, m, h->w_nuis.p + (size_t)a * n_nuis, n_nuis, L.sn_ia, L.sn_ib,
                                                    h->w_snW.p, pst, ld);
              CB_LAUNCH_CHECK();
              h->n_launches += 1;
            }
            CholParams cp;
            cp.n = n; cp.nr = nr; cp.ld = ld; cp.np = m; cp.assemble = pre ? 0 : 1; cp.S = L.sn;
            cp.nuis = h->w_nuis.p + (size_t)a * n_nuis; cp.n_nuis = n_nuis; cp.ia = L.sn_ia; cp.ib = L.sn_ib;
            cp.W = h->w_snW.p; cp.pt_stride = pst; cp.status = h->w_snbad.p;
            // four 4-warp CTAs per SM cover each other's serial phases slightly better than two 8-warp ones once the launch
            // fills all of them (>= 592 CTAs: 9.35 against 9.56 us/point); below that the 8-warp form wins
            const bool w4 = h->sn_chol_warps == 4 || (h->sn_chol_warps == 0 && m >= 592);
            if (pre && h->sn_chol_kernel_gen == 2 && w4) sn_chol2_kernel<4><<<m, 128, C2Cfg<4>::SMEM, s>>>(cp);
            else if (pre && h->sn_chol_kernel_gen == 2) sn_chol2_kernel<8><<<m, 256, C2Cfg<8>::SMEM, s>>>(cp);
            else sn_chol_kernel<<<m, 256, CH_SMEM, s>>>(cp);
            CB_LAUNCH_CHECK();
            sn_final_kernel<<<m, 256, 0, s>>>(m, n, L.sn.twoscriptm, h->w_snW.p + (size_t)n * ld, pst, ld, h->w_snbad.p,
                                              h->w_ll.p + (size_t)a * nlike + li, nlike);
            CB_LAUNCH_CHECK();
            h->n_launches += 3;
          }
        } else {
          // covariance independent of the nuisance parameters: cached inverse, one DMMA GEMM per batch
          h->w_resid.alloc((size_t)npts * n);
          h->w_T.alloc((size_t)npts * n);
          sn_prep_kernel<<<npts, 256, 0, s>>>(L.sn, npts, h->w_da.p, nzt, L.z_off, h->w_nuis.p, n_nuis, -1, -1,
                                              h->w_resid.p, (size_t)n, n, 0);
          CB_LAUNCH_CHECK();
          dgemm(s, false, false, npts, n, n, 1.0, h->w_resid.p, n, L.sn_vinv.p, n, h->w_T.p, n, &h->n_launches);
          sn_final_cached_kernel<<<npts, 256, 0, s>>>(npts, n, h->w_T.p, h->w_resid.p, L.sn_sum_vinv, h->w_ll.p + li, nlike);
          CB_LAUNCH_CHECK();
          h->n_launches += 2;
        }
      }
    }
    total_kernel<<<(npts + 127) / 128, 128, 0, s>>>(npts, nlike, h->w_ll.p, d_status, h->w_total.p);
    CB_LAUNCH_CHECK();
    h->n_launches += 1;
  }
  if (loglikes) CB_CUDA(cudaMemcpyAsync(loglikes, h->w_ll.p, sizeof(double) * npts * nlike, cudaMemcpyDeviceToHost, s));
  if (total) CB_CUDA(cudaMemcpyAsync(total, h->w_total.p, sizeof(double) * npts, cudaMemcpyDeviceToHost, s));
  if (status && d_status) CB_CUDA(cudaMemcpyAsync(status, d_status, sizeof(int) * npts, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  if (status && !d_status) for (int i = 0; i < npts; i++) status[i] = 0;
  return 0;
}

int cb200_loglike_batch(cb200_handle* h, int first, int npts, const double* nuisance, int n_nuis, double* loglikes,
                        double* total, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (first < 0 || npts <= 0 || first + npts > h->cfg.max_points) return fail(h, "loglike_batch: point range exceeds max_points");
  if (h->n_cmb_likes > 0 && !h->r_cls_out.p) return fail(h, "loglike_batch: no resident Cls");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  const bool have_cls = h->r_cls_out.p != nullptr;
  const int rc = loglike_device(h, first, npts, have_cls ? h->r_cls_out.p + (size_t)first * 5 * (h->cfg.lmax_out + 1) : nullptr,
                                have_cls ? h->r_status.p + first : nullptr, nuisance, n_nuis, loglikes, total, status);
  finish_results(h);   // asynchronous result copies of earlier cb200_powers calls are complete when -lnL is returned
  return rc;
  CB_API_END(h)
}

int cb200_loglike_cls(cb200_handle* h, int npts, const double* cls, const double* nuisance, int n_nuis,
                      double* loglikes, double* total, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)npts * 5 * (h->cfg.lmax_out + 1);
  h->w_cls_in.alloc(n);
  CB_CUDA(cudaMemcpyAsync(h->w_cls_in.p, cls, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  return loglike_device(h, h->r_bg.p ? 0 : -1, npts, h->w_cls_in.p, nullptr, nuisance, n_nuis, loglikes, total, status);
  CB_API_END(h)
}

int cb200_eval_batch(cb200_handle* h, const cb200_param_layout* L, int first, int npts, const double* params,
                     double* loglike, double* likelihoods, double* prior, int* status) {
  if (!h) return -1;
  CB_API_BEGIN
  if (!L || !params || !loglike || npts <= 0 || L->num_params <= 0) return fail(h, "eval_batch: bad arguments");
  const double logZero = 1e30;
  const int np_ = L->num_params;
  const double T = L->temperature > 0 ? L->temperature : 1.0;
  const int n_like = (int)h->likes.size();
  if (n_like == 0) return fail(h, "eval_batch: no likelihood registered");
  if (L->i_nuis_first + L->n_nuis > np_) return fail(h, "eval_batch: nuisance columns outside the parameter vector");
  auto col = [&](const double* P, int i, double def) { return i >= 0 ? P[i] : def; };
  std::vector<double> ip((size_t)npts * 10), al(npts), ap(npts), nuis((size_t)npts * std::max(L->n_nuis, 1), 0.0), pr(npts, 0.0);
  std::vector<int> st(npts, 0), stp(npts, 0), stl(npts, 0);
  for (int i = 0; i < npts; i++) {
    const double* P = params + (size_t)i * np_;
    // GetLogLikeBounds (calclike.f90:97-109)
    for (int j = 0; j < np_; j++)
      if ((L->pmax && P[j] > L->pmax[j]) || (L->pmin && P[j] < L->pmin[j])) { st[i] = 1; break; }
    // GetLogPriors (calclike.f90:111-134)
    double lp = 0;
    if (L->prior_std)
      for (int j = 0; j < np_; j++)
        if ((!L->use_prior || L->use_prior[j]) && L->prior_std[j] != 0) {
          const double d = (P[j] - (L->prior_mean ? L->prior_mean[j] : 0.0)) / L->prior_std[j];
          lp += d * d;
        }
    for (int c = 0; c < L->n_lincomb; c++)
      if (L->lincomb_std[c] != 0) {
        double dot = 0;
        for (int j = 0; j < np_; j++) dot += L->lincomb[(size_t)c * np_ + j] * P[j];
        const double d = (dot - L->lincomb_mean[c]) / L->lincomb_std[c];
        lp += d * d;
      }
    pr[i] = lp / 2;
    // CAMBCalc_SetCAMBInitPower (Calculator_CAMB.f90:839-877)
    double* q = &ip[(size_t)i * 10];
    q[0] = 1e-10 * std::exp(col(P, L->i_logA, L->def_logA));
    q[1] = col(P, L->i_ns, L->def_ns);
    q[2] = col(P, L->i_nrun, L->def_nrun);
    q[3] = col(P, L->i_nrunrun, L->def_nrunrun);
    q[4] = col(P, L->i_r, L->def_r);
    q[5] = col(P, L->i_nt, L->def_nt);
    q[6] = col(P, L->i_ntrun, L->def_ntrun);
    q[7] = L->pivot_scalar; q[8] = L->pivot_tensor; q[9] = L->inflation_consistency ? 1.0 : 0.0;
    al[i] = col(P, L->i_Alens, L->def_Alens);
    ap[i] = col(P, L->i_Aphiphi, L->def_Aphiphi);
    for (int j = 0; j < L->n_nuis; j++) nuis[(size_t)i * L->n_nuis + j] = P[L->i_nuis_first + j];
  }
  int rc = 0;
  if (h->kind[0].active) {
    // change mask (CalcLike_Cosmology.f90:59-94): recompute the spectra of a point only after a slow (new sources) or
    // semi-slow (initial power, ALens, Aphiphi) change; runs of such points go to cb200_powers together
    const int MP = h->cfg.max_points;
    if ((int)h->ev_epoch.size() < MP) { h->ev_epoch.assign(MP, 0); h->ev_key.assign((size_t)MP * 12, 0.0); }
    if ((int)h->src_epoch.size() < MP) h->src_epoch.assign(MP, 0);
    if (first < 0 || first + npts > MP) return fail(h, "eval_batch: point range exceeds max_points");
    std::vector<char> need(npts, 0);
    for (int i = 0; i < npts; i++) {
      if (st[i] == 1) continue;                               // out of bounds: the reference returns before any theory
      const int slot = first + i;
      double key[12];
      std::copy(&ip[(size_t)i * 10], &ip[(size_t)i * 10] + 10, key);
      key[10] = al[i]; key[11] = ap[i];
      const bool same = h->ev_epoch[slot] != 0 && h->ev_epoch[slot] == h->src_epoch[slot] + 1 &&
                        std::equal(key, key + 12, &h->ev_key[(size_t)slot * 12]);
      need[i] = !same;
    }
    for (int a = 0; a < npts;) {
      if (!need[a]) { a++; continue; }
      int b = a;
      while (b < npts && need[b]) b++;
      rc = cb200_powers(h, first + a, b - a, &ip[(size_t)a * 10], &al[a], &ap[a], nullptr, nullptr, &stp[a]);
      if (rc) return rc;
      for (int i = a; i < b; i++) {                           // (cb200_powers cleared these entries)
        const int slot = first + i;
        std::copy(&ip[(size_t)i * 10], &ip[(size_t)i * 10] + 10, &h->ev_key[(size_t)slot * 12]);
        h->ev_key[(size_t)slot * 12 + 10] = al[i]; h->ev_key[(size_t)slot * 12 + 11] = ap[i];
        h->ev_epoch[slot] = (stp[i] == 0) ? h->src_epoch[slot] + 1 : 0;   // a rejected point is never reused
      }
      h->ev_powers += b - a;
      a = b;
    }
    for (int i = 0; i < npts; i++) if (!need[i] && st[i] != 1) h->ev_reused++;
  }
  std::vector<double> ll((size_t)npts * n_like), tot(npts);
  rc = cb200_loglike_batch(h, first, npts, L->n_nuis ? nuis.data() : nullptr, L->n_nuis, ll.data(), tot.data(), stl.data());
  if (rc) return rc;
  for (int i = 0; i < npts; i++) {
    double v;
    if (st[i] == 1) v = logZero;                        // out of bounds: nothing else is looked at
    else {
      if (stp[i] != 0) st[i] = 1 + stp[i];
      else if (stl[i] != 0) st[i] = 1 + stl[i];
      bool zero = st[i] != 0;
      double sum = 0;
      for (int k = 0; k < n_like; k++) {
        const double x = ll[(size_t)i * n_like + k];
        if (!(x < logZero)) zero = true;                // logZero (or NaN) from a likelihood rejects the point
        sum += x;
      }
      v = zero ? logZero : sum / T + pr[i] / T;
    }
    loglike[i] = v;
    if (status) status[i] = st[i];
    if (prior) prior[i] = pr[i];
  }
  if (likelihoods) std::copy(ll.begin(), ll.end(), likelihoods);
  return 0;
  CB_API_END(h)
}

int cb200_test_like_batch(cb200_handle* h, int npts, int n, const double* x, const double* center, const double* covinv,
                          double* loglike) {
  if (!h) return -1;
  CB_API_BEGIN
  if (npts <= 0 || n <= 0 || !x || !covinv || !loglike) return fail(h, "test_like_batch: bad arguments");
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  std::vector<double> d((size_t)npts * n);
  for (int i = 0; i < npts; i++)
    for (int j = 0; j < n; j++) d[(size_t)i * n + j] = x[(size_t)i * n + j] - (center ? center[j] : 0.0);
  DevBuf<double> dX, dC, dT, dO;
  dX.upload(d.data(), d.size(), s);
  dC.upload(covinv, (size_t)n * n, s);
  dT.alloc((size_t)npts * n);
  dO.alloc(npts);
  cb200_handle::Scope sc(h, PH_LIKE);
  dgemm(s, false, false, npts, n, n, 1.0, dX.p, n, dC.p, n, dT.p, n, &h->n_launches);   // T = X covinv
  rowdot_kernel<<<(npts + 3) / 4, 128, 0, s>>>(npts, n, dT.p, dX.p, 0.5, dO.p, 1, 0);
  CB_LAUNCH_CHECK();
  h->n_launches += 1;
  CB_CUDA(cudaMemcpyAsync(loglike, dO.p, sizeof(double) * npts, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  return 0;
  CB_API_END(h)
}

int cb200_get_timing(cb200_handle* h, cb200_timing* t, int reset) {
  if (!h || !t) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  CB_CUDA(cudaStreamSynchronize(h->stream));
  float ms[PH_COUNT];
  for (int p = 0; p < PH_COUNT; p++) {
    ms[p] = 0;
    for (auto& e : h->ev[p]) {
      float x = 0;
      cudaEventElapsedTime(&x, e.first, e.second);
      ms[p] += x;
    }
  }
  t->ms_spline = ms[PH_SPLINE]; t->ms_project = ms[PH_PROJECT]; t->ms_contract = ms[PH_CONTRACT];
  t->ms_interp = ms[PH_INTERP]; t->ms_lens = ms[PH_LENS]; t->ms_like = ms[PH_LIKE]; t->ms_background = ms[PH_BG];
  t->ms_total = 0;
  for (int p = 0; p < PH_COUNT; p++) t->ms_total += ms[p];
  t->n_launches = h->n_launches;
  unsigned long long tr = 0;
  if (h->d_triples.p) CB_CUDA(cudaMemcpy(&tr, h->d_triples.p, sizeof(tr), cudaMemcpyDeviceToHost));
  t->proj_triples = (long long)tr;
  unsigned long long rs[16] = {0};
  if (h->d_ring_stats.p) CB_CUDA(cudaMemcpy(rs, h->d_ring_stats.p, sizeof(rs), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 6; i++) t->phase_cycles[i] = (long long)rs[4 + i];
  t->ring_slabs = (long long)rs[0]; t->ring_direct = (long long)rs[1]; t->ring_rows = (long long)rs[2];
  t->ring_pairs = (long long)rs[3];
  t->proj_mask_mismatch = (h->proj_kernel == 4) ? (long long)rs[1] : 0;
  t->eval_points_powers = h->ev_powers;
  t->eval_points_reused = h->ev_reused;
  if (getenv("CB200_DEBUG_WAIT")) fprintf(stderr, "[cb200] consumer warp0 wait cycles: all %llu, before last metadata arrive %llu, before ring-warp arrive %llu\n", rs[10], rs[11], rs[12]);
  if (reset) {
    for (int p = 0; p < PH_COUNT; p++) {
      for (auto& e : h->ev[p]) { h->ev_pool.push_back(e.first); h->ev_pool.push_back(e.second); }
      h->ev[p].clear();
    }
    h->n_launches = 0;
    h->ev_powers = h->ev_reused = 0;
    if (h->d_triples.p) h->d_triples.zero(h->stream);
    if (h->d_ring_stats.p) h->d_ring_stats.zero(h->stream);
  }
  return 0;
  CB_API_END(h)
}

// ---- FP64 peak micro-kernels -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* out) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double seed, double* out) {
  double c[8][2];
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = seed;
  const double a = 0.5 + threadIdx.x * 1e-6, b = 1e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) dmma_m8n8k4(c[k][0], c[k][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

int cb200_measure_fp64_peaks(cb200_handle* h, double* dfma_tflops, double* dmma_tflops) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  cudaDeviceProp prop;
  CB_CUDA(cudaGetDeviceProperties(&prop, h->cfg.device));
  DevBuf<double> out;
  out.alloc(1);
  const int blocks = prop.multiProcessorCount * 8, iters = 20000;
  cudaEvent_t a = h->get_event(), b = h->get_event();
  float ms = 0;
  double best_f = 0, best_m = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a, h->stream);
    dfma_peak_kernel<<<blocks, 256, 0, h->stream>>>(iters, 1.0, out.p);
    cudaEventRecord(b, h->stream);
    CB_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&ms, a, b);
    best_f = std::max(best_f, 2.0 * 8 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12);
    cudaEventRecord(a, h->stream);
    dmma_peak_kernel<<<blocks, 256, 0, h->stream>>>(iters / 4, 1.0, out.p);
    cudaEventRecord(b, h->stream);
    CB_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&ms, a, b);
    best_m = std::max(best_m, 2.0 * 256 * 8 * (iters / 4) * 8.0 * blocks / (ms * 1e-3) / 1e12);
  }
  h->ev_pool.push_back(a); h->ev_pool.push_back(b);
  if (dfma_tflops) *dfma_tflops = best_f;
  if (dmma_tflops) *dmma_tflops = best_m;
  return 0;
  CB_API_END(h)
}

int cb200_set_option(cb200_handle* h, const char* name, double value) {
  if (!h || !name) return -1;
  std::string n(name);
  if (n == "count_triples") h->count_triples = value != 0;
  else if (n == "keep_transfers") h->keep_transfers = value != 0;
  else if (n == "ring_stats") h->ring_stats = value != 0;
  else if (n == "async_upload") h->async_upload = value != 0;
  else if (n == "async_results") h->async_results = value != 0;
  else if (n == "sn_preassemble") h->sn_preassemble = value != 0;
  else if (n == "sn_chol_kernel") h->sn_chol_kernel_gen = (int)value;
  else if (n == "sn_chol_warps") h->sn_chol_warps = ((int)value == 4) ? 4 : ((int)value == 8 ? 8 : 0);
  else if (n == "sn_chunk") h->sn_chunk = std::max(1, (int)value);
  else if (n == "bk_scalar_foregrounds") h->bk_scalar_foregrounds = value != 0;
  else if (n == "spline_kernel") h->spline_kernel = (value == 1) ? 1 : 2;
  else if (n == "proj_kernel") h->proj_kernel = (value >= 1 && value <= 4) ? (int)value : 4;
  else return fail(h, "set_option: unknown option " + n);
  return 0;
}

int cb200_timer_start(cb200_handle* h) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  if (!h->t_a) { h->t_a = h->get_event(); h->t_b = h->get_event(); }
  CB_CUDA(cudaStreamSynchronize(h->stream));
  CB_CUDA(cudaEventRecord(h->t_a, h->stream));
  return 0;
  CB_API_END(h)
}

int cb200_timer_stop(cb200_handle* h, float* ms) {
  if (!h || !h->t_a) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  CB_CUDA(cudaEventRecord(h->t_b, h->stream));
  CB_CUDA(cudaEventSynchronize(h->t_b));
  float x = 0;
  CB_CUDA(cudaEventElapsedTime(&x, h->t_a, h->t_b));
  if (ms) *ms = x;
  return 0;
  CB_API_END(h)
}

int cb200_sync(cb200_handle* h) {
  if (!h) return -1;
  CB_API_BEGIN
  CB_CUDA(cudaSetDevice(h->cfg.device));
  CB_CUDA(cudaStreamSynchronize(h->copy_stream));
  CB_CUDA(cudaStreamSynchronize(h->stream));
  finish_results(h);
  return 0;
  CB_API_END(h)
}

}  // extern "C"
