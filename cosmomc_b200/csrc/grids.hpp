// Host-side grid builders of the B200 path (product code).
//
// These produce the integer-valued artefacts that must be bit-identical to the reference:
//   * piecewise linear/log sample grids with region merging  (reference: camb/utils.F90:179-470)
//   * the sparse multipole set                                (reference: camb/modules.f90:791-950)
//   * Bessel-table abscissae                                  (reference: camb/bessels.f90:64-77)
//   * integration wavenumbers, time steps, source wavenumbers (reference: camb/cmbmain.f90:1221-1293,
//     camb/modules.f90:2994-3027, camb/cmbmain.f90:794-849)
// All arithmetic is plain IEEE double without FMA contraction (compile with -ffp-contract=off / nvcc
// host code does not contract), because point counts come from truncating FP expressions.
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

namespace cb200 {

struct Segment {     // one uniformly (lin or log) sampled stretch of a grid
  double lo, hi;     // [lo, hi)
  double step;       // spacing (in log units when is_log)
  int n;             // number of samples in the stretch (the sample at hi belongs to the next one)
  int first;         // 1-based index of the sample at lo
  bool is_log;
  double step_min, step_max;  // smallest / largest linear spacing inside
};

// A sampling grid assembled from requested stretches; overlapping requests are resolved to the finer
// spacing and slivers are merged into their neighbours (tolerance 0.1 of a bin, as in the reference).
class SampleGrid {
 public:
  static constexpr double kTol = 0.1;
  std::vector<Segment> seg;
  std::vector<double> x, dx;
  double lowest = 0, highest = 0;
  int npoints = 0;

  void clear() { seg.clear(); req_.clear(); x.clear(); dx.clear(); npoints = 0; }

  void add_spacing(double a, double b, double approx_step, bool is_log = false) {
    if (!(b > a)) throw std::runtime_error("SampleGrid: end must exceed start");
    if (!(approx_step > 0)) throw std::runtime_error("SampleGrid: step must be positive");
    double span = is_log ? std::log(b / a) : (b - a);
    int n = std::max(1, (int)(span / approx_step + 1.0 - kTol));
    add_count(a, b, n, is_log);
  }

  void add_count(double a, double b, int n, bool is_log = false) {
    if (!(b > a)) throw std::runtime_error("SampleGrid: end must exceed start");
    if (n <= 0) throw std::runtime_error("SampleGrid: count must be positive");
    // the requests seen by each rebuild are the *previous resolved segments* plus the new one,
    // exactly like the reference which re-merges its current region list with the newcomer.
    req_.assign(seg.begin(), seg.end());
    Segment s{};
    s.lo = a; s.hi = b; s.n = n; s.is_log = is_log;
    s.step = (is_log ? std::log(b / a) : (b - a)) / n;
    req_.push_back(s);
    rebuild();
  }

  // 1-based index of the last sample <= v, evaluated with the reference's arithmetic.
  int index_of(double v) const {
    for (const Segment& s : seg)
      if (v < s.hi && v >= s.lo)
        return s.first + (int)((s.is_log ? std::log(v / s.lo) : (v - s.lo)) / s.step);
    if (v >= highest) return npoints;
    throw std::runtime_error("SampleGrid::index_of: value below grid");
  }

  void materialise(bool half_weight_ends = true) {
    x.assign(npoints, 0.0);
    int k = 0;
    for (const Segment& s : seg)
      for (int j = 0; j < s.n; j++) x[k++] = s.is_log ? s.lo * std::exp(j * s.step) : s.lo + s.step * j;
    x[k++] = highest;
    if (k != npoints) throw std::runtime_error("SampleGrid: inconsistent point count");
    dx.assign(npoints, 0.0);
    for (int i = 1; i + 1 < npoints; i++) dx[i] = (x[i + 1] - x[i - 1]) / 2;
    double e0 = x[1] - x[0], e1 = x[npoints - 1] - x[npoints - 2];
    dx[0] = half_weight_ends ? e0 / 2 : e0;
    dx[npoints - 1] = half_weight_ends ? e1 / 2 : e1;
  }

 private:
  std::vector<Segment> req_;

  static void set_lin_extents(Segment& s) {
    if (s.is_log) {
      if (s.n == 1) s.step_min = s.step_max = s.hi - s.lo;
      else { s.step_min = s.lo * (std::exp(s.step) - 1); s.step_max = s.hi * (1 - std::exp(-s.step)); }
    } else s.step_min = s.step_max = s.step;
  }

  void rebuild() {
    // 1. ordered, de-duplicated break points.  The reference inserts each request's two ends into a
    //    sorted list with strict '<' comparisons, appending when nothing is larger; the resulting list
    //    is the sorted multiset of all ends, then exact duplicates are dropped.
    std::vector<double> ends;
    for (const Segment& r : req_) { insert_sorted(ends, r.lo); insert_sorted(ends, r.hi); }
    std::vector<double> br;
    for (double e : ends) if (br.empty() || e != br.back()) br.push_back(e);
    lowest = br.front(); highest = br.back();
    const double widest = highest - lowest;

    // 2. resolve each elementary interval to the finest request covering its left end
    std::vector<Segment> out;
    std::vector<double> asked;
    for (size_t i = 0; i + 1 < br.size(); i++) {
      Segment s{};
      s.lo = br[i]; s.hi = br[i + 1]; s.is_log = false;
      double d = widest;
      for (const Segment& r : req_) {
        if (!(s.lo >= r.lo && s.lo < r.hi)) continue;
        if (r.is_log) {
          if (s.is_log) d = std::min(d, r.step);
          else {
            double fine = s.lo * (std::exp(r.step) - 1);
            if (fine < d) {
              double coarse = s.hi * (1 - std::exp(-r.step));
              if (d < coarse) d = fine;
              else { s.is_log = true; d = r.step; }
            }
          }
        } else {
          if (s.is_log) {
            double coarse = s.hi * (1 - std::exp(-d));
            if (r.step < coarse) {
              double fine = s.lo * (std::exp(d) - 1);
              if (fine < r.step) { s.is_log = false; d = fine; }
              else d = -std::log(1 - r.step / s.hi);
            }
          } else d = std::min(d, r.step);
        }
      }
      double span = s.is_log ? std::log(s.hi / s.lo) : (s.hi - s.lo);
      if (d >= span) { s.step = span; s.n = 1; }
      else { s.n = std::max(1, (int)(span / d + 1.0 - kTol)); s.step = span / s.n; }
      set_lin_extents(s);
      out.push_back(s);
      asked.push_back(d);
    }

    // 3. absorb single-sample slivers into a neighbour whose spacing is compatible (scan from the top)
    for (int i = (int)out.size() - 1; i >= 0; i--) {
      Segment& s = out[i];
      if (s.n != 1) continue;
      double width = s.hi - s.lo, lo_req, hi_req;
      if (s.is_log) { lo_req = s.lo * (std::exp(asked[i]) - 1); hi_req = s.hi * (1 - std::exp(-asked[i])); }
      else lo_req = hi_req = asked[i];
      if (i != (int)out.size() - 1) {
        Segment& up = out[i + 1];
        if (asked[i] >= s.step && width <= up.step_min && up.step_min <= hi_req) {
          up.lo = s.lo;
          if (width > up.step_min * kTol) up.n += 1;
          up.step = (up.is_log ? std::log(up.hi / up.lo) : (up.hi - up.lo)) / up.n;
          out.erase(out.begin() + i);  // 'asked' deliberately keeps its indexing (as the reference does)
          continue;
        }
      }
      if (i != 0) {
        Segment& dn = out[i - 1];
        if (asked[i] >= s.step && width <= dn.step_max && dn.step_max <= lo_req) {
          dn.hi = s.hi;
          if (width > dn.step_max * kTol) dn.n += 1;
          dn.step = (dn.is_log ? std::log(dn.hi / dn.lo) : (dn.hi - dn.lo)) / dn.n;
          out.erase(out.begin() + i);
        }
      }
    }

    // 4. indices
    int next = 1;
    for (Segment& s : out) { s.first = next; next += s.n; set_lin_extents(s); }
    npoints = next;
    seg.swap(out);
  }

  static void insert_sorted(std::vector<double>& v, double e) {
    auto it = v.begin();
    while (it != v.end() && !(e < *it)) ++it;
    v.insert(it, e);
  }
};

inline int round_half_away(double v) { return (int)std::lround(v); }

// Sparse multipole set at which C_l is computed (others are splined).  Flat, non-log sampling.
struct LSampleOpts {
  double l_sample_boost = 1;
  bool accurate_reion = true;
};
inline std::vector<int> make_l_samples(int max_l, const LSampleOpts& o = LSampleOpts()) {
  std::vector<int> l;
  const double sc = 1.0 / o.l_sample_boost;
  if (o.l_sample_boost >= 50) { for (int v = 2; v <= max_l; v++) l.push_back(v); return l; }
  for (int v = 2; v <= 10; v++) l.push_back(v);
  int step, bot, top;
  if (o.accurate_reion) {
    for (int v = 11; v <= 37; v += (o.l_sample_boost > 1 ? 1 : 2)) l.push_back(v);
    step = std::max(round_half_away(5 * sc), 2); bot = 40; top = bot + step * 10;
  } else {
    if (o.l_sample_boost > 1) for (int v = 11; v <= 15; v++) l.push_back(v);
    else { l.push_back(12); l.push_back(15); }
    step = std::max(round_half_away(10 * sc), 3); bot = 15 + std::max(step / 2, 2); top = bot + step * 7;
  }
  for (int v = bot; v <= top; v += step) l.push_back(v);
  auto clip_to_max = [&]() {  // drop entries above max_l and make max_l the last one
    while (!l.empty() && l.back() > max_l) l.pop_back();
    if (l.back() < max_l) l.push_back(max_l);
  };
  step = std::max(round_half_away(20 * sc), 4);
  bot = l.back() + step; top = bot + step * 2;
  for (int v = bot; v <= top; v += step) l.push_back(v);
  if (l.back() >= max_l) { clip_to_max(); return l; }
  step = std::max(round_half_away(25 * sc), 4);
  bot = l.back() + step; top = bot + step;
  for (int v = bot; v <= top; v += step) l.push_back(v);
  if (l.back() >= max_l) { clip_to_max(); return l; }
  step = std::max(round_half_away(50 * sc), 7);  // use_spline_template = true
  bot = l.back() + step; top = std::min(5000, max_l);
  for (int v = bot; v <= top; v += step) l.push_back(v);
  if (max_l > 5000) {
    step = std::max(round_half_away(400 * sc), 50);
    int v = l.back();
    for (;;) { v += step; if (v > max_l) break; l.push_back(v); step = round_half_away(step * (double)1.5f); }
  }
  if (l.back() != max_l) l.push_back(max_l);
  return l;
}

// Bessel-table abscissae for arguments up to kmax (reference: bessels.f90:64-77)
inline void make_bessel_x(SampleGrid& g, double max_eta_k, double accuracy_boost = 1) {
  int kmax = (int)max_eta_k + 1;
  g.clear();
  g.add_spacing(0., 1., 0.01);
  g.add_spacing(1., 5., 0.1);
  g.add_spacing(5., 25., 0.2);
  g.add_spacing(25., 150., 0.5 / accuracy_boost);
  g.add_spacing(150., (double)kmax, 0.8 / accuracy_boost);
  g.materialise();
}

// Integration wavenumbers for the flat line-of-sight projection (reference: cmbmain.f90:1221-1293)
inline void make_q_grid(SampleGrid& g, double tau0, double max_eta_k, int max_l, bool high_accuracy = true,
                        double boost = 1) {
  const double qmax = max_eta_k / tau0, qmin = 0.1 / tau0 / boost;
  const double qmax_int = std::min(qmax, max_eta_k / tau0);
  const int lognum = round_half_away(10 * boost), nlin = round_half_away(600 * boost);
  const double dk0 = 1.8 / tau0 / boost;
  double dk = 3. / tau0 / boost;
  if (high_accuracy) dk = dk / (double)1.6f;
  const double k_log_end = lognum * dk0, k_lin_end = nlin * dk0;
  const double dk_fine = (double)0.04f / boost;
  g.clear();
  g.add_spacing(qmin, k_log_end, 1. / lognum, true);
  g.add_spacing(k_log_end, std::min(qmax_int, k_lin_end), dk0);
  if (qmax_int > k_lin_end) {
    double k_dk_end = std::max(3000, 2 * max_l) / tau0;
    g.add_spacing(k_lin_end, std::min(qmax_int, k_dk_end), dk);
    if (qmax_int > k_dk_end) g.add_spacing(k_dk_end, qmax_int, dk_fine);
  }
  g.materialise(true);
}

// Conformal-time samples at which sources are stored (reference: modules.f90:2994-3027, with
// dtaurec from cmbmain.f90:742-745 and modules.f90:2910-2915).
inline void make_time_steps(SampleGrid& g, double taurst, double taurend, double tau0, double max_eta_k,
                            bool tensors, double reion_start, double reion_complete, double boost = 1) {
  double dtaurec = 4 / (max_eta_k / tau0) / boost;
  dtaurec = std::min(dtaurec, tensors ? taurst / 160 : taurst / 40) / boost;
  g.clear();
  g.add_spacing(taurst, taurend, dtaurec);
  double dtau0 = tensors ? std::max(taurst / 40, tau0 / 2000. / boost) : tau0 / 500. / boost;
  g.add_spacing(taurend, tau0, dtau0);
  if (reion_start > 0) g.add_count(reion_start, reion_complete, (int)(50 * boost));
  g.materialise(true);
}

// Wavenumbers at which the Boltzmann sources are evolved (reference: cmbmain.f90:794-849), flat.
inline void make_source_k(SampleGrid& g, double tau0, double taurst, double max_eta_k, bool tensors, int max_l,
                          double boost = 1) {
  const double qmax = max_eta_k / tau0, qmin = 0.1 / tau0 / boost;
  double dlnk0 = 2. / 10 / boost;  // scalars with reionisation and accurate polarisation
  if (tensors) dlnk0 = 5. / 10 / boost;
  dlnk0 = dlnk0 / 2;  // AccurateReionization
  double dkn1 = 0.6 / taurst / boost, dkn2 = 0.9 / taurst / boost;
  dkn2 = dkn2 / (double)1.2f;  // HighAccuracyDefault
  if (tensors) { dkn1 = dkn1 * 0.8; dlnk0 = dlnk0 / 2; dkn2 = dkn2 * 0.85; }
  const double q_log_end = dkn1 / dlnk0;
  const double q_switch = (double)(2 * 6.3f) / taurst;
  double q_cmb = 2 * 3000 / tau0 * boost;
  if (max_l > 5000) q_cmb = q_cmb * (double)1.4f;
  double dksmooth = q_cmb / 2 / (boost * boost) / 6;
  g.clear();
  g.add_spacing(qmin, q_log_end, dlnk0, true);
  g.add_spacing(q_log_end, std::min(qmax, q_switch), dkn1);
  if (qmax > q_switch) {
    g.add_spacing(q_switch, std::min(q_cmb, qmax), dkn2);
    if (qmax > q_cmb) g.add_spacing(q_cmb, qmax, std::log(1 + dksmooth / q_cmb), true);
  }
  g.materialise(true);
}

}  // namespace cb200
