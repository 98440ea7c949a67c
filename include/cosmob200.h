/* cosmob200 — C ABI of the B200-native theory + likelihood path behind CosmoMC's calculator /
 * likelihood plug-in surface.  Plain pointers and sizes only; the library owns device memory, the caller
 * owns every host buffer.  All arrays are C-contiguous, point-major.  One handle per host thread; there are
 * no globals (the reference's CAMB state is global and thread-unsafe, camb/cmbmain.f90:7-8).
 *
 * Return value of every entry point: 0 = OK, >0 = soft error (CAMB-style: the caller maps it to `error`
 * and rejects the point, source/Calculator_CAMB.f90:205-211), <0 = usage / CUDA error (cb200_last_error()).
 * The library never exits or aborts the process.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef COSMOB200_H
#define COSMOB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_VERSION 2
#define CB200_LOGZERO 1e30 /* source/settings.f90:114 */

typedef struct cb200_handle cb200_handle;

/* Mirrors what CAMBCalc_InitCAMBParams derives from CosmoSettings (source/Calculator_CAMB.f90:729-836). */
typedef struct cb200_config {
  int struct_size;           /* sizeof(cb200_config) as the CALLER's binding lays it out; set by cb200_default_config.
                                cb200_create refuses a value other than its own sizeof (a Fortran / ctypes mirror that
                                fell out of step with this header fails loudly instead of reading garbage) */
  int device;                /* CUDA device ordinal */
  int lmax_computed_cl;      /* CosmoSettings%lmax_computed_cl (batch3: 2500) */
  int cmb_lensing;           /* CosmoSettings%CMB_Lensing */
  int use_lensing_potential; /* CosmoSettings%use_lensing_potential */
  int use_nonlinear_lensing; /* CosmoSettings%use_nonlinear_lensing */
  int compute_tensors;       /* CosmoSettings%compute_tensors */
  int lmax_tensor;           /* CosmoSettings%lmax_tensor (600) */
  int accurate_bb;           /* CAMB_Calculator%accurate_BB */
  double k_eta_max_scalar;   /* CAMB_Calculator%k_eta_max_scalar, <=0: default rule */
  double accuracy_level;     /* AccuracyLevel (only 1 supported) */
  int lmax_out;              /* CosmoSettings%lmax: highest l of the returned Cls (>= lmax_computed_cl ok) */
  int highl_norm_first_call; /* 1: keep the reference's SAVEd highL_norm (Calculator_CAMB.f90:358,396);
                                0 (default): renormalise the l>lmax_computed tail per point */
  int max_points;            /* capacity of resident source storage (points) */
  int chunk_points;          /* points processed per internal pass (work-buffer size); 0 = auto */
  int n_tau_max, n_k_max, n_q_max; /* capacities per point; 0 = defaults 768 / 256 / 3072 */
  int n_tau_max_tensor, n_k_max_tensor, n_q_max_tensor; /* same for the tensor pass; 0 = 2304 / 128 / 1024 */
} cb200_config;

typedef struct cb200_info {
  int max_l, max_eta_k;       /* CAMB Max_l, Max_eta_k after InitCAMBParams */
  int max_l_tensor, max_eta_k_tensor;
  int n_lsamp, n_lsamp_tensor; /* l-sample counts (camb/modules.f90:791-950) */
  int num_xx;                  /* Bessel abscissae (camb/bessels.f90:64-77) */
  int lmax_lensed;             /* camb/lensing.f90:153-157 */
  int lens_lmax;               /* lmax of the lensing sums (lmax_extrap, camb/lensing.f90:96-101) */
  int lens_npoints;            /* theta samples actually integrated (camb/lensing.f90:163-177) */
  int lens_jmax;               /* sampled l in the correlation sums (camb/lensing.f90:183-189) */
  int n_tau_max, n_k_max, n_q_max, max_points, chunk_points;
  int num_xx_tensor;           /* Bessel abscissae of the tensor pass (kmax = Max_eta_k_tensor) */
} cb200_info;

void cb200_default_config(cb200_config* cfg);
/* sizeof(cb200_config) / CB200_VERSION of the loaded library, for bindings that want to check before anything else */
int cb200_config_size(void);
int cb200_abi_version(void);
int cb200_create(const cb200_config* cfg, cb200_handle** out);
void cb200_destroy(cb200_handle* h);
const char* cb200_last_error(const cb200_handle* h);
int cb200_get_info(const cb200_handle* h, cb200_info* info);
/* l-sample set (replaces lSamp from initlval, camb/modules.f90:791). kind: 0 scalar, 1 tensor */
int cb200_get_lsamples(const cb200_handle* h, int kind, int* l, int* n);

/* Fiducial templates read by the reference at run time:
 *   highl_unlensed [4][8001] TT,EE,TE,PP   camb/modules.f90:1162-1185 (HighLExtrapTemplate_lenspotentialCls.dat)
 *   highl_lensed   [4][n_l]  TT,EE,BB,TE   source/Calculator_CAMB.f90:966-990 (HighL_lensedCls.dat, muK^2) */
int cb200_set_templates(cb200_handle* h, const double* highl_unlensed, const double* highl_lensed, int n_l);

/* ---- host grid helpers (bit-exact restatement of the reference's Ranges-based grids) -------------------
 * kind: 0 scalar, 1 tensor.  Each returns the number of samples in *n (arrays sized max_n). */
int cb200_make_q_grid(const cb200_handle* h, int kind, double tau0, int max_n, double* q, double* dq, int* n);
    /* camb/cmbmain.f90:1221-1293 SetkValuesForInt */
int cb200_make_time_steps(const cb200_handle* h, int kind, double tau0, double taurst, double taurend,
                          double reion_tau_start, double reion_tau_complete, int max_n, double* tau,
                          double* dtau, int* n); /* camb/modules.f90:2994-3027 SetTimeSteps */
int cb200_make_source_k(const cb200_handle* h, int kind, double tau0, double taurst, int max_n, double* k,
                        int* n);                 /* camb/cmbmain.f90:794-849 SetkValuesForSources */
/* Generic grid builder for tests: ops[i] = {kind(0 spacing,1 count), start, end, step_or_count, is_log} */
int cb200_grid_build(int nops, const double* ops, int max_n, double* x, double* dx, int* n, int n_query,
                     const double* query, int* index_out);
/* Bessel table read-back (camb/bessels.f90:50-120): x [num_xx], ajl/ajlpr [n_lsamp][num_xx] */
int cb200_get_bessel_table(const cb200_handle* h, int kind, double* x, double* ajl, double* ajlpr);

/* ---- calculator: slow step ------------------------------------------------------------------------------
 * Replaces the part of CAMBCalc_GetNewTransferData after the source ODEs (source/Calculator_CAMB.f90:179-218
 * -> camb/cmbmain.f90:238-263).  Inputs per point (what CAMB holds in globals after DoSourcek):
 *   thermo [npts][5] = tau0, taurst, taurend, reion_tau_start (<=0: none), reion_tau_complete
 *   n_k [npts], k [npts][n_k_max]                     Evolve_q%points
 *   src  [npts][n_tau][n_src=3][n_k] packed per point with strides (n_tau_max, 3, n_k_max)  == Src(k,s,tau)
 * The time-step grid and q grid are rebuilt inside from `thermo` exactly as the reference does.
 * `first` = index of the first point slot to fill (resident storage), kind 0 scalar / 1 tensor.
 * src_is_device != 0: `src` is a device pointer (used by the HBM-resident benchmark leg). */
int cb200_upload_sources(cb200_handle* h, int kind, int first, int npts, const double* thermo, const int* n_k,
                         const double* k, const double* src, int src_is_device);
/* Same hand-off with the sources PACKED at their exact sizes, the way CAMB holds them: point i contributes
 * Src(1:n_k[i], 1:3, 1:n_tau[i]) (= C [n_tau][3][n_k]) back to back in src_packed, no padding (7 % fewer bytes over
 * PCIe than the padded form at the batch3 sizes, and no padding copy on the Fortran side).  n_tau[i] must equal the
 * number of time steps the library derives from thermo[i] (checked; camb/modules.f90:2994-3027).  A device kernel
 * scatters the block into the resident padded layout. */
int cb200_upload_sources_packed(cb200_handle* h, int kind, int first, int npts, const double* thermo, const int* n_tau,
                                const int* n_k, const double* k, const double* src_packed);

/* ---- calculator: semi-slow step -------------------------------------------------------------------------
 * Replaces CAMBCalc_GetNewPowerData (source/Calculator_CAMB.f90:220-275): k-contraction, l-interpolation,
 * lensing, unit conversion for points [first, first+npts).
 *   initpower [npts][10] = As(absolute), ns, nrun, nrunrun, r, nt, ntrun, pivot_k, tensor_pivot_k,
 *                          inflation_consistency      (CAMBCalc_SetCAMBInitPower, :839-877)
 *   alens [npts] (ALens), aphiphi [npts] (Aphiphi; NULL = 1)
 *   cls_out [npts][5][lmax_out+1]  TT,TE,EE,BB,PP in CosmoMC units (NULL: keep on device only)
 *   derived_out [npts][4] rms deflection (arcmin), tensor ratio_02, ratio_BB, AT  (NULL ok)
 *   status [npts] 0 ok / >0 rejected (NaN or negative TT/EE/BB, :239-256) */
int cb200_powers(cb200_handle* h, int first, int npts, const double* initpower, const double* alens,
                 const double* aphiphi, double* cls_out, double* derived_out, int* status);

/* Same step when the transfer functions are SHARED by the whole batch: CosmoMC calls GetNewPowerData without
 * GetNewTransferData when only the initial-power block moved (source/CalcLike_Cosmology.f90:73-85; e.g. BK15 chains
 * at fixed cosmology).  Sources of resident point `src_point` (scalar, and tensor when compute_tensors) are projected
 * once, the k-contraction of the npts initial-power points is one FP64 tensor-pipe GEMM; outputs go to the resident
 * slots [first, first+npts). */
int cb200_powers_shared(cb200_handle* h, int src_point, int first, int npts, const double* initpower,
                        const double* alens, const double* aphiphi, double* cls_out, double* derived_out, int* status);

/* Intermediate read-backs for parity tests (device -> host copies of the last cb200_powers call):
 *   what: 0 iCl [6][n_lsamp], 1 Cl_scalar [6][max_l+1], 2 Cl_lensed [4][max_l+1] (dimensionless),
 *         3 transfers Delta [n_q][n_lsamp_pad][3] (only if cb200_keep_transfers(h,1)), 4 q, 5 dq, 6 tau, 7 dtau,
 *         8 tensor iCl [4][n_lsamp_tensor], 9 Cl_tensor [4][lmax_tensor+1] TT,EE,BB,TE (dimensionless),
 *         10 the resident lensing-potential source of the point [n_tau][n_k_max] */
int cb200_debug_fetch(cb200_handle* h, int what, int point, int max_n, double* out, int* n);
int cb200_keep_transfers(cb200_handle* h, int on);

/* ---- likelihoods ----------------------------------------------------------------------------------------
 * Each cb200_like_add_* takes the dense arrays the reference's ReadIni resolves a .dataset to (parsing is
 * host-side, cosmomc_b200/datasets.py or the Fortran ReadIni) and returns like_id = position in the list
 * (TLikelihoodList order, source/GeneralTypes.f90:129-146). */

/* native plik-lite (source/CMB.f90:208-329): nb[3] bins used for TT,TE,EE (first nb[i] bins of the common table),
 * blmin/blmax absolute l per bin, weights [lmax_w+1] indexed by l and already x 2pi/(l(l+1)),
 * invcov [nused][nused], x_data [nused].  nuisance: calPlanck. */
int cb200_like_add_pliklite(cb200_handle* h, const int* nb, int nbins_tab, const int* blmin, const int* blmax,
                            const double* weights, int lmax_w, const double* invcov, const double* x_data,
                            int cal_index /* position of calPlanck in the nuisance vector, -1 none */,
                            int* like_id);

/* binned CMBLikes (source/CMBlikes.f90:1165-1256), dense form:
 *   binned[bin][c] = sum_{X,l} W[bin][c][X][l] * Cl_X(l)/cal_X - offset[bin][c]   (X over TT,TE,EE,BB,PP;
 *   cal_X = calPlanck^2 for CMB spectra, 1 for PP; window + linear-correction windows folded into W,
 *   FiducialCorrection into offset; CMBlikes.f90:981-995)
 *   then per bin: C(+noise) -> gaussian (C-Chat) or HL transform -> vecp[cl_use] -> bigX; chi2 = X^T invcov X
 *   (+ (ln cal / prior)^2 when log_cal_prior>0).   like_approx: 1 HL, 2 gaussian.
 *   W [nbins][ncl][5][lmax_w+1]; noise/chat/sqrt_fid [nbins][nmaps][nmaps] (noise, sqrt_fid may be NULL). */
int cb200_like_add_cmblikes(cb200_handle* h, int nmaps, int nbins, int ncl_used, const int* cl_use_index,
                            int like_approx, int lmax_w, const double* W, const double* offset,
                            const double* noise, const double* chat, const double* sqrt_fid,
                            const double* invcov, double log_cal_prior,
                            int cal_index /* position of the calibration parameter in the nuisance vector, -1 none */,
                            int* like_id);

/* BICEP/Keck foreground model attached to a registered cmblikes likelihood (TBK_planck extends TCMBLikes and
 * overrides AddForegrounds, source/CMB_BK_Planck.f90:21-31,229-340): dust + synchrotron + correlated component with
 * bandpass-integrated frequency scalings (:109-183), band-centre errors and optional decorrelation (:187-227).
 *   map_field [nmaps] 0-based theory field of each used map (1 = E, 2 = B); bc_class [nmaps] 0 none / 1 '95' / 2 '150' /
 *   3 '220' (which gamma_* parameter applies, :266-274); bandpasses concatenated: bp_offset [nmaps+1], bp_nu / bp_R /
 *   bp_dnu; th_dust, th_sync, nu_bar [nmaps] from TBK_planck_Read_Bandpass (:72-105);
 *   lform_*: 0 flat, 1 lin, 2 quad; fgW [nbins][ncl][lmax+1] the band-power window of every map pair (zero for pairs
 *   that are not EE or BB); nuis_offset: position of BBdust (first of the 16 BK15.paramnames) in the nuisance vector. */
int cb200_like_set_bk_foregrounds(cb200_handle* h, int like_id, int nmaps, const int* map_field, const int* bc_class,
                                  const int* bp_offset, const double* bp_nu, const double* bp_R, const double* bp_dnu,
                                  const double* th_dust, const double* th_sync, const double* nu_bar, double fpivot_dust,
                                  double fpivot_sync, const double* fpivot_dust_decorr, const double* fpivot_sync_decorr,
                                  int lform_dust, int lform_sync, int lmin, int lmax, const double* fgW, int nuis_offset);

/* ---- background functions and background-only likelihoods ---------------------------------------------
 * bg [npts][16] per point = H0, omegab, omegac, omegan, omegav, w, tcmb, nu_massless_degeneracy, n_eigenstates,
 *   nu_mass_degeneracies[3], nu_mass_fractions[3], rdrag  — i.e. what CAMBCalc_CMBToCAMB + CAMBParams_Set hold in
 *   CP / grho* after SetParamsForBackground (source/Calculator_CAMB.f90:84-129,151-177; camb/modules.f90:300-375),
 *   plus r_drag from the thermal history (Theory%derived_parameters(derived_rdrag), source/bao.f90:237-248).
 *   cosmomc_b200/params.py builds it from CosmoMC's parameters.
 *
 * cb200_background replaces TCosmologyCalculator%AngularDiameterDistance / Hofz / ComovingRadialDistance /
 *   CMBToTheta and CAMB's DeltaPhysicalTimeGyr (source/Calculator_Cosmology.f90:17-39; camb/modules.f90:519-751):
 *   DA, H [npts][nz] (Mpc, Mpc^-1; NULL to skip), scalars [npts][3] = tau0, age/Gyr, CosmomcTheta (NULL to skip). */
int cb200_background(cb200_handle* h, int npts, const double* bg, int nz, const double* z, double* DA, double* H,
                     double* scalars);
/* ---- non-linear lensing rescale and sigma_8 ---------------------------------------------------
 * cb200_nonlinear_lensing replaces, for a batch, CAMB's MakeNonlinearSources (camb/cmbmain.f90:1145-1204) with the
 * halofit ratios of NonLinear_GetNonLinRatios (camb/halofit_ppf.f90:96-352, Takahashi 2012) on the matter power table of
 * Transfer_GetMatterPowerData (camb/modules.f90:1882-2074), and Transfer_Get_SigmaR at 8 Mpc/h (:2202-2268), i.e. what
 * happens to Src(k, 3, tau) between the Boltzmann ODEs and cb200_powers when use_nonlinear_lensing = T, plus the sigma_8
 * of the chain's derived block.  Call it after cb200_upload_sources and before cb200_powers (it multiplies the resident
 * lensing source of points [first, first+npts) in place: once per upload).
 *   initpower [npts][10] as for cb200_powers; cosmo [npts][6] = h, Omega_c + Omega_b + Omega_nu, Omega_Lambda,
 *   Omega_nu / Omega_m, w, wa; kh [npts][n_kt] the transfer wavenumbers in h/Mpc (the first n_k of them are the source
 *   wavenumbers of the point, as in CAMB where MT%q_trans starts with Evolve_q); z [n_z] the NLL redshifts, descending;
 *   transfer [npts][n_z][n_kt] = MT%TransferData(transfer_power_var, k, z); tautf [npts][n_z] their conformal times
 *   (needed when rescale_sources != 0).  Outputs (NULL to skip): sigma8 [npts][n_z], ratio [npts][n_z][n_kt] =
 *   sqrt(P_NL / P_L), spec [npts][n_z][3] = k_NL, n_eff, curvature, status [npts] (349 = halofit's "totally crazy
 *   non-linear" exit, global_error_flag of the reference). */
int cb200_nonlinear_lensing(cb200_handle* h, int first, int npts, const double* initpower, const double* cosmo, int n_kt,
                            int n_z, const double* kh, const double* z, const double* transfer, const double* tautf,
                            int rescale_sources, double* sigma8, double* ratio, double* spec, int* status);

/* ---- thermal history ---------------------------------------------------------------------
 * cb200_thermo replaces, for a batch, what CAMB does between CAMBParams_Set and the source ODEs
 * (camb/modules.f90:376-400 Nnow / akthom / adotrad, camb/reionization.f90:139-199 Reionization_Init incl. the
 * z_re-from-optical-depth bisection that CAMB_GetZreFromTau drives (source/Calculator_CAMB.f90 GetZreFromTau),
 * camb/recfast.f90:460-722 Recombination_init, camb/modules.f90:2682-2992 inithermo) and returns what the rest of the
 * path consumes: the time-grid scalars of cb200_upload_sources' `thermo` rows and ThermoDerivedParams
 * (Theory%derived_parameters: r_drag for source/bao.f90:237-248, z_star, theta_star, ... of the chain's derived block).
 *   bg         [npts][16]  as for cb200_background
 *   thermo_in  [npts][8]   Y_He, z_re (used when optical depth = 0), optical depth tau (> 0: z_re by bisection),
 *                          Max_eta_k, WantTensors, Transfer kmax in h/Mpc (0: WantTransfer = F; CosmoMC with
 *                          use_nonlinear_lensing: 5), AccuracyBoost (0 = 1), reserved
 *   thermo_out [npts][32]  tau0, taurst, taurend, reionisation tau_start, tau_complete, dtaurec, tau_maxvis, z_re, z_star,
 *                          z_drag, actual_opt_depth, status, then ThermoDerivedParams(1:13) = age/Gyr, zstar, rstar,
 *                          100 thetastar, DAstar/Gpc, zdrag, rdrag, kD, 100 thetaD, zEQ, kEQ, 100 thetaEQ, 100 theta_rs_EQ
 *   status     [npts]      0, 1 = error_reionization, 2 = error_recombination (camb/constants.f90 Errors); may be NULL */
int cb200_thermo(cb200_handle* h, int npts, const double* bg, const double* thermo_in, double* thermo_out, int* status);
/* theta -> H0: the bisection of ThetaParameterization%ParamArrayToTheoryParams on CMBToTheta
 * (source/CosmologyParameterizations.f90:134-176, camb/modules.f90:729-751 CosmomcTheta).
 *   cosmo [npts][8] = ombh2, omch2, omnuh2, omk, w, 100 theta_MC, H0_min, H0_max ; nu [npts][8] = bg[7..15) (the neutrino
 *   split, independent of H0) ; bg [npts][16] in: only bg[15] (rdrag) is kept, out: the solved rows (all zero but bg[15]
 *   when theta is out of range: H0 = 0, the point is rejected as in the reference). */
int cb200_theta_to_background(cb200_handle* h, int npts, const double* cosmo, const double* nu, double tcmb, double* bg);

/* make bg resident for points [first, first+npts): input of the background likelihoods in cb200_loglike_batch
 * (replaces Calculator%SetParamsForBackground / GetNewBackgroundData, source/Calculator_CAMB.f90:151-177) */
int cb200_set_background(cb200_handle* h, int first, int npts, const double* bg);

/* BAO (source/bao.f90:113-308): kind 0 = TBAOLikelihood (quadratic form with invcov [num_bao][num_bao]),
 * kind 1 = MGSLikelihood (tabulated chi^2(alpha), :390-410).  type[i]: 1-based measurement type codes of
 * bao.f90:29-35 (1 Az, 2 DV_over_rs, 3 rs_over_DV, 4 DA_over_rs, 5 F_AP, 7 bao_Hz_rs, 8 bao_Hz_rs_103,
 * 10 DM_over_rs).  fixed_rs > 0: BAO_fixed_rs. */
int cb200_like_add_bao(cb200_handle* h, int kind, int num_bao, const int* type, const double* z, const double* obs,
                       const double* invcov, double rs_rescale, double fixed_rs, const double* alpha_prob, int n_alpha,
                       int* like_id);
/* HST (source/HST.f90:47-59): (H0 - H0_obs)^2 / 2 sigma^2, or via D_A(zeff) when zeff > 0 */
int cb200_like_add_hst(cb200_handle* h, double H0, double H0_err, double zeff, double angconversion, int* like_id);
/* JLA / Pantheon (source/supernovae_JLA.f90:773-866,1028-1228).
 *   cols [11][nsn] = zcmb, zhel, mag, stretch, colour, pre_vars (jla_prep :911-920), stretch_var, colour_var,
 *                    cov_mag_stretch, cov_mag_colour, cov_stretch_colour
 *   A1, A2 [nsn] masks of the two-scriptM fit (ignored when !twoscriptm)
 *   cov[6] = mag, stretch, colour, mag_stretch, mag_colour, stretch_colour covariance blocks [nsn][nsn] (NULL: absent)
 *   alpha_index / beta_index: positions of alpha, beta in the nuisance vector (-1: fixed to 0) */
int cb200_like_add_sn(cb200_handle* h, int nsn, const double* cols, const double* A1, const double* A2, int twoscriptm,
                      const double* const* cov, int alpha_index, int beta_index, int* like_id);

/* -lnL of every registered likelihood for points [first, first+npts) using the Cls resident on the device
 * from the last cb200_powers (replaces TheoryLike_LogLikeWithTheorySet, source/calclike.f90:357-389).
 *   nuisance [npts][n_nuis_total] in registration order; loglikes [npts][n_like]; total [npts] */
int cb200_loglike_batch(cb200_handle* h, int first, int npts, const double* nuisance, int n_nuis,
                        double* loglikes, double* total, int* status);

/* same for host-supplied Cls (cls [npts][5][lmax_out+1]); used by the ctypes shim / importance sampling */
int cb200_loglike_cls(cb200_handle* h, int npts, const double* cls, const double* nuisance, int n_nuis,
                      double* loglikes, double* total, int* status);

/* ---- batch entry: bounds + priors + theory + likelihoods ------------------------------------------------
 * The reference evaluates ONE parameter point per call (TLikeCalculator%GetLogLike, source/calclike.f90:136-151);
 * this is the same control flow for a batch of points whose sources are resident (cb200_upload_sources):
 *   hard bounds (GetLogLikeBounds :97-109) -> logZero = 1e30, else sum of the registered likelihoods / temperature
 *   (TheoryLike_GetLogLikeMain :293-318, AddLikeTemp :80-94; any soft error or logZero likelihood -> logZero) +
 *   Gaussian and linear-combination priors / temperature (GetLogPriors :111-134).
 * The columns of `params` that feed the initial power spectrum follow CAMBCalc_SetCAMBInitPower
 * (source/Calculator_CAMB.f90:839-877): A_s = 1e-10 exp(logA); a column index of -1 takes the def_* value. */
typedef struct cb200_param_layout {
  int num_params;
  const double* pmin;           /* [num_params] BaseParams%PMin */
  const double* pmax;           /* [num_params] BaseParams%PMax */
  const double* prior_mean;     /* [num_params] GaussPriors%mean (NULL: no Gaussian priors) */
  const double* prior_std;      /* [num_params] GaussPriors%std, 0 = none */
  const unsigned char* use_prior; /* [num_params] varying(i) .or. include_fixed_parameter_priors; NULL = all */
  int n_lincomb;                /* BaseParams%LinearCombinations */
  const double* lincomb;        /* [n_lincomb][num_params] */
  const double* lincomb_mean;   /* [n_lincomb] */
  const double* lincomb_std;    /* [n_lincomb], 0 = none */
  double temperature;           /* TLikeCalculator%Temperature */
  int i_logA, i_ns, i_nrun, i_nrunrun, i_r, i_nt, i_ntrun, i_Alens, i_Aphiphi;
  double def_logA, def_ns, def_nrun, def_nrunrun, def_r, def_nt, def_ntrun, def_Alens, def_Aphiphi;
  double pivot_scalar, pivot_tensor;
  int inflation_consistency;
  int i_nuis_first, n_nuis;     /* nuisance parameters of the likelihoods = params[:, i_nuis_first : +n_nuis] */
} cb200_param_layout;
/* Change mask (Cosmo_CalculateRequiredTheoryChanges, source/CalcLike_Cosmology.f90:59-94), per point: the spectra are
 * recomputed only if the point's sources were re-uploaded since its last evaluation (slow change) or its initial-power /
 * ALens / Aphiphi values differ from the ones its resident spectra were computed with (semi-slow change); a point whose
 * nuisance parameters alone moved (fast change) goes straight to the likelihoods, and an out-of-bounds point is never
 * evaluated.  A direct cb200_powers / cb200_powers_shared call on a slot invalidates its entry.
 * params [npts][num_params]; loglike [npts] (-ln L incl. priors, 1e30 = rejected); likelihoods [npts][n_like] (may be
 * NULL); prior [npts] (may be NULL; the un-tempered prior term); status [npts] 0 ok / 1 out of bounds / >1 soft error */
int cb200_eval_batch(cb200_handle* h, const cb200_param_layout* layout, int first, int npts, const double* params,
                     double* loglike, double* likelihoods, double* prior, int* status);
/* TLikeCalculator_TestLikelihoodFunction (source/calclike.f90:180-199): -ln L = (x - c)^T covinv (x - c) / 2 for
 * x [npts][n]; used by `test_likelihood = T` runs and by the adaptive-MCMC tests (one DMMA GEMM + row dots) */
int cb200_test_like_batch(cb200_handle* h, int npts, int n, const double* x, const double* center, const double* covinv,
                          double* loglike);

/* ---- timing / counters (DebugMsgs timings of camb/cmbmain.f90:152-165,265-269, lensing.f90:516) ------- */
typedef struct cb200_timing {
  float ms_spline, ms_project, ms_contract, ms_interp, ms_lens, ms_like, ms_total;
  long long n_launches;   /* kernels launched by this library since the last reset */
  long long proj_triples; /* (q,l,tau) triples integrated by the last projection (if counting enabled) */
  long long ring_slabs, ring_direct, ring_rows, ring_pairs; /* windowed projection statistics (option "ring_stats") */
  long long phase_cycles[6]; /* per-warp clock64 sums: prologue, prefetch, barrier wait, ring fill, accumulate, metadata */
  float ms_background;       /* K5 distance kernels */
  long long proj_mask_mismatch; /* proj_kernel 4 with "ring_stats": (pair, lane) activity masks that differ from the exact windows; must be 0 */
  long long eval_points_powers; /* cb200_eval_batch: points whose spectra were recomputed (slow or semi-slow change) ... */
  long long eval_points_reused; /* ... and points whose resident spectra were reused (only nuisance parameters moved) */
} cb200_timing;
int cb200_get_timing(cb200_handle* h, cb200_timing* t, int reset);
int cb200_sync(cb200_handle* h);
/* options: "async_upload" (0/1: cb200_upload_sources returns with the source copy in flight on a separate stream, so
 * that the upload of one block of points overlaps cb200_powers of the previous block; the source buffer must then
 * stay untouched until cb200_sync or a call that returns results), "count_triples" (0/1), "keep_transfers" (0/1), "ring_stats" (0/1), "proj_kernel" (1 = L2 gathers, 2 = windowed warp-per-pair,
 * 3 = quarter-warp pairs per 32-multipole chunk, 4 = all multipoles per quarter-warp, producer/consumer warps; default),
 * "sn_chunk" (points per launch of the supernova kernels, default 1024; 4.4 MB of work matrix per point),
 * "sn_chol_warps" (warps per CTA of the per-point Cholesky: 8, 4, or 0 = by launch size; default 0) */
int cb200_set_option(cb200_handle* h, const char* name, double value);
/* CUDA-event stopwatch on the library's stream (device-side timing of whole calls) */
int cb200_timer_start(cb200_handle* h);
int cb200_timer_stop(cb200_handle* h, float* ms);
/* measured FP64 peaks of this GPU (TFLOP/s): vector DFMA and tensor-pipe DMMA micro-kernels; the driver-written
 * MEASURED_PEAKS.json holds only HBM and bf16 numbers, so the FP64 roofline denominators are measured here. */
int cb200_measure_fp64_peaks(cb200_handle* h, double* dfma_tflops, double* dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif
