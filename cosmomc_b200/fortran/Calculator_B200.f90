    !Calculator_B200: CosmoMC calculator plug-in that keeps CAMB for the Boltzmann source ODEs and hands everything
    !after them (source spline, line-of-sight projection, k-contraction, l-interpolation, lensing, unit conversion)
    !to the B200 library through the ISO_C_BINDING interfaces below (include/cosmob200.h).
    !
    !Drop-in pattern: same as source/Calculator_PICO.f90:18-27 (extend CAMB_Calculator, override the two slow-step
    !procedures), selected with  cosmology_calculator = B200  by the branch shown in INTEGRATION.md
    !(source/CosmologyConfig.f90:40-52).
    !
    !NOT compiled in this repository's CI: neither the build container nor the GPU box has a Fortran compiler
    !(SURVEY section 0).  It is written against the reference's module interfaces as they stand at
    !source/Calculator_CAMB.f90:179-275 and camb/cmbmain.f90:90-121,198-263.

    module Calculator_B200
    use, intrinsic :: iso_c_binding
    use CosmologyTypes
    use CosmoTheory
    use CAMB, only : CAMB_GetResults, CAMBParams, CAMB_SetDefParams
    use Calculator_CAMB
    use settings
    implicit none
    private

    !-- mirror of cb200_config (include/cosmob200.h:25-48).  The C layout this type must reproduce is stated here and
    !   checked WITHOUT a Fortran compiler by tests/test_abi_layout.py (gcc _Static_assert on the header against these
    !   numbers, and a field-by-field comparison of this declaration with the header):
    !   ABI-LAYOUT sizeof=96 struct_size=0 device=4 lmax_computed_cl=8 cmb_lensing=12 use_lensing_potential=16
    !   ABI-LAYOUT use_nonlinear_lensing=20 compute_tensors=24 lmax_tensor=28 accurate_bb=32 k_eta_max_scalar=40
    !   ABI-LAYOUT accuracy_level=48 lmax_out=56 highl_norm_first_call=60 max_points=64 chunk_points=68 n_tau_max=72
    !   ABI-LAYOUT n_k_max=76 n_q_max=80 n_tau_max_tensor=84 n_k_max_tensor=88 n_q_max_tensor=92
    !   cb200_default_config writes struct_size = 96; cb200_create returns -3 for any other value.
    type, bind(C) :: cb200_config
        integer(c_int) :: struct_size, device, lmax_computed_cl, cmb_lensing, use_lensing_potential, use_nonlinear_lensing
        integer(c_int) :: compute_tensors, lmax_tensor, accurate_bb
        real(c_double) :: k_eta_max_scalar, accuracy_level
        integer(c_int) :: lmax_out, highl_norm_first_call, max_points, chunk_points, n_tau_max, n_k_max, n_q_max
        integer(c_int) :: n_tau_max_tensor, n_k_max_tensor, n_q_max_tensor
    end type cb200_config

    interface
    subroutine cb200_default_config(cfg) bind(C, name='cb200_default_config')
    import :: cb200_config
    type(cb200_config), intent(out) :: cfg
    end subroutine
    integer(c_int) function cb200_create(cfg, h) bind(C, name='cb200_create')
    import :: cb200_config, c_ptr, c_int
    type(cb200_config), intent(in) :: cfg
    type(c_ptr), intent(out) :: h
    end function
    subroutine cb200_destroy(h) bind(C, name='cb200_destroy')
    import :: c_ptr
    type(c_ptr), value :: h
    end subroutine
    integer(c_int) function cb200_set_templates(h, unlensed, lensed, n_l) bind(C, name='cb200_set_templates')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    real(c_double), intent(in) :: unlensed(*), lensed(*)
    integer(c_int), value :: n_l
    end function
    integer(c_int) function cb200_upload_sources(h, kind, first, npts, thermo, n_k, k, src, src_is_device) &
        bind(C, name='cb200_upload_sources')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: kind, first, npts, src_is_device
    real(c_double), intent(in) :: thermo(*), k(*), src(*)
    integer(c_int), intent(in) :: n_k(*)
    end function
    integer(c_int) function cb200_upload_sources_packed(h, kind, first, npts, thermo, n_tau, n_k, k, src_packed) &
        bind(C, name='cb200_upload_sources_packed')
    !sources at their exact sizes: CAMB's Src(1:n_k, 1:3, 1:n_tau) IS the packed layout (no padding copy on this side)
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: kind, first, npts
    real(c_double), intent(in) :: thermo(*), k(*), src_packed(*)
    integer(c_int), intent(in) :: n_tau(*), n_k(*)
    end function
    integer(c_int) function cb200_powers(h, first, npts, initpower, alens, aphiphi, cls_out, derived_out, status) &
        bind(C, name='cb200_powers')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: first, npts
    real(c_double), intent(in) :: initpower(*), alens(*), aphiphi(*)
    real(c_double), intent(out) :: cls_out(*), derived_out(*)
    integer(c_int), intent(out) :: status(*)
    end function
    integer(c_int) function cb200_set_option(h, name, value) bind(C, name='cb200_set_option')
    !"async_upload" = 1: uploads overlap the evaluation of the previous block (INTEGRATION.md section 7)
    import :: c_ptr, c_int, c_double, c_char
    type(c_ptr), value :: h
    character(kind=c_char), intent(in) :: name(*)
    real(c_double), value :: value
    end function
    integer(c_int) function cb200_eval_batch(h, layout, first, npts, params, loglike, likelihoods, prior, status) &
        bind(C, name='cb200_eval_batch')
    !TLikeCalculator%GetLogLike (source/calclike.f90:97-151) for npts points at once; layout = cb200_param_layout
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h, layout
    integer(c_int), value :: first, npts
    real(c_double), intent(in) :: params(*)
    real(c_double), intent(out) :: loglike(*), likelihoods(*), prior(*)
    integer(c_int), intent(out) :: status(*)
    end function
    integer(c_int) function cb200_nonlinear_lensing(h, first, npts, initpower, cosmo, n_kt, n_z, kh, z, transfer, tautf, &
        rescale_sources, sigma8, ratio, spec, status) bind(C, name='cb200_nonlinear_lensing')
    !MakeNonlinearSources + halofit ratios + sigma_8 for npts points (camb/cmbmain.f90:1145-1204, halofit_ppf.f90:96-352,
    !modules.f90:2202-2268); transfer(n_kt, n_z, npts) = MT%TransferData(transfer_power_var, :, :) at the NLL redshifts
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: first, npts, n_kt, n_z, rescale_sources
    real(c_double), intent(in) :: initpower(*), cosmo(*), kh(*), z(*), transfer(*), tautf(*)
    type(c_ptr), value :: sigma8, ratio, spec, status
    end function
    integer(c_int) function cb200_thermo(h, npts, bg, thermo_in, thermo_out, status) bind(C, name='cb200_thermo')
    !thermal history of npts points (RECFAST, reionisation, inithermo: camb/modules.f90:2682-2992); thermo_out(13:25, i) =
    !ThermoDerivedParams, thermo_out(8, i) = z_re from the optical depth (what GetZreFromTau returns)
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: npts
    real(c_double), intent(in) :: bg(*), thermo_in(*)
    real(c_double), intent(out) :: thermo_out(*)
    integer(c_int), intent(out) :: status(*)
    end function
    integer(c_int) function cb200_theta_to_background(h, npts, cosmo, nu, tcmb, bg) bind(C, name='cb200_theta_to_background')
    !the H0 bisection of ThetaParameterization%ParamArrayToTheoryParams (source/CosmologyParameterizations.f90:134-176)
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: npts
    real(c_double), intent(in) :: cosmo(*), nu(*)
    real(c_double), value :: tcmb
    real(c_double), intent(inout) :: bg(*)
    end function
    end interface

    !the likelihood plug-ins (Likelihood_B200.f90) register with, and evaluate on, the calculator's handle
    type(c_ptr), save :: shared_handle = c_null_ptr
    abstract interface
    subroutine no_arg_sub()
    end subroutine
    end interface
    procedure(no_arg_sub), pointer, save :: invalidate_like_cache => null()

    Type, extends(CAMB_Calculator) :: B200_Calculator
        type(c_ptr) :: handle = c_null_ptr
        integer :: n_tau_max = 768, n_k_max = 256, n_tau_max_tensor = 2304, n_k_max_tensor = 128
    contains
    procedure :: InitForLikelihoods => B200_InitForLikelihoods
    procedure :: GetNewTransferData => B200_GetNewTransferData
    procedure :: GetNewPowerData => B200_GetNewPowerData
    procedure :: VersionTraceOutput => B200_VersionTraceOutput
    procedure :: GetZreFromTau => B200_GetZreFromTau
    end type B200_Calculator

    public B200_Calculator, b200_shared_handle, invalidate_like_cache
    contains

    function B200_GetZreFromTau(this, CMB, tau) result(zre)
    !TCosmologyCalculator%GetZreFromTau (source/Calculator_CAMB.f90 CAMBCalc_GetZreFromTau -> CAMB_GetZreFromTau ->
    !Reionization_zreFromOptDepth, camb/reionization.f90:256-291) through the batched thermal history with npts = 1.
    !A driver that owns many chains calls cb200_thermo once for all of them instead (INTEGRATION.md section 8).
    class(B200_Calculator) :: this
    class(CMBParams) :: CMB
    real(mcp), intent(in) :: tau
    real(mcp) :: zre
    type(CAMBParams) :: P
    real(c_double) :: bg(16), tin(8), tout(32)
    integer(c_int) :: st(1), rc

    call this%CMBToCAMB(CMB, P)
    bg = 0
    bg(1) = P%H0; bg(2) = P%omegab; bg(3) = P%omegac; bg(4) = P%omegan; bg(5) = P%omegav; bg(6) = -1; bg(7) = P%tcmb
    bg(8) = P%Num_Nu_massless; bg(9) = P%Nu_mass_eigenstates
    bg(10:9+P%Nu_mass_eigenstates) = P%Nu_mass_degeneracies(1:P%Nu_mass_eigenstates)
    bg(13:12+P%Nu_mass_eigenstates) = P%Nu_mass_fractions(1:P%Nu_mass_eigenstates)
    tin = [P%YHe, 0._c_double, real(tau, c_double), real(P%Max_eta_k, c_double), 0._c_double, &
        real(P%Transfer%kmax, c_double), 1._c_double, 0._c_double]
    rc = cb200_thermo(this%handle, 1_c_int, bg, tin, tout, st)
    if (rc /= 0 .or. st(1) /= 0) call MpiStop('B200: thermal history failed in GetZreFromTau')
    zre = tout(8)
    end function B200_GetZreFromTau

    function b200_shared_handle() result(h)
    type(c_ptr) :: h
    if (.not. c_associated(shared_handle)) call MpiStop('B200: likelihood registered before the calculator was initialised')
    h = shared_handle
    end function

    subroutine B200_InitForLikelihoods(this)
    !Called once after the likelihoods fixed CosmoSettings (source/Calculator_CAMB.f90:926-946)
    class(B200_Calculator) :: this
    type(cb200_config) :: cfg
    real(c_double), allocatable :: unl(:,:), lens(:,:)
    integer L

    call this%CAMB_Calculator%InitForLikelihoods()
    call cb200_default_config(cfg)
    cfg%lmax_computed_cl = CosmoSettings%lmax_computed_cl
    cfg%cmb_lensing = merge(1, 0, CosmoSettings%CMB_Lensing)
    cfg%use_lensing_potential = merge(1, 0, CosmoSettings%use_lensing_potential)
    cfg%use_nonlinear_lensing = merge(1, 0, CosmoSettings%use_nonlinear_lensing)
    cfg%compute_tensors = merge(1, 0, CosmoSettings%compute_tensors)
    cfg%lmax_tensor = CosmoSettings%lmax_tensor
    cfg%accurate_bb = merge(1, 0, this%accurate_BB)
    cfg%k_eta_max_scalar = this%k_eta_max_scalar
    cfg%accuracy_level = AccuracyLevel
    cfg%lmax_out = CosmoSettings%lmax
    cfg%highl_norm_first_call = 1          !keep the reference's SAVEd highL_norm behaviour bit for bit
    cfg%max_points = 1
    cfg%n_tau_max = this%n_tau_max
    cfg%n_k_max = this%n_k_max
    cfg%n_tau_max_tensor = this%n_tau_max_tensor
    cfg%n_k_max_tensor = this%n_k_max_tensor
    !cfg%struct_size was set by cb200_default_config; cb200_create returns -3 if this mirror is out of step with the header
    if (cb200_create(cfg, this%handle) /= 0) call MpiStop('B200: cb200_create failed (no CUDA device, or stale cb200_config mirror)')
    shared_handle = this%handle
    !templates: camb/modules.f90:1162-1185 (highL_CL_template) and source/Calculator_CAMB.f90:966-990
    allocate(unl(0:8000,4), lens(0:CosmoSettings%lmax,4))
    unl = 0; lens = 0
    call CheckLoadedHighLTemplate
    do L = lmin, 8000
        unl(L,1:4) = highL_CL_template(L, C_Temp:C_Phi)
    end do
    if (allocated(this%highL_lensedCL_template)) then
        do L = 2, CosmoSettings%lmax
            lens(L,1:4) = this%highL_lensedCL_template(L,1:4)
        end do
    end if
    if (cb200_set_templates(this%handle, unl, lens, CosmoSettings%lmax+1) /= 0) call MpiStop('B200: templates')
    end subroutine B200_InitForLikelihoods

    subroutine B200_GetNewTransferData(this, CMB, Info, Theory, error)
    !Slow step (source/Calculator_CAMB.f90:179-218): CAMB evolves the sources, then the library takes over.
    !Requires cosmomc_b200/fortran/camb_sources_only.patch (INTEGRATION.md section 3): with cmbmain_sources_only = .true.
    !cmbmain returns after the DoSourcek loop, TransferOut and MakeNonlinearSources (camb/cmbmain.f90:198-233), i.e. with
    !the non-linear lensing rescale already applied to Src, and leaves Src, Evolve_q allocated.  CAMB_GetResults runs one
    !cmbmain pass per perturbation type (camb/camb.f90:124-180); the hook returns before FreeSourceMem, so the passes are
    !driven one at a time here and the sources of each are handed over before the next pass allocates its own.
    use CAMBmain, only : Src, Evolve_q, SourceNum, cmbmain_sources_only, FreeSourceMem
    use ModelParams, only : CP, taurst, taurend        !camb/modules.f90:27-36,179 (module ModelParams is public)
    class(B200_Calculator) :: this
    class(CMBParams) CMB
    class(TTheoryIntermediateCache), pointer :: Info
    class(TCosmoTheoryPredictions) :: Theory
    integer error
    type(CAMBParams) P, P1
    integer pass

    select type (Info)
    class is (CAMBTransferCache)
        call this%CMBToCAMB(CMB, P)
        P%OnlyTransfers = .true.
        cmbmain_sources_only = .true.
        do pass = 0, merge(1, 0, CosmoSettings%compute_tensors)
            P1 = P
            if (pass == 0) then
                P1%WantTensors = .false.                       !scalar pass: also the matter transfer functions (sigma8)
            else
                P1%WantScalars = .false.; P1%WantTransfer = .false.
            end if
            call CAMB_GetResults(P1, error)
            if (error == 0) call upload_pass(int(pass, c_int))
            call FreeSourceMem
            if (error /= 0) exit
        end do
        cmbmain_sources_only = .false.
        if (error == 0) call this%SetDerived(Theory)
    end select

    contains

    subroutine upload_pass(kind)
    integer(c_int), intent(in) :: kind
    real(c_double) :: thermo(5)
    real(c_double), allocatable :: k(:)
    integer(c_int) :: nk(1), nt(1)
    thermo = [CP%tau0, taurst, taurend, &
        merge(CP%ReionHist%tau_start, -1._dl, CP%Reion%Reionization), CP%ReionHist%tau_complete]
    nk(1) = Evolve_q%npoints
    nt(1) = size(Src, 3)
    if (SourceNum /= 3) call MpiStop('B200: expects three sources per perturbation type (T, E, lensing potential / B)')
    allocate(k(merge(this%n_k_max_tensor, this%n_k_max, kind == 1)))   !k is padded to n_k_max; Src is not
    k = Evolve_q%points(nk(1))
    k(1:nk(1)) = Evolve_q%points(1:nk(1))
    !Src(k,s,tau), contiguous and exactly (n_k, 3, n_tau): C [n_tau][3][n_k], the packed layout of the header
    if (cb200_upload_sources_packed(this%handle, kind, 0_c_int, 1_c_int, thermo, nt, nk, k, Src) /= 0) error = 1
    end subroutine upload_pass

    end subroutine B200_GetNewTransferData

    subroutine B200_GetNewPowerData(this, CMB, Info, Theory, error)
    !Semi-slow step (source/Calculator_CAMB.f90:220-275 + SetPowersFromCAMB :349-463) on the GPU
    class(B200_Calculator) :: this
    class(CMBParams) :: CMB
    class(TTheoryIntermediateCache), pointer :: Info
    class(TCosmoTheoryPredictions) :: Theory
    integer error
    real(c_double) :: ip(10), alens(1), aphi(1), derived(4)
    real(c_double), allocatable :: cls(:,:)
    integer(c_int) :: st(1)
    integer lmx

    lmx = CosmoSettings%lmax
    allocate(cls(0:lmx, 5))
    ip = [cl_norm*CMB%InitPower(As_index), CMB%InitPower(ns_index), CMB%InitPower(nrun_index), &
        CMB%InitPower(nrunrun_index), CMB%InitPower(amp_ratio_index), CMB%InitPower(nt_index), &
        CMB%InitPower(ntrun_index), CosmoSettings%pivot_k, CosmoSettings%tensor_pivot_k, &
        merge(1._dl, 0._dl, CosmoSettings%inflation_consistency)]
    alens = CMB%ALens
    aphi = CMB%InitPower(Aphiphi_index)
    if (cb200_powers(this%handle, 0_c_int, 1_c_int, ip, alens, aphi, cls, derived, st) /= 0) then
        error = 1
        return
    end if
    error = st(1)
    call B200_InvalidateLikeCache_hook()          !new C_l are resident: Likelihood_B200's cached -lnL are stale
    if (error /= 0) return
    !order of cls columns: TT, TE, EE, BB, PP  ->  Theory%Cls(i,j)%CL   (source/CosmoTheory.f90:23-52)
    if (allocated(Theory%Cls(1,1)%CL)) Theory%Cls(1,1)%CL(2:) = cls(2:ubound(Theory%Cls(1,1)%CL,1), 1)
    if (allocated(Theory%Cls(2,1)%CL)) Theory%Cls(2,1)%CL(2:) = cls(2:ubound(Theory%Cls(2,1)%CL,1), 2)
    if (allocated(Theory%Cls(2,2)%CL)) Theory%Cls(2,2)%CL(2:) = cls(2:ubound(Theory%Cls(2,2)%CL,1), 3)
    if (allocated(Theory%Cls(3,3)%CL)) Theory%Cls(3,3)%CL(2:) = cls(2:ubound(Theory%Cls(3,3)%CL,1), 4)
    if (allocated(Theory%Cls(CL_Phi,CL_Phi)%CL)) &
        Theory%Cls(CL_Phi,CL_Phi)%CL(2:) = cls(2:ubound(Theory%Cls(CL_Phi,CL_Phi)%CL,1), 5)
    Theory%Lensing_rms_deflect = derived(1)
    end subroutine B200_GetNewPowerData

    subroutine B200_InvalidateLikeCache_hook()
    !Likelihood_B200 uses this module, so the dependency cannot point the other way: it sets this procedure pointer when
    !its first likelihood registers
    if (associated(invalidate_like_cache)) call invalidate_like_cache()
    end subroutine

    subroutine B200_VersionTraceOutput(this, ReadValues)
    use IniObjects
    class(B200_Calculator) :: this
    class(TNameValueList) :: ReadValues
    call this%CAMB_Calculator%VersionTraceOutput(ReadValues)
    call ReadValues%Add('Compiled_B200_projection', 'cosmob200 v1')
    end subroutine B200_VersionTraceOutput

    end module Calculator_B200
