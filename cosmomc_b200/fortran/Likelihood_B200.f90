    !Likelihood_B200: TDataLikelihood plug-ins whose LogLike runs on the B200 library.
    !
    !Pattern (source/Likelihood_Cosmology.f90:35-85): CosmoMC calls  like%GetLogLike(Params, Theory, DataParams), which type-casts
    !and calls the overridable  LogLike(CMB, Theory, DataParams).  The types below extend the reference's own likelihood
    !classes, so data-set parsing (ReadIni, ReadClArr, ReadBinWindows, ReadCovmat: source/CMBlikes.f90:371-859,
    !source/CMB.f90:208-303) stays the reference's; once the arrays are read they are resolved to the dense form of
    !include/cosmob200.h and registered with the calculator's handle, and LogLike becomes one call with npts = 1 that uses
    !the C_l the calculator left RESIDENT on the device (cb200_loglike_batch), so no spectrum crosses PCIe per likelihood.
    !
    !Selection: in CMBLikelihood_Add (source/CMB.f90:85-100) allocate TB200CMBLikes / TB200PlikLite instead of TCMBLikes /
    !TPlikLiteLikelihood when  cosmology_calculator = B200  (two lines, shown in INTEGRATION.md section 4).
    !
    !NOT compiled in this repository (no Fortran compiler in the build image or on the GPU box, see
    !profiles/r02_probe_gpu_host_no_fortran.log); written against the reference interfaces at the lines cited.

    module Likelihood_B200
    use, intrinsic :: iso_c_binding
    use settings
    use CosmologyTypes
    use CosmoTheory
    use Likelihood_Cosmology
    use CMBLikes
    use CMBLikelihoods, only : TPlikLiteLikelihood
    use Calculator_B200, only : b200_shared_handle, invalidate_like_cache
    implicit none
    private

    interface
    integer(c_int) function cb200_like_add_cmblikes(h, nmaps, nbins, ncl_used, cl_use_index, like_approx, lmax_w, W, &
        offset, noise, chat, sqrt_fid, invcov, log_cal_prior, cal_index, like_id) bind(C, name='cb200_like_add_cmblikes')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: nmaps, nbins, ncl_used, like_approx, lmax_w, cal_index
    integer(c_int), intent(in) :: cl_use_index(*)
    real(c_double), intent(in) :: W(*), offset(*), chat(*), invcov(*)
    type(c_ptr), value :: noise, sqrt_fid          !c_null_ptr when the data set has none
    real(c_double), value :: log_cal_prior
    integer(c_int), intent(out) :: like_id
    end function
    integer(c_int) function cb200_like_add_pliklite(h, nb, nbins_tab, blmin, blmax, weights, lmax_w, invcov, x_data, &
        cal_index, like_id) bind(C, name='cb200_like_add_pliklite')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), intent(in) :: nb(3), blmin(*), blmax(*)
    integer(c_int), value :: nbins_tab, lmax_w, cal_index
    real(c_double), intent(in) :: weights(*), invcov(*), x_data(*)
    integer(c_int), intent(out) :: like_id
    end function
    integer(c_int) function cb200_loglike_batch(h, first, npts, nuisance, n_nuis, loglikes, total, status) &
        bind(C, name='cb200_loglike_batch')
    import :: c_ptr, c_int, c_double
    type(c_ptr), value :: h
    integer(c_int), value :: first, npts, n_nuis
    real(c_double), intent(in) :: nuisance(*)
    real(c_double), intent(out) :: loglikes(*), total(*)
    integer(c_int), intent(out) :: status(*)
    end function
    end interface

    !Every B200 likelihood registers with the calculator's handle; cb200_loglike_batch evaluates ALL registered
    !likelihoods of a point in one pass, so the first LogLike call of a point fills this cache and the others read it.
    integer, parameter :: max_b200_likes = 32
    integer :: n_b200_likes = 0, n_b200_nuis = 0
    real(c_double) :: cached_loglikes(max_b200_likes)
    real(c_double), allocatable :: cached_nuis(:)
    logical :: cache_valid = .false.

    Type, extends(TCMBLikes) :: TB200CMBLikes
        integer(c_int) :: like_id = -1
        integer :: nuis_offset = 0                 !position of this likelihood's DataParams in the library's nuisance vector
    contains
    procedure :: ReadIni => B200CMBLikes_ReadIni
    procedure :: LogLike => B200CMBLikes_LogLike
    end Type TB200CMBLikes

    Type, extends(TPlikLiteLikelihood) :: TB200PlikLite
        integer(c_int) :: like_id = -1
        integer :: nuis_offset = 0
    contains
    procedure :: ReadIni => B200PlikLite_ReadIni
    procedure :: LogLike => B200PlikLite_LogLike
    end Type TB200PlikLite

    public TB200CMBLikes, TB200PlikLite, B200_InvalidateLikeCache
    contains

    subroutine B200_InvalidateLikeCache()
    !called by B200_GetNewPowerData: new C_l are resident, cached -lnL are stale
    cache_valid = .false.
    end subroutine

    integer function theory_slot(fi, fj)
    !device C_l layout [TT, TE, EE, BB, PP] for a pair of theory fields (CL_T=1, CL_E=2, CL_B=3, CL_Phi=4); 0 = never allocated
    integer, intent(in) :: fi, fj
    integer a, b
    a = min(fi, fj); b = max(fi, fj)
    theory_slot = 0
    if (a==CL_T .and. b==CL_T) theory_slot = 1
    if (a==CL_T .and. b==CL_E) theory_slot = 2
    if (a==CL_E .and. b==CL_E) theory_slot = 3
    if (a==CL_B .and. b==CL_B) theory_slot = 4
    if (a==CL_Phi .and. b==CL_Phi) theory_slot = 5
    end function

    subroutine fold_windows(this, BW, W, lmax_w)
    !dense W(l, spec, cl, bin) += window * [spectrum slot of (ix_in)] ; TBinWindows_bin (source/CMBlikes.f90:1230-1256)
    class(TCMBLikes) :: this
    Type(TBinWindows), intent(in) :: BW
    integer, intent(in) :: lmax_w
    real(c_double), intent(inout) :: W(0:lmax_w, 5, this%ncl, this%nbins_used)
    integer win_ix, ix_in(2), ix_out, slot, bin, fi, fj
    do win_ix = 1, size(BW%bin_cols_in, 2)
        ix_in = BW%bin_cols_in(:, win_ix)
        ix_out = BW%bin_cols_out(win_ix)
        if (ix_out <= 0) cycle
        fi = this%map_fields(this%required_order(ix_in(1)))
        fj = this%map_fields(this%required_order(ix_in(2)))
        slot = theory_slot(fi, fj)
        if (slot == 0) cycle                          !spectrum the theory never allocates (TB, EB): contributes zero
        do bin = 1, this%nbins_used
            W(BW%lmin:BW%lmax, slot, ix_out, bin) = W(BW%lmin:BW%lmax, slot, ix_out, bin) + &
                BW%W(:, win_ix, this%bin_min + bin - 1)
        end do
    end do
    end subroutine

    subroutine B200CMBLikes_ReadIni(this, Ini)
    class(TB200CMBLikes) :: this
    class(TSettingIni) :: Ini
    real(c_double), allocatable, target :: W(:,:,:,:), offset(:,:), chat(:,:,:), noise(:,:,:), sfid(:,:,:)
    type(c_ptr) :: pnoise, psfid
    integer bin, lmax_w, nb

    call this%TCMBLikes%ReadIni(Ini)                 !the reference parses the .dataset (source/CMBlikes.f90:371-750)
    if (.not. this%binned) call MpiStop('B200: unbinned CMBlikes data sets are not on the GPU path')
    if (this%like_approx == like_approx_fullsky_exact) call MpiStop('B200: exact full-sky likelihood is not binned')
    if (this%has_foregrounds) call MpiStop('B200: use TB200BKPlanck for data sets with foreground models')
    if (this%aberration_coeff /= 0) call MpiStop('B200: aberration_coeff is not folded into the dense windows')
    nb = this%nbins_used
    lmax_w = this%pcl_lmax
    allocate(W(0:lmax_w, 5, this%ncl, nb), offset(this%ncl, nb), chat(this%nmaps, this%nmaps, nb))
    W = 0; offset = 0
    call fold_windows(this, this%binWindows, W, lmax_w)
    if (allocated(this%binCorrectionWindows%W)) then  !GetBinnedMapCls, source/CMBlikes.f90:981-995
        call fold_windows(this, this%binCorrectionWindows, W, lmax_w)
        do bin = 1, nb
            offset(:, bin) = this%FiducialCorrection(:, this%bin_min + bin - 1)
        end do
    end if
    pnoise = c_null_ptr; psfid = c_null_ptr
    do bin = 1, nb
        chat(:,:,bin) = this%ChatM(this%bin_min + bin - 1)%M
    end do
    if (allocated(this%NoiseM)) then
        allocate(noise(this%nmaps, this%nmaps, nb))
        do bin = 1, nb
            noise(:,:,bin) = this%NoiseM(this%bin_min + bin - 1)%M
        end do
        pnoise = c_loc(noise)
    end if
    if (allocated(this%sqrt_fiducial)) then
        allocate(sfid(this%nmaps, this%nmaps, nb))
        do bin = 1, nb
            sfid(:,:,bin) = this%sqrt_fiducial(this%bin_min + bin - 1)%M
        end do
        psfid = c_loc(sfid)
    end if
    !Fortran W(l, spec, cl, bin) is C W[bin][cl][spec][l]; symmetric [nmaps][nmaps] blocks need no transpose
    this%nuis_offset = n_b200_nuis
    invalidate_like_cache => B200_InvalidateLikeCache
    if (cb200_like_add_cmblikes(b200_shared_handle(), int(this%nmaps, c_int), int(nb, c_int), int(this%ncl_used, c_int), &
        int(this%cl_use_index - 1, c_int), int(this%like_approx, c_int), int(lmax_w, c_int), W, offset, pnoise, chat, &
        psfid, this%inv_covariance, real(this%log_calibration_prior, c_double), &
        int(merge(this%nuis_offset + this%calibration_index - 1, -1, this%calibration_index > 0), c_int), &
        this%like_id) /= 0) call MpiStop('B200: cb200_like_add_cmblikes failed')
    n_b200_nuis = n_b200_nuis + this%nuisance_params%nnames
    n_b200_likes = n_b200_likes + 1
    end subroutine B200CMBLikes_ReadIni

    function b200_loglike_of(like_id, nuis_offset, DataParams) result(LogLike)
    !one pass over ALL registered likelihoods for the point whose C_l are resident (slot 0); cached until new C_l arrive
    integer(c_int), intent(in) :: like_id
    integer, intent(in) :: nuis_offset
    real(mcp), intent(in) :: DataParams(:)
    real(mcp) LogLike
    real(c_double) :: total(1)
    integer(c_int) :: st(1)

    if (.not. allocated(cached_nuis)) then
        allocate(cached_nuis(max(1, n_b200_nuis)))
        cached_nuis = 0
    end if
    if (size(DataParams) > 0) then
        if (any(cached_nuis(nuis_offset+1:nuis_offset+size(DataParams)) /= DataParams)) cache_valid = .false.
        cached_nuis(nuis_offset+1:nuis_offset+size(DataParams)) = DataParams
    end if
    if (.not. cache_valid) then
        if (cb200_loglike_batch(b200_shared_handle(), 0_c_int, 1_c_int, cached_nuis, int(n_b200_nuis, c_int), &
            cached_loglikes, total, st) /= 0) then
            LogLike = logZero
            return
        end if
        cache_valid = (st(1) == 0)
        if (st(1) /= 0) then
            LogLike = logZero                            !soft error: the point is rejected, the run goes on
            return
        end if
    end if
    LogLike = cached_loglikes(like_id + 1)
    end function

    function B200CMBLikes_LogLike(this, CMB, Theory, DataParams) result(LogLike)
    !replaces CMBLikes_LogLike (source/CMBlikes.f90:1165-1227); Theory%Cls is not read: the same C_l are on the device
    class(TB200CMBLikes) :: this
    Class(CMBParams) CMB
    Class(TCosmoTheoryPredictions), target :: Theory
    real(mcp) DataParams(:)
    real(mcp) LogLike
    LogLike = b200_loglike_of(this%like_id, this%nuis_offset, DataParams)
    end function

    subroutine B200PlikLite_ReadIni(this, Ini)
    !TPlikLiteLikelihood_ReadIni (source/CMB.f90:208-303) reads blmin/blmax, the weights (already x 2pi/(l(l+1)), indexed
    !from plmin), X_data and the inverse covariance of the used bins; they go to the device as they are.  The library
    !takes "the first nb(i) bins of spectrum i" (the default selection); an L-range cut that keeps the lowest bins is
    !that; any other subset (bins_for_L_range with a lower cut) is refused.
    class(TB200PlikLite) :: this
    class(TSettingIni) :: Ini
    integer(c_int) :: nb(3)
    real(c_double), allocatable :: w(:)
    integer i, j, lmax_w
    call this%TPlikLiteLikelihood%ReadIni(Ini)
    nb = 0
    do i = 1, 3
        if (this%used(i)) then
            nb(i) = size(this%used_bins(i)%bins)
            do j = 1, nb(i)
                if (this%used_bins(i)%bins(j) /= j) call MpiStop('B200 plik-lite: only bin selections 1..n are supported')
            end do
        end if
    end do
    lmax_w = ubound(this%weights, 1)
    allocate(w(0:lmax_w))                      !library indexes the weights by l from 0
    w = 0
    w(this%plmin:lmax_w) = this%weights(this%plmin:lmax_w)
    this%nuis_offset = n_b200_nuis
    invalidate_like_cache => B200_InvalidateLikeCache
    if (cb200_like_add_pliklite(b200_shared_handle(), nb, int(size(this%blmin), c_int), int(this%blmin, c_int), &
        int(this%blmax, c_int), w, int(lmax_w, c_int), this%invcov, this%X_data, &
        int(this%nuis_offset, c_int), this%like_id) /= 0) call MpiStop('B200: cb200_like_add_pliklite failed')
    n_b200_nuis = n_b200_nuis + this%nuisance_params%nnames
    n_b200_likes = n_b200_likes + 1
    end subroutine

    function B200PlikLite_LogLike(this, CMB, Theory, DataParams) result(LogLike)
    !replaces TPlikLiteLikelihood_LogLike (source/CMB.f90:305-329)
    class(TB200PlikLite) :: this
    Class(CMBParams) CMB
    Class(TCosmoTheoryPredictions), target :: Theory
    real(mcp) DataParams(:)
    real(mcp) LogLike
    LogLike = b200_loglike_of(this%like_id, this%nuis_offset, DataParams)
    end function

    end module Likelihood_B200
