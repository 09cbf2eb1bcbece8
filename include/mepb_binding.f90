! mepb_binding.f90 - ISO_C_BINDING interface to libmepb.so (include/mepb.h) for the reference's Fortran host.
!
! SOURCE ONLY: this image has no Fortran compiler, so the module has never been compiled here.  The C side of every symbol
! is exercised through ctypes (tests/test_abi.py, tests/test_gpu_parity.py), which binds the same ABI.  INTEGRATION.md shows
! how core_worker (src/core.f90:127-292) and band_worker (src/band_calc.f90:62-76) of the reference call it.
module mepb_binding
   use, intrinsic :: iso_c_binding
   implicit none
   type, bind(C) :: mepb_model                 ! include/mepb.h: mepb_model
      integer(c_int32_t) :: nw, nR
      type(c_ptr)        :: R                  ! R_vect(3,nR)              file_io.f90:724
      type(c_ptr)        :: H                  ! H_tb(nw,nw,nR) complex(dp) file_io.f90:732-741
      type(c_ptr)        :: r_pos              ! c_null_ptr (position operator unsupported)
      type(c_ptr)        :: wf_centers         ! wf_centers(3,nw) or c_null_ptr   wann_interp.f90:192
      real(c_double)     :: a_latt(9)          ! a_latt(3,3) as stored (column-major)
   end type
   type, bind(C) :: mepb_params                ! include/mepb.h: mepb_params  (globals of input_paras.f90:49-68)
      integer(c_int32_t) :: mp_grid(3)
      integer(c_int32_t) :: do_gauge_trafo, use_cart_velo
      integer(c_int32_t) :: do_mep, do_kubo, do_ahc, do_opt, do_write_mep_bands
      integer(c_int32_t) :: valence_bands, n_hw, n_eF
      real(c_double)     :: hw_min, hw_max, eF_min, eF_max       ! Hartree
      real(c_double)     :: T_kelvin, eta, kubo_tol
   end type
   type, bind(C) :: mepb_result                ! include/mepb.h: mepb_result
      integer(c_int64_t) :: n_k
      integer(c_int32_t) :: eF_idx             ! 0-based
      integer(c_int64_t) :: n_degenerate
      type(c_ptr) :: Ne
      type(c_ptr) :: mep_ic, mep_lc, mep_cs, mep_bands
      type(c_ptr) :: kubo_mep_ic, kubo_mep_lc, kubo_mep_cs
      type(c_ptr) :: ahc, ahc_velo, ohc_velo, opt_s, opt_a, photo2
   end type
   interface
      integer(c_int) function mepb_create(model, params, device, ctx) bind(C, name="mepb_create")
         import; type(mepb_model), intent(in) :: model; type(mepb_params), intent(in) :: params
         integer(c_int), value :: device; type(c_ptr), intent(out) :: ctx
      end function
      subroutine mepb_destroy(ctx) bind(C, name="mepb_destroy")
         import; type(c_ptr), value :: ctx
      end subroutine
      integer(c_int) function mepb_run_mesh(ctx, res) bind(C, name="mepb_run_mesh")
         import; type(c_ptr), value :: ctx; type(mepb_result), intent(inout) :: res
      end function
      integer(c_int) function mepb_reset(ctx) bind(C, name="mepb_reset")
         import; type(c_ptr), value :: ctx
      end function
      integer(c_int) function mepb_accumulate(ctx, lin_begin, lin_end) bind(C, name="mepb_accumulate")
         import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: lin_begin, lin_end
      end function
      integer(c_int64_t) function mepb_partials_size(ctx) bind(C, name="mepb_partials_size")
         import; type(c_ptr), value :: ctx
      end function
      integer(c_int) function mepb_get_partials(ctx, buf) bind(C, name="mepb_get_partials")
         import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: buf(*)
      end function
      integer(c_int) function mepb_set_partials(ctx, buf) bind(C, name="mepb_set_partials")
         import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: buf(*)
      end function
      integer(c_int) function mepb_finalize(ctx, res) bind(C, name="mepb_finalize")
         import; type(c_ptr), value :: ctx; type(mepb_result), intent(inout) :: res
      end function
      integer(c_int) function mepb_interp(ctx, nk, kpt, e_k, V_ka) bind(C, name="mepb_interp")   ! band_calc.f90:67-75
         import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: nk
         real(c_double), intent(in) :: kpt(3,*); real(c_double), intent(out) :: e_k(*); type(c_ptr), value :: V_ka
      end function
      function mepb_last_error() bind(C, name="mepb_last_error") result(msg)
         import; type(c_ptr) :: msg
      end function
   end interface
end module
