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
      type(c_ptr)        :: r_pos              ! r_tb(3,nw,nw,nR) complex(dp), Bohr (file_io.f90:782-790) or c_null_ptr
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
      integer(c_int32_t) :: do_gyro            ! jobs%do_gyro
      integer(c_int32_t) :: posop_mode         ! 0: PRB 74 195118 eqs. 21-31; 1: curvature / velocity corrections as coded
   end type
   type, bind(C) :: mepb_result                ! include/mepb.h: mepb_result
      integer(c_int64_t) :: n_k
      integer(c_int32_t) :: eF_idx             ! 0-based
      integer(c_int64_t) :: n_degenerate
      type(c_ptr) :: Ne
      type(c_ptr) :: mep_ic, mep_lc, mep_cs, mep_bands
      type(c_ptr) :: kubo_mep_ic, kubo_mep_lc, kubo_mep_cs
      type(c_ptr) :: ahc, ahc_velo, ohc_velo, opt_s, opt_a, photo2
      type(c_ptr) :: gyro_c, gyro_d, gyro_dw
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
      ! k ownership by round-robin blocks (mpi_ki_selector, mpi_comm.f90:80-99, at block granularity) and the in-library NCCL
      ! all-reduce that replaces the mpi_reduce_sum calls of core.f90:467-526
      integer(c_int) function mepb_accumulate_strided(ctx, lin_begin, lin_end, blk, rank, nranks) bind(C, name="mepb_accumulate_strided")
         import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: lin_begin, lin_end, blk; integer(c_int), value :: rank, nranks
      end function
      integer(c_int64_t) function mepb_balanced_block(ctx, nk, nranks) bind(C, name="mepb_balanced_block")
         import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: nk; integer(c_int), value :: nranks
      end function
      integer(c_int) function mepb_comm_unique_id(id128, nbytes) bind(C, name="mepb_comm_unique_id")
         import; character(kind=c_char), intent(out) :: id128(128); integer(c_int), value :: nbytes
      end function
      integer(c_int) function mepb_comm_init_rank(ctx, nranks, rank, id128) bind(C, name="mepb_comm_init_rank")
         import; type(c_ptr), value :: ctx; integer(c_int), value :: nranks, rank; character(kind=c_char), intent(in) :: id128(128)
      end function
      integer(c_int) function mepb_allreduce_partials(ctx) bind(C, name="mepb_allreduce_partials")
         import; type(c_ptr), value :: ctx
      end function
      integer(c_int) function mepb_update_model(ctx, model) bind(C, name="mepb_update_model")
         import; type(c_ptr), value :: ctx; type(mepb_model), intent(in) :: model
      end function
      integer(c_int) function mepb_update_params(ctx, params) bind(C, name="mepb_update_params")
         import; type(c_ptr), value :: ctx; type(mepb_params), intent(in) :: params
      end function
      function mepb_last_error() bind(C, name="mepb_last_error") result(msg)
         import; type(c_ptr) :: msg
      end function
   end interface
end module
