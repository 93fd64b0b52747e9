! shim/mo_gas_optics_utils_shim.F90 - the two NON-C symbols of the extern-mode kernel interface.
!
! rte/kernels/api/mo_gas_optics_utils.F90:38-66 declares `get_layer_mass` (subroutine) and `get_layer_number`
! (ARRAY-VALUED function) as plain external Fortran procedures: their link names are compiler-mangled
! (`get_layer_mass_`, `get_layer_number_`) and the function result travels by a compiler-specific hidden
! argument, so a C library cannot export them portably.  Compile this file with the HOST's Fortran compiler and
! link it between librte/librrtmgp (built with -DRTE_KERNEL_MODE=extern) and librte_rrtmgp_b200.so:
!
!     $FC -c shim/mo_gas_optics_utils_shim.F90 -I<rte-rrtmgp module dir>
!     $FC host.o mo_gas_optics_utils_shim.o -lrrtmgp -lrte -L<repo>/rte_rrtmgp_b200/lib -lrte_rrtmgp_b200
!
! They forward to the library's C-callable kernels (include/rrtmgp_b200_ext.h: rrtmgpb_get_col_dry,
! rrtmgpb_get_layer_mass), which restate rte/kernels/mo_gas_optics_utils.F90:97-152.
! (Cannot be compiled in this repository's image: it has no Fortran compiler.  The C side is tested in
! tests/test_kernels_parity.py::test_glue_kernels and tests/test_ssm_kernels.py.)
function get_layer_number(ncol, nlay, vmr_h2o, plev) result(col_dry)
  use iso_c_binding, only: c_int, c_double, c_float
  use mo_rte_kind,   only: wp
  implicit none
  integer,                           intent(in) :: ncol, nlay
  real(wp), dimension(ncol, nlay  ), intent(in) :: vmr_h2o
  real(wp), dimension(ncol, nlay+1), intent(in) :: plev
  real(wp), dimension(ncol, nlay)               :: col_dry
  interface
    subroutine rrtmgpb_get_col_dry(ncol, nlay, vmr_h2o, plev, col_dry) bind(C, name="rrtmgpb_get_col_dry")
      import :: c_int, wp
      integer(c_int), value :: ncol, nlay
      real(wp), intent(in)  :: vmr_h2o(*), plev(*)
      real(wp), intent(out) :: col_dry(*)
    end subroutine rrtmgpb_get_col_dry
  end interface
  call rrtmgpb_get_col_dry(int(ncol, c_int), int(nlay, c_int), vmr_h2o, plev, col_dry)
end function get_layer_number

subroutine get_layer_mass(ncol, nlay, ngas, vmr, plev, mol_weights, m_dry, layer_mass)
  use iso_c_binding, only: c_int
  use mo_rte_kind,   only: wp
  implicit none
  integer,                                 intent(in ) :: ncol, nlay, ngas
  real(wp), dimension(ngas, ncol, nlay  ), intent(in ) :: vmr
  real(wp), dimension(      ncol, nlay+1), intent(in ) :: plev
  real(wp), dimension(ngas),               intent(in ) :: mol_weights
  real(wp),                                intent(in ) :: m_dry
  real(wp), dimension(ngas, ncol, nlay),   intent(out) :: layer_mass
  interface
    subroutine rrtmgpb_get_layer_mass(ncol, nlay, ngas, vmr, plev, mol_weights, m_dry, layer_mass) &
        bind(C, name="rrtmgpb_get_layer_mass")
      import :: c_int, wp
      integer(c_int), value :: ncol, nlay, ngas
      real(wp), intent(in)  :: vmr(*), plev(*), mol_weights(*)
      real(wp), value       :: m_dry
      real(wp), intent(out) :: layer_mass(*)
    end subroutine rrtmgpb_get_layer_mass
  end interface
  call rrtmgpb_get_layer_mass(int(ncol, c_int), int(nlay, c_int), int(ngas, c_int), vmr, plev, mol_weights, m_dry, layer_mass)
end subroutine get_layer_mass
