! run_sph_b200.f90 -- Fortran host for the B200 SPH step engine (ISO_C_BINDING over include/sph_b200.h).
!
! Keeps the reference's program shell: `program run_sph`, read_data_from_file, read_params_from_file,
! make_save, the `do while (t < end_time)` loop, the per-step print and the save cadence
! (SUMMER_SPH.f90:594-738, 863-955 | "SUMMER_SPH - Variable.f90":729-942, 1076-1191).  The loop body
! (F:886-928 | V:1120-1162) is ONE call: sph_step.
!
! NOT COMPILED IN THE BUILD IMAGE (no gfortran/flang/nvfortran there, SURVEY.md 8(c)); the C++ twin
! host/run_sph.cpp exercises the same entry points.  Build where a compiler exists:
!     make -C host run_sph_b200
module sph_b200_c
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int32_t), parameter :: SPH_MODE_FIXED_H = 0, SPH_MODE_VARIABLE_H = 1
  integer(c_int32_t), parameter :: SPH_FLAG_SOFT_USES_HI = 2, SPH_FLAG_SINK_MERGE_SPIN = 4     ! OR-ed into sph_params%mode

  type, bind(C) :: sph_params          ! mirrors `struct sph_params` (and V's type(param), V:54-64)
    integer(c_int32_t) :: mode, max_depth, nq, n_ranks
    real(c_double)     :: h_fixed, bounding_size, theta, gamma, eta, convergence_criteria
    real(c_double)     :: max_length, timestep_scale, end_time, sink_radius
    integer(c_int32_t) :: theta_override, decomposition
  end type sph_params

  interface
    integer(c_int) function sph_default_params(mode, p) bind(C, name="sph_default_params")
      import :: c_int, c_int32_t, sph_params
      integer(c_int32_t), value :: mode
      type(sph_params), intent(out) :: p
    end function
    integer(c_int) function sph_create(p, device, ctx) bind(C, name="sph_create")
      import :: c_int, c_int32_t, c_ptr, sph_params
      type(sph_params), intent(in) :: p
      integer(c_int32_t), value :: device
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function sph_destroy(ctx) bind(C, name="sph_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    function sph_last_error(ctx) result(msg) bind(C, name="sph_last_error")
      import :: c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr) :: msg
    end function
    integer(c_int) function sph_upload(ctx, n_gas, x, y, z, vx, vy, vz, u, m, alpha, h, n_sink, &
                                       sx, sy, sz, svx, svy, svz, sm, srad) bind(C, name="sph_upload")
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int64_t), value :: n_gas
      real(c_double), intent(in) :: x(*), y(*), z(*), vx(*), vy(*), vz(*), u(*), m(*), alpha(*), h(*)
      integer(c_int32_t), value :: n_sink
      real(c_double), intent(in) :: sx(*), sy(*), sz(*), svx(*), svy(*), svz(*), sm(*), srad(*)
    end function
    integer(c_int) function sph_step(ctx, dt, t, n_gas, n_sink) bind(C, name="sph_step")
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: dt, t
      integer(c_int64_t), intent(out) :: n_gas
      integer(c_int32_t), intent(out) :: n_sink
    end function
    ! upload + one loop body + download in one call (include/sph_b200.h: sph_step_host); gas_in / gas_out = 10 column
    ! pointers (x y z vx vy vz u m alpha h), sink_in / sink_out = 8 (x y z vx vy vz m radius)
    integer(c_int) function sph_step_host(ctx, n_gas, gas_in, n_sink, sink_in, dt, t, gas_out, sink_out, n_gas_out, n_sink_out) &
        bind(C, name="sph_step_host")
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int64_t), value :: n_gas
      type(c_ptr), intent(in) :: gas_in(10), sink_in(8), gas_out(10), sink_out(8)
      integer(c_int32_t), value :: n_sink
      real(c_double), intent(inout) :: dt, t
      integer(c_int64_t), intent(out) :: n_gas_out
      integer(c_int32_t), intent(out) :: n_sink_out
    end function
    integer(c_int) function sph_sizes(ctx, n_gas, n_sink) bind(C, name="sph_sizes")
      import :: c_int, c_int32_t, c_int64_t, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int64_t), intent(out) :: n_gas
      integer(c_int32_t), intent(out) :: n_sink
    end function
    integer(c_int) function sph_download(ctx, x, y, z, vx, vy, vz, u, m, alpha, h, &
                                         sx, sy, sz, svx, svy, svz, sm, srad) bind(C, name="sph_download")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: x(*), y(*), z(*), vx(*), vy(*), vz(*), u(*), m(*), alpha(*), h(*)
      real(c_double), intent(out) :: sx(*), sy(*), sz(*), svx(*), svy(*), svz(*), sm(*), srad(*)
    end function
    integer(c_int) function sph_conserved(ctx, out, n_out) bind(C, name="sph_conserved")
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: out(*)          ! E_kin E_int E_pot P(3) L(3) M E_pot_gas E_pot_sink
      integer(c_int32_t), value :: n_out
    end function
    ! column-density image of the resident gas (the Density_Image.py counterpart); image holds nu*nv doubles, row iv, column iu
    integer(c_int) function sph_column_density(ctx, axis, u0, u1, v0, v1, nu, nv, image) bind(C, name="sph_column_density")
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: axis, nu, nv
      real(c_double), value :: u0, u1, v0, v1
      real(c_double), intent(out) :: image(*)
    end function
    ! sink spins (all zero unless the context was created with mode + SPH_FLAG_SINK_MERGE_SPIN)
    integer(c_int) function sph_download_sink_spin(ctx, spin_x, spin_y, spin_z) bind(C, name="sph_download_sink_spin")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: spin_x(*), spin_y(*), spin_z(*)
    end function
  end interface
end module sph_b200_c

program run_sph
  use, intrinsic :: iso_c_binding
  use sph_b200_c
  implicit none
  integer, parameter :: dp = kind(1.0d0)
  character(len=256) :: filename, header_line
  type(sph_params) :: p
  type(c_ptr) :: ctx
  real(dp), allocatable :: x(:), y(:), z(:), vx(:), vy(:), vz(:), u(:), m(:), alpha(:), h(:)
  real(dp), allocatable :: sx(:), sy(:), sz(:), svx(:), svy(:), svz(:), sm(:), srad(:)
  real(dp), allocatable :: row(:,:)
  real(dp) :: t, dt, v(10), cons0(12), cons1(12), e0, e1
  integer :: status, num_lines, i, nb, ns, t_test, rc, ncol
  integer(c_int64_t) :: n_gas
  integer(c_int32_t) :: n_sink
  logical :: variable

  variable = (command_argument_count() >= 1)          ! any argument selects the variable-h reference
  filename = merge('disc_20k_low_vel.txt', 'disc_12000_2.txt    ', variable)     ! V:1181 | F:946
  rc = sph_default_params(merge(SPH_MODE_VARIABLE_H, SPH_MODE_FIXED_H, variable), p)
  ncol = merge(10, 8, variable)
  if (variable) call read_params_from_file('parameters.txt', p)                  ! V:1182-1184 (the fixed-h program has no parameter file)

  ! ---- read_data_from_file (F:594-716): header skipped, first pass counts records, second pass reads
  open(unit=10, file=trim(filename), status='old', action='read', iostat=status)
  if (status /= 0) then
    write(*,*) 'Error opening file: ', trim(filename); stop 1
  end if
  read(10, '(A)', iostat=status) header_line
  num_lines = 0
  do
    read(10, *, iostat=status)
    if (status /= 0) exit
    num_lines = num_lines + 1
  end do
  close(10)
  allocate(row(10, num_lines)); row = 0.0_dp
  open(unit=10, file=trim(filename), status='old', action='read')
  read(10, '(A)') header_line
  do i = 1, num_lines
    v = 0.0_dp
    read(10, *, iostat=status) v(1:ncol)
    row(:, i) = v
    if (status /= 0) exit                              ! trailing 8-column sink rows in V (SURVEY.md 5)
  end do
  close(10)
  nb = count(row(7, :) /= 0.0_dp); ns = count(row(7, :) == 0.0_dp)
  allocate(x(nb), y(nb), z(nb), vx(nb), vy(nb), vz(nb), u(nb), m(nb), alpha(nb), h(nb))
  allocate(sx(max(ns,1)), sy(max(ns,1)), sz(max(ns,1)), svx(max(ns,1)), svy(max(ns,1)), svz(max(ns,1)), sm(max(ns,1)), srad(max(ns,1)))
  nb = 0; ns = 0
  do i = 1, num_lines
    if (row(7, i) /= 0.0_dp) then
      nb = nb + 1
      x(nb) = row(1, i); y(nb) = row(2, i); z(nb) = row(3, i)
      vx(nb) = row(4, i); vy(nb) = row(5, i); vz(nb) = row(6, i)
      u(nb) = row(7, i); m(nb) = row(8, i)
      alpha(nb) = merge(row(9, i), 0.0_dp, variable)   ! F:681
      h(nb) = merge(row(10, i), p%h_fixed, variable)
    else
      ns = ns + 1
      sx(ns) = row(1, i); sy(ns) = row(2, i); sz(ns) = row(3, i)
      svx(ns) = row(4, i); svy(ns) = row(5, i); svz(ns) = row(6, i)
      sm(ns) = row(8, i); srad(ns) = p%sink_radius      ! F:694 | V:830
    end if
  end do
  write(*,*) 'Successfully read ', nb, ' bodies and ', max(ns, 1), ' sinks from ', trim(filename), '.'

  ! ---- simulate (F:863-930): the loop body is one call into the engine
  rc = sph_create(p, 0_c_int32_t, ctx)
  if (rc /= 0) stop 'sph_create failed (no CUDA device? there is no CPU fallback)'
  rc = sph_upload(ctx, int(nb, c_int64_t), x, y, z, vx, vy, vz, u, m, alpha, h, int(ns, c_int32_t), &
                  sx, sy, sz, svx, svy, svz, sm, srad)
  if (rc /= 0) stop 'sph_upload failed'
  rc = sph_conserved(ctx, cons0, 12_c_int32_t)           ! not in the reference: conserved sums for the drift report
  t = 0.0_dp; dt = 1.0e-2_dp; t_test = 0                 ! F:871-875
  n_gas = nb
  do while (t < p%end_time)                              ! F:879
    if (t > t_test * p%end_time / 1000) then             ! F:881 (t_list(0) := 0)
      call make_save(t_test)
      t_test = t_test + 1
    end if
    print *, "SPH Particles:", n_gas, "dt :", dt, "time : ", t      ! F:891
    rc = sph_step(ctx, dt, t, n_gas, n_sink)
    if (rc /= 0) stop 'sph_step failed'
  end do
  rc = sph_conserved(ctx, cons1, 12_c_int32_t)
  e0 = sum(cons0(1:3)); e1 = sum(cons1(1:3))
  print *, "Energy:", e0, "->", e1, " dE/|E0| =", (e1 - e0) / abs(e0)
  print *, "|dP| =", norm2(cons1(4:6) - cons0(4:6)), " |dL|/|L0| =", norm2(cons1(7:9) - cons0(7:9)) / norm2(cons0(7:9))
  rc = sph_destroy(ctx)

contains

  ! V:854-919: one header line, then rows of `bounding_size max_depth theta gamma eta convergence_criteria max_length
  ! timestep_scale end_time`; every row is read, the LAST one stands (V:900-918).  A missing or empty file leaves the
  ! defaults and says so, as the reference does (V:880-893).  theta is stored but, like in the reference (V:1029 passes
  ! the literal 0.5), it does not reach the opening test: theta_override stays 0.
  subroutine read_params_from_file(fname, params)
    character(len=*), intent(in) :: fname
    type(sph_params), intent(inout) :: params
    character(len=256) :: hdr
    integer :: st, nl, k, max_depth
    real(dp) :: bounding_size, theta, gamma, eta, convergence_criteria, max_length, timestep_scale, end_time
    nl = 0
    open(unit=11, file=fname, status='old', action='read', iostat=st)
    if (st /= 0) then
      write(*,*) 'Error opening file: ', trim(fname); return
    end if
    read(11, '(A)', iostat=st) hdr
    do
      read(11, *, iostat=st)
      if (st /= 0) exit
      nl = nl + 1
    end do
    close(11)
    if (nl == 0) then
      write(*,*) 'No data found in file: ', trim(fname); return
    end if
    open(unit=11, file=fname, status='old', action='read', iostat=st)
    read(11, '(A)', iostat=st) hdr
    do k = 1, nl
      read(11, *, iostat=st) bounding_size, max_depth, theta, gamma, eta, convergence_criteria, max_length, timestep_scale, end_time
      if (st /= 0) then
        write(*,*) 'Error reading line ', k; exit
      end if
    end do
    close(11)
    params%bounding_size = bounding_size; params%max_depth = int(max_depth, c_int32_t); params%theta = theta
    params%gamma = gamma; params%eta = eta; params%convergence_criteria = convergence_criteria
    params%max_length = max_length; params%timestep_scale = timestep_scale; params%end_time = end_time
    write(*,*) 'Successfully read parameters from', trim(fname), '.'
  end subroutine read_params_from_file

  subroutine make_save(number)                           ! F:719-738 | V:921-942
    integer, intent(in) :: number
    integer :: io, k
    integer(c_int64_t) :: ng
    integer(c_int32_t) :: nsk
    character(len=256) :: savename
    real(dp), allocatable :: a(:,:), s(:,:)
    rc = sph_sizes(ctx, ng, nsk)
    allocate(a(ng, 10), s(max(nsk, 1), 8))
    rc = sph_download(ctx, a(:,1), a(:,2), a(:,3), a(:,4), a(:,5), a(:,6), a(:,7), a(:,8), a(:,9), a(:,10), &
                      s(:,1), s(:,2), s(:,3), s(:,4), s(:,5), s(:,6), s(:,7), s(:,8))
    write(savename, '(A,I0,A)') 'save', number, '.txt'
    open(newunit=io, file=savename, status="new", action="write")
    if (variable) then
      write(io, *) 'x  ','y  ','z  ','vx  ','vy ','vz ','energy ','mass  ','alpha  ', 'smoothing'
    else
      write(io, *) 'x  ','y  ','z  ','vx  ','vy ','vz ','energy ','mass  ','alpha  '
    end if
    do k = 1, int(ng)
      if (variable) then
        write(io, *) a(k, 1:10)
      else
        write(io, *) a(k, 1:9)
      end if
    end do
    do k = 1, nsk
      write(io, *) s(k, 1:6), 0.0_dp, s(k, 7)
    end do
    close(io)
  end subroutine make_save

end program run_sph
