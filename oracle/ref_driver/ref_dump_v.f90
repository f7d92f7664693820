! ref_dump_v.f90 - dump driver for the REAL reference, variable-h program (test infrastructure; SURVEY.md 8(c)(ii)).
! Linked against SPH_routines_module exactly as "/root/reference/SUMMER_SPH - Variable.f90" defines it (module text
! extracted at build time by scripts/build_ref_oracle.sh; nothing of the reference is copied into this repository).
! Reads parameters.txt and ics.txt (10 columns) with the reference's own readers, runs ONE evaluation in the order of
! Variable.f90:1128-1132, then calc_smoothing (:1152) on that tree, and writes dump.bin (stream, native endian):
!   int32 n, int32 ns, then n x [rho omega P c ax ay az udot alphadot h_after_calc_smoothing] as 10 arrays of
!   real(8), then ns x [ax ay az].
program ref_dump
  use SPH_routines_module
  implicit none
  type(particle), allocatable :: bodies(:)
  type(sink), allocatable :: sinks(:)
  type(branch), allocatable :: root
  type(param) :: params
  character(len=256) :: filename, pname
  real(dp), allocatable :: keep(:,:)
  integer :: i, k
  call init_kernel_table()
  call init_grav_kernel_table()
  pname = 'parameters.txt'
  filename = 'ics.txt'
  call read_params_from_file(pname, params)
  call read_data_from_file(filename, bodies, sinks)
  do i = 1, size(bodies)
    bodies(i)%number = i                                   ! Variable.f90:1120-1122
  end do
  allocate(root)
  call create_tree(root, bodies, params%max_depth)
  call get_density(root, bodies)
  call get_pressure_and_sound_speed(bodies, params%gamma)
  call find_forces(root, bodies, sinks)
  allocate(keep(9, size(bodies)))
  do i = 1, size(bodies)
    keep(1, i) = bodies(i)%density; keep(2, i) = bodies(i)%omega; keep(3, i) = bodies(i)%pressure
    keep(4, i) = bodies(i)%sound_speed; keep(5:7, i) = bodies(i)%acceleration
    keep(8, i) = bodies(i)%internal_energy_rate; keep(9, i) = bodies(i)%alpha_rate
  end do
  call calc_smoothing(root, bodies, params%eta, params%convergence_criteria, params%max_length)   ! :1152
  open(unit=20, file='dump.bin', access='stream', form='unformatted', status='replace')
  write(20) int(size(bodies), 4), int(size(sinks), 4)
  do k = 1, 9
    write(20) (keep(k, i), i = 1, size(bodies))
  end do
  write(20) (bodies(i)%s_length, i = 1, size(bodies))
  do k = 1, 3
    write(20) (sinks(i)%acceleration(k), i = 1, size(sinks))
  end do
  close(20)
end program ref_dump
