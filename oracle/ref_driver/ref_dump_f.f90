! ref_dump_f.f90 - dump driver for the REAL reference, fixed-h program (test infrastructure; SURVEY.md 8(c)(ii)).
! Linked against SPH_routines_module exactly as /root/reference/SUMMER_SPH.f90 defines it (scripts/build_ref_oracle.sh
! extracts the module text at build time into oracle/_ref/build/; no reference source is copied into this repository).
! Reads ics.txt with the reference's own reader, runs ONE evaluation in the order of SUMMER_SPH.f90:894-898 and writes
! the per-particle quantities make_save never writes to dump.bin (stream, native endian):
!   int32 n, int32 ns, then n x [rho P c ax ay az udot alphadot] as 8 arrays of real(8), then ns x [ax ay az].
program ref_dump
  use SPH_routines_module
  implicit none
  type(particle), allocatable :: bodies(:)
  type(sink), allocatable :: sinks(:)
  type(branch), allocatable :: root
  character(len=256) :: filename
  integer :: i, k
  call init_kernel_table()
  call init_grav_kernel_table()
  filename = 'ics.txt'
  call read_data_from_file(filename, bodies, sinks)
  do i = 1, size(bodies)
    bodies(i)%number = i                                   ! SUMMER_SPH.f90:886-888
  end do
  allocate(root)
  call create_tree(root, bodies, max_depth)                 ! :894-895
  call get_density(root, bodies)                            ! :896
  call get_pressure_and_sound_speed(bodies)                 ! :897
  call find_forces(root, bodies, sinks)                     ! :898
  open(unit=20, file='dump.bin', access='stream', form='unformatted', status='replace')
  write(20) int(size(bodies), 4), int(size(sinks), 4)
  write(20) (bodies(i)%density, i = 1, size(bodies))
  write(20) (bodies(i)%pressure, i = 1, size(bodies))
  write(20) (bodies(i)%sound_speed, i = 1, size(bodies))
  do k = 1, 3
    write(20) (bodies(i)%acceleration(k), i = 1, size(bodies))
  end do
  write(20) (bodies(i)%internal_energy_rate, i = 1, size(bodies))
  write(20) (bodies(i)%alpha_rate, i = 1, size(bodies))
  do k = 1, 3
    write(20) (sinks(i)%acceleration(k), i = 1, size(sinks))
  end do
  close(20)
end program ref_dump
