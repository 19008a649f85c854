!!  TransformIntegralsG.f90 -- ISO_C_BINDING shim that selects the B200 transformer (liblowdin_itgpu.so,
!!  include/lowdin_it.h) from openLOWDIN as integralsTransformationMethod = "G".
!!
!!  Drop this file into src/integralsTransformation/ of the reference, add the `case ("G")` lines of INTEGRATION.md
!!  section 2 and link with -llowdin_itgpu.  NOT COMPILED IN THIS REPOSITORY'S BUILD ENVIRONMENT (no Fortran compiler is
!!  installed there): the same call sequence is exercised from C++ (openlowdin_b200/csrc/host_mirror.cpp, function
!!  run_and_write) and from Python (tests/test_host_mirror.py), and the record layout written below is checked byte for
!!  byte against gfortran's sequential unformatted format in tests/test_host_mirror.py::test_moint_record_layouts.
!!
!!  It keeps the reference's own conventions, so nothing else in the program changes:
!!    * same subroutine signatures as TransformIntegralsC_atomicToMolecularOfOneSpecie / OfTwoSpecies
!!      (TransformIntegralsC.f90:141, :728), minus the unused density / auxiliary matrices;
!!    * window table and `symmetric` flag from TransformIntegralsC_checkMOIntegralType / checkInterMOIntegralType
!!      (TransformIntegralsC.f90:1436-1963): the transformer object of method C is reused for that.  Both routines are
!!      private in TransformIntegralsC_ today: add their two names to its `public ::` list (TransformIntegralsC.f90:81-86);
!!    * AO integrals are read from the per-thread stream files of TransformIntegralsC.f90:231-298 / :852-972 and handed to
!!      the library as raw bytes, 64 MiB at a time (blocks of int32 p,q,r,s; real64 v; terminator p = -1 decoded on the device);
!!    * <prefix>moint.dat is written in method C's layout (TransformIntegralsC.f90:419-456), so
!!      ReadTransformedIntegrals needs only `case ("C","G")`.
module TransformIntegralsG_
  use, intrinsic :: iso_c_binding
  use MolecularSystem_
  use InputCI_
  use Matrix_
  use Exception_
  use CONTROL_
  use String_
  use TransformIntegralsC_
  use omp_lib
  implicit none

  type, public :: TransformIntegralsG
     type(c_ptr) :: handle = c_null_ptr
     type(TransformIntegralsC) :: tables          !! window tables + partialTransform of method C
  end type TransformIntegralsG

  integer(c_int), parameter :: LOWDIN_IT_CONV_C = 0_c_int

  interface   !! include/lowdin_it.h -- every function returns int, 0 = ok
     integer(c_int) function lowdin_it_create(device, handle) bind(C, name="lowdin_it_create")
       import :: c_int, c_ptr
       integer(c_int), value :: device
       type(c_ptr) :: handle
     end function lowdin_it_create
     integer(c_int) function lowdin_it_destroy(handle) bind(C, name="lowdin_it_destroy")
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function lowdin_it_destroy
     type(c_ptr) function lowdin_it_last_error(handle) bind(C, name="lowdin_it_last_error")
       import :: c_ptr
       type(c_ptr), value :: handle
     end function lowdin_it_last_error
     integer(c_int) function lowdin_it_set_species(handle, slot, nao, coeff, ldc, ncols) bind(C, name="lowdin_it_set_species")
       import :: c_int, c_ptr
       type(c_ptr), value :: handle, coeff
       integer(c_int), value :: slot, nao, ldc, ncols
     end function lowdin_it_set_species
     integer(c_int) function lowdin_it_ao_begin(handle, slotA, slotB, swapped) bind(C, name="lowdin_it_ao_begin")
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: slotA, slotB, swapped
     end function lowdin_it_ao_begin
     integer(c_int) function lowdin_it_ao_push_stacks(handle, p, q, r, s, v, n) bind(C, name="lowdin_it_ao_push_stacks")
       import :: c_int, c_int64_t, c_ptr
       type(c_ptr), value :: handle, p, q, r, s, v
       integer(c_int64_t), value :: n
     end function lowdin_it_ao_push_stacks
     integer(c_int) function lowdin_it_ao_push_blocks(handle, blocks, nblocks, stackSize) bind(C, name="lowdin_it_ao_push_blocks")
       import :: c_int, c_int64_t, c_ptr
       type(c_ptr), value :: handle, blocks
       integer(c_int64_t), value :: nblocks
       integer(c_int), value :: stackSize
     end function lowdin_it_ao_push_blocks
     integer(c_int) function lowdin_it_ao_end(handle) bind(C, name="lowdin_it_ao_end")
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function lowdin_it_ao_end
     integer(c_int) function lowdin_it_transform(handle, slotA, slotB, win, conv, symmetric, tol) bind(C, name="lowdin_it_transform")
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: handle, win
       integer(c_int), value :: slotA, slotB, conv, symmetric
       real(c_double), value :: tol
     end function lowdin_it_transform
     integer(c_int) function lowdin_it_result_count(handle, count) bind(C, name="lowdin_it_result_count")
       import :: c_int, c_int64_t, c_ptr
       type(c_ptr), value :: handle
       integer(c_int64_t) :: count
     end function lowdin_it_result_count
     integer(c_int) function lowdin_it_download_quads(handle, p, q, r, s, v) bind(C, name="lowdin_it_download_quads")
       import :: c_int, c_ptr
       type(c_ptr), value :: handle, p, q, r, s, v
     end function lowdin_it_download_quads
     integer(c_size_t) function c_strlen(str) bind(C, name="strlen")
       import :: c_size_t, c_ptr
       type(c_ptr), value :: str
     end function c_strlen
  end interface

  public :: TransformIntegralsG_constructor, TransformIntegralsG_destructor, &
       TransformIntegralsG_atomicToMolecularOfOneSpecie, TransformIntegralsG_atomicToMolecularOfTwoSpecies

contains

  subroutine TransformIntegralsG_constructor(this, partial)
    implicit none
    type(TransformIntegralsG) :: this
    character(*) :: partial

    call TransformIntegralsC_constructor(this%tables, partial)
    !! no usable GPU -> ERROR exception; the library has no CPU path
    if (lowdin_it_create(0_c_int, this%handle) /= 0) call TransformIntegralsG_fail(c_null_ptr)
  end subroutine TransformIntegralsG_constructor

  subroutine TransformIntegralsG_destructor(this)
    implicit none
    type(TransformIntegralsG) :: this
    integer(c_int) :: rc
    rc = lowdin_it_destroy(this%handle)
    this%handle = c_null_ptr
  end subroutine TransformIntegralsG_destructor

  !> One species: replaces TransformIntegralsC_atomicToMolecularOfOneSpecie (TransformIntegralsC.f90:141-471)
  subroutine TransformIntegralsG_atomicToMolecularOfOneSpecie(this, coefficientsOfAtomicOrbitals, speciesID, nameOfSpecie)
    implicit none
    type(TransformIntegralsG) :: this
    type(Matrix), target :: coefficientsOfAtomicOrbitals
    integer :: speciesID
    character(*) :: nameOfSpecie
    logical :: symmetric
    integer(c_int), target :: win(8)
    integer(c_int) :: nao, ncols

    this%tables%prefixOfFile = ""//trim(nameOfSpecie)
    this%tables%specieID = speciesID
    nao = size(coefficientsOfAtomicOrbitals%values, dim=1)
    ncols = size(coefficientsOfAtomicOrbitals%values, dim=2)
    this%tables%numberOfContractions = nao
    call TransformIntegralsC_checkMOIntegralType(speciesID, this%tables, symmetric)
    win = (/ this%tables%p_l, this%tables%p_u, this%tables%q_l, this%tables%q_u, &
             this%tables%r_l, this%tables%r_u, this%tables%s_l, this%tables%s_u /)

    if (lowdin_it_set_species(this%handle, 0_c_int, nao, c_loc(coefficientsOfAtomicOrbitals%values(1,1)), nao, ncols) /= 0) &
         call TransformIntegralsG_fail(this%handle)

    if (lowdin_it_ao_begin(this%handle, 0_c_int, 0_c_int, 0_c_int) /= 0) call TransformIntegralsG_fail(this%handle)
    if (trim(nameOfSpecie) == "E-BETA") then          !! TransformIntegralsC.f90:241-245
       call TransformIntegralsG_pushStreams(this, "E-ALPHA")
    else
       call TransformIntegralsG_pushStreams(this, trim(nameOfSpecie))
    end if
    if (lowdin_it_ao_end(this%handle) /= 0) call TransformIntegralsG_fail(this%handle)

    if (lowdin_it_transform(this%handle, 0_c_int, 0_c_int, c_loc(win), LOWDIN_IT_CONV_C, merge(1_c_int, 0_c_int, symmetric), &
         1.0E-10_c_double) /= 0) call TransformIntegralsG_fail(this%handle)

    call TransformIntegralsG_writeMOIntegrals(this, trim(this%tables%prefixOfFile)//"moint.dat")
  end subroutine TransformIntegralsG_atomicToMolecularOfOneSpecie

  !> Two species: replaces TransformIntegralsC_atomicToMolecularOfTwoSpecies (TransformIntegralsC.f90:728-1168).
  !! The program calls it with the species of fewer occupied orbitals first (IntegralTransformation.f90:322-334); when that
  !! order is the reverse of the molecular-system order the stream on disk holds (B B|A A) and is loaded with swapped = 1
  !! (TransformIntegralsC.f90:906-972).
  subroutine TransformIntegralsG_atomicToMolecularOfTwoSpecies(this, coefficientsOfAtomicOrbitals, otherCoefficientsOfAtomicOrbitals, &
       speciesID, nameOfSpecie, otherSpeciesID, nameOfOtherSpecie)
    implicit none
    type(TransformIntegralsG) :: this
    type(Matrix), target :: coefficientsOfAtomicOrbitals, otherCoefficientsOfAtomicOrbitals
    integer :: speciesID, otherSpeciesID
    character(*) :: nameOfSpecie, nameOfOtherSpecie
    logical :: symmetric
    integer(c_int), target :: win(8)
    integer(c_int) :: nao, ncols, onao, oncols, swapped
    character(100) :: first, second

    this%tables%prefixOfFile = ""//trim(nameOfSpecie)//"."//trim(nameOfOtherSpecie)
    nao = size(coefficientsOfAtomicOrbitals%values, dim=1)
    ncols = size(coefficientsOfAtomicOrbitals%values, dim=2)
    onao = size(otherCoefficientsOfAtomicOrbitals%values, dim=1)
    oncols = size(otherCoefficientsOfAtomicOrbitals%values, dim=2)
    this%tables%numberOfContractions = nao
    this%tables%otherNumberOfContractions = onao
    call TransformIntegralsC_checkInterMOIntegralType(speciesID, otherSpeciesID, this%tables, symmetric)
    win = (/ this%tables%p_l, this%tables%p_u, this%tables%q_l, this%tables%q_u, &
             this%tables%r_l, this%tables%r_u, this%tables%s_l, this%tables%s_u /)

    if (lowdin_it_set_species(this%handle, 0_c_int, nao, c_loc(coefficientsOfAtomicOrbitals%values(1,1)), nao, ncols) /= 0) &
         call TransformIntegralsG_fail(this%handle)
    if (lowdin_it_set_species(this%handle, 1_c_int, onao, c_loc(otherCoefficientsOfAtomicOrbitals%values(1,1)), onao, oncols) /= 0) &
         call TransformIntegralsG_fail(this%handle)

    !! the stream was written for the pair in molecular-system order; E-BETA reads E-ALPHA's (TransformIntegralsC.f90:838-848, :906-915)
    if (speciesID < otherSpeciesID) then
       first = trim(nameOfSpecie); second = trim(nameOfOtherSpecie); swapped = 0_c_int
    else
       first = trim(nameOfOtherSpecie); second = trim(nameOfSpecie); swapped = 1_c_int
    end if
    if (trim(first) == "E-ALPHA" .and. trim(second) == "E-BETA") then
       continue
    else if (trim(second) == "E-BETA") then
       second = "E-ALPHA"
    else if (trim(first) == "E-BETA") then
       first = "E-ALPHA"
    end if

    if (lowdin_it_ao_begin(this%handle, 0_c_int, 1_c_int, swapped) /= 0) call TransformIntegralsG_fail(this%handle)
    call TransformIntegralsG_pushStreams(this, trim(first)//"."//trim(second))
    if (lowdin_it_ao_end(this%handle) /= 0) call TransformIntegralsG_fail(this%handle)

    if (lowdin_it_transform(this%handle, 0_c_int, 1_c_int, c_loc(win), LOWDIN_IT_CONV_C, merge(1_c_int, 0_c_int, symmetric), &
         1.0E-10_c_double) /= 0) call TransformIntegralsG_fail(this%handle)

    call TransformIntegralsG_writeMOIntegrals(this, trim(this%tables%prefixOfFile)//"moint.dat")
  end subroutine TransformIntegralsG_atomicToMolecularOfTwoSpecies

  !> Reads <tid><stem>.ints of every producer thread in pieces of up to 64 MiB and passes the RAW BYTES to the library
  !! (lowdin_it_ao_push_blocks): the blocks `pp, qq, rr, ss, shellIntegrals` that the reader loops of
  !! TransformIntegralsC.f90:247-296 consume one by one are decoded on the device -- terminator (p = -1, :279-280), index range
  !! check and the scatter to ioff(min)+max (:264-272) -- so the host only moves bytes (55 GB/s from pinned memory on B200).
  subroutine TransformIntegralsG_pushStreams(this, stem)
    implicit none
    type(TransformIntegralsG) :: this
    character(*) :: stem
    integer(c_int8_t), allocatable, target :: raw(:)
    character(50) :: fileid
    integer :: nfiles, tid, unitid, status, stackSize
    integer(8) :: filesize, nstacks, istack, perRead, n
    logical :: existFile

    stackSize = CONTROL_instance%INTEGRAL_STACK_SIZE
    perRead = max(1_8, (64_8 * 1024_8 * 1024_8) / (24_8 * stackSize))     !! stacks per read
    allocate(raw(perRead * 24_8 * stackSize))
    nfiles = omp_get_max_threads()          !! lowdin-ints.x wrote one stream per OpenMP thread (Libint2Iface.cpp:282-286)
    unitid = 40
    do tid = 0, nfiles - 1
       write(fileid,*) tid
       fileid = trim(adjustl(fileid))
       inquire(file=trim(fileid)//trim(stem)//".ints", exist=existFile)
       if (.not. existFile) cycle
       open(unit=unitid, file=trim(fileid)//trim(stem)//".ints", status='old', access='stream', form='unformatted')
       inquire(unit=unitid, size=filesize)
       nstacks = filesize/24/stackSize        !! 24 bytes per integral (TransformIntegralsC.f90:251-253)
       istack = 0
       do while (istack < nstacks)
          n = min(perRead, nstacks - istack)
          read(unit=unitid, iostat=status) raw(1:n * 24_8 * stackSize)
          if (status /= 0) exit
          !! one call per file piece; the terminator of the file's last stack ends the call's stream on the device
          if (lowdin_it_ao_push_blocks(this%handle, c_loc(raw), int(n, c_int64_t), int(stackSize, c_int)) /= 0) &
               call TransformIntegralsG_fail(this%handle)
          istack = istack + n
       end do
       close(unitid)
    end do
    deallocate(raw)
  end subroutine TransformIntegralsG_pushStreams

  !> Downloads the kept integrals (|x| > 1E-10, p,q,r,s loop order) and writes them in method C's record layout:
  !! stacks pp,qq,rr,ss,auxIntegrals of INTEGRAL_STACK_SIZE, the last one carrying pp(m+1) = -1 (TransformIntegralsC.f90:419-456)
  subroutine TransformIntegralsG_writeMOIntegrals(this, fileName)
    implicit none
    type(TransformIntegralsG) :: this
    character(*) :: fileName
    integer(c_int), allocatable, target :: p(:), q(:), r(:), s(:)
    real(c_double), allocatable, target :: v(:)
    integer :: pp(CONTROL_instance%INTEGRAL_STACK_SIZE), qq(CONTROL_instance%INTEGRAL_STACK_SIZE)
    integer :: rr(CONTROL_instance%INTEGRAL_STACK_SIZE), ss(CONTROL_instance%INTEGRAL_STACK_SIZE)
    real(8) :: auxIntegrals(CONTROL_instance%INTEGRAL_STACK_SIZE)
    integer(c_int64_t) :: n, k
    integer :: m, stackSize

    stackSize = CONTROL_instance%INTEGRAL_STACK_SIZE
    if (lowdin_it_result_count(this%handle, n) /= 0) call TransformIntegralsG_fail(this%handle)
    allocate(p(max(n,1_8)), q(max(n,1_8)), r(max(n,1_8)), s(max(n,1_8)), v(max(n,1_8)))
    if (lowdin_it_download_quads(this%handle, c_loc(p), c_loc(q), c_loc(r), c_loc(s), c_loc(v)) /= 0) &
         call TransformIntegralsG_fail(this%handle)

    open(unit=CONTROL_instance%UNIT_FOR_MP2_INTEGRALS_FILE, file=trim(fileName), &
         status='replace', access='sequential', form='unformatted')
    pp = 0; qq = 0; rr = 0; ss = 0; auxIntegrals = 0.0_8
    m = 0
    do k = 1, n
       m = m + 1
       pp(m) = p(k); qq(m) = q(k); rr(m) = r(k); ss(m) = s(k); auxIntegrals(m) = v(k)
       if (m == stackSize) then
          write(CONTROL_instance%UNIT_FOR_MP2_INTEGRALS_FILE) pp, qq, rr, ss, auxIntegrals
          m = 0
          pp = 0; qq = 0; rr = 0; ss = 0; auxIntegrals = 0.0_8
       end if
    end do
    pp(m+1) = -1
    write(CONTROL_instance%UNIT_FOR_MP2_INTEGRALS_FILE) pp, qq, rr, ss, auxIntegrals
    close(CONTROL_instance%UNIT_FOR_MP2_INTEGRALS_FILE)

    write(*,"(T4,A36,I12)") "Non-zero transformed integrals: ", n
    deallocate(p, q, r, s, v)
  end subroutine TransformIntegralsG_writeMOIntegrals

  !> lowdin_it_last_error -> the reference's ERROR exception (prints and stops, as TransformIntegralsC.f90:2030-2044)
  subroutine TransformIntegralsG_fail(handle)
    implicit none
    type(c_ptr) :: handle
    type(Exception) :: ex
    type(c_ptr) :: cmsg
    character(kind=c_char), pointer :: chars(:)
    character(512) :: message
    integer :: i, n

    message = "liblowdin_itgpu failed"
    cmsg = lowdin_it_last_error(handle)
    if (c_associated(cmsg)) then
       n = min(int(c_strlen(cmsg)), 512)
       call c_f_pointer(cmsg, chars, (/ n /))
       message = ""
       do i = 1, n
          message(i:i) = chars(i)
       end do
    end if
    call Exception_constructor(ex, ERROR)
    call Exception_setDebugDescription(ex, "Class object TransformIntegralsG")
    call Exception_setDescription(ex, trim(message))
    call Exception_show(ex)
  end subroutine TransformIntegralsG_fail

end module TransformIntegralsG_
