! prealpha_b200_bindc.f90 -- ISO_C_BINDING view of include/prealpha_b200.h (ABI version 2).
!
! The interface block a prealpha maintainer adds to bind Module_Distribution / Module_Distances / Module_Cluster to
! libprealpha_b200.so (see INTEGRATION.md for the call sites).  Struct layouts mirror the C header field by field; the
! C side of the same ABI is exercised by prealpha_b200/csrc/pa_frontend.cpp and, through ctypes, by prealpha_b200/capi.py
! (tests/test_host_cpu.py checks those layouts against the header as compiled by gcc).  There is no Fortran compiler in
! the build image of this repository, so this file is shipped uncompiled.
MODULE prealpha_b200_bindc
   USE, INTRINSIC :: ISO_C_BINDING
   IMPLICIT NONE
   INTEGER(C_INT), PARAMETER :: PA_ABI_VERSION = 2
   INTEGER(C_INT32_T), PARAMETER :: PA_MODE_CDF = 0, PA_MODE_PDF = 1, PA_MODE_CHARGE_ARM = 2
   INTEGER(C_INT32_T), PARAMETER :: PA_EXEC_FORCE_GENERAL = 1, PA_EXEC_NO_CELLS = 2, PA_EXEC_NO_SYMMETRY = 4, PA_EXEC_NO_F32 = 8
   INTEGER(C_INT), PARAMETER :: PA_ERR_NO_DEVICE = -1, PA_ERR_CUDA = -2, PA_ERR_BAD_ARG = -3, PA_ERR_UNSUPPORTED = -4, PA_ERR_IO = -5

   TYPE, BIND(C) :: pa_type                      ! molecule_list(n), Module_Molecular.f90:374-398
      INTEGER(C_INT32_T) :: charge, natoms, nmol, pad0
      REAL(C_DOUBLE)     :: mass
      REAL(C_FLOAT)      :: realcharge
      INTEGER(C_INT32_T) :: pad1
      TYPE(C_PTR)        :: atom_mass, atom_charge     ! REAL(8) (natoms)
      TYPE(C_PTR)        :: xyz                        ! REAL(4) trajectory(atom,mol,step)%coordinates = C order [step][mol][atom][3]
   END TYPE pa_type
   TYPE, BIND(C) :: pa_traj
      INTEGER(C_INT32_T) :: ntypes, nsteps
      TYPE(C_PTR)        :: types                      ! pa_type(ntypes)
      REAL(C_DOUBLE)     :: box_lo(3), box_hi(3), max_dist_sq
      INTEGER(C_INT32_T) :: com_boost, wrap_trajectory
   END TYPE pa_traj
   TYPE, BIND(C) :: pa_job_request
      INTEGER(C_INT32_T) :: mode, ref_xyz(3), ref_type, obs_type, bin_a, bin_b, sampling_interval
      INTEGER(C_INT32_T) :: weigh_charge, use_coc, subtract_uniform, maxdist_optimize
      REAL(C_FLOAT)      :: maxdist
      INTEGER(C_INT32_T) :: normalise_clm
   END TYPE pa_job_request
   TYPE, BIND(C) :: pa_job
      INTEGER(C_INT32_T) :: mode, ref_type, obs_type, ref_xyz(3)
      REAL(C_DOUBLE)     :: ref_vec(3), rot(9)
      REAL(C_FLOAT)      :: maxdist, step_a, step_b, shift
      INTEGER(C_INT32_T) :: bin_a, bin_b, sampling_interval, weigh_charge, use_coc, subtract_uniform, aligned, normalise_clm
   END TYPE pa_job
   TYPE, BIND(C) :: pa_exec
      INTEGER(C_INT32_T) :: device, shard_rank, shard_count, frames_per_batch, flags
      INTEGER(C_INT32_T) :: tile_shard_rank, tile_shard_count, ndevices
   END TYPE pa_exec
   TYPE, BIND(C) :: pa_stats
      INTEGER(C_INT64_T) :: pairs_evaluated, pairs_binned, frames_done, kernel_launches, exception_pairs, h2d_bytes, d2h_bytes
      REAL(C_DOUBLE)     :: kernel_ms, total_ms
      INTEGER(C_INT64_T) :: cell_pair_evals, smem_atomics, unique_pairs_counted, dominant_launches
      REAL(C_DOUBLE)     :: dominant_ms
      INTEGER(C_INT64_T) :: dominant_pair_evals, dominant_atomics
   END TYPE pa_stats
   TYPE, BIND(C) :: pa_dist_subjob
      INTEGER(C_INT32_T) :: n_ref, n_obs
      TYPE(C_PTR)        :: ref, obs                   ! INTEGER(4) (2,n): (molecule type, atom index), 1-based
      INTEGER(C_INT64_T) :: weight
   END TYPE pa_dist_subjob
   TYPE, BIND(C) :: pa_dist_job
      INTEGER(C_INT32_T) :: intermolecular, nsubjobs
      TYPE(C_PTR)        :: subjobs
      INTEGER(C_INT32_T) :: nsteps, sampling_interval, calc_exp, calc_ffc, calc_stdev
      REAL(C_FLOAT)      :: maxdist, exponent
   END TYPE pa_dist_job
   TYPE, BIND(C) :: pa_dist_result
      REAL(C_DOUBLE)     :: closest_average, closest_stdev, exponential_average, exponential_stdev, exponential_weight
      REAL(C_DOUBLE)     :: ffc_average, ffcexp_average, ffc_weight
      INTEGER(C_INT64_T) :: missed_neighbours
   END TYPE pa_dist_result
   TYPE, BIND(C) :: pa_contact_result
      REAL(C_FLOAT)      :: largest_intramolecular, smallest_intramolecular, smallest_intermolecular
      INTEGER(C_INT32_T) :: pad
      INTEGER(C_INT64_T) :: pairs_evaluated
   END TYPE pa_contact_result
   TYPE, BIND(C) :: pa_neigh_request
      INTEGER(C_INT32_T) :: timestep, n_a, n_b
      TYPE(C_PTR)        :: a, b                       ! INTEGER(4) (3,n): (molecule type, molecule index, atom index)
      TYPE(C_PTR)        :: a_class, b_class           ! optional INTEGER(4) (n), 0-based rows / columns of cutoff_sq
      INTEGER(C_INT32_T) :: n_class_a, n_class_b
      TYPE(C_PTR)        :: cutoff_sq                  ! REAL(4) (n_class_b, n_class_a)
      INTEGER(C_INT32_T) :: same_list, skip_same_molecule
   END TYPE pa_neigh_request
   TYPE, BIND(C) :: pa_neigh_pair
      INTEGER(C_INT32_T) :: ia, ib                     ! 0-based positions in the lists
   END TYPE pa_neigh_pair

   INTERFACE
      INTEGER(C_INT) FUNCTION pa_abi_version() BIND(C, NAME="pa_abi_version")
         IMPORT :: C_INT
      END FUNCTION pa_abi_version
      INTEGER(C_INT) FUNCTION pa_device_count() BIND(C, NAME="pa_device_count")
         IMPORT :: C_INT
      END FUNCTION pa_device_count
      ! step sizes, shift, normalised reference vector, rotation (Module_Distribution.f90:353-377, 656-675)
      INTEGER(C_INT) FUNCTION pa_job_setup(traj, req, job) BIND(C, NAME="pa_job_setup")
         IMPORT :: C_INT, pa_traj, pa_job_request, pa_job
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_job_request), INTENT(IN) :: req
         TYPE(pa_job), INTENT(OUT) :: job
      END FUNCTION pa_job_setup
      ! replaces CALL iterate_timesteps_parallelised (Module_Distribution.f90:691)
      INTEGER(C_INT) FUNCTION pa_distribution_histogram(traj, job, hist, within_bin, out_of_bounds) &
         BIND(C, NAME="pa_distribution_histogram")
         IMPORT :: C_INT, C_INT64_T, pa_traj, pa_job
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_job),  INTENT(IN) :: job
         INTEGER(C_INT64_T), INTENT(OUT) :: hist(*)    ! (bin_a,bin_b), a fastest = Fortran layout
         INTEGER(C_INT64_T), INTENT(OUT) :: within_bin, out_of_bounds
      END FUNCTION pa_distribution_histogram
      ! several reference lines in one pass, chosen device(s) / shards; hists = array of C_LOC(hist_k)
      INTEGER(C_INT) FUNCTION pa_distribution_histogram_multi(traj, jobs, njobs, exec, hists, within_bin, out_of_bounds, stats) &
         BIND(C, NAME="pa_distribution_histogram_multi")
         IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_PTR, pa_traj, pa_job, pa_exec, pa_stats
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_job),  INTENT(IN) :: jobs(*)
         INTEGER(C_INT32_T), VALUE :: njobs
         TYPE(pa_exec), INTENT(IN) :: exec
         TYPE(C_PTR), INTENT(IN) :: hists(*)
         INTEGER(C_INT64_T), INTENT(OUT) :: within_bin(*), out_of_bounds(*)
         TYPE(pa_stats), INTENT(OUT) :: stats
      END FUNCTION pa_distribution_histogram_multi
      ! compute_inter/intramolecular_distances + post-processing (Module_Distances.f90:856-1091, 1161-1243)
      INTEGER(C_INT) FUNCTION pa_distance_subjobs(traj, job, exec, results) BIND(C, NAME="pa_distance_subjobs")
         IMPORT :: C_INT, pa_traj, pa_dist_job, pa_exec, pa_dist_result
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_dist_job), INTENT(IN) :: job
         TYPE(pa_exec), INTENT(IN) :: exec
         TYPE(pa_dist_result), INTENT(OUT) :: results(*)
      END FUNCTION pa_distance_subjobs
      ! give_intramolecular_distances + give_intermolecular_contact_distance (Module_Molecular.f90:1271-1367)
      INTEGER(C_INT) FUNCTION pa_contact_distance(traj, timestep, molecule_type, exec, res) BIND(C, NAME="pa_contact_distance")
         IMPORT :: C_INT, C_INT32_T, pa_traj, pa_exec, pa_contact_result
         TYPE(pa_traj), INTENT(IN) :: traj
         INTEGER(C_INT32_T), VALUE :: timestep, molecule_type
         TYPE(pa_exec), INTENT(IN) :: exec
         TYPE(pa_contact_result), INTENT(OUT) :: res
      END FUNCTION pa_contact_distance
      ! the cutoff test of the pair loops of Module_Cluster.f90:1092-1100 / Module_Speciation.f90:1407-1431
      INTEGER(C_INT) FUNCTION pa_neighbour_pairs(traj, req, exec, pairs, capacity, count, tests) BIND(C, NAME="pa_neighbour_pairs")
         IMPORT :: C_INT, C_INT64_T, pa_traj, pa_neigh_request, pa_exec, pa_neigh_pair
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_neigh_request), INTENT(IN) :: req
         TYPE(pa_exec), INTENT(IN) :: exec
         TYPE(pa_neigh_pair), INTENT(OUT) :: pairs(*)
         INTEGER(C_INT64_T), VALUE :: capacity
         INTEGER(C_INT64_T), INTENT(OUT) :: count, tests
      END FUNCTION pa_neighbour_pairs
      ! transfer_histogram / write_distribution_functions (Module_Distribution.f90:897-977, 982-1185)
      INTEGER(C_INT) FUNCTION pa_transfer_histogram(traj, job, hist, within_bin, out_of_bounds, verbose, g, ideal_density) &
         BIND(C, NAME="pa_transfer_histogram")
         IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_FLOAT, pa_traj, pa_job
         TYPE(pa_traj), INTENT(IN) :: traj
         TYPE(pa_job),  INTENT(IN) :: job
         INTEGER(C_INT64_T), INTENT(IN) :: hist(*)
         INTEGER(C_INT64_T), VALUE :: within_bin, out_of_bounds
         INTEGER(C_INT32_T), VALUE :: verbose
         REAL(C_FLOAT), INTENT(OUT) :: g(*), ideal_density
      END FUNCTION pa_transfer_histogram
      ! the file-level drop-in: runs a general.inp for the keywords of this path (Module_Main.f90:2882-3666)
      INTEGER(C_INT) FUNCTION pa_run_general_input(general_inp_path, exec) BIND(C, NAME="pa_run_general_input")
         IMPORT :: C_INT, C_CHAR, pa_exec
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: general_inp_path(*)      ! NUL-terminated
         TYPE(pa_exec), INTENT(IN) :: exec
      END FUNCTION pa_run_general_input
      SUBROUTINE pa_trim() BIND(C, NAME="pa_trim")
      END SUBROUTINE pa_trim
   END INTERFACE
END MODULE prealpha_b200_bindc
