/*
 * prealpha_b200 -- C ABI of the B200-native pair-distribution path.
 *
 * Drop-in boundary for ONE path of FPhilippi/prealpha: the `distribution`
 * analysis (Module_Distribution: pdf -> RDF trace / number integral, 2-D
 * distance-angle pdf, cylindrical cdf) and the `distance` analysis
 * (Module_Distances) that shares its minimum-image primitive.
 *
 * The reference is a single Fortran translation unit without any FFI; its
 * modules exchange state through module variables.  Every entry point below
 * names the reference routine whose state/behaviour it replaces.  Paths are
 * relative to the reference checkout, shorthand as in SURVEY.md:
 *   Distribution:n = Development/Module_Distribution.f90:n
 *   Molecular:n    = Development/Module_Molecular.f90:n
 *   Distances:n    = Development/Module_Distances.f90:n
 *   Angles:n       = Development/Module_Angles.f90:n
 *   SETTINGS:n     = Development/Module_SETTINGS.f90:n
 *   Main:n         = Development/Module_Main.f90:n
 *
 * Conventions
 *   - plain C, plain pointers and sizes; the caller owns every buffer, the
 *     library never frees or retains caller memory after a call returns
 *     (sessions copy what they need to the device).
 *   - return value: 0 on success, otherwise a prealpha error code
 *     (report_error numbers, SETTINGS:383-1037) or a PA_ERR_* value < 0 for
 *     conditions the reference has no code for (no CUDA device, ...).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     PA_ERR_NO_DEVICE when no sm_100 device is usable.
 *   - indices in structs are 1-based where the reference's input files are
 *     1-based (molecule type indices), 0-based for C arrays.
 */
#ifndef PREALPHA_B200_H
#define PREALPHA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 2

/* operation modes, Distribution:343-347 ("cdf|cydf", "pdf|podf") */
#define PA_MODE_CDF 0
#define PA_MODE_PDF 1
#define PA_MODE_CHARGE_ARM 2          /* "charge_arm": polar distribution of the charge arm / dipole vector (O(N) per frame) */

/* library-specific failures (negative; prealpha's own codes are >= 0) */
#define PA_ERR_NO_DEVICE   (-1)  /* no usable CUDA device / kernel image          */
#define PA_ERR_CUDA        (-2)  /* a CUDA runtime call failed (see pa_last_error) */
#define PA_ERR_BAD_ARG     (-3)  /* NULL pointer / inconsistent sizes             */
#define PA_ERR_UNSUPPORTED (-4)  /* reference behaviour is undefined there (see   */
                                 /* randomised reference vector,       */
                                 /* 2-D binning beyond maximum_distance_squared)  */
#define PA_ERR_IO          (-5)

/* prealpha error codes that can be returned on this path (SETTINGS:383-1037) */
#define PA_E_TYPE_INVALID       33  /* Distribution:415-427 */
#define PA_E_TRAJ_TOO_SHORT     37  /* Distribution:1191    */
#define PA_E_NO_BOX             41  /* Main:3487            */
#define PA_E_DIST_INPUT_FORMAT  99  /* Distribution:342-344,414 */
#define PA_E_BIN_OVERFLOW      101  /* Distribution:904-905 */
#define PA_E_NO_DIST_INPUT     132  /* Distribution:396     */

/* One molecule type: molecule_list(n) of Molecular:374-398.
 * xyz is trajectory(atom,mol,step)%coordinates, i.e. C order
 * [step][mol][atom][3] float32 (Molecular:366-368, column-major Fortran). */
typedef struct pa_type {
    int32_t charge;            /* %charge                                   */
    int32_t natoms;            /* %number_of_atoms                          */
    int32_t nmol;              /* %total_molecule_count                     */
    int32_t _pad0;
    double  mass;              /* %mass (REAL8 sum of REAL4 element masses) */
    float   realcharge;        /* %realcharge (REAL4)                       */
    int32_t _pad1;
    const double *atom_mass;   /* %list_of_atom_masses(natoms), REAL8       */
    const double *atom_charge; /* %list_of_atom_charges(natoms), REAL8      */
    const float  *xyz;         /* [nsteps][nmol][natoms][3]                 */
} pa_type;

/* The loaded trajectory + box: module variables of MOLECULAR. */
typedef struct pa_traj {
    int32_t ntypes;            /* number_of_molecule_types                  */
    int32_t nsteps;            /* number_of_steps                           */
    const pa_type *types;
    double  box_lo[3];         /* box_dimensions(1,:)  Molecular:399        */
    double  box_hi[3];         /* box_dimensions(2,:)                       */
    double  max_dist_sq;       /* maximum_distance_squared (Molecular:938,  */
                               /* 22500 for xyz: Molecular:4883)            */
    int32_t com_boost;         /* use_firstatom_as_com  Molecular:3556-3562 */
    int32_t wrap_trajectory;   /* WRAP_TRAJECTORY (global switch, Main:260) */
} pa_traj;

/* What the user writes in the distribution input file for ONE reference line
 * plus the keyword block (Distribution:400-616).  pa_job_setup() turns it
 * into a pa_job with the reference's REAL(4) arithmetic. */
typedef struct pa_job_request {
    int32_t mode;              /* PA_MODE_*                                 */
    int32_t ref_xyz[3];        /* references(1:3,n)  integer vector         */
    int32_t ref_type;          /* references(4,n)  1-based                  */
    int32_t obs_type;          /* references(5,n)  1-based, <1 = all types  */
    int32_t bin_a, bin_b;      /* bin_count_a / bin_count_b (1..1000)       */
    int32_t sampling_interval;
    int32_t weigh_charge;      /* logical                                   */
    int32_t use_coc;           /* logical                                   */
    int32_t subtract_uniform;  /* logical (only affects output, row W)      */
    int32_t maxdist_optimize;  /* 1: maxdist = REAL4(box_limit/2.0); charge_arm: from the largest arm at step 1 (Distribution:1232-1269) */
    float   maxdist;           /* used when maxdist_optimize == 0           */
    int32_t normalise_clm;     /* charge_arm only: divide by mass * radius of gyration^2 (Distribution:864-867) */
} pa_job_request;

/* Fully resolved job: the locals of make_distribution_functions
 * (Distribution:636-692) as the hot loop reads them. */
typedef struct pa_job {
    int32_t mode;
    int32_t ref_type, obs_type;      /* 1-based; obs_type < 1 = all types   */
    int32_t ref_xyz[3];
    double  ref_vec[3];              /* normalised reference_vector         */
    double  rot[9];                  /* ANGLES rotation, row-major R(k,:)   */
    float   maxdist, step_a, step_b, shift;   /* REAL4, Distribution:353-377,656 */
    int32_t bin_a, bin_b, sampling_interval;
    int32_t weigh_charge, use_coc, subtract_uniform;
    int32_t aligned;                 /* prepare_rotation's `aligned`        */
    int32_t normalise_clm;           /* charge_arm only                     */
} pa_job;

/* How to run: which device(s), and which shard of the work this process owns.
 * Frames are independent units (Distribution:717-718 runs them SCHEDULE(STATIC,1)
 * over OpenMP threads); here frame shard r of n takes sampled frame k when
 * k % n == r.  A single huge frame gives that loop no parallelism, so a second,
 * independent split exists below the frame: tile shard r of n takes the work
 * items (tile of reference molecules, slab of observed cells) w with w % n == r
 * (cell-sorted kernels) or the tiles of reference molecules t with t % n == r
 * (brute-force kernels).  Histograms of different shards add (integers), in
 * any combination of the two splits.
 * ndevices > 1 (one-shot entry points and the file-level drop-in): the call
 * itself drives devices device .. device+ndevices-1 with one host thread each --
 * the analogue of `set_threads` (Main:3313-3330) for the parallel region
 * Distribution:704-753 -- sharding frames when there are at least ndevices
 * sampled frames and tiles otherwise, and combines the int64 accumulators on
 * `device` over NVLink: peer copies + an add kernel for the usual few KB, one
 * ncclReduce from 4 MiB of accumulators on (environment PA_NCCL_MIN_BYTES).
 * ndevices < 0: all usable devices.  The caller's
 * shard_* / tile_shard_* then describe this process's share of a larger job
 * and are subdivided further. */
typedef struct pa_exec {
    int32_t device;            /* CUDA ordinal (first device when ndevices > 1) */
    int32_t shard_rank;        /* frame shard 0..shard_count-1              */
    int32_t shard_count;       /* >= 1 (0 is read as 1)                     */
    int32_t frames_per_batch;  /* 0 = choose                                */
    int32_t flags;             /* PA_EXEC_*                                 */
    int32_t tile_shard_rank;   /* tile shard 0..tile_shard_count-1          */
    int32_t tile_shard_count;  /* >= 1 (0 is read as 1)                     */
    int32_t ndevices;          /* 0/1: one device; N: N devices; <0: all    */
} pa_exec;

#define PA_EXEC_FORCE_GENERAL  1   /* use the general (2-D capable) kernel even for RDF jobs */
#define PA_EXEC_NO_CELLS       2   /* disable the cell-sorted tile path (brute force tiles)   */
#define PA_EXEC_NO_SYMMETRY    4   /* evaluate every ordered pair even when (i,j)/(j,i) are provably bit-identical */
#define PA_EXEC_NO_F32         8   /* cell kernel: approximate classes from fp64 differences only (no float32 segments) */

/* Per-call statistics (all optional outputs). */
typedef struct pa_stats {
    int64_t pairs_evaluated;   /* ordered (ref,obs,frame) triples the reference's loop visits */
    int64_t pairs_binned;      /* = within_bin                                                */
    int64_t frames_done;
    int64_t kernel_launches;   /* launches of OUR kernels                                     */
    int64_t exception_pairs;   /* pairs re-evaluated by the exact per-pair kernel             */
    int64_t h2d_bytes, d2h_bytes;
    double  kernel_ms;         /* CUDA-event time of the pair kernels on their stream         */
    double  total_ms;          /* CUDA-event time upload..download                            */
    /* counted by the cell-sorted RDF kernels themselves (0 when another kernel did the work): */
    int64_t cell_pair_evals;   /* pair distances actually evaluated after cell pruning, each unordered pair of a
                                  symmetric / merged pass once                                */
    int64_t smem_atomics;      /* shared-memory atomic increments issued (lanes)              */
    int64_t unique_pairs_counted; /* pairs that landed in a bin, each evaluated pair once     */
    int64_t dominant_launches; /* launches of the float32 cell-sorted RDF kernel (the dominant kernel of large RDF jobs) */
    double  dominant_ms;       /* CUDA-event time of those launches alone                     */
    int64_t dominant_pair_evals, dominant_atomics;   /* their share of cell_pair_evals / smem_atomics */
} pa_stats;

/* ---- library ---------------------------------------------------------- */
int         pa_abi_version(void);
const char *pa_last_error(void);             /* thread-local message of the last failure */
int         pa_device_count(void);           /* usable sm_100 devices, 0 if none         */

/* ---- row S / Angles: job set-up (host, float32 exact) -------------------
 * Replaces Distribution:353-377 (step sizes), :526-527 (maxdist_optimize),
 * :656-675 (shift, normalised reference vector, prepare_rotation Angles:64). */
int pa_job_setup(const pa_traj *traj, const pa_job_request *req, pa_job *job);

/* ---- rows C,C',Wv,D,F,I: THE hot path -----------------------------------
 * Replaces the body of make_distribution_functions between histogram zeroing
 * and transfer_histogram (Distribution:691-692): iterate_timesteps_parallelised
 * (:696-760) + fill_histogram (:762-845) + give_smallest_distance_squared
 * (Molecular:1430-1469) + give_center_of_mass / _charge (Molecular:2804, 2953)
 * + wrap_vector (Molecular:1560).
 * hist has bin_a*bin_b entries, a fastest (= Fortran (bin_a,bin_b) layout) and
 * is overwritten.  traj->types[*].xyz are HOST pointers; frames are staged
 * through pinned buffers and copied with cudaMemcpyAsync on a side stream. */
int pa_distribution_histogram(const pa_traj *traj, const pa_job *job,
                              int64_t *hist, int64_t *within_bin, int64_t *out_of_bounds);

/* Same for several jobs over the same trajectory in one pass (frames are
 * uploaded and pre-processed once), on a chosen device / frame shard.
 * hists[k] has jobs[k].bin_a*jobs[k].bin_b entries. */
int pa_distribution_histogram_multi(const pa_traj *traj, const pa_job *jobs, int32_t njobs,
                                    const pa_exec *exec,
                                    int64_t *const *hists, int64_t *within_bin,
                                    int64_t *out_of_bounds, pa_stats *stats);

/* ---- device-resident sessions (streaming + benchmarking) ----------------
 * A session owns device memory for a batch of frames.  upload copies frames
 * [first_step, first_step+nframes) of `traj` (host pointers) to HBM; run
 * executes the jobs on the resident frames and ACCUMULATES into the session's
 * device histograms; download copies hist/within/oob back (and optionally
 * zeroes the accumulators).  This is the additive streaming API of SURVEY 8b
 * (pa_begin / pa_push_frames / pa_finish). */
typedef struct pa_session pa_session;
int  pa_session_create(const pa_traj *traj_meta, const pa_job *jobs, int32_t njobs,
                       const pa_exec *exec, int32_t max_frames, pa_session **out);
int  pa_session_upload(pa_session *s, const pa_traj *traj, int32_t first_step, int32_t nframes,
                       int32_t step_stride);
int  pa_session_run(pa_session *s);                                  /* all resident frames   */
int  pa_session_run_frames(pa_session *s, int32_t first, int32_t count); /* a resident sub-range   */
int  pa_session_download(pa_session *s, int64_t *const *hists, int64_t *within_bin,
                         int64_t *out_of_bounds, int32_t reset);
int  pa_session_stats(pa_session *s, pa_stats *stats);
/* raw device pointer + length (int64 words) of the accumulators of all jobs:
 * layout per job = [bin_a*bin_b hist][within][oob]; lets the caller reduce
 * across GPUs with NCCL without a host round trip. */
int  pa_session_device_accumulators(pa_session *s, void **dev_ptr, int64_t *nwords);
void pa_session_destroy(pa_session *s);

/* ---- row N: normalisation (host, float32 exact) -------------------------
 * Replaces transfer_histogram (Distribution:897-977).  g has bin_a*bin_b
 * floats (a fastest).  ideal_density receives the value the writer uses
 * (after the VERBOSE_OUTPUT rescale of :928-930 when verbose != 0). */
int pa_transfer_histogram(const pa_traj *traj, const pa_job *job, const int64_t *hist,
                          int64_t within_bin, int64_t out_of_bounds, int32_t verbose,
                          float *g, float *ideal_density);

/* ---- row W: output files -------------------------------------------------
 * Replaces write_distribution_functions (Distribution:982-1185).
 * path_prefix = TRIM(PATH_OUTPUT)//TRIM(ADJUSTL(OUTPUT_PREFIX)); g is modified
 * in place when subtract_uniform is set, exactly like the reference. */
int pa_write_distribution(const pa_traj *traj, const pa_job *job, int32_t reference_number,
                          const char *path_prefix, const char *input_file_name,
                          float *g, float ideal_density);

/* gfortran list-directed REAL(4) item (17 characters, no terminator added
 * beyond NUL); buf must hold >= 32 bytes.  Returns the length written. */
int pa_format_real4(float value, char *buf);

/* ---- row X: `distance` analysis (Module_Distances) -----------------------
 * Replaces compute_intermolecular_distances / compute_intramolecular_distances
 * (Distances:976-1091 / 856-974) including the second pass for the standard deviations
 * and the post-processing of perform_distance_analysis (Distances:1161-1243).
 * A subjob is one line of the distance input after wildcard expansion
 * (prepare_subjob, Distances:500-609): the reference atoms indices_1 and the observed
 * atoms indices_2 as (molecule type, atom index) pairs, both 1-based.
 * The GPU evaluates the minimum-image distances and returns those below the cutoff in the
 * reference's loop order; the sums of exp / r^-3 / r^-6 terms are then formed on the host as
 * the same sequential chain of REAL(8) additions, so every result is bit-identical. */
typedef struct pa_dist_subjob {
    int32_t n_ref, n_obs;
    const int32_t *ref;           /* [n_ref][2] = (molecule type, atom index)   indices_1       */
    const int32_t *obs;           /* [n_obs][2]                                  indices_2       */
    int64_t weight;               /* reference atoms expected per step (Distances:810-845)        */
} pa_dist_subjob;
typedef struct pa_dist_job {
    int32_t intermolecular;       /* 1 = "inter", 0 = "intra"  (Distances:314-320)                */
    int32_t nsubjobs;
    const pa_dist_subjob *subjobs;
    int32_t nsteps;               /* steps 1..nsteps are used (default 1, Distances:9)            */
    int32_t sampling_interval;
    int32_t calc_exp, calc_ffc, calc_stdev;
    float   maxdist;              /* REAL4 cutoff; maxdist_squared = maxdist**2 in REAL4          */
    float   exponent;             /* k of exp(-k r), REAL4                                        */
} pa_dist_job;
typedef struct pa_dist_result {   /* list_of_distances(n)%... after post-processing               */
    double closest_average, closest_stdev;
    double exponential_average, exponential_stdev, exponential_weight;
    double ffc_average, ffcexp_average, ffc_weight;
    int64_t missed_neighbours;
} pa_dist_result;
int pa_distance_subjobs(const pa_traj *traj, const pa_dist_job *job, const pa_exec *exec, pa_dist_result *results);

/* ---- contact distances --------------------------------------
 * Replaces contact_distance / print_distances (Debug:586-623) for one molecule
 * type at one time step: give_intramolecular_distances (Molecular:1271-1292) and
 * give_intermolecular_contact_distance (Molecular:1295-1367), both built on
 * give_smallest_atom_distance_squared (Molecular:1380-1427).  The intermolecular
 * search is the reference's O(N_atoms(type) * N_atoms(all)) brute force; the
 * results are the REAL(4) values the reference prints with F0.3.
 * timestep and type are 1-based; the caller applies check_timestep (Debug:46-56). */
typedef struct pa_contact_result {
    float largest_intramolecular, smallest_intramolecular, smallest_intermolecular;
    int32_t pad_;
    int64_t pairs_evaluated;      /* atom pairs of the intermolecular search */
} pa_contact_result;
int pa_contact_distance(const pa_traj *traj, int32_t timestep, int32_t type, const pa_exec *exec, pa_contact_result *result);

/* Page-locked staging buffers of finished sessions are kept for the next call of the same size (at most four buffers,
 * 1 GiB); pa_trim() gives them back to the system. */
void pa_trim(void);

/* ---- neighbour pairs -----------------------------------------
 * The distance test that the cluster and speciation analyses run for every pair
 * of listed atoms: give_smallest_atom_distance_squared(...) < squared cutoff in
 * REAL(4) (Cluster:1094-1100 with one cutoff per pass; Speciation:1407-1431 with
 * one cutoff per (acceptor atom, donor atom) combination).  Returns every
 * qualifying pair of atom instances, sorted by (ia, ib) = the order in which the
 * reference's nested loops meet them, so that the sequential bookkeeping
 * (cluster merging, species clipboard) can replay them on the host.
 * Instances are (molecule type, molecule index, atom index) triples, 1-based. */
typedef struct pa_neigh_request {
    int32_t timestep;            /* 1-based                                                     */
    int32_t n_a, n_b;            /* atom instances in list a (outer loop) and b (inner loop)    */
    const int32_t *a, *b;        /* [n][3]                                                      */
    const int32_t *a_class;      /* optional [n_a], 0-based row of cutoff_sq (NULL: row 0)      */
    const int32_t *b_class;      /* optional [n_b], 0-based column of cutoff_sq (NULL: col 0)   */
    int32_t n_class_a, n_class_b;
    const float *cutoff_sq;      /* [n_class_a][n_class_b] REAL(4) squared cutoffs              */
    int32_t same_list;           /* 1: b is a; each unordered pair once, ib > ia (Cluster:1092) */
    int32_t skip_same_molecule;  /* 1: atoms of one molecule are never paired (Speciation:1407) */
} pa_neigh_request;
typedef struct pa_neigh_pair { int32_t ia, ib; } pa_neigh_pair;   /* 0-based positions in the lists */
/* Writes at most `capacity` pairs; *count is the number found (call again with a larger buffer when
 * it exceeds capacity).  *tests = distance tests carried out. */
int pa_neighbour_pairs(const pa_traj *traj, const pa_neigh_request *req, const pa_exec *exec,
                       pa_neigh_pair *out, int64_t capacity, int64_t *count, int64_t *tests);

/* ---- file-level drop-in ----------------------------------
 * Runs a general.inp the way `./prealpha general.inp` does for the keywords
 * of this path (Main:103-357 header, Main:2882-3666 body; unknown keywords
 * are reported as error 20 and skipped, like the reference).  Returns the
 * number of errors/warnings encountered (the reference's final counter), or
 * PA_ERR_CUDA (< 0) when a GPU entry point failed along the way. */
int pa_run_general_input(const char *general_inp_path, const pa_exec *exec);

#ifdef __cplusplus
}
#endif
#endif /* PREALPHA_B200_H */
