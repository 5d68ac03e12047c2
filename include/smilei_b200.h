/* smilei_b200.h — C ABI of the B200-native PIC time-step hot path.
 *
 * This is the drop-in boundary: the entry points a Smilei build would bind for
 * 3D Cartesian gather + push + Esirkepov deposit + Yee FDTD when
 * Main.gpu_computing = True.  Plain C types only; every pointer is documented as a
 * HOST or a DEVICE pointer.  The reference's own precedent for such a boundary is
 *   extern "C" void currentDeposition3DOnDevice(...)  (src/Projector/Projector3D2OrderGPU.cpp:55-88,
 *   defined in src/Projector/Projector3D2OrderGPUKernel.cpp:40-166).
 *
 * Error convention: every function returns 0 on success and non-zero on failure;
 * sb200_last_error() returns a description.  Nothing aborts (the reference's ERROR()
 * macro raises SIGABRT, src/Tools/Tools.h:127-133 — the C++ adapter in
 * include/smilei_b200_operators.hpp maps a non-zero return onto it).
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * All `path:line` citations are relative to the reference tree (/root/reference).
 */
#ifndef SMILEI_B200_H
#define SMILEI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 1

/* Field identifiers; layout of each = the reference's Field3D (src/Field/Field3D.cpp:177-216):
 * contiguous double[nx*ny*nz], index (i*ny+j)*nz+k, dims per component as in
 * src/ElectroMagn/ElectroMagn3D.cpp:115-123,156-159. */
enum sb200_field {
    SB200_EX = 0, SB200_EY, SB200_EZ,
    SB200_BX, SB200_BY, SB200_BZ,
    SB200_BXM, SB200_BYM, SB200_BZM,
    SB200_JX, SB200_JY, SB200_JZ,
    SB200_RHO,
    SB200_NFIELDS
};

/* Species.pusher values of the namelist handled here (src/Pusher/PusherFactory.h:49-70). */
enum sb200_pusher { SB200_PUSHER_BORIS = 0, SB200_PUSHER_VAY = 1, SB200_PUSHER_HIGUERACARY = 2 };

/* Geometry of one patch = the constants the reference operators capture at construction
 * (src/Interpolator/Interpolator3D2Order.cpp:15-22, src/Pusher/Pusher.cpp:5-29,
 *  src/Projector/Projector3D2Order.cpp:18-42, src/ElectroMagnSolver/Solver3D.h:14-24,
 *  src/Patch/Patch.cpp:136-165). */
typedef struct sb200_grid {
    int    n[3];            /* Params::patch_size_ : cells of this patch per dimension          */
    int    oversize[3];     /* Params::oversize    : ghost cells = interpolation order          */
    double cell_length[3];  /* Params::cell_length                                              */
    double dt;              /* Params::timestep                                                 */
    int    pcoord[3];       /* Patch::Pcoordinates : position of the patch in the patch grid    */
    int    npatch[3];       /* Params::number_of_patches                                        */
    int    interp_order;    /* Main.interpolation_order : 2 or 4                                */
} sb200_grid;

typedef struct sb200_patch sb200_patch;   /* opaque: device fields + per-species SoA + scratch  */

const char *sb200_last_error( void );
int  sb200_abi_version( void );
int  sb200_device_count( int *count );

/* ---- lifetime ------------------------------------------------------------------------ */
/* replaces PatchesFactory::create -> Patch3D + ElectroMagn3D + Species allocation
 * (src/Smilei.cpp:276, src/ElectroMagn/ElectroMagn3D.cpp:88-230). */
int  sb200_patch_create( sb200_patch **out, const sb200_grid *grid, int n_species, int device );
int  sb200_patch_destroy( sb200_patch *p );
/* cudaStream_t every later call is enqueued on (NULL = legacy default stream). */
int  sb200_patch_set_stream( sb200_patch *p, void *cuda_stream );
int  sb200_patch_synchronize( sb200_patch *p );

/* ---- species ------------------------------------------------------------------------- */
/* Species::mass_, Species.pusher, and the SoA capacity to reserve
 * (src/Species/Species.cpp:260-307 initOperators, src/Particles/Particles.h:526-566). */
int  sb200_species_config( sb200_patch *p, int ispec, double mass, int pusher, size_t capacity );
/* Particle boundary conditions of the species at the six GLOBAL box sides, order xmin xmax ymin ymax zmin zmax
 * (Species::boundary_conditions_, src/ParticleBC/PartBoundCond.cpp:99-245): SB200_PBC_PERIODIC (exchange with
 * the neighbour across the box) or SB200_PBC_REMOVE (remove_particle_inf/sup,
 * src/ParticleBC/BoundaryConditionType.cpp:204-294).  Sides of the patch that are not global box sides always
 * tag for exchange (internal_inf/sup).  Default: all periodic. */
#define SB200_PBC_PERIODIC 0
#define SB200_PBC_REMOVE   1
int  sb200_species_set_bc( sb200_patch *p, int ispec, const int bc[6] );
/* mass * sum of w*(gamma-1) over the particles removed at the boundaries since the last reset
 * (Species::nrj_bc_lost, src/Species/Species.cpp:757-762). */
int  sb200_species_lost_energy( sb200_patch *p, int ispec, double *lost, int reset );
/* HOST -> device import of the SoA columns (Particles::Position/Momentum/Weight/Charge). */
int  sb200_species_set( sb200_patch *p, int ispec,
                        const double *x, const double *y, const double *z,
                        const double *px, const double *py, const double *pz,
                        const double *w, const short *q, size_t n );
/* device -> HOST export; any pointer may be NULL; `keys` = Particles::cell_keys. */
int  sb200_species_get( sb200_patch *p, int ispec,
                        double *x, double *y, double *z, double *px, double *py, double *pz,
                        double *w, short *q, int *keys, size_t n );
int  sb200_species_count( sb200_patch *p, int ispec, size_t *n );
/* DEVICE pointers to the live SoA columns, the seam the reference exposes through
 * Particles::getPtrPosition/Momentum/Weight/Charge/CellKeys (src/Particles/Particles.h:441-478).
 * column: 0..2 position, 3..5 momentum, 6 weight, 7 charge (short*), 8 cell_keys (int*). */
int  sb200_species_device_ptr( sb200_patch *p, int ispec, int column, void **dev_ptr );
/* first_index / last_index per primal-node cell after sb200_sort (Particles::first_index,
 * src/Species/SpeciesV.cpp:645-652): HOST int[ncells+1], ncells = (n0+1)(n1+1)(n2+1). */
int  sb200_species_first_index( sb200_patch *p, int ispec, int *first, size_t n );

/* ---- fields -------------------------------------------------------------------------- */
int  sb200_field_size( sb200_patch *p, int field_id, size_t *n, int dims[3] );
int  sb200_field_set( sb200_patch *p, int field_id, const double *host, size_t n );
int  sb200_field_get( sb200_patch *p, int field_id, double *host, size_t n );
/* DEVICE pointer + padded shape of a field array (the seam the reference exposes through
 * Field::data_ mapped with GetDevicePointer, src/Tools/gpu.h:18-44).  Element (i,j,k) is at
 * (i*alloc[1]+j)*alloc[2]+k; alloc may be NULL. */
int  sb200_field_device_ptr( sb200_patch *p, int field_id, void **dev_ptr, int alloc[3] );

/* ---- the time step, in the order of src/Smilei.cpp:519-649 ----------------------------- */
/* ElectroMagn::restartRhoJ (src/ElectroMagn/ElectroMagn.cpp:402-408): Jx,Jy,Jz,rho <- 0. */
int  sb200_restart_rhoJ( sb200_patch *p );

/* ElectroMagn::Jx_s / Jy_s / Jz_s / rho_s of one species (src/ElectroMagn/ElectroMagn.h:129-141): the arrays a field
 * diagnostic of that species asks for.  mask bit k (Jx_s Jy_s Jz_s rho_s) allocates / frees array k.  While a species
 * owns an array, sb200_dynamics with SB200_DYN_DIAG_RHO deposits that component THERE instead of into the total
 * (Projector3D2Order.cpp:756-763, Projector3D4Order.cpp:706-713); sb200_restart_rhoJ clears it
 * (ElectroMagn::restartRhoJs, ElectroMagn.cpp:410-436).  The arrays are addressed by every sb200_field_* and sb200_halo_*
 * call through the field id SB200_SPECIES_FIELD(ispec, k). */
#define SB200_SPECIES_FIELD( ispec, k ) ( SB200_NFIELDS + 4*( ispec ) + ( k ) )
int  sb200_species_diag_fields( sb200_patch *p, int ispec, int mask );
/* ElectroMagn3D::computeTotalRhoJ (src/ElectroMagn/ElectroMagn3D.cpp:1753-1799): Jx Jy Jz rho += every species' own
 * arrays (called after the species' dynamics on a diag step, VectorPatch.cpp: computeCharge / sumDensities). */
int  sb200_compute_total_rhoJ( sb200_patch *p );

#define SB200_DYN_KEEP_SCRATCH 1   /* also write Epart/Bpart/invgf/iold/deltaold (SmileiMPI.h:213-221) */
#define SB200_DYN_DIAG_RHO     2   /* diag step: also deposit rho (Projector3D2Order.cpp:349-521)       */
/* Species::dynamics for one species (src/Species/Species.cpp:524-875): fused
 *   Interpolator3D{2,4}Order::fieldsWrapper + Pusher{Boris,Vay,HigueraCary}::operator()
 *   + PartBoundCond::apply (internal_inf/sup tagging) + SpeciesV::computeParticleCellKeys
 *   + Projector3D{2,4}Order::currentsAndDensityWrapper.
 * Requires the species to be cell-sorted (sb200_sort). */
int  sb200_dynamics( sb200_patch *p, int ispec, int flags );
/* device -> HOST copy of the scratch written with SB200_DYN_KEEP_SCRATCH; component-major,
 * stride n (SmileiMPI::dynamics_*). Any pointer may be NULL. */
int  sb200_scratch_get( sb200_patch *p, double *Epart, double *Bpart, double *invgf,
                        int *iold, double *deltaold, size_t n );

/* VectorPatch::solveMaxwell (src/Patch/VectorPatch.cpp:1013-1023):
 *   ElectroMagn3D::saveMagneticFields + MA_Solver3D_norm + MF_Solver3D_Yee.
 * B_m is left = (B_new+B_old)/2 on points the B halo exchange never overwrites and
 * = B_old on the exchanged ghost planes; sb200_center_B completes those after the
 * exchange, so that after sb200_center_B all of B_m equals the reference's
 * ElectroMagn3D::centerMagneticFields (src/ElectroMagn/ElectroMagn3D.cpp:1191-1293). */
int  sb200_maxwell( sb200_patch *p );
/* Silver-Mueller absorbing / injecting boundary on one GLOBAL box side (i_boundary = 0..5: xmin xmax ymin ymax
 * zmin zmax): ElectroMagnBC3D_SM::apply (src/ElectroMagnBC/ElectroMagnBC3D_SM.cpp:141-376) with zero external
 * fields.  `k` = EM_BCs_k of that side (incidence vector).  db1 / db2 are the summed laser amplitudes
 * Laser::getAmplitude0/1 on the face (HOST arrays of n_p[axis1]*n_d[axis2] and n_d[axis1]*n_p[axis2] doubles,
 * row-major as the reference's b1/b2; NULL = no laser).  is_boundary = { axis1 min, axis1 max, axis2 min,
 * axis2 max }: 1 where the patch has no neighbour on that side of the face's own axes (Patch::isBoundary; the
 * sweeps skip the first / last index there), axis1 = (axis0==0 ? 1 : 0), axis2 = (axis0==2 ? 1 : 2).
 * A no-op on a patch that does not touch the side.  Call after the B halo exchange and before sb200_center_B
 * (Smilei.cpp:649 finalizeSyncAndBCFields). */
int  sb200_apply_SM( sb200_patch *p, int i_boundary, const double k[3], const int is_boundary[4],
                     const double *db1, const double *db2 );
/* Moving window along x (SimWindow::shift, src/MovWindow/SimWindow.cpp:98-550), cell-granular: the patch slides
 * by `ncells` cells.  E, B, B_m move down by ncells planes and the planes entering on the right are zero (a patch
 * created by the window starts with zero fields, SimWindow.cpp:224); the patch origin advances
 * (Patch::initStep3 with n_moved, src/Patch/Patch.cpp:159-163); particles left behind (x < new xmin) are
 * dropped on the patch at the left end of the box and tagged for the -x neighbour elsewhere (then
 * sb200_leaving_count / sb200_leaving_pack(dim 0, side 0) hand them over); every species is marked unsorted.
 * With several patches along x the caller packs the planes the -x neighbour needs BEFORE the shift
 * (sb200_halo_pack, planes [2*oversize+1+dual, +ncells) of E, B, B_m) and unpacks what the +x neighbour sent into
 * the last ncells real planes AFTER it.  The caller then appends the particles of the cells uncovered at the
 * right end of the box (sb200_species_append) and sorts. */
int  sb200_window_shift( sb200_patch *p, int ncells );
/* ParticleCreator on the DEVICE for position_initialization "regular" + momentum_initialization "cold"
 * (src/Particles/ParticleCreator.cpp:627-667, 840-851) with the same count in every kept cell — what a moving window
 * creates every shift in the laser-wake benchmarks (SimWindow.cpp:372-392).  `cells` = flat indices (x slowest) of the
 * kept cells inside `box` cells starting at position `origin`; weight / charge per particle of each kept cell (the
 * density and charge profiles are evaluated by the caller); regular_inv = 1/regular_number or 1/pow(nppc,1/3) as the
 * reference computes it.  Appends ncells * prod(regular_number) particles; the species becomes unsorted. */
int  sb200_species_append_regular( sb200_patch *p, int ispec, const double origin[3], const int box[3],
                                   const int regular_number[3], const double regular_inv[3],
                                   const int *cells, const double *weight, const short *charge, size_t ncells );
/* HOST -> device append of n particles at the end of a species (ParticleCreator::create on the cells a moving
 * window uncovers, SimWindow.cpp:372-392); the species becomes unsorted. */
int  sb200_species_append( sb200_patch *p, int ispec,
                           const double *x, const double *y, const double *z,
                           const double *px, const double *py, const double *pz,
                           const double *w, const short *q, size_t n );
int  sb200_center_B( sb200_patch *p );

/* SpeciesV::computeParticleCellKeys histogram + SpeciesV::sortParticles
 * (src/Species/SpeciesV.cpp:599-855): drops particles with key<0, orders the rest by
 * cell key with a stable counting sort (canonical order), rebuilds first_index. */
int  sb200_sort( sb200_patch *p, int ispec );

/* DiagnosticScalar Ukin per species and Uelm (src/Diagnostic/DiagnosticScalar.cpp:435-587,
 * 658-691; Field3D::norm2 src/Field/Field3D.cpp:230-250).  HOST outputs. */
int  sb200_energy( sb200_patch *p, double *ukin_per_species, double *uelm );

/* ---- halo hooks for the exchange layer (NCCL or an MPI adapter) ------------------------ */
/* Pack `nplanes` planes starting at `first_plane` along `dim` of one field into / out of a
 * contiguous DEVICE buffer (plane-major, then the two other dims in field order).
 * Replaces Field3D::extract_fields_{sum,exch} / inject_fields_{sum,exch}
 * (src/Field/Field3D.h:115-119; slab sizes src/Patch/SyncVectorPatch.cpp:235-237,1483). */
#define SB200_UNPACK_COPY 0
#define SB200_UNPACK_ADD  1
int  sb200_halo_plane_elems( sb200_patch *p, int field_id, int dim, size_t *elems_per_plane );
int  sb200_halo_pack  ( sb200_patch *p, int field_id, int dim, int first_plane, int nplanes, double *dev_buf );
int  sb200_halo_unpack( sb200_patch *p, int field_id, int dim, int first_plane, int nplanes, const double *dev_buf, int mode );
/* same-GPU neighbour (a periodic dimension held by one patch):
 * SyncVectorPatch::sumAllComponents local branch (SyncVectorPatch.cpp:263-311) and
 * exchangeAllComponentsAlong{X,Y,Z} local branch (:1483-1527). */
int  sb200_halo_sum_self( sb200_patch *p, int field_id, int dim );
int  sb200_halo_exchange_self( sb200_patch *p, int field_id, int dim );

/* ---- particle migration hooks (SyncVectorPatch::initExchParticles / finalizeExchParticlesAndSort,
 *      src/Patch/SyncVectorPatch.cpp:27-113; Patch::exchNbrOfParticles src/Patch/Patch.cpp:560-594) ---- */
#define SB200_PARTICLE_RECORD_DOUBLES 8   /* x y z px py pz w q(as double) */
/* counts[2*dim+side] = particles tagged -2-2*dim-side among [0, count). HOST output. */
int  sb200_leaving_count( sb200_patch *p, int ispec, int counts[6] );
/* Pack the particles tagged for (dim, side), in index order, as records of
 * SB200_PARTICLE_RECORD_DOUBLES doubles into a DEVICE buffer; `wrap` is added to the
 * position along `dim` when the particle crosses the global box (Patch::prepareParticles,
 * src/Patch/Patch.cpp:633-650).  *n_packed receives the record count (HOST). */
int  sb200_leaving_pack( sb200_patch *p, int ispec, int dim, int side, double wrap,
                         double *dev_buf, size_t max_records, size_t *n_packed );
/* Same, for a caller that already knows the count (from sb200_leaving_count): no device->host round trip.
 * Fails if the hint exceeds what the tag list can hold (then use sb200_leaving_pack). */
int  sb200_leaving_pack_known( sb200_patch *p, int ispec, int dim, int side, double wrap,
                               double *dev_buf, size_t max_records, size_t n_known );
/* Append `n` records to the species, tagging them again (a corner particle is forwarded
 * in the next dimension, Patch::cornersParticles src/Patch/Patch.cpp:727-800). */
int  sb200_arriving_unpack( sb200_patch *p, int ispec, const double *dev_buf, size_t n );

/* ---- synthetic input (bench / smoke) ------------------------------------------------------ */
/* Fill species `ispec` on the device with a uniform thermal plasma: ppc[0]*ppc[1]*ppc[2]
 * particles per cell at the reference's "regular" positions (ParticleCreator.cpp:661-667),
 * weight = density*cell_volume/nppc (:259,933-939), Maxwellian momenta of temperature T
 * (units of m_e c^2).  The species must be configured with enough capacity; it is left
 * unsorted (call sb200_sort). */
int  sb200_species_init_thermal( sb200_patch *p, int ispec, const int ppc[3], double density, int charge,
                                 double temperature, unsigned long long seed );

/* ---- initial particles on the reference's random streams (HOST, init time only) --- */
/* Hilbert index of the patch at (x,y,z) in a box of 2^m0 x 2^m1 x 2^m2 patches: replaces
 * generalhilbertindex (src/DomainDecomposition/Hilbert_functions.cpp:246-296), which
 * HilbertDomainDecomposition3D::getDomainId calls (HilbertDomainDecomposition.cpp:83-87).  A patch's
 * stream is xorshift32 seeded with random_seed + this index (src/Patch/Patch.cpp:129, Random.h:91-140). */
int  sb200_hilbert_index3d( unsigned m0, unsigned m1, unsigned m2, int x, int y, int z, unsigned *hindex );
/* Particles of ONE species in the box[3] cells starting at box_min (one reference patch), cell by cell
 * in the order and with the arithmetic of ParticleCreator::create (src/Particles/ParticleCreator.cpp:
 * 300-338): createPosition (:611-745; position_init 0 regular, 1 random, 2 centered, 3 = positions
 * already in x,y,z, copied from another species), createMomentum (:818-851; momentum_init 0 cold,
 * 1 maxwell-juettner with ParticleCreator::maxwellJuttner :1002-1071 and its two tables), createWeight
 * (:933-939), createCharge (:964-974).  Per cell (row-major, z fastest): nppc, n_real = |density| x
 * cell volume (0 = empty cell), charge, temperature (units of m_e c^2; divided by `mass` here).
 * *rng_state is the patch's xorshift32 state, read and written back, so that the next species
 * continues the stream as the reference does.  All pointers are HOST pointers. */
int  sb200_create_particles_ref( unsigned int *rng_state, int position_init, int momentum_init,
                                 const int box[3], const double box_min[3], const double cell_length[3],
                                 const int *nppc, const double *n_real, const double *charge, const double *temperature,
                                 double mass, const int regular_number[3],
                                 const double *lnInvF, const double *lnInvH,
                                 double *x, double *y, double *z, double *px, double *py, double *pz,
                                 double *w, short *q, size_t capacity, size_t *n_created );

/* ---- debugging -------------------------------------------------------------------------- */
/* HOST int[8] counters, cleared by sb200_sort: [0] particles outside the patch without a tag at
 * sort time, [1] particles found outside the cell their sort key says during sb200_dynamics. */
int  sb200_debug_flags( sb200_patch *p, int flags[8] );
/* number of CUDA kernels this library has launched since it was loaded (HOST output). */
int  sb200_launch_count( unsigned long long *n );

#ifdef __cplusplus
}
#endif
#endif /* SMILEI_B200_H */
