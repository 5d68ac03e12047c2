/* smilei_oracle.h — CPU restatement of Smilei's 3D Cartesian PIC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement of the
 * reference algorithm, function by function, with the reference's own operation order.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (smilei_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every operator here is checked bit-for-bit against the
 * reference's own translation units compiled from /root/reference (oracle/_ref,
 * built by oracle/ref_build/build_ref.sh) in tests/test_oracle_vs_ref.py, and the
 * outputs of that reference build are committed as fixtures under tests/golden/.
 *
 * All paths below are relative to /root/reference/src.
 *
 * Array conventions (identical to the reference):
 *   - a field component is one contiguous double[nx*ny*nz], index (i*ny+j)*nz+k
 *     (Field/Field3D.cpp:177-216);  p[d] = n[d]+2*o[d]+1 (primal), d = p+1 (dual)
 *     (ElectroMagn/ElectroMagn.cpp:44-49).
 *     Ex(d,p,p) Ey(p,d,p) Ez(p,p,d) Bx(p,d,d) By(d,p,d) Bz(d,d,p), J like E, rho(p,p,p).
 *   - particles are SoA: x,y,z,px,py,pz,w (double), q (short), key (int)
 *     (Particles/Particles.h:526-566).
 *   - scratch is component-major with stride N: Epart[c*N+p], Bpart, iold, deltaold
 *     (SmileiMPI/SmileiMPI.h:213-221,272-279).
 */
#ifndef SMILEI_ORACLE_H
#define SMILEI_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int    n[3];          /* cells of this patch per dim (params.patch_size_)            */
    int    o[3];          /* oversize = interpolation order (Params/Params.cpp:1202-1207) */
    double cell[3];       /* cell_length                                                 */
    double dt;            /* timestep                                                    */
    int    pcoord[3];     /* patch coordinates in the patch grid (Patch::Pcoordinates)    */
    int    npatch[3];     /* number_of_patches                                           */
    int    n_moved;       /* cells the moving window has advanced along x (SimWindow::n_moved)  */
} orc_grid;

/* geometry helpers (Patch/Patch.cpp:136-165) */
void orc_patch_bounds( const orc_grid *g, double *min_local, double *max_local, int *cell_start_gc );
void orc_dims( const orc_grid *g, int *p, int *d );
long orc_field_size( const orc_grid *g, int field_id ); /* 0..2 E/J like, 3..5 B like, 6 rho */

/* a3-a6 gather */
void orc_interp( const orc_grid *g, int order,
                 const double *Ex, const double *Ey, const double *Ez,
                 const double *Bxm, const double *Bym, const double *Bzm,
                 const double *x, const double *y, const double *z, int nparts, int istart, int iend,
                 double *Epart, double *Bpart, int *iold, double *deltaold );

/* a7-a9 push: pusher 0 boris, 1 vay, 2 higueracary */
void orc_push( const orc_grid *g, int pusher, double mass,
               double *x, double *y, double *z, double *px, double *py, double *pz,
               const short *q, int nparts, int istart, int iend,
               const double *Epart, const double *Bpart, double *invgf );

/* a10 boundary tagging (periodic / inter-patch) */
void orc_bc_tag( const orc_grid *g, const double *x, const double *y, const double *z,
                 int *keys, int istart, int iend );

/* a10b particle boundary conditions with `remove` at global box sides */
void orc_bc_apply( const orc_grid *g, const int *bc_remove, const double *x, const double *y, const double *z,
                   const double *px, const double *py, const double *pz, const double *w, short *q,
                   int *cell_keys, int imin, int imax, double *energy_lost );

/* Silver-Mueller boundary condition on one global box face (zero external fields) */
void orc_apply_SM( const orc_grid *g, int i_boundary, const double *K, const int *is_boundary,
                   const double *Ex, const double *Ey, const double *Ez, double *Bx, double *By, double *Bz,
                   const double *db1, const double *db2 );


/* a11-a14 deposit */
void orc_project( const orc_grid *g, int order, double *Jx, double *Jy, double *Jz,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold );
/* a12 diag-step deposit with rho (order 2 only) */
void orc_project_rho_o2( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold );
void orc_project_rho_o4( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold );
/* currentsAndDensityWrapper with diag_flag (either order); the arrays are the totals or the species' own */
void orc_project_rho( const orc_grid *g, int order, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold );
/* ElectroMagn3D::computeTotalRhoJ for one species (null arrays are skipped) */
void orc_compute_total_rhoJ( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                             const double *Jx_s, const double *Jy_s, const double *Jz_s, const double *rho_s );

/* a16-a19 Maxwell */
void orc_save_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                 double *Bxm, double *Bym, double *Bzm );
void orc_maxwell_ampere( const orc_grid *g, double *Ex, double *Ey, double *Ez,
                         const double *Bx, const double *By, const double *Bz,
                         const double *Jx, const double *Jy, const double *Jz );
void orc_maxwell_faraday( const orc_grid *g, const double *Ex, const double *Ey, const double *Ez,
                          double *Bx, double *By, double *Bz );
void orc_center_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                   double *Bxm, double *Bym, double *Bzm );

/* a20-a21 keys + sort */
void orc_cell_keys( const orc_grid *g, const double *x, const double *y, const double *z,
                    int *keys, int *count, int istart, int iend );
/* stable counting sort on keys>=0; perm[new]=old; returns number kept; first[ncells+1] */
int  orc_counting_sort_perm( const int *keys, int nparts, int ncells, int *first, int *perm );

/* a23 energies */
double orc_ukin( double mass, const double *px, const double *py, const double *pz, const double *w, int nparts );
double orc_field_norm2( const orc_grid *g, const double *f, int dualx, int dualy, int dualz );
double orc_uelm( const orc_grid *g, const double *Ex, const double *Ey, const double *Ez,
                 const double *Bxm, const double *Bym, const double *Bzm );

/* halo semantics between two patches adjacent along `dim`, L on the min side of R
 * (Patch/SyncVectorPatch.cpp:263-311 sum, :1483-1527 exchange).  L==R is the
 * single-patch periodic case. */
void orc_sum_pair( const orc_grid *g, int dim, int dualx, int dualy, int dualz, double *L, double *R );
void orc_exchange_pair( const orc_grid *g, int dim, int dualx, int dualy, int dualz, double *L, double *R );

#ifdef __cplusplus
}
#endif
#endif
