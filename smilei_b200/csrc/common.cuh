// common.cuh — shared host/device definitions of the B200 PIC hot path.
//
// Device data layout:
//  * every field component lives in its own HBM array of identical padded shape
//    (AX, AY, AZ) = (d[0], d[1], roundup(d[2],8)), element (i,j,k) at (i*AY+j)*AZ+k, where
//    d = n+2*oversize+2 is the reference's dual dimension.  Only the sub-box
//    [0,dims_c) of component c is meaningful; the padding is kept at zero.  A common
//    shape means one index expression serves all 13 arrays and rows start 128-B aligned.
//  * particles are SoA columns (x y z px py pz w : double, q : short, key : int), kept
//    sorted by primal-node cell key between steps; first[c] .. first[c+1] is the run of
//    cell c (row-major cells, (n0+1)(n1+1)(n2+1) of them).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>
#include "../../include/smilei_b200.h"

namespace sb200 {

struct GridDev {
    int    n[3];        // cells
    int    o[3];        // oversize
    int    order;
    int    p[3], d[3];  // primal / dual dims
    int    ncell[3];    // n+1 node-centred cells per dim
    int    begin[3];    // Patch::cell_starting_global_index = pcoord*n - o
    int    ax, ay, az;  // allocated (padded) shape
    long long sx, sy;   // strides in elements
    double cell[3];
    double dxi[3];      // 1/cell_length
    double d_ov_dt[3];  // cell_length/dt
    double dt, dts2;
    double dt_ov_d[3];
    double xmin[3], xmax[3];     // Patch::min_local_/max_local_
    double min_loc_round[3];     // round(min_local*dxi), SpeciesV.cpp:800-802
    double inv_cell_volume, cell_volume;
    int    pcoord[3], npatch[3];
};

// dual flags per field id (ElectroMagn3D.cpp:115-123)
__host__ __device__ inline int field_dual( int id, int dim )
{
    // a species' own Jx_s Jy_s Jz_s rho_s (SB200_SPECIES_FIELD) are shaped like Jx Jy Jz rho
    if( id >= SB200_NFIELDS ) id = SB200_JX + ( id - SB200_NFIELDS ) % 4;
    // E/J: dual along own direction; B/Bm: primal along own direction, dual elsewhere
    switch( id ) {
        case SB200_EX: case SB200_JX: return dim==0;
        case SB200_EY: case SB200_JY: return dim==1;
        case SB200_EZ: case SB200_JZ: return dim==2;
        case SB200_BX: case SB200_BXM: return dim!=0;
        case SB200_BY: case SB200_BYM: return dim!=1;
        case SB200_BZ: case SB200_BZM: return dim!=2;
        default: return 0;
    }
}

struct SpeciesDev {
    double mass = 1.;
    int    pusher = 0;
    size_t cap = 0, n = 0;
    double *col[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x y z px py pz w
    short  *q = nullptr;
    int    *key = nullptr;
    int    *first = nullptr;       // ncells+1, valid after sort
    int    *count = nullptr;       // ncells+1: histogram of the NEW keys, filled by the dynamics kernel / arrivals
    bool   count_valid = false;
    bool   sorted = false;
    int    maxcount = 0;       // most particles in one cell (after the last sort)
    double qwmax = 0.;         // max |charge*weight| seen in this species
    int    *leave_idx = nullptr;   // [6][leave_cap] indices of the particles tagged -2..-7, in the order the atomics served them
    size_t leave_cap = 0;
    unsigned long long *d_qwmax = nullptr;   // device copy kept up to date by imports / arrivals (bits of a positive double)
    // Deferred gather: sb200_sort leaves the columns where they are and only produces the sorted order
    // perm[j] = index of the particle that belongs in slot j.  The next dynamics kernel reads its particles
    // through perm and writes them, pushed, at their sorted slots of the spare column set (one pass over the
    // particle data per step instead of two); any other consumer calls materialize() first.
    size_t n_sorted = 0;                   // particles [0, n_sorted) are the cell runs `first` describes (0: no valid runs); later ones were appended
    bool   window_tagged = false;          // leavers were tagged by a window shift: they sit anywhere in the first cells along x
    int    bc[6] = { 0, 0, 0, 0, 0, 0 };   // SB200_PBC_* per global box side
    double *d_lost = nullptr;              // sum of w*(gamma-1) of removed particles
    int    *perm = nullptr;
    size_t perm_cap = 0;
    bool   perm_pending = false;
    // the species' own Jx_s Jy_s Jz_s rho_s (ElectroMagn::Jx_s .. rho_s): allocated on request, the target of the
    // deposit on diag steps (Projector3D2Order.cpp:756-759), nullptr = deposit into the totals
    double *fs[4] = { nullptr, nullptr, nullptr, nullptr };
};

struct ParticleBuf {               // spare SoA set the sort scatters into, then swaps with the species
    size_t cap = 0;
    double *col[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    short  *q = nullptr;
    int    *key = nullptr;
};

void set_error( const std::string &s );
extern unsigned long long g_launches;   // kernels launched by this library (bench.py: gpu_launches)

#define SB200_CUDA( call ) do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { \
        sb200::set_error( std::string( #call ) + ": " + cudaGetErrorString( e_ ) + " (" + __FILE__ + ":" + std::to_string( __LINE__ ) + ")" ); \
        return 1; } } while( 0 )
#define SB200_CHECK( cond, msg ) do { if( !( cond ) ) { sb200::set_error( std::string( msg ) ); return 1; } } while( 0 )

} // namespace sb200

struct sb200_patch {
    sb200_grid        grid;
    sb200::GridDev    gd;
    int               device = 0;
    cudaStream_t      stream = nullptr;
    int               nspec = 0;
    int               n_moved = 0;           // cells the moving window has advanced (SimWindow::n_moved)
    size_t            falloc = 0;            // elements per field array
    double           *f[SB200_NFIELDS] = {};
    sb200::SpeciesDev *sp = nullptr;
    sb200::ParticleBuf spare;
    size_t            ncells = 0;
    // sort workspace
    int              *cursor = nullptr;      // ncells
    int              *perm = nullptr;        // capacity
    size_t            perm_cap = 0;
    int              *blocksums = nullptr;   size_t blocksums_cap = 0;
    // staging / reductions
    double           *stage = nullptr;       size_t stage_cap = 0;   // device, elements
    double           *red = nullptr;         // device partial sums
    int              *leave_counts = nullptr;// device int[8]
    int              *iflags = nullptr;      // device int[8] : error / overflow flags
    int              *d_maxcount = nullptr;  // device int: max particles per cell of the species being sorted
    // optional scratch (SB200_DYN_KEEP_SCRATCH)
    double           *sc_E = nullptr, *sc_B = nullptr, *sc_invgf = nullptr, *sc_delta = nullptr;
    int              *sc_iold = nullptr;     size_t sc_cap = 0;
};

namespace sb200 {
// device array of a field id: one of the 13 patch fields or a species' own array (SB200_SPECIES_FIELD); nullptr if
// the id is out of range or the species array was not requested
inline double *field_ptr( sb200_patch *p, int id )
{
    if( !p || id < 0 ) return nullptr;
    if( id < SB200_NFIELDS ) return p->f[id];
    const int is = ( id - SB200_NFIELDS )/4;
    return is < p->nspec ? p->sp[is].fs[( id - SB200_NFIELDS ) % 4] : nullptr;
}
// implemented across the .cu files; all enqueue on p->stream
int launch_maxwell( sb200_patch *p );
int launch_center_shell( sb200_patch *p );
int launch_dynamics( sb200_patch *p, int ispec, int flags );
int launch_sort( sb200_patch *p, int ispec );
int launch_rho( sb200_patch *p, int ispec );
int launch_energy( sb200_patch *p, double *ukin, double *uelm );
int ensure_spare( sb200_patch *p, size_t cap );
int grow_species( sb200_patch *p, int ispec, size_t need );        // room for `need` particles (larger arrays, device copy)
int ensure_perm( sb200_patch *p, size_t cap );
void swap_with_spare( sb200_patch *p, SpeciesDev &s );
int materialize( sb200_patch *p, int ispec );                    // apply a pending sort permutation to the columns (k_gather + swap)
int ensure_stage( sb200_patch *p, size_t elems );
int exclusive_scan_int( sb200_patch *p, int *data, size_t n );   // in place, device
int update_qwmax( sb200_patch *p, int ispec, size_t first, size_t n );   // fold |q*w| of particles [first, first+n) into d_qwmax
}
