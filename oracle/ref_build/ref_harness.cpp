/* ref_harness.cpp — C entry points that drive the REFERENCE's own operator classes.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled, together with the reference's
 * hot-path translation units taken from where they lie under /root/reference/src,
 * into oracle/_ref/libsmilei_ref.so by oracle/ref_build/build_ref.sh.  It contains
 * no reference code: it builds the minimal object graph the reference operators read
 * (Params, Patch, Species, SmileiMPI, ElectroMagn, Field3D, Particles) and calls
 *   Interpolator3D2Order / Interpolator3D4Order ::fieldsWrapper
 *   PusherBoris / PusherVay / PusherHigueraCary ::operator()
 *   Projector3D2Order / Projector3D4Order ::currentsAndDensityWrapper
 *   MA_Solver3D_norm / MF_Solver3D_Yee ::operator()
 *   ElectroMagn3D::saveMagneticFields / centerMagneticFields
 *   SpeciesV::computeParticleCellKeys, internal_inf / internal_sup, remove_particle_inf / _sup, Field3D::norm2
 *   ElectroMagnBC3D_SM (constructor + apply, with array-backed LaserProfile objects)
 * exactly as Species::dynamics / VectorPatch::solveMaxwell do.
 *
 * The big driver classes (Params, Patch, Species, SmileiMPI, ElectroMagn3D) cannot be
 * constructed without MPI, HDF5 and an embedded Python namelist, so they are
 * materialised as zero-initialised storage whose few members the operators read are
 * then set explicitly (the same values Params.cpp / Patch.cpp would compute).
 * Particles and Field3D are real, constructor-built reference objects.
 */
/* every standard header the reference headers pull in, included before the access
 * override below so that the override only affects the reference's own classes */
#include <Python.h>
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <ostream>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <utility>
#include <vector>
#include <omp.h>
#include <unistd.h>
#include <sys/time.h>
#define private public
#define protected public
#include "Params.h"
#include "Patch.h"
#include "Species.h"
#include "SpeciesV.h"
#include "SmileiMPI.h"
#include "ElectroMagn.h"
#include "ElectroMagn3D.h"
#include "Field3D.h"
#include "Particles.h"
#include "Interpolator3D2Order.h"
#include "Interpolator3D4Order.h"
#include "PusherBoris.h"
#include "PusherVay.h"
#include "PusherHigueraCary.h"
#include "Projector3D2Order.h"
#include "Projector3D4Order.h"
#include "Interpolator3D2OrderV.h"
#include "Projector3D2OrderV.h"
#include "MA_Solver3D_norm.h"
#include "MF_Solver3D_Yee.h"
#include "BoundaryConditionType.h"
#include "ElectroMagnBC3D_SM.h"
#include "Laser.h"
#undef private
#undef protected

#include <cstdlib>
#include <cstring>
#include <new>
#include <omp.h>

extern "C" {

typedef struct {
    int    n[3];
    int    o[3];
    double cell[3];
    double dt;
    int    pcoord[3];
    int    npatch[3];
    int    n_moved;
} orc_grid;   /* same POD as oracle/smilei_oracle.h */

}

namespace {

template<class T> T *blank()
{
    void *m = std::calloc( 1, sizeof( T ) + 64 );
    return reinterpret_cast<T *>( m );
}

struct Ctx {
    Params      *params;
    Patch       *patch;
    SpeciesV    *species;
    SmileiMPI   *smpi;
    ElectroMagn3D *em;
    orc_grid     g;
};

Ctx *make_ctx( const orc_grid *g, double mass, int nthreads )
{
    Ctx *c = new Ctx;
    c->g = *g;
    c->params = blank<Params>();
    Params &P = *c->params;
    new( &P.cell_length ) std::vector<double>( g->cell, g->cell+3 );
    new( &P.patch_size_ ) std::vector<unsigned int>( g->n, g->n+3 );
    new( &P.oversize ) std::vector<unsigned int>( g->o, g->o+3 );
    new( &P.number_of_patches ) std::vector<unsigned int>( g->npatch, g->npatch+3 );
    new( &P.vectorization_mode ) std::string( "off" );
    new( &P.geometry ) std::string( "3Dcartesian" );
    P.timestep = g->dt;
    P.nDim_particle = 3;
    P.nDim_field = 3;
    P.cell_sorting_ = false;
    P.is_pxr = false;
    P.Friedman_filter = false;
    /* Params::compute, Params.cpp:1172-1186 */
    P.cell_volume = 1.0;
    for( int i=0; i<3; i++ ) P.cell_volume *= P.cell_length[i];

    c->patch = blank<Patch>();
    Patch &pt = *c->patch;
    new( &pt.cell_starting_global_index ) std::vector<int>( 3 );
    new( &pt.min_local_ ) std::vector<double>( 3 );
    new( &pt.max_local_ ) std::vector<double>( 3 );
    new( &pt.Pcoordinates ) std::vector<unsigned int>( g->pcoord, g->pcoord+3 );
    /* Patch::initStep3, Patch.cpp:146-153 */
    for( int i=0; i<3; i++ ) {
        pt.min_local_[i] = ( pt.Pcoordinates[i]   )*( P.patch_size_[i]*P.cell_length[i] );
        pt.max_local_[i] = ( pt.Pcoordinates[i]+1 )*( P.patch_size_[i]*P.cell_length[i] );
        pt.cell_starting_global_index[i]  = pt.Pcoordinates[i]*P.patch_size_[i];
        pt.cell_starting_global_index[i] -= P.oversize[i];
    }
    /* Patch::initStep3 with n_moved (moving window), Patch.cpp:159-163 */
    pt.cell_starting_global_index[0] += g->n_moved;
    pt.min_local_[0] += g->n_moved*P.cell_length[0];
    pt.max_local_[0] += g->n_moved*P.cell_length[0];

    c->species = blank<SpeciesV>();
    SpeciesV &S = *c->species;
    new( &S.min_loc_vec ) std::vector<double>( pt.min_local_ );
    S.mass_ = mass;
    S.nDim_field = 3;
    for( int i=0; i<3; i++ ) {
        S.dx_inv_[i] = 1./P.cell_length[i];              /* Species.cpp: dx_inv_ */
        S.length_[i] = P.patch_size_[i]+1;               /* SpeciesV.cpp:72-77   */
    }
    S.particles = new Particles();

    c->smpi = blank<SmileiMPI>();
    SmileiMPI &M = *c->smpi;
    new( &M.dynamics_Epart ) std::vector<std::vector<double>>( nthreads );
    new( &M.dynamics_Bpart ) std::vector<std::vector<double>>( nthreads );
    new( &M.dynamics_invgf ) std::vector<std::vector<double>>( nthreads );
    new( &M.dynamics_iold ) std::vector<std::vector<int>>( nthreads );
    new( &M.dynamics_deltaold ) std::vector<std::vector<double>>( nthreads );
    M.use_BTIS3 = false;

    c->em = blank<ElectroMagn3D>();
    ElectroMagn3D &E = *c->em;
    new( &E.dimPrim ) std::vector<unsigned int>( 3 );
    new( &E.dimDual ) std::vector<unsigned int>( 3 );
    new( &E.Jx_s ) std::vector<Field *>( 1, ( Field * )NULL );
    new( &E.Jy_s ) std::vector<Field *>( 1, ( Field * )NULL );
    new( &E.Jz_s ) std::vector<Field *>( 1, ( Field * )NULL );
    new( &E.rho_s ) std::vector<Field *>( 1, ( Field * )NULL );
    /* ElectroMagn::ElectroMagn, ElectroMagn.cpp:44-49 */
    for( int i=0; i<3; i++ ) {
        E.dimPrim[i] = g->n[i] + 2*g->o[i] + 1;
        E.dimDual[i] = g->n[i] + 2*g->o[i] + 2;
    }
    std::vector<unsigned int> dp = E.dimPrim;
    /* ElectroMagn3D::initElectroMagn3DQuantities, ElectroMagn3D.cpp:115-123,156-159 */
    E.Ex_  = new Field3D( dp, 0, false );
    E.Ey_  = new Field3D( dp, 1, false );
    E.Ez_  = new Field3D( dp, 2, false );
    E.Bx_  = new Field3D( dp, 0, true );
    E.By_  = new Field3D( dp, 1, true );
    E.Bz_  = new Field3D( dp, 2, true );
    E.Bx_m = new Field3D( dp, 0, true );
    E.By_m = new Field3D( dp, 1, true );
    E.Bz_m = new Field3D( dp, 2, true );
    E.Jx_  = new Field3D( dp, 0, false );
    E.Jy_  = new Field3D( dp, 1, false );
    E.Jz_  = new Field3D( dp, 2, false );
    E.rho_ = new Field3D( dp );
    /* ElectroMagn3D.cpp:190-229 */
    for( unsigned int i=0 ; i<3 ; i++ ) {
        for( unsigned int isDual=0 ; isDual<2 ; isDual++ ) {
            E.istart[i][isDual] = g->o[i] + ( g->pcoord[i]!=0 ? 1 : 0 );
            int b = g->n[i] + 1 + isDual;
            if( g->npatch[i]!=1 ) {
                if( ( !isDual ) && ( g->pcoord[i]!=0 ) ) {
                    b--;
                } else if( isDual ) {
                    b--;
                    if( ( g->pcoord[i]!=0 ) && ( g->pcoord[i]!=g->npatch[i]-1 ) ) b--;
                }
            }
            E.bufsize[i][isDual] = b;
        }
    }
    return c;
}

void free_ctx( Ctx *c )
{
    ElectroMagn3D &E = *c->em;
    Field *f[13] = { E.Ex_, E.Ey_, E.Ez_, E.Bx_, E.By_, E.Bz_, E.Bx_m, E.By_m, E.Bz_m, E.Jx_, E.Jy_, E.Jz_, E.rho_ };
    for( int i=0; i<13; i++ ) delete f[i];
    delete c->species->particles;
    /* blank<> storage is intentionally leaked member-wise: test process only. */
    std::free( c->em );
    std::free( c->smpi );
    std::free( c->species );
    std::free( c->patch );
    std::free( c->params );
    delete c;
}

void load( Field *f, const double *src )
{
    if( src ) std::memcpy( f->data_, src, sizeof( double )*f->number_of_points_ );
}
void store( Field *f, double *dst )
{
    if( dst ) std::memcpy( dst, f->data_, sizeof( double )*f->number_of_points_ );
}

void set_particles( Ctx *c, const double *x, const double *y, const double *z,
                    const double *px, const double *py, const double *pz,
                    const double *w, const short *q, int n )
{
    Particles &p = *c->species->particles;
    p.initialize( n, 3, false );
    const double *pos[3] = { x, y, z }, *mom[3] = { px, py, pz };
    for( int d=0; d<3; d++ ) {
        if( pos[d] ) std::memcpy( p.Position[d].data(), pos[d], sizeof( double )*n );
        if( mom[d] ) std::memcpy( p.Momentum[d].data(), mom[d], sizeof( double )*n );
    }
    if( w ) std::memcpy( p.Weight.data(), w, sizeof( double )*n );
    if( q ) std::memcpy( p.Charge.data(), q, sizeof( short )*n );
    p.first_index.assign( 1, 0 );
    p.last_index.assign( 1, n );
    p.cell_keys.assign( n, 0 );
}

void resize_scratch( Ctx *c, int n )
{
    /* SmileiMPI::resizeBuffers, SmileiMPI.h:272-279 */
    SmileiMPI &M = *c->smpi;
    M.dynamics_Epart[0].resize( 3*n );
    M.dynamics_Bpart[0].resize( 3*n );
    M.dynamics_invgf[0].resize( n );
    M.dynamics_iold[0].resize( 3*n );
    M.dynamics_deltaold[0].resize( 3*n );
}

} // namespace

extern "C" {

/* target of the generated stubs for symbols the hot path never calls (build_ref.sh) */
void sb200_ref_trap_report( const char *name )
{
    std::fprintf( stderr, "libsmilei_ref: call into un-built reference symbol %s\n", name );
    std::abort();
}

void ref_interp( const orc_grid *g, int order,
                 const double *Ex, const double *Ey, const double *Ez,
                 const double *Bxm, const double *Bym, const double *Bzm,
                 const double *x, const double *y, const double *z, int nparts, int istart, int iend,
                 double *Epart, double *Bpart, int *iold, double *deltaold )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Ex_, Ex ); load( c->em->Ey_, Ey ); load( c->em->Ez_, Ez );
    load( c->em->Bx_m, Bxm ); load( c->em->By_m, Bym ); load( c->em->Bz_m, Bzm );
    set_particles( c, x, y, z, 0, 0, 0, 0, 0, nparts );
    resize_scratch( c, nparts );
    Interpolator *I = order==2 ? ( Interpolator * )new Interpolator3D2Order( *c->params, c->patch )
                      : ( Interpolator * )new Interpolator3D4Order( *c->params, c->patch );
    I->fieldsWrapper( c->em, *c->species->particles, c->smpi, &istart, &iend, 0 );
    std::memcpy( Epart, c->smpi->dynamics_Epart[0].data(), sizeof( double )*3*nparts );
    std::memcpy( Bpart, c->smpi->dynamics_Bpart[0].data(), sizeof( double )*3*nparts );
    std::memcpy( iold, c->smpi->dynamics_iold[0].data(), sizeof( int )*3*nparts );
    std::memcpy( deltaold, c->smpi->dynamics_deltaold[0].data(), sizeof( double )*3*nparts );
    delete I;
    free_ctx( c );
}

void ref_push( const orc_grid *g, int pusher, double mass,
               double *x, double *y, double *z, double *px, double *py, double *pz,
               const short *q, int nparts, int istart, int iend,
               const double *Epart, const double *Bpart, double *invgf )
{
    Ctx *c = make_ctx( g, mass, 1 );
    set_particles( c, x, y, z, px, py, pz, 0, q, nparts );
    resize_scratch( c, nparts );
    std::memcpy( c->smpi->dynamics_Epart[0].data(), Epart, sizeof( double )*3*nparts );
    std::memcpy( c->smpi->dynamics_Bpart[0].data(), Bpart, sizeof( double )*3*nparts );
    Pusher *P = pusher==0 ? ( Pusher * )new PusherBoris( *c->params, c->species )
                : pusher==1 ? ( Pusher * )new PusherVay( *c->params, c->species )
                : ( Pusher * )new PusherHigueraCary( *c->params, c->species );
    ( *P )( *c->species->particles, c->smpi, istart, iend, 0 );
    Particles &p = *c->species->particles;
    double *pos[3] = { x, y, z }, *mom[3] = { px, py, pz };
    for( int d=0; d<3; d++ ) {
        std::memcpy( pos[d], p.Position[d].data(), sizeof( double )*nparts );
        std::memcpy( mom[d], p.Momentum[d].data(), sizeof( double )*nparts );
    }
    std::memcpy( invgf, c->smpi->dynamics_invgf[0].data(), sizeof( double )*nparts );
    delete P;
    free_ctx( c );
}

void ref_bc_tag( const orc_grid *g, const double *x, const double *y, const double *z,
                 int *keys, int imin, int imax )
{
    Ctx *c = make_ctx( g, 1., 1 );
    set_particles( c, x, y, z, 0, 0, 0, 0, 0, imax );
    Particles &p = *c->species->particles;
    std::vector<double> invgf;
    double e = 0.;
    /* PartBoundCond::apply, PartBoundCond.h:38-76 (periodic => patch bounds, PartBoundCond.cpp:44-69) */
    int *ck = p.getPtrCellKeys();
    for( int i=imin; i<imax; i++ ) ck[i] = 0;
    for( int d=0; d<3; d++ ) {
        internal_inf( c->species, imin, imax, d, c->patch->min_local_[d], g->dt, invgf, NULL, e );
        internal_sup( c->species, imin, imax, d, c->patch->max_local_[d], g->dt, invgf, NULL, e );
    }
    std::memcpy( keys, ck, sizeof( int )*imax );
    free_ctx( c );
}

void ref_project( const orc_grid *g, int order, double *Jx, double *Jy, double *Jz,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Jx_, Jx ); load( c->em->Jy_, Jy ); load( c->em->Jz_, Jz );
    set_particles( c, x, y, z, 0, 0, 0, w, q, nparts );
    resize_scratch( c, nparts );
    std::memcpy( c->smpi->dynamics_iold[0].data(), iold, sizeof( int )*3*nparts );
    std::memcpy( c->smpi->dynamics_deltaold[0].data(), deltaold, sizeof( double )*3*nparts );
    Projector *P = order==2 ? ( Projector * )new Projector3D2Order( *c->params, c->patch )
                   : ( Projector * )new Projector3D4Order( *c->params, c->patch );
    P->currentsAndDensityWrapper( c->em, *c->species->particles, c->smpi, istart, iend, 0, false, false, 0 );
    store( c->em->Jx_, Jx ); store( c->em->Jy_, Jy ); store( c->em->Jz_, Jz );
    delete P;
    free_ctx( c );
}

void ref_project_rho_o2( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Jx_, Jx ); load( c->em->Jy_, Jy ); load( c->em->Jz_, Jz ); load( c->em->rho_, rho );
    set_particles( c, x, y, z, 0, 0, 0, w, q, nparts );
    resize_scratch( c, nparts );
    std::memcpy( c->smpi->dynamics_iold[0].data(), iold, sizeof( int )*3*nparts );
    std::memcpy( c->smpi->dynamics_deltaold[0].data(), deltaold, sizeof( double )*3*nparts );
    Projector *P = new Projector3D2Order( *c->params, c->patch );
    /* diag_flag = true with no per-species arrays: projects into the totals with rho */
    P->currentsAndDensityWrapper( c->em, *c->species->particles, c->smpi, istart, iend, 0, true, false, 0 );
    store( c->em->Jx_, Jx ); store( c->em->Jy_, Jy ); store( c->em->Jz_, Jz ); store( c->em->rho_, rho );
    delete P;
    free_ctx( c );
}

/* currentsAndDensityWrapper with diag_flag = true, either order; with species_arrays != 0 the species owns
 * Jx_s .. rho_s (Projector3D2Order.cpp:756-759, Projector3D4Order.cpp:706-709): the deposit goes there, the
 * totals (Jx .. rho as handed in) are untouched until ElectroMagn3D::computeTotalRhoJ (ElectroMagn3D.cpp:1753)
 * adds the species arrays into them.  Outputs: totals in Jx..rho, species arrays in sJx..srho. */
void ref_project_rho_species( const orc_grid *g, int order, int species_arrays,
                  double *Jx, double *Jy, double *Jz, double *rho,
                  double *sJx, double *sJy, double *sJz, double *srho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold, int compute_total )
{
    Ctx *c = make_ctx( g, 1., 1 );
    ElectroMagn3D &E = *c->em;
    load( E.Jx_, Jx ); load( E.Jy_, Jy ); load( E.Jz_, Jz ); load( E.rho_, rho );
    if( species_arrays ) {
        std::vector<unsigned int> dp = E.dimPrim;
        E.Jx_s[0] = new Field3D( dp, 0, false ); E.Jy_s[0] = new Field3D( dp, 1, false );
        E.Jz_s[0] = new Field3D( dp, 2, false ); E.rho_s[0] = new Field3D( dp );
        load( E.Jx_s[0], sJx ); load( E.Jy_s[0], sJy ); load( E.Jz_s[0], sJz ); load( E.rho_s[0], srho );
    }
    *const_cast<unsigned int *>( &E.n_species ) = 1;
    set_particles( c, x, y, z, 0, 0, 0, w, q, nparts );
    resize_scratch( c, nparts );
    std::memcpy( c->smpi->dynamics_iold[0].data(), iold, sizeof( int )*3*nparts );
    std::memcpy( c->smpi->dynamics_deltaold[0].data(), deltaold, sizeof( double )*3*nparts );
    Projector *P = order == 2 ? ( Projector * )new Projector3D2Order( *c->params, c->patch )
                              : ( Projector * )new Projector3D4Order( *c->params, c->patch );
    P->currentsAndDensityWrapper( c->em, *c->species->particles, c->smpi, istart, iend, 0, true, false, 0 );
    if( compute_total ) E.ElectroMagn3D::computeTotalRhoJ();
    store( E.Jx_, Jx ); store( E.Jy_, Jy ); store( E.Jz_, Jz ); store( E.rho_, rho );
    if( species_arrays ) {
        store( E.Jx_s[0], sJx ); store( E.Jy_s[0], sJy ); store( E.Jz_s[0], sJz ); store( E.rho_s[0], srho );
        delete E.Jx_s[0]; delete E.Jy_s[0]; delete E.Jz_s[0]; delete E.rho_s[0];
        E.Jx_s[0] = E.Jy_s[0] = E.Jz_s[0] = E.rho_s[0] = NULL;
    }
    delete P;
    free_ctx( c );
}

void ref_save_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                 double *Bxm, double *Bym, double *Bzm )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Bx_, Bx ); load( c->em->By_, By ); load( c->em->Bz_, Bz );
    c->em->ElectroMagn3D::saveMagneticFields( false );
    store( c->em->Bx_m, Bxm ); store( c->em->By_m, Bym ); store( c->em->Bz_m, Bzm );
    free_ctx( c );
}

void ref_maxwell_ampere( const orc_grid *g, double *Ex, double *Ey, double *Ez,
                         const double *Bx, const double *By, const double *Bz,
                         const double *Jx, const double *Jy, const double *Jz )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Ex_, Ex ); load( c->em->Ey_, Ey ); load( c->em->Ez_, Ez );
    load( c->em->Bx_, Bx ); load( c->em->By_, By ); load( c->em->Bz_, Bz );
    load( c->em->Jx_, Jx ); load( c->em->Jy_, Jy ); load( c->em->Jz_, Jz );
    MA_Solver3D_norm S( *c->params );
    S( c->em );
    store( c->em->Ex_, Ex ); store( c->em->Ey_, Ey ); store( c->em->Ez_, Ez );
    free_ctx( c );
}

void ref_maxwell_faraday( const orc_grid *g, const double *Ex, const double *Ey, const double *Ez,
                          double *Bx, double *By, double *Bz )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Ex_, Ex ); load( c->em->Ey_, Ey ); load( c->em->Ez_, Ez );
    load( c->em->Bx_, Bx ); load( c->em->By_, By ); load( c->em->Bz_, Bz );
    MF_Solver3D_Yee S( *c->params );
    S( c->em );
    store( c->em->Bx_, Bx ); store( c->em->By_, By ); store( c->em->Bz_, Bz );
    free_ctx( c );
}

void ref_center_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                   double *Bxm, double *Bym, double *Bzm )
{
    Ctx *c = make_ctx( g, 1., 1 );
    load( c->em->Bx_, Bx ); load( c->em->By_, By ); load( c->em->Bz_, Bz );
    load( c->em->Bx_m, Bxm ); load( c->em->By_m, Bym ); load( c->em->Bz_m, Bzm );
    c->em->ElectroMagn3D::centerMagneticFields();
    store( c->em->Bx_m, Bxm ); store( c->em->By_m, Bym ); store( c->em->Bz_m, Bzm );
    free_ctx( c );
}

void ref_cell_keys( const orc_grid *g, const double *x, const double *y, const double *z,
                    int *keys, int *count, int istart, int iend )
{
    Ctx *c = make_ctx( g, 1., 1 );
    set_particles( c, x, y, z, 0, 0, 0, 0, 0, iend );
    std::vector<int> dummy;
    int *cnt = count;
    if( !cnt ) {
        dummy.assign( ( size_t )( g->n[0]+1 )*( g->n[1]+1 )*( g->n[2]+1 ), 0 );
        cnt = dummy.data();
    }
    c->species->SpeciesV::computeParticleCellKeys( *c->params, c->species->particles, keys, cnt, istart, iend );
    free_ctx( c );
}

extern "C" char _ZTV8SpeciesV[];   /* vtable for SpeciesV (defined by the reference's SpeciesV.cpp) */

/* ------------------------------------------------------------------------------
 * The reference's OWN sort: SpeciesV::computeParticleCellKeys( params ) on the resident particles (those whose
 * cell_keys entry is >= 0; tagged leavers carry a negative key, Species.cpp:757-775) followed by
 * SpeciesV::sortParticles (SpeciesV.cpp:599-762) with the arrivals of the six neighbours in
 * MPI_buffer_.partRecv[dim][side].  Inputs: n resident particles + tags (< 0: leaver), narr[6] arrivals given
 * back to back in a*.  Outputs: the species after the sort (cap entries available), its cell keys and
 * first_index (ncell entries) / last_index of the last cell.  Returns the particle count after the sort.
 * ------------------------------------------------------------------------------ */
int ref_sort( const orc_grid *g, int n,
              const double *x, const double *y, const double *z, const double *px, const double *py, const double *pz,
              const double *w, const short *q, const int *tags,
              const int *narr, const double *ax, const double *ay, const double *az, const double *apx, const double *apy,
              const double *apz, const double *aw, const short *aq,
              int cap, double *ox, double *oy, double *oz, double *opx, double *opy, double *opz, double *ow, short *oq,
              int *okeys, int *first_index )
{
    Ctx *c = make_ctx( g, 1., 1 );
    SpeciesV &S = *c->species;
    Params &P = *c->params;
    P.keep_position_old = false;
    S.nDim_particle = 3;
    set_particles( c, x, y, z, px, py, pz, w, q, n );
    Particles &p = *S.particles;
    for( int i=0; i<n; i++ ) p.cell_keys[i] = tags[i] < 0 ? tags[i] : 0;
    const unsigned int ncell = ( g->n[0]+1 )*( g->n[1]+1 )*( g->n[2]+1 );
    new( &S.count ) std::vector<int>( ncell, 0 );
    p.first_index.assign( ncell, 0 );
    p.last_index.assign( ncell, 0 );
    /* SpeciesV::computeParticleCellKeys( Params & ), SpeciesV.cpp:857-867: keys and counts of the resident particles */
    S.SpeciesV::computeParticleCellKeys( P, &p, &p.cell_keys[0], &S.count[0], 0, n );
    new( &S.MPI_buffer_.partRecv ) std::vector<std::vector<Particles *>>( 3, std::vector<Particles *>( 2, ( Particles * )NULL ) );
    int off = 0;
    for( int d=0; d<3; d++ ) {
        for( int s=0; s<2; s++ ) {
            Particles *b = new Particles();
            const int m = narr[2*d+s];
            b->initialize( m, 3, false );
            const double *src[7] = { ax, ay, az, apx, apy, apz, aw };
            for( int k=0; k<3; k++ ) {
                if( m ) std::memcpy( b->Position[k].data(), src[k]+off, sizeof( double )*m );
                if( m ) std::memcpy( b->Momentum[k].data(), src[3+k]+off, sizeof( double )*m );
            }
            if( m ) std::memcpy( b->Weight.data(), aw+off, sizeof( double )*m );
            if( m ) std::memcpy( b->Charge.data(), aq+off, sizeof( short )*m );
            b->cell_keys.assign( m, 0 );
            S.MPI_buffer_.partRecv[d][s] = b;
            off += m;
        }
    }
    /* sortParticles calls computeParticleCellKeys through the vtable: the calloc'ed SpeciesV has none, so it is
       given the class's own (Itanium ABI: the object's vptr points 2 words into _ZTV8SpeciesV) */
    *reinterpret_cast<void **>( &S ) = reinterpret_cast<void *>( _ZTV8SpeciesV + 2*sizeof( void * ) );
    S.SpeciesV::sortParticles( P );
    const int nout = ( int )p.size();
    if( nout <= cap ) {
        double *dst[3] = { ox, oy, oz }, *dm[3] = { opx, opy, opz };
        for( int k=0; k<3; k++ ) {
            std::memcpy( dst[k], p.Position[k].data(), sizeof( double )*nout );
            std::memcpy( dm[k], p.Momentum[k].data(), sizeof( double )*nout );
        }
        std::memcpy( ow, p.Weight.data(), sizeof( double )*nout );
        std::memcpy( oq, p.Charge.data(), sizeof( short )*nout );
        std::memcpy( okeys, p.cell_keys.data(), sizeof( int )*nout );
        for( unsigned int ic=0; ic<ncell; ic++ ) first_index[ic] = p.first_index[ic];
        first_index[ncell] = p.last_index[ncell-1];
    }
    for( int d=0; d<3; d++ ) for( int s=0; s<2; s++ ) delete S.MPI_buffer_.partRecv[d][s];
    free_ctx( c );
    return nout;
}

double ref_field_norm2( const orc_grid *g, const double *f, int dualx, int dualy, int dualz )
{
    Ctx *c = make_ctx( g, 1., 1 );
    ElectroMagn3D &E = *c->em;
    Field *F = NULL;
    int code = dualx*4+dualy*2+dualz;
    switch( code ) {
        case 4: F = E.Ex_; break;
        case 2: F = E.Ey_; break;
        case 1: F = E.Ez_; break;
        case 3: F = E.Bx_; break;
        case 5: F = E.By_; break;
        case 6: F = E.Bz_; break;
        default: F = E.rho_; break;
    }
    load( F, f );
    double r = F->norm2( E.istart, E.bufsize );
    free_ctx( c );
    return r;
}

/* ------------------------------------------------------------------------------
 * Timed CPU baseline: the reference's gather + push + BC-tag + deposit on many small
 * patches, one OpenMP thread per patch at a time (VectorPatch::dynamicsWithoutTasks,
 * VectorPatch.cpp:4777 `omp for schedule(runtime)` over patches), operators called in
 * the order of Species::dynamics (Species.cpp:591,727,757,782).
 * All patches share the same field content and particle set (synthetic); each thread
 * works on its own private copy so that memory traffic is realistic.
 * Returns seconds for `nsteps` steps over `npatches` patches.
 * ------------------------------------------------------------------------------ */
double ref_time_dynamics( const orc_grid *g, int order, int pusher, double mass,
                          const double *fields6,   /* Ex,Ey,Ez,Bxm,Bym,Bzm concatenated (compact) */
                          const double *x, const double *y, const double *z,
                          const double *px, const double *py, const double *pz,
                          const double *w, const short *q, int nparts,
                          int npatches, int nsteps, int nthreads, double *checksum )
{
    std::vector<Ctx *> ctx( npatches );
    std::vector<Interpolator *> I( npatches );
    std::vector<Pusher *> Pu( npatches );
    std::vector<Projector *> Pr( npatches );
    for( int ip=0; ip<npatches; ip++ ) {
        Ctx *c = ctx[ip] = make_ctx( g, mass, 1 );
        const double *f = fields6;
        Field *dst[6] = { c->em->Ex_, c->em->Ey_, c->em->Ez_, c->em->Bx_m, c->em->By_m, c->em->Bz_m };
        for( int k=0; k<6; k++ ) { load( dst[k], f ); f += dst[k]->number_of_points_; }
        set_particles( c, x, y, z, px, py, pz, w, q, nparts );
        resize_scratch( c, nparts );
        I[ip]  = order==2 ? ( Interpolator * )new Interpolator3D2Order( *c->params, c->patch )
                 : ( Interpolator * )new Interpolator3D4Order( *c->params, c->patch );
        Pu[ip] = pusher==0 ? ( Pusher * )new PusherBoris( *c->params, c->species )
                 : pusher==1 ? ( Pusher * )new PusherVay( *c->params, c->species )
                 : ( Pusher * )new PusherHigueraCary( *c->params, c->species );
        Pr[ip] = order==2 ? ( Projector * )new Projector3D2Order( *c->params, c->patch )
                 : ( Projector * )new Projector3D4Order( *c->params, c->patch );
    }
    std::vector<double> x0( x, x+nparts ), y0( y, y+nparts ), z0( z, z+nparts );
    omp_set_num_threads( nthreads );
    double t0 = omp_get_wtime();
    for( int it=0; it<nsteps; it++ ) {
        #pragma omp parallel for schedule(dynamic,1)
        for( int ip=0; ip<npatches; ip++ ) {
            Ctx *c = ctx[ip];
            Particles &p = *c->species->particles;
            int istart = 0, iend = nparts;
            /* keep particles inside the patch so that iold stays valid: restore positions */
            std::memcpy( p.Position[0].data(), x0.data(), sizeof( double )*nparts );
            std::memcpy( p.Position[1].data(), y0.data(), sizeof( double )*nparts );
            std::memcpy( p.Position[2].data(), z0.data(), sizeof( double )*nparts );
            c->em->Jx_->put_to( 0. ); c->em->Jy_->put_to( 0. ); c->em->Jz_->put_to( 0. );
            I[ip]->fieldsWrapper( c->em, p, c->smpi, &istart, &iend, 0 );
            ( *Pu[ip] )( p, c->smpi, 0, nparts, 0 );
            int *ck = p.getPtrCellKeys();
            std::vector<double> invgf;
            double e = 0.;
            for( int i=0; i<nparts; i++ ) ck[i] = 0;
            for( int d=0; d<3; d++ ) {
                internal_inf( c->species, 0, nparts, d, c->patch->min_local_[d], g->dt, invgf, NULL, e );
                internal_sup( c->species, 0, nparts, d, c->patch->max_local_[d], g->dt, invgf, NULL, e );
            }
            Pr[ip]->currentsAndDensityWrapper( c->em, p, c->smpi, 0, nparts, 0, false, false, 0 );
        }
    }
    double t1 = omp_get_wtime();
    if( checksum ) {
        double s = 0.;
        Field *J = ctx[0]->em->Jx_;
        for( unsigned int i=0; i<J->number_of_points_; i++ ) s += J->data_[i];
        *checksum = s;
    }
    for( int ip=0; ip<npatches; ip++ ) {
        delete I[ip]; delete Pu[ip]; delete Pr[ip];
        free_ctx( ctx[ip] );
    }
    return t1-t0;
}

/* ------------------------------------------------------------------------------
 * Timed CPU baseline, the reference's VECTORISED path (SpeciesV::dynamics, SpeciesV.cpp:131-563, what a production CPU
 * run uses with vectorization_mode = "on"): per patch and per step, in that function's order,
 *     Interpolator3D2OrderV::fieldsWrapper per cell          (SpeciesV.cpp:208-212)
 *     PusherBoris / Vay / HigueraCary over the patch         (:375-378; the pushers carry `omp simd` loops)
 *     [particles that left the periodic patch are wrapped back — what the exchange with the periodic neighbour does]
 *     SpeciesV::computeParticleCellKeys                      (:472-477)
 *     Projector3D2OrderV::currentsAndDensityWrapper per cell (:493-500)
 *     SpeciesV::sortParticles                                (importAndSortParticles, SpeciesV.cpp:599-762)
 * on npatches DISTINCT patches (each its own fields and particles: the sample leaves the caches when npatches x
 * patch is large), one OpenMP thread per patch at a time (VectorPatch.cpp:4777).  The particles handed in must be
 * cell-sorted with first_index (ncell+1 entries).  Returns seconds for nsteps steps.
 * ------------------------------------------------------------------------------ */
double ref_time_dynamics_V( const orc_grid *g, int pusher, double mass,
                            const double *fields6,
                            const double *x, const double *y, const double *z,
                            const double *px, const double *py, const double *pz,
                            const double *w, const short *q, const int *first_index, int nparts,
                            int npatches, int nsteps, int nthreads, int with_sort, double *checksum, double *J_out )
{
    const unsigned int ncell = ( g->n[0]+1 )*( g->n[1]+1 )*( g->n[2]+1 );
    std::vector<Ctx *> ctx( npatches );
    std::vector<Interpolator *> I( npatches );
    std::vector<Pusher *> Pu( npatches );
    std::vector<Projector *> Pr( npatches );
    for( int ip=0; ip<npatches; ip++ ) {
        Ctx *c = ctx[ip] = make_ctx( g, mass, 1 );
        Params &P = *c->params;
        P.keep_position_old = false;
        new( &P.vectorization_mode ) std::string( "on" );
        const double *f = fields6;
        Field *dst[6] = { c->em->Ex_, c->em->Ey_, c->em->Ez_, c->em->Bx_m, c->em->By_m, c->em->Bz_m };
        for( int k=0; k<6; k++ ) { load( dst[k], f ); f += dst[k]->number_of_points_; }
        set_particles( c, x, y, z, px, py, pz, w, q, nparts );
        resize_scratch( c, nparts );
        SpeciesV &S = *c->species;
        S.nDim_particle = 3;
        new( &S.count ) std::vector<int>( ncell, 0 );
        new( &S.MPI_buffer_.partRecv ) std::vector<std::vector<Particles *>>( 3, std::vector<Particles *>( 2, ( Particles * )NULL ) );
        for( int d=0; d<3; d++ ) for( int sd=0; sd<2; sd++ ) { S.MPI_buffer_.partRecv[d][sd] = new Particles(); S.MPI_buffer_.partRecv[d][sd]->initialize( 0, 3, false ); }
        *reinterpret_cast<void **>( &S ) = reinterpret_cast<void *>( _ZTV8SpeciesV + 2*sizeof( void * ) );
        Particles &p = *S.particles;
        p.first_index.assign( first_index, first_index + ncell );
        p.last_index.assign( first_index + 1, first_index + ncell + 1 );
        I[ip]  = new Interpolator3D2OrderV( P, c->patch );
        Pu[ip] = pusher==0 ? ( Pusher * )new PusherBoris( P, c->species )
                 : pusher==1 ? ( Pusher * )new PusherVay( P, c->species )
                 : ( Pusher * )new PusherHigueraCary( P, c->species );
        Pr[ip] = new Projector3D2OrderV( P, c->patch );
    }
    omp_set_num_threads( nthreads );
    double t0 = omp_get_wtime();
    for( int it=0; it<nsteps; it++ ) {
        #pragma omp parallel for schedule(dynamic,1)
        for( int ip=0; ip<npatches; ip++ ) {
            Ctx *c = ctx[ip];
            SpeciesV &S = *c->species;
            Particles &p = *S.particles;
            const int n = ( int )p.size();
            c->em->Jx_->put_to( 0. ); c->em->Jy_->put_to( 0. ); c->em->Jz_->put_to( 0. );
            for( unsigned int scell=0; scell<ncell; scell++ )
                I[ip]->fieldsWrapper( c->em, p, c->smpi, &p.first_index[scell], &p.last_index[scell], 0, scell, 0 );
            ( *Pu[ip] )( p, c->smpi, 0, n, 0, 0 );
            /* PartBoundCond::apply: internal_inf / internal_sup tag the leavers (cell_keys = -1), Species.cpp:757 */
            for( int i=0; i<n; i++ ) p.cell_keys[i] = 0;
            {
                std::vector<double> invgf;
                double e = 0.;
                for( int d=0; d<3; d++ ) {
                    internal_inf( c->species, 0, n, d, c->patch->min_local_[d], g->dt, invgf, NULL, e );
                    internal_sup( c->species, 0, n, d, c->patch->max_local_[d], g->dt, invgf, NULL, e );
                }
            }
            for( unsigned int ic=0; ic<ncell; ic++ ) S.count[ic] = 0;
            S.SpeciesV::computeParticleCellKeys( *c->params, &p, &p.cell_keys[0], &S.count[0], 0, n );
            for( unsigned int scell=0; scell<ncell; scell++ )
                Pr[ip]->currentsAndDensityWrapper( c->em, p, c->smpi, p.first_index[scell], p.last_index[scell], 0, false, false, 0, scell, 0 );
            /* the exchange: the patch spans the periodic sample box, so its leavers come back through the receive
               buffer of the first neighbour, wrapped across the box (SmileiMPI / Patch::exchParticles) */
            if( with_sort ) {
                Particles &rcv = *S.MPI_buffer_.partRecv[0][0];
                rcv.resize( 0, 3, false );
                for( int i=0; i<n; i++ ) {
                    if( p.cell_keys[i] >= 0 ) continue;
                    p.copyParticle( i, rcv );
                    const int k = rcv.size() - 1;
                    for( int d=0; d<3; d++ ) {
                        const double lo = c->patch->min_local_[d], hi = c->patch->max_local_[d];
                        double &X = rcv.Position[d][k];
                        if( X < lo ) X += hi - lo; else if( X >= hi ) X -= hi - lo;
                    }
                }
                rcv.cell_keys.assign( rcv.size(), 0 );
            }
            if( with_sort ) S.SpeciesV::sortParticles( *c->params );
        }
    }
    double t1 = omp_get_wtime();
    if( checksum ) {
        double s = 0.;
        Field *J = ctx[0]->em->Jx_;
        for( unsigned int i=0; i<J->number_of_points_; i++ ) s += J->data_[i];
        *checksum = s;
    }
    if( J_out ) {          /* Jx Jy Jz of patch 0 after the last step, back to back (for checks against the scalar operators) */
        Field *J[3] = { ctx[0]->em->Jx_, ctx[0]->em->Jy_, ctx[0]->em->Jz_ };
        for( int k=0; k<3; k++ ) { store( J[k], J_out ); J_out += J[k]->number_of_points_; }
    }
    for( int ip=0; ip<npatches; ip++ ) {
        for( int d=0; d<3; d++ ) for( int sd=0; sd<2; sd++ ) delete ctx[ip]->species->MPI_buffer_.partRecv[d][sd];
        delete I[ip]; delete Pu[ip]; delete Pr[ip];
        free_ctx( ctx[ip] );
    }
    return t1-t0;
}

/* Timed CPU baseline for the Maxwell solve: saveMagneticFields + MA + MF + centerMagneticFields
 * (VectorPatch::solveMaxwell, VectorPatch.cpp:1013-1023,1170) over npatches patches. */
double ref_time_maxwell( const orc_grid *g, int npatches, int nsteps, int nthreads )
{
    std::vector<Ctx *> ctx( npatches );
    for( int ip=0; ip<npatches; ip++ ) {
        Ctx *c = ctx[ip] = make_ctx( g, 1., 1 );
        Field *all[9] = { c->em->Ex_, c->em->Ey_, c->em->Ez_, c->em->Bx_, c->em->By_, c->em->Bz_, c->em->Jx_, c->em->Jy_, c->em->Jz_ };
        for( int k=0; k<9; k++ )
            for( unsigned int i=0; i<all[k]->number_of_points_; i++ ) all[k]->data_[i] = 1e-3*( ( i*7+k )%13 );
    }
    MA_Solver3D_norm MA( *ctx[0]->params );
    MF_Solver3D_Yee MF( *ctx[0]->params );
    omp_set_num_threads( nthreads );
    double t0 = omp_get_wtime();
    for( int it=0; it<nsteps; it++ ) {
        #pragma omp parallel for schedule(static)
        for( int ip=0; ip<npatches; ip++ ) {
            ctx[ip]->em->ElectroMagn3D::saveMagneticFields( false );
            MA( ctx[ip]->em );
        }
        #pragma omp parallel for schedule(static)
        for( int ip=0; ip<npatches; ip++ ) MF( ctx[ip]->em );
        #pragma omp parallel for schedule(static)
        for( int ip=0; ip<npatches; ip++ ) ctx[ip]->em->ElectroMagn3D::centerMagneticFields();
    }
    double t1 = omp_get_wtime();
    for( int ip=0; ip<npatches; ip++ ) free_ctx( ctx[ip] );
    return t1-t0;
}

/* PartBoundCond::apply (PartBoundCond.h:38-76) with remove_particle_inf/sup on the sides flagged in bc_remove and
 * internal_inf/sup elsewhere; limits = patch bounds (PartBoundCond.cpp:44-69). */
void ref_bc_apply( const orc_grid *g, const int *bc_remove, const double *x, const double *y, const double *z,
                   const double *px, const double *py, const double *pz, const double *w, short *q,
                   int *keys, int imin, int imax, double *energy_lost )
{
    Ctx *c = make_ctx( g, 1., 1 );
    set_particles( c, x, y, z, px, py, pz, w, q, imax );
    Particles &p = *c->species->particles;
    std::vector<double> invgf;
    double energy_tot = 0.;
    int *ck = p.getPtrCellKeys();
    for( int i=imin; i<imax; i++ ) ck[i] = 0;
    for( int d=0; d<3; d++ ) {
        double e = 0.;
        if( bc_remove[2*d] ) remove_particle_inf( c->species, imin, imax, d, c->patch->min_local_[d], g->dt, invgf, NULL, e );
        else internal_inf( c->species, imin, imax, d, c->patch->min_local_[d], g->dt, invgf, NULL, e );
        energy_tot += e;
        if( bc_remove[2*d+1] ) remove_particle_sup( c->species, imin, imax, d, c->patch->max_local_[d], g->dt, invgf, NULL, e );
        else internal_sup( c->species, imin, imax, d, c->patch->max_local_[d], g->dt, invgf, NULL, e );
        energy_tot += e;
    }
    std::memcpy( keys, ck, sizeof( int )*imax );
    std::memcpy( q, p.Charge.data(), sizeof( short )*imax );
    *energy_lost = energy_tot;
    free_ctx( c );
}

} // extern "C"

namespace {
/* a LaserProfile whose amplitude on the face is read from an array (what Laser::getAmplitude0/1 returns is
 * whatever the profile object computes; the reference sums it into b1/b2, ElectroMagnBC3D_SM.cpp:189-199,265-275) */
struct ArrayProfile : public LaserProfile {
    const double *a; int ld;
    ArrayProfile( const double *a_, int ld_ ) : a( a_ ), ld( ld_ ) {}
    double getAmplitude( std::vector<double>, double, int j, int k ) override { return a ? a[j*ld + k] : 0.; }
};
}

extern "C" {

void ref_apply_SM( const orc_grid *g, int i_boundary, const double *K, const int *is_boundary,
                   const double *Ex, const double *Ey, const double *Ez, double *Bx, double *By, double *Bz,
                   const double *db1, const double *db2 )
{
    Ctx *c = make_ctx( g, 1., 1 );
    Params &P = *c->params;
    Patch &pt = *c->patch;
    ElectroMagn3D &E = *c->em;
    const int axis0 = i_boundary/2, axis1 = axis0 == 0 ? 1 : 0, axis2 = axis0 == 2 ? 1 : 2;
    new( &P.EM_BCs_k ) std::vector<std::vector<double>>( 6, std::vector<double>( K, K+3 ) );
    new( &pt.size_ ) std::vector<unsigned int>( g->n, g->n+3 );
    new( &pt.oversize ) std::vector<unsigned int>( g->o, g->o+3 );
    new( &E.oversize ) std::vector<unsigned int>( g->o, g->o+3 );
    /* Patch::isBoundary reads neighbor_[axis][side] == MPI_PROC_NULL */
    new( &pt.neighbor_ ) std::vector<std::vector<int>>( 3, std::vector<int>( 2, 0 ) );
    const bool at_side = ( i_boundary % 2 ) == 0 ? g->pcoord[axis0] == 0 : g->pcoord[axis0] == g->npatch[axis0]-1;
    pt.neighbor_[axis0][i_boundary % 2] = at_side ? MPI_PROC_NULL : 0;
    pt.neighbor_[axis1][0] = is_boundary[0] ? MPI_PROC_NULL : 0;
    pt.neighbor_[axis1][1] = is_boundary[1] ? MPI_PROC_NULL : 0;
    pt.neighbor_[axis2][0] = is_boundary[2] ? MPI_PROC_NULL : 0;
    pt.neighbor_[axis2][1] = is_boundary[3] ? MPI_PROC_NULL : 0;
    load( E.Ex_, Ex ); load( E.Ey_, Ey ); load( E.Ez_, Ez );
    load( E.Bx_, Bx ); load( E.By_, By ); load( E.Bz_, Bz );
    {
        ElectroMagnBC3D_SM bc( P, &pt, ( unsigned int )i_boundary );
        Laser *laser = NULL;
        if( db1 || db2 ) {
            const int n2d = g->n[axis2] + 2*g->o[axis2] + 2, n2p = n2d - 1;
            laser = blank<Laser>();
            new( &laser->profiles ) std::vector<LaserProfile *>();
            laser->profiles.push_back( new ArrayProfile( db1, n2d ) );
            laser->profiles.push_back( new ArrayProfile( db2, n2p ) );
            bc.vecLaser.push_back( laser );
        }
        bc.apply( &E, 0., &pt );
        bc.vecLaser.clear();        /* the Laser stand-in is not the destructor's to delete */
    }
    store( E.Bx_, Bx ); store( E.By_, By ); store( E.Bz_, Bz );
    free_ctx( c );
}

} // extern "C"
