/* ref_creator_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the REFERENCE's own particle creation and patch numbering, compiled from where they lie under
 * /root/reference/src (src/Particles/ParticleCreator.cpp, src/DomainDecomposition/Hilbert_functions.cpp,
 * src/Tools/Random.h), so that the product's host-side restatement (smilei_b200/csrc/creator.cu) can be
 * pinned on them bit for bit.  Same blank-object technique as ref_harness.cpp: the static
 * member functions ParticleCreator::createPosition / createMomentum / createWeight / createCharge read a
 * handful of Species / Params members, set here as Species::Species and Params::Params set them.
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
#include "Species.h"
#include "SpeciesV.h"
#include "Particles.h"
#include "ParticleCreator.h"
#include "Random.h"
#include "Hilbert_functions.h"
#undef private
#undef protected

#include <cstdlib>
#include <cstring>
#include <new>

namespace {
template<class T> T *blank_obj()
{
    void *m = std::calloc( 1, sizeof( T ) + 64 );
    return reinterpret_cast<T *>( m );
}
}

extern "C" {

unsigned ref_hilbert_index3d( unsigned m0, unsigned m1, unsigned m2, int x, int y, int z )
{
    return generalhilbertindex( m0, m1, m2, x, y, z );
}

/* One cell of ParticleCreator::create's loop (ParticleCreator.cpp:308-336) on the patch stream `state`
 * (Patch::rand_, Patch.cpp:129).  position_init: "regular" | "random" | "centered" | "" (positions kept). */
int ref_create_cell( unsigned *state, const char *position_init, const char *momentum_init, unsigned nPart,
                     const double indexes_in[3], const double cell[3], double mass, const double temp_in[3],
                     double n_real, double charge, const int *regular_number,
                     double *x, double *y, double *z, double *px, double *py, double *pz, double *w, short *q )
{
    static Params   *P = NULL;
    static SpeciesV *S = NULL;
    if( !P ) {
        P = blank_obj<Params>();
        new( &P->geometry ) std::string( "3Dcartesian" );
        new( &P->cell_length ) std::vector<double>( 3 );
        S = blank_obj<SpeciesV>();
        new( &S->cell_length ) std::vector<double>( 3 );
        new( &S->name_ ) std::string( "harness" );
        S->nDim_particle = 3;
        S->nDim_field = 3;
        S->inv_nDim_particles = 1./( ( double )S->nDim_particle );      /* Species.cpp:107 */
        S->radial_velocity_profile_ = false;
    }
    for( int i=0; i<3; i++ ) {
        P->cell_length[i] = cell[i];
        S->cell_length[i] = cell[i];
    }
    S->mass_ = mass;
    Random rand( 1 );
    rand.xorshift32_state = *state;
    Particles part;
    part.initialize( nPart, 3, false );
    double indexes[3] = { indexes_in[0], indexes_in[1], indexes_in[2] };
    double temp[3] = { temp_in[0], temp_in[1], temp_in[2] }, vel[3] = { 0., 0., 0. };
    std::vector<int> regular;
    if( regular_number && regular_number[0] > 0 ) regular.assign( regular_number, regular_number+3 );
    std::string pinit( position_init );
    if( pinit.size() ) {
        ParticleCreator::createPosition( pinit, regular, &part, S, nPart, 0, indexes, *P, &rand );
    }
    ParticleCreator::createMomentum( std::string( momentum_init ), &part, S, nPart, 0, temp, vel, &rand );
    ParticleCreator::createWeight( &part, nPart, 0, n_real, *P, pinit == "regular" );
    ParticleCreator::createCharge( &part, S, nPart, 0, charge );
    if( pinit.size() ) {
        std::memcpy( x, part.Position[0].data(), sizeof( double )*nPart );
        std::memcpy( y, part.Position[1].data(), sizeof( double )*nPart );
        std::memcpy( z, part.Position[2].data(), sizeof( double )*nPart );
    }
    std::memcpy( px, part.Momentum[0].data(), sizeof( double )*nPart );
    std::memcpy( py, part.Momentum[1].data(), sizeof( double )*nPart );
    std::memcpy( pz, part.Momentum[2].data(), sizeof( double )*nPart );
    std::memcpy( w, part.Weight.data(), sizeof( double )*nPart );
    std::memcpy( q, part.Charge.data(), sizeof( short )*nPart );
    *state = rand.xorshift32_state;
    return 0;
}

/* The two tabulated inverse cumulative functions of ParticleCreator::maxwellJuttner. */
const double *ref_lnInvF( void ) { return ParticleCreator::lnInvF; }
const double *ref_lnInvH( void ) { return ParticleCreator::lnInvH; }

} // extern "C"
