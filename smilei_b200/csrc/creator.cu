// creator.cu — host side, init time only: the reference's per-patch random streams.
//
// A namelist run of the reference starts from particles drawn patch by patch from a xorshift32 stream
// seeded with random_seed + the patch's Hilbert index (src/Patch/Patch.cpp:129, src/Tools/Random.h:91-140).
// To start the GPU path from the SAME particles for any rank layout, the driver walks the reference's
// patches inside its box and asks this file for the particles of each one, cell by cell and species by
// species, in the order of ParticleCreator::create (src/Particles/ParticleCreator.cpp:300-338).
// Nothing here runs on the device and nothing here is on the time-step path.
#include "common.cuh"
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Random (src/Tools/Random.h:91-140): xorshift32, zero seed replaced by 2^32-1
struct Stream {
    uint32_t s;
    explicit Stream( uint32_t state ) : s( state ? state : 4294967295u ) {}   // also how a saved state is resumed (never 0)
    uint32_t next() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
    double uniform()     { return next() * ( 1. / 4294967296. ); }              // ]0,1]
    double uniform2()    { return next() * ( 2. / 4294967296. ) - 1.; }         // ]-1,1]
    double uniform_2pi() { return next() * ( 2. * M_PI / 4294967296. ); }       // ]0,2pi]
};

// ---------------------------------------------------------------------------------------------
// Compact Hilbert index (C. Hamilton, Tech. Rep. CS-2006-07) as the reference uses it
// (src/DomainDecomposition/Hilbert_functions.cpp:10-133, 178-208, 246-296), written once for n = 2 or 3.
inline unsigned rot_r( unsigned v, unsigned s, unsigned n ) { return ( ( v >> s ) | ( v << ( n - s ) ) ) & ( ( 1u << n ) - 1u ); }
inline unsigned rot_l( unsigned v, unsigned s, unsigned n ) { return ( ( v << s ) | ( v >> ( n - s ) ) ) & ( ( 1u << n ) - 1u ); }
inline unsigned gray( unsigned i ) { return i ^ ( i >> 1 ); }
inline unsigned gray_inv( unsigned g )
{
    unsigned i = g;
    for( unsigned j = 1; ( 1u << j ) <= g; ++j ) i ^= g >> j;
    return i;
}
inline unsigned trailing_ones( unsigned i ) { unsigned k = 0; while( i & 1u ) { i >>= 1; ++k; } return k; }
inline unsigned intra_direction( unsigned w, unsigned n )
{
    if( w == 0 ) return 0;
    return ( ( w & 1u ) ? trailing_ones( w ) : trailing_ones( w - 1 ) ) % n;
}
inline unsigned entry_point( unsigned w ) { return w ? gray( 2 * ( ( w - 1 ) / 2 ) ) : 0; }

// index inside a cube of 2^m patches per side; e, d = entry point and direction, updated
unsigned cube_index( unsigned n, unsigned m, const unsigned *p, unsigned &e, unsigned &d )
{
    unsigned h = 0;
    for( int i = ( int )m - 1; i >= 0; --i ) {
        unsigned l = 0;
        for( unsigned a = 0; a < n; ++a ) l |= ( ( p[a] >> i ) & 1u ) << a;
        l = rot_r( l ^ e, d + 1, n );
        unsigned w = gray_inv( l );
        e ^= rot_l( entry_point( w ), d + 1, n );
        d = ( d + intra_direction( w, n ) + 1 ) % n;
        h = ( h << n ) | w;
    }
    return h;
}

// rectangle of 2^m0 x 2^m1 (Hilbert_functions.cpp:178-208): the long side is cut into squares first
unsigned rect_index( unsigned m0, unsigned m1, unsigned x, unsigned y, unsigned &e, unsigned &d )
{
    unsigned p[2] = { x, y }, h = 0;
    unsigned longd = m0 >= m1 ? 0 : 1, mmin = m0 >= m1 ? m1 : m0, mmax = m0 >= m1 ? m0 : m1;
    d = longd;
    for( int i = ( int )mmax - 1; i >= ( int )mmin; --i ) {
        unsigned l = ( p[longd] >> i ) & 1u;
        h += l * ( 1u << ( i + mmin ) );
        p[longd] -= l * ( 1u << i );
    }
    if( mmin > 0 ) h += cube_index( 2, mmin, p, e, d );
    return h;
}

// ---------------------------------------------------------------------------------------------
// ParticleCreator::maxwellJuttner (ParticleCreator.cpp:1002-1071): energies gamma-1 from the tabulated
// inverse cumulative functions; T in units of m c^2
void mj_energies( Stream &R, unsigned n, double T, const double *lnInvF, const double *lnInvH, double *energy )
{
    if( T < 0.1 ) {                                                       // Maxwell-Boltzmann
        const double inv_dU = 999. / ( 2. + 19. );
        for( unsigned i = 0; i < n; ++i ) {
            double U = R.uniform();
            double lnlnU = std::log( -std::log( U ) );
            double invF;
            if( lnlnU > 2. ) {
                invF = 3. * std::sqrt( M_PI ) / 4. * std::cbrt( U * U );
            } else if( lnlnU < -19. ) {
                invF = 1.;
            } else {
                double I = ( lnlnU + 19. ) * inv_dU;
                unsigned k = ( unsigned )I;
                double r = I - ( double )k;
                invF = std::exp( lnInvF[k] + r * ( lnInvF[k + 1] - lnInvF[k] ) );
            }
            energy[i] = T * invF;
        }
    } else {                                                              // Maxwell-Juttner, rejection on beta
        const double inv_dU = 999. / ( 12. + 30. );
        double invT = 1. / T;
        double H0 = -invT + std::log( 1. + invT + 0.5 * invT * invT );
        for( unsigned i = 0; i < n; ++i ) {
            double U, gamma;
            do {
                U = R.uniform();
                double lnU = std::log( -std::log( 1. - U ) - H0 ), invH;
                if( lnU < -26. ) {
                    invH = std::cbrt( -6. * U );
                } else if( lnU > 12. ) {
                    invH = -U + 11.35 * std::pow( -U, 0.06 );
                } else {
                    double I = ( lnU + 30. ) * inv_dU;
                    unsigned k = ( unsigned )I;
                    double r = I - ( double )k;
                    invH = std::exp( lnInvH[k] + r * ( lnInvH[k + 1] - lnInvH[k] ) );
                }
                gamma = T * invH;
                U = R.uniform();
            } while( U >= std::sqrt( 1. - 1. / ( gamma * gamma ) ) );
            energy[i] = gamma - 1.;
        }
    }
}

} // namespace

extern "C" {

int sb200_hilbert_index3d( unsigned m0, unsigned m1, unsigned m2, int x, int y, int z, unsigned *hindex )
{
    if( !hindex || m0 > 10 || m1 > 10 || m2 > 10 ) { sb200::set_error( "sb200_hilbert_index3d: bad argument" ); return 1; }
    if( x < 0 || x >= ( 1 << m0 ) || y < 0 || y >= ( 1 << m1 ) || z < 0 || z >= ( 1 << m2 ) ) {
        sb200::set_error( "sb200_hilbert_index3d: patch coordinates outside the box (the reference returns MPI_PROC_NULL)" );
        return 1;
    }
    unsigned mi[3] = { m0, m1, m2 }, p[3] = { ( unsigned )x, ( unsigned )y, ( unsigned )z };
    // longest, middle and shortest dimension (Hilbert_functions.cpp:266-279)
    unsigned dmax = ( m0 >= m1 && m0 >= m2 ) ? 0 : ( ( m1 > m0 && m1 >= m2 ) ? 1 : 2 );
    unsigned a = ( dmax + 1 ) % 3, b = ( dmax + 2 ) % 3;
    unsigned dmed = mi[a] >= mi[b] ? a : b, dmin = mi[a] >= mi[b] ? b : a;
    unsigned e = 0, d = 0, mm = mi[dmin];
    // the box flattened along its shortest dimension is a rectangle of cubes; then inside the cube
    unsigned h = rect_index( mi[dmax] - mm, mi[dmed] - mm, p[dmax] >> mm, p[dmed] >> mm, e, d ) * ( 1u << ( 3 * mm ) );
    unsigned mask = ( 1u << mm ) - 1u;
    unsigned q[3] = { p[dmax] & mask, p[dmed] & mask, p[dmin] & mask };
    h += cube_index( 3, mm, q, e, d );
    *hindex = h;
    return 0;
}

int sb200_create_particles_ref( unsigned int *rng_state, int position_init, int momentum_init,
                                const int box[3], const double box_min[3], const double cell_length[3],
                                const int *nppc, const double *n_real, const double *charge, const double *temperature,
                                double mass, const int regular_number[3],
                                const double *lnInvF, const double *lnInvH,
                                double *x, double *y, double *z, double *px, double *py, double *pz,
                                double *w, short *q, size_t capacity, size_t *n_created )
{
    if( !rng_state || !box || !box_min || !cell_length || !nppc || !n_real || !charge || !n_created
        || !x || !y || !z || !px || !py || !pz || !w || !q ) {
        sb200::set_error( "sb200_create_particles_ref: null argument" );
        return 1;
    }
    if( position_init < 0 || position_init > 3 || momentum_init < 0 || momentum_init > 1 ) {
        sb200::set_error( "sb200_create_particles_ref: position_init in {0 regular,1 random,2 centered,3 keep}, momentum_init in {0 cold,1 maxwell-juettner}" );
        return 1;
    }
    if( momentum_init == 1 && ( !temperature || !lnInvF || !lnInvH || !( mass > 0. ) ) ) {
        sb200::set_error( "sb200_create_particles_ref: maxwell-juettner needs temperature, the two tables and mass > 0" );
        return 1;
    }
    Stream R( *rng_state );                                 // the patch's stream goes on where the previous species left it
    double *pos[3] = { x, y, z };
    std::vector<double> energy;
    size_t ip = 0;
    for( int i = 0; i < box[0]; ++i ) for( int j = 0; j < box[1]; ++j ) for( int k = 0; k < box[2]; ++k ) {
        size_t c = ( ( size_t )i * box[1] + j ) * box[2] + k;
        if( !( n_real[c] > 0. ) || nppc[c] <= 0 ) continue;                                // ParticleCreator.cpp:308
        unsigned n = ( unsigned )nppc[c];
        if( ip + n > capacity ) { sb200::set_error( "sb200_create_particles_ref: capacity too small" ); return 1; }
        int ijk[3] = { i, j, k };
        double origin[3];
        for( int d = 0; d < 3; ++d ) origin[d] = ( unsigned )ijk[d] * cell_length[d] + box_min[d];   // :311-317
        // --- createPosition (:611-745)
        if( position_init == 0 ) {
            int cnt[3]; double inv[3];
            if( regular_number && regular_number[0] > 0 ) {
                if( ( unsigned )( regular_number[0] * regular_number[1] * regular_number[2] ) != n ) {
                    sb200::set_error( "The number of particles required per cell and per dimension is not coherent with the total number of particles per cell." );
                    return 1;
                }
                for( int d = 0; d < 3; ++d ) { cnt[d] = regular_number[d]; inv[d] = 1. / ( double )cnt[d]; }
            } else {
                const double coeff = std::pow( ( double )n, 1. / 3. );
                if( n != ( unsigned )std::floor( std::pow( std::round( coeff ), 3. ) ) ) {
                    sb200::set_error( "Impossible to put the particles regularly spaced in one cell. Use a square number, or `position_initialization = 'random'`" );
                    return 1;
                }
                for( int d = 0; d < 3; ++d ) { cnt[d] = ( int )coeff; inv[d] = 1. / coeff; }
            }
            for( unsigned p = 0; p < n; ++p ) {
                int r = ( int )p;
                for( int d = 0; d < 3; ++d ) {
                    pos[d][ip + p] = origin[d] + cell_length[d] * 0.975 * inv[d] * ( 0.5 + r % cnt[d] );
                    r /= cnt[d];
                }
            }
        } else if( position_init == 1 ) {
            for( unsigned p = 0; p < n; ++p )
                for( int d = 0; d < 3; ++d ) pos[d][ip + p] = origin[d] + R.uniform() * cell_length[d];
        } else if( position_init == 2 ) {
            for( unsigned p = 0; p < n; ++p )
                for( int d = 0; d < 3; ++d ) pos[d][ip + p] = origin[d] + 0.5 * cell_length[d];
        }
        // --- createMomentum (:818-851)
        if( momentum_init == 0 ) {
            for( unsigned p = 0; p < n; ++p ) px[ip + p] = py[ip + p] = pz[ip + p] = 0.;
        } else {
            energy.resize( n );
            mj_energies( R, n, temperature[c] / mass, lnInvF, lnInvH, energy.data() );
            for( unsigned p = 0; p < n; ++p ) {
                double phi   = std::acos( -R.uniform2() );
                double theta = R.uniform_2pi();
                double psm   = std::sqrt( ( 1.0 + energy[p] ) * ( 1.0 + energy[p] ) - 1.0 );
                px[ip + p] = psm * std::cos( theta ) * std::sin( phi );
                py[ip + p] = psm * std::sin( theta ) * std::sin( phi );
                pz[ip + p] = psm * std::cos( phi );
            }
        }
        // --- createWeight (:933-939), createCharge (:964-974, integer charges)
        double wt = n_real[c] / n;
        short  Z  = ( short )charge[c];
        if( charge[c] - ( double )Z != 0. ) { sb200::set_error( "sb200_create_particles_ref: non-integer charge profiles are not supported" ); return 1; }
        for( unsigned p = 0; p < n; ++p ) { w[ip + p] = wt; q[ip + p] = Z; }
        ip += n;
    }
    *rng_state = R.s;
    *n_created = ip;
    return 0;
}

} // extern "C"
