// init.cu — device-side creation of the synthetic thermal plasma used by bench.py / smoke().
//
// Positions and weights follow the reference's ParticleCreator for position_initialization =
// "regular" with regular_number = ppc[3] (src/Particles/ParticleCreator.cpp:661-667:
//   x = cell_origin + cell_length*0.975*(0.5 + i%c)/c ), weight = density*cell_volume/nppc
// (:259, :933-939), charge as createCharge (:964-974).  Momenta are a non-relativistic
// Maxwellian of temperature T (per component sigma = sqrt(T/mass)) from a counter-based
// generator (SplitMix64 + Box-Muller), i.e. NOT the reference's xorshift32 Maxwell-Juttner
// stream: this initial state is synthetic by construction (bench.py says data = "synthetic");
// parity tests import explicit arrays through sb200_species_set instead.
#include "common.cuh"

namespace sb200 {

__device__ __forceinline__ unsigned long long splitmix64( unsigned long long x )
{
    x += 0x9E3779B97F4A7C15ull;
    x = ( x ^ ( x >> 30 ) )*0xBF58476D1CE4E5B9ull;
    x = ( x ^ ( x >> 27 ) )*0x94D049BB133111EBull;
    return x ^ ( x >> 31 );
}
__device__ __forceinline__ double u01( unsigned long long h ) { return ( ( h >> 11 ) + 0.5 )*( 1.0/9007199254740992.0 ); }

static unsigned long long splitmix64_host_seed( unsigned long long seed, int ispec )
{
    unsigned long long x = seed*0x9E3779B97F4A7C15ull + ( unsigned long long )( ispec+1 )*0xD1B54A32D192ED03ull;
    x = ( x ^ ( x >> 30 ) )*0xBF58476D1CE4E5B9ull;
    x = ( x ^ ( x >> 27 ) )*0x94D049BB133111EBull;
    return x ^ ( x >> 31 );
}

struct InitCols { double *c[7]; short *q; int *key; };

__global__ void __launch_bounds__( 256 ) k_init_thermal( GridDev g, InitCols out, int pc0, int pc1, int pc2,
        double weight, short charge, double sigma, unsigned long long seed, unsigned long long gid0, size_t n )
{
    const int nppc = pc0*pc1*pc2;
    for( size_t t = blockIdx.x*( size_t )blockDim.x + threadIdx.x; t < n; t += ( size_t )gridDim.x*blockDim.x ) {
        const size_t cell = t / nppc;
        int i = ( int )( t - cell*nppc );
        const int kc = ( int )( cell % g.n[2] );
        const size_t r = cell / g.n[2];
        const int jc = ( int )( r % g.n[1] );
        const int ic = ( int )( r / g.n[1] );
        const int ii = i % pc0; i /= pc0;
        const int jj = i % pc1; i /= pc1;
        const int kk = i;
        // cell origin in global coordinates = (pcoord*n + ic)*cell_length
        const double x = ( double )( g.pcoord[0]*g.n[0] + ic )*g.cell[0] + g.cell[0]*0.975*( 1./( double )pc0 )*( 0.5 + ii );
        const double y = ( double )( g.pcoord[1]*g.n[1] + jc )*g.cell[1] + g.cell[1]*0.975*( 1./( double )pc1 )*( 0.5 + jj );
        const double z = ( double )( g.pcoord[2]*g.n[2] + kc )*g.cell[2] + g.cell[2]*0.975*( 1./( double )pc2 )*( 0.5 + kk );
        const unsigned long long id = gid0 + t;
        const unsigned long long h0 = splitmix64( seed ^ ( id*4 + 0 ) ), h1 = splitmix64( seed ^ ( id*4 + 1 ) );
        const unsigned long long h2 = splitmix64( seed ^ ( id*4 + 2 ) ), h3 = splitmix64( seed ^ ( id*4 + 3 ) );
        const double r0 = sqrt( -2.*log( u01( h0 ) ) ), a0 = 6.283185307179586*u01( h1 );
        const double r1 = sqrt( -2.*log( u01( h2 ) ) ), a1 = 6.283185307179586*u01( h3 );
        out.c[0][t] = x; out.c[1][t] = y; out.c[2][t] = z;
        out.c[3][t] = sigma*r0*cos( a0 );
        out.c[4][t] = sigma*r0*sin( a0 );
        out.c[5][t] = sigma*r1*cos( a1 );
        out.c[6][t] = weight;
        out.q[t] = charge;
        out.key[t] = 0;
    }
}

} // namespace sb200

using namespace sb200;

extern "C" int sb200_species_init_thermal( sb200_patch *p, int ispec, const int ppc[3], double density, int charge,
        double temperature, unsigned long long seed )
{
    SB200_CHECK( p && ppc && ispec >= 0 && ispec < p->nspec, "sb200_species_init_thermal: bad arguments" );
    SB200_CHECK( ppc[0] > 0 && ppc[1] > 0 && ppc[2] > 0 && density > 0. && temperature >= 0., "sb200_species_init_thermal: bad plasma parameters" );
    SpeciesDev &s = p->sp[ispec];
    const GridDev &g = p->gd;
    const int nppc = ppc[0]*ppc[1]*ppc[2];
    const size_t n = ( size_t )g.n[0]*g.n[1]*g.n[2]*nppc;
    SB200_CHECK( n <= s.cap, "sb200_species_init_thermal: capacity too small (sb200_species_config first)" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    InitCols out;
    for( int c=0; c<7; c++ ) out.c[c] = s.col[c];
    out.q = s.q; out.key = s.key;
    const double weight = density*g.cell_volume/( double )nppc;
    const double sigma = sqrt( temperature/s.mass );
    // global particle id offset so that every patch of a decomposition draws distinct numbers
    const unsigned long long patch_lin = ( ( unsigned long long )g.pcoord[0]*g.npatch[1] + g.pcoord[1] )*g.npatch[2] + g.pcoord[2];
    k_init_thermal<<<148*16, 256, 0, p->stream>>>( g, out, ppc[0], ppc[1], ppc[2], weight, ( short )charge, sigma,
            splitmix64_host_seed( seed, ispec ), patch_lin*n, n );
            sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    s.n = n;
    s.n_sorted = 0;
    s.sorted = false;
    s.count_valid = false;
    s.perm_pending = false;
    SB200_CUDA( cudaMemsetAsync( s.d_qwmax, 0, sizeof( unsigned long long ), p->stream ) );
    return update_qwmax( p, ispec, 0, n );
}
