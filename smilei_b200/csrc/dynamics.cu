// dynamics.cu — Species::dynamics as ONE kernel per species: gather + push + boundary tag +
// next cell key + Esirkepov deposit, on cell-sorted SoA particles.
//
// Restates, per particle (file:line relative to the reference's src/):
//   Interpolator3D2Order::coeffs/compute/fieldsWrapper   Interpolator/Interpolator3D2Order.h:54-158, .cpp:163-285
//   Interpolator3D4Order::coeffs/compute/fieldsWrapper   Interpolator/Interpolator3D4Order.h:40-144, .cpp:159-217
//   PusherBoris / PusherVay / PusherHigueraCary          Pusher/PusherBoris.cpp:82-127, PusherVay.cpp:92-170, PusherHigueraCary.cpp:93-162
//   PartBoundCond::apply + internal_inf/internal_sup      ParticleBC/PartBoundCond.h:38-76, BoundaryConditionType.cpp:15-57
//   SpeciesV::computeParticleCellKeys                     Species/SpeciesV.cpp:796-812
//   Projector3D2Order::currents / Projector3D4Order       Projector/Projector3D2Order.cpp:55-343, Projector3D4Order.cpp:49-233
//
// Design: one CTA per tile of 4 x 4 x 8 primal-node cells.  Because the particles are
// sorted by cell key (z fastest), the tile's particles are 16 contiguous runs.  The CTA receives the E and
// B_m stencil boxes of the tile by TMA (cp.async.bulk.tensor), keeps a private fixed-point J box in shared
// memory and hands it to the TMA unit once (cp.reduce.async.bulk.tensor .add).  Order 2: k_dynamics_o2,
// order 4: k_dynamics_o4 (producer warps walk the particle stream, consumer warps form the per-cell current
// sums).  k_dynamics_cg (8-lane cell groups, shuffle transpose-reduction: round 1's order-4 kernel) and k_dynamics
// (one thread per particle, full Esirkepov window) are compiled only with -DSB200_AB_KERNELS, for A/B checks.
// Particle traffic is the
// algorithmic 110 B: read 7 doubles + 1 short, write 6 doubles + 1 int; Epart/Bpart/iold/deltaold/invgf
// never exist in HBM unless SB200_DYN_KEEP_SCRATCH asks for them.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cmath>

namespace sb200 {

constexpr int DYN_THREADS = 256;

template<int ORDER> struct Shape;
// Interpolator3D2Order.h:107-123
template<> struct Shape<2> {
    __device__ static __forceinline__ void w( double d, double *c )
    {
        const double d2 = d*d;
        c[0] = 0.5*( d2 - d + 0.25 );
        c[1] = 0.75 - d2;
        c[2] = 0.5*( d2 + d + 0.25 );
    }
};
// Interpolator3D4Order.h:69-73 with the constants of Interpolator3D4Order.cpp:24-34
template<> struct Shape<4> {
    __device__ static __forceinline__ void w( double d, double *c )
    {
        const double d2 = d*d, d3 = d2*d, d4 = d3*d;
        c[0] = 1.0/384.0   - 1.0/48.0*d  + 1.0/16.0*d2 - 1.0/12.0*d3 + 1.0/24.0*d4;
        c[1] = 19.0/96.0   - 11.0/24.0*d + 1.0/4.0*d2  + 1.0/6.0*d3  - 1.0/6.0*d4;
        c[2] = 115.0/192.0 - 5.0/8.0*d2  + 1.0/4.0*d4;
        c[3] = 19.0/96.0   + 11.0/24.0*d + 1.0/4.0*d2  - 1.0/6.0*d3  - 1.0/6.0*d4;
        c[4] = 1.0/384.0   + 1.0/48.0*d  + 1.0/16.0*d2 + 1.0/12.0*d3 + 1.0/24.0*d4;
    }
};

template<int ORDER, int TX_ = 4, int TY_ = 4, int TZ_ = 8> struct Tile {
    static constexpr int O  = ORDER;
    static constexpr int TX = TX_, TY = TY_, TZ = TZ_;
    static constexpr int H  = ORDER/2;
    static constexpr int NW = ORDER+1;            // gather points per dim
    static constexpr int WD = ORDER+3;            // Esirkepov window per dim (5 or 7)
    // staged field box; FZ has one spare point and is rounded up to even: a box row is a multiple of 16 B and
    // the box may start one element early so that its first element is 16-B aligned in HBM (TMA requirements)
    static constexpr int FX = TX+2*H+1, FY = TY+2*H+1, FZ = ( TZ+2*H+2 + 1 )/2*2;
    // J accumulation box; along z it may start one element early so that the TMA reduction of the box into
    // HBM begins on a 16-B boundary (needed when oversize-H-1 is odd), hence the spare and the even extent
    // At order 4 the even extent would be 16 doubles = 32 banks: the five rows j of a consumer's (cell, component)
    // lanes would add into the SAME bank (measured: 13 wavefronts per ATOMS instead of 1).  Two spare elements
    // (rows 36 words apart) spread them; the TMA reduction adds the two zero columns or clips them at the array edge.
    static constexpr int JX = TX+2*H+2, JY = TY+2*H+2, JZ = ( TZ+2*H+2 + ( ( ORDER-H-1 ) & 1 ) + 1 )/2*2 + ( ORDER == 4 ? 2 : 0 );
    static constexpr int FVOL = FX*FY*FZ, JVOL = JX*JY*JZ;
    static constexpr int FBOX = ( FVOL + 15 )/16*16;          // box strides in shared memory: 128-B aligned (TMA source / destination)
    static constexpr int JBOX = ( JVOL + 15 )/16*16;
    static constexpr size_t SMEM = ( size_t )( 6*FBOX + 3*JBOX )*sizeof( double );
};

// The tile's J box accumulates in 64-bit FIXED POINT: value*jscale rounded to an integer and added with the
// native 64-bit integer shared-memory atomic (sm_100 has no native double atomic on shared memory: a
// double atomicAdd there is a compare-and-swap loop).  jscale is a power of two chosen per launch from a
// rigorous bound of |J| in the box (see launch_dynamics), so nothing can overflow and the rounding error
// of one contribution is <= 2^-62 of that bound, i.e. far below 1 ulp of typical J values.  Integer
// addition is associative, so the box sums are also bitwise reproducible.
typedef unsigned long long jbox_t;
// 64-bit add as two NATIVE 32-bit shared-memory atomics (ATOMS.ADD) with carry: sm_100 implements a 64-bit
// atomicAdd on shared memory as a compare-and-swap loop (ATOMS.CAST.SPIN), measured 1.5-2.5x slower than this
// pair and much worse under address conflicts (tools/smem_atomic_bench.cu).  Every wrap of the low word is
// seen by exactly one adder, so the sum of carries is exact whatever the interleaving.
__device__ __forceinline__ void jadd_scaled( jbox_t *p, double vs )
{
    const long long iv = __double2ll_rn( vs );
    unsigned *w = reinterpret_cast<unsigned *>( p );
    const unsigned lo = ( unsigned )iv, hi = ( unsigned )( iv >> 32 );
    const unsigned old = atomicAdd( w, lo );
    atomicAdd( w+1, hi + ( ( old + lo ) < old ? 1u : 0u ) );
}
__device__ __forceinline__ void jadd( jbox_t *p, double v, double jscale ) { jadd_scaled( p, v*jscale ); }

struct DynArgs {
    double *col[7];           // columns written (pushed particles at their sorted slots)
    short  *q;
    int    *key;
    const double *in[7];      // columns read: == col unless a sort order is pending, then the unsorted set read through perm
    const short  *qin;
    const int    *perm;       // pending sort order (slot -> source index) or nullptr
    const int *first;
    const double *F[6];       // Ex Ey Ez Bxm Bym Bzm
    double *J[3];
    int    *count;            // histogram of the new cell keys (input of the next sort)
    int    *leave_counts;
    int    *leave_idx;        // [6][leave_cap]
    int     leave_cap;
    int    *iflags;
    double *sc_E, *sc_B, *sc_invgf, *sc_delta;
    int    *sc_iold;
    size_t  n;
    double  one_over_mass;
    double  jscale, jinv;     // fixed-point scale of the J box and its inverse
    int     tiles[3];
    int     any_remove;       // some entry of bc_remove is set
    int     bc_remove[6];     // 1: `remove` particle BC on this side (xmin xmax ymin ymax zmin zmax) AND the patch touches that global edge
    double *lost;             // accumulates w*(gamma-1) of the removed particles
};

// a particle tagged for exchange is counted per box side (sb200_leaving_pack compacts the boundary layer of the
// side itself, halo.cu: no index list is kept)
__device__ __forceinline__ void note_leaver( const DynArgs &a, int tag, size_t ip )
{
    atomicAdd( &a.leave_counts[-tag-2], 1 );
}

// PartBoundCond::apply on one particle (ParticleBC/PartBoundCond.h:38-76): the six boundary functions run in the
// order xmin xmax ymin ymax zmin zmax.  internal_inf/sup (BoundaryConditionType.cpp:15-57) tag a particle for
// exchange only while its key is still >= 0; remove_particle_inf/sup (:204-294) act whatever the key is: key = -1,
// charge = 0 (so the projector that follows deposits nothing) and w*(gamma-1) is added to the lost energy — once per
// boundary the particle is beyond, as in the reference.
template<bool REMOVE>
__device__ __forceinline__ int boundary_tag( const DynArgs &a, const GridDev &g, const double *npos, double px, double py, double pz,
                                             double weight, size_t ip, bool &removed )
{
    int tag = 0, nrem = 0;
    removed = false;
    if( !REMOVE ) {
        // all sides exchange (periodic box or inner patch): the first side the particle is beyond wins
#pragma unroll
        for( int d=0; d<3; d++ ) {
            if( tag == 0 ) {
                if( npos[d] < g.xmin[d] ) tag = -2 - 2*d;
                else if( npos[d] >= g.xmax[d] ) tag = -3 - 2*d;
            }
        }
        return tag;
    }
#pragma unroll
    for( int d=0; d<3; d++ ) {
        if( a.bc_remove[2*d] ) { if( npos[d] < g.xmin[d] ) { tag = -1; nrem++; } }
        else if( tag == 0 && npos[d] < g.xmin[d] ) tag = -2 - 2*d;
        if( a.bc_remove[2*d+1] ) { if( npos[d] >= g.xmax[d] ) { tag = -1; nrem++; } }
        else if( tag == 0 && npos[d] >= g.xmax[d] ) tag = -3 - 2*d;
    }
    removed = nrem > 0;
    if( removed ) {
        const double gf = sqrt( 1.0 + px*px + py*py + pz*pz );
        atomicAdd( a.lost, ( double )nrem*( weight*( gf - 1.0 ) ) );
        a.q[ip] = 0;
    }
    return tag;
}
__device__ __forceinline__ int boundary_tag( const DynArgs &a, const GridDev &g, const double *npos, double px, double py, double pz,
                                             double weight, size_t ip, bool &removed )
{
    return a.any_remove ? boundary_tag<true>( a, g, npos, px, py, pz, weight, ip, removed )
                        : boundary_tag<false>( a, g, npos, px, py, pz, weight, ip, removed );
}

// separable gather of one component from its staged box: sum_i cx[i] sum_j cy[j] sum_k cz[k] F
template<class T>
__device__ __forceinline__ double gather( const double *__restrict__ sF, const double *cx, const double *cy, const double *cz,
                                          int sx, int sy, int sz )
{
    double acc = 0.;
#pragma unroll
    for( int i=0; i<T::NW; i++ ) {
        double ai = 0.;
#pragma unroll
        for( int j=0; j<T::NW; j++ ) {
            const double *row = sF + ( ( sx - T::H + i )*T::FY + ( sy - T::H + j ) )*T::FZ + ( sz - T::H );
            double aj = 0.;
#pragma unroll
            for( int k=0; k<T::NW; k++ ) aj += cz[k]*row[k];
            ai += cy[j]*aj;
        }
        acc += cx[i]*ai;
    }
    return acc;
}

template<int PUSHER>
__device__ __forceinline__ void push( double cmd, double dt, double &px, double &py, double &pz,
                                      double Ex, double Ey, double Ez, double Bx, double By, double Bz,
                                      double &dxp, double &dyp, double &dzp, double &invgf_out )
{
    if( PUSHER == SB200_PUSHER_BORIS ) {
        double pxsm = cmd*Ex, pysm = cmd*Ey, pzsm = cmd*Ez;
        const double umx = px + pxsm, umy = py + pysm, umz = pz + pzsm;
        // rsqrt / reciprocal instead of sqrt + division: the same values to 1-2 ulp (the parity bar of the push is
        // 1e-12) for a third of the instructions and a much shorter dependent chain
        double local_invgf = cmd*rsqrt( 1.0 + umx*umx + umy*umy + umz*umz );
        const double Tx = local_invgf*Bx, Ty = local_invgf*By, Tz = local_invgf*Bz;
        const double inv_det_T = __drcp_rn( 1.0 + Tx*Tx + Ty*Ty + Tz*Tz );
        pxsm += ( ( 1.0+Tx*Tx-Ty*Ty-Tz*Tz )*umx + 2.0*( Tx*Ty+Tz )*umy + 2.0*( Tz*Tx-Ty )*umz )*inv_det_T;
        pysm += ( 2.0*( Tx*Ty-Tz )*umx + ( 1.0-Tx*Tx+Ty*Ty-Tz*Tz )*umy + 2.0*( Ty*Tz+Tx )*umz )*inv_det_T;
        pzsm += ( 2.0*( Tz*Tx+Ty )*umx + 2.0*( Ty*Tz-Tx )*umy + ( 1.0-Tx*Tx-Ty*Ty+Tz*Tz )*umz )*inv_det_T;
        local_invgf = rsqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
        invgf_out = local_invgf;
        px = pxsm; py = pysm; pz = pzsm;
        local_invgf *= dt;
        dxp = pxsm*local_invgf; dyp = pysm*local_invgf; dzp = pzsm*local_invgf;
    } else if( PUSHER == SB200_PUSHER_VAY ) {
        double invgf = rsqrt( 1.0 + px*px + py*py + pz*pz );      // rsqrt / reciprocal as in the Boris branch
        double upx = px + 2.*cmd*Ex, upy = py + 2.*cmd*Ey, upz = pz + 2.*cmd*Ez;
        double Tx = cmd*Bx, Ty = cmd*By, Tz = cmd*Bz;
        upx += invgf*( py*Tz - pz*Ty );
        upy += invgf*( pz*Tx - px*Tz );
        upz += invgf*( px*Ty - py*Tx );
        double alpha = 1.0 + upx*upx + upy*upy + upz*upz;
        const double T2 = Tx*Tx + Ty*Ty + Tz*Tz;
        double s = alpha - T2;
        double us2 = upx*Tx + upy*Ty + upz*Tz;
        us2 = us2*us2;
        alpha = rsqrt( 0.5*( s + sqrt( s*s + 4.0*( T2 + us2 ) ) ) );
        Tx *= alpha; Ty *= alpha; Tz *= alpha;
        s = __drcp_rn( 1.0 + Tx*Tx + Ty*Ty + Tz*Tz );
        alpha = upx*Tx + upy*Ty + upz*Tz;
        const double pxsm = s*( upx + alpha*Tx + Tz*upy - Ty*upz );
        const double pysm = s*( upy + alpha*Ty + Tx*upz - Tz*upx );
        const double pzsm = s*( upz + alpha*Tz + Ty*upx - Tx*upy );
        invgf = rsqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
        invgf_out = invgf;
        px = pxsm; py = pysm; pz = pzsm;
        dxp = dt*pxsm*invgf; dyp = dt*pysm*invgf; dzp = dt*pzsm*invgf;
    } else {
        double pxsm = cmd*Ex, pysm = cmd*Ey, pzsm = cmd*Ez;
        const double umx = px + pxsm, umy = py + pysm, umz = pz + pzsm;
        const double gfm2 = 1.0 + umx*umx + umy*umy + umz*umz;
        double Tx = cmd*Bx, Ty = cmd*By, Tz = cmd*Bz;
        const double beta2 = Tx*Tx + Ty*Ty + Tz*Tz;
        const double Tum = Tx*umx + Ty*umy + Tz*umz;
        const double local_invgf = rsqrt( 0.5*( gfm2 - beta2 + sqrt( ( gfm2-beta2 )*( gfm2-beta2 ) + 4.0*( beta2 + Tum*Tum ) ) ) );
        Tx *= local_invgf; Ty *= local_invgf; Tz *= local_invgf;
        const double Tx2 = Tx*Tx, Ty2 = Ty*Ty, Tz2 = Tz*Tz, TxTy = Tx*Ty, TyTz = Ty*Tz, TzTx = Tz*Tx;
        const double inv_det_T = __drcp_rn( 1.0 + Tx2 + Ty2 + Tz2 );
        const double upx = ( ( 1.0+Tx2-Ty2-Tz2 )*umx + 2.0*( TxTy+Tz )*umy + 2.0*( TzTx-Ty )*umz )*inv_det_T;
        const double upy = ( 2.0*( TxTy-Tz )*umx + ( 1.0-Tx2+Ty2-Tz2 )*umy + 2.0*( TyTz+Tx )*umz )*inv_det_T;
        const double upz = ( 2.0*( TzTx+Ty )*umx + 2.0*( TyTz-Tx )*umy + ( 1.0-Tx2-Ty2+Tz2 )*umz )*inv_det_T;
        pxsm += upx; pysm += upy; pzsm += upz;
        const double invgf = rsqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
        invgf_out = invgf;
        px = pxsm; py = pysm; pz = pzsm;
        dxp = dt*pxsm*invgf; dyp = dt*pysm*invgf; dzp = dt*pzsm*invgf;
    }
}

// S1 on the WD-point window from the NW weights `w` at shift s = ip - ipo in {-1,0,1}
// (Projector3D2Order.cpp:124-152: Sx1[ip_m_ipo+1 .. +3] = weights)
template<class T>
__device__ __forceinline__ void place_S1( const double *w, int shift, double *S1 )
{
#pragma unroll
    for( int s=0; s<T::WD; s++ ) {
        const double a = ( s-1 >= 0 && s-1 < T::NW ) ? w[( s-1 >= 0 && s-1 < T::NW ) ? s-1 : 0] : 0.;   // shift  0
        const double b = ( s-2 >= 0 && s-2 < T::NW ) ? w[( s-2 >= 0 && s-2 < T::NW ) ? s-2 : 0] : 0.;   // shift +1
        const double c = ( s   >= 0 && s   < T::NW ) ? w[( s   >= 0 && s   < T::NW ) ? s   : 0] : 0.;   // shift -1
        S1[s] = shift == 0 ? a : ( shift > 0 ? b : c );
    }
}

// General Esirkepov deposit of one particle into the tile's J box (any cell crossing):
// Projector3D2Order::currents (Projector3D2Order.cpp:160-340) / Projector3D4Order::currents
// (Projector3D4Order.cpp:191-230) in outer-product form: J[i][j][k] += C[i]*W[j][k] with
// C[i] = -cr * sum_{i'<i} DS[i'] (the reference's running sum over the flux direction).
template<class T>
__device__ __forceinline__ void esirkepov_general( jbox_t *jb, const double ( &S0 )[3][T::WD], const double ( &DS )[3][T::WD], const double *cr, double jscale )
{
    const double third = 1./3.;
    // Jx: flux along x, weights over (y,z)
    {
        double C[T::WD];
        double run = 0.;
        C[0] = 0.;
#pragma unroll
        for( int i=1; i<T::WD; i++ ) { run -= cr[0]*DS[0][i-1]; C[i] = run; }
#pragma unroll
        for( int j=0; j<T::WD; j++ ) {
#pragma unroll
            for( int k=0; k<T::WD; k++ ) {
                const double W = S0[1][j]*S0[2][k] + 0.5*DS[1][j]*S0[2][k] + 0.5*DS[2][k]*S0[1][j] + third*DS[1][j]*DS[2][k];
                if( W != 0. ) {
#pragma unroll
                    for( int i=1; i<T::WD; i++ ) {
                        const double v = C[i]*W;
                        if( v != 0. ) jadd( jb + 0*T::JBOX + ( i*T::JY + j )*T::JZ + k, v, jscale );
                    }
                }
            }
        }
    }
    // Jy: flux along y, weights over (z,x)
    {
        double C[T::WD];
        double run = 0.;
        C[0] = 0.;
#pragma unroll
        for( int j=1; j<T::WD; j++ ) { run -= cr[1]*DS[1][j-1]; C[j] = run; }
#pragma unroll
        for( int i=0; i<T::WD; i++ ) {
#pragma unroll
            for( int k=0; k<T::WD; k++ ) {
                const double W = S0[2][k]*S0[0][i] + 0.5*DS[2][k]*S0[0][i] + 0.5*DS[0][i]*S0[2][k] + third*DS[2][k]*DS[0][i];
                if( W != 0. ) {
#pragma unroll
                    for( int j=1; j<T::WD; j++ ) {
                        const double v = C[j]*W;
                        if( v != 0. ) jadd( jb + 1*T::JBOX + ( i*T::JY + j )*T::JZ + k, v, jscale );
                    }
                }
            }
        }
    }
    // Jz: flux along z, weights over (x,y)
    {
        double C[T::WD];
        double run = 0.;
        C[0] = 0.;
#pragma unroll
        for( int k=1; k<T::WD; k++ ) { run -= cr[2]*DS[2][k-1]; C[k] = run; }
#pragma unroll
        for( int i=0; i<T::WD; i++ ) {
#pragma unroll
            for( int j=0; j<T::WD; j++ ) {
                const double W = S0[0][i]*S0[1][j] + 0.5*DS[0][i]*S0[1][j] + 0.5*DS[1][j]*S0[0][i] + third*DS[0][i]*DS[1][j];
                if( W != 0. ) {
#pragma unroll
                    for( int k=1; k<T::WD; k++ ) {
                        const double v = C[k]*W;
                        if( v != 0. ) jadd( jb + 2*T::JBOX + ( i*T::JY + j )*T::JZ + k, v, jscale );
                    }
                }
            }
        }
    }
}

#ifdef SB200_AB_KERNELS
template<int ORDER, int PUSHER, bool SCRATCH>
__global__ void __launch_bounds__( DYN_THREADS ) k_dynamics( const GridDev g, const DynArgs a )
{
    using T = Tile<ORDER>;
    extern __shared__ double smem[];
    double *sF = smem;                    // 6 boxes of FVOL
    jbox_t *sJ = reinterpret_cast<jbox_t *>( smem + 6*T::FBOX );        // 3 boxes of JVOL (fixed point)
    __shared__ int row_off[T::TX*T::TY+1];
    __shared__ int row_base[T::TX*T::TY];

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int tz = b % a.tiles[2]; b /= a.tiles[2];
    const int ty = b % a.tiles[1];
    const int tx = b / a.tiles[1];
    const int c0[3] = { tx*T::TX, ty*T::TY, tz*T::TZ };

    // particle runs of the tile
    if( tid < T::TX*T::TY ) {
        const int ix = c0[0] + tid/T::TY, iy = c0[1] + tid%T::TY;
        int beg = 0, end = 0;
        if( ix < g.ncell[0] && iy < g.ncell[1] ) {
            const int kz1 = min( c0[2]+T::TZ, g.ncell[2] );
            const int cell = ( ix*g.ncell[1] + iy )*g.ncell[2] + c0[2];
            beg = a.first[cell];
            end = a.first[cell + ( kz1 - c0[2] )];
        }
        row_base[tid] = beg;
        row_off[tid+1] = end - beg;
    }
    __syncthreads();
    if( tid == 0 ) {
        int s = 0;
        row_off[0] = 0;
        for( int r=0; r<T::TX*T::TY; r++ ) { s += row_off[r+1]; row_off[r+1] = s; }
    }
    __syncthreads();
    const int total = row_off[T::TX*T::TY];
    if( total == 0 ) return;

    // stage the field boxes: box index s <-> array index gs + s, gs = c0 + o - H
    const int gs[3] = { c0[0] + g.o[0] - T::H, c0[1] + g.o[1] - T::H, c0[2] + g.o[2] - T::H };
    for( int t = tid; t < 6*T::FVOL; t += DYN_THREADS ) {
        const int c = t / T::FVOL;
        int r = t - c*T::FVOL;
        const int k = r % T::FZ; r /= T::FZ;
        const int j = r % T::FY;
        const int i = r / T::FY;
        const int gi = gs[0]+i, gj = gs[1]+j, gk = gs[2]+k;
        double v = 0.;
        if( gi < g.ax && gj < g.ay && gk < g.az ) v = a.F[c][gi*g.sx + gj*g.sy + gk];
        sF[c*T::FBOX + ( t - c*T::FVOL )] = v;
    }
    for( int t = tid; t < 3*T::JBOX; t += DYN_THREADS ) sJ[t] = 0ull;
    __syncthreads();

    for( int wi = tid; wi < total; wi += DYN_THREADS ) {
        // row of this work item (binary search in row_off)
        int lo = 0, hi = T::TX*T::TY;
        while( hi - lo > 1 ) { const int mid = ( lo+hi ) >> 1; if( row_off[mid] <= wi ) lo = mid; else hi = mid; }
        const size_t ip = ( size_t )row_base[lo] + ( size_t )( wi - row_off[lo] );

        double pos[3] = { a.col[0][ip], a.col[1][ip], a.col[2][ip] };
        double px = a.col[3][ip], py = a.col[4][ip], pz = a.col[5][ip];
        const double weight = a.col[6][ip];
        const short charge = a.q[ip];

        // ---- coefficients at the old position (Interpolator3D{2,4}Order::coeffs)
        double cp[3][T::NW], cd[3][T::NW], delta_p[3];
        int sp[3], sd[3], cl[3];   // box indices (primal, dual) and cell offset in the tile
        bool bad = false;
#pragma unroll
        for( int d=0; d<3; d++ ) {
            const double pn = pos[d]*g.dxi[d];
            const int ipn = ( int )round( pn );
            delta_p[d] = pn - ( double )ipn;
            Shape<ORDER>::w( delta_p[d], cp[d] );
            const int idn = ( int )round( pn + 0.5 );
            const double dd = pn - ( double )idn + 0.5;
            Shape<ORDER>::w( dd, cd[d] );
            int c = ipn - g.begin[d] - g.o[d] - c0[d];       // cell offset inside the tile
            const int tdim = d==0 ? T::TX : d==1 ? T::TY : T::TZ;
            if( c < 0 || c >= tdim ) { bad = true; c = c < 0 ? 0 : tdim-1; }
            cl[d] = c;
            sp[d] = c + T::H;
            sd[d] = sp[d] + ( idn - ipn );
        }
        if( bad ) atomicAdd( &a.iflags[1], 1 );   // particle not in the cell its sort key says

        // ---- gather (fieldsWrapper): Ex(d,p,p) Ey(p,d,p) Ez(p,p,d) Bx(p,d,d) By(d,p,d) Bz(d,d,p)
        const double Ex = gather<T>( sF+0*T::FBOX, cd[0], cp[1], cp[2], sd[0], sp[1], sp[2] );
        const double Ey = gather<T>( sF+1*T::FBOX, cp[0], cd[1], cp[2], sp[0], sd[1], sp[2] );
        const double Ez = gather<T>( sF+2*T::FBOX, cp[0], cp[1], cd[2], sp[0], sp[1], sd[2] );
        const double Bx = gather<T>( sF+3*T::FBOX, cp[0], cd[1], cd[2], sp[0], sd[1], sd[2] );
        const double By = gather<T>( sF+4*T::FBOX, cd[0], cp[1], cd[2], sd[0], sp[1], sd[2] );
        const double Bz = gather<T>( sF+5*T::FBOX, cd[0], cd[1], cp[2], sd[0], sd[1], sp[2] );

        // ---- push
        const double cmd = ( double )charge*a.one_over_mass*g.dts2;
        double dxp, dyp, dzp, invgf;
        push<PUSHER>( cmd, g.dt, px, py, pz, Ex, Ey, Ez, Bx, By, Bz, dxp, dyp, dzp, invgf );
        double npos[3] = { pos[0] + dxp, pos[1] + dyp, pos[2] + dzp };

        a.col[0][ip] = npos[0]; a.col[1][ip] = npos[1]; a.col[2][ip] = npos[2];
        a.col[3][ip] = px; a.col[4][ip] = py; a.col[5][ip] = pz;

        if( SCRATCH ) {
            a.sc_E[0*a.n+ip] = Ex; a.sc_E[1*a.n+ip] = Ey; a.sc_E[2*a.n+ip] = Ez;
            a.sc_B[0*a.n+ip] = Bx; a.sc_B[1*a.n+ip] = By; a.sc_B[2*a.n+ip] = Bz;
            a.sc_invgf[ip] = invgf;
#pragma unroll
            for( int d=0; d<3; d++ ) {
                a.sc_iold[d*a.n+ip] = cl[d] + c0[d] + g.o[d];
                a.sc_delta[d*a.n+ip] = delta_p[d];
            }
        }

        // ---- S0 / S1 / DS on the Esirkepov window (Projector3D2Order.cpp:99-159)
        double S0[3][T::WD], DS[3][T::WD];
        int    nkey[3];
        bool   removed;
        const int tag = boundary_tag( a, g, npos, px, py, pz, weight, ip, removed );
#pragma unroll
        for( int d=0; d<3; d++ ) {
            const double pn = npos[d]*g.dxi[d];
            const int ipn = ( int )round( pn );
            const double dl = pn - ( double )ipn;
            double w1[T::NW], S1[T::WD];
            Shape<ORDER>::w( dl, w1 );
            const int shift = ipn - g.begin[d] - ( cl[d] + c0[d] + g.o[d] );    // ip - ipo - i_domain_begin
            place_S1<T>( w1, shift, S1 );
            S0[d][0] = 0.; S0[d][T::WD-1] = 0.;
#pragma unroll
            for( int s=0; s<T::NW; s++ ) S0[d][s+1] = cp[d][s];
#pragma unroll
            for( int s=0; s<T::WD; s++ ) DS[d][s] = S1[s] - S0[d][s];
            nkey[d] = ( int )( ( double )ipn - g.min_loc_round[d] );
        }
        int key = tag;
        if( tag == 0 ) { key = ( nkey[0]*g.ncell[1] + nkey[1] )*g.ncell[2] + nkey[2]; atomicAdd( &a.count[key], 1 ); }
        else if( tag < -1 ) note_leaver( a, tag, ip );
        a.key[ip] = key;

        // ---- currents (Esirkepov), accumulated in the tile's J box
        const double charge_weight = removed ? 0. : g.inv_cell_volume*( double )charge*weight;
        const double cr[3] = { charge_weight*g.d_ov_dt[0], charge_weight*g.d_ov_dt[1], charge_weight*g.d_ov_dt[2] };
        jbox_t *jb = sJ + ( cl[0]*T::JY + cl[1] )*T::JZ + cl[2];
        esirkepov_general<T>( jb, S0, DS, cr, a.jscale );
    }
    __syncthreads();

    // flush the J box: box index s <-> array index c0 + o - H - 1 + s
    const int js[3] = { c0[0] + g.o[0] - T::H - 1, c0[1] + g.o[1] - T::H - 1, c0[2] + g.o[2] - T::H - 1 };
    for( int t = tid; t < 3*T::JVOL; t += DYN_THREADS ) {
        const int c = t / T::JVOL;
        int r = t - c*T::JVOL;
        const long long iv = ( long long )sJ[c*T::JBOX + r];
        if( iv == 0 ) continue;
        const double v = ( double )iv*a.jinv;
        const int k = r % T::JZ; r /= T::JZ;
        const int j = r % T::JY;
        const int i = r / T::JY;
        const int gi = js[0]+i, gj = js[1]+j, gk = js[2]+k;
        if( gi >= 0 && gj >= 0 && gk >= 0 && gi < g.ax && gj < g.ay && gk < g.az )
            atomicAdd( a.J[c] + gi*g.sx + gj*g.sy + gk, v );
    }
}

template<int ORDER, int PUSHER, bool SCRATCH>
static int launch_one( sb200_patch *p, const DynArgs &a, int ntiles )
{
    using T = Tile<ORDER>;
    auto kern = k_dynamics<ORDER, PUSHER, SCRATCH>;
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )T::SMEM ) );
    kern<<<ntiles, DYN_THREADS, T::SMEM, p->stream>>>( p->gd, a );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

template<int ORDER, int PUSHER>
static int launch_scratch( sb200_patch *p, const DynArgs &a, int ntiles, bool scratch )
{
    return scratch ? launch_one<ORDER, PUSHER, true>( p, a, ntiles ) : launch_one<ORDER, PUSHER, false>( p, a, ntiles );
}

template<int ORDER>
static int launch_pusher( sb200_patch *p, const DynArgs &a, int ntiles, int pusher, bool scratch )
{
    switch( pusher ) {
        case SB200_PUSHER_BORIS: return launch_scratch<ORDER, SB200_PUSHER_BORIS>( p, a, ntiles, scratch );
        case SB200_PUSHER_VAY: return launch_scratch<ORDER, SB200_PUSHER_VAY>( p, a, ntiles, scratch );
        default: return launch_scratch<ORDER, SB200_PUSHER_HIGUERACARY>( p, a, ntiles, scratch );
    }
}

#endif // SB200_AB_KERNELS

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers: raw PTX for sm_100a -----------------------------------
struct FieldMaps { CUtensorMap m[6]; CUtensorMap j[3]; };   // Ex Ey Ez Bxm Bym Bzm boxes (FZ,FY,FX); Jx Jy Jz boxes (JZ,JY,JX)

__device__ __forceinline__ unsigned smem_u32( const void *p ) { return ( unsigned )__cvta_generic_to_shared( p ); }
__device__ __forceinline__ void tma_bar_init( unsigned long long *bar, unsigned count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"( smem_u32( bar ) ), "r"( count ) : "memory" );
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
}
__device__ __forceinline__ void tma_expect( unsigned long long *bar, unsigned bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void tma_load_3d( void *dst, const CUtensorMap *map, unsigned long long *bar, int x0, int x1, int x2 )
{
    asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  :: "r"( smem_u32( dst ) ), "l"( ( unsigned long long )map ), "r"( smem_u32( bar ) ), "r"( x0 ), "r"( x1 ), "r"( x2 ) : "memory" );
}
// shared -> global element-wise ADD of a box (f64), performed by the TMA unit / L2 atomically per element;
// elements of the box that fall outside the tensor are dropped
__device__ __forceinline__ void tma_reduce_add_3d( const CUtensorMap *map, const void *src, int x0, int x1, int x2 )
{
    asm volatile( "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                  :: "l"( ( unsigned long long )map ), "r"( smem_u32( src ) ), "r"( x0 ), "r"( x1 ), "r"( x2 ) : "memory" );
}
__device__ __forceinline__ void tma_store_fence() { asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); }
__device__ __forceinline__ void tma_commit_and_wait_read()
{
    asm volatile( "cp.async.bulk.commit_group;" ::: "memory" );
    asm volatile( "cp.async.bulk.wait_group.read 0;" ::: "memory" );
}
__device__ __forceinline__ void tma_wait( unsigned long long *bar, unsigned phase )
{
    asm volatile( "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                  :: "r"( smem_u32( bar ) ), "r"( phase ) : "memory" );
}

// =================================================================================================
// Cell-group kernel, orders 2 and 4.
//
// What binds the general kernel above is not HBM: every particle issues (ORDER+2)*(ORDER+3)^2 accumulations
// per current component on shared memory, all lanes of a warp on the same few addresses because sorted
// neighbours sit in the same cell.  This kernel removes them from the common case:
//   * work item = (cell, round): G = 8 consecutive lanes take up to 8 particles OF THE SAME CELL;
//   * a particle whose primal node does not change during the step (the large majority in a thermal
//     plasma: |dx| << 1 cell) deposits only on the NW^3 nodes its shape covers (NW = ORDER+1): per current
//     component NW-1 flux points x NW x NW values.  They are formed in registers, NV at a time, summed over
//     the 8 lanes with a shuffle transpose-reduction (order 2: 18 -> 9 -> 5 -> 3 values per lane; order 4:
//     25 -> 13 -> 7 -> 4, once per flux point), and only the per-cell sums reach the tile's J box;
//   * particles that do change cell go to a per-warp queue; whenever four are pending the warp deposits them,
//     one per 8-lane group: their Esirkepov window is exactly NW+1 points wide per dimension (S0 on 1..NW, S1
//     shifted by -1/0/+1), each lane takes a share of the (NW+1)^2 transverse positions and the NW non-zero
//     flux points of each component.
// Gather, push, boundary tagging and next-key computation are as in the general kernel.
// =================================================================================================
constexpr int GRP = 8;                                   // lanes per cell group
constexpr int XQ = 16;                                   // cell-crossers a warp can keep pending
constexpr int XQD = 9*XQ + XQ/2;                         // doubles per warp queue: deltaold[3], new pos[3], cr[3] (SoA) + XQ ints

template<int ORDER> struct CG;
template<> struct CG<2> {
    using T = Tile<2, 4, 4, 8>;
    static constexpr int NSL = 2;      // flux points reduced together
    static constexpr int MINB = 2;     // CTAs per SM asked of the compiler
    static constexpr int LOG2CELLS = 7;
};
template<> struct CG<4> {
    using T = Tile<4, 4, 4, 8>;
    static constexpr int NSL = 1;
    static constexpr int MINB = 2;
    static constexpr int LOG2CELLS = 7;
};
template<int ORDER> struct CGDim {
    using T = typename CG<ORDER>::T;
    static constexpr int NW = ORDER+1;                     // non-zero shape points
    static constexpr int NSL = CG<ORDER>::NSL;
    static constexpr int NPASS = ( NW-1 )/NSL;             // reductions per component
    static constexpr int NV = NSL*NW*NW;                   // values reduced together (18 or 25)
    static constexpr int N1 = ( NV+1 )/2, N2 = ( N1+1 )/2, NS = ( N2+1 )/2;   // after each of the 3 steps
    static constexpr int WX = NW+1;                        // crosser window width
    static constexpr int XSCR = 6*WX;                      // crosser scratch per lane group: S0[3][WX], DS[3][WX]
    static constexpr int NCELL = T::TX*T::TY*T::TZ;
    static constexpr int CPT = ( NCELL + DYN_THREADS - 1 )/DYN_THREADS;
    static constexpr size_t BYTES = ( size_t )( 6*T::FBOX + 3*T::JBOX + XSCR*( DYN_THREADS/GRP ) + XQD*( DYN_THREADS/32 ) )*sizeof( double );
    static constexpr unsigned TMA_BYTES = 6u*T::FVOL*sizeof( double );   // bytes the six box loads deliver
};
using TileO2 = CG<2>::T;

// one step of the transpose-reduction: N values -> (N+1)/2 values; lanes with `upper` keep the second half
template<int N>
__device__ __forceinline__ void xr_step( double *v, int lane_mask, bool upper )
{
    constexpr int H = ( N+1 )/2;
#pragma unroll
    for( int i=0; i<H; i++ ) {
        const double a = v[i];
        const double b = ( i+H < N ) ? v[i+H] : 0.;
        const double send = upper ? a : b;
        const double keep = upper ? b : a;
        v[i] = keep + __shfl_xor_sync( 0xffffffffu, send, lane_mask );
    }
}

// select w[t] for a run-time t (registers cannot be indexed dynamically), 0 outside [0,N)
template<int N>
__device__ __forceinline__ double pick( const double *w, int t )
{
    double r = 0.;
#pragma unroll
    for( int i=0; i<N; i++ ) r = t == i ? w[i] : r;
    return r;
}

// One pass over the warp's crosser queue: lane group `grp` (0..3) deposits entry head+grp if it exists.
// The Esirkepov window of a particle that changed cell is exactly WX = NW+1 points wide per dimension,
// starting at lo = (shift<0 ? 0 : 1): S0 on window points 1..NW, S1 on 1+shift..NW+shift.  Lanes 0..2 of
// the group evaluate S0/DS of one dimension each into the group's scratch; then every lane takes its share
// of the WX x WX transverse positions and the NW non-zero flux points (lo+1..lo+NW) of each component.
template<int ORDER, int XQ = sb200::XQ>
__device__ __forceinline__ void cross_pass( jbox_t *sJ, const double *xq, const int *xqm, double *xscr, int qh, int qn,
                                            int gl, int grp, double jscale )
{
    using D = CGDim<ORDER>;
    using T = typename D::T;
    constexpr int NW = D::NW, WX = D::WX;
    const bool work = grp < qn;
    const int e = ( qh + ( work ? grp : 0 ) ) % XQ;
    const int meta = xqm[e];
    const int cellt = meta & 0xffff, bsh = meta >> 16;
    if( work && gl < 3 ) {
        const int d = gl;
        const double dl0 = xq[( 0+d )*XQ+e];
        const double pn  = xq[( 3+d )*XQ+e];
        const int shift = ( ( bsh >> ( 2*d ) ) & 3 ) - 1;
        double w0[NW], w1[NW];
        Shape<ORDER>::w( dl0, w0 );
        Shape<ORDER>::w( pn - round( pn ), w1 );
        const int lo_ = shift < 0 ? 0 : 1;
#pragma unroll
        for( int s=0; s<WX; s++ ) {
            const double s0 = pick<NW>( w0, lo_ + s - 1 );
            const double s1 = pick<NW>( w1, lo_ + s - 1 - shift );
            xscr[d*WX+s] = s0;
            xscr[3*WX+d*WX+s] = s1 - s0;
        }
    }
    __syncwarp();
    if( work ) {
        const double third = 1./3.;
        const int cl[3] = { cellt / ( T::TZ*T::TY ), ( cellt / T::TZ ) % T::TY, cellt % T::TZ };
        const int lo0 = ( ( bsh      ) & 3 ) == 0 ? 0 : 1;
        const int lo1 = ( ( bsh >> 2 ) & 3 ) == 0 ? 0 : 1;
        const int lo2 = ( ( bsh >> 4 ) & 3 ) == 0 ? 0 : 1;
        const double *S0x = xscr, *S0y = xscr+WX, *S0z = xscr+2*WX, *DSx = xscr+3*WX, *DSy = xscr+4*WX, *DSz = xscr+5*WX;
        jbox_t *xb = sJ + ( ( cl[0]+lo0 )*T::JY + ( cl[1]+lo1 ) )*T::JZ + ( cl[2]+lo2 );
        const double c0x = xq[6*XQ+e]*jscale, c1y = xq[7*XQ+e]*jscale, c2z = xq[8*XQ+e]*jscale;   // fixed-point scale folded in
        double Cx[NW], Cy[NW], Cz[NW];     // flux coefficients at window points lo+1 .. lo+NW
        {
            double rx = 0., ry = 0., rz = 0.;
#pragma unroll
            for( int f=0; f<NW; f++ ) {
                rx -= c0x*DSx[f]; ry -= c1y*DSy[f]; rz -= c2z*DSz[f];
                Cx[f] = rx; Cy[f] = ry; Cz[f] = rz;
            }
        }
        // NOT unrolled: the body holds 3*NW inlined fixed-point adds; unrolled (5 copies at order 4) the kernel no
        // longer fits the instruction cache and two thirds of the stall samples of an electron launch were
        // instruction fetches
#pragma unroll 1
        for( int h=0; h<( WX*WX + GRP - 1 )/GRP; h++ ) {
            const int pp = gl + GRP*h;
            if( pp < WX*WX ) {
                const int aa = pp / WX, bb = pp % WX;
                const double Az = S0z[bb] + 0.5*DSz[bb], Bz = 0.5*S0z[bb] + third*DSz[bb];
                const double Wx = S0y[aa]*Az + DSy[aa]*Bz;                                                       // Jx: W(j=aa,k=bb)
                const double Wy = S0x[aa]*Az + DSx[aa]*Bz;                                                       // Jy: W(i=aa,k=bb)
                const double Wz = S0x[aa]*( S0y[bb] + 0.5*DSy[bb] ) + DSx[aa]*( 0.5*S0y[bb] + third*DSy[bb] );   // Jz: W(i=aa,j=bb)
                jbox_t *qx = xb + 0*T::JBOX + ( 1*T::JY + aa )*T::JZ + bb;      // flux points lo0+1..
                jbox_t *qy = xb + 1*T::JBOX + ( aa*T::JY + 1 )*T::JZ + bb;      // flux points lo1+1..
                jbox_t *qz = xb + 2*T::JBOX + ( aa*T::JY + bb )*T::JZ + 1;      // flux points lo2+1..
#pragma unroll
                for( int f=0; f<NW; f++ ) {
                    jadd_scaled( qx + f*T::JY*T::JZ, Cx[f]*Wx );
                    jadd_scaled( qy + f*T::JZ, Cy[f]*Wy );
                    jadd_scaled( qz + f, Cz[f]*Wz );
                }
            }
        }
    }
    __syncwarp();
}

template<int ORDER, int PUSHER, bool SCRATCH>
__global__ void __launch_bounds__( DYN_THREADS, CG<ORDER>::MINB ) k_dynamics_cg( const GridDev g, const DynArgs a, const __grid_constant__ FieldMaps tm )
{
    using D = CGDim<ORDER>;
    using T = typename D::T;
    constexpr int NW = D::NW, NCELL = D::NCELL, CPT = D::CPT, NV = D::NV, NS = D::NS, NSL = D::NSL, NPASS = D::NPASS;
    constexpr bool COMPACT = ORDER == 4;      // loops over components / flux passes kept rolled: code size (see the deposit)
    static_assert( !COMPACT || NSL == 1, "the compact deposit advances one flux point per pass" );
    extern __shared__ __align__( 128 ) double smem[];
    double *sF = smem;
    jbox_t *sJ = reinterpret_cast<jbox_t *>( smem + 6*T::FBOX );
    double *xscr = smem + 6*T::FBOX + 3*T::JBOX + D::XSCR*( threadIdx.x/GRP );      // this lane group's crosser scratch
    __shared__ __align__( 8 ) unsigned long long tma_bar;
    __shared__ int cell_first[NCELL];
    __shared__ int cell_cnt[NCELL];
    __shared__ int round_off[NCELL+1];
    __shared__ int warp_tot[DYN_THREADS/32];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    int b = blockIdx.x;
    const int tz = b % a.tiles[2]; b /= a.tiles[2];
    const int ty = b % a.tiles[1];
    const int tx = b / a.tiles[1];
    const int c0[3] = { tx*T::TX, ty*T::TY, tz*T::TZ };

    // the boxes start at an even z index (16-B aligned rows): one element early when the natural start is odd
    const int zs = ( c0[2] + g.o[2] - T::H ) & 1;          // field boxes
    const int zj = ( c0[2] + g.o[2] - T::H - 1 ) & 1;      // J box
    // ---- TMA: one elected thread asks for the six E/B_m stencil boxes of the tile (box index s <-> array
    //      index c0 + o - H + s); they land while the CTA sets up its cell runs
    if( tid == 0 ) tma_bar_init( &tma_bar, 1 );
    __syncthreads();
    if( tid == 0 ) {
        tma_expect( &tma_bar, D::TMA_BYTES );
#pragma unroll
        for( int c=0; c<6; c++ )
            tma_load_3d( sF + c*T::FBOX, &tm.m[c], &tma_bar, c0[2] + g.o[2] - T::H - zs, c0[1] + g.o[1] - T::H, c0[0] + g.o[0] - T::H );
    }

    // ---- per-cell particle runs and the prefix sum of their rounds
    {
        int run = 0;
        int rounds[CPT];
#pragma unroll
        for( int e=0; e<CPT; e++ ) {
            const int ct = tid*CPT + e;
            int beg = 0, cnt = 0;
            if( ct < NCELL ) {
                const int lz = ct % T::TZ, ly = ( ct / T::TZ ) % T::TY, lx = ct / ( T::TZ*T::TY );
                const int ix = c0[0]+lx, iy = c0[1]+ly, iz = c0[2]+lz;
                if( ix < g.ncell[0] && iy < g.ncell[1] && iz < g.ncell[2] ) {
                    const int cell = ( ix*g.ncell[1] + iy )*g.ncell[2] + iz;
                    beg = a.first[cell];
                    cnt = a.first[cell+1] - beg;
                }
                cell_first[ct] = beg;
                cell_cnt[ct] = cnt;
            }
            rounds[e] = ( cnt + GRP - 1 )/GRP;
            run += rounds[e];
        }
        int inc = run;
#pragma unroll
        for( int d=1; d<32; d<<=1 ) { const int u = __shfl_up_sync( 0xffffffffu, inc, d ); if( lane >= d ) inc += u; }
        if( lane == 31 ) warp_tot[tid >> 5] = inc;
        __syncthreads();
        int off = inc - run;
        for( int w=0; w<( tid >> 5 ); w++ ) off += warp_tot[w];
#pragma unroll
        for( int e=0; e<CPT; e++ ) {
            off += rounds[e];
            if( tid*CPT + e < NCELL ) round_off[tid*CPT + e + 1] = off;
        }
        if( tid == 0 ) round_off[0] = 0;
    }
    __syncthreads();
    const int nrounds = round_off[NCELL];
    if( nrounds == 0 ) {
        tma_wait( &tma_bar, 0 );      // the boxes must have landed before this CTA's shared memory is released
        return;
    }

    // ---- the six field boxes arrive by TMA (issued before the per-cell set-up, see above); clear the J box
    //      meanwhile and wait for the boxes.  Out-of-range box elements are zero-filled by the TMA unit.
    for( int t = tid; t < 3*T::JBOX; t += DYN_THREADS ) sJ[t] = 0ull;
    tma_wait( &tma_bar, 0 );
    __syncthreads();

    // ---- lane geometry of the transpose-reduction: which of the NV values each of this lane's NS final
    //      slots holds, as an offset in the J box relative to the cell base (for the first flux pass)
    const int gl = tid & ( GRP-1 );
    const int gid = tid / GRP;
    const bool up4 = gl & 4, up2 = gl & 2, up1 = gl & 1;
    int joff[3][NS];
#pragma unroll
    for( int r=0; r<NS; r++ ) {
        int idx = r;
        bool ok = true;
        idx += up1 ? NS : 0;     ok = ok && idx < D::N2;
        idx += up2 ? D::N2 : 0;  ok = ok && idx < D::N1;
        idx += up4 ? D::N1 : 0;  ok = ok && idx < NV;
        const int f = idx / ( NW*NW ), aa = ( idx / NW ) % NW, bb = idx % NW;
        // Jx: flux i = 2+f, (j,k) = (1+aa, 1+bb);  Jy: flux j = 2+f, (i,k) = (1+aa, 1+bb);  Jz: flux k = 2+f, (i,j) = (1+aa, 1+bb)
        joff[0][r] = ok ? 0*T::JBOX + ( ( 2+f )*T::JY + ( 1+aa ) )*T::JZ + ( 1+bb ) : -1;
        joff[1][r] = ok ? 1*T::JBOX + ( ( 1+aa )*T::JY + ( 2+f ) )*T::JZ + ( 1+bb ) : -1;
        joff[2][r] = ok ? 2*T::JBOX + ( ( 1+aa )*T::JY + ( 1+bb ) )*T::JZ + ( 2+f ) : -1;
    }
    const int fstride[3] = { NSL*T::JY*T::JZ, NSL*T::JZ, NSL };      // J-box stride of one flux pass per component

    double *xq = smem + 6*T::FBOX + 3*T::JBOX + D::XSCR*( DYN_THREADS/GRP ) + XQD*( tid >> 5 );   // this warp's queue
    int *xqm = reinterpret_cast<int *>( xq + 9*XQ );
    int qh = 0, qn = 0;                                       // queue head / pending entries (warp-uniform)

    constexpr int NGROUPS = DYN_THREADS/GRP;
    const int niter = ( nrounds + NGROUPS - 1 )/NGROUPS;
    for( int it = 0; it < niter; it++ ) {
        const int wi = it*NGROUPS + gid;
        const bool have = wi < nrounds;
        // cell of this round: last c with round_off[c] <= wi
        int lo = 0, hi = NCELL;
        const int wq = have ? wi : 0;
#pragma unroll
        for( int st=0; st<CG<ORDER>::LOG2CELLS; st++ ) { const int mid = ( lo+hi ) >> 1; if( round_off[mid] <= wq ) lo = mid; else hi = mid; }
        const int cellt = lo;
        const int cl[3] = { cellt / ( T::TZ*T::TY ), ( cellt / T::TZ ) % T::TY, cellt % T::TZ };
        const int slot = ( wq - round_off[cellt] )*GRP + gl;
        const bool active = have && slot < cell_cnt[cellt];
        const size_t ip = ( size_t )cell_first[cellt] + ( size_t )( active ? slot : 0 );

        double S0[3][NW], DS[3][NW], cr[3] = { 0., 0., 0. }, xdelta[3] = { 0., 0., 0. }, xnpos[3] = { 0., 0., 0. };
        int shifts = 0;                       // (shift+1) per dimension, 2 bits each
        bool fast = false;
        if( active ) {
            const size_t is = a.perm ? ( size_t )a.perm[ip] : ip;       // deferred gather of the sort (sort.cu)
            double pos[3] = { a.in[0][is], a.in[1][is], a.in[2][is] };
            double px = a.in[3][is], py = a.in[4][is], pz = a.in[5][is];
            const double weight = a.in[6][is];
            const short charge = a.qin[is];
            if( a.perm ) { a.col[6][ip] = weight; a.q[ip] = charge; }

            double cd[3][NW];
            int sp[3], sd[3];
#pragma unroll
            for( int d=0; d<3; d++ ) {
                const double pn = pos[d]*g.dxi[d];
                const int ipn = ( int )round( pn );
                xdelta[d] = pn - ( double )ipn;
                Shape<ORDER>::w( xdelta[d], S0[d] );          // primal coefficients = S0 of the deposit
                const int idn = ( int )round( pn + 0.5 );
                const double dd = pn - ( double )idn + 0.5;
                Shape<ORDER>::w( dd, cd[d] );
                if( ipn - g.begin[d] - g.o[d] - c0[d] != cl[d] ) atomicAdd( &a.iflags[1], 1 );
                sp[d] = cl[d] + T::H + ( d == 2 ? zs : 0 );
                sd[d] = sp[d] + ( idn - ipn );
            }
            double Ex, Ey, Ez, Bx, By, Bz;
            if( COMPACT ) {
                // one loop over the six components instead of six inlined gathers (750 FMA + 750 loads at order 4):
                // the weights of a component are picked per dimension (dual along its own direction for E, along
                // the two others for B), ElectroMagn3D.cpp:115-123
                double EB[6];
#pragma unroll 1
                for( int c=0; c<6; c++ ) {
                    const bool ux = c < 3 ? c == 0 : c != 3, uy = c < 3 ? c == 1 : c != 4, uz = c < 3 ? c == 2 : c != 5;
                    double wx[NW], wy[NW], wz[NW];
#pragma unroll
                    for( int s=0; s<NW; s++ ) {
                        wx[s] = ux ? cd[0][s] : S0[0][s];
                        wy[s] = uy ? cd[1][s] : S0[1][s];
                        wz[s] = uz ? cd[2][s] : S0[2][s];
                    }
                    EB[c] = gather<T>( sF + c*T::FBOX, wx, wy, wz, ux ? sd[0] : sp[0], uy ? sd[1] : sp[1], uz ? sd[2] : sp[2] );
                }
                Ex = EB[0]; Ey = EB[1]; Ez = EB[2]; Bx = EB[3]; By = EB[4]; Bz = EB[5];
            } else {
                Ex = gather<T>( sF+0*T::FBOX, cd[0], S0[1], S0[2], sd[0], sp[1], sp[2] );
                Ey = gather<T>( sF+1*T::FBOX, S0[0], cd[1], S0[2], sp[0], sd[1], sp[2] );
                Ez = gather<T>( sF+2*T::FBOX, S0[0], S0[1], cd[2], sp[0], sp[1], sd[2] );
                Bx = gather<T>( sF+3*T::FBOX, S0[0], cd[1], cd[2], sp[0], sd[1], sd[2] );
                By = gather<T>( sF+4*T::FBOX, cd[0], S0[1], cd[2], sd[0], sp[1], sd[2] );
                Bz = gather<T>( sF+5*T::FBOX, cd[0], cd[1], S0[2], sd[0], sd[1], sp[2] );
            }

            const double cmd = ( double )charge*a.one_over_mass*g.dts2;
            double dxp, dyp, dzp, invgf;
            push<PUSHER>( cmd, g.dt, px, py, pz, Ex, Ey, Ez, Bx, By, Bz, dxp, dyp, dzp, invgf );
            const double npos[3] = { pos[0] + dxp, pos[1] + dyp, pos[2] + dzp };
            a.col[0][ip] = npos[0]; a.col[1][ip] = npos[1]; a.col[2][ip] = npos[2];
            a.col[3][ip] = px; a.col[4][ip] = py; a.col[5][ip] = pz;
            if( SCRATCH ) {
                a.sc_E[0*a.n+ip] = Ex; a.sc_E[1*a.n+ip] = Ey; a.sc_E[2*a.n+ip] = Ez;
                a.sc_B[0*a.n+ip] = Bx; a.sc_B[1*a.n+ip] = By; a.sc_B[2*a.n+ip] = Bz;
                a.sc_invgf[ip] = invgf;
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    a.sc_iold[d*a.n+ip] = cl[d] + c0[d] + g.o[d];
                    a.sc_delta[d*a.n+ip] = xdelta[d];
                }
            }

            // new shape factors, tag / next key
            int nkey[3];
            bool same = true, removed;
            const int tag = boundary_tag( a, g, npos, px, py, pz, weight, ip, removed );
#pragma unroll
            for( int d=0; d<3; d++ ) {
                const double pn = npos[d]*g.dxi[d];
                const int ipn = ( int )round( pn );
                double w1[NW];
                Shape<ORDER>::w( pn - ( double )ipn, w1 );
                const int shift = ipn - g.begin[d] - ( cl[d] + c0[d] + g.o[d] );
                shifts |= ( shift+1 ) << ( 2*d );
                same = same && shift == 0;
                nkey[d] = ( int )( ( double )ipn - g.min_loc_round[d] );
                xnpos[d] = pn;
#pragma unroll
                for( int s=0; s<NW; s++ ) DS[d][s] = w1[s] - S0[d][s];
            }
            int key = tag;
            if( tag == 0 ) { key = ( nkey[0]*g.ncell[1] + nkey[1] )*g.ncell[2] + nkey[2]; atomicAdd( &a.count[key], 1 ); }
            else if( tag < -1 ) note_leaver( a, tag, ip );
            a.key[ip] = key;

            const double charge_weight = removed ? 0. : g.inv_cell_volume*( double )charge*weight;
            cr[0] = charge_weight*g.d_ov_dt[0]; cr[1] = charge_weight*g.d_ov_dt[1]; cr[2] = charge_weight*g.d_ov_dt[2];
            fast = same;
        }
        jbox_t *jb = sJ + ( cl[0]*T::JY + cl[1] )*T::JZ + cl[2] + zj;
        // two rounds of the same cell often sit in neighbouring groups of the warp: their sums are added up
        // below instead of letting their adds collide on identical addresses
        const int mycell = have ? cellt : -1 - gid;
        const bool same8 = __shfl_xor_sync( 0xffffffffu, mycell, 8 ) == mycell;
        const bool own8 = have && !( same8 && ( lane & 8 ) );
        // every lane must execute both shuffles: no short-circuit evaluation around them
        const int c16 = __shfl_xor_sync( 0xffffffffu, mycell, 16 );
        const int o16own = __shfl_xor_sync( 0xffffffffu, ( int )own8, 16 );
        const bool same16 = c16 == mycell && own8 && o16own != 0;
        const bool owner = own8 && !( same16 && ( lane & 16 ) );
        // ---- non-crossing particles: (NW-1) x NW x NW values per component, NV at a time, summed over the 8
        //      lanes of the cell group in registers (all 32 lanes take part in the shuffles)
        if( COMPACT ) {
            // the same arithmetic with the component loop and the flux-pass loop NOT unrolled (the unrolled body
            // is 12 copies of ~330 instructions at order 4: the kernel did not fit the instruction cache)
#pragma unroll 1
            for( int c=0; c<3; c++ ) {
                const double third = 1./3.;
                double Sa[NW], Da[NW], A[NW], B[NW], Dc[NW-1];
#pragma unroll
                for( int k=0; k<NW; k++ ) {
                    Sa[k] = c == 0 ? S0[1][k] : S0[0][k];                  // slow transverse dimension: y for Jx, x otherwise
                    Da[k] = c == 0 ? DS[1][k] : DS[0][k];
                    const double sb = c == 2 ? S0[1][k] : S0[2][k];       // fast transverse dimension: y for Jz, z otherwise
                    const double db_ = c == 2 ? DS[1][k] : DS[2][k];
                    A[k] = sb + 0.5*db_; B[k] = 0.5*sb + third*db_;
                }
#pragma unroll
                for( int f=0; f<NW-1; f++ ) Dc[f] = c == 0 ? DS[0][f] : ( c == 1 ? DS[1][f] : DS[2][f] );
                const double crc = c == 0 ? cr[0] : ( c == 1 ? cr[1] : cr[2] );
                const int fs = c == 0 ? fstride[0] : ( c == 1 ? fstride[1] : fstride[2] );
                int jo[NS];
#pragma unroll
                for( int r=0; r<NS; r++ ) jo[r] = c == 0 ? joff[0][r] : ( c == 1 ? joff[1][r] : joff[2][r] );
                double run = 0.;
#pragma unroll 1
                for( int ps=0; ps<NPASS; ps++ ) {
                    run -= crc*Dc[0];                                      // flux coefficient of this pass (NSL == 1)
#pragma unroll
                    for( int f=0; f<NW-2; f++ ) Dc[f] = Dc[f+1];
                    double v[NV];
#pragma unroll
                    for( int j=0; j<NW; j++ )
#pragma unroll
                        for( int k=0; k<NW; k++ )
                            v[j*NW + k] = fast ? run*( Sa[j]*A[k] + Da[j]*B[k] ) : 0.;
                    xr_step<NV>( v, 4, up4 ); xr_step<D::N1>( v, 2, up2 ); xr_step<D::N2>( v, 1, up1 );
#pragma unroll
                    for( int r=0; r<NS; r++ ) {
                        const double o8 = __shfl_xor_sync( 0xffffffffu, v[r], 8 );
                        if( same8 ) v[r] += o8;
                        const double o16 = __shfl_xor_sync( 0xffffffffu, v[r], 16 );
                        if( same16 ) v[r] += o16;
                        if( owner && jo[r] >= 0 && v[r] != 0. ) jadd( jb + jo[r] + ps*fs, v[r], a.jscale );
                    }
                }
            }
        } else {
#pragma unroll
        for( int c=0; c<3; c++ ) {
            const int da = c == 0 ? 1 : 0, db = c == 2 ? 1 : 2;      // transverse dimensions (slow, fast)
            const double third = 1./3.;
            double A[NW], B[NW], Cf[NW-1];
            if( fast ) {
#pragma unroll
                for( int k=0; k<NW; k++ ) { A[k] = S0[db][k] + 0.5*DS[db][k]; B[k] = 0.5*S0[db][k] + third*DS[db][k]; }
                double run = 0.;
#pragma unroll
                for( int f=0; f<NW-1; f++ ) { run -= cr[c]*DS[c][f]; Cf[f] = run; }
            }
#pragma unroll
            for( int ps=0; ps<NPASS; ps++ ) {
                double v[NV];
                if( fast ) {
#pragma unroll
                    for( int fl=0; fl<NSL; fl++ )
#pragma unroll
                        for( int j=0; j<NW; j++ )
#pragma unroll
                            for( int k=0; k<NW; k++ )
                                v[( fl*NW + j )*NW + k] = Cf[ps*NSL+fl]*( S0[da][j]*A[k] + DS[da][j]*B[k] );
                } else {
#pragma unroll
                    for( int i=0; i<NV; i++ ) v[i] = 0.;
                }
                xr_step<NV>( v, 4, up4 ); xr_step<D::N1>( v, 2, up2 ); xr_step<D::N2>( v, 1, up1 );
#pragma unroll
                for( int r=0; r<NS; r++ ) {
                    const double o8 = __shfl_xor_sync( 0xffffffffu, v[r], 8 );
                    if( same8 ) v[r] += o8;
                    const double o16 = __shfl_xor_sync( 0xffffffffu, v[r], 16 );
                    if( same16 ) v[r] += o16;
                    if( owner && joff[c][r] >= 0 && v[r] != 0. ) jadd( jb + joff[c][r] + ps*fstride[c], v[r], a.jscale );
                }
            }
        }
        }
        // ---- particles that changed cell (Projector3D2Order.cpp:124-340 with ip_m_ipo != 0) go to the warp's
        //      queue; whenever 4 are pending the warp deposits them, one per lane group (cross_pass)
        unsigned xmask = __ballot_sync( 0xffffffffu, active && !fast );
        while( xmask ) {                                                  // warp-uniform
            const int room = XQ - qn;
            const int rank = __popc( xmask & ( ( 1u << lane ) - 1u ) );
            const bool mineq = ( ( xmask >> lane ) & 1u ) && rank < room;
            if( mineq ) {
                const int e = ( qh + qn + rank ) % XQ;
#pragma unroll
                for( int d=0; d<3; d++ ) { xq[( 0+d )*XQ+e] = xdelta[d]; xq[( 3+d )*XQ+e] = xnpos[d]; xq[( 6+d )*XQ+e] = cr[d]; }
                xqm[e] = cellt | ( shifts << 16 );
            }
            const unsigned done = __ballot_sync( 0xffffffffu, mineq );
            xmask &= ~done;
            qn += __popc( done );
            __syncwarp();
            while( qn >= 4 || ( xmask && qn > 0 ) ) {
                cross_pass<ORDER>( sJ + zj, xq, xqm, xscr, qh, qn, gl, lane >> 3, a.jscale );
                const int took = qn < 4 ? qn : 4;
                qh = ( qh + took ) % XQ;
                qn -= took;
            }
        }
    }
    while( qn > 0 ) {                                                     // drain what is left in the queue
        cross_pass<ORDER>( sJ + zj, xq, xqm, xscr, qh, qn, gl, lane >> 3, a.jscale );
        const int took = qn < 4 ? qn : 4;
        qh = ( qh + took ) % XQ;
        qn -= took;
    }
    __syncthreads();

    // ---- flush the J box: convert the fixed-point sums to double in place, then ONE elected thread hands the three
    //      boxes to the TMA unit, which adds them into Jx, Jy, Jz in HBM (cp.reduce.async.bulk.tensor .add)
    for( int t = tid; t < 3*T::JBOX; t += DYN_THREADS ) {
        const long long iv = ( long long )sJ[t];
        reinterpret_cast<double *>( sJ )[t] = ( double )iv*a.jinv;
    }
    tma_store_fence();            // make the generic-proxy writes visible to the async proxy
    __syncthreads();
    if( tid == 0 ) {
#pragma unroll
        for( int c=0; c<3; c++ )
            tma_reduce_add_3d( &tm.j[c], sJ + c*T::JBOX, c0[2] + g.o[2] - T::H - 1 - zj, c0[1] + g.o[1] - T::H - 1, c0[0] + g.o[0] - T::H - 1 );
        tma_commit_and_wait_read();      // shared memory may be released once the TMA unit has read the boxes
    }
}


// =================================================================================================
// Order-2 production kernel: a particle STREAM walked by producer warps, the per-cell
// current sums formed by a consumer warp.
//
// One CTA (320 threads) owns a tile of 4 x 4 x 8 primal-node cells.  The tile's 16 z-rows of 8 cells are split
// between two GROUPS (x-slabs {0,1} and {2,3}); a group is 4 producer warps + 1 consumer warp.  Because the
// particles are sorted by cell key with z fastest, the particles of a row are one contiguous run; the group's
// stream is its 8 runs one after the other, and producer lane L of round r takes stream item 128 r + L: every
// lane holds a particle whatever the cell occupancies are (no lock-step on the fullest cell of a warp), and a
// warp's loads and stores are contiguous.
//   producer  lane = one particle: gather, push, tag, key, new shape.  What the deposit needs — per dimension
//             M = (S0+S1)/2 and DS/sqrt(12) on the 3 home nodes, the two flux coefficients per component
//             (already in fixed-point units) — goes to a 26-double record in shared memory;
//   consumer  lane = one current COMPONENT of one cell of the row being streamed (8 cells x 3 components): it
//             walks the records of its cell in the window and accumulates its 2 x 3 x 3 values in registers,
//                 J_c[f][j][k] += Cf_c[f] * ( M_a[j] M_b[k] + DS_a[j] DS_b[k] / 12 )
//             (the Esirkepov weight S0S0 + (DS S0 + S0 DS)/2 + DS DS/3 around the mid-point shape); when its
//             cell is exhausted the 18 sums go to the tile's J box (fixed-point adds) and the lane moves to
//             the same z of the next row.
// Producers and consumer hand the record buffer back and forth with two named barriers per group (FULL /
// EMPTY); the buffer is single: the consumer reads round r while the producers gather and push round r+1,
// and they only wait for EMPTY right before they write the records of round r+1.
// A particle that moved to the next node in ONE dimension still deposits its home part through its record;
// the 21 values that fall outside the home window are added by an 8-lane octet of its own warp, driven by a
// table (value = C * (P*B0 + Q*B1), five record slots and a J-box offset per entry).  Particles that moved
// in 2 or 3 dimensions stage zero flux coefficients and take the warp's queue (cross_pass).
// =================================================================================================
namespace o2 {
using T = CG<2>::T;
constexpr int NW = 3;
constexpr int NPROD = 256;                               // producer threads: 2 groups x 4 warps
constexpr int NTHR = NPROD + 64;                         // + one consumer warp per group
constexpr int GROUP = 128;                               // stream items per window = producer lanes of a group
constexpr int GTHR = GROUP + 32;                         // threads meeting at a group's named barriers
constexpr int NCELL = T::TX*T::TY*T::TZ;                 // 128 cells per tile
constexpr int GCELLS = NCELL/2;                          // cells per group (x-slabs {0,1} / {2,3})
constexpr int ROWS = GCELLS/T::TZ;                       // 8 rows per group
static_assert( T::TX == 4 && T::TY == 4 && T::TZ == 8, "the row / group arithmetic below is written for 4 x 4 x 8 tiles" );
constexpr int REC = 26;                                  // doubles per record: 18 M/DS + 6 Cf + outer weight + outer flux coefficient
// slot L of a group's record buffer; one 16-byte skew every 16 slots: with exactly 16 particles per cell the 8
// lane quads of the consumer read slots 16 apart, which would otherwise all sit in the same banks
__device__ __forceinline__ int rec_off( int L ) { return L*REC + 2*( L >> 4 ); }
__device__ __forceinline__ unsigned rec_byte( int L ) { return ( unsigned )( L*( REC*8 ) + ( ( L >> 4 ) << 4 ) ); }     // 8*rec_off(L)
// predicated 16-byte shared-memory load (no branch: the compiler keeps it where it is written; a lane that is switched
// off keeps the old value and asks nothing of the shared-memory pipe)
__device__ __forceinline__ void lds_v2( double2 &v, unsigned addr, int on )
{
    asm volatile( "{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p ld.shared.v2.f64 {%0, %1}, [%2];\n}" : "+d"( v.x ), "+d"( v.y ) : "r"( addr ), "r"( on ) );
}
constexpr int RECBUF = GROUP*REC + 2*( GROUP/16 );       // doubles per group
constexpr int DIRECT_PPC = 6;                            // a group with fewer particles per cell on average deposits lane by lane (see the kernel)
constexpr int XQ2 = 8;                                   // crosser queue entries per producer warp
constexpr int XQD2 = 9*XQ2 + XQ2/2;
constexpr int XSCR2 = 4*CGDim<2>::XSCR;                  // cross_pass scratch per producer warp (4 octets)
constexpr int XTAB = 6*24;                               // outer-part table: (dimension, direction) x 24 items, one int2 each
constexpr double K12 = 0.28867513459481288225;           // 1/sqrt(12)
constexpr size_t BYTES = ( size_t )( 6*T::FBOX + 3*T::JBOX + 2*RECBUF + 8*XQD2 + 8*XSCR2 + XTAB )*sizeof( double );
constexpr unsigned TMA_BYTES = 6u*T::FVOL*sizeof( double );

__device__ __forceinline__ void mbar_arrive( unsigned long long *bar ) { asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"( smem_u32( bar ) ) : "memory" ); }

// The windows of a group's stream, computed identically by every warp of the group: as many WHOLE cells as fit in
// GROUP items, at most 8 (one per lane quad of the consumer, which then walks each cell in one go and never carries
// a cell over a window boundary); a cell of more than GROUP items is taken in pieces.  ws/q0: first item / first cell
// of the window; returns false when the stream is exhausted, else we (end item) and q1 (end cell; == q0 for a piece).
__device__ __forceinline__ bool next_window( const int *coff, int total, int lane, int &ws, int &q0, int &we, int &q1 )
{
    while( ws < total ) {
        const int qq = min( q0 + 1 + ( lane & 7 ), GCELLS );            // candidate ends q0+1 .. q0+8
        const int nfit = __popc( __ballot_sync( 0xffffffffu, coff[qq] - ws <= GROUP ) & 0xffu );
        if( nfit == 0 ) { q1 = q0; we = ws + GROUP; return true; }        // a piece of a very full cell
        q1 = min( q0 + nfit, GCELLS );
        we = coff[q1];
        if( we > ws ) return true;
        q0 = q1;                                                          // only empty cells: skip them
    }
    return false;
}

// Interpolator3D2Order.h:107-123 in 5 operations: c2 - c0 = d exactly
__device__ __forceinline__ void shape2( double d, double *c )
{
    const double t = fma( d, d, 0.25 );
    c[0] = 0.5*( t - d );
    c[1] = fma( -d, d, 0.75 );
    c[2] = c[0] + d;
}

// entry `it` (0..23) of the outer-part table for a particle that moved to the next node along dimension d,
// upwards (up = 1) or downwards.  Items 0..8: the flux component d at the flux point outside the home window,
// on the 3 x 3 home nodes of the two other dimensions; items 9..20: 2 flux points x 3 nodes of each of the two
// other components on the node plane outside the home window.  value = rec[iC] * ( cP rec[iP] rec[iB] + cQ rec[iQ] rec[iB+3] ).
__device__ __forceinline__ int2 xtab_entry( int m, int it )
{
    const int d = m % 3, up = m / 3;
    const int a1 = d == 0 ? 1 : 0, a2 = d == 2 ? 1 : 2;
    const int st[3] = { T::JY*T::JZ, T::JZ, 1 };
    int comp = 0, iC = 0, iP = 0, iQ = 0, iB = 0, isb = 0, off = 0, valid = 1;
    if( it < 9 ) {
        const int j = it/3, kk = it - 3*j;
        comp = d; iC = 25; iP = 6*a1 + j; iQ = iP + 3; iB = 6*a2 + kk;
        off = ( up ? 4 : 1 )*st[d] + ( 1+j )*st[a1] + ( 1+kk )*st[a2];
    } else if( it < 21 ) {
        const int t = it - 9, second = t >= 6 ? 1 : 0, u = t - 6*second, f = u/3, kk = u - 3*f;
        const int bdim = second ? a1 : a2;
        comp = second ? a2 : a1;
        iC = 18 + 2*comp + f; iP = 24; iQ = 24; iB = 6*bdim + kk; isb = 1;
        off = ( 2+f )*st[comp] + ( up ? 4 : 0 )*st[d] + ( 1+kk )*st[bdim];
    } else valid = 0;
    return make_int2( iC | ( iP << 5 ) | ( iQ << 10 ) | ( iB << 15 ) | ( isb << 20 ) | ( valid << 21 ), comp*T::JBOX + off );
}
// sparse group (see k_dynamics_o2): the lane that wrote a record adds its sums to the J box itself, one current component
// after the other.  Kept out of line: it is the rare path and must not weigh on the register allocation of the main one.
__device__ __noinline__ void self_consume( const double *rec, jbox_t *jcell )
{
#pragma unroll 1
    for( int cc=0; cc<3; cc++ ) {
        const int da = cc == 0 ? 1 : 0, db = cc == 2 ? 1 : 2;
        const int sf_ = cc == 0 ? T::JY*T::JZ : ( cc == 1 ? T::JZ : 1 );
        const int sa_ = cc == 0 ? T::JZ : T::JY*T::JZ;
        const int sb_ = cc == 2 ? T::JZ : 1;
        const double2 *qa = reinterpret_cast<const double2 *>( rec + 6*da );
        const double2 *qb = reinterpret_cast<const double2 *>( rec + 6*db );
        const double2 a0 = qa[0], a1 = qa[1], a2 = qa[2], b0 = qb[0], b1 = qb[1], b2 = qb[2];
        const double2 cf = *reinterpret_cast<const double2 *>( rec + 18 + 2*cc );
        const double Ma[NW] = { a0.x, a0.y, a1.x }, Da[NW] = { a1.y, a2.x, a2.y };
        const double Mb[NW] = { b0.x, b0.y, b1.x }, Db[NW] = { b1.y, b2.x, b2.y };
        jbox_t *jb = jcell + cc*T::JBOX + 2*sf_ + sa_ + sb_;
        if( cf.x != 0. || cf.y != 0. ) {
#pragma unroll
            for( int j=0; j<NW; j++ )
#pragma unroll
                for( int k_=0; k_<NW; k_++ ) {
                    const double W = fma( Ma[j], Mb[k_], Da[j]*Db[k_] );
                    jadd_scaled( jb + j*sa_ + k_*sb_, cf.x*W );
                    jadd_scaled( jb + sf_ + j*sa_ + k_*sb_, cf.y*W );
                }
        }
    }
}
}

template<int PUSHER, bool SCRATCH, bool REMOVE>
__global__ void __launch_bounds__( o2::NTHR, 2 ) k_dynamics_o2( const GridDev g, const DynArgs a, const __grid_constant__ FieldMaps tm )
{
    using namespace o2;
    extern __shared__ __align__( 128 ) double smem[];
    double *sF = smem;
    jbox_t *sJ = reinterpret_cast<jbox_t *>( smem + 6*T::FBOX );
    double *recbase = smem + 6*T::FBOX + 3*T::JBOX;
    double *xqbase = recbase + 2*RECBUF;
    double *xscrbase = xqbase + 8*XQD2;
    int2 *xtab = reinterpret_cast<int2 *>( xscrbase + 8*XSCR2 );
    __shared__ __align__( 8 ) unsigned long long tma_bar;
    __shared__ int cell_first[NCELL];
    __shared__ int cell_cnt[NCELL];
    __shared__ int cell_off[2][GCELLS+1];                // start of each cell in its group's stream
    __shared__ unsigned short xsrc[8][32];               // per producer warp: the lanes holding a single-dimension mover
    __shared__ __align__( 8 ) unsigned long long full_bar[2], empty_bar[2];   // per group: records written (4 producer warps) / records consumed

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    int b = blockIdx.x;
    const int tz = b % a.tiles[2]; b /= a.tiles[2];
    const int ty = b % a.tiles[1];
    const int tx = b / a.tiles[1];
    const int c0[3] = { tx*T::TX, ty*T::TY, tz*T::TZ };
    const int zs = ( c0[2] + g.o[2] - T::H ) & 1;          // field boxes start on an even z index
    const int zj = ( c0[2] + g.o[2] - T::H - 1 ) & 1;      // J box too

    if( tid == 0 ) {
        tma_bar_init( &tma_bar, 1 );
        tma_bar_init( &full_bar[0], 4 ); tma_bar_init( &full_bar[1], 4 );
        tma_bar_init( &empty_bar[0], 1 ); tma_bar_init( &empty_bar[1], 1 );
    }
    __syncthreads();
    if( tid == 0 ) {
        tma_expect( &tma_bar, TMA_BYTES );
#pragma unroll
        for( int c=0; c<6; c++ )
            tma_load_3d( sF + c*T::FBOX, &tm.m[c], &tma_bar, c0[2] + g.o[2] - T::H - zs, c0[1] + g.o[1] - T::H, c0[0] + g.o[0] - T::H );
    }
    int mine = 0;
    if( tid < NCELL ) {
        const int lz = tid % T::TZ, ly = ( tid / T::TZ ) % T::TY, lx = tid / ( T::TZ*T::TY );
        const int ix = c0[0]+lx, iy = c0[1]+ly, iz = c0[2]+lz;
        int beg = 0, cnt = 0;
        if( ix < g.ncell[0] && iy < g.ncell[1] && iz < g.ncell[2] ) {
            const int cell = ( ix*g.ncell[1] + iy )*g.ncell[2] + iz;
            beg = a.first[cell];
            cnt = a.first[cell+1] - beg;
        }
        cell_first[tid] = beg;
        cell_cnt[tid] = cnt;
        mine = cnt;
    } else if( tid - NCELL < XTAB ) {
        xtab[tid - NCELL] = xtab_entry( ( tid - NCELL )/24, ( tid - NCELL )%24 );
    }
    for( int t = tid; t < 3*T::JBOX; t += NTHR ) sJ[t] = 0ull;
    const int any = __syncthreads_or( mine );
    tma_wait( &tma_bar, 0 );          // the boxes must have landed before this CTA's shared memory is used or released
    if( !any ) return;
    // stream offsets of the cells: exclusive prefix sum of the 64 counts of each group (warp w < 2 scans group w)
    if( warp < 2 ) {
        const int a0 = cell_cnt[GCELLS*warp + 2*lane], a1 = cell_cnt[GCELLS*warp + 2*lane + 1];
        int inc = a0 + a1;
#pragma unroll
        for( int d=1; d<32; d<<=1 ) { const int u = __shfl_up_sync( 0xffffffffu, inc, d ); if( lane >= d ) inc += u; }
        cell_off[warp][2*lane] = inc - a0 - a1;
        cell_off[warp][2*lane+1] = inc - a1;
        if( lane == 31 ) cell_off[warp][GCELLS] = inc;
    }
    __syncthreads();

    if( warp < 8 ) {
        // ============================================================ producers
        const int grp = warp >> 2;
        const int L = tid & ( GROUP-1 );
        const int *coff = cell_off[grp];
        const int total = coff[GCELLS];
        double *recbuf = recbase + grp*RECBUF;
        double *rec = recbuf + rec_off( L );
        double *xq = xqbase + XQD2*warp;
        int *xqm = reinterpret_cast<int *>( xq + 9*XQ2 );
        double *xscr = xscrbase + XSCR2*warp + CGDim<2>::XSCR*( lane >> 3 );
        const int base[3] = { g.begin[0] + g.o[0] + c0[0], g.begin[1] + g.o[1] + c0[1], g.begin[2] + g.o[2] + c0[2] };
        int qh = 0, qn = 0, bad = 0;
        int ws = 0, q0 = 0, we = 0, q1 = 0;
        // SPARSE group (fewer than DIRECT_PPC particles per cell on average, e.g. the 1-per-cell plasma of a laser
        // wake): windows of whole cells would leave most lanes idle, so the windows are plain stretches of GROUP
        // particles whatever their cells, and every lane adds the sums of ITS OWN record to the J box (no consumer,
        // no hand-off; with so few particles per cell there is nothing to accumulate in registers anyway)
        const bool direct = total < DIRECT_PPC*GCELLS;
        auto window = [&]( int &ws_, int &q0_, int &we_, int &q1_ ) -> bool {
            if( !direct ) return next_window( coff, total, lane, ws_, q0_, we_, q1_ );
            we_ = min( ws_ + GROUP, total ); q1_ = q0_;
            return ws_ < total;
        };
        // slot (in the sorted order) and source index (through the pending sort order) of this lane's particle in a
        // window; looked up one window ahead so that the dependent load of the sort order is off the critical path
        auto locate = [&]( int ws_, int q0_, int we_, int &row_, int &ip_, int &is_ ) {
            const int s_ = ws_ + L;
            row_ = q0_ >> 3;                                            // the window's cells lie in at most two rows
            ip_ = -1; is_ = -1;
            if( s_ < we_ ) {
                if( direct ) {
                    row_ = 0;
#pragma unroll
                    for( int k_=1; k_<ROWS; k_++ ) row_ += s_ >= coff[T::TZ*k_];
                } else row_ += s_ >= coff[T::TZ*( row_+1 )];
                ip_ = cell_first[GCELLS*grp + T::TZ*row_] + ( s_ - coff[T::TZ*row_] );
                is_ = a.perm ? a.perm[ip_] : ip_;
            }
        };
        double pos[3] = { 0., 0., 0. };
        bool more = window( ws, q0, we, q1 );
        int row = 0, ipi = -1, isi = -1;
        if( more ) locate( ws, q0, we, row, ipi, isi );
        if( isi >= 0 ) { pos[0] = a.in[0][isi]; pos[1] = a.in[1][isi]; pos[2] = a.in[2][isi]; }

#pragma unroll 1
        for( int r = 0; more; r++ ) {
            int nws = we, nq0 = q1, nwe = 0, nq1 = 0, nrow = 0, nip = -1, nis = -1;
            const bool nmore = window( nws, nq0, nwe, nq1 );
            if( nmore ) locate( nws, nq0, nwe, nrow, nip, nis );
            const bool active = ipi >= 0;
            double S0[3][NW], dl1[3] = { 0., 0., 0. }, cr[3] = { 0., 0., 0. }, xdelta[3] = { 0., 0., 0. }, xnpos[3] = { 0., 0., 0. };
            int shifts = 0x15, cellt = 0, nx = 0;
            // ---------------- part A: gather, push, tag, key
            if( active ) {
                const size_t ip = ( size_t )ipi, is = ( size_t )isi;
                double cd[3][NW];
                int sp[3], sd[3], cl[3];
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    const double pn = __dmul_rn( pos[d], g.dxi[d] );       // no contraction into the subtraction below: deltaold is bit-exact
                    const int ipn = ( int )round( pn );
                    xdelta[d] = pn - ( double )ipn;
                    shape2( xdelta[d], S0[d] );
                    const int idn = ( int )round( pn + 0.5 );
                    shape2( pn - ( double )idn + 0.5, cd[d] );
                    cl[d] = ipn - base[d];
                    sd[d] = idn - ipn;
                }
                // the particle must sit in the row the stream says (its sort key): x and y of the row, z inside the tile
                {
                    const int lx = 2*grp + ( row >> 2 ), ly = row & 3;
                    if( cl[0] != lx || cl[1] != ly || ( unsigned )cl[2] >= ( unsigned )T::TZ ) {
                        bad++;
                        cl[0] = lx; cl[1] = ly; cl[2] = cl[2] < 0 ? 0 : ( cl[2] >= T::TZ ? T::TZ-1 : cl[2] );
                    }
                }
                cellt = ( cl[0]*T::TY + cl[1] )*T::TZ + cl[2];
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    sp[d] = cl[d] + T::H + ( d == 2 ? zs : 0 );
                    sd[d] += sp[d];
                }
                const double Ex = gather<T>( sF+0*T::FBOX, cd[0], S0[1], S0[2], sd[0], sp[1], sp[2] );
                const double Ey = gather<T>( sF+1*T::FBOX, S0[0], cd[1], S0[2], sp[0], sd[1], sp[2] );
                const double Ez = gather<T>( sF+2*T::FBOX, S0[0], S0[1], cd[2], sp[0], sp[1], sd[2] );
                // the momenta are not needed before the push: their loads are issued here, after the electric field, so that
                // they do not hold registers during the first half of the gather (the empty asm ties the index to Ez)
                size_t isl = is;
                asm volatile( "" : "+l"( isl ) : "d"( Ez ) );
                double px = a.in[3][isl], py = a.in[4][isl], pz = a.in[5][isl];
                const short charge = a.qin[isl];
                const double Bx = gather<T>( sF+3*T::FBOX, S0[0], cd[1], cd[2], sp[0], sd[1], sd[2] );
                const double By = gather<T>( sF+4*T::FBOX, cd[0], S0[1], cd[2], sd[0], sp[1], sd[2] );
                const double Bz = gather<T>( sF+5*T::FBOX, cd[0], cd[1], S0[2], sd[0], sd[1], sp[2] );

                const double cmd = ( double )charge*a.one_over_mass*g.dts2;
                double dxp, dyp, dzp, invgf;
                push<PUSHER>( cmd, g.dt, px, py, pz, Ex, Ey, Ez, Bx, By, Bz, dxp, dyp, dzp, invgf );
                const double npos[3] = { pos[0] + dxp, pos[1] + dyp, pos[2] + dzp };
                a.col[0][ip] = npos[0]; a.col[1][ip] = npos[1]; a.col[2][ip] = npos[2];
                a.col[3][ip] = px; a.col[4][ip] = py; a.col[5][ip] = pz;
                const double weight = a.in[6][is];
                if( a.perm ) { a.col[6][ip] = weight; a.q[ip] = charge; }
                if( SCRATCH ) {
                    a.sc_E[0*a.n+ip] = Ex; a.sc_E[1*a.n+ip] = Ey; a.sc_E[2*a.n+ip] = Ez;
                    a.sc_B[0*a.n+ip] = Bx; a.sc_B[1*a.n+ip] = By; a.sc_B[2*a.n+ip] = Bz;
                    a.sc_invgf[ip] = invgf;
#pragma unroll
                    for( int d=0; d<3; d++ ) {
                        a.sc_iold[d*a.n+ip] = cl[d] + c0[d] + g.o[d];
                        a.sc_delta[d*a.n+ip] = xdelta[d];
                    }
                }

                bool removed;
                const int tag = boundary_tag<REMOVE>( a, g, npos, px, py, pz, weight, ip, removed );
                const double charge_weight = removed ? 0. : g.inv_cell_volume*( double )charge*weight*a.jscale;
                cr[0] = charge_weight*g.d_ov_dt[0]; cr[1] = charge_weight*g.d_ov_dt[1]; cr[2] = charge_weight*g.d_ov_dt[2];

                int nkey[3];
                shifts = 0;
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    const double pn = __dmul_rn( npos[d], g.dxi[d] );
                    const int ipn = ( int )round( pn );
                    dl1[d] = pn - ( double )ipn;
                    const int shift = ipn - base[d] - cl[d];
                    shifts |= ( shift+1 ) << ( 2*d );
                    nx += shift != 0;
                    nkey[d] = ( int )( ( double )ipn - g.min_loc_round[d] );
                    xnpos[d] = pn;
                }
                int key = tag;
                if( tag == 0 ) { key = ( nkey[0]*g.ncell[1] + nkey[1] )*g.ncell[2] + nkey[2]; atomicAdd( &a.count[key], 1 ); }
                else if( tag < -1 ) atomicAdd( &a.leave_counts[-tag-2], 1 );
                a.key[ip] = key;
            }
            // ---------------- the positions of the next window's particle: in flight during the rest of this round
            if( nis >= 0 ) { pos[0] = a.in[0][nis]; pos[1] = a.in[1][nis]; pos[2] = a.in[2][nis]; }
            // ---------------- the consumer must have finished with the records of the previous round
            if( r > 0 && !direct ) tma_wait( &empty_bar[grp], ( r-1 ) & 1 );
            // ---------------- part B: new shape, the record of the deposit.  Per dimension M[3], DS[3]/sqrt(12) on
            //                  the HOME nodes (the 3 nodes of S0) and the flux coefficients at the 2 home flux
            //                  points.  A particle that moved to the next node along a dimension has S1 shifted by
            //                  one node: two of its three weights still fall on home nodes, the third (s1e) on the
            //                  node just outside; the flux running sum (Projector3D2Order.cpp:215-228) starts one
            //                  node earlier when the shift is negative.
            int xmeta = 0;
            if( __all_sync( 0xffffffffu, shifts == 0x15 ) ) {
                // no particle of the warp changed node (nearly every warp of a slow species): S1 sits on the home nodes,
                // none of the selections of the general case below
                if( active ) {
#pragma unroll
                    for( int d=0; d<3; d++ ) {
                        double w1[NW], s0[NW];
                        shape2( dl1[d], w1 );
                        shape2( xdelta[d], s0 );
                        const double ds0 = w1[0] - s0[0], ds1 = w1[1] - s0[1], ds2 = w1[2] - s0[2];
                        double2 *r2 = reinterpret_cast<double2 *>( rec + 6*d );
                        r2[0] = make_double2( fma( 0.5, ds0, s0[0] ), fma( 0.5, ds1, s0[1] ) );
                        r2[1] = make_double2( fma( 0.5, ds2, s0[2] ), ds0*K12 );
                        r2[2] = make_double2( ds1*K12, ds2*K12 );
                        const double cf0 = -cr[d]*( 0. + ds0 );
                        const double cf1 = fma( -cr[d], ds1, cf0 );
                        *reinterpret_cast<double2 *>( rec + 18 + 2*d ) = make_double2( cf0, cf1 );
                    }
                    *reinterpret_cast<double2 *>( rec + 24 ) = make_double2( 0., 0. );
                }
            } else if( active ) {
                const bool home = nx <= 1;           // the consumer deposits the home part; movers in 2+ dimensions go to the queue whole
                double xe = 0., xc = 0.;
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    double w1[NW], s0[NW];
                    shape2( dl1[d], w1 );
                    shape2( xdelta[d], s0 );                       // recomputed rather than kept in registers since the gather
                    const int shift = ( ( shifts >> ( 2*d ) ) & 3 ) - 1;
                    double s1a = w1[0], s1b = w1[1], s1c = w1[2], s1e = 0.;
                    if( shift > 0 ) { s1e = w1[2]; s1c = w1[1]; s1b = w1[0]; s1a = 0.; }
                    else if( shift < 0 ) { s1e = w1[0]; s1a = w1[1]; s1b = w1[2]; s1c = 0.; }
                    const double ds0 = s1a - s0[0], ds1 = s1b - s0[1], ds2 = s1c - s0[2];
                    double2 *r2 = reinterpret_cast<double2 *>( rec + 6*d );
                    r2[0] = make_double2( fma( 0.5, ds0, s0[0] ), fma( 0.5, ds1, s0[1] ) );
                    r2[1] = make_double2( fma( 0.5, ds2, s0[2] ), ds0*K12 );
                    r2[2] = make_double2( ds1*K12, ds2*K12 );
                    const double crd = home ? cr[d] : 0.;
                    const double cf0 = -crd*( ( shift < 0 ? s1e : 0. ) + ds0 );
                    const double cf1 = fma( -crd, ds1, cf0 );
                    *reinterpret_cast<double2 *>( rec + 18 + 2*d ) = make_double2( cf0, cf1 );
                    if( shift != 0 ) {
                        xmeta = d + ( shift > 0 ? 3 : 0 );
                        xe = s1e;
                        xc = shift > 0 ? fma( -crd, ds2, cf1 ) : -crd*s1e;     // flux coefficient at the point outside the home window
                    }
                }
                *reinterpret_cast<double2 *>( rec + 24 ) = make_double2( xe, xc );
            }
            __syncwarp();
            if( !direct ) {
                if( lane == 0 ) mbar_arrive( &full_bar[grp] );
            } else if( active ) {
                o2::self_consume( rec, sJ + zj + ( ( cellt >> 5 )*T::JY + ( ( cellt >> 3 ) & 3 ) )*T::JZ + ( cellt & 7 ) );
            }
            // ---------------- particles that moved to the next node in ONE dimension: the 21 values outside the
            //                  home window, one particle per 8-lane octet, 3 passes of 8 table items
            const bool one = active && nx == 1;
            const unsigned rem = __ballot_sync( 0xffffffffu, one );
            if( rem ) {
                if( one ) xsrc[warp][__popc( rem & ( ( 1u << lane ) - 1u ) )] = ( unsigned short )( lane | ( xmeta << 5 ) | ( cellt << 8 ) );
                __syncwarp();
                const int n1 = __popc( rem );
#pragma unroll 1
                for( int k = lane >> 3; k < n1; k += 4 ) {
                    const int e = xsrc[warp][k];
                    const double *rc = recbuf + rec_off( ( L & ~31 ) + ( e & 31 ) );
                    const int2 *tb = xtab + 24*( ( e >> 5 ) & 7 ) + ( lane & 7 );
                    const int cx = e >> 8;
                    jbox_t *jbx = sJ + zj + ( ( cx >> 5 )*T::JY + ( ( cx >> 3 ) & 3 ) )*T::JZ + ( cx & 7 );
#pragma unroll
                    for( int h=0; h<3; h++ ) {
                        const int2 en = tb[8*h];
                        if( en.x & ( 1 << 21 ) ) {
                            const bool isb = en.x & ( 1 << 20 );
                            const double C = rc[en.x & 31];
                            const double P = rc[( en.x >> 5 ) & 31]*( isb ? 0.5 : 1.0 );
                            const double Q = rc[( en.x >> 10 ) & 31]*( isb ? K12 : 1.0 );
                            const double *bp = rc + ( ( en.x >> 15 ) & 31 );
                            jadd_scaled( jbx + en.y, C*fma( P, bp[0], Q*bp[3] ) );
                        }
                    }
                }
                __syncwarp();
            }
            // ---------------- particles that moved in 2 or 3 dimensions: the warp's queue, 4 at a time (cross_pass)
            unsigned xmask = __ballot_sync( 0xffffffffu, active && nx > 1 );
            while( xmask ) {
                const int room = XQ2 - qn;
                const int rank = __popc( xmask & ( ( 1u << lane ) - 1u ) );
                const bool mineq = ( ( xmask >> lane ) & 1u ) && rank < room;
                if( mineq ) {
                    const int e = ( qh + qn + rank ) % XQ2;
#pragma unroll
                    for( int d=0; d<3; d++ ) { xq[( 0+d )*XQ2+e] = xdelta[d]; xq[( 3+d )*XQ2+e] = xnpos[d]; xq[( 6+d )*XQ2+e] = cr[d]; }
                    xqm[e] = cellt | ( shifts << 16 );
                }
                const unsigned done = __ballot_sync( 0xffffffffu, mineq );
                xmask &= ~done;
                qn += __popc( done );
                __syncwarp();
                while( qn >= 4 || ( xmask && qn > 0 ) ) {
                    cross_pass<2, XQ2>( sJ + zj, xq, xqm, xscr, qh, qn, lane & 7, lane >> 3, 1.0 );
                    const int took = qn < 4 ? qn : 4;
                    qh = ( qh + took ) % XQ2;
                    qn -= took;
                }
            }
            ws = nws; q0 = nq0; we = nwe; q1 = nq1; more = nmore; row = nrow; ipi = nip; isi = nis;
        }
        while( qn > 0 ) {
            cross_pass<2, XQ2>( sJ + zj, xq, xqm, xscr, qh, qn, lane & 7, lane >> 3, 1.0 );
            const int took = qn < 4 ? qn : 4;
            qh = ( qh + took ) % XQ2;
            qn -= took;
        }
        if( bad ) atomicAdd( &a.iflags[1], bad );   // particles not in the cell their sort key says
    } else {
        // ============================================================ consumer of group warp-8
        const int grp = warp - 8;
        const int *coff = cell_off[grp];
        const int total = coff[GCELLS];
        const double *recbuf = recbase + grp*RECBUF;
        const int c = lane >> 2;                               // z of this lane's cells
        const bool wk = ( lane & 3 ) < 3;                      // lane 3 of a quad idles
        const int cc = wk ? ( lane & 3 ) : 0;                  // current component
        const int da = cc == 0 ? 1 : 0, db = cc == 2 ? 1 : 2;  // its transverse dimensions
        const int oa = 6*da, ob = 6*db, oc = 18 + 2*cc;
        unsigned pa = smem_u32( recbuf + oa ), pb = smem_u32( recbuf + ob ), pc = smem_u32( recbuf + oc );   // this lane's three pieces of record 0
        asm volatile( "" : "+r"( pa ), "+r"( pb ), "+r"( pc ) );      // kept in registers: the compiler would rebuild them from the thread index in every turn
        // J-box strides of the flux index and of the two transverse indices, and the offset of (f,j,k) = (0,0,0)
        const int sf_ = cc == 0 ? T::JY*T::JZ : ( cc == 1 ? T::JZ : 1 );
        const int sa_ = cc == 0 ? T::JZ : T::JY*T::JZ;
        const int sb_ = cc == 2 ? T::JZ : 1;
        const int jo_ = cc*T::JBOX + 2*sf_ + sa_ + sb_;
        double acc[2][NW][NW];
#pragma unroll
        for( int f=0; f<2; f++ )
#pragma unroll
            for( int j=0; j<NW; j++ )
#pragma unroll
                for( int k=0; k<NW; k++ ) acc[f][j][k] = 0.;
        // Pull the group's particle columns into L2 while the producers work on the first window (this warp has
        // nothing to consume yet): one lane per cell, in stream order, asks for the lines its cell's particles sit in
        // (through the pending sort order their source slots are the old slots of the same particles, i.e. one
        // short stretch per cell plus the few that moved in).
#pragma unroll 1
        for( int ct = GCELLS*grp + lane; ct < GCELLS*( grp+1 ); ct += 32 ) {
            const int cnt = cell_cnt[ct];
            if( cnt == 0 ) continue;
            size_t lo = ( size_t )cell_first[ct], hi = lo + ( size_t )cnt - 1;
            if( a.perm ) {
                const size_t p0 = ( size_t )a.perm[lo], p1 = ( size_t )a.perm[hi];
                asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.perm + lo + 32 ) );
                lo = p0 < p1 ? p0 : p1; hi = p0 < p1 ? p1 : p0;
                if( hi - lo > 64 ) hi = lo + 64;                 // movers from far away: not worth chasing
            }
#pragma unroll 1
            for( int c=0; c<7; c++ )
                for( size_t i = lo & ~( size_t )15; i <= hi; i += 16 ) asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.in[c] + i ) );
            asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.qin + lo ) );
        }
        int ws = 0, q0 = 0, we, q1;
        const bool direct = total < DIRECT_PPC*GCELLS;        // sparse group: the producers deposit themselves
#pragma unroll 1
        for( int r = 0; !direct && next_window( coff, total, lane, ws, q0, we, q1 ); r++, ws = we, q0 = q1 ) {
            // The window holds m <= 8 whole cells (or a piece of one).  A lane quad takes one cell; when the cells are
            // dense (m <= 4) 2, 4 or 8 quads SHARE a cell, each taking every 2nd / 4th / 8th of its records, so that
            // the walk stays ~GROUP/8 records long whatever the number of particles per cell.  The sums go to the
            // J box at the end of every window.
            const int m = max( q1 - q0, 1 );
            const int sh = m > 4 ? 0 : ( m > 2 ? 1 : ( m > 1 ? 2 : 3 ) );
            const int idx = c >> sh;
            const int q = q0 + idx;
            const bool has = wk && idx < m;
            int s = 0, hi = 0;
            if( has ) {
                s = max( coff[q], ws ) - ws + ( c & ( ( 1 << sh ) - 1 ) );
                hi = min( coff[q+1], we ) - ws;
            }
            const int step = 1 << sh;
            const bool fin = s < hi;
            tma_wait( &full_bar[grp], r & 1 );
            // software pipeline: the record of the next particle is loaded while the sums of this one are formed (its 28
            // registers are free as soon as the nine W are known).  The loads are predicated instructions on shared-memory
            // addresses kept in registers: as a branch around plain loads the compiler put them AFTER the 18 accumulations
            // and rebuilt the three lane offsets from the thread index in every turn (83 instructions per record where the
            // arithmetic needs 36).
            double2 a0, a1, a2, b0, b1, b2, cf;
            a0 = a1 = a2 = b0 = b1 = b2 = cf = make_double2( 0., 0. );
            {
                const unsigned o = rec_byte( s );
                const int on = s < hi;
                lds_v2( a0, pa + o, on ); lds_v2( a1, pa + o + 16, on ); lds_v2( a2, pa + o + 32, on );
                lds_v2( b0, pb + o, on ); lds_v2( b1, pb + o + 16, on ); lds_v2( b2, pb + o + 32, on );
                lds_v2( cf, pc + o, on );
            }
#pragma unroll 1
            while( __any_sync( 0xffffffffu, s < hi ) ) {
                double W[NW][NW];
                {
                    const double Ma[NW] = { a0.x, a0.y, a1.x }, Da[NW] = { a1.y, a2.x, a2.y };
                    const double Mb[NW] = { b0.x, b0.y, b1.x }, Db[NW] = { b1.y, b2.x, b2.y };
#pragma unroll
                    for( int j=0; j<NW; j++ )
#pragma unroll
                        for( int k=0; k<NW; k++ ) W[j][k] = fma( Ma[j], Mb[k], Da[j]*Db[k] );
                }
                const double c0_ = s < hi ? cf.x : 0., c1_ = s < hi ? cf.y : 0.;     // a lane past its cell adds nothing
                s += step;
                {
                    const unsigned o = rec_byte( s );
                    const int on = s < hi;
                    lds_v2( a0, pa + o, on ); lds_v2( a1, pa + o + 16, on ); lds_v2( a2, pa + o + 32, on );
                    lds_v2( b0, pb + o, on ); lds_v2( b1, pb + o + 16, on ); lds_v2( b2, pb + o + 32, on );
                    lds_v2( cf, pc + o, on );
                }
#pragma unroll
                for( int j=0; j<NW; j++ )
#pragma unroll
                    for( int k=0; k<NW; k++ ) {
                        acc[0][j][k] = fma( c0_, W[j][k], acc[0][j][k] );
                        acc[1][j][k] = fma( c1_, W[j][k], acc[1][j][k] );
                    }
            }
            __syncwarp();
            if( lane == 0 ) mbar_arrive( &empty_bar[grp] );    // the producers may write the records of the next window
            if( fin ) {
                const int row = q >> 3;
                jbox_t *jb = sJ + zj + ( ( 2*grp + ( row >> 2 ) )*T::JY + ( row & 3 ) )*T::JZ + ( q & 7 ) + jo_;
#pragma unroll
                for( int f=0; f<2; f++ )
#pragma unroll
                    for( int j=0; j<NW; j++ )
#pragma unroll
                        for( int k=0; k<NW; k++ ) { jadd_scaled( jb + f*sf_ + j*sa_ + k*sb_, acc[f][j][k] ); acc[f][j][k] = 0.; }
            }
        }
    }
    __syncthreads();

    // ---------------- flush the J box: convert the fixed-point sums to double in place, then ONE elected thread hands
    //                  the three boxes to the TMA unit, which adds them into Jx, Jy, Jz in HBM
    for( int t = tid; t < 3*T::JBOX; t += NTHR ) {
        const long long iv = ( long long )sJ[t];
        reinterpret_cast<double *>( sJ )[t] = ( double )iv*a.jinv;
    }
    tma_store_fence();
    __syncthreads();
    if( tid == 0 ) {
#pragma unroll
        for( int c=0; c<3; c++ )
            tma_reduce_add_3d( &tm.j[c], sJ + c*T::JBOX, c0[2] + g.o[2] - T::H - 1 - zj, c0[1] + g.o[1] - T::H - 1, c0[0] + g.o[0] - T::H - 1 );
        tma_commit_and_wait_read();
    }
}

// =================================================================================================
// Order-4 production kernel: the design of k_dynamics_o2 (particle stream walked by producer warps, per-cell current
// sums formed by consumer warps) with the order-4 sizes.
//
// Interpolator3D4Order / Projector3D4Order work on 5 nodes per dimension: the home window of a (cell, component) is
// 4 flux points x 5 x 5 nodes = 100 values, too many for one lane.  The consumer lane is therefore one ROW of that
// window — (cell, component, j), 4 x 5 = 20 running sums — 15 lanes per cell, two cells per consumer warp, four
// consumer warps per group.  One CTA (512 threads: 8 producer + 8 consumer warps, 128 registers each) per SM: the
// order-4 boxes (6 x 9 x 9 x 14 field values, 3 x 10 x 10 x 16 currents) and the 46-double records of two groups
// fill the shared memory of an SM.
//   record   per dimension M[5] = (S0+S1)/2 and DS[5]/sqrt(12) on the home nodes; 4 flux coefficients per component
//            (fixed-point units); the weight on the node outside the home window and the flux coefficient of the
//            flux point outside it for a particle that moved to the next node in one dimension;
//   consumer J_c[f][j][k] += Cf_c[f] * ( M_a[j] M_b[k] + DS_a[j] DS_b[k] / 12 ),   f < 4, k < 5, one j per lane;
//   one-dimension movers: 25 + 2 x 20 = 65 values outside the home window, table-driven, one particle per octet;
//   movers in 2 or 3 dimensions: the warp's queue (cross_pass<4>).
// The gather keeps its loop over the six components rolled (the unrolled body is 750 loads + 930 FMA).
// =================================================================================================
namespace o4 {
using T = CG<4>::T;
constexpr int NW = 5;
constexpr int NF = NW - 1;                               // home flux points
constexpr int NPROD = 256;
constexpr int NCONS = 256;                               // 4 consumer warps per group
constexpr int NTHR = NPROD + NCONS;
constexpr int GROUP = 128;
constexpr int NCELL = T::TX*T::TY*T::TZ;
constexpr int GCELLS = NCELL/2;
static_assert( T::TX == 4 && T::TY == 4 && T::TZ == 8, "the row / group arithmetic below is written for 4 x 4 x 8 tiles" );
constexpr int CFO = 6*NW;                                // record: cf[3][NF] after the 3 x (M[5], DS[5])
constexpr int SEO = CFO + 3*NF;                          // weight on the outer node
constexpr int XCO = SEO + 1;                             // flux coefficient at the outer flux point
constexpr int REC = 46;                                  // 44 used; 46 doubles = 92 words: consecutive slots fall in distinct 16-byte bank groups
__device__ __forceinline__ int rec_off( int L ) { return L*REC + 2*( L >> 4 ); }
constexpr int RECBUF = GROUP*REC + 2*( GROUP/16 );
constexpr int XQ4 = 8;
constexpr int XQD4 = 9*XQ4 + XQ4/2;
constexpr int XSCR4 = 4*CGDim<4>::XSCR;                  // cross_pass scratch per producer warp (4 octets x 36)
constexpr int NITEM = NW*NW + 2*NF*NW;                   // 65 values outside the home window of a one-dimension mover
constexpr int NPASS = ( NITEM + 7 )/8;                   // 9 passes of an octet
constexpr int XTAB = 6*8*NPASS;                          // table entries (int2 each)
constexpr double K12 = 0.28867513459481288225;
constexpr size_t BYTES = ( size_t )( 6*T::FBOX + 3*T::JBOX + 2*RECBUF + 8*XQD4 + 8*XSCR4 + XTAB )*sizeof( double );
constexpr unsigned TMA_BYTES = 6u*T::FVOL*sizeof( double );

// windows as in o2 (next_window of namespace o2 works on GROUP = 128 items and at most 8 cells)
using sb200::o2::next_window;
using sb200::o2::mbar_arrive;

// entry `it` (0 .. 8*NPASS-1) of the outer-part table, see o2::xtab_entry: value = rec[iC] * ( cP rec[iP] rec[iB] + cQ rec[iQ] rec[iB+NW] )
__device__ __forceinline__ int2 xtab_entry( int m, int it )
{
    const int d = m % 3, up = m / 3;
    const int a1 = d == 0 ? 1 : 0, a2 = d == 2 ? 1 : 2;
    const int st[3] = { T::JY*T::JZ, T::JZ, 1 };
    int comp = 0, iC = 0, iP = 0, iQ = 0, iB = 0, isb = 0, off = 0, valid = 1;
    if( it < NW*NW ) {                                   // flux component d at the outer flux point, 5 x 5 home nodes
        const int j = it/NW, kk = it - NW*j;
        comp = d; iC = XCO; iP = 2*NW*a1 + j; iQ = iP + NW; iB = 2*NW*a2 + kk;
        off = ( up ? NW+1 : 1 )*st[d] + ( 1+j )*st[a1] + ( 1+kk )*st[a2];
    } else if( it < NITEM ) {                            // the two other components on the outer node plane: 4 flux points x 5 nodes
        const int t = it - NW*NW, second = t >= NF*NW ? 1 : 0, u = t - NF*NW*second, f = u/NW, kk = u - NW*f;
        const int bdim = second ? a1 : a2;
        comp = second ? a2 : a1;
        iC = CFO + NF*comp + f; iP = SEO; iQ = SEO; iB = 2*NW*bdim + kk; isb = 1;
        off = ( 2+f )*st[comp] + ( up ? NW+1 : 0 )*st[d] + ( 1+kk )*st[bdim];
    } else valid = 0;
    return make_int2( iC | ( iP << 6 ) | ( iQ << 12 ) | ( iB << 18 ) | ( isb << 24 ) | ( valid << 25 ), comp*T::JBOX + off );
}
}

template<int PUSHER, bool SCRATCH, bool REMOVE>
__global__ void __launch_bounds__( o4::NTHR, 1 ) k_dynamics_o4( const GridDev g, const DynArgs a, const __grid_constant__ FieldMaps tm )
{
    using namespace o4;
    extern __shared__ __align__( 128 ) double smem[];
    double *sF = smem;
    jbox_t *sJ = reinterpret_cast<jbox_t *>( smem + 6*T::FBOX );
    double *recbase = smem + 6*T::FBOX + 3*T::JBOX;
    double *xqbase = recbase + 2*RECBUF;
    double *xscrbase = xqbase + 8*XQD4;
    int2 *xtab = reinterpret_cast<int2 *>( xscrbase + 8*XSCR4 );
    __shared__ __align__( 8 ) unsigned long long tma_bar;
    __shared__ int cell_first[NCELL];
    __shared__ int cell_cnt[NCELL];
    __shared__ int cell_off[2][GCELLS+1];
    __shared__ unsigned short xsrc[8][32];
    __shared__ __align__( 8 ) unsigned long long full_bar[2], empty_bar[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    int b = blockIdx.x;
    const int tz = b % a.tiles[2]; b /= a.tiles[2];
    const int ty = b % a.tiles[1];
    const int tx = b / a.tiles[1];
    const int c0[3] = { tx*T::TX, ty*T::TY, tz*T::TZ };
    const int zs = ( c0[2] + g.o[2] - T::H ) & 1;
    const int zj = ( c0[2] + g.o[2] - T::H - 1 ) & 1;

    if( tid == 0 ) {
        tma_bar_init( &tma_bar, 1 );
        tma_bar_init( &full_bar[0], 4 ); tma_bar_init( &full_bar[1], 4 );       // 4 producer warps write a window
        tma_bar_init( &empty_bar[0], 4 ); tma_bar_init( &empty_bar[1], 4 );     // 4 consumer warps read it
    }
    __syncthreads();
    if( tid == 0 ) {
        tma_expect( &tma_bar, TMA_BYTES );
#pragma unroll
        for( int c=0; c<6; c++ )
            tma_load_3d( sF + c*T::FBOX, &tm.m[c], &tma_bar, c0[2] + g.o[2] - T::H - zs, c0[1] + g.o[1] - T::H, c0[0] + g.o[0] - T::H );
    }
    int mine = 0;
    if( tid < NCELL ) {
        const int lz = tid % T::TZ, ly = ( tid / T::TZ ) % T::TY, lx = tid / ( T::TZ*T::TY );
        const int ix = c0[0]+lx, iy = c0[1]+ly, iz = c0[2]+lz;
        int beg = 0, cnt = 0;
        if( ix < g.ncell[0] && iy < g.ncell[1] && iz < g.ncell[2] ) {
            const int cell = ( ix*g.ncell[1] + iy )*g.ncell[2] + iz;
            beg = a.first[cell];
            cnt = a.first[cell+1] - beg;
        }
        cell_first[tid] = beg;
        cell_cnt[tid] = cnt;
        mine = cnt;
    } else {
        for( int t = tid - NCELL; t < XTAB; t += NTHR - NCELL ) xtab[t] = xtab_entry( t/( 8*NPASS ), t%( 8*NPASS ) );
    }
    for( int t = tid; t < 3*T::JBOX; t += NTHR ) sJ[t] = 0ull;
    const int any = __syncthreads_or( mine );
    tma_wait( &tma_bar, 0 );
    if( !any ) return;
    if( warp < 2 ) {
        const int a0 = cell_cnt[GCELLS*warp + 2*lane], a1 = cell_cnt[GCELLS*warp + 2*lane + 1];
        int inc = a0 + a1;
#pragma unroll
        for( int d=1; d<32; d<<=1 ) { const int u = __shfl_up_sync( 0xffffffffu, inc, d ); if( lane >= d ) inc += u; }
        cell_off[warp][2*lane] = inc - a0 - a1;
        cell_off[warp][2*lane+1] = inc - a1;
        if( lane == 31 ) cell_off[warp][GCELLS] = inc;
    }
    __syncthreads();

    if( warp < 8 ) {
        // ============================================================ producers
        const int grp = warp >> 2;
        const int L = tid & ( GROUP-1 );
        const int *coff = cell_off[grp];
        const int total = coff[GCELLS];
        double *recbuf = recbase + grp*RECBUF;
        double *rec = recbuf + rec_off( L );
        double *xq = xqbase + XQD4*warp;
        int *xqm = reinterpret_cast<int *>( xq + 9*XQ4 );
        double *xscr = xscrbase + XSCR4*warp + CGDim<4>::XSCR*( lane >> 3 );
        const int base[3] = { g.begin[0] + g.o[0] + c0[0], g.begin[1] + g.o[1] + c0[1], g.begin[2] + g.o[2] + c0[2] };
        int qh = 0, qn = 0, bad = 0;
        int ws = 0, q0 = 0, we = 0, q1 = 0;
        const bool direct = total < o2::DIRECT_PPC*GCELLS;       // sparse group: see k_dynamics_o2
        auto window = [&]( int &ws_, int &q0_, int &we_, int &q1_ ) -> bool {
            if( !direct ) return next_window( coff, total, lane, ws_, q0_, we_, q1_ );
            we_ = min( ws_ + GROUP, total ); q1_ = q0_;
            return ws_ < total;
        };
        auto locate = [&]( int ws_, int q0_, int we_, int &row_, int &ip_, int &is_ ) {
            const int s_ = ws_ + L;
            row_ = q0_ >> 3;
            ip_ = -1; is_ = -1;
            if( s_ < we_ ) {
                if( direct ) {
                    row_ = 0;
#pragma unroll
                    for( int k_=1; k_<8; k_++ ) row_ += s_ >= coff[T::TZ*k_];
                } else row_ += s_ >= coff[T::TZ*( row_+1 )];
                ip_ = cell_first[GCELLS*grp + T::TZ*row_] + ( s_ - coff[T::TZ*row_] );
                is_ = a.perm ? a.perm[ip_] : ip_;
            }
        };
        bool more = window( ws, q0, we, q1 );
        int row = 0, ipi = -1, isi = -1;
        if( more ) locate( ws, q0, we, row, ipi, isi );

#pragma unroll 1
        for( int r = 0; more; r++ ) {
            int nws = we, nq0 = q1, nwe = 0, nq1 = 0, nrow = 0, nip = -1, nis = -1;
            const bool nmore = window( nws, nq0, nwe, nq1 );
            if( nmore ) locate( nws, nq0, nwe, nrow, nip, nis );
            const bool active = ipi >= 0;
            double dl1[3] = { 0., 0., 0. }, cr[3] = { 0., 0., 0. }, xdelta[3] = { 0., 0., 0. }, xnpos[3] = { 0., 0., 0. };
            int shifts = 0x15, cellt = 0, nx = 0;
            // ---------------- part A: gather, push, tag, key
            if( active ) {
                const size_t ip = ( size_t )ipi, is = ( size_t )isi;
                const double pos[3] = { a.in[0][is], a.in[1][is], a.in[2][is] };
                double px = a.in[3][is], py = a.in[4][is], pz = a.in[5][is];
                const short charge = a.qin[is];

                double S0[3][NW], cd[3][NW];
                int sp[3], sd[3], cl[3];
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    const double pn = __dmul_rn( pos[d], g.dxi[d] );
                    const int ipn = ( int )round( pn );
                    xdelta[d] = pn - ( double )ipn;
                    Shape<4>::w( xdelta[d], S0[d] );
                    const int idn = ( int )round( pn + 0.5 );
                    Shape<4>::w( pn - ( double )idn + 0.5, cd[d] );
                    cl[d] = ipn - base[d];
                    sd[d] = idn - ipn;
                }
                {
                    const int lx = 2*grp + ( row >> 2 ), ly = row & 3;
                    if( cl[0] != lx || cl[1] != ly || ( unsigned )cl[2] >= ( unsigned )T::TZ ) {
                        bad++;
                        cl[0] = lx; cl[1] = ly; cl[2] = cl[2] < 0 ? 0 : ( cl[2] >= T::TZ ? T::TZ-1 : cl[2] );
                    }
                }
                cellt = ( cl[0]*T::TY + cl[1] )*T::TZ + cl[2];
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    sp[d] = cl[d] + T::H + ( d == 2 ? zs : 0 );
                    sd[d] += sp[d];
                }
                double EB[6];
                EB[0] = gather<T>( sF+0*T::FBOX, cd[0], S0[1], S0[2], sd[0], sp[1], sp[2] );
                EB[1] = gather<T>( sF+1*T::FBOX, S0[0], cd[1], S0[2], sp[0], sd[1], sp[2] );
                EB[2] = gather<T>( sF+2*T::FBOX, S0[0], S0[1], cd[2], sp[0], sp[1], sd[2] );
                EB[3] = gather<T>( sF+3*T::FBOX, S0[0], cd[1], cd[2], sp[0], sd[1], sd[2] );
                EB[4] = gather<T>( sF+4*T::FBOX, cd[0], S0[1], cd[2], sd[0], sp[1], sd[2] );
                EB[5] = gather<T>( sF+5*T::FBOX, cd[0], cd[1], S0[2], sd[0], sd[1], sp[2] );
                const double Ex = EB[0], Ey = EB[1], Ez = EB[2], Bx = EB[3], By = EB[4], Bz = EB[5];

                const double cmd = ( double )charge*a.one_over_mass*g.dts2;
                double dxp, dyp, dzp, invgf;
                push<PUSHER>( cmd, g.dt, px, py, pz, Ex, Ey, Ez, Bx, By, Bz, dxp, dyp, dzp, invgf );
                const double npos[3] = { pos[0] + dxp, pos[1] + dyp, pos[2] + dzp };
                a.col[0][ip] = npos[0]; a.col[1][ip] = npos[1]; a.col[2][ip] = npos[2];
                a.col[3][ip] = px; a.col[4][ip] = py; a.col[5][ip] = pz;
                const double weight = a.in[6][is];
                if( a.perm ) { a.col[6][ip] = weight; a.q[ip] = charge; }
                if( SCRATCH ) {
                    a.sc_E[0*a.n+ip] = Ex; a.sc_E[1*a.n+ip] = Ey; a.sc_E[2*a.n+ip] = Ez;
                    a.sc_B[0*a.n+ip] = Bx; a.sc_B[1*a.n+ip] = By; a.sc_B[2*a.n+ip] = Bz;
                    a.sc_invgf[ip] = invgf;
#pragma unroll
                    for( int d=0; d<3; d++ ) {
                        a.sc_iold[d*a.n+ip] = cl[d] + c0[d] + g.o[d];
                        a.sc_delta[d*a.n+ip] = xdelta[d];
                    }
                }

                bool removed;
                const int tag = boundary_tag<REMOVE>( a, g, npos, px, py, pz, weight, ip, removed );
                const double charge_weight = removed ? 0. : g.inv_cell_volume*( double )charge*weight*a.jscale;
                cr[0] = charge_weight*g.d_ov_dt[0]; cr[1] = charge_weight*g.d_ov_dt[1]; cr[2] = charge_weight*g.d_ov_dt[2];

                int nkey[3];
                shifts = 0;
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    const double pn = __dmul_rn( npos[d], g.dxi[d] );
                    const int ipn = ( int )round( pn );
                    dl1[d] = pn - ( double )ipn;
                    const int shift = ipn - base[d] - cl[d];
                    shifts |= ( shift+1 ) << ( 2*d );
                    nx += shift != 0;
                    nkey[d] = ( int )( ( double )ipn - g.min_loc_round[d] );
                    xnpos[d] = pn;
                }
                int key = tag;
                if( tag == 0 ) { key = ( nkey[0]*g.ncell[1] + nkey[1] )*g.ncell[2] + nkey[2]; atomicAdd( &a.count[key], 1 ); }
                else if( tag < -1 ) atomicAdd( &a.leave_counts[-tag-2], 1 );
                a.key[ip] = key;
            }
            // ---------------- the consumers must have finished with the records of the previous round
            if( r > 0 && !direct ) tma_wait( &empty_bar[grp], ( r-1 ) & 1 );
            // ---------------- part B: new shape, the record of the deposit (see k_dynamics_o2), 5 home nodes per dimension
            int xmeta = 0;
            if( active ) {
                const bool home = nx <= 1;
                double xe = 0., xc = 0.;
#pragma unroll
                for( int d=0; d<3; d++ ) {
                    double w1[NW], s0[NW], s1[NW], ds[NW];
                    Shape<4>::w( dl1[d], w1 );
                    Shape<4>::w( xdelta[d], s0 );
                    const int shift = ( ( shifts >> ( 2*d ) ) & 3 ) - 1;
                    double s1e = 0.;
#pragma unroll
                    for( int s=0; s<NW; s++ ) s1[s] = w1[s];
                    if( shift > 0 ) {
                        s1e = w1[NW-1];
#pragma unroll
                        for( int s=NW-1; s>0; s-- ) s1[s] = w1[s-1];
                        s1[0] = 0.;
                    } else if( shift < 0 ) {
                        s1e = w1[0];
#pragma unroll
                        for( int s=0; s<NW-1; s++ ) s1[s] = w1[s+1];
                        s1[NW-1] = 0.;
                    }
                    double *rd = rec + 2*NW*d;
#pragma unroll
                    for( int s=0; s<NW; s++ ) {
                        ds[s] = s1[s] - s0[s];
                        rd[s] = fma( 0.5, ds[s], s0[s] );
                        rd[NW+s] = ds[s]*K12;
                    }
                    const double crd = home ? cr[d] : 0.;
                    double run = -crd*( ( shift < 0 ? s1e : 0. ) + ds[0] );
                    double *rcf = rec + CFO + NF*d;
                    rcf[0] = run;
#pragma unroll
                    for( int f=1; f<NF; f++ ) { run = fma( -crd, ds[f], run ); rcf[f] = run; }
                    if( shift != 0 ) {
                        xmeta = d + ( shift > 0 ? 3 : 0 );
                        xe = s1e;
                        xc = shift > 0 ? fma( -crd, ds[NW-1], run ) : -crd*s1e;
                    }
                }
                rec[SEO] = xe; rec[XCO] = xc;
            }
            __syncwarp();
            if( !direct ) {
                if( lane == 0 ) mbar_arrive( &full_bar[grp] );
            } else if( active ) {
                // sparse group: this lane is the consumer of its own record, one (component, row) after the other
                jbox_t *jcell = sJ + zj + ( ( cellt >> 5 )*T::JY + ( ( cellt >> 3 ) & 3 ) )*T::JZ + ( cellt & 7 );
#pragma unroll 1
                for( int cj=0; cj<3*NW; cj++ ) {
                    const int cc = cj / NW, j = cj - NW*cc;
                    const int da = cc == 0 ? 1 : 0, db = cc == 2 ? 1 : 2;
                    const int sf_ = cc == 0 ? T::JY*T::JZ : ( cc == 1 ? T::JZ : 1 );
                    const int sa_ = cc == 0 ? T::JZ : T::JY*T::JZ;
                    const int sb_ = cc == 2 ? T::JZ : 1;
                    const double ma = rec[2*NW*da + j], dj = rec[2*NW*da + NW + j];
                    const double *rb = rec + 2*NW*db, *rcf = rec + CFO + NF*cc;
                    jbox_t *jb = jcell + cc*T::JBOX + 2*sf_ + ( 1+j )*sa_ + sb_;
                    if( rcf[0] != 0. || rcf[NF-1] != 0. ) {
#pragma unroll
                        for( int k_=0; k_<NW; k_++ ) {
                            const double W = fma( ma, rb[k_], dj*rb[NW+k_] );
#pragma unroll
                            for( int f=0; f<NF; f++ ) jadd_scaled( jb + f*sf_ + k_*sb_, rcf[f]*W );
                        }
                    }
                }
            }
            // ---------------- one-dimension movers: the 65 values outside the home window, one particle per octet
            const bool one = active && nx == 1;
            const unsigned rem = __ballot_sync( 0xffffffffu, one );
            if( rem ) {
                if( one ) xsrc[warp][__popc( rem & ( ( 1u << lane ) - 1u ) )] = ( unsigned short )( lane | ( xmeta << 5 ) | ( cellt << 8 ) );
                __syncwarp();
                const int n1 = __popc( rem );
#pragma unroll 1
                for( int k = lane >> 3; k < n1; k += 4 ) {
                    const int e = xsrc[warp][k];
                    const double *rc = recbuf + rec_off( ( L & ~31 ) + ( e & 31 ) );
                    const int2 *tb = xtab + 8*NPASS*( ( e >> 5 ) & 7 ) + ( lane & 7 );
                    const int cx = e >> 8;
                    jbox_t *jbx = sJ + zj + ( ( cx >> 5 )*T::JY + ( ( cx >> 3 ) & 3 ) )*T::JZ + ( cx & 7 );
#pragma unroll 1
                    for( int h=0; h<NPASS; h++ ) {
                        const int2 en = tb[8*h];
                        if( en.x & ( 1 << 25 ) ) {
                            const bool isb = en.x & ( 1 << 24 );
                            const double C = rc[en.x & 63];
                            const double P = rc[( en.x >> 6 ) & 63]*( isb ? 0.5 : 1.0 );
                            const double Q = rc[( en.x >> 12 ) & 63]*( isb ? K12 : 1.0 );
                            const double *bp = rc + ( ( en.x >> 18 ) & 63 );
                            jadd_scaled( jbx + en.y, C*fma( P, bp[0], Q*bp[NW] ) );
                        }
                    }
                }
                __syncwarp();
            }
            // ---------------- movers in 2 or 3 dimensions: the warp's queue, 4 at a time (cross_pass)
            unsigned xmask = __ballot_sync( 0xffffffffu, active && nx > 1 );
            while( xmask ) {
                const int room = XQ4 - qn;
                const int rank = __popc( xmask & ( ( 1u << lane ) - 1u ) );
                const bool mineq = ( ( xmask >> lane ) & 1u ) && rank < room;
                if( mineq ) {
                    const int e = ( qh + qn + rank ) % XQ4;
#pragma unroll
                    for( int d=0; d<3; d++ ) { xq[( 0+d )*XQ4+e] = xdelta[d]; xq[( 3+d )*XQ4+e] = xnpos[d]; xq[( 6+d )*XQ4+e] = cr[d]; }
                    xqm[e] = cellt | ( shifts << 16 );
                }
                const unsigned done = __ballot_sync( 0xffffffffu, mineq );
                xmask &= ~done;
                qn += __popc( done );
                __syncwarp();
                while( qn >= 4 || ( xmask && qn > 0 ) ) {
                    cross_pass<4, XQ4>( sJ + zj, xq, xqm, xscr, qh, qn, lane & 7, lane >> 3, 1.0 );
                    const int took = qn < 4 ? qn : 4;
                    qh = ( qh + took ) % XQ4;
                    qn -= took;
                }
            }
            ws = nws; q0 = nq0; we = nwe; q1 = nq1; more = nmore; row = nrow; ipi = nip; isi = nis;
        }
        while( qn > 0 ) {
            cross_pass<4, XQ4>( sJ + zj, xq, xqm, xscr, qh, qn, lane & 7, lane >> 3, 1.0 );
            const int took = qn < 4 ? qn : 4;
            qh = ( qh + took ) % XQ4;
            qn -= took;
        }
        if( bad ) atomicAdd( &a.iflags[1], bad );
    } else {
        // ============================================================ consumers: warp 8 + 4 grp + cw, cells 2 cw and 2 cw + 1 of a window
        const int grp = ( warp - 8 ) >> 2, cw = ( warp - 8 ) & 3;
        const int *coff = cell_off[grp];
        const int total = coff[GCELLS];
        const double *recbuf = recbase + grp*RECBUF;
        const int half = lane >> 4, hl = lane & 15;            // cell of the pair, lane in the half warp
        const int c = 2*cw + half;                             // z of this lane's cells
        const bool wk = hl < 15;
        const int cc = wk ? hl / NW : 0;                       // current component
        const int j = wk ? hl - NW*cc : 0;                     // its row of the transverse window
        const int da = cc == 0 ? 1 : 0, db = cc == 2 ? 1 : 2;
        const int oa = 2*NW*da + j, ob = 2*NW*db, oc = CFO + NF*cc;
        const int sf_ = cc == 0 ? T::JY*T::JZ : ( cc == 1 ? T::JZ : 1 );
        const int sa_ = cc == 0 ? T::JZ : T::JY*T::JZ;
        const int sb_ = cc == 2 ? T::JZ : 1;
        const int jo_ = cc*T::JBOX + 2*sf_ + ( 1+j )*sa_ + sb_;
        double acc[NF][NW];
#pragma unroll
        for( int f=0; f<NF; f++ )
#pragma unroll
            for( int k=0; k<NW; k++ ) acc[f][k] = 0.;
        // L2 prefetch of the group's particle columns (see k_dynamics_o2): the four consumer warps share the 64 cells
#pragma unroll 1
        for( int ct = GCELLS*grp + 32*cw + lane; ct < GCELLS*( grp+1 ); ct += 128 ) {
            const int cnt = cell_cnt[ct];
            if( cnt == 0 ) continue;
            size_t lo = ( size_t )cell_first[ct], hi = lo + ( size_t )cnt - 1;
            if( a.perm ) {
                const size_t p0 = ( size_t )a.perm[lo], p1 = ( size_t )a.perm[hi];
                asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.perm + lo + 32 ) );
                lo = p0 < p1 ? p0 : p1; hi = p0 < p1 ? p1 : p0;
                if( hi - lo > 64 ) hi = lo + 64;
            }
#pragma unroll 1
            for( int cc2=0; cc2<7; cc2++ )
                for( size_t i = lo & ~( size_t )15; i <= hi; i += 16 ) asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.in[cc2] + i ) );
            asm volatile( "prefetch.global.L2 [%0];" :: "l"( a.qin + lo ) );
        }
        int ws = 0, q0 = 0, we, q1;
        const bool direct = total < o2::DIRECT_PPC*GCELLS;
#pragma unroll 1
        for( int r = 0; next_window( coff, total, lane, ws, q0, we, q1 ); r++, ws = we, q0 = q1 ) {
            if( direct ) break;                                // sparse group: the producers deposit themselves
            // m <= 8 whole cells (or a piece of one) in the window; dense cells are shared by 2, 4 or 8 cell slots, each
            // taking every 2nd / 4th / 8th record (see k_dynamics_o2); the sums go to the J box at the end of every window
            const int m = max( q1 - q0, 1 );
            const int sh = m > 4 ? 0 : ( m > 2 ? 1 : ( m > 1 ? 2 : 3 ) );
            const int idx = c >> sh;
            const int q = q0 + idx;
            const bool has = wk && idx < m;
            int s = 0, hi = 0;
            if( has ) {
                s = max( coff[q], ws ) - ws + ( c & ( ( 1 << sh ) - 1 ) );
                hi = min( coff[q+1], we ) - ws;
            }
            const int step = 1 << sh;
            const bool fin = s < hi;
            tma_wait( &full_bar[grp], r & 1 );
#pragma unroll 1
            while( __any_sync( 0xffffffffu, s < hi ) ) {
                if( s < hi ) {
                    const double *rc = recbuf + rec_off( s );
                    const double ma = rc[oa], dj = rc[oa+NW];
                    const double2 *qb = reinterpret_cast<const double2 *>( rc + ob );
                    const double2 b0 = qb[0], b1 = qb[1], b2 = qb[2], b3 = qb[3], b4 = qb[4];
                    const double Mb[NW] = { b0.x, b0.y, b1.x, b1.y, b2.x }, Db[NW] = { b2.y, b3.x, b3.y, b4.x, b4.y };
                    double cf[NF];
#pragma unroll
                    for( int f=0; f<NF; f++ ) cf[f] = rc[oc+f];
#pragma unroll
                    for( int k=0; k<NW; k++ ) {
                        const double W = fma( ma, Mb[k], dj*Db[k] );
#pragma unroll
                        for( int f=0; f<NF; f++ ) acc[f][k] = fma( cf[f], W, acc[f][k] );
                    }
                }
                s += step;
            }
            __syncwarp();
            if( lane == 0 ) mbar_arrive( &empty_bar[grp] );
            if( fin ) {
                const int row = q >> 3;
                jbox_t *jb = sJ + zj + ( ( 2*grp + ( row >> 2 ) )*T::JY + ( row & 3 ) )*T::JZ + ( q & 7 ) + jo_;
#pragma unroll
                for( int f=0; f<NF; f++ )
#pragma unroll
                    for( int k=0; k<NW; k++ ) { jadd_scaled( jb + f*sf_ + k*sb_, acc[f][k] ); acc[f][k] = 0.; }
            }
        }
    }
    __syncthreads();

    for( int t = tid; t < 3*T::JBOX; t += NTHR ) {
        const long long iv = ( long long )sJ[t];
        reinterpret_cast<double *>( sJ )[t] = ( double )iv*a.jinv;
    }
    tma_store_fence();
    __syncthreads();
    if( tid == 0 ) {
#pragma unroll
        for( int c=0; c<3; c++ )
            tma_reduce_add_3d( &tm.j[c], sJ + c*T::JBOX, c0[2] + g.o[2] - T::H - 1 - zj, c0[1] + g.o[1] - T::H - 1, c0[0] + g.o[0] - T::H - 1 );
        tma_commit_and_wait_read();
    }
}

// Tensor maps of the six gathered fields for a given box: element (k,j,i) innermost first, row pitch AZ*8 B
// (a multiple of 128 B by construction of the padded layout), out-of-bounds elements read as zero.
static int field_maps( sb200_patch *p, double *const *J, int fx, int fy, int fz, int jx, int jy, int jz, FieldMaps &out )
{
    typedef CUresult ( *encode_t )( CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
    static encode_t encode = nullptr;
    if( !encode ) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        SB200_CUDA( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q ) );
        SB200_CHECK( fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver" );
        encode = ( encode_t )fn;
    }
    const GridDev &g = p->gd;
    const int fid[6] = { SB200_EX, SB200_EY, SB200_EZ, SB200_BXM, SB200_BYM, SB200_BZM };
    const cuuint64_t dims[3] = { ( cuuint64_t )g.az, ( cuuint64_t )g.ay, ( cuuint64_t )g.ax };
    const cuuint64_t strides[2] = { ( cuuint64_t )g.az*sizeof( double ), ( cuuint64_t )g.ay*g.az*sizeof( double ) };
    const cuuint32_t box[3] = { ( cuuint32_t )fz, ( cuuint32_t )fy, ( cuuint32_t )fx };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    for( int c=0; c<6; c++ ) {
        const CUresult r = encode( &out.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p->f[fid[c]], dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
        SB200_CHECK( r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for a field box" );
    }
    const cuuint32_t jbox[3] = { ( cuuint32_t )jz, ( cuuint32_t )jy, ( cuuint32_t )jx };
    for( int c=0; c<3; c++ ) {
        const CUresult r = encode( &out.j[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, J[c], dims, strides, jbox, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
        SB200_CHECK( r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for a J box" );
    }
    return 0;
}

template<int ORDER, int PUSHER, bool SCRATCH>
static int launch_cg( sb200_patch *p, const DynArgs &a, int ntiles )
{
    using T = typename CG<ORDER>::T;
    FieldMaps tm;
    if( field_maps( p, a.J, T::FX, T::FY, T::FZ, T::JX, T::JY, T::JZ, tm ) ) return 1;
    auto kern = k_dynamics_cg<ORDER, PUSHER, SCRATCH>;
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )CGDim<ORDER>::BYTES ) );
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared ) );
    kern<<<ntiles, DYN_THREADS, CGDim<ORDER>::BYTES, p->stream>>>( p->gd, a, tm );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

template<int PUSHER, bool SCRATCH, bool REMOVE>
static int launch_o2( sb200_patch *p, const DynArgs &a, int ntiles )
{
    using T = o2::T;
    FieldMaps tm;
    if( field_maps( p, a.J, T::FX, T::FY, T::FZ, T::JX, T::JY, T::JZ, tm ) ) return 1;
    auto kern = k_dynamics_o2<PUSHER, SCRATCH, REMOVE>;
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )o2::BYTES ) );
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared ) );
    kern<<<ntiles, o2::NTHR, o2::BYTES, p->stream>>>( p->gd, a, tm );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

template<int PUSHER>
static int launch_o2_flags( sb200_patch *p, const DynArgs &a, int ntiles, bool scratch )
{
    if( a.any_remove ) return scratch ? launch_o2<PUSHER, true, true>( p, a, ntiles ) : launch_o2<PUSHER, false, true>( p, a, ntiles );
    return scratch ? launch_o2<PUSHER, true, false>( p, a, ntiles ) : launch_o2<PUSHER, false, false>( p, a, ntiles );
}

static int launch_o2_pusher( sb200_patch *p, const DynArgs &a, int ntiles, int pusher, bool scratch )
{
    switch( pusher ) {
        case SB200_PUSHER_BORIS: return launch_o2_flags<SB200_PUSHER_BORIS>( p, a, ntiles, scratch );
        case SB200_PUSHER_VAY: return launch_o2_flags<SB200_PUSHER_VAY>( p, a, ntiles, scratch );
        default: return launch_o2_flags<SB200_PUSHER_HIGUERACARY>( p, a, ntiles, scratch );
    }
}

template<int PUSHER, bool SCRATCH, bool REMOVE>
static int launch_o4( sb200_patch *p, const DynArgs &a, int ntiles )
{
    using T = o4::T;
    FieldMaps tm;
    if( field_maps( p, a.J, T::FX, T::FY, T::FZ, T::JX, T::JY, T::JZ, tm ) ) return 1;
    auto kern = k_dynamics_o4<PUSHER, SCRATCH, REMOVE>;
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )o4::BYTES ) );
    SB200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared ) );
    kern<<<ntiles, o4::NTHR, o4::BYTES, p->stream>>>( p->gd, a, tm );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

template<int PUSHER>
static int launch_o4_flags( sb200_patch *p, const DynArgs &a, int ntiles, bool scratch )
{
    if( a.any_remove ) return scratch ? launch_o4<PUSHER, true, true>( p, a, ntiles ) : launch_o4<PUSHER, false, true>( p, a, ntiles );
    return scratch ? launch_o4<PUSHER, true, false>( p, a, ntiles ) : launch_o4<PUSHER, false, false>( p, a, ntiles );
}

static int launch_o4_pusher( sb200_patch *p, const DynArgs &a, int ntiles, int pusher, bool scratch )
{
    switch( pusher ) {
        case SB200_PUSHER_BORIS: return launch_o4_flags<SB200_PUSHER_BORIS>( p, a, ntiles, scratch );
        case SB200_PUSHER_VAY: return launch_o4_flags<SB200_PUSHER_VAY>( p, a, ntiles, scratch );
        default: return launch_o4_flags<SB200_PUSHER_HIGUERACARY>( p, a, ntiles, scratch );
    }
}

template<int ORDER>
static int launch_cg_pusher( sb200_patch *p, const DynArgs &a, int ntiles, int pusher, bool scratch )
{
    switch( pusher ) {
        case SB200_PUSHER_BORIS: return scratch ? launch_cg<ORDER, SB200_PUSHER_BORIS, true>( p, a, ntiles ) : launch_cg<ORDER, SB200_PUSHER_BORIS, false>( p, a, ntiles );
        case SB200_PUSHER_VAY: return scratch ? launch_cg<ORDER, SB200_PUSHER_VAY, true>( p, a, ntiles ) : launch_cg<ORDER, SB200_PUSHER_VAY, false>( p, a, ntiles );
        default: return scratch ? launch_cg<ORDER, SB200_PUSHER_HIGUERACARY, true>( p, a, ntiles ) : launch_cg<ORDER, SB200_PUSHER_HIGUERACARY, false>( p, a, ntiles );
    }
}

int launch_dynamics( sb200_patch *p, int ispec, int flags )
{
    SpeciesDev &s = p->sp[ispec];
    SB200_CUDA( cudaMemsetAsync( p->leave_counts + 8*ispec, 0, 8*sizeof( int ), p->stream ) );
    SB200_CUDA( cudaMemsetAsync( s.count, 0, ( p->ncells+1 )*sizeof( int ), p->stream ) );
    s.count_valid = true;
    if( s.n == 0 ) return 0;
    const GridDev &g = p->gd;
    const bool scratch = ( flags & SB200_DYN_KEEP_SCRATCH ) != 0;
    if( scratch && p->sc_cap < s.n ) {
        void *old[] = { p->sc_E, p->sc_B, p->sc_invgf, p->sc_delta, p->sc_iold };
        for( void *m : old ) if( m ) cudaFree( m );
        p->sc_cap = 0;
        SB200_CUDA( cudaMalloc( &p->sc_E, 3*s.n*sizeof( double ) ) );
        SB200_CUDA( cudaMalloc( &p->sc_B, 3*s.n*sizeof( double ) ) );
        SB200_CUDA( cudaMalloc( &p->sc_invgf, s.n*sizeof( double ) ) );
        SB200_CUDA( cudaMalloc( &p->sc_delta, 3*s.n*sizeof( double ) ) );
        SB200_CUDA( cudaMalloc( &p->sc_iold, 3*s.n*sizeof( int ) ) );
        p->sc_cap = s.n;
    }
#ifdef SB200_AB_KERNELS
    if( materialize( p, ispec ) ) return 1;    // A/B build: the one-thread-per-particle kernel works in place
#endif
    DynArgs a;
    const bool deferred = s.perm_pending;      // read through the pending sort order, write the sorted spare set
    if( deferred ) {
        if( ensure_spare( p, s.cap ) ) return 1;
        for( int c=0; c<7; c++ ) { a.in[c] = s.col[c]; a.col[c] = p->spare.col[c]; }
        a.qin = s.q; a.q = p->spare.q; a.key = p->spare.key; a.perm = s.perm;
        swap_with_spare( p, s );               // from here on s.col is the set the kernel writes
        s.perm_pending = false;
    } else {
        for( int c=0; c<7; c++ ) { a.in[c] = s.col[c]; a.col[c] = s.col[c]; }
        a.qin = s.q; a.q = s.q; a.key = s.key; a.perm = nullptr;
    }
    a.first = s.first;
    const int fid[6] = { SB200_EX, SB200_EY, SB200_EZ, SB200_BXM, SB200_BYM, SB200_BZM };
    for( int c=0; c<6; c++ ) a.F[c] = p->f[fid[c]];
    // diag step: the species' own Jx_s Jy_s Jz_s when it has them (Projector3D2Order.cpp:756-758), else the totals
    for( int c=0; c<3; c++ ) a.J[c] = ( ( flags & SB200_DYN_DIAG_RHO ) && s.fs[c] ) ? s.fs[c] : p->f[SB200_JX+c];
    a.count = s.count;
    a.leave_counts = p->leave_counts + 8*ispec;
    a.leave_idx = s.leave_idx;
    a.leave_cap = ( int )s.leave_cap;
    a.iflags = p->iflags;
    a.sc_E = p->sc_E; a.sc_B = p->sc_B; a.sc_invgf = p->sc_invgf; a.sc_delta = p->sc_delta; a.sc_iold = p->sc_iold;
    a.n = s.n;
    for( int d=0; d<3; d++ ) {       // `remove` acts at the global box sides only (PartBoundCond.cpp:99-245: patch->isBoundary)
        a.bc_remove[2*d]   = s.bc[2*d]   == SB200_PBC_REMOVE && p->gd.pcoord[d] == 0;
        a.bc_remove[2*d+1] = s.bc[2*d+1] == SB200_PBC_REMOVE && p->gd.pcoord[d] == p->gd.npatch[d]-1;
    }
    a.any_remove = 0;
    for( int i=0; i<6; i++ ) a.any_remove |= a.bc_remove[i];
    a.lost = s.d_lost;
    a.one_over_mass = 1.0/s.mass;                      // Pusher.cpp:20
    {
        // |J box entry| <= (particles whose window can reach a node) * max|q w|/V * max(d/dt): a node is reached
        // from (2H+3)^3 cells, each holding at most s.maxcount particles; |C| <= cr and |W| <= 1.
        const int reach = g.order + 3;
        const double dmax = fmax( g.d_ov_dt[0], fmax( g.d_ov_dt[1], g.d_ov_dt[2] ) );
        const double bound = ( double )reach*reach*reach*( double )( s.maxcount > 0 ? s.maxcount : 1 )*s.qwmax*g.inv_cell_volume*dmax;
        int e = 0;
        frexp( bound > 0. ? bound : 1., &e );          // bound < 2^e
        a.jscale = ldexp( 1.0, 61 - e );                // bound*jscale < 2^61
        a.jinv = ldexp( 1.0, e - 61 );
    }
#ifndef SB200_AB_KERNELS
    if( g.order == 2 ) {
        using T = CG<2>::T;
        a.tiles[0] = ( g.ncell[0] + T::TX - 1 )/T::TX;
        a.tiles[1] = ( g.ncell[1] + T::TY - 1 )/T::TY;
        a.tiles[2] = ( g.ncell[2] + T::TZ - 1 )/T::TZ;
        return launch_o2_pusher( p, a, a.tiles[0]*a.tiles[1]*a.tiles[2], s.pusher, scratch );
    }
    using T = CG<4>::T;
    a.tiles[0] = ( g.ncell[0] + T::TX - 1 )/T::TX;
    a.tiles[1] = ( g.ncell[1] + T::TY - 1 )/T::TY;
    a.tiles[2] = ( g.ncell[2] + T::TZ - 1 )/T::TZ;
    return launch_o4_pusher( p, a, a.tiles[0]*a.tiles[1]*a.tiles[2], s.pusher, scratch );
#else
    using T = Tile<2>;                        // same tile footprint for both orders of the general kernel
    a.tiles[0] = ( g.ncell[0] + T::TX - 1 )/T::TX;
    a.tiles[1] = ( g.ncell[1] + T::TY - 1 )/T::TY;
    a.tiles[2] = ( g.ncell[2] + T::TZ - 1 )/T::TZ;
    const int ntiles = a.tiles[0]*a.tiles[1]*a.tiles[2];
    if( g.order == 2 ) return launch_pusher<2>( p, a, ntiles, s.pusher, scratch );
    return launch_pusher<4>( p, a, ntiles, s.pusher, scratch );
#endif
}

} // namespace sb200
