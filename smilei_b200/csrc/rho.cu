// rho.cu — diag-step charge deposit (SB200_DYN_DIAG_RHO).
//
// Projector3D2Order::currentsAndDensity (src/Projector/Projector3D2Order.cpp:509-519) and its order-4
// counterpart add  rho[i][j][k] += charge_weight * Sx1[i]*Sy1[j]*Sz1[k]  with S1 the shape at the NEW
// position, i.e. on the (ORDER+1)^3 primal nodes around the particle's new primal node.  Only steps with
// field diagnostics run it, so it is a plain one-thread-per-particle kernel with red.global.add.f64; the
// currents of such a step are deposited by the normal fused kernel (same values: currentsAndDensity and
// currents differ only in loop form).
#include "common.cuh"

namespace sb200 {

template<int ORDER> struct RhoShape;
template<> struct RhoShape<2> {
    __device__ static __forceinline__ void w( double d, double *c )
    {
        const double d2 = d*d;
        c[0] = 0.5*( d2 - d + 0.25 ); c[1] = 0.75 - d2; c[2] = 0.5*( d2 + d + 0.25 );
    }
};
template<> struct RhoShape<4> {
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

template<int ORDER>
__global__ void __launch_bounds__( 256 ) k_deposit_rho( GridDev g, const double *__restrict__ x, const double *__restrict__ y,
        const double *__restrict__ z, const double *__restrict__ w, const short *__restrict__ q, size_t n, double *__restrict__ rho )
{
    constexpr int NW = ORDER+1, H = ORDER/2;
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x ) {
        const double pos[3] = { x[i], y[i], z[i] };
        double S[3][NW];
        int base[3];
#pragma unroll
        for( int d=0; d<3; d++ ) {
            const double pn = pos[d]*g.dxi[d];
            const int ipn = ( int )round( pn );
            RhoShape<ORDER>::w( pn - ( double )ipn, S[d] );
            base[d] = ipn - g.begin[d] - H;
        }
        const double charge_weight = g.inv_cell_volume*( double )q[i]*w[i];
#pragma unroll
        for( int a=0; a<NW; a++ ) {
            const int gi = base[0]+a;
            if( gi < 0 || gi >= g.p[0] ) continue;
#pragma unroll
            for( int b=0; b<NW; b++ ) {
                const int gj = base[1]+b;
                if( gj < 0 || gj >= g.p[1] ) continue;
#pragma unroll
                for( int c=0; c<NW; c++ ) {
                    const int gk = base[2]+c;
                    if( gk < 0 || gk >= g.p[2] ) continue;
                    atomicAdd( rho + gi*g.sx + gj*g.sy + gk, charge_weight*S[0][a]*S[1][b]*S[2][c] );
                }
            }
        }
    }
}

int launch_rho( sb200_patch *p, int ispec )
{
    if( materialize( p, ispec ) ) return 1;
    SpeciesDev &s = p->sp[ispec];
    if( s.n == 0 ) return 0;
    const unsigned blocks = ( unsigned )( ( s.n + 255 )/256 < 148*16 ? ( s.n + 255 )/256 : 148*16 );
    double *rho = s.fs[3] ? s.fs[3] : p->f[SB200_RHO];        // the species' own rho_s when it has one (Projector3D2Order.cpp:759)
    if( p->gd.order == 2 ) k_deposit_rho<2><<<blocks, 256, 0, p->stream>>>( p->gd, s.col[0], s.col[1], s.col[2], s.col[6], s.q, s.n, rho );
    else k_deposit_rho<4><<<blocks, 256, 0, p->stream>>>( p->gd, s.col[0], s.col[1], s.col[2], s.col[6], s.q, s.n, rho );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

} // namespace sb200
