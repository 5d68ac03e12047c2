// embc.cu — Silver-Mueller absorbing / injecting boundary condition on one global box face.
//
// Restates ElectroMagnBC3D_SM (src/ElectroMagnBC/ElectroMagnBC3D_SM.cpp): the coefficients of the
// constructor (:61-75) and the two face sweeps of apply() (:141-376) — the tangential B components on the
// boundary plane are rebuilt from the tangential E on the plane, their own value one plane inside, the normal
// B component and the laser amplitudes injected through the face.  External (static) fields are not on this
// path: B_ext = 0, which leaves the reference's expressions unchanged in value.
// One thread per face point; the arithmetic uses __dmul_rn/__dadd_rn in the reference's left-to-right order,
// so the result is bit-identical to the reference's.
#include "common.cuh"
#include <cmath>

namespace sb200 {

struct SMArgs {
    double *Bt;            // tangential component being rebuilt (B1 or B2)
    const double *Et;      // the tangential E it pairs with (E2 for B1, E1 for B2)
    const double *Bn;      // normal component B0
    const double *db;      // laser amplitudes on the face (device) or nullptr
    long long s0, sj, sk;  // element strides along the normal axis, the face's first and second axis
    long long sn;          // stride of the +1 neighbour of Bn (sj for B1, sk for B2)
    int iB, p0, p1;        // plane written, plane of Et/Bn, plane of Bt read
    int j0, j1, k0, k1;    // ranges of the face indices
    int ldb;               // row length of db
    double cE, beta, gamma, cN1, cN0;   // coefficient of Et (+-Alpha), Beta, Gamma, (Delta|Zeta), (Epsilon|Eta)
};

__global__ void __launch_bounds__( 256 ) k_silver_muller( const __grid_constant__ SMArgs a )
{
    const int nk = a.k1 - a.k0;
    const int total = ( a.j1 - a.j0 )*nk;
    for( int t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x ) {
        const int j = a.j0 + t / nk, k = a.k0 + t % nk;
        const long long f = j*a.sj + k*a.sk;
        // ElectroMagnBC3D_SM.cpp:221-227 (and the five sibling sweeps): evaluated left to right
        double v = __dmul_rn( a.cE, a.Et[a.p0*a.s0 + f] );
        v = __dadd_rn( v, __dmul_rn( a.beta, a.Bt[a.p1*a.s0 + f] ) );
        v = __dadd_rn( v, __dmul_rn( a.gamma, a.db ? a.db[j*a.ldb + k] : 0. ) );
        v = __dadd_rn( v, __dmul_rn( a.cN1, a.Bn[a.p0*a.s0 + f + a.sn] ) );
        v = __dadd_rn( v, __dmul_rn( a.cN0, a.Bn[a.p0*a.s0 + f] ) );
        a.Bt[a.iB*a.s0 + f] = v;
    }
}

} // namespace sb200

using namespace sb200;

extern "C" int sb200_apply_SM( sb200_patch *p, int i_boundary, const double kvec[3], const int is_boundary[4],
                               const double *db1, const double *db2 )
{
    SB200_CHECK( p && kvec && is_boundary && i_boundary >= 0 && i_boundary < 6, "sb200_apply_SM: bad arguments" );
    const GridDev &g = p->gd;
    const int axis0 = i_boundary/2, axis1 = axis0 == 0 ? 1 : 0, axis2 = axis0 == 2 ? 1 : 2;
    const int side = i_boundary % 2;
    // ElectroMagnBC3D_SM::apply acts only where patch->isBoundary( i_boundary_ ) (:143)
    if( side == 0 ? g.pcoord[axis0] != 0 : g.pcoord[axis0] != g.npatch[axis0]-1 ) return 0;
    SB200_CUDA( cudaSetDevice( p->device ) );
    const double sign = ( double )side*2 - 1.;
    const double dtd[3] = { g.dt/g.cell[0], g.dt/g.cell[1], g.dt/g.cell[2] };        // ElectroMagnBC.cpp:33
    // constructor, ElectroMagnBC3D_SM.cpp:61-75
    const double Knorm = sqrt( kvec[0]*kvec[0] + kvec[1]*kvec[1] + kvec[2]*kvec[2] );
    SB200_CHECK( Knorm > 0., "sb200_apply_SM: null incidence vector" );
    const double omega = 1.;
    const double k0 = omega*kvec[axis0] / Knorm, k1 = omega*kvec[axis1] / Knorm, k2 = omega*kvec[axis2] / Knorm;
    const double factor = 1.0 / ( k0 - sign * dtd[axis0] );
    const double Alpha = 2.0 * factor;
    const double Beta = - ( k0 + sign * dtd[axis0] ) * factor;
    const double Gamma = 4.0 * k0 * factor;
    const double Delta = - ( k1 + dtd[axis1] ) * factor;
    const double Epsilon = - ( k1 - dtd[axis1] ) * factor;
    const double Zeta = - ( k2 + dtd[axis2] ) * factor;
    const double Eta = - ( k2 - dtd[axis2] ) * factor;
    // plane indices, ElectroMagnBC3D_SM.cpp:24-34 and :169-171
    const int iB_0 = side == 0 ? 0 : g.p[axis0] - 1;          // iB_[axis0_]: plane of the normal component / tangential E
    const int iB_t = side == 0 ? 0 : g.d[axis0] - 1;          // iB_[axis1_] = iB_[axis2_]: plane of the tangential components
    const int p0 = iB_0, p1 = iB_t - ( int )sign, iB1 = iB_t;
    const long long str[3] = { g.sx, g.sy, 1 };
    const int n1p = g.p[axis1], n1d = g.d[axis1], n2p = g.p[axis2], n2d = g.d[axis2];
    double *E[3] = { p->f[SB200_EX], p->f[SB200_EY], p->f[SB200_EZ] };
    double *B[3] = { p->f[SB200_BX], p->f[SB200_BY], p->f[SB200_BZ] };
    // laser amplitudes: host -> staging
    const size_t nb1 = ( size_t )n1p*n2d, nb2 = ( size_t )n1d*n2p;
    double *d1 = nullptr, *d2 = nullptr;
    if( db1 || db2 ) {
        if( ensure_stage( p, nb1 + nb2 ) ) return 1;
        if( db1 ) { d1 = p->stage; SB200_CUDA( cudaMemcpyAsync( d1, db1, nb1*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) ); }
        if( db2 ) { d2 = p->stage + nb1; SB200_CUDA( cudaMemcpyAsync( d2, db2, nb2*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) ); }
    }
    SMArgs a;
    a.s0 = str[axis0]; a.sj = str[axis1]; a.sk = str[axis2];
    a.iB = iB1; a.p0 = p0; a.p1 = p1;
    a.beta = Beta; a.gamma = Gamma;
    // B1 (:203-261): sign of Alpha is + on x and z faces, - on y faces
    a.Bt = B[axis1]; a.Et = E[axis2]; a.Bn = B[axis0]; a.db = d1; a.ldb = n2d;
    a.sn = a.sj;
    a.cE = axis0 == 1 ? -Alpha : Alpha; a.cN1 = Delta; a.cN0 = Epsilon;
    a.j0 = is_boundary[0]; a.j1 = n1p - is_boundary[1]; a.k0 = is_boundary[2]; a.k1 = n2d - is_boundary[3];
    if( a.j1 > a.j0 && a.k1 > a.k0 ) {
        const int total = ( a.j1 - a.j0 )*( a.k1 - a.k0 );
        k_silver_muller<<<( total + 255 )/256, 256, 0, p->stream>>>( a );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
    }
    // B2 (:283-341): sign of Alpha is - on x and z faces, + on y faces
    a.Bt = B[axis2]; a.Et = E[axis1]; a.db = d2; a.ldb = n2p;
    a.sn = a.sk;
    a.cE = axis0 == 1 ? Alpha : -Alpha; a.cN1 = Zeta; a.cN0 = Eta;
    a.j0 = is_boundary[0]; a.j1 = n1d - is_boundary[1]; a.k0 = is_boundary[2]; a.k1 = n2p - is_boundary[3];
    if( a.j1 > a.j0 && a.k1 > a.k0 ) {
        const int total = ( a.j1 - a.j0 )*( a.k1 - a.k0 );
        k_silver_muller<<<( total + 255 )/256, 256, 0, p->stream>>>( a );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
    }
    if( db1 || db2 ) SB200_CUDA( cudaStreamSynchronize( p->stream ) );       // the host arrays may be reused by the caller
    return 0;
}
