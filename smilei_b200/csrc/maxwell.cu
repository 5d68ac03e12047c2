// maxwell.cu — Yee FDTD on the padded field layout.
//
//  k_ampere          : MA_Solver3D_norm::operator()  (src/ElectroMagnSolver/MA_Solver3D_norm.cpp:18-115)
//  k_faraday_center  : ElectroMagn3D::saveMagneticFields (ElectroMagn3D.cpp:1048-1118)
//                      + MF_Solver3D_Yee::operator()  (src/ElectroMagnSolver/MF_Solver3D_Yee.cpp:18-111)
//                      + ElectroMagn3D::centerMagneticFields (ElectroMagn3D.cpp:1191-1293) on the
//                        points no halo exchange overwrites
//  k_center_shell    : centerMagneticFields on the exchanged ghost planes (after the B halo)
//
// Algorithmic traffic: E pass reads 3E+3B+3J, writes 3E (96 B/cell); B pass reads 3E+3B,
// writes 3B+3B_m (96 B/cell) -> 192 B per cell-update, against 288 B for the reference's
// four separate sweeps.  HBM-bound; arithmetic uses __dmul_rn/__dadd_rn so that no FMA
// contraction happens and results are bit-identical to the reference's expression order.
#include "common.cuh"

namespace sb200 {

// Two k-consecutive points per thread (16-B loads/stores; AZ is a multiple of 16).
__global__ void __launch_bounds__( 256 ) k_ampere( GridDev g,
        double *__restrict__ Ex, double *__restrict__ Ey, double *__restrict__ Ez,
        const double *__restrict__ Bx, const double *__restrict__ By, const double *__restrict__ Bz,
        const double *__restrict__ Jx, const double *__restrict__ Jy, const double *__restrict__ Jz )
{
    const int halfz = g.az >> 1;
    const unsigned total = ( unsigned )( g.ax*g.ay )*( unsigned )halfz;      // < 2^32 (checked by launch_maxwell): 32-bit divisions, the 64-bit ones are emulated
    const double mdt = -g.dt;
    const double ddx = g.dt_ov_d[0], ddy = g.dt_ov_d[1], ddz = g.dt_ov_d[2];
    for( unsigned t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x ) {
        const int kh = ( int )( t % ( unsigned )halfz );
        const unsigned r = t / ( unsigned )halfz;
        const int j = ( int )( r % g.ay );
        const int i = ( int )( r / g.ay );
        const int k = kh*2;
        const long long idx = i*g.sx + j*g.sy + k;
        const bool ip = i < g.p[0], jp = j < g.p[1];
        if( k >= g.d[2] ) continue;
        // values at k, k+1 and (for the z-differences) k+2
        const double2 bx = *reinterpret_cast<const double2 *>( Bx+idx );
        const double2 by = *reinterpret_cast<const double2 *>( By+idx );
        const double2 bz = *reinterpret_cast<const double2 *>( Bz+idx );
        const bool k2ok = k+2 < g.az;
        const double bx2 = k2ok ? Bx[idx+2] : 0., by2 = k2ok ? By[idx+2] : 0.;
        // Ex^(d,p,p):  Ex += -dt*Jx + dt/dy*(Bz[j+1]-Bz) - dt/dz*(By[k+1]-By)
        if( jp ) {
            const double2 bzj = *reinterpret_cast<const double2 *>( Bz+idx+g.sy );
            double2 e = *reinterpret_cast<double2 *>( Ex+idx );
            const double2 jx = *reinterpret_cast<const double2 *>( Jx+idx );
            if( k < g.p[2] )
                e.x = __dadd_rn( e.x, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jx.x ), __dmul_rn( ddy, __dadd_rn( bzj.x, -bz.x ) ) ), -__dmul_rn( ddz, __dadd_rn( by.y, -by.x ) ) ) );
            if( k+1 < g.p[2] )
                e.y = __dadd_rn( e.y, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jx.y ), __dmul_rn( ddy, __dadd_rn( bzj.y, -bz.y ) ) ), -__dmul_rn( ddz, __dadd_rn( by2, -by.y ) ) ) );
            *reinterpret_cast<double2 *>( Ex+idx ) = e;
        }
        // Ey^(p,d,p):  Ey += -dt*Jy - dt/dx*(Bz[i+1]-Bz) + dt/dz*(Bx[k+1]-Bx)
        if( ip ) {
            const double2 bzi = *reinterpret_cast<const double2 *>( Bz+idx+g.sx );
            double2 e = *reinterpret_cast<double2 *>( Ey+idx );
            const double2 jy = *reinterpret_cast<const double2 *>( Jy+idx );
            if( k < g.p[2] )
                e.x = __dadd_rn( e.x, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jy.x ), -__dmul_rn( ddx, __dadd_rn( bzi.x, -bz.x ) ) ), __dmul_rn( ddz, __dadd_rn( bx.y, -bx.x ) ) ) );
            if( k+1 < g.p[2] )
                e.y = __dadd_rn( e.y, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jy.y ), -__dmul_rn( ddx, __dadd_rn( bzi.y, -bz.y ) ) ), __dmul_rn( ddz, __dadd_rn( bx2, -bx.y ) ) ) );
            *reinterpret_cast<double2 *>( Ey+idx ) = e;
        }
        // Ez^(p,p,d):  Ez += -dt*Jz + dt/dx*(By[i+1]-By) - dt/dy*(Bx[j+1]-Bx)
        if( ip && jp ) {
            const double2 byi = *reinterpret_cast<const double2 *>( By+idx+g.sx );
            const double2 bxj = *reinterpret_cast<const double2 *>( Bx+idx+g.sy );
            double2 e = *reinterpret_cast<double2 *>( Ez+idx );
            const double2 jz = *reinterpret_cast<const double2 *>( Jz+idx );
            e.x = __dadd_rn( e.x, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jz.x ), __dmul_rn( ddx, __dadd_rn( byi.x, -by.x ) ) ), -__dmul_rn( ddy, __dadd_rn( bxj.x, -bx.x ) ) ) );
            if( k+1 < g.d[2] )
                e.y = __dadd_rn( e.y, __dadd_rn( __dadd_rn( __dmul_rn( mdt, jz.y ), __dmul_rn( ddx, __dadd_rn( byi.y, -by.y ) ) ), -__dmul_rn( ddy, __dadd_rn( bxj.y, -bx.y ) ) ) );
            *reinterpret_cast<double2 *>( Ez+idx ) = e;
        }
    }
}

__device__ __forceinline__ bool in_shell( int v, int o, int dim ) { return v < o || v >= dim-o; }

__global__ void __launch_bounds__( 256 ) k_faraday_center( GridDev g,
        const double *__restrict__ Ex, const double *__restrict__ Ey, const double *__restrict__ Ez,
        double *__restrict__ Bx, double *__restrict__ By, double *__restrict__ Bz,
        double *__restrict__ Bxm, double *__restrict__ Bym, double *__restrict__ Bzm )
{
    const int halfz = g.az >> 1;
    const unsigned total = ( unsigned )( g.ax*g.ay )*( unsigned )halfz;      // < 2^32 (checked by launch_maxwell): 32-bit divisions, the 64-bit ones are emulated
    const double ddx = g.dt_ov_d[0], ddy = g.dt_ov_d[1], ddz = g.dt_ov_d[2];
    for( unsigned t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x ) {
        const int kh = ( int )( t % ( unsigned )halfz );
        const unsigned r = t / ( unsigned )halfz;
        const int j = ( int )( r % g.ay );
        const int i = ( int )( r / g.ay );
        const int k = kh*2;
        if( k >= g.d[2] ) continue;
        const long long idx = i*g.sx + j*g.sy + k;
        const bool ip = i < g.p[0], jp = j < g.p[1];
        const bool i_int = i >= 1 && i < g.d[0]-1, j_int = j >= 1 && j < g.d[1]-1;
        const bool i_sh = in_shell( i, g.o[0], g.d[0] ), j_sh = in_shell( j, g.o[1], g.d[1] );
        const double2 ex = *reinterpret_cast<const double2 *>( Ex+idx );
        const double2 ey = *reinterpret_cast<const double2 *>( Ey+idx );
        const double2 ez = *reinterpret_cast<const double2 *>( Ez+idx );
        const double exm = k > 0 ? Ex[idx-1] : 0., eym = k > 0 ? Ey[idx-1] : 0.;
        // every operand of the three updates is asked for here, before the first of them is used (the index of a
        // neighbour that does not exist is replaced by the point's own: loaded, never used): behind the conditions of
        // the updates the loads formed three dependent groups per point, each one waiting for a DRAM round trip
        const long long oy = j >= 1 ? g.sy : 0, ox = i >= 1 ? g.sx : 0;
        const double2 ezj = *reinterpret_cast<const double2 *>( Ez+idx-oy );
        const double2 ezi = *reinterpret_cast<const double2 *>( Ez+idx-ox );
        const double2 eyi = *reinterpret_cast<const double2 *>( Ey+idx-ox );
        const double2 exj = *reinterpret_cast<const double2 *>( Ex+idx-oy );
        const double2 bx0 = *reinterpret_cast<const double2 *>( Bx+idx );
        const double2 by0 = *reinterpret_cast<const double2 *>( By+idx );
        const double2 bz0 = *reinterpret_cast<const double2 *>( Bz+idx );
        bool kint[2], ksh[2], kp[2];
        for( int s=0; s<2; s++ ) {
            kint[s] = ( k+s ) >= 1 && ( k+s ) < g.d[2]-1;
            ksh[s]  = in_shell( k+s, g.o[2], g.d[2] );
            kp[s]   = ( k+s ) < g.p[2];
        }
        const bool kd1 = k+1 < g.d[2];
        // Bx^(p,d,d):  Bx += -dt/dy*(Ez-Ez[j-1]) + dt/dz*(Ey-Ey[k-1]),  j in [1,d1-1), k in [1,d2-1)
        if( ip ) {
            double2 b = bx0;
            const double2 old = b;
            if( j_int ) {
                if( kint[0] ) b.x = __dadd_rn( b.x, __dadd_rn( -__dmul_rn( ddy, __dadd_rn( ez.x, -ezj.x ) ), __dmul_rn( ddz, __dadd_rn( ey.x, -eym ) ) ) );
                if( kint[1] ) b.y = __dadd_rn( b.y, __dadd_rn( -__dmul_rn( ddy, __dadd_rn( ez.y, -ezj.y ) ), __dmul_rn( ddz, __dadd_rn( ey.y, -ey.x ) ) ) );
                *reinterpret_cast<double2 *>( Bx+idx ) = b;
            }
            double2 m;
            m.x = ( j_sh || ksh[0] ) ? old.x : __dmul_rn( __dadd_rn( b.x, old.x ), 0.5 );
            m.y = ( j_sh || ksh[1] ) ? old.y : __dmul_rn( __dadd_rn( b.y, old.y ), 0.5 );
            if( !kd1 ) m.y = 0.;
            *reinterpret_cast<double2 *>( Bxm+idx ) = m;
        }
        // By^(d,p,d):  By += -dt/dz*(Ex-Ex[k-1]) + dt/dx*(Ez-Ez[i-1]),  i in [1,d0-1), k in [1,d2-1)
        if( jp ) {
            double2 b = by0;
            const double2 old = b;
            if( i_int ) {
                if( kint[0] ) b.x = __dadd_rn( b.x, __dadd_rn( -__dmul_rn( ddz, __dadd_rn( ex.x, -exm ) ), __dmul_rn( ddx, __dadd_rn( ez.x, -ezi.x ) ) ) );
                if( kint[1] ) b.y = __dadd_rn( b.y, __dadd_rn( -__dmul_rn( ddz, __dadd_rn( ex.y, -ex.x ) ), __dmul_rn( ddx, __dadd_rn( ez.y, -ezi.y ) ) ) );
                *reinterpret_cast<double2 *>( By+idx ) = b;
            }
            double2 m;
            m.x = ( i_sh || ksh[0] ) ? old.x : __dmul_rn( __dadd_rn( b.x, old.x ), 0.5 );
            m.y = ( i_sh || ksh[1] ) ? old.y : __dmul_rn( __dadd_rn( b.y, old.y ), 0.5 );
            if( !kd1 ) m.y = 0.;
            *reinterpret_cast<double2 *>( Bym+idx ) = m;
        }
        // Bz^(d,d,p):  Bz += -dt/dx*(Ey-Ey[i-1]) + dt/dy*(Ex-Ex[j-1]),  i in [1,d0-1), j in [1,d1-1), k in [0,p2)
        if( kp[0] ) {
            double2 b = bz0;
            const double2 old = b;
            if( i_int && j_int ) {
                b.x = __dadd_rn( b.x, __dadd_rn( -__dmul_rn( ddx, __dadd_rn( ey.x, -eyi.x ) ), __dmul_rn( ddy, __dadd_rn( ex.x, -exj.x ) ) ) );
                if( kp[1] ) b.y = __dadd_rn( b.y, __dadd_rn( -__dmul_rn( ddx, __dadd_rn( ey.y, -eyi.y ) ), __dmul_rn( ddy, __dadd_rn( ex.y, -exj.y ) ) ) );
                *reinterpret_cast<double2 *>( Bz+idx ) = b;
            }
            double2 m;
            m.x = ( i_sh || j_sh ) ? old.x : __dmul_rn( __dadd_rn( b.x, old.x ), 0.5 );
            m.y = ( i_sh || j_sh ) ? old.y : __dmul_rn( __dadd_rn( b.y, old.y ), 0.5 );
            if( !kp[1] ) m.y = 0.;
            *reinterpret_cast<double2 *>( Bzm+idx ) = m;
        }
    }
}

// centerMagneticFields on the ghost planes the B exchange may have overwritten:
// B_m (holding B_old there) <- (B + B_m)*0.5.  Only the shell is enumerated: for a component with dual
// dimensions (a,b) the shell is { a in an edge } U { a interior, b in an edge }, an edge being the first and
// last `o` planes of the dual extent; the third dimension is primal and runs over its full extent.
struct ShellGeom {
    int ext[2][3];        // per region: extents of the local box (x,y,z order)
    int edge[2][3];       // per region and dim: 1 = local index enumerates the 2*o edge planes, 0 = interior/full
    int off[2][3];        // per region and dim: offset added to an interior/full local index
    int dimsz[3];         // real extent of the component per dim
    int o[3];
    long long count[2];
};

struct ShellArgs { ShellGeom sg[3]; const double *B[3]; double *M[3]; };

__global__ void __launch_bounds__( 256 ) k_center_shell( const __grid_constant__ ShellArgs sa, long long sx, long long sy )
{
    // everything is read from the constant bank with the block-uniform component index (no per-thread copy of the structs)
    const ShellGeom &sg = sa.sg[blockIdx.y];
    const double *__restrict__ B = sa.B[blockIdx.y];
    double *__restrict__ M = sa.M[blockIdx.y];
    // 32-bit index arithmetic: the shell holds far fewer than 2^31 points and 64-bit divisions are emulated
    const unsigned c0 = ( unsigned )sg.count[0], total = c0 + ( unsigned )sg.count[1];
    for( unsigned t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x ) {
        const int r = t < c0 ? 0 : 1;
        unsigned u = r == 0 ? t : t - c0;
        int loc[3];
        const unsigned e2 = ( unsigned )sg.ext[r][2], e1 = ( unsigned )sg.ext[r][1];
        loc[2] = ( int )( u % e2 ); u /= e2;
        loc[1] = ( int )( u % e1 );
        loc[0] = ( int )( u / e1 );
        int gi[3];
#pragma unroll
        for( int d=0; d<3; d++ )
            gi[d] = sg.edge[r][d] ? ( loc[d] < sg.o[d] ? loc[d] : sg.dimsz[d] - 2*sg.o[d] + loc[d] ) : loc[d] + sg.off[r][d];
        const long long idx = gi[0]*sx + gi[1]*sy + gi[2];
        M[idx] = __dmul_rn( __dadd_rn( B[idx], M[idx] ), 0.5 );
    }
}

// persistent grid: exactly as many CTAs as are resident at once (SMs x occupancy), each striding over the box
template<class K> static int resident_grid( K kern )
{
    int per_sm = 1, dev = 0, sms = 148;
    cudaGetDevice( &dev );
    cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
    cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, kern, 256, 0 );
    return sms*( per_sm > 0 ? per_sm : 1 );
}

int launch_maxwell( sb200_patch *p )
{
    static const int blocks_a = resident_grid( k_ampere ), blocks_f = resident_grid( k_faraday_center );
    SB200_CHECK( ( double )p->gd.ax*p->gd.ay*( p->gd.az/2 ) < 4.0e9, "sb200_maxwell: patch too large for 32-bit point indices" );
    const int blocks = blocks_a;
    k_ampere<<<blocks, 256, 0, p->stream>>>( p->gd, p->f[SB200_EX], p->f[SB200_EY], p->f[SB200_EZ],
            p->f[SB200_BX], p->f[SB200_BY], p->f[SB200_BZ], p->f[SB200_JX], p->f[SB200_JY], p->f[SB200_JZ] );
            sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    k_faraday_center<<<blocks_f, 256, 0, p->stream>>>( p->gd, p->f[SB200_EX], p->f[SB200_EY], p->f[SB200_EZ],
            p->f[SB200_BX], p->f[SB200_BY], p->f[SB200_BZ], p->f[SB200_BXM], p->f[SB200_BYM], p->f[SB200_BZM] );
            sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int launch_center_shell( sb200_patch *p )
{
    const GridDev &g = p->gd;
    ShellArgs sa;
    ShellGeom *sg = sa.sg;
    long long most = 0;
    for( int c=0; c<3; c++ ) {
        // component c (Bx,By,Bz) is primal along c and dual along the two others, da < db
        const int da = c == 0 ? 1 : 0, db = c == 2 ? 1 : 2;
        ShellGeom &q = sg[c];
        for( int d=0; d<3; d++ ) { q.dimsz[d] = d == c ? g.p[d] : g.d[d]; q.o[d] = g.o[d]; }
        for( int r=0; r<2; r++ )
            for( int d=0; d<3; d++ ) { q.ext[r][d] = q.dimsz[d]; q.edge[r][d] = 0; q.off[r][d] = 0; }
        // region 0: a in the edges, b and c full
        q.ext[0][da] = 2*g.o[da]; q.edge[0][da] = 1;
        // region 1: a interior, b in the edges, c full
        q.ext[1][da] = q.dimsz[da] - 2*g.o[da]; q.off[1][da] = g.o[da];
        q.ext[1][db] = 2*g.o[db]; q.edge[1][db] = 1;
        for( int r=0; r<2; r++ ) q.count[r] = ( long long )q.ext[r][0]*q.ext[r][1]*q.ext[r][2];
        if( q.count[0] + q.count[1] > most ) most = q.count[0] + q.count[1];
    }
    long long nb = ( most + 255 )/256;
    if( nb > 148*8 ) nb = 148*8;
    if( nb < 1 ) nb = 1;
    dim3 grid( ( unsigned )nb, 3 );
    for( int c=0; c<3; c++ ) { sa.B[c] = p->f[SB200_BX+c]; sa.M[c] = p->f[SB200_BXM+c]; }
    k_center_shell<<<grid, 256, 0, p->stream>>>( sa, g.sx, g.sy );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

} // namespace sb200
