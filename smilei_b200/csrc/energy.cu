// energy.cu — the parity observables of DiagnosticScalar:
//   Ukin_s = mass * sum_p w (sqrt(1+p^2) - 1)        (src/Diagnostic/DiagnosticScalar.cpp:497-510)
//   Uelm   = sum_f 0.5*cell_volume*norm2(f), f in Ex,Ey,Ez,Bx_m,By_m,Bz_m over the
//            non-duplicated window istart/bufsize     (:658-691; Field3D::norm2 src/Field/Field3D.cpp:230-250;
//            window src/ElectroMagn/ElectroMagn3D.cpp:190-229)
// Two-stage deterministic reduction (fixed grid, fixed tree), double precision.
#include "common.cuh"
#include <vector>

namespace sb200 {

constexpr int RED_BLOCKS = 592, RED_T = 256;

__device__ __forceinline__ double block_sum( double v )
{
    __shared__ double sh[RED_T/32];
#pragma unroll
    for( int d=16; d>0; d>>=1 ) v += __shfl_down_sync( 0xffffffffu, v, d );
    __syncthreads();
    if( ( threadIdx.x & 31 ) == 0 ) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    if( threadIdx.x == 0 ) for( int i=0; i<RED_T/32; i++ ) t += sh[i];
    return t;
}

__global__ void __launch_bounds__( RED_T ) k_ukin( const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
        const double *__restrict__ w, size_t n, double *__restrict__ partial )
{
    double acc = 0.;
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x ) {
        const double a = px[i], b = py[i], c = pz[i];
        const double gamma = sqrt( 1. + a*a + b*b + c*c );
        acc += w[i]*( gamma - 1.0 );
    }
    const double t = block_sum( acc );
    if( threadIdx.x == 0 ) partial[blockIdx.x] = t;
}

struct Win { int s[3], e[3]; };

__global__ void __launch_bounds__( RED_T ) k_norm2( const double *__restrict__ f, Win w, long long sx, long long sy, double *__restrict__ partial )
{
    const int nk = w.e[2]-w.s[2], nj = w.e[1]-w.s[1], ni = w.e[0]-w.s[0];
    const long long total = ( long long )ni*nj*nk;
    double acc = 0.;
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x ) {
        const int k = w.s[2] + ( int )( t % nk );
        const long long r = t / nk;
        const int j = w.s[1] + ( int )( r % nj );
        const int i = w.s[0] + ( int )( r / nj );
        const double v = f[i*sx + j*sy + k];
        acc += v*v;
    }
    const double t = block_sum( acc );
    if( threadIdx.x == 0 ) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__( RED_T ) k_final( const double *__restrict__ partial, int n, double scale, double *__restrict__ out )
{
    double acc = 0.;
    for( int i = threadIdx.x; i < n; i += blockDim.x ) acc += partial[i];
    const double t = block_sum( acc );
    if( threadIdx.x == 0 ) *out = t*scale;
}

int launch_energy( sb200_patch *p, double *ukin, double *uelm )
{
    const GridDev &g = p->gd;
    double *partial = p->red;            // RED_BLOCKS entries
    double *res = p->red + 1024;         // up to 64 species + 6 fields
    std::vector<double> host( p->nspec + 6, 0. );
    if( ukin ) {
        for( int s=0; s<p->nspec; s++ ) {
            if( materialize( p, s ) ) return 1;       // the sum tree is defined on the sorted order
            SpeciesDev &S = p->sp[s];
            k_ukin<<<RED_BLOCKS, RED_T, 0, p->stream>>>( S.col[3], S.col[4], S.col[5], S.col[6], S.n, partial );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
            k_final<<<1, RED_T, 0, p->stream>>>( partial, RED_BLOCKS, S.mass, res+s );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
        }
    }
    if( uelm ) {
        const int ids[6] = { SB200_EX, SB200_EY, SB200_EZ, SB200_BXM, SB200_BYM, SB200_BZM };
        for( int f=0; f<6; f++ ) {
            Win w;
            for( int i=0; i<3; i++ ) {
                const int isDual = field_dual( ids[f], i );
                int istart = g.o[i] + ( g.pcoord[i] != 0 ? 1 : 0 );
                int bufsize = g.n[i] + 1 + isDual;
                if( g.npatch[i] != 1 ) {
                    if( !isDual && g.pcoord[i] != 0 ) bufsize--;
                    else if( isDual ) {
                        bufsize--;
                        if( g.pcoord[i] != 0 && g.pcoord[i] != g.npatch[i]-1 ) bufsize--;
                    }
                }
                w.s[i] = istart; w.e[i] = istart+bufsize;
            }
            k_norm2<<<RED_BLOCKS, RED_T, 0, p->stream>>>( p->f[ids[f]], w, g.sx, g.sy, partial );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
            k_final<<<1, RED_T, 0, p->stream>>>( partial, RED_BLOCKS, 0.5*g.cell_volume, res+p->nspec+f );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
        }
    }
    SB200_CUDA( cudaMemcpyAsync( host.data(), res, host.size()*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    if( ukin ) for( int s=0; s<p->nspec; s++ ) ukin[s] = host[s];
    if( uelm ) { double u = 0.; for( int f=0; f<6; f++ ) u += host[p->nspec+f]; *uelm = u; }
    return 0;
}

} // namespace sb200
