// Shared-memory accumulation microbenchmark: how many SM cycles does one lane-level accumulate cost?
//   mode 0: atomicAdd(unsigned long long) distinct addresses per lane (stride 1)
//   mode 1: atomicAdd(unsigned long long) all lanes of a warp on the same address
//   mode 2: atomicAdd(double) distinct addresses (CAS loop)
//   mode 3: atomicAdd(double) same address per group of 8 lanes
//   mode 4: plain LDS + DADD + STS (no atomic), distinct addresses
//   mode 5: atomicAdd(unsigned long long) same address per group of 8 lanes
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE>
__global__ void k( double *out, int iters )
{
    __shared__ unsigned long long s[4096];
    for( int i=threadIdx.x; i<4096; i+=blockDim.x ) s[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *sd = reinterpret_cast<double *>( s );
    for( int it=0; it<iters; it++ ) {
#pragma unroll
        for( int u=0; u<8; u++ ) {
            const int base = ( w*401 + it*37 + u*67 ) & 2047;
            if( MODE == 0 ) atomicAdd( &s[base + lane], 3ull );
            if( MODE == 1 ) atomicAdd( &s[base], 3ull );
            if( MODE == 2 ) atomicAdd( &sd[base + lane], 1.5 );
            if( MODE == 3 ) atomicAdd( &sd[base + ( lane >> 3 )*9], 1.5 );
            if( MODE == 4 ) sd[base + lane] += 1.5;
            if( MODE == 5 ) atomicAdd( &s[base + ( lane >> 3 )*9], 3ull );
            unsigned *s32 = reinterpret_cast<unsigned *>( s );
            if( MODE == 6 ) atomicAdd( &s32[2*( base + lane )], 3u );
            if( MODE == 7 ) atomicAdd( &s32[2*( base + ( lane >> 3 )*9 )], 3u );
            if( MODE == 8 ) atomicAdd( &s32[2*base], 3u );
            if( MODE >= 9 ) {   // 64-bit add as two native 32-bit atomics with carry
                const int idx = MODE == 9 ? base + lane : MODE == 10 ? base + ( lane >> 3 )*9 : base;
                const unsigned lo = 0xfffffff0u + lane, hi = 1u;
                const unsigned old = atomicAdd( &s32[2*idx], lo );
                const unsigned carry = ( old + lo ) < old ? 1u : 0u;
                atomicAdd( &s32[2*idx+1], hi + carry );
            }
        }
    }
    __syncthreads();
    double acc = 0;
    for( int i=threadIdx.x; i<4096; i+=blockDim.x ) acc += ( double )s[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = acc;
}
template<int MODE> void run( const char *name )
{
    double *d; cudaMalloc( &d, 148*4*256*sizeof( double ) );
    cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
    const int iters = 4096;
    k<MODE><<<148*4, 256>>>( d, 16 );
    cudaEventRecord( e0 );
    k<MODE><<<148*4, 256>>>( d, iters );
    cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
    float ms; cudaEventElapsedTime( &ms, e0, e1 );
    // lane-ops per SM: 4 blocks * 256 threads * iters * 8
    const double laneops = 4.*256*iters*8;
    printf( "%-52s %8.3f ms  %6.2f SM-cycles per lane-op @1.9GHz\n", name, ms, ms*1e-3*1.9e9/laneops );
    cudaFree( d );
}
int main()
{
    run<0>( "u64 atomicAdd, distinct addresses" );
    run<1>( "u64 atomicAdd, one address per warp" );
    run<5>( "u64 atomicAdd, one address per 8 lanes" );
    run<2>( "double atomicAdd (CAS), distinct addresses" );
    run<3>( "double atomicAdd (CAS), one address per 8 lanes" );
    run<4>( "plain LDS+DADD+STS, distinct addresses" );
    run<6>( "u32 atomicAdd (native), distinct addresses" );
    run<7>( "u32 atomicAdd (native), one address per 8 lanes" );
    run<8>( "u32 atomicAdd (native), one address per warp" );
    run<9>( "2x u32 carry add, distinct addresses" );
    run<10>( "2x u32 carry add, one address per 8 lanes" );
    run<11>( "2x u32 carry add, one address per warp" );
    return 0;
}
