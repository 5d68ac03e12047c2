// DFMA / DADD / DMUL dependent-issue latency and single-warp throughput with k independent chains (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template<int CH>
__global__ void k_chain( double *out, long long *cyc, int iters, double a, double b )
{
    double v[CH];
#pragma unroll
    for( int c=0; c<CH; c++ ) v[c] = threadIdx.x + c;
    long long t0 = clock64();
#pragma unroll 1
    for( int i=0; i<iters; i++ ) {
#pragma unroll
        for( int u=0; u<8; u++ )
#pragma unroll
            for( int c=0; c<CH; c++ ) v[c] = fma( v[c], a, b );
    }
    long long t1 = clock64();
    double s = 0.;
#pragma unroll
    for( int c=0; c<CH; c++ ) s += v[c];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
    if( threadIdx.x == 0 && blockIdx.x == 0 ) *cyc = t1 - t0;
}
template<int CH> void run( int warps )
{
    double *out; long long *cyc, h;
    cudaMalloc( &out, 1024*sizeof( double ) ); cudaMalloc( &cyc, 8 );
    const int iters = 2000;
    k_chain<CH><<<1, 32*warps>>>( out, cyc, iters, 1.0000001, 1e-9 );
    k_chain<CH><<<1, 32*warps>>>( out, cyc, iters, 1.0000001, 1e-9 );
    cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost );
    printf( "warps/SM %2d chains %2d : %.2f cycles per DFMA per warp  (%.2f cycles per chain step)\n", warps, CH, ( double )h/( iters*8.0*CH ), ( double )h/( iters*8.0 ) );
    cudaFree( out ); cudaFree( cyc );
}
int main()
{
    run<1>( 1 ); run<2>( 1 ); run<4>( 1 ); run<8>( 1 ); run<16>( 1 );
    run<1>( 4 ); run<2>( 4 ); run<4>( 4 ); run<8>( 4 );
    run<1>( 16 ); run<2>( 16 ); run<4>( 16 );
    run<1>( 20 ); run<2>( 20 ); run<4>( 20 );
    return 0;
}
