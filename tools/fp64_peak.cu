// DFMA issue-rate microbenchmark: measures the FP64 FMA peak of the device (the number
// SURVEY.md §7 asks for; MEASURED_PEAKS.json only holds HBM and bf16).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k( double *out, double a, double b, int iters )
{
    double x0 = threadIdx.x, x1 = x0+1, x2 = x0+2, x3 = x0+3, x4 = x0+4, x5 = x0+5, x6 = x0+6, x7 = x0+7;
    for( int i=0; i<iters; i++ ) {
        x0 = fma( x0, a, b ); x1 = fma( x1, a, b ); x2 = fma( x2, a, b ); x3 = fma( x3, a, b );
        x4 = fma( x4, a, b ); x5 = fma( x5, a, b ); x6 = fma( x6, a, b ); x7 = fma( x7, a, b );
    }
    out[blockIdx.x*blockDim.x+threadIdx.x] = x0+x1+x2+x3+x4+x5+x6+x7;
}
int main()
{
    double *d; cudaMalloc( &d, 148*16*256*sizeof( double ) );
    cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
    const int iters = 1<<16;
    for( int rep=0; rep<3; rep++ ) {
        cudaEventRecord( e0 );
        k<<<148*16, 256>>>( d, 1.0000001, 1e-9, iters );
        cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
        float ms; cudaEventElapsedTime( &ms, e0, e1 );
        double fma = 148.*16*256*8.*iters;
        printf( "DFMA: %.2f ms  %.2f TFMA/s  (%.2f TFLOP/s fp64)\n", ms, fma/ms/1e9, 2*fma/ms/1e9 );
    }
    return 0;
}
