// How many shared-memory wavefronts does a warp-wide 64-bit / 128-bit load take when many lanes read the same
// address?  (The gather of k_dynamics_* reads a field box with lanes = particles sorted by cell: 2-4 distinct
// addresses per instruction.)  One CTA of 128 threads per pattern; every warp issues N independent loads; the
// printed figure is SM cycles per warp-level load instruction (1.0 = one wavefront each, the pipe's peak).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lds_pattern_bench tools/lds_pattern_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template<int W> __device__ __forceinline__ void ld( unsigned addr, double &acc );
template<> __device__ __forceinline__ void ld<8>( unsigned addr, double &acc )
{
    double v; asm volatile( "ld.volatile.shared.f64 %0, [%1];" : "=d"( v ) : "r"( addr ) ); acc += v;
}
template<> __device__ __forceinline__ void ld<16>( unsigned addr, double &acc )
{
    double v, w; asm volatile( "ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"( v ), "=d"( w ) : "r"( addr ) ); acc += v + w;
}

// pattern p: byte offset of lane l
__device__ int offset_of( int p, int l )
{
    switch( p ) {
    case 0: return 0;                                 // all lanes one address
    case 1: return ( l >> 4 )*8;                      // half warps: A, A+8
    case 2: return ( l & 1 )*8;                       // alternating A, A+8
    case 3: return ( l >> 3 )*8;                      // 4 blocks of 8 lanes: 4 consecutive doubles
    case 4: return ( l >> 4 )*112;                    // half warps 14 doubles apart (next row of an order-4 box)
    case 5: return ( l >> 4 )*8 + ( ( l >> 2 ) & 1 )*112;   // two cells x two rows
    case 6: return l*8;                               // all distinct, consecutive
    case 7: return ( l >> 4 )*16;                     // half warps: A, A+16
    case 8: return ( l >> 3 )*16;                     // 4 blocks: 16 B apart
    case 9: return ( l >> 4 )*1008;                   // half warps one x-plane apart (126 doubles)
    case 10: return ( l % 3 )*8;                      // 3 addresses interleaved
    case 11: return ( l >> 4 )*8 + ( l & 1 )*1008;    // two cells x two planes
    case 12: return ( l >> 2 )*8;                     // 8 blocks of 4 lanes: 8 consecutive doubles
    case 13: return ( l >> 1 )*8;                     // 16 pairs: 16 consecutive doubles
    case 14: return ( l >> 4 )*128;                   // half warps 128 B apart (same banks!)
    case 15: return l*16;                             // stride 2 doubles
    case 16: return l*128;                            // stride 16 doubles: every lane the same banks
    case 17: return ( l & 7 )*128;                    // 8 addresses in the same banks
    case 18: return ( l & 3 )*128;                    // 4 addresses in the same banks
    case 19: return ( l & 1 )*128;                    // 2 addresses in the same banks
    case 20: return ( ( 0x5a3c96e1u >> l ) & 1 )*8;                    // 2 addresses, irregular lanes
    case 21: return ( ( 0x5a3c96e1u >> l ) & 1 )*8 + ( l >> 4 )*16;    // 2 cells (half warps) x 2 irregular
    case 22: return ( ( l >> 2 ) & 1 )*8;                              // blocks of 4 lanes alternating A, A+8
    case 23: return ( l % 5 )*8;
    case 24: return ( l % 3 == 0 )*8;                                  // 2 addresses, period 3
    case 25: return ( ( 0x5a3c96e1u >> l ) & 1 )*8 + ( ( 0x0ff0f00fu >> l ) & 1 )*112;   // irregular in z and y (14 doubles)
    case 26: return ( l / 11 )*8 + ( ( 0x5a3c96e1u >> l ) & 1 )*112;   // 3 cells (11 lanes each) x irregular y
    case 27: return ( l / 11 )*8;                                      // 3 cells of 11 lanes
    case 28: return ( l < 13 ? 0 : 8 );                                // 2 cells: 13 + 19 lanes
    case 29: return ( ( 0x5a3c96e1u >> l ) & 1 )*1008;                 // irregular in x (126 doubles)
    case 30: return l < 16 ? 0 : ( l & 1 )*8;                          // second half alternating
    case 31: return l == 0 ? 8 : 0;                                    // one outlier lane
    case 32: return l < 8 ? 0 : 8;                                     // 8 + 24
    case 33: return l < 12 ? 0 : 8;                                    // 12 + 20
    case 34: return l < 14 ? 0 : 8;                                    // 14 + 18
    case 35: return ( ( l >> 1 ) % 3 )*8;                              // pairs, period 3
    case 36: return ( ( l >> 2 ) % 3 )*8;                              // quads, period 3
    case 37: return ( ( l >> 3 ) % 3 )*8;                              // octets: A, A+8, A+16, A
    case 38: return ( ( 0x5a3c96e1u >> ( l >> 1 ) ) & 1 )*8;           // pairs, irregular
    case 39: return ( ( 0x5a3c96e1u >> ( l >> 2 ) ) & 1 )*8;           // quads, irregular
    case 40: return ( ( 0x6u >> ( l >> 3 ) ) & 1 )*8;                  // octets: A, A+8, A+8, A
    case 41: return l == 31 ? 8 : 0;                                   // last lane outlier
    case 42: return l == 16 ? 8 : 0;                                   // lane 16 outlier
    case 43: return ( l & 15 ) == 0 ? 8 : 0;                           // lanes 0 and 16 outliers
    case 44: return ( l & 15 ) < 5 ? 0 : 8;                            // same irregular split in both halves
    case 45: return ( l & 15 ) % 3 * 8;                                // period 3 in each half, same in both
    default: return 0;
    }
}

template<int W> __global__ void k( int p, int n, long long *cycles, double *sink )
{
    extern __shared__ __align__( 16 ) double sm[];
    for( int i = threadIdx.x; i < 2048; i += blockDim.x ) sm[i] = i;
    __syncthreads();
    const unsigned base = ( unsigned )__cvta_generic_to_shared( sm ) + offset_of( p, threadIdx.x & 31 )*( W/8 ) + ( threadIdx.x >> 5 )*16;
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
    __syncthreads();
    const long long t0 = clock64();
    for( int i = 0; i < n; i += 8 ) {
        ld<W>( base, a0 ); ld<W>( base + 16, a1 ); ld<W>( base + 32, a2 ); ld<W>( base + 48, a3 );
        ld<W>( base + 64, a0 ); ld<W>( base + 80, a1 ); ld<W>( base + 96, a2 ); ld<W>( base + 112, a3 );
    }
    __syncthreads();
    const long long t1 = clock64();
    if( threadIdx.x == 0 ) *cycles = t1 - t0;
    sink[threadIdx.x] = a0 + a1 + a2 + a3;
}

int main()
{
    long long *c; double *s;
    cudaMalloc( &c, 8 ); cudaMalloc( &s, 8*128 );
    const int n = 8192;
    const char *names[] = { "all lanes one address", "half warps A, A+8", "alternating A, A+8", "4 blocks of 8 lanes, consecutive doubles",
        "half warps 14 doubles apart", "2 cells x 2 rows (14 doubles)", "32 distinct consecutive", "half warps A, A+16", "4 blocks 16 B apart",
        "half warps 126 doubles apart", "3 addresses interleaved", "2 cells x 2 planes (126 doubles)", "8 blocks of 4 lanes", "16 pairs", "half warps 128 B apart", "stride 2 doubles", "stride 16 doubles (same banks)", "8 addresses, same banks", "4 addresses, same banks", "2 addresses, same banks", "2 addresses irregular lanes", "2 half-warp cells x 2 irregular", "blocks of 4 alternating", "l % 5", "2 addresses period 3", "irregular z and y", "3 cells x irregular y", "3 cells of 11 lanes", "2 cells 13+19 lanes", "irregular x planes", "second half alternating", "lane 0 outlier", "8 + 24", "12 + 20", "14 + 18", "pairs period 3", "quads period 3", "octets A,A+8,A+16,A", "pairs irregular", "quads irregular", "octets A,A+8,A+8,A", "lane 31 outlier", "lane 16 outlier", "lanes 0 and 16 outliers", "5+11 split in both halves", "period 3 in each half (same)" };
    for( int NT = 512; NT <= 512; NT *= 4 )
    for( int w = 8; w <= 8; w += 8 )
        for( int p = 30; p < 46; p++ ) {
            long long h = 0;
            for( int rep = 0; rep < 2; rep++ ) {
                if( w == 8 ) k<8><<<1, NT, 16384>>>( p, n, c, s ); else k<16><<<1, NT, 16384>>>( p, n, c, s );
                cudaMemcpy( &h, c, 8, cudaMemcpyDeviceToHost );
            }
            printf( "%3d threads LDS.%-3d %-45s %.2f cycles per warp load\n", NT, w*8, names[p], ( double )h/( ( NT/32 )*( double )n ) );
        }
    printf( "%s\n", cudaGetErrorString( cudaGetLastError() ) );
    return 0;
}
