// sort.cu — cell keys, histogram, exclusive scan and the stable counting sort of the SoA.
//
// Restates SpeciesV::computeParticleCellKeys (src/Species/SpeciesV.cpp:766-855) and the
// result of SpeciesV::sortParticles (:599-762): particles with key<0 (tagged leavers) are
// dropped, the others are ordered by cell key; first_index/last_index are the prefix sum of
// the per-cell counts (:645-652).  Within a cell the order is the STABLE one (ascending
// original index) — the canonical order of this build; the reference's in-place cycle sort
// leaves an algorithm-dependent order within a cell.
//
// Pipeline (all on p->stream):
//   k_keys_hist   key (if not already valid) + count[key]++                 4(+24) B/particle read
//   scan          first = exclusive_scan(count)                             ncells ints
//   k_scatter_idx slot = first[key] + atomicAdd(cursor[key]) ; perm[slot]=i  8 B/particle
//   k_cell_sort   ascending insertion sort of each cell's run of perm        (restores stability)
//   k_gather      out[c][j] = in[c][perm[j]] for the 9 columns               124 B/particle
#include "common.cuh"
#include <cstring>

namespace sb200 {

__device__ __forceinline__ int cell_key_of( const GridDev &g, double x, double y, double z )
{
    // SpeciesV.cpp:806-812 — same int arithmetic, round() = half away from zero
    int key = ( int )( round( x*g.dxi[0] ) - g.min_loc_round[0] );
    key *= g.ncell[1];
    key += ( int )( round( y*g.dxi[1] ) - g.min_loc_round[1] );
    key *= g.ncell[2];
    key += ( int )( round( z*g.dxi[2] ) - g.min_loc_round[2] );
    return key;
}

__global__ void __launch_bounds__( 256 ) k_keys_hist( GridDev g, const double *__restrict__ x, const double *__restrict__ y,
        const double *__restrict__ z, int *__restrict__ key, int *__restrict__ count, size_t n, int recompute, int ncells, int *__restrict__ iflags )
{
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x ) {
        int k = key[i];
        if( k < 0 ) continue;
        if( recompute ) {
            k = cell_key_of( g, x[i], y[i], z[i] );
            if( k < 0 || k >= ncells ) {       // a particle outside the patch that nobody tagged
                atomicAdd( &iflags[0], 1 );
                k = -1;
            }
            key[i] = k;
            if( k < 0 ) continue;
        }
        atomicAdd( &count[k], 1 );
    }
}

// ---------------------------------------------------------------- exclusive scan (int)
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_B = SCAN_T*SCAN_E;

__global__ void __launch_bounds__( SCAN_T ) k_scan_block( int *__restrict__ data, int *__restrict__ sums, size_t n )
{
    __shared__ int warp_tot[SCAN_T/32];
    const size_t base = ( size_t )blockIdx.x*SCAN_B + ( size_t )threadIdx.x*SCAN_E;
    int v[SCAN_E], t = 0;
#pragma unroll
    for( int e=0; e<SCAN_E; e++ ) { v[e] = base+e < n ? data[base+e] : 0; t += v[e]; }
    // inclusive warp scan of the thread totals
    int inc = t;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for( int d=1; d<32; d<<=1 ) { int u = __shfl_up_sync( 0xffffffffu, inc, d ); if( lane >= d ) inc += u; }
    if( lane == 31 ) warp_tot[w] = inc;
    __syncthreads();
    int woff = 0;
    for( int i=0; i<w; i++ ) woff += warp_tot[i];
    int run = woff + inc - t;
#pragma unroll
    for( int e=0; e<SCAN_E; e++ ) { if( base+e < n ) data[base+e] = run; run += v[e]; }
    if( threadIdx.x == SCAN_T-1 && sums ) sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__( SCAN_T ) k_scan_add( int *__restrict__ data, const int *__restrict__ sums, size_t n )
{
    const int off = sums[blockIdx.x];
    const size_t base = ( size_t )blockIdx.x*SCAN_B + ( size_t )threadIdx.x*SCAN_E;
#pragma unroll
    for( int e=0; e<SCAN_E; e++ ) if( base+e < n ) data[base+e] += off;
}

static int scan_rec( sb200_patch *p, int *data, size_t n, int *ws )
{
    const size_t nb = ( n + SCAN_B - 1 )/SCAN_B;
    if( nb == 1 ) {
        k_scan_block<<<1, SCAN_T, 0, p->stream>>>( data, nullptr, n );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
        return 0;
    }
    k_scan_block<<<( unsigned )nb, SCAN_T, 0, p->stream>>>( data, ws, n );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    if( scan_rec( p, ws, nb, ws+nb ) ) return 1;
    k_scan_add<<<( unsigned )nb, SCAN_T, 0, p->stream>>>( data, ws, n );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int exclusive_scan_int( sb200_patch *p, int *data, size_t n )
{
    if( n == 0 ) return 0;
    size_t need = 0, m = n;
    while( m > 1 ) { m = ( m + SCAN_B - 1 )/SCAN_B; need += m; if( m == 1 ) break; }
    need += 16;
    if( need > p->blocksums_cap ) {
        if( p->blocksums ) cudaFree( p->blocksums );
        p->blocksums = nullptr; p->blocksums_cap = 0;
        SB200_CUDA( cudaMalloc( &p->blocksums, need*sizeof( int ) ) );
        p->blocksums_cap = need;
    }
    return scan_rec( p, data, n, p->blocksums );
}

// ---------------------------------------------------------------- scatter / per-cell order / gather
// slot of particle i in the run of its new cell.  Lanes of a warp holding the same key take consecutive slots
// in lane (= index) order from ONE atomic on the cell's cursor, so a cell whose particles sit in one warp
// comes out already ordered; k_cell_sort repairs the cells that straddle warps or received movers.
__global__ void __launch_bounds__( 256 ) k_scatter_idx( const int *__restrict__ key, const int *__restrict__ first,
        int *__restrict__ cursor, int *__restrict__ perm, size_t n )
{
    const int lane = threadIdx.x & 31;
    // a warp walks ONE contiguous chunk of the particle list, 32 at a time: its successive atomics on a cursor are
    // served in program order (the shuffle below consumes each result before the next is issued), so the only cells
    // that come out with swapped blocks are the few that straddle two chunks - not every cell that straddles two
    // groups of 32 particles, as with a grid-stride loop
    const size_t nblk = ( n + 31 )/32, nwarp = ( size_t )gridDim.x*( blockDim.x >> 5 );
    const size_t per = ( nblk + nwarp - 1 )/nwarp;
    const size_t b0 = ( blockIdx.x*( size_t )( blockDim.x >> 5 ) + ( threadIdx.x >> 5 ) )*per;
    const size_t b1 = b0 + per < nblk ? b0 + per : nblk;
    // A warp's walk is a chain of dependent latencies (key load, atomic round trip, store).  Four stretches are taken per
    // turn, their atomics in flight together, and the keys of the next turn are loaded before this one's atomics are
    // issued.  (Atomics of one warp on one cursor leave the SM in program order; should two ever be served the other
    // way round, k_cell_sort repairs the run like any other.)
    auto load = [&]( size_t b ) -> int { const size_t i = b*32 + lane; return b < b1 && i < n ? key[i] : -1; };
    constexpr int U = 4;                                   // stretches per turn
    int kn[U];
#pragma unroll
    for( int u=0; u<U; u++ ) kn[u] = load( b0 + u );
    const unsigned below = ( 1u << lane ) - 1u;
    for( size_t b = b0; b < b1; b += U ) {
        int k[U], f[U], base[U], lead[U];
        unsigned peers[U];
#pragma unroll
        for( int u=0; u<U; u++ ) { k[u] = kn[u]; kn[u] = load( b + U + u ); }
#pragma unroll
        for( int u=0; u<U; u++ ) {
            f[u] = k[u] >= 0 ? first[k[u]] : 0;           // independent of the atomics
            peers[u] = __match_any_sync( 0xffffffffu, k[u] );
            lead[u] = __ffs( peers[u] ) - 1;
        }
#pragma unroll
        for( int u=0; u<U; u++ ) {
            base[u] = 0;
            if( lane == lead[u] && k[u] >= 0 ) base[u] = atomicAdd( &cursor[k[u]], __popc( peers[u] ) );
        }
#pragma unroll
        for( int u=0; u<U; u++ ) {
            base[u] = __shfl_sync( 0xffffffffu, base[u], lead[u] );
            if( k[u] >= 0 ) perm[f[u] + base[u] + __popc( peers[u] & below )] = ( int )( ( b+u )*32 + lane );
        }
    }
}

// Restore ascending index order inside every cell's run of perm (the atomic cursor of k_scatter_idx serves
// slots in arbitrary order).  A warp takes 32 consecutive cells = one contiguous stretch of perm, reads it
// coalesced into shared memory and looks for inversions between neighbours of the same cell; particles of a
// cell mostly arrive in order.  Each lane then insertion-sorts its own cell in shared memory (the runs are
// nearly ordered, so this is a handful of moves) and the warp writes the stretch back coalesced.
__global__ void __launch_bounds__( 256 ) k_cell_sort( const int *__restrict__ first, int *__restrict__ perm, int ncells )
{
    constexpr int SW = 1024;                       // entries of a warp's stretch kept in shared memory
    __shared__ int stretch[8][SW + SW/16 + 1];     // skewed: entry p sits at p + p/16, so lanes working on cells of ~16 entries hit different banks
    const int lane = threadIdx.x & 31;
    int *sm = stretch[threadIdx.x >> 5];
    const int nwarps = ( gridDim.x*blockDim.x ) >> 5;
    for( int w = ( blockIdx.x*blockDim.x + threadIdx.x ) >> 5; w*32 < ncells; w += nwarps ) {
        const int c = w*32 + lane;
        const int cb = first[min( c, ncells )], ce = first[min( c+1, ncells )];
        // the warp's 32 cells are taken in `nsub` sub-stretches of 32/nsub cells, so that a sub-stretch fits the
        // shared-memory staging at any number of particles per cell (one pass up to SW entries, e.g. 16 per cell;
        // four passes of ~512 entries at 64 per cell)
        const int total = __shfl_sync( 0xffffffffu, ce, 31 ) - __shfl_sync( 0xffffffffu, cb, 0 );
        int nsub = 1;
        if( total > SW ) while( nsub < 32 && nsub*( SW/2 ) < total ) nsub <<= 1;
        const int G = 32/nsub;
      for( int sub = 0; sub < nsub; sub++ ) {
        const int base = __shfl_sync( 0xffffffffu, cb, sub*G ), end = __shfl_sync( 0xffffffffu, ce, sub*G + G - 1 );
        const bool staged = end - base <= SW;
        const bool member = lane >= sub*G && lane < sub*G + G;      // my cell belongs to this sub-stretch
        bool mybad = false;
        __syncwarp();
        if( staged ) {
            // all the loads of the stretch are issued back to back (nothing below depends on them until the barrier)
            for( int q = lane; q < end - base; q += 32 ) sm[q + ( q >> 4 )] = perm[base + q];
            __syncwarp();
        }
        for( int j0 = base; j0 < end; j0 += 32 ) {
            const int j = j0 + lane;
            int v = 0x7fffffff;
            if( j < end ) { const int q = j - base; v = staged ? sm[q + ( q >> 4 )] : perm[j]; }
            int vn = __shfl_down_sync( 0xffffffffu, v, 1 );
            if( lane == 31 && j+1 < end ) { const int q = j + 1 - base; vn = staged ? sm[q + ( q >> 4 )] : perm[j+1]; }
            else if( lane == 31 ) vn = 0x7fffffff;
            // positions of the chunk that are the LAST entry of a cell (the next entry starts another cell): lane c
            // knows where its own cell ends
            const int last = ce - 1 - j0;
            const unsigned ends = __reduce_or_sync( 0xffffffffu, ( ce > cb && last >= 0 && last < 32 ) ? ( 1u << last ) : 0u );
            const bool inv = j+1 < end && v > vn && !( ( ends >> lane ) & 1u );
            const unsigned invmask = __ballot_sync( 0xffffffffu, inv );
            // does an inversion fall inside my own cell's run?
            const int lo = max( cb - j0, 0 ), hi = min( ce - j0, 32 );
            if( invmask && hi > lo && member ) {
                const unsigned range = ( hi >= 32 ? 0xffffffffu : ( ( 1u << hi ) - 1u ) ) & ~( ( 1u << lo ) - 1u );
                mybad = mybad || ( invmask & range ) != 0u;
            }
        }
        const unsigned bad = __ballot_sync( 0xffffffffu, mybad );
        if( bad == 0 ) continue;
        __syncwarp();
        if( ( bad >> lane ) & 1u ) {
            if( staged ) {
                const int b0 = cb - base, e0 = ce - base;
                int prev = sm[b0 + ( b0 >> 4 )];                       // the largest entry so far = the one just before i
                for( int i = b0+1; i < e0; i++ ) {
                    const int v = sm[i + ( i >> 4 )];
                    if( v > prev ) { prev = v; continue; }             // in order (the common case): nothing to move
                    sm[i + ( i >> 4 )] = prev;
                    int j = i-2;
                    while( j >= b0 ) {
                        const int u = sm[j + ( j >> 4 )];
                        if( u <= v ) break;
                        sm[j+1 + ( ( j+1 ) >> 4 )] = u;
                        j--;
                    }
                    sm[j+1 + ( ( j+1 ) >> 4 )] = v;
                }
            } else {
                for( int i = cb+1; i < ce; i++ ) {
                    const int v = perm[i];
                    int j = i-1;
                    while( j >= cb ) {
                        const int u = perm[j];
                        if( u <= v ) break;
                        perm[j+1] = u;
                        j--;
                    }
                    perm[j+1] = v;
                }
            }
        }
        __syncwarp();
        if( staged )
            for( int q = lane; q < end - base; q += 32 ) perm[base + q] = sm[q + ( q >> 4 )];
        __syncwarp();
      }
    }
}

__global__ void __launch_bounds__( 256 ) k_max_count( const int *__restrict__ first, int ncells, int *__restrict__ out )
{
    int m = 0;
    for( int c = blockIdx.x*blockDim.x + threadIdx.x; c < ncells; c += gridDim.x*blockDim.x ) m = max( m, first[c+1] - first[c] );
#pragma unroll
    for( int d=16; d>0; d>>=1 ) m = max( m, __shfl_xor_sync( 0xffffffffu, m, d ) );
    if( ( threadIdx.x & 31 ) == 0 && m > 0 ) atomicMax( out, m );
}

// max |q*w| over a range of particles, folded into a device scalar holding the bits of a positive double
// (positive doubles order like their bit patterns)
__global__ void __launch_bounds__( 256 ) k_qwmax( const double *__restrict__ w, const short *__restrict__ q, size_t first, size_t n,
        unsigned long long *__restrict__ out )
{
    double m = 0.;
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x )
        m = fmax( m, fabs( ( double )q[first+i]*w[first+i] ) );
#pragma unroll
    for( int d=16; d>0; d>>=1 ) m = fmax( m, __shfl_xor_sync( 0xffffffffu, m, d ) );
    if( ( threadIdx.x & 31 ) == 0 && m > 0. ) atomicMax( out, ( unsigned long long )__double_as_longlong( m ) );
}

int update_qwmax( sb200_patch *p, int ispec, size_t first, size_t n )
{
    SpeciesDev &s = p->sp[ispec];
    if( n == 0 ) return 0;
    const unsigned blocks = ( unsigned )( ( n + 255 )/256 < 148*8 ? ( n + 255 )/256 : 148*8 );
    k_qwmax<<<blocks, 256, 0, p->stream>>>( s.col[6], s.q, first, n, s.d_qwmax );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

struct Cols { double *c[7]; short *q; int *key; };

// exchange the species' column set with the patch's spare set
void swap_with_spare( sb200_patch *p, SpeciesDev &s )
{
    for( int c=0; c<7; c++ ) { double *t = s.col[c]; s.col[c] = p->spare.col[c]; p->spare.col[c] = t; }
    { short *t = s.q; s.q = p->spare.q; p->spare.q = t; }
    { int *t = s.key; s.key = p->spare.key; p->spare.key = t; }
    const size_t tc = s.cap; s.cap = p->spare.cap; p->spare.cap = tc;
}

// out[c][j] = in[c][perm[j]] for the 9 columns.  Four consecutive output slots per thread: 36 independent
// loads in flight per thread and 32-B stores; perm is the identity plus small shifts except around movers, so
// the reads are almost as contiguous as the writes.
__global__ void __launch_bounds__( 256 ) k_gather( Cols in, Cols out, const int *__restrict__ perm, size_t n )
{
    const size_t nq = ( n + 3 )/4;
    for( size_t t = blockIdx.x*( size_t )blockDim.x + threadIdx.x; t < nq; t += ( size_t )gridDim.x*blockDim.x ) {
        const size_t j = 4*t;
        if( j + 3 < n ) {
            const int4 s = *reinterpret_cast<const int4 *>( perm + j );
#pragma unroll
            for( int c=0; c<7; c++ ) {
                const double a0 = in.c[c][s.x], a1 = in.c[c][s.y], a2 = in.c[c][s.z], a3 = in.c[c][s.w];
                *reinterpret_cast<double4 *>( out.c[c] + j ) = make_double4( a0, a1, a2, a3 );
            }
            const short q0 = in.q[s.x], q1 = in.q[s.y], q2 = in.q[s.z], q3 = in.q[s.w];
            *reinterpret_cast<short4 *>( out.q + j ) = make_short4( q0, q1, q2, q3 );
            const int k0 = in.key[s.x], k1 = in.key[s.y], k2 = in.key[s.z], k3 = in.key[s.w];
            *reinterpret_cast<int4 *>( out.key + j ) = make_int4( k0, k1, k2, k3 );
        } else {
            for( size_t jj = j; jj < n; jj++ ) {
                const int s = perm[jj];
#pragma unroll
                for( int c=0; c<7; c++ ) out.c[c][jj] = in.c[c][s];
                out.q[jj] = in.q[s];
                out.key[jj] = in.key[s];
            }
        }
    }
}

static int ensure_species_perm( SpeciesDev &s )
{
    if( s.perm_cap >= s.cap && s.perm ) return 0;
    if( s.perm ) cudaFree( s.perm );
    s.perm = nullptr; s.perm_cap = 0;
    SB200_CUDA( cudaMalloc( &s.perm, ( s.cap > 0 ? s.cap : 1 )*sizeof( int ) ) );
    s.perm_cap = s.cap;
    return 0;
}

// Apply a pending sort permutation: out[c][j] = in[c][perm[j]] into the spare set, then swap the sets.
int materialize( sb200_patch *p, int ispec )
{
    SpeciesDev &s = p->sp[ispec];
    if( !s.perm_pending ) return 0;
    s.perm_pending = false;
    if( s.n == 0 ) return 0;
    if( ensure_spare( p, s.cap ) ) return 1;
    Cols in, out;
    for( int c=0; c<7; c++ ) { in.c[c] = s.col[c]; out.c[c] = p->spare.col[c]; }
    in.q = s.q; in.key = s.key; out.q = p->spare.q; out.key = p->spare.key;
    const size_t nq = ( s.n + 3 )/4;
    const unsigned blocks = ( unsigned )( ( nq + 255 )/256 < 148*32 ? ( nq + 255 )/256 : 148*32 );
    k_gather<<<blocks, 256, 0, p->stream>>>( in, out, s.perm, s.n );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    swap_with_spare( p, s );
    return 0;
}

int launch_sort( sb200_patch *p, int ispec )
{
    SpeciesDev &s = p->sp[ispec];
    if( materialize( p, ispec ) ) return 1;          // a sort on top of an unapplied sort: apply the first one
    const size_t n = s.n;
    const int ncells = ( int )p->ncells;
    // the histogram of the new keys was accumulated by the dynamics kernel and the arrival unpack when the
    // species went through a step; a freshly imported species gets it here
    const bool have_hist = s.sorted && s.count_valid;
    if( !have_hist ) SB200_CUDA( cudaMemsetAsync( s.count, 0, ( p->ncells+1 )*sizeof( int ), p->stream ) );
    SB200_CUDA( cudaMemsetAsync( p->cursor, 0, ( p->ncells+1 )*sizeof( int ), p->stream ) );
    SB200_CUDA( cudaMemsetAsync( p->iflags, 0, 8*sizeof( int ), p->stream ) );
    if( n > 0 ) {
        if( ensure_species_perm( s ) ) return 1;
        const unsigned blocks = ( unsigned )( ( n + 255 )/256 < 148*16 ? ( n + 255 )/256 : 148*16 );
        // keys written by the fused dynamics kernel / arriving_unpack are already final; a
        // freshly imported species (keys all 0, unsorted) gets them computed here
        const int recompute = s.sorted ? 0 : 1;
        if( !have_hist ) {
            k_keys_hist<<<blocks, 256, 0, p->stream>>>( p->gd, s.col[0], s.col[1], s.col[2], s.key, s.count, n, recompute, ncells, p->iflags );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
        }
    }
    // first = exclusive scan(count) over ncells+1 entries (last = total kept)
    SB200_CUDA( cudaMemcpyAsync( s.first, s.count, ( p->ncells+1 )*sizeof( int ), cudaMemcpyDeviceToDevice, p->stream ) );
    if( exclusive_scan_int( p, s.first, p->ncells+1 ) ) return 1;
    int kept = 0, flags[8], maxcount = 0;
    unsigned long long qwbits = 0;
    SB200_CUDA( cudaMemsetAsync( p->d_maxcount, 0, sizeof( int ), p->stream ) );
    k_max_count<<<148*8, 256, 0, p->stream>>>( s.first, ncells, p->d_maxcount );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    SB200_CUDA( cudaMemcpyAsync( &maxcount, p->d_maxcount, sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( &qwbits, s.d_qwmax, sizeof( qwbits ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( &kept, s.first + p->ncells, sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( flags, p->iflags, 8*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    if( n > 0 ) {
        const unsigned blocks = ( unsigned )( ( n + 255 )/256 < 148*8 ? ( n + 255 )/256 : 148*8 );
        k_scatter_idx<<<blocks, 256, 0, p->stream>>>( s.key, s.first, p->cursor, s.perm, n );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
        // persistent grid: as many CTAs as are resident at once (the loop strides over the cells; a partial second
        // wave would run at a fraction of the occupancy)
        static int cs_blocks = 0;
        if( !cs_blocks ) {
            int per_sm = 1, dev = 0, sms = 148;
            cudaGetDevice( &dev );
            cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
            cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, k_cell_sort, 256, 0 );
            cs_blocks = sms*( per_sm > 0 ? per_sm : 1 );
        }
        k_cell_sort<<<cs_blocks, 256, 0, p->stream>>>( s.first, s.perm, ncells );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
    }
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    SB200_CHECK( flags[0] == 0, "sb200_sort: particles outside the patch without a leaving tag (positions must lie in [min,max) of the patch)" );
    // the columns stay where they are: the next dynamics kernel reads them through perm (materialize() for anyone else)
    s.n = ( size_t )kept;
    s.n_sorted = ( size_t )kept;
    s.window_tagged = false;
    s.perm_pending = kept > 0;
    s.sorted = true;
    s.count_valid = false;
    s.maxcount = maxcount;
    { double v; memcpy( &v, &qwbits, sizeof( v ) ); s.qwmax = v; }
    return 0;
}

} // namespace sb200
