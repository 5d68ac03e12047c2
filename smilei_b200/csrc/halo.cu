// halo.cu — guard-cell slabs and migrating particles: the pack/unpack hooks the exchange
// layer (NCCL over NVLink, or an MPI adapter) calls, and the same-GPU periodic wrap.
//
// Slab semantics restate SyncVectorPatch::sumAllComponents (src/Patch/SyncVectorPatch.cpp:203-…,
// local branch :263-311) and exchangeAllComponentsAlong{X,Y,Z} (:1441-1527, :1577-…, :1708-…);
// particle semantics restate Patch::prepareParticles / exchParticles / cornersParticles
// (src/Patch/Patch.cpp:624-800) and PartBoundCond::apply (src/ParticleBC/PartBoundCond.h:38-76).
#include "common.cuh"

namespace sb200 {

struct SlabGeom { int dims[3]; int dim; int first, nplanes; long long sx, sy; };

__device__ __forceinline__ long long slab_index( const SlabGeom &s, long long t )
{
    // t enumerates (plane, a, b) with plane along s.dim and (a,b) the two other dims in field order
    int e[3];
    int da = s.dim==0 ? 1 : 0, db = s.dim==2 ? 1 : 2;
    const int nb = s.dims[db], na = s.dims[da];
    e[db] = ( int )( t % nb );
    long long r = t / nb;
    e[da] = ( int )( r % na );
    e[s.dim] = s.first + ( int )( r / na );
    return e[0]*s.sx + e[1]*s.sy + e[2];
}

__global__ void __launch_bounds__( 256 ) k_slab_pack( SlabGeom s, const double *__restrict__ f, double *__restrict__ buf, long long total )
{
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x )
        buf[t] = f[slab_index( s, t )];
}

__global__ void __launch_bounds__( 256 ) k_slab_unpack( SlabGeom s, double *__restrict__ f, const double *__restrict__ buf, long long total, int add )
{
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x ) {
        const long long i = slab_index( s, t );
        f[i] = add ? f[i] + buf[t] : buf[t];
    }
}

// L == R periodic wrap of a sum: planes [n, n+gsp) += planes [0, gsp), both keep the sum
__global__ void __launch_bounds__( 256 ) k_slab_sum_self( SlabGeom s, double *__restrict__ f, long long total, long long shift )
{
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x ) {
        const long long i = slab_index( s, t );
        const double v = f[i+shift] + f[i];
        f[i+shift] = v;
        f[i] = v;
    }
}

// L == R periodic wrap of an exchange: [0,o) <- [n,n+o) ; [n+gsp, n+gsp+o) <- [gsp, gsp+o)
__global__ void __launch_bounds__( 256 ) k_slab_exch_self( SlabGeom s, double *__restrict__ f, long long total, long long shift, long long gshift )
{
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x ) {
        const long long i = slab_index( s, t );
        f[i] = f[i+shift];
        f[i+shift+gshift] = f[i+gshift];
    }
}

static int slab_geom( sb200_patch *p, int field_id, int dim, int first, int nplanes, SlabGeom &s, long long &total )
{
    SB200_CHECK( p && field_ptr( p, field_id ) && dim >= 0 && dim < 3, "halo: bad field or dimension (or a species array that was not requested)" );
    const GridDev &g = p->gd;
    for( int i=0; i<3; i++ ) s.dims[i] = field_dual( field_id, i ) ? g.d[i] : g.p[i];
    SB200_CHECK( first >= 0 && nplanes >= 0 && first+nplanes <= s.dims[dim], "halo: plane range outside the field" );
    s.dim = dim; s.first = first; s.nplanes = nplanes; s.sx = g.sx; s.sy = g.sy;
    total = ( long long )nplanes;
    for( int i=0; i<3; i++ ) if( i != dim ) total *= s.dims[i];
    return 0;
}

static unsigned nblocks( long long total ) { long long b = ( total + 255 )/256; return ( unsigned )( b < 148*8 ? ( b > 0 ? b : 1 ) : 148*8 ); }

// ---------------------------------------------------------------- particles
__device__ __forceinline__ int tag_or_key( const GridDev &g, double x, double y, double z )
{
    // PartBoundCond::apply order x,y,z; internal_inf: pos < min -> -2-2d ; internal_sup: pos >= max -> -3-2d
    if( x <  g.xmin[0] ) return -2;
    if( x >= g.xmax[0] ) return -3;
    if( y <  g.xmin[1] ) return -4;
    if( y >= g.xmax[1] ) return -5;
    if( z <  g.xmin[2] ) return -6;
    if( z >= g.xmax[2] ) return -7;
    int key = ( int )( round( x*g.dxi[0] ) - g.min_loc_round[0] );
    key *= g.ncell[1];
    key += ( int )( round( y*g.dxi[1] ) - g.min_loc_round[1] );
    key *= g.ncell[2];
    key += ( int )( round( z*g.dxi[2] ) - g.min_loc_round[2] );
    return key;
}

constexpr int CP_T = 256, CP_E = 8, CP_B = CP_T*CP_E;

// pass 1: matches per block of CP_B consecutive particles
__global__ void __launch_bounds__( CP_T ) k_leave_count( const int *__restrict__ key, size_t n, int tag, int *__restrict__ blockcount )
{
    __shared__ int tot;
    if( threadIdx.x == 0 ) tot = 0;
    __syncthreads();
    const size_t base = ( size_t )blockIdx.x*CP_B + ( size_t )threadIdx.x*CP_E;
    int c = 0;
#pragma unroll
    for( int e=0; e<CP_E; e++ ) if( base+e < n && key[base+e] == tag ) c++;
    if( c ) atomicAdd( &tot, c );
    __syncthreads();
    if( threadIdx.x == 0 ) blockcount[blockIdx.x] = tot;
}

struct ConstCols { const double *c[7]; const short *q; };

// pass 2: stable write of the records, index order
__global__ void __launch_bounds__( CP_T ) k_leave_write( ConstCols in, const int *__restrict__ key, size_t n, int tag, int dim, double wrap,
        double lo, double hi, const int *__restrict__ blockoff, double *__restrict__ buf, size_t max_records )
{
    __shared__ int warp_tot[CP_T/32];
    const size_t base = ( size_t )blockIdx.x*CP_B + ( size_t )threadIdx.x*CP_E;
    int c = 0;
    bool m[CP_E];
#pragma unroll
    for( int e=0; e<CP_E; e++ ) { m[e] = base+e < n && key[base+e] == tag; c += m[e]; }
    int inc = c;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for( int d=1; d<32; d<<=1 ) { int u = __shfl_up_sync( 0xffffffffu, inc, d ); if( lane >= d ) inc += u; }
    if( lane == 31 ) warp_tot[w] = inc;
    __syncthreads();
    int off = blockoff[blockIdx.x] + inc - c;
    for( int i=0; i<w; i++ ) off += warp_tot[i];
#pragma unroll
    for( int e=0; e<CP_E; e++ ) {
        if( !m[e] ) continue;
        if( ( size_t )off < max_records ) {
            const size_t i = base+e;
            double *r = buf + ( size_t )off*SB200_PARTICLE_RECORD_DOUBLES;
#pragma unroll
            for( int cc=0; cc<7; cc++ ) r[cc] = in.c[cc][i];
            // Patch::prepareParticles (Patch.cpp:633-650): wrap only positions that left the global box
            if( wrap > 0. ) { if( r[dim] < lo ) r[dim] += wrap; }
            else if( wrap < 0. ) { if( r[dim] >= hi ) r[dim] += wrap; }
            r[7] = ( double )in.q[i];
        }
        off++;
    }
}

struct MutCols { double *c[7]; short *q; int *key; };

__global__ void __launch_bounds__( 256 ) k_arrive( GridDev g, MutCols out, size_t n0, const double *__restrict__ buf, size_t n,
        int *__restrict__ leave_counts, int *__restrict__ leave_idx, int leave_cap, int *__restrict__ count )
{
    for( size_t t = blockIdx.x*( size_t )blockDim.x + threadIdx.x; t < n; t += ( size_t )gridDim.x*blockDim.x ) {
        const double *r = buf + t*SB200_PARTICLE_RECORD_DOUBLES;
        const size_t i = n0 + t;
#pragma unroll
        for( int cc=0; cc<7; cc++ ) out.c[cc][i] = r[cc];
        out.q[i] = ( short )r[7];
        const int k = tag_or_key( g, r[0], r[1], r[2] );
        out.key[i] = k;
        if( k < 0 ) {
            const int c = atomicAdd( &leave_counts[-k-2], 1 );
            if( c < leave_cap ) leave_idx[( size_t )( -k-2 )*leave_cap + c] = ( int )i;
        } else if( count ) {
            atomicAdd( &count[k], 1 );
        }
    }
}

// ---------------------------------------------------------------- leavers of one box side, in index order
// A particle tagged for side (dim, side) sat, before the push, in one of the two outermost node layers of that
// side (it moves less than a cell per step), or it is an arrival appended behind the sorted runs.  The layers'
// cells are enumerated in increasing cell index, then the appended tail in chunks: counting the tagged particles
// per entry, scanning, and writing the records at the scanned offsets is a stable compaction of the boundary
// layer only — a few hundred thousand keys instead of all of them, and no quadratic ranking of a leaver list.
struct LayerGeom { int dim, l0, l1, nc0, nc1, nc2; unsigned ncells_layer; unsigned ntail_chunks; size_t n_sorted, n; };

__device__ __forceinline__ int layer_cell( const LayerGeom &G, unsigned e )
{
    if( G.dim == 0 ) { const unsigned A = ( unsigned )G.nc1*G.nc2; const unsigned l = e / A, r = e - l*A; return ( int )( ( l ? G.l1 : G.l0 )*A + r ); }
    if( G.dim == 1 ) { const unsigned iz = e % G.nc2, t = e / G.nc2, l = t & 1u, ix = t >> 1; return ( int )( ( ix*G.nc1 + ( l ? G.l1 : G.l0 ) )*G.nc2 + iz ); }
    const unsigned l = e & 1u, t = e >> 1;
    return ( int )( t*G.nc2 + ( l ? G.l1 : G.l0 ) );
}

__global__ void __launch_bounds__( 256 ) k_layer_count( LayerGeom G, const int *__restrict__ first, const int *__restrict__ key, int tag, int *__restrict__ cnt )
{
    const unsigned total = G.ncells_layer + G.ntail_chunks;
    for( unsigned e = blockIdx.x*blockDim.x + threadIdx.x; e < total; e += gridDim.x*blockDim.x ) {
        size_t b, en;
        if( e < G.ncells_layer ) { const int c = layer_cell( G, e ); b = ( size_t )first[c]; en = ( size_t )first[c+1]; }
        else { b = G.n_sorted + ( size_t )( e - G.ncells_layer )*256; en = b + 256 < G.n ? b + 256 : G.n; }
        int c = 0;
        for( size_t i=b; i<en; i++ ) c += key[i] == tag;
        cnt[e] = c;
    }
}

__global__ void __launch_bounds__( 256 ) k_layer_write( LayerGeom G, ConstCols in, const int *__restrict__ first, const int *__restrict__ key, int tag,
        int dim, double wrap, double lo, double hi, const int *__restrict__ off, double *__restrict__ buf, size_t max_records )
{
    const unsigned total = G.ncells_layer + G.ntail_chunks;
    for( unsigned e = blockIdx.x*blockDim.x + threadIdx.x; e < total; e += gridDim.x*blockDim.x ) {
        size_t b, en;
        if( e < G.ncells_layer ) { const int c = layer_cell( G, e ); b = ( size_t )first[c]; en = ( size_t )first[c+1]; }
        else { b = G.n_sorted + ( size_t )( e - G.ncells_layer )*256; en = b + 256 < G.n ? b + 256 : G.n; }
        size_t o = ( size_t )off[e];
        for( size_t i=b; i<en; i++ ) {
            if( key[i] != tag ) continue;
            if( o < max_records ) {
                double *r = buf + o*SB200_PARTICLE_RECORD_DOUBLES;
#pragma unroll
                for( int cc=0; cc<7; cc++ ) r[cc] = in.c[cc][i];
                if( wrap > 0. ) { if( r[dim] < lo ) r[dim] += wrap; }          // Patch::prepareParticles, Patch.cpp:633-650
                else if( wrap < 0. ) { if( r[dim] >= hi ) r[dim] += wrap; }
                r[7] = ( double )in.q[i];
            }
            o++;
        }
    }
}

} // namespace sb200

using namespace sb200;

// stable pack of the leavers of one side from the boundary layer (see k_layer_count); total_out may be null
static int pack_boundary_layer( sb200_patch *p, SpeciesDev &s, int tag, int dim, int side, double wrap, double hi,
                                double *dev_buf, size_t max_records, int *d_total_out )
{
    const GridDev &g = p->gd;
    LayerGeom G;
    G.dim = dim; G.nc0 = g.ncell[0]; G.nc1 = g.ncell[1]; G.nc2 = g.ncell[2];
    const int nc = g.ncell[dim];
    G.l0 = side == 0 ? 0 : nc-2; G.l1 = side == 0 ? 1 : nc-1;
    G.ncells_layer = 2u*( unsigned )( ( size_t )G.nc0*G.nc1*G.nc2 / nc );
    G.n_sorted = s.n_sorted; G.n = s.n;
    G.ntail_chunks = ( unsigned )( ( s.n - s.n_sorted + 255 )/256 );
    const size_t ne = ( size_t )G.ncells_layer + G.ntail_chunks;
    if( ensure_perm( p, ne + 1 ) ) return 1;
    int *cnt = p->perm;                       // scratch (the sort orders live in the species' own perm arrays)
    const unsigned blocks = ( unsigned )( ( ne + 255 )/256 < 148*8 ? ( ne + 255 )/256 : 148*8 );
    k_layer_count<<<blocks, 256, 0, p->stream>>>( G, s.first, s.key, tag, cnt );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    SB200_CUDA( cudaMemsetAsync( cnt + ne, 0, sizeof( int ), p->stream ) );
    if( exclusive_scan_int( p, cnt, ne + 1 ) ) return 1;
    if( d_total_out ) SB200_CUDA( cudaMemcpyAsync( d_total_out, cnt + ne, sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    ConstCols in;
    for( int c=0; c<7; c++ ) in.c[c] = s.col[c];
    in.q = s.q;
    k_layer_write<<<blocks, 256, 0, p->stream>>>( G, in, s.first, s.key, tag, dim, wrap, 0., hi, cnt, dev_buf, max_records );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

extern "C" {

int sb200_halo_plane_elems( sb200_patch *p, int field_id, int dim, size_t *elems )
{
    SlabGeom s; long long total;
    if( slab_geom( p, field_id, dim, 0, 1, s, total ) ) return 1;
    SB200_CHECK( elems, "sb200_halo_plane_elems: null pointer" );
    *elems = ( size_t )total;
    return 0;
}

int sb200_halo_pack( sb200_patch *p, int field_id, int dim, int first_plane, int nplanes, double *dev_buf )
{
    SlabGeom s; long long total;
    if( slab_geom( p, field_id, dim, first_plane, nplanes, s, total ) ) return 1;
    SB200_CHECK( dev_buf, "sb200_halo_pack: null buffer" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( total == 0 ) return 0;
    k_slab_pack<<<nblocks( total ), 256, 0, p->stream>>>( s, field_ptr( p, field_id ), dev_buf, total );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int sb200_halo_unpack( sb200_patch *p, int field_id, int dim, int first_plane, int nplanes, const double *dev_buf, int mode )
{
    SlabGeom s; long long total;
    if( slab_geom( p, field_id, dim, first_plane, nplanes, s, total ) ) return 1;
    SB200_CHECK( dev_buf, "sb200_halo_unpack: null buffer" );
    SB200_CHECK( mode == SB200_UNPACK_COPY || mode == SB200_UNPACK_ADD, "sb200_halo_unpack: bad mode" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( total == 0 ) return 0;
    k_slab_unpack<<<nblocks( total ), 256, 0, p->stream>>>( s, field_ptr( p, field_id ), dev_buf, total, mode == SB200_UNPACK_ADD );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int sb200_halo_sum_self( sb200_patch *p, int field_id, int dim )
{
    SB200_CHECK( p && dim >= 0 && dim < 3, "sb200_halo_sum_self: bad arguments" );
    const GridDev &g = p->gd;
    const int gsp = 1 + 2*g.o[dim] + field_dual( field_id, dim );      // SyncVectorPatch.cpp:284
    SlabGeom s; long long total;
    if( slab_geom( p, field_id, dim, 0, gsp, s, total ) ) return 1;
    SB200_CHECK( g.n[dim] >= gsp, "sb200_halo_sum_self: patch too small for a self wrap" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    const long long stride = dim==0 ? g.sx : dim==1 ? g.sy : 1;
    k_slab_sum_self<<<nblocks( total ), 256, 0, p->stream>>>( s, field_ptr( p, field_id ), total, ( long long )g.n[dim]*stride );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int sb200_halo_exchange_self( sb200_patch *p, int field_id, int dim )
{
    SB200_CHECK( p && dim >= 0 && dim < 3, "sb200_halo_exchange_self: bad arguments" );
    const GridDev &g = p->gd;
    const int o = g.o[dim];
    const int gsp = o + 1 + field_dual( field_id, dim );               // SyncVectorPatch.cpp:1501
    SlabGeom s; long long total;
    if( slab_geom( p, field_id, dim, 0, o, s, total ) ) return 1;
    SB200_CUDA( cudaSetDevice( p->device ) );
    const long long stride = dim==0 ? g.sx : dim==1 ? g.sy : 1;
    k_slab_exch_self<<<nblocks( total ), 256, 0, p->stream>>>( s, field_ptr( p, field_id ), total, ( long long )g.n[dim]*stride, ( long long )gsp*stride );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    return 0;
}

int sb200_leaving_count( sb200_patch *p, int ispec, int counts[6] )
{
    SB200_CHECK( p && counts && ispec >= 0 && ispec < p->nspec, "sb200_leaving_count: bad arguments" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SB200_CUDA( cudaMemcpyAsync( counts, p->leave_counts + 8*ispec, 6*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

// stable compaction of the particles tagged `tag` over ALL keys (no cell runs needed): count per block, scan, write.
// known < 0: the total is read back and returned in *total; else it is taken as known (no host round trip).
static int pack_all_keys( sb200_patch *p, SpeciesDev &s, const ConstCols &in, int tag, int dim, double wrap, double hi,
                          double *dev_buf, size_t max_records, long long known, size_t *total_out )
{
    const size_t nb = ( s.n + CP_B - 1 )/CP_B;
    if( ensure_perm( p, nb + 1 > s.cap ? nb + 1 : s.cap ) ) return 1;
    int *bc = p->perm;             // perm is free between sorts
    k_leave_count<<<( unsigned )nb, CP_T, 0, p->stream>>>( s.key, s.n, tag, bc );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    SB200_CUDA( cudaMemsetAsync( bc+nb, 0, sizeof( int ), p->stream ) );
    if( exclusive_scan_int( p, bc, nb+1 ) ) return 1;
    int total = ( int )known;
    if( known < 0 ) {
        SB200_CUDA( cudaMemcpyAsync( &total, bc+nb, sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
        SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    }
    SB200_CHECK( ( size_t )total <= max_records, "sb200_leaving_pack: buffer too small for the leaving particles" );
    if( total > 0 ) {
        SB200_CHECK( dev_buf, "sb200_leaving_pack: null buffer" );
        k_leave_write<<<( unsigned )nb, CP_T, 0, p->stream>>>( in, s.key, s.n, tag, dim, wrap, 0., hi, bc, dev_buf, max_records );
        sb200::g_launches++;
        SB200_CUDA( cudaGetLastError() );
    }
    if( total_out ) *total_out = ( size_t )total;
    return 0;
}

int sb200_leaving_pack( sb200_patch *p, int ispec, int dim, int side, double wrap, double *dev_buf, size_t max_records, size_t *n_packed )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec && dim >= 0 && dim < 3 && ( side==0 || side==1 ), "sb200_leaving_pack: bad arguments" );
    SB200_CHECK( n_packed, "sb200_leaving_pack: null count pointer" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SpeciesDev &s = p->sp[ispec];
    *n_packed = 0;
    if( s.n == 0 ) return 0;
    if( materialize( p, ispec ) ) return 1;
    const int tag = -2 - 2*dim - side;
    const GridDev &g = p->gd;
    const double hi = g.cell[dim]*( double )( g.n[dim]*g.npatch[dim] );   // Patch.cpp:626: cell_length*global_size
    ConstCols in;
    for( int c=0; c<7; c++ ) in.c[c] = s.col[c];
    in.q = s.q;
    if( s.n_sorted > 0 && !s.window_tagged ) {
        // the leavers sit in the boundary layer of that side (or in the appended tail): compact those cells only
        int cnt = 0;
        SB200_CUDA( cudaMemcpyAsync( &cnt, p->leave_counts + 8*ispec + ( -tag-2 ), sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
        SB200_CUDA( cudaStreamSynchronize( p->stream ) );
        SB200_CHECK( ( size_t )cnt <= max_records, "sb200_leaving_pack: buffer too small for the leaving particles" );
        if( cnt > 0 ) {
            SB200_CHECK( dev_buf, "sb200_leaving_pack: null buffer" );
            if( pack_boundary_layer( p, s, tag, dim, side, wrap, hi, dev_buf, max_records, nullptr ) ) return 1;
        }
        *n_packed = ( size_t )cnt;
        return 0;
    }
    // general path (no valid cell runs, or leavers tagged by a window shift): stable compaction over all keys
    return pack_all_keys( p, s, in, tag, dim, wrap, hi, dev_buf, max_records, -1, n_packed );
}

int sb200_leaving_pack_known( sb200_patch *p, int ispec, int dim, int side, double wrap, double *dev_buf, size_t max_records, size_t n_known )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec && dim >= 0 && dim < 3 && ( side==0 || side==1 ), "sb200_leaving_pack_known: bad arguments" );
    SpeciesDev &s = p->sp[ispec];
    SB200_CHECK( n_known <= max_records, "sb200_leaving_pack_known: buffer too small for the leaving particles" );
    if( n_known == 0 ) return 0;
    SB200_CHECK( dev_buf, "sb200_leaving_pack_known: null buffer" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    const int tag = -2 - 2*dim - side;
    const GridDev &g = p->gd;
    const double hi = g.cell[dim]*( double )( g.n[dim]*g.npatch[dim] );
    if( s.n_sorted == 0 || s.window_tagged ) {
        // a species that was empty at its last sort (vacuum / slab outside this rank) or whose leavers a window shift
        // tagged has no cell runs to walk: its leavers — corner particles just received and re-tagged, typically —
        // are compacted over all keys, with the count the caller already knows
        ConstCols in;
        for( int c=0; c<7; c++ ) in.c[c] = s.col[c];
        in.q = s.q;
        return pack_all_keys( p, s, in, tag, dim, wrap, hi, dev_buf, max_records, ( long long )n_known, nullptr );
    }
    return pack_boundary_layer( p, s, tag, dim, side, wrap, hi, dev_buf, max_records, nullptr );
}

int sb200_arriving_unpack( sb200_patch *p, int ispec, const double *dev_buf, size_t n )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_arriving_unpack: bad arguments" );
    SpeciesDev &s = p->sp[ispec];
    if( n == 0 ) return 0;
    SB200_CHECK( dev_buf, "sb200_arriving_unpack: null buffer" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    if( grow_species( p, ispec, s.n + n ) ) return 1;            // Particles::resize of the reference: room on demand
    MutCols out;
    for( int c=0; c<7; c++ ) out.c[c] = s.col[c];
    out.q = s.q; out.key = s.key;
    const unsigned blocks = ( unsigned )( ( n + 255 )/256 < 148*8 ? ( n + 255 )/256 : 148*8 );
    k_arrive<<<blocks, 256, 0, p->stream>>>( p->gd, out, s.n, dev_buf, n, p->leave_counts + 8*ispec, s.leave_idx, ( int )s.leave_cap, s.count_valid ? s.count : nullptr );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    const size_t n0 = s.n;
    s.n += n;
    return update_qwmax( p, ispec, n0, n );
}

} // extern "C"
