// patch.cu — patch handle, device memory, host<->device import/export of the
// reference-layout field and particle arrays.
#include "common.cuh"
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace sb200 {

static thread_local std::string g_err;
unsigned long long g_launches = 0;
void set_error( const std::string &s ) { g_err = s; }

static void fill_grid( const sb200_grid &g, GridDev &d )
{
    d.order = g.interp_order;
    for( int i=0; i<3; i++ ) {
        d.n[i] = g.n[i];
        d.o[i] = g.oversize[i];
        d.p[i] = g.n[i] + 2*g.oversize[i] + 1;            // ElectroMagn.cpp:47
        d.d[i] = g.n[i] + 2*g.oversize[i] + 2;            // ElectroMagn.cpp:48
        d.ncell[i] = g.n[i] + 1;                          // SpeciesV.cpp:72-77
        d.begin[i] = g.pcoord[i]*g.n[i] - g.oversize[i];  // Patch.cpp:148-150
        d.cell[i] = g.cell_length[i];
        d.dxi[i] = 1.0/g.cell_length[i];                  // Interpolator3D2Order.cpp:18-20
        d.d_ov_dt[i] = g.cell_length[i]/g.dt;             // Projector3D2Order.cpp:21-25
        d.dt_ov_d[i] = g.dt/g.cell_length[i];             // Solver3D.h:20-22
        d.xmin[i] = ( g.pcoord[i]   )*( g.n[i]*g.cell_length[i] );   // Patch.cpp:146
        d.xmax[i] = ( g.pcoord[i]+1 )*( g.n[i]*g.cell_length[i] );   // Patch.cpp:147
        d.min_loc_round[i] = std::round( d.xmin[i]*d.dxi[i] );       // SpeciesV.cpp:800-802
        d.pcoord[i] = g.pcoord[i];
        d.npatch[i] = g.npatch[i];
    }
    d.dt = g.dt;
    d.dts2 = g.dt/2.;                                     // Pusher.cpp:23
    d.cell_volume = 1.0;
    for( int i=0; i<3; i++ ) d.cell_volume *= g.cell_length[i];      // Params.cpp:1172-1186
    d.inv_cell_volume = 1./d.cell_volume;                 // Projector.cpp:7
    d.ax = d.d[0];
    d.ay = d.d[1];
    d.az = ( d.d[2] + 7 )/8*8;          // rows start on a 64-B boundary (two 32-B sectors); 16-B vector accesses and the TMA strides need az even
    d.sy = d.az;
    d.sx = ( long long )d.ay*d.az;
}

int ensure_stage( sb200_patch *p, size_t elems )
{
    if( elems <= p->stage_cap ) return 0;
    if( p->stage ) SB200_CUDA( cudaFree( p->stage ) );
    p->stage = nullptr; p->stage_cap = 0;
    SB200_CUDA( cudaMalloc( &p->stage, elems*sizeof( double ) ) );
    p->stage_cap = elems;
    return 0;
}

static int alloc_particle_cols( double **col, short **q, int **key, size_t cap )
{
    for( int c=0; c<7; c++ ) SB200_CUDA( cudaMalloc( &col[c], cap*sizeof( double ) ) );
    SB200_CUDA( cudaMalloc( q, cap*sizeof( short ) ) );
    SB200_CUDA( cudaMalloc( key, cap*sizeof( int ) ) );
    return 0;
}
static void free_particle_cols( double **col, short **q, int **key )
{
    for( int c=0; c<7; c++ ) { if( col[c] ) cudaFree( col[c] ); col[c] = nullptr; }
    if( *q ) cudaFree( *q ); *q = nullptr;
    if( *key ) cudaFree( *key ); *key = nullptr;
}

int ensure_spare( sb200_patch *p, size_t cap )
{
    if( cap <= p->spare.cap ) return 0;
    free_particle_cols( p->spare.col, &p->spare.q, &p->spare.key );
    p->spare.cap = 0;
    if( alloc_particle_cols( p->spare.col, &p->spare.q, &p->spare.key, cap ) ) return 1;
    p->spare.cap = cap;
    return 0;
}

// The reference resizes Particles whenever arrivals or created particles need room (Particles::resize,
// SpeciesV.cpp:660-665).  Here: a larger column set, the live particles copied on the device, the old set freed.
int grow_species( sb200_patch *p, int ispec, size_t need )
{
    SpeciesDev &s = p->sp[ispec];
    if( need <= s.cap ) return 0;
    SB200_CHECK( need < ( size_t )2000000000u, "species capacity must fit int indices" );
    if( materialize( p, ispec ) ) return 1;                // no pending sort order refers to the old arrays afterwards
    size_t cap = s.cap + s.cap/2 + 4096;
    if( cap < need ) cap = need;
    if( cap >= ( size_t )2000000000u ) cap = ( size_t )2000000000u - 1;
    double *col[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    short *q = nullptr;
    int *key = nullptr;
    if( alloc_particle_cols( col, &q, &key, cap ) ) return 1;
    if( s.n > 0 ) {
        for( int c=0; c<7; c++ ) SB200_CUDA( cudaMemcpyAsync( col[c], s.col[c], s.n*sizeof( double ), cudaMemcpyDeviceToDevice, p->stream ) );
        SB200_CUDA( cudaMemcpyAsync( q, s.q, s.n*sizeof( short ), cudaMemcpyDeviceToDevice, p->stream ) );
        SB200_CUDA( cudaMemcpyAsync( key, s.key, s.n*sizeof( int ), cudaMemcpyDeviceToDevice, p->stream ) );
    }
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    free_particle_cols( s.col, &s.q, &s.key );
    for( int c=0; c<7; c++ ) s.col[c] = col[c];
    s.q = q; s.key = key; s.cap = cap;
    return 0;
}

int ensure_perm( sb200_patch *p, size_t cap )
{
    if( cap <= p->perm_cap ) return 0;
    if( p->perm ) cudaFree( p->perm );
    p->perm = nullptr; p->perm_cap = 0;
    SB200_CUDA( cudaMalloc( &p->perm, cap*sizeof( int ) ) );
    p->perm_cap = cap;
    return 0;
}

// compact (reference) <-> padded (device) field conversion through the staging buffer
__global__ void k_field_pad( const double *__restrict__ compact, double *__restrict__ padded,
                             int nx, int ny, int nz, long long sx, long long sy, int to_padded )
{
    long long total = ( long long )nx*ny*nz;
    for( long long t = blockIdx.x*( long long )blockDim.x + threadIdx.x; t < total; t += ( long long )gridDim.x*blockDim.x ) {
        int k = ( int )( t % nz );
        long long r = t / nz;
        int j = ( int )( r % ny );
        int i = ( int )( r / ny );
        long long pi = i*sx + j*sy + k;
        if( to_padded ) padded[pi] = compact[t];
        else ( ( double * )compact )[t] = padded[pi];
    }
}

static void field_dims( const GridDev &g, int id, int dims[3] )
{
    for( int i=0; i<3; i++ ) dims[i] = field_dual( id, i ) ? g.d[i] : g.p[i];
}

} // namespace sb200

using namespace sb200;

extern "C" {

const char *sb200_last_error( void ) { return g_err.c_str(); }
int sb200_abi_version( void ) { return SB200_ABI_VERSION; }

int sb200_device_count( int *count )
{
    SB200_CHECK( count, "sb200_device_count: null pointer" );
    SB200_CUDA( cudaGetDeviceCount( count ) );
    return 0;
}

int sb200_patch_create( sb200_patch **out, const sb200_grid *grid, int n_species, int device )
{
    SB200_CHECK( out && grid, "sb200_patch_create: null pointer" );
    SB200_CHECK( grid->interp_order==2 || grid->interp_order==4, "sb200_patch_create: interpolation_order must be 2 or 4" );
    for( int i=0; i<3; i++ ) {
        SB200_CHECK( grid->n[i] >= 2*grid->oversize[i]+2, "sb200_patch_create: patch_size must be >= 2*oversize+2 cells" );
        SB200_CHECK( grid->oversize[i] >= grid->interp_order, "sb200_patch_create: oversize must be >= interpolation_order" );
        SB200_CHECK( grid->cell_length[i] > 0., "sb200_patch_create: cell_length must be > 0" );
        SB200_CHECK( grid->npatch[i] >= 1 && grid->pcoord[i] >= 0 && grid->pcoord[i] < grid->npatch[i], "sb200_patch_create: bad patch coordinates" );
    }
    SB200_CHECK( grid->dt > 0., "sb200_patch_create: timestep must be > 0" );
    SB200_CHECK( n_species >= 0 && n_species <= 64, "sb200_patch_create: bad species count" );
    int ndev = 0;
    SB200_CUDA( cudaGetDeviceCount( &ndev ) );
    SB200_CHECK( ndev > 0, "sb200_patch_create: no CUDA device (there is no CPU fallback)" );
    SB200_CHECK( device >= 0 && device < ndev, "sb200_patch_create: bad device index" );
    SB200_CUDA( cudaSetDevice( device ) );
    cudaDeviceProp prop;
    SB200_CUDA( cudaGetDeviceProperties( &prop, device ) );
    SB200_CHECK( prop.major >= 10, "sb200_patch_create: kernels are built for sm_100a only" );

    sb200_patch *p = new sb200_patch();
    p->grid = *grid;
    fill_grid( *grid, p->gd );
    p->device = device;
    p->nspec = n_species;
    p->sp = new SpeciesDev[n_species > 0 ? n_species : 1];
    const GridDev &g = p->gd;
    SB200_CHECK( ( double )g.ncell[0]*g.ncell[1]*g.ncell[2] < 2.0e9, "sb200_patch_create: too many cells for int keys" );
    p->ncells = ( size_t )g.ncell[0]*g.ncell[1]*g.ncell[2];
    p->falloc = ( size_t )g.ax*g.ay*g.az;
    for( int f=0; f<SB200_NFIELDS; f++ ) {
        SB200_CUDA( cudaMalloc( &p->f[f], p->falloc*sizeof( double ) ) );
        SB200_CUDA( cudaMemset( p->f[f], 0, p->falloc*sizeof( double ) ) );
    }
    SB200_CUDA( cudaMalloc( &p->cursor, ( p->ncells+1 )*sizeof( int ) ) );
    SB200_CUDA( cudaMalloc( &p->red, 4096*sizeof( double ) ) );
    SB200_CUDA( cudaMalloc( &p->leave_counts, 8*( n_species+1 )*sizeof( int ) ) );
    SB200_CUDA( cudaMalloc( &p->iflags, 8*sizeof( int ) ) );
    SB200_CUDA( cudaMalloc( &p->d_maxcount, sizeof( int ) ) );
    SB200_CUDA( cudaMemset( p->leave_counts, 0, 8*( n_species+1 )*sizeof( int ) ) );
    SB200_CUDA( cudaMemset( p->iflags, 0, 8*sizeof( int ) ) );
    *out = p;
    return 0;
}

int sb200_patch_destroy( sb200_patch *p )
{
    if( !p ) return 0;
    cudaSetDevice( p->device );
    cudaDeviceSynchronize();
    for( int f=0; f<SB200_NFIELDS; f++ ) if( p->f[f] ) cudaFree( p->f[f] );
    for( int s=0; s<p->nspec; s++ ) {
        free_particle_cols( p->sp[s].col, &p->sp[s].q, &p->sp[s].key );
        if( p->sp[s].first ) cudaFree( p->sp[s].first );
        if( p->sp[s].count ) cudaFree( p->sp[s].count );
        if( p->sp[s].d_qwmax ) cudaFree( p->sp[s].d_qwmax );
        if( p->sp[s].leave_idx ) cudaFree( p->sp[s].leave_idx );
        for( int k=0; k<4; k++ ) if( p->sp[s].fs[k] ) cudaFree( p->sp[s].fs[k] );
        if( p->sp[s].perm ) cudaFree( p->sp[s].perm );
        if( p->sp[s].d_lost ) cudaFree( p->sp[s].d_lost );
    }
    free_particle_cols( p->spare.col, &p->spare.q, &p->spare.key );
    void *misc[] = { p->cursor, p->perm, p->blocksums, p->stage, p->red, p->leave_counts, p->iflags, p->d_maxcount,
                     p->sc_E, p->sc_B, p->sc_invgf, p->sc_delta, p->sc_iold };
    for( void *m : misc ) if( m ) cudaFree( m );
    delete[] p->sp;
    delete p;
    return 0;
}

int sb200_patch_set_stream( sb200_patch *p, void *cuda_stream )
{
    SB200_CHECK( p, "sb200_patch_set_stream: null patch" );
    p->stream = ( cudaStream_t )cuda_stream;
    return 0;
}

int sb200_patch_synchronize( sb200_patch *p )
{
    SB200_CHECK( p, "sb200_patch_synchronize: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_species_config( sb200_patch *p, int ispec, double mass, int pusher, size_t capacity )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_species_config: bad species index" );
    SB200_CHECK( pusher >= 0 && pusher <= 2, "sb200_species_config: pusher must be boris(0), vay(1) or higueracary(2)" );
    SB200_CHECK( mass > 0., "sb200_species_config: only massive species are on this path (mass > 0)" );
    SB200_CHECK( capacity < ( size_t )2000000000u, "sb200_species_config: capacity must fit int indices" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SpeciesDev &s = p->sp[ispec];
    s.mass = mass;
    s.pusher = pusher;
    if( capacity > s.cap ) {
        if( s.n == 0 ) {
            free_particle_cols( s.col, &s.q, &s.key );
            if( alloc_particle_cols( s.col, &s.q, &s.key, capacity ) ) return 1;
            s.cap = capacity;
        } else if( grow_species( p, ispec, capacity ) ) return 1;
    }
    {
        const size_t want = capacity/64 > ( size_t )65536 ? capacity/64 : ( size_t )65536;
        if( want > s.leave_cap ) {
            if( s.leave_idx ) cudaFree( s.leave_idx );
            s.leave_idx = nullptr; s.leave_cap = 0;
            SB200_CUDA( cudaMalloc( &s.leave_idx, 6*want*sizeof( int ) ) );
            s.leave_cap = want;
        }
    }
    if( !s.d_qwmax ) {
        SB200_CUDA( cudaMalloc( &s.d_qwmax, sizeof( unsigned long long ) ) );
        SB200_CUDA( cudaMemset( s.d_qwmax, 0, sizeof( unsigned long long ) ) );
    }
    if( !s.d_lost ) {
        SB200_CUDA( cudaMalloc( &s.d_lost, sizeof( double ) ) );
        SB200_CUDA( cudaMemset( s.d_lost, 0, sizeof( double ) ) );
    }
    if( !s.count ) {
        SB200_CUDA( cudaMalloc( &s.count, ( p->ncells+1 )*sizeof( int ) ) );
        SB200_CUDA( cudaMemset( s.count, 0, ( p->ncells+1 )*sizeof( int ) ) );
    }
    if( !s.first ) {
        SB200_CUDA( cudaMalloc( &s.first, ( p->ncells+1 )*sizeof( int ) ) );
        SB200_CUDA( cudaMemset( s.first, 0, ( p->ncells+1 )*sizeof( int ) ) );
    }
    return 0;
}

int sb200_species_set_bc( sb200_patch *p, int ispec, const int bc[6] )
{
    SB200_CHECK( p && bc && ispec >= 0 && ispec < p->nspec, "sb200_species_set_bc: bad arguments" );
    for( int i=0; i<6; i++ )
        SB200_CHECK( bc[i] == SB200_PBC_PERIODIC || bc[i] == SB200_PBC_REMOVE,
                     "sb200_species_set_bc: only `periodic` and `remove` particle boundary conditions are on this path" );
    for( int i=0; i<6; i++ ) p->sp[ispec].bc[i] = bc[i];
    return 0;
}

int sb200_species_lost_energy( sb200_patch *p, int ispec, double *lost, int reset )
{
    SB200_CHECK( p && lost && ispec >= 0 && ispec < p->nspec, "sb200_species_lost_energy: bad arguments" );
    SpeciesDev &s = p->sp[ispec];
    SB200_CHECK( s.d_lost, "sb200_species_lost_energy: species not configured" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    double v = 0.;
    SB200_CUDA( cudaMemcpyAsync( &v, s.d_lost, sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    if( reset ) SB200_CUDA( cudaMemsetAsync( s.d_lost, 0, sizeof( double ), p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    *lost = s.mass*v;
    return 0;
}

int sb200_species_set( sb200_patch *p, int ispec,
                       const double *x, const double *y, const double *z,
                       const double *px, const double *py, const double *pz,
                       const double *w, const short *q, size_t n )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_species_set: bad species index" );
    SpeciesDev &s = p->sp[ispec];
    SB200_CHECK( n <= s.cap, "sb200_species_set: more particles than the configured capacity" );
    SB200_CHECK( n == 0 || ( x && y && z && px && py && pz && w && q ), "sb200_species_set: null column" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    const double *src[7] = { x, y, z, px, py, pz, w };
    for( int c=0; c<7; c++ ) SB200_CUDA( cudaMemcpyAsync( s.col[c], src[c], n*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( s.q, q, n*sizeof( short ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemsetAsync( s.key, 0, n*sizeof( int ), p->stream ) );
    SB200_CUDA( cudaMemsetAsync( s.d_qwmax, 0, sizeof( unsigned long long ), p->stream ) );
    s.n = n;
    s.n_sorted = 0;
    s.count_valid = false;
    s.perm_pending = false;
    if( update_qwmax( p, ispec, 0, n ) ) return 1;
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    s.sorted = false;
    return 0;
}

int sb200_species_append( sb200_patch *p, int ispec,
                          const double *x, const double *y, const double *z,
                          const double *px, const double *py, const double *pz,
                          const double *w, const short *q, size_t n )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_species_append: bad species index" );
    if( n == 0 ) return 0;
    SB200_CHECK( x && y && z && px && py && pz && w && q, "sb200_species_append: null column" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    SpeciesDev &s = p->sp[ispec];
    if( grow_species( p, ispec, s.n + n ) ) return 1;            // Particles::resize of the reference: room on demand
    const double *src[7] = { x, y, z, px, py, pz, w };
    for( int c=0; c<7; c++ ) SB200_CUDA( cudaMemcpyAsync( s.col[c] + s.n, src[c], n*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( s.q + s.n, q, n*sizeof( short ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemsetAsync( s.key + s.n, 0, n*sizeof( int ), p->stream ) );
    const size_t n0 = s.n;
    s.n += n;
    s.sorted = false;
    s.count_valid = false;
    if( update_qwmax( p, ispec, n0, n ) ) return 1;
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );       // the host arrays may be reused by the caller
    return 0;
}

// ParticleCreator for the common laser-wake case, on the device: position_initialization "regular"
// (ParticleCreator.cpp:627-667: x = x_cell + dx*0.975*(0.5 + i%c)/c per dimension), momentum_initialization "cold",
// the same number of particles in every kept cell.  The profiles (density, charge) are the caller's business: it
// hands the flat index of every kept cell inside `box`, its weight per particle and its charge.  The arithmetic is
// written without contraction in the order of the host creator of this repository, so both produce the same doubles.
__global__ void __launch_bounds__( 256 ) k_create_regular( double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
        double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz, double *__restrict__ w, short *__restrict__ q,
        int *__restrict__ key, const int *__restrict__ cells, const double *__restrict__ wcell, const short *__restrict__ qcell,
        size_t ntot, int nppc, int b1, int b2, double o0, double o1, double o2, double d0, double d1, double d2,
        int c0, int c1, int c2, double f0, double f1, double f2 )
{
    for( size_t t = blockIdx.x*( size_t )blockDim.x + threadIdx.x; t < ntot; t += ( size_t )gridDim.x*blockDim.x ) {
        const size_t ic = t / ( size_t )nppc;
        int i = ( int )( t - ic*( size_t )nppc );
        const int cell = cells[ic];
        const int k2 = cell % b2, k1 = ( cell / b2 ) % b1, k0 = cell / ( b2*b1 );
        // origin + ci*cell_length + (cell_length*0.975*inv)*(0.5 + i % c)
        x[t] = __dadd_rn( __dadd_rn( o0, __dmul_rn( ( double )k0, d0 ) ), __dmul_rn( f0, __dadd_rn( 0.5, ( double )( i % c0 ) ) ) );
        i /= c0;
        y[t] = __dadd_rn( __dadd_rn( o1, __dmul_rn( ( double )k1, d1 ) ), __dmul_rn( f1, __dadd_rn( 0.5, ( double )( i % c1 ) ) ) );
        i /= c1;
        z[t] = __dadd_rn( __dadd_rn( o2, __dmul_rn( ( double )k2, d2 ) ), __dmul_rn( f2, __dadd_rn( 0.5, ( double )( i % c2 ) ) ) );
        px[t] = 0.; py[t] = 0.; pz[t] = 0.;
        w[t] = wcell[ic];
        q[t] = qcell[ic];
        key[t] = 0;
    }
}

int sb200_species_append_regular( sb200_patch *p, int ispec, const double origin[3], const int box[3], const int regular_number[3],
                                  const double regular_inv[3], const int *cells, const double *weight, const short *charge, size_t ncells )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec && origin && box && regular_number && regular_inv, "sb200_species_append_regular: bad arguments" );
    if( ncells == 0 ) return 0;
    SB200_CHECK( cells && weight && charge, "sb200_species_append_regular: null cell list" );
    const int nppc = regular_number[0]*regular_number[1]*regular_number[2];
    SB200_CHECK( nppc > 0, "sb200_species_append_regular: regular_number must be positive" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    SpeciesDev &s = p->sp[ispec];
    const size_t n = ncells*( size_t )nppc;
    if( grow_species( p, ispec, s.n + n ) ) return 1;
    // the three per-cell lists travel through the staging buffer: ints | doubles | shorts
    const size_t need = ( ncells*sizeof( int ) + 7 )/8 + ncells + ( ncells*sizeof( short ) + 7 )/8 + 8;
    if( ensure_stage( p, need ) ) return 1;
    int *d_cells = reinterpret_cast<int *>( p->stage );
    double *d_w = p->stage + ( ncells*sizeof( int ) + 7 )/8;
    short *d_q = reinterpret_cast<short *>( d_w + ncells );
    SB200_CUDA( cudaMemcpyAsync( d_cells, cells, ncells*sizeof( int ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( d_w, weight, ncells*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemcpyAsync( d_q, charge, ncells*sizeof( short ), cudaMemcpyHostToDevice, p->stream ) );
    const GridDev &g = p->gd;
    double f[3];
    for( int d=0; d<3; d++ ) f[d] = g.cell[d]*0.975*regular_inv[d];          // left to right, as the host creator
    const unsigned blocks = ( unsigned )( ( n + 255 )/256 < 148*8 ? ( n + 255 )/256 : 148*8 );
    k_create_regular<<<blocks, 256, 0, p->stream>>>( s.col[0] + s.n, s.col[1] + s.n, s.col[2] + s.n, s.col[3] + s.n, s.col[4] + s.n,
            s.col[5] + s.n, s.col[6] + s.n, s.q + s.n, s.key + s.n, d_cells, d_w, d_q, n, nppc, box[1], box[2],
            origin[0], origin[1], origin[2], g.cell[0], g.cell[1], g.cell[2],
            regular_number[0], regular_number[1], regular_number[2], f[0], f[1], f[2] );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    const size_t n0 = s.n;
    s.n += n;
    s.sorted = false;
    s.count_valid = false;
    if( update_qwmax( p, ispec, n0, n ) ) return 1;
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );       // the host lists may be reused by the caller
    return 0;
}

// particles the window leaves behind: dropped on the patch at the left end of the box, tagged for the -x neighbour
// (and listed, as the dynamics kernel lists its leavers) elsewhere
__global__ void __launch_bounds__( 256 ) k_window_tag( const double *__restrict__ x, int *__restrict__ key, size_t n, double xmin_new,
        int tag_left, int *__restrict__ leave_count, int *__restrict__ leave_idx, int leave_cap )
{
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x ) {
        int k = 0;
        if( key[i] < 0 ) k = -1;
        else if( x[i] < xmin_new ) {
            k = tag_left;
            if( tag_left == -2 ) {
                const int c = atomicAdd( leave_count, 1 );
                if( c < leave_cap ) leave_idx[c] = ( int )i;
            }
        }
        key[i] = k;
    }
}

int sb200_window_shift( sb200_patch *p, int ncells )
{
    SB200_CHECK( p && ncells > 0, "sb200_window_shift: bad arguments" );
    SB200_CHECK( ncells < p->gd.n[0], "sb200_window_shift: shift larger than the patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    GridDev &g = p->gd;
    // ---- fields: plane i <- plane i+ncells (x is the slowest dimension: one contiguous block), through the staging buffer
    const size_t plane = ( size_t )g.sx, keep = ( size_t )( g.ax - ncells )*plane;
    if( ensure_stage( p, keep ) ) return 1;
    const int ids[9] = { SB200_EX, SB200_EY, SB200_EZ, SB200_BX, SB200_BY, SB200_BZ, SB200_BXM, SB200_BYM, SB200_BZM };
    for( int f=0; f<9; f++ ) {
        double *a = p->f[ids[f]];
        SB200_CUDA( cudaMemcpyAsync( p->stage, a + ( size_t )ncells*plane, keep*sizeof( double ), cudaMemcpyDeviceToDevice, p->stream ) );
        SB200_CUDA( cudaMemcpyAsync( a, p->stage, keep*sizeof( double ), cudaMemcpyDeviceToDevice, p->stream ) );
        SB200_CUDA( cudaMemsetAsync( a + keep, 0, ( size_t )ncells*plane*sizeof( double ), p->stream ) );
    }
    // ---- origin of the patch (Patch::initStep3 with n_moved, Patch.cpp:159-163)
    p->n_moved += ncells;
    g.begin[0] = g.pcoord[0]*g.n[0] - g.o[0] + p->n_moved;
    g.xmin[0] = ( g.pcoord[0]   )*( g.n[0]*g.cell[0] );
    g.xmax[0] = ( g.pcoord[0]+1 )*( g.n[0]*g.cell[0] );
    g.xmin[0] += p->n_moved*g.cell[0];
    g.xmax[0] += p->n_moved*g.cell[0];
    g.min_loc_round[0] = std::round( g.xmin[0]*g.dxi[0] );
    // ---- particles left behind are dropped; keys are recomputed from the new origin by the next sort
    for( int is=0; is<p->nspec; is++ ) {
        if( materialize( p, is ) ) return 1;
        SpeciesDev &s = p->sp[is];
        SB200_CUDA( cudaMemsetAsync( p->leave_counts + 8*is, 0, 8*sizeof( int ), p->stream ) );
        if( s.n > 0 ) {
            const unsigned blocks = ( unsigned )( ( s.n + 255 )/256 < 148*16 ? ( s.n + 255 )/256 : 148*16 );
            const int tag_left = g.pcoord[0] > 0 ? -2 : -1;      // a -x neighbour takes them over, or they leave the box
            k_window_tag<<<blocks, 256, 0, p->stream>>>( s.col[0], s.key, s.n, g.xmin[0], tag_left, p->leave_counts + 8*is,
                    s.leave_idx, ( int )s.leave_cap );
            sb200::g_launches++;
            SB200_CUDA( cudaGetLastError() );
        }
        s.sorted = false;
        s.count_valid = false;
        s.window_tagged = true;
    }
    return 0;
}

int sb200_species_get( sb200_patch *p, int ispec,
                       double *x, double *y, double *z, double *px, double *py, double *pz,
                       double *w, short *q, int *keys, size_t n )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_species_get: bad species index" );
    SpeciesDev &s = p->sp[ispec];
    SB200_CHECK( n <= s.n, "sb200_species_get: asking for more particles than the species holds" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    double *dst[7] = { x, y, z, px, py, pz, w };
    for( int c=0; c<7; c++ ) if( dst[c] ) SB200_CUDA( cudaMemcpyAsync( dst[c], s.col[c], n*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    if( q ) SB200_CUDA( cudaMemcpyAsync( q, s.q, n*sizeof( short ), cudaMemcpyDeviceToHost, p->stream ) );
    if( keys ) SB200_CUDA( cudaMemcpyAsync( keys, s.key, n*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_species_count( sb200_patch *p, int ispec, size_t *n )
{
    SB200_CHECK( p && n && ispec >= 0 && ispec < p->nspec, "sb200_species_count: bad arguments" );
    *n = p->sp[ispec].n;
    return 0;
}

int sb200_species_device_ptr( sb200_patch *p, int ispec, int column, void **dev_ptr )
{
    SB200_CHECK( p && dev_ptr && ispec >= 0 && ispec < p->nspec, "sb200_species_device_ptr: bad arguments" );
    SpeciesDev &s = p->sp[ispec];
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( materialize( p, ispec ) ) return 1;
    if( column >= 0 && column < 7 ) *dev_ptr = s.col[column];
    else if( column == 7 ) *dev_ptr = s.q;
    else if( column == 8 ) *dev_ptr = s.key;
    else SB200_CHECK( false, "sb200_species_device_ptr: column must be 0..8" );
    return 0;
}

int sb200_species_first_index( sb200_patch *p, int ispec, int *first, size_t n )
{
    SB200_CHECK( p && first && ispec >= 0 && ispec < p->nspec, "sb200_species_first_index: bad arguments" );
    SB200_CHECK( n == p->ncells+1, "sb200_species_first_index: n must be ncells+1" );
    SB200_CHECK( p->sp[ispec].sorted, "sb200_species_first_index: species is not sorted" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SB200_CUDA( cudaMemcpyAsync( first, p->sp[ispec].first, n*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_field_size( sb200_patch *p, int field_id, size_t *n, int dims[3] )
{
    SB200_CHECK( p && field_id >= 0 && field_id < SB200_NFIELDS + 4*p->nspec, "sb200_field_size: bad field id" );
    int d[3];
    field_dims( p->gd, field_id, d );
    if( n ) *n = ( size_t )d[0]*d[1]*d[2];
    if( dims ) { dims[0] = d[0]; dims[1] = d[1]; dims[2] = d[2]; }
    return 0;
}

int sb200_field_set( sb200_patch *p, int field_id, const double *host, size_t n )
{
    SB200_CHECK( p && host && field_ptr( p, field_id ), "sb200_field_set: bad arguments (or a species array that was not requested)" );
    int d[3];
    field_dims( p->gd, field_id, d );
    size_t total = ( size_t )d[0]*d[1]*d[2];
    SB200_CHECK( n == total, "sb200_field_set: size does not match the field dims" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( ensure_stage( p, total ) ) return 1;
    SB200_CUDA( cudaMemcpyAsync( p->stage, host, total*sizeof( double ), cudaMemcpyHostToDevice, p->stream ) );
    SB200_CUDA( cudaMemsetAsync( field_ptr( p, field_id ), 0, p->falloc*sizeof( double ), p->stream ) );
    k_field_pad<<<1184, 256, 0, p->stream>>>( p->stage, field_ptr( p, field_id ), d[0], d[1], d[2], p->gd.sx, p->gd.sy, 1 );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_field_get( sb200_patch *p, int field_id, double *host, size_t n )
{
    SB200_CHECK( p && host && field_ptr( p, field_id ), "sb200_field_get: bad arguments (or a species array that was not requested)" );
    int d[3];
    field_dims( p->gd, field_id, d );
    size_t total = ( size_t )d[0]*d[1]*d[2];
    SB200_CHECK( n == total, "sb200_field_get: size does not match the field dims" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( ensure_stage( p, total ) ) return 1;
    k_field_pad<<<1184, 256, 0, p->stream>>>( p->stage, field_ptr( p, field_id ), d[0], d[1], d[2], p->gd.sx, p->gd.sy, 0 );
    sb200::g_launches++;
    SB200_CUDA( cudaGetLastError() );
    SB200_CUDA( cudaMemcpyAsync( host, p->stage, total*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_field_device_ptr( sb200_patch *p, int field_id, void **dev_ptr, int alloc[3] )
{
    SB200_CHECK( p && dev_ptr && field_ptr( p, field_id ), "sb200_field_device_ptr: bad arguments (or a species array that was not requested)" );
    *dev_ptr = field_ptr( p, field_id );
    if( alloc ) { alloc[0] = p->gd.ax; alloc[1] = p->gd.ay; alloc[2] = p->gd.az; }
    return 0;
}

int sb200_launch_count( unsigned long long *n )
{
    SB200_CHECK( n, "sb200_launch_count: null pointer" );
    *n = g_launches;
    return 0;
}

int sb200_debug_flags( sb200_patch *p, int flags[8] )
{
    SB200_CHECK( p && flags, "sb200_debug_flags: bad arguments" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SB200_CUDA( cudaMemcpyAsync( flags, p->iflags, 8*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_restart_rhoJ( sb200_patch *p )
{
    SB200_CHECK( p, "sb200_restart_rhoJ: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    for( int f=SB200_JX; f<=SB200_RHO; f++ ) SB200_CUDA( cudaMemsetAsync( p->f[f], 0, p->falloc*sizeof( double ), p->stream ) );
    // ElectroMagn::restartRhoJs (ElectroMagn.cpp:410-436): the species' own arrays too
    for( int is=0; is<p->nspec; is++ )
        for( int k=0; k<4; k++ )
            if( p->sp[is].fs[k] ) SB200_CUDA( cudaMemsetAsync( p->sp[is].fs[k], 0, p->falloc*sizeof( double ), p->stream ) );
    return 0;
}

// ElectroMagn::Jx_s .. rho_s of one species: allocated when a field diagnostic asks for them
// (ElectroMagn.cpp: the per-species arrays exist only for the species a DiagFields names)
int sb200_species_diag_fields( sb200_patch *p, int ispec, int mask )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec && mask >= 0 && mask < 16, "sb200_species_diag_fields: bad arguments" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    SpeciesDev &s = p->sp[ispec];
    for( int k=0; k<4; k++ ) {
        const bool want = ( mask >> k ) & 1;
        if( want && !s.fs[k] ) {
            SB200_CUDA( cudaMalloc( &s.fs[k], p->falloc*sizeof( double ) ) );
            SB200_CUDA( cudaMemsetAsync( s.fs[k], 0, p->falloc*sizeof( double ), p->stream ) );
        } else if( !want && s.fs[k] ) {
            SB200_CUDA( cudaStreamSynchronize( p->stream ) );
            cudaFree( s.fs[k] );
            s.fs[k] = nullptr;
        }
    }
    return 0;
}

__global__ void __launch_bounds__( 256 ) k_add_into( double *__restrict__ tot, const double *__restrict__ sp, size_t n )
{
    for( size_t i = blockIdx.x*( size_t )blockDim.x + threadIdx.x; i < n; i += ( size_t )gridDim.x*blockDim.x ) tot[i] += sp[i];
}

// ElectroMagn3D::computeTotalRhoJ (ElectroMagn3D.cpp:1753-1799): totals += every species' own arrays, species by
// species (the padding of the device layout is zero in both, so the whole allocation is added)
int sb200_compute_total_rhoJ( sb200_patch *p )
{
    SB200_CHECK( p, "sb200_compute_total_rhoJ: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    for( int is=0; is<p->nspec; is++ )
        for( int k=0; k<4; k++ )
            if( p->sp[is].fs[k] ) {
                k_add_into<<<148*8, 256, 0, p->stream>>>( p->f[SB200_JX+k], p->sp[is].fs[k], p->falloc );
                sb200::g_launches++;
                SB200_CUDA( cudaGetLastError() );
            }
    return 0;
}

int sb200_dynamics( sb200_patch *p, int ispec, int flags )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_dynamics: bad species index" );
    SB200_CHECK( p->sp[ispec].sorted || p->sp[ispec].n == 0, "sb200_dynamics: species must be cell-sorted (call sb200_sort)" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    if( launch_dynamics( p, ispec, flags ) ) return 1;
    // diag step: rho from the new positions (currentsAndDensity, Projector3D2Order.cpp:509-519)
    if( flags & SB200_DYN_DIAG_RHO ) return launch_rho( p, ispec );      // (the dynamics kernel consumed any pending sort order)
    return 0;
}

int sb200_scratch_get( sb200_patch *p, double *Epart, double *Bpart, double *invgf, int *iold, double *deltaold, size_t n )
{
    SB200_CHECK( p, "sb200_scratch_get: null patch" );
    SB200_CHECK( n <= p->sc_cap && p->sc_E, "sb200_scratch_get: no scratch of that size (run sb200_dynamics with SB200_DYN_KEEP_SCRATCH)" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    // device scratch is component-major with stride sc_n == n of the last dynamics call
    if( Epart ) SB200_CUDA( cudaMemcpyAsync( Epart, p->sc_E, 3*n*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    if( Bpart ) SB200_CUDA( cudaMemcpyAsync( Bpart, p->sc_B, 3*n*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    if( invgf ) SB200_CUDA( cudaMemcpyAsync( invgf, p->sc_invgf, n*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    if( iold ) SB200_CUDA( cudaMemcpyAsync( iold, p->sc_iold, 3*n*sizeof( int ), cudaMemcpyDeviceToHost, p->stream ) );
    if( deltaold ) SB200_CUDA( cudaMemcpyAsync( deltaold, p->sc_delta, 3*n*sizeof( double ), cudaMemcpyDeviceToHost, p->stream ) );
    SB200_CUDA( cudaStreamSynchronize( p->stream ) );
    return 0;
}

int sb200_maxwell( sb200_patch *p )
{
    SB200_CHECK( p, "sb200_maxwell: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    return launch_maxwell( p );
}

int sb200_center_B( sb200_patch *p )
{
    SB200_CHECK( p, "sb200_center_B: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    return launch_center_shell( p );
}

int sb200_sort( sb200_patch *p, int ispec )
{
    SB200_CHECK( p && ispec >= 0 && ispec < p->nspec, "sb200_sort: bad species index" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    return launch_sort( p, ispec );
}

int sb200_energy( sb200_patch *p, double *ukin_per_species, double *uelm )
{
    SB200_CHECK( p, "sb200_energy: null patch" );
    SB200_CUDA( cudaSetDevice( p->device ) );
    return launch_energy( p, ukin_per_species, uelm );
}

} // extern "C"
