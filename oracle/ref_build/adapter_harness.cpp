/* adapter_harness.cpp — EXECUTES include/smilei_b200_operators.hpp.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/ref_build/build_adapter.sh into oracle/_ref/libsmilei_adapter.so from
 *   - the reference's own translation units (the objects of libsmilei_ref.so, compiled in place from /root/reference),
 *   - the object graph of ref_harness.cpp (fabricated Params / Patch / SpeciesV / SmileiMPI, REAL Particles,
 *     Field3D and ElectroMagn3D members),
 *   - the adapter header of this repository, and libsmilei_b200.so (the product's C ABI).
 *
 * adapter_step() instantiates the B200 operator subclasses and drives them THROUGH THE REFERENCE'S BASE-CLASS
 * POINTERS, i.e. through the vtables the reference's factories hand to Species and ElectroMagn:
 *     Interpolator*  -> fieldsWrapper            (Interpolator.h:23;  Species.cpp:591)
 *     Pusher*        -> operator()               (Pusher.h:33;        Species.cpp:727)
 *     Projector*     -> currentsAndDensityWrapper (Projector.h:44;     Species.cpp:782)
 *     Solver*        -> operator() (MA, then MF) (Solver.h:22;        VectorPatch.cpp:1017,1023)
 * in the order of Species::dynamics / VectorPatch::solveMaxwell, on real reference Particles / Field3D objects:
 * they are uploaded with sb200_species_set / sb200_field_set (what Patch::finishCreation + the adapter's
 * Bridge::attach do in a real build) and downloaded afterwards, so that the caller can compare them with what the
 * reference's own operators (ref_interp / ref_push / ref_project / ref_maxwell_* of libsmilei_ref.so) make of the
 * same objects.
 */
#include "ref_harness.cpp"

#include "smilei_b200_operators.hpp"

namespace {
int field_id_of( int k ) { const int ids[12] = { SB200_EX, SB200_EY, SB200_EZ, SB200_BX, SB200_BY, SB200_BZ, SB200_BXM, SB200_BYM, SB200_BZM, SB200_JX, SB200_JY, SB200_JZ }; return ids[k]; }
}

extern "C" {

int adapter_device_count()
{
    int n = 0;
    return sb200_device_count( &n ) == 0 ? n : -1;
}

/* One step of the hot path through the adapter.  fields[12]: Ex Ey Ez Bx By Bz Bxm Bym Bzm Jx Jy Jz in the reference's
 * compact layout (in: state before the step; out: after).  Particles: in = any order; out = after the step, in the
 * cell-sorted order the device works in; x0..pz0 receive the sorted particles BEFORE the push (the order the
 * comparison must use).  Returns the particle count, or -1 with the message of sb200_last_error() on stderr. */
int adapter_step( const orc_grid *g, int order, int pusher, double mass, double *fields[12],
                  int nparts, double *x, double *y, double *z, double *px, double *py, double *pz, double *w, short *q,
                  double *x0, double *y0, double *z0, double *px0, double *py0, double *pz0, int *keys )
{
    using namespace smilei_b200;
    Ctx *c = make_ctx( g, mass, 1 );
    Params &P = *c->params;
    P.interpolation_order = order;
    ElectroMagn3D &E = *c->em;
    Field *F[12] = { E.Ex_, E.Ey_, E.Ez_, E.Bx_, E.By_, E.Bz_, E.Bx_m, E.By_m, E.Bz_m, E.Jx_, E.Jy_, E.Jz_ };
    for( int k=0; k<12; k++ ) load( F[k], fields[k] );
    set_particles( c, x, y, z, px, py, pz, w, q, nparts );
    Particles &part = *c->species->particles;

    /* ---- what Patch::finishCreation does in an adapted build: one device handle per patch, fields and species on it */
    sb200_patch *h = Bridge::attach( P, c->patch, 1, 0 );
    int rc = 0;
    for( int k=0; k<12 && !rc; k++ ) rc = sb200_field_set( h, field_id_of( k ), F[k]->data_, F[k]->number_of_points_ );
    const char *pname[3] = { "boris", "vay", "higueracary" };
    rc = rc || sb200_species_config( h, 0, mass, pusher, ( size_t )nparts + 16 );
    rc = rc || sb200_species_set( h, 0, part.Position[0].data(), part.Position[1].data(), part.Position[2].data(),
                                  part.Momentum[0].data(), part.Momentum[1].data(), part.Momentum[2].data(),
                                  part.Weight.data(), part.Charge.data(), ( size_t )nparts );
    rc = rc || sb200_sort( h, 0 );                /* SpeciesV keeps its particles cell-sorted; so does the device */
    size_t n = 0;
    rc = rc || sb200_species_count( h, 0, &n );
    rc = rc || ( n != ( size_t )nparts );
    rc = rc || sb200_species_get( h, 0, x0, y0, z0, px0, py0, pz0, w, q, keys, n );
    if( rc ) { std::fprintf( stderr, "adapter_step (%s): %s\n", pname[pusher], sb200_last_error() ); return -1; }

    /* ---- the operators, created as the adapted factories create them and held by BASE-CLASS pointers */
    Interpolator *Interp = new Interpolator3DB200( P, c->patch );
    Pusher       *Push   = new PusherB200( P, c->species );
    Projector    *Proj   = new Projector3DB200( P, c->patch );
    Solver       *MA     = new MA_Solver3D_B200( P );
    Solver       *MF     = new MF_Solver3D_B200( P, c->patch );

    /* ---- Species::dynamics (Species.cpp:524-875), one bin [0, n) */
    int istart = 0, iend = ( int )n;
    Interp->fieldsWrapper( &E, part, c->smpi, &istart, &iend, 0 );
    ( *Push )( part, c->smpi, istart, iend, 0 );
    Proj->currentsAndDensityWrapper( &E, part, c->smpi, istart, iend, 0, false, false, 0 );
    /* ---- VectorPatch::solveMaxwell (VectorPatch.cpp:1013-1023) and centerMagneticFields (Smilei.cpp:649) */
    ( *MA )( &E );
    ( *MF )( &E );
    rc = sb200_center_B( h );

    /* ---- back into the reference's objects */
    rc = rc || sb200_species_get( h, 0, x, y, z, px, py, pz, w, q, keys, n );
    for( int k=0; k<12 && !rc; k++ ) {
        rc = sb200_field_get( h, field_id_of( k ), F[k]->data_, F[k]->number_of_points_ );
        store( F[k], fields[k] );
    }
    if( rc ) std::fprintf( stderr, "adapter_step: %s\n", sb200_last_error() );
    delete MF; delete MA; delete Proj; delete Push; delete Interp;
    Bridge::detach( c->patch );
    free_ctx( c );
    return rc ? -1 : ( int )n;
}

}
