// smilei_b200_operators.hpp — the C++ adapter a Smilei build adds to run its 3D Cartesian hot path on
// the B200 library: operator subclasses of the reference's own base classes that forward to the C ABI
// of include/smilei_b200.h.
//
// This header is compiled AGAINST THE REFERENCE TREE (its include paths), nothing in it is used by the
// product library itself.  tests/test_capi_symbols.py type-checks it with the reference headers when
// /root/reference is present, and the adapter harness of the test tree (adapter_harness.cpp) EXECUTES it: the classes below are created
// on the reference's Params / Patch / Species / ElectroMagn3D objects and driven through Interpolator* / Pusher* /
// Projector* / Solver* (tests/test_gpu_parity.py::test_adapter_executes_through_reference_vtable, on a B200).
// INTEGRATION.md shows the four factory branches that return these classes.
//
//   Interpolator3D2OrderB200 / Interpolator3D4OrderB200 : Interpolator3D   (src/Interpolator/Interpolator3D.h)
//   PusherB200                                          : Pusher           (src/Pusher/Pusher.h)
//   Projector3DB200                                     : Projector3D      (src/Projector/Projector3D.h)
//   MA_Solver3D_B200 / MF_Solver3D_B200                 : Solver3D         (src/ElectroMagnSolver/Solver3D.h)
//
// Gather, push, boundary tagging and deposit are ONE kernel on the device (sb200_dynamics).  The
// interpolator and pusher adapters therefore only record that their stage was requested; the projector
// adapter launches the fused kernel.  That is legal because nothing reads the inter-operator scratch
// (SmileiMPI::dynamics_Epart/Bpart/invgf/iold/deltaold) between the three calls when ionization, radiation
// and pair creation are off (src/Species/Species.cpp:596-676) — the adapter refuses to be created otherwise.
#ifndef SMILEI_B200_OPERATORS_HPP
#define SMILEI_B200_OPERATORS_HPP

#include <unordered_map>

#include "smilei_b200.h"

#include "ElectroMagn.h"
#include "Interpolator3D.h"
#include "Params.h"
#include "Particles.h"
#include "ElectroMagnBC3D.h"
#include "Laser.h"
#include "Patch.h"
#include "Projector3D.h"
#include "Pusher.h"
#include "Solver3D.h"
#include "Species.h"
#include "Tools.h"

namespace smilei_b200 {

#define SB200_OR_ERROR( call ) do { if( ( call ) != 0 ) { ERROR( "smilei_b200: " << sb200_last_error() ); } } while( 0 )

//! One device handle per Patch (GPU builds run 1 patch per rank per GPU, doc/Sphinx/Understand/GPU_offloading.rst:40-47).
class Bridge
{
public:
    static std::unordered_map<const Patch *, sb200_patch *> &handles()
    {
        static std::unordered_map<const Patch *, sb200_patch *> h;
        return h;
    }

    //! Called from Patch::finishCreation once Params and the species list are known.
    static sb200_patch *attach( Params &params, Patch *patch, int n_species, int device )
    {
        sb200_grid g;
        for( int i=0; i<3; i++ ) {
            g.n[i]           = ( int )params.patch_size_[i];
            g.oversize[i]    = ( int )params.oversize[i];
            g.cell_length[i] = params.cell_length[i];
            g.pcoord[i]      = ( int )patch->Pcoordinates[i];
            g.npatch[i]      = ( int )params.number_of_patches[i];
        }
        g.dt = params.timestep;
        g.interp_order = ( int )params.interpolation_order;
        sb200_patch *h = NULL;
        SB200_OR_ERROR( sb200_patch_create( &h, &g, n_species, device ) );
        handles()[patch] = h;
        return h;
    }

    static sb200_patch *of( const Patch *patch )
    {
        auto it = handles().find( patch );
        if( it == handles().end() ) {
            ERROR( "smilei_b200: patch has no device handle (Bridge::attach was not called)" );
        }
        return it->second;
    }

    static void detach( Patch *patch )
    {
        auto it = handles().find( patch );
        if( it != handles().end() ) {
            sb200_patch_destroy( it->second );
            handles().erase( it );
        }
    }
};

//! Stage bookkeeping of one (patch, species) inside a step: interpolator -> pusher -> projector.
struct Stage {
    bool interpolated = false;
    bool pushed = false;
};

inline Stage &stage_of( Particles &particles )
{
    static std::unordered_map<const Particles *, Stage> s;
    return s[&particles];
}

// ---------------------------------------------------------------------------------------------------
class Interpolator3DB200 : public Interpolator3D
{
public:
    Interpolator3DB200( Params &, Patch *patch ) : Interpolator3D( patch ) {}

    void fieldsWrapper( ElectroMagn *, Particles &particles, SmileiMPI *, int *, int *, int, unsigned int = 0, int = 0 ) override
    {
        stage_of( particles ).interpolated = true;      // fused: see the header comment
    }
    void fieldsAndCurrents( ElectroMagn *, Particles &, SmileiMPI *, int *, int *, int, LocalFields *, double * ) override
    {
        ERROR( "smilei_b200: fieldsAndCurrents (probe/ionization interpolation) is not on the B200 hot path" );
    }
    void fieldsSelection( ElectroMagn *, Particles &, double *, int, std::vector<unsigned int> * ) override
    {
        ERROR( "smilei_b200: fieldsSelection (tracked particles) is not on the B200 hot path" );
    }
    void oneField( Field **, Particles &, int *, int *, double *, double * = NULL, double * = NULL, double * = NULL ) override
    {
        ERROR( "smilei_b200: oneField is not on the B200 hot path" );
    }
};
typedef Interpolator3DB200 Interpolator3D2OrderB200;
typedef Interpolator3DB200 Interpolator3D4OrderB200;

// ---------------------------------------------------------------------------------------------------
class PusherB200 : public Pusher
{
public:
    PusherB200( Params &params, Species *species ) : Pusher( params, species ) {}

    void operator()( Particles &particles, SmileiMPI *, int, int, int, int = 0 ) override
    {
        Stage &s = stage_of( particles );
        if( !s.interpolated ) {
            ERROR( "smilei_b200: pusher called before the interpolator of the same step" );
        }
        s.pushed = true;
    }
};

// ---------------------------------------------------------------------------------------------------
class Projector3DB200 : public Projector3D
{
public:
    Projector3DB200( Params &params, Patch *patch ) : Projector3D( params, patch ), patch_( patch ) {}

    void currentsAndDensityWrapper( ElectroMagn *, Particles &particles, SmileiMPI *, int, int, int, bool diag_flag, bool is_spectral,
                                    int ispec, int = 0, int = 0 ) override
    {
        Stage &s = stage_of( particles );
        if( !s.interpolated || !s.pushed ) {
            ERROR( "smilei_b200: projector called without the interpolator and pusher stages of the same step" );
        }
        if( is_spectral ) {
            ERROR( "smilei_b200: spectral solvers are not on the B200 hot path" );
        }
        SB200_OR_ERROR( sb200_dynamics( Bridge::of( patch_ ), ispec, diag_flag ? SB200_DYN_DIAG_RHO : 0 ) );
        s = Stage();
    }
    void ionizationCurrents( Field *, Field *, Field *, Particles &, int, LocalFields ) override
    {
        ERROR( "smilei_b200: ionization is not on the B200 hot path" );
    }

private:
    Patch *patch_;
};

// ---------------------------------------------------------------------------------------------------
//! MA and MF are launched together (E sweep then B sweep, B_m centred in the same pass): the Ampère
//! adapter arms, the Faraday adapter fires; VectorPatch::solveMaxwell calls them in that order
//! (src/Patch/VectorPatch.cpp:1017,1023).
class MA_Solver3D_B200 : public Solver3D
{
public:
    MA_Solver3D_B200( Params &params ) : Solver3D( params ) {}
    void operator()( ElectroMagn *fields ) override { armed()[fields] = true; }
    static std::unordered_map<const ElectroMagn *, bool> &armed()
    {
        static std::unordered_map<const ElectroMagn *, bool> a;
        return a;
    }
};

class MF_Solver3D_B200 : public Solver3D
{
public:
    MF_Solver3D_B200( Params &params, Patch *patch ) : Solver3D( params ), patch_( patch ) {}
    void operator()( ElectroMagn *fields ) override
    {
        if( !MA_Solver3D_B200::armed()[fields] ) {
            ERROR( "smilei_b200: MF solver called before the MA solver of the same step" );
        }
        MA_Solver3D_B200::armed()[fields] = false;
        SB200_OR_ERROR( sb200_maxwell( Bridge::of( patch_ ) ) );
    }

private:
    Patch *patch_;
};

// ---------------------------------------------------------------------------------------------------
//! Silver-Mueller side: replaces ElectroMagnBC3D_SM (src/ElectroMagnBC/ElectroMagnBC3D_SM.cpp) where
//! ElectroMagnBC_Factory::create (src/ElectroMagnBC/ElectroMagnBC_Factory.h) returns it for "silver-muller".
//! The laser amplitudes stay the reference's business (Laser / LaserProfile objects of vecLaser, Python profiles
//! included): they are summed on the host exactly as ElectroMagnBC3D_SM::apply does (:189-199, :265-275) and
//! handed to the device kernel.  External fields (B_val) are not on this path.
class ElectroMagnBC3D_SM_B200 : public ElectroMagnBC3D
{
public:
    ElectroMagnBC3D_SM_B200( Params &params, Patch *patch, unsigned int i_boundary )
        : ElectroMagnBC3D( params, patch, i_boundary )
    {
        axis0_ = i_boundary / 2;
        axis1_ = axis0_ == 0 ? 1 : 0;
        axis2_ = axis0_ == 2 ? 1 : 2;
        for( int i=0; i<3; i++ ) k_[i] = params.EM_BCs_k[i_boundary][i];
    }
    void apply( ElectroMagn *EMfields, double time_dual, Patch *patch ) override
    {
        if( !patch->isBoundary( i_boundary_ ) ) return;
        const int isb[4] = { patch->isBoundary( axis1_, 0 ), patch->isBoundary( axis1_, 1 ),
                             patch->isBoundary( axis2_, 0 ), patch->isBoundary( axis2_, 1 ) };
        const unsigned int n1p = n_p[axis1_], n1d = n_d[axis1_], n2p = n_p[axis2_], n2d = n_d[axis2_];
        std::vector<double> b1, b2, pos( 2 );
        if( !vecLaser.empty() ) {
            b1.assign( n1p*n2d, 0. );
            b2.assign( n1d*n2p, 0. );
            for( unsigned int j=isb[0]; j<n1p-isb[1]; j++ ) {
                pos[0] = patch->getDomainLocalMin( axis1_ ) + ( ( int )j - ( int )EMfields->oversize[axis1_] )*d[axis1_];
                for( unsigned int k=isb[2]; k<n2d-isb[3]; k++ ) {
                    pos[1] = patch->getDomainLocalMin( axis2_ ) + ( ( int )k - 0.5 - ( int )EMfields->oversize[axis2_] )*d[axis2_];
                    for( unsigned int il=0; il<vecLaser.size(); il++ ) b1[j*n2d+k] += vecLaser[il]->getAmplitude0( pos, time_dual, j, k );
                }
            }
            for( unsigned int j=isb[0]; j<n1d-isb[1]; j++ ) {
                pos[0] = patch->getDomainLocalMin( axis1_ ) + ( ( int )j - 0.5 - ( int )EMfields->oversize[axis1_] )*d[axis1_];
                for( unsigned int k=isb[2]; k<n2p-isb[3]; k++ ) {
                    pos[1] = patch->getDomainLocalMin( axis2_ ) + ( ( int )k - ( int )EMfields->oversize[axis2_] )*d[axis2_];
                    for( unsigned int il=0; il<vecLaser.size(); il++ ) b2[j*n2p+k] += vecLaser[il]->getAmplitude1( pos, time_dual, j, k );
                }
            }
        }
        SB200_OR_ERROR( sb200_apply_SM( Bridge::of( patch ), ( int )i_boundary_, k_, isb,
                                        b1.empty() ? NULL : b1.data(), b2.empty() ? NULL : b2.data() ) );
    }
    void save_fields( Field *, Patch * ) override {}
    void disableExternalFields() override {}

private:
    unsigned int axis0_, axis1_, axis2_;
    double k_[3];
};

//! Particle boundary conditions of a species: PartBoundCond (src/ParticleBC/PartBoundCond.cpp:99-245) chooses per
//! box side among internal / remove / reflective ...; the fused kernel applies `periodic` and `remove` itself, so the
//! adapter only forwards the species' choice (called from Species::initOperators next to the three factories).
inline void forward_particle_bc( Patch *patch, Species *species, int ispec )
{
    int bc[6];
    for( int d=0; d<3; d++ ) {
        for( int s=0; s<2; s++ ) {
            const std::string &name = species->boundary_conditions_[d][s];
            if( name == "periodic" ) bc[2*d+s] = SB200_PBC_PERIODIC;
            else if( name == "remove" ) bc[2*d+s] = SB200_PBC_REMOVE;
            else { ERROR( "smilei_b200: particle boundary condition `" << name << "` is not on the B200 path (periodic, remove)" ); }
        }
    }
    SB200_OR_ERROR( sb200_species_set_bc( Bridge::of( patch ), ispec, bc ) );
}

} // namespace smilei_b200

#endif
