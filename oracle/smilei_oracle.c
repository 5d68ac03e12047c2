/* smilei_oracle.c — CPU restatement of Smilei's 3D Cartesian PIC hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see smilei_oracle.h).  Parity status: PINNED against the
 * reference's own translation units (oracle/_ref) — tests/test_oracle_vs_ref.py.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/src).  Operation order follows the reference statement by
 * statement so that, compiled without FMA contraction (-ffp-contract=off), results
 * are bit-identical to the reference compiled the same way.
 */
#include "smilei_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

/* ------------------------------------------------------------------------- */
/* geometry                                                                   */
/* ------------------------------------------------------------------------- */

/* Patch::initStep3, Patch/Patch.cpp:146-153 */
void orc_patch_bounds( const orc_grid *g, double *min_local, double *max_local, int *cell_start_gc )
{
    for( int i=0; i<3; i++ ) {
        min_local[i] = ( g->pcoord[i]   )*( g->n[i]*g->cell[i] );
        max_local[i] = ( g->pcoord[i]+1 )*( g->n[i]*g->cell[i] );
        cell_start_gc[i] = g->pcoord[i]*g->n[i] - g->o[i];
    }
    /* moving window: Patch::initStep3 with n_moved, Patch/Patch.cpp:159-163 */
    cell_start_gc[0] += g->n_moved;
    min_local[0] += g->n_moved*g->cell[0];
    max_local[0] += g->n_moved*g->cell[0];
}

/* ElectroMagn::ElectroMagn, ElectroMagn/ElectroMagn.cpp:44-49 */
void orc_dims( const orc_grid *g, int *p, int *d )
{
    for( int i=0; i<3; i++ ) {
        p[i] = g->n[i] + 2*g->o[i] + 1;
        d[i] = g->n[i] + 2*g->o[i] + 2;
    }
}

static void comp_dims( const orc_grid *g, int dualx, int dualy, int dualz, int *dims )
{
    int p[3], d[3];
    orc_dims( g, p, d );
    dims[0] = dualx ? d[0] : p[0];
    dims[1] = dualy ? d[1] : p[1];
    dims[2] = dualz ? d[2] : p[2];
}

/* ElectroMagn3D::initElectroMagn3DQuantities, ElectroMagn/ElectroMagn3D.cpp:115-123 */
long orc_field_size( const orc_grid *g, int id )
{
    static const int dual[7][3] = { {1,0,0},{0,1,0},{0,0,1}, {0,1,1},{1,0,1},{1,1,0}, {0,0,0} };
    int dims[3];
    comp_dims( g, dual[id][0], dual[id][1], dual[id][2], dims );
    return ( long )dims[0]*dims[1]*dims[2];
}

/* ------------------------------------------------------------------------- */
/* a3-a6 gather                                                               */
/* ------------------------------------------------------------------------- */

/* Interpolator3D2Order::coeffs weights, Interpolator/Interpolator3D2Order.h:107-123 */
static inline void w2( double delta, double *c )
{
    double delta2 = delta*delta;
    c[0] = 0.5 * ( delta2-delta+0.25 );
    c[1] = 0.75 - delta2;
    c[2] = 0.5 * ( delta2+delta+0.25 );
}

/* Interpolator3D4Order::coeffs weights, Interpolator/Interpolator3D4Order.h:69-73;
 * constants as stored doubles, Interpolator3D4Order.cpp:24-34 */
static inline void w4( double delta, double *c )
{
    const double dble_1_ov_384 = 1.0/384.0, dble_1_ov_48 = 1.0/48.0, dble_1_ov_16 = 1.0/16.0,
                 dble_1_ov_12 = 1.0/12.0, dble_1_ov_24 = 1.0/24.0, dble_19_ov_96 = 19.0/96.0,
                 dble_11_ov_24 = 11.0/24.0, dble_1_ov_4 = 1.0/4.0, dble_1_ov_6 = 1.0/6.0,
                 dble_115_ov_192 = 115.0/192.0, dble_5_ov_8 = 5.0/8.0;
    double delta2 = delta*delta;
    double delta3 = delta2*delta;
    double delta4 = delta3*delta;
    c[0] = dble_1_ov_384   - dble_1_ov_48  * delta  + dble_1_ov_16 * delta2 - dble_1_ov_12 * delta3 + dble_1_ov_24 * delta4;
    c[1] = dble_19_ov_96   - dble_11_ov_24 * delta  + dble_1_ov_4  * delta2 + dble_1_ov_6  * delta3 - dble_1_ov_6  * delta4;
    c[2] = dble_115_ov_192 - dble_5_ov_8   * delta2 + dble_1_ov_4  * delta4;
    c[3] = dble_19_ov_96   + dble_11_ov_24 * delta  + dble_1_ov_4  * delta2 - dble_1_ov_6  * delta3 - dble_1_ov_6  * delta4;
    c[4] = dble_1_ov_384   + dble_1_ov_48  * delta  + dble_1_ov_16 * delta2 + dble_1_ov_12 * delta3 + dble_1_ov_24 * delta4;
}

/* Interpolator3D2Order::compute, Interpolator3D2Order.h:54-76 (half = 1) and
 * Interpolator3D4Order::compute, Interpolator3D4Order.h:40-52 (half = 2):
 * accumulation order iloc -> jloc -> kloc, product left to right. */
static inline double compute( int half, const double *cx, const double *cy, const double *cz,
                              const double *f, int idx, int idy, int idz, int ny, int nz )
{
    double interp_res = 0.;
    for( int iloc=-half ; iloc<=half ; iloc++ ) {
        for( int jloc=-half ; jloc<=half ; jloc++ ) {
            for( int kloc=-half ; kloc<=half ; kloc++ ) {
                interp_res += cx[iloc] * cy[jloc] * cz[kloc] * f[ ( idx+iloc )*ny*nz + ( idy+jloc )*nz + ( idz+kloc ) ];
            }
        }
    }
    return interp_res;
}

/* Interpolator3D2Order::fieldsWrapper, Interpolator3D2Order.cpp:163-285;
 * Interpolator3D4Order::fieldsWrapper, Interpolator3D4Order.cpp:159-217 */
void orc_interp( const orc_grid *g, int order,
                 const double *Ex, const double *Ey, const double *Ez,
                 const double *Bxm, const double *Bym, const double *Bzm,
                 const double *x, const double *y, const double *z, int nparts, int istart, int iend,
                 double *Epart, double *Bpart, int *iold, double *deltaold )
{
    double d_inv[3], mn[3], mx[3];
    int begin[3], p[3], d[3];
    for( int i=0; i<3; i++ ) d_inv[i] = 1.0/g->cell[i];            /* Interpolator3D2Order.cpp:18-20 */
    orc_patch_bounds( g, mn, mx, begin );                          /* Interpolator3D.cpp:15-17       */
    orc_dims( g, p, d );
    const int half = order/2;
    const int nw = order+1;

    for( int ipart=istart ; ipart<iend ; ipart++ ) {
        const double pn[3] = { x[ipart]*d_inv[0], y[ipart]*d_inv[1], z[ipart]*d_inv[2] };
        int idx_p[3], idx_d[3];
        double delta_p[3];
        double cp[3][5], cd[3][5];
        for( int c=0; c<3; c++ ) {
            idx_p[c] = ( int )round( pn[c] );
            delta_p[c] = pn[c] - ( double )idx_p[c];
            if( order==2 ) w2( delta_p[c], cp[c] ); else w4( delta_p[c], cp[c] );
            idx_p[c] = idx_p[c] - begin[c];
        }
        for( int c=0; c<3; c++ ) {
            idx_d[c] = ( int )round( pn[c]+0.5 );
            double delta = pn[c] - ( double )idx_d[c] + 0.5;
            if( order==2 ) w2( delta, cd[c] ); else w4( delta, cd[c] );
            idx_d[c] = idx_d[c] - begin[c];
        }
        ( void )nw;
        const double *cxp = &cp[0][half], *cyp = &cp[1][half], *czp = &cp[2][half];
        const double *cxd = &cd[0][half], *cyd = &cd[1][half], *czd = &cd[2][half];
        /* Ex(d,p,p) Ey(p,d,p) Ez(p,p,d) Bx(p,d,d) By(d,p,d) Bz(d,d,p) */
        Epart[0*nparts+ipart] = compute( half, cxd, cyp, czp, Ex,  idx_d[0], idx_p[1], idx_p[2], p[1], p[2] );
        Epart[1*nparts+ipart] = compute( half, cxp, cyd, czp, Ey,  idx_p[0], idx_d[1], idx_p[2], d[1], p[2] );
        Epart[2*nparts+ipart] = compute( half, cxp, cyp, czd, Ez,  idx_p[0], idx_p[1], idx_d[2], p[1], d[2] );
        Bpart[0*nparts+ipart] = compute( half, cxp, cyd, czd, Bxm, idx_p[0], idx_d[1], idx_d[2], d[1], d[2] );
        Bpart[1*nparts+ipart] = compute( half, cxd, cyp, czd, Bym, idx_d[0], idx_p[1], idx_d[2], p[1], d[2] );
        Bpart[2*nparts+ipart] = compute( half, cxd, cyd, czp, Bzm, idx_d[0], idx_d[1], idx_p[2], d[1], p[2] );
        for( int c=0; c<3; c++ ) {
            iold[c*nparts+ipart]     = idx_p[c];
            deltaold[c*nparts+ipart] = delta_p[c];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a7-a9 push                                                                 */
/* ------------------------------------------------------------------------- */

void orc_push( const orc_grid *g, int pusher, double mass,
               double *position_x, double *position_y, double *position_z,
               double *momentum_x, double *momentum_y, double *momentum_z,
               const short *charge, int nparts, int istart, int iend,
               const double *Epart, const double *Bpart, double *invgf )
{
    /* Pusher::Pusher, Pusher/Pusher.cpp:19-27 */
    const double one_over_mass_ = mass > 0. ? 1.0/mass : 0.;
    const double dt   = g->dt;
    const double dts2 = g->dt/2.;
    const double *Ex = &Epart[0*nparts], *Ey = &Epart[1*nparts], *Ez = &Epart[2*nparts];
    const double *Bx = &Bpart[0*nparts], *By = &Bpart[1*nparts], *Bz = &Bpart[2*nparts];

    if( pusher==0 ) {
        /* PusherBoris::operator(), Pusher/PusherBoris.cpp:82-127 */
        for( int ipart=istart ; ipart<iend; ipart++ ) {
            const double charge_over_mass_dts2 = ( double )( charge[ipart] )*one_over_mass_*dts2;
            double pxsm = charge_over_mass_dts2*( Ex[ipart] );
            double pysm = charge_over_mass_dts2*( Ey[ipart] );
            double pzsm = charge_over_mass_dts2*( Ez[ipart] );
            const double umx = momentum_x[ipart] + pxsm;
            const double umy = momentum_y[ipart] + pysm;
            const double umz = momentum_z[ipart] + pzsm;
            double local_invgf     = charge_over_mass_dts2 / sqrt( 1.0 + umx*umx + umy*umy + umz*umz );
            const double Tx        = local_invgf * ( Bx[ipart] );
            const double Ty        = local_invgf * ( By[ipart] );
            const double Tz        = local_invgf * ( Bz[ipart] );
            const double inv_det_T = 1.0/( 1.0+Tx*Tx+Ty*Ty+Tz*Tz );
            pxsm += ( ( 1.0+Tx*Tx-Ty*Ty-Tz*Tz )* umx  +      2.0*( Tx*Ty+Tz )* umy  +      2.0*( Tz*Tx-Ty )* umz )*inv_det_T;
            pysm += ( 2.0*( Tx*Ty-Tz )* umx  + ( 1.0-Tx*Tx+Ty*Ty-Tz*Tz )* umy  +      2.0*( Ty*Tz+Tx )* umz )*inv_det_T;
            pzsm += ( 2.0*( Tz*Tx+Ty )* umx  +      2.0*( Ty*Tz-Tx )* umy  + ( 1.0-Tx*Tx-Ty*Ty+Tz*Tz )* umz )*inv_det_T;
            local_invgf = 1. / sqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
            invgf[ipart] = local_invgf;
            momentum_x[ipart] = pxsm;
            momentum_y[ipart] = pysm;
            momentum_z[ipart] = pzsm;
            local_invgf *= dt;
            position_x[ipart] += pxsm*local_invgf;
            position_y[ipart] += pysm*local_invgf;
            position_z[ipart] += pzsm*local_invgf;
        }
    } else if( pusher==1 ) {
        /* PusherVay::operator(), Pusher/PusherVay.cpp:92-170 */
        for( int ipart=istart ; ipart<iend; ipart++ ) {
            const double charge_over_mass_dts2 = ( double )( charge[ipart] )*one_over_mass_*dts2;
            invgf[ipart] = 1./sqrt( 1.0 + momentum_x[ipart]*momentum_x[ipart]
                                    + momentum_y[ipart]*momentum_y[ipart]
                                    + momentum_z[ipart]*momentum_z[ipart] );
            double upx = momentum_x[ipart] + 2.*charge_over_mass_dts2*( Ex[ipart] );
            double upy = momentum_y[ipart] + 2.*charge_over_mass_dts2*( Ey[ipart] );
            double upz = momentum_z[ipart] + 2.*charge_over_mass_dts2*( Ez[ipart] );
            double Tx  = charge_over_mass_dts2* ( Bx[ipart] );
            double Ty  = charge_over_mass_dts2* ( By[ipart] );
            double Tz  = charge_over_mass_dts2* ( Bz[ipart] );
            upx += invgf[ipart]*( momentum_y[ipart]*Tz - momentum_z[ipart]*Ty );
            upy += invgf[ipart]*( momentum_z[ipart]*Tx - momentum_x[ipart]*Tz );
            upz += invgf[ipart]*( momentum_x[ipart]*Ty - momentum_y[ipart]*Tx );
            double alpha = 1.0 + upx*upx + upy*upy + upz*upz;
            const double T2    = Tx*Tx + Ty*Ty + Tz*Tz;
            double s     = alpha - T2;
            double us2   = upx*Tx + upy*Ty + upz*Tz;
            us2   = us2*us2;
            alpha = 1.0/sqrt( 0.5*( s + sqrt( s*s + 4.0*( T2 + us2 ) ) ) );
            Tx *= alpha;
            Ty *= alpha;
            Tz *= alpha;
            s = 1.0/( 1.0+Tx*Tx+Ty*Ty+Tz*Tz );
            alpha   = upx*Tx + upy*Ty + upz*Tz;
            const double pxsm = s*( upx + alpha*Tx + Tz*upy - Ty*upz );
            const double pysm = s*( upy + alpha*Ty + Tx*upz - Tz*upx );
            const double pzsm = s*( upz + alpha*Tz + Ty*upx - Tx*upy );
            invgf[ipart] = 1.0 / sqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
            momentum_x[ipart] = pxsm;
            momentum_y[ipart] = pysm;
            momentum_z[ipart] = pzsm;
            position_x[ipart] += dt*momentum_x[ipart]*invgf[ipart];
            position_y[ipart] += dt*momentum_y[ipart]*invgf[ipart];
            position_z[ipart] += dt*momentum_z[ipart]*invgf[ipart];
        }
    } else {
        /* PusherHigueraCary::operator(), Pusher/PusherHigueraCary.cpp:93-162 */
        for( int ipart=istart ; ipart<iend; ipart++ ) {
            const double charge_over_mass_dts2 = ( double )( charge[ipart] )*one_over_mass_*dts2;
            double pxsm = charge_over_mass_dts2*( Ex[ipart] );
            double pysm = charge_over_mass_dts2*( Ey[ipart] );
            double pzsm = charge_over_mass_dts2*( Ez[ipart] );
            const double umx = momentum_x[ipart] + pxsm;
            const double umy = momentum_y[ipart] + pysm;
            const double umz = momentum_z[ipart] + pzsm;
            const double gfm2 = ( 1.0 + umx*umx + umy*umy + umz*umz );
            double Tx    = charge_over_mass_dts2 * ( Bx[ipart] );
            double Ty    = charge_over_mass_dts2 * ( By[ipart] );
            double Tz    = charge_over_mass_dts2 * ( Bz[ipart] );
            const double beta2 = Tx*Tx + Ty*Ty + Tz*Tz;
            const double Tum = Tx*umx + Ty*umy + Tz*umz;
            const double local_invgf = 1./sqrt( 0.5*( gfm2 - beta2 +
                                                sqrt( ( gfm2 - beta2 )*( gfm2 - beta2 ) + 4.0*( beta2 + Tum * Tum ) ) ) );
            Tx    *= local_invgf;
            Ty    *= local_invgf;
            Tz    *= local_invgf;
            const double Tx2   = Tx*Tx;
            const double Ty2   = Ty*Ty;
            const double Tz2   = Tz*Tz;
            const double TxTy  = Tx*Ty;
            const double TyTz  = Ty*Tz;
            const double TzTx  = Tz*Tx;
            const double inv_det_T = 1.0/( 1.0+Tx2+Ty2+Tz2 );
            const double upx = ( ( 1.0+Tx2-Ty2-Tz2 )* umx  +      2.0*( TxTy+Tz )* umy  +      2.0*( TzTx-Ty )* umz )*inv_det_T;
            const double upy = ( 2.0*( TxTy-Tz )* umx  + ( 1.0-Tx2+Ty2-Tz2 )* umy  +      2.0*( TyTz+Tx )* umz )*inv_det_T;
            const double upz = ( 2.0*( TzTx+Ty )* umx  +      2.0*( TyTz-Tx )* umy  + ( 1.0-Tx2-Ty2+Tz2 )* umz )*inv_det_T;
            pxsm += upx;
            pysm += upy;
            pzsm += upz;
            invgf[ipart] = 1. / sqrt( 1.0 + pxsm*pxsm + pysm*pysm + pzsm*pzsm );
            momentum_x[ipart] = pxsm;
            momentum_y[ipart] = pysm;
            momentum_z[ipart] = pzsm;
            position_x[ipart] += dt*momentum_x[ipart]*invgf[ipart];
            position_y[ipart] += dt*momentum_y[ipart]*invgf[ipart];
            position_z[ipart] += dt*momentum_z[ipart]*invgf[ipart];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a10 boundary tagging                                                       */
/* ------------------------------------------------------------------------- */

/* PartBoundCond::apply, ParticleBC/PartBoundCond.h:38-76 with bc_* = internal_inf /
 * internal_sup (ParticleBC/BoundaryConditionType.cpp:15-57); limits are the patch
 * bounds when the EM BC is periodic (ParticleBC/PartBoundCond.cpp:44-69). */
void orc_bc_tag( const orc_grid *g, const double *x, const double *y, const double *z,
                 int *cell_keys, int imin, int imax )
{
    double mn[3], mx[3];
    int begin[3];
    orc_patch_bounds( g, mn, mx, begin );
    const double *pos[3] = { x, y, z };
    for( int ipart=imin; ipart<imax; ipart++ ) cell_keys[ipart] = 0;
    for( int direction=0; direction<3; direction++ ) {
        for( int ipart=imin ; ipart<imax ; ipart++ ) {
            if( cell_keys[ ipart ] >= 0 && pos[direction][ ipart ] < mn[direction] ) {
                cell_keys[ ipart ] = -2 - 2 * direction;
            }
        }
        for( int ipart=imin ; ipart<imax ; ipart++ ) {
            if( cell_keys[ ipart ] >= 0 && pos[direction][ ipart ] >= mx[direction] ) {
                cell_keys[ ipart ] = -3 - 2 * direction;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a11-a14 deposit                                                            */
/* ------------------------------------------------------------------------- */

static const double one_third = 1./3.;   /* Projector/Projector3D.h:47 */

/* Projector3D2Order::currents, Projector/Projector3D2Order.cpp:55-343 */
static void currents_o2( double *Jx, double *Jy, double *Jz,
                         double xp, double yp, double zp, short charge, double weight,
                         const int *iold, const double *deltaold, int nparts,
                         double inv_cell_volume, const double *d_inv, const double *d_ov_dt,
                         const int *begin, int nprimy, int nprimz )
{
    double charge_weight = inv_cell_volume * ( double )( charge )*weight;
    double crx_p = charge_weight*d_ov_dt[0];
    double cry_p = charge_weight*d_ov_dt[1];
    double crz_p = charge_weight*d_ov_dt[2];

    double xpn, ypn, zpn;
    double delta, delta2;
    double Sx0[5], Sx1[5], Sy0[5], Sy1[5], Sz0[5], Sz1[5], DSx[5], DSy[5], DSz[5];
    double tmpJx[5][5], tmpJy[5][5], tmpJz[5][5];

    for( unsigned int i=0; i<5; i++ ) {
        Sx1[i] = 0.;
        Sy1[i] = 0.;
        Sz1[i] = 0.;
    }
    memset( tmpJx, 0, sizeof( tmpJx ) );
    memset( tmpJy, 0, sizeof( tmpJy ) );
    memset( tmpJz, 0, sizeof( tmpJz ) );

    delta = deltaold[0*nparts];
    delta2 = delta*delta;
    Sx0[0] = 0.;
    Sx0[1] = 0.5 * ( delta2-delta+0.25 );
    Sx0[2] = 0.75-delta2;
    Sx0[3] = 0.5 * ( delta2+delta+0.25 );
    Sx0[4] = 0.;

    delta = deltaold[1*nparts];
    delta2 = delta*delta;
    Sy0[0] = 0.;
    Sy0[1] = 0.5 * ( delta2-delta+0.25 );
    Sy0[2] = 0.75-delta2;
    Sy0[3] = 0.5 * ( delta2+delta+0.25 );
    Sy0[4] = 0.;

    delta = deltaold[2*nparts];
    delta2 = delta*delta;
    Sz0[0] = 0.;
    Sz0[1] = 0.5 * ( delta2-delta+0.25 );
    Sz0[2] = 0.75-delta2;
    Sz0[3] = 0.5 * ( delta2+delta+0.25 );
    Sz0[4] = 0.;

    xpn = xp * d_inv[0];
    int ip = ( int )round( xpn );
    int ipo = iold[0*nparts];
    int ip_m_ipo = ip-ipo-begin[0];
    delta  = xpn - ( double )ip;
    delta2 = delta*delta;
    Sx1[ip_m_ipo+1] = 0.5 * ( delta2-delta+0.25 );
    Sx1[ip_m_ipo+2] = 0.75-delta2;
    Sx1[ip_m_ipo+3] = 0.5 * ( delta2+delta+0.25 );

    ypn = yp * d_inv[1];
    int jp = ( int )round( ypn );
    int jpo = iold[1*nparts];
    int jp_m_jpo = jp-jpo-begin[1];
    delta  = ypn - ( double )jp;
    delta2 = delta*delta;
    Sy1[jp_m_jpo+1] = 0.5 * ( delta2-delta+0.25 );
    Sy1[jp_m_jpo+2] = 0.75-delta2;
    Sy1[jp_m_jpo+3] = 0.5 * ( delta2+delta+0.25 );

    zpn = zp * d_inv[2];
    int kp = ( int )round( zpn );
    int kpo = iold[2*nparts];
    int kp_m_kpo = kp-kpo-begin[2];
    delta  = zpn - ( double )kp;
    delta2 = delta*delta;
    Sz1[kp_m_kpo+1] = 0.5 * ( delta2-delta+0.25 );
    Sz1[kp_m_kpo+2] = 0.75-delta2;
    Sz1[kp_m_kpo+3] = 0.5 * ( delta2+delta+0.25 );

    for( unsigned int i=0; i < 5; i++ ) {
        DSx[i] = Sx1[i] - Sx0[i];
        DSy[i] = Sy1[i] - Sy0[i];
        DSz[i] = Sz1[i] - Sz0[i];
    }

    ipo -= 2;
    jpo -= 2;
    kpo -= 2;

    int linindex, linindex_x, linindex_y;
    double tmp, tmp2;
    double vtmp[5];

    /* Jx^(d,p,p) */
    int  z_size = nprimz;
    int yz_size = nprimz*nprimy;
    int linindex0 = ipo*yz_size+jpo*z_size+kpo;
    tmp = 0.;
    linindex = linindex0;
    tmp2 = crx_p * ( one_third*Sy1[0]*Sz1[0] );
    for( int i=1 ; i<5 ; i++ ) {
        tmp -= DSx[i-1] * tmp2;
        linindex += yz_size;
        Jx [linindex] += tmp;
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int k=1 ; k<5 ; k++ ) {
        linindex_x += 1;
        linindex    = linindex_x;
        tmp = crx_p * ( 0.5*Sy1[0]*Sz0[k] + one_third*Sy1[0]*DSz[k] );
        for( int i=1 ; i<5 ; i++ ) {
            vtmp[k] -= DSx[i-1] * tmp;
            linindex += yz_size;
            Jx [linindex] += vtmp[k];
        }
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int j=1 ; j<5 ; j++ ) {
        linindex_x += z_size;
        linindex    = linindex_x;
        tmp = crx_p * ( 0.5*Sz1[0]*Sy0[j] + one_third*DSy[j]*Sz1[0] );
        for( int i=1 ; i<5 ; i++ ) {
            vtmp[j] -= DSx[i-1] * tmp;
            linindex += yz_size;
            Jx [linindex] += vtmp[j];
        }
    }
    linindex_x = linindex0;
    for( int j=1 ; j<5 ; j++ ) {
        linindex_x += z_size;
        linindex_y  = linindex_x;
        for( int k=1 ; k<5 ; k++ ) {
            linindex_y += 1;
            linindex    = linindex_y;
            tmp = crx_p * ( Sy0[j]*Sz0[k] + 0.5*DSy[j]*Sz0[k] + 0.5*DSz[k]*Sy0[j] + one_third*DSy[j]*DSz[k] );
            for( int i=1 ; i<5 ; i++ ) {
                tmpJx[j][k] -= DSx[i-1] * tmp;
                linindex += yz_size;
                Jx [linindex] += tmpJx[j][k];
            }
        }
    }

    /* Jy^(p,d,p) */
    yz_size = nprimz*( nprimy+1 );
    linindex0 = ipo*yz_size+jpo*z_size+kpo;
    tmp = 0.;
    linindex = linindex0;
    tmp2 = cry_p * ( one_third*Sz1[0]*Sx1[0] );
    for( int j=1 ; j<5 ; j++ ) {
        tmp -= DSy[j-1] * tmp2;
        linindex += z_size;
        Jy [linindex] += tmp;
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int k=1 ; k<5 ; k++ ) {
        linindex_x += 1;
        linindex    = linindex_x;
        tmp  = cry_p * ( 0.5*Sx1[0]*Sz0[k] + one_third*DSz[k]*Sx1[0] );
        for( int j=1 ; j<5 ; j++ ) {
            vtmp[k] -= DSy[j-1] * tmp;
            linindex += z_size;
            Jy [linindex] += vtmp[k];
        }
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int i=1 ; i<5 ; i++ ) {
        linindex_x += yz_size;
        linindex    = linindex_x;
        tmp = cry_p * ( 0.5*Sz1[0]*Sx0[i] + one_third*Sz1[0]*DSx[i] );
        for( int j=1 ; j<5 ; j++ ) {
            vtmp[i] -= DSy[j-1] * tmp;
            linindex += z_size;
            Jy [linindex] += vtmp[i];
        }
    }
    linindex_x = linindex0;
    for( int i=1 ; i<5 ; i++ ) {
        linindex_x += yz_size;
        linindex_y  = linindex_x;
        for( int k=1 ; k<5 ; k++ ) {
            linindex_y += 1;
            linindex    = linindex_y;
            tmp = cry_p * ( Sz0[k]*Sx0[i] + 0.5*DSz[k]*Sx0[i] + 0.5*DSx[i]*Sz0[k] + one_third*DSz[k]*DSx[i] );
            for( int j=1 ; j<5 ; j++ ) {
                tmpJy[i][k] -= DSy[j-1] * tmp;
                linindex +=z_size;
                Jy [linindex] += tmpJy[i][k];
            }
        }
    }

    /* Jz^(p,p,d) */
    z_size =  nprimz+1;
    yz_size = ( nprimz+1 )*nprimy;
    linindex0 = ipo*yz_size+jpo*z_size+kpo;
    tmp = 0.;
    linindex = linindex0;
    tmp2 = crz_p * ( one_third*Sx1[0]*Sy1[0] );
    for( int k=1 ; k<5 ; k++ ) {
        tmp -= DSz[k-1] * tmp2;
        linindex += 1;
        Jz [linindex] += tmp;
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int j=1 ; j<5 ; j++ ) {
        linindex_x += z_size;
        linindex    = linindex_x;
        tmp = crz_p * ( 0.5*Sx1[0]*Sy0[j] + one_third*Sx1[0]*DSy[j] );
        for( int k=1 ; k<5 ; k++ ) {
            vtmp[j] -= DSz[k-1] * tmp;
            linindex += 1;
            Jz [linindex] += vtmp[j];
        }
    }
    for( unsigned int i=0 ; i<5 ; i++ ) vtmp[i] = 0.;
    linindex_x = linindex0;
    for( int i=1 ; i<5 ; i++ ) {
        linindex_x += yz_size;
        linindex    = linindex_x;
        tmp = crz_p * ( 0.5*Sy1[0]*Sx0[i] + one_third*DSx[i]*Sy1[0] );
        for( int k=1 ; k<5 ; k++ ) {
            vtmp[i] -= DSz[k-1] * tmp;
            linindex += 1;
            Jz [linindex] += vtmp[i];
        }
    }
    linindex_x = linindex0;
    for( int i=1 ; i<5 ; i++ ) {
        linindex_x += yz_size;
        linindex_y  = linindex_x;
        for( int j=1 ; j<5 ; j++ ) {
            linindex_y += z_size;
            linindex    = linindex_y;
            tmp = crz_p*( Sx0[i]*Sy0[j] + 0.5*DSx[i]*Sy0[j] + 0.5*DSy[j]*Sx0[i] + one_third*DSx[i]*DSy[j] );
            for( int k=1 ; k<5 ; k++ ) {
                tmpJz[i][j] -= DSz[k-1] * tmp;
                linindex += 1;
                Jz [linindex] += tmpJz[i][j];
            }
        }
    }
}

/* Projector3D4Order::currents, Projector/Projector3D4Order.cpp:49-233 */
static void currents_o4( double *Jx, double *Jy, double *Jz,
                         double xp, double yp, double zp, short charge, double weight,
                         const int *iold, const double *deltaold, int nparts,
                         double inv_cell_volume, const double *d_inv, const double *d_ov_dt,
                         const int *begin, int nprimy, int nprimz )
{
    double charge_weight = inv_cell_volume * ( double )( charge )*weight;
    double crx_p = charge_weight*d_ov_dt[0];
    double cry_p = charge_weight*d_ov_dt[1];
    double crz_p = charge_weight*d_ov_dt[2];

    double S0[3][7], S1[3][7], DS[3][7];
    double tmpJx[7][7], tmpJy[7][7], tmpJz[7][7];
    memset( S1, 0, sizeof( S1 ) );
    memset( tmpJx, 0, sizeof( tmpJx ) );
    memset( tmpJy, 0, sizeof( tmpJy ) );
    memset( tmpJz, 0, sizeof( tmpJz ) );

    const double pos[3] = { xp, yp, zp };
    int po[3];
    for( int c=0; c<3; c++ ) {
        double w[5];
        w4( deltaold[c*nparts], w );
        S0[c][0] = 0.;
        for( int i=0; i<5; i++ ) S0[c][i+1] = w[i];
        S0[c][6] = 0.;
    }
    for( int c=0; c<3; c++ ) {
        double pn = pos[c] * d_inv[c];
        int ip = ( int )round( pn );
        po[c] = iold[c*nparts];
        int ip_m_ipo = ip-po[c]-begin[c];
        double delta  = pn - ( double )ip;
        double w[5];
        w4( delta, w );
        for( int i=0; i<5; i++ ) S1[c][ip_m_ipo+1+i] = w[i];
    }
    for( unsigned int i=0; i < 7; i++ ) {
        DS[0][i] = S1[0][i] - S0[0][i];
        DS[1][i] = S1[1][i] - S0[1][i];
        DS[2][i] = S1[2][i] - S0[2][i];
    }
    const double *Sx0 = S0[0], *Sy0 = S0[1], *Sz0 = S0[2];
    const double *DSx = DS[0], *DSy = DS[1], *DSz = DS[2];
    int ipo = po[0]-3, jpo = po[1]-3, kpo = po[2]-3;
    int iloc, jloc, kloc, linindex;

    /* Jx^(d,p,p) */
    for( unsigned int i=1 ; i<7 ; i++ ) {
        iloc = i+ipo;
        for( unsigned int j=0 ; j<7 ; j++ ) {
            jloc = j+jpo;
            for( unsigned int k=0 ; k<7 ; k++ ) {
                tmpJx[j][k] -= crx_p * DSx[i-1] * ( Sy0[j]*Sz0[k] + 0.5*DSy[j]*Sz0[k] + 0.5*DSz[k]*Sy0[j] + one_third*DSy[j]*DSz[k] );
                kloc = k+kpo;
                linindex = iloc*nprimz*nprimy+jloc*nprimz+kloc;
                Jx [linindex] += tmpJx[j][k];
            }
        }
    }
    /* Jy^(p,d,p) */
    for( unsigned int i=0 ; i<7 ; i++ ) {
        iloc = i+ipo;
        for( unsigned int j=1 ; j<7 ; j++ ) {
            jloc = j+jpo;
            for( unsigned int k=0 ; k<7 ; k++ ) {
                tmpJy[i][k] -= cry_p * DSy[j-1] * ( Sz0[k]*Sx0[i] + 0.5*DSz[k]*Sx0[i] + 0.5*DSx[i]*Sz0[k] + one_third*DSz[k]*DSx[i] );
                kloc = k+kpo;
                linindex = iloc*nprimz*( nprimy+1 )+jloc*nprimz+kloc;
                Jy [linindex] += tmpJy[i][k];
            }
        }
    }
    /* Jz^(p,p,d) */
    for( unsigned int i=0 ; i<7 ; i++ ) {
        iloc = i+ipo;
        for( unsigned int j=0 ; j<7 ; j++ ) {
            jloc = j+jpo;
            for( unsigned int k=1 ; k<7 ; k++ ) {
                tmpJz[i][j] -= crz_p * DSz[k-1] * ( Sx0[i]*Sy0[j] + 0.5*DSx[i]*Sy0[j] + 0.5*DSy[j]*Sx0[i] + one_third*DSx[i]*DSy[j] );
                kloc = k+kpo;
                linindex = iloc*( nprimz+1 )*nprimy+jloc*( nprimz+1 )+kloc;
                Jz [linindex] += tmpJz[i][j];
            }
        }
    }
}

/* Projector3D{2,4}Order::currentsAndDensityWrapper (diag_flag = false, !is_spectral),
 * Projector3D2Order.cpp:733-752, Projector3D4Order.cpp:682-701 */
void orc_project( const orc_grid *g, int order, double *Jx, double *Jy, double *Jz,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold )
{
    /* Projector3D2Order::Projector3D2Order, Projector3D2Order.cpp:18-42; Projector.cpp:7 */
    double d_inv[3], d_ov_dt[3], mn[3], mx[3];
    int begin[3], p[3], d[3];
    const double cell_volume = 1.0 * g->cell[0] * g->cell[1] * g->cell[2];  /* Params.cpp:1172-1186 */
    const double inv_cell_volume = 1. / cell_volume;
    for( int i=0; i<3; i++ ) {
        d_inv[i]   = 1.0/g->cell[i];
        d_ov_dt[i] = g->cell[i] / g->dt;
    }
    orc_patch_bounds( g, mn, mx, begin );
    orc_dims( g, p, d );
    for( int ipart=istart ; ipart<iend; ipart++ ) {
        if( order==2 )
            currents_o2( Jx, Jy, Jz, x[ipart], y[ipart], z[ipart], q[ipart], w[ipart], &iold[ipart], &deltaold[ipart],
                         nparts, inv_cell_volume, d_inv, d_ov_dt, begin, p[1], p[2] );
        else
            currents_o4( Jx, Jy, Jz, x[ipart], y[ipart], z[ipart], q[ipart], w[ipart], &iold[ipart], &deltaold[ipart],
                         nparts, inv_cell_volume, d_inv, d_ov_dt, begin, p[1], p[2] );
    }
}

/* Projector3D2Order::currentsAndDensity, Projector/Projector3D2Order.cpp:349-521 */
void orc_project_rho_o2( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold_, const double *deltaold_ )
{
    double d_inv[3], d_ov_dt[3], mn[3], mx[3];
    int begin[3], p[3], d[3];
    const double cell_volume = 1.0 * g->cell[0] * g->cell[1] * g->cell[2];
    const double inv_cell_volume = 1. / cell_volume;
    for( int i=0; i<3; i++ ) {
        d_inv[i]   = 1.0/g->cell[i];
        d_ov_dt[i] = g->cell[i] / g->dt;
    }
    orc_patch_bounds( g, mn, mx, begin );
    orc_dims( g, p, d );
    const int nprimy = p[1], nprimz = p[2];

    for( int ipart=istart ; ipart<iend; ipart++ ) {
        const int *iold = &iold_[ipart];
        const double *deltaold = &deltaold_[ipart];
        double charge_weight = inv_cell_volume * ( double )( q[ipart] )*w[ipart];
        double crx_p = charge_weight*d_ov_dt[0];
        double cry_p = charge_weight*d_ov_dt[1];
        double crz_p = charge_weight*d_ov_dt[2];
        double S0[3][5], S1[3][5], DS[3][5];
        double tmpJx[5][5], tmpJy[5][5], tmpJz[5][5];
        memset( S1, 0, sizeof( S1 ) );
        memset( tmpJx, 0, sizeof( tmpJx ) );
        memset( tmpJy, 0, sizeof( tmpJy ) );
        memset( tmpJz, 0, sizeof( tmpJz ) );
        const double pos[3] = { x[ipart], y[ipart], z[ipart] };
        int po[3];
        for( int c=0; c<3; c++ ) {
            double delta = deltaold[c*nparts];
            double delta2 = delta*delta;
            S0[c][0] = 0.;
            S0[c][1] = 0.5 * ( delta2-delta+0.25 );
            S0[c][2] = 0.75-delta2;
            S0[c][3] = 0.5 * ( delta2+delta+0.25 );
            S0[c][4] = 0.;
        }
        for( int c=0; c<3; c++ ) {
            double pn = pos[c] * d_inv[c];
            int ip = ( int )round( pn );
            po[c] = iold[c*nparts];
            int ip_m_ipo = ip-po[c]-begin[c];
            double delta  = pn - ( double )ip;
            double delta2 = delta*delta;
            S1[c][ip_m_ipo+1] = 0.5 * ( delta2-delta+0.25 );
            S1[c][ip_m_ipo+2] = 0.75-delta2;
            S1[c][ip_m_ipo+3] = 0.5 * ( delta2+delta+0.25 );
        }
        for( unsigned int i=0; i < 5; i++ ) {
            DS[0][i] = S1[0][i] - S0[0][i];
            DS[1][i] = S1[1][i] - S0[1][i];
            DS[2][i] = S1[2][i] - S0[2][i];
        }
        const double *Sx0 = S0[0], *Sy0 = S0[1], *Sz0 = S0[2];
        const double *Sx1 = S1[0], *Sy1 = S1[1], *Sz1 = S1[2];
        const double *DSx = DS[0], *DSy = DS[1], *DSz = DS[2];
        int ipo = po[0]-2, jpo = po[1]-2, kpo = po[2]-2;
        int iloc, jloc, kloc, linindex;
        /* Projector3D2Order.cpp:467-506 */
        for( unsigned int i=1 ; i<5 ; i++ ) {
            iloc = i+ipo;
            for( unsigned int j=0 ; j<5 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<5 ; k++ ) {
                    tmpJx[j][k] -= crx_p * DSx[i-1] * ( Sy0[j]*Sz0[k] + 0.5*DSy[j]*Sz0[k] + 0.5*DSz[k]*Sy0[j] + one_third*DSy[j]*DSz[k] );
                    kloc = k+kpo;
                    linindex = iloc*nprimz*nprimy+jloc*nprimz+kloc;
                    Jx [linindex] += tmpJx[j][k];
                }
            }
        }
        for( unsigned int i=0 ; i<5 ; i++ ) {
            iloc = i+ipo;
            for( unsigned int j=1 ; j<5 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<5 ; k++ ) {
                    tmpJy[i][k] -= cry_p * DSy[j-1] * ( Sz0[k]*Sx0[i] + 0.5*DSz[k]*Sx0[i] + 0.5*DSx[i]*Sz0[k] + one_third*DSz[k]*DSx[i] );
                    kloc = k+kpo;
                    linindex = iloc*nprimz*( nprimy+1 )+jloc*nprimz+kloc;
                    Jy [linindex] += tmpJy[i][k];
                }
            }
        }
        for( unsigned int i=0 ; i<5 ; i++ ) {
            iloc = i+ipo;
            for( unsigned int j=0 ; j<5 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=1 ; k<5 ; k++ ) {
                    tmpJz[i][j] -= crz_p * DSz[k-1] * ( Sx0[i]*Sy0[j] + 0.5*DSx[i]*Sy0[j] + 0.5*DSy[j]*Sx0[i] + one_third*DSx[i]*DSy[j] );
                    kloc = k+kpo;
                    linindex = iloc*( nprimz+1 )*nprimy+jloc*( nprimz+1 )+kloc;
                    Jz [linindex] += tmpJz[i][j];
                }
            }
        }
        /* Projector3D2Order.cpp:509-519 */
        for( unsigned int i=0 ; i<5 ; i++ ) {
            iloc = i+ipo;
            for( unsigned int j=0 ; j<5 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<5 ; k++ ) {
                    kloc = k+kpo;
                    linindex = iloc*nprimz*nprimy+jloc*nprimz+kloc;
                    rho[linindex] += charge_weight * Sx1[i]*Sy1[j]*Sz1[k];
                }
            }
        }
    }
}

/* Projector3D4Order::currentsAndDensity, Projector/Projector3D4Order.cpp:239-432: the order-4 currents of
 * `currents` in the same loop form plus rho on the 7-point window (S1 only). */
void orc_project_rho_o4( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold_, const double *deltaold_ )
{
    double d_inv[3], d_ov_dt[3], mn[3], mx[3];
    int begin[3], p[3], d[3];
    const double cell_volume = 1.0 * g->cell[0] * g->cell[1] * g->cell[2];
    const double inv_cell_volume = 1. / cell_volume;
    for( int i=0; i<3; i++ ) {
        d_inv[i]   = 1.0/g->cell[i];
        d_ov_dt[i] = g->cell[i] / g->dt;
    }
    orc_patch_bounds( g, mn, mx, begin );
    orc_dims( g, p, d );
    const int nprimy = p[1], nprimz = p[2];

    for( int ipart=istart ; ipart<iend; ipart++ ) {
        const int *iold = &iold_[ipart];
        const double *deltaold = &deltaold_[ipart];
        double charge_weight = inv_cell_volume * ( double )( q[ipart] )*w[ipart];
        double crx_p = charge_weight*d_ov_dt[0];
        double cry_p = charge_weight*d_ov_dt[1];
        double crz_p = charge_weight*d_ov_dt[2];
        double S0[3][7], S1[3][7], DS[3][7];
        double tmpJx[7][7], tmpJy[7][7], tmpJz[7][7];
        memset( S1, 0, sizeof( S1 ) );
        memset( tmpJx, 0, sizeof( tmpJx ) );
        memset( tmpJy, 0, sizeof( tmpJy ) );
        memset( tmpJz, 0, sizeof( tmpJz ) );
        const double pos[3] = { x[ipart], y[ipart], z[ipart] };
        int po[3];
        for( int c=0; c<3; c++ ) {                 /* :284-318 */
            S0[c][0] = 0.;
            w4( deltaold[c*nparts], &S0[c][1] );
            S0[c][6] = 0.;
        }
        for( int c=0; c<3; c++ ) {                 /* :321-361 */
            double pn = pos[c] * d_inv[c];
            int ip = ( int )round( pn );
            po[c] = iold[c*nparts];
            int ip_m_ipo = ip-po[c]-begin[c];
            w4( pn - ( double )ip, &S1[c][ip_m_ipo+1] );
        }
        for( unsigned int i=0; i < 7; i++ ) {
            DS[0][i] = S1[0][i] - S0[0][i];
            DS[1][i] = S1[1][i] - S0[1][i];
            DS[2][i] = S1[2][i] - S0[2][i];
        }
        const double *Sx0 = S0[0], *Sy0 = S0[1], *Sz0 = S0[2];
        const double *Sx1 = S1[0], *Sy1 = S1[1], *Sz1 = S1[2];
        const double *DSx = DS[0], *DSy = DS[1], *DSz = DS[2];
        int ipo = po[0]-3, jpo = po[1]-3, kpo = po[2]-3;       /* :373-377 */
        int iloc, jloc, kloc, linindex;
        for( unsigned int i=1 ; i<7 ; i++ ) {                  /* :382-393 */
            iloc = i+ipo;
            for( unsigned int j=0 ; j<7 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<7 ; k++ ) {
                    tmpJx[j][k] -= crx_p * DSx[i-1] * ( Sy0[j]*Sz0[k] + 0.5*DSy[j]*Sz0[k] + 0.5*DSz[k]*Sy0[j] + one_third*DSy[j]*DSz[k] );
                    kloc = k+kpo;
                    linindex = iloc*nprimz*nprimy+jloc*nprimz+kloc;
                    Jx [linindex] += tmpJx[j][k];
                }
            }
        }
        for( unsigned int i=0 ; i<7 ; i++ ) {                  /* :396-407 */
            iloc = i+ipo;
            for( unsigned int j=1 ; j<7 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<7 ; k++ ) {
                    tmpJy[i][k] -= cry_p * DSy[j-1] * ( Sz0[k]*Sx0[i] + 0.5*DSz[k]*Sx0[i] + 0.5*DSx[i]*Sz0[k] + one_third*DSz[k]*DSx[i] );
                    kloc = k+kpo;
                    linindex = iloc*nprimz*( nprimy+1 )+jloc*nprimz+kloc;
                    Jy [linindex] += tmpJy[i][k];
                }
            }
        }
        for( unsigned int i=0 ; i<7 ; i++ ) {                  /* :410-421 */
            iloc = i+ipo;
            for( unsigned int j=0 ; j<7 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=1 ; k<7 ; k++ ) {
                    tmpJz[i][j] -= crz_p * DSz[k-1] * ( Sx0[i]*Sy0[j] + 0.5*DSx[i]*Sy0[j] + 0.5*DSy[j]*Sx0[i] + one_third*DSx[i]*DSy[j] );
                    kloc = k+kpo;
                    linindex = iloc*( nprimz+1 )*nprimy+jloc*( nprimz+1 )+kloc;
                    Jz [linindex] += tmpJz[i][j];
                }
            }
        }
        for( unsigned int i=0 ; i<7 ; i++ ) {                  /* :424-434 */
            iloc = i+ipo;
            for( unsigned int j=0 ; j<7 ; j++ ) {
                jloc = j+jpo;
                for( unsigned int k=0 ; k<7 ; k++ ) {
                    kloc = k+kpo;
                    linindex = iloc*nprimz*nprimy+jloc*nprimz+kloc;
                    rho[linindex] += charge_weight * Sx1[i]*Sy1[j]*Sz1[k];
                }
            }
        }
    }
}

/* Projector3D{2,4}Order::currentsAndDensityWrapper with diag_flag = true (Projector3D2Order.cpp:753-763,
 * Projector3D4Order.cpp:703-713): the arrays handed in are either the totals or the species' own Jx_s .. rho_s. */
void orc_project_rho( const orc_grid *g, int order, double *Jx, double *Jy, double *Jz, double *rho,
                  const double *x, const double *y, const double *z,
                  const short *q, const double *w, int nparts, int istart, int iend,
                  const int *iold, const double *deltaold )
{
    if( order == 2 ) orc_project_rho_o2( g, Jx, Jy, Jz, rho, x, y, z, q, w, nparts, istart, iend, iold, deltaold );
    else orc_project_rho_o4( g, Jx, Jy, Jz, rho, x, y, z, q, w, nparts, istart, iend, iold, deltaold );
}

/* ElectroMagn3D::computeTotalRhoJ, ElectroMagn/ElectroMagn3D.cpp:1753-1799, one species: total += species array */
void orc_compute_total_rhoJ( const orc_grid *g, double *Jx, double *Jy, double *Jz, double *rho,
                             const double *Jx_s, const double *Jy_s, const double *Jz_s, const double *rho_s )
{
    double *tot[4] = { Jx, Jy, Jz, rho };
    const double *sp[4] = { Jx_s, Jy_s, Jz_s, rho_s };
    const int id[4] = { 0, 1, 2, 6 };       /* orc_field_size ids: Jx Jy Jz as Ex Ey Ez, rho all primal */
    for( int a=0; a<4; a++ ) {
        if( !sp[a] ) continue;
        const size_t n = orc_field_size( g, id[a] );
        for( size_t i=0; i<n; i++ ) tot[a][i] += sp[a][i];
    }
}

/* ------------------------------------------------------------------------- */
/* a16-a19 Maxwell                                                            */
/* ------------------------------------------------------------------------- */

/* ElectroMagn3D::saveMagneticFields, ElectroMagn/ElectroMagn3D.cpp:1048-1118 */
void orc_save_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                 double *Bxm, double *Bym, double *Bzm )
{
    memcpy( Bxm, Bx, orc_field_size( g, 3 )*sizeof( double ) );
    memcpy( Bym, By, orc_field_size( g, 4 )*sizeof( double ) );
    memcpy( Bzm, Bz, orc_field_size( g, 5 )*sizeof( double ) );
}

/* MA_Solver3D_norm::operator(), ElectroMagnSolver/MA_Solver3D_norm.cpp:18-115 */
void orc_maxwell_ampere( const orc_grid *g, double *Ex3D, double *Ey3D, double *Ez3D,
                         const double *Bx3D, const double *By3D, const double *Bz3D,
                         const double *Jx3D, const double *Jy3D, const double *Jz3D )
{
    int p[3], d[3];
    orc_dims( g, p, d );
    const unsigned int nx_p = p[0], nx_d = d[0], ny_p = p[1], ny_d = d[1], nz_p = p[2], nz_d = d[2];
    /* Solver3D::Solver3D, ElectroMagnSolver/Solver3D.h:14-24 */
    const double dt = g->dt;
    const double dt_ov_dx = g->dt / g->cell[0];
    const double dt_ov_dy = g->dt / g->cell[1];
    const double dt_ov_dz = g->dt / g->cell[2];

    for( unsigned int i=0 ; i<nx_d ; i++ ) {
        for( unsigned int j=0 ; j<ny_p ; j++ ) {
            for( unsigned int k=0 ; k<nz_p ; k++ ) {
                Ex3D[ i*( ny_p*nz_p ) + j*( nz_p ) + k ] += -dt*Jx3D[ i*( ny_p*nz_p ) + j*( nz_p ) + k ]
                        +                 dt_ov_dy * ( Bz3D[ i*( ny_d*nz_p ) + ( j+1 )*( nz_p ) + k   ] - Bz3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] )
                        -                 dt_ov_dz * ( By3D[ i*( ny_p*nz_d ) +  j   *( nz_d ) + k+1 ] - By3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] );
            }
        }
    }
    for( unsigned int i=0 ; i<nx_p ; i++ ) {
        for( unsigned int j=0 ; j<ny_d ; j++ ) {
            for( unsigned int k=0 ; k<nz_p ; k++ ) {
                Ey3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] += -dt*Jy3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ]
                        -                  dt_ov_dx * ( Bz3D[ ( i+1 )*( ny_d*nz_p ) + j*( nz_p ) + k   ] - Bz3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] )
                        +                  dt_ov_dz * ( Bx3D[  i   *( ny_d*nz_d ) + j*( nz_d ) + k+1 ] - Bx3D[ i*( ny_d*nz_d ) + j*( nz_d ) + k ] );
            }
        }
    }
    for( unsigned int i=0 ;  i<nx_p ; i++ ) {
        for( unsigned int j=0 ; j<ny_p ; j++ ) {
            for( unsigned int k=0 ; k<nz_d ; k++ ) {
                Ez3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] += -dt*Jz3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ]
                        +                  dt_ov_dx * ( By3D[ ( i+1 )*( ny_p*nz_d ) +  j   *( nz_d ) + k ] - By3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] )
                        -                  dt_ov_dy * ( Bx3D[  i   *( ny_d*nz_d ) + ( j+1 )*( nz_d ) + k ] - Bx3D[ i*( ny_d*nz_d ) + j*( nz_d ) + k ] );
            }
        }
    }
}

/* MF_Solver3D_Yee::operator(), ElectroMagnSolver/MF_Solver3D_Yee.cpp:18-111 */
void orc_maxwell_faraday( const orc_grid *g, const double *Ex3D, const double *Ey3D, const double *Ez3D,
                          double *Bx3D, double *By3D, double *Bz3D )
{
    int p[3], d[3];
    orc_dims( g, p, d );
    const unsigned int nx_p = p[0], nx_d = d[0], ny_p = p[1], ny_d = d[1], nz_p = p[2], nz_d = d[2];
    const double dt_ov_dx = g->dt / g->cell[0];
    const double dt_ov_dy = g->dt / g->cell[1];
    const double dt_ov_dz = g->dt / g->cell[2];

    for( unsigned int i=0 ; i<nx_p;  i++ ) {
        for( unsigned int j=1 ; j<ny_d-1 ; j++ ) {
            for( unsigned int k=1 ; k<nz_d-1 ; k++ ) {
                Bx3D[ i*( ny_d*nz_d ) + j*( nz_d ) + k ] += -dt_ov_dy * ( Ez3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] - Ez3D[ i*( ny_p*nz_d ) + ( j-1 )*( nz_d ) + k   ] )
                        +   dt_ov_dz * ( Ey3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] - Ey3D[ i*( ny_d*nz_p ) +  j   *( nz_p ) + k-1 ] );
            }
        }
    }
    for( unsigned int i=1 ; i<nx_d-1 ; i++ ) {
        for( unsigned int j=0 ; j<ny_p ; j++ ) {
            for( unsigned int k=1 ; k<nz_d-1 ; k++ ) {
                By3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] += -dt_ov_dz * ( Ex3D[ i*( ny_p*nz_p ) + j*( nz_p ) + k ] - Ex3D[  i   *( ny_p*nz_p ) + j*( nz_p ) + k-1 ] )
                        +   dt_ov_dx * ( Ez3D[ i*( ny_p*nz_d ) + j*( nz_d ) + k ] - Ez3D[ ( i-1 )*( ny_p*nz_d ) + j*( nz_d ) + k   ] );
            }
        }
    }
    for( unsigned int i=1 ; i<nx_d-1 ; i++ ) {
        for( unsigned int j=1 ; j<ny_d-1 ; j++ ) {
            for( unsigned int k=0 ; k<nz_p ; k++ ) {
                Bz3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] += -dt_ov_dx * ( Ey3D[ i*( ny_d*nz_p ) + j*( nz_p ) + k ] - Ey3D[ ( i-1 )*( ny_d*nz_p ) +  j   *( nz_p ) + k ] )
                        +   dt_ov_dy * ( Ex3D[ i*( ny_p*nz_p ) + j*( nz_p ) + k ] - Ex3D[  i   *( ny_p*nz_p ) + ( j-1 )*( nz_p ) + k ] );
            }
        }
    }
}

/* ElectroMagn3D::centerMagneticFields, ElectroMagn/ElectroMagn3D.cpp:1191-1293 */
void orc_center_B( const orc_grid *g, const double *Bx, const double *By, const double *Bz,
                   double *Bxm, double *Bym, double *Bzm )
{
    long n;
    n = orc_field_size( g, 3 );
    for( long i=0; i<n; i++ ) Bxm[i] = ( Bx[i] + Bxm[i] )*0.5;
    n = orc_field_size( g, 4 );
    for( long i=0; i<n; i++ ) Bym[i] = ( By[i] + Bym[i] )*0.5;
    n = orc_field_size( g, 5 );
    for( long i=0; i<n; i++ ) Bzm[i] = ( Bz[i] + Bzm[i] )*0.5;
}

/* ------------------------------------------------------------------------- */
/* a20-a21 keys + sort                                                        */
/* ------------------------------------------------------------------------- */

/* SpeciesV::computeParticleCellKeys, Species/SpeciesV.cpp:766-855 (nDim_field == 3).
 * length_[i] = patch_size_[i]+1 (SpeciesV.cpp:72-77); the key arithmetic is done in
 * int exactly as the reference does (double -> int conversion on assignment). */
void orc_cell_keys( const orc_grid *g, const double *x, const double *y, const double *z,
                    int *cell_keys, int *count, int istart, int iend )
{
    double mn[3], mx[3], dx_inv[3];
    int begin[3];
    orc_patch_bounds( g, mn, mx, begin );
    for( int i=0; i<3; i++ ) dx_inv[i] = 1./g->cell[i];          /* Species.cpp: dx_inv_ = 1/cell_length */
    const unsigned int length1 = g->n[1]+1, length2 = g->n[2]+1;
    double min_loc_x = round( mn[0] * dx_inv[0] );
    double min_loc_y = round( mn[1] * dx_inv[1] );
    double min_loc_z = round( mn[2] * dx_inv[2] );
    for( int iPart=istart; iPart < iend ; iPart++ ) {
        if( cell_keys[iPart] >= 0 ) {
            cell_keys[iPart]  = round( x[iPart] * dx_inv[0] )- min_loc_x ;
            cell_keys[iPart] *= length1;
            cell_keys[iPart] += round( y[iPart] * dx_inv[1] )- min_loc_y ;
            cell_keys[iPart] *= length2;
            cell_keys[iPart] += round( z[iPart] * dx_inv[2] )- min_loc_z ;
        }
    }
    if( count ) {
        for( int iPart=istart; iPart < iend ; iPart++ ) {
            if( cell_keys[iPart] >= 0 ) {
                count[cell_keys[iPart]] ++;
            }
        }
    }
}

/* Canonical order of this build: stable counting sort on the key above; particles
 * with key<0 (leavers) are dropped.  first[] is the prefix sum the reference builds in
 * SpeciesV::sortParticles (Species/SpeciesV.cpp:645-652: first_index / last_index). */
int orc_counting_sort_perm( const int *keys, int nparts, int ncells, int *first, int *perm )
{
    int *cursor = ( int * )calloc( ( size_t )ncells+1, sizeof( int ) );
    for( int i=0; i<nparts; i++ ) if( keys[i]>=0 ) cursor[keys[i]+1]++;
    first[0] = 0;
    for( int c=0; c<ncells; c++ ) first[c+1] = first[c] + cursor[c+1];
    for( int c=0; c<ncells; c++ ) cursor[c] = first[c];
    for( int i=0; i<nparts; i++ ) if( keys[i]>=0 ) perm[cursor[keys[i]]++] = i;
    int kept = first[ncells];
    free( cursor );
    return kept;
}

/* ------------------------------------------------------------------------- */
/* a23 energies                                                               */
/* ------------------------------------------------------------------------- */

/* DiagnosticScalar::compute, Diagnostic/DiagnosticScalar.cpp:497-510 (CPU mode, mass>0) */
double orc_ukin( double mass, const double *px, const double *py, const double *pz, const double *w, int nPart )
{
    double ener_tot = 0.0;
    for( int iPart=0 ; iPart<nPart; iPart++ ) {
        const double gamma = sqrt( 1 + px[iPart]*px[iPart] + py[iPart]*py[iPart] + pz[iPart]*pz[iPart] );
        ener_tot += w[iPart] * ( gamma - 1.0 );
    }
    ener_tot *= mass;
    return ener_tot;
}

/* Field3D::norm2, Field/Field3D.cpp:230-250 with istart/bufsize from
 * ElectroMagn3D::initElectroMagn3DQuantities, ElectroMagn/ElectroMagn3D.cpp:190-229 */
double orc_field_norm2( const orc_grid *g, const double *f, int dualx, int dualy, int dualz )
{
    int dims[3];
    const int isDual[3] = { dualx, dualy, dualz };
    comp_dims( g, dualx, dualy, dualz, dims );
    int s[3], e[3];
    for( int i=0; i<3; i++ ) {
        int istart = g->o[i];
        if( g->pcoord[i]!=0 ) istart += 1;
        int bufsize = g->n[i] + 1 + isDual[i];
        if( g->npatch[i]!=1 ) {
            if( ( !isDual[i] ) && ( g->pcoord[i]!=0 ) ) {
                bufsize--;
            } else if( isDual[i] ) {
                bufsize--;
                if( ( g->pcoord[i]!=0 ) && ( g->pcoord[i]!=g->npatch[i]-1 ) ) bufsize--;
            }
        }
        s[i] = istart;
        e[i] = istart+bufsize;
    }
    double nrj = 0.;
    for( int i=s[0] ; i<e[0] ; i++ )
        for( int j=s[1] ; j<e[1] ; j++ )
            for( int k=s[2] ; k<e[2] ; k++ ) {
                double v = f[ ( ( long )i*dims[1]+j )*dims[2]+k ];
                nrj += v*v;
            }
    return nrj;
}

/* DiagnosticScalar::compute, Diagnostic/DiagnosticScalar.cpp:658-691: fields Ex,Ey,Ez,Bx_m,By_m,Bz_m */
double orc_uelm( const orc_grid *g, const double *Ex, const double *Ey, const double *Ez,
                 const double *Bxm, const double *Bym, const double *Bzm )
{
    const double cell_volume = 1.0 * g->cell[0] * g->cell[1] * g->cell[2];
    const double *f[6] = { Ex, Ey, Ez, Bxm, Bym, Bzm };
    static const int dual[6][3] = { {1,0,0},{0,1,0},{0,0,1}, {0,1,1},{1,0,1},{1,1,0} };
    double Uelm = 0.;
    for( int ifield=0; ifield<6; ifield++ ) {
        double Uem = orc_field_norm2( g, f[ifield], dual[ifield][0], dual[ifield][1], dual[ifield][2] );
        Uem *= 0.5*cell_volume;
        Uelm += Uem;
    }
    return Uelm;
}

/* ------------------------------------------------------------------------- */
/* halo semantics                                                             */
/* ------------------------------------------------------------------------- */

/* SyncVectorPatch::sumAllComponents, local-neighbour branch,
 * Patch/SyncVectorPatch.cpp:263-311 (x), :395-440 (y), z alike:
 * planes [n, n+gsp) of L and [0, gsp) of R along `dim` are summed, both keep the sum;
 * gsp = 1+2*oversize+isDual[dim]; full extent in the other two dims. */
void orc_sum_pair( const orc_grid *g, int dim, int dualx, int dualy, int dualz, double *L, double *R )
{
    int dims[3];
    const int isDual[3] = { dualx, dualy, dualz };
    comp_dims( g, dualx, dualy, dualz, dims );
    const int gsp = 1+2*g->o[dim]+isDual[dim];
    const long stride[3] = { ( long )dims[1]*dims[2], dims[2], 1 };
    int lo[3] = {0,0,0}, hi[3] = { dims[0], dims[1], dims[2] };
    hi[dim] = gsp;
    const long shift = ( long )g->n[dim]*stride[dim];
    for( int i=lo[0]; i<hi[0]; i++ )
        for( int j=lo[1]; j<hi[1]; j++ )
            for( int k=lo[2]; k<hi[2]; k++ ) {
                long idx = i*stride[0]+j*stride[1]+k;
                L[idx+shift] += R[idx];
                R[idx] = L[idx+shift];
            }
}

/* SyncVectorPatch::exchangeAllComponentsAlong{X,Y,Z}, local-neighbour branch,
 * Patch/SyncVectorPatch.cpp:1483-1527 (x), :1630-1660 (y):
 * R[0,o) <- L[n, n+o) ;  L[n+gsp, n+gsp+o) <- R[gsp, gsp+o),  gsp = o+1+isDual[dim]. */
void orc_exchange_pair( const orc_grid *g, int dim, int dualx, int dualy, int dualz, double *L, double *R )
{
    int dims[3];
    const int isDual[3] = { dualx, dualy, dualz };
    comp_dims( g, dualx, dualy, dualz, dims );
    const int o = g->o[dim];
    const int gsp = o+1+isDual[dim];
    const long stride[3] = { ( long )dims[1]*dims[2], dims[2], 1 };
    int hi[3] = { dims[0], dims[1], dims[2] };
    hi[dim] = o;
    const long shift = ( long )g->n[dim]*stride[dim];
    const long gshift = ( long )gsp*stride[dim];
    for( int i=0; i<hi[0]; i++ )
        for( int j=0; j<hi[1]; j++ )
            for( int k=0; k<hi[2]; k++ ) {
                long idx = i*stride[0]+j*stride[1]+k;
                R[idx] = L[idx+shift];
                L[idx+shift+gshift] = R[idx+gshift];
            }
}

/* ------------------------------------------------------------------------- */
/* PartBoundCond::apply with `remove` conditions at global box sides
 * (ParticleBC/PartBoundCond.h:38-76; remove_particle_inf/sup, ParticleBC/BoundaryConditionType.cpp:204-294;
 * internal_inf/sup, :15-57).  bc_remove[2*d+s] != 0: side s of dimension d is a global box side with the
 * `remove` condition (the limits are then max/min of the global and the patch bounds, PartBoundCond.cpp:44-69,
 * i.e. the patch bounds).  energy_lost receives the reference's energy_tot (sum of w*(gamma-1)).            */
void orc_bc_apply( const orc_grid *g, const int *bc_remove, const double *x, const double *y, const double *z,
                   const double *px, const double *py, const double *pz, const double *w, short *q,
                   int *cell_keys, int imin, int imax, double *energy_lost )
{
    double mn[3], mx[3];
    int begin[3];
    orc_patch_bounds( g, mn, mx, begin );
    const double *position[3] = { x, y, z };
    double energy_tot = 0.;
    for( int ipart=imin; ipart<imax; ipart++ ) cell_keys[ipart] = 0;
    for( int direction=0; direction<3; direction++ ) {
        for( int side=0; side<2; side++ ) {
            double change_in_energy = 0.0;
            const double limit = side==0 ? mn[direction] : mx[direction];
            for( int ipart=imin ; ipart<imax ; ipart++ ) {
                const int beyond = side==0 ? ( position[direction][ipart] < limit ) : ( position[direction][ipart] >= limit );
                if( bc_remove[2*direction+side] ) {
                    if( beyond ) {
                        const double LorentzFactor = sqrt( 1.0 + px[ipart]*px[ipart] + py[ipart]*py[ipart] + pz[ipart]*pz[ipart] );
                        change_in_energy += w[ipart] * ( LorentzFactor - 1.0 );
                        q[ipart] = 0;
                        cell_keys[ipart] = -1;
                    }
                } else if( cell_keys[ipart] >= 0 && beyond ) {
                    cell_keys[ipart] = ( side==0 ? -2 : -3 ) - 2 * direction;
                }
            }
            energy_tot += change_in_energy;
        }
    }
    *energy_lost = energy_tot;
}

/* ------------------------------------------------------------------------- */
/* ElectroMagnBC3D_SM (ElectroMagnBC/ElectroMagnBC3D_SM.cpp): constructor coefficients (:61-75) and apply()
 * (:141-376) with zero external fields (B_val = 0).  Fields in the reference's compact layout; db1 / db2 are
 * the laser amplitude arrays b1 (n1p x n2d) and b2 (n1d x n2p) the reference fills from Laser::getAmplitude0/1,
 * or NULL.  is_boundary = { isBoundary1min, isBoundary1max, isBoundary2min, isBoundary2max }.              */
void orc_apply_SM( const orc_grid *g, int i_boundary, const double *K, const int *is_boundary,
                   const double *Ex, const double *Ey, const double *Ez, double *Bx, double *By, double *Bz,
                   const double *db1, const double *db2 )
{
    int n_p[3], n_d[3];
    orc_dims( g, n_p, n_d );
    const int axis0_ = i_boundary / 2;
    const int axis1_ = axis0_ == 0 ? 1 : 0;
    const int axis2_ = axis0_ == 2 ? 1 : 2;
    const double sign_ = ( double )( i_boundary % 2 ) *2 - 1.;
    /* acts only on a patch at that global side (patch->isBoundary( i_boundary_ ), :143) */
    if( sign_ < 0 ? g->pcoord[axis0_] != 0 : g->pcoord[axis0_] != g->npatch[axis0_]-1 ) return;
    int iB_[3] = { 0, 0, 0 };
    if( sign_ > 0 ) {
        iB_[axis0_] = n_p[axis0_] - 1;
        iB_[axis1_] = n_d[axis0_] - 1;
        iB_[axis2_] = n_d[axis0_] - 1;
    }
    double dt_ov_d[3];
    for( int i=0; i<3; i++ ) dt_ov_d[i] = g->dt / g->cell[i];
    const double Knorm = sqrt( K[0]*K[0] + K[1]*K[1] + K[2]*K[2] ) ;
    const double omega = 1.;
    const double k0 = omega*K[axis0_] / Knorm;
    const double k1 = omega*K[axis1_] / Knorm;
    const double k2 = omega*K[axis2_] / Knorm;
    const double factor = 1.0 / ( k0 - sign_ * dt_ov_d[axis0_] );
    const double Alpha_   = 2.0 * factor;
    const double Beta_    = - ( k0 + sign_ * dt_ov_d[axis0_] ) * factor;
    const double Gamma_   = 4.0 * k0 * factor;
    const double Delta_   = - ( k1 + dt_ov_d[axis1_] ) * factor;
    const double Epsilon_ = - ( k1 - dt_ov_d[axis1_] ) * factor;
    const double Zeta_    = - ( k2 + dt_ov_d[axis2_] ) * factor;
    const double Eta_     = - ( k2 - dt_ov_d[axis2_] ) * factor;

    const double *E[3] = { Ex, Ey, Ez };
    double *B[3] = { Bx, By, Bz };
    const double *E1 = E[axis1_], *E2 = E[axis2_], *B0 = B[axis0_];
    double *B1 = B[axis1_], *B2 = B[axis2_];
    const unsigned int nz_p = n_p[2], nz_d = n_d[2];
    const unsigned int nyz_pp = n_p[1]*n_p[2], nyz_pd = n_p[1]*n_d[2], nyz_dp = n_d[1]*n_p[2], nyz_dd = n_d[1]*n_d[2];
    const unsigned int n1p = n_p[axis1_], n1d = n_d[axis1_], n2p = n_p[axis2_], n2d = n_d[axis2_];
    const unsigned int p0 = iB_[axis0_];
    const unsigned int p1 = iB_[axis1_] - sign_;
    const unsigned int iB1 = iB_[axis1_];
    const unsigned int b1min = is_boundary[0], b1max = is_boundary[1], b2min = is_boundary[2], b2max = is_boundary[3];
#define DB1( j, k ) ( db1 ? db1[( j )*n2d + ( k )] : 0. )
#define DB2( j, k ) ( db2 ? db2[( j )*n2p + ( k )] : 0. )
    /* B1 */
    if( axis0_ == 0 ) {
        for( unsigned int j=b1min; j<n1p-b1max ; j++ )
            for( unsigned int k=b2min ; k<n2d-b2max ; k++ )
                B1[ iB1*nyz_pd + j*nz_d + k ]
                    = Alpha_   *  E2[ p0*nyz_pd + j*nz_d + k ]
                    + Beta_    *( B1[ p1*nyz_pd + j*nz_d + k ]-0. )
                    + Gamma_   * DB1( j, k )
                    + Delta_   *( B0[ p0*nyz_dd + (j+1)*nz_d + k ]-0. )
                    + Epsilon_ *( B0[ p0*nyz_dd +  j   *nz_d + k ]-0. )
                    + 0.;
    } else if( axis0_ == 1 ) {
        for( unsigned int i=b1min; i<n1p-b1max ; i++ )
            for( unsigned int k=b2min ; k<n2d-b2max ; k++ )
                B1[ i*nyz_dd + iB1*nz_d + k ]
                    =-Alpha_   *  E2[ i*nyz_pd + p0*nz_d + k ]
                    + Beta_    *( B1[ i*nyz_dd + p1*nz_d + k ]-0. )
                    + Gamma_   * DB1( i, k )
                    + Delta_   *( B0[ (i+1)*nyz_pd + p0*nz_d + k ]-0. )
                    + Epsilon_ *( B0[  i   *nyz_pd + p0*nz_d + k ]-0. )
                    + 0.;
    } else {
        for( unsigned int i=b1min; i<n1p-b1max ; i++ )
            for( unsigned int j=b2min ; j<n2d-b2max ; j++ )
                B1[ i*nyz_dd + j*nz_d + iB1 ]
                    = Alpha_   *  E2[ i*nyz_dp + j*nz_p + p0 ]
                    + Beta_    *( B1[ i*nyz_dd + j*nz_d + p1 ]-0. )
                    + Gamma_   * DB1( i, j )
                    + Delta_   *( B0[ (i+1)*nyz_dp + j*nz_p + p0 ]-0. )
                    + Epsilon_ *( B0[  i   *nyz_dp + j*nz_p + p0 ]-0. )
                    + 0.;
    }
    /* B2 */
    if( axis0_ == 0 ) {
        for( unsigned int j=b1min; j<n1d-b1max ; j++ )
            for( unsigned int k=b2min; k<n2p-b2max ; k++ )
                B2[ iB1*nyz_dp + j*nz_p + k ]
                    = -Alpha_ *  E1[ p0*nyz_dp + j*nz_p + k ]
                    +  Beta_  *( B2[ p1*nyz_dp + j*nz_p + k ]-0. )
                    +  Gamma_ * DB2( j, k )
                    +  Zeta_  *( B0[ p0*nyz_dd + j*nz_d + k+1 ]-0. )
                    +  Eta_   *( B0[ p0*nyz_dd + j*nz_d + k   ]-0. )
                    +  0.;
    } else if( axis0_ == 1 ) {
        for( unsigned int i=b1min; i<n1d-b1max ; i++ )
            for( unsigned int k=b2min; k<n2p-b2max ; k++ )
                B2[ i*nyz_dp + iB1*nz_p + k ]
                    =  Alpha_ *  E1[ i*nyz_pp + p0*nz_p + k ]
                    +  Beta_  *( B2[ i*nyz_dp + p1*nz_p + k ]-0. )
                    +  Gamma_ * DB2( i, k )
                    +  Zeta_  *( B0[ i*nyz_pd + p0*nz_d + k+1 ]-0. )
                    +  Eta_   *( B0[ i*nyz_pd + p0*nz_d + k   ]-0. )
                    +  0.;
    } else {
        for( unsigned int i=b1min; i<n1d-b1max ; i++ )
            for( unsigned int j=b2min; j<n2p-b2max ; j++ )
                B2[ i*nyz_pd + j*nz_d + iB1 ]
                    = -Alpha_ *  E1[ i*nyz_pp + j*nz_p + p0 ]
                    +  Beta_  *( B2[ i*nyz_pd + j*nz_d  + p1 ]-0. )
                    +  Gamma_ * DB2( i, j )
                    +  Zeta_  *( B0[ i*nyz_dp + (j+1)*nz_p + p0 ]-0. )
                    +  Eta_   *( B0[ i*nyz_dp +  j   *nz_p + p0 ]-0. )
                    +  0.;
    }
#undef DB1
#undef DB2
}
