/*
 * edge_physics.h -- per-edge arithmetic of the nalu-wind edge kernels, written
 * for registers: every function takes the two end-node states by value-struct
 * and returns the compact per-edge result the row reduction needs.
 *
 * Operation order follows the reference lambdas so that results agree with the
 * reference to rounding (tolerance 1e-12 relative, see tests/):
 *   continuity_edge / mdot_edge : src/edge_kernels/ContinuityEdgeSolverAlg.C:109-194,
 *                                 src/ngp_algorithms/MdotEdgeAlg.C:117-190
 *   scalar_edge                 : src/edge_kernels/ScalarEdgeSolverAlg.C:85-205
 *   momentum_edge               : src/edge_kernels/MomentumEdgeSolverAlg.C:105-312
 *   peclet_edge                 : src/edge_kernels/MomentumEdgePecletAlg.C:74-101
 *   peclet_eval                 : src/PecletFunction.C:41-45, 68-71
 *   van_leer                    : include/edge_kernels/EdgeKernelUtils.h:18-24
 *
 * NW_HD is __host__ __device__ under nvcc and empty under g++ (the plan
 * emulator used by the CPU test-suite compiles this header with g++).
 */
#ifndef NW_EDGE_PHYSICS_H
#define NW_EDGE_PHYSICS_H

#include <math.h>

#include "nw_types.h"

#if defined(__CUDACC__)
#define NW_HD __host__ __device__ __forceinline__
#else
#define NW_HD inline
#endif

namespace nw {

/* Reciprocal for the divisions of the edge kernels.  The FP64 pipe is the
 * second roofline of this path, and nvcc's IEEE divide
 * costs ~12 FP64 instructions plus a branchy slow path that zero numerators
 * (limiter on a flat field) fall into.  On the device: MUFU.RCP64H seed
 * (rel. error <= 2^-23) + two Newton steps = 1 MUFU + 4 DFMA, <= 2 ulp for
 * normal-range arguments -- every denominator here (a.dx, rho, momentum_diag,
 * dq^2 + 1e-16, D + 1e-16, relaxation factors) is normal-range for valid
 * input; 0 / denormal / inf arguments give inf / NaN exactly where the
 * reference's divide gives inf / NaN downstream.  The 1e-12 parity bar is
 * three orders of magnitude above this.  Host (CPU walk-through): 1.0 / x. */
NW_HD double
nw_rcp(double x)
{
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

/* exactly-rounded, never-contracted fp64 ops for the Peclet chain (see
 * peclet_eval); g++ on baseline x86-64 does not contract either */
#if defined(__CUDA_ARCH__)
#define NW_XMUL(a, b) __dmul_rn((a), (b))
#define NW_XADD(a, b) __dadd_rn((a), (b))
#define NW_XDIV(a, b) __ddiv_rn((a), (b))
#else
#define NW_XMUL(a, b) ((a) * (b))
#define NW_XADD(a, b) ((a) + (b))
#define NW_XDIV(a, b) ((a) / (b))
#endif

/* LowMach::udiag_post_processing, one node (src/LowMachEquationSystem.C:
 * 2783-2790): the extracted momentum diagonal per unit mass with the velocity
 * relaxation taken out again; every operation rounded on its own */
NW_HD double
udiag_post_value(
  double udiag, double rho, double dualVol, double projTimeScale, double alphaU)
{
  const double udiagTmp = NW_XDIV(udiag, NW_XMUL(rho, dualVol));
  return NW_XADD(
    NW_XMUL(NW_XADD(udiagTmp, -projTimeScale), alphaU), projTimeScale);
}

NW_HD double
peclet_eval(const nw_peclet_fn& f, double pecnum)
{
  if (f.form == NW_PECLET_CLASSIC) {
    /* Exactly-rounded, un-contracted arithmetic on purpose: at high Peclet
     * number the factor is 1 - O(1e-11) and the kernels use (1 - pecfac), so a
     * 1-ulp difference in pecfac (FMA contraction of 5 + m*m, or a reciprocal-
     * multiply divide) shows up at ~1e-11 relative in off-diagonal entries.
     * Worse, where 5/m^2 is about half an ulp of 1 the rounding of pecfac
     * flips on any perturbation of the Peclet number, so the whole chain
     * (peclet_number too) uses IEEE mul/add/div in the reference's order: the
     * fused-Peclet path and the scalar kernel then see bit-identical factors
     * to the MomentumEdgePecletAlg edge field (src/PecletFunction.C:41-45). */
    const double modPeclet = NW_XMUL(f.a, pecnum);
    const double m2 = NW_XMUL(modPeclet, modPeclet);
    return NW_XDIV(m2, NW_XADD(5.0, m2));
  }
  return 0.50 * (1.0 + tanh((pecnum - f.a) * nw_rcp(f.b)));
}

NW_HD double
van_leer(double dqm, double dqp, double eps)
{
  return (2.0 * (dqm * dqp + fabs(dqm * dqp))) *
         nw_rcp((dqm + dqp) * (dqm + dqp) + eps);
}

/* ---- node state bundles (what each kernel stages per node) ---- */

template <int ND>
struct ContNode /* continuity + mdot: 3*ND + 3 doubles */
{
  double x[ND], u[ND], g[ND], rho, p, ud;
};
static const int kContNodeComps3 = 12;

template <int ND>
struct ScalNode /* scalar: 3*ND + 3 doubles */
{
  double x[ND], v[ND], dq[ND], q, rho, mu;
};

template <int ND>
struct MomNode /* momentum: 2*ND + ND*ND + 3 doubles */
{
  double x[ND], u[ND], g[ND * ND], mu, rho, mask;
};

template <int ND>
struct PecNode
{
  double x[ND], v[ND], rho, mu;
};

/* ---- mdot / continuity ---- */

template <int ND>
struct MdotCore
{
  double tmdot;         /* un-scaled edge mass flow rate (MdotEdgeAlg) */
  double asq_inv_axdx;  /* asq * inv_axdx */
  double projTimeScale;
  double rhoIp;
};

template <int ND>
NW_HD MdotCore<ND>
mdot_core(
  const ContNode<ND>& L,
  const ContNode<ND>& R,
  const double* av,
  double nocFac,
  double interpTogether)
{
  const double om_interpTogether = 1.0 - interpTogether;
  MdotCore<ND> r;
  /* one reciprocal per node instead of a divide per use: differs from the
   * reference's g/udiag by <= 1 ulp per term, far inside the 1e-12 bar, and
   * removes 2*ND fp64 divides (the FP64 pipe is the second roofline here) */
  const double invUdL = nw_rcp(L.ud), invUdR = nw_rcp(R.ud);
  r.projTimeScale = 0.5 * (invUdL + invUdR);
  r.rhoIp = 0.5 * (L.rho + R.rho);
  double axdx = 0.0, asq = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    asq += av[d] * av[d];
    axdx += av[d] * dxj;
  }
  const double inv_axdx = nw_rcp(axdx);
  double tmdot = -r.projTimeScale * (R.p - L.p) * asq * inv_axdx;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    const double kxj = av[d] - asq * inv_axdx * dxj;
    const double rhoUjIp = 0.5 * (R.rho * R.u[d] + L.rho * L.u[d]);
    const double ujIp = 0.5 * (R.u[d] + L.u[d]);
    const double GjIp = 0.5 * (R.g[d] * invUdR + L.g[d] * invUdL);
    tmdot += (interpTogether * rhoUjIp + om_interpTogether * r.rhoIp * ujIp +
              GjIp) *
               av[d] -
             kxj * GjIp * nocFac;
  }
  r.tmdot = tmdot;
  r.asq_inv_axdx = asq * inv_axdx;
  return r;
}

/* mdot_core plus the optional terms of MdotEdgeAlg / ContinuityEdgeSolverAlg
 * (src/ngp_algorithms/MdotEdgeAlg.C:153-163, 175-180;
 * src/edge_kernels/ContinuityEdgeSolverAlg.C:147-158, 172-177): balanced
 * buoyancy forcing (gravity, source, source mask) and the GCL term
 * (edge_face_velocity_mag).  Used by the *_ext kernels only. */
template <int ND>
struct ContExtra
{
  bool balanced, gcl;
  double gravity[ND];
  double smaskL, smaskR;
  double srcL[ND], srcR[ND];
  double faceVelMag;
};

template <int ND>
NW_HD MdotCore<ND>
mdot_core_ext(
  const ContNode<ND>& L,
  const ContNode<ND>& R,
  const ContExtra<ND>& x,
  const double* av,
  double nocFac,
  double interpTogether)
{
  const double om_interpTogether = 1.0 - interpTogether;
  MdotCore<ND> r;
  const double invUdL = nw_rcp(L.ud), invUdR = nw_rcp(R.ud);
  r.projTimeScale = 0.5 * (invUdL + invUdR);
  r.rhoIp = 0.5 * (L.rho + R.rho);
  double axdx = 0.0, asq = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    asq += av[d] * av[d];
    axdx += av[d] * dxj;
  }
  const double inv_axdx = nw_rcp(axdx);
  double tmdot = -r.projTimeScale * (R.p - L.p) * asq * inv_axdx;
  if (x.balanced) {
    const double masked_weights = 0.5 * (x.smaskL + x.smaskR);
#pragma unroll
    for (int d = 0; d < ND; ++d)
      tmdot += r.projTimeScale * av[d] * x.gravity[d] * r.rhoIp * masked_weights;
  }
  if (x.gcl)
    tmdot -= r.rhoIp * x.faceVelMag;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    const double kxj = av[d] - asq * inv_axdx * dxj;
    const double rhoUjIp = 0.5 * (R.rho * R.u[d] + L.rho * L.u[d]);
    const double ujIp = 0.5 * (R.u[d] + L.u[d]);
    double GjIp = 0.5 * (R.g[d] * invUdR + L.g[d] * invUdL);
    if (x.balanced)
      GjIp -= 0.5 * ((x.smaskR * x.srcR[d]) * invUdR + (x.smaskL * x.srcL[d]) * invUdL);
    tmdot += (interpTogether * rhoUjIp + om_interpTogether * r.rhoIp * ujIp +
              GjIp) *
               av[d] -
             kxj * GjIp * nocFac;
  }
  r.tmdot = tmdot;
  r.asq_inv_axdx = asq * inv_axdx;
  return r;
}

/* result: res[0] = lhsfac, res[1] = tmdot (scaled).
 * lhs = [[-f, +f], [+f, -f]], rhs = [-m, +m]. */
template <int ND>
NW_HD void
continuity_edge(
  const ContNode<ND>& L,
  const ContNode<ND>& R,
  const double* av,
  const nw_continuity_opts& o,
  double& lhsfac,
  double& tmdot_out)
{
  const double solveInc = o.solve_incompressible;
  const double om_solveInc = 1.0 - solveInc;
  MdotCore<ND> c = mdot_core<ND>(L, R, av, o.noc_fac, o.interp_together);
  const double denScale = nw_rcp(c.rhoIp) * solveInc + om_solveInc;
  const double invTauScale = o.gamma1 * nw_rcp(o.dt); /* 1 / (dt / gamma1) */
  double tmdot = c.tmdot;
  tmdot *= invTauScale;
  tmdot *= denScale;
  /* -asq * inv_axdx * projTimeScale * denScale / tauScale, left to right */
  lhsfac = -c.asq_inv_axdx * c.projTimeScale * denScale * invTauScale;
  tmdot_out = tmdot;
}

/* ---- Peclet factor ---- */

template <int ND>
NW_HD double
peclet_number(const PecNode<ND>& L, const PecNode<ND>& R, double eps)
{
  double udotx = 0.0;
  const double diffIp =
    NW_XMUL(0.5, NW_XADD(NW_XDIV(L.mu, L.rho), NW_XDIV(R.mu, R.rho)));
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    udotx = NW_XADD(
      udotx, NW_XMUL(NW_XMUL(0.5, dxj), NW_XADD(R.v[d], L.v[d])));
  }
  return NW_XDIV(fabs(udotx), NW_XADD(diffIp, eps));
}

/* ---- scalar ---- */

/* result: a[4] = lhs(0,0),lhs(0,1),lhs(1,0),lhs(1,1); flux: rhs = [-flux,+flux] */
template <int ND, bool DEF>
NW_HD void
scalar_edge_t(
  const ScalNode<ND>& L,
  const ScalNode<ND>& R,
  const double* av,
  double mdot,
  const nw_scalar_opts& o,
  double* a,
  double& flux)
{
  const double eps = o.eps;
  const double alpha = o.alpha;
  const double alphaUpw = o.alpha_upw;
  const double hoUpwind = o.ho_upwind;
  const double invRelax = nw_rcp(o.relax_fac); /* warp-uniform */
  const double om_alpha = 1.0 - alpha;
  const double om_alphaUpw = 1.0 - alphaUpw;

  const double viscIp = 0.5 * (L.mu + R.mu);

  double axdx = 0.0, asq = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    asq += av[d] * av[d];
    axdx += av[d] * dxj;
  }
  const double inv_axdx = nw_rcp(axdx);

  double dqL = 0.0, dqR = 0.0, nonOrth = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    dqL += 0.5 * dxj * L.dq[d];
    dqR += 0.5 * dxj * R.dq[d];
    const double kxj = av[d] - asq * inv_axdx * dxj;
    nonOrth += -viscIp * kxj * 0.5 * (R.dq[d] + L.dq[d]);
  }

  /* ScalarEdgeSolverAlg.C:128-143: same Peclet number as the momentum Peclet
   * algorithm, with D = diffFluxCoeff / rho */
  PecNode<ND> pl, pr;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    pl.x[d] = L.x[d];
    pr.x[d] = R.x[d];
    pl.v[d] = L.v[d];
    pr.v[d] = R.v[d];
  }
  pl.rho = L.rho;
  pr.rho = R.rho;
  pl.mu = L.mu;
  pr.mu = R.mu;
  const double pecfac = peclet_eval(o.pf, peclet_number<ND>(pl, pr, eps));
  const double om_pecfac = 1.0 - pecfac;

  double limitL = 1.0, limitR = 1.0;
  if (o.use_limiter) {
    const double dq = R.q - L.q;
    const double dqML = 4.0 * dqL - dq;
    const double dqMR = 4.0 * dqR - dq;
    limitL = van_leer(dqML, dq, eps);
    limitR = van_leer(dqMR, dq, eps);
  }

  const double qIpL = DEF ? L.q + dqL * limitL : L.q + dqL * hoUpwind * limitL;
  const double qIpR = DEF ? R.q - dqR * limitR : R.q - dqR * hoUpwind * limitR;

  const double lhsfac = -viscIp * asq * inv_axdx;
  const double diffFlux = lhsfac * (R.q - L.q) + nonOrth;

  double a00 = -lhsfac * invRelax;
  double a01 = lhsfac;
  double a10 = lhsfac;
  double a11 = -lhsfac * invRelax;

  const double qIp = 0.5 * (R.q + L.q);
  double qUpw, qCds;
  if (DEF) {
    qUpw = (mdot > 0) ? qIpL : qIpR;
    qCds = qIp;
  } else {
    qUpw = (mdot > 0) ? (alphaUpw * qIpL + om_alphaUpw * qIp)
                      : (alphaUpw * qIpR + om_alphaUpw * qIp);
    const double qHatL = (alpha * qIpL + om_alpha * qIp);
    const double qHatR = (alpha * qIpR + om_alpha * qIp);
    qCds = 0.5 * (qHatL + qHatR);
  }
  const double adv_flux = mdot * (pecfac * qUpw + om_pecfac * qCds);

  /* rhs(0) = -diffFlux - adv_flux, rhs(1) = diffFlux + adv_flux; negation is
   * exact, so one number carries both. */
  flux = diffFlux + adv_flux;

  double alhsfac = DEF ? 0.5 * (mdot + fabs(mdot)) * pecfac
                       : 0.5 * (mdot + fabs(mdot)) * pecfac * alphaUpw +
                           0.5 * alpha * om_pecfac * mdot;
  a00 += alhsfac * invRelax;
  a10 -= alhsfac;

  alhsfac = DEF ? 0.5 * (mdot - fabs(mdot)) * pecfac
                : 0.5 * (mdot - fabs(mdot)) * pecfac * alphaUpw +
                    0.5 * alpha * om_pecfac * mdot;
  a11 -= alhsfac * invRelax;
  a01 += alhsfac;

  alhsfac = DEF ? 0.5 * mdot * om_pecfac
                : 0.5 * mdot * (pecfac * om_alphaUpw + om_pecfac * om_alpha);
  a00 += alhsfac * invRelax;
  a01 += alhsfac;
  a10 -= alhsfac;
  a11 -= alhsfac * invRelax;

  a[0] = a00;
  a[1] = a01;
  a[2] = a10;
  a[3] = a11;
}

template <int ND>
NW_HD void
scalar_edge(
  const ScalNode<ND>& L,
  const ScalNode<ND>& R,
  const double* av,
  double mdot,
  const nw_scalar_opts& o,
  double* a,
  double& flux)
{
  if (o.alpha == 0.0 && o.alpha_upw == 1.0 && o.ho_upwind == 1.0)
    scalar_edge_t<ND, true>(L, R, av, mdot, o, a, flux);
  else
    scalar_edge_t<ND, false>(L, R, av, mdot, o, a, flux);
}

/* ---- momentum ---- */

/* Per-edge momentum result.  The 2ND x 2ND block the reference builds is
 *   lhs(rowL(i), colL(j)) = d_ij * sLL - NS_ij / relax
 *   lhs(rowL(i), colR(j)) = d_ij * sLR + NS_ij
 *   lhs(rowR(i), colL(j)) = d_ij * sRL + NS_ij
 *   lhs(rowR(i), colR(j)) = d_ij * sRR - NS_ij / relax
 * with NS_ij = -viscIp * av[i] * av[j] * inv_axdx, accumulated in the
 * reference's order (same-component terms first, then the NS terms for
 * j = 0..ND-1).  flux[i]: rhs(rowL(i)) = -flux[i], rhs(rowR(i)) = +flux[i]. */
template <int ND>
struct MomResult
{
  double sLL, sLR, sRL, sRR; /* same-component advection+diffusion part */
  double flux[ND];
  double viscIp, inv_axdx;
};

/* DEF: the deck's default upwinding options, alpha = 0, alpha_upw = 1,
 * hoUpwind = 1 (every regression deck of SURVEY 8 runs them).  The blending
 * products then reduce exactly -- x * 1 = x, 0 * y = +-0 and a + (+-0) = a for
 * finite y, a != 0 -- so the specialised path gives the same bits as the
 * general one (at most the sign of an exact zero differs) with ~10 % fewer
 * FP64 instructions; momentum_edge_any picks the path (warp-uniform). */
template <int ND, bool DEF, bool VOF = false>
NW_HD void
momentum_edge_t(
  const MomNode<ND>& L,
  const MomNode<ND>& R,
  const double* av,
  double mdot,
  double pecfac,
  const nw_momentum_opts& o,
  MomResult<ND>& res)
{
  static_assert(!(DEF && VOF), "the VOF branch changes alphaUpw per edge");
  const double eps = o.eps;
  const double includeDivU = o.include_divu;
  const double alpha = o.alpha;
  double alphaUpw = o.alpha_upw;
  const double hoUpwind = o.ho_upwind;
  const double invRelaxU = nw_rcp(o.relax_fac); /* warp-uniform */
  const double om_alpha = 1.0 - alpha;
  double om_alphaUpw = 1.0 - alphaUpw;
  double density_upwinding_factor = 1.0; /* has_vof == 0 */

  const double viscIp = 0.5 * (L.mu + R.mu);

  double axdx = 0.0, asq = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double dxj = R.x[d] - L.x[d];
    asq += av[d] * av[d];
    axdx += av[d] * dxj;
  }
  const double inv_axdx = nw_rcp(axdx);

  double duL[ND], duR[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    duL[i] = 0.0;
    duR[i] = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const double dxj = 0.5 * (R.x[j] - L.x[j]);
      duL[i] += dxj * L.g[i * ND + j];
      duR[i] += dxj * R.g[i * ND + j];
    }
  }

  double limitL[ND], limitR[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    limitL[d] = 1.0;
    limitR[d] = 1.0;
  }
  if (o.use_limiter) {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const double du = R.u[d] - L.u[d];
      const double duML = 4.0 * duL[d] - du;
      const double duMR = 4.0 * duR[d] - du;
      limitL[d] = van_leer(duML, du, eps);
      limitR[d] = van_leer(duMR, du, eps);
    }
  }

  double om_pecfac = 1.0 - pecfac;
  if (VOF) {
    /* upwinding switch for multiphase cases: full upwinding across an
     * interface (src/edge_kernels/MomentumEdgeSolverAlg.C:174-192) */
    const double min_density = fmin(L.rho, R.rho);
    const double density_differential = fabs(L.rho - R.rho) / min_density;
    density_upwinding_factor = 1.0 - erf(6.0 * density_differential);
    alphaUpw = density_upwinding_factor * alphaUpw + (1.0 - density_upwinding_factor);
    om_alphaUpw = 1.0 - alphaUpw;
    pecfac = 1.0 - density_upwinding_factor + density_upwinding_factor * pecfac;
    om_pecfac = 1.0 - pecfac;
  }

  double uIpL[ND], uIpR[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    if (DEF) {
      uIpL[d] = L.u[d] + duL[d] * limitL[d];
      uIpR[d] = R.u[d] - duR[d] * limitR[d];
    } else {
      uIpL[d] = L.u[d] + duL[d] * hoUpwind * limitL[d] * density_upwinding_factor;
      uIpR[d] = R.u[d] - duR[d] * hoUpwind * limitR[d] * density_upwinding_factor;
    }
  }

  double duidxj[ND][ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const double dui = R.u[i] - L.u[i];
    double gjuidx = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const double dxj = R.x[j] - L.x[j];
      const double gjui = 0.5 * (R.g[i * ND + j] + L.g[i * ND + j]);
      gjuidx += gjui * dxj;
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const double gjui = 0.5 * (R.g[i * ND + j] + L.g[i * ND + j]);
      duidxj[i][j] = gjui + (dui - gjuidx) * av[j] * inv_axdx;
    }
  }

  const double dlhsfac = -viscIp * asq * inv_axdx;
  const double maskNode = fmin(L.mask, R.mask);

#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const double uiIp = 0.5 * (R.u[i] + L.u[i]);
    double uiUpw, uiCds;
    if (DEF) {
      uiUpw = (mdot > 0.0) ? uIpL[i] : uIpR[i];
      uiCds = uiIp; /* 0.5 * (uiIp + uiIp) */
    } else {
      uiUpw = (mdot > 0.0) ? (alphaUpw * uIpL[i] + om_alphaUpw * uiIp)
                           : (alphaUpw * uIpR[i] + om_alphaUpw * uiIp);
      const double uiHatL = (alpha * uIpL[i] + om_alpha * uiIp);
      const double uiHatR = (alpha * uIpR[i] + om_alpha * uiIp);
      uiCds = 0.5 * (uiHatL + uiHatR);
    }
    const double adv_flux = mdot * (pecfac * uiUpw + om_pecfac * uiCds);

    /* divU term: the reference always forms (sum_j duidxj[j][j]) * 2/3 mu a_i *
     * includeDivU; with includeDivU == 0 (warp-uniform option) that is an exact
     * +-0 added to a sum, so skipping it changes no value (at most the sign of
     * an exact zero) and saves ~7 % of this kernel's FP64 instructions */
    double diff_flux = 0.0;
    if (includeDivU != 0.0) {
#pragma unroll
      for (int j = 0; j < ND; ++j)
        diff_flux += duidxj[j][j];
      diff_flux *= 2.0 / 3.0 * viscIp * av[i] * includeDivU;
    }
#pragma unroll
    for (int j = 0; j < ND; ++j)
      diff_flux += -viscIp * (duidxj[i][j] + duidxj[j][i]) * av[j];

    res.flux[i] = adv_flux + diff_flux * maskNode;
  }

  /* same-component Jacobian terms: identical for every i, accumulated in the
   * reference's order onto a zeroed entry (AssembleEdgeSolverAlgorithm.h:89) */
  double sLL = 0.0, sLR = 0.0, sRL = 0.0, sRR = 0.0;
  double alhsfac = DEF ? 0.5 * (mdot + fabs(mdot)) * pecfac
                       : 0.5 * (mdot + fabs(mdot)) * pecfac * alphaUpw +
                           0.5 * alpha * om_pecfac * mdot;
  sLL += alhsfac * invRelaxU;
  sRL -= alhsfac;

  alhsfac = DEF ? 0.5 * (mdot - fabs(mdot)) * pecfac
                : 0.5 * (mdot - fabs(mdot)) * pecfac * alphaUpw +
                    0.5 * alpha * om_pecfac * mdot;
  sRR -= alhsfac * invRelaxU;
  sLR += alhsfac;

  alhsfac = DEF ? 0.5 * mdot * om_pecfac
                : 0.5 * mdot * (pecfac * om_alphaUpw + om_pecfac * om_alpha);
  sLL += alhsfac * invRelaxU;
  sLR += alhsfac;
  sRL -= alhsfac;
  sRR -= alhsfac * invRelaxU;

  sLL -= dlhsfac * invRelaxU;
  sLR += dlhsfac;
  sRL += dlhsfac;
  sRR -= dlhsfac * invRelaxU;

  res.sLL = sLL;
  res.sLR = sLR;
  res.sRL = sRL;
  res.sRR = sRR;
  res.viscIp = viscIp;
  res.inv_axdx = inv_axdx;
}

/* realm_has_vof_: mdot is massFlowRate + massVofBalancedFlowRate (summed by the
 * caller of the kernel, MomentumEdgeSolverAlg.C:124-125), the upwinding factors
 * depend on the density jump of the edge */
template <int ND>
NW_HD void
momentum_edge_vof(
  const MomNode<ND>& L,
  const MomNode<ND>& R,
  const double* av,
  double mdot,
  double pecfac,
  const nw_momentum_opts& o,
  MomResult<ND>& res)
{
  momentum_edge_t<ND, false, true>(L, R, av, mdot, pecfac, o, res);
}

template <int ND>
NW_HD void
momentum_edge(
  const MomNode<ND>& L,
  const MomNode<ND>& R,
  const double* av,
  double mdot,
  double pecfac,
  const nw_momentum_opts& o,
  MomResult<ND>& res)
{
  if (o.alpha == 0.0 && o.alpha_upw == 1.0 && o.ho_upwind == 1.0)
    momentum_edge_t<ND, true>(L, R, av, mdot, pecfac, o, res);
  else
    momentum_edge_t<ND, false>(L, R, av, mdot, pecfac, o, res);
}

/* entry (i,j) of the four ND x ND sub-blocks, reference accumulation order:
 * for row component i the j-loop adds NS_i0, NS_i1, NS_i2 in turn, and the
 * same-component part was added before the j-loop (MomentumEdgeSolverAlg.C:275-310). */
template <int ND>
NW_HD void
momentum_block_entry(
  const MomResult<ND>& r,
  const double* av,
  double relaxFacU,
  int i,
  int j,
  double& LL,
  double& LR,
  double& RL,
  double& RR)
{
  const double lhsfacNS = -r.viscIp * av[i] * av[j] * r.inv_axdx;
  const double s = (i == j) ? 1.0 : 0.0;
  const double invRelaxU = nw_rcp(relaxFacU);
  LL = s * r.sLL - lhsfacNS * invRelaxU;
  LR = s * r.sLR + lhsfacNS;
  RL = s * r.sRL + lhsfacNS;
  RR = s * r.sRR - lhsfacNS * invRelaxU;
}

} // namespace nw

#endif
