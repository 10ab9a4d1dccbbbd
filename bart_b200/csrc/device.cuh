// device.cuh -- device-side data model shared by kernels.cu / builder.cu / transit.cu.
//
// HBM layout (all fp64 unless stated):
//   grid   o[layer][temp][wave][gms]   the opacity file (o[layer][temp][mol][wave],
//                                      opacity.c:418-421) with the molecule axis moved innermost
//                                      at upload time (gms = molecules per sample, padded to an
//                                      even count when > 1): thread <-> wavenumber reads one
//                                      (layer, temperature) plane as a contiguous stream of
//                                      16-byte vector loads carrying all molecules of a sample.
//   ciaPQ[file][temp][wave][4]         CIA tables pre-folded through the wavenumber spline: (value,
//                                      temperature second derivative) at node k and at node k+1,
//                                      so one 32-byte load serves a (layer, wavenumber) cell.
//   profiles[model][(1+nspec)*nlayer]  the caller's per-model input, verbatim.
//   tab[model][field][depth]           per-model, per-layer coefficients written by atm_prep and
//                                      staged into shared memory (one bulk-async copy) by the
//                                      column kernels; depth 0 = top layer.
//   spectra[model][wave]
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define BART_HD __host__ __device__ __forceinline__
#else
#define BART_HD inline
#endif

namespace bart {

constexpr int kMaxGridMol = 16;
constexpr int kMaxCia = 4;
constexpr int kMaxAng = 16;
constexpr int kMaxSpec = 64;
constexpr int kTaylorN = 12;   // polynomial degree 11

// One model table = nlayer records of `nf()` doubles, one record per depth (0 = top layer), so a
// column kernel reads everything it needs for a layer from one contiguous, 16-byte aligned
// shared-memory record with compile-time offsets (warp-wide broadcast loads, LDS.128 for pairs).
//   [0] 1/T      [1] exp(planck_step/T): ratio of the Planck exponentials of two columns planck_cols apart
//   [2] scattering coef (x wn^4)   [3] cloud    [4..6] Simpson panel coefficients   [7] trapezoid half-width
//   [8] T        [9] radius (file units)
//   [10] byte offset (int64 bits) of grid plane (layer, it)   [11] spare
//   [12+2m, 13+2m]  W0, W1 of grid molecule m:  rho*(t1-T)/(t1-t0), rho*(T-t0)/(t1-t0)
//   [cia(f) .. +5]  CIA file f: table byte offset (int64 bits), bracket index, 4 cubic coefficients
struct TabLayout {
  int nl, ngmol, ncia;
  static constexpr int INVT = 0, PF = 1, SCAT = 2, CLOUD = 3, SA = 4, SB = 5, SC = 6, TR = 7,
                       T = 8, RAD = 9, GOFF = 10, W = 12;
  BART_HD int cia(int f) const { return W + 2 * ngmol + 6 * f; }
  BART_HD int nf() const { return W + 2 * ngmol + 6 * ncia; }
  BART_HD int stride() const { return nf() * nl; }   // nf() is even: 16-byte granularity holds
};

struct DevConfig {
  int nlayer, nspec, nwave, ntemp, ngmol, ncia, nang;
  int gms;                  // grid doubles per wavenumber sample (ngmol, padded to even when > 1)
  int eclipse, transparent, modlevel;
  const double *grid;
  const double *gtemp;
  const double *wn;
  const double *press;      // [nlayer] atmosphere-file units, bottom -> top
  const double *mass;       // [nspec]
  const double *pol;        // [nspec]
  int gmol_spec[kMaxGridMol];
  const double *ciaPQ[kMaxCia];
  const double *ciaT[kMaxCia];
  int cia_nt[kMaxCia];
  int cia_nspec[kMaxCia];
  int cia_spec[kMaxCia][2];
  double pfct, rfct, gsurf, p0, toomuch;
  int ref_layer;            // layer nearest to the reference pressure p0 (column_math.cuh ref_layer_of)
  double inv_mu[kMaxAng];   // 1/cos(angle)
  double wgt[kMaxAng];      // sin^2(g_{a+1}) - sin^2(g_a)
  double inv_srad2;         // 1/R*^2 (cm^-2)
  // hemispheric transmission D(tau) = sum_a wgt[a] exp(-tau inv_mu[a]) (column_math.cuh):
  double exp_a[kMaxAng];    // inv_mu[a] * N/ln2, the exponent in units of the exp table
  double taylor[kTaylorN];  // D on [0, tau_small] as a polynomial in u = tau ser_s - 1 (fill_angle_consts)
  double tau_small;         // warp-uniform switch to the polynomial
  double d0, ser_s;         // D(0); 2 / tau_small
  double tau_clamp;         // exp arguments stay above -700
  int sq_src, sq_dst;       // angles with inv_mu[sq_dst] = 2 inv_mu[sq_src] (exp by squaring), or -1
  // weighted angle exponentials wgt[a] exp(-tau inv_mu[a]) (column_math.cuh exp_w): the weight is
  // folded into a per-angle copy of the 2^(j/N) table (fill_ecl_exp_table), so the angle sum rides
  // on the last multiply-add of each exponential
  double sq_coef;           // wgt[sq_dst] / wgt[sq_src]^2
  // 1: some layer/wavenumber of this configuration can have a Planck exponent hc wn/(k T) above 690
  // or below 0.1 -> the column kernel with the per-column clamp and the degree-5 exponential
  int planck_generic;
  // Planck chaining: a thread's columns are planck_cols samples apart on the uniform wavenumber
  // grid, so exp(c2 wn'/T) = exp(c2 wn/T) * exp(planck_step/T), planck_step = (hc/k) planck_cols dwn
  int planck_cols;          // 0 = no chaining
  double planck_step;
  // line-by-line mode (no opacity file, tau.c:163-175): `grid` is the per-batch buffer
  // ext[model][layer][wave] filled by the builder kernels (one "molecule" with weight 1, second
  // bracket weight 0); gtemp = {TLI tmin, TLI tmax}; atm_prep also stores the mass densities
  int lbl;
  int lbl_model0;           // global index of the launch's first model
  double *lbl_dens;         // [model][layer][nspec]
  TabLayout lay;
};

// per-model knobs (BARTfunc.py:350-360 sets these before each run_transit)
struct Knobs {
  const double *r0;          // [M] or nullptr -> r0_all
  const double *cloudtop;    // [M] or nullptr
  const int *scat_flag;      // [M] or nullptr
  const double *scat_logext; // [M] or nullptr
  const double *radius_file; // [nlayer] or nullptr: use these radii (file units, bottom -> top)
                             // instead of the hydrostatic ones -- the command-line program runs
                             // do_transit on the atmosphere file as read, without reloadatm
                             // (transit.c:230-242, readatm.c:499)
  double r0_all;
  int cloud_flag_all; double cloudext_all, cloudtop_all, cloudbot_all;
  int scat_flag_all; double scat_logext_all;
};

}  // namespace bart
