// device.cuh -- device-side data model shared by kernels.cu / builder.cu / transit.cu.
//
// HBM layout (all fp64 unless stated):
//   grid   o[layer][temp][mol][wave]   the opacity file's own order (opacity.c:418-421): the
//                                      wavenumber axis is contiguous, so thread <-> wavenumber
//                                      gives perfectly coalesced streams with no transpose.
//   ciaP/ciaQ[file][temp][wave]        CIA tables pre-folded through the wavenumber spline.
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

// field indices inside one model table (each field is `nlayer` doubles, depth-indexed)
struct TabLayout {
  int nl, ngmol, ncia;
  BART_HD int T() const { return 0; }
  BART_HD int IT() const { return 1; }                        // bracket index (exact int)
  BART_HD int W0(int m) const { return 2 + 2 * m; }          // rho*(t1-T)/(t1-t0)
  BART_HD int W1(int m) const { return 3 + 2 * m; }          // rho*(T-t0)/(t1-t0)
  BART_HD int CIAK(int f) const { return 2 + 2 * ngmol + 5 * f; }
  BART_HD int CIAC(int f, int c) const { return 3 + 2 * ngmol + 5 * f + c; }
  BART_HD int SCAT() const { return 2 + 2 * ngmol + 5 * ncia; }
  BART_HD int CLOUD() const { return SCAT() + 1; }
  BART_HD int SA() const { return SCAT() + 2; }              // Simpson panel coefficients
  BART_HD int SB() const { return SCAT() + 3; }
  BART_HD int SC() const { return SCAT() + 4; }
  BART_HD int TR() const { return SCAT() + 5; }              // trapezoid half-width
  BART_HD int RAD() const { return SCAT() + 6; }             // radius (file units)
  BART_HD int nfields() const { return SCAT() + 7; }
  // doubles per model, padded to a multiple of 2 (16-byte bulk-copy granularity)
  BART_HD int stride() const { int n = nfields() * nl; return (n + 1) & ~1; }
};

struct DevConfig {
  int nlayer, nspec, nwave, ntemp, ngmol, ncia, nang;
  int eclipse, transparent;
  const double *grid;
  const double *gtemp;
  const double *wn;
  const double *press;      // [nlayer] atmosphere-file units, bottom -> top
  const double *mass;       // [nspec]
  const double *pol;        // [nspec]
  int gmol_spec[kMaxGridMol];
  const double *ciaP[kMaxCia];
  const double *ciaQ[kMaxCia];
  const double *ciaT[kMaxCia];
  int cia_nt[kMaxCia];
  int cia_nspec[kMaxCia];
  int cia_spec[kMaxCia][2];
  double pfct, rfct, gsurf, p0, toomuch;
  double inv_mu[kMaxAng];   // 1/cos(angle)
  double wgt[kMaxAng];      // sin^2(g_{a+1}) - sin^2(g_a)
  double inv_srad2;         // 1/R*^2 (cm^-2)
  TabLayout lay;
};

// per-model knobs (BARTfunc.py:350-360 sets these before each run_transit)
struct Knobs {
  const double *r0;          // [M] or nullptr -> r0_all
  const double *cloudtop;    // [M] or nullptr
  const int *scat_flag;      // [M] or nullptr
  const double *scat_logext; // [M] or nullptr
  double r0_all;
  int cloud_flag_all; double cloudext_all, cloudtop_all, cloudbot_all;
  int scat_flag_all; double scat_logext_all;
};

}  // namespace bart
