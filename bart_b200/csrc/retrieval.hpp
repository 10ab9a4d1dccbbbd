// retrieval.hpp -- the retrieval loop around the forward model, on the device (retrieval.cu):
//   K-1 convert_params_kernel   input converter, code/BARTfunc.py:320-360 + code/PT.py:589-750
//   K7  demc_propose_kernel     DE-MC proposal, modules/MCcubed/MCcubed/mc/mcmc.py:524-575
//   K8  chisq_accept_kernel     chi-squared (src_c/chisq.c:111-142, include/stats.h:72-103),
//                               Metropolis rule, best fit, trace (mcmc.py:590-625)
#pragma once
#include "device.cuh"
#include <cuda_runtime.h>

namespace bart {

enum { REJ_TBOUNDS = 16, REJ_ABUND = 32 };
enum { PT_ISO = 0, PT_LINE = 1, PT_ADIABATIC = 2 };
constexpr int kMaxPars = 64;

// Input converter set-up (BARTfunc.py:139-222).  Arrays live on the device.
struct ConvConfig {
  int ready;
  int pt_type, npt, nrad, ncloud, nray, nmolfit, nmetals, npars;
  int nlayer, nspec;
  int imol[kMaxGridMol];
  int imetals[kMaxSpec];
  int iH2, iHe;
  double tmin, tmax;
  double rstar, tstar, tint, sma, grav;    // PT_line arguments (BARTfunc.py:206-211); tint final
  const double *press_bar;                 // [nlayer] atmosphere-file order (bottom -> top)
  const double *base;                      // [nspec][nlayer] abundances of the atmosphere file
  const double *ratio;                     // [nlayer] H2/He
};

// per-model knob arrays written by the converter (BARTfunc.py:350-360)
struct ConvKnobs {
  double *r0, *cloudtop, *scat_logext;
  int *scat_flag;
};

void launch_convert_params(const ConvConfig &cc, const double *params, int npars, double *profiles,
                           int n_in, int *status, const ConvKnobs &kn, int nmodels, cudaStream_t s);

// DE-MC state (device pointers), one population of `nchains` chains
struct McmcDev {
  int nchains, npars, nfree, ndata, nprior, chainsize, burnin, nold;
  double gamma, fepsilon;
  int ifree[kMaxPars];
  int share_dst[kMaxPars], share_src[kMaxPars], nshare;
  int iprior[kMaxPars];
  const double *pmin, *pmax, *prior, *priorlow, *data, *uncert;
  double *params, *nextp, *currchisq, *nextchisq, *c2, *bestp, *bestchisq, *bestmodel;
  double *numaccept, *allparams;
  int *outbounds, *outflag, *iter;
  // random streams of this run, MC3's shapes (mcmc.py:484-507)
  const double *support;   // [chainsize][nchains][nfree]
  const int *r1, *r2;      // [nchains][chainsize]
  const double *unif;      // [chainsize][nchains]
  const double *ugamma;    // [chainsize][nchains]
};

// where chain c's model sits in the gathered band-flux buffer: rank blocks of `pad` rows
struct ModelMap { int world, base, extra, pad; };

void launch_demc_propose(const McmcDev &mc, cudaStream_t s);
// first = 1: initial state (mcmc.py:310-345): chi-squared of the current parameters, best fit
void launch_chisq_accept(const McmcDev &mc, const double *models, ModelMap map, int first,
                         cudaStream_t s);

}  // namespace bart
