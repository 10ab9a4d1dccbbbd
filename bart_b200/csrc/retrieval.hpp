// retrieval.hpp -- the retrieval loop around the forward model, on the device (retrieval.cu):
//   K-1 convert_params_kernel   input converter, code/BARTfunc.py:320-360 + code/PT.py:589-750
//   K7  demc_propose_kernel     DE-MC proposal, modules/MCcubed/MCcubed/mc/mcmc.py:524-575
//   K8  chisq_accept_kernel     chi-squared (src_c/chisq.c:111-142, include/stats.h:72-103),
//                               Metropolis rule, best fit, trace (mcmc.py:590-625), snooker
//                               Metropolis factor (603-609) and Z update (653-660)
//   K9  snooker_propose_kernel  DE-MC-with-snooker proposal from the sample history Z
//                               (ter Braak & Vrugt 2008; mcmc.py:527-561)
//   K10 zrow_chisq_kernel       chi-squared of the initial Z samples (mcmc.py:426-460)
#pragma once
#include "device.cuh"
#include <cuda_runtime.h>

namespace bart {

enum { REJ_TBOUNDS = 16, REJ_ABUND = 32, REJ_ENERGY = 128, REJ_PTMODEL = 256 };
enum { PT_ISO = 0, PT_LINE = 1, PT_ADIABATIC = 2, PT_MADHU_NOINV = 3, PT_MADHU_INV = 4, PT_PIETTE = 5 };
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
  // PT models that smooth over the layers (PT.py:157-586 Madhusudhan & Seager, 752-812 Piette):
  double p_top, p_bot;                     // min / max pressure (bar)
  int smooth_r;                            // Gaussian kernel radius, int(4 sigma + 0.5)
  const double *smooth_w;                  // [2 r + 1] normalised weights
  const double *node_x;                    // [nlayer] log10 p (Piette's interpolation abscissa)
  const int *node_seg;                     // [nlayer] knot interval of every layer
  double node_t[8];                        // Piette's eight knots (log10 p of the node layers)
};

// per-model knob arrays written by the converter (BARTfunc.py:350-360)
struct ConvKnobs {
  double *r0, *cloudtop, *scat_logext;
  int *scat_flag;
};

void launch_convert_params(const ConvConfig &cc, const double *params, int npars, double *profiles,
                           int n_in, int *status, const ConvKnobs &kn, int nmodels, cudaStream_t s);

// Energy-balance test of BARTfunc.py:366-383 on the spectra of a batch: a model whose outgoing
// energy trapz(spectrum, wn) * out_scale exceeds e_in gets REJ_ENERGY (its band fluxes then come
// out as -1, like every rejected model).  Models already rejected are left alone.
void launch_energy_balance(const double *spectra, const double *wn, int nwave, double out_scale,
                           double e_in, int *status, int nmodels, cudaStream_t s);

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
  // MC3's `savemodel` trace (mcmc.py:636-651): the model of every chain's CURRENT state after each
  // generation, allmodel[chain][datum][iteration].  curmodel starts as zeros, like the reference's
  // `allmodel[~accepted,:,i+nold-1]` with i + nold - 1 = -1 at the first generation.
  double *curmodel, *allmodel;
  int *outbounds, *outflag, *iter;
  // random streams of this run, MC3's shapes (mcmc.py:484-507)
  const double *support;   // [chainsize][nchains][nfree]
  const int *r1, *r2;      // [nchains][chainsize]
  const double *unif;      // [chainsize][nchains]
  const double *ugamma;    // [chainsize][nchains]
  // walk = 1: snooker (mcmc.py:357-460,527-561,603-609,653-660)
  int walk, hsize, thinning;
  double *Z;               // [zcap][nchains][npars] sample history; rows carry the chains'
                           // initial non-free columns like the reference's (mcmc.py:419,425)
  double *Zchisq;          // [zcap][nchains]
  int *zsize;              // rows filled so far (device counter, advances with the generations)
  double *zbest;           // [1 + npars + ndata] best initial Z sample: chisq, params, model
  const int *i1, *i2;      // [chainsize][nchains] flat (row, chain) indices into Z
  const int *iz, *ic;      // [chainsize][nchains]
  const double *usn;       // [sum of snooker chains][nfree] uniform(1.2, 2.2) factors
  const int *usn_off;      // [chainsize + 1] first row of generation i in usn
  int *noproj;             // [nchains] scratch: z == current state (mcmc.py:543)
  int *slot;               // [nchains] scratch: row of the chain's factors in usn
};

// where chain c's model sits in the gathered band-flux buffer: rank blocks of `pad` rows
struct ModelMap { int world, base, extra, pad; };

void launch_demc_propose(const McmcDev &mc, cudaStream_t s);
void launch_snooker_propose(const McmcDev &mc, cudaStream_t s);
// chi-squared of Z row `row` (models of its nchains samples); keeps the running best in mc.zbest.
// last != 0: afterwards adopt zbest as the best fit when it beats the chains' (mcmc.py:462-470)
void launch_zrow_chisq(const McmcDev &mc, const double *models, ModelMap map, int row, int last,
                       cudaStream_t s);
// first = 1: initial state (mcmc.py:310-345): chi-squared of the current parameters, best fit
void launch_chisq_accept(const McmcDev &mc, const double *models, ModelMap map, int first,
                         cudaStream_t s);

}  // namespace bart
