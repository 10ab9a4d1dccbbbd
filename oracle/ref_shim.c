/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiled with the reference's own headers and flags and linked with the UNMODIFIED reference
 * objects into oracle/_ref/libtransit_ref.so (see oracle/Makefile).  It adds no arithmetic: it
 * only (i) replays the reference's own stage sequence of do_transit()
 * (modules/transit/transit/src/transit.c:125-214) WITHOUT the trailing frees so the
 * intermediates stay readable, and (ii) hands out raw pointers to reference globals so the
 * oracle restatement and the CUDA path can be compared stage by stage at full fp64 precision
 * (the reference's `savefiles yes` text dumps carry only ~10 digits).
 *
 * Never imported by the product; only tests/, smoke() and bench.py's cpu_baseline leg load it.
 */
#include <transit.h>

extern struct transit transit;
extern int init_run;

static int ref_kept = 0;

static void ref_release(void){
  if (!ref_kept) return;
  /* Same frees as transit.c:202-207                                         */
  free(transit.save.ext);
  freemem_samp(&transit.ips);
  freemem_idexrefrac(transit.ds.ir,  &transit.pi);
  freemem_extinction(transit.ds.ex,  &transit.pi);
  freemem_tau(       transit.ds.tau, &transit.pi);
  freemem_outputray( transit.ds.out, &transit.pi);
  ref_kept = 0;
}

/* Run one forward model exactly as run_transit() would (transit.c:118-214) but keep the
 * per-call arrays alive until the next ref_run_keep()/ref_release_model() call.           */
void ref_run_keep(double *re_input, double *transit_out){
  int i;
  ref_release();
  fw(reloadatm, <0, &transit, re_input);
  fw(makeipsample, <0, &transit);
  fw(interpcs, !=0, &transit);
  fw(idxrefrac, !=0, &transit);
  fw(extwn, !=0, &transit);
  fw(init_optdepth, !=0, &transit);
  fw(tau, !=0, &transit);
  if (strcmp(transit.sol->name, "eclipse") == 0){
    for (i=0; i < transit.ann; i++){
      transit.angleIndex = i;
      fw(emergent_intens, !=0, &transit);
    }
    fw(flux, !=0, &transit);
    freemem_intensityGrid(transit.ds.intens, &transit.pi);
  }
  else{
    fw(modulation, !=0, &transit);
  }
  for (i=0; i < transit.wns.n; i++)
    transit_out[i] = transit.ds.out->o[i];
  ref_kept = 1;
}

void ref_release_model(void){ ref_release(); }

/* ---- raw views of reference state (valid after transit_init / ref_run_keep) ---- */
long    ref_nlayers(void){ return (long)transit.rads.n; }
long    ref_nwave(void)  { return (long)transit.wns.n; }
long    ref_nmol(void)   { return (long)transit.ds.mol->nmol; }
int     ref_is_eclipse(void){ return strcmp(transit.sol->name, "eclipse") == 0; }
double *ref_radius(void) { return transit.rads.v; }          /* [nlayer], units rads.fct   */
double  ref_radfct(void) { return transit.rads.fct; }
double *ref_temp(void)   { return transit.atm.t; }           /* [nlayer]                   */
double *ref_press(void)  { return transit.atm.p; }           /* [nlayer], units atm.pfct   */
double  ref_pfct(void)   { return transit.atm.pfct; }
double *ref_mm(void)     { return transit.atm.mm; }          /* [nlayer]                   */
double *ref_density(int imol){ return transit.ds.mol->molec[imol].d; }  /* [nlayer]        */
double *ref_abund(int imol)  { return transit.ds.mol->molec[imol].q; }  /* [nlayer]        */
int     ref_molid(int imol)  { return transit.ds.mol->ID[imol]; }
double  ref_molmass(int imol){ return transit.ds.mol->mass[imol]; }
double *ref_ext(void)    { return transit.ds.ex->e[0]; }     /* [nlayer][nwave]            */
short  *ref_ext_computed(void){ return (short *)transit.ds.ex->computed; }
double *ref_cia(void)    { return transit.ds.cross->e[0]; }  /* [nwave][nlayer]            */
double *ref_tau(void)    { return transit.ds.tau->t[0]; }    /* [nwave][nlayer]            */
long   *ref_last(void)   { return transit.ds.tau->last; }    /* [nwave]                    */
double  ref_toomuch(void){ return transit.ds.tau->toomuch; }
int     ref_nangles(void){ return transit.ann; }
double *ref_angles(void) { return transit.angles; }
double  ref_starrad_cm(void){ return transit.ds.sg->starrad * transit.ds.sg->starradfct; }
double  ref_p0(void)     { return transit.p0; }
double  ref_r0(void)     { return transit.r0; }
double  ref_gsurf(void)  { return transit.gsurf; }

/* Opacity grid as read/built by the reference (opacity.c:432-503)                          */
long    ref_op_dims(long *dims){
  struct opacity *op = transit.ds.op;
  dims[0] = op->Nmol; dims[1] = op->Ntemp; dims[2] = op->Nlayer; dims[3] = op->Nwave;
  return 0;
}
double *ref_op_temp(void){ return transit.ds.op->temp; }
int    *ref_op_molid(void){ return transit.ds.op->molID; }
double *ref_op_row(long r, long t, long m){ return transit.ds.op->o[r][t][m]; }

/* Voigt-profile table built by calcprofiles() (opacity.c:218-277), builder runs only.      */
long    ref_prof_size(int idop, int ilor){ return (long)transit.ds.op->profsize[idop][ilor]; }
float  *ref_prof(int idop, int ilor){ return transit.ds.op->profile[idop][ilor]; }
double *ref_adop(void){ return transit.ds.op->aDop; }
double *ref_alor(void){ return transit.ds.op->aLor; }
int     ref_ndop(void){ return (int)transit.ds.op->nDop; }
int     ref_nlor(void){ return (int)transit.ds.op->nLor; }

/* Oversampled wavenumber grid (makesample.c:77-104) for bin-index parity                   */
long    ref_nowns(void){ return (long)transit.owns.n; }
double *ref_owns(void){ return transit.owns.v; }
int     ref_osamp(void){ return transit.owns.o; }

/* Direct access to pu numerics for unit-level pinning of the restatement                   */
double  ref_voigt_table(int nwn, double dwn, double alphaL, double alphaD, float *out){
  float *p = out;
  return (double)voigtn(nwn, dwn, alphaL, alphaD, &p, -1,
                        nwn > _voigt_maxelements ? VOIGT_QUICK : 0);
}
