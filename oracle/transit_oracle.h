/* oracle/transit_oracle.h -- TEST INFRASTRUCTURE ONLY (see transit_oracle.c header). */
#ifndef TRANSIT_ORACLE_H
#define TRANSIT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Static description of one configuration (what transit_init leaves behind). */
typedef struct {
  int nlayer, nspec, nwave;
  const double *press;      /* [nlayer] atmosphere-file units, bottom -> top            */
  double pfct, rfct;        /* unit factors to cgs                                      */
  const double *mass;       /* [nspec] g/mol                                            */
  const double *pol;        /* [nspec] polarizability (A^3), scattering flag 2 only     */
  const double *wn;         /* [nwave] cm-1                                             */
  double gsurf, p0, r0;     /* cm/s2; file pressure units; radius in rfct units         */
  /* opacity grid */
  int ntemp, ngmol;
  const double *gtemp;      /* [ntemp]                                                  */
  const int *gmol_spec;     /* [ngmol] index into species for each grid molecule        */
  const double *grid;       /* [nlayer][ntemp][ngmol][nwave]                            */
  /* CIA */
  int ncia;
  const int *cia_nwn;       /* [ncia]                                                   */
  const int *cia_nt;        /* [ncia]                                                   */
  const double *const *cia_wn;   /* [ncia][nwn]                                         */
  const double *const *cia_t;    /* [ncia][nt]                                          */
  const double *const *cia_tab;  /* [ncia][nwn*nt] row = wavenumber                     */
  const int *cia_nspec;     /* [ncia] 1 or 2                                            */
  const int *cia_spec;      /* [ncia*2] species indices                                 */
  /* ray solution */
  double toomuch;
  int nangle;
  const double *angles_deg; /* [nangle]                                                 */
  double starrad_cm;
  int transparent;
  /* per-call knobs (set_cloudtop / set_scattering state)                               */
  int cloud_flag; double cloudext, cloudtop, cloudbot;
  int scat_flag;  double scat_logext;
  int modlevel;             /* transit: 1 (modulation1) or -1 (modulationm1)            */
} orc_config;

/* Optional intermediates; any pointer may be NULL. */
typedef struct {
  double *radius;   /* [nlayer]                    */
  double *temp;     /* [nlayer] (after the identity resample) */
  double *mm;       /* [nlayer]                    */
  double *dens;     /* [nspec][nlayer]             */
  double *ext;      /* [nlayer][nwave] molecular   */
  double *cia;      /* [nwave][nlayer]             */
  double *tau;      /* [nwave][nlayer]             */
  long   *last;     /* [nwave]                     */
  double *intens;   /* [nangle][nwave] eclipse     */
} orc_inter;

int orc_forward(const orc_config *cfg, int eclipse, const double *input, double *spectrum,
                orc_inter *inter);

/* building blocks, exported for unit-level pinning against oracle/_ref */
int    orc_binsearchapprox(const double *a, double v, int lo, int hi);
void   orc_spline_init(double *z, const double *x, const double *y, long n);
double orc_splinterp_pt(const double *z, long n, const double *x, const double *y, double xo);
void   orc_splinterp(long n, const double *xi, const double *yi, long nx, const double *xo,
                     double *yo);
void   orc_radpress(double g0, double p0, double r0, const double *temp, const double *mu,
                    const double *press, double *radius, int nlayer, double rfct);
double orc_simps_path(const double *s, const double *y, int n);
double orc_eclipsetau(const double *rad, double *ex, int nlayer, int rs);
double orc_totaltau1(double b, double *rad, double *ex, long nrad);
double orc_modulationm1(const double *tau, long last, double toomuch, const double *ipv,
                        double ipfct, double srad);
double orc_modulation1(const double *tau, long last, double toomuch, const double *ipv,
                       long ipn, double ipfct, double srad, int transparent);

/* line-by-line builder (stage d) */
int  orc_voigtn(int nwn, double dwn, double alphaL, double alphaD, float *vpro, int quick);
long orc_profile_halfsize(double dwn, double dop, double lor, float ta, long nowns);

typedef struct {
  long   nlines;
  const double *wl_um;   /* [nlines] wavelength, micron                       */
  const double *elow;    /* [nlines] cm-1                                     */
  const double *gf;      /* [nlines]                                          */
  const short  *isoid;   /* [nlines]                                          */
  int    niso;
  const double *iso_mass;   /* [niso]                                         */
  const double *iso_ratio;  /* [niso]                                         */
  const int    *iso_spec;   /* [niso] species index of the isotope's molecule */
  const int    *iso_gmol;   /* [niso] output (grid molecule) index            */
  int    ngmol;
  int    nspec;
  const double *spec_mass;  /* [nspec]                                        */
  const double *spec_radius;/* [nspec] cm                                     */
  /* sampling */
  double wn_lo;    /* wns.i                                                   */
  double dwn;      /* wns.d                                                   */
  long   nwave;
  int    osamp;
  long   nowns;    /* oversampled count                                       */
  /* Voigt table */
  int nDop, nLor;
  const double *aDop, *aLor;
  const long   *profsize;      /* [nDop*nLor] half sizes                      */
  const float *const *profile; /* [nDop*nLor] pointers                        */
  double ethresh;
} orc_lbl;

int orc_computemolext(const orc_lbl *L, double temp, const double *density, const double *Z,
                      double *k /*[ngmol][nwave]*/, long *trace_iown /*[nlines] or NULL*/,
                      long *counts /*[3] nadd,nskip,neval or NULL*/);

/* permol = 0 (line-by-line forward mode): k[nwave], densities folded in */
int orc_computemolext_total(const orc_lbl *L, double temp, const double *density, const double *Z,
                            double *k /*[nwave]*/, long *counts);

/* forward model without an opacity file (tau.c:163-175,253-264): the grid fields of orc_config
   are ignored; molecular extinction comes from computemolext(permol=0) per layer */
typedef struct {
  const orc_lbl *lbl;
  const int *iso_nt;             /* [niso] temperature nodes of the isotope's database   */
  const double *const *iso_T;    /* [niso][nt]                                          */
  const double *const *iso_Z;    /* [niso][nt] partition function                       */
} orc_lbl_fwd;
int orc_forward_lbl(const orc_config *cfg, const orc_lbl_fwd *F, int eclipse, const double *input,
                    double *spectrum, orc_inter *inter);

#ifdef __cplusplus
}
#endif
#endif
