/* oracle/transit_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * called from the product (bart_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * Plain-C, single-threaded, fp64 RESTATEMENT of the reference `transit` forward-model path
 * (exosports/BART, modules/transit/{transit,pu}/src).  Every function cites the reference
 * file:line whose arithmetic it follows, including the reference's quirks (nearest-index
 * binary search, top-aligned Simpson panels, the parabolic self-interpolation of the bottom
 * sample, the >0 clamp on CIA, float32 Voigt tables).  It is deliberately written the slow,
 * literal way (O(nlayer^2) optical depth per column, spline solves per layer) so that the
 * CUDA path's algebraic short-cuts (prefix scans, pre-folded CIA splines) are checked
 * against something that does not share them.
 *
 * PINNING: validated against the compiled reference itself (oracle/_ref/libtransit_ref.so,
 * built from /root/reference by oracle/Makefile) on the configurations in tests/golden/ --
 * see tests/test_oracle_vs_reference.py and tests/golden/make_golden.py.  The reference's own
 * test-suite holds no executable known-answer tests for this path; its
 * analytic slant-path formulas (transit/test/test_slantpath.c:177-307) are used as
 * additional KATs in tests/test_oracle_analytic.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "transit_oracle.h"

/* constants: transit/include/constants_tr.h:20-46 */
#define O_PI      3.141592653589793
#define O_AMU     1.66053886e-24
#define O_LS      2.99792458e10
#define O_KB      1.380658e-16
#define O_H       6.6260755e-27
#define O_EC      4.8032068e-10
#define O_ME      9.1093897e-28
#define O_AMAGAT  2.68678e19
#define O_E0H2    4.911e-23
#define O_NAVO    6.02214076e23
#define O_MICRON  1e-4
#define O_SIGCTE  (O_PI*O_EC*O_EC/O_LS/O_LS/O_ME/O_AMU)
#define O_EXPCTE  (O_H*O_LS/O_KB)
#define O_SQRTLN2 0.83255461115769775635

/* ------------------------------------------------------------------------------------ */
/* pu/src/iomisc.c:1088-1108 -- index of the element NEAREST to v between lo and hi       */
int orc_binsearchapprox(const double *a, double v, int lo, int hi){
  while (hi - lo > 1){
    int mid = (hi + lo)/2;
    if (a[mid] > v) hi = mid; else lo = mid;
  }
  if (hi == lo) return lo;      /* the reference exits here; callers never trigger it */
  return (fabs(a[hi] - v) < fabs(a[lo] - v)) ? hi : lo;
}

/* pu/src/numerical.c:16-45 (binsearchie), used by slantpath.c:37 through `binsearch` */
static int bsearch_ie(const double *arr, long i, long f, double val){
  if (arr[i] > val) return -1;
  if (arr[f] < val) return -2;
  if (arr[f] == val) return -5;
  if (i == f && arr[i] != val) return -3;
  while (f - i > 1){
    long m = (f + i) >> 1;
    if (arr[m] > val) f = m; else i = m;
  }
  return (int)i;
}

/* ------------------------------------------------------------------------------------ */
/* Natural cubic spline: pu/src/spline.c:12-48 (tri), 186-206 (spline_init)               */
void orc_spline_init(double *z, const double *x, const double *y, long n){
  double *h = malloc(sizeof(double)*(n-1)), *b = malloc(sizeof(double)*(n-1));
  double *u = calloc(n-1, sizeof(double)),  *v = calloc(n-1, sizeof(double));
  long i;
  for (i=0; i<n-1; i++){ h[i] = x[i+1]-x[i]; b[i] = (y[i+1]-y[i])/h[i]; }
  if (n > 2){
    u[1] = 2*(h[1]+h[0]);
    v[1] = 6*(b[1]-b[0]);
    for (i=2; i<n-1; i++){
      u[i] = 2*(h[i]+h[i-1]) - h[i-1]*h[i-1]/u[i-1];
      v[i] = 6*(b[i]-b[i-1]) - v[i-1]*h[i-1]/u[i-1];
    }
  }
  z[0] = z[n-1] = 0.0;
  for (i=n-2; i>0; i--) z[i] = (v[i] - h[i]*z[i+1])/u[i];
  free(h); free(b); free(u); free(v);
}

/* pu/src/spline.c:131-183 */
double orc_splinterp_pt(const double *z, long n, const double *x, const double *y, double xo){
  int k = orc_binsearchapprox(x, xo, 0, (int)n-1);
  if (k == n-1 || xo < x[k]) k--;
  double h = x[k+1]-x[k], dy = y[k+1]-y[k];
  if (x[k] == xo) return y[k];
  if (h > 0){
    double dx = xo - x[k];
    double a = (z[k+1]-z[k])/(6*h), b = 0.5*z[k], c = dy/h - h/6*(z[k+1]+2*z[k]);
    return y[k] + dx*(c + dx*(b + dx*a));
  }
  return 0.0;
}

/* pu/src/spline.c:55-128 (spline3 + splinterp): note the different polynomial form */
void orc_splinterp(long n, const double *xi, const double *yi, long nx, const double *xo,
                   double *yo){
  double *z = malloc(sizeof(double)*n);
  long j;
  orc_spline_init(z, xi, yi, n);
  for (j=0; j<nx; j++){
    int k = orc_binsearchapprox(xi, xo[j], 0, (int)n-1);
    if (k == n-1 || xo[j] < xi[k]) k--;
    double h = xi[k+1]-xi[k], dx = xo[j]-xi[k];
    double B = (yi[k+1]-yi[k])/h - h/6*(z[k+1] + 2*z[k]);
    yo[j] = yi[k] + dx*B + dx*dx*0.5*z[k] + dx*dx*dx*(z[k+1]-z[k])/(6*h);
  }
  free(z);
}

/* ------------------------------------------------------------------------------------ */
/* transit/src/readatm.c:787-865 -- hydrostatic radii                                     */
void orc_radpress(double g0, double p0, double r0, const double *temp, const double *mu,
                  const double *press, double *radius, int nlayer, double rfct){
  int i, i0 = -1;
  double best = 1e37, t0, m0, g;
  for (i=0; i<nlayer; i++)
    if (fabs(press[i]-p0) < best){ i0 = i; best = fabs(press[i]-p0); }
  if (press[i0] > p0){
    double lr = log(press[i0+1]/press[i0]), lp = log(p0/press[i0]);
    t0 = temp[i0] + ((temp[i0+1]-temp[i0])/lr)*lp;
    m0 = mu[i0]   + ((mu[i0+1]  -mu[i0]  )/lr)*lp;
    radius[i0] = r0 + 0.5*(temp[i0]/mu[i0] + t0/m0)*(O_KB/O_AMU*lp/g0)/rfct;
  } else {
    double lr = log(press[i0-1]/press[i0]), lp = log(p0/press[i0]);
    t0 = temp[i0] + ((temp[i0-1]-temp[i0])/lr)*lp;
    m0 = mu[i0]   + ((mu[i0-1]  -mu[i0]  )/lr)*lp;
    radius[i0] = r0 - 0.5*(temp[i0]/mu[i0] + t0/m0)*(O_KB/O_AMU*log(press[i0]/p0)/g0)/rfct;
  }
  g = g0*pow(r0/radius[i0], 2);
  for (i=i0-1; i>=0; i--){
    radius[i] = radius[i+1] - 0.5*(temp[i]/mu[i] + temp[i+1]/mu[i+1])*
                (O_KB/O_AMU*log(press[i]/press[i+1])/g)/rfct;
    g = g*pow(radius[i+1]/radius[i], 2);
  }
  g = g0*pow(r0/radius[i0], 2);
  for (i=i0+1; i<nlayer; i++){
    radius[i] = radius[i-1] + 0.5*(temp[i]/mu[i] + temp[i-1]/mu[i-1])*
                (O_KB/O_AMU*log(press[i-1]/press[i])/g)/rfct;
    g = g*pow(radius[i-1]/radius[i], 2);
  }
}

/* ------------------------------------------------------------------------------------ */
/* pu/src/numerical.c:390-525 (geth, simps, simpson, makeh): Simpson on unequal spacing,
 * panels aligned to the LAST point; an even point count integrates the FIRST interval by
 * trapezoid.  `s` are abscissae, `y` ordinates.                                          */
double orc_simps_path(const double *s, const double *y, int n){
  if (n == 1) return 0.0;
  if (n == 2) return (s[1]-s[0])*(y[0]+y[1])/2;
  int even = (n % 2 == 0), np = (n-1)/2, i;
  double res = 0.0;
  for (i=0; i<np; i++){
    int j = 2*i + even;
    double h0 = s[j+1]-s[j], h1 = s[j+2]-s[j+1];
    double hsum = h0+h1, hratio = h1/h0, hfactor = hsum*hsum/(h0*h1);
    res += (y[j]*(2.0-hratio) + y[j+1]*hfactor + y[j+2]*(2.0-1.0/hratio))*hsum;
  }
  res /= 6.0;
  if (even) res += (s[1]-s[0])*(y[0]+y[1])/2;
  return res;
}

/* pu/src/numerical.c:182-195 */
static double interp_parab(const double *x, const double *y, double xr){
  const double dx = x[1]-x[0];
  const double x0 = x[0]/dx;
  const double my = y[0]+y[2]-2*y[1];
  const double a  = my/(2.0*dx*dx);
  const double b  = (y[2]-y[1]-(x0+1.5)*my)/dx;
  const double c  = y[0] + x0*(y[2]-4*y[1]+3*y[0]+x0*my)/2.0;
  return xr*xr*a + xr*b + c;
}

/* transit/src/eclipse.c:28-105 -- vertical optical depth from layer rs to the top, / rfct.
 * `ex` is MODIFIED exactly as the reference modifies it (bottom sample replaced by its own
 * parabolic interpolant and not restored when more than two points are used).            */
double orc_eclipsetau(const double *rad_all, double *ex_all, int nlayer, int rs){
  if (rs == nlayer-1) return 0.0;
  const double *rad = rad_all + rs;
  double *ex = ex_all + rs;
  int n = nlayer - rs, i;
  const double keep = ex[0];
  double r3[3], x3[3];
  if (n == 2) ex[0] = interp_parab(rad-1, ex-1, rad[0]);
  else        ex[0] = interp_parab(rad,   ex,   rad[0]);
  const double *yy = ex, *rr = rad;
  if (n == 2){
    x3[0] = ex[0]; x3[2] = ex[1]; x3[1] = (ex[1]+ex[0])/2.0;
    r3[0] = rad[0]; r3[2] = rad[1]; r3[1] = (rad[0]+rad[1])/2.0;
    ex[0] = keep;
    yy = x3; rr = r3; n = 3;
  }
  double *s = malloc(sizeof(double)*n);
  s[0] = 0.0;
  for (i=1; i<n; i++) s[i] = s[i-1] + (rr[i]-rr[i-1]);
  double res = orc_simps_path(s, yy, n);
  free(s);
  return res;
}

/* transit/src/slantpath.c:18-108 -- chord optical depth at impact parameter b, / rfct.    */
double orc_totaltau1(double b, double *rad_all, double *ex_all, long nrad){
  double r0 = b;                       /* refraction index is 1 (idxrefraction.c:30-55) */
  int rs = bsearch_ie(rad_all, 0, nrad-1, r0), i;
  if (rs == -5 || rs == -2) return 0.0;
  if (rs < 0){ fprintf(stderr, "orc_totaltau1: b outside the atmosphere\n"); exit(1); }
  double *rad = rad_all + rs, *ex = ex_all + rs;
  long n = nrad - rs;
  const double keep_ex = ex[0], keep_rad = rad[0];
  double r3[3], x3[3];
  if (n == 2) ex[0] = interp_parab(rad-1, ex-1, r0);
  else        ex[0] = interp_parab(rad,   ex,   r0);
  rad[0] = r0;
  const double *yy = ex, *rr = rad;
  if (n == 2){
    x3[0] = ex[0]; x3[2] = ex[1]; x3[1] = (ex[0]+ex[1])/2.0;
    r3[0] = rad[0]; r3[2] = rad[1]; r3[1] = (rad[0]+rad[1])/2.0;
    rad[0] = keep_rad; ex[0] = keep_ex;
    yy = x3; rr = r3; n = 3;
  }
  double *s = malloc(sizeof(double)*n);
  s[0] = 0.0;
  for (i=1; i<n; i++) s[i] = sqrt(rr[i]*rr[i] - r0*r0);
  double res = orc_simps_path(s, yy, (int)n);
  free(s);
  ex[0] = keep_ex; rad[0] = keep_rad;
  return 2*res;
}

/* transit/src/slantpath.c:446-473 + pu/src/numerical.c:203-211 (interp_line): modlevel -1, the
   radius where tau = toomuch by linear interpolation; -1 when toomuch was not reached (the
   reference then exits, slantpath.c:308-316) */
double orc_modulationm1(const double *tau, long last, double toomuch, const double *ipv_in,
                        double ipfct, double srad){
  long i, ini;
  double ipv[2] = {0.0, 0.0};
  if (tau[last] < toomuch) return -1;
  ini = ++last - 2;
  if (ini < 0) ini = 0;
  for (i=ini; i<last; i++) ipv[i-ini] = ipv_in[i]*ipfct;
  {
    const double *x = tau + ini;
    const double dx = x[1] - x[0];
    const double m = (ipv[1] - ipv[0]) / dx;
    const double muchrad = ipv[0] + (toomuch - x[0]) * m;
    return muchrad * muchrad / (srad*srad);
  }
}

/* transit/src/slantpath.c:350-436 */
double orc_modulation1(const double *tau, long last, double toomuch, const double *ipv_in,
                       long ipn, double ipfct, double srad, int transparent){
  long ipn1 = ipn-1, i;
  const double maxtau = tau[last] > toomuch ? tau[last] : toomuch;
  double *rinteg = calloc(ipn, sizeof(double)), *ipv = calloc(ipn, sizeof(double));
  for (i=0; i<=last; i++){
    ipv[ipn1-i] = ipv_in[i]*ipfct;
    rinteg[ipn1-i] = exp(-tau[i])*ipv[ipn1-i];
  }
  last += 1;
  if (last > ipn1) last = ipn1;
  for (; i<=last; i++){
    ipv[ipn1-i] = ipv_in[i]*ipfct;
    rinteg[ipn1-i] = 0;
  }
  last++;
  if (last < 3){ fprintf(stderr, "orc_modulation1: fewer than 3 points\n"); exit(1); }
  double res = orc_simps_path(ipv+ipn-last, rinteg+ipn-last, (int)last);
  res = ipv[ipn1]*ipv[ipn1] - 2.0*res;
  if (transparent) res -= exp(-maxtau)*ipv[ipn-last]*ipv[ipn-last];
  res *= 1.0/(srad*srad);
  free(rinteg); free(ipv);
  return res;
}

/* ------------------------------------------------------------------------------------ */
/* transit/src/crosssec.c:353-428 (bicubicinterpolate) + 271-344 (interpcs)               */
static void cia_file(const orc_config *c, int f, const double *temp, const double *dens,
                     double *e_cs /*[nwave][nlayer], accumulated*/){
  int nx1 = c->cia_nwn[f], nx2 = c->cia_nt[f], nl = c->nlayer, nw = c->nwave, i, j, k;
  const double *x1 = c->cia_wn[f], *x2 = c->cia_t[f], *src = c->cia_tab[f];
  double *res = calloc((size_t)nw*nl, sizeof(double));
  double *f2 = malloc(sizeof(double)*(size_t)nl*nx1);
  double *z1 = malloc(sizeof(double)*nx2), *z2 = malloc(sizeof(double)*nx1);
  int ok = !(c->wn[0] > x1[nx1-1] || c->wn[nw-1] < x1[0] || temp[0] > x2[nx2-1]
             || temp[nl-1] < x2[0]);
  /* note: the reference's range test uses t2[0] and t2[nt2-1] only (crosssec.c:375);
     interpcs has already exited if any layer is out of range (293-309)                  */
  if (ok){
    int fi = 0, li = nw, fj = 0, lj = nl;
    while (c->wn[fi] < x1[0]) fi++;
    for (i=0; i<li; i++) if (c->wn[i] > x1[nx1-1]) li = i;
    while (temp[fj] < x2[0]) fj++;
    for (j=0; j<lj; j++) if (temp[j] > x2[nx2-1]) lj = j;
    for (i=0; i<nx1; i++){
      orc_spline_init(z1, x2, src + (size_t)i*nx2, nx2);
      for (j=fj; j<lj; j++)
        f2[(size_t)j*nx1+i] = orc_splinterp_pt(z1, nx2, x2, src + (size_t)i*nx2, temp[j]);
    }
    for (j=fj; j<lj; j++){
      orc_spline_init(z2, x1, f2 + (size_t)j*nx1, nx1);
      for (i=fi; i<li; i++)
        res[(size_t)i*nl+j] += orc_splinterp_pt(z2, nx1, x1, f2 + (size_t)j*nx1, c->wn[i]);
    }
  }
  for (j=0; j<nl; j++){
    double d = 1.0;
    for (k=0; k<c->cia_nspec[f]; k++){
      int s = c->cia_spec[2*f+k];
      d *= dens[(size_t)s*nl+j]/(O_AMU*c->mass[s]*O_AMAGAT);
    }
    for (i=0; i<nw; i++)
      if (res[(size_t)i*nl+j] > 0) e_cs[(size_t)i*nl+j] += res[(size_t)i*nl+j]*d;
  }
  free(res); free(f2); free(z1); free(z2);
}

/* ------------------------------------------------------------------------------------ */
/* transit/src/extinction.c:586-624 */
static void ext_scat(const orc_config *c, const double *press, const double *temp,
                     const double *dens, double wn, double *e){
  int i, j, n = c->nlayer;
  switch (c->scat_flag){
  case 1:
    for (i=0; i<n; i++)
      e[i] = pow(10.0, c->scat_logext)*O_E0H2*press[i]/temp[i]*pow(wn, 4);
    break;
  case 2:
    for (i=0; i<n; i++){
      e[i] = 0.0;
      for (j=0; j<c->nspec; j++)
        e[i] += O_PI*8e-32/3.*pow(c->pol[j], 2)*pow(2.*O_PI*wn*O_MICRON, 4)*
                dens[(size_t)j*n+i]/c->mass[j]*O_NAVO;
    }
    break;
  default:
    for (i=0; i<n; i++) e[i] = 0.0;
  }
}

/* transit/src/extinction.c:629-693, flag 1 (constant extinction) only: flags 2-5 read the
 * uninitialised `mean_dens` VLA of tau.c:127-131,203 in the reference and are therefore not
 * reproducible.                                                                          */
static void ext_cloud(const orc_config *c, const double *press, double *e){
  int i, n = c->nlayer;
  double top = pow(10, c->cloudtop), bot = pow(10, c->cloudbot);
  if (!c->cloudext || c->cloud_flag == 0){
    /* cloudext==0 zeroes (extinction.c:652-655); flag 0 with ext!=0 leaves e untouched in
       the switch, i.e. zero here because the caller's array starts from the previous wn  */
    for (i=0; i<n; i++) e[i] = 0.0;
    return;
  }
  for (i=n-1; i>=0; i--){
    if (press[i] >= top) break;
    e[i] = 0.0;
  }
  for (; i>=0; i--){
    if (press[i] >= bot) break;
    e[i] = c->cloudext;
  }
  for (; i>=0; i--) e[i] = 0.0;
}

/* ------------------------------------------------------------------------------------ */
static int computemolext_impl(const orc_lbl *L, double temp, const double *density, const double *Z,
                              double *k, long *trace_iown, long *counts, int permol);

static int forward_impl(const orc_config *c, const orc_lbl_fwd *F, int eclipse, const double *input,
                        double *spectrum, orc_inter *inter){
  const int nl = c->nlayer, ns = c->nspec, nw = c->nwave;
  int i, j, w, m;
  double *t_in = malloc(sizeof(double)*nl), *mm_in = malloc(sizeof(double)*nl);
  double *d_in = malloc(sizeof(double)*(size_t)ns*nl);
  double *rad = malloc(sizeof(double)*nl);
  double *temp = malloc(sizeof(double)*nl), *press = malloc(sizeof(double)*nl);
  double *mm = malloc(sizeof(double)*nl), *dens = malloc(sizeof(double)*(size_t)ns*nl);

  /* reloadatm: readatm.c:722-784.  input = [T | q_0 | q_1 ...], each nlayer long          */
  for (i=0; i<nl; i++){
    t_in[i] = input[i];
    double mu = 0.0;                                       /* checkaddmm, number abundances */
    for (j=0; j<ns; j++) mu += input[(size_t)nl*(j+1)+i]*c->mass[j];
    mm_in[i] = mu;
    for (j=0; j<ns; j++){                                  /* stateeqnford transit.h:58-69  */
      double rho = O_AMU*input[(size_t)nl*(j+1)+i]*(c->press[i]*c->pfct)/O_KB/t_in[i];
      d_in[(size_t)j*nl+i] = rho*c->mass[j];
    }
  }
  orc_radpress(c->gsurf, c->p0, c->r0, t_in, mm_in, c->press, rad, nl, c->rfct);

  /* makeradsample with raddelt -1: makesample.c:472-531 -- spline "resample" onto the same
     radii (identity up to rounding at the last knot)                                     */
  orc_splinterp(nl, rad, t_in, nl, rad, temp);
  orc_splinterp(nl, rad, c->press, nl, rad, press);
  orc_splinterp(nl, rad, mm_in, nl, rad, mm);
  for (j=0; j<ns; j++)
    orc_splinterp(nl, rad, d_in+(size_t)j*nl, nl, rad, dens+(size_t)j*nl);

  /* interpcs: crosssec.c:271-344 */
  double *e_cs = calloc((size_t)nw*nl, sizeof(double));
  for (m=0; m<c->ncia; m++) cia_file(c, m, temp, dens, e_cs);

  /* molecular extinction for every layer: extinction.c:534-581                            */
  double *e = calloc((size_t)nl*nw, sizeof(double));
  if (F){
    /* no opacity file: tau.c:163-175,253-264 computes each layer line by line at the layer's own
       temperature, with Z_iso(T) splined from the TLI tables (makesample.c:534-544)         */
    const orc_lbl *L = F->lbl;
    double *dl = malloc(sizeof(double)*ns), *Zl = malloc(sizeof(double)*L->niso);
    for (i=0; i<nl; i++){
      for (j=0; j<ns; j++) dl[j] = dens[(size_t)j*nl+i];
      for (j=0; j<L->niso; j++)
        orc_splinterp(F->iso_nt[j], F->iso_T[j], F->iso_Z[j], 1, temp+i, Zl+j);
      computemolext_impl(L, temp[i], dl, Zl, e+(size_t)i*nw, NULL, NULL, 0);
    }
    free(dl); free(Zl);
  }
  for (i=0; i<nl && !F; i++){
    double T = temp[i];
    int it = orc_binsearchapprox(c->gtemp, T, 0, c->ntemp-1);
    /* the reference passes hi = Ntemp (one past the end); identical for T < gtemp[Ntemp-1] */
    if (T < c->gtemp[it]) it--;
    if (it > c->ntemp-2) it = c->ntemp-2;
    double t0 = c->gtemp[it], t1 = c->gtemp[it+1];
    for (w=0; w<nw; w++)
      for (m=0; m<c->ngmol; m++){
        const double *lo = c->grid + (((size_t)i*c->ntemp + it  )*c->ngmol + m)*nw;
        const double *hi = c->grid + (((size_t)i*c->ntemp + it+1)*c->ngmol + m)*nw;
        double ext = (lo[w]*(t1-T) + hi[w]*(T-t0))/(t1-t0);
        e[(size_t)i*nw+w] += dens[(size_t)c->gmol_spec[m]*nl+i]*ext;
      }
  }

  /* tau: tau.c:216-305 */
  double *tau = calloc((size_t)nw*nl, sizeof(double));
  long *last = calloc(nw, sizeof(long));
  double *er = malloc(sizeof(double)*nl), *e_s = malloc(sizeof(double)*nl);
  double *e_c = calloc(nl, sizeof(double));
  double *radw = malloc(sizeof(double)*nl);
  for (w=0; w<nw; w++){
    double *tw = tau + (size_t)w*nl;
    long ri;
    ext_scat(c, press, temp, dens, c->wn[w], e_s);
    ext_cloud(c, press, e_c);
    for (i=0; i<nl; i++) er[i] = e[(size_t)i*nw+w] + e_s[i] + e_c[i] + e_cs[(size_t)w*nl+i];
    last[w] = nl-1;
    for (ri=0; ri<nl; ri++){
      int rs = nl-1-(int)ri;           /* h[ri] = r[nl-1-ri]; binsearchapprox finds it exactly */
      if (eclipse) tw[ri] = c->rfct*orc_eclipsetau(rad, er, nl, rs);
      else {
        memcpy(radw, rad, sizeof(double)*nl);
        tw[ri] = c->rfct*orc_totaltau1(rad[rs], radw, er, nl);
      }
      if (tw[ri] > c->toomuch){ last[w] = ri; break; }
    }
  }

  if (eclipse){
    /* eclipse.c:117-160 (eclipse_intens), 242-287 (flux)                                  */
    int na = c->nangle, a;
    double *B = malloc(sizeof(double)*nl), *dt = malloc(sizeof(double)*nl);
    double *area = malloc(sizeof(double)*(na+1));
    area[0] = 0.0*(O_PI/180.0); area[na] = 90.0*(O_PI/180.0);
    for (a=1; a<na; a++) area[a] = (c->angles_deg[a-1]+c->angles_deg[a])*(O_PI/180.0)/2.0;
    for (w=0; w<nw; w++) spectrum[w] = 0.0;
    for (a=0; a<na; a++){
      double ang = c->angles_deg[a]*(O_PI/180.0);
      double wgt = pow(sin(area[a+1]), 2.0) - pow(sin(area[a]), 2.0);
      for (w=0; w<nw; w++){
        const double *tw = tau + (size_t)w*nl;
        long L = last[w], k;
        double wv = c->wn[w];
        for (k=0; k<=L; k++){
          dt[k] = exp(-tw[k]/cos(ang));
          B[k] = (2.0*O_H*pow(wv, 3.0)*O_LS*O_LS)/(exp(O_H*wv*O_LS/(O_KB*temp[nl-1-k])) - 1.0);
        }
        double trap = 0.0;                            /* numerical.c:154-172 */
        for (k=0; k<L; k++) trap += (dt[k+1]-dt[k])*(B[k+1]+B[k]);
        double I = B[L]*dt[L] - 0.5*trap;
        if (inter && inter->intens) inter->intens[(size_t)a*nw+w] = I;
        spectrum[w] += O_PI*I*wgt;
      }
    }
    free(B); free(dt); free(area);
  } else {
    double *ipv = malloc(sizeof(double)*nl);
    for (i=0; i<nl; i++) ipv[i] = rad[nl-1-i];          /* makesample.c:564-574 */
    for (w=0; w<nw; w++)
      spectrum[w] = c->modlevel == -1 ?
        orc_modulationm1(tau+(size_t)w*nl, last[w], c->toomuch, ipv, c->rfct, c->starrad_cm) :
        orc_modulation1(tau+(size_t)w*nl, last[w], c->toomuch, ipv, nl, c->rfct,
                        c->starrad_cm, c->transparent);
    free(ipv);
  }

  if (inter){
    if (inter->radius) memcpy(inter->radius, rad, sizeof(double)*nl);
    if (inter->temp)   memcpy(inter->temp, temp, sizeof(double)*nl);
    if (inter->mm)     memcpy(inter->mm, mm, sizeof(double)*nl);
    if (inter->dens)   memcpy(inter->dens, dens, sizeof(double)*(size_t)ns*nl);
    if (inter->ext)    memcpy(inter->ext, e, sizeof(double)*(size_t)nl*nw);
    if (inter->cia)    memcpy(inter->cia, e_cs, sizeof(double)*(size_t)nl*nw);
    if (inter->tau)    memcpy(inter->tau, tau, sizeof(double)*(size_t)nl*nw);
    if (inter->last)   memcpy(inter->last, last, sizeof(long)*nw);
  }
  free(t_in); free(mm_in); free(d_in); free(rad); free(temp); free(press); free(mm);
  free(dens); free(e_cs); free(e); free(tau); free(last); free(er); free(e_s); free(e_c);
  free(radw);
  return 0;
}

int orc_forward(const orc_config *c, int eclipse, const double *input, double *spectrum,
                orc_inter *inter){
  return forward_impl(c, NULL, eclipse, input, spectrum, inter);
}

int orc_forward_lbl(const orc_config *c, const orc_lbl_fwd *F, int eclipse, const double *input,
                    double *spectrum, orc_inter *inter){
  return forward_impl(c, F, eclipse, input, spectrum, inter);
}

/* ====================================================================================== */
/* Line-by-line builder (stage d)                                                         */

/* Voigt function of pu/src/voigt.c:132-200, K(x, y) * sqrt(ln2/pi) / alphaD, as float.
   Region I (x < 3, y < 1.8): w(z) = exp(-z^2) (1 + (2i/sqrt(pi)) int_0^z exp(t^2) dt) with the
   integral as its Maclaurin series sum_k z^(2k+1) / (k! (2k+1)), carried here as the running power
   p_k = -i z^(2k+1) (long double, like the reference).  Regions II / III: Pierluissi's sums over
   poles, Re sum_j c_j / (z^2 - s_j) written out in real arithmetic.  The order of the floating-point
   operations follows the reference's expressions.                                            */
typedef struct { long double re, im; } cxl;
static float voigtxy(double x, double y, double alphaD){
  static const double pole2[3][2] = {{0.46131350, 0.19016350}, {0.09999216, 1.78449270},
                                     {0.002883894, 5.52534370}};      /* {weight, shift}, region II  */
  static const double pole3[2][2] = {{0.51242424, 0.27525510}, {0.05176536, 2.72474500}};
  const double norm = 0.46971863934982566689 / alphaD;                /* sqrt(ln 2 / pi) / alphaD    */
  const double two_over_sqrtpi = 1.12837916709551257389;
  const cxl zsq = {(long double)(x*x - y*y), (long double)(2*x*y)};   /* z^2, z = x + i y            */
  if (x < 3 && y < 1.8){
    const int nterms = (x < 1 ? 15 : (int)(6.842*x + 8.0)) + 1;
    cxl pw = {y, -x};                                                 /* -i z                        */
    cxl sum = pw;
    long double kfact = 1.0L;
    int k;
    for (k = 1; k <= nterms; k++){
      const cxl nx = {pw.re*zsq.re - pw.im*zsq.im, pw.re*zsq.im + pw.im*zsq.re};   /* pw * z^2      */
      kfact *= k;
      const long double coef = 1.0L/(kfact*(2*k + 1));
      sum.im += nx.im*coef;
      sum.re += nx.re*coef;
      pw = nx;
    }
    return (float)(norm*exp(-zsq.re)*(cosl(zsq.im)*(1 - sum.re*two_over_sqrtpi) -
                                      sinl(zsq.im)*sum.im*two_over_sqrtpi));
  }
  {
    const long double im2 = zsq.im*zsq.im, xim = zsq.im*x;
    const int np = (x < 5 && y < 5) ? 3 : 2;
    const double (*pole)[2] = np == 3 ? pole2 : pole3;
    long double acc = 0.0L;
    int j;
    for (j = 0; j < np; j++){
      const long double d = zsq.re - pole[j][1];
      const long double term = pole[j][0]*((xim - d*y)/(d*d + im2));
      acc = j == 0 ? term : acc + term;
    }
    return (float)(norm*acc);
  }
}

/* pu/src/voigt.c:369-483 (voigtn), 489-554 (meanintegSimp / meanintegTrap, float math)    */
int orc_voigtn(int nwn, double dwn, double alphaL, double alphaD, float *vpro, int quick){
  double y = O_SQRTLN2*alphaL/alphaD;
  double ddwn = 2.0*dwn/(nwn-1);
  int nint = 50, i, o;
  double dint = alphaD/(nint-1);
  if (ddwn < dint || quick){ dint = ddwn; nint = nwn+1; }
  else {
    nint = (int)(ddwn/dint) + 1;
    if (nint & 1) nint++;
    nint = nwn*nint + 1;
    dint = 2.0*dwn/(nint-1);
  }
  float *a = calloc(nint, sizeof(float));
  if (!a) return -1;
  for (i=0; i<nint; i++)
    a[i] = voigtxy(O_SQRTLN2*fabs(dint*i - dwn)/alphaD, y, alphaD);
  if (quick){
    for (i=0; i<nwn; i++) vpro[i] = a[i];
  } else {
    int ipo = (nint-1)/nwn;             /* fine intervals per output bin */
    const float *in = a;
    if ((ipo+1) & 1){                   /* odd point count: Simpson */
      for (o=0; o<nwn; o++, in+=ipo){
        float acc = 0;
        for (i=1; i<ipo; i+=2) acc += in[i];
        acc *= 2;
        for (i=2; i<ipo; i+=2) acc += in[i];
        acc *= 2;
        acc += in[0] + in[ipo];
        acc /= (ipo*3.0);
        vpro[o] = acc;
      }
    } else {
      for (o=0; o<nwn; o++, in+=ipo){
        float acc = 0;
        for (i=1; i<ipo; i++) acc += in[i];
        acc = (acc + (in[0]+in[ipo])/2.0)/(double)ipo;
        vpro[o] = acc;
      }
    }
  }
  free(a);
  return 1;
}

/* transit/src/extinction.c:8-57 (getprofile): half-size of the profile                    */
long orc_profile_halfsize(double dwn, double dop, double lor, float ta, long nowns){
  double big = dop; if (big < lor) big = lor;
  double wvgt = big*ta;
  int nvgt = 2*(long)(wvgt/dwn + 0.5) + 1;
  if (nvgt < 2) nvgt = 3;
  if (nvgt > 2*nowns) nvgt = 2*(int)nowns + 1;
  return nvgt/2;
}

/* transit/src/extinction.c:281-529.  permol = 1 is the opacity-grid call (opacity.c:397):
   one row per line-list molecule, no density factor.  permol = 0 is the line-by-line forward
   call (tau.c:163-175,253-264): the species index m stays 0 for every isotope (extinction.c:
   296,405-406,433-434), so there is ONE strongest-line reference across all molecules, every
   profile lands in row 0, and each strength is multiplied by its species' density (472-473). */
static int computemolext_impl(const orc_lbl *L, double temp, const double *density, const double *Z,
                              double *k, long *trace_iown, long *counts, int permol){
  int niso = L->niso, i, j;
  long ln, nl = L->nlines, nwn = L->nwave;
  double dwn = L->dwn, odwn = L->dwn/L->osamp;
  double *alphal = calloc(niso, sizeof(double)), *alphad = calloc(niso, sizeof(double));
  int *idop = calloc(niso, sizeof(int)), *ilor = calloc(niso, sizeof(int));
  int nout = permol ? L->ngmol : 1;
  double *kmax = calloc(nout, sizeof(double));
  long nadd = 0, nskip = 0, neval = 0;
  double own_last = L->wn_lo + (L->nowns-1)*odwn;   /* owns.v[onwn-1], makesample.c:97-104 */
  double wn0 = L->wn_lo;                            /* wn[0] */
  memset(k, 0, sizeof(double)*(size_t)nout*nwn);

  double fdoppler = sqrt(2*O_KB*temp/O_AMU)*O_SQRTLN2/O_LS;
  double florentz = sqrt(2*O_KB*temp/O_PI/O_AMU)/(O_AMU*O_LS);
  for (i=0; i<niso; i++){
    alphal[i] = 0.0;
    for (j=0; j<L->nspec; j++){
      double cs = L->spec_radius[j] + L->spec_radius[L->iso_spec[i]];
      alphal[i] += density[j]/L->spec_mass[j]*cs*cs*sqrt(1/L->iso_mass[i] + 1/L->spec_mass[j]);
    }
    alphal[i] *= florentz;
    alphad[i] = fdoppler/sqrt(L->iso_mass[i]);
    /* the reference passes hi = nDop / nLor (one past the end); clamp like the CUDA path */
    idop[i] = orc_binsearchapprox(L->aDop, alphad[i]*wn0, 0, L->nDop-1);
    ilor[i] = orc_binsearchapprox(L->aLor, alphal[i],     0, L->nLor-1);
  }
  /* pass 1: strongest line per output molecule (400-427) */
  for (ln=0; ln<nl; ln++){
    double wavn = 1.0/(L->wl_um[ln]*1e-4);
    i = L->isoid[ln];
    int m = permol ? L->iso_gmol[i] : 0;
    if (wavn < L->wn_lo || wavn > own_last) continue;
    double pk = L->iso_ratio[i]*O_SIGCTE*L->gf[ln]*exp(-O_EXPCTE*L->elow[ln]/temp)*
                (1-exp(-O_EXPCTE*wavn/temp))/L->iso_mass[i]/Z[i];
    if (kmax[m] == 0) kmax[m] = pk; else kmax[m] = fmax(kmax[m], pk);
  }
  /* pass 2 (430-511) */
  for (ln=0; ln<nl; ln++){
    double wavn = 1.0/(L->wl_um[ln]*1e-4);
    i = L->isoid[ln];
    int m = permol ? L->iso_gmol[i] : 0;
    if (trace_iown) trace_iown[ln] = -1;
    if (wavn < L->wn_lo || wavn > own_last) continue;
    double pk = L->gf[ln]*exp(-O_EXPCTE*L->elow[ln]/temp)*(1-exp(-O_EXPCTE*wavn/temp));
    int iown = (int)((wavn - L->wn_lo)/odwn);
    double v0 = L->wn_lo + iown*odwn, v1 = L->wn_lo + (iown+1)*odwn;
    if (fabs(wavn - v1) < fabs(wavn - v0)) iown++;
    if (trace_iown) trace_iown[ln] = iown;
    double vnode = L->wn_lo + iown*odwn;
    while (ln != nl-1 && L->isoid[ln+1] == i){
      double nxt = 1.0/(L->wl_um[ln+1]*1e-4);
      if (fabs(nxt - vnode) < odwn){
        nadd++; ln++;
        if (trace_iown) trace_iown[ln] = -2 - iown;      /* co-added into iown */
        pk += L->gf[ln]*exp(-O_EXPCTE*L->elow[ln]/temp)*(1-exp(-O_EXPCTE*nxt/temp));
      } else break;
    }
    pk *= O_SIGCTE*L->iso_ratio[i]/(L->iso_mass[i]*Z[i]);
    if (pk < L->ethresh*kmax[m]){ nskip++; continue; }
    if (!permol) pk *= density[L->iso_spec[i]];
    int idwn = (int)((wavn - L->wn_lo)/dwn);
    if (alphad[i]*wavn/alphal[i] >= 1e-1)
      idop[i] = orc_binsearchapprox(L->aDop, alphad[i]*wavn, 0, L->nDop-1);
    long ps = L->profsize[(size_t)idop[i]*L->nLor + ilor[i]];
    const float *prof = L->profile[(size_t)idop[i]*L->nLor + ilor[i]];
    long subw = iown - (long)idwn*L->osamp;
    long offset = iown - ps;
    long minj = idwn - (ps - subw)/L->osamp;
    long maxj = idwn + (ps + subw)/L->osamp;
    if (minj < 0) minj = 0;
    if (maxj >= nwn) maxj = nwn-1;
    long bj = (long)L->osamp*minj - offset, jj;
    for (jj=minj; jj<=maxj; jj++){
      if (bj > 2*ps) break;
      if (bj >= 0) k[(size_t)m*nwn + jj] += pk*prof[bj];
      bj += L->osamp;
    }
    neval++;
  }
  if (counts){ counts[0] = nadd; counts[1] = nskip; counts[2] = neval; }

  free(alphal); free(alphad); free(idop); free(ilor); free(kmax);
  return 0;
}

int orc_computemolext(const orc_lbl *L, double temp, const double *density, const double *Z,
                      double *k, long *trace_iown, long *counts){
  return computemolext_impl(L, temp, density, Z, k, trace_iown, counts, 1);
}

int orc_computemolext_total(const orc_lbl *L, double temp, const double *density, const double *Z,
                            double *k, long *counts){
  return computemolext_impl(L, temp, density, Z, k, NULL, counts, 0);
}
