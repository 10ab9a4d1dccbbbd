// column_math.cuh -- the arithmetic of the forward model, written once as inline
// host+device functions.  The CUDA kernels in kernels.cu are thin wrappers (thread mapping,
// shared-memory staging) around these; tests/cpu_emu compiles the same functions for the host
// as a TEST-ONLY aid to debug the math without a GPU (the product never runs them on the CPU).
//
// Reference citations are into exosports/BART, modules/transit/{transit,pu}/src.
#pragma once
#include "device.cuh"
#include <cmath>
#include <cstring>

namespace bart {

constexpr double cPI = 3.141592653589793;
constexpr double cAMU = 1.66053886e-24;
constexpr double cLS = 2.99792458e10;
constexpr double cKB = 1.380658e-16;
constexpr double cH = 6.6260755e-27;
constexpr double cAMAGAT = 2.68678e19;
constexpr double cE0H2 = 4.911e-23;
constexpr double cNAVO = 6.02214076e23;
constexpr double cMICRON = 1e-4;

enum { REJ_TGRID = 1, REJ_TCIA = 2, REJ_SUMQ = 4, REJ_FEWPTS = 8 };

// ---------------------------------------------------------------------------------------
// fp64 exp and reciprocal for the column kernels.  libdevice's exp() costs ~45 issue slots per
// call on sm_100 (half of them integer moves that materialise the polynomial coefficients),
// and the eclipse column needs six per (layer, wavenumber) cell.  This version keeps the
// coefficients as immediate/constant-bank operands of the FMAs, uses a Cody-Waite reduction and
// builds the 2^n scaling in the exponent field.
BART_HD double bits_to_double(long long b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d; memcpy(&d, &b, 8); return d;
#endif
}
BART_HD long long double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  long long b; memcpy(&b, &d, 8); return b;
#endif
}

// 16-byte pair load from a table record (records and pair offsets are 16-byte aligned)
struct alignas(16) D2 { double x, y; };
BART_HD D2 ld2(const double *p) { return *reinterpret_cast<const D2 *>(p); }

// Table-driven variant used by the column kernels: exp(x) = 2^n * 2^(j/16) * e^r with
// m = round(16 x / ln2) = 16 n + j and |r| <= ln2/32, so a degree-6 polynomial suffices
// (truncation 4e-16) -- 11 fp64 instructions instead of 19.  `tab16` holds 2^(j/16), j = 0..15,
// in shared memory: 128 bytes = one full bank row, so divergent lanes never conflict.
// GUARD_LO: x may be below -708 (result 0); GUARD_HI: x may exceed 709 (result saturates).  The
// guards are selects on the result (cheaper than clamping the argument); out-of-range arguments
// only ever produce a discarded product.
constexpr int kExpTabSize = 16;
BART_HD void fill_exp_table(double *tab16) {
  // 2^(j/16) correctly rounded
  const double v[kExpTabSize] = {
      1.0, 1.0442737824274138403, 1.0905077326652576592, 1.1387886347566916537,
      1.1892071150027210667, 1.2418578120734840486, 1.2968395546510096659, 1.3542555469368927283,
      1.4142135623730950488, 1.4768261459394993114, 1.5422108254079408236, 1.6104903319492543082,
      1.6817928305074290861, 1.7562521603732994831, 1.8340080864093424635, 1.9152065613971472939};
  for (int j = 0; j < kExpTabSize; j++) tab16[j] = v[j];
}

template <bool GUARD_LO, bool GUARD_HI>
BART_HD double fast_exp_t(double x, const double *tab16) {
  const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52
  const double t = fma(x, 23.083120654223414518, MAGIC);          // low word = round(16 x / ln2)
  const double kf = t - MAGIC;
  double r = fma(kf, -4.33216987730702385306e-02, x);             // (ln2 high part) / 16, exact product
  r = fma(kf, -1.19263433079411731251e-11, r);                    // (ln2 low part) / 16
  double p = 1.38888888888888888889e-03;                          // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);
  p = fma(p, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int m = (int)double_to_bits(t);                           // low word of t
  const long long sc = double_to_bits(tab16[m & (kExpTabSize - 1)]) + ((long long)(m >> 4) << 52);
  double res = p * bits_to_double(sc);
  if (GUARD_LO) res = x < -708.0 ? 0.0 : res;
  if (GUARD_HI) res = x > 709.0 ? 1.0e308 : res;
  return res;
}
BART_HD double fast_exp(double x, const double *t) { return fast_exp_t<true, true>(x, t); }
BART_HD double fast_exp_neg(double x, const double *t) { return fast_exp_t<true, false>(x, t); }  // x <= 0
BART_HD double fast_exp_pos(double x, const double *t) { return fast_exp_t<false, true>(x, t); }  // x >= 0

// 1/d for finite d > 0: hardware seed (>= 20 bits) + three Newton steps (full fp64 precision); the compiler's IEEE divide
// is ~3x the issue slots because of its special-case branches.
BART_HD double fast_rcp(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / d;
#endif
}

struct KnobVals {
  double r0;
  int cloud_flag; double cloudext, cloudtop, cloudbot;
  int scat_flag; double scat_logext;
};

BART_HD KnobVals knobs_for(const Knobs &k, int m) {
  KnobVals v;
  v.r0 = k.r0 ? k.r0[m] : k.r0_all;
  if (k.cloudtop) {                       // set_cloudtop semantics, transit.c:103-109
    v.cloud_flag = 1; v.cloudext = 100.0; v.cloudtop = k.cloudtop[m]; v.cloudbot = k.cloudtop[m] + 10;
  } else {
    v.cloud_flag = k.cloud_flag_all; v.cloudext = k.cloudext_all;
    v.cloudtop = k.cloudtop_all; v.cloudbot = k.cloudbot_all;
  }
  v.scat_flag = k.scat_flag ? k.scat_flag[m] : k.scat_flag_all;
  v.scat_logext = k.scat_logext ? k.scat_logext[m] : k.scat_logext_all;
  return v;
}

// ---------------------------------------------------------------------------------------
// atm_prep stage 1, one layer: mean molecular mass and mass densities.
// Reference: checkaddmm readatm.c:122-159 (number abundances), stateeqnford transit.h:58-69,
// reloadatm readatm.c:722-784.  rho[j*rho_stride] receives the density of species j.
BART_HD int prep_layer(const DevConfig &c, const double *in, int l, double *rho, int rho_stride,
                       double *mu_out) {
  const int nl = c.nlayer;
  const double T = in[l];
  const double p = c.press[l] * c.pfct;
  double mu = 0.0, sumq = 0.0;
  for (int j = 0; j < c.nspec; j++) {
    const double q = in[(size_t)nl * (j + 1) + l];
    mu += q * c.mass[j];
    sumq += q;
    const double r = cAMU * q * p / cKB / T;
    rho[(size_t)j * rho_stride] = r * c.mass[j];
  }
  *mu_out = mu;
  return sumq > 1.001 ? REJ_SUMQ : 0;
}

// atm_prep stage 2: hydrostatic radii, sequential in the layer index.
// Reference: radpress readatm.c:787-865 (same expression order).
BART_HD void hydrostatic_radii(const DevConfig &c, double r0, const double *temp, const double *mu,
                               double *radius) {
  const int nl = c.nlayer;
  const double *pr = c.press;
  const double p0 = c.p0, g0 = c.gsurf, rfct = c.rfct;
  int i0 = 0;
  double best = 1e37;
  for (int i = 0; i < nl; i++) {
    const double d = fabs(pr[i] - p0);
    if (d < best) { i0 = i; best = d; }
  }
  if (pr[i0] > p0) {
    const int i1 = i0 + 1 < nl ? i0 + 1 : i0;
    const double lr = log(pr[i1] / pr[i0]), lp = log(p0 / pr[i0]);
    const double t0 = temp[i0] + ((temp[i1] - temp[i0]) / lr) * lp;
    const double m0 = mu[i0] + ((mu[i1] - mu[i0]) / lr) * lp;
    radius[i0] = r0 + 0.5 * (temp[i0] / mu[i0] + t0 / m0) * (cKB / cAMU * lp / g0) / rfct;
  } else {
    const int i1 = i0 > 0 ? i0 - 1 : i0;
    const double lr = log(pr[i1] / pr[i0]), lp = log(p0 / pr[i0]);
    const double t0 = temp[i0] + ((temp[i1] - temp[i0]) / lr) * lp;
    const double m0 = mu[i0] + ((mu[i1] - mu[i0]) / lr) * lp;
    radius[i0] = r0 - 0.5 * (temp[i0] / mu[i0] + t0 / m0) * (cKB / cAMU * log(pr[i0] / p0) / g0) / rfct;
  }
  double ratio = r0 / radius[i0];
  double g = g0 * (ratio * ratio);
  for (int i = i0 - 1; i >= 0; i--) {
    radius[i] = radius[i + 1] - 0.5 * (temp[i] / mu[i] + temp[i + 1] / mu[i + 1]) *
                (cKB / cAMU * log(pr[i] / pr[i + 1]) / g) / rfct;
    ratio = radius[i + 1] / radius[i];
    g = g * (ratio * ratio);
  }
  ratio = r0 / radius[i0];
  g = g0 * (ratio * ratio);
  for (int i = i0 + 1; i < nl; i++) {
    radius[i] = radius[i - 1] + 0.5 * (temp[i] / mu[i] + temp[i - 1] / mu[i - 1]) *
                (cKB / cAMU * log(pr[i - 1] / pr[i]) / g) / rfct;
    ratio = radius[i - 1] / radius[i];
    g = g * (ratio * ratio);
  }
}

// floor-bracket search: the reference finds the NEAREST node (iomisc.c:1088-1108) and steps
// down when the value is below it (extinction.c:560-564, spline.c:149-154), i.e. the largest k
// with x[k] <= v, clamped so that k+1 is a valid node.
BART_HD int bracket(const double *x, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (hi + lo) >> 1;
    if (x[mid] > v) hi = mid; else lo = mid;
  }
  if (lo > n - 2) lo = n - 2;
  if (lo < 0) lo = 0;
  return lo;
}

// atm_prep stage 3, one depth d (0 = top): every per-layer coefficient the column kernels need,
// written as one record (layout: TabLayout).  temp/radius are indexed by layer (bottom -> top);
// rho[j*rho_stride + layer].
BART_HD int prep_table_row(const DevConfig &c, const KnobVals &kv, int d, const double *temp,
                           const double *rho, int rho_stride, const double *radius, double *tab) {
  const TabLayout &L = c.lay;
  const int nl = c.nlayer;
  const int l = nl - 1 - d;
  const double T = temp[l];
  double *row = tab + (size_t)d * L.nf();
  int status = 0;
  row[L.T] = T;
  row[L.INVT] = 1.0 / T;
  row[L.RAD] = radius[l];

  // opacity-grid bracket and folded weights (interpolmolext, extinction.c:534-581)
  if (T < c.gtemp[0] || T > c.gtemp[c.ntemp - 1]) status |= REJ_TGRID;
  const int it = bracket(c.gtemp, c.ntemp, T);
  const double t0 = c.gtemp[it], t1 = c.gtemp[it + 1];
  row[L.GOFF] = bits_to_double((((long long)l * c.ntemp + it) * c.ngmol) * (long long)c.nwave);
  for (int m = 0; m < c.ngmol; m++) {
    const double r = rho[(size_t)c.gmol_spec[m] * rho_stride + l];
    row[L.W + 2 * m] = r * (t1 - T) / (t1 - t0);
    row[L.W + 2 * m + 1] = r * (T - t0) / (t1 - t0);
  }

  // CIA: cubic-spline-in-T coefficients (splinterp_pt, spline.c:131-183) applied to the
  // wavenumber-pre-splined tables, times the density product (interpcs, crosssec.c:321-336)
  for (int f = 0; f < c.ncia; f++) {
    const double *x = c.ciaT[f];
    const int nt = c.cia_nt[f];
    if (T < x[0] || T > x[nt - 1]) status |= REJ_TCIA;
    const int k = bracket(x, nt, T);
    double dens = 1.0;
    for (int s = 0; s < c.cia_nspec[f]; s++) {
      const int sp = c.cia_spec[f][s];
      dens *= rho[(size_t)sp * rho_stride + l] / (cAMU * c.mass[sp] * cAMAGAT);
    }
    double cy0, cy1, cz0, cz1;
    if (x[k] == T) { cy0 = 1.0; cy1 = 0.0; cz0 = 0.0; cz1 = 0.0; }
    else {
      const double h = x[k + 1] - x[k], dx = T - x[k];
      const double u = dx / h;
      cy0 = 1.0 - u;
      cy1 = u;
      cz0 = dx * (-h / 3.0 + dx * (0.5 - dx / (6.0 * h)));
      cz1 = dx * (-h / 6.0 + dx * dx / (6.0 * h));
    }
    double *cr = row + L.cia(f);
    cr[0] = bits_to_double((long long)k * c.nwave);
    cr[1] = (double)k;
    cr[2] = cy0 * dens;
    cr[3] = cy1 * dens;
    cr[4] = cz0 * dens;
    cr[5] = cz1 * dens;
  }

  // scattering (computeextscat, extinction.c:586-624): coefficient of wn^4
  double sc = 0.0;
  if (kv.scat_flag == 1) sc = pow(10.0, kv.scat_logext) * cE0H2 * c.press[l] / T;
  else if (kv.scat_flag == 2) {
    const double k4 = (2.0 * cPI * cMICRON) * (2.0 * cPI * cMICRON) * (2.0 * cPI * cMICRON) *
                      (2.0 * cPI * cMICRON);
    for (int j = 0; j < c.nspec; j++)
      sc += cPI * 8e-32 / 3.0 * (c.pol[j] * c.pol[j]) * k4 * rho[(size_t)j * rho_stride + l] /
            c.mass[j] * cNAVO;
  }
  row[L.SCAT] = sc;

  // gray cloud deck (computeextcloud flag 1, extinction.c:629-693)
  double cl = 0.0;
  if (kv.cloud_flag == 1 && kv.cloudext != 0.0) {
    const double top = pow(10.0, kv.cloudtop), bot = pow(10.0, kv.cloudbot);
    if (c.press[l] >= top && c.press[l] < bot) cl = kv.cloudext;
  }
  row[L.CLOUD] = cl;

  // Simpson / trapezoid coefficients on the radius spacing (geth + simpson, numerical.c:390-525),
  // top-aligned panels: the panel ending at even depth d spans depths d-2, d-1, d.
  double sa = 0.0, sb = 0.0, scf = 0.0, tr = 0.0;
  if (d >= 1) {
    const double h0 = radius[l + 1] - radius[l];            // interval (d, d-1)
    tr = c.rfct * h0 / 2.0;
    if (d >= 2) {
      const double h1 = radius[l + 2] - radius[l + 1];      // interval (d-1, d-2)
      const double hsum = h0 + h1, hratio = h1 / h0, hfactor = hsum * hsum / (h0 * h1);
      const double s6 = c.rfct * hsum / 6.0;
      sa = (2.0 - hratio) * s6;
      sb = hfactor * s6;
      scf = (2.0 - 1.0 / hratio) * s6;
    }
  }
  row[L.SA] = sa;
  row[L.SB] = sb;
  row[L.SC] = scf;
  row[L.TR] = tr;
  return status;
}

// ---------------------------------------------------------------------------------------
// Total extinction of one (depth, wavenumber) cell: opacity-grid lookup with temperature
// interpolation and abundance scaling (extinction.c:534-581), + scattering + cloud + CIA in the
// reference's summation order (tau.c:231-232).  `row` is the depth's table record.  NMOL / NCIA
// are compile-time counts (0 = take them from the configuration at run time).
template <int NMOL, int NCIA>
BART_HD double cell_extinction(const DevConfig &c, const double *row, int w, double wn4,
                               bool mol_only) {
  typedef TabLayout L;
  const int ngmol = NMOL > 0 ? NMOL : c.ngmol;
  const int ncia = NCIA >= 0 ? NCIA : c.ncia;
  const size_t nw = (size_t)c.nwave;
  const D2 head = ld2(row + L::INVT);                        // (1/T, grid offset)
  const double *lo = c.grid + double_to_bits(head.y) + w;
  const double *hi = lo + (size_t)ngmol * nw;
  double e;
  {
    const D2 wt = ld2(row + L::W);
    e = wt.x * lo[0] + wt.y * hi[0];
  }
#pragma unroll
  for (int m = 1; m < (NMOL > 0 ? NMOL : kMaxGridMol); m++)
    if (m < ngmol) {
      const D2 wt = ld2(row + L::W + 2 * m);
      e += wt.x * lo[m * nw] + wt.y * hi[m * nw];
    }
  if (mol_only) return e;
  double ecs = 0.0;
  const double *cr = row + L::W + 2 * ngmol;
#pragma unroll
  for (int f = 0; f < (NCIA >= 0 ? NCIA : kMaxCia); f++) {
    if (f < ncia) {
      const long long off = double_to_bits(cr[6 * f]) + w;
      const double *P = c.ciaP[f] + off;
      const double *Q = c.ciaQ[f] + off;
      const D2 cy = ld2(cr + 6 * f + 2), cz = ld2(cr + 6 * f + 4);
      const double v = cy.x * P[0] + cy.y * P[nw] + cz.x * Q[0] + cz.y * Q[nw];
      if (v > 0) ecs += v;
    }
  }
  const D2 sc = ld2(row + L::SCAT);                          // (scattering coefficient, cloud)
  return e + sc.x * wn4 + sc.y + ecs;
}

// ---------------------------------------------------------------------------------------
// Eclipse geometry, one (model, wavenumber) column: fused lookup + optical-depth scan +
// emergent intensity + hemispheric flux.
//   tau: eclipsetau eclipse.c:28-105 restated as a top-aligned Simpson prefix scan (the
//        reference re-integrates from each depth to the top; its panels are always aligned to
//        the top sample, numerical.c:454-525, so tau(d) = sum of completed panels (+ one
//        trapezoid when d is odd));  stop at the first tau > toomuch (tau.c:277-287).
//   intensity: eclipse_intens eclipse.c:117-160; flux eclipse.c:242-287.
template <int NMOL, int NCIA, int NANG, bool KEEP>
BART_HD double eclipse_column(const DevConfig &c, const double *tab, const double *etab, int w,
                              double *tau_keep, int *last_keep) {
  typedef TabLayout L;
  const int nl = c.nlayer;
  const int nf = c.lay.nf();
  const int nang = NANG > 0 ? NANG : c.nang;
  const double wn = c.wn[w];
  const double wn4 = (wn * wn) * (wn * wn);
  const double c1 = 2.0 * cH * (wn * wn * wn) * cLS * cLS;
  const double c2 = cH * wn * cLS / cKB;
  double trap[NANG > 0 ? NANG : kMaxAng], dprev[NANG > 0 ? NANG : kMaxAng];
#pragma unroll
  for (int a = 0; a < (NANG > 0 ? NANG : kMaxAng); a++) { trap[a] = 0.0; dprev[a] = 1.0; }
  double S = 0.0, er1 = 0.0, er2 = 0.0, Bprev = 0.0;
  int last = nl - 1;
  const double *row = tab;
  for (int d = 0; d < nl; d++, row += nf) {
    const double er = cell_extinction<NMOL, NCIA>(c, row, w, wn4, false);
    double tau;
    if (d == 0) tau = 0.0;
    else {
      const D2 s1 = ld2(row + L::SA), s2 = ld2(row + L::SC);   // (SA, SB), (SC, TR)
      if (d & 1) tau = S + s2.y * (er + er1);
      else {
        S += s1.x * er + s1.y * er1 + s2.x * er2;
        tau = S;
      }
    }
    er2 = er1; er1 = er;
    if (KEEP) tau_keep[d] = tau;
    const double B = c1 * fast_rcp(fast_exp_pos(c2 * row[L::INVT], etab) - 1.0);
    const double Bs = B + Bprev;
#pragma unroll
    for (int a = 0; a < (NANG > 0 ? NANG : kMaxAng); a++) {
      if (a < nang) {
        const double dt = fast_exp_neg(-tau * c.inv_mu[a], etab);
        trap[a] += (dt - dprev[a]) * Bs;       // d = 0: dprev = 1 = dt, contributes exactly 0
        dprev[a] = dt;
      }
    }
    Bprev = B;
    if (tau > c.toomuch) { last = d; break; }
  }
  if (KEEP) *last_keep = last;
  double flux = 0.0;
#pragma unroll
  for (int a = 0; a < (NANG > 0 ? NANG : kMaxAng); a++)
    if (a < nang) flux += cPI * (Bprev * dprev[a] - 0.5 * trap[a]) * c.wgt[a];
  return flux;
}

// ---------------------------------------------------------------------------------------
// Transit geometry.  Chord weights of one model, row d: tau(d) = sum_i wt[i] * er[i], i = 0..d,
// restating totaltau1 (slantpath.c:18-108): abscissa s_i = sqrt(r_i^2 - b^2) along the chord at
// impact parameter b = r(depth d), top-aligned Simpson panels with a trapezoid on the bottom
// interval when the count is even, the two-point case through the reference's 3-point
// construction, result x2 (both halves of the chord) and x rfct (tau.c:274).
BART_HD void transit_weight_row(const DevConfig &c, const double *tab, int d, double *wt) {
  const int nf = c.lay.nf();
  auto rad = [&](int i) { return tab[(size_t)i * nf + TabLayout::RAD]; };   // depth-indexed radii
  for (int i = 0; i <= d; i++) wt[i] = 0.0;
  if (d == 0) return;
  const double b = rad(d);
  const double f = 2.0 * c.rfct;
  if (d == 1) {
    const double rm = (rad(1) + rad(0)) / 2.0;
    const double s1 = sqrt(rm * rm - b * b), s2 = sqrt(rad(0) * rad(0) - b * b);
    const double h0 = s1, h1 = s2 - s1;
    const double hsum = h0 + h1, hratio = h1 / h0, hfactor = hsum * hsum / (h0 * h1);
    const double a0 = (2.0 - hratio) * hsum / 6.0, a1 = hfactor * hsum / 6.0,
                 a2 = (2.0 - 1.0 / hratio) * hsum / 6.0;
    wt[1] = f * (a0 + 0.5 * a1);                             // bottom sample (depth 1)
    wt[0] = f * (a2 + 0.5 * a1);
    return;
  }
  // s at depth i
  auto sdep = [&](int i) { return i == d ? 0.0 : sqrt(rad(i) * rad(i) - b * b); };
  for (int p = 0; 2 * p + 2 <= d; p++) {
    const double sB = sdep(2 * p + 2), sM = sdep(2 * p + 1), sT = sdep(2 * p);
    const double h0 = sM - sB, h1 = sT - sM;
    const double hsum = h0 + h1, hratio = h1 / h0, hfactor = hsum * hsum / (h0 * h1);
    wt[2 * p + 2] += f * (2.0 - hratio) * hsum / 6.0;
    wt[2 * p + 1] += f * hfactor * hsum / 6.0;
    wt[2 * p]     += f * (2.0 - 1.0 / hratio) * hsum / 6.0;
  }
  if (d & 1) {                                               // even count: bottom trapezoid
    const double h = sdep(d - 1);
    wt[d]     += f * h / 2.0;
    wt[d - 1] += f * h / 2.0;
  }
}

// Transit column: tau(d) by the chord weights, stop at toomuch, then the modulation integral
// (modulation1, slantpath.c:350-436) as a top-aligned Simpson scan over impact parameter.
// `er` is per-thread scratch with stride `es` (shared memory in the kernel).
template <int NMOL, int NCIA, bool KEEP>
BART_HD double transit_column(const DevConfig &c, const double *tab, const double *etab,
                              const double *wts, int w, double *er, int es, double *tau_keep,
                              int *last_keep, int *status) {
  typedef TabLayout L;
  const int nl = c.nlayer;
  const int nf = c.lay.nf();
  const double wn = c.wn[w];
  const double wn4 = (wn * wn) * (wn * wn);
  double S = 0.0, f1 = 0.0, f2 = 0.0, tau = 0.0;
  int last = nl - 1;
  int d;
  for (d = 0; d < nl; d++) {
    const double *row = tab + (size_t)d * nf;
    er[(size_t)d * es] = cell_extinction<NMOL, NCIA>(c, row, w, wn4, false);
    const double *wr = wts + (size_t)d * (d + 1) / 2;
    tau = 0.0;
    for (int i = 0; i <= d; i++) tau += wr[i] * er[(size_t)i * es];
    if (KEEP) tau_keep[d] = tau;
    const double bd = row[L::RAD] * c.rfct;
    const double fd = fast_exp_neg(-tau, etab) * bd;
    if (d >= 2 && !(d & 1)) S += row[L::SA] * fd + row[L::SB] * f1 + row[L::SC] * f2;
    f2 = f1; f1 = fd;
    if (tau > c.toomuch) { last = d; break; }
  }
  if (KEEP) *last_keep = last;
  int n;                                                       // number of integration points
  if (last < nl - 1) {
    const int dd = last + 1;                                   // appended zero-integrand point
    const double *row = tab + (size_t)dd * nf;
    if (dd >= 2 && !(dd & 1)) S += row[L::SB] * f1 + row[L::SC] * f2;
    f2 = f1; f1 = 0.0;
    n = dd + 1;
  } else n = nl;
  if (n < 3) { *status |= REJ_FEWPTS; return -1.0; }
  if (!(n & 1)) S += tab[(size_t)(n - 1) * nf + L::TR] * (f1 + f2);
  const double btop = tab[L::RAD] * c.rfct;
  double res = btop * btop - 2.0 * S;
  if (c.transparent) {
    const double maxtau = tau > c.toomuch ? tau : c.toomuch;
    const double bl = tab[(size_t)(n - 1) * nf + L::RAD] * c.rfct;
    res -= fast_exp_neg(-maxtau, etab) * bl * bl;
  }
  return res * c.inv_srad2;
}

}  // namespace bart
