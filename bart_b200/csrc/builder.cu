// builder.cu -- stage (d): the line-by-line opacity-grid builder (`transit --justOpacity`).
//
// Reference: calcprofiles / calcopacity (transit/src/opacity.c:218-427), getprofile and
// computemolext(permol=1) (transit/src/extinction.c:8-57, 281-529), voigtn / voigtxy
// (pu/src/voigt.c:132-200, 369-554).  The reference loops layers x temperatures x lines on one
// core; here:
//
//   K5  voigt_table_kernel     one thread per output bin of every unique (Doppler, Lorentz)
//                              profile: fine samples by the 3-region Pierluissi approximation,
//                              bin-averaged in float32 with the reference's operation order.
//   line_index_kernel          per line: wavenumber, oversampled/coarse bin indices with
//                              non-contracted IEEE ops (bit-exact indices).
//   (host)                     co-add grouping of neighbouring lines -- temperature independent,
//                              sequential by construction (extinction.c:450-462), done once.
//   K6a kmax_kernel            per temperature: strongest line per molecule (extinction.c:400-427)
//   K6b strength_kernel        per temperature: co-added strength per group; these do not depend
//                              on the layer, so they are computed once per T, not per cell.
//   K6c widths_kernel          per (layer, isotope): Lorentz/Doppler widths, table indices, the
//                              carried Doppler index of the reference's sequential loop.
//   K6d accumulate_kernel      per (layer, 128-bin wavenumber tile): GATHER over the candidate
//                              groups in line order -- no atomics, summation order identical to
//                              the reference, profile samples fetched with the reference's
//                              stride-`wnosamp` indexing.
//
// The temperature axis is the sharding axis (bart_build_opacity_slice): planes are independent.
#include "builder.hpp"
#include "device.cuh"
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>
#include <map>
#include <string>
#include <chrono>
#include <unistd.h>
#include <fcntl.h>

namespace bart {

#define BCUDA(call)                                                                        \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_),     \
                                __FILE__, __LINE__, #call);                                \
  } while (0)

constexpr int kMaxIso = 64;
constexpr int kAccThreads = 128;

struct BuilderState {
  // sampling
  int nwave = 0, osamp = 0, nlayer = 0, nspec = 0, niso = 0, ngmol = 0, ntemp = 0;
  long long nowns = 0;
  double wn_lo = 0, dwn = 0, odwn = 0;
  std::vector<double> temps;
  // isotopes
  std::vector<int> iso_spec, iso_gmol, gmol_id;
  std::vector<double> ziso;                 // [niso][ntemp]
  // lines / groups
  long long nlines = 0, ngroups = 0, neval = 0;
  double *d_wl = nullptr, *d_elow = nullptr, *d_gf = nullptr, *d_wavn = nullptr;   // raw lines
  double *d_c_wavn = nullptr, *d_c_elow = nullptr, *d_c_gf = nullptr;             // grouped lines
  short *d_isoid = nullptr;
  int *d_iown = nullptr, *d_idwn = nullptr;
  unsigned char *d_inrange = nullptr;
  long long *d_gstart = nullptr;            // [ngroups+1]
  int *d_giown = nullptr, *d_gidwn = nullptr;
  short *d_giso = nullptr;
  double *d_gwavn = nullptr, *d_gS = nullptr;
  std::vector<long long> iso_gbeg;          // [niso+1] group range per isotope
  std::vector<int> h_iown;                  // per-line trace (leader bin, or -2-bin when co-added)
  // Voigt table
  int nDop = 0, nLor = 0;
  std::vector<double> aDop, aLor;
  std::vector<long long> prof_off, prof_size;     // [nDop*nLor] offset (floats) and half-size
  float *d_prof = nullptr;
  long long prof_total = 0;
  double *d_aDop = nullptr, *d_aLor = nullptr;
  long long *d_prof_off = nullptr, *d_prof_size = nullptr;
  // per-layer state
  double *d_density = nullptr;              // [nlayer][nspec] for the current temperature
  double *d_kmax = nullptr;                 // [ngmol]
  double *d_out = nullptr;                  // [nlayer][ngmol][nwave] for the current temperature
  void *d_cellinfo = nullptr;
  bool lines_loaded = false, profiles_ready = false;
  // per-phase wall/device milliseconds (bart_builder_phase_ms)
  std::map<std::string, double> phase_ms;
};

// Times a phase of the build on stream `s` with CUDA events (the builder synchronises between
// phases anyway, so this costs nothing on the data path).
struct PhaseTimer {
  BuilderState *b; const char *name; cudaStream_t s; cudaEvent_t e0, e1;
  PhaseTimer(BuilderState *b_, const char *n, cudaStream_t s_) : b(b_), name(n), s(s_) {
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s);
  }
  ~PhaseTimer() {
    cudaEventRecord(e1, s); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    b->phase_ms[name] += ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
};

// ---------------------------------------------------------------------------------------
// Voigt function (voigt.c:132-200).  Region I is a power series the reference evaluates in
// 80-bit long double; fp64 here (difference <= 1e-10 relative before the float32 rounding).
__constant__ double c_ferf[64];

__device__ float voigtxy_dev(double x, double y, double alphaD) {
  const double A1 = 0.46131350, A2 = 0.19016350, A3 = 0.09999216, A4 = 1.78449270,
               A5 = 0.002883894, A6 = 5.52534370, B1 = 0.51242424, B2 = 0.27525510,
               B3 = 0.05176536, B4 = 2.72474500;
  const double SQRTLN2PI = 0.46971863934982566689, TWOOSQRTPI = 1.12837916709551257389;
  const double x2y2 = x * x - y * y, xy2 = 2 * x * y;
  if (x < 3 && y < 1.8) {
    double sinxy, cosxy;
    sincos(xy2, &sinxy, &cosxy);
    const int n = (x < 1 ? 15 : (int)(6.842 * x + 8.0)) + 1;
    double orr = y, oi = -x, ar = y, ai = -x;
    for (int i = 1; i <= n; i++) {
      const double ni = orr * xy2 + oi * x2y2;
      const double nr = orr * x2y2 - oi * xy2;
      ai += ni * c_ferf[i];
      ar += nr * c_ferf[i];
      oi = ni; orr = nr;
    }
    return (float)(SQRTLN2PI / alphaD * exp(-x2y2) *
                   (cosxy * (1 - ar * TWOOSQRTPI) - sinxy * ai * TWOOSQRTPI));
  }
  const double ar = xy2 * xy2, nr = xy2 * x;
  if (x < 5 && y < 5) {
    const double ni = x2y2 - A2, ai = x2y2 - A4, oi = x2y2 - A6;
    return (float)(SQRTLN2PI / alphaD * (A1 * ((nr - ni * y) / (ni * ni + ar)) +
                                         A3 * ((nr - ai * y) / (ai * ai + ar)) +
                                         A5 * ((nr - oi * y) / (oi * oi + ar))));
  }
  const double ni = x2y2 - B2, ai = x2y2 - B4;
  return (float)(SQRTLN2PI / alphaD * (B1 * ((nr - ni * y) / (ni * ni + ar)) +
                                       B3 * ((nr - ai * y) / (ai * ai + ar))));
}

struct ProfJob {          // one unique profile
  long long off;          // offset into the float pool
  int nwn;                // 2*halfsize+1
  int ipo;                // fine intervals per bin (1 => adjacent average / quick)
  int quick;
  double dint, dwn_half, alphaL, alphaD;
};

// voigtn (voigt.c:369-483) + meanintegSimp/Trap (489-554): thread per output bin.
__global__ void voigt_table_kernel(const ProfJob *jobs, float *pool) {
  const ProfJob jb = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= jb.nwn) return;
  const double y = 0.83255461115769775635 * jb.alphaL / jb.alphaD;
  const long long base = (long long)k * jb.ipo;
  auto sample = [&](long long i) {
    const double xx = 0.83255461115769775635 * fabs(jb.dint * (double)i - jb.dwn_half) / jb.alphaD;
    return voigtxy_dev(xx, y, jb.alphaD);
  };
  float out;
  if (jb.quick) out = sample(base);
  else if (jb.ipo & 1) {                       // even point count: trapezoid (meanintegTrap)
    float acc = 0;
    for (int i = 1; i < jb.ipo; i++) acc += sample(base + i);
    const float ends = sample(base) + sample(base + jb.ipo);
    out = (float)((acc + ends / 2.0) / (double)jb.ipo);
  } else {                                     // odd point count: Simpson (meanintegSimp)
    float acc = 0;
    for (int i = 1; i < jb.ipo; i += 2) acc += sample(base + i);
    acc *= 2;
    for (int i = 2; i < jb.ipo; i += 2) acc += sample(base + i);
    acc *= 2;
    acc += sample(base) + sample(base + jb.ipo);
    out = (float)(acc / (jb.ipo * 3.0));
  }
  pool[jb.off + k] = out;
}

// ---------------------------------------------------------------------------------------
// Per-line indices (extinction.c:431-447, 476).  Non-contracted IEEE operations so that the
// integer truncations and the nearest-node test see the same doubles as a plain C build.
__global__ void line_index_kernel(const double *wl, long long n, double wn_lo, double own_last,
                                  double odwn, double dwn, double *wavn_out, int *iown_out,
                                  int *idwn_out, unsigned char *inrange) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double wavn = __ddiv_rn(1.0, __dmul_rn(wl[i], 1e-4));
  wavn_out[i] = wavn;
  const bool in = !(wavn < wn_lo || wavn > own_last);
  inrange[i] = in ? 1 : 0;
  int iown = 0, idwn = 0;
  if (in) {
    iown = (int)__ddiv_rn(__dadd_rn(wavn, -wn_lo), odwn);
    const double v0 = __dadd_rn(wn_lo, __dmul_rn((double)iown, odwn));
    const double v1 = __dadd_rn(wn_lo, __dmul_rn((double)(iown + 1), odwn));
    if (fabs(__dadd_rn(wavn, -v1)) < fabs(__dadd_rn(wavn, -v0))) iown++;
    idwn = (int)__ddiv_rn(__dadd_rn(wavn, -wn_lo), dwn);
  }
  iown_out[i] = iown;
  idwn_out[i] = idwn;
}

// K6a: strongest individual line per output molecule at temperature T (extinction.c:400-427)
__global__ void kmax_kernel(const double *wavn, const double *elow, const double *gf,
                            const short *isoid, const unsigned char *inrange, long long n, double T,
                            const double *iso_fac /*[niso] ratio*SIGCTE/(mass*Z)*/,
                            const int *iso_gmol, unsigned long long *kmax_bits, int ngmol) {
  __shared__ double s_max[kMaxGridMol];
  if (threadIdx.x < kMaxGridMol) s_max[threadIdx.x] = 0.0;
  __syncthreads();
  double loc[kMaxGridMol];
#pragma unroll
  for (int m = 0; m < kMaxGridMol; m++) loc[m] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (!inrange[i]) continue;
    const int iso = isoid[i];
    const double pk = iso_fac[iso] * gf[i] * exp(-kEXPCTE * elow[i] / T) *
                      (1 - exp(-kEXPCTE * wavn[i] / T));
    const int m = iso_gmol[iso];
#pragma unroll
    for (int q = 0; q < kMaxGridMol; q++) if (q == m) loc[q] = fmax(loc[q], pk);
  }
#pragma unroll
  for (int m = 0; m < kMaxGridMol; m++)
    if (m < ngmol && loc[m] > 0)
      atomicMax((unsigned long long *)&s_max[m], (unsigned long long)__double_as_longlong(loc[m]));
  __syncthreads();
  if (threadIdx.x < ngmol && s_max[threadIdx.x] > 0)
    atomicMax(&kmax_bits[threadIdx.x], (unsigned long long)__double_as_longlong(s_max[threadIdx.x]));
}

// K6b: co-added group strength at temperature T (extinction.c:439-464), thread per group,
// lines of a group summed in file order.
__global__ void strength_kernel(const long long *gstart, const short *giso, const double *wavn,
                                const double *elow, const double *gf, long long ngroups, double T,
                                const double *iso_fac2 /*[niso] SIGCTE*ratio/(mass*Z)*/,
                                double *gS) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const long long b = gstart[g], e = gstart[g + 1];
  double pk = 0.0;
  for (long long i = b; i < e; i++) {
    const double term = gf[i] * exp(-kEXPCTE * elow[i] / T) * (1 - exp(-kEXPCTE * wavn[i] / T));
    pk = (i == b) ? term : pk + term;
  }
  gS[g] = pk * iso_fac2[giso[g]];
}

struct CellIso {          // per (layer, isotope) for the current temperature
  double alphal, alphad;  // Lorentz width; Doppler width / wavenumber
  int ilor, idop0;        // table indices (idop0: at wn[0])
  int idop_carry;         // Doppler index used by lines below the re-pick threshold
  int hwbins;             // max profile half-width in coarse bins (+2)
  long long gsplit;       // first group (in the isotope's range) with alphad*wavn/alphal < 0.1
};

__device__ int nearest_dev(const double *a, double v, int lo, int hi) {
  while (hi - lo > 1) {
    const int mid = (hi + lo) >> 1;
    if (a[mid] > v) hi = mid; else lo = mid;
  }
  if (hi == lo) return lo;
  return fabs(a[hi] - v) < fabs(a[lo] - v) ? hi : lo;
}

// K6c: widths and table indices per (layer, isotope) (extinction.c:365-396, 478-483).
__global__ void widths_kernel(CellIso *cells, int nlayer, int niso, int nspec, double T,
                              const double *density /*[nlayer][nspec]*/, const double *spec_mass,
                              const double *spec_radius, const double *iso_mass, const int *iso_spec,
                              const int *iso_gmol, const double *aDop, const double *aLor, int nDop,
                              int nLor, const long long *prof_size, int osamp, double wn0,
                              const long long *iso_gbeg, const double *gwavn, const double *gS,
                              const double *kmax, double ethresh) {
  const int r = blockIdx.x, i = threadIdx.x;
  if (r >= nlayer || i >= niso) return;
  const double fdoppler = sqrt(2 * kKB * T / kAMU) * kSQRTLN2 / kLS;
  const double florentz = sqrt(2 * kKB * T / kPI / kAMU) / (kAMU * kLS);
  double al = 0.0;
  for (int j = 0; j < nspec; j++) {
    const double cs = spec_radius[j] + spec_radius[iso_spec[i]];
    al += density[(size_t)r * nspec + j] / spec_mass[j] * cs * cs *
          sqrt(1 / iso_mass[i] + 1 / spec_mass[j]);
  }
  al *= florentz;
  const double ad = fdoppler / sqrt(iso_mass[i]);
  CellIso c;
  c.alphal = al; c.alphad = ad;
  // the reference searches with hi = nDop / nLor (one past the end, extinction.c:394-395);
  // clamped to the last valid entry here
  c.idop0 = nearest_dev(aDop, ad * wn0, 0, nDop - 1);
  c.ilor = nearest_dev(aLor, al, 0, nLor - 1);
  // lines are sorted by decreasing wavenumber inside an isotope, so "alphad*wavn/alphal >= 0.1"
  // holds for a prefix of the isotope's groups
  const long long gb = iso_gbeg[i], ge = iso_gbeg[i + 1];
  long long lo = gb, hi = ge;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (ad * gwavn[mid] / al >= 1e-1) lo = mid + 1; else hi = mid;
  }
  c.gsplit = lo;
  // the sequential loop keeps the Doppler index of the last line that was re-picked AND evaluated
  // (weak lines `continue` before the re-pick, extinction.c:467-483)
  int carry = c.idop0;
  const double thr = ethresh * kmax[iso_gmol[i]];
  for (long long g = lo - 1; g >= gb; g--)
    if (!(gS[g] < thr)) { carry = nearest_dev(aDop, ad * gwavn[g], 0, nDop - 1); break; }
  c.idop_carry = carry;
  long long hw = 0;
  for (int d = 0; d < nDop; d++) hw = max(hw, prof_size[(size_t)d * nLor + c.ilor]);
  c.hwbins = (int)(hw / osamp) + 2;
  cells[(size_t)r * niso + i] = c;
}

// K6d: gather-accumulate.  Block = (128-bin tile, layer); thread <-> coarse bin j.  For every
// isotope the candidate groups (leaders whose coarse bin lies within the widest profile of the
// tile) are staged 128 at a time into shared memory -- one group per thread: Doppler index,
// profile pointer, bin range -- and then every thread walks the staged groups in line order,
// adding S * profile[wnosamp*j - offset] when its bin is inside the group's range
// (extinction.c:486-509).
struct StagedGroup {
  const float *prof;
  double S;
  int offset, ps2, minj, maxj;
};

__global__ void __launch_bounds__(kAccThreads)
accumulate_kernel(const CellIso *cells, int niso, int ngmol, int nwave, int osamp,
                  const int *iso_gmol, const long long *iso_gbeg, const int *giown,
                  const int *gidwn, const double *gwavn, const double *gS, const double *kmax,
                  double ethresh, const double *aDop, int nDop, int nLor,
                  const long long *prof_off, const long long *prof_size, const float *pool,
                  double *out /*[nlayer][ngmol][nwave]*/, unsigned long long *neval) {
  __shared__ StagedGroup s_g[kAccThreads];
  __shared__ double s_aDop[128];
  const int r = blockIdx.y;
  const int j0 = blockIdx.x * kAccThreads;
  const int j = j0 + threadIdx.x;
  for (int k = threadIdx.x; k < nDop && k < 128; k += blockDim.x) s_aDop[k] = aDop[k];
  __syncthreads();
  int cur_mol = -1;
  double acc = 0.0;
  unsigned long long my_eval = 0;
  for (int iso = 0; iso < niso; iso++) {
    const int m = iso_gmol[iso];
    if (m != cur_mol) {
      if (cur_mol >= 0 && j < nwave) out[((size_t)r * ngmol + cur_mol) * nwave + j] = acc;
      // isotopes of one molecule are contiguous (TLI database order); a molecule that re-appears
      // continues from what was stored
      acc = (j < nwave && m >= 0) ? out[((size_t)r * ngmol + m) * nwave + j] : 0.0;
      cur_mol = m;
    }
    const CellIso c = cells[(size_t)r * niso + iso];
    const long long gb = iso_gbeg[iso], ge = iso_gbeg[iso + 1];
    if (gb == ge) continue;
    // candidate range: leader coarse bins in [j0 - hw, j0 + 127 + hw]; gidwn is non-increasing
    const int hi_bin = j0 + kAccThreads - 1 + c.hwbins, lo_bin = j0 - c.hwbins;
    long long a = gb, b = ge;
    while (a < b) { const long long mid = (a + b) >> 1; if (gidwn[mid] > hi_bin) a = mid + 1; else b = mid; }
    const long long first = a;
    b = ge;
    while (a < b) { const long long mid = (a + b) >> 1; if (gidwn[mid] >= lo_bin) a = mid + 1; else b = mid; }
    const long long last = a;                                   // exclusive
    const double thr = ethresh * kmax[m];
    for (long long base = first; base < last; base += kAccThreads) {
      const long long g = base + threadIdx.x;
      StagedGroup sg;
      sg.prof = nullptr; sg.S = 0.0; sg.offset = 0; sg.ps2 = -1; sg.minj = 1; sg.maxj = 0;
      if (g < last) {
        const double S = gS[g];
        if (!(S < thr)) {                                       // weak-line cut (467-470)
          int idop = c.idop_carry;
          if (g < c.gsplit) idop = nearest_dev(s_aDop, c.alphad * gwavn[g], 0, nDop - 1);
          const size_t pi = (size_t)idop * nLor + c.ilor;
          const int ps = (int)prof_size[pi];
          const int iown = giown[g], idwn = gidwn[g];
          const int subw = iown - idwn * osamp;
          sg.offset = iown - ps;
          sg.minj = idwn - (ps - subw) / osamp;
          sg.maxj = idwn + (ps + subw) / osamp;
          if (sg.minj < 0) sg.minj = 0;
          if (sg.maxj >= nwave) sg.maxj = nwave - 1;
          sg.ps2 = 2 * ps;
          sg.prof = pool + prof_off[pi];
          sg.S = S;
          if (blockIdx.x == 0) my_eval++;                       // counted once per layer
        }
      }
      __syncthreads();
      s_g[threadIdx.x] = sg;
      __syncthreads();
      const int cnt = (int)min((long long)kAccThreads, last - base);
      for (int q = 0; q < cnt; q++) {
        const StagedGroup &t = s_g[q];
        if (j >= t.minj && j <= t.maxj) {
          const int bj = osamp * j - t.offset;
          if (bj >= 0 && bj <= t.ps2) acc += t.S * (double)t.prof[bj];
        }
      }
    }
  }
  if (cur_mol >= 0 && j < nwave) out[((size_t)r * ngmol + cur_mol) * nwave + j] = acc;
  if (neval && my_eval) atomicAdd(neval, my_eval);
}

// ---------------------------------------------------------------------------------------
// host side
template <class T> static T *dev_upload(const std::vector<T> &v) {
  T *p = nullptr;
  BCUDA(cudaMalloc((void **)&p, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) BCUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

static void setup_static(BuilderState *b, const Options &o, const Atmosphere &a, const Molecules &mol,
                         Tli &t, const std::vector<double> &wn) {
  b->nwave = (int)wn.size();
  b->osamp = o.wnosamp;
  b->nlayer = a.nlayer();
  b->nspec = a.nspec();
  b->wn_lo = wn[0];
  b->dwn = o.wndelt;
  b->odwn = o.wndelt / o.wnosamp;                                  // owns.d / owns.o
  b->nowns = (long long)(b->nwave - 1) * o.wnosamp + 1;            // makesample1, makesample.c:93
  b->temps = make_sampling(o.tlow, o.thigh, o.tempdelt, 1);        // maketempsample
  b->ntemp = (int)b->temps.size();
  if (b->temps.front() < t.tmin)
    fail("The opacity file attempted to sample a temperature (%.1f K) below the lowest allowed TLI "
         "temperature (%.1f K).", b->temps.front(), t.tmin);
  if (b->temps.back() > t.tmax)
    fail("The opacity file attempted to sample a temperature (%.1f K) beyond the highest allowed "
         "TLI temperature (%.1f K).", b->temps.back(), t.tmax);
  b->niso = t.niso();
  if (b->niso > kMaxIso) fail("at most %d isotopes are supported", kMaxIso);
  // setimol (readlineinfo.c:249-278) and the molID list of calcopacity (opacity.c:353-361)
  b->iso_spec.assign(b->niso, -1);
  b->iso_gmol.assign(b->niso, -1);
  b->gmol_id.clear();
  for (int i = 0; i < b->niso; i++) {
    const std::string &mn = t.db[t.iso_db[i]].molname;
    for (int j = 0; j < a.nspec(); j++) if (a.species[j] == mn) b->iso_spec[i] = j;
    if (b->iso_spec[i] < 0) fail("TLI molecule '%s' is not among the atmospheric species.", mn.c_str());
    const int id = mol.id[b->iso_spec[i]];
    int g = -1;
    for (size_t k = 0; k < b->gmol_id.size(); k++) if (b->gmol_id[k] == id) g = (int)k;
    if (g < 0) { b->gmol_id.push_back(id); g = (int)b->gmol_id.size() - 1; }
    b->iso_gmol[i] = g;
  }
  b->ngmol = (int)b->gmol_id.size();
  if (b->ngmol > kMaxGridMol) fail("at most %d line-list molecules are supported", kMaxGridMol);
  // partition functions on the temperature grid (opacity.c:325-339)
  b->ziso.assign((size_t)b->niso * b->ntemp, 0.0);
  for (int i = 0; i < b->niso; i++) {
    const std::vector<double> &T = t.db[t.iso_db[i]].T;
    std::vector<double> z(T.size());
    spline_second_derivs(T.data(), t.iso_Z[i].data(), (long)T.size(), z.data());
    for (int k = 0; k < b->ntemp; k++)
      b->ziso[(size_t)i * b->ntemp + k] = spline_eval(z.data(), (long)T.size(), T.data(),
                                                      t.iso_Z[i].data(), b->temps[k]);
  }
}

// calcprofiles (opacity.c:218-277) + getprofile (extinction.c:8-57)
static void build_profiles(BuilderState *b, const Options &o, cudaStream_t s) {
  if (b->profiles_ready) return;
  b->nDop = o.ndop; b->nLor = o.nlor;
  if (b->nDop > 128) fail("at most 128 Doppler-width samples are supported");
  auto logspace = [](double lo, double hi, int n) {             // iomisc.c:1064-1083
    std::vector<double> v(n);
    const double l0 = log10(lo), l1 = log10(hi), st = (l1 - l0) / (n - 1.0);
    for (int i = 0; i < n; i++) v[i] = pow(10, l0 + i * st);
    return v;
  };
  b->aDop = logspace(o.dmin, o.dmax, b->nDop);
  b->aLor = logspace(o.lmin, o.lmax, b->nLor);
  const int np = b->nDop * b->nLor;
  b->prof_off.assign(np, 0); b->prof_size.assign(np, 0);
  std::vector<ProfJob> jobs;
  long long total = 0;
  const double dwn = b->odwn;
  const float ta = o.nwidth;
  for (int i = 0; i < b->nDop; i++)
    for (int j = 0; j < b->nLor; j++) {
      const int p = i * b->nLor + j;
      if (b->aDop[i] * 10.0 < b->aLor[j] && i != 0) {             // reuse the previous Doppler row
        b->prof_off[p] = b->prof_off[p - b->nLor];
        b->prof_size[p] = b->prof_size[p - b->nLor];
        continue;
      }
      const float dop = (float)b->aDop[i], lor = (float)b->aLor[j];   // PREC_VOIGT arguments
      double big = dop; if (big < lor) big = lor;
      const double wvgt = big * ta;
      int nvgt = 2 * (long)(wvgt / dwn + 0.5) + 1;
      if (nvgt < 2) nvgt = 3;
      if (nvgt > 2 * b->nowns) nvgt = 2 * (int)b->nowns + 1;
      ProfJob jb;
      jb.off = total; jb.nwn = nvgt;
      jb.alphaL = lor; jb.alphaD = dop;
      jb.dwn_half = dwn * (long)(nvgt / 2);
      jb.quick = nvgt > 99999 ? 1 : 0;                            // _voigt_maxelements
      // voigtn's choice of the fine sampling (voigt.c:393-431)
      const double ddwn = 2.0 * jb.dwn_half / (nvgt - 1);
      double dint = jb.alphaD / 49.0;
      if (ddwn < dint || jb.quick) { jb.dint = ddwn; jb.ipo = 1; }
      else {
        int nint = (int)(ddwn / dint) + 1;
        if (nint & 1) nint++;
        const long long ntot = (long long)nvgt * nint + 1;
        jb.dint = 2.0 * jb.dwn_half / (double)(ntot - 1);
        jb.ipo = nint;
      }
      b->prof_off[p] = total;
      b->prof_size[p] = nvgt / 2;
      total += nvgt;
      jobs.push_back(jb);
    }
  b->prof_total = total;
  BCUDA(cudaMalloc((void **)&b->d_prof, std::max<long long>(1, total) * sizeof(float)));
  // series coefficients 1/(n!(2n+1)) (the table of voigt.c:47-108)
  double ferf[64];
  long double fac = 1.0L;
  ferf[0] = 1.0;
  for (int n = 1; n < 64; n++) { fac *= n; ferf[n] = (double)(1.0L / (fac * (2 * n + 1))); }
  BCUDA(cudaMemcpyToSymbol(c_ferf, ferf, sizeof(ferf)));
  ProfJob *d_jobs = dev_upload(jobs);
  int maxn = 0;
  for (auto &j : jobs) maxn = std::max(maxn, j.nwn);
  // grid.y is limited to 65535 jobs; 60x60 profiles fit
  const int chunk = 32768;
  PhaseTimer pt(b, "voigt_table", s);
  for (size_t j0 = 0; j0 < jobs.size(); j0 += chunk) {
    const int nj = (int)std::min<size_t>(chunk, jobs.size() - j0);
    dim3 grid((maxn + 127) / 128, nj);
    voigt_table_kernel<<<grid, 128, 0, s>>>(d_jobs + j0, b->d_prof);
  }
  BCUDA(cudaGetLastError());
  BCUDA(cudaStreamSynchronize(s));
  cudaFree(d_jobs);
  b->d_aDop = dev_upload(b->aDop);
  b->d_aLor = dev_upload(b->aLor);
  b->d_prof_off = dev_upload(b->prof_off);
  b->d_prof_size = dev_upload(b->prof_size);
  b->profiles_ready = true;
}

static void load_lines(BuilderState *b, const Options &o, Tli &t, const std::vector<double> &wn,
                       cudaStream_t s) {
  if (b->lines_loaded) return;
  double lo = wn.front(), hi_hint;
  // readdatarng selects with the HINTED limits wns.i / wns.f (readlineinfo.c:435-436)
  hi_hint = o.wnhigh > 0 ? o.wnhigh * o.wnfct : 1.0 / (o.wllow * o.wlfct);
  auto tp0 = std::chrono::steady_clock::now();
  read_tli_lines(o.linedb, t, lo, hi_hint);
  auto tp1 = std::chrono::steady_clock::now();
  b->phase_ms["read_tli_host"] += std::chrono::duration<double, std::milli>(tp1 - tp0).count();
  const long long n = (long long)t.wl.size();
  b->nlines = n;
  b->d_wl = dev_upload(t.wl); b->d_elow = dev_upload(t.elow); b->d_gf = dev_upload(t.gf);
  b->d_isoid = dev_upload(t.isoid);
  BCUDA(cudaMalloc((void **)&b->d_wavn, std::max<long long>(1, n) * 8));
  BCUDA(cudaMalloc((void **)&b->d_iown, std::max<long long>(1, n) * 4));
  BCUDA(cudaMalloc((void **)&b->d_idwn, std::max<long long>(1, n) * 4));
  BCUDA(cudaMalloc((void **)&b->d_inrange, std::max<long long>(1, n)));
  const double own_last = b->wn_lo + (double)(b->nowns - 1) * b->odwn;
  if (n > 0) {
    line_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b->d_wl, n, b->wn_lo, own_last, b->odwn,
                                                                   b->dwn, b->d_wavn, b->d_iown, b->d_idwn,
                                                                   b->d_inrange);
    BCUDA(cudaGetLastError());
  }
  std::vector<double> wavn(n);
  std::vector<int> iown(n), idwn(n);
  std::vector<unsigned char> inr(n);
  BCUDA(cudaStreamSynchronize(s));
  if (n > 0) {
    BCUDA(cudaMemcpy(wavn.data(), b->d_wavn, n * 8, cudaMemcpyDeviceToHost));
    BCUDA(cudaMemcpy(iown.data(), b->d_iown, n * 4, cudaMemcpyDeviceToHost));
    BCUDA(cudaMemcpy(idwn.data(), b->d_idwn, n * 4, cudaMemcpyDeviceToHost));
    BCUDA(cudaMemcpy(inr.data(), b->d_inrange, n, cudaMemcpyDeviceToHost));
  }
  // co-add grouping (extinction.c:450-462): a leader absorbs the following lines of the same
  // isotope while |wavn - owns[iown_leader]| < odwn.  Out-of-range lines never lead, but are
  // absorbed when they follow a leader (the reference does not re-test the range in its while
  // loop).  The grouping does not depend on temperature or layer, so it is done once, here.
  auto tg0 = std::chrono::steady_clock::now();
  std::vector<long long> bounds;            // [ngroups+1] into the grouped line arrays
  std::vector<int> giown, gidwn;
  std::vector<short> giso;
  std::vector<double> gwavn, c_wavn, c_elow, c_gf;
  c_wavn.reserve(n); c_elow.reserve(n); c_gf.reserve(n);
  b->h_iown.assign(n, -1);
  b->iso_gbeg.assign(b->niso + 1, 0);
  int prev_iso = -1;
  for (long long ln = 0; ln < n; ln++) {
    const int iso = t.isoid[ln];
    if (iso < 0 || iso >= b->niso) fail("TLI line %lld has isotope index %d outside [0,%d)", ln, iso, b->niso);
    if (iso != prev_iso) {
      if (iso < prev_iso) fail("TLI lines are not grouped by isotope");
      for (int k = prev_iso + 1; k <= iso; k++) b->iso_gbeg[k] = (long long)giown.size();
      prev_iso = iso;
    }
    if (!inr[ln]) continue;
    const int io = iown[ln];
    const double vnode = b->wn_lo + (double)io * b->odwn;
    long long e = ln + 1;
    while (e < n && t.isoid[e] == iso && std::fabs(wavn[e] - vnode) < b->odwn) e++;
    bounds.push_back((long long)c_wavn.size());
    for (long long k = ln; k < e; k++) {
      c_wavn.push_back(wavn[k]); c_elow.push_back(t.elow[k]); c_gf.push_back(t.gf[k]);
    }
    giown.push_back(io); gidwn.push_back(idwn[ln]); giso.push_back((short)iso); gwavn.push_back(wavn[ln]);
    b->h_iown[ln] = io;
    for (long long k = ln + 1; k < e; k++) b->h_iown[k] = -2 - io;
    ln = e - 1;
  }
  bounds.push_back((long long)c_wavn.size());
  for (int k = prev_iso + 1; k <= b->niso; k++) b->iso_gbeg[k] = (long long)giown.size();
  b->ngroups = (long long)giown.size();
  b->phase_ms["grouping_host"] += std::chrono::duration<double, std::milli>(
      std::chrono::steady_clock::now() - tg0).count();
  b->d_gstart = dev_upload(bounds);
  b->d_giown = dev_upload(giown); b->d_gidwn = dev_upload(gidwn); b->d_giso = dev_upload(giso);
  b->d_gwavn = dev_upload(gwavn);
  b->d_c_wavn = dev_upload(c_wavn); b->d_c_elow = dev_upload(c_elow); b->d_c_gf = dev_upload(c_gf);
  BCUDA(cudaMalloc((void **)&b->d_gS, std::max<long long>(1, b->ngroups) * 8));
  cudaFree(b->d_wl); b->d_wl = nullptr;     // wavelengths are no longer needed on the device
  b->lines_loaded = true;
}

// One temperature plane for all layers: out[layer][mol][wave] on the device.
static void build_temperature(BuilderState *b, const Options &o, const Atmosphere &a,
                              const Molecules &mol, const Tli &t, int it, cudaStream_t s,
                              const int *d_iso_spec, const int *d_iso_gmol, const double *d_iso_mass,
                              const double *d_spec_mass, const double *d_spec_radius,
                              const long long *d_iso_gbeg, unsigned long long *d_neval) {
  const double T = b->temps[it];
  const int nl = b->nlayer, ns = b->nspec;
  // densities: stateeqnford with number abundances (transit.h:58-69, opacity.c:390-394)
  std::vector<double> dens((size_t)nl * ns);
  for (int r = 0; r < nl; r++)
    for (int j = 0; j < ns; j++) {
      const double rho = kAMU * a.q[(size_t)j * nl + r] * (a.press[r] * a.pfct) / kKB / T;
      dens[(size_t)r * ns + j] = rho * mol.mass[j];
    }
  BCUDA(cudaMemcpyAsync(b->d_density, dens.data(), dens.size() * 8, cudaMemcpyHostToDevice, s));
  // per-isotope factors at this temperature: pass 1 (extinction.c:412-418) uses
  // ratio*SIGCTE*...*/mass/Z per line, pass 2 (464) SIGCTE*ratio/(mass*Z) per group
  std::vector<double> facfull(b->niso), fac2(b->niso);
  for (int i = 0; i < b->niso; i++) {
    const double Z = b->ziso[(size_t)i * b->ntemp + it];
    facfull[i] = t.iso_ratio[i] * kSIGCTE / t.iso_mass[i] / Z;
    fac2[i] = kSIGCTE * t.iso_ratio[i] / (t.iso_mass[i] * Z);
  }
  double *d_facfull = dev_upload(facfull), *d_fac2 = dev_upload(fac2);
  BCUDA(cudaMemsetAsync(b->d_kmax, 0, kMaxGridMol * 8, s));
  if (b->nlines > 0) {
    PhaseTimer pt(b, "kmax", s);
    kmax_kernel<<<148 * 4, 256, 0, s>>>(b->d_wavn, b->d_elow, b->d_gf, b->d_isoid, b->d_inrange, b->nlines,
                                        T, d_facfull, d_iso_gmol, (unsigned long long *)b->d_kmax, b->ngmol);
    BCUDA(cudaGetLastError());
  }
  if (b->ngroups > 0) {
    PhaseTimer pt(b, "strength", s);
    strength_kernel<<<(unsigned)((b->ngroups + 255) / 256), 256, 0, s>>>(
        b->d_gstart, b->d_giso, b->d_c_wavn, b->d_c_elow, b->d_c_gf, b->ngroups, T, d_fac2, b->d_gS);
    BCUDA(cudaGetLastError());
  }
  CellIso *cells = (CellIso *)b->d_cellinfo;
  {
  PhaseTimer pt(b, "widths", s);
  widths_kernel<<<nl, std::max(32, b->niso), 0, s>>>(
      cells, nl, b->niso, ns, T, b->d_density, d_spec_mass, d_spec_radius, d_iso_mass, d_iso_spec,
      d_iso_gmol, b->d_aDop, b->d_aLor, b->nDop, b->nLor, b->d_prof_size, b->osamp, b->wn_lo,
      d_iso_gbeg, b->d_gwavn, b->d_gS, b->d_kmax, o.ethreshold);
  BCUDA(cudaGetLastError());
  }
  PhaseTimer pt(b, "accumulate", s);
  BCUDA(cudaMemsetAsync(b->d_out, 0, (size_t)nl * b->ngmol * b->nwave * 8, s));
  dim3 grid((b->nwave + kAccThreads - 1) / kAccThreads, nl);
  accumulate_kernel<<<grid, kAccThreads, 0, s>>>(
      cells, b->niso, b->ngmol, b->nwave, b->osamp, d_iso_gmol, d_iso_gbeg, b->d_giown, b->d_gidwn,
      b->d_gwavn, b->d_gS, b->d_kmax, o.ethreshold, b->d_aDop, b->nDop, b->nLor, b->d_prof_off,
      b->d_prof_size, b->d_prof, b->d_out, d_neval);
  BCUDA(cudaGetLastError());
  BCUDA(cudaStreamSynchronize(s));
  cudaFree(d_fac2); cudaFree(d_facfull);
}

static void ensure_builder(BuilderState *&b, const Options &o, const Atmosphere &a,
                           const Molecules &m, Tli &t, const std::vector<double> &wn, cudaStream_t s) {
  if (!b) b = new BuilderState();
  if (!t.present) fail("the opacity-grid builder needs a TLI line list (linedb)");
  if (b->nwave == 0) setup_static(b, o, a, m, t, wn);
  build_profiles(b, o, s);
  load_lines(b, o, t, wn, s);
  if (!b->d_density) {
    BCUDA(cudaMalloc((void **)&b->d_density, (size_t)b->nlayer * b->nspec * 8));
    BCUDA(cudaMalloc((void **)&b->d_kmax, kMaxGridMol * 8));
    BCUDA(cudaMalloc((void **)&b->d_out, (size_t)b->nlayer * b->ngmol * b->nwave * 8));
    BCUDA(cudaMalloc((void **)&b->d_cellinfo, (size_t)b->nlayer * b->niso * sizeof(CellIso)));
  }
}

void builder_slice(BuilderState *&b, const Options &o, const Atmosphere &a, const Molecules &m,
                   Tli &t, const std::vector<double> &wn, cudaStream_t s, int t_begin, int t_end,
                   double *host_out) {
  ensure_builder(b, o, a, m, t, wn, s);
  if (t_begin < 0 || t_end > b->ntemp || t_begin > t_end)
    fail("temperature slice [%d, %d) outside the grid of %d temperatures", t_begin, t_end, b->ntemp);
  int *d_iso_spec = dev_upload(b->iso_spec), *d_iso_gmol = dev_upload(b->iso_gmol);
  double *d_iso_mass = dev_upload(t.iso_mass), *d_spec_mass = dev_upload(m.mass),
         *d_spec_radius = dev_upload(m.radius_cm);
  long long *d_iso_gbeg = dev_upload(b->iso_gbeg);
  unsigned long long *d_neval = nullptr;
  BCUDA(cudaMalloc((void **)&d_neval, 8));
  BCUDA(cudaMemset(d_neval, 0, 8));
  const int nt = t_end - t_begin, nl = b->nlayer;
  const size_t plane = (size_t)b->ngmol * b->nwave;
  std::vector<double> tmp((size_t)nl * plane);
  for (int it = t_begin; it < t_end; it++) {
    build_temperature(b, o, a, m, t, it, s, d_iso_spec, d_iso_gmol, d_iso_mass, d_spec_mass,
                      d_spec_radius, d_iso_gbeg, d_neval);
    auto td0 = std::chrono::steady_clock::now();
    BCUDA(cudaMemcpy(tmp.data(), b->d_out, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (int r = 0; r < nl; r++)
      memcpy(host_out + ((size_t)r * nt + (it - t_begin)) * plane, tmp.data() + (size_t)r * plane, plane * 8);
    b->phase_ms["d2h"] += std::chrono::duration<double, std::milli>(
        std::chrono::steady_clock::now() - td0).count();
  }
  unsigned long long ne = 0;
  BCUDA(cudaMemcpy(&ne, d_neval, 8, cudaMemcpyDeviceToHost));
  b->neval += (long long)ne;
  cudaFree(d_iso_spec); cudaFree(d_iso_gmol); cudaFree(d_iso_mass); cudaFree(d_spec_mass);
  cudaFree(d_spec_radius); cudaFree(d_iso_gbeg); cudaFree(d_neval);
}

void builder_run_and_write(BuilderState *&b, const Options &o, const Atmosphere &a,
                           const Molecules &m, Tli &t, const std::vector<double> &wn,
                           cudaStream_t s, const std::string &path) {
  ensure_builder(b, o, a, m, t, wn, s);
  OpacityGrid g;
  g.nmol = b->ngmol; g.ntemp = b->ntemp; g.nlayer = b->nlayer; g.nwave = b->nwave;
  g.molid = b->gmol_id; g.temp = b->temps; g.wn = wn;
  g.press.resize(b->nlayer);
  // opacity.c:344-346.  The reference stores the pressures after its identity spline resample
  // (makesample.c:507-531, compiled -ffast-math): the top layer differs by ~1 ulp from the file value.
  for (int r = 0; r < b->nlayer; r++) g.press[r] = a.press[r] * a.pfct;
  // $BART_TSLICE="begin:end" builds only a slice of the temperature axis and writes its planes
  // in place (one process per GPU shards the axis; rank 0 writes the header first)
  int t0 = 0, t1 = b->ntemp;
  bool header = true;
  if (const char *e = getenv("BART_TSLICE")) {
    if (sscanf(e, "%d:%d", &t0, &t1) != 2) fail("BART_TSLICE must be 'begin:end'");
    header = t0 == 0;
  }
  const size_t plane = (size_t)b->ngmol * b->nwave;
  std::vector<double> slab((size_t)b->nlayer * (t1 - t0) * plane);
  builder_slice(b, o, a, m, t, wn, s, t0, t1, slab.data());
  if (t0 == 0 && t1 == b->ntemp) { write_opacity_file(path, g, slab.data()); return; }
  int fd = open(path.c_str(), O_WRONLY | O_CREAT, 0644);
  if (fd < 0) fail("Opacity filename '%s' cannot be opened for writing.", path.c_str());
  const long long hdr = 4 * sizeof(long) + g.nmol * sizeof(int) + (g.ntemp + g.nlayer + g.nwave) * 8;
  if (header) {
    std::vector<char> h(hdr);
    char *p = h.data();
    long dims[4] = {g.nmol, g.ntemp, g.nlayer, g.nwave};
    memcpy(p, dims, sizeof(dims)); p += sizeof(dims);
    memcpy(p, g.molid.data(), g.nmol * sizeof(int)); p += g.nmol * sizeof(int);
    memcpy(p, g.temp.data(), g.ntemp * 8); p += g.ntemp * 8;
    memcpy(p, g.press.data(), g.nlayer * 8); p += g.nlayer * 8;
    memcpy(p, g.wn.data(), g.nwave * 8);
    if (pwrite(fd, h.data(), hdr, 0) != hdr) fail("short write on '%s'", path.c_str());
  }
  const int nt = t1 - t0;
  for (int r = 0; r < b->nlayer; r++) {
    const long long off = hdr + ((long long)r * g.ntemp + t0) * (long long)plane * 8;
    const long long len = (long long)nt * plane * 8;
    if (pwrite(fd, slab.data() + (size_t)r * nt * plane, len, off) != len)
      fail("short write on '%s'", path.c_str());
  }
  close(fd);
}

long long builder_stats(BuilderState *b, long long *nlines, long long *ngroups, long long *neval) {
  if (!b) return -1;
  if (nlines) *nlines = b->nlines;
  if (ngroups) *ngroups = b->ngroups;
  if (neval) *neval = b->neval;
  return b->nlines;
}

double builder_phase_ms(BuilderState *b, const char *name) {
  if (!b) return -1.0;
  auto it = b->phase_ms.find(name);
  return it == b->phase_ms.end() ? 0.0 : it->second;
}

long long builder_line_bins(BuilderState *b, long long *iown_out, long long capacity) {
  const long long n = (long long)b->h_iown.size();
  for (long long i = 0; i < n && i < capacity; i++) iown_out[i] = b->h_iown[i];
  return n;
}

int builder_profile(BuilderState *b, int idop, int ilor, float *out, long long capacity,
                    long long *halfsize) {
  if (!b->profiles_ready) fail("Voigt profiles have not been computed");
  if (idop < 0 || idop >= b->nDop || ilor < 0 || ilor >= b->nLor) fail("profile index out of range");
  const size_t p = (size_t)idop * b->nLor + ilor;
  const long long n = 2 * b->prof_size[p] + 1;
  if (halfsize) *halfsize = b->prof_size[p];
  if (out) {
    if (capacity < n) fail("profile buffer too small (%lld needed)", n);
    BCUDA(cudaMemcpy(out, b->d_prof + b->prof_off[p], n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return 0;
}

void builder_free(BuilderState *b) {
  if (!b) return;
  void *ptrs[] = {b->d_wl, b->d_elow, b->d_gf, b->d_wavn, b->d_c_wavn, b->d_c_elow, b->d_c_gf,
                  b->d_isoid, b->d_iown, b->d_idwn, b->d_inrange, b->d_gstart, b->d_giown,
                  b->d_gidwn, b->d_giso, b->d_gwavn, b->d_gS, b->d_prof, b->d_aDop, b->d_aLor,
                  b->d_prof_off, b->d_prof_size, b->d_density, b->d_kmax, b->d_out, b->d_cellinfo};
  for (void *p : ptrs) if (p) cudaFree(p);
  delete b;
}

}  // namespace bart
